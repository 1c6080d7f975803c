"""reference_lib.py -- TEST INFRASTRUCTURE ONLY.

ctypes access to oracle/_ref/libw2x_reference.so: the reference's OWN hot-path sources
(/root/reference/src/modelHandler.cpp, convertRoutine.cpp) compiled against the OpenCV API shim in
oracle/cvshim (recipe: oracle/Makefile; built only where /root/reference exists, the prebuilt file
travels to the GPU box).  What runs here is the reference's real control flow -- picojson model loading,
thread partition, layer loop, padding, block split / crop / stitch -- over the shim's restated fp32
arithmetic; oracle/ref_cv2.py covers the complementary half (real OpenCV arithmetic, restated flow)."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(_HERE, "_ref", "libw2x_reference.so")
_lib = None


def available() -> bool:
    return os.path.exists(SO)


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise FileNotFoundError(SO + " (built by `make -C oracle` where /root/reference is present)")
        L = C.CDLL(SO)
        fp = C.POINTER(C.c_float)
        L.w2xr_load.argtypes = [C.c_char_p, C.POINTER(C.c_void_p)]
        L.w2xr_free.argtypes = [C.c_void_p]
        L.w2xr_free.restype = None
        L.w2xr_layer_dims.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.w2xr_config.argtypes = [C.c_int, C.c_int]
        L.w2xr_convert.argtypes = [C.c_void_p, fp, C.c_int, C.c_int, C.c_long, fp, C.c_long, C.c_int]
        L.w2xr_filter.argtypes = [C.c_void_p, C.c_int, fp, C.c_int, C.c_int, fp]
        L.w2xr_convert_log.argtypes = [C.c_void_p, fp, C.c_int, C.c_int, C.c_long, fp, C.c_long, C.c_int, C.c_char_p, C.c_int]
        _lib = L
    return _lib


def configure(n_job: int = 4, block_exp: int = 9):
    """modelUtility::setNumberOfJobs / setBlockSizeExp2Square (process-wide singleton, as in the reference)."""
    if lib().w2xr_config(n_job, block_exp) != 0:
        raise ValueError("modelUtility rejected the configuration")


class ReferenceModels:
    """std::vector<std::unique_ptr<w2xc::Model>> filled by modelUtility::generateModelFromJSON."""

    def __init__(self, json_path: str):
        h = C.c_void_p()
        n = lib().w2xr_load(os.fsencode(json_path), C.byref(h))
        if n < 0:
            raise RuntimeError("generateModelFromJSON failed for " + json_path)
        self._h, self.n = h, n

    def close(self):
        if self._h:
            lib().w2xr_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def dims(self):
        out = []
        for i in range(self.n):
            a, b = C.c_int(), C.c_int()
            lib().w2xr_layer_dims(self._h, i, C.byref(a), C.byref(b))
            out.append((a.value, b.value))
        return out

    def convert(self, plane, block_splitting=True):
        """w2xc::convertWithModels(inputPlane, outputPlane, models, blockSplitting)"""
        x = np.ascontiguousarray(plane, np.float32)
        h, w = x.shape
        out = np.empty((h, w), np.float32)
        fp = C.POINTER(C.c_float)
        rc = lib().w2xr_convert(self._h, x.ctypes.data_as(fp), w, h, w, out.ctypes.data_as(fp), w, int(bool(block_splitting)))
        if rc != 0:
            raise RuntimeError(f"convertWithModels failed (rc={rc})")
        return out

    def convert_with_log(self, plane, block_splitting=True, cap=1 << 20):
        """convertWithModels plus the text the reference printed to std::cout while doing it (layer and block progress)."""
        x = np.ascontiguousarray(plane, np.float32)
        h, w = x.shape
        out = np.empty((h, w), np.float32)
        fp = C.POINTER(C.c_float)
        buf = C.create_string_buffer(cap)
        rc = lib().w2xr_convert_log(self._h, x.ctypes.data_as(fp), w, h, w, out.ctypes.data_as(fp), w, int(bool(block_splitting)), buf, cap)
        if rc != 0:
            raise RuntimeError(f"convertWithModels failed (rc={rc})")
        return out, buf.value.decode()

    def filter(self, layer, in_planes):
        """w2xc::Model::filter of one layer on planar [n_in][h][w] input"""
        x = np.ascontiguousarray(in_planes, np.float32)
        n_in, n_out = self.dims[layer]
        assert x.ndim == 3 and x.shape[0] == n_in
        out = np.empty((n_out, x.shape[1], x.shape[2]), np.float32)
        fp = C.POINTER(C.c_float)
        rc = lib().w2xr_filter(self._h, layer, x.ctypes.data_as(fp), x.shape[2], x.shape[1], out.ctypes.data_as(fp))
        if rc != 0:
            raise RuntimeError(f"Model::filter failed (rc={rc})")
        return out
