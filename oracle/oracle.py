"""oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

ctypes front end of oracle/w2x_oracle.c (the plain-C restatement of the reference's
Model::filter / convertWithModels, see that file's header for the file:line map) plus
the small model-fixture helpers the tests share.  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import this module.
"""
from __future__ import annotations

import ctypes as C
import json
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "libw2x_oracle.so")

MODEL_NAMES = ("scale2.0x", "noise1", "noise2")
GOLDEN_DIR = os.path.join(os.path.dirname(_HERE), "tests", "golden")


class _Layer(C.Structure):
    _fields_ = [("n_in", C.c_int), ("n_out", C.c_int), ("k", C.c_int),
                ("w", C.POINTER(C.c_float)), ("b", C.POINTER(C.c_double))]


def build(force: bool = False) -> str:
    """Compile oracle/w2x_oracle.c -> oracle/_ref/libw2x_oracle.so (gcc only) and, where the reference tree is present
    (the authoring container), the reference's own hot-path sources against oracle/cvshim -> oracle/_ref/libw2x_reference.so
    (see oracle/reference_lib.py; on the GPU box the prebuilt file is used as it is)."""
    src = os.path.join(_HERE, "w2x_oracle.c")
    ref_so = os.path.join(_HERE, "_ref", "libw2x_reference.so")
    have_ref = os.path.exists("/root/reference/src/modelHandler.cpp")
    ref_missing = have_ref and not os.path.exists(ref_so)
    if force or ref_missing or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "all"])
    elif have_ref:
        subprocess.check_call(["make", "-s", "-C", _HERE, "all"])      # dependency-driven: e.g. the option-A binary after a product rebuild
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = C.CDLL(_SO)
        fp, ip = C.POINTER(C.c_float), C.POINTER(C.c_int)
        L.w2xo_filter.argtypes = [C.POINTER(_Layer), fp, fp, C.c_int, C.c_int, C.c_int]
        L.w2xo_filter.restype = C.c_int
        L.w2xo_convert.argtypes = [C.POINTER(_Layer), C.c_int, fp, C.c_int, C.c_int, C.c_long, fp,
                                   C.c_int, C.c_int, C.c_int, C.c_int]
        L.w2xo_convert.restype = C.c_int
        L.w2xo_convert_basic.argtypes = [C.POINTER(_Layer), C.c_int, fp, C.c_int, C.c_int, C.c_long,
                                         fp, C.c_int]
        L.w2xo_convert_basic.restype = C.c_int
        L.w2xo_block_table.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, ip, ip, ip]
        L.w2xo_block_table.restype = C.c_int
        L.w2xo_pad_replicate.argtypes = [fp, C.c_int, C.c_int, C.c_long, C.c_int, fp]
        L.w2xo_pad_replicate.restype = None
        _lib = L
    return _lib


class OracleModel:
    """A 7-layer model as the oracle sees it: fp32 weights [o][i][3][3], fp64 biases."""

    def __init__(self, weights, biases):
        self.weights = [np.ascontiguousarray(w, np.float32) for w in weights]
        self.biases = [np.ascontiguousarray(b, np.float64) for b in biases]
        self._arr = (_Layer * len(self.weights))()
        for i, (w, b) in enumerate(zip(self.weights, self.biases)):
            assert w.ndim == 4 and w.shape[2] == w.shape[3] == 3 and b.shape == (w.shape[0],)
            self._arr[i] = _Layer(w.shape[1], w.shape[0], 3,
                                  w.ctypes.data_as(C.POINTER(C.c_float)),
                                  b.ctypes.data_as(C.POINTER(C.c_double)))

    def __len__(self):
        return len(self.weights)

    @property
    def dims(self):
        return [(w.shape[1], w.shape[0]) for w in self.weights]

    # --- construction ------------------------------------------------------
    @classmethod
    def from_json(cls, path):
        """Reference JSON format (src/modelHandler.cpp:74-115): double -> float weights."""
        with open(path) as f:
            root = json.load(f)
        ws = [np.asarray(o["weight"], np.float64).astype(np.float32) for o in root]
        bs = [np.asarray(o["bias"], np.float64) for o in root]
        return cls(ws, bs)

    @classmethod
    def from_npz(cls, path):
        z = np.load(path)
        n = int(z["n_layers"])
        return cls([z[f"w{i}"] for i in range(n)], [z[f"b{i}"] for i in range(n)])

    @classmethod
    def golden(cls, name):
        return cls.from_npz(os.path.join(GOLDEN_DIR, "models", f"{name}_model.npz"))

    @classmethod
    def random(cls, dims, seed=0, scale=None):
        rng = np.random.default_rng(seed)
        ws, bs = [], []
        for (ci, co) in dims:
            s = scale if scale is not None else 1.0 / np.sqrt(9.0 * ci)
            ws.append((rng.standard_normal((co, ci, 3, 3)) * s).astype(np.float32))
            bs.append((rng.standard_normal(co) * 0.05).astype(np.float32).astype(np.float64))
        return cls(ws, bs)

    def save_npz(self, path):
        d = {"n_layers": np.int32(len(self))}
        for i, (w, b) in enumerate(zip(self.weights, self.biases)):
            d[f"w{i}"] = w
            d[f"b{i}"] = b
        np.savez_compressed(path, **d)

    def write_json(self, path):
        """Write the reference's JSON model format (keys and nesting of models/*.json).
        Weights are written as repr(float64(float32)) so a strtod-equivalent loader
        followed by double->float gives back exactly these fp32 values."""
        root = []
        for w, b in zip(self.weights, self.biases):
            root.append({"weight": w.astype(np.float64).tolist(), "nOutputPlane": int(w.shape[0]),
                         "kW": 3, "kH": 3, "bias": b.tolist(), "nInputPlane": int(w.shape[1])})
        with open(path, "w") as f:
            json.dump(root, f, separators=(",", ":"))

    # --- the restated reference functions ---------------------------------
    def filter(self, layer, in_planes, n_job=4):
        """Model::filter on dense planar input [n_in][h][w] -> [n_out][h][w]."""
        x = np.ascontiguousarray(in_planes, np.float32)
        ci, co = self.dims[layer]
        assert x.ndim == 3 and x.shape[0] == ci
        out = np.empty((co, x.shape[1], x.shape[2]), np.float32)
        fp = C.POINTER(C.c_float)
        rc = lib().w2xo_filter(C.byref(self._arr[layer]), x.ctypes.data_as(fp), out.ctypes.data_as(fp),
                               x.shape[2], x.shape[1], n_job)
        if rc != 0:
            raise RuntimeError(f"w2xo_filter rc={rc}")
        return out

    def convert(self, plane, block_splitting=True, block=(512, 512), n_job=4):
        """convertWithModels on one fp32 plane (may be a strided view along rows)."""
        x = np.asarray(plane, np.float32)
        if x.strides[1] != 4 or x.strides[0] % 4:
            x = np.ascontiguousarray(x)
        h, w = x.shape
        out = np.empty((h, w), np.float32)
        fp = C.POINTER(C.c_float)
        rc = lib().w2xo_convert(self._arr, len(self), C.cast(x.ctypes.data, fp), w, h, x.strides[0] // 4,
                                out.ctypes.data_as(fp), int(bool(block_splitting)), block[0], block[1],
                                n_job)
        if rc != 0:
            raise RuntimeError(f"w2xo_convert rc={rc}")
        return out


def block_table(w, h, bw=512, bh=512, n_model=7):
    n = lib().w2xo_block_table(w, h, bw, bh, n_model, None, None, None)
    tab = np.zeros((n, 8), np.int32)
    sc, sr = C.c_int(), C.c_int()
    lib().w2xo_block_table(w, h, bw, bh, n_model, tab.ctypes.data_as(C.POINTER(C.c_int)),
                           C.byref(sc), C.byref(sr))
    return tab, sc.value, sr.value


def seeded_plane(w, h, seed, kind="uniform"):
    """The synthetic Y planes SURVEY.md section 8(d) names: uniform [0,1) noise (the hard case for
    the tolerance gate) or a smooth u8/255 image."""
    rng = np.random.default_rng(seed)
    if kind == "uniform":
        return rng.random((h, w), dtype=np.float32)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    img = 0.5 + 0.25 * np.sin(xx / 9.0 + seed) * np.cos(yy / 7.0) + 0.2 * np.sin((xx + yy) / 23.0)
    img += 0.02 * rng.standard_normal((h, w))
    return (np.clip(np.round(img * 255.0), 0, 255) / 255.0).astype(np.float32)
