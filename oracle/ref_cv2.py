"""ref_cv2.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

The reference's hot path re-stated call for call on the reference's own arithmetic
backend (OpenCV, through the Python ``cv2`` module), because the reference C++
cannot be compiled in this image (no OpenCV C++ headers/libraries).

Every cv2 call below is the Python spelling of the cv:: call at the cited line of
/root/reference/src (WL-Amigo/waifu2x-converter-cpp):

  Model.filter / Model._filter_worker  <- src/modelHandler.cpp:26-72, :117-159
  convert_with_models_basic            <- src/convertRoutine.cpp:53-82
  convert_with_models_block_split      <- src/convertRoutine.cpp:84-169
  convert_with_models                  <- src/convertRoutine.cpp:21-51
  load_models_json                     <- src/modelHandler.cpp:74-115, :170-197,
                                          src/modelHandler.hpp:48-71

It is used (a) to generate the golden vectors under tests/golden/ (the reference
itself ships none), (b) by tests to pin oracle/w2x_oracle.c, and (c) by
``bench.py --impl reference`` / ``cpu_baseline`` as the reference's CPU path timed
on the host cores.  The product never imports it.
"""
from __future__ import annotations

import json
import math
import threading

import numpy as np

try:  # cv2 is present in the build image; keep import errors readable elsewhere
    import cv2
except Exception as e:  # pragma: no cover
    cv2 = None
    _cv2_err = e


def _need_cv2():
    if cv2 is None:  # pragma: no cover
        raise RuntimeError(f"cv2 is not importable: {_cv2_err}")


class ModelUtility:
    """modelUtility singleton (src/modelHandler.hpp:92-113): nJob=4, block 512x512."""
    n_job = 4
    block_w = 512
    block_h = 512


class Model:
    """One conv layer (class Model, src/modelHandler.hpp:24-90)."""

    def __init__(self, obj: dict):
        self.n_in = int(obj["nInputPlane"])           # hpp:50
        self.n_out = int(obj["nOutputPlane"])          # hpp:51
        self.k = int(obj["kW"])
        if self.k != int(obj["kH"]):                   # hpp:52-58 -> exit(-1)
            raise SystemExit(-1)
        w = np.asarray(obj["weight"], dtype=np.float64)        # [o][i][ky][kx]
        assert w.shape == (self.n_out, self.n_in, self.k, self.k)
        self.weights = w.astype(np.float32)            # at<float>() = double, cpp:96-97
        self.biases = np.asarray(obj["bias"], dtype=np.float64)  # kept double, cpp:111

    # Model::filterWorker, src/modelHandler.cpp:117-159
    def _filter_worker(self, in_planes, out_planes, begin, n_works):
        _need_cv2()
        cv2.ocl.setUseOpenCL(False)                    # :121
        size = in_planes[0].shape
        for op in range(begin, begin + n_works):       # :127
            acc = np.zeros(size, np.float32)           # :131-132
            for ip in range(self.n_in):                # :134
                tmp = cv2.filter2D(in_planes[ip], -1, self.weights[op, ip], anchor=(-1, -1),
                                   delta=0.0, borderType=cv2.BORDER_REPLICATE)   # :141-142
                acc = cv2.add(acc, tmp)                # :144
            acc = cv2.add(acc, float(self.biases[op]))  # :147 (double scalar on CV_32F)
            more = cv2.max(acc, 0.0)                   # :150
            less = cv2.min(acc, 0.0)                   # :151
            acc = cv2.scaleAdd(less, 0.1, more)        # :152
            out_planes[op] = acc.reshape(size).astype(np.float32, copy=True)   # :153-154

    # Model::filter, src/modelHandler.cpp:26-72
    def filter(self, in_planes, n_job=None):
        if len(in_planes) != self.n_in:                # :29-35 -> false
            return None
        n_job = ModelUtility.n_job if n_job is None else n_job
        out_planes = [None] * self.n_out
        wpt = self.n_out // n_job                      # :47
        threads = []
        for idx in range(n_job):                       # :48-65
            if not (idx == n_job - 1 and wpt * n_job != self.n_out):
                args = (in_planes, out_planes, wpt * idx, wpt)
            else:
                args = (in_planes, out_planes, wpt * idx, self.n_out - wpt * idx)
            t = threading.Thread(target=self._filter_worker, args=args)
            t.start()
            threads.append(t)
        for t in threads:                              # :67-69
            t.join()
        return out_planes


def load_models_json(path: str):
    """modelUtility::generateModelFromJSON (src/modelHandler.cpp:170-197).
    Python's float() is correctly rounded like strtod (include/picojson.h:788)."""
    with open(path, "r") as f:
        root = json.load(f)
    return [Model(o) for o in root]


def convert_with_models_basic(in_plane, models, n_job=None, log=None):
    planes = [np.ascontiguousarray(in_plane, dtype=np.float32)]   # ROI -> the data read is identical
    for index, m in enumerate(models):                 # :66
        if log is not None:
            log(f"Iteration #{index + 1}...")          # :67
        planes = m.filter(planes, n_job)               # :68
        if planes is None:
            raise SystemExit(-1)                       # :69
    return planes[0].copy()                            # :78


def block_table(w, h, bw, bh, n_model):
    """Index arithmetic of convertWithModelsBlockSplit, same row layout as
    w2xo_block_table: (r, c, in_y0, in_y1, in_x0, in_x1, out_y0, out_x0)."""
    sc = int(math.ceil(np.float32(w) / np.float32(bw - 2 * n_model)))    # :100-102 (float math)
    sr = int(math.ceil(np.float32(h) / np.float32(bh - 2 * n_model)))    # :103-105
    pw, ph = w + 2 * n_model, h + 2 * n_model
    rows = []
    for r in range(sr):
        y0 = r * (bh - 2 * n_model)
        y1 = ph if r == sr - 1 else y0 + bh
        for c in range(sc):
            x0 = c * (bw - 2 * n_model)
            x1 = pw if c == sc - 1 else x0 + bw
            rows.append((r, c, y0, y1, x0, x1, r * (bh - 2 * n_model), c * (bh - 2 * n_model)))
    return rows, sc, sr


def convert_with_models_block_split(in_plane, models, n_job=None, log=None):
    _need_cv2()
    n = len(models)
    h, w = in_plane.shape
    bw, bh = ModelUtility.block_w, ModelUtility.block_h
    temp = cv2.copyMakeBorder(in_plane, n, n, n, n, cv2.BORDER_REPLICATE)   # :96
    rows, _, _ = block_table(w, h, bw, bh, n)
    out = np.zeros((h, w), np.float32)                                      # :113
    for (r, c, y0, y1, x0, x1, oy, ox) in rows:
        if log is not None:
            log(f"start process block ({c},{r}) ...")                       # :133-134
        block_out = convert_with_models_basic(temp[y0:y1, x0:x1], models, n_job, log)   # :135
        src = block_out[n:block_out.shape[0] - n, n:block_out.shape[1] - n]  # :143-147
        out[oy:oy + src.shape[0], ox:ox + src.shape[1]] = src               # :148-161
    return out


def convert_with_models(in_plane, models, block_splitting=True, n_job=None, log=None):
    _need_cv2()
    in_plane = np.asarray(in_plane, dtype=np.float32)
    h, w = in_plane.shape
    bw, bh = ModelUtility.block_w, ModelUtility.block_h
    require = (w * h) > (bw * bh * 3) // 2                                  # :25-26
    if block_splitting and require:
        return convert_with_models_block_split(in_plane, models, n_job, log)
    n = len(models)
    temp = cv2.copyMakeBorder(in_plane, n, n, n, n, cv2.BORDER_REPLICATE)   # :35
    full = convert_with_models_basic(temp, models, n_job, log)              # :38
    return full[n:h + n, n:w + n].copy()                                    # :40-46
