"""overlap_probe.py -- can the HBM-bound layers of one half-plane hide under the tensor-bound layers of the other?
Two contexts on ONE device, each limited to half the SMs (w2x_debug_set_num_sms), each converting its own 4096x2048
half, phase-shifted by a preliminary partial pass on the second context; against one context on all SMs converting the
whole 4096x4096 plane.  Device-resident buffers, K back-to-back passes, wall clock around a full drain (run under gpurun)."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import w2x_loader  # noqa: E402
from oracle import oracle  # noqa: E402  (model fixture only)

w2x = w2x_loader.load()
om = oracle.OracleModel.golden("scale2.0x")
m = w2x.Model.from_arrays(om.weights, om.biases)
K = 10
W = 4096
dev = torch.device("cuda:0")


def buf(h):
    return torch.rand((h, W), device=dev, dtype=torch.float32), torch.empty((h, W), device=dev, dtype=torch.float32)


def run(ctxs, hs, pre_rows, sms):
    bufs = [buf(h) for h in hs]
    pre = buf(pre_rows) if pre_rows else None
    for c in ctxs:
        c.debug_set_num_sms(sms)
    for c, (a, b), h in zip(ctxs, bufs, hs):            # warm-up
        c.convert_plane_device(m, a.data_ptr(), W, h, W * 4, b.data_ptr(), W * 4)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    if pre is not None:                                  # phase shift: the LAST context starts with a partial pass
        ctxs[-1].convert_plane_device(m, pre[0].data_ptr(), W, pre_rows, W * 4, pre[1].data_ptr(), W * 4)
    for _ in range(K):
        for c, (a, b), h in zip(ctxs, bufs, hs):
            c.convert_plane_device(m, a.data_ptr(), W, h, W * 4, b.data_ptr(), W * 4)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    px = K * sum(hs) * W + (pre_rows * W if pre_rows else 0)
    return px / dt / 1e6, dt * 1e3 / K


one = w2x.Context(0, engine=w2x.ENGINE_TC)
print("1 context, all SMs, 4096x4096:        %.1f Mpix/s  (%.2f ms per pass)" % run([one], [4096], 0, 0), flush=True)
print("1 context, all SMs, 2 x 4096x2048:    %.1f Mpix/s  (%.2f ms per pair)" % run([one, one], [2048, 2048], 0, 0), flush=True)
a, b = w2x.Context(0, engine=w2x.ENGINE_TC), w2x.Context(0, engine=w2x.ENGINE_TC)
for pre in (0, 512, 1024, 1536):
    for sms in (74, 0):
        r = run([a, b], [2048, 2048], pre, sms)
        print("2 contexts, %3s SMs each, phase shift %4d rows: %.1f Mpix/s  (%.2f ms per pair of halves)" % (sms or "all", pre, r[0], r[1]), flush=True)
