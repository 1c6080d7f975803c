"""strip_check.py -- row-strip kernel vs the 16x16-tile kernel vs the CPU oracle on a few sizes, then per-layer times of both
at 4096x4096 (run under gpurun):  python tools/strip_check.py [size]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import w2x_loader  # noqa: E402
from oracle import oracle  # noqa: E402

w2x = w2x_loader.load()
om = oracle.OracleModel.golden("scale2.0x")
m = w2x.Model.from_arrays(om.weights, om.biases)
for prec, pname in ((w2x.PRECISION_F16_F8X2, "f16+f8x2"), (w2x.PRECISION_F16X3, "f16x3")):
    ctx = w2x.Context(0, engine=w2x.ENGINE_TC)
    ctx.set_precision(prec)
    for (w, h) in ((40, 30), (130, 70), (300, 200)):
        x = oracle.seeded_plane(w, h, 3, "uniform")
        ref = om.convert(x, n_job=os.cpu_count() or 4)
        ctx.debug_set_strip(True)
        ys = ctx.convert_plane(m, x)
        ctx.debug_set_strip(False)
        yt = ctx.convert_plane(m, x)
        print(f"{pname} {w}x{h}: strip vs oracle {np.abs(ys - ref).max():.2e}  tile vs oracle {np.abs(yt - ref).max():.2e}  strip vs tile {np.abs(ys - yt).max():.2e}", flush=True)
    size = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    x = oracle.seeded_plane(size, size, 1, "uniform")
    for strip in (True, False):
        ctx.debug_set_strip(strip)
        ctx.debug_set_host_bands(1)
        ctx.convert_plane(m, x)
        ctx.set_timing(True)
        for _ in range(3):
            ctx.convert_plane(m, x)
        t = ctx.layer_times()
        ctx.set_timing(False)
        print(f"{pname} {size}^2 strip={strip}: per-layer ms", [round(a / max(b, 1), 3) for a, b, _ in t], "sum", round(sum(a / max(b, 1) for a, b, _ in t), 3), [n for _, _, n in t][1:4], flush=True)
    ctx.close()
