"""parity_report.py -- max-abs error of every engine / precision against the reference goldens and, at the BASELINE sizes,
of the default precision over the WHOLE plane against the fp32 engine (run under gpurun):

    python tools/parity_report.py > gpurun_out/parity_report.txt
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import w2x_loader  # noqa: E402
from oracle import oracle  # noqa: E402

w2x = w2x_loader.load()
ctxs = {"fp32": w2x.Context(0, engine=w2x.ENGINE_FP32), "tc f16x3": w2x.Context(0, engine=w2x.ENGINE_TC), "tc f16+f8x2": w2x.Context(0, engine=w2x.ENGINE_TC)}
ctxs["tc f16x3"].set_precision(w2x.PRECISION_F16X3)
ctxs["tc f16+f8x2"].set_precision(w2x.PRECISION_F16_F8X2)
print("max-abs error vs the reference's OpenCV output (tests/golden/cfg1_*.npy, 256x256), gate 1e-4")
for name, kind in (("scale2.0x", "uniform"), ("scale2.0x", "smooth"), ("noise1", "uniform"), ("noise2", "uniform")):
    om = oracle.OracleModel.golden(name)
    m = w2x.Model.from_arrays(om.weights, om.biases)
    x = oracle.seeded_plane(256, 256, 0, kind)
    g = np.load(os.path.join(ROOT, "tests", "golden", f"cfg1_{name}_{kind}.npy"))
    print(f"  {name:10s} {kind:8s} " + "  ".join(f"{k}: {np.abs(c.convert_plane(m, x) - g).max():.2e}" for k, c in ctxs.items()))
print("whole plane, white noise (default_rng), max-abs vs the fp32 engine (pinned to the oracle at 5e-6):")
for size, seed in ((2048, 7), (4096, 1), (8192, 2)):
    x = oracle.seeded_plane(size, size, seed, "uniform")
    for name in oracle.MODEL_NAMES:
        om = oracle.OracleModel.golden(name)
        m = w2x.Model.from_arrays(om.weights, om.biases)
        ref = ctxs["fp32"].convert_plane(m, x)
        print(f"  {size}x{size} {name:10s} " + "  ".join(f"{k}: {np.abs(c.convert_plane(m, x) - ref).max():.2e}" for k, c in ctxs.items() if k != "fp32"), flush=True)
