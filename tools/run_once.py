"""run_once.py -- N whole-plane passes of the hot path (for ncu captures): python tools/run_once.py [size] [passes] [engine]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import w2x_loader  # noqa: E402
from oracle import oracle  # noqa: E402  (model fixture + synthetic plane only)

w2x = w2x_loader.load()
size = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
passes = int(sys.argv[2]) if len(sys.argv) > 2 else 2
engine = {"tc": w2x.ENGINE_TC, "fp32": w2x.ENGINE_FP32}[sys.argv[3] if len(sys.argv) > 3 else "tc"]
om = oracle.OracleModel.golden("scale2.0x")
m = w2x.Model.from_arrays(om.weights, om.biases)
ctx = w2x.Context(0, engine=engine)      # precision: library default, or W2X_PRECISION=f16x3
if os.environ.get("W2X_ONE_BAND"):
    ctx.debug_set_host_bands(1)            # one whole-plane launch per layer (what bench.py's device-resident leg times)
x = oracle.seeded_plane(size, size, 1, "uniform")
for _ in range(passes):
    y = ctx.convert_plane(m, x)
print("ok", float(y.mean()))
