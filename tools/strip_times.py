"""strip_times.py -- per-layer milliseconds of one 4096x4096 pass, three repetitions (timing experiments with
-DW2X_EPI_EXPERIMENTS builds: W2X_DEBUG_STRIP / W2X_DEBUG_EPI; results of such runs are wrong by design)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import w2x_loader  # noqa: E402
from oracle import oracle  # noqa: E402

w2x = w2x_loader.load()
size = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
om = oracle.OracleModel.golden("scale2.0x")
m = w2x.Model.from_arrays(om.weights, om.biases)
ctx = w2x.Context(0, engine=w2x.ENGINE_TC)
ctx.debug_set_host_bands(1)
x = oracle.seeded_plane(size, size, 1, "uniform")
ctx.convert_plane(m, x)
ctx.set_timing(True)
for _ in range(3):
    ctx.convert_plane(m, x)
t = ctx.layer_times()
print(f"W2X_DEBUG_STRIP={os.environ.get('W2X_DEBUG_STRIP', '0')} W2X_STRIP_ROWS={os.environ.get('W2X_STRIP_ROWS', '32')}: per-layer ms",
      [round(a / max(b, 1), 3) for a, b, _ in t], "sum", round(sum(a / max(b, 1) for a, b, _ in t), 3))
