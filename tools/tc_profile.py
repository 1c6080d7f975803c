"""tc_profile.py -- per-role wait/work cycle breakdown of the tcgen05 layer kernels (run under gpurun).

    python tools/tc_profile.py [size] > gpurun_out/tc_profile.txt
Uses the kernel's own clock64 counters (w2x_debug_tc_profile_*): for every layer, cycles per
tile-set the MMA issuer spent waiting for accumulators / staged activations / weight stages, what
the producers waited for, and how long an epilogue pass takes."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import w2x_loader  # noqa: E402
from oracle import oracle  # noqa: E402  (model fixture only)

w2x = w2x_loader.load()
size = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
om = oracle.OracleModel.golden("scale2.0x")
m = w2x.Model.from_arrays(om.weights, om.biases)
ctx = w2x.Context(0, engine=w2x.ENGINE_TC)
x = oracle.seeded_plane(size, size, 1, "uniform")
mode = int(sys.argv[2]) if len(sys.argv) > 2 else 0      # 0 = f16x3, 1 = f16 + 2 x e4m3 corrections
ctx.set_precision(mode)
ctx.debug_set_pair(not (len(sys.argv) > 3 and sys.argv[3] == "single"))
ctx.convert_plane(m, x)
ctx.convert_plane(m, x)
ctx.set_timing(True)
ctx.convert_plane(m, x)
clean = ctx.layer_times()
print(f"size {size}x{size} precision {mode}; per-layer ms without counters:", [round(t[0], 3) for t in clean], "sum", round(sum(t[0] for t in clean), 3))
ctx.debug_tc_profile_enable(True)
ctx.convert_plane(m, x)
times = ctx.layer_times()
print(f"per-layer ms with counters:", [round(t[0], 3) for t in times])
print("layer  ms     cyc/unit     mma_wait_acc  mma_wait_a  mma_wait_b  issue+other | aprod_wait bprod_wait | epi_wait epi_work  (cycles per unit = 128-pixel strip on L1-L3, pair of 16x16 tile-sets on L4-L5; per-CTA average)")
for li in range(1, 6):
    d = ctx.debug_tc_profile_read(li)
    n = max(d["tilesets"], 1)
    tot = d["total"] / n
    rest = (d["total"] - d["mma_wait_acc"] - d["mma_wait_a"] - d["mma_wait_b"]) / n
    print(f"L{li}   {times[li][0]:7.3f} {tot:10.0f} {d['mma_wait_acc']/n:12.0f} {d['mma_wait_a']/n:11.0f} {d['mma_wait_b']/n:11.0f} {rest:12.0f} | "
          f"{d['aprod_wait']/n:10.0f} {d['bprod_wait']/n:10.0f} | {d['epi_wait']/n:8.0f} {d['epi_work']/n:8.0f}   ctas={d['ctas']} tilesets/cta={n:.0f}")
ctx.debug_tc_profile_enable(False)
