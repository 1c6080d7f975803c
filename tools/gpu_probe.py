"""gpu_probe.py -- first-contact diagnostics on a B200 (run under gpurun; writes to stdout).

For each engine and each MMA issue variant, runs per-layer Model::filter and the whole path on
small inputs and prints the max-abs error against the CPU oracle.  Each tcgen05 trial runs in its
own subprocess with a timeout, so a trapped/hung kernel cannot take the whole probe down."""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CHILD = r'''
import sys, json, time
import numpy as np
sys.path.insert(0, %(root)r)
import w2x_loader
from oracle import oracle
w2x = w2x_loader.load()
engine, mma_mode, what = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
om = oracle.OracleModel.golden("scale2.0x")
m = w2x.Model.from_arrays(om.weights, om.biases)
ctx = w2x.Context(0, engine=engine)
ctx.debug_set_mma_mode(mma_mode)
res = {}
if what == "layers":
    z = np.load(%(root)r + "/tests/golden/layers_32x24.npz")
    for li in range(7):
        if engine == 2 and li in (0, 6):
            continue
        out = ctx.filter_layer(m, li, z["in%%d" %% li])
        res["layer%%d" %% li] = float(np.abs(out - z["out%%d" %% li]).max())
else:
    n = int(what)
    x = oracle.seeded_plane(n, n, 0, "uniform")
    y = ctx.convert_plane(m, x)
    ctx.set_timing(True)
    t = time.time(); y = ctx.convert_plane(m, x); dt2 = time.time() - t
    layer_ms = [round(v[0], 3) for v in ctx.layer_times()]
    ref = om.convert(x, n_job=16) if n <= 512 else None
    res["convert%%d" %% n] = {"err": None if ref is None else float(np.abs(y - ref).max()), "second_s": round(dt2, 4),
                             "layer_ms": layer_ms, "finite": bool(np.isfinite(y).all())}
print("RESULT " + json.dumps(res))
'''


def run(engine, mma_mode, what, timeout=120):
    code = CHILD % {"root": ROOT}
    t = time.time()
    try:
        p = subprocess.run([sys.executable, "-c", code, str(engine), str(mma_mode), what], capture_output=True, text=True, timeout=timeout)
        tail = [l for l in p.stdout.splitlines() if l.startswith("RESULT ")]
        out = tail[-1][7:] if tail else ("rc=%d stderr=%s" % (p.returncode, p.stderr[-600:]))
    except subprocess.TimeoutExpired:
        out = "TIMEOUT"
    print(f"engine={engine} mma_mode={mma_mode} what={what} ({time.time() - t:.1f}s): {out}", flush=True)


if __name__ == "__main__":
    subprocess.run(["nvidia-smi", "--query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total", "--format=csv"])
    modes = [int(a) for a in sys.argv[1:]] or [0, 1, 2]
    for mode in modes:
        run(2, mode, "layers")
        run(2, mode, "256")
        run(2, mode, "4096")
