"""multi_debug.py -- stage-by-stage diagnosis of the peer exchange on one device (development aid)."""
import os
import sys

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import numpy as np  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import w2x_loader  # noqa: E402
from oracle import oracle  # noqa: E402

w2x = w2x_loader.load()
om = oracle.OracleModel.golden("scale2.0x")
m = w2x.Model.from_arrays(om.weights, om.biases)
W, H = 170, 150
x = oracle.seeded_plane(W, H, 31, "uniform")
nb = 2
cuts = [0, 70, H]
ctxs = [w2x.Context(0, engine=w2x.ENGINE_TC) for _ in range(nb)]
bands = [w2x.Band(ctxs[b], m, W, cuts[b + 1] - cuts[b], b > 0, b < nb - 1) for b in range(nb)]
for b in range(nb):
    bands[b].connect_local(bands[b - 1] if b > 0 else None, bands[b + 1] if b < nb - 1 else None)
d_in = [torch.from_numpy(np.ascontiguousarray(x[cuts[b]:cuts[b + 1]])).cuda() for b in range(nb)]
outs = [torch.zeros_like(t) for t in d_in]


def sync(tag):
    for i, c in enumerate(ctxs):
        try:
            c.synchronize()
        except Exception as e:
            print(f"FAIL after {tag} on ctx {i}: {e}", flush=True)
            sys.exit(1)
    print("ok:", tag, flush=True)


for b in range(nb):
    bands[b].load_rows(d_in[b].data_ptr(), W * 4)
sync("load_rows")
for b in range(nb):
    bands[b].exchange(-1)
sync("exchange(-1)")
for k in range(bands[0].steps):
    for b in range(nb):
        bands[b].step(k)
    sync(f"step({k})")
    for b in range(nb):
        bands[b].exchange(k)
    sync(f"exchange({k})")
for b in range(nb):
    bands[b].finish(outs[b].data_ptr(), W * 4)
sync("finish")
ref = w2x.Context(0, engine=w2x.ENGINE_TC)
whole = ref.convert_plane(m, x)
got = np.concatenate([o.cpu().numpy() for o in outs])
print("max abs diff vs one GPU:", float(np.abs(got - whole).max()), "equal:", bool(np.array_equal(got, whole)))
