"""multi_selftest.py -- the multi-GPU machinery on whatever GPUs are visible (run under gpurun; tests/test_gpu_multi.py runs it):

    python tools/multi_selftest.py [n_bands]

With ONE visible device the bands / contexts all live on it (peer pointers = local pointers); the exchange kernels of
different bands then have to run CONCURRENTLY on that device, so the process needs more hardware work queues than streams:
CUDA_DEVICE_MAX_CONNECTIONS=32 must be in the environment before CUDA initialises (two streams sharing a queue would put
one band's layers behind another band's waiting exchange kernel).  With N >= n_bands devices every band gets its own GPU."""
import os
import sys

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import numpy as np  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import w2x_loader  # noqa: E402
from oracle import oracle  # noqa: E402  (model fixtures + synthetic planes only)

w2x = w2x_loader.load()
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 3
ndev = torch.cuda.device_count()
devs = [b % ndev for b in range(nb)] if ndev < nb else list(range(nb))
models = {n: w2x.Model.from_arrays(om.weights, om.biases) for n, om in ((n, oracle.OracleModel.golden(n)) for n in ("scale2.0x", "noise1"))}

# ---- w2x_band_connect_local + w2x_band_run: per-layer halo exchange inside the library ----
W, H = 170, 150
x = oracle.seeded_plane(W, H, 31, "uniform")
for precision in (1, 0):
    ref = w2x.Context(0, engine=w2x.ENGINE_TC)
    ref.set_precision(precision)
    whole = ref.convert_plane(models["scale2.0x"], x)
    ref.close()
    cuts = [H * b // nb for b in range(nb + 1)]
    cuts[1] -= 3                                       # uneven bands
    ctxs = [w2x.Context(d, engine=w2x.ENGINE_TC) for d in devs]
    for c in ctxs:
        c.set_precision(precision)
    bands = [w2x.Band(ctxs[b], models["scale2.0x"], W, cuts[b + 1] - cuts[b], b > 0, b < nb - 1) for b in range(nb)]
    for b in range(nb):
        bands[b].connect_local(bands[b - 1] if b > 0 else None, bands[b + 1] if b < nb - 1 else None)
    d_in = [torch.from_numpy(np.ascontiguousarray(x[cuts[b]:cuts[b + 1]])).to(f"cuda:{devs[b]}") for b in range(nb)]
    outs = [torch.zeros_like(t) for t in d_in]
    for rep in range(3):                               # repeated passes reuse frames and flags
        for b in range(nb):
            bands[b].run(d_in[b].data_ptr(), W * 4, outs[b].data_ptr(), W * 4)
        for c in ctxs:
            c.synchronize()
        got = np.concatenate([o.cpu().numpy() for o in outs])
        assert np.array_equal(got, whole), (precision, rep, float(np.abs(got - whole).max()))
    for b in bands:
        b.close()
    for c in ctxs:
        c.close()
    print(f"band sessions + peer exchange, precision {precision}, devices {devs}: bit-identical to one GPU", flush=True)

# ---- w2x_slab_*: host rows in / out, copy pipeline over sub-bands, outer edges exchanged with the neighbour slabs ----
W2, H2 = 190, 420
x2 = oracle.seeded_plane(W2, H2, 9, "uniform")
ref = w2x.Context(0, engine=w2x.ENGINE_TC)
whole2 = ref.convert_plane(models["noise1"], x2)
ref.close()
cuts = [H2 * b // nb for b in range(nb + 1)]
ctxs = [w2x.Context(d, engine=w2x.ENGINE_TC) for d in devs]
slabs = [w2x.Slab(ctxs[b], models["noise1"], W2, cuts[b + 1] - cuts[b], b > 0, b < nb - 1, order=b & 1, n_sub=2 + (b == 1)) for b in range(nb)]
for b in range(nb):
    slabs[b].connect_local(slabs[b - 1] if b > 0 else None, slabs[b + 1] if b < nb - 1 else None)
h_in = torch.from_numpy(x2).pin_memory().numpy()
h_out = torch.zeros((H2, W2)).pin_memory().numpy()
for rep in range(3):
    h_out[:] = 0
    for b in range(nb):
        slabs[b].convert_async(h_in[cuts[b]:cuts[b + 1]], h_out[cuts[b]:cuts[b + 1]])
    for sl in slabs:
        sl.synchronize()
    assert np.array_equal(h_out, whole2), (rep, float(np.abs(h_out - whole2).max()))
for sl in slabs:
    sl.close()
for c in ctxs:
    c.close()
print(f"slabs (sub-band copy pipeline, alternating order) on devices {devs}: bit-identical to one GPU", flush=True)

# ---- w2x_multi_*: the one-process driver ----
x = oracle.seeded_plane(260, 1700, 5, "uniform")             # tall enough for sub-bands inside every slab
single = w2x.Context(0)
want = single.convert_plane(models["noise1"], x)
tiles = np.stack([oracle.seeded_plane(64, 48, 200 + t, "uniform") for t in range(7)])
want_tiles = single.convert_tiles(models["noise1"], tiles)
small = oracle.seeded_plane(40, 30, 6, "uniform")
want_small = single.convert_plane(models["noise1"], small)
single.close()
multi = w2x.Multi(devs)
for _ in range(2):
    assert np.array_equal(multi.convert_plane(models["noise1"], x), want)
assert np.array_equal(multi.convert_plane(models["noise1"], small), want_small)      # too small to cut: first context
assert np.array_equal(multi.convert_tiles(models["noise1"], tiles), want_tiles)
multi.close()
print(f"w2x_multi_convert_plane / w2x_multi_convert_tiles on devices {devs}: bit-identical to one GPU")
print("multi selftest ok")
