"""bench_configs.py -- the other BASELINE.json configurations, measured on one GPU (run under gpurun).

    python tools/bench_configs.py > gpurun_out/configs.json
cfg2  1920x1080 RGB, noise1 + scale2.0x: the two Y-plane passes main.cpp would issue (1920x1080 with noise1, then 3840x2160 with scale2.0x)
cfg4  8192x8192 Y plane, scale2.0x (single GPU baseline of the halo-tiled configuration)
cfg5  64 tiles of 512x512, noise2 (per-GPU rate of the one-tile-per-GPU configuration)
Device-resident planes, CUDA events, 3 warm-up + 5 timed repetitions each."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import w2x_loader  # noqa: E402
from oracle import oracle  # noqa: E402  (model fixtures + synthetic planes only)

w2x = w2x_loader.load()
models = {n: w2x.Model.from_arrays(oracle.OracleModel.golden(n).weights, oracle.OracleModel.golden(n).biases) for n in oracle.MODEL_NAMES}
ctx = w2x.Context(0)
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
ctx.set_stream(stream.cuda_stream)


def timed(fn, reps=5, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        fn()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def plane(w, h, seed):
    return torch.from_numpy(oracle.seeded_plane(w, h, seed, "uniform")).cuda()


out = {}
a, b = plane(1920, 1080, 1), plane(3840, 2160, 2)
oa, ob = torch.empty_like(a), torch.empty_like(b)


def cfg2():
    ctx.convert_plane_device(models["noise1"], a.data_ptr(), 1920, 1080, 1920 * 4, oa.data_ptr(), 1920 * 4)
    ctx.convert_plane_device(models["scale2.0x"], b.data_ptr(), 3840, 2160, 3840 * 4, ob.data_ptr(), 3840 * 4)


ms = timed(cfg2)
pix = 1920 * 1080 + 3840 * 2160
out["cfg2_1080p_noise1_plus_scale2x"] = {"ms_per_image": ms, "conv_Mpix_per_s": pix / ms / 1e3, "images_per_s": 1e3 / ms,
                                         "algorithmic_TFLOP_per_s": 574272 * pix / ms / 1e9}
c = plane(8192, 8192, 2)
oc = torch.empty_like(c)
ms = timed(lambda: ctx.convert_plane_device(models["scale2.0x"], c.data_ptr(), 8192, 8192, 8192 * 4, oc.data_ptr(), 8192 * 4), reps=3, warm=2)
out["cfg4_8192_single_gpu"] = {"ms": ms, "Mpix_per_s": 8192 * 8192 / ms / 1e3}
tiles = torch.from_numpy(np.stack([oracle.seeded_plane(512, 512, 3 + i, "uniform") for i in range(8)])).cuda()
ot = torch.empty_like(tiles)


def cfg5():
    for i in range(8):
        ctx.convert_plane_device(models["noise2"], tiles[i].data_ptr(), 512, 512, 512 * 4, ot[i].data_ptr(), 512 * 4)


ms = timed(cfg5)
out["cfg5_512_tiles_noise2_one_gpu"] = {"ms_per_tile": ms / 8, "tiles_per_s": 8e3 / ms, "Mpix_per_s": 8 * 512 * 512 / ms / 1e3,
                                        "note": "tiles run back to back on one GPU; 64 tiles over 8 GPUs = 8 rounds of this"}
print(json.dumps(out, indent=1))
