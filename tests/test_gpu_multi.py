"""Multi-GPU machinery of the product (csrc/engine_band.cu) on ONE device -- every mechanism is the same code the
multi-GPU runs use, only the "neighbour" lives on the same GPU (separate contexts / streams / processes):

  * w2x_convert_tiles: a batch of independent planes as one stacked frame is bit-identical to plane-by-plane calls
    (the reference's block loop, src/convertRoutine.cpp:114-165; BASELINE config 5);
  * w2x_band_connect_local + w2x_band_run: row bands with the per-layer halo exchange done INSIDE the library through
    peer-mapped frames (one exchange kernel per layer: rows, flag, wait) reproduce the whole-plane bits;
  * w2x_multi_*: the one-process N-context driver;
  * w2x_band_export / w2x_band_connect: the same between PROCESSES over CUDA IPC (what bench.py's torchrun ranks do).
"""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def models(w2x, oracle_models):
    return {n: w2x.Model.from_arrays(om.weights, om.biases) for n, om in oracle_models.items()}


@pytest.mark.parametrize("engine,precision", [("tc", 1), ("tc", 0), ("fp32", 1)])
def test_batched_tiles_equal_plane_by_plane(w2x, models, oracle_mod, engine, precision):
    ctx = w2x.Context(0, engine=w2x.ENGINE_TC if engine == "tc" else w2x.ENGINE_FP32)
    ctx.set_precision(precision)
    try:
        tiles = np.stack([oracle_mod.seeded_plane(96, 80, 100 + t, "uniform") for t in range(5)])
        batch = ctx.convert_tiles(models["noise2"], tiles)
        for t in range(5):
            assert np.array_equal(batch[t], ctx.convert_plane(models["noise2"], tiles[t], block_splitting=False)), t
        # a scratch limit smaller than the batch: processed in groups, same bits
        ctx.set_scratch_limit(128 * (96 + 14) * (80 + 14) * 4 * 2)
        assert np.array_equal(ctx.convert_tiles(models["noise2"], tiles), batch)
    finally:
        ctx.close()


@pytest.mark.parametrize("precision", [1, 0])
def test_peer_exchange_inside_the_library_three_bands(w2x, models, oracle_mod, precision):
    import torch
    W, H = 170, 150
    x = oracle_mod.seeded_plane(W, H, 31, "uniform")
    ref_ctx = w2x.Context(0, engine=w2x.ENGINE_TC)
    ref_ctx.set_precision(precision)
    whole = ref_ctx.convert_plane(models["scale2.0x"], x)
    ref_ctx.close()
    cuts = [0, 47, 101, H]
    ctxs = [w2x.Context(0, engine=w2x.ENGINE_TC) for _ in range(3)]       # own stream each: the three bands run concurrently
    for c in ctxs:
        c.set_precision(precision)
    bands = [w2x.Band(ctxs[b], models["scale2.0x"], W, cuts[b + 1] - cuts[b], b > 0, b < 2) for b in range(3)]
    for b in range(3):
        bands[b].connect_local(bands[b - 1] if b > 0 else None, bands[b + 1] if b < 2 else None)
    d_in = torch.from_numpy(x).cuda()
    outs = [torch.zeros((cuts[b + 1] - cuts[b], W), device="cuda") for b in range(3)]
    for rep in range(3):                                                  # repeated passes reuse frames and flags
        for b in range(3):
            bands[b].run(d_in[cuts[b]:].data_ptr(), W * 4, outs[b].data_ptr(), W * 4)
        for c in ctxs:
            c.synchronize()
        assert np.array_equal(torch.cat(outs).cpu().numpy(), whole), rep
    for b in bands:
        b.close()
    for c in ctxs:
        c.close()


def test_multi_driver_on_repeated_device(w2x, models, oracle_mod):
    """w2x_multi_* with the same device listed three times: three contexts, three bands, peer pointers = local pointers."""
    x = oracle_mod.seeded_plane(260, 300, 5, "uniform")
    single = w2x.Context(0)
    want = single.convert_plane(models["noise1"], x)
    tiles = np.stack([oracle_mod.seeded_plane(64, 48, 200 + t, "uniform") for t in range(7)])
    want_tiles = single.convert_tiles(models["noise1"], tiles)
    single.close()
    multi = w2x.Multi([0, 0, 0])
    try:
        for _ in range(2):
            assert np.array_equal(multi.convert_plane(models["noise1"], x), want)
        small = oracle_mod.seeded_plane(40, 30, 6, "uniform")             # too small to cut: runs on the first context
        one = w2x.Context(0)
        assert np.array_equal(multi.convert_plane(models["noise1"], small), one.convert_plane(models["noise1"], small))
        one.close()
        assert np.array_equal(multi.convert_tiles(models["noise1"], tiles), want_tiles)
    finally:
        multi.close()


@pytest.mark.parametrize("world,precision", [(2, 1), (3, 0)])
def test_ipc_ranks_in_separate_processes(w2x, models, oracle_mod, tmp_path, world, precision):
    W, H = 200, 190
    x = oracle_mod.seeded_plane(W, H, 77, "uniform")
    ctx = w2x.Context(0, engine=w2x.ENGINE_TC)
    ctx.set_precision(precision)
    whole = ctx.convert_plane(models["noise2"], x)
    ctx.close()
    procs = [subprocess.Popen([sys.executable, os.path.join(ROOT, "tools", "ipc_band_worker.py"), str(r), str(world), str(tmp_path), str(W), str(H),
                               "noise2", str(precision)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(world)]
    logs = [p.communicate(timeout=300)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), logs
    got = np.concatenate([np.load(tmp_path / f"out_{r}.npy") for r in range(world)])
    assert np.array_equal(got, whole)
