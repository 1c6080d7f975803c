"""Multi-GPU host logic on CPU: world_size-2 and -3 gloo runs of the row-band partition + halo
exchange (waifu2x-converter-cpp_b200/bands.py).  The band compute is done by the CPU oracle here
(tests may use it as the checker); the GPU band entry itself is covered by
test_gpu_parity.py::test_device_entry_points_and_band_mode."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, W, H, outdir):
    sys.path.insert(0, ROOT)
    import w2x_loader
    from oracle import oracle
    w2x = w2x_loader.load()
    from w2x_b200 import bands
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        plane = oracle.seeded_plane(W, H, 42, "uniform")             # every rank can build the truth
        r0, n = bands.partition_rows(H, world)[rank]
        ra, rb = bands.halo_rows(rank, world)
        ext = torch.full((ra + n + rb, W), float("nan"))
        ext[ra:ra + n] = torch.from_numpy(plane[r0:r0 + n])
        bands.exchange_halos(ext, n, rank, world)
        assert torch.equal(ext, torch.from_numpy(plane[r0 - ra:r0 + n + rb])), "halo rows are not the neighbour's rows"
        om = oracle.OracleModel.golden("noise1")
        out = om.convert(ext.numpy(), n_job=2)[ra:ra + n]
        np.save(os.path.join(outdir, f"band{rank}.npy"), out)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_band_partition_and_halo_exchange(tmp_path, oracle_mod, world):
    W, H = 40, 53
    mp.spawn(_worker, args=(world, _free_port(), W, H, str(tmp_path)), nprocs=world, join=True)
    whole = oracle_mod.OracleModel.golden("noise1").convert(oracle_mod.seeded_plane(W, H, 42, "uniform"), n_job=4)
    got = np.concatenate([np.load(tmp_path / f"band{r}.npy") for r in range(world)])
    assert np.array_equal(got, whole)


def test_partition_rows():
    import w2x_loader
    w2x_loader.load()
    from w2x_b200 import bands
    assert bands.partition_rows(10, 3) == [(0, 4), (4, 3), (7, 3)]
    assert bands.partition_rows(4096 * 4, 4) == [(i * 4096, 4096) for i in range(4)]
    assert bands.halo_rows(0, 1) == (0, 0) and bands.halo_rows(0, 2) == (0, 7) and bands.halo_rows(1, 3) == (7, 7)
    with pytest.raises(ValueError):
        bands.partition_rows(2, 3)
