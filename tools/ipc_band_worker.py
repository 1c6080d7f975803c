"""ipc_band_worker.py -- one rank of a multi-process row-band job (CUDA IPC peer memory, no torch.distributed):

    python tools/ipc_band_worker.py <rank> <world> <dir> <width> <height> <model> <precision> [device]

Each rank owns rows [H*rank/world, H*(rank+1)/world) of the seeded plane, exports its band session's buffers into <dir>,
maps its neighbours', runs w2x_band_run twice (the second pass exercises buffer reuse across passes) and writes its rows
to <dir>/out_<rank>.npy.  tests/test_gpu_multi.py launches the ranks and compares the stitched plane with one GPU."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import w2x_loader  # noqa: E402
from oracle import oracle  # noqa: E402  (model fixture + synthetic plane only)

rank, world, d, W, H, name, prec = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], int(sys.argv[4]), int(sys.argv[5]), sys.argv[6], int(sys.argv[7])
device = int(sys.argv[8]) if len(sys.argv) > 8 else 0
import torch  # noqa: E402  (device buffers only)

w2x = w2x_loader.load()
om = oracle.OracleModel.golden(name)
m = w2x.Model.from_arrays(om.weights, om.biases)
torch.cuda.set_device(device)
ctx = w2x.Context(device, engine=w2x.ENGINE_TC)
ctx.set_precision(prec)
x = oracle.seeded_plane(W, H, 77, "uniform")
r0, r1 = H * rank // world, H * (rank + 1) // world
band = w2x.Band(ctx, m, W, r1 - r0, rank > 0, rank < world - 1)
open(os.path.join(d, f"blob_{rank}.tmp"), "wb").write(band.export())
os.replace(os.path.join(d, f"blob_{rank}.tmp"), os.path.join(d, f"blob_{rank}"))


def blob(r):
    p = os.path.join(d, f"blob_{r}")
    t0 = time.time()
    while not os.path.exists(p):
        if time.time() - t0 > 120:
            raise SystemExit(f"rank {rank}: neighbour {r} never exported")
        time.sleep(0.02)
    return open(p, "rb").read()


band.connect(blob(rank - 1) if rank > 0 else None, blob(rank + 1) if rank < world - 1 else None)
d_in = torch.from_numpy(np.ascontiguousarray(x[r0:r1])).cuda()
d_out = torch.zeros_like(d_in)
for _ in range(2):
    band.run(d_in.data_ptr(), W * 4, d_out.data_ptr(), W * 4)
ctx.synchronize()
np.save(os.path.join(d, f"out_{rank}.npy"), d_out.cpu().numpy())
# keep the exported buffers alive until every neighbour has finished with them
open(os.path.join(d, f"done_{rank}"), "w").write("1")
t0 = time.time()
while not all(os.path.exists(os.path.join(d, f"done_{r}")) for r in range(world)) and time.time() - t0 < 120:
    time.sleep(0.02)
band.close()
ctx.close()
