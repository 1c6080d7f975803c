"""bands.py -- host logic of the multi-GPU row-band mode (one process per GPU).

The conv stack has a 15x15 receptive field, so a plane shards into contiguous row bands that need
only n_model (=7) input rows from each neighbour.  Every rank owns one band; `exchange_halos` trades
those rows with torch.distributed point-to-point ops (NCCL over NVLink on GPUs, gloo in the CPU
tests), after which `w2x_convert_band_device` runs the band with no further communication -- the
"one-shot input halo, recompute" variant of SURVEY.md section 8(e): +14/(H/G) redundant rows
instead of a per-layer exchange.  At the image border the library replicates like the reference
(src/convertRoutine.cpp:35,96), so halo = 0 there.
"""
from __future__ import annotations


def partition_rows(height: int, world: int):
    """Contiguous, near-equal bands: [(row0, rows)] per rank (earlier ranks take the remainder)."""
    if world < 1 or height < world:
        raise ValueError("need at least one row per rank")
    base, rem = divmod(height, world)
    out, r0 = [], 0
    for r in range(world):
        n = base + (1 if r < rem else 0)
        out.append((r0, n))
        r0 += n
    return out


def halo_rows(rank: int, world: int, n_model: int = 7):
    """(rows_above, rows_below) of real neighbour data this rank needs."""
    return (n_model if rank > 0 else 0), (n_model if rank < world - 1 else 0)


def exchange_halos(ext, band_rows: int, rank: int, world: int, n_model: int = 7, dist=None):
    """ext: 2-D tensor [rows_above + band_rows + rows_below, W] whose middle holds this rank's band.
    Fills the halo rows from the neighbours.  Requires every band to have >= n_model rows."""
    if world == 1:
        return
    if dist is None:
        import torch.distributed as dist
    ra, rb = halo_rows(rank, world, n_model)
    if band_rows < n_model:
        raise ValueError("band thinner than the halo")
    band = ext[ra:ra + band_rows]
    ops = []
    if rank > 0:
        ops.append(dist.P2POp(dist.isend, band[:n_model].contiguous(), rank - 1))
        ops.append(dist.P2POp(dist.irecv, ext[:ra], rank - 1))
    if rank < world - 1:
        ops.append(dist.P2POp(dist.isend, band[band_rows - n_model:].contiguous(), rank + 1))
        ops.append(dist.P2POp(dist.irecv, ext[ra + band_rows:], rank + 1))
    for req in dist.batch_isend_irecv(ops):
        req.wait()


def run_band_per_layer(band, ext_in, ext_stride_bytes, d_out, out_stride_bytes, rank, world, dist=None, torch=None):
    """The per-layer variant (north_star): `band` is a w2x Band session; ext_in = device pointer of this rank's band
    input preceded / followed by ONE real neighbour row where a neighbour exists.  After every layer each rank
    sends its boundary row of the fresh activation to the neighbour and receives the neighbour's into its halo row."""
    if dist is None:
        import torch.distributed as dist
    if torch is None:
        import torch
    from .capi import DevBytes
    # The NCCL sends / receives are ordered on torch's CURRENT stream; the layer kernels must be queued on the same stream or
    # a row could be sent before its layer wrote it (and the next layer could read a halo row before it landed).
    cur = torch.cuda.current_stream().cuda_stream
    if cur == 0:
        raise RuntimeError("run_band_per_layer needs a non-default torch stream (stream handle 0 means 'the context's own stream')")
    band._ctx.set_stream(cur)
    band.load(ext_in, ext_stride_bytes)
    views = band.__dict__.setdefault("_p2p_views", {})     # zero-copy tensor views of the session's halo rows, built once
    for k in range(band.steps):
        band.step(k)
        if world == 1:
            continue
        if k not in views:
            v = []
            for (su, ru, sd, rd, nb) in band.halo(k):
                if rank > 0:
                    v.append((torch.as_tensor(DevBytes(su, nb), device="cuda"), torch.as_tensor(DevBytes(ru, nb), device="cuda"), rank - 1))
                if rank < world - 1:
                    v.append((torch.as_tensor(DevBytes(sd, nb), device="cuda"), torch.as_tensor(DevBytes(rd, nb), device="cuda"), rank + 1))
            views[k] = v
        ops = []
        for ts, tr, peer in views[k]:
            ops += [dist.P2POp(dist.isend, ts, peer), dist.P2POp(dist.irecv, tr, peer)]
        for req in dist.batch_isend_irecv(ops):
            req.wait()       # NCCL: orders the current stream after the transfer, does not block the host
    band.finish(d_out, out_stride_bytes)
