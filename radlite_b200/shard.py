"""Multi-GPU sharding of the line ray-tracer: one process per GPU, no data-path collective.

Every (line, ray, channel) of the path is independent (main.F:1043 line loop, telescope.F:532-533 ray
loops), so ranks take disjoint pieces and the only exchange is the final gather of the results on
rank 0:

* multi-line spectra (BASELINE configs 2, 4, 5): contiguous line blocks, split exactly like the
  reference drivers split lines over processes (``split_lines`` = pyradlite radlite.py:1163-1169,
  ``split_lines_idl`` = PRO/line_run.pro:79); rank 0 gathers ``flux[nlines][nfr]``.
* single-line renders (configs 1, 3): contiguous camera-ring blocks (``split_rings``); every rank
  returns the per-ring flux contributions of its block, rank 0 adds the (disjoint) arrays and runs
  the reference's index-ordered ring sum once -- bit-identical to the one-GPU result.

The engine is anything with the call shapes of ``radlite_b200.api.Renderer`` (``render``,
``render_rings``, ``flux_from_rings``); the collectives are ``torch.distributed`` (NCCL on the GPU
box, gloo in the CPU tests).  Nothing here computes any part of the path.
"""
from __future__ import annotations

import numpy as np


def split_lines(numlines: int, numcores: int):
    """[start, stop) line intervals per rank, radlite.py:1163-1169: numlines // numcores each, the
    remainder spread over the first ranks."""
    base, rem = divmod(int(numlines), int(numcores))
    sizes = [base + (1 if k < rem else 0) for k in range(numcores)]
    edges = np.concatenate(([0], np.cumsum(sizes)))
    return [(int(edges[k]), int(edges[k + 1])) for k in range(numcores)]


def split_lines_idl(nlines: int, ncores: int):
    """line_run.pro:78-79: ncores = min(ncores, nlines); subN = CEIL(nlines / ncores) lines per
    core, the last core takes what is left."""
    ncores = min(int(ncores), int(nlines))
    sub = -(-int(nlines) // ncores)
    return [(min(nlines, k * sub), min(nlines, (k + 1) * sub)) for k in range(ncores)]


def split_rings(nrr: int, world: int, cost=None):
    """Contiguous camera-ring blocks [lo, hi] (inclusive; ring 0 = the central beam) per rank,
    balanced by ``cost[ir]`` (default: equal cost per ring).  Ranks beyond the ring count get an
    empty block (lo > hi)."""
    n = nrr + 1
    cost = np.ones(n) if cost is None else np.asarray(cost, dtype=np.float64)
    assert cost.shape == (n,) and np.all(cost >= 0)
    cum = np.concatenate(([0.0], np.cumsum(cost)))
    total = cum[-1] if cum[-1] > 0 else 1.0
    edges = [0]
    for k in range(1, world):
        e = int(np.searchsorted(cum, total * k / world, side="left"))
        edges.append(min(n, max(edges[-1], e)))
    edges.append(n)
    return [(edges[k], edges[k + 1] - 1) for k in range(world)]


def rebalance_rings(cost, blocks, times):
    """One correction of the per-ring work estimate from measured times: the estimate of every ring of block k
    is scaled so that the block's sum equals the time rank k needed (``times[k]``, any unit), and the rings are
    cut again.  A rank's time also holds a part that does not move with its rings (launches, per-line tables),
    which this model spreads over the block's rings, so a step under-corrects a little and the iteration
    approaches equal times from one side; three or four steps settle it.  Returns (cost, blocks)."""
    cost = np.array(cost, dtype=np.float64)
    times = np.asarray(times, dtype=np.float64)
    world = len(blocks)
    for k, (a, b) in enumerate(blocks):
        if b >= a and times[k] > 0:
            cost[a:b + 1] *= times[k] / max(cost[a:b + 1].sum(), 1e-300)
    return cost, split_rings(len(cost) - 1, world, cost)


def _dist():
    import torch.distributed as dist
    return dist


def _gather_rows(local: np.ndarray, rows_per_rank, rank: int, world: int, device):
    """Gather row blocks of unequal height on rank 0 (padded to the largest block)."""
    import torch
    dist = _dist()
    width = local.shape[1:]
    hmax = max(max(rows_per_rank), 1)
    buf = torch.zeros((hmax,) + tuple(width), dtype=torch.float64, device=device)
    if local.shape[0]:
        buf[:local.shape[0]] = torch.from_numpy(np.ascontiguousarray(local)).to(device)
    parts = [torch.zeros_like(buf) for _ in range(world)] if rank == 0 else None
    dist.gather(buf, parts, dst=0)
    if rank != 0:
        return None
    return np.concatenate([p[:h].cpu().numpy() for p, h in zip(parts, rows_per_rank)], axis=0)


def render_spectrum_sharded(engine, nlines, nfr, vmax_kms, dist_cm, rank=0, world=1, device="cpu",
                            split=split_lines):
    """One spectrum of ``nlines`` lines, line blocks over ``world`` ranks.  Returns flux[nlines, nfr]
    on rank 0 (None elsewhere).  Mirrors the drivers' one-process-per-line-chunk runs; the gather
    replaces their concatenation of the per-core linespectrum_moldata_<n>.dat files."""
    blocks = split(nlines, world)
    i0, i1 = blocks[rank] if rank < len(blocks) else (nlines, nlines)
    if i1 > i0:
        local = engine.render(i0 + 1, i1 - i0, nfr, vmax_kms, dist_cm)["flux"]
    else:
        local = np.zeros((0, nfr))
    if world == 1:
        return local
    rows = [b[1] - b[0] for b in blocks] + [0] * (world - len(blocks))
    return _gather_rows(local, rows, rank, world, device)


def render_line_ring_sharded(engine, iline0, nl, nfr, vmax_kms, dist_cm, rank=0, world=1, device="cpu",
                             cost=None, image=None):
    """Lines iline0..iline0+nl-1 with the camera rings split over ``world`` ranks.  Returns
    flux[nl, nfr] on rank 0 (None elsewhere), bit-identical to the unsharded render.  ``image``
    (optional full-size cube) receives this rank's rows."""
    import torch
    nrr, _, _ = engine.camera_dims()
    lo, hi = split_rings(nrr, world, cost)[rank]
    if hi >= lo:
        rs = engine.render_rings(iline0, nl, nfr, vmax_kms, dist_cm, lo, hi, image=image)
    else:
        rs = np.zeros((nl, nrr + 1, nfr))
    if world > 1:
        # blocks are disjoint and the other rows are exactly 0: the sum is a concatenation
        t = torch.from_numpy(rs).to(device)
        _dist().reduce(t, dst=0, op=_dist().ReduceOp.SUM)
        rs = t.cpu().numpy()
        if rank != 0:
            return None
    return engine.flux_from_rings(rs, dist_cm)
