"""Host-side sharding logic (radlite_b200/shard.py): line splits like the reference drivers, ring
blocks, and the world_size-2 gather/reduce over gloo with the CPU oracle standing in for the GPU
engine.  The sharded results must be bit-identical to the unsharded render."""
import os
import sys
import tempfile

import numpy as np
import pytest

from helpers import OracleEngine, tiny
from radlite_b200 import shard, synth

HERE = os.path.dirname(os.path.abspath(__file__))


def _ref_split(numlines, numcores):
    # restatement of pyradlite radlite.py:1163-1169 with its own numpy calls
    t = np.array([numlines // numcores] * numcores)
    t[0:numlines % numcores] += 1
    c = np.concatenate((np.array([0]), np.cumsum(t)))
    return [(int(c[i]), int(c[i + 1])) for i in range(numcores)]


@pytest.mark.parametrize("n,cores", [(100, 1), (100, 2), (100, 8), (26, 3), (5, 8), (2000, 8), (1, 1)])
def test_split_lines_matches_pyradlite(n, cores):
    s = shard.split_lines(n, cores)
    assert s == _ref_split(n, cores)
    assert s[0][0] == 0 and s[-1][1] == n and all(a[1] == b[0] for a, b in zip(s, s[1:]))


def test_split_lines_tutorial_case():
    # tutorial_RadliteModel.ipynb: 26 lines on 3 cores -> 9 / 9 / 8
    assert [b - a for a, b in shard.split_lines(26, 3)] == [9, 9, 8]


@pytest.mark.parametrize("n,cores", [(100, 8), (26, 3), (5, 8), (7, 2)])
def test_split_lines_idl(n, cores):
    s = shard.split_lines_idl(n, cores)
    sub = int(np.ceil(n / float(min(cores, n))))
    assert all(b - a <= sub for a, b in s) and s[0][0] == 0 and max(b for _, b in s) == n
    assert sum(b - a for a, b in s) == n


@pytest.mark.parametrize("nrr,world", [(269, 1), (269, 2), (269, 8), (3, 8), (169, 4)])
def test_split_rings_cover(nrr, world):
    b = shard.split_rings(nrr, world)
    seen = [ir for lo, hi in b for ir in range(lo, hi + 1)]
    assert seen == list(range(nrr + 1))
    cost = np.arange(nrr + 1, dtype=float) + 1.0
    b = shard.split_rings(nrr, world, cost)
    seen = [ir for lo, hi in b for ir in range(lo, hi + 1)]
    assert seen == list(range(nrr + 1))
    if world > 1 and nrr > 50:
        loads = [cost[lo:hi + 1].sum() for lo, hi in b]
        assert max(loads) < 1.2 * cost.sum() / world


def _worker(rank, world, port, outdir):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        m = tiny(2, nlines=5)
        eng = OracleEngine(m)
        flux = shard.render_spectrum_sharded(eng, m.nlines, m.nfr, m.passband, synth.PARSEC, rank, world)
        m1 = tiny(1)
        eng1 = OracleEngine(m1)
        nrr, nphi, _ = eng1.camera_dims()
        cube = np.zeros((1, nrr + 1, nphi, m1.nfr))
        f1 = shard.render_line_ring_sharded(eng1, 1, 1, m1.nfr, m1.passband, synth.PARSEC, rank, world,
                                            image=cube)
        np.save(os.path.join(outdir, f"cube{rank}.npy"), cube)
        if rank == 0:
            np.save(os.path.join(outdir, "flux.npy"), flux)
            np.save(os.path.join(outdir, "flux1.npy"), f1)
        else:
            assert flux is None and f1 is None
    finally:
        dist.destroy_process_group()


def test_world2_gloo_line_and_ring_sharding(oracle_cls):
    import torch.multiprocessing as mp
    world = 2
    port = 29500 + (os.getpid() % 2000)
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_worker, args=(world, port, d), nprocs=world, join=True)
        flux = np.load(os.path.join(d, "flux.npy"))
        flux1 = np.load(os.path.join(d, "flux1.npy"))
        cubes = [np.load(os.path.join(d, f"cube{r}.npy")) for r in range(world)]
    m = tiny(2, nlines=5)
    o = oracle_cls()
    o.load_model(m)
    ref = o.render(1, m.nlines, m.nfr, m.passband, synth.PARSEC)["flux"]
    assert np.array_equal(flux, ref)
    m1 = tiny(1)
    o1 = oracle_cls()
    o1.load_model(m1)
    ref1 = o1.render(1, 1, m1.nfr, m1.passband, synth.PARSEC, want_image=True)
    assert np.array_equal(flux1, ref1["flux"])
    # the ranks' cube slabs are disjoint and tile the full cube
    nrr = ref1["image"].shape[1] - 1
    blocks = shard.split_rings(nrr, world)
    for r, (lo, hi) in enumerate(blocks):
        assert np.array_equal(cubes[r][:, lo:hi + 1], ref1["image"][:, lo:hi + 1])
        other = np.delete(cubes[r], np.s_[lo:hi + 1], axis=1)
        assert not other.any()


def test_single_rank_paths_equal_unsharded(oracle_cls):
    m = tiny(2, nlines=3)
    eng = OracleEngine(m)
    a = shard.render_spectrum_sharded(eng, m.nlines, m.nfr, m.passband, synth.PARSEC)
    b = shard.render_line_ring_sharded(eng, 1, m.nlines, m.nfr, m.passband, synth.PARSEC)
    o = oracle_cls()
    o.load_model(m)
    ref = o.render(1, m.nlines, m.nfr, m.passband, synth.PARSEC)["flux"]
    assert np.array_equal(a, ref) and np.array_equal(b, ref)


def test_rebalance_rings_converges_with_a_fixed_part():
    """Ranks whose time is a fixed part plus the (unknown) true cost of their rings: a few corrections of a
    wrong estimate bring the slowest rank within a few per cent of the mean."""
    rng = np.random.default_rng(5)
    nrr, world, fixed = 269, 8, 2.4
    ir = np.arange(nrr + 1)
    true = 0.02 + 0.3 * np.exp(-((ir - 70) / 60.0) ** 2) + 0.1 * (ir > 200)
    true *= 36.0 / true.sum()
    est = np.ones(nrr + 1)  # a poor first estimate
    blocks = shard.split_rings(nrr, world, est)

    def measure(blocks):
        return np.array([fixed + true[a:b + 1].sum() for a, b in blocks]) * (1 + 0.005 * rng.standard_normal(world))

    t0 = measure(blocks)
    for _ in range(5):
        est, blocks = shard.rebalance_rings(est, blocks, measure(blocks))
        assert blocks[0][0] == 0 and blocks[-1][1] == nrr
        assert all(blocks[k + 1][0] == blocks[k][1] + 1 for k in range(world - 1))
    t = measure(blocks)
    assert t.max() < 0.75 * t0.max()
    assert t.max() / t.mean() < 1.06, (t, blocks)
