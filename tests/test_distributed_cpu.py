"""world_size-2 (and 4) runs of the slab pipeline's exchange choreography over
gloo on CPU tensors, with oracle-backed stages, against the single-process
oracle.  The CUDA stages themselves are covered by the -m gpu tests."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _particles(n, box, seed, var_mass):
    rng = np.random.default_rng(seed)
    pos = ((rng.random((n, 3)) * 1.1 - 0.05) * box).astype(np.float32)       # some outside the box
    masses = (10.0 ** rng.uniform(-1, 1, n)).astype(np.float32) if var_mass else None
    return pos, masses


def _slab_local_particles(n, box, dims, seed, jitter_cells):
    """Sorted by x, so index-range shards are x-slabs up to a jitter of a few cells."""
    rng = np.random.default_rng(seed)
    pos = (rng.random((n, 3)) * box).astype(np.float32)
    pos = pos[np.argsort(pos[:, 0], kind="stable")]
    pos[:, 0] = np.mod(pos[:, 0] + rng.uniform(-jitter_cells, jitter_cells, n) * (box / dims), box).astype(np.float32)
    return np.ascontiguousarray(pos)


def _wide_worker(rank, world, port, dims, n, box, ghost, jitter, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from genpk_b200.distributed import SlabPipeline
        from tests.slab_numpy_stages import NumpyStages
        pos = _slab_local_particles(n, box, dims, 9, jitter)
        lo, hi = rank * n // world, (rank + 1) * n // world
        pipe = SlabPipeline(dims, NumpyStages(dims, world, rank, ghost))
        assert pipe.placement == "local"
        shard = torch.from_numpy(pos[lo:hi].reshape(-1).copy())
        p, c, k = pipe.pk(shard, None, 0.5, box, 0.5 * n, dims)
        first = pipe.placement
        p2, c2, k2 = pipe.pk(shard, None, 0.5, box, 0.5 * n, dims)         # the decision is sticky
        assert np.array_equal(c, c2) and np.allclose(p, p2, rtol=1e-12)
        if rank == 0:
            np.savez(os.path.join(out_dir, "out.npz"), p=p, c=c, k=k, placement=first)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,dims,ghost,jitter,expect", [(2, 32, 4, 2.5, "local"), (4, 32, 3, 2.0, "local"),
                                                            (2, 32, 2, 6.0, "route"), (4, 16, 1, 3.0, "route")])
def test_wide_ghost_slabs_skip_the_particle_exchange(tmp_path, port, world, dims, ghost, jitter, expect):
    """Slab-local shards (sorted by x, stragglers within the ghosts) are deposited without
    routing; stragglers beyond the ghosts make every rank fall back to the routed path."""
    n, box = 8000, 64.0
    mp.spawn(_wide_worker, args=(world, _free_port(), dims, n, box, ghost, jitter, str(tmp_path)), nprocs=world, join=True)
    got = np.load(tmp_path / "out.npz")
    assert str(got["placement"]) == expect
    pos = _slab_local_particles(n, box, dims, 9, jitter)
    _, pr, cr, kr = port.pk(box, dims, pos, None, 0.5, 0.5 * n, dims)
    assert np.array_equal(got["c"], cr)
    nz = cr > 0
    np.testing.assert_allclose(got["p"][nz], pr[nz], rtol=1e-9)


def _worker(rank, world, port, dims, n, box, var_mass, out_dir, fused=False, scatter=False):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from genpk_b200.distributed import SlabPipeline
        from tests.slab_numpy_stages import NumpyStages
        pos, masses = _particles(n, box, 123, var_mass)
        lo, hi = rank * n // world, (rank + 1) * n // world                 # rank-sharded, arbitrary order
        tm = float(masses.astype(np.float64).sum()) if var_mass else 0.5 * n
        pipe = SlabPipeline(dims, NumpyStages(dims, world, rank, 0, fused, scatter))
        p, c, k = pipe.pk(torch.from_numpy(pos[lo:hi].reshape(-1).copy()),
                          torch.from_numpy(masses[lo:hi].copy()) if var_mass else None, 0.5, box, tm, dims)
        if rank == 0:
            np.savez(os.path.join(out_dir, "out.npz"), p=p, c=c, k=k)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,dims,var_mass", [(2, 16, False), (2, 32, True), (4, 16, True)])
def test_slab_pipeline_over_gloo(tmp_path, port, world, dims, var_mass):
    n, box = 6000, 100.0
    mp.spawn(_worker, args=(world, _free_port(), dims, n, box, var_mass, str(tmp_path)), nprocs=world, join=True)
    got = np.load(tmp_path / "out.npz")
    pos, masses = _particles(n, box, 123, var_mass)
    tm = float(masses.astype(np.float64).sum()) if var_mass else 0.5 * n
    _, pr, cr, kr = port.pk(box, dims, pos, masses, 0.5, tm, dims)
    assert np.array_equal(got["c"], cr)
    nz = cr > 0
    np.testing.assert_allclose(got["p"][nz], pr[nz], rtol=1e-9)
    np.testing.assert_allclose(got["k"][nz], kr[nz], rtol=1e-12)


@pytest.mark.parametrize("world,dims,scatter", [(2, 16, False), (2, 16, True), (4, 32, True)])
def test_fused_and_scatter_branches_over_gloo(tmp_path, port, world, dims, scatter):
    """SlabPipeline.pk's other two branches: x pass + binning as one stage after the all-to-all,
    and the y pass that delivers its rows itself between two barriers (no transpose step)."""
    n, box = 6000, 100.0
    mp.spawn(_worker, args=(world, _free_port(), dims, n, box, True, str(tmp_path), True, scatter), nprocs=world, join=True)
    got = np.load(tmp_path / "out.npz")
    pos, masses = _particles(n, box, 123, True)
    tm = float(masses.astype(np.float64).sum())
    _, pr, cr, kr = port.pk(box, dims, pos, masses, 0.5, tm, dims)
    assert np.array_equal(got["c"], cr)
    nz = cr > 0
    np.testing.assert_allclose(got["p"][nz], pr[nz], rtol=1e-9)
    np.testing.assert_allclose(got["k"][nz], kr[nz], rtol=1e-12)


def test_single_rank_pipeline_without_process_group(port):
    sys.path.insert(0, ROOT)
    from genpk_b200.distributed import SlabPipeline
    from tests.slab_numpy_stages import NumpyStages
    dims, n, box = 16, 3000, 10.0
    pos, _ = _particles(n, box, 5, False)
    pipe = SlabPipeline(dims, NumpyStages(dims, 1, 0))
    p, c, k = pipe.pk(torch.from_numpy(pos.reshape(-1).copy()), None, 1.0, box, float(n), dims)
    _, pr, cr, kr = port.pk(box, dims, pos, None, 1.0, float(n), dims)
    assert np.array_equal(c, cr)
    nz = cr > 0
    np.testing.assert_allclose(p[nz], pr[nz], rtol=1e-9)
