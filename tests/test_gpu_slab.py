"""Slab (multi-GPU) stages of the C ABI.

* On ONE GPU: P slab contexts live on the same device and the three exchange
  steps are done by hand with tensor copies, so the routing, ghost, pack and
  transposed-binning kernels are checked wherever a single B200 is available.
* On >= 2 GPUs: the real thing, one process per GPU over NCCL."""
import os
import socket
import sys

import numpy as np
import pytest

import genpk_b200 as gp
from genpk_b200 import api
from oracle.oracle import padded_shape

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _particles(n, box, seed, var_mass):
    rng = np.random.default_rng(seed)
    pos = ((rng.random((n, 3)) * 1.1 - 0.05) * box).astype(np.float32)
    masses = (10.0 ** rng.uniform(-1, 1, n)).astype(np.float32) if var_mass else None
    return pos, masses


@pytest.mark.parametrize("P,dims,var_mass,fixed", [(2, 32, False, False), (4, 64, True, False), (2, 64, True, True),
                                                  (8, 64, False, True)])
def test_slab_stages_emulated_on_one_gpu(port, P, dims, var_mass, fixed):
    import torch
    from genpk_b200.distributed import CudaStages
    dev = torch.device("cuda", 0)
    n, box, nrbins = 50000, 250.0, dims
    pos, masses = _particles(n, box, 77, var_mass)
    tm = float(masses.astype(np.float64).sum()) if var_mass else 0.5 * n
    flags = api.FLAG_FIXED_POINT if fixed else 0
    st = [CudaStages(dims, P, r, dev, flags) for r in range(P)]
    try:
        # route every rank's shard, then hand each run to its owner (the all-to-all-v)
        runs = [[None] * P for _ in range(P)]
        mruns = [[None] * P for _ in range(P)]
        for r in range(P):
            lo, hi = r * n // P, (r + 1) * n // P
            dp = torch.from_numpy(pos[lo:hi].reshape(-1).copy()).to(dev)
            dm = torch.from_numpy(masses[lo:hi].copy()).to(dev) if var_mass else None
            spos, smass, counts = st[r].route(dp, dm, box)
            c = counts.cpu().tolist()
            assert sum(c) == hi - lo
            off = np.concatenate([[0], np.cumsum(c)])
            for s in range(P):
                runs[r][s] = spos[3 * off[s]: 3 * off[s + 1]]
                mruns[r][s] = smass[off[s]: off[s + 1]] if var_mass else None
        for s in range(P):
            st[s].zero()
            rp = torch.cat([runs[r][s] for r in range(P)])
            rm = torch.cat([mruns[r][s] for r in range(P)]) if var_mass else None
            st[s].deposit(rp, rm, 0.5, box)
        # ghost ring shift
        ghosts = [st[r].ghost_plane().clone() for r in range(P)]
        for r in range(P):
            st[(r + 1) % P].ghost_accumulate(ghosts[r])
        # the assembled real grid equals the single-GPU deposit
        with gp.Context(dims, flags=flags) as one:
            one.grid_zero()
            one.deposit(pos, masses, 0.5, box)
            if fixed:
                whole = one.grid_download_fixed().reshape(padded_shape(dims))
                for r in range(P):
                    part = st[r].ctx.grid_download_fixed().reshape(dims // P + 1, dims, 2 * (dims // 2 + 1))
                    assert np.array_equal(part[:-1], whole[r * dims // P: (r + 1) * dims // P]), "slab grid not bit-exact"
            else:
                whole = one.grid_download().reshape(padded_shape(dims))
                for r in range(P):
                    part = st[r].ctx.grid_download().reshape(dims // P + 1, dims, 2 * (dims // 2 + 1))
                    np.testing.assert_allclose(part[:-1], whole[r * dims // P: (r + 1) * dims // P], rtol=1e-11, atol=1e-13)
            one.fft()
            p1, c1, k1 = one.power(nrbins, tm, tm)
            one.synchronize()
        # FFT with the transpose done by hand
        nx = ny = dims // P
        nc = dims // 2 + 1
        blk = nx * ny * nc * 2
        sends = []
        for r in range(P):
            st[r].fft_yz()
            sends.append(st[r].pack().clone())
        total = torch.zeros(3 * nrbins, dtype=torch.float64, device=dev)
        for s in range(P):
            spec = st[s].spectrum_buffer()
            for r in range(P):
                spec[r * blk: (r + 1) * blk] = sends[r][s * blk: (s + 1) * blk]
            st[s].fft_x(spec)
            total += st[s].power_partial(spec, None, nrbins)[:3 * nrbins]
            st[s].check()
        p, c, k = api.power_finalize(total.cpu().numpy(), nrbins, tm, tm)
        assert np.array_equal(c, c1)
        nz = c1 > 0
        np.testing.assert_allclose(p[nz], p1[nz], rtol=1e-9)
        np.testing.assert_allclose(k[nz], k1[nz], rtol=1e-12)
        # and against the CPU oracle
        _, pr, cr, kr = port.pk(box, dims, pos, masses, 0.5, tm, nrbins)
        assert np.array_equal(c, cr)
        np.testing.assert_allclose(p[nz], pr[nz], rtol=1e-5)
    finally:
        for s in st:
            s.close()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _nccl_worker(rank, world, port, dims, n, box, out_dir, scatter=False):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from genpk_b200.distributed import CudaStages, SlabPipeline
        pos, masses = _particles(n, box, 99, True)
        lo, hi = rank * n // world, (rank + 1) * n // world
        tm = float(masses.astype(np.float64).sum())
        dev = torch.device("cuda", rank)
        stages = CudaStages(dims, world, rank, dev)
        pipe = SlabPipeline(dims, stages)
        if scatter:
            assert stages.enable_scatter(), "CUDA IPC exchange of the transposed blocks failed"
            pipe.pk(torch.from_numpy(pos[lo:hi].reshape(-1).copy()).to(dev),        # twice: blocks are reused
                    torch.from_numpy(masses[lo:hi].copy()).to(dev), 0.0, box, tm, dims)
        p, c, k = pipe.pk(torch.from_numpy(pos[lo:hi].reshape(-1).copy()).to(dev),
                          torch.from_numpy(masses[lo:hi].copy()).to(dev), 0.0, box, tm, dims)
        stages.check()
        if rank == 0:
            np.savez(os.path.join(out_dir, "out.npz"), p=p, c=c, k=k)
        stages.close()
    finally:
        dist.destroy_process_group()


def test_slab_pipeline_over_nccl(tmp_path, port):
    import torch
    import torch.multiprocessing as mp
    world = min(torch.cuda.device_count(), 8)
    if world < 2:
        pytest.skip("needs >= 2 GPUs (run under gpurun --gpus 2)")
    dims, n, box = 128, 400000, 100.0
    mp.spawn(_nccl_worker, args=(world, _free_port(), dims, n, box, str(tmp_path)), nprocs=world, join=True)
    got = np.load(tmp_path / "out.npz")
    pos, masses = _particles(n, box, 99, True)
    tm = float(masses.astype(np.float64).sum())
    _, pr, cr, kr = port.pk(box, dims, pos, masses, 0.0, tm, dims)
    assert np.array_equal(got["c"], cr)
    nz = cr > 0
    np.testing.assert_allclose(got["p"][nz], pr[nz], rtol=1e-5)
    np.testing.assert_allclose(got["k"][nz], kr[nz], rtol=1e-5)


def test_slab_pipeline_over_nccl_with_peer_stores(tmp_path, port):
    """One process per GPU: the y pass of every rank stores into the other ranks' transposed
    blocks through CUDA IPC mappings (NVLink), then the fused x pass; against the CPU oracle."""
    import torch
    import torch.multiprocessing as mp
    world = min(torch.cuda.device_count(), 8)
    if world < 2:
        pytest.skip("needs >= 2 GPUs (run under gpurun --gpus 2)")
    dims, n, box = 256, 400000, 100.0
    mp.spawn(_nccl_worker, args=(world, _free_port(), dims, n, box, str(tmp_path), True), nprocs=world, join=True)
    got = np.load(tmp_path / "out.npz")
    pos, masses = _particles(n, box, 99, True)
    tm = float(masses.astype(np.float64).sum())
    _, pr, cr, kr = port.pk(box, dims, pos, masses, 0.0, tm, dims)
    assert np.array_equal(got["c"], cr)
    nz = cr > 0
    np.testing.assert_allclose(got["p"][nz], pr[nz], rtol=1e-5)
    np.testing.assert_allclose(got["k"][nz], kr[nz], rtol=1e-5)


@pytest.mark.parametrize("P,ghost", [(2, 6), (4, 8), (8, 4)])
def test_wide_ghost_slabs_emulated_on_one_gpu(port, P, ghost):
    """Wide-ghost slab contexts: every rank deposits its index-range shard of a displaced
    lattice as it is (no particle exchange); stragglers up to `ghost` planes outside the slab
    land in the ghost planes on either side and the two-sided ghost exchange puts them where
    they belong.  Fixed-point mode: the assembled grid equals the single-GPU grid bit for bit."""
    import torch
    from genpk_b200.distributed import CudaStages
    dev = torch.device("cuda", 0)
    n_side = dims = 64
    box = 640.0
    n = n_side ** 3
    d = torch.empty(3 * n, dtype=torch.float32, device=dev)
    api.synth_particles_dev(api.SYNTH_LATTICE, 1, n_side, 0, n, box, dims, d.data_ptr())
    torch.cuda.synchronize()
    rng = np.random.default_rng(8)
    pos = d.cpu().numpy().reshape(-1, 3)
    pos[:, 0] = np.mod(pos[:, 0] + rng.uniform(-(ghost - 1.5), ghost - 1.5, n).astype(np.float32) * np.float32(box / dims), box)
    pos[:, 1:] += rng.uniform(-3, 3, (n, 2)).astype(np.float32)
    pos = np.ascontiguousarray(pos, np.float32)
    want = np.zeros(padded_shape(dims), np.int64)
    port.fieldize_fixed(box, dims, want, pos, None, 1.0, 1, 40)
    per = n // P
    st = [CudaStages(dims, P, r, dev, api.FLAG_FIXED_POINT, ghost) for r in range(P)]
    nx = dims // P
    try:
        for r in range(P):
            st[r].zero()
            st[r].deposit(torch.from_numpy(pos[r * per:(r + 1) * per].reshape(-1).copy()).to(dev), None, 1.0, box)
            assert st[r].rejected() == 0
        up = [st[r].ghost_plane(0, 1).clone() for r in range(P)]
        down = [st[r].ghost_plane(0, 0).clone() for r in range(P)]
        for r in range(P):
            st[(r + 1) % P].ghost_accumulate(up[r], 0, 0)
            st[(r - 1) % P].ghost_accumulate(down[r], 0, 1)
        for r in range(P):
            part = st[r].ctx.grid_download_fixed().reshape(nx + 2 * ghost, dims, 2 * (dims // 2 + 1))
            assert np.array_equal(part[ghost:ghost + nx], want[r * nx:(r + 1) * nx]), f"slab {r} differs"
        # a straggler beyond the ghosts is rejected, not silently dropped or misplaced
        far = np.array([[((0 * nx + nx + ghost + 2) % dims + 0.5) * box / dims, 1.0, 1.0]], np.float32)
        st[0].deposit(torch.from_numpy(far.reshape(-1)).to(dev), None, 1.0, box)
        assert st[0].rejected() == 1
    finally:
        for s in st:
            s.close()


def _nccl_wide_worker(rank, world, port, dims, n_side, box, ghost, out_dir):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from genpk_b200.distributed import CudaStages, SlabPipeline
        n = n_side ** 3
        first, count = rank * n // world, (rank + 1) * n // world - rank * n // world
        dpos = torch.empty(3 * count, dtype=torch.float32, device=dev)
        api.synth_particles_dev(api.SYNTH_CLUSTERED, 42, n_side, first, count, box, dims, dpos.data_ptr())
        torch.cuda.synchronize()
        stages = CudaStages(dims, world, rank, dev, 0, ghost)
        pipe = SlabPipeline(dims, stages)
        p, c, k = pipe.pk(dpos, None, 1.0, box, float(n), dims)
        stages.check()
        if rank == 0:
            np.savez(os.path.join(out_dir, "out.npz"), p=p, c=c, k=k, placement=pipe.placement)
        stages.close()
    finally:
        dist.destroy_process_group()


def test_wide_ghost_pipeline_over_nccl(tmp_path, port):
    """The clustered lattice sharded by index range over the GPUs of the box, no particle
    exchange (placement stays 'local'), against the single-process oracle."""
    import torch
    import torch.multiprocessing as mp
    world = min(torch.cuda.device_count(), 8)
    if world < 2:
        pytest.skip("needs >= 2 GPUs (run under gpurun --gpus 2)")
    n_side = dims = 128
    box, ghost = 1000.0, 16
    mp.spawn(_nccl_wide_worker, args=(world, _free_port(), dims, n_side, box, ghost, str(tmp_path)), nprocs=world, join=True)
    got = np.load(tmp_path / "out.npz")
    assert str(got["placement"]) == "local"
    n = n_side ** 3
    d = torch.empty(3 * n, dtype=torch.float32, device="cuda:0")
    api.synth_particles_dev(api.SYNTH_CLUSTERED, 42, n_side, 0, n, box, dims, d.data_ptr())
    torch.cuda.synchronize()
    _, pr, cr, kr = port.pk(box, dims, d.cpu().numpy().reshape(-1, 3), None, 1.0, float(n), dims)
    assert np.array_equal(got["c"], cr)
    nz = cr > 0
    np.testing.assert_allclose(got["p"][nz], pr[nz], rtol=1e-5)
    np.testing.assert_allclose(got["k"][nz], kr[nz], rtol=1e-5)


@pytest.mark.parametrize("P,ghost,mode", [(2, 6, "sweep"), (4, 8, "sweep"), (8, 4, "march"), (2, 0, "sweep"), (4, 0, "direct")])
def test_ghost_pull_emulated_on_one_gpu(port, P, ghost, mode):
    """Ghost exchange by peer loads (genpk_ghost_pull) with the neighbours' grids mapped as plain pointers
    inside one process: every rank adds its neighbours' ghost planes into its own outermost planes, only the
    planes their deposits touched (the sweep kernel tracks them; march / direct mark every plane).
    Fixed-point mode: the assembled grid equals the single-GPU grid bit for bit."""
    import torch
    from genpk_b200.distributed import CudaStages
    dev = torch.device("cuda", 0)
    n_side = dims = 64
    box = 640.0
    n = n_side ** 3
    d = torch.empty(3 * n, dtype=torch.float32, device=dev)
    api.synth_particles_dev(api.SYNTH_LATTICE, 1, n_side, 0, n, box, dims, d.data_ptr())
    torch.cuda.synchronize()
    rng = np.random.default_rng(8)
    pos = d.cpu().numpy().reshape(-1, 3)
    reach = (ghost - 1.5) if ghost else 0.45
    pos[:, 0] = np.mod(pos[:, 0] + rng.uniform(-reach, reach, n).astype(np.float32) * np.float32(box / dims), box)
    if not ghost:
        pos[:, 0] = np.clip(pos[:, 0], 0.01, box - 0.01)
    pos[:, 1:] += rng.uniform(-3, 3, (n, 2)).astype(np.float32)
    pos = np.ascontiguousarray(pos, np.float32)
    want = np.zeros(padded_shape(dims), np.int64)
    port.fieldize_fixed(box, dims, want, pos, None, 1.0, 1, 40)
    per = n // P
    st = [CudaStages(dims, P, r, dev, api.FLAG_FIXED_POINT, ghost) for r in range(P)]
    nx = dims // P
    try:
        for r in range(P):
            st[r].ctx.set_deposit_mode({"sweep": api.DEPOSIT_SWEEP, "march": api.DEPOSIT_MARCH, "direct": api.DEPOSIT_DIRECT}[mode])
            st[r].ctx.set_lattice_hint(n_side, n_side)
            st[r].zero()
            st[r].deposit(torch.from_numpy(pos[r * per:(r + 1) * per].reshape(-1).copy()).to(dev), None, 1.0, box)
            assert st[r].rejected() == 0
        for r in range(P):
            st[r].ctx.slab_set_grid_peer(0, None, st[(r - 1) % P].ctx.grid_ptr())
            if ghost:
                st[r].ctx.slab_set_grid_peer(1, None, st[(r + 1) % P].ctx.grid_ptr())
            assert st[r].ctx.ghost_pull_ready()
        torch.cuda.synchronize()
        for r in range(P):
            st[r].ghost_pull()
        torch.cuda.synchronize()
        for r in range(P):
            part = st[r].ctx.grid_download_fixed().reshape(nx + (2 * ghost if ghost else 1), dims, 2 * (dims // 2 + 1))
            assert np.array_equal(part[ghost:ghost + nx], want[r * nx:(r + 1) * nx]), f"slab {r} differs"
    finally:
        for s in st:
            s.close()


def _nccl_pull_worker(rank, world, port, dims, n_side, box, ghost, out_dir, host_shards):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from genpk_b200.distributed import CudaStages, SlabPipeline
        n = n_side ** 3
        first, count = rank * n // world, (rank + 1) * n // world - rank * n // world
        dpos = torch.empty(3 * count, dtype=torch.float32, device=dev)
        api.synth_particles_dev(api.SYNTH_CLUSTERED, 42, n_side, first, count, box, dims, dpos.data_ptr())
        torch.cuda.synchronize()
        stages = CudaStages(dims, world, rank, dev, 0, ghost)
        pipe = SlabPipeline(dims, stages)
        assert stages.enable_scatter() and stages.enable_ghost_pull()
        shard = dpos
        if host_shards:
            shard = torch.empty(3 * count, dtype=torch.float32, pin_memory=True)
            shard.copy_(dpos)
            torch.cuda.synchronize()
        for _ in range(2):                                        # twice: blocks, ghosts and counters are reused
            p, c, k = pipe.pk(shard, None, 1.0, box, float(n), dims)
        stages.check()
        if rank == 0:
            np.savez(os.path.join(out_dir, "out.npz"), p=p, c=c, k=k, placement=pipe.placement)
        stages.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("ghost,host_shards,placement", [(16, False, "local"), (16, True, "local"), (2, False, "route")])
def test_peer_pull_pipeline_over_nccl(tmp_path, port, ghost, host_shards, placement):
    """The whole multi-GPU step with no NCCL data collective but the final all-reduce: ghost planes pulled
    from the neighbours' memory, transpose stored into the owners' blocks, rejected-particle check riding
    with the sums (ghost = 2 is too thin for the displacements: the pipeline must notice and route)."""
    import torch
    import torch.multiprocessing as mp
    world = min(torch.cuda.device_count(), 8)
    if world < 2:
        pytest.skip("needs >= 2 GPUs (run under gpurun --gpus 2)")
    n_side = dims = 256
    box = 1000.0
    mp.spawn(_nccl_pull_worker, args=(world, _free_port(), dims, n_side, box, ghost, str(tmp_path), host_shards), nprocs=world,
             join=True)
    got = np.load(tmp_path / "out.npz")
    assert str(got["placement"]) == placement
    n = n_side ** 3
    d = torch.empty(3 * n, dtype=torch.float32, device="cuda:0")
    api.synth_particles_dev(api.SYNTH_CLUSTERED, 42, n_side, 0, n, box, dims, d.data_ptr())
    torch.cuda.synchronize()
    _, pr, cr, kr = port.pk(box, dims, d.cpu().numpy().reshape(-1, 3), None, 1.0, float(n), dims)
    assert np.array_equal(got["c"], cr)
    nz = cr > 0
    np.testing.assert_allclose(got["p"][nz], pr[nz], rtol=1e-5)
    np.testing.assert_allclose(got["k"][nz], kr[nz], rtol=1e-5)
