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
            total += st[s].power_partial(spec, None, nrbins)
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


def _nccl_worker(rank, world, port, dims, n, box, out_dir):
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
