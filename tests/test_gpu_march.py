"""Lattice-march deposit (deposit_march.cu) against the CPU oracle.

The march kernel merges neighbouring particles' corner contributions in registers
before they reach L2; every merge is validated per lane, so the grid must equal the
reference's for ANY input and ANY lattice hint.  Fixed-point mode makes that check
bit-exact: a contribution counted twice or dropped changes an integer."""
import numpy as np
import pytest

import genpk_b200 as gp
from genpk_b200 import api
from oracle.oracle import padded_shape

pytestmark = pytest.mark.gpu


def synth(kind, n_side, dims, box, first=0, count=None):
    import torch
    n = n_side ** 3 if count is None else count
    d = torch.empty(3 * n, dtype=torch.float32, device="cuda")
    api.synth_particles_dev(kind, 42, n_side, first, n, box, dims, d.data_ptr())
    torch.cuda.synchronize()
    return d, d.cpu().numpy().reshape(-1, 3)


def assert_grid_close(got, want):
    got, want = np.asarray(got).reshape(-1), np.asarray(want).reshape(-1)
    floor = 1e-6 * np.abs(want).mean()
    bad = np.abs(got - want) > 1e-6 * np.abs(want) + floor
    assert not bad.any(), f"{bad.sum()} cells differ; worst {np.abs(got - want).max()}"


def fixed_want(port, box, dims, pos, masses, cmass, S=40):
    want = np.zeros(padded_shape(dims), np.int64)
    port.fieldize_fixed(box, dims, want, np.ascontiguousarray(pos), masses, cmass, 1, S)
    return want.reshape(-1)


# (n_side, dims): one particle per cell (every hand-over fires), two cells per particle
# (none does), non-power-of-two, row shorter than a warp, row of 32 and 33 (segment edges)
@pytest.mark.parametrize("n_side,dims", [(48, 48), (32, 64), (40, 40), (20, 24), (32, 32), (33, 33), (63, 64), (64, 64)])
@pytest.mark.parametrize("hint", ["right", "probe", "wrong", "none"])
def test_march_fixed_point_bit_exact(port, n_side, dims, hint):
    box = 1000.0
    dpos, pos = synth(api.SYNTH_CLUSTERED, n_side, dims, box)
    n = n_side ** 3
    want = fixed_want(port, box, dims, pos, None, 0.75)
    with gp.Context(dims, flags=api.FLAG_FIXED_POINT) as ctx:
        ctx.set_deposit_mode(api.DEPOSIT_MARCH)
        if hint == "right":
            ctx.set_lattice_hint(n_side, n_side)
        elif hint == "wrong":
            ctx.set_lattice_hint(n_side + 5, 7)          # still exact, just fewer merges
        elif hint == "none":
            ctx.set_lattice_hint(n, 1)                   # one long row: z hand-overs only
        ctx.grid_zero()
        ctx.deposit_dev(dpos.data_ptr(), n, 0, 0.75, box)
        got = ctx.grid_download_fixed()
        ctx.synchronize()
    assert np.array_equal(got, want), f"{(got != want).sum()} cells differ"


@pytest.mark.parametrize("ry,rx", [(1, 1), (3, 2), (8, 8), (16, 5), (32, 64)])
def test_march_block_shapes(port, ry, rx):
    n_side, dims, box = 44, 44, 100.0
    dpos, pos = synth(api.SYNTH_CLUSTERED, n_side, dims, box)
    want = fixed_want(port, box, dims, pos, None, 1.0)
    with gp.Context(dims, flags=api.FLAG_FIXED_POINT) as ctx:
        ctx.set_deposit_mode(api.DEPOSIT_MARCH)
        ctx.set_lattice_hint(n_side, n_side)
        ctx.set_option(api.OPT_MARCH_RY, ry)
        ctx.set_option(api.OPT_MARCH_RX, rx)
        ctx.grid_zero()
        ctx.deposit_dev(dpos.data_ptr(), n_side ** 3, 0, 1.0, box)
        got = ctx.grid_download_fixed()
        ctx.synchronize()
    assert np.array_equal(got, want)


@pytest.mark.parametrize("fixed", [False, True])
def test_march_random_particles_masses_and_chunks(port, fixed):
    """Incoherent input, per-particle masses, out-of-box positions, a NaN-free ragged
    tail, and additive calls that cut the array at arbitrary points."""
    import torch
    rng = np.random.default_rng(3)
    dims, n, box = 32, 70001, 10.0
    pos = ((rng.random((n, 3)) * 1.3 - 0.15) * box).astype(np.float32)
    masses = (10.0 ** rng.uniform(-2, 1, n)).astype(np.float32)
    flags = api.FLAG_FIXED_POINT if fixed else 0
    dp = torch.from_numpy(pos.reshape(-1).copy()).cuda()
    dm = torch.from_numpy(masses).cuda()
    with gp.Context(dims, flags=flags) as ctx:
        ctx.set_deposit_mode(api.DEPOSIT_MARCH)
        ctx.set_lattice_hint(97, 13)
        ctx.grid_zero()
        cuts = [0, 1, 40, 40 + 97 * 13 * 3 + 5, 33333, n]
        for lo, hi in zip(cuts[:-1], cuts[1:]):
            ctx.deposit_dev(dp.data_ptr() + 12 * lo, hi - lo, dm.data_ptr() + 4 * lo, 0.0, box)
        if fixed:
            got = ctx.grid_download_fixed()
            assert np.array_equal(got, fixed_want(port, box, dims, pos, masses, 0.0))
        else:
            want = np.zeros(padded_shape(dims))
            port.fieldize(box, dims, want, pos, masses, 0.0, 1)
            assert_grid_close(ctx.grid_download(), want)
        ctx.synchronize()


def test_march_fp64_vs_oracle_and_auto(port):
    """fp64 accumulation (tolerance 1e-6) through AUTO: the probe must find the lattice."""
    n_side, dims, box = 128, 128, 250.0          # smooth enough at this size: ~60 % of hand-overs fire
    dpos, pos = synth(api.SYNTH_CLUSTERED, n_side, dims, box)
    n = n_side ** 3
    want = np.zeros(padded_shape(dims))
    port.fieldize(box, dims, want, pos, None, 1.0, 1)
    with gp.Context(dims) as ctx:
        ctx.grid_zero()
        ctx.deposit_dev(dpos.data_ptr(), n, 0, 1.0, box)
        got = ctx.grid_download()
        ctx.synchronize()
        o = ctx.last_order()
    assert o["lattice"] == 1 and o["n0"] == n_side and o["n1"] == n_side, o
    assert_grid_close(got, want)
    assert abs(got.sum() - n) <= 1e-10 * n


def test_march_probe_finds_rows_without_a_cube(port):
    """A lattice whose particle count is no cube (and no hint): the row length comes
    from the backward jumps of z along the array."""
    import torch
    n0, n1, n2, dims, box = 80, 50, 20, 80, 80.0
    p = np.arange(n0 * n1 * n2)
    iz, iy, ix = p % n0, (p // n0) % n1, p // (n0 * n1)          # p = (ix*n1 + iy)*n0 + iz
    # unit spacing, smooth sub-cell displacement
    pos = np.stack([ix + 0.5 + 0.3 * np.sin(iy / 9.0), iy + 0.5 + 0.3 * np.cos(iz / 7.0), iz + 0.5 + 0.3 * np.sin(ix / 5.0)],
                   axis=1).astype(np.float32)
    n = len(pos)
    want = fixed_want(port, box, dims, pos, None, 1.0)
    dp = torch.from_numpy(pos.reshape(-1).copy()).cuda()
    with gp.Context(dims, flags=api.FLAG_FIXED_POINT) as ctx:
        ctx.grid_zero()
        ctx.deposit_dev(dp.data_ptr(), n, 0, 1.0, box)
        got = ctx.grid_download_fixed()
        ctx.synchronize()
        o = ctx.last_order()
    assert o["lattice"] == 1 and o["n0"] == n0, o
    assert np.array_equal(got, want)


@pytest.mark.parametrize("P", [2, 4])
def test_march_in_slab_contexts(port, P):
    """x-slab contexts (ghost plane on the high-x side): each rank marches over its own
    slab of the lattice; slabs + ghost ring shift reassemble the single-GPU grid."""
    import torch
    from genpk_b200.distributed import CudaStages
    n_side = dims = 48
    box = 48.0
    dev = torch.device("cuda", 0)
    # sub-cell displacement keeps every particle inside its own rank's slab
    dpos, pos = synth(api.SYNTH_LATTICE, n_side, dims, box)
    rng = np.random.default_rng(1)
    pos = (pos + rng.uniform(-0.45, 0.45, pos.shape)).astype(np.float32)
    pos = np.clip(pos, 0.01, box - 0.01).astype(np.float32)
    want = fixed_want(port, box, dims, pos, None, 1.0).reshape(padded_shape(dims))
    per = n_side ** 3 // P
    st = [CudaStages(dims, P, r, dev, api.FLAG_FIXED_POINT) for r in range(P)]
    try:
        for r in range(P):
            st[r].ctx.set_deposit_mode(api.DEPOSIT_MARCH)
            st[r].ctx.set_lattice_hint(n_side, n_side)
            st[r].zero()
            shard = torch.from_numpy(pos[r * per:(r + 1) * per].reshape(-1).copy()).to(dev)
            st[r].deposit(shard, None, 1.0, box)
        ghosts = [st[r].ghost_plane().clone() for r in range(P)]
        for r in range(P):
            st[(r + 1) % P].ghost_accumulate(ghosts[r])
        for r in range(P):
            part = st[r].ctx.grid_download_fixed().reshape(dims // P + 1, dims, 2 * (dims // 2 + 1))
            assert np.array_equal(part[:-1], want[r * dims // P:(r + 1) * dims // P])
            st[r].check()
    finally:
        for s in st:
            s.close()


@pytest.mark.parametrize("n_side,dims", [(210, 210), (256, 256)])
def test_host_chunks_of_a_lattice_bit_exact(n_side, dims):
    """genpk_deposit on HOST particles uploads 2^23-particle chunks and plans the deposit once,
    on the first chunk; with a lattice the later chunks are cut on lattice-plane boundaries
    (210^2 does not divide 2^23).  Fixed-point sums must equal the one-launch device deposit."""
    box = 640.0
    dpos, pos = synth(api.SYNTH_CLUSTERED, n_side, dims, box)
    n = n_side ** 3
    assert n > (1 << 23)
    with gp.Context(dims, flags=api.FLAG_FIXED_POINT) as ctx:
        ctx.grid_zero()
        ctx.deposit_dev(dpos.data_ptr(), n, 0, 1.0, box)
        want = ctx.grid_download_fixed()
        ctx.grid_zero()
        ctx.deposit(pos, None, 1.0, box)                 # numpy array = host buffer
        got = ctx.grid_download_fixed()
        ctx.synchronize()
        o = ctx.last_order()
    assert o["lattice"] == 1 and o["n0"] == n_side, o
    assert np.array_equal(got, want)
