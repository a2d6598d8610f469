"""Lattice-sweep deposit (deposit_sweep.cu) against the CPU oracle.

The sweep kernel merges neighbouring particles' contributions in registers (y), shared memory
(x) and by one shuffle (z), and -- when it follows genpk_grid_zero -- clears the grid ahead of its
own front instead of a memset ("zero ahead"), leaving particles beyond the cleared front to a
clean-up kernel.  Every merge is validated per lane and every particle is deposited exactly once
whatever the window, so the grid must equal the reference's for ANY input, ANY lattice hint and
ANY window.  Fixed-point mode makes that check bit-exact: a contribution counted twice, dropped,
or wiped by a late zero store changes an integer."""
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


def poison(ctx):
    """Fill the grid with garbage so that a plane the zero-ahead sweep forgot to clear shows."""
    ctx.grid_upload(np.full(ctx.grid_doubles(), 1.0e300))


# (n_side, dims): one particle per cell (every hand-over fires), two cells per particle (none does),
# non-power-of-two, row shorter than a warp, rows of 32 and 33 (segment edges), two particles per cell
@pytest.mark.parametrize("n_side,dims", [(48, 48), (32, 64), (40, 40), (20, 24), (32, 32), (33, 33), (63, 64), (64, 64), (64, 32)])
@pytest.mark.parametrize("hint", ["right", "probe", "wrong", "none"])
@pytest.mark.parametrize("za", [0, 1])
def test_sweep_fixed_point_bit_exact(port, n_side, dims, hint, za):
    box = 1000.0
    dpos, pos = synth(api.SYNTH_CLUSTERED, n_side, dims, box)
    n = n_side ** 3
    want = fixed_want(port, box, dims, pos, None, 0.75)
    with gp.Context(dims, flags=api.FLAG_FIXED_POINT) as ctx:
        ctx.set_deposit_mode(api.DEPOSIT_SWEEP)
        ctx.set_option(api.OPT_ZERO_AHEAD, za)
        ctx.set_option(api.OPT_SWEEP_RX, 0 if za else 8)     # zero ahead needs the persistent sweep
        if hint == "right":
            ctx.set_lattice_hint(n_side, n_side)
        elif hint == "wrong":
            ctx.set_lattice_hint(n_side + 5, 7)          # still exact, just fewer merges
        elif hint == "none":
            ctx.set_lattice_hint(n, 1)                   # one long row: z hand-overs only
        poison(ctx)
        ctx.grid_zero()
        ctx.deposit_dev(dpos.data_ptr(), n, 0, 0.75, box)
        got = ctx.grid_download_fixed()
        ctx.synchronize()
        sw = ctx.last_sweep()
    assert np.array_equal(got, want), f"{(got != want).sum()} cells differ ({sw})"
    if za == 0:
        assert sw["zero_ahead"] == 0


@pytest.mark.parametrize("window,slack,defcap", [(1, 0, 4096), (1, 2, 4096), (2, 1, 4096), (3, 3, 4096), (40, 2, 4096),
                                                 (1, 1, 1), (2, 2, 3)])
@pytest.mark.parametrize("fixed", [True, False])
def test_zero_ahead_windows_and_deferred_particles(port, window, slack, defcap, fixed):
    """Windows far smaller than the displacements (rms 2 cells): most of the deposit goes through the
    deferred list -- and, with a list capacity of a few entries, through the overflow rescan."""
    n_side = dims = 96
    box = 500.0
    dpos, pos = synth(api.SYNTH_CLUSTERED, n_side, dims, box)
    n = n_side ** 3
    with gp.Context(dims, flags=api.FLAG_FIXED_POINT if fixed else 0) as ctx:
        ctx.set_deposit_mode(api.DEPOSIT_SWEEP)
        ctx.set_lattice_hint(n_side, n_side)
        ctx.set_option(api.OPT_SWEEP_RX, 0)
        ctx.set_option(api.OPT_ZA_WINDOW, window)
        ctx.set_option(api.OPT_ZA_SLACK, slack)
        ctx.set_option(api.OPT_ZA_DEFERRED, defcap)
        poison(ctx)
        ctx.grid_zero()
        ctx.deposit_dev(dpos.data_ptr(), n, 0, 1.0, box)
        sw = ctx.last_sweep()
        assert sw["zero_ahead"] == 1 and sw["window"] == window, sw
        if fixed:
            got = ctx.grid_download_fixed()
            assert np.array_equal(got, fixed_want(port, box, dims, pos, None, 1.0))
        else:
            want = np.zeros(padded_shape(dims))
            port.fieldize(box, dims, want, pos, None, 1.0, 1)
            assert_grid_close(ctx.grid_download(), want)
        ctx.synchronize()


@pytest.mark.parametrize("couple,zero_ctas,za", [(0, 0, 1), (1, 1, 1), (3, 7, 1), (40, 200, 1), (0, 0, 0), (1, 0, 0), (5, 0, 0)])
def test_sweep_coupling_and_zero_ctas(port, couple, zero_ctas, za):
    """The sweep's warps wait for each other (coupling) and for the CTAs that clear planes ahead of them;
    any setting of the two must give the same sums."""
    n_side = dims = 72
    box = 300.0
    dpos, pos = synth(api.SYNTH_CLUSTERED, n_side, dims, box)
    want = fixed_want(port, box, dims, pos, None, 1.0)
    with gp.Context(dims, flags=api.FLAG_FIXED_POINT) as ctx:
        ctx.set_deposit_mode(api.DEPOSIT_SWEEP)
        ctx.set_lattice_hint(n_side, n_side)
        ctx.set_option(api.OPT_SWEEP_RX, 0)
        ctx.set_option(api.OPT_SWEEP_COUPLE, couple)
        ctx.set_option(api.OPT_SWEEP_COUPLE_STEP, 1 + couple % 3)
        ctx.set_option(api.OPT_ZA_ZERO_CTAS, zero_ctas)
        ctx.set_option(api.OPT_ZERO_AHEAD, za)
        poison(ctx)
        ctx.grid_zero()
        ctx.deposit_dev(dpos.data_ptr(), n_side ** 3, 0, 1.0, box)
        got = ctx.grid_download_fixed()
        ctx.synchronize()
        assert ctx.last_sweep()["zero_ahead"] == za
    assert np.array_equal(got, want)


@pytest.mark.parametrize("ry", [1, 2, 3, 7, 10, 21])
@pytest.mark.parametrize("rx", [0, 1, 3, 8, 100])
def test_sweep_column_heights_and_task_lengths(port, ry, rx):
    n_side, dims, box = 44, 44, 100.0
    dpos, pos = synth(api.SYNTH_CLUSTERED, n_side, dims, box)
    want = fixed_want(port, box, dims, pos, None, 1.0)
    with gp.Context(dims, flags=api.FLAG_FIXED_POINT) as ctx:
        ctx.set_deposit_mode(api.DEPOSIT_SWEEP)
        ctx.set_lattice_hint(n_side, n_side)
        ctx.set_option(api.OPT_SWEEP_RY, ry)
        ctx.set_option(api.OPT_SWEEP_RX, rx)
        poison(ctx)
        ctx.grid_zero()
        ctx.deposit_dev(dpos.data_ptr(), n_side ** 3, 0, 1.0, box)
        got = ctx.grid_download_fixed()
        ctx.synchronize()
        assert ctx.last_sweep()["ry"] >= min(ry, 21)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("fixed", [False, True])
def test_sweep_random_particles_masses_and_chunks(port, fixed):
    """Incoherent input, per-particle masses, out-of-box and non-finite positions, a ragged tail, and
    additive calls that cut the array at arbitrary points (the first one on a pending zero)."""
    import torch
    rng = np.random.default_rng(3)
    dims, n, box = 32, 70001, 10.0
    pos = ((rng.random((n, 3)) * 1.3 - 0.15) * box).astype(np.float32)
    masses = (10.0 ** rng.uniform(-2, 1, n)).astype(np.float32)
    flags = api.FLAG_FIXED_POINT if fixed else 0
    dp = torch.from_numpy(pos.reshape(-1).copy()).cuda()
    dm = torch.from_numpy(masses).cuda()
    with gp.Context(dims, flags=flags) as ctx:
        ctx.set_deposit_mode(api.DEPOSIT_SWEEP)
        ctx.set_lattice_hint(97, 13)
        poison(ctx)
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


def test_sweep_rejects_non_finite_positions(port):
    import torch
    n_side = dims = 40
    box = 40.0
    dpos, pos = synth(api.SYNTH_CLUSTERED, n_side, dims, box)
    pos = pos.copy()
    bad = [5, 1234, 40 * 40 * 7 + 31, 40 * 40 * 7 + 32, n_side ** 3 - 1]
    pos[bad[0], 0] = np.nan
    pos[bad[1], 1] = np.inf
    pos[bad[2], 2] = -np.inf
    pos[bad[3], 0] = 3.0e38
    pos[bad[4], 1] = np.nan
    keep = np.ones(len(pos), bool)
    keep[bad] = False
    want = fixed_want(port, box, dims, pos[keep], None, 1.0)
    dp = torch.from_numpy(pos.reshape(-1).copy()).cuda()
    with gp.Context(dims, flags=api.FLAG_FIXED_POINT) as ctx:
        ctx.set_deposit_mode(api.DEPOSIT_SWEEP)
        ctx.set_lattice_hint(n_side, n_side)
        ctx.grid_zero()
        ctx.deposit_dev(dp.data_ptr(), len(pos), 0, 1.0, box)
        got = ctx.grid_download_fixed()
        assert ctx.take_rejected() == len(bad)
    assert np.array_equal(got, want)


def test_sweep_fp64_vs_oracle_and_auto(port):
    """fp64 accumulation (tolerance 1e-6) through AUTO with GENPK_OPT_SWEEP: the probe must find the lattice
    and AUTO must pick the sweep kernel."""
    n_side, dims, box = 128, 128, 250.0
    dpos, pos = synth(api.SYNTH_CLUSTERED, n_side, dims, box)
    n = n_side ** 3
    want = np.zeros(padded_shape(dims))
    port.fieldize(box, dims, want, pos, None, 1.0, 1)
    with gp.Context(dims) as ctx:
        ctx.set_option(api.OPT_SWEEP, 1)                     # AUTO: the sweep kernel instead of the march kernel
        poison(ctx)
        ctx.grid_zero()
        ctx.deposit_dev(dpos.data_ptr(), n, 0, 1.0, box)
        got = ctx.grid_download()
        ctx.synchronize()
        o, sw = ctx.last_order(), ctx.last_sweep()
    assert o["lattice"] == 1 and o["n0"] == n_side and o["n1"] == n_side, o
    assert sw["columns"] > 0, sw                             # AUTO took the sweep kernel
    assert_grid_close(got, want)
    assert abs(got.sum() - n) <= 1e-10 * n


def test_lazy_zero_is_an_immediate_zero_to_every_reader(port):
    """genpk_grid_zero is carried out lazily; whoever looks at the grid next must see zeros."""
    dims = 24
    with gp.Context(dims) as ctx:
        poison(ctx)
        ctx.grid_zero()
        assert not ctx.grid_download().any()
        poison(ctx)
        ctx.grid_zero()
        ctx.fft()
        assert not ctx.grid_download().any()
        poison(ctx)
        ctx.grid_zero()
        p, c, k = ctx.fft_power(dims, 1.0, 1.0)
        assert not p.any()
        ctx.synchronize()


def test_additive_deposit_after_zero_ahead(port):
    """Stars into baryons (gen-pk.cpp:228-230): a second deposit without a zero in between adds."""
    n_side = dims = 64
    box = 64.0
    dpos, pos = synth(api.SYNTH_CLUSTERED, n_side, dims, box)
    n = n_side ** 3
    want = fixed_want(port, box, dims, np.concatenate([pos, pos[: n // 3]]), None, 1.0)
    with gp.Context(dims, flags=api.FLAG_FIXED_POINT) as ctx:
        ctx.set_deposit_mode(api.DEPOSIT_SWEEP)
        ctx.set_lattice_hint(n_side, n_side)
        ctx.set_option(api.OPT_SWEEP_RX, 0)
        poison(ctx)
        ctx.grid_zero()
        ctx.deposit_dev(dpos.data_ptr(), n, 0, 1.0, box)
        assert ctx.last_sweep()["zero_ahead"] == 1
        ctx.deposit_dev(dpos.data_ptr(), n // 3, 0, 1.0, box)
        got = ctx.grid_download_fixed()
        ctx.synchronize()
    assert np.array_equal(got, want)


@pytest.mark.parametrize("P,ghost", [(2, 0), (4, 0), (2, 6), (4, 4)])
@pytest.mark.parametrize("za", [0, 1])
def test_sweep_in_slab_contexts(port, P, ghost, za):
    """x-slab contexts: each rank sweeps its own slab of the lattice (plain slabs: one ghost plane on the
    high-x side, sub-cell displacements; wide-ghost slabs: Zel'dovich displacements reaching into the
    ghosts on both sides); slabs + ghost exchange reassemble the single-GPU grid."""
    import torch
    from genpk_b200.distributed import CudaStages
    n_side = dims = 48
    box = 48.0
    dev = torch.device("cuda", 0)
    if ghost == 0:
        dpos, pos = synth(api.SYNTH_LATTICE, n_side, dims, box)
        rng = np.random.default_rng(1)
        pos = (pos + rng.uniform(-0.45, 0.45, pos.shape)).astype(np.float32)
        pos = np.clip(pos, 0.01, box - 0.01).astype(np.float32)
    else:
        dpos, pos = synth(api.SYNTH_CLUSTERED, n_side, 24, box)          # rms ~2 cells of a 24-grid = ~4 cells here
    want = fixed_want(port, box, dims, pos, None, 1.0).reshape(padded_shape(dims))
    per = n_side ** 3 // P
    nx = dims // P
    st = [CudaStages(dims, P, r, dev, api.FLAG_FIXED_POINT, ghost) for r in range(P)]
    try:
        rejected = 0
        for r in range(P):
            st[r].ctx.set_deposit_mode(api.DEPOSIT_SWEEP)
            st[r].ctx.set_option(api.OPT_ZERO_AHEAD, za)
            st[r].ctx.set_option(api.OPT_SWEEP_RX, 0 if za else 8)
            st[r].ctx.set_lattice_hint(n_side, n_side)
            st[r].ctx.grid_upload(np.full(st[r].ctx.grid_doubles(), 1.0e300))
            st[r].zero()
            shard = torch.from_numpy(pos[r * per:(r + 1) * per].reshape(-1).copy()).to(dev)
            st[r].deposit(shard, None, 1.0, box)
            rejected += st[r].rejected()
            assert st[r].ctx.last_sweep()["zero_ahead"] == za
        if ghost and rejected:
            pytest.skip(f"{rejected} particles beyond the ghosts: the pipeline would route instead")
        up = [st[r].ghost_plane(0, 1).clone() for r in range(P)]
        down = [st[r].ghost_plane(0, 0).clone() for r in range(P)] if ghost else None
        for r in range(P):
            st[(r + 1) % P].ghost_accumulate(up[r], 0, 0)
            if ghost:
                st[(r - 1) % P].ghost_accumulate(down[r], 0, 1)
        glo = ghost
        for r in range(P):
            part = st[r].ctx.grid_download_fixed().reshape(-1, dims, 2 * (dims // 2 + 1))
            assert np.array_equal(part[glo:glo + nx], want[r * nx:(r + 1) * nx]), f"rank {r}"
            st[r].check()
    finally:
        for s in st:
            s.close()


def test_fixed_mode_empty_rank_keeps_its_neighbours_ghost_mass(port):
    """A slab rank that deposits nothing must still add its neighbour's int64 ghost planes as integers
    (the grid representation comes from the context's mode, not from whether a deposit ran)."""
    import torch
    from genpk_b200.distributed import CudaStages
    dims, P, box = 32, 2, 32.0
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(5)
    # every particle in rank 0's slab, many of them in its last plane (their clouds reach rank 1's first plane)
    pos = np.stack([rng.uniform(14.0, 15.999, 4000), rng.uniform(0, box, 4000), rng.uniform(0, box, 4000)], axis=1).astype(np.float32)
    want = fixed_want(port, box, dims, pos, None, 1.0).reshape(padded_shape(dims))
    st = [CudaStages(dims, P, r, dev, api.FLAG_FIXED_POINT) for r in range(P)]
    try:
        for r in range(P):
            st[r].zero()
        st[0].deposit(torch.from_numpy(pos.reshape(-1).copy()).to(dev), None, 1.0, box)
        ghosts = [st[r].ghost_plane().clone() for r in range(P)]
        for r in range(P):
            st[(r + 1) % P].ghost_accumulate(ghosts[r])
        for r in range(P):
            part = st[r].ctx.grid_download_fixed().reshape(dims // P + 1, dims, 2 * (dims // 2 + 1))
            assert np.array_equal(part[:-1], want[r * dims // P:(r + 1) * dims // P]), f"rank {r}"
        assert want[dims // P].any(), "the test must put mass into rank 1's first plane"
        # and the conversion to double before the FFT sees integers on the empty rank too
        st[1].fft_yz()
        spec = st[1].ctx.grid_download().reshape(dims // P + 1, dims, 2 * (dims // 2 + 1))
        assert abs(spec[0, 0, 0] - want[dims // P][:, :dims].sum() / 2.0 ** 40) < 1e-9 * max(1.0, abs(spec[0, 0, 0]))
    finally:
        for s in st:
            s.close()


@pytest.mark.parametrize("n_side,dims", [(210, 210), (256, 256)])
def test_host_chunks_of_a_lattice_bit_exact(n_side, dims):
    """genpk_deposit on HOST particles uploads 2^23-particle chunks and plans the deposit once, on the
    first chunk; with a lattice the later chunks are cut on lattice-plane boundaries.  Fixed-point sums
    must equal the one-launch device deposit (which uses zero ahead; the chunks cannot)."""
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
