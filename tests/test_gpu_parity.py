"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle,
the reference's known answers (test.cpp) and the committed golden vectors.

Tolerances (BASELINE.json north_star): mode counts bit-exact; density grid
relative 1e-6 (bit-exact in fixed-point mode); P(k) and k_eff relative 1e-5."""
import math
import os

import numpy as np
import pytest

import genpk_b200 as gp
from genpk_b200 import api
from oracle.oracle import padded_shape, rfftn_padded

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GRID_RTOL = 1e-6
PK_RTOL = 1e-5


def near(x, y, rel=1e-5):
    return abs(x - y) <= max(abs(x), abs(y)) * rel       # FLOATS_NEAR_TO, test.cpp:25-26


def assert_grid_close(got, want):
    """|g-r| <= 1e-6*|r| + 1e-6*mean(|r|) per cell (parity protocol 2)."""
    got, want = np.asarray(got).reshape(-1), np.asarray(want).reshape(-1)
    floor = GRID_RTOL * np.abs(want).mean()
    bad = np.abs(got - want) > GRID_RTOL * np.abs(want) + floor
    assert not bad.any(), f"{bad.sum()} cells differ; worst {np.abs(got - want).max()}"


def assert_pk_close(got, want):
    (p, c, k), (pr, cr, kr) = got, want
    assert np.array_equal(c, cr), "mode counts differ"
    nz = cr > 0
    np.testing.assert_allclose(p[nz], pr[nz], rtol=PK_RTOL, atol=0)
    np.testing.assert_allclose(k[nz], kr[nz], rtol=PK_RTOL, atol=0)
    assert np.all(p[~nz] == 0) and np.all(k[~nz] == 0)


@pytest.fixture(scope="module")
def orc(request):
    """The reference's object code when it was built, else the pinned port."""
    from oracle.oracle import Oracle, have_reference
    request.getfixturevalue("port")
    return Oracle("reference") if have_reference() else Oracle("port")


# ----------------------------------------------------------------------------------
# the reference's own known answers, through the reference-signature shims
# ----------------------------------------------------------------------------------
def test_check_fieldize():
    # test.cpp:31-50
    dims = 5
    field = np.zeros(2 * dims * dims * (dims // 2 + 1))
    pos = (np.arange(30) / 3.0).astype(np.float32)
    masses = np.full(30, 10.0, np.float32)
    assert gp.fieldize(10, dims, field, 10, pos, masses, 10.0, 1) == 0
    assert near(field[0], 8.61111)
    assert field[3] == 0 and field[20] == 0 and field[125] == 0
    assert near(field[124], 1.66666)


def test_check_powerspectrum():
    # test.cpp:59-86, with cuFFT where the reference calls FFTW
    field = np.zeros(2 * 4 * 4 * 3)
    for i in range(32):
        field[6 * (i // 4) + i % 4] = 1
    field[0] = 2
    gp.r2c_3d(4, field)
    pw, count, keffs = np.empty(10), np.empty(10, np.int32), np.empty(10)
    assert gp.powerspectrum(4, field, field, 10, pw, count, keffs, 64.0, 64.0) == 0
    assert near(keffs[2], math.sqrt(2))
    assert count[2] == 12 and count[1] == 0 and count[0] == 6
    assert near(pw[0], 0.0677526) and abs(pw[1]) < 1e-12
    assert near(pw[2], 0.000565561) and near(pw[9], 0.0550908)
    assert list(count) == [6, 0, 12, 8, 0, 15, 12, 9, 0, 1]


def test_check_read_fieldize_values():
    # test.cpp:88-100 at the fieldize() boundary: the arrays the reference's reader
    # produced (golden), deposited by the GPU, baryons then stars into one field.
    g = np.load(os.path.join(GOLD, "test_g2_snap.npz"))
    field = np.zeros(padded_shape(4))
    gp.fieldize(3000.0, 4, field, len(g["pos0"]), g["pos0"], g["masses0"], 0.0, 1)
    gp.fieldize(3000.0, 4, field, len(g["pos4"]), g["pos4"], g["masses4"], 0.0, 1)
    tm = float(g["rf4_total_mass"])
    f = field.reshape(-1)
    assert abs(f[10] / tm) < 1e-5
    assert near(f[0] / tm, 0.026873) and near(f[15] / tm, 0.0188683)
    assert_grid_close(field, g["rf4_grid"])


# ----------------------------------------------------------------------------------
# golden vectors produced by the reference's object code
# ----------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["a", "b", "c"])
def test_golden_fieldize(name):
    g = np.load(os.path.join(GOLD, "fieldize_cases.npz"))
    dims, box, pos, masses = int(g[f"{name}_dims"]), float(g[f"{name}_box"]), g[f"{name}_pos"], g[f"{name}_masses"]
    out = np.zeros(padded_shape(dims))
    gp.fieldize(box, dims, out, len(pos), pos, None, 0.75, 1)
    assert_grid_close(out, g[f"{name}_grid_const"])
    out = np.zeros(padded_shape(dims))
    gp.fieldize(box, dims, out, len(pos), pos, masses, 0.0, 1)
    assert_grid_close(out, g[f"{name}_grid_var"])


@pytest.mark.parametrize("name", ["s16", "s12", "s32"])
def test_golden_powerspectrum(name):
    g = np.load(os.path.join(GOLD, "powerspectrum_cases.npz"))
    dims, nrbins = int(g[f"{name}_dims"]), int(g[f"{name}_nrbins"])
    a, b = np.ascontiguousarray(g[f"{name}_a"]), np.ascontiguousarray(g[f"{name}_b"])
    p, c, k = np.empty(nrbins), np.empty(nrbins, np.int32), np.empty(nrbins)
    gp.powerspectrum(dims, a, a, nrbins, p, c, k, 3.0, 3.0)
    assert_pk_close((p, c, k), (g[f"{name}_power"], g[f"{name}_count"], g[f"{name}_keffs"]))
    gp.powerspectrum(dims, a, b, nrbins, p, c, k, 3.0, 5.0)
    assert np.array_equal(c, g[f"{name}_xcount"])
    np.testing.assert_allclose(p, g[f"{name}_xpower"], rtol=PK_RTOL, atol=1e-9 * np.abs(g[f"{name}_xpower"]).max())


@pytest.mark.parametrize("ptype", [0, 1, 4])
def test_golden_snapshot_pk(ptype):
    """BASELINE config 1 (test_g2_snap, grid 32) at the fieldize()/powerspectrum() boundary."""
    g = np.load(os.path.join(GOLD, "test_g2_snap.npz"))
    pos = g[f"pos{ptype}"]
    masses = g[f"masses{ptype}"] if f"masses{ptype}" in g.files else None
    mass = float(g["mass"][ptype])
    tm = float(g[f"total_mass{ptype}"])
    with gp.Context(32) as ctx:
        ctx.grid_zero()
        ctx.deposit(pos, masses, mass, float(g["box"]))
        grid = ctx.grid_download()
        assert_grid_close(grid, g[f"grid{ptype}"])
        ctx.fft()
        got = ctx.power(32, tm, tm)
        ctx.synchronize()
    assert_pk_close(got, (g[f"power{ptype}"], g[f"count{ptype}"], g[f"keffs{ptype}"]))


# ----------------------------------------------------------------------------------
# differential tests against the oracle on seeded inputs
# ----------------------------------------------------------------------------------
@pytest.mark.parametrize("dims,n,box,spread", [(8, 1000, 25.0, 1.0), (32, 20000, 3000.0, 1.5), (33, 5000, 1.0, 1.0),
                                               (64, 200000, 100.0, 1.0), (128, 300000, 1000.0, 1.0)])
@pytest.mark.parametrize("var_mass", [False, True])
def test_fieldize_vs_oracle(orc, dims, n, box, spread, var_mass):
    rng = np.random.default_rng(dims * 7 + n)
    pos = ((rng.random((n, 3)) * spread - (spread - 1) / 2) * box).astype(np.float32)    # incl. out-of-box wrap
    masses = (10.0 ** rng.uniform(-2, 1, n)).astype(np.float32) if var_mass else None
    for extra in ((0, 1) if dims % 2 == 0 else (1,)):
        want = np.zeros(dims * dims * 2 * (dims // 2 + extra))
        got = np.zeros_like(want)
        orc.fieldize(box, dims, want, pos, masses, 0.37, extra)
        gp.fieldize(box, dims, got, n, pos, masses, 0.37, extra)
        assert_grid_close(got, want)
        total = masses.astype(np.float64).sum() if var_mass else 0.37 * n
        assert abs(got.sum() - total) <= 1e-12 * total * 8


def test_fieldize_edge_cases(orc):
    dims, box = 16, 16.0
    out = np.zeros(padded_shape(dims))
    gp.fieldize(box, dims, out, 0, np.zeros(0, np.float32), None, 1.0, 1)          # empty input
    assert not out.any()
    # exactly on cell corners, on the box edge, negative, and far outside (periodic wrap)
    pos = np.array([[0, 0, 0], [16, 16, 16], [15.5, 15.5, 15.5], [-0.25, 3, 3], [-16.0, -32.0, 48.0],
                    [1e-7, 7.9999995, 15.999999], [100.3, -57.9, 0.5]], np.float32)
    want = np.zeros(padded_shape(dims))
    orc.fieldize(box, dims, want, pos, None, 2.0, 1)
    gp.fieldize(box, dims, out, len(pos), pos, None, 2.0, 1)
    assert_grid_close(out, want)
    # accumulation into a non-zero grid (chunked callers, read_fieldize.cpp:51-93)
    gp.fieldize(box, dims, out, len(pos), pos, None, 2.0, 1)
    assert_grid_close(out, 2 * want)


@pytest.mark.parametrize("mode", [api.DEPOSIT_DIRECT, api.DEPOSIT_SORTED, api.DEPOSIT_AUTO])
@pytest.mark.parametrize("var_mass", [False, True])
def test_fixed_point_bit_exact(port, mode, var_mass):
    rng = np.random.default_rng(11)
    dims, n, box, S = 64, 150000, 50.0, 40
    pos = ((rng.random((n, 3)) * 1.2 - 0.1) * box).astype(np.float32)
    masses = (10.0 ** rng.uniform(-3, 0, n)).astype(np.float32) if var_mass else None
    want = np.zeros(padded_shape(dims), np.int64)
    port.fieldize_fixed(box, dims, want, pos, masses, 1.0, 1, S)
    with gp.Context(dims, flags=api.FLAG_FIXED_POINT) as ctx:
        ctx.set_deposit_mode(mode)
        ctx.set_scale_bits(S)
        ctx.grid_zero()
        ctx.deposit(pos[: n // 3], None if masses is None else masses[: n // 3], 1.0, box)    # additive calls
        ctx.deposit(pos[n // 3:], None if masses is None else masses[n // 3:], 1.0, box)
        got = ctx.grid_download_fixed()
        assert np.array_equal(got, want.reshape(-1)), "fixed-point grid is not bit-exact"
        dbl = ctx.grid_download()
        assert np.array_equal(dbl, port.fixed_to_double(want, S).reshape(-1))
        # the same particles in another order give the same integers
        perm = rng.permutation(n)
        ctx.grid_zero()
        ctx.deposit(pos[perm], None if masses is None else masses[perm], 1.0, box)
        assert np.array_equal(ctx.grid_download_fixed(), want.reshape(-1))
        ctx.synchronize()


@pytest.mark.parametrize("dims", [64, 96])
def test_sorted_deposit_equals_direct(orc, dims):
    rng = np.random.default_rng(5)
    n, box = 400000, 10.0
    pos = (rng.random((n, 3)) * box).astype(np.float32)
    masses = rng.random(n).astype(np.float32)
    want = np.zeros(padded_shape(dims))
    orc.fieldize(box, dims, want, pos, masses, 0.0, 1)
    grids = {}
    for mode in (api.DEPOSIT_DIRECT, api.DEPOSIT_SORTED):
        with gp.Context(dims) as ctx:
            ctx.set_deposit_mode(mode)
            ctx.grid_zero()
            ctx.deposit(pos, masses, 0.0, box)
            grids[mode] = ctx.grid_download()
            ctx.synchronize()
        assert_grid_close(grids[mode], want)
    np.testing.assert_allclose(grids[api.DEPOSIT_DIRECT], grids[api.DEPOSIT_SORTED], rtol=1e-11, atol=1e-13)


@pytest.mark.parametrize("dims", [4, 5, 16, 30, 64, 96])
def test_fft_vs_pocketfft(dims):
    rng = np.random.default_rng(dims)
    field = np.zeros(padded_shape(dims))
    field[:, :, :dims] = rng.standard_normal((dims, dims, dims))
    want = rfftn_padded(field, dims)
    buf = field.reshape(-1).copy()
    gp.r2c_3d(dims, buf)
    got = buf.view(np.complex128).reshape(dims, dims, dims // 2 + 1)
    np.testing.assert_allclose(got, want, rtol=0, atol=1e-11 * np.abs(want).max())


@pytest.mark.parametrize("dims,nrbins", [(4, 10), (5, 5), (8, 8), (30, 30), (32, 32), (64, 64), (64, 17), (128, 128),
                                         (160, 160)])
def test_powerspectrum_vs_oracle(orc, dims, nrbins):
    rng = np.random.default_rng(dims + nrbins)
    shape = (dims, dims, dims // 2 + 1)
    a = rng.standard_normal(shape) + 1j * rng.standard_normal(shape)
    b = rng.standard_normal(shape) + 1j * rng.standard_normal(shape)
    for other, tm2 in ((None, 3.5), (b, 2.25)):
        _, pr, cr, kr = orc.powerspectrum(dims, a, other, nrbins, 3.5, tm2)
        p, c, k = np.empty(nrbins), np.empty(nrbins, np.int32), np.empty(nrbins)
        gp.powerspectrum(dims, a, a if other is None else other, nrbins, p, c, k, 3.5, tm2)
        assert np.array_equal(c, cr)
        if dims % 2 == 0:
            assert c.sum() == dims ** 3 - 1
        scale = np.abs(pr).max()
        np.testing.assert_allclose(p, pr, rtol=PK_RTOL if other is None else 1e-4, atol=1e-10 * scale)
        np.testing.assert_allclose(k, kr, rtol=1e-12)


@pytest.mark.parametrize("kind", ["uniform", "clustered"])
@pytest.mark.parametrize("fixed", [False, True])
def test_pipeline_pk_vs_oracle(orc, kind, fixed):
    """deposit -> FFT -> binning end to end on 64^3 particles / 128^3 grid."""
    import torch
    n_side, dims, box = 64, 128, 1000.0
    n = n_side ** 3
    dpos = torch.empty(n * 3, dtype=torch.float32, device="cuda")
    api.synth_particles_dev(api.SYNTH_UNIFORM_RANDOM if kind == "uniform" else api.SYNTH_CLUSTERED, 42, n_side, 0, n,
                            box, dims, dpos.data_ptr())
    torch.cuda.synchronize()
    pos = dpos.cpu().numpy()
    field, pr, cr, kr = orc.pk(box, dims, pos, None, 1.0, float(n), dims)
    with gp.Context(dims, flags=api.FLAG_FIXED_POINT if fixed else 0) as ctx:
        ctx.grid_zero()
        ctx.deposit_dev(dpos.data_ptr(), n, 0, 1.0, box)
        assert_grid_close(ctx.grid_download(), field)
        ctx.fft()
        got = ctx.power(dims, float(n), float(n))
        ctx.synchronize()
        assert_pk_close(got, (pr, cr, kr))
        # the one-call host-array entry point gives the same answer
        got2 = ctx.pk_from_particles(pos, None, 1.0, box, float(n), dims)
        assert_pk_close(got2, (pr, cr, kr))


def test_cross_spectrum_two_fields(orc):
    rng = np.random.default_rng(9)
    dims, box, n = 32, 10.0, 30000
    p1 = (rng.random((n, 3)) * box).astype(np.float32)
    p2 = (p1 + rng.normal(0, 0.2, (n, 3))).astype(np.float32)
    f1, f2 = np.zeros(padded_shape(dims)), np.zeros(padded_shape(dims))
    orc.fieldize(box, dims, f1, p1, None, 1.0, 1)
    orc.fieldize(box, dims, f2, p2, None, 2.0, 1)
    _, pr, cr, kr = orc.powerspectrum(dims, rfftn_padded(f1, dims), rfftn_padded(f2, dims), dims, n, 2.0 * n)
    with gp.Context(dims, flags=api.FLAG_TWO_FIELDS) as ctx:
        ctx.grid_zero(0)
        ctx.grid_zero(1)
        ctx.deposit(p1, None, 1.0, box, which=0)
        ctx.deposit(p2, None, 2.0, box, which=1)
        ctx.fft(0)
        ctx.fft(1)
        p, c, k = ctx.power(dims, n, 2.0 * n, a=0, b=1)
        ctx.synchronize()
    assert np.array_equal(c, cr)
    np.testing.assert_allclose(p, pr, rtol=PK_RTOL, atol=1e-9 * np.abs(pr).max())


def test_error_paths():
    with pytest.raises(gp.GenPKError):
        gp.Context(30, nranks=4, rank=0)                  # dims not divisible by nranks
    with gp.Context(16) as ctx:
        with pytest.raises(gp.GenPKError):
            ctx.grid_zero(1)                               # second field was not requested
        ctx.grid_zero()
        bad = np.array([[np.nan, 1, 1], [1, 1, 1]], np.float32)
        ctx.deposit(bad, None, 1.0, 16.0)
        with pytest.raises(gp.GenPKError):
            ctx.synchronize()                              # non-finite particle is reported, not deposited
        assert abs(ctx.grid_download().sum() - 1.0) < 1e-12


# ----------------------------------------------------------------------------------
# synthetic generators
# ----------------------------------------------------------------------------------
def _mix64(z):
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def test_synth_generators():
    import torch
    n_side, box, dims = 32, 1000.0, 64
    n = n_side ** 3
    d = torch.empty(n * 3, dtype=torch.float32, device="cuda")
    api.synth_particles_dev(api.SYNTH_UNIFORM_RANDOM, 42, n_side, 0, n, box, dims, d.data_ptr())
    u = d.cpu().numpy()
    with np.errstate(over="ignore"):
        s = _mix64(np.uint64(42) + np.uint64(0x9E3779B97F4A7C15)) * np.uint64(0x9E3779B97F4A7C15)
        h = _mix64(np.arange(3 * n, dtype=np.uint64) + s)
    want = np.float32(box) * ((h >> np.uint64(40)).astype(np.float32) * np.float32(2.0 ** -24))
    assert np.array_equal(u, want)                         # bit-identical host mirror
    # any sub-range reproduces the same particles
    api.synth_particles_dev(api.SYNTH_UNIFORM_RANDOM, 42, n_side, 1000, 500, box, dims, d.data_ptr())
    assert np.array_equal(d[:1500].cpu().numpy(), want[3000:4500])
    api.synth_particles_dev(api.SYNTH_LATTICE, 42, n_side, 0, n, box, dims, d.data_ptr())
    lat = d.cpu().numpy().reshape(n_side, n_side, n_side, 3)
    q = ((np.arange(n_side) + 0.5) * (box / n_side)).astype(np.float32)
    assert np.allclose(lat[..., 2], q[None, None, :], rtol=1e-6) and np.allclose(lat[..., 0], q[:, None, None], rtol=1e-6)
    api.synth_particles_dev(api.SYNTH_CLUSTERED, 42, n_side, 0, n, box, dims, d.data_ptr())
    cl = d.cpu().numpy().reshape(n_side, n_side, n_side, 3)
    disp = (cl - lat + box / 2) % box - box / 2
    rms_cells = np.sqrt((disp ** 2).mean(axis=(0, 1, 2))) / (box / dims)
    assert np.all(rms_cells > 1.0) and np.all(rms_cells < 3.5), rms_cells
    assert cl.min() >= 0 and cl.max() <= box


# ----------------------------------------------------------------------------------
# full BASELINE size (config 2: 256^3 particles -> 512^3 grid) through
# size-independent properties; the oracle only supplies the mode-count table
# ----------------------------------------------------------------------------------
def test_full_size_properties():
    import torch
    n_side, dims, box = 256, 512, 1000.0
    n = n_side ** 3
    gold = np.load(os.path.join(GOLD, "mode_counts.npz"))
    dpos = torch.empty(n * 3, dtype=torch.float32, device="cuda")
    api.synth_particles_dev(api.SYNTH_UNIFORM_RANDOM, 42, n_side, 0, n, box, dims, dpos.data_ptr())
    results = {}
    for label, flags, mode in (("direct", 0, api.DEPOSIT_DIRECT), ("sorted", 0, api.DEPOSIT_SORTED),
                               ("fixed", api.FLAG_FIXED_POINT, api.DEPOSIT_AUTO)):
        with gp.Context(dims, flags=flags) as ctx:
            ctx.set_deposit_mode(mode)
            ctx.grid_zero()
            ctx.deposit_dev(dpos.data_ptr(), n, 0, 1.0, box)
            grid = ctx.grid_download().reshape(padded_shape(dims))
            assert abs(grid[:, :, :dims].sum() - n) <= 1e-9 * n          # mass conservation
            assert not grid[:, :, dims:].any()                           # FFTW padding untouched
            assert grid.min() >= 0
            ctx.fft()
            p, c, k = ctx.power(dims, float(n), float(n))
            ctx.synchronize()
            assert np.array_equal(c.astype(np.int64), gold["count512"])  # bit-exact mode counts
            assert c.sum() == dims ** 3 - 1
            nz = c > 0
            np.testing.assert_allclose(k[nz], gold["ksum512"][nz] / gold["count512"][nz], rtol=1e-12)
            results[label] = p
    # Poisson sample of a uniform field: P(k) ~ 1/N in box-volume units once the CIC window is deconvolved
    # (low k only: towards the Nyquist frequency aliasing adds power)
    kmean = gold["ksum512"] / np.maximum(gold["count512"], 1)
    nz = (gold["count512"] > 1000) & (kmean < dims / 8)
    shot = results["direct"][nz] * n
    assert 0.9 < np.median(shot) < 1.1
    np.testing.assert_allclose(results["sorted"], results["direct"], rtol=1e-9, atol=1e-20)
    np.testing.assert_allclose(results["fixed"], results["direct"], rtol=1e-7, atol=1e-20)


@pytest.mark.parametrize("var_mass", [False, True])
def test_double_precision_positions_are_narrowed_like_the_reference(var_mass):
    """genpk_deposit_f64: f8 positions narrowed on the GPU as read_fieldize_bigfile.cpp:93-94 narrows
    them on the host ((float) of each double, round to nearest): the fixed-point grid must be
    bit-identical to narrowing with numpy and calling genpk_deposit; host and device-resident input."""
    import torch
    dims, box, n = 96, 250.0, 300000
    rng = np.random.default_rng(11)
    pos64 = (rng.random((n, 3)) * 1.1 - 0.05) * box                    # full double mantissas, some outside the box
    masses = (10.0 ** rng.uniform(-1, 1, n)).astype(np.float32) if var_mass else None
    with gp.Context(dims, flags=api.FLAG_FIXED_POINT) as ctx:
        ctx.grid_zero()
        ctx.deposit(pos64.astype(np.float32), masses, 0.5, box)
        want = ctx.grid_download_fixed()
        ctx.grid_zero()
        ctx.deposit_f64(pos64, masses, 0.5, box)
        got = ctx.grid_download_fixed()
        dpos = torch.from_numpy(pos64.reshape(-1).copy()).cuda()
        dm = torch.from_numpy(masses).cuda() if var_mass else None
        ctx.grid_zero()
        ctx.deposit_f64_dev(dpos.data_ptr(), n, dm.data_ptr() if var_mass else 0, 0.5, box)
        got_dev = ctx.grid_download_fixed()
        ctx.synchronize()
    assert want.any()
    assert np.array_equal(got, want)
    assert np.array_equal(got_dev, want)


@pytest.mark.parametrize("var_mass", [False, True])
def test_double_precision_positions_used_as_they_are(port, var_mass):
    """GENPK_OPT_F64_POSITIONS=1: genpk_deposit_f64 hands the doubles to the deposit un-narrowed, which is what a
    reference built with -DDOUBLE_PRECISION_SNAP does (gen-pk.h:25-29, read_fieldize.cpp:24-25).  Fixed-point
    grid bit-exact against the CPU restatement (pinned to that reference build in tests/test_oracle.py), fp64
    grid within 1e-6 of the reference objects themselves; host and device-resident input."""
    import torch
    from oracle.oracle import have_reference_f64, ref_fieldize_f64
    dims, box, n = 96, 250.0, 300000
    rng = np.random.default_rng(12)
    pos64 = (rng.random((n, 3)) * 1.1 - 0.05) * box
    masses = (10.0 ** rng.uniform(-1, 1, n)).astype(np.float32) if var_mass else None
    want = np.zeros(padded_shape(dims), np.int64)
    port.fieldize_fixed_f64(box, dims, want, pos64, masses, 0.5, 1, 40)
    narrowed = np.zeros(padded_shape(dims), np.int64)
    port.fieldize_fixed(box, dims, narrowed, pos64.astype(np.float32), masses, 0.5, 1, 40)
    assert not np.array_equal(want, narrowed)
    with gp.Context(dims, flags=api.FLAG_FIXED_POINT) as ctx:
        ctx.set_option(api.OPT_F64_POSITIONS, 1)
        ctx.grid_zero()
        ctx.deposit_f64(pos64, masses, 0.5, box)
        got = ctx.grid_download_fixed()
        dpos = torch.from_numpy(pos64.reshape(-1).copy()).cuda()
        dm = torch.from_numpy(masses).cuda() if var_mass else None
        ctx.grid_zero()
        ctx.deposit_f64_dev(dpos.data_ptr(), n, dm.data_ptr() if var_mass else 0, 0.5, box)
        got_dev = ctx.grid_download_fixed()
        ctx.synchronize()
    assert np.array_equal(got, want.reshape(-1))
    assert np.array_equal(got_dev, want.reshape(-1))
    if have_reference_f64():
        ref = np.zeros(padded_shape(dims))
        ref_fieldize_f64(box, dims, ref, pos64, masses, 0.5, 1)
        with gp.Context(dims) as ctx:
            ctx.set_option(api.OPT_F64_POSITIONS, 1)
            ctx.grid_zero()
            ctx.deposit_f64(pos64, masses, 0.5, box)
            g = ctx.grid_download()
            ctx.synchronize()
        tol = 1e-6 * np.abs(ref).reshape(-1) + 1e-6 * np.abs(ref).mean()
        assert np.all(np.abs(g - ref.reshape(-1)) <= tol)


def test_automatic_fixed_point_scale_follows_the_mass_unit(port):
    """GENPK_OPT_SCALE_BITS = -1: particle masses of 1e-8 (the stars of test_g2_snap) lose five digits at the default
    2^40; the automatic scale is 40 - ceil(log2(max mass)), latched at the first deposit after a zero."""
    rng = np.random.default_rng(4)
    dims, box, n = 32, 100.0, 5000
    pos = (rng.random((n, 3)) * box).astype(np.float32)
    masses = (10.0 ** rng.uniform(-9, -7.2, n)).astype(np.float32)
    want = np.zeros(padded_shape(dims))
    port.fieldize(box, dims, want, pos, masses, 0.0, 1)
    with gp.Context(dims, flags=api.FLAG_FIXED_POINT) as ctx:
        ctx.grid_zero()
        ctx.deposit(pos, masses, 0.0, box)
        coarse = ctx.grid_download()
        assert ctx.grid_scale_bits() == 40
        ctx.set_scale_bits(-1)
        ctx.grid_zero()
        ctx.deposit(pos, masses, 0.0, box)
        fine = ctx.grid_download()
        bits = ctx.grid_scale_bits()
        q = ctx.grid_download_fixed()
        ctx.synchronize()
    assert bits == 40 - int(np.ceil(np.log2(float(masses.max()))))
    ref_q = np.zeros(padded_shape(dims), np.int64)
    port.fieldize_fixed(box, dims, ref_q, pos, masses, 0.0, 1, bits)
    assert np.array_equal(q, ref_q.reshape(-1))                   # the same rule, bit for bit, at the chosen scale
    tol = 1e-9 * np.abs(want).max()
    assert np.abs(fine - want.reshape(-1)).max() <= tol
    assert np.abs(coarse - want.reshape(-1)).max() > 100 * tol    # (what the default scale does to such masses)
