"""Pins the CPU oracle (oracle/genpk_oracle.c) to (1) the known answers of the
reference's own test.cpp and (2) the reference's object code in oracle/_ref."""
import math

import numpy as np
import pytest

from oracle.oracle import padded_shape, rfftn_padded


def near(x, y, rel=1e-5):
    # FLOATS_NEAR_TO, test.cpp:25-26
    return abs(x - y) <= max(abs(x), abs(y)) * rel


BACKENDS = ["port", "ref"]


@pytest.fixture(params=BACKENDS)
def orc(request):
    return request.getfixturevalue(request.param)


def test_check_fieldize(orc):
    # test.cpp:31-50
    dims = 5
    field = np.zeros(2 * dims * dims * (dims // 2 + 1))
    pos = (np.arange(30) / 3.0).astype(np.float32)
    masses = np.full(30, 10.0, np.float32)
    orc.fieldize(10, dims, field, pos, masses, 10.0, 1)
    assert near(field[0], 8.61111)
    assert field[3] == 0 and field[20] == 0 and field[125] == 0
    assert near(field[124], 1.66666)
    assert near(field[0], 8.61111088, 1e-8) and near(field[124], 1.66666468, 1e-8)


def test_check_invwindow(orc):
    # test.cpp:52-57
    assert near(orc.invwindow(0, 3, 4, 5), 71.8177719)
    assert near(orc.invwindow(4, 4, 4, 5), 6111.20801)
    assert orc.invwindow(1, 1, 1, 0) == 0


def _ps_field():
    field = np.zeros(2 * 4 * 4 * 3)
    for i in range(32):
        field[6 * (i // 4) + i % 4] = 1
    field[0] = 2
    return field


@pytest.mark.parametrize("fft", ["pocketfft", "naive"])
def test_check_powerspectrum(orc, port, fft):
    # test.cpp:59-86 (FFTW3 replaced by pocketfft / a direct DFT)
    field = _ps_field()
    spec = rfftn_padded(field, 4) if fft == "pocketfft" else port.naive_r2c_3d(4, field)
    rc, pw, count, keffs = orc.powerspectrum(4, spec, None, 10, 64.0, 64.0)
    assert rc == 0
    assert near(keffs[2], math.sqrt(2))
    assert count[2] == 12 and count[1] == 0 and count[0] == 6
    assert near(pw[0], 0.0677526)
    assert abs(pw[1]) < 1e-12
    assert near(pw[2], 0.000565561)
    assert near(pw[9], 0.0550908)
    # full-precision values recorded in SURVEY App. B
    assert list(count) == [6, 0, 12, 8, 0, 15, 12, 9, 0, 1]
    exp_p = [0.067752553897088813, 0.00056556067269081668, 0.00086079319063426922, 0.0021070631414486104,
             0.0034431720417974223, 0.012198117621116564, 0.055090752668758756]
    for b, e in zip([0, 2, 3, 5, 6, 7, 9], exp_p):
        assert near(pw[b], e, 1e-10)


def test_port_matches_reference_fieldize(port, ref):
    rng = np.random.default_rng(7)
    for dims, n, box in [(8, 1000, 25.0), (32, 20000, 3000.0), (33, 5000, 1.0)]:
        pos = (rng.random((n, 3)) * box * 1.2 - 0.1 * box).astype(np.float32)   # incl. out-of-box
        masses = rng.random(n).astype(np.float32)
        for mm in (None, masses):
            for extra in (0, 1) if dims % 2 == 0 else (1,):
                a = np.zeros(dims * dims * 2 * (dims // 2 + extra))
                b = np.zeros_like(a)
                port.fieldize(box, dims, a, pos, mm, 0.37, extra)
                ref.fieldize(box, dims, b, pos, mm, 0.37, extra)
                np.testing.assert_allclose(a, b, rtol=1e-12, atol=1e-14 * np.abs(b).max())


def test_port_matches_reference_invwindow(port, ref):
    for n in (4, 5, 32, 512, 3072):
        for k in [(0, 0, 0), (1, 0, 0), (n // 2, n // 2, n // 2), (-(n // 2) + 1, 3 % n, 1), (n // 3, -(n // 4), n // 5)]:
            assert port.invwindow(*k, n) == ref.invwindow(*k, n)


@pytest.mark.parametrize("dims,nrbins", [(4, 10), (8, 8), (32, 32), (64, 64), (128, 128)])
def test_port_matches_reference_powerspectrum(port, ref, dims, nrbins):
    rng = np.random.default_rng(dims)
    spec = rng.standard_normal((dims, dims, dims // 2 + 1)) + 1j * rng.standard_normal((dims, dims, dims // 2 + 1))
    spec2 = rng.standard_normal(spec.shape) + 1j * rng.standard_normal(spec.shape)
    for other in (None, spec2):
        _, p1, c1, k1 = port.powerspectrum(dims, spec, other, nrbins, 3.5, 2.25)
        _, p2, c2, k2 = ref.powerspectrum(dims, spec, other, nrbins, 3.5, 2.25)
        assert np.array_equal(c1, c2)
        assert c1.sum() == dims ** 3 - 1
        np.testing.assert_allclose(p1, p2, rtol=1e-11, atol=1e-300)
        np.testing.assert_allclose(k1, k2, rtol=1e-12)


def test_mode_counts_match_reference(port, ref):
    for dims in (32, 256):
        spec = np.zeros((dims, dims, dims // 2 + 1), np.complex128)
        _, _, c_ref, k_ref = ref.powerspectrum(dims, spec, None, dims, 1.0, 1.0)
        c, ks = port.mode_counts(dims, dims)
        assert np.array_equal(c, c_ref.astype(np.int64))
        nz = c > 0
        np.testing.assert_allclose(ks[nz] / c[nz], k_ref[nz], rtol=1e-12)


def test_fixed_point_restatement(port):
    rng = np.random.default_rng(3)
    dims, n, box, S = 16, 5000, 100.0, 40
    pos = (rng.random((n, 3)) * box).astype(np.float32)
    q = np.zeros(padded_shape(dims), np.int64)
    port.fieldize_fixed(box, dims, q, pos, None, 1.0, 1, S)
    f = np.zeros(padded_shape(dims))
    port.fieldize(box, dims, f, pos, None, 1.0, 1)
    g = port.fixed_to_double(q, S)
    np.testing.assert_allclose(g, f, rtol=0, atol=n * 2.0 ** -S)
    # order independence: permuting the particles gives the same integers
    q2 = np.zeros_like(q)
    port.fieldize_fixed(box, dims, q2, pos[rng.permutation(n)], None, 1.0, 1, S)
    assert np.array_equal(q, q2)
    assert abs(g.sum() - n) < 8 * n * 2.0 ** -S


def test_double_precision_restatement_against_the_reference_built_with_double_precision_snap(port):
    """oracle_fieldize_fixed_f64 (the parity target of GENPK_OPT_F64_POSITIONS=1) against fieldize.cpp compiled
    unmodified with -DDOUBLE_PRECISION_SNAP: positions with full double mantissas, some outside the box."""
    from oracle.oracle import have_reference_f64, padded_shape, ref_fieldize_f64
    if not have_reference_f64():
        pytest.skip("oracle/_ref/libgenpk_ref_f64.so not built (needs /root/reference)")
    rng = np.random.default_rng(21)
    dims, box, n = 24, 50.0, 20000
    pos = (rng.random((n, 3)) * 1.2 - 0.1) * box
    masses = (10.0 ** rng.uniform(-1, 1, n)).astype(np.float32)
    for m in (None, masses):
        want = np.zeros(padded_shape(dims))
        assert ref_fieldize_f64(box, dims, want, pos, m, 0.7, 1) == 0
        q = np.zeros(padded_shape(dims), np.int64)
        port.fieldize_fixed_f64(box, dims, q, pos, m, 0.7, 1, 40)
        got = port.fixed_to_double(q, 40)
        np.testing.assert_allclose(got, want, rtol=1e-9, atol=1e-9 * want.mean())
        # and it is NOT what narrowing the positions to float first gives
        narrowed = np.zeros(padded_shape(dims), np.int64)
        port.fieldize_fixed(box, dims, narrowed, pos.astype(np.float32), m, 0.7, 1, 40)
        assert not np.array_equal(narrowed, q)
