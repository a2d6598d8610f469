"""CPU-only checks of the product's host side: the C-ABI library loads and
exports every symbol the header declares, and the plan-time host arithmetic
(window table, bin edges) and the kept host utilities agree with the oracle.
No device compute is called here."""
import ctypes
import math
import os
import re

import numpy as np
import pytest

import genpk_b200 as gp
from genpk_b200 import _lib, api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module", autouse=True)
def built():
    if not os.path.exists(gp.LIB_PATH):
        gp.build()
    return gp.load()


def header_functions():
    text = open(os.path.join(ROOT, "include", "genpk_cuda.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(genpk_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(built):
    names = header_functions()
    assert len(names) >= 30
    raw = ctypes.CDLL(gp.LIB_PATH)
    for n in names:
        assert hasattr(raw, n), f"{n} declared in include/genpk_cuda.h but not exported"
    # and the Python binding table covers the header exactly
    assert sorted(_lib.SIGNATURES) == names


def test_no_cuda_device_fails_loudly(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(gp.GenPKError):
        gp.Context(16)
    field = np.zeros(2 * 4 * 4 * 3)
    with pytest.raises(gp.GenPKError):
        gp.fieldize(10.0, 4, field, 1, np.zeros(3, np.float32), None, 1.0, 1)


def test_null_handles_are_errors_not_crashes(built):
    """Every handle entry point of the fused / peer-store path refuses a null context with an
    error code and a message (no GPU needed: nothing is launched)."""
    import ctypes as C
    lib = gp.load()
    buf = C.create_string_buffer(64)
    out = (C.c_double * 4)()
    cnt = (C.c_int * 4)()
    assert lib.genpk_fft_power(None, 0, 4, out, cnt, out, 1.0, 1.0) != 0
    assert lib.genpk_deposit_f64(None, 0, out, None, 1, 1.0, 1.0, 0) != 0
    assert lib.genpk_slab_fft_yz_scatter(None, 0) != 0
    assert lib.genpk_slab_fftx_power_partial(None, out, 4, out) != 0
    assert lib.genpk_ipc_export(None, buf) != 0
    assert lib.genpk_slab_set_peer(None, 0, buf, None) != 0
    assert not lib.genpk_slab_recv_buffer(None, None)
    assert lib.genpk_fused_xpass_supported(None, 4) == 0
    assert lib.genpk_slab_scatter_supported(None) == 0
    assert gp.last_error() if hasattr(gp, "last_error") else True


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "genpk_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp", "Makefile")):
                src = open(os.path.join(dirpath, f), errors="replace").read()
                for pat in (r"import\s+oracle", r"from\s+oracle", r"oracle[/.]\w*\.(so|py|c)\b", r"libgenpk_oracle",
                            r"libgenpk_ref", r"#include\s+\"[^\"]*oracle", r"\bdlopen\b"):
                    assert not re.search(pat, src), f"{f} reaches for the oracle ({pat})"


def test_invwindow_known_answers():
    # test.cpp:52-57
    def near(x, y):
        return abs(x - y) <= max(abs(x), abs(y)) / 1e5
    assert near(gp.invwindow(0, 3, 4, 5), 71.8177719)
    assert near(gp.invwindow(4, 4, 4, 5), 6111.20801)
    assert gp.invwindow(1, 1, 1, 0) == 0


@pytest.mark.parametrize("which", ["port", "ref"])
def test_invwindow_bit_exact_vs_oracle(request, which):
    orc = request.getfixturevalue(which)
    rng = np.random.default_rng(1)
    for n in (4, 5, 32, 100, 512, 1024, 2048, 3072):
        ks = [(0, 0, 0), (n // 2, n // 2, n // 2), (1, 0, -1)]
        ks += [tuple(int(v) for v in rng.integers(-(n // 2) + 1, n // 2 + 1, 3)) for _ in range(200)]
        for k in ks:
            assert gp.invwindow(*k, n) == orc.invwindow(*k, n), (k, n)


def test_window_table_every_k(ref):
    for n in (32, 512, 1024):
        for k in range(n // 2 + 1):
            assert gp.invwindow(k, 0, 0, n) == ref.invwindow(k, 0, 0, n)


def _counts_from_thresholds(dims, nrbins, thresh):
    """Mode counts implied by the bin edges, enumerating the half-complex grid."""
    half = dims // 2
    idx = np.arange(dims)
    kv = np.where(idx <= half, idx, idx - dims).astype(np.int64)
    kz = np.arange(half + 1, dtype=np.int64)
    k2 = (kv[:, None, None] ** 2 + kv[None, :, None] ** 2 + kz[None, None, :] ** 2)
    mult = np.where((kz == 0) | (kz == half), 1, 2)[None, None, :] * np.ones_like(k2)
    sel = k2 > 0
    bins = np.searchsorted(thresh.astype(np.int64), k2[sel], side="right") - 1
    return np.bincount(bins, weights=mult[sel], minlength=nrbins).astype(np.int64)


@pytest.mark.parametrize("dims,nrbins", [(4, 10), (5, 5), (32, 32), (64, 64), (128, 128), (256, 256)])
def test_bin_edges_reproduce_reference_counts(ref, dims, nrbins):
    thresh = api.bin_thresholds(dims, nrbins)
    assert thresh[0] == 1 and thresh[nrbins] == 3 * (dims // 2) ** 2 + 1
    assert np.all(np.diff(thresh.astype(np.int64)) >= 0)
    spec = np.zeros((dims, dims, dims // 2 + 1), np.complex128)
    _, _, c_ref, _ = ref.powerspectrum(dims, spec, None, nrbins, 1.0, 1.0)
    got = _counts_from_thresholds(dims, nrbins, thresh)
    assert np.array_equal(got, c_ref.astype(np.int64))


@pytest.mark.parametrize("dims", [512, 1024, 2048, 3072])
def test_bin_edges_match_oracle_expression_at_every_edge(port, dims):
    """For the large grids of BASELINE.json: each edge k2 and its predecessor fall
    in the bins the oracle's own expression gives."""
    nrbins = dims
    thresh = api.bin_thresholds(dims, nrbins).astype(np.int64)
    k2max = 3 * (dims // 2) ** 2
    for b in range(nrbins):
        t = int(thresh[b])
        if t > k2max:
            continue
        assert port.bin_of_k2(dims, nrbins, t) >= b
        if t > 1:
            assert port.bin_of_k2(dims, nrbins, t - 1) < b
    assert port.bin_of_k2(dims, nrbins, k2max) < nrbins


def test_bin_rule_source_flag():
    a = api.bin_thresholds(64, 64, 0)
    b = api.bin_thresholds(64, 64, api.FLAG_BINRULE_SOURCE)
    assert a.shape == b.shape and a[0] == b[0] == 1
    assert np.abs(a.astype(np.int64) - b.astype(np.int64)).max() <= 1    # same rule up to libm rounding at an edge


def test_power_finalize_matches_reference_normalisation():
    nrbins = 6
    sums = np.zeros(3 * nrbins)
    sums[:nrbins] = [8.0, 0.0, 3.0, 1.5, 0.0, 9.0]
    sums[nrbins:2 * nrbins] = [6.0, 0.0, 12 * math.sqrt(2), 5.0, 0.0, 2.0]
    sums[2 * nrbins:] = [6, 0, 12, 3, 0, 1]
    p, c, k = api.power_finalize(sums, nrbins, 4.0, 2.0)
    assert list(c) == [6, 0, 12, 3, 0, 1]
    for b in range(nrbins):
        if c[b]:
            assert p[b] == sums[b] / (4.0 * 2.0) / c[b]          # two divisions, powerspectrum.c:104-105
            assert k[b] == sums[nrbins + b] / c[b]
        else:
            assert p[b] == 0 and k[b] == 0


def test_host_utilities(tmp_path):
    # test.cpp:102-113
    assert gp.type_str(0) == "by" and gp.type_str(1) == "DM" and gp.type_str(2) == "nu"
    assert gp.type_str(4) == "st" and gp.type_str(3) == "xx" and gp.type_str(5) == "xx"
    assert gp.nexttwo(12) == 16 and gp.nexttwo(8) == 8 and gp.nexttwo(1) == 1
    # gen-pk.cpp:169-172 on the bundled snapshot's particle counts
    assert gp.grid_dims_for([4039, 4096, 0, 0, 57, 0]) == 32
    assert gp.grid_dims_for([0, 1024 ** 3, 0, 0, 0, 0]) == 2048
    assert gp.grid_dims_for([0, 2048 ** 3, 0, 0, 0, 0]) == 3072
    f = tmp_path / "PK-DM-x"
    gp.print_pk(str(f), 3, np.array([1.0, 2.0, 3.5]), np.array([0.5, 9.0, 1e-7]), np.array([6, 0, 12], np.int32))
    assert f.read_text() == "1.000000e+00\t5.000000e-01\t6\n3.500000e+00\t1.000000e-07\t12\n"


def test_print_pk_matches_reference_bytes(tmp_path):
    import ctypes as C
    from oracle.oracle import REF_SO, have_reference
    if not have_reference():
        pytest.skip("oracle/_ref not built")
    lib = C.CDLL(REF_SO)
    rng = np.random.default_rng(5)
    n = 40
    keffs = rng.random(n) * 100
    power = 10.0 ** rng.uniform(-14, 3, n)
    count = rng.integers(0, 5000, n).astype(np.int32)
    count[::7] = 0
    a, b = tmp_path / "ours", tmp_path / "theirs"
    gp.print_pk(str(a), n, keffs, power, count)
    lib.ref_print_pk(str(b).encode(), n, keffs.ctypes.data_as(C.c_void_p), power.ctypes.data_as(C.c_void_p),
                     count.ctypes.data_as(C.c_void_p))
    assert a.read_bytes() == b.read_bytes()


def test_dropin_exports_the_reference_link_symbols(built):
    """libgenpk_dropin.so must define exactly what gen-pk's other objects import from
    fieldize.o, powerspectrum.o and libfftw3 (gen-pk.h:93-119; gen-pk.cpp:176-193,233,363)."""
    import subprocess
    path = os.path.join(ROOT, "genpk_b200", "libgenpk_dropin.so")
    assert os.path.exists(path)
    out = subprocess.run(["nm", "-D", "--defined-only", path], capture_output=True, text=True, check=True).stdout
    defined = {line.split()[-1] for line in out.splitlines() if line.strip()}
    for sym in ("_Z8fieldizediPdlPfS0_di", "invwindow", "powerspectrum", "fftw_malloc", "fftw_free", "fftw_init_threads",
                "fftw_plan_with_nthreads", "fftw_plan_dft_r2c_3d", "fftw_execute", "fftw_destroy_plan"):
        assert sym in defined, sym
    refhost = os.path.join(ROOT, "oracle", "_ref", "libgenpk_refhost.so")
    if os.path.exists(refhost):
        und = subprocess.run(["nm", "-D", "--undefined-only", refhost], capture_output=True, text=True, check=True).stdout
        assert "_Z8fieldizediPdlPfS0_di" in und      # the reference's reader really imports fieldize()


def test_rebin_min_modes_is_the_mode_weighted_merge():
    """genpk_rebin_min_modes (the reference's to-do "rebinning for min modes/bin", gen-pk.cpp:27-31): every output bin
    holds at least min_modes modes, nothing is lost, and the merged values are the mode-weighted means."""
    from genpk_b200 import api
    rng = np.random.default_rng(11)
    nrbins = 64
    count = rng.integers(0, 40, nrbins).astype(np.int32)
    count[[3, 4, 17]] = 0
    power = np.where(count > 0, rng.random(nrbins) * 10, 0.0)
    keffs = np.where(count > 0, np.sort(rng.random(nrbins) * 100), 0.0)
    for min_modes in (1, 25, 100, 10 ** 6):
        p, c, k = api.rebin_min_modes(power, count, keffs, min_modes)
        assert c.sum() == count.sum()
        assert len(c) >= 1 and (c[:-1] >= min(min_modes, count.sum())).all() or len(c) == 1
        if len(c) > 1:
            assert (c >= min_modes).all()
        np.testing.assert_allclose((p * c).sum(), (power * count).sum(), rtol=1e-13)
        np.testing.assert_allclose((k * c).sum(), (keffs * count).sum(), rtol=1e-13)
        assert (np.diff(k) > 0).all()
    p, c, k = api.rebin_min_modes(power, count, keffs, 1)
    nz = count > 0
    assert np.array_equal(c, count[nz]) and np.array_equal(p, power[nz]) and np.array_equal(k, keffs[nz])
    # by hand: bins of 10, 5, 20 modes with at least 12 per bin -> (10+5), then 20
    p, c, k = api.rebin_min_modes([1.0, 4.0, 2.0], [10, 5, 20], [1.0, 2.0, 3.0], 12)
    assert list(c) == [15, 20]
    np.testing.assert_allclose(p, [(10 * 1.0 + 5 * 4.0) / 15, 2.0])
    np.testing.assert_allclose(k, [(10 * 1.0 + 5 * 2.0) / 15, 3.0])
    # a short tail joins the last full bin
    p, c, k = api.rebin_min_modes([1.0, 4.0, 2.0], [10, 5, 3], [1.0, 2.0, 3.0], 12)
    assert list(c) == [18]
