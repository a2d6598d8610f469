"""CPU check of the fused x-pass (genpk_b200/csrc/fftx_core.cuh): the kernel's own phase
functions -- index maps, twiddles, butterflies, exchange swizzles, +-kx folding and the bin
walk -- compiled for the host (tests/fftx_emu.cpp) and run thread by thread, against
numpy.fft and a direct restatement of powerspectrum.c:56-99 on the tile."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from genpk_b200 import api

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    out = tmp_path_factory.mktemp("fftx") / "libfftx_emu.so"
    subprocess.run(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-o", str(out),
                    os.path.join(HERE, "fftx_emu.cpp")], check=True)
    lib = ctypes.CDLL(str(out))
    lib.fftx_emu_columns.restype = ctypes.c_int
    lib.fftx_emu_tile.restype = ctypes.c_int
    return lib


def _tables(n, nrbins):
    thresh = api.bin_thresholds(n, nrbins).astype(np.uint32)
    k = np.arange(n // 2 + 1, dtype=np.float64)
    with np.errstate(invalid="ignore", divide="ignore"):
        iw = np.where(k > 0, np.pi * k / (n * np.sin(np.pi * k / np.float32(n))), 1.0)
    return thresh, iw.astype(np.float32)


def _direct_bins(spec, n, kj, kz0, nc, thresh, iw, nrbins):
    """powerspectrum.c:56-99 on the tile's columns."""
    p = np.zeros(nrbins)
    ncols = spec.shape[1]
    idx = np.arange(n)
    kx = np.where(idx <= n // 2, idx, idx - n)
    for c in range(ncols):
        kz = kz0 + c
        if kz >= nc:
            continue
        mult = 1 if kz in (0, n // 2) else 2
        k2 = kx.astype(np.int64) ** 2 + kj * kj + kz * kz
        w = (iw[np.abs(kx)] * iw[abs(kj)]) * iw[kz]                 # float32 products, fieldize.cpp:129-132
        w = w.astype(np.float64) ** 4
        mod2 = spec[:, c].real ** 2 + spec[:, c].imag ** 2
        ok = k2 > 0
        b = np.searchsorted(thresh, k2[ok], side="right") - 1
        np.add.at(p, b, mult * mod2[ok] * w[ok])
    return p


@pytest.mark.parametrize("plan,nrbins", [(256, 1), (256, 37), (512, 2 * 512 + 3), (1024, 100)])
def test_tile_bins_for_other_bin_counts(emu, plan, nrbins):
    """nrbins is independent of the grid side in the C ABI (gen-pk uses nrbins = dims): one bin,
    few bins (long runs), more bins than |k| values (many empty bins)."""
    test_tile_fft_and_bins(emu, plan, -7, 16, nrbins)


@pytest.mark.parametrize("plan", [256, 512, 1024, 2048, -1024])      # -1024: the 8192-mode tile of the 1024 plan
@pytest.mark.parametrize("kj,kz0", [(0, 0), (-3, 8), (5, None)])
def test_tile_fft_and_bins(emu, plan, kj, kz0, nrbins=None):
    ncols = emu.fftx_emu_columns(plan)
    n = abs(plan)
    assert ncols * n in (4096, 8192)
    nc = n // 2 + 1
    if kz0 is None:                                              # the last column group: only the Nyquist column is valid
        kz0 = (nc // ncols) * ncols
    rng = np.random.default_rng(n + 7 * kj + kz0)
    tile = rng.standard_normal((n, ncols)) + 1j * rng.standard_normal((n, ncols))
    tile *= np.exp(rng.uniform(-6, 6, (n, 1)))                   # a wide dynamic range
    tw = np.exp(-2j * np.pi * np.arange(n) / n)
    nrbins = n if nrbins is None else nrbins
    thresh, iw = _tables(n, nrbins)
    half_bpu = np.float32(0.5 * (nrbins - 1) / np.log(np.sqrt(3.0) * n / 2.0)) if nrbins > 1 else np.float32(0)
    spec = np.zeros((n, ncols), dtype=np.complex128)
    mod2 = np.zeros((n, ncols))
    sp = np.zeros(nrbins)
    tin = np.ascontiguousarray(tile)
    twc = np.ascontiguousarray(tw)
    rc = emu.fftx_emu_tile(ctypes.c_int(plan), tin.ctypes.data_as(ctypes.c_void_p), twc.ctypes.data_as(ctypes.c_void_p),
                           spec.ctypes.data_as(ctypes.c_void_p), mod2.ctypes.data_as(ctypes.c_void_p),
                           ctypes.c_int(kj), ctypes.c_int(kz0), ctypes.c_int(nc),
                           iw.ctypes.data_as(ctypes.c_void_p), thresh.ctypes.data_as(ctypes.c_void_p),
                           ctypes.c_int(nrbins), ctypes.c_float(half_bpu), sp.ctypes.data_as(ctypes.c_void_p))
    assert rc == 0
    ref = np.fft.fft(tile, axis=0)
    scale = np.abs(ref).max()
    assert np.abs(spec - ref).max() <= 1e-13 * scale
    want = _direct_bins(ref, n, kj, kz0, nc, thresh, iw, nrbins)
    nz = want > 0
    assert np.array_equal(sp > 0, nz)
    assert np.allclose(sp[nz], want[nz], rtol=1e-11, atol=0)


@pytest.mark.parametrize("plan", [256, 512, 1024, -1024, 2048])
def test_row_tile_is_the_real_transform(emu, plan):
    """The z tile of fft_zy_kernel (half-length complex FFT + the pair untangling of rfft_pair) against numpy.fft.rfft,
    in place on padded rows."""
    emu.fftx_emu_row_count.restype = ctypes.c_int
    emu.fftx_emu_rows.restype = ctypes.c_int
    n = abs(plan)
    rows = emu.fftx_emu_row_count(plan)
    assert rows * (n // 2) in (4096, 8192)
    rng = np.random.default_rng(plan + 4096)
    x = rng.standard_normal((rows, n)) * np.exp(rng.uniform(-6, 6, (rows, 1)))
    buf = np.zeros((rows, n + 2))
    buf[:, :n] = x
    buf[:, n:] = 123.0                                           # the padding doubles are overwritten, never read
    tw = np.ascontiguousarray(np.exp(-2j * np.pi * np.arange(n) / n))
    twh = np.ascontiguousarray(np.exp(-2j * np.pi * np.arange(n // 2) / (n // 2)))
    rc = emu.fftx_emu_rows(ctypes.c_int(plan), buf.ctypes.data_as(ctypes.c_void_p), tw.ctypes.data_as(ctypes.c_void_p),
                           twh.ctypes.data_as(ctypes.c_void_p))
    assert rc == 0
    got = buf.view(np.complex128)
    ref = np.fft.rfft(x, axis=1)
    assert got.shape == ref.shape
    scale = np.abs(ref).max(axis=1, keepdims=True)
    assert (np.abs(got - ref) <= 1e-13 * scale).all()
    assert (got[:, 0].imag == 0).all() and (got[:, -1].imag == 0).all()


@pytest.mark.parametrize("n", [512, 1024])
@pytest.mark.parametrize("kj,kz0", [(0, 0), (-3, 8), (5, None)])
def test_two_pass_tile_fft_and_bins(emu, n, kj, kz0):
    """The two-pass plan (Plan2: radix 32, one exchange, 32 elements per thread) of fftx_power2_kernel against numpy.fft
    and the direct bins, tile of 8 columns."""
    emu.fftx_emu_tile2.restype = ctypes.c_int
    ncols, nc = 8, n // 2 + 1
    if kz0 is None:
        kz0 = (nc // ncols) * ncols
    rng = np.random.default_rng(3 * n + 7 * kj + kz0)
    tile = rng.standard_normal((n, ncols)) + 1j * rng.standard_normal((n, ncols))
    tile *= np.exp(rng.uniform(-6, 6, (n, 1)))
    tw = np.ascontiguousarray(np.exp(-2j * np.pi * np.arange(n) / n))
    nrbins = n
    thresh, iw = _tables(n, nrbins)
    half_bpu = np.float32(0.5 * (nrbins - 1) / np.log(np.sqrt(3.0) * n / 2.0))
    spec = np.zeros((n, ncols), dtype=np.complex128)
    sp = np.zeros(nrbins)
    tin = np.ascontiguousarray(tile)
    rc = emu.fftx_emu_tile2(ctypes.c_int(n), tin.ctypes.data_as(ctypes.c_void_p), tw.ctypes.data_as(ctypes.c_void_p),
                            spec.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(kj), ctypes.c_int(kz0), ctypes.c_int(nc),
                            iw.ctypes.data_as(ctypes.c_void_p), thresh.ctypes.data_as(ctypes.c_void_p),
                            ctypes.c_int(nrbins), ctypes.c_float(half_bpu), sp.ctypes.data_as(ctypes.c_void_p))
    assert rc == 0
    ref = np.fft.fft(tile, axis=0)
    assert np.abs(spec - ref).max() <= 1e-13 * np.abs(ref).max()
    want = _direct_bins(ref, n, kj, kz0, nc, thresh, iw, nrbins)
    nz = want > 0
    assert np.array_equal(sp > 0, nz)
    assert np.allclose(sp[nz], want[nz], rtol=1e-11, atol=0)
