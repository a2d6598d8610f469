"""The drop-in boundary tested with the REFERENCE's own callers.

genpk_b200/libgenpk_dropin.so exports the link-time symbols the reference's
fieldize.o, powerspectrum.o and -lfftw3 provide (gen-pk.h:93-119, gen-pk.cpp:176-193).
oracle/_ref/libgenpk_refhost.so is the reference's host side compiled unmodified
(read_fieldize.cpp, GadgetReader, utils.cpp) with fieldize() left undefined; loaded
after the drop-in, its read_fieldize() calls OUR fieldize().  The loop below is the
per-type loop of gen-pk.cpp:204-239 on the bundled test_g2_snap (BASELINE config 1)."""
import ctypes as C
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DROPIN = os.path.join(ROOT, "genpk_b200", "libgenpk_dropin.so")
REFHOST = os.path.join(ROOT, "oracle", "_ref", "libgenpk_refhost.so")
SNAP = os.path.join(ROOT, "oracle", "_ref", "test_g2_snap")
GOLD = os.path.join(ROOT, "tests", "golden")
FIELDIZE = "_Z8fieldizediPdlPfS0_di"          # int fieldize(double,int,double*,long,float*,float*,double,int)


@pytest.fixture(scope="module")
def dropin():
    lib = C.CDLL(DROPIN, mode=C.RTLD_GLOBAL)
    lib.fftw_malloc.restype = C.c_void_p
    lib.fftw_malloc.argtypes = [C.c_size_t]
    lib.fftw_free.argtypes = [C.c_void_p]
    lib.fftw_plan_dft_r2c_3d.restype = C.c_void_p
    lib.fftw_plan_dft_r2c_3d.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_uint]
    lib.fftw_execute.argtypes = [C.c_void_p]
    lib.fftw_destroy_plan.argtypes = [C.c_void_p]
    lib.powerspectrum.restype = C.c_int
    lib.powerspectrum.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.c_double, C.c_double]
    lib.invwindow.restype = C.c_double
    lib.invwindow.argtypes = [C.c_int64] * 4
    f = getattr(lib, FIELDIZE)
    f.restype = C.c_int
    f.argtypes = [C.c_double, C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_double, C.c_int]
    return lib


def near(x, y, rel=1e-5):
    return abs(x - y) <= max(abs(x), abs(y)) * rel


def test_reference_known_answers_through_reference_symbols(dropin):
    """test.cpp:31-86 against the drop-in's symbols (host buffers the caller owns and reads)."""
    fieldize = getattr(dropin, FIELDIZE)
    dims = 5
    field = np.zeros(2 * dims * dims * (dims // 2 + 1))
    pos = (np.arange(30) / 3.0).astype(np.float32)
    masses = np.full(30, 10.0, np.float32)
    assert fieldize(10.0, dims, field.ctypes.data, 10, pos.ctypes.data, masses.ctypes.data, 10.0, 1) == 0
    assert near(field[0], 8.61111) and field[3] == 0 and field[20] == 0 and near(field[124], 1.66666)
    assert near(dropin.invwindow(0, 3, 4, 5), 71.8177719) and near(dropin.invwindow(4, 4, 4, 5), 6111.20801)
    assert dropin.invwindow(1, 1, 1, 0) == 0
    # check_powerspectrum: the caller fills the field on the host, plans in place, executes, bins
    d = 4
    n = 2 * d * d * (d // 2 + 1)
    buf = dropin.fftw_malloc(n * 8)
    f = np.ctypeslib.as_array((C.c_double * n).from_address(buf))
    f[:] = 0
    for i in range(32):
        f[6 * (i // 4) + i % 4] = 1
    f[0] = 2
    plan = dropin.fftw_plan_dft_r2c_3d(d, d, d, buf, buf, 0)
    assert plan
    dropin.fftw_execute(plan)
    power, count, keffs = np.zeros(10), np.zeros(10, np.int32), np.zeros(10)
    assert dropin.powerspectrum(d, buf, buf, 10, power.ctypes.data, count.ctypes.data, keffs.ctypes.data, 64.0, 64.0) == 0
    assert near(keffs[2], 2 ** 0.5) and count[2] == 12 and count[1] == 0 and count[0] == 6
    assert near(power[0], 0.0677526) and power[1] == 0 and near(power[2], 0.000565561) and near(power[9], 0.0550908)
    dropin.fftw_destroy_plan(plan)
    dropin.fftw_free(buf)


@pytest.mark.skipif(not os.path.exists(REFHOST), reason="oracle/_ref/libgenpk_refhost.so not built")
def test_gen_pk_loop_with_reference_reader(dropin, tmp_path):
    ref = C.CDLL(REFHOST, mode=C.RTLD_GLOBAL)          # its undefined fieldize() binds to the drop-in's
    ref.ref_snap_open.restype = C.c_void_p
    ref.ref_snap_open.argtypes = [C.c_char_p]
    ref.ref_snap_close.argtypes = [C.c_void_p]
    ref.ref_snap_npart.restype = C.c_int64
    ref.ref_snap_npart.argtypes = [C.c_void_p, C.c_int]
    ref.ref_snap_box.restype = C.c_double
    ref.ref_snap_box.argtypes = [C.c_void_p]
    ref.ref_read_fieldize.restype = C.c_int
    ref.ref_read_fieldize.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_int, C.POINTER(C.c_double)]
    ref.ref_print_pk.restype = C.c_int
    ref.ref_print_pk.argtypes = [C.c_char_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    gold = np.load(os.path.join(GOLD, "test_g2_snap.npz"))
    snap = ref.ref_snap_open(SNAP.encode())
    assert snap
    npart = [ref.ref_snap_npart(snap, t) for t in range(6)]
    assert npart == [4039, 4096, 0, 0, 57, 0]
    box = ref.ref_snap_box(snap)
    dims = nrbins = 32                                  # gen-pk.cpp:169-173 on these counts
    n = 2 * dims * dims * (dims // 2 + 1)
    field = dropin.fftw_malloc(n * 8)                   # gen-pk.cpp:176-181
    plan = dropin.fftw_plan_dft_r2c_3d(dims, dims, dims, field, field, 0)      # :193
    assert plan
    power, count, keffs = np.zeros(nrbins), np.zeros(nrbins, np.int32), np.zeros(nrbins)
    for t in range(6):                                  # :204
        if npart[t] == 0:
            continue
        C.memset(field, 0, n * 8)                       # :208
        tm = C.c_double(0.0)
        assert ref.ref_read_fieldize(field, snap, t, box, dims, C.byref(tm)) == 0    # :227 -> our fieldize()
        assert near(tm.value, float(gold[f"total_mass{t}"]), 1e-12)
        dropin.fftw_execute(plan)                       # :233
        rc = dropin.powerspectrum(dims, field, field, nrbins, power.ctypes.data, count.ctypes.data, keffs.ctypes.data,
                                  tm.value, tm.value)   # :234
        assert rc == 0
        assert np.array_equal(count, gold[f"count{t}"]), "mode counts differ from the reference"
        nz = count > 0
        np.testing.assert_allclose(power[nz], gold[f"power{t}"][nz], rtol=1e-5)
        np.testing.assert_allclose(keffs[nz], gold[f"keffs{t}"][nz], rtol=1e-5)
        out = tmp_path / f"PK-{t}"
        ref.ref_print_pk(str(out).encode(), nrbins, keffs.ctypes.data, power.ctypes.data, count.ctypes.data)   # :238
        rows = np.loadtxt(out)
        assert rows.shape == (int(nz.sum()), 3)
        assert np.array_equal(rows[:, 2].astype(np.int64), count[nz])
        # the host bytes behind the registered field are never touched by the drop-in
        assert not np.ctypeslib.as_array((C.c_double * n).from_address(field)).any()
    dropin.fftw_destroy_plan(plan)
    dropin.fftw_free(field)
    ref.ref_snap_close(snap)
