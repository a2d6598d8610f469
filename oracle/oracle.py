"""TEST INFRASTRUCTURE ONLY -- ctypes front end to the CPU oracle.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module.  Nothing under genpk_b200/ does.

Two back ends with the same call surface:

* ``Oracle("port")``       oracle/libgenpk_oracle.so, our C restatement
                           (genpk_oracle.c, genpk_oracle_strict.c);
* ``Oracle("reference")``  oracle/_ref/libgenpk_ref.so, the reference's own
                           object code (fieldize.cpp, powerspectrum.c, ...
                           compiled unmodified by oracle/Makefile).

The 3-D r2c transform is *not* in either: the reference delegates it to FFTW3
(gen-pk.cpp:193,233; ``-lfftw3`` Makefile:17, version unpinned, not vendored and
not installed in this image).  Its published contract -- unnormalised forward
DFT, exp(-2 pi i ...), real -> half complex, x slowest -- is restated here with
pocketfft (``scipy.fft.rfftn`` / ``numpy.fft.rfftn``).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(_HERE, "libgenpk_oracle.so")
REF_SO = os.path.join(_HERE, "_ref", "libgenpk_ref.so")
REF_F64_SO = os.path.join(_HERE, "_ref", "libgenpk_ref_f64.so")     # fieldize.cpp built with -DDOUBLE_PRECISION_SNAP
REF_SNAPSHOT = os.path.join(_HERE, "_ref", "test_g2_snap")

_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")


def build(ref: bool = True) -> None:
    """Compile the oracle (and oracle/_ref when /root/reference exists)."""
    targets = ["oracle"] + (["ref"] if ref else [])
    subprocess.run(["make", "-s", "-C", _HERE] + targets, check=True)


def have_reference() -> bool:
    return os.path.exists(REF_SO)


def padded_shape(dims: int, extra: int = 1):
    return (dims, dims, 2 * (dims // 2 + extra))


def rfftn_padded(field: np.ndarray, dims: int, workers: int = -1) -> np.ndarray:
    """FFTW-style in-place r2c restated out of place: padded real [d][d][d+2] ->
    complex128 [d][d][d/2+1] (gen-pk.cpp:193,233)."""
    real = np.ascontiguousarray(field.reshape(padded_shape(dims))[:, :, :dims])
    try:
        import scipy.fft as sfft
        return np.ascontiguousarray(sfft.rfftn(real, workers=workers))
    except ImportError:  # pragma: no cover
        return np.ascontiguousarray(np.fft.rfftn(real))


class Oracle:
    def __init__(self, kind: str = "port"):
        self.kind = kind
        if kind == "port":
            if not os.path.exists(PORT_SO):
                build(ref=False)
            self.lib = C.CDLL(PORT_SO)
            self._fieldize = self.lib.oracle_fieldize
            self._invwindow = self.lib.oracle_invwindow
            self._powerspectrum = self.lib.oracle_powerspectrum
        elif kind == "reference":
            if not os.path.exists(REF_SO):
                raise FileNotFoundError(REF_SO + " (run `make -C oracle ref` where /root/reference exists)")
            self.lib = C.CDLL(REF_SO)
            self._fieldize = self.lib.ref_fieldize
            self._invwindow = self.lib.invwindow
            self._powerspectrum = self.lib.powerspectrum
        else:
            raise ValueError(kind)
        self._fieldize.restype = C.c_int
        self._fieldize.argtypes = [C.c_double, C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                   C.c_double, C.c_int]
        self._invwindow.restype = C.c_double
        self._invwindow.argtypes = [C.c_int64] * 4
        self._powerspectrum.restype = C.c_int
        self._powerspectrum.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_int, _f64p, _i32p, _f64p,
                                        C.c_double, C.c_double]

    # -- print_pk(), utils.cpp:7-21 (the reference's own object code; reference kind only) ------
    def print_pk(self, filename, nrbins, keffs, power, count):
        if self.kind != "reference":
            raise RuntimeError("print_pk: only the reference's objects carry it")
        f = self.lib.ref_print_pk
        f.restype = C.c_int
        f.argtypes = [C.c_char_p, C.c_int, _f64p, _f64p, _i32p]
        return f(str(filename).encode(), int(nrbins), np.ascontiguousarray(keffs, np.float64),
                 np.ascontiguousarray(power, np.float64), np.ascontiguousarray(count, np.int32))

    # -- fieldize(), fieldize.cpp:46 ------------------------------------------------
    def fieldize(self, boxsize, dims, out, positions, masses=None, mass=1.0, extra=1):
        positions = np.ascontiguousarray(positions, dtype=np.float32)
        n = positions.size // 3
        assert out.dtype == np.float64 and out.flags.c_contiguous
        assert out.size >= dims * dims * 2 * (dims // 2 + extra)
        mp = None
        if masses is not None:
            masses = np.ascontiguousarray(masses, dtype=np.float32)
            assert masses.size >= n
            mp = masses.ctypes.data
        return self._fieldize(float(boxsize), int(dims), out.ctypes.data, n, positions.ctypes.data, mp,
                              float(mass), int(extra))

    # -- invwindow(), fieldize.cpp:125 ----------------------------------------------
    def invwindow(self, kx, ky, kz, n):
        return self._invwindow(int(kx), int(ky), int(kz), int(n))

    # -- powerspectrum(), powerspectrum.c:35 ----------------------------------------
    def powerspectrum(self, dims, f1, f2, nrbins, total_mass, total_mass2):
        f1 = np.ascontiguousarray(f1, dtype=np.complex128)
        f2 = f1 if f2 is None else np.ascontiguousarray(f2, dtype=np.complex128)
        assert f1.size == dims * dims * (dims // 2 + 1) == f2.size
        power = np.empty(nrbins, np.float64)
        count = np.empty(nrbins, np.int32)
        keffs = np.empty(nrbins, np.float64)
        rc = self._powerspectrum(int(dims), f1.ctypes.data, f2.ctypes.data, int(nrbins), power, count, keffs,
                                 float(total_mass), float(total_mass2))
        return rc, power, count, keffs

    # -- whole path: deposit -> FFT -> binning (gen-pk.cpp:208-234) -------------------
    def pk(self, boxsize, dims, positions, masses=None, mass=1.0, total_mass=None, nrbins=None):
        field = np.zeros(padded_shape(dims), np.float64)
        self.fieldize(boxsize, dims, field, positions, masses, mass, 1)
        if total_mass is None:
            n = np.asarray(positions).size // 3
            total_mass = float(np.sum(np.asarray(masses[:n], np.float64))) if masses is not None else mass * n
        spec = rfftn_padded(field, dims)
        rc, power, count, keffs = self.powerspectrum(dims, spec, None, nrbins or dims, total_mass, total_mass)
        return field, power, count, keffs

    # -- port-only helpers ------------------------------------------------------------
    def fieldize_fixed(self, boxsize, dims, out_q, positions, masses=None, mass=1.0, extra=1, scale_bits=40):
        assert self.kind == "port"
        positions = np.ascontiguousarray(positions, dtype=np.float32)
        n = positions.size // 3
        assert out_q.dtype == np.int64 and out_q.flags.c_contiguous
        mp = None
        if masses is not None:
            masses = np.ascontiguousarray(masses, dtype=np.float32)
            mp = masses.ctypes.data
        f = self.lib.oracle_fieldize_fixed
        f.restype = C.c_int
        f.argtypes = [C.c_double, C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_double, C.c_int,
                      C.c_int]
        return f(float(boxsize), int(dims), out_q.ctypes.data, n, positions.ctypes.data, mp, float(mass),
                 int(extra), int(scale_bits))

    def fieldize_fixed_f64(self, boxsize, dims, out_q, positions, masses=None, mass=1.0, extra=1, scale_bits=40):
        """Fixed-point rule on double-precision positions used as they are (DOUBLE_PRECISION_SNAP)."""
        assert self.kind == "port"
        positions = np.ascontiguousarray(positions, dtype=np.float64)
        n = positions.size // 3
        assert out_q.dtype == np.int64 and out_q.flags.c_contiguous
        mp = None
        if masses is not None:
            masses = np.ascontiguousarray(masses, dtype=np.float32)
            mp = masses.ctypes.data
        f = self.lib.oracle_fieldize_fixed_f64
        f.restype = C.c_int
        f.argtypes = [C.c_double, C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_double, C.c_int,
                      C.c_int]
        return f(float(boxsize), int(dims), out_q.ctypes.data, n, positions.ctypes.data, mp, float(mass),
                 int(extra), int(scale_bits))

    def fixed_to_double(self, q, scale_bits):
        assert self.kind == "port"
        out = np.empty(q.shape, np.float64)
        f = self.lib.oracle_fixed_to_double
        f.restype = None
        f.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
        f(q.ctypes.data, out.ctypes.data, q.size, int(scale_bits))
        return out

    def mode_counts(self, dims, nrbins):
        """Per-bin mode count and sum of |k| without any field memory."""
        assert self.kind == "port"
        count = np.zeros(nrbins, np.int64)
        ksum = np.zeros(nrbins, np.float64)
        f = self.lib.oracle_mode_counts
        f.restype = C.c_int
        f.argtypes = [C.c_int64, C.c_int, _i64p, _f64p]
        f(int(dims), int(nrbins), count, ksum)
        return count, ksum

    def bin_of_k2(self, dims, nrbins, k2):
        assert self.kind == "port"
        f = self.lib.oracle_bin_of_k2
        f.restype = C.c_int
        f.argtypes = [C.c_int64, C.c_int, C.c_int64]
        return f(int(dims), int(nrbins), int(k2))

    def naive_r2c_3d(self, n, field):
        assert self.kind == "port"
        out = np.zeros((n, n, n // 2 + 1), np.complex128)
        f = self.lib.oracle_naive_r2c_3d
        f.restype = C.c_int
        f.argtypes = [C.c_int, C.c_void_p, C.c_void_p]
        fld = np.ascontiguousarray(field, np.float64)
        f(int(n), fld.ctypes.data, out.ctypes.data)
        return out


def have_reference_f64() -> bool:
    return os.path.exists(REF_F64_SO)


def ref_fieldize_f64(boxsize, dims, out, positions, masses=None, mass=1.0, extra=1):
    """fieldize() of the reference compiled with -DDOUBLE_PRECISION_SNAP (gen-pk.h:25-29): positions and
    masses are double arrays, used as they are (fieldize.cpp:46, 63-69)."""
    lib = C.CDLL(REF_F64_SO)
    f = lib.ref_fieldize_f64
    f.restype = C.c_int
    f.argtypes = [C.c_double, C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_double, C.c_int]
    positions = np.ascontiguousarray(positions, dtype=np.float64)
    n = positions.size // 3
    mp = None
    if masses is not None:
        masses = np.ascontiguousarray(masses, dtype=np.float64)
        mp = masses.ctypes.data
    return f(float(boxsize), int(dims), out.ctypes.data, n, positions.ctypes.data, mp, float(mass), int(extra))


class RefSnapshot:
    """The reference's GadgetReader + read_fieldize on a Gadget snapshot
    (read_fieldize.cpp:18-97), reference back end only."""

    def __init__(self, base: str = REF_SNAPSHOT):
        self.lib = C.CDLL(REF_SO)
        L = self.lib
        L.ref_snap_open.restype = C.c_void_p
        L.ref_snap_open.argtypes = [C.c_char_p]
        L.ref_snap_close.argtypes = [C.c_void_p]
        L.ref_snap_numfiles.argtypes = [C.c_void_p]
        L.ref_snap_npart.restype = C.c_int64
        L.ref_snap_npart.argtypes = [C.c_void_p, C.c_int]
        for name in ("ref_snap_mass",):
            getattr(L, name).restype = C.c_double
            getattr(L, name).argtypes = [C.c_void_p, C.c_int]
        for name in ("ref_snap_box", "ref_snap_redshift", "ref_snap_omega0"):
            getattr(L, name).restype = C.c_double
            getattr(L, name).argtypes = [C.c_void_p]
        L.ref_snap_getblock.restype = C.c_int64
        L.ref_snap_getblock.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int]
        L.ref_read_fieldize.restype = C.c_int
        L.ref_read_fieldize.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_int, C.c_void_p]
        self.h = L.ref_snap_open(base.encode())

    def close(self):
        if self.h:
            self.lib.ref_snap_close(self.h)
            self.h = None

    def numfiles(self):
        return self.lib.ref_snap_numfiles(self.h)

    def npart(self, t):
        return self.lib.ref_snap_npart(self.h, t)

    def mass(self, t):
        return self.lib.ref_snap_mass(self.h, t)

    def box(self):
        return self.lib.ref_snap_box(self.h)

    def redshift(self):
        return self.lib.ref_snap_redshift(self.h)

    def omega0(self):
        return self.lib.ref_snap_omega0(self.h)

    def get_block(self, name: str, n: int, start: int, skip_type: int, width: int):
        buf = np.zeros(n * width, np.float32)
        got = self.lib.ref_snap_getblock(self.h, name.encode(), buf.ctypes.data, n, start, skip_type)
        return got, buf

    def read_fieldize(self, field, ptype, box, dims, total_mass):
        tm = C.c_double(total_mass)
        rc = self.lib.ref_read_fieldize(field.ctypes.data, self.h, ptype, float(box), int(dims), C.byref(tm))
        return rc, tm.value
