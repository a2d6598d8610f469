"""Python host side above the C ABI: same names, argument meaning and error
behaviour as the reference's numerics interface (gen-pk.h:35-119).

numpy arrays are host buffers; the `*_dev` methods of Context take raw device
pointers (ints), e.g. ``torch_tensor.data_ptr()``.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import check

FLAG_FIXED_POINT = 0x1
FLAG_TWO_FIELDS = 0x2
FLAG_BINRULE_SOURCE = 0x4
OPT_DEPOSIT, OPT_SCALE_BITS, OPT_POWER = 1, 2, 3
OPT_LATTICE_N0, OPT_LATTICE_N1, OPT_MARCH_RY, OPT_MARCH_RX = 4, 5, 6, 7
OPT_FUSED_XPASS = 8
OPT_FFT_YZ_BATCH = 9
OPT_OWN_YPASS = 10
OPT_SWEEP, OPT_SWEEP_RY, OPT_ZERO_AHEAD, OPT_ZA_WINDOW, OPT_ZA_SLACK, OPT_ZA_DEFERRED = 11, 12, 13, 14, 15, 16
OPT_ZA_ZERO_CTAS, OPT_SWEEP_COUPLE, OPT_SWEEP_COUPLE_STEP, OPT_SWEEP_POLL_WEAK, OPT_SWEEP_RX = 17, 18, 19, 20, 21
OPT_F64_POSITIONS = 23
OPT_TMA = 24
OPT_ZERO_AFTER_POWER = 25
OPT_FUSED_ZY = 26
OPT_ZY_LAG = 27
POWER_CACHED, POWER_FUSED = 0, 1
DEPOSIT_AUTO, DEPOSIT_DIRECT, DEPOSIT_SORTED, DEPOSIT_TILED, DEPOSIT_MARCH, DEPOSIT_SWEEP = 0, 1, 2, 3, 4, 5
STAGE_DEPOSIT, STAGE_FFT, STAGE_POWER, STAGE_SORT, STAGE_ZERO = 0, 1, 2, 3, 4
SYNTH_UNIFORM_RANDOM, SYNTH_LATTICE, SYNTH_CLUSTERED = 0, 1, 2
FIELD_DIMS = 3072          # gen-pk.cpp:63


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


# ---- small host utilities of utils.cpp, kept for the per-type loop ------------------
def nexttwo(n: int) -> int:
    """Next power of two >= n (utils.cpp:35-42)."""
    n = int(n) - 1
    i = 1
    while i < 32:
        n |= n >> i
        i <<= 1
    return n + 1


def grid_dims_for(npart_total) -> int:
    """Grid rule of gen-pk.cpp:169-172: max over types of min(2*nexttwo((int)cbrt(N)), 3072)."""
    dims = 0
    for n in npart_total:
        c = np.cbrt(float(n))
        tmp = 2 * nexttwo(int(c)) if n > 0 else 2 * nexttwo(0)
        dims = max(dims, min(tmp, FIELD_DIMS))
    return dims


def type_str(ptype: int) -> str:
    """utils.cpp:57-72."""
    return {0: "by", 1: "DM", 2: "nu", 4: "st"}.get(ptype, "xx")


def rebin_min_modes(power, count, keffs, min_modes: int):
    """Merge neighbouring bins until each holds at least min_modes modes (the reference's to-do, gen-pk.cpp:27-31);
    mode-weighted means.  Returns new (power, count, keffs) arrays of the merged length."""
    lib = _lib.load()
    p = np.ascontiguousarray(power, np.float64).copy()
    c = np.ascontiguousarray(count, np.int32).copy()
    k = np.ascontiguousarray(keffs, np.float64).copy()
    n = lib.genpk_rebin_min_modes(len(p), p.ctypes.data, c.ctypes.data, k.ctypes.data, int(min_modes))
    if n < 0:
        raise _lib.GenPKError("genpk_rebin_min_modes: " + _lib.last_error())
    return p[:n], c[:n], k[:n]


def print_pk(filename: str, nrbins: int, keffs, power, count) -> int:
    """Three-column text output of utils.cpp:7-21 ("%e\\t%e\\t%d\\n" for non-empty bins)."""
    try:
        fd = open(filename, "w")
    except OSError:
        import sys
        sys.stderr.write("Error opening file: %s\n" % filename)
        return 0
    with fd:
        for i in range(nrbins):
            if count[i]:
                fd.write("%e\t%e\t%d\n" % (keffs[i], power[i], count[i]))
    return nrbins


# ---- reference-signature entry points (host buffers) -------------------------------
def fieldize(boxsize, dims, out, segment_particles, positions, masses, mass, extra) -> int:
    """fieldize() of fieldize.cpp:46 on the GPU: accumulates into the host grid `out`."""
    lib = _lib.load()
    positions = _f32(positions)
    n = int(segment_particles)
    assert positions.size >= 3 * n
    assert out.dtype == np.float64 and out.flags.c_contiguous
    assert out.size >= dims * dims * 2 * (dims // 2 + extra)
    mp = None
    if masses is not None:
        masses = _f32(masses)
        assert masses.size >= n
        mp = masses.ctypes.data
    rc = lib.genpk_fieldize(float(boxsize), int(dims), out.ctypes.data, n, positions.ctypes.data, mp, float(mass),
                            int(extra))
    check(rc, "genpk_fieldize")
    return rc


def invwindow(kx, ky, kz, n) -> float:
    """invwindow() of fieldize.cpp:125."""
    return _lib.load().genpk_invwindow(int(kx), int(ky), int(kz), int(n))


def r2c_3d(dims, field) -> None:
    """fftw_plan_dft_r2c_3d + fftw_execute of gen-pk.cpp:193,233, in place on a host buffer."""
    assert field.dtype == np.float64 and field.flags.c_contiguous
    assert field.size >= 2 * dims * dims * (dims // 2 + 1)
    check(_lib.load().genpk_r2c_3d(int(dims), field.ctypes.data), "genpk_r2c_3d")


def powerspectrum(dims, outfield, outfield2, nrbins, power, count, keffs, total_mass, total_mass2) -> int:
    """powerspectrum() of powerspectrum.c:35; outfield2 may be the same array."""
    lib = _lib.load()
    a = outfield.view(np.float64).reshape(-1)
    b = a if outfield2 is outfield or outfield2 is None else outfield2.view(np.float64).reshape(-1)
    need = 2 * dims * dims * (dims // 2 + 1)
    assert a.size >= need and b.size >= need and a.flags.c_contiguous and b.flags.c_contiguous
    assert power.dtype == np.float64 and keffs.dtype == np.float64 and count.dtype == np.int32
    rc = lib.genpk_powerspectrum(int(dims), a.ctypes.data, b.ctypes.data, int(nrbins), power.ctypes.data,
                                 count.ctypes.data, keffs.ctypes.data, float(total_mass), float(total_mass2))
    check(rc, "genpk_powerspectrum")
    return rc


# ---- handle API ----------------------------------------------------------------------
class Context:
    """One GPU, one grid (or x-slab of it) resident in HBM: zero -> deposit* -> fft -> power."""

    def __init__(self, dims: int, device: int = -1, flags: int = 0, nranks: int = 1, rank: int = 0,
                 ghost_planes: int = 0):
        self.lib = _lib.load()
        self.dims, self.nranks, self.rank, self.flags = int(dims), int(nranks), int(rank), int(flags)
        self.ghost_planes = int(ghost_planes) if nranks > 1 else 0
        if nranks == 1:
            self.h = self.lib.genpk_create(self.dims, int(device), self.flags)
        else:
            self.h = self.lib.genpk_create_slab_wide(self.dims, int(device), self.nranks, self.rank, self.flags,
                                                     self.ghost_planes)
        if not self.h:
            raise _lib.GenPKError("genpk_create failed: " + _lib.last_error())
        self.nc = self.dims // 2 + 1
        self.fd = 2 * self.nc
        self.nx = self.dims // self.nranks

    def close(self):
        if getattr(self, "h", None):
            self.lib.genpk_destroy(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # plumbing
    def set_stream(self, cuda_stream: int):
        check(self.lib.genpk_set_stream(self.h, cuda_stream), "genpk_set_stream")

    def set_option(self, option: int, value: int):
        check(self.lib.genpk_set_option(self.h, option, value), "genpk_set_option")

    def set_deposit_mode(self, mode: int):
        self.set_option(OPT_DEPOSIT, mode)

    def set_lattice_hint(self, n0: int, n1: int = 0):
        """Particle p sits near lattice site (ix,iy,iz), p = (ix*n1 + iy)*n0 + iz (0 = probe)."""
        self.set_option(OPT_LATTICE_N0, n0)
        self.set_option(OPT_LATTICE_N1, n1)

    def last_order(self) -> dict:
        """Verdict of the last order probe (diagnostics)."""
        out = np.zeros(7, np.int64)
        check(self.lib.genpk_last_order(self.h, out.ctypes.data), "genpk_last_order")
        keys = ("coherent", "lattice", "n0", "n1", "score_z", "score_y", "score_x")
        return {k: int(v) for k, v in zip(keys, out)}

    def last_sweep(self) -> dict:
        """The last lattice-sweep deposit: rows per column, columns (= warps), whether it cleared the grid
        ahead of its own front (zero ahead), and the window in planes (diagnostics)."""
        out = np.zeros(4, np.int64)
        check(self.lib.genpk_last_sweep(self.h, out.ctypes.data), "genpk_last_sweep")
        return {k: int(v) for k, v in zip(("ry", "columns", "zero_ahead", "window"), out)}

    def set_power_mode(self, mode: int):
        self.set_option(OPT_POWER, mode)

    def set_scale_bits(self, bits: int):
        """Fixed-point scale: q = llrint(w * 2^bits); -1 = from the particle masses of the first deposit after a zero."""
        self.set_option(OPT_SCALE_BITS, bits)

    def grid_scale_bits(self, which: int = 0) -> int:
        return int(self.lib.genpk_grid_scale_bits(self.h, which))

    def synchronize(self):
        check(self.lib.genpk_synchronize(self.h), "genpk_synchronize")

    def stage_ms(self, stage: int) -> float:
        ms = C.c_float(0)
        check(self.lib.genpk_stage_ms(self.h, stage, C.byref(ms)), "genpk_stage_ms")
        return ms.value

    def stage_total_ms(self, stage: int):
        """(sum of ms, number of recorded instances) since stage_reset()."""
        ms, n = C.c_float(0), C.c_int64(0)
        check(self.lib.genpk_stage_total_ms(self.h, stage, C.byref(ms), C.byref(n)), "genpk_stage_total_ms")
        return ms.value, n.value

    def stage_reset(self):
        check(self.lib.genpk_stage_reset(self.h), "genpk_stage_reset")

    def launch_count(self) -> int:
        return self.lib.genpk_launch_count(self.h)

    def library_calls(self) -> int:
        """cuFFT executions so far (the default paths at 256..2048 make none)."""
        return self.lib.genpk_library_calls(self.h)

    # the per-type step of gen-pk.cpp:208-234
    def grid_zero(self, which: int = 0):
        check(self.lib.genpk_grid_zero(self.h, which), "genpk_grid_zero")

    def deposit(self, positions, masses=None, mass=1.0, boxsize=1.0, which: int = 0, n=None):
        positions = _f32(positions)
        n = positions.size // 3 if n is None else int(n)
        mp = None
        if masses is not None:
            masses = _f32(masses)
            assert masses.size >= n
            mp = masses.ctypes.data
        check(self.lib.genpk_deposit(self.h, which, positions.ctypes.data, mp, n, float(mass), float(boxsize), 0),
              "genpk_deposit")

    def deposit_f64(self, positions, masses=None, mass=1.0, boxsize=1.0, which: int = 0):
        """Host positions in double precision, narrowed on the GPU (read_fieldize_bigfile.cpp:93-94)."""
        positions = np.ascontiguousarray(positions, dtype=np.float64)
        n = positions.size // 3
        mp = None
        if masses is not None:
            masses = _f32(masses)
            assert masses.size >= n
            mp = masses.ctypes.data
        check(self.lib.genpk_deposit_f64(self.h, which, positions.ctypes.data, mp, n, float(mass), float(boxsize), 0),
              "genpk_deposit_f64")

    def deposit_f64_dev(self, pos_ptr: int, n: int, mass_ptr: int = 0, mass=1.0, boxsize=1.0, which: int = 0):
        check(self.lib.genpk_deposit_f64(self.h, which, pos_ptr, mass_ptr or None, int(n), float(mass), float(boxsize), 1),
              "genpk_deposit_f64")

    def deposit_dev(self, pos_ptr: int, n: int, mass_ptr: int = 0, mass=1.0, boxsize=1.0, which: int = 0):
        check(self.lib.genpk_deposit(self.h, which, pos_ptr, mass_ptr or None, int(n), float(mass), float(boxsize), 1),
              "genpk_deposit")

    def deposit_host_ptr(self, pos_ptr: int, n: int, mass_ptr: int = 0, mass=1.0, boxsize=1.0, which: int = 0):
        check(self.lib.genpk_deposit(self.h, which, pos_ptr, mass_ptr or None, int(n), float(mass), float(boxsize), 0),
              "genpk_deposit")

    def fft(self, which: int = 0):
        check(self.lib.genpk_fft(self.h, which), "genpk_fft")

    def power(self, nrbins=None, total_mass=1.0, total_mass2=None, a: int = 0, b: int = 0):
        nrbins = self.dims if nrbins is None else int(nrbins)
        total_mass2 = total_mass if total_mass2 is None else total_mass2
        power = np.zeros(nrbins, np.float64)
        count = np.zeros(nrbins, np.int32)
        keffs = np.zeros(nrbins, np.float64)
        check(self.lib.genpk_power(self.h, a, b, nrbins, power.ctypes.data, count.ctypes.data, keffs.ctypes.data,
                                   float(total_mass), float(total_mass2)), "genpk_power")
        return power, count, keffs

    def fft_power(self, nrbins=None, total_mass=1.0, total_mass2=None, which: int = 0):
        """genpk_fft + genpk_power as one call; the x transform and the binning share one
        kernel when the grid side allows (the grid then holds the (y,z)-transformed planes)."""
        nrbins = self.dims if nrbins is None else int(nrbins)
        total_mass2 = total_mass if total_mass2 is None else total_mass2
        power = np.zeros(nrbins, np.float64)
        count = np.zeros(nrbins, np.int32)
        keffs = np.zeros(nrbins, np.float64)
        check(self.lib.genpk_fft_power(self.h, which, nrbins, power.ctypes.data, count.ctypes.data, keffs.ctypes.data,
                                       float(total_mass), float(total_mass2)), "genpk_fft_power")
        return power, count, keffs

    def fft_power_cross(self, nrbins=None, total_mass=1.0, total_mass2=1.0, a: int = 0, b: int = 1, other=None):
        """Cross spectrum of grid `a` of this context and grid `b` of `other` (default: this context, which then
        needs FLAG_TWO_FIELDS) straight from the deposited real grids: both transforms and the binning,
        on the fused path when the grid side allows (genpk_fft_power_cross)."""
        nrbins = self.dims if nrbins is None else int(nrbins)
        power = np.zeros(nrbins, np.float64)
        count = np.zeros(nrbins, np.int32)
        keffs = np.zeros(nrbins, np.float64)
        hb = (other or self).h
        check(self.lib.genpk_fft_power_cross(self.h, a, hb, b, nrbins, power.ctypes.data, count.ctypes.data,
                                             keffs.ctypes.data, float(total_mass), float(total_mass2)), "genpk_fft_power_cross")
        return power, count, keffs

    def fused_xpass_supported(self, nrbins=None) -> bool:
        nrbins = self.dims if nrbins is None else int(nrbins)
        return bool(self.lib.genpk_fused_xpass_supported(self.h, nrbins))

    def pk_from_particles(self, positions, masses=None, mass=1.0, boxsize=1.0, total_mass=None, nrbins=None):
        positions = _f32(positions)
        n = positions.size // 3
        nrbins = self.dims if nrbins is None else int(nrbins)
        mp = None
        if masses is not None:
            masses = _f32(masses)
            mp = masses.ctypes.data
        if total_mass is None:
            total_mass = float(np.sum(masses[:n], dtype=np.float64)) if masses is not None else mass * n
        power = np.zeros(nrbins, np.float64)
        count = np.zeros(nrbins, np.int32)
        keffs = np.zeros(nrbins, np.float64)
        check(self.lib.genpk_pk_from_particles(self.h, positions.ctypes.data, mp, n, float(mass), float(boxsize),
                                               float(total_mass), nrbins, power.ctypes.data, count.ctypes.data,
                                               keffs.ctypes.data), "genpk_pk_from_particles")
        return power, count, keffs

    def pk_from_particles_ptr(self, pos_host_ptr: int, n: int, mass=1.0, boxsize=1.0, total_mass=1.0, nrbins=None,
                              mass_host_ptr: int = 0):
        """genpk_pk_from_particles on raw HOST pointers (e.g. a pinned torch tensor)."""
        nrbins = self.dims if nrbins is None else int(nrbins)
        power = np.zeros(nrbins, np.float64)
        count = np.zeros(nrbins, np.int32)
        keffs = np.zeros(nrbins, np.float64)
        check(self.lib.genpk_pk_from_particles(self.h, pos_host_ptr, mass_host_ptr or None, int(n), float(mass),
                                               float(boxsize), float(total_mass), nrbins, power.ctypes.data,
                                               count.ctypes.data, keffs.ctypes.data), "genpk_pk_from_particles")
        return power, count, keffs

    # parity helpers
    def grid_doubles(self) -> int:
        return self.lib.genpk_grid_doubles(self.h)

    def grid_download(self, which: int = 0) -> np.ndarray:
        out = np.empty(self.grid_doubles(), np.float64)
        check(self.lib.genpk_grid_download(self.h, which, out.ctypes.data), "genpk_grid_download")
        return out

    def grid_download_fixed(self, which: int = 0) -> np.ndarray:
        out = np.empty(self.grid_doubles(), np.int64)
        check(self.lib.genpk_grid_download_fixed(self.h, which, out.ctypes.data), "genpk_grid_download_fixed")
        return out

    def grid_upload(self, host: np.ndarray, which: int = 0):
        host = np.ascontiguousarray(host.view(np.float64).reshape(-1))
        assert host.size == self.grid_doubles()
        check(self.lib.genpk_grid_upload(self.h, which, host.ctypes.data), "genpk_grid_upload")

    def grid_ptr(self, which: int = 0) -> int:
        p = self.lib.genpk_grid_device_ptr(self.h, which)
        if not p:
            raise _lib.GenPKError(_lib.last_error())
        return p

    # slab stages (device pointers)
    def route_particles(self, pos_ptr, mass_ptr, n, boxsize, spos_ptr, smass_ptr, counts_ptr):
        check(self.lib.genpk_route_particles(self.h, pos_ptr, mass_ptr or None, int(n), float(boxsize), spos_ptr,
                                             smass_ptr or None, counts_ptr), "genpk_route_particles")

    def take_rejected(self) -> int:
        """Particles the deposits rejected since the last call (waits for the stream)."""
        n = C.c_uint64(0)
        check(self.lib.genpk_take_rejected(self.h, C.byref(n)), "genpk_take_rejected")
        return int(n.value)

    def owned_offset(self) -> int:
        return int(self.lib.genpk_grid_owned_offset(self.h))

    def ghost_side_ptr(self, side: int, which: int = 0):
        nbytes = C.c_size_t(0)
        p = self.lib.genpk_ghost_side_ptr(self.h, which, side, C.byref(nbytes))
        if not p:
            raise _lib.GenPKError(_lib.last_error())
        return p, nbytes.value

    def ghost_side_accumulate(self, side: int, recv_ptr: int, which: int = 0):
        check(self.lib.genpk_ghost_side_accumulate(self.h, which, side, recv_ptr), "genpk_ghost_side_accumulate")

    def ghost_ptr(self, which: int = 0):
        nbytes = C.c_size_t(0)
        p = self.lib.genpk_ghost_ptr(self.h, which, C.byref(nbytes))
        if not p:
            raise _lib.GenPKError(_lib.last_error())
        return p, nbytes.value

    def ghost_accumulate(self, recv_ptr: int, which: int = 0):
        check(self.lib.genpk_ghost_accumulate(self.h, which, recv_ptr), "genpk_ghost_accumulate")

    def slab_fft_yz(self, which: int = 0):
        check(self.lib.genpk_slab_fft_yz(self.h, which), "genpk_slab_fft_yz")

    def slab_pack(self, send_ptr: int, which: int = 0):
        check(self.lib.genpk_slab_pack(self.h, which, send_ptr), "genpk_slab_pack")

    def slab_fft_x(self, recv_ptr: int):
        check(self.lib.genpk_slab_fft_x(self.h, recv_ptr), "genpk_slab_fft_x")

    def slab_spectrum_bytes(self) -> int:
        return self.lib.genpk_slab_spectrum_bytes(self.h)

    def slab_power_partial(self, spec_a_ptr: int, spec_b_ptr: int, nrbins: int, sums_ptr: int):
        check(self.lib.genpk_slab_power_partial(self.h, spec_a_ptr, spec_b_ptr or None, int(nrbins), sums_ptr),
              "genpk_slab_power_partial")


    def slab_fftx_power_partial(self, spec_yz_ptr: int, nrbins: int, sums_ptr: int):
        check(self.lib.genpk_slab_fftx_power_partial(self.h, spec_yz_ptr, int(nrbins), sums_ptr),
              "genpk_slab_fftx_power_partial")


    # transpose fused into the y pass (peer stores)
    IPC_HANDLE_BYTES = 64

    def slab_recv_buffer(self):
        """(device pointer, bytes) of this rank's library-owned transposed block [dims][ny][nc]."""
        nbytes = C.c_size_t(0)
        p = self.lib.genpk_slab_recv_buffer(self.h, C.byref(nbytes))
        if not p:
            raise _lib.GenPKError(_lib.last_error())
        return p, nbytes.value

    def ipc_export(self) -> bytes:
        buf = C.create_string_buffer(self.IPC_HANDLE_BYTES)
        check(self.lib.genpk_ipc_export(self.h, buf), "genpk_ipc_export")
        return buf.raw

    def slab_set_peer(self, rank: int, ipc_handle: bytes = None, same_process_ptr: int = 0):
        h = C.create_string_buffer(ipc_handle, self.IPC_HANDLE_BYTES) if ipc_handle is not None else None
        check(self.lib.genpk_slab_set_peer(self.h, int(rank), h, same_process_ptr or None), "genpk_slab_set_peer")

    def ipc_export_grid(self, which: int = 0) -> bytes:
        buf = C.create_string_buffer(self.IPC_HANDLE_BYTES)
        check(self.lib.genpk_ipc_export_grid(self.h, which, buf), "genpk_ipc_export_grid")
        return buf.raw

    def slab_set_grid_peer(self, side: int, ipc_handle: bytes = None, same_process_ptr: int = 0, which: int = 0):
        """side 0: the grid of rank-1, side 1: the grid of rank+1 (ghost exchange by peer loads)."""
        h = C.create_string_buffer(ipc_handle, self.IPC_HANDLE_BYTES) if ipc_handle is not None else None
        check(self.lib.genpk_slab_set_grid_peer(self.h, which, int(side), h, same_process_ptr or None),
              "genpk_slab_set_grid_peer")

    def ghost_pull_ready(self, which: int = 0) -> bool:
        return bool(self.lib.genpk_ghost_pull_ready(self.h, which))

    def ghost_pull(self, which: int = 0):
        check(self.lib.genpk_ghost_pull(self.h, which), "genpk_ghost_pull")

    def rejected_to(self, dst_ptr: int):
        """Stream-ordered: (double) rejected-particle count -> *dst_ptr, counter cleared."""
        check(self.lib.genpk_rejected_to(self.h, dst_ptr), "genpk_rejected_to")

    def slab_scatter_supported(self) -> bool:
        return bool(self.lib.genpk_slab_scatter_supported(self.h))

    def slab_fft_yz_scatter(self, which: int = 0):
        check(self.lib.genpk_slab_fft_yz_scatter(self.h, which), "genpk_slab_fft_yz_scatter")


class MultiContext:
    """N GPUs of one box behind one handle (genpk_multi_*): single process, slab decomposition in x, host
    particles in any order.  devices: ordinals of the GPUs (several slabs may share one)."""

    def __init__(self, dims: int, ngpus: int, devices=None, flags: int = 0):
        self.lib = _lib.load()
        self.dims, self.ngpus = int(dims), int(ngpus)
        dev = None
        if devices is not None:
            dev = (C.c_int * self.ngpus)(*[int(d) for d in devices])
        self.h = self.lib.genpk_multi_create(self.dims, self.ngpus, dev, int(flags))
        if not self.h:
            raise _lib.GenPKError("genpk_multi_create failed: " + _lib.last_error())

    def close(self):
        if getattr(self, "h", None):
            self.lib.genpk_multi_destroy(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_option(self, option: int, value: int):
        check(self.lib.genpk_multi_set_option(self.h, option, value), "genpk_multi_set_option")

    def grid_zero(self):
        check(self.lib.genpk_multi_grid_zero(self.h), "genpk_multi_grid_zero")

    def deposit(self, positions, masses=None, mass=1.0, boxsize=1.0):
        positions = _f32(positions)
        n = positions.size // 3
        mp = None
        if masses is not None:
            masses = _f32(masses)
            mp = masses.ctypes.data
        check(self.lib.genpk_multi_deposit(self.h, positions.ctypes.data, mp, n, float(mass), float(boxsize)),
              "genpk_multi_deposit")

    def deposit_host_ptr(self, pos_ptr: int, n: int, mass_ptr: int = 0, mass=1.0, boxsize=1.0):
        check(self.lib.genpk_multi_deposit(self.h, pos_ptr, mass_ptr or None, int(n), float(mass), float(boxsize)),
              "genpk_multi_deposit")

    def fft_power(self, nrbins=None, total_mass=1.0, total_mass2=None):
        nrbins = self.dims if nrbins is None else int(nrbins)
        total_mass2 = total_mass if total_mass2 is None else total_mass2
        power = np.zeros(nrbins, np.float64)
        count = np.zeros(nrbins, np.int32)
        keffs = np.zeros(nrbins, np.float64)
        check(self.lib.genpk_multi_fft_power(self.h, nrbins, power.ctypes.data, count.ctypes.data, keffs.ctypes.data,
                                             float(total_mass), float(total_mass2)), "genpk_multi_fft_power")
        return power, count, keffs

    def pk_from_particles_ptr(self, pos_host_ptr: int, n: int, mass=1.0, boxsize=1.0, total_mass=1.0, nrbins=None):
        nrbins = self.dims if nrbins is None else int(nrbins)
        power = np.zeros(nrbins, np.float64)
        count = np.zeros(nrbins, np.int32)
        keffs = np.zeros(nrbins, np.float64)
        check(self.lib.genpk_multi_pk_from_particles(self.h, pos_host_ptr, None, int(n), float(mass), float(boxsize),
                                                     float(total_mass), nrbins, power.ctypes.data, count.ctypes.data,
                                                     keffs.ctypes.data), "genpk_multi_pk_from_particles")
        return power, count, keffs

    def rank_context_handle(self, rank: int) -> int:
        return self.lib.genpk_multi_rank_ctx(self.h, int(rank))


def power_finalize(sums: np.ndarray, nrbins: int, total_mass: float, total_mass2: float):
    """powerspectrum.c:102-108 applied to all-reduced raw sums [3][nrbins]."""
    sums = np.ascontiguousarray(sums, np.float64)
    power = np.zeros(nrbins, np.float64)
    count = np.zeros(nrbins, np.int32)
    keffs = np.zeros(nrbins, np.float64)
    check(_lib.load().genpk_power_finalize(sums.ctypes.data, int(nrbins), float(total_mass), float(total_mass2),
                                           power.ctypes.data, count.ctypes.data, keffs.ctypes.data),
          "genpk_power_finalize")
    return power, count, keffs


def bin_thresholds(dims: int, nrbins: int, flags: int = 0) -> np.ndarray:
    """Plan-time bin edges in k^2 (host arithmetic; see genpk_bin_thresholds)."""
    out = np.zeros(nrbins + 1, np.uint32)
    check(_lib.load().genpk_bin_thresholds(int(dims), int(nrbins), int(flags), out.ctypes.data), "genpk_bin_thresholds")
    return out


def synth_particles_dev(kind: int, seed: int, n_side: int, first: int, count: int, boxsize: float, grid_dims: float,
                        pos_ptr: int, stream: int = 0):
    check(_lib.load().genpk_synth_particles(int(kind), int(seed), int(n_side), int(first), int(count), float(boxsize),
                                            float(grid_dims), pos_ptr, stream or None), "genpk_synth_particles")
