"""ctypes binding of libgenpk_cuda.so (include/genpk_cuda.h).

The library is the product: if it is missing or does not export the ABI this
module raises -- there is no CPU or PyTorch fallback anywhere in this package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
# GENPK_LIB: another build of the same library (kernel-variant measurements), never a fallback
LIB_PATH = os.environ.get("GENPK_LIB") or os.path.join(_HERE, "libgenpk_cuda.so")
CSRC = os.path.join(_HERE, "csrc")

c_f32p = C.c_void_p
c_f64p = C.c_void_p
c_i32p = C.c_void_p
c_i64p = C.c_void_p

# name -> (restype, argtypes); mirrors include/genpk_cuda.h one to one
SIGNATURES = {
    # 1. reference-signature shims
    "genpk_fieldize": (C.c_int, [C.c_double, C.c_int, c_f64p, C.c_int64, c_f32p, c_f32p, C.c_double, C.c_int]),
    "genpk_invwindow": (C.c_double, [C.c_int64, C.c_int64, C.c_int64, C.c_int64]),
    "genpk_r2c_3d": (C.c_int, [C.c_int, c_f64p]),
    "genpk_powerspectrum": (C.c_int, [C.c_int64, c_f64p, c_f64p, C.c_int, c_f64p, c_i32p, c_f64p, C.c_double,
                                      C.c_double]),
    # 2. handle API
    "genpk_create": (C.c_void_p, [C.c_int, C.c_int, C.c_uint]),
    "genpk_destroy": (None, [C.c_void_p]),
    "genpk_last_error": (C.c_char_p, []),
    "genpk_abi_version": (C.c_int, []),
    "genpk_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "genpk_set_option": (C.c_int, [C.c_void_p, C.c_int, C.c_int64]),
    "genpk_synchronize": (C.c_int, [C.c_void_p]),
    "genpk_grid_zero": (C.c_int, [C.c_void_p, C.c_int]),
    "genpk_deposit": (C.c_int, [C.c_void_p, C.c_int, c_f32p, c_f32p, C.c_int64, C.c_double, C.c_double, C.c_int]),
    "genpk_deposit_f64": (C.c_int, [C.c_void_p, C.c_int, c_f64p, c_f32p, C.c_int64, C.c_double, C.c_double, C.c_int]),
    "genpk_fft": (C.c_int, [C.c_void_p, C.c_int]),
    "genpk_power": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, c_f64p, c_i32p, c_f64p, C.c_double, C.c_double]),
    "genpk_power_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, c_f64p, c_i32p, c_f64p, C.c_double,
                                  C.c_double]),
    "genpk_fft_power": (C.c_int, [C.c_void_p, C.c_int, C.c_int, c_f64p, c_i32p, c_f64p, C.c_double, C.c_double]),
    "genpk_fft_power_cross": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, c_f64p, c_i32p, c_f64p, C.c_double,
                                        C.c_double]),
    "genpk_fused_xpass_supported": (C.c_int, [C.c_void_p, C.c_int]),
    "genpk_pk_from_particles": (C.c_int, [C.c_void_p, c_f32p, c_f32p, C.c_int64, C.c_double, C.c_double, C.c_double,
                                          C.c_int, c_f64p, c_i32p, c_f64p]),
    "genpk_grid_doubles": (C.c_size_t, [C.c_void_p]),
    "genpk_grid_download": (C.c_int, [C.c_void_p, C.c_int, c_f64p]),
    "genpk_grid_upload": (C.c_int, [C.c_void_p, C.c_int, c_f64p]),
    "genpk_grid_download_fixed": (C.c_int, [C.c_void_p, C.c_int, c_i64p]),
    "genpk_grid_device_ptr": (C.c_void_p, [C.c_void_p, C.c_int]),
    "genpk_stage_ms": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_float)]),
    "genpk_stage_total_ms": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_int64)]),
    "genpk_stage_reset": (C.c_int, [C.c_void_p]),
    "genpk_launch_count": (C.c_int64, [C.c_void_p]),
    "genpk_library_calls": (C.c_int64, [C.c_void_p]),
    "genpk_grid_scale_bits": (C.c_int, [C.c_void_p, C.c_int]),
    "genpk_last_order": (C.c_int, [C.c_void_p, c_i64p]),
    "genpk_last_sweep": (C.c_int, [C.c_void_p, c_i64p]),
    # 3. slab stages
    "genpk_create_slab": (C.c_void_p, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint]),
    "genpk_create_slab_wide": (C.c_void_p, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint, C.c_int]),
    "genpk_grid_owned_offset": (C.c_size_t, [C.c_void_p]),
    "genpk_take_rejected": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64)]),
    "genpk_ghost_side_ptr": (C.c_void_p, [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_size_t)]),
    "genpk_ghost_side_accumulate": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "genpk_route_particles": (C.c_int, [C.c_void_p, c_f32p, c_f32p, C.c_int64, C.c_double, c_f32p, c_f32p, c_i64p]),
    "genpk_ghost_ptr": (C.c_void_p, [C.c_void_p, C.c_int, C.POINTER(C.c_size_t)]),
    "genpk_ghost_accumulate": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "genpk_ipc_export_grid": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "genpk_slab_set_grid_peer": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "genpk_ghost_pull_ready": (C.c_int, [C.c_void_p, C.c_int]),
    "genpk_ghost_pull": (C.c_int, [C.c_void_p, C.c_int]),
    "genpk_rejected_to": (C.c_int, [C.c_void_p, C.c_void_p]),
    "genpk_slab_fft_yz": (C.c_int, [C.c_void_p, C.c_int]),
    "genpk_slab_pack": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "genpk_slab_fft_x": (C.c_int, [C.c_void_p, C.c_void_p]),
    "genpk_slab_spectrum_bytes": (C.c_size_t, [C.c_void_p]),
    "genpk_slab_power_partial": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, c_f64p]),
    "genpk_slab_fftx_power_partial": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, c_f64p]),
    "genpk_slab_recv_buffer": (C.c_void_p, [C.c_void_p, C.POINTER(C.c_size_t)]),
    "genpk_ipc_export": (C.c_int, [C.c_void_p, C.c_void_p]),
    "genpk_slab_set_peer": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "genpk_slab_scatter_supported": (C.c_int, [C.c_void_p]),
    "genpk_slab_fft_yz_scatter": (C.c_int, [C.c_void_p, C.c_int]),
    # 4. N GPUs behind one handle
    "genpk_multi_create": (C.c_void_p, [C.c_int, C.c_int, C.c_void_p, C.c_uint]),
    "genpk_multi_destroy": (None, [C.c_void_p]),
    "genpk_multi_ngpus": (C.c_int, [C.c_void_p]),
    "genpk_multi_rank_ctx": (C.c_void_p, [C.c_void_p, C.c_int]),
    "genpk_multi_set_option": (C.c_int, [C.c_void_p, C.c_int, C.c_int64]),
    "genpk_multi_grid_zero": (C.c_int, [C.c_void_p]),
    "genpk_multi_deposit": (C.c_int, [C.c_void_p, c_f32p, c_f32p, C.c_int64, C.c_double, C.c_double]),
    "genpk_multi_fft_power": (C.c_int, [C.c_void_p, C.c_int, c_f64p, c_i32p, c_f64p, C.c_double, C.c_double]),
    "genpk_multi_pk_from_particles": (C.c_int, [C.c_void_p, c_f32p, c_f32p, C.c_int64, C.c_double, C.c_double, C.c_double,
                                                C.c_int, c_f64p, c_i32p, c_f64p]),
    "genpk_power_finalize": (C.c_int, [c_f64p, C.c_int, C.c_double, C.c_double, c_f64p, c_i32p, c_f64p]),
    "genpk_rebin_min_modes": (C.c_int, [C.c_int, c_f64p, c_i32p, c_f64p, C.c_int64]),
    "genpk_bin_thresholds": (C.c_int, [C.c_int, C.c_int, C.c_uint, C.c_void_p]),
    # synthetic particle sets
    "genpk_synth_particles": (C.c_int, [C.c_int, C.c_uint64, C.c_int64, C.c_int64, C.c_int64, C.c_double, C.c_double,
                                        c_f32p, C.c_void_p]),
}

_lib = None


class GenPKError(RuntimeError):
    pass


def build(verbose: bool = False) -> str:
    """Compile libgenpk_cuda.so in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    out = None if verbose else subprocess.DEVNULL
    subprocess.run(["make", "-C", CSRC, "-j8"], check=True, stdout=out)
    return LIB_PATH


def load():
    """Load the CUDA library and bind every symbol of the header; raises when absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise GenPKError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                         "(there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name, None)
        if fn is None:
            raise GenPKError(f"{LIB_PATH} does not export {name}")
        fn.restype = res
        fn.argtypes = args
    if lib.genpk_abi_version() != 1:
        raise GenPKError("libgenpk_cuda.so ABI version mismatch")
    _lib = lib
    return lib


def last_error() -> str:
    return load().genpk_last_error().decode(errors="replace")


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise GenPKError(f"{what} failed (code {rc}): {last_error()}")
