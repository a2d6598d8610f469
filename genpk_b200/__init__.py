"""genpk_b200 -- B200 (sm_100a) implementation of GenPK's P(k) hot path.

Host-side mirror of the reference's function boundary (gen-pk.h:93-119) over the
C ABI of libgenpk_cuda.so.  Device work only happens inside that library.
"""
from ._lib import GenPKError, build, load, LIB_PATH
from .api import (Context, fieldize, invwindow, powerspectrum, r2c_3d, nexttwo, grid_dims_for, type_str, print_pk,
                  FLAG_FIXED_POINT, FLAG_TWO_FIELDS, FLAG_BINRULE_SOURCE, DEPOSIT_AUTO, DEPOSIT_DIRECT, DEPOSIT_SORTED,
                  DEPOSIT_TILED, DEPOSIT_MARCH, SYNTH_UNIFORM_RANDOM, SYNTH_LATTICE, SYNTH_CLUSTERED)

__all__ = ["GenPKError", "build", "load", "LIB_PATH", "Context", "fieldize", "invwindow", "powerspectrum", "r2c_3d",
           "nexttwo", "grid_dims_for", "type_str", "print_pk"]
