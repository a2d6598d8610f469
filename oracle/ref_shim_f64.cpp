/* TEST INFRASTRUCTURE ONLY.
 *
 * Doorway into the reference's fieldize() compiled UNMODIFIED with -DDOUBLE_PRECISION_SNAP
 * (gen-pk.h:25-29: particle arrays are double; read_fieldize.cpp:24-25 hands them to fieldize()
 * as they are).  Built by oracle/Makefile into oracle/_ref/libgenpk_ref_f64.so.  No arithmetic here.
 */
#include "gen-pk.h"

extern "C" int ref_fieldize_f64(double boxsize, int dims, double *out, int64_t n, double *positions, double *masses,
                                double mass, int extra)
{
    return fieldize(boxsize, dims, out, n, positions, masses, mass, extra);
}
