/* TEST INFRASTRUCTURE ONLY (see genpk_oracle.c for the rules).
 *
 * CPU restatement of the DETERMINISTIC fixed-point deposit mode.  This mode
 * has no counterpart in the reference (BASELINE.json north_star adds it); it is
 * defined here and in DESIGN.md, and the CUDA kernel must reproduce it bit for
 * bit.  Built WITHOUT -ffast-math and with -ffp-contract=off so that every
 * operation below is one IEEE-754 round-to-nearest operation in the written
 * order.
 *
 * Rule (per particle, geometry exactly as fieldize.cpp:63-92):
 *     x  = (double)pos_f32 * (dims / boxsize)          one multiply
 *     f  = floor(x);  d = x - f;  t = 1.0 - d
 *     w  = ((m * ax) * ay) * az                         ax in {t,d}, left to right
 *     q  = llrint(w * 2^scale_bits)                     round half to even
 *     grid_q[cell] += q                                 int64, order independent
 * and the real grid is (double)grid_q * 2^-scale_bits.
 */
#include <math.h>
#include <stdint.h>
#include <stddef.h>

static inline int wrap_cell(int c, int dims)
{
    c %= dims;
    return c < 0 ? c + dims : c;
}

int oracle_fieldize_fixed(double boxsize, int dims, int64_t *out_q, int64_t n,
                          const float *positions, const float *masses, double mass,
                          int extra, int scale_bits)
{
    const size_t fd = 2 * (size_t)(dims / 2 + extra);
    const size_t plane = fd * (size_t)dims;
    const double units = dims / boxsize;
    const double scale = ldexp(1.0, scale_bits);
    for (int64_t p = 0; p < n; p++) {
        const double m = masses ? (double)masses[p] : mass;
        int lo[3], hi[3];
        double wl[3], wh[3];
        for (int a = 0; a < 3; a++) {
            const double x = (double)positions[3 * p + a] * units;
            const double fl = floor(x);
            const int f = (int)fl;
            wh[a] = x - fl;
            wl[a] = 1.0 - wh[a];
            hi[a] = wrap_cell(f + 1, dims);
            lo[a] = wrap_cell(f, dims);
        }
        for (int c = 0; c < 8; c++) {
            const int sx = c & 1, sy = (c >> 1) & 1, sz = (c >> 2) & 1;
            double w = m * (sx ? wh[0] : wl[0]);
            w = w * (sy ? wh[1] : wl[1]);
            w = w * (sz ? wh[2] : wl[2]);
            const int64_t q = llrint(w * scale);
            const size_t idx = plane * (size_t)(sx ? hi[0] : lo[0]) + fd * (size_t)(sy ? hi[1] : lo[1])
                             + (size_t)(sz ? hi[2] : lo[2]);
            out_q[idx] += q;
        }
    }
    return 0;
}

/* (double)q * 2^-scale_bits for every cell. */
void oracle_fixed_to_double(const int64_t *q, double *out, size_t ncell, int scale_bits)
{
    const double inv = ldexp(1.0, -scale_bits);
    for (size_t i = 0; i < ncell; i++)
        out[i] = (double)q[i] * inv;
}

/* The same rule for double-precision positions used as they are -- what a reference built with
 * -DDOUBLE_PRECISION_SNAP hands to fieldize() (gen-pk.h:25-29: GENPK_FLOAT_TYPE = double;
 * fieldize.cpp:66 multiplies the double position itself).  Masses stay float32 arrays cast to
 * the particle type in that build; here they are taken as float32, which is how they are stored. */
int oracle_fieldize_fixed_f64(double boxsize, int dims, int64_t *out_q, int64_t n,
                              const double *positions, const float *masses, double mass,
                              int extra, int scale_bits)
{
    const size_t fd = 2 * (size_t)(dims / 2 + extra);
    const size_t plane = fd * (size_t)dims;
    const double units = dims / boxsize;
    const double scale = ldexp(1.0, scale_bits);
    for (int64_t p = 0; p < n; p++) {
        const double m = masses ? (double)masses[p] : mass;
        int lo[3], hi[3];
        double wl[3], wh[3];
        for (int a = 0; a < 3; a++) {
            const double x = positions[3 * p + a] * units;
            const double fl = floor(x);
            const int f = (int)fl;
            wh[a] = x - fl;
            wl[a] = 1.0 - wh[a];
            hi[a] = wrap_cell(f + 1, dims);
            lo[a] = wrap_cell(f, dims);
        }
        for (int c = 0; c < 8; c++) {
            const int sx = c & 1, sy = (c >> 1) & 1, sz = (c >> 2) & 1;
            double w = m * (sx ? wh[0] : wl[0]);
            w = w * (sy ? wh[1] : wl[1]);
            w = w * (sz ? wh[2] : wl[2]);
            const int64_t q = llrint(w * scale);
            const size_t idx = plane * (size_t)(sx ? hi[0] : lo[0]) + fd * (size_t)(sy ? hi[1] : lo[1])
                             + (size_t)(sz ? hi[2] : lo[2]);
            out_q[idx] += q;
        }
    }
    return 0;
}
