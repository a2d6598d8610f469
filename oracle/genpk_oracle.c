/* TEST INFRASTRUCTURE ONLY -- never linked into, imported by, or executed from
 * the product path (genpk_b200/, include/).  Only tests/, the smoke check in
 * __graft_entry__.py and the cpu_baseline / --impl reference legs of bench.py
 * may load this.
 *
 * CPU restatement of the three functions on GenPK's P(k) hot path, written
 * from the behaviour of the reference (citations are file:line into
 * /root/reference).  Built with the reference's own arithmetic flags
 * (-O2 -ffast-math -fopenmp, reference Makefile:30,36-37).
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this file against
 *   (1) every numeric known answer of the reference's test.cpp:31-113, and
 *   (2) the reference's own object code (oracle/_ref/libgenpk_ref.so, built by
 *       oracle/Makefile from the unmodified sources) on seeded inputs,
 * and tests/golden/ holds vectors produced by (2).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.1415926535897932384626433832795
#endif

/* Periodic wrap of a cell coordinate that may be negative or >= dims
 * (fieldize.cpp:70-75: C remainder, then +dims if negative). */
static inline int wrap_cell(int c, int dims)
{
    c %= dims;
    return c < 0 ? c + dims : c;
}

/* Cloud-in-cell mass assignment, restating fieldize() (fieldize.cpp:46-114).
 *   out       padded real grid, index = (dims*fd)*X + fd*Y + Z with
 *             fd = 2*(dims/2+extra) (fieldize.cpp:48-51); accumulated into.
 *   positions AoS float32 [n][3]; masses optional per-particle float32.
 * Particles are visited in index order (the reference with one OpenMP thread
 * does the same; with more threads its order is schedule dependent). */
int oracle_fieldize(double boxsize, int dims, double *out, int64_t n,
                    const float *positions, const float *masses, double mass, int extra)
{
    const size_t fd = 2 * (size_t)(dims / 2 + extra);
    const size_t plane = fd * (size_t)dims;
    const double units = dims / boxsize;                     /* fieldize.cpp:52 */
    for (int64_t p = 0; p < n; p++) {
        const double m = masses ? (double)masses[p] : mass;  /* fieldize.cpp:63 */
        int lo[3], hi[3];
        double wl[3], wh[3];
        for (int a = 0; a < 3; a++) {
            const double x = positions[3 * p + a] * units;   /* :66 */
            const int f = (int)floor(x);                     /* :67 */
            wh[a] = x - f;                                   /* :68  dx */
            wl[a] = 1.0 - wh[a];                             /* :69  tx */
            hi[a] = wrap_cell(f + 1, dims);                  /* :70-72 */
            lo[a] = wrap_cell(f, dims);                      /* :73-75 */
        }
        /* Eight corners; bit 0 of c selects x, bit 1 y, bit 2 z, matching the
         * order of fieldize.cpp:77-92. */
        for (int c = 0; c < 8; c++) {
            const int sx = c & 1, sy = (c >> 1) & 1, sz = (c >> 2) & 1;
            const double w = m * (sx ? wh[0] : wl[0]) * (sy ? wh[1] : wl[1]) * (sz ? wh[2] : wl[2]);
            const size_t idx = plane * (size_t)(sx ? hi[0] : lo[0]) + fd * (size_t)(sy ? hi[1] : lo[1])
                             + (size_t)(sz ? hi[2] : lo[2]);
            out[idx] += w;
        }
    }
    return 0;
}

/* One axis of the inverse CIC window: pi k / (n sin(pi k / n)), 1 at k=0
 * (fieldize.cpp:117-121; the (float)n cast is exact for n < 2^24). */
static inline double oned_invwindow(int64_t k, int64_t n)
{
    return k ? M_PI * k / (n * sin(M_PI * k / (float)n)) : 1.0;
}

/* invwindow() (fieldize.cpp:125-133): each axis factor is narrowed to float32,
 * the product is taken in float32, and the square is taken in double
 * (C++ pow(float,int) promotes). */
double oracle_invwindow(int64_t kx, int64_t ky, int64_t kz, int64_t n)
{
    if (n == 0)
        return 0;
    const float ax = (float)oned_invwindow(kx, n);
    const float ay = (float)oned_invwindow(ky, n);
    const float az = (float)oned_invwindow(kz, n);
    const float prod = ax * ay * az;
    return (double)prod * (double)prod;
}

/* Signed wavenumber of FFT index i (KVAL, powerspectrum.c:33: i==dims/2 stays +dims/2). */
static inline int64_t kval(int64_t i, int64_t dims) { return i <= dims / 2 ? i : i - dims; }

struct bin_acc { double *p, *k; int64_t *c; };

static inline void add_mode(struct bin_acc *acc, int nrbins, double binsperunit, int64_t dims,
                            int64_t ki, int64_t kj, int64_t kz, int mult,
                            const double *a, const double *b)
{
    /* kk = sqrt(ki^2+kj^2+kz^2); bin = floor(binsperunit*log(kk))
     * (powerspectrum.c:64-66, 74-75, 83-84). */
    const double kk = sqrt((double)(ki * ki) + (double)(kj * kj) + (double)(kz * kz));
    const int64_t bin = (int64_t)floor(binsperunit * log(kk));
    if (bin < 0 || bin >= nrbins)
        abort();                                              /* assert, :67,76,85 */
    if (a) {
        const double w = oracle_invwindow(ki, kj, kz, dims);
        acc->p[bin] += mult * (a[0] * b[0] + a[1] * b[1]) * (w * w);   /* :68,77,86 */
    }
    acc->k[bin] += mult * kk;
    acc->c[bin] += mult;
}

static int spectrum_pass(int64_t dims, const double *f1, const double *f2, int nrbins,
                         double *psum, double *ksum, int64_t *csum)
{
    const double binsperunit = (nrbins - 1) / log(sqrt(3) * dims / 2.0);   /* powerspectrum.c:38 */
    const int64_t nc = dims / 2 + 1;
    memset(psum, 0, nrbins * sizeof(double));
    memset(ksum, 0, nrbins * sizeof(double));
    memset(csum, 0, nrbins * sizeof(int64_t));
    #pragma omp parallel
    {
        struct bin_acc acc;
        acc.p = calloc(nrbins, sizeof(double));
        acc.k = calloc(nrbins, sizeof(double));
        acc.c = calloc(nrbins, sizeof(int64_t));
        #pragma omp for nowait
        for (int64_t i = 0; i < dims; i++) {
            for (int64_t j = 0; j < dims; j++) {
                const int64_t row = (i * dims + j) * nc;
                const int64_t ki = kval(i, dims), kj = kval(j, dims);
                for (int64_t k = 0; k < nc; k++) {
                    /* kz=0 and kz=dims/2 planes are their own Hermitian mirror:
                     * weight 1, all others 2; the DC mode is dropped
                     * (powerspectrum.c:59-89). */
                    if (k == 0 && ki == 0 && kj == 0)
                        continue;
                    const int mult = (k == 0 || k == dims / 2) ? 1 : 2;
                    const double *a = f1 ? f1 + 2 * (row + k) : NULL;
                    const double *b = f2 ? f2 + 2 * (row + k) : NULL;
                    add_mode(&acc, nrbins, binsperunit, dims, ki, kj, kval(k, dims), mult, a, b);
                }
            }
        }
        #pragma omp critical
        for (int b = 0; b < nrbins; b++) {                    /* :93-100 */
            psum[b] += acc.p[b];
            ksum[b] += acc.k[b];
            csum[b] += acc.c[b];
        }
        free(acc.p); free(acc.k); free(acc.c);
    }
    return 0;
}

/* powerspectrum() (powerspectrum.c:35-110).  f1,f2: interleaved complex double
 * [dims][dims][dims/2+1]; may alias.  Outputs zeroed here, normalised by
 * total_mass*total_mass2 and by the mode count (two divisions, :104-105). */
int oracle_powerspectrum(int64_t dims, const double *f1, const double *f2, int nrbins,
                         double *power, int *count, double *keffs, double total_mass, double total_mass2)
{
    int64_t *c64 = malloc(nrbins * sizeof(int64_t));
    spectrum_pass(dims, f1, f2, nrbins, power, keffs, c64);
    for (int b = 0; b < nrbins; b++) {
        count[b] = (int)c64[b];
        if (count[b]) {
            power[b] /= total_mass * total_mass2;
            power[b] /= count[b];
            keffs[b] /= count[b];
        }
    }
    free(c64);
    return 0;
}

/* Geometry only: per-bin mode count and sum of |k| with no field memory, for
 * grids too large to hold on the host (same expressions as above). */
int oracle_mode_counts(int64_t dims, int nrbins, int64_t *count, double *ksum)
{
    double *dummy = malloc(nrbins * sizeof(double));
    spectrum_pass(dims, NULL, NULL, nrbins, dummy, ksum, count);
    free(dummy);
    return 0;
}

/* The bin of a squared integer wavenumber, as the functions above compute it. */
int oracle_bin_of_k2(int64_t dims, int nrbins, int64_t k2)
{
    const double binsperunit = (nrbins - 1) / log(sqrt(3) * dims / 2.0);
    const double kk = sqrt((double)k2);
    return (int)floor(binsperunit * log(kk));
}

/* Unnormalised forward r2c DFT by direct summation, O(n^6): only for the 4^3
 * known answer of test.cpp:59-86, where FFTW3 (absent here) is the reference's
 * transform.  in: padded real [n][n][2(n/2+1)]; out: complex [n][n][n/2+1]. */
int oracle_naive_r2c_3d(int n, const double *in, double *out)
{
    const int nc = n / 2 + 1, fd = 2 * nc;
    for (int a = 0; a < n; a++)
    for (int b = 0; b < n; b++)
    for (int c = 0; c < nc; c++) {
        double re = 0, im = 0;
        for (int x = 0; x < n; x++)
        for (int y = 0; y < n; y++)
        for (int z = 0; z < n; z++) {
            const double ph = -2.0 * M_PI * ((double)(a * x + b * y + c * z) / n);
            const double v = in[((size_t)x * n + y) * fd + z];
            re += v * cos(ph);
            im += v * sin(ph);
        }
        out[2 * (((size_t)a * n + b) * nc + c)] = re;
        out[2 * (((size_t)a * n + b) * nc + c) + 1] = im;
    }
    return 0;
}
