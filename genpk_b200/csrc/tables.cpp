// Host-built plan-time tables for the fused binning kernel.
//
// Two things on the binning path depend on the host libm in a way that mode
// counts only stay bit-exact if the GPU never re-derives them:
//   * the log-|k| bin of a mode, floor(binsperunit*log(kk)) (powerspectrum.c:38,66,75,84),
//     which is a pure function of the integer k2 = ki^2+kj^2+kz^2;
//   * the float32-rounded per-axis inverse CIC window (fieldize.cpp:117-131).
// Both are evaluated here once per (dims, nrbins) with the same libm the
// reference links, in O(k2max) / O(dims) work, and shipped to the GPU as an
// integer threshold table and a float table.  Compiled WITHOUT -ffast-math and
// with the expressions kept out of line so that gcc cannot vectorise them onto
// libmvec or re-associate them.
#include <math.h>
#include <stdint.h>

#include "common.cuh"

#ifndef M_PI
#define M_PI 3.1415926535897932384626433832795
#endif

namespace genpk {

// fieldize.cpp:117-121, narrowed to float as at fieldize.cpp:129-131.
__attribute__((noinline)) float oned_invwindow_f32(int64_t k, int64_t n)
{
    if (k == 0)
        return 1.0f;
    const double w = M_PI * k / (n * sin(M_PI * k / (float)n));
    return (float)w;
}

__attribute__((noinline)) static double binsperunit_of(int dims, int nrbins)
{
    // powerspectrum.c:38
    return (nrbins - 1) / log(sqrt(3.0) * dims / 2.0);
}

static inline int bin_from_bpu(double bpu, int64_t k2, unsigned rule)
{
    if (rule == 0) {
        // The reference's Makefile:30 builds powerspectrum.c with -ffast-math, under
        // which gcc rewrites log(sqrt(x)) as 0.5*log(x) and folds the 0.5 into
        // binsperunit (verified in the disassembly of oracle/_ref).
        const double half_bpu = bpu * 0.5;
        return (int)floor(half_bpu * log((double)k2));
    }
    return (int)floor(bpu * log(sqrt((double)k2)));
}

int reference_bin_of_k2(int dims, int nrbins, int64_t k2, unsigned rule)
{
    return bin_from_bpu(binsperunit_of(dims, nrbins), k2, rule);
}

int build_bin_tables(int dims, int nrbins, unsigned rule, BinTables *out)
{
    if (dims < 1 || nrbins < 1) {
        set_error("build_bin_tables: bad dims=%d nrbins=%d", dims, nrbins);
        return 1;
    }
    out->dims = dims;
    out->nrbins = nrbins;
    out->rule = rule;
    const int half = dims / 2;
    out->iw1d.resize(half + 1);
    for (int k = 0; k <= half; k++)
        out->iw1d[k] = oned_invwindow_f32(k, dims);

    // Largest |k|^2 on the grid: each signed component lies in [-(dims-1)/2, dims/2].
    const int64_t k2max = 3 * (int64_t)half * half;
    out->thresh.assign((size_t)nrbins + 1, (uint32_t)(k2max + 1));
    out->monotone = true;
    int prev = -1;
    const double bpu = binsperunit_of(dims, nrbins);
    for (int64_t k2 = 1; k2 <= k2max; k2++) {
        const int b = bin_from_bpu(bpu, k2, rule);
        if (b < 0 || b >= nrbins) {
            set_error("bin %d of k2=%lld outside [0,%d) (the reference asserts here, powerspectrum.c:67)",
                      b, (long long)k2, nrbins);
            return 1;
        }
        if (b < prev) {
            out->monotone = false;
            set_error("bin rule not monotone in k2 at dims=%d nrbins=%d k2=%lld", dims, nrbins, (long long)k2);
            return 1;
        }
        // every bin in (prev, b] starts at this k2
        for (int t = prev + 1; t <= b; t++)
            out->thresh[t] = (uint32_t)k2;
        prev = b;
    }
    // bins above the last populated one keep k2max+1 (never reached)
    return 0;
}

}  // namespace genpk
