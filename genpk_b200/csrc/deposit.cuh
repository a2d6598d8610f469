// Shared pieces of the CIC deposit kernels (deposit.cu, deposit_march.cu).
#pragma once
#include "common.cuh"

namespace genpk {

struct DepositArgs {
    const float *pos;
    const float *mass;     // may be null
    int64_t n;
    double cmass;
    double units;          // dims / boxsize
    double scale;          // 2^scale_bits (fixed-point mode)
    int dims, fd;
    int x0;                // slab: first owned global x plane
    int slab;              // 1: the grid is a slab with ghost planes, 0: the whole periodic grid
    int ghost_lo;          // ghost planes below the owned ones; local plane = (X - x0) + ghost_lo
    int xl_max;            // largest admissible local plane of the low-x corner (its +1 neighbour must exist)
    size_t plane;          // doubles per x plane = dims*fd
    void *grid;
    unsigned long long *errors;
};

// deposit_march.cu
struct OrderInfo {
    int coherent;         // consecutive particles are near each other on the grid
    int lattice;          // 1: lattice order with row length n0 (and n1 rows per plane when > 0)
    long long n0, n1;
    int score_z, score_y, score_x, samples;   // per mille of exact one-cell hand-overs, diagnostics
};
int probe_order(genpk_ctx *ctx, const float *pos, int64_t n, double units, OrderInfo *info);
int launch_march(genpk_ctx *ctx, const DepositArgs &a, long long n0, long long n1);

struct AxisCell {
    int lo, hi;
    double wl, wh;
    bool ok;
};

__device__ __forceinline__ AxisCell axis_cell(float p, double units, int dims)
{
    AxisCell c;
    const double x = __dmul_rn((double)p, units);           // fieldize.cpp:66
    const double fl = floor(x);                             // :67
    c.wh = __dsub_rn(x, fl);                                // :68  dx
    c.wl = __dsub_rn(1.0, c.wh);                            // :69  tx
    c.ok = fabs(x) < 2.0e9;                                 // false for NaN/inf/out of int range
    int f = c.ok ? (int)fl : 0;
    if ((unsigned)f >= (unsigned)dims) {                    // :70-75 periodic wrap, negative fix-up
        f %= dims;
        if (f < 0)
            f += dims;
    }
    c.lo = f;
    c.hi = (f + 1 == dims) ? 0 : f + 1;
    return c;
}

// Local plane index of global cell X in a slab whose first owned plane is x0 and that
// stores ghost_lo ghost planes below it: the periodic image of X - x0 nearest to the slab.
__device__ __forceinline__ int slab_plane(int X, int x0, int ghost_lo, int dims)
{
    int d = X - x0;
    if (d < -ghost_lo)
        d += dims;
    else if (d >= dims - ghost_lo)
        d -= dims;
    return d + ghost_lo;
}

// A contribution is a double (fp64 mode) or llrint(w*2^S) as int64 (fixed-point mode);
// the quantisation happens per contribution, before any merging, so merged
// integer sums are bit-identical to eight separate adds.
template <bool FIXED> struct Acc;
template <> struct Acc<false> {
    typedef double type;
    static __device__ __forceinline__ double make(double w, double) { return w; }
    static __device__ __forceinline__ void red(void *grid, size_t idx, double v)
    {
        atomicAdd(reinterpret_cast<double *>(grid) + idx, v);               // REDG.E.ADD.F64
    }
};
template <> struct Acc<true> {
    typedef long long type;
    static __device__ __forceinline__ long long make(double w, double scale) { return __double2ll_rn(__dmul_rn(w, scale)); }
    static __device__ __forceinline__ void red(void *grid, size_t idx, long long v)
    {
        atomicAdd(reinterpret_cast<unsigned long long *>(grid) + idx, (unsigned long long)v);   // REDG.E.ADD.64
    }
};

}  // namespace genpk
