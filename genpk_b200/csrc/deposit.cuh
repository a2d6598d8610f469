// Shared pieces of the CIC deposit kernels (deposit.cu, deposit_march.cu).
#pragma once
#include "common.cuh"

namespace genpk {

struct DepositArgs {
    const float *pos;
    const double *pos64;   // double-precision positions used un-narrowed (DOUBLE_PRECISION_SNAP, gen-pk.h:25-29), else null
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
    int *touched;          // slab: {lowest, highest} local x plane any deposit wrote since the grid was cleared (null: not tracked)
};

// deposit_march.cu
struct OrderInfo {
    int coherent;         // consecutive particles are near each other on the grid
    int lattice;          // 1: lattice order with row length n0 (and n1 rows per plane when > 0)
    long long n0, n1;
    int score_z, score_y, score_x, samples;   // per mille of exact one-cell hand-overs, diagnostics
    // lattice only: where the particles of lattice plane x sit along x, in (local) grid planes, relative to
    // floor((x + 1/2) * planes_per_lattice_plane): rounded mean and largest deviation from it over the samples
    int dx_valid, dx_mean, dx_dev;
};
int probe_order(genpk_ctx *ctx, const float *pos, int64_t n, double units, OrderInfo *info);
int launch_march(genpk_ctx *ctx, const DepositArgs &a, long long n0, long long n1);
// deposit_sweep.cu.  za: the grid has NOT been cleared, the launch clears it ahead of its own front.
// Returns with *za_possible false (and nothing launched) when za was asked for but this lattice cannot
// be swept in one resident wave; the caller then clears the grid and calls again with za = false.
int launch_sweep(genpk_ctx *ctx, const DepositArgs &a, long long n0, long long n1, bool za, const OrderInfo *info,
                 bool *za_possible);

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

// The same for a double-precision position (a DOUBLE_PRECISION_SNAP build of the reference: GENPK_FLOAT_TYPE
// is double and fieldize.cpp:66 multiplies the double itself).
__device__ __forceinline__ AxisCell axis_cell(double p, double units, int dims)
{
    AxisCell c;
    const double x = __dmul_rn(p, units);                   // fieldize.cpp:66
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

// One particle, eight reductions: the arithmetic of deposit_direct_kernel (fieldize.cpp:63-92) for the
// stragglers of the sweep kernel.  Rejected (non-finite, outside the slab) particles are skipped.
// TRACK: record the planes written in a.touched with atomics (callers that deposit many particles this way
// track the range themselves: two same-address atomics per particle serialise in one L2 slice).
template <bool FIXED, bool TRACK = true>
__device__ __forceinline__ void deposit_single(const DepositArgs &a, float px, float py, float pz, double m)
{
    const AxisCell cx = axis_cell(px, a.units, a.dims);
    const AxisCell cy = axis_cell(py, a.units, a.dims);
    const AxisCell cz = axis_cell(pz, a.units, a.dims);
    if (!(cx.ok && cy.ok && cz.ok))
        return;
    int xl = cx.lo, xh = cx.hi;
    if (a.slab) {
        xl = slab_plane(cx.lo, a.x0, a.ghost_lo, a.dims);
        if (xl < 0 || xl > a.xl_max)
            return;
        xh = xl + 1;
        if (TRACK && a.touched) {
            atomicMin(a.touched, xl);
            atomicMax(a.touched + 1, xh);
        }
    }
    const double mx0 = __dmul_rn(m, cx.wl), mx1 = __dmul_rn(m, cx.wh);
    const double w00 = __dmul_rn(mx0, cy.wl), w10 = __dmul_rn(mx1, cy.wl);
    const double w01 = __dmul_rn(mx0, cy.wh), w11 = __dmul_rn(mx1, cy.wh);
    const size_t bx0 = a.plane * (size_t)xl, bx1 = a.plane * (size_t)xh;
    const size_t by0 = (size_t)a.fd * cy.lo, by1 = (size_t)a.fd * cy.hi;
    Acc<FIXED>::red(a.grid, bx0 + by0 + cz.lo, Acc<FIXED>::make(__dmul_rn(w00, cz.wl), a.scale));
    Acc<FIXED>::red(a.grid, bx1 + by0 + cz.lo, Acc<FIXED>::make(__dmul_rn(w10, cz.wl), a.scale));
    Acc<FIXED>::red(a.grid, bx0 + by1 + cz.lo, Acc<FIXED>::make(__dmul_rn(w01, cz.wl), a.scale));
    Acc<FIXED>::red(a.grid, bx1 + by1 + cz.lo, Acc<FIXED>::make(__dmul_rn(w11, cz.wl), a.scale));
    Acc<FIXED>::red(a.grid, bx0 + by0 + cz.hi, Acc<FIXED>::make(__dmul_rn(w00, cz.wh), a.scale));
    Acc<FIXED>::red(a.grid, bx1 + by0 + cz.hi, Acc<FIXED>::make(__dmul_rn(w10, cz.wh), a.scale));
    Acc<FIXED>::red(a.grid, bx0 + by1 + cz.hi, Acc<FIXED>::make(__dmul_rn(w01, cz.wh), a.scale));
    Acc<FIXED>::red(a.grid, bx1 + by1 + cz.hi, Acc<FIXED>::make(__dmul_rn(w11, cz.wh), a.scale));
}

// Out-of-box cell index -> [0, dims) (fieldize.cpp:70-75); kept out of line, it is rare.
static __device__ __noinline__ int wrap_cell(int f, int dims)
{
    f %= dims;
    return f < 0 ? f + dims : f;
}

// Same arithmetic as axis_cell (x = p*units, f = floor(x), d = x - f, t = 1 - d) with the
// floor taken by a round-down add of 1.5*2^52: the low word of the sum is floor(x) as an
// int and sum - magic is floor(x) as a double, both exact for |x| < 2^31.  Two additions
// instead of three 64-bit conversions, which run at a fraction of the FP64 add rate.
__device__ __forceinline__ void axis_fast(float p, double units, int &f, double &wl, double &wh)
{
    const double magic = 6755399441055744.0;
    const double x = __dmul_rn((double)p, units);           // fieldize.cpp:66
    const double t = __dadd_rd(x, magic);
    f = __double2loint(t);                                  // :67
    wh = __dsub_rn(x, __dsub_rn(t, magic));                 // :68  dx
    wl = __dsub_rn(1.0, wh);                                // :69  tx
}

// The same with the high word of x + 1.5*2^52, which is 0x43380000 exactly when 0 <= x < 2^32.
__device__ __forceinline__ void axis_fast_hi(float p, double units, int &f, int &hi, double &wl, double &wh)
{
    const double magic = 6755399441055744.0;
    const double x = __dmul_rn((double)p, units);           // fieldize.cpp:66
    const double t = __dadd_rd(x, magic);
    f = __double2loint(t);                                  // :67
    hi = __double2hiint(t);
    wh = __dsub_rn(x, __dsub_rn(t, magic));                 // :68  dx
    wl = __dsub_rn(1.0, wh);                                // :69  tx
}

// lo += from when take (one predicated add instead of add + two selects)
__device__ __forceinline__ void add_if(double &lo, double from, int take)
{
    asm("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %2, 0;\n\t@p add.rn.f64 %0, %0, %1;\n\t}" : "+d"(lo) : "d"(from), "r"(take));
}
__device__ __forceinline__ void add_if(long long &lo, long long from, int take)
{
    asm("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %2, 0;\n\t@p add.s64 %0, %0, %1;\n\t}" : "+l"(lo) : "l"(from), "r"(take));
}

}  // namespace genpk
