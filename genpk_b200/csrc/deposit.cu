// Cloud-in-cell mass assignment on the GPU -- replaces fieldize() (fieldize.cpp:46-114).
//
// Numerics follow the reference exactly: per axis
//   x = (double)pos_f32 * (dims/box);  f = floor(x);  d = x - f;  t = 1 - d
//   cells f mod dims and (f+1) mod dims (any finite position wraps periodically)
//   weight = ((m*ax)*ay)*az with ax in {t,d}
// accumulated into the padded grid at (dims*fd)*X + fd*Y + Z.
//
// Accumulation is order independent in fixed-point mode (int64 adds of
// llrint(w*2^S)) and ordinary fp64 red.add otherwise.
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
    int x0, nx, ghost;     // slab: owned planes [x0, x0+nx), ghost plane stored at local index nx
    size_t plane;          // doubles per x plane = dims*fd
    void *grid;
    unsigned long long *errors;
    const int *use_sorted; // optional device flag: 1 = read sorted_pos/sorted_mass instead of pos/mass
    const float *sorted_pos;
    const float *sorted_mass;
};

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

// One particle per thread, CTA b owns particles [256 b, 256 b + 256): consecutive
// particles stay consecutive in time, which keeps brick-sorted and snapshot-ordered
// input L2-local.  Eight reductions per particle, minus the ones merged below.
//
// Neighbour merge: in lattice / snapshot order the next particle usually sits one
// cell further along z, so its four low-z corners are this particle's four high-z
// corners.  Those four contributions are handed to the next lane by shuffle and
// leave with its reductions: ~4 instead of 8 red.add per particle on coherent input,
// unchanged results (the sum per cell is the same set of terms).
template <bool FIXED>
__global__ void __launch_bounds__(256) deposit_direct_kernel(DepositArgs a)
{
    typedef typename Acc<FIXED>::type acc_t;
    const int lane = threadIdx.x & 31;
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = p < a.n;
    const float *pos = a.pos;
    const float *mass = a.mass;
    if (a.use_sorted && *a.use_sorted) {       // device-side choice made by the coherence probe
        pos = a.sorted_pos;
        mass = a.mass ? a.sorted_mass : nullptr;
    }
    float px = 0.f, py = 0.f, pz = 0.f;
    double m = a.cmass;
    if (live) {
        px = pos[3 * p];
        py = pos[3 * p + 1];
        pz = pos[3 * p + 2];
        if (mass)
            m = (double)mass[p];                                          // fieldize.cpp:63
    }
    const AxisCell cx = axis_cell(px, a.units, a.dims);
    const AxisCell cy = axis_cell(py, a.units, a.dims);
    const AxisCell cz = axis_cell(pz, a.units, a.dims);
    int xl = cx.lo - a.x0, xh;
    bool ok = live && cx.ok && cy.ok && cz.ok;
    if (a.ghost) {                           // slab: the +1 neighbour may be the ghost plane
        ok = ok && xl >= 0 && xl < a.nx;
        xh = xl + 1;
    } else {
        xh = cx.hi;
    }
    if (live && !ok)
        atomicAdd(a.errors, 1ull);
    const double mx0 = __dmul_rn(m, cx.wl), mx1 = __dmul_rn(m, cx.wh);
    const double w00 = __dmul_rn(mx0, cy.wl), w10 = __dmul_rn(mx1, cy.wl);
    const double w01 = __dmul_rn(mx0, cy.wh), w11 = __dmul_rn(mx1, cy.wh);
    // weights in the order of fieldize.cpp:77-84
    acc_t lo[4], hi[4];
    lo[0] = Acc<FIXED>::make(__dmul_rn(w00, cz.wl), a.scale);
    lo[1] = Acc<FIXED>::make(__dmul_rn(w10, cz.wl), a.scale);
    lo[2] = Acc<FIXED>::make(__dmul_rn(w01, cz.wl), a.scale);
    lo[3] = Acc<FIXED>::make(__dmul_rn(w11, cz.wl), a.scale);
    hi[0] = Acc<FIXED>::make(__dmul_rn(w00, cz.wh), a.scale);
    hi[1] = Acc<FIXED>::make(__dmul_rn(w10, cz.wh), a.scale);
    hi[2] = Acc<FIXED>::make(__dmul_rn(w01, cz.wh), a.scale);
    hi[3] = Acc<FIXED>::make(__dmul_rn(w11, cz.wh), a.scale);

    // (x,y) column id and z cells; -1/-2 never match
    const long long col = ((long long)cx.lo * a.dims + cy.lo) * a.dims;
    const long long key_lo = ok ? col + cz.lo : -1;
    const long long key_hi = ok ? col + cz.hi : -2;
    const long long next_lo = __shfl_down_sync(0xffffffffu, key_lo, 1);
    const bool give = lane < 31 && next_lo == key_hi;      // my high-z corners are the next lane's low-z corners
    const bool take = __shfl_up_sync(0xffffffffu, (int)give, 1) && lane > 0;
#pragma unroll
    for (int c = 0; c < 4; c++) {
        const acc_t from_prev = __shfl_up_sync(0xffffffffu, hi[c], 1);
        if (take)
            lo[c] += from_prev;
    }
    if (!ok)
        return;
    const size_t bx0 = a.plane * (size_t)xl, bx1 = a.plane * (size_t)xh;
    const size_t by0 = (size_t)a.fd * cy.lo, by1 = (size_t)a.fd * cy.hi;
    // cell order of fieldize.cpp:85-92
    Acc<FIXED>::red(a.grid, bx0 + by0 + cz.lo, lo[0]);
    Acc<FIXED>::red(a.grid, bx1 + by0 + cz.lo, lo[1]);
    Acc<FIXED>::red(a.grid, bx0 + by1 + cz.lo, lo[2]);
    Acc<FIXED>::red(a.grid, bx1 + by1 + cz.lo, lo[3]);
    if (!give) {
        Acc<FIXED>::red(a.grid, bx0 + by0 + cz.hi, hi[0]);
        Acc<FIXED>::red(a.grid, bx1 + by0 + cz.hi, hi[1]);
        Acc<FIXED>::red(a.grid, bx0 + by1 + cz.hi, hi[2]);
        Acc<FIXED>::red(a.grid, bx1 + by1 + cz.hi, hi[3]);
    }
}

// Spatial coherence probe: are consecutive particles near each other on the grid?
// Samples pairs (i, i+1) across the run; writes *use_sorted = 0 when at least 60 % of
// them are within 4 cells in every axis (lattice, snapshot or already-sorted order:
// sorting would only cost bandwidth), else 1.  The decision stays on the device: the
// sort kernels return at once when it is 0 and the deposit reads the original array.
__global__ void __launch_bounds__(1024) coherence_probe_kernel(const float *pos, int64_t n, double units, int dims,
                                                               int *use_sorted)
{
    __shared__ int s_near;
    if (threadIdx.x == 0)
        s_near = 0;
    __syncthreads();
    const int samples = 8 * 1024;
    int near = 0, tried = 0;
    for (int s = threadIdx.x; s < samples; s += blockDim.x) {
        const int64_t i = (int64_t)((double)s / samples * (double)(n - 1));
        if (i + 1 >= n)
            continue;
        tried++;
        bool all = true;
        for (int ax = 0; ax < 3; ax++) {
            const AxisCell c0 = axis_cell(pos[3 * i + ax], units, dims);
            const AxisCell c1 = axis_cell(pos[3 * (i + 1) + ax], units, dims);
            int d = abs(c0.lo - c1.lo);
            d = min(d, dims - d);
            all = all && d <= 4;
        }
        near += all ? 1 : 0;
    }
    atomicAdd(&s_near, near);
    __syncthreads();
    if (threadIdx.x == 0) {
        const int total = (int)min((int64_t)samples, n > 1 ? n - 1 : 0);
        *use_sorted = (total > 0 && s_near * 10 >= total * 6) ? 0 : 1;
    }
    (void)tried;
}

// ---------------------------------------------------------------------------------
// Counting sort of particles into bricks whose grid footprint fits in L2, so that
// the reductions of the deposit kernel hit L2-resident lines instead of issuing a
// DRAM read-modify-write per 32-byte sector.
// ---------------------------------------------------------------------------------
struct BrickMap {
    double units;
    int dims;
    int bx, by;        // brick extent in x planes and y rows
    int nby;           // bricks along y
    int nbricks;
};

__device__ __forceinline__ int brick_of(const BrickMap &bm, float px, float py)
{
    const AxisCell cx = axis_cell(px, bm.units, bm.dims);
    const AxisCell cy = axis_cell(py, bm.units, bm.dims);
    return (cx.lo / bm.bx) * bm.nby + (cy.lo / bm.by);
}

constexpr int SORT_THREADS = 512;
constexpr int SORT_ITEMS = 8;
constexpr int SORT_MAX_KEYS = 4096;   // 12 B of shared memory per key in the scatter

__global__ void __launch_bounds__(SORT_THREADS) brick_histogram_kernel(const float *pos, int64_t n, BrickMap bm,
                                                                      unsigned long long *counts, const int *enabled)
{
    extern __shared__ unsigned s_hist[];
    if (enabled && !*enabled)           // the coherence probe said sorting is not needed
        return;
    for (int k = threadIdx.x; k < bm.nbricks; k += SORT_THREADS)
        s_hist[k] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * SORT_THREADS * SORT_ITEMS;
#pragma unroll
    for (int it = 0; it < SORT_ITEMS; it++) {
        const int64_t p = base + it * SORT_THREADS + threadIdx.x;
        if (p < n)
            atomicAdd(&s_hist[brick_of(bm, pos[3 * p], pos[3 * p + 1])], 1u);
    }
    __syncthreads();
    for (int k = threadIdx.x; k < bm.nbricks; k += SORT_THREADS)
        if (s_hist[k])
            atomicAdd(&counts[k], (unsigned long long)s_hist[k]);
}

// Exclusive scan of at most SORT_MAX_KEYS counters, in place (one CTA).
__global__ void __launch_bounds__(1024) exclusive_scan_kernel(unsigned long long *v, int n)
{
    __shared__ unsigned long long s_part[1024];
    const int per = (n + 1023) / 1024;
    const int lo = threadIdx.x * per, hi = min(lo + per, n);
    unsigned long long sum = 0;
    for (int i = lo; i < hi; i++)
        sum += v[i];
    s_part[threadIdx.x] = sum;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {
        unsigned long long t = threadIdx.x >= off ? s_part[threadIdx.x - off] : 0;
        __syncthreads();
        s_part[threadIdx.x] += t;
        __syncthreads();
    }
    unsigned long long run = s_part[threadIdx.x] - sum;
    for (int i = lo; i < hi; i++) {
        const unsigned long long c = v[i];
        v[i] = run;
        run += c;
    }
}

__global__ void __launch_bounds__(SORT_THREADS) brick_scatter_kernel(const float *pos, const float *mass, int64_t n,
                                                                    BrickMap bm, unsigned long long *cursors,
                                                                    float *out_pos, float *out_mass, const int *enabled)
{
    extern __shared__ unsigned long long s_mem[];
    if (enabled && !*enabled)
        return;
    unsigned long long *s_base = s_mem;                                  // [nbricks]
    unsigned *s_cnt = reinterpret_cast<unsigned *>(s_mem + bm.nbricks);  // [nbricks]
    for (int k = threadIdx.x; k < bm.nbricks; k += SORT_THREADS)
        s_cnt[k] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * SORT_THREADS * SORT_ITEMS;
    float px[SORT_ITEMS], py[SORT_ITEMS], pz[SORT_ITEMS];
    int key[SORT_ITEMS];
    unsigned rank[SORT_ITEMS];
#pragma unroll
    for (int it = 0; it < SORT_ITEMS; it++) {
        const int64_t p = base + it * SORT_THREADS + threadIdx.x;
        key[it] = -1;
        if (p < n) {
            px[it] = pos[3 * p];
            py[it] = pos[3 * p + 1];
            pz[it] = pos[3 * p + 2];
            key[it] = brick_of(bm, px[it], py[it]);
            rank[it] = atomicAdd(&s_cnt[key[it]], 1u);
        }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < bm.nbricks; k += SORT_THREADS)
        if (s_cnt[k])
            s_base[k] = atomicAdd(&cursors[k], (unsigned long long)s_cnt[k]);
    __syncthreads();
#pragma unroll
    for (int it = 0; it < SORT_ITEMS; it++) {
        if (key[it] >= 0) {
            const unsigned long long dst = s_base[key[it]] + rank[it];
            out_pos[3 * dst] = px[it];
            out_pos[3 * dst + 1] = py[it];
            out_pos[3 * dst + 2] = pz[it];
            if (mass)
                out_mass[dst] = mass[base + it * SORT_THREADS + threadIdx.x];
        }
    }
}

__global__ void fixed_to_double_kernel(double *grid, size_t n, double inv_scale)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const long long q = reinterpret_cast<const long long *>(grid)[i];
        grid[i] = __dmul_rn((double)q, inv_scale);
    }
}

static int launch_direct(genpk_ctx *ctx, const DepositArgs &a)
{
    if (a.n <= 0)
        return 0;
    const int threads = 256;
    const int64_t want = (a.n + threads - 1) / threads;
    if (want > 0x7fffffffLL) {
        set_error("deposit: %lld particles in one call exceed the launch grid; split the call", (long long)a.n);
        return 1;
    }
    const int blocks = (int)want;
    if (ctx->fixed)
        deposit_direct_kernel<true><<<blocks, threads, 0, ctx->stream>>>(a);
    else
        deposit_direct_kernel<false><<<blocks, threads, 0, ctx->stream>>>(a);
    ctx->launches++;
    GENPK_CUDA_OK(cudaGetLastError());
    return 0;
}

static BrickMap choose_bricks(const genpk_ctx *ctx, double units)
{
    // A brick's footprint (its planes/rows plus the +1 halo touched by CIC)
    // should stay within a fraction of L2.
    const SlabGeom &g = ctx->g;
    const size_t budget = ctx->l2_bytes ? ctx->l2_bytes / 3 : (size_t)40 << 20;
    const size_t plane_bytes = g.plane() * sizeof(double);
    BrickMap bm;
    bm.units = units;
    bm.dims = g.dims;
    if (2 * plane_bytes <= budget) {
        bm.by = g.dims;
        bm.bx = (int)(budget / plane_bytes) - 1;
        if (bm.bx > g.dims)
            bm.bx = g.dims;
    } else {
        bm.bx = 1;
        const size_t row_bytes = (size_t)g.fd * sizeof(double) * 2;   // two planes per row
        bm.by = (int)(budget / row_bytes) - 1;
        if (bm.by < 1)
            bm.by = 1;
        if (bm.by > g.dims)
            bm.by = g.dims;
    }
    bm.nby = (g.dims + bm.by - 1) / bm.by;
    const int nbx = (g.dims + bm.bx - 1) / bm.bx;     // global x bricks (slab ranks see a subset)
    bm.nbricks = nbx * bm.nby;
    return bm;
}

// Counting sort of a particle run by brick id into out_pos/out_mass.  When
// counts_out is given it receives the per-brick particle counts (int64).
__global__ void copy_counts_kernel(const unsigned long long *in, int64_t *out, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
        out[i] = (int64_t)in[i];
}

static int sort_by_brick(genpk_ctx *ctx, const float *pos, const float *mass, int64_t n, const BrickMap &bm,
                         float *out_pos, float *out_mass, int64_t *counts_out, const int *enabled)
{
    if (bm.nbricks > ctx->brick_cap) {
        if (ctx->d_brick_counts) cudaFree(ctx->d_brick_counts);
        ctx->d_brick_counts = nullptr;
        ctx->brick_cap = 0;
        GENPK_CUDA_OK(cudaMalloc(&ctx->d_brick_counts, (size_t)bm.nbricks * sizeof(unsigned long long)));
        ctx->brick_cap = bm.nbricks;
    }
    unsigned long long *counts = reinterpret_cast<unsigned long long *>(ctx->d_brick_counts);
    GENPK_CUDA_OK(cudaMemsetAsync(counts, 0, (size_t)bm.nbricks * sizeof(unsigned long long), ctx->stream));
    const int64_t per_block = (int64_t)SORT_THREADS * SORT_ITEMS;
    const int blocks = (int)((n + per_block - 1) / per_block);
    if (blocks > 0) {
        brick_histogram_kernel<<<blocks, SORT_THREADS, bm.nbricks * sizeof(unsigned), ctx->stream>>>(pos, n, bm, counts, enabled);
        ctx->launches++;
    }
    if (counts_out) {
        copy_counts_kernel<<<(bm.nbricks + 255) / 256, 256, 0, ctx->stream>>>(counts, counts_out, bm.nbricks);
        ctx->launches++;
    }
    exclusive_scan_kernel<<<1, 1024, 0, ctx->stream>>>(counts, bm.nbricks);
    ctx->launches++;
    if (blocks > 0) {
        const size_t smem = (size_t)bm.nbricks * (sizeof(unsigned long long) + sizeof(unsigned));
        brick_scatter_kernel<<<blocks, SORT_THREADS, smem, ctx->stream>>>(pos, mass, n, bm, counts, out_pos,
                                                                          mass ? out_mass : nullptr, enabled);
        ctx->launches++;
    }
    GENPK_CUDA_OK(cudaGetLastError());
    return 0;
}

static int ensure_sorted_scratch(genpk_ctx *ctx, int64_t n, bool with_mass)
{
    if (n > ctx->sorted_cap) {
        if (ctx->d_sorted_pos) cudaFree(ctx->d_sorted_pos);
        if (ctx->d_sorted_mass) cudaFree(ctx->d_sorted_mass);
        ctx->d_sorted_pos = nullptr;
        ctx->d_sorted_mass = nullptr;
        ctx->sorted_cap = 0;
        GENPK_CUDA_OK(cudaMalloc(&ctx->d_sorted_pos, (size_t)n * 3 * sizeof(float)));
        ctx->sorted_cap = n;
    }
    if (with_mass && !ctx->d_sorted_mass)
        GENPK_CUDA_OK(cudaMalloc(&ctx->d_sorted_mass, (size_t)ctx->sorted_cap * sizeof(float)));
    if (!ctx->d_use_sorted)
        GENPK_CUDA_OK(cudaMalloc(&ctx->d_use_sorted, sizeof(int)));
    return 0;
}

// Destination rank of each particle = owner of x plane floor(x*dims/box) mod dims.
int route_particles(genpk_ctx *ctx, const float *pos, const float *mass, int64_t n, double boxsize,
                    float *spos, float *smass, int64_t *counts)
{
    const SlabGeom &g = ctx->g;
    BrickMap bm;
    bm.units = g.dims / boxsize;
    bm.dims = g.dims;
    bm.bx = g.nx;
    bm.by = g.dims;
    bm.nby = 1;
    bm.nbricks = g.nranks;
    return sort_by_brick(ctx, pos, mass, n, bm, spos, smass, counts, nullptr);
}

int deposit_device(genpk_ctx *ctx, int which, const float *pos, const float *masses, int64_t n,
                   double mass, double boxsize)
{
    const SlabGeom &g = ctx->g;
    if (n <= 0)
        return 0;
    if (!(boxsize > 0)) {
        set_error("deposit: boxsize must be positive");
        return 1;
    }
    DepositArgs a;
    a.pos = pos;
    a.mass = masses;
    a.n = n;
    a.cmass = mass;
    a.units = g.dims / boxsize;                       // fieldize.cpp:52
    a.scale = ldexp(1.0, ctx->scale_bits);
    a.dims = g.dims;
    a.fd = g.fd;
    a.x0 = g.x0;
    a.nx = g.nx;
    a.ghost = g.ghost;
    a.plane = g.plane();
    a.grid = ctx->grid[which];
    a.errors = ctx->d_errors;
    a.use_sorted = nullptr;
    a.sorted_pos = nullptr;
    a.sorted_mass = nullptr;
    if (ctx->fixed)
        ctx->grid_is_fixed[which] = true;

    // DIRECT: reductions straight from the caller's order.  SORTED: always brick-sort
    // first.  AUTO: nothing to gain from sorting when the owned grid fits in L2;
    // otherwise a device-side probe of spatial coherence decides, without a host sync.
    int mode = ctx->deposit_mode;
    const size_t grid_bytes = g.grid_doubles() * sizeof(double);
    const bool fits_l2 = ctx->l2_bytes && grid_bytes <= ctx->l2_bytes / 2;
    if (mode == GENPK_DEPOSIT_TILED)
        mode = GENPK_DEPOSIT_AUTO;
    if (mode == GENPK_DEPOSIT_AUTO && fits_l2)
        mode = GENPK_DEPOSIT_DIRECT;
    if (mode != GENPK_DEPOSIT_DIRECT) {
        const BrickMap bm = choose_bricks(ctx, a.units);
        if (bm.nbricks > 1 && bm.nbricks <= SORT_MAX_KEYS) {
            stage_begin(ctx, ST_SORT);
            if (int rc = ensure_sorted_scratch(ctx, n, masses != nullptr))
                return rc;
            const int *enabled = nullptr;
            if (mode == GENPK_DEPOSIT_AUTO) {
                coherence_probe_kernel<<<1, 1024, 0, ctx->stream>>>(pos, n, a.units, g.dims, ctx->d_use_sorted);
                ctx->launches++;
                enabled = ctx->d_use_sorted;
            }
            if (int rc = sort_by_brick(ctx, pos, masses, n, bm, ctx->d_sorted_pos, ctx->d_sorted_mass, nullptr, enabled))
                return rc;
            stage_end(ctx, ST_SORT);
            if (enabled) {
                a.use_sorted = enabled;
                a.sorted_pos = ctx->d_sorted_pos;
                a.sorted_mass = ctx->d_sorted_mass;
            } else {
                a.pos = ctx->d_sorted_pos;
                a.mass = masses ? ctx->d_sorted_mass : nullptr;
            }
        }
    }
    return launch_direct(ctx, a);
}

int fixed_to_double(genpk_ctx *ctx, int which)
{
    if (!ctx->grid_is_fixed[which])
        return 0;
    const size_t n = ctx->g.grid_doubles();
    fixed_to_double_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(ctx->grid[which], n,
                                                                      ldexp(1.0, -ctx->scale_bits));
    ctx->launches++;
    GENPK_CUDA_OK(cudaGetLastError());
    ctx->grid_is_fixed[which] = false;
    return 0;
}

// Host-array shim behind genpk_fieldize(): standalone grid with arbitrary z stride.
int fieldize_host_shim(double boxsize, int dims, double *out, int64_t n, const float *positions,
                       const float *masses, double mass, int extra)
{
    if (n <= 0)
        return 0;
    if (dims < 1 || !(boxsize > 0) || (extra != 0 && extra != 1)) {
        set_error("fieldize: bad arguments dims=%d box=%g extra=%d", dims, boxsize, extra);
        return 1;
    }
    const int fd = 2 * (dims / 2 + extra);              // fieldize.cpp:48
    const size_t cells = (size_t)dims * dims * fd;
    double *d_grid = nullptr;
    float *d_pos = nullptr, *d_mass = nullptr;
    unsigned long long *d_err = nullptr;
    int rc = 1;
    do {
        if (cudaMalloc(&d_grid, cells * sizeof(double)) != cudaSuccess) break;
        if (cudaMalloc(&d_pos, (size_t)n * 3 * sizeof(float)) != cudaSuccess) break;
        if (masses && cudaMalloc(&d_mass, (size_t)n * sizeof(float)) != cudaSuccess) break;
        if (cudaMalloc(&d_err, sizeof(unsigned long long)) != cudaSuccess) break;
        if (cudaMemset(d_err, 0, sizeof(unsigned long long)) != cudaSuccess) break;
        if (cudaMemcpy(d_grid, out, cells * sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess) break;
        if (cudaMemcpy(d_pos, positions, (size_t)n * 3 * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) break;
        if (masses && cudaMemcpy(d_mass, masses, (size_t)n * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) break;
        DepositArgs a;
        a.pos = d_pos;
        a.mass = d_mass;
        a.n = n;
        a.cmass = mass;
        a.units = dims / boxsize;
        a.scale = 1.0;
        a.dims = dims;
        a.fd = fd;
        a.x0 = 0;
        a.nx = dims;
        a.ghost = 0;
        a.plane = (size_t)dims * fd;
        a.grid = d_grid;
        a.errors = d_err;
        a.use_sorted = nullptr;
        a.sorted_pos = nullptr;
        a.sorted_mass = nullptr;
        const int threads = 256;
        const int64_t blocks = (n + threads - 1) / threads;
        if (blocks > 0x7fffffffLL) break;
        deposit_direct_kernel<false><<<(int)blocks, threads>>>(a);
        if (cudaGetLastError() != cudaSuccess) break;
        if (cudaMemcpy(out, d_grid, cells * sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess) break;
        rc = 0;
    } while (0);
    if (rc) {
        cudaError_t e = cudaGetLastError();
        set_error("genpk_fieldize: CUDA failure (%s)", cudaGetErrorString(e));
    }
    cudaFree(d_grid);
    cudaFree(d_pos);
    cudaFree(d_mass);
    cudaFree(d_err);
    return rc;
}

}  // namespace genpk
