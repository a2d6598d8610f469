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
#include "deposit.cuh"

namespace genpk {

// One particle per thread, CTA b owns particles [256 b, 256 b + 256): consecutive
// particles stay consecutive in time, which keeps brick-sorted and snapshot-ordered
// input L2-local.  Eight reductions per particle, minus the ones merged below.
//
// Neighbour merge: in lattice / snapshot order the next particle usually sits one
// cell further along z, so its four low-z corners are this particle's four high-z
// corners.  Those four contributions are handed to the next lane by shuffle and
// leave with its reductions: ~4 instead of 8 red.add per particle on coherent input,
// unchanged results (the sum per cell is the same set of terms).
template <bool FIXED, typename pos_t = float>
__global__ void __launch_bounds__(256) deposit_direct_kernel(DepositArgs a)
{
    typedef typename Acc<FIXED>::type acc_t;
    const int lane = threadIdx.x & 31;
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = p < a.n;
    const pos_t *pos = sizeof(pos_t) == 8 ? reinterpret_cast<const pos_t *>(a.pos64) : reinterpret_cast<const pos_t *>(a.pos);
    const float *mass = a.mass;
    pos_t px = 0, py = 0, pz = 0;
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
    int xl = cx.lo, xh;
    bool ok = live && cx.ok && cy.ok && cz.ok;
    if (a.slab) {                            // slab: the +1 neighbour may be a ghost plane
        xl = slab_plane(cx.lo, a.x0, a.ghost_lo, a.dims);
        ok = ok && xl >= 0 && xl <= a.xl_max;
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
                                                                      unsigned long long *counts)
{
    extern __shared__ unsigned s_hist[];
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
                                                                    float *out_pos, float *out_mass)
{
    extern __shared__ unsigned long long s_mem[];
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
    if (a.pos64) {
        if (ctx->fixed)
            deposit_direct_kernel<true, double><<<blocks, threads, 0, ctx->stream>>>(a);
        else
            deposit_direct_kernel<false, double><<<blocks, threads, 0, ctx->stream>>>(a);
    } else if (ctx->fixed)
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
                         float *out_pos, float *out_mass, int64_t *counts_out)
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
        brick_histogram_kernel<<<blocks, SORT_THREADS, bm.nbricks * sizeof(unsigned), ctx->stream>>>(pos, n, bm, counts);
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
                                                                          mass ? out_mass : nullptr);
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
    return sort_by_brick(ctx, pos, mass, n, bm, spos, smass, counts);
}

int deposit_device(genpk_ctx *ctx, int which, const float *pos, const float *masses, int64_t n,
                   double mass, double boxsize, const DepositPlan *plan)
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
    a.pos64 = nullptr;
    a.mass = masses;
    a.n = n;
    a.cmass = mass;
    a.units = g.dims / boxsize;                       // fieldize.cpp:52
    ctx->grid_clean[which] = false;
    if (int rc = latch_scale(ctx, which, masses, n, mass))
        return rc;
    a.scale = ldexp(1.0, ctx->grid_scale_bits[which]);
    a.dims = g.dims;
    a.fd = g.fd;
    a.x0 = g.x0;
    a.slab = g.nranks > 1 ? 1 : 0;
    a.ghost_lo = g.ghost_lo;
    a.xl_max = g.ghost_lo + g.nx + g.ghost_hi - 2;
    a.plane = g.plane();
    a.grid = ctx->grid[which];
    a.errors = ctx->d_errors;
    a.touched = a.slab ? touched_ptr(ctx, which) : nullptr;
    if (ctx->fixed)
        ctx->grid_is_fixed[which] = true;

    // DIRECT: reductions straight from the caller's order.  SORTED: brick-sort first.
    // MARCH: lattice-march kernel (deposit_march.cu).  AUTO: a probe of the head of
    // the stream decides -- lattice order => MARCH, merely coherent (or a grid that
    // fits in L2) => DIRECT, incoherent => SORTED.  The verdict is one small D2H.
    DepositPlan local;
    if (!plan) {
        const bool hit = ctx->plan_key_pos == pos && ctx->plan_key_n == n && ctx->plan_key_box == boxsize &&
                         ctx->plan_key_mode == ctx->deposit_mode * 2 + ctx->sweep;
        if (hit) {
            local.mode = ctx->plan_cached.mode;
            local.n0 = ctx->plan_cached.n0;
            local.n1 = ctx->plan_cached.n1;
            local.have_dx = ctx->plan_cached.have_dx;
            local.dx_mean = ctx->plan_cached.dx_mean;
            local.dx_dev = ctx->plan_cached.dx_dev;
        } else {
            // the grid can be cleared while the probe runs and its verdict travels to the host
            const bool may_sweep = ctx->deposit_mode == GENPK_DEPOSIT_SWEEP || (ctx->deposit_mode == GENPK_DEPOSIT_AUTO && ctx->sweep);
            if (ctx->zero_pending[which] && !(ctx->zero_ahead && ctx->sweep_rx == 0 && may_sweep))
                if (int rc = materialize_zero(ctx, which)) return rc;
            if (int rc = deposit_plan(ctx, pos, n, boxsize, &local))
                return rc;
            ctx->plan_key_pos = pos;
            ctx->plan_key_n = n;
            ctx->plan_key_box = boxsize;
            ctx->plan_key_mode = ctx->deposit_mode * 2 + ctx->sweep;
            ctx->plan_cached.mode = local.mode;
            ctx->plan_cached.n0 = local.n0;
            ctx->plan_cached.n1 = local.n1;
            ctx->plan_cached.have_dx = local.have_dx;
            ctx->plan_cached.dx_mean = local.dx_mean;
            ctx->plan_cached.dx_dev = local.dx_dev;
        }
        plan = &local;
    }
    if (plan->mode == GENPK_DEPOSIT_SWEEP) {
        OrderInfo info = {};
        info.dx_valid = plan->have_dx ? 1 : 0;
        info.dx_mean = plan->dx_mean;
        info.dx_dev = plan->dx_dev;
        // a pending genpk_grid_zero: the sweep clears the grid ahead of its own front when it can
        if (ctx->zero_pending[which] && ctx->zero_ahead && ctx->sweep_rx == 0) {
            bool possible = true;
            if (a.slab)
                if (int rc = touched_set(ctx, which, 0x7fffffff, -1)) return rc;     // cleared grid: nothing touched yet
            if (int rc = launch_sweep(ctx, a, plan->n0, plan->n1, true, &info, &possible))
                return rc;
            if (possible) {
                ctx->zero_pending[which] = false;
                return 0;
            }
        }
        if (int rc = materialize_zero(ctx, which))
            return rc;
        return launch_sweep(ctx, a, plan->n0, plan->n1, false, &info, nullptr);
    }
    if (int rc = materialize_zero(ctx, which))
        return rc;
    if (plan->mode == GENPK_DEPOSIT_MARCH)               // (tracks the planes it writes, like the sweep kernel)
        return launch_march(ctx, a, plan->n0, plan->n1);
    if (a.slab)                                          // the direct kernel does not track what it touches
        if (int rc = touched_set(ctx, which, 0, g.ghost_lo + g.nx + g.ghost_hi - 1)) return rc;
    if (plan->mode == GENPK_DEPOSIT_SORTED) {
        const BrickMap bm = choose_bricks(ctx, a.units);
        if (bm.nbricks > 1 && bm.nbricks <= SORT_MAX_KEYS) {
            {
                StageScope scope(ctx, ST_SORT);
                if (int rc = ensure_sorted_scratch(ctx, n, masses != nullptr))
                    return rc;
                if (int rc = sort_by_brick(ctx, pos, masses, n, bm, ctx->d_sorted_pos, ctx->d_sorted_mass, nullptr))
                    return rc;
            }
            a.pos = ctx->d_sorted_pos;
            a.mass = masses ? ctx->d_sorted_mass : nullptr;
        }
    }
    return launch_direct(ctx, a);
}

// Double-precision positions used as they are (what a DOUBLE_PRECISION_SNAP build of the reference hands to
// fieldize(), read_fieldize.cpp:24-25): one thread per particle, eight reductions minus the z hand-over.
int deposit_device_f64(genpk_ctx *ctx, int which, const double *pos, const float *masses, int64_t n, double mass,
                       double boxsize)
{
    const SlabGeom &g = ctx->g;
    if (n <= 0)
        return 0;
    if (!(boxsize > 0)) {
        set_error("deposit: boxsize must be positive");
        return 1;
    }
    if (int rc = materialize_zero(ctx, which))
        return rc;
    DepositArgs a;
    a.pos = nullptr;
    a.pos64 = pos;
    a.mass = masses;
    a.n = n;
    a.cmass = mass;
    a.units = g.dims / boxsize;
    ctx->grid_clean[which] = false;
    if (int rc = latch_scale(ctx, which, masses, n, mass))
        return rc;
    a.scale = ldexp(1.0, ctx->grid_scale_bits[which]);
    a.dims = g.dims;
    a.fd = g.fd;
    a.x0 = g.x0;
    a.slab = g.nranks > 1 ? 1 : 0;
    a.ghost_lo = g.ghost_lo;
    a.xl_max = g.ghost_lo + g.nx + g.ghost_hi - 2;
    a.plane = g.plane();
    a.grid = ctx->grid[which];
    a.errors = ctx->d_errors;
    a.touched = a.slab ? touched_ptr(ctx, which) : nullptr;
    if (ctx->fixed)
        ctx->grid_is_fixed[which] = true;
    if (a.slab)
        if (int rc = touched_set(ctx, which, 0, g.ghost_lo + g.nx + g.ghost_hi - 1)) return rc;
    return launch_direct(ctx, a);
}

// Which kernels a deposit of these particles will use (the order probe runs here when the
// mode is AUTO or MARCH without a hint).  A plan may be reused for the following chunks of
// the same stream: it is only ever a performance choice, every kernel is exact for any input.
int deposit_plan(genpk_ctx *ctx, const float *pos, int64_t n, double boxsize, DepositPlan *plan)
{
    const SlabGeom &g = ctx->g;
    int mode = ctx->deposit_mode;
    if (mode == GENPK_DEPOSIT_TILED)
        mode = GENPK_DEPOSIT_AUTO;
    const size_t grid_bytes = g.grid_doubles() * sizeof(double);
    const bool fits_l2 = ctx->l2_bytes && grid_bytes <= ctx->l2_bytes / 2;
    long long n0 = ctx->lattice_n0, n1 = ctx->lattice_n1;
    if (mode == GENPK_DEPOSIT_AUTO && n < (1 << 16))
        mode = GENPK_DEPOSIT_DIRECT;
    plan->have_dx = false;
    const bool lattice_mode = mode == GENPK_DEPOSIT_MARCH || mode == GENPK_DEPOSIT_SWEEP;
    if (mode == GENPK_DEPOSIT_AUTO || (lattice_mode && n0 <= 0) || (mode == GENPK_DEPOSIT_SWEEP && ctx->zero_ahead)) {
        OrderInfo info;
        if (int rc = probe_order(ctx, pos, n, g.dims / boxsize, &info))
            return rc;
        ctx->last_order[0] = info.coherent;
        ctx->last_order[1] = info.lattice;
        ctx->last_order[2] = info.n0;
        ctx->last_order[3] = info.n1;
        ctx->last_order[4] = info.score_z;
        ctx->last_order[5] = info.score_y;
        ctx->last_order[6] = info.score_x;
        if (!(lattice_mode && n0 > 0)) {                  // (a caller's hint stands; the probe only adds the x offsets then)
            n0 = info.lattice ? info.n0 : 0;
            n1 = info.lattice ? info.n1 : 0;
        }
        if (info.lattice && info.dx_valid && info.n0 == n0 && (info.n1 == n1 || n1 == 0)) {
            plan->have_dx = true;
            plan->dx_mean = info.dx_mean;
            plan->dx_dev = info.dx_dev;
        }
        if (mode == GENPK_DEPOSIT_AUTO)
            mode = info.lattice ? (ctx->sweep ? GENPK_DEPOSIT_SWEEP : GENPK_DEPOSIT_MARCH)
                                : ((info.coherent || fits_l2) ? GENPK_DEPOSIT_DIRECT : GENPK_DEPOSIT_SORTED);
    }
    plan->mode = mode;
    plan->n0 = n0;
    plan->n1 = n1;
    return 0;
}

__global__ void max_mass_kernel(const float *m, int64_t n, float *out)
{
    float best = 0.f;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float v = fabsf(m[i]);
        if (v < 3.0e38f && v > best) best = v;              // (NaN and inf compare false / are skipped)
    }
    for (int o = 16; o > 0; o >>= 1)
        best = fmaxf(best, __shfl_xor_sync(0xffffffffu, best, o));
    if ((threadIdx.x & 31) == 0 && best > 0.f)
        atomicMax(reinterpret_cast<int *>(out), __float_as_int(best));   // positive floats order like their bits
}

// The scale of a grid's fixed-point sums is fixed by the first deposit after genpk_grid_zero.  A given number of
// bits (GENPK_OPT_SCALE_BITS >= 0) is used as it is.  -1: S = 40 - ceil(log2(largest particle mass of this
// deposit)), so that one contribution keeps ~40 significant bits whatever the mass unit (stellar masses of 1e-8
// lose everything at S = 40) and 2^23 such particles fit in a cell.
int latch_scale(genpk_ctx *ctx, int which, const float *masses_dev, int64_t n, double mass)
{
    if (!ctx->fixed || ctx->grid_scale_latched[which])
        return 0;
    int bits = ctx->scale_bits;
    if (bits < 0) {
        double biggest = fabs(mass);
        if (masses_dev) {
            if (!ctx->d_maxmass)
                GENPK_CUDA_OK(cudaMalloc(&ctx->d_maxmass, sizeof(float)));
            GENPK_CUDA_OK(cudaMemsetAsync(ctx->d_maxmass, 0, sizeof(float), ctx->stream));
            max_mass_kernel<<<ctx->sm_count * 4, 256, 0, ctx->stream>>>(masses_dev, n, ctx->d_maxmass);
            ctx->launches++;
            float h = 0.f;
            GENPK_CUDA_OK(cudaMemcpyAsync(&h, ctx->d_maxmass, sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
            GENPK_CUDA_OK(cudaStreamSynchronize(ctx->stream));
            biggest = h;
        }
        bits = 40;
        if (biggest > 0 && biggest < 1e300)
            bits = 40 - (int)ceil(log2(biggest));
        bits = bits < 0 ? 0 : (bits > 400 ? 400 : bits);       // only the double 2^bits has to exist
    }
    ctx->grid_scale_bits[which] = bits;
    ctx->grid_scale_latched[which] = true;
    return 0;
}

int materialize_zero(genpk_ctx *ctx, int which)
{
    ctx->grid_clean[which] = false;                          // whoever asks is about to read or write the grid
    if (!ctx->zero_pending[which])
        return 0;
    ctx->zero_pending[which] = false;
    {
        StageScope scope(ctx, ST_ZERO);
        GENPK_CUDA_OK(cudaMemsetAsync(ctx->grid[which], 0, ctx->g.grid_doubles() * sizeof(double), ctx->stream));
    }
    if (ctx->g.nranks > 1)
        return touched_set(ctx, which, 0x7fffffff, -1);      // nothing touched yet
    return 0;
}

int fixed_to_double(genpk_ctx *ctx, int which)
{
    if (int rc = materialize_zero(ctx, which))
        return rc;
    if (!ctx->grid_is_fixed[which])
        return 0;
    // the owned planes only: after the ghost exchange nothing reads the ghost planes of this rank but its
    // neighbours, which may still be pulling them (as int64) while this rank goes on to its transform
    const size_t n = ctx->g.owned_doubles();
    fixed_to_double_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(ctx->grid[which] + ctx->g.owned_offset(), n,
                                                                      ldexp(1.0, -ctx->grid_scale_bits[which]));
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
        a.pos64 = nullptr;
        a.mass = d_mass;
        a.n = n;
        a.cmass = mass;
        a.units = dims / boxsize;
        a.scale = 1.0;
        a.dims = dims;
        a.fd = fd;
        a.x0 = 0;
        a.slab = 0;
        a.ghost_lo = 0;
        a.xl_max = dims - 1;
        a.plane = (size_t)dims * fd;
        a.grid = d_grid;
        a.errors = d_err;
        a.touched = nullptr;
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
