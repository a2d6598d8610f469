// Lattice-sweep CIC deposit: persistent warps, one (z segment, y block) column of the particle
// lattice each, swept along x through every lattice plane -- the default path for snapshot /
// lattice-ordered input (particle p = (ix*n1 + iy)*n0 + iz sits near lattice site (ix,iy,iz), z
// fastest: BASELINE configs 3 and 5, "Zel'dovich-displaced").  Numerics are those of
// deposit_direct_kernel (fieldize.cpp:46-114) contribution by contribution; what changes is how
// few instructions, reductions and DRAM bytes a particle costs:
//
//   * merge order y -> x -> z.  The four high-y sums of a row stay in registers and join the
//     next row's low-y sums; the two high-x sums of every row wait in a per-thread shared-memory
//     slot for the same row of the next plane; only ONE value per particle (its high-z sum)
//     crosses lanes by shuffle.  On a regular lattice one red.add per particle leaves the SM,
//     32 lanes to 32 consecutive doubles.  (deposit_march_kernel merges z first: four shuffled
//     doubles per particle, and its carries end every 8 planes.)
//   * every hand-over is validated by comparing linear cell indices, so ANY input gives the right
//     sums -- an irregular neighbour just flushes what was carried with its own red.add.
//   * particle rows are read straight into registers one step ahead (three coalesced LDG per
//     step, addresses advance by a constant): no staging buffer, no per-copy address arithmetic.
//   * first-touch zeroing ("zero ahead"): when the grid is to be zeroed first (genpk_grid_zero
//     followed by one deposit call) no memset runs.  The sweep moves through the grid as a front
//     of x planes; every warp clears its share of the plane `ahead` planes in front of the
//     expected position of its current lattice plane, completion is counted per plane, and a
//     warp deposits a lattice plane only once every grid plane up to that distance is known to
//     be clear.  The zero lines are still in L2 when the reductions arrive: no 8.6 GB memset
//     write, no 8.6 GB fill read.  A particle whose cell lies beyond the cleared front (a
//     displacement larger than the window) is not deposited by the sweep but recorded in a
//     per-warp list and deposited by deposit_deferred_kernel afterwards -- exact for any input,
//     the window is only ever a performance choice.
#include <cooperative_groups.h>

#include "deposit.cuh"

namespace genpk {

#ifndef GENPK_SWEEP_THREADS
#define GENPK_SWEEP_THREADS 128      // 4 warps: five CTAs per SM = 20 warps at 96 registers per thread (24 warps leave 80, which
                                     // spills; register files are handed out in units of 4 warps, so 7-warp CTAs do not help)
#endif
constexpr int SWEEP_THREADS = GENPK_SWEEP_THREADS;
constexpr int SWEEP_WARPS = SWEEP_THREADS / 32;
constexpr int SWEEP_RY_MAX = 21;         // x-carry slots per thread: (ry + 1) * 20 B * 128 threads <= 56 KB

#ifndef GENPK_SWEEP_MAXREG
#define GENPK_SWEEP_MAXREG 96
#endif

struct SweepArgs {
    long long n0, n1;            // lattice row length, rows per plane (z fastest)
    long long x_begin, x_end;    // lattice planes swept
    long long row_bytes;         // 12 * n0: bytes between lattice rows of the particle array
    long long plane_bytes;       // 12 * n0 * n1: bytes between lattice planes
    int rx;                      // > 0: blockIdx.y splits the sweep into blocks of rx lattice planes (independent tasks in
                                 // launch order: the front is then rx planes thick without any waiting); 0: one sweep
    int ry;                      // lattice rows per column
    int nzs;                     // 31-particle segments per row
    int ncols;                   // nzs * ceil(n1 / ry) columns, one warp each
    // ---- zero ahead (ZA) ----
    int za_periodic;             // 1: plane u of the zeroing order is grid plane (za_base + u) mod dims; 0: plane u itself (slab)
    int za_base;
    int za_umax;                 // planes of the grid (all are cleared by the launch)
    int za_upre;                 // planes [0, za_upre) were cleared before the launch
    int za_ahead;                // planes up to uc(x) + za_ahead (exclusive) must be clear before lattice plane x is deposited
    int za_slack;                // lattice planes between clearing a plane and first needing it
    long long za_uc0;            // expected plane (in u) of lattice plane x_begin
    unsigned long long za_gstep; // grid planes per lattice plane, 32.32 fixed point
    int n_zero_ctas;             // the first n_zero_ctas CTAs of the launch only clear planes (zero ahead)
    int couple_step;             // every couple_step lattice planes a warp leaves an arrival mark ...
    int couple;                  // ... and waits until every warp has left the mark `couple` marks back (0: never waits)
    int poll_weak;               // probe the marks with weak L1-bypassing loads instead of relaxed.gpu ones
    unsigned *arrived;           // [marks] warps that have left mark j (= started lattice plane j * couple_step)
    unsigned *zdone;             // [za_umax] zero CTAs that have cleared their share of plane u
    size_t zero_units;           // 16-byte units per grid plane
    unsigned *def_count;         // [SWEEP_DEF_LISTS] deferred particles per list (a column appends to list col % LISTS)
    unsigned *def_list;          // [SWEEP_DEF_LISTS][def_cap] particle indices (64 bit)
    unsigned *def_overflow;      // set when a list ran full: the clean-up rescans every particle
    int def_cap;
};

// expected plane (u coordinates) of lattice plane x: the same integer arithmetic everywhere
__host__ __device__ __forceinline__ long long sweep_uc(const SweepArgs &g, long long x)
{
    return g.za_uc0 + (long long)((((unsigned long long)(x - g.x_begin)) * g.za_gstep + 0x80000000ull) >> 32);
}
// lattice plane x may deposit into u-planes [.., utest): INT_MAX once every plane is clear
__host__ __device__ __forceinline__ int sweep_utest(const SweepArgs &g, long long x)
{
    const long long f = sweep_uc(g, x) + g.za_ahead;
    return f >= g.za_umax ? 0x7fffffff : (int)f;
}
__device__ __forceinline__ bool sweep_za_pass(const SweepArgs &g, int xl, int dims, int utest)
{
    int u = xl - g.za_base;
    if (g.za_periodic && u < 0)
        u += dims;
    return utest == 0x7fffffff || u + 2 <= utest;      // both x planes of the cloud are clear
}

template <bool FIXED> struct Acc2;
template <> struct Acc2<false> { typedef double2 type; };
template <> struct Acc2<true> { typedef longlong2 type; };

// red.global.add without a return value, spelled out: with the acquire loads and fences of the
// zero-ahead variant in the same kernel the compiler turns atomicAdd() into ATOMG, which has a
// return path.  (ptxas never predicates REDG.F64 -- `@p red.global.add.f64` comes out as a branch
// around it -- so the emission sites are few and each sits behind one branch.)
__device__ __forceinline__ void red_add(double *p, double v) { asm volatile("red.global.add.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory"); }
__device__ __forceinline__ void red_add(long long *p, long long v)
{
    asm volatile("red.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

constexpr int SWEEP_DEF_LISTS = 1024;    // deferred-particle lists (a column appends to list col % 1024)

// Spin until *counter >= want.  Relaxed loads with a back-off: what follows the wait are reductions
// performed at L2, which the loop's exit orders after the load (no L1 involved, so no acquire --
// an acquire load here costs a whole-L1 invalidate per probe).
__device__ __forceinline__ unsigned poll_load(const unsigned *counter, int weak)
{
    unsigned seen;
    // weak: a plain load that bypasses L1 (ld.global.cg).  A relaxed.gpu load is "strong": it queues behind the
    // thread's own outstanding reductions, and a probe then costs microseconds.
    if (weak)
        asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory");
    else
        asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory");
    return seen;
}
__device__ __forceinline__ void poll_at_least(const unsigned *counter, unsigned want, int weak)
{
    unsigned seen = poll_load(counter, weak);
    while (seen < want) {
        __nanosleep(64);
        seen = poll_load(counter, weak);
    }
}

// Zero ahead, the clearing side: the first n_zero_ctas CTAs of the launch do nothing but clear grid
// planes in order, each its slice of every plane, `slack` lattice planes ahead of the leading warp
// of the sweep (arrived[x] > 0 says some warp has started lattice plane x), and count themselves
// into zdone[u] once their slice of plane u is visible.
__device__ __noinline__ void sweep_zero_role(const DepositArgs &a, const SweepArgs &g)
{
    __shared__ int s_lead;
    const int tid = threadIdx.x, nthreads = blockDim.x, z = blockIdx.x;
    const int n_planes = (int)(g.x_end - g.x_begin);
    const size_t share = (g.zero_units + g.n_zero_ctas - 1) / g.n_zero_ctas;
    const size_t lo = (size_t)z * share;
    size_t hi = lo + share;
    if (hi > g.zero_units) hi = g.zero_units;
    const int step = g.couple_step;
    const int n_marks = (n_planes + step - 1) / step;
    int lead = 0;                                             // the sweep's leaders have started lattice plane lead * step
    for (int u = g.za_upre; u < g.za_umax; u++) {
        // plane u comes into reach when the leaders start a lattice plane x with uc(x + slack) + ahead > u
        if (tid == 0) {
            while (lead < n_marks - 1 && sweep_uc(g, g.x_begin + (long long)lead * step + g.za_slack) + g.za_ahead <= u) {
                poll_at_least(g.arrived + lead + 1, 1u, g.poll_weak);
                lead++;
            }
            s_lead = lead;
        }
        __syncthreads();
        lead = s_lead;
        int pl = u;
        if (!a.slab) {
            pl += g.za_base;
            if (pl >= a.dims) pl -= a.dims;
        }
        uint4 *base = reinterpret_cast<uint4 *>(reinterpret_cast<double *>(a.grid) + (size_t)pl * a.plane);
        for (size_t i = lo + tid; i < hi; i += nthreads)
            base[i] = make_uint4(0u, 0u, 0u, 0u);
        __syncthreads();
        if (tid == 0) {
            __threadfence();
            atomicAdd(&g.zdone[u], 1u);
        }
    }
}

// A particle whose cloud touches the periodic wrap of an axis (a cell index dims-1 or beyond, or a
// negative one), found by the sweep: deposited on the spot with eight reductions and kept out of
// the carries, so that every carried sum belongs to a cell whose +1 neighbours are plain +1 /
// +fd / +plane steps.  0.3 % of the particles at 1024^3.
template <bool FIXED>
__device__ __noinline__ void sweep_edge_particle(const DepositArgs &a, float px, float py, float pz, double m)
{
    deposit_single<FIXED, false>(a, px, py, pz, m);          // (the sweep tracks the touched planes in registers)
}

template <bool FIXED, typename key_t, bool FULL, bool MASS, bool ZA>
__global__ void __maxnreg__(GENPK_SWEEP_MAXREG) deposit_sweep_kernel(const __grid_constant__ DepositArgs a,
                                                                                       const __grid_constant__ SweepArgs g)
{
    typedef typename Acc<FIXED>::type acc_t;
    typedef typename Acc2<FIXED>::type acc2_t;
    constexpr key_t INVALID = ~(key_t)0;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // x-carry slots [ry+1][threads]: (high-x sums of the z0 and z1 cells) and the cell they belong to
    acc2_t *const xv0 = reinterpret_cast<acc2_t *>(smem_raw) + tid;
    key_t *const xk0 = reinterpret_cast<key_t *>(reinterpret_cast<acc2_t *>(smem_raw) + (size_t)(g.ry + 1) * SWEEP_THREADS) + tid;

    if (ZA && (int)blockIdx.x < g.n_zero_ctas) {
        sweep_zero_role(a, g);
        return;
    }
    const int col = ((int)blockIdx.x - (ZA ? g.n_zero_ctas : 0)) * SWEEP_WARPS + warp;
    if (col >= g.ncols)
        return;
    const int zseg = col % g.nzs, yb = col / g.nzs;
    const long long y0 = (long long)yb * g.ry;
    const int ry_eff = (int)((g.n1 - y0) < g.ry ? (g.n1 - y0) : g.ry);
    const long long iz = 31LL * zseg + lane;                 // lane 0 repeats the previous segment's lane 31
    const bool pair_in_row = iz + 1 < g.n0;
    // flags: 1 the lane has a particle in this row; 2 owner lane (lane 0 of later segments only carries
    // z1 sums); 4 this lane emits its z1 sums on a FULL lattice.  The z1 sums of the pair (lane, lane+1)
    // belong to this warp when lane < 31; a row's last particle has no pair and emits them in its owner
    // lane.  Otherwise (!FULL) the array may end inside a row and the rule varies per step.
    int flags = (iz < g.n0 ? 1 : 0) | ((lane > 0 || zseg == 0) ? 2 : 0);
    flags |= (pair_in_row ? lane < 31 : (flags & 2) != 0) ? 4 : 0;
    asm volatile("" : "+r"(flags));                          // keep it a register: not rematerialised from 64-bit compares
    const bool lane_in_row = flags & 1, owner_lane = flags & 2, emit_b_full = flags & 4;
    const float pos_limit = (float)(2.0e9 / a.units);        // |x| < 2e9 cells, as in axis_cell (rare path only)
    const int dims = a.dims;
    const double units = a.units;
    const key_t kplane = (key_t)a.plane, kfd = (key_t)a.fd;
    acc_t *const grid = reinterpret_cast<acc_t *>(a.grid);

    for (int s = 0; s <= ry_eff; s++)
        xk0[s * SWEEP_THREADS] = INVALID;

    // ---- particle rows: read straight into registers, SWEEP_AHEAD steps ahead of the row being deposited ----
    // Streaming loads (ld.global.cs = evict first in L1 and L2): the particles must not push grid lines out
    // of L2.  (A cp.async ring with an L2 evict-first cache hint is what deposit_march_kernel uses; here ptxas
    // 12.9 reads the hint descriptor of the prologue copies from uniform registers nobody wrote, and the GPU
    // answers with "illegal instruction".)
    long long p = (g.x_begin * g.n1 + y0) * g.n0 + iz;       // particle of the row being DEPOSITED (tracked when needed)
    long long lp_index = p;                                  // particle of the row being LOADED (tracked when !FULL)
    const long long plane_inc = (g.n1 - ry_eff + 1) * g.n0;
    // byte steps of the load pointer: a row, and what the last row of a block adds on top of it
    // (opaque to the compiler, which would otherwise redo the 64-bit products every step)
    long long block_adj_bytes = 12 * (plane_inc - g.n0);
    asm volatile("" : "+l"(block_adj_bytes));
    const char *lp = reinterpret_cast<const char *>(a.pos + 3 * p);
    const float *lm = MASS ? a.mass + p : nullptr;
    // the lattice planes of this task
    long long x_first = g.x_begin, x_last = g.x_end;
    if (g.rx > 0) {
        x_first = g.x_begin + (long long)blockIdx.y * g.rx;
        x_last = x_first + g.rx < g.x_end ? x_first + g.rx : g.x_end;
        const long long skip = (x_first - g.x_begin) * g.n1 * g.n0;
        p += skip;
        lp_index += skip;
        lp += 12 * skip;
        if (MASS) lm += skip;
    }
    const int n_planes = (int)(x_last - x_first);
    int l_r = 0;                                             // row (within the block) of the row lp points at
    // L2 prefetch: the same row one lattice plane further on (ry_eff steps ahead of the load, which itself
    // runs one row ahead of the deposit); no cursor of its own
    float ax = 0.f, ay = 0.f, az = 0.f, am = 0.f;             // the row in flight
    auto issue_load = [&](bool prefetch_ok) {
        if (lane_in_row && (FULL || lp_index < a.n)) {
            if (FULL && prefetch_ok)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(lp + g.plane_bytes));
            ax = __ldcs(reinterpret_cast<const float *>(lp));
            ay = __ldcs(reinterpret_cast<const float *>(lp) + 1);
            az = __ldcs(reinterpret_cast<const float *>(lp) + 2);
            if (MASS)
                am = __ldcs(lm);
        }
        const bool last_row = ++l_r == ry_eff;
        lp += g.row_bytes;
        if (last_row) {
            lp += block_adj_bytes;
            l_r = 0;
        }
        if (!FULL || MASS) {
            const long long inc = last_row ? plane_inc : g.n0;
            lp_index += inc;
            if (MASS) lm += inc;
        }
    };
    if (n_planes > 0)
        issue_load(x_first + 2 < g.x_end);

    // One row slot: merge what the previous plane left for this row (x), leave this row's high-x
    // sums for the next plane, then hand the z1 sum to the next lane (z) and emit.
    //   v00, v01: sums of the cells `cell` and `cell + 1`;  v10, v11: of `cell + plane` (+ 1)
    // (carried sums only ever belong to cells whose +1 neighbours are plain steps, see above)
    auto row_slot = [&](acc2_t *xv, key_t *xk, bool has, bool emit_b, key_t cell, acc_t v00, acc_t v10, acc_t v01, acc_t v11) {
        const key_t old = *xk;
        const acc2_t s = *xv;
        const bool merged = has && old == cell;
        if (merged) {
            v00 += s.x;
            v01 += s.y;
        }
        // a carry nobody took: out with its own reductions (sums this lane must not emit are zero)
        if (old != INVALID && !merged) {
            red_add(grid + (size_t)old, s.x);
            red_add(grid + (size_t)(key_t)(old + 1), s.y);
        }
        if (has) {
            acc2_t t;
            t.x = v10;
            t.y = v11;
            *xv = t;
        }
        *xk = has ? (key_t)(cell + kplane) : INVALID;
        const key_t key_b = (key_t)(cell + 1);
        const key_t next_cell = __shfl_down_sync(0xffffffffu, has ? cell : INVALID, 1);
        const bool give = lane < 31 && has && next_cell == key_b;
        const int gave = __shfl_up_sync(0xffffffffu, (int)give, 1);
        const acc_t from_prev = __shfl_up_sync(0xffffffffu, v01, 1);
        if (lane > 0 && gave)
            v00 += from_prev;
        if (has && owner_lane)
            red_add(grid + (size_t)cell, v00);
        if (has && !give && emit_b)
            red_add(grid + (size_t)key_b, v01);
    };

    // ---- zero ahead / coupling: planes known to be clear by this warp ----
    int ready = g.za_upre;
    int to_mark = 1, mark = 0;                               // planes until the next arrival mark / its index
    unsigned long long gacc = 0x80000000ull;                 // (x - x_begin) * gstep + 1/2

    acc_t c00 = 0, c10 = 0, c01 = 0, c11 = 0;                // y-carry: the four high-y sums of the previous row
    key_t yc_key = INVALID;
    unsigned n_rejected = 0;
    int t_lo = 0x7fffffff, t_hi = -1;                        // slab: lowest / highest low-x plane of a deposited cloud

    for (int xi = 0; xi < n_planes; xi++) {
        // zero ahead: this lattice plane may deposit into local planes xl with (xl - za_lo) mod dims <= za_span
        int za_lo = 0, za_span = 0x7fffffff;
        // Coupling: every warp announces the lattice plane it starts and starts plane x only when all
        // warps have started plane x - couple, so the front of the sweep stays `couple` planes thick
        // (L2 locality of the reductions; the zero CTAs clear just ahead of the leaders).
        if (g.arrived && --to_mark == 0) {
            to_mark = g.couple_step;
            if (lane == 0) {
                asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(g.arrived + mark), "r"(1u) : "memory");
                if (g.couple > 0 && mark >= g.couple)
                    poll_at_least(g.arrived + (mark - g.couple), (unsigned)g.ncols, g.poll_weak);
            }
            mark++;
            __syncwarp();
        }
        if (ZA) {
            const long long front = g.za_uc0 + (long long)(gacc >> 32) + g.za_ahead;
            const int need = front > g.za_umax ? g.za_umax : (int)front;
            za_lo = g.za_base;
            za_span = front >= g.za_umax ? 0x7fffffff : (int)front - 2;          // u + 2 <= front
            if (ready < need) {
                if (lane == 0)
                    for (int u = ready; u < need; u++)
                        poll_at_least(g.zdone + u, (unsigned)g.n_zero_ctas, g.poll_weak);
                ready = need;
                __syncwarp();
            }
            gacc += g.za_gstep;
        }
        acc2_t *xv = xv0;
        key_t *xk = xk0;
        const bool last_plane = xi + 1 == n_planes;
        const bool pf_ok = x_first + xi + 3 < g.x_end;        // (the load cursor may already be one plane on)
        for (int r = 0; r < ry_eff; r++) {
            const float px = ax, py = ay, pz = az;
            double m = a.cmass;
            if (MASS)
                m = (double)am;                                             // fieldize.cpp:63
            if (!(last_plane && r + 1 == ry_eff))
                issue_load(pf_ok);                                          // the next step's row
            const bool live = lane_in_row && (FULL || p < a.n);
            const bool emit_b = FULL ? emit_b_full : ((pair_in_row && p + 1 < a.n) ? lane < 31 : owner_lane);
            const long long p_now = p;
            if (!FULL || MASS)
                p += (r + 1 == ry_eff) ? plane_inc : g.n0;
            int fx, fy, fz, hx, hy, hz;
            double tx, dx, ty, dy, tz, dz;
            axis_fast_hi(px, units, fx, hx, tx, dx);
            axis_fast_hi(py, units, fy, hy, ty, dy);
            axis_fast_hi(pz, units, fz, hz, tz, dz);
            // x + 1.5*2^52 has the high word 0x43380000 exactly when 0 <= x < 2^32 (and x is finite): one test
            // for "finite, non-negative, floor exact" on all three axes
            const bool in_range = (((hx ^ 0x43380000) | (hy ^ 0x43380000) | (hz ^ 0x43380000)) == 0);
            int xl = fx;
            if (a.slab)                                                     // slab: the +1 neighbour may be a ghost plane
                xl = slab_plane(fx, a.x0, a.ghost_lo, dims);
            // the regular case: every cell index in [0, dims-2] (slab: local plane in [0, xl_max]), so the
            // cloud's +1 neighbours are plain steps
            const unsigned top = (unsigned)(dims - 1);
            const bool inside = ((unsigned)xl < (a.slab ? (unsigned)(a.xl_max + 1) : top)) & ((unsigned)fy < top) & ((unsigned)fz < top);
            bool pass = true;
            if (ZA) {
                // both x planes of the cloud must already be clear; else the clean-up pass deposits the particle
                const unsigned u1 = (unsigned)(xl - za_lo), u2 = u1 + (unsigned)dims;
                pass = (u1 <= (unsigned)za_span) | (!a.slab & (u2 <= (unsigned)za_span));
            }
            bool ok = in_range && inside && pass;
            if (live && !ok) {
                // rare: rejected (non-finite, outside the slab), at the periodic wrap, or beyond the cleared front
                if (owner_lane) {
                    bool deposit_now = false;
                    int wl_edge = 0;
                    ok = fabsf(px) < pos_limit && fabsf(py) < pos_limit && fabsf(pz) < pos_limit;   // as axis_cell
                    if (!ok) {
                        n_rejected++;
                    } else {
                        // wrap every axis as fieldize.cpp:70-75 does, then look again
                        const int wx = (unsigned)fx >= (unsigned)dims ? wrap_cell(fx, dims) : fx;
                        int wl = wx;
                        bool in_slab = true;
                        if (a.slab) {
                            wl = slab_plane(wx, a.x0, a.ghost_lo, dims);
                            in_slab = wl >= 0 && wl <= a.xl_max;
                        }
                        if (!in_slab) {
                            n_rejected++;
                        } else {
                            bool wpass = true;
                            if (ZA) {
                                const unsigned u1 = (unsigned)(wl - za_lo), u2 = u1 + (unsigned)dims;
                                wpass = (u1 <= (unsigned)za_span) | (!a.slab & (u2 <= (unsigned)za_span));
                            }
                            deposit_now = wpass;
                            wl_edge = wl;
                            if (ZA && !wpass) {
                                const int li = col & (SWEEP_DEF_LISTS - 1);
                                const unsigned slot = atomicAdd(&g.def_count[li], 1u);
                                // (p is only tracked when the step needs it)
                                const long long q = (!FULL || MASS) ? p_now : ((x_first + xi) * g.n1 + y0 + r) * g.n0 + 31LL * zseg + lane;
                                if (slot < (unsigned)g.def_cap)
                                    reinterpret_cast<unsigned long long *>(g.def_list)[(size_t)li * g.def_cap + slot] = (unsigned long long)q;
                                else
                                    *g.def_overflow = 1u;
                            }
                        }
                    }
                    if (deposit_now) {
                        sweep_edge_particle<FIXED>(a, px, py, pz, m);
                        if (a.slab) {
                            t_lo = wl_edge < t_lo ? wl_edge : t_lo;
                            t_hi = wl_edge > t_hi ? wl_edge : t_hi;
                        }
                    }
                }
                ok = false;
            }
            ok = ok && live;
            if (a.slab && ok) {                                             // planes this rank's ghost exchange has to move
                t_lo = xl < t_lo ? xl : t_lo;
                t_hi = xl > t_hi ? xl : t_hi;
            }
            // sums this lane must not emit are zero from the start: lane 0 of a later segment repeats the
            // previous segment's last particle and only carries its z1 sums; the z1 sums of a lane whose
            // pair belongs to the next segment are carried there.  Every carry can then be flushed as it is.
            if (!owner_lane)
                tz = 0.0;
            if (!emit_b)
                dz = 0.0;
            const double mx0 = __dmul_rn(m, tx), mx1 = __dmul_rn(m, dx);
            // weights in the order of fieldize.cpp:77-84; first the four low-y cells ...
            const double w00 = __dmul_rn(mx0, ty), w10 = __dmul_rn(mx1, ty);
            acc_t v00 = Acc<FIXED>::make(__dmul_rn(w00, tz), a.scale);        // (x0, y0, z0)
            acc_t v10 = Acc<FIXED>::make(__dmul_rn(w10, tz), a.scale);        // (x1, y0, z0)
            acc_t v01 = Acc<FIXED>::make(__dmul_rn(w00, dz), a.scale);        // (x0, y0, z1)
            acc_t v11 = Acc<FIXED>::make(__dmul_rn(w10, dz), a.scale);        // (x1, y0, z1)
            const key_t cell = ((key_t)xl * (key_t)dims + (key_t)fy) * kfd + (key_t)fz;
            // ---- y: the previous row's high-y sums are this row's low-y sums ----
            {
                const bool merged = ok && yc_key == cell;
                if (merged) {
                    v00 += c00;
                    v10 += c10;
                    v01 += c01;
                    v11 += c11;
                }
                if (yc_key != INVALID && !merged) {                         // a carry nobody took
                    red_add(grid + (size_t)yc_key, c00);
                    red_add(grid + (size_t)(key_t)(yc_key + kplane), c10);
                    red_add(grid + (size_t)(key_t)(yc_key + 1), c01);
                    red_add(grid + (size_t)(key_t)(yc_key + kplane + 1), c11);
                }
            }
            // ... then the four high-y cells, which wait for the next row
            const double w01 = __dmul_rn(mx0, dy), w11 = __dmul_rn(mx1, dy);
            c00 = Acc<FIXED>::make(__dmul_rn(w01, tz), a.scale);              // (x0, y1, z0)
            c10 = Acc<FIXED>::make(__dmul_rn(w11, tz), a.scale);              // (x1, y1, z0)
            c01 = Acc<FIXED>::make(__dmul_rn(w01, dz), a.scale);              // (x0, y1, z1)
            c11 = Acc<FIXED>::make(__dmul_rn(w11, dz), a.scale);              // (x1, y1, z1)
            yc_key = ok ? (key_t)(cell + kfd) : INVALID;

            // ---- x (slot r of the plane carry), then z ----
            row_slot(xv, xk, ok, emit_b, cell, v00, v10, v01, v11);
            xv += SWEEP_THREADS;
            xk += SWEEP_THREADS;
        }
        // the block's last row: its high-y sums leave through slot ry_eff (cell = that row's (x0, y1, z0))
        row_slot(xv, xk, yc_key != INVALID, FULL ? emit_b_full : true, yc_key, c00, c10, c01, c11);
        yc_key = INVALID;
    }
    // ---- what the last plane left behind ----
    for (int s = 0; s <= ry_eff; s++) {
        const key_t old = xk0[s * SWEEP_THREADS];
        const acc2_t t = xv0[s * SWEEP_THREADS];
        if (old != INVALID) {
            red_add(grid + (size_t)old, t.x);
            red_add(grid + (size_t)(key_t)(old + 1), t.y);
        }
    }
    if (n_rejected)
        atomicAdd(a.errors, (unsigned long long)n_rejected);
    if (a.touched) {
        t_lo = __reduce_min_sync(0xffffffffu, t_lo);
        t_hi = __reduce_max_sync(0xffffffffu, t_hi);
        if (lane == 0 && t_hi >= 0) {
            atomicMin(a.touched, t_lo);
            atomicMax(a.touched + 1, t_hi + 1);
        }
    }
}

// ---------------------------------------------------------------------------------
// The particles a zero-ahead sweep left out (their cells lay beyond the cleared front when
// their lattice plane passed): deposited here, after the sweep, with eight reductions each.
// One warp per column list; when a list overflowed, every particle is tested again instead.
// ---------------------------------------------------------------------------------
template <bool FIXED>
__global__ void __launch_bounds__(256) deposit_deferred_kernel(const __grid_constant__ DepositArgs a,
                                                               const __grid_constant__ SweepArgs g)
{
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long n_threads = (long long)gridDim.x * blockDim.x;
    if (*g.def_overflow == 0u) {
        const unsigned long long *lists = reinterpret_cast<const unsigned long long *>(g.def_list);
        // thread t takes entry t / LISTS of list t % LISTS: consecutive threads, different lists
        for (long long t = tid;; t += n_threads) {
            const int li = (int)(t & (SWEEP_DEF_LISTS - 1));
            const long long e = t / SWEEP_DEF_LISTS;
            if (e >= g.def_cap)
                break;
            if (e < (long long)g.def_count[li]) {
                const long long p = (long long)lists[(size_t)li * g.def_cap + e];
                deposit_single<FIXED>(a, a.pos[3 * p], a.pos[3 * p + 1], a.pos[3 * p + 2], a.mass ? (double)a.mass[p] : a.cmass);
            }
        }
        return;
    }
    // a list ran full: test every particle of the swept range the way the sweep did
    const long long first = g.x_begin * g.n1 * g.n0;
    long long last = g.x_end * g.n1 * g.n0;
    if (last > a.n) last = a.n;
    const long long per_plane = g.n0 * g.n1;
    const float pos_limit = (float)(2.0e9 / a.units);
    for (long long p = first + tid; p < last; p += n_threads) {
        const float px = a.pos[3 * p], py = a.pos[3 * p + 1], pz = a.pos[3 * p + 2];
        if (!(fabsf(px) < pos_limit && fabsf(py) < pos_limit && fabsf(pz) < pos_limit))
            continue;                                         // counted as rejected by the sweep
        const AxisCell cx = axis_cell(px, a.units, a.dims);
        int xl = cx.lo;
        if (a.slab) {
            xl = slab_plane(cx.lo, a.x0, a.ghost_lo, a.dims);
            if (xl < 0 || xl > a.xl_max)
                continue;                                     // rejected by the sweep
        }
        if (!sweep_za_pass(g, xl, a.dims, sweep_utest(g, p / per_plane)))
            deposit_single<FIXED>(a, px, py, pz, a.mass ? (double)a.mass[p] : a.cmass);
    }
}

// ---------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------
template <bool FIXED, typename key_t, bool FULL, bool MASS, bool ZA> static const void *sweep_fn()
{
    return reinterpret_cast<const void *>(&deposit_sweep_kernel<FIXED, key_t, FULL, MASS, ZA>);
}

template <bool FIXED, typename key_t> static const void *sweep_pick(bool full, bool mass, bool za)
{
    const int sel = (full ? 4 : 0) | (mass ? 2 : 0) | (za ? 1 : 0);
    switch (sel) {
    case 0: return sweep_fn<FIXED, key_t, false, false, false>();
    case 1: return sweep_fn<FIXED, key_t, false, false, true>();
    case 2: return sweep_fn<FIXED, key_t, false, true, false>();
    case 3: return sweep_fn<FIXED, key_t, false, true, true>();
    case 4: return sweep_fn<FIXED, key_t, true, false, false>();
    case 5: return sweep_fn<FIXED, key_t, true, false, true>();
    case 6: return sweep_fn<FIXED, key_t, true, true, false>();
    default: return sweep_fn<FIXED, key_t, true, true, true>();
    }
}

static size_t sweep_smem(int ry, bool fixed, bool key32)
{
    (void)fixed;
    return (size_t)(ry + 1) * SWEEP_THREADS * (16 + (key32 ? 4 : 8));
}

// device scratch of the coupled / zero-ahead sweep: arrived[n_planes], zdone[umax], overflow flag;
// deferred-particle counters and lists
static int ensure_za_scratch(genpk_ctx *ctx, long long n_planes, int umax, int def_cap)
{
    const size_t words = (size_t)n_planes + (size_t)umax + 8;
    if (words > (size_t)ctx->za_zdone_cap) {
        if (ctx->d_za_zdone) cudaFree(ctx->d_za_zdone);
        ctx->d_za_zdone = nullptr;
        ctx->za_zdone_cap = 0;
        GENPK_CUDA_OK(cudaMalloc(&ctx->d_za_zdone, words * sizeof(unsigned)));
        ctx->za_zdone_cap = (int)words;
    }
    const size_t need = (size_t)SWEEP_DEF_LISTS * (2 * (size_t)def_cap + 2);        // counters, then 64-bit entries
    if (need > ctx->za_def_cap) {
        if (ctx->d_za_def) cudaFree(ctx->d_za_def);
        ctx->d_za_def = nullptr;
        ctx->za_def_cap = 0;
        GENPK_CUDA_OK(cudaMalloc(&ctx->d_za_def, need * sizeof(unsigned)));
        ctx->za_def_cap = need;
    }
    return 0;
}

// n0: particles per lattice row; n1: rows per plane (0 = unknown: the whole array is one plane).
int launch_sweep(genpk_ctx *ctx, const DepositArgs &a, long long n0, long long n1, bool za, const OrderInfo *info,
                 bool *za_possible)
{
    if (za_possible) *za_possible = true;
    if (a.n <= 0)
        return 0;
    SweepArgs g = {};
    if (n0 < 1 || n0 > a.n) n0 = a.n;
    const long long rows = (a.n + n0 - 1) / n0;
    if (n1 < 1 || n1 > rows) n1 = rows;
    const long long n2 = (rows + n1 - 1) / n1;
    g.n0 = n0;
    g.n1 = n1;
    g.row_bytes = 12 * n0;
    g.plane_bytes = 12 * n0 * n1;
    g.x_begin = 0;
    g.x_end = n2;
    g.nzs = n0 >= 2 ? (int)((n0 - 2) / 31 + 1) : 1;
    if (n2 >= (1 << 21) || (long long)g.nzs * n1 > 0x3fffffffLL) {
        set_error("deposit: lattice %lld x %lld x %lld exceeds the sweep kernel's index range", n0, n1, n2);
        return 1;
    }
    const bool full = n0 * n1 * n2 == a.n;
    const bool key32 = ctx->g.grid_doubles() < 0xfffffff0ull;
    const bool mass = a.mass != nullptr;

    // rows per column: as few as keep every column's warp resident at once (the sweep is then one
    // wave of persistent warps, which zero ahead needs and which keeps the front one plane thick)
    auto pick = [&](bool with_za) {
        return ctx->fixed ? (key32 ? sweep_pick<true, uint32_t>(full, mass, with_za) : sweep_pick<true, unsigned long long>(full, mass, with_za))
                          : (key32 ? sweep_pick<false, uint32_t>(full, mass, with_za) : sweep_pick<false, unsigned long long>(full, mass, with_za));
    };
    const void *kern = pick(za);
    int ry = 0;
    const int ry_cap = (int)(n1 < SWEEP_RY_MAX ? n1 : SWEEP_RY_MAX);
    // ---- task mode: blocks of rx lattice planes x columns, independent CTAs in launch order ----
    if (ctx->sweep_rx > 0 && !za) {
        ry = ctx->sweep_ry > 0 ? ctx->sweep_ry : 8;
        if (ry > ry_cap) ry = ry_cap;
        g.ry = ry;
        g.rx = (int)(n2 < ctx->sweep_rx ? n2 : ctx->sweep_rx);
        g.ncols = (int)(g.nzs * ((n1 + ry - 1) / ry));
        const size_t smem = sweep_smem(ry, ctx->fixed, key32);
        GENPK_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const long long bx = ((long long)g.ncols + SWEEP_WARPS - 1) / SWEEP_WARPS, by = (n2 + g.rx - 1) / g.rx;
        if (bx > 0x7fffffffLL || by > 65535) {
            set_error("deposit: %lld x %lld sweep tasks exceed the launch grid", bx, by);
            return 1;
        }
        DepositArgs args = a;
        ctx->last_sweep[0] = ry;
        ctx->last_sweep[1] = g.ncols;
        ctx->last_sweep[2] = 0;
        ctx->last_sweep[3] = 0;
        void *params[] = {(void *)&args, (void *)&g};
        GENPK_CUDA_OK(cudaLaunchKernel(kern, dim3((unsigned)bx, (unsigned)by), dim3(SWEEP_THREADS), params, smem, ctx->stream));
        ctx->launches++;
        return 0;
    }
    // zero ahead: a few CTAs of the launch do nothing but clear planes ahead of the sweep
    const int n_zero = za ? (ctx->za_zero_ctas > 0 ? ctx->za_zero_ctas : ctx->sm_count / 3) : 0;
    // (at least 4 rows when the plane has them: every block of rows pays one extra slot for its last carry)
    const int ry_lo = ctx->sweep_ry > 0 ? (ctx->sweep_ry < ry_cap ? ctx->sweep_ry : ry_cap) : (ry_cap < 4 ? ry_cap : 4);
    for (int t = ry_lo; t <= ry_cap; t++) {
        const size_t smem = sweep_smem(t, ctx->fixed, key32);
        if (smem > (size_t)ctx->smem_optin)
            break;
        GENPK_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int per_sm = 0;
        GENPK_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, SWEEP_THREADS, smem));
        const long long cols = (long long)g.nzs * ((n1 + t - 1) / t);
        const long long ctas = (cols + SWEEP_WARPS - 1) / SWEEP_WARPS + n_zero;
        if (per_sm >= 1 && ctas <= (long long)per_sm * ctx->sm_count) {
            ry = t;
            break;
        }
    }
    // one resident wave: the warps can wait for each other, so the sweep is coupled (a front a few planes
    // thick) and may clear the grid ahead of itself; more columns than resident warps: several
    // uncoupled waves over a grid that was cleared before
    const bool one_wave = ry > 0;
    if (!one_wave) {
        if (za) {
            if (za_possible) *za_possible = false;
            return 0;
        }
        ry = ctx->sweep_ry > 0 ? ctx->sweep_ry : 10;
        if (ry > ry_cap) ry = ry_cap;
        while (ry > 1 && sweep_smem(ry, ctx->fixed, key32) * 5 > (size_t)ctx->smem_optin)     // five CTAs per SM
            ry--;
    }
    g.ry = ry;
    const long long nyb = (n1 + ry - 1) / ry;
    g.ncols = (int)(g.nzs * nyb);
    g.n_zero_ctas = n_zero;
    g.couple = one_wave ? ctx->sweep_couple : 0;
    g.couple_step = ctx->sweep_couple_step > 0 ? ctx->sweep_couple_step : 1;
    g.poll_weak = ctx->sweep_poll_weak;
    const size_t smem = sweep_smem(ry, ctx->fixed, key32);
    GENPK_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long blocks = ((long long)g.ncols + SWEEP_WARPS - 1) / SWEEP_WARPS + n_zero;

    DepositArgs args = a;
    ctx->last_sweep[0] = ry;
    ctx->last_sweep[1] = g.ncols;
    ctx->last_sweep[2] = za ? 1 : 0;
    ctx->last_sweep[3] = 0;
    const SlabGeom &sg = ctx->g;
    const int n_grid_planes = sg.ghost_lo + sg.nx + sg.ghost_hi;
    if (one_wave && (za || g.couple > 0)) {
        if (int rc = ensure_za_scratch(ctx, n2, n_grid_planes, ctx->za_def_per_col)) return rc;
        g.arrived = ctx->d_za_zdone;
        g.zdone = ctx->d_za_zdone + n2;
        g.def_overflow = g.zdone + n_grid_planes;
        GENPK_CUDA_OK(cudaMemsetAsync(ctx->d_za_zdone, 0, ((size_t)n2 + n_grid_planes + 8) * sizeof(unsigned), ctx->stream));
    }
    if (!za) {
        void *params[] = {(void *)&args, (void *)&g};
        if (g.arrived)
            GENPK_CUDA_OK(cudaLaunchCooperativeKernel(kern, dim3((unsigned)blocks), dim3(SWEEP_THREADS), params, smem, ctx->stream));
        else
            GENPK_CUDA_OK(cudaLaunchKernel(kern, dim3((unsigned)blocks), dim3(SWEEP_THREADS), params, smem, ctx->stream));
        ctx->launches++;
        return 0;
    }

    // ---- zero ahead ----
    const int n_planes = n_grid_planes;
    // window: planes past the expected one that are kept clear -- a little more than the largest
    // displacement the order probe saw (particles beyond it take the deferred path)
    int window = ctx->za_window > 0 ? ctx->za_window : (info && info->dx_valid ? info->dx_dev * 5 / 4 + 1 : 6);
    if (window < 1) window = 1;
    if (window > n_planes / 2 - 2) window = n_planes / 2 - 2 > 0 ? n_planes / 2 - 2 : 0;
    const int ahead = window + 2;
    ctx->last_sweep[3] = window;
    g.za_periodic = a.slab ? 0 : 1;
    g.za_umax = n_planes;
    g.za_ahead = ahead;
    g.za_slack = ctx->za_slack > g.couple_step ? ctx->za_slack : g.couple_step;   // the leaders are known to one mark
    // grid planes per lattice plane of the GLOBAL lattice (a slab rank holds 1/nranks of its planes)
    const double planes_per = (double)sg.nx / (double)n2;
    g.za_gstep = (unsigned long long)llround(planes_per * 4294967296.0);
    // expected plane of the first lattice plane: measured by the order probe, else the lattice site itself
    long long centre = (long long)floor(0.5 * planes_per) + (info && info->dx_valid ? info->dx_mean : sg.ghost_lo);
    if (g.za_periodic) {
        long long base = centre - (ahead - 2);                  // plane u = 0: `window` planes behind the first expected plane
        base %= sg.dims;
        if (base < 0) base += sg.dims;
        g.za_base = (int)base;
        g.za_uc0 = ahead - 2;
    } else {
        g.za_base = 0;
        g.za_uc0 = centre;
    }
    long long upre = sweep_uc(g, g.x_begin) + g.za_ahead;          // what the first lattice plane needs
    if (upre < 0) upre = 0;
    if (upre > g.za_umax) upre = g.za_umax;
    g.za_upre = (int)upre;
    g.zero_units = a.plane * sizeof(double) / 16;
    g.def_count = ctx->d_za_def;
    g.def_list = ctx->d_za_def + 2 * SWEEP_DEF_LISTS;                         // 8-byte aligned
    g.def_cap = ctx->za_def_per_col;
    GENPK_CUDA_OK(cudaMemsetAsync(ctx->d_za_def, 0, SWEEP_DEF_LISTS * sizeof(unsigned), ctx->stream));
    // planes the first lattice planes need at once: cleared here (up to two pieces of the periodic grid)
    {
        const size_t plane_bytes = a.plane * sizeof(double);
        char *gb = reinterpret_cast<char *>(a.grid);
        if (g.za_periodic) {
            const int first = g.za_base, cnt = g.za_upre;
            const int c1 = first + cnt <= sg.dims ? cnt : sg.dims - first;
            if (c1 > 0) GENPK_CUDA_OK(cudaMemsetAsync(gb + (size_t)first * plane_bytes, 0, (size_t)c1 * plane_bytes, ctx->stream));
            if (cnt - c1 > 0) GENPK_CUDA_OK(cudaMemsetAsync(gb, 0, (size_t)(cnt - c1) * plane_bytes, ctx->stream));
        } else if (g.za_upre > 0) {
            GENPK_CUDA_OK(cudaMemsetAsync(gb, 0, (size_t)g.za_upre * plane_bytes, ctx->stream));
        }
    }
    void *params[] = {(void *)&args, (void *)&g};
    GENPK_CUDA_OK(cudaLaunchCooperativeKernel(kern, dim3((unsigned)blocks), dim3(SWEEP_THREADS), params, smem, ctx->stream));
    ctx->launches++;
    const int cleanup_blocks = ctx->sm_count * 8;
    if (ctx->fixed)
        deposit_deferred_kernel<true><<<cleanup_blocks, 256, 0, ctx->stream>>>(args, g);
    else
        deposit_deferred_kernel<false><<<cleanup_blocks, 256, 0, ctx->stream>>>(args, g);
    ctx->launches++;
    GENPK_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace genpk
