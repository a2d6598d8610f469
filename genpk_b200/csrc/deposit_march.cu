// Lattice-march CIC deposit: the fast path for snapshot / lattice-ordered input
// (particle p = (ix*n1 + iy)*n0 + iz sits near lattice site (ix,iy,iz), z fastest),
// which is what initial-condition and ID-ordered snapshots look like and what
// BASELINE configs 3 and 5 name ("Zel'dovich-displaced").  Same numerics as
// deposit_direct_kernel (fieldize.cpp:46-114); only the number of reductions that
// reach L2 changes.
//
// The L2 atomic units, not HBM, bound the direct kernel (ncu: lts throughput 78 %,
// ~1.6 red sectors per particle, profiles/r01/s2_ncu_full_c3_v1_summary.txt): L2
// retires ~1.7e11 red sectors/s whatever the number of lanes in a sector.  So the
// eight corner contributions of a particle are merged with its lattice neighbours'
// in registers before anything is sent:
//
//   z: a warp holds 32 consecutive particles of one lattice row; a lane's four
//      high-z corners are handed to the next lane by shuffle (its low-z corners).
//      Warps overlap by one particle (stride 31) so the hand-over never crosses a
//      warp boundary.
//   y: the warp then marches over `ry` consecutive rows; the two high-y sums stay in
//      registers and join the next row's low-y sums.
//   x: it repeats that for `rx` consecutive lattice planes; the high-x sum of every
//      row waits in a per-thread shared-memory slot for the same row of the next plane.
//
// Every hand-over is validated per lane by comparing linear cell indices, so ANY
// input gives the right sums -- an irregular lane (neighbour not exactly one cell
// further) just flushes what it carried with its own red.add.  On a regular lattice
// one red.add per particle leaves the SM, 32 lanes to 32 consecutive doubles.
//
// Particle rows are staged through shared memory with cp.async (4 rows in flight per
// warp, evict-first in L2) so the loads are fully coalesced and asynchronous.
#include "deposit.cuh"

namespace genpk {

constexpr int MARCH_THREADS = 256;
constexpr int MARCH_WARPS = MARCH_THREADS / 32;
constexpr int MARCH_STAGES = 4;          // particle rows in flight per warp
constexpr int MARCH_ROW_FLOATS = 128;    // 96 position floats + 32 masses per staged row
#ifndef GENPK_MARCH_MINB
#define GENPK_MARCH_MINB 3               // resident CTAs per SM the register allocation must allow (3 -> 80 registers)
#endif

struct MarchGeom {
    long long n0, n1, n2;        // lattice extents, z fastest: p = (ix*n1 + iy)*n0 + iz
    int ry, rx;                  // rows / planes marched by one warp
    int nzs;                     // 31-particle segments per row
    int band_rows;               // rows per y band (multiple of ry): bounds the L2 footprint of a layer
    int yb_per_band;
    long long nbands, nxb;
    long long tasks_per_layer;   // nzs * yb_per_band
    long long n_tasks;           // nbands * nxb * tasks_per_layer
};

__device__ __forceinline__ void cp_async4(void *smem_dst, const void *gsrc, unsigned long long policy)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global.L2::cache_hint [%0], [%1], 4, %2;\n" ::"r"(d), "l"(gsrc), "l"(policy)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

// FULL: every lattice site the launch touches holds a particle (n == n0*n1*n2), so no
// per-step bound checks on the particle index.  MASS: per-particle masses.
template <bool FIXED, typename key_t, bool FULL, bool MASS>
__global__ void __launch_bounds__(MARCH_THREADS, GENPK_MARCH_MINB) deposit_march_kernel(const __grid_constant__ DepositArgs a,
                                                                         const __grid_constant__ MarchGeom g)
{
    typedef typename Acc<FIXED>::type acc_t;
    constexpr key_t INVALID = ~(key_t)0;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // x-carry slots [ry+1][threads] (value, key), then the staging rows of each warp
    acc_t *const xc_val0 = reinterpret_cast<acc_t *>(smem_raw) + tid;
    key_t *const xc_key0 = reinterpret_cast<key_t *>(reinterpret_cast<acc_t *>(smem_raw) + (size_t)(g.ry + 1) * MARCH_THREADS) + tid;
    float *const stage = reinterpret_cast<float *>(reinterpret_cast<key_t *>(reinterpret_cast<acc_t *>(smem_raw) +
                                                                             (size_t)(g.ry + 1) * MARCH_THREADS) +
                                                   (size_t)(g.ry + 1) * MARCH_THREADS) +
                         (size_t)warp * MARCH_STAGES * MARCH_ROW_FLOATS;

    // ---- which part of the lattice this warp marches over ----
    const long long task = (long long)blockIdx.x * MARCH_WARPS + warp;
    if (task >= g.n_tasks)
        return;
    const long long layer = task / g.tasks_per_layer;
    const int t_in = (int)(task - layer * g.tasks_per_layer);
    const int zs = t_in % g.nzs, ybl = t_in / g.nzs;
    const long long band = layer / g.nxb, xb = layer - band * g.nxb;
    const long long y0 = band * g.band_rows + (long long)ybl * g.ry;
    if (y0 >= g.n1)
        return;
    long long y_end = y0 + g.ry;
    if (y_end > (band + 1) * g.band_rows) y_end = (band + 1) * g.band_rows;
    if (y_end > g.n1) y_end = g.n1;
    const int ry_eff = (int)(y_end - y0);
    const long long x0 = xb * g.rx;
    const long long x_end = (x0 + g.rx < g.n2) ? x0 + g.rx : g.n2;
    const int nsteps = (int)(x_end - x0) * ry_eff;
    const long long iz = 31LL * zs + lane;                   // lane 0 repeats the previous segment's lane 31
    const bool lane_in_row = iz < g.n0;
    const bool pair_in_row = iz + 1 < g.n0;
    const bool owner_lane = lane > 0 || zs == 0;             // lane 0 of later segments only hands its high-z corners on
    // particle of lane 0 at step (plane x0, row y0); +n0 per row, +plane_inc at the end of a plane
    const long long p_first = (x0 * g.n1 + y0) * g.n0 + 31LL * zs;
    const long long plane_inc = (g.n1 - ry_eff + 1) * g.n0;
    const float pos_limit = (float)(2.0e9 / a.units);        // |x| < 2e9 cells, as in axis_cell
    const key_t kplane = (key_t)a.plane, kfd = (key_t)a.fd;
    acc_t *const grid = reinterpret_cast<acc_t *>(a.grid);
    const int dims = a.dims;
    const double units = a.units;

    unsigned long long policy;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;\n" : "=l"(policy));

    for (int s = 0; s <= ry_eff; s++)
        xc_key0[s * MARCH_THREADS] = INVALID;

    // ---- cp.async pipeline over the (plane, row) steps ----
    long long pl = p_first;                  // load cursor: particle of lane 0
    const float *lsrc = a.pos + 3 * p_first + lane;
    const float *lmass = MASS ? a.mass + p_first + lane : nullptr;
    int l_r = 0;
    // FULL still has to stop the last row's 32-lane window at the end of the arrays
    const bool window_safe = FULL && ((x_end - 1) * g.n1 + y_end - 1) * g.n0 + 31LL * zs + 32 <= a.n;
    auto issue_load = [&](int q) {
        if (q < nsteps) {
            float *dst = stage + (q & (MARCH_STAGES - 1)) * MARCH_ROW_FLOATS + lane;
            if (window_safe || pl + 32 <= a.n) {
                cp_async4(dst, lsrc, policy);
                cp_async4(dst + 32, lsrc + 32, policy);
                cp_async4(dst + 64, lsrc + 64, policy);
                if (MASS)
                    cp_async4(dst + 96, lmass, policy);
            } else {
                const long long fmax = 3 * a.n, f = 3 * pl + lane;
                if (f < fmax) cp_async4(dst, lsrc, policy);
                if (f + 32 < fmax) cp_async4(dst + 32, lsrc + 32, policy);
                if (f + 64 < fmax) cp_async4(dst + 64, lsrc + 64, policy);
                if (MASS && pl + lane < a.n)
                    cp_async4(dst + 96, lmass, policy);
            }
            const long long inc = (++l_r == ry_eff) ? plane_inc : g.n0;
            if (l_r == ry_eff) l_r = 0;
            pl += inc;
            lsrc += 3 * inc;
            if (MASS) lmass += inc;
        }
        cp_async_commit();
    };
#pragma unroll
    for (int q = 0; q < MARCH_STAGES - 1; q++)
        issue_load(q);

    // y-carry: the two high-y sums of the previous row
    bool yc_has = false;
    key_t yc_kc = 0, yc_kd = 0;
    acc_t yc_c = 0, yc_d = 0;

    // one slot of the x-carry: merge what the previous plane left for this row, emit the
    // low-x sum, leave the high-x sum for the next plane
    auto x_slot = [&](acc_t *xv_p, key_t *xk_p, bool has, key_t ka, acc_t va, key_t kb, acc_t vb) {
        const key_t xk = *xk_p;
        if (xk != INVALID) {
            const acc_t xv = *xv_p;
            if (has && xk == ka)
                va += xv;
            else
                Acc<FIXED>::red(grid, (size_t)xk, xv);
        }
        if (has) {
            Acc<FIXED>::red(grid, (size_t)ka, va);
            *xv_p = vb;
        }
        *xk_p = has ? kb : INVALID;
    };

    int t_lo = 0x7fffffff, t_hi = -1;        // slab: lowest / highest low-x plane of a deposited cloud
    int r = 0;
    long long pc = p_first + lane;           // this lane's particle at the current step (only read when !FULL)
    acc_t *xv_p = xc_val0;
    key_t *xk_p = xc_key0;
    const float *my_stage = stage + 3 * lane;
    const bool emit_rule_full = pair_in_row ? lane < 31 : owner_lane;
    for (int q = 0; q < nsteps; q++) {
        issue_load(q + MARCH_STAGES - 1);
        cp_async_wait<MARCH_STAGES - 1>();
        __syncwarp();
        const int slot = (q & (MARCH_STAGES - 1)) * MARCH_ROW_FLOATS;
        const bool live = lane_in_row && (FULL || pc < a.n);
        float px = my_stage[slot], py = my_stage[slot + 1], pz = my_stage[slot + 2];
        double m = a.cmass;
        if (MASS)
            m = (double)stage[slot + 96 + lane];                             // fieldize.cpp:63
        __syncwarp();                                                        // slot may be refilled from here on
        bool ok = live && fabsf(px) < pos_limit && fabsf(py) < pos_limit && fabsf(pz) < pos_limit;
        if (!ok)
            px = py = pz = 0.f;                                              // stale / non-finite staging data
        int fx, fy, fz;
        double tx, dx, ty, dy, tz, dz;
        axis_fast(px, units, fx, tx, dx);
        axis_fast(py, units, fy, ty, dy);
        axis_fast(pz, units, fz, tz, dz);
        if (((unsigned)fx >= (unsigned)dims) | ((unsigned)fy >= (unsigned)dims) | ((unsigned)fz >= (unsigned)dims)) {
            fx = wrap_cell(fx, dims);                                        // periodic wrap, fieldize.cpp:70-75
            fy = wrap_cell(fy, dims);
            fz = wrap_cell(fz, dims);
        }
        int xl = fx;
        int xstep = 1;                                                       // planes from the low-x to the high-x corner
        if (a.slab) {                                                        // slab: the +1 neighbour may be a ghost plane
            xl = slab_plane(fx, a.x0, a.ghost_lo, dims);
            ok = ok && xl >= 0 && xl <= a.xl_max;
        } else if (fx + 1 == dims) {
            xstep = 1 - dims;
        }
        if (live && !ok && owner_lane)
            atomicAdd(a.errors, 1ull);
        if (a.slab && ok) {                                                  // planes the ghost exchange has to move
            t_lo = xl < t_lo ? xl : t_lo;
            t_hi = xl > t_hi ? xl : t_hi;
        }
        const int ystep = fy + 1 == dims ? 1 - dims : 1;
        const int zoff = fz + 1 == dims ? 1 - dims : 1;                      // +1, or back to 0 at the wrap
        const double mx0 = __dmul_rn(m, tx), mx1 = __dmul_rn(m, dx);
        const double w00 = __dmul_rn(mx0, ty), w10 = __dmul_rn(mx1, ty);
        const double w01 = __dmul_rn(mx0, dy), w11 = __dmul_rn(mx1, dy);
        // weights in the order of fieldize.cpp:77-84
        acc_t lo[4], hi[4];
        lo[0] = Acc<FIXED>::make(__dmul_rn(w00, tz), a.scale);
        lo[1] = Acc<FIXED>::make(__dmul_rn(w10, tz), a.scale);
        lo[2] = Acc<FIXED>::make(__dmul_rn(w01, tz), a.scale);
        lo[3] = Acc<FIXED>::make(__dmul_rn(w11, tz), a.scale);
        hi[0] = Acc<FIXED>::make(__dmul_rn(w00, dz), a.scale);
        hi[1] = Acc<FIXED>::make(__dmul_rn(w10, dz), a.scale);
        hi[2] = Acc<FIXED>::make(__dmul_rn(w01, dz), a.scale);
        hi[3] = Acc<FIXED>::make(__dmul_rn(w11, dz), a.scale);
        // linear cell indices of the four low-z corners (cell order of fieldize.cpp:85-92)
        const key_t cell = ((key_t)xl * (key_t)dims + (key_t)fy) * kfd + (key_t)fz;
        const key_t ka = ok ? cell : INVALID;
        const key_t kb = cell + (key_t)xstep * kplane;
        const key_t kc = cell + (key_t)ystep * kfd;
        const key_t kd = kb + (key_t)ystep * kfd;

        // ---- z: hand the high-z corners to the next lane when it sits one cell further ----
        const key_t next_ka = __shfl_down_sync(0xffffffffu, ka, 1);
        const bool give = lane < 31 && ok && next_ka == (key_t)(cell + (key_t)zoff);
        const int given = __shfl_up_sync(0xffffffffu, (int)give, 1);         // every lane takes part in the shuffle
        const int take = lane > 0 ? given : 0;
#pragma unroll
        for (int c = 0; c < 4; c++)
            add_if(lo[c], __shfl_up_sync(0xffffffffu, hi[c], 1), take);
        // high-z corners nobody took: the pair (lane, lane+1) belongs to this warp when
        // lane < 31; a row's last particle has no pair and flushes in its owner lane
        const bool emit_rule = FULL ? emit_rule_full : ((pair_in_row && pc + 1 < a.n) ? lane < 31 : owner_lane);
        if (ok && !give && emit_rule) {
            Acc<FIXED>::red(grid, (size_t)(key_t)(cell + (key_t)zoff), hi[0]);
            Acc<FIXED>::red(grid, (size_t)(key_t)(kb + (key_t)zoff), hi[1]);
            Acc<FIXED>::red(grid, (size_t)(key_t)(kc + (key_t)zoff), hi[2]);
            Acc<FIXED>::red(grid, (size_t)(key_t)(kd + (key_t)zoff), hi[3]);
        }
        const bool has = ok && owner_lane;

        // ---- y: the previous row's high-y sums are this row's low-y sums ----
        if (yc_has) {
            if (has && yc_kc == ka) {
                lo[0] += yc_c;
                lo[1] += yc_d;
            } else {
                Acc<FIXED>::red(grid, (size_t)yc_kc, yc_c);
                Acc<FIXED>::red(grid, (size_t)yc_kd, yc_d);
            }
        }
        yc_has = has;
        yc_kc = kc;
        yc_kd = kd;
        yc_c = lo[2];
        yc_d = lo[3];

        // ---- x: slot r of the plane carry ----
        x_slot(xv_p, xk_p, has, ka, lo[0], kb, lo[1]);
        xv_p += MARCH_THREADS;
        xk_p += MARCH_THREADS;
        if (++r == ry_eff) {
            // the last row's high-y sums leave through slot ry_eff
            x_slot(xv_p, xk_p, yc_has, yc_kc, yc_c, yc_kd, yc_d);
            yc_has = false;
            r = 0;
            xv_p = xc_val0;
            xk_p = xc_key0;
            if (!FULL) pc += plane_inc;
        } else {
            if (!FULL) pc += g.n0;
        }
    }
    // ---- what the last plane left behind ----
    for (int s = 0; s <= ry_eff; s++) {
        const key_t xk = xc_key0[s * MARCH_THREADS];
        if (xk != INVALID)
            Acc<FIXED>::red(grid, (size_t)xk, xc_val0[s * MARCH_THREADS]);
    }
    if (a.touched) {                         // slab: {lowest, highest} plane written, behind the grid allocation
        t_lo = __reduce_min_sync(0xffffffffu, t_lo);
        t_hi = __reduce_max_sync(0xffffffffu, t_hi);
        if (lane == 0 && t_hi >= 0) {
            atomicMin(a.touched, t_lo);
            atomicMax(a.touched + 1, t_hi + 1);
        }
    }
}

// ---------------------------------------------------------------------------------
// Order probe: is the particle array spatially coherent, and is it a lattice?
// One CTA.  (1) pairs (i, i+1) within 4 cells => coherent (no brick sort needed).
// (2) candidate row lengths n0 -- caller hints, the cube root of the particle count,
// and the mean distance between the big backward jumps of z along the array -- are
// scored by how often particle p+n0 lands exactly one cell further in y than p (the
// condition for the y hand-over above); likewise n0*n1 for x.  The result is only
// ever a performance choice: the march kernel is exact for any (n0, n1).
// ---------------------------------------------------------------------------------
constexpr int PROBE_CANDS = 4;
struct ProbeArgs {
    const float *pos;
    long long n;
    double units;
    int dims;
    long long cand_n0[PROBE_CANDS], cand_n1[PROBE_CANDS];   // 0 = unused
    OrderInfo *out;
    // where lattice plane x sits along x (for the sweep kernel's zero-ahead window)
    int slab, x0, ghost_lo, n_local_planes;                 // slab geometry (slab = 0: the whole periodic grid)
    int nx;                                                 // grid planes this context owns
};

__device__ __forceinline__ void probe_cells(const float *pos, long long p, double units, int dims, int c[3])
{
    for (int ax = 0; ax < 3; ax++)
        c[ax] = axis_cell(pos[3 * p + ax], units, dims).lo;
}

__device__ __forceinline__ bool step_is(const int a[3], const int b[3], int dims, int axis)
{
    // b is exactly one cell further than a along `axis` and in the same cell otherwise
    for (int ax = 0; ax < 3; ax++) {
        const int want = ax == axis ? (a[ax] + 1 == dims ? 0 : a[ax] + 1) : a[ax];
        if (b[ax] != want)
            return false;
    }
    return true;
}

__global__ void __launch_bounds__(1024) order_probe_kernel(ProbeArgs A)
{
    __shared__ int s_near, s_tried;
    __shared__ int s_y[PROBE_CANDS + 1], s_x[PROBE_CANDS + 1], s_z;
    __shared__ long long s_jmin, s_jmax;
    __shared__ int s_jcount;
    __shared__ long long s_det_n0;
    __shared__ long long s_lat_n0, s_lat_n1;
    __shared__ int s_dref, s_dmin, s_dmax, s_dcount;
    __shared__ long long s_dsum;
    const int tid = threadIdx.x;
    if (tid == 0) {
        s_near = s_tried = s_z = 0;
        s_jmin = 0x7fffffffffffffffLL;
        s_jmax = -1;
        s_jcount = 0;
        s_det_n0 = 0;
    }
    if (tid <= PROBE_CANDS) {
        s_y[tid] = 0;
        s_x[tid] = 0;
    }
    __syncthreads();
    const long long n = A.n;
    const int samples = 4096;
    // (1) coherence of consecutive particles, (z) exact z hand-over rate
    // sample positions scattered by a multiplicative hash (an even stride would always
    // land on the same place of a lattice row)
    auto sample_at = [&](int s) { return (long long)(((unsigned long long)(s + 1) * 0x9E3779B97F4A7C15ull >> 20) % (unsigned long long)(n - 1)); };
    for (int s = tid; s < samples && n > 1; s += blockDim.x) {
        const long long i = sample_at(s);
        int c0[3], c1[3];
        probe_cells(A.pos, i, A.units, A.dims, c0);
        probe_cells(A.pos, i + 1, A.units, A.dims, c1);
        bool all = true;
        for (int ax = 0; ax < 3; ax++) {
            int d = abs(c0[ax] - c1[ax]);
            d = min(d, A.dims - d);
            all = all && d <= 4;
        }
        atomicAdd(&s_tried, 1);
        if (all) atomicAdd(&s_near, 1);
        if (step_is(c0, c1, A.dims, 2)) atomicAdd(&s_z, 1);
    }
    // (2a) row length from the backward jumps of z over the head of the array
    const long long window = n - 1 < (1LL << 18) ? n - 1 : (1LL << 18);
    for (long long i = tid; i < window; i += blockDim.x) {
        const int z0 = axis_cell(A.pos[3 * i + 2], A.units, A.dims).lo;
        const int z1 = axis_cell(A.pos[3 * (i + 1) + 2], A.units, A.dims).lo;
        if (z1 + A.dims / 2 < z0) {
            atomicMin(&s_jmin, i);
            atomicMax(&s_jmax, i);
            atomicAdd(&s_jcount, 1);
        }
    }
    __syncthreads();
    if (tid == 0 && s_jcount >= 2)
        s_det_n0 = (long long)llrint((double)(s_jmax - s_jmin) / (double)(s_jcount - 1));
    __syncthreads();
    // (2b) score the candidates
    for (int c = 0; c <= PROBE_CANDS; c++) {
        const long long n0 = c < PROBE_CANDS ? A.cand_n0[c] : s_det_n0;
        long long n1 = c < PROBE_CANDS ? A.cand_n1[c] : 0;
        if (n0 < 2 || n0 >= n)
            continue;
        if (n1 < 1)
            n1 = n0;                                   // cubic lattice guess
        for (int s = tid; s < samples; s += blockDim.x) {
            const long long i = sample_at(s);
            int c0[3], c1[3];
            if (i + n0 < n) {
                probe_cells(A.pos, i, A.units, A.dims, c0);
                probe_cells(A.pos, i + n0, A.units, A.dims, c1);
                if (step_is(c0, c1, A.dims, 1)) atomicAdd(&s_y[c], 1);
                if (i + n0 * n1 < n) {
                    probe_cells(A.pos, i + n0 * n1, A.units, A.dims, c1);
                    if (step_is(c0, c1, A.dims, 0)) atomicAdd(&s_x[c], 1);
                }
            }
        }
    }
    __syncthreads();
    if (tid == 0) {
        OrderInfo o;
        const int tried = s_tried > 0 ? s_tried : 1;
        o.coherent = (s_tried > 0 && s_near * 10 >= s_tried * 6) ? 1 : 0;
        o.samples = s_tried;
        o.score_z = (int)(1000LL * s_z / tried);
        int best = -1;
        for (int c = 0; c <= PROBE_CANDS; c++)
            if (s_y[c] > 0 && (best < 0 || s_y[c] > s_y[best]))
                best = c;
        o.lattice = 0;
        o.n0 = o.n1 = 0;
        o.score_y = o.score_x = 0;
        if (best >= 0) {
            o.score_y = (int)(1000LL * s_y[best] / tried);
            o.score_x = (int)(1000LL * s_x[best] / tried);
            if (o.coherent && o.score_y >= 400) {
                o.lattice = 1;
                o.n0 = best < PROBE_CANDS ? A.cand_n0[best] : s_det_n0;
                const long long n1 = best < PROBE_CANDS && A.cand_n1[best] > 0 ? A.cand_n1[best] : o.n0;
                o.n1 = o.score_x >= 400 ? n1 : 0;
            }
        }
        o.dx_valid = o.dx_mean = o.dx_dev = 0;
        *A.out = o;
        s_lat_n0 = o.lattice ? o.n0 : 0;
        s_lat_n1 = o.lattice ? (o.n1 > 0 ? o.n1 : (n + o.n0 - 1) / o.n0) : 0;
        s_dref = 0x7fffffff;
        s_dmin = 0x7fffffff;
        s_dmax = -0x7fffffff;
        s_dcount = 0;
        s_dsum = 0;
    }
    __syncthreads();
    // (3) lattice: offset along x between a particle's (local) grid plane and the plane its lattice
    // plane is expected at, floor((x_lat + 1/2) * nx / n2): mean and extremes over the samples
    if (s_lat_n0 == 0)
        return;
    const long long per_plane = s_lat_n0 * s_lat_n1;
    const long long n2 = (n + per_plane - 1) / per_plane;
    const double planes_per = (double)A.nx / (double)n2;
    auto offset_of = [&](long long i, int &d) {
        const AxisCell cx = axis_cell(A.pos[3 * i], A.units, A.dims);
        if (!cx.ok)
            return false;
        int xl = cx.lo;
        if (A.slab) {
            xl = slab_plane(cx.lo, A.x0, A.ghost_lo, A.dims);
            if (xl < 0 || xl >= A.n_local_planes)
                return false;
        }
        d = xl - (int)floor(((double)(i / per_plane) + 0.5) * planes_per);
        return true;
    };
    if (tid == 0) {
        int d;
        for (int s = 0; s < 64; s++)
            if (offset_of(sample_at(s), d)) {
                s_dref = d;
                break;
            }
    }
    __syncthreads();
    if (s_dref == 0x7fffffff)
        return;
    for (int s = tid; s < samples && n > 1; s += blockDim.x) {
        int d;
        if (!offset_of(sample_at(s), d))
            continue;
        d -= s_dref;
        if (!A.slab) {                                      // periodic: the image nearest to the reference offset
            if (d > A.dims / 2) d -= A.dims;
            if (d < -A.dims / 2) d += A.dims;
        }
        atomicMin(&s_dmin, d);
        atomicMax(&s_dmax, d);
        atomicAdd(&s_dcount, 1);
        atomicAdd((unsigned long long *)&s_dsum, (unsigned long long)(long long)d);
    }
    __syncthreads();
    if (tid == 0 && s_dcount > 0) {
        const double mean = (double)s_dsum / (double)s_dcount;
        const int m = (int)floor(mean + 0.5);
        A.out->dx_valid = 1;
        A.out->dx_mean = s_dref + m;
        const int up = s_dmax - m, down = m - s_dmin;
        A.out->dx_dev = up > down ? up : down;
    }
}

static long long exact_cbrt(long long n)
{
    long long c = (long long)llround(cbrt((double)n));
    for (long long t = c - 1; t <= c + 1; t++)
        if (t > 0 && t * t * t == n)
            return t;
    return 0;
}

// Runs the probe on the head of the stream and returns its verdict (one small D2H;
// the caller decides which kernels to launch).
int probe_order(genpk_ctx *ctx, const float *pos, int64_t n, double units, OrderInfo *info)
{
    if (!ctx->d_order)
        GENPK_CUDA_OK(cudaMalloc(&ctx->d_order, sizeof(OrderInfo)));
    ProbeArgs A;
    A.pos = pos;
    A.n = n;
    A.units = units;
    A.dims = ctx->g.dims;
    for (int c = 0; c < PROBE_CANDS; c++)
        A.cand_n0[c] = A.cand_n1[c] = 0;
    int c = 0;
    if (ctx->lattice_n0 > 0) {
        A.cand_n0[c] = ctx->lattice_n0;
        A.cand_n1[c] = ctx->lattice_n1;
        c++;
    }
    if (long long k = exact_cbrt(n))
        A.cand_n0[c++] = k;
    if (ctx->g.nranks > 1)
        if (long long k = exact_cbrt((long long)n * ctx->g.nranks))      // an x-slab of a cubic lattice
            A.cand_n0[c++] = k;
    A.out = reinterpret_cast<OrderInfo *>(ctx->d_order);
    A.slab = ctx->g.nranks > 1 ? 1 : 0;
    A.x0 = ctx->g.x0;
    A.ghost_lo = ctx->g.ghost_lo;
    A.n_local_planes = ctx->g.ghost_lo + ctx->g.nx + ctx->g.ghost_hi;
    A.nx = ctx->g.nx;
    order_probe_kernel<<<1, 1024, 0, ctx->stream>>>(A);
    ctx->launches++;
    GENPK_CUDA_OK(cudaGetLastError());
    GENPK_CUDA_OK(cudaMemcpyAsync(info, ctx->d_order, sizeof(OrderInfo), cudaMemcpyDeviceToHost, ctx->stream));
    GENPK_CUDA_OK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

template <bool FIXED, typename key_t, bool FULL, bool MASS>
static int launch_march_t(genpk_ctx *ctx, const DepositArgs &a, const MarchGeom &g)
{
    auto kern = deposit_march_kernel<FIXED, key_t, FULL, MASS>;
    const size_t smem = (size_t)(g.ry + 1) * MARCH_THREADS * (sizeof(typename Acc<FIXED>::type) + sizeof(key_t)) +
                        (size_t)MARCH_WARPS * MARCH_STAGES * MARCH_ROW_FLOATS * sizeof(float);
    GENPK_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long blocks = (g.n_tasks + MARCH_WARPS - 1) / MARCH_WARPS;
    if (blocks > 0x7fffffffLL) {
        set_error("deposit: %lld march tasks exceed the launch grid; split the call", g.n_tasks);
        return 1;
    }
    kern<<<(int)blocks, MARCH_THREADS, smem, ctx->stream>>>(a, g);
    ctx->launches++;
    GENPK_CUDA_OK(cudaGetLastError());
    return 0;
}

template <bool FIXED, typename key_t>
static int launch_march_k(genpk_ctx *ctx, const DepositArgs &a, const MarchGeom &g)
{
    const bool full = g.n0 * g.n1 * g.n2 == a.n;
    if (a.mass)
        return full ? launch_march_t<FIXED, key_t, true, true>(ctx, a, g) : launch_march_t<FIXED, key_t, false, true>(ctx, a, g);
    return full ? launch_march_t<FIXED, key_t, true, false>(ctx, a, g) : launch_march_t<FIXED, key_t, false, false>(ctx, a, g);
}

// n0: particles per lattice row; n1: rows per plane (0 = unknown: the whole array is one plane).
int launch_march(genpk_ctx *ctx, const DepositArgs &a, long long n0, long long n1)
{
    if (a.n <= 0)
        return 0;
    MarchGeom g;
    if (n0 < 1 || n0 > a.n) n0 = a.n;
    const long long rows = (a.n + n0 - 1) / n0;
    if (n1 < 1 || n1 > rows) n1 = rows;
    g.n0 = n0;
    g.n1 = n1;
    g.n2 = (rows + n1 - 1) / n1;
    g.ry = (int)(n1 < ctx->march_ry ? n1 : ctx->march_ry);
    g.rx = (int)(g.n2 < ctx->march_rx ? g.n2 : ctx->march_rx);
    g.nzs = n0 >= 2 ? (int)((n0 - 2) / 31 + 1) : 1;
    // y bands: about eight grid planes of one band should fit in a fraction of L2
    const size_t row_bytes = (size_t)a.fd * sizeof(double);
    const size_t budget = ctx->l2_bytes ? ctx->l2_bytes * 5 / 8 : (size_t)64 << 20;
    long long band = (long long)(budget / (8 * row_bytes));
    band = band / g.ry * g.ry;
    if (band < g.ry) band = g.ry;
    if (band > n1) band = (n1 + g.ry - 1) / g.ry * g.ry;
    g.band_rows = (int)band;
    g.yb_per_band = g.band_rows / g.ry;
    g.nbands = (n1 + band - 1) / band;
    g.nxb = (g.n2 + g.rx - 1) / g.rx;
    g.tasks_per_layer = (long long)g.nzs * g.yb_per_band;
    g.n_tasks = g.nbands * g.nxb * g.tasks_per_layer;
    const bool key32 = ctx->g.grid_doubles() < 0xfffffff0ull;
    if (ctx->fixed)
        return key32 ? launch_march_k<true, uint32_t>(ctx, a, g) : launch_march_k<true, unsigned long long>(ctx, a, g);
    return key32 ? launch_march_k<false, uint32_t>(ctx, a, g) : launch_march_k<false, unsigned long long>(ctx, a, g);
}

}  // namespace genpk
