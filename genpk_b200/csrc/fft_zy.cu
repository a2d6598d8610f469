// The (y,z) part of the r2c transform (gen-pk.cpp:193,233: fftw_plan_dft_r2c_3d) as ONE persistent kernel:
// real-to-complex along z on the contiguous rows of a plane, then complex along y on its columns, the y tiles of a
// plane scheduled a couple of planes behind its z tiles so that they find the rows in L2.  The padded real grid is
// read from HBM once and the (y,z)-transformed spectrum written once (16 B/cell in all), where the two-kernel route
// (batched 1-D library transform along z + fft_cols_kernel along y) moves 32 B/cell.
//
// z tile: CZ consecutive rows of a plane.  In the FFTW padded layout they ARE one contiguous block of
//   CZ * (dims/2 + 1) * 16 bytes (a row of dims reals plus its two padding doubles is exactly the dims/2 + 1 complex
//   values the row becomes), so a tile comes in with ONE bulk copy.  Then: the half-length complex FFT of
//   z[n] = x[2n] + i x[2n+1] with the three register passes of a Plan (fftx_core.cuh), a third exchange that hands
//   every thread eight pairs (Z[k], Z[dims/2 - k]), the untangling of rfft_pair, and the results stored from the
//   registers over the rows they came from (64-byte runs).  The tile buffer is free as soon as the registers are
//   loaded, so the next tile arrives while this one is worked on.  A grid deposited in int64 fixed point is
//   converted on the way into the registers.
// y tile: exactly the tile of fft_cols_kernel (N = dims values of y for C adjacent kz columns), filled by TMA
//   tensor copies once the plane's z tiles have all been stored (one counter per plane), stored from registers --
//   in place, or to the owner ranks' transposed blocks (SCATTER, the multi-GPU transpose).
//
// Tiles are dealt round-robin in a fixed order (slot s -> CTA s % gridDim.x): plane by plane, the z tiles of plane p
// and then the y tiles of plane p - lag.  A y tile waits only for z tiles that come earlier in that order, and z tiles
// wait for nothing, so with every CTA resident (cooperative launch) the schedule cannot deadlock.
#include <cooperative_groups.h>

#include <type_traits>

#include "fft_tiles.cuh"

namespace genpk {

using namespace fftx;

struct FftZyArgs {
    alignas(64) CUtensorMap tmap;   // the planes as {2*nc doubles, dims rows, n_planes slabs}; box {2*CY, min(dims,256), 1}
    double2 *spec;            // [n_planes][dims][nc], transformed in place (or read only, with SCATTER)
    const double2 *tw;        // exp(-2 pi i t / dims), t < dims
    const double2 *tw_half;   // exp(-2 pi i t / (dims/2)), t < dims/2
    int dims, nc;
    long long plane_stride;   // modes per plane: dims * nc
    int n_planes;
    int zt, yt;               // z tiles and y tiles per plane
    int lag;                  // the y tiles of plane p follow the z tiles of plane p + lag
    int n_slots;              // (n_planes + lag) * (zt + yt)
    int *rows_done;           // per plane: warps of z tiles that have stored their rows (zero at launch)
    int from_fixed;           // the rows hold int64 fixed-point sums
    double inv_scale;
    // SCATTER: see FftColsArgs (fftx_power.cu)
    double2 *peer[GENPK_MAX_PEERS];
    int ny_shift, dst_pitch, x0;
};

__device__ __forceinline__ void bulk_load(void *smem_dst, const void *gsrc, unsigned bytes, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc), "r"(bytes),
                   "r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void bulk_store(void *gdst, const void *smem_src, unsigned bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 ::"l"(gdst), "r"((unsigned)__cvta_generic_to_shared(smem_src)), "r"(bytes) : "memory");
}
// The counters are read past L1 and never through an acquire: an acquire load (and __threadfence) invalidates the
// whole L1 -- the twiddle table with it (measured: the passes then wait on L2 for every twiddle).  What the reader
// needs ordered after the counter is a TMA copy, which does not go through L1 and is issued behind a branch on the value.
__device__ __forceinline__ int ld_counter(const int *p)
{
    int v;
    asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// stores of this CTA that the calling thread has observed (through a barrier) -> visible before the increment
__device__ __forceinline__ void announce(int *counter)
{
    asm volatile("fence.proxy.async.global;" ::: "memory");
    asm volatile("red.release.gpu.global.add.s32 [%0], 1;" ::"l"(counter) : "memory");
}

template <class PLZ, class PLY, bool SCATTER>
__global__ void __launch_bounds__(PLY::THREADS, PLY::TILE == 4096 ? 2 : 1) fft_zy_kernel(const __grid_constant__ FftZyArgs A)
{
    static_assert(PLZ::THREADS == PLY::THREADS && PLZ::TILE == PLY::TILE, "both tile kinds run on the same CTA");
    static_assert(2 * PLZ::N == PLY::N, "the row transform is the half-length complex one");
    constexpr int TILE_MODES = PLY::TILE;
    constexpr int NZ = PLZ::N, CZ = PLZ::C, TZ = PLZ::T;           // half-length transform, rows per tile, threads per row
    constexpr int NY = PLY::N, CY = PLY::C;
    constexpr int ROW = NZ + 1;                                     // complex values of a padded row
    constexpr unsigned Z_BYTES = (unsigned)CZ * ROW * 16;
    constexpr size_t R0_BYTES = ((size_t)TILE_MODES * 16 + (size_t)CZ * 16 + 127) / 128 * 128;
    constexpr int Y_BOX_ROWS = NY < TMA_BOX_ROWS ? NY : TMA_BOX_ROWS;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    cd *const R0 = reinterpret_cast<cd *>(smem_raw);                // the tile: rows (z) or [NY][CY] (y)
    cd *const E = reinterpret_cast<cd *>(smem_raw + R0_BYTES);      // half-size exchange buffer
    __shared__ __align__(8) unsigned long long tile_bar;

    const int tid = threadIdx.x;
    if (tid == 0) {
        mbar_init(&tile_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    unsigned tile_parity = 0;

    const int per = A.zt + A.yt;
    const int G = (int)gridDim.x;
    const int rows_target = A.zt;                                   // every z tile of a plane is announced once
    // slot -> (kind, plane, tile index); false for the empty slots at both ends of the schedule
    struct Slot {
        int s, plane, idx;
        bool is_z;
    };
    auto decode = [&](int s, Slot &d) -> bool {
        const int P = (int)((unsigned)s / (unsigned)per), r = s - P * per;
        d.s = s;
        d.is_z = r < A.zt;
        d.plane = d.is_z ? P : P - A.lag;
        d.idx = d.is_z ? r : r - A.zt;
        return d.plane >= 0 && d.plane < A.n_planes;
    };
    auto next_slot = [&](int s, Slot &d) {
        while (s < A.n_slots && !decode(s, d))
            s += G;
        d.s = s;
    };
    // One thread asks for the tile of slot d.  A y tile needs the plane's rows: `seen` is a value of the plane's counter
    // read earlier (its latency has passed); if that is not enough the thread either waits (wait = true) or gives up.
    auto issue = [&](const Slot &d, int seen, bool wait) -> bool {
        if (d.s >= A.n_slots)
            return true;
        if (d.is_z) {
            mbar_expect_tx(&tile_bar, Z_BYTES);
            bulk_load(R0, A.spec + (size_t)d.plane * A.plane_stride + (size_t)d.idx * CZ * ROW, Z_BYTES, &tile_bar);
            return true;
        }
        if (seen < rows_target) {
            if (!wait)
                return false;
            while (ld_counter(A.rows_done + d.plane) < rows_target)
                __nanosleep(100);
        }
        asm volatile("fence.proxy.async.global;" ::: "memory");       // rows stored by other SMs' threads, read by the copy engine
        mbar_expect_tx(&tile_bar, (unsigned)(TILE_MODES * 16));
#pragma unroll 1
        for (int y = 0; y < NY; y += Y_BOX_ROWS)
            tma_load_3d(R0 + (size_t)y * CY, &A.tmap, 2 * d.idx * CY, y, d.plane, &tile_bar);
        return true;
    };
    auto peek = [&](const Slot &d) -> int {
        int seen = 0;
        if (tid == 0 && d.s < A.n_slots && !d.is_z)
            seen = ld_counter(A.rows_done + d.plane);
        return seen;
    };

    // The rows of a z tile are announced by thread 0 one tile later, at the point of the next tile where every warp is
    // known to be past its stores and the stores have had a microsecond to land (the release then costs little).
    __shared__ int next_ready;
    int owed = -1;                                                   // thread 0: plane of a stored, not yet announced z tile
    Slot cur, nxt;
    next_slot((int)blockIdx.x, cur);
    if (tid == 0)
        issue(cur, 0, true);
    while (cur.s < A.n_slots) {
        next_slot(cur.s + G, nxt);
        const int seen = peek(nxt);                                  // (in flight during the first pass)
        mbar_wait(&tile_bar, tile_parity);
        tile_parity ^= 1u;
        cd v[EPT], w[EPT];
        if (cur.is_z) {
            const int c = tid % CZ, t = tid / CZ;
            const cd *const row = R0 + c * ROW;
#pragma unroll
            for (int i = 0; i < EPT; i++)
                v[i] = row[PLZ::load_n(t, i)];
            if (A.from_fixed) {
#pragma unroll
                for (int i = 0; i < EPT; i++) {
                    v[i].x = (double)__double_as_longlong(v[i].x) * A.inv_scale;
                    v[i].y = (double)__double_as_longlong(v[i].y) * A.inv_scale;
                }
            }
            PLZ::pass1(v, t, A.tw_half);
            if (tid == 0)
                next_ready = nxt.s >= A.n_slots || nxt.is_z || seen >= rows_target;
            __syncthreads();                                         // everyone has taken its part of the tile; E is free
            // The next tile, while this one is worked on -- unless it is a y tile whose rows are not all there yet: those
            // may be this very tile's, so it is asked for after they have been announced (below).
            const bool asked = next_ready != 0;
            if (tid == 0) {
                if (owed >= 0)
                    announce(A.rows_done + owed);
                owed = -1;
                if (asked)
                    issue(nxt, seen, false);
            }
            int b2w[4], b2r[4];
#pragma unroll
            for (int q = 0; q < 4; q++) {
                b2w[q] = PLZ::ex2_w_base(t, q);
                b2r[q] = PLZ::ex2_r_base(t, q);
            }
            const int b1r = PLZ::ex1_r_base(t);
            exchange<NZ, CZ, false, PLZ::ONE_UNIT2>(E, v, w, c, [&](int i) { return t + PLZ::ex1_w_part(i); },
                           [&](int i) { return b1r + PLZ::ex1_r_part(i); });
            PLZ::pass2(w, t, A.tw_half);
            __syncthreads();
            exchange<NZ, CZ, PLZ::ONE_UNIT2, false>(E, w, v, c, [&](int i) { return b2w[(i % PLZ::R2) & 3] + PLZ::ex2_w_part(i); },
                           [&](int i) { return b2r[(i % PLZ::R3) & 3] + PLZ::ex2_r_part(i); });
            PLZ::pass3(v);
            // Third exchange: Z leaves in natural order and comes back as the pairs (k, NZ - k), k = t + TZ*j, of this
            // thread -- lower half of the index space first (all the k), then the upper half (the partners and Z[NZ/2]).
            // Which half an element of v falls in is a property of its register index (zhalf, checked on the host).
            constexpr int PAIRS = NZ / 2 / TZ;
            static_assert(PAIRS * 2 == EPT, "a thread untangles eight pairs");
            const int kb = PLZ::out_k_base(t);
            cd mid = make_double2(0.0, 0.0);
            __syncthreads();
#pragma unroll
            for (int i = 0; i < EPT; i++)
                if (!PLZ::out_k_upper(i))
                    E[(kb + PLZ::out_k_part(i)) * CZ + c] = v[i];
            __syncthreads();
#pragma unroll
            for (int j = 0; j < PAIRS; j++)
                w[2 * j] = E[(t + TZ * j) * CZ + c];
            __syncthreads();
#pragma unroll
            for (int i = 0; i < EPT; i++)
                if (PLZ::out_k_upper(i))
                    E[(kb + PLZ::out_k_part(i) - NZ / 2) * CZ + c] = v[i];
            __syncthreads();
#pragma unroll
            for (int j = 0; j < PAIRS; j++) {
                const int k = t + TZ * j;
                w[2 * j + 1] = k == 0 ? w[0] : E[(NZ / 2 - k) * CZ + c];
            }
            if (t == 0)
                mid = E[c];
            double2 *const dst = A.spec + (size_t)cur.plane * A.plane_stride + (size_t)(cur.idx * CZ + c) * ROW;
            const cd wt = A.tw[t];                                   // exp(-2 pi i t / dims); the pairs' twiddles are 1/32-turn steps from it
            auto untangle = [&](auto J) {
                constexpr int j = decltype(J)::value;
                const int k = t + TZ * j;
                cd xk, xm;
                rfft_pair(w[2 * j], w[2 * j + 1], rfft_step<j>(wt), &xk, &xm);
                dst[k] = xk;
                dst[NZ - k] = xm;
            };
            untangle(std::integral_constant<int, 0>());
            untangle(std::integral_constant<int, 1>());
            untangle(std::integral_constant<int, 2>());
            untangle(std::integral_constant<int, 3>());
            untangle(std::integral_constant<int, 4>());
            untangle(std::integral_constant<int, 5>());
            untangle(std::integral_constant<int, 6>());
            untangle(std::integral_constant<int, 7>());
            if (t == 0)
                dst[NZ / 2] = make_double2(mid.x, -mid.y);
            if (asked) {
                owed = cur.plane;                                    // announced during the next tile
            } else {
                // (rare) the next tile may need these very rows: announce them now, then wait for the plane
                __syncthreads();
                if (tid == 0) {
                    announce(A.rows_done + cur.plane);
                    issue(nxt, 0, true);
                }
            }
        } else {
            const int c = tid % CY, t = tid / CY;
            const int kz = cur.idx * CY + c;
            const bool valid = kz < A.nc;
#pragma unroll
            for (int i = 0; i < EPT; i++)
                v[i] = valid ? R0[PLY::load_n(t, i) * CY + c] : make_double2(0.0, 0.0);
            PLY::pass1(v, t, A.tw);
            __syncthreads();                                         // everyone has taken its part of the tile; E is free
            if (tid == 0) {
                if (owed >= 0)
                    announce(A.rows_done + owed);
                owed = -1;
                issue(nxt, seen, true);
            }
            int b2w[4], b2r[4];
#pragma unroll
            for (int q = 0; q < 4; q++) {
                b2w[q] = PLY::ex2_w_base(t, q);
                b2r[q] = PLY::ex2_r_base(t, q);
            }
            const int b1r = PLY::ex1_r_base(t);
            exchange<NY, CY, false, PLY::ONE_UNIT2>(E, v, w, c, [&](int i) { return t + PLY::ex1_w_part(i); },
                           [&](int i) { return b1r + PLY::ex1_r_part(i); });
            PLY::pass2(w, t, A.tw);
            __syncthreads();
            exchange<NY, CY, PLY::ONE_UNIT2, false>(E, w, v, c, [&](int i) { return b2w[(i % PLY::R2) & 3] + PLY::ex2_w_part(i); },
                           [&](int i) { return b2r[(i % PLY::R3) & 3] + PLY::ex2_r_part(i); });
            PLY::pass3(v);
            const int kb = PLY::out_k_base(t);
            if (valid) {
                if (SCATTER) {
                    const int ny = 1 << A.ny_shift;
                    const size_t plane_off = (size_t)(A.x0 + cur.plane) * ny * A.dst_pitch + kz;
#pragma unroll
                    for (int i = 0; i < EPT; i++) {
                        const int ky = kb + PLY::out_k_part(i);
                        A.peer[ky >> A.ny_shift][plane_off + (size_t)(ky & (ny - 1)) * A.dst_pitch] = v[i];
                    }
                } else {
                    double2 *dst = A.spec + (size_t)cur.plane * A.plane_stride + kz;
#pragma unroll
                    for (int i = 0; i < EPT; i++)
                        dst[(size_t)(kb + PLY::out_k_part(i)) * A.nc] = v[i];
                }
            }
        }
        cur = nxt;
    }
    __syncthreads();                                                 // (a last z tile: every warp is past its stores)
    if (tid == 0 && owed >= 0)
        announce(A.rows_done + owed);
}

template <class PLZ, class PLY, bool SCATTER> static int launch_zy(genpk_ctx *ctx, FftZyArgs &A)
{
    auto kern = fft_zy_kernel<PLZ, PLY, SCATTER>;
    constexpr size_t R0_BYTES = ((size_t)PLY::TILE * 16 + (size_t)PLZ::C * 16 + 127) / 128 * 128;
    const size_t smem = R0_BYTES + (size_t)PLY::TILE * 8;
    A.zt = A.dims / PLZ::C;
    A.yt = (A.nc + PLY::C - 1) / PLY::C;
    A.n_slots = (A.n_planes + A.lag) * (A.zt + A.yt);
    const int box_rows = PLY::N < TMA_BOX_ROWS ? PLY::N : TMA_BOX_ROWS;
    if (int rc = make_tile_map(&A.tmap, A.spec, A.nc, PLY::N, A.nc, A.n_planes, A.plane_stride, PLY::C, box_rows, 1)) return rc;
    GENPK_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    GENPK_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, PLY::THREADS, smem));
    if (per_sm < 1) {
        set_error("fused (y,z) pass: %zu bytes of shared memory do not fit", smem);
        return 1;
    }
    long long ctas = (long long)ctx->sm_count * per_sm;
    if (ctas > A.n_slots) ctas = A.n_slots;
    if (ctas < 1) ctas = 1;
    void *params[] = {(void *)&A};
    // cooperative: every CTA is resident, which is what the wait of a y tile on earlier z tiles relies on
    GENPK_CUDA_OK(cudaLaunchCooperativeKernel((const void *)kern, dim3((unsigned)ctas), dim3(PLY::THREADS), params, smem, ctx->stream));
    ctx->launches++;
    return 0;
}

bool fft_zy_supported(const genpk_ctx *ctx)
{
    const int d = ctx->g.dims;
    return ctx->fused_zy != 0 && ctx->use_tma && ctx->own_ypass != 0 && (d == 256 || d == 512 || d == 1024 || d == 2048) &&
           (size_t)8192 * 24 + 1024 <= (size_t)ctx->smem_optin && ctx->coop_launch && tma_available();
}

template <bool SCATTER> static int zy_dispatch(genpk_ctx *ctx, FftZyArgs &A)
{
    switch (ctx->g.dims) {
    case 256: return launch_zy<Plan<2, 8, 8, 4096>, Plan<4, 8, 8, 4096>, SCATTER>(ctx, A);
    case 512: return launch_zy<Plan<4, 8, 8, 4096>, Plan<8, 8, 8, 4096>, SCATTER>(ctx, A);
    case 1024:
        // 8192-mode tiles (one CTA of 512 threads per SM, 128-byte y-tile rows) for the scatter variant, whose rows
        // leave as peer stores over NVLink: C3 on two GPUs 3.65 ms against 5.41 ms with 64-byte stores (profiles/r02)
        if (ctx->fused_zy == 2 || SCATTER)
            return launch_zy<Plan<8, 8, 8, 8192>, Plan<16, 8, 8, 8192>, SCATTER>(ctx, A);
        return launch_zy<Plan<8, 8, 8, 4096>, Plan<16, 8, 8, 4096>, SCATTER>(ctx, A);
    case 2048: return launch_zy<Plan<16, 8, 8, 8192>, Plan<16, 16, 8, 8192>, SCATTER>(ctx, A);
    }
    set_error("fused (y,z) pass: unsupported grid side %d", ctx->g.dims);
    return 1;
}

// (y,z) transform of n_planes padded real planes starting at `planes`, in place -- or, with scatter, the y pass's
// results go to the owner ranks' transposed blocks (the planes then hold the z-transformed rows).
int fft_zy(genpk_ctx *ctx, double *planes, int n_planes, bool scatter, bool from_fixed, int scale_bits)
{
    const SlabGeom &g = ctx->g;
    if (int rc = ensure_twiddles(ctx)) return rc;
    if (!ctx->d_rows_done || ctx->rows_done_n < n_planes) {
        if (ctx->d_rows_done) cudaFree(ctx->d_rows_done);
        ctx->d_rows_done = nullptr;
        GENPK_CUDA_OK(cudaMalloc(&ctx->d_rows_done, (size_t)n_planes * sizeof(int)));
        ctx->rows_done_n = n_planes;
    }
    GENPK_CUDA_OK(cudaMemsetAsync(ctx->d_rows_done, 0, (size_t)n_planes * sizeof(int), ctx->stream));
    FftZyArgs A = {};
    A.spec = reinterpret_cast<double2 *>(planes);
    A.tw = reinterpret_cast<const double2 *>(ctx->d_twiddle);
    A.tw_half = A.tw + g.dims;
    A.dims = g.dims;
    A.nc = g.nc;
    A.plane_stride = (long long)g.dims * g.nc;
    A.n_planes = n_planes;
    A.lag = ctx->zy_lag < 1 ? 1 : ctx->zy_lag;
    A.rows_done = ctx->d_rows_done;
    A.from_fixed = from_fixed ? 1 : 0;
    A.inv_scale = ldexp(1.0, -scale_bits);
    if (scatter) {
        const int ny = g.dims / g.nranks;
        if (!ctx->peers_set || g.nranks > GENPK_MAX_PEERS || (ny & (ny - 1)) != 0) {
            set_error("scatter (y,z) pass: peers not set, more than %d ranks, or dims/nranks not a power of two", GENPK_MAX_PEERS);
            return 1;
        }
        for (int r = 0; r < g.nranks; r++)
            A.peer[r] = reinterpret_cast<double2 *>(ctx->peer_recv[r]);
        while ((1 << A.ny_shift) < ny) A.ny_shift++;
        A.x0 = g.x0;
        A.dst_pitch = recv_row_pitch(ctx);
        return zy_dispatch<true>(ctx, A);
    }
    return zy_dispatch<false>(ctx, A);
}

}  // namespace genpk
