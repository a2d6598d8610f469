// Fused x-pass: the last 1-D transform of the r2c FFT (gen-pk.cpp:233) and the
// |delta_k|^2 binning of powerspectrum() (powerspectrum.c:35-110) in ONE kernel.
//
// Input: the spectrum after the batched 2-D (y,z) transform, [x][n_mid][nc] complex
// doubles.  A tile is the N x-values of C adjacent kz columns of one ky row (N*C = 8192
// modes = 128 KB; 4096 for the small grids).  Persistent CTAs, 16 modes per thread:
//
//   cp.async   next tile  -> shared memory (natural [x][c] layout), issued after pass 1 } overlapped with
//   registers  <- this tile; three register passes of the length-N FFT                  } everything below
//   two exchanges between the passes through a half-tile buffer of complex doubles (lower
//     half of the index space, then the upper; layouts chosen so both sides are conflict free)
//   |X|^2 -> the same buffer in kx order; bin walk along |kx| with the +-kx modes folded
//     (forward-only threshold walk) into the CTA's histograms
//
// (Measured alternatives, profiles/r01: loading the next tile straight into registers during
// the bin walk stalls on the load queue; 4096-mode tiles with two CTAs per SM at 1024 overfetch.)
// so the x-transformed spectrum is never written and never re-read: this pass moves
// 16 B/mode once (8.6 GB at 1024^3) where cuFFT's in-place x pass plus bin_power_kernel
// move 48 B/mode.  Thread-level arithmetic lives in fftx_core.cuh, which is also compiled
// for the host and checked against numpy (tests/test_fftx_core.py).
//
// Supported: dims in {256, 512, 1024, 2048}, auto spectra.  Everything else (other sizes,
// cross spectra, callers that want the transformed grid) keeps the cuFFT + bin_power path.
#include "fft_tiles.cuh"

namespace genpk {

using namespace fftx;

struct FftxArgs {
    alignas(64) CUtensorMap tmap;   // the spectrum as {2*nc doubles, n_mid rows, dims slabs}; box {2C, 1, 256}
    alignas(64) CUtensorMap zmap;   // the same tensor with box {2C, 1, ZERO_BOX_ROWS}: zero stores behind the read
    int use_tma;
    int halves_delay_ns;      // fftx_power_halves_kernel: head start of the first half
    int zero_after;           // 1: every tile is overwritten with zeros once it has been read (the grid is clear for the next deposit)
    const double2 *spec;      // [dims][n_mid][nc]
    const double2 *tw;        // exp(-2 pi i t / dims), t < dims
    int dims, nc, n_mid, mid0;
    int row_pitch;            // modes between consecutive mid rows (nc, or the padded pitch of a transposed block)
    long long x_stride;       // modes between consecutive x: n_mid * row_pitch
    int groups;               // column groups per (x, mid) row: ceil(nc / C)
    long long n_tiles;        // n_mid * groups
    int nrbins;
    const float *iw1d;
    const uint32_t *thresh;
    float half_bpu;
    double *sums;             // nrbins P sums, accumulated into
    int hists;                // histograms per CTA (1 or 2: even / odd tile columns)
};

typedef CUresult (*tmap_encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                   const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// cuTensorMapEncodeTiled through the runtime's driver entry point (the library does not link libcuda)
bool tma_available();
static tmap_encode_fn tmap_encoder()
{
    static tmap_encode_fn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<tmap_encode_fn>(p);
    }
    return fn;
}

bool tma_available() { return tmap_encoder() != nullptr; }

// Tensor of complex doubles viewed as doubles: inner extent 2*cols (valid columns), `rows` rows `row_pitch` complex
// apart, `slabs` slabs `slab_pitch` complex apart; box = {2*C doubles, box_rows, box_slabs}.
int make_tile_map(CUtensorMap *map, const void *base, long long cols, long long rows, long long row_pitch, long long slabs,
                         long long slab_pitch, int C, int box_rows, int box_slabs)
{
    tmap_encode_fn enc = tmap_encoder();
    if (!enc) {
        set_error("TMA: cuTensorMapEncodeTiled is not available from this driver");
        return 1;
    }
    const cuuint64_t dims[3] = {(cuuint64_t)(2 * cols), (cuuint64_t)rows, (cuuint64_t)slabs};
    const cuuint64_t strides[2] = {(cuuint64_t)row_pitch * 16, (cuuint64_t)slab_pitch * 16};
    const cuuint32_t box[3] = {(cuuint32_t)(2 * C), (cuuint32_t)box_rows, (cuuint32_t)box_slabs};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<void *>(base), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("TMA: cuTensorMapEncodeTiled failed with %d (cols %lld rows %lld pitch %lld)", (int)r, cols, rows, row_pitch);
        return 1;
    }
    return 0;
}

template <class PL>
__global__ void __launch_bounds__(PL::THREADS, PL::TILE == 4096 ? 2 : 1) fftx_power_kernel(const __grid_constant__ FftxArgs A)
{
    constexpr int N = PL::N, C = PL::C, TILE_MODES = PL::TILE, CTA_THREADS = PL::THREADS;
    constexpr int R2 = PL::R2, R3 = PL::R3;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    cd *const stage = reinterpret_cast<cd *>(smem_raw);                                 // [N][C], the next tile
    cd *const E = reinterpret_cast<cd *>(smem_raw + (size_t)TILE_MODES * 16);           // [N/2][C] complex exchange
    double *const P = reinterpret_cast<double *>(E);                                    // [N][C] |X|^2, same bytes
    double *const sP = reinterpret_cast<double *>(smem_raw + (size_t)TILE_MODES * 24);  // [nrbins][hists]
    unsigned *const sT = reinterpret_cast<unsigned *>(sP + (size_t)A.nrbins * A.hists); // nrbins + 1
    float *const sW = reinterpret_cast<float *>(sT + A.nrbins + 1);                     // dims/2 + 1

    const int tid = threadIdx.x;
    const int c = tid % C, t = tid / C;
    for (int i = tid; i < A.nrbins * A.hists; i += CTA_THREADS)
        sP[i] = 0.0;
    for (int i = tid; i <= A.nrbins; i += CTA_THREADS)
        sT[i] = A.thresh[i];
    for (int i = tid; i <= N / 2; i += CTA_THREADS)
        sW[i] = A.iw1d[i];

    // per-thread bases of the exchange maps
    const int b_ex1r = PL::ex1_r_base(t);
    int b_ex2w[4], b_ex2r[4];
#pragma unroll
    for (int s = 0; s < 4; s++) {
        b_ex2w[s] = PL::ex2_w_base(t, s);
        b_ex2r[s] = PL::ex2_r_base(t, s);
    }
    const int kb = PL::out_k_base(t);
    double *const hist = sP + (c & (A.hists - 1));                   // neighbouring kz columns: different histograms

    // a thread copies exactly the 16 staging slots it later reads: no barrier guards the staging buffer
    // tile -> (column group g, mid row m), advanced without divisions: tiles go up by gridDim.x
    const int step_g = (int)(gridDim.x % (unsigned)A.groups), step_m = (int)(gridDim.x / (unsigned)A.groups);
    auto advance = [&](int &g, int &m) {
        g += step_g;
        m += step_m;
        if (g >= A.groups) {
            g -= A.groups;
            m++;
        }
    };
    __shared__ __align__(8) unsigned long long tile_bar;
    __shared__ __align__(128) double zero_block[ZERO_BOX_ROWS * C * 2];
    if (A.use_tma && tid == 0) {
        mbar_init(&tile_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (A.zero_after) {
        for (int i = tid; i < ZERO_BOX_ROWS * C * 2; i += CTA_THREADS)
            zero_block[i] = 0.0;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the copy engine reads what these stores wrote
    }
    __syncthreads();
    unsigned tile_parity = 0;
    auto issue = [&](int g, int m) {
        if (A.use_tma) {
            // one thread, N/256 bulk tensor copies: x runs along the third tensor dimension
            if (tid == 0 && m < A.n_mid) {
                mbar_expect_tx(&tile_bar, (unsigned)(TILE_MODES * 16));
#pragma unroll 1
                for (int x = 0; x < N; x += TMA_BOX_ROWS)
                    tma_load_3d(stage + (size_t)x * C, &A.tmap, 2 * g * C, m, x, &tile_bar);
            }
            return;
        }
        if (m < A.n_mid) {
            const int kz = g * C + c;
            if (kz < A.nc) {
                const double2 *src = A.spec + (size_t)m * A.row_pitch + kz;
#pragma unroll
                for (int i = 0; i < EPT; i++) {
                    const int n = PL::load_n(t, i);
                    cp_async16(stage + n * C + c, src + (size_t)n * A.x_stride);
                }
            }
        }
        cp_async_commit_group();
    };

    int g = (int)(blockIdx.x % (unsigned)A.groups), m = (int)(blockIdx.x / (unsigned)A.groups);
    int g_next = g, m_next = m;
    issue(g, m);
    for (; m < A.n_mid; g = g_next, m = m_next) {
        advance(g_next, m_next);
        const int kz = g * C + c;
        const bool valid = kz < A.nc;
        int kj = A.mid0 + m;
        kj = kj <= A.dims / 2 ? kj : kj - A.dims;                    // KVAL, powerspectrum.c:33

        cd v[EPT], w[EPT];
        if (A.use_tma) {
            mbar_wait(&tile_bar, tile_parity);
            tile_parity ^= 1u;
            // the tile has left global memory: zeros go back in its place, N/32 bulk stores of 32 rows spread over
            // the warps (each from the same block of zeros; nothing ever waits on them before the kernel ends)
            if (A.zero_after && (tid & 31) == 0) {
                constexpr int PER_WARP = N / ZERO_BOX_ROWS / (CTA_THREADS / 32);
                static_assert(PER_WARP >= 1 && PER_WARP * (CTA_THREADS / 32) * ZERO_BOX_ROWS == N, "zero stores do not tile the column");
#pragma unroll
                for (int j = 0; j < PER_WARP; j++)
                    tma_store_3d(&A.zmap, 2 * g * C, m, ((tid >> 5) * PER_WARP + j) * ZERO_BOX_ROWS, zero_block);
                tma_store_commit();
            }
        } else {
            cp_async_wait_all();
        }
#pragma unroll
        for (int i = 0; i < EPT; i++)
            v[i] = valid ? stage[PL::load_n(t, i) * C + c] : make_double2(0.0, 0.0);
        loads_landed(v);
        if (!A.use_tma)
            issue(g_next, m_next);                                   // the slots are free: the next tile has the whole tile time to arrive
        PL::pass1(v, t, A.tw);

        __syncthreads();                                             // the previous tile's bin walk has left P
        if (A.use_tma)
            issue(g_next, m_next);                                   // (every thread has taken its part of the staged tile)
        exchange<N, C, false, PL::ONE_UNIT2>(E, v, w, c, [&](int i) { return t + PL::ex1_w_part(i); },
                       [&](int i) { return b_ex1r + PL::ex1_r_part(i); });
        PL::pass2(w, t, A.tw);
        __syncthreads();
        exchange<N, C, PL::ONE_UNIT2, false>(E, w, v, c, [&](int i) { return b_ex2w[(i % R2) & 3] + PL::ex2_w_part(i); },
                       [&](int i) { return b_ex2r[(i % R3) & 3] + PL::ex2_r_part(i); });
        PL::pass3(v);
        __syncthreads();
#pragma unroll
        for (int i = 0; i < EPT; i++)
            P[PL::slot(kb + PL::out_k_part(i)) * C + c] = fma(v[i].x, v[i].x, v[i].y * v[i].y);
        __syncthreads();
        if (valid)
            bin_walk<PL>(P, t, c, kj, kz, A.dims / 2, sW, sT, A.nrbins, A.half_bpu, hist, A.hists);
    }
    cp_async_wait_all();
    if (A.zero_after && (tid & 31) == 0)
        tma_store_wait_all();
    __syncthreads();
    for (int i = tid; i < A.nrbins; i += CTA_THREADS) {
        double sum = 0.0;
        for (int j = 0; j < A.hists; j++)
            sum += sP[i * A.hists + j];
        if (sum != 0.0)
            atomicAdd(&A.sums[i], sum);
    }
}

// ---------------------------------------------------------------------------------
// The same pass with the CTA split in two halves of 256 threads that share the 8192-mode tile (128-byte rows from
// HBM) but work on four of its eight columns each, with their own exchange buffers and named barriers -- so that one
// half's shared-memory exchanges can run under the other half's FP64 passes (GENPK_OPT_FUSED_XPASS = 3, 1024 only).
// The halves are started half a tile apart; the tile buffer is refilled when both have taken their columns
// (`empty`, 512 arrivals), so they stay within one tile of each other.
// ---------------------------------------------------------------------------------
template <class PL>
__global__ void __launch_bounds__(2 * PL::THREADS, 1) fftx_power_halves_kernel(const __grid_constant__ FftxArgs A)
{
    constexpr int N = PL::N, C = PL::C, CW = 2 * PL::C, HT = PL::THREADS, TILE_MODES = 2 * PL::TILE;
    constexpr int R2 = PL::R2, R3 = PL::R3;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    cd *const stage = reinterpret_cast<cd *>(smem_raw);                                 // [N][CW], the next tile
    const int tid = threadIdx.x, half = tid / HT, ht = tid % HT;
    cd *const E = reinterpret_cast<cd *>(smem_raw + (size_t)TILE_MODES * 16) + (size_t)half * (N / 2) * C;   // [N/2][C] per half
    double *const P = reinterpret_cast<double *>(E);                                    // [N][C] |X|^2, same bytes
    double *const sP = reinterpret_cast<double *>(smem_raw + (size_t)TILE_MODES * 24);  // [nrbins][2]: one histogram per half
    unsigned *const sT = reinterpret_cast<unsigned *>(sP + (size_t)A.nrbins * 2);       // nrbins + 1
    float *const sW = reinterpret_cast<float *>(sT + A.nrbins + 1);                     // dims/2 + 1
    __shared__ __align__(8) unsigned long long full_bar, empty_bar;

    const int c = ht % C, t = ht / C, col = half * C + c;
    for (int i = tid; i < A.nrbins * 2; i += 2 * HT)
        sP[i] = 0.0;
    for (int i = tid; i <= A.nrbins; i += 2 * HT)
        sT[i] = A.thresh[i];
    for (int i = tid; i <= N / 2; i += 2 * HT)
        sW[i] = A.iw1d[i];
    if (tid == 0) {
        mbar_init(&full_bar, 1);
        mbar_init(&empty_bar, 2 * HT);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const NamedSync hsync{1 + half, HT};
    const int b_ex1r = PL::ex1_r_base(t);
    int b_ex2w[4], b_ex2r[4];
#pragma unroll
    for (int s = 0; s < 4; s++) {
        b_ex2w[s] = PL::ex2_w_base(t, s);
        b_ex2r[s] = PL::ex2_r_base(t, s);
    }
    const int kb = PL::out_k_base(t);
    double *const hist = sP + half;

    const int step_g = (int)(gridDim.x % (unsigned)A.groups), step_m = (int)(gridDim.x / (unsigned)A.groups);
    auto advance = [&](int &g, int &m) {
        g += step_g;
        m += step_m;
        if (g >= A.groups) {
            g -= A.groups;
            m++;
        }
    };
    auto issue = [&](int g, int m) {                                 // one thread
        if (m < A.n_mid) {
            mbar_expect_tx(&full_bar, (unsigned)(TILE_MODES * 16));
#pragma unroll 1
            for (int x = 0; x < N; x += TMA_BOX_ROWS)
                tma_load_3d(stage + (size_t)x * CW, &A.tmap, 2 * g * CW, m, x, &full_bar);
        }
    };
    int g = (int)(blockIdx.x % (unsigned)A.groups), m = (int)(blockIdx.x / (unsigned)A.groups);
    int g_next = g, m_next = m;
    unsigned parity = 0;
    if (tid == 0)
        issue(g, m);
    if (half == 1)
        __nanosleep(A.halves_delay_ns);                              // half a tile behind from the start
    for (; m < A.n_mid; g = g_next, m = m_next) {
        advance(g_next, m_next);
        const int kz = g * CW + col;
        const bool valid = kz < A.nc;
        int kj = A.mid0 + m;
        kj = kj <= A.dims / 2 ? kj : kj - A.dims;                    // KVAL, powerspectrum.c:33
        cd v[EPT], w[EPT];
        mbar_wait(&full_bar, parity);
#pragma unroll
        for (int i = 0; i < EPT; i++)
            v[i] = valid ? stage[PL::load_n(t, i) * CW + col] : make_double2(0.0, 0.0);
        loads_landed(v);
        mbar_arrive(&empty_bar);                                     // this thread's columns have left the tile buffer
        // The next tile is asked for by the half that runs behind: when it has taken its columns the leading half
        // has long taken its own, so nobody waits here -- and the leading half never waits for the follower.
        if (tid == HT) {
            mbar_wait(&empty_bar, parity);
            issue(g_next, m_next);
        }
        parity ^= 1u;
        PL::pass1(v, t, A.tw);
        hsync();                                                     // this half's previous bin walk has left P
        exchange<N, C, false, PL::ONE_UNIT2>(E, v, w, c, [&](int i) { return t + PL::ex1_w_part(i); },
                       [&](int i) { return b_ex1r + PL::ex1_r_part(i); }, hsync);
        PL::pass2(w, t, A.tw);
        hsync();
        exchange<N, C, PL::ONE_UNIT2, false>(E, w, v, c, [&](int i) { return b_ex2w[(i % R2) & 3] + PL::ex2_w_part(i); },
                       [&](int i) { return b_ex2r[(i % R3) & 3] + PL::ex2_r_part(i); }, hsync);
        PL::pass3(v);
        hsync();
#pragma unroll
        for (int i = 0; i < EPT; i++)
            P[PL::slot(kb + PL::out_k_part(i)) * C + c] = fma(v[i].x, v[i].x, v[i].y * v[i].y);
        hsync();
        if (valid)
            bin_walk<PL>(P, t, c, kj, kz, A.dims / 2, sW, sT, A.nrbins, A.half_bpu, hist, 2);
    }
    __syncthreads();
    for (int i = tid; i < A.nrbins; i += 2 * HT) {
        const double sum = sP[2 * i] + sP[2 * i + 1];
        if (sum != 0.0)
            atomicAdd(&A.sums[i], sum);
    }
}

// ---------------------------------------------------------------------------------
// The x pass with the two-pass plan (fftx_core.cuh: Plan2): 32 modes per thread, a radix-32 register pass, ONE
// exchange (through the tile buffer), a radix-R2 register pass -- half the threads of fftx_power_kernel with twice
// the registers, a third less shared-memory traffic and half the barriers per tile (GENPK_OPT_FUSED_XPASS = 4).
// ---------------------------------------------------------------------------------
__device__ __forceinline__ void loads_landed32(const fftx::cd *v)
{
#pragma unroll
    for (int i = 0; i < 32; i += 4)
        asm volatile("" ::"d"(v[i].x), "d"(v[i].y), "d"(v[i + 1].x), "d"(v[i + 1].y), "d"(v[i + 2].x), "d"(v[i + 2].y),
                     "d"(v[i + 3].x), "d"(v[i + 3].y)
                     : "memory");
}

template <class PL>
__global__ void __launch_bounds__(PL::THREADS, 1) fftx_power2_kernel(const __grid_constant__ FftxArgs A)
{
    constexpr int N = PL::N, C = PL::C, TILE_MODES = PL::TILE, CTA_THREADS = PL::THREADS, E2 = PL::EPT2;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    cd *const stage = reinterpret_cast<cd *>(smem_raw);                                 // [N][C], the next tile
    cd *const E = reinterpret_cast<cd *>(smem_raw + (size_t)TILE_MODES * 16);           // [N/2][C] complex exchange
    double *const P = reinterpret_cast<double *>(E);                                    // [N][C] |X|^2, same bytes
    double *const sP = reinterpret_cast<double *>(smem_raw + (size_t)TILE_MODES * 24);  // [nrbins][hists]
    unsigned *const sT = reinterpret_cast<unsigned *>(sP + (size_t)A.nrbins * A.hists); // nrbins + 1
    float *const sW = reinterpret_cast<float *>(sT + A.nrbins + 1);                     // dims/2 + 1
    __shared__ __align__(8) unsigned long long tile_bar;

    const int tid = threadIdx.x;
    const int c = tid % C, t = tid / C;
    for (int i = tid; i < A.nrbins * A.hists; i += CTA_THREADS)
        sP[i] = 0.0;
    for (int i = tid; i <= A.nrbins; i += CTA_THREADS)
        sT[i] = A.thresh[i];
    for (int i = tid; i <= N / 2; i += CTA_THREADS)
        sW[i] = A.iw1d[i];
    if (tid == 0) {
        mbar_init(&tile_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    double *const hist = sP + (c & (A.hists - 1));
    const int step_g = (int)(gridDim.x % (unsigned)A.groups), step_m = (int)(gridDim.x / (unsigned)A.groups);
    auto advance = [&](int &g, int &m) {
        g += step_g;
        m += step_m;
        if (g >= A.groups) {
            g -= A.groups;
            m++;
        }
    };
    auto issue = [&](int g, int m) {
        if (tid == 0 && m < A.n_mid) {
            mbar_expect_tx(&tile_bar, (unsigned)(TILE_MODES * 16));
#pragma unroll 1
            for (int x = 0; x < N; x += TMA_BOX_ROWS)
                tma_load_3d(stage + (size_t)x * C, &A.tmap, 2 * g * C, m, x, &tile_bar);
        }
    };
    int g = (int)(blockIdx.x % (unsigned)A.groups), m = (int)(blockIdx.x / (unsigned)A.groups);
    int g_next = g, m_next = m;
    unsigned parity = 0;
    issue(g, m);
    for (; m < A.n_mid; g = g_next, m = m_next) {
        advance(g_next, m_next);
        const int kz = g * C + c;
        const bool valid = kz < A.nc;
        int kj = A.mid0 + m;
        kj = kj <= A.dims / 2 ? kj : kj - A.dims;                    // KVAL, powerspectrum.c:33
        cd v[E2];
        mbar_wait(&tile_bar, parity);
        parity ^= 1u;
#pragma unroll
        for (int i = 0; i < E2; i++)
            v[i] = valid ? stage[PL::load_n(t, i) * C + c] : make_double2(0.0, 0.0);
        PL::pass1(v, t, A.tw);
        __syncthreads();                                             // everyone has taken its part of the tile
        // The exchange goes through the tile buffer itself, all 32 registers at once: with 32 modes per thread there
        // are no registers for a second copy, which an exchange through a half-size buffer needs (a thread would
        // receive while it still holds what it has not sent).  The next tile is asked for once the buffer has been
        // read back; its fill runs under the second pass and the bin walk.
#pragma unroll
        for (int i = 0; i < E2; i++)
            stage[PL::ex_w(t, i) * C + c] = v[i];
        __syncthreads();
#pragma unroll
        for (int i = 0; i < E2; i++)
            v[i] = stage[PL::ex_r(t, i) * C + c];
        loads_landed32(v);
        __syncthreads();
        if (tid == 0)
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        issue(g_next, m_next);
        PL::pass2(v);
#pragma unroll
        for (int i = 0; i < E2; i++)
            P[PL::slot(PL::out_k(t, i)) * C + c] = fma(v[i].x, v[i].x, v[i].y * v[i].y);
        __syncthreads();
        if (valid) {
            bin_walk<PL>(P, 2 * t, c, kj, kz, A.dims / 2, sW, sT, A.nrbins, A.half_bpu, hist, A.hists);
            bin_walk<PL>(P, 2 * t + 1, c, kj, kz, A.dims / 2, sW, sT, A.nrbins, A.half_bpu, hist, A.hists);
        }
    }
    __syncthreads();
    for (int i = tid; i < A.nrbins; i += CTA_THREADS) {
        double sum = 0.0;
        for (int j = 0; j < A.hists; j++)
            sum += sP[i * A.hists + j];
        if (sum != 0.0)
            atomicAdd(&A.sums[i], sum);
    }
}

// ---------------------------------------------------------------------------------
// The same tile machinery as a plain in-place column transform: length-N FFTs along the
// MIDDLE axis (y) of [n_planes][N][nc] complex planes, i.e. the y pass of the (y,z)
// transform after cuFFT's batched 1-D r2c along z.  Reads and writes every mode once in
// rows of C*16 bytes (cuFFT's strided pass reaches about half the copy bandwidth here).
// ---------------------------------------------------------------------------------
struct FftColsArgs {
    alignas(64) CUtensorMap tmap;   // the planes as {2*nc doubles, N rows, n_planes slabs}; box {2C, 256, 1}
    int use_tma;
    double2 *spec;            // [n_planes][N][nc], transformed in place
    const double2 *tw;
    int nc;
    long long plane_stride;   // modes per plane: N * nc
    int groups;               // column groups per plane: ceil(nc / C)
    int n_planes;
    long long n_tiles;        // n_planes * groups
    // SCATTER (multi-GPU transpose fused into the pass): row ky of local plane o is written to
    // rank ky / ny, into its [dims][ny][nc] block at plane x0 + o, row ky % ny -- a peer store
    // over NVLink (or a local store for the rank's own rows) instead of an in-place store.
    double2 *peer[GENPK_MAX_PEERS];
    int ny_shift;             // log2(ny)
    int dst_pitch;            // modes between consecutive rows of a transposed block
    int x0;                   // first global x plane of this rank
};

template <class PL, bool SCATTER>
__global__ void __launch_bounds__(PL::THREADS, PL::TILE == 4096 ? 2 : 1) fft_cols_kernel(const __grid_constant__ FftColsArgs A)
{
    constexpr int N = PL::N, C = PL::C, TILE_MODES = PL::TILE;
    constexpr int R2 = PL::R2, R3 = PL::R3;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    cd *const stage = reinterpret_cast<cd *>(smem_raw);                                 // [N][C], the next tile
    cd *const E = reinterpret_cast<cd *>(smem_raw + (size_t)TILE_MODES * 16);           // [N/2][C] complex exchange
    const int tid = threadIdx.x;
    const int c = tid % C, t = tid / C;
    const int b_ex1r = PL::ex1_r_base(t);
    int b_ex2w[4], b_ex2r[4];
#pragma unroll
    for (int s = 0; s < 4; s++) {
        b_ex2w[s] = PL::ex2_w_base(t, s);
        b_ex2r[s] = PL::ex2_r_base(t, s);
    }
    const int kb = PL::out_k_base(t);

    // tile -> (column group g, plane o), advanced without divisions: tiles go up by gridDim.x
    const int step_g = (int)(gridDim.x % (unsigned)A.groups), step_o = (int)(gridDim.x / (unsigned)A.groups);
    auto advance = [&](int &g, int &o) {
        g += step_g;
        o += step_o;
        if (g >= A.groups) {
            g -= A.groups;
            o++;
        }
    };
    __shared__ __align__(8) unsigned long long tile_bar;
    if (A.use_tma && tid == 0) {
        mbar_init(&tile_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    unsigned tile_parity = 0;
    auto issue = [&](int g, int o) {
        if (A.use_tma) {
            // one thread, N/256 bulk tensor copies of 256 rows x C*16 bytes each
            if (tid == 0 && o < A.n_planes) {
                mbar_expect_tx(&tile_bar, (unsigned)(TILE_MODES * 16));
#pragma unroll 1
                for (int y = 0; y < N; y += TMA_BOX_ROWS)
                    tma_load_3d(stage + (size_t)y * C, &A.tmap, 2 * g * C, y, o, &tile_bar);
            }
            return;
        }
        if (o < A.n_planes) {
            const int kz = g * C + c;
            if (kz < A.nc) {
                const double2 *src = A.spec + (size_t)o * A.plane_stride + kz;
#pragma unroll
                for (int i = 0; i < EPT; i++) {
                    const int n = PL::load_n(t, i);
                    cp_async16(stage + n * C + c, src + (size_t)n * A.nc);
                }
            }
        }
        cp_async_commit_group();
    };

    int g = (int)(blockIdx.x % (unsigned)A.groups), o = (int)(blockIdx.x / (unsigned)A.groups);
    int g_next = g, o_next = o;
    issue(g, o);
    for (; o < A.n_planes; g = g_next, o = o_next) {
        advance(g_next, o_next);
        const int kz = g * C + c;
        const bool valid = kz < A.nc;
        cd v[EPT], w[EPT];
        if (A.use_tma) {
            mbar_wait(&tile_bar, tile_parity);
            tile_parity ^= 1u;
        } else {
            cp_async_wait_all();
        }
#pragma unroll
        for (int i = 0; i < EPT; i++)
            v[i] = valid ? stage[PL::load_n(t, i) * C + c] : make_double2(0.0, 0.0);
        PL::pass1(v, t, A.tw);
        // (issued after pass 1 here: right after the fill these loads would queue up behind the
        // previous tile's store burst -- measured 3.9 -> 5.3 ms, profiles/r01/s29)
        if (!A.use_tma)
            issue(g_next, o_next);                                   // another tile: never the rows written below
        __syncthreads();                                             // the previous tile's last exchange read is over
        if (A.use_tma)
            issue(g_next, o_next);                                   // (every thread has taken its part of the staged tile)
        exchange<N, C, false, PL::ONE_UNIT2>(E, v, w, c, [&](int i) { return t + PL::ex1_w_part(i); },
                       [&](int i) { return b_ex1r + PL::ex1_r_part(i); });
        PL::pass2(w, t, A.tw);
        __syncthreads();
        exchange<N, C, PL::ONE_UNIT2, false>(E, w, v, c, [&](int i) { return b_ex2w[(i % R2) & 3] + PL::ex2_w_part(i); },
                       [&](int i) { return b_ex2r[(i % R3) & 3] + PL::ex2_r_part(i); });
        PL::pass3(v);
        if (valid) {
            if (SCATTER) {
                const int ny = 1 << A.ny_shift;
                const size_t plane_off = (size_t)(A.x0 + o) * ny * A.dst_pitch + kz;
#pragma unroll
                for (int i = 0; i < EPT; i++) {
                    const int ky = kb + PL::out_k_part(i);
                    A.peer[ky >> A.ny_shift][plane_off + (size_t)(ky & (ny - 1)) * A.dst_pitch] = v[i];
                }
            } else {
                double2 *dst = A.spec + (size_t)o * A.plane_stride + kz;
#pragma unroll
                for (int i = 0; i < EPT; i++)
                    dst[(size_t)(kb + PL::out_k_part(i)) * A.nc] = v[i];
            }
        }
    }
    cp_async_wait_all();
}

// 8192-mode tiles (one CTA of 512 threads per SM) keep a tile row at 128 B for 1024 and 64 B for
// 2048; the smaller grids use 4096-mode tiles, two CTAs per SM.  fused_xpass == 2 asks for the
// 4096-mode tile at 1024 as well (64-B rows: measured slower, kept for A/B measurements).
static int fftx_tile_modes(const genpk_ctx *ctx)
{
    const int dims = ctx->g.dims;
    return (dims == 2048 || (dims == 1024 && ctx->fused_xpass != 2)) ? 8192 : 4096;
}

static size_t fftx_smem_bytes_h(const genpk_ctx *ctx, int nrbins, int hists)
{
    const int dims = ctx->g.dims;
    const size_t tile = (size_t)fftx_tile_modes(ctx);
    // staging tile + half-size complex exchange buffer + histograms + threshold and window tables
    return tile * 16 + tile * 8 + (size_t)nrbins * 8 * hists + (size_t)(nrbins + 1) * 4 + (size_t)(dims / 2 + 1) * 4 + 16;
}

// the kernel's static shared memory: the block of zeros behind the zero stores, the tile barrier
constexpr size_t FFTX_STATIC_SMEM = ZERO_BOX_ROWS * 8 * 16 + 128;

// two histograms (even / odd tile columns) when they fit
static int fftx_hists(const genpk_ctx *ctx, int nrbins)
{
    return fftx_smem_bytes_h(ctx, nrbins, 2) + FFTX_STATIC_SMEM <= (size_t)ctx->smem_optin ? 2 : 1;
}

size_t fftx_smem_bytes(const genpk_ctx *ctx, int nrbins) { return fftx_smem_bytes_h(ctx, nrbins, fftx_hists(ctx, nrbins)); }

bool fftx_supported(const genpk_ctx *ctx, int nrbins)
{
    const int d = ctx->g.dims;
    if (ctx->fused_xpass == 0)
        return false;
    if (d != 256 && d != 512 && d != 1024 && d != 2048)
        return false;
    return nrbins >= 1 && fftx_smem_bytes(ctx, nrbins) + FFTX_STATIC_SMEM <= (size_t)ctx->smem_optin;
}

// Rows of the library-owned transposed block start on 128-byte boundaries (nc = dims/2+1 is odd:
// with the natural pitch every other row would straddle sectors, and 128-byte peer stores
// would split into two NVLink packets).
int recv_row_pitch(const genpk_ctx *ctx) { return (ctx->g.nc + 7) / 8 * 8; }

int ensure_twiddles(genpk_ctx *ctx)
{
    const int d = ctx->g.dims;
    if (ctx->d_twiddle && ctx->twiddle_n == d)
        return 0;
    if (ctx->d_twiddle) cudaFree(ctx->d_twiddle);
    ctx->d_twiddle = nullptr;
    // [0, d): exp(-2 pi i t / d);  [d, d + d/2): exp(-2 pi i t / (d/2)), the table of the half-length row transforms
    std::vector<double> h(2 * (size_t)(d + d / 2));
    const long double tau = 6.283185307179586476925286766559005768L;
    for (int t = 0; t < d; t++) {
        // exact values on the axes and diagonals, long-double libm elsewhere
        const long double a = tau * (long double)t / (long double)d;
        double cr = (double)cosl(a), si = (double)sinl(a);
        if (4 * t % d == 0) {
            const int q = 4 * t / d;                       // multiples of pi/2
            cr = q == 0 ? 1.0 : (q == 2 ? -1.0 : 0.0);
            si = q == 1 ? 1.0 : (q == 3 ? -1.0 : 0.0);
        }
        h[2 * (size_t)t] = cr;
        h[2 * (size_t)t + 1] = -si;
    }
    for (int t = 0; t < d / 2; t++) {
        h[2 * (size_t)(d + t)] = h[2 * (size_t)(2 * t)];
        h[2 * (size_t)(d + t) + 1] = h[2 * (size_t)(2 * t) + 1];
    }
    GENPK_CUDA_OK(cudaMalloc(&ctx->d_twiddle, h.size() * sizeof(double)));
    GENPK_CUDA_OK(cudaMemcpyAsync(ctx->d_twiddle, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    GENPK_CUDA_OK(cudaStreamSynchronize(ctx->stream));
    ctx->twiddle_n = d;
    return 0;
}

template <class PL> static int launch_fftx(genpk_ctx *ctx, const FftxArgs &A, size_t smem)
{
    auto kern = fftx_power_kernel<PL>;
    GENPK_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    GENPK_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, PL::THREADS, smem));
    if (per_sm < 1) {
        set_error("fused x pass: %zu bytes of shared memory do not fit", smem);
        return 1;
    }
    long long ctas = (long long)ctx->sm_count * per_sm;          // persistent: every CTA resident
    if (ctas > A.n_tiles) ctas = A.n_tiles;
    if (ctas < 1) ctas = 1;
    kern<<<(int)ctas, PL::THREADS, smem, ctx->stream>>>(A);
    ctx->launches++;
    GENPK_CUDA_OK(cudaGetLastError());
    return 0;
}

// P sums of the block [dims][n_mid][nc] of a (y,z)-transformed spectrum whose first mid
// row is global ky index mid0, with the x transform done on the fly.  sums_dev: 3*nrbins
// doubles (P from this pass, K and N from the cached geometry pass).
int fftx_power_raw(genpk_ctx *ctx, const double *spec_yz, int n_mid, int mid0, int nrbins, double *sums_dev, int row_pitch,
                   bool *zero_after)
{
    const bool want_zero = zero_after && *zero_after;
    if (zero_after) *zero_after = false;
    if (!fftx_supported(ctx, nrbins)) {
        set_error("fused x pass: unsupported grid side %d / nrbins %d", ctx->g.dims, nrbins);
        return 1;
    }
    if (int rc = ensure_tables(ctx, nrbins)) return rc;
    if (int rc = ensure_twiddles(ctx)) return rc;
    if (int rc = power_seed_sums(ctx, ctx->g.dims, 0, n_mid, mid0, nrbins, sums_dev)) return rc;
    FftxArgs A;
    A.spec = reinterpret_cast<const double2 *>(spec_yz);
    A.tw = reinterpret_cast<const double2 *>(ctx->d_twiddle);
    A.dims = ctx->g.dims;
    A.nc = ctx->g.nc;
    A.n_mid = n_mid;
    A.mid0 = mid0;
    A.row_pitch = row_pitch > 0 ? row_pitch : A.nc;
    A.x_stride = (long long)n_mid * A.row_pitch;
    A.nrbins = nrbins;
    A.iw1d = ctx->d_iw1d;
    A.thresh = ctx->d_thresh;
    A.half_bpu = nrbins > 1 ? (float)(0.5 * (nrbins - 1) / log(sqrt(3.0) * A.dims / 2.0)) : 0.f;
    A.sums = sums_dev;
    A.hists = fftx_hists(ctx, nrbins);
    const size_t smem = fftx_smem_bytes(ctx, nrbins);
    A.use_tma = 0;
    A.zero_after = 0;
    int tma_rc = 0;
    auto tiles = [&](int C) {
        A.groups = (A.nc + C - 1) / C;
        A.n_tiles = (long long)n_mid * A.groups;
        if (ctx->use_tma && tmap_encoder()) {
            // {2*nc doubles} x {n_mid rows, row_pitch apart} x {dims values of x, x_stride apart}; box {2C, 1, 256}
            tma_rc = make_tile_map(&A.tmap, spec_yz, A.nc, n_mid, A.row_pitch, A.dims, A.x_stride, C, 1, TMA_BOX_ROWS);
            A.use_tma = tma_rc == 0 ? 1 : 0;
            // zero stores behind the read (the caller's block is every row of the tensor: rows of 2*nc doubles,
            // which in the padded grid layout is the whole row)
            if (A.use_tma && want_zero &&
                make_tile_map(&A.zmap, spec_yz, A.nc, n_mid, A.row_pitch, A.dims, A.x_stride, C, 1, ZERO_BOX_ROWS) == 0) {
                A.zero_after = 1;
                *zero_after = true;
            }
        }
    };
    // (tests/fftx_emu.cpp instantiates the same plans on the host)
    typedef Plan<4, 8, 8, 4096> P256;
    typedef Plan<8, 8, 8, 4096> P512;
    typedef Plan<16, 8, 8, 4096> P1024;
    typedef Plan<16, 16, 8, 8192> P2048;
    typedef Plan<16, 8, 8, 8192> P1024W;
    if (A.dims == 1024 && ctx->fused_xpass == 4 && ctx->use_tma && tmap_encoder()) {
        typedef Plan2<32, 8192> Q1024;
        tiles(Q1024::C);
        if (A.use_tma) {
            A.zero_after = 0;
            if (zero_after) *zero_after = false;
            auto kern = fftx_power2_kernel<Q1024>;
            GENPK_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            long long ctas = ctx->sm_count;
            if (ctas > A.n_tiles) ctas = A.n_tiles;
            kern<<<(int)ctas, Q1024::THREADS, smem, ctx->stream>>>(A);
            ctx->launches++;
            GENPK_CUDA_OK(cudaGetLastError());
            return 0;
        }
    }
    if (A.dims == 1024 && ctx->fused_xpass == 3 && ctx->use_tma && tmap_encoder()) {
        // two halves of 256 threads on one 8192-mode tile (see fftx_power_halves_kernel)
        tiles(2 * P1024::C);
        if (A.use_tma) {
            A.zero_after = 0;
            if (zero_after) *zero_after = false;
            auto kern = fftx_power_halves_kernel<P1024>;
            const char *dl = getenv("GENPK_XHALVES_DELAY_NS");
            A.halves_delay_ns = dl ? atoi(dl) : 3000;
            const size_t smem2 = (size_t)8192 * 24 + (size_t)nrbins * 16 + (size_t)(nrbins + 1) * 4 + (size_t)(A.dims / 2 + 1) * 4 + 16;
            if (smem2 + 256 <= (size_t)ctx->smem_optin) {
                GENPK_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
                long long ctas = ctx->sm_count;
                if (ctas > A.n_tiles) ctas = A.n_tiles;
                kern<<<(int)ctas, 2 * P1024::THREADS, smem2, ctx->stream>>>(A);
                ctx->launches++;
                GENPK_CUDA_OK(cudaGetLastError());
                return 0;
            }
        }
    }
    if (A.dims == 1024 && fftx_tile_modes(ctx) == 8192) {
        tiles(P1024W::C);
        return launch_fftx<P1024W>(ctx, A, smem);
    }
    switch (A.dims) {
    case 256: tiles(P256::C); return launch_fftx<P256>(ctx, A, smem);
    case 512: tiles(P512::C); return launch_fftx<P512>(ctx, A, smem);
    case 1024: tiles(P1024::C); return launch_fftx<P1024>(ctx, A, smem);
    case 2048: tiles(P2048::C); return launch_fftx<P2048>(ctx, A, smem);
    }
    return 1;
}

template <class PL, bool SCATTER> static int launch_cols(genpk_ctx *ctx, FftColsArgs &A, int n_planes)
{
    auto kern = fft_cols_kernel<PL, SCATTER>;
    const size_t smem = (size_t)PL::TILE * 24;                   // staging tile + half-size exchange buffer
    A.groups = (A.nc + PL::C - 1) / PL::C;
    A.n_planes = n_planes;
    A.n_tiles = (long long)n_planes * A.groups;
    A.use_tma = 0;
    if (ctx->use_tma && tmap_encoder())
        A.use_tma = make_tile_map(&A.tmap, A.spec, A.nc, PL::N, A.nc, n_planes, A.plane_stride, PL::C, TMA_BOX_ROWS, 1) == 0 ? 1 : 0;
    GENPK_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    GENPK_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, PL::THREADS, smem));
    if (per_sm < 1) {
        set_error("column FFT: %zu bytes of shared memory do not fit", smem);
        return 1;
    }
    long long ctas = (long long)ctx->sm_count * per_sm;
    if (ctas > A.n_tiles) ctas = A.n_tiles;
    if (ctas < 1) ctas = 1;
    kern<<<(int)ctas, PL::THREADS, smem, ctx->stream>>>(A);
    ctx->launches++;
    GENPK_CUDA_OK(cudaGetLastError());
    return 0;
}

bool fft_cols_supported(const genpk_ctx *ctx)
{
    const int d = ctx->g.dims;
    return ctx->own_ypass != 0 && (d == 256 || d == 512 || d == 1024 || d == 2048) &&
           (size_t)8192 * 24 <= (size_t)ctx->smem_optin;
}

template <bool SCATTER> static int cols_dispatch(genpk_ctx *ctx, FftColsArgs &A, int n_planes)
{
    switch (ctx->g.dims) {
    case 256: return launch_cols<Plan<4, 8, 8, 4096>, SCATTER>(ctx, A, n_planes);
    case 512: return launch_cols<Plan<8, 8, 8, 4096>, SCATTER>(ctx, A, n_planes);
    case 1024: return launch_cols<Plan<16, 8, 8, 8192>, SCATTER>(ctx, A, n_planes);
    case 2048: return launch_cols<Plan<16, 16, 8, 8192>, SCATTER>(ctx, A, n_planes);
    }
    set_error("column FFT: unsupported grid side %d", ctx->g.dims);
    return 1;
}

// In-place FFT along y of n_planes planes [dims][nc] starting at spec.
int fft_cols_y(genpk_ctx *ctx, double *spec, int n_planes)
{
    if (int rc = ensure_twiddles(ctx)) return rc;
    FftColsArgs A = {};
    A.spec = reinterpret_cast<double2 *>(spec);
    A.tw = reinterpret_cast<const double2 *>(ctx->d_twiddle);
    A.nc = ctx->g.nc;
    A.plane_stride = (long long)ctx->g.dims * ctx->g.nc;
    return cols_dispatch<false>(ctx, A, n_planes);
}

// The y pass of this rank's planes with the transpose fused in: results go straight to the
// [dims][ny][nc] blocks of their owner ranks (peer[r], set by genpk_slab_set_peers).
int fft_cols_y_scatter(genpk_ctx *ctx, double *spec, int n_planes)
{
    const SlabGeom &g = ctx->g;
    const int ny = g.dims / g.nranks;
    if (!ctx->peers_set || g.nranks > GENPK_MAX_PEERS || (ny & (ny - 1)) != 0) {
        set_error("scatter y pass: peers not set, more than %d ranks, or dims/nranks not a power of two", GENPK_MAX_PEERS);
        return 1;
    }
    if (int rc = ensure_twiddles(ctx)) return rc;
    FftColsArgs A = {};
    A.spec = reinterpret_cast<double2 *>(spec);
    A.tw = reinterpret_cast<const double2 *>(ctx->d_twiddle);
    A.nc = g.nc;
    A.plane_stride = (long long)g.dims * g.nc;
    for (int r = 0; r < g.nranks; r++)
        A.peer[r] = reinterpret_cast<double2 *>(ctx->peer_recv[r]);
    A.ny_shift = 0;
    while ((1 << A.ny_shift) < ny) A.ny_shift++;
    A.x0 = g.x0;
    A.dst_pitch = recv_row_pitch(ctx);
    return cols_dispatch<true>(ctx, A, n_planes);
}

}  // namespace genpk
