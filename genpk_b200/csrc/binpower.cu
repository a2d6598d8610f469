// Fused |delta_k|^2 binning -- replaces powerspectrum() (powerspectrum.c:35-110)
// and the per-mode invwindow() calls inside it (fieldize.cpp:125-133).
//
// One streaming pass over the half-complex spectrum [outer][mid][kz] (16 B per
// mode, read once): per mode
//     k2   = ki^2 + kj^2 + kz^2                      (exact integer)
//     bin  = max{b : thresh[b] <= k2}                (host-built table == floor(bpu*log|k|))
//     P   += mult * (re1*re2 + im1*im2) * W^4,  W = (float)((iw[ki]*iw[kj])*iw[kz])
//     K   += mult * sqrt(k2);   N += mult;   mult = 1 on the kz=0 and kz=dims/2 planes, else 2
// The DC mode is skipped (powerspectrum.c:63).
//
// Mapping: persistent warps.  A "panel" unit is 32 consecutive kz (one lane
// each, so every load is a coalesced 512 B) times JC consecutive mid steps that
// the thread walks; along that walk |k| changes slowly, so the thread keeps a
// run (bin, P, K, N) in registers and only touches the CTA's shared-memory
// histogram when the bin changes.
//
// Mirror rows: bin, window and multiplicity depend on |ki|, |kj| only, so the up
// to four rows (+-ki, +-kj) of one kz share everything but the data.  A step
// loads all of them (four independent coalesced streams), adds their |delta|^2
// and does the bin / window / run bookkeeping once.  The outer mirror needs the
// whole outer axis in the block (always true: x is complete on every rank after
// the transpose), the mid mirror the whole mid axis (single-rank contexts).
//
// The nc%32 leftover columns (the Nyquist column for power-of-two grids) are
// "tail" units with lanes on rows.  Shared histograms are flushed to global
// with one red.add per non-empty bin per CTA.
//
// K and N depend on the grid geometry only, not on the data.  In the default
// mode they are produced once per (dims, nrbins, block) by the same kernel
// instantiated without loads (GEOM), cached in the context like an FFT plan's
// twiddles, and the per-spectrum pass (DATA) carries P alone; the FUSED
// instantiation does all three in one pass.
#include "common.cuh"

namespace genpk {

struct PowerArgs {
    const double2 *a;
    const double2 *b;
    int dims, nc;
    int n_outer, outer0, n_mid, mid0;
    int mirror_outer, mirror_mid;   // block holds the whole axis: fold +-k rows into one step
    int n_oc, n_mc;       // outer / mid classes (dims/2+1 when mirrored, else the extent)
    int nrbins;
    int jc;               // mid steps per unit
    int n_mchunks;        // ceil(n_mc / jc)
    int n_panels;         // nc / 32
    int n_tail;           // nc % 32
    long long n_units;
    const float *iw1d;
    const uint32_t *thresh;
    float half_bpu;       // first guess of the bin only; the threshold walk makes it exact
    double *sums;         // [3][nrbins]: P, K, N
};

enum PowerKind { PK_FUSED = 0, PK_DATA = 1, PK_GEOM = 2 };

__device__ __forceinline__ int kval(int i, int dims) { return i <= dims / 2 ? i : i - dims; }   // powerspectrum.c:33

struct Run {
    int bin;
    unsigned lo, hi;      // thresh[bin], thresh[bin+1]
    double p, k;
    unsigned n;
};

template <int KIND>
__device__ __forceinline__ void run_flush(const Run &r, int mult, double *sP, double *sK, unsigned *sN)
{
    if (r.n) {
        if (KIND != PK_GEOM)
            atomicAdd(&sP[r.bin], r.p * mult);
        if (KIND != PK_DATA) {
            atomicAdd(&sK[r.bin], r.k * mult);
            atomicAdd(&sN[r.bin], r.n * (unsigned)mult);
        }
    }
}

__device__ __forceinline__ void run_seek(Run &r, unsigned k2, const unsigned *sT, int nrbins, float half_bpu, bool guess)
{
    int b = r.bin;
    if (guess) {
        b = (int)(half_bpu * __logf((float)k2));
        b = max(0, min(b, nrbins - 1));
    }
    while (k2 >= sT[b + 1]) b++;
    while (k2 < sT[b]) b--;
    r.bin = b;
    r.lo = sT[b];
    r.hi = sT[b + 1];
    r.p = 0.0;
    r.k = 0.0;
    r.n = 0;
}

// mod2sum = sum over the mirror rows of re1*re2+im1*im2; rows = how many there were.
template <int KIND>
__device__ __forceinline__ void run_add(Run &r, unsigned k2, double mod2sum, int rows, float fwin, int mult,
                                        const unsigned *sT, int nrbins, float half_bpu, double *sP, double *sK,
                                        unsigned *sN)
{
    if (k2 == 0)
        return;                                             // DC mode, powerspectrum.c:63
    if (k2 < r.lo || k2 >= r.hi) {
        run_flush<KIND>(r, mult, sP, sK, sN);
        run_seek(r, k2, sT, nrbins, half_bpu, r.hi == 0);
    }
    if (KIND != PK_GEOM) {
        double w = (double)fwin;                            // float product promoted, fieldize.cpp:132
        w = w * w;                                          // invwindow() = prod^2
        w = w * w;                                          // pow(invwindow,2), powerspectrum.c:68
        r.p = fma(mod2sum, w, r.p);
    }
    if (KIND != PK_DATA)
        r.k = fma((double)rows, sqrt((double)k2), r.k);
    r.n += (unsigned)rows;
}

template <bool CROSS>
__device__ __forceinline__ double mod2(double2 va, double2 vb)
{
    return CROSS ? fma(va.x, vb.x, va.y * vb.y) : fma(va.x, va.x, va.y * va.y);
}

constexpr int POWER_THREADS = 256;
constexpr int POWER_UNROLL = 2;       // steps in flight per thread: up to 4 rows x 2 steps x 16 B

template <int KIND, bool CROSS>
__global__ void __launch_bounds__(POWER_THREADS) bin_power_kernel(PowerArgs A)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *sP = reinterpret_cast<double *>(smem_raw);
    double *sK = sP + A.nrbins;
    unsigned *sN = reinterpret_cast<unsigned *>(sK + A.nrbins);
    unsigned *sT = sN + A.nrbins;                           // nrbins + 1
    float *sW = reinterpret_cast<float *>(sT + A.nrbins + 1);   // dims/2 + 1
    const int half = A.dims / 2;
    for (int i = threadIdx.x; i < A.nrbins; i += POWER_THREADS) {
        sP[i] = 0.0;
        sK[i] = 0.0;
        sN[i] = 0;
    }
    for (int i = threadIdx.x; i <= A.nrbins; i += POWER_THREADS)
        sT[i] = A.thresh[i];
    for (int i = threadIdx.x; i <= half; i += POWER_THREADS)
        sW[i] = A.iw1d[i];
    __syncthreads();

    const int lane = threadIdx.x & 31;
    const int warps_per_cta = POWER_THREADS / 32;
    const long long warp_global = (long long)blockIdx.x * warps_per_cta + (threadIdx.x >> 5);
    const long long warp_total = (long long)gridDim.x * warps_per_cta;
    const int units_per_chunk = A.n_panels + (A.n_tail ? 1 : 0);
    const long long row_stride = A.nc;                       // modes per (outer, mid) row
    const long long outer_stride = (long long)A.n_mid * A.nc;

    for (long long u = warp_global; u < A.n_units; u += warp_total) {
        const int p = (int)(u % units_per_chunk);
        const long long oc = u / units_per_chunk;
        const int mc = (int)(oc % A.n_mchunks);
        const int o = (int)(oc / A.n_mchunks);              // outer class
        // outer rows of this class: o and, when mirrored and distinct, dims-o
        const int o_b = A.mirror_outer ? (A.dims - o) % A.dims : o;
        const bool has_ob = o_b != o;
        const int ki = kval(A.outer0 + o, A.dims);
        const float fi = sW[abs(ki)];
        const int mbeg = mc * A.jc, mend = min(A.n_mc, mbeg + A.jc);
        Run r;
        r.bin = 0; r.lo = 0; r.hi = 0; r.p = 0.0; r.k = 0.0; r.n = 0;

        if (p < A.n_panels) {
            // ---- panel unit: lane = kz, walk over mid classes ----
            const int kz = p * 32 + lane;
            const int mult = (kz == 0 || kz == half) ? 1 : 2;          // powerspectrum.c:59-89
            const float fz = sW[kz];
            const unsigned base2 = (unsigned)(ki * ki) + (unsigned)(kz * kz);
            const size_t oa_off = (size_t)o * outer_stride + kz, ob_off = (size_t)o_b * outer_stride + kz;
            for (int m0 = mbeg; m0 < mend; m0 += POWER_UNROLL) {
                double2 va[POWER_UNROLL][4], vb[POWER_UNROLL][4];
                int rows[POWER_UNROLL];
                if (KIND != PK_GEOM) {
#pragma unroll
                    for (int t = 0; t < POWER_UNROLL; t++) {
                        const int m = m0 + t;
                        if (m < mend) {
                            const int j_b = A.mirror_mid ? (A.dims - m) % A.dims : m;
                            const bool has_jb = j_b != m;
                            const size_t ja = (size_t)m * row_stride, jb = (size_t)j_b * row_stride;
                            va[t][0] = __ldcs(A.a + oa_off + ja);
                            if (CROSS) vb[t][0] = __ldcs(A.b + oa_off + ja);
                            if (has_ob) {
                                va[t][1] = __ldcs(A.a + ob_off + ja);
                                if (CROSS) vb[t][1] = __ldcs(A.b + ob_off + ja);
                            }
                            if (has_jb) {
                                va[t][2] = __ldcs(A.a + oa_off + jb);
                                if (CROSS) vb[t][2] = __ldcs(A.b + oa_off + jb);
                                if (has_ob) {
                                    va[t][3] = __ldcs(A.a + ob_off + jb);
                                    if (CROSS) vb[t][3] = __ldcs(A.b + ob_off + jb);
                                }
                            }
                        }
                    }
                }
#pragma unroll
                for (int t = 0; t < POWER_UNROLL; t++) {
                    const int m = m0 + t;
                    if (m < mend) {
                        const int j_b = A.mirror_mid ? (A.dims - m) % A.dims : m;
                        const bool has_jb = j_b != m;
                        const int kj = kval(A.mid0 + m, A.dims);
                        const float fwin = __fmul_rn(__fmul_rn(fi, sW[abs(kj)]), fz);   // (iwx*iwy)*iwz in float
                        rows[t] = (has_ob ? 2 : 1) * (has_jb ? 2 : 1);
                        double s = 0.0;
                        if (KIND != PK_GEOM) {
                            s = mod2<CROSS>(va[t][0], vb[t][0]);
                            if (has_ob) s += mod2<CROSS>(va[t][1], vb[t][1]);
                            if (has_jb) {
                                s += mod2<CROSS>(va[t][2], vb[t][2]);
                                if (has_ob) s += mod2<CROSS>(va[t][3], vb[t][3]);
                            }
                        }
                        run_add<KIND>(r, base2 + (unsigned)(kj * kj), s, rows[t], fwin, mult, sT, A.nrbins, A.half_bpu,
                                      sP, sK, sN);
                    }
                }
            }
            run_flush<KIND>(r, mult, sP, sK, sN);
        } else {
            // ---- tail unit: lane = mid class, walk over the leftover kz columns ----
            for (int mb = mbeg; mb < mend; mb += 32) {
                const int m = mb + lane;
                if (m < mend) {
                    const int j_b = A.mirror_mid ? (A.dims - m) % A.dims : m;
                    const bool has_jb = j_b != m;
                    const int kj = kval(A.mid0 + m, A.dims);
                    const float fij = __fmul_rn(fi, sW[abs(kj)]);
                    const unsigned base2 = (unsigned)(ki * ki) + (unsigned)(kj * kj);
                    const int rows = (has_ob ? 2 : 1) * (has_jb ? 2 : 1);
                    const size_t raa = (size_t)o * outer_stride + (size_t)m * row_stride;
                    const size_t rba = (size_t)o_b * outer_stride + (size_t)m * row_stride;
                    const size_t rab = (size_t)o * outer_stride + (size_t)j_b * row_stride;
                    const size_t rbb = (size_t)o_b * outer_stride + (size_t)j_b * row_stride;
                    for (int kz = A.n_panels * 32; kz < A.nc; kz++) {
                        const int mult = (kz == 0 || kz == half) ? 1 : 2;
                        double s = 0.0;
                        if (KIND != PK_GEOM) {
                            s = mod2<CROSS>(__ldcs(A.a + raa + kz), CROSS ? __ldcs(A.b + raa + kz) : make_double2(0, 0));
                            if (has_ob)
                                s += mod2<CROSS>(__ldcs(A.a + rba + kz), CROSS ? __ldcs(A.b + rba + kz) : make_double2(0, 0));
                            if (has_jb) {
                                s += mod2<CROSS>(__ldcs(A.a + rab + kz), CROSS ? __ldcs(A.b + rab + kz) : make_double2(0, 0));
                                if (has_ob)
                                    s += mod2<CROSS>(__ldcs(A.a + rbb + kz), CROSS ? __ldcs(A.b + rbb + kz) : make_double2(0, 0));
                            }
                        }
                        r.bin = 0; r.lo = 0; r.hi = 0; r.p = 0.0; r.k = 0.0; r.n = 0;
                        run_add<KIND>(r, base2 + (unsigned)(kz * kz), s, rows, __fmul_rn(fij, sW[kz]), mult, sT, A.nrbins,
                                      A.half_bpu, sP, sK, sN);
                        run_flush<KIND>(r, mult, sP, sK, sN);
                    }
                }
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < A.nrbins; i += POWER_THREADS) {
        if (KIND != PK_GEOM && sP[i] != 0.0)
            atomicAdd(&A.sums[i], sP[i]);
        if (KIND != PK_DATA && sN[i]) {
            atomicAdd(&A.sums[A.nrbins + i], sK[i]);
            atomicAdd(&A.sums[2 * A.nrbins + i], (double)sN[i]);     // exact: integers < 2^53
        }
    }
}

// Frees the device side of the bin tables and forgets the cached geometry: the next
// ensure_tables rebuilds everything (also the state a failed rebuild leaves behind).
static void drop_tables(genpk_ctx *ctx)
{
    if (ctx->d_thresh) cudaFree(ctx->d_thresh);
    if (ctx->d_iw1d) cudaFree(ctx->d_iw1d);
    if (ctx->d_sums) cudaFree(ctx->d_sums);
    if (ctx->d_geom) cudaFree(ctx->d_geom);
    if (ctx->h_sums) cudaFreeHost(ctx->h_sums);
    ctx->d_thresh = nullptr; ctx->d_iw1d = nullptr; ctx->d_sums = nullptr; ctx->h_sums = nullptr; ctx->d_geom = nullptr;
    ctx->geom_valid = false;
    ctx->sums_cap = 0;
    ctx->tables.nrbins = 0;
    ctx->tables.dims = 0;
}

static int upload_tables(genpk_ctx *ctx, const BinTables &t, int nrbins)
{
    GENPK_CUDA_OK(cudaMalloc(&ctx->d_thresh, t.thresh.size() * sizeof(uint32_t)));
    GENPK_CUDA_OK(cudaMalloc(&ctx->d_iw1d, t.iw1d.size() * sizeof(float)));
    GENPK_CUDA_OK(cudaMalloc(&ctx->d_sums, (size_t)3 * nrbins * sizeof(double)));
    GENPK_CUDA_OK(cudaMalloc(&ctx->d_geom, (size_t)3 * nrbins * sizeof(double)));
    GENPK_CUDA_OK(cudaMallocHost(&ctx->h_sums, (size_t)3 * nrbins * sizeof(double)));
    GENPK_CUDA_OK(cudaMemcpyAsync(ctx->d_thresh, t.thresh.data(), t.thresh.size() * sizeof(uint32_t), cudaMemcpyHostToDevice,
                                  ctx->stream));
    GENPK_CUDA_OK(cudaMemcpyAsync(ctx->d_iw1d, t.iw1d.data(), t.iw1d.size() * sizeof(float), cudaMemcpyHostToDevice,
                                  ctx->stream));
    GENPK_CUDA_OK(cudaStreamSynchronize(ctx->stream));     // the host vectors may be rebuilt later
    return 0;
}

int ensure_tables(genpk_ctx *ctx, int nrbins)
{
    const unsigned rule = (ctx->flags & GENPK_FLAG_BINRULE_SOURCE) ? 1u : 0u;
    if (ctx->tables.nrbins == nrbins && ctx->tables.dims == ctx->g.dims && ctx->tables.rule == rule && ctx->d_thresh)
        return 0;
    // built aside and committed only once every allocation and upload has succeeded: a failure
    // leaves no half-updated cache behind (the next call starts from scratch)
    BinTables fresh;
    if (int rc = build_bin_tables(ctx->g.dims, nrbins, rule, &fresh))
        return rc;
    drop_tables(ctx);
    if (int rc = upload_tables(ctx, fresh, nrbins)) {
        drop_tables(ctx);
        return rc;
    }
    ctx->tables = fresh;
    ctx->sums_cap = nrbins;
    return 0;
}

template <int KIND, bool CROSS>
static int launch_power(genpk_ctx *ctx, const PowerArgs &A, size_t smem)
{
    auto kern = bin_power_kernel<KIND, CROSS>;
    GENPK_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    GENPK_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, POWER_THREADS, smem));
    if (per_sm < 1) {
        set_error("bin_power: %zu bytes of shared memory do not fit (nrbins=%d dims=%d)", smem, A.nrbins, A.dims);
        return 1;
    }
    long long ctas = (long long)per_sm * ctx->sm_count;
    const long long need = (A.n_units + POWER_THREADS / 32 - 1) / (POWER_THREADS / 32);
    if (ctas > need) ctas = need;
    if (ctas < 1) ctas = 1;
    kern<<<(int)ctas, POWER_THREADS, smem, ctx->stream>>>(A);
    ctx->launches++;
    GENPK_CUDA_OK(cudaGetLastError());
    return 0;
}

static void fill_power_args(genpk_ctx *ctx, PowerArgs &A, int n_outer, int outer0, int n_mid, int mid0, int nrbins)
{
    A.a = A.b = nullptr;
    A.dims = ctx->g.dims;
    A.nc = ctx->g.nc;
    A.n_outer = n_outer;
    A.outer0 = outer0;
    A.n_mid = n_mid;
    A.mid0 = mid0;
    A.nrbins = nrbins;
    A.mirror_outer = (outer0 == 0 && n_outer == A.dims) ? 1 : 0;
    A.mirror_mid = (mid0 == 0 && n_mid == A.dims) ? 1 : 0;
    A.n_oc = A.mirror_outer ? A.dims / 2 + 1 : n_outer;
    A.n_mc = A.mirror_mid ? A.dims / 2 + 1 : n_mid;
    A.jc = 32;
    A.n_mchunks = (A.n_mc + A.jc - 1) / A.jc;
    A.n_panels = A.nc / 32;
    A.n_tail = A.nc % 32;
    A.n_units = (long long)A.n_oc * A.n_mchunks * (A.n_panels + (A.n_tail ? 1 : 0));
    A.iw1d = ctx->d_iw1d;
    A.thresh = ctx->d_thresh;
    A.half_bpu = nrbins > 1 ? (float)(0.5 * (nrbins - 1) / log(sqrt(3.0) * A.dims / 2.0)) : 0.f;
    A.sums = nullptr;
}

static size_t power_smem(const PowerArgs &A)
{
    return (size_t)A.nrbins * (8 + 8 + 4) + (size_t)(A.nrbins + 1) * 4 + (size_t)(A.dims / 2 + 1) * 4 + 16;
}

// sums_dev = {0, K, N}: the P part cleared, sum|k| and the mode counts of the block
// [n_outer][n_mid][nc] copied from the geometry pass cached in the context (run here on a miss).
int power_seed_sums(genpk_ctx *ctx, int n_outer, int outer0, int n_mid, int mid0, int nrbins, double *sums_dev)
{
    if (int rc = ensure_tables(ctx, nrbins))
        return rc;
    const size_t sums_bytes = (size_t)3 * nrbins * sizeof(double);
    const long long key[5] = {nrbins, n_outer, outer0, n_mid, mid0};
    bool hit = ctx->geom_valid;
    for (int i = 0; i < 5 && hit; i++)
        hit = ctx->geom_key[i] == key[i];
    if (!hit) {
        GENPK_CUDA_OK(cudaMemsetAsync(ctx->d_geom, 0, sums_bytes, ctx->stream));
        PowerArgs G;
        fill_power_args(ctx, G, n_outer, outer0, n_mid, mid0, nrbins);
        G.sums = ctx->d_geom;
        if (int rc = launch_power<PK_GEOM, false>(ctx, G, power_smem(G)))
            return rc;
        for (int i = 0; i < 5; i++)
            ctx->geom_key[i] = key[i];
        ctx->geom_valid = true;
    }
    GENPK_CUDA_OK(cudaMemsetAsync(sums_dev, 0, (size_t)nrbins * sizeof(double), ctx->stream));
    GENPK_CUDA_OK(cudaMemcpyAsync(sums_dev + nrbins, ctx->d_geom + nrbins, (size_t)2 * nrbins * sizeof(double),
                                  cudaMemcpyDeviceToDevice, ctx->stream));
    return 0;
}

// Raw per-bin sums of a spectrum block [n_outer][n_mid][nc] whose first outer
// (mid) index is global FFT index outer0 (mid0).  sums_dev: 3*nrbins doubles.
int power_raw(genpk_ctx *ctx, const double *spec_a, const double *spec_b, int n_outer, int outer0, int n_mid,
              int mid0, int nrbins, double *sums_dev)
{
    if (int rc = ensure_tables(ctx, nrbins))
        return rc;
    PowerArgs A;
    fill_power_args(ctx, A, n_outer, outer0, n_mid, mid0, nrbins);
    A.a = reinterpret_cast<const double2 *>(spec_a);
    A.b = reinterpret_cast<const double2 *>(spec_b);
    A.sums = sums_dev;
    const size_t smem = power_smem(A);
    const bool cross = spec_b != spec_a;

    if (ctx->power_mode == GENPK_POWER_FUSED) {
        GENPK_CUDA_OK(cudaMemsetAsync(sums_dev, 0, (size_t)3 * nrbins * sizeof(double), ctx->stream));
        return cross ? launch_power<PK_FUSED, true>(ctx, A, smem) : launch_power<PK_FUSED, false>(ctx, A, smem);
    }
    // P from the data pass; K and N from the geometry sums cached per block of the spectrum
    if (int rc = power_seed_sums(ctx, n_outer, outer0, n_mid, mid0, nrbins, sums_dev))
        return rc;
    return cross ? launch_power<PK_DATA, true>(ctx, A, smem) : launch_power<PK_DATA, false>(ctx, A, smem);
}

}  // namespace genpk
