// C ABI of libgenpk_cuda.so (see include/genpk_cuda.h).
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <new>
#include <vector>

#include "common.cuh"

namespace genpk {

static thread_local char g_error[1024] = "";

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

int fieldize_host_shim(double boxsize, int dims, double *out, int64_t n, const float *positions,
                       const float *masses, double mass, int extra);

static bool check_which(const genpk_ctx *ctx, int which, const char *fn)
{
    if (!ctx) {
        set_error("%s: null context", fn);
        return false;
    }
    if (which < 0 || which > 1 || !ctx->grid[which]) {
        set_error("%s: grid %d not allocated (create the context with GENPK_FLAG_TWO_FIELDS for grid 1)", fn, which);
        return false;
    }
    return true;
}

static genpk_ctx *create_common(int dims, int device, int nranks, int rank, unsigned flags, int ghost_planes)
{
    if (ghost_planes < 0 || (nranks > 1 && dims % nranks == 0 &&
                             (ghost_planes > dims / nranks || dims / nranks + 2 * ghost_planes > dims))) {
        set_error("genpk_create: ghost_planes=%d must satisfy ghost <= dims/nranks and dims/nranks + 2*ghost <= dims",
                  ghost_planes);
        return nullptr;
    }
    if (dims < 1 || nranks < 1 || rank < 0 || rank >= nranks || dims % nranks != 0) {
        set_error("genpk_create: bad geometry dims=%d nranks=%d rank=%d (dims must be divisible by nranks)", dims,
                  nranks, rank);
        return nullptr;
    }
    if (device >= 0 && cudaSetDevice(device) != cudaSuccess) {
        set_error("genpk_create: cudaSetDevice(%d) failed: %s", device, cudaGetErrorString(cudaGetLastError()));
        return nullptr;
    }
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) {
        set_error("genpk_create: no CUDA device: %s", cudaGetErrorString(cudaGetLastError()));
        return nullptr;
    }
    genpk_ctx *ctx = new (std::nothrow) genpk_ctx();
    if (!ctx) {
        set_error("genpk_create: out of host memory");
        return nullptr;
    }
    ctx->device = dev;
    ctx->flags = flags;

    ctx->fixed = (flags & GENPK_FLAG_FIXED_POINT) != 0;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, dev) == cudaSuccess) {
        ctx->sm_count = prop.multiProcessorCount;
        ctx->l2_bytes = (size_t)prop.l2CacheSize;
        ctx->smem_optin = (int)prop.sharedMemPerBlockOptin;
    }
    cudaDeviceGetAttribute(&ctx->coop_launch, cudaDevAttrCooperativeLaunch, dev);
    SlabGeom &g = ctx->g;
    g.dims = dims;
    g.nranks = nranks;
    g.rank = rank;
    g.nx = dims / nranks;
    g.x0 = rank * g.nx;
    // plain slab: one ghost plane above (CIC reaches one plane up).  Wide slab: G planes on both
    // sides, so a rank can deposit a slab-local particle shard whose stragglers sit up to G planes
    // outside its slab without any particle exchange.
    g.ghost_lo = nranks > 1 ? ghost_planes : 0;
    g.ghost_hi = nranks > 1 ? (ghost_planes > 0 ? ghost_planes : 1) : 0;
    g.nc = dims / 2 + 1;
    g.fd = 2 * g.nc;
    const int ngrids = (flags & GENPK_FLAG_TWO_FIELDS) ? 2 : 1;
    bool ok = true;
    for (int i = 0; i < ngrids && ok; i++) {
        ok = cudaMalloc(&ctx->grid[i], g.grid_doubles() * sizeof(double) + GRID_TAIL_BYTES) == cudaSuccess;
        // touched-plane range: "every plane" until a deposit that tracks it says otherwise
        const int full[2] = {0, g.ghost_lo + g.nx + g.ghost_hi - 1};
        ok = ok && cudaMemcpy(ctx->grid[i] + g.grid_doubles(), full, sizeof(full), cudaMemcpyHostToDevice) == cudaSuccess;
    }
    ok = ok && cudaMalloc(&ctx->d_errors, sizeof(unsigned long long)) == cudaSuccess;
    ok = ok && cudaMemset(ctx->d_errors, 0, sizeof(unsigned long long)) == cudaSuccess;
    ok = ok && cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking) == cudaSuccess;
    for (int i = 0; i < ST_COUNT && ok; i++)
        for (int s = 0; s < genpk_ctx::EV_SLOTS && ok; s++)
            ok = cudaEventCreate(&ctx->ev_begin[i][s]) == cudaSuccess && cudaEventCreate(&ctx->ev_end[i][s]) == cudaSuccess;
    for (int i = 0; i < 2 && ok; i++)
        ok = cudaEventCreateWithFlags(&ctx->stage_free[i], cudaEventDisableTiming) == cudaSuccess;
    if (!ok) {
        set_error("genpk_create: CUDA allocation failed for dims=%d (%zu bytes per grid): %s", dims,
                  g.grid_doubles() * sizeof(double), cudaGetErrorString(cudaGetLastError()));
        genpk_destroy(ctx);
        return nullptr;
    }
    return ctx;
}

}  // namespace genpk

using namespace genpk;

extern "C" {

const char *genpk_last_error(void) { return g_error; }
int genpk_abi_version(void) { return GENPK_ABI_VERSION; }

genpk_ctx *genpk_create(int dims, int device, unsigned flags) { return create_common(dims, device, 1, 0, flags, 0); }

genpk_ctx *genpk_create_slab(int dims, int device, int nranks, int rank, unsigned flags)
{
    return create_common(dims, device, nranks, rank, flags, 0);
}

genpk_ctx *genpk_create_slab_wide(int dims, int device, int nranks, int rank, unsigned flags, int ghost_planes)
{
    return create_common(dims, device, nranks, rank, flags, ghost_planes);
}

void genpk_destroy(genpk_ctx *ctx)
{
    if (!ctx)
        return;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    fft_release(ctx);
    for (int i = 0; i < 2; i++) {
        if (ctx->grid[i]) cudaFree(ctx->grid[i]);
        if (ctx->d_stage_pos[i]) cudaFree(ctx->d_stage_pos[i]);
        if (ctx->d_stage_mass[i]) cudaFree(ctx->d_stage_mass[i]);
        if (ctx->d_stage_pos64[i]) cudaFree(ctx->d_stage_pos64[i]);
        if (ctx->stage_free[i]) cudaEventDestroy(ctx->stage_free[i]);
    }
    if (ctx->d_iw1d) cudaFree(ctx->d_iw1d);
    if (ctx->d_thresh) cudaFree(ctx->d_thresh);
    if (ctx->d_sums) cudaFree(ctx->d_sums);
    if (ctx->d_geom) cudaFree(ctx->d_geom);
    if (ctx->h_sums) cudaFreeHost(ctx->h_sums);
    if (ctx->d_sorted_pos) cudaFree(ctx->d_sorted_pos);
    if (ctx->d_sorted_mass) cudaFree(ctx->d_sorted_mass);
    if (ctx->d_brick_counts) cudaFree(ctx->d_brick_counts);
    if (ctx->d_errors) cudaFree(ctx->d_errors);
    if (ctx->d_order) cudaFree(ctx->d_order);
    if (ctx->d_maxmass) cudaFree(ctx->d_maxmass);
    if (ctx->d_za_zdone) cudaFree(ctx->d_za_zdone);
    if (ctx->d_rows_done) cudaFree(ctx->d_rows_done);
    if (ctx->d_za_def) cudaFree(ctx->d_za_def);
    if (ctx->d_twiddle) cudaFree(ctx->d_twiddle);
    for (int r = 0; r < GENPK_MAX_PEERS; r++)
        if (ctx->peer_opened[r] && ctx->peer_recv[r]) cudaIpcCloseMemHandle(ctx->peer_recv[r]);
    for (int w = 0; w < 2; w++)
        for (int side = 0; side < 2; side++)
            if (ctx->grid_peer_opened[w][side] && ctx->grid_peer[w][side]) {
                // the same allocation may be mapped for both sides (two ranks): close it once
                if (side == 1 && ctx->grid_peer[w][0] == ctx->grid_peer[w][1] && ctx->grid_peer_opened[w][0]) continue;
                cudaIpcCloseMemHandle(ctx->grid_peer[w][side]);
            }
    if (ctx->d_recv) cudaFree(ctx->d_recv);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    for (int i = 0; i < ST_COUNT; i++)
        for (int s = 0; s < genpk_ctx::EV_SLOTS; s++) {
            if (ctx->ev_begin[i][s]) cudaEventDestroy(ctx->ev_begin[i][s]);
            if (ctx->ev_end[i][s]) cudaEventDestroy(ctx->ev_end[i][s]);
        }
    delete ctx;
}

int genpk_set_stream(genpk_ctx *ctx, void *cuda_stream)
{
    if (!ctx) { set_error("genpk_set_stream: null context"); return 1; }
    ctx->stream = (cudaStream_t)cuda_stream;
    return 0;
}

int genpk_set_option(genpk_ctx *ctx, int option, int64_t value)
{
    if (!ctx) { set_error("genpk_set_option: null context"); return 1; }
    ctx->plan_key_pos = nullptr;                                 // a cached deposit plan may no longer be the one wanted
    switch (option) {
    case GENPK_OPT_DEPOSIT:
        if (value < GENPK_DEPOSIT_AUTO || value > GENPK_DEPOSIT_SWEEP) break;
        ctx->deposit_mode = (int)value;
        return 0;
    case GENPK_OPT_SCALE_BITS:
        if (value < -1 || value > 62) break;
        ctx->scale_bits = (int)value;
        return 0;
    case GENPK_OPT_LATTICE_N0:
        if (value < 0) break;
        ctx->lattice_n0 = value;
        return 0;
    case GENPK_OPT_LATTICE_N1:
        if (value < 0) break;
        ctx->lattice_n1 = value;
        return 0;
    case GENPK_OPT_MARCH_RY:
        if (value < 1 || value > 32) break;
        ctx->march_ry = (int)value;
        return 0;
    case GENPK_OPT_MARCH_RX:
        if (value < 1 || value > 4096) break;
        ctx->march_rx = (int)value;
        return 0;
    case GENPK_OPT_POWER:
        if (value != GENPK_POWER_CACHED && value != GENPK_POWER_FUSED) break;
        ctx->power_mode = (int)value;
        return 0;
    case GENPK_OPT_FFT_YZ_BATCH:
        if (value < 0 || value > (1 << 20)) break;
        ctx->fft_yz_batch = (int)value;
        return 0;
    case GENPK_OPT_OWN_YPASS:
        if (value != 0 && value != 1) break;
        ctx->own_ypass = (int)value;
        return 0;
    case GENPK_OPT_TMA:
        if (value != 0 && value != 1) break;
        ctx->use_tma = (int)value;
        return 0;
    case GENPK_OPT_FUSED_ZY:
        if (value < 0 || value > 2) break;
        ctx->fused_zy = (int)value;
        return 0;
    case GENPK_OPT_ZY_LAG:
        if (value < 1 || value > 64) break;
        ctx->zy_lag = (int)value;
        return 0;
    case GENPK_OPT_ZERO_AFTER_POWER:
        if (value != 0 && value != 1) break;
        ctx->zero_after_power = (int)value;
        return 0;
    case GENPK_OPT_FUSED_XPASS:
        if (value < 0 || value > 4) break;
        ctx->fused_xpass = (int)value;
        return 0;
    case GENPK_OPT_F64_POSITIONS:
        if (value != 0 && value != 1) break;
        ctx->f64_exact = (int)value;
        return 0;
    case GENPK_OPT_SWEEP:
        if (value != 0 && value != 1) break;
        ctx->sweep = (int)value;
        return 0;
    case GENPK_OPT_SWEEP_RY:
        if (value < 0 || value > 64) break;
        ctx->sweep_ry = (int)value;
        return 0;
    case GENPK_OPT_ZERO_AHEAD:
        if (value != 0 && value != 1) break;
        ctx->zero_ahead = (int)value;
        return 0;
    case GENPK_OPT_ZA_WINDOW:
        if (value < 0 || value > 4096) break;
        ctx->za_window = (int)value;
        return 0;
    case GENPK_OPT_ZA_SLACK:
        if (value < 0 || value > 64) break;
        ctx->za_slack = (int)value;
        return 0;
    case GENPK_OPT_ZA_ZERO_CTAS:
        if (value < 0 || value > 1024) break;
        ctx->za_zero_ctas = (int)value;
        return 0;
    case GENPK_OPT_SWEEP_COUPLE:
        if (value < 0 || value > 4096) break;
        ctx->sweep_couple = (int)value;
        return 0;
    case GENPK_OPT_SWEEP_RX:
        if (value < 0 || value > 65535) break;
        ctx->sweep_rx = (int)value;
        return 0;
    case GENPK_OPT_SWEEP_COUPLE_STEP:
        if (value < 1 || value > 4096) break;
        ctx->sweep_couple_step = (int)value;
        return 0;
    case GENPK_OPT_SWEEP_POLL_WEAK:
        if (value != 0 && value != 1) break;
        ctx->sweep_poll_weak = (int)value;
        return 0;
    case GENPK_OPT_ZA_DEFERRED:
        if (value < 1 || value > (1 << 20)) break;
        ctx->za_def_per_col = (int)value;
        return 0;
    }
    set_error("genpk_set_option: bad option %d / value %lld", option, (long long)value);
    return 1;
}

int genpk_synchronize(genpk_ctx *ctx)
{
    if (!ctx) { set_error("genpk_synchronize: null context"); return 1; }
    GENPK_CUDA_OK(cudaStreamSynchronize(ctx->stream));
    unsigned long long bad = 0;
    GENPK_CUDA_OK(cudaMemcpy(&bad, ctx->d_errors, sizeof(bad), cudaMemcpyDeviceToHost));
    if (bad) {
        cudaMemset(ctx->d_errors, 0, sizeof(bad));
        set_error("%llu particles rejected (non-finite position, or outside this rank's x-slab)", bad);
        return 3;
    }
    return 0;
}

__global__ void rejected_to_kernel(unsigned long long *errors, double *dst)
{
    *dst = (double)*errors;
    *errors = 0ull;
}

int genpk_rejected_to(genpk_ctx *ctx, double *dst_dev)
{
    if (!ctx || !dst_dev) { set_error("genpk_rejected_to: bad arguments"); return 1; }
    rejected_to_kernel<<<1, 1, 0, ctx->stream>>>(ctx->d_errors, dst_dev);
    ctx->launches++;
    GENPK_CUDA_OK(cudaGetLastError());
    return 0;
}

int genpk_take_rejected(genpk_ctx *ctx, uint64_t *rejected)
{
    if (!ctx || !rejected) { set_error("genpk_take_rejected: bad arguments"); return 1; }
    GENPK_CUDA_OK(cudaStreamSynchronize(ctx->stream));
    unsigned long long bad = 0;
    GENPK_CUDA_OK(cudaMemcpy(&bad, ctx->d_errors, sizeof(bad), cudaMemcpyDeviceToHost));
    if (bad)
        GENPK_CUDA_OK(cudaMemset(ctx->d_errors, 0, sizeof(bad)));
    *rejected = bad;
    return 0;
}

int genpk_grid_zero(genpk_ctx *ctx, int which)
{
    if (!check_which(ctx, which, "genpk_grid_zero")) return 1;
    // Lazy: the memset runs when the grid is next read or deposited into -- unless that next use is
    // a lattice sweep, which clears the grid ahead of its own front instead (deposit_sweep.cu) and
    // saves one write and one read of the whole grid.
    // A grid that the fused x pass has just left all zeros (GENPK_OPT_ZERO_AFTER_POWER) needs nothing.
    ctx->zero_pending[which] = !ctx->grid_clean[which];
    ctx->grid_scale_latched[which] = false;                      // the next deposit fixes the fixed-point scale anew
    if (!ctx->zero_ahead)
        if (int rc = materialize_zero(ctx, which)) return rc;
    // an all-zero grid is valid in either representation: take it from the context's mode, so a
    // slab rank that deposits nothing still adds its neighbours' int64 ghost planes as integers
    ctx->grid_is_fixed[which] = ctx->fixed;
    return 0;
}

static int ensure_stage(genpk_ctx *ctx, int64_t cap, bool with_mass)
{
    if (cap > ctx->stage_cap) {
        for (int i = 0; i < 2; i++) {
            if (ctx->d_stage_pos[i]) cudaFree(ctx->d_stage_pos[i]);
            if (ctx->d_stage_mass[i]) cudaFree(ctx->d_stage_mass[i]);
            ctx->d_stage_pos[i] = nullptr;
            ctx->d_stage_mass[i] = nullptr;
        }
        ctx->stage_cap = 0;
        for (int i = 0; i < 2; i++)
            GENPK_CUDA_OK(cudaMalloc(&ctx->d_stage_pos[i], (size_t)cap * 3 * sizeof(float)));
        ctx->stage_cap = cap;
    }
    if (with_mass && !ctx->d_stage_mass[0])
        for (int i = 0; i < 2; i++)
            GENPK_CUDA_OK(cudaMalloc(&ctx->d_stage_mass[i], (size_t)ctx->stage_cap * sizeof(float)));
    return 0;
}

// positions[i] = (float)((double *)pos)[i]: the narrowing of read_fieldize_bigfile.cpp:93-94 (round to nearest)
__global__ void narrow_f64_kernel(const double *src, float *dst, size_t n)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        dst[i] = __double2float_rn(src[i]);
}

static int ensure_stage64(genpk_ctx *ctx, int64_t cap)
{
    if (cap > ctx->stage64_cap) {
        for (int i = 0; i < 2; i++) {
            if (ctx->d_stage_pos64[i]) cudaFree(ctx->d_stage_pos64[i]);
            ctx->d_stage_pos64[i] = nullptr;
        }
        ctx->stage64_cap = 0;
        for (int i = 0; i < 2; i++)
            GENPK_CUDA_OK(cudaMalloc(&ctx->d_stage_pos64[i], (size_t)cap * 3 * sizeof(double)));
        ctx->stage64_cap = cap;
    }
    return 0;
}

// Chunk loop shared by genpk_deposit (float positions) and genpk_deposit_f64 (double positions,
// narrowed on the GPU).  Host particles: chunks go up on the copy stream into two device staging
// buffers while the previous chunk is being deposited (chunk loop of read_fieldize.cpp:51-93,
// overlapped).  The deposit is planned once, on the first chunk (the only host synchronisation of
// the loop); when that finds a lattice the following chunks end on lattice-plane (or row)
// boundaries so every chunk marches from a row start.
static int deposit_chunks(genpk_ctx *ctx, int which, const void *positions, bool f64, const float *masses, int64_t n,
                          double mass, double boxsize, int on_device)
{
    const int64_t chunk = n < ((int64_t)1 << 23) ? n : ((int64_t)1 << 23);
    // a pending genpk_grid_zero: chunks are partial sweeps, so the grid is cleared by a memset here
    // (it overlaps the first upload); one-call device-resident deposits clear it ahead of their front
    if (n > chunk || !on_device)
        if (int rc = materialize_zero(ctx, which)) return rc;
    if (int rc = ensure_stage(ctx, chunk, masses != nullptr && !on_device)) return rc;
    if (f64 && !on_device)
        if (int rc = ensure_stage64(ctx, chunk)) return rc;
    const float *pos32 = reinterpret_cast<const float *>(positions);
    const double *pos64 = reinterpret_cast<const double *>(positions);
    int rc = 0, buf = 0;
    DepositPlan plan;
    bool planned = false;
    int64_t unit = 1;
    for (int64_t off = 0; off < n && !rc; buf ^= 1) {
        int64_t m = (n - off) < chunk ? (n - off) : chunk;
        if (planned && off + m < n && m > unit)
            m = m / unit * unit;
        const float *dmass = masses ? masses + off : nullptr;
        if (!on_device) {
            GENPK_CUDA_OK(cudaStreamWaitEvent(ctx->copy_stream, ctx->stage_free[buf], 0));
            if (f64)
                GENPK_CUDA_OK(cudaMemcpyAsync(ctx->d_stage_pos64[buf], pos64 + 3 * off, (size_t)m * 3 * sizeof(double),
                                              cudaMemcpyHostToDevice, ctx->copy_stream));
            else
                GENPK_CUDA_OK(cudaMemcpyAsync(ctx->d_stage_pos[buf], pos32 + 3 * off, (size_t)m * 3 * sizeof(float),
                                              cudaMemcpyHostToDevice, ctx->copy_stream));
            if (masses) {
                GENPK_CUDA_OK(cudaMemcpyAsync(ctx->d_stage_mass[buf], masses + off, (size_t)m * sizeof(float),
                                              cudaMemcpyHostToDevice, ctx->copy_stream));
                dmass = ctx->d_stage_mass[buf];
            }
            cudaEvent_t up;   // chunk uploaded
            GENPK_CUDA_OK(cudaEventCreateWithFlags(&up, cudaEventDisableTiming));
            GENPK_CUDA_OK(cudaEventRecord(up, ctx->copy_stream));
            GENPK_CUDA_OK(cudaStreamWaitEvent(ctx->stream, up, 0));
            GENPK_CUDA_OK(cudaEventDestroy(up));
        }
        if (f64 && ctx->f64_exact) {
            // a DOUBLE_PRECISION_SNAP build of the reference: the doubles go to the deposit as they are
            const double *src = on_device ? pos64 + 3 * off : ctx->d_stage_pos64[buf];
            rc = deposit_device_f64(ctx, which, src, dmass, m, mass, boxsize);
            if (!on_device)
                GENPK_CUDA_OK(cudaEventRecord(ctx->stage_free[buf], ctx->stream));
            off += m;
            continue;
        }
        if (f64) {
            const double *src = on_device ? pos64 + 3 * off : ctx->d_stage_pos64[buf];
            narrow_f64_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(src, ctx->d_stage_pos[buf], (size_t)m * 3);
            ctx->launches++;
            GENPK_CUDA_OK(cudaGetLastError());
        }
        const float *dpos = (f64 || !on_device) ? ctx->d_stage_pos[buf] : pos32 + 3 * off;
        if (!planned) {
            if ((rc = deposit_plan(ctx, dpos, m, boxsize, &plan))) break;
            planned = true;
            if ((plan.mode == GENPK_DEPOSIT_MARCH || plan.mode == GENPK_DEPOSIT_SWEEP) && plan.n0 > 0) {
                const int64_t plane = plan.n1 > 0 ? plan.n0 * plan.n1 : 0;
                unit = (plane > 0 && plane <= chunk) ? plane : (plan.n0 <= chunk ? plan.n0 : 1);
                if (off + m < n && m > unit)
                    m = m / unit * unit;           // the tail of this chunk is handled again with the next one
            }
        }
        rc = deposit_device(ctx, which, dpos, dmass, m, mass, boxsize, &plan);
        if (!on_device || f64)
            GENPK_CUDA_OK(cudaEventRecord(ctx->stage_free[buf], ctx->stream));
        off += m;
    }
    if (!on_device) {
        // the uploads read the caller's buffers asynchronously: they are over when this call returns
        // (the deposits of the last chunks may still be running on the context's stream)
        GENPK_CUDA_OK(cudaStreamSynchronize(ctx->copy_stream));
    }
    return rc;
}

int genpk_deposit(genpk_ctx *ctx, int which, const float *positions, const float *masses, int64_t n, double mass,
                  double boxsize, int on_device)
{
    if (!check_which(ctx, which, "genpk_deposit")) return 1;
    if (n < 0 || (n > 0 && !positions)) { set_error("genpk_deposit: bad particle array"); return 1; }
    if (n == 0) return 0;
    StageScope scope(ctx, ST_DEPOSIT);
    if (on_device)
        return deposit_device(ctx, which, positions, masses, n, mass, boxsize);        // one launch over the resident set
    return deposit_chunks(ctx, which, positions, false, masses, n, mass, boxsize, 0);
}

int genpk_deposit_f64(genpk_ctx *ctx, int which, const double *positions, const float *masses, int64_t n, double mass,
                      double boxsize, int on_device)
{
    if (!check_which(ctx, which, "genpk_deposit_f64")) return 1;
    if (n < 0 || (n > 0 && !positions)) { set_error("genpk_deposit_f64: bad particle array"); return 1; }
    if (n == 0) return 0;
    StageScope scope(ctx, ST_DEPOSIT);
    return deposit_chunks(ctx, which, positions, true, masses, n, mass, boxsize, on_device);
}

int genpk_fft(genpk_ctx *ctx, int which)
{
    if (!check_which(ctx, which, "genpk_fft")) return 1;
    if (ctx->g.nranks != 1) {
        set_error("genpk_fft: slab contexts use genpk_slab_fft_yz / genpk_slab_pack / genpk_slab_fft_x");
        return 1;
    }
    {
        StageScope scope(ctx, ST_FFT);
        if (int rc = fixed_to_double(ctx, which)) return rc;
        if (int rc = fft_3d(ctx, which)) return rc;
    }
    return 0;
}

int genpk_power_finalize(const double *sums, int nrbins, double total_mass, double total_mass2, double *power,
                         int *count, double *keffs)
{
    if (!sums || !power || !count || !keffs || nrbins < 1) { set_error("genpk_power_finalize: bad arguments"); return 1; }
    for (int b = 0; b < nrbins; b++) {
        const long long c = (long long)sums[2 * (size_t)nrbins + b];
        count[b] = (int)c;                                   // int count[], powerspectrum.c:35
        power[b] = sums[b];
        keffs[b] = sums[(size_t)nrbins + b];
        if (count[b]) {                                      // powerspectrum.c:102-108
            power[b] /= total_mass * total_mass2;
            power[b] /= count[b];
            keffs[b] /= count[b];
        }
    }
    return 0;
}

int genpk_rebin_min_modes(int nrbins, double *power, int *count, double *keffs, int64_t min_modes)
{
    if (nrbins < 1 || !power || !count || !keffs) { set_error("genpk_rebin_min_modes: bad arguments"); return -1; }
    int out = 0, merged = 0, only = -1;                      // input bins in the open output bin; the bin if there is one
    double p_sum = 0, k_sum = 0;
    long long n_sum = 0;
    auto close = [&]() {
        if (merged == 1) {                                   // a bin on its own keeps its values bit for bit
            power[out] = power[only];
            keffs[out] = keffs[only];
        } else {
            power[out] = p_sum / (double)n_sum;
            keffs[out] = k_sum / (double)n_sum;
        }
        count[out] = (int)n_sum;
        out++;
        p_sum = k_sum = 0;
        n_sum = 0;
        merged = 0;
    };
    auto add = [&](int b) {
        const double c = (double)count[b];
        p_sum += c * power[b];                               // = the raw sum over the bin's modes / (tm1 * tm2)
        k_sum += c * keffs[b];
        n_sum += count[b];
        merged++;
        only = b;
    };
    for (int b = 0; b < nrbins; b++) {
        if (count[b] <= 0)
            continue;
        add(b);                                              // (b >= out: nothing that is still to be read is overwritten)
        if (n_sum >= min_modes)
            close();
    }
    if (n_sum > 0) {                                         // a short tail joins the last full bin
        if (out > 0) {
            out--;
            add(out);
            merged++;                                        // (never a bin on its own)
        }
        close();
    }
    for (int b = out; b < nrbins; b++) {
        power[b] = keffs[b] = 0;
        count[b] = 0;
    }
    return out;
}

static int power_on(genpk_ctx *ctx, const double *spec_a, const double *spec_b, int nrbins, double *power, int *count,
                    double *keffs, double total_mass, double total_mass2)
{
    if (nrbins < 1 || !power || !count || !keffs) { set_error("genpk_power: bad arguments"); return 1; }
    if (ctx->g.nranks != 1) { set_error("genpk_power: slab contexts use genpk_slab_power_partial"); return 1; }
    if (int rc = ensure_tables(ctx, nrbins)) return rc;
    {
        StageScope scope(ctx, ST_POWER);
        if (int rc = power_raw(ctx, spec_a, spec_b, ctx->g.dims, 0, ctx->g.dims, 0, nrbins, ctx->d_sums)) return rc;
    }
    GENPK_CUDA_OK(cudaMemcpyAsync(ctx->h_sums, ctx->d_sums, (size_t)3 * nrbins * sizeof(double), cudaMemcpyDeviceToHost,
                                  ctx->stream));
    GENPK_CUDA_OK(cudaStreamSynchronize(ctx->stream));
    return genpk_power_finalize(ctx->h_sums, nrbins, total_mass, total_mass2, power, count, keffs);
}

int genpk_power(genpk_ctx *ctx, int a, int b, int nrbins, double *power, int *count, double *keffs, double total_mass,
                double total_mass2)
{
    if (!check_which(ctx, a, "genpk_power") || !check_which(ctx, b, "genpk_power")) return 1;
    if (int rc = materialize_zero(ctx, a)) return rc;
    if (int rc = materialize_zero(ctx, b)) return rc;
    return power_on(ctx, ctx->grid[a], ctx->grid[b], nrbins, power, count, keffs, total_mass, total_mass2);
}

int genpk_power_dev(genpk_ctx *ctx, const void *spec_a_dev, const void *spec_b_dev, int nrbins, double *power,
                    int *count, double *keffs, double total_mass, double total_mass2)
{
    if (!ctx || !spec_a_dev) { set_error("genpk_power_dev: bad arguments"); return 1; }
    const double *a = reinterpret_cast<const double *>(spec_a_dev);
    const double *b = spec_b_dev ? reinterpret_cast<const double *>(spec_b_dev) : a;
    return power_on(ctx, a, b, nrbins, power, count, keffs, total_mass, total_mass2);
}

int genpk_fft_power(genpk_ctx *ctx, int which, int nrbins, double *power, int *count, double *keffs, double total_mass,
                    double total_mass2)
{
    if (!check_which(ctx, which, "genpk_fft_power")) return 1;
    if (nrbins < 1 || !power || !count || !keffs) { set_error("genpk_fft_power: bad arguments"); return 1; }
    if (ctx->g.nranks != 1) { set_error("genpk_fft_power: slab contexts use genpk_slab_fftx_power_partial"); return 1; }
    if (!fftx_supported(ctx, nrbins)) {                    // other grid sides: the library transform + the binning pass
        if (int rc = genpk_fft(ctx, which)) return rc;
        return genpk_power(ctx, which, which, nrbins, power, count, keffs, total_mass, total_mass2);
    }
    if (int rc = ensure_tables(ctx, nrbins)) return rc;
    {
        StageScope scope(ctx, ST_FFT);
        if (int rc = fft_yz(ctx, which)) return rc;
    }
    {
        StageScope scope(ctx, ST_POWER);
        bool zeroed = ctx->zero_after_power != 0;
        if (int rc = fftx_power_raw(ctx, ctx->grid[which], ctx->g.dims, 0, nrbins, ctx->d_sums, 0, &zeroed)) return rc;
        ctx->grid_clean[which] = zeroed;                   // (nothing of the grid's content survives this call either way)
    }
    GENPK_CUDA_OK(cudaMemcpyAsync(ctx->h_sums, ctx->d_sums, (size_t)3 * nrbins * sizeof(double), cudaMemcpyDeviceToHost,
                                  ctx->stream));
    GENPK_CUDA_OK(cudaStreamSynchronize(ctx->stream));
    return genpk_power_finalize(ctx->h_sums, nrbins, total_mass, total_mass2, power, count, keffs);
}

// a <- a + b, b <- a - b over two padded real grids (cross spectra by polarisation, see below)
__global__ void sum_diff_kernel(double *a, double *b, size_t n)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double x = a[i], y = b[i];
        a[i] = x + y;
        b[i] = x - y;
    }
}

// Cross spectrum of two real-space grids on the fused path.  powerspectrum() wants, per mode,
// re1*re2 + im1*im2 (powerspectrum.c:68,77,86) = (|F1 + F2|^2 - |F1 - F2|^2) / 4, and the transform is
// linear: the grids are replaced by their sum and difference, each goes through the fused auto path
// (z pass, y pass, x pass + binning in one kernel, the x-transformed spectrum never written), and the
// per-bin sums are combined on the host.  Two transforms either way; what is saved is the x pass's write
// and the binning pass's read of both spectra (64 B/mode).  Rounding: the two auto sums are accurate to
// a few ulp each, so the cross sum carries an absolute error of ~1e-15 * the auto power of the bin --
// far inside the 1e-5 relative tolerance unless the fields are uncorrelated to ten digits.
int genpk_fft_power_cross(genpk_ctx *ctx_a, int a, genpk_ctx *ctx_b, int b, int nrbins, double *power, int *count,
                          double *keffs, double total_mass, double total_mass2)
{
    if (!check_which(ctx_a, a, "genpk_fft_power_cross") || !check_which(ctx_b, b, "genpk_fft_power_cross")) return 1;
    if (nrbins < 1 || !power || !count || !keffs) { set_error("genpk_fft_power_cross: bad arguments"); return 1; }
    if (ctx_a == ctx_b && a == b)
        return genpk_fft_power(ctx_a, a, nrbins, power, count, keffs, total_mass, total_mass2);
    if (ctx_a->g.nranks != 1 || ctx_b->g.nranks != 1 || ctx_a->g.dims != ctx_b->g.dims || ctx_a->device != ctx_b->device) {
        set_error("genpk_fft_power_cross: two single-GPU grids of the same side on the same device are needed");
        return 1;
    }
    if (!fftx_supported(ctx_a, nrbins) || !fftx_supported(ctx_b, nrbins)) {
        // other grid sides: the library transform of both fields + the two-field binning pass
        if (int rc = genpk_fft(ctx_a, a)) return rc;
        if (int rc = genpk_fft(ctx_b, b)) return rc;
        return genpk_power_dev(ctx_a, ctx_a->grid[a], ctx_b->grid[b], nrbins, power, count, keffs, total_mass, total_mass2);
    }
    // everything below runs on ctx_a's stream
    cudaStream_t keep = ctx_b->stream;
    if (ctx_b != ctx_a) {
        GENPK_CUDA_OK(cudaStreamSynchronize(ctx_b->stream));
        ctx_b->stream = ctx_a->stream;
    }
    int rc = 0;
    std::vector<double> sum_pass((size_t)3 * nrbins);
    do {
        if ((rc = ensure_tables(ctx_a, nrbins)) || (rc = ensure_tables(ctx_b, nrbins))) break;
        {
            StageScope scope(ctx_a, ST_FFT);
            if ((rc = fixed_to_double(ctx_a, a)) || (rc = fixed_to_double(ctx_b, b))) break;
            sum_diff_kernel<<<ctx_a->sm_count * 8, 256, 0, ctx_a->stream>>>(ctx_a->grid[a], ctx_b->grid[b], ctx_a->g.grid_doubles());
            ctx_a->launches++;
            if ((rc = fft_yz(ctx_a, a)) || (rc = fft_yz(ctx_b, b))) break;
        }
        {
            StageScope scope(ctx_a, ST_POWER);
            bool zeroed = ctx_a->zero_after_power != 0;
            if ((rc = fftx_power_raw(ctx_a, ctx_a->grid[a], ctx_a->g.dims, 0, nrbins, ctx_a->d_sums, 0, &zeroed))) break;
            ctx_a->grid_clean[a] = zeroed;
            if (cudaMemcpyAsync(ctx_a->h_sums, ctx_a->d_sums, (size_t)3 * nrbins * sizeof(double), cudaMemcpyDeviceToHost,
                                ctx_a->stream) != cudaSuccess || cudaStreamSynchronize(ctx_a->stream) != cudaSuccess) {
                set_error("genpk_fft_power_cross: copying the sums failed: %s", cudaGetErrorString(cudaGetLastError()));
                rc = 1;
                break;
            }
            memcpy(sum_pass.data(), ctx_a->h_sums, sum_pass.size() * sizeof(double));
            // (the second pass may share ctx_a's buffers when both fields live in one context)
            double *d_sums_b = ctx_b->d_sums, *h_sums_b = ctx_b->h_sums;
            zeroed = ctx_b->zero_after_power != 0;
            if ((rc = fftx_power_raw(ctx_b, ctx_b->grid[b], ctx_b->g.dims, 0, nrbins, d_sums_b, 0, &zeroed))) break;
            ctx_b->grid_clean[b] = zeroed;
            if (cudaMemcpyAsync(h_sums_b, d_sums_b, (size_t)3 * nrbins * sizeof(double), cudaMemcpyDeviceToHost,
                                ctx_a->stream) != cudaSuccess || cudaStreamSynchronize(ctx_a->stream) != cudaSuccess) {
                set_error("genpk_fft_power_cross: copying the sums failed: %s", cudaGetErrorString(cudaGetLastError()));
                rc = 1;
                break;
            }
            for (int i = 0; i < nrbins; i++)
                sum_pass[i] = 0.25 * (sum_pass[i] - h_sums_b[i]);           // P part; sum|k| and counts are geometry
        }
    } while (0);
    ctx_b->stream = keep;
    if (rc) return rc;
    return genpk_power_finalize(sum_pass.data(), nrbins, total_mass, total_mass2, power, count, keffs);
}

int genpk_fused_xpass_supported(const genpk_ctx *ctx, int nrbins) { return ctx && fftx_supported(ctx, nrbins) ? 1 : 0; }

int genpk_pk_from_particles(genpk_ctx *ctx, const float *positions, const float *masses, int64_t n, double mass,
                            double boxsize, double total_mass, int nrbins, double *power, int *count, double *keffs)
{
    if (int rc = genpk_grid_zero(ctx, 0)) return rc;
    if (int rc = genpk_deposit(ctx, 0, positions, masses, n, mass, boxsize, 0)) return rc;
    if (int rc = genpk_fft_power(ctx, 0, nrbins, power, count, keffs, total_mass, total_mass)) return rc;
    return genpk_synchronize(ctx);
}

size_t genpk_grid_doubles(const genpk_ctx *ctx) { return ctx ? ctx->g.grid_doubles() : 0; }

size_t genpk_grid_owned_offset(const genpk_ctx *ctx) { return ctx ? ctx->g.owned_offset() : 0; }

void *genpk_grid_device_ptr(genpk_ctx *ctx, int which)
{
    if (!check_which(ctx, which, "genpk_grid_device_ptr")) return nullptr;
    if (materialize_zero(ctx, which)) return nullptr;          // the caller is about to look at the grid
    return ctx->grid[which];
}

int genpk_grid_download(genpk_ctx *ctx, int which, double *host)
{
    if (!check_which(ctx, which, "genpk_grid_download")) return 1;
    if (int rc = materialize_zero(ctx, which)) return rc;
    const size_t n = ctx->g.grid_doubles();
    GENPK_CUDA_OK(cudaMemcpyAsync(host, ctx->grid[which], n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    GENPK_CUDA_OK(cudaStreamSynchronize(ctx->stream));
    if (ctx->grid_is_fixed[which]) {                         // exact: (double)q * 2^-S
        const double inv = ldexp(1.0, -ctx->grid_scale_bits[which]);
        for (size_t i = 0; i < n; i++) {
            long long q;
            memcpy(&q, &host[i], sizeof(q));
            host[i] = (double)q * inv;
        }
    }
    return 0;
}

int genpk_grid_download_fixed(genpk_ctx *ctx, int which, int64_t *host)
{
    if (!check_which(ctx, which, "genpk_grid_download_fixed")) return 1;
    if (int rc = materialize_zero(ctx, which)) return rc;
    if (!ctx->grid_is_fixed[which] && ctx->fixed) {
        // an untouched (zeroed) grid is valid fixed-point data too
    } else if (!ctx->grid_is_fixed[which]) {
        set_error("genpk_grid_download_fixed: grid %d does not hold fixed-point sums", which);
        return 1;
    }
    GENPK_CUDA_OK(cudaMemcpyAsync(host, ctx->grid[which], ctx->g.grid_doubles() * sizeof(double), cudaMemcpyDeviceToHost,
                                  ctx->stream));
    GENPK_CUDA_OK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int genpk_grid_upload(genpk_ctx *ctx, int which, const double *host)
{
    if (!check_which(ctx, which, "genpk_grid_upload")) return 1;
    ctx->zero_pending[which] = false;                            // overwritten as a whole
    ctx->grid_clean[which] = false;
    GENPK_CUDA_OK(cudaMemcpyAsync(ctx->grid[which], host, ctx->g.grid_doubles() * sizeof(double), cudaMemcpyHostToDevice,
                                  ctx->stream));
    GENPK_CUDA_OK(cudaStreamSynchronize(ctx->stream));
    ctx->grid_is_fixed[which] = false;
    return 0;
}

int genpk_stage_ms(genpk_ctx *ctx, int stage, float *ms)
{
    if (!ctx || stage < 0 || stage >= ST_COUNT || !ms) { set_error("genpk_stage_ms: bad arguments"); return 1; }
    *ms = 0.f;
    if (ctx->ev_count[stage] == 0) return 0;
    const int s = (int)((ctx->ev_count[stage] - 1) % genpk_ctx::EV_SLOTS);
    GENPK_CUDA_OK(cudaEventSynchronize(ctx->ev_end[stage][s]));
    GENPK_CUDA_OK(cudaEventElapsedTime(ms, ctx->ev_begin[stage][s], ctx->ev_end[stage][s]));
    return 0;
}

int genpk_stage_total_ms(genpk_ctx *ctx, int stage, float *total_ms, int64_t *records)
{
    if (!ctx || stage < 0 || stage >= ST_COUNT || !total_ms || !records) { set_error("genpk_stage_total_ms: bad arguments"); return 1; }
    const int64_t n = ctx->ev_count[stage] < genpk_ctx::EV_SLOTS ? ctx->ev_count[stage] : genpk_ctx::EV_SLOTS;
    double sum = 0;
    for (int64_t i = 0; i < n; i++) {
        const int s = (int)((ctx->ev_count[stage] - 1 - i) % genpk_ctx::EV_SLOTS);
        float ms = 0.f;
        GENPK_CUDA_OK(cudaEventSynchronize(ctx->ev_end[stage][s]));
        GENPK_CUDA_OK(cudaEventElapsedTime(&ms, ctx->ev_begin[stage][s], ctx->ev_end[stage][s]));
        sum += ms;
    }
    *total_ms = (float)sum;
    *records = n;
    return 0;
}

int genpk_stage_reset(genpk_ctx *ctx)
{
    if (!ctx) { set_error("genpk_stage_reset: null context"); return 1; }
    for (int i = 0; i < ST_COUNT; i++) ctx->ev_count[i] = 0;
    return 0;
}

int64_t genpk_launch_count(const genpk_ctx *ctx) { return ctx ? ctx->launches : 0; }
int64_t genpk_library_calls(const genpk_ctx *ctx) { return ctx ? ctx->lib_calls : 0; }

int genpk_last_order(const genpk_ctx *ctx, int64_t out[7])
{
    if (!ctx || !out) { set_error("genpk_last_order: bad arguments"); return 1; }
    for (int i = 0; i < 7; i++) out[i] = ctx->last_order[i];
    return 0;
}

int genpk_grid_scale_bits(const genpk_ctx *ctx, int which)
{
    return ctx && which >= 0 && which <= 1 ? ctx->grid_scale_bits[which] : -1;
}

int genpk_last_sweep(const genpk_ctx *ctx, int64_t out[4])
{
    if (!ctx || !out) { set_error("genpk_last_sweep: bad arguments"); return 1; }
    for (int i = 0; i < 4; i++) out[i] = ctx->last_sweep[i];
    return 0;
}

/* ---------------- slab stages ---------------- */

int genpk_route_particles(genpk_ctx *ctx, const float *pos_dev, const float *mass_dev, int64_t n, double boxsize,
                          float *sorted_pos_dev, float *sorted_mass_dev, int64_t *counts_dev)
{
    if (!ctx || n < 0 || !counts_dev || (n > 0 && (!pos_dev || !sorted_pos_dev))) {
        set_error("genpk_route_particles: bad arguments");
        return 1;
    }
    return route_particles(ctx, pos_dev, mass_dev, n, boxsize, sorted_pos_dev, sorted_mass_dev, counts_dev);
}

void *genpk_ghost_side_ptr(genpk_ctx *ctx, int which, int side, size_t *bytes)
{
    if (!check_which(ctx, which, "genpk_ghost_side_ptr")) return nullptr;
    if (materialize_zero(ctx, which)) return nullptr;
    const SlabGeom &g = ctx->g;
    const int planes = side ? g.ghost_hi : g.ghost_lo;
    if (side < 0 || side > 1 || planes == 0) {
        set_error("genpk_ghost_side_ptr: this context stores no ghost planes on side %d", side);
        return nullptr;
    }
    if (bytes) *bytes = g.plane() * sizeof(double) * (size_t)planes;
    return ctx->grid[which] + (side ? g.owned_offset() + g.owned_doubles() : 0);
}

void *genpk_ghost_ptr(genpk_ctx *ctx, int which, size_t *bytes) { return genpk_ghost_side_ptr(ctx, which, 1, bytes); }

int genpk_ghost_side_accumulate(genpk_ctx *ctx, int which, int side, const void *recv_planes_dev)
{
    if (!check_which(ctx, which, "genpk_ghost_side_accumulate")) return 1;
    if (side < 0 || side > 1 || !recv_planes_dev) { set_error("genpk_ghost_side_accumulate: bad arguments"); return 1; }
    if (int rc = materialize_zero(ctx, which)) return rc;
    return ghost_accumulate(ctx, which, side, recv_planes_dev);
}

int genpk_ghost_accumulate(genpk_ctx *ctx, int which, const void *recv_plane_dev)
{
    return genpk_ghost_side_accumulate(ctx, which, 0, recv_plane_dev);
}

/* ---- ghost exchange by peer loads ---- */

int genpk_ipc_export_grid(genpk_ctx *ctx, int which, void *handle_out)
{
    if (!check_which(ctx, which, "genpk_ipc_export_grid")) return 1;
    if (!handle_out) { set_error("genpk_ipc_export_grid: null handle"); return 1; }
    cudaIpcMemHandle_t h;
    GENPK_CUDA_OK(cudaIpcGetMemHandle(&h, ctx->grid[which]));
    memcpy(handle_out, &h, sizeof(h));
    return 0;
}

int genpk_slab_set_grid_peer(genpk_ctx *ctx, int which, int side, const void *ipc_handle, void *same_process_ptr)
{
    if (!check_which(ctx, which, "genpk_slab_set_grid_peer")) return 1;
    if (side < 0 || side > 1 || (!ipc_handle && !same_process_ptr)) { set_error("genpk_slab_set_grid_peer: bad arguments"); return 1; }
    if (ctx->grid_peer_opened[which][side] && ctx->grid_peer[which][side] &&
        !(ctx->grid_peer[which][side ^ 1] == ctx->grid_peer[which][side] && ctx->grid_peer_opened[which][side ^ 1]))
        cudaIpcCloseMemHandle(ctx->grid_peer[which][side]);
    ctx->grid_peer[which][side] = nullptr;
    ctx->grid_peer_opened[which][side] = false;
    if (same_process_ptr) {
        ctx->grid_peer[which][side] = same_process_ptr;
        return 0;
    }
    cudaIpcMemHandle_t h;
    memcpy(&h, ipc_handle, sizeof(h));
    // two ranks: both neighbours are the same peer, and a handle can be opened only once per process
    if (ctx->grid_peer_opened[which][side ^ 1] && ctx->grid_peer[which][side ^ 1] &&
        memcmp(&h, ctx->grid_peer_handle[which][side ^ 1], sizeof(h)) == 0) {
        ctx->grid_peer[which][side] = ctx->grid_peer[which][side ^ 1];
        ctx->grid_peer_opened[which][side] = true;
        memcpy(ctx->grid_peer_handle[which][side], &h, sizeof(h));
        return 0;
    }
    void *p = nullptr;
    GENPK_CUDA_OK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    ctx->grid_peer[which][side] = p;
    ctx->grid_peer_opened[which][side] = true;
    memcpy(ctx->grid_peer_handle[which][side], &h, sizeof(h));
    return 0;
}

int genpk_ghost_pull_ready(const genpk_ctx *ctx, int which)
{
    if (!ctx || which < 0 || which > 1 || ctx->g.nranks < 2) return 0;
    return ctx->grid_peer[which][0] && (ctx->g.ghost_lo == 0 || ctx->grid_peer[which][1]) ? 1 : 0;
}

int genpk_ghost_pull(genpk_ctx *ctx, int which)
{
    if (!check_which(ctx, which, "genpk_ghost_pull")) return 1;
    if (!genpk_ghost_pull_ready(ctx, which)) { set_error("genpk_ghost_pull: the neighbours' grids are not mapped"); return 1; }
    if (int rc = materialize_zero(ctx, which)) return rc;
    return ghost_pull(ctx, which);
}

int genpk_slab_fft_yz(genpk_ctx *ctx, int which)
{
    if (!check_which(ctx, which, "genpk_slab_fft_yz")) return 1;
    {
        StageScope scope(ctx, ST_FFT);
        if (int rc = fft_yz(ctx, which)) return rc;
    }
    return 0;
}

int genpk_slab_pack(genpk_ctx *ctx, int which, void *send_dev)
{
    if (!check_which(ctx, which, "genpk_slab_pack")) return 1;
    if (int rc = materialize_zero(ctx, which)) return rc;
    return slab_pack(ctx, which, send_dev);
}

int genpk_slab_fft_x(genpk_ctx *ctx, void *recv_dev)
{
    if (!ctx || !recv_dev) { set_error("genpk_slab_fft_x: bad arguments"); return 1; }
    if (recv_dev == ctx->d_recv) {
        set_error("genpk_slab_fft_x: the library-owned transposed block has padded rows; use genpk_slab_fftx_power_partial");
        return 1;
    }
    return fft_x(ctx, recv_dev);
}

size_t genpk_slab_spectrum_bytes(const genpk_ctx *ctx)
{
    return ctx ? (size_t)ctx->g.dims * (ctx->g.dims / ctx->g.nranks) * ctx->g.nc * 2 * sizeof(double) : 0;
}

int genpk_slab_power_partial(genpk_ctx *ctx, const void *spec_a_dev, const void *spec_b_dev, int nrbins, double *sums_dev)
{
    if (!ctx || !spec_a_dev || !sums_dev || nrbins < 1) { set_error("genpk_slab_power_partial: bad arguments"); return 1; }
    const int ny = ctx->g.dims / ctx->g.nranks;
    {
        StageScope scope(ctx, ST_POWER);
        if (int rc = power_raw(ctx, (const double *)spec_a_dev, (const double *)(spec_b_dev ? spec_b_dev : spec_a_dev),
                               ctx->g.dims, 0, ny, ctx->g.rank * ny, nrbins, sums_dev))
            return rc;
    }
    return 0;
}

int genpk_slab_fftx_power_partial(genpk_ctx *ctx, const void *spec_yz_dev, int nrbins, double *sums_dev)
{
    if (!ctx || !spec_yz_dev || !sums_dev || nrbins < 1) { set_error("genpk_slab_fftx_power_partial: bad arguments"); return 1; }
    const int ny = ctx->g.dims / ctx->g.nranks;
    {
        StageScope scope(ctx, ST_POWER);
        // the library-owned transposed block (filled by genpk_slab_fft_yz_scatter) has padded rows
        const int pitch = spec_yz_dev == ctx->d_recv ? recv_row_pitch(ctx) : 0;
        if (int rc = fftx_power_raw(ctx, (const double *)spec_yz_dev, ny, ctx->g.rank * ny, nrbins, sums_dev, pitch)) return rc;
    }
    return 0;
}

void *genpk_slab_recv_buffer(genpk_ctx *ctx, size_t *bytes)
{
    if (!ctx) { set_error("genpk_slab_recv_buffer: null context"); return nullptr; }
    const size_t n = (size_t)ctx->g.dims * (ctx->g.dims / ctx->g.nranks) * recv_row_pitch(ctx) * 2 * sizeof(double);
    if (!ctx->d_recv && cudaMalloc(&ctx->d_recv, n) != cudaSuccess) {
        set_error("genpk_slab_recv_buffer: cudaMalloc(%zu) failed: %s", n, cudaGetErrorString(cudaGetLastError()));
        return nullptr;
    }
    if (bytes) *bytes = n;
    return ctx->d_recv;
}

int genpk_ipc_export(genpk_ctx *ctx, void *handle_out)
{
    if (!ctx || !handle_out) { set_error("genpk_ipc_export: bad arguments"); return 1; }
    if (!genpk_slab_recv_buffer(ctx, nullptr)) return 1;
    static_assert(sizeof(cudaIpcMemHandle_t) == GENPK_IPC_HANDLE_BYTES, "IPC handle size");
    cudaIpcMemHandle_t h;
    GENPK_CUDA_OK(cudaIpcGetMemHandle(&h, ctx->d_recv));
    memcpy(handle_out, &h, sizeof(h));
    return 0;
}

int genpk_slab_set_peer(genpk_ctx *ctx, int rank, const void *ipc_handle, void *same_process_ptr)
{
    if (!ctx || rank < 0 || rank >= ctx->g.nranks || rank >= GENPK_MAX_PEERS || (!ipc_handle && !same_process_ptr)) {
        set_error("genpk_slab_set_peer: bad arguments");
        return 1;
    }
    if (ctx->peer_opened[rank] && ctx->peer_recv[rank]) cudaIpcCloseMemHandle(ctx->peer_recv[rank]);
    ctx->peer_opened[rank] = false;
    ctx->peer_recv[rank] = nullptr;
    if (rank == ctx->g.rank) {
        if (!genpk_slab_recv_buffer(ctx, nullptr)) return 1;
        ctx->peer_recv[rank] = ctx->d_recv;                       // own rows: local stores
    } else if (same_process_ptr) {
        ctx->peer_recv[rank] = same_process_ptr;
    } else {
        cudaIpcMemHandle_t h;
        memcpy(&h, ipc_handle, sizeof(h));
        void *p = nullptr;
        GENPK_CUDA_OK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        ctx->peer_recv[rank] = p;
        ctx->peer_opened[rank] = true;
    }
    bool all = true;
    for (int r = 0; r < ctx->g.nranks; r++)
        all = all && ctx->peer_recv[r] != nullptr;
    ctx->peers_set = all;
    return 0;
}

int genpk_slab_scatter_supported(const genpk_ctx *ctx)
{
    if (!ctx) return 0;
    const int ny = ctx->g.dims / ctx->g.nranks;
    return fft_cols_supported(ctx) && ctx->g.nranks <= GENPK_MAX_PEERS && (ny & (ny - 1)) == 0 ? 1 : 0;
}

int genpk_slab_fft_yz_scatter(genpk_ctx *ctx, int which)
{
    if (!check_which(ctx, which, "genpk_slab_fft_yz_scatter")) return 1;
    if (!genpk_slab_scatter_supported(ctx)) { set_error("genpk_slab_fft_yz_scatter: unsupported geometry"); return 1; }
    {
        StageScope scope(ctx, ST_FFT);
        if (fft_zy_supported(ctx)) {
            if (int rc = materialize_zero(ctx, which)) return rc;
            const bool fixed = ctx->grid_is_fixed[which];
            ctx->grid_is_fixed[which] = false;
            return fft_zy(ctx, ctx->grid[which] + ctx->g.owned_offset(), ctx->g.nx, true, fixed, ctx->grid_scale_bits[which]);
        }
        if (int rc = fixed_to_double(ctx, which)) return rc;
        if (int rc = fft_z_rows(ctx, which)) return rc;
        if (int rc = fft_cols_y_scatter(ctx, ctx->grid[which] + ctx->g.owned_offset(), ctx->g.nx)) return rc;
    }
    return 0;
}

int genpk_bin_thresholds(int dims, int nrbins, unsigned flags, uint32_t *thresh_out)
{
    if (!thresh_out) { set_error("genpk_bin_thresholds: null output"); return 1; }
    BinTables t;
    if (int rc = build_bin_tables(dims, nrbins, (flags & GENPK_FLAG_BINRULE_SOURCE) ? 1u : 0u, &t)) return rc;
    memcpy(thresh_out, t.thresh.data(), t.thresh.size() * sizeof(uint32_t));
    return 0;
}

/* ---------------- reference-signature shims on host buffers ---------------- */

int genpk_fieldize(double boxsize, int dims, double *out, int64_t segment_particles, const float *positions,
                   const float *masses, double mass, int extra)
{
    if (!out || (segment_particles > 0 && !positions)) { set_error("genpk_fieldize: null buffer"); return 1; }
    return fieldize_host_shim(boxsize, dims, out, segment_particles, positions, masses, mass, extra);
}

double genpk_invwindow(int64_t kx, int64_t ky, int64_t kz, int64_t n)
{
    if (n == 0)                                              // fieldize.cpp:127
        return 0;
    const float iwx = oned_invwindow_f32(kx, n), iwy = oned_invwindow_f32(ky, n), iwz = oned_invwindow_f32(kz, n);
    const float prod = (iwx * iwy) * iwz;
    return (double)prod * (double)prod;
}

int genpk_r2c_3d(int dims, double *field)
{
    if (!field) { set_error("genpk_r2c_3d: null buffer"); return 1; }
    genpk_ctx *ctx = genpk_create(dims, -1, 0);
    if (!ctx) return 1;
    int rc = genpk_grid_upload(ctx, 0, field);
    if (!rc) rc = genpk_fft(ctx, 0);
    if (!rc) rc = genpk_grid_download(ctx, 0, field);
    genpk_destroy(ctx);
    return rc;
}

int genpk_powerspectrum(int64_t dims, const double *outfield, const double *outfield2, int nrbins, double *power,
                        int *count, double *keffs, double total_mass, double total_mass2)
{
    if (!outfield || !outfield2) { set_error("genpk_powerspectrum: null field"); return 1; }
    const bool cross = outfield2 != outfield;
    genpk_ctx *ctx = genpk_create((int)dims, -1, cross ? GENPK_FLAG_TWO_FIELDS : 0);
    if (!ctx) return 1;
    int rc = genpk_grid_upload(ctx, 0, outfield);
    if (!rc && cross) rc = genpk_grid_upload(ctx, 1, outfield2);
    if (!rc) rc = genpk_power(ctx, 0, cross ? 1 : 0, nrbins, power, count, keffs, total_mass, total_mass2);
    genpk_destroy(ctx);
    return rc;
}

}  // extern "C"
