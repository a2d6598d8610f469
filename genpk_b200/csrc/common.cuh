// Shared declarations of libgenpk_cuda.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cufft.h>
#include <nvtx3/nvToolsExt.h>
#include <stdint.h>
#include <stddef.h>
#include <string>
#include <vector>

#include "../../include/genpk_cuda.h"

namespace genpk {

void set_error(const char *fmt, ...);

#define GENPK_CUDA_OK(expr)                                                                   \
    do {                                                                                      \
        cudaError_t e__ = (expr);                                                             \
        if (e__ != cudaSuccess) {                                                             \
            genpk::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e__)); \
            return 1;                                                                         \
        }                                                                                     \
    } while (0)

#define GENPK_CUFFT_OK(expr)                                                                  \
    do {                                                                                      \
        cufftResult r__ = (expr);                                                             \
        if (r__ != CUFFT_SUCCESS) {                                                           \
            genpk::set_error("%s:%d: %s -> cufft error %d", __FILE__, __LINE__, #expr, (int)r__); \
            return 2;                                                                         \
        }                                                                                     \
    } while (0)

// Host-built tables for the binning kernel (tables.cpp).
struct BinTables {
    int dims = 0, nrbins = 0;
    unsigned rule = 0;
    std::vector<float> iw1d;          // iw1d[k] = (float) onedinvwindow(k, dims), k = 0..dims/2
    std::vector<uint32_t> thresh;     // thresh[b] = smallest k2 >= 1 whose bin is >= b; thresh[nrbins] = k2max+1
    bool monotone = true;
};
// rule 0: bin = floor((0.5*bpu)*log((double)k2))  (what gcc -O2 -ffast-math makes of powerspectrum.c:66)
// rule 1: bin = floor(bpu*log(sqrt((double)k2)))  (powerspectrum.c:66 as written)
int reference_bin_of_k2(int dims, int nrbins, int64_t k2, unsigned rule);
int build_bin_tables(int dims, int nrbins, unsigned rule, BinTables *out);
float oned_invwindow_f32(int64_t k, int64_t n);

// Geometry of the part of the grid / spectrum one context owns.
struct SlabGeom {
    int dims = 0;      // global grid side
    int nranks = 1, rank = 0;
    int nx = 0;        // local x planes (dims / nranks)
    int x0 = 0;        // first global x plane
    int ghost_lo = 0;  // ghost planes stored below the owned ones (wide-ghost slabs)
    int ghost_hi = 0;  // ghost planes stored above: 1 for plain slabs (CIC reaches one plane up), G for wide ones
    int fd = 0;        // padded z stride in doubles, 2*(dims/2+1)
    int nc = 0;        // complex z extent, dims/2+1
    size_t plane() const { return (size_t)dims * fd; }                    // doubles per x plane
    size_t grid_doubles() const { return plane() * (size_t)(ghost_lo + nx + ghost_hi); }
    size_t owned_doubles() const { return plane() * (size_t)nx; }
    size_t owned_offset() const { return plane() * (size_t)ghost_lo; }    // first owned plane inside the allocation
};

constexpr size_t GRID_TAIL_BYTES = 256;
enum Stage { ST_DEPOSIT = 0, ST_FFT = 1, ST_POWER = 2, ST_SORT = 3, ST_ZERO = 4, ST_COUNT = 5 };

}  // namespace genpk

namespace genpk {
struct DepositPlanPod {
    int mode = 0;
    long long n0 = 0, n1 = 0;
    bool have_dx = false;
    int dx_mean = 0, dx_dev = 0;
};
}  // namespace genpk

struct genpk_ctx {
    genpk::SlabGeom g;
    unsigned flags = 0;
    int device = 0;
    int sm_count = 148;
    size_t l2_bytes = 0;
    cudaStream_t stream = nullptr;
    bool fixed = false;
    int scale_bits = 40;              // GENPK_OPT_SCALE_BITS: q = llrint(w * 2^bits); -1 = chosen per grid from the particle masses
    int grid_scale_bits[2] = {40, 40};   // what each grid's sums are scaled by (latched at the first deposit after a zero)
    bool grid_scale_latched[2] = {false, false};
    float *d_maxmass = nullptr;       // scratch of the per-particle-mass maximum
    int deposit_mode = GENPK_DEPOSIT_AUTO;
    int f64_exact = 0;                // genpk_deposit_f64 uses the doubles un-narrowed (a DOUBLE_PRECISION_SNAP reference)

    // each grid allocation ends with GRID_TAIL_BYTES of bookkeeping that travels with its IPC handle:
    // int32 {lowest, highest} local x plane written since the grid was cleared (slab contexts)
    double *grid[2] = {nullptr, nullptr};
    bool grid_is_fixed[2] = {false, false};   // grid currently holds int64 fixed-point sums
    // ghost exchange by peer loads: the ring neighbours' grid allocations ([which][0]: rank-1, [1]: rank+1)
    void *grid_peer[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};
    bool grid_peer_opened[2][2] = {{false, false}, {false, false}};
    unsigned char grid_peer_handle[2][2][64] = {};

    // cuFFT
    cufftHandle plan3d = 0, plan_yz = 0, plan_x = 0, plan_z = 0;
    bool have_plan3d = false, have_plan_yz = false, have_plan_x = false, have_plan_z = false;
    int fft_yz_batch = 0;             // planes per 2-D cuFFT call (0: the whole slab in one call)
    int plan_yz_batch = 0;
    void *fft_work = nullptr;
    size_t fft_work_bytes = 0;

    // binning
    genpk::BinTables tables;
    float *d_iw1d = nullptr;
    uint32_t *d_thresh = nullptr;
    double *d_sums = nullptr;         // 3*nrbins raw sums
    double *h_sums = nullptr;         // pinned
    int sums_cap = 0;
    int power_mode = GENPK_POWER_CACHED;
    double *d_geom = nullptr;         // cached geometry sums (K, N) of the last spectrum block
    bool geom_valid = false;
    long long geom_key[5] = {0, 0, 0, 0, 0};

    // deposit scratch
    float *d_stage_pos[2] = {nullptr, nullptr};
    float *d_stage_mass[2] = {nullptr, nullptr};
    double *d_stage_pos64[2] = {nullptr, nullptr};   // double-precision positions before the narrowing kernel
    int64_t stage64_cap = 0;
    cudaEvent_t stage_free[2] = {nullptr, nullptr};
    cudaStream_t copy_stream = nullptr;
    int64_t stage_cap = 0;            // particles per staging buffer
    float *d_sorted_pos = nullptr;
    float *d_sorted_mass = nullptr;
    int64_t sorted_cap = 0;
    uint32_t *d_brick_counts = nullptr;   // histogram / cursors
    int64_t brick_cap = 0;
    unsigned long long *d_errors = nullptr;   // device-side counter of rejected particles
    void *d_order = nullptr;                  // OrderInfo written by the order probe
    long long lattice_n0 = 0, lattice_n1 = 0; // caller's hint: particles per lattice row, rows per plane
    int march_ry = 8, march_rx = 8;           // rows / planes one warp marches over
    // lattice sweep (deposit_sweep.cu)
    int sweep = 0;                            // 1: AUTO picks the sweep kernel for lattice input; 0 (default): the march kernel
                                              // (measured equal on C3, profiles/r02; the march kernel is the longer-serving one)
    int sweep_ry = 0;                         // rows per column (0: 8 in task mode; as few as keep every column resident otherwise)
    int sweep_rx = 8;                         // > 0: task mode, blocks of rx lattice planes; 0: one persistent sweep (coupled, may zero ahead)
    int zero_ahead = 1;                       // genpk_grid_zero is lazy; a sweep that follows clears the grid ahead of its front
    int za_window = 0;                        // planes ahead of the expected plane kept clear (0: from the order probe)
    int za_slack = 2;                         // lattice planes between clearing a plane and first needing it
    int za_def_per_col = 4096;                // deferred-particle list entries per list
    int za_zero_ctas = 0;                     // CTAs that only clear planes (0: a third of the SM count)
    int sweep_couple_step = 4;                // sweep warps leave an arrival mark every so many lattice planes ...
    int sweep_couple = 2;                     // ... and wait for everybody's mark `couple` marks back (0: uncoupled)
    int sweep_poll_weak = 1;                  // marks are probed with weak L1-bypassing loads
    unsigned *d_za_zdone = nullptr;
    int za_zdone_cap = 0;
    unsigned *d_za_def = nullptr;
    size_t za_def_cap = 0;
    bool grid_clean[2] = {false, false};      // the grid is all zeros already (the fused x pass left it so): genpk_grid_zero has nothing to do
    int zero_after_power = 0;                 // GENPK_OPT_ZERO_AFTER_POWER
    bool zero_pending[2] = {false, false};    // genpk_grid_zero has been asked for but not yet carried out
    long long last_sweep[4] = {0, 0, 0, 0};   // rows per column, columns, zero ahead used, window (diagnostics)

    // fused x pass (fftx_power.cu)
    int fused_xpass = 1;                      // 0: always cuFFT's x pass + bin_power_kernel
    int use_tma = 1;                          // column kernels fill their tiles with bulk tensor copies (0: per-thread cp.async)
    int fused_zy = 1;                         // GENPK_OPT_FUSED_ZY: z rows and y columns in one persistent kernel (fft_zy.cu); 2: 8192-mode tiles at 1024
    int zy_lag = 3;                           // planes between a plane's z tiles and its y tiles in that kernel's schedule
    int coop_launch = 0;                      // cudaDevAttrCooperativeLaunch
    int *d_rows_done = nullptr;
    int rows_done_n = 0;
    int own_ypass = 1;                        // 1: (y,z) transform on our own kernels (fft_zy_kernel; with fused_zy = 0: cuFFT 1-D r2c along z + fft_cols_kernel)
    int smem_optin = 0;                       // opt-in shared memory per CTA of this device
    double *d_twiddle = nullptr;              // exp(-2 pi i t/dims), t < dims
    // transpose fused into the y pass: every rank's [dims][ny][nc] block (own entry = d_recv)
    void *d_recv = nullptr;                   // library-owned transposed block of this rank
    void *peer_recv[GENPK_MAX_PEERS] = {};
    bool peer_opened[GENPK_MAX_PEERS] = {};   // entry came from cudaIpcOpenMemHandle
    bool peers_set = false;
    int twiddle_n = 0;
    long long last_order[7] = {0, 0, 0, 0, 0, 0, 0};   // last probe verdict (diagnostics)
    // the plan of the last probed stream: a deposit of the same device array (pointer, count, box) reuses it
    // instead of probing and synchronising again (a plan is only ever a performance choice)
    const void *plan_key_pos = nullptr;
    int64_t plan_key_n = 0;
    double plan_key_box = 0;
    int plan_key_mode = -1;
    genpk::DepositPlanPod plan_cached;

    // timing: a ring of event pairs per stage, summed on request (no host sync while recording)
    static constexpr int EV_SLOTS = 128;
    cudaEvent_t ev_begin[genpk::ST_COUNT][EV_SLOTS] = {}, ev_end[genpk::ST_COUNT][EV_SLOTS] = {};
    int64_t ev_count[genpk::ST_COUNT] = {};   // records since the last reset
    int64_t launches = 0;     // our own kernels
    int64_t lib_calls = 0;    // cuFFT executions
};

namespace genpk {

// deposit.cu
struct DepositPlan {
    int mode = 0;                 // GENPK_DEPOSIT_DIRECT / SORTED / MARCH / SWEEP
    long long n0 = 0, n1 = 0;     // lattice row length / rows per plane (MARCH, SWEEP)
    bool have_dx = false;         // the order probe measured where lattice planes sit along x
    int dx_mean = 0, dx_dev = 0;
};
int deposit_plan(genpk_ctx *ctx, const float *pos, int64_t n, double boxsize, DepositPlan *plan);
// plan == nullptr: planned here (one order probe, one small D2H)
int deposit_device(genpk_ctx *ctx, int which, const float *pos, const float *masses, int64_t n,
                   double mass, double boxsize, const DepositPlan *plan = nullptr);
int deposit_device_f64(genpk_ctx *ctx, int which, const double *pos, const float *masses, int64_t n, double mass,
                       double boxsize);
int fixed_to_double(genpk_ctx *ctx, int which);
// fixed-point mode: the scale of grid `which` (latched at the first deposit after a zero; automatic scales
// look at the masses of that deposit)
int latch_scale(genpk_ctx *ctx, int which, const float *masses_dev, int64_t n, double mass);
// carries out a pending genpk_grid_zero (every reader of the grid calls this first)
int materialize_zero(genpk_ctx *ctx, int which);
// binpower.cu
int ensure_tables(genpk_ctx *ctx, int nrbins);
int power_raw(genpk_ctx *ctx, const double *spec_a, const double *spec_b, int n_outer, int outer0,
              int n_mid, int mid0, int nrbins, double *sums_dev);
int power_seed_sums(genpk_ctx *ctx, int n_outer, int outer0, int n_mid, int mid0, int nrbins, double *sums_dev);
// fftx_power.cu
bool fftx_supported(const genpk_ctx *ctx, int nrbins);
// zero_after (in: wanted, out: done): the block is overwritten with zeros as it is read
int fftx_power_raw(genpk_ctx *ctx, const double *spec_yz, int n_mid, int mid0, int nrbins, double *sums_dev, int row_pitch = 0,
                   bool *zero_after = nullptr);
int recv_row_pitch(const genpk_ctx *ctx);
bool fft_cols_supported(const genpk_ctx *ctx);
int fft_cols_y(genpk_ctx *ctx, double *spec, int n_planes);
int fft_cols_y_scatter(genpk_ctx *ctx, double *spec, int n_planes);
bool fft_zy_supported(const genpk_ctx *ctx);
int fft_zy(genpk_ctx *ctx, double *planes, int n_planes, bool scatter, bool from_fixed, int scale_bits);
int fft_z_rows(genpk_ctx *ctx, int which);
// fft.cu
int fft_3d(genpk_ctx *ctx, int which);
int fft_yz(genpk_ctx *ctx, int which);
int fft_x(genpk_ctx *ctx, void *recv);
void fft_release(genpk_ctx *ctx);
// slab.cu
int route_particles(genpk_ctx *ctx, const float *pos, const float *mass, int64_t n, double boxsize,
                    float *spos, float *smass, int64_t *counts);
int ghost_accumulate(genpk_ctx *ctx, int which, int side, const void *recv);
int ghost_pull(genpk_ctx *ctx, int which);
int touched_set(genpk_ctx *ctx, int which, int lo, int hi);      // stream-ordered write of the touched-plane range
inline int *touched_ptr(genpk_ctx *ctx, int which)
{
    return reinterpret_cast<int *>(ctx->grid[which] + ctx->g.grid_doubles());
}
int slab_pack(genpk_ctx *ctx, int which, void *send);

// A stage = one CUDA-event pair on the context's stream (summed on request) and one NVTX range on the
// calling thread (genpk:deposit, genpk:fft, ... in a timeline; header-only NVTX, a no-op without a tool).
inline const char *stage_name(int st)
{
    static const char *const names[ST_COUNT] = {"genpk:deposit", "genpk:fft", "genpk:power", "genpk:sort", "genpk:zero"};
    return st >= 0 && st < ST_COUNT ? names[st] : "genpk:?";
}
inline void stage_begin(genpk_ctx *ctx, int st)
{
    nvtxRangePushA(stage_name(st));
    cudaEventRecord(ctx->ev_begin[st][ctx->ev_count[st] % genpk_ctx::EV_SLOTS], ctx->stream);
}
inline void stage_end(genpk_ctx *ctx, int st)
{
    cudaEventRecord(ctx->ev_end[st][ctx->ev_count[st] % genpk_ctx::EV_SLOTS], ctx->stream);
    ctx->ev_count[st]++;
    nvtxRangePop();
}

// scope form: the range and the event pair close on every return path
struct StageScope {
    genpk_ctx *ctx;
    int st;
    StageScope(genpk_ctx *c, int s) : ctx(c), st(s) { stage_begin(ctx, st); }
    ~StageScope() { stage_end(ctx, st); }
    StageScope(const StageScope &) = delete;
    StageScope &operator=(const StageScope &) = delete;
};

}  // namespace genpk
