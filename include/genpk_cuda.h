/* genpk_cuda.h -- C ABI of libgenpk_cuda.so, the B200 (sm_100a) implementation
 * of GenPK's matter-power-spectrum hot path:
 *
 *     cloud-in-cell deposit  ->  3-D r2c FFT  ->  |delta_k|^2 binning
 *
 * It replaces, behind the reference's own link-time function boundary
 * (gen-pk.h:93-119), what happens between the memset at gen-pk.cpp:208 and the
 * return of powerspectrum() at gen-pk.cpp:234.  All entry points are
 * extern "C", take plain pointers and sizes, and return 0 on success and a
 * nonzero code on failure (genpk_last_error() has the message); nothing here
 * aborts the process and nothing falls back to the CPU.
 *
 * Three layers:
 *   1. reference-signature shims on HOST buffers (drop-in for the callers in
 *      read_fieldize*.cpp, gen-pk.cpp and test.cpp);
 *   2. a handle API that keeps the grid resident in HBM across
 *      zero -> deposit* -> fft -> power (the fast path the host CLI uses);
 *   3. slab-stage entry points on DEVICE buffers for the multi-GPU pipeline,
 *      where the caller owns the inter-GPU exchange steps.
 *
 * Layout conventions (same as the reference / FFTW in-place r2c):
 *   real grid   double [dims][dims][fd],  fd = 2*(dims/2+1), x slowest, z fastest
 *   spectrum    complex double [dims][dims][dims/2+1] in the same bytes
 *   particles   float32 AoS [n][3]; optional float32 masses [n]
 */
#ifndef GENPK_CUDA_H
#define GENPK_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GENPK_ABI_VERSION 1

/* ---- flags for genpk_create / genpk_create_slab ---------------------------------- */
#define GENPK_FLAG_FIXED_POINT   0x1u  /* deterministic int64 fixed-point accumulation */
#define GENPK_FLAG_TWO_FIELDS    0x2u  /* allocate a second grid (cross spectra, gen-pk.cpp:263,314)   */
#define GENPK_FLAG_BINRULE_SOURCE 0x4u /* bin = floor(b*log(sqrt(k2))) evaluated as written in
                                          powerspectrum.c:66 instead of the 0.5*b*log(k2) form that
                                          gcc -ffast-math (the reference's Makefile:30) compiles it to */

/* ---- deposit algorithm selection (genpk_set_option(ctx, GENPK_OPT_DEPOSIT, v)) ---- */
#define GENPK_OPT_DEPOSIT        1
#define GENPK_DEPOSIT_AUTO       0     /* pick from particle density and spatial coherence */
#define GENPK_DEPOSIT_DIRECT     1     /* one thread per particle, 8 global red.add          */
#define GENPK_DEPOSIT_SORTED     2     /* counting sort into L2-sized bricks, then deposit   */
#define GENPK_DEPOSIT_TILED      3     /* reserved (alias of AUTO)                              */
#define GENPK_DEPOSIT_MARCH      4     /* lattice-ordered input: neighbour contributions merged in
                                          registers along z, y and x before one red.add per particle */
#define GENPK_DEPOSIT_SWEEP      5     /* lattice-ordered input: persistent warps sweep the lattice along x; merges in
                                          registers (y), shared memory (x) and one shuffle (z); as one persistent wave
                                          (GENPK_OPT_SWEEP_RX = 0) it can clear the grid ahead of its own front.  20 % fewer
                                          instructions than MARCH and the same time: AUTO takes MARCH */
#define GENPK_OPT_SCALE_BITS     2     /* fixed-point mode: q = llrint(w * 2^bits), default 40.  -1: chosen per grid at the first
                                          deposit after genpk_grid_zero as 40 - ceil(log2(largest particle mass of that
                                          deposit)), which keeps ~40 significant bits per contribution whatever the mass
                                          unit and 2^23 such particles per cell of headroom (single-GPU contexts; a
                                          multi-GPU handle picks one scale for all slabs).  genpk_grid_scale_bits() tells */
/* Lattice hint for GENPK_DEPOSIT_MARCH / AUTO: particle p sits near lattice site
 * (ix,iy,iz) with p = (ix*N1 + iy)*N0 + iz.  0 = let the order probe find it.  Only
 * ever a performance hint: results do not depend on it. */
#define GENPK_OPT_LATTICE_N0     4
#define GENPK_OPT_LATTICE_N1     5
#define GENPK_OPT_MARCH_RY       6     /* lattice rows one warp marches over (default 8)   */
#define GENPK_OPT_MARCH_RX       7     /* lattice planes one warp marches over (default 8) */
#define GENPK_OPT_SWEEP         11     /* 1: AUTO uses the sweep kernel for lattice input; 0 (default): the march kernel */
#define GENPK_OPT_SWEEP_RY      12     /* lattice rows per sweep column (0 = as few as keep all columns resident) */
#define GENPK_OPT_ZERO_AHEAD    13     /* 1 (default): genpk_grid_zero is carried out lazily -- by a memset when the grid is next used, or
                                          by a persistent sweep deposit itself (first-touch zeroing ahead of its front, only
                                          with GENPK_DEPOSIT_SWEEP and GENPK_OPT_SWEEP_RX = 0); 0: memset at once */
#define GENPK_OPT_ZA_WINDOW     14     /* zero ahead: grid planes past a lattice plane's expected position that are kept
                                          clear (0 = from the displacements the order probe saw).  Performance only:
                                          particles displaced further are deposited by a clean-up pass */
#define GENPK_OPT_ZA_SLACK      15     /* zero ahead: lattice planes between clearing a grid plane and first needing it */
#define GENPK_OPT_ZA_DEFERRED   16     /* zero ahead: entries per clean-up list (1024 lists; default 4096) */
#define GENPK_OPT_ZA_ZERO_CTAS  17     /* zero ahead: CTAs of the sweep launch that only clear planes (0 = a third of the SMs) */
#define GENPK_OPT_SWEEP_COUPLE  18     /* sweep warps leave an arrival mark every STEP lattice planes and go on only when every
                                          warp has left the mark N marks back (default 2; 0 = uncoupled): keeps the front of
                                          the sweep about N*STEP planes thick, which is what keeps the reductions in L2 */
#define GENPK_OPT_SWEEP_COUPLE_STEP 19 /* STEP above (default 4) */
#define GENPK_OPT_SWEEP_RX      21     /* N > 0 (default 8): the sweep runs as independent tasks of N lattice planes x one block of
                                          rows x one segment, in launch order (no waiting; the grid is cleared by a memset);
                                          0: one persistent sweep over all planes, coupled, which can clear the grid ahead of itself */
#define GENPK_OPT_SWEEP_POLL_WEAK 20   /* 1 (default): marks are probed with weak L1-bypassing loads; 0: relaxed.gpu loads */
#define GENPK_OPT_F64_POSITIONS 23     /* genpk_deposit_f64: 0 (default) narrows the doubles to float first, as read_fieldize_bigfile.cpp
                                          :93-94 does; 1 uses them as they are, which is what a reference built with
                                          -DDOUBLE_PRECISION_SNAP (gen-pk.h:25-29, read_fieldize.cpp:24-25) hands to fieldize() */
/* ---- binning pass selection (genpk_set_option(ctx, GENPK_OPT_POWER, v)) ------------ */
#define GENPK_OPT_POWER          3
#define GENPK_POWER_CACHED       0     /* sum|k| and mode counts per bin depend on the grid only: computed
                                          on the GPU once per (dims, nrbins) and cached in the context;
                                          the per-spectrum pass accumulates P alone (default)          */
#define GENPK_POWER_FUSED        1     /* P, sum|k| and counts in one pass every call                  */
#define GENPK_OPT_FFT_YZ_BATCH   9     /* x planes per 2-D cuFFT call of the (y,z) transform; groups of a few
                                          planes keep the z pass's output in L2 for the y pass (0 = all
                                          planes in one call; must divide the local plane count) */
#define GENPK_OPT_OWN_YPASS     10     /* 1 (default): the (y,z) transform (genpk_slab_fft_yz, genpk_fft_power) runs on our own kernels
                                          for grid sides 256/512/1024/2048 -- the fused kernel of GENPK_OPT_FUSED_ZY, or with
                                          that option off cuFFT's 1-D r2c along z + our in-place column pass along y;
                                          0: cuFFT's 2-D plan */
/* ---- fused x pass (genpk_fft_power, genpk_slab_fftx_power_partial) ------------------ */
#define GENPK_OPT_FUSED_XPASS    8     /* 1 (default): last FFT pass and binning in one kernel when the
                                          grid side allows; 0: always cuFFT's x pass + the binning pass;
                                          2: as 1 with 4096-mode tiles at 1024 (measurements);
                                          3: at 1024, the CTA as two halves of 256 threads out of step on one 8192-mode tile;
                                          4: at 1024, the two-pass plan (32 modes per thread, one exchange, 256 threads) */

#define GENPK_OPT_TMA           24     /* 1 (default): the column kernels (y pass, fused x pass) fill their shared-memory tiles with
                                          TMA bulk tensor copies behind an mbarrier; 0: per-thread cp.async (measurements) */
#define GENPK_OPT_FUSED_ZY      26     /* 1 (default): the (y,z) part of the transform is ONE persistent kernel -- r2c along z on the rows
                                          of a plane, then the column pass along y two planes behind, out of L2 (fft_zy.cu): the grid
                                          is read once and written once.  2: the same with 8192-mode tiles at 1024.  0: the library's
                                          batched 1-D r2c along z followed by fft_cols_kernel along y (two trips through HBM). */
#define GENPK_OPT_ZY_LAG        27     /* planes between a plane's z tiles and its y tiles in that kernel's schedule (default 3) */
#define GENPK_OPT_ZERO_AFTER_POWER 25  /* 1: the fused x pass of genpk_fft_power / genpk_fft_power_cross / genpk_pk_from_particles
                                          overwrites every tile of the grid with zeros (TMA bulk tensor stores) right after reading it,
                                          and the genpk_grid_zero that follows has nothing left to do; needs GENPK_OPT_TMA.  0 (default):
                                          measured at 1024^3 the stores cost the x pass 1.45 ms and the memset they replace 1.19 ms
                                          (profiles/r02/README.md).  The grid content after those calls is unspecified either way. */

typedef struct genpk_ctx genpk_ctx;
#define GENPK_MAX_PEERS 16
#define GENPK_IPC_HANDLE_BYTES 64

/* ================= 1. reference-signature shims (host buffers) ====================== */

/* Replaces fieldize() (fieldize.cpp:46, gen-pk.h:93).  Same arguments, same
 * meaning: `out` is a caller-owned, caller-zeroed host grid that is ACCUMULATED
 * into; extra=1 selects the FFTW padded z stride 2*(dims/2+1), extra=0 the
 * unpadded stride 2*(dims/2).  Copies `out` up, deposits on the GPU, copies it
 * back.  Returns 0 (the reference always returns 0) or nonzero on a CUDA error. */
int genpk_fieldize(double boxsize, int dims, double *out, int64_t segment_particles,
                   const float *positions, const float *masses, double mass, int extra);

/* Replaces invwindow() (fieldize.cpp:125, gen-pk.h:104): ((float)wx*(float)wy*(float)wz)^2,
 * 0 when n==0.  Host arithmetic, identical roundings to the table the binning
 * kernel consumes; kept so test.cpp:52-57 can be run against this library. */
double genpk_invwindow(int64_t kx, int64_t ky, int64_t kz, int64_t n);

/* Replaces fftw_plan_dft_r2c_3d(d,d,d,field,(fftw_complex*)field,FFTW_ESTIMATE) +
 * fftw_execute() (gen-pk.cpp:193,233): unnormalised forward in-place r2c on a
 * host buffer of 2*d*d*(d/2+1) doubles, computed with cuFFT on the GPU. */
int genpk_r2c_3d(int dims, double *field);

/* Replaces powerspectrum() (powerspectrum.c:35, gen-pk.h:119).  outfield and
 * outfield2 are host arrays of dims*dims*(dims/2+1) interleaved complex doubles
 * and may alias.  power/count/keffs (nrbins each) are overwritten. */
int genpk_powerspectrum(int64_t dims, const double *outfield, const double *outfield2, int nrbins,
                        double *power, int *count, double *keffs, double total_mass, double total_mass2);

/* ================= 2. handle API (grid resident in HBM) ============================= */

/* One context = one GPU, one cubic grid of side `dims` (or one x-slab of it),
 * its cuFFT plans and its scratch.  device<0 keeps the current device. */
genpk_ctx *genpk_create(int dims, int device, unsigned flags);
void genpk_destroy(genpk_ctx *ctx);
const char *genpk_last_error(void);
int genpk_abi_version(void);

/* All work of the context is issued on `cuda_stream` (a cudaStream_t; NULL =
 * the legacy default stream).  Lets a host that owns streams (e.g. torch) order
 * our kernels against its own copies and collectives. */
int genpk_set_stream(genpk_ctx *ctx, void *cuda_stream);
int genpk_set_option(genpk_ctx *ctx, int option, int64_t value);
int genpk_synchronize(genpk_ctx *ctx);

/* memset(field,0,...) of gen-pk.cpp:208 for grid `which` (0 or 1).  Carried out lazily (see
 * GENPK_OPT_ZERO_AHEAD): observable behaviour is that of an immediate memset. */
int genpk_grid_zero(genpk_ctx *ctx, int which);

/* fieldize() into the resident grid `which`; additive across calls until the
 * next genpk_grid_zero (chunk loop read_fieldize.cpp:51-93; stars into baryons
 * gen-pk.cpp:228-230).  positions/masses are host pointers when on_device==0 or device
 * pointers when on_device!=0.  Host arrays go up in chunks of 2^23 particles straight from
 * the caller's memory on a copy stream, double-buffered on the device so that chunk k+1
 * uploads while chunk k is deposited (pin the arrays -- cudaHostAlloc / cudaHostRegister --
 * to get the overlap; pageable memory is staged by the driver); every upload has completed
 * when the call returns, so the arrays may be reused at once.  masses may be NULL (constant
 * `mass`). */
int genpk_deposit(genpk_ctx *ctx, int which, const float *positions, const float *masses,
                  int64_t n, double mass, double boxsize, int on_device);

/* The same deposit for double-precision positions (bigfile `Position` blocks of dtype f8,
 * DOUBLE_PRECISION_SNAP Gadget files): narrowed to float on the GPU exactly as
 * read_fieldize_bigfile.cpp:93-94 does on the host (`positions[i] = ((double *)pos)[i]`, round
 * to nearest), then deposited; identical grids to narrowing on the host and calling
 * genpk_deposit.  Masses stay float (read_fieldize_bigfile.cpp:98-118). */
int genpk_deposit_f64(genpk_ctx *ctx, int which, const double *positions, const float *masses,
                      int64_t n, double mass, double boxsize, int on_device);

/* fftw_execute() of gen-pk.cpp:233 on grid `which`, in place (cuFFT D2Z).  In
 * fixed-point mode the int64 grid is converted to double first. */
int genpk_fft(genpk_ctx *ctx, int which);

/* powerspectrum(dims, grid a, grid b, ...) (gen-pk.cpp:234,297,348) on the
 * resident spectra; results to host arrays of nrbins entries. */
int genpk_power(genpk_ctx *ctx, int a, int b, int nrbins, double *power, int *count, double *keffs,
                double total_mass, double total_mass2);

/* The same binning on two device-resident spectra [dims][dims][dims/2+1] that need
 * not be this context's own grids (e.g. the spectrum of another context on the same
 * GPU: the two-snapshot cross spectrum of gen-pk.cpp:295-297).  spec_b_dev NULL = auto. */
int genpk_power_dev(genpk_ctx *ctx, const void *spec_a_dev, const void *spec_b_dev, int nrbins,
                    double *power, int *count, double *keffs, double total_mass, double total_mass2);

/* fftw_execute() + powerspectrum() of gen-pk.cpp:233-234 as one call on grid `which` (auto
 * spectrum).  For grid sides 256/512/1024/2048 the batched 2-D (y,z) cuFFT transform is
 * followed by ONE kernel that does the x transform and the binning, so the x-transformed
 * spectrum is never written to memory: afterwards the grid holds the (y,z)-transformed
 * planes, not the 3-D spectrum (call genpk_fft + genpk_power when the spectrum itself is
 * wanted).  Other grid sides run genpk_fft + genpk_power.  Same results as those two calls
 * up to the summation order inside a bin. */
int genpk_fft_power(genpk_ctx *ctx, int which, int nrbins, double *power, int *count, double *keffs,
                    double total_mass, double total_mass2);
/* fftw_execute() of both fields + powerspectrum(dims, field a, field b, ...) of gen-pk.cpp:295-297 / 345-348 as
 * one call: the cross spectrum of grid `a` of ctx_a and grid `b` of ctx_b (the same context with
 * GENPK_FLAG_TWO_FIELDS, or two contexts on one device -- the two-snapshot mode).  On the fused grid sides
 * the two real grids are replaced by their sum and difference and each goes through the fused auto path:
 * re1*re2 + im1*im2 = (|F1+F2|^2 - |F1-F2|^2)/4 per mode, so neither x-transformed spectrum is ever written.
 * Afterwards the grids hold (y,z)-transformed planes of the sum and the difference.  Other grid sides:
 * genpk_fft on both + the two-field binning pass. */
int genpk_fft_power_cross(genpk_ctx *ctx_a, int a, genpk_ctx *ctx_b, int b, int nrbins, double *power, int *count,
                          double *keffs, double total_mass, double total_mass2);
/* 1 when genpk_fft_power / genpk_slab_fftx_power_partial take the fused path for this context. */
int genpk_fused_xpass_supported(const genpk_ctx *ctx, int nrbins);

/* Whole per-type step of gen-pk.cpp:208-234 in one call on host particle
 * arrays: zero, deposit, FFT, binning, results on the host. */
int genpk_pk_from_particles(genpk_ctx *ctx, const float *positions, const float *masses, int64_t n,
                            double mass, double boxsize, double total_mass, int nrbins,
                            double *power, int *count, double *keffs);

/* Parity / debugging: copy grid `which` to/from a host buffer of
 * genpk_grid_doubles(ctx) doubles (real padded grid before genpk_fft, the
 * interleaved spectrum after it).  In fixed-point mode before genpk_fft the
 * download converts to double; genpk_grid_download_fixed returns the raw int64. */
size_t genpk_grid_doubles(const genpk_ctx *ctx);
int genpk_grid_download(genpk_ctx *ctx, int which, double *host);
int genpk_grid_upload(genpk_ctx *ctx, int which, const double *host);
int genpk_grid_download_fixed(genpk_ctx *ctx, int which, int64_t *host);
void *genpk_grid_device_ptr(genpk_ctx *ctx, int which);

/* Elapsed GPU milliseconds of the most recent deposit / fft / power stage of
 * this context (CUDA events on the context's stream), and kernels launched
 * since creation. */
#define GENPK_STAGE_DEPOSIT 0
#define GENPK_STAGE_FFT     1
#define GENPK_STAGE_POWER   2
#define GENPK_STAGE_SORT    3   /* part of DEPOSIT: the brick sort */
#define GENPK_STAGE_ZERO    4   /* genpk_grid_zero */
int genpk_stage_ms(genpk_ctx *ctx, int stage, float *ms);
/* Sum over the (up to 128 most recent) recorded instances of a stage since
 * genpk_stage_reset, and how many were summed.  Recording never synchronises
 * the host; this call does. */
int genpk_stage_total_ms(genpk_ctx *ctx, int stage, float *total_ms, int64_t *records);
int genpk_stage_reset(genpk_ctx *ctx);
int64_t genpk_launch_count(const genpk_ctx *ctx);     /* kernels of this library launched so far */
int64_t genpk_library_calls(const genpk_ctx *ctx);    /* cuFFT executions so far (0 on the default 256..2048 paths) */
/* Verdict of the last order probe of this context: {coherent, lattice, n0, n1,
 * score_z, score_y, score_x} (scores per mille; diagnostics for the bench line). */
int genpk_grid_scale_bits(const genpk_ctx *ctx, int which);    /* the scale the grid's fixed-point sums carry now */
int genpk_last_order(const genpk_ctx *ctx, int64_t out[7]);
/* The last sweep deposit of this context: {rows per column, columns (= warps), zero ahead used (0/1),
 * zero-ahead window in planes}. */
int genpk_last_sweep(const genpk_ctx *ctx, int64_t out[4]);

/* ================= 3. slab stages for the multi-GPU pipeline ======================== */
/* Rank `rank` of `nranks` owns x-planes [rank*dims/nranks, (rank+1)*dims/nranks)
 * of the real grid plus one ghost plane on the high-x side, and after the
 * transpose the ky rows [rank*dims/nranks, ...) of the spectrum.  dims must be
 * divisible by nranks.  The caller (one process per GPU) performs the three
 * exchange steps between the stages with its own collectives:
 *
 *   genpk_route_particles   -> all-to-all-v of particle runs      (caller)
 *   genpk_deposit(on_device)
 *   genpk_ghost_ptr         -> ring shift of the ghost plane      (caller)
 *   genpk_ghost_accumulate
 *   genpk_slab_fft_yz, genpk_slab_pack
 *                           -> all-to-all of the packed blocks    (caller)
 *   genpk_slab_fft_x
 *   genpk_slab_power_partial-> all-reduce of 3*nrbins doubles     (caller)
 *   genpk_power_finalize
 */
genpk_ctx *genpk_create_slab(int dims, int device, int nranks, int rank, unsigned flags);

/* Wide-ghost slab: ghost_planes planes are stored on BOTH sides of the owned ones
 * (ghost_planes <= dims/nranks).  A rank whose particle shard is already slab-local
 * up to stragglers -- snapshot / lattice order sharded by index range -- deposits it
 * without any particle exchange: contributions up to ghost_planes planes outside the
 * slab land in the ghosts and travel with the two-sided ghost exchange below.
 * Particles further out are counted by genpk_take_rejected() and NOT deposited; the
 * caller then falls back to genpk_route_particles.  ghost_planes = 0 is genpk_create_slab. */
genpk_ctx *genpk_create_slab_wide(int dims, int device, int nranks, int rank, unsigned flags, int ghost_planes);
/* Doubles between the start of the grid allocation (genpk_grid_device_ptr,
 * genpk_grid_download) and the first owned plane: ghost_planes * dims * fd. */
size_t genpk_grid_owned_offset(const genpk_ctx *ctx);
/* Waits for the context's stream, then returns and clears the number of particles the
 * deposits rejected (outside the slab + ghosts, or non-finite). */
int genpk_take_rejected(genpk_ctx *ctx, uint64_t *rejected);

/* Destination rank of every particle: floor(x*dims/box) wrapped, / (dims/nranks).
 * Writes counts[nranks] (device int64) and the particles grouped by destination
 * into sorted_pos (device, n*3 floats) / sorted_mass (or NULL). */
int genpk_route_particles(genpk_ctx *ctx, const float *pos_dev, const float *mass_dev, int64_t n,
                          double boxsize, float *sorted_pos_dev, float *sorted_mass_dev,
                          int64_t *counts_dev);

/* Device pointer and byte size of the ghost plane (send buffer of the ring
 * shift), and accumulation of a received plane into local plane 0. */
void *genpk_ghost_ptr(genpk_ctx *ctx, int which, size_t *bytes);
int genpk_ghost_accumulate(genpk_ctx *ctx, int which, const void *recv_plane_dev);
/* The same for either side of a (wide-ghost) slab.  side 1 = high-x ghosts, sent to rank+1,
 * which adds them into its first owned planes with accumulate(side 0); side 0 = low-x ghosts,
 * sent to rank-1, which adds them into its last owned planes with accumulate(side 1). */
void *genpk_ghost_side_ptr(genpk_ctx *ctx, int which, int side, size_t *bytes);
int genpk_ghost_side_accumulate(genpk_ctx *ctx, int which, int side, const void *recv_planes_dev);

/* ---- ghost exchange by peer loads (no collective, no staging buffers) ------------------------
 * Every grid allocation can be exported with CUDA IPC (genpk_ipc_export_grid) and mapped by the two
 * ring neighbours (genpk_slab_set_grid_peer: side 0 = the grid of rank-1, side 1 = of rank+1; plain
 * pointers inside one process).  Once every rank has deposited (the caller's barrier),
 * genpk_ghost_pull adds the neighbours' ghost planes into this rank's outermost owned planes,
 * reading them straight from the neighbours' memory over NVLink -- and only the planes their
 * deposits wrote (each allocation carries the {lowest, highest} plane written since it was cleared;
 * the sweep kernel tracks it, the other deposit kernels mark every plane).  Replaces
 * genpk_ghost_side_ptr + send/recv + genpk_ghost_side_accumulate.  The neighbours must not clear or
 * deposit into their grids again before every pull of the step is over (one more barrier, e.g. the one
 * that precedes the transpose). */
int genpk_ipc_export_grid(genpk_ctx *ctx, int which, void *handle_out /* GENPK_IPC_HANDLE_BYTES */);
int genpk_slab_set_grid_peer(genpk_ctx *ctx, int which, int side, const void *ipc_handle /* or NULL */,
                             void *same_process_ptr);
int genpk_ghost_pull_ready(const genpk_ctx *ctx, int which);
int genpk_ghost_pull(genpk_ctx *ctx, int which);
/* Stream-ordered: writes the number of particles the deposits rejected so far (as a double) to
 * dst_dev and clears the counter -- lets a pipeline carry the count along with its per-bin sums and
 * look at it once per step on the host instead of synchronising after every deposit. */
int genpk_rejected_to(genpk_ctx *ctx, double *dst_dev);

/* Batched 2-D D2Z over the local x-planes (in place). */
int genpk_slab_fft_yz(genpk_ctx *ctx, int which);
/* Reorder [x_local][y][kz] into nranks contiguous blocks [dest][x_local][y_local][kz]. */
int genpk_slab_pack(genpk_ctx *ctx, int which, void *send_dev);
/* 1-D Z2Z along x on the received [x][y_local][kz] array (in place in recv_dev). */
int genpk_slab_fft_x(genpk_ctx *ctx, void *recv_dev);
size_t genpk_slab_spectrum_bytes(const genpk_ctx *ctx);

/* Raw per-bin sums of this rank's part of the spectrum: sums_dev holds
 * 3*nrbins doubles = [sum w*P*W^2][sum w*|k|][sum w] (counts are exact integers
 * below 2^53).  spec_a/spec_b are [dims][ny_local][dims/2+1] device arrays. */
int genpk_slab_power_partial(genpk_ctx *ctx, const void *spec_a_dev, const void *spec_b_dev, int nrbins,
                             double *sums_dev);

/* genpk_slab_fft_x + genpk_slab_power_partial in one kernel (auto spectrum): spec_yz_dev is
 * the received [dims][ny_local][dims/2+1] array BEFORE the x transform and is left
 * untouched.  Fails (nonzero) when genpk_fused_xpass_supported() is 0. */
int genpk_slab_fftx_power_partial(genpk_ctx *ctx, const void *spec_yz_dev, int nrbins, double *sums_dev);

/* ---- transpose fused into the y pass (peer stores over NVLink) -------------------------
 * Each rank owns one transposed block [dims][ny_local][pitch], pitch = dims/2+1 rounded up to 8 modes
 * so that rows start on 128-byte boundaries (genpk_slab_recv_buffer, allocated by the library so that
 * it can be shared; only genpk_slab_fft_yz_scatter writes it and genpk_slab_fftx_power_partial reads it).  Once every rank knows every block's
 * address -- genpk_ipc_export / genpk_ipc_open between processes, plain pointers inside one
 * process -- genpk_slab_fft_yz_scatter does the z pass, then the y pass whose results are
 * stored straight into their owner's block: no pack kernel, no all-to-all.  The caller
 * orders the ranks: nobody may still read its block when the scatter starts, and everybody
 * must have finished scattering before genpk_slab_fftx_power_partial reads a block (two barriers, e.g. 1-element all-reduces on the stream).
 * Needs grid side 256/512/1024/2048, dims/nranks a power of two, nranks <= GENPK_MAX_PEERS. */
void *genpk_slab_recv_buffer(genpk_ctx *ctx, size_t *bytes);
int genpk_ipc_export(genpk_ctx *ctx, void *handle_out /* GENPK_IPC_HANDLE_BYTES */);
int genpk_slab_set_peer(genpk_ctx *ctx, int rank, const void *ipc_handle /* or NULL */, void *same_process_ptr);
int genpk_slab_scatter_supported(const genpk_ctx *ctx);
int genpk_slab_fft_yz_scatter(genpk_ctx *ctx, int which);

/* ================= 4. N GPUs of one box behind one handle ========================================
 * The per-particle-type loop of gen-pk.cpp:202-239 on N GPUs, single process, no collective library: the grid is
 * slab-decomposed in x (dims divisible by ngpus); host particles in ANY order are split N ways, every GPU uploads
 * its part over its own PCIe link and groups it by owner slab on the device, the runs move with peer copies over
 * NVLink, each GPU deposits what it owns; the ghost plane is pulled from the neighbour's memory, the FFT transpose
 * is stored straight into the owners' blocks by the y pass (grid sides 256..2048, dims/ngpus a power of two; peer
 * copies of packed blocks otherwise), the last FFT pass is fused with the binning, and the per-bin partial sums are
 * added on the host.  devices: ngpus device ordinals, or NULL for the visible GPUs spread evenly (4 of 8: 0, 2, 4, 6;
 * wrapping when there are fewer than ngpus: several slabs may share a device).  Same results as the single-GPU handle API (counts bit-exact; the fixed-point grid bit for bit). */
typedef struct genpk_multi genpk_multi;
genpk_multi *genpk_multi_create(int dims, int ngpus, const int *devices, unsigned flags);
void genpk_multi_destroy(genpk_multi *m);
int genpk_multi_ngpus(const genpk_multi *m);
genpk_ctx *genpk_multi_rank_ctx(genpk_multi *m, int rank);              /* the slab context of one GPU (diagnostics) */
int genpk_multi_set_option(genpk_multi *m, int option, int64_t value);  /* genpk_set_option on every slab */
int genpk_multi_grid_zero(genpk_multi *m);                              /* memset(field, 0, ...), gen-pk.cpp:208 */
/* fieldize() (additive across calls until the next genpk_multi_grid_zero): host arrays, any particle order. */
int genpk_multi_deposit(genpk_multi *m, const float *positions, const float *masses, int64_t n, double mass, double boxsize);
/* fftw_execute + powerspectrum (gen-pk.cpp:233-234); power/count/keffs on the host. */
int genpk_multi_fft_power(genpk_multi *m, int nrbins, double *power, int *count, double *keffs, double total_mass,
                          double total_mass2);
/* zero + deposit + fft_power in one call. */
int genpk_multi_pk_from_particles(genpk_multi *m, const float *positions, const float *masses, int64_t n, double mass,
                                  double boxsize, double total_mass, int nrbins, double *power, int *count, double *keffs);

/* Normalisation of powerspectrum.c:102-108 applied to reduced raw sums (host). */
int genpk_power_finalize(const double *sums_host, int nrbins, double total_mass, double total_mass2,
                         double *power, int *count, double *keffs);

/* Beyond the reference (its own to-do list, gen-pk.cpp:27-31: "rebinning for min modes/bin"): merge neighbouring
 * bins of a finished spectrum, from low k upwards, until every output bin holds at least min_modes modes (a tail with
 * fewer joins the last bin).  Merged values are the mode-weighted means, i.e. exactly what powerspectrum() would have
 * produced with the wider bin: P = sum(count_i P_i) / sum(count_i), k_eff likewise.  In place; empty input bins are
 * skipped.  Returns the number of output bins (<= nrbins), written to the front of the three arrays; the rest is
 * zeroed.  Host only. */
int genpk_rebin_min_modes(int nrbins, double *power, int *count, double *keffs, int64_t min_modes);

/* Host-built bin edges the binning kernel consumes (plan-time table, no GPU
 * needed): thresh_out[b], b = 0..nrbins, is the smallest k2 = ki^2+kj^2+kz^2 >= 1
 * whose bin floor(binsperunit*log|k|) (powerspectrum.c:38,66) is >= b, and
 * thresh_out[nrbins] = 3*(dims/2)^2 + 1.  flags: GENPK_FLAG_BINRULE_SOURCE or 0. */
int genpk_bin_thresholds(int dims, int nrbins, unsigned flags, uint32_t *thresh_out);

/* ================= synthetic particle sets (BASELINE.json configs) ================== */
#define GENPK_SYNTH_UNIFORM_RANDOM 0   /* x = box * hash24(seed, 3p+a) / 2^24, random order       */
#define GENPK_SYNTH_LATTICE        1   /* q = (i+1/2) box/n, z fastest                              */
#define GENPK_SYNTH_CLUSTERED      2   /* lattice + sum of 32 plane-wave displacements, rms ~2 cells*/
/* Fills pos_dev[count][3] with particles first..first+count-1 of an n_side^3 set. */
int genpk_synth_particles(int kind, uint64_t seed, int64_t n_side, int64_t first, int64_t count,
                          double boxsize, double grid_dims, float *pos_dev, void *cuda_stream);

#ifdef __cplusplus
}
#endif
#endif /* GENPK_CUDA_H */
