// gen-pk -- the GenPK command line on top of libgenpk_cuda.so.
//
// Host side of the reference kept as it is for the user: the flags of gen-pk.cpp:95-117
// (-i -j -o -c -s -h), format sniffing (bigfile, else Gadget; gen-pk.cpp:138-167), the grid
// rule (:169-173), the per-particle-type loop (:202-239), the two cross-spectrum modes
// (:240-356), the stdout lines and the three-column output files of utils.cpp:7-21.  The
// numerics between the memset at :208 and the return of powerspectrum() at :234 run on the
// GPU through the C ABI (include/genpk_cuda.h); nothing is computed on the host.
//
// Extensions (defaults keep the reference's behaviour):
//   -g DIMS            FFT grid side instead of the rule of gen-pk.cpp:169-172
//   --fixed            deterministic int64 fixed-point accumulation
//   --gpus N           x-slab decomposition over N GPUs of this box (one process, genpk_multi_*)
//   --synthetic K:N[:SEED]  no snapshot: N^3 device-generated particles, K = uniform|lattice|clustered
//   --json FILE        per-type stage timings
//   --info             print the header lines and the grid side, then exit (no GPU needed)
//   --dump TYPE FILE   write the float32 positions (FILE) and masses (FILE.mass) of one type
//                      exactly as they would be handed to the deposit, then exit (no GPU needed)
#include <cuda_runtime.h>
#include <getopt.h>
#include <limits.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <future>
#include <memory>
#include <string>
#include <vector>

#include "../../include/genpk_cuda.h"
#include "snapshot.hpp"

using namespace genpk_host;

static const int64_t FIELD_DIMS = 3072;     // gen-pk.cpp:63
enum { BARYON_TYPE = 0, DM_TYPE = 1, NEUTRINO_TYPE = 2, STARS_TYPE = 4 };

static int nexttwo(int n)                   // utils.cpp:35-42
{
    n--;
    for (unsigned i = 1; i < sizeof(int) * CHAR_BIT; i <<= 1)
        n |= n >> i;
    return ++n;
}

static std::string type_str(int type)       // utils.cpp:57-72
{
    switch (type) {
    case BARYON_TYPE: return "by";
    case DM_TYPE: return "DM";
    case NEUTRINO_TYPE: return "nu";
    case STARS_TYPE: return "st";
    default: return "xx";
    }
}

// Post-processing beyond the reference (its own to-do list, gen-pk.cpp:27-31); both off by default.
static int64_t g_min_modes = 0;      // --min-modes N: neighbouring bins merged until each holds >= N modes
static int g_fold = 1;               // --fold F: the box folded F times onto itself (small-scale power): k_eff scales by F

static int print_pk(const std::string &filename, int nrbins, const double *keffs, const double *power, const int *count)
{
    FILE *fd = fopen(filename.c_str(), "w");     // utils.cpp:7-21
    if (!fd) {
        fprintf(stderr, "Error opening file: %s\n", filename.c_str());
        return 0;
    }
    std::vector<double> k(keffs, keffs + nrbins), p(power, power + nrbins);
    std::vector<int> c(count, count + nrbins);
    int nout = nrbins;
    if (g_min_modes > 1)
        nout = genpk_rebin_min_modes(nrbins, p.data(), c.data(), k.data(), g_min_modes);
    for (int i = 0; i < nout; i++)
        if (c[i])
            fprintf(fd, "%e\t%e\t%d\n", k[i] * g_fold, p[i], c[i]);
    fclose(fd);
    return nrbins;
}

static void help()
{
    fprintf(stderr,
            "Usage: ./gen-pk -i filenames -j other_filenames -o outdir -c (optional) cross-corr type\n"
            "Outputs one file per particle type, with the name PK-$TYPE-$INPUT\n"
            "Each output file has three columns, for each bin, k_eff, P(k) and N_modes\n"
            "if -j is specified it cross-correlates files specified under -i with files specifed under\n"
            "-j (one output per particle type).\n"
            "If -c is specified the code computes the cross-correlation of that particle type with \n"
            "the CDM (type 1) (within the file specified by -i)\n"
            "-s (1,0) Determines whether stars are included in the baryon type.\n"
            "B200 extensions: -g dims | --fixed | --gpus N | --synthetic kind:n[:seed] | --json file | --info | --dump type file\n"
            "  --min-modes N  merge neighbouring bins until each holds at least N modes\n"
            "  --fold F       fold the box F times onto itself before the deposit (power at F times smaller scales;\n"
            "                 k_eff is printed in fundamental modes of the full box)\n"
            "(HDF5 snapshots are not supported in this build: the image has no HDF5)\n");
}

// One opened snapshot of either format.
struct Source {
    std::unique_ptr<GadgetSnapshot> gadget;
    std::unique_ptr<BigfileSnapshot> big;
    int64_t npart[N_TYPE] = {0, 0, 0, 0, 0, 0};
    double mass[N_TYPE] = {0, 0, 0, 0, 0, 0};
    double box = 0, redshift = 0, omega0 = 0, atime = 0, h100 = 0;
    bool is_big() const { return (bool)big; }
};

static bool open_source(const std::string &path, Source *s)
{
    std::unique_ptr<BigfileSnapshot> b(new BigfileSnapshot(path));
    if (b->ok()) {                                          // gen-pk.cpp:139-148
        if (!b->attr_i8("TotNumPart", s->npart, N_TYPE) || !b->attr_f8("MassTable", s->mass, N_TYPE) ||
            !b->attr_f8("Time", &s->atime, 1) || !b->attr_f8("HubbleParam", &s->h100, 1) ||
            !b->attr_f8("Omega0", &s->omega0, 1) || !b->attr_f8("BoxSize", &s->box, 1)) {
            fprintf(stderr, "Failed to read attr: %s\n", b->error().c_str());
            fprintf(stderr, "Could not load header\n");
            return false;
        }
        s->redshift = 1 / s->atime - 1;
        printf("NumPart=[%ld,%ld,%ld,%ld,%ld,%ld], ", (long)s->npart[0], (long)s->npart[1], (long)s->npart[2],
               (long)s->npart[3], (long)s->npart[4], (long)s->npart[5]);
        printf("Masses=[%g %g %g %g %g %g], ", s->mass[0], s->mass[1], s->mass[2], s->mass[3], s->mass[4], s->mass[5]);
        printf("Redshift=%g, Ω_M=%g\n", s->redshift, s->omega0);
        printf("Expansion factor = %f\n", s->atime);
        printf("Hubble = %g Box=%g \n", s->h100, s->box);
        s->big = std::move(b);
        return true;
    }
    std::unique_ptr<GadgetSnapshot> g(new GadgetSnapshot(path));
    if (!g->ok()) {
        fprintf(stderr, "Could not open %s: %s\n", path.c_str(), g->error().c_str());
        return false;
    }
    for (int t = 0; t < N_TYPE; t++) {                      // gen-pk.cpp:157-160
        s->npart[t] = g->npart(t);
        s->mass[t] = g->header().mass[t];
    }
    s->box = g->header().BoxSize;
    s->redshift = g->header().redshift;
    s->omega0 = g->header().Omega0;
    printf("Boxsize=%g, ", s->box);
    printf("NPart=(%g,%g,%g,%g,%g,%g)**3\n", cbrt(s->npart[0]), cbrt(s->npart[1]), cbrt(s->npart[2]), cbrt(s->npart[3]),
           cbrt(s->npart[4]), cbrt(s->npart[5]));
    printf("Masses=[%g %g %g ]\n", s->mass[0], s->mass[1], s->mass[2]);
    printf("redshift=%g, Ω_M=%g\n", s->redshift, s->omega0);
    s->gadget = std::move(g);
    return true;
}

// Consumer of particle chunks: the GPU deposit, or a file dump.
struct Sink {
    genpk_ctx *ctx = nullptr;
    genpk_multi *multi = nullptr;      // --gpus N: the chunk is split over the GPUs and routed on the device
    int which = 0;
    FILE *dump_pos = nullptr, *dump_mass = nullptr;
    int put(const float *pos, const float *masses, int64_t n, double mass, double box)
    {
        if (dump_pos) {
            fwrite(pos, sizeof(float), 3 * (size_t)n, dump_pos);
            if (masses && dump_mass)
                fwrite(masses, sizeof(float), (size_t)n, dump_mass);
            return 0;
        }
        if (multi ? genpk_multi_deposit(multi, pos, masses, n, mass, box) : genpk_deposit(ctx, which, pos, masses, n, mass, box, 0)) {
            fprintf(stderr, "deposit failed: %s\n", genpk_last_error());
            return 1;
        }
        return 0;
    }
    // double-precision positions go up as they are stored and are narrowed on the GPU
    // (the dump mode keeps the host narrowing: it shows what the reference's reader hands over)
    bool takes_f64() const { return dump_pos == nullptr && multi == nullptr; }
    int put64(const double *pos, const float *masses, int64_t n, double mass, double box)
    {
        if (genpk_deposit_f64(ctx, which, pos, masses, n, mass, box, 0)) {
            fprintf(stderr, "deposit failed: %s\n", genpk_last_error());
            return 1;
        }
        return 0;
    }
};

// Page-locked host memory for the chunks on their way to the GPU (the copies run at PCIe speed and overlap the
// reads); plain memory if the driver refuses.
struct PinnedBuf {
    void *p = nullptr;
    bool pinned = false;
    size_t bytes = 0;
    bool reserve(size_t n)
    {
        if (n <= bytes)
            return true;
        release();
        if (cudaHostAlloc(&p, n, cudaHostAllocPortable) == cudaSuccess) {
            pinned = true;
        } else {
            cudaGetLastError();
            p = malloc(n);
            pinned = false;
        }
        bytes = p ? n : 0;
        return p != nullptr;
    }
    void release()
    {
        if (p && pinned) cudaFreeHost(p);
        if (p && !pinned) free(p);
        p = nullptr;
        bytes = 0;
    }
    ~PinnedBuf() { release(); }
};

// One chunk in flight: positions (float32, or float64 as stored) and masses.
struct ChunkBuf {
    PinnedBuf pos, mass;
    int64_t n = 0;
    bool ok = true;
};

// read_fieldize() (read_fieldize.cpp:18-97) and read_fieldize_bigfile()
// (read_fieldize_bigfile.cpp:64-125): stream one particle type into the sink and
// accumulate total_mass the way the reference does, quirks included.  The reference reads a chunk of 2^26
// particles, deposits it, reads the next; here a reader thread fills one page-locked buffer (the pieces of a chunk
// that live in different files fetched in parallel, parallel_read.cpp) while the previous one is on its way to the
// GPU.  Chunks of 2^24 particles; the mass sums are taken in the reference's order and grouping.
static int read_deposit(const Source &src, int type, double box, Sink &sink, double *total_mass)
{
    const int64_t npart_total = src.npart[type];
    if (npart_total == 0)
        return 1;
    const double mass = src.mass[type];
    const int64_t ref_chunk = std::min<int64_t>(npart_total, (int64_t)1 << 26);   // read_fieldize.cpp:45
    int64_t chunk = (int64_t)1 << 24;
    if (const char *env = getenv("GENPK_READ_CHUNK"))                              // particles per chunk (tests: many small chunks)
        if (atoll(env) > 0)
            chunk = atoll(env);
    while (ref_chunk % chunk != 0 && chunk > 1)                                    // the reference's chunk boundaries are chunk boundaries
        chunk = ref_chunk % chunk;                                                 // (Euclid: ends at a divisor of ref_chunk)
    chunk = std::min(chunk, npart_total);
    const bool with_mass = mass == 0;
    ChunkBuf bufs[2];

    BigBlockInfo bp, bm;
    bool f64 = false;
    int skip_type = 0;
    if (src.is_big()) {
        char name[32];
        snprintf(name, sizeof(name), "%d/Position", type);
        if (!src.big->open_block(name, &bp) || bp.nmemb != 3 || bp.rows < npart_total) {
            fprintf(stderr, "Failed to open block at %s:%s\n", name, src.big->error().c_str());
            return 1;
        }
        if (with_mass) {
            snprintf(name, sizeof(name), "%d/Mass", type);
            if (!src.big->open_block(name, &bm) || bm.nmemb != 1 || bm.rows < npart_total) {
                fprintf(stderr, "Failed to open block at %s:%s\n", name, src.big->error().c_str());
                return 1;
            }
        }
        f64 = bp.dtype == "<f8" && sink.takes_f64();                               // stored doubles: narrowed on the GPU
    } else {
        const GadgetSnapshot &snap = *src.gadget;
        skip_type = ((1 << N_TYPE) - 1) - (1 << type);                            // read_fieldize.cpp:27,37
        const int64_t parts_in_pos = snap.block_parts("POS ");
        if (parts_in_pos == 0 || snap.block_bytes("POS ") / parts_in_pos != 3 * (int64_t)sizeof(float)) {
            fprintf(stderr, "The pos array uses %ld bytes per particle, instead of %lu.\n"
                            " Double-precision snapshots are not supported by this build.\n",
                    parts_in_pos ? (long)(snap.block_bytes("POS ") / parts_in_pos) : 0L, 3 * sizeof(float));
            return 1;
        }
    }
    for (ChunkBuf &b : bufs) {
        if (!b.pos.reserve((size_t)chunk * 3 * (f64 ? sizeof(double) : sizeof(float))) ||
            (with_mass && !b.mass.reserve((size_t)chunk * sizeof(float)))) {
            fprintf(stderr, "out of host memory for the read buffers\n");
            return 1;
        }
    }
    // the reader: particles [first, first + n) of this type into b
    auto fill = [&](ChunkBuf *b, int64_t first, int64_t n) {
        b->n = n;
        b->ok = true;
        if (src.is_big()) {
            if (!(f64 ? src.big->read_f64_raw(bp, first, n, (double *)b->pos.p) : src.big->read_f32(bp, first, n, (float *)b->pos.p)) ||
                (with_mass && !src.big->read_f32(bm, first, n, (float *)b->mass.p))) {
                fprintf(stderr, "Failed to read from block: %s\n", src.big->error().c_str());
                b->ok = false;
            }
            return;
        }
        if (src.gadget->get_block("POS ", b->pos.p, n, first, skip_type) != n) {
            fprintf(stderr, "Error reading particle data for type %d\n", type);
            b->ok = false;
        } else if (with_mass && src.gadget->get_block("MASS", b->mass.p, n, first, skip_type) != n) {
            fprintf(stderr, "Error reading mass data for type %d\n", type);
            b->ok = false;
        }
    };
    const int64_t nchunks = (npart_total + chunk - 1) / chunk;
    std::future<void> pending = std::async(std::launch::async, fill, &bufs[0], (int64_t)0, std::min(chunk, npart_total));
    double group_mass = 0;                                                         // total_mass_this_file of the reference's chunk
    int64_t in_group = 0;
    for (int64_t i = 0; i < nchunks; i++) {
        pending.get();
        ChunkBuf &b = bufs[i & 1];
        if (i + 1 < nchunks) {
            const int64_t first = (i + 1) * chunk;
            pending = std::async(std::launch::async, fill, &bufs[(i + 1) & 1], first, std::min(chunk, npart_total - first));
        }
        int rc = b.ok ? 0 : 1;
        if (!rc) {
            const float *masses = with_mass ? (const float *)b.mass.p : nullptr;
            if (with_mass)
                for (int64_t k = 0; k < b.n; k++)
                    group_mass += masses[k];
            // (the deposit returns once the chunk's copies have left this buffer)
            rc = f64 ? sink.put64((const double *)b.pos.p, masses, b.n, mass, box) : sink.put((const float *)b.pos.p, masses, b.n, mass, box);
        }
        if (rc) {
            if (pending.valid())
                pending.get();
            return 1;
        }
        in_group += b.n;
        // the sums of a 2^26 chunk of the reference (or of the whole type on the bigfile path) are closed here
        const bool group_ends = src.is_big() ? i + 1 == nchunks : (in_group == ref_chunk || i + 1 == nchunks);
        if (group_ends) {
            if (src.is_big()) {
                *total_mass += group_mass;                                         // read_fieldize_bigfile.cpp:110-112
                *total_mass += mass * npart_total;                                 // :120 (no "+1" on this path)
            } else {
                if (with_mass)
                    *total_mass += group_mass;                                     // read_fieldize.cpp:72-75
                *total_mass += mass * in_group;                                    // :77
            }
            group_mass = 0;
            in_group = 0;
        }
    }
    if (!src.is_big())
        *total_mass += 1;                                                          // read_fieldize.cpp:94
    return 0;
}

struct Timing {
    std::string label;
    double wall_ms = 0;
    float zero = 0, deposit = 0, fft = 0, power = 0;
};

// Stage times of everything since the last genpk_stage_reset (all the chunks of a type, not the last one).  With
// --gpus N these are the times of the first GPU's slab.
static void stage_times(genpk_ctx *ctx, Timing *t)
{
    int64_t records = 0;
    genpk_stage_total_ms(ctx, GENPK_STAGE_ZERO, &t->zero, &records);
    genpk_stage_total_ms(ctx, GENPK_STAGE_DEPOSIT, &t->deposit, &records);
    genpk_stage_total_ms(ctx, GENPK_STAGE_FFT, &t->fft, &records);
    genpk_stage_total_ms(ctx, GENPK_STAGE_POWER, &t->power, &records);
}

static double now_ms()
{
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

int main(int argc, char *argv[])
{
    std::string infiles, jinfiles, outdir, synthetic, json_path, dump_path;
    int crosstype = -1, dump_type = -1;
    bool stars_are_baryons = false, fixed = false, info_only = false;
    int64_t grid_override = 0;
    int ngpus = 1;
    static const option long_opts[] = {{"fixed", no_argument, nullptr, 1000},    {"synthetic", required_argument, nullptr, 1001},
                                       {"json", required_argument, nullptr, 1002}, {"info", no_argument, nullptr, 1003},
                                       {"dump", required_argument, nullptr, 1004}, {"gpus", required_argument, nullptr, 1005},
                                       {"min-modes", required_argument, nullptr, 1006}, {"fold", required_argument, nullptr, 1007},
                                       {nullptr, 0, nullptr, 0}};
    int c;
    while ((c = getopt_long(argc, argv, "i:j:o:c:s:g:h", long_opts, nullptr)) != -1) {
        switch (c) {
        case 'o': outdir = optarg; break;
        case 'i': infiles = optarg; break;
        case 'j': jinfiles = optarg; break;
        case 'c': crosstype = atoi(optarg); break;
        case 's': stars_are_baryons = atoi(optarg) != 0; break;
        case 'g': grid_override = atoll(optarg); break;
        case 1000: fixed = true; break;
        case 1001: synthetic = optarg; break;
        case 1002: json_path = optarg; break;
        case 1003: info_only = true; break;
        case 1004:
            dump_type = atoi(optarg);
            if (optind < argc)
                dump_path = argv[optind++];
            break;
        case 1005: ngpus = atoi(optarg); break;
        case 1006: g_min_modes = atoll(optarg); break;
        case 1007:
            g_fold = atoi(optarg);
            if (g_fold < 1) {
                fprintf(stderr, "--fold wants a positive integer\n");
                return 1;
            }
            break;
        case 'h':
        default:
            help();
            return 0;
        }
    }
    const bool need_out = !info_only && dump_type < 0;
    if ((infiles.empty() && synthetic.empty()) || (need_out && outdir.empty())) {
        help();
        return 0;
    }

    // ---- input -----------------------------------------------------------------------
    Source src, src2;
    int64_t syn_side = 0;
    int syn_kind = GENPK_SYNTH_CLUSTERED;
    uint64_t syn_seed = 42;
    if (!synthetic.empty()) {
        char kind[32] = "";
        long long n = 0;
        unsigned long long seed = 42;
        if (sscanf(synthetic.c_str(), "%31[^:]:%lld:%llu", kind, &n, &seed) < 2 || n < 1) {
            fprintf(stderr, "--synthetic wants kind:n[:seed], kind = uniform|lattice|clustered\n");
            return 1;
        }
        syn_kind = !strcmp(kind, "uniform") ? GENPK_SYNTH_UNIFORM_RANDOM : !strcmp(kind, "lattice") ? GENPK_SYNTH_LATTICE
                                                                                                    : GENPK_SYNTH_CLUSTERED;
        syn_side = n;
        syn_seed = seed;
        src.npart[DM_TYPE] = n * n * n;
        src.mass[DM_TYPE] = 1.0;
        src.box = 1000.0;
        infiles = "synthetic-" + std::string(kind) + "-" + std::to_string(n);
        printf("Boxsize=%g, synthetic %s %lld^3 particles, seed %llu\n", src.box, kind, n, seed);
    } else if (!open_source(infiles, &src)) {
        return 1;
    }
    if (!jinfiles.empty() && !open_source(jinfiles, &src2))
        return 1;
    // --fold F: the deposit wraps every position periodically (fieldize.cpp:70-75), so a box F times smaller IS the fold
    const double box = src.box / g_fold;

    // ---- grid side: gen-pk.cpp:169-173 ---------------------------------------------------
    int64_t field_dims = 0;
    for (int type = 0; type < N_TYPE; type++) {
        const int64_t tmp = 2 * (int64_t)nexttwo((int)cbrt((double)src.npart[type]));
        field_dims = std::max(field_dims, std::min(tmp, FIELD_DIMS));
    }
    if (grid_override > 0)
        field_dims = grid_override;
    const int nrbins = (int)field_dims;
    printf("FFT grid dimension: %lu\n", (unsigned long)field_dims);
    if (info_only)
        return 0;
    if (dump_type >= 0) {
        if (dump_type >= N_TYPE || dump_path.empty() || src.npart[dump_type] == 0 || synthetic.size()) {
            fprintf(stderr, "--dump TYPE FILE: type not present\n");
            return 1;
        }
        Sink sink;
        sink.dump_pos = fopen(dump_path.c_str(), "wb");
        sink.dump_mass = src.mass[dump_type] == 0 ? fopen((dump_path + ".mass").c_str(), "wb") : nullptr;
        if (!sink.dump_pos)
            return 1;
        double tm = 0;
        const int rc = read_deposit(src, dump_type, box, sink, &tm);
        fclose(sink.dump_pos);
        if (sink.dump_mass)
            fclose(sink.dump_mass);
        printf("total_mass in type %d = %.17g\n", dump_type, tm);
        return rc;
    }

    // ---- GPU context: the field of gen-pk.cpp:176-193 lives in HBM ---------------------------
    const bool two_fields = crosstype >= 0 || !jinfiles.empty();
    unsigned flags = (fixed ? GENPK_FLAG_FIXED_POINT : 0) | (two_fields ? GENPK_FLAG_TWO_FIELDS : 0);
    genpk_multi *multi = nullptr;
    if (ngpus > 1) {
        // --gpus N: x-slab decomposition over N GPUs of this box, one process (genpk_multi_*)
        if (two_fields || !synthetic.empty()) {
            fprintf(stderr, "--gpus applies to the per-type spectra of a snapshot (no -c / -j / --synthetic)\n");
            return 1;
        }
        multi = genpk_multi_create((int)field_dims, ngpus, nullptr, flags);
        if (!multi) {
            fprintf(stderr, "Error setting up %d GPUs: %s\n", ngpus, genpk_last_error());
            return 1;
        }
    }
    genpk_ctx *ctx = multi ? genpk_multi_rank_ctx(multi, 0) : genpk_create((int)field_dims, -1, flags);
    if (!ctx) {
        fprintf(stderr, "Error allocating memory for grid: %s\n", genpk_last_error());
        return 1;
    }
    if (fixed) {
        // --fixed: the scale of the integer sums follows the mass unit of each particle type
        if (multi ? genpk_multi_set_option(multi, GENPK_OPT_SCALE_BITS, -1) : genpk_set_option(ctx, GENPK_OPT_SCALE_BITS, -1)) {
            fprintf(stderr, "%s\n", genpk_last_error());
            return 1;
        }
    }
    std::vector<double> power(nrbins), keffs(nrbins);
    std::vector<int> count(nrbins);
    const size_t last = infiles.find_last_of("/\\");
    const std::string base = infiles.substr(last == std::string::npos ? 0 : last + 1);
    std::vector<Timing> timings;
    int status = 0;

    // both transforms and the two-field binning in one call (fftw_execute x2 + powerspectrum, gen-pk.cpp:295-297,
    // 345-348): on the fused grid sides neither x-transformed spectrum is ever written
    auto run_power = [&](int a, int b, double tm1, double tm2, const std::string &filename, Timing *t) {
        if (genpk_fft_power_cross(ctx, a, ctx, b, nrbins, power.data(), count.data(), keffs.data(), tm1, tm2)) {
            fprintf(stderr, "powerspectrum failed: %s\n", genpk_last_error());
            return 1;
        }
        if (genpk_synchronize(ctx)) {
            fprintf(stderr, "%s\n", genpk_last_error());
            return 1;
        }
        stage_times(ctx, t);
        print_pk(filename, nrbins, keffs.data(), power.data(), count.data());
        return 0;
    };

    if (crosstype < 0 && jinfiles.empty()) {
        // ---- one spectrum per particle type: gen-pk.cpp:202-239 ------------------------------
        for (int type = 0; type < N_TYPE; type++) {
            if (src.npart[type] == 0)
                continue;
            Timing t;
            t.label = type_str(type);
            genpk_stage_reset(ctx);
            const double t0 = now_ms();
            if (multi)
                genpk_multi_grid_zero(multi);
            else
                genpk_grid_zero(ctx, 0);                                           // :208
            double total_mass = 0;
            if (syn_side) {
                float *dpos = nullptr;
                const int64_t n = src.npart[type];
                if (cudaMalloc(&dpos, (size_t)n * 12) != cudaSuccess) {
                    fprintf(stderr, "Error allocating %ld MB of device memory for the synthetic particles\n",
                            (long)(n * 12 >> 20));
                    status = 1;
                    break;
                }
                if (genpk_synth_particles(syn_kind, syn_seed, syn_side, 0, n, src.box, (double)field_dims, dpos, nullptr) ||
                    genpk_deposit(ctx, 0, dpos, nullptr, n, src.mass[type], box, 1) || genpk_synchronize(ctx)) {
                    fprintf(stderr, "synthetic deposit failed: %s\n", genpk_last_error());
                    status = 1;
                }
                cudaFree(dpos);
                if (status)
                    break;
                total_mass = src.mass[type] * (double)n;
            } else {
                Sink sink;
                sink.ctx = ctx;
                sink.multi = multi;
                read_deposit(src, type, box, sink, &total_mass);                   // :221,227
                if (type == BARYON_TYPE && stars_are_baryons)
                    read_deposit(src, STARS_TYPE, box, sink, &total_mass);         // :228-230
            }
            printf("total_mass in type %d = %g\n", type, total_mass);             // :232
            // :233-238: fftw_execute + powerspectrum as one call (the last FFT pass and the binning share a
            // kernel for grid sides 256/512/1024/2048; the library's 3-D plan + binning pass otherwise)
            if (multi ? genpk_multi_fft_power(multi, nrbins, power.data(), count.data(), keffs.data(), total_mass, total_mass)
                      : (genpk_fft_power(ctx, 0, nrbins, power.data(), count.data(), keffs.data(), total_mass, total_mass) ||
                         genpk_synchronize(ctx))) {
                fprintf(stderr, "FFT / powerspectrum failed: %s\n", genpk_last_error());
                status = 1;
                break;
            }
            stage_times(ctx, &t);
            print_pk(outdir + "/PK-" + type_str(type) + "-" + base, nrbins, keffs.data(), power.data(), count.data());
            t.wall_ms = now_ms() - t0;
            timings.push_back(t);
        }
    } else if (!jinfiles.empty()) {
        // ---- same type in two snapshots: gen-pk.cpp:240-303 (PX- files) ----------------------
        for (int type = 0; type < N_TYPE; type++) {
            if (src.npart[type] == 0)
                continue;
            Timing t;
            t.label = "x" + type_str(type);
            genpk_stage_reset(ctx);
            const double t0 = now_ms();
            genpk_grid_zero(ctx, 0);
            genpk_grid_zero(ctx, 1);
            double tm1 = 0, tm2 = 0;
            Sink s1, s2;
            s1.ctx = s2.ctx = ctx;
            s2.which = 1;
            if (read_deposit(src, type, box, s1, &tm1) || read_deposit(src2, type, box, s2, &tm2))
                continue;
            if (run_power(0, 1, tm1, tm2, outdir + "/PX-" + type_str(type) + "-" + base, &t))
                continue;
            t.wall_ms = now_ms() - t0;
            timings.push_back(t);
        }
    } else {
        // ---- DM x another type inside one snapshot: gen-pk.cpp:304-356 -------------------------
        if (crosstype >= N_TYPE || src.npart[DM_TYPE] == 0 || src.npart[crosstype] == 0) {
            fprintf(stderr, "Can't cross-correlate types not present in snapshot\n");
            genpk_destroy(ctx);
            return 1;
        }
        Timing t;
        t.label = "DMx" + type_str(crosstype);
        genpk_stage_reset(ctx);
        const double t0 = now_ms();
        genpk_grid_zero(ctx, 0);
        genpk_grid_zero(ctx, 1);
        double tm1 = 0, tm2 = 0;
        Sink s1, s2;
        s1.ctx = s2.ctx = ctx;
        s2.which = 1;
        read_deposit(src, DM_TYPE, box, s1, &tm1);
        read_deposit(src, crosstype, box, s2, &tm2);
        if (!run_power(0, 1, tm1, tm2, outdir + "/PK-DMx" + type_str(crosstype) + "-" + base, &t)) {
            t.wall_ms = now_ms() - t0;
            timings.push_back(t);
        }
    }

    if (!json_path.empty()) {
        FILE *fj = fopen(json_path.c_str(), "w");
        if (fj) {
            fprintf(fj, "{\"grid\": %ld, \"nrbins\": %d, \"accumulation\": \"%s\", \"spectra\": [", (long)field_dims, nrbins,
                    fixed ? "int64 fixed-point" : "fp64");
            for (size_t i = 0; i < timings.size(); i++)
                fprintf(fj, "%s{\"type\": \"%s\", \"wall_ms\": %.3f, \"zero_ms\": %.3f, \"deposit_ms\": %.3f, \"fft_ms\": %.3f, "
                            "\"binning_ms\": %.3f}",
                        i ? ", " : "", timings[i].label.c_str(), timings[i].wall_ms, timings[i].zero, timings[i].deposit,
                        timings[i].fft, timings[i].power);
            fprintf(fj, "]}\n");
            fclose(fj);
        }
    }
    if (multi)
        genpk_multi_destroy(multi);
    else
        genpk_destroy(ctx);
    return status;
}
