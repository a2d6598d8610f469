// Multi-GPU P(k) behind the C ABI: one process, one slab context per GPU, no collective library.
//
// The per-particle-type loop of gen-pk.cpp:202-239 on N GPUs of one box: the grid is
// slab-decomposed in x, particles arrive in ANY order from the host and are routed on the device
// (bucket by owner slab, runs moved with peer copies over NVLink), each GPU deposits what it owns,
// the ghost plane is pulled from the neighbour's memory, the FFT transpose is stored straight into the
// owners' blocks by the y pass (or moved by peer copies on grid sides the column kernels do not
// cover), the last FFT pass is fused with the binning, and the 3*nrbins partial sums are added on
// the host.  Cross-GPU ordering is by CUDA events (cudaStreamWaitEvent works across the devices of one
// process); one host thread per GPU drives the uploads so that the PCIe links run side by side.
//
// N contexts may share a device (devices = {0, 0, ...}): that is how the path is tested on one GPU.
#include <math.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <thread>
#include <vector>

#include "common.cuh"

namespace genpk {

#define MULTI_OK(expr)                                                                          \
    do {                                                                                        \
        cudaError_t e__ = (expr);                                                               \
        if (e__ != cudaSuccess) {                                                               \
            genpk::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e__)); \
            return 1;                                                                           \
        }                                                                                       \
    } while (0)

struct RankBuf {
    float *in_pos = nullptr, *in_mass = nullptr;        // this rank's share of the call's particles, as uploaded
    float *sorted_pos = nullptr, *sorted_mass = nullptr;   // ... grouped by owner slab
    int64_t *d_counts = nullptr;
    int64_t cap = 0;
    bool has_mass = false;
    float *recv_pos = nullptr, *recv_mass = nullptr;    // particles this rank owns, gathered from every rank
    int64_t recv_cap = 0;
    bool recv_has_mass = false;
    void *send = nullptr, *spec = nullptr;              // pack + peer-copy transpose (grid sides without the scatter pass)
    double *d_sums = nullptr;
    double *h_sums = nullptr;                           // pinned
    int sums_cap = 0;
};

}  // namespace genpk

struct genpk_multi {
    int n = 0, dims = 0;
    unsigned flags = 0;
    std::vector<int> dev;
    std::vector<genpk_ctx *> ctx;
    std::vector<cudaStream_t> stream;
    std::vector<genpk::RankBuf> buf;
    std::vector<cudaEvent_t> ev_a, ev_b, ev_c;          // per rank: deposit done / scatter (or pack) done / spectrum consumed
    bool scatter = false;
    bool scale_latched = false;                         // fixed point: one scale for all slabs, fixed by the first deposit after a zero
    std::vector<std::vector<int64_t>> counts;           // [source][dest] of the current call
    std::string thread_error;
};

namespace genpk {

static int use_device(const genpk_multi *m, int r) { return cudaSetDevice(m->dev[r]) == cudaSuccess ? 0 : 1; }

template <class F> static int for_each_rank_parallel(genpk_multi *m, F fn)
{
    std::vector<int> rc(m->n, 0);
    std::vector<std::string> err(m->n);
    std::vector<std::thread> th;
    for (int r = 0; r < m->n; r++)
        th.emplace_back([&, r]() {
            if (use_device(m, r)) {
                rc[r] = 1;
                err[r] = "cudaSetDevice failed";
                return;
            }
            rc[r] = fn(r);
            if (rc[r])
                err[r] = genpk_last_error();            // (the message is thread-local)
        });
    for (auto &t : th)
        t.join();
    for (int r = 0; r < m->n; r++)
        if (rc[r]) {
            set_error("rank %d: %s", r, err[r].c_str());
            return rc[r];
        }
    return 0;
}

static int grow_in(genpk_multi *m, int r, int64_t cap, bool mass)
{
    RankBuf &b = m->buf[r];
    if (cap > b.cap) {
        cudaFree(b.in_pos); cudaFree(b.sorted_pos); cudaFree(b.in_mass); cudaFree(b.sorted_mass);
        b.in_pos = b.sorted_pos = b.in_mass = b.sorted_mass = nullptr;
        b.cap = 0;
        b.has_mass = false;
        MULTI_OK(cudaMalloc(&b.in_pos, (size_t)cap * 12));
        MULTI_OK(cudaMalloc(&b.sorted_pos, (size_t)cap * 12));
        b.cap = cap;
    }
    if (mass && !b.has_mass) {
        MULTI_OK(cudaMalloc(&b.in_mass, (size_t)b.cap * 4));
        MULTI_OK(cudaMalloc(&b.sorted_mass, (size_t)b.cap * 4));
        b.has_mass = true;
    }
    if (!b.d_counts)
        MULTI_OK(cudaMalloc(&b.d_counts, (size_t)m->n * sizeof(int64_t)));
    return 0;
}

static int grow_recv(genpk_multi *m, int r, int64_t cap, bool mass)
{
    RankBuf &b = m->buf[r];
    if (cap > b.recv_cap) {
        cudaFree(b.recv_pos); cudaFree(b.recv_mass);
        b.recv_pos = b.recv_mass = nullptr;
        b.recv_cap = 0;
        b.recv_has_mass = false;
        const int64_t want = cap + cap / 8 + 1024;
        MULTI_OK(cudaMalloc(&b.recv_pos, (size_t)want * 12));
        b.recv_cap = want;
    }
    if (mass && !b.recv_has_mass) {
        MULTI_OK(cudaMalloc(&b.recv_mass, (size_t)b.recv_cap * 4));
        b.recv_has_mass = true;
    }
    return 0;
}

static int grow_sums(genpk_multi *m, int r, int nrbins)
{
    RankBuf &b = m->buf[r];
    if (nrbins > b.sums_cap) {
        cudaFree(b.d_sums);
        if (b.h_sums) cudaFreeHost(b.h_sums);
        b.d_sums = b.h_sums = nullptr;
        b.sums_cap = 0;
        MULTI_OK(cudaMalloc(&b.d_sums, (size_t)(3 * nrbins + 1) * sizeof(double)));
        MULTI_OK(cudaMallocHost(&b.h_sums, (size_t)(3 * nrbins + 1) * sizeof(double)));
        b.sums_cap = nrbins;
    }
    return 0;
}

}  // namespace genpk

using namespace genpk;

extern "C" {

void genpk_multi_destroy(genpk_multi *m)
{
    if (!m)
        return;
    for (int r = 0; r < (int)m->ctx.size(); r++) {
        cudaSetDevice(m->dev[r]);
        cudaDeviceSynchronize();
        RankBuf &b = m->buf[r];
        cudaFree(b.in_pos); cudaFree(b.in_mass); cudaFree(b.sorted_pos); cudaFree(b.sorted_mass); cudaFree(b.d_counts);
        cudaFree(b.recv_pos); cudaFree(b.recv_mass); cudaFree(b.send); cudaFree(b.spec); cudaFree(b.d_sums);
        if (b.h_sums) cudaFreeHost(b.h_sums);
        if (m->ctx[r]) genpk_destroy(m->ctx[r]);
        if (r < (int)m->stream.size() && m->stream[r]) cudaStreamDestroy(m->stream[r]);
        for (auto *v : {&m->ev_a, &m->ev_b, &m->ev_c})
            if (r < (int)v->size() && (*v)[r]) cudaEventDestroy((*v)[r]);
    }
    delete m;
}

genpk_multi *genpk_multi_create(int dims, int ngpus, const int *devices, unsigned flags)
{
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) {
        set_error("genpk_multi_create: no CUDA device: %s", cudaGetErrorString(cudaGetLastError()));
        return nullptr;
    }
    if (ngpus < 1 || ngpus > GENPK_MAX_PEERS || dims < 1 || dims % ngpus != 0) {
        set_error("genpk_multi_create: bad geometry dims=%d ngpus=%d (dims must be divisible by ngpus <= %d)", dims, ngpus,
                  GENPK_MAX_PEERS);
        return nullptr;
    }
    if (flags & GENPK_FLAG_TWO_FIELDS) {
        set_error("genpk_multi_create: one field per multi-GPU context (cross spectra: two contexts)");
        return nullptr;
    }
    genpk_multi *m = new genpk_multi();
    m->n = ngpus;
    m->dims = dims;
    m->flags = flags;
    m->dev.resize(ngpus);
    m->ctx.assign(ngpus, nullptr);
    m->stream.assign(ngpus, nullptr);
    m->buf.resize(ngpus);
    m->ev_a.assign(ngpus, nullptr);
    m->ev_b.assign(ngpus, nullptr);
    m->ev_c.assign(ngpus, nullptr);
    m->counts.assign(ngpus, std::vector<int64_t>(ngpus, 0));
    bool ok = true;
    for (int r = 0; r < ngpus && ok; r++) {
        // no list: spread over the visible GPUs (N = 4 of 8: 0, 2, 4, 6 -- neighbours tend to share a PCIe bridge to the
        // host, and every slab uploads its share of a chunk at once), or share devices when there are fewer than N
        const int stride = (ndev >= ngpus && ndev % ngpus == 0) ? ndev / ngpus : 1;
        m->dev[r] = devices ? devices[r] : (r * stride) % ndev;
        ok = m->dev[r] >= 0 && m->dev[r] < ndev;
        if (!ok) set_error("genpk_multi_create: device %d does not exist", m->dev[r]);
    }
    // peer access between every pair of distinct devices
    for (int r = 0; r < ngpus && ok; r++)
        for (int s = 0; s < ngpus && ok; s++)
            if (m->dev[r] != m->dev[s]) {
                int can = 0;
                cudaDeviceCanAccessPeer(&can, m->dev[r], m->dev[s]);
                if (!can) {
                    set_error("genpk_multi_create: device %d cannot access device %d", m->dev[r], m->dev[s]);
                    ok = false;
                    break;
                }
                cudaSetDevice(m->dev[r]);
                const cudaError_t e = cudaDeviceEnablePeerAccess(m->dev[s], 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
                    set_error("genpk_multi_create: cudaDeviceEnablePeerAccess(%d -> %d): %s", m->dev[r], m->dev[s], cudaGetErrorString(e));
                    ok = false;
                }
                cudaGetLastError();
            }
    for (int r = 0; r < ngpus && ok; r++) {
        m->ctx[r] = ngpus == 1 ? genpk_create(dims, m->dev[r], flags) : genpk_create_slab(dims, m->dev[r], ngpus, r, flags);
        ok = m->ctx[r] != nullptr;
        ok = ok && cudaSetDevice(m->dev[r]) == cudaSuccess && cudaStreamCreateWithFlags(&m->stream[r], cudaStreamNonBlocking) == cudaSuccess;
        ok = ok && genpk_set_stream(m->ctx[r], m->stream[r]) == 0;
        ok = ok && cudaEventCreateWithFlags(&m->ev_a[r], cudaEventDisableTiming) == cudaSuccess;
        ok = ok && cudaEventCreateWithFlags(&m->ev_b[r], cudaEventDisableTiming) == cudaSuccess;
        ok = ok && cudaEventCreateWithFlags(&m->ev_c[r], cudaEventDisableTiming) == cudaSuccess;
        if (!ok && m->ctx[r]) set_error("genpk_multi_create: stream / event creation failed on device %d", m->dev[r]);
    }
    if (ok && ngpus > 1) {
        // ghost pull: the low neighbour's grid; transpose: every rank's transposed block when the y pass can scatter
        m->scatter = true;
        for (int r = 0; r < ngpus; r++)
            m->scatter = m->scatter && genpk_slab_scatter_supported(m->ctx[r]) && genpk_fused_xpass_supported(m->ctx[r], dims);
        std::vector<void *> grid_ptr(ngpus, nullptr), recv_ptr(ngpus, nullptr);
        for (int r = 0; r < ngpus && ok; r++) {                  // (allocations and memsets belong to the owner's device)
            cudaSetDevice(m->dev[r]);
            grid_ptr[r] = genpk_grid_device_ptr(m->ctx[r], 0);
            ok = grid_ptr[r] != nullptr;
            if (ok && m->scatter) {
                recv_ptr[r] = genpk_slab_recv_buffer(m->ctx[r], nullptr);
                ok = recv_ptr[r] != nullptr;
            }
        }
        for (int r = 0; r < ngpus && ok; r++) {
            cudaSetDevice(m->dev[r]);
            ok = genpk_slab_set_grid_peer(m->ctx[r], 0, 0, nullptr, grid_ptr[(r + ngpus - 1) % ngpus]) == 0;
            if (m->scatter)
                for (int s = 0; s < ngpus && ok; s++)
                    ok = genpk_slab_set_peer(m->ctx[r], s, nullptr, recv_ptr[s]) == 0;
        }
        for (int r = 0; r < ngpus && ok; r++) {
            cudaSetDevice(m->dev[r]);
            ok = cudaStreamSynchronize(m->stream[r]) == cudaSuccess;
        }
    }
    if (!ok) {
        const std::string keep = genpk_last_error();
        genpk_multi_destroy(m);
        set_error("%s", keep.c_str());
        return nullptr;
    }
    return m;
}

int genpk_multi_ngpus(const genpk_multi *m) { return m ? m->n : 0; }
genpk_ctx *genpk_multi_rank_ctx(genpk_multi *m, int rank) { return m && rank >= 0 && rank < m->n ? m->ctx[rank] : nullptr; }

int genpk_multi_set_option(genpk_multi *m, int option, int64_t value)
{
    if (!m) { set_error("genpk_multi_set_option: null context"); return 1; }
    for (int r = 0; r < m->n; r++)
        if (int rc = genpk_set_option(m->ctx[r], option, value)) return rc;
    return 0;
}

int genpk_multi_grid_zero(genpk_multi *m)
{
    if (!m) { set_error("genpk_multi_grid_zero: null context"); return 1; }
    for (int r = 0; r < m->n; r++) {
        if (use_device(m, r)) { set_error("cudaSetDevice failed"); return 1; }
        if (int rc = genpk_grid_zero(m->ctx[r], 0)) return rc;
    }
    m->scale_latched = false;
    return 0;
}

// fieldize() over N GPUs: host particles in any order.  Rank q uploads the q-th part of the array over its own
// PCIe link and groups it by owner slab on the device; then every rank gathers its runs from all ranks with
// peer copies and deposits them.  Additive across calls (the chunk loop of read_fieldize.cpp:51-93).
int genpk_multi_deposit(genpk_multi *m, const float *positions, const float *masses, int64_t n, double mass, double boxsize)
{
    if (!m) { set_error("genpk_multi_deposit: null context"); return 1; }
    if (n < 0 || (n > 0 && !positions)) { set_error("genpk_multi_deposit: bad particle array"); return 1; }
    if (n == 0) return 0;
    if (m->n == 1) {
        if (use_device(m, 0)) { set_error("cudaSetDevice failed"); return 1; }
        return genpk_deposit(m->ctx[0], 0, positions, masses, n, mass, boxsize, 0);
    }
    const int P = m->n;
    if (m->ctx[0]->fixed && !m->scale_latched) {
        // every slab must scale its fixed-point sums alike (ghost planes are added as integers): the automatic
        // scale (GENPK_OPT_SCALE_BITS = -1) is taken here, from all the masses of this first deposit
        int bits = m->ctx[0]->scale_bits;
        if (bits < 0) {
            double biggest = fabs(mass);
            if (masses) {
                float h = 0.f;
                for (int64_t i = 0; i < n; i++) {
                    const float v = fabsf(masses[i]);
                    if (v < 3.0e38f && v > h) h = v;
                }
                biggest = h;
            }
            bits = 40;
            if (biggest > 0 && biggest < 1e300) bits = 40 - (int)ceil(log2(biggest));
            bits = bits < 0 ? 0 : (bits > 400 ? 400 : bits);
        }
        for (int r = 0; r < P; r++) {
            m->ctx[r]->grid_scale_bits[0] = bits;
            m->ctx[r]->grid_scale_latched[0] = true;
        }
        m->scale_latched = true;
    }
    // ---- phase 1: upload + route, every rank its own share ----
    int rc = for_each_rank_parallel(m, [&](int q) -> int {
        const int64_t lo = n * q / P, cnt = n * (q + 1) / P - lo;
        std::fill(m->counts[q].begin(), m->counts[q].end(), 0);
        if (cnt == 0) return 0;
        if (int e = grow_in(m, q, cnt, masses != nullptr)) return e;
        RankBuf &b = m->buf[q];
        MULTI_OK(cudaMemcpyAsync(b.in_pos, positions + 3 * lo, (size_t)cnt * 12, cudaMemcpyHostToDevice, m->stream[q]));
        if (masses)
            MULTI_OK(cudaMemcpyAsync(b.in_mass, masses + lo, (size_t)cnt * 4, cudaMemcpyHostToDevice, m->stream[q]));
        if (int e = genpk_route_particles(m->ctx[q], b.in_pos, masses ? b.in_mass : nullptr, cnt, boxsize, b.sorted_pos,
                                          masses ? b.sorted_mass : nullptr, b.d_counts))
            return e;
        MULTI_OK(cudaMemcpyAsync(m->counts[q].data(), b.d_counts, (size_t)P * sizeof(int64_t), cudaMemcpyDeviceToHost, m->stream[q]));
        MULTI_OK(cudaStreamSynchronize(m->stream[q]));
        return 0;
    });
    if (rc) return rc;
    // ---- phase 2: every rank gathers what it owns and deposits it ----
    return for_each_rank_parallel(m, [&](int s) -> int {
        int64_t total = 0;
        for (int q = 0; q < P; q++) total += m->counts[q][s];
        if (total == 0) return 0;
        if (int e = grow_recv(m, s, total, masses != nullptr)) return e;
        RankBuf &d = m->buf[s];
        int64_t at = 0;
        for (int q = 0; q < P; q++) {
            const int64_t c = m->counts[q][s];
            if (c == 0) continue;
            int64_t off = 0;
            for (int t = 0; t < s; t++) off += m->counts[q][t];
            MULTI_OK(cudaMemcpyPeerAsync(d.recv_pos + 3 * at, m->dev[s], m->buf[q].sorted_pos + 3 * off, m->dev[q], (size_t)c * 12, m->stream[s]));
            if (masses)
                MULTI_OK(cudaMemcpyPeerAsync(d.recv_mass + at, m->dev[s], m->buf[q].sorted_mass + off, m->dev[q], (size_t)c * 4, m->stream[s]));
            at += c;
        }
        if (int e = genpk_deposit(m->ctx[s], 0, d.recv_pos, masses ? d.recv_mass : nullptr, total, mass, boxsize, 1)) return e;
        // the sources' sorted buffers are free again once these copies have run
        MULTI_OK(cudaStreamSynchronize(m->stream[s]));
        return 0;
    });
}

// fftw_execute + powerspectrum of gen-pk.cpp:233-234 over the N slabs; results on the host.
int genpk_multi_fft_power(genpk_multi *m, int nrbins, double *power, int *count, double *keffs, double total_mass, double total_mass2)
{
    if (!m || nrbins < 1 || !power || !count || !keffs) { set_error("genpk_multi_fft_power: bad arguments"); return 1; }
    if (m->n == 1) {
        if (use_device(m, 0)) { set_error("cudaSetDevice failed"); return 1; }
        if (int rc = genpk_fft_power(m->ctx[0], 0, nrbins, power, count, keffs, total_mass, total_mass2)) return rc;
        return genpk_synchronize(m->ctx[0]);
    }
    const int P = m->n;
    const bool fused = m->scatter && genpk_fused_xpass_supported(m->ctx[0], nrbins);
    // every rank's deposits are done (event a); the ghost plane of the low neighbour is pulled after that
    for (int r = 0; r < P; r++) {
        if (use_device(m, r)) { set_error("cudaSetDevice failed"); return 1; }
        if (int rc = grow_sums(m, r, nrbins)) return rc;
        MULTI_OK(cudaEventRecord(m->ev_a[r], m->stream[r]));
    }
    for (int r = 0; r < P; r++) {
        if (use_device(m, r)) { set_error("cudaSetDevice failed"); return 1; }
        MULTI_OK(cudaStreamWaitEvent(m->stream[r], m->ev_a[(r + P - 1) % P], 0));
        if (int rc = genpk_ghost_pull(m->ctx[r], 0)) return rc;
        if (fused) {
            // nobody may still read its transposed block (the previous call's x pass: event c) when the scatter starts
            for (int s = 0; s < P; s++)
                MULTI_OK(cudaStreamWaitEvent(m->stream[r], m->ev_c[s], 0));
            if (int rc = genpk_slab_fft_yz_scatter(m->ctx[r], 0)) return rc;
        } else {
            RankBuf &b = m->buf[r];
            const size_t bytes = genpk_slab_spectrum_bytes(m->ctx[r]);
            if (!b.send) MULTI_OK(cudaMalloc(&b.send, bytes));
            if (!b.spec) MULTI_OK(cudaMalloc(&b.spec, bytes));
            if (int rc = genpk_slab_fft_yz(m->ctx[r], 0)) return rc;
            if (int rc = genpk_slab_pack(m->ctx[r], 0, b.send)) return rc;
        }
        MULTI_OK(cudaEventRecord(m->ev_b[r], m->stream[r]));
    }
    for (int r = 0; r < P; r++) {
        if (use_device(m, r)) { set_error("cudaSetDevice failed"); return 1; }
        RankBuf &b = m->buf[r];
        for (int s = 0; s < P; s++)
            MULTI_OK(cudaStreamWaitEvent(m->stream[r], m->ev_b[s], 0));     // every rank has stored / packed
        if (fused) {
            if (int rc = genpk_slab_fftx_power_partial(m->ctx[r], genpk_slab_recv_buffer(m->ctx[r], nullptr), nrbins, b.d_sums)) return rc;
        } else {
            // the transpose by peer copies: block r of every rank's packed buffer
            const size_t blk = genpk_slab_spectrum_bytes(m->ctx[r]) / P;
            for (int s = 0; s < P; s++)
                MULTI_OK(cudaMemcpyPeerAsync((char *)b.spec + (size_t)s * blk, m->dev[r], (char *)m->buf[s].send + (size_t)r * blk, m->dev[s], blk,
                                             m->stream[r]));
            if (int rc = genpk_slab_fft_x(m->ctx[r], b.spec)) return rc;
            if (int rc = genpk_slab_power_partial(m->ctx[r], b.spec, nullptr, nrbins, b.d_sums)) return rc;
        }
        if (int rc = genpk_rejected_to(m->ctx[r], b.d_sums + 3 * nrbins)) return rc;
        MULTI_OK(cudaMemcpyAsync(b.h_sums, b.d_sums, (size_t)(3 * nrbins + 1) * sizeof(double), cudaMemcpyDeviceToHost, m->stream[r]));
        MULTI_OK(cudaEventRecord(m->ev_c[r], m->stream[r]));
    }
    std::vector<double> sums((size_t)3 * nrbins, 0.0);
    double rejected = 0;
    for (int r = 0; r < P; r++) {
        if (use_device(m, r)) { set_error("cudaSetDevice failed"); return 1; }
        MULTI_OK(cudaStreamSynchronize(m->stream[r]));
        for (int i = 0; i < 3 * nrbins; i++) sums[i] += m->buf[r].h_sums[i];
        rejected += m->buf[r].h_sums[3 * nrbins];
    }
    if (rejected > 0) {
        set_error("%.0f particles rejected (non-finite positions)", rejected);
        return 3;
    }
    return genpk_power_finalize(sums.data(), nrbins, total_mass, total_mass2, power, count, keffs);
}

// The whole per-type step of gen-pk.cpp:208-234 on host particle arrays, N GPUs.
int genpk_multi_pk_from_particles(genpk_multi *m, const float *positions, const float *masses, int64_t n, double mass, double boxsize,
                                  double total_mass, int nrbins, double *power, int *count, double *keffs)
{
    if (int rc = genpk_multi_grid_zero(m)) return rc;
    if (int rc = genpk_multi_deposit(m, positions, masses, n, mass, boxsize)) return rc;
    return genpk_multi_fft_power(m, nrbins, power, count, keffs, total_mass, total_mass);
}

}  // extern "C"
