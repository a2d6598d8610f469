// Host emulation of one tile of fftx_power_kernel (genpk_b200/csrc/fftx_core.cuh): the
// same phase functions run thread by thread over plain arrays in place of shared memory.
// Built by tests/test_fftx_core.py with g++; test infrastructure only.
#include <string.h>
#include <vector>

#include "../genpk_b200/csrc/fftx_core.cuh"

using namespace genpk::fftx;

template <class PL>
static int emu_tile(const double *tile_in, const double *twiddle, double *spec_out, double *mod2_out, int kj, int kz0,
                    int nc, const float *sW, const unsigned *sT, int nrbins, float half_bpu, double *sP)
{
    constexpr int N = PL::N, C = PL::C, T = PL::T;
    const cd *in = reinterpret_cast<const cd *>(tile_in);        // [N][C]
    const cd *tw = reinterpret_cast<const cd *>(twiddle);        // [N]
    std::vector<cd> regs((size_t)PL::THREADS * EPT), regs2((size_t)PL::THREADS * EPT);
    std::vector<double> Ere((size_t)N * C), Eim((size_t)N * C), P((size_t)N * C);
    auto col = [](int tid) { return tid % C; };
    auto thr = [](int tid) { return tid / C; };
    // the base + part forms the kernel uses are the index maps
    for (int t = 0; t < T; t++)
        for (int i = 0; i < EPT; i++) {
            if (PL::ex1_w(t, i) != t + PL::ex1_w_part(i)) return 10;
            if (PL::ex1_r(t, i) != PL::ex1_r_base(t) + PL::ex1_r_part(i)) return 11;
            if (PL::ex2_w(t, i) != PL::ex2_w_base(t, (i % PL::R2) & 3) + PL::ex2_w_part(i)) return 12;
            if (PL::ex2_r(t, i) != PL::ex2_r_base(t, (i % PL::R3) & 3) + PL::ex2_r_part(i)) return 13;
            if (PL::out_k(t, i) != PL::out_k_base(t) + PL::out_k_part(i)) return 14;
        }
    // the kernel exchanges through a half-size buffer, lower half of the index space first: on the
    // pass-2 side of a plan whose thread owns one radix-R2 unit, all 16 indices of a thread must
    // share a half (exchange<..., W_UNI / R_UNI> in fftx_power.cu relies on it)
    if (PL::ONE_UNIT2)
        for (int t = 0; t < T; t++)
            for (int i = 1; i < EPT; i++) {
                if ((PL::ex1_r(t, i) >= N / 2) != (PL::ex1_r(t, 0) >= N / 2)) return 15;
                if ((PL::ex2_w(t, i) >= N / 2) != (PL::ex2_w(t, 0) >= N / 2)) return 16;
            }
    // fill + pass 1 + exchange-1 write
    for (int tid = 0; tid < PL::THREADS; tid++) {
        const int c = col(tid), t = thr(tid);
        cd *v = &regs[(size_t)tid * EPT];
        for (int i = 0; i < EPT; i++)
            v[i] = in[(size_t)PL::load_n(t, i) * C + c];
        PL::pass1(v, t, tw);
        for (int i = 0; i < EPT; i++) {
            Ere[(size_t)PL::ex1_w(t, i) * C + c] = v[i].x;
            Eim[(size_t)PL::ex1_w(t, i) * C + c] = v[i].y;
        }
    }
    // exchange-1 read + pass 2
    for (int tid = 0; tid < PL::THREADS; tid++) {
        const int c = col(tid), t = thr(tid);
        cd *w = &regs2[(size_t)tid * EPT];
        for (int i = 0; i < EPT; i++) {
            w[i].x = Ere[(size_t)PL::ex1_r(t, i) * C + c];
            w[i].y = Eim[(size_t)PL::ex1_r(t, i) * C + c];
        }
        PL::pass2(w, t, tw);
    }
    // exchange-2 write (after every thread has read exchange 1)
    for (int tid = 0; tid < PL::THREADS; tid++) {
        const int c = col(tid), t = thr(tid);
        const cd *w = &regs2[(size_t)tid * EPT];
        for (int i = 0; i < EPT; i++) {
            Ere[(size_t)PL::ex2_w(t, i) * C + c] = w[i].x;
            Eim[(size_t)PL::ex2_w(t, i) * C + c] = w[i].y;
        }
    }
    // exchange-2 read + pass 3 + |X|^2
    std::vector<char> hit((size_t)N * C, 0);
    for (int tid = 0; tid < PL::THREADS; tid++) {
        const int c = col(tid), t = thr(tid);
        cd *v = &regs[(size_t)tid * EPT];
        for (int i = 0; i < EPT; i++) {
            v[i].x = Ere[(size_t)PL::ex2_r(t, i) * C + c];
            v[i].y = Eim[(size_t)PL::ex2_r(t, i) * C + c];
        }
        PL::pass3(v);
        for (int i = 0; i < EPT; i++) {
            const int k = PL::out_k(t, i);
            if (k < 0 || k >= N || hit[(size_t)k * C + c])
                return 1;                                        // out_k must be a bijection per column
            hit[(size_t)k * C + c] = 1;
            spec_out[2 * ((size_t)k * C + c)] = v[i].x;
            spec_out[2 * ((size_t)k * C + c) + 1] = v[i].y;
            const int sl = PL::slot(k);
            if (sl < 0 || sl >= N)
                return 2;
            P[(size_t)sl * C + c] = v[i].x * v[i].x + v[i].y * v[i].y;
            mod2_out[(size_t)k * C + c] = P[(size_t)sl * C + c];
        }
    }
    // bin walk
    for (int tid = 0; tid < PL::THREADS; tid++) {
        const int c = col(tid), t = thr(tid);
        if (kz0 + c < nc)
            bin_walk<PL>(P.data(), t, c, kj, kz0 + c, N / 2, sW, sT, nrbins, half_bpu, sP, 1);
    }
    (void)T;
    return 0;
}

// the plans fftx_power.cu launches (keep in step with fftx_power_raw)
typedef Plan<4, 8, 8, 4096> P256;
typedef Plan<8, 8, 8, 4096> P512;
typedef Plan<16, 8, 8, 4096> P1024;
typedef Plan<16, 16, 8, 8192> P2048;
typedef Plan<16, 8, 8, 8192> P1024W;      // n = -1024: the one-CTA-per-SM tile

extern "C" int fftx_emu_columns(int n)
{
    switch (n) {
    case 256: return P256::C;
    case 512: return P512::C;
    case 1024: return P1024::C;
    case 2048: return P2048::C;
    case -1024: return P1024W::C;
    }
    return 0;
}

// tile_in: [N][C] complex (x, column); twiddle: exp(-2 pi i t/N), t < N; spec_out: [N][C] complex in
// natural k order; mod2_out: [N][C]; sP: nrbins doubles, accumulated into.
extern "C" int fftx_emu_tile(int n, const double *tile_in, const double *twiddle, double *spec_out, double *mod2_out,
                             int kj, int kz0, int nc, const float *sW, const unsigned *sT, int nrbins, float half_bpu,
                             double *sP)
{
    switch (n) {
    case 256: return emu_tile<P256>(tile_in, twiddle, spec_out, mod2_out, kj, kz0, nc, sW, sT, nrbins, half_bpu, sP);
    case 512: return emu_tile<P512>(tile_in, twiddle, spec_out, mod2_out, kj, kz0, nc, sW, sT, nrbins, half_bpu, sP);
    case 1024: return emu_tile<P1024>(tile_in, twiddle, spec_out, mod2_out, kj, kz0, nc, sW, sT, nrbins, half_bpu, sP);
    case 2048: return emu_tile<P2048>(tile_in, twiddle, spec_out, mod2_out, kj, kz0, nc, sW, sT, nrbins, half_bpu, sP);
    case -1024: return emu_tile<P1024W>(tile_in, twiddle, spec_out, mod2_out, kj, kz0, nc, sW, sT, nrbins, half_bpu, sP);
    }
    return -1;
}

// ---- one z tile of fft_zy_kernel (genpk_b200/csrc/fft_zy.cu): CZ padded rows of dims reals, each transformed in place
// into dims/2 + 1 complex values -- the half-length complex FFT of z[n] = x[2n] + i x[2n+1] with the plan's register
// passes, a third exchange that hands every thread eight pairs (Z[k], Z[dims/2 - k]), rfft_pair, rows written in place.
template <class PLZ> static int emu_rows(double *rows, const double *twiddle, const double *twiddle_half)
{
    constexpr int NZ = PLZ::N, CZ = PLZ::C, TZ = PLZ::T, ROW = NZ + 1;
    cd *R0 = reinterpret_cast<cd *>(rows);                         // [CZ][ROW]
    const cd *tw = reinterpret_cast<const cd *>(twiddle), *twh = reinterpret_cast<const cd *>(twiddle_half);
    std::vector<cd> regs((size_t)PLZ::THREADS * EPT), regs2((size_t)PLZ::THREADS * EPT), E((size_t)NZ * CZ);
    for (int tid = 0; tid < PLZ::THREADS; tid++) {
        const int c = tid % CZ, t = tid / CZ;
        cd *v = &regs[(size_t)tid * EPT];
        for (int i = 0; i < EPT; i++)
            v[i] = R0[(size_t)c * ROW + PLZ::load_n(t, i)];
        PLZ::pass1(v, t, twh);
    }
    for (int tid = 0; tid < PLZ::THREADS; tid++)
        for (int i = 0; i < EPT; i++)
            E[(size_t)(tid / CZ + PLZ::ex1_w_part(i)) * CZ + tid % CZ] = regs[(size_t)tid * EPT + i];
    for (int tid = 0; tid < PLZ::THREADS; tid++) {
        const int c = tid % CZ, t = tid / CZ;
        cd *w = &regs2[(size_t)tid * EPT];
        for (int i = 0; i < EPT; i++)
            w[i] = E[(size_t)(PLZ::ex1_r_base(t) + PLZ::ex1_r_part(i)) * CZ + c];
        PLZ::pass2(w, t, twh);
    }
    for (int tid = 0; tid < PLZ::THREADS; tid++)
        for (int i = 0; i < EPT; i++)
            E[(size_t)(PLZ::ex2_w_base(tid / CZ, (i % PLZ::R2) & 3) + PLZ::ex2_w_part(i)) * CZ + tid % CZ] = regs2[(size_t)tid * EPT + i];
    for (int tid = 0; tid < PLZ::THREADS; tid++) {
        const int c = tid % CZ, t = tid / CZ;
        cd *v = &regs[(size_t)tid * EPT];
        for (int i = 0; i < EPT; i++)
            v[i] = E[(size_t)(PLZ::ex2_r_base(t, (i % PLZ::R3) & 3) + PLZ::ex2_r_part(i)) * CZ + c];
        PLZ::pass3(v);
    }
    // third exchange, through a half-size buffer: lower half of the index space (all the k = t + TZ*j), then the upper
    // half (the partners NZ - k and Z[NZ/2]); which half a register goes to must follow from its index alone
    constexpr int PAIRS = NZ / 2 / TZ;
    if (PAIRS * 2 != EPT) return 20;
    std::vector<cd> H((size_t)(NZ / 2) * CZ), pa((size_t)PLZ::THREADS * PAIRS), pb((size_t)PLZ::THREADS * PAIRS), mid(CZ);
    for (int half = 0; half < 2; half++) {
        for (int tid = 0; tid < PLZ::THREADS; tid++)
            for (int i = 0; i < EPT; i++) {
                const int k = PLZ::out_k_base(tid / CZ) + PLZ::out_k_part(i);
                if (PLZ::out_k_upper(i) != (k >= NZ / 2)) return 21;
                if (PLZ::out_k_upper(i) == (half == 1))
                    H[(size_t)(k - half * (NZ / 2)) * CZ + tid % CZ] = regs[(size_t)tid * EPT + i];
            }
        for (int tid = 0; tid < PLZ::THREADS; tid++) {
            const int c = tid % CZ, t = tid / CZ;
            for (int j = 0; j < PAIRS; j++) {
                const int k = t + TZ * j;
                if (!half)
                    pa[(size_t)tid * PAIRS + j] = H[(size_t)k * CZ + c];
                else
                    pb[(size_t)tid * PAIRS + j] = k == 0 ? pa[(size_t)tid * PAIRS] : H[(size_t)(NZ / 2 - k) * CZ + c];
            }
            if (half && t == 0)
                mid[c] = H[c];
        }
    }
    std::vector<char> touched((size_t)CZ * ROW, 0);
    for (int tid = 0; tid < PLZ::THREADS; tid++) {
        const int c = tid % CZ, t = tid / CZ;
        cd *row = R0 + (size_t)c * ROW;
        const cd steps[8] = {rfft_step<0>(tw[t]), rfft_step<1>(tw[t]), rfft_step<2>(tw[t]), rfft_step<3>(tw[t]),
                             rfft_step<4>(tw[t]), rfft_step<5>(tw[t]), rfft_step<6>(tw[t]), rfft_step<7>(tw[t])};
        for (int j = 0; j < PAIRS; j++) {
            const int k = t + TZ * j;
            cd xk, xm;
            rfft_pair(pa[(size_t)tid * PAIRS + j], pb[(size_t)tid * PAIRS + j], steps[j], &xk, &xm);
            if (touched[(size_t)c * ROW + k] || touched[(size_t)c * ROW + NZ - k]) return 22;
            row[k] = xk;
            row[NZ - k] = xm;
            touched[(size_t)c * ROW + k] = touched[(size_t)c * ROW + NZ - k] = 1;
        }
        if (t == 0) {
            if (touched[(size_t)c * ROW + NZ / 2]) return 22;
            row[NZ / 2].x = mid[c].x;
            row[NZ / 2].y = -mid[c].y;
            touched[(size_t)c * ROW + NZ / 2] = 1;
        }
    }
    for (size_t i = 0; i < touched.size(); i++)
        if (!touched[i]) return 23;                                 // every output of every row was produced
    return 0;
}

// the row plans fft_zy.cu launches (keep in step with zy_dispatch); n = -1024 / 2048: the 8192-mode tiles
extern "C" int fftx_emu_row_count(int n)
{
    switch (n) {
    case 256: return Plan<2, 8, 8, 4096>::C;
    case 512: return Plan<4, 8, 8, 4096>::C;
    case 1024: return Plan<8, 8, 8, 4096>::C;
    case -1024: return Plan<8, 8, 8, 8192>::C;
    case 2048: return Plan<16, 8, 8, 8192>::C;
    }
    return 0;
}

// rows: [C][n + 2] doubles, transformed in place; twiddle: exp(-2 pi i t / n), t < n; twiddle_half: exp(-2 pi i t / (n/2))
extern "C" int fftx_emu_rows(int n, double *rows, const double *twiddle, const double *twiddle_half)
{
    switch (n) {
    case 256: return emu_rows<Plan<2, 8, 8, 4096>>(rows, twiddle, twiddle_half);
    case 512: return emu_rows<Plan<4, 8, 8, 4096>>(rows, twiddle, twiddle_half);
    case 1024: return emu_rows<Plan<8, 8, 8, 4096>>(rows, twiddle, twiddle_half);
    case -1024: return emu_rows<Plan<8, 8, 8, 8192>>(rows, twiddle, twiddle_half);
    case 2048: return emu_rows<Plan<16, 8, 8, 8192>>(rows, twiddle, twiddle_half);
    }
    return -1;
}

// ---- one tile of fftx_power2_kernel: the two-pass plan (32 elements per thread, one exchange) ----
template <class PL>
static int emu_tile2(const double *tile_in, const double *twiddle, double *spec_out, int kj, int kz0, int nc, const float *sW,
                     const unsigned *sT, int nrbins, float half_bpu, double *sP)
{
    constexpr int N = PL::N, C = PL::C, E2 = PL::EPT2;
    const cd *in = reinterpret_cast<const cd *>(tile_in);        // [N][C]
    const cd *tw = reinterpret_cast<const cd *>(twiddle);        // [N]
    std::vector<cd> regs((size_t)PL::THREADS * E2), X((size_t)N * C);
    std::vector<double> P((size_t)N * C);
    for (int tid = 0; tid < PL::THREADS; tid++) {
        const int c = tid % C, t = tid / C;
        cd *v = &regs[(size_t)tid * E2];
        for (int i = 0; i < E2; i++)
            v[i] = in[(size_t)PL::load_n(t, i) * C + c];
        PL::pass1(v, t, tw);
    }
    std::vector<char> hit((size_t)N * C, 0);
    for (int tid = 0; tid < PL::THREADS; tid++)
        for (int i = 0; i < E2; i++) {
            const size_t at = (size_t)PL::ex_w(tid / C, i) * C + tid % C;
            if (hit[at]) return 30;                              // the exchange map is a bijection per column
            hit[at] = 1;
            X[at] = regs[(size_t)tid * E2 + i];
        }
    std::fill(hit.begin(), hit.end(), 0);
    for (int tid = 0; tid < PL::THREADS; tid++) {
        const int c = tid % C, t = tid / C;
        cd w[E2];
        for (int i = 0; i < E2; i++)
            w[i] = X[(size_t)PL::ex_r(t, i) * C + c];
        PL::pass2(w);
        for (int i = 0; i < E2; i++) {
            const int k = PL::out_k(t, i);
            if (k < 0 || k >= N || hit[(size_t)k * C + c]) return 31;
            hit[(size_t)k * C + c] = 1;
            spec_out[2 * ((size_t)k * C + c)] = w[i].x;
            spec_out[2 * ((size_t)k * C + c) + 1] = w[i].y;
            const int sl = PL::slot(k);
            if (sl < 0 || sl >= N) return 32;
            P[(size_t)sl * C + c] = w[i].x * w[i].x + w[i].y * w[i].y;
        }
    }
    // a thread walks two blocks of eight |kx|: 8*(2t) .. and 8*(2t+1) ..
    for (int tid = 0; tid < PL::THREADS; tid++) {
        const int c = tid % C, t = tid / C;
        if (kz0 + c < nc)
            for (int h = 0; h < 2; h++)
                bin_walk<PL>(P.data(), 2 * t + h, c, kj, kz0 + c, N / 2, sW, sT, nrbins, half_bpu, sP, 1);
    }
    return 0;
}

extern "C" int fftx_emu_tile2(int n, const double *tile_in, const double *twiddle, double *spec_out, int kj, int kz0, int nc,
                              const float *sW, const unsigned *sT, int nrbins, float half_bpu, double *sP)
{
    switch (n) {
    case 512: return emu_tile2<Plan2<16, 4096>>(tile_in, twiddle, spec_out, kj, kz0, nc, sW, sT, nrbins, half_bpu, sP);
    case 1024: return emu_tile2<Plan2<32, 8192>>(tile_in, twiddle, spec_out, kj, kz0, nc, sW, sT, nrbins, half_bpu, sP);
    }
    return -1;
}
