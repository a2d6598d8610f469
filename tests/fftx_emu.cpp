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
