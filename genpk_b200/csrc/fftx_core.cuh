// Thread-level pieces of the fused x-pass: a length-N complex FFT along x of a
// tile [N][C] of the spectrum (C adjacent kz columns, N*C = 8192 modes), followed by
// the |delta_k|^2 binning of powerspectrum.c:35-110 on the tile -- so the x-transformed
// spectrum is never written back to HBM.
//
// This header compiles for the device (fftx_power.cu) AND for the host
// (tests/fftx_emu.cpp, which runs the phases thread by thread over a plain array in
// place of shared memory and checks them against numpy): the index maps, twiddles,
// butterflies, swizzles and the bin walk are verified on the CPU, the kernel adds only
// the data movement and the barriers between phases.
//
// FFT: N = R1*R2*R3 (Cooley-Tukey, three register passes, two exchanges), forward and
// unnormalised, exp(-2 pi i n k / N) -- the convention of FFTW/cuFFT (gen-pk.cpp:193,233).
// A thread owns 16 elements in every pass; T = N/16 threads work on one column.
//
//   n = M1*j + n2        (M1 = R2*R3)   pass 1: DFT-R1 over j,  times W_N^(n2*k1)
//   n2 = R3*m1 + m2                     pass 2: DFT-R2 over m1, times W_N^(R1*m2*q1)
//                                       pass 3: DFT-R3 over m2
//   k = k1 + R1*(q1 + R2*q2)
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define FX_HD __host__ __device__ __forceinline__
#else
#define FX_HD inline
#endif

namespace genpk {
namespace fftx {

#if defined(__CUDACC__)
typedef double2 cd;
#else
struct alignas(16) cd {
    double x, y;
};
#endif

FX_HD cd cadd(cd a, cd b) { cd r; r.x = a.x + b.x; r.y = a.y + b.y; return r; }
FX_HD cd csub(cd a, cd b) { cd r; r.x = a.x - b.x; r.y = a.y - b.y; return r; }
FX_HD cd cmul(cd a, cd w)
{
    cd r;
    r.x = a.x * w.x - a.y * w.y;
    r.y = a.x * w.y + a.y * w.x;
    return r;
}

// a * exp(-2 pi i K/16), K in [0, 8), constants folded at compile time
template <int K> FX_HD cd mulw16(cd a)
{
    const double c1 = 0.92387953251128673848, s1 = 0.38268343236508978178, h = 0.70710678118654752440;
    cd r;
    if (K == 0) { r = a; }
    else if (K == 4) { r.x = a.y; r.y = -a.x; }
    else if (K == 2) { r.x = (a.x + a.y) * h; r.y = (a.y - a.x) * h; }
    else if (K == 6) { r.x = (a.y - a.x) * h; r.y = -(a.x + a.y) * h; }
    else if (K == 1) { r.x = a.x * c1 + a.y * s1; r.y = a.y * c1 - a.x * s1; }
    else if (K == 3) { r.x = a.x * s1 + a.y * c1; r.y = a.y * s1 - a.x * c1; }
    else if (K == 5) { r.x = a.y * c1 - a.x * s1; r.y = -(a.y * s1) - a.x * c1; }
    else { r.x = a.y * s1 - a.x * c1; r.y = -(a.y * c1) - a.x * s1; }
    return r;
}

FX_HD cd csqr(cd a)
{
    cd r;
    r.x = (a.x + a.y) * (a.x - a.y);
    r.y = 2.0 * (a.x * a.y);
    return r;
}

// v[k] *= w^k for k = 1 .. R-1, the powers built from w by squaring and multiplying (w, w^2, w^4, w^8 and the
// products along the binary digits of k: at most five roundings deep).  One table lookup per butterfly unit instead
// of R-1: the scattered 16-byte twiddle loads were a third of the kernels' load/store wavefronts and most of their
// scoreboard stalls (profiles/r02/README.md).
// v[k] *= w^k for k = 1 .. 31: w^k = (w^8)^(k/8) * w^(k%8), ten powers kept, at most five roundings deep
FX_HD void mul_twiddle_powers32(cd *v, cd w)
{
    cd b[8], a[4];
    b[1] = w;
    b[2] = csqr(w);
    b[3] = cmul(b[2], w);
    b[4] = csqr(b[2]);
    b[5] = cmul(b[4], w);
    b[6] = csqr(b[3]);
    b[7] = cmul(b[6], w);
    a[1] = csqr(b[4]);
    a[2] = csqr(a[1]);
    a[3] = cmul(a[2], a[1]);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int k = 1; k < 32; k++) {
        const int hi = k / 8, lo = k % 8;
        cd t;
        if (hi == 0) t = b[lo];
        else if (lo == 0) t = a[hi];
        else t = cmul(a[hi], b[lo]);
        v[k] = cmul(v[k], t);
    }
}

template <int R> FX_HD void mul_twiddle_powers(cd *v, cd w)
{
    static_assert(R == 2 || R == 4 || R == 8 || R == 16, "radix");
    v[1] = cmul(v[1], w);
    if (R >= 4) {
        const cd w2 = csqr(w);
        v[2] = cmul(v[2], w2);
        v[3] = cmul(v[3], cmul(w2, w));
        if (R >= 8) {
            const cd w4 = csqr(w2);
            v[4] = cmul(v[4], w4);
            v[5] = cmul(v[5], cmul(w4, w));
            const cd w6 = cmul(w4, w2);
            v[6] = cmul(v[6], w6);
            v[7] = cmul(v[7], cmul(w6, w));
            if (R >= 16) {
                const cd w8 = csqr(w4);
                v[8] = cmul(v[8], w8);
                v[9] = cmul(v[9], cmul(w8, w));
                const cd w10 = cmul(w8, w2);
                v[10] = cmul(v[10], w10);
                v[11] = cmul(v[11], cmul(w10, w));
                const cd w12 = cmul(w8, w4);
                v[12] = cmul(v[12], w12);
                v[13] = cmul(v[13], cmul(w12, w));
                const cd w14 = cmul(w8, w6);
                v[14] = cmul(v[14], w14);
                v[15] = cmul(v[15], cmul(w14, w));
            }
        }
    }
}

// a * exp(-2 pi i K/32), K in [0, 16): the even ones are the sixteenth roots above
template <int K> FX_HD cd mulw32(cd a)
{
    static_assert(K >= 0 && K < 16, "first half turn");
    if (K % 2 == 0)
        return mulw16<K / 2>(a);
    constexpr double cs[16][2] = {{1.00000000000000000000, 0.00000000000000000000}, {0.98078528040323043058, 0.19509032201612824808}, {0.92387953251128673848, 0.38268343236508978178}, {0.83146961230254523567, 0.55557023301960217765}, {0.70710678118654757274, 0.70710678118654746172}, {0.55557023301960228867, 0.83146961230254523567}, {0.38268343236508983729, 0.92387953251128673848}, {0.19509032201612833135, 0.98078528040323043058}, {0.00000000000000006123, 1.00000000000000000000}, {-0.19509032201612819257, 0.98078528040323043058}, {-0.38268343236508972627, 0.92387953251128673848}, {-0.55557023301960195560, 0.83146961230254545772}, {-0.70710678118654746172, 0.70710678118654757274}, {-0.83146961230254534669, 0.55557023301960217765}, {-0.92387953251128673848, 0.38268343236508989280}, {-0.98078528040323043058, 0.19509032201612860891}};
    cd r;
    r.x = a.x * cs[K][0] + a.y * cs[K][1];
    r.y = a.y * cs[K][0] - a.x * cs[K][1];
    return r;
}

template <int R> struct Dft;
template <int R, int K> struct Comb {
    static FX_HD void run(cd *v, const cd *e, const cd *o)
    {
        const cd t = mulw32<K * (32 / R)>(o[K]);
        v[K] = cadd(e[K], t);
        v[K + R / 2] = csub(e[K], t);
        Comb<R, K - 1>::run(v, e, o);
    }
};
template <int R> struct Comb<R, -1> {
    static FX_HD void run(cd *, const cd *, const cd *) {}
};
// forward DFT of R values in place, natural order in and out (radix-2 decimation in time,
// fully unrolled: every index below is a compile-time constant)
template <int R> struct Dft {
    static FX_HD void run(cd *v)
    {
        cd e[R / 2], o[R / 2];
#pragma unroll
        for (int k = 0; k < R / 2; k++) {
            e[k] = v[2 * k];
            o[k] = v[2 * k + 1];
        }
        Dft<R / 2>::run(e);
        Dft<R / 2>::run(o);
        Comb<R, R / 2 - 1>::run(v, e, o);
    }
};
template <> struct Dft<1> {
    static FX_HD void run(cd *) {}
};

constexpr int EPT = 16;               // elements per thread

// Real-input transform of length 2*Nc from the length-Nc complex transform Z of z[n] = x[2n] + i x[2n+1]:
// with a = Z[k], b = Z[Nc - k] (b = a for k = 0) and w = exp(-2 pi i k / (2 Nc)),
//   E = (a + conj b) / 2,  O = -i (a - conj b) / 2      (transforms of the even and the odd samples)
//   X[k] = E + w O,        X[Nc - k] = conj(E - w O)
// k = 0 gives X[0] = Re a + Im a and X[Nc] = Re a - Im a; the self-paired k = Nc/2 gives conj(a).
// exp(-2 pi i j / 32), j < 8: a thread's pairs are k = t + (NZ/16) j, so its twiddles exp(-2 pi i k / (2 NZ)) are
// exp(-2 pi i t / (2 NZ)) times these, whatever the length
template <int J> FX_HD cd rfft_step(cd w)
{
    static_assert(J >= 0 && J < 8, "eight pairs per thread");
    constexpr double cs[8][2] = {{1.00000000000000000000, -0.00000000000000000000}, {0.98078528040323043058, -0.19509032201612824808}, {0.92387953251128673848, -0.38268343236508978178}, {0.83146961230254523567, -0.55557023301960217765}, {0.70710678118654757274, -0.70710678118654746172}, {0.55557023301960228867, -0.83146961230254523567}, {0.38268343236508983729, -0.92387953251128673848}, {0.19509032201612833135, -0.98078528040323043058}};
    if (J == 0) return w;
    cd c;
    c.x = cs[J][0];
    c.y = cs[J][1];
    return cmul(w, c);
}

FX_HD void rfft_pair(cd a, cd b, cd w, cd *xk, cd *xm)
{
    cd e, o;
    e.x = 0.5 * (a.x + b.x);
    e.y = 0.5 * (a.y - b.y);
    o.x = 0.5 * (a.y + b.y);
    o.y = -0.5 * (a.x - b.x);
    const cd wo = cmul(o, w);
    xk->x = e.x + wo.x;
    xk->y = e.y + wo.y;
    xm->x = e.x - wo.x;
    xm->y = -(e.y - wo.y);
}

// TILE: modes per tile (16 B each).  4096 -> 256 threads and ~110 KB of shared memory per
// CTA, two CTAs per SM whose barriers and exchanges interleave; 8192 -> one CTA of 512.
template <int R1_, int R2_, int R3_, int TILE_> struct Plan {
    static constexpr int R1 = R1_, R2 = R2_, R3 = R3_;
    static constexpr int N = R1 * R2 * R3;
    static constexpr int M1 = R2 * R3;
    static constexpr int TILE = TILE_;
    static constexpr int THREADS = TILE / EPT;
    static constexpr int T = N / EPT;             // threads per column
    static constexpr int C = TILE / N;            // columns per tile
    static_assert(C >= 4, "a tile row is at least 64 contiguous bytes");
    static constexpr bool ONE_UNIT2 = EPT == R2;  // a thread owns one radix-R2 unit in pass 2: its 16 exchange indices share a half
    static constexpr int LOG_R1 = R1 == 16 ? 4 : (R1 == 8 ? 3 : (R1 == 4 ? 2 : 1));
    static_assert(EPT % R1 == 0 && EPT % R2 == 0 && EPT % R3 == 0, "a thread owns whole butterflies");
    static_assert(R2 >= 4 && R3 >= 4, "the exchange swizzle uses two bits");

    // x index of a thread's i-th element before pass 1
    static FX_HD int load_n(int t, int i) { return M1 * (i % R1) + t + T * (i / R1); }

    static FX_HD void pass1(cd *v, int t, const cd *tw)
    {
#pragma unroll
        for (int u = 0; u < EPT / R1; u++) {
            Dft<R1>::run(v + u * R1);
            mul_twiddle_powers<R1>(v + u * R1, tw[t + T * u]);                 // W_N^(n2*k1), n2 = t + T*u
        }
    }
    // exchange 1: y[k1][n2]
    static FX_HD int ex1_w(int t, int i) { return (i % R1) * M1 + t + T * (i / R1); }
    static FX_HD int ex1_r(int t, int i)
    {
        const int v2 = t + T * (i / R2), m1 = i % R2;
        return (v2 / R3) * M1 + R3 * m1 + (v2 % R3);
    }
    static FX_HD void pass2(cd *w, int t, const cd *tw)
    {
#pragma unroll
        for (int u = 0; u < EPT / R2; u++) {
            Dft<R2>::run(w + u * R2);
            mul_twiddle_powers<R2>(w + u * R2, tw[R1 * ((t + T * u) % R3)]);   // W_N^(R1*m2*q1), m2 = (t + T*u) % R3
        }
    }
    // exchange 2: z[k1][q1][m2], m2 swizzled by q1 so that both the writers (lanes along
    // m2) and the readers (lanes along q1) spread over the banks
    static FX_HD int ex2_w(int t, int i)
    {
        const int v2 = t + T * (i / R2), q1 = i % R2;
        return (v2 / R3) * M1 + q1 * R3 + ((v2 % R3) ^ (q1 & 3));
    }
    static FX_HD int ex2_r(int t, int i)
    {
        const int v3 = t + T * (i / R3), m2 = i % R3;
        const int q1 = v3 % R2;
        return (v3 / R2) * M1 + q1 * R3 + (m2 ^ (q1 & 3));
    }
    static FX_HD void pass3(cd *v)
    {
#pragma unroll
        for (int u = 0; u < EPT / R3; u++)
            Dft<R3>::run(v + u * R3);
    }
    // FFT index of a thread's i-th element after pass 3
    static FX_HD int out_k(int t, int i)
    {
        const int v3 = t + T * (i / R3), q2 = i % R3;
        return (v3 / R2) + R1 * ((v3 % R2) + R2 * q2);
    }
    // The same maps split into a per-thread base and a compile-time part, base(t) + part(i),
    // which is what the kernel keeps in registers (tests/fftx_emu.cpp checks them against
    // the definitions above for every t and i).
    static_assert(T % R2 == 0 && T % R3 == 0, "a thread's units share m2 (pass 2) and q1 (pass 3)");
    static FX_HD int ex1_w_part(int i) { return (i % R1) * M1 + T * (i / R1); }                       // base: t
    static FX_HD int ex1_r_base(int t) { return (t / R3) * M1 + (t % R3); }
    static FX_HD int ex1_r_part(int i) { return ((T * (i / R2)) / R3) * M1 + R3 * (i % R2); }
    static FX_HD int ex2_w_base(int t, int s) { return (t / R3) * M1 + ((t % R3) ^ s); }             // s = (i % R2) & 3
    static FX_HD int ex2_w_part(int i) { return ((T * (i / R2)) / R3) * M1 + (i % R2) * R3; }
    static FX_HD int ex2_r_base(int t, int j) { return (t / R2) * M1 + (t % R2) * R3 + (j ^ (t & 3)); }   // j = (i % R3) & 3
    static FX_HD int ex2_r_part(int i) { return ((T * (i / R3)) / R2) * M1 + ((i % R3) & ~3); }
    static FX_HD int out_k_base(int t) { return t / R2 + R1 * (t % R2); }
    static FX_HD int out_k_part(int i) { return (T / R2) * (i / R3) + R1 * R2 * (i % R3); }
    // out_k >= N/2 depends on the register index alone: out_k_base(t) + (T/R2)*(i/R3) < R1*R2 (checked in fftx_emu.cpp)
    static FX_HD bool out_k_upper(int i) { return (i % R3) >= R3 / 2; }
    // where |X[k]|^2 waits for the bin walk (bank swizzle only; any bijection is correct)
    static FX_HD int slot(int k) { return k ^ (((k >> 3) ^ (R1 == 8 ? 0 : (k >> LOG_R1))) & 3); }
};

// Two-pass plan: N = 32 * R2 with 32 elements per thread -- one radix-32 pass, ONE exchange, one radix-R2 pass --
// for the kernels that hold a tile with half as many threads and twice the registers (fftx_power2_kernel).
//   n = R2*j + n2 (j < 32)      pass 1: DFT-32 over j, times W_N^(n2*k1)
//   k = k1 + 32*k2              pass 2: DFT-R2 over n2
template <int R2_, int TILE_> struct Plan2 {
    static constexpr int R1 = 32, R2 = R2_;
    static constexpr int EPT2 = 32;
    static constexpr int N = R1 * R2;
    static constexpr int TILE = TILE_;
    static constexpr int THREADS = TILE / EPT2;
    static constexpr int T = N / EPT2;            // threads per column
    static constexpr int C = TILE / N;
    static constexpr int LOG_R1 = 5;
    static_assert(R2 == 16 || R2 == 32, "N = 512 or 1024");
    static_assert(T == R2, "a thread owns n2 = t in pass 1");
    static FX_HD int load_n(int t, int i) { return R2 * i + t; }
    static FX_HD void pass1(cd *v, int t, const cd *tw)
    {
        Dft<32>::run(v);
        mul_twiddle_powers32(v, tw[t]);                              // W_N^(n2*k1), n2 = t
    }
    // exchange: y[k1][n2] at k1*R2 + n2
    static FX_HD int ex_w(int t, int i) { return i * R2 + t; }
    // pass 2: EPT2/R2 units; unit u of thread t works on k1 = t + T*u
    static FX_HD int ex_r(int t, int i) { return (t + T * (i / R2)) * R2 + (i % R2); }
    static FX_HD void pass2(cd *w)
    {
#pragma unroll
        for (int u = 0; u < EPT2 / R2; u++)
            Dft<R2>::run(w + u * R2);
    }
    static FX_HD int out_k(int t, int i) { return t + T * (i / R2) + R1 * (i % R2); }
    // where |X[k]|^2 waits for the bin walk: neighbouring blocks of 16 on opposite bank halves
    static FX_HD int slot(int k) { return k ^ ((k >> 4) & 1); }
};

// ---------------------------------------------------------------------------------
// Bin walk (powerspectrum.c:56-99 for one (ky, kz) column of the tile): thread t owns
// |kx| = 8t+1 .. 8t+8 (thread 0 also kx = 0); the +kx and -kx modes share bin and window,
// so their |X|^2 are added first.  Along the walk |k| only grows, so the bin only moves
// forward: the thread keeps a run (bin, sum) in registers and touches its column's
// histogram when the bin changes.  Histograms are private to a tile column (hist[b*hstride]):
// the lanes of a warp that walk the same |kx| range of neighbouring kz columns -- and so
// change bins together -- never meet on an address.
// ---------------------------------------------------------------------------------
#if defined(__CUDA_ARCH__)
#define FX_HIST_ADD(ptr, val) atomicAdd((ptr), (val))
#define FX_LOGF(x) __logf(x)
#define FX_FMUL(a, b) __fmul_rn((a), (b))
#else
#define FX_HIST_ADD(ptr, val) (*(ptr) += (val))
#define FX_LOGF(x) logf(x)
#define FX_FMUL(a, b) ((a) * (b))
#endif

// P: the tile's |X|^2 in slot order, [N][C] doubles.  c: column of this thread, kz its global kz.
template <class PL>
FX_HD void bin_walk(const double *P, int t, int c, int kj, int kz, int dims_half, const float *sW, const unsigned *sT,
                    int nrbins, float half_bpu, double *hist, int hstride)
{
    constexpr int N = PL::N, C = PL::C;
    const double mult = (kz == 0 || kz == dims_half) ? 1.0 : 2.0;      // powerspectrum.c:59-89
    const float fj = sW[kj < 0 ? -kj : kj], fz = sW[kz];
    const unsigned base2 = (unsigned)(kj * kj) + (unsigned)(kz * kz);
    // first step: kx = 0 belongs to thread 0, and the DC mode is skipped (powerspectrum.c:63)
    const int s0 = (t == 0 && base2 != 0) ? 0 : 1;
    // bin of the first mode: a float guess made exact by walking the threshold table
    const unsigned k2first = base2 + (unsigned)((8 * t + s0) * (8 * t + s0));
    int b = (int)(half_bpu * FX_LOGF((float)k2first));
    b = b < 0 ? 0 : (b > nrbins - 1 ? nrbins - 1 : b);
    while (k2first >= sT[b + 1]) b++;
    while (k2first < sT[b]) b--;
    // every operand of the walk first (independent loads), then the sequential run logic on
    // registers: the histogram atomics below would otherwise order themselves between the loads
    double mv[9];
    float wv[9];
    // |X|^2 slots of the blocks this thread reads: 8t.., 8(t+1) and the mirror block of -kx.
    // With R1 >= 8 the swizzle of slot() is constant over an aligned block of 8.
    const int blk_m = N / 8 - t - 1;
    const int g0 = PL::slot(8 * t) ^ (8 * t), g1 = PL::slot(8 * t + 8) ^ (8 * t + 8), gm = PL::slot(8 * blk_m) ^ (8 * blk_m);
#pragma unroll
    for (int s = 0; s <= 8; s++) {
        const int a = 8 * t + s;                                        // |kx|
        int sa, sm;
        if (PL::LOG_R1 >= 3) {
            sa = s < 8 ? 8 * t + (s ^ g0) : 8 * t + 8 + g1;
            sm = 8 * blk_m + (((8 - s) & 7) ^ gm);                      // N - a = 8*blk_m + (8 - s), s >= 1
        } else {
            sa = PL::slot(a);
            sm = PL::slot(s > 0 ? N - a : 0);
        }
        double m = P[sa * C + c];
        if (s > 0 && a < N / 2)
            m += P[sm * C + c];
        mv[s] = m;
        wv[s] = sW[a];
    }
    unsigned hi = sT[b + 1];
    double p = 0.0;
#pragma unroll
    for (int s = 0; s <= 8; s++) {
        if (s >= s0) {
            const int a = 8 * t + s;
            const unsigned k2 = base2 + (unsigned)(a * a);
            if (k2 >= hi) {
                FX_HIST_ADD(&hist[b * hstride], p * mult);
                do b++; while (k2 >= sT[b + 1]);
                hi = sT[b + 1];
                p = 0.0;
            }
            double w = (double)FX_FMUL(FX_FMUL(wv[s], fj), fz);         // (iwx*iwy)*iwz in float, promoted: fieldize.cpp:129-132
            w = w * w;                                                  // invwindow() = prod^2
            w = w * w;                                                  // pow(invwindow,2), powerspectrum.c:68
            p = fma(mv[s], w, p);
        }
    }
    FX_HIST_ADD(&hist[b * hstride], p * mult);
}

}  // namespace fftx
}  // namespace genpk
