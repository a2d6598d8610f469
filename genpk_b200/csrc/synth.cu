// Synthetic particle sets of the shapes BASELINE.json names, generated on the
// device so that 10^9-particle inputs never cross PCIe.  Every particle is a
// pure function of (kind, seed, n_side, particle index): any sub-range can be
// produced on any rank, and parity tests copy the same floats to the CPU oracle.
//
//   UNIFORM_RANDOM  x_a = box * (hash(seed, 3p+a) >> 40) * 2^-24   spatially incoherent order
//   LATTICE         q   = (i + 1/2) * box / n_side, p = (ix*n + iy)*n + iz  (z fastest)
//   CLUSTERED       q + psi(q),  psi = sum_{m<32} A_m nhat_m sin(2 pi n_m.q/box + phi_m):
//                   a smooth Zel'dovich-like displacement with per-component rms of
//                   ~2 grid cells, integer wave vectors |n_m| <= 8, A_m ~ |n_m|^-1.5
#include "common.cuh"

namespace genpk {

__host__ __device__ __forceinline__ uint64_t mix64(uint64_t z)
{
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

__host__ __device__ __forceinline__ uint64_t hash2(uint64_t seed, uint64_t idx)
{
    return mix64(idx + mix64(seed + 0x9E3779B97F4A7C15ull) * 0x9E3779B97F4A7C15ull);
}

constexpr int SYNTH_MODES = 32;
struct WaveSet {
    float nx[SYNTH_MODES], ny[SYNTH_MODES], nz[SYNTH_MODES];   // integer wave vector
    float ax[SYNTH_MODES], ay[SYNTH_MODES], az[SYNTH_MODES];   // amplitude * unit vector, in box units
    float phase[SYNTH_MODES];                                  // in turns
};

__global__ void __launch_bounds__(256) synth_kernel(int kind, uint64_t seed, int64_t n_side, int64_t first,
                                                    int64_t count, float box, WaveSet ws, float *pos)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < count; t += stride) {
        const int64_t p = first + t;
        float x, y, z;
        if (kind == GENPK_SYNTH_UNIFORM_RANDOM) {
            const float s = 1.0f / 16777216.0f;
            x = __fmul_rn(box, __fmul_rn((float)(hash2(seed, 3 * (uint64_t)p) >> 40), s));
            y = __fmul_rn(box, __fmul_rn((float)(hash2(seed, 3 * (uint64_t)p + 1) >> 40), s));
            z = __fmul_rn(box, __fmul_rn((float)(hash2(seed, 3 * (uint64_t)p + 2) >> 40), s));
        } else {
            const int64_t iz = p % n_side, iy = (p / n_side) % n_side, ix = p / (n_side * n_side);
            const double inv = 1.0 / (double)n_side;
            // lattice site in box units
            double qx = ((double)ix + 0.5) * inv, qy = ((double)iy + 0.5) * inv, qz = ((double)iz + 0.5) * inv;
            if (kind == GENPK_SYNTH_CLUSTERED) {
                double dx = 0, dy = 0, dz = 0;
#pragma unroll 4
                for (int m = 0; m < SYNTH_MODES; m++) {
                    double ph = ws.nx[m] * qx + ws.ny[m] * qy + ws.nz[m] * qz + ws.phase[m];
                    ph -= floor(ph);
                    const float s = sinpif(2.0f * (float)ph);
                    dx += ws.ax[m] * s;
                    dy += ws.ay[m] * s;
                    dz += ws.az[m] * s;
                }
                qx += dx; qy += dy; qz += dz;
                qx -= floor(qx); qy -= floor(qy); qz -= floor(qz);
            }
            x = (float)(qx * box);
            y = (float)(qy * box);
            z = (float)(qz * box);
        }
        pos[3 * t] = x;
        pos[3 * t + 1] = y;
        pos[3 * t + 2] = z;
    }
}

static WaveSet make_waves(uint64_t seed, double grid_dims)
{
    WaveSet ws;
    double amp[SYNTH_MODES], sum2 = 0;
    uint64_t ctr = 0;
    for (int m = 0; m < SYNTH_MODES; m++) {
        int n[3];
        double len;
        do {
            for (int a = 0; a < 3; a++)
                n[a] = (int)(hash2(seed ^ 0xC1057E7ull, ctr++) % 17) - 8;
            len = sqrt((double)(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]));
        } while (len < 0.5 || len > 8.0);
        amp[m] = pow(len, -1.5);
        sum2 += amp[m] * amp[m];
        ws.nx[m] = (float)n[0]; ws.ny[m] = (float)n[1]; ws.nz[m] = (float)n[2];
        ws.ax[m] = (float)(n[0] / len); ws.ay[m] = (float)(n[1] / len); ws.az[m] = (float)(n[2] / len);
        ws.phase[m] = (float)((hash2(seed ^ 0xC1057E7ull, ctr++) >> 40) / 16777216.0);
    }
    // per-component rms of psi = sqrt(sum A^2 / 6); target 2 cells = 2/grid_dims box units
    const double norm = (2.0 / grid_dims) / sqrt(sum2 / 6.0);
    for (int m = 0; m < SYNTH_MODES; m++) {
        const float a = (float)(amp[m] * norm);
        ws.ax[m] *= a; ws.ay[m] *= a; ws.az[m] *= a;
    }
    return ws;
}

}  // namespace genpk

extern "C" int genpk_synth_particles(int kind, uint64_t seed, int64_t n_side, int64_t first, int64_t count,
                                     double boxsize, double grid_dims, float *pos_dev, void *cuda_stream)
{
    using namespace genpk;
    if (kind < 0 || kind > 2 || n_side < 1 || count < 0 || first < 0 || !(boxsize > 0) || !(grid_dims >= 1)) {
        set_error("genpk_synth_particles: bad arguments");
        return 1;
    }
    if (count == 0)
        return 0;
    WaveSet ws = {};
    if (kind == GENPK_SYNTH_CLUSTERED)
        ws = make_waves(seed, grid_dims);
    int64_t blocks = (count + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    synth_kernel<<<(int)blocks, 256, 0, (cudaStream_t)cuda_stream>>>(kind, seed, n_side, first, count, (float)boxsize, ws, pos_dev);
    GENPK_CUDA_OK(cudaGetLastError());
    return 0;
}
