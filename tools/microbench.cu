// Primitive throughputs on B200 that decide the deposit / binning kernel designs
// (DESIGN.md "Measured primitives").  Standalone: nvcc -arch=sm_100a microbench.cu.
//   1. streaming copy bandwidth (sanity vs MEASURED_PEAKS.json)
//   2. global red.add.f64 / red.add.u64 : random addresses in an L2-resident
//      window, random over a DRAM-sized array, CIC-like coherent pattern
//   3. shared-memory accumulation: f64 atomicAdd (CAS loop), u64 atomicAdd (CAS
//      loop), 2 x native u32 atomicAdd with carry, f32 atomicAdd
//   4. __match_any_sync rate
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint64_t mix64(uint64_t z)
{
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

__global__ void copy_kernel(const double2 *__restrict__ a, double2 *__restrict__ b, size_t n)
{
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        b[i] = a[i];
}

// mode 0: uniformly random address in [0, span); mode 1: CIC-like: thread t touches
// 8 cells around a base that advances smoothly with t (z fastest)
template <typename T>
__global__ void red_kernel(T *grid, size_t span, size_t nops, int mode, int dims)
{
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nops / 8; i += stride) {
        if (mode == 0) {
#pragma unroll
            for (int c = 0; c < 8; c++) {
                size_t a = mix64(i * 8 + c) % span;
                atomicAdd(&grid[a], (T)1);
            }
        } else {
            // particle i sits in cell (x,y,z) of a dims^3 grid with z fastest, 1 particle per cell
            size_t z = i % dims, y = (i / dims) % dims, x = (i / ((size_t)dims * dims)) % dims;
            size_t fd = dims + 2;
#pragma unroll
            for (int c = 0; c < 8; c++) {
                size_t xx = (x + (c & 1)) % dims, yy = (y + ((c >> 1) & 1)) % dims, zz = (z + (c >> 2)) % dims;
                atomicAdd(&grid[(xx * dims + yy) * fd + zz], (T)1);
            }
        }
    }
}

// Shared-memory accumulation into a TILE-cell tile, random cells, NOPS per thread.
template <int KIND>   // 0 f64 CAS-add, 1 u64 add, 2 2xu32 with carry, 3 f32 add, 4 u32 add
__global__ void __launch_bounds__(256) smem_kernel(double *out, int tile_cells, int nops, int conflict_shift)
{
    extern __shared__ double s_tile[];
    for (int i = threadIdx.x; i < tile_cells; i += blockDim.x)
        s_tile[i] = 0;
    __syncthreads();
    uint64_t h = mix64(blockIdx.x * 1024ull + threadIdx.x);
    for (int k = 0; k < nops; k++) {
        h = mix64(h + k);
        // conflict_shift > 0 makes groups of 2^shift lanes hit the same cell
        uint64_t hh = conflict_shift ? mix64((uint64_t)(threadIdx.x >> conflict_shift) * 7919 + k + blockIdx.x) : h;
        int a = (int)(hh % (uint64_t)tile_cells);
        if (KIND == 0) {
            atomicAdd(&s_tile[a], 1.0);
        } else if (KIND == 1) {
            atomicAdd(reinterpret_cast<unsigned long long *>(s_tile) + a, 1ull << 20);
        } else if (KIND == 2) {
            unsigned *w = reinterpret_cast<unsigned *>(s_tile) + 2 * a;
            const unsigned lo = 0x90000000u, hi = 3u;
            unsigned old = atomicAdd(w, lo);
            unsigned carry = (old + lo < old) ? 1u : 0u;
            atomicAdd(w + 1, hi + carry);
        } else if (KIND == 3) {
            atomicAdd(reinterpret_cast<float *>(s_tile) + a, 1.0f);
        } else {
            atomicAdd(reinterpret_cast<unsigned *>(s_tile) + a, 1u);
        }
    }
    __syncthreads();
    double acc = 0;
    for (int i = threadIdx.x; i < tile_cells; i += blockDim.x)
        acc += s_tile[i];
    if (acc == 123.456)
        out[0] = acc;
}

__global__ void match_kernel(int *out, int nops)
{
    unsigned acc = 0;
    uint64_t h = mix64(blockIdx.x * 1024ull + threadIdx.x);
    for (int k = 0; k < nops; k++) {
        h = mix64(h);
        acc += __match_any_sync(0xffffffffu, (int)(h & 15));
    }
    if (acc == 0x12345u)
        out[0] = acc;
}

// TMA bulk reduction: cp.reduce.async.bulk.global.shared::cta.add.f64 of `bytes` per op
// from a CTA-private shared buffer into a window of the grid.  One elected thread per
// warp issues; span_cells bounds the window (L2-resident or not).
__global__ void __launch_bounds__(256) bulk_red_kernel(double *grid, size_t span_cells, int bytes, int nops, int coherent)
{
    extern __shared__ __align__(128) double s_buf[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cells = bytes / 8;
    double *mine = s_buf + warp * cells;
    for (int i = lane; i < cells; i += 32)
        mine[i] = 1.0;
    __syncwarp();
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (lane == 0) {
        const unsigned saddr = (unsigned)__cvta_generic_to_shared(mine);
        uint64_t h = mix64(blockIdx.x * 64ull + warp);
        size_t base = (size_t)(blockIdx.x * 8 + warp) * (size_t)cells * nops;
        for (int k = 0; k < nops; k++) {
            size_t cell;
            if (coherent) {
                cell = (base + (size_t)k * cells) % (span_cells - cells);
            } else {
                h = mix64(h + k);
                cell = (h % (span_cells - cells));
            }
            cell &= ~(size_t)1;                                   // 16-byte alignment
            double *dst = grid + cell;
            asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [%0], [%1], %2;" ::"l"(dst), "r"(saddr),
                         "r"(bytes)
                         : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            if ((k & 7) == 7)
                asm volatile("cp.async.bulk.wait_group.read 4;" ::: "memory");
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}

template <typename F>
static float time_ms(F f, int reps = 3)
{
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    f();
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; r++) {
        cudaEventRecord(a);
        f();
        cudaEventRecord(b);
        CK(cudaEventSynchronize(b));
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    return best;
}

int main()
{
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    printf("device %s SMs %d L2 %d MB smem/SM %zu KB\n", prop.name, prop.multiProcessorCount, prop.l2CacheSize >> 20,
           prop.sharedMemPerMultiprocessor >> 10);
    const int sms = prop.multiProcessorCount;

    // 1. copy
    {
        size_t n = (size_t)1 << 26;   // 1 GiB of double2
        double2 *a, *b;
        CK(cudaMalloc(&a, n * 16));
        CK(cudaMalloc(&b, n * 16));
        CK(cudaMemset(a, 0, n * 16));
        float ms = time_ms([&] { copy_kernel<<<sms * 16, 512>>>(a, b, n); });
        printf("copy 1GiB->1GiB: %.3f ms  %.0f GB/s (read+write)\n", ms, 2.0 * n * 16 / ms / 1e6);
        ms = time_ms([&] { cudaMemsetAsync(b, 0, n * 16); });
        printf("memset 1GiB: %.3f ms  %.0f GB/s\n", ms, 1.0 * n * 16 / ms / 1e6);
        cudaFree(a);
        cudaFree(b);
    }
    // 2. global red
    {
        size_t cells = (size_t)1 << 28;   // 2 GiB of doubles
        double *g;
        CK(cudaMalloc(&g, cells * 8));
        CK(cudaMemset(g, 0, cells * 8));
        size_t nops = (size_t)1 << 28;
        const size_t spans[] = {(size_t)1 << 20, (size_t)1 << 22, (size_t)1 << 23, (size_t)1 << 24, (size_t)1 << 27, (size_t)1 << 28};
        for (size_t span : spans) {
            float ms = time_ms([&] { red_kernel<double><<<sms * 16, 256>>>(g, span, nops, 0, 0); });
            printf("red.f64 random over %6zu MB: %8.3f ms  %7.1f Gred/s\n", span * 8 >> 20, ms, nops / ms / 1e6);
            ms = time_ms([&] { red_kernel<unsigned long long><<<sms * 16, 256>>>((unsigned long long *)g, span, nops, 0, 0); });
            printf("red.u64 random over %6zu MB: %8.3f ms  %7.1f Gred/s\n", span * 8 >> 20, ms, nops / ms / 1e6);
        }
        for (size_t span : {(size_t)1 << 22, (size_t)1 << 28}) {
            for (int bytes : {32, 64, 128, 256, 512, 2048}) {
                for (int coherent : {0, 1}) {
                    const int nops = 2048, ctas = sms * 4;
                    float ms = time_ms([&] { bulk_red_kernel<<<ctas, 256, 8 * bytes>>>(g, span, bytes, nops, coherent); });
                    double ops = (double)ctas * 8 * nops;
                    printf("TMA bulk red.f64 %4d B/op %s over %5zu MB: %8.3f ms  %7.2f Gop/s  %8.1f Gcell/s  %7.0f GB/s\n", bytes,
                           coherent ? "coherent" : "random  ", span * 8 >> 20, ms, ops / ms / 1e6, ops * (bytes / 8) / ms / 1e6,
                           ops * bytes / ms / 1e6);
                }
            }
        }
        for (int dims : {256, 512}) {
            size_t np = (size_t)dims * dims * dims;
            float ms = time_ms([&] { red_kernel<double><<<sms * 16, 256>>>(g, 0, np * 8, 1, dims); });
            printf("red.f64 CIC-coherent %d^3 (1/cell): %8.3f ms  %7.1f Gred/s  %.1f Gpart/s\n", dims, ms, np * 8 / ms / 1e6, np / ms / 1e6);
        }
        cudaFree(g);
    }
    // 3. shared-memory accumulation
    {
        double *out;
        CK(cudaMalloc(&out, 8));
        const char *names[] = {"f64 atomicAdd(CAS)", "u64 atomicAdd(CAS)", "2x u32 add+carry", "f32 atomicAdd(CAS)", "u32 atomicAdd"};
        const int nops = 4096;
        for (int tile : {4096, 512}) {
            for (int shift : {0, 2, 5}) {
                const int ctas = sms * 4;
                float ms[5];
                ms[0] = time_ms([&] { smem_kernel<0><<<ctas, 256, 4096 * 8>>>(out, tile, nops, shift); });
                ms[1] = time_ms([&] { smem_kernel<1><<<ctas, 256, 4096 * 8>>>(out, tile, nops, shift); });
                ms[2] = time_ms([&] { smem_kernel<2><<<ctas, 256, 4096 * 8>>>(out, tile, nops, shift); });
                ms[3] = time_ms([&] { smem_kernel<3><<<ctas, 256, 4096 * 8>>>(out, tile, nops, shift); });
                ms[4] = time_ms([&] { smem_kernel<4><<<ctas, 256, 4096 * 8>>>(out, tile, nops, shift); });
                for (int k = 0; k < 5; k++) {
                    double total = (double)ctas * 256 * nops;
                    printf("smem %-20s tile %4d cells, %2d lanes/cell: %8.3f ms  %7.1f Gadd/s  (%.2f add/clk/SM @1.9GHz)\n", names[k],
                           tile, 1 << shift, ms[k], total / ms[k] / 1e6, total / ms[k] / 1e6 / sms / 1.9);
                }
            }
        }
        int *iout;
        CK(cudaMalloc(&iout, 4));
        float ms = time_ms([&] { match_kernel<<<sms * 8, 256>>>(iout, 4096); });
        printf("match_any: %.3f ms  %.1f G warp-ops/s (%.2f per clk per SM)\n", ms, (double)sms * 8 * 8 * 4096 / ms / 1e6,
               (double)sms * 8 * 8 * 4096 / ms / 1e6 / sms / 1.9);
    }
    return 0;
}
