// Device-side pieces shared by the tile kernels (fused x pass, column pass, fused z+y pass): mbarrier and TMA
// wrappers, cp.async, and the two-half exchange through shared memory between the register passes of a Plan.
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "fftx_core.cuh"

namespace genpk {

// ---- TMA: a tile [N][C] of complex doubles is N rows of C*16 contiguous bytes at a fixed pitch -- a 3-D tensor
// box {C complex, rows, 1}.  One elected thread arms an mbarrier with the tile's byte count and issues
// N/256 bulk tensor copies (a box dimension holds at most 256); nobody computes an address, nobody waits on
// a copy it issued itself.  Columns past the end of a row are zero-filled by the copy engine.
constexpr int TMA_BOX_ROWS = 256;
constexpr int ZERO_BOX_ROWS = 32;   // rows of a zero store: the source is a C*16*32-byte block of zeros (4 KB at C = 8)

__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity)
{
    const unsigned addr = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("{\n\t.reg .pred p;\n\tWAIT:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE;\n\tbra WAIT;\n\tDONE:\n\t}"
                 ::"r"(addr), "r"(parity) : "memory");
}
// box at coordinates (c0 doubles along a row, c1, c2) of the 3-D tensor behind `map` -> shared memory
__device__ __forceinline__ void tma_load_3d(void *smem_dst, const CUtensorMap *map, int c0, int c1, int c2, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(map), "r"(c0), "r"(c1), "r"(c2),
                   "r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}

// shared memory box -> the box at (c0, c1, c2) of the tensor; completion is tracked by the bulk async-group
__device__ __forceinline__ void tma_store_3d(const CUtensorMap *map, int c0, int c1, int c2, const void *smem_src)
{
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];"
                 ::"l"(map), "r"(c0), "r"(c1), "r"(c2), "r"((unsigned)__cvta_generic_to_shared(smem_src)) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit_group() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

// The staged tile has been read into registers: wait for those shared-memory loads to
// complete (an empty asm that consumes the registers) so that the slots may be refilled.
__device__ __forceinline__ void loads_landed(const fftx::cd *v)
{
#pragma unroll
    for (int i = 0; i < fftx::EPT; i += 4)
        asm volatile("" ::"d"(v[i].x), "d"(v[i].y), "d"(v[i + 1].x), "d"(v[i + 1].y), "d"(v[i + 2].x), "d"(v[i + 2].y),
                     "d"(v[i + 3].x), "d"(v[i + 3].y)
                     : "memory");
}

// One exchange through the half-size buffer E ([N/2][C] complex): the lower half of the index
// space first, then the upper half.  WI(i) / RI(i): element index of register i on the writing /
// reading side.  Which half an index falls in is a compile-time property of i (the predicates
// fold away), except on the pass-2 side of a plan whose thread owns a single radix-R2 unit
// (2048): there all 16 indices of a thread lie in the same half (W_UNI / R_UNI).
struct BlockSync {
    __device__ __forceinline__ void operator()() const { __syncthreads(); }
};
// named barrier `id` over `count` threads (the two halves of a CTA that run out of step with one another)
struct NamedSync {
    int id, count;
    __device__ __forceinline__ void operator()() const { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
};

template <int N, int C, bool W_UNI, bool R_UNI, class WI, class RI, class SYNC = BlockSync>
__device__ __forceinline__ void exchange(fftx::cd *E, const fftx::cd *src, fftx::cd *dst, int c, WI wi, RI ri, SYNC sync = SYNC())
{
#pragma unroll
    for (int half = 0; half < 2; half++) {
        if (half)
            sync();                                    // the lower half has been read
        if (W_UNI) {
            if ((wi(0) >= N / 2) == (half == 1)) {
#pragma unroll
                for (int i = 0; i < fftx::EPT; i++)
                    E[(wi(i) - half * (N / 2)) * C + c] = src[i];
            }
        } else {
#pragma unroll
            for (int i = 0; i < fftx::EPT; i++) {
                const int idx = wi(i);
                if ((idx >= N / 2) == (half == 1))
                    E[(idx - half * (N / 2)) * C + c] = src[i];
            }
        }
        sync();
        if (R_UNI) {
            if ((ri(0) >= N / 2) == (half == 1)) {
#pragma unroll
                for (int i = 0; i < fftx::EPT; i++)
                    dst[i] = E[(ri(i) - half * (N / 2)) * C + c];
            }
        } else {
#pragma unroll
            for (int i = 0; i < fftx::EPT; i++) {
                const int idx = ri(i);
                if ((idx >= N / 2) == (half == 1))
                    dst[i] = E[(idx - half * (N / 2)) * C + c];
            }
        }
    }
}


// Tensor of complex doubles viewed as doubles: inner extent 2*cols (valid columns), `rows` rows `row_pitch` complex
// apart, `slabs` slabs `slab_pitch` complex apart; box = {2*C doubles, box_rows, box_slabs}.  (fftx_power.cu)
bool tma_available();
int make_tile_map(CUtensorMap *map, const void *base, long long cols, long long rows, long long row_pitch, long long slabs,
                  long long slab_pitch, int C, int box_rows, int box_slabs);
int ensure_twiddles(genpk_ctx *ctx);

}  // namespace genpk
