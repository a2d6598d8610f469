// Kernels either side of the multi-GPU exchange steps: the ghost
// plane produced by the deposit on the high-x side of a slab is added into the
// neighbour's first plane, and the 2-D transformed slab is regrouped into one
// contiguous block per destination rank for the all-to-all transpose.
#include "common.cuh"

namespace genpk {

__global__ void plane_add_f64_kernel(double *dst, const double *src, size_t n)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        dst[i] += src[i];
}

__global__ void plane_add_i64_kernel(long long *dst, const long long *src, size_t n)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        dst[i] += src[i];
}

// side 0: planes received from rank-1 (its high ghosts) are added into my first owned planes;
// side 1: planes received from rank+1 (its low ghosts) into my last owned planes.
int ghost_accumulate(genpk_ctx *ctx, int which, int side, const void *recv)
{
    const SlabGeom &g = ctx->g;
    const int planes = side ? g.ghost_lo : g.ghost_hi;       // what the neighbour on that side sends
    if (planes == 0)
        return 0;
    const size_t n = g.plane() * (size_t)planes;
    double *dst = ctx->grid[which] + g.owned_offset() + (side ? g.plane() * (size_t)(g.nx - planes) : 0);
    const int blocks = ctx->sm_count * 8;
    if (ctx->grid_is_fixed[which])
        plane_add_i64_kernel<<<blocks, 256, 0, ctx->stream>>>(reinterpret_cast<long long *>(dst),
                                                              reinterpret_cast<const long long *>(recv), n);
    else
        plane_add_f64_kernel<<<blocks, 256, 0, ctx->stream>>>(dst, reinterpret_cast<const double *>(recv), n);
    ctx->launches++;
    GENPK_CUDA_OK(cudaGetLastError());
    return 0;
}

__global__ void touched_set_kernel(int *touched, int lo, int hi)
{
    touched[0] = lo;
    touched[1] = hi;
}

int touched_set(genpk_ctx *ctx, int which, int lo, int hi)
{
    touched_set_kernel<<<1, 1, 0, ctx->stream>>>(touched_ptr(ctx, which), lo, hi);
    ctx->launches++;
    GENPK_CUDA_OK(cudaGetLastError());
    return 0;
}

// Ghost exchange without a collective: this rank adds, into its outermost owned planes, the ghost
// planes of its ring neighbours -- read straight from their memory over NVLink (mapped with CUDA
// IPC), and only the planes their deposits touched (the {lowest, highest} written plane sits behind
// each grid allocation).  blockIdx.y: ghost plane slot (first the ghost_hi planes above rank-1's
// slab, then the ghost_lo planes below rank+1's); blocks of untouched planes leave at once.
struct PullArgs {
    void *mine;
    const void *prev, *next;      // the neighbours' grid allocations (next may be null: no low ghosts)
    size_t plane, grid_doubles;   // doubles per plane / per allocation (the touched range follows it)
    int nx, ghost_lo, ghost_hi;
    int fixed;
};

__global__ void __launch_bounds__(256) ghost_pull_kernel(PullArgs A)
{
    const int slot = blockIdx.y;
    const bool from_prev = slot < A.ghost_hi;
    const double *peer = reinterpret_cast<const double *>(from_prev ? A.prev : A.next);
    const int src_plane = from_prev ? A.ghost_lo + A.nx + slot : slot - A.ghost_hi;
    const int dst_plane = from_prev ? A.ghost_lo + slot : A.nx + (slot - A.ghost_hi);   // (= ghost_lo + nx - ghost_lo + j)
    const int *range = reinterpret_cast<const int *>(peer + A.grid_doubles);
    const int lo = range[0], hi = range[1];
    if (src_plane < lo || src_plane > hi)
        return;
    const size_t units = A.plane / 2;                                      // 16-byte units per plane
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    if (A.fixed) {
        const longlong2 *src = reinterpret_cast<const longlong2 *>(peer + (size_t)src_plane * A.plane);
        longlong2 *dst = reinterpret_cast<longlong2 *>(reinterpret_cast<double *>(A.mine) + (size_t)dst_plane * A.plane);
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < units; i += stride) {
            const longlong2 v = src[i];
            longlong2 d = dst[i];
            d.x += v.x;
            d.y += v.y;
            dst[i] = d;
        }
    } else {
        const double2 *src = reinterpret_cast<const double2 *>(peer + (size_t)src_plane * A.plane);
        double2 *dst = reinterpret_cast<double2 *>(reinterpret_cast<double *>(A.mine) + (size_t)dst_plane * A.plane);
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < units; i += stride) {
            const double2 v = src[i];
            double2 d = dst[i];
            d.x += v.x;
            d.y += v.y;
            dst[i] = d;
        }
    }
}

int ghost_pull(genpk_ctx *ctx, int which)
{
    const SlabGeom &g = ctx->g;
    PullArgs A;
    A.mine = ctx->grid[which];
    A.prev = ctx->grid_peer[which][0];
    A.next = ctx->grid_peer[which][1];
    A.plane = g.plane();
    A.grid_doubles = g.grid_doubles();
    A.nx = g.nx;
    A.ghost_lo = g.ghost_lo;
    A.ghost_hi = g.ghost_hi;
    A.fixed = ctx->grid_is_fixed[which] ? 1 : 0;
    const int slots = g.ghost_hi + g.ghost_lo;
    if (slots == 0)
        return 0;
    // enough CTAs per plane to keep NVLink busy when only a few planes are touched
    const unsigned per_plane = (unsigned)((ctx->sm_count * 4 + slots - 1) / slots);
    ghost_pull_kernel<<<dim3(per_plane < 8 ? 8 : per_plane, (unsigned)slots), 256, 0, ctx->stream>>>(A);
    ctx->launches++;
    GENPK_CUDA_OK(cudaGetLastError());
    return 0;
}

// in : [nx][dims][nc] complex   (after the batched 2-D D2Z)
// out: [nranks][nx][ny][nc]     block s goes to rank s, which owns y in [s*ny, (s+1)*ny)
__global__ void __launch_bounds__(256) slab_pack_kernel(const double2 *in, double2 *out, int nx, int dims, int nc, int ny)
{
    const long long rows = (long long)nx * dims;
    const int lane_rows = blockDim.x / 32;
    for (long long row = (long long)blockIdx.x * lane_rows + (threadIdx.x >> 5); row < rows;
         row += (long long)gridDim.x * lane_rows) {
        const int xl = (int)(row / dims), y = (int)(row % dims);
        const int s = y / ny, yl = y % ny;
        const double2 *src = in + (size_t)row * nc;
        double2 *dst = out + (((size_t)s * nx + xl) * ny + yl) * nc;
        for (int k = threadIdx.x & 31; k < nc; k += 32)
            dst[k] = src[k];
    }
}

int slab_pack(genpk_ctx *ctx, int which, void *send)
{
    const SlabGeom &g = ctx->g;
    const int ny = g.dims / g.nranks;
    slab_pack_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(reinterpret_cast<const double2 *>(ctx->grid[which] + g.owned_offset()),
                                                                reinterpret_cast<double2 *>(send), g.nx, g.dims, g.nc, ny);
    ctx->launches++;
    GENPK_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace genpk
