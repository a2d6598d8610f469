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
