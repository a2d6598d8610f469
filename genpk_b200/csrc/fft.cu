// 3-D real-to-complex transform -- replaces fftw_plan_dft_r2c_3d(...) + fftw_execute()
// (gen-pk.cpp:193,233): unnormalised, forward (exp(-2 pi i ...)), in place on the
// FFTW padded layout, which is also cuFFT's in-place D2Z layout.
//
// Single GPU: one cufftPlan3d D2Z.  Slab mode: batched 2-D D2Z over the local x
// planes, then (after the caller's transpose) strided 1-D Z2Z along x.
// All plans share one work area owned by the context.
#include "common.cuh"

namespace genpk {

static int grow_work(genpk_ctx *ctx, size_t bytes)
{
    if (bytes <= ctx->fft_work_bytes)
        return 0;
    if (ctx->fft_work) cudaFree(ctx->fft_work);
    ctx->fft_work = nullptr;
    ctx->fft_work_bytes = 0;
    GENPK_CUDA_OK(cudaMalloc(&ctx->fft_work, bytes));
    ctx->fft_work_bytes = bytes;
    return 0;
}

static int attach_work(genpk_ctx *ctx)
{
    if (ctx->have_plan3d) GENPK_CUFFT_OK(cufftSetWorkArea(ctx->plan3d, ctx->fft_work));
    if (ctx->have_plan_yz) GENPK_CUFFT_OK(cufftSetWorkArea(ctx->plan_yz, ctx->fft_work));
    if (ctx->have_plan_x) GENPK_CUFFT_OK(cufftSetWorkArea(ctx->plan_x, ctx->fft_work));
    if (ctx->have_plan_z) GENPK_CUFFT_OK(cufftSetWorkArea(ctx->plan_z, ctx->fft_work));
    return 0;
}

int fft_3d(genpk_ctx *ctx, int which)
{
    const int d = ctx->g.dims;
    if (!ctx->have_plan3d) {
        size_t ws = 0;
        GENPK_CUFFT_OK(cufftCreate(&ctx->plan3d));
        GENPK_CUFFT_OK(cufftSetAutoAllocation(ctx->plan3d, 0));
        GENPK_CUFFT_OK(cufftMakePlan3d(ctx->plan3d, d, d, d, CUFFT_D2Z, &ws));
        ctx->have_plan3d = true;
        if (int rc = grow_work(ctx, ws)) return rc;
        if (int rc = attach_work(ctx)) return rc;
    }
    GENPK_CUFFT_OK(cufftSetStream(ctx->plan3d, ctx->stream));
    GENPK_CUFFT_OK(cufftExecD2Z(ctx->plan3d, ctx->grid[which], reinterpret_cast<cufftDoubleComplex *>(ctx->grid[which])));
    ctx->lib_calls++;
    return 0;
}

// (y,z) transform of the local planes as cuFFT's batched 1-D r2c along z (contiguous rows, in
// place on the padded layout: fft_z_rows) followed by our own in-place column pass along y.
int fft_z_rows(genpk_ctx *ctx, int which)
{
    const SlabGeom &g = ctx->g;
    if (!ctx->have_plan_z) {
        size_t ws = 0;
        long long n[1] = {g.dims};
        long long inembed[1] = {g.fd};
        long long onembed[1] = {g.nc};
        GENPK_CUFFT_OK(cufftCreate(&ctx->plan_z));
        GENPK_CUFFT_OK(cufftSetAutoAllocation(ctx->plan_z, 0));
        GENPK_CUFFT_OK(cufftMakePlanMany64(ctx->plan_z, 1, n, inembed, 1, g.fd, onembed, 1, g.nc, CUFFT_D2Z,
                                           (long long)g.nx * g.dims, &ws));
        ctx->have_plan_z = true;
        if (int rc = grow_work(ctx, ws)) return rc;
        if (int rc = attach_work(ctx)) return rc;
    }
    GENPK_CUFFT_OK(cufftSetStream(ctx->plan_z, ctx->stream));
    double *owned = ctx->grid[which] + g.owned_offset();
    GENPK_CUFFT_OK(cufftExecD2Z(ctx->plan_z, owned, reinterpret_cast<cufftDoubleComplex *>(owned)));
    ctx->lib_calls++;
    return 0;
}

static int fft_yz_own(genpk_ctx *ctx, int which)
{
    if (int rc = fft_z_rows(ctx, which)) return rc;
    return fft_cols_y(ctx, ctx->grid[which] + ctx->g.owned_offset(), ctx->g.nx);
}

int fft_yz(genpk_ctx *ctx, int which)
{
    const SlabGeom &g = ctx->g;
    if (int rc = materialize_zero(ctx, which)) return rc;
    if (fft_zy_supported(ctx)) {
        // one persistent kernel, int64 fixed-point sums converted as the rows are read (fft_zy.cu)
        const bool fixed = ctx->grid_is_fixed[which];
        ctx->grid_is_fixed[which] = false;
        return fft_zy(ctx, ctx->grid[which] + g.owned_offset(), g.nx, false, fixed, ctx->grid_scale_bits[which]);
    }
    if (int rc = fixed_to_double(ctx, which)) return rc;
    if (fft_cols_supported(ctx))
        return fft_yz_own(ctx, which);
    // Planes per cuFFT call.  One call over the whole slab runs the z pass over every plane
    // and then the y pass over every plane: 4 trips through HBM.  Groups of a few planes
    // (tens of MB) keep the z pass's output in L2 for the y pass: 2 trips.
    int batch = g.nx;
    if (ctx->fft_yz_batch > 0 && ctx->fft_yz_batch < g.nx && g.nx % ctx->fft_yz_batch == 0)
        batch = ctx->fft_yz_batch;
    if (ctx->have_plan_yz && ctx->plan_yz_batch != batch) {
        cufftDestroy(ctx->plan_yz);
        ctx->have_plan_yz = false;
    }
    if (!ctx->have_plan_yz) {
        size_t ws = 0;
        long long n[2] = {g.dims, g.dims};
        long long inembed[2] = {g.dims, g.fd};
        long long onembed[2] = {g.dims, g.nc};
        GENPK_CUFFT_OK(cufftCreate(&ctx->plan_yz));
        GENPK_CUFFT_OK(cufftSetAutoAllocation(ctx->plan_yz, 0));
        GENPK_CUFFT_OK(cufftMakePlanMany64(ctx->plan_yz, 2, n, inembed, 1, (long long)g.dims * g.fd, onembed, 1,
                                           (long long)g.dims * g.nc, CUFFT_D2Z, batch, &ws));
        ctx->have_plan_yz = true;
        ctx->plan_yz_batch = batch;
        if (int rc = grow_work(ctx, ws)) return rc;
        if (int rc = attach_work(ctx)) return rc;
    }
    GENPK_CUFFT_OK(cufftSetStream(ctx->plan_yz, ctx->stream));
    double *owned = ctx->grid[which] + g.owned_offset();
    for (int x = 0; x < g.nx; x += batch) {
        double *p = owned + (size_t)x * g.plane();
        GENPK_CUFFT_OK(cufftExecD2Z(ctx->plan_yz, p, reinterpret_cast<cufftDoubleComplex *>(p)));
        ctx->lib_calls++;
    }
    return 0;
}

int fft_x(genpk_ctx *ctx, void *recv)
{
    const SlabGeom &g = ctx->g;
    const long long ny = g.dims / g.nranks;
    const long long inner = ny * g.nc;                 // stride between consecutive x
    if (!ctx->have_plan_x) {
        size_t ws = 0;
        long long n[1] = {g.dims};
        long long embed[1] = {g.dims};
        GENPK_CUFFT_OK(cufftCreate(&ctx->plan_x));
        GENPK_CUFFT_OK(cufftSetAutoAllocation(ctx->plan_x, 0));
        GENPK_CUFFT_OK(cufftMakePlanMany64(ctx->plan_x, 1, n, embed, inner, 1, embed, inner, 1, CUFFT_Z2Z, inner, &ws));
        ctx->have_plan_x = true;
        if (int rc = grow_work(ctx, ws)) return rc;
        if (int rc = attach_work(ctx)) return rc;
    }
    GENPK_CUFFT_OK(cufftSetStream(ctx->plan_x, ctx->stream));
    cufftDoubleComplex *p = reinterpret_cast<cufftDoubleComplex *>(recv);
    GENPK_CUFFT_OK(cufftExecZ2Z(ctx->plan_x, p, p, CUFFT_FORWARD));
    ctx->lib_calls++;
    return 0;
}

void fft_release(genpk_ctx *ctx)
{
    if (ctx->have_plan3d) cufftDestroy(ctx->plan3d);
    if (ctx->have_plan_yz) cufftDestroy(ctx->plan_yz);
    if (ctx->have_plan_x) cufftDestroy(ctx->plan_x);
    if (ctx->have_plan_z) cufftDestroy(ctx->plan_z);
    ctx->have_plan3d = ctx->have_plan_yz = ctx->have_plan_x = ctx->have_plan_z = false;
    if (ctx->fft_work) cudaFree(ctx->fft_work);
    ctx->fft_work = nullptr;
    ctx->fft_work_bytes = 0;
}

}  // namespace genpk
