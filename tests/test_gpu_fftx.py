"""Fused x-pass (fftx_power_kernel: last FFT pass + binning in one kernel) through the C ABI:
against the CPU oracle at 256^3, against the unfused cuFFT + bin_power path at the
BASELINE grid sides, and on transposed slab blocks (the multi-GPU layout) up to 2048."""
import numpy as np
import pytest

import genpk_b200 as gp
from genpk_b200 import api

pytestmark = pytest.mark.gpu


def _particles(n, box, seed, var_mass=False):
    rng = np.random.default_rng(seed)
    pos = ((rng.random((n, 3)) * 1.1 - 0.05) * box).astype(np.float32)
    masses = (10.0 ** rng.uniform(-1, 1, n)).astype(np.float32) if var_mass else None
    return pos, masses


def test_fused_vs_oracle_256(ref, port):
    """The fused path against the reference's own object code (oracle/_ref) and the C restatement."""
    dims, box, n = 256, 500.0, 300000
    pos, masses = _particles(n, box, 5, True)
    tm = float(masses.astype(np.float64).sum())
    _, pr, cr, kr = ref.pk(box, dims, pos, masses, 1.0, tm, dims)
    _, pp, cp, kp = port.pk(box, dims, pos, masses, 1.0, tm, dims)
    assert np.array_equal(cp, cr)
    np.testing.assert_allclose(pp[cr > 0], pr[cr > 0], rtol=1e-9, atol=0)
    with gp.Context(dims) as ctx:
        assert ctx.fused_xpass_supported(dims)
        launches0 = ctx.launch_count()
        ctx.grid_zero()
        ctx.deposit(pos, masses, 1.0, box)
        p, c, k = ctx.fft_power(dims, tm, tm)
        ctx.synchronize()
        assert ctx.launch_count() > launches0
    assert np.array_equal(c, cr), "mode counts differ from the oracle"
    nz = cr > 0
    np.testing.assert_allclose(p[nz], pr[nz], rtol=1e-5, atol=0)          # north_star tolerance for P(k)
    np.testing.assert_allclose(k[nz], kr[nz], rtol=1e-5, atol=0)


@pytest.mark.parametrize("dims,nrbins,tile", [(256, 256, 1), (256, 37, 1), (512, 512, 1), (1024, 1024, 1), (1024, 1024, 2),
                                              (1024, 1024, 3), (1024, 77, 3), (1024, 1024, 4), (1024, 77, 4)])
def test_fused_vs_unfused(dims, nrbins, tile):
    box, n = 1000.0, 400000
    pos, _ = _particles(n, box, dims)
    with gp.Context(dims) as ctx:
        ctx.set_option(api.OPT_FUSED_XPASS, tile)                    # 2: 4096-mode tiles at 1024
        ctx.grid_zero()
        ctx.deposit(pos, None, 1.0, box)
        ctx.fft()
        p0, c0, k0 = ctx.power(nrbins, float(n), float(n))
        ctx.grid_zero()
        ctx.deposit(pos, None, 1.0, box)
        p1, c1, k1 = ctx.fft_power(nrbins, float(n), float(n))
        # the switch really selects the library path, and that gives the same answer too
        ctx.set_option(api.OPT_FUSED_XPASS, 0)
        assert not ctx.fused_xpass_supported(nrbins)
        ctx.grid_zero()
        ctx.deposit(pos, None, 1.0, box)
        p2, c2, k2 = ctx.fft_power(nrbins, float(n), float(n))
        ctx.synchronize()
    assert int(c0.astype(np.int64).sum()) == dims ** 3 - 1
    assert np.array_equal(c1, c0) and np.array_equal(c2, c0)
    nz = c0 > 0
    np.testing.assert_allclose(p1[nz], p0[nz], rtol=1e-9, atol=0)
    np.testing.assert_array_equal(k1, k0)
    np.testing.assert_allclose(p2[nz], p0[nz], rtol=1e-10, atol=0)       # fp64 red.add order differs between deposits


def test_unsupported_sizes_fall_back():
    dims, box, n = 96, 100.0, 20000
    pos, _ = _particles(n, box, 3)
    with gp.Context(dims, flags=api.FLAG_FIXED_POINT) as ctx:         # fixed point: the two deposits are bit-identical
        assert not ctx.fused_xpass_supported(dims)
        ctx.grid_zero()
        ctx.deposit(pos, None, 1.0, box)
        ctx.fft()
        p0, c0, k0 = ctx.power(dims, float(n), float(n))
        ctx.grid_zero()
        ctx.deposit(pos, None, 1.0, box)
        p1, c1, k1 = ctx.fft_power(dims, float(n), float(n))
    assert np.array_equal(c1, c0)
    np.testing.assert_allclose(p1, p0, rtol=1e-12, atol=0)           # same kernels; the bin sums are atomics


@pytest.mark.parametrize("dims,P,rank", [(256, 2, 1), (512, 4, 3), (1024, 8, 5), (2048, 8, 2), (2048, 8, 0)])
def test_fused_slab_block(dims, P, rank):
    """A transposed block [dims][dims/P][nc] of random data: slab_fft_x + slab_power_partial
    against the fused kernel (which must leave the block untouched)."""
    import torch
    dev = torch.device("cuda", 0)
    nrbins = dims
    ctx = api.Context(dims, 0, 0, P, rank)
    try:
        ctx.set_stream(torch.cuda.current_stream(dev).cuda_stream)
        nd = ctx.slab_spectrum_bytes() // 8
        gen = torch.Generator(device=dev).manual_seed(dims + rank)
        spec = torch.randn(nd, dtype=torch.float64, device=dev, generator=gen)
        spec *= torch.exp(4.0 * torch.randn(nd, dtype=torch.float64, device=dev, generator=gen))
        keep = spec.clone()
        fused = torch.zeros(3 * nrbins, dtype=torch.float64, device=dev)
        ctx.slab_fftx_power_partial(spec.data_ptr(), nrbins, fused.data_ptr())
        torch.cuda.synchronize()
        assert torch.equal(spec, keep), "the fused pass must not write the block"
        del keep
        plain = torch.zeros(3 * nrbins, dtype=torch.float64, device=dev)
        ctx.slab_fft_x(spec.data_ptr())
        ctx.slab_power_partial(spec.data_ptr(), 0, nrbins, plain.data_ptr())
        torch.cuda.synchronize()
        f, q = fused.cpu().numpy().reshape(3, nrbins), plain.cpu().numpy().reshape(3, nrbins)
        np.testing.assert_array_equal(f[1:], q[1:])                  # sum |k| and counts: the same cached geometry pass
        nz = q[2] > 0
        np.testing.assert_allclose(f[0][nz], q[0][nz], rtol=1e-9, atol=0)
        assert np.all(f[0][~nz] == 0)
    finally:
        ctx.close()


@pytest.mark.parametrize("dims,P", [(256, 1), (512, 1), (512, 4), (1024, 1), (2048, 16)])
def test_own_y_pass_vs_cufft_2d(dims, P):
    """genpk_slab_fft_yz: cuFFT 1-D r2c along z + fft_cols_kernel along y (default) against
    cuFFT's batched 2-D plan, on the same random planes (device-resident comparison)."""
    import torch
    from genpk_b200.distributed import _DevMem
    dev = torch.device("cuda", 0)
    outs = []
    for own in (1, 0):
        ctx = api.Context(dims, 0, 0, P, P - 1)
        try:
            ctx.set_stream(torch.cuda.current_stream(dev).cuda_stream)
            ctx.set_option(api.OPT_OWN_YPASS, own)
            nd = ctx.grid_doubles()
            grid = torch.as_tensor(_DevMem(ctx.grid_ptr(), nd * 8), device=dev)
            gen = torch.Generator(device=dev).manual_seed(dims)
            grid.copy_(torch.randn(nd, dtype=torch.float64, device=dev, generator=gen))
            launches0 = ctx.launch_count()
            ctx.slab_fft_yz()
            torch.cuda.synchronize()
            assert (ctx.launch_count() > launches0) == bool(own)
            off = ctx.owned_offset()
            owned = (dims // P) * dims * 2 * (dims // 2 + 1)
            outs.append(grid[off:off + owned].clone())
        finally:
            ctx.close()
    a, b = outs
    scale = float(b.abs().max())
    assert float((a - b).abs().max()) <= 1e-12 * scale


@pytest.mark.parametrize("dims,P", [(256, 2), (256, 8), (512, 4), (1024, 2)])
def test_scatter_y_pass_emulated_on_one_gpu(dims, P):
    """genpk_slab_fft_yz_scatter: the y pass stores its results straight into the owner ranks'
    transposed blocks (peer pointers; here P slab contexts in one process on one GPU), then the
    fused x pass bins each block.  Against fft_yz + pack + hand-made all-to-all + cuFFT x pass +
    bin_power_kernel on the same random slabs."""
    import torch
    from genpk_b200.distributed import CudaStages, _DevMem
    dev = torch.device("cuda", 0)
    nrbins = dims
    st = [CudaStages(dims, P, r, dev) for r in range(P)]
    try:
        grids = []
        for r in range(P):
            nd = st[r].ctx.grid_doubles()
            g = torch.as_tensor(_DevMem(st[r].ctx.grid_ptr(), nd * 8), device=dev)
            gen = torch.Generator(device=dev).manual_seed(1000 * dims + r)
            g.copy_(torch.randn(nd, dtype=torch.float64, device=dev, generator=gen))
            grids.append((g, g.clone()))
        # (a) library path
        nx = ny = dims // P
        nc = dims // 2 + 1
        blk = nx * ny * nc * 2
        sends = []
        for r in range(P):
            st[r].ctx.set_option(api.OPT_OWN_YPASS, 0)
            st[r].fft_yz()
            sends.append(st[r].pack().clone())
        want = torch.zeros(3 * nrbins, dtype=torch.float64, device=dev)
        for s in range(P):
            spec = st[s].spectrum_buffer()
            for r in range(P):
                spec[r * blk:(r + 1) * blk] = sends[r][s * blk:(s + 1) * blk]
            st[s].fft_x(spec)
            want += st[s].power_partial(spec, None, nrbins)[:3 * nrbins]
        torch.cuda.synchronize()
        # (b) scatter path on the same input
        ptrs = [st[r].ctx.slab_recv_buffer()[0] for r in range(P)]
        for r in range(P):
            grids[r][0].copy_(grids[r][1])
            st[r].ctx.set_option(api.OPT_OWN_YPASS, 1)
            assert st[r].ctx.slab_scatter_supported()
            for s in range(P):
                st[r].ctx.slab_set_peer(s, None, ptrs[s])
        for r in range(P):
            st[r].fft_yz_scatter()
        torch.cuda.synchronize()                                   # "barrier": every rank has stored
        got = torch.zeros(3 * nrbins, dtype=torch.float64, device=dev)
        for s in range(P):
            got += st[s].fftx_power_partial(st[s].recv_block(), nrbins)[:3 * nrbins]
        torch.cuda.synchronize()
        w, g = want.cpu().numpy().reshape(3, nrbins), got.cpu().numpy().reshape(3, nrbins)
        np.testing.assert_array_equal(g[1:], w[1:])
        assert int(w[2].sum()) == dims ** 3 - 1
        nz = w[2] > 0
        np.testing.assert_allclose(g[0][nz], w[0][nz], rtol=1e-9, atol=0)
    finally:
        for s in st:
            s.close()


@pytest.mark.parametrize("dims", [256, 512, 1024])
def test_tma_tile_fills_give_the_same_bits_as_cp_async(dims):
    """GENPK_OPT_TMA: the column kernels fill their tiles with bulk tensor copies (one thread, an mbarrier)
    or with per-thread cp.async; the arithmetic is the same, so P(k) must be bit-identical (fixed-point
    deposit: identical grids), for the y pass + fused x pass and for the last, partly empty column group."""
    box, n = 1000.0, 300000
    pos, _ = _particles(n, box, 17 + dims)
    out = {}
    for tma in (1, 0):
        with gp.Context(dims, flags=api.FLAG_FIXED_POINT) as ctx:
            ctx.set_option(api.OPT_TMA, tma)
            ctx.grid_zero()
            ctx.deposit(pos, None, 1.0, box)
            out[tma] = ctx.fft_power(dims, float(n), float(n))
            ctx.synchronize()
    assert np.array_equal(out[1][1], out[0][1])
    assert int(out[1][1].astype(np.int64).sum()) == dims ** 3 - 1
    np.testing.assert_allclose(out[1][0], out[0][0], rtol=1e-12, atol=0)      # bin sums are shared-memory atomics: order may differ
    np.testing.assert_allclose(out[1][2], out[0][2], rtol=1e-12, atol=0)      # (the geometry sums are atomics too)


@pytest.mark.parametrize("dims,two_ctx", [(256, False), (256, True), (96, False), (512, False)])
def test_cross_spectrum_on_the_fused_path(ref, dims, two_ctx):
    """genpk_fft_power_cross: powerspectrum(f1, f2) of gen-pk.cpp:295-297 / 345-348 from two deposited grids.  On the
    fused grid sides the grids are replaced by their sum and difference and go through the fused auto path
    (re1*re2 + im1*im2 = (|F1+F2|^2 - |F1-F2|^2)/4); 96 takes the library route.  Against the reference's own
    powerspectrum() on pocketfft transforms of the reference's own fieldize() grids, correlated and anticorrelated
    fields, two fields in one context and two contexts on one device."""
    from oracle.oracle import padded_shape, rfftn_padded
    box, n = 700.0, 250000
    rng = np.random.default_rng(dims)
    pos1 = (rng.random((n, 3)) * box).astype(np.float32)
    # the second field: the first one displaced a little, plus unrelated particles
    pos2 = np.concatenate([np.mod(pos1[: n // 2] + rng.normal(0, 0.3 * box / dims, (n // 2, 3)), box),
                           rng.random((n // 3, 3)) * box]).astype(np.float32)
    m2 = (10.0 ** rng.uniform(-1, 0.5, len(pos2))).astype(np.float32)
    tm1, tm2 = float(n), float(m2.astype(np.float64).sum())
    f1 = np.zeros(padded_shape(dims))
    f2 = np.zeros(padded_shape(dims))
    ref.fieldize(box, dims, f1, pos1, None, 1.0, 1)
    ref.fieldize(box, dims, f2, pos2, m2, 0.0, 1)
    rc, pr, cr, kr = ref.powerspectrum(dims, rfftn_padded(f1, dims), rfftn_padded(f2, dims), dims, tm1, tm2)
    assert rc == 0
    if two_ctx:
        with gp.Context(dims) as c1, gp.Context(dims) as c2:
            c1.grid_zero()
            c2.grid_zero()
            c1.deposit(pos1, None, 1.0, box)
            c2.deposit(pos2, m2, 0.0, box)
            p, c, k = c1.fft_power_cross(dims, tm1, tm2, 0, 0, other=c2)
            c1.synchronize()
            c2.synchronize()
    else:
        with gp.Context(dims, flags=api.FLAG_TWO_FIELDS) as ctx:
            ctx.grid_zero(0)
            ctx.grid_zero(1)
            ctx.deposit(pos1, None, 1.0, box, 0)
            ctx.deposit(pos2, m2, 0.0, box, 1)
            p, c, k = ctx.fft_power_cross(dims, tm1, tm2, 0, 1)
            # and the library route gives the same spectrum
            ctx.grid_zero(0)
            ctx.grid_zero(1)
            ctx.deposit(pos1, None, 1.0, box, 0)
            ctx.deposit(pos2, m2, 0.0, box, 1)
            ctx.fft(0)
            ctx.fft(1)
            p0, c0, k0 = ctx.power(dims, tm1, tm2, 0, 1)
            ctx.synchronize()
        assert np.array_equal(c0, c)
        scale = np.abs(p0).max()
        np.testing.assert_allclose(p, p0, rtol=1e-7, atol=1e-12 * scale)
    assert np.array_equal(c, cr)
    nz = cr > 0
    # relative 1e-5 where the cross power is not a cancellation of the two auto powers (the bins with few modes at
    # low k can be: there the absolute floor applies)
    np.testing.assert_allclose(p[nz], pr[nz], rtol=1e-5, atol=1e-9 * np.abs(pr[nz]).max())
    np.testing.assert_allclose(k[nz], kr[nz], rtol=1e-5, atol=0)


@pytest.mark.parametrize("dims,fixed", [(256, False), (512, False), (1024, False), (2048, False), (256, True)])
def test_fused_x_pass_leaves_the_grid_zero(dims, fixed):
    """GENPK_OPT_ZERO_AFTER_POWER: the fused x pass stores zeros behind every tile it reads (TMA bulk stores), the
    genpk_grid_zero that follows is free, and a second P(k) of the same particles is the first one bit for bit."""
    box, n = 300.0, 200000
    pos, masses = _particles(n, box, dims + 7, True)
    tm = float(masses.astype(np.float64).sum())
    with gp.Context(dims, flags=api.FLAG_FIXED_POINT if fixed else 0) as ctx:
        ctx.set_option(api.OPT_DEPOSIT, api.DEPOSIT_SORTED)          # (an order-independent deposit: the two passes can be compared bit for bit)
        ctx.set_option(api.OPT_ZERO_AFTER_POWER, 1)
        ctx.grid_zero()
        ctx.deposit(pos, masses, 1.0, box)
        p1, c1, k1 = ctx.fft_power(dims, tm, tm)
        ctx.stage_reset()
        ctx.grid_zero()
        ctx.deposit(pos, masses, 1.0, box)
        assert ctx.stage_total_ms(api.STAGE_ZERO)[1] == 0, "a memset ran although the x pass had cleared the grid"
        p2, c2, k2 = ctx.fft_power(dims, tm, tm)
        ctx.grid_zero()
        if dims <= 1024:                                             # (a 2048^3 grid is 69 GB of host memory)
            left = ctx.grid_download()
            assert not left.any(), "the x pass left something behind"
            del left
        # switched off: the memset is back and nothing else changes
        ctx.set_option(api.OPT_ZERO_AFTER_POWER, 0)
        ctx.deposit(pos, masses, 1.0, box)
        p3, c3, k3 = ctx.fft_power(dims, tm, tm)
        ctx.stage_reset()
        ctx.grid_zero()
        ctx.deposit(pos, masses, 1.0, box)
        assert ctx.stage_total_ms(api.STAGE_ZERO)[1] == 1
        p4, c4, k4 = ctx.fft_power(dims, tm, tm)
        ctx.synchronize()
    for p, c, k in ((p2, c2, k2), (p3, c3, k3), (p4, c4, k4)):
        assert np.array_equal(c, c1)
        # (the deposit is bit-reproducible in fixed point, the per-bin sums are fp64 atomics either way: order of additions)
        np.testing.assert_allclose(p[c1 > 0], p1[c1 > 0], rtol=1e-12 if fixed else 1e-10, atol=0)


@pytest.mark.parametrize("dims,P,how,lag", [(256, 1, 1, 2), (512, 1, 1, 2), (512, 4, 1, 1), (1024, 1, 1, 2), (1024, 1, 2, 2), (1024, 4, 1, 5),
                                            (2048, 16, 1, 2), (256, 8, 1, 3)])
def test_fused_zy_kernel_vs_the_two_kernel_route(dims, P, how, lag):
    """fft_zy_kernel (z rows and y columns of a plane in one persistent kernel, GENPK_OPT_FUSED_ZY) against the batched
    1-D library transform along z + fft_cols_kernel along y, on the same random planes; slab contexts included."""
    import torch
    from genpk_b200.distributed import _DevMem
    dev = torch.device("cuda", 0)
    outs = []
    for fused in (how, 0):
        ctx = api.Context(dims, 0, 0, P, P - 1)
        try:
            ctx.set_stream(torch.cuda.current_stream(dev).cuda_stream)
            ctx.set_option(api.OPT_FUSED_ZY, fused)
            ctx.set_option(api.OPT_ZY_LAG, lag)
            nd = ctx.grid_doubles()
            grid = torch.as_tensor(_DevMem(ctx.grid_ptr(), nd * 8), device=dev)
            gen = torch.Generator(device=dev).manual_seed(dims + P)
            grid.copy_(torch.randn(nd, dtype=torch.float64, device=dev, generator=gen) *
                       torch.exp(4 * torch.randn(nd, dtype=torch.float64, device=dev, generator=gen)))
            ctx.slab_fft_yz()
            torch.cuda.synchronize()
            off = ctx.owned_offset()
            owned = (dims // P) * dims * 2 * (dims // 2 + 1)
            outs.append(grid[off:off + owned].clone())
            if P > 1:                                            # the ghost planes are not part of the transform
                ghosts = torch.cat([grid[:off], grid[off + owned:nd]])
                outs.append(ghosts.clone())
        finally:
            ctx.close()
    if P > 1:
        a, ga, b, gb = outs
        assert torch.equal(ga, gb)
    else:
        a, b = outs
    # per plane: every plane is an independent 2-D transform with its own dynamic range
    plane = dims * 2 * (dims // 2 + 1)
    a, b = a.view(-1, plane), b.view(-1, plane)
    scale = b.abs().amax(dim=1, keepdim=True)
    assert bool(((a - b).abs() <= 2e-13 * scale).all())


@pytest.mark.parametrize("dims", [256, 1024])
def test_fused_zy_reads_fixed_point_rows(dims):
    """A grid deposited in int64 fixed point goes into fft_zy_kernel as it is (converted on the way into the
    registers); P(k) agrees within rounding with converting first and taking the two-kernel route."""
    box, n = 300.0, 200000
    pos, masses = _particles(n, box, dims + 11, True)
    tm = float(masses.astype(np.float64).sum())
    res = []
    for fused in (1, 0):
        with gp.Context(dims, flags=api.FLAG_FIXED_POINT) as ctx:
            ctx.set_option(api.OPT_FUSED_ZY, fused)
            ctx.grid_zero()
            ctx.deposit(pos, masses, 1.0, box)
            res.append(ctx.fft_power(dims, tm, tm))
            ctx.synchronize()
    (p1, c1, k1), (p0, c0, k0) = res
    assert np.array_equal(c1, c0)
    np.testing.assert_allclose(p1[c0 > 0], p0[c0 > 0], rtol=1e-11, atol=0)
