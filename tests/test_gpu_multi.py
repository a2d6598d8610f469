"""N GPUs behind one C-ABI handle (genpk_multi_*, csrc/multi.cu): single process, particles in any order routed
on the device, ghost plane pulled from the neighbour's memory, transpose by peer stores (or peer copies),
partial sums added on the host.  With fewer GPUs than slabs the slabs share devices -- the code path is the
same, which is how the driver's one-GPU box tests it; on a multi-GPU box every slab gets its own GPU."""
import numpy as np
import pytest

import genpk_b200 as gp
from genpk_b200 import api

pytestmark = pytest.mark.gpu


def _devices(n):
    import torch
    have = torch.cuda.device_count()
    return [r % have for r in range(n)]


def _particles(n, box, seed, var_mass=False):
    rng = np.random.default_rng(seed)
    pos = ((rng.random((n, 3)) * 1.1 - 0.05) * box).astype(np.float32)
    masses = (10.0 ** rng.uniform(-1, 1, n)).astype(np.float32) if var_mass else None
    return pos, masses


@pytest.mark.parametrize("P,dims,var_mass", [(2, 64, False), (4, 64, True), (8, 96, True), (2, 256, True), (4, 256, False),
                                             (1, 48, True)])
def test_multi_pk_against_the_reference(ref, P, dims, var_mass):
    """Random particles (some outside the box), chunks in two deposit calls: P(k) against the reference's own
    object code; 64 / 96 take the pack + peer-copy transpose, 256 the peer-store y pass + fused x pass."""
    box, n = 300.0, 200000
    pos, masses = _particles(n, box, 100 + P + dims, var_mass)
    tm = float(masses.astype(np.float64).sum()) if var_mass else 0.5 * n
    _, pr, cr, kr = ref.pk(box, dims, pos, masses, 0.5, tm, dims)
    with api.MultiContext(dims, P, _devices(P)) as m:
        for _ in range(2):                                   # twice: buffers, events and blocks are reused
            m.grid_zero()
            cut = n // 3
            m.deposit(pos[:cut], masses[:cut] if var_mass else None, 0.5, box)
            m.deposit(pos[cut:], masses[cut:] if var_mass else None, 0.5, box)
            p, c, k = m.fft_power(dims, tm, tm)
    assert np.array_equal(c, cr)
    nz = cr > 0
    np.testing.assert_allclose(p[nz], pr[nz], rtol=1e-5, atol=0)
    np.testing.assert_allclose(k[nz], kr[nz], rtol=1e-5, atol=0)


@pytest.mark.parametrize("P", [2, 4])
def test_multi_fixed_point_equals_one_gpu_bit_for_bit(P):
    """Deterministic accumulation: the N-slab P(k) sums come from grids that equal the single-GPU grid bit for
    bit, so the spectra agree to the rounding of the partial-sum order only."""
    import torch
    n_side = dims = 128
    box = 1000.0
    n = n_side ** 3
    d = torch.empty(3 * n, dtype=torch.float32, device="cuda:0")
    api.synth_particles_dev(api.SYNTH_CLUSTERED, 42, n_side, 0, n, box, dims, d.data_ptr())
    torch.cuda.synchronize()
    pos = d.cpu().numpy()
    with gp.Context(dims, 0, api.FLAG_FIXED_POINT) as ctx:
        ctx.grid_zero()
        ctx.deposit(pos, None, 1.0, box)
        p1, c1, k1 = ctx.fft_power(dims, float(n), float(n))
        ctx.synchronize()
    with api.MultiContext(dims, P, _devices(P), api.FLAG_FIXED_POINT) as m:
        m.grid_zero()
        m.deposit(pos, None, 1.0, box)
        p, c, k = m.fft_power(dims, float(n), float(n))
    assert np.array_equal(c, c1)
    nz = c1 > 0
    np.testing.assert_allclose(p[nz], p1[nz], rtol=1e-9, atol=0)
    np.testing.assert_allclose(k[nz], k1[nz], rtol=1e-12, atol=0)


def test_multi_rejects_non_finite_positions_loudly():
    dims, box = 32, 10.0
    pos, _ = _particles(1000, box, 3)
    pos[17, 1] = np.nan
    with api.MultiContext(dims, 2, _devices(2)) as m:
        m.grid_zero()
        m.deposit(pos, None, 1.0, box)
        with pytest.raises(Exception, match="rejected"):
            m.fft_power(dims, 1000.0, 1000.0)
