"""A/B timing of the FFT + binning variants on one resident C3 particle set (measurement tool, not product).
usage: python tools/ab_fft.py [dims n_side]"""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import genpk_b200 as gp
from genpk_b200 import api

dims = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
n_side = int(sys.argv[2]) if len(sys.argv) > 2 else dims
box = 1000.0
torch.cuda.set_device(0)
n = n_side ** 3
dpos = torch.empty(3 * n, dtype=torch.float32, device="cuda")
api.synth_particles_dev(api.SYNTH_CLUSTERED, 42, n_side, 0, n, box, dims, dpos.data_ptr())
torch.cuda.synchronize()
ref = None
with gp.Context(dims, 0) as ctx:
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    for name, fused, own_y in [("unfused: cuFFT 3-D + bin_power_kernel", 0, 0), ("fused x pass, cuFFT 2-D (y,z)", 1, 0),
                               ("fused x pass, cuFFT z + own y pass", 1, 1), ("fused x pass 4096-mode tile, own y pass", 2, 1)]:
        ctx.set_option(api.OPT_FUSED_XPASS, fused)
        ctx.set_option(api.OPT_OWN_YPASS, own_y)
        def step():
            ctx.grid_zero()
            ctx.deposit_dev(dpos.data_ptr(), n, 0, 1.0, box)
            return ctx.fft_power(dims, float(n), float(n))
        for _ in range(3):
            out = step()
        ctx.stage_reset()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        K = 5
        for _ in range(K):
            out = step()
        e1.record(); torch.cuda.synchronize()
        st = {k: ctx.stage_total_ms(v)[0] / K for k, v in (("fft", api.STAGE_FFT), ("power", api.STAGE_POWER), ("deposit", api.STAGE_DEPOSIT))}
        if ref is None:
            ref = out
        nz = ref[1] > 0
        err = float(np.max(np.abs(out[0][nz] / ref[0][nz] - 1)))
        print(json.dumps({"variant": name, "step_ms": round(e0.elapsed_time(e1) / K, 3), **{k: round(v, 3) for k, v in st.items()},
                          "counts_equal": bool(np.array_equal(out[1], ref[1])), "max_rel_dP_vs_first": err}), flush=True)
