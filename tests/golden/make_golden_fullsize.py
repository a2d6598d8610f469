#!/usr/bin/env python
"""Makes tests/golden/fullsize_pk.npz: P(k), mode counts and k_eff of BASELINE configs C2 and C3
at their FULL sizes, computed by the reference's own object code (oracle/_ref/libgenpk_ref.so:
fieldize.cpp + powerspectrum.c compiled unmodified; 3-D r2c = pocketfft stand-in for FFTW3).

Run on a GPU box (the particle sets are the device generator's, copied to the host so that the
fixture is pinned to the exact float32 positions the benchmark deposits):

    python tests/golden/make_golden_fullsize.py gpurun_out/fullsize_pk.npz

C3 needs about 40 GB of host memory and a few minutes of CPU time.  Only the particle
generation touches the GPU; nothing of libgenpk_cuda's deposit / FFT / binning is involved.
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main(out):
    import torch

    from genpk_b200 import api
    from oracle.oracle import Oracle, padded_shape, rfftn_padded
    ref = Oracle("reference")
    box = 1000.0
    res = {}
    for name, kind, n_side, dims in (("c2", api.SYNTH_UNIFORM_RANDOM, 256, 512), ("c3", api.SYNTH_CLUSTERED, 1024, 1024)):
        n = n_side ** 3
        dpos = torch.empty(3 * n, dtype=torch.float32, device="cuda")
        api.synth_particles_dev(kind, 42, n_side, 0, n, box, dims, dpos.data_ptr(), torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        pos = dpos.cpu().numpy()
        del dpos
        torch.cuda.empty_cache()
        t0 = time.perf_counter()
        field = np.zeros(padded_shape(dims), np.float64)
        ref.fieldize(box, dims, field, pos, None, 1.0, 1)
        t1 = time.perf_counter()
        res[f"{name}_pos_xor"] = np.bitwise_xor.reduce(pos.view(np.uint32))
        res[f"{name}_pos_sum"] = pos.astype(np.float64).sum()
        del pos
        res[f"{name}_grid_sum"] = field[:, :, :dims].sum()
        res[f"{name}_grid_sample"] = field[::max(1, dims // 8), ::max(1, dims // 8), :: max(1, dims // 8)].copy()
        spec = rfftn_padded(field, dims)
        t2 = time.perf_counter()
        del field
        rc, p, c, k = ref.powerspectrum(dims, spec, None, dims, float(n), float(n))
        t3 = time.perf_counter()
        del spec
        assert rc == 0 and int(c.astype(np.int64).sum()) == dims ** 3 - 1
        res[f"{name}_power"], res[f"{name}_count"], res[f"{name}_keffs"] = p, c, k
        res[f"{name}_cpu_seconds"] = np.array([t1 - t0, t2 - t1, t3 - t2])
        print(f"{name}: {n} particles -> {dims}^3, reference deposit {t1 - t0:.1f} s, fft {t2 - t1:.1f} s, "
              f"binning {t3 - t2:.1f} s on {len(os.sched_getaffinity(0))} threads", flush=True)
    np.savez_compressed(out, **res)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "fullsize_pk.npz"))
