#!/bin/bash
# Round 2, session 11 (2 GPUs): the single-process multi-GPU C ABI, gen-pk --gpus, slab tests again
mkdir -p gpurun_out
echo "== multi + cli + slab tests"; timeout 1500 python -m pytest tests/test_gpu_multi.py tests/test_gpu_cli.py tests/test_gpu_slab.py tests/test_gpu_fftx.py -m gpu -q --timeout 300 > gpurun_out/r2s11_pytest.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/r2s11_pytest.log
echo "== multi C ABI timing, C3 host particles"
cat > /tmp/multi_bench.py <<'PY'
import sys, time, numpy as np, torch
sys.path.insert(0, ".")
from genpk_b200 import api
n_side = dims = 1024; n = n_side**3; box = 1000.0
for P in (1, 2):
    d = torch.empty(3*n, dtype=torch.float32, device="cuda:0")
    api.synth_particles_dev(api.SYNTH_CLUSTERED, 42, n_side, 0, n, box, dims, d.data_ptr()); torch.cuda.synchronize()
    h = torch.empty(3*n, dtype=torch.float32, pin_memory=True); h.copy_(d); torch.cuda.synchronize(); del d; torch.cuda.empty_cache()
    with api.MultiContext(dims, P) as m:
        for it in range(3):
            t0 = time.perf_counter()
            p, c, k = m.pk_from_particles_ptr(h.data_ptr(), n, 1.0, box, float(n), dims)
            t1 = time.perf_counter()
            print(f"genpk_multi_pk_from_particles P={P} it={it}: {1e3*(t1-t0):.1f} ms, counts {int(c.astype(np.int64).sum())}", flush=True)
    del h
PY
timeout 600 python /tmp/multi_bench.py 2>&1 | tail -8
