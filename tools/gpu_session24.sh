#!/bin/bash
# Session 24 (2 GPUs): transposed blocks with 128-byte-aligned rows: emulated scatter tests, NCCL test, C3 N=2.
mkdir -p gpurun_out
echo "== pytest scatter"; timeout 300 python -m pytest tests/test_gpu_fftx.py -q -x -k "scatter or slab_block" > gpurun_out/s24_pytest.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/s24_pytest.log
echo "== pytest nccl"; timeout 300 python -m pytest tests/test_gpu_slab.py -q -x -k "peer_stores" > gpurun_out/s24_pytest_nccl.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/s24_pytest_nccl.log
echo "== bench c3 N=2"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e --check-mass > gpurun_out/s24_bench_c3_n2.json 2> gpurun_out/s24_bench_c3_n2.err
echo "rc=$?"; python - <<'PY'
import json
d=json.loads(open("gpurun_out/s24_bench_c3_n2.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("n_gpus","value","ms_per_step")}, {k:round(v,3) for k,v in d["stage_ms"].items()})
PY
grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" gpurun_out/s24_bench_c3_n2.err | tail -3
