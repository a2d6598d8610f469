#!/bin/bash
# 2-GPU: wide-ghost slab tests, N=2 bench
mkdir -p gpurun_out
echo "== pytest slab"; timeout 600 python -m pytest tests/test_gpu_slab.py tests/test_gpu_march.py -x -q --timeout 180 > gpurun_out/pytest_slab.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/pytest_slab.log
for w in c3 c2; do
echo "== bench N=2 $w"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --workload $w --steps 5 --warmup 3 --no-e2e > gpurun_out/bench_n2_$w.json 2> gpurun_out/bench_n2_$w.err
echo "rc=$?"; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_n2_$w.json").read().strip().splitlines()[-1])
    print({k:d[k] for k in ("value","ms_per_step","stage_ms","gpu_launches")}, d["config"]["parallelism"], d["config"].get("order_probe"))
except Exception as e: print("ERR", e)
PY
grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" gpurun_out/bench_n2_$w.err | tail -5
done
