#!/bin/bash
# Session 23 (8 GPUs): C5 (2048^3 -> 2048^3) and C3 on 8 GPUs, NCCL peer-store test at world=8, C3 at N=4.
mkdir -p gpurun_out
nvidia-smi -L | wc -l
show() {
python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k:d[k] for k in ("n_gpus","value","ms_per_step","gpu_launches")}, {k:round(v,3) for k,v in d["stage_ms"].items()}, d["config"]["parallelism"], "e2e", d["e2e"] and round(d["e2e"]["ms_per_step"],2))
except Exception as e: print("ERR", e)
PY
}
echo "== bench c5 N=8"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --workload c5 --steps 5 --warmup 3 --check-mass > gpurun_out/s23_bench_c5_n8.json 2> gpurun_out/s23_bench_c5_n8.err
echo "rc=$?"; show gpurun_out/s23_bench_c5_n8.json; grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" gpurun_out/s23_bench_c5_n8.err | tail -5
echo "== bench c3 N=8"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 8 --steps 10 --warmup 3 --check-mass > gpurun_out/s23_bench_c3_n8.json 2> gpurun_out/s23_bench_c3_n8.err
echo "rc=$?"; show gpurun_out/s23_bench_c3_n8.json; grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" gpurun_out/s23_bench_c3_n8.err | tail -4
echo "== bench c3 N=4"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 4 --steps 10 --warmup 3 --no-e2e > gpurun_out/s23_bench_c3_n4.json 2> gpurun_out/s23_bench_c3_n4.err
echo "rc=$?"; show gpurun_out/s23_bench_c3_n4.json; grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" gpurun_out/s23_bench_c3_n4.err | tail -4
echo "== pytest slab (nccl, world=8)"; timeout 300 python -m pytest tests/test_gpu_slab.py -x -q -k nccl > gpurun_out/s23_pytest_slab.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/s23_pytest_slab.log
