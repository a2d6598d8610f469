#!/bin/bash
# Session 18 (2 GPUs): NCCL slab test + C3 at N=2 with the fused x pass and own y pass; exchange-step timings.
mkdir -p gpurun_out
nvidia-smi -L | head -4
echo "== pytest slab (nccl, world=2)"; timeout 300 python -m pytest tests/test_gpu_slab.py -x -q -k nccl > gpurun_out/s18_pytest_slab.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/s18_pytest_slab.log
show() {
python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k:d[k] for k in ("n_gpus","value","ms_per_step","gpu_launches")}, {k:round(v,3) for k,v in d["stage_ms"].items()}, d["config"]["parallelism"], "e2e", d["e2e"] and round(d["e2e"]["ms_per_step"],2))
except Exception as e: print("ERR", e)
PY
}
echo "== bench c3 N=2"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/s18_bench_c3_n2.json 2> gpurun_out/s18_bench_c3_n2.err
echo "rc=$?"; show gpurun_out/s18_bench_c3_n2.json; grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" gpurun_out/s18_bench_c3_n2.err | tail -3
echo "== bench c3 N=2, library FFT passes"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e --no-fused-xpass --no-own-ypass > gpurun_out/s18_bench_c3_n2_lib.json 2> gpurun_out/s18_bench_c3_n2_lib.err
echo "rc=$?"; show gpurun_out/s18_bench_c3_n2_lib.json; grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" gpurun_out/s18_bench_c3_n2_lib.err | tail -3
