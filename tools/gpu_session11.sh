#!/bin/bash
# Session 11 (8 GPUs): NCCL slab tests at world=8, C3 scaling curve N=2,4,8, C5 (2048^3) on 8 GPUs.
mkdir -p gpurun_out
nvidia-smi -L | head -8
echo "== pytest slab (nccl, world=8)"; timeout 300 python -m pytest tests/test_gpu_slab.py -x -q --timeout 180 -k nccl > gpurun_out/s11_pytest_slab.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/s11_pytest_slab.log
show() {
python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k:d[k] for k in ("n_gpus","value","ms_per_step","gpu_launches")}, {k:round(v,3) for k,v in d["stage_ms"].items()}, d["config"]["parallelism"], "e2e", d["e2e"] and round(d["e2e"]["ms_per_step"],2))
except Exception as e: print("ERR", e)
PY
}
for n in 8 4 2; do
echo "== bench c3 N=$n"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/s11_bench_c3_n$n.json 2> gpurun_out/s11_bench_c3_n$n.err
echo "rc=$?"; show gpurun_out/s11_bench_c3_n$n.json; grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" gpurun_out/s11_bench_c3_n$n.err | tail -3
done
echo "== bench c5 N=8"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29527 bench.py --gpus 8 --workload c5 --steps 5 --warmup 3 --no-e2e > gpurun_out/s11_bench_c5_n8.json 2> gpurun_out/s11_bench_c5_n8.err
echo "rc=$?"; show gpurun_out/s11_bench_c5_n8.json; grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" gpurun_out/s11_bench_c5_n8.err | tail -5
echo "== bench c5 N=8 fixed-point"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29528 bench.py --gpus 8 --workload c5 --steps 3 --warmup 3 --no-e2e --fixed-point > gpurun_out/s11_bench_c5_n8_fixed.json 2> gpurun_out/s11_bench_c5_n8_fixed.err
echo "rc=$?"; show gpurun_out/s11_bench_c5_n8_fixed.json
