#!/bin/bash
# Session 19 (2 GPUs): transpose fused into the y pass (peer stores over NVLink).
mkdir -p gpurun_out
echo "== pytest fftx (1 GPU, emulated scatter)"; timeout 600 python -m pytest tests/test_gpu_fftx.py -q -x -k "scatter or own_y" > gpurun_out/s19_pytest_fftx.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/s19_pytest_fftx.log
echo "== pytest slab (nccl, world=2)"; timeout 400 python -m pytest tests/test_gpu_slab.py -x -q -k nccl > gpurun_out/s19_pytest_slab.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/s19_pytest_slab.log
show() {
python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k:d[k] for k in ("n_gpus","value","ms_per_step","gpu_launches")}, {k:round(v,3) for k,v in d["stage_ms"].items()}, d["config"]["parallelism"], "e2e", d["e2e"] and round(d["e2e"]["ms_per_step"],2))
except Exception as e: print("ERR", e)
PY
}
echo "== bench c3 N=2 (peer stores)"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/s19_bench_c3_n2.json 2> gpurun_out/s19_bench_c3_n2.err
echo "rc=$?"; show gpurun_out/s19_bench_c3_n2.json; grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" gpurun_out/s19_bench_c3_n2.err | tail -5
echo "== bench c3 N=2 (pack + all-to-all)"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e --no-scatter > gpurun_out/s19_bench_c3_n2_a2a.json 2> gpurun_out/s19_bench_c3_n2_a2a.err
echo "rc=$?"; show gpurun_out/s19_bench_c3_n2_a2a.json; grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" gpurun_out/s19_bench_c3_n2_a2a.err | tail -3
