#!/bin/bash
# Session 26 (2 GPUs): C5 and C3 on two GPUs with 128-byte-aligned transposed rows.
# peer-store transpose end to end, mass conservation; C3 N=2 final numbers.
mkdir -p gpurun_out
show() {
python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k:d[k] for k in ("n_gpus","value","ms_per_step","gpu_launches")}, {k:round(v,3) for k,v in d["stage_ms"].items()}, d["config"]["parallelism"], "e2e", d["e2e"] and round(d["e2e"]["ms_per_step"],2))
except Exception as e: print("ERR", e)
PY
}
echo "== bench c5 N=2"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 2 --workload c5 --steps 3 --warmup 3 --no-e2e --check-mass > gpurun_out/s26_bench_c5_n2.json 2> gpurun_out/s26_bench_c5_n2.err
echo "rc=$?"; show gpurun_out/s26_bench_c5_n2.json; grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" gpurun_out/s26_bench_c5_n2.err | tail -6
echo "== bench c3 N=2"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 2 --steps 10 --warmup 3 --check-mass > gpurun_out/s26_bench_c3_n2.json 2> gpurun_out/s26_bench_c3_n2.err
echo "rc=$?"; show gpurun_out/s26_bench_c3_n2.json; grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" gpurun_out/s26_bench_c3_n2.err | tail -4
