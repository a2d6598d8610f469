#!/bin/bash
# Round 2, session 8 (2 GPUs): ghost pull over NVLink, deferred rejection check, host-shard overlap
mkdir -p gpurun_out
nvidia-smi --query-gpu=name --format=csv,noheader | head -3
echo "== slab tests (emulated + NCCL)"; timeout 1200 python -m pytest tests/test_gpu_slab.py -m gpu -x -q --timeout 300 > gpurun_out/r2s8_pytest_slab.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/r2s8_pytest_slab.log
show() {
python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({"ms":round(d["ms_per_step"],3), "e2e":d.get("e2e") and round(d["e2e"]["ms_per_step"],2)}, {k:round(v,3) for k,v in d.get("stage_ms",{}).items()}, d["config"].get("parallelism"), d.get("self_check") and list(d["self_check"]))
except Exception as e: print("ERR", e)
PY
}
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline"
run() { name=$1; shift; echo "== $name: $*"; timeout 400 $T "$@" > gpurun_out/r2s8_$name.json 2> gpurun_out/r2s8_$name.err; echo "rc=$?"; show gpurun_out/r2s8_$name.json; tail -2 gpurun_out/r2s8_$name.err | cut -c1-300; }
run c3_n2_pull --check-mass
run c3_n2_sendrecv --no-ghost-pull --no-e2e
run c3_n2_pull_fixed --fixed-point --no-e2e
run c3_n2_march --no-sweep --no-e2e
