#!/bin/bash
# Round 2, session 23 (2 GPUs): NCCL / multi-handle tests; scatter with 8192-mode tiles in the fused (y,z) kernel
mkdir -p gpurun_out
echo "== pytest gpu: slab, multi, cli"; timeout 1500 python -m pytest tests/test_gpu_slab.py tests/test_gpu_multi.py tests/test_gpu_cli.py -m gpu -q --timeout 900 > gpurun_out/r2s23_pytest.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/r2s23_pytest.log
show() {
python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ("n_gpus","value","ms_per_step","gpu_launches")}, {k:round(v,3) for k,v in d.get("stage_ms",{}).items()}, "e2e", d.get("e2e") and round(d["e2e"].get("ms_per_step",0),2), d.get("self_check") and list(d["self_check"]))
except Exception as e: print("ERR", e)
PY
}
for zy in 2 1 0; do
echo "== N=2 fused-zy $zy"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$zy bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --fused-zy $zy > gpurun_out/r2s23_n2_zy$zy.json 2> gpurun_out/r2s23_n2_zy$zy.err; echo "rc=$?"; show gpurun_out/r2s23_n2_zy$zy.json
done
