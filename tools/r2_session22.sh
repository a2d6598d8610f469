#!/bin/bash
# Round 2, session 22 (2 GPUs): NCCL / multi-handle tests with the fused (y,z) scatter kernel, bench at N = 1 and 2
mkdir -p gpurun_out
echo "== pytest gpu: slab, multi, cli, distributed"; timeout 1500 python -m pytest tests/test_gpu_slab.py tests/test_gpu_multi.py tests/test_gpu_cli.py tests/test_gpu_distributed.py -m gpu -q --timeout 900 > gpurun_out/r2s22_pytest.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/r2s22_pytest.log
show() {
python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ("n_gpus","value","ms_per_step","gpu_launches")}, {k:round(v,3) for k,v in d.get("stage_ms",{}).items()}, "e2e", d.get("e2e") and round(d["e2e"].get("ms_per_step",0),2), d.get("self_check") and list(d["self_check"]))
except Exception as e: print("ERR", e)
PY
}
echo "== N=2"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2s22_n2.json 2> gpurun_out/r2s22_n2.err; echo "rc=$?"; show gpurun_out/r2s22_n2.json; tail -2 gpurun_out/r2s22_n2.err
echo "== N=2 library z + scatter cols (fused-zy 0)"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --fused-zy 0 > gpurun_out/r2s22_n2_zy0.json 2> gpurun_out/r2s22_n2_zy0.err; echo "rc=$?"; show gpurun_out/r2s22_n2_zy0.json
echo "== N=1"
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2s22_n1.json 2> gpurun_out/r2s22_n1.err; echo "rc=$?"; show gpurun_out/r2s22_n1.json
