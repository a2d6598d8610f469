#!/bin/bash
# Round 2, session 27 (8 GPUs): multi-GPU tests at world 8 visible GPUs, the curve with ranks spread over the GPUs
mkdir -p gpurun_out
echo "== pytest gpu: slab, multi, cli"; timeout 1200 python -m pytest tests/test_gpu_slab.py tests/test_gpu_multi.py tests/test_gpu_cli.py -m gpu -q --timeout 900 > gpurun_out/r2s27_pytest.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/r2s27_pytest.log
show() {
python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ("n_gpus","value","ms_per_step")}, {k:round(v,3) for k,v in d.get("stage_ms",{}).items()}, "e2e", d.get("e2e") and round(d["e2e"].get("ms_per_step",0),2), d["config"].get("gpus_used"), d.get("self_check") and list(d["self_check"]))
except Exception as e: print("ERR", e)
PY
}
for n in 8 4 2; do
echo "== C3 N=$n"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2956$n bench.py --gpus $n --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2s27_c3_n$n.json 2> gpurun_out/r2s27_c3_n$n.err; echo "rc=$?"; show gpurun_out/r2s27_c3_n$n.json
done
echo "== gen-pk --gpus 8 on the bundled snapshot (fixed point) vs 1 GPU"
ls tests/data 2>/dev/null | head -3
