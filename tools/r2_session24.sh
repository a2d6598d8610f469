#!/bin/bash
# Round 2, session 24 (8 GPUs): the 1 -> 8 curve on C3, C5 on 8
mkdir -p gpurun_out
show() {
python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ("n_gpus","value","ms_per_step","gpu_launches")}, {k:round(v,3) for k,v in d.get("stage_ms",{}).items()}, "e2e", d.get("e2e") and round(d["e2e"].get("ms_per_step",0),2), d.get("self_check") and list(d["self_check"]))
except Exception as e: print("ERR", e)
PY
}
for n in 8 4 2; do
echo "== C3 N=$n"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n bench.py --gpus $n --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2s24_c3_n$n.json 2> gpurun_out/r2s24_c3_n$n.err; echo "rc=$?"; show gpurun_out/r2s24_c3_n$n.json
done
echo "== C3 N=1"
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2s24_c3_n1.json 2> gpurun_out/r2s24_c3_n1.err; echo "rc=$?"; show gpurun_out/r2s24_c3_n1.json
echo "== C5 N=8"
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --workload c5 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2s24_c5_n8.json 2> gpurun_out/r2s24_c5_n8.err; echo "rc=$?"; show gpurun_out/r2s24_c5_n8.json; tail -2 gpurun_out/r2s24_c5_n8.err
