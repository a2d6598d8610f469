#!/bin/bash
# Round 2, session 25 (8 GPUs): e2e leg with every rank bound to its GPU's NUMA node
mkdir -p gpurun_out
show() {
python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ("n_gpus","value","ms_per_step")}, "e2e", d.get("e2e") and (round(d["e2e"].get("ms_per_step",0),2), d["e2e"].get("host_affinity")), d.get("self_check") and list(d["self_check"]))
except Exception as e: print("ERR", e)
PY
}
lscpu | grep -i -E "numa|socket|^CPU\(s\)" | head -8
nvidia-smi topo -m 2>/dev/null | head -14
for n in 8 4; do
echo "== C3 N=$n"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2954$n bench.py --gpus $n --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2s25_c3_n$n.json 2> gpurun_out/r2s25_c3_n$n.err; echo "rc=$?"; show gpurun_out/r2s25_c3_n$n.json
done
