#!/bin/bash
# Round 2, session 14: zero stores by whole rows (last reader of a mid row) vs tile by tile vs memset
mkdir -p gpurun_out
echo "== pytest fftx"; timeout 900 python -m pytest tests/test_gpu_fftx.py -m gpu -q -x --timeout 600 2>&1 | tail -5
for za in 1 2 0; do
timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --zero-after $za 2>/dev/null | tail -1 > gpurun_out/r2s14_zero_after_$za.json; python - <<PY
import json
d=json.loads(open("gpurun_out/r2s14_zero_after_$za.json").read()); print("zero-after $za", d["ms_per_step"], d["stage_ms"], d.get("self_check",{}).get("pk",{}).get("max_rel_power"))
PY
done
