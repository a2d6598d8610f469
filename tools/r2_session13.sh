#!/bin/bash
# Round 2, session 13: zero stores behind the fused x pass (UTMASTG) -- tests and A/B bench
mkdir -p gpurun_out
echo "== pytest fftx + parity"; timeout 900 python -m pytest tests/test_gpu_fftx.py tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -x --timeout 600 2>&1 | tail -5
echo "== bench zero-after"; timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e 2>/dev/null | tail -1 > gpurun_out/r2s13_zero_after.json; python - <<'PY'
import json
for f in ("r2s13_zero_after",):
    d=json.loads(open(f"gpurun_out/{f}.json").read()); print(f, d["ms_per_step"], d["stage_ms"], d.get("self_check"))
PY
echo "== bench no-zero-after"; timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-zero-after 2>/dev/null | tail -1 > gpurun_out/r2s13_no_zero_after.json; python - <<'PY'
import json
for f in ("r2s13_no_zero_after",):
    d=json.loads(open(f"gpurun_out/{f}.json").read()); print(f, d["ms_per_step"], d["stage_ms"])
PY
