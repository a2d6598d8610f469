#!/bin/bash
# Round 2, session 10: TMA tile fills in the column kernels
mkdir -p gpurun_out
echo "== fftx tests"; timeout 1200 python -m pytest tests/test_gpu_fftx.py tests/test_gpu_sweep.py tests/test_gpu_dropin.py tests/test_gpu_cli.py tests/test_gpu_parity.py -m gpu -q --timeout 300 > gpurun_out/r2s10_pytest.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/r2s10_pytest.log
show() {
python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k:round(v,3) for k,v in d.get("stage_ms",{}).items() if k in ("deposit","fft","binning")}, round(d["ms_per_step"],3), d.get("self_check") and list(d["self_check"]))
except Exception as e: print("ERR", e)
PY
}
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e"
run() { name=$1; shift; echo "== $name: $*"; timeout 200 $B "$@" > gpurun_out/r2s10_$name.json 2> gpurun_out/r2s10_$name.err; echo "rc=$?"; show gpurun_out/r2s10_$name.json; tail -1 gpurun_out/r2s10_$name.err | cut -c1-200; }
run tma
run cpasync --no-tma
run tma_c2 --workload c2
run cpasync_c2 --workload c2 --no-tma
