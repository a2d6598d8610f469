#!/bin/bash
# Session 13: fused x-pass with two CTAs per SM (4096-mode tiles) vs one (8192): parity, bench A/B, ncu.
mkdir -p gpurun_out
echo "== pytest fftx + march"; timeout 900 python -m pytest tests/test_gpu_fftx.py tests/test_gpu_march.py -q > gpurun_out/s13_pytest_fftx.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/s13_pytest_fftx.log
show() {
python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, {k:round(v,3) for k,v in d["stage_ms"].items()}, {k:round(v["frac"],3) for k,v in d["roofline_all"].items()}, "e2e", d.get("e2e") and round(d["e2e"]["ms_per_step"],2), d.get("clocks"))
except Exception as e: print("ERR", e)
PY
}
echo "== bench c3 fused, 2 CTAs/SM"
timeout 300 python bench.py --no-e2e --no-cpu-baseline > gpurun_out/s13_bench_c3.json 2> gpurun_out/s13_bench_c3.err; echo "rc=$?"; show gpurun_out/s13_bench_c3.json; tail -3 gpurun_out/s13_bench_c3.err
echo "== bench c3 fused, wide tile"
timeout 300 python bench.py --xpass-wide-tile --no-e2e --no-cpu-baseline > gpurun_out/s13_bench_c3_wide.json 2> gpurun_out/s13_bench_c3_wide.err; echo "rc=$?"; show gpurun_out/s13_bench_c3_wide.json; tail -3 gpurun_out/s13_bench_c3_wide.err
echo "== bench c2 fused"
timeout 300 python bench.py --workload c2 --no-cpu-baseline --no-e2e > gpurun_out/s13_bench_c2.json 2> gpurun_out/s13_bench_c2.err; echo "rc=$?"; show gpurun_out/s13_bench_c2.json; tail -3 gpurun_out/s13_bench_c2.err
echo "== ncu full: fftx_power_kernel (c3)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"fftx_power" -s 2 -c 1 -o gpurun_out/s13_prof_fftx -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/s13_ncu_full.log 2>&1; echo "rc=$?"; ls -la gpurun_out/s13*.ncu-rep
