#!/bin/bash
# Session 12: fused x-pass/binning kernel (fftx_power_kernel): parity, bench A/B, ncu.
mkdir -p gpurun_out
echo "== pytest fftx"; timeout 600 python -m pytest tests/test_gpu_fftx.py -x -q > gpurun_out/s12_pytest_fftx.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/s12_pytest_fftx.log
echo "== pytest gpu (rest)"; timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_fftx.py > gpurun_out/s12_pytest_gpu.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/s12_pytest_gpu.log
show() {
python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, {k:round(v,3) for k,v in d["stage_ms"].items()}, {k:round(v["frac"],3) for k,v in d["roofline_all"].items()}, "e2e", d.get("e2e") and round(d["e2e"]["ms_per_step"],2), d.get("clocks"))
except Exception as e: print("ERR", e)
PY
}
echo "== bench c3 fused (default line)"
timeout 600 python bench.py > gpurun_out/s12_bench_default.json 2> gpurun_out/s12_bench_default.err; echo "rc=$?"; show gpurun_out/s12_bench_default.json; tail -3 gpurun_out/s12_bench_default.err
echo "== bench c3 unfused"
timeout 300 python bench.py --no-fused-xpass --no-e2e --no-cpu-baseline > gpurun_out/s12_bench_c3_unfused.json 2> gpurun_out/s12_bench_c3_unfused.err; echo "rc=$?"; show gpurun_out/s12_bench_c3_unfused.json; tail -3 gpurun_out/s12_bench_c3_unfused.err
echo "== bench c2 fused"
timeout 300 python bench.py --workload c2 --no-cpu-baseline > gpurun_out/s12_bench_c2.json 2> gpurun_out/s12_bench_c2.err; echo "rc=$?"; show gpurun_out/s12_bench_c2.json; tail -3 gpurun_out/s12_bench_c2.err
echo "== ncu launch list (c3)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/s12_launches_c3.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/s12_ncu_list.log 2>&1; echo "rc=$?"
echo "== ncu full: fftx_power_kernel (c3)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"fftx_power" -s 2 -c 1 -o gpurun_out/s12_prof_fftx -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/s12_ncu_full.log 2>&1; echo "rc=$?"; ls -la gpurun_out/*.ncu-rep
