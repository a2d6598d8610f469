#!/bin/bash
# Session 21: round-1 closing evidence on one GPU: full GPU suite, default bench line (e2e, cpu_baseline),
# reference arm, C2 line, ncu launch list and full captures of the three dominant kernels.
mkdir -p gpurun_out
echo "== pytest gpu (all)"; timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/s21_pytest_gpu.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/s21_pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
show() {
python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ("value","ms_per_step","gpu_launches")}, {k:round(v,3) for k,v in d.get("stage_ms",{}).items()}, {k:round(v["frac"],3) for k,v in d.get("roofline_all",{}).items()}, "e2e", d.get("e2e") and round(d["e2e"].get("ms_per_step",0),2), d.get("clocks"), d.get("cpu_baseline",{}).get("value"))
except Exception as e: print("ERR", e)
PY
}
echo "== bench reference arm"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/s21_bench_reference.json 2> gpurun_out/s21_bench_reference.err; echo "rc=$?"; show gpurun_out/s21_bench_reference.json
echo "== bench default"
timeout 600 python bench.py > gpurun_out/s21_bench_default.json 2> gpurun_out/s21_bench_default.err; echo "rc=$?"; show gpurun_out/s21_bench_default.json; tail -3 gpurun_out/s21_bench_default.err
echo "== bench c2"
timeout 300 python bench.py --workload c2 --no-cpu-baseline > gpurun_out/s21_bench_c2.json 2> gpurun_out/s21_bench_c2.err; echo "rc=$?"; show gpurun_out/s21_bench_c2.json; tail -3 gpurun_out/s21_bench_c2.err
echo "== bench c3 library FFT + binning"
timeout 300 python bench.py --no-fused-xpass --no-own-ypass --no-e2e --no-cpu-baseline > gpurun_out/s21_bench_c3_libfft.json 2> gpurun_out/s21_bench_c3_libfft.err; echo "rc=$?"; show gpurun_out/s21_bench_c3_libfft.json
echo "== ncu launch list (c3)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/s21_launches_c3.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/s21_ncu_list.log 2>&1; echo "rc=$?"
echo "== ncu full: march + fft_cols + fftx_power (c3)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"deposit_march|fftx_power|fft_cols" -s 6 -c 3 -o gpurun_out/s21_prof_c3 -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/s21_ncu_full.log 2>&1; echo "rc=$?"; ls -la gpurun_out/s21*.ncu-rep
