#!/bin/bash
# Session 20: march kernel with register-prefetched loads (no smem staging), SLAB template; full GPU suite.
mkdir -p gpurun_out
echo "== pytest gpu (all)"; timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/s20_pytest_gpu.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/s20_pytest_gpu.log
show() {
python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, {k:round(v,3) for k,v in d["stage_ms"].items()}, {k:round(v["frac"],3) for k,v in d["roofline_all"].items()}, "e2e", d.get("e2e") and round(d["e2e"]["ms_per_step"],2))
except Exception as e: print("ERR", e)
PY
}
echo "== bench c3"
timeout 300 python bench.py --no-e2e --no-cpu-baseline > gpurun_out/s20_bench_c3.json 2> gpurun_out/s20_bench_c3.err; echo "rc=$?"; show gpurun_out/s20_bench_c3.json; tail -3 gpurun_out/s20_bench_c3.err
echo "== bench c3, march kernel built for 2 CTAs/SM (128 registers)"
GENPK_LIB=$PWD/genpk_b200/libgenpk_cuda_minb2.so timeout 300 python bench.py --no-e2e --no-cpu-baseline > gpurun_out/s20_bench_c3_minb2.json 2> gpurun_out/s20_bench_c3_minb2.err; echo "rc=$?"; show gpurun_out/s20_bench_c3_minb2.json; tail -3 gpurun_out/s20_bench_c3_minb2.err
echo "== bench c3 fixed point"
timeout 300 python bench.py --no-e2e --no-cpu-baseline --fixed-point > gpurun_out/s20_bench_c3_fixed.json 2> gpurun_out/s20_bench_c3_fixed.err; echo "rc=$?"; show gpurun_out/s20_bench_c3_fixed.json; tail -3 gpurun_out/s20_bench_c3_fixed.err
echo "== ncu full: march (c3)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"deposit_march" -s 3 -c 1 -o gpurun_out/s20_prof_march -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/s20_ncu_full.log 2>&1; echo "rc=$?"; ls -la gpurun_out/s20*.ncu-rep
