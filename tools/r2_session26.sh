#!/bin/bash
# Round 2, session 26 (1 GPU): full GPU suite, smoke, default bench + reference arm, C2 line, launch list,
# ncu --set full of the three dominant kernels (-> profiles/ncu_traffic.json), compute-sanitizer on the new kernel
mkdir -p gpurun_out
echo "== pytest gpu (all)"; timeout 1800 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r2s26_pytest_gpu.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/r2s26_pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
show() {
python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ("value","ms_per_step","gpu_launches","cufft_execs_per_step")}, {k:round(v,3) for k,v in d.get("stage_ms",{}).items()}, {k:round(v["frac"],3) for k,v in d.get("roofline_all",{}).items()}, "e2e", d.get("e2e") and round(d["e2e"].get("ms_per_step",0),2), d.get("clocks"), d.get("cpu_baseline",{}).get("value"), d.get("self_check") and list(d["self_check"]))
except Exception as e: print("ERR", e)
PY
}
echo "== bench reference arm"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2s26_bench_reference.json 2> gpurun_out/r2s26_bench_reference.err; echo "rc=$?"; show gpurun_out/r2s26_bench_reference.json
echo "== bench default"
timeout 600 python bench.py > gpurun_out/r2s26_bench_default.json 2> gpurun_out/r2s26_bench_default.err; echo "rc=$?"; show gpurun_out/r2s26_bench_default.json; tail -3 gpurun_out/r2s26_bench_default.err
echo "== bench c2"
timeout 300 python bench.py --workload c2 --no-cpu-baseline > gpurun_out/r2s26_bench_c2.json 2> gpurun_out/r2s26_bench_c2.err; echo "rc=$?"; show gpurun_out/r2s26_bench_c2.json; tail -3 gpurun_out/r2s26_bench_c2.err
echo "== ncu launch list (c3)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2s26_launches_c3.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2s26_ncu_list.log 2>&1; echo "rc=$?"
echo "== ncu --set full: march, fft_zy, fftx_power (one launch each)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"deposit_march_kernel|fft_zy_kernel|fftx_power_kernel" -c 3 -o gpurun_out/r2s26_c3 --force-overwrite python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-self-check > gpurun_out/r2s26_ncu_full.log 2>&1; echo "rc=$?"; ls -la gpurun_out/r2s26_c3.ncu-rep
echo "== compute-sanitizer (256^3 through march + fft_zy + fftx_power; slab scatter variant emulated on one GPU)"
cat > /tmp/san.py <<'PY'
import sys, torch, numpy as np
sys.path.insert(0, ".")
import genpk_b200 as gp
from genpk_b200 import api
n_side = dims = 256
n = n_side ** 3
d = torch.empty(3 * n, dtype=torch.float32, device="cuda")
api.synth_particles_dev(api.SYNTH_CLUSTERED, 42, n_side, 0, n, 1000.0, dims, d.data_ptr()); torch.cuda.synchronize()
for fixed in (0, 1):
    with gp.Context(dims, flags=api.FLAG_FIXED_POINT if fixed else 0) as ctx:
        ctx.grid_zero(); ctx.deposit_dev(d.data_ptr(), n, 0, 1.0, 1000.0)
        p, c, k = ctx.fft_power(dims, float(n), float(n)); ctx.synchronize()
        print("fixed", fixed, "library calls", ctx.library_calls(), "counts", int(c.astype(np.int64).sum()), "P[10]", p[10])
PY
timeout 900 compute-sanitizer --tool memcheck python /tmp/san.py > gpurun_out/r2s26_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -3 gpurun_out/r2s26_sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck python /tmp/san.py > gpurun_out/r2s26_sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -3 gpurun_out/r2s26_sanitizer_racecheck.log
