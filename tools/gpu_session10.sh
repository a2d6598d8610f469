#!/bin/bash
# Session 10 (1 GPU): full GPU parity suite, default bench line (e2e + cpu_baseline), reference arm,
# ncu launch list of the default command, ncu --set full of the deposit and binning kernels.
mkdir -p gpurun_out
nvidia-smi -L
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -x -q --timeout 180 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu.log
echo "== smoke"; timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench default (c3)"
timeout 600 python bench.py > gpurun_out/s10_bench_default.json 2> gpurun_out/s10_bench_default.err; echo "rc=$?"
tail -c 3000 gpurun_out/s10_bench_default.json; tail -3 gpurun_out/s10_bench_default.err
echo "== bench c2"
timeout 300 python bench.py --workload c2 > gpurun_out/s10_bench_c2.json 2> gpurun_out/s10_bench_c2.err; echo "rc=$?"
tail -c 1500 gpurun_out/s10_bench_c2.json
echo "== reference arm"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/s10_bench_reference.json 2> gpurun_out/s10_bench_reference.err; echo "rc=$?"
tail -c 1500 gpurun_out/s10_bench_reference.json
echo "== ncu launch list (default command, short)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/s10_launches_c3.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_c3.log 2>&1
echo "ncu rc=$?"; tail -12 gpurun_out/s10_launches_c3.csv | cut -c1-200
echo "== ncu full: march + binning (c3)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"deposit_march|bin_power" -s 4 -c 4 -o gpurun_out/s10_prof_c3 -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full_c3.log 2>&1
echo "ncu full rc=$?"; ls -la gpurun_out/*.ncu-rep
