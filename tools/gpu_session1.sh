#!/bin/bash
# First GPU session: parity tests, smoke, primitive microbenchmarks, first bench lines, ncu launch list.
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvidia-smi.txt 2>&1
nproc > gpurun_out/host.txt; free -g >> gpurun_out/host.txt; grep -m1 "model name" /proc/cpuinfo >> gpurun_out/host.txt
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== microbench"; timeout 600 tools/microbench > gpurun_out/microbench.txt 2>&1; tail -70 gpurun_out/microbench.txt
for w in "c2 direct" "c2 sorted" "c3 direct" "c3 sorted"; do
  set -- $w
  echo "== bench $1 $2"
  timeout 900 python bench.py --workload $1 --deposit $2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$1_$2.json 2> gpurun_out/bench_$1_$2.err
  echo "rc=$?"; tail -c 1500 gpurun_out/bench_$1_$2.json; tail -3 gpurun_out/bench_$1_$2.err
done
echo "== bench c2 fixed"; timeout 600 python bench.py --workload c2 --fixed-point --steps 5 --no-cpu-baseline --no-e2e > gpurun_out/bench_c2_fixed.json 2>&1; tail -c 600 gpurun_out/bench_c2_fixed.json
echo "== ncu launch list (c2)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_c2.csv python bench.py --workload c2 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_c2.log 2>&1
echo "ncu rc=$?"; tail -5 gpurun_out/launches_c2.csv
