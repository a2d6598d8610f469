#!/bin/bash
# 2-GPU: NCCL slab pipeline test, bench at N=2, FFT decomposition timings
mkdir -p gpurun_out
nvidia-smi -L
echo "== fft variants"; timeout 200 python tools/fft_variants.py 1024 2>&1 | tail -8
echo "== pytest slab (nccl)"; timeout 400 python -m pytest tests/test_gpu_slab.py -x -q --timeout 180 > gpurun_out/pytest_slab.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/pytest_slab.log
for w in c3 c2; do
echo "== bench N=2 $w"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --workload $w --steps 5 --warmup 3 --no-e2e > gpurun_out/bench_n2_$w.json 2> gpurun_out/bench_n2_$w.err
echo "rc=$?"; tail -c 2500 gpurun_out/bench_n2_$w.json; tail -5 gpurun_out/bench_n2_$w.err
done
