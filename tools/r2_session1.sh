#!/bin/bash
# Round 2, session 1: new sweep deposit kernel -- correctness first (tight timeouts: the zero-ahead
# variant spins on counters), then A/B against the march kernel on C3, windows / slack, and the
# full-size parity tests + golden fixture.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader
free -g | head -2
echo "== sweep tests"; timeout 900 python -m pytest tests/test_gpu_sweep.py -m gpu -x -q --timeout 120 > gpurun_out/r2s1_pytest_sweep.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/r2s1_pytest_sweep.log
show() {
python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ("value","ms_per_step","gpu_launches")}, {k:round(v,3) for k,v in d.get("stage_ms",{}).items()}, {k:round(v["frac"],3) for k,v in d.get("roofline_all",{}).items()}, d["config"].get("sweep"), d.get("self_check") and list(d["self_check"]))
except Exception as e: print("ERR", e)
PY
}
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-self-check"
echo "== c3 march (round-1 kernel)"; timeout 300 $B --no-sweep > gpurun_out/r2s1_c3_march.json 2> gpurun_out/r2s1_c3_march.err; echo "rc=$?"; show gpurun_out/r2s1_c3_march.json
echo "== c3 sweep, memset"; timeout 300 $B --no-zero-ahead > gpurun_out/r2s1_c3_sweep_memset.json 2> gpurun_out/r2s1_c3_sweep_memset.err; echo "rc=$?"; show gpurun_out/r2s1_c3_sweep_memset.json; tail -2 gpurun_out/r2s1_c3_sweep_memset.err
echo "== c3 sweep, zero ahead (auto window)"; timeout 300 $B > gpurun_out/r2s1_c3_sweep_za.json 2> gpurun_out/r2s1_c3_sweep_za.err; echo "rc=$?"; show gpurun_out/r2s1_c3_sweep_za.json; tail -2 gpurun_out/r2s1_c3_sweep_za.err
for W in 4 6 8 12; do for A in 1 3; do
echo "== c3 sweep za window $W slack $A"; timeout 200 $B --za-window $W --za-slack $A > gpurun_out/r2s1_c3_za_w${W}_a${A}.json 2>/dev/null; echo "rc=$?"; show gpurun_out/r2s1_c3_za_w${W}_a${A}.json
done; done
echo "== c3 sweep fixed point"; timeout 300 $B --fixed-point > gpurun_out/r2s1_c3_sweep_fixed.json 2>/dev/null; echo "rc=$?"; show gpurun_out/r2s1_c3_sweep_fixed.json
echo "== golden full-size fixture (reference objects on the host cores)"
timeout 1500 python tests/golden/make_golden_fullsize.py gpurun_out/fullsize_pk.npz 2>&1 | tail -4
cp gpurun_out/fullsize_pk.npz tests/golden/fullsize_pk.npz 2>/dev/null
echo "== full-size parity tests"; timeout 1500 python -m pytest tests/test_gpu_fullsize.py -m gpu -x -q --timeout 900 > gpurun_out/r2s1_pytest_fullsize.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/r2s1_pytest_fullsize.log
echo "== ncu full: sweep (c3, zero ahead)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"deposit_sweep" -s 3 -c 1 -o gpurun_out/r2s1_prof_sweep -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-self-check > gpurun_out/r2s1_ncu.log 2>&1; echo "rc=$?"; ls -la gpurun_out/*.ncu-rep
