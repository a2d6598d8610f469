#!/bin/bash
# Round 2, session 28: x pass with 4096-mode tiles (two CTAs per SM) now that the tiles are TMA-filled with 128-byte L2 promotion
mkdir -p gpurun_out
run() { name=$1; shift; timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline "$@" 2>/dev/null | tail -1 > gpurun_out/r2s28_$name.json; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2s28_$name.json").read()); print("$name", round(d["ms_per_step"],3), {k:round(v,3) for k,v in d["stage_ms"].items()}, d["clocks"]["sm_mhz"], (d.get("self_check") or {}).get("pk",{}).get("max_rel_power"))
except Exception as e: print("$name failed", e)
PY
}
run wide
run narrow --xpass-narrow-tile
run wide2
run narrow2 --xpass-narrow-tile
