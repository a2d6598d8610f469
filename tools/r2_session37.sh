#!/bin/bash
# Round 2, session 37: the C++ gen-pk host at full size (C3 from the in-memory generator), fp64 and fixed point
mkdir -p gpurun_out /tmp/pkout
for mode in "" "--fixed"; do
  genpk_b200/bin/gen-pk --synthetic clustered:1024:42 -g 1024 -o /tmp/pkout --json gpurun_out/r2s37_genpk_c3${mode/--/_}.json $mode > gpurun_out/r2s37_genpk_c3${mode/--/_}.log 2>&1; echo "rc=$?"
  tail -3 gpurun_out/r2s37_genpk_c3${mode/--/_}.log | head -2; cat gpurun_out/r2s37_genpk_c3${mode/--/_}.json | head -c 600; echo
  wc -l /tmp/pkout/PK-DM-*; head -3 /tmp/pkout/PK-DM-*
done
