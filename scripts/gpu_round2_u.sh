#!/bin/bash
# Round 2, GPU call U: local-estimation A/B (split gathers, 4-cell bursts), leap threshold on C3 after the branchless counts
mkdir -p gpurun_out
V=gpurun_out/r02u_variants.log; : > $V
run() { timeout 300 python scripts/profile_case.py "$@" 2>&1 | grep -E "BEST|Error|error|Traceback" >> $V; }
for lib in "" lesplit leb4; do
  export MCB_LIB_VARIANT=$lib; [ -z "$lib" ] && unset MCB_LIB_VARIANT
  run --case c3 --photons 16000000 --views --batches 2 --tag "c3views_lib=${lib:-default}"
  run --case c3 --photons 16000000 --views --batches 2 --leap 2 --tag "c3views_leap2_lib=${lib:-default}"
done
unset MCB_LIB_VARIANT
for leap in 2 3 4; do run --case c3 --photons 125000000 --batches 2 --leap $leap --tag c3_leap; done
run --case c5 --photons 125000000 --batches 2 --leap 2 --tag c5_leap2
run --case c5 --photons 125000000 --batches 2 --leap 4 --tag c5_leap4
cat $V
