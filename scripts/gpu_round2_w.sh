#!/bin/bash
# Round 2, GPU call W: the inverse table staged in shared memory (north star wording) against the L1/L2-served table
mkdir -p gpurun_out
V=gpurun_out/r02w_variants.log; : > $V
run() { timeout 300 python scripts/profile_case.py "$@" 2>&1 | grep -E "BEST|Error|error|Traceback" >> $V; }
for lib in "" sminv; do
  export MCB_LIB_VARIANT=$lib; [ -z "$lib" ] && unset MCB_LIB_VARIANT
  run --case c3 --photons 125000000 --batches 2 --tag "c3_lib=${lib:-default}"
  run --case c3 --photons 125000000 --batches 2 --blocks-per-sm 3 --tag "c3_occ3_lib=${lib:-default}"
  run --case c3 --photons 125000000 --batches 2 --blocks-per-sm 4 --tag "c3_occ4_lib=${lib:-default}"
done
cat $V
