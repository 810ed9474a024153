#!/bin/bash
# Round 2, GPU call L: A/B in one process environment: leaps with per-warp counters / without counters / compiled out
mkdir -p gpurun_out
V=gpurun_out/r02l_variants.log; : > $V
run() { timeout 300 python scripts/profile_case.py "$@" 2>&1 | grep -E "BEST|Error|error|Traceback" >> $V; }
for rep in 1 2; do
for lib in "" nostats noleap; do
  export MCB_LIB_VARIANT=$lib; [ -z "$lib" ] && unset MCB_LIB_VARIANT
  run --case c3 --photons 125000000 --batches 2 --tag "c3_lib=${lib:-default}"
  run --case c3 --photons 125000000 --batches 2 --leap 3 --leap-lanes 1 --tag "c3_lanes1_lib=${lib:-default}"
  run --case c5 --photons 20000000 --batches 2 --tag "c5_lib=${lib:-default}"
  run --case c3 --photons 16000000 --views --batches 2 --tag "c3views_lib=${lib:-default}"
  run --case c3mie --photons 64000000 --batches 2 --tag "c3mie_lib=${lib:-default}"
done; done
unset MCB_LIB_VARIANT
cat $V
