#!/bin/bash
# Round 2, GPU call X: preferred shared-memory carve-out of the pool flux kernel (how much of the 256 KB stays L1)
mkdir -p gpurun_out
V=gpurun_out/r02x2_variants.log; : > $V
run() { timeout 300 python scripts/profile_case.py "$@" 2>&1 | grep -E "BEST|Error|error|Traceback" >> $V; }
export MCB_LIB_VARIANT=carve
for c in 0 40 50 55 60 75 100; do
  run --case c3 --photons 125000000 --batches 2 --park-threshold $c --tag "c3_carveout=$c"
done
run --case c5 --photons 125000000 --batches 2 --park-threshold 50 --tag "c5_carveout=50"
run --case c5 --photons 125000000 --batches 2 --park-threshold 0 --tag "c5_carveout=0"
cat $V
