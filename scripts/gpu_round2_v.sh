#!/bin/bash
# Round 2, GPU call V: the cropped C5 field as a persisting L2 window (experiment)
mkdir -p gpurun_out
V=gpurun_out/r02v_variants.log; : > $V
run() { timeout 300 python scripts/profile_case.py "$@" 2>&1 | grep -E "BEST|Error|error|Traceback|persisting" >> $V; }
run --case c5 --photons 125000000 --batches 2 --ext-mask 2 --tag c5_crop
run --case c5 --photons 125000000 --batches 3 --ext-mask 3 --tag c5_crop_persist
run --case c5 --photons 125000000 --batches 2 --ext-mask 2 --tag c5_crop
run --case c5 --photons 125000000 --batches 3 --ext-mask 3 --tag c5_crop_persist
cat $V
