#!/bin/bash
# Round 2, GPU call F: the whole GPU suite on the current tree; C5 with whole-layer clear-sky flags (no bitmap look-up in cloud-free layers)
mkdir -p gpurun_out
timeout 2700 python -m pytest tests -q -m gpu > gpurun_out/r02f_pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/r02f_pytest_gpu.log
tail -8 gpurun_out/r02f_pytest_gpu.log | cut -c1-300
V=gpurun_out/r02f_variants.log; : > $V
run() { timeout 300 python scripts/profile_case.py "$@" 2>&1 | grep -E "BEST|Error|error|Traceback" >> $V; }
for occ in 6 7; do for burst in 8 44; do
  run --case c5 --photons 20000000 --kernel 2 --blocks-per-sm $occ --burst $burst --batches 2 --tag pool_c5_layerflag
done; done
run --case c5 --photons 20000000 --kernel 1 --batches 2 --tag park_c5_layerflag
run --case c3 --photons 125000000 --batches 2 --tag c3_default
run --case c3mie --photons 64000000 --batches 2 --tag c3mie_default
run --case c1 --photons 64000000 --batches 2 --tag c1_default
run --case c2 --photons 64000000 --batches 2 --tag c2_default
run --case c4 --photons 64000000 --batches 2 --tag c4_default
run --case c2 --photons 16000000 --views --batches 2 --tag c2_views_default
run --case c2 --photons 16000000 --views --batches 2 --kernel 2 --tag c2_views_pool
cat $V
