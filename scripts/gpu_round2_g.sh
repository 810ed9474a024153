#!/bin/bash
# Round 2, GPU call G: vacuum leaps (distance-encoded field) -- tests, then the sweep; C5 whole-layer flags
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_leap.py tests/test_gpu_pool.py tests/test_gpu_bounds.py -q -x > gpurun_out/r02g_pytest_leap.log 2>&1; echo "rc=$?" >> gpurun_out/r02g_pytest_leap.log
tail -30 gpurun_out/r02g_pytest_leap.log | cut -c1-300
V=gpurun_out/r02g_variants.log; : > $V
run() { timeout 300 python scripts/profile_case.py "$@" 2>&1 | grep -E "BEST|Error|error|Traceback|batch 1" >> $V; }
for leap in -1 0 2 3 6 10; do
  run --case c3 --photons 125000000 --batches 2 --leap $leap --tag c3_leap
done
run --case c3 --photons 125000000 --batches 2 --leap 0 --blocks-per-sm 6 --tag c3_leap_occ6
run --case c3 --photons 125000000 --batches 2 --leap 0 --burst 8 --tag c3_leap_b8
for leap in -1 0 3 8; do
  run --case c3 --photons 16000000 --views --batches 2 --kernel 2 --blocks-per-sm 5 --leap $leap --tag c3_views_pool_leap
done
run --case c3 --photons 16000000 --views --batches 2 --kernel 2 --blocks-per-sm 5 --layout 2 --tag c3_views_pool_leap_brick
run --case c3 --photons 16000000 --views --batches 2 --kernel 2 --blocks-per-sm 4 --tag c3_views_pool_leap_occ4
run --case c3mie --photons 64000000 --batches 2 --tag c3mie_default
for occ in 6 7; do for burst in 8 44; do
  run --case c5 --photons 20000000 --kernel 2 --blocks-per-sm $occ --burst $burst --batches 2 --tag pool_c5_layerflag
done; done
run --case c1 --photons 64000000 --batches 2 --tag c1_default
run --case c2 --photons 64000000 --batches 2 --tag c2_default
run --case c4 --photons 64000000 --batches 2 --tag c4_default
run --case c2 --photons 16000000 --views --batches 2 --tag c2_views_default
cat $V
timeout 2400 python -m pytest tests -q -m gpu --deselect tests/test_gpu_leap.py --deselect tests/test_gpu_pool.py --deselect tests/test_gpu_bounds.py > gpurun_out/r02g_pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/r02g_pytest_gpu.log
tail -12 gpurun_out/r02g_pytest_gpu.log | cut -c1-300
