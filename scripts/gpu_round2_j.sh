#!/bin/bash
# Round 2, GPU call J: leap and burst in the same iteration (no landing gather, no lost iteration)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_leap.py tests/test_gpu_pool.py tests/test_gpu_bounds.py -q -s > gpurun_out/r02j_pytest_leap.log 2>&1; echo "rc=$?" >> gpurun_out/r02j_pytest_leap.log
grep -E "passed|failed|leaps per photon|^FAILED|^E  " gpurun_out/r02j_pytest_leap.log | cut -c1-300 | head -40
V=gpurun_out/r02j_variants.log; : > $V
run() { timeout 300 python scripts/profile_case.py "$@" 2>&1 | grep -E "BEST|Error|error|Traceback" >> $V; }
run --case c3 --photons 125000000 --batches 2 --leap -1 --tag c3_noleap
for lanes in 0 4 8; do for leap in 3 4 6 8; do
  run --case c3 --photons 125000000 --batches 2 --leap $leap --leap-lanes $lanes --tag c3_leap
done; done
run --case c3 --photons 125000000 --batches 2 --leap 4 --blocks-per-sm 6 --tag c3_leap_occ6
run --case c5 --photons 20000000 --batches 2 --leap -1 --tag c5_noleap
for lanes in 0 4 8; do
  run --case c5 --photons 20000000 --batches 2 --leap 4 --leap-lanes $lanes --tag c5_leap
done
run --case c5 --photons 20000000 --batches 2 --leap 4 --burst 44 --tag c5_leap_b44
run --case c5 --photons 20000000 --batches 2 --leap 4 --blocks-per-sm 7 --tag c5_leap_occ7
run --case c3 --photons 16000000 --views --batches 2 --kernel 2 --blocks-per-sm 5 --leap -1 --tag c3_views_pool_noleap
for lanes in 0 4 8; do
  run --case c3 --photons 16000000 --views --batches 2 --kernel 2 --blocks-per-sm 5 --leap 4 --leap-lanes $lanes --tag c3_views_pool_leap
done
cat $V
