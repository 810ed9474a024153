#!/bin/bash
# Round 2, GPU call H: leaps limited only by the boundary ahead; layer leaps through the clear sky of bitmap-marched fields (C5)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_leap.py tests/test_gpu_pool.py tests/test_gpu_bounds.py tests/test_cpp_host.py -q > gpurun_out/r02h_pytest_leap.log 2>&1; echo "rc=$?" >> gpurun_out/r02h_pytest_leap.log
tail -30 gpurun_out/r02h_pytest_leap.log | cut -c1-300
V=gpurun_out/r02h_variants.log; : > $V
run() { timeout 300 python scripts/profile_case.py "$@" 2>&1 | grep -E "BEST|Error|error|Traceback" >> $V; }
for leap in -1 0 3 6 8 12; do
  run --case c3 --photons 125000000 --batches 2 --leap $leap --tag c3_leap
done
for leap in -1 0 8 16; do
  run --case c3 --photons 16000000 --views --batches 2 --kernel 2 --blocks-per-sm 5 --leap $leap --tag c3_views_pool_leap
done
for leap in -1 0 8 16 32; do
  run --case c5 --photons 20000000 --batches 2 --leap $leap --tag c5_leap
done
run --case c5 --photons 20000000 --batches 2 --leap 0 --blocks-per-sm 7 --tag c5_leap_occ7
run --case c5 --photons 20000000 --batches 2 --leap 0 --burst 44 --tag c5_leap_b44
cat $V
timeout 600 python -m pytest tests/test_gpu_headline.py -q -k "maps" > gpurun_out/r02h_pytest_headline.log 2>&1; echo "rc=$?" >> gpurun_out/r02h_pytest_headline.log
tail -5 gpurun_out/r02h_pytest_headline.log | cut -c1-300
