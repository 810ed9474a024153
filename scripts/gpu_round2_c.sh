#!/bin/bash
# Round 2, GPU call C: pool kernel v2 (deferred decode, legs set up in the event phase, split gathers)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pool.py -q -x > gpurun_out/r02c_pytest_pool.log 2>&1; echo "rc=$?" >> gpurun_out/r02c_pytest_pool.log
tail -3 gpurun_out/r02c_pytest_pool.log
V=gpurun_out/r02c_variants.log; : > $V
run() { timeout 300 python scripts/profile_case.py "$@" 2>&1 | grep -E "BEST|Error|error|Traceback" >> $V; }
for case in c3 c5 c1 c4 c3mie; do
  ph=16000000; [ $case = c5 ] && ph=8000000
  run --case $case --photons $ph --kernel 1 --tag park
  for occ in 6 8; do for burst in 8 44 4; do
    run --case $case --photons $ph --kernel 2 --blocks-per-sm $occ --burst $burst --tag pool2
  done; done
done
run --case c3 --photons 125000000 --kernel 1 --batches 2 --tag park_full
run --case c3 --photons 125000000 --kernel 2 --blocks-per-sm 6 --burst 8 --batches 2 --tag pool2_full
run --case c3 --photons 125000000 --kernel 2 --blocks-per-sm 6 --burst 44 --batches 2 --tag pool2_full
run --case c3 --photons 125000000 --kernel 2 --blocks-per-sm 8 --burst 44 --batches 2 --tag pool2_full
cat $V
for b in 8 44; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'pool_kernel' -s 1 -c 1 -f \
    -o gpurun_out/r02c_prof_c3_pool2_b$b python scripts/profile_case.py --case c3 --photons 16000000 --batches 2 --kernel 2 --burst $b \
    > gpurun_out/r02c_ncu_pool2_b$b.log 2>&1
done
# baselines for the next steps: local estimation and C5 through bench.py
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --views --photons 16000000 > gpurun_out/r02c_bench_views.log 2>&1
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --workload c5 --photons 20000000 > gpurun_out/r02c_bench_c5.log 2>&1
tail -2 gpurun_out/r02c_bench_views.log gpurun_out/r02c_bench_c5.log
