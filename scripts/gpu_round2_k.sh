#!/bin/bash
# Round 2, GPU call K: whole GPU suite on the leap build (pool LE default), tuning around the defaults, bench lines, ncu captures
mkdir -p gpurun_out
timeout 2700 python -m pytest tests -q -m gpu > gpurun_out/r02k_pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/r02k_pytest_gpu.log
tail -12 gpurun_out/r02k_pytest_gpu.log | cut -c1-300
V=gpurun_out/r02k_variants.log; : > $V
run() { timeout 300 python scripts/profile_case.py "$@" 2>&1 | grep -E "BEST|Error|error|Traceback" >> $V; }
for lanes in 6 8 12; do for leap in 2 3 4; do
  run --case c3 --photons 125000000 --batches 2 --leap $leap --leap-lanes $lanes --tag c3_leap
done; done
for lanes in 8 12 16; do
  run --case c5 --photons 20000000 --batches 2 --leap 3 --leap-lanes $lanes --tag c5_leap
  run --case c3 --photons 16000000 --views --batches 2 --leap 3 --leap-lanes $lanes --tag c3_views_leap
done
run --case c3mie --photons 64000000 --batches 2 --tag c3mie_default
run --case c2 --photons 16000000 --views --batches 2 --tag c2_views_default
cat $V
timeout 900 python bench.py > gpurun_out/r02k_bench_c3.log 2> gpurun_out/r02k_bench_c3.err
timeout 900 python bench.py --views --no-cpu-baseline > gpurun_out/r02k_bench_views.log 2> gpurun_out/r02k_bench_views.err
timeout 900 python bench.py --workload c5 --photons 20000000 --no-cpu-baseline > gpurun_out/r02k_bench_c5.log 2> gpurun_out/r02k_bench_c5.err
timeout 900 python bench.py --workload broadband --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02k_bench_bb.log 2> gpurun_out/r02k_bench_bb.err
for f in c3 views c5 bb; do echo "== $f"; tail -c 600 gpurun_out/r02k_bench_$f.log; tail -3 gpurun_out/r02k_bench_$f.err; done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'pool_kernel' -s 1 -c 1 -f \
    -o gpurun_out/r02k_prof_c3_pool_leap python scripts/profile_case.py --case c3 --photons 16000000 --batches 2 > gpurun_out/r02k_ncu_c3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'pool_le_kernel' -s 1 -c 1 -f \
    -o gpurun_out/r02k_prof_c3_views_pool_leap python scripts/profile_case.py --case c3 --photons 4000000 --views --batches 2 > gpurun_out/r02k_ncu_views.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'pool_kernel' -s 1 -c 1 -f \
    -o gpurun_out/r02k_prof_c5_pool_leap python scripts/profile_case.py --case c5 --photons 10000000 --batches 2 > gpurun_out/r02k_ncu_c5.log 2>&1
ls -la gpurun_out/r02k*
