#!/bin/bash
# Round 2, GPU call M: whole GPU suite (LEAP kernel variants chosen at staging, counters in the debug build only, device table builders), bench lines, ncu captures
mkdir -p gpurun_out
timeout 2700 python -m pytest tests -q -m gpu > gpurun_out/r02m_pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/r02m_pytest_gpu.log
tail -12 gpurun_out/r02m_pytest_gpu.log | cut -c1-300
V=gpurun_out/r02m_variants.log; : > $V
run() { timeout 300 python scripts/profile_case.py "$@" 2>&1 | grep -E "BEST|Error|error|Traceback" >> $V; }
for leap in -1 0 2 4 6; do
  run --case c3 --photons 125000000 --batches 2 --leap $leap --tag c3_leap
done
run --case c3 --photons 125000000 --batches 2 --leap-lanes 8 --tag c3_leap_lanes8
for leap in -1 0; do
  run --case c5 --photons 20000000 --batches 2 --leap $leap --tag c5_leap
  run --case c3 --photons 16000000 --views --batches 2 --leap $leap --tag c3_views_leap
done
run --case c3mie --photons 64000000 --batches 2 --tag c3mie_default
run --case c2 --photons 16000000 --views --batches 2 --tag c2_views_default
cat $V
timeout 900 python bench.py > gpurun_out/r02m_bench_c3.log 2> gpurun_out/r02m_bench_c3.err
timeout 900 python bench.py --views --no-cpu-baseline > gpurun_out/r02m_bench_views.log 2> gpurun_out/r02m_bench_views.err
timeout 900 python bench.py --workload c5 --photons 20000000 --no-cpu-baseline > gpurun_out/r02m_bench_c5.log 2> gpurun_out/r02m_bench_c5.err
timeout 900 python bench.py --workload broadband --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02m_bench_bb.log 2> gpurun_out/r02m_bench_bb.err
for f in c3 views c5 bb; do echo "== $f"; tail -c 600 gpurun_out/r02m_bench_$f.log; tail -3 gpurun_out/r02m_bench_$f.err; done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'pool_kernel' -s 1 -c 1 -f \
    -o gpurun_out/r02m_prof_c3_pool_leap python scripts/profile_case.py --case c3 --photons 16000000 --batches 2 > gpurun_out/r02m_ncu_c3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'pool_le_kernel' -s 1 -c 1 -f \
    -o gpurun_out/r02m_prof_c3_views_pool_leap python scripts/profile_case.py --case c3 --photons 4000000 --views --batches 2 > gpurun_out/r02m_ncu_views.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'pool_kernel' -s 1 -c 1 -f \
    -o gpurun_out/r02m_prof_c5_pool_leap python scripts/profile_case.py --case c5 --photons 10000000 --batches 2 > gpurun_out/r02m_ncu_c5.log 2>&1
ls -la gpurun_out/r02m*
