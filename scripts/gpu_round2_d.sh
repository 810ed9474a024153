#!/bin/bash
# Round 2, GPU call D: the whole GPU suite, bench lines of every workload, 7-CTA pool variant
mkdir -p gpurun_out
timeout 2700 python -m pytest tests -q -m gpu > gpurun_out/r02d_pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/r02d_pytest_gpu.log
tail -6 gpurun_out/r02d_pytest_gpu.log
V=gpurun_out/r02d_variants.log; : > $V
run() { timeout 300 python scripts/profile_case.py "$@" 2>&1 | grep -E "BEST|Error|error|Traceback" >> $V; }
for occ in 6 7; do for burst in 8 44; do
  run --case c3 --photons 125000000 --kernel 2 --blocks-per-sm $occ --burst $burst --batches 2 --tag pool2_full
  run --case c5 --photons 20000000 --kernel 2 --blocks-per-sm $occ --burst $burst --batches 2 --tag pool2_c5
done; done
run --case c3 --photons 16000000 --views --batches 2 --tag le_base
run --case c2 --photons 16000000 --views --batches 2 --tag le_base
cat $V
timeout 900 python bench.py > gpurun_out/r02d_bench_c3.log 2> gpurun_out/r02d_bench_c3.err
timeout 900 python bench.py --views --no-cpu-baseline > gpurun_out/r02d_bench_views.log 2> gpurun_out/r02d_bench_views.err
timeout 900 python bench.py --workload c5 --photons 20000000 --no-cpu-baseline > gpurun_out/r02d_bench_c5.log 2> gpurun_out/r02d_bench_c5.err
timeout 900 python bench.py --workload broadband --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02d_bench_bb.log 2> gpurun_out/r02d_bench_bb.err
for f in c3 views c5 bb; do echo "== $f"; tail -c 700 gpurun_out/r02d_bench_$f.log; tail -3 gpurun_out/r02d_bench_$f.err; done
