#!/bin/bash
# Round 2, GPU call S: final validation of the tree -- whole GPU suite, smoke, bench lines of every workload, ncu of the C5 variant
mkdir -p gpurun_out
timeout 2700 python -m pytest tests -q -m gpu > gpurun_out/r02s_pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/r02s_pytest_gpu.log
tail -8 gpurun_out/r02s_pytest_gpu.log | cut -c1-300
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02s_smoke.log 2>&1; tail -2 gpurun_out/r02s_smoke.log | cut -c1-300
timeout 900 python bench.py > gpurun_out/r02s_bench_c3.log 2> gpurun_out/r02s_bench_c3.err
timeout 900 python bench.py --views --no-cpu-baseline > gpurun_out/r02s_bench_views.log 2> gpurun_out/r02s_bench_views.err
timeout 900 python bench.py --workload c5 --no-cpu-baseline > gpurun_out/r02s_bench_c5.log 2> gpurun_out/r02s_bench_c5.err
timeout 900 python bench.py --workload broadband --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02s_bench_bb.log 2> gpurun_out/r02s_bench_bb.err
for f in c3 views c5 bb; do echo "== $f"; tail -c 400 gpurun_out/r02s_bench_$f.log; tail -3 gpurun_out/r02s_bench_$f.err; done
V=gpurun_out/r02s_variants.log; : > $V
run() { timeout 300 python scripts/profile_case.py "$@" 2>&1 | grep -E "BEST|Error|error|Traceback" >> $V; }
run --case c5 --photons 125000000 --batches 2 --ext-mask 1 --tag c5_bitmap_big
run --case c5 --photons 125000000 --batches 2 --tag c5_default_big
run --case c3 --photons 125000000 --batches 2 --tag c3_default
run --case c3mie --photons 64000000 --batches 2 --tag c3mie_default
cat $V
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'pool_kernel' -s 1 -c 1 -f \
    -o gpurun_out/r02s_prof_c5_crop python scripts/profile_case.py --case c5 --photons 10000000 --batches 2 > gpurun_out/r02s_ncu_crop.log 2>&1
ls -la gpurun_out/r02s*
