#!/bin/bash
# Round 2, GPU call N: the whole GPU suite (new table-builder tests), bench lines with the leap share, ncu --set full of the bench-size launches (DRAM traffic per launch)
mkdir -p gpurun_out
timeout 2700 python -m pytest tests -q -m gpu > gpurun_out/r02n_pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/r02n_pytest_gpu.log
tail -12 gpurun_out/r02n_pytest_gpu.log | cut -c1-300
timeout 900 python bench.py > gpurun_out/r02n_bench_c3.log 2> gpurun_out/r02n_bench_c3.err
timeout 900 python bench.py --views --no-cpu-baseline > gpurun_out/r02n_bench_views.log 2> gpurun_out/r02n_bench_views.err
timeout 900 python bench.py --workload c5 --no-cpu-baseline > gpurun_out/r02n_bench_c5.log 2> gpurun_out/r02n_bench_c5.err
timeout 900 python bench.py --workload broadband --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02n_bench_bb.log 2> gpurun_out/r02n_bench_bb.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02n_bench_ref.log 2> gpurun_out/r02n_bench_ref.err
for f in c3 views c5 bb ref; do echo "== $f"; tail -c 500 gpurun_out/r02n_bench_$f.log; tail -3 gpurun_out/r02n_bench_$f.err; done
# launch list of the bench command (kernel shares), then one full capture of a timed launch per workload
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02n_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02n_ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'pool_kernel' -s 3 -c 1 -f -o gpurun_out/r02n_prof_bench_c3 \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r02n_ncu_bench_c3.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'pool_le_kernel' -s 3 -c 1 -f -o gpurun_out/r02n_prof_bench_views \
    python bench.py --views --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r02n_ncu_bench_views.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'pool_kernel' -s 3 -c 1 -f -o gpurun_out/r02n_prof_bench_c5 \
    python bench.py --workload c5 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r02n_ncu_bench_c5.log 2>&1
ls -la gpurun_out/r02n*
