#!/bin/bash
# usage: scripts/profile_round.sh TAG   -- one-GPU evidence for profiles/: plain runs first, then the ncu launch list
# of the bench command, one ncu --set full capture of the dominant kernel at profile size (source-level) and one at
# bench size (per-launch DRAM traffic for bench.py's roofline.traffic)
TAG=${1:-rXX}
mkdir -p gpurun_out
set -x
python scripts/profile_case.py --case c3 --photons 4000000 --batches 2 > gpurun_out/prof_plain_$TAG.log 2>&1 || exit 1
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_plain_$TAG.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:batch_kernel -s 1 -c 1 -f -o gpurun_out/prof_c3_$TAG \
    python scripts/profile_case.py --case c3 --photons 4000000 --batches 2 > gpurun_out/ncu_$TAG.log 2>&1
ncu --set full --clock-control none -k regex:batch_kernel -s 3 -c 1 -f -o gpurun_out/prof_bench_$TAG \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_$TAG.log 2>&1
ls -la gpurun_out/prof_c3_$TAG.ncu-rep gpurun_out/prof_bench_$TAG.ncu-rep
