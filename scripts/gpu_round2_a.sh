#!/bin/bash
# Round 2, GPU call A: compiler probe, the full GPU test suite, kernel variants (park vs pool), the gather ceiling.
mkdir -p gpurun_out
bash scripts/gpu_probe_compilers.sh gpurun_out/r02_compiler_probe.log > /dev/null 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/r02a_smi.log 2>&1
# (1) pool-vs-park parity first: it gates everything else
timeout 900 python -m pytest tests/test_gpu_pool.py -x -q > gpurun_out/r02a_pytest_pool.log 2>&1
echo "pool tests rc=$?" >> gpurun_out/r02a_pytest_pool.log
# (2) gather ceiling
timeout 300 python scripts/gather_probe.py > gpurun_out/r02a_gather_probe.log 2>&1
# (3) kernel variants
V=gpurun_out/r02a_variants.log; : > $V
run() { timeout 300 python scripts/profile_case.py "$@" 2>&1 | grep -E "BEST|Error|error|Traceback" >> $V; }
for case in c3 c3mie c5 c1 c4; do
  ph=16000000; [ $case = c5 ] && ph=8000000
  run --case $case --photons $ph --kernel 1 --tag park
  run --case $case --photons $ph --kernel 1 --blocks-per-sm 6 --tag park6
  for occ in 6 8; do for burst in 8 4; do
    run --case $case --photons $ph --kernel 2 --blocks-per-sm $occ --burst $burst --tag pool
  done; done
done
run --case c3 --photons 16000000 --kernel 2 --blocks-per-sm 5 --burst 8 --tag pool_occ5
run --case c3 --photons 16000000 --kernel 2 --blocks-per-sm 4 --burst 8 --tag pool_occ4
run --case c3 --photons 16000000 --kernel 2 --layout 1 --blocks-per-sm 6 --tag pool_linear
run --case c3 --photons 125000000 --kernel 1 --batches 2 --tag park_full
run --case c3 --photons 125000000 --kernel 2 --blocks-per-sm 6 --batches 2 --tag pool_full6
run --case c3 --photons 125000000 --kernel 2 --blocks-per-sm 8 --batches 2 --tag pool_full8
cat $V
# (4) the rest of the GPU suite
timeout 2400 python -m pytest tests -x -q -m gpu --deselect tests/test_gpu_pool.py > gpurun_out/r02a_pytest_gpu.log 2>&1
echo "gpu tests rc=$?" >> gpurun_out/r02a_pytest_gpu.log
tail -5 gpurun_out/r02a_pytest_pool.log gpurun_out/r02a_pytest_gpu.log
cat gpurun_out/r02a_gather_probe.log
