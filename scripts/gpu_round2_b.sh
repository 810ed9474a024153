#!/bin/bash
# Round 2, GPU call B: pool + headline tests, extended gather probe, ncu --set full of the park and the pool kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pool.py -q > gpurun_out/r02b_pytest_pool.log 2>&1; echo "rc=$?" >> gpurun_out/r02b_pytest_pool.log
timeout 300 python scripts/gather_probe.py --full > gpurun_out/r02b_gather_probe.log 2>&1
for k in 1 2; do
  name=park; [ $k = 2 ] && name=pool
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:'batch_kernel|pool_kernel' -s 1 -c 1 -f \
      -o gpurun_out/r02b_prof_c3_$name python scripts/profile_case.py --case c3 --photons 16000000 --batches 2 --kernel $k \
      > gpurun_out/r02b_ncu_$name.log 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'batch_kernel|pool_kernel' -s 1 -c 1 -f \
    -o gpurun_out/r02b_prof_c5_pool python scripts/profile_case.py --case c5 --photons 8000000 --batches 2 --kernel 2 \
    > gpurun_out/r02b_ncu_c5_pool.log 2>&1
timeout 2400 python -m pytest tests -q -m gpu --deselect tests/test_gpu_pool.py > gpurun_out/r02b_pytest_gpu.log 2>&1
echo "rc=$?" >> gpurun_out/r02b_pytest_gpu.log
tail -4 gpurun_out/r02b_pytest_pool.log gpurun_out/r02b_pytest_gpu.log
ls -la gpurun_out/*.ncu-rep
