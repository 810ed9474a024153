#!/bin/bash
# Round 2, GPU call E: local estimation on the pool organisation
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_pool.py tests/test_gpu_bounds.py tests/test_cpp_host.py "tests/test_gpu_trace.py::test_trace_parity" -q -x > gpurun_out/r02e_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02e_pytest.log
tail -25 gpurun_out/r02e_pytest.log | cut -c1-250
V=gpurun_out/r02e_variants.log; : > $V
run() { timeout 400 python scripts/profile_case.py "$@" 2>&1 | grep -E "BEST|Error|error|Traceback" >> $V; }
run --case c3 --photons 16000000 --views --batches 2 --kernel 1 --tag le_queue
for occ in 4 5 6; do for lay in 1 2; do
  run --case c3 --photons 16000000 --views --batches 2 --kernel 2 --blocks-per-sm $occ --layout $lay --tag le_pool
done; done
run --case c3mie --photons 16000000 --views --batches 2 --kernel 1 --tag le_queue
run --case c3mie --photons 16000000 --views --batches 2 --kernel 2 --blocks-per-sm 5 --tag le_pool
run --case c5 --photons 4000000 --views --batches 2 --kernel 1 --tag le_queue
run --case c5 --photons 4000000 --views --batches 2 --kernel 2 --blocks-per-sm 5 --tag le_pool
run --case c3 --photons 125000000 --kernel 0 --batches 2 --tag flux_default
cat $V
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'pool_le_kernel' -s 1 -c 1 -f \
    -o gpurun_out/r02e_prof_c3_views_pool python scripts/profile_case.py --case c3 --photons 4000000 --views --batches 2 --kernel 2 \
    > gpurun_out/r02e_ncu_le.log 2>&1
ls -la gpurun_out/r02e*
