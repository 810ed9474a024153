#!/bin/bash
# Round 2, GPU call T: staging cost of the compact structures (multi-block scan, merged read-back) and the broadband run with / without them
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pool.py tests/test_gpu_bounds.py tests/test_gpu_leap.py tests/test_broadband.py tests/test_assemble_optics.py -q -x -m gpu > gpurun_out/r02t_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02t_pytest.log
grep -E "passed|failed|^FAILED|^E  |rc=" gpurun_out/r02t_pytest.log | cut -c1-300 | head -20
for k in 1 0; do echo "== assemble timing, tuneExtMask=$k"; timeout 600 python scripts/assemble_timing.py $k 2>&1 | tail -9; done
for k in 1 0; do echo "== broadband, tuneExtMask=$k"; for rep in 1 2; do timeout 600 python scripts/broadband_bench.py --ext-mask $k 2>&1 | grep -E "spectral run|Error|Traceback"; done; done
timeout 900 python bench.py --workload broadband --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02t_bench_bb.log 2> gpurun_out/r02t_bench_bb.err; tail -c 300 gpurun_out/r02t_bench_bb.log
