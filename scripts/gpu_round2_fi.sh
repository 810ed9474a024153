#!/bin/bash
# Round 2, GPU call FI: the first-interaction tests (kernels against the independent deterministic 3-D answers)
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_first_interaction.py -q -s -m gpu --durations=5 > gpurun_out/r02fi_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02fi_pytest.log
grep -v "^$" gpurun_out/r02fi_pytest.log | tail -60 | cut -c1-250
