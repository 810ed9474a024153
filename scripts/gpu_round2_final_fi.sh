#!/bin/bash
# Round 2: the first-interaction tests' printed statistics on the final build, and the seam diagnostic (no odd entries expected)
mkdir -p gpurun_out
timeout 50 python -m pytest tests/test_gpu_first_interaction.py -q -s -m gpu > gpurun_out/r02final_first_interaction.log 2>&1; echo "rc=$?" >> gpurun_out/r02final_first_interaction.log
tail -2 gpurun_out/r02final_first_interaction.log
timeout 30 python scripts/diag_stretched_le.py default > gpurun_out/r02final_diag.log 2>&1; cut -c1-400 gpurun_out/r02final_diag.log | tail -4
