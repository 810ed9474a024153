#!/bin/bash
# Round 2, the last seconds of GPU time: the thermal emission-radiance test of the first-interaction family
mkdir -p gpurun_out
timeout 33 python -m pytest tests/test_gpu_first_interaction.py -q -s -m gpu -k thermal > gpurun_out/r02final_thermal_fi.log 2>&1; echo "rc=$?" >> gpurun_out/r02final_thermal_fi.log
grep -v "^$" gpurun_out/r02final_thermal_fi.log | tail -45 | cut -c1-330
