#!/bin/bash
# Round 2, GPU call Z2: the full-size C5 parity test on the default path
mkdir -p gpurun_out
timeout 1500 python -m pytest "tests/test_gpu_headline.py::test_full_size_c5_default_path_matches_reference_kernel" -q --durations=3 > gpurun_out/r02z2_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02z2_pytest.log
tail -15 gpurun_out/r02z2_pytest.log | cut -c1-300
