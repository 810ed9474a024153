#!/bin/bash
# Round 2, GPU call AA: tests after the staging clean-up; ncu captures of the final kernels (C3 flux, C3 + views, C5)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pool.py tests/test_gpu_bounds.py tests/test_gpu_leap.py "tests/test_gpu_headline.py::test_full_size_c5_default_path_matches_reference_kernel" -q -x > gpurun_out/r02aa_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02aa_pytest.log
grep -E "passed|failed|^FAILED|^E  |rc=" gpurun_out/r02aa_pytest.log | cut -c1-300 | head
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'pool_kernel' -s 1 -c 1 -f \
    -o gpurun_out/r02aa_prof_c3 python scripts/profile_case.py --case c3 --photons 16000000 --batches 2 > gpurun_out/r02aa_ncu_c3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'pool_le_kernel' -s 1 -c 1 -f \
    -o gpurun_out/r02aa_prof_views python scripts/profile_case.py --case c3 --photons 4000000 --views --batches 2 > gpurun_out/r02aa_ncu_views.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'pool_kernel' -s 1 -c 1 -f \
    -o gpurun_out/r02aa_prof_c5 python scripts/profile_case.py --case c5 --photons 10000000 --batches 2 > gpurun_out/r02aa_ncu_c5.log 2>&1
ls -la gpurun_out/r02aa*
