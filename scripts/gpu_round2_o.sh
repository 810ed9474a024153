#!/bin/bash
# Round 2, GPU call O: column-compressed storage for fields too large for L2 (C5): tests, then A/B against the bitmap
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_pool.py tests/test_gpu_bounds.py tests/test_gpu_leap.py "tests/test_gpu_headline.py::test_column_compressed_storage_matches_reference_kernel_maps" -q -x > gpurun_out/r02o_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02o_pytest.log
grep -E "passed|failed|^FAILED|^E  |rc=" gpurun_out/r02o_pytest.log | cut -c1-300 | head -30
V=gpurun_out/r02o_variants.log; : > $V
run() { timeout 300 python scripts/profile_case.py "$@" 2>&1 | grep -E "BEST|Error|error|Traceback" >> $V; }
run --case c5 --photons 20000000 --batches 2 --ext-mask 1 --tag c5_bitmap
for occ in 6 7; do for burst in 8 44; do
  run --case c5 --photons 20000000 --batches 2 --ext-mask 2 --blocks-per-sm $occ --burst $burst --tag c5_columns
done; done
run --case c5 --photons 20000000 --batches 2 --ext-mask 2 --leap -1 --tag c5_columns_noleap
run --case c5 --photons 125000000 --batches 2 --ext-mask 1 --tag c5_bitmap_big
run --case c5 --photons 125000000 --batches 2 --ext-mask 2 --tag c5_columns_big
cat $V
