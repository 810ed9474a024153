#!/bin/bash
# Round 2, GPU call Y: the pool local-estimation kernel on grids narrower than the ghost shell (C2 step cloud + 5 views)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pool.py tests/test_gpu_bounds.py -q -x > gpurun_out/r02y2_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02y2_pytest.log
grep -E "passed|failed|^FAILED|^E  |rc=" gpurun_out/r02y2_pytest.log | cut -c1-300 | head -20
V=gpurun_out/r02y2_variants.log; : > $V
run() { timeout 300 python scripts/profile_case.py "$@" 2>&1 | grep -E "BEST|Error|error|Traceback" >> $V; }
run --case c2 --photons 16000000 --views --batches 2 --kernel 1 --tag c2_views_queue
for occ in 4 5 6; do run --case c2 --photons 16000000 --views --batches 2 --kernel 2 --blocks-per-sm $occ --tag c2_views_pool; done
run --case c2 --photons 64000000 --views --batches 2 --kernel 1 --tag c2_views_queue_big
run --case c2 --photons 64000000 --views --batches 2 --kernel 2 --tag c2_views_pool_big
cat $V
