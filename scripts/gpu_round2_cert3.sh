#!/bin/bash
# Round 2, last GPU call: the whole GPU suite (with the seam fix of the edge-table grids) and smoke() on the final tree
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu > gpurun_out/r02cert3_pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/r02cert3_pytest_gpu.log
tail -6 gpurun_out/r02cert3_pytest_gpu.log | cut -c1-300
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02cert3_smoke.log 2>&1; tail -2 gpurun_out/r02cert3_smoke.log | cut -c1-300
