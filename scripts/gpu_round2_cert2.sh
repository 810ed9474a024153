#!/bin/bash
# Round 2, last GPU call: certification of the final tree -- the first-interaction tests at their final photon counts, the
# whole GPU suite, smoke() and the default bench line
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_first_interaction.py -q -s -m gpu > gpurun_out/r02cert_first_interaction.log 2>&1; echo "rc=$?" >> gpurun_out/r02cert_first_interaction.log
tail -3 gpurun_out/r02cert_first_interaction.log | cut -c1-300
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/r02cert_pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/r02cert_pytest_gpu.log
tail -6 gpurun_out/r02cert_pytest_gpu.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02cert_smoke.log 2>&1; tail -2 gpurun_out/r02cert_smoke.log | cut -c1-300
timeout 600 python bench.py > gpurun_out/r02cert_bench_c3.log 2> gpurun_out/r02cert_bench_c3.err
tail -c 400 gpurun_out/r02cert_bench_c3.log; tail -3 gpurun_out/r02cert_bench_c3.err
