#!/bin/bash
# Round 2, GPU call Z: certification of the final tree -- whole GPU suite, smoke(), the default bench line and the reference arm
mkdir -p gpurun_out
timeout 2700 python -m pytest tests -q -m gpu > gpurun_out/r02z_pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/r02z_pytest_gpu.log
tail -6 gpurun_out/r02z_pytest_gpu.log | cut -c1-300
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02z_smoke.log 2>&1; tail -2 gpurun_out/r02z_smoke.log | cut -c1-300
timeout 900 python bench.py > gpurun_out/r02z_bench_c3.log 2> gpurun_out/r02z_bench_c3.err
timeout 900 python bench.py --views --no-cpu-baseline > gpurun_out/r02z_bench_views.log 2> gpurun_out/r02z_bench_views.err
for f in c3 views; do echo "== $f"; tail -c 300 gpurun_out/r02z_bench_$f.log; tail -3 gpurun_out/r02z_bench_$f.err; done
