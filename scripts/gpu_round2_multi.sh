#!/bin/bash
# Round 2, multi-GPU call (one 8-GPU box): the workloads the scaling driver does not run -- C3 + 5 views, C5, C5 broadband --
# at 8 GPUs, then at 2 and 4 GPUs side by side on disjoint GPUs; the 2-GPU tests of the C-ABI exchange on the last two.
mkdir -p gpurun_out
tr() { # n port tag args...
  n=$1; port=$2; tag=$3; shift 3
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port \
      bench.py --gpus $n --no-cpu-baseline "$@" > gpurun_out/r02x_${tag}_n$n.log 2> gpurun_out/r02x_${tag}_n$n.err
  grep "^{" gpurun_out/r02x_${tag}_n$n.log | tail -1 | python -c "import sys, json; d = json.loads(sys.stdin.read()); print('$tag n=$n value %.4g e2e %.4g ms %.1f' % (d['value'], d['e2e']['value'], d['ms_per_step']))" || tail -3 gpurun_out/r02x_${tag}_n$n.err
}
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8; nproc
tr 8 29511 c3
tr 8 29512 views --views
tr 8 29513 c5 --workload c5 --photons 20000000
tr 8 29514 bb --workload broadband --steps 2 --warmup 1
( export CUDA_VISIBLE_DEVICES=0,1
  tr 2 29521 views --views; tr 2 29522 c5 --workload c5 --photons 20000000; tr 2 29523 bb --workload broadband --steps 2 --warmup 1 ) &
( export CUDA_VISIBLE_DEVICES=2,3,4,5
  tr 4 29531 views --views; tr 4 29532 c5 --workload c5 --photons 20000000; tr 4 29533 bb --workload broadband --steps 2 --warmup 1 ) &
( export CUDA_VISIBLE_DEVICES=6,7
  timeout 900 python -m pytest tests/test_cpp_host.py tests/test_gpu_api.py -q -m gpu > gpurun_out/r02x_pytest_2gpu.log 2>&1; tail -3 gpurun_out/r02x_pytest_2gpu.log ) &
wait
ls gpurun_out/r02x_*
