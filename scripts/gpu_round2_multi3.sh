#!/bin/bash
# Round 2, third multi-GPU call (one 4-GPU box), final build: C3 + views and broadband at 4 and 2 GPUs
mkdir -p gpurun_out
tr() { n=$1; port=$2; tag=$3; shift 3
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port \
      bench.py --gpus $n --no-cpu-baseline "$@" > gpurun_out/r02w_${tag}_n$n.log 2> gpurun_out/r02w_${tag}_n$n.err
  grep "^{" gpurun_out/r02w_${tag}_n$n.log | tail -1 | python -c "import sys, json; d = json.loads(sys.stdin.read()); print('$tag n=$n value %.4g e2e %.4g ms %.1f' % (d['value'], d['e2e']['value'], d['ms_per_step']))" || tail -3 gpurun_out/r02w_${tag}_n$n.err
}
tr 4 29711 views --views --steps 3
tr 4 29712 bb --workload broadband --steps 2 --warmup 1
( export CUDA_VISIBLE_DEVICES=0,1; tr 2 29721 views --views --steps 3 ) &
( export CUDA_VISIBLE_DEVICES=2,3; tr 2 29731 bb --workload broadband --steps 2 --warmup 1 ) &
wait
