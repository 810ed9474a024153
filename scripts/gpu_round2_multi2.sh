#!/bin/bash
# Round 2, second multi-GPU call (one 8-GPU box), final build: C3, C3 + views, C5 (default step), broadband at 8 GPUs; C5 and views at 2 and 4
mkdir -p gpurun_out
tr() { n=$1; port=$2; tag=$3; shift 3
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port \
      bench.py --gpus $n --no-cpu-baseline "$@" > gpurun_out/r02y_${tag}_n$n.log 2> gpurun_out/r02y_${tag}_n$n.err
  grep "^{" gpurun_out/r02y_${tag}_n$n.log | tail -1 | python -c "import sys, json; d = json.loads(sys.stdin.read()); print('$tag n=$n value %.4g e2e %.4g ms %.1f' % (d['value'], d['e2e']['value'], d['ms_per_step']))" || tail -3 gpurun_out/r02y_${tag}_n$n.err
}
tr 8 29611 c3
tr 8 29612 views --views
tr 8 29613 c5 --workload c5
tr 8 29614 bb --workload broadband --steps 2 --warmup 1
( export CUDA_VISIBLE_DEVICES=0,1; tr 2 29621 c5 --workload c5; tr 2 29622 c3 ) &
( export CUDA_VISIBLE_DEVICES=2,3,4,5; tr 4 29631 c5 --workload c5; tr 4 29632 c3 ) &
wait
ls gpurun_out/r02y_*
