#!/bin/bash
# tuning sweep of the fast kernel on one B200 (burst length x blocks/SM x park threshold)
mkdir -p gpurun_out
python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu_v5.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu_v5.log
for burst in 4 8; do for occ in 4 5 6; do for park in 12 16 20; do
  echo "burst $burst occ $occ park $park"
  MCB_BURST=$burst MCB_BLOCKS_PER_SM=$occ MCB_PARK_THRESHOLD=$park python scripts/profile_case.py --case c3 --photons 8000000 --batches 2 | grep "batch 1"
done; done; done
for c in c1 c2 c3mie c5; do echo "case $c"; python scripts/profile_case.py --case $c --photons 8000000 --batches 2 | grep "batch 1"; done
echo "c2 views"; python scripts/profile_case.py --case c2 --views --photons 2000000 --batches 2 | grep "batch 1"
echo "c3 views"; python scripts/profile_case.py --case c3 --views --photons 1000000 --batches 2 | grep "batch 1"
