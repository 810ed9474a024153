#!/bin/bash
# tuning sweep of the throughput kernels on one B200: blocks/SM x park threshold for the flux kernel, the local-
# estimation knobs, every case of scripts/profile_case.py (run under gpurun; each line is one measured batch)
mkdir -p gpurun_out
for occ in 6 8; do for park in 12 14 16 18 20; do
  echo "flux: occ $occ park $park"
  MCB_BLOCKS_PER_SM=$occ MCB_PARK_THRESHOLD=$park python scripts/profile_case.py --case c3 --photons 16000000 --batches 3 | grep "batch 2"
done; done
echo "flux, x-fastest field"; MCB_LAYOUT=linear python scripts/profile_case.py --case c3 --photons 16000000 --batches 3 | grep "batch 2"
for carry in -1 0 12 24; do
  echo "views: carry $carry"; MCB_LE_CARRY=$carry python scripts/profile_case.py --case c3 --views --photons 2000000 --batches 3 | grep "batch 2"
done
for c in c1 c2 c3mie c5; do echo "case $c"; python scripts/profile_case.py --case $c --photons 16000000 --batches 3 | grep "batch 2"; done
echo "c2 views"; python scripts/profile_case.py --case c2 --views --photons 2000000 --batches 3 | grep "batch 2"
echo "stretched grid"; python scripts/profile_irregular.py stretched | grep "batch 2"
echo "thermal source"; python scripts/profile_lw.py | grep "batch 2"
