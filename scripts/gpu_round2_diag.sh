#!/bin/bash
mkdir -p gpurun_out
( timeout 150 python scripts/diag_stretched_le.py default n1e6 roulette nadir down slant default:irregular
  MCB_LIB_DEBUG=1 timeout 90 python scripts/diag_stretched_le.py dbg ) > gpurun_out/r02diag.log 2>&1
cut -c1-700 gpurun_out/r02diag.log | tail -40
