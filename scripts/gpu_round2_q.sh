#!/bin/bash
# Round 2, GPU call Q: ncu capture of the layer-cropped variant on C5 next to the bitmap variant (same build)
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'pool_kernel' -s 1 -c 1 -f \
    -o gpurun_out/r02q_prof_c5_crop python scripts/profile_case.py --case c5 --photons 10000000 --batches 2 --ext-mask 2 --blocks-per-sm 6 --burst 44 > gpurun_out/r02q_ncu_crop.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'pool_kernel' -s 1 -c 1 -f \
    -o gpurun_out/r02q_prof_c5_bitmap python scripts/profile_case.py --case c5 --photons 10000000 --batches 2 --ext-mask 1 > gpurun_out/r02q_ncu_bitmap.log 2>&1
ls -la gpurun_out/r02q*
