#!/bin/bash
# ncu --set full of the long-list resolve (predict, one image, ~8 k candidates); raw page exported on the box
set -u
mkdir -p gpurun_out
P="python tools/profile_predict.py --mu -9.5 --calls 3"
$P > gpurun_out/plain_pred.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:'nms_resolve_stream_kernel|nms_mask_kernel|radix_sort_kernel' -s 3 -c 3 -o gpurun_out/r02_long --force-overwrite $P > gpurun_out/ncu_long.log 2>&1
ncu -i gpurun_out/r02_long.ncu-rep --page raw --csv > gpurun_out/r02_long_raw.csv 2>/dev/null
tail -2 gpurun_out/ncu_long.log; ls -la gpurun_out/r02_long_raw.csv gpurun_out/r02_long.ncu-rep
