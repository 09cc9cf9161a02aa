#!/bin/bash
# ncu --set full of K4 with every anchor a candidate (block-aggregated append) and of the trained-like stream; raw pages exported on the box
set -u
mkdir -p gpurun_out
DD="python tools/bench_detect.py --mu -4 --steps 3 --warmup 3"
$DD > gpurun_out/plain_det_dense.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:'decode_filter_kernel' -s 3 -c 1 -o gpurun_out/r02_k4_dense --force-overwrite $DD > gpurun_out/ncu_k4d.log 2>&1
ncu -i gpurun_out/r02_k4_dense.ncu-rep --page raw --csv > gpurun_out/r02_k4_dense_raw.csv 2>/dev/null; rm -f gpurun_out/r02_k4_dense.ncu-rep
DS="python tools/bench_detect.py --mu -10.5 --steps 3 --warmup 3"
$DS > gpurun_out/plain_det_sparse.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:'decode_filter_kernel' -s 3 -c 1 -o gpurun_out/r02_k4_sparse --force-overwrite $DS > gpurun_out/ncu_k4s.log 2>&1
ncu -i gpurun_out/r02_k4_sparse.ncu-rep --page raw --csv > gpurun_out/r02_k4_sparse_raw.csv 2>/dev/null; rm -f gpurun_out/r02_k4_sparse.ncu-rep
tail -2 gpurun_out/ncu_k4d.log gpurun_out/ncu_k4s.log; ls -la gpurun_out/r02_k4_*_raw.csv
