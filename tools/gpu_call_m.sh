#!/bin/bash
# A/B of build variants through the C ABI (tools/bench_kernels.py), two interleaved rounds
set -u
mkdir -p gpurun_out
: > gpurun_out/ab_kernels.jsonl
for round in 1 2; do
for v in in-tree "$@"; do
  if [ "$v" = in-tree ]; then timeout 200 python tools/bench_kernels.py 2>/dev/null | tee -a gpurun_out/ab_kernels.jsonl
  else CLDET_LIBRARY=build/variants/libcldet_$v.so timeout 200 python tools/bench_kernels.py 2>/dev/null | tee -a gpurun_out/ab_kernels.jsonl; fi
done; done
