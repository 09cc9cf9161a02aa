#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | cut -c1-300
( time timeout 900 python bench.py --impl reference > gpurun_out/bench_r2_ref.json 2> gpurun_out/bench_r2_ref.err ) 2>&1 | grep real
cut -c1-1100 gpurun_out/bench_r2_ref.json
