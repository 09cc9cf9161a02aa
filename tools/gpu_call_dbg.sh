#!/bin/bash
# scratch GPU call: private histogram copies in the top-k select
set -u
mkdir -p gpurun_out
show() { python -c "import json,sys;d=json.load(open('$1'));print('$2',round(d['ms_per_step'],4),{k:round(v,4) for k,v in d['stage_ms'].items()},round(d['roofline']['frac'],4))"; }
for i in 1 2; do
timeout 300 python tools/bench_detect.py --mu -4 > gpurun_out/det_dense_sub.json 2>/dev/null; show gpurun_out/det_dense_sub.json "mu=-4"
timeout 300 python tools/bench_detect.py --mu -10.5 > gpurun_out/det_sparse_sub.json 2>/dev/null; show gpurun_out/det_sparse_sub.json "mu=-10.5"
done
timeout 300 python tools/bench_detect.py --mu -7.5 > gpurun_out/det_mid_sub.json 2>/dev/null; show gpurun_out/det_mid_sub.json "mu=-7.5"
timeout 900 python -m pytest tests/test_detect_gpu.py tests/test_guards_gpu.py tests/test_robustness_gpu.py -m gpu -x -q 2>&1 | tail -3 | cut -c1-300
