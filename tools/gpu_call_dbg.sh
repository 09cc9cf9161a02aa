#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_detect_gpu.py tests/test_guards_gpu.py tests/test_robustness_gpu.py -m gpu -x -q 2>&1 | tail -4 | cut -c1-300
for mu in -10.5 -4; do timeout 300 python tools/bench_detect.py --mu $mu --head > gpurun_out/det_$mu.json 2>/dev/null; python -c "import json;d=json.load(open('gpurun_out/det_$mu.json'));print('mu=$mu',d['ms_per_step'],d['stage_ms'],round(d['roofline']['frac'],4),d['conv_layout']['head_layout_ms'])"; done
