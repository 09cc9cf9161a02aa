#!/bin/bash
# scratch GPU call: detection tests + A/B of the short-list kernels (one-launch NMS, bitonic sort)
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_detect_gpu.py tests/test_guards_gpu.py tests/test_robustness_gpu.py tests/test_reference_snapshot.py -m gpu -x -q 2>&1 | tail -15 | cut -c1-400
show() { python -c "import json,sys;d=json.load(open('$1'));print('$2',round(d['ms_per_step'],4),{k:round(v,4) for k,v in d['stage_ms'].items()},round(d['roofline']['frac'],4))"; }
for cfg in "1 1" "1 0" "1 1" "1 0"; do set -- $cfg
  CLDET_NMS_FUSED=$1 CLDET_BUCKET_RANK=$2 timeout 300 python tools/bench_detect.py --mu -10.5 > gpurun_out/det_f$1_b$2.json 2>/dev/null; show gpurun_out/det_f$1_b$2.json "mu=-10.5 fused=$1 bucket=$2"
done
timeout 300 python tools/bench_detect.py --mu -4 > gpurun_out/det_dense.json 2>/dev/null; show gpurun_out/det_dense.json "mu=-4 default"
timeout 600 python bench.py --steps 50 --warmup 5 --no-configs --no-cpu-baseline > gpurun_out/bench_short.json 2>/dev/null; python -c "
import json;d=json.load(open('gpurun_out/bench_short.json'))
for k in ('all_anchors_candidates','trained_like'):
    e=d['decode'][k]; print('bench.py',k,round(e['ms_per_step'],4),{a:round(b,4) for a,b in e['stage_ms'].items()},round(e['roofline']['frac'],3))
print({k:round(v['device_resident_ms_per_image'],3) for k,v in d['decode']['predict_batch1_reference_mode'].items()})
print(d['decode'].get('nms_vs_torchvision'))
"
