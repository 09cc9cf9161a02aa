#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 300 python - <<'PY' 2>&1 | tail -12
import json, sys, torch
sys.path.insert(0, '.')
from tools.bench_detect import measure_nms_h2h, measure_predict
dev = torch.device('cuda', 0)
for k in (1000, 8000, 40000):
    print(json.dumps(measure_nms_h2h(dev, k)))
for mu in (-10.5, -9.5, -8.5):
    r = measure_predict(dev, mu, cpu_images=0)
    print(json.dumps({k: r[k] for k in ('candidates_per_image', 'kept_per_image', 'e2e_ms_per_image', 'device_resident_ms_per_image', 'gpu_eager_ms_per_image')}))
PY
timeout 120 python tools/bench_detect.py --mu -10.5 > gpurun_out/det_default.json 2>/dev/null; python -c "import json;d=json.load(open('gpurun_out/det_default.json'));print('default(smem) ',d['ms_per_step'],d['stage_ms'])"
CLDET_NMS_RESOLVE=stream timeout 120 python tools/bench_detect.py --mu -10.5 > gpurun_out/det_stream.json 2>/dev/null; python -c "import json;d=json.load(open('gpurun_out/det_stream.json'));print('stream        ',d['ms_per_step'],d['stage_ms'])"
D="python tools/bench_detect.py --mu -4.0 --steps 3 --warmup 3"
timeout 200 $D > gpurun_out/plain_dense.log 2>&1 && \
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:'decode_filter_kernel' -s 3 -c 1 -o gpurun_out/r02_k4_dense --force-overwrite $D > gpurun_out/ncu_k4.log 2>&1
ls -la gpurun_out/r02_k4_dense.ncu-rep
