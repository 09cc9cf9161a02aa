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
for mu in -10.5 -4.0; do
timeout 120 python tools/bench_detect.py --mu $mu --head > gpurun_out/det_$mu.json 2>/dev/null; python -c "import json;d=json.load(open('gpurun_out/det_$mu.json'));print('mu=$mu',d['ms_per_step'],d['stage_ms'],d['roofline']['frac'],d.get('conv_layout'))"
done
# racecheck once per round on the smallest case that runs every kernel family (smoke: loss fwd+bwd, assign, decode, sort, NMS)
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python __graft_entry__.py smoke > gpurun_out/r02_racecheck_smoke.log 2>&1; echo "racecheck rc=$?"; tail -6 gpurun_out/r02_racecheck_smoke.log
