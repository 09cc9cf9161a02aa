#!/bin/bash
set -u
mkdir -p gpurun_out
( time python bench.py > gpurun_out/bench_r2_n1.json 2> gpurun_out/bench_r2_n1.err ) 2>&1 | tail -3; echo "bench rc=$?"
( time python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_r2_ref.json 2> gpurun_out/bench_r2_ref.err ) 2>&1 | tail -3
tail -3 gpurun_out/bench_r2_n1.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2_n1.json'))
print({k:d[k] for k in ('value','ms_per_step','host_enqueue_us_per_step','gpu_launches')}, d['roofline']['kernel_ms'], d['roofline']['frac'], d['roofline']['step_frac'], d['e2e'])
print(d['clocks']); print(d['cpu_baseline']); print(d['gpu_eager_baseline'])
print(json.dumps({k:(round(v['ms_per_step'],4),round(v['value']),round(v['roofline_frac'],3)) for k,v in d['configs'].items()}))
dec=d['decode']
for k in ('all_anchors_candidates','trained_like'):
    print(k, dec[k]['ms_per_step'], dec[k]['stage_ms'], dec[k]['roofline']['frac'], dec[k].get('gpu_eager_baseline',{}).get('value'))
print(json.dumps(dec['predict_batch1_reference_mode'])[:2500])
print(json.dumps(dec['nms_vs_torchvision']))
r=json.load(open('gpurun_out/bench_r2_ref.json')); print(r['value'], r['cpu_baseline'])
PY
