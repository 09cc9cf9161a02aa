#!/bin/bash
# final validation pass (1 GPU), the driver's own commands: every GPU test, smoke, bench.py with the default flags
# (the reference arm -- unchanged code -- was last run by the previous version of this script: profiles/bench_r2_ref.json)
set -u
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -1
( time timeout 900 python bench.py > gpurun_out/bench_r2_n1.json 2> gpurun_out/bench_r2_n1.err ) 2>&1 | grep real; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2_n1.json'))
print({k:d[k] for k in ('value','ms_per_step','steps')}, d['roofline']['kernel_ms'], round(d['roofline']['frac'],4), d['e2e']['value'], d['clocks'])
print(d['cpu_baseline'], d.get('cpu_baseline_port',{}).get('value'))
for k,v in d['decode']['predict_batch1_reference_mode'].items():
    print(k, {q: (v[q] if not isinstance(v[q], dict) else v[q].get('ms_per_image', v[q])) for q in ('e2e_ms_per_image','device_resident_ms_per_image','gpu_eager_ms_per_image','reference_predict_same_gpu','reference_predict_cpu') if q in v})
print({k:(round(v['ms_per_step'],4),{a:round(b,4) for a,b in v['stage_ms'].items()},round(v['roofline']['frac'],3)) for k,v in d['decode'].items() if 'stage_ms' in v})
print({k:(round(v['cldet_ms'],4),round(v['torchvision_ms'],3),round(v['speedup'],2),v['identical_keep']) for k,v in d['decode']['nms_vs_torchvision'].items()})
print({k:(round(v['ms_per_step'],4),round(v['value'])) for k,v in d['configs'].items()})
PY
tail -2 gpurun_out/bench_r2_n1.err
