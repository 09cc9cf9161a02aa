#!/bin/bash
# final validation + evidence pass (1 GPU): every GPU test, smoke, bench.py with the default flags, launch lists and the
# ncu --set full capture of the post-filter chain (each after a plain run of the same command), detection side benches
set -u
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -1
( time timeout 900 python bench.py > gpurun_out/bench_r2_n1.json 2> gpurun_out/bench_r2_n1.err ) 2>&1 | grep real; echo "bench rc=$?"
D="python tools/bench_detect.py --mu -10.5 --steps 10 --warmup 3"
$D > gpurun_out/plain_det.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_detect_sparse_launches.csv $D > gpurun_out/ncu_d.log 2>&1
P="python tools/profile_predict.py --mu -9.5 --calls 3"
$P > gpurun_out/plain_pred.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_predict8k_launches.csv $P > gpurun_out/ncu_p.log 2>&1
C="python tools/bench_detect.py --mu -10.5 --steps 3 --warmup 3"
$C > gpurun_out/plain_chain.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:'rank_sort_kernel|nms_fused_kernel|select_fused_kernel' -s 9 -c 3 -o gpurun_out/r02_chain --force-overwrite $C > gpurun_out/ncu_chain.log 2>&1
ncu -i gpurun_out/r02_chain.ncu-rep --page raw --csv > gpurun_out/r02_chain_raw.csv 2>/dev/null
timeout 300 python tools/bench_detect.py --head > gpurun_out/detect_dense.json 2>/dev/null
timeout 300 python tools/bench_detect.py --mu -10.5 --head > gpurun_out/detect_sparse.json 2>/dev/null
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2_n1.json'))
print({k:d[k] for k in ('value','ms_per_step','steps')}, d['roofline']['kernel_ms'], round(d['roofline']['frac'],4), d['e2e']['value'], d['clocks'])
print(d['cpu_baseline'], d.get('cpu_baseline_port',{}).get('value'))
for k,v in d['decode']['predict_batch1_reference_mode'].items():
    print(k, {q: (v[q] if not isinstance(v[q], dict) else v[q].get('ms_per_image', v[q])) for q in ('device_resident_ms_per_image','gpu_eager_ms_per_image','reference_predict_same_gpu','reference_predict_cpu') if q in v})
print({k:(round(v['ms_per_step'],4),{a:round(b,4) for a,b in v['stage_ms'].items()},round(v['roofline']['frac'],3)) for k,v in d['decode'].items() if 'stage_ms' in v})
print(d['decode'].get('nms_vs_torchvision'))
for f in ('detect_dense','detect_sparse'):
    e=json.load(open('gpurun_out/%s.json'%f)); print(f, round(e['ms_per_step'],4), {a:round(b,4) for a,b in e['stage_ms'].items()}, round(e['roofline']['frac'],4), e.get('conv_layout'))
PY
tail -2 gpurun_out/bench_r2_n1.err
