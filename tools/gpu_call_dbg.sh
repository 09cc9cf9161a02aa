#!/bin/bash
# scratch GPU call: long-list resolve with the parallel-round chain + the wider prepare kernel
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_detect_gpu.py tests/test_guards_gpu.py tests/test_robustness_gpu.py tests/test_reference_snapshot.py -m gpu -x -q 2>&1 | tail -4 | cut -c1-400
timeout 600 python bench.py --steps 50 --warmup 5 --no-configs --no-cpu-baseline > gpurun_out/bench_short.json 2>/dev/null; python -c "
import json;d=json.load(open('gpurun_out/bench_short.json'))
for k in ('all_anchors_candidates','trained_like'):
    e=d['decode'][k]; print('bench.py',k,round(e['ms_per_step'],4),{a:round(b,4) for a,b in e['stage_ms'].items()},round(e['roofline']['frac'],3))
print({k:round(v['device_resident_ms_per_image'],3) for k,v in d['decode']['predict_batch1_reference_mode'].items()})
print({k:(round(v['cldet_ms'],4),round(v['speedup'],2),v['identical_keep']) for k,v in d['decode']['nms_vs_torchvision'].items()})
"
P="python tools/profile_predict.py --mu -9.5 --calls 3"
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_predict8k_launches.csv $P > gpurun_out/ncu_p.log 2>&1
python - <<'P'
import csv,collections
rows=[r for r in csv.reader(open('gpurun_out/r02_predict8k_launches.csv')) if len(r)>10]
hdr=rows[0]; k=hdr.index('Kernel Name'); v=hdr.index('Metric Value')
agg=collections.defaultdict(list)
for r in rows[1:]:
    try: agg[r[k].split('(')[0]].append(float(r[v].replace(',','')))
    except: pass
for n,x in agg.items():
    if 'cldet' in n: print(n[:50].ljust(50), len(x), round(sum(x)/len(x)/1000,2),'us')
P
