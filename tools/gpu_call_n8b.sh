#!/bin/bash
set -u
N=${1:-8}
mkdir -p gpurun_out
ls oracle/_ref/retinanet 2>&1 | head -5
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29711"
$TR bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_r2_n$N.json 2> gpurun_out/bench_r2_n$N.err; echo "bench rc=$?"
$TR bench.py --gpus $N --steps 200 --warmup 10 --no-configs > gpurun_out/bench_r2_n${N}_200.json 2> gpurun_out/bench_r2_n${N}_200.err; echo "bench200 rc=$?"
python - <<PY
import json
for f in ('gpurun_out/bench_r2_n$N.json','gpurun_out/bench_r2_n${N}_200.json'):
    d=json.load(open(f))
    print({k:d[k] for k in ('value','ms_per_step','host_enqueue_us_per_step','n_gpus','steps')}, d['roofline']['kernel_ms'], d['e2e']['value'], d['config']['collective'][:20], d['clocks'])
    print(json.dumps(d.get('sharded_parity')))
    if 'configs' in d: print(json.dumps({k:(v['ms_per_step'],v['value'],v['roofline_frac']) for k,v in d['configs'].items()}))
PY
