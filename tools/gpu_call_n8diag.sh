#!/bin/bash
set -u
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29711"
B="bench.py --gpus $N --steps 20 --warmup 5 --quick"
CLDET_BENCH_TRACE=base $TR $B 2>/dev/null | tail -1
CLDET_BENCH_TRACE=nosampler $TR $B --no-clock-sampler 2>/dev/null | tail -1
CLDET_BENCH_TRACE=noevents $TR $B --no-clock-sampler --no-kernel-events 2>/dev/null | tail -1
CLDET_BENCH_TRACE=nccl $TR $B --no-clock-sampler --collective nccl 2>/dev/null | tail -1
CLDET_BENCH_TRACE=sepwait CLDET_PEER_SEPARATE_WAIT=1 $TR $B --no-clock-sampler 2>/dev/null | tail -1
CLDET_BENCH_TRACE=base200 $TR bench.py --gpus $N --steps 200 --warmup 10 --quick 2>/dev/null | tail -1
python - <<'PY'
import json, glob
for tag in ('base','nosampler','noevents','nccl','sepwait','base200'):
    files = sorted(glob.glob('gpurun_out/trace_%s_rank*.json' % tag))
    if not files: continue
    print('==', tag)
    for f in files[:8]:
        d = json.load(open(f))
        dev = d['device_ms_since_t0']; host = d['host_ms_since_first']
        dd = [round(dev[0],3)] + [round(dev[i]-dev[i-1],3) for i in range(1,len(dev))]
        hh = [round(host[i]-host[i-1],3) for i in range(1,len(host))]
        big = [(i,x) for i,x in enumerate(dd) if x > 0.6]
        print(f[-11:-5], 'dev total %.2f' % dev[-1], 'steps>0.6ms:', big[:8], '| host gaps>0.6ms:', [(i,x) for i,x in enumerate(hh) if x > 0.6][:8])
PY
