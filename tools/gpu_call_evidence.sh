#!/bin/bash
# Round-2 evidence pass (1 GPU): tests, smoke, both bench arms, launch lists, ncu --set full captures (each after a plain run of
# the same command), side benches.  Everything lands in gpurun_out/; tools/summarise_profiles.py turns it into profiles/.
set -u
mkdir -p gpurun_out
ls oracle/_ref/retinanet | head -3
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 900 python bench.py --steps 200 --warmup 20 > gpurun_out/bench_r2_n1.json 2> gpurun_out/bench_r2_n1.err; echo "bench rc=$?"
timeout 900 python bench.py --impl reference --steps 20 --warmup 2 > gpurun_out/bench_r2_ref.json 2> gpurun_out/bench_r2_ref.err; echo "ref rc=$?"
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-decode --no-configs"
$B > gpurun_out/plain_loss.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_loss_launches.csv $B > gpurun_out/ncu_l.log 2>&1
B2="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-decode --no-configs"
$B2 > gpurun_out/plain_loss2.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:'focal_loss_kernel|gt_scatter_kernel' -c 2 -o gpurun_out/r02_loss --force-overwrite $B2 > gpurun_out/ncu_f.log 2>&1
ncu -i gpurun_out/r02_loss.ncu-rep --page raw --csv > gpurun_out/r02_loss_raw.csv 2>/dev/null; rm -f gpurun_out/r02_loss.ncu-rep
H="python tools/bench_kernels.py --steps 3 --only head_probs"
$H > gpurun_out/plain_k3h.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:'focal_loss_head_kernel' -s 3 -c 1 -o gpurun_out/r02_k3h --force-overwrite $H > gpurun_out/ncu_k3h.log 2>&1
ncu -i gpurun_out/r02_k3h.ncu-rep --page raw --csv > gpurun_out/r02_k3h_raw.csv 2>/dev/null; rm -f gpurun_out/r02_k3h.ncu-rep
HL="python tools/bench_kernels.py --steps 3 --only cat_logits"
$HL > gpurun_out/plain_k3l.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:'focal_loss_kernel' -s 3 -c 1 -o gpurun_out/r02_k3_logits --force-overwrite $HL > gpurun_out/ncu_k3l.log 2>&1
ncu -i gpurun_out/r02_k3_logits.ncu-rep --page raw --csv > gpurun_out/r02_k3_logits_raw.csv 2>/dev/null; rm -f gpurun_out/r02_k3_logits.ncu-rep
DD="python tools/bench_detect.py --mu -4 --steps 3 --warmup 3"
$DD > gpurun_out/plain_det_dense.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:'decode_filter_kernel' -s 3 -c 1 -o gpurun_out/r02_k4_dense --force-overwrite $DD > gpurun_out/ncu_k4d.log 2>&1
ncu -i gpurun_out/r02_k4_dense.ncu-rep --page raw --csv > gpurun_out/r02_k4_dense_raw.csv 2>/dev/null; rm -f gpurun_out/r02_k4_dense.ncu-rep
D="python tools/bench_detect.py --mu -10.5 --steps 10 --warmup 3"
$D > gpurun_out/plain_det.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_detect_sparse_launches.csv $D > gpurun_out/ncu_d.log 2>&1
P="python tools/profile_predict.py --mu -9.5 --calls 3"
$P > gpurun_out/plain_pred.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_predict8k_launches.csv $P > gpurun_out/ncu_p.log 2>&1
timeout 200 python tools/bench_kernels.py > gpurun_out/bench_kernels.json 2>/dev/null
timeout 300 python tools/bench_head_layout.py > gpurun_out/bench_head_layout.json 2> gpurun_out/bhl.err
timeout 300 python tools/bench_head_layout.py --logits >> gpurun_out/bench_head_layout.json 2>> gpurun_out/bhl.err
timeout 300 python tools/bench_detect.py --head > gpurun_out/detect_dense.json 2>/dev/null
timeout 300 python tools/bench_detect.py --mu -10.5 --head > gpurun_out/detect_sparse.json 2>/dev/null
timeout 300 python tools/bench_api.py > gpurun_out/bench_api.jsonl 2>/dev/null
timeout 300 python tools/bench_host.py > gpurun_out/bench_host.json 2>/dev/null
timeout 300 python tools/bench_logits.py > gpurun_out/bench_logits.json 2>/dev/null
timeout 300 python tools/bench_distill.py > gpurun_out/bench_distill.json 2>/dev/null
cut -c1-300 gpurun_out/bench_r2_n1.json; echo; cut -c1-700 gpurun_out/bench_r2_ref.json; echo; cat gpurun_out/bench_kernels.json
for f in detect_dense detect_sparse; do python -c "import json;d=json.load(open('gpurun_out/$f.json'));print('$f',d['ms_per_step'],d['stage_ms'],round(d['roofline']['frac'],4),d.get('conv_layout'))"; done
du -sh gpurun_out; ls gpurun_out | head -60
