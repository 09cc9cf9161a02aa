#!/bin/bash
# round-2 GPU pass A (1 GPU): tests, smoke, default bench line, API host-overhead bench
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python __graft_entry__.py smoke 2>&1 | tail -2
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err; echo "bench rc=$?"
tail -5 gpurun_out/bench_a.err
python tools/bench_api.py > gpurun_out/bench_api.jsonl 2> gpurun_out/bench_api.err; echo "api rc=$?"
cat gpurun_out/bench_api.jsonl; tail -3 gpurun_out/bench_api.err
cut -c1-1500 gpurun_out/bench_a.json
