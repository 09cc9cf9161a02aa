#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python tools/bench_host.py 2>/dev/null | tee gpurun_out/bench_host.json
timeout 400 python tools/probe_host.py > gpurun_out/probe_host.log 2>&1; echo "probe rc=$?"
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python __graft_entry__.py smoke > gpurun_out/r02_racecheck_smoke.log 2>&1; echo "racecheck rc=$?"; tail -6 gpurun_out/r02_racecheck_smoke.log
