"""The driver-facing contract of bench.py that can be checked without a GPU: the reference arm prints exactly one JSON
line on stdout with the keys the round-end comparison reads, and the product arm refuses to run without CUDA."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    env = dict(os.environ, CLDET_BENCH_BUDGET_S='4')
    r = subprocess.run([sys.executable, 'bench.py', '--impl', 'reference', '--steps', '1', '--warmup', '0'], cwd=ROOT, env=env,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['unit'] == 'images/s' and d['higher_is_better'] is True and d['value'] > 0
    snapshot = os.path.exists(os.path.join(ROOT, 'oracle', '_ref', 'retinanet', 'losses.bytecode'))
    assert d['cpu_baseline']['kind'] == ('reference' if snapshot else 'port')
    assert d['cpu_baseline']['cores'] >= 1 and d['cpu_baseline']['value'] == d['value']
    if snapshot:          # the numpy port is reported beside the unmodified reference
        assert d['cpu_baseline_port']['kind'] == 'port' and d['cpu_baseline_port']['value'] > 0
    assert d['e2e'] == {'value': d['value'], 'unit': 'images/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert d['gpu_launches'] == 0 and d['vs_baseline'] is None and 'workload' in d['config']


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK='1', WORLD_SIZE='2', LOCAL_RANK='1')
    r = subprocess.run([sys.executable, 'bench.py', '--impl', 'reference', '--gpus', '2', '--steps', '1', '--warmup', '0'], cwd=ROOT,
                       env=env, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ''


def test_product_arm_needs_cuda():
    import torch
    if torch.cuda.is_available():
        return
    r = subprocess.run([sys.executable, 'bench.py', '--steps', '1', '--warmup', '0'], cwd=ROOT, capture_output=True, text=True,
                       timeout=300)
    assert r.returncode != 0 and 'no CPU fallback' in (r.stderr + r.stdout)
