#!/usr/bin/env python
"""Benchmark of the detection-head hot path (BASELINE.json metric: detection-head images/sec + % HBM roofline).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--no-cpu-baseline] [--no-decode]

Default workload = BASELINE config 3: COCO-shaped focal loss fwd+bwd, 16 x 800x1333, C=80, A=200700 per GPU
(weak scaling: every rank processes its own 16 images; one NCCL all-gather of the per-image loss terms per step).
A step = anchors cached -> IoU/assign -> fused focal + smooth-L1 loss AND gradients -> backward weight check.

value : device-resident throughput through the C ABI (inputs already in HBM), CUDA-event timed, max over ranks.
e2e   : the same metric through the public Python drop-in (FocalLoss.forward + autograd) with HOST inputs: every step
        copies cls/reg/annotations from pinned host memory and reads the loss back.
roofline: the fused loss kernel's algorithmic bytes / its own CUDA-event time inside the timed loop.
cpu_baseline: the UNMODIFIED reference (byte-code snapshot oracle/_ref, built by oracle/build_ref.py where /root/reference exists)
        timed on this box's host cores on a bounded sample, rank 0 only (kind "reference"); the numpy oracle port beside it as
        cpu_baseline_port.  Without the snapshot the port is the baseline (kind "port").
gpu_eager_baseline (N=1, baseline leg): the reference's own execution model (eager torch ops + autograd, oracle/torch_eager.py)
        on the same GPU and batch -- informative, like cpu_baseline.
decode (N=1): BASELINE config 4, the other half of the metric: decode + threshold + top-1000 + per-class NMS over 32 images.
--impl reference: times that CPU baseline only (the snapshot when present, else the port; /root/reference is not on the box).
"""
import argparse
import ctypes
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

H, W, C, N_PER_GPU, GMAX = 800, 1333, 80, 16, 20
NOMINAL_HBM_GBS = 8000.0      # "also report against nominal 8 TB/s"
CPU_THREADS_CAP = 16          # both arms time the CPU port with min(host cores, 16) threads
# the SAME string in both arms (the driver compares the two lines' config.workload)
WORKLOAD = 'coco_loss_fwd_bwd: 16 x 800x1333 per GPU, C=80, A=200700, G<=20 (BASELINE config 3)'
METRIC = 'detection-head images/sec (loss fwd+bwd: IoU assign + focal + smooth-L1, grads for cls and reg)'


def synth_annotations(rng, n, gmax, h, w, c, empty=(0,), exact=False, pseudo_split=None):
    """SURVEY 8(d) GT generator.  exact: every image has exactly gmax boxes (config 5).  pseudo_split = P: the first rows of
    an image are new-class GT (label >= P), the remaining rows pseudo-labels of old classes (label < P) -- the layout the
    dataset produces when it appends the Labeler's boxes (dataloader.py:129-136); padding rows are -1."""
    ann = np.full((n, gmax, 5), -1.0, np.float32)
    for j in range(n):
        if j in empty:
            continue
        g = gmax if exact else int(rng.integers(1, gmax + 1))
        x1 = rng.uniform(0, 0.7 * w, g)
        y1 = rng.uniform(0, 0.7 * h, g)
        bw = rng.uniform(16, 0.3 * w + 16, g)
        bh = rng.uniform(16, 0.3 * h + 16, g)
        if pseudo_split:
            k = max(1, g // 2)
            lab = np.concatenate([rng.integers(pseudo_split, c, k), rng.integers(0, pseudo_split, g - k)])
        else:
            lab = rng.integers(0, c, g)
        ann[j, :g] = np.stack([x1, y1, x1 + bw, y1 + bh, lab.astype(np.float64)], 1)
    return ann


def peak_hbm():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        try:
            return float(json.load(open(p))['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
        except Exception:
            pass
    return 6650.0, 'fallback (B200_PROFILING.md)'


def ncu_traffic(kernel):
    p = os.path.join(ROOT, 'profiles', 'traffic.json')
    if os.path.exists(p):
        try:
            return json.load(open(p)).get(kernel)
        except Exception:
            return None
    return None


class ClockSampler:
    """Polls NVML (the API behind nvidia-smi) for SM clock and throttle reasons while the timed region runs."""
    REASONS = {0x4: 'sw_power_cap', 0x8: 'hw_slowdown', 0x20: 'sw_thermal_slowdown', 0x40: 'hw_thermal_slowdown',
               0x80: 'hw_power_brake_slowdown', 0x2: 'applications_clocks_setting', 0x10: 'sync_boost'}

    def __init__(self, index):
        self.samples, self.reasons, self.ok = [], set(), False
        self.recording = False
        self._stop = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.max_mhz = None
        self.t = None

    def _run(self):
        while not self._stop.is_set():
            try:
                mhz = self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)
                r = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                if self.recording:          # the thread is started (and its first NVML calls paid) BEFORE the timed region
                    self.samples.append(mhz)
                    for bit, name in self.REASONS.items():
                        if r & bit:
                            self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.002)

    def warm(self):
        """Start polling without recording: thread start-up and NVML's first-call costs (measured: one rank's main thread lost
        11 ms to them, and on N > 1 every rank waits for the slowest one every step) stay outside the timed region."""
        self.recording = False
        self.start(record=False)

    def start(self, record=True):
        self.recording = record
        if self.ok and self.t is None:
            self._stop.clear()
            self.t = threading.Thread(target=self._run, daemon=True)
            self.t.start()

    def stop(self):
        self.recording = False
        if self.t is not None:
            self._stop.set()
            self.t.join()
            self.t = None

    def summary(self):
        if not self.samples:
            return {'sm_mhz': None, 'sm_max_mhz': self.max_mhz, 'reasons': sorted(self.reasons), 'samples': 0}
        return {'sm_mhz': statistics.median(self.samples), 'sm_max_mhz': self.max_mhz, 'reasons': sorted(self.reasons),
                'samples': len(self.samples)}


# --------------------------------------------------------------------------------------------------------------
# CPU port of the reference (oracle) -- used ONLY for the reported baseline / --impl reference
# --------------------------------------------------------------------------------------------------------------
def cpu_reference_step(images, anchors, threads):
    """One pass of the reference algorithm (numpy port) over `images` = list of (probs[1,A',C], reg[1,A',4], ann[1,G,5])."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import head_oracle as O
    params = O.OracleParams()
    n = len(images)

    def one(item):
        p, r, a = item
        out = O.focal_loss(p, r, anchors, a, 0, params, w_bg=[1.0 / n], w_fg=[1.0 / n], w_reg=1.0 / n)
        return float(out['bg'][0]), float(out['fg'][0])
    if threads <= 1:
        return [one(it) for it in images]
    with ThreadPoolExecutor(threads) as ex:
        return list(ex.map(one, images))


def make_cpu_images(count, frac=1.0, seed=1234):
    """`count` COCO-shaped images; frac < 1 keeps only the first frac*A anchors of each (every anchor is independent
    work, so throughput in images/s is (count*frac)/time)."""
    from oracle import head_oracle as O
    rng = np.random.default_rng(seed)
    a_full = O.num_anchors(H, W)
    a = max(1, int(a_full * frac))
    anchors = O.anchors_for_image(H, W)[:, :a]
    imgs = []
    for i in range(count):
        logits = rng.normal(-4.0, 2.0, (1, a, C)).astype(np.float32)
        probs = (1.0 / (1.0 + np.exp(-logits))).astype(np.float32)
        reg = rng.normal(0, 1, (1, a, 4)).astype(np.float32)
        imgs.append((probs, reg, synth_annotations(rng, 1, GMAX, H, W, C, empty=())))
    return imgs, anchors, a / a_full


def time_cpu_reference(steps, warmup, images_per_step, threads, frac=1.0):
    imgs, anchors, f = make_cpu_images(images_per_step, frac)
    for _ in range(warmup):
        cpu_reference_step(imgs, anchors, threads)
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_reference_step(imgs, anchors, threads)
    dt = time.perf_counter() - t0
    return images_per_step * f * steps / dt, dt / steps, f


def time_gpu_eager(probs, reg, anchors, ann, n, steps=3):
    """Informative second baseline (part of the baseline leg, rank 0, N=1): the reference's own execution model -- one
    eager ATen op at a time, a dense [A,C] target matrix, boolean-mask gathers, autograd backward -- restated with
    plain torch calls (oracle/torch_eager.py, bit-identical to the reference on the CPU fixtures) and run on the same
    B200 over the same device-resident batch.  Not our product path; reported beside cpu_baseline."""
    import torch
    from oracle import torch_eager as E

    def step():
        p = probs.detach().requires_grad_(True)
        r = reg.detach().requires_grad_(True)
        bg, fg, rl = E.focal_loss(p, r, anchors, ann)
        (bg.mean() + fg.mean() + rl.mean()).backward()
        return p.grad, r.grad

    try:
        step()
        torch.cuda.synchronize()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(steps):
            step()
        t1.record()
        torch.cuda.synchronize()
        ms = t0.elapsed_time(t1) / steps
    except Exception as e:  # noqa: BLE001  (e.g. out of memory on a shared box): the line is informative only
        return {'unavailable': repr(e)[:200]}
    finally:
        torch.cuda.empty_cache()
    return {'value': n / (ms * 1e-3), 'unit': 'images/s', 'ms_per_step': ms, 'steps': steps, 'kind': 'port',
            'sample': '%d steps x %d COCO-shaped images (the bench batch itself), torch-eager restatement of FocalLoss '
                      'fwd + autograd bwd on the same GPU, inputs resident' % (steps, n)}


def decode_section(dev, with_eager):
    """BASELINE config 4 (the second half of the metric): eval-mode decode + threshold + top-1000 + per-class NMS over
    32 x 800x1333, C=80, device-resident logits, through the C ABI (tools/bench_detect.py).  Two synthetic logit
    distributions: SURVEY 8(d)'s N(-4, 2) (every anchor passes the 0.05 threshold: 200 700 candidates per image, the
    worst case) and a trained-like N(-10.5, 2) (~1.3 k candidates per image).  With `with_eager` the torch-eager
    restatement of ResNet.predict + torchvision.ops.batched_nms (the reference's mode: one image per call, no top-k)
    is timed on the same GPU over the same logits as part of the baseline leg."""
    import argparse as _ap

    import torch
    from tools.bench_detect import measure
    out = {}
    for name, mu, eager_images in (('all_anchors_candidates', -4.0, 2), ('trained_like', -10.5, 8)):
        a = _ap.Namespace(steps=10, warmup=3, images=32, mu=mu, topk=1000, classes=C, head=True)
        line, (logits, reg, anchors, h, w) = measure(a, dev, return_inputs=True)
        entry = {'value': line['value'], 'unit': 'images/s', 'ms_per_step': line['ms_per_step'], 'steps': a.steps,
                 'workload': line['config']['workload'], 'candidates_per_image': line['config']['candidates_per_image'],
                 'kept_per_image': line['config']['kept_per_image'], 'stage_ms': line['stage_ms'], 'roofline': line['roofline'],
                 # the same detections from the head's raw conv outputs (detect_batch_head) vs the reference's layout ops + detect_batch
                 'conv_layout': line.get('conv_layout')}
        if with_eager:
            try:
                from oracle import torch_eager as E
                E.predict(logits[:1], reg[:1], anchors, h, w)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                for j in range(eager_images):
                    s_, l_, b_ = E.predict(logits[j:j + 1], reg[j:j + 1], anchors, h, w)
                torch.cuda.synchronize()
                dt = time.perf_counter() - t0
                entry['gpu_eager_baseline'] = {
                    'value': eager_images / dt, 'unit': 'images/s', 'kind': 'port',
                    'sample': '%d images, one predict() call each (torch-eager sigmoid/decode/clip/max/threshold + '
                              'torchvision batched_nms 0.5, no top-k: the reference has none), same GPU, same logits' % eager_images}
            except Exception as e:  # noqa: BLE001
                entry['gpu_eager_baseline'] = {'unavailable': repr(e)[:200]}
        out[name] = entry
        del logits, reg
        torch.cuda.empty_cache()
    # the boundary as the reference calls it: predict, one image per call, no top-k, host head outputs in, detections out
    from tools.bench_detect import measure_nms_h2h, measure_predict, reference_predict_times
    out['predict_batch1_reference_mode'] = {}
    levels = (('~1.3k candidates', -10.5), ('~8k candidates', -9.5), ('~40k candidates', -8.5))
    for label, mu in levels:
        try:
            out['predict_batch1_reference_mode'][label] = measure_predict(dev, mu, cpu_images=2 if with_eager else 0)
        except Exception as e:  # noqa: BLE001
            out['predict_batch1_reference_mode'][label] = {'unavailable': repr(e)[:200]}
        torch.cuda.empty_cache()
    if with_eager:
        # the UNMODIFIED ResNet.predict on the same head-output distribution: on this GPU as written, and on the host cores
        ref = reference_predict_times([mu for _, mu in levels])
        for label, mu in levels:
            out['predict_batch1_reference_mode'][label].update(ref[mu])
    # K6 (+ its sort) against torchvision's own CUDA nms on identical inputs
    out['nms_vs_torchvision'] = {}
    for k in (1000, 8000, 40000):
        try:
            out['nms_vs_torchvision'][str(k)] = measure_nms_h2h(dev, k)
        except Exception as e:  # noqa: BLE001
            out['nms_vs_torchvision'][str(k)] = {'unavailable': repr(e)[:200]}
    return out


def reference_snapshot_available():
    return os.path.exists(os.path.join(ROOT, 'oracle', '_ref', 'retinanet', 'losses.bytecode'))


def time_unmodified_reference(images, frac, steps, warmup, threads, timeout=900):
    """The UNMODIFIED reference (oracle/_ref: byte-code snapshot of retinanet/losses.py built by oracle/build_ref.py) on the host
    cores, in its own process (its CPU shim patches torch globally).  Returns oracle.ref_runner's JSON dict."""
    import subprocess
    cmd = [sys.executable, '-m', 'oracle.ref_runner', '--images', str(images), '--frac', '%.6f' % frac, '--steps', str(steps),
           '--warmup', str(warmup), '--threads', str(threads)]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=timeout)
    if r.returncode != 0:
        raise RuntimeError('oracle.ref_runner failed: %s' % r.stderr[-500:])
    return json.loads(r.stdout.strip().splitlines()[-1])


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on the box's host cores.  With the byte-code snapshot
    oracle/_ref present (built where /root/reference exists; it travels with the repo) this is the UNMODIFIED FocalLoss.forward +
    autograd backward (kind "reference"); the numpy port is timed beside it as `cpu_baseline_port`.  Without the snapshot the
    port is the arm (kind "port")."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    threads = max(1, min(cores, CPU_THREADS_CAP))
    budget = float(os.environ.get('CLDET_BENCH_BUDGET_S', '150'))      # seconds of CPU work for the whole run (tests shrink it)
    total_steps = args.steps + args.warmup
    port = None
    if reference_snapshot_available():
        kind = 'reference'
        images_per_step = 1
        probe = time_unmodified_reference(1, 0.25, 1, 1, threads)           # one image, a quarter of the anchors, after a warm-up call
        full_step = probe['s_per_step'] / probe['anchor_fraction']
        frac = max(0.02, min(1.0, 0.8 * budget / (full_step * total_steps)))
        res = time_unmodified_reference(images_per_step, frac, args.steps, args.warmup, threads, timeout=max(900, 20 * budget))
        val, per_step, f = res['value'], res['s_per_step'], res['anchor_fraction']
        sample = ('%d image x %.0f%% of the anchors per step (800x1333, C=80, A=200700 full), UNMODIFIED reference FocalLoss.forward '
                  '+ autograd backward (retinanet/losses.py byte-code snapshot oracle/_ref, CPU shim of SURVEY 8c), torch %s, '
                  '%d intra-op threads' % (images_per_step, f * 100, res['torch'], res['threads']))
        pv, pstep, pf = time_cpu_reference(1, 0, threads, threads, max(0.02, min(1.0, 0.2 * budget / max(full_step / 15.0, 1e-3))))
        port = {'value': pv, 'unit': 'images/s', 'cores': threads, 'kind': 'port', 'host_cores': cores,
                'sample': '1 step x %d images x %.0f%% of the anchors, numpy port of FocalLoss fwd+bwd, %d threads'
                          % (threads, pf * 100, threads)}
    else:
        kind = 'port'
        images_per_step = threads
        # keep the whole run to a few minutes: probe a quarter-size step, then pick the anchor fraction of the per-step sample
        _, probe_step, f0 = time_cpu_reference(1, 0, images_per_step, threads, 0.25)
        full_step = probe_step / f0
        frac = max(0.02, min(1.0, budget / (full_step * total_steps)))
        val, per_step, f = time_cpu_reference(args.steps, args.warmup, images_per_step, threads, frac)
        sample = ('%d images x %.0f%% of the anchors per step (800x1333, C=80, A=200700 full), numpy port of FocalLoss fwd+bwd, '
                  '%d threads' % (images_per_step, f * 100, threads))
    line = {'impl': 'reference', 'metric': METRIC, 'value': val, 'unit': 'images/s', 'n_gpus': args.gpus, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': per_step * 1e3, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': WORKLOAD, 'images_per_step': images_per_step * f},
            'cpu_baseline': {'value': val, 'unit': 'images/s', 'cores': threads, 'kind': kind, 'sample': sample,
                             'host_cores': cores},
            'e2e': {'value': val, 'unit': 'images/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    if port is not None:
        line['cpu_baseline_port'] = port
    emit(line)
    return 0


def bind_near_gpu(index):
    """Best effort, for the e2e leg on multi-socket boxes: run this rank on the CPUs NVML reports as local to its GPU and
    prefer that NUMA node for the pinned staging buffers, so that eight ranks do not pull their host->device copies across
    the socket interconnect.  Returns a short description for the JSON line; never raises."""
    note = []
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, 16)                 # 1024 CPUs
        cpus = {64 * i + b for i, wd in enumerate(words) for b in range(64) if (int(wd) >> b) & 1}
        usable = cpus & os.sched_getaffinity(0)
        if usable:
            os.sched_setaffinity(0, usable)
            note.append('cpus=%d' % len(usable))
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        node_file = '/sys/bus/pci/devices/%s/numa_node' % bus.lower()[-12:]
        node = int(open(node_file).read().strip()) if os.path.exists(node_file) else -1
        if node >= 0:
            libc = ctypes.CDLL(None, use_errno=True)
            mask = (ctypes.c_ulong * 16)()
            mask[node // 64] = 1 << (node % 64)
            # set_mempolicy(MPOL_PREFERRED = 1, nodemask, maxnode): syscall 238 on x86_64
            rc = libc.syscall(238, 1, mask, 1024)
            note.append('numa_node=%d%s' % (node, '' if rc == 0 else ' (mempolicy refused)'))
    except Exception as e:  # noqa: BLE001
        note.append('unavailable: %s' % type(e).__name__)
    return ', '.join(note) if note else 'none'


# --------------------------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------------------------
def make_inputs(dev, cfg, seed):
    """Seeded synthetic batch of one BASELINE config: probabilities sigmoid(N(-4,2)), reg N(0,1), random GT."""
    import torch

    import cl_object_detection_b200 as cld
    anchors = cld.generate_anchors(cfg['h'], cfg['w'], dev)
    a = anchors.shape[1]
    gen = torch.Generator(device=dev).manual_seed(seed)
    probs = torch.sigmoid(torch.randn(cfg['n'], a, cfg['c'], device=dev, generator=gen) * 2.0 - 4.0)
    reg = torch.randn(cfg['n'], a, 4, device=dev, generator=gen)
    ann_np = synth_annotations(np.random.default_rng(seed), cfg['n'], cfg['gmax'], cfg['h'], cfg['w'], cfg['c'],
                               empty=cfg.get('empty', (0,)), exact=cfg.get('exact_g', False), pseudo_split=cfg.get('past') or None)
    return probs, reg, ann_np, anchors


class LossStep:
    """One step of the loss path THROUGH THE PUBLIC DROP-IN: FocalLoss.forward (or ShardedFocalLoss.forward on N > 1 ranks) ->
    torch.ops.cldet.focal_loss (C++ op layer -> C ABI -> kernels), then autograd backward of its outputs with the upstream
    gradients the caller's reduction delivers (IL_Loss: .mean() of every term, losses.py:584-588 -> 1/N_global per image)."""

    def __init__(self, probs, reg, ann, anchors, state, params, n_global, module):
        import torch
        self.torch = torch
        self.p = probs.detach().requires_grad_(True)
        self.r = reg.detach().requires_grad_(True)
        self.ann, self.anchors, self.state, self.params, self.module = ann, anchors, state, params, module
        dev = probs.device
        self.g_rows = torch.full((n_global,), 1.0 / n_global, device=dev)      # dL/dbg_j = dL/dfg_j of the caller's mean
        self.g_one = torch.ones(1, device=dev)                                 # dL/dreg_loss
        self.out = None

    def __call__(self):
        out = self.module(self.p, self.r, self.anchors, self.ann, self.state, self.params)
        bg, fg = out['cls_loss']
        self.out = out
        return self.torch.autograd.grad([bg, fg, out['reg_loss']], [self.p, self.r], [self.g_rows, self.g_rows, self.g_one])


def time_loop(step, steps, warmup, barrier, hook=None, sampler=None):
    """`steps` timed iterations after `warmup`, CUDA events on the current stream, barrier + synchronize on both sides; one
    more untimed step directly before the timed region aligns the ranks (on N > 1 it ends in the peer exchange)."""
    import torch
    if sampler is not None:
        sampler.warm()           # polling thread up and NVML warm before anything is timed
    for i in range(max(warmup, 3)):
        if hook is not None and i == 0:
            hook(0)              # the in-call event records pay their first-use cost here, not in timed step 0
        step()
    barrier()
    step()
    if sampler is not None:
        sampler.start()          # SM clock + throttle reasons are RECORDED during the timed region only
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    trace = os.environ.get('CLDET_BENCH_TRACE')          # diagnostics: per-step device and host time lines of every rank
    marks, host = [], []
    t0.record()
    for i in range(steps):
        if hook is not None:
            hook(i)
        if trace:
            host.append(time.perf_counter())
        step()
        if trace:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            marks.append(e)
    t1.record()
    barrier()
    if sampler is not None:
        sampler.stop()
    if trace:
        os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
        with open(os.path.join(ROOT, 'gpurun_out', 'trace_%s_rank%s.json' % (trace, os.environ.get('RANK', '0'))), 'w') as f:
            json.dump({'device_ms_since_t0': [t0.elapsed_time(e) for e in marks],
                       'host_ms_since_first': [(h - host[0]) * 1e3 for h in host]}, f)
    return t0.elapsed_time(t1)


LOSS_CONFIGS = {
    1: dict(name="1: VOC '20' state 0, 2 x 512x512, C=20", n=2, h=512, w=512, c=20, gmax=20, state=0, past=0),
    2: dict(name="2: VOC '15_1' state 1 + pseudo-label GT rows, 16 x 512x512, C=16 (15 old + 1 new)", n=16, h=512, w=512, c=16,
            gmax=20, state=1, past=15),
    3: dict(name='3: COCO-shaped, 16 x 800x1333, C=80', n=N_PER_GPU, h=H, w=W, c=C, gmax=GMAX, state=0, past=0),
    5: dict(name='5: dense-GT stress, 8 x 1333x1333 per GPU, C=80, exactly 100 GT boxes per image', n=8, h=1333, w=1333, c=80,
            gmax=100, state=0, past=0, exact_g=True, empty=()),
}


def config_lines(dev, world, rank, steps, barrier, reduce_max, make_module):
    """Device-resident lines for the other loss configs of BASELINE.json through the same public step: 1 and 2 (VOC shapes; one
    GPU only -- they are L2-resident, launch-bound problems) and 5 (dense GT) on every N, image-sharded like config 3."""
    import torch

    import cl_object_detection_b200 as cld
    peak, _ = peak_hbm()
    out = {}
    for cid in ((1, 2, 5) if world == 1 else (5,)):
        cfg = LOSS_CONFIGS[cid]
        probs, reg, ann_np, anchors = make_inputs(dev, cfg, 1000 * cid + rank)
        ann = torch.from_numpy(ann_np).to(dev)
        params = cld.HeadParams([0, cfg['past']]) if cfg['state'] else cld.HeadParams()
        n_global = cfg['n'] * world
        step = LossStep(probs, reg, ann, anchors, cfg['state'], params, n_global, make_module())
        ms = reduce_max(time_loop(step, steps, 5, barrier)) / steps
        a = anchors.shape[1]
        path_bytes = cfg['n'] * (8 * a * cfg['c'] + 48 * a + 20 * cfg['gmax'])
        gbs = path_bytes / (ms * 1e-3) / 1e9
        out[str(cid)] = {'workload': cfg['name'], 'value': n_global / (ms * 1e-3), 'unit': 'images/s', 'ms_per_step': ms,
                         'steps': steps, 'images_per_gpu': cfg['n'], 'anchors': a, 'n_gpus': world,
                         'path_bytes_per_gpu_step': path_bytes, 'achieved_gbs_per_gpu': gbs, 'roofline_frac': gbs / peak,
                         'note': 'working set %.0f MB per GPU%s' % (2 * probs.numel() * 4 / 1e6,
                                                                    ' (fits the 126 MB L2: latency-bound, not a roofline case)'
                                                                    if 2 * probs.numel() * 4 < 126e6 else '')}
        mod = step.module
        del step, probs, reg
        if world > 1:
            dist_barrier_close(mod, barrier)
        torch.cuda.empty_cache()
    return out


def dist_barrier_close(module, barrier):
    barrier()
    if hasattr(module, 'close'):
        module.close()


def sharded_parity(dev, world, rank, sharded, cfg):
    """Driver-visible correctness of the image-sharded path (N > 1), on config 3's own batch:
    (a) for two consecutive steps (both parities of the exchange buffers) the [4, N_global] rows delivered by the FUSED peer
        exchange must be bit-equal to an NCCL all-gather of every rank's locally computed [4, n] rows;
    (b) rank 0 recomputes the GLOBAL batch on its one GPU; IL_Loss's clip_loss reduction (losses.py:572-588: bg.mean() +
        fg[fg >= clip].mean() + reg.mean(), clip = the median fg so that half the images are masked) must give the same loss,
        and the gradients of rank 0's own shard must match what the sharded run produced."""
    import torch
    import torch.distributed as dist

    import cl_object_detection_b200 as cld
    n = cfg['n']
    params = cld.HeadParams()
    probs, reg, ann_np, anchors = make_inputs(dev, cfg, 3000 + rank)
    ann = torch.from_numpy(ann_np).to(dev)
    local_fl = cld.FocalLoss()
    bit_equal = True
    with torch.no_grad():
        for _ in range(2):
            out = sharded(probs, reg, anchors, ann, 0, params)
            peer_rows = torch.stack([out['cls_loss'][0], out['cls_loss'][1], sharded.local_loss.last_reg_per_image])
            lo = local_fl(probs, reg, anchors, ann, 0, params)
            mine = torch.stack([lo['cls_loss'][0], lo['cls_loss'][1], local_fl.last_reg_per_image]).contiguous()     # [3, n]
            allr = torch.empty((world, 3, n), device=dev)
            dist.all_gather_into_tensor(allr, mine)
            nccl_rows = allr.permute(1, 0, 2).reshape(3, world * n)
            bit_equal = bit_equal and bool(torch.equal(peer_rows, nccl_rows))

    def reduction(o, clip):
        bg, fg = o['cls_loss']
        m = fg >= clip
        return bg.mean() + (fg[m].mean() if bool(m.any()) else fg.sum() * 0) + o['reg_loss'].mean()

    p = probs.detach().requires_grad_(True)
    r = reg.detach().requires_grad_(True)
    out = sharded(p, r, anchors, ann, 0, params)
    clip = float(out['cls_loss'][1].detach().median())
    loss = reduction(out, clip)
    gp, gr = torch.autograd.grad(loss, [p, r])
    flag = torch.tensor([1 if bit_equal else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    res = {'terms_bit_equal': bool(int(flag.item())), 'checked_parities': 2, 'world': world,
           'reduction': 'IL_Loss clip_loss (losses.py:572-588), clip = median fg'}
    if rank == 0:
        parts = [make_inputs(dev, cfg, 3000 + q) for q in range(world)]
        gp_all = torch.cat([x[0] for x in parts]).requires_grad_(True)
        gr_all = torch.cat([x[1] for x in parts]).requires_grad_(True)
        ann_all = torch.from_numpy(np.concatenate([x[2] for x in parts])).to(dev)
        del parts
        single = cld.FocalLoss()(gp_all, gr_all, anchors, ann_all, 0, params)
        ref = reduction(single, clip)
        rp, rr = torch.autograd.grad(ref, [gp_all, gr_all])

        def max_rel(got, want):
            d = (got - want).abs()
            return float((d / want.abs().clamp_min(1e-30)).max()) if bool((d > 0).any()) else 0.0
        res.update({'loss_rel': abs(float(loss) - float(ref)) / abs(float(ref)),
                    'grad_cls_max_rel': max_rel(gp, rp[:n]), 'grad_reg_max_rel': max_rel(gr, rr[:n]),
                    'grad_bit_equal': bool(torch.equal(gp, rp[:n]) and torch.equal(gr, rr[:n])),
                    'global_rows_bit_equal_single_gpu': bool(torch.equal(out['cls_loss'][0], single['cls_loss'][0]) and
                                                             torch.equal(out['cls_loss'][1], single['cls_loss'][1]))})
        res['grad_max_rel'] = max(res['grad_cls_max_rel'], res['grad_reg_max_rel'])
        del gp_all, gr_all, rp, rr, single
    torch.cuda.empty_cache()
    dist.barrier()
    return res


def run_ours(args):
    import torch
    import torch.distributed as dist

    import cl_object_detection_b200 as cld
    from cl_object_detection_b200 import _lib

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device (there is no CPU fallback for the product path)')
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    numa = bind_near_gpu(local_rank) if world > 1 else 'not applied (1 GPU)'
    if world > 1:
        # NCCL writes its banner / debug lines to stdout unless told otherwise: keep stdout for the one JSON line
        os.environ.setdefault('NCCL_DEBUG_FILE', '/dev/stderr')
        dist.init_process_group('nccl', device_id=dev)
    lib = _lib.load()
    cld.load_ops()

    cfg = LOSS_CONFIGS[3]
    n = cfg['n']
    n_global = n * world
    probs, reg, ann_np, anchors = make_inputs(dev, cfg, 3000 + rank)
    a = anchors.shape[1]
    ann = torch.from_numpy(ann_np).to(dev)
    params = cld.HeadParams()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_max(x):
        if world == 1:
            return float(x)
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # N > 1: every rank needs every image's (bg, fg, reg) terms (IL_Loss's mean / clip_loss mask).  ShardedFocalLoss: the loss
    # kernel itself pushes them into all ranks' gather buffers over NVLink peer stores and a one-block kernel waits for the
    # arrivals and copies them out (`--collective nccl`: an NCCL all-gather after the kernel instead).
    def make_module():
        if world == 1:
            return cld.FocalLoss()
        return cld.ShardedFocalLoss(use_peer_memory=(args.collective == 'peer'))

    module = make_module()
    step = LossStep(probs, reg, ann, anchors, 0, params, n_global, module)
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    for tri in ev:
        for e in tri:
            e.record()          # materialise the CUDA event handles
    torch.cuda.synchronize()

    HOOK_EVERY = 4

    def hook(i):
        # the next fused call of this thread records events around its two kernels: the loss kernel is timed inside the real
        # step.  Every 4th step only: an event record between two kernels also keeps the second from being scheduled while
        # the first drains (programmatic dependent launch), which the other steps keep.
        if i % HOOK_EVERY == 0:
            lib.cldet_focal_loss_profile_events(ev[i][0].cuda_event, ev[i][1].cuda_event, ev[i][2].cuda_event)

    sampler = ClockSampler(local_rank) if not args.no_clock_sampler else None
    total_ms = time_loop(step, args.steps, args.warmup, barrier, hook if not args.no_kernel_events else None, sampler)
    if args.quick:          # diagnostics: the timed loop only
        total_ms = reduce_max(total_ms)
        if rank == 0:
            emit({'quick': True, 'n_gpus': world, 'steps': args.steps, 'ms_per_step': total_ms / args.steps,
                  'value': n_global / (total_ms / args.steps * 1e-3), 'collective': args.collective})
        if world > 1:
            dist.barrier()
            module.close()
            dist.destroy_process_group()
        return 0
    collective = 'none'
    if world > 1:
        used_peer = bool(module._peer and any(v for v in module._peer.values()))
        collective = 'peer' if used_peer else 'nccl'
        if used_peer:
            for pg in module._peer.values():
                pg.check()          # a timed-out exchange raises here (status word in mapped host memory)
    hooked = [t for i, t in enumerate(ev) if i % HOOK_EVERY == 0]
    assign_ms = sum(t[0].elapsed_time(t[1]) for t in hooked) / len(hooked)
    loss_ms = sum(t[1].elapsed_time(t[2]) for t in hooked) / len(hooked)
    if args.steps < 300:
        # the timed region is a few milliseconds -- shorter than a handful of sampling periods: keep sampling the clocks under the
        # very same load for a FIXED number of extra steps (the same on every rank: the ranks exchange terms every step)
        sampler.start()
        for _ in range(600):
            step()
        torch.cuda.synchronize()
        sampler.stop()
    clocks = sampler.summary()
    # host time to ENQUEUE one step through the drop-in (how far the CPU is from being the limit)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20):
        step()
    host_us = (time.perf_counter() - t0) / 20 * 1e6
    barrier()

    # ---- end to end through the public drop-in, host buffers in, losses out ----
    e2e_steps = max(3, min(args.steps, 10))
    h_probs = probs.cpu().pin_memory()
    h_reg = reg.cpu().pin_memory()
    h_ann = torch.from_numpy(ann_np).pin_memory()
    h_out = torch.empty(3, dtype=torch.float32).pin_memory()
    d_probs = torch.empty_like(probs)
    d_reg = torch.empty_like(reg)
    d_ann = torch.empty_like(ann)

    def e2e_step():
        d_probs.copy_(h_probs, non_blocking=True)
        d_reg.copy_(h_reg, non_blocking=True)
        d_ann.copy_(h_ann, non_blocking=True)
        p = d_probs.detach().requires_grad_(True)
        r = d_reg.detach().requires_grad_(True)
        out = module(p, r, anchors, d_ann, 0, params)          # rows of the GLOBAL batch on N > 1
        bg, fg = out['cls_loss']
        terms = torch.stack([bg.mean(), fg.mean(), out['reg_loss'].mean()])      # the caller's reductions (losses.py:584-588)
        g = torch.autograd.grad(terms.sum(), [p, r])
        h_out.copy_(terms.detach(), non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return g

    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_s = reduce_max((time.perf_counter() - t0) / e2e_steps)
    h2d = h_probs.numel() * 4 + h_reg.numel() * 4 + h_ann.numel() * 4
    d2h = h_out.numel() * 4
    del h_probs, h_reg, d_probs, d_reg

    # ---- max over ranks ----
    total_ms, loss_ms, assign_ms = reduce_max(total_ms), reduce_max(loss_ms), reduce_max(assign_ms)
    ms_per_step = total_ms / args.steps
    value = n_global / (ms_per_step * 1e-3)

    peak, peak_src = peak_hbm()
    kernel_bytes = n * (8 * a * C + 20 * a)                    # loss kernel: p read + dL/dp write + assignment word + dL/dreg
    path_bytes = n * (8 * a * C + 48 * a + 20 * GMAX)          # SURVEY 8(d) figure for the whole loss path
    achieved = kernel_bytes / (loss_ms * 1e-3) / 1e9
    traffic = ncu_traffic('focal_loss_kernel')
    roofline = {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                'traffic': traffic, 'kernel': 'focal_loss_kernel<VEC=8,GAMMA2,no IL variants,GRAD,probabilities> (256-bit LDG/STG)', 'kernel_ms': loss_ms,
                'assign_kernel_ms': assign_ms, 'algorithmic_bytes_per_launch': kernel_bytes, 'peak_source': peak_src,
                'frac_nominal': achieved / NOMINAL_HBM_GBS, 'nominal_peak': NOMINAL_HBM_GBS,
                'path_frac': path_bytes / ((loss_ms + assign_ms) * 1e-3) / 1e9 / peak,
                'step_frac': path_bytes / (ms_per_step * 1e-3) / 1e9 / peak,
                'step_frac_nominal': path_bytes / (ms_per_step * 1e-3) / 1e9 / NOMINAL_HBM_GBS}

    cfg_lines = config_lines(dev, world, rank, 50, barrier, reduce_max, make_module) if not args.no_configs else None
    parity = sharded_parity(dev, world, rank, module, cfg) if (world > 1 and collective == 'peer') else None

    if rank == 0:
        cores = os.cpu_count() or 1
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            thr = max(1, min(cores, CPU_THREADS_CAP))
            v, per, _ = time_cpu_reference(steps=2, warmup=0, images_per_step=thr, threads=thr)
            cpu_port = {'value': v, 'unit': 'images/s', 'cores': thr, 'kind': 'port', 'host_cores': cores,
                        'sample': '2 steps x %d COCO-shaped images (800x1333, C=80, A=200700), numpy port of FocalLoss fwd+bwd, '
                                  '%d threads' % (thr, thr)}
            cpu = cpu_port
            if reference_snapshot_available():
                try:
                    res = time_unmodified_reference(2, 1.0, 2, 1, thr)
                    cpu = {'value': res['value'], 'unit': 'images/s', 'cores': thr, 'kind': 'reference', 'host_cores': cores,
                           'sample': '2 steps x 2 COCO-shaped images (800x1333, C=80, A=200700), UNMODIFIED reference FocalLoss.forward '
                                     '+ autograd backward (byte-code snapshot oracle/_ref), torch %s, %d intra-op threads'
                                     % (res['torch'], res['threads'])}
                except Exception as e:  # noqa: BLE001
                    cpu_port['reference_snapshot_error'] = repr(e)[:200]
        eager = None
        if world == 1 and not args.no_cpu_baseline:
            eager = time_gpu_eager(probs, reg, anchors, ann, n)
        line = {'metric': METRIC, 'value': value, 'unit': 'images/s', 'n_gpus': world, 'steps': args.steps,
                'warmup': max(args.warmup, 3), 'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'weak',
                'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
                'config': {'workload': WORKLOAD, 'images_per_gpu': n, 'global_batch': n_global,
                           'parallelism': 'image-sharded dp%d' % world,
                           'api': 'FocalLoss.forward -> torch.ops.cldet.focal_loss (C++ op layer over the C ABI) + autograd backward'
                                  if world == 1 else 'ShardedFocalLoss.forward -> torch.ops.cldet.focal_loss + autograd backward',
                           'collective': {'none': 'none (1 GPU)', 'peer': 'fused into the loss kernel: NVLink peer stores + arrival counters',
                                          'nccl': 'NCCL all-gather of the [4,N] terms'}[collective],
                           'l2': 'inputs (%.2f GB per GPU) are larger than the 126 MB L2; no flush needed' % (probs.numel() * 4 / 1e9),
                           'host_binding_rank0': numa},
                'clocks': clocks,
                'e2e': {'value': n_global / e2e_s, 'unit': 'images/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                        'ms_per_step': e2e_s * 1e3, 'steps': e2e_steps},
                'gpu_launches': (4 if collective == 'peer' else 3) * args.steps,
                'host_enqueue_us_per_step': host_us,
                'roofline': roofline}
        if cfg_lines is not None:
            line['configs'] = cfg_lines
        if parity is not None:
            line['sharded_parity'] = parity
        if cpu is not None:
            line['cpu_baseline'] = cpu
            if cpu is not cpu_port:
                line['cpu_baseline_port'] = cpu_port
        if eager is not None:
            line['gpu_eager_baseline'] = eager
        if world == 1 and not args.no_decode:
            del probs, reg, step
            torch.cuda.empty_cache()
            line['decode'] = decode_section(dev, with_eager=not args.no_cpu_baseline)
        emit(line)
    if world > 1:
        dist.barrier()
        module.close()
        dist.destroy_process_group()
    return 0


_JSON_FD = None


def _claim_stdout():
    """stdout must carry exactly ONE JSON line, but libraries under us print there too (NCCL's version banner, warnings):
    point fd 1 at stderr for the whole run and keep the real stdout for emit()."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + '\n').encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=100)
    ap.add_argument('--warmup', type=int, default=10)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--no-cpu-baseline', action='store_true', help='skip the CPU port and the GPU-eager baseline legs')
    ap.add_argument('--no-decode', action='store_true', help='skip the BASELINE config 4 (decode + NMS) section')
    ap.add_argument('--no-configs', action='store_true', help='skip the lines for BASELINE configs 1, 2 and 5')
    ap.add_argument('--collective', default='peer', choices=['peer', 'nccl'])
    ap.add_argument('--quick', action='store_true', help='diagnostics: run the timed loop only and print a short line')
    ap.add_argument('--no-clock-sampler', action='store_true', help='diagnostics: do not poll NVML during the timed region')
    ap.add_argument('--no-kernel-events', action='store_true', help='diagnostics: no in-call event records (kernel_ms unavailable)')
    args = ap.parse_args()
    if (args.no_clock_sampler or args.no_kernel_events) and not args.quick:
        ap.error('--no-clock-sampler / --no-kernel-events are diagnostics of the --quick mode (the full line needs both)')
    if args.impl == 'reference':
        return run_reference(args)
    return run_ours(args)


if __name__ == '__main__':
    sys.exit(main())
