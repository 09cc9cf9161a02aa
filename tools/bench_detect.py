#!/usr/bin/env python
"""Eval-mode detection output benchmark (BASELINE config 4): decode + score threshold + top-1000 + per-class NMS on a
COCO-shaped batch (32 x 800x1333, C=80, A=200700).  Prints one JSON line; per-stage times come from CUDA events around
the C-ABI calls.  Not the default bench line (bench.py measures the loss path the metric's target is quoted on)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import cl_object_detection_b200 as cld  # noqa: E402
from cl_object_detection_b200 import _lib  # noqa: E402


def ncu_traffic(kernel):
    """dram bytes per launch from the committed ncu --set full capture (profiles/traffic.json; trained-like logits), or None."""
    try:
        return json.load(open(os.path.join(ROOT, 'profiles', 'traffic.json'))).get(kernel)
    except Exception:
        return None


def measure(args, dev=None, return_inputs=False):
    """One JSON-able dict for the configuration in `args` (steps, warmup, images, mu, topk, classes)."""
    if dev is None:
        dev = torch.device('cuda', 0)
        torch.cuda.set_device(dev)
    lib = _lib.load()
    h, w, c, n, topk = 800, 1333, args.classes, args.images, args.topk
    anchors = cld.generate_anchors(h, w, dev)
    a = anchors.shape[1]
    gen = torch.Generator(device=dev).manual_seed(4000)
    logits = torch.randn(n, a, c, device=dev, generator=gen) * 2.0 + args.mu
    reg = torch.randn(n, a, 4, device=dev, generator=gen) * 0.5
    st = torch.cuda.current_stream().cuda_stream
    cap = min(topk, a)
    counts = torch.zeros(n, dtype=torch.int32, device=dev)
    cand = torch.empty((n, a, 32), dtype=torch.uint8, device=dev)
    keys = torch.empty((n, a), dtype=torch.int64, device=dev)
    sorted_c = torch.empty((n, cap, 32), dtype=torch.uint8, device=dev)
    sorted_counts = torch.empty(n, dtype=torch.int32, device=dev)
    sws = torch.empty(lib.cldet_sort_workspace_bytes(n, a, topk), dtype=torch.uint8, device=dev)
    nws = torch.empty(lib.cldet_nms_workspace_bytes(n, cap), dtype=torch.uint8, device=dev)
    keep = torch.empty((n, cap), dtype=torch.int32, device=dev)
    keep_counts = torch.empty(n, dtype=torch.int32, device=dev)
    scores = torch.empty((n, cap), device=dev)
    labels = torch.empty((n, cap), dtype=torch.int64, device=dev)
    boxes = torch.empty((n, cap, 4), device=dev)
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(args.steps)]

    def step(i=None):
        counts.zero_()
        if i is not None:
            ev[i][0].record()
        _lib.check(lib.cldet_decode_filter(logits.data_ptr(), 1, reg.data_ptr(), anchors.data_ptr(), n, a, c, h, w, 0.05,
                                           cand.data_ptr(), keys.data_ptr(), a, counts.data_ptr(), st))
        if i is not None:
            ev[i][1].record()
        _lib.check(lib.cldet_sort_candidates(cand.data_ptr(), keys.data_ptr(), counts.data_ptr(), n, a, a, topk,
                                             sorted_c.data_ptr(), cap, sorted_counts.data_ptr(), sws.data_ptr(), sws.numel(), st))
        if i is not None:
            ev[i][2].record()
        # NMS + gather of the kept candidates (one chain: the block that resolves an image gathers it)
        _lib.check(lib.cldet_nms_gather_sorted(sorted_c.data_ptr(), sorted_counts.data_ptr(), n, cap, cap, 0.5, 0, 100000,
                                               keep.data_ptr(), keep_counts.data_ptr(), scores.data_ptr(), labels.data_ptr(),
                                               boxes.data_ptr(), nws.data_ptr(), nws.numel(), st))
        if i is not None:
            ev[i][3].record()

    for _ in range(max(3, args.warmup)):
        step()
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for i in range(args.steps):
        step(i)
    t1.record()
    torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / args.steps
    stage = [sum(e[k].elapsed_time(e[k + 1]) for e in ev) / args.steps for k in range(3)]
    kept = keep_counts.float().mean().item()
    ncand = counts.float().mean().item()
    peak = 6544.7
    try:
        peak = float(json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs'])
    except Exception:
        pass
    filt_bytes = n * (4 * a * c) + counts.sum().item() * (32 + 40)
    line = {'metric': 'detection-head images/sec (decode + threshold + top-k + per-class NMS)', 'value': n / (ms * 1e-3),
            'unit': 'images/s', 'n_gpus': 1, 'steps': args.steps, 'ms_per_step': ms, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': 'eval decode: %d x 800x1333, C=%d, A=%d, thr 0.05, top-%d, NMS 0.5 (BASELINE config 4)' % (n, c, a, topk),
                       'logit_mean': args.mu, 'candidates_per_image': ncand, 'kept_per_image': kept},
            'stage_ms': {'decode_filter': stage[0], 'select_sort': stage[1], 'nms_gather': stage[2]},
            'roofline': {'bound': 'hbm', 'kernel': 'decode_filter_kernel<4>', 'achieved': filt_bytes / (stage[0] * 1e-3) / 1e9,
                         'peak': peak, 'unit': 'GB/s', 'frac': filt_bytes / (stage[0] * 1e-3) / 1e9 / peak,
                         'algorithmic_bytes_per_launch': filt_bytes, 'traffic': ncu_traffic('decode_filter_kernel')}}
    if getattr(args, 'head', False):
        line['conv_layout'] = measure_head(args, dev, logits, reg, anchors, h, w)
    if return_inputs:
        return line, (logits, reg, anchors, h, w)
    return line


def measure_head(args, dev, logits, reg, anchors, h, w):
    """SURVEY 8f row f1, eval side: head outputs in conv layout -> detections.  ref-layout = the reference's per-level
    permute + contiguous + view and torch.cat (model.py:125-130, 170-184, 472-474) followed by detect_batch; head-layout =
    detect_batch_head on the conv outputs as they are."""
    from cl_object_detection_b200 import detect as D
    n, a, c = logits.shape
    shapes = [((h + 2 ** l - 1) // 2 ** l, (w + 2 ** l - 1) // 2 ** l) for l in range(3, 8)]

    def to_levels(x, per):
        out, off = [], 0
        for hl, wl in shapes:
            cnt = hl * wl * 9
            out.append(x[:, off:off + cnt].reshape(n, hl, wl, 9 * per).permute(0, 3, 1, 2).contiguous())
            off += cnt
        return out

    cls_lv, reg_lv = to_levels(logits, c), to_levels(reg, 4)

    def cat_layout(levels, per):
        return torch.cat([t.permute(0, 2, 3, 1).contiguous().view(t.shape[0], -1, per) for t in levels], dim=1)

    def ref_step():
        return D.detect_batch(cat_layout(cls_lv, c), cat_layout(reg_lv, 4), anchors, h, w, pre_nms_topk=args.topk, return_padded=True)

    def head_step():
        return D.detect_batch_head(cls_lv, reg_lv, anchors, h, w, pre_nms_topk=args.topk, return_padded=True)

    def timeit(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(args.steps):
            fn()
        t1.record()
        torch.cuda.synchronize()
        return t0.elapsed_time(t1) / args.steps

    r0, r1 = ref_step(), head_step()
    # padded outputs: compare the valid prefix of every image (the padding is uninitialised memory)
    same = torch.equal(r0[3], r1[3])
    if same:
        for j, kj in enumerate(r0[3].tolist()):
            same = same and all(torch.equal(x0[j, :kj], x1[j, :kj]) for x0, x1 in zip(r0[:3], r1[:3]))
    ms_ref, ms_head = timeit(ref_step), timeit(head_step)
    return {'ref_layout_ms': ms_ref, 'head_layout_ms': ms_head, 'speedup': ms_ref / ms_head,
            'images_per_s_head_layout': n / (ms_head * 1e-3), 'identical_detections': bool(same)}


def measure_predict(dev, mu, images=6, cpu_images=2):
    """The eval boundary AS THE REFERENCE CALLS IT (evaluator.py:324-329 -> ResNet.predict, model.py:507-550): one image per
    call, no top-k, head outputs of that image arriving from the HOST (pinned) and the detections read back with .cpu().
    Timed wall clock per call, after warm-up.  cpu_baseline: the torch-eager restatement of predict + torchvision's CPU
    batched_nms on the box's host cores over the same image (what the reference's own code does on a CPU)."""
    import time

    from cl_object_detection_b200 import detect as D
    from oracle import torch_eager as E
    h, w, c = 800, 1333, 80
    anchors = cld.generate_anchors(h, w, dev)
    a = anchors.shape[1]
    gen = torch.Generator(device=dev).manual_seed(int(-mu * 100))
    logits = torch.randn(images, a, c, device=dev, generator=gen) * 2.0 + mu
    reg = torch.randn(images, a, 4, device=dev, generator=gen) * 0.3
    h_logits = [logits[j:j + 1].cpu().pin_memory() for j in range(images)]
    h_reg = [reg[j:j + 1].cpu().pin_memory() for j in range(images)]
    d_logits = torch.empty_like(logits[:1])
    d_reg = torch.empty_like(reg[:1])

    def call(j):
        d_logits.copy_(h_logits[j], non_blocking=True)
        d_reg.copy_(h_reg[j], non_blocking=True)
        s, l, b = D.detect_batch(d_logits, d_reg, anchors, h, w)[0]          # reference mode: pre_nms_topk=0
        return s.cpu(), l.cpu(), b.cpu()

    for j in range(min(3, images)):
        call(j)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    kept = 0
    for j in range(images):
        kept += call(j)[0].shape[0]
    dt = (time.perf_counter() - t0) / images
    # the same call with the head outputs already on the device (what a real evaluator has: the model ran on the GPU)
    torch.cuda.synchronize()
    per_call = []
    for rep in range(3):          # the first pass over the slices still pays one-time costs (allocator growth): report the last
        per_call = []
        for j in range(images):
            t0 = time.perf_counter()
            s, l, b = D.detect_batch(logits[j:j + 1], reg[j:j + 1], anchors, h, w)[0]
            s.cpu(), l.cpu(), b.cpu()
            per_call.append(time.perf_counter() - t0)
    dt_dev = sum(per_call) / images
    cand = int(((torch.sigmoid(logits).amax(dim=2)) > 0.05).sum().item()) / images
    out = {'workload': 'predict, batch 1, no top-k (reference mode): 800x1333, C=80, A=%d, logit mean %.1f' % (a, mu),
           'candidates_per_image': cand, 'kept_per_image': kept / images,
           'e2e_ms_per_image': dt * 1e3, 'e2e_images_per_s': 1.0 / dt, 'h2d_bytes_per_image': (a * c + a * 4) * 4,
           'device_resident_ms_per_image': dt_dev * 1e3, 'device_resident_images_per_s': 1.0 / dt_dev,
           'device_resident_per_call_ms': [round(x * 1e3, 3) for x in per_call]}
    # same-GPU eager baseline: torch ops + torchvision.ops.batched_nms (the reference's own execution model)
    E.predict(logits[:1], reg[:1], anchors, h, w)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for j in range(images):
        s, l, b = E.predict(logits[j:j + 1], reg[j:j + 1], anchors, h, w)
        s.cpu(), l.cpu(), b.cpu()
    out['gpu_eager_ms_per_image'] = (time.perf_counter() - t0) / images * 1e3
    if cpu_images:
        cl, cr, ca = [x.cpu() for x in (logits[:cpu_images], reg[:cpu_images], anchors)]
        t0 = time.perf_counter()
        for j in range(cpu_images):
            E.predict(cl[j:j + 1], cr[j:j + 1], ca, h, w)
        cdt = (time.perf_counter() - t0) / cpu_images
        out['cpu_baseline'] = {'value': 1.0 / cdt, 'unit': 'images/s', 'ms_per_image': cdt * 1e3, 'cores': torch.get_num_threads(),
                               'kind': 'port', 'sample': '%d images, torch-eager restatement of ResNet.predict + torchvision CPU '
                                                         'batched_nms on the host cores' % cpu_images}
    return out


def reference_predict_times(mus, images=2):
    """The UNMODIFIED ResNet.predict (oracle/_ref snapshot, forward stubbed) on COCO-shaped head outputs, one process per
    device: as written on cuda:0, and on the host cores with the GPUs hidden.  Returns {mu: {key: entry}} (empty without the
    snapshot)."""
    import subprocess
    out = {mu: {} for mu in mus}
    if not os.path.exists(os.path.join(ROOT, 'oracle', '_ref', 'retinanet', 'model.bytecode')):
        return out
    for device, key in (('cuda', 'reference_predict_same_gpu'), ('cpu', 'reference_predict_cpu')):
        try:
            r = subprocess.run([sys.executable, '-m', 'oracle.ref_runner', '--predict-time'] + [str(m) for m in mus] +
                               ['--images', str(images), '--device', device, '--threads', str(min(os.cpu_count() or 1, 16))],
                               cwd=ROOT, capture_output=True, text=True, timeout=900)
            for mu, res in zip(mus, json.loads(r.stdout.strip().splitlines()[-1])):
                out[mu][key] = {'value': res['value'], 'unit': 'images/s', 'ms_per_image': res['ms_per_image'], 'kind': 'reference',
                                'cores': res['threads'], 'kept_per_image': res['kept_per_image'],
                                'sample': '%d predict() calls of the unmodified reference (model.py:494-605, forward stubbed), %s, '
                                          'detections read back with .cpu()' % (res['images'], 'cuda:0 as written' if device == 'cuda'
                                                                                else 'host cores, GPUs hidden')}
        except Exception as e:  # noqa: BLE001
            for mu in mus:
                out[mu][key] = {'unavailable': repr(e)[:200]}
    return out


def measure_nms_h2h(dev, k, iters=10):
    """cldet batched_nms vs torchvision.ops.batched_nms on IDENTICAL (boxes, scores, idxs) device tensors ('the
    kernel to beat on the same box'), K boxes drawn from the decode of a COCO-shaped image.  CUDA-event time per call,
    including each implementation's own sort; results must be identical."""
    import torchvision

    from cl_object_detection_b200 import detect as D
    h, w = 800, 1333
    gen = torch.Generator(device=dev).manual_seed(k)
    anchors = cld.generate_anchors(h, w, dev)[0]
    pick = torch.randperm(anchors.shape[0], device=dev, generator=gen)[:k]
    reg = torch.randn(1, anchors.shape[0], 4, device=dev, generator=gen) * 0.3
    boxes = D.decode_boxes(anchors.unsqueeze(0), reg, clip_to=(h, w))[0][pick].contiguous()
    scores = ((torch.randperm(k, device=dev, generator=gen).float() + 1) / (k + 1)).contiguous()        # distinct
    idxs = torch.randint(0, 80, (k,), device=dev, generator=gen)

    def timeit(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters

    ours = D.batched_nms(boxes, scores, idxs, 0.5)
    theirs = torchvision.ops.batched_nms(boxes, scores, idxs, 0.5)
    ms_ours = timeit(lambda: D.batched_nms(boxes, scores, idxs, 0.5))
    ms_tv = timeit(lambda: torchvision.ops.batched_nms(boxes, scores, idxs, 0.5))
    return {'boxes': k, 'kept': int(ours.shape[0]), 'cldet_ms': ms_ours, 'torchvision_ms': ms_tv, 'speedup': ms_tv / ms_ours,
            'identical_keep': bool(torch.equal(ours, theirs))}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--steps', type=int, default=30)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--images', type=int, default=32)
    ap.add_argument('--mu', type=float, default=-4.0)
    ap.add_argument('--topk', type=int, default=1000)
    ap.add_argument('--classes', type=int, default=80)
    ap.add_argument('--head', action='store_true', help='also time the conv-layout entry (detect_batch_head)')
    print(json.dumps(measure(ap.parse_args())))


if __name__ == '__main__':
    main()
