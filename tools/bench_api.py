#!/usr/bin/env python
"""Device-resident step time through the PUBLIC Python drop-in (FocalLoss.forward + autograd backward, the caller's mean
reduction) next to the raw C-ABI step, for BASELINE configs 2 and 3: shows what the host-side wrapper costs."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import cl_object_detection_b200 as cld  # noqa: E402
from bench import synth_annotations  # noqa: E402


def run(name, n, h, w, c, g, steps=100):
    dev = torch.device('cuda', 0)
    anchors = cld.generate_anchors(h, w, dev)
    a = anchors.shape[1]
    gen = torch.Generator(device=dev).manual_seed(1)
    probs = torch.sigmoid(torch.randn(n, a, c, device=dev, generator=gen) * 2 - 4)
    reg = torch.randn(n, a, 4, device=dev, generator=gen)
    ann = torch.from_numpy(synth_annotations(np.random.default_rng(1), n, g, h, w, c)).to(dev)
    params = cld.HeadParams()
    fl = cld.FocalLoss()

    def step():
        p = probs.detach().requires_grad_(True)
        r = reg.detach().requires_grad_(True)
        out = fl(p, r, anchors, ann, 0, params)
        bg, fg = out['cls_loss']
        return torch.autograd.grad(bg.mean() + fg.mean() + out['reg_loss'].mean(), [p, r])
    for _ in range(10):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    host_ms = (time.perf_counter() - t0) / steps * 1e3      # time to ENQUEUE a step
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    # the same step captured once in a CUDA graph (the drop-in has no host sync and allocates through torch's allocator)
    sp = probs.detach().clone().requires_grad_(True)
    sr = reg.detach().clone().requires_grad_(True)

    def graph_step():
        out = fl(sp, sr, anchors, ann, 0, params)
        bg, fg = out['cls_loss']
        sp.grad = None
        sr.grad = None
        (bg.mean() + fg.mean() + out['reg_loss'].mean()).backward()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            graph_step()
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        graph_step()
    ref = step()
    graph.replay()
    torch.cuda.synchronize()
    same = bool(torch.equal(sp.grad, ref[0]) and torch.equal(sr.grad, ref[1]))
    e0.record()
    for _ in range(steps):
        graph.replay()
    e1.record()
    torch.cuda.synchronize()
    gms = e0.elapsed_time(e1) / steps
    print(json.dumps({'config': name, 'api_ms_per_step': ms, 'host_enqueue_ms_per_step': host_ms, 'images_per_s': n / (ms * 1e-3),
                      'cuda_graph_ms_per_step': gms, 'cuda_graph_images_per_s': n / (gms * 1e-3), 'graph_grads_identical': same}))


if __name__ == '__main__':
    run('2 VOC 512x512 N=16 C=16', 16, 512, 512, 16, 20)
    run('3 COCO 800x1333 N=16 C=80', 16, 800, 1333, 80, 20)
