#!/usr/bin/env python
"""Where the HOST time of the two public boundaries goes (1 GPU): cProfile of FocalLoss fwd+bwd at the VOC shape and of
predict (detect_batch, one image, no top-k) at ~1.3 k candidates, plus a torch.profiler table of the predict call."""
import cProfile
import io
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import cl_object_detection_b200 as cld  # noqa: E402
from bench import synth_annotations  # noqa: E402
from cl_object_detection_b200 import detect as D  # noqa: E402


def top(pr, n=28):
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats('cumulative').print_stats(n)
    print('\n'.join(l for l in s.getvalue().splitlines() if l.strip())[:6000])


def loss_probe():
    dev = torch.device('cuda', 0)
    n, h, w, c, g = 16, 512, 512, 16, 20
    anchors = cld.generate_anchors(h, w, dev)
    a = anchors.shape[1]
    gen = torch.Generator(device=dev).manual_seed(1)
    p = torch.sigmoid(torch.randn(n, a, c, device=dev, generator=gen) * 2 - 4).requires_grad_(True)
    r = torch.randn(n, a, 4, device=dev, generator=gen).requires_grad_(True)
    ann = torch.from_numpy(synth_annotations(np.random.default_rng(1), n, g, h, w, c)).to(dev)
    params = cld.HeadParams()
    fl = cld.FocalLoss()
    g_rows = torch.full((n,), 1.0 / n, device=dev)
    g_one = torch.ones(1, device=dev)

    def fwd():
        return fl(p, r, anchors, ann, 0, params)

    def fwd_bwd():
        out = fl(p, r, anchors, ann, 0, params)
        bg, fg = out['cls_loss']
        torch.autograd.grad([bg, fg, out['reg_loss']], [p, r], [g_rows, g_rows, g_one])

    for name, fn in (('fwd', fwd), ('fwd_bwd', fwd_bwd)):
        for _ in range(50):
            fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(400):
            fn()
            if i % 40 == 39:
                torch.cuda.synchronize()
        print('loss %s: %.1f us per call (host, incl. a sync every 40)' % (name, (time.perf_counter() - t0) / 400 * 1e6))
    pr = cProfile.Profile()
    pr.enable()
    for i in range(400):
        fwd_bwd()
        if i % 40 == 39:
            torch.cuda.synchronize()
    pr.disable()
    top(pr)


def predict_probe(mu=-10.5):
    dev = torch.device('cuda', 0)
    h, w, c = 800, 1333, 80
    anchors = cld.generate_anchors(h, w, dev)
    gen = torch.Generator(device=dev).manual_seed(1050)
    logits = torch.randn(1, anchors.shape[1], c, device=dev, generator=gen) * 2.0 + mu
    reg = torch.randn(1, anchors.shape[1], 4, device=dev, generator=gen) * 0.3

    def call():
        s, l, b = D.detect_batch(logits, reg, anchors, h, w)[0]
        return s.cpu(), l.cpu(), b.cpu()
    for _ in range(5):
        call()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20):
        call()
    print('predict mu=%.1f: %.3f ms per call, %d kept' % (mu, (time.perf_counter() - t0) / 20 * 1e3, call()[0].shape[0]))
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(20):
        call()
    pr.disable()
    top(pr, 22)
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        for _ in range(5):
            call()
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by='cuda_time_total', row_limit=22, max_name_column_width=60)[:9000])


if __name__ == '__main__':
    loss_probe()
    predict_probe(-10.5)
    predict_probe(-9.5)
