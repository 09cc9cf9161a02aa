#!/usr/bin/env python
"""logits-in fused loss vs the reference's call pattern sigmoid -> FocalLoss -> SigmoidBackward
(BASELINE config 3 shape, 16 x 800x1333, C=80).  CUDA-event timings of the three arrangements."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import json  # noqa: E402

import numpy as np  # noqa: E402
import torch  # noqa: E402

import cl_object_detection_b200 as cld  # noqa: E402
from bench import synth_annotations  # noqa: E402


def timeit(fn, steps=30, warmup=5):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / steps


def main():
    dev = torch.device('cuda', 0)
    h, w, c, n, g = 800, 1333, 80, 16, 20
    anchors = cld.generate_anchors(h, w, dev)
    a = anchors.shape[1]
    gen = torch.Generator(device=dev).manual_seed(1)
    logits = torch.randn(n, a, c, device=dev, generator=gen) * 2 - 4
    reg = torch.randn(n, a, 4, device=dev, generator=gen)
    ann = torch.from_numpy(synth_annotations(np.random.default_rng(1), n, g, h, w, c)).to(dev)
    params = cld.HeadParams()
    fl_p, fl_x = cld.FocalLoss(), cld.FocalLoss(from_logits=True)

    def total(out):
        bg, fg = out['cls_loss']
        return bg.mean() + fg.mean() + out['reg_loss'].mean()

    def probs_path():      # the reference's arrangement with our probability-entry kernels
        x = logits.detach().requires_grad_(True)
        r = reg.detach().requires_grad_(True)
        return torch.autograd.grad(total(fl_p(torch.sigmoid(x), r, anchors, ann, 0, params)), [x, r])

    def logits_path():
        x = logits.detach().requires_grad_(True)
        r = reg.detach().requires_grad_(True)
        return torch.autograd.grad(total(fl_x(x, r, anchors, ann, 0, params)), [x, r])
    t_p, t_x = timeit(probs_path), timeit(logits_path)
    g1, g2 = probs_path(), logits_path()
    err = float(((g1[0] - g2[0]).abs() / (g1[0].abs() + 1e-12 * g1[0].abs().max())).max())
    print(json.dumps({'workload': '16 x 800x1333, C=80: dL/dlogits via autograd', 'sigmoid+probs_entry_ms': t_p,
                      'logits_entry_ms': t_x, 'speedup': t_p / t_x, 'images_per_s_logits_entry': n / (t_x * 1e-3),
                      'max_rel_grad_diff': err}))


if __name__ == '__main__':
    main()
