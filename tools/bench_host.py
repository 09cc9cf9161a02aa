#!/usr/bin/env python
"""Host-side cost of the public loss drop-in, piece by piece (1 GPU): how many microseconds of CPU time one FocalLoss forward /
forward + backward takes to ENQUEUE at the VOC shape (BASELINE config 2, where the kernels take ~50 us and the host is the
limit).  Prints one JSON line."""
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
from cl_object_detection_b200.losses import _check_cuda_f32, _focal_loss_op  # noqa: E402
from cl_object_detection_b200.params import loss_param_args  # noqa: E402


def per_call_us(fn, iters=300, sync_every=50):
    for _ in range(20):
        fn()
    torch.cuda.synchronize()
    total = 0.0
    done = 0
    while done < iters:
        t0 = time.perf_counter()
        for _ in range(sync_every):
            fn()
        total += time.perf_counter() - t0
        done += sync_every
        torch.cuda.synchronize()          # keep the launch queue short: measure enqueue cost, not back-pressure
    return total / done * 1e6


def main():
    dev = torch.device('cuda', 0)
    n, h, w, c, g = 16, 512, 512, 16, 20
    anchors = cld.generate_anchors(h, w, dev)
    a = anchors.shape[1]
    gen = torch.Generator(device=dev).manual_seed(1)
    probs = torch.sigmoid(torch.randn(n, a, c, device=dev, generator=gen) * 2 - 4)
    reg = torch.randn(n, a, 4, device=dev, generator=gen)
    ann = torch.from_numpy(synth_annotations(np.random.default_rng(1), n, g, h, w, c)).to(dev)
    params = cld.HeadParams()
    fl = cld.FocalLoss()
    p = probs.detach().requires_grad_(True)
    r = reg.detach().requires_grad_(True)
    g_rows = torch.full((n,), 1.0 / n, device=dev)
    g_one = torch.ones(1, device=dev)
    hint = fl._hint(n, dev)
    op = _focal_loss_op()
    lp = loss_param_args(params, 0, c)

    def fwd_nograd():
        with torch.no_grad():
            fl(probs, reg, anchors, ann, 0, params)

    def fwd_grad():
        fl(p, r, anchors, ann, 0, params)

    def fwd_bwd():
        out = fl(p, r, anchors, ann, 0, params)
        bg, fg = out['cls_loss']
        torch.autograd.grad([bg, fg, out['reg_loss']], [p, r], [g_rows, g_rows, g_one])

    def fwd_bwd_caller_means():
        out = fl(p, r, anchors, ann, 0, params)
        bg, fg = out['cls_loss']
        torch.autograd.grad(bg.mean() + fg.mean() + out['reg_loss'].mean(), [p, r])

    def raw_op_nograd():
        with torch.no_grad():
            op(probs, reg, anchors, ann, hint, *lp, h, w, False, False, False, [])

    def raw_op_grad():
        op(p, r, anchors, ann, hint, *lp, h, w, False, False, False, [])

    def python_checks():
        for name, t in (('a', probs), ('b', reg), ('c', anchors), ('d', ann)):
            _check_cuda_f32(name, t)
        loss_param_args(params, 0, c)

    res = {'shape': 'VOC 15+1: 16 x 512x512, C=16, A=%d' % a,
           'python_arg_checks_us': per_call_us(python_checks),
           'op_forward_no_grad_us': per_call_us(raw_op_nograd),
           'op_forward_with_autograd_node_us': per_call_us(raw_op_grad),
           'FocalLoss_forward_no_grad_us': per_call_us(fwd_nograd),
           'FocalLoss_forward_us': per_call_us(fwd_grad),
           'FocalLoss_forward_backward_us': per_call_us(fwd_bwd),
           'forward_backward_with_caller_means_us': per_call_us(fwd_bwd_caller_means)}
    print(json.dumps(res))


if __name__ == '__main__':
    main()
