#!/usr/bin/env python
"""SURVEY 8f row f2 measurement: fused head-distillation terms vs the eager torch statements of losses.py:705-737
(VOC 15+1 shape: N=16, 512x512, C=16, P=15), forward + backward, CUDA events."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))

import torch  # noqa: E402

import cl_object_detection_b200 as cld  # noqa: E402
from make_golden_f2 import il_loss_distill_terms  # noqa: E402


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
    n, c, p = 16, 16, 15
    a = cld.num_anchors(512, 512)
    gen = torch.Generator(device=dev).manual_seed(2)
    cls = torch.randn(n, a, c, device=dev, generator=gen) * 2 - 4
    prev = cls[:, :, :p] + torch.randn(n, a, p, device=dev, generator=gen) * 0.5
    reg = torch.randn(n, a, 4, device=dev, generator=gen)
    preg = reg + torch.randn(n, a, 4, device=dev, generator=gen) * 0.5
    bg = torch.rand(n, a, device=dev, generator=gen) > 0.02

    def eager():
        x = cls.detach().requires_grad_(True)
        r = reg.detach().requires_grad_(True)
        lc, lr = il_loss_distill_terms(x, r, prev, preg, bg, p, False, False)
        return torch.autograd.grad(lc + lr, [x, r])

    def fused():
        x = cls.detach().requires_grad_(True)
        r = reg.detach().requires_grad_(True)
        out = cld.head_distillation(x, r, prev, preg, bg)
        return torch.autograd.grad(out['dist_cls_loss'] + out['dist_reg_loss'], [x, r])
    te, tf = timeit(eager), timeit(fused)
    ge, gf = eager(), fused()
    err = float(((ge[0] - gf[0]).abs().max() / ge[0].abs().max()))
    print(json.dumps({'workload': 'IL_Loss distillation terms, N=16, A=%d, C=16, P=15, fwd+bwd' % a, 'eager_torch_ms': te,
                      'fused_ms': tf, 'speedup': te / tf, 'max_grad_err_rel_to_scale': err}))


if __name__ == '__main__':
    main()
