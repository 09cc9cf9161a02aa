#!/usr/bin/env python
"""SURVEY 8(f) row f1, second half: head -> loss segment on the conv outputs' own layout.

Times, for a COCO-shaped batch (16 x 800x1333, C=80), forward + backward down to the gradients of the ten conv outputs of
  ref-layout : the reference's layout ops (per-level permute + contiguous + view, torch.cat over the levels,
               retinanet/model.py:125-130, 170-184, 472-474; with --logits also classifier_act = Sigmoid) feeding the
               concatenated-layout drop-in FocalLoss, autograd carrying the gradients back through those ops;
  head-layout: FocalLoss.forward_head on the conv outputs as they are.
Prints one JSON line."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import cl_object_detection_b200 as cld  # noqa: E402
from bench import synth_annotations  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--steps', type=int, default=30)
    ap.add_argument('--images', type=int, default=16)
    ap.add_argument('--classes', type=int, default=80)
    ap.add_argument('--height', type=int, default=800)
    ap.add_argument('--width', type=int, default=1333)
    ap.add_argument('--logits', action='store_true')
    args = ap.parse_args()
    dev = torch.device('cuda', 0)
    torch.cuda.set_device(dev)
    h, w, c, n = args.height, args.width, args.classes, args.images
    anchors = cld.generate_anchors(h, w, dev)
    gen = torch.Generator(device=dev).manual_seed(5000)
    shapes = [((h + 2 ** l - 1) // 2 ** l, (w + 2 ** l - 1) // 2 ** l) for l in range(3, 8)]
    cls_lv = [torch.randn(n, 9 * c, hl, wl, device=dev, generator=gen) * 2 - 4 for hl, wl in shapes]
    if not args.logits:
        cls_lv = [torch.sigmoid(t) for t in cls_lv]
    reg_lv = [torch.randn(n, 36, hl, wl, device=dev, generator=gen) for hl, wl in shapes]
    ann = torch.from_numpy(synth_annotations(np.random.default_rng(5000), n, 20, h, w, c)).to(dev)
    params = cld.HeadParams()
    fl_probs = cld.FocalLoss()
    fl_head = cld.FocalLoss(from_logits=args.logits)

    def cat_layout(levels, per_anchor):
        return torch.cat([t.permute(0, 2, 3, 1).contiguous().view(t.shape[0], -1, per_anchor) for t in levels], dim=1)

    def leaves():
        return [t.detach().requires_grad_(True) for t in cls_lv], [t.detach().requires_grad_(True) for t in reg_lv]

    def ref_step():
        cl, rl = leaves()
        cls = cat_layout(cl, c)
        if args.logits:
            cls = torch.sigmoid(cls)
        out = fl_probs(cls, cat_layout(rl, 4), anchors, ann, 0, params)
        (out['cls_loss'][0].mean() + out['cls_loss'][1].mean() + out['reg_loss'].mean()).backward()
        return cl[0].grad

    def head_step():
        cl, rl = leaves()
        out = fl_head.forward_head(cl, rl, anchors, ann, 0, params, (h, w))
        (out['cls_loss'][0].mean() + out['cls_loss'][1].mean() + out['reg_loss'].mean()).backward()
        return cl[0].grad

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

    # the C ABI alone (scatter + fused kernel, preallocated outputs): the device time without autograd / allocator work
    from cl_object_detection_b200 import _lib
    from cl_object_detection_b200.params import to_loss_params
    lib = _lib.load()
    a = anchors.shape[1]
    lp = to_loss_params(params, 0, c)
    lp.cls_is_logits = int(args.logits)
    lp.image_height, lp.image_width = h, w
    weights = torch.full((4, n), 1.0 / n, device=dev)
    baked = torch.empty_like(weights)
    gcls = [torch.empty_like(t) for t in cls_lv]
    greg = [torch.empty_like(t) for t in reg_lv]
    losses = torch.empty((4, n), device=dev)
    meta = torch.empty((n, a), dtype=torch.int32, device=dev)
    npos = torch.zeros(n, dtype=torch.int32, device=dev)
    nvalid = torch.zeros(n, dtype=torch.int32, device=dev)
    ws = torch.zeros(lib.cldet_focal_loss_workspace_bytes(n, a), dtype=torch.uint8, device=dev)
    pc, pr, pgc, pgr = (_lib.ptr_array(x) for x in (cls_lv, reg_lv, gcls, greg))
    st = torch.cuda.current_stream().cuda_stream

    def raw_step():
        _lib.check(lib.cldet_focal_loss_head(pc, pr, 5, h, w, anchors.data_ptr(), ann.data_ptr(), n, c, ann.shape[1], lp,
                                             weights.data_ptr(), baked.data_ptr(), pgc, pgr, losses.data_ptr(), meta.data_ptr(),
                                             None, npos.data_ptr(), nvalid.data_ptr(), None, None, ws.data_ptr(), ws.numel(), st))

    g_ref, g_head = ref_step(), head_step()
    diff = ((g_ref - g_head).abs() / (g_ref.abs() + 1e-30)).max().item()
    ms_ref, ms_head, ms_raw = timeit(ref_step), timeit(head_step), timeit(raw_step)
    elems = sum(t.numel() for t in cls_lv)
    print(json.dumps({'workload': '%d x %dx%d, C=%d, %s in: head outputs -> losses -> gradients of the head outputs'
                                  % (n, h, w, c, 'logits' if args.logits else 'probabilities'),
                      'ref_layout_ms': ms_ref, 'head_layout_ms': ms_head, 'speedup': ms_ref / ms_head,
                      'images_per_s_head_layout': n / (ms_head * 1e-3),
                      'head_layout_c_abi_ms': ms_raw, 'head_layout_c_abi_GBps_algorithmic': 8.0 * elems / (ms_raw * 1e-3) / 1e9,
                      'max_rel_grad_diff_level3': diff}))


if __name__ == '__main__':
    main()
