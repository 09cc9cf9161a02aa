#!/usr/bin/env python
"""Device time of the four fused-loss entries through the C ABI alone (ctypes, preallocated outputs, CUDA events): the
concatenated layout (cldet_focal_loss) and the conv layout (cldet_focal_loss_head), each with probabilities and with logits
in, on BASELINE config 3's shape (16 x 800x1333, C=80).  Each call = GT-centric assignment launch + the fused loss kernel.
Works with A/B builds: CLDET_LIBRARY=build/variants/libcldet_NAME.so python tools/bench_kernels.py.  One JSON line."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import cl_object_detection_b200 as cld  # noqa: E402
from bench import peak_hbm, synth_annotations  # noqa: E402
from cl_object_detection_b200 import _lib  # noqa: E402
from cl_object_detection_b200.params import to_loss_params  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--steps', type=int, default=50)
    ap.add_argument('--only', default='', help='comma list of cat_probs,cat_logits,head_probs,head_logits')
    args = ap.parse_args()
    dev = torch.device('cuda', 0)
    torch.cuda.set_device(dev)
    n, h, w, c = 16, 800, 1333, 80
    lib = _lib.load()
    anchors = cld.generate_anchors(h, w, dev)
    a = anchors.shape[1]
    gen = torch.Generator(device=dev).manual_seed(7)
    ann = torch.from_numpy(synth_annotations(np.random.default_rng(7), n, 20, h, w, c)).to(dev)
    params = cld.HeadParams()
    weights = torch.full((4, n), 1.0 / n, device=dev)
    baked = torch.empty_like(weights)
    losses = torch.empty((4, n), device=dev)
    meta = torch.empty((n, a), dtype=torch.int32, device=dev)
    npos = torch.zeros(n, dtype=torch.int32, device=dev)
    nvalid = torch.zeros(n, dtype=torch.int32, device=dev)
    ws = torch.zeros(lib.cldet_focal_loss_workspace_bytes(n, a), dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    peak, _ = peak_hbm()
    only = set(filter(None, args.only.split(',')))
    out = {'library': os.environ.get('CLDET_LIBRARY', 'in-tree')}

    def timeit(fn):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(args.steps):
            fn()
        t1.record()
        torch.cuda.synchronize()
        return t0.elapsed_time(t1) / args.steps

    def report(name, ms, elems):
        gbs = (8.0 * elems + 20.0 * n * a) / (ms * 1e-3) / 1e9
        out[name] = {'ms': round(ms, 5), 'GBps_algorithmic': round(gbs, 1), 'frac_of_measured_peak': round(gbs / peak, 4)}

    for logits in (False, True):
        name = 'cat_logits' if logits else 'cat_probs'
        if only and name not in only:
            continue
        x = torch.randn(n, a, c, device=dev, generator=gen) * 2 - 4
        if not logits:
            x = torch.sigmoid(x)
        reg = torch.randn(n, a, 4, device=dev, generator=gen)
        g, gr = torch.empty_like(x), torch.empty_like(reg)
        lp = to_loss_params(params, 0, c)
        lp.cls_is_logits = int(logits)
        lp.image_height, lp.image_width = h, w

        def step():
            _lib.check(lib.cldet_focal_loss(x.data_ptr(), reg.data_ptr(), anchors.data_ptr(), ann.data_ptr(), n, a, c, ann.shape[1],
                                            lp, weights.data_ptr(), baked.data_ptr(), g.data_ptr(), gr.data_ptr(), losses.data_ptr(),
                                            meta.data_ptr(), None, npos.data_ptr(), nvalid.data_ptr(), None, None, ws.data_ptr(),
                                            ws.numel(), st))
        report(name, timeit(step), x.numel())
        del x, reg, g, gr
        torch.cuda.empty_cache()

    shapes = [((h + 2 ** l - 1) // 2 ** l, (w + 2 ** l - 1) // 2 ** l) for l in range(3, 8)]
    for logits in (False, True):
        name = 'head_logits' if logits else 'head_probs'
        if only and name not in only:
            continue
        cls_lv = [torch.randn(n, 9 * c, hl, wl, device=dev, generator=gen) * 2 - 4 for hl, wl in shapes]
        if not logits:
            cls_lv = [torch.sigmoid(t) for t in cls_lv]
        reg_lv = [torch.randn(n, 36, hl, wl, device=dev, generator=gen) for hl, wl in shapes]
        gcls = [torch.empty_like(t) for t in cls_lv]
        greg = [torch.empty_like(t) for t in reg_lv]
        lp = to_loss_params(params, 0, c)
        lp.cls_is_logits = int(logits)
        lp.image_height, lp.image_width = h, w
        pc, pr, pgc, pgr = (_lib.ptr_array(t) for t in (cls_lv, reg_lv, gcls, greg))

        def step():
            _lib.check(lib.cldet_focal_loss_head(pc, pr, 5, h, w, anchors.data_ptr(), ann.data_ptr(), n, c, ann.shape[1], lp,
                                                 weights.data_ptr(), baked.data_ptr(), pgc, pgr, losses.data_ptr(), meta.data_ptr(),
                                                 None, npos.data_ptr(), nvalid.data_ptr(), None, None, ws.data_ptr(), ws.numel(), st))
        report(name, timeit(step), sum(t.numel() for t in cls_lv))
        del cls_lv, reg_lv, gcls, greg
        torch.cuda.empty_cache()
    print(json.dumps(out))


if __name__ == '__main__':
    main()
