#!/usr/bin/env python
"""Loss-path timing over the BASELINE.json configs (1, 2, 3, 5) on one GPU: device-resident steps through the fused C-ABI call
(assign + loss/grad + backward check), CUDA events, one JSON line per config.  Configs 1-2 are L2-resident / latency-bound
(reported, not judged against the HBM roofline); config 5 is the dense-GT stress shape."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import cl_object_detection_b200 as cld  # noqa: E402
from cl_object_detection_b200 import _lib  # noqa: E402
from cl_object_detection_b200.params import to_loss_params  # noqa: E402
from bench import synth_annotations  # noqa: E402

CONFIGS = [
    dict(name='1 VOC state0', n=2, h=512, w=512, c=20, g=10, state=0, past=(0,)),
    dict(name='2 VOC 15_1 state1 + pseudo GT', n=16, h=512, w=512, c=16, g=20, state=1, past=(0, 15)),
    dict(name='2b same, C=20', n=16, h=512, w=512, c=20, g=20, state=1, past=(0, 15)),
    dict(name='3 COCO-shaped', n=16, h=800, w=1333, c=80, g=20, state=0, past=(0,)),
    dict(name='5 dense GT stress', n=8, h=1333, w=1333, c=80, g=100, state=0, past=(0,)),
]


def main():
    dev = torch.device('cuda', 0)
    lib = _lib.load()
    peak = 6544.7
    for cfg in CONFIGS:
        n, h, w, c, g = cfg['n'], cfg['h'], cfg['w'], cfg['c'], cfg['g']
        anchors = cld.generate_anchors(h, w, dev)
        a = anchors.shape[1]
        gen = torch.Generator(device=dev).manual_seed(1)
        probs = torch.sigmoid(torch.randn(n, a, c, device=dev, generator=gen) * 2 - 4)
        reg = torch.randn(n, a, 4, device=dev, generator=gen)
        ann_np = synth_annotations(np.random.default_rng(1), n, g, h, w, c, empty=(0,) if cfg['g'] < 100 else ())
        if cfg['g'] == 100:
            ann_np = synth_annotations(np.random.default_rng(1), n, g, h, w, c, empty=())
            for j in range(n):                       # exactly 100 boxes per image
                m = ann_np[j, :, 4] == -1
                ann_np[j, m] = ann_np[j, ~m][:m.sum()] if (~m).sum() >= m.sum() else ann_np[j, ~m][np.arange(m.sum()) % (~m).sum()]
        ann = torch.from_numpy(ann_np).to(dev)
        params = cld.HeadParams(list(cfg['past']), persuado_label=cfg['state'] > 0)
        lp = to_loss_params(params, cfg['state'], c)
        lp.image_height, lp.image_width = h, w
        weights = torch.full((4, n), 1.0 / n, device=dev)
        baked = weights.clone()
        gcls, greg = torch.empty_like(probs), torch.empty_like(reg)
        losses = torch.empty((4, n), device=dev)
        meta = torch.empty((n, a), dtype=torch.int32, device=dev)
        npos = torch.empty(n, dtype=torch.int32, device=dev)
        nvalid = torch.empty(n, dtype=torch.int32, device=dev)
        ws = torch.zeros(lib.cldet_focal_loss_workspace_bytes(n, a), dtype=torch.uint8, device=dev)
        st = torch.cuda.current_stream().cuda_stream

        def step():
            _lib.check(lib.cldet_focal_loss(probs.data_ptr(), reg.data_ptr(), anchors.data_ptr(), ann.data_ptr(), n, a, c, g, lp,
                                            weights.data_ptr(), baked.data_ptr(), gcls.data_ptr(), greg.data_ptr(), losses.data_ptr(),
                                            meta.data_ptr(), None, npos.data_ptr(), nvalid.data_ptr(), None, None,
                                            ws.data_ptr(), ws.numel(), st))
            _lib.check(lib.cldet_focal_loss_reweight(probs.data_ptr(), reg.data_ptr(), anchors.data_ptr(), ann.data_ptr(), n, a,
                                                     c, g, lp, weights.data_ptr(), baked.data_ptr(), gcls.data_ptr(),
                                                     greg.data_ptr(), meta.data_ptr(), None, npos.data_ptr(), ws.data_ptr(),
                                                     ws.numel(), st))
        for _ in range(10):
            step()
        torch.cuda.synchronize()
        steps = 100
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        path_bytes = n * (8 * a * c + 48 * a + 20 * g)
        print(json.dumps({'config': cfg['name'], 'N': n, 'HxW': '%dx%d' % (h, w), 'C': c, 'A': a, 'G': g, 'ms_per_step': ms,
                          'images_per_s': n / (ms * 1e-3), 'npos_mean': float(npos.float().mean()),
                          'path_GBps': path_bytes / (ms * 1e-3) / 1e9, 'path_frac_of_measured_peak': path_bytes / (ms * 1e-3) / 1e9 / peak,
                          'working_set_MB': (probs.numel() * 8 + reg.numel() * 8) / 1e6}))


if __name__ == '__main__':
    main()
