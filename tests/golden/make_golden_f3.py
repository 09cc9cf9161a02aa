"""Golden vectors for SURVEY 8(f) row f3 -- the OTHER users of calc_iou + max in the reference -- produced by running the
unmodified reference (build container only, same CPU shim as make_golden.py):
  IL_method/mas.py        Output_norm.forward (:35-67)      (+ autograd gradients)
  IL_method/prototype.py  ProtoTyper._get_positive (:24-47)
"""
import os
import sys
import types

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import make_golden as mg  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    mg.install_cpu_shim()
    sys.modules['matplotlib.pyplot'].fill = None          # mas.py:3 imports an unused name
    from retinanet.anchors import Anchors
    from IL_method.mas import Output_norm
    from IL_method.prototype import ProtoTyper

    rng = np.random.default_rng(901)
    h, w, C, N, G = 96, 128, 6, 3, 5
    anchors = Anchors()(torch.zeros(1, 3, h, w))
    A = anchors.shape[1]
    cls = rng.uniform(0, 1, (N, A, C)).astype(np.float32)
    reg = rng.normal(0, 1, (N, A, 4)).astype(np.float32)
    ann = mg.make_gt(rng, N, G, h, w, C, empty=())          # these callers crash on images without GT
    tc = torch.from_numpy(cls).requires_grad_(True)
    tr = torch.from_numpy(reg).requires_grad_(True)
    out = Output_norm().forward(tc, tr, anchors, torch.from_numpy(ann))
    (out['regression'] * 0.7 + out['classification'] * 0.3).backward()
    stub = types.SimpleNamespace(thresold=0.45, num_anchors=9)
    pos, tgt = ProtoTyper._get_positive(stub, anchors, torch.from_numpy(ann))
    np.savez_compressed(os.path.join(OUT, 'f3_iou_users.npz'), h=h, w=w, cls=cls, reg=reg, ann=ann,
                        norm_regression=out['regression'].detach().numpy(), norm_classification=out['classification'].detach().numpy(),
                        grad_cls=tc.grad.numpy(), grad_reg=tr.grad.numpy(), proto_threshold=0.45,
                        proto_positive=pos.numpy(), proto_targets=tgt.numpy())
    print('f3 golden written', out['regression'].item(), out['classification'].item(), pos.shape, tgt.shape, int(pos.sum()))


if __name__ == '__main__':
    main()
