"""Generate the golden fixtures in this directory by RUNNING THE UNMODIFIED REFERENCE.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

The reference has no tests or golden vectors of its own, so parity is
pinned on outputs of the reference's own code for the path:
  retinanet/anchors.py  Anchors.forward           retinanet/losses.py  calc_iou, FocalLoss.forward (+autograd)
  retinanet/utils.py    BBoxTransform, ClipBoxes   retinanet/model.py   ResNet.predict (forward stubbed)
  IL_method/persuado_label.py  Labeler.predict     torchvision.ops      nms, batched_nms (third party, 0.26.0)
The reference hard-codes cuda:0; a test-side shim (no edits to reference source) makes it run on CPU:
torch.ones/zeros drop the `device=` kwarg and Tensor.cuda() is the identity.
Inputs come from numpy PCG64 streams with fixed seeds and are stored next to the outputs.
"""
import hashlib
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get('CLDET_REFERENCE', '/root/reference')
OUT = os.path.dirname(os.path.abspath(__file__))


def install_cpu_shim():
    for name in ('ones', 'zeros'):
        orig = getattr(torch, name)

        def wrapped(*a, _orig=orig, **k):
            k.pop('device', None)
            return _orig(*a, **k)
        setattr(torch, name, wrapped)
    torch.Tensor.cuda = lambda self, *a, **k: self
    for mod in ('pycocotools', 'pycocotools.coco', 'pycocotools.cocoeval', 'skimage', 'skimage.io', 'skimage.transform',
                'skimage.color', 'matplotlib', 'matplotlib.pyplot', 'cv2'):
        if mod not in sys.modules:
            try:
                __import__(mod)
            except Exception:
                m = types.ModuleType(mod)
                m.COCO = object
                m.COCOeval = object
                sys.modules[mod] = m
    sys.path.insert(0, REF)


class Params:
    """Duck type of preprocessing/params.py Params: params[key] (None if absent) and params.states."""

    def __init__(self, num_past_class, **kw):
        self.d = dict(alpha=0.25, gamma=2.0, distill=False, enhance_on_new=False, ignore_past_class=False,
                      new_ignore_past_class=False, decrease_positive_by_IOU=False, decrease_positive=1.0,
                      persuado_label=False)
        self.d.update(kw)
        self.states = [{'num_past_class': n} for n in num_past_class]

    def __getitem__(self, k):
        return self.d.get(k, None)


def make_gt(rng, n_img, gmax, height, width, num_classes, empty=(), min_side=16.0, labels=None):
    ann = np.full((n_img, gmax, 5), -1.0, dtype=np.float32)
    for j in range(n_img):
        if j in empty:
            continue
        g = int(rng.integers(1, gmax + 1))
        x1 = rng.uniform(0, 0.7 * width, g)
        y1 = rng.uniform(0, 0.7 * height, g)
        w = rng.uniform(min_side, 0.45 * width + min_side, g)
        h = rng.uniform(min_side, 0.45 * height + min_side, g)
        lab = rng.integers(0, num_classes, g) if labels is None else labels(rng, g)
        rows = np.stack([x1, y1, x1 + w, y1 + h, lab.astype(np.float64)], 1).astype(np.float32)
        # scatter the valid rows among the pad rows to exercise order-preserving compaction
        pos = np.sort(rng.choice(gmax, g, replace=False))
        ann[j, pos] = rows
    return ann


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    install_cpu_shim()
    import torchvision
    from retinanet.anchors import Anchors
    from retinanet.losses import FocalLoss, calc_iou
    from retinanet.utils import BBoxTransform, ClipBoxes
    from retinanet import model as ref_model
    from IL_method.persuado_label import Labeler

    torch.manual_seed(0)
    torch.set_num_threads(4)
    meta = dict(torch=torch.__version__, torchvision=torchvision.__version__, device='cpu')

    # ---------------- anchors ----------------
    anc = {}
    for (h, w) in [(64, 96), (128, 160), (33, 70), (512, 512), (800, 1333), (1333, 1333), (608, 1024)]:
        a = Anchors()(torch.zeros(1, 3, h, w)).numpy()
        anc[f'{h}x{w}_shape'] = np.array(a.shape)
        anc[f'{h}x{w}_sha256'] = np.array(sha(a))
        if a.shape[1] < 5000:
            anc[f'{h}x{w}'] = a
        else:
            idx = np.random.default_rng(h * 10000 + w).choice(a.shape[1], 2048, replace=False)
            idx.sort()
            anc[f'{h}x{w}_idx'] = idx
            anc[f'{h}x{w}_rows'] = a[0, idx]
    np.savez_compressed(os.path.join(OUT, 'anchors.npz'), **anc)

    # ---------------- calc_iou ----------------
    rng = np.random.default_rng(101)
    a = Anchors()(torch.zeros(1, 3, 128, 160)).numpy()[0]
    g = make_gt(rng, 1, 9, 128, 160, 5)[0]
    g = g[g[:, 4] != -1][:, :4]
    g = np.concatenate([g, a[1234:1235], np.array([[5, 5, 5, 9]], np.float32)])   # exact match + zero-area box
    iou = calc_iou(torch.from_numpy(a), torch.from_numpy(g)).numpy()
    np.savez_compressed(os.path.join(OUT, 'calc_iou.npz'), a=a, b=g, iou=iou)

    # ---------------- FocalLoss fwd + autograd bwd ----------------
    def run_focal(name, h, w, C, N, gmax, seed, cur_state, params, empty=(1,), labels=None, progress=-1, sigma=3.0):
        rng = np.random.default_rng(seed)
        anchors = Anchors()(torch.zeros(1, 3, h, w))
        A = anchors.shape[1]
        logits = rng.normal(-3.0, sigma, (N, A, C)).astype(np.float32)
        cls = (1.0 / (1.0 + np.exp(-logits.astype(np.float64)))).astype(np.float32)
        # force a few exact clamp-boundary and out-of-band values
        flat = cls.reshape(-1)
        flat[:6] = np.array([1e-4, 1 - 1e-4, 0.0, 1.0, 5e-5, 0.99995], np.float32)
        reg = rng.normal(0, 1, (N, A, 4)).astype(np.float32)
        ann = make_gt(rng, N, gmax, h, w, C, empty=empty, labels=labels)
        wb = rng.uniform(0.2, 1.0, N)
        wf = rng.uniform(0.2, 1.0, N)
        wr, we = 0.7, 0.3
        tc = torch.from_numpy(cls).requires_grad_(True)
        tr = torch.from_numpy(reg).requires_grad_(True)
        out = FocalLoss().forward(tc, tr, anchors, torch.from_numpy(ann), cur_state, params, progress)
        bg, fg = out['cls_loss']
        loss = (bg * torch.from_numpy(wb).float()).sum() + (fg * torch.from_numpy(wf).float()).sum() + wr * out['reg_loss'].sum()
        if 'enhance_on_new_loss' in out:
            loss = loss + we * out['enhance_on_new_loss']
        loss.backward()
        # assignment ground truth straight from the reference's building blocks (losses.py:309-330)
        state = np.full((N, A), 3, np.uint8)
        argmax = np.full((N, A), -1, np.int32)
        for j in range(N):
            b = torch.from_numpy(ann[j])
            b = b[b[:, 4] != -1]
            if b.shape[0] == 0:
                continue
            im, ia = torch.max(calc_iou(anchors[0], b[:, :4]), dim=1)
            s = np.full(A, 2, np.uint8)
            s[torch.lt(im, 0.4).numpy()] = 0
            s[torch.ge(im, 0.5).numpy()] = 1
            state[j], argmax[j] = s, ia.numpy().astype(np.int32)
        d = dict(h=h, w=w, cur_state=cur_state, cls=cls, reg=reg, ann=ann, wb=wb, wf=wf, wr=wr, we=we, progress=progress,
                 bg=bg.detach().numpy(), fg=fg.detach().numpy(), reg_loss=out['reg_loss'].detach().numpy(),
                 grad_cls=tc.grad.numpy(), grad_reg=(tr.grad.numpy() if tr.grad is not None else np.zeros_like(reg)), state=state, argmax=argmax,
                 params_keys=np.array(list(params.d.keys())), params_vals=np.array([float(v) for v in params.d.values()]),
                 num_past_class=np.array([s['num_past_class'] for s in params.states]))
        if 'bg_masks' in out:
            d['bg_masks'] = out['bg_masks'].numpy()
        if 'enhance_on_new_loss' in out:
            d['enhance_on_new_loss'] = out['enhance_on_new_loss'].detach().numpy()
        np.savez_compressed(os.path.join(OUT, f'focal_{name}.npz'), **d)
        print(name, 'npos', [(int((state[j] == 1).sum())) for j in range(N)], 'bg', d['bg'], 'fg', d['fg'], 'reg', d['reg_loss'])

    run_focal('state0_voc', 128, 160, 20, 3, 8, 201, 0, Params([0]))
    run_focal('state0_allvalid', 64, 96, 6, 2, 4, 202, 0, Params([0]), empty=())
    run_focal('state0_allempty', 64, 96, 6, 2, 4, 203, 0, Params([0]), empty=(0, 1))
    run_focal('state0_gamma15', 64, 96, 6, 2, 4, 204, 0, Params([0], gamma=1.5, alpha=0.4), empty=())
    # incremental state 1 of scenario 15+1: labels: first rows new class (4), others pseudo (old classes 0..3)
    pl = lambda rng, g: np.concatenate([[4], rng.integers(0, 4, g - 1)]) if g > 0 else np.zeros(0)
    run_focal('il_default_pseudo', 64, 96, 5, 3, 6, 205, 1, Params([0, 4], persuado_label=True), labels=pl, progress=0.5)
    run_focal('il_ignore_past', 64, 96, 5, 3, 6, 206, 1, Params([0, 4], ignore_past_class=True), labels=pl)
    run_focal('il_new_ignore_past', 64, 96, 5, 3, 6, 207, 1, Params([0, 4], ignore_past_class=True, new_ignore_past_class=True),
              labels=pl, sigma=2.0)
    run_focal('il_distill_enhance', 64, 96, 5, 3, 6, 208, 1, Params([0, 4], distill=True, enhance_on_new=True), labels=pl, empty=())
    run_focal('il_decrease_positive', 64, 96, 5, 3, 6, 209, 1, Params([0, 4], decrease_positive=0.8), labels=pl)
    run_focal('il_decrease_by_iou', 64, 96, 5, 3, 6, 210, 1, Params([0, 4], decrease_positive_by_IOU=True), labels=pl)
    run_focal('il_all_flags', 64, 96, 5, 3, 6, 211, 2,
              Params([0, 3, 4], ignore_past_class=True, new_ignore_past_class=True, enhance_on_new=True, distill=True,
                     decrease_positive_by_IOU=True), labels=pl, empty=())

    # ---------------- decode / clip ----------------
    rng = np.random.default_rng(301)
    anchors = Anchors()(torch.zeros(1, 3, 128, 160))
    reg = rng.normal(0, 0.7, (2, anchors.shape[1], 4)).astype(np.float32)
    dec = BBoxTransform()(anchors, torch.from_numpy(reg))
    clipped = ClipBoxes()(dec.clone(), torch.zeros(2, 3, 128, 160))
    np.savez_compressed(os.path.join(OUT, 'decode.npz'), h=128, w=160, reg=reg, decoded=dec.numpy(), clipped=clipped.numpy())

    # ---------------- predict (model.py:494-550) with forward stubbed ----------------
    class StubModel(ref_model.ResNet):
        def __init__(self):
            torch.nn.Module.__init__(self)
            self.regressBoxes = BBoxTransform()
            self.clipBoxes = ClipBoxes()

        def forward(self, img_batch, return_feat=False, return_anchor=True, enable_act=False):
            return self._cls, self._reg, self._anchors

    def run_predict(name, h, w, C, seed, mu, sigma_reg=0.3):
        rng = np.random.default_rng(seed)
        anchors = Anchors()(torch.zeros(1, 3, h, w))
        A = anchors.shape[1]
        logits = rng.normal(mu, 2.0, (1, A, C)).astype(np.float32)
        logits.reshape(-1)[:C] = 30.0            # a row of saturated ties -> first-index argmax
        reg = rng.normal(0, sigma_reg, (1, A, 4)).astype(np.float32)
        m = StubModel()
        m._cls, m._reg, m._anchors = torch.from_numpy(logits), torch.from_numpy(reg), anchors
        with torch.no_grad():
            scores, labels, boxes = m.predict(torch.zeros(1, 3, h, w))
            probs = torch.sigmoid(torch.from_numpy(logits))
            s2, b2, l2 = Labeler.predict(m, torch.zeros(1, 3, h, w), probs, torch.from_numpy(reg), anchors)
        assert torch.equal(scores, s2) and torch.equal(boxes, b2) and torch.equal(labels, l2)
        np.savez_compressed(os.path.join(OUT, f'predict_{name}.npz'), h=h, w=w, logits=logits, reg=reg, probs=probs.numpy(),
                            scores=scores.numpy(), labels=labels.numpy(), boxes=boxes.numpy(), device_rule='cpu')
        print(name, 'kept', scores.shape[0])

    run_predict('trick', 128, 160, 20, 401, -6.0)       # few candidates: coordinate-trick branch (numel <= 4000 on CPU)
    run_predict('vanilla', 128, 160, 20, 402, -4.0)     # >1000 candidates: vanilla branch on CPU
    run_predict('none', 64, 96, 6, 403, -14.0)          # one candidate row only (the forced 30.0 row)

    # ---------------- torchvision nms / batched_nms (third party oracle of record) ----------------
    from torchvision.ops import nms as tv_nms
    from torchvision.ops.boxes import _batched_nms_coordinate_trick, _batched_nms_vanilla
    rng = np.random.default_rng(501)
    d = {}
    for t, (K, ncls, span) in enumerate([(300, 5, 200.0), (1500, 20, 300.0), (64, 1, 60.0), (1, 1, 10.0), (129, 80, 50.0)]):
        x1 = rng.uniform(0, span, K)
        y1 = rng.uniform(0, span, K)
        bw = rng.uniform(4, span * 0.4, K)
        bh = rng.uniform(4, span * 0.4, K)
        boxes = np.stack([x1, y1, x1 + bw, y1 + bh], 1).astype(np.float32)
        if K > 8:
            boxes[5] = boxes[3]                   # exact duplicates
            boxes[7, 2] = boxes[7, 0] - 1.0       # negative-width box
        scores = rng.uniform(0.05, 1.0, K).astype(np.float32)
        idxs = rng.integers(0, ncls, K)
        d[f'boxes{t}'], d[f'idxs{t}'] = boxes, idxs
        tb, ti = torch.from_numpy(boxes), torch.from_numpy(idxs)
        d[f'scores{t}_unique'] = scores
        d[f'vanilla{t}'] = _batched_nms_vanilla(tb, torch.from_numpy(scores), ti, 0.5).numpy()   # no ties: order well defined
        d[f'trick{t}_unique'] = _batched_nms_coordinate_trick(tb, torch.from_numpy(scores), ti, 0.5).numpy()
        tied = np.round(scores * 20) / 20          # many exact score ties
        tied = tied.astype(np.float32)
        d[f'scores{t}_tied'] = tied
        d[f'nms{t}'] = tv_nms(tb, torch.from_numpy(tied), 0.5).numpy()
        d[f'nms{t}_thr03'] = tv_nms(tb, torch.from_numpy(tied), 0.3).numpy()
        d[f'trick{t}'] = _batched_nms_coordinate_trick(tb, torch.from_numpy(tied), ti, 0.5).numpy()
    np.savez_compressed(os.path.join(OUT, 'nms.npz'), **d)

    np.savez_compressed(os.path.join(OUT, 'meta.npz'), **{k: np.array(v) for k, v in meta.items()})
    print('done', meta)


if __name__ == '__main__':
    main()
