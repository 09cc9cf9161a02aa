"""TEST / BASELINE INFRASTRUCTURE -- never imported by the product path.

Runs the UNMODIFIED reference (the byte-code snapshot oracle/_ref, see oracle/build_ref.py) on the HOST cores:
FocalLoss.forward + autograd backward (retinanet/losses.py:252-452) with the caller's reduction (IL_Loss: .mean() of every
term, losses.py:584-588).  The reference hard-codes cuda:0 (`torch.ones(..., device=torch.device('cuda:0'))`, `.cuda()`);
SURVEY 8(c)'s shim makes it run on a CPU without touching its code: torch.ones / torch.zeros drop the `device=` keyword
and Tensor.cuda() is the identity.  The shim patches torch globally, so this module is meant to run in its OWN process:

    python -m oracle.ref_runner --images 2 --frac 0.25 --steps 1 --warmup 0 --threads 16      -> one JSON line
"""
import argparse
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.path.join(HERE, '_ref')


def install_cpu_shim():
    import torch
    if getattr(torch, '_cldet_ref_shim', False):
        return
    for name in ('ones', 'zeros'):
        orig = getattr(torch, name)

        def wrapped(*a, _orig=orig, **k):
            k.pop('device', None)
            return _orig(*a, **k)
        setattr(torch, name, wrapped)
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch._cldet_ref_shim = True


EXT = '.bytecode'


def import_snapshot():
    """Import the snapshot's modules as the package `retinanet` (byte-code files under a neutral extension, loaded with
    importlib's SourcelessFileLoader); raises if the snapshot was never built."""
    import importlib.machinery
    import importlib.util
    if 'retinanet.losses' in sys.modules:
        return sys.modules['retinanet.losses'], sys.modules['retinanet.anchors']
    pkg_dir = os.path.join(REF, 'retinanet')
    if not os.path.exists(os.path.join(pkg_dir, 'losses' + EXT)):
        raise RuntimeError('oracle/_ref is missing: run `python -m oracle.build_ref` where /root/reference exists')

    def load(name, fname, is_pkg=False):
        path = os.path.join(pkg_dir, fname + EXT)
        loader = importlib.machinery.SourcelessFileLoader(name, path)
        spec = importlib.util.spec_from_file_location(name, path, loader=loader, submodule_search_locations=[pkg_dir] if is_pkg else None)
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        loader.exec_module(mod)
        return mod
    load('retinanet', '__init__', is_pkg=True)
    return load('retinanet.losses', 'losses'), load('retinanet.anchors', 'anchors')


def import_snapshot_model():
    """retinanet.model of the snapshot (ResNet.predict, model.py:494-605) with the modules it imports."""
    import importlib.machinery
    import importlib.util
    import_snapshot()
    if 'retinanet.model' in sys.modules:
        return sys.modules['retinanet.model']

    def load(name, rel, is_pkg=False):
        path = os.path.join(REF, rel + EXT)
        loader = importlib.machinery.SourcelessFileLoader(name, path)
        spec = importlib.util.spec_from_file_location(name, path, loader=loader,
                                                      submodule_search_locations=[os.path.dirname(path)] if is_pkg else None)
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        loader.exec_module(mod)
        return mod
    load('preprocessing', 'preprocessing/__init__', is_pkg=True)
    load('preprocessing.debug', 'preprocessing/debug')
    load('retinanet.utils', 'retinanet/utils')
    return load('retinanet.model', 'retinanet/model')


def stub_model():
    """The reference's ResNet with its constructor and forward stubbed (no backbone): predict() itself -- everything after
    self.forward, model.py:507-605 -- is the unmodified reference code.  Same stub as tests/golden/make_golden.py."""
    import torch
    ref_model = import_snapshot_model()
    utils = sys.modules['retinanet.utils']

    class StubModel(ref_model.ResNet):
        def __init__(self):
            torch.nn.Module.__init__(self)
            self.regressBoxes = utils.BBoxTransform()
            self.clipBoxes = utils.ClipBoxes()

        def forward(self, img_batch, return_feat=False, return_anchor=True, enable_act=False):
            return self._cls, self._reg, self._anchors
    return StubModel()


def run_predict_npz(path_in, path_out, device):
    """Parity mode for the eval half: logits [1,A,C], reg [1,A,4], h, w from an .npz -> the unmodified ResNet.predict's
    [scores, labels, boxes].  device='cuda': exactly as written, on cuda:0; 'cpu': GPUs hidden, the reference's CPU branches."""
    import torch
    d = np.load(path_in, allow_pickle=False)
    if device == 'cpu':
        install_cpu_shim()
    dev = torch.device('cuda:0' if device == 'cuda' else 'cpu')
    m = stub_model()
    h, w = int(d['h']), int(d['w'])
    img = torch.zeros(1, 3, h, w, device=dev)
    m._cls, m._reg = torch.from_numpy(d['logits']).to(dev), torch.from_numpy(d['reg']).to(dev)
    m._anchors = sys.modules['retinanet.anchors'].Anchors()(img)
    with torch.no_grad():
        scores, labels, boxes = m.predict(img)
    np.savez(path_out, scores=scores.cpu().numpy(), labels=labels.cpu().numpy(), boxes=boxes.cpu().numpy())


def time_predict(mu, images, device, height=800, width=1333, classes=80, seed=7):
    """Wall time per predict() call of the unmodified reference (forward stubbed) on COCO-shaped head outputs,
    logits ~ N(mu, 2): what evaluator.py:324-329 pays per image after the network, detections read back with .cpu()."""
    import torch
    if device == 'cpu':
        install_cpu_shim()
    dev = torch.device('cuda:0' if device == 'cuda' else 'cpu')
    m = stub_model()
    img = torch.zeros(1, 3, height, width, device=dev)
    m._anchors = sys.modules['retinanet.anchors'].Anchors()(img)
    a = m._anchors.shape[1]
    gen = torch.Generator(device=dev).manual_seed(seed)
    logits = torch.randn(images + 1, a, classes, device=dev, generator=gen) * 2.0 + mu
    reg = torch.randn(images + 1, a, 4, device=dev, generator=gen) * 0.3
    kept = 0
    t0 = 0.0
    with torch.no_grad():
        for j in range(images + 1):          # the first call is the warm-up
            if j == 1:
                if dev.type == 'cuda':
                    torch.cuda.synchronize()
                t0 = time.perf_counter()
            m._cls, m._reg = logits[j:j + 1], reg[j:j + 1]
            s, l, b = m.predict(img)
            s, l, b = s.cpu(), l.cpu(), b.cpu()
            if j >= 1:
                kept += s.shape[0]
    dt = (time.perf_counter() - t0) / images
    return {'ms_per_image': dt * 1e3, 'value': 1.0 / dt, 'unit': 'images/s', 'images': images, 'kept_per_image': kept / images,
            'device': device, 'threads': torch.get_num_threads(), 'torch': torch.__version__}


def load_reference():
    """(FocalLoss class, calc_iou, Anchors class) of the snapshot behind the CPU shim."""
    install_cpu_shim()
    losses, anchors = import_snapshot()
    return losses.FocalLoss, losses.calc_iou, anchors.Anchors


class Params:
    """Duck type of the reference's preprocessing/params.py Params with main.py:116-177's CLI defaults."""

    def __init__(self, num_past_class=(0,), **kw):
        self.d = dict(alpha=0.25, gamma=2.0, distill=False, enhance_on_new=False, ignore_past_class=False,
                      new_ignore_past_class=False, decrease_positive_by_IOU=False, decrease_positive=1.0, persuado_label=False)
        self.d.update(kw)
        self.states = [{'num_past_class': n} for n in num_past_class]

    def __getitem__(self, k):
        return self.d.get(k, None)


def focal_step(focal, cls, reg, anchors, ann, cur_state=0, params=None):
    """One fwd + bwd of the reference module; returns (bg[N], fg[N], reg_loss[1], dL/dcls, dL/dreg) as numpy."""
    import torch
    p = torch.from_numpy(cls).requires_grad_(True)
    r = torch.from_numpy(reg).requires_grad_(True)
    out = focal(p, r, torch.from_numpy(anchors), torch.from_numpy(ann), cur_state, params or Params())
    bg, fg = out['cls_loss']
    (bg.mean() + fg.mean() + out['reg_loss'].mean()).backward()
    # no positive anchor anywhere: the reference's regression terms are constants and autograd leaves r.grad unset
    gr = r.grad.numpy() if r.grad is not None else np.zeros_like(reg)
    return bg.detach().numpy(), fg.detach().numpy(), out['reg_loss'].detach().numpy(), p.grad.numpy(), gr


def time_reference(images, frac, steps, warmup, threads, height=800, width=1333, classes=80, gmax=20, seed=1234):
    """`images` COCO-shaped images per step, the first frac*A anchors of each (every anchor is independent work, so the
    rate in images/s is images*frac*steps / time); same generator as bench.make_cpu_images."""
    import torch
    sys.path.insert(0, ROOT)
    from bench import synth_annotations
    from oracle import head_oracle as O
    FocalLoss, _, _ = load_reference()
    torch.set_num_threads(max(1, threads))
    rng = np.random.default_rng(seed)
    a_full = O.num_anchors(height, width)
    a = max(1, int(a_full * frac))
    anchors = O.anchors_for_image(height, width)[:, :a].copy()
    logits = rng.normal(-4.0, 2.0, (images, a, classes)).astype(np.float32)
    cls = (1.0 / (1.0 + np.exp(-logits))).astype(np.float32)
    reg = rng.normal(0, 1, (images, a, 4)).astype(np.float32)
    ann = synth_annotations(rng, images, gmax, height, width, classes, empty=())
    focal = FocalLoss()
    for _ in range(warmup):
        focal_step(focal, cls, reg, anchors, ann)
    t0 = time.perf_counter()
    for _ in range(steps):
        focal_step(focal, cls, reg, anchors, ann)
    dt = time.perf_counter() - t0
    f = a / a_full
    return {'value': images * f * steps / dt, 'unit': 'images/s', 's_per_step': dt / steps, 'images': images, 'anchor_fraction': f,
            'steps': steps, 'threads': torch.get_num_threads(), 'torch': torch.__version__}


def run_npz(path_in, path_out, device):
    """Parity mode: inputs from an .npz (cls, reg, anchors, ann, cur_state, num_past_class, params_keys, params_vals), outputs of
    the unmodified reference to another .npz.  device='cuda' runs the reference exactly as written (it hard-codes cuda:0 --
    the oracle of record on a GPU box); device='cpu' goes through the shim."""
    import torch
    d = np.load(path_in, allow_pickle=False)
    if device == 'cpu':
        FocalLoss, _, _ = load_reference()
    else:
        FocalLoss = import_snapshot()[0].FocalLoss
    kw = {}
    for k, v in zip(d['params_keys'], d['params_vals']):
        k = str(k)
        kw[k] = float(v) if k in ('alpha', 'gamma', 'decrease_positive') else bool(v)
    params = Params([int(x) for x in d['num_past_class']], **kw)
    dev = torch.device('cuda:0' if device == 'cuda' else 'cpu')
    p = torch.from_numpy(d['cls']).to(dev).requires_grad_(True)
    r = torch.from_numpy(d['reg']).to(dev).requires_grad_(True)
    out = FocalLoss()(p, r, torch.from_numpy(d['anchors']).to(dev), torch.from_numpy(d['ann']).to(dev), int(d['cur_state']), params)
    bg, fg = out['cls_loss']
    w_bg = torch.from_numpy(d['w_bg']).to(dev)
    w_fg = torch.from_numpy(d['w_fg']).to(dev)
    loss = (bg * w_bg).sum() + (fg * w_fg).sum() + out['reg_loss'].sum() * float(d['w_reg'])
    if 'enhance_on_new_loss' in out:
        loss = loss + out['enhance_on_new_loss']
    loss.backward()
    res = {'bg': bg.detach().cpu().numpy(), 'fg': fg.detach().cpu().numpy(), 'reg_loss': out['reg_loss'].detach().cpu().numpy(),
           'grad_cls': p.grad.cpu().numpy(), 'grad_reg': (r.grad.cpu().numpy() if r.grad is not None else np.zeros_like(d['reg']))}
    if 'enhance_on_new_loss' in out:
        res['enhance'] = np.asarray(float(out['enhance_on_new_loss']), np.float32)
    if 'bg_masks' in out:
        res['bg_masks'] = out['bg_masks'].cpu().numpy()
    np.savez(path_out, **res)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--npz', nargs=2, metavar=('IN', 'OUT'), help='parity mode: run the reference on the inputs of IN, write OUT')
    ap.add_argument('--device', default='cpu', choices=['cpu', 'cuda'])
    ap.add_argument('--predict-npz', nargs=2, metavar=('IN', 'OUT'), help='parity mode of the eval half (ResNet.predict)')
    ap.add_argument('--predict-time', type=float, nargs='+', metavar='MU',
                    help='time predict() on logits ~ N(MU, 2), --images calls per MU; prints one JSON list')
    ap.add_argument('--images', type=int, default=1)
    ap.add_argument('--frac', type=float, default=1.0)
    ap.add_argument('--steps', type=int, default=1)
    ap.add_argument('--warmup', type=int, default=0)
    ap.add_argument('--threads', type=int, default=os.cpu_count() or 1)
    a = ap.parse_args()
    if a.device == 'cpu':
        # the CPU runs must take the reference's OWN CPU branches (`if torch.cuda.is_available(): ... torch.cuda.FloatTensor(...)`,
        # losses.py:116, 146, 193, 424, 439): hide the GPUs of a GPU box from this process before torch is imported
        os.environ['CUDA_VISIBLE_DEVICES'] = ''
    if a.npz:
        run_npz(a.npz[0], a.npz[1], a.device)
        return
    if a.predict_npz:
        run_predict_npz(a.predict_npz[0], a.predict_npz[1], a.device)
        return
    if a.predict_time is not None:
        import torch
        torch.set_num_threads(max(1, a.threads))
        print(json.dumps([time_predict(mu, a.images, a.device) for mu in a.predict_time]))
        return
    print(json.dumps(time_reference(a.images, a.frac, a.steps, a.warmup, a.threads)))


if __name__ == '__main__':
    main()
