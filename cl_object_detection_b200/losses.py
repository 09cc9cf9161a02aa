"""Drop-in for the loss half of the detection head: retinanet/losses.py calc_iou (:4-21) and
FocalLoss.forward (:252-452), backed by the fused sm_100a kernels in libcldet.so.

Differences in HOW (not WHAT): the per-image Python loop, the [A,G] IoU matrix, the dense [A,C] target
matrix and ~200 eager kernels per image are replaced by two launches for the whole batch; the gradient
w.r.t. the probabilities and the regression outputs is produced IN THE FORWARD PASS with the upstream
weights the caller is expected to apply (IL_Loss takes .mean() of each term, losses.py:584-588), and
autograd's backward only verifies those weights on the device and patches the (rare) images where they
differ (clip_loss masking, losses.py:575-581).
"""
import threading

import torch
import torch.nn as nn

from . import _lib, _ops
from .anchors import grid_of
from .params import loss_param_args, to_loss_params


def _check_cuda_f32(name, t):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError('%s must be a CUDA tensor: this path has no CPU implementation' % name)
    if t.dtype != torch.float32:
        raise TypeError('%s must be float32 (got %s)' % (name, t.dtype))
    return t if t.is_contiguous() else t.contiguous()


_FOCAL_OP = None


def _focal_loss_op():
    """torch.ops.cldet.focal_loss.default (the overload itself: skips the packet's overload resolution on every call)."""
    global _FOCAL_OP
    if _FOCAL_OP is None:
        _FOCAL_OP = _ops.load().focal_loss.default
    return _FOCAL_OP


def _stream():
    return torch.cuda.current_stream().cuda_stream


def calc_iou(a, b):
    """Pairwise IoU [A,G] in fp32 with the reference's exact op order (losses.py:4-21)."""
    a = _check_cuda_f32('a', a)
    b = _check_cuda_f32('b', b)
    if a.dim() != 2 or a.shape[1] != 4 or b.dim() != 2 or b.shape[1] != 4:
        raise ValueError('calc_iou expects [A,4] and [G,4] boxes')
    with _DeviceGuard(a.device):
        out = torch.empty((a.shape[0], b.shape[0]), dtype=torch.float32, device=a.device)
        _lib.check(_lib.load().cldet_calc_iou(a.data_ptr(), a.shape[0], b.data_ptr(), b.shape[0], out.data_ptr(), _stream()))
    return out


def iou_assign(anchors, annotations, num_classes, want_argmax=True, want_iou_max=True):
    """Anchor-to-GT assignment for a batch (losses.py:287-288, 309-341) without touching the class map.

    anchors [1,A,4] or [A,4]; annotations [N,G,5] (pad rows label == -1).
    Returns dict(meta uint32-as-int32 [N,A], state uint8 [N,A] (0 bg, 1 pos, 2 ignore, 3 empty image),
    argmax int32 [N,A] (compacted GT index, -1 for empty images), iou_max [N,A], npos int32 [N], nvalid int32 [N]).
    """
    anchors = _check_cuda_f32('anchors', anchors).reshape(-1, 4)
    annotations = _check_cuda_f32('annotations', annotations)
    if annotations.dim() != 3 or annotations.shape[2] != 5:
        raise ValueError('annotations must be [N,G,5]')
    n, g = annotations.shape[0], annotations.shape[1]
    a = anchors.shape[0]
    dev = anchors.device
    with _DeviceGuard(dev):
        meta = torch.empty((n, a), dtype=torch.int32, device=dev)
        argmax = torch.empty((n, a), dtype=torch.int32, device=dev) if want_argmax else None
        iou_max = torch.empty((n, a), dtype=torch.float32, device=dev) if want_iou_max else None
        npos = torch.zeros(n, dtype=torch.int32, device=dev)
        nvalid = torch.empty(n, dtype=torch.int32, device=dev)
        _lib.check(_lib.load().cldet_iou_assign(anchors.data_ptr(), a, annotations.data_ptr(), n, g, int(num_classes),
                                                meta.data_ptr(), _lib.ptr(argmax), _lib.ptr(iou_max), npos.data_ptr(),
                                                nvalid.data_ptr(), _stream()))
    return dict(meta=meta, state=(meta & 3).to(torch.uint8), label=(meta >> 3) & 0x1fff, argmax=argmax,
                iou_max=iou_max, npos=npos, nvalid=nvalid)


_tls = threading.local()


def _workspace(dev, stream, n, a):
    """Zero-initialised scratch for (device, stream, N, A), cached per host thread.  libcldet leaves the header of a
    workspace zeroed after every call, so it is cleared exactly once; calls that share it are ordered by the stream."""
    cache = getattr(_tls, 'ws', None)
    if cache is None:
        cache = _tls.ws = {}
    key = (dev.index, stream, n, a)
    ws = cache.get(key)
    if ws is None:
        if len(cache) > 8:
            cache.clear()
        nbytes = _lib.load().cldet_focal_loss_workspace_bytes(n, a)
        ws = cache[key] = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
    return ws


def _drop_workspaces():
    _tls.ws = {}


class _DeviceGuard:
    """`with torch.cuda.device(dev)` only when dev is not already current (the context manager costs ~10 us of host time)."""

    def __init__(self, dev):
        self.ctx = None if torch.cuda.current_device() == dev.index else torch.cuda.device(dev)

    def __enter__(self):
        if self.ctx is not None:
            self.ctx.__enter__()

    def __exit__(self, *exc):
        if self.ctx is not None:
            self.ctx.__exit__(*exc)


def _row(t):
    """(device pointer, element stride) of an upstream-gradient row; None -> (NULL, 0) = zeros."""
    if t is None:
        return None, 0
    if t.dtype != torch.float32:
        t = t.to(torch.float32)
    return t, (t.stride(0) if t.dim() else 0)


class _FocalLossHeadFn(torch.autograd.Function):
    """The fused loss on the head's raw conv outputs (cldet_focal_loss_head).  Tensor inputs: 5 classification levels
    [N, 9*C, H_l, W_l] followed by 5 regression levels [N, 36, H_l, W_l]; outputs as _FocalLossFn."""

    @staticmethod
    def forward(ctx, anchors, annotations, lp, hint, want_bg_mask, check_labels, hw, *levels):
        lib = _lib.load()
        nl = len(levels) // 2
        cls_lv, reg_lv = levels[:nl], levels[nl:]
        n = cls_lv[0].shape[0]
        c = cls_lv[0].shape[1] // 9
        a = anchors.shape[1]
        g = annotations.shape[1]
        dev = cls_lv[0].device
        need_grad = any(ctx.needs_input_grad[7:])
        with _DeviceGuard(dev):
            stream = _stream()
            losses = torch.empty((4, n), dtype=torch.float32, device=dev)
            meta = torch.empty((n, a), dtype=torch.int32, device=dev)
            iou_max = torch.empty((n, a), dtype=torch.float32, device=dev) if lp.decrease_positive_by_iou else None
            counts = torch.empty((2, n), dtype=torch.int32, device=dev)
            npos, nvalid = counts[0], counts[1]
            bg_mask = torch.empty((n, a), dtype=torch.uint8, device=dev) if want_bg_mask else None
            status = torch.empty(1, dtype=torch.int32, device=dev) if check_labels else None
            ws = _workspace(dev, stream, n, a)
            if need_grad:
                baked = torch.empty((4, n), dtype=torch.float32, device=dev)
                gcls = [torch.empty_like(t) for t in cls_lv]
                greg = [torch.empty_like(t) for t in reg_lv]
            else:
                hint = baked = gcls = greg = None
            try:
                _lib.check(lib.cldet_focal_loss_head(
                    _lib.ptr_array(cls_lv), _lib.ptr_array(reg_lv), nl, hw[0], hw[1], anchors.data_ptr(), annotations.data_ptr(),
                    n, c, g, lp, _lib.ptr(hint), _lib.ptr(baked), _lib.ptr_array(gcls), _lib.ptr_array(greg), losses.data_ptr(),
                    meta.data_ptr(), _lib.ptr(iou_max), npos.data_ptr(), nvalid.data_ptr(), _lib.ptr(bg_mask), _lib.ptr(status),
                    ws.data_ptr(), ws.numel(), stream))
            except Exception:
                _drop_workspaces()
                raise
        if check_labels and int(status.item()) != 0:
            raise IndexError('a GT label is outside [0, %d): the reference indexes the class dimension with it '
                             '(losses.py:341)' % c)
        ctx.lp, ctx.hw, ctx.nl, ctx.shape = lp, hw, nl, (n, a, c, g)
        ctx.backward_calls = 0
        if need_grad:
            ctx.save_for_backward(anchors, annotations, baked, meta, npos, *cls_lv, *reg_lv, *gcls, *greg)
            ctx.iou_max = iou_max
            ctx.ws = ws
        ctx.mark_non_differentiable(npos, nvalid)
        outs = (losses[0], losses[1], losses[2], losses[3], npos, nvalid)
        if want_bg_mask:
            ctx.mark_non_differentiable(bg_mask)
            outs = outs + (bg_mask,)
        return outs

    @staticmethod
    def backward(ctx, g_bg, g_fg, g_reg, g_enh, *unused):
        saved = ctx.saved_tensors
        anchors, annotations, baked, meta, npos = saved[:5]
        nl = ctx.nl
        cls_lv, reg_lv = saved[5:5 + nl], saved[5 + nl:5 + 2 * nl]
        gcls, greg = saved[5 + 2 * nl:5 + 3 * nl], saved[5 + 3 * nl:5 + 4 * nl]
        n, a, c, g = ctx.shape
        rows = [_row(t) for t in (g_bg, g_fg, g_reg, g_enh)]
        with _DeviceGuard(anchors.device):
            _lib.check(_lib.load().cldet_focal_loss_head_reweight(
                _lib.ptr_array(cls_lv), _lib.ptr_array(reg_lv), nl, ctx.hw[0], ctx.hw[1], anchors.data_ptr(),
                annotations.data_ptr(), n, c, g, ctx.lp, _lib.ptr(rows[0][0]), rows[0][1], _lib.ptr(rows[1][0]), rows[1][1],
                _lib.ptr(rows[2][0]), rows[2][1], _lib.ptr(rows[3][0]), rows[3][1], baked.data_ptr(), _lib.ptr_array(gcls),
                _lib.ptr_array(greg), meta.data_ptr(), _lib.ptr(ctx.iou_max), npos.data_ptr(), ctx.ws.data_ptr(),
                ctx.ws.numel(), _stream()))
        ctx.backward_calls += 1
        grads = list(gcls) + list(greg)
        if ctx.backward_calls > 1:
            grads = [t.clone() for t in grads]
        return (None,) * 7 + tuple(grads)


class FocalLoss(nn.Module):
    """Same call signature and result dict as the reference module (losses.py:252-253, 444-452).

    forward(classifications[N,A,C] probs, regressions[N,A,4], anchors[1,A,4], annotations[N,G,5], cur_state, params,
            progress=-1) -> {'cls_loss': (bg[N], fg[N]), 'reg_loss': [1], ['bg_masks': bool[M,A]],
                             ['enhance_on_new_loss': scalar]}

    `progress` is accepted and ignored: in the reference it only feeds a statement that has no effect
    (losses.py:388-392 multiplies a temporary).  Inputs are never modified.

    upstream_hint: the dL/d(term) the caller will apply -- 'mean' (1/N for bg, fg and the regression mean; what IL_Loss
    does) or a [4,N] tensor (rows bg, fg, per-image reg, enhance).  A wrong hint costs a re-weighting pass in backward,
    never a wrong gradient.
    check_labels=True adds a host sync to raise IndexError on out-of-range GT labels like the reference does.
    from_logits=True (beyond the reference's signature): `classifications` holds the classification
    head's raw LOGITS (after any BiC correction); the kernel applies ATen's sigmoid itself and returns dL/dlogits, which
    replaces `self.classifier_act(classification)` (losses.py:566, 633-647) and its SigmoidBackward: 8 B/element of HBM
    traffic instead of 28.
    """

    def __init__(self, upstream_hint='mean', check_labels=False, from_logits=False):
        super().__init__()
        self.upstream_hint = upstream_hint
        self.check_labels = check_labels
        self.from_logits = bool(from_logits)
        self._hint_cache = {}

    def _hint(self, n, device):
        if isinstance(self.upstream_hint, torch.Tensor):
            if tuple(self.upstream_hint.shape) != (4, n):
                raise ValueError('upstream_hint must be [4, N]')
            return self.upstream_hint.to(device=device, dtype=torch.float32).contiguous()
        if self.upstream_hint != 'mean':
            raise ValueError("upstream_hint must be 'mean' or a [4,N] tensor")
        key = (n, device.index)
        w = self._hint_cache.get(key)
        if w is None:
            w = torch.full((4, n), 1.0 / n, dtype=torch.float32)
            w[3] = 1.0
            w = w.to(device)
            self._hint_cache[key] = w
        return w

    def forward(self, classifications, regressions, anchors, annotations, cur_state: int, params, progress=-1, peer=None):
        cls = _check_cuda_f32('classifications', classifications)
        reg = _check_cuda_f32('regressions', regressions)
        anc = _check_cuda_f32('anchors', anchors)
        ann = _check_cuda_f32('annotations', annotations)
        if cls.dim() != 3 or reg.dim() != 3 or reg.shape[2] != 4 or reg.shape[:2] != cls.shape[:2]:
            raise ValueError('classifications must be [N,A,C] and regressions [N,A,4]')
        if anc.dim() != 3 or anc.shape[0] != 1 or anc.shape[1] != cls.shape[1] or anc.shape[2] != 4:
            raise ValueError('anchors must be [1,A,4]')
        if ann.dim() != 3 or ann.shape[0] != cls.shape[0] or ann.shape[2] != 5:
            raise ValueError('annotations must be [N,G,5]')
        if ann.shape[1] == 0:
            raise ValueError('annotations needs at least one (possibly padding) row; the collater emits [N,1,5] of -1')
        n, _, c = cls.shape
        lp = loss_param_args(params, int(cur_state), c)
        grid = grid_of(anc)            # anchors made by our Anchors module: GT-centric assignment
        h, w = grid if grid is not None else (0, 0)
        incremental = cur_state > 0
        want_mask = bool(incremental and params['distill'])
        # one call into the C++ op layer (csrc/cldet_torch.cpp): allocation, the two kernel launches and the autograd node
        outs = _focal_loss_op()(cls, reg, anc, ann, self._hint(n, cls.device), *lp, h, w, self.from_logits, want_mask,
                                self.check_labels, [] if peer is None else peer.descriptor())
        bg, fg, reg_j, enh_j, reg_loss, npos, nvalid, meta = outs[:8]
        if self.check_labels and int(outs[-1].item()) != 0:
            raise IndexError('a GT label is outside [0, %d): the reference indexes the class dimension with it '
                             '(losses.py:341)' % c)
        result = {'cls_loss': (bg, fg), 'reg_loss': reg_loss}          # losses.py:444-445
        if incremental:
            if want_mask:
                # the reference appends a mask only for images that have GT (quirk Q6): M <= N rows
                result['bg_masks'] = outs[8].bool()[nvalid > 0]
            if params['enhance_on_new']:
                result['enhance_on_new_loss'] = enh_j.sum()
        # plain attributes, set past nn.Module.__setattr__ (its parameter / buffer / submodule bookkeeping costs ~2 us per tensor)
        d = self.__dict__
        d['last_npos'], d['last_nvalid'], d['last_reg_per_image'], d['last_meta'] = npos, nvalid, reg_j, meta
        return result

    def forward_head(self, cls_levels, reg_levels, anchors, annotations, cur_state: int, params, image_size, progress=-1):
        """SURVEY 8(f) row f1, second half: FocalLoss on the head's RAW conv outputs (beyond the reference's signature).

        cls_levels / reg_levels: the five per-level results of the classification / regression output convolutions as they
        come out of the conv, [N, 9*C, H_l, W_l] and [N, 36, H_l, W_l] (contiguous NCHW) -- i.e. `self.output(out)` (after
        `self.output_act` unless from_logits=True) of ClassificationModel / RegressionModel BEFORE their
        permute + contiguous + view (retinanet/model.py:125-130, 170-184) and before ResNet.forward's torch.cat
        (model.py:472-474).  image_size = (H, W) of the input batch; anchors = Anchors()(img) for it.  Returns the same
        dict as forward(); gradients flow to the level tensors in their own layout."""
        cls_lv = [_check_cuda_f32('cls_levels[%d]' % i, t) for i, t in enumerate(cls_levels)]
        reg_lv = [_check_cuda_f32('reg_levels[%d]' % i, t) for i, t in enumerate(reg_levels)]
        anc = _check_cuda_f32('anchors', anchors)
        ann = _check_cuda_f32('annotations', annotations)
        h, w = int(image_size[0]), int(image_size[1])
        if len(cls_lv) != 5 or len(reg_lv) != 5:
            raise ValueError('expected the 5 pyramid levels (3..7) of both heads')
        n = cls_lv[0].shape[0]
        if cls_lv[0].dim() != 4 or cls_lv[0].shape[1] % 9 != 0:
            raise ValueError('classification levels must be [N, 9*C, H_l, W_l]')
        c = cls_lv[0].shape[1] // 9
        total = 0
        for l in range(5):
            hl, wl = (h + 2 ** (l + 3) - 1) // 2 ** (l + 3), (w + 2 ** (l + 3) - 1) // 2 ** (l + 3)
            if tuple(cls_lv[l].shape) != (n, 9 * c, hl, wl) or tuple(reg_lv[l].shape) != (n, 36, hl, wl):
                raise ValueError('level %d: expected cls [%d,%d,%d,%d] and reg [%d,36,%d,%d] for a %dx%d input'
                                 % (l + 3, n, 9 * c, hl, wl, n, hl, wl, h, w))
            total += 9 * hl * wl
        if anc.dim() != 3 or anc.shape[0] != 1 or anc.shape[1] != total or anc.shape[2] != 4:
            raise ValueError('anchors must be [1,%d,4] for a %dx%d input' % (total, h, w))
        if grid_of(anc) != (h, w):
            raise ValueError('forward_head needs the anchors made by cl_object_detection_b200.Anchors for this image size')
        if ann.dim() != 3 or ann.shape[0] != n or ann.shape[2] != 5 or ann.shape[1] == 0:
            raise ValueError('annotations must be [N,G>=1,5]')
        lp = to_loss_params(params, int(cur_state), c)
        lp.cls_is_logits = int(self.from_logits)
        lp.image_height, lp.image_width = h, w
        incremental = cur_state > 0
        want_mask = bool(incremental and params['distill'])
        outs = _FocalLossHeadFn.apply(anc, ann, lp, self._hint(n, anc.device), want_mask, self.check_labels, (h, w),
                                      *cls_lv, *reg_lv)
        bg, fg, reg_j, enh_j, npos, nvalid = outs[:6]
        result = {'cls_loss': (bg, fg), 'reg_loss': reg_j.mean(dim=0, keepdim=True)}
        if incremental:
            if params['distill']:
                result['bg_masks'] = outs[6].bool()[nvalid > 0]
            if params['enhance_on_new']:
                result['enhance_on_new_loss'] = enh_j.sum()
        d = self.__dict__
        d['last_npos'], d['last_nvalid'], d['last_reg_per_image'] = npos, nvalid, reg_j
        return result
