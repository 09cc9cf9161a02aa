"""the head-distillation terms of IL_Loss (retinanet/losses.py:705-737) fused into two kernels.

    out = head_distillation(classification, regression, prev_classification, prev_regression, bg_masks,
                            distill_logits=params['distill_logits'], ignore_GD=params['ignore_GD'])
    result['dist_cls_loss'], result['dist_reg_loss'] = out['dist_cls_loss'], out['dist_reg_loss']

classification [N,A,C] are the current model's LOGITS (all C columns; only the first P = prev_classification.shape[2]
take part, losses.py:705), prev_classification [N,A,P] the previous model's logits, bg_masks the bool [N,A] mask FocalLoss
returns with distill=True.  Differentiable w.r.t. classification and regression, no host synchronisation.
"""
import torch

from . import _lib
from .losses import _DeviceGuard, _check_cuda_f32, _stream


class _HeadDistillFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, cls, reg, prev_cls, prev_reg, bg_mask, use_logits, ignore_gd):
        lib = _lib.load()
        n, a, c = cls.shape
        p = prev_cls.shape[2]
        dev = cls.device
        with _DeviceGuard(dev):
            losses = torch.empty(2, dtype=torch.float32, device=dev)
            counts = torch.empty(2, dtype=torch.float32, device=dev)
            ws = torch.empty(lib.cldet_distill_workspace_bytes(n, a), dtype=torch.uint8, device=dev)
            _lib.check(lib.cldet_distill_forward(cls.data_ptr(), prev_cls.data_ptr(), reg.data_ptr(), prev_reg.data_ptr(),
                                                 bg_mask.data_ptr(), n, a, c, p, int(use_logits), int(ignore_gd),
                                                 losses.data_ptr(), counts.data_ptr(), ws.data_ptr(), ws.numel(), _stream()))
        ctx.save_for_backward(cls, reg, prev_cls, prev_reg, bg_mask, counts)
        ctx.flags = (int(use_logits), int(ignore_gd))
        return losses[0], losses[1]

    @staticmethod
    def backward(ctx, g_cls, g_reg):
        cls, reg, prev_cls, prev_reg, bg_mask, counts = ctx.saved_tensors
        n, a, c = cls.shape
        p = prev_cls.shape[2]
        dev = cls.device
        g_cls = None if g_cls is None else g_cls.to(torch.float32).contiguous()
        g_reg = None if g_reg is None else g_reg.to(torch.float32).contiguous()
        with _DeviceGuard(dev):
            dcls = torch.empty_like(cls)
            dreg = torch.empty_like(reg)
            _lib.check(_lib.load().cldet_distill_backward(
                cls.data_ptr(), prev_cls.data_ptr(), reg.data_ptr(), prev_reg.data_ptr(), bg_mask.data_ptr(), n, a, c, p,
                ctx.flags[0], ctx.flags[1], counts.data_ptr(), _lib.ptr(g_cls), _lib.ptr(g_reg), dcls.data_ptr(),
                dreg.data_ptr(), _stream()))
        return dcls, dreg, None, None, None, None, None


def head_distillation(classification, regression, prev_classification, prev_regression, bg_masks, distill_logits=False,
                      ignore_GD=False):
    cls = _check_cuda_f32('classification', classification)
    reg = _check_cuda_f32('regression', regression)
    pcls = _check_cuda_f32('prev_classification', prev_classification.detach())
    preg = _check_cuda_f32('prev_regression', prev_regression.detach())
    if cls.dim() != 3 or pcls.dim() != 3 or pcls.shape[:2] != cls.shape[:2] or pcls.shape[2] > cls.shape[2] or pcls.shape[2] < 1:
        raise ValueError('classification [N,A,C] and prev_classification [N,A,P] with 1 <= P <= C expected')
    if reg.shape != preg.shape or reg.shape[:2] != cls.shape[:2] or reg.shape[2] != 4:
        raise ValueError('regression and prev_regression must be [N,A,4]')
    if not bg_masks.is_cuda or tuple(bg_masks.shape) != tuple(cls.shape[:2]):
        raise ValueError('bg_masks must be a CUDA [N,A] mask (FocalLoss returns fewer rows when an image has no GT: the '
                         'reference fails on that batch too, losses.py:719)')
    mask = bg_masks.to(torch.uint8).contiguous() if bg_masks.dtype != torch.uint8 else bg_masks.contiguous()
    dc, dr = _HeadDistillFn.apply(cls, reg, pcls, preg, mask, bool(distill_logits), bool(ignore_GD))
    return {'dist_cls_loss': dc, 'dist_reg_loss': dr}


class _EnhanceErrorFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, cls, past, method):
        lib = _lib.load()
        n, a, c = cls.shape
        dev = cls.device
        with _DeviceGuard(dev):
            loss = torch.empty(1, dtype=torch.float32, device=dev)
            count = torch.empty(1, dtype=torch.float32, device=dev)
            ws_bytes = lib.cldet_enhance_error_workspace_bytes(cls.numel())
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
            _lib.check(lib.cldet_enhance_error_forward(cls.data_ptr(), n, a, c, past, method, loss.data_ptr(), count.data_ptr(),
                                                       ws.data_ptr(), ws_bytes, _stream()))
        ctx.save_for_backward(cls, count)
        ctx.args = (past, method)
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        cls, count = ctx.saved_tensors
        n, a, c = cls.shape
        g = g.to(torch.float32).contiguous()
        with _DeviceGuard(cls.device):
            grad = torch.empty_like(cls)
            _lib.check(_lib.load().cldet_enhance_error_backward(cls.data_ptr(), n, a, c, ctx.args[0], ctx.args[1], count.data_ptr(),
                                                                g.data_ptr(), grad.data_ptr(), _stream()))
        return grad, None, None


def enhance_error(classification, past_class_num, method='L2'):
    """SURVEY 8(f) row f2, second half -- `enhance_error` on replay batches (retinanet/losses.py:590-603):

        classification = classification[:, :, past_class_num:]; classification = classification[classification > 0.05]
        enhance_loss = {L1: abs, L2: pow 2, L3: pow 3}(classification).sum() / max(classification.shape[0], 1)

    classification [N,A,C] are class PROBABILITIES (that branch runs the model with enable_act=True).  One fused pass instead of
    a slice copy, a boolean gather (host sync) and three elementwise kernels; differentiable, no synchronisation."""
    cls = _check_cuda_f32('classification', classification)
    if cls.dim() != 3:
        raise ValueError('classification must be [N,A,C]')
    m = {'L1': 1, 'L2': 2, 'L3': 3}.get(str(method).upper())
    if m is None:            # the reference leaves enhance_loss unbound for any other string (UnboundLocalError)
        raise ValueError("enhance_error_method must be 'L1', 'L2' or 'L3'")
    past = int(past_class_num)
    if past < 0 or past > cls.shape[2]:
        raise IndexError('past_class_num outside [0, C]')
    return _EnhanceErrorFn.apply(cls, past, m)
