"""the other callers of `calc_iou` + `torch.max` in the reference, served by the K2 assign kernel
instead of a materialised [A,G] IoU matrix per image:

  IL_method/mas.py:35-67        Output_norm.forward        (positive mask -> mean |regression|, sum cls^2)
  IL_method/prototype.py:24-47  ProtoTyper._get_positive   (positive mask at a custom threshold + assigned class)
  IL_method/weight_init.py:75-115 Weight_similarity.forward (positive mask + assigned class; uses match_anchors)

The matching (IoU, max, first-index argmax, label gather) runs in one launch for the whole batch; what remains is a few
autograd-visible torch reductions over the selected rows, which keeps these modules differentiable like the originals.
"""
import torch
import torch.nn as nn

from . import _lib
from .losses import _DeviceGuard, _check_cuda_f32, _stream, iou_assign


def match_anchors(anchors, annotations, threshold=0.5, num_classes=8191):
    """Per image: IoU_max, IoU_argmax = max(calc_iou(anchors, valid GT), 1); positive = IoU_max >= threshold;
    target = label of the assigned GT.  Returns dict(positive bool [N,A], targets int64 [N,A], iou_max [N,A],
    argmax int32 [N,A] (compacted GT index), nvalid int32 [N]).  Images without GT: nothing positive (the reference's
    callers crash on them)."""
    asg = iou_assign(anchors, annotations, num_classes, want_argmax=True, want_iou_max=True)
    positive = (asg['iou_max'] >= threshold) & (asg['nvalid'] > 0).unsqueeze(1)
    return dict(positive=positive, targets=asg['label'].to(torch.int64), iou_max=asg['iou_max'], argmax=asg['argmax'],
                nvalid=asg['nvalid'])


def get_positive(anchors, annotations, threshold, num_anchors):
    """ProtoTyper._get_positive (prototype.py:24-47): (positive_indices bool [N, A/num_anchors, num_anchors],
    targets int64 of the same shape)."""
    m = match_anchors(anchors, annotations, threshold)
    n = annotations.shape[0]
    return m['positive'].view(n, -1, num_anchors), m['targets'].view(n, -1, num_anchors)


class _MaskedAbsMeanFn(torch.autograd.Function):
    """sum_j mean(|reg[j][positive_j]|) over images with at least one positive anchor (mas.py:52-55), no host sync."""

    @staticmethod
    def forward(ctx, reg, positive):
        lib = _lib.load()
        n, a, _ = reg.shape
        dev = reg.device
        with _DeviceGuard(dev):
            terms = torch.empty(n, dtype=torch.float32, device=dev)
            counts = torch.empty(n, dtype=torch.float32, device=dev)
            _lib.check(lib.cldet_masked_abs_mean_forward(reg.data_ptr(), positive.data_ptr(), n, a, terms.data_ptr(),
                                                         counts.data_ptr(), _stream()))
        ctx.save_for_backward(reg, positive, counts)
        return terms

    @staticmethod
    def backward(ctx, g):
        reg, positive, counts = ctx.saved_tensors
        n, a, _ = reg.shape
        g = g.to(torch.float32)
        stride = g.stride(0) if g.dim() else 0
        with _DeviceGuard(reg.device):
            grad = torch.empty_like(reg)
            _lib.check(_lib.load().cldet_masked_abs_mean_backward(reg.data_ptr(), positive.data_ptr(), n, a, counts.data_ptr(),
                                                                  g.data_ptr(), stride, grad.data_ptr(), _stream()))
        return grad, None


class OutputNorm(nn.Module):
    """Drop-in for IL_method/mas.py Output_norm (:35-67): same forward signature and result dict, differentiable w.r.t.
    classifications and regressions.  No per-image Python loop and no host synchronisation: the matching is one K2 launch for
    the batch, the per-image masked mean one more (cldet_masked_abs_mean_*)."""

    def forward(self, classifications, regressions, anchors, annotations):
        n = classifications.shape[0]
        positive = match_anchors(anchors, annotations, 0.5)['positive'].to(torch.uint8).contiguous()
        reg = _check_cuda_f32('regressions', regressions)
        terms = _MaskedAbsMeanFn.apply(reg, positive)                    # [N]: per-image mean, 0 for images without positives
        result = {'regression': terms.sum() / n,
                  'classification': torch.sum(torch.pow(classifications, 2)) / (n * classifications.shape[2])}
        return result


class WeightSimilarity(object):
    """Drop-in for IL_method/weight_init.py Weight_similarity (:75-115): forward(img_batch, annotations) ->
    (classification[K,C] normalised rows, assigned labels[K] float) for image 0 of the batch, or None when it has no GT.
    calc_iou + max + the label gather are the K2 kernel (match_anchors); the row selection has a data-dependent size, so --
    like the reference -- it ends in one boolean index."""

    def __init__(self, model, new_class_num, old_class_num, thresold=0.5):
        self.model = model
        self.new_class_num = new_class_num
        self.old_class_num = old_class_num
        self.thresold = thresold

    def forward(self, img_batch, annotations):
        classifications, _, anchors = self.model(img_batch, return_feat=False, return_anchor=True, enable_act=True)
        m = match_anchors(anchors, annotations[:1], 0.5)
        classification = torch.clamp(classifications[0], 1e-4, 1.0 - 1e-4)
        if int(m['nvalid'][0]) == 0:
            return None
        rowsum = torch.sum(classification, dim=1)
        indices = torch.logical_and(m['positive'][0], torch.ge(rowsum, self.thresold))
        classification = classification[indices, :]
        classification = classification / torch.sum(classification, dim=1).unsqueeze(dim=1)
        return classification, m['targets'][0][indices].to(torch.float32)
