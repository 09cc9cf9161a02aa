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

from .losses import iou_assign


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


class OutputNorm(nn.Module):
    """Drop-in for IL_method/mas.py Output_norm (:35-67): same forward signature and result dict, differentiable w.r.t.
    classifications and regressions."""

    def forward(self, classifications, regressions, anchors, annotations):
        n = classifications.shape[0]
        positive = match_anchors(anchors, annotations, 0.5)['positive']
        reg_term = regressions.new_zeros(())
        counts = positive.sum(dim=1)
        for j in range(n):                                   # per-image mean over that image's positive rows (mas.py:52-55)
            if int(counts[j]) > 0:
                reg_term = reg_term + regressions[j][positive[j]].abs().mean()
        result = {'regression': reg_term / n,
                  'classification': torch.sum(torch.pow(classifications, 2)) / (n * classifications.shape[2])}
        return result
