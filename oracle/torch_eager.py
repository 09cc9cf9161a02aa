"""Torch-eager restatement of the reference's detection-head path (TEST INFRASTRUCTURE -- not product code).

Why a second oracle: the reference does not run this path on the CPU in production, it runs it on the GPU as a long
sequence of eager ATen ops (~200 launches per image, a dense [A,C] target matrix, boolean-mask gathers
with host syncs).  /root/reference cannot travel to the GPU box, so this file restates that op-by-op evaluation with
plain torch calls so that

  * the `-m gpu` tests can compare the CUDA kernels with an autograd-derived result AT FULL SIZE on the same device
    (the numpy oracle needs minutes for 200 k anchors x 80 classes), and
  * bench.py's baseline leg can report what the reference's own execution model (eager ops + autograd +
    torchvision NMS) achieves on the very same B200, beside the CPU number.

Only tests/, __graft_entry__.smoke() and bench.py's baseline leg may import this module.  Nothing under
cl_object_detection_b200/ does.

Pinned (tests/test_oracle_golden.py::test_torch_eager_*) against the golden vectors generated from the unmodified
reference: state-0, default-flag incremental AND every IL-flag FocalLoss fixture (ignore_past_class, new_ignore_past_class,
enhance_on_new, decrease_positive, decrease_positive_by_IOU, all together: losses and autograd gradients), decode + clip, and
predict.

Reference statements restated here (paths relative to /root/reference):
  pairwise_iou          retinanet/losses.py:4-21
  focal_loss            retinanet/losses.py:252-452 (default flags), backward = torch autograd, like the reference
  decode_boxes          retinanet/utils.py:102-126
  clip_boxes            retinanet/utils.py:134-144
  predict               retinanet/model.py:507-550 (image 0, score > 0.05, torchvision.ops.batched_nms 0.5)
"""
import torch


def pairwise_iou(boxes_a, boxes_b):
    """[A,4] x [G,4] -> [A,G]; the broadcast temporaries are materialised one op at a time, as eager torch does."""
    area_b = (boxes_b[:, 2] - boxes_b[:, 0]) * (boxes_b[:, 3] - boxes_b[:, 1])
    ax1, ay1, ax2, ay2 = (boxes_a[:, k].unsqueeze(1) for k in range(4))
    iw = torch.min(ax2, boxes_b[:, 2]) - torch.max(ax1, boxes_b[:, 0])
    ih = torch.min(ay2, boxes_b[:, 3]) - torch.max(ay1, boxes_b[:, 1])
    iw = torch.clamp(iw, min=0)
    ih = torch.clamp(ih, min=0)
    union = ((boxes_a[:, 2] - boxes_a[:, 0]) * (boxes_a[:, 3] - boxes_a[:, 1])).unsqueeze(1) + area_b - iw * ih
    union = torch.clamp(union, min=1e-8)
    return (iw * ih) / union


class ILFlags:
    """The incremental-state switches FocalLoss.forward reads from `params` (losses.py:317-384); all off = state 0."""

    def __init__(self, past=0, ignore_past_class=False, new_ignore_past_class=False, enhance_on_new=False, decrease_positive=1.0,
                 decrease_positive_by_iou=False):
        self.past = int(past)
        self.ignore_past_class = bool(ignore_past_class)
        self.new_ignore_past_class = bool(new_ignore_past_class)
        self.enhance_on_new = bool(enhance_on_new)
        self.decrease_positive = float(decrease_positive)
        self.decrease_positive_by_iou = bool(decrease_positive_by_iou)


def _image_terms(prob, reg, anchor, geom, gt, alpha, gamma, il=None):
    """(bg, fg, reg, enhance) of one image with at least one GT row; every line is one or two eager kernels.
    il = ILFlags for an incremental state (losses.py:317-384), None for state 0."""
    dev = prob.device
    iou = pairwise_iou(anchor, gt[:, :4])
    best, which = torch.max(iou, dim=1)
    target = torch.full(prob.shape, -1.0, device=dev)
    background = best < 0.4
    positive = best >= 0.5
    if il is None or not il.ignore_past_class:
        target[background, :] = 0
    else:
        # old-class columns of background anchors stay "ignore" (:319-322) unless the anchor's old-class mass is small (:323-328)
        target[background, il.past:] = 0
        if il.new_ignore_past_class:
            old_mass = prob[:, :il.past].sum(dim=1)
            target[background & (old_mass < 0.5), :il.past] = 0
    npos = positive.sum()
    matched = gt[which, :]
    target[positive, :] = 0
    target[positive, matched[positive, 4].long()] = 1
    is_one = target == 1.0
    if il is None:
        base = torch.where(is_one, 1.0 - prob, prob)
    elif il.decrease_positive_by_iou:
        base = torch.where(is_one, 1.0 - prob, prob)
        mid = positive & (best <= 0.7)                                   # :354
        mid_one = torch.zeros(prob.shape, device=dev)
        mid_one[mid, matched[mid, 4].long()] = 1
        upper = torch.clip(best + 0.2, 1e-4, 1 - 1e-4).unsqueeze(1)      # :361
        capped = torch.where(prob >= upper, torch.full(prob.shape, 1e-4, device=dev), torch.abs(prob - upper))
        base = torch.where(mid_one == 1, capped, base)
    else:
        s = il.decrease_positive                                         # :365-366
        base = torch.where(is_one, s - torch.clip(prob, 0, s), prob)
    weight = torch.full(prob.shape, alpha, device=dev) * torch.pow(base, gamma)
    bce = -(target * torch.log(prob) + (1.0 - target) * torch.log(1.0 - prob))
    loss = torch.where(target != -1.0, weight * bce, torch.zeros(prob.shape, device=dev))
    enhance = torch.zeros((), device=dev)
    if il is not None and il.enhance_on_new:                             # :380-384
        new_cols = prob[background, il.past:]
        hot = new_cols > 0.05
        if int(hot.sum()) != 0:
            enhance = torch.pow(new_cols[hot], 2).sum()
    norm = torch.clamp(npos.float(), min=1.0)
    bg = loss[target == 0.0].sum() / norm
    fg = loss[is_one].sum() / norm
    if int(npos) == 0:                                   # host sync, as in the reference
        return bg, fg, torch.zeros((), device=dev), enhance
    aw, ah, acx, acy = (g[positive] for g in geom)
    rows = matched[positive, :]
    gw = rows[:, 2] - rows[:, 0]
    gh = rows[:, 3] - rows[:, 1]
    gcx = rows[:, 0] + 0.5 * gw
    gcy = rows[:, 1] + 0.5 * gh
    gw = torch.clamp(gw, min=1)
    gh = torch.clamp(gh, min=1)
    t = torch.stack(((gcx - acx) / aw, (gcy - acy) / ah, torch.log(gw / aw), torch.log(gh / ah))).t()
    t = t / torch.tensor([[0.1, 0.1, 0.2, 0.2]], device=dev)
    diff = torch.abs(t - reg[positive, :])
    sl1 = torch.where(diff <= 1.0 / 9.0, 0.5 * 9.0 * torch.pow(diff, 2), diff - 0.5 / 9.0)
    return bg, fg, sl1.mean(), enhance


def focal_loss(classifications, regressions, anchors, annotations, alpha=0.25, gamma=2.0, il=None):
    """-> (bg[N], fg[N], reg_loss[1]) -- plus the enhance_on_new sum as a 4th value when `il` (ILFlags) is given;
    differentiable through torch autograd.  FocalLoss.forward for state 0 (il=None) or an incremental state."""
    dev = classifications.device
    anchor = anchors[0]
    aw = anchor[:, 2] - anchor[:, 0]
    ah = anchor[:, 3] - anchor[:, 1]
    geom = (aw, ah, anchor[:, 0] + 0.5 * aw, anchor[:, 1] + 0.5 * ah)
    bgs, fgs, regs = [], [], []
    enhance = torch.zeros((), device=dev)
    for j in range(classifications.shape[0]):
        rows = annotations[j]
        gt = rows[rows[:, 4] != -1]
        prob = torch.clamp(classifications[j], 1e-4, 1.0 - 1e-4)
        if gt.shape[0] == 0:
            # image without GT: (1 - alpha) weighting, no normaliser (quirk Q1)
            w = (1.0 - torch.full(prob.shape, alpha, device=dev)) * torch.pow(prob, gamma)
            bgs.append((w * -torch.log(1.0 - prob)).sum())
            fgs.append(torch.zeros((), device=dev))
            regs.append(torch.zeros((), device=dev))
            continue
        bg, fg, rg, enh = _image_terms(prob, regressions[j], anchor, geom, gt, alpha, gamma, il)
        bgs.append(bg)
        fgs.append(fg)
        regs.append(rg)
        enhance = enhance + enh
    out = (torch.stack(bgs), torch.stack(fgs), torch.stack(regs).mean(dim=0, keepdim=True))
    return out if il is None else out + (enhance,)


def decode_boxes(anchors, deltas):
    """anchors [1,A,4], deltas [N,A,4] -> boxes [N,A,4]; std (0.1,0.1,0.2,0.2), mean 0."""
    w = anchors[:, :, 2] - anchors[:, :, 0]
    h = anchors[:, :, 3] - anchors[:, :, 1]
    cx = anchors[:, :, 0] + 0.5 * w
    cy = anchors[:, :, 1] + 0.5 * h
    dx = deltas[:, :, 0] * 0.1 + 0.0
    dy = deltas[:, :, 1] * 0.1 + 0.0
    dw = deltas[:, :, 2] * 0.2 + 0.0
    dh = deltas[:, :, 3] * 0.2 + 0.0
    pcx = cx + dx * w
    pcy = cy + dy * h
    pw = torch.exp(dw) * w
    ph = torch.exp(dh) * h
    return torch.stack([pcx - 0.5 * pw, pcy - 0.5 * ph, pcx + 0.5 * pw, pcy + 0.5 * ph], dim=2)


def clip_boxes(boxes, height, width):
    boxes[:, :, 0] = torch.clamp(boxes[:, :, 0], min=0)
    boxes[:, :, 1] = torch.clamp(boxes[:, :, 1], min=0)
    boxes[:, :, 2] = torch.clamp(boxes[:, :, 2], max=width)
    boxes[:, :, 3] = torch.clamp(boxes[:, :, 3], max=height)
    return boxes


def predict(logits, regressions, anchors, height, width, image=0, score_threshold=0.05, iou_threshold=0.5):
    """Eval-mode detection output of ONE image -> (scores, labels, boxes), score-descending."""
    from torchvision.ops import batched_nms
    probs = torch.sigmoid(logits)
    boxes = clip_boxes(decode_boxes(anchors, regressions), height, width)
    scores, labels = probs[image].max(dim=1)
    keep = scores > score_threshold
    scores, labels, cand = scores[keep], labels[keep], boxes[image][keep]
    if scores.numel() == 0:
        return scores, labels, cand
    order = batched_nms(cand, scores, labels, iou_threshold)
    return scores[order], labels[order], cand[order]
