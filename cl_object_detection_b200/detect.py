"""Drop-in for the eval half of the detection head: retinanet/utils.py BBoxTransform (:82-126) and ClipBoxes
(:129-144), the part of ResNet.predict after self.forward (retinanet/model.py:507-550), Labeler.predict
(IL_method/persuado_label.py:99-127) and torchvision.ops.nms / batched_nms as called there (model.py:540).

The pipeline is decode_filter (class max + sigmoid + threshold + decode of the survivors only) -> ordering
(+ optional radix-select top-k) -> 64x64 bitmask NMS -> gather, all in libcldet.so.  The reference handles one
image per call; `detect_batch` runs a whole batch through the same kernels at once.
"""
import torch
import torch.nn as nn

from . import _lib, _ops
from .losses import _check_cuda_f32, _stream

NMS_MODE_TORCHVISION = 0   # coordinate trick unless 4*K > limit (torchvision.ops.boxes.batched_nms)
NMS_MODE_TRICK = 1
NMS_MODE_VANILLA = 2
CUDA_VANILLA_NUMEL_LIMIT = 100_000   # torchvision 0.26 on CUDA; 4000 on CPU
_CAND_BYTES = 32


class BBoxTransform(nn.Module):
    """forward(boxes[1,A,4], deltas[N,A,4]) -> [N,A,4]; mean 0, std (0.1,0.1,0.2,0.2) (utils.py:82-126)."""

    def __init__(self, mean=None, std=None):
        super().__init__()
        if mean is not None or std is not None:
            raise NotImplementedError('only the reference defaults mean=0, std=(0.1,0.1,0.2,0.2) are implemented')

    def forward(self, boxes, deltas):
        return decode_boxes(boxes, deltas, clip_to=None)


class ClipBoxes(nn.Module):
    """forward(boxes, img): clamps IN PLACE like the reference (utils.py:134-144) and returns boxes."""

    def __init__(self, width=None, height=None):
        super().__init__()

    def forward(self, boxes, img):
        height, width = int(img.shape[2]), int(img.shape[3])
        b = _check_cuda_f32('boxes', boxes)
        if b.data_ptr() != boxes.data_ptr():
            raise ValueError('ClipBoxes works in place and needs a contiguous float32 CUDA tensor')
        with torch.cuda.device(b.device):
            _lib.check(_lib.load().cldet_clip_boxes(b.data_ptr(), b.numel() // 4, height, width, _stream()))
        return boxes


def decode_boxes(anchors, deltas, clip_to=None):
    """BBoxTransform (+ ClipBoxes when clip_to=(height, width))."""
    anc = _check_cuda_f32('anchors', anchors).reshape(-1, 4)
    d = _check_cuda_f32('deltas', deltas)
    if d.dim() != 3 or d.shape[2] != 4 or d.shape[1] != anc.shape[0]:
        raise ValueError('deltas must be [N,A,4] with A matching the anchors')
    h, w = (0, 0) if clip_to is None else (int(clip_to[0]), int(clip_to[1]))
    with torch.cuda.device(d.device):
        out = torch.empty_like(d)
        _lib.check(_lib.load().cldet_decode_boxes(anc.data_ptr(), d.data_ptr(), d.shape[0], d.shape[1],
                                                  0 if clip_to is None else 1, h, w, out.data_ptr(), _stream()))
    return out


def _aligned(t):
    return t if t.data_ptr() % 16 == 0 else t.clone()


def batched_nms(boxes, scores, idxs, iou_threshold, mode=NMS_MODE_TORCHVISION,
                vanilla_numel_limit=CUDA_VANILLA_NUMEL_LIMIT):
    """torchvision.ops.batched_nms(boxes[K,4], scores[K], idxs[K], thr) -> int64 keep indices, score-descending
    (ties by ascending index).  idxs=None gives torchvision.ops.nms."""
    b = _aligned(_check_cuda_f32('boxes', boxes).reshape(-1, 4))
    s = _check_cuda_f32('scores', scores)
    k = b.shape[0]
    if s.shape[0] != k:
        raise ValueError('boxes and scores disagree')
    if k == 0:
        return torch.empty(0, dtype=torch.int64, device=b.device)
    keep, count = _ops.load().batched_nms(b, s, idxs, float(iou_threshold), int(mode), int(vanilla_numel_limit))
    return keep[:int(count.item())]


def nms(boxes, scores, iou_threshold):
    return batched_nms(boxes, scores, None, iou_threshold, mode=NMS_MODE_TRICK)


def detect_batch(cls, regressions, anchors, height, width, is_logits=True, score_thresh=0.05, iou_threshold=0.5,
                 pre_nms_topk=0, nms_mode=NMS_MODE_TORCHVISION, vanilla_numel_limit=CUDA_VANILLA_NUMEL_LIMIT,
                 return_padded=False):
    """Detection output for EVERY image of the batch.

    cls [N,A,C] logits (is_logits=True: sigmoid is applied like model.py:507) or probabilities (Labeler.predict);
    regressions [N,A,4]; anchors [1,A,4].  Returns a list of (scores[K'], labels[K'] int64, boxes[K',4]) per image in NMS
    order, or with return_padded=True the dense tensors (scores[N,cap], labels[N,cap], boxes[N,cap,4], counts[N]) without
    any per-image slicing.  pre_nms_topk > 0 keeps only the best k candidates per image before NMS (the reference has
    no such stage; 0 reproduces it) and makes the whole pipeline free of host synchronisation until the final counts.
    """
    c = _check_cuda_f32('classifications', cls)
    r = _aligned(_check_cuda_f32('regressions', regressions))
    anc = _aligned(_check_cuda_f32('anchors', anchors).reshape(-1, 4))
    if c.dim() != 3 or r.dim() != 3 or r.shape[2] != 4 or r.shape[:2] != c.shape[:2] or anc.shape[0] != c.shape[1]:
        raise ValueError('expected cls [N,A,C], regressions [N,A,4], anchors [1,A,4]')
    n = c.shape[0]
    # one call into the C++ op layer (csrc/cldet_torch.cpp `detect`): allocation + K4 -> K5 -> K6 -> gather launches, no sync
    topk = int(pre_nms_topk) if pre_nms_topk and pre_nms_topk > 0 else 0
    args = (c, r, anc, int(height), int(width), bool(is_logits), float(score_thresh), float(iou_threshold), topk, int(nms_mode),
            int(vanilla_numel_limit))
    op = _ops.load().detect
    scores, labels, boxes, keep_counts, cand_counts = op(*args, 0)
    if topk and return_padded:
        return scores, labels, boxes, keep_counts
    # ONE device->host read: the kept counts (needed to slice) together with the candidate counts.  Without a top-k the op
    # sized its buffers optimistically; an image with more candidates than that (an untrained model) repeats the call at the
    # exact size -- the reference-faithful mode needs no other synchronisation.
    host = torch.stack((keep_counts, cand_counts)).cpu()
    if not topk and int(host[1].max()) > scores.shape[1]:
        scores, labels, boxes, keep_counts, cand_counts = op(*args, int(host[1].max()))
        host = torch.stack((keep_counts, cand_counts)).cpu()
    if return_padded:
        return scores, labels, boxes, keep_counts
    kc = host[0].tolist()
    return [(scores[j, :kc[j]], labels[j, :kc[j]], boxes[j, :kc[j]]) for j in range(n)]


def detect_batch_head(cls_levels, reg_levels, anchors, height, width, is_logits=True, score_thresh=0.05, iou_threshold=0.5,
                      pre_nms_topk=0, nms_mode=NMS_MODE_TORCHVISION, vanilla_numel_limit=CUDA_VANILLA_NUMEL_LIMIT,
                      return_padded=False):
    """detect_batch on the head's RAW conv outputs (eval side; beyond the reference's signature).

    cls_levels / reg_levels: the five per-level results of the classification / regression output convolutions,
    [N, 9*C, H_l, W_l] logits (or probabilities with is_logits=False) and [N, 36, H_l, W_l], contiguous NCHW -- what
    ClassificationModel / RegressionModel hold BEFORE their permute + contiguous + view (retinanet/model.py:125-130, 170-184)
    and before ResNet.forward's torch.cat (model.py:472-474).  anchors = Anchors()(img) for the (height, width) input.
    Same results as detect_batch on the reshaped + concatenated tensors, bit for bit."""
    cl = [_check_cuda_f32('cls_levels[%d]' % i, t) for i, t in enumerate(cls_levels)]
    rl = [_check_cuda_f32('reg_levels[%d]' % i, t) for i, t in enumerate(reg_levels)]
    anc = _aligned(_check_cuda_f32('anchors', anchors).reshape(-1, 4))
    h, w = int(height), int(width)
    if len(cl) != 5 or len(rl) != 5 or cl[0].dim() != 4 or cl[0].shape[1] % 9 != 0:
        raise ValueError('expected the 5 pyramid levels (3..7) of both heads, classification [N, 9*C, H_l, W_l]')
    n, nc = cl[0].shape[0], cl[0].shape[1] // 9
    a = 0
    for l in range(5):
        hl, wl = (h + 2 ** (l + 3) - 1) // 2 ** (l + 3), (w + 2 ** (l + 3) - 1) // 2 ** (l + 3)
        if tuple(cl[l].shape) != (n, 9 * nc, hl, wl) or tuple(rl[l].shape) != (n, 36, hl, wl):
            raise ValueError('level %d: expected cls [%d,%d,%d,%d] and reg [%d,36,%d,%d] for a %dx%d input'
                             % (l + 3, n, 9 * nc, hl, wl, n, hl, wl, h, w))
        a += 9 * hl * wl
    if anc.shape[0] != a:
        raise ValueError('anchors must be [1,%d,4] for a %dx%d input' % (a, h, w))
    dev = cl[0].device
    lib = _lib.load()
    pc, pr = _lib.ptr_array(cl), _lib.ptr_array(rl)

    def run_filter(cand, keys, counts, st):
        _lib.check(lib.cldet_decode_filter_head(pc, pr, 5, h, w, int(bool(is_logits)), anc.data_ptr(), n, nc, float(score_thresh),
                                                cand.data_ptr(), keys.data_ptr(), a, counts.data_ptr(), st))

    return _detect_pipeline(run_filter, n, a, dev, iou_threshold, pre_nms_topk, nms_mode, vanilla_numel_limit, return_padded)


def _detect_pipeline(run_filter, n, a, dev, iou_threshold, pre_nms_topk, nms_mode, vanilla_numel_limit, return_padded):
    """Candidate filter (K4 / K4h, supplied by the caller) -> K5 select + rank sort -> K6 NMS -> gather."""
    lib = _lib.load()
    topk = int(pre_nms_topk) if pre_nms_topk and pre_nms_topk > 0 else 0
    with torch.cuda.device(dev):
        st = _stream()
        counts = torch.zeros(n, dtype=torch.int32, device=dev)
        cand = torch.empty((n, a, _CAND_BYTES), dtype=torch.uint8, device=dev)
        keys = torch.empty((n, a), dtype=torch.int64, device=dev)
        run_filter(cand, keys, counts, st)
        if topk:
            max_count, cap = a, min(topk, a)
        else:
            max_count = int(counts.max().item())     # the one mid-pipeline sync of the reference-faithful mode
            cap = max_count
        if cap == 0:
            empty = (torch.empty(0, device=dev), torch.empty(0, dtype=torch.int64, device=dev), torch.empty((0, 4), device=dev))
            if return_padded:
                return (torch.empty((n, 0), device=dev), torch.empty((n, 0), dtype=torch.int64, device=dev),
                        torch.empty((n, 0, 4), device=dev), torch.zeros(n, dtype=torch.int32, device=dev))
            return [empty for _ in range(n)]
        sorted_c = torch.empty((n, cap, _CAND_BYTES), dtype=torch.uint8, device=dev)
        sorted_counts = torch.empty(n, dtype=torch.int32, device=dev)
        sws_bytes = lib.cldet_sort_workspace_bytes(n, max_count, topk)
        sws = torch.empty(sws_bytes, dtype=torch.uint8, device=dev)
        _lib.check(lib.cldet_sort_candidates(cand.data_ptr(), keys.data_ptr(), counts.data_ptr(), n, a, max_count, topk,
                                             sorted_c.data_ptr(), cap, sorted_counts.data_ptr(), sws.data_ptr(), sws_bytes, st))
        nws_bytes = lib.cldet_nms_workspace_bytes(n, cap)
        nws = torch.empty(nws_bytes, dtype=torch.uint8, device=dev)
        keep = torch.empty((n, cap), dtype=torch.int32, device=dev)
        keep_counts = torch.empty(n, dtype=torch.int32, device=dev)
        scores = torch.empty((n, cap), dtype=torch.float32, device=dev)
        labels = torch.empty((n, cap), dtype=torch.int64, device=dev)
        boxes = torch.empty((n, cap, 4), dtype=torch.float32, device=dev)
        _lib.check(lib.cldet_nms_gather_sorted(sorted_c.data_ptr(), sorted_counts.data_ptr(), n, cap, cap, float(iou_threshold),
                                               int(nms_mode), int(vanilla_numel_limit), keep.data_ptr(), keep_counts.data_ptr(),
                                               scores.data_ptr(), labels.data_ptr(), boxes.data_ptr(), nws.data_ptr(), nws_bytes, st))
        if return_padded:
            return scores, labels, boxes, keep_counts
        kc = keep_counts.cpu().tolist()
        return [(scores[j, :kc[j]], labels[j, :kc[j]], boxes[j, :kc[j]]) for j in range(n)]


def predict_from_head(classification, regression, anchors, img_batch, thresh=None, **kw):
    """The part of ResNet.predict after self.forward and the optional BiC correction (model.py:507-550): logits in,
    [scores, labels(int64), boxes] of image 0 out, score-descending.  `thresh` is validated and then ignored exactly
    like the reference (model.py:524-530 overwrites it with 0.05)."""
    if thresh is not None and len(thresh) != classification.shape[2]:
        raise ValueError('Parameter Thresh  must contain {} elements!'.format(classification.shape[2]))
    h, w = int(img_batch.shape[2]), int(img_batch.shape[3])
    s, l, b = detect_batch(classification[:1], regression[:1], anchors, h, w, is_logits=True, **kw)[0]
    return [s, l, b]


def predict(model, img_batch, thresh=None, method=None, bic=None, **kw):
    """Drop-in for ResNet.predict (model.py:494): `ResNet.predict = cl_object_detection_b200.detect.predict`."""
    classification, regression, anchors = model.forward(img_batch, return_feat=False, return_anchor=True, enable_act=False)
    if bic:
        classification = bic.bic_correction(classification)
    return predict_from_head(classification, regression, anchors, img_batch, thresh, **kw)


def labeler_predict(img_batch, classifications, regressions, anchors, **kw):
    """Labeler.predict (persuado_label.py:99-127): probabilities in, (scores, boxes, labels) of image 0 out; empty
    result = three empty float tensors like the reference's torch.tensor([])."""
    h, w = int(img_batch.shape[2]), int(img_batch.shape[3])
    s, l, b = detect_batch(classifications[:1], regressions[:1], anchors, h, w, is_logits=False, **kw)[0]
    if s.shape[0] == 0:
        return torch.tensor([]), torch.tensor([]), torch.tensor([])
    return s, b, l


def coco_results(padded, scales, image_ids=None, label_to_coco_label=None, score_threshold=0.05):
    """SURVEY 8(f) row f4 -- the evaluator's per-detection Python loop (evaluator.py:329-361) for a whole batch:
    boxes / scale, xyxy -> xywh, keep score >= threshold, on the device; ONE device->host copy of the compact records.

    padded: (scores[N,cap], labels[N,cap], boxes[N,cap,4], counts[N]) from detect_batch(..., return_padded=True);
    scales: per-image resize scale (sequence or tensor).  Returns the reference's list of dicts
    {'image_id', 'category_id', 'score', 'bbox'} (image_ids / label_to_coco_label default to identity).
    """
    scores, labels, boxes, counts = padded
    n, cap = scores.shape
    dev = scores.device
    if cap == 0:
        return []
    lib = _lib.load()
    with torch.cuda.device(dev):
        sc = torch.as_tensor(scales, dtype=torch.float32).to(dev).contiguous()
        if sc.numel() != n:
            raise ValueError('one scale per image expected')
        rec = torch.empty((n * cap, 8), dtype=torch.int32, device=dev)
        offsets = torch.empty(n + 1, dtype=torch.int32, device=dev)
        _lib.check(lib.cldet_coco_results(scores.contiguous().data_ptr(), labels.contiguous().data_ptr(),
                                          boxes.contiguous().data_ptr(), counts.contiguous().data_ptr(), sc.data_ptr(), n, cap,
                                          float(score_threshold), rec.data_ptr(), offsets.data_ptr(), _stream()))
        total = int(offsets[n].item())
        host = rec[:total].cpu()
    ints = host.numpy()
    floats = host.view(torch.float32).numpy()
    out = []
    for k in range(total):
        j, lab = int(ints[k, 0]), int(ints[k, 1])
        out.append({'image_id': j if image_ids is None else image_ids[j],
                    'category_id': lab if label_to_coco_label is None else label_to_coco_label(lab),
                    'score': float(floats[k, 2]),
                    'bbox': [float(floats[k, 3]), float(floats[k, 4]), float(floats[k, 5]), float(floats[k, 6])]})
    return out
