"""pseudo-label GT -- the INPUT FORMAT of the loss path and the generator's post-filter.

* merge_pseudo_labels / collate_annotations: host-side mirrors of retinanet/dataloader.py:119-147 and :348-359 -- pseudo
  boxes (COCO xywh, old-class ids) are appended AFTER the real GT rows, everything becomes xyxy, and the batch is padded
  with -1 rows to [N, Gmax, 5] float32.  The assign kernel treats pseudo rows as ordinary GT (the reference's `progress`
  down-weighting is a no-op, losses.py:388-392).
* filter_pseudo_labels: the tail of Labeler.get_persuado_label (IL_method/persuado_label.py:52-81) on the device: keep
  score > 0.7, divide by the resize scale, drop boxes whose max fp64 IoU with a real GT is >= 0.35, xyxy -> xywh.
"""
import numpy as np
import torch

from . import _lib
from .losses import _DeviceGuard, _stream

DEFAULT_SCORE_THRESHOLD = 0.7      # persuado_label.py:12
DEFAULT_IOU_THRESHOLD = 0.35       # persuado_label.py:13


def merge_pseudo_labels(real_xywh_label, pseudo_xywh_label):
    """[g,5] real rows then [p,5] pseudo rows, each (x, y, w, h, label) in fp64 -> [g+p,5] (x1, y1, x2, y2, label)."""
    rows = [np.asarray(r, dtype=np.float64).reshape(-1, 5) for r in (real_xywh_label, pseudo_xywh_label)]
    ann = np.concatenate(rows, axis=0)
    ann[:, 2] = ann[:, 0] + ann[:, 2]
    ann[:, 3] = ann[:, 1] + ann[:, 3]
    return ann


def collate_annotations(annots):
    """List of per-image [g_i,5] arrays/tensors -> float32 tensor [N, max(g_i) or 1, 5] padded with -1 (collater)."""
    arrays = [a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a) for a in annots]
    gmax = max((a.shape[0] for a in arrays), default=0)
    out = torch.full((len(arrays), max(gmax, 1), 5), -1.0, dtype=torch.float32)
    for i, a in enumerate(arrays):
        if a.shape[0] > 0:
            out[i, :a.shape[0]] = torch.from_numpy(np.ascontiguousarray(a)).to(torch.float32)
    return out


def filter_pseudo_labels(scores, boxes, labels, annotations, scale, score_threshold=DEFAULT_SCORE_THRESHOLD,
                         iou_threshold=DEFAULT_IOU_THRESHOLD):
    """scores[K], boxes[K,4] (xyxy, resized image frame), labels[K]: Labeler.predict's outputs; annotations [G,5] the image's
    real GT in the same frame (rows with label -1 are padding), any float dtype; scale: the resize factor.
    Returns (scores, boxes_xywh float32 in the ORIGINAL image frame, labels) of the accepted pseudo labels."""
    if scores.numel() == 0:
        return scores, boxes.reshape(0, 4), labels
    dev = boxes.device
    if dev.type != 'cuda':
        raise RuntimeError('filter_pseudo_labels runs on CUDA tensors only')
    keep = scores > score_threshold                                   # :54
    b = boxes[keep].to(torch.float32) / scale                         # :55  (true fp32 division)
    s, l = scores[keep], labels[keep]
    if b.shape[0] > 0:
        ann = annotations.to(dev)
        ann = ann[ann[..., -1] != -1]                                 # :63
        if ann.shape[0] > 0:
            gd = (ann[..., :4].to(torch.float64) / scale).contiguous()    # :64-66
            bd = b.to(torch.float64).contiguous()
            with _DeviceGuard(dev):
                mx = torch.empty(bd.shape[0], dtype=torch.float64, device=dev)
                _lib.check(_lib.load().cldet_iou_max_f64(bd.data_ptr(), bd.shape[0], gd.data_ptr(), gd.shape[0], mx.data_ptr(),
                                                         None, _stream()))
            ok = mx < iou_threshold                                   # :72
            b, s, l = b[ok], s[ok], l[ok]
    b = b.clone()
    if b.shape[0] > 0:                                                # :82-84  xyxy -> xywh
        b[:, 2] -= b[:, 0]
        b[:, 3] -= b[:, 1]
    return s, b, l
