"""Golden vectors for SURVEY 8 row a9 (pseudo-label GT format), from the unmodified reference (build container only):
  retinanet/dataloader.py collater (:327-364) on per-image annotation arrays that already contain merged pseudo rows,
and the torch-CPU statements of the pseudo-label post-filter (IL_method/persuado_label.py:52-81: score > 0.7, boxes / scale,
max IoU with the real GT < 0.35 via the reference's calc_iou on float64 annotations, xyxy -> xywh)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import make_golden as mg  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    mg.install_cpu_shim()
    from retinanet.dataloader import collater
    from retinanet.losses import calc_iou
    rng = np.random.default_rng(77)
    # --- collater: three images, 2 / 0 / 5 annotation rows (real rows then pseudo rows, fp64 xyxy+label) ---
    annots = []
    for g in (2, 0, 5):
        x1, y1 = rng.uniform(0, 80, g), rng.uniform(0, 60, g)
        a = np.stack([x1, y1, x1 + rng.uniform(5, 40, g), y1 + rng.uniform(5, 40, g), rng.integers(0, 6, g).astype(np.float64)], 1).reshape(-1, 5)
        annots.append(a)
    data = [{'img': torch.zeros(64 + 32 * i, 96, 3), 'annot': torch.from_numpy(a), 'scale': 1.0 + 0.1 * i, 'num_persuado_labels': -1}
            for i, a in enumerate(annots)]
    batch = collater(data)
    empty = collater([{'img': torch.zeros(64, 96, 3), 'annot': torch.zeros(0, 5, dtype=torch.float64), 'scale': 1.0, 'num_persuado_labels': -1}])
    d = {'collated': batch['annot'].numpy(), 'collated_empty': empty['annot'].numpy()}
    for i, a in enumerate(annots):
        d['annot%d' % i] = a
    # --- pseudo-label post filter (persuado_label.py:52-81) ---
    K, G = 40, 4
    scale = 1.2345
    x1, y1 = rng.uniform(0, 300, K), rng.uniform(0, 200, K)
    boxes = np.stack([x1, y1, x1 + rng.uniform(10, 120, K), y1 + rng.uniform(10, 120, K)], 1).astype(np.float32)
    scores = rng.uniform(0.4, 1.0, K).astype(np.float32)
    labels = rng.integers(0, 15, K)
    gx, gy = rng.uniform(0, 300, G), rng.uniform(0, 200, G)
    gt = np.stack([gx, gy, gx + rng.uniform(30, 150, G), gy + rng.uniform(30, 150, G), rng.integers(15, 16, G).astype(np.float64)], 1)
    gt = np.concatenate([gt, -np.ones((2, 5))])                       # padded rows as the dataset delivers them
    boxes[:3] = (gt[:3, :4] * 1.0).astype(np.float32)                 # some predictions coincide with real GT
    pb, ps, pl = torch.from_numpy(boxes), torch.from_numpy(scores), torch.from_numpy(labels)
    annotations = torch.from_numpy(gt.copy()).unsqueeze(0)
    mask = ps > 0.7
    pb = pb[mask] / scale
    ps, pl = ps[mask], pl[mask]
    annotation = annotations[0, ...]
    annotation = annotation[annotation[..., -1] != -1]
    gd_boxes = annotation[..., :4]
    gd_boxes /= scale
    iou = calc_iou(pb, gd_boxes)
    max_iou, _ = iou.max(dim=1)
    keep = max_iou < 0.35
    pb, ps, pl = pb[keep], ps[keep], pl[keep]
    pb = pb.clone()
    pb[:, 2] -= pb[:, 0]
    pb[:, 3] -= pb[:, 1]
    d.update(pl_boxes=boxes, pl_scores=scores, pl_labels=labels, pl_gt=gt, pl_scale=scale, out_boxes=pb.numpy(), out_scores=ps.numpy(),
             out_labels=pl.numpy(), out_max_iou=max_iou.numpy())
    np.savez_compressed(os.path.join(OUT, 'a9_pseudo_labels.npz'), **d)
    print('collated', batch['annot'].shape, empty['annot'].shape, 'pseudo kept', pb.shape, pb.dtype, max_iou.dtype)


if __name__ == '__main__':
    main()
