"""Shared helpers for the parity tests: golden loading, oracle params, seeded synthetic inputs."""
import os

import numpy as np

from oracle import head_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')

FOCAL_CASES = ['state0_voc', 'state0_allvalid', 'state0_allempty', 'state0_gamma15', 'il_default_pseudo',
               'il_ignore_past', 'il_new_ignore_past', 'il_distill_enhance', 'il_decrease_positive',
               'il_decrease_by_iou', 'il_all_flags']


def load(name):
    return np.load(os.path.join(GOLDEN, name + '.npz'), allow_pickle=False)


def golden_params(g, cls=O.OracleParams):
    """Rebuild the params object a focal fixture was generated with."""
    kw = {}
    for k, v in zip(g['params_keys'], g['params_vals']):
        k = str(k)
        kw[k] = float(v) if k in ('alpha', 'gamma', 'decrease_positive') else bool(v)
    return cls(num_past_class=[int(x) for x in g['num_past_class']], **kw)


def rel_err(a, b, floor=0.0):
    """max |a-b| / max(|b|, floor) elementwise."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    if a.size == 0:
        return 0.0
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor if floor > 0 else 1e-300)))


def synth_gt(rng, n_img, gmax, height, width, num_classes, empty=(), exact=None, pseudo_split=None):
    """SURVEY 8(d) GT generator: x1,y1 ~ U(0,0.7*W/H), w,h ~ U(16, 0.3*W/H+16), pad rows = -1."""
    ann = np.full((n_img, gmax, 5), -1.0, dtype=np.float32)
    for j in range(n_img):
        if j in empty:
            continue
        g = gmax if exact else int(rng.integers(1, gmax + 1))
        x1 = rng.uniform(0, 0.7 * width, g)
        y1 = rng.uniform(0, 0.7 * height, g)
        w = rng.uniform(16, 0.3 * width + 16, g)
        h = rng.uniform(16, 0.3 * height + 16, g)
        if pseudo_split is None:
            lab = rng.integers(0, num_classes, g)
        else:  # first rows = new classes (>= past), remaining rows = pseudo labels of old classes
            k = max(1, g // 2)
            lab = np.concatenate([rng.integers(pseudo_split, num_classes, k), rng.integers(0, pseudo_split, g - k)])
        ann[j, :g] = np.stack([x1, y1, x1 + w, y1 + h, lab.astype(np.float64)], 1).astype(np.float32)
    return ann


def synth_head(rng, n_img, num_anchors, num_classes, mu=-4.0, sigma=2.0, reg_sigma=1.0):
    logits = rng.normal(mu, sigma, (n_img, num_anchors, num_classes)).astype(np.float32)
    probs = (1.0 / (1.0 + np.exp(-logits.astype(np.float64)))).astype(np.float32)
    reg = rng.normal(0, reg_sigma, (n_img, num_anchors, 4)).astype(np.float32)
    return logits, probs, reg


# worst errors observed by the tolerance helpers during a session (written out by conftest.pytest_sessionfinish)
OBSERVED = {}
