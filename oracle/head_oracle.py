"""CPU oracle for the detection-head hot path (TEST INFRASTRUCTURE -- not product code).

This file is a numpy restatement of the reference's algorithm for the path named by
BASELINE.json:north_star.  It exists only to CHECK the CUDA path: the only importers
allowed are tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs.  Nothing under cl_object_detection_b200/ imports it, and the product path raises
when the CUDA library is missing instead of falling back to this file.

Parity pin: the reference ships no tests or golden vectors, so this
oracle is pinned against outputs of the reference ITSELF, imported from /root/reference in
the build container by tests/golden/make_golden.py (committed) and stored as small
fixtures under tests/golden/*.npz.  tests/test_oracle_golden.py checks every function here
against those fixtures; tests/test_reference_snapshot.py checks the loss half against the LIVE
reference (byte-code snapshot oracle/_ref, oracle/build_ref.py) on fresh seeded inputs.

Arithmetic notes.  Everything that decides an integer/boolean result (anchors, IoU,
assignment thresholds, argmax, NMS) follows the reference's fp32 op order exactly, one
rounding per op, no FMA contraction -- numpy float32 ufuncs give exactly that.  Sums
(loss reductions) are accumulated in float64 and rounded once; the reference accumulates
in fp32 with torch's own reduction tree, so those agree to ~1e-7 relative, inside the
1e-5 tolerance north_star states.

Each function cites the reference file:line it restates (paths relative to /root/reference).
"""
from __future__ import annotations

import numpy as np

F32 = np.float32


# --------------------------------------------------------------------------------------
# a1  anchors  (retinanet/anchors.py:21-40, 42-73, 109-129)
# --------------------------------------------------------------------------------------
PYRAMID_LEVELS = (3, 4, 5, 6, 7)
RATIOS = (0.5, 1.0, 2.0)
SCALES = (2.0 ** 0, 2.0 ** (1.0 / 3.0), 2.0 ** (2.0 / 3.0))


def base_anchors(base_size: float) -> np.ndarray:
    """9 base boxes (ratio-major, scale-minor) centred on 0, fp64.  anchors.py:42-73."""
    out = np.zeros((9, 4), dtype=np.float64)
    k = 0
    for r in RATIOS:
        for s in SCALES:
            side = base_size * s                 # anchors.py:60
            area = side * side                   # :63
            w = np.sqrt(area / r)                # :66
            h = w * r                            # :67
            out[k] = (0.0 - w * 0.5, 0.0 - h * 0.5, w - w * 0.5, h - h * 0.5)  # :70-71
            k += 1
    return out


def level_shapes(height: int, width: int):
    """Feature-map (H_l, W_l) per level: integer ceil-div.  anchors.py:25."""
    return [((height + 2 ** l - 1) // 2 ** l, (width + 2 ** l - 1) // 2 ** l) for l in PYRAMID_LEVELS]


def anchors_for_image(height: int, width: int) -> np.ndarray:
    """[1, A, 4] float32 anchors; fp64 math, ONE final cast.  anchors.py:21-40, :109-129."""
    chunks = []
    for l, (hl, wl) in zip(PYRAMID_LEVELS, level_shapes(height, width)):
        stride = 2 ** l
        base = base_anchors(2 ** (l + 2))
        sx = (np.arange(wl, dtype=np.float64) + 0.5) * stride
        sy = (np.arange(hl, dtype=np.float64) + 0.5) * stride
        gx, gy = np.meshgrid(sx, sy)             # x fastest, then y
        shifts = np.stack([gx.ravel(), gy.ravel(), gx.ravel(), gy.ravel()], axis=1)  # [K,4]
        chunks.append((shifts[:, None, :] + base[None, :, :]).reshape(-1, 4))
    return np.concatenate(chunks, axis=0).astype(F32)[None]


def num_anchors(height: int, width: int) -> int:
    return 9 * sum(h * w for h, w in level_shapes(height, width))


# --------------------------------------------------------------------------------------
# a2  calc_iou  (retinanet/losses.py:4-21)
# --------------------------------------------------------------------------------------
def calc_iou(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """Pairwise IoU [A,G] in fp32, one rounding per op, same op order as losses.py:4-21."""
    a = np.asarray(a, dtype=F32)
    b = np.asarray(b, dtype=F32)
    area = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    iw = np.minimum(a[:, 2:3], b[:, 2]) - np.maximum(a[:, 0:1], b[:, 0])
    ih = np.minimum(a[:, 3:4], b[:, 3]) - np.maximum(a[:, 1:2], b[:, 1])
    iw = np.maximum(iw, F32(0))
    ih = np.maximum(ih, F32(0))
    ua = ((a[:, 2] - a[:, 0]) * (a[:, 3] - a[:, 1]))[:, None] + area - iw * ih
    ua = np.maximum(ua, F32(1e-8))
    return (iw * ih) / ua


# --------------------------------------------------------------------------------------
# a3/a4  GT filter + assignment  (retinanet/losses.py:287-288, 309-341)
# --------------------------------------------------------------------------------------
ST_BG, ST_POS, ST_IGNORE = 0, 1, 2


def assign(anchors: np.ndarray, annotations_j: np.ndarray):
    """Assignment for ONE image.

    Returns dict(valid=int, state[A] uint8 in {0 bg,1 pos,2 ignore}, argmax[A] int32 (index
    into the COMPACTED GT list, first maximal index = torch.max semantics), iou_max[A] f32,
    label[A] int32 (label of the assigned GT), npos=int, gt = compacted GT rows [G',5]).
    Empty image (no valid GT): valid=0 and the per-anchor arrays are None (losses.py:292).
    """
    ann = np.asarray(annotations_j, dtype=F32)
    gt = ann[ann[:, 4] != F32(-1)]                       # losses.py:287-288
    if gt.shape[0] == 0:
        return dict(valid=0, state=None, argmax=None, iou_max=None, label=None, npos=0, gt=gt)
    iou = calc_iou(anchors, gt[:, :4])                   # :309
    argmax = np.argmax(iou, axis=1).astype(np.int32)     # first max index, like torch.max  :310
    iou_max = iou[np.arange(iou.shape[0]), argmax]
    state = np.full(iou.shape[0], ST_IGNORE, dtype=np.uint8)
    state[iou_max < F32(0.4)] = ST_BG                    # :316
    pos = iou_max >= F32(0.5)                            # :330
    state[pos] = ST_POS
    label = gt[argmax, 4].astype(np.int64).astype(np.int32)   # .long()  :341
    return dict(valid=int(gt.shape[0]), state=state, argmax=argmax, iou_max=iou_max,
                label=label, npos=int(pos.sum()), gt=gt)


# --------------------------------------------------------------------------------------
# params duck type  (preprocessing/params.py:174-178, :17-21)
# --------------------------------------------------------------------------------------
class OracleParams:
    """params[key] -> value or None; params.states[k]['num_past_class'].  CLI defaults of main.py:116-177."""

    DEFAULTS = dict(alpha=0.25, gamma=2.0, distill=False, enhance_on_new=False, ignore_past_class=False,
                    new_ignore_past_class=False, decrease_positive_by_IOU=False, decrease_positive=1.0,
                    persuado_label=False)

    def __init__(self, num_past_class=(0,), **kw):
        self._d = dict(self.DEFAULTS)
        self._d.update(kw)
        self.states = [{'num_past_class': int(n)} for n in num_past_class]

    def __getitem__(self, key):
        return self._d.get(key, None)


# --------------------------------------------------------------------------------------
# a5-a8  FocalLoss forward + analytic backward  (retinanet/losses.py:252-452)
# --------------------------------------------------------------------------------------
def _pow(x, gamma):
    # torch.pow(x, 2.0) is computed as x*x by ATen; other exponents go through powf.
    if float(gamma) == 2.0:
        return x * x
    return np.power(x, F32(gamma), dtype=F32)


def _dpow(x, gamma):
    """d/dx x**gamma as autograd forms it (gamma * x**(gamma-1))."""
    if float(gamma) == 2.0:
        return F32(2.0) * x
    return F32(gamma) * np.power(x, F32(gamma - 1.0), dtype=F32)


def focal_loss(classifications, regressions, anchors, annotations, cur_state, params, progress=-1,
               w_bg=None, w_fg=None, w_reg=1.0, w_enh=1.0, want_grads=True, from_logits=False):
    """Restatement of FocalLoss.forward (losses.py:252-452) plus its autograd backward.

    Inputs as the reference: classifications [N,A,C] fp32 PROBABILITIES, regressions [N,A,4],
    anchors [1,A,4], annotations [N,G,5] (pad rows = -1).  `progress` is accepted and has no
    effect (the statement at losses.py:388-392 writes into a temporary).

    Upstream weights for the backward: w_bg[N], w_fg[N] = dL/d(bg_j), dL/d(fg_j); w_reg =
    dL/d(reg_loss[0]); w_enh = dL/d(enhance_on_new_loss).  Defaults 1/N, 1/N, 1, 1 -- what
    IL_Loss applies without clip_loss (losses.py:584-588).

    Returns dict: bg[N], fg[N], reg_loss[1], npos[N], [bg_masks bool[M,A]], [enhance_on_new_loss],
    grad_cls[N,A,C], grad_reg[N,A,4], and per-image assignment dicts under 'assign'.
    """
    cls = np.asarray(classifications, dtype=F32)
    if from_logits:
        # caller pattern of IL_Loss: focal_loss(self.classifier_act(classification), ...) with classifier_act = Sigmoid
        # (losses.py:566, 633-647); grad_cls is then dL/dlogits = dL/dp * (1 - p) * p (SigmoidBackward)
        out = focal_loss(sigmoid(cls), regressions, anchors, annotations, cur_state, params, progress, w_bg, w_fg, w_reg,
                         w_enh, want_grads, from_logits=False)
        if want_grads:
            p = sigmoid(cls)
            out['grad_cls'] = (out['grad_cls'] * (F32(1.0) - p)) * p
        return out
    reg = np.asarray(regressions, dtype=F32)
    anc = np.asarray(anchors, dtype=F32)[0]
    ann = np.asarray(annotations, dtype=F32)
    N, A, C = cls.shape
    alpha = F32(params['alpha'])
    gamma = float(params['gamma'])
    incremental = cur_state > 0
    distill = bool(incremental and params['distill'])
    enhance = bool(incremental and params['enhance_on_new'])
    if w_bg is None:
        w_bg = np.full(N, 1.0 / N)
    if w_fg is None:
        w_fg = np.full(N, 1.0 / N)
    w_bg = np.asarray(w_bg, dtype=np.float64)
    w_fg = np.asarray(w_fg, dtype=np.float64)

    aw = anc[:, 2] - anc[:, 0]                      # losses.py:276-280
    ah = anc[:, 3] - anc[:, 1]
    acx = anc[:, 0] + F32(0.5) * aw
    acy = anc[:, 1] + F32(0.5) * ah

    bg = np.zeros(N, dtype=F32)
    fg = np.zeros(N, dtype=F32)
    regl = np.zeros(N, dtype=F32)
    npos_out = np.zeros(N, dtype=np.int32)
    bg_masks = []
    enh = np.float64(0.0)
    gcls = np.zeros_like(cls) if want_grads else None
    greg = np.zeros_like(reg) if want_grads else None
    assigns = []

    lo, hi = F32(1e-4), F32(1.0 - 1e-4)
    for j in range(N):
        p_raw = cls[j]
        p = np.clip(p_raw, lo, hi)                  # :290
        band = (p_raw >= lo) & (p_raw <= hi)        # clamp backward pass-band (inclusive)
        asg = assign(anc, ann[j])
        assigns.append(asg)
        if asg['valid'] == 0:                       # :292-307  (Q1: 1-alpha, NOT normalised)
            one_m_alpha = F32(1.0) - alpha
            fw = one_m_alpha * _pow(p, gamma)
            bce = -np.log(F32(1.0) - p)
            bg[j] = F32(np.sum((fw * bce).astype(np.float64)))
            if want_grads:
                d = one_m_alpha * (_dpow(p, gamma) * bce + _pow(p, gamma) / (F32(1.0) - p))
                gcls[j] = np.where(band, d * F32(w_bg[j]), F32(0))
            continue

        state, label, iou_max = asg['state'], asg['label'], asg['iou_max']
        npos = asg['npos']
        npos_out[j] = npos
        is_bg = state == ST_BG
        is_pos = state == ST_POS
        past = int(params.states[cur_state]['num_past_class'])   # :317

        # dense targets, exactly the values the reference builds (:313-341)
        targets = np.full((A, C), -1, dtype=np.int8)
        if (not incremental) or (incremental and not params['ignore_past_class']):
            targets[is_bg, :] = 0                                    # :319
        else:
            targets[is_bg, past:] = 0                                # :321
            if params['new_ignore_past_class']:
                old_prod = np.sum(p[:, :past], axis=1, dtype=F32)    # :326
                sel = is_bg & (old_prod < F32(0.5))
                targets[sel, :past] = 0                              # :327
        if distill:
            bg_masks.append(~is_pos)                                 # :334-335
        targets[is_pos, :] = 0                                       # :340
        pos_idx = np.nonzero(is_pos)[0]
        if np.any(label[pos_idx] < 0) or np.any(label[pos_idx] >= C):
            raise IndexError('GT label out of range for the class dimension (Q8)')
        targets[pos_idx, label[pos_idx]] = 1                         # :341
        t1 = targets == 1
        t0 = targets == 0

        # focal weight f and df/dp for the t==1 elements (:352-366); others use p
        one = F32(1.0)
        if not incremental:
            f1 = one - p
            df1 = np.full_like(p, -1.0)
        elif params['decrease_positive_by_IOU']:
            f1 = one - p
            df1 = np.full_like(p, -1.0)
            mid = (iou_max <= F32(0.7)) & is_pos                      # :354
            tm = np.zeros((A, C), dtype=bool)
            mid_idx = np.nonzero(mid)[0]
            tm[mid_idx, label[mid_idx]] = True                       # :358-359
            upper = np.clip(iou_max + F32(0.2), lo, hi)[:, None]     # :361
            ge = p >= upper
            f_mid = np.where(ge, F32(1e-4), np.abs(p - upper))       # :362
            df_mid = np.where(ge, F32(0.0), np.sign(p - upper).astype(F32))
            f1 = np.where(tm, f_mid, f1)
            df1 = np.where(tm, df_mid, df1)
        else:
            s = F32(params['decrease_positive'])                      # :365
            pc = np.clip(p, F32(0.0), s)
            f1 = s - pc                                              # :366
            df1 = np.where((p >= F32(0.0)) & (p <= s), F32(-1.0), F32(0.0))

        fwt = np.where(t1, f1, p)
        focal_weight = alpha * _pow(fwt, gamma)                      # :369 (Q1: alpha for both)
        with np.errstate(divide='ignore', invalid='ignore'):
            logp = np.log(p)
            log1mp = np.log(one - p)
        tf = targets.astype(F32)
        bce = -(tf * logp + (one - tf) * log1mp)                     # :370
        cls_loss = focal_weight * bce
        cls_loss = np.where(targets != -1, cls_loss, F32(0.0))       # :374-377
        n = F32(max(float(npos), 1.0))                               # :395 clamp(min=1)
        bg[j] = F32(np.sum(cls_loss[t0].astype(np.float64))) / n     # :395
        fg[j] = F32(np.sum(cls_loss[t1].astype(np.float64))) / n     # :396

        if enhance:                                                  # :380-384
            sub = p[is_bg, past:]
            fn = sub > F32(0.05)
            enh += np.sum((sub[fn] * sub[fn]).astype(np.float64))

        if want_grads:
            # t==1:  alpha*( f^g' * f' * (-ln p) - f^g / p );  t==0:  alpha*( p^g' * (-ln(1-p)) + p^g/(1-p) )
            d1 = alpha * (_dpow(f1, gamma) * df1 * (-logp) - _pow(f1, gamma) / p)
            d0 = alpha * (_dpow(p, gamma) * (-log1mp) + _pow(p, gamma) / (one - p))
            g = np.where(t1, d1 * F32(w_fg[j] / float(n)), np.where(t0, d0 * F32(w_bg[j] / float(n)), F32(0.0)))
            if enhance:
                em = np.zeros((A, C), dtype=bool)
                em[:, past:] = (is_bg[:, None] & (p[:, past:] > F32(0.05)))
                g = g + np.where(em, F32(2.0) * p * F32(w_enh), F32(0.0))
            gcls[j] = np.where(band, g, F32(0.0))

        # regression (:398-442)
        if npos > 0:
            g4 = asg['gt'][asg['argmax'][pos_idx], :4]
            gw = g4[:, 2] - g4[:, 0]
            gh = g4[:, 3] - g4[:, 1]
            gcx = g4[:, 0] + F32(0.5) * gw                           # un-clamped w/h for the centre (Q4)
            gcy = g4[:, 1] + F32(0.5) * gh
            gw = np.maximum(gw, F32(1.0))
            gh = np.maximum(gh, F32(1.0))
            tdx = (gcx - acx[pos_idx]) / aw[pos_idx]
            tdy = (gcy - acy[pos_idx]) / ah[pos_idx]
            tdw = np.log(gw / aw[pos_idx])
            tdh = np.log(gh / ah[pos_idx])
            t = np.stack([tdx, tdy, tdw, tdh], axis=1) / np.array([0.1, 0.1, 0.2, 0.2], dtype=F32)
            e = t - reg[j][pos_idx]
            d = np.abs(e)
            small = d <= F32(1.0 / 9.0)
            l = np.where(small, F32(4.5) * (d * d), d - F32(0.5 / 9.0))
            regl[j] = F32(np.sum(l.astype(np.float64)) / (4.0 * npos))
            if want_grads:
                dl = np.where(small, F32(9.0) * d, F32(1.0)) * (-np.sign(e).astype(F32))
                greg[j][pos_idx] = dl * F32(float(w_reg) / N / (4.0 * npos))

    out = dict(bg=bg, fg=fg, reg_loss=np.array([np.mean(regl.astype(np.float64))], dtype=F32),
               reg_per_image=regl, npos=npos_out, assign=assigns)
    if distill:
        out['bg_masks'] = np.stack(bg_masks) if bg_masks else np.zeros((0, A), dtype=bool)
    if enhance:
        out['enhance_on_new_loss'] = F32(enh)
    if want_grads:
        out['grad_cls'] = gcls
        out['grad_reg'] = greg
    return out


# --------------------------------------------------------------------------------------
# a10/a11  decode + clip  (retinanet/utils.py:102-126, 134-144)
# --------------------------------------------------------------------------------------
def bbox_transform(anchors, deltas):
    """BBoxTransform.forward with mean 0, std (0.1,0.1,0.2,0.2).  utils.py:102-126."""
    b = np.asarray(anchors, dtype=F32)
    d = np.asarray(deltas, dtype=F32)
    std = np.array([0.1, 0.1, 0.2, 0.2], dtype=F32)
    zero = F32(0.0)
    w = b[:, :, 2] - b[:, :, 0]
    h = b[:, :, 3] - b[:, :, 1]
    cx = b[:, :, 0] + F32(0.5) * w
    cy = b[:, :, 1] + F32(0.5) * h
    dx = d[:, :, 0] * std[0] + zero
    dy = d[:, :, 1] * std[1] + zero
    dw = d[:, :, 2] * std[2] + zero
    dh = d[:, :, 3] * std[3] + zero
    pcx = cx + dx * w
    pcy = cy + dy * h
    pw = np.exp(dw) * w
    ph = np.exp(dh) * h
    return np.stack([pcx - F32(0.5) * pw, pcy - F32(0.5) * ph, pcx + F32(0.5) * pw, pcy + F32(0.5) * ph], axis=2)


def clip_boxes(boxes, height, width):
    """ClipBoxes.forward: x1,y1 >= 0; x2 <= W; y2 <= H (no other bounds).  utils.py:134-144."""
    b = np.array(boxes, dtype=F32, copy=True)
    b[:, :, 0] = np.maximum(b[:, :, 0], F32(0))
    b[:, :, 1] = np.maximum(b[:, :, 1], F32(0))
    b[:, :, 2] = np.minimum(b[:, :, 2], F32(width))
    b[:, :, 3] = np.minimum(b[:, :, 3], F32(height))
    return b


# --------------------------------------------------------------------------------------
# a13  NMS  (third party: torchvision 0.26.0 ops/boxes.py batched_nms, csrc/ops/cpu/nms_kernel.cpp)
# torchvision is NOT under /root/reference and the reference pins no version; the oracle of
# record is the installed torchvision 0.26.0+cu128.  Call sites: retinanet/model.py:540,
# IL_method/persuado_label.py:116.
# --------------------------------------------------------------------------------------
def nms(boxes, scores, iou_threshold):
    """Greedy NMS: stable descending score sort; suppress iff inter/((Sa+Sb)-inter) > thr (strict).

    Returns keep indices (int64) into `boxes`, in processing (score-descending) order.
    """
    boxes = np.asarray(boxes, dtype=F32).reshape(-1, 4)
    scores = np.asarray(scores, dtype=F32)
    K = boxes.shape[0]
    if K == 0:
        return np.zeros(0, dtype=np.int64)
    order = np.argsort(-scores.astype(np.float64), kind='stable')
    x1, y1, x2, y2 = (boxes[order, i] for i in range(4))
    areas = (x2 - x1) * (y2 - y1)
    thr = F32(iou_threshold)
    suppressed = np.zeros(K, dtype=bool)
    keep = []
    for i in range(K):
        if suppressed[i]:
            continue
        keep.append(order[i])
        if i + 1 == K:
            break
        xx1 = np.maximum(x1[i], x1[i + 1:])
        yy1 = np.maximum(y1[i], y1[i + 1:])
        xx2 = np.minimum(x2[i], x2[i + 1:])
        yy2 = np.minimum(y2[i], y2[i + 1:])
        w = np.maximum(xx2 - xx1, F32(0))
        h = np.maximum(yy2 - yy1, F32(0))
        inter = w * h
        with np.errstate(divide='ignore', invalid='ignore'):
            ovr = inter / ((areas[i] + areas[i + 1:]) - inter)
        suppressed[i + 1:] |= ovr > thr
    return np.asarray(keep, dtype=np.int64)


def batched_nms(boxes, scores, idxs, iou_threshold, device_rule='cuda'):
    """torchvision.ops.batched_nms: coordinate trick unless numel > 100000 (cuda) / 4000 (cpu).

    Vanilla branch: per-class nms on raw coordinates, result ordered by descending score
    (ties by ascending index here; torchvision's final sort is not stable, so tests avoid
    score ties in that regime).
    """
    boxes = np.asarray(boxes, dtype=F32).reshape(-1, 4)
    scores = np.asarray(scores, dtype=F32)
    idxs = np.asarray(idxs)
    limit = 100_000 if device_rule == 'cuda' else 4000
    if boxes.size == 0:
        return np.zeros(0, dtype=np.int64)
    if boxes.size > limit:
        mask = np.zeros(scores.shape[0], dtype=bool)
        for c in np.unique(idxs):
            cur = np.nonzero(idxs == c)[0]
            mask[cur[nms(boxes[cur], scores[cur], iou_threshold)]] = True
        ki = np.nonzero(mask)[0]
        return ki[np.argsort(-scores[ki].astype(np.float64), kind='stable')]
    max_coordinate = boxes.max()
    offsets = idxs.astype(F32) * (max_coordinate + F32(1))
    return nms(boxes + offsets[:, None], scores, iou_threshold)


# --------------------------------------------------------------------------------------
# a12 / a12'  eval-mode detection output  (retinanet/model.py:494-550; IL_method/persuado_label.py:99-127)
# --------------------------------------------------------------------------------------
def sigmoid(x):
    x = np.asarray(x, dtype=F32)
    return F32(1.0) / (F32(1.0) + np.exp(-x))


def detect(cls, regressions, anchors, height, width, is_logits=True, score_thresh=0.05,
           iou_threshold=0.5, pre_nms_topk=0, device_rule='cuda', image=0):
    """ResNet.predict after self.forward (model.py:507-550) for one image of the batch.

    cls: [N,A,C] logits (is_logits=True, model.py:507 applies Sigmoid) or probabilities
    (Labeler.predict, persuado_label.py:99).  Returns (scores[K'], labels[K'] int64,
    boxes[K',4]) in NMS order, plus the candidate arrays for parity checks.
    pre_nms_topk > 0 keeps only the top-k candidates by (score desc, anchor asc) before NMS
    -- a stage the reference does not have (Q7); 0 disables it.
    """
    cls = np.asarray(cls, dtype=F32)
    prob = sigmoid(cls[image]) if is_logits else cls[image]
    boxes = clip_boxes(bbox_transform(anchors, np.asarray(regressions, dtype=F32)[image:image + 1]), height, width)[0]
    label = np.argmax(prob, axis=1)                       # first max index
    score = prob[np.arange(prob.shape[0]), label]
    m = score > F32(score_thresh)                         # model.py:536 / persuado_label.py:109
    cand_anchor = np.nonzero(m)[0]
    c_scores, c_labels, c_boxes = score[m], label[m].astype(np.int64), boxes[m]
    if pre_nms_topk and c_scores.shape[0] > pre_nms_topk:
        top = np.argsort(-c_scores.astype(np.float64), kind='stable')[:pre_nms_topk]
        top = np.sort(top)                                # keep anchor order among survivors
        cand_anchor, c_scores, c_labels, c_boxes = cand_anchor[top], c_scores[top], c_labels[top], c_boxes[top]
    keep = batched_nms(c_boxes, c_scores, c_labels, iou_threshold, device_rule)
    return dict(scores=c_scores[keep], labels=c_labels[keep], boxes=c_boxes[keep].reshape(-1, 4), keep=keep,
                cand_anchor=cand_anchor, cand_scores=c_scores, cand_labels=c_labels, cand_boxes=c_boxes)


# --------------------------------------------------------------------------------------
# a9  pseudo-label merge + collate format  (retinanet/dataloader.py:129-147, 348-359;
#     IL_method/persuado_label.py:54-91)
# --------------------------------------------------------------------------------------
def merge_pseudo_labels(real_xywh_label, pseudo_xywh_label):
    """Rows = real GT then pseudo GT, xywh -> xyxy in fp64 (dataloader.py:119-142)."""
    rows = [np.asarray(r, dtype=np.float64).reshape(-1, 5) for r in (real_xywh_label, pseudo_xywh_label)]
    ann = np.concatenate(rows, axis=0)
    ann[:, 2] = ann[:, 0] + ann[:, 2]
    ann[:, 3] = ann[:, 1] + ann[:, 3]
    return ann


def collate_annotations(annots):
    """Pad per-image [g_i,5] fp64 rows with -1 to [N,Gmax,5] fp32 (dataloader.py:348-359)."""
    gmax = max((a.shape[0] for a in annots), default=0)
    out = np.full((len(annots), max(gmax, 1), 5), -1.0, dtype=F32)
    for i, a in enumerate(annots):
        if a.shape[0] > 0:
            out[i, :a.shape[0]] = a.astype(F32)
    return out


def filter_pseudo_labels(scores, boxes, labels, real_gt_xyxy, scale, score_thresh=0.7, iou_thresh=0.35):
    """Labeler.get_persuado_label post-filter (persuado_label.py:54-75): keep score > 0.7, divide
    boxes by the resize scale, drop any with max IoU >= 0.35 against a real GT (IoU in fp64)."""
    m = np.asarray(scores) > F32(score_thresh)
    s, b, l = np.asarray(scores)[m], np.asarray(boxes, dtype=F32)[m] / F32(scale), np.asarray(labels)[m]
    if s.shape[0] and np.asarray(real_gt_xyxy).shape[0]:
        bd = b.astype(np.float64)
        gd = np.asarray(real_gt_xyxy, dtype=np.float64)
        area = (gd[:, 2] - gd[:, 0]) * (gd[:, 3] - gd[:, 1])
        iw = np.clip(np.minimum(bd[:, 2:3], gd[:, 2]) - np.maximum(bd[:, 0:1], gd[:, 0]), 0, None)
        ih = np.clip(np.minimum(bd[:, 3:4], gd[:, 3]) - np.maximum(bd[:, 1:2], gd[:, 1]), 0, None)
        ua = np.clip(((bd[:, 2] - bd[:, 0]) * (bd[:, 3] - bd[:, 1]))[:, None] + area - iw * ih, 1e-8, None)
        ok = (iw * ih / ua).max(axis=1) < iou_thresh
        s, b, l = s[ok], b[ok], l[ok]
    return s, b, l


# --------------------------------------------------------------------------------------
# 8(f) row f3  other calc_iou + max users  (IL_method/mas.py:35-67, IL_method/prototype.py:24-47)
# --------------------------------------------------------------------------------------
def output_norm(classifications, regressions, anchors, annotations):
    """MAS Output_norm.forward: mean over images of mean|regression[positive]| (images without positives add 0) and
    sum(cls^2) / (N * C).  Returns (regression_term, classification_term, grad_cls, grad_reg) for unit upstream weights."""
    cls = np.asarray(classifications, dtype=F32)
    reg = np.asarray(regressions, dtype=F32)
    n, a, c = cls.shape
    rterm = 0.0
    greg = np.zeros_like(reg)
    for j in range(n):
        asg = assign(np.asarray(anchors, dtype=F32)[0], annotations[j])
        pos = asg['iou_max'] >= F32(0.5)
        k = int(pos.sum())
        if k > 0:
            rterm += float(np.mean(np.abs(reg[j][pos]).astype(np.float64)))
            greg[j][pos] = np.sign(reg[j][pos]) / F32(4 * k * n)
    cterm = float(np.sum((cls * cls).astype(np.float64)) / (n * c))
    return F32(rterm / n), F32(cterm), (F32(2.0) * cls / F32(n * c)), greg


def get_positive(anchors, annotations, threshold, num_anchors):
    """ProtoTyper._get_positive: positive mask at a custom threshold and the assigned GT class, viewed [N, cells, 9]."""
    pos, tgt = [], []
    for j in range(annotations.shape[0]):
        asg = assign(np.asarray(anchors, dtype=F32)[0], annotations[j])
        pos.append((asg['iou_max'] >= F32(threshold)).reshape(-1, num_anchors))
        tgt.append(asg['label'].astype(np.int64).reshape(-1, num_anchors))
    return np.stack(pos), np.stack(tgt)


# --------------------------------------------------------------------------------------
# 8(f) row f2  head-distillation terms of IL_Loss  (retinanet/losses.py:705-737)
# --------------------------------------------------------------------------------------
def head_distillation(classification, regression, prev_classification, prev_regression, bg_masks, distill_logits=False,
                      ignore_gd=False, g_cls=1.0, g_reg=1.0):
    """Returns (dist_cls_loss, dist_reg_loss, grad_classification[N,A,C], grad_regression[N,A,4])."""
    cls = np.asarray(classification, dtype=F32)
    prev = np.asarray(prev_classification, dtype=F32)
    reg = np.asarray(regression, dtype=F32)
    preg = np.asarray(prev_regression, dtype=F32)
    bgm = np.asarray(bg_masks).astype(bool)
    P = prev.shape[2]
    cur_l = cls[:, :, :P]                                          # :705
    prev_p = sigmoid(prev)
    fg = prev_p > F32(0.05)                                        # :712 / :717
    if distill_logits:
        a, b = prev, cur_l
    else:
        a, b = prev_p, sigmoid(cur_l)                              # :714-715
    reg_mask = bgm & fg.any(axis=2)                                # :720
    d = (preg[reg_mask] - reg[reg_mask]).astype(np.float64)
    z = np.abs(d)
    k_reg = d.size
    reg_loss = np.sum(np.where(z < 1.0, 0.5 * z * z, z - 0.5)) / k_reg if k_reg else np.nan      # SmoothL1Loss, beta 1
    greg = np.zeros_like(reg)
    if k_reg:
        greg[reg_mask] = (-np.clip(d, -1.0, 1.0) * (g_reg / k_reg)).astype(F32)
    sel = np.broadcast_to(reg_mask[:, :, None], fg.shape) if ignore_gd else fg               # :724-727
    e = (a[sel] - b[sel]).astype(np.float64)
    k_cls = e.size
    cls_loss = np.sum(e * e) / k_cls if k_cls else np.nan
    gcls = np.zeros_like(cls)
    if k_cls:
        gsel = -2.0 * e * (g_cls / k_cls)
        if not distill_logits:
            bp = b[sel].astype(np.float64)
            gsel = gsel * (1.0 - bp) * bp
        tmp = np.zeros(fg.shape, dtype=F32)
        tmp[sel] = gsel.astype(F32)
        gcls[:, :, :P] = tmp
    return F32(cls_loss), F32(reg_loss), gcls, greg


def enhance_error(classification, past_class_num, method='L2', g=1.0):
    """`enhance_error` on replay batches (retinanet/losses.py:590-603): over the new-class columns, the probabilities > 0.05
    contribute |p|, p^2 or p^3; loss = sum / max(count, 1).  Returns (loss, grad_classification) for upstream weight g."""
    cls = np.asarray(classification, dtype=F32)
    sel = np.zeros(cls.shape, dtype=bool)
    sel[:, :, past_class_num:] = cls[:, :, past_class_num:] > F32(0.05)          # :591-592
    p = cls[sel].astype(np.float64)
    m = str(method).upper()
    denom = max(p.size, 1)                                                       # :600  max(classification.shape[0], 1)
    if m == 'L1':
        loss, d = np.abs(p).sum() / denom, np.ones_like(p)
    elif m == 'L2':
        loss, d = (p * p).sum() / denom, 2.0 * p
    elif m == 'L3':
        loss, d = (p * p * p).sum() / denom, 3.0 * p * p
    else:
        raise ValueError(method)
    grad = np.zeros_like(cls)
    grad[sel] = (d * (g / denom)).astype(F32)
    return F32(loss), grad


def clip_loss_reduce(bg, fg, clip):
    """IL_Loss's reductions of the per-image terms with clip_loss (losses.py:575-583, 651-659): bg.mean(), and the mean of the
    fg terms >= clip (0 when none).  Returns (cls_bg_loss, cls_fg_loss, w_bg[N], w_fg[N]) -- the weights are d/d(bg_j), d/d(fg_j)."""
    bg = np.asarray(bg, dtype=F32)
    fg = np.asarray(fg, dtype=F32)
    n = bg.shape[0]
    mask = fg >= F32(clip)
    k = int(mask.sum())
    w_fg = np.where(mask, 1.0 / k, 0.0) if k else np.zeros(n)
    fg_term = F32(fg[mask].astype(np.float64).mean()) if k else F32(0)
    return F32(bg.astype(np.float64).mean()), fg_term, np.full(n, 1.0 / n), w_fg


def weight_similarity(classifications, anchors, annotations, threshold=0.5):
    """Weight_similarity.forward (IL_method/weight_init.py:82-115) on image 0: rows of the clamped class probabilities of
    positive anchors whose row sum >= threshold, normalised to sum 1, and the assigned GT labels.  None without GT."""
    cls = np.clip(np.asarray(classifications, dtype=F32)[0], F32(1e-4), F32(1.0 - 1e-4))
    ann = np.asarray(annotations, dtype=F32)[0]
    if not np.any(ann[:, 4] != -1):
        return None
    asg = assign(np.asarray(anchors, dtype=F32)[0], ann)
    rowsum = cls.sum(axis=1, dtype=F32)
    idx = (asg['iou_max'] >= F32(0.5)) & (rowsum >= F32(threshold))
    rows = cls[idx]
    return rows / rows.sum(axis=1, dtype=F32)[:, None], asg['label'][idx].astype(F32)


# --------------------------------------------------------------------------------------
# 8(f) row f4  evaluator post-processing  (evaluator.py:329-361)
# --------------------------------------------------------------------------------------
def coco_results(detections, scales, score_threshold=0.05):
    """detections: list over images of (scores, labels, boxes xyxy).  boxes /= scale (fp32 true division -- CPU ATen),
    boxes[:,2:] -= boxes[:,:2], skip score < threshold.  Returns a list of (image, label, score, [x,y,w,h])."""
    out = []
    for j, (s, l, b) in enumerate(detections):
        b = np.asarray(b, dtype=F32).reshape(-1, 4) / F32(scales[j])
        b = b.copy()
        b[:, 2] -= b[:, 0]
        b[:, 3] -= b[:, 1]
        for i in range(b.shape[0]):
            if F32(s[i]) < F32(score_threshold):
                continue
            out.append((j, int(l[i]), float(s[i]), [float(v) for v in b[i]]))
    return out
