"""GPU parity tests of the eval half of the path (decode/clip, class max + threshold, top-k, per-class NMS) against the CPU
oracle and the golden vectors of the unmodified reference / torchvision.  All calls go through the C ABI.

Bars: labels, candidate sets and NMS keep indices BIT-EXACT; scores bit-exact against ATen's own sigmoid on the same device;
decoded boxes within 2 ulp of the row's largest magnitude when compared with a CPU exp() (bit-exact vs the same-device exp).
"""
import numpy as np
import pytest
import torch

import cl_object_detection_b200 as cld
from cl_object_detection_b200 import detect as D
from oracle import head_oracle as O
from tests.helpers import load

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def cu(x):
    return torch.from_numpy(np.ascontiguousarray(x)).to(DEV)


def boxes_close(got, ref):
    got, ref = np.asarray(got, np.float64), np.asarray(ref, np.float64)
    assert got.shape == ref.shape
    if got.size == 0:
        return
    scale = np.max(np.abs(ref), axis=-1, keepdims=True)
    worst = float((np.abs(got - ref) / np.maximum(scale, 1e-30)).max())
    assert worst <= 2.4e-7, worst


def test_decode_and_clip_vs_golden_and_oracle():
    g = load('decode')
    h, w = int(g['h']), int(g['w'])
    anchors = cld.generate_anchors(h, w, DEV)
    dec = D.BBoxTransform()(anchors, cu(g['reg']))
    boxes_close(dec.cpu().numpy(), g['decoded'])
    boxes_close(dec.cpu().numpy(), O.bbox_transform(O.anchors_for_image(h, w), g['reg']))
    # clip is exact given the same decoded input, and in place like the reference
    t = cu(g['decoded']).clone()
    out = D.ClipBoxes()(t, torch.zeros(2, 3, h, w))
    assert out is t and np.array_equal(t.cpu().numpy(), g['clipped'])
    both = D.decode_boxes(anchors, cu(g['reg']), clip_to=(h, w))
    assert np.array_equal(both.cpu().numpy(), O.clip_boxes(dec.cpu().numpy(), h, w))
    # same-device reference: torch's own exp on the GPU -> bit exact
    a = anchors[0]
    r = cu(g['reg'])
    wd, hg = a[:, 2] - a[:, 0], a[:, 3] - a[:, 1]
    cx, cy = a[:, 0] + 0.5 * wd, a[:, 1] + 0.5 * hg
    pcx, pcy = cx + (r[:, :, 0] * 0.1 + 0) * wd, cy + (r[:, :, 1] * 0.1 + 0) * hg
    pw, ph = torch.exp(r[:, :, 2] * 0.2 + 0) * wd, torch.exp(r[:, :, 3] * 0.2 + 0) * hg
    ref = torch.stack([pcx - 0.5 * pw, pcy - 0.5 * ph, pcx + 0.5 * pw, pcy + 0.5 * ph], dim=2)
    assert torch.equal(dec, ref)


@pytest.mark.parametrize('t', range(5))
def test_nms_golden_torchvision(t):
    g = load('nms')
    boxes, idxs = cu(g[f'boxes{t}']), cu(g[f'idxs{t}'])
    tied, uniq = cu(g[f'scores{t}_tied']), cu(g[f'scores{t}_unique'])
    assert np.array_equal(D.nms(boxes, tied, 0.5).cpu().numpy(), g[f'nms{t}'])
    assert np.array_equal(D.nms(boxes, tied, 0.3).cpu().numpy(), g[f'nms{t}_thr03'])
    assert np.array_equal(D.batched_nms(boxes, tied, idxs, 0.5, mode=D.NMS_MODE_TRICK).cpu().numpy(), g[f'trick{t}'])
    assert np.array_equal(D.batched_nms(boxes, uniq, idxs, 0.5, mode=D.NMS_MODE_TRICK).cpu().numpy(), g[f'trick{t}_unique'])
    assert np.array_equal(D.batched_nms(boxes, uniq, idxs, 0.5, mode=D.NMS_MODE_VANILLA).cpu().numpy(), g[f'vanilla{t}'])
    # torchvision's own switch: numel > limit -> vanilla
    auto = D.batched_nms(boxes, uniq, idxs, 0.5, vanilla_numel_limit=4000).cpu().numpy()
    assert np.array_equal(auto, g[f'vanilla{t}'] if boxes.numel() > 4000 else g[f'trick{t}_unique'])


@pytest.mark.parametrize('K,ncls,seed', [(0, 1, 0), (1, 1, 1), (63, 3, 2), (64, 3, 3), (65, 3, 4), (1000, 20, 5), (4097, 80, 6),
                                          (1023, 1, 7), (1024, 7, 8), (1025, 7, 9), (2048, 40, 10), (2049, 40, 11)])
def test_nms_random_vs_oracle(K, ncls, seed):
    rng = np.random.default_rng(seed)
    x1, y1 = rng.uniform(0, 300, K), rng.uniform(0, 300, K)
    boxes = np.stack([x1, y1, x1 + rng.uniform(2, 120, K), y1 + rng.uniform(2, 120, K)], 1).astype(np.float32).reshape(-1, 4)
    scores = (np.round(rng.uniform(0.05, 1, K) * 50) / 50).astype(np.float32)      # heavy score ties
    idxs = rng.integers(0, ncls, K)
    for mode, rule in ((D.NMS_MODE_TRICK, None), (D.NMS_MODE_VANILLA, None)):
        got = D.batched_nms(cu(boxes), cu(scores), cu(idxs), 0.5, mode=mode).cpu().numpy()
        if mode == D.NMS_MODE_TRICK:
            ref = O.batched_nms(boxes, scores, idxs, 0.5, 'cuda') if boxes.size <= 100000 else None
        else:
            ref = O.batched_nms(boxes, scores, idxs, 0.5, 'cpu') if boxes.size > 4000 else None
        if ref is not None:
            assert np.array_equal(got, ref)
    if K:
        assert np.array_equal(D.nms(cu(boxes), cu(scores), 0.5).cpu().numpy(), O.nms(boxes, scores, 0.5))


def same_device_rule(name):
    return dict(trick=4000, vanilla=4000, none=4000)[name]


@pytest.mark.parametrize('name', ['trick', 'vanilla', 'none'])
def test_predict_golden_reference(name):
    """ResNet.predict / Labeler.predict outputs recorded from the unmodified reference (CPU run: torchvision switches to the
    vanilla branch above 4000 coordinates, so the same limit is passed here)."""
    g = load('predict_' + name)
    h, w = int(g['h']), int(g['w'])
    anchors = cld.generate_anchors(h, w, DEV)
    img = torch.zeros(1, 3, h, w, device=DEV)
    s, b, l = D.labeler_predict(img, cu(g['probs']), cu(g['reg']), anchors, vanilla_numel_limit=4000)
    assert np.array_equal(s.cpu().numpy(), g['scores'])
    assert np.array_equal(l.cpu().numpy(), g['labels']) and l.dtype == torch.int64
    boxes_close(b.cpu().numpy(), g['boxes'])
    # logits entry (ResNet.predict): sigmoid on the device; ATen's sigmoid on the same device must agree bit for bit
    s2, l2, b2 = D.predict_from_head(cu(g['logits']), cu(g['reg']), anchors, img, vanilla_numel_limit=4000)
    probs_dev = torch.sigmoid(cu(g['logits']))
    ref = O.detect(probs_dev.cpu().numpy(), g['reg'], O.anchors_for_image(h, w), h, w, is_logits=False, device_rule='cpu')
    assert np.array_equal(s2.cpu().numpy(), ref['scores'])
    assert np.array_equal(l2.cpu().numpy(), ref['labels'])
    boxes_close(b2.cpu().numpy(), ref['boxes'])
    with pytest.raises(ValueError):
        D.predict_from_head(cu(g['logits']), cu(g['reg']), anchors, img, thresh=[0.05])


def test_predict_method_dropin_and_empty():
    h, w, C = 64, 96, 6
    anchors = cld.generate_anchors(h, w, DEV)
    A = anchors.shape[1]
    gen = torch.Generator(device=DEV).manual_seed(1)
    logits = torch.randn(1, A, C, device=DEV, generator=gen) - 12.0      # nothing passes 0.05
    reg = torch.randn(1, A, 4, device=DEV, generator=gen) * 0.3

    class Model:
        def forward(self, img_batch, return_feat=False, return_anchor=True, enable_act=False):
            return logits, reg, anchors
    img = torch.zeros(1, 3, h, w, device=DEV)
    s, l, b = D.predict(Model(), img)
    assert s.shape == (0,) and l.shape == (0,) and l.dtype == torch.int64 and b.shape == (0, 4)
    out = D.labeler_predict(img, torch.sigmoid(logits), reg, anchors)
    assert all(t.numel() == 0 for t in out)


@pytest.mark.parametrize('C,topk,mu', [(80, 1000, -4.0), (80, 0, -7.5), (20, 300, -5.0), (7, 50, -3.0), (5, 0, -6.0)])
def test_detect_batch_vs_oracle(C, topk, mu):
    """Seeded batches incl. class counts that are not multiples of 4, top-k on/off, heavy exact score ties from saturated
    logits; per-image results equal to the oracle's on the same-device sigmoid."""
    h, w, N = 256, 320, 3
    anchors = cld.generate_anchors(h, w, DEV)
    A = anchors.shape[1]
    gen = torch.Generator(device=DEV).manual_seed(C * 13 + topk)
    logits = torch.randn(N, A, C, device=DEV, generator=gen) * 2 + mu
    logits[0, :40] = 25.0                     # saturated rows: every class ties at 1.0 -> label 0, anchor-order ties
    logits[1, 5, :] = -2.0                    # a full row of equal logits above the threshold
    reg = torch.randn(N, A, 4, device=DEV, generator=gen) * 0.5
    got = D.detect_batch(logits, reg, anchors, h, w, pre_nms_topk=topk)
    probs = torch.sigmoid(logits).cpu().numpy()
    oa = O.anchors_for_image(h, w)
    for j in range(N):
        ref = O.detect(probs, reg.cpu().numpy(), oa, h, w, is_logits=False, pre_nms_topk=topk, image=j)
        s, l, b = got[j]
        assert np.array_equal(s.cpu().numpy(), ref['scores']), (j, s.shape, ref['scores'].shape)
        assert np.array_equal(l.cpu().numpy(), ref['labels'])
        boxes_close(b.cpu().numpy(), ref['boxes'])
    # padded form agrees with the list form
    ps, pl, pb, pc = D.detect_batch(logits, reg, anchors, h, w, pre_nms_topk=topk, return_padded=True)
    for j in range(N):
        k = int(pc[j])
        assert k == got[j][0].shape[0] and torch.equal(ps[j, :k], got[j][0]) and torch.equal(pb[j, :k], got[j][2])


def test_full_size_coco_shape_topk1000():
    """BASELINE config 4 shape (800x1333, C=80, top-1000, per-class NMS), N=2: oracle comparison + NMS invariants."""
    h, w, C, N, topk = 800, 1333, 80, 2, 1000
    anchors = cld.generate_anchors(h, w, DEV)
    A = anchors.shape[1]
    gen = torch.Generator(device=DEV).manual_seed(4)
    logits = torch.randn(N, A, C, device=DEV, generator=gen) * 2 - 4
    reg = torch.randn(N, A, 4, device=DEV, generator=gen) * 0.5
    got = D.detect_batch(logits, reg, anchors, h, w, pre_nms_topk=topk)
    probs = torch.sigmoid(logits[:1]).cpu().numpy()
    ref = O.detect(probs, reg[:1].cpu().numpy(), O.anchors_for_image(h, w), h, w, is_logits=False, pre_nms_topk=topk)
    s, l, b = got[0]
    assert np.array_equal(s.cpu().numpy(), ref['scores']) and np.array_equal(l.cpu().numpy(), ref['labels'])
    boxes_close(b.cpu().numpy(), ref['boxes'])
    for s, l, b in got:
        assert 0 < s.shape[0] <= topk
        assert torch.all(s[:-1] >= s[1:])                        # score-descending
        # idempotence: NMS of the output keeps everything
        again = D.batched_nms(b, s, l, 0.5)
        assert again.shape[0] == s.shape[0]


@pytest.mark.parametrize('mu,reg_sigma', [(-10.5, 0.5), (-9.5, 0.3), (-8.5, 0.3)])
def test_full_size_vs_torch_eager_and_torchvision_on_the_same_device(mu, reg_sigma):
    """BASELINE config 4 shape in the reference's own mode (no top-k) against the torch-eager restatement of
    ResNet.predict + torchvision.ops.batched_nms run on the SAME GPU: ~1.3 k, ~8 k and ~40 k candidates per image (the
    last one crosses torchvision's 100 000-element switch to per-class NMS).  Same-device sigmoid and exp: everything
    bit-exact, including the order."""
    from oracle import torch_eager as E
    h, w, C, N = 800, 1333, 80, 3
    anchors = cld.generate_anchors(h, w, DEV)
    A = anchors.shape[1]
    gen = torch.Generator(device=DEV).manual_seed(int(-mu * 10))
    logits = torch.randn(N, A, C, device=DEV, generator=gen) * 2 + mu
    reg = torch.randn(N, A, 4, device=DEV, generator=gen) * reg_sigma
    got = D.detect_batch(logits, reg, anchors, h, w)
    for j in range(N):
        s, l, b = E.predict(logits[j:j + 1], reg[j:j + 1], anchors, h, w)
        assert s.shape[0] > 100
        assert torch.equal(got[j][0], s) and torch.equal(got[j][1], l)
        assert torch.equal(got[j][2], b)


def _conv_levels(x, h, w, per_anchor):
    """[N, A, per_anchor] in the reference's concatenated order -> five [N, 9*per_anchor, H_l, W_l] conv-layout tensors
    (the inverse of ClassificationModel / RegressionModel's permute + view and ResNet.forward's cat)."""
    out, off, n = [], 0, x.shape[0]
    for l in range(3, 8):
        hl, wl = (h + 2 ** l - 1) // 2 ** l, (w + 2 ** l - 1) // 2 ** l
        cnt = hl * wl * 9
        out.append(x[:, off:off + cnt].reshape(n, hl, wl, 9 * per_anchor).permute(0, 3, 1, 2).contiguous())
        off += cnt
    assert off == x.shape[1]
    return out


@pytest.mark.parametrize('h,w,C,N,mu,topk,logits', [(800, 1333, 80, 3, -10.5, 0, True), (800, 1333, 80, 2, -4.0, 1000, True),
                                                    (512, 512, 20, 4, -6.0, 0, True), (33, 70, 7, 2, -3.0, 50, True),
                                                    (608, 1024, 16, 2, -5.0, 300, False), (1, 1, 3, 2, 0.0, 0, True)])
def test_conv_layout_filter_equals_concatenated_path(h, w, C, N, mu, topk, logits):
    """SURVEY 8f row f1, eval side: detect_batch_head on the head's raw conv outputs vs detect_batch on the reference's
    reshaped + concatenated tensors -- scores, labels, boxes and order bit for bit (quantised logits give exact ties in the
    class maximum, saturated logits exercise the equal-probability rule)."""
    anchors = cld.generate_anchors(h, w, DEV)
    A = anchors.shape[1]
    gen = torch.Generator(device=DEV).manual_seed(h + C)
    x = torch.randn(N, A, C, device=DEV, generator=gen) * 2 + mu
    x[0] = torch.round(x[0] * 4) / 4                              # exact ties between classes
    if N > 1:
        x[1, ::7] += 16.0                                         # saturated sigmoid: several classes with p == 1.0f
    if not logits:
        x = torch.sigmoid(x)
    reg = torch.randn(N, A, 4, device=DEV, generator=gen) * 0.4
    ref = D.detect_batch(x, reg, anchors, h, w, is_logits=logits, pre_nms_topk=topk)
    got = D.detect_batch_head(_conv_levels(x, h, w, C), _conv_levels(reg, h, w, 4), anchors, h, w, is_logits=logits,
                              pre_nms_topk=topk)
    for (s0, l0, b0), (s1, l1, b1) in zip(ref, got):
        assert torch.equal(s0, s1) and torch.equal(l0, l1) and torch.equal(b0, b1)
    assert sum(int(s.shape[0]) for s, _, _ in ref) > 0 or A < 100


def test_f4_coco_results_vs_oracle_and_torch_cpu():
    """evaluator post-processing (evaluator.py:329-361) on the device, bit-exact with the CPU statements."""
    h, w, C, N = 256, 320, 7, 3
    anchors = cld.generate_anchors(h, w, DEV)
    A = anchors.shape[1]
    gen = torch.Generator(device=DEV).manual_seed(21)
    logits = torch.randn(N, A, C, device=DEV, generator=gen) * 2 - 5
    logits[2] = -20.0                                           # an image without detections
    reg = torch.randn(N, A, 4, device=DEV, generator=gen) * 0.4
    scales = [1.3371234567, 0.731, 2.0]
    padded = D.detect_batch(logits, reg, anchors, h, w, pre_nms_topk=200, return_padded=True)
    got = D.coco_results(padded, scales, image_ids=[101, 102, 103], label_to_coco_label=lambda x: x + 1, score_threshold=0.3)
    dets = D.detect_batch(logits, reg, anchors, h, w, pre_nms_topk=200)
    ref = O.coco_results([(s.cpu().numpy(), l.cpu().numpy(), b.cpu().numpy()) for s, l, b in dets], scales, 0.3)
    assert len(got) == len(ref) and len(got) > 0
    for g_, r_ in zip(got, ref):
        assert g_['image_id'] == 101 + r_[0] and g_['category_id'] == r_[1] + 1
        assert g_['score'] == r_[2] and g_['bbox'] == r_[3]
    # and against the evaluator's own torch-CPU statements
    k = 0
    for j, (s, l, b) in enumerate(dets):
        b = b.cpu().clone()
        b /= scales[j]
        if b.shape[0] > 0:
            b[:, 2] -= b[:, 0]
            b[:, 3] -= b[:, 1]
        for i in range(b.shape[0]):
            if float(s[i]) < 0.3:
                continue
            assert got[k]['bbox'] == b[i].tolist() and got[k]['score'] == float(s[i])
            k += 1
    assert k == len(got)


def test_a9_pseudo_label_filter_golden():
    """Tail of Labeler.get_persuado_label (persuado_label.py:52-81) vs the reference's own torch statements (golden)."""
    g = load('a9_pseudo_labels')
    s, b, l = cld.filter_pseudo_labels(cu(g['pl_scores']), cu(g['pl_boxes']), cu(g['pl_labels']), cu(g['pl_gt']), float(g['pl_scale']))
    assert np.array_equal(s.cpu().numpy(), g['out_scores']) and np.array_equal(l.cpu().numpy(), g['out_labels'])
    # `boxes / scale` runs on the device like in the reference (persuado_label.py:55 divides CUDA tensors): ATen's CUDA kernel
    # multiplies by the reciprocal of a scalar divisor, the CPU kernel that produced the golden vectors divides: <= 1 ulp.
    assert np.allclose(b.cpu().numpy(), g['out_boxes'], rtol=2.4e-7, atol=6.2e-5)   # w = x2 - x1: 2 ulp of the coordinate (<= 512)
    e = cld.filter_pseudo_labels(torch.empty(0, device=DEV), torch.empty((0, 4), device=DEV), torch.empty(0, dtype=torch.int64, device=DEV),
                                 cu(g['pl_gt']), 1.0)
    assert e[0].numel() == 0 and e[1].shape == (0, 4)


def _sort_through_abi(scores, anchors_idx, counts, topk=0):
    """cldet_sort_candidates on hand-made candidate lists: scores [N, cap] float32, anchors_idx [N, cap] int32 (distinct per
    image), counts [N].  Returns (sorted anchor ids [N, cap], sorted score bits, sorted_counts)."""
    from cl_object_detection_b200 import _lib
    lib = _lib.load()
    n, cap = scores.shape
    cand = torch.zeros((n, cap, 8), dtype=torch.float32, device=DEV)
    cand[:, :, 4] = scores
    cand.view(torch.int32)[:, :, 6] = anchors_idx
    bits = scores.view(torch.int32).to(torch.int64) & 0xFFFFFFFF
    ordered = torch.where(bits >= 0x80000000, (~bits) & 0xFFFFFFFF, bits | 0x80000000)
    keys = (ordered << 32) | ((0xFFFFFFFF - anchors_idx.to(torch.int64)) & 0xFFFFFFFF)
    keys = keys.contiguous()
    cnt = torch.as_tensor(counts, dtype=torch.int32, device=DEV)
    max_count = int(max(counts))
    out_cap = min(topk, cap) if topk else max_count
    sorted_c = torch.zeros((n, out_cap, 8), dtype=torch.float32, device=DEV)
    sorted_counts = torch.empty(n, dtype=torch.int32, device=DEV)
    ws_bytes = lib.cldet_sort_workspace_bytes(n, max_count, topk)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=DEV)
    _lib.check(lib.cldet_sort_candidates(cand.data_ptr(), keys.data_ptr(), cnt.data_ptr(), n, cap, max_count, topk,
                                         sorted_c.data_ptr(), out_cap, sorted_counts.data_ptr(), ws.data_ptr(), ws_bytes,
                                         torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    return sorted_c.view(torch.int32)[:, :, 6], sorted_c[:, :, 4], sorted_counts, keys


@pytest.mark.parametrize('counts', [[2049, 4096, 4097], [100000, 1, 0], [12345, 8192, 2048], [40000]])
@pytest.mark.parametrize('kind', ['random', 'all_equal', 'few_values', 'negative_and_zero'])
def test_radix_sort_orders_like_a_stable_descending_sort(counts, kind):
    """The long-list ordering (segmented LSD radix sort, one CTA per image) against torch.sort of the 64-bit keys: score
    descending, equal scores by ascending anchor -- for ragged counts around the tile size, a list where EVERY score is equal
    (the case that made the pairwise rank sort quadratic), heavy ties, and scores <= 0 (the key transform's other branch)."""
    n, cap = len(counts), max(max(counts), 1)
    gen = torch.Generator(device=DEV).manual_seed(sum(counts) + len(kind))
    if kind == 'random':
        scores = torch.rand(n, cap, device=DEV, generator=gen)
    elif kind == 'all_equal':
        scores = torch.full((n, cap), 0.75, device=DEV)
    elif kind == 'few_values':
        scores = torch.randint(0, 7, (n, cap), device=DEV, generator=gen).float() / 8 + 0.06
    else:
        scores = torch.randn(n, cap, device=DEV, generator=gen)
        scores[:, ::5] = 0.0
        scores[:, 1::11] = -0.0
    # distinct anchor ids per image, in scrambled order (the filter appends candidates in arbitrary order)
    ids = torch.stack([torch.randperm(cap + 1000, device=DEV, generator=gen)[:cap] for _ in range(n)]).to(torch.int32)
    got_ids, got_scores, got_counts, keys = _sort_through_abi(scores, ids, counts)
    assert got_counts.tolist() == counts
    for j, c in enumerate(counts):
        if c == 0:
            continue
        # reference order: descending unsigned 64-bit key (flip the top bit to compare as signed int64)
        order = torch.argsort(keys[j, :c] ^ (-0x8000000000000000), descending=True, stable=True)
        assert torch.equal(got_ids[j, :c], ids[j, :c][order]), (kind, j, c)
        assert torch.equal(got_scores[j, :c].view(torch.int32), scores[j, :c][order].view(torch.int32))


@pytest.mark.parametrize('counts', [[1, 2, 3, 1000], [2047, 2048, 1025, 0], [1500, 3000, 5000, 64]])
@pytest.mark.parametrize('kind', ['random', 'all_equal', 'few_values'])
@pytest.mark.parametrize('topk', [0, 1000])
def test_short_list_sort_and_topk(counts, kind, topk):
    """Lists of <= 2048 candidates are ranked by one block through a bucket pass, longer ones by the pairwise rank / radix sort, and with a
    top-k by radix select first (ties at the k-th score survive the select, so an all-equal list reaches the ordering pass whole):
    every image of a ragged batch must come out in the order of a stable descending sort of the keys, cut at top-k."""
    n, cap = len(counts), max(max(counts), 1)
    gen = torch.Generator(device=DEV).manual_seed(sum(counts) + len(kind) + topk)
    if kind == 'random':
        scores = torch.rand(n, cap, device=DEV, generator=gen)
    elif kind == 'all_equal':
        scores = torch.full((n, cap), 0.75, device=DEV)
    else:
        scores = torch.randint(0, 7, (n, cap), device=DEV, generator=gen).float() / 8 + 0.06
    ids = torch.stack([torch.randperm(cap + 1000, device=DEV, generator=gen)[:cap] for _ in range(n)]).to(torch.int32)
    got_ids, got_scores, got_counts, keys = _sort_through_abi(scores, ids, counts, topk=topk)
    want_counts = [min(c, topk) if topk else c for c in counts]
    assert got_counts.tolist() == want_counts
    for j, c in enumerate(counts):
        if c == 0:
            continue
        order = torch.argsort(keys[j, :c] ^ (-0x8000000000000000), descending=True, stable=True)[:want_counts[j]]
        assert torch.equal(got_ids[j, :want_counts[j]], ids[j, :c][order]), (kind, j, c)
        assert torch.equal(got_scores[j, :want_counts[j]].view(torch.int32), scores[j, :c][order].view(torch.int32))


@pytest.mark.parametrize('k', [500, 1024, 3000])
@pytest.mark.parametrize('case', ['negative', 'huge', 'many_labels', 'plain'])
def test_coordinate_trick_corner_cases_vs_torchvision_same_device(k, case):
    """torchvision's coordinate trick adds label * (max + 1) in fp32.  With ordinary boxes that keeps classes apart and the mask
    kernels skip cross-label pairs on the label compare; with negative coordinates, huge coordinates (the offsets round and
    classes DO interact) or huge label values that proof fails and the full test must run.  Either way the keep list has to equal
    torchvision's own kernel on the same GPU -- for the one-launch short-list path (k <= 1024) and the three-launch path."""
    import torchvision
    gen = torch.Generator(device=DEV).manual_seed(k + len(case))
    centers = torch.rand(k // 15 + 1, 2, device=DEV, generator=gen) * 500
    which = torch.randint(0, centers.shape[0], (k,), device=DEV, generator=gen)
    xy = centers[which] + torch.randn(k, 2, device=DEV, generator=gen) * 5
    wh = torch.rand(k, 2, device=DEV, generator=gen) * 50 + 8
    idxs = torch.randint(0, 6, (k,), device=DEV, generator=gen)
    if case == 'negative':
        xy = xy - 250.0
    elif case == 'huge':
        xy = xy + 3.0e7          # fp32 spacing 2 at this magnitude: offsets and boxes round, classes merge
        wh = wh * 4
    elif case == 'many_labels':
        idxs = idxs * 100000     # (label_max + 1) * (max + 1) far beyond 2^21
    boxes = torch.cat([xy, xy + wh], dim=1).contiguous()
    scores = ((torch.randperm(k, device=DEV, generator=gen).float() + 1) / (k + 1)).contiguous()      # distinct
    ref = torchvision.ops.boxes._batched_nms_coordinate_trick(boxes, scores, idxs, 0.5)
    got = D.batched_nms(boxes, scores, idxs, 0.5, mode=D.NMS_MODE_TRICK)
    assert torch.equal(got, ref), (got.shape, ref.shape)
    assert 0 < got.shape[0] < k


@pytest.mark.parametrize('k,trick', [(3000, True), (10000, True), (10000, False), (30000, False)])
def test_batched_nms_long_lists_vs_torchvision_same_device(k, trick):
    """cldet batched_nms (radix sort + streamed resolve for K > 2048 / > 1216) vs torchvision.ops on the same GPU: clustered
    boxes with duplicated scores, both torchvision branches (coordinate trick / per-class)."""
    import torchvision
    gen = torch.Generator(device=DEV).manual_seed(k + int(trick))
    centers = torch.rand(k // 20 + 1, 2, device=DEV, generator=gen) * 600
    which = torch.randint(0, centers.shape[0], (k,), device=DEV, generator=gen)
    xy = centers[which] + torch.randn(k, 2, device=DEV, generator=gen) * 6
    wh = torch.rand(k, 2, device=DEV, generator=gen) * 60 + 8
    boxes = torch.cat([xy, xy + wh], dim=1).contiguous()
    scores = (torch.randint(0, 4000, (k,), device=DEV, generator=gen).float() / 4000).contiguous()      # many exact ties
    idxs = torch.randint(0, 5, (k,), device=DEV, generator=gen)
    if trick:
        ref = torchvision.ops.boxes._batched_nms_coordinate_trick(boxes, scores, idxs, 0.5)
        got = D.batched_nms(boxes, scores, idxs, 0.5, mode=D.NMS_MODE_TRICK)
    else:
        ref = torchvision.ops.boxes._batched_nms_vanilla(boxes, scores, idxs, 0.5)
        got = D.batched_nms(boxes, scores, idxs, 0.5, mode=D.NMS_MODE_VANILLA)
    assert got.shape[0] > 10
    assert torch.all(scores[got][:-1] >= scores[got][1:])
    # exact agreement with the oracle semantics (stable by index) on a distinct-score copy of the problem
    sd = ((torch.randperm(k, device=DEV, generator=gen).float() + 1) / (k + 1)).contiguous()      # distinct by construction
    if trick:
        ref_d = torchvision.ops.boxes._batched_nms_coordinate_trick(boxes, sd, idxs, 0.5)
        got_d = D.batched_nms(boxes, sd, idxs, 0.5, mode=D.NMS_MODE_TRICK)
    else:
        ref_d = torchvision.ops.boxes._batched_nms_vanilla(boxes, sd, idxs, 0.5)
        got_d = D.batched_nms(boxes, sd, idxs, 0.5, mode=D.NMS_MODE_VANILLA)
    assert torch.equal(got_d, ref_d), (got_d.shape, ref_d.shape)
    # with ties: the CPU oracle defines the order (stable sort), the kept set must match it exactly
    want = O.batched_nms(boxes.cpu().numpy(), scores.cpu().numpy(), idxs.cpu().numpy(), 0.5, 'cuda' if trick else 'cpu')     # rule picks the branch
    assert np.array_equal(got.cpu().numpy(), want)
    assert ref.shape[0] > 0


def test_nms_reports_counts_larger_than_the_workspace():
    """A caller whose counts exceed the max_count the workspace was sized for gets an error marker (keep count -1), never
    silently truncated detections."""
    from cl_object_detection_b200 import _lib
    lib = _lib.load()
    n, cap = 2, 512
    sorted_c = torch.zeros((n, cap, 8), dtype=torch.float32, device=DEV)
    sorted_c[:, :, 2:4] = 1.0
    counts = torch.tensor([100, 400], dtype=torch.int32, device=DEV)
    for max_count in (128, 2000):      # shared-memory resolve and (via a larger workspace than the list) the normal case
        ws_bytes = lib.cldet_nms_workspace_bytes(n, max_count)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=DEV)
        keep = torch.empty((n, cap), dtype=torch.int32, device=DEV)
        kc = torch.empty(n, dtype=torch.int32, device=DEV)
        _lib.check(lib.cldet_nms_sorted(sorted_c.data_ptr(), counts.data_ptr(), n, cap, max_count, 0.5, 1, 100000, keep.data_ptr(),
                                        kc.data_ptr(), ws.data_ptr(), ws_bytes, torch.cuda.current_stream().cuda_stream))
        got = kc.tolist()
        if max_count == 128:
            assert got[0] == 1 and got[1] == -1, got          # 100 identical boxes -> 1 kept; 400 > 128 -> error marker
        else:
            assert got == [1, 1], got
