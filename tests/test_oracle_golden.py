"""Pin the CPU oracle (oracle/head_oracle.py) to the golden vectors produced by the unmodified
reference (tests/golden/make_golden.py).  CPU only; runs in the `-m "not gpu"` suite."""
import hashlib

import numpy as np
import pytest

from oracle import head_oracle as O
from tests.helpers import FOCAL_CASES, golden_params, load, rel_err

SHAPES = [(64, 96), (128, 160), (33, 70), (512, 512), (800, 1333), (1333, 1333), (608, 1024)]


@pytest.mark.parametrize('h,w', SHAPES)
def test_anchors_bit_exact(h, w):
    g = load('anchors')
    a = O.anchors_for_image(h, w)
    assert tuple(g[f'{h}x{w}_shape']) == a.shape
    assert a.shape[1] == O.num_anchors(h, w)
    assert hashlib.sha256(a.tobytes()).hexdigest() == str(g[f'{h}x{w}_sha256'])
    if f'{h}x{w}' in g:
        assert np.array_equal(a, g[f'{h}x{w}'])
    else:
        assert np.array_equal(a[0, g[f'{h}x{w}_idx']], g[f'{h}x{w}_rows'])


def test_anchor_counts_match_survey():
    assert O.num_anchors(512, 512) == 49104
    assert O.num_anchors(800, 1333) == 200700
    assert O.num_anchors(1333, 1333) == 335439


def test_calc_iou_bit_exact():
    g = load('calc_iou')
    assert np.array_equal(O.calc_iou(g['a'], g['b']), g['iou'])


@pytest.mark.parametrize('case', FOCAL_CASES)
def test_focal_loss_matches_reference(case):
    g = load('focal_' + case)
    params = golden_params(g)
    anchors = O.anchors_for_image(int(g['h']), int(g['w']))
    out = O.focal_loss(g['cls'], g['reg'], anchors, g['ann'], int(g['cur_state']), params, float(g['progress']),
                       w_bg=g['wb'], w_fg=g['wf'], w_reg=float(g['wr']), w_enh=float(g['we']))
    # assignments: bit exact
    for j, asg in enumerate(out['assign']):
        if asg['valid'] == 0:
            assert np.all(g['state'][j] == 3)
            continue
        assert np.array_equal(asg['state'], g['state'][j])
        assert np.array_equal(asg['argmax'], g['argmax'][j])
    # losses and gradients: 1e-5 relative (north_star tolerance)
    assert rel_err(out['bg'], g['bg'], 1e-30) < 1e-5
    assert rel_err(out['fg'], g['fg'], 1e-30) < 1e-5
    assert rel_err(out['reg_loss'], g['reg_loss'], 1e-30) < 1e-5
    gc, gr = g['grad_cls'], g['grad_reg']
    assert np.array_equal(out['grad_cls'] == 0, gc == 0), 'zero-gradient pattern (ignore / out-of-band) differs'
    assert np.max(np.abs(out['grad_cls'] - gc) / (np.abs(gc) + 1e-12 * np.abs(gc).max())) < 1e-5
    # smooth-L1's quadratic zone has gradient 9*(t - r): the subtraction cancels, so a 1-ulp difference in log()
    # shows up amplified on SMALL gradients; bound those against the gradient scale instead (1e-5 of max).
    assert float((np.abs(out['grad_reg'] - gr) - 1e-5 * np.abs(gr)).max()) <= 1e-5 * float(np.abs(gr).max())
    if 'bg_masks' in g:
        assert np.array_equal(out['bg_masks'], g['bg_masks'])
    if 'enhance_on_new_loss' in g:
        assert rel_err(out['enhance_on_new_loss'], g['enhance_on_new_loss'], 1e-30) < 1e-5


def test_progress_argument_is_a_noop():
    g = load('focal_il_default_pseudo')
    params = golden_params(g)
    anchors = O.anchors_for_image(int(g['h']), int(g['w']))
    a = O.focal_loss(g['cls'], g['reg'], anchors, g['ann'], 1, params, 0.5)
    b = O.focal_loss(g['cls'], g['reg'], anchors, g['ann'], 1, params, -1)
    assert np.array_equal(a['bg'], b['bg']) and np.array_equal(a['grad_cls'], b['grad_cls'])


def test_decode_and_clip():
    g = load('decode')
    anchors = O.anchors_for_image(int(g['h']), int(g['w']))
    dec = O.bbox_transform(anchors, g['reg'])
    # exp() differs by <= 1 ulp between numpy and torch CPU; everything else is exact.  x1 = cx - 0.5*w cancels,
    # so the bound is 2 ulp of the LARGEST magnitude in the row (the width/centre), not of the coordinate itself.
    scale = np.max(np.abs(g['decoded']), axis=2, keepdims=True)
    assert np.all(np.abs(dec - g['decoded']) <= 2.4e-7 * scale)
    clipped = O.clip_boxes(g['decoded'], int(g['h']), int(g['w']))
    assert np.array_equal(clipped, g['clipped'])


@pytest.mark.parametrize('t', range(5))
def test_nms_matches_torchvision(t):
    g = load('nms')
    boxes, idxs = g[f'boxes{t}'], g[f'idxs{t}']
    assert np.array_equal(O.nms(boxes, g[f'scores{t}_tied'], 0.5), g[f'nms{t}'])
    assert np.array_equal(O.nms(boxes, g[f'scores{t}_tied'], 0.3), g[f'nms{t}_thr03'])
    assert np.array_equal(O.batched_nms(boxes, g[f'scores{t}_tied'], idxs, 0.5, 'cuda'), g[f'trick{t}'])
    assert np.array_equal(O.batched_nms(boxes, g[f'scores{t}_unique'], idxs, 0.5, 'cuda'), g[f'trick{t}_unique'])
    # force the vanilla branch through the cpu rule when K*4 > 4000, else call it via a tiny limit
    if boxes.size > 4000:
        assert np.array_equal(O.batched_nms(boxes, g[f'scores{t}_unique'], idxs, 0.5, 'cpu'), g[f'vanilla{t}'])


@pytest.mark.parametrize('name', ['trick', 'vanilla', 'none'])
def test_detect_matches_reference_predict(name):
    g = load('predict_' + name)
    h, w = int(g['h']), int(g['w'])
    anchors = O.anchors_for_image(h, w)
    # feed the reference's own probabilities so the comparison does not hinge on a 1-ulp sigmoid/exp difference
    out = O.detect(g['probs'], g['reg'], anchors, h, w, is_logits=False, device_rule=str(g['device_rule']))
    assert out['scores'].shape == g['scores'].shape
    assert np.array_equal(out['labels'], g['labels'])
    assert np.array_equal(out['scores'], g['scores'])
    assert np.max(np.abs(out['boxes'] - g['boxes'])) <= 1e-4   # exp() ulp only
    lg = O.detect(g['logits'], g['reg'], anchors, h, w, is_logits=True, device_rule=str(g['device_rule']))
    assert abs(lg['scores'].shape[0] - g['scores'].shape[0]) <= 2


def test_collate_and_merge_format():
    real = np.array([[10, 20, 30, 40, 4]], dtype=np.float64)
    pseudo = np.array([[5, 6, 7, 8, 1], [1, 2, 3, 4, 0]], dtype=np.float64)
    m = O.merge_pseudo_labels(real, pseudo)
    assert m.shape == (3, 5) and np.array_equal(m[0], [10, 20, 40, 60, 4]) and np.array_equal(m[2], [1, 2, 4, 6, 0])
    out = O.collate_annotations([m, np.zeros((0, 5))])
    assert out.shape == (2, 3, 5) and out.dtype == np.float32 and np.all(out[1] == -1)
    assert O.collate_annotations([np.zeros((0, 5))]).shape == (1, 1, 5)


def test_f3_iou_users_match_reference():
    g = load('f3_iou_users')
    anchors = O.anchors_for_image(int(g['h']), int(g['w']))
    r, c, gc, gr = O.output_norm(g['cls'], g['reg'], anchors, g['ann'])
    assert abs(float(r) - float(g['norm_regression'])) <= 1e-6 * abs(float(g['norm_regression']))
    assert abs(float(c) - float(g['norm_classification'])) <= 1e-6 * abs(float(g['norm_classification']))
    assert np.allclose(0.3 * gc, g['grad_cls'], rtol=1e-6, atol=0)
    assert np.allclose(0.7 * gr, g['grad_reg'], rtol=1e-6, atol=0)
    pos, tgt = O.get_positive(anchors, g['ann'], float(g['proto_threshold']), 9)
    assert np.array_equal(pos, g['proto_positive']) and np.array_equal(tgt, g['proto_targets'])


@pytest.mark.parametrize('name,dl,ig', [('probs', False, False), ('logits', True, False), ('probs_ignoregd', False, True),
                                        ('logits_ignoregd', True, True)])
def test_f2_head_distillation_matches_torch_restatement(name, dl, ig):
    g = load('f2_distill')
    lc, lr, gc, gr = O.head_distillation(g['cls'], g['reg'], g['prev'], g['preg'], g['bg'], dl, ig, g_cls=0.6, g_reg=1.7)
    assert rel_err(lc, g[name + '_cls_loss'], 1e-30) < 1e-5 and rel_err(lr, g[name + '_reg_loss'], 1e-30) < 1e-5
    ref_c, ref_r = g[name + '_grad_cls'], g[name + '_grad_reg']
    assert np.array_equal(gc == 0, ref_c == 0) and np.array_equal(gr == 0, ref_r == 0)
    assert float(np.max(np.abs(gc - ref_c) - 1e-5 * np.abs(ref_c))) <= 1e-5 * float(np.abs(ref_c).max())
    assert float(np.max(np.abs(gr - ref_r) - 1e-5 * np.abs(ref_r))) <= 1e-6 * float(np.abs(ref_r).max())


def _il_step_oracle(g, name, dl=None, ig=None, method=None):
    """The oracle's composition of one IL_Loss.forward call on the fixture's inputs; returns (terms dict, grad_cls, grad_reg)
    for the fixture's weights."""
    wts = dict(zip([str(k) for k in g['weight_keys']], [float(v) for v in g['weight_vals']]))
    P = int(g['P'])
    anchors = O.anchors_for_image(int(g['h']), int(g['w']))
    logits, reg, ann = g['logits'], g['reg'], g['ann']
    n = logits.shape[0]
    replay = method is not None
    params = O.OracleParams(num_past_class=[0, P], distill=not replay)
    state = 0 if replay else 1
    clip = 0.003 if replay else 0.03
    probe = O.focal_loss(logits, reg, anchors, ann, state, params, want_grads=False, from_logits=True)
    bg_l, fg_l, w_bg, w_fg = O.clip_loss_reduce(probe['bg'], probe['fg'], clip)
    out = O.focal_loss(logits, reg, anchors, ann, state, params, w_bg=w_bg * wts['cls_bg_loss'], w_fg=w_fg * wts['cls_fg_loss'],
                       w_reg=wts['reg_loss'], from_logits=True)
    terms = {'cls_bg_loss': bg_l, 'cls_fg_loss': fg_l, 'reg_loss': out['reg_loss'][0]}
    gc, gr = out['grad_cls'].copy(), out['grad_reg'].copy()
    if replay:
        p = O.sigmoid(logits)
        loss, gp = O.enhance_error(p, P, method, g=wts['enhance_loss'])
        terms['enhance_loss'] = loss
        gc = gc + (gp * (np.float32(1.0) - p)) * p
    else:
        assert out['bg_masks'].shape[0] == n
        lc, lr, dgc, dgr = O.head_distillation(logits, reg, g['prev_logits'], g['prev_reg'], out['bg_masks'], dl, ig,
                                               g_cls=wts['dist_cls_loss'], g_reg=wts['dist_reg_loss'])
        terms.update(dist_cls_loss=lc, dist_reg_loss=lr)
        gc, gr = gc + dgc, gr + dgr
    return terms, gc, gr


def _assert_grad_close(got, ref, rel=1e-5, floor=2e-7):
    got, ref = np.asarray(got, np.float64), np.asarray(ref, np.float64)
    assert float(np.max(np.abs(got - ref) - rel * np.abs(ref))) <= floor * float(np.abs(ref).max())


@pytest.mark.parametrize('name,dl,ig', [('probs', False, False), ('logits', True, False), ('probs_ignoregd', False, True),
                                        ('logits_ignoregd', True, True)])
def test_f2_il_loss_distillation_step_matches_reference_il_loss(name, dl, ig):
    """Row f2 pinned on the reference ITSELF: tests/golden/make_golden_f2_ref.py ran the unmodified IL_Loss.forward
    (losses.py:633-737) with stub models; the oracle's Sigmoid -> FocalLoss -> clip_loss -> distillation composition must
    reproduce every returned term and the autograd gradients of their weighted sum."""
    g = load('f2_il_loss_reference')
    terms, gc, gr = _il_step_oracle(g, name, dl, ig)
    for k, v in terms.items():
        assert rel_err(v, g['distill_%s_%s' % (name, k)], 1e-30) < 1e-5, k
    _assert_grad_close(gc, g['distill_%s_grad_cls' % name])
    _assert_grad_close(gr, g['distill_%s_grad_reg' % name])


@pytest.mark.parametrize('method', ['L1', 'L2', 'L3'])
def test_f2_enhance_error_replay_step_matches_reference_il_loss(method):
    """enhance_error (losses.py:590-603) inside the reference's replay branch (:566-603), same fixture generator."""
    g = load('f2_il_loss_reference')
    terms, gc, gr = _il_step_oracle(g, method, method=method)
    for k, v in terms.items():
        assert rel_err(v, g['replay_%s_%s' % (method, k)], 1e-30) < 1e-5, k
    _assert_grad_close(gc, g['replay_%s_grad_cls' % method])
    _assert_grad_close(gr, g['replay_%s_grad_reg' % method])


def test_f3_weight_similarity_matches_reference():
    g = load('f2_il_loss_reference')
    anchors = O.anchors_for_image(int(g['h']), int(g['w']))
    sc, lab = O.weight_similarity(g['ws_probs'], anchors, g['ann'])
    assert np.array_equal(lab, g['ws_labels']) and sc.shape == g['ws_scores'].shape
    assert np.allclose(sc, g['ws_scores'], rtol=1e-6, atol=0)
    assert O.weight_similarity(g['ws_probs'], anchors, np.full_like(g['ann'], -1.0)) is None


def test_a9_collate_and_pseudo_filter_match_reference():
    g = load('a9_pseudo_labels')
    annots = [g['annot0'], g['annot1'], g['annot2']]
    assert np.array_equal(O.collate_annotations(annots), g['collated'])
    assert np.array_equal(O.collate_annotations([np.zeros((0, 5))]), g['collated_empty'])
    s, b, l = O.filter_pseudo_labels(g['pl_scores'], g['pl_boxes'], g['pl_labels'], g['pl_gt'][g['pl_gt'][:, 4] != -1][:, :4] / float(g['pl_scale']),
                                     float(g['pl_scale']))
    b = b.copy()
    b[:, 2] -= b[:, 0]
    b[:, 3] -= b[:, 1]
    assert np.array_equal(s, g['out_scores']) and np.array_equal(l, g['out_labels']) and np.array_equal(b, g['out_boxes'])


# ---- the torch-eager restatement (oracle/torch_eager.py): same fixtures, gradients through autograd ----
def eager_flags(g, params):
    """ILFlags of a focal fixture (None for state 0), as oracle/torch_eager.py takes them."""
    from oracle import torch_eager as E
    state = int(g['cur_state'])
    if state == 0:
        return None
    return E.ILFlags(past=int(g['num_past_class'][state]), ignore_past_class=params['ignore_past_class'],
                     new_ignore_past_class=params['new_ignore_past_class'], enhance_on_new=params['enhance_on_new'],
                     decrease_positive=params['decrease_positive'], decrease_positive_by_iou=params['decrease_positive_by_IOU'])


@pytest.mark.parametrize('case', FOCAL_CASES)
def test_torch_eager_focal_matches_reference(case):
    """The torch-eager restatement against EVERY focal fixture of the unmodified reference -- state 0 and all IL flag
    variants -- so that the -m gpu tests can use it as a full-size, same-device oracle for the flags too."""
    import torch
    from oracle import torch_eager as E
    g = load('focal_' + case)
    params = golden_params(g)
    il = eager_flags(g, params)
    anchors = torch.from_numpy(O.anchors_for_image(int(g['h']), int(g['w'])))
    cls = torch.from_numpy(g['cls']).requires_grad_(True)
    reg = torch.from_numpy(g['reg']).requires_grad_(True)
    out = E.focal_loss(cls, reg, anchors, torch.from_numpy(g['ann']), params['alpha'], params['gamma'], il)
    bg, fg, rl = out[:3]
    # the fixture's upstream weights: dL/dbg_j = wb_j, dL/dfg_j = wf_j, dL/dreg_loss = wr, dL/d(enhance_on_new_loss) = we
    total = (bg * torch.from_numpy(g['wb'])).sum() + (fg * torch.from_numpy(g['wf'])).sum() + rl.sum() * float(g['wr'])
    if il is not None and il.enhance_on_new:
        total = total + out[3] * float(g['we'])
        assert rel_err(out[3].detach().numpy(), g['enhance_on_new_loss'], 1e-30) < 1e-6
    total.backward()
    # the same torch ops in the same order as the reference on the same CPU: identical bits
    assert np.array_equal(bg.detach().numpy(), g['bg'])
    assert np.array_equal(fg.detach().numpy(), g['fg'])
    assert np.array_equal(rl.detach().numpy(), g['reg_loss'])
    assert rel_err(cls.grad.numpy(), g['grad_cls'], 1e-30) < 1e-6
    assert np.array_equal(cls.grad.numpy() == 0, g['grad_cls'] == 0)
    reg_grad = np.zeros_like(g['grad_reg']) if reg.grad is None else reg.grad.numpy()     # no positives anywhere: no graph
    assert float(np.abs(reg_grad - g['grad_reg']).max()) <= 1e-6 * float(np.abs(g['grad_reg']).max())


def test_torch_eager_decode_and_predict_match_reference():
    import torch
    from oracle import torch_eager as E
    g = load('decode')
    h, w = int(g['h']), int(g['w'])
    anchors = torch.from_numpy(O.anchors_for_image(h, w))
    dec = E.decode_boxes(anchors, torch.from_numpy(g['reg']))
    assert np.array_equal(dec.numpy(), g['decoded'])
    assert np.array_equal(E.clip_boxes(dec.clone(), h, w).numpy(), g['clipped'])
    for name in ('trick', 'none'):
        p = load('predict_' + name)
        h, w = int(p['h']), int(p['w'])
        anchors = torch.from_numpy(O.anchors_for_image(h, w))
        s, l, b = E.predict(torch.from_numpy(p['logits']), torch.from_numpy(p['reg']), anchors, h, w)
        assert np.array_equal(s.numpy(), p['scores']) and np.array_equal(l.numpy(), p['labels'])
        assert np.array_equal(b.numpy(), p['boxes'])


@pytest.mark.parametrize('seed,h,w,C,N,G', [(0, 128, 160, 8, 3, 6), (1, 96, 96, 20, 2, 12), (2, 33, 70, 3, 2, 4)])
def test_two_oracles_agree_on_random_inputs(seed, h, w, C, N, G):
    """The numpy oracle (analytic gradients) and the torch-eager restatement (autograd) are independent statements of the same
    reference code: on seeded random inputs beyond the fixtures they must agree on assignments-derived losses and on every
    gradient element within the north_star tolerance."""
    import torch
    from oracle import torch_eager as E
    from tests.helpers import synth_gt, synth_head
    rng = np.random.default_rng(seed)
    anchors = O.anchors_for_image(h, w)
    A = anchors.shape[1]
    _, probs, reg = synth_head(rng, N, A, C, mu=-3.0)
    ann = synth_gt(rng, N, G, h, w, C, empty=(N - 1,))
    wb, wf = rng.uniform(0.5, 1.5, N), rng.uniform(0.5, 1.5, N)
    ref = O.focal_loss(probs, reg, anchors, ann, 0, O.OracleParams(), w_bg=wb, w_fg=wf, w_reg=0.7)
    cls_t = torch.from_numpy(probs).requires_grad_(True)
    reg_t = torch.from_numpy(reg).requires_grad_(True)
    bg, fg, rl = E.focal_loss(cls_t, reg_t, torch.from_numpy(anchors), torch.from_numpy(ann))
    ((bg * torch.from_numpy(wb).float()).sum() + (fg * torch.from_numpy(wf).float()).sum() + 0.7 * rl.sum()).backward()
    assert rel_err(bg.detach().numpy(), ref['bg'], 1e-30) < 1e-5 and rel_err(fg.detach().numpy(), ref['fg'], 1e-30) < 1e-5
    assert rel_err(rl.detach().numpy(), ref['reg_loss'], 1e-30) < 1e-5
    gc = cls_t.grad.numpy()
    assert np.array_equal(gc == 0, ref['grad_cls'] == 0)
    assert np.max(np.abs(gc - ref['grad_cls']) / (np.abs(ref['grad_cls']) + 1e-12 * np.abs(ref['grad_cls']).max())) < 1e-5
    gr = np.zeros_like(ref['grad_reg']) if reg_t.grad is None else reg_t.grad.numpy()
    assert float((np.abs(gr - ref['grad_reg']) - 1e-5 * np.abs(ref['grad_reg'])).max()) <= 1e-5 * float(np.abs(ref['grad_reg']).max())
