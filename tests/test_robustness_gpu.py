"""GPU robustness tests: thread safety (the reference's evaluator calls predict from up to 10 threads, evaluator.py:400-422),
CUDA-graph capture of the sync-free paths, repeated calls on the cached workspace, dense-GT stress shape (BASELINE config 5),
multi-tile GT (> 256 rows), large candidate counts in the reference-faithful (no top-k) mode."""
import threading

import numpy as np
import pytest
import torch

import cl_object_detection_b200 as cld
from cl_object_detection_b200 import _lib
from cl_object_detection_b200 import detect as D
from cl_object_detection_b200.params import to_loss_params
from oracle import head_oracle as O
from tests.helpers import synth_gt, synth_head

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def cu(x):
    return torch.from_numpy(np.ascontiguousarray(x)).to(DEV)


def test_predict_and_loss_from_ten_threads():
    h, w, C = 192, 256, 20
    anchors = cld.generate_anchors(h, w, DEV)
    A = anchors.shape[1]
    rng = np.random.default_rng(11)
    cases = []
    for t in range(10):
        logits, probs, reg = synth_head(rng, 1, A, C, mu=-5.0)
        ann = synth_gt(rng, 1, 6, h, w, C)
        cases.append((logits, probs, reg * 0.4, ann))
    want_det = [O.detect(torch.sigmoid(cu(c[0])).cpu().numpy(), c[2], O.anchors_for_image(h, w), h, w, is_logits=False) for c in cases]
    want_loss = [O.focal_loss(c[1], c[2], O.anchors_for_image(h, w), c[3], 0, O.OracleParams(), want_grads=False) for c in cases]
    errors = []

    def worker(t):
        try:
            stream = torch.cuda.Stream(device=DEV)
            with torch.cuda.stream(stream):
                for _ in range(5):
                    logits, probs, reg, ann = (cu(x) for x in cases[t])
                    s, l, b = D.predict_from_head(logits, reg, anchors, torch.zeros(1, 3, h, w, device=DEV))
                    assert np.array_equal(s.cpu().numpy(), want_det[t]['scores'])
                    assert np.array_equal(l.cpu().numpy(), want_det[t]['labels'])
                    with torch.no_grad():
                        out = cld.FocalLoss()(probs, reg, anchors, ann, 0, cld.HeadParams())
                    assert np.allclose(out['cls_loss'][0].cpu().numpy(), want_loss[t]['bg'], rtol=1e-5, atol=0)
                    assert np.allclose(out['cls_loss'][1].cpu().numpy(), want_loss[t]['fg'], rtol=1e-5, atol=0)
        except Exception as e:  # noqa: BLE001
            errors.append((t, repr(e)))
    threads = [threading.Thread(target=worker, args=(t,)) for t in range(10)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert not errors, errors


def test_loss_path_is_cuda_graph_capturable_and_repeatable():
    """The fused loss call has no host sync and no per-call memset: capture it in a CUDA graph and replay."""
    h, w, C, N, G = 160, 192, 8, 3, 7
    rng = np.random.default_rng(5)
    anchors = cld.generate_anchors(h, w, DEV)
    A = anchors.shape[1]
    _, probs, reg = synth_head(rng, N, A, C, mu=-3.0)
    ann = synth_gt(rng, N, G, h, w, C, empty=(1,))
    lib = _lib.load()
    lp = to_loss_params(cld.HeadParams(), 0, C)
    p, r, an = cu(probs), cu(reg), cu(ann)
    weights = torch.full((4, N), 1.0 / N, device=DEV)
    gcls, greg = torch.empty_like(p), torch.empty_like(r)
    losses = torch.empty((4, N), device=DEV)
    meta = torch.empty((N, A), dtype=torch.int32, device=DEV)
    npos = torch.empty(N, dtype=torch.int32, device=DEV)
    nvalid = torch.empty(N, dtype=torch.int32, device=DEV)
    ws = torch.zeros(lib.cldet_focal_loss_workspace_bytes(N, A), dtype=torch.uint8, device=DEV)

    def call():
        _lib.check(lib.cldet_focal_loss(p.data_ptr(), r.data_ptr(), anchors.data_ptr(), an.data_ptr(), N, A, C, G, lp,
                                        weights.data_ptr(), None, gcls.data_ptr(), greg.data_ptr(), losses.data_ptr(), meta.data_ptr(),
                                        None, npos.data_ptr(), nvalid.data_ptr(), None, None, ws.data_ptr(), ws.numel(),
                                        torch.cuda.current_stream().cuda_stream))
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        call()                                   # warm-up outside capture
        torch.cuda.current_stream().synchronize()
        first = (losses.clone(), gcls.clone(), npos.clone())
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            call()
        for _ in range(3):
            losses.zero_(); gcls.zero_(); npos.zero_()
            g.replay()
        torch.cuda.current_stream().synchronize()
    assert torch.equal(losses, first[0]) and torch.equal(gcls, first[1]) and torch.equal(npos, first[2])
    ref = O.focal_loss(probs, reg, O.anchors_for_image(h, w), ann, 0, O.OracleParams(), want_grads=False)
    assert np.allclose(losses[0].cpu().numpy(), ref['bg'], rtol=1e-5, atol=0)
    assert np.array_equal(npos.cpu().numpy(), ref['npos'])
    assert int(ws[: 12 * N].sum()) == 0          # the workspace header is left zeroed


def test_dense_gt_stress_config5_shape():
    """BASELINE config 5: 1333x1333, C=80, exactly 100 GT boxes per image; assignment bit-exact, loss within 1e-5."""
    h = w = 1333
    C, N, G = 80, 1, 100
    rng = np.random.default_rng(55)
    oa = O.anchors_for_image(h, w)
    A = oa.shape[1]
    ann = synth_gt(rng, N, G, h, w, C, exact=True)
    gen = torch.Generator(device=DEV).manual_seed(55)
    probs = torch.sigmoid(torch.randn(N, A, C, device=DEV, generator=gen) * 2 - 4)
    reg = torch.randn(N, A, 4, device=DEV, generator=gen)
    anchors = cld.generate_anchors(h, w, DEV)
    asg = cld.iou_assign(anchors, cu(ann), C)
    ref = O.assign(oa[0], ann[0])
    assert np.array_equal(asg['state'][0].cpu().numpy(), ref['state'])
    assert np.array_equal(asg['argmax'][0].cpu().numpy(), ref['argmax'])
    assert int(asg['npos'][0]) == ref['npos'] and ref['npos'] > 1000
    with torch.no_grad():
        out = cld.FocalLoss()(probs, reg, anchors, cu(ann), 0, cld.HeadParams())
    want = O.focal_loss(probs.cpu().numpy(), reg.cpu().numpy(), oa, ann, 0, O.OracleParams(), want_grads=False)
    assert np.allclose(out['cls_loss'][0].cpu().numpy(), want['bg'], rtol=1e-5, atol=0)
    assert np.allclose(out['cls_loss'][1].cpu().numpy(), want['fg'], rtol=1e-5, atol=0)
    assert np.allclose(out['reg_loss'].cpu().numpy(), want['reg_loss'], rtol=1e-5, atol=0)


def test_many_gt_rows_multi_tile_and_max_rows():
    """More GT rows than one shared-memory tile (256) incl. scattered padding, and the table overflow path (> 256 valid rows)."""
    h, w, C = 320, 320, 4
    rng = np.random.default_rng(77)
    oa = O.anchors_for_image(h, w)
    ann = synth_gt(rng, 2, 700, h, w, C, exact=True)
    ann[0, rng.choice(700, 300, replace=False)] = -1        # 400 valid rows, scattered
    ann[1, 5:] = -1                                          # 5 valid rows among 700
    got = cld.iou_assign(cld.generate_anchors(h, w, DEV), cu(ann), C)
    for j in range(2):
        ref = O.assign(oa[0], ann[j])
        assert int(got['nvalid'][j]) == ref['valid']
        assert np.array_equal(got['state'][j].cpu().numpy(), ref['state'])
        assert np.array_equal(got['argmax'][j].cpu().numpy(), ref['argmax'])
        pos = ref['state'] == 1
        assert np.array_equal(got['label'][j].cpu().numpy()[pos], ref['label'][pos])
        # the raw row stored in the assignment word points at the same GT box as the compacted index
        valid_rows = np.nonzero(ann[j, :, 4] != -1)[0]
        raw = (got['meta'][j].cpu().numpy().view(np.uint32) >> 16).astype(np.int64)
        assert np.array_equal(raw[pos], valid_rows[ref['argmax'][pos]])


def test_large_candidate_count_reference_mode():
    """No top-k (the reference's behaviour) with ~12k candidates in one image: vanilla/trick switch + large bitmask."""
    h, w, C = 512, 512, 20
    anchors = cld.generate_anchors(h, w, DEV)
    A = anchors.shape[1]
    gen = torch.Generator(device=DEV).manual_seed(9)
    logits = torch.randn(1, A, C, device=DEV, generator=gen) * 2 - 6.0
    reg = torch.randn(1, A, 4, device=DEV, generator=gen) * 0.3
    s, l, b = D.detect_batch(logits, reg, anchors, h, w)[0]
    probs = torch.sigmoid(logits).cpu().numpy()
    ref = O.detect(probs, reg.cpu().numpy(), O.anchors_for_image(h, w), h, w, is_logits=False, device_rule='cuda')
    assert ref['cand_scores'].shape[0] > 8000
    assert np.array_equal(s.cpu().numpy(), ref['scores']) and np.array_equal(l.cpu().numpy(), ref['labels'])


def test_conv_layout_entries_from_threads_and_in_a_cuda_graph():
    """The conv-layout entry points: six host threads on their own streams, then the fused call captured in
    a CUDA graph and replayed (no host sync, no per-call memset, scratch left clean)."""
    h, w, C, N, G = 160, 192, 8, 2, 5
    rng = np.random.default_rng(21)
    anchors = cld.generate_anchors(h, w, DEV)
    shapes = [((h + 2 ** l - 1) // 2 ** l, (w + 2 ** l - 1) // 2 ** l) for l in range(3, 8)]
    gen = torch.Generator(device=DEV).manual_seed(21)
    cls_lv = [torch.sigmoid(torch.randn(N, 9 * C, hl, wl, device=DEV, generator=gen) * 2 - 3) for hl, wl in shapes]
    reg_lv = [torch.randn(N, 36, hl, wl, device=DEV, generator=gen) * 0.4 for hl, wl in shapes]
    ann = cu(synth_gt(rng, N, G, h, w, C))
    fl = cld.FocalLoss()
    with torch.no_grad():
        want = fl.forward_head(cls_lv, reg_lv, anchors, ann, 0, cld.HeadParams(), (h, w))
        want_det = D.detect_batch_head(cls_lv, reg_lv, anchors, h, w, is_logits=False)
    torch.cuda.synchronize()
    errors = []

    def worker(t):
        try:
            with torch.cuda.stream(torch.cuda.Stream(device=DEV)):
                for _ in range(5):
                    with torch.no_grad():
                        out = cld.FocalLoss().forward_head(cls_lv, reg_lv, anchors, ann, 0, cld.HeadParams(), (h, w))
                    det = D.detect_batch_head(cls_lv, reg_lv, anchors, h, w, is_logits=False)
                    assert torch.equal(out['cls_loss'][0], want['cls_loss'][0]) and torch.equal(out['cls_loss'][1], want['cls_loss'][1])
                    for (s0, l0, b0), (s1, l1, b1) in zip(want_det, det):
                        assert torch.equal(s0, s1) and torch.equal(l0, l1) and torch.equal(b0, b1)
        except Exception as e:  # noqa: BLE001
            errors.append((t, repr(e)))
    threads = [threading.Thread(target=worker, args=(t,)) for t in range(6)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert not errors, errors

    lib = _lib.load()
    A = anchors.shape[1]
    lp = to_loss_params(cld.HeadParams(), 0, C)
    lp.image_height, lp.image_width = h, w
    weights = torch.full((4, N), 1.0 / N, device=DEV)
    gcls, greg = [torch.empty_like(t) for t in cls_lv], [torch.empty_like(t) for t in reg_lv]
    losses = torch.empty((4, N), device=DEV)
    meta = torch.empty((N, A), dtype=torch.int32, device=DEV)
    npos = torch.empty(N, dtype=torch.int32, device=DEV)
    nvalid = torch.empty(N, dtype=torch.int32, device=DEV)
    ws = torch.zeros(lib.cldet_focal_loss_workspace_bytes(N, A), dtype=torch.uint8, device=DEV)
    pc, pr, pgc, pgr = (_lib.ptr_array(x) for x in (cls_lv, reg_lv, gcls, greg))

    def call():
        _lib.check(lib.cldet_focal_loss_head(pc, pr, 5, h, w, anchors.data_ptr(), ann.data_ptr(), N, C, G, lp, weights.data_ptr(), None,
                                             pgc, pgr, losses.data_ptr(), meta.data_ptr(), None, npos.data_ptr(), nvalid.data_ptr(),
                                             None, None, ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream))
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        call()
        torch.cuda.current_stream().synchronize()
        first = (losses.clone(), gcls[0].clone(), greg[0].clone())
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            call()
        for _ in range(3):
            losses.zero_(); gcls[0].zero_(); greg[0].zero_()
            g.replay()
        torch.cuda.current_stream().synchronize()
    assert torch.equal(losses, first[0]) and torch.equal(gcls[0], first[1]) and torch.equal(greg[0], first[2])
    assert torch.equal(losses[0], want['cls_loss'][0]) and torch.equal(losses[1], want['cls_loss'][1])
    # counters / accumulators (header) and the assignment keys + bitmap (after the per-block partials) are left zeroed
    header = (12 * N + 255) // 256 * 256
    partials = (N * ((A + 31) // 32 + 72) * 16 + 255) // 256 * 256
    assert int(ws[:12 * N].sum()) == 0 and int(ws[header + partials:].sum()) == 0
