"""Out-of-bounds WRITE detection without compute-sanitizer (closed on this pool): every output buffer of the main C-ABI calls
is carved out of a larger allocation with sentinel-filled guard bands on both sides; after the call the bands must be
untouched.  Shapes are chosen ragged on purpose (anchor counts that are not multiples of any block/tile size, all three
vector widths of the loss kernel, top-k on and off)."""
import numpy as np
import pytest
import torch

import cl_object_detection_b200 as cld
from cl_object_detection_b200 import _lib
from cl_object_detection_b200.params import to_loss_params
from tests.helpers import synth_gt

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'
GUARD = 4096      # bytes on each side
SENT = 0xA5


class Guarded:
    def __init__(self):
        self.items = []

    def alloc(self, shape, dtype):
        nbytes = int(np.prod(shape)) * torch.empty(0, dtype=dtype).element_size()
        pad = (-nbytes) % 256
        raw = torch.full((GUARD + nbytes + pad + GUARD,), SENT, dtype=torch.uint8, device=DEV)
        view = raw[GUARD:GUARD + nbytes].view(dtype).view(shape)
        self.items.append((raw, nbytes))
        return view

    def check(self):
        for k, (raw, nbytes) in enumerate(self.items):
            assert bool((raw[:GUARD] == SENT).all()), 'buffer %d: write before the start' % k
            assert bool((raw[GUARD + nbytes:] == SENT).all()), 'buffer %d: write past the end' % k


@pytest.mark.parametrize('h,w,C,N,G', [(200, 264, 80, 3, 9), (136, 200, 20, 2, 5), (96, 104, 7, 2, 3), (72, 72, 16, 5, 300)])
def test_loss_path_writes_stay_in_bounds(h, w, C, N, G):
    lib = _lib.load()
    rng = np.random.default_rng(C)
    anchors = cld.generate_anchors(h, w, DEV)
    A = anchors.shape[1]
    probs = torch.rand(N, A, C, device=DEV) * 0.2
    reg = torch.randn(N, A, 4, device=DEV)
    ann = torch.from_numpy(synth_gt(rng, N, G, h, w, C, empty=(1,))).to(DEV)
    g = Guarded()
    gcls, greg = g.alloc((N, A, C), torch.float32), g.alloc((N, A, 4), torch.float32)
    losses, baked = g.alloc((4, N), torch.float32), g.alloc((4, N), torch.float32)
    meta, iou = g.alloc((N, A), torch.int32), g.alloc((N, A), torch.float32)
    npos, nvalid = g.alloc((N,), torch.int32), g.alloc((N,), torch.int32)
    mask, status = g.alloc((N, A), torch.uint8), g.alloc((1,), torch.int32)
    ws_bytes = lib.cldet_focal_loss_workspace_bytes(N, A)
    ws = g.alloc((ws_bytes,), torch.uint8)
    ws.zero_()
    weights = torch.full((4, N), 0.5, device=DEV)
    st = torch.cuda.current_stream().cuda_stream
    for variant in range(2):
        params = cld.HeadParams([0, max(1, C // 2)], ignore_past_class=True, new_ignore_past_class=True, enhance_on_new=True,
                                decrease_positive_by_IOU=True, distill=True) if variant else cld.HeadParams()
        lp = to_loss_params(params, variant, C)
        _lib.check(lib.cldet_focal_loss(probs.data_ptr(), reg.data_ptr(), anchors.data_ptr(), ann.data_ptr(), N, A, C, G, lp,
                                        weights.data_ptr(), baked.data_ptr(), gcls.data_ptr(), greg.data_ptr(), losses.data_ptr(),
                                        meta.data_ptr(), iou.data_ptr(), npos.data_ptr(), nvalid.data_ptr(), mask.data_ptr(),
                                        status.data_ptr(), ws.data_ptr(), ws_bytes, st))
        w2 = weights * 2
        _lib.check(lib.cldet_focal_loss_reweight(probs.data_ptr(), reg.data_ptr(), anchors.data_ptr(), ann.data_ptr(), N, A, C, G,
                                                 lp, w2.data_ptr(), baked.data_ptr(), gcls.data_ptr(), greg.data_ptr(),
                                                 meta.data_ptr(), iou.data_ptr(), npos.data_ptr(), ws.data_ptr(), ws_bytes, st))
        torch.cuda.synchronize()
        g.check()
        assert torch.isfinite(losses).all() and torch.isfinite(gcls).all()
        assert torch.equal(baked, w2)
    # status codes instead of faults for bad arguments
    assert lib.cldet_focal_loss(probs.data_ptr(), reg.data_ptr(), anchors.data_ptr(), ann.data_ptr(), N, A, C, G, lp,
                                weights.data_ptr(), None, gcls.data_ptr(), greg.data_ptr(), losses.data_ptr(), meta.data_ptr(),
                                iou.data_ptr(), npos.data_ptr(), nvalid.data_ptr(), None, None, ws.data_ptr(), 16, st) == 2
    assert lib.cldet_focal_loss(probs.data_ptr(), reg.data_ptr(), anchors.data_ptr(), ann.data_ptr(), N, A, 0, G, lp,
                                weights.data_ptr(), None, gcls.data_ptr(), greg.data_ptr(), losses.data_ptr(), meta.data_ptr(),
                                iou.data_ptr(), npos.data_ptr(), nvalid.data_ptr(), None, None, ws.data_ptr(), ws_bytes, st) == 1


@pytest.mark.parametrize('h,w,C,N,G', [(200, 264, 80, 3, 9), (136, 200, 20, 2, 5), (96, 104, 7, 2, 3), (33, 70, 4, 2, 6), (1, 1, 3, 2, 2)])
def test_head_layout_writes_stay_in_bounds(h, w, C, N, G):
    """cldet_focal_loss_head / _reweight: ten per-level gradient tensors + the anchor-keyed outputs, all guard-banded; level
    planes here are ragged (not multiples of 4 floats, shorter than a block's 512 positions, down to 1 x 1)."""
    lib = _lib.load()
    rng = np.random.default_rng(C + h)
    anchors = cld.generate_anchors(h, w, DEV)
    A = anchors.shape[1]
    shapes = [((h + 2 ** l - 1) // 2 ** l, (w + 2 ** l - 1) // 2 ** l) for l in range(3, 8)]
    cls_lv = [torch.rand(N, 9 * C, hl, wl, device=DEV) * 0.2 for hl, wl in shapes]
    reg_lv = [torch.randn(N, 36, hl, wl, device=DEV) for hl, wl in shapes]
    ann = torch.from_numpy(synth_gt(rng, N, G, max(h, 32), max(w, 32), C, empty=(1,))).to(DEV)
    g = Guarded()
    gcls = [g.alloc(tuple(t.shape), torch.float32) for t in cls_lv]
    greg = [g.alloc(tuple(t.shape), torch.float32) for t in reg_lv]
    losses, baked = g.alloc((4, N), torch.float32), g.alloc((4, N), torch.float32)
    meta, iou = g.alloc((N, A), torch.int32), g.alloc((N, A), torch.float32)
    npos, nvalid = g.alloc((N,), torch.int32), g.alloc((N,), torch.int32)
    mask, status = g.alloc((N, A), torch.uint8), g.alloc((1,), torch.int32)
    ws_bytes = lib.cldet_focal_loss_workspace_bytes(N, A)
    ws = g.alloc((ws_bytes,), torch.uint8)
    ws.zero_()
    weights = torch.full((4, N), 0.5, device=DEV)
    st = torch.cuda.current_stream().cuda_stream
    pc, pr, pgc, pgr = (_lib.ptr_array(x) for x in (cls_lv, reg_lv, gcls, greg))
    for variant in range(2):
        params = cld.HeadParams([0, max(1, C // 2)], ignore_past_class=True, new_ignore_past_class=True, enhance_on_new=True,
                                decrease_positive_by_IOU=True, distill=True) if variant else cld.HeadParams()
        lp = to_loss_params(params, variant, C)
        _lib.check(lib.cldet_focal_loss_head(pc, pr, 5, h, w, anchors.data_ptr(), ann.data_ptr(), N, C, G, lp, weights.data_ptr(),
                                             baked.data_ptr(), pgc, pgr, losses.data_ptr(), meta.data_ptr(), iou.data_ptr(),
                                             npos.data_ptr(), nvalid.data_ptr(), mask.data_ptr(), status.data_ptr(), ws.data_ptr(),
                                             ws_bytes, st))
        w2 = weights * 2
        _lib.check(lib.cldet_focal_loss_head_reweight(pc, pr, 5, h, w, anchors.data_ptr(), ann.data_ptr(), N, C, G, lp,
                                                      w2[0].data_ptr(), 1, w2[1].data_ptr(), 1, w2[2].data_ptr(), 1, w2[3].data_ptr(), 1,
                                                      baked.data_ptr(), pgc, pgr, meta.data_ptr(), iou.data_ptr(), npos.data_ptr(),
                                                      ws.data_ptr(), ws_bytes, st))
        torch.cuda.synchronize()
        g.check()
        assert torch.isfinite(losses).all() and all(torch.isfinite(t).all() for t in gcls + greg)
        assert torch.equal(baked, w2)
    # the scratch (keys, counters) is left zeroed: a second pair of calls gives the same losses
    first = losses.clone()
    _lib.check(lib.cldet_focal_loss_head(pc, pr, 5, h, w, anchors.data_ptr(), ann.data_ptr(), N, C, G, lp, weights.data_ptr(),
                                         baked.data_ptr(), pgc, pgr, losses.data_ptr(), meta.data_ptr(), iou.data_ptr(),
                                         npos.data_ptr(), nvalid.data_ptr(), mask.data_ptr(), status.data_ptr(), ws.data_ptr(),
                                         ws_bytes, st))
    torch.cuda.synchronize()
    assert torch.equal(first, losses)
    # status codes instead of faults for bad arguments
    assert lib.cldet_focal_loss_head(pc, pr, 4, h, w, anchors.data_ptr(), ann.data_ptr(), N, C, G, lp, weights.data_ptr(),
                                     baked.data_ptr(), pgc, pgr, losses.data_ptr(), meta.data_ptr(), iou.data_ptr(), npos.data_ptr(),
                                     nvalid.data_ptr(), None, None, ws.data_ptr(), ws_bytes, st) == 1
    assert lib.cldet_focal_loss_head(pc, pr, 5, h, w, anchors.data_ptr(), ann.data_ptr(), N, C, G, lp, weights.data_ptr(),
                                     baked.data_ptr(), pgc, pgr, losses.data_ptr(), meta.data_ptr(), iou.data_ptr(), npos.data_ptr(),
                                     nvalid.data_ptr(), None, None, ws.data_ptr(), 16, st) == 2


@pytest.mark.parametrize('C,topk', [(80, 100), (20, 0), (7, 33)])
def test_detection_pipeline_writes_stay_in_bounds(C, topk):
    lib = _lib.load()
    h, w, N = 200, 264, 3
    anchors = cld.generate_anchors(h, w, DEV)
    A = anchors.shape[1]
    gen = torch.Generator(device=DEV).manual_seed(C)
    logits = torch.randn(N, A, C, device=DEV, generator=gen) * 2 - 4
    reg = torch.randn(N, A, 4, device=DEV, generator=gen) * 0.4
    g = Guarded()
    counts = g.alloc((N,), torch.int32)
    counts.zero_()
    cand, keys = g.alloc((N, A, 32), torch.uint8), g.alloc((N, A), torch.int64)
    st = torch.cuda.current_stream().cuda_stream
    _lib.check(lib.cldet_decode_filter(logits.data_ptr(), 1, reg.data_ptr(), anchors.data_ptr(), N, A, C, h, w, 0.05,
                                       cand.data_ptr(), keys.data_ptr(), A, counts.data_ptr(), st))
    kmax = int(counts.max().item())
    cap = min(topk, A) if topk else kmax
    max_count = A if topk else kmax
    sorted_c, sorted_counts = g.alloc((N, cap, 32), torch.uint8), g.alloc((N,), torch.int32)
    sws = g.alloc((lib.cldet_sort_workspace_bytes(N, max_count, topk),), torch.uint8)
    _lib.check(lib.cldet_sort_candidates(cand.data_ptr(), keys.data_ptr(), counts.data_ptr(), N, A, max_count, topk,
                                         sorted_c.data_ptr(), cap, sorted_counts.data_ptr(), sws.data_ptr(), sws.numel(), st))
    nws = g.alloc((lib.cldet_nms_workspace_bytes(N, cap),), torch.uint8)
    keep, keep_counts = g.alloc((N, cap), torch.int32), g.alloc((N,), torch.int32)
    _lib.check(lib.cldet_nms_sorted(sorted_c.data_ptr(), sorted_counts.data_ptr(), N, cap, cap, 0.5, 0, 100000, keep.data_ptr(),
                                    keep_counts.data_ptr(), nws.data_ptr(), nws.numel(), st))
    scores, labels = g.alloc((N, cap), torch.float32), g.alloc((N, cap), torch.int64)
    boxes = g.alloc((N, cap, 4), torch.float32)
    _lib.check(lib.cldet_gather_detections(sorted_c.data_ptr(), keep.data_ptr(), keep_counts.data_ptr(), N, cap, cap,
                                           scores.data_ptr(), labels.data_ptr(), boxes.data_ptr(), st))
    rec, offs = g.alloc((N * cap, 8), torch.int32), g.alloc((N + 1,), torch.int32)
    scales = torch.tensor([1.0, 1.5, 0.7], device=DEV)
    _lib.check(lib.cldet_coco_results(scores.data_ptr(), labels.data_ptr(), boxes.data_ptr(), keep_counts.data_ptr(),
                                      scales.data_ptr(), N, cap, 0.2, rec.data_ptr(), offs.data_ptr(), st))
    torch.cuda.synchronize()
    g.check()
    assert int(keep_counts.min()) > 0 and int(offs[N]) <= int(keep_counts.sum())


@pytest.mark.parametrize('h,w,C', [(200, 264, 80), (96, 104, 7), (33, 70, 4)])
def test_conv_layout_filter_writes_stay_in_bounds(h, w, C):
    """cldet_decode_filter_head with a candidate capacity SMALLER than the number of candidates: counts may exceed the
    capacity, records and keys must not be written past it."""
    lib = _lib.load()
    N = 2
    anchors = cld.generate_anchors(h, w, DEV)
    A = anchors.shape[1]
    shapes = [((h + 2 ** l - 1) // 2 ** l, (w + 2 ** l - 1) // 2 ** l) for l in range(3, 8)]
    cls_lv = [torch.randn(N, 9 * C, hl, wl, device=DEV) for hl, wl in shapes]          # ~every anchor is a candidate
    reg_lv = [torch.randn(N, 36, hl, wl, device=DEV) * 0.3 for hl, wl in shapes]
    cap = max(1, A // 3)
    g = Guarded()
    cand, keys = g.alloc((N, cap, 32), torch.uint8), g.alloc((N, cap), torch.int64)
    counts = g.alloc((N,), torch.int32)
    counts.zero_()
    st = torch.cuda.current_stream().cuda_stream
    _lib.check(lib.cldet_decode_filter_head(_lib.ptr_array(cls_lv), _lib.ptr_array(reg_lv), 5, h, w, 1, anchors.data_ptr(), N, C,
                                            0.05, cand.data_ptr(), keys.data_ptr(), cap, counts.data_ptr(), st))
    torch.cuda.synchronize()
    g.check()
    assert int(counts.min()) > cap
    assert lib.cldet_decode_filter_head(_lib.ptr_array(cls_lv), _lib.ptr_array(reg_lv), 4, h, w, 1, anchors.data_ptr(), N, C, 0.05,
                                        cand.data_ptr(), keys.data_ptr(), cap, counts.data_ptr(), st) == 1
