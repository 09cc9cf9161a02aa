"""Golden vectors for SURVEY 8(f) rows f2 and f3 produced by RUNNING THE UNMODIFIED REFERENCE CODE (build container only):

  retinanet/losses.py  IL_Loss.forward (:518-739) itself -- instantiated with a SimpleNamespace trainer whose `model` /
                       `prev_model` are stubs returning seeded synthetic head outputs -- for
                         * the incremental branch with distillation (:633-737): Sigmoid -> FocalLoss (+bg_masks) -> clip_loss
                           -> dist_cls_loss / dist_reg_loss, all four (distill_logits, ignore_GD) combinations, and
                         * the replay branch (:566-603) with enhance_error L1 / L2 / L3;
                       results AND the autograd gradients of a weighted sum of every returned term w.r.t. the head outputs.
  IL_method/weight_init.py  Weight_similarity.forward (:82-115) with a stub model.

This replaces the restatement in make_golden_f2.py as the pin of row f2 (that fixture stays as a second check).
Same CPU shim as make_golden.py (no edits to reference source).
"""
import os
import sys
import types

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import make_golden as mg  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
WEIGHTS = dict(cls_bg_loss=1.0, cls_fg_loss=0.9, reg_loss=1.1, dist_cls_loss=0.6, dist_reg_loss=1.7, enhance_loss=0.8)


def main():
    mg.install_cpu_shim()
    from retinanet.anchors import Anchors
    from retinanet.losses import IL_Loss
    from IL_method.weight_init import Weight_similarity

    torch.set_num_threads(2)
    rng = np.random.default_rng(2202)
    h, w, C, P, N, G = 96, 128, 6, 4, 3, 5
    anchors = Anchors()(torch.zeros(1, 3, h, w))
    A = anchors.shape[1]
    logits = rng.normal(-3.0, 2.0, (N, A, C)).astype(np.float32)
    reg = rng.normal(0, 1.0, (N, A, 4)).astype(np.float32)
    prev_logits = (logits[:, :, :P] + rng.normal(0, 0.7, (N, A, P))).astype(np.float32)
    prev_reg = (reg + rng.normal(0, 0.9, (N, A, 4))).astype(np.float32)
    ann = mg.make_gt(rng, N, G, h, w, C, empty=())       # distillation needs GT in every image (bg_masks has one row per image with GT)
    feats = [rng.normal(0, 1, (N, 8, 2, 2)).astype(np.float32)]
    prev_feats = [rng.normal(0, 1, (N, 8, 2, 2)).astype(np.float32)]
    d = dict(h=h, w=w, P=P, logits=logits, reg=reg, prev_logits=prev_logits, prev_reg=prev_reg, ann=ann,
             weight_keys=np.array(list(WEIGHTS)), weight_vals=np.array(list(WEIGHTS.values()), np.float32))

    def run(params, cur_state, is_replay):
        tc = torch.from_numpy(logits).requires_grad_(True)
        tr = torch.from_numpy(reg).requires_grad_(True)

        def model(img, return_feat=False, return_anchor=True, enable_act=True):
            cls_out = torch.sigmoid(tc) if enable_act else tc
            if return_feat:
                return cls_out, tr, [torch.from_numpy(f) for f in feats], anchors
            return cls_out, tr, anchors

        def prev_model(img, return_feat=True, return_anchor=False, enable_act=False):
            return torch.from_numpy(prev_logits), torch.from_numpy(prev_reg), [torch.from_numpy(f) for f in prev_feats]

        trainer = types.SimpleNamespace(model=model, prev_model=prev_model, params=params, cur_state=cur_state, cur_warm_stage=-1,
                                        bic=None, cur_epoch=1, end_epoch=10, protoTyper=None)
        res = IL_Loss(trainer).forward(torch.zeros(N, 3, h, w), torch.from_numpy(ann), is_replay=is_replay)
        total = sum(WEIGHTS[k] * v for k, v in res.items() if k in WEIGHTS)
        total.backward()
        return res, tc.grad.numpy(), tr.grad.numpy()

    # ---- incremental state with distillation: IL_Loss.forward :633-737 ----
    for name, (dl, ig) in dict(probs=(False, False), logits=(True, False), probs_ignoregd=(False, True),
                               logits_ignoregd=(True, True)).items():
        params = mg.Params([0, P], distill=True, distill_logits=dl, ignore_GD=ig, clip_loss=True, clip_cls_loss=0.03,
                           clip_replay_cls_loss=0.003, prototype_loss=False, classifier_loss=False, bic=False,
                           enhance_error=False, warm_layers=[])
        res, gc, gr = run(params, 1, False)
        for k in ('cls_bg_loss', 'cls_fg_loss', 'reg_loss', 'dist_cls_loss', 'dist_reg_loss'):
            d['distill_%s_%s' % (name, k)] = res[k].detach().numpy()
        d['distill_%s_grad_cls' % name] = gc
        d['distill_%s_grad_reg' % name] = gr
        print('distill', name, {k: float(v) for k, v in res.items() if k in WEIGHTS})

    # ---- replay batch with enhance_error: IL_Loss.forward :566-603 ----
    for method in ('L1', 'L2', 'L3'):
        params = mg.Params([0, P], distill=False, clip_loss=True, clip_cls_loss=0.03, clip_replay_cls_loss=0.003,
                           prototype_loss=False, classifier_loss=False, bic=False, enhance_error=True,
                           enhance_error_method=method.lower(), warm_layers=[])
        res, gc, gr = run(params, 1, True)
        for k in ('cls_bg_loss', 'cls_fg_loss', 'reg_loss', 'enhance_loss'):
            d['replay_%s_%s' % (method, k)] = res[k].detach().numpy()
        d['replay_%s_grad_cls' % method] = gc
        d['replay_%s_grad_reg' % method] = gr
        print('replay', method, {k: float(v) for k, v in res.items() if k in WEIGHTS})

    # ---- Weight_similarity.forward (weight_init.py:82-115) ----
    probs_ws = rng.uniform(0, 0.45, (2, A, C)).astype(np.float32)       # row sums straddle the 0.5 threshold

    def ws_model(img, return_feat=False, return_anchor=True, enable_act=True):
        return torch.from_numpy(probs_ws), None, anchors

    ws = Weight_similarity(ws_model, C - P, P)
    sc, lab = ws.forward(torch.zeros(2, 3, h, w), torch.from_numpy(ann))
    d.update(ws_probs=probs_ws, ws_scores=sc.numpy(), ws_labels=lab.numpy())
    ann_empty = np.full_like(ann, -1.0)
    assert ws.forward(torch.zeros(2, 3, h, w), torch.from_numpy(ann_empty)) is None
    print('weight_similarity', sc.shape, lab.shape)
    np.savez_compressed(os.path.join(OUT, 'f2_il_loss_reference.npz'), **d)


if __name__ == '__main__':
    main()
