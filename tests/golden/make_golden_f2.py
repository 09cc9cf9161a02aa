"""Golden vectors for SURVEY 8(f) row f2 (head-distillation terms of IL_Loss, retinanet/losses.py:705-737).

IL_Loss.forward cannot be driven stand-alone (it needs the trainer, two full models and the dataset), so this script
evaluates the SAME torch calls those lines make -- nn.Sigmoid, boolean-mask indexing, nn.SmoothL1Loss(), nn.MSELoss(),
autograd -- on seeded synthetic head outputs.  Build container only.
"""
import os

import numpy as np
import torch
import torch.nn as nn

OUT = os.path.dirname(os.path.abspath(__file__))


def il_loss_distill_terms(classification, regression, prev_classification, prev_regression, bg_masks, past_class_num,
                          distill_logits, ignore_GD):
    classifier_act, smooth_l1 = nn.Sigmoid(), nn.SmoothL1Loss()
    classification = classification[:, :, :past_class_num]
    if distill_logits:
        prev_fg_mask = classifier_act(prev_classification) > 0.05
    else:
        prev_classification = classifier_act(prev_classification)
        classification = classifier_act(classification)
        prev_fg_mask = prev_classification > 0.05
    reg_mask = torch.logical_and(bg_masks, prev_fg_mask.any(dim=2))
    dist_reg_loss = smooth_l1(prev_regression[reg_mask], regression[reg_mask])
    if ignore_GD:
        dist_class_loss = nn.MSELoss()(prev_classification[reg_mask], classification[reg_mask])
    else:
        dist_class_loss = nn.MSELoss()(prev_classification[prev_fg_mask], classification[prev_fg_mask])
    return dist_class_loss, dist_reg_loss


def main():
    torch.set_num_threads(2)
    rng = np.random.default_rng(1201)
    N, A, C, P = 3, 1161, 6, 4
    d = {}
    cls = rng.normal(-3.0, 2.0, (N, A, C)).astype(np.float32)
    prev = (cls[:, :, :P] + rng.normal(0, 0.7, (N, A, P))).astype(np.float32)
    reg = rng.normal(0, 1.0, (N, A, 4)).astype(np.float32)
    preg = (reg + rng.normal(0, 0.9, (N, A, 4))).astype(np.float32)
    bg = rng.uniform(0, 1, (N, A)) > 0.05
    d.update(cls=cls, prev=prev, reg=reg, preg=preg, bg=bg, P=P)
    for name, (dl, ig) in dict(probs=(False, False), logits=(True, False), probs_ignoregd=(False, True),
                               logits_ignoregd=(True, True)).items():
        tc = torch.from_numpy(cls).requires_grad_(True)
        tr = torch.from_numpy(reg).requires_grad_(True)
        lc, lr = il_loss_distill_terms(tc, tr, torch.from_numpy(prev), torch.from_numpy(preg), torch.from_numpy(bg), P, dl, ig)
        (0.6 * lc + 1.7 * lr).backward()
        d[name + '_cls_loss'] = lc.detach().numpy()
        d[name + '_reg_loss'] = lr.detach().numpy()
        d[name + '_grad_cls'] = tc.grad.numpy()
        d[name + '_grad_reg'] = tr.grad.numpy()
        print(name, float(lc), float(lr))
    np.savez_compressed(os.path.join(OUT, 'f2_distill.npz'), **d)


if __name__ == '__main__':
    main()
