#!/usr/bin/env python
"""Small driver for profiling the reference-mode eval boundary (predict: one image per call, no top-k) under ncu:
a few detect_batch calls on one COCO-shaped image with ~8 k (default) candidates."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import cl_object_detection_b200 as cld  # noqa: E402
from cl_object_detection_b200 import detect as D  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--mu', type=float, default=-9.5)
    ap.add_argument('--calls', type=int, default=3)
    a = ap.parse_args()
    dev = torch.device('cuda', 0)
    h, w, c = 800, 1333, 80
    anchors = cld.generate_anchors(h, w, dev)
    gen = torch.Generator(device=dev).manual_seed(int(-a.mu * 100))
    logits = torch.randn(1, anchors.shape[1], c, device=dev, generator=gen) * 2.0 + a.mu
    reg = torch.randn(1, anchors.shape[1], 4, device=dev, generator=gen) * 0.3
    for _ in range(a.calls):
        s, l, b = D.detect_batch(logits, reg, anchors, h, w)[0]
    torch.cuda.synchronize()
    print('candidates kept', int(s.shape[0]))


if __name__ == '__main__':
    main()
