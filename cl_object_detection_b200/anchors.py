"""Drop-in for retinanet/anchors.py Anchors (reference :6-40): same constructor defaults, same
forward(image) -> [1, A, 4] float32 CUDA tensor, but generated ON the device by the K1 kernel and cached
per (H, W, device) -- the reference recomputes it with numpy on the host and uploads it on every forward.
"""
import ctypes
import threading

import torch
import torch.nn as nn

from . import _lib


def num_anchors(height, width):
    out = ctypes.c_int64(0)
    _lib.check(_lib.load().cldet_num_anchors(int(height), int(width), ctypes.byref(out)))
    return out.value


def generate_anchors(height, width, device=None):
    """[1, A, 4] fp32 anchors for an image of (height, width), bit-exact with the reference."""
    device = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
    if device.type != 'cuda':
        raise RuntimeError('cl_object_detection_b200 runs on CUDA devices only (no CPU path)')
    lib = _lib.load()
    a = num_anchors(height, width)
    with torch.cuda.device(device):
        out = torch.empty((1, a, 4), dtype=torch.float32, device=device)
        _lib.check(lib.cldet_anchors(int(height), int(width), out.data_ptr(), torch.cuda.current_stream().cuda_stream))
    return out


class Anchors(nn.Module):
    """Same interface as the reference module; only the default pyramid (levels 3-7) is supported, which is
    the only configuration the reference ever constructs (model.py:304)."""

    def __init__(self, pyramid_levels=None, strides=None, sizes=None, ratios=None, scales=None):
        super().__init__()
        if any(x is not None for x in (pyramid_levels, strides, sizes, ratios, scales)):
            raise NotImplementedError('only the reference default anchor configuration is implemented')
        self._cache = {}
        self._lock = threading.Lock()

    def forward(self, image):
        h, w = int(image.shape[2]), int(image.shape[3])
        dev = image.device if image.is_cuda else torch.device('cuda', torch.cuda.current_device())
        key = (h, w, dev.index)
        with self._lock:
            hit = self._cache.get(key)
        if hit is not None:
            return hit
        out = generate_anchors(h, w, dev)
        with self._lock:
            self._cache[key] = out
        return out
