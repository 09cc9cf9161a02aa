"""Drop-in for retinanet/anchors.py Anchors (reference :6-40): same constructor defaults, same
forward(image) -> [1, A, 4] float32 CUDA tensor, but generated ON the device by the K1 kernel and cached
per (H, W, device) -- the reference recomputes it with numpy on the host and uploads it on every forward.
"""
import ctypes
import threading
import weakref

import torch
import torch.nn as nn

from . import _lib


# anchors tensors produced here, by storage address: lets FocalLoss know that an `anchors` argument is the standard grid of
# an (H, W) image, which unlocks the GT-centric assignment kernel (include/cldet.h, cldet_loss_params.image_height)
_GRIDS = {}
_GRIDS_LOCK = threading.Lock()


def _register_grid(t, height, width):
    with _GRIDS_LOCK:
        if len(_GRIDS) > 64:
            for k in [k for k, (ref, _, _) in _GRIDS.items() if ref() is None]:
                del _GRIDS[k]
        _GRIDS[t.data_ptr()] = (weakref.ref(t), int(height), int(width))


def grid_of(anchors):
    """(height, width) if `anchors` is (a view of) a live tensor made by generate_anchors / Anchors, else None."""
    hit = _GRIDS.get(anchors.data_ptr())
    if hit is None:
        return None
    ref, h, w = hit
    t = ref()
    if t is None or t.numel() != anchors.numel() or t.device != anchors.device:
        return None
    return h, w


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
    _register_grid(out, height, width)
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
