"""Loader of the torch custom-op layer `_cldet_torch.so` (csrc/cldet_torch.cpp): `torch.ops.cldet.focal_loss`,
`.detect`, `.batched_nms`.  The layer is host C++ over the C ABI of libcldet.so (include/cldet.h); it owns tensor
allocation, the workspace cache, the current stream and the autograd node of the fused loss.

There is NO fallback: if the library has not been built this raises.  Nothing here imports the test oracle.
"""
import os
import threading

import torch

from . import _lib

_PKG = os.path.dirname(os.path.abspath(__file__))
OPS_PATH = os.environ.get('CLDET_OPS_LIBRARY', os.path.join(_PKG, '_cldet_torch.so'))

_lock = threading.Lock()
_ops = None


def load():
    """`torch.ops.cldet`, loading `_cldet_torch.so` (and with it libcldet.so) on first use."""
    global _ops
    if _ops is not None:
        return _ops
    with _lock:
        if _ops is not None:
            return _ops
        _lib.load()          # same existence / ABI checks as the ctypes binding; also maps libcldet.so before the op layer
        if not os.path.exists(OPS_PATH):
            raise _lib.CldetError('%s is missing: build it with `python -m cl_object_detection_b200.build` '
                                  '(or __graft_entry__.build()). There is no CPU/PyTorch fallback for this path.' % OPS_PATH)
        torch.ops.load_library(OPS_PATH)
        ops = torch.ops.cldet
        if ops.abi_version() != 1:
            raise _lib.CldetError('_cldet_torch.so was built against a different libcldet ABI')
        _ops = ops
    return _ops
