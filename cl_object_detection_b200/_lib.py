"""ctypes binding of libcldet.so (the C ABI declared in include/cldet.h).

There is NO fallback: if the library is missing or a call fails, this module raises.  Nothing here
imports the test oracle.
"""
import ctypes
import os
import threading

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('CLDET_LIBRARY', os.path.join(_PKG, 'libcldet.so'))   # override: A/B builds only

_lock = threading.Lock()
_lib = None


class CldetError(RuntimeError):
    pass


class PeerExchange(ctypes.Structure):
    """struct cldet_peer_exchange (include/cldet.h)."""
    _fields_ = [('d_peer_terms', ctypes.c_void_p), ('d_peer_flags', ctypes.c_void_p), ('rank', ctypes.c_int32),
                ('world', ctypes.c_int32), ('parity', ctypes.c_int32), ('timeout_ms', ctypes.c_int32),
                ('target_arrivals', ctypes.c_uint32), ('reserved', ctypes.c_int32), ('d_flags_local', ctypes.c_void_p),
                ('d_terms_local', ctypes.c_void_p), ('d_wait_out', ctypes.c_void_p), ('d_wait_status', ctypes.c_void_p)]


class LossParams(ctypes.Structure):
    """struct cldet_loss_params (include/cldet.h)."""
    _fields_ = [('alpha', ctypes.c_float), ('gamma', ctypes.c_float), ('incremental', ctypes.c_int32),
                ('past_class_num', ctypes.c_int32), ('ignore_past_class', ctypes.c_int32),
                ('new_ignore_past_class', ctypes.c_int32), ('decrease_positive_by_iou', ctypes.c_int32),
                ('enhance_on_new', ctypes.c_int32), ('decrease_positive', ctypes.c_float), ('image_height', ctypes.c_int32),
                ('image_width', ctypes.c_int32), ('cls_is_logits', ctypes.c_int32)]


_P = ctypes.c_void_p
_I = ctypes.c_int
_L = ctypes.c_int64
_F = ctypes.c_float
_Z = ctypes.c_size_t

# name -> (restype, argtypes).  Must list every function include/cldet.h declares (tests check this).
SIGNATURES = {
    'cldet_abi_version': (_I, []),
    'cldet_status_string': (ctypes.c_char_p, [_I]),
    'cldet_last_cuda_error': (ctypes.c_char_p, []),
    'cldet_num_anchors': (_I, [_I, _I, ctypes.POINTER(_L)]),
    'cldet_anchors': (_I, [_I, _I, _P, _P]),
    'cldet_iou_assign': (_I, [_P, _L, _P, _I, _I, _I, _P, _P, _P, _P, _P, _P]),
    'cldet_iou_max_f64': (_I, [_P, _L, _P, _I, _P, _P, _P]),
    'cldet_calc_iou': (_I, [_P, _L, _P, _I, _P, _P]),
    'cldet_focal_loss_workspace_bytes': (_Z, [_I, _L]),
    'cldet_focal_loss': (_I, [_P, _P, _P, _P, _I, _L, _I, _I, ctypes.POINTER(LossParams), _P, _P, _P, _P, _P, _P, _P, _P,
                              _P, _P, _P, _P, _Z, _P]),
    'cldet_focal_loss_sharded': (_I, [_P, _P, _P, _P, _I, _L, _I, _I, ctypes.POINTER(LossParams), _P, _P, _P, _P, _P, _P, _P,
                                      _P, _P, _P, _P, _P, _Z, ctypes.POINTER(PeerExchange), _P, _P]),
    'cldet_peer_alloc': (_I, [_Z, ctypes.POINTER(ctypes.c_void_p), ctypes.c_char_p]),
    'cldet_peer_open': (_I, [ctypes.c_char_p, ctypes.POINTER(ctypes.c_void_p)]),
    'cldet_peer_close': (_I, [_P]),
    'cldet_peer_free': (_I, [_P]),
    'cldet_enable_peer_access': (_I, [_I]),
    'cldet_peer_wait': (_I, [_P, _P, _I, _I, _I, ctypes.c_uint32, _I, _P, _P, _P, _P]),
    'cldet_focal_loss_profile_events': (_I, [_P, _P, _P]),
    'cldet_focal_loss_from_assignment': (_I, [_P, _P, _P, _P, _I, _L, _I, _I, ctypes.POINTER(LossParams), _P, _P, _P, _P,
                                              _P, _P, _P, _P, _P, _P, _P, _Z, _P]),
    'cldet_focal_loss_reweight_rows': (_I, [_P, _P, _P, _P, _I, _L, _I, _I, ctypes.POINTER(LossParams), _P, _L, _P, _L, _P,
                                            _L, _P, _L, _P, _F, _P, _P, _P, _P, _P, _P, _P, _Z, _P]),
    'cldet_focal_loss_reweight': (_I, [_P, _P, _P, _P, _I, _L, _I, _I, ctypes.POINTER(LossParams), _P, _P, _P, _P, _P,
                                       _P, _P, _P, _Z, _P]),
    'cldet_focal_loss_head': (_I, [_P, _P, _I, _I, _I, _P, _P, _I, _I, _I, ctypes.POINTER(LossParams), _P, _P, _P, _P, _P, _P,
                                   _P, _P, _P, _P, _P, _P, _Z, _P]),
    'cldet_focal_loss_head_reweight': (_I, [_P, _P, _I, _I, _I, _P, _P, _I, _I, _I, ctypes.POINTER(LossParams), _P, _L, _P, _L,
                                            _P, _L, _P, _L, _P, _P, _P, _P, _P, _P, _P, _Z, _P]),
    'cldet_distill_workspace_bytes': (_Z, [_I, _L]),
    'cldet_distill_forward': (_I, [_P, _P, _P, _P, _P, _I, _L, _I, _I, _I, _I, _P, _P, _P, _Z, _P]),
    'cldet_distill_backward': (_I, [_P, _P, _P, _P, _P, _I, _L, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P]),
    'cldet_enhance_error_workspace_bytes': (_Z, [_L]),
    'cldet_enhance_error_forward': (_I, [_P, _I, _L, _I, _I, _I, _P, _P, _P, _Z, _P]),
    'cldet_enhance_error_backward': (_I, [_P, _I, _L, _I, _I, _I, _P, _P, _P, _P]),
    'cldet_masked_abs_mean_forward': (_I, [_P, _P, _I, _L, _P, _P, _P]),
    'cldet_masked_abs_mean_backward': (_I, [_P, _P, _I, _L, _P, _P, _L, _P, _P]),
    'cldet_decode_boxes': (_I, [_P, _P, _I, _L, _I, _I, _I, _P, _P]),
    'cldet_clip_boxes': (_I, [_P, _L, _I, _I, _P]),
    'cldet_decode_filter': (_I, [_P, _I, _P, _P, _I, _L, _I, _I, _I, _F, _P, _P, _L, _P, _P]),
    'cldet_decode_filter_head': (_I, [_P, _P, _I, _I, _I, _I, _P, _I, _I, _F, _P, _P, _L, _P, _P]),
    'cldet_sort_workspace_bytes': (_Z, [_I, _L, _I]),
    'cldet_sort_candidates': (_I, [_P, _P, _P, _I, _L, _L, _I, _P, _L, _P, _P, _Z, _P]),
    'cldet_nms_workspace_bytes': (_Z, [_I, _L]),
    'cldet_nms_sorted': (_I, [_P, _P, _I, _L, _L, _F, _I, _L, _P, _P, _P, _Z, _P]),
    'cldet_nms_gather_sorted': (_I, [_P, _P, _I, _L, _L, _F, _I, _L, _P, _P, _P, _P, _P, _P, _Z, _P]),
    'cldet_batched_nms_workspace_bytes': (_Z, [_L]),
    'cldet_batched_nms': (_I, [_P, _P, _P, _L, _F, _I, _L, _P, _P, _P, _Z, _P]),
    'cldet_coco_results': (_I, [_P, _P, _P, _P, _P, _I, _L, _F, _P, _P, _P]),
    'cldet_gather_detections': (_I, [_P, _P, _P, _I, _L, _L, _P, _P, _P, _P]),
}


# entry points declared in the header whose kernels are not in this build yet (emptied as they land)
_PENDING = set()


def load():
    """Load libcldet.so once.  Raises CldetError if it has not been built (no fallback path exists)."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise CldetError('%s is missing: build it with `python -m cl_object_detection_b200.build` '
                             '(or __graft_entry__.build()). There is no CPU/PyTorch fallback for this path.' % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            try:
                fn = getattr(lib, name)
            except AttributeError as e:
                if name in _PENDING:
                    continue
                raise CldetError('libcldet.so does not export %s (stale build?)' % name) from e
            fn.restype = res
            fn.argtypes = args
        if lib.cldet_abi_version() != 1:
            raise CldetError('libcldet.so ABI version mismatch')
        _lib = lib
    return _lib


def check(status):
    if status == 0:
        return
    lib = load()
    msg = lib.cldet_status_string(status).decode()
    if status == 3:
        msg += ': ' + lib.cldet_last_cuda_error().decode()
    raise CldetError('libcldet: ' + msg)


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def ptr_array(tensors):
    """Host array of device pointers (`const float* const*` in the C ABI) for a list of tensors; None -> NULL."""
    if tensors is None:
        return None
    return (ctypes.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])
