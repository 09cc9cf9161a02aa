"""B200-native detection-head path for the incremental-learning RetinaNet of EonianCoda/CL_object_detection.

Host side mirrors the reference's Python interfaces (Anchors, calc_iou, FocalLoss, BBoxTransform, ClipBoxes,
predict); the arithmetic runs in hand-written sm_100a CUDA kernels behind the C ABI in include/cldet.h.
"""
from ._lib import CldetError, LIB_PATH, load as load_library  # noqa: F401
from ._ops import OPS_PATH, load as load_ops  # noqa: F401
from .params import HeadParams  # noqa: F401
from .anchors import Anchors, generate_anchors, num_anchors  # noqa: F401
from .losses import FocalLoss, calc_iou, iou_assign  # noqa: F401
from .dist import ShardedFocalLoss, gather_terms, shard_sizes, shard_slice  # noqa: F401
from . import detect  # noqa: F401
from .distill import enhance_error, head_distillation  # noqa: F401
from .pseudo_label import collate_annotations, filter_pseudo_labels, merge_pseudo_labels  # noqa: F401
from .matching import OutputNorm, WeightSimilarity, get_positive, match_anchors  # noqa: F401
from .detect import BBoxTransform, ClipBoxes, batched_nms, nms, detect_batch, detect_batch_head, predict, labeler_predict, coco_results  # noqa: F401

__all__ = ['enhance_error', 'WeightSimilarity', 'collate_annotations', 'filter_pseudo_labels', 'merge_pseudo_labels', 'coco_results', 'head_distillation', 'OutputNorm', 'get_positive', 'match_anchors', 'ShardedFocalLoss', 'gather_terms', 'shard_sizes', 'shard_slice', 'BBoxTransform', 'ClipBoxes', 'batched_nms', 'nms',
           'detect_batch', 'detect_batch_head', 'predict', 'labeler_predict', 'detect', 'Anchors', 'generate_anchors', 'num_anchors', 'FocalLoss', 'calc_iou', 'iou_assign', 'HeadParams',
           'CldetError', 'load_library', 'LIB_PATH', 'load_ops', 'OPS_PATH']
