"""Host-side mirror of the pieces of preprocessing/params.py the head path reads.

FocalLoss.forward reads `params[key]` (None when absent, params.py:174-178) and
`params.states[cur_state]['num_past_class']` (losses.py:254-255, 264-266, 317-323, 353, 365, 388).
Any object with that duck type -- including the reference's own Params -- is accepted by the drop-in;
HeadParams is a minimal one carrying the CLI defaults of main.py:116-177.
"""
from ._lib import LossParams


class HeadParams:
    DEFAULTS = dict(alpha=0.25, gamma=2.0, distill=False, enhance_on_new=False, ignore_past_class=False,
                    new_ignore_past_class=False, decrease_positive_by_IOU=False, decrease_positive=1.0,
                    persuado_label=False, clip_loss=True, clip_cls_loss=0.03, clip_replay_cls_loss=0.003)

    def __init__(self, num_past_class=(0,), **overrides):
        self._d = dict(self.DEFAULTS)
        self._d.update(overrides)
        self.states = [{'num_past_class': int(n)} for n in num_past_class]

    def __getitem__(self, key):
        return self._d.get(key, None)

    def __setitem__(self, key, value):
        self._d[key] = value


def loss_param_args(params, cur_state, num_classes):
    """The reference's params object as the scalar arguments of the loss entry points, with the reference's error behaviour:
    (alpha, gamma, incremental, past_class_num, ignore_past_class, new_ignore_past_class, decrease_positive_by_iou,
    enhance_on_new, decrease_positive) -- the order of struct cldet_loss_params / torch.ops.cldet.focal_loss."""
    alpha, gamma = params['alpha'], params['gamma']
    if alpha is None or gamma is None:
        raise TypeError("params['alpha'] / params['gamma'] must be set (losses.py:254-255)")
    if cur_state <= 0:
        return float(alpha), float(gamma), 0, 0, 0, 0, 0, 0, 1.0
    past = int(params.states[cur_state]['num_past_class'])
    if past < 0 or past > num_classes:
        raise IndexError('num_past_class %d outside [0, %d]' % (past, num_classes))
    ignore = bool(params['ignore_past_class'])
    by_iou = bool(params['decrease_positive_by_IOU'])
    dec = 1.0
    if not by_iou:
        dp = params['decrease_positive']
        if dp is None:   # the reference evaluates `None - Tensor` here (losses.py:366)
            raise TypeError("params['decrease_positive'] must be set in incremental states (losses.py:365-366)")
        dec = float(dp)
    return (float(alpha), float(gamma), 1, past, int(ignore), int(ignore and bool(params['new_ignore_past_class'])), int(by_iou),
            int(bool(params['enhance_on_new'])), dec)


def to_loss_params(params, cur_state, num_classes):
    """Translate the reference's params object into struct cldet_loss_params (ctypes), with the reference's error behaviour."""
    v = loss_param_args(params, cur_state, num_classes)
    lp = LossParams()
    (lp.alpha, lp.gamma, lp.incremental, lp.past_class_num, lp.ignore_past_class, lp.new_ignore_past_class,
     lp.decrease_positive_by_iou, lp.enhance_on_new, lp.decrease_positive) = v
    return lp
