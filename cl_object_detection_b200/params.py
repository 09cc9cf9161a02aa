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


def to_loss_params(params, cur_state, num_classes):
    """Translate the reference's params object into struct cldet_loss_params, with the reference's error behaviour."""
    alpha, gamma = params['alpha'], params['gamma']
    if alpha is None or gamma is None:
        raise TypeError("params['alpha'] / params['gamma'] must be set (losses.py:254-255)")
    incremental = cur_state > 0
    lp = LossParams()
    lp.alpha, lp.gamma = float(alpha), float(gamma)
    lp.incremental = int(incremental)
    lp.decrease_positive = 1.0
    if incremental:
        past = int(params.states[cur_state]['num_past_class'])
        if past < 0 or past > num_classes:
            raise IndexError('num_past_class %d outside [0, %d]' % (past, num_classes))
        lp.past_class_num = past
        lp.ignore_past_class = int(bool(params['ignore_past_class']))
        lp.new_ignore_past_class = int(bool(params['ignore_past_class']) and bool(params['new_ignore_past_class']))
        lp.decrease_positive_by_iou = int(bool(params['decrease_positive_by_IOU']))
        lp.enhance_on_new = int(bool(params['enhance_on_new']))
        if not lp.decrease_positive_by_iou:
            dp = params['decrease_positive']
            if dp is None:   # the reference evaluates `None - Tensor` here (losses.py:366)
                raise TypeError("params['decrease_positive'] must be set in incremental states (losses.py:365-366)")
            lp.decrease_positive = float(dp)
    return lp
