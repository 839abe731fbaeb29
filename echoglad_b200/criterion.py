"""Drop-in criteria: same constructor kwargs and `.compute(pred_y, y, valid)` contract as the reference's
`WeightedBCEWithLogitsLoss` / `ExpectedLandmarkMSE` (src/core/criterion.py:30-34, 67-161), evaluated by
the fused sm_100a loss kernels (no numpy weight tensor, no host round trip)."""
from __future__ import annotations

from . import ops


class WeightedBCEWithLogitsLoss(object):
    """CRITERIA['WeightedBceWithLogits'] (src/builders/criterion_builder.py:11)."""

    def __init__(self, reduction, ones_weight, loss_weight, **_):
        if reduction != 'none':
            # the reference multiplies an elementwise weight into the loss, which only works for 'none'
            raise ValueError("WeightedBCEWithLogitsLoss expects reduction='none' (configs/default.yml:38)")
        self.ones_weight = ones_weight
        self.loss_weight = loss_weight

    def compute(self, pred_y, y, valid=None):
        if valid is None:
            raise AttributeError("'NoneType' object has no attribute 'view'")  # reference behaviour (criterion.py:15)
        with ops.device_of(pred_y):
            return ops.WeightedBCEWithLogits.apply(pred_y, y, valid, float(self.ones_weight), float(self.loss_weight))


class ExpectedLandmarkMSE(object):
    """CRITERIA['ExpectedLandmarkMse'] (src/builders/criterion_builder.py:12)."""

    def __init__(self, loss_weight=1, batch_size=2, frame_size=128, num_aux_graphs=6, use_main_graph_only=False,
                 num_output_channels=4):
        self.loss_weight = loss_weight
        self.batch_size = batch_size
        self.frame_size = frame_size
        self.num_aux_graphs = num_aux_graphs
        self.num_output_channels = num_output_channels
        self.use_main_graph_only = use_main_graph_only
        if use_main_graph_only:
            self.grid_sizes = [frame_size]
        else:
            self.grid_sizes = [2 ** k for k in range(1, num_aux_graphs + 1)] + [frame_size]

    def compute(self, pred_y, y, valid):
        with ops.device_of(pred_y):
            return ops.ExpectedLandmarkMSEFn.apply(pred_y, y, valid, int(self.batch_size),
                                                   int(self.num_output_channels), tuple(self.grid_sizes),
                                                   float(self.loss_weight))


class MAE(object):
    """CRITERIA['mae'] = the 'coordinate' loss of `use_coordinate_graph` (src/core/criterion.py:52-64,
    src/builders/criterion_builder.py:40-41): loss_weight * mean |pred - y| over the [4B, 2] coordinates
    (eg_mae: loss and gradient in one launch)."""

    def __init__(self, loss_weight=1):
        self.loss_weight = loss_weight

    def compute(self, pred_y, y):
        with ops.device_of(pred_y):
            return ops.MAELoss.apply(pred_y, y.to(pred_y.device), float(self.loss_weight))
