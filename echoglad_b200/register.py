"""Plug-in point: registers the B200-native classes under the reference's own registry keys, so that
`model_builder.build` / `criterion_builder.build` (src/builders/model_builder.py:6-26,
src/builders/criterion_builder.py:6-35) construct them from an unmodified config.

    import src.builders.model_builder as mb, src.builders.criterion_builder as cb
    import echoglad_b200.register as reg
    reg.patch(mb.MODELS, cb.CRITERIA)
"""
from __future__ import annotations

from .criterion import MAE, ExpectedLandmarkMSE, WeightedBCEWithLogitsLoss
from .evaluator import LandmarkExpectedCoordiantesEvaluator
from .modules import HierarchicalPatchModel, UNETHierarchicalPatchModel

MODEL_KEYS = {
    'unet_hierarchical_patch': UNETHierarchicalPatchModel,
    'hierarchicalpatch': HierarchicalPatchModel,
}
CRITERION_KEYS = {
    'WeightedBceWithLogits': WeightedBCEWithLogitsLoss,
    'ExpectedLandmarkMse': ExpectedLandmarkMSE,
    'mae': MAE,  # the 'coordinate' loss (src/builders/criterion_builder.py:40-41)
}
EVALUATOR_KEYS = {
    'landmarkcoorderror': LandmarkExpectedCoordiantesEvaluator,  # src/builders/evaluator_builder.py:9
}


def patch(models: dict, criteria: dict, evaluators: dict = None) -> None:
    """Overwrites the hot-path entries of the reference registries in place; everything else is untouched."""
    models.update(MODEL_KEYS)
    criteria.update(CRITERION_KEYS)
    if evaluators is not None:
        evaluators.update(EVALUATOR_KEYS)
