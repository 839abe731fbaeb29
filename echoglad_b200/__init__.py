"""echoglad_b200 — B200-native (sm_100a) EchoGLAD GNN hot path behind the reference's module API.

Importing the package loads libechoglad_b200.so; there is no CPU or eager fallback.
"""
from ._lib import EchogladError, lib  # noqa: F401  (fails loudly when the CUDA library is missing)
from .criterion import MAE, ExpectedLandmarkMSE, WeightedBCEWithLogitsLoss  # noqa: F401
from .evaluator import LandmarkExpectedCoordiantesEvaluator  # noqa: F401
from .graph import DeviceGraph, GraphMeta, HierGraphSpec  # noqa: F401
from .modules import CNN, HierarchicalPatchModel, UNETHierarchicalPatchModel  # noqa: F401

__version__ = "0.1.0"
