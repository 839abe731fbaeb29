"""Boundary test (CPU, build container only): the drop-in classes bound into the reference's REAL registries.

`register.patch` is applied to `src.builders.model_builder.MODELS`, `criterion_builder.CRITERIA` and
`evaluator_builder.EVALUATORS` of the unmodified reference (imported through oracle/ref_shim.py), and the model /
criteria / evaluators are then constructed by the reference's own `build(config, logger)` functions from the parsed
`configs/default.yml`, with the keys the engine injects (src/engine.py:93-100, :140-160).  The module built this way
must accept, with `strict=True`, the `state_dict` of the module the UNPATCHED reference builds from the same config.
Skipped when /root/reference is absent (GPU box)."""
import copy
import os

import pytest
import torch

from oracle import ref_shim

pytestmark = pytest.mark.skipif(not ref_shim.reference_available(), reason="reference tree not present")


class _Logger:
    def infov(self, *a, **k):
        pass

    info = warning = error = infov


def _engine_model_config(cfg):
    """The kwargs injection of Engine._build (src/engine.py:93-100)."""
    mc = copy.deepcopy(cfg["model"])
    data = cfg["data"]
    mc["landmark"].update({"frame_size": data["transform"]["image_size"], "num_aux_graphs": data["num_aux_graphs"],
                           "use_coordinate_graph": data.get("use_coordinate_graph", False),
                           "use_connection_nodes": data.get("use_connection_nodes", False),
                           "use_main_graph_only": data.get("use_main_graph_only", False),
                           "num_output_channels": 4})
    return mc


@pytest.fixture()
def ref_builders():
    import importlib
    ref_shim.load()
    mb = importlib.import_module("src.builders.model_builder")
    cb = importlib.import_module("src.builders.criterion_builder")
    eb = importlib.import_module("src.builders.evaluator_builder")
    saved = (dict(mb.MODELS), dict(cb.CRITERIA), dict(eb.EVALUATORS))
    yield mb, cb, eb
    for d, s in zip((mb.MODELS, cb.CRITERIA, eb.EVALUATORS), saved):
        d.clear()
        d.update(s)


def _default_yml():
    import yaml
    with open(os.path.join(ref_shim.REFERENCE_ROOT, "configs", "default.yml")) as fh:
        return yaml.safe_load(fh)


def test_patched_reference_builders_construct_the_drop_in_classes(ref_builders):
    import echoglad_b200 as eg
    from echoglad_b200 import register
    mb, cb, eb = ref_builders
    cfg = _default_yml()
    torch.manual_seed(200)
    ref_models = mb.build(_engine_model_config(cfg), _Logger())           # the reference's own classes
    assert type(ref_models["landmark"]).__module__ == "src.core.models"
    ref_sd = ref_models["landmark"].state_dict()

    register.patch(mb.MODELS, cb.CRITERIA, eb.EVALUATORS)
    models = mb.build(_engine_model_config(cfg), _Logger())
    landmark = models["landmark"]
    assert isinstance(landmark, eg.UNETHierarchicalPatchModel)
    assert type(models["embedder"]).__module__ == "src.core.models"       # the embedder stays the reference's (PyTorch)
    # same keys, shapes and dtypes; strict load of the reference-built module's state_dict, and back
    ours = landmark.state_dict()
    assert list(ours.keys()) == list(ref_sd.keys())
    for k, v in ref_sd.items():
        assert ours[k].shape == v.shape and ours[k].dtype == v.dtype, k
    landmark.load_state_dict(ref_sd, strict=True)
    ref_models["landmark"].load_state_dict(landmark.state_dict(), strict=True)
    for k, v in landmark.state_dict().items():
        assert torch.equal(v, ref_sd[k]), k
    assert sum(p.numel() for p in landmark.parameters()) == sum(p.numel() for p in ref_models["landmark"].parameters())

    # criteria through criterion_builder.build with the keys the engine injects (src/engine.py:140-149)
    ccfg = copy.deepcopy(cfg["train"]["criterion"])
    ccfg.update({"batch_size": 2, "frame_size": 224, "num_aux_graphs": 7, "use_main_graph_only": False,
                 "use_coordinate_graph": True, "num_output_channels": 4})
    crit = cb.build(ccfg, _Logger())
    assert isinstance(crit["WeightedBceWithLogits"], eg.WeightedBCEWithLogitsLoss)
    assert isinstance(crit["ExpectedLandmarkMse"], eg.ExpectedLandmarkMSE)
    assert isinstance(crit["coordinate"], eg.MAE)
    assert crit["ExpectedLandmarkMse"].grid_sizes == [2, 4, 8, 16, 32, 64, 128, 224]
    assert crit["WeightedBceWithLogits"].ones_weight == 9000 and crit["ExpectedLandmarkMse"].loss_weight == 10

    # evaluator through evaluator_builder.build (src/engine.py:151-158)
    ecfg = {"standards": ["landmarkcoorderror"], "batch_size": 2, "frame_size": 224, "use_coordinate_graph": False}
    ev = eb.build(ecfg, _Logger())
    assert isinstance(ev["landmarkcoorderror"], eg.LandmarkExpectedCoordiantesEvaluator)


@pytest.mark.parametrize("name,extra", [
    ("hierarchicalpatch", dict(use_coordinate_graph=True)),
    ("hierarchicalpatch", dict(use_connection_nodes=True, gnn_jk_mode="max")),
    ("unet_hierarchical_patch", dict(use_main_graph_only=True)),
])
def test_other_registry_entries_and_flags_round_trip(ref_builders, name, extra):
    """The base `hierarchicalpatch` entry and the graph flags: same state_dict layout as the reference class."""
    from echoglad_b200 import register
    mb, cb, eb = ref_builders
    cfg = _default_yml()
    mc = _engine_model_config(cfg)
    mc["landmark"]["name"] = name
    mc["landmark"].update(extra)
    if name == "hierarchicalpatch":
        mc["landmark"].pop("encoder_embedding_widths")
        mc["landmark"].pop("encoder_embedding_dims")
        mc["landmark"]["frame_size"] = 16
        mc["landmark"]["num_aux_graphs"] = 3
    ref = mb.build(mc, _Logger())["landmark"]
    register.patch(mb.MODELS, cb.CRITERIA)
    ours = mb.build(mc, _Logger())["landmark"]
    assert type(ours).__module__.startswith("echoglad_b200")
    ours.load_state_dict(ref.state_dict(), strict=True)
    assert list(ours.state_dict().keys()) == list(ref.state_dict().keys())


def test_evaluator_visual_helpers_match_the_reference_class(ref_builders):
    """`get_softmaxed_heatmap` / `create_overlay_image`, which Engine.log_heatmap_wandb calls through the evaluator
    (src/engine.py:575), against the reference class on the same inputs (bit-exact: same torch ops)."""
    import sys
    import numpy as np
    import echoglad_b200 as eg
    _, _, eb = ref_builders
    plt = sys.modules["matplotlib.pyplot"]
    if not hasattr(plt, "get_cmap"):  # stubbed matplotlib of the build container
        plt.get_cmap = lambda name: None
    ref = eb.EVALUATORS["landmarkcoorderror"](None, 2, 16, False)
    ours = eg.LandmarkExpectedCoordiantesEvaluator(None, 2, 16, False)
    g = torch.Generator().manual_seed(3)
    y = torch.randn(2, 84 + 256, 4, generator=g)
    assert torch.equal(ref.get_softmaxed_heatmap(y), ours.get_softmaxed_heatmap(y))
    for x in (torch.rand(16, 16, generator=g), torch.randn(16, 16, generator=g)):
        h = torch.rand(16, 16, 4, generator=g)
        assert np.array_equal(np.asarray(ref.create_overlay_image(x, h)), np.asarray(ours.create_overlay_image(x, h)))
