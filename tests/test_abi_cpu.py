"""CPU tier: the C-ABI library loads and exports every symbol include/echoglad_b200.h declares; host-side
closed forms (graph, node types) are bit-exact against the golden vectors; host logic of the drop-in
modules (constructor kwargs, state_dict layout, registry patching) — no GPU compute calls."""
import ctypes
import hashlib
import json
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="module")
def eg():
    sys.path.insert(0, ROOT)
    import __graft_entry__
    __graft_entry__.build()
    import echoglad_b200
    return echoglad_b200


def test_every_declared_symbol_is_exported_and_bound(eg):
    header = open(os.path.join(ROOT, "include", "echoglad_b200.h")).read()
    declared = set(re.findall(r"\b(eg_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 30
    from echoglad_b200 import _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert eg.lib.eg_workspace_bytes() > 0
    assert b"sm_100a" in eg.lib.eg_version()


def test_missing_library_fails_loudly():
    code = ("import echoglad_b200._lib as l, os, sys\n")
    env = dict(os.environ, PYTHONPATH=ROOT)
    src = ("import importlib, sys, os\n"
           "import echoglad_b200._lib as L\n"
           "L.LIB_PATH = '/nonexistent/libechoglad_b200.so'\n"
           "try:\n    L._load()\nexcept ImportError as e:\n    print('IMPORTERROR', e)\n")
    out = subprocess.run([sys.executable, "-c", src], capture_output=True, text=True, env=env, cwd=ROOT)
    assert "IMPORTERROR" in out.stdout and "no CPU fallback" in out.stdout, out.stdout + out.stderr


def _spec(eg, key):
    s, n, mo, co, cn, mt, at = key.split("_")
    return eg.HierGraphSpec(frame_size=int(s[1:]), num_aux_graphs=max(int(n[1:]), 1),
                            use_main_graph_only=mo == "mo1", use_coordinate_graph=co == "co1",
                            use_connection_nodes=cn == "cn1", main_graph_type=mt, aux_graph_type=at)


def test_host_closed_form_graph_bit_exact(eg):
    z = np.load(os.path.join(GOLDEN, "graphs_small.npz"))
    keys = sorted({k.split("/")[0] for k in z.files})
    for k in keys:
        spec = _spec(eg, k)
        ei = spec.host_edge_index(1).numpy()
        assert np.array_equal(ei, z[k + "/edge_index"].astype(np.int64)), k
        assert np.array_equal(spec.host_node_type(1), z[k + "/node_type"].astype(np.float64)), k
        meta = spec.info()
        assert meta.num_edges == ei.shape[1] and meta.num_nodes == z[k + "/node_type"].shape[0]
    # batched = block-diagonal offsets (PyG collate)
    spec = _spec(eg, "S12_n3_mo0_co0_cn0_grid_grid")
    one, three = spec.host_edge_index(1), spec.host_edge_index(3)
    n = spec.info().num_nodes
    assert torch.equal(three, torch.cat([one + b * n for b in range(3)], dim=1))


def test_host_closed_form_graph_hashes_224(eg):
    h = json.load(open(os.path.join(GOLDEN, "graph_hashes.json")))
    for k, v in h["specs"].items():
        spec = _spec(eg, k)
        ei = spec.host_edge_index(1).numpy()
        assert ei.shape[1] == v["num_edges"]
        assert hashlib.sha256(ei.tobytes()).hexdigest() == v["edge_index_sha256"], k
        assert hashlib.sha256(spec.host_node_type(1).tobytes()).hexdigest() == v["node_type_sha256"], k
    meta = eg.HierGraphSpec().info()
    assert (meta.num_nodes, meta.num_edges, meta.max_degree, meta.crop_offset) == (72020, 430200, 10, 8)
    assert meta.level_offset == (0, 4, 20, 84, 340, 1364, 5460, 21844)
    big = eg.HierGraphSpec(frame_size=448, num_aux_graphs=8).info()
    assert (big.num_nodes, big.num_edges) == (288084, 1724664)


@pytest.mark.parametrize("kw,classes", [
    (dict(), (408, 155)),                                     # default.yml: 392 main + 16 childless aux tiles; 155 tiles with children / ragged
    (dict(use_main_graph_only=True), (392, 0)),
    (dict(frame_size=448, num_aux_graphs=8), (1688, 563)),  # BASELINE configs[3]
    (dict(use_coordinate_graph=True), (408, 155)),
    (dict(use_connection_nodes=True), None),
    (dict(frame_size=64, num_aux_graphs=5), None),
    (dict(frame_size=32, num_aux_graphs=4, main_graph_type="grid-diagonal", aux_graph_type="grid-diagonal"), None),
    (dict(frame_size=12, num_aux_graphs=3), None),
])
def test_gather_plan_is_consistent_and_tiles_land_in_their_kernel_class(eg, kw, classes):
    """Host-only self check of the tile table + gather plan of the fused kernel against the closed-form neighbour
    lists and gcn_norm weights, and the class every tile runs in (lattice / general): every main-level tile must be a
    lattice tile (a general tile costs twice the time)."""
    from echoglad_b200._lib import lib
    spec = eg.HierGraphSpec(**kw)
    stats = (ctypes.c_int64 * 8)()
    assert lib.eg_graph_plan_check(ctypes.byref(spec.c_spec()), stats) == 0
    tiles, plan_rows, csr_rows = stats[0], stats[1], stats[2]
    assert plan_rows + csr_rows == spec.info().num_nodes
    assert stats[6] + stats[7] == tiles
    if classes is not None:
        assert (stats[6], stats[7]) == classes


@pytest.mark.parametrize("kw,want", [
    # (usable, plain patches, patches with children by direct loads, CSR tiles, patches reading their unit's pool, units)
    (dict(), (1, 408, 42, 1, 112, 171)),                      # default.yml: 112 families (aux-128 patch + its main patches)
    (dict(use_main_graph_only=True), (1, 392, 0, 0, 0, 392)),
    (dict(frame_size=448, num_aux_graphs=8), (1, 1688, 170, 1, 392, 683)),
    (dict(use_coordinate_graph=True), (1, 408, 42, 1, 112, 171)),
    (dict(use_connection_nodes=True), (0, 0, 0, 0, 0, 0)),    # hubs: the gather plan runs instead
    (dict(frame_size=64, num_aux_graphs=5), None),
    (dict(frame_size=32, num_aux_graphs=4, main_graph_type="grid-diagonal", aux_graph_type="grid-diagonal"), (0, 0, 0, 0, 0, 0)),
    (dict(frame_size=12, num_aux_graphs=3), None),
    (dict(frame_size=96, num_aux_graphs=6), None),
])
def test_patch_plan_reproduces_the_csr(eg, kw, want):
    """Host-only replay of the TMA path of the fused kernel: box position -> node id -> weight for every node of every
    patch tile against the closed-form neighbour list and gcn_norm weights (bit-equal floats)."""
    from echoglad_b200._lib import lib
    spec = eg.HierGraphSpec(**kw)
    stats = (ctypes.c_int64 * 6)()
    assert lib.eg_graph_patch_check(ctypes.byref(spec.c_spec()), stats) == 0
    if want is not None:
        assert tuple(stats) == want
    if stats[0]:
        full = (ctypes.c_int64 * 8)()
        lib.eg_graph_plan_check(ctypes.byref(spec.c_spec()), full)
        assert stats[1] + stats[2] + stats[3] + stats[4] == full[0]


def test_malformed_spec_is_rejected(eg):
    with pytest.raises(eg.EchogladError, match="malformed"):
        eg.HierGraphSpec(frame_size=448, num_aux_graphs=7).info()  # 2^7 < 448/2
    with pytest.raises(eg.EchogladError):
        eg.HierGraphSpec(frame_size=1, num_aux_graphs=1).info()
    with pytest.raises(ValueError):
        eg.HierGraphSpec(main_graph_type="hex").info()


DEFAULT_KW = dict(frame_size=224, gnn_dropout_p=0.5, classifier_dropout_p=0.5, node_embedding_dim=128,
                  node_hidden_dim=128, num_output_channels=4, num_gnn_layers=3, num_aux_graphs=7,
                  gnn_jk_mode='last', classifier_hidden_dim=32, residual=True, use_coordinate_graph=False,
                  output_activation='logit', use_connection_nodes=False, use_main_graph_only=False)


def test_state_dict_layout_is_the_references(eg):
    from oracle import restated as R
    m = eg.UNETHierarchicalPatchModel(encoder_embedding_widths=[128, 64, 32, 16, 8, 4, 2],
                                      encoder_embedding_dims=[8, 16, 32, 64, 128, 256, 512], **DEFAULT_KW)
    sd = m.state_dict()
    assert sum(p.numel() for p in m.parameters()) == 8073948
    assert sd["gnn_layers.0.module_0.lin.weight"].shape == (128, 128)
    assert sd["gnn_layers.2.module_1.num_batches_tracked"].dtype == torch.int64
    assert sd["node_classifiers.3.8.weight"].shape == (1, 16)
    assert sd["linears.7.weight"].shape == (128, 4, 1, 1)
    ref_layout = R.init_landmark_state(R.Cfg())  # loaded strict=True into the REFERENCE class by make_golden.py
    assert set(sd) == set(ref_layout)
    for k in sd:
        assert sd[k].shape == ref_layout[k].shape, k
    m.load_state_dict(ref_layout, strict=True)
    base = eg.HierarchicalPatchModel(**DEFAULT_KW)
    assert set(base.state_dict()) == {k for k in sd if k.startswith(("gnn_layers", "node_classifiers"))}
    emb = eg.CNN(out_channels=[4], kernel_sizes=[3], pool_sizes=[1], cnn_dropout_p=0.1)
    emb.load_state_dict(R.init_embedder_state(4), strict=True)
    assert sum(p.numel() for p in emb.parameters()) == 56


def test_unsupported_configs_fail_loudly_and_cpu_inputs_are_rejected(eg):
    kw = dict(DEFAULT_KW)
    # the reference constructor defaults (64 / 16, src/core/models.py:290-296) build the generic-width path
    small = eg.HierarchicalPatchModel(**{**kw, "node_hidden_dim": 64, "classifier_hidden_dim": 16})
    assert small.state_dict()["gnn_layers.0.module_0.lin.weight"].shape == (64, 128)
    assert small.state_dict()["node_classifiers.3.4.weight"].shape == (8, 16)
    with pytest.raises(NotImplementedError):
        eg.HierarchicalPatchModel(**{**kw, "node_hidden_dim": 96})
    with pytest.raises(NotImplementedError):
        eg.HierarchicalPatchModel(**{**kw, "node_hidden_dim": 64, "use_coordinate_graph": True})
    coord = eg.HierarchicalPatchModel(**{**kw, "use_coordinate_graph": True})  # optional branch: 3 coordinate MLPs
    assert [k for k in coord.state_dict() if k.startswith("node_coordinate_mlp.2.8.")] == \
        ["node_coordinate_mlp.2.8.weight", "node_coordinate_mlp.2.8.bias"]
    assert coord.state_dict()["node_coordinate_mlp.0.0.weight"].shape == (32, 136)
    with pytest.raises(TypeError):
        eg.HierarchicalPatchModel(**{**kw, "output_activation": "tanh"})
    with pytest.raises(AssertionError):
        eg.HierarchicalPatchModel(**{**kw, "gnn_jk_mode": "mean"})
    m = eg.HierarchicalPatchModel(**{**kw, "frame_size": 12, "num_aux_graphs": 3})
    with pytest.raises(eg.EchogladError, match="no CPU fallback"):
        m(x=torch.randn(1, 128, 12, 12))
    bce = eg.WeightedBCEWithLogitsLoss(reduction='none', ones_weight=9000, loss_weight=1)
    with pytest.raises(eg.EchogladError):
        bce.compute(torch.zeros(1, 4, 4), torch.zeros(1, 4, 4), torch.ones(4, 4))


def test_registry_patch_replaces_only_hot_path_entries(eg):
    from echoglad_b200 import register
    models = {'cnn': object, 'unet_hierarchical_patch': object, 'hierarchicalpatch': object, 'unet': object}
    criteria = {'mse': object, 'WeightedBceWithLogits': object, 'ExpectedLandmarkMse': object}
    register.patch(models, criteria)
    assert models['unet_hierarchical_patch'] is eg.UNETHierarchicalPatchModel
    assert models['hierarchicalpatch'] is eg.HierarchicalPatchModel
    assert models['cnn'] is object and models['unet'] is object
    assert criteria['WeightedBceWithLogits'] is eg.WeightedBCEWithLogitsLoss
    assert criteria['ExpectedLandmarkMse'] is eg.ExpectedLandmarkMSE and criteria['mse'] is object
    # builder-style construction: MODELS[name](**config['landmark']) with the engine-injected keys
    cfg = dict(encoder_embedding_widths=[128, 64, 32, 16, 8, 4, 2], encoder_embedding_dims=[8, 16, 32, 64, 128, 256, 512],
               **DEFAULT_KW)
    assert isinstance(models['unet_hierarchical_patch'](**cfg), torch.nn.Module)
    c = criteria['ExpectedLandmarkMse'](batch_size=2, frame_size=224, num_aux_graphs=7, use_main_graph_only=False,
                                        num_output_channels=4, loss_weight=10)
    assert c.grid_sizes == [2, 4, 8, 16, 32, 64, 128, 224]


def test_argument_validation_happens_before_any_cuda_call(eg):
    """Entry points reject malformed calls with EG_ERR_INVALID and a message from eg_last_error(), without
    touching the (absent) GPU: NULL pointers, unsupported channel counts, evaluator shapes."""
    lib = eg.lib
    assert lib.eg_level_embed_fwd(None, 1, 0, 4, None, None, None, None, None) < 0
    assert b"eg_level_embed_fwd" in lib.eg_last_error()
    assert lib.eg_level_embed_supported(None, 0, 4) == 0
    assert lib.eg_expected_coords(2, 3, 100, 4, 1, 1, None, 1, 1, 1, None) < 0  # 3 channels
    assert b"4 landmark channels" in lib.eg_last_error()
    assert lib.eg_expected_coords(2, 4, 10, 4, 1, 1, None, 1, 1, 1, None) < 0   # frame^2 > nodes per frame
    assert lib.eg_linear128_wgrad(0, None, None, None, None, None, 0, None) < 0
    assert lib.eg_gcn_conv_fwd(None, 1, None, None, None, None, None, None, None, 0, None) < 0
    assert lib.eg_classifier_fwd(4, None, None, None, None, None, None, None, None, None, None, 0, None) < 0
    assert b"eg_classifier_fwd" in lib.eg_last_error()
    assert lib.eg_classifier_bwd(4, None, None, None, None, None, None, None, None, None, None, None, None, None,
                                 None, 0, None) < 0
    assert b"eg_classifier_bwd" in lib.eg_last_error()


def test_registry_patch_with_evaluators(eg):
    from echoglad_b200 import register
    models, criteria, evaluators = {}, {}, {'accuracy': object, 'landmarkcoorderror': object}
    register.patch(models, criteria, evaluators)
    assert evaluators['landmarkcoorderror'] is eg.LandmarkExpectedCoordiantesEvaluator and evaluators['accuracy'] is object
    ev = evaluators['landmarkcoorderror'](logger=None, batch_size=2, frame_size=224, use_coord_graph=False)
    assert set(ev.coordinate_errors) == {'ivs', 'lvid_top', 'lvid_bot', 'lvpw'} and ev.get_predictions() == {}
    with pytest.raises(eg.EchogladError, match="no CPU fallback"):
        ev.update(torch.zeros(8, 4), torch.zeros(8, 4), torch.ones(2), torch.ones(2), torch.ones(8, 4))


def test_evaluator_host_arithmetic_matches_reference_golden(eg, monkeypatch):
    """Host half of the drop-in evaluator (errors in mm, width MAE / MPE, accumulation over batches) against the
    values the reference's own evaluator produced; the device half (eg_expected_coords) is replaced by the oracle's
    restatement here and tested against it under -m gpu."""
    from echoglad_b200 import evaluator as E
    from oracle import restated as R
    z = np.load(os.path.join(GOLDEN, "evaluator_expected_coords.npz"))

    def fake(y_pred, y_true, valid, batch, frame):
        p, g, v = R.expected_coords(y_pred, y_true, valid, batch, frame)
        return p, g.to(torch.int32), v

    monkeypatch.setattr(E, "expected_coords", fake)
    for name, (frame, naux, batch) in {"S16_n3_B3": (16, 3, 3), "S28_n4_B2": (28, 4, 2)}.items():
        ev = eg.LandmarkExpectedCoordiantesEvaluator(None, batch, frame, False)
        for step in range(2):
            pre = f"{name}/step{step}/"
            y = torch.cat([R.node_labels(c, frame, naux) for c in z[pre + "coords"]], dim=0)
            ev.update(torch.from_numpy(z[pre + "logits"]), y, torch.from_numpy(z[pre + "pix2mm_x"]),
                      torch.from_numpy(z[pre + "pix2mm_y"]), torch.from_numpy(z[pre + "valid"]))
            for k, v in ev.get_last().items():
                want = float(z[pre + "last/" + k])
                assert abs(v - want) <= 1e-5 * max(abs(want), 1.0), (name, step, k)
            for k, v in ev.get_predictions()["widths"].items():
                assert np.allclose(v.numpy(), z[pre + "width/" + k], rtol=1e-5, atol=1e-5), k
        for k, v in ev.compute().items():
            want = float(z[f"{name}/compute/{k}"])
            assert abs(v - want) <= 1e-5 * max(abs(want), 1.0), (name, k)
        assert abs(ev.get_sum_of_width_MAE() - sum(float(z[f"{name}/compute/{k}"]) for k in ("ivs_w", "lvid_w", "lvpw_w"))) < 1e-3


def test_pyramid_blocks_fall_back_to_plain_pytorch_off_the_device(eg):
    """conv -> ReLU -> BatchNorm2d blocks of the pyramid and the embedder block: where the BatchNorm2d kernels do not
    apply (CPU tensors here; eval mode, > 64 channels or odd plane sizes on the device) the modules run the textbook
    PyTorch composition with the reference's parameters (src/core/models.py:71-260,841-876)."""
    import torch.nn.functional as TF
    from echoglad_b200.modules import _DownConv
    torch.manual_seed(0)
    blk = _DownConv(3, 5, 4).train()
    x = torch.randn(2, 3, 9, 7)
    want = TF.batch_norm(TF.relu(blk.conv1(x)), None, None, blk.BN1.weight, blk.BN1.bias, True, 0.1, 1e-5)
    want = TF.batch_norm(TF.relu(blk.conv2(want)), None, None, blk.BN2.weight, blk.BN2.bias, True, 0.1, 1e-5)
    want = TF.adaptive_max_pool2d(want, 4)
    assert torch.allclose(blk(x), want, atol=1e-6)
    assert int(blk.BN1.num_batches_tracked) == 1 and not blk.BN1.fused_ok(x)
    emb = eg.CNN(out_channels=[4], kernel_sizes=[3], pool_sizes=[1], cnn_dropout_p=0.0).train()
    b = emb.conv[0][0]
    f = torch.randn(2, 1, 10, 6)
    z = TF.batch_norm(b.conv(f), None, None, b.bn.weight, b.bn.bias, True, 0.1, 1e-5) + b.one_by_one_cnn(f)
    assert torch.allclose(emb(f), TF.relu(z), atol=1e-6)
