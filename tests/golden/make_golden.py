"""Mints the golden vectors under tests/golden/ by running the reference's OWN code
(`/root/reference/src/core/{datasets,models,criterion}.py`, imported unmodified through
`oracle/ref_shim.py`).  Runs only in the build container (the reference tree is not on the GPU box);
the outputs are committed.  Usage:  python tests/golden/make_golden.py [--skip-big]

Files written:
  graphs_small.npz   edge_index / node_type of reference `create_graphs` + `from_networkx` for small specs
  graph_hashes.json  sha256 of the int64 edge_index / float64 node_type bytes for the big specs
  labels.npz         reference `create_node_labels` outputs (incl. the -1 wrap-around quirk)
  model_*.npz        logits / losses / gradients of the reference modules + criteria with a
                     state_dict produced by `oracle.restated.init_landmark_state` (strict load =>
                     the key layout is the reference's)
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import ref_shim  # noqa: E402
from oracle import restated as R  # noqa: E402

SMALL_SPECS = [
    # (frame, naux, main_only, coord, conn, main_type, aux_type)
    (4, 1, False, False, False, "grid", "grid"),
    (8, 2, False, False, False, "grid", "grid"),
    (16, 3, False, False, False, "grid", "grid"),
    (12, 3, False, False, False, "grid", "grid"),
    (24, 4, False, False, False, "grid", "grid"),
    (28, 4, False, False, False, "grid", "grid"),
    (32, 4, False, False, False, "grid", "grid"),
    (16, 0, True, False, False, "grid", "grid"),
    (9, 0, True, False, False, "grid-diagonal", "grid"),
    (16, 3, False, True, False, "grid", "grid"),
    (16, 3, False, False, True, "grid", "grid"),
    (12, 3, False, True, True, "grid", "grid"),
    (8, 2, False, False, False, "grid-diagonal", "grid-diagonal"),
    (12, 3, False, False, False, "grid-diagonal", "grid-diagonal"),
    (16, 3, False, False, False, "grid-diagonal", "grid"),
    (16, 3, False, False, False, "grid", "grid-diagonal"),
    (12, 3, False, False, True, "grid-diagonal", "grid-diagonal"),
]
BIG_SPECS = [
    (56, 5, False, False, False, "grid", "grid"),
    (224, 7, False, False, False, "grid", "grid"),
    (224, 0, True, False, False, "grid", "grid"),
    (224, 7, False, True, False, "grid", "grid"),
    (224, 7, False, False, True, "grid", "grid"),
    (448, 8, False, False, False, "grid", "grid"),   # BASELINE configs[3]: 2x resolution, centre crop c = 16
]


def spec_key(s):
    f, n, mo, co, cn, mt, at = s
    return f"S{f}_n{n}_mo{int(mo)}_co{int(co)}_cn{int(cn)}_{mt}_{at}"


def ref_graph(s):
    f, n, mo, co, cn, mt, at = s
    return ref_shim.reference_graph(f, max(n, 1), main_only=mo, coord=co, conn=cn, main_type=mt, aux_type=at)


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def make_graphs(skip_big: bool, big_only: str = ""):
    """big_only: mint only the BIG_SPECS entry with this key and merge it into graph_hashes.json."""
    import networkx as nx

    if big_only:
        path = os.path.join(HERE, "graph_hashes.json")
        hashes = json.load(open(path))
        (s,) = [s for s in BIG_SPECS if spec_key(s) == big_only]
        t = time.time()
        ei, nt = ref_graph(s)
        e = ei.numpy().astype(np.int64)
        hashes["specs"][spec_key(s)] = {
            "num_nodes": int(nt.shape[0]), "num_edges": int(e.shape[1]),
            "edge_index_sha256": sha(e), "node_type_sha256": sha(nt.astype(np.float64)),
            "first_edges": e[:, :8].T.tolist()}
        print(spec_key(s), nt.shape[0], e.shape[1], f"{time.time() - t:.1f}s", flush=True)
        with open(path, "w") as fh:
            json.dump(hashes, fh, indent=1)
        return

    out = {}
    for s in SMALL_SPECS:
        ei, nt = ref_graph(s)
        out[spec_key(s) + "/edge_index"] = ei.numpy().astype(np.int32)
        out[spec_key(s) + "/node_type"] = nt.astype(np.int8)
    np.savez_compressed(os.path.join(HERE, "graphs_small.npz"), **out)
    if skip_big:
        return
    hashes = {"networkx": nx.__version__, "specs": {}}
    for s in BIG_SPECS:
        t = time.time()
        ei, nt = ref_graph(s)
        e = ei.numpy().astype(np.int64)
        hashes["specs"][spec_key(s)] = {
            "num_nodes": int(nt.shape[0]), "num_edges": int(e.shape[1]),
            "edge_index_sha256": sha(e), "node_type_sha256": sha(nt.astype(np.float64)),
            "first_edges": e[:, :8].T.tolist()}
        print(spec_key(s), nt.shape[0], e.shape[1], f"{time.time() - t:.1f}s", flush=True)
    with open(os.path.join(HERE, "graph_hashes.json"), "w") as fh:
        json.dump(hashes, fh, indent=1)


def make_labels():
    out = {}
    cases = {
        "S224_n7": (224, 7, False, [[0, 0], [223, 223], [111, 112], [56, 167]]),
        "S224_n7_wrap": (224, 7, False, [[-1, 5], [7, -1], [28, 28], [195, 3]]),
        "S12_n3": (12, 3, False, [[0, 11], [6, 6], [3, 9], [11, 0]]),
        "S16_mo": (16, 1, True, [[0, 15], [8, 8], [3, 9], [15, 0]]),
        "S28_n4": (28, 4, False, [[7, 14], [21, 27], [13, 13], [1, 26]]),
    }
    for k, (f, n, mo, coords) in cases.items():
        c = np.array(coords)
        out[k + "/coords"] = c
        out[k + "/y"] = ref_shim.reference_labels(c, f, n, main_only=mo).numpy().astype(np.int8)
        out[k + "/meta"] = np.array([f, n, int(mo)])
    np.savez_compressed(os.path.join(HERE, "labels.npz"), **out)


GNN_PREFIXES = ("gnn_layers.", "node_classifiers.", "node_coordinate_mlp.")


def run_reference_model(variant, cfg_kw, batch, seed, *, training, coords_seed=7):
    """Reference module + criteria forward/backward.  Returns dict of numpy arrays.  The input x is
    not stored: it is `torch.randn(batch, cin, S, S, generator=manual_seed(seed + 1))`."""
    _, models, criterion = ref_shim.load()
    cfg = R.Cfg(variant=variant, **cfg_kw)
    sd = R.init_landmark_state(cfg, seed=seed)
    ctor = dict(frame_size=cfg.frame_size, gnn_dropout_p=cfg.gnn_dropout_p,
                classifier_dropout_p=cfg.classifier_dropout_p, node_embedding_dim=cfg.node_embedding_dim,
                node_hidden_dim=cfg.node_hidden_dim, num_output_channels=4,
                num_gnn_layers=cfg.num_gnn_layers, num_aux_graphs=cfg.num_aux_graphs,
                gnn_jk_mode=cfg.gnn_jk_mode, classifier_hidden_dim=cfg.classifier_hidden_dim,
                residual=cfg.residual, use_coordinate_graph=cfg.use_coordinate_graph,
                output_activation=cfg.output_activation,
                use_connection_nodes=cfg.use_connection_nodes, use_main_graph_only=cfg.use_main_graph_only)
    if variant == "unet":
        model = models.UNETHierarchicalPatchModel(encoder_embedding_widths=[128, 64, 32, 16, 8, 4, 2],
                                                  encoder_embedding_dims=[8, 16, 32, 64, 128, 256, 512], **ctor)
        cin = 4
    else:
        model = models.HierarchicalPatchModel(**ctor)
        cin = cfg.node_embedding_dim
    model.load_state_dict(sd, strict=True)  # proves the key layout
    model.train(training)

    ei1, nt1 = ref_shim.reference_graph(cfg.frame_size, cfg.num_aux_graphs, main_only=cfg.use_main_graph_only,
                                        conn=cfg.use_connection_nodes, coord=cfg.use_coordinate_graph)
    n = nt1.shape[0]
    ei = R.batch_edge_index(ei1, n, batch)
    node_type = torch.tensor(np.tile(nt1, batch))
    batch_idx = torch.arange(batch).repeat_interleave(n)

    g = torch.Generator().manual_seed(seed + 1)
    x = torch.randn(batch, cin, cfg.frame_size, cfg.frame_size, generator=g, requires_grad=True)
    rng = np.random.default_rng(coords_seed)
    coords = rng.integers(0, cfg.frame_size, size=(batch, 4, 2))
    y = torch.cat([ref_shim.reference_labels(c, cfg.frame_size, cfg.num_aux_graphs,
                                             main_only=cfg.use_main_graph_only) for c in coords], dim=0)
    valid = torch.ones_like(y)
    if batch > 1:  # make the valid mask non-trivial: frame 1, channel 2 invalid everywhere
        n0 = y.shape[0] // batch
        valid[n0:2 * n0, 2] = 0.0

    node_coords = None
    if cfg.use_coordinate_graph:  # initial coordinates as the data set provides them (src/core/datasets.py:99) + jitter
        base = torch.tensor([[99.99, 112.57], [142.71, 90.67], [151.18, 86.25], [91.81, 117.91]]) * (cfg.frame_size / 224.0)
        node_coords = (base.repeat(batch, 1) + torch.rand(4 * batch, 2, generator=g)).clamp(0, cfg.frame_size - 1)
    logits, coord_pred = model(x=x, node_coords=None if node_coords is None else node_coords.clone(), edge_index=ei,
                               batch_idx=batch_idx, node_type=node_type)
    bce = criterion.WeightedBCEWithLogitsLoss(reduction="none", ones_weight=9000, loss_weight=1)
    elm = criterion.ExpectedLandmarkMSE(loss_weight=10, batch_size=batch, frame_size=cfg.frame_size,
                                        num_aux_graphs=cfg.num_aux_graphs,
                                        use_main_graph_only=cfg.use_main_graph_only, num_output_channels=4)
    pv, yv = logits.view(batch, -1, 4), y.view(batch, -1, 4)
    l_bce = bce.compute(pv, yv, valid)
    l_elm = elm.compute(pv, yv, valid)
    total = l_bce + l_elm
    extra = {}
    if cfg.use_coordinate_graph:
        l_mae = criterion.MAE().compute(coord_pred, torch.from_numpy(coords).float().view(-1, 2))
        total = total + l_mae
        extra = {"node_coords_in": node_coords.numpy(), "node_coords_out": coord_pred.detach().numpy(),
                 "loss_mae": l_mae.detach().numpy()}
    out = {**extra, "coords": coords, "valid": valid.numpy().astype(np.int8),
           "logits": logits.detach().numpy(), "loss_bce": l_bce.detach().numpy(),
           "loss_elmse": l_elm.detach().numpy(), "seed": np.array(seed), "batch": np.array(batch)}
    if training:
        total.backward()
        out["grad_x"] = x.grad.numpy()
        for k, p in model.named_parameters():
            if k.startswith(GNN_PREFIXES):
                out["grad/" + k] = p.grad.numpy()
            elif p.grad is not None:  # UNet part: checksums only (the tensors are large)
                out["gradsum/" + k] = np.array([p.grad.double().sum().item(), p.grad.double().abs().sum().item()])
        for k, v in model.state_dict().items():
            if k.startswith(GNN_PREFIXES) and "running_" in k:
                out["stat/" + k] = v.numpy()
    return out


def make_default_width_model():
    """The reference CONSTRUCTOR defaults node_hidden_dim=64 / classifier_hidden_dim=16 (src/core/models.py:290-296):
    layer 0 maps 128 -> 64 without a residual (widths differ, :434), the classifiers are 64 -> 16 -> 8 -> 1."""
    o = run_reference_model("avgpool", dict(frame_size=12, num_aux_graphs=3, gnn_dropout_p=0.0, classifier_dropout_p=0.0,
                                            node_hidden_dim=64, classifier_hidden_dim=16),
                            batch=3, seed=15, training=True)
    np.savez_compressed(os.path.join(HERE, "model_avgpool_S12_n3_h64_c16_train.npz"), **o)
    print("h64/c16", float(o["loss_bce"]), float(o["loss_elmse"]), flush=True)


def make_models(skip_big: bool):
    tiny = dict(frame_size=12, num_aux_graphs=3, gnn_dropout_p=0.0, classifier_dropout_p=0.0)
    for training in (False, True):
        tag = "train" if training else "eval"
        o = run_reference_model("avgpool", tiny, batch=3, seed=11, training=training)
        np.savez_compressed(os.path.join(HERE, f"model_avgpool_S12_n3_{tag}.npz"), **o)
        print("avgpool", tag, float(o["loss_bce"]), float(o["loss_elmse"]), flush=True)
    # jk=max, no residual, sigmoid head off, main-only ablation
    o = run_reference_model("avgpool", dict(frame_size=10, num_aux_graphs=1, use_main_graph_only=True,
                                            gnn_dropout_p=0.0, classifier_dropout_p=0.0, gnn_jk_mode="max",
                                            residual=False, num_gnn_layers=2),
                            batch=2, seed=12, training=True)
    np.savez_compressed(os.path.join(HERE, "model_avgpool_S10_mainonly_jkmax_train.npz"), **o)
    o = run_reference_model("avgpool", dict(frame_size=12, num_aux_graphs=3, use_connection_nodes=True,
                                            gnn_dropout_p=0.0, classifier_dropout_p=0.0),
                            batch=2, seed=13, training=True)
    np.savez_compressed(os.path.join(HERE, "model_avgpool_S12_n3_conn_train.npz"), **o)
    o = run_reference_model("avgpool", dict(frame_size=12, num_aux_graphs=3, use_coordinate_graph=True,
                                            gnn_dropout_p=0.0, classifier_dropout_p=0.0),
                            batch=3, seed=14, training=True)
    np.savez_compressed(os.path.join(HERE, "model_avgpool_S12_n3_coord_train.npz"), **o)
    print("coord", float(o["loss_bce"]), float(o["loss_elmse"]), float(o["loss_mae"]), flush=True)
    make_default_width_model()
    unet = dict(frame_size=16, num_aux_graphs=3, gnn_dropout_p=0.0, classifier_dropout_p=0.0)
    for training in (False, True):
        tag = "train" if training else "eval"
        o = run_reference_model("unet", unet, batch=2, seed=21, training=training)
        np.savez_compressed(os.path.join(HERE, f"model_unet_S16_n3_{tag}.npz"), **o)
        print("unet", tag, float(o["loss_bce"]), float(o["loss_elmse"]), flush=True)
    if skip_big:
        return
    t = time.time()
    o = run_reference_model("unet", dict(frame_size=224, num_aux_graphs=7, gnn_dropout_p=0.0,
                                         classifier_dropout_p=0.0), batch=1, seed=31, training=True)
    o.pop("grad_x")
    o.pop("valid")
    o["logits"] = o["logits"][::97].copy()  # strided subsample keeps the fixture small
    np.savez_compressed(os.path.join(HERE, "model_unet_S224_n7_train.npz"), **o)
    print("default.yml B=1", float(o["loss_bce"]), float(o["loss_elmse"]), f"{time.time() - t:.1f}s", flush=True)


def make_evaluator():
    """LandmarkExpectedCoordiantesEvaluator.update / compute (src/core/evaluators.py:291-391,430-449) run by the
    reference's own class on seeded logits / labels / valid masks: two batches per case, the second with one
    landmark invalid in one frame and one landmark invalid in all frames."""
    import importlib
    from oracle import ref_shim
    ref_shim.load()
    ev_mod = importlib.import_module("src.core.evaluators")
    out = {}
    for name, (frame, naux, batch) in {"S16_n3_B3": (16, 3, 3), "S28_n4_B2": (28, 4, 2)}.items():
        ev = ev_mod.LandmarkExpectedCoordiantesEvaluator(logger=None, batch_size=batch, frame_size=frame,
                                                         use_coord_graph=False)
        n0 = sum(4 ** k for k in range(1, naux + 1)) + frame * frame
        gen = torch.Generator().manual_seed(frame)
        for step in range(2):
            rng = np.random.default_rng(100 * frame + step)
            coords = rng.integers(0, frame, size=(batch, 4, 2))
            y = torch.cat([R.node_labels(c, frame, naux) for c in coords], dim=0)
            logits = torch.randn(batch * n0, 4, generator=gen) * 3.0
            valid = torch.ones(batch, n0, 4)
            if step == 1:
                valid[0, :, 1] = 0.0
                valid[:, :, 2] = 0.0
            valid = valid.view(batch * n0, 4)
            px = torch.rand(batch, generator=gen) + 0.5
            py = torch.rand(batch, generator=gen) + 0.5
            ev.update(logits, y, px, py, valid)
            pre = f"{name}/step{step}/"
            out[pre + "logits"] = logits.numpy()
            out[pre + "coords"] = coords
            out[pre + "valid"] = valid.numpy()
            out[pre + "pix2mm_x"] = px.numpy()
            out[pre + "pix2mm_y"] = py.numpy()
            last = ev.get_last()
            for k, v in last.items():
                out[pre + "last/" + k] = np.float64(v)
            for k, v in ev.detailed_performance["coordinates"].items():
                out[pre + "coord/" + k] = v.numpy()
            for k, v in ev.detailed_performance["widths"].items():
                out[pre + "width/" + k] = v.numpy()
        for k, v in ev.compute().items():
            out[f"{name}/compute/{k}"] = np.float64(v)
    np.savez_compressed(os.path.join(HERE, "evaluator_expected_coords.npz"), **out)
    print("evaluator", len(out), "arrays", flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--skip-big", action="store_true")
    ap.add_argument("--only", default="")
    ap.add_argument("--big-spec", default="", help="with --only graphs: mint just this BIG_SPECS key into graph_hashes.json")
    a = ap.parse_args()
    torch.manual_seed(0)
    torch.backends.cudnn.allow_tf32 = False
    if a.only in ("", "graphs"):
        make_graphs(a.skip_big, a.big_spec)
    if a.only in ("", "labels"):
        make_labels()
    if a.only in ("", "models", "models-small"):
        make_models(a.skip_big or a.only == "models-small")
    if a.only in ("", "evaluator"):
        make_evaluator()
