"""CPU tier: pins `oracle/restated.py` against the golden vectors minted from the reference's own
code (tests/golden/make_golden.py).  No GPU, no reference tree needed."""
import hashlib
import json
import os

import numpy as np
import pytest
import torch

from oracle import restated as R
from tests._golden import GOLDEN, MODEL_CASES, close, load_case, grads_close


def _parse(key):
    s, n, mo, co, cn, mt, at = key.split("_")
    return dict(frame=int(s[1:]), naux=max(int(n[1:]), 1), main_only=mo == "mo1", coord=co == "co1",
                conn=cn == "cn1", main_type=mt, aux_type=at)


def test_graph_small_bit_exact():
    z = np.load(os.path.join(GOLDEN, "graphs_small.npz"))
    keys = sorted({k.split("/")[0] for k in z.files})
    assert len(keys) >= 15
    for k in keys:
        p = _parse(k)
        ei, nt = R.build_edge_index(p.pop("frame"), p.pop("naux"), **p)
        assert np.array_equal(ei.numpy(), z[k + "/edge_index"].astype(np.int64)), k
        assert np.array_equal(nt, z[k + "/node_type"].astype(np.float64)), k


def test_graph_56_hash():
    h = json.load(open(os.path.join(GOLDEN, "graph_hashes.json")))["specs"]["S56_n5_mo0_co0_cn0_grid_grid"]
    ei, nt = R.build_edge_index(56, 5)
    assert ei.shape[1] == h["num_edges"] and nt.shape[0] == h["num_nodes"]
    assert hashlib.sha256(ei.numpy().astype(np.int64).tobytes()).hexdigest() == h["edge_index_sha256"]


def test_labels_bit_exact():
    z = np.load(os.path.join(GOLDEN, "labels.npz"))
    for k in sorted({k.split("/")[0] for k in z.files}):
        f, n, mo = (int(v) for v in z[k + "/meta"])
        y = R.node_labels(z[k + "/coords"], f, n, bool(mo))
        assert np.array_equal(y.numpy().astype(np.int8), z[k + "/y"]), k


def test_gcn_conv_matches_dense_formula():
    """Known-answer: D^-1/2 (A+I) D^-1/2 X W^T + b with a dense adjacency (fp64)."""
    ei, nt = R.build_edge_index(12, 3)
    n = nt.shape[0]
    g = torch.Generator().manual_seed(3)
    x = torch.randn(n, 16, generator=g, dtype=torch.float64)
    w = torch.randn(8, 16, generator=g, dtype=torch.float64)
    b = torch.randn(8, generator=g, dtype=torch.float64)
    a = torch.zeros(n, n, dtype=torch.float64)
    a[ei[1], ei[0]] = 1.0
    a += torch.eye(n, dtype=torch.float64)
    d = a.sum(1).pow(-0.5)
    dense = (d[:, None] * a * d[None, :]) @ x @ w.t() + b
    got = R.gcn_conv(x, ei, w, b)
    assert torch.allclose(got, dense, rtol=1e-12, atol=1e-12)
    # constant features on the isolated 2x2 grid: every node has deg 3 -> A_hat 1 = 1
    ei2, _ = R.build_edge_index(2, 1, main_only=True)
    ones = torch.ones(4, 1, dtype=torch.float64)
    out = R.gcn_conv(ones, ei2, torch.ones(1, 1, dtype=torch.float64), None)
    assert torch.allclose(out, ones)


@pytest.mark.parametrize("stem", [s for s in MODEL_CASES if "S224" not in s])
def test_model_and_losses_match_reference(stem):
    c = load_case(stem)
    cfg, z = c["cfg"], c["z"]
    sd = R.clone_state(c["sd"], requires_grad=c["training"])
    x = c["x"].clone().requires_grad_(c["training"])
    ei1, nt1 = R.build_edge_index(cfg.frame_size, cfg.num_aux_graphs, main_only=cfg.use_main_graph_only,
                                  conn=cfg.use_connection_nodes)
    n = nt1.shape[0]
    ei = R.batch_edge_index(ei1, n, c["batch"])
    logits = R.landmark_forward(sd, cfg, x, ei, np.tile(nt1, c["batch"]), c["training"])
    losses = R.total_loss(logits, c["y"], c["valid"], cfg, c["batch"])
    ok, worst = close(logits.detach(), z["logits"], 1e-4, 1e-5)
    assert ok, f"logits {worst}"
    assert abs(losses["WeightedBceWithLogits"].item() - float(z["loss_bce"])) <= 1e-5 * abs(float(z["loss_bce"]))
    assert abs(losses["ExpectedLandmarkMse"].item() - float(z["loss_elmse"])) <= 1e-5 * abs(float(z["loss_elmse"]))
    if c["training"]:
        losses["total"].backward()
        ok, worst = close(x.grad, z["grad_x"], 1e-3, 1e-4)
        assert ok, f"grad_x {worst}"
        want = {k[5:]: z[k] for k in z.files if k.startswith("grad/")}
        bad = grads_close({k: sd[k].grad for k in want}, want)
        assert not bad, bad
        for k in z.files:
            if k.startswith("grad/"):
                continue
            elif k.startswith("stat/"):
                ok, worst = close(sd[k[5:]], z[k], 1e-5, 1e-6)
                assert ok, f"{k} {worst}"
            elif k.startswith("gradsum/"):
                g = sd[k[8:]].grad.double()
                assert abs(g.abs().sum().item() - z[k][1]) <= 1e-3 * abs(z[k][1]) + 1e-12, k


def test_default_yml_full_graph_matches_reference():
    """default.yml (224 px, 7 aux, UNet variant), B=1, train mode with dropout p=0: the oracle's
    closed networkx restatement and model against the reference run (logits strided by 97)."""
    c = load_case("model_unet_S224_n7_train")
    cfg, z = c["cfg"], c["z"]
    h = json.load(open(os.path.join(GOLDEN, "graph_hashes.json")))["specs"]["S224_n7_mo0_co0_cn0_grid_grid"]
    ei, nt = R.build_edge_index(224, 7)
    assert hashlib.sha256(ei.numpy().astype(np.int64).tobytes()).hexdigest() == h["edge_index_sha256"]
    assert hashlib.sha256(nt.astype(np.float64).tobytes()).hexdigest() == h["node_type_sha256"]
    sd = R.clone_state(c["sd"], requires_grad=True)
    logits = R.landmark_forward(sd, cfg, c["x"], ei, nt, True)
    losses = R.total_loss(logits, c["y"], c["valid"], cfg, 1)
    ok, worst = close(logits.detach()[::97], z["logits"], 1e-4, 1e-5)
    assert ok, worst
    assert abs(losses["WeightedBceWithLogits"].item() - float(z["loss_bce"])) <= 1e-5 * float(z["loss_bce"])
    assert abs(losses["ExpectedLandmarkMse"].item() - float(z["loss_elmse"])) <= 1e-5 * float(z["loss_elmse"])
    losses["total"].backward()
    want = {k[5:]: z[k] for k in z.files if k.startswith("grad/")}
    bad = grads_close({k: sd[k].grad for k in want}, want)
    assert not bad, bad


def _evaluator_cases():
    z = np.load(os.path.join(GOLDEN, "evaluator_expected_coords.npz"))
    cases = {"S16_n3_B3": (16, 3, 3), "S28_n4_B2": (28, 4, 2)}
    return z, cases


def test_expected_coordinate_evaluator_restatement_matches_reference_golden():
    """oracle `expected_coords` + `expected_coord_metrics` against values minted by the reference's own
    LandmarkExpectedCoordiantesEvaluator (src/core/evaluators.py:291-391), incl. invalid landmarks."""
    z, cases = _evaluator_cases()
    for name, (frame, naux, batch) in cases.items():
        for step in range(2):
            pre = f"{name}/step{step}/"
            logits = torch.from_numpy(z[pre + "logits"])
            valid = torch.from_numpy(z[pre + "valid"])
            y = torch.cat([R.node_labels(c, frame, naux) for c in z[pre + "coords"]], dim=0)
            preds, gt, vs = R.expected_coords(logits, y, valid, batch, frame)
            assert np.array_equal(gt.numpy(), z[pre + "coords"])  # arg-max of a one-hot map is the landmark pixel
            got = R.expected_coord_metrics(preds, gt, vs, torch.from_numpy(z[pre + "pix2mm_x"]),
                                           torch.from_numpy(z[pre + "pix2mm_y"]))
            for k in ("lvid_top", "lvid_bot", "lvpw", "ivs", "ivs_w", "lvid_w", "lvpw_w", "ivs_mpe", "lvid_mpe", "lvpw_mpe"):
                want = float(z[pre + "last/" + k])
                assert abs(got[k] - want) <= 1e-5 * max(abs(want), 1.0), (name, step, k, got[k], want)
            assert np.allclose(preds[:, 3].numpy(), z[pre + "coord/pred_ivs"], rtol=1e-5, atol=1e-5)


def test_coordinate_graph_branch_matches_reference():
    """`use_coordinate_graph=True`: K4 coordinate nodes, per-layer coordinate MLP,
    bilinear re-sampling, MAE 'coordinate' loss — oracle against values minted by the reference module
    (src/core/models.py:438-473,539-553; src/core/criterion.py:52-64)."""
    z = np.load(os.path.join(GOLDEN, "model_avgpool_S12_n3_coord_train.npz"))
    cfg = R.Cfg(variant="avgpool", frame_size=12, num_aux_graphs=3, gnn_dropout_p=0.0, classifier_dropout_p=0.0,
                use_coordinate_graph=True)
    batch, seed = int(z["batch"]), int(z["seed"])
    sd = R.clone_state(R.init_landmark_state(cfg, seed=seed), requires_grad=True)
    x = torch.randn(batch, 128, 12, 12, generator=torch.Generator().manual_seed(seed + 1)).requires_grad_(True)
    ei1, nt1 = R.build_edge_index(12, 3, coord=True)
    n = nt1.shape[0]
    logits, coords = R.landmark_forward(sd, cfg, x, R.batch_edge_index(ei1, n, batch), np.tile(nt1, batch), True,
                                        node_coords=torch.from_numpy(z["node_coords_in"]))
    ok, worst = close(logits.detach(), z["logits"], 1e-4, 1e-5)
    assert ok, f"logits {worst}"
    ok, worst = close(coords.detach(), z["node_coords_out"], 1e-5, 1e-6)
    assert ok, f"coords {worst}"
    y = torch.cat([R.node_labels(c, 12, 3) for c in z["coords"]], dim=0)
    valid = torch.from_numpy(z["valid"].astype(np.float32))
    losses = R.total_loss(logits, y, valid, cfg, batch)
    mae = R.mae_loss(coords, torch.from_numpy(z["coords"]).float().view(-1, 2))
    assert abs(mae.item() - float(z["loss_mae"])) <= 1e-5 * float(z["loss_mae"])
    assert abs(losses["WeightedBceWithLogits"].item() - float(z["loss_bce"])) <= 1e-5 * abs(float(z["loss_bce"]))
    (losses["total"] + mae).backward()
    ok, worst = close(x.grad, z["grad_x"], 1e-3, 1e-4)
    assert ok, f"grad_x {worst}"
    want = {k[5:]: z[k] for k in z.files if k.startswith("grad/")}
    assert any(k.startswith("node_coordinate_mlp.") for k in want)
    bad = grads_close({k: sd[k].grad for k in want}, want)
    assert not bad, bad
