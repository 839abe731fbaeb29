"""Helpers shared by the golden-vector tests: rebuild the exact inputs `make_golden.py` used."""
import os

import numpy as np
import torch

from oracle import restated as R

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

MODEL_CASES = {
    # file stem -> (variant, cfg kwargs, training)
    "model_avgpool_S12_n3_eval": ("avgpool", dict(frame_size=12, num_aux_graphs=3), False),
    "model_avgpool_S12_n3_train": ("avgpool", dict(frame_size=12, num_aux_graphs=3), True),
    "model_avgpool_S10_mainonly_jkmax_train": ("avgpool", dict(frame_size=10, num_aux_graphs=1,
                                                               use_main_graph_only=True, gnn_jk_mode="max",
                                                               residual=False, num_gnn_layers=2), True),
    "model_avgpool_S12_n3_conn_train": ("avgpool", dict(frame_size=12, num_aux_graphs=3,
                                                        use_connection_nodes=True), True),
    "model_avgpool_S12_n3_h64_c16_train": ("avgpool", dict(frame_size=12, num_aux_graphs=3, node_hidden_dim=64,
                                                           classifier_hidden_dim=16), True),
    "model_unet_S16_n3_eval": ("unet", dict(frame_size=16, num_aux_graphs=3), False),
    "model_unet_S16_n3_train": ("unet", dict(frame_size=16, num_aux_graphs=3), True),
    "model_unet_S224_n7_train": ("unet", dict(frame_size=224, num_aux_graphs=7), True),
}


def load_case(stem):
    variant, kw, training = MODEL_CASES[stem]
    z = np.load(os.path.join(GOLDEN, stem + ".npz"))
    cfg = R.Cfg(variant=variant, gnn_dropout_p=0.0, classifier_dropout_p=0.0, **kw)
    seed, batch = int(z["seed"]), int(z["batch"])
    sd = R.init_landmark_state(cfg, seed=seed)
    cin = 4 if variant == "unet" else cfg.node_embedding_dim
    g = torch.Generator().manual_seed(seed + 1)
    x = torch.randn(batch, cin, cfg.frame_size, cfg.frame_size, generator=g)
    coords = z["coords"]
    y = torch.cat([R.node_labels(c, cfg.frame_size, cfg.num_aux_graphs, cfg.use_main_graph_only)
                   for c in coords], dim=0)
    if "valid" in z:
        valid = torch.from_numpy(z["valid"].astype(np.float32))
    else:
        valid = torch.ones_like(y)
    return dict(cfg=cfg, sd=sd, x=x, y=y, valid=valid, batch=batch, training=training, z=z)


def close(a, b, rtol=1e-4, atol_frac=1e-5):
    """|a-b| <= rtol*|b| + atol_frac*max|b|  (SURVEY.md §7.3 tolerance form)."""
    a = torch.as_tensor(a, dtype=torch.float64).reshape(-1)
    b = torch.as_tensor(b, dtype=torch.float64).reshape(-1)
    scale = b.abs().max().item() if b.numel() else 0.0
    err = (a - b).abs()
    bound = rtol * b.abs() + atol_frac * scale + 1e-30
    worst = (err / bound).max().item() if b.numel() else 0.0
    return worst <= 1.0, worst


def structurally_zero(key: str) -> bool:
    """Parameters whose true gradient is exactly 0 in train mode, so both sides hold only rounding
    noise: a bias feeding straight into a train-mode BatchNorm (GCNConv bias, classifier Linear 0/4
    bias).  (The last GNN BN bias under jk='last' is zero too, but is left to the generic floor.)"""
    if key.endswith("module_0.bias"):
        return True
    parts = key.split(".")
    return parts[0] == "node_classifiers" and parts[2] in ("0", "4") and parts[3] == "bias"


def grads_close(got: dict, want: dict, rtol=1e-3, atol_frac=1e-4, noise_frac=1e-6, zero_frac=1e-3):
    """Compares gradient dicts key by key with `close`.  Tensors that are pure rounding noise on both
    sides pass when both stay below noise_frac * (largest gradient entry of the whole set), or
    zero_frac * that for the `structurally_zero` keys.  Returns the list of failures."""
    gmax = max(float(np.abs(np.asarray(v)).max()) for v in want.values())
    bad = []
    for k, w in want.items():
        g = got[k]
        ok, worst = close(g, w, rtol, atol_frac)
        if not ok:
            floor = (zero_frac if structurally_zero(k) else noise_frac) * gmax
            if float(torch.as_tensor(g).abs().max()) <= floor and float(np.abs(np.asarray(w)).max()) <= floor:
                continue
            bad.append((k, worst))
    return bad
