"""Diagnostic: per-level d(map) agreement of the default.yml batch-2 hot path vs the CPU oracle."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import echoglad_b200 as eg
from echoglad_b200 import ops
from oracle import restated as R
sys.path.insert(0, os.path.join(ROOT, "tests"))
from tests.test_gpu_parity import _build_module, _criteria

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
DEV = torch.device("cuda", 0)
cfg = R.Cfg(gnn_dropout_p=0.0, classifier_dropout_p=0.0)
batch = 2
frames, coords, y, valid = R.synthetic_batch(2, 224, 7, seed=200)
sd = R.init_landmark_state(cfg, seed=200)
esd = R.init_embedder_state(4, seed=201)
model = _build_module(cfg, "unet").to(DEV)
model.load_state_dict(sd, strict=True)
model.train()
emb = eg.CNN(out_channels=[4], kernel_sizes=[3], pool_sizes=[1], cnn_dropout_p=0.0).to(DEV)
emb.load_state_dict(esd, strict=True)
emb.train()
x = emb(frames.to(DEV))
maps = [m.detach().requires_grad_(True) for m in model.pyramid(x)]
graph = eg.DeviceGraph.get(model.graph_spec, DEV)
feats = ops.PackNodes.apply(graph, None, None, *maps)
hid = model.gnn_stack(feats, graph, batch)
logits = model.classify(hid)
bce, elm = _criteria(cfg, batch)
pv, yv = logits.view(batch, -1, 4), y.to(DEV).view(batch, -1, 4)
(bce.compute(pv, yv, valid.to(DEV)) + elm.compute(pv, yv, valid.to(DEV))).backward()
osd = R.clone_state(sd, requires_grad=True)
cmaps = [m.detach().cpu().requires_grad_(True) for m in maps]
ei, nt = R.build_edge_index(224, 7)
n = nt.shape[0]
lo = R.landmark_forward(osd, cfg, None, R.batch_edge_index(ei, n, batch), np.tile(nt, batch), True,
                        node_feats=R.pack_nodes(cfg, cmaps))
R.total_loss(lo, y, valid, cfg, batch)["total"].backward()
print("legacy" if os.environ.get("EG_LEGACY_MMA") == "1" else "tcgen05")
for lvl, (a, b) in enumerate(zip(maps, cmaps)):
    ga, gb = a.grad.cpu().double(), b.grad.double()
    err = (ga - gb).abs()
    bound = 1e-3 * gb.abs() + 1e-4 * gb.abs().max()
    bad = err > bound
    rms = float((ga - gb).pow(2).mean().sqrt() / gb.pow(2).mean().sqrt())
    where = bad.nonzero()
    pix = sorted({(int(w[0]), int(w[2]), int(w[3])) for w in where})[:12]
    print(f"level {lvl} shape {tuple(ga.shape)}: viol {int(bad.sum())} rms {rms:.2e} max_ratio {float((err / bound).max()):.2f} pixels {pix}")
