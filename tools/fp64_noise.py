"""fp32 rounding noise of the full-size default.yml case: runs the CPU oracle in fp32 and in fp64 on identical
pyramid maps and counts d(map) entries outside the parity bound (basis of the outlier allowance in
tests/test_gpu_parity.py::test_default_yml_batch2_hot_path_against_oracle).  python tools/fp64_noise.py"""
import sys, numpy as np, torch
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import restated as R
torch.set_num_threads(16)
cfg = R.Cfg(gnn_dropout_p=0.0, classifier_dropout_p=0.0)
batch=2
frames, coords, y, valid = R.synthetic_batch(2, 224, 7, seed=200)
sd = R.init_landmark_state(cfg, seed=200); esd = R.init_embedder_state(4, seed=201)
x = R.embedder_forward(esd, frames, True, 0.0)
maps = [m.detach() for m in R.unet_pyramid(sd, cfg, x, True)]
ei, nt = R.build_edge_index(224, 7); n = nt.shape[0]
bei = R.batch_edge_index(ei, n, batch); ntb = np.tile(nt, batch)
def run(dt):
    osd = {k:(v.detach().to(dt) if v.is_floating_point() else v.clone()) for k,v in sd.items()}
    for k,v in osd.items():
        if v.is_floating_point() and "running" not in k: v.requires_grad_(True)
    for v in []:
        pass
    cm = [m.to(dt).detach().clone().requires_grad_(True) for m in maps]
    feats = R.pack_nodes(cfg, cm)
    lo = R.landmark_forward(osd, cfg, None, bei, ntb, True, node_feats=feats)
    want = R.total_loss(lo, y.to(dt), valid.to(dt), cfg, batch)
    want["total"].backward()
    return lo.detach(), [m.grad for m in cm], osd
l32,g32,s32 = run(torch.float32)
l64,g64,s64 = run(torch.float64)
print('logits max rel', ((l32.double()-l64).abs().max()/l64.abs().max()).item())
for lvl,(a,b) in enumerate(zip(g32,g64)):
    ga, gb = a.double().reshape(-1), b.reshape(-1)
    bound = 1e-3*gb.abs()+1e-4*gb.abs().max()
    viol = int(((ga-gb).abs()>bound).sum()); rms=float((ga-gb).pow(2).mean().sqrt()/gb.pow(2).mean().sqrt())
    print(lvl, ga.numel(), viol, f"{rms:.2e}")
