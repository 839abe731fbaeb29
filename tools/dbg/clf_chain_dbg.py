"""Debug aid: worst error / bound ratio of every output of the fused classifier chain for a few row counts."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from echoglad_b200 import ops
from tests._golden import close
DEV = torch.device("cuda", 0)

def run(rows, training, drop_p, sigmoid):
    gen = torch.Generator().manual_seed(rows)
    r = lambda *s: torch.randn(*s, generator=gen)
    h = r(rows, 128)
    prm = dict(w1=r(128, 128) * 0.15, b1=r(128) * 0.1, g1=torch.rand(128, generator=gen) + 0.5, be1=r(128) * 0.2,
               w2=r(4, 16, 32) * 0.3, b2=r(4, 16) * 0.1, g2=torch.rand(64, generator=gen) + 0.5, be2=r(64) * 0.2,
               w3=r(4, 16) * 0.4, b3=r(4) * 0.1)
    run_ = dict(m1=r(128) * 0.1, v1=torch.rand(128, generator=gen) + 0.5, m2=r(64) * 0.1, v2=torch.rand(64, generator=gen) + 0.5)
    dout = r(rows, 4)
    seed, eps = 4242, 1e-5
    dev = {k: v.to(DEV).requires_grad_(True) for k, v in prm.items()}
    hd = h.to(DEV).requires_grad_(True)
    out, m1, v1, m2, v2 = ops.ClassifierHeads.apply(
        hd, dev["w1"], dev["b1"], dev["g1"], dev["be1"], run_["m1"].to(DEV), run_["v1"].to(DEV), dev["w2"], dev["b2"],
        dev["g2"], dev["be2"], run_["m2"].to(DEV), run_["v2"].to(DEV), dev["w3"], dev["b3"], training, eps, drop_p, seed, sigmoid)
    out.backward(dout.to(DEV))
    ref = {k: v.double().requires_grad_(True) for k, v in prm.items()}
    hr = h.double().requires_grad_(True)
    p = drop_p if training else 0.0
    mask1 = ops.dropout_mask(rows, 128, p, seed, DEV).cpu().double()
    mask2 = ops.dropout_mask(rows, 64, p, seed + 1, DEV).cpu().double()
    z1 = hr @ ref["w1"].t() + ref["b1"]; z1.retain_grad()
    mu1, va1 = (z1.mean(0), z1.var(0, unbiased=False)) if training else (run_["m1"].double(), run_["v1"].double())
    a1 = torch.relu(((z1 - mu1) / torch.sqrt(va1 + eps) * ref["g1"] + ref["be1"]) * mask1)
    z2 = (torch.einsum("rki,kji->rkj", a1.view(rows, 4, 32), ref["w2"]) + ref["b2"]).reshape(rows, 64)
    mu2, va2 = (z2.mean(0), z2.var(0, unbiased=False)) if training else (run_["m2"].double(), run_["v2"].double())
    a2 = torch.relu(((z2 - mu2) / torch.sqrt(va2 + eps) * ref["g2"] + ref["be2"]) * mask2)
    want = torch.einsum("rkj,kj->rk", a2.view(rows, 4, 16), ref["w3"]) + ref["b3"]
    if sigmoid:
        want = torch.sigmoid(want)
    want.backward(dout.double())
    res = {"out": close(out.detach().cpu(), want.detach(), 1e-4, 1e-5)[1], "dh": close(hd.grad.cpu(), hr.grad, 1e-3, 1e-4)[1]}
    for k in prm:
        res["d" + k] = close(dev[k].grad.cpu(), ref[k].grad, 1e-3, 1e-4)[1]
    # where is dh wrong?
    err = (hd.grad.cpu().double() - hr.grad).abs().max(dim=1).values
    bad = torch.nonzero(err > 1e-3 * hr.grad.abs().max()).flatten()
    print(f"rows={rows} train={training} p={drop_p} sig={sigmoid}: " + " ".join(f"{k}={v:.2g}" for k, v in res.items()))
    if bad.numel():
        print("   bad dh rows:", bad.numel(), "first", bad[:12].tolist(), "last", bad[-5:].tolist())

ROWS = [int(a) for a in sys.argv[1:]] or [1000, 9472, 9473, 20000, 40000, 80000, 144040]
for rows in ROWS:
    run(rows, False, 0.0, False)
if len(sys.argv) == 1:
    run(144040, True, 0.0, False)
