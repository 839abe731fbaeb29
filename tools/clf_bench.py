"""Classifier chain (ops.ClassifierHeads: eg_classifier_fwd / eg_classifier_bwd) alone at default.yml / batch 64, for
ncu captures of its kernels:  python tools/clf_bench.py [--batch 64] [--iters 3]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from echoglad_b200 import ops  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--iters", type=int, default=3)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    rows = args.batch * 72020
    gen = torch.Generator(device=dev).manual_seed(0)
    r = lambda *s: torch.randn(*s, device=dev, generator=gen)  # noqa: E731
    h = r(rows, 128).requires_grad_()
    prm = [r(128, 128) * 0.1, r(128), torch.rand(128, device=dev) + 0.5, r(128), torch.zeros(128, device=dev),
           torch.ones(128, device=dev), r(4, 16, 32) * 0.2, r(4, 16), torch.rand(64, device=dev) + 0.5, r(64),
           torch.zeros(64, device=dev), torch.ones(64, device=dev), r(4, 16) * 0.3, r(4)]
    for i in (0, 1, 2, 3, 6, 7, 8, 9, 12, 13):
        prm[i].requires_grad_()
    dout = r(rows, 4)
    for it in range(args.iters):
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        out = ops.ClassifierHeads.apply(h, *prm, True, 1e-5, 0.5, 11 + it, False)[0]
        out.backward(dout)
        t1.record()
        torch.cuda.synchronize()
        print(f"classifier chain fwd+bwd: {t0.elapsed_time(t1):.3f} ms")
        h.grad = None


if __name__ == "__main__":
    main()
