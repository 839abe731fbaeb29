"""Train-mode BatchNorm2d of the PyTorch pyramid (modules._BNTrain2d) at its full-resolution shapes: time of
forward + backward and the distance of the gradients from an fp64 evaluation.
    python tools/bn2d_bench.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from echoglad_b200 import modules  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    for shape in [(64, 4, 224, 224), (64, 8, 224, 224), (64, 8, 128, 128), (64, 16, 128, 128), (64, 16, 64, 64)]:
        gen = torch.Generator(device=dev).manual_seed(shape[1])
        x = (torch.randn(*shape, device=dev, generator=gen) * 1.7).requires_grad_()
        w = (torch.rand(shape[1], device=dev, generator=gen) + 0.5).requires_grad_()
        b = torch.randn(shape[1], device=dev, generator=gen).requires_grad_()
        dy = torch.randn(*shape, device=dev, generator=gen)

        from echoglad_b200 import ops

        def run_old():  # relu + the PyTorch composition
            y, _, _ = modules._BNTrain2d.apply(torch.relu(x), w, b, 1e-5)
            return torch.autograd.grad(y, (x, w, b), dy)

        def run():  # eg_bn2d_fwd / eg_bn2d_bwd with the ReLU folded in
            y, _, _ = ops.BN2dTrain.apply(x, w, b, 1e-5, True, None)
            return torch.autograd.grad(y, (x, w, b), dy)

        for _ in range(3):
            run_old()
        torch.cuda.synchronize()
        o0, o1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        o0.record()
        for _ in range(10):
            run_old()
        o1.record()
        torch.cuda.synchronize()
        print(f"{shape}: relu + PyTorch composition fwd+bwd {o0.elapsed_time(o1) / 10:.3f} ms")

        for _ in range(3):
            g = run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            g = run()
        e1.record()
        torch.cuda.synchronize()
        xd, wd, bd = (t.detach().double().requires_grad_() for t in (x, w, b))
        yd = torch.nn.functional.batch_norm(xd.relu(), None, None, wd, bd, True, 0.0, 1e-5)
        gd = torch.autograd.grad(yd, (xd, wd, bd), dy.double())
        err = [float(((a.double() - r).abs().max() / r.abs().max())) for a, r in zip(g, gd)]
        print(f"{shape}: fwd+bwd {e0.elapsed_time(e1) / 10:.3f} ms   max-norm err dx {err[0]:.2e} dgamma {err[1]:.2e} dbeta {err[2]:.2e}")


if __name__ == "__main__":
    main()
