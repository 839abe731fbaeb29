#!/usr/bin/env python
"""Practical HBM ceilings on this box for the access mixes of the elementwise kernels (development aid):
copy (1R+1W), add (2R+1W), a 2-read reduction and addcmul (3R+1W) in plain torch at the default.yml batch-64 size."""
import torch

dev = torch.device("cuda", 0)
rows = 4609280
a = torch.randn(rows, 128, device=dev)
b = torch.randn(rows, 128, device=dev)
c = torch.empty_like(a)
U = a.numel() * 4


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


for name, fn, n in [("copy 1R1W", lambda: c.copy_(a), 2), ("add 2R1W", lambda: torch.add(a, b, out=c), 3),
                    ("dot 2R", lambda: torch.dot(a.view(-1)[:2**31 - 8], b.view(-1)[:2**31 - 8]), 2 * (2**31 - 8) * 4 / U),
                    ("sum 1R", lambda: a.sum(), 1)]:
    ms = timeit(fn)
    print(f"{name:10s} {ms:7.3f} ms  {n * U / ms / 1e6:8.1f} GB/s")
