#!/usr/bin/env python
"""Per-level timing of the generic strided transforms on the small pyramid levels (development aid)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import echoglad_b200 as eg  # noqa: E402
from echoglad_b200 import ops  # noqa: E402


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


def main():
    dev = torch.device("cuda", 0)
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    g = eg.DeviceGraph.get(eg.HierGraphSpec(frame_size=224, num_aux_graphs=7), dev)
    x = torch.randn(B * g.meta.num_nodes, 128, device=dev)
    dx = torch.randn_like(x)
    ws = ops._ws(dev)
    st = torch.cuda.current_stream().cuda_stream
    cins = [512, 256, 128, 64, 32, 16]
    tot = [0.0, 0.0, 0.0]
    for l, cin in enumerate(cins):
        s = g.meta.level_size[l]
        raw = torch.randn(B, cin, s, s, device=dev)
        d_raw = torch.empty_like(raw)
        w = torch.randn(128, cin, device=dev) * 0.1
        b = torch.randn(128, device=dev)
        dw, db = torch.empty_like(w), torch.empty_like(b)
        rows = B * s * s
        f = timeit(lambda: ops.linear_generic(rows, cin, 128, ops._view_nchw(raw), w, True, ops._view_level(x, g, l), bias=b, relu=True, stream=st))
        d = timeit(lambda: ops.linear_generic(rows, 128, cin, ops._view_level(dx, g, l), w, False, ops._view_nchw(d_raw), gate=ops._view_level(x, g, l), stream=st))
        wg = timeit(lambda: ops.linear_generic_wgrad(rows, cin, 128, ops._view_level(dx, g, l), ops._view_nchw(raw), dw, db, ws, gate=ops._view_level(x, g, l), stream=st))
        print(f"level {l} s={s:3d} cin={cin:3d} rows={rows:7d}: fwd {f:7.1f} us  dgrad {d:7.1f} us  wgrad {wg:7.1f} us")
        for i, v in enumerate((f, d, wg)):
            tot[i] += v
    print(f"total: fwd {tot[0]:.1f} us dgrad {tot[1]:.1f} us wgrad {tot[2]:.1f} us")


if __name__ == "__main__":
    main()
