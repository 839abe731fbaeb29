#!/usr/bin/env python
"""Per-entry-point device timing of the C-ABI library at the default.yml shapes (CUDA events, L2-exceeding
working sets).  Development aid; the judged numbers come from bench.py.

    python tools/kernel_bench.py [--batch 64] [--iters 10] [--only name,name]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import echoglad_b200 as eg  # noqa: E402
from echoglad_b200 import ops  # noqa: E402
from echoglad_b200._lib import WORKSPACE_BYTES, check, lib  # noqa: E402


def timeit(fn, iters, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--only", default="")
    ap.add_argument("--main-only", action="store_true", help="use_main_graph_only graph (all tiles are lattice tiles)")
    ap.add_argument("--frame", type=int, default=224, help="frame size (BASELINE configs[3]: 448 with --naux 8)")
    ap.add_argument("--naux", type=int, default=7)
    ap.add_argument("--conn", action="store_true", help="use_connection_nodes (hub rows: gather plan, CSR rows)")
    ap.add_argument("--diag", action="store_true", help="grid-diagonal lattices (gather plan, 8-neighbour rows)")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    kw = dict(main_graph_type="grid-diagonal", aux_graph_type="grid-diagonal") if args.diag else {}
    g = eg.DeviceGraph.get(eg.HierGraphSpec(frame_size=args.frame, num_aux_graphs=args.naux,
                                            use_main_graph_only=args.main_only, use_connection_nodes=args.conn, **kw), dev)
    B, N = args.batch, g.meta.num_nodes
    rows = B * N
    U = rows * 128 * 4
    gen = torch.Generator(device=dev).manual_seed(0)
    x = torch.randn(rows, 128, device=dev, generator=gen)
    dy = torch.randn(rows, 128, device=dev, generator=gen)
    w = torch.randn(128, 128, device=dev, generator=gen) * 0.1
    bias = torch.randn(128, device=dev, generator=gen)
    gamma = torch.rand(128, device=dev, generator=gen) + 0.5
    beta = torch.randn(128, device=dev, generator=gen)
    h = torch.empty_like(x)
    out = torch.empty_like(x)
    scratch = torch.empty_like(x)
    mean, var = torch.empty(128, device=dev), torch.empty(128, device=dev)
    dw = torch.empty(128, 128, device=dev)
    ws = torch.empty(WORKSPACE_BYTES, dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    P = lambda t: t.data_ptr()  # noqa: E731

    cases = {
        "aggregate": (lambda: check(lib.eg_gcn_aggregate(g.handle, B, 128, P(x), P(h), st)), 2 * U),
        "gcn_conv_fwd": (lambda: check(lib.eg_gcn_conv_fwd(g.handle, B, P(x), P(w), P(bias), P(h), P(mean), P(var),
                                                           P(ws), WORKSPACE_BYTES, st)), 2 * U),
        "gcn_conv_bwd": (lambda: check(lib.eg_gcn_conv_bwd(g.handle, B, P(x), P(w), P(dy), P(h), P(out), P(dw), None,
                                                           P(scratch), P(ws), WORKSPACE_BYTES, st)), 6 * U),
        # eval-mode layer in one launch (reads X twice: gathered + residual rows, writes Y)
        "gcn_layer_eval": (lambda: check(lib.eg_gcn_layer_eval_fwd(g.handle, B, P(x), P(w), P(bias), P(gamma), P(beta),
                                                                   P(mean), P(var), 1e-5, 1, 1, P(out), P(ws),
                                                                   WORKSPACE_BYTES, st)), 3 * U),
        "gcn_layer_eval_nores": (lambda: check(lib.eg_gcn_layer_eval_fwd(g.handle, B, P(x), P(w), P(bias), P(gamma),
                                                                         P(beta), P(mean), P(var), 1e-5, 1, 0, P(out),
                                                                         P(ws), WORKSPACE_BYTES, st)), 2 * U),
        "gcn_bwd_nowgrad": (lambda: check(lib.eg_gcn_conv_bwd(g.handle, B, P(x), P(w), P(dy), P(h), P(out), None, None,
                                                              P(scratch), P(ws), WORKSPACE_BYTES, st)), 4 * U),
        "linear128": (lambda: check(lib.eg_linear128(rows, P(x), P(w), 1, P(bias), None, P(h), P(mean), P(var),
                                                     P(ws), WORKSPACE_BYTES, st)), 2 * U),
        "wgrad128": (lambda: check(lib.eg_linear128_wgrad(rows, P(dy), P(x), P(dw), None, P(ws), WORKSPACE_BYTES,
                                                          st)), 2 * U),
        "bn_act_fwd": (lambda: check(lib.eg_bn_act_fwd(rows, 128, P(h), P(mean), P(var), P(gamma), P(beta), 1e-5,
                                                       0.5, 7, 1, P(x), P(out), st)), 3 * U),
        "bn_act_bwd": (lambda: check(lib.eg_bn_act_bwd(rows, 128, P(dy), P(h), P(mean), P(var), P(gamma), P(beta),
                                                       1e-5, 0.5, 7, 1, 1, P(out), P(mean.clone()), P(var.clone()),
                                                       P(ws), WORKSPACE_BYTES, st)), 5 * U),
    }
    w2 = torch.randn(4, 16, 32, device=dev, generator=gen) * 0.3
    b2 = torch.randn(4, 16, device=dev, generator=gen)
    z2 = torch.empty(rows, 64, device=dev)
    dz2 = torch.randn(rows, 64, device=dev, generator=gen)
    dw2, db2 = torch.empty(4, 16, 32, device=dev), torch.empty(4, 16, device=dev)
    m2, v2 = torch.empty(64, device=dev), torch.empty(64, device=dev)
    cases["clf_mid_fwd"] = (lambda: check(lib.eg_clf_mid_fwd(rows, P(x), P(w2), P(b2), P(z2), P(m2), P(v2), P(ws),
                                                             WORKSPACE_BYTES, st)), U * 3 // 2)
    cases["clf_mid_bwd"] = (lambda: check(lib.eg_clf_mid_bwd(rows, P(x), P(w2), P(dz2), P(out), P(dw2), P(db2), P(ws),
                                                             WORKSPACE_BYTES, st)), U * 5 // 2)
    only = [s for s in args.only.split(",") if s]
    check(lib.eg_col_stats(rows, 128, P(x), P(mean), P(var), P(ws), WORKSPACE_BYTES, st))
    res = {}
    for name, (fn, nbytes) in cases.items():
        if only and name not in only:
            continue
        ms = timeit(fn, args.iters)
        res[name] = {"ms": round(ms, 4), "algo_GB": round(nbytes / 1e9, 3), "GBps": round(nbytes / ms / 1e6, 1)}
        if hasattr(lib, "eg_tc_debug_read") and name in ("gcn_conv_fwd", "linear128", "gcn_conv_bwd"):
            import ctypes as C
            buf = (C.c_longlong * (148 * 20))()
            torch.cuda.synchronize()
            lib.eg_tc_debug_read(buf)
            import numpy as np
            a = np.array(buf).reshape(148, 20).mean(0)
            print(f"   tc wait cycles/CTA: compute-wait-operand {a[0]:.0f} compute-wait-raw {a[5]:.0f} loader-wait-raw-empty {a[6]:.0f} "
                  f"compute-fence {a[7]:.0f} lattice-chunk-busy {a[8]:.0f} over {a[9]:.0f} chunks mma-wait-acc-empty {a[1]:.0f} mma-wait-full {a[2]:.0f} epilogue-wait-acc-full {a[3]:.0f} total {a[4]:.0f} | compute0 span {a[10]:.0f} lattice-prologue {a[16]:.0f} release-tail {a[17]:.0f} setup {a[18]:.0f} lattice-loop-total {a[19]:.0f} | compute15: busy {a[11]:.0f} wait-operand {a[12]:.0f} wait-raw {a[13]:.0f} fence {a[15]:.0f} span {a[14]:.0f}")
        print(f"{name:14s} {ms:8.3f} ms   {nbytes / 1e9:6.2f} GB algorithmic   {nbytes / ms / 1e6:8.1f} GB/s", flush=True)
    print(json.dumps({"batch": B, "rows": rows, "results": res}))


if __name__ == "__main__":
    main()
