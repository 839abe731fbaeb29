#!/usr/bin/env python
"""C1' timing: default.yml landmark module + `use_coordinate_graph` (4 K4 coordinate nodes per frame,
per-layer coordinate update, extra MAE loss), module + loss level, forward + backward, against the same module without
the flag.  CUDA events on the current stream; per-kernel times of the coordinate kernels from the library's profile
hooks.

    python tools/coord_bench.py [--batches 2,64] [--iters 10]
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
from echoglad_b200._lib import profile_enable, profile_report  # noqa: E402


def step_fn(model, x, coords, y, valid, batch, bce, elm, mae, target):
    def step():
        logits, out = model(x=x, node_coords=coords)
        pv = logits.view(batch, -1, 4)
        loss = bce.compute(pv, y, valid) + elm.compute(pv, y, valid)
        if out is not None:
            loss = loss + mae.compute(out, target)
        loss.backward()
        model.zero_grad(set_to_none=True)
    return step


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
    ap.add_argument("--batches", default="2,64")
    ap.add_argument("--iters", type=int, default=10)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    kw = dict(frame_size=224, gnn_dropout_p=0.5, classifier_dropout_p=0.5, node_embedding_dim=128, node_hidden_dim=128,
              num_output_channels=4, num_gnn_layers=3, num_aux_graphs=7, gnn_jk_mode='last', classifier_hidden_dim=32,
              residual=True, output_activation='logit')
    res = {}
    for batch in [int(b) for b in args.batches.split(",")]:
        gen = torch.Generator(device=dev).manual_seed(200)
        x = torch.randn(batch, 128, 224, 224, device=dev, generator=gen).requires_grad_(True)
        lm = torch.randint(0, 224, (batch, 4, 2), device=dev, generator=gen, dtype=torch.int32)
        y = ops.node_labels(lm, 224, [2 ** k for k in range(1, 8)] + [224]).view(batch, -1, 4)
        valid = torch.ones(y.numel() // 4, 4, device=dev)
        bce = eg.WeightedBCEWithLogitsLoss(reduction='none', ones_weight=9000, loss_weight=1)
        elm = eg.ExpectedLandmarkMSE(loss_weight=10, batch_size=batch, frame_size=224, num_aux_graphs=7)
        mae = eg.MAE(loss_weight=1)
        # initial coordinates of the reference's data sets (src/core/datasets.py:99)
        init = torch.tensor([[99.99, 112.57], [142.71, 90.67], [151.18, 86.25], [91.81, 117.91]], device=dev)
        entry = {}
        for flag in (False, True):
            torch.manual_seed(200)
            model = eg.HierarchicalPatchModel(use_coordinate_graph=flag, **kw).to(dev).train()
            coords = init.repeat(batch, 1) if flag else None
            fn = step_fn(model, x, coords, y, valid, batch, bce, elm, mae, lm.view(-1, 2).float())
            ms = timeit(fn, args.iters)
            entry["with_coordinate_graph" if flag else "plain"] = round(ms, 3)
            if flag:
                profile_enable(True)
                for _ in range(args.iters):
                    fn()
                torch.cuda.synchronize()
                rep = profile_report()
                profile_enable(False)
                entry["coordinate_kernels_ms_per_step"] = {
                    k: round(v[0] / args.iters, 4) for k, v in rep.items() if k.startswith(("coord_", "mae"))}
            del model
        entry["extra_ms_per_step"] = round(entry["with_coordinate_graph"] - entry["plain"], 3)
        res[f"batch{batch}"] = entry
        print(f"batch {batch}: {entry}", flush=True)
        del x, y, valid
        torch.cuda.empty_cache()
    print(json.dumps({"workload": "default.yml landmark module (avg-pool pyramid of a [B,128,224,224] embedder output) + "
                                  "both losses (+ MAE), fwd+bwd, ms per step", "results": res}))


if __name__ == "__main__":
    main()
