#!/usr/bin/env python
"""Headline benchmark: frames/s of one default.yml training step (embedder -> UNet pyramid -> packing ->
3 GCN layers -> 4 classifier heads -> WeightedBceWithLogits + ExpectedLandmarkMse -> backward -> Adam)
on synthetic DummyDataset-shaped frames, plus the HBM roofline of the aggregation kernel.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl native|reference]
    torchrun ... bench.py --gpus N ...            (one rank per GPU, weak scaling: B frames per GPU)

Prints ONE JSON line (rank 0).  `--impl reference` times the CPU oracle port of the reference path
(oracle/restated.py) on the host cores for the same metric.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "frames/sec fwd+bwd (default.yml, synthetic)"
F = 128
N_NODES = 72020  # default.yml graph (224 px, 7 aux levels)

LANDMARK_KW = dict(frame_size=224, gnn_dropout_p=0.5, classifier_dropout_p=0.5, node_embedding_dim=128,
                   node_hidden_dim=128, num_output_channels=4, num_gnn_layers=3, num_aux_graphs=7,
                   gnn_jk_mode='last', classifier_hidden_dim=32, residual=True, use_coordinate_graph=False,
                   output_activation='logit', use_connection_nodes=False, use_main_graph_only=False)


def workload_config(batch: int, world: int, cudnn_tf32: bool = False) -> dict:
    """`config` of the JSON line: identical on both arms (the reference arm times a bounded sample of this
    workload and says which in `cpu_baseline.sample`)."""
    return {"workload": f"default.yml full hierarchical graph, batch {batch} per GPU, training step "
                        "(embedder + UNet + GNN stack + classifiers + both losses, fwd+bwd, Adam)",
            "frame_size": 224, "num_aux_graphs": 7, "num_gnn_layers": 3, "batch_per_gpu": batch,
            "global_batch": batch * world, "nodes_per_step": batch * world * N_NODES,
            "parallelism": f"dp{world} (frames sharded; flat gradient buffer, 3 groups all-reduced over NCCL from autograd hooks on a side stream)",
            "l2": "working set (2.36 GB per node tensor) >> 126 MB L2, no flush",
            "precision": "fp32 storage; split-precision tensor-core transforms (tf32 x tf32 main term + both correction terms as one bf16 MMA, fp32 accumulate: within 2e-6 of fp64); "
                         f"cuDNN TF32 {'on' if cudnn_tf32 else 'off'}"}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "200", "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self) -> dict:
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(", ") for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except ValueError:
                continue
            for n, v in zip(names, r[5:9]):
                if v.strip().lower() == "active":
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------
# CPU oracle arm (cpu_baseline leg and --impl reference)
# ---------------------------------------------------------------------------------------------------------

def cpu_oracle_step_factory(batch: int):
    """One training step of the reference path restated on the CPU (oracle/restated.py): embedder ->
    landmark -> both losses -> backward -> Adam.  Only bench.py's baseline legs execute this."""
    from oracle import restated as R
    import numpy as np

    # nothing of echoglad_b200 (and so no libechoglad_b200.so) is imported on this arm: the graph comes from the
    # oracle's networkx restatement of `create_graphs` + `from_networkx` (built once, outside the timed steps, as
    # the reference's data set builds it once per sample in its loader workers)
    torch.set_num_threads(os.cpu_count() or 1)
    cfg = R.Cfg()
    sd = R.clone_state(R.init_landmark_state(cfg, seed=200), requires_grad=True)
    esd = R.clone_state(R.init_embedder_state(4, seed=201), requires_grad=True)
    params = [v for v in list(sd.values()) + list(esd.values()) if v.requires_grad]
    opt = torch.optim.Adam(params, lr=1e-3, weight_decay=1e-4)
    frames, coords, y, valid = R.synthetic_batch(batch, 224, 7, seed=200)
    ei1, nt1 = R.build_edge_index(224, 7)
    ei = R.batch_edge_index(ei1, nt1.shape[0], batch)
    node_type = np.tile(nt1, batch)

    def step():
        opt.zero_grad(set_to_none=True)
        x = R.embedder_forward(esd, frames, True, 0.1)
        logits = R.landmark_forward(sd, cfg, x, ei, node_type, True)
        loss = R.total_loss(logits, y, valid, cfg, batch)["total"]
        loss.backward()
        opt.step()
        return float(loss.detach())

    return step


def time_cpu_oracle(batch: int, steps: int, warmup: int):
    step = cpu_oracle_step_factory(batch)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return batch * steps / dt, dt / steps


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    batch = 2  # bounded sample: 2 of the workload's frames per step (the CPU rate is per frame; batch 64 needs ~100 GB)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    fps, sec = time_cpu_oracle(batch, max(1, args.steps), max(0, args.warmup))
    cores = torch.get_num_threads()
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": max(1, args.steps), "warmup": max(0, args.warmup), "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.batch, max(world, args.gpus), bool(args.cudnn_tf32)),
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
                         "sample": f"bounded sample of the workload: {max(1, args.steps)} training steps of {batch} "
                                   f"frames each instead of {args.batch} (same graph, model, losses and optimizer; "
                                   f"oracle/restated.py on torch CPU with {cores} threads, graph from its networkx "
                                   "restatement); the reference itself needs torch_geometric, absent here"},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------
# native arm
# ---------------------------------------------------------------------------------------------------------

def run_native(args):
    import echoglad_b200 as eg
    from echoglad_b200 import _lib, dist as egdist, ops, synthetic
    import torch.distributed as dist

    rank, local, world = egdist.init_from_env("nccl")
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    torch.backends.cudnn.allow_tf32 = bool(args.cudnn_tf32)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.benchmark = True
    if args.scaling == "strong":
        if args.batch % world:
            raise SystemExit(f"--scaling strong: global batch {args.batch} does not divide over {world} ranks")
        B = args.batch // world   # the GLOBAL batch is fixed; every rank takes its shard of frames
    else:
        B = args.batch            # weak: --batch frames per GPU

    torch.manual_seed(200)
    embedder = eg.CNN(out_channels=[4], kernel_sizes=[3], pool_sizes=[1], cnn_dropout_p=0.1).to(dev)
    landmark = eg.UNETHierarchicalPatchModel(encoder_embedding_widths=[128, 64, 32, 16, 8, 4, 2],
                                             encoder_embedding_dims=[8, 16, 32, 64, 128, 256, 512],
                                             **LANDMARK_KW).to(dev)
    embedder.train(); landmark.train()
    params = list(embedder.parameters()) + list(landmark.parameters())
    # gradient groups in backward order; with more than one rank each group's all-reduce is launched from autograd
    # hooks on a side stream as soon as its gradients exist (overlaps the rest of the backward)
    bucket = egdist.FlatGradBucket(params, groups=egdist.backward_order_groups(embedder, landmark))
    opt = torch.optim.Adam(params, lr=1e-3, weight_decay=1e-4, fused=True)
    bce = eg.WeightedBCEWithLogitsLoss(reduction='none', ones_weight=9000, loss_weight=1)
    spec = landmark.graph_spec
    graph = eg.DeviceGraph.get(spec, dev)
    assert graph.meta.num_nodes == N_NODES

    def build_steps(B):
        """(step_resident, step_e2e, frames_h, coords_h) for B frames per rank."""
        elm = eg.ExpectedLandmarkMSE(loss_weight=10, batch_size=B, frame_size=224, num_aux_graphs=7,
                                     use_main_graph_only=False, num_output_channels=4)
        frames_h, coords_h = synthetic.host_batch(B, 224, seed=200 + rank)
        frames_d, coords_d = frames_h.to(dev), coords_h.to(dev)
        y_d, valid_d = synthetic.device_labels(coords_d, spec)
        loss_h = torch.empty((), pin_memory=True)

        def fwd_bwd(frames, y, valid):
            bucket.zero()
            logits, _ = landmark(x=embedder(frames))
            pv, yv = logits.view(B, -1, 4), y.view(B, -1, 4)
            loss = bce.compute(pv, yv, valid) + elm.compute(pv, yv, valid)
            loss.backward()
            bucket.all_reduce_mean()
            opt.step()
            return loss

        def step_resident():
            return fwd_bwd(frames_d, y_d, valid_d)

        # e2e: every step copies its frames + landmark coordinates from pinned host memory and reads its loss back.
        # The copies are double-buffered on a copy stream (what a pinned, prefetching DataLoader does): step i computes
        # on buffer i % 2 while the input of step i + 1 lands in the other one -- whose last reader, step i - 1, has
        # finished, because every step ends with the synchronising loss read.  One H2D and one D2H per step.
        copy_stream = torch.cuda.Stream(device=dev)
        in_bufs = [(frames_d, coords_d), (torch.empty_like(frames_d), torch.empty_like(coords_d))]
        in_ready = [torch.cuda.Event(), torch.cuda.Event()]
        state = {"k": 0, "primed": False}

        def prefetch(k):
            with torch.cuda.stream(copy_stream):
                in_bufs[k][0].copy_(frames_h, non_blocking=True)
                in_bufs[k][1].copy_(coords_h, non_blocking=True)
                in_ready[k].record(copy_stream)

        def step_e2e():
            k = state["k"]
            if not state["primed"]:  # very first call: nothing was prefetched yet
                prefetch(k)
                state["primed"] = True
            prefetch(1 - k)  # the next step's input, in flight during this step
            torch.cuda.current_stream().wait_event(in_ready[k])
            frames, coords = in_bufs[k]
            y, valid = synthetic.device_labels(coords, spec)
            loss = fwd_bwd(frames, y, valid)
            loss_h.copy_(loss.detach(), non_blocking=False)  # D2H read of the step's result (synchronises)
            state["k"] = 1 - k
            return loss_h

        return step_resident, step_e2e, frames_h, coords_h

    step_resident, step_e2e, frames_h, coords_h = build_steps(B)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for _ in range(args.warmup):
        step_resident()
    torch.cuda.synchronize()

    if args.torch_profile:  # per-op CUDA time of one steady-state step (diagnostic, not a bench value)
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as p:
            step_resident()
            torch.cuda.synchronize()
        with open(args.torch_profile, "w") as fh:
            fh.write(p.key_averages().table(sort_by="cuda_time_total", row_limit=60, max_name_column_width=80))
    if args.ncu_range:  # `ncu --profile-from-start off`: only the steps below are captured
        torch.cuda.profiler.start()
        for _ in range(args.steps):
            step_resident()
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return

    sampler = ClockSampler(local) if rank == 0 else None
    _lib.profile_enable(True)
    launches0 = int(_lib.lib.eg_launch_count())
    ms = timed(step_resident, args.steps)
    launches = int(_lib.lib.eg_launch_count()) - launches0
    prof = _lib.profile_report()
    _lib.profile_enable(False)
    clocks = sampler.stop() if sampler else None

    for _ in range(2):
        step_e2e()
    ms_e2e = timed(step_e2e, args.steps)

    # informational only (never the headline): the same step with the PyTorch default `cudnn.allow_tf32 = True`
    # for the out-of-scope UNet / embedder convolutions, i.e. what the unmodified reference runs on an Ampere+ GPU
    ms_tf32 = None
    if not args.cudnn_tf32:
        torch.backends.cudnn.allow_tf32 = True
        for _ in range(3):
            step_resident()
        ms_tf32 = timed(step_resident, args.steps)
        torch.backends.cudnn.allow_tf32 = False

    # second regime on more than one rank (informational, same JSON line): STRONG scaling of the same step, the global
    # batch fixed at --batch and sharded over the ranks
    strong = None
    if world > 1 and args.scaling == "weak" and args.batch % world == 0 and not args.no_strong:
        Bs = args.batch // world
        s_res, _, _, _ = build_steps(Bs)
        for _ in range(3):
            s_res()
        _lib.profile_enable(True)
        ms_s = timed(s_res, args.steps)
        prof_s = _lib.profile_report()
        _lib.profile_enable(False)
        strong = {"global_batch": args.batch, "batch_per_gpu": Bs, "ms_per_step": ms_s / args.steps,
                  "frames_per_s": args.batch * args.steps / (ms_s / 1e3),
                  "native_kernel_ms_per_step_rank0": sum(v[0] for v in prof_s.values()) / args.steps,
                  "kernels_ms_per_step_rank0": {k: v[0] / args.steps for k, v in prof_s.items()}}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    frames_total = B * world * args.steps
    value = frames_total / (ms / 1e3)
    hbm_peak, peak_src = peaks()
    # dominant kernel: the fused tcgen05 message-passing + transform kernel (gather A_hat X -> split-precision MMA ->
    # bias / statistics / residual epilogue).  Algorithmic bytes per launch: forward = every input
    # row read once + every output row written once (+ the 64 KB weight) = 2 U; the backward launch also reads the
    # residual gradient and writes the A_hat dH side output = 4 U.  Index bytes are overhead and not counted.
    u_bytes = B * N_NODES * F * 4
    agg_ms, agg_n = prof.get("gcn_tc_fwd", (0.0, 0))
    a_agg = 2 * u_bytes + F * F * 4
    achieved = (a_agg * agg_n) / (agg_ms / 1e3) / 1e9 if agg_ms > 0 else None
    bwd_ms, bwd_n = prof.get("gcn_tc_bwd", (0.0, 0))
    a_bwd = 4 * u_bytes + F * F * 4
    achieved_bwd = (a_bwd * bwd_n) / (bwd_ms / 1e3) / 1e9 if bwd_ms > 0 else None
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")  # dram__bytes_read+write per launch from the last ncu --set full capture
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get(f"gcn_tc_batch{B}")
        except Exception:
            traffic = None
    kernels = {k: {"ms_per_step": v[0] / args.steps, "launches_per_step": v[1] / args.steps} for k, v in prof.items()}
    native_ms = sum(v[0] for v in prof.values()) / args.steps

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        fps, sec = time_cpu_oracle(2, 1, 1)
        cpu_baseline = {"value": fps, "unit": "frames/s", "cores": torch.get_num_threads(), "kind": "port",
                        "sample": "1 warm-up + 1 timed training step of batch 2 (oracle/restated.py on the host "
                                  f"CPU, {sec:.1f} s/step)"}

    line = {
        "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(B, world, bool(args.cudnn_tf32)),
        "clocks": clocks,
        "e2e": {"value": frames_total / (ms_e2e / 1e3), "unit": "frames/s",
                "h2d_bytes_per_step": int(frames_h.numel() * 4 + coords_h.numel() * 4) * world,
                "d2h_bytes_per_step": 4 * world, "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": launches,
        "roofline": {"bound": "hbm", "kernel": "gcn_tc forward launches (fused A_hat X gather, TMA patch staging + split-precision tcgen05 transform + bias/BN statistics)",
                     "achieved": achieved, "peak": hbm_peak,
                     "unit": "GB/s", "frac": (achieved / hbm_peak) if achieved else None, "traffic": traffic,
                     "peak_source": peak_src, "algorithmic_bytes_per_launch": a_agg,
                     "launches_per_step": agg_n / args.steps, "avg_launch_ms": (agg_ms / agg_n) if agg_n else None,
                     "backward_launches": {"achieved": achieved_bwd, "frac": (achieved_bwd / hbm_peak) if achieved_bwd else None,
                                           "algorithmic_bytes_per_launch": a_bwd, "launches_per_step": bwd_n / args.steps,
                                           "avg_launch_ms": (bwd_ms / bwd_n) if bwd_n else None}},
        "cpu_baseline": cpu_baseline,
        "strong_scaling": strong,
        "kernels": kernels,
        "native_kernel_ms_per_step": native_ms,
        "alt": {"note": "informational: same step with cudnn.allow_tf32=True (PyTorch default) for the UNet/embedder convolutions",
                "frames_per_s": (frames_total / (ms_tf32 / 1e3)) if ms_tf32 else None,
                "ms_per_step": (ms_tf32 / args.steps) if ms_tf32 else None},
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=64, help="frames per GPU (weak scaling) / global batch (strong)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak (default, the driver's scaling run): --batch frames per GPU; strong: --batch frames in "
                         "total, sharded over the ranks.  A weak multi-rank run also reports the strong regime under "
                         "`strong_scaling`")
    ap.add_argument("--no-strong", action="store_true", help="skip the informational strong-scaling leg of a weak run")
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--cudnn-tf32", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ncu-range", action="store_true", help="wrap --steps steps in cudaProfilerStart/Stop and exit")
    ap.add_argument("--torch-profile", default="", help="write a torch.profiler table of one step to this file")
    args = ap.parse_args()
    # stdout carries exactly ONE line (the JSON): libraries that print to fd 1 from native code (NCCL's version banner
    # at communicator creation) are routed to stderr until the line is printed.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real_stdout, "w")
    if args.impl == "reference":
        run_reference(args)
    else:
        if args.warmup < 3:
            args.warmup = 3
        run_native(args)


if __name__ == "__main__":
    main()
