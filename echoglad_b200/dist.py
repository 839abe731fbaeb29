"""Data-parallel plumbing: one process per GPU, frames sharded across ranks, ONE flat fp32 gradient
bucket all-reduced over NCCL (NVLink 5 / NVSwitch) per step.

Replaces the reference's single-process `torch_geometric.nn.DataParallel` / `nn.DataParallel`
(src/engine.py:105-110), which re-broadcasts all parameters and gathers logits to GPU 0 every step.
The batched graph is block-diagonal, so frames are independent through the whole GNN path.  BatchNorm
statistics stay per rank, exactly as they are per replica in the reference's DataParallel.

Loss normalisers are NOT per replica in the reference: DataParallel gathers the logits and `compute_loss` runs
once, so `sum(valid)` (BCE) and the per-level `num_valid` (expected-landmark MSE) are GLOBAL sums
(src/engine.py:589-598).  Here every rank normalises by its own sums and the gradients are averaged, which
equals the global loss exactly when all ranks hold the same normalisers -- true for the all-ones `valid` of the
synthetic benchmark and for equal shards of fully labelled frames.  For unevenly distributed invalid landmarks,
`global_normaliser_scale` gives the factor that turns a rank's `S_r / V_r` into its share of `sum S / sum V`
(exact for the BCE term; for the expected-landmark term it is exact when `valid` is constant over the nodes of
a frame and channel and the per-channel counts agree across ranks, otherwise an approximation).
"""
from __future__ import annotations

import os
from typing import Iterable, List, Optional

import torch
import torch.distributed as dist


def init_from_env(backend: Optional[str] = None) -> tuple:
    """(rank, local_rank, world_size); initialises torch.distributed when WORLD_SIZE > 1."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, local, world


def shard_range(num_frames: int, rank: int, world: int) -> range:
    """Frames [r*B/G, (r+1)*B/G) of a global batch; the global batch must divide evenly so that per-rank
    loss normalisers (sum(valid), num_valid) average to the global loss."""
    if num_frames % world != 0:
        raise ValueError(f"global batch {num_frames} does not divide over {world} ranks")
    per = num_frames // world
    return range(rank * per, (rank + 1) * per)


def global_normaliser_scale(local_normaliser: torch.Tensor, group=None) -> torch.Tensor:
    """G * V_r / sum_r V_r as a 0-dim tensor on the normaliser's device (1 when not distributed).  Multiplying rank
    r's loss `S_r / V_r` by it before `backward()` makes the averaged gradient equal the gradient of the reference's
    global `sum_r S_r / sum_r V_r`:  (1/G) sum_r [G V_r / V] grad(S_r) / V_r = sum_r grad(S_r) / V."""
    v = local_normaliser.detach().to(torch.float32).reshape(())
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return torch.ones_like(v)
    total = v.clone()
    dist.all_reduce(total, op=dist.ReduceOp.SUM, group=group)
    return v * dist.get_world_size(group) / total.clamp_min(torch.finfo(torch.float32).tiny)


class FlatGradBucket:
    """All gradients of `params` end up in one contiguous fp32 buffer, so the data-parallel reduction is a single
    all-reduce and the optimizer reads views of that buffer.

    Step protocol: `zero()` (drops the `.grad`s, so autograd WRITES fresh gradients instead of launching one
    accumulate kernel per parameter into a zeroed buffer: ~230 tiny kernels per step on the default.yml model),
    `loss.backward()`, `all_reduce_mean()` (multi-tensor copy of the fresh gradients into the flat buffer, one
    all-reduce, division by the world size; afterwards every `p.grad` is its view of the flat buffer)."""

    def __init__(self, params: Iterable[torch.nn.Parameter]):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no trainable parameters")
        dev = self.params[0].device
        total = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        self.views: List[torch.Tensor] = []
        off = 0
        for p in self.params:
            n = p.numel()
            self.views.append(self.flat[off:off + n].view_as(p))
            p.grad = self.views[-1]
            off += n

    @property
    def nbytes(self) -> int:
        return self.flat.numel() * 4

    def zero(self) -> None:
        for p in self.params:
            p.grad = None

    def gather(self) -> None:
        """Fresh gradients -> flat buffer; `.grad` = the views.  A parameter that received no gradient contributes
        zeros to the reduction and KEEPS `.grad = None`, so the optimizer skips it (no weight decay / momentum on
        unused parameters, as in the reference's single-process step); every rank runs the same graph, so the set
        of unused parameters is the same everywhere."""
        src, dst = [], []
        for p, v in zip(self.params, self.views):
            if p.grad is None:
                v.zero_()
            elif p.grad.data_ptr() != v.data_ptr():
                src.append(p.grad)
                dst.append(v)
        if src:
            torch._foreach_copy_(dst, src)
        for p, v in zip(self.params, self.views):
            if p.grad is not None:
                p.grad = v

    def all_reduce_mean(self, group=None) -> None:
        self.gather()
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
            self.flat.div_(dist.get_world_size(group))
