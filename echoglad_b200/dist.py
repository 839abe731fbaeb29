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
    """All gradients of `params` end up in one contiguous fp32 buffer whose GROUPS (contiguous slices) are all-reduced
    over NCCL as soon as the backward has produced them, and the optimizer reads views of that buffer.

    Step protocol: `zero()` (drops the `.grad`s, so autograd WRITES fresh gradients instead of launching one
    accumulate kernel per parameter into a zeroed buffer: ~230 tiny kernels per step on the default.yml model),
    `loss.backward()`, `all_reduce_mean()` (afterwards every `p.grad` is its view of the flat buffer, averaged over the
    ranks).

    `groups`: lists of parameters in the order the backward finishes them (default: one group).  With more than one
    rank, a post-accumulate hook on every parameter counts a group down; when its last gradient has been written the
    group's gradients are copied into its slice (one multi-tensor copy) and the slice's all-reduce is launched on a
    SIDE stream, so the exchange of the GNN-scope gradients (finished first: they are last in the forward) and of
    the decoder / encoder gradients overlaps the rest of the backward ("launched as soon as backward
    produces the bucket").  `all_reduce_mean()` then only launches what is left, waits and divides."""

    def __init__(self, params: Iterable[torch.nn.Parameter], groups: Optional[List[List[torch.nn.Parameter]]] = None,
                 overlap: Optional[bool] = None, group=None):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no trainable parameters")
        if groups is None:
            groups = [self.params]
        groups = [[p for p in g if p.requires_grad] for g in groups]
        groups = [g for g in groups if g]
        ids = [id(p) for g in groups for p in g]
        if sorted(ids) != sorted(id(p) for p in self.params) or len(set(ids)) != len(ids):
            raise ValueError("groups must partition the trainable parameters")
        self.params = [p for g in groups for p in g]  # flat-buffer order = group order
        self.pg = group
        dev = self.params[0].device
        total = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        self.views: List[torch.Tensor] = []
        self.group_of = {}
        self.group_params: List[List[torch.nn.Parameter]] = groups
        self.group_views: List[List[torch.Tensor]] = []
        self.group_slice: List[torch.Tensor] = []
        off = 0
        for gi, g in enumerate(groups):
            start, gv = off, []
            for p in g:
                n = p.numel()
                v = self.flat[off:off + n].view_as(p)
                self.views.append(v)
                gv.append(v)
                p.grad = v
                self.group_of[id(p)] = gi
                off += n
            self.group_views.append(gv)
            self.group_slice.append(self.flat[start:off])
        world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.overlap = (world > 1) if overlap is None else (bool(overlap) and world > 1)
        self._pending = [0] * len(groups)
        self._launched = [False] * len(groups)
        self._works = []
        self._side = torch.cuda.Stream(device=dev) if (self.overlap and dev.type == "cuda") else None
        self._hooks = []
        if self.overlap:
            for p in self.params:
                self._hooks.append(p.register_post_accumulate_grad_hook(self._on_grad))

    @property
    def nbytes(self) -> int:
        return self.flat.numel() * 4

    def zero(self) -> None:
        for p in self.params:
            p.grad = None
        self._pending = [len(g) for g in self.group_params]
        self._launched = [False] * len(self.group_params)
        self._works = []

    # ---- one group: fresh gradients -> its slice of the flat buffer, then (world > 1) its all-reduce ------------------
    def _gather_group(self, gi: int) -> None:
        src, dst = [], []
        for p, v in zip(self.group_params[gi], self.group_views[gi]):
            if p.grad is None:
                v.zero_()
            elif p.grad.data_ptr() != v.data_ptr():
                src.append(p.grad)
                dst.append(v)
        if src:
            torch._foreach_copy_(dst, src)

    def _launch(self, gi: int) -> None:
        if self._launched[gi]:
            return
        self._launched[gi] = True
        distributed = dist.is_available() and dist.is_initialized() and dist.get_world_size(self.pg) > 1
        if self._side is not None and distributed:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.flat.device))  # the gradients of this group have been enqueued
            with torch.cuda.stream(self._side):
                self._side.wait_event(ev)
                self._gather_group(gi)
                self._works.append(dist.all_reduce(self.group_slice[gi], op=dist.ReduceOp.SUM, group=self.pg,
                                                   async_op=True))
        else:
            self._gather_group(gi)
            if distributed:
                self._works.append(dist.all_reduce(self.group_slice[gi], op=dist.ReduceOp.SUM, group=self.pg,
                                                   async_op=True))

    def _on_grad(self, p: torch.nn.Parameter) -> None:
        gi = self.group_of[id(p)]
        self._pending[gi] -= 1
        if self._pending[gi] == 0:
            self._launch(gi)

    def gather(self) -> None:
        """Fresh gradients -> flat buffer; `.grad` = the views.  A parameter that received no gradient contributes
        zeros to the reduction and KEEPS `.grad = None`, so the optimizer skips it (no weight decay / momentum on
        unused parameters, as in the reference's single-process step); every rank runs the same graph, so the set
        of unused parameters is the same everywhere."""
        for gi in range(len(self.group_params)):
            self._launch(gi)
        for w in self._works:
            w.wait()  # the current stream waits for the collective
        self._works = []
        if self._side is not None:
            torch.cuda.current_stream(self.flat.device).wait_stream(self._side)
        for p, v in zip(self.params, self.views):
            if p.grad is not None:
                p.grad = v

    def all_reduce_mean(self, group=None) -> None:
        if group is not None:
            self.pg = group
        self.gather()
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(self.pg) > 1:
            self.flat.div_(dist.get_world_size(self.pg))


def backward_order_groups(embedder: torch.nn.Module, landmark: torch.nn.Module) -> List[List[torch.nn.Parameter]]:
    """Parameter groups of the default.yml model in the order its backward finishes them: the GNN scope (classifier
    heads, GNN layers, coordinate MLPs, level-embedding 1x1 convs: 0.3 MB), the UNet decoder, the UNet encoder + the
    CNN embedder.  Parameters of other sub-modules go with the last group."""
    first, second, last = [], [], []
    for name, p in landmark.named_parameters():
        top = name.split(".")[0]
        if top in ("gnn_layers", "node_classifiers", "node_coordinate_mlp", "linears"):
            first.append(p)
        elif top == "up_convs":
            second.append(p)
        else:
            last.append(p)
    last += list(embedder.parameters())
    return [g for g in (first, second, last) if g]
