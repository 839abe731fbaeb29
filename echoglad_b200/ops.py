"""torch.autograd wrappers over the C ABI (ctypes; raw device pointers; torch's current stream).

PyTorch is plumbing here — device memory, streams, autograd bookkeeping — every FLOP and byte of the
GNN path runs in the hand-written sm_100a kernels of libechoglad_b200.so.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import torch

from ._lib import WORKSPACE_BYTES, EchogladError, check, lib
from .graph import DeviceGraph

F = 128


def _stream(t: torch.Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _f32(t: torch.Tensor, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise EchogladError(f"{name} must be a CUDA tensor: echoglad_b200 has no CPU fallback")
    if t.dtype != torch.float32:
        raise EchogladError(f"{name} must be float32 (got {t.dtype})")
    return t.contiguous()


def device_of(t: torch.Tensor):
    """Context manager: the tensor's CUDA device becomes the current one (the C ABI launches on the current device)."""
    if not t.is_cuda:
        raise EchogladError("expected a CUDA tensor: echoglad_b200 has no CPU fallback")
    return torch.cuda.device(t.device)


_WS_CACHE = {}


def _ws(device) -> torch.Tensor:
    """The library's scratch area (per-CTA partials of reductions), cached per (device, stream): every entry point
    finishes with its workspace inside the call, and calls on one stream are ordered, so one buffer per stream is
    enough (it was a fresh 10 MB allocation per call)."""
    device = torch.device(device)
    idx = device.index if device.index is not None else torch.cuda.current_device()
    key = (idx, torch.cuda.current_stream(idx).cuda_stream)
    t = _WS_CACHE.get(key)
    if t is None:
        t = _WS_CACHE[key] = torch.empty(WORKSPACE_BYTES, dtype=torch.uint8, device=torch.device("cuda", idx))
    return t


# ---------------------------------------------------------------------------------------------------------
# plain (non-autograd) calls, used by tests / bench and by the Functions below
# ---------------------------------------------------------------------------------------------------------

def gcn_aggregate(graph: DeviceGraph, batch: int, x: torch.Tensor) -> torch.Tensor:
    x = _f32(x, "x")
    out = torch.empty_like(x)
    check(lib.eg_gcn_aggregate(graph.handle, batch, x.shape[1], x.data_ptr(), out.data_ptr(), _stream(x)),
          "eg_gcn_aggregate")
    return out


def linear128(a: torch.Tensor, w: torch.Tensor, trans_w: bool, bias=None, addend=None, stats: bool = False):
    a, w = _f32(a, "a"), _f32(w, "w")
    out = torch.empty_like(a)
    mean = var = None
    ws = None
    if stats:
        mean = torch.empty(F, device=a.device)
        var = torch.empty(F, device=a.device)
        ws = _ws(a.device)
    check(lib.eg_linear128(a.shape[0], a.data_ptr(), w.data_ptr(), int(trans_w), _ptr(bias), _ptr(addend),
                           out.data_ptr(), _ptr(mean), _ptr(var), _ptr(ws), WORKSPACE_BYTES if stats else 0,
                           _stream(a)), "eg_linear128")
    return (out, mean, var) if stats else out


def linear128_wgrad(g: torch.Tensor, x: torch.Tensor, want_bias: bool = True):
    g, x = _f32(g, "g"), _f32(x, "x")
    dw = torch.empty(F, F, device=g.device)
    db = torch.empty(F, device=g.device) if want_bias else None
    ws = _ws(g.device)
    check(lib.eg_linear128_wgrad(g.shape[0], g.data_ptr(), x.data_ptr(), dw.data_ptr(), _ptr(db), ws.data_ptr(),
                                 WORKSPACE_BYTES, _stream(g)), "eg_linear128_wgrad")
    return dw, db


def col_stats(z: torch.Tensor):
    z = _f32(z, "z")
    mean = torch.empty(z.shape[1], device=z.device)
    var = torch.empty(z.shape[1], device=z.device)
    ws = _ws(z.device)
    check(lib.eg_col_stats(z.shape[0], z.shape[1], z.data_ptr(), mean.data_ptr(), var.data_ptr(), ws.data_ptr(),
                           WORKSPACE_BYTES, _stream(z)), "eg_col_stats")
    return mean, var


def dropout_mask(rows: int, cols: int, p: float, seed: int, device) -> torch.Tensor:
    m = torch.empty(rows, cols, device=device)
    check(lib.eg_dropout_mask(rows, cols, float(p), int(seed), m.data_ptr(),
                              torch.cuda.current_stream(device).cuda_stream), "eg_dropout_mask")
    return m


def bn_act_fwd(h, mean, var, gamma, beta, eps, drop_p, seed, relu, res=None):
    y = torch.empty_like(h)
    check(lib.eg_bn_act_fwd(h.shape[0], h.shape[1], h.data_ptr(), mean.data_ptr(), var.data_ptr(),
                            gamma.data_ptr(), beta.data_ptr(), float(eps), float(drop_p), int(seed), int(relu),
                            _ptr(res), y.data_ptr(), _stream(h)), "eg_bn_act_fwd")
    return y


def bn_act_bwd(dy, h, mean, var, gamma, beta, eps, drop_p, seed, relu, batch_stats):
    dh = torch.empty_like(h)
    dgamma = torch.empty_like(gamma)
    dbeta = torch.empty_like(beta)
    ws = _ws(h.device)
    check(lib.eg_bn_act_bwd(h.shape[0], h.shape[1], dy.data_ptr(), h.data_ptr(), mean.data_ptr(), var.data_ptr(),
                            gamma.data_ptr(), beta.data_ptr(), float(eps), float(drop_p), int(seed), int(relu),
                            int(batch_stats), dh.data_ptr(), dgamma.data_ptr(), dbeta.data_ptr(), ws.data_ptr(),
                            WORKSPACE_BYTES, _stream(h)), "eg_bn_act_bwd")
    return dh, dgamma, dbeta


def node_labels(coords: torch.Tensor, frame_size: int, level_size: Sequence[int], validate: bool = True) -> torch.Tensor:
    """coords int32[B, C, 2] (h, w) on device -> y float32[B, n0, C] (create_node_labels on device).
    validate=True raises IndexError, as the reference does (src/core/datasets.py:536-537), when a coordinate is
    >= frame_size (one host sync); validate=False skips the sync -- such labels are NaN-poisoned by the kernel, so a
    loss computed from them is NaN rather than silently wrong."""
    if not coords.is_cuda:
        raise EchogladError("coords must be a CUDA tensor")
    coords = coords.to(torch.int32).contiguous()
    b, c, _ = coords.shape
    n0 = sum(s * s for s in level_size)
    y = torch.empty(b, n0, c, device=coords.device)
    ls = (C.c_int32 * len(level_size))(*level_size)
    oob = torch.zeros(1, dtype=torch.int32, device=coords.device) if validate else None
    check(lib.eg_node_labels(b, c, frame_size, len(level_size), ls, coords.data_ptr(), y.data_ptr(), _ptr(oob),
                             _stream(coords)), "eg_node_labels")
    if validate and int(oob.item()):
        raise IndexError(f"{int(oob.item())} landmark coordinate(s) outside [-{frame_size}, {frame_size}): the "
                         "reference's create_node_labels raises IndexError for them (src/core/datasets.py:536-537)")
    return y


# ---------------------------------------------------------------------------------------------------------
# autograd Functions
# ---------------------------------------------------------------------------------------------------------

class PackNodes(torch.autograd.Function):
    """NCHW pyramid maps (+ optional connection / coordinate rows) -> node-major [B*N, 128]."""

    @staticmethod
    def forward(ctx, graph: DeviceGraph, head, tail, *maps):
        maps = [_f32(m, "map") for m in maps]
        meta = graph.meta
        if len(maps) != meta.num_levels:
            raise EchogladError(f"expected {meta.num_levels} level maps, got {len(maps)}")
        batch = maps[0].shape[0]
        for m, s in zip(maps, meta.level_size):
            if tuple(m.shape) != (batch, F, s, s):
                raise EchogladError(f"level map has shape {tuple(m.shape)}, expected {(batch, F, s, s)}")
        head = None if head is None else _f32(head, "head")
        tail = None if tail is None else _f32(tail, "tail")
        x = torch.empty(batch * meta.num_nodes, F, device=maps[0].device)
        arr = (C.c_void_p * len(maps))(*[m.data_ptr() for m in maps])
        check(lib.eg_pack_nodes(graph.handle, batch, arr, _ptr(head), _ptr(tail), x.data_ptr(), _stream(x)),
              "eg_pack_nodes")
        ctx.graph, ctx.batch = graph, batch
        ctx.has_head, ctx.has_tail = head is not None, tail is not None
        ctx.map_shapes = [m.shape for m in maps]
        return x

    @staticmethod
    def backward(ctx, dx):
        dx = _f32(dx, "dx")
        meta = ctx.graph.meta
        need = ctx.needs_input_grad
        d_maps: List[Optional[torch.Tensor]] = []
        for i, shp in enumerate(ctx.map_shapes):
            d_maps.append(torch.empty(shp, device=dx.device) if need[3 + i] else None)
        d_head = torch.empty(ctx.batch, meta.first_pixel_node, F, device=dx.device) if ctx.has_head else None
        d_tail = torch.empty(ctx.batch, meta.num_coord_nodes, F, device=dx.device) if ctx.has_tail else None
        arr = (C.c_void_p * len(d_maps))(*[_ptr(m) for m in d_maps])
        check(lib.eg_pack_nodes_grad(ctx.graph.handle, ctx.batch, dx.data_ptr(), arr, _ptr(d_head), _ptr(d_tail),
                                     _stream(dx)), "eg_pack_nodes_grad")
        return (None, d_head if need[1] else None, d_tail if need[2] else None, *d_maps)


def _view_rows(t: torch.Tensor):
    """eg_view of a contiguous row-major [rows, C] tensor."""
    from ._lib import View
    return View(t.data_ptr(), max(int(t.shape[0]), 1), 0, int(t.shape[1]), 1)


def _view_level(x: torch.Tensor, graph: DeviceGraph, level: int):
    """eg_view of lattice level `level` inside the node tensor x [B*N, F]."""
    from ._lib import View
    meta = graph.meta
    f = int(x.shape[1])
    s = meta.level_size[level]
    return View(x.data_ptr() + meta.level_offset[level] * f * 4, s * s, meta.num_nodes * f, f, 1)


def _view_nchw(m: torch.Tensor):
    """eg_view of a contiguous NCHW map [B, C, s, s] seen as [B*s*s, C]."""
    from ._lib import View
    p = int(m.shape[2] * m.shape[3])
    return View(m.data_ptr(), p, int(m.shape[1]) * p, 1, p)


def linear_generic(rows, k, n, a, w, trans_w, y, bias=None, gate=None, addend=None, relu=False, stream=None):
    """y = act(a_eff op(w) + bias + addend) over eg_views (see include/echoglad_b200.h: eg_linear_fwd)."""
    ref = lambda v: None if v is None else C.byref(v)  # noqa: E731
    check(lib.eg_linear_fwd(rows, k, n, C.byref(a), ref(gate), w.data_ptr(), int(trans_w), _ptr(bias), ref(addend),
                            int(relu), C.byref(y), stream), "eg_linear_fwd")


def linear_generic_wgrad(rows, k, n, g, a, dw, db, ws, gate=None, stream=None):
    check(lib.eg_linear_wgrad(rows, k, n, C.byref(g), None if gate is None else C.byref(gate), C.byref(a),
                              dw.data_ptr(), _ptr(db), ws.data_ptr(), WORKSPACE_BYTES, stream), "eg_linear_wgrad")


class EmbedPackNodes(torch.autograd.Function):
    """Node features of the UNet variant: per lattice level `relu(conv1x1(feature map))` packed node-major
    (reference src/core/models.py:708-710,722-756), no intermediate [B,128,s,s] map.  Levels whose raw map is narrow
    (cin 4 / 8: the main grid and the 128x128 level, 92 % of the nodes) run the eg_level_embed kernels; the small
    levels (cin 16..512) run the generic strided transform eg_linear_fwd straight from the NCHW map into the level's
    node rows (backward: eg_linear_fwd with the ReLU gate for d_in, eg_linear_wgrad).  A level NOT listed in
    fused_levels arrives already embedded ([B,128,s,s]) and goes through eg_pack_nodes.

    args: graph, fused_levels (tuple of level indices), then for every level l either
          (raw_map [B,cin,s,s], weight [128,cin,1,1], bias [128]) if l is fused, or (map [B,128,s,s], None, None)."""

    @staticmethod
    def forward(ctx, graph: DeviceGraph, fused_levels, *args):
        meta = graph.meta
        if len(args) != 3 * meta.num_levels:
            raise EchogladError(f"expected {3 * meta.num_levels} tensors, got {len(args)}")
        fused = set(fused_levels)
        maps = [_f32(args[3 * l], "map") for l in range(meta.num_levels)]
        batch = maps[0].shape[0]
        dev = maps[0].device
        x = torch.empty(batch * meta.num_nodes, F, device=dev)
        st = _stream(x)
        plain = []
        for l, s_l in enumerate(meta.level_size):
            if l in fused:
                plain.append(None)
                continue
            if tuple(maps[l].shape) != (batch, F, s_l, s_l):
                raise EchogladError(f"level {l} map has shape {tuple(maps[l].shape)}, expected {(batch, F, s_l, s_l)}")
            plain.append(maps[l])
        if any(m is not None for m in plain) or meta.first_pixel_node:
            arr = (C.c_void_p * len(plain))(*[_ptr(m) for m in plain])
            check(lib.eg_pack_nodes(graph.handle, batch, arr, None, None, x.data_ptr(), st), "eg_pack_nodes")
        saved, kinds = [], []
        for l in sorted(fused):
            w, b = _f32(args[3 * l + 1], "weight"), _f32(args[3 * l + 2], "bias")
            cin, s_l = maps[l].shape[1], meta.level_size[l]
            if tuple(maps[l].shape) != (batch, cin, s_l, s_l) or w.numel() != F * cin:
                raise EchogladError(f"level {l}: raw map {tuple(maps[l].shape)} / weight {tuple(w.shape)} mismatch")
            tc = bool(lib.eg_level_embed_supported(graph.handle, l, cin))
            if tc:
                check(lib.eg_level_embed_fwd(graph.handle, batch, l, cin, maps[l].data_ptr(), w.data_ptr(),
                                             b.data_ptr(), x.data_ptr(), st), "eg_level_embed_fwd")
            else:
                linear_generic(batch * s_l * s_l, cin, F, _view_nchw(maps[l]), w, True, _view_level(x, graph, l),
                               bias=b, relu=True, stream=st)
            kinds.append(tc)
            saved += [maps[l], w, b]
        ctx.save_for_backward(*saved)
        ctx.graph, ctx.batch, ctx.fused, ctx.kinds = graph, batch, sorted(fused), kinds
        ctx.shapes = [m.shape for m in maps]
        # the generic levels' ReLU mask is read from the output rows themselves.  Not through save_for_backward: the
        # coordinate rows of x may be rewritten in place afterwards (CoordSample), which never touches a level's rows.
        ctx.x_nodes = x.detach() if not all(kinds) else None
        return x

    @staticmethod
    def backward(ctx, dx):
        dx = _f32(dx, "dx")
        meta = ctx.graph.meta
        need = ctx.needs_input_grad
        st = _stream(dx)
        grads = [None] * (3 * meta.num_levels)
        d_plain = []
        for l in range(meta.num_levels):
            if l in ctx.fused or not need[2 + 3 * l]:
                d_plain.append(None)
            else:
                grads[3 * l] = torch.empty(ctx.shapes[l], device=dx.device)
                d_plain.append(grads[3 * l])
        if any(m is not None for m in d_plain):
            arr = (C.c_void_p * len(d_plain))(*[_ptr(m) for m in d_plain])
            check(lib.eg_pack_nodes_grad(ctx.graph.handle, ctx.batch, dx.data_ptr(), arr, None, None, st),
                  "eg_pack_nodes_grad")
        ws = _ws(dx.device)
        for k, l in enumerate(ctx.fused):
            raw, w, b = ctx.saved_tensors[3 * k:3 * k + 3]
            cin, s_l = raw.shape[1], meta.level_size[l]
            d_raw = torch.empty_like(raw) if need[2 + 3 * l] else None
            dw, db = torch.empty_like(w), torch.empty_like(b)
            if ctx.kinds[k]:
                check(lib.eg_level_embed_bwd(ctx.graph.handle, ctx.batch, l, cin, raw.data_ptr(), w.data_ptr(),
                                             b.data_ptr(), dx.data_ptr(), _ptr(d_raw), dw.data_ptr(), db.data_ptr(),
                                             ws.data_ptr(), WORKSPACE_BYTES, st), "eg_level_embed_bwd")
            else:
                rows = ctx.batch * s_l * s_l
                g, gate = _view_level(dx, ctx.graph, l), _view_level(ctx.x_nodes, ctx.graph, l)
                if d_raw is not None:  # d_in = (dX masked by the ReLU) W, written straight into the NCHW layout
                    linear_generic(rows, F, cin, g, w, False, _view_nchw(d_raw), gate=gate, stream=st)
                linear_generic_wgrad(rows, cin, F, g, _view_nchw(raw), dw, db, ws, gate=gate, stream=st)
            grads[3 * l], grads[3 * l + 1], grads[3 * l + 2] = d_raw, dw, db
        return (None, None, *grads)


class LinearGeneric(torch.autograd.Function):
    """y = x W^T + b for row-major tensors of ANY width (eg_linear_fwd / eg_linear_wgrad, fp32 FMA): the dense
    transforms of a module whose widths are not the tensor-core kernels' 128 / 32 / 4 (the reference constructor
    defaults are node_hidden_dim=64, classifier_hidden_dim=16, src/core/models.py:290-296)."""

    @staticmethod
    def forward(ctx, x, w, bias):
        x, w = _f32(x, "x"), _f32(w, "W")
        rows, k = x.shape
        n = w.shape[0]
        if w.shape[1] != k:
            raise EchogladError(f"LinearGeneric: x {tuple(x.shape)} does not match W {tuple(w.shape)}")
        y = torch.empty(rows, n, device=x.device)
        linear_generic(rows, k, n, _view_rows(x), w, True, _view_rows(y), bias=bias, stream=_stream(x))
        ctx.save_for_backward(x, w)
        ctx.has_bias = bias is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        dy = _f32(dy, "dy")
        rows, k = x.shape
        n = w.shape[0]
        st = _stream(dy)
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x)
            linear_generic(rows, n, k, _view_rows(dy), w, False, _view_rows(dx), stream=st)
        if ctx.needs_input_grad[1] or ctx.has_bias:
            dw = torch.empty_like(w)
            db = torch.empty(n, device=w.device) if ctx.has_bias else None
            linear_generic_wgrad(rows, k, n, _view_rows(dy), _view_rows(x), dw, db, _ws(w.device), stream=st)
        return dx, dw, db


class BNAct(torch.autograd.Function):
    """BatchNorm1d (batch or running statistics) -> Dropout -> ReLU|Identity (+ residual) on [rows, cols] with
    cols in {4, 8, ..., 128} (eg_col_stats / eg_bn_act_fwd / eg_bn_act_bwd).  Returns (y, batch_mean, batch_var)."""

    @staticmethod
    def forward(ctx, z, gamma, beta, mean_in, var_in, training: bool, eps: float, drop_p: float, seed: int, relu: bool,
                res):
        z = _f32(z, "z")
        if training:
            mean, var = col_stats(z)
        else:
            mean, var = _f32(mean_in, "running_mean"), _f32(var_in, "running_var")
        p = float(drop_p) if training else 0.0
        y = bn_act_fwd(z, mean, var, gamma, beta, eps, p, seed, relu, res)
        ctx.save_for_backward(z, gamma, beta, mean, var)
        ctx.cfg = (training, eps, p, seed, relu, res is not None)
        ctx.mark_non_differentiable(mean, var)
        return y, mean, var

    @staticmethod
    def backward(ctx, dy, _dm, _dv):
        z, gamma, beta, mean, var = ctx.saved_tensors
        training, eps, p, seed, relu, has_res = ctx.cfg
        dy = _f32(dy, "dy")
        dz, dgamma, dbeta = bn_act_bwd(dy, z, mean, var, gamma, beta, eps, p, seed, relu, training)
        return (dz, dgamma, dbeta) + (None,) * 7 + (dy if has_res else None,)


class Aggregate(torch.autograd.Function):
    """y = A_hat x (A_hat symmetric => backward is the same kernel)."""

    @staticmethod
    def forward(ctx, graph: DeviceGraph, batch: int, x):
        ctx.graph, ctx.batch = graph, batch
        return gcn_aggregate(graph, batch, x)

    @staticmethod
    def backward(ctx, dy):
        return None, None, gcn_aggregate(ctx.graph, ctx.batch, dy.contiguous())


# Test hook: when set to a list, every layer appends the sign pattern (bool tensor) of its ReLU input so a
# parity test can impose the same pattern on the CPU oracle (see oracle/restated.py:_relu).  Never set in
# production; costs one extra bn_act_fwd per GCN layer while active.
CAPTURE_RELU = None


class GCNLayer(torch.autograd.Function):
    """One reference GNN layer: GCNConv -> BatchNorm1d -> Dropout -> ReLU|Identity (+ residual)
    (reference src/core/models.py:329-335,431-435).  Returns (Y, batch_mean, batch_var)."""

    @staticmethod
    def forward(ctx, graph: DeviceGraph, batch: int, x, w, bias, gamma, beta, mean_in, var_in, training: bool,
                eps: float, drop_p: float, seed: int, relu: bool, residual: bool):
        x, w = _f32(x, "x"), _f32(w, "W")
        rows = batch * graph.meta.num_nodes
        if tuple(x.shape) != (rows, F) or tuple(w.shape) != (F, F):
            raise EchogladError(f"GCNLayer: x {tuple(x.shape)} / W {tuple(w.shape)} do not match rows={rows}, F={F}")
        ws = _ws(x.device)
        if not training and not residual and CAPTURE_RELU is None and not any(ctx.needs_input_grad):
            # inference (model.eval() under torch.no_grad(), src/engine.py:343-350) of a layer WITHOUT the residual
            # branch: the whole layer is one launch and the pre-activation is never written (1.22 ms against 2.32 ms at
            # default.yml / batch 64).  With the residual the epilogue's extra row reads make the single launch the
            # slower route (2.65 ms, profiles/r02_patch_experiments.txt r02ac/ad), so that case keeps the two launches.
            mean, var = _f32(mean_in, "running_mean"), _f32(var_in, "running_var")
            y = torch.empty_like(x)
            check(lib.eg_gcn_layer_eval_fwd(graph.handle, batch, x.data_ptr(), w.data_ptr(), _ptr(bias),
                                            _f32(gamma, "gamma").data_ptr(), _f32(beta, "beta").data_ptr(),
                                            mean.data_ptr(), var.data_ptr(), float(eps), int(relu), int(residual),
                                            y.data_ptr(), ws.data_ptr(), WORKSPACE_BYTES, _stream(x)),
                  "eg_gcn_layer_eval_fwd")
            ctx.mark_non_differentiable(mean, var)
            return y, mean, var
        h = torch.empty_like(x)
        if training:
            mean = torch.empty(F, device=x.device)
            var = torch.empty(F, device=x.device)
        else:
            mean, var = _f32(mean_in, "running_mean"), _f32(var_in, "running_var")
        check(lib.eg_gcn_conv_fwd(graph.handle, batch, x.data_ptr(), w.data_ptr(), _ptr(bias), h.data_ptr(),
                                  mean.data_ptr() if training else None, var.data_ptr() if training else None,
                                  ws.data_ptr(), WORKSPACE_BYTES, _stream(x)), "eg_gcn_conv_fwd")
        p = float(drop_p) if training else 0.0
        y = bn_act_fwd(h, mean, var, gamma, beta, eps, p, seed, relu, x if residual else None)
        if CAPTURE_RELU is not None and relu:
            CAPTURE_RELU.append(("gnn", (y if not residual else
                                         bn_act_fwd(h, mean, var, gamma, beta, eps, 0.0, seed, relu, None)) > 0))
        ctx.save_for_backward(x, h, w, gamma, beta, mean, var)
        ctx.cfg = (graph, batch, training, eps, p, seed, relu, residual)
        ctx.mark_non_differentiable(mean, var)
        return y, mean, var

    @staticmethod
    def backward(ctx, dy, _dmean, _dvar):
        x, h, w, gamma, beta, mean, var = ctx.saved_tensors
        graph, batch, training, eps, p, seed, relu, residual = ctx.cfg
        dy = _f32(dy, "dy")
        dh, dgamma, dbeta = bn_act_bwd(dy, h, mean, var, gamma, beta, eps, p, seed, relu, training)
        need_dx, need_dw = ctx.needs_input_grad[2], ctx.needs_input_grad[3]
        dx = torch.empty_like(x) if need_dx else None
        dw = torch.empty_like(w) if need_dw else None
        if need_dx or need_dw:
            scratch = torch.empty_like(x)
            ws = _ws(x.device)
            check(lib.eg_gcn_conv_bwd(graph.handle, batch, x.data_ptr(), w.data_ptr(), dh.data_ptr(),
                                      dy.data_ptr() if residual else None, _ptr(dx), _ptr(dw), None,
                                      scratch.data_ptr(), ws.data_ptr(), WORKSPACE_BYTES, _stream(x)),
                  "eg_gcn_conv_bwd")
        # d bias = column sums of dH: identically 0 through a train-mode BN, gamma*invstd*dbeta in eval mode
        if training:
            dbias = torch.zeros_like(gamma)
        else:
            dbias = gamma * torch.rsqrt(var + eps) * dbeta
        return (None, None, dx, dw, dbias, dgamma, dbeta) + (None,) * 8


class BN2dTrain(torch.autograd.Function):
    """Train-mode BatchNorm2d over an NCHW map with few channels, optionally of relu(x + pre_bias):
    y = BN(relu(x + pre_bias)) in two launches per direction (eg_bn2d_fwd / eg_bn2d_bwd) -- the
    conv3x3 -> ReLU -> BatchNorm2d blocks of the UNet pyramid at full resolution (reference src/core/models.py:841-876),
    `pre_bias` being the bias of a convolution that was run without it.  Returns (y, batch mean, biased batch variance)."""

    @staticmethod
    def forward(ctx, x, weight, bias, eps: float, relu_in: bool, pre_bias=None):
        x, weight, bias = _f32(x, "x"), _f32(weight, "weight"), _f32(bias, "bias")
        pre_bias = None if pre_bias is None else _f32(pre_bias, "pre_bias")
        n, c, h, w = x.shape
        y = torch.empty_like(x)
        mean, var = torch.empty(c, device=x.device), torch.empty(c, device=x.device)
        ws = _ws(x.device)
        check(lib.eg_bn2d_fwd(n, c, h * w, x.data_ptr(), _ptr(pre_bias), int(relu_in), weight.data_ptr(),
                              bias.data_ptr(), float(eps), y.data_ptr(), mean.data_ptr(), var.data_ptr(), ws.data_ptr(),
                              WORKSPACE_BYTES, _stream(x)), "eg_bn2d_fwd")
        ctx.save_for_backward(x, weight, mean, var, pre_bias)
        ctx.cfg = (float(eps), bool(relu_in))
        ctx.mark_non_differentiable(mean, var)
        return y, mean, var

    @staticmethod
    def backward(ctx, dy, _dmean, _dvar):
        x, weight, mean, var, pre_bias = ctx.saved_tensors
        eps, relu_in = ctx.cfg
        dy = _f32(dy, "dy")
        n, c, h, w = x.shape
        want_pb = pre_bias is not None and ctx.needs_input_grad[5]
        dx = torch.empty_like(x) if (ctx.needs_input_grad[0] or want_pb) else None
        dgamma, dbeta = torch.empty(c, device=x.device), torch.empty(c, device=x.device)
        dpb = torch.empty(c, device=x.device) if want_pb else None
        ws = _ws(x.device)
        check(lib.eg_bn2d_bwd(n, c, h * w, x.data_ptr(), _ptr(pre_bias), int(relu_in), dy.data_ptr(), mean.data_ptr(),
                              var.data_ptr(), weight.data_ptr(), eps, _ptr(dx), dgamma.data_ptr(), dbeta.data_ptr(),
                              _ptr(dpb), ws.data_ptr(), WORKSPACE_BYTES, _stream(x)), "eg_bn2d_bwd")
        return (dx if ctx.needs_input_grad[0] else None), dgamma, dbeta, None, None, dpb


class ClassifierHeads(torch.autograd.Function):
    """The four node classifiers (reference src/core/models.py:363-377,488-490) as ONE chain per direction
    (eg_classifier_fwd / eg_classifier_bwd): stacked / block-diagonal transforms whose BatchNorm / ReLU / Dropout are
    recomputed from the two saved pre-activations z1 [rows,128] and z2 [rows,64]; no activated tensor is written.
    Inputs are the STACKED parameters: w1 [128,128] (4 x [32,128]), b1 [128], g1/be1 [128],
    w2 [4,16,32], b2 [4,16], g2/be2 [64], w3 [4,16], b3 [4].
    Returns (out [rows,4], mean1, var1, mean2, var2)."""

    @staticmethod
    def _params(w1, b1, g1, be1, w2, b2, g2, be2, w3, b3, eps, p, seed, training, sigmoid):
        from ._lib import ClassifierParams
        return ClassifierParams(w1.data_ptr(), b1.data_ptr(), g1.data_ptr(), be1.data_ptr(), w2.data_ptr(),
                                b2.data_ptr(), g2.data_ptr(), be2.data_ptr(), w3.data_ptr(), b3.data_ptr(),
                                float(eps), float(p), int(seed), int(training), int(sigmoid))

    @staticmethod
    def forward(ctx, h, w1, b1, g1, be1, m1_in, v1_in, w2, b2, g2, be2, m2_in, v2_in, w3, b3, training: bool,
                eps: float, drop_p: float, seed: int, sigmoid: bool):
        h = _f32(h, "h")
        rows, dev, st = h.shape[0], h.device, _stream(h)
        if h.shape[1] != F:
            raise EchogladError(f"ClassifierHeads: h must be [rows, {F}], got {tuple(h.shape)}")
        p = float(drop_p) if training else 0.0
        prm = [_f32(t, "classifier parameter") for t in (w1, b1, g1, be1, w2, b2, g2, be2, w3, b3)]
        if training:
            m1, v1 = torch.empty(F, device=dev), torch.empty(F, device=dev)
            m2, v2 = torch.empty(64, device=dev), torch.empty(64, device=dev)
        else:
            m1, v1, m2, v2 = (_f32(t, "running statistics") for t in (m1_in, v1_in, m2_in, v2_in))
        z1 = torch.empty(rows, F, device=dev)
        z2 = torch.empty(rows, 64, device=dev)
        out = torch.empty(rows, 4, device=dev)
        cp = ClassifierHeads._params(*prm, eps, p, seed, training, sigmoid)
        ws = _ws(dev)
        check(lib.eg_classifier_fwd(rows, h.data_ptr(), C.byref(cp), m1.data_ptr(), v1.data_ptr(), m2.data_ptr(),
                                    v2.data_ptr(), z1.data_ptr(), z2.data_ptr(), out.data_ptr(), ws.data_ptr(),
                                    WORKSPACE_BYTES, st), "eg_classifier_fwd")
        if CAPTURE_RELU is not None:  # sign patterns of the two ReLU inputs (same arithmetic as the chain's kernels)
            CAPTURE_RELU.append(("clf_a", bn_act_fwd(z1, m1, v1, prm[2], prm[3], eps, 0.0, seed, True) > 0))
            CAPTURE_RELU.append(("clf_b", bn_act_fwd(z2, m2, v2, prm[6], prm[7], eps, 0.0, seed, True) > 0))
        ctx.save_for_backward(h, *prm, m1, v1, m2, v2, z1, z2, out)
        ctx.cfg = (training, eps, p, seed, sigmoid)
        ctx.mark_non_differentiable(m1, v1, m2, v2)
        return out, m1, v1, m2, v2

    @staticmethod
    def backward(ctx, dout, *_):
        h, w1, b1, g1, be1, w2, b2, g2, be2, w3, b3, m1, v1, m2, v2, z1, z2, out = ctx.saved_tensors
        training, eps, p, seed, sigmoid = ctx.cfg
        from ._lib import ClassifierGrads
        dout = _f32(dout, "dout")
        rows, dev, st = h.shape[0], h.device, _stream(h)
        cp = ClassifierHeads._params(w1, b1, g1, be1, w2, b2, g2, be2, w3, b3, eps, p, seed, training, sigmoid)
        grads = [torch.empty_like(t) for t in (w1, b1, g1, be1, w2, b2, g2, be2, w3, b3)]
        cg = ClassifierGrads(*[t.data_ptr() for t in grads])
        scratch = torch.empty(rows * (F + 64), device=dev)  # [rows,128] g1 -> dz1, then [rows,64] dz2
        dh = torch.empty_like(h) if ctx.needs_input_grad[0] else None
        ws = _ws(dev)
        check(lib.eg_classifier_bwd(rows, h.data_ptr(), C.byref(cp), m1.data_ptr(), v1.data_ptr(), m2.data_ptr(),
                                    v2.data_ptr(), z1.data_ptr(), z2.data_ptr(), out.data_ptr(), dout.data_ptr(),
                                    scratch.data_ptr(), _ptr(dh), C.byref(cg), ws.data_ptr(), WORKSPACE_BYTES, st),
              "eg_classifier_bwd")
        dw1, db1, dg1, dbe1, dw2, db2, dg2, dbe2, dw3, db3 = grads
        return (dh, dw1, db1, dg1, dbe1, None, None, dw2, db2, dg2, dbe2, None, None, dw3, db3) + (None,) * 5


class WeightedBCEWithLogits(torch.autograd.Function):
    """loss_weight * sum(valid * w(y) * bce(x, y)) / sum(valid)   (reference src/core/criterion.py:13-27)."""

    @staticmethod
    def forward(ctx, logits, y, valid, ones_weight: float, loss_weight: float):
        x = _f32(logits, "logits")
        y = _f32(y, "y")
        valid = _f32(valid, "valid")
        if y.numel() != x.numel() or valid.numel() != x.numel():
            raise EchogladError("WeightedBCEWithLogits: logits / y / valid element counts differ")
        loss = torch.empty((), device=x.device)
        grad = torch.empty_like(x) if logits.requires_grad else None
        ws = _ws(x.device)
        check(lib.eg_bce_multilevel(x.numel(), x.data_ptr(), y.data_ptr(), valid.data_ptr(), float(ones_weight),
                                    float(loss_weight), loss.data_ptr(), _ptr(grad), ws.data_ptr(), WORKSPACE_BYTES,
                                    _stream(x)), "eg_bce_multilevel")
        ctx.save_for_backward(grad)
        ctx.shape = logits.shape
        return loss

    @staticmethod
    def backward(ctx, dloss):
        (grad,) = ctx.saved_tensors
        return (grad * dloss).view(ctx.shape), None, None, None, None


class ExpectedLandmarkMSEFn(torch.autograd.Function):
    """Per-level soft-argmax MSE (reference src/core/criterion.py:93-151)."""

    @staticmethod
    def forward(ctx, logits, y, valid, batch: int, channels: int, level_size: tuple, loss_weight: float):
        x = _f32(logits, "logits")
        y = _f32(y, "y")
        valid = _f32(valid, "valid")
        n0 = sum(s * s for s in level_size)
        if x.numel() != batch * n0 * channels or y.numel() != x.numel() or valid.numel() != x.numel():
            raise EchogladError(f"ExpectedLandmarkMSE: expected {batch}x{n0}x{channels} elements, got {x.numel()}")
        loss = torch.empty((), device=x.device)
        grad = torch.empty_like(x) if logits.requires_grad else None
        ws = _ws(x.device)
        ls = (C.c_int32 * len(level_size))(*level_size)
        check(lib.eg_expected_landmark_mse(batch, channels, len(level_size), ls, x.data_ptr(), y.data_ptr(),
                                           valid.data_ptr(), float(loss_weight), loss.data_ptr(), _ptr(grad),
                                           ws.data_ptr(), WORKSPACE_BYTES, _stream(x)), "eg_expected_landmark_mse")
        ctx.save_for_backward(grad)
        ctx.shape = logits.shape
        return loss

    @staticmethod
    def backward(ctx, dloss):
        (grad,) = ctx.saved_tensors
        return (grad * dloss).view(ctx.shape), None, None, None, None, None, None


# ---------------------------------------------------------------------------------------------------------
# coordinate-graph branch (use_coordinate_graph): in-place Functions over the node tensor
# ---------------------------------------------------------------------------------------------------------

def _coord_geom(graph: DeviceGraph, frame_size: int):
    """(nodes per frame, first coordinate row, first main-lattice row) inside a frame."""
    meta = graph.meta
    if meta.num_coord_nodes != 4:
        raise EchogladError("the graph has no coordinate nodes (use_coordinate_graph=False)")
    return (meta.num_nodes, meta.num_nodes - meta.num_coord_nodes,
            meta.first_pixel_node + meta.num_pixel_nodes - frame_size * frame_size)


def _own_grad(dy: torch.Tensor) -> torch.Tensor:
    """The backward kernels rewrite the incoming gradient of the node tensor in place (4 + 16 rows per frame of a
    multi-GB tensor).  COORD_INPLACE_GRAD = False clones it first (one extra pass; for callers that keep a
    reference to that gradient, e.g. retain_grad() on the updated tensor)."""
    dy = _f32(dy, "dy")
    return dy if COORD_INPLACE_GRAD else dy.clone()


COORD_INPLACE_GRAD = True


class CoordSample(torch.autograd.Function):
    """Initial coordinate-node features: rows of the 4 coordinate nodes = bilinear sample of the frame's main-level
    rows at `coords` [4B,2] (src/core/models.py:526-527, 743-744, 539-553).  In place on x (eg_coord_sample_fwd)."""

    @staticmethod
    def forward(ctx, x, coords, graph: DeviceGraph, batch: int, frame_size: int):
        x = _f32(x, "x")
        ctx.coords_shape = coords.shape
        coords = _f32(coords.reshape(-1, 2), "node_coords")
        n, c0, m0 = _coord_geom(graph, frame_size)
        if x.shape != (batch * n, F) or coords.shape[0] != 4 * batch:
            raise EchogladError(f"CoordSample: x {tuple(x.shape)} / coords {tuple(coords.shape)} for batch {batch}")
        check(lib.eg_coord_sample_fwd(x.data_ptr(), batch, n, c0, m0, frame_size, coords.data_ptr(), _stream(x)),
              "eg_coord_sample_fwd")
        ctx.mark_dirty(x)
        ctx.save_for_backward(x, coords)
        ctx.cfg = (batch, n, c0, m0, frame_size)
        return x

    @staticmethod
    def backward(ctx, dy):
        x, coords = ctx.saved_tensors
        batch, n, c0, m0, s = ctx.cfg
        dy = _own_grad(dy)
        dc = torch.empty_like(coords) if ctx.needs_input_grad[1] else None
        check(lib.eg_coord_sample_bwd(dy.data_ptr(), x.data_ptr(), batch, n, c0, m0, s, coords.data_ptr(), _ptr(dc),
                                      _stream(dy)), "eg_coord_sample_bwd")
        return dy, (None if dc is None else dc.view(ctx.coords_shape)), None, None, None


class CoordUpdate(torch.autograd.Function):
    """Coordinate update after a GNN layer (src/core/models.py:438-473) as one kernel per direction: relative
    positions + coordinate-node embeddings -> node_coordinate_mlp[i] -> clamp -> re-sample -> coordinate rows of y
    overwritten in place.  Returns (y, new coords [4B,2], mean1, var1, mean2, var2)."""

    @staticmethod
    def forward(ctx, y, coords, graph: DeviceGraph, batch: int, frame_size: int, w1, b1, g1, be1, m1_in, v1_in,
                w2, b2, g2, be2, m2_in, v2_in, w3, b3, training: bool, eps: float, drop_p: float, seed: int):
        from ._lib import CoordMlpParams
        y = _f32(y, "y")
        ctx.coords_shape = coords.shape
        coords = _f32(coords.reshape(-1, 2), "node_coords")
        n, c0, m0 = _coord_geom(graph, frame_size)
        r, dev = 4 * batch, y.device
        if y.shape != (batch * n, F) or coords.shape[0] != r:
            raise EchogladError(f"CoordUpdate: y {tuple(y.shape)} / coords {tuple(coords.shape)} for batch {batch}")
        prm = [_f32(t, "coordinate MLP parameter") for t in (w1, b1, g1, be1, w2, b2, g2, be2, w3, b3)]
        if tuple(prm[0].shape) != (32, F + 8) or tuple(prm[4].shape) != (16, 32) or tuple(prm[8].shape) != (2, 16):
            raise EchogladError("CoordUpdate: node_coordinate_mlp must be Linear(136,32) / (32,16) / (16,2) "
                                "(classifier_hidden_dim == 32)")
        p = float(drop_p) if training else 0.0
        if training:
            m1, v1, m2, v2 = (torch.empty(k, device=dev) for k in (32, 32, 16, 16))
        else:
            m1, v1, m2, v2 = (_f32(t, "running statistics") for t in (m1_in, v1_in, m2_in, v2_in))
        feat_in = torch.empty(r, F, device=dev)
        z1, z2 = torch.empty(r, 32, device=dev), torch.empty(r, 16, device=dev)
        pre, out = torch.empty(r, 2, device=dev), torch.empty(r, 2, device=dev)
        cp = CoordMlpParams(*[t.data_ptr() for t in prm], float(eps), p, int(seed), int(training), 0)
        check(lib.eg_coord_update_fwd(y.data_ptr(), batch, n, c0, m0, frame_size, coords.data_ptr(), C.byref(cp),
                                      m1.data_ptr(), v1.data_ptr(), m2.data_ptr(), v2.data_ptr(), feat_in.data_ptr(),
                                      z1.data_ptr(), z2.data_ptr(), pre.data_ptr(), out.data_ptr(), _stream(y)),
              "eg_coord_update_fwd")
        ctx.mark_dirty(y)
        ctx.mark_non_differentiable(m1, v1, m2, v2)
        ctx.save_for_backward(y, coords, *prm, m1, v1, m2, v2, feat_in, z1, z2, pre, out)
        ctx.cfg = (batch, n, c0, m0, frame_size, training, eps, p, seed)
        return y, out, m1, v1, m2, v2

    @staticmethod
    def backward(ctx, dy, dcoords, *_):
        from ._lib import ClassifierGrads, CoordMlpParams
        y, coords, *rest = ctx.saved_tensors
        prm, (m1, v1, m2, v2, feat_in, z1, z2, pre, out) = rest[:10], rest[10:]
        batch, n, c0, m0, s, training, eps, p, seed = ctx.cfg
        dy = _own_grad(dy)
        dcoords = None if dcoords is None else _f32(dcoords.reshape(-1, 2), "dcoords")
        dev = dy.device
        grads = [torch.empty_like(t) for t in prm]
        cg = ClassifierGrads(*[t.data_ptr() for t in grads])  # same ten pointers as eg_coord_mlp_grads
        cp = CoordMlpParams(*[t.data_ptr() for t in prm], float(eps), p, int(seed), int(training), 0)
        scratch = torch.empty(4 * batch * 64, device=dev)
        dc_in = torch.empty_like(coords) if ctx.needs_input_grad[1] else None
        check(lib.eg_coord_update_bwd(dy.data_ptr(), _ptr(dcoords), y.data_ptr(), batch, n, c0, m0, s,
                                      coords.data_ptr(), C.byref(cp), m1.data_ptr(), v1.data_ptr(), m2.data_ptr(),
                                      v2.data_ptr(), feat_in.data_ptr(), z1.data_ptr(), z2.data_ptr(), pre.data_ptr(),
                                      out.data_ptr(), scratch.data_ptr(), C.byref(cg), _ptr(dc_in), _stream(dy)),
              "eg_coord_update_bwd")
        dw1, db1, dg1, dbe1, dw2, db2, dg2, dbe2, dw3, db3 = grads
        return (dy, None if dc_in is None else dc_in.view(ctx.coords_shape), None, None, None, dw1, db1, dg1, dbe1, None, None, dw2, db2, dg2, dbe2, None, None,
                dw3, db3, None, None, None, None)


class MAELoss(torch.autograd.Function):
    """loss_weight * mean |pred - y|  (reference src/core/criterion.py:52-64), loss and gradient in one launch."""

    @staticmethod
    def forward(ctx, pred, y, loss_weight: float):
        p = _f32(pred, "pred")
        y = _f32(y.to(torch.float32), "y")
        if p.numel() != y.numel() or p.numel() == 0:
            raise EchogladError("MAE: pred / y element counts differ")
        loss = torch.empty((), device=p.device)
        grad = torch.empty_like(p) if pred.requires_grad else None
        check(lib.eg_mae(p.numel(), p.data_ptr(), y.data_ptr(), float(loss_weight), loss.data_ptr(), _ptr(grad),
                         _stream(p)), "eg_mae")
        ctx.save_for_backward(grad)
        ctx.shape = pred.shape
        return loss

    @staticmethod
    def backward(ctx, dloss):
        (grad,) = ctx.saved_tensors
        return (grad * dloss).view(ctx.shape), None, None
