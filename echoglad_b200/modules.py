"""Drop-in `nn.Module`s for the reference's landmark models, with IDENTICAL constructor kwargs, forward
signature and `state_dict` layout, whose GNN stack / classifiers run on the sm_100a
kernels.  The CNN embedder and the UNet pyramid stay in PyTorch (north star: "CNN embedder left in
PyTorch"); they are written here only so the replacement is self-contained.

Reference classes replaced:
  UNETHierarchicalPatchModel  src/core/models.py:639-756   (MODELS['unet_hierarchical_patch'])
  HierarchicalPatchModel      src/core/models.py:262-553   (MODELS['hierarchicalpatch'])
  CNN (embedder)              src/core/models.py:161-260   (MODELS['cnn'])  — plain PyTorch, out of scope
"""
from __future__ import annotations

import math
from typing import List, Optional

import torch
import torch.nn as nn
import torch.nn.functional as TF

from . import ops
from ._lib import EchogladError
from .graph import DeviceGraph, HierGraphSpec

F = 128
BN_EPS = 1e-5
BN_MOMENTUM = 0.1


# ---------------------------------------------------------------------------------------------------------
# parameter containers (names chosen so that state_dict keys equal the reference's)
# ---------------------------------------------------------------------------------------------------------

class _GCNConvParams(nn.Module):
    """Holds PyG GCNConv's parameters: `lin.weight` [out,in] (glorot, no bias) and `bias` (zeros)."""

    def __init__(self, fin: int, fout: int):
        super().__init__()
        self.lin = nn.Linear(fin, fout, bias=False)
        self.bias = nn.Parameter(torch.zeros(fout))
        a = math.sqrt(6.0 / (fin + fout))
        nn.init.uniform_(self.lin.weight, -a, a)


class _GNNBlock(nn.Module):
    """`Sequential('x, edge_index', [GCNConv, BatchNorm1d, Dropout, ReLU|Identity])` of the reference
    (src/core/models.py:329-335): children `module_0` (conv params) and `module_1` (BatchNorm1d buffers)."""

    def __init__(self, fin: int, fout: int, dropout_p: float, relu: bool):
        super().__init__()
        self.module_0 = _GCNConvParams(fin, fout)
        self.module_1 = nn.BatchNorm1d(fout)
        self.dropout_p = float(dropout_p)
        self.relu = relu


def _update_running(bn: nn.BatchNorm1d, mean: torch.Tensor, var: torch.Tensor, n: int, sl=slice(None)) -> None:
    """torch BatchNorm train-mode bookkeeping: momentum update with the UNBIASED variance."""
    with torch.no_grad():
        bn.num_batches_tracked += 1
        if bn.momentum is None:  # torch: cumulative moving average, factor 1 / num_batches_tracked
            m = 1.0 / float(bn.num_batches_tracked)
        else:
            m = bn.momentum
        unbiased = var[sl] * (n / max(n - 1, 1))
        bn.running_mean.mul_(1 - m).add_(mean[sl], alpha=m)
        bn.running_var.mul_(1 - m).add_(unbiased, alpha=m)


# ---------------------------------------------------------------------------------------------------------
# landmark model
# ---------------------------------------------------------------------------------------------------------

class HierarchicalPatchModel(nn.Module):
    """Base landmark model: average-pooled pyramid of the embedder output as node features."""

    def __init__(self,
                 frame_size: int = 32,
                 gnn_dropout_p: float = 0.0,
                 classifier_dropout_p: float = 0.0,
                 node_embedding_dim: int = 128,
                 node_hidden_dim: int = 64,
                 num_output_channels: int = 4,
                 num_gnn_layers: int = 3,
                 num_aux_graphs: int = 4,
                 gnn_jk_mode: str = 'last',
                 classifier_hidden_dim: int = 16,
                 residual: bool = True,
                 use_coordinate_graph: bool = False,
                 output_activation: str = 'sigmoid',
                 use_connection_nodes=False,
                 use_main_graph_only=False,
                 main_graph_type: str = 'grid',
                 aux_graph_type: str = 'grid'):
        super().__init__()
        assert gnn_jk_mode in ['last', 'max', 'cat'], "Only last, max or cat jumping knowledge mode is supported."
        if output_activation not in ('sigmoid', 'logit'):
            raise TypeError(f"invalid output_activation:{output_activation}")  # reference raises a TypeError too
        # Widths.  node_embedding_dim == node_hidden_dim == 128, classifier_hidden_dim == 32, 4 channels (default.yml)
        # run the tensor-core kernels; other widths (e.g. the reference constructor defaults 64 / 16,
        # src/core/models.py:290-296) run the same graph / BatchNorm / loss kernels with the generic fp32 transforms
        # (eg_linear_fwd / eg_linear_wgrad) and the CSR aggregation (eg_gcn_aggregate).
        self._wide = (node_embedding_dim == F and node_hidden_dim == F and classifier_hidden_dim == 32 and
                      num_output_channels == 4)
        if not self._wide:
            if node_embedding_dim != F:
                raise NotImplementedError(f"node_embedding_dim must be {F}: the pyramid / packing kernels move "
                                          f"{F}-wide node rows (got {node_embedding_dim})")
            if node_hidden_dim not in (64, 128):
                raise NotImplementedError(f"node_hidden_dim must be 64 or 128 (eg_gcn_aggregate / eg_bn_act widths); "
                                          f"got {node_hidden_dim}")
            if classifier_hidden_dim not in (8, 16, 32, 64, 128):
                raise NotImplementedError("classifier_hidden_dim must be one of 8, 16, 32, 64, 128 (eg_bn_act widths); "
                                          f"got {classifier_hidden_dim}")
            if use_coordinate_graph:
                raise NotImplementedError("use_coordinate_graph is built for node_hidden_dim == 128 and "
                                          "classifier_hidden_dim == 32 (eg_coord_update kernels)")
        if gnn_jk_mode == 'cat':
            raise NotImplementedError("gnn_jk_mode='cat' feeds (L+1)*128 features into Linear(128, .) and fails in "
                                      "the reference as well (src/core/models.py:364,479-482)")

        self.gnn_layers = nn.ModuleList(
            _GNNBlock(node_embedding_dim if i == 0 else node_hidden_dim, node_hidden_dim, gnn_dropout_p,
                      relu=(i != num_gnn_layers - 1)) for i in range(num_gnn_layers))
        # coordinate regressors, one per GNN layer (src/core/models.py:337-350); same Sequential indices -> same keys
        self.node_coordinate_mlp = nn.ModuleList()
        if use_coordinate_graph:
            ch = classifier_hidden_dim
            for _ in range(num_gnn_layers):
                self.node_coordinate_mlp.append(nn.Sequential(
                    nn.Linear(node_hidden_dim + 8, ch), nn.BatchNorm1d(ch), nn.ReLU(inplace=True),
                    nn.Dropout(p=classifier_dropout_p), nn.Linear(ch, ch // 2), nn.BatchNorm1d(ch // 2),
                    nn.ReLU(inplace=True), nn.Dropout(p=classifier_dropout_p), nn.Linear(ch // 2, 2), nn.Identity()))
        self.output_activation = output_activation
        last = nn.Sigmoid() if output_activation == 'sigmoid' else nn.Identity()
        h = classifier_hidden_dim
        # never called as modules: they only hold the parameters under the reference's key names
        self.node_classifiers = nn.ModuleList(
            nn.Sequential(nn.Linear(node_hidden_dim, h), nn.BatchNorm1d(h), nn.ReLU(inplace=True),
                          nn.Dropout(p=classifier_dropout_p), nn.Linear(h, h // 2), nn.BatchNorm1d(h // 2),
                          nn.ReLU(inplace=True), nn.Dropout(p=classifier_dropout_p), nn.Linear(h // 2, 1), last)
            for _ in range(num_output_channels))
        self.jk = None  # 'max' is evaluated inline; kept for attribute parity with the reference

        self.frame_size = frame_size
        self.residual = residual
        self.num_gnn_layers = num_gnn_layers
        self.node_embedding_dim = node_embedding_dim
        self.num_aux_graphs = num_aux_graphs
        self.use_coordinate_graph = use_coordinate_graph
        self.use_connection_nodes = use_connection_nodes
        self.use_main_graph_only = use_main_graph_only
        self.gnn_jk_mode = gnn_jk_mode
        self.classifier_dropout_p = float(classifier_dropout_p)
        self.graph_spec = HierGraphSpec(frame_size=frame_size, num_aux_graphs=num_aux_graphs,
                                        use_main_graph_only=bool(use_main_graph_only),
                                        use_coordinate_graph=bool(use_coordinate_graph),
                                        use_connection_nodes=bool(use_connection_nodes),
                                        main_graph_type=main_graph_type, aux_graph_type=aux_graph_type)
        self._step = 0              # training forwards so far (advances the dropout stream; not part of state_dict,
        self.dropout_seed = 0x5EED  # like torch's own RNG state, which the reference does not checkpoint either)
        self._dropout_stream = None

    # -- node features ---------------------------------------------------------------------------------------
    def pyramid(self, x: torch.Tensor) -> List[torch.Tensor]:
        """[B,128,S,S] embedder output -> list of level maps [B,128,s_l,s_l] (src/core/models.py:512-524)."""
        maps = [] if self.use_main_graph_only else [TF.adaptive_avg_pool2d(x, 2 ** k)
                                                    for k in range(1, self.num_aux_graphs + 1)]
        return maps + [x]

    def connection_rows(self, maps: List[torch.Tensor]) -> torch.Tensor:
        """[B, naux+1, 128]: the frame mean repeated (base variant, src/core/models.py:531-534)."""
        return maps[-1].mean(dim=(2, 3)).unsqueeze(1).repeat(1, self.num_aux_graphs + 1, 1)

    def create_node_pixels(self, x: torch.Tensor, graph: DeviceGraph, node_coords=None) -> torch.Tensor:
        maps = self.pyramid(x)
        head = self.connection_rows(maps) if graph.meta.first_pixel_node else None
        feats = ops.PackNodes.apply(graph, head, None, *maps)
        return self.sample_coordinate_rows(feats, graph, node_coords)

    def sample_coordinate_rows(self, feats: torch.Tensor, graph: DeviceGraph, node_coords) -> torch.Tensor:
        """Coordinate nodes start as the bilinear sample of the frame's main-level features at the initial
        coordinates (src/core/models.py:526-527, 743-744): written into the packed tensor in place."""
        if not self.use_coordinate_graph:
            return feats
        batch = feats.shape[0] // graph.meta.num_nodes
        return ops.CoordSample.apply(feats, node_coords, graph, batch, self.frame_size)

    # -- forward ---------------------------------------------------------------------------------------------
    def forward(self, data_batch=None, x: torch.Tensor = None, node_coords: torch.Tensor = None,
                edge_index: torch.Tensor = None, node_type=None, batch_idx: torch.Tensor = None):
        # The kernels launch on the CURRENT CUDA device: make it the inputs' device for the whole forward, so a
        # module living on cuda:1 of a multi-GPU process works (the backward runs on autograd's per-device thread,
        # which already does this).  One process per GPU (torchrun) remains the supported multi-GPU route.
        ref = x if x is not None else (data_batch[0].x if isinstance(data_batch, (list, tuple)) else
                                       getattr(data_batch, "x", None))
        if ref is None or not ref.is_cuda:
            raise EchogladError("the landmark module needs CUDA inputs: echoglad_b200 has no CPU fallback")
        with torch.cuda.device(ref.device):
            return self._forward(data_batch, x, node_coords, edge_index, node_type, batch_idx)

    def _forward(self, data_batch, x, node_coords, edge_index, node_type, batch_idx):
        if data_batch is not None:
            # the reference's two data_batch forms (src/core/models.py:408-413, src/engine.py:243-248): a collated
            # PyG `Batch` (attributes x / edge_index / batch / node_type / node_coords), or -- on the multi-GPU route,
            # where PyG's DataParallel would have collated it -- the raw list of per-frame `Data` objects
            if isinstance(data_batch, (list, tuple)):
                x = torch.cat([d.x for d in data_batch], dim=0)
                graph1 = DeviceGraph.get(self.graph_spec, x.device)
                graph1.validate_edge_index_once(getattr(data_batch[0], "edge_index", None), 1)
                edge_index = None
                if self.use_coordinate_graph and node_coords is None:
                    node_coords = torch.cat([d.node_coords.reshape(-1, 2) for d in data_batch], dim=0)
            else:
                x, edge_index = data_batch.x, data_batch.edge_index
                if self.use_coordinate_graph and node_coords is None:
                    node_coords = data_batch.node_coords
        if x is None or not x.is_cuda:
            raise EchogladError("the landmark module needs CUDA inputs: echoglad_b200 has no CPU fallback")
        batch = x.shape[0]
        graph = DeviceGraph.get(self.graph_spec, x.device)
        graph.validate_edge_index_once(edge_index, batch)

        coords = None
        if self.use_coordinate_graph:
            if node_coords is None:
                raise EchogladError("use_coordinate_graph=True needs node_coords [4*B, 2]")
            coords = node_coords.to(torch.float32).reshape(batch, 4, 2)  # not modified in place (the reference does)
            feats = self.create_node_pixels(x.float(), graph, coords)
            h, coords = self.gnn_stack(feats, graph, batch, coords)
        else:
            feats = self.create_node_pixels(x.float(), graph)
            h = self.gnn_stack(feats, graph, batch)
        if graph.meta.num_pixel_nodes != graph.meta.num_nodes:  # drop connection / coordinate rows
            a = graph.meta.first_pixel_node
            h = h.view(batch, graph.meta.num_nodes, h.shape[1])[:, a:a + graph.meta.num_pixel_nodes].reshape(-1, h.shape[1])
        out = self.classify(h)
        if self.training:
            self._step += 1
        return out.squeeze(1), (None if coords is None else coords.reshape(batch * 4, 2))

    def _stream_id(self) -> int:
        """Distinguishes dropout streams of different runs / replicas: torch's global seed (the engine sets it from
        `train.seed`, src/engine.py:31-33) and the data-parallel rank, so ranks do not apply the same masks."""
        if self._dropout_stream is None:
            rank = 0
            if torch.distributed.is_available() and torch.distributed.is_initialized():
                rank = torch.distributed.get_rank()
            self._dropout_stream = (torch.initial_seed() * 0x2545F4914F6CDD1D + rank * 0x9E3779B97F4A7C15) & ((1 << 63) - 1)
        return self._dropout_stream

    def _seed(self, salt: int) -> int:
        return (self.dropout_seed * 0x9E3779B1 + self._stream_id() + self._step * 1000003 + salt * 7919) & 0x7FFFFFFFFFFFFFFF

    def update_coordinates(self, i: int, y: torch.Tensor, coords: torch.Tensor, graph: DeviceGraph, batch: int):
        """Coordinate update after GNN layer i (src/core/models.py:438-473), one kernel (eg_coord_update_fwd):
        relative-position features + the coordinate nodes' embeddings -> node_coordinate_mlp[i] -> delta; clamp; the
        coordinate nodes' embeddings are re-sampled from the main-level rows of `y` at the new coordinates and
        written back in place (4 rows per frame).  No host sync (the reference does three np.where per layer)."""
        mlp = self.node_coordinate_mlp[i]
        bn1, bn2 = mlp[1], mlp[5]
        y, new, m1, v1, m2, v2 = ops.CoordUpdate.apply(
            y, coords, graph, batch, self.frame_size, mlp[0].weight, mlp[0].bias, bn1.weight, bn1.bias,
            bn1.running_mean, bn1.running_var, mlp[4].weight, mlp[4].bias, bn2.weight, bn2.bias, bn2.running_mean,
            bn2.running_var, mlp[8].weight, mlp[8].bias, self.training, bn1.eps,
            self.classifier_dropout_p if self.training else 0.0, self._seed(201 + 2 * i))
        if self.training:
            _update_running(bn1, m1, v1, 4 * batch)
            _update_running(bn2, m2, v2, 4 * batch)
        return y, new.view(batch, 4, 2)

    def gnn_stack(self, feats: torch.Tensor, graph: DeviceGraph, batch: int, coords: torch.Tensor = None):
        hidden = [feats]
        rows = feats.shape[0]
        for i, blk in enumerate(self.gnn_layers):
            bn = blk.module_1
            use_batch_stats = self.training or not bn.track_running_stats
            if self._wide:
                y, mean, var = ops.GCNLayer.apply(
                    graph, batch, hidden[i], blk.module_0.lin.weight, blk.module_0.bias, bn.weight, bn.bias,
                    bn.running_mean, bn.running_var, use_batch_stats, bn.eps,
                    blk.dropout_p if self.training else 0.0, self._seed(i), blk.relu, bool(self.residual))
            else:
                # generic widths: (A_hat X) W^T + b with the CSR aggregation and the fp32 transform, then the same
                # BatchNorm / Dropout / ReLU kernels; the residual only where the widths agree (src/core/models.py:434)
                w = blk.module_0.lin.weight
                z = ops.LinearGeneric.apply(ops.Aggregate.apply(graph, batch, hidden[i]), w, blk.module_0.bias)
                res = hidden[i] if (self.residual and w.shape[0] == w.shape[1]) else None
                y, mean, var = ops.BNAct.apply(z, bn.weight, bn.bias, bn.running_mean, bn.running_var, use_batch_stats,
                                               bn.eps, blk.dropout_p if self.training else 0.0, self._seed(i), blk.relu,
                                               res)
            if self.training and bn.track_running_stats:
                _update_running(bn, mean, var, rows)
            if coords is not None:
                y, coords = self.update_coordinates(i, y, coords, graph, batch)
            hidden.append(y)
        if self.gnn_jk_mode == 'max':
            out = hidden[0]
            for t in hidden[1:]:
                out = torch.maximum(out, t)
        else:
            out = hidden[-1]
        return out if coords is None else (out, coords)

    def _classify_generic(self, h: torch.Tensor) -> torch.Tensor:
        """Node classifiers of any width: per head Linear -> BN -> ReLU -> Dropout -> Linear -> BN -> ReLU -> Dropout
        -> Linear (src/core/models.py:363-377) on the generic transforms + the BatchNorm / activation kernels."""
        rows = h.shape[0]
        p = self.classifier_dropout_p if self.training else 0.0
        outs = []
        for k, c in enumerate(self.node_classifiers):
            a = h
            for j, (lin, bn) in enumerate(((c[0], c[1]), (c[4], c[5]))):
                z = ops.LinearGeneric.apply(a, lin.weight, lin.bias)
                use_batch_stats = self.training or not bn.track_running_stats
                a, mean, var = ops.BNAct.apply(z, bn.weight, bn.bias, bn.running_mean, bn.running_var, use_batch_stats,
                                               bn.eps, p, self._seed(101 + 10 * k + j), True, None)
                if self.training and bn.track_running_stats:
                    _update_running(bn, mean, var, rows)
            outs.append(ops.LinearGeneric.apply(a, c[8].weight, c[8].bias))
        out = torch.cat(outs, dim=1)
        return torch.sigmoid(out) if self.output_activation == 'sigmoid' else out

    def classify(self, h: torch.Tensor) -> torch.Tensor:
        if not self._wide:
            return self._classify_generic(h)
        clf = self.node_classifiers
        cat = torch.cat
        w1 = cat([c[0].weight for c in clf]); b1 = cat([c[0].bias for c in clf])
        g1 = cat([c[1].weight for c in clf]); be1 = cat([c[1].bias for c in clf])
        w2 = torch.stack([c[4].weight for c in clf]); b2 = torch.stack([c[4].bias for c in clf])
        g2 = cat([c[5].weight for c in clf]); be2 = cat([c[5].bias for c in clf])
        w3 = cat([c[8].weight for c in clf]); b3 = cat([c[8].bias for c in clf])
        with torch.no_grad():
            m1 = cat([c[1].running_mean for c in clf]); v1 = cat([c[1].running_var for c in clf])
            m2 = cat([c[5].running_mean for c in clf]); v2 = cat([c[5].running_var for c in clf])
        out, bm1, bv1, bm2, bv2 = ops.ClassifierHeads.apply(
            h, w1, b1, g1, be1, m1, v1, w2, b2, g2, be2, m2, v2, w3, b3, self.training, BN_EPS,
            self.classifier_dropout_p if self.training else 0.0, self._seed(101),
            self.output_activation == 'sigmoid')
        if self.training:
            rows = h.shape[0]
            for k, c in enumerate(clf):
                _update_running(c[1], bm1, bv1, rows, slice(32 * k, 32 * k + 32))
                _update_running(c[5], bm2, bv2, rows, slice(16 * k, 16 * k + 16))
        return out


class _BNTrain2d(torch.autograd.Function):
    """Train-mode BatchNorm2d written as a few full-width PyTorch passes.  cuDNN's spatial BN kernels
    (bn_fw_tr_1C11 / bn_bw_1C11) run ONE CTA per channel, so on the 4..8-channel full-resolution levels of
    this UNet they use 8 of 148 SMs: measured 43 ms of a 157 ms batch-64 step (profiles/r01a_launches.txt).
    Out of the hot-path scope (the pyramid stays PyTorch); same arithmetic as nn.BatchNorm2d."""

    @staticmethod
    def forward(ctx, x, weight, bias, eps):
        var, mean = torch.var_mean(x, dim=(0, 2, 3), unbiased=False)
        invstd = torch.rsqrt(var + eps)
        scale = weight * invstd
        y = torch.addcmul((bias - mean * scale).view(1, -1, 1, 1), x, scale.view(1, -1, 1, 1))
        ctx.save_for_backward(x, weight, mean, invstd)
        ctx.mark_non_differentiable(mean, var)
        return y, mean, var

    @staticmethod
    def backward(ctx, dy, _dm, _dv):
        x, weight, mean, invstd = ctx.saved_tensors
        N, C, H, W = x.shape
        n = N * H * W
        dy = dy.contiguous()
        # 3 + 5 tensor passes instead of the 20 of the textbook composition (xhat materialised, five in-place updates):
        # sum dy; sum dy x as N*C dot products (no product tensor); dx = a dy + b x + c per channel in two kernels
        dbeta = dy.sum(dim=(0, 2, 3))
        s_dyx = torch.bmm(dy.view(N * C, 1, H * W), x.view(N * C, H * W, 1)).view(N, C).sum(dim=0)
        dgamma = (s_dyx - mean * dbeta) * invstd
        dx = None
        if ctx.needs_input_grad[0]:
            a = weight * invstd
            b = -a * invstd * dgamma / n
            c = -a * dbeta / n - b * mean
            dx = torch.addcmul(c.view(1, -1, 1, 1), dy, a.view(1, -1, 1, 1))
            dx.addcmul_(x, b.view(1, -1, 1, 1))
        return dx, dgamma, dbeta, None


class _BatchNorm2d(nn.BatchNorm2d):
    """nn.BatchNorm2d (same parameters / buffers / state_dict keys) whose CUDA train-mode path on maps with at most 64
    channels is `ops.BN2dTrain` (two launches per direction; `relu_in=True` folds the preceding ReLU in, `pre_bias` the
    bias of a convolution that was run without it), or `_BNTrain2d` where the plane size is not a multiple of 4.
    `forward(x, relu_in=True, pre_bias=b)` = BN(relu(x + b))."""

    def fused_ok(self, x: torch.Tensor) -> bool:
        """Whether `forward` on a map like `x` (device, dtype, channels, plane size) takes the two-launch kernels."""
        plane = x.shape[2] * x.shape[3]
        units = x.shape[0] * self.num_features * -(-plane // 4096)  # 16 KB work units: one double each of scratch
        return bool(self.training and x.is_cuda and self.track_running_stats and self.momentum is not None
                    and self.num_features <= 64 and x.dtype == torch.float32 and plane % 4 == 0
                    and units * 8 <= ops.WORKSPACE_BYTES // 2)

    def forward(self, x, relu_in: bool = False, pre_bias: Optional[torch.Tensor] = None):
        if self.fused_ok(x):
            y, mean, var = ops.BN2dTrain.apply(x, self.weight, self.bias, self.eps, relu_in, pre_bias)
        else:
            if pre_bias is not None:
                x = x + pre_bias.view(1, -1, 1, 1)
            if relu_in:
                x = TF.relu(x)
            if not (self.training and x.is_cuda and self.track_running_stats and self.momentum is not None
                    and x.shape[1] <= 16):
                return super().forward(x)
            y, mean, var = _BNTrain2d.apply(x, self.weight, self.bias, self.eps)
        with torch.no_grad():
            n = x.numel() // x.shape[1]
            self.running_mean.mul_(1 - self.momentum).add_(mean, alpha=self.momentum)
            self.running_var.mul_(1 - self.momentum).add_(var * (n / max(n - 1, 1)), alpha=self.momentum)
            self.num_batches_tracked += 1
        return y


def _conv_relu_bn(conv: nn.Conv2d, bn: _BatchNorm2d, x: torch.Tensor, relu: bool = True) -> torch.Tensor:
    """bn(relu(conv(x))) (`relu=False`: bn(conv(x))).  Where the BatchNorm takes its two-launch kernels the convolution
    runs WITHOUT its bias, which the BatchNorm kernels add on load (and whose gradient they return): no bias-add pass, no
    bias-gradient reduction."""
    if conv.bias is not None and conv.padding_mode == "zeros" and bn.fused_ok(x) \
            and (x.shape[2] * x.shape[3]) == _conv_out_plane(conv, x):
        z = TF.conv2d(x, conv.weight, None, conv.stride, conv.padding, conv.dilation, conv.groups)
        return bn(z, relu_in=relu, pre_bias=conv.bias)
    return bn(conv(x), relu_in=relu)


def _conv_out_plane(conv: nn.Conv2d, x: torch.Tensor) -> int:
    def out(size, k, s, p, d):
        return (size + 2 * p - d * (k - 1) - 1) // s + 1
    return out(x.shape[2], conv.kernel_size[0], conv.stride[0], conv.padding[0], conv.dilation[0]) * \
        out(x.shape[3], conv.kernel_size[1], conv.stride[1], conv.padding[1], conv.dilation[1])


class _DownConv(nn.Module):
    """conv3x3-ReLU-BN twice, then AdaptiveMaxPool to `out_size` (reference DownConv, models.py:841-856)."""

    def __init__(self, cin: int, cout: int, out_size: int):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.BN1 = _BatchNorm2d(cout)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.BN2 = _BatchNorm2d(cout)
        self.out_size = out_size

    def forward(self, x):
        x = _conv_relu_bn(self.conv1, self.BN1, x)
        x = _conv_relu_bn(self.conv2, self.BN2, x)
        return TF.adaptive_max_pool2d(x, self.out_size)


class _UpConv(nn.Module):
    """nearest upsample, conv-ReLU-BN, concat skip, conv-ReLU-BN (reference UpConv, models.py:859-876)."""

    def __init__(self, cin: int, cout: int, out_size: int):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.BN1 = _BatchNorm2d(cout)
        self.conv2 = nn.Conv2d(cin, cout, 3, padding=1)
        self.BN2 = _BatchNorm2d(cout)
        self.out_size = out_size

    def forward(self, x, skip):
        x = TF.interpolate(x, size=self.out_size)
        x = _conv_relu_bn(self.conv1, self.BN1, x)
        return _conv_relu_bn(self.conv2, self.BN2, torch.cat([x, skip], dim=1))


class UNETHierarchicalPatchModel(HierarchicalPatchModel):
    """default.yml landmark model: node features come from the decoder maps of a 7-level UNet."""

    def __init__(self, encoder_embedding_widths: list = None, encoder_embedding_dims=None, **kwargs):
        super().__init__(**kwargs)
        if encoder_embedding_widths is None:
            encoder_embedding_widths = [128, 64, 32, 16, 8, 4, 2]
        # reference quirk kept: any truthy value is overwritten by the default list (models.py:652-653)
        if encoder_embedding_dims:
            encoder_embedding_dims = [8, 16, 32, 64, 128, 256, 512]
        if self.num_aux_graphs > len(encoder_embedding_widths):
            raise TypeError(f"num_aux_graphs:{self.num_aux_graphs} is larger than total number of usable "
                            f"intermediate layers:{len(encoder_embedding_widths)}")
        dims = list(encoder_embedding_dims)
        self.down_convs = nn.ModuleList(_DownConv(f // 2, f, encoder_embedding_widths[i])
                                        for i, f in enumerate(dims))
        up_sizes = list(reversed(encoder_embedding_widths))[1:] + [self.frame_size]
        self.up_convs = nn.ModuleList(_UpConv(f, f // 2, up_sizes[i]) for i, f in enumerate(reversed(dims)))
        feats_in = list(reversed(dims)) + [dims[0] // 2]
        self.linears = nn.ModuleList(nn.Conv2d(c, self.node_embedding_dim, kernel_size=1) for c in feats_in)
        self.fuse_level_embed = True  # False: PyTorch conv1x1 + ReLU + eg_pack_nodes for every level (A/B, tests)

    def decoder_features(self, x: torch.Tensor):
        """UNet encoder / decoder (PyTorch): decoder maps of all 8 scales and the indices the graph uses."""
        skips = []
        for down in self.down_convs:
            skips.append(x)
            x = down(x)
        feats = [x]
        for up in self.up_convs:
            x = up(x, skips.pop())
            feats.append(x)
        naux = 0 if self.use_main_graph_only else self.num_aux_graphs
        return feats, list(range(naux)) + [len(feats) - 1]

    def pyramid(self, x: torch.Tensor) -> List[torch.Tensor]:
        feats, used = self.decoder_features(x)
        return [TF.relu(self.linears[i](feats[i])) for i in used]

    def create_node_pixels(self, x: torch.Tensor, graph: DeviceGraph, node_coords=None) -> torch.Tensor:
        """Fused route (no connection nodes): the raw decoder maps of ALL levels go through 1x1 conv + ReLU + packing
        in one pass per level (eg_level_embed for the two big levels, eg_linear_fwd for the small ones; SURVEY.md
        §8(f) row 1); no [B,128,s,s] map is materialised."""
        meta = graph.meta
        if meta.first_pixel_node or not self.fuse_level_embed:
            return super().create_node_pixels(x, graph, node_coords)
        feats, used = self.decoder_features(x)
        args = []
        for i in used:  # every level: raw decoder map + its 1x1 conv (tensor-core-free kernels pick by cin)
            args += [feats[i], self.linears[i].weight, self.linears[i].bias]
        fused = range(len(used))
        return self.sample_coordinate_rows(ops.EmbedPackNodes.apply(graph, tuple(fused), *args), graph, node_coords)

    def connection_rows(self, maps: List[torch.Tensor]) -> torch.Tensor:
        """One connection node per used level = its spatial mean (src/core/models.py:735-752)."""
        return torch.stack([m.mean(dim=(2, 3)) for m in maps], dim=1)


class CNN(nn.Module):
    """The default.yml embedder (`cnn`, one residual block 1 -> 4 channels).  Plain PyTorch: out of the
    hot-path scope, present so that engine-level runs are self-contained.  Same state_dict keys as the
    reference (`conv.{i}.0.{one_by_one_cnn,conv,bn}.*`, src/core/models.py:71-260)."""

    class _Block(nn.Module):
        def __init__(self, cin, cout, k, pool, p):
            super().__init__()
            self.one_by_one_cnn = nn.Conv2d(cin, cout, 1) if cin != cout else None
            self.conv = nn.Conv2d(cin, cout, k, padding=(k - 1) // 2)
            self.bn = _BatchNorm2d(cout)
            self.pool = nn.MaxPool2d(pool)
            self.pool_size = pool
            self.dropout = nn.Dropout2d(p)

        def forward(self, x):
            res = x if self.one_by_one_cnn is None else self.one_by_one_cnn(x)
            y = _conv_relu_bn(self.conv, self.bn, x, relu=False) + res
            if self.pool_size != 1:  # (MaxPool2d(1) is the identity, but costs a pass, an int64 index map and a backward)
                y = self.pool(y)
            return self.dropout(TF.relu(y))

    def __init__(self, out_channels: list, kernel_sizes: list = None, pool_sizes: list = None,
                 fc_output_dim: list = None, cnn_dropout_p: float = 0.0):
        super().__init__()
        n = len(out_channels)
        kernel_sizes = kernel_sizes or [3] * n
        pool_sizes = pool_sizes or [1] * n
        if fc_output_dim is not None:
            raise NotImplementedError("fc_output_dim is unused by default.yml")
        chans = [1] + list(out_channels)
        self.conv = nn.Sequential(*[nn.Sequential(CNN._Block(chans[i], chans[i + 1], kernel_sizes[i],
                                                             pool_sizes[i], cnn_dropout_p)) for i in range(n)])
        self.output_fc = None

    def forward(self, x):
        return self.conv(x)
