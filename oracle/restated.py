"""TEST INFRASTRUCTURE ONLY — CPU oracle for the EchoGLAD GNN hot path.

A plain torch / numpy / networkx restatement of the reference algorithm, written in a functional
style over a `state_dict` (the checkpoint layout is the compatibility contract).
Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may
import it; `echoglad_b200/*` never does (the product fails loudly without its CUDA library).

PINNING: the reference has no tests and its message-passing arithmetic lives in the un-vendored
third-party `torch_geometric==2.0.2` / `torch_scatter==2.0.9` (README.md:42 of the reference), which
are not installable offline.  This restatement is pinned against golden vectors minted by running the
reference's own `src/core/{datasets,models,criterion}.py` in the build container through
`oracle/ref_shim.py` (graph builder, label builder, UNet, packing, classifiers and both losses are
the reference's real code; only `GCNConv`/`Sequential`/`from_networkx` are restated from the PyG
2.0.2 semantics) — see `tests/golden/make_golden.py` and `tests/test_oracle_golden.py`.  Because the
PyG kernels themselves could not be executed, the GCNConv arithmetic is "parity unpinned" against
PyG proper; it is anchored on the published GCN formula D^-1/2 (A+I) D^-1/2 X W^T + b and on a dense
known-answer test.

Every function cites the reference file:line it follows (paths relative to /root/reference).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

Tensor = torch.Tensor

# --------------------------------------------------------------------------------------------
# PyG 2.0.2 restatements (third-party, un-vendored)
# --------------------------------------------------------------------------------------------


def gcn_norm(edge_index: Tensor, num_nodes: int, dtype=torch.float32) -> Tuple[Tensor, Tensor]:
    """PyG 2.0.2 `gcn_norm(improved=False, add_self_loops=True)`: self-loops (weight 1) appended
    after all real edges, deg = scatter_add(w, col), dis = deg^-1/2 (inf -> 0),
    w' = dis[row] * w * dis[col].  Used by GCNConv at src/core/models.py:330,431."""
    loops = torch.arange(num_nodes, dtype=edge_index.dtype, device=edge_index.device)
    ei = torch.cat([edge_index, torch.stack([loops, loops])], dim=1)
    w = torch.ones(ei.shape[1], dtype=dtype, device=ei.device)
    row, col = ei[0], ei[1]
    deg = torch.zeros(num_nodes, dtype=dtype, device=ei.device).index_add_(0, col, w)
    dis = deg.pow(-0.5)
    dis = dis.masked_fill(dis == float("inf"), 0.0)
    return ei, dis[row] * w * dis[col]


def gcn_conv(x: Tensor, edge_index: Tensor, weight: Tensor, bias: Optional[Tensor]) -> Tensor:
    """PyG 2.0.2 `GCNConv.forward` (cached=False): x' = lin(x); message = w_e * x'[row];
    aggregate = scatter-add at col (flow source_to_target); + bias."""
    n = x.shape[0]
    ei, w = gcn_norm(edge_index, n, x.dtype)
    z = x @ weight.t()
    out = torch.zeros_like(z).index_add_(0, ei[1], w.unsqueeze(1) * z[ei[0]])
    return out if bias is None else out + bias


class GCNConvRestated(nn.Module):
    """Module form of `gcn_conv` with PyG's parameter names (`lin.weight` glorot, `bias` zeros)."""

    def __init__(self, in_channels: int, out_channels: int, **_):
        super().__init__()
        self.lin = nn.Linear(in_channels, out_channels, bias=False)
        self.bias = nn.Parameter(torch.zeros(out_channels))
        a = math.sqrt(6.0 / (in_channels + out_channels))
        nn.init.uniform_(self.lin.weight, -a, a)

    def forward(self, x, edge_index):
        return gcn_conv(x, edge_index, self.lin.weight, self.bias)


class PygSequential(nn.Module):
    """`torch_geometric.nn.Sequential('x, edge_index', [...])`: children are registered as
    `module_{i}`; entries given as (module, 'x, edge_index -> x') receive the edge_index too."""

    def __init__(self, input_args: str, modules: Sequence):
        super().__init__()
        self._takes_graph: List[bool] = []
        for i, entry in enumerate(modules):
            if isinstance(entry, (tuple, list)):
                m, desc = entry
                self._takes_graph.append("edge_index" in desc.split("->")[0])
            else:
                m = entry
                self._takes_graph.append(False)
            self.add_module(f"module_{i}", m)

    def forward(self, x, edge_index):
        for i, g in enumerate(self._takes_graph):
            m = getattr(self, f"module_{i}")
            x = m(x, edge_index) if g else m(x)
        return x


class JumpingKnowledge(nn.Module):
    def __init__(self, mode: str, **_):
        super().__init__()
        self.mode = mode

    def forward(self, xs):
        if self.mode == "cat":
            return torch.cat(xs, dim=-1)
        return torch.stack(xs, dim=-1).max(dim=-1)[0]


def from_networkx_edge_index(G) -> Tuple[Tensor, int]:
    """PyG 2.0.2 `from_networkx` as used at src/core/datasets.py:258 (no node/edge attributes):
    relabel to consecutive ints in node order, make directed, list the edges."""
    import networkx as nx

    G = nx.convert_node_labels_to_integers(G)
    G = G.to_directed() if not nx.is_directed(G) else G
    ei = torch.tensor(list(G.edges), dtype=torch.long).t().contiguous()
    return ei.view(2, -1), G.number_of_nodes()


# --------------------------------------------------------------------------------------------
# Static hierarchical graph (src/core/datasets.py:375-521, copies at :739, :1142, :1441)
# --------------------------------------------------------------------------------------------


def level_sizes(frame: int, naux: int, main_only: bool = False) -> List[int]:
    return [frame] if main_only else [2 ** k for k in range(1, naux + 1)] + [frame]


def build_graph_nx(frame: int, naux: int, *, main_only=False, coord=False, conn=False,
                   main_type="grid", aux_type="grid"):
    """Restates `create_graphs` with the same networkx primitives (grid_graph over disjoint label
    ranges, compose, add_edges_from) so that adjacency insertion order — and hence the edge order
    `from_networkx` emits — is the reference's.  Returns (nx.Graph, node_type float64[N])."""
    import networkx as nx

    def lattice(lo: int, p: int, diag: bool):
        g = nx.grid_graph(dim=[range(lo, lo + p), range(lo, lo + p)])  # :399 / :426
        if diag:  # :403-411 / :430-436
            r = range(lo, lo + p - 1)
            g.add_edges_from([((x, y), (x + 1, y + 1)) for x in r for y in r]
                             + [((x + 1, y), (x, y + 1)) for x in r for y in r])
        return g

    def fan_out(parents: np.ndarray, children: np.ndarray):
        # parent (x, y) <-> the 2x2 block of children, row-major  (:471-493, :495-521)
        out = []
        for x in range(parents.shape[0]):
            for y in range(parents.shape[1]):
                blk = children[2 * x:2 * x + 2, 2 * y:2 * y + 2].reshape(-1, 2)
                out += [(tuple(parents[x, y]), tuple(c)) for c in blk]
        return out

    pieces, extra, types = [], [], []
    cursor = 0
    grids = []  # lattices only (no connection graph), coarse -> fine -> main
    if not main_only:
        if conn:  # :386-390
            k = nx.complete_graph(range(naux + 1))
            pieces.append(k)
            types.append(np.full(k.number_of_nodes(), 2.0))
            cursor = k.number_of_nodes()
        for lvl in range(1, naux + 1):
            p = 2 ** lvl
            g = lattice(cursor, p, aux_type == "grid-diagonal")
            pieces.append(g)
            grids.append(g)
            types.append(np.zeros(p * p))
            cursor += p  # == last node's last coordinate + 1  (:419)
        for lvl in range(1, naux):  # :421-422
            p = 2 ** lvl
            extra += fan_out(np.array(grids[lvl - 1].nodes).reshape(p, p, 2),
                             np.array(grids[lvl].nodes).reshape(2 * p, 2 * p, 2))
    main = lattice(cursor, frame, main_type == "grid-diagonal")
    pieces.append(main)
    grids.append(main)
    types.append(np.zeros(frame * frame))
    cursor += frame
    if not main_only:
        p = 2 ** naux
        src = np.array(grids[naux - 1].nodes).reshape(p, p, 2)
        c = (p - frame // 2) // 2  # centre crop (:502); negative c slices like python does
        src = src[c:c + frame // 2, c:c + frame // 2]
        extra += fan_out(src, np.array(main.nodes).reshape(frame, frame, 2))
        if conn:  # :448-452 — connection node g-1 <-> every node of aux level g, g=1..naux-1
            for lvl in range(1, naux):
                extra += [(lvl - 1, v) for v in list(pieces[lvl])]
        if coord:  # :455-460 — K4, not connected to anything else
            k = nx.complete_graph(range(cursor, cursor + 4))
            pieces.append(k)
            types.append(np.ones(4))
    whole = pieces[0]
    for g in pieces[1:]:
        whole = nx.compose(whole, g)  # :463-464
    whole.add_edges_from(extra)  # :467
    return whole, np.concatenate(types)


def build_edge_index(frame: int, naux: int, **kw) -> Tuple[Tensor, np.ndarray]:
    g, node_type = build_graph_nx(frame, naux, **kw)
    ei, n = from_networkx_edge_index(g)
    assert n == node_type.shape[0]
    return ei, node_type


def batch_edge_index(edge_index: Tensor, num_nodes: int, batch: int) -> Tensor:
    """PyG `Batch.from_data_list`: block-diagonal offsets by cumulative node count."""
    return torch.cat([edge_index + b * num_nodes for b in range(batch)], dim=1)


# --------------------------------------------------------------------------------------------
# Labels (src/core/datasets.py:523-549) and DummyDataset inputs (:1381-1439)
# --------------------------------------------------------------------------------------------


def node_labels(coords: np.ndarray, frame: int, naux: int, main_only: bool = False) -> Tensor:
    """coords int[(L,2)] in (h, w) -> y float32[N, L]; one-hot per level via np.digitize into 2^k
    bins over [0, frame], then the pixel-level one-hot.  numpy negative-index wrap-around and the
    IndexError for coordinate == frame are inherited deliberately."""
    cols = []
    for hw in np.asarray(coords):
        parts = []
        if not main_only:
            for k in range(1, naux + 1):
                edges = np.linspace(0, frame, 2 ** k + 1)
                ij = np.digitize(hw, bins=edges) - 1
                m = np.zeros((2 ** k, 2 ** k))
                m[ij[0], ij[1]] = 1.0
                parts.append(m.reshape(-1))
        m = np.zeros((frame, frame))
        m[hw[0], hw[1]] = 1.0
        parts.append(m.reshape(-1))
        cols.append(np.concatenate(parts))
    return torch.tensor(np.stack(cols, axis=1), dtype=torch.float32)


def synthetic_batch(batch: int, frame: int, naux: int, *, main_only=False, seed=200):
    """DummyDataset-style inputs (src/core/datasets.py:1381-1415): randn frames, 4 integer (h, w)
    landmarks per frame uniform in [0, frame-1], valid = 1.  Returns frames[B,1,S,S], coords
    int64[B,4,2], y[B*N,4], valid[B*N,4]."""
    g = torch.Generator().manual_seed(seed)
    frames = torch.randn(batch, 1, frame, frame, generator=g)
    rng = np.random.default_rng(seed)
    coords = rng.integers(0, frame, size=(batch, 4, 2))
    y = torch.cat([node_labels(c, frame, naux, main_only) for c in coords], dim=0)
    return frames, torch.from_numpy(coords), y, torch.ones_like(y)


# --------------------------------------------------------------------------------------------
# Model (functional over the reference state_dict layout)
# --------------------------------------------------------------------------------------------


class Cfg:
    """Constructor kwargs of the reference landmark module (src/core/models.py:286-301,644-647)."""

    def __init__(self, **kw):
        d = dict(frame_size=224, gnn_dropout_p=0.5, classifier_dropout_p=0.5, node_embedding_dim=128,
                 node_hidden_dim=128, num_output_channels=4, num_gnn_layers=3, num_aux_graphs=7,
                 gnn_jk_mode="last", classifier_hidden_dim=32, residual=True,
                 use_coordinate_graph=False, output_activation="logit", use_connection_nodes=False,
                 use_main_graph_only=False, variant="unet")
        d.update(kw)
        self.__dict__.update(d)


def _bn(sd, pfx, x, training, momentum=0.1, eps=1e-5):
    return F.batch_norm(x, sd[pfx + "running_mean"], sd[pfx + "running_var"], sd[pfx + "weight"],
                        sd[pfx + "bias"], training, momentum, eps)


def _drop(x, p, training, masks, key):
    if masks is not None and key in masks:  # externally supplied keep-mask (already scaled)
        return x * masks[key]
    return F.dropout(x, p, training)


def _relu(x, masks, key):
    """ReLU; when the caller supplies the sign pattern of another implementation (masks['relu:'+key], bool),
    that pattern is imposed instead, so both sides differentiate the SAME piecewise-linear function (a
    pre-activation within rounding of 0 otherwise flips between any two fp32 implementations and changes the
    gradient of its whole 2-hop neighbourhood).  The largest |x| at a disagreeing position is recorded in
    masks['relu_margin'] so the test can check that every disagreement is a rounding-level one."""
    if masks is not None and ("relu:" + key) in masks:
        m = masks["relu:" + key]
        bad = (x.detach() > 0) != m
        if bool(bad.any()):
            masks.setdefault("relu_margin", []).append((key, int(bad.sum()), float(x.detach()[bad].abs().max())))
        return x * m.to(x.dtype)
    return F.relu(x)


def embedder_forward(sd: Dict[str, Tensor], frames: Tensor, training: bool, dropout_p: float = 0.0):
    """default.yml embedder: one CNNResBlock 1->C (src/core/models.py:137-158): conv3x3 -> BN ->
    + 1x1 skip -> MaxPool(1) -> ReLU -> Dropout2d."""
    p = "conv.0.0."
    res = F.conv2d(frames, sd[p + "one_by_one_cnn.weight"], sd[p + "one_by_one_cnn.bias"])
    k = sd[p + "conv.weight"].shape[-1]
    x = F.conv2d(frames, sd[p + "conv.weight"], sd[p + "conv.bias"], padding=(k - 1) // 2)
    x = F.relu(_bn(sd, p + "bn.", x, training) + res)
    return F.dropout2d(x, dropout_p, training)


def unet_pyramid(sd, cfg: Cfg, x: Tensor, training: bool,
                 widths=(128, 64, 32, 16, 8, 4, 2)) -> List[Tensor]:
    """src/core/models.py:693-710 with DownConv/UpConv (:841-876): 7 encoder stages (conv-relu-BN x2
    then AdaptiveMaxPool to `widths[i]`), 7 decoder stages (nearest upsample, conv1-relu-BN, concat
    skip, conv2-relu-BN), then a 1x1 conv + ReLU on each of the 8 decoder maps."""
    def conv(pfx, t):
        return F.relu(F.conv2d(t, sd[pfx + "weight"], sd[pfx + "bias"], padding=1))

    skips = []
    for i, w in enumerate(widths):
        skips.append(x)
        p = f"down_convs.{i}."
        x = _bn(sd, p + "BN1.", conv(p + "conv1.", x), training)
        x = _bn(sd, p + "BN2.", conv(p + "conv2.", x), training)
        x = F.adaptive_max_pool2d(x, w)
    feats = [x]
    up_sizes = list(reversed(widths))[1:] + [cfg.frame_size]
    for i, s in enumerate(up_sizes):
        p = f"up_convs.{i}."
        x = F.interpolate(x, size=s)  # nn.Upsample default: nearest
        x = _bn(sd, p + "BN1.", conv(p + "conv1.", x), training)
        x = torch.cat([x, skips.pop()], dim=1)
        x = _bn(sd, p + "BN2.", conv(p + "conv2.", x), training)
        feats.append(x)
    return [F.relu(F.conv2d(f, sd[f"linears.{i}.weight"], sd[f"linears.{i}.bias"]))
            for i, f in enumerate(feats)]


def avgpool_pyramid(cfg: Cfg, x: Tensor) -> List[Tensor]:
    """Base `hierarchicalpatch` features (src/core/models.py:512-524): adaptive average pools of the
    embedder output to 2^k, plus the full-resolution map."""
    maps = [] if cfg.use_main_graph_only else [F.adaptive_avg_pool2d(x, 2 ** k)
                                                for k in range(1, cfg.num_aux_graphs + 1)]
    return maps + [x]


def bilinear_tent(coords: Tensor, fmap: Tensor) -> Tensor:
    """`bilinear_interpolation` (src/core/models.py:539-553): dense tent weights over the whole map.
    coords [4,2] (h, w); fmap [C,S,S] -> [4,C]."""
    s = fmap.shape[-1]
    grid = torch.arange(s, device=coords.device, dtype=coords.dtype)
    wh = F.relu(1 - (coords[:, 0:1] - grid).abs())  # [4,S]
    ww = F.relu(1 - (coords[:, 1:2] - grid).abs())
    return torch.einsum("ph,pw,chw->pc", wh, ww, fmap)


def pack_nodes(cfg: Cfg, maps: List[Tensor], node_coords: Optional[Tensor] = None) -> Tensor:
    """Node-major packing (src/core/models.py:722-756): per frame [connection nodes][aux1..auxn]
    [main][coordinate nodes], each level row-major (h, w), features = channels."""
    B = maps[0].shape[0]
    naux = 0 if cfg.use_main_graph_only else cfg.num_aux_graphs
    use = (maps[:naux] + [maps[-1]])
    rows = []
    for b in range(B):
        per = [m[b].permute(1, 2, 0).reshape(-1, m.shape[1]) for m in use]
        if cfg.use_connection_nodes and not cfg.use_main_graph_only:
            if cfg.variant == "unet":  # :735-752 — one connection node per used level = its spatial mean
                per = [torch.stack([m[b].mean(dim=(1, 2)) for m in use])] + per
            else:  # base variant (:531-534): the frame mean repeated naux+1 times
                per = [maps[-1][b].mean(dim=(1, 2)).unsqueeze(0).repeat(naux + 1, 1)] + per
        if cfg.use_coordinate_graph:
            per.append(bilinear_tent(node_coords[b], maps[-1][b]))
        rows.append(torch.cat(per, dim=0))
    return torch.cat(rows, dim=0)


def coordinate_mlp(sd, cfg: Cfg, i: int, feats: Tensor, training: bool, masks=None) -> Tensor:
    """node_coordinate_mlp[i] (src/core/models.py:337-350): Linear(136,32)-BN-ReLU-Drop-Linear(32,16)-BN-ReLU-Drop-
    Linear(16,2).  masks (optional): externally supplied keep-masks `cmlp{i}a` [R,32] / `cmlp{i}b` [R,16]."""
    p = f"node_coordinate_mlp.{i}."
    z = F.linear(feats, sd[p + "0.weight"], sd[p + "0.bias"])
    z = _drop(F.relu(_bn(sd, p + "1.", z, training)), cfg.classifier_dropout_p, training, masks, f"cmlp{i}a")
    z = F.linear(z, sd[p + "4.weight"], sd[p + "4.bias"])
    z = _drop(F.relu(_bn(sd, p + "5.", z, training)), cfg.classifier_dropout_p, training, masks, f"cmlp{i}b")
    return F.linear(z, sd[p + "8.weight"], sd[p + "8.bias"])


def coordinate_update(sd, cfg: Cfg, i: int, h: Tensor, coords: Tensor, coord_rows: Tensor, pixel_rows: Tensor,
                      training: bool, masks=None):
    """The coordinate update after GNN layer i (src/core/models.py:438-473): h [Nt,F], coords [B,4,2] ->
    (h with the coordinate rows re-sampled, new coords)."""
    B, S = coords.shape[0], cfg.frame_size
    rel = -(coords.unsqueeze(2) - coords.unsqueeze(1))          # [B,4,4,2]: -(c_j - c_k) per frame (:441-444)
    shape_feats = rel.reshape(B * 4, 8)
    delta = coordinate_mlp(sd, cfg, i, torch.cat((h[coord_rows], shape_feats), dim=1), training, masks)
    coords = torch.clamp(coords + delta.view(B, 4, 2), min=0, max=S - 1)
    main = h[pixel_rows].view(B, -1, h.shape[1])[:, -S * S:, :].permute(0, 2, 1).reshape(B, -1, S, S)
    new = torch.cat([bilinear_tent(coords[b], main[b]) for b in range(B)], dim=0)
    return h.index_copy(0, coord_rows, new), coords


def gnn_stack(sd, cfg: Cfg, feats: Tensor, edge_index: Tensor, training: bool, masks=None,
              return_hidden: bool = False, node_coords: Optional[Tensor] = None, node_type=None):
    """src/core/models.py:425-482: L x [GCNConv -> BN -> Dropout -> ReLU|Identity] + identity residual when widths
    match; with `use_coordinate_graph` the coordinate update of :438-473 after every layer (relative-position
    features + the coordinate nodes' embeddings -> MLP -> coordinate delta, clamp, re-sample the coordinate
    nodes' embeddings from the main-level rows of h); JK last|max.  node_coords: [B,4,2] (h, w), not modified;
    with coordinates the result is (out, new_coords[B,4,2])."""
    hidden = [feats]
    L = cfg.num_gnn_layers
    coords = node_coords
    if cfg.use_coordinate_graph:
        nt = torch.as_tensor(np.asarray(node_type))
        coord_rows = torch.nonzero(nt == 1).squeeze(1)
        pixel_rows = torch.nonzero(nt == 0).squeeze(1)
    for i in range(L):
        p = f"gnn_layers.{i}."
        h = gcn_conv(hidden[i], edge_index, sd[p + "module_0.lin.weight"], sd[p + "module_0.bias"])
        h = _bn(sd, p + "module_1.", h, training)
        h = _drop(h, cfg.gnn_dropout_p, training, masks, f"gnn{i}")
        if i != L - 1:
            h = _relu(h, masks, f"gnn{i}")
        if cfg.residual and h.shape[1] == hidden[i].shape[1]:
            h = h + hidden[i]
        if cfg.use_coordinate_graph:
            h, coords = coordinate_update(sd, cfg, i, h, coords, coord_rows, pixel_rows, training, masks)
        hidden.append(h)
    if cfg.gnn_jk_mode == "max":
        out = torch.stack(hidden, dim=-1).max(dim=-1)[0]
    elif cfg.gnn_jk_mode == "cat":
        out = torch.cat(hidden, dim=-1)
    else:
        out = hidden[-1]
    if cfg.use_coordinate_graph:
        return out, coords
    return (out, hidden) if return_hidden else out


def classifiers(sd, cfg: Cfg, h: Tensor, training: bool, masks=None) -> Tensor:
    """src/core/models.py:363-377,488-490: per output channel Linear-BN-ReLU-Drop-Linear-BN-ReLU-
    Drop-Linear(-Sigmoid); outputs concatenated on dim 1."""
    outs = []
    for k in range(cfg.num_output_channels):
        p = f"node_classifiers.{k}."
        z = F.linear(h, sd[p + "0.weight"], sd[p + "0.bias"])
        z = _relu(_bn(sd, p + "1.", z, training), masks, f"clf{k}a")
        z = _drop(z, cfg.classifier_dropout_p, training, masks, f"clf{k}a")
        z = F.linear(z, sd[p + "4.weight"], sd[p + "4.bias"])
        z = _relu(_bn(sd, p + "5.", z, training), masks, f"clf{k}b")
        z = _drop(z, cfg.classifier_dropout_p, training, masks, f"clf{k}b")
        z = F.linear(z, sd[p + "8.weight"], sd[p + "8.bias"])
        outs.append(torch.sigmoid(z) if cfg.output_activation == "sigmoid" else z)
    return torch.cat(outs, dim=1)


def landmark_forward(sd, cfg: Cfg, x: Tensor, edge_index: Tensor, node_type: np.ndarray,
                     training: bool, masks=None, node_feats: Optional[Tensor] = None,
                     node_coords: Optional[Tensor] = None):
    """`HierarchicalPatchModel.forward` (src/core/models.py:394-496).  x is the embedder output [B,C,S,S];
    node_type is the batched per-node type vector; returns logits [B*N0, num_output_channels], and with
    `use_coordinate_graph` (node_coords [4B,2] or [B,4,2]) the pair (logits, coords [4B,2])."""
    if cfg.use_coordinate_graph:
        node_coords = node_coords.reshape(-1, 4, 2)
    if node_feats is None:
        maps = unet_pyramid(sd, cfg, x, training) if cfg.variant == "unet" else avgpool_pyramid(cfg, x)
        node_feats = pack_nodes(cfg, maps, node_coords)
    keep = np.where(np.asarray(node_type) == 0)[0]
    if cfg.use_coordinate_graph:
        h, coords = gnn_stack(sd, cfg, node_feats, edge_index, training, masks, node_coords=node_coords,
                              node_type=node_type)
        h = h[torch.from_numpy(keep)]
        return classifiers(sd, cfg, h, training, masks).squeeze(1), coords.reshape(-1, 2)
    h = gnn_stack(sd, cfg, node_feats, edge_index, training, masks)
    if keep.shape[0] != h.shape[0]:
        h = h[torch.from_numpy(keep)]
    return classifiers(sd, cfg, h, training, masks).squeeze(1)


def mae_loss(pred: Tensor, y: Tensor, loss_weight: float = 1.0) -> Tensor:
    """criterion 'coordinate' = MAE (src/core/criterion.py:52-64, src/builders/criterion_builder.py:40-41)."""
    return loss_weight * F.l1_loss(pred, y)


# --------------------------------------------------------------------------------------------
# Losses (src/core/criterion.py:6-34, 67-161; src/engine.py:582-600)
# --------------------------------------------------------------------------------------------


def weighted_bce_with_logits(pred: Tensor, y: Tensor, valid: Tensor, ones_weight: float = 9000.0,
                             loss_weight: float = 1.0) -> Tensor:
    """pred, y: [B, n, C]; valid: anything viewable to that.  loss_weight * sum(bce * w * valid) /
    sum(valid), w = ones_weight where y == 1 (src/core/criterion.py:13-27)."""
    loss = F.binary_cross_entropy_with_logits(pred, y, reduction="none")
    valid = valid.view(pred.shape)
    if ones_weight > 1:
        loss = torch.where(y == 1, torch.full_like(loss, ones_weight), torch.ones_like(loss)) * loss
    return loss_weight * (loss * valid).sum() / valid.sum()


def expected_landmark_mse(pred: Tensor, y: Tensor, valid: Tensor, *, batch_size: int, frame_size: int,
                          num_aux_graphs: int, use_main_graph_only: bool = False,
                          num_output_channels: int = 4, loss_weight: float = 1.0) -> Tensor:
    """src/core/criterion.py:93-151: per level, soft-argmax of softmax(pred over the level's nodes)
    vs argmax of the GT heat-map, normalised by grid size, squared, valid-weighted batch mean."""
    C = num_output_channels
    pred, y, valid = (t.reshape(batch_size, -1, C) for t in (pred, y, valid))
    total = pred.new_zeros(())
    start = 0
    for g in level_sizes(frame_size, num_aux_graphs, use_main_graph_only):
        end = start + g * g
        gt_map = y[:, start:end].view(batch_size, g, g, C)
        v = valid[:, start:end].permute(0, 2, 1).mean(dim=-1, keepdim=True)  # [B,C,1]
        nv = v.sum(dim=0, keepdim=True)
        nv = torch.where(nv == 0, torch.ones_like(nv), nv)
        gt_h = gt_map.max(dim=2)[0].max(dim=1)[1]  # argmax over h of (max over w)   [B,C]
        gt_w = gt_map.max(dim=1)[0].max(dim=1)[1]
        gt = torch.stack([gt_h, gt_w], dim=2).to(pred.dtype)
        p = torch.softmax(pred[:, start:end], dim=1).view(batch_size, g, g, C)
        ramp = torch.linspace(0, g - 1, g, device=pred.device, dtype=pred.dtype)
        e_h = (p * ramp.view(1, g, 1, 1)).sum(dim=(1, 2))
        e_w = (p * ramp.view(1, 1, g, 1)).sum(dim=(1, 2))
        e = torch.stack([e_h, e_w], dim=2)
        d = ((e / g - gt / g) ** 2) * v
        total = total + (d.sum(dim=0, keepdim=True) / nv).sum()
        start = end
    return total * loss_weight


def total_loss(logits: Tensor, y: Tensor, valid: Tensor, cfg: Cfg, batch: int, *, ones_weight=9000.0,
               bce_weight=1.0, elmse_weight=10.0) -> Dict[str, Tensor]:
    """`Engine.compute_loss` with the default.yml criteria (configs/default.yml:35-41)."""
    C = cfg.num_output_channels
    pred = logits.view(batch, -1, C)
    yy = y.view(batch, -1, C)
    out = {"WeightedBceWithLogits": weighted_bce_with_logits(pred, yy, valid, ones_weight, bce_weight),
           "ExpectedLandmarkMse": expected_landmark_mse(
               pred, yy, valid, batch_size=batch, frame_size=cfg.frame_size,
               num_aux_graphs=cfg.num_aux_graphs, use_main_graph_only=cfg.use_main_graph_only,
               num_output_channels=C, loss_weight=elmse_weight)}
    out["total"] = out["WeightedBceWithLogits"] + out["ExpectedLandmarkMse"]
    return out


# --------------------------------------------------------------------------------------------
# Random-init state_dicts in the reference layout — torch default initialisers
# --------------------------------------------------------------------------------------------


def init_landmark_state(cfg: Cfg, seed: int = 200, embed_channels: int = 4) -> Dict[str, Tensor]:
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, Tensor] = {}

    def uni(shape, bound):
        return (torch.rand(shape, generator=g) * 2 - 1) * bound

    def linear(pfx, out_f, in_f, ksz=None, bias=True):
        fan_in = in_f * (ksz * ksz if ksz else 1)
        shape = (out_f, in_f, ksz, ksz) if ksz else (out_f, in_f)
        sd[pfx + "weight"] = uni(shape, 1.0 / math.sqrt(fan_in))
        if bias:
            sd[pfx + "bias"] = uni((out_f,), 1.0 / math.sqrt(fan_in))

    def bn(pfx, n):
        sd[pfx + "weight"] = torch.ones(n) + 0.1 * uni((n,), 1.0)
        sd[pfx + "bias"] = 0.1 * uni((n,), 1.0)
        sd[pfx + "running_mean"] = 0.1 * uni((n,), 1.0)
        sd[pfx + "running_var"] = torch.ones(n) + 0.2 * uni((n,), 1.0)
        sd[pfx + "num_batches_tracked"] = torch.zeros((), dtype=torch.long)

    E, H, Ch = cfg.node_embedding_dim, cfg.node_hidden_dim, cfg.classifier_hidden_dim
    for i in range(cfg.num_gnn_layers):
        fin = E if i == 0 else H
        p = f"gnn_layers.{i}."
        sd[p + "module_0.bias"] = 0.05 * uni((H,), 1.0)
        sd[p + "module_0.lin.weight"] = uni((H, fin), math.sqrt(6.0 / (fin + H)))
        bn(p + "module_1.", H)
    for k in range(cfg.num_output_channels):
        p = f"node_classifiers.{k}."
        linear(p + "0.", Ch, H)
        bn(p + "1.", Ch)
        linear(p + "4.", Ch // 2, Ch)
        bn(p + "5.", Ch // 2)
        linear(p + "8.", 1, Ch // 2)
    if cfg.variant == "unet":
        dims = [8, 16, 32, 64, 128, 256, 512]
        for i, f in enumerate(dims):
            p = f"down_convs.{i}."
            linear(p + "conv1.", f, f // 2, 3)
            bn(p + "BN1.", f)
            linear(p + "conv2.", f, f, 3)
            bn(p + "BN2.", f)
        for i, f in enumerate(reversed(dims)):
            p = f"up_convs.{i}."
            linear(p + "conv1.", f // 2, f, 3)
            bn(p + "BN1.", f // 2)
            linear(p + "conv2.", f // 2, f, 3)
            bn(p + "BN2.", f // 2)
        for i, f in enumerate(list(reversed(dims)) + [dims[0] // 2]):
            linear(f"linears.{i}.", E, f, 1)
    if cfg.use_coordinate_graph:  # src/core/models.py:337-350; drawn last so the other tensors keep their values
        for i in range(cfg.num_gnn_layers):
            p = f"node_coordinate_mlp.{i}."
            linear(p + "0.", Ch, H + 8)
            bn(p + "1.", Ch)
            linear(p + "4.", Ch // 2, Ch)
            bn(p + "5.", Ch // 2)
            linear(p + "8.", 2, Ch // 2)
    return sd


def init_embedder_state(out_channels: int = 4, seed: int = 201) -> Dict[str, Tensor]:
    g = torch.Generator().manual_seed(seed)

    def uni(shape, bound):
        return (torch.rand(shape, generator=g) * 2 - 1) * bound

    p = "conv.0.0."
    return {p + "one_by_one_cnn.weight": uni((out_channels, 1, 1, 1), 1.0),
            p + "one_by_one_cnn.bias": uni((out_channels,), 1.0),
            p + "conv.weight": uni((out_channels, 1, 3, 3), 1 / 3.0),
            p + "conv.bias": uni((out_channels,), 1 / 3.0),
            p + "bn.weight": torch.ones(out_channels), p + "bn.bias": torch.zeros(out_channels),
            p + "bn.running_mean": torch.zeros(out_channels),
            p + "bn.running_var": torch.ones(out_channels),
            p + "bn.num_batches_tracked": torch.zeros((), dtype=torch.long)}


def clone_state(sd: Dict[str, Tensor], requires_grad: bool = False) -> Dict[str, Tensor]:
    out = {}
    for k, v in sd.items():
        t = v.detach().clone()
        if requires_grad and t.is_floating_point() and "running_" not in k:
            t.requires_grad_(True)
        out[k] = t
    return out


# ---------------------------------------------------------------------------------------------------------
# post-path metric: LandmarkExpectedCoordiantesEvaluator (src/core/evaluators.py:239-483), SURVEY.md §8(f) row 4
# ---------------------------------------------------------------------------------------------------------

def expected_coords(logits: Tensor, y: Tensor, valid: Tensor, batch: int, frame: int):
    """src/core/evaluators.py:310-348: on the main level (the last frame^2 nodes of every frame) the softmax over
    the nodes of each channel, its expected (h, w), the arg-max position of the label heat map (first maximum
    along each axis) and the mean of `valid`.  Returns preds[B,4,2], gt[B,4,2] (int64), valid_subset[B,4]."""
    c = logits.shape[-1]
    yp = logits.view(batch, -1, c).detach()[:, -frame * frame:, :]
    yt = y.view(batch, -1, c).detach()[:, -frame * frame:, :].view(batch, frame, frame, c)
    vs = valid.view(batch, -1, c)[:, -frame * frame:, :].permute(0, 2, 1).mean(dim=-1)
    max_along_w, _ = torch.max(yt, dim=-2)
    max_along_h, _ = torch.max(yt, dim=-3)
    _, gt_h = torch.max(max_along_w, dim=-2)
    _, gt_w = torch.max(max_along_h, dim=-2)
    gt = torch.cat((gt_h.unsqueeze(2), gt_w.unsqueeze(2)), dim=2)
    heat = torch.softmax(yp, dim=1).view(batch, frame, frame, c)
    line = torch.linspace(0, frame - 1, frame)
    ph = torch.sum(heat * line.view(1, -1, 1, 1), dim=(1, 2))
    pw = torch.sum(heat * line.view(1, 1, -1, 1), dim=(1, 2))
    return torch.cat((ph.unsqueeze(2), pw.unsqueeze(2)), dim=2), gt, vs


def expected_coord_metrics(preds: Tensor, gt: Tensor, valid_subset: Tensor, pix2mm_x: Tensor, pix2mm_y: Tensor) -> dict:
    """src/core/evaluators.py:350-391,395-428: per-landmark errors in mm and width MAE / MPE of one batch, from the
    [B,4,2] coordinates.  Landmark order lvid_top, lvid_bot, lvpw, ivs.  Returns the values `get_last()` reports
    plus 'valid' (4 bools) and the width dictionary."""
    def plen(x0, y0, x1, y1, mx, my):
        return torch.sqrt(((x0 - x1) * mx) ** 2 + ((y0 - y1) * my) ** 2)

    gt = gt.to(preds.dtype)
    nvs = valid_subset.sum(dim=0, keepdim=True)
    valid_flags = [(nvs[0, k] > 0).item() for k in range(4)]
    nvs = nvs.clone()
    nvs[nvs == 0] = 1
    err = plen(gt[:, :, 1], gt[:, :, 0], preds[:, :, 1], preds[:, :, 0], pix2mm_x.unsqueeze(1), pix2mm_y.unsqueeze(1))
    err = (err * valid_subset).sum(dim=0) / nvs[0]
    w = {}
    for tag, t in (("pred", preds), ("gt", gt)):
        w[f"{tag}_ivs_mm"] = plen(t[:, 3, 1], t[:, 3, 0], t[:, 0, 1], t[:, 0, 0], pix2mm_x, pix2mm_y)
        w[f"{tag}_lvid_mm"] = plen(t[:, 0, 1], t[:, 0, 0], t[:, 1, 1], t[:, 1, 0], pix2mm_x, pix2mm_y)
        w[f"{tag}_lvpw_mm"] = plen(t[:, 1, 1], t[:, 1, 0], t[:, 2, 1], t[:, 2, 0], pix2mm_x, pix2mm_y)
    out = {"lvid_top": err[0].item(), "lvid_bot": err[1].item(), "lvpw": err[2].item(), "ivs": err[3].item(),
           "valid": valid_flags, "widths": w}
    scale = {"lvid": valid_subset[:, 0] * valid_subset[:, 1] / torch.min(nvs[0, 0], nvs[0, 1]),
             "ivs": valid_subset[:, 3] / nvs[0, 3], "lvpw": valid_subset[:, 2] / nvs[0, 2]}
    for k in ("ivs", "lvid", "lvpw"):
        d = torch.abs(w[f"pred_{k}_mm"] - w[f"gt_{k}_mm"])
        out[f"{k}_w"] = (d * scale[k]).sum().item()
        out[f"{k}_mpe"] = (100 * d / w[f"gt_{k}_mm"] * scale[k]).sum().item()
    return out
