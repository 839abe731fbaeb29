"""Static hierarchical graph of EchoGLAD as a device-resident CSR (one frame; batches are block-diagonal).

Replaces `create_graphs` (reference src/core/datasets.py:375-521), PyG `from_networkx` (:258) and the
`Batch.from_data_list` edge offsets: the graph is a closed form of (frame_size, num_aux_graphs, flags),
built once per (spec, device) by `eg_graph_create`.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Dict, List, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import GraphInfo, GraphSpec, check, lib


@dataclass(frozen=True)
class HierGraphSpec:
    """The `data.*` keys of configs/default.yml:67-75 that define the graph."""
    frame_size: int = 224
    num_aux_graphs: int = 7
    use_main_graph_only: bool = False
    use_coordinate_graph: bool = False
    use_connection_nodes: bool = False
    main_graph_type: str = "grid"
    aux_graph_type: str = "grid"

    def c_spec(self) -> GraphSpec:
        for t in (self.main_graph_type, self.aux_graph_type):
            if t not in ("grid", "grid-diagonal"):
                raise ValueError(f"unknown graph type {t!r} (reference supports 'grid' / 'grid-diagonal')")
        return GraphSpec(int(self.frame_size), int(self.num_aux_graphs), int(self.use_main_graph_only),
                         int(self.use_coordinate_graph), int(self.use_connection_nodes),
                         int(self.main_graph_type == "grid-diagonal"), int(self.aux_graph_type == "grid-diagonal"))

    # ---- host-side closed forms (no GPU) -------------------------------------------------------------
    def info(self) -> "GraphMeta":
        gi = GraphInfo()
        check(lib.eg_graph_spec_info(C.byref(self.c_spec()), C.byref(gi)), "eg_graph_spec_info")
        return GraphMeta.from_c(gi)

    def host_edge_index(self, batch: int = 1) -> torch.Tensor:
        """int64[2, batch*E] in the reference's (networkx/PyG) edge order, computed on the host."""
        meta = self.info()
        out = np.empty((2, batch * meta.num_edges), dtype=np.int64)
        check(lib.eg_graph_host_edge_index(C.byref(self.c_spec()), batch, out.ctypes.data_as(C.c_void_p)),
              "eg_graph_host_edge_index")
        return torch.from_numpy(out)

    def host_node_type(self, batch: int = 1) -> np.ndarray:
        meta = self.info()
        out = np.empty(batch * meta.num_nodes, dtype=np.float64)
        check(lib.eg_graph_host_node_type(C.byref(self.c_spec()), batch, out.ctypes.data_as(C.c_void_p)),
              "eg_graph_host_node_type")
        return out


@dataclass(frozen=True)
class GraphMeta:
    num_nodes: int
    num_edges: int
    num_pixel_nodes: int
    first_pixel_node: int
    num_coord_nodes: int
    level_size: Tuple[int, ...]
    level_offset: Tuple[int, ...]
    max_degree: int
    crop_offset: int

    @staticmethod
    def from_c(gi: GraphInfo) -> "GraphMeta":
        n = gi.num_levels
        return GraphMeta(gi.num_nodes, gi.num_edges, gi.num_pixel_nodes, gi.first_pixel_node,
                         gi.num_coord_nodes, tuple(gi.level_size[:n]), tuple(gi.level_offset[:n]),
                         gi.max_degree, gi.crop_offset)

    @property
    def num_levels(self) -> int:
        return len(self.level_size)


class DeviceGraph:
    """Owns an `eg_graph*` (device CSR + normalisation weights) for one spec on one CUDA device."""

    _cache: Dict[Tuple[HierGraphSpec, int], "DeviceGraph"] = {}

    def __init__(self, spec: HierGraphSpec, device: torch.device):
        device = torch.device(device)
        if device.type != "cuda":
            raise _lib.EchogladError("echoglad_b200 has no CPU path: the graph must live on a CUDA device")
        self.spec = spec
        self.device = torch.device("cuda", device.index if device.index is not None else torch.cuda.current_device())
        handle = C.c_void_p()
        check(lib.eg_graph_create(C.byref(spec.c_spec()), self.device.index, C.byref(handle)), "eg_graph_create")
        self.handle = handle
        gi = GraphInfo()
        check(lib.eg_graph_get_info(handle, C.byref(gi)), "eg_graph_get_info")
        self.meta = GraphMeta.from_c(gi)
        self._checked_edge_index = False

    @classmethod
    def get(cls, spec: HierGraphSpec, device) -> "DeviceGraph":
        device = torch.device(device)
        idx = device.index if device.index is not None else torch.cuda.current_device()
        key = (spec, idx)
        g = cls._cache.get(key)
        if g is None:
            g = cls._cache[key] = DeviceGraph(spec, torch.device("cuda", idx))
        return g

    def __del__(self):
        h = getattr(self, "handle", None)
        if h is not None and h.value:
            try:
                lib.eg_graph_destroy(h)
            except Exception:
                pass
            self.handle = None

    # ---- exports / validation ----------------------------------------------------------------------------
    def edge_index(self, batch: int = 1) -> torch.Tensor:
        """Device int64[2, batch*E], bit-exact with the reference loader's batched edge_index."""
        out = torch.empty((2, batch * self.meta.num_edges), dtype=torch.int64, device=self.device)
        check(lib.eg_graph_export_edge_index(self.handle, batch, out.data_ptr(), _stream(self.device)),
              "eg_graph_export_edge_index")
        return out

    def csr(self):
        """(rowptr int32[N+1], col int32[nnz], w float32[nnz], dis float32[N]) copied to torch tensors."""
        ptrs = [C.c_void_p() for _ in range(4)]
        check(lib.eg_graph_csr(self.handle, *[C.byref(p) for p in ptrs]), "eg_graph_csr")
        n, nnz = self.meta.num_nodes, self.meta.num_edges + self.meta.num_nodes
        shapes = [(n + 1, torch.int32), (nnz, torch.int32), (nnz, torch.float32), (n, torch.float32)]
        outs = []
        for p, (cnt, dt) in zip(ptrs, shapes):
            t = torch.empty(cnt, dtype=dt, device=self.device)
            _memcpy_d2d(t.data_ptr(), p.value, cnt * t.element_size(), self.device)
            outs.append(t)
        return tuple(outs)

    def count_edge_index_mismatches(self, edge_index: torch.Tensor, batch: int) -> int:
        ei = edge_index.to(device=self.device, dtype=torch.int64).contiguous()
        flag = torch.zeros(1, dtype=torch.int32, device=self.device)
        check(lib.eg_graph_check_edge_index(self.handle, batch, ei.data_ptr(), ei.shape[1], flag.data_ptr(),
                                            _stream(self.device)), "eg_graph_check_edge_index")
        return int(flag.item())

    def validate_edge_index_once(self, edge_index, batch: int) -> None:
        """The drop-in module ignores the caller's edge_index (the graph is static) but checks on the first
        call that it describes the same graph the reference would have built."""
        if self._checked_edge_index or edge_index is None:
            return
        bad = self.count_edge_index_mismatches(edge_index, batch)
        if bad != 0:
            raise _lib.EchogladError(
                f"edge_index passed to the landmark module does not match the static graph for {self.spec} "
                f"({'shape mismatch' if bad < 0 else str(bad) + ' mismatching entries'})")
        self._checked_edge_index = True


def _stream(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def _memcpy_d2d(dst: int, src: int, nbytes: int, device: torch.device) -> None:
    # goes through torch so that no second CUDA runtime binding is needed: wrap src as a tensor
    src_t = _as_tensor(src, nbytes, device)
    dst_t = _as_tensor(dst, nbytes, device)
    dst_t.copy_(src_t)


def _as_tensor(ptr: int, nbytes: int, device: torch.device) -> torch.Tensor:
    class _Holder:
        pass
    h = _Holder()
    h.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}
    return torch.as_tensor(h, device=device)
