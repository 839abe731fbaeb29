"""TEST INFRASTRUCTURE ONLY — never imported by the product path (echoglad_b200/*).

Imports the *unmodified* reference sources (`/root/reference/src/core/{datasets,models,criterion}.py`)
with `sys.modules` stubs for the third-party packages that are absent from this image
(torch_geometric 2.0.2, torch_scatter, matplotlib, torchsummary, imageio).  Everything that is
not a stub below is then the reference's own code, which is what `tests/golden/make_golden.py`
runs to mint the golden vectors that pin `oracle/restated.py`.

The reference tree only exists in the build container (not on the GPU box), so nothing that runs
under `-m gpu`, `smoke()` or `bench.py` may import this module.

Stubbed pieces (restated from the torch_geometric 2.0.2 sources, which are not available offline):
  * `torch_geometric.nn.GCNConv`          -> `oracle.restated.GCNConvRestated`
  * `torch_geometric.nn.Sequential`       -> `oracle.restated.PygSequential` (children `module_{i}`)
  * `torch_geometric.nn.JumpingKnowledge` -> `oracle.restated.JumpingKnowledge`
  * `torch_geometric.utils.from_networkx` -> `oracle.restated.from_networkx_edge_index` wrapped in a namespace
  * `torch_geometric.data.Dataset`        -> empty base class
Reference call sites of these: src/core/models.py:5,329-335,382,431 ; src/core/datasets.py:9-10,258.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("ECHOGLAD_REFERENCE", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "src", "core", "models.py"))


def _install_stubs() -> None:
    from oracle import restated as R

    def mod(name: str) -> types.ModuleType:
        m = types.ModuleType(name)
        sys.modules[name] = m
        return m

    if "torch_geometric" not in sys.modules:
        tg = mod("torch_geometric")
        tg_nn = mod("torch_geometric.nn")
        tg_data = mod("torch_geometric.data")
        tg_utils = mod("torch_geometric.utils")
        tg_loader = mod("torch_geometric.loader")
        tg.nn, tg.data, tg.utils, tg.loader = tg_nn, tg_data, tg_utils, tg_loader

        tg_nn.GCNConv = R.GCNConvRestated
        tg_nn.Sequential = R.PygSequential
        tg_nn.JumpingKnowledge = R.JumpingKnowledge
        tg_nn.global_add_pool = None  # imported by the reference, never called
        tg_nn.DataParallel = None

        class _Dataset:  # torch_geometric.data.Dataset: only `super().__init__()` is used
            def __init__(self, *a, **k):
                pass

        class _Data(types.SimpleNamespace):
            pass

        def _from_networkx(G):
            ei, n = R.from_networkx_edge_index(G)
            return _Data(edge_index=ei, num_nodes=n)

        tg_data.Dataset = _Dataset
        tg_data.Data = _Data
        tg_utils.from_networkx = _from_networkx
        tg_loader.DataLoader = None
        tg_loader.DataListLoader = None

    if "matplotlib" not in sys.modules:
        try:
            importlib.import_module("matplotlib.pyplot")
        except Exception:
            mpl = mod("matplotlib")
            plt = mod("matplotlib.pyplot")
            plt.new_figure_manager = None
            mpl.pyplot = plt
    for name, attrs in (("torchsummary", {"summary": None}), ("imageio", {})):
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:
                m = mod(name)
                for k, v in attrs.items():
                    setattr(m, k, v)


def load():
    """Returns (datasets, models, criterion) modules of the reference."""
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    datasets = importlib.import_module("src.core.datasets")
    models = importlib.import_module("src.core.models")
    criterion = importlib.import_module("src.core.criterion")
    return datasets, models, criterion


def reference_graph(frame_size, num_aux_graphs, *, main_only=False, coord=False, conn=False,
                    main_type="grid", aux_type="grid"):
    """Runs the reference's own `create_graphs` (src/core/datasets.py:1441) + `from_networkx`
    (:1392) and returns (edge_index int64[2,E], node_type float64[N])."""
    import numpy as np

    datasets, _, _ = load()
    ds = object.__new__(datasets.DummyDataset)
    ds.num_aux_graphs = num_aux_graphs
    ds.frame_size = frame_size
    ds.use_coordinate_graph = coord
    ds.use_connection_nodes = conn
    ds.use_main_graph_only = main_only
    graphs, node_type = ds.create_graphs(main_type, aux_type)
    g = datasets.from_networkx(graphs)
    return g.edge_index, np.asarray(node_type, dtype=np.float64)


def reference_labels(coords, frame_size, num_aux_graphs, *, main_only=False):
    """Reference `create_node_labels` (src/core/datasets.py:1586) for a (4,2) coord array -> y[N,4]."""
    import torch

    datasets, _, _ = load()
    ds = object.__new__(datasets.DummyDataset)
    ds.num_aux_graphs = num_aux_graphs
    ds.frame_size = frame_size
    ds.use_main_graph_only = main_only
    return torch.cat([ds.create_node_labels(c) for c in coords], dim=1)
