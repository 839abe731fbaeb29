"""ctypes binding of the C-ABI library (include/echoglad_b200.h).

There is no CPU fallback: if `libechoglad_b200.so` is missing or does not export a declared symbol the
import raises.  Build it with `python -m echoglad_b200.build` (nvcc, sm_100a).
"""
from __future__ import annotations

import ctypes as C
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("EG_LIB_PATH") or os.path.join(_PKG, "libechoglad_b200.so")  # override: dev builds

EG_MAX_LEVELS = 16


class GraphSpec(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "frame_size", "num_aux_graphs", "use_main_graph_only", "use_coordinate_graph",
        "use_connection_nodes", "main_diagonal", "aux_diagonal")]


class GraphInfo(C.Structure):
    _fields_ = [("num_nodes", C.c_int32), ("num_edges", C.c_int32), ("num_pixel_nodes", C.c_int32),
                ("first_pixel_node", C.c_int32), ("num_coord_nodes", C.c_int32), ("num_levels", C.c_int32),
                ("level_size", C.c_int32 * EG_MAX_LEVELS), ("level_offset", C.c_int32 * EG_MAX_LEVELS),
                ("max_degree", C.c_int32), ("crop_offset", C.c_int32)]


class ClassifierParams(C.Structure):
    """eg_classifier_params (include/echoglad_b200.h)."""
    _fields_ = [(n, C.c_void_p) for n in ("w1", "b1", "g1", "be1", "w2", "b2", "g2", "be2", "w3", "b3")] + \
               [("eps", C.c_float), ("drop_p", C.c_float), ("seed", C.c_uint64), ("batch_stats", C.c_int32),
                ("sigmoid", C.c_int32)]


class CoordMlpParams(C.Structure):
    """eg_coord_mlp_params."""
    _fields_ = [(n, C.c_void_p) for n in ("w1", "b1", "g1", "be1", "w2", "b2", "g2", "be2", "w3", "b3")] + \
               [("eps", C.c_float), ("drop_p", C.c_float), ("seed", C.c_uint64), ("batch_stats", C.c_int32),
                ("reserved", C.c_int32)]


class View(C.Structure):
    """eg_view: element (r, c) = base[(r // frame_rows) * frame_stride + (r % frame_rows) * row_stride + c * col_stride]."""
    _fields_ = [("base", C.c_void_p), ("frame_rows", C.c_int64), ("frame_stride", C.c_int64),
                ("row_stride", C.c_int64), ("col_stride", C.c_int64)]


class ClassifierGrads(C.Structure):
    """eg_classifier_grads."""
    _fields_ = [(n, C.c_void_p) for n in ("dw1", "db1", "dg1", "dbe1", "dw2", "db2", "dg2", "dbe2", "dw3", "db3")]


_P = C.c_void_p
_I = C.c_int
_L = C.c_int64
_F = C.c_float
_U64 = C.c_uint64
_SZ = C.c_size_t

# name -> (restype, argtypes); must list every symbol the header declares
SIGNATURES = {
    "eg_version": (C.c_char_p, []),
    "eg_last_error": (C.c_char_p, []),
    "eg_workspace_bytes": (_SZ, []),
    "eg_graph_create": (_I, [C.POINTER(GraphSpec), _I, C.POINTER(_P)]),
    "eg_graph_destroy": (None, [_P]),
    "eg_graph_get_info": (_I, [_P, C.POINTER(GraphInfo)]),
    "eg_graph_spec_info": (_I, [C.POINTER(GraphSpec), C.POINTER(GraphInfo)]),
    "eg_graph_csr": (_I, [_P, C.POINTER(_P), C.POINTER(_P), C.POINTER(_P), C.POINTER(_P)]),
    "eg_graph_tiles": (_I, [_P, C.POINTER(_P), C.POINTER(C.c_int32)]),
    "eg_graph_plan_check": (_I, [C.POINTER(GraphSpec), C.POINTER(C.c_int64)]),
    "eg_graph_patch_check": (_I, [C.POINTER(GraphSpec), C.POINTER(C.c_int64)]),
    "eg_gcn_plan_select": (_I, [_I]),
    "eg_graph_export_edge_index": (_I, [_P, _I, _P, _P]),
    "eg_graph_host_edge_index": (_I, [C.POINTER(GraphSpec), _I, _P]),
    "eg_graph_host_node_type": (_I, [C.POINTER(GraphSpec), _I, _P]),
    "eg_graph_check_edge_index": (_I, [_P, _I, _P, _L, _P, _P]),
    "eg_pack_nodes": (_I, [_P, _I, C.POINTER(_P), _P, _P, _P, _P]),
    "eg_pack_nodes_grad": (_I, [_P, _I, _P, C.POINTER(_P), _P, _P, _P]),
    "eg_level_embed_supported": (_I, [_P, _I, _I]),
    "eg_level_embed_fwd": (_I, [_P, _I, _I, _I, _P, _P, _P, _P, _P]),
    "eg_level_embed_bwd": (_I, [_P, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _SZ, _P]),
    "eg_gcn_aggregate": (_I, [_P, _I, _I, _P, _P, _P]),
    "eg_gcn_conv_fwd": (_I, [_P, _I, _P, _P, _P, _P, _P, _P, _P, _SZ, _P]),
    "eg_gcn_conv_bwd": (_I, [_P, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _SZ, _P]),
    "eg_gcn_layer_eval_fwd": (_I, [_P, _I, _P, _P, _P, _P, _P, _P, _P, C.c_float, _I, _I, _P, _P, _SZ, _P]),
    "eg_bn_act_fwd": (_I, [_L, _I, _P, _P, _P, _P, _P, _F, _F, _U64, _I, _P, _P, _P]),
    "eg_bn_act_bwd": (_I, [_L, _I, _P, _P, _P, _P, _P, _P, _F, _F, _U64, _I, _I, _P, _P, _P, _P, _SZ, _P]),
    "eg_dropout_mask": (_I, [_L, _I, _F, _U64, _P, _P]),
    "eg_bn2d_fwd": (_I, [_I, _I, _L, _P, _P, _I, _P, _P, _F, _P, _P, _P, _P, _SZ, _P]),
    "eg_bn2d_bwd": (_I, [_I, _I, _L, _P, _P, _I, _P, _P, _P, _P, _F, _P, _P, _P, _P, _P, _SZ, _P]),
    "eg_col_stats": (_I, [_L, _I, _P, _P, _P, _P, _SZ, _P]),
    "eg_linear128": (_I, [_L, _P, _P, _I, _P, _P, _P, _P, _P, _P, _SZ, _P]),
    "eg_linear128_wgrad": (_I, [_L, _P, _P, _P, _P, _P, _SZ, _P]),
    "eg_clf_mid_fwd": (_I, [_L, _P, _P, _P, _P, _P, _P, _P, _SZ, _P]),
    "eg_clf_mid_bwd": (_I, [_L, _P, _P, _P, _P, _P, _P, _P, _SZ, _P]),
    "eg_clf_out_fwd": (_I, [_L, _P, _P, _P, _I, _P, _P]),
    "eg_clf_out_bwd": (_I, [_L, _P, _P, _P, _P, _I, _P, _P, _P, _P, _SZ, _P]),
    "eg_classifier_fwd": (_I, [_L, _P, C.POINTER(ClassifierParams), _P, _P, _P, _P, _P, _P, _P, _P, _SZ, _P]),
    "eg_classifier_bwd": (_I, [_L, _P, C.POINTER(ClassifierParams), _P, _P, _P, _P, _P, _P, _P, _P, _P, _P,
                               C.POINTER(ClassifierGrads), _P, _SZ, _P]),
    "eg_coord_sample_fwd": (_I, [_P, _I, _I, _I, _I, _I, _P, _P]),
    "eg_coord_sample_bwd": (_I, [_P, _P, _I, _I, _I, _I, _I, _P, _P, _P]),
    "eg_coord_update_fwd": (_I, [_P, _I, _I, _I, _I, _I, _P, C.POINTER(CoordMlpParams)] + [_P] * 10),
    "eg_coord_update_bwd": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _P, C.POINTER(CoordMlpParams)] + [_P] * 9 +
                            [_P, C.POINTER(ClassifierGrads), _P, _P]),
    "eg_mae": (_I, [_L, _P, _P, _F, _P, _P, _P]),
    "eg_linear_fwd": (_I, [_L, _I, _I, C.POINTER(View), C.POINTER(View), _P, _I, _P, C.POINTER(View), _I,
                           C.POINTER(View), _P]),
    "eg_linear_wgrad": (_I, [_L, _I, _I, C.POINTER(View), C.POINTER(View), C.POINTER(View), _P, _P, _P, _SZ, _P]),
    "eg_bce_multilevel": (_I, [_L, _P, _P, _P, _F, _F, _P, _P, _P, _SZ, _P]),
    "eg_expected_landmark_mse": (_I, [_I, _I, _I, C.POINTER(C.c_int32), _P, _P, _P, _F, _P, _P, _P, _SZ, _P]),
    "eg_node_labels": (_I, [_I, _I, _I, _I, C.POINTER(C.c_int32), _P, _P, _P, _P]),
    "eg_expected_coords": (_I, [_I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P]),
    "eg_profile_enable": (_I, [_I]),
    "eg_profile_read": (_I, [C.c_char_p, C.POINTER(C.c_double), C.POINTER(_L)]),
    "eg_profile_names": (_I, [C.c_char_p, _SZ]),
    "eg_launch_count": (_L, []),
}


class EchogladError(RuntimeError):
    pass


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: the CUDA extension is mandatory (no CPU fallback). "
            "Build it with `python -m echoglad_b200.build`.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:
            raise ImportError(f"{LIB_PATH} does not export {name}") from e
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()
WORKSPACE_BYTES = int(lib.eg_workspace_bytes())


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib.eg_last_error().decode(errors="replace")
        raise EchogladError(f"{what or 'echoglad_b200'} failed (code {rc}): {msg}")


def profile_enable(on: bool) -> None:
    check(lib.eg_profile_enable(int(on)), "eg_profile_enable")


def profile_report() -> dict:
    """{name: (total_ms, spans)} for everything recorded since profile_enable(True)."""
    buf = C.create_string_buffer(4096)
    check(lib.eg_profile_names(buf, 4096), "eg_profile_names")
    out = {}
    for name in filter(None, buf.value.decode().split(",")):
        ms, cnt = C.c_double(), _L()
        check(lib.eg_profile_read(name.encode(), C.byref(ms), C.byref(cnt)), "eg_profile_read")
        out[name] = (ms.value, cnt.value)
    return out
