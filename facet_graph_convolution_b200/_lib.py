"""ctypes binding of libfacetconv_b200.so (the C ABI in include/facetconv_b200.h).

There is no CPU fallback: if the shared library has not been built, or no CUDA device is
present when a compute entry point is called, the call raises.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libfacetconv_b200.so")

_lib = None


class ConvShape(C.Structure):
    """Mirror of ``fgc_conv_shape``."""

    _fields_ = [(n, C.c_int32) for n in ("B", "N", "K", "Cin", "Cw", "Ca0", "Ca", "Cout", "M")]


class FacetConvError(RuntimeError):
    pass


p = C.c_void_p
i32 = C.c_int
i64 = C.c_int64
f32 = C.c_float
sz = C.c_size_t
PS = C.POINTER(ConvShape)

# name -> (restype, argtypes); every symbol include/facetconv_b200.h declares
SIGNATURES = {
    "fgc_version": (i32, []),
    "fgc_last_error": (C.c_char_p, []),
    "fgc_device_count": (i32, []),
    "fgc_launch_count": (C.c_uint64, []),
    "fgc_profile_begin": (i32, [p]),
    "fgc_profile_end": (i32, [C.c_char_p, sz]),
    "fgc_conv_fwd_workspace": (sz, [PS]),
    "fgc_conv_bwd_workspace": (sz, [PS]),
    "fgc_conv_fwd": (i32, [PS, p, p, p, p, p, p, p, p, i32, i32, f32, p, sz, p]),
    "fgc_conv_plan_bytes": (sz, [i32, i32, i32, i32]),
    "fgc_build_conv_plan": (i32, [p, i32, i32, i32, i32, p, sz, p]),
    "fgc_conv_fwd_up_supported": (i32, [PS, i32]),
    "fgc_conv_fwd_up": (i32, [PS, p, p, p, p, p, p, p, p, i32, i32, f32, i32, p, sz, p]),
    "fgc_conv_fwd_planned": (i32, [PS, p, p, p, p, p, p, p, p, p, i32, i32, f32, p, sz, p]),
    "fgc_debug_trace": (i32, [C.POINTER(i64), i32]),
    "fgc_reverse_adj_workspace": (sz, [i32, i32, i32]),
    "fgc_build_reverse_adj": (i32, [p, i32, i32, i32, p, p, C.POINTER(i64), p, sz, p]),
    "fgc_conv_bwd": (i32, [PS, p, p, p, p, p, p, p, p, p, p, p, p, p, p, p, i32, p, sz, p]),
    "fgc_build_reverse_padded": (i32, [p, p, i32, i32, i32, i32, p, p]),
    "fgc_conv_bwd_planned": (i32, [PS, p, p, p, p, p, p, p, i32, p, p, p, p, p, p, p, p, p, p, p, i32, p, sz, p, sz, p]),
    "fgc_gather_rows": (i32, [p, p, p, i32, i32, i32, i32, p]),
    "fgc_assignments": (i32, [PS, p, p, p, p, p, p, p, sz, p]),
    "fgc_pool_max": (i32, [p, p, i64, i32, i32, p]),
    "fgc_pool_max_bwd": (i32, [p, p, p, p, i64, i32, i32, p]),
    "fgc_pool_avg_ignore_zeros": (i32, [p, p, i32, i64, i32, i32, p]),
    "fgc_upsample": (i32, [p, p, i64, i32, i32, p]),
    "fgc_upsample_bwd": (i32, [p, p, i64, i32, i32, p]),
    "fgc_lrelu": (i32, [p, p, i64, f32, p]),
    "fgc_lrelu_bwd": (i32, [p, p, p, i64, f32, p]),
    "fgc_concat2": (i32, [p, p, p, i64, i32, i32, p]),
    "fgc_split2": (i32, [p, p, p, i64, i32, i32, p]),
    "fgc_gather_perm": (i32, [p, p, p, i64, i32, p]),
    "fgc_push_rows": (i32, [p, p, p, i64, i32, p]),
    "fgc_greedy_pairing": (i32, [p, p, p, i64, p, i64, p, i32, i32, p, p, p]),
    "fgc_grow_patch": (i32, [p, i64, i32, i64, i64, p, i64, p, i64, p, p, p]),
    "fgc_face_features": (i32, [p, p, i64, i64, i32, p, p, sz, p]),
    "fgc_faces_adj_workspace": (sz, [i64, i64]),
    "fgc_build_faces_adj": (i32, [p, i64, i64, i32, p, p, i32, p, sz, p]),
    "fgc_edge_maps_workspace": (sz, [i64, i64]),
    "fgc_build_edge_maps": (i32, [p, i64, i64, i32, p, p, p, p, sz, p]),
    "fgc_lin_fwd": (i32, [p, p, p, p, i64, i32, i32, i32, f32, p]),
    "fgc_lin_bwd": (i32, [p, p, p, p, p, p, i64, i32, i32, p, sz, p]),
    "fgc_lin_bwd_workspace": (sz, [i64, i32, i32]),
    "fgc_mlp_head_workspace": (sz, [i64, i32, i32, i32]),
    "fgc_mlp_head_fwd": (i32, [p, p, p, p, p, p, i64, i32, i32, i32, f32, p, sz, p]),
    "fgc_net_param_count": (i32, []),
    "fgc_net_prepared_bytes": (sz, []),
    "fgc_net_prepare": (i32, [p, i32, p, sz, p]),
    "fgc_net_fwd_workspace": (sz, [i32, i32, i32]),
    "fgc_net_fwd": (i32, [i32, i32, i32, p, p, p, p, p, i32, p, p, p, sz, p]),
    "fgc_normalize_workspace": (sz, [i64]),
    "fgc_normalize_rows": (i32, [p, p, i64, p, sz, p]),
    "fgc_normalize_rows_bwd": (i32, [p, p, p, i64, p, sz, p]),
    "fgc_normalize_rows_segmented": (i32, [p, p, i32, i64, p, p, sz, p]),
    "fgc_face_normals_loss": (i32, [p, p, p, p, i64, f32, p, sz, p]),
    "fgc_point_set_loss_workspace": (sz, [i32, i64, i64, i32, i32]),
    "fgc_point_set_loss": (i32, [p, p, i32, i64, i64, p, i32, p, i32, i32, p, p, p, sz, p]),
    "fgc_vertex_update_workspace": (sz, [i64]),
    "fgc_vertex_update_edges": (i32, [p, p, p, p, p, i64, i64, i64, i32, i32, f32, p, sz, p]),
    "fgc_vertex_update_edges_range": (i32, [p, p, p, p, p, i64, i64, i64, i32, i64, i64, f32, p]),
    "fgc_vertex_update_ms_workspace": (sz, [i64, i64]),
    "fgc_vertex_update_ms": (i32, [p, p, p, p, p, i64, i64, i32, i32, i32, i32, p, sz, p]),
    "fgc_vertex_update_ms_bwd_workspace": (sz, [i64, i64, i32, i32]),
    "fgc_vertex_update_ms_bwd": (i32, [p, p, p, p, i64, i64, i32, i32, i32, i32, p, p, p, p, p, p, p, p, sz, p]),
    "fgc_conv_fwd_host": (i32, [PS, p, p, p, p, p, p, p, p, i32, i32, f32, i32]),
    "fgc_conv_fwd_bwd_host": (i32, [PS, p, p, p, p, p, p, p, p, p, p, p, p, p, p, p, i32, i32]),
    "fgc_host_release": (None, []),
}


def lib():
    """Loads the shared library once; raises FacetConvError when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise FacetConvError(
                "facet_graph_convolution_b200: %s is missing -- build it with "
                "`python -m facet_graph_convolution_b200.build` (there is no CPU fallback)" % LIB_PATH)
        h = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(h, name)  # AttributeError => header and library out of sync
            fn.restype = res
            fn.argtypes = args
        _lib = h
    return _lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = lib().fgc_last_error().decode("utf-8", "replace")
        raise FacetConvError("%s failed (code %d): %s" % (what or "facetconv call", rc, msg))


def require_device():
    if lib().fgc_device_count() < 1:
        raise FacetConvError("facet_graph_convolution_b200: no CUDA device visible (no CPU fallback)")
