"""ctypes binding of recast_rs_b200/librecast_b200.so — every symbol include/recast_b200.h declares.

The library is built in-tree by recast_rs_b200/build.py (nvcc, sm_100a).  There is no fallback: a
missing library raises, and rcb_create fails with RCB_ECUDA on a box without a GPU.
"""
from __future__ import annotations

import ctypes
import os

from .config import RcbConfig

_HERE = os.path.dirname(os.path.abspath(__file__))
# RCB_DEBUG_SO=1: the -DRCB_BOUNDS build (checked accessors in the hand-laid-out kernels; see rcb_debug_bounds_report)
LIB_PATH = os.path.join(_HERE, "librecast_b200_dbg.so" if os.environ.get("RCB_DEBUG_SO") == "1" else "librecast_b200.so")

# name -> (restype, argtypes); must list EVERY function declared in include/recast_b200.h
_vp, _i, _u32, _u64, _sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_uint32, ctypes.c_uint64, ctypes.c_size_t
_pp = ctypes.POINTER(ctypes.c_void_p)


class RcbTotals(ctypes.Structure):
    _fields_ = [(n, ctypes.c_uint64) for n in ("tiles", "triangles", "vertices", "cells", "fragments", "spans",
                                               "connections", "walkable_spans", "result_bytes")]


class RcbTileView(ctypes.Structure):
    _fields_ = [
        ("width", ctypes.c_int32), ("height", ctypes.c_int32),
        ("cell_count", _u32), ("span_count", _u32), ("conn_count", _u32), ("walkable_span_count", _u32),
        ("max_height", ctypes.c_uint16), ("max_distance", ctypes.c_uint16), ("status", _u32),
        ("hf_cell_offset", _vp), ("span_min", _vp), ("span_max", _vp), ("span_area", _vp),
        ("cell_index", _vp), ("cell_count_arr", _vp), ("areas", _vp), ("con", _vp), ("first_conn", _vp),
        ("conn_ref", _vp), ("conn_dir", _vp), ("conn_next", _vp), ("dist", _vp),
        ("reg", _vp), ("max_regions", _u32), ("_pad", _u32),
        ("conn_mask", _vp), ("conn_ref16", _vp), ("cell_count16", _vp), ("con8", _vp), ("conn_off8", _vp),
    ]


class RcbAreaOp(ctypes.Structure):
    _fields_ = [("type", ctypes.c_int32), ("area_id", _u32), ("a", ctypes.c_float * 3), ("b", ctypes.c_float * 3),
                ("first_vert", _u32), ("num_verts", _u32)]


class RcbStageTime(ctypes.Structure):
    _fields_ = [("name", ctypes.c_char * 40), ("ms", ctypes.c_float), ("calls", _u32)]


SYMBOLS = {
    "rcb_profile_enable": (_i, [_vp, _i]),
    "rcb_profile_read": (_i, [_vp, ctypes.POINTER(RcbStageTime), _i]),
    "rcb_profile_reset": (_i, [_vp]),
    "rcb_create": (_i, [_i, _pp]),
    "rcb_destroy": (None, [_vp]),
    "rcb_last_error": (ctypes.c_char_p, [_vp]),
    "rcb_set_stream": (_i, [_vp, _vp]),
    "rcb_use_own_stream": (_i, [_vp]),
    "rcb_launch_count": (_u64, [_vp]),
    "rcb_ctx_stats": (_i, [_vp, ctypes.POINTER(ctypes.c_uint64)]),
    "rcb_debug_bounds_report": (_i, [_vp, ctypes.POINTER(ctypes.c_uint64), _i]),
    "rcb_config_default": (None, [ctypes.POINTER(RcbConfig)]),
    "rcb_config_calculate_grid_size": (None, [ctypes.POINTER(RcbConfig), _vp, _vp]),
    "rcb_config_validate": (_i, [ctypes.POINTER(RcbConfig)]),
    "rcb_upload_mesh": (_i, [_vp, _vp, _sz, _vp, _sz]),
    "rcb_set_mesh_device": (_i, [_vp, _vp, _sz, _vp, _sz]),
    "rcb_upload_mesh_u16": (_i, [_vp, _vp, _sz, _vp, _sz]),
    "rcb_upload_triangle_soup": (_i, [_vp, _vp, _sz]),
    "rcb_build_tiles": (_i, [_vp, _vp, _i, _u32, _i, _pp]),
    "rcb_build_tiles_streamed": (_i, [_vp, _vp, _i, _u32, _i, _i, _pp]),
    "rcb_rasterize": (_i, [_vp, _vp, _i, _vp, ctypes.c_int32, _u32, _i, _pp]),
    "rcb_rasterize_onto": (_i, [_vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, ctypes.c_int32, _u32, _i, _pp]),
    "rcb_build_from_spans": (_i, [_vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _u32, _i, _pp]),
    "rcb_build_tiles_with_volumes": (_i, [_vp, _vp, _i, _u32, _i, _vp, _i, _vp, _sz, _pp]),
    "rcb_apply_convex_volumes": (_i, [_vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _i, _vp, _sz, _u32, _i, _pp]),
    "rcb_build_from_span_data": (_i, [_vp, _vp, _i, _vp, _vp, _i, _u32, _i, _pp]),
    "rcb_result_span_data_size": (_u64, [_vp]),
    "rcb_result_serialize_spans": (_i, [_vp, _i, _vp, _u64, _vp]),
    "rcb_apply_area_ops": (_i, [_vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _sz, _u32, _i, _pp]),
    "rcb_build_tilecache_layers": (_i, [_vp, _vp, _i, _vp, _i, ctypes.c_float, ctypes.c_float, _u32, _i, _pp]),
    "rcb_decompress_tiles": (_i, [_vp, _vp, _vp, _i, _vp, _u64, _vp]),
    "rcb_build_tilecache_compressed": (_i, [_vp, _vp, _vp, _i, _vp, _vp, _vp, _i, ctypes.c_float, ctypes.c_float, _u32, _i, _pp]),
    "rcb_result_totals": (_i, [_vp, ctypes.POINTER(RcbTotals)]),
    "rcb_result_download": (_i, [_vp]),
    "rcb_result_tile": (_i, [_vp, _i, ctypes.POINTER(RcbTileView)]),
    "rcb_result_tile_device": (_i, [_vp, _i, ctypes.POINTER(RcbTileView)]),
    "rcb_result_free": (None, [_vp]),
    "rcb_multi_create": (_i, [ctypes.POINTER(ctypes.c_int), _i, _pp]),
    "rcb_multi_destroy": (None, [_vp]),
    "rcb_multi_last_error": (ctypes.c_char_p, [_vp]),
    "rcb_multi_device_count": (_i, [_vp]),
    "rcb_multi_upload_mesh": (_i, [_vp, _vp, _sz, _vp, _sz]),
    "rcb_multi_build_tiles": (_i, [_vp, _vp, _i, _u32, _i, _i, _i, _pp]),
    "rcb_multi_result_download": (_i, [_vp]),
    "rcb_multi_result_totals": (_i, [_vp, ctypes.POINTER(RcbTotals), ctypes.POINTER(ctypes.c_int), _i]),
    "rcb_multi_result_tile": (_i, [_vp, _i, ctypes.POINTER(RcbTileView)]),
    "rcb_multi_result_tile_device": (_i, [_vp, _i, ctypes.POINTER(RcbTileView), ctypes.POINTER(ctypes.c_int)]),
    "rcb_multi_result_free": (None, [_vp]),
}

_lib = None


def lib():
    """Load the native library (built by build.py).  Raises if it is missing — no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: run `python -m recast_rs_b200.build` (or __graft_entry__.build()). "
                "recast_rs_b200 has no CPU fallback.")
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)  # AttributeError if the library does not export it
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib
