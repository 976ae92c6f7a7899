"""Host-side mirror of the reference's voxelization interface over the C ABI.

Same names, argument meaning and error behaviour as crates/recast/src/lib.rs (RecastBuilder),
heightfield.rs (Heightfield filters), compact_heightfield.rs (CompactHeightfield) and area.rs
(erode_walkable_area); every method is one call into librecast_b200.so.  Nothing here computes on
the CPU and nothing imports oracle/.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np

from . import _ffi
from .config import RcbConfig, RecastConfig, configs_to_array
from .errors import check

STAGE_RASTER, STAGE_FILTER, STAGE_COMPACT, STAGE_ERODE, STAGE_DIST = 1, 2, 4, 8, 16
FILTER_LOW_HANGING, FILTER_LEDGE, FILTER_LOW_HEIGHT = 0x100, 0x200, 0x400
STAGE_ALL_BUILD = STAGE_RASTER | STAGE_FILTER | STAGE_COMPACT | STAGE_DIST
STAGE_REGIONS = 0x40  # CompactHeightfield::build_regions (watershed.rs:364-510); implies COMPACT | DIST
STAGE_READD = 0x20  # re-insert given spans through Heightfield::add_span (checkpoint.rs:29-60)
RESULT_SLIM = 0x1000  # rcb_result_download copies only what the host cannot rebuild (see TileVoxels.expand_slim)
WIRE_VOXEL_TILE, WIRE_DYNAMIC_TILE = 0, 1
NONE = 0xFFFFFFFF


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _np_from(ptr, count, dtype):
    if not ptr or count == 0:
        return np.zeros(0, dtype=dtype)
    nbytes = count * np.dtype(dtype).itemsize
    buf = (ctypes.c_char * nbytes).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype, count=count).copy()


@dataclass
class TileVoxels:
    """One tile's output: heightfield CSR + compact heightfield (flat SoA of rcb_tile_view)."""
    width: int
    height: int
    span_count: int
    conn_count: int
    walkable_span_count: int
    max_height: int
    max_distance: int
    status: int
    hf_cell_offset: np.ndarray
    span_min: np.ndarray
    span_max: np.ndarray
    span_area: np.ndarray
    cell_index: Optional[np.ndarray] = None
    cell_count: Optional[np.ndarray] = None
    areas: Optional[np.ndarray] = None
    con: Optional[np.ndarray] = None
    first_conn: Optional[np.ndarray] = None
    conn_ref: Optional[np.ndarray] = None
    conn_dir: Optional[np.ndarray] = None
    conn_next: Optional[np.ndarray] = None
    dist: Optional[np.ndarray] = None
    reg: Optional[np.ndarray] = None  # CompactSpan.reg, when REGIONS ran
    max_regions: int = 0
    conn_mask: Optional[np.ndarray] = None   # RESULT_SLIM: u8[S], bit d = the span is connected in direction d
    conn_ref16: Optional[np.ndarray] = None  # RESULT_SLIM: u16[K] conn_ref (None when a tile holds 65536+ spans)
    cell_count16: Optional[np.ndarray] = None  # RESULT_SLIM: u16[C] cell_count
    con8: Optional[np.ndarray] = None  # RESULT_SLIM: u8[4S] con
    conn_off8: Optional[np.ndarray] = None  # RESULT_SLIM: u8[K] offset of the connected span inside its (neighbour) cell

    def expand_slim(self) -> "TileVoxels":
        """Rebuilds, on the host, the arrays a RESULT_SLIM download leaves out, from the layout rule of
        compact_heightfield.rs:375-389: connections lie in span order and, inside a span, in descending direction
        (dir 7..0 are pushed with `next` -> the one pushed before, `first_connection` = the last pushed).  Pure index
        bookkeeping of the binding layer (the Rust shim does the same loop while it fills Vec<CompactSpan>)."""
        if self.conn_mask is None:
            return self
        S = self.span_count
        if self.cell_count is None:
            self.cell_count = self.cell_count16.astype(np.uint32)
        if self.areas is None:  # nothing wrote CompactHeightfield.areas: still the copy of CompactSpan.area it starts as
            self.areas = self.span_area.copy()
        self.hf_cell_offset = np.concatenate([[0], np.cumsum(self.cell_count, dtype=np.uint64)]).astype(np.uint32)
        self.cell_index = np.where(self.cell_count > 0, self.hf_cell_offset[:-1], NONE).astype(np.uint32)
        bits = np.unpackbits(self.conn_mask.reshape(-1, 1), axis=1)  # column j = direction 7 - j: already descending
        cnt = bits.sum(axis=1, dtype=np.int64)
        ends = np.cumsum(cnt)
        self.first_conn = np.where(cnt > 0, ends - 1, NONE).astype(np.uint32)
        self.conn_dir = (7 - np.nonzero(bits)[1]).astype(np.uint8)
        K = int(ends[-1]) if S else 0
        nxt = np.arange(-1, K - 1, dtype=np.int64)
        nxt[(ends - cnt)[cnt > 0]] = NONE  # the first connection pushed of every block
        self.conn_next = nxt.astype(np.uint32)
        if self.conn_off8 is not None:
            # conn_ref = first span of the neighbour cell (by direction) + offset; con = the offsets of the cardinal directions
            span_cell = np.repeat(np.arange(self.cell_count.size, dtype=np.int64), self.cell_count)
            owner = np.repeat(np.arange(S, dtype=np.int64), cnt)  # the span every connection belongs to
            d = self.conn_dir.astype(np.int64)
            dx = np.array([-1, 0, 1, 1, 1, 0, -1, -1], np.int64)  # compact_heightfield.rs:109-111
            dz = np.array([-1, -1, -1, 0, 1, 1, 1, 0], np.int64)
            ncell = span_cell[owner] + dz[d] * self.width + dx[d]
            off = self.conn_off8.astype(np.int64)
            self.conn_ref = (self.cell_index[ncell].astype(np.int64) + off).astype(np.uint32)
            con = np.full((S, 4), 63, np.uint16)
            d4 = np.array([-1, 3, -1, 2, -1, 1, -1, 0], np.int64)[d]  # 8-direction -> con slot (:397-427), -1: diagonal
            card = d4 >= 0
            con[owner[card], d4[card]] = off[card].astype(np.uint16)
            self.con = con.reshape(-1)
        else:
            if self.conn_ref is None:
                self.conn_ref = self.conn_ref16.astype(np.uint32)
            if self.con is None:
                self.con = self.con8.astype(np.uint16)
        assert K == self.conn_count
        return self


def _tile_from_view(v, expand: bool = True) -> "TileVoxels":
    """TileVoxels (host copies) of a filled rcb_tile_view with host pointers."""
    C, S, K = v.cell_count, v.span_count, v.conn_count
    tv = TileVoxels(v.width, v.height, S, K, v.walkable_span_count, v.max_height, v.max_distance, v.status,
                    _np_from(v.hf_cell_offset, C + 1, np.uint32) if v.hf_cell_offset else None, _np_from(v.span_min, S, np.int16),
                    _np_from(v.span_max, S, np.int16), _np_from(v.span_area, S, np.uint8))
    if v.dist:
        tv.dist = _np_from(v.dist, S, np.uint16)
        if v.conn_mask and not v.first_conn:  # RESULT_SLIM host view: the binding layer rebuilds the rest (expand_slim)
            tv.hf_cell_offset = None
            tv.conn_mask = _np_from(v.conn_mask, S, np.uint8)
            tv.areas = _np_from(v.areas, S, np.uint8) if v.areas else None
            if v.cell_count16:
                tv.cell_count16 = _np_from(v.cell_count16, C, np.uint16)
            else:
                tv.cell_count = _np_from(v.cell_count_arr, C, np.uint32)
            if v.conn_off8:
                tv.conn_off8 = _np_from(v.conn_off8, K, np.uint8)
            else:  # an offset passed 255: references and con in the narrowest form that holds them
                if v.conn_ref16:
                    tv.conn_ref16 = _np_from(v.conn_ref16, K, np.uint16)
                else:
                    tv.conn_ref = _np_from(v.conn_ref, K, np.uint32)
                if v.con8:
                    tv.con8 = _np_from(v.con8, 4 * S, np.uint8)
                else:
                    tv.con = _np_from(v.con, 4 * S, np.uint16)
            if expand:
                tv.expand_slim()
        else:
            tv.cell_index = _np_from(v.cell_index, C, np.uint32)
            tv.cell_count = _np_from(v.cell_count_arr, C, np.uint32)
            tv.areas = _np_from(v.areas, S, np.uint8)
            tv.con = _np_from(v.con, 4 * S, np.uint16)
            tv.first_conn = _np_from(v.first_conn, S, np.uint32)
            tv.conn_ref = _np_from(v.conn_ref, K, np.uint32)
            tv.conn_dir = _np_from(v.conn_dir, K, np.uint8)
            tv.conn_next = _np_from(v.conn_next, K, np.uint32)
        if v.reg:
            tv.reg, tv.max_regions = _np_from(v.reg, S, np.uint16), int(v.max_regions)
    return tv

class BatchResult:
    """Owner of an rcb_result (device arena + optional pinned host copy)."""

    def __init__(self, ctx: "Context", handle):
        self._ctx, self._h = ctx, handle

    def totals(self) -> dict:
        t = _ffi.RcbTotals()
        check(_ffi.lib().rcb_result_totals(self._h, ctypes.byref(t)), self._ctx._h)
        return {n: int(getattr(t, n)) for n, _ in _ffi.RcbTotals._fields_}

    def download(self) -> "BatchResult":
        check(_ffi.lib().rcb_result_download(self._h), self._ctx._h)
        return self

    def tile(self, i: int, expand: bool = True) -> TileVoxels:
        """One tile of a downloaded result.  For a RESULT_SLIM result `expand` rebuilds the arrays the download left out."""
        v = _ffi.RcbTileView()
        check(_ffi.lib().rcb_result_tile(self._h, i, ctypes.byref(v)), self._ctx._h)
        return _tile_from_view(v, expand)

    def span_data_size(self) -> int:
        return int(_ffi.lib().rcb_result_span_data_size(self._h))

    def serialize_spans(self, fmt: int = 0, out: Optional[np.ndarray] = None):
        """Span data of every tile (VoxelTile::serialize_spans, io/voxel_tile.rs:70-118; fmt 1 = little-endian twin).
        Returns (bytes as uint8 array, tile_offsets[ntiles+1]).  `out`: caller's uint8 buffer (page-locked memory
        makes the copy run at PCIe speed)."""
        n = self.span_data_size()
        if out is None:
            out = np.empty(max(n, 1), np.uint8)
        if out.dtype != np.uint8 or out.size < n or not out.flags.c_contiguous:
            raise ValueError(f"out must be a contiguous uint8 array of at least {n} bytes")
        off = np.zeros(self.totals()["tiles"] + 1, np.uint64)
        check(_ffi.lib().rcb_result_serialize_spans(self._h, fmt, _ptr(out), n, _ptr(off)), self._ctx._h)
        return out[:n], off

    def tile_device_view(self, i: int) -> _ffi.RcbTileView:
        v = _ffi.RcbTileView()
        check(_ffi.lib().rcb_result_tile_device(self._h, i, ctypes.byref(v)), self._ctx._h)
        return v

    def free(self):
        # safe after Context.close(): rcb_destroy releases the arenas of results that are still alive and leaves the
        # handle valid for exactly this call
        if self._h:
            _ffi.lib().rcb_result_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.free()


class Context:
    """One rcb_ctx: a device, a stream, scratch pools.  One per host thread (reference types are !Send)."""

    def __init__(self, device: int = 0):
        h = ctypes.c_void_p()
        check(_ffi.lib().rcb_create(device, ctypes.byref(h)), None,
              f"rcb_create(device={device}) failed: no CUDA device (recast_rs_b200 has no CPU fallback)")
        self._h = h
        self.device = device
        self._mesh_keep = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def close(self):
        if self._h:
            _ffi.lib().rcb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream: int | None):
        """Run on the given cudaStream_t handle (0 / None = the legacy default stream)."""
        check(_ffi.lib().rcb_set_stream(self._h, ctypes.c_void_p(cuda_stream or 0)), self._h)

    def use_own_stream(self):
        check(_ffi.lib().rcb_use_own_stream(self._h), self._h)

    def profile_enable(self, on: bool = True):
        check(_ffi.lib().rcb_profile_enable(self._h, int(on)), self._h)

    def profile_reset(self):
        check(_ffi.lib().rcb_profile_reset(self._h), self._h)

    def profile_read(self) -> dict:
        """{kernel name: (summed device ms, launches)} since the last reset."""
        arr = (_ffi.RcbStageTime * 64)()
        n = _ffi.lib().rcb_profile_read(self._h, arr, 64)
        return {arr[i].name.decode(): (float(arr[i].ms), int(arr[i].calls)) for i in range(min(n, 64))}

    def launch_count(self) -> int:
        return int(_ffi.lib().rcb_launch_count(self._h))

    def stats(self) -> dict:
        """{'speculative': batches issued without a host round trip, 'rebuilt': of those, redone because a size was short,
        'cached_batches': configuration batches the context remembers}."""
        a = (ctypes.c_uint64 * 4)()
        check(_ffi.lib().rcb_ctx_stats(self._h, a), self._h)
        return {"speculative": int(a[0]), "rebuilt": int(a[1]), "cached_batches": int(a[2])}

    def bounds_report(self, reset: bool = False) -> dict:
        """The -DRCB_BOUNDS build's verdict (RCB_DEBUG_SO=1): {'instrumented', 'violations', 'first': (file id, line, index, size)}."""
        a = (ctypes.c_uint64 * 4)()
        r = _ffi.lib().rcb_debug_bounds_report(self._h, a, int(reset))
        check(min(r, 0), self._h)
        return {"instrumented": r == 1, "violations": int(a[0]), "first": (int(a[1]) >> 32, int(a[1]) & 0xFFFFFFFF, int(a[2]), int(a[3]))}

    def upload_mesh(self, vertices, indices):
        v = np.ascontiguousarray(vertices, dtype=np.float32).reshape(-1)
        t = np.ascontiguousarray(indices, dtype=np.int32).reshape(-1)
        check(_ffi.lib().rcb_upload_mesh(self._h, _ptr(v), v.size, _ptr(t), t.size), self._h)

    def upload_mesh_u16(self, vertices, indices_u16):
        """The mesh of rasterize_triangles_u16 (rasterization.rs:432): u16 indices, widened on the device."""
        v = np.ascontiguousarray(vertices, dtype=np.float32).reshape(-1)
        t = np.ascontiguousarray(indices_u16, dtype=np.uint16).reshape(-1)
        check(_ffi.lib().rcb_upload_mesh_u16(self._h, _ptr(v), v.size, _ptr(t), t.size), self._h)

    def upload_triangle_soup(self, vertices):
        """The mesh of rasterize_triangle_soup (rasterization.rs:459): 9 floats per triangle, no index list."""
        v = np.ascontiguousarray(vertices, dtype=np.float32).reshape(-1)
        if v.size % 9:
            raise ValueError("a triangle soup holds 9 floats per triangle")
        check(_ffi.lib().rcb_upload_triangle_soup(self._h, _ptr(v), v.size // 9), self._h)

    def set_mesh_device(self, d_verts_ptr: int, nfloats: int, d_tris_ptr: int, nidx: int, keep=None):
        self._mesh_keep = keep  # keeps the owning tensors alive
        check(_ffi.lib().rcb_set_mesh_device(self._h, ctypes.c_void_p(d_verts_ptr), nfloats,
                                             ctypes.c_void_p(d_tris_ptr), nidx), self._h)

    def build_tiles(self, cfgs, stage_mask=STAGE_ALL_BUILD, erode_radius=0) -> BatchResult:
        arr = cfgs if isinstance(cfgs, ctypes.Array) else configs_to_array(cfgs)
        out = ctypes.c_void_p()
        check(_ffi.lib().rcb_build_tiles(self._h, ctypes.cast(arr, ctypes.c_void_p), len(arr), stage_mask, erode_radius,
                                         ctypes.byref(out)), self._h)
        return BatchResult(self, out)

    def rasterize(self, cfgs, tri_areas, flag_merge_threshold, stage_mask=STAGE_RASTER, erode_radius=0) -> BatchResult:
        arr = cfgs if isinstance(cfgs, ctypes.Array) else configs_to_array(cfgs)
        a = np.ascontiguousarray(tri_areas, dtype=np.uint8)
        out = ctypes.c_void_p()
        check(_ffi.lib().rcb_rasterize(self._h, ctypes.cast(arr, ctypes.c_void_p), len(arr), _ptr(a), flag_merge_threshold,
                                       stage_mask, erode_radius, ctypes.byref(out)), self._h)
        return BatchResult(self, out)

    def build_tiles_streamed(self, cfgs, stage_mask: int = STAGE_ALL_BUILD, erode_radius: int = 0, groups: int = 8) -> BatchResult:
        """build_tiles + download with the device-to-host copy of finished tile groups overlapped with the
        computation of the next ones; the result is already in host memory."""
        arr = cfgs if isinstance(cfgs, ctypes.Array) else configs_to_array(cfgs)
        out = ctypes.c_void_p()
        check(_ffi.lib().rcb_build_tiles_streamed(self._h, ctypes.cast(arr, ctypes.c_void_p), len(arr), stage_mask, erode_radius, groups,
                                                  ctypes.byref(out)), self._h)
        return BatchResult(self, out)

    def rasterize_onto(self, cfgs, cell_offset, span_min, span_max, span_area, stage_mask: int = 0, tri_areas=None,
                       flag_merge_threshold: int = 1, erode_radius: int = 0) -> BatchResult:
        """Rasterize the uploaded mesh onto heightfields that already hold spans (rasterization.rs:405-428 / lib.rs:186-290 on a
        non-empty Heightfield).  tri_areas None: areas by slope, as RecastBuilder does."""
        arr = cfgs if isinstance(cfgs, ctypes.Array) else configs_to_array(cfgs)
        co = np.ascontiguousarray(cell_offset, dtype=np.uint32)
        a = np.ascontiguousarray(span_min, dtype=np.int16)
        b = np.ascontiguousarray(span_max, dtype=np.int16)
        c = np.ascontiguousarray(span_area, dtype=np.uint8)
        ta = None if tri_areas is None else np.ascontiguousarray(tri_areas, dtype=np.uint8)
        out = ctypes.c_void_p()
        check(_ffi.lib().rcb_rasterize_onto(self._h, ctypes.cast(arr, ctypes.c_void_p), len(arr), _ptr(co), _ptr(a), _ptr(b), _ptr(c), _ptr(ta),
                                            flag_merge_threshold, stage_mask, erode_radius, ctypes.byref(out)), self._h)
        return BatchResult(self, out)

    def build_from_spans(self, cfgs, cell_offset, span_min, span_max, span_area, stage_mask, erode_radius=0,
                         areas_override=None) -> BatchResult:
        arr = cfgs if isinstance(cfgs, ctypes.Array) else configs_to_array(cfgs)
        co = np.ascontiguousarray(cell_offset, dtype=np.uint32)
        a = np.ascontiguousarray(span_min, dtype=np.int16)
        b = np.ascontiguousarray(span_max, dtype=np.int16)
        c = np.ascontiguousarray(span_area, dtype=np.uint8)
        ov = None if areas_override is None else np.ascontiguousarray(areas_override, dtype=np.uint8)
        out = ctypes.c_void_p()
        check(_ffi.lib().rcb_build_from_spans(self._h, ctypes.cast(arr, ctypes.c_void_p), len(arr), _ptr(co), _ptr(a),
                                              _ptr(b), _ptr(c), _ptr(ov), stage_mask, erode_radius, ctypes.byref(out)),
              self._h)
        return BatchResult(self, out)


    def build_tiles_with_volumes(self, cfgs, volumes, stage_mask: int = STAGE_ALL_BUILD, erode_radius: int = 0) -> BatchResult:
        """RecastBuilder::build_mesh_with_volumes' voxel half (lib.rs:293-316): build_tiles with a ConvexVolumeSet applied to
        the solid heightfield after the filters (convex_volume.rs:392-430)."""
        arr = cfgs if isinstance(cfgs, ctypes.Array) else configs_to_array(cfgs)
        va, vv = volumes.pack()
        out = ctypes.c_void_p()
        check(_ffi.lib().rcb_build_tiles_with_volumes(self._h, ctypes.cast(arr, ctypes.c_void_p), len(arr), stage_mask, erode_radius,
                                                      ctypes.cast(va, ctypes.c_void_p), len(volumes.volumes), _ptr(vv) if vv.size else None,
                                                      vv.size, ctypes.byref(out)), self._h)
        return BatchResult(self, out)

    def apply_convex_volumes(self, cfgs, cell_offset, span_min, span_max, span_area, volumes, stage_mask: int = 0,
                             erode_radius: int = 0) -> BatchResult:
        """ConvexVolumeSet::apply_to_heightfield (convex_volume.rs:392-430) on existing heightfields (CSR), later stages optional."""
        arr = cfgs if isinstance(cfgs, ctypes.Array) else configs_to_array(cfgs)
        co = np.ascontiguousarray(cell_offset, dtype=np.uint32)
        a = np.ascontiguousarray(span_min, dtype=np.int16)
        b = np.ascontiguousarray(span_max, dtype=np.int16)
        c = np.ascontiguousarray(span_area, dtype=np.uint8)
        va, vv = volumes.pack()
        out = ctypes.c_void_p()
        check(_ffi.lib().rcb_apply_convex_volumes(self._h, ctypes.cast(arr, ctypes.c_void_p), len(arr), _ptr(co), _ptr(a), _ptr(b), _ptr(c),
                                                  ctypes.cast(va, ctypes.c_void_p), len(volumes.volumes), _ptr(vv) if vv.size else None, vv.size,
                                                  stage_mask, erode_radius, ctypes.byref(out)), self._h)
        return BatchResult(self, out)

    def build_from_span_data(self, cfgs, data, tile_offsets, fmt: int, stage_mask: int = 0, erode_radius: int = 0) -> BatchResult:
        """Heightfields from serialized span data: fmt WIRE_VOXEL_TILE = VoxelTile::deserialize_spans
        (io/voxel_tile.rs:120-176), WIRE_DYNAMIC_TILE = DynamicTile::reconstruct_heightfield (dynamic_tile.rs:147-257)."""
        arr = cfgs if isinstance(cfgs, ctypes.Array) else configs_to_array(cfgs)
        d = np.ascontiguousarray(np.frombuffer(data, np.uint8) if isinstance(data, (bytes, bytearray)) else data, dtype=np.uint8)
        off = np.ascontiguousarray(tile_offsets, dtype=np.uint64)
        if off.size != len(arr) + 1:
            raise ValueError("tile_offsets must hold ntiles + 1 entries")
        out = ctypes.c_void_p()
        check(_ffi.lib().rcb_build_from_span_data(self._h, ctypes.cast(arr, ctypes.c_void_p), len(arr), _ptr(d) if d.size else None,
                                                  _ptr(off), fmt, stage_mask, erode_radius, ctypes.byref(out)), self._h)
        return BatchResult(self, out)

    def apply_area_ops(self, cfgs, cell_offset, span_min, span_max, span_area, ops, areas_override=None, stage_mask=STAGE_COMPACT,
                       erode_radius=0) -> BatchResult:
        """ops: list of ("median",) | ("box", bmin, bmax, area) | ("cylinder", pos, radius, height, area) |
        ("poly", verts[n,3], min_y, max_y, area), applied in order (area.rs:238, 298, 544, 359)."""
        arr = cfgs if isinstance(cfgs, ctypes.Array) else configs_to_array(cfgs)
        co = np.ascontiguousarray(cell_offset, dtype=np.uint32)
        a = np.ascontiguousarray(span_min, dtype=np.int16)
        b = np.ascontiguousarray(span_max, dtype=np.int16)
        c = np.ascontiguousarray(span_area, dtype=np.uint8)
        ov = None if areas_override is None else np.ascontiguousarray(areas_override, dtype=np.uint8)
        oa = (_ffi.RcbAreaOp * max(len(ops), 1))()
        poly = []
        f3 = lambda v: (ctypes.c_float * 3)(*[float(x) for x in v])  # noqa: E731
        for i, op in enumerate(ops):
            kind = op[0]
            if kind == "median":
                oa[i].type = 0
            elif kind == "box":
                oa[i].type, oa[i].a, oa[i].b, oa[i].area_id = 1, f3(op[1]), f3(op[2]), int(op[3])
            elif kind == "cylinder":
                oa[i].type, oa[i].a, oa[i].b, oa[i].area_id = 2, f3(op[1]), f3((op[2], op[3], 0.0)), int(op[4])
            elif kind == "poly":
                v = np.asarray(op[1], dtype=np.float32).reshape(-1, 3)
                oa[i].type, oa[i].a, oa[i].b, oa[i].area_id = 3, f3((0.0, op[2], 0.0)), f3((0.0, op[3], 0.0)), int(op[4])
                oa[i].first_vert, oa[i].num_verts = sum(len(p) for p in poly), len(v)
                poly.append(v)
            else:
                raise ValueError(kind)
        pv = np.ascontiguousarray(np.concatenate(poly).reshape(-1) if poly else np.zeros(0, np.float32), dtype=np.float32)
        out = ctypes.c_void_p()
        check(_ffi.lib().rcb_apply_area_ops(self._h, ctypes.cast(arr, ctypes.c_void_p), len(arr), _ptr(co), _ptr(a), _ptr(b), _ptr(c),
                                            _ptr(ov), ctypes.cast(oa, ctypes.c_void_p), len(ops), _ptr(pv), pv.size, stage_mask,
                                            erode_radius, ctypes.byref(out)), self._h)
        return BatchResult(self, out)

    def build_tilecache_layers(self, layers, obstacles, cs: float, ch: float, stage_mask=STAGE_COMPACT, erode_radius=0
                               ) -> BatchResult:
        """TileCacheBuilder::layer_to_compact_heightfield + rasterize_obstacles for a batch of layers
        (detour-tilecache/src/tile_cache_builder.rs:100-126, 175-251)."""
        from . import tilecache as tcm

        # `layers` may be a tilecache.Packed (C structs built once, e.g. outside a timed loop)
        pk = layers if isinstance(layers, tcm.Packed) else tcm.Packed(layers, obstacles)
        out = ctypes.c_void_p()
        check(_ffi.lib().rcb_build_tilecache_layers(self._h, ctypes.cast(pk.layers, ctypes.c_void_p), pk.nlayers,
                                                    ctypes.cast(pk.obstacles, ctypes.c_void_p), pk.nobstacles, cs, ch, stage_mask,
                                                    erode_radius, ctypes.byref(out)), self._h)
        return BatchResult(self, out)


    def decompress_tiles(self, blobs: Sequence[bytes]):
        """TileCache::decompress_tile (detour-tilecache/src/tile_cache.rs:854-872) for a batch of compressed tiles;
        returns the decoded tiles as a list of bytes.  Raises errors.DetourFailure for a blob lz4_flex rejects."""
        buf, off = _join_blobs(blobs)
        n = len(blobs)
        ooff = np.zeros(n + 1, np.uint64)
        check(_ffi.lib().rcb_decompress_tiles(self._h, _ptr(buf) if buf.size else None, _ptr(off), n, None, 0, _ptr(ooff)), self._h)
        out = np.empty(max(int(ooff[-1]), 1), np.uint8)
        check(_ffi.lib().rcb_decompress_tiles(self._h, _ptr(buf) if buf.size else None, _ptr(off), n, _ptr(out), int(ooff[-1]), _ptr(ooff)),
              self._h)
        raw = out.tobytes()
        return [raw[int(ooff[i]):int(ooff[i + 1])] for i in range(n)]

    def build_tilecache_compressed(self, blobs, obstacles, obstacle_ranges, cs: float, ch: float, stage_mask=STAGE_COMPACT,
                                   erode_radius=0) -> BatchResult:
        """decompress_tile -> TileCacheLayer::from_bytes -> layer_to_compact_heightfield + rasterize_obstacles, all on the
        device (tile_cache.rs:854-872, tile_cache_data.rs:281-320, tile_cache_builder.rs:100-251).
        blobs: list of compressed tiles, or a tilecache.CompressedTiles built once; obstacle_ranges: per layer
        (first, count) into `obstacles` (None: no obstacles)."""
        from . import tilecache as tcm

        ct = blobs if isinstance(blobs, tcm.CompressedTiles) else tcm.CompressedTiles(blobs, obstacles, obstacle_ranges)
        out = ctypes.c_void_p()
        check(_ffi.lib().rcb_build_tilecache_compressed(self._h, _ptr(ct.buf) if ct.buf.size else None, _ptr(ct.off), ct.n,
                                                        _ptr(ct.first), _ptr(ct.count), ctypes.cast(ct.obstacles, ctypes.c_void_p)
                                                        if ct.nobstacles else None, ct.nobstacles, cs, ch, stage_mask, erode_radius,
                                                        ctypes.byref(out)), self._h)
        return BatchResult(self, out)


class MultiResult:
    """Owner of an rcb_multi_result: the groups of a multi-device build, addressed by the caller's tile index."""

    def __init__(self, mctx: "MultiContext", handle):
        self._m, self._h = mctx, handle

    def _check(self, status):
        if status != 0:
            raw = _ffi.lib().rcb_multi_last_error(self._m._h) if self._m._h else b""
            check(status, None, raw.decode("utf-8", "replace") if raw else "")

    def totals(self) -> dict:
        t = _ffi.RcbTotals()
        per = (ctypes.c_int * self._m.ndevices)()
        self._check(_ffi.lib().rcb_multi_result_totals(self._h, ctypes.byref(t), per, self._m.ndevices))
        d = {n: int(getattr(t, n)) for n, _ in _ffi.RcbTotals._fields_}
        d["tiles_per_device"] = [int(x) for x in per]
        return d

    def download(self) -> "MultiResult":
        self._check(_ffi.lib().rcb_multi_result_download(self._h))
        return self

    def tile(self, i: int, expand: bool = True) -> TileVoxels:
        v = _ffi.RcbTileView()
        self._check(_ffi.lib().rcb_multi_result_tile(self._h, i, ctypes.byref(v)))
        return _tile_from_view(v, expand)

    def tile_device(self, i: int):
        """(device ordinal, rcb_tile_view with device pointers) of tile i."""
        v, dev = _ffi.RcbTileView(), ctypes.c_int(-1)
        self._check(_ffi.lib().rcb_multi_result_tile_device(self._h, i, ctypes.byref(v), ctypes.byref(dev)))
        return int(dev.value), v

    def free(self):
        if self._h and self._m._h:
            _ffi.lib().rcb_multi_result_free(self._h)
        self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.free()


class MultiContext:
    """rcb_multi: the devices of one box behind one handle.  The multi-tile loop of detour-dynamic
    (DynamicNavMesh::build, dynamic_navmesh.rs:558-578) with the tiles dealt to the GPUs group by group from a shared
    queue; results stay on the device that built them (or are downloaded by that device's thread)."""

    def __init__(self, devices: Sequence[int]):
        arr = (ctypes.c_int * len(devices))(*[int(d) for d in devices])
        h = ctypes.c_void_p()
        check(_ffi.lib().rcb_multi_create(arr, len(devices), ctypes.byref(h)), None,
              f"rcb_multi_create({list(devices)}) failed: no such CUDA device (recast_rs_b200 has no CPU fallback)")
        self._h, self.ndevices, self._live = h, len(devices), []

    def _check(self, status):
        if status != 0:
            raw = _ffi.lib().rcb_multi_last_error(self._h)
            check(status, None, raw.decode("utf-8", "replace") if raw else "")

    def bounds_report(self, reset: bool = False) -> dict:
        """The -DRCB_BOUNDS build's verdict (RCB_DEBUG_SO=1): {'instrumented', 'violations', 'first': (file id, line, index, size)}."""
        a = (ctypes.c_uint64 * 4)()
        r = _ffi.lib().rcb_debug_bounds_report(self._h, a, int(reset))
        check(min(r, 0), self._h)
        return {"instrumented": r == 1, "violations": int(a[0]), "first": (int(a[1]) >> 32, int(a[1]) & 0xFFFFFFFF, int(a[2]), int(a[3]))}

    def upload_mesh(self, vertices, indices):
        v = np.ascontiguousarray(vertices, dtype=np.float32).reshape(-1)
        t = np.ascontiguousarray(indices, dtype=np.int32).reshape(-1)
        self._check(_ffi.lib().rcb_multi_upload_mesh(self._h, _ptr(v), v.size, _ptr(t), t.size))

    def build_tiles(self, cfgs, stage_mask=STAGE_ALL_BUILD, erode_radius=0, tiles_per_group: int = 0, download: bool = False) -> MultiResult:
        arr = cfgs if isinstance(cfgs, ctypes.Array) else configs_to_array(cfgs)
        out = ctypes.c_void_p()
        self._check(_ffi.lib().rcb_multi_build_tiles(self._h, ctypes.cast(arr, ctypes.c_void_p), len(arr), stage_mask, erode_radius,
                                                     tiles_per_group, int(download), ctypes.byref(out)))
        r = MultiResult(self, out)
        self._live.append(r)
        return r

    def close(self):
        if self._h:
            for r in self._live:  # results hold arenas of the per-device contexts: release them first
                r.free()
            self._live = []
            _ffi.lib().rcb_multi_destroy(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _join_blobs(blobs):
    off = np.zeros(len(blobs) + 1, np.uint64)
    off[1:] = np.cumsum([len(b) for b in blobs], dtype=np.uint64)
    return np.frombuffer(b"".join(bytes(b) for b in blobs), np.uint8), off


_default_ctx: Optional[Context] = None


def default_context() -> Context:
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context(0)
    return _default_ctx


# ---- reference-shaped objects ------------------------------------------------------------------------
class Heightfield:
    """`Heightfield` (heightfield.rs:58-76) as CSR arrays; the filters run on the GPU."""

    def __init__(self, width, height, bmin, bmax, cs, ch, cell_offset=None, span_min=None, span_max=None,
                 span_area=None, ctx: Optional[Context] = None):
        self.width, self.height = int(width), int(height)
        self.bmin, self.bmax = tuple(float(x) for x in bmin), tuple(float(x) for x in bmax)
        self.cs, self.ch = float(cs), float(ch)
        n = self.width * self.height
        self.cell_offset = np.zeros(n + 1, np.uint32) if cell_offset is None else np.asarray(cell_offset, np.uint32)
        self.span_min = np.zeros(0, np.int16) if span_min is None else np.asarray(span_min, np.int16)
        self.span_max = np.zeros(0, np.int16) if span_max is None else np.asarray(span_max, np.int16)
        self.span_area = np.zeros(0, np.uint8) if span_area is None else np.asarray(span_area, np.uint8).copy()
        self._ctx = ctx

    @classmethod
    def new(cls, width, height, bmin, bmax, cs, ch, ctx=None):  # heightfield.rs:80
        return cls(width, height, bmin, bmax, cs, ch, ctx=ctx)

    def _cfg(self, **kw) -> RecastConfig:
        return RecastConfig(width=self.width, height=self.height, cs=self.cs, ch=self.ch, bmin=self.bmin, bmax=self.bmax,
                            **kw)

    def column(self, x, z):
        c = z * self.width + x
        return [(int(self.span_min[i]), int(self.span_max[i]), int(self.span_area[i]))
                for i in range(int(self.cell_offset[c]), int(self.cell_offset[c + 1]))]

    def get_span_count(self) -> int:  # heightfield.rs:1178
        return int(self.span_min.size)

    def _run_filter(self, bits, height, climb):
        ctx = self._ctx or default_context()
        cfg = self._cfg(walkable_height=int(height), walkable_climb=int(climb))
        with ctx.build_from_spans([cfg], self.cell_offset, self.span_min, self.span_max, self.span_area, bits) as r:
            self.span_area = r.download().tile(0).span_area

    def filter_low_hanging_walkable_obstacles(self, walkable_climb: int):  # heightfield.rs:286
        self._run_filter(FILTER_LOW_HANGING, 0, walkable_climb)

    def filter_ledge_spans(self, walkable_height: int, walkable_climb: int):  # heightfield.rs:332
        self._run_filter(FILTER_LEDGE, walkable_height, walkable_climb)

    def filter_walkable_low_height_spans(self, walkable_height: int):  # heightfield.rs:492
        self._run_filter(FILTER_LOW_HEIGHT, walkable_height, 0)


class CompactHeightfield:
    """`CompactHeightfield` (compact_heightfield.rs:129-170)."""

    def __init__(self, hf: Heightfield, tv: TileVoxels, ctx: Context):
        self._hf, self._ctx = hf, ctx
        self.width, self.height = hf.width, hf.height
        self.bmin, self.bmax, self.cs, self.ch = hf.bmin, hf.bmax, hf.cs, hf.ch
        self._take(tv)
        self.max_distance = 0
        self.dist = np.zeros(self.span_min.size, np.uint16)

    def _take(self, tv: TileVoxels):
        self.cell_index, self.cell_count = tv.cell_index, tv.cell_count
        self.span_min, self.span_max, self.span_area = tv.span_min, tv.span_max, tv.span_area
        self.areas, self.con, self.first_conn = tv.areas, tv.con, tv.first_conn
        self.conn_ref, self.conn_dir, self.conn_next = tv.conn_ref, tv.conn_dir, tv.conn_next
        self.walkable_span_count, self.max_height = tv.walkable_span_count, tv.max_height
        self.span_count, self.cell_total = tv.span_count, self.width * self.height

    @classmethod
    def build_from_heightfield(cls, hf: Heightfield, ctx: Optional[Context] = None):  # compact_heightfield.rs:175
        ctx = ctx or hf._ctx or default_context()
        with ctx.build_from_spans([hf._cfg()], hf.cell_offset, hf.span_min, hf.span_max, hf.span_area, STAGE_COMPACT) as r:
            return cls(hf, r.download().tile(0), ctx)

    def get_neighbor(self, span_idx: int, dir8: int) -> Optional[int]:  # compact_heightfield.rs:495
        c = int(self.first_conn[span_idx])
        while c != NONE:
            if int(self.conn_dir[c]) == dir8:
                return int(self.conn_ref[c])
            c = int(self.conn_next[c])
        return None

    def _rerun(self, stage, radius=0) -> TileVoxels:
        hf = self._hf
        with self._ctx.build_from_spans([hf._cfg()], hf.cell_offset, hf.span_min, hf.span_max, hf.span_area,
                                        STAGE_COMPACT | stage, radius, areas_override=self.areas) as r:
            return r.download().tile(0)

    def build_distance_field(self):  # compact_heightfield.rs:514
        tv = self._rerun(STAGE_DIST)
        self.dist, self.max_distance = tv.dist, tv.max_distance

    def build_regions(self, border_size: int, min_region_area: int, merge_region_area: int):
        """CompactHeightfield::build_regions (compact_heightfield.rs:770-778 -> watershed.rs:364-510): rebuilds the distance
        field, then fills `reg` (CompactSpan.reg) and `max_regions`."""
        hf = self._hf
        cfg = hf._cfg(border_size=int(border_size), min_region_area=int(min_region_area), merge_region_area=int(merge_region_area))
        with self._ctx.build_from_spans([cfg], hf.cell_offset, hf.span_min, hf.span_max, hf.span_area, STAGE_COMPACT | STAGE_REGIONS,
                                        areas_override=self.areas) as r:
            tv = r.download().tile(0)
        self.dist, self.max_distance, self.reg, self.max_regions = tv.dist, tv.max_distance, tv.reg, tv.max_regions

    def erode_walkable_area(self, radius: int):  # area.rs:73
        self.areas = self._rerun(STAGE_ERODE, radius).areas

    def _area_ops(self, ops):
        hf = self._hf
        with self._ctx.apply_area_ops([hf._cfg()], hf.cell_offset, hf.span_min, hf.span_max, hf.span_area, ops,
                                      areas_override=self.areas) as r:
            self.areas = r.download().tile(0).areas

    def median_filter_walkable_area(self):  # area.rs:238
        self._area_ops([("median",)])

    def mark_box_area(self, bmin, bmax, area_id: int):  # area.rs:298
        self._area_ops([("box", bmin, bmax, area_id)])

    def mark_cylinder_area(self, position, radius: float, height: float, area_id: int):  # area.rs:544
        self._area_ops([("cylinder", position, radius, height, area_id)])

    def mark_convex_poly_area(self, verts, min_y: float, max_y: float, area_id: int):  # area.rs:359
        self._area_ops([("poly", verts, min_y, max_y, area_id)])


def erode_walkable_area(chf: CompactHeightfield, erosion_radius: int):  # area.rs:73 (free function)
    chf.erode_walkable_area(erosion_radius)


def median_filter_walkable_area(chf: CompactHeightfield):  # area.rs:238
    chf.median_filter_walkable_area()


def mark_box_area(chf: CompactHeightfield, bmin, bmax, area_id: int):  # area.rs:298
    chf.mark_box_area(bmin, bmax, area_id)


def mark_cylinder_area(chf: CompactHeightfield, position, radius: float, height: float, area_id: int):  # area.rs:544
    chf.mark_cylinder_area(position, radius, height, area_id)


def mark_convex_poly_area(chf: CompactHeightfield, verts, num_verts: int, min_y: float, max_y: float, area_id: int):  # area.rs:359
    chf.mark_convex_poly_area(np.asarray(verts, np.float32).reshape(-1, 3)[:num_verts], min_y, max_y, area_id)


class RecastBuilder:
    """`RecastBuilder` (lib.rs:51-359), voxelization half."""

    def __init__(self, config: RecastConfig, ctx: Optional[Context] = None):
        self._config = config
        self._ctx = ctx

    @classmethod
    def new(cls, config: RecastConfig) -> "RecastBuilder":
        return cls(config)

    def config(self) -> RecastConfig:
        return self._config

    def _context(self) -> Context:
        return self._ctx or default_context()

    def build_heightfield(self, vertices, indices) -> Heightfield:  # lib.rs:116-135
        c = self._config
        hf = Heightfield.new(c.width, c.height, c.bmin, c.bmax, c.cs, c.ch, ctx=self._context())
        self.rasterize_triangles(hf, vertices, indices)
        return hf

    def rasterize_triangles(self, heightfield: Heightfield, vertices, indices) -> None:  # lib.rs:186-290
        ctx = self._context()
        ctx.upload_mesh(vertices, indices)
        c = self._config
        cfg = RecastConfig(**{**vars(c), "width": heightfield.width, "height": heightfield.height, "bmin": heightfield.bmin,
                              "bmax": heightfield.bmax, "cs": heightfield.cs, "ch": heightfield.ch})
        if heightfield.get_span_count() != 0:  # e.g. a restored checkpoint that gets its colliders back
            hf = heightfield
            with ctx.rasterize_onto([cfg], hf.cell_offset, hf.span_min, hf.span_max, hf.span_area, STAGE_FILTER) as r:
                tv = r.download().tile(0)
        else:
            with ctx.build_tiles([cfg], STAGE_RASTER | STAGE_FILTER) as r:
                tv = r.download().tile(0)
        heightfield.cell_offset, heightfield.span_min = tv.hf_cell_offset, tv.span_min
        heightfield.span_max, heightfield.span_area = tv.span_max, tv.span_area

    def build_heightfield_with_volumes(self, vertices, indices, convex_volumes) -> Heightfield:  # lib.rs:346-361
        c = self._config
        ctx = self._context()
        ctx.upload_mesh(vertices, indices)
        with ctx.build_tiles_with_volumes([c], convex_volumes, STAGE_RASTER | STAGE_FILTER) as r:
            tv = r.download().tile(0)
        return Heightfield(c.width, c.height, c.bmin, c.bmax, c.cs, c.ch, tv.hf_cell_offset, tv.span_min, tv.span_max, tv.span_area, ctx=ctx)

    def build_compact_heightfield(self, heightfield: Heightfield) -> CompactHeightfield:
        return CompactHeightfield.build_from_heightfield(heightfield, self._context())

    # The batch entry the reference lacks: one launch sequence for many tiles.
    def build_tiles(self, vertices, indices, tiles: Sequence[RecastConfig], stage_mask=STAGE_ALL_BUILD,
                    erode_radius=0) -> List[TileVoxels]:
        ctx = self._context()
        ctx.upload_mesh(vertices, indices)
        with ctx.build_tiles(tiles, stage_mask, erode_radius) as r:
            r.download()
            return [r.tile(i) for i in range(len(tiles))]
