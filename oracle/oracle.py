"""ctypes wrapper of oracle/liboracle.so — the CPU restatement of the reference hot path.

TEST INFRASTRUCTURE ONLY (see oracle.cpp header).  Imported by tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs — never by recast_rs_b200.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
import sys
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(_HERE))
from recast_rs_b200.config import RcbConfig, RecastConfig, configs_to_array  # noqa: E402  (plain data only)

NONE = 0xFFFFFFFF
_lib = None


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "oracle.cpp")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return so


def lib():
    global _lib
    if _lib is None:
        L = ctypes.CDLL(build())
        vp, i32, i16, u8, u32, f32 = (ctypes.c_void_p, ctypes.c_int32, ctypes.c_int16, ctypes.c_uint8,
                                      ctypes.c_uint32, ctypes.c_float)
        L.orc_hf_new.restype = vp
        L.orc_hf_new.argtypes = [i32, i32, vp, vp, f32, f32]
        L.orc_hf_free.argtypes = [vp]
        L.orc_add_span.argtypes = [vp, i32, i32, i16, i16, u8, i32]
        L.orc_hf_add_span.argtypes = [vp, i32, i32, i16, i16, u8]
        L.orc_rasterize_triangle.argtypes = [vp, vp, vp, vp, u8, i32]
        L.orc_rasterize_triangles.argtypes = [vp, vp, ctypes.c_size_t, vp, vp, ctypes.c_size_t, i32]
        L.orc_builder_rasterize.argtypes = [vp, vp, vp, ctypes.c_size_t, vp, ctypes.c_size_t, ctypes.c_int]
        L.orc_filter_low_hanging_walkable_obstacles.argtypes = [vp, i16]
        L.orc_filter_ledge_spans.argtypes = [vp, i16, i16]
        L.orc_filter_walkable_low_height_spans.argtypes = [vp, i16]
        L.orc_hf_poly_overflow.argtypes = [vp]
        L.orc_hf_span_count.restype = ctypes.c_uint64
        L.orc_hf_span_count.argtypes = [vp]
        L.orc_hf_export.argtypes = [vp, vp, vp, vp, vp]
        L.orc_hf_from_csr.restype = vp
        L.orc_hf_from_csr.argtypes = [i32, i32, vp, vp, f32, f32, vp, vp, vp, vp]
        L.orc_chf_build.restype = vp
        L.orc_chf_build.argtypes = [vp]
        L.orc_chf_free.argtypes = [vp]
        L.orc_chf_build_distance_field.argtypes = [vp]
        L.orc_chf_erode_walkable_area.argtypes = [vp, i32]
        L.orc_chf_get_neighbor.restype = u32
        L.orc_chf_get_neighbor.argtypes = [vp, u32, u8]
        L.orc_chf_sizes.argtypes = [vp, vp]
        L.orc_chf_export.argtypes = [vp] * 13
        L.orc_config_default.argtypes = [vp]
        L.orc_config_calculate_grid_size.argtypes = [vp, vp, vp]
        L.orc_config_validate.argtypes = [vp]
        L.orc_build_tiles.restype = ctypes.c_double
        L.orc_build_tiles.argtypes = [vp, ctypes.c_size_t, vp, ctypes.c_size_t, vp, ctypes.c_int, u32,
                                      ctypes.c_int, ctypes.c_int, ctypes.c_int, vp]
        L.orc_tilecache_build.restype = vp
        L.orc_tilecache_build.argtypes = [vp, vp, f32, f32, vp]
        L.orc_tilecache_build_batch.restype = ctypes.c_double
        L.orc_tilecache_build_batch.argtypes = [vp, ctypes.c_int, vp, f32, f32, ctypes.c_int, vp]
        L.orc_chf_median_filter_walkable_area.argtypes = [vp]
        L.orc_chf_mark_box_area.argtypes = [vp, vp, vp, u8]
        L.orc_chf_mark_cylinder_area.argtypes = [vp, vp, f32, f32, u8]
        L.orc_chf_mark_convex_poly_area.argtypes = [vp, vp, u32, f32, f32, u8]
        L.orc_chf_set_areas.argtypes = [vp, vp]
        L.orc_hardware_concurrency.restype = ctypes.c_int
        L.orc_lz4_prepended_size.restype = ctypes.c_int64
        L.orc_lz4_prepended_size.argtypes = [vp, ctypes.c_uint64]
        L.orc_lz4_decompress_size_prepended.argtypes = [vp, ctypes.c_uint64, vp, ctypes.c_uint64, vp]
        L.orc_tilecache_layer_from_bytes.argtypes = [vp, ctypes.c_uint64, vp, vp]
        L.orc_tilecache_build_compressed_batch.restype = ctypes.c_double
        L.orc_tilecache_build_compressed_batch.argtypes = [vp, vp, ctypes.c_int, vp, vp, vp, f32, f32, ctypes.c_int, vp]
        L.orc_chf_build_regions.argtypes = [vp, i32, i32, i32, vp, vp]
        L.orc_convex_is_convex.argtypes = [vp, u32]
        L.orc_convex_volume_validate.argtypes = [vp, u32, f32, f32]
        L.orc_convex_contains_point.argtypes = [vp, u32, f32, f32, vp]
        L.orc_hf_apply_convex_volumes.argtypes = [vp, vp, ctypes.c_int, vp]
        L.orc_voxel_serialize_spans.argtypes = [vp, ctypes.c_int, vp]
        L.orc_voxel_deserialize_spans.argtypes = [vp, vp, ctypes.c_uint64]
        L.orc_dynamic_reconstruct_heightfield.argtypes = [vp, vp, ctypes.c_uint64]
        L.orc_checkpoint_clone_heightfield.restype = vp
        L.orc_checkpoint_clone_heightfield.argtypes = [vp]
        _lib = L
    return _lib


def _p(a: np.ndarray):
    return a.ctypes.data_as(ctypes.c_void_p)


def _f3(v):
    return np.ascontiguousarray(np.asarray(v, dtype=np.float32).reshape(3))


class OracleError(Exception):
    def __init__(self, status):
        super().__init__(f"oracle status {status}")
        self.status = status


@dataclass
class HfArrays:
    """Heightfield as CSR (same arrays as rcb_tile_view)."""
    cell_offset: np.ndarray
    span_min: np.ndarray
    span_max: np.ndarray
    span_area: np.ndarray

    def column(self, x, z, width):
        c = z * width + x
        a, b = int(self.cell_offset[c]), int(self.cell_offset[c + 1])
        return [(int(self.span_min[i]), int(self.span_max[i]), int(self.span_area[i])) for i in range(a, b)]


class Heightfield:
    """Restated `Heightfield` (heightfield.rs:58-99) + the free functions of rasterization.rs."""

    def __init__(self, width, height, bmin, bmax, cs, ch, _handle=None):
        self.width, self.height = int(width), int(height)
        self.bmin, self.bmax = _f3(bmin), _f3(bmax)
        self.cs, self.ch = float(cs), float(ch)
        self._h = _handle or lib().orc_hf_new(self.width, self.height, _p(self.bmin), _p(self.bmax), self.cs, self.ch)

    @classmethod
    def from_config(cls, cfg: RecastConfig):
        return cls(cfg.width, cfg.height, cfg.bmin, cfg.bmax, cfg.cs, cfg.ch)

    @classmethod
    def from_csr(cls, width, height, bmin, bmax, cs, ch, cell_offset, smin, smax, area):
        bmin, bmax = _f3(bmin), _f3(bmax)
        co = np.ascontiguousarray(cell_offset, dtype=np.uint32)
        a = np.ascontiguousarray(smin, dtype=np.int16)
        b = np.ascontiguousarray(smax, dtype=np.int16)
        c = np.ascontiguousarray(area, dtype=np.uint8)
        h = lib().orc_hf_from_csr(width, height, _p(bmin), _p(bmax), cs, ch, _p(co), _p(a), _p(b), _p(c))
        return cls(width, height, bmin, bmax, cs, ch, _handle=h)

    def __del__(self):
        if getattr(self, "_h", None) and lib is not None:  # module globals may be gone at interpreter exit
            lib().orc_hf_free(self._h)
            self._h = None

    # rasterization.rs:23 add_span
    def add_span_raster(self, x, z, smin, smax, area, flag_merge_threshold):
        lib().orc_add_span(self._h, x, z, smin, smax, area, flag_merge_threshold)

    # heightfield.rs:102 Heightfield::add_span
    def add_span(self, x, z, smin, smax, area):
        st = lib().orc_hf_add_span(self._h, x, z, smin, smax, area)
        if st:
            raise OracleError(st)

    def rasterize_triangle(self, v0, v1, v2, area, flag_merge_threshold):
        a, b, c = _f3(v0), _f3(v1), _f3(v2)
        lib().orc_rasterize_triangle(self._h, _p(a), _p(b), _p(c), area, flag_merge_threshold)

    def rasterize_triangles(self, verts, tris, areas, flag_merge_threshold):
        v = np.ascontiguousarray(verts, dtype=np.float32).reshape(-1)
        t = np.ascontiguousarray(tris, dtype=np.int32).reshape(-1)
        a = np.ascontiguousarray(areas, dtype=np.uint8)
        st = lib().orc_rasterize_triangles(self._h, _p(v), v.size, _p(t), _p(a), t.size // 3, flag_merge_threshold)
        if st:
            raise OracleError(st)

    # lib.rs:186 RecastBuilder::rasterize_triangles
    def builder_rasterize(self, cfg: RecastConfig, verts, indices, run_filters=True):
        v = np.ascontiguousarray(verts, dtype=np.float32).reshape(-1)
        t = np.ascontiguousarray(indices, dtype=np.int32).reshape(-1)
        c = cfg.to_c()
        st = lib().orc_builder_rasterize(self._h, ctypes.byref(c), _p(v), v.size, _p(t), t.size, int(run_filters))
        if st:
            raise OracleError(st)

    def filter_low_hanging_walkable_obstacles(self, climb):
        lib().orc_filter_low_hanging_walkable_obstacles(self._h, climb)

    def filter_ledge_spans(self, height, climb):
        lib().orc_filter_ledge_spans(self._h, height, climb)

    def filter_walkable_low_height_spans(self, height):
        lib().orc_filter_walkable_low_height_spans(self._h, height)

    def apply_convex_volumes(self, volume_set):
        """ConvexVolumeSet::apply_to_heightfield (convex_volume.rs:392-430); volume_set: recast_rs_b200.convex_volume.ConvexVolumeSet."""
        va, vv = volume_set.pack()
        lib().orc_hf_apply_convex_volumes(self._h, ctypes.cast(va, ctypes.c_void_p), len(volume_set.volumes), _p(vv) if vv.size else None)

    # ---- wire formats (detour-dynamic) ----
    def serialize_spans(self, big_endian=True) -> np.ndarray:
        """VoxelTile::serialize_spans (io/voxel_tile.rs:70-118); big_endian=False is the byte-swapped twin."""
        out = np.zeros(2 * self.width * self.height + 12 * self.span_count(), dtype=np.uint8)
        lib().orc_voxel_serialize_spans(self._h, 1 if big_endian else 0, _p(out))
        return out

    def deserialize_spans(self, data):
        """VoxelTile::deserialize_spans (io/voxel_tile.rs:120-176) into this (empty) heightfield."""
        d = np.ascontiguousarray(data, dtype=np.uint8)
        lib().orc_voxel_deserialize_spans(self._h, _p(d), d.size)

    def reconstruct_heightfield(self, data):
        """DynamicTile::reconstruct_heightfield (dynamic_tile.rs:147-257); raises OracleError on Err."""
        d = np.ascontiguousarray(data, dtype=np.uint8)
        st = lib().orc_dynamic_reconstruct_heightfield(self._h, _p(d), d.size)
        if st:
            raise OracleError(st)

    def checkpoint_clone(self):
        """DynamicTileCheckpoint::clone_heightfield (checkpoint.rs:29-60)."""
        h = lib().orc_checkpoint_clone_heightfield(self._h)
        return Heightfield(self.width, self.height, self.bmin, self.bmax, self.cs, self.ch, _handle=h)

    @property
    def poly_overflow(self):
        return bool(lib().orc_hf_poly_overflow(self._h))

    def span_count(self):
        return int(lib().orc_hf_span_count(self._h))

    def export(self) -> HfArrays:
        n = self.span_count()
        co = np.zeros(self.width * self.height + 1, dtype=np.uint32)
        a = np.zeros(n, dtype=np.int16)
        b = np.zeros(n, dtype=np.int16)
        c = np.zeros(n, dtype=np.uint8)
        lib().orc_hf_export(self._h, _p(co), _p(a), _p(b), _p(c))
        return HfArrays(co, a, b, c)


@dataclass
class ChfArrays:
    cell_index: np.ndarray
    cell_count: np.ndarray
    span_min: np.ndarray
    span_max: np.ndarray
    span_area: np.ndarray
    areas: np.ndarray
    con: np.ndarray
    first_conn: np.ndarray
    conn_ref: np.ndarray
    conn_dir: np.ndarray
    conn_next: np.ndarray
    dist: np.ndarray
    walkable_span_count: int
    max_height: int
    max_distance: int


class CompactHeightfield:
    """Restated `CompactHeightfield` (compact_heightfield.rs:129-170)."""

    def __init__(self, handle, width, height):
        self._h, self.width, self.height = handle, width, height

    @classmethod
    def build_from_heightfield(cls, hf: Heightfield):
        return cls(lib().orc_chf_build(hf._h), hf.width, hf.height)

    def __del__(self):
        if getattr(self, "_h", None) and lib is not None:  # module globals may be gone at interpreter exit
            lib().orc_chf_free(self._h)
            self._h = None

    def build_distance_field(self):
        lib().orc_chf_build_distance_field(self._h)

    def erode_walkable_area(self, radius):
        lib().orc_chf_erode_walkable_area(self._h, radius)

    # area.rs:238 / :298 / :544 / :359
    def median_filter_walkable_area(self):
        lib().orc_chf_median_filter_walkable_area(self._h)

    def mark_box_area(self, bmin, bmax, area_id):
        a, b = _f3(bmin), _f3(bmax)
        lib().orc_chf_mark_box_area(self._h, _p(a), _p(b), area_id)

    def mark_cylinder_area(self, position, radius, height, area_id):
        a = _f3(position)
        lib().orc_chf_mark_cylinder_area(self._h, _p(a), radius, height, area_id)

    def mark_convex_poly_area(self, verts, min_y, max_y, area_id):
        v = np.ascontiguousarray(verts, dtype=np.float32).reshape(-1)
        lib().orc_chf_mark_convex_poly_area(self._h, _p(v), v.size // 3, min_y, max_y, area_id)

    def set_areas(self, areas):
        a = np.ascontiguousarray(areas, dtype=np.uint8)
        lib().orc_chf_set_areas(self._h, _p(a))

    def build_regions(self, border_size=0, min_region_area=8, merge_region_area=20):
        """CompactHeightfield::build_regions -> watershed::build_regions_watershed (watershed.rs:364-510).
        Returns (reg u16[S], max_regions); raises OracleError(-6) on region id overflow."""
        sizes = np.zeros(8, dtype=np.uint64)
        lib().orc_chf_sizes(self._h, _p(sizes))
        reg = np.zeros(max(int(sizes[1]), 1), dtype=np.uint16)
        mr = ctypes.c_uint16(0)
        st = lib().orc_chf_build_regions(self._h, border_size, min_region_area, merge_region_area, _p(reg), ctypes.byref(mr))
        if st:
            raise OracleError(st)
        return reg[:int(sizes[1])], int(mr.value)

    def get_neighbor(self, span, dir8):
        r = lib().orc_chf_get_neighbor(self._h, span, dir8)
        return None if r == NONE else int(r)

    def export(self) -> ChfArrays:
        sz = np.zeros(6, dtype=np.uint64)
        lib().orc_chf_sizes(self._h, _p(sz))
        C, S, K = int(sz[0]), int(sz[1]), int(sz[2])
        arrs = [np.zeros(C, np.uint32), np.zeros(C, np.uint32), np.zeros(S, np.int16), np.zeros(S, np.int16),
                np.zeros(S, np.uint8), np.zeros(S, np.uint8), np.zeros(S * 4, np.uint16), np.zeros(S, np.uint32),
                np.zeros(K, np.uint32), np.zeros(K, np.uint8), np.zeros(K, np.uint32), np.zeros(S, np.uint16)]
        lib().orc_chf_export(self._h, *[_p(a) for a in arrs])
        return ChfArrays(*arrs, walkable_span_count=int(sz[3]), max_height=int(sz[4]), max_distance=int(sz[5]))


# ---- config.rs ---------------------------------------------------------------------------------
def config_default() -> RecastConfig:
    c = RcbConfig()
    lib().orc_config_default(ctypes.byref(c))
    return RecastConfig.from_c(c)


def calculate_grid_size(cfg: RecastConfig, bmin, bmax) -> RecastConfig:
    c = cfg.to_c()
    a, b = _f3(bmin), _f3(bmax)
    lib().orc_config_calculate_grid_size(ctypes.byref(c), _p(a), _p(b))
    return RecastConfig.from_c(c)


def config_validate(cfg: RecastConfig) -> int:
    c = cfg.to_c()
    return lib().orc_config_validate(ctypes.byref(c))


# ---- per-tile pipeline, returns (HfArrays, ChfArrays|None) ----------------------------------------
def build_tile(cfg: RecastConfig, verts, indices, stage_mask=1 | 2 | 4 | 16, erode_radius=0):
    hf = Heightfield.from_config(cfg)
    if stage_mask & 1:
        hf.builder_rasterize(cfg, verts, indices, run_filters=bool(stage_mask & 2))
    hfa = hf.export()
    if not stage_mask & 4:
        return hfa, None
    chf = CompactHeightfield.build_from_heightfield(hf)
    if stage_mask & 8:
        chf.erode_walkable_area(erode_radius)
    if stage_mask & 16:
        chf.build_distance_field()
    return hfa, chf.export()


def build_tiles_timed(verts, indices, cfgs, stage_mask=1 | 2 | 4 | 16, erode_radius=0, nthreads=1, prebinned=False):
    """CPU baseline: returns (seconds, totals dict)."""
    v = np.ascontiguousarray(verts, dtype=np.float32).reshape(-1)
    t = np.ascontiguousarray(indices, dtype=np.int32).reshape(-1)
    arr = configs_to_array(cfgs)
    tot = np.zeros(4, dtype=np.uint64)
    s = lib().orc_build_tiles(_p(v), v.size, _p(t), t.size, ctypes.cast(arr, ctypes.c_void_p), len(cfgs),
                              stage_mask, erode_radius, nthreads, int(prebinned), _p(tot))
    if s < 0:
        raise OracleError(int(s))
    return s, {"spans": int(tot[0]), "connections": int(tot[1]), "walkable": int(tot[2]), "checksum": int(tot[3])}


def hardware_concurrency() -> int:
    return int(lib().orc_hardware_concurrency())


# ---- tile-cache layers (detour-tilecache/src/tile_cache_builder.rs:100-320, 422-549) -------------------------
def tilecache_build(layers, obstacles, cs, ch):
    """Per layer: ChfArrays, or the rcb_status (int) the reference's Err maps to."""
    from recast_rs_b200 import tilecache as tcm

    la, oa, keep = tcm.pack(layers, obstacles)
    out = []
    for i, l in enumerate(layers):
        st = ctypes.c_int(0)
        h = lib().orc_tilecache_build(ctypes.byref(la[i]), ctypes.cast(oa, ctypes.c_void_p), cs, ch, ctypes.byref(st))
        if not h:
            out.append(int(st.value))
            continue
        out.append(CompactHeightfield(h, l.width, l.height).export())
    del keep
    return out


def tilecache_build_timed(layers, obstacles, cs, ch, nthreads=1):
    from recast_rs_b200 import tilecache as tcm

    la, oa, keep = tcm.pack(layers, obstacles)
    tot = np.zeros(2, dtype=np.uint64)
    s = lib().orc_tilecache_build_batch(ctypes.cast(la, ctypes.c_void_p), len(layers), ctypes.cast(oa, ctypes.c_void_p), cs, ch,
                                        nthreads, _p(tot))
    del keep
    return s, {"spans": int(tot[0]), "connections": int(tot[1])}


# ---- tile-cache front half (detour-tilecache) ----------------------------------------------------------
def decompress_tile(blob) -> bytes:
    """TileCache::decompress_tile (tile_cache.rs:854-872); OracleError(-7) where lz4_flex returns Err."""
    d = np.frombuffer(bytes(blob), np.uint8)
    if d.size == 0:
        return b""
    size = lib().orc_lz4_prepended_size(_p(d), d.size)
    if size < 0:
        raise OracleError(-7)
    out = np.zeros(max(size, 1), np.uint8)
    prod = ctypes.c_uint64(0)
    st = lib().orc_lz4_decompress_size_prepended(_p(d), d.size, _p(out), size, ctypes.byref(prod))
    if st:
        raise OracleError(st)
    return out[:prod.value].tobytes()


def tilecache_layer_from_bytes(data, first_obstacle=0, obstacle_count=0):
    """TileCacheLayer::from_bytes (tile_cache_data.rs:281-320) -> tilecache.TileCacheLayer; OracleError(-8) on Err."""
    from recast_rs_b200 import tilecache as tcm

    d = np.frombuffer(bytes(data), np.uint8)
    hdr = tcm.RcbTcLayer()
    cons = ctypes.c_uint32(0)
    st = lib().orc_tilecache_layer_from_bytes(_p(d) if d.size else None, d.size, ctypes.byref(hdr), ctypes.byref(cons))
    if st:
        raise OracleError(st)
    ro = hdr.regons - d.ctypes.data
    ao = hdr.areas - d.ctypes.data
    return tcm.TileCacheLayer(tuple(hdr.bmin), tuple(hdr.bmax), hdr.hmin, hdr.hmax, hdr.width, hdr.height,
                              d[ro:ro + hdr.regons_len].copy(), d[ao:ao + hdr.areas_len].copy(), first_obstacle, obstacle_count)


def tilecache_build_compressed(blobs, obstacles, ranges, cs, ch):
    """Per compressed tile: ChfArrays, or the rcb_status of the first step that fails."""
    out = []
    for i, b in enumerate(blobs):
        fo, oc = ranges[i] if ranges is not None else (0, 0)
        try:
            layer = tilecache_layer_from_bytes(decompress_tile(b), fo, oc)
        except OracleError as e:
            out.append(e.status)
            continue
        out.append(tilecache_build([layer], obstacles, cs, ch)[0])
    return out


def tilecache_build_compressed_timed(ct, cs, ch, nthreads=1):
    """ct: recast_rs_b200.tilecache.CompressedTiles.  Returns (seconds, totals)."""
    tot = np.zeros(3, dtype=np.uint64)
    s = lib().orc_tilecache_build_compressed_batch(_p(ct.buf), _p(ct.off), ct.n, _p(ct.first) if ct.first is not None else None,
                                                   _p(ct.count) if ct.count is not None else None,
                                                   ctypes.cast(ct.obstacles, ctypes.c_void_p), cs, ch, nthreads, _p(tot))
    return s, {"spans": int(tot[0]), "connections": int(tot[1]), "failed": int(tot[2])}


# ---- convex volumes (crates/recast/src/convex_volume.rs) ----------------------------------------------
def convex_is_convex(verts) -> bool:
    v = np.ascontiguousarray(verts, dtype=np.float32).reshape(-1, 3)
    return bool(lib().orc_convex_is_convex(_p(v), len(v)))


def convex_volume_validate(verts, min_height, max_height) -> int:
    v = np.ascontiguousarray(verts, dtype=np.float32).reshape(-1, 3)
    return int(lib().orc_convex_volume_validate(_p(v), len(v), min_height, max_height))


def convex_contains_point(verts, min_height, max_height, point) -> bool:
    v = np.ascontiguousarray(verts, dtype=np.float32).reshape(-1, 3)
    p = _f3(point)
    return bool(lib().orc_convex_contains_point(_p(v), len(v), min_height, max_height, _p(p)))
