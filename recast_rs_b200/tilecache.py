"""Tile-cache layer inputs (SURVEY §8(a) row TC): plain-data mirrors of TileCacheLayer / TileCacheLayerHeader
(crates/detour-tilecache/src/tile_cache_data.rs:21-52, 238-247) and Obstacle / ObstacleData / ObstacleState
(tile_cache.rs:109-152), plus their packing into the C-ABI structs rcb_tc_layer / rcb_tc_obstacle.
No CUDA here; used by builder.Context.build_tilecache_layers and by the oracle wrapper."""
from __future__ import annotations

import ctypes
import struct
from dataclasses import dataclass, field
from typing import List, Sequence, Tuple

import numpy as np

CYLINDER, BOX, ORIENTED_BOX = 0, 1, 2
EMPTY, PROCESSING, PROCESSED, REMOVING = 0, 1, 2, 3


class RcbTcLayer(ctypes.Structure):
    _fields_ = [("bmin", ctypes.c_float * 3), ("bmax", ctypes.c_float * 3), ("hmin", ctypes.c_uint16), ("hmax", ctypes.c_uint16),
                ("width", ctypes.c_uint8), ("height", ctypes.c_uint8), ("_pad", ctypes.c_uint16),
                ("regons", ctypes.c_void_p), ("areas", ctypes.c_void_p), ("regons_len", ctypes.c_uint32),
                ("areas_len", ctypes.c_uint32), ("first_obstacle", ctypes.c_uint32), ("obstacle_count", ctypes.c_uint32)]


class RcbTcObstacle(ctypes.Structure):
    _fields_ = [("type", ctypes.c_int32), ("state", ctypes.c_int32), ("a", ctypes.c_float * 3), ("b", ctypes.c_float * 3),
                ("rot_aux", ctypes.c_float * 2)]


@dataclass
class TileCacheLayer:
    bmin: Sequence[float]
    bmax: Sequence[float]
    hmin: int
    hmax: int
    width: int
    height: int
    regons: np.ndarray  # uint8 [width*height], 0 = empty
    areas: np.ndarray   # uint8 [width*height]
    first_obstacle: int = 0
    obstacle_count: int = 0
    cons: "np.ndarray | None" = None  # uint8 [width*height] neighbour-connection nibbles (TileCacheLayer.cons); unused by the voxel path


TILECACHE_MAGIC, TILECACHE_VERSION = 0x4C494554, 1  # tile_cache_data.rs:10-13
_LAYER_HDR = struct.Struct("<IIiii6fHH6B")  # TileCacheLayerHeader::to_bytes (tile_cache_data.rs:94-125): 54 bytes


def layer_to_bytes(layer: TileCacheLayer, cons=None, tx=0, ty=0, tlayer=0, minx=0, maxx=0, miny=0, maxy=0) -> bytes:
    """TileCacheLayer::to_bytes (tile_cache_data.rs:261-278): header, three u32 lengths, regons, areas, cons."""
    r = np.ascontiguousarray(layer.regons, np.uint8).tobytes()
    a = np.ascontiguousarray(layer.areas, np.uint8).tobytes()
    c = b"" if cons is None else np.ascontiguousarray(cons, np.uint8).tobytes()
    h = _LAYER_HDR.pack(TILECACHE_MAGIC, TILECACHE_VERSION, tx, ty, tlayer, *[float(x) for x in layer.bmin], *[float(x) for x in layer.bmax],
                        int(layer.hmin), int(layer.hmax), int(layer.width), int(layer.height), minx, maxx, miny, maxy)
    return h + struct.pack("<III", len(r), len(a), len(c)) + r + a + c


def layer_from_bytes(data: bytes) -> TileCacheLayer:
    """TileCacheLayer::from_bytes (tile_cache_data.rs:281-320), host mirror; ValueError where the reference returns
    Error::Detour(InvalidParam)."""
    if len(data) < 66:
        raise ValueError("InvalidParam")
    f = _LAYER_HDR.unpack_from(data, 0)
    if f[0] != TILECACHE_MAGIC or f[1] != TILECACHE_VERSION:
        raise ValueError("InvalidParam")
    rl, al, cl = struct.unpack_from("<III", data, 54)
    if 66 + rl + al + cl > len(data):
        raise ValueError("InvalidParam")
    return TileCacheLayer(f[5:8], f[8:11], f[11], f[12], f[13], f[14], np.frombuffer(data, np.uint8, rl, 66).copy(),
                          np.frombuffer(data, np.uint8, al, 66 + rl).copy())


class CompressedTiles:
    """Compressed tiles (TileCache.compressed_tiles, tile_cache.rs:34) joined into one buffer with offsets, plus the
    obstacle table and the per-layer obstacle ranges: what a native caller hands to rcb_build_tilecache_compressed."""

    def __init__(self, blobs: Sequence[bytes], obstacles: Sequence["Obstacle"] = (), obstacle_ranges=None):
        self.n = len(blobs)
        self.off = np.zeros(self.n + 1, np.uint64)
        self.off[1:] = np.cumsum([len(b) for b in blobs], dtype=np.uint64)
        self.buf = np.frombuffer(b"".join(bytes(b) for b in blobs), np.uint8)
        _, self.obstacles, self._keep = pack([], obstacles or [])
        self.nobstacles = len(obstacles or [])
        if self.nobstacles:
            rng = obstacle_ranges if obstacle_ranges is not None else [(0, 0)] * self.n
            self.first = np.ascontiguousarray([r[0] for r in rng], dtype=np.uint32)
            self.count = np.ascontiguousarray([r[1] for r in rng], dtype=np.uint32)
        else:
            self.first = self.count = None


@dataclass
class Obstacle:
    type: int
    state: int = PROCESSED
    a: Sequence[float] = (0.0, 0.0, 0.0)   # cylinder pos | box bmin | oriented-box center
    b: Sequence[float] = (0.0, 0.0, 0.0)   # (radius, height, -) | box bmax | half extents
    rot_aux: Sequence[float] = (0.0, 0.0)

    @classmethod
    def cylinder(cls, pos, radius, height, state=PROCESSED):
        return cls(CYLINDER, state, tuple(pos), (radius, height, 0.0))

    @classmethod
    def box(cls, bmin, bmax, state=PROCESSED):
        return cls(BOX, state, tuple(bmin), tuple(bmax))

    @classmethod
    def oriented_box(cls, center, half_extents, y_radians, state=PROCESSED):
        """rot_aux as the reference defines it (tile_cache.rs:149): [cos(a/2)sin(-a/2), cos(a/2)cos(a/2) - 0.5], in f32."""
        f = np.float32
        c, s = f(np.cos(f(0.5) * f(y_radians))), f(np.sin(f(-0.5) * f(y_radians)))
        return cls(ORIENTED_BOX, state, tuple(center), tuple(half_extents), (float(f(c * s)), float(f(f(c * c) - f(0.5)))))


def pack(layers: Sequence[TileCacheLayer], obstacles: Sequence[Obstacle]):
    """-> (rcb_tc_layer[], rcb_tc_obstacle[], keepalive list of the numpy buffers the structs point to)."""
    la = (RcbTcLayer * max(len(layers), 1))()
    keep = []
    for i, l in enumerate(layers):
        r = np.ascontiguousarray(l.regons, dtype=np.uint8).reshape(-1)
        a = np.ascontiguousarray(l.areas, dtype=np.uint8).reshape(-1)
        keep += [r, a]
        la[i].bmin = (ctypes.c_float * 3)(*[float(x) for x in l.bmin])
        la[i].bmax = (ctypes.c_float * 3)(*[float(x) for x in l.bmax])
        la[i].hmin, la[i].hmax, la[i].width, la[i].height = int(l.hmin), int(l.hmax), int(l.width), int(l.height)
        la[i].regons = r.ctypes.data if r.size else None
        la[i].areas = a.ctypes.data if a.size else None
        la[i].regons_len, la[i].areas_len = r.size, a.size
        la[i].first_obstacle, la[i].obstacle_count = int(l.first_obstacle), int(l.obstacle_count)
    oa = (RcbTcObstacle * max(len(obstacles), 1))()
    for i, o in enumerate(obstacles):
        oa[i].type, oa[i].state = int(o.type), int(o.state)
        oa[i].a = (ctypes.c_float * 3)(*[float(x) for x in o.a])
        oa[i].b = (ctypes.c_float * 3)(*[float(x) for x in o.b])
        oa[i].rot_aux = (ctypes.c_float * 2)(*[float(x) for x in o.rot_aux])
    return la, oa, keep


class Packed:
    """rcb_tc_layer[] / rcb_tc_obstacle[] built once from the Python-side descriptions (what a native caller
    holds already); keeps the numpy buffers the structs point to alive."""

    def __init__(self, layers: Sequence[TileCacheLayer], obstacles: Sequence[Obstacle]):
        self.layers, self.obstacles, self._keep = pack(layers, obstacles)
        self.nlayers, self.nobstacles = len(layers), len(obstacles)


def config5(ntiles: int = 256, size: int = 48, cs: float = 0.3, ch: float = 0.2, seed: int = 1
            ) -> Tuple[List[TileCacheLayer], List[Obstacle]]:
    """BASELINE config 5: `ntiles` dirty tile-cache layers of size x size cells, regons = 1, areas = 63,
    hmin = 10, bmin from a square tile grid, 1-3 hashed cylinder / box / oriented-box obstacles per tile."""
    from .synth import hash_unit

    side = int(np.ceil(np.sqrt(ntiles)))
    layers, obstacles = [], []
    ext = size * cs
    for t in range(ntiles):
        tx, tz = t % side, t // side
        bmin = (np.float32(tx * ext), np.float32(0.0), np.float32(tz * ext))
        bmax = (np.float32((tx + 1) * ext), np.float32(8.0), np.float32((tz + 1) * ext))
        h = lambda k: float(hash_unit(np.array([t]), np.array([k]), seed)[0]) * 0.5 + 0.5  # noqa: E731
        nobs = 1 + int(h(0) * 3) % 3
        first = len(obstacles)
        for k in range(nobs):
            cx, cz = float(bmin[0]) + ext * h(10 + k), float(bmin[2]) + ext * h(20 + k)
            kind = int(h(30 + k) * 3) % 3
            if kind == 0:
                obstacles.append(Obstacle.cylinder((cx, 1.5, cz), 0.5 + 2.0 * h(40 + k), 2.0))
            elif kind == 1:
                w, d = 0.5 + 2.5 * h(50 + k), 0.5 + 2.5 * h(60 + k)
                obstacles.append(Obstacle.box((cx - w, 1.0, cz - d), (cx + w, 4.0, cz + d)))
            else:
                obstacles.append(Obstacle.oriented_box((cx, 2.5, cz), (0.5 + 2.0 * h(70 + k), 1.5, 0.4 + 1.5 * h(80 + k)), 3.0 * h(90 + k)))
        layers.append(TileCacheLayer(bmin, bmax, 10, 11, size, size, np.ones(size * size, np.uint8),
                                     np.full(size * size, 63, np.uint8), first, nobs))
    return layers, obstacles


def config5_varied(ntiles: int = 256, size: int = 48, cs: float = 0.3, ch: float = 0.2, seed: int = 1
                   ) -> Tuple[List[TileCacheLayer], List[Obstacle]]:
    """config5 with layer content that looks like a built tile cache instead of constants: region ids in
    patches with ~10 % empty cells, areas mostly 63 with patches of other ids, `cons` = the 4-neighbour
    occupancy nibble.  Same obstacles as config5.  LZ4 gets these to roughly a third of their size."""
    layers, obstacles = config5(ntiles, size, cs, ch, seed)
    for t, l in enumerate(layers):
        rng = np.random.default_rng(seed * 100003 + t)
        coarse = rng.integers(1, 9, (size // 6 + 2, size // 6 + 2), dtype=np.uint8)
        reg = np.kron(coarse, np.ones((6, 6), np.uint8))[:size, :size].copy()
        holes = np.kron((rng.random((size // 4 + 1, size // 4 + 1)) < 0.1).astype(np.uint8), np.ones((4, 4), np.uint8))[:size, :size]
        reg[holes == 1] = 0
        reg[rng.random((size, size)) < 0.01] = 0
        area = np.full((size, size), 63, np.uint8)
        patch = np.kron(rng.integers(0, 12, (size // 8 + 1, size // 8 + 1), dtype=np.uint8), np.ones((8, 8), np.uint8))[:size, :size]
        area[patch == 1] = 5
        area[patch == 2] = 0
        occ = np.pad(reg != 0, 1)
        cons = (occ[1:-1, :-2] * 1 + occ[2:, 1:-1] * 2 + occ[1:-1, 2:] * 4 + occ[:-2, 1:-1] * 8).astype(np.uint8)
        cons[reg == 0] = 0
        l.regons, l.areas, l.cons = reg.reshape(-1), area.reshape(-1), cons.reshape(-1)
        l.hmin = int(rng.integers(0, 40))
        l.hmax = l.hmin + 1
    return layers, obstacles
