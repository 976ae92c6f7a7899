"""Tile-cache layer inputs (SURVEY §8(a) row TC): plain-data mirrors of TileCacheLayer / TileCacheLayerHeader
(crates/detour-tilecache/src/tile_cache_data.rs:21-52, 238-247) and Obstacle / ObstacleData / ObstacleState
(tile_cache.rs:109-152), plus their packing into the C-ABI structs rcb_tc_layer / rcb_tc_obstacle.
No CUDA here; used by builder.Context.build_tilecache_layers and by the oracle wrapper."""
from __future__ import annotations

import ctypes
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
