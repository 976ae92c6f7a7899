"""Convex volumes (crates/recast/src/convex_volume.rs), host mirror + the device call.

ConvexVolume / ConvexVolumeSet keep the reference's names and checks; `ConvexVolumeSet.apply_to_heightfield`
(convex_volume.rs:392-430) and `RecastBuilder.build_heightfield_with_volumes` (lib.rs:346-361) run on the GPU
through rcb_apply_convex_volumes / rcb_build_tiles_with_volumes."""
from __future__ import annotations

import ctypes
from dataclasses import dataclass, field
from typing import List, Sequence

import numpy as np

from .errors import InvalidMesh

MAX_CONVEX_VOLUME_VERTS = 32  # convex_volume.rs:12
f32 = np.float32


class RcbConvexVolume(ctypes.Structure):
    _fields_ = [("first_vert", ctypes.c_uint32), ("num_verts", ctypes.c_uint32), ("min_height", ctypes.c_float),
                ("max_height", ctypes.c_float), ("area_type", ctypes.c_uint32)]


@dataclass
class ConvexVolume:
    """convex_volume.rs:15-25; build with `new`, `from_box` or `from_cylinder` to get the reference's checks."""
    vertices: np.ndarray  # float32 [n, 3]
    min_height: float
    max_height: float
    area_type: int

    @classmethod
    def new(cls, vertices, min_height, max_height, area_type) -> "ConvexVolume":  # :28-68
        v = np.ascontiguousarray(vertices, dtype=np.float32).reshape(-1, 3)
        if len(v) < 3:
            raise InvalidMesh("Convex volume requires at least 3 vertices")
        if len(v) > MAX_CONVEX_VOLUME_VERTS:
            raise InvalidMesh(f"Convex volume has too many vertices: {len(v)} (max: {MAX_CONVEX_VOLUME_VERTS})")
        if f32(min_height) > f32(max_height):
            raise InvalidMesh("Convex volume min_height > max_height")
        if not cls.is_convex(v):
            raise InvalidMesh("Vertices do not form a convex polygon")
        return cls(v, float(f32(min_height)), float(f32(max_height)), int(area_type) & 0xFF)

    @classmethod
    def from_box(cls, center, half_extents, area_type) -> "ConvexVolume":  # :71-90
        c, h = np.asarray(center, f32), np.asarray(half_extents, f32)
        x0, x1, z0, z1 = c[0] - h[0], c[0] + h[0], c[2] - h[2], c[2] + h[2]
        return cls.new([(x0, c[1], z0), (x0, c[1], z1), (x1, c[1], z1), (x1, c[1], z0)], c[1] - h[1], c[1] + h[1], area_type)

    @classmethod
    def from_cylinder(cls, center, radius, height, segments, area_type) -> "ConvexVolume":  # :93-127
        if segments < 3:
            raise InvalidMesh("Cylinder requires at least 3 segments")
        if segments > MAX_CONVEX_VOLUME_VERTS:
            raise InvalidMesh(f"Cylinder has too many segments: {segments} (max: {MAX_CONVEX_VOLUME_VERTS})")
        c, r, hh = np.asarray(center, f32), f32(radius), f32(height)
        step = f32(f32(2.0) * f32(np.pi)) / f32(segments)
        verts = []
        for i in range(segments):
            a = f32(i) * step
            verts.append((c[0] + r * f32(np.cos(a)), c[1], c[2] + r * f32(np.sin(a))))
        return cls.new(verts, c[1] - hh * f32(0.5), c[1] + hh * f32(0.5), area_type)

    @staticmethod
    def is_convex(vertices) -> bool:  # :166-199
        v = np.asarray(vertices, f32).reshape(-1, 3)
        n = len(v)
        if n < 3:
            return False
        sign = f32(0.0)
        for i in range(n):
            j, k = (i + 1) % n, (i + 2) % n
            dx1, dz1 = v[j][0] - v[i][0], v[j][2] - v[i][2]
            dx2, dz2 = v[k][0] - v[j][0], v[k][2] - v[j][2]
            cross = f32(dx1 * dz2) - f32(dz1 * dx2)
            if cross != 0.0:
                if sign == 0.0:
                    sign = cross
                elif (cross > 0.0) != (sign > 0.0):
                    return False
        return True

    def contains_point(self, point) -> bool:  # :130-164
        p = np.asarray(point, f32)
        if p[1] < f32(self.min_height) or p[1] > f32(self.max_height):
            return False
        x, z, v = p[0], p[2], self.vertices
        inside, n = False, len(v)
        j = n - 1
        for i in range(n):
            vi, vj = v[i], v[j]
            j = i
            if (vi[2] > z) == (vj[2] > z):
                continue
            if x >= f32(f32(f32(vj[0] - vi[0]) * f32(z - vi[2])) / f32(vj[2] - vi[2])) + vi[0]:
                continue
            inside = not inside
        return inside


@dataclass
class ConvexVolumeSet:
    """convex_volume.rs:346-431."""
    volumes: List[ConvexVolume] = field(default_factory=list)

    def add_volume(self, volume: ConvexVolume) -> None:
        self.volumes.append(volume)

    def find_volumes_at_point(self, point) -> List[int]:  # :363-372
        return [i for i, v in enumerate(self.volumes) if v.contains_point(point)]

    def get_area_at_point(self, point, default_area: int) -> int:  # :375-389, later volumes win
        area = default_area
        for v in self.volumes:
            if v.contains_point(point):
                area = v.area_type
        return area

    def pack(self):
        """-> (rcb_convex_volume[], float32 vertex array) for the C ABI."""
        arr = (RcbConvexVolume * max(len(self.volumes), 1))()
        verts, first = [], 0
        for i, v in enumerate(self.volumes):
            arr[i].first_vert, arr[i].num_verts = first, len(v.vertices)
            arr[i].min_height, arr[i].max_height, arr[i].area_type = v.min_height, v.max_height, v.area_type
            verts.append(np.asarray(v.vertices, np.float32).reshape(-1, 3))
            first += len(v.vertices)
        flat = np.ascontiguousarray(np.concatenate(verts).reshape(-1) if verts else np.zeros(0, np.float32), dtype=np.float32)
        return arr, flat

    def apply_to_heightfield(self, heightfield, ctx=None) -> None:
        """convex_volume.rs:392-430, on the GPU: span.area of `heightfield` (builder.Heightfield) is rewritten in place."""
        from .builder import default_context

        ctx = ctx or heightfield._ctx or default_context()
        hf = heightfield
        with ctx.apply_convex_volumes([hf._cfg()], hf.cell_offset, hf.span_min, hf.span_max, hf.span_area, self, 0) as r:
            hf.span_area = r.download().tile(0).span_area.copy()
