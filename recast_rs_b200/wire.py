"""Heightfield wire formats of detour-dynamic, on the GPU.

Mirrors, with the reference's names:
  VoxelTile            detour-dynamic/src/io/voxel_tile.rs:9-259   (from_heightfield, heightfield, write_to, read_from)
  DynamicTile.reconstruct_heightfield   detour-dynamic/src/dynamic_tile.rs:147-257
  DynamicTileCheckpoint.clone_heightfield   detour-dynamic/src/checkpoint.rs:29-60

The per-tile methods call the batched C ABI with one tile; `voxel_tiles_from_result` and
`heightfields_from_voxel_tiles` are the batched forms (one launch sequence for any number of tiles).
Byte order: VoxelTile writes and reads big-endian; DynamicTile.reconstruct_heightfield reads little-endian
(the reference's own pair disagrees — reproduced, not repaired).
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field
from typing import BinaryIO, List, Optional, Sequence, Set, Tuple

import numpy as np

from .builder import (STAGE_READD, WIRE_DYNAMIC_TILE, WIRE_VOXEL_TILE, BatchResult, Context, Heightfield, default_context)
from .config import RecastConfig

_HDR = struct.Struct(">iiiii" + "f" * 8 + "I")  # voxel_tile.rs:178-199: everything big-endian


def _hf_cfg(width, depth, bmin, bmax, cs, ch) -> RecastConfig:
    return RecastConfig(width=int(width), height=int(depth), cs=float(cs), ch=float(ch), bmin=tuple(bmin), bmax=tuple(bmax))


def _hf_from_tile(tv, cfg: RecastConfig, ctx: Context) -> Heightfield:
    return Heightfield(cfg.width, cfg.height, cfg.bmin, cfg.bmax, cfg.cs, cfg.ch, tv.hf_cell_offset.copy(), tv.span_min.copy(),
                       tv.span_max.copy(), tv.span_area.copy(), ctx=ctx)


@dataclass
class VoxelTile:
    """io/voxel_tile.rs:9-21."""
    tile_x: int
    tile_z: int
    width: int
    depth: int
    bounds_min: Tuple[float, float, float]
    bounds_max: Tuple[float, float, float]
    cell_size: float
    cell_height: float
    border_size: int = 0
    span_data: bytes = b""

    @classmethod
    def from_heightfield(cls, tile_x: int, tile_z: int, heightfield: Heightfield, ctx: Optional[Context] = None) -> "VoxelTile":
        """io/voxel_tile.rs:50-66 (border_size is not stored in a Heightfield: 0)."""
        ctx = ctx or heightfield._ctx or default_context()
        hf = heightfield
        with ctx.build_from_spans([hf._cfg()], hf.cell_offset, hf.span_min, hf.span_max, hf.span_area, 0) as r:
            data, _ = r.serialize_spans(WIRE_VOXEL_TILE)
        return cls(tile_x, tile_z, hf.width, hf.height, hf.bmin, hf.bmax, hf.cs, hf.ch, 0, data.tobytes())

    def _cfg(self) -> RecastConfig:
        return _hf_cfg(self.width, self.depth, self.bounds_min, self.bounds_max, self.cell_size, self.cell_height)

    def heightfield(self, ctx: Optional[Context] = None) -> Heightfield:
        """io/voxel_tile.rs:68, 120-176: lenient big-endian reader."""
        return heightfields_from_voxel_tiles([self], WIRE_VOXEL_TILE, ctx)[0]

    def write_to(self, w: BinaryIO) -> None:
        """io/voxel_tile.rs:178-199."""
        w.write(_HDR.pack(self.tile_x, self.tile_z, self.width, self.depth, self.border_size, *self.bounds_min, *self.bounds_max,
                          self.cell_size, self.cell_height, len(self.span_data)))
        w.write(self.span_data)

    @classmethod
    def read_from(cls, r: BinaryIO) -> "VoxelTile":
        """io/voxel_tile.rs:201-258 (`read_exact`: a short stream is an error)."""
        h = r.read(_HDR.size)
        if len(h) != _HDR.size:
            raise EOFError("failed to fill whole buffer")
        f = _HDR.unpack(h)
        data = r.read(f[13])
        if len(data) != f[13]:
            raise EOFError("failed to fill whole buffer")
        return cls(f[0], f[1], f[2], f[3], tuple(f[5:8]), tuple(f[8:11]), f[11], f[12], f[4], data)


def voxel_tiles_from_result(result: BatchResult, cfgs: Sequence[RecastConfig], coords: Optional[Sequence[Tuple[int, int]]] = None,
                            fmt: int = WIRE_VOXEL_TILE) -> List[VoxelTile]:
    """VoxelTile::from_heightfield for every tile of a batch result: one serialization kernel, one copy."""
    data, off = result.serialize_spans(fmt)
    raw = data.tobytes()
    out = []
    for i, c in enumerate(cfgs):
        tx, tz = coords[i] if coords is not None else (i, 0)
        out.append(VoxelTile(tx, tz, c.width, c.height, tuple(c.bmin), tuple(c.bmax), c.cs, c.ch, 0, raw[int(off[i]):int(off[i + 1])]))
    return out


def build_from_voxel_tiles(tiles: Sequence[VoxelTile], fmt: int, stage_mask: int = 0, erode_radius: int = 0,
                           ctx: Optional[Context] = None, cfgs: Optional[Sequence[RecastConfig]] = None) -> BatchResult:
    """Parse a batch of span streams on the device and run the later stages of stage_mask on the heightfields."""
    ctx = ctx or default_context()
    cfgs = list(cfgs) if cfgs is not None else [t._cfg() for t in tiles]
    off = np.zeros(len(tiles) + 1, np.uint64)
    off[1:] = np.cumsum([len(t.span_data) for t in tiles], dtype=np.uint64)
    buf = np.frombuffer(b"".join(t.span_data for t in tiles), np.uint8)  # the device reader is bytewise: no alignment needed
    return ctx.build_from_span_data(cfgs, buf, off, fmt, stage_mask, erode_radius)


def heightfields_from_voxel_tiles(tiles: Sequence[VoxelTile], fmt: int = WIRE_VOXEL_TILE, ctx: Optional[Context] = None) -> List[Heightfield]:
    ctx = ctx or default_context()
    cfgs = [t._cfg() for t in tiles]
    r = build_from_voxel_tiles(tiles, fmt, 0, 0, ctx, cfgs)
    with r:
        r.download()
        return [_hf_from_tile(r.tile(i), cfgs[i], ctx) for i in range(len(tiles))]


class DynamicTile:
    """The heightfield-reconstruction half of detour-dynamic's DynamicTile (dynamic_tile.rs:106-257)."""

    @staticmethod
    def reconstruct_heightfield(voxel_tile: VoxelTile, ctx: Optional[Context] = None) -> Heightfield:
        """dynamic_tile.rs:147-257: strict little-endian reader; raises errors.Recast / errors.NavMeshGeneration."""
        return heightfields_from_voxel_tiles([voxel_tile], WIRE_DYNAMIC_TILE, ctx)[0]


@dataclass
class DynamicTileCheckpoint:
    """checkpoint.rs:13-19."""
    heightfield: Heightfield
    colliders: Set[int] = field(default_factory=set)

    @classmethod
    def new(cls, heightfield: Heightfield, colliders: Set[int], ctx: Optional[Context] = None) -> "DynamicTileCheckpoint":
        return cls(cls.clone_heightfield(heightfield, ctx), set(colliders))  # checkpoint.rs:22-27

    @staticmethod
    def clone_heightfield(source: Heightfield, ctx: Optional[Context] = None) -> Heightfield:
        """checkpoint.rs:29-60: every span re-added through Heightfield::add_span (may merge touching same-area spans)."""
        ctx = ctx or source._ctx or default_context()
        cfg = source._cfg()
        with ctx.build_from_spans([cfg], source.cell_offset, source.span_min, source.span_max, source.span_area, STAGE_READD) as r:
            return _hf_from_tile(r.download().tile(0), cfg, ctx)

    def is_valid_for(self, active_colliders: Set[int]) -> bool:  # checkpoint.rs:62-68
        return set(active_colliders) <= self.colliders
