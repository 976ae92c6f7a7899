"""Tile sharding across GPUs.

Tiles are independent units: a tile's result depends only on the mesh and its own RecastConfig
(rasterization clips against the tile's own bounds, out-of-bounds neighbours are ledges,
`border_size` is 0 on every multi-tile path of the reference: detour-dynamic/src/dynamic_tile.rs:549).
So ranks never exchange data on the build path; the only communication is the optional gather of
per-tile counts at the end.

Pure host logic (numpy + optional torch.distributed); no CUDA here.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np

from .config import RecastConfig


def shard_tiles(ntiles: int, rank: int, world: int, mode: str = "block_cyclic") -> np.ndarray:
    """Tile ids owned by `rank`.  block_cyclic: tile i -> rank i % world;
    slab: contiguous ranges (row-major tile order => bands of tile rows, best for pre-binned meshes)."""
    ids = np.arange(ntiles, dtype=np.int64)
    if mode == "block_cyclic":
        return ids[rank::world]
    if mode == "slab":
        lo = (ntiles * rank) // world
        hi = (ntiles * (rank + 1)) // world
        return ids[lo:hi]
    raise ValueError(mode)


def prebin_mesh(verts: np.ndarray, tris: np.ndarray, cfgs: Sequence[RecastConfig], compact_vertices: bool = True
                ) -> Tuple[np.ndarray, np.ndarray]:
    """Sub-mesh containing, IN THE ORIGINAL ORDER, every triangle whose AABB touches the union AABB of the
    given tiles (inclusive test in f32, the same comparison as rasterization.rs:370-377, so no triangle the
    reference would rasterize into these tiles is dropped).  Order matters: span areas depend on it.
    compact_vertices: keep only the vertices those triangles use (in their original order) and renumber the
    indices, so a rank uploads its slab's share of the vertex array instead of the whole world's; the
    coordinates are untouched, so every tile's result is the same bytes.  Returns (verts, tris_subset)."""
    if len(cfgs) == 0:
        return (verts[:0] if compact_vertices else verts), tris[:0]
    v = verts.reshape(-1, 3)
    t = tris.reshape(-1, 3)
    bmin = np.array([c.bmin for c in cfgs], dtype=np.float32).min(axis=0)
    bmax = np.array([c.bmax for c in cfgs], dtype=np.float32).max(axis=0)
    tv = v[t]  # [T,3,3]
    tmin, tmax = tv.min(axis=1), tv.max(axis=1)
    keep = np.all((tmin <= bmax) & (tmax >= bmin), axis=1)
    sub = np.ascontiguousarray(t[keep])
    if not compact_vertices:
        return verts, sub.reshape(-1)
    used = np.zeros(v.shape[0], dtype=bool)
    used[sub.reshape(-1)] = True
    remap = np.cumsum(used, dtype=np.int64) - 1  # old vertex id -> new id (order preserved)
    return np.ascontiguousarray(v[used]).reshape(-1), remap[sub].astype(np.int32).reshape(-1)


def tile_costs(verts: np.ndarray, tris: np.ndarray, cfgs: Sequence[RecastConfig]) -> np.ndarray:
    """Estimated build cost per tile: the cell footprints of the triangles over it (what the clipper's work and
    the fragment count follow) plus its cells (what everything behind the clipper follows).  Cheap: every
    triangle's footprint goes to the tile holding the centre of its AABB."""
    n = len(cfgs)
    if n == 0:
        return np.zeros(0)
    v = verts.reshape(-1, 3)
    tv = v[tris.reshape(-1, 3)]
    tmin, tmax = tv.min(axis=1), tv.max(axis=1)
    cs = float(cfgs[0].cs)
    foot = (np.floor((tmax[:, 0] - tmin[:, 0]) / cs) + 1.0) * (np.floor((tmax[:, 2] - tmin[:, 2]) / cs) + 1.0)
    bmin = np.array([c.bmin for c in cfgs], dtype=np.float64)
    bmax = np.array([c.bmax for c in cfgs], dtype=np.float64)
    xs, zs = np.unique(bmin[:, 0]), np.unique(bmin[:, 2])  # the tiles form a grid (synth.tile_grid)
    cx, cz = 0.5 * (tmin[:, 0] + tmax[:, 0]), 0.5 * (tmin[:, 2] + tmax[:, 2])
    ix = np.clip(np.searchsorted(xs, cx, side="right") - 1, 0, len(xs) - 1)
    iz = np.clip(np.searchsorted(zs, cz, side="right") - 1, 0, len(zs) - 1)
    grid = np.zeros((len(zs), len(xs)))
    np.add.at(grid, (iz, ix), foot)
    tx = np.searchsorted(xs, bmin[:, 0])
    tz = np.searchsorted(zs, bmin[:, 2])
    cells = np.array([c.width * c.height for c in cfgs], dtype=np.float64)
    return grid[tz, tx] + 2.0 * cells


def cut_by_cost(cost: np.ndarray, world: int) -> List[int]:
    """world + 1 tile indices: slab k = tiles [cuts[k], cuts[k+1]), cut where the cumulative cost crosses k/world of the total."""
    cum = np.cumsum(cost)
    total = cum[-1] if len(cum) else 0.0
    cuts = [0] + [int(np.searchsorted(cum, total * k / world)) for k in range(1, world)] + [len(cost)]
    for k in range(1, len(cuts)):
        cuts[k] = max(cuts[k], cuts[k - 1])
    return cuts


def shard_tiles_weighted(verts: np.ndarray, tris: np.ndarray, cfgs: Sequence[RecastConfig], rank: int, world: int,
                         cost: np.ndarray | None = None) -> np.ndarray:
    """Contiguous slabs like shard_tiles(mode='slab'), but cut where the cumulative estimated COST (tile_costs, or the
    given per-tile costs) crosses k/world of the total: uneven worlds (config 3's buildings) then take about the same
    time on every rank.  The static alternative to a work-stealing queue (BASELINE.json north_star allows either)."""
    cuts = cut_by_cost(tile_costs(verts, tris, cfgs) if cost is None else cost, world)
    return np.arange(cuts[rank], cuts[rank + 1], dtype=np.int64)


def refine_costs(cost: np.ndarray, cuts: Sequence[int], measured_ms: Sequence[float]) -> np.ndarray:
    """One round of feedback for a static partition: every slab's tiles are re-priced so that the slab's estimated cost
    matches the time its rank measured (the estimate's error is mostly a per-region factor: how many buildings, how
    steep).  Cutting the re-priced costs again (cut_by_cost) moves tiles from slow ranks to fast ones."""
    out = np.asarray(cost, dtype=np.float64).copy()
    for k, ms in enumerate(measured_ms):
        lo, hi = cuts[k], cuts[k + 1]
        est = out[lo:hi].sum()
        if hi > lo and est > 0 and ms > 0:
            out[lo:hi] *= ms / est
    return out


def gather_tile_counts(local_ids: np.ndarray, local_counts: np.ndarray, ntiles: int, world: int) -> np.ndarray:
    """All-gather per-tile counters (rows = tiles, columns = e.g. spans, connections, walkable) so every
    rank holds the whole table in tile order.  Works with any initialised torch.distributed backend
    (nccl on GPUs, gloo in the CPU tests); with world == 1 it is a local scatter."""
    cols = local_counts.shape[1] if local_counts.ndim == 2 else 1
    table = np.zeros((ntiles, cols), dtype=np.int64)
    lc = local_counts.reshape(len(local_ids), cols)
    if world == 1:
        table[local_ids] = lc
        return table
    import torch
    import torch.distributed as dist

    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    n_max = (ntiles + world - 1) // world + 1
    buf = torch.full((n_max, cols + 1), -1, dtype=torch.int64, device=dev)
    buf[:len(local_ids), 0] = torch.from_numpy(np.asarray(local_ids, dtype=np.int64)).to(dev)
    buf[:len(local_ids), 1:] = torch.from_numpy(lc.astype(np.int64)).to(dev)
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf)
    for o in out:
        o = o.cpu().numpy()
        valid = o[:, 0] >= 0
        table[o[valid, 0]] = o[valid, 1:]
    return table


def weak_scaling_world(base_nq: int, world: int) -> int:
    """Quads per side so that the world has ~world x the triangles of the single-GPU workload."""
    return int(round(base_nq * float(np.sqrt(world))))
