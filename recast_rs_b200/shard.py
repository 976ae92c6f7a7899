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


def prebin_mesh(verts: np.ndarray, tris: np.ndarray, cfgs: Sequence[RecastConfig]) -> Tuple[np.ndarray, np.ndarray]:
    """Sub-mesh containing, IN THE ORIGINAL ORDER, every triangle whose AABB touches the union AABB of the
    given tiles (inclusive test in f32, the same comparison as rasterization.rs:370-377, so no triangle the
    reference would rasterize into these tiles is dropped).  Order matters: span areas depend on it.
    Vertices are kept whole (indices stay valid); returns (verts, tris_subset)."""
    if len(cfgs) == 0:
        return verts, tris[:0]
    v = verts.reshape(-1, 3)
    t = tris.reshape(-1, 3)
    bmin = np.array([c.bmin for c in cfgs], dtype=np.float32).min(axis=0)
    bmax = np.array([c.bmax for c in cfgs], dtype=np.float32).max(axis=0)
    tv = v[t]  # [T,3,3]
    tmin, tmax = tv.min(axis=1), tv.max(axis=1)
    keep = np.all((tmin <= bmax) & (tmax >= bmin), axis=1)
    return verts, np.ascontiguousarray(t[keep]).reshape(-1)


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
