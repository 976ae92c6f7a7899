"""Deterministic synthetic inputs named by BASELINE.md §3 / SURVEY.md §8(d).

Meshes are generated once on the host in float64 (libm via numpy) and rounded to f32; the identical
buffers feed the oracle and the GPU, so generator numerics never enter a parity check.
No library RNG: the roughness term is a splitmix64 hash of (i, j, seed).
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np

from .config import RecastConfig

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(x: np.ndarray) -> np.ndarray:
    """splitmix64 finaliser on uint64 arrays (wrapping arithmetic)."""
    with np.errstate(over="ignore"):
        x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M64
        z = x
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
        return z ^ (z >> np.uint64(31))


def hash_unit(i: np.ndarray, j: np.ndarray, seed: int) -> np.ndarray:
    """h(i,j,seed) in [-1, 1)."""
    with np.errstate(over="ignore"):
        k = (i.astype(np.uint64) * np.uint64(0x100000001B3)) ^ (j.astype(np.uint64) << np.uint64(32)) ^ np.uint64(
            seed * 0x9E3779B1 + 12345)
    r = splitmix64(k)
    return (r >> np.uint64(11)).astype(np.float64) * (2.0 / 9007199254740992.0) - 1.0


def terrain_height(x: np.ndarray, z: np.ndarray, i: np.ndarray, j: np.ndarray, A: float, seed: int) -> np.ndarray:
    return (A * np.sin(0.11 * x) * np.cos(0.13 * z) + 0.25 * A * np.sin(0.53 * x + 0.71 * z)
            + 0.05 * A * hash_unit(i, j, seed))


def terrain(nq: int, s: float = 1.2, A: float = 4.0, seed: int = 1, y_offset: float = 0.0
            ) -> Tuple[np.ndarray, np.ndarray]:
    """(nq+1)^2 vertices at (i*s, y, j*s); two triangles per quad (a,c,b),(b,c,d).
    Returns (verts f32 [V*3], indices i32 [T*3])."""
    n1 = nq + 1
    j, i = np.meshgrid(np.arange(n1), np.arange(n1), indexing="ij")
    x = i.astype(np.float64) * s
    z = j.astype(np.float64) * s
    y = terrain_height(x, z, i, j, A, seed) + y_offset
    verts = np.stack([x, y, z], axis=-1).reshape(-1, 3).astype(np.float32)
    jq, iq = np.meshgrid(np.arange(nq), np.arange(nq), indexing="ij")
    a = (jq * n1 + iq).reshape(-1)
    b, c = a + 1, a + n1
    d = c + 1
    tris = np.stack([a, c, b, b, c, d], axis=-1).reshape(-1, 3).astype(np.int32)
    return verts.reshape(-1), tris.reshape(-1)


def boxes_on_terrain(nq: int, s: float, A: float, seed: int, lattice: int, vbase: int
                     ) -> Tuple[np.ndarray, np.ndarray]:
    """Axis-aligned boxes (12 triangles each) on a hashed jittered lattice x lattice grid: footprint
    3-12 units, height 3-15 units, base at terrain height (config 3 'buildings')."""
    extent = nq * s
    gy, gx = np.meshgrid(np.arange(lattice), np.arange(lattice), indexing="ij")
    gx, gy = gx.reshape(-1), gy.reshape(-1)
    pitch = extent / lattice
    h1 = hash_unit(gx, gy, seed + 101) * 0.5 + 0.5
    h2 = hash_unit(gx, gy, seed + 202) * 0.5 + 0.5
    h3 = hash_unit(gx, gy, seed + 303) * 0.5 + 0.5
    h4 = hash_unit(gx, gy, seed + 404) * 0.5 + 0.5
    h5 = hash_unit(gx, gy, seed + 505) * 0.5 + 0.5
    cx = (gx + 0.25 + 0.5 * h1) * pitch
    cz = (gy + 0.25 + 0.5 * h2) * pitch
    fx = 3.0 + 9.0 * h3
    fz = 3.0 + 9.0 * h4
    hh = 3.0 + 12.0 * h5
    x0, x1 = np.clip(cx - fx / 2, 0, extent), np.clip(cx + fx / 2, 0, extent)
    z0, z1 = np.clip(cz - fz / 2, 0, extent), np.clip(cz + fz / 2, 0, extent)
    ii = np.rint(cx / s).astype(np.int64)
    jj = np.rint(cz / s).astype(np.int64)
    y0 = terrain_height(ii * s, jj * s, ii, jj, A, seed) - 0.5
    y1 = y0 + hh
    corners = np.stack([
        np.stack([x0, y0, z0], -1), np.stack([x1, y0, z0], -1), np.stack([x1, y0, z1], -1), np.stack([x0, y0, z1], -1),
        np.stack([x0, y1, z0], -1), np.stack([x1, y1, z0], -1), np.stack([x1, y1, z1], -1), np.stack([x0, y1, z1], -1),
    ], axis=1)  # [B, 8, 3]
    quad = np.array([[0, 1, 2], [0, 2, 3], [4, 6, 5], [4, 7, 6], [0, 5, 1], [0, 4, 5], [1, 6, 2], [1, 5, 6],
                     [2, 7, 3], [2, 6, 7], [3, 4, 0], [3, 7, 4]], dtype=np.int64)
    nb = corners.shape[0]
    tris = (quad[None, :, :] + (np.arange(nb) * 8)[:, None, None] + vbase).astype(np.int32)
    return corners.reshape(-1).astype(np.float32), tris.reshape(-1)


def open_world(nq: int = 2200, s: float = 1.2, A: float = 4.0, seed: int = 1, lattice: int = 164):
    """Config 3: terrain(nq) + lattice^2 boxes x 12 triangles."""
    v, t = terrain(nq, s, A, seed)
    bv, bt = boxes_on_terrain(nq, s, A, seed, lattice, v.size // 3)
    return np.concatenate([v, bv]), np.concatenate([t, bt])


def interior(floors: int = 16, nq: int = 177, s: float = 1.2, A: float = 0.2, seed: int = 1, dy: float = 3.0):
    """Config 4: stacked floors, floor k = terrain(nq, s, A) offset y += dy*k."""
    vs, ts, base = [], [], 0
    for k in range(floors):
        v, t = terrain(nq, s, A, seed + k, y_offset=dy * k)
        vs.append(v)
        ts.append(t + base)
        base += v.size // 3
    return np.concatenate(vs), np.concatenate(ts).astype(np.int32)


def mesh_bounds(verts: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    v = verts.reshape(-1, 3)
    return v.min(axis=0).astype(np.float32), v.max(axis=0).astype(np.float32)


def single_tile_config(verts: np.ndarray, base: RecastConfig | None = None) -> RecastConfig:
    """Config 1: bounds = mesh AABB; grid from RecastConfig::calculate_grid_size (config.rs:85-107),
    evaluated here in f32 exactly as the reference does."""
    cfg = RecastConfig(**(vars(base) if base else {}))
    bmin, bmax = mesh_bounds(verts)
    f = np.float32
    with np.errstate(all="ignore"):
        wf = f(f(bmax[0] - bmin[0]) / f(cfg.cs))
        hf = f(f(bmax[2] - bmin[2]) / f(cfg.cs))
    cfg.bmin, cfg.bmax = tuple(float(x) for x in bmin), tuple(float(x) for x in bmax)
    cfg.width = int(wf) + 1 if f(wf - np.trunc(wf)) == 0 else int(np.ceil(wf))
    cfg.height = int(hf) + 1 if f(hf - np.trunc(hf)) == 0 else int(np.ceil(hf))
    return cfg


def tile_grid(verts: np.ndarray, tile_cells: int, base: RecastConfig | None = None) -> List[RecastConfig]:
    """Tiles of tile_cells x tile_cells cells over the mesh AABB:
    tbmin.x = bmin.x + (tx*n) as f32 * cs, tbmax.x = bmin.x + ((tx+1)*n) as f32 * cs (f32, mul then
    add), same for z; y-range = mesh y-range; last row/column partial; border 0.  Row-major in
    (tz, tx)."""
    whole = single_tile_config(verts, base)
    f = np.float32
    cs = f(whole.cs)
    bmin = [f(x) for x in whole.bmin]
    bmax = [f(x) for x in whole.bmax]
    ntx = (whole.width + tile_cells - 1) // tile_cells
    ntz = (whole.height + tile_cells - 1) // tile_cells
    out = []
    for tz in range(ntz):
        for tx in range(ntx):
            c = RecastConfig(**vars(whole))
            c.width = min(tile_cells, whole.width - tx * tile_cells)
            c.height = min(tile_cells, whole.height - tz * tile_cells)
            x0 = f(bmin[0] + f(f(tx * tile_cells) * cs))
            x1 = f(bmin[0] + f(f((tx + 1) * tile_cells) * cs))
            z0 = f(bmin[2] + f(f(tz * tile_cells) * cs))
            z1 = f(bmin[2] + f(f((tz + 1) * tile_cells) * cs))
            c.bmin = (float(x0), float(bmin[1]), float(z0))
            c.bmax = (float(x1), float(bmax[1]), float(z1))
            out.append(c)
    return out


# ---- the five BASELINE.json configurations -------------------------------------------------------
def config1(seed=1, nq=224):
    v, t = terrain(nq, 1.2, 4.0, seed)
    return v, t, [single_tile_config(v)]


def config2(seed=1, nq=708, tile=64):
    v, t = terrain(nq, 1.2, 4.0, seed)
    return v, t, tile_grid(v, tile)


def config3(seed=1, nq=2200, tile=128, lattice=164):
    v, t = open_world(nq, 1.2, 4.0, seed, lattice)
    return v, t, tile_grid(v, tile)


def config4(seed=1, floors=16, nq=177, tile=64):
    v, t = interior(floors, nq, 1.2, 0.2, seed)
    return v, t, tile_grid(v, tile)
