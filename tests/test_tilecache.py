"""Tile-cache layer path (SURVEY §8(a) row TC, BASELINE config 5): oracle vs the independent Python
restatement on CPU, and the CUDA path vs the oracle on the GPU."""
import numpy as np
import pytest

from oracle import pymodel as pm
from recast_rs_b200 import tilecache as tc

NONE = 0xFFFFFFFF


def _random_case(seed, nlayers=5):
    rng = np.random.default_rng(seed)
    layers, obstacles = [], []
    cs, ch = [(0.3, 0.2), (0.5, 0.25)][seed % 2]
    for t in range(nlayers):
        w, h = int(rng.integers(3, 14)), int(rng.integers(3, 14))
        bmin = (float(rng.uniform(-5, 5)), float(rng.uniform(-1, 1)), float(rng.uniform(-5, 5)))
        bmax = (bmin[0] + w * cs, bmin[1] + 6.0, bmin[2] + h * cs)
        regons = (rng.uniform(size=w * h) > 0.15).astype(np.uint8) * rng.integers(1, 5, w * h).astype(np.uint8)
        areas = rng.choice([0, 1, 63], size=w * h, p=[0.1, 0.3, 0.6]).astype(np.uint8)
        first = len(obstacles)
        for _ in range(int(rng.integers(0, 5))):
            cx, cz = rng.uniform(bmin[0] - 1, bmax[0] + 1), rng.uniform(bmin[2] - 1, bmax[2] + 1)
            cy = bmin[1] + float(rng.uniform(0, 4))
            kind = int(rng.integers(0, 3))
            state = int(rng.choice([tc.EMPTY, tc.PROCESSING, tc.PROCESSED, tc.PROCESSED, tc.REMOVING]))
            if kind == 0:
                obstacles.append(tc.Obstacle.cylinder((cx, cy, cz), float(rng.uniform(0.2, 2.0)), float(rng.uniform(0.1, 3.0)), state))
            elif kind == 1:
                obstacles.append(tc.Obstacle.box((cx - rng.uniform(0.1, 1.5), cy, cz - rng.uniform(0.1, 1.5)),
                                                 (cx + rng.uniform(0.1, 1.5), cy + rng.uniform(0.1, 3), cz + rng.uniform(0.1, 1.5)), state))
            else:
                obstacles.append(tc.Obstacle.oriented_box((cx, cy, cz), (float(rng.uniform(0.2, 2)), float(rng.uniform(0.2, 2)),
                                                                        float(rng.uniform(0.2, 2))), float(rng.uniform(0, 6.28)), state))
        hmin = int(rng.choice([0, 3, 10, 200]))
        layers.append(tc.TileCacheLayer(bmin, bmax, hmin, hmin + 1, w, h, regons, areas, first, len(obstacles) - first))
    return layers, obstacles, cs, ch


def _eq_pm(a, c):
    assert a.cell_index.tolist() == [NONE if i is None else i for (i, _) in c.cells]
    assert a.span_min.tolist() == [s["min"] for s in c.spans] and a.span_max.tolist() == [s["max"] for s in c.spans]
    assert a.span_area.tolist() == [s["area"] for s in c.spans]
    assert a.areas.tolist() == c.areas
    assert a.conn_ref.tolist() == [n for (n, _, _) in c.connections]
    assert a.con.reshape(-1, 4).tolist() == [s["con"] for s in c.spans]


@pytest.mark.parametrize("seed", range(4))
def test_oracle_vs_pymodel_tilecache(oracle, seed):
    layers, obstacles, cs, ch = _random_case(seed)
    got = oracle.tilecache_build(layers, obstacles, cs, ch)
    for l, a in zip(layers, got):
        _eq_pm(a, pm.tilecache_build(l, obstacles, cs, ch))


def _poison_obstacles(rng, obstacles):
    """One parameter of every obstacle becomes NaN, +-inf, +-3e38 or 0 (nothing in tile_cache_builder.rs:175-549 checks)."""
    import dataclasses

    vals = [np.nan, np.inf, -np.inf, 3e38, -3e38, 0.0]
    out = []
    for o in obstacles:
        a, b, rot = list(o.a), list(o.b), list(o.rot_aux)
        val = float(vals[int(rng.integers(0, len(vals)))])
        k = int(rng.integers(0, 3))
        if k == 0:
            a[int(rng.integers(0, 3))] = val
        elif k == 1:
            b[int(rng.integers(0, 3))] = val
        else:
            rot[int(rng.integers(0, 2))] = val
        out.append(dataclasses.replace(o, a=tuple(a), b=tuple(b), rot_aux=tuple(rot)))
    return out


@pytest.mark.parametrize("seed", range(12))
def test_oracle_vs_pymodel_tilecache_with_non_finite_obstacles(oracle, seed):
    layers, obstacles, cs, ch = _random_case(seed)
    obstacles = _poison_obstacles(np.random.default_rng(4000 + seed), obstacles)
    got = oracle.tilecache_build(layers, obstacles, cs, ch)
    with np.errstate(all="ignore"):
        for l, a in zip(layers, got):
            _eq_pm(a, pm.tilecache_build(l, obstacles, cs, ch))


def test_obstacles_clear_span_area_only(oracle):
    layers, obstacles = tc.config5(9, 16)
    got = oracle.tilecache_build(layers, obstacles, 0.3, 0.2)
    for l, a in zip(layers, got):
        _eq_pm(a, pm.tilecache_build(l, obstacles, 0.3, 0.2))
    # CompactSpan.area is cleared under obstacles while chf.areas keeps the layer's area (tile_cache_builder.rs:309)
    assert sum(int(((a.span_area == 0) & (a.areas == 63)).sum()) for a in got) > 0
    assert all((a.areas == 63).all() for a in got)


def test_tilecache_errors(oracle):
    l = tc.TileCacheLayer((0, 0, 0), (1, 1, 1), 10, 11, 2, 2, np.ones(3, np.uint8), np.ones(4, np.uint8))
    assert oracle.tilecache_build([l], [], 0.5, 0.5) == [-1]  # tile_cache_builder.rs:138-143
    l = tc.TileCacheLayer((0, 0, 0), (1, 1, 1), 32767, 32768 & 0xFFFF, 2, 2, np.ones(4, np.uint8), np.ones(4, np.uint8))
    assert oracle.tilecache_build([l], [], 0.5, 0.5) == [-2]  # heightfield.rs:110-115: (32767, -32768)
    l.regons = np.zeros(4, np.uint8)
    assert not isinstance(oracle.tilecache_build([l], [], 0.5, 0.5)[0], int)  # no span => add_span never runs


def _assert_tile(tv, a, what):
    for k in ("span_min", "span_max", "span_area", "cell_index", "cell_count", "areas", "con", "first_conn", "conn_ref", "conn_dir",
              "conn_next", "dist"):
        x, y = getattr(tv, k), getattr(a, k)
        assert np.array_equal(x, y), f"{what}: {k}"
    assert (tv.max_height, tv.walkable_span_count) == (a.max_height, a.walkable_span_count), what


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(4))
def test_gpu_tilecache_random(oracle, seed):
    import recast_rs_b200.builder as rb

    layers, obstacles, cs, ch = _random_case(100 + seed, nlayers=12)
    ctx = rb.Context(0)
    with ctx.build_tilecache_layers(layers, obstacles, cs, ch) as r:
        r.download()
        want = oracle.tilecache_build(layers, obstacles, cs, ch)
        for i, a in enumerate(want):
            _assert_tile(r.tile(i), a, f"layer {i}")
    ctx.close()


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(3))
def test_gpu_tilecache_non_finite_obstacles(oracle, seed):
    import recast_rs_b200.builder as rb

    layers, obstacles, cs, ch = _random_case(200 + seed, nlayers=12)
    obstacles = _poison_obstacles(np.random.default_rng(4200 + seed), obstacles)
    ctx = rb.Context(0)
    with ctx.build_tilecache_layers(layers, obstacles, cs, ch) as r:
        r.download()
        for i, a in enumerate(oracle.tilecache_build(layers, obstacles, cs, ch)):
            _assert_tile(r.tile(i), a, f"layer {i}")
    ctx.close()


@pytest.mark.gpu
def test_gpu_tilecache_config5(oracle):
    """BASELINE config 5 at full size: 256 dirty 48x48 layers with 1-3 obstacles each."""
    import recast_rs_b200 as r
    import recast_rs_b200.builder as rb

    layers, obstacles = tc.config5(256, 48)
    ctx = rb.Context(0)
    with ctx.build_tilecache_layers(layers, obstacles, 0.3, 0.2) as res:
        res.download()
        tot = res.totals()
        assert tot["tiles"] == 256 and tot["spans"] == 256 * 48 * 48
        want = oracle.tilecache_build(layers, obstacles, 0.3, 0.2)
        for i, a in enumerate(want):
            _assert_tile(res.tile(i), a, f"layer {i}")
        assert sum(int((a.span_area == 0).sum()) for a in want) > 0
    bad = tc.TileCacheLayer((0, 0, 0), (1, 1, 1), 10, 11, 2, 2, np.ones(3, np.uint8), np.ones(4, np.uint8))
    with pytest.raises(r.InvalidMesh):
        ctx.build_tilecache_layers([bad], [], 0.5, 0.5)
    bad = tc.TileCacheLayer((0, 0, 0), (1, 1, 1), 32767, 0, 2, 2, np.ones(4, np.uint8), np.ones(4, np.uint8))
    with pytest.raises(r.NavMeshGeneration):
        ctx.build_tilecache_layers([bad], [], 0.5, 0.5)
    ctx.close()
