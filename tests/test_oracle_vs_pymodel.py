"""Differential test: oracle.cpp (C++) against oracle/pymodel.py (independent plain-Python
restatement with the reference's own data structures) on seeded random inputs.  CPU only."""
import numpy as np
import pytest

from oracle import pymodel as pm
from recast_rs_b200.config import RecastConfig


def _random_soup(rng, ntri, lo, hi, ylo, yhi, small=0.5):
    """Triangles of mixed size, some degenerate, some outside the bounds."""
    verts = []
    for _ in range(ntri):
        c = rng.uniform(lo - 1.0, hi + 1.0, size=2)
        kind = rng.integers(0, 6)
        ext = {0: small, 1: 2.0, 2: 6.0, 3: small, 4: 1.0, 5: 12.0}[int(kind)]
        tri = []
        for _k in range(3):
            tri.append((c[0] + rng.uniform(-ext, ext), rng.uniform(ylo, yhi), c[1] + rng.uniform(-ext, ext)))
        if kind == 3:  # degenerate: repeated vertex
            tri[2] = tri[1]
        if kind == 4:  # axis aligned vertical wall
            tri[1] = (tri[0][0], tri[1][1], tri[1][2])
            tri[2] = (tri[0][0], tri[2][1], tri[2][2])
        verts += tri
    v = np.array(verts, dtype=np.float32).reshape(-1)
    idx = np.arange(ntri * 3, dtype=np.int32)
    return v, idx


def _cfg(w, h, bmin, cs, ch, ymax, **kw):
    c = RecastConfig(width=w, height=h, cs=cs, ch=ch, bmin=bmin,
                     bmax=(bmin[0] + w * cs, ymax, bmin[2] + h * cs), **kw)
    return c


def _compare_hf(orc_hf, pm_hf):
    a = orc_hf.export()
    off, mn, mx, ar = pm_hf.export()
    np.testing.assert_array_equal(a.cell_offset, off)
    np.testing.assert_array_equal(a.span_min, mn)
    np.testing.assert_array_equal(a.span_max, mx)
    np.testing.assert_array_equal(a.span_area, ar)


def _compare_chf(a, c: pm.Compact):
    NONE = 0xFFFFFFFF
    assert a.cell_index.tolist() == [NONE if i is None else i for (i, _) in c.cells]
    assert a.cell_count.tolist() == [n for (_, n) in c.cells]
    assert a.span_min.tolist() == [s["min"] for s in c.spans]
    assert a.span_max.tolist() == [s["max"] for s in c.spans]
    assert a.span_area.tolist() == [s["area"] for s in c.spans]
    assert a.areas.tolist() == c.areas
    assert a.con.reshape(-1, 4).tolist() == [s["con"] for s in c.spans]
    assert a.first_conn.tolist() == [NONE if s["first"] is None else s["first"] for s in c.spans]
    assert a.conn_ref.tolist() == [n for (n, _, _) in c.connections]
    assert a.conn_dir.tolist() == [d for (_, d, _) in c.connections]
    assert a.conn_next.tolist() == [NONE if x is None else x for (_, _, x) in c.connections]
    assert a.dist.tolist() == c.dist
    assert a.max_height == c.max_height
    assert a.max_distance == c.max_distance
    assert a.walkable_span_count == c.walkable_span_count


@pytest.mark.parametrize("seed", range(6))
def test_builder_pipeline_random_soup(oracle, seed):
    rng = np.random.default_rng(1000 + seed)
    w, h = int(rng.integers(5, 12)), int(rng.integers(5, 12))
    cs, ch = [(0.3, 0.2), (0.5, 0.25), (1.0, 1.0)][seed % 3]
    bmin = (float(rng.uniform(-3, 3)), float(rng.uniform(-1, 0.5)), float(rng.uniform(-3, 3)))
    cfg = _cfg(w, h, bmin, cs, ch, bmin[1] + 6.0, walkable_height=int(rng.integers(1, 5)),
               walkable_climb=int(rng.integers(0, 3)), walkable_slope_angle=float(rng.choice([30.0, 45.0, 60.0])))
    v, idx = _random_soup(rng, 60, min(bmin[0], bmin[2]), max(bmin[0] + w * cs, bmin[2] + h * cs), bmin[1] - 1.0,
                          bmin[1] + 7.0, small=cs)
    for run_filters in (False, True):
        ohf = oracle.Heightfield.from_config(cfg)
        ohf.builder_rasterize(cfg, v, idx, run_filters=run_filters)
        assert not ohf.poly_overflow
        phf = pm.Heightfield(w, h, cfg.bmin, cfg.bmax, cs, ch)
        pm.builder_rasterize(phf, cfg, v, idx, run_filters=run_filters)
        _compare_hf(ohf, phf)
    ochf = oracle.CompactHeightfield.build_from_heightfield(ohf)
    pchf = pm.Compact(phf)
    _compare_chf(ochf.export(), pchf)
    ochf.build_distance_field()
    pchf.build_distance_field()
    _compare_chf(ochf.export(), pchf)
    ochf.erode_walkable_area(1 + seed % 2)
    pchf.erode_walkable_area(1 + seed % 2)
    _compare_chf(ochf.export(), pchf)
    ochf.build_distance_field()  # distance field after erosion (areas changed, connections kept)
    pchf.build_distance_field()
    _compare_chf(ochf.export(), pchf)


@pytest.mark.parametrize("seed", range(8))
def test_non_finite_vertices(oracle, seed):
    """NaN, infinities and 3e38 in the vertex array: nothing in the reference rejects them, so its float rules decide -
    f32::min / max return the other operand for a NaN (Python's min() does not: the model has its own), `as i32` / `as i16`
    saturate with NaN -> 0, comparisons with NaN are false.  The two restatements were written independently and must
    agree here as well (they did not until the model stopped using Python's min / max)."""
    from tests.parity import poison_vertices

    rng = np.random.default_rng(5000 + seed)
    w, h = int(rng.integers(5, 12)), int(rng.integers(5, 12))
    cs, ch = [(0.3, 0.2), (0.5, 0.25), (1.0, 1.0)][seed % 3]
    bmin = (float(rng.uniform(-3, 3)), float(rng.uniform(-1, 0.5)), float(rng.uniform(-3, 3)))
    cfg = _cfg(w, h, bmin, cs, ch, bmin[1] + 6.0, walkable_height=int(rng.integers(1, 5)), walkable_climb=int(rng.integers(0, 3)))
    v, idx = _random_soup(rng, 60, min(bmin[0], bmin[2]), max(bmin[0] + w * cs, bmin[2] + h * cs), bmin[1] - 1.0, bmin[1] + 7.0, small=cs)
    v = poison_vertices(rng, v, 0.1)
    assert not np.isfinite(v).all()
    for run_filters in (False, True):
        ohf = oracle.Heightfield.from_config(cfg)
        ohf.builder_rasterize(cfg, v, idx, run_filters=run_filters)
        phf = pm.Heightfield(w, h, cfg.bmin, cfg.bmax, cs, ch)
        pm.builder_rasterize(phf, cfg, v, idx, run_filters=run_filters)
        _compare_hf(ohf, phf)
    ochf, pchf = oracle.CompactHeightfield.build_from_heightfield(ohf), pm.Compact(phf)
    ochf.build_distance_field(), pchf.build_distance_field()
    _compare_chf(ochf.export(), pchf)


@pytest.mark.parametrize("seed", range(24))
def test_odd_configurations(oracle, seed):
    """Cell sizes from a centimetre to a hundred metres, cell heights from a millimetre to fifty metres, origins ten million
    units away (where f32 has no fraction left), height ranges that saturate i16, slope limits of 0, 0.001, 89.99 and 90
    degrees, tiles of a single cell: both restatements through rasterization, filters, compaction, distance field, erosion."""
    rng = np.random.default_rng(20000 + seed)
    w, h = int(rng.integers(1, 10)), int(rng.integers(1, 10))
    cs = float(rng.choice([0.01, 0.3, 0.5, 1.0, 7.5, 100.0]))
    ch = float(rng.choice([0.001, 0.2, 0.25, 1.0, 50.0]))
    org = float(rng.choice([0.0, -3.0, 1e4, -1e6, 1e7]))
    bmin = (org + float(rng.uniform(-1, 1)) * cs, float(rng.choice([0.0, -5.0, 1e4, -1e5])) * ch / 0.25, org + float(rng.uniform(-1, 1)) * cs)
    slope = float(rng.choice([0.0, 0.001, 30.0, 45.0, 89.99, 90.0]))
    cfg = _cfg(w, h, bmin, cs, ch, bmin[1] + float(rng.choice([6.0, 1e5])) * ch, walkable_height=int(rng.integers(0, 6)),
               walkable_climb=int(rng.integers(0, 4)), walkable_slope_angle=slope)
    lo = min(bmin[0], bmin[2])
    v, idx = _random_soup(rng, 40, lo, lo + max(w, h) * cs, bmin[1] - 2 * ch, bmin[1] + float(rng.choice([30.0, 4e4, 2e5])) * ch, small=cs)
    ohf = oracle.Heightfield.from_config(cfg)
    ohf.builder_rasterize(cfg, v, idx, run_filters=True)
    phf = pm.Heightfield(w, h, cfg.bmin, cfg.bmax, cs, ch)
    with np.errstate(all="ignore"):
        pm.builder_rasterize(phf, cfg, v, idx, run_filters=True)
    _compare_hf(ohf, phf)
    o, p = oracle.CompactHeightfield.build_from_heightfield(ohf), pm.Compact(phf)
    o.build_distance_field(), p.build_distance_field()
    _compare_chf(o.export(), p)
    o.erode_walkable_area(1), p.erode_walkable_area(1)
    _compare_chf(o.export(), p)


@pytest.mark.parametrize("seed", range(4))
def test_add_span_rules_random_columns(oracle, seed):
    """Both merge rules (rasterization.rs:51 and heightfield.rs:144) on random columns, incl.
    inverted spans (smax < smin) which the rasterizer can emit (no check at rasterization.rs:346-350)."""
    rng = np.random.default_rng(2000 + seed)
    w = h = 4
    o1 = oracle.Heightfield(w, h, (0, 0, 0), (4, 50, 4), 1.0, 1.0)
    p1 = pm.Heightfield(w, h, (0, 0, 0), (4, 50, 4), 1.0, 1.0)
    o2 = oracle.Heightfield(w, h, (0, 0, 0), (4, 50, 4), 1.0, 1.0)
    p2 = pm.Heightfield(w, h, (0, 0, 0), (4, 50, 4), 1.0, 1.0)
    for _ in range(400):
        x, z = int(rng.integers(-1, 5)), int(rng.integers(-1, 5))
        a = int(rng.integers(0, 40))
        b = a + int(rng.integers(-2, 8))
        area = int(rng.integers(0, 4))
        thr = int(rng.integers(0, 3))
        o1.add_span_raster(x, z, a, b, area, thr)
        pm.add_span(p1, x, z, a, b, area, thr)
        if 0 <= x < w and 0 <= z < h and a <= b:
            o2.add_span(x, z, a, b, area)
            p2.add_span(x, z, a, b, area)
    _compare_hf(o1, p1)
    _compare_hf(o2, p2)
    # the whole downstream chain on the AS2-built (possibly unsorted / overlapping) columns
    for (o, p) in ((o1, p1), (o2, p2)):
        o.filter_low_hanging_walkable_obstacles(2), pm.filter_low_hanging(p, 2)
        o.filter_ledge_spans(3, 2), pm.filter_ledge(p, 3, 2)
        o.filter_walkable_low_height_spans(3), pm.filter_low_height(p, 3)
        _compare_hf(o, p)
        oc, pc = oracle.CompactHeightfield.build_from_heightfield(o), pm.Compact(p)
        oc.build_distance_field(), pc.build_distance_field()
        _compare_chf(oc.export(), pc)


def test_span_area_is_order_dependent_geometry_is_not(oracle):
    """SURVEY §7 'hard parts': permuting triangles may change span area, never span geometry."""
    rng = np.random.default_rng(7)
    cfg = _cfg(8, 8, (0.0, 0.0, 0.0), 0.5, 0.25, 6.0)
    v, idx = _random_soup(rng, 80, 0.0, 4.0, 0.0, 5.0)
    base = oracle.Heightfield.from_config(cfg)
    base.builder_rasterize(cfg, v, idx, run_filters=False)
    b = base.export()
    changed = 0
    for k in range(8):
        perm = np.random.default_rng(k).permutation(idx.size // 3)
        pidx = idx.reshape(-1, 3)[perm].reshape(-1)
        hf = oracle.Heightfield.from_config(cfg)
        hf.builder_rasterize(cfg, v, pidx, run_filters=False)
        e = hf.export()
        np.testing.assert_array_equal(e.cell_offset, b.cell_offset)
        np.testing.assert_array_equal(e.span_min, b.span_min)
        np.testing.assert_array_equal(e.span_max, b.span_max)
        changed += int((e.span_area != b.span_area).sum())
    assert changed > 0
