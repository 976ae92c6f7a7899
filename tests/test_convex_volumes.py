"""Convex volumes (second half): ConvexVolume / ConvexVolumeSet (crates/recast/src/convex_volume.rs)
and their use by RecastBuilder::build_heightfield_with_volumes / build_mesh_with_volumes (lib.rs:293-361).
Pinned by the reference's own unit tests (convex_volume.rs:434-558), restated below against the oracle and the host
mirror; then oracle vs the Python restatement, and the CUDA path vs the oracle."""
import numpy as np
import pytest

from oracle import pymodel as pm
from recast_rs_b200 import errors
from recast_rs_b200.config import RecastConfig
from recast_rs_b200.convex_volume import ConvexVolume, ConvexVolumeSet
from tests.parity import random_soup


# ---- the reference's tests --------------------------------------------------------------------------------------
def test_ref_convex_volume_creation(oracle):  # convex_volume.rs:434-451
    sq = [(0, 0, 0), (10, 0, 0), (10, 0, 10), (0, 0, 10)]
    assert oracle.convex_volume_validate(sq, 0.0, 5.0) == 0
    ConvexVolume.new(sq, 0.0, 5.0, 1)
    assert oracle.convex_volume_validate(sq[:2], 0.0, 5.0) == -1
    with pytest.raises(errors.InvalidMesh):
        ConvexVolume.new(sq[:2], 0.0, 5.0, 1)
    with pytest.raises(errors.InvalidMesh):
        ConvexVolume.new(sq, 6.0, 5.0, 1)
    with pytest.raises(errors.InvalidMesh):
        ConvexVolume.new([(i, 0, i * i) for i in range(33)], 0.0, 5.0, 1)


def test_ref_box_volume(oracle):  # :453-470
    v = ConvexVolume.from_box((5.0, 2.5, 5.0), (5.0, 2.5, 5.0), 2)
    inside = [(5.0, 2.5, 5.0), (1.0, 1.0, 1.0), (9.0, 4.0, 9.0)]
    outside = [(-1.0, 2.5, 5.0), (5.0, 6.0, 5.0), (5.0, -1.0, 5.0)]
    for p in inside:
        assert v.contains_point(p) and oracle.convex_contains_point(v.vertices, v.min_height, v.max_height, p)
        assert pm.convex_contains_point(v.vertices.tolist(), v.min_height, v.max_height, p)
    for p in outside:
        assert not v.contains_point(p) and not oracle.convex_contains_point(v.vertices, v.min_height, v.max_height, p)
        assert not pm.convex_contains_point(v.vertices.tolist(), v.min_height, v.max_height, p)


def test_ref_cylinder_volume(oracle):  # :472-493
    v = ConvexVolume.from_cylinder((0.0, 5.0, 0.0), 10.0, 10.0, 8, 3)
    checks = {(0.0, 5.0, 0.0): True, (9.0, 5.0, 0.0): True, (15.0, 5.0, 0.0): False, (0.0, 0.1, 0.0): True, (0.0, 9.9, 0.0): True,
              (0.0, -0.1, 0.0): False, (0.0, 10.1, 0.0): False}
    for p, want in checks.items():
        assert v.contains_point(p) == want, p
        assert oracle.convex_contains_point(v.vertices, v.min_height, v.max_height, p) == want, p


def test_ref_convex_check(oracle):  # :495-515
    sq = [(0, 0, 0), (2, 0, 0), (2, 0, 2), (0, 0, 2)]
    ell = [(0, 0, 0), (2, 0, 0), (2, 0, 1), (1, 0, 1), (1, 0, 2), (0, 0, 2)]
    assert oracle.convex_is_convex(sq) and ConvexVolume.is_convex(sq)
    assert not oracle.convex_is_convex(ell) and not ConvexVolume.is_convex(ell)


def test_ref_volume_set():  # :517-541
    s = ConvexVolumeSet()
    s.add_volume(ConvexVolume.from_box((5.0, 0.0, 5.0), (5.0, 5.0, 5.0), 1))
    s.add_volume(ConvexVolume.from_box((7.0, 0.0, 7.0), (3.0, 5.0, 3.0), 2))
    assert s.get_area_at_point((1.0, 0.0, 1.0), 0) == 1
    assert s.get_area_at_point((7.0, 0.0, 7.0), 0) == 2
    assert s.get_area_at_point((20.0, 0.0, 20.0), 0) == 0
    assert s.find_volumes_at_point((7.0, 0.0, 7.0)) == [0, 1]


# ---- oracle vs Python restatement, CUDA vs oracle ------------------------------------------------------------------
def _scene(seed):
    rng = np.random.default_rng(seed)
    w, h, cs, ch = int(rng.integers(10, 30)), int(rng.integers(10, 30)), 0.5, 0.25
    cfg = RecastConfig(width=w, height=h, cs=cs, ch=ch, bmin=(1.0, -1.0, -2.0), bmax=(1.0 + w * cs, 7.0, -2.0 + h * cs), walkable_height=3,
                       walkable_climb=2)
    v, idx = random_soup(rng, 120, 0.0, max(w, h) * cs, -0.5, 4.0, small=cs * 3)
    vs = ConvexVolumeSet()
    for _ in range(int(rng.integers(1, 7))):
        cx, cz, cy = rng.uniform(1, 1 + w * cs), rng.uniform(-2, -2 + h * cs), rng.uniform(-1, 4)
        kind = int(rng.integers(0, 3))
        area = int(rng.choice([0, 1, 5, 63, 200]))
        if kind == 0:
            vs.add_volume(ConvexVolume.from_box((cx, cy, cz), (rng.uniform(0.3, 4), rng.uniform(0.2, 3), rng.uniform(0.3, 4)), area))
        elif kind == 1:
            vs.add_volume(ConvexVolume.from_cylinder((cx, cy, cz), float(rng.uniform(0.5, 4)), float(rng.uniform(0.5, 5)), int(rng.integers(3, 33)), area))
        else:
            n = int(rng.integers(3, 9))
            ang = np.sort(rng.uniform(0, 2 * np.pi, n))
            rad = rng.uniform(0.8, 4)
            pts = [(cx + rad * np.cos(a), cy, cz + rad * np.sin(a)) for a in ang]
            try:
                vs.add_volume(ConvexVolume.new(pts, cy - rng.uniform(0, 2), cy + rng.uniform(0, 3), area))
            except errors.InvalidMesh:
                pass
    return cfg, v, idx, vs


@pytest.mark.parametrize("seed", range(6))
def test_oracle_vs_pymodel_apply_volumes(oracle, seed):
    cfg, v, idx, vs = _scene(seed)
    ohf = oracle.Heightfield.from_config(cfg)
    ohf.builder_rasterize(cfg, v, idx)
    before = ohf.export().span_area.copy()
    ohf.apply_convex_volumes(vs)
    phf = pm.Heightfield(cfg.width, cfg.height, cfg.bmin, cfg.bmax, cfg.cs, cfg.ch)
    pm.builder_rasterize(phf, cfg, v, idx)
    pm.apply_convex_volumes(phf, [(vol.vertices.tolist(), vol.min_height, vol.max_height, vol.area_type) for vol in vs.volumes])
    a, b = ohf.export(), phf.export()
    assert np.array_equal(a.span_area, b[3]) and np.array_equal(a.span_min, b[1])
    if seed < 3:
        assert (a.span_area != before).any()


@pytest.mark.parametrize("seed", range(12))
def test_oracle_vs_pymodel_volumes_with_non_finite_values(oracle, seed):
    """Volumes that carry a NaN, an infinity or 3e38 in a vertex or a height (built past ConvexVolume::new, whose `min > max`
    test lets NaN through anyway): contains_point's comparisons (convex_volume.rs:130-164) decide, the same way in both
    restatements."""
    cfg, v, idx, vs = _scene(seed)
    rng = np.random.default_rng(7000 + seed)
    vals = [np.nan, np.inf, -np.inf, 3e38, -3e38, 0.0]
    ps = ConvexVolumeSet()
    for vol in vs.volumes:
        verts = np.array(vol.vertices, dtype=np.float32).reshape(-1, 3).copy()
        mn, mx = vol.min_height, vol.max_height
        k, val = int(rng.integers(0, 3)), float(vals[int(rng.integers(0, len(vals)))])
        if k == 0:
            verts[int(rng.integers(0, len(verts))), int(rng.integers(0, 3))] = val
        elif k == 1:
            mn = val
        else:
            mx = val
        ps.add_volume(ConvexVolume(verts, mn, mx, vol.area_type))
    ohf = oracle.Heightfield.from_config(cfg)
    ohf.builder_rasterize(cfg, v, idx)
    phf = pm.Heightfield(cfg.width, cfg.height, cfg.bmin, cfg.bmax, cfg.cs, cfg.ch)
    with np.errstate(all="ignore"):
        pm.builder_rasterize(phf, cfg, v, idx)
        ohf.apply_convex_volumes(ps)
        pm.apply_convex_volumes(phf, [(np.asarray(x.vertices).tolist(), x.min_height, x.max_height, x.area_type) for x in ps.volumes])
    assert np.array_equal(ohf.export().span_area, phf.export()[3])


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(6))
def test_gpu_volumes(oracle, seed):
    import recast_rs_b200.builder as rb
    from tests.parity import assert_tile_equal

    cfg, v, idx, vs = _scene(50 + seed)
    ohf = oracle.Heightfield.from_config(cfg)
    ohf.builder_rasterize(cfg, v, idx)
    plain = ohf.export()
    ohf.apply_convex_volumes(vs)
    want = ohf.export()
    oc = oracle.CompactHeightfield.build_from_heightfield(ohf)
    reg, mr = oc.build_regions(cfg.border_size, cfg.min_region_area, cfg.merge_region_area)
    with rb.Context(0) as ctx:
        # RecastBuilder::build_heightfield_with_volumes (lib.rs:346-361)
        hf = rb.RecastBuilder(cfg, ctx).build_heightfield_with_volumes(v, idx, vs)
        np.testing.assert_array_equal(hf.span_area, want.span_area)
        np.testing.assert_array_equal(hf.span_min, want.span_min)
        # ConvexVolumeSet::apply_to_heightfield on an existing heightfield
        hf2 = rb.Heightfield(cfg.width, cfg.height, cfg.bmin, cfg.bmax, cfg.cs, cfg.ch, plain.cell_offset, plain.span_min, plain.span_max,
                             plain.span_area, ctx=ctx)
        vs.apply_to_heightfield(hf2)
        np.testing.assert_array_equal(hf2.span_area, want.span_area)
        # build_mesh_with_volumes' voxel half (lib.rs:293-316): volumes, compaction, regions
        ctx.upload_mesh(v, idx)
        with ctx.build_tiles_with_volumes([cfg], vs, rb.STAGE_ALL_BUILD | rb.STAGE_REGIONS) as r:
            tv = r.download().tile(0)
        assert_tile_equal(tv, want, oc.export(), what="with volumes")
        np.testing.assert_array_equal(tv.reg, reg)
        assert tv.max_regions == mr
        # what ConvexVolume::new refuses
        bad = ConvexVolumeSet([ConvexVolume(np.array([(0, 0, 0), (2, 0, 0), (2, 0, 1), (1, 0, 1), (1, 0, 2), (0, 0, 2)], np.float32), 0.0, 1.0, 1)])
        with pytest.raises(errors.InvalidMesh):
            ctx.build_tiles_with_volumes([cfg], bad, rb.STAGE_RASTER | rb.STAGE_FILTER)
