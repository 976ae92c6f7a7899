"""Region building: CompactHeightfield::build_regions -> watershed::build_regions_watershed
(watershed.rs:364-746).  The reference has no test for watershed.rs ("parity unpinned"): on the CPU the oracle is
cross-checked against the independent Python restatement; on the GPU the CUDA path (one CTA per tile: parallel
expansion, breadth-first floods, atomics for the region tables) is compared span for span with the oracle's
sequential program."""
import numpy as np
import pytest

from oracle import pymodel as pm
from recast_rs_b200 import synth
from recast_rs_b200.config import RecastConfig
from tests.parity import random_soup

PARAMS = [(0, 8, 20), (0, 2, 0), (2, 4, 10), (0, 1, 50), (3, 0, 5)]


def _scene(seed):
    rng = np.random.default_rng(seed)
    w, h = int(rng.integers(8, 30)), int(rng.integers(8, 30))
    cs, ch = 0.5, 0.25
    cfg = RecastConfig(width=w, height=h, cs=cs, ch=ch, bmin=(0, -1, 0), bmax=(w * cs, 8, h * cs), walkable_height=3, walkable_climb=2)
    if seed % 3 == 0:
        v, idx = random_soup(rng, 80, 0.0, max(w, h) * cs, -0.5, 4.0, small=cs * 3)
    elif seed % 3 == 1:
        v, idx = synth.terrain(int(max(w, h) * cs / 1.2) + 2, 1.2, 2.0, seed)
    else:
        v, idx = synth.interior(3, int(max(w, h) * cs / 1.2) + 2, 1.2, 0.3, seed, 1.5)
    return cfg, v, idx


@pytest.mark.parametrize("seed", range(9))
def test_oracle_vs_pymodel_regions(oracle, seed):
    cfg, v, idx = _scene(seed)
    ohf = oracle.Heightfield.from_config(cfg)
    ohf.builder_rasterize(cfg, v, idx)
    phf = pm.Heightfield(cfg.width, cfg.height, cfg.bmin, cfg.bmax, cfg.cs, cfg.ch)
    pm.builder_rasterize(phf, cfg, v, idx)
    for (b, mn, mg) in PARAMS[:4]:
        oc, pc = oracle.CompactHeightfield.build_from_heightfield(ohf), pm.Compact(phf)
        reg, mr = oc.build_regions(b, mn, mg)
        preg, pmr = pm.build_regions_watershed(pc, b, mn, mg)
        assert mr == pmr and reg.tolist() == preg, (b, mn, mg)
        # ids are compact: 1..n without gaps (border ids keep their flag)
        ids = sorted(set(int(r) for r in reg if r and not r & 0x8000))
        assert all(1 <= i <= mr for i in ids)
        assert (reg[oc.export().areas == 0] == 0).all()


ODD_PARAMS = [(0, 0, 0), (1, 1, 1), (-1, 8, 20), (3, -5, 20), (3, 8, -1), (100, 8, 20), (2, 100000, 20), (2, 0, 100000), (5, 1, 0),
              (0, 2147483647, 2147483647), (-100, -100, -100)]


@pytest.mark.parametrize("seed", [0, 4, 8])
def test_oracle_vs_pymodel_regions_with_odd_parameters(oracle, seed):
    """border_size, min_region_area and merge_region_area are plain i32 that nothing validates (watershed.rs:364-375): borders
    wider than the tile, negative sizes, areas nothing can reach.  The two restatements must take them the same way."""
    cfg, v, idx = _scene(seed)
    ohf = oracle.Heightfield.from_config(cfg)
    ohf.builder_rasterize(cfg, v, idx)
    phf = pm.Heightfield(cfg.width, cfg.height, cfg.bmin, cfg.bmax, cfg.cs, cfg.ch)
    pm.builder_rasterize(phf, cfg, v, idx)
    for (b, mn, mg) in ODD_PARAMS:
        oc, pc = oracle.CompactHeightfield.build_from_heightfield(ohf), pm.Compact(phf)
        reg, mr = oc.build_regions(b, mn, mg)
        preg, pmr = pm.build_regions_watershed(pc, b, mn, mg)
        assert mr == pmr and reg.tolist() == preg, (b, mn, mg)


def test_regions_basic_properties(oracle):
    """Regions never span two plateaus that no connection joins; min_region_area only ever removes regions; every
    walkable span that keeps a region has it from a connected neighbourhood (no label without a span count)."""
    cfg = RecastConfig(width=20, height=20, cs=0.5, ch=0.25, bmin=(0, 0, 0), bmax=(10, 4, 10), walkable_height=3, walkable_climb=1)
    quad = lambda x0, z0, x1, z1, y: [(x0, y, z0), (x1, y, z1), (x1, y, z0), (x0, y, z0), (x0, y, z1), (x1, y, z1)]  # noqa: E731
    v2 = np.array(quad(0, 0, 4, 10, 1.0) + quad(6, 0, 10, 10, 3.0), np.float32).reshape(-1)
    hf = oracle.Heightfield.from_config(cfg)
    hf.builder_rasterize(cfg, v2, np.arange(12, dtype=np.int32))
    c = oracle.CompactHeightfield.build_from_heightfield(hf)
    reg, mr = c.build_regions(0, 2, 0)
    e = c.export()
    low, high = set(reg[e.span_min < 8].tolist()) - {0}, set(reg[e.span_min >= 8].tolist()) - {0}
    assert low and high and not (low & high) and max(low | high) < mr
    n = []
    for min_area in (1, 8, 60, 1000):
        r, _ = oracle.CompactHeightfield.build_from_heightfield(hf).build_regions(0, min_area, 0)
        n.append(len(set(r.tolist()) - {0}))
    assert n == sorted(n, reverse=True) and n[-1] == 0 and n[0] > 0


# ---- GPU ---------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(9))
def test_gpu_regions_single_tiles(oracle, seed):
    import recast_rs_b200.builder as rb

    cfg, v, idx = _scene(100 + seed)
    ohf = oracle.Heightfield.from_config(cfg)
    ohf.builder_rasterize(cfg, v, idx)
    e = ohf.export()
    with rb.Context(0) as ctx:
        hf = rb.Heightfield(cfg.width, cfg.height, cfg.bmin, cfg.bmax, cfg.cs, cfg.ch, e.cell_offset, e.span_min, e.span_max, e.span_area, ctx=ctx)
        for (b, mn, mg) in PARAMS:
            want, wmr = oracle.CompactHeightfield.build_from_heightfield(ohf).build_regions(b, mn, mg)
            chf = rb.CompactHeightfield.build_from_heightfield(hf)
            chf.build_regions(b, mn, mg)
            assert chf.max_regions == wmr, (b, mn, mg)
            np.testing.assert_array_equal(chf.reg, want, err_msg=str((b, mn, mg)))


@pytest.mark.gpu
def test_gpu_regions_with_odd_parameters(oracle):
    import recast_rs_b200.builder as rb

    cfg, v, idx = _scene(104)
    ohf = oracle.Heightfield.from_config(cfg)
    ohf.builder_rasterize(cfg, v, idx)
    e = ohf.export()
    with rb.Context(0) as ctx:
        hf = rb.Heightfield(cfg.width, cfg.height, cfg.bmin, cfg.bmax, cfg.cs, cfg.ch, e.cell_offset, e.span_min, e.span_max, e.span_area, ctx=ctx)
        for (b, mn, mg) in ODD_PARAMS:
            want, wmr = oracle.CompactHeightfield.build_from_heightfield(ohf).build_regions(b, mn, mg)
            chf = rb.CompactHeightfield.build_from_heightfield(hf)
            chf.build_regions(b, mn, mg)
            assert chf.max_regions == wmr, (b, mn, mg)
            np.testing.assert_array_equal(chf.reg, want, err_msg=str((b, mn, mg)))


@pytest.mark.gpu
@pytest.mark.parametrize("force_global", [False, True])
def test_gpu_regions_tiled_world_both_back_ends(oracle, force_global):
    """RecastBuilder::build_mesh's voxel half + regions over a tiled terrain, tile-resident and global back ends."""
    import recast_rs_b200.builder as rb
    from tests.test_gpu_fullsize import _context

    v, t = synth.terrain(60, 1.2, 4.0, 5)
    cfgs = synth.tile_grid(v, 48, RecastConfig(cs=0.3, ch=0.2, min_region_area=8, merge_region_area=20))
    ctx = _context(force_global)
    ctx.upload_mesh(v, t)
    with ctx.build_tiles(cfgs, rb.STAGE_ALL_BUILD | rb.STAGE_REGIONS) as r:
        r.download()
        for i, c in enumerate(cfgs):
            ohf = oracle.Heightfield.from_config(c)
            ohf.builder_rasterize(c, v, t)
            oc = oracle.CompactHeightfield.build_from_heightfield(ohf)
            want, wmr = oc.build_regions(c.border_size, c.min_region_area, c.merge_region_area)
            tv = r.tile(i)
            assert tv.max_regions == wmr, i
            np.testing.assert_array_equal(tv.reg, want, err_msg=f"tile {i}")
            np.testing.assert_array_equal(tv.dist, oc.export().dist)
    ctx.close()


@pytest.mark.gpu
def test_gpu_regions_config2_invariants_and_sample(oracle):
    """Full size (2025 tiles): both back ends agree on every span; a sample of tiles against the oracle."""
    import recast_rs_b200.builder as rb
    from tests.test_gpu_fullsize import _context

    v, t, cfgs = synth.config2()
    out = []
    for fg in (False, True):
        ctx = _context(fg)
        ctx.upload_mesh(v, t)
        with ctx.build_tiles(cfgs, rb.STAGE_ALL_BUILD | rb.STAGE_REGIONS) as r:
            r.download()
            out.append([(r.tile(i).reg.copy(), r.tile(i).max_regions, r.tile(i).areas.copy()) for i in range(len(cfgs))])
        ctx.close()
    for i, (a, b) in enumerate(zip(*out)):
        assert a[1] == b[1] and np.array_equal(a[0], b[0]), i
        assert (a[0][a[2] == 0] == 0).all()
        ids = np.unique(a[0][a[0] != 0])
        assert ids.size == 0 or (ids[0] == 1 and ids[-1] == ids.size)  # compact ids
    for i in (0, 517, 1033, 2024):
        ohf = oracle.Heightfield.from_config(cfgs[i])
        ohf.builder_rasterize(cfgs[i], v, t)
        want, wmr = oracle.CompactHeightfield.build_from_heightfield(ohf).build_regions(cfgs[i].border_size, cfgs[i].min_region_area,
                                                                                         cfgs[i].merge_region_area)
        assert out[0][i][1] == wmr
        np.testing.assert_array_equal(out[0][i][0], want)


@pytest.mark.gpu
def test_gpu_regions_tall_tile_global_memory_path(oracle):
    """A tile with more than 65535 spans (18 floors over 64x64 cells): 32-bit neighbour table, labels in global memory."""
    import recast_rs_b200.builder as rb

    v, t = synth.interior(18, 17, 1.2, 0.2, 3, 3.0)
    cfg = synth.single_tile_config(v, RecastConfig(cs=0.3, ch=0.2, min_region_area=8, merge_region_area=20))
    ohf = oracle.Heightfield.from_config(cfg)
    ohf.builder_rasterize(cfg, v, t)
    oc = oracle.CompactHeightfield.build_from_heightfield(ohf)
    want, wmr = oc.build_regions(0, 8, 20)
    assert want.size > 65535
    with rb.Context(0) as ctx:
        ctx.upload_mesh(v, t)
        with ctx.build_tiles([cfg], rb.STAGE_ALL_BUILD | rb.STAGE_REGIONS) as r:
            tv = r.download().tile(0)
    assert tv.max_regions == wmr
    np.testing.assert_array_equal(tv.reg, want)


@pytest.mark.gpu
def test_gpu_regions_mixed_batch_one_huge_tile_among_small_ones(oracle):
    """One tile past 65535 spans in a batch of small ones: the width of the neighbour table is a property of the batch
    (16- and 32-bit planes of different tiles would overlap in the shared scratch)."""
    import recast_rs_b200.builder as rb

    v, t = synth.interior(18, 17, 1.2, 0.2, 3, 3.0)
    base = RecastConfig(cs=0.3, ch=0.2, min_region_area=8, merge_region_area=20)
    big = synth.single_tile_config(v, base)
    small = synth.tile_grid(v, 24, base)[:5]
    cfgs = small[:2] + [big] + small[2:]
    with rb.Context(0) as ctx:
        ctx.upload_mesh(v, t)
        with ctx.build_tiles(cfgs, rb.STAGE_ALL_BUILD | rb.STAGE_REGIONS) as r:
            r.download()
            sizes = [r.tile(i).reg.size for i in range(len(cfgs))]
            assert max(sizes) > 65535 and min(sizes) < 65535
            for i, c in enumerate(cfgs):
                ohf = oracle.Heightfield.from_config(c)
                ohf.builder_rasterize(c, v, t)
                want, wmr = oracle.CompactHeightfield.build_from_heightfield(ohf).build_regions(c.border_size, c.min_region_area,
                                                                                                 c.merge_region_area)
                assert r.tile(i).max_regions == wmr, i
                np.testing.assert_array_equal(r.tile(i).reg, want, err_msg=f"tile {i}")
