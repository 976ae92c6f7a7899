"""GPU parity tests proper: the CUDA path (through the C ABI) against the CPU oracle, bit-exact.
Run on the B200 box with `pytest -m gpu`."""
import numpy as np
import pytest

from tests.parity import assert_tile_equal, random_soup

pytestmark = pytest.mark.gpu

DIR_E = 3


@pytest.fixture(scope="module")
def rb():
    import recast_rs_b200.builder as b

    return b


@pytest.fixture(scope="module")
def ctx(rb):
    c = rb.Context(0)
    yield c
    c.close()


def _oracle_tiles(oracle, verts, idx, cfgs, stage_mask, erode_radius=0):
    return [oracle.build_tile(c, verts, idx, stage_mask, erode_radius) for c in cfgs]


def _check_batch(oracle, rb, ctx, verts, idx, cfgs, stage_mask, erode_radius=0):
    ctx.upload_mesh(verts, idx)
    with ctx.build_tiles(cfgs, stage_mask, erode_radius) as r:
        r.download()
        tot = r.totals()
        ref = _oracle_tiles(oracle, verts, idx, cfgs, stage_mask, erode_radius)
        S = K = 0
        for i, (hfa, chfa) in enumerate(ref):
            assert_tile_equal(r.tile(i), hfa, chfa, what=f"tile {i}")
            S += hfa.span_min.size
            K += 0 if chfa is None else chfa.conn_ref.size
        assert tot["spans"] == S and tot["connections"] == K and tot["tiles"] == len(cfgs)
    return tot


# ---- the reference's own unit tests, through the product API -------------------------------------------
def _fixture_hf(oracle, rb, ctx, w, h, bmax, spans):
    """Build the fixture with the reference's Heightfield::add_span rule (oracle restatement of
    heightfield.rs:102) and hand its columns to the product as a Heightfield."""
    o = oracle.Heightfield(w, h, (0, 0, 0), bmax, 1.0, 1.0)
    for s in spans:
        o.add_span(*s)
    e = o.export()
    return o, rb.Heightfield(w, h, (0, 0, 0), bmax, 1.0, 1.0, e.cell_offset, e.span_min, e.span_max, e.span_area, ctx=ctx)


def test_ref_filter_low_hanging_walkable_obstacles(oracle, rb, ctx):  # heightfield.rs:1417-1452
    _, hf = _fixture_hf(oracle, rb, ctx, 5, 5, (5, 10, 5), [(2, 2, 0, 5, 1), (2, 2, 6, 8, 0), (2, 2, 9, 15, 1)])
    hf.filter_low_hanging_walkable_obstacles(3)
    assert hf.column(2, 2) == [(0, 5, 1), (6, 8, 1), (9, 15, 1)]


def test_ref_compact_heightfield_creation(oracle, rb, ctx):  # compact_heightfield.rs:1028-1059 + KAT-1
    o, hf = _fixture_hf(oracle, rb, ctx, 3, 3, (3, 3, 3), [(0, 0, 0, 2, 1), (0, 0, 3, 5, 1), (1, 0, 0, 2, 1), (0, 1, 0, 2, 1)])
    chf = rb.CompactHeightfield.build_from_heightfield(hf)
    assert (chf.width, chf.height) == (3, 3)
    assert chf.cell_index.size == 9 and chf.span_min.size == 4 and chf.conn_ref.size > 0
    conns = list(zip(chf.conn_ref.tolist(), chf.conn_dir.tolist(), chf.conn_next.tolist()))
    N = 0xFFFFFFFF
    assert conns == [(3, 5, N), (2, 3, 0), (0, 7, N), (3, 6, 2), (2, 2, N), (0, 1, 4)]
    assert chf.con.reshape(-1, 4).tolist() == [[63, 0, 0, 63], [63, 63, 63, 63], [0, 63, 63, 63], [63, 63, 63, 0]]
    assert chf.first_conn.tolist() == [1, N, 3, 5]


def test_ref_get_neighbor(oracle, rb, ctx):  # compact_heightfield.rs:1093-1117
    _, hf = _fixture_hf(oracle, rb, ctx, 2, 1, (2, 1, 1), [(0, 0, 0, 1, 1), (1, 0, 0, 1, 1)])
    chf = rb.CompactHeightfield.build_from_heightfield(hf)
    assert chf.get_neighbor(0, DIR_E) == 1


def test_ref_add_span_merge_via_rasterize(oracle, rb, ctx):
    """rasterization.rs:507-530: (10,20,a1) then (15,25,a2), thr 1 -> (10,25,a2) — reproduced through
    the explicit-area rasterizer with two flat triangles covering one cell at those heights."""
    import recast_rs_b200 as r

    cfg = r.RecastConfig(width=1, height=1, cs=1.0, ch=1.0, bmin=(0, 0, 0), bmax=(1, 40, 1))
    # plane y = a + 9.5 z over the whole cell [0,1]^2: y-range [a, a+9.5] -> span (floor, ceil).
    # (the slope is along z: column 0 keeps polygon parts left of the tile, rasterization.rs:281)
    def tri(a, b=9.5):
        return [(-2, a - 2 * b, -2), (6, a - 2 * b, -2), (-2, a + 6 * b, 6)]
    v = np.array(tri(10.25) + tri(15.25), np.float32).reshape(-1)
    idx = np.arange(6, dtype=np.int32)
    ctx.upload_mesh(v, idx)
    with ctx.rasterize([cfg], np.array([1, 2], np.uint8), 1) as res:
        tv = res.download().tile(0)
    o = oracle.Heightfield.from_config(cfg)
    o.rasterize_triangles(v, idx, np.array([1, 2], np.uint8), 1)
    assert_tile_equal(tv, o.export())
    assert (tv.span_min.tolist(), tv.span_max.tolist(), tv.span_area.tolist()) == ([10], [25], [2])


# ---- KATs 2-4 --------------------------------------------------------------------------
def _flat(oracle, rb, ctx, n):
    return _fixture_hf(oracle, rb, ctx, n, n, (n, 10, n), [(x, z, 0, 5, 1) for z in range(n) for x in range(n)])


def _rows(a, n):
    return ["".join(str(int(v)) for v in a[z * n:(z + 1) * n]) for z in range(n)]


def test_kat2_kat3_distance(oracle, rb, ctx):
    _, hf = _flat(oracle, rb, ctx, 5)
    chf = rb.CompactHeightfield.build_from_heightfield(hf)
    chf.build_distance_field()
    assert chf.conn_ref.size == 144 and chf.max_distance == 3
    assert _rows(chf.dist, 5) == ["00000", "02220", "02120", "02120", "00000"]
    _, hf = _flat(oracle, rb, ctx, 8)
    chf = rb.CompactHeightfield.build_from_heightfield(hf)
    chf.build_distance_field()
    assert chf.max_distance == 6
    assert _rows(chf.dist, 8) == ["00000000", "02222220", "02222220", "02333220", "02344320", "02343220", "02233220",
                                  "00000000"]


def test_kat4_erode(oracle, rb, ctx):
    _, hf = _flat(oracle, rb, ctx, 8)
    chf = rb.CompactHeightfield.build_from_heightfield(hf)
    rb.erode_walkable_area(chf, 2)
    keep = {(x, z) for z in range(8) for x in range(8) if chf.areas[z * 8 + x] == 1}
    assert keep == {(x, z) for x in (2, 3, 4) for z in (3, 4, 5, 6)}
    assert (chf.span_area == 1).all()
    chf.build_distance_field()  # distance field on eroded areas with the ORIGINAL connections
    o, _ = _flat(oracle, rb, ctx, 8)
    oc = oracle.CompactHeightfield.build_from_heightfield(o)
    oc.erode_walkable_area(2)
    oc.build_distance_field()
    e = oc.export()
    np.testing.assert_array_equal(chf.dist, e.dist)
    assert chf.max_distance == e.max_distance


# ---- randomized differential tests ----------------------------------------------------------------------
@pytest.mark.parametrize("seed", range(6))
def test_random_soup_tiles(oracle, rb, ctx, seed):
    import recast_rs_b200 as r
    from recast_rs_b200 import synth

    rng = np.random.default_rng(500 + seed)
    cs, ch = [(0.3, 0.2), (0.5, 0.25), (1.0, 1.0)][seed % 3]
    v, idx = random_soup(rng, 400, 0.0, 40 * cs, -1.0, 8.0, small=cs)
    base = r.RecastConfig(cs=cs, ch=ch, walkable_height=int(rng.integers(1, 5)), walkable_climb=int(rng.integers(0, 3)),
                          walkable_slope_angle=float(rng.choice([30.0, 45.0, 60.0])))
    cfgs = synth.tile_grid(v, [8, 16, 13][seed % 3], base)
    _check_batch(oracle, rb, ctx, v, idx, cfgs, rb.STAGE_RASTER)
    _check_batch(oracle, rb, ctx, v, idx, cfgs, rb.STAGE_RASTER | rb.STAGE_FILTER)
    _check_batch(oracle, rb, ctx, v, idx, cfgs, rb.STAGE_ALL_BUILD)
    _check_batch(oracle, rb, ctx, v, idx, cfgs, rb.STAGE_ALL_BUILD | rb.STAGE_ERODE, erode_radius=1 + seed % 2)


@pytest.mark.parametrize("seed", range(4))
def test_non_finite_vertices(oracle, rb, ctx, seed, monkeypatch):
    """NaN, +-inf and +-3e38 coordinates (tests/parity.py::poison_vertices): the reference checks nothing, its float rules
    decide, and the clipper has to follow them - NaN-ignoring min / max, saturating casts, false comparisons - on both back
    ends.  (The leading-column skip and the NaN-start min / max of k_clip_items were argued from these rules.)"""
    import recast_rs_b200 as r
    from recast_rs_b200 import synth
    from tests.parity import poison_vertices

    rng = np.random.default_rng(900 + seed)
    cs, ch = [(0.3, 0.2), (0.5, 0.25)][seed % 2]
    v, idx = random_soup(rng, 400, 0.0, 40 * cs, -1.0, 8.0, small=cs)
    base = r.RecastConfig(cs=cs, ch=ch, walkable_height=int(rng.integers(1, 5)), walkable_climb=int(rng.integers(0, 3)))
    cfgs = synth.tile_grid(v, [16, 13][seed % 2], base)  # the grid of the clean mesh
    v = poison_vertices(rng, v)
    assert not np.isfinite(v).all()
    _check_batch(oracle, rb, ctx, v, idx, cfgs, rb.STAGE_RASTER)
    _check_batch(oracle, rb, ctx, v, idx, cfgs, rb.STAGE_ALL_BUILD)
    if seed == 0:
        monkeypatch.setenv("RCB_FORCE_GLOBAL", "1")
        cg = rb.Context(0)
        monkeypatch.delenv("RCB_FORCE_GLOBAL")
        _check_batch(oracle, rb, cg, v, idx, cfgs, rb.STAGE_ALL_BUILD)
        cg.close()


@pytest.mark.parametrize("seed", range(3))
def test_vertices_on_cell_lines_and_tile_borders(oracle, rb, ctx, seed, monkeypatch):
    """Every vertex on a multiple of the cell size (and many on tile borders): the `== offset` side of divide_poly
    (rasterization.rs:189-216: the vertex goes to BOTH halves, no intersection is computed), the inclusive bounds test and the
    z = -1 row of a tile's footprint all get exercised on every triangle, on both back ends."""
    import recast_rs_b200 as r

    rng = np.random.default_rng(40 + seed)
    cs = [0.5, 1.0, 0.25][seed]
    step = [1.0, 1.0, 0.5][seed]
    ntri, n = 300, int(24 / step)
    pts = rng.integers(0, n + 1, size=(ntri, 3, 2)).astype(np.float32) * np.float32(step)
    near = rng.integers(-3, 4, size=(ntri, 2, 2)).astype(np.float32) * np.float32(step)  # keep the triangles small
    pts[:, 1:, :] = np.clip(pts[:, :1, :] + near, 0, 24)
    ys = rng.integers(0, 9, size=(ntri, 3)).astype(np.float32) * np.float32(0.75)
    v = np.stack([pts[..., 0], ys, pts[..., 1]], axis=-1).reshape(-1).astype(np.float32)
    idx = np.arange(3 * ntri, dtype=np.int32)
    tile = 8.0  # world units per tile: borders on the lattice
    cfgs = []
    for tz in range(3):
        for tx in range(3):
            c = r.RecastConfig(cs=cs, ch=0.25, walkable_height=3, walkable_climb=2)
            c.bmin, c.bmax = (tx * tile, -0.5, tz * tile), ((tx + 1) * tile, 7.0, (tz + 1) * tile)
            c.width = c.height = int(tile / cs)
            cfgs.append(c)
    _check_batch(oracle, rb, ctx, v, idx, cfgs, rb.STAGE_RASTER)
    _check_batch(oracle, rb, ctx, v, idx, cfgs, rb.STAGE_ALL_BUILD)
    if seed == 0:
        monkeypatch.setenv("RCB_FORCE_GLOBAL", "1")
        cg = rb.Context(0)
        monkeypatch.delenv("RCB_FORCE_GLOBAL")
        _check_batch(oracle, rb, cg, v, idx, cfgs, rb.STAGE_ALL_BUILD)
        cg.close()


def test_walkable_parameters_wrap_like_the_reference(oracle, rb, ctx):
    """RecastConfig::walkable_height / walkable_climb are i32 and reach the filters through `as i16` (lib.rs:278-287), which
    WRAPS (40000 -> -25536, 65540 -> 4); nothing validates them (config.rs:110-136).  One tile per odd pair, same mesh."""
    import recast_rs_b200 as r

    rng = np.random.default_rng(77)
    v, idx = random_soup(rng, 250, 0.0, 8.0, -1.0, 7.0, small=0.5)
    cfgs = []
    for wh, wc in [(0, 0), (1, 0), (255, 3), (32767, 32767), (40000, 40000), (70000, 2), (-3, -1), (3, 40000), (65540, 65537), (2, -2)]:
        c = r.RecastConfig(cs=0.5, ch=0.25, walkable_height=wh, walkable_climb=wc)
        c.bmin, c.bmax, c.width, c.height = (0.0, -0.5, 0.0), (8.0, 7.5, 8.0), 16, 16
        cfgs.append(c)
    _check_batch(oracle, rb, ctx, v, idx, cfgs, rb.STAGE_RASTER | rb.STAGE_FILTER)
    _check_batch(oracle, rb, ctx, v, idx, cfgs, rb.STAGE_ALL_BUILD)


@pytest.mark.parametrize("seed", range(2))
def test_tiny_triangles_around_the_degenerate_threshold(oracle, rb, ctx, seed):
    """Triangles of 1e-5 .. 3e-3 units: their cross products straddle f32::EPSILON, below which lib.rs:258-260 skips the
    triangle; above it the slope test runs on a normal whose length is barely representable.  glam's length (x*x + y*y + z*z
    in that order, sqrt) and normalize (times the reciprocal) have to be followed to the bit for the same triangles to drop."""
    import recast_rs_b200 as r

    rng = np.random.default_rng(60 + seed)
    ntri = 3000
    c = rng.uniform(0.2, 7.8, size=(ntri, 1, 3)).astype(np.float32)
    ext = (10.0 ** rng.uniform(-5.0, -2.5, size=(ntri, 1, 1))).astype(np.float32)
    v = c + rng.uniform(-1, 1, size=(ntri, 3, 3)).astype(np.float32) * ext
    v[..., 1] = np.abs(v[..., 1]) % 4.0
    v, idx = v.reshape(-1).astype(np.float32), np.arange(3 * ntri, dtype=np.int32)
    cfgs = []
    for tz in range(2):
        for tx in range(2):
            cfg = r.RecastConfig(cs=0.25, ch=0.25)
            cfg.bmin, cfg.bmax, cfg.width, cfg.height = (4.0 * tx, -0.5, 4.0 * tz), (4.0 * tx + 4.0, 6.0, 4.0 * tz + 4.0), 16, 16
            cfgs.append(cfg)
    tot = _check_batch(oracle, rb, ctx, v, idx, cfgs, rb.STAGE_RASTER)
    assert 0 < tot["fragments"] < ntri * 4  # some triangles were dropped as degenerate, some were not
    _check_batch(oracle, rb, ctx, v, idx, cfgs, rb.STAGE_ALL_BUILD)


def test_erosion_radius_wraps_like_the_reference(oracle, rb, ctx, monkeypatch):
    """erode_walkable_area compares against `(erosion_radius * 2) as u8` (area.rs:226): 128 -> 0 (nothing eroded), 130 -> 4,
    -1 -> 254 (everything within reach of a border goes).  Both back ends."""
    import recast_rs_b200 as r
    from recast_rs_b200 import synth

    v, idx = synth.terrain(40, 1.0, 2.0, 9)
    cfgs = synth.tile_grid(v, 32, r.RecastConfig(cs=0.5, ch=0.25))
    monkeypatch.setenv("RCB_FORCE_GLOBAL", "1")
    cg = rb.Context(0)
    monkeypatch.delenv("RCB_FORCE_GLOBAL")
    for radius in (0, 1, 127, 128, 130, -1):
        for c in (ctx, cg):
            _check_batch(oracle, rb, c, v, idx, cfgs, rb.STAGE_ALL_BUILD | rb.STAGE_ERODE, erode_radius=radius)
    cg.close()


def test_boundary_distances_saturate_in_eight_bits(oracle, rb, ctx):
    """One open 300x300-cell floor: erode_walkable_area keeps its distances in u8 with saturating adds (area.rs:142-219), so
    everything further than 127 cells from the border sits at 255; a radius of 100 (threshold 200) then eats a 100-cell rim.
    The distance field proper (u16) is far from saturating here and must come out as the oracle's as well."""
    import recast_rs_b200 as r

    v = np.array([0, 1, 0, 150, 1, 0, 150, 1, 150, 0, 1, 150], np.float32)
    idx = np.array([0, 2, 1, 0, 3, 2], np.int32)
    cfg = r.RecastConfig(cs=0.5, ch=0.25)
    cfg.bmin, cfg.bmax, cfg.width, cfg.height = (0.0, 0.0, 0.0), (150.0, 5.0, 150.0), 300, 300
    for radius in (1, 100):
        tot = _check_batch(oracle, rb, ctx, v, idx, [cfg], rb.STAGE_ALL_BUILD | rb.STAGE_ERODE, erode_radius=radius)
        assert tot["spans"] == 300 * 300


def test_permuted_triangle_order(oracle, rb, ctx):
    """Span area is order dependent: the GPU must replay fragments in triangle order."""
    import recast_rs_b200 as r
    from recast_rs_b200 import synth

    rng = np.random.default_rng(42)
    v, idx = random_soup(rng, 300, 0.0, 6.0, 0.0, 5.0)
    cfgs = synth.tile_grid(v, 12, r.RecastConfig(cs=0.5, ch=0.25))
    for k in range(3):
        perm = np.random.default_rng(k).permutation(idx.size // 3)
        pidx = idx.reshape(-1, 3)[perm].reshape(-1).copy()
        _check_batch(oracle, rb, ctx, v, pidx, cfgs, rb.STAGE_RASTER | rb.STAGE_FILTER)


@pytest.mark.parametrize("seed", [1, 2])
def test_terrain_tiles(oracle, rb, ctx, seed):
    from recast_rs_b200 import synth

    v, t = synth.terrain(40, 1.2, 4.0, seed)
    cfgs = synth.tile_grid(v, 64)
    tot = _check_batch(oracle, rb, ctx, v, t, cfgs, rb.STAGE_ALL_BUILD)
    assert tot["triangles"] == 3200 and tot["fragments"] > tot["spans"] > 0


def test_terrain_single_tile_config1_shape(oracle, rb, ctx):
    from recast_rs_b200 import synth

    v, t, cfgs = synth.config1(seed=1, nq=48)  # config 1 scaled down: one tile over the whole mesh
    assert len(cfgs) == 1
    _check_batch(oracle, rb, ctx, v, t, cfgs, rb.STAGE_ALL_BUILD)


def test_interior_multilevel(oracle, rb, ctx):
    from recast_rs_b200 import synth

    v, t, cfgs = synth.config4(seed=1, floors=6, nq=24, tile=32)
    tot = _check_batch(oracle, rb, ctx, v, t, cfgs, rb.STAGE_ALL_BUILD | rb.STAGE_ERODE, erode_radius=1)
    assert tot["spans"] > 3 * tot["cells"]  # many spans per column


def test_open_world_boxes(oracle, rb, ctx):
    from recast_rs_b200 import synth

    v, t = synth.open_world(nq=60, lattice=5)
    cfgs = synth.tile_grid(v, 128)
    _check_batch(oracle, rb, ctx, v, t, cfgs, rb.STAGE_ALL_BUILD)


# ---- edge cases and error behaviour (lib.rs:195-233, config.rs:110-136) -----------------------------------
def test_empty_inputs_and_errors(rb, ctx):
    import recast_rs_b200 as r

    cfg = r.RecastConfig(width=4, height=4, cs=1.0, ch=1.0, bmin=(0, 0, 0), bmax=(4, 4, 4))
    ctx.upload_mesh(np.zeros(0, np.float32), np.zeros(0, np.int32))  # lib.rs:195 -> Ok, empty heightfield
    with ctx.build_tiles([cfg], rb.STAGE_ALL_BUILD) as res:
        tv = res.download().tile(0)
        assert tv.span_count == 0 and tv.hf_cell_offset.tolist() == [0] * 17 and (tv.cell_index == 0xFFFFFFFF).all()
    v = np.array([0, 0, 0, 1, 0, 0, 0, 0, 1], np.float32)
    ctx.upload_mesh(v, np.array([0, 1], np.int32))  # lib.rs:205
    with pytest.raises(r.NavMeshGeneration):
        ctx.build_tiles([cfg], rb.STAGE_RASTER)
    ctx.upload_mesh(v[:8], np.array([0, 1, 2], np.int32))  # lib.rs:199
    with pytest.raises(r.NavMeshGeneration):
        ctx.build_tiles([cfg], rb.STAGE_RASTER)
    ctx.upload_mesh(v, np.array([0, 1, 3], np.int32))  # lib.rs:224-233
    with pytest.raises(r.NavMeshGeneration):
        ctx.build_tiles([cfg], rb.STAGE_RASTER)
    ctx.upload_mesh(v, np.array([0, -1, 2], np.int32))
    with pytest.raises(r.NavMeshGeneration):
        ctx.build_tiles([cfg], rb.STAGE_RASTER)
    bad = r.RecastConfig(width=0, height=4, cs=1.0, ch=1.0)
    ctx.upload_mesh(v, np.array([0, 1, 2], np.int32))
    with pytest.raises(r.InvalidMesh):
        ctx.build_tiles([bad], rb.STAGE_RASTER)
    with ctx.build_tiles([], rb.STAGE_ALL_BUILD) as res:  # zero tiles
        assert res.totals()["tiles"] == 0


def test_recast_builder_facade(oracle, rb, ctx):
    """RecastBuilder::build_heightfield (lib.rs:116) — the reference has no test for it; compare with the oracle."""
    import recast_rs_b200 as r
    from recast_rs_b200 import synth

    v, t, cfgs = synth.config1(seed=3, nq=20)
    builder = r.RecastBuilder(cfgs[0], ctx)
    hf = builder.build_heightfield(v, t)
    o = oracle.Heightfield.from_config(cfgs[0])
    o.builder_rasterize(cfgs[0], v, t, run_filters=True)
    e = o.export()
    np.testing.assert_array_equal(hf.cell_offset, e.cell_offset)
    np.testing.assert_array_equal(hf.span_area, e.span_area)
    chf = builder.build_compact_heightfield(hf)
    chf.build_distance_field()
    oc = oracle.CompactHeightfield.build_from_heightfield(o)
    oc.build_distance_field()
    ea = oc.export()
    np.testing.assert_array_equal(chf.dist, ea.dist)
    np.testing.assert_array_equal(chf.conn_ref, ea.conn_ref)
    assert chf.max_distance == ea.max_distance


def test_precision_collapse_far_from_origin(oracle, rb, ctx):
    """Coordinates so large that consecutive cell boundaries round to the same float: clip polygons pick up
    extra on-the-line vertices.  The fast clipper's 6-vertex slots may overflow; the library must then fall
    back to the wide clipper and still match the oracle bit for bit."""
    import recast_rs_b200 as r
    from recast_rs_b200 import synth

    rng = np.random.default_rng(77)
    base = np.float32(4.0e6)  # ulp(4e6) = 0.5 > cs
    v, idx = random_soup(rng, 200, 0.0, 12.0, 0.0, 5.0, small=0.5)
    v = v.reshape(-1, 3).copy()
    v[:, 0] += base
    v[:, 2] += base
    v = v.astype(np.float32).reshape(-1)
    cfgs = synth.tile_grid(v, 16, r.RecastConfig(cs=0.3, ch=0.2))
    cfgs = cfgs[:40]
    _check_batch(oracle, rb, ctx, v, idx, cfgs, rb.STAGE_ALL_BUILD)


def test_sparse_world_mostly_empty_tiles(oracle, rb, ctx):
    """Few small triangles over many tiles: most tiles receive no fragment at all (tile-resident back end)."""
    import recast_rs_b200 as r
    from recast_rs_b200 import synth

    rng = np.random.default_rng(5)
    v, idx = random_soup(rng, 12, 0.0, 30.0, 0.0, 3.0, small=0.4)
    v = np.concatenate([v, np.array([0, 0, 0, 0.1, 0, 0, 0, 0, 0.1, 30, 0, 30, 30.1, 0, 30, 30, 0, 30.1], np.float32)])  # pin the AABB
    idx = np.concatenate([idx, np.arange(36, 42, dtype=np.int32)])
    cfgs = synth.tile_grid(v, 8, r.RecastConfig(cs=0.5, ch=0.25))
    tot = _check_batch(oracle, rb, ctx, v, idx, cfgs, rb.STAGE_ALL_BUILD)
    assert tot["tiles"] > 30 and tot["spans"] < tot["cells"] / 4


@pytest.mark.parametrize("shape", [(1, 1), (1, 37), (53, 1), (7, 5), (64, 3)])
def test_odd_tile_shapes(oracle, rb, ctx, shape):
    """Tiles that are one cell wide / tall or otherwise not warp-shaped."""
    import recast_rs_b200 as r

    w, h = shape
    rng = np.random.default_rng(w * 100 + h)
    cs, ch = 0.5, 0.25
    v, idx = random_soup(rng, 60, 0.0, max(w, h) * cs, 0.0, 4.0, small=0.3)
    cfgs = [r.RecastConfig(width=w, height=h, cs=cs, ch=ch, bmin=(0.0, -0.5, 0.0), bmax=(w * cs, 5.0, h * cs)),
            r.RecastConfig(width=h, height=w, cs=cs, ch=ch, bmin=(0.25, -0.5, 0.1), bmax=(0.25 + h * cs, 5.0, 0.1 + w * cs))]
    _check_batch(oracle, rb, ctx, v, idx, cfgs, rb.STAGE_ALL_BUILD | rb.STAGE_ERODE, erode_radius=1)


def test_heterogeneous_tile_sizes_in_one_batch(oracle, rb, ctx):
    """Arbitrary tile AABBs / sizes / agent parameters in one batch (overlapping tiles allowed)."""
    import recast_rs_b200 as r

    rng = np.random.default_rng(11)
    v, idx = random_soup(rng, 300, 0.0, 20.0, -1.0, 6.0, small=0.4)
    cfgs = []
    for k in range(9):
        w, h = int(rng.integers(3, 40)), int(rng.integers(3, 40))
        cs = float(rng.choice([0.25, 0.5, 0.3]))
        bx, bz = float(rng.uniform(-2, 12)), float(rng.uniform(-2, 12))
        cfgs.append(r.RecastConfig(width=w, height=h, cs=cs, ch=float(rng.choice([0.2, 0.5])), bmin=(bx, float(rng.uniform(-1, 1)), bz),
                                   bmax=(bx + w * cs, float(rng.uniform(4, 7)), bz + h * cs), walkable_height=int(rng.integers(1, 6)),
                                   walkable_climb=int(rng.integers(0, 4)), walkable_slope_angle=float(rng.uniform(20, 70))))
    _check_batch(oracle, rb, ctx, v, idx, cfgs, rb.STAGE_ALL_BUILD)


# ---- ordered / unordered columns: the bisection fast path of the span-parallel link and ledge filter --------------
def _columns_case(rng, w, h, depth, ordered):
    """CSR heightfield with up to `depth` spans per column.  ordered: disjoint ascending spans (what the rasterizer
    leaves); else: some columns shuffled and overlapping, as Heightfield::add_span or a caller's CSR can hold."""
    off, mn, mx, ar = [0], [], [], []
    for _ in range(w * h):
        n = int(rng.integers(0, depth + 1))
        y, col = int(rng.integers(0, 6)), []
        for _ in range(n):
            t = int(rng.integers(1, 5))
            col.append((y, y + t, int(rng.choice([0, 1, 1, 1, 63]))))
            y += t + int(rng.integers(0, 9))
        if not ordered and n > 1 and rng.random() < 0.4:
            rng.shuffle(col)
            col = [(a - int(rng.integers(0, 3)), b + int(rng.integers(0, 3)), c) for a, b, c in col]
        for a, b, c in col:
            mn.append(a); mx.append(b); ar.append(c)
        off.append(len(mn))
    return np.array(off, np.uint32), np.array(mn, np.int16), np.array(mx, np.int16), np.array(ar, np.uint8)


@pytest.mark.parametrize("ordered", [True, False])
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_deep_columns_filters_link_distance(oracle, rb, ctx, seed, ordered):
    rng = np.random.default_rng(100 + seed)
    w, h = 23, 17
    off, mn, mx, ar = _columns_case(rng, w, h, 24 if seed else 5, ordered)
    bmax = (w * 0.5, 200.0, h * 0.5)
    height, climb = (4, 2) if seed != 2 else (9, 5)
    o = oracle.Heightfield.from_csr(w, h, (0, 0, 0), bmax, 0.5, 0.5, off, mn, mx, ar)
    o.filter_low_hanging_walkable_obstacles(climb)
    o.filter_ledge_spans(height, climb)
    o.filter_walkable_low_height_spans(height)
    hfa = o.export()
    oc = oracle.CompactHeightfield.build_from_heightfield(o)
    oc.erode_walkable_area(1)
    oc.build_distance_field()
    chfa = oc.export()
    from recast_rs_b200.config import RecastConfig
    cfg = RecastConfig(width=w, height=h, cs=0.5, ch=0.5, bmin=(0, 0, 0), bmax=bmax, walkable_height=height, walkable_climb=climb)
    bits = rb.STAGE_FILTER | rb.STAGE_COMPACT | rb.STAGE_ERODE | rb.STAGE_DIST
    with ctx.build_from_spans([cfg], off, mn, mx, ar, bits, 1) as r:
        r.download()
        assert_tile_equal(r.tile(0), hfa, chfa, what=f"deep columns ordered={ordered}")


def test_global_raster_with_16_byte_fragment_records(oracle, rb, monkeypatch):
    """The global rasterizer packs fragments into 8-byte records while triangle indices fit 24 bits and into 16-byte
    ones beyond; RCB_WIDE_RECORDS forces the wide form so that it is exercised without a 16 M-triangle mesh."""
    from recast_rs_b200 import synth
    from recast_rs_b200.config import RecastConfig
    from tests.test_gpu_fullsize import _context

    v, t = synth.interior(5, 14, 1.2, 0.3, 7, 1.5)
    cfgs = synth.tile_grid(v, 32, RecastConfig(cs=0.3, ch=0.2))
    monkeypatch.setenv("RCB_WIDE_RECORDS", "1")
    c = _context(True)
    try:
        _check_batch(oracle, rb, c, v, t, cfgs, rb.STAGE_ALL_BUILD)
    finally:
        c.close()


def test_u16_indices_and_triangle_soup_entries(oracle):
    """rasterize_triangles_u16 (rasterization.rs:432) and rasterize_triangle_soup (:459): the same rasterizer behind two
    other index forms.  Both must give the heightfield of the i32-indexed call on the same triangles."""
    import recast_rs_b200.builder as rb
    from recast_rs_b200.config import RecastConfig

    rng = np.random.default_rng(11)
    v, t = random_soup(rng, 500, 0.0, 12.0, 0.0, 5.0)
    areas = rng.integers(0, 4, size=t.size // 3).astype(np.uint8)
    cfg = RecastConfig(width=40, height=40, cs=0.3, ch=0.2, bmin=(0.0, 0.0, 0.0), bmax=(12.0, 6.0, 12.0))
    ohf = oracle.Heightfield.from_config(cfg)
    ohf.rasterize_triangles(v, t, areas, 2)
    want = ohf.export()
    ctx = rb.Context(0)
    got = []
    ctx.upload_mesh(v, t)
    with ctx.rasterize([cfg], areas, 2) as r:
        got.append(r.download().tile(0))
    ctx.upload_mesh_u16(v, t.astype(np.uint16))
    with ctx.rasterize([cfg], areas, 2) as r:
        got.append(r.download().tile(0))
    ctx.upload_triangle_soup(v.reshape(-1, 3)[t].reshape(-1))
    with ctx.rasterize([cfg], areas, 2) as r:
        got.append(r.download().tile(0))
    for g in got:
        np.testing.assert_array_equal(g.hf_cell_offset, want.cell_offset)
        np.testing.assert_array_equal(g.span_min, want.span_min)
        np.testing.assert_array_equal(g.span_max, want.span_max)
        np.testing.assert_array_equal(g.span_area, want.span_area)
    with pytest.raises(ValueError):
        ctx.upload_triangle_soup(np.zeros(10, np.float32))
    ctx.close()
