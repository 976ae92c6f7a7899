"""Pins the CPU oracle against every expectation the reference's own unit tests hold for the hot
path and against the hand-derived KATs of SURVEY.md §8(c).  CPU only.

Reference tests mirrored (paths under /root/reference/crates/recast/src/):
  rasterization.rs:484-504  test_add_span_simple
  rasterization.rs:507-530  test_add_span_merge
  rasterization.rs:533-559  test_rasterize_triangle
  heightfield.rs:1206-1219  test_heightfield_creation
  heightfield.rs:1222-1250  test_add_span
  heightfield.rs:1417-1452  test_filter_low_hanging_walkable_obstacles
  compact_heightfield.rs:1028-1059 test_compact_heightfield_creation
  compact_heightfield.rs:1093-1117 test_get_neighbor
"""
import numpy as np

DIR_NW, DIR_N, DIR_NE, DIR_E, DIR_SE, DIR_S, DIR_SW, DIR_W = range(8)
NONE = 0xFFFFFFFF


def _hf10(orc):
    return orc.Heightfield(10, 10, (0, 0, 0), (10, 10, 10), 1.0, 1.0)


def test_add_span_simple(oracle):
    hf = _hf10(oracle)
    hf.add_span_raster(5, 5, 10, 20, 1, 1)
    assert hf.export().column(5, 5, 10) == [(10, 20, 1)]


def test_add_span_merge(oracle):
    hf = _hf10(oracle)
    hf.add_span_raster(5, 5, 10, 20, 1, 1)
    hf.add_span_raster(5, 5, 15, 25, 2, 1)
    assert hf.export().column(5, 5, 10) == [(10, 25, 2)]


def test_add_span_out_of_bounds_ignored(oracle):
    hf = _hf10(oracle)
    for x, z in [(-1, 0), (0, -1), (10, 0), (0, 10)]:
        hf.add_span_raster(x, z, 0, 1, 1, 1)  # rasterization.rs:33 silently ignored
    assert hf.span_count() == 0


def test_rasterize_triangle_some_span(oracle):
    hf = _hf10(oracle)
    hf.rasterize_triangle((2, 0, 2), (8, 0, 2), (5, 2, 8), 1, 1)
    assert hf.span_count() > 0
    assert not hf.poly_overflow


def test_heightfield_creation_and_add_span(oracle):
    hf = _hf10(oracle)
    e = hf.export()
    assert e.cell_offset.size == 10 * 10 + 1 and e.span_min.size == 0
    hf.add_span(5, 5, 3, 6, 1)
    assert hf.export().column(5, 5, 10) == [(3, 6, 1)]


def test_heightfield_add_span_errors(oracle):
    import pytest

    hf = _hf10(oracle)
    with pytest.raises(oracle.OracleError) as e:
        hf.add_span(10, 0, 0, 1, 1)  # heightfield.rs:103-108
    assert e.value.status == -2
    with pytest.raises(oracle.OracleError):
        hf.add_span(0, 0, 5, 1, 1)  # heightfield.rs:110-115


def test_filter_low_hanging_walkable_obstacles(oracle):
    hf = oracle.Heightfield(5, 5, (0, 0, 0), (5, 10, 5), 1.0, 1.0)
    hf.add_span(2, 2, 0, 5, 1)
    hf.add_span(2, 2, 6, 8, 0)
    hf.add_span(2, 2, 9, 15, 1)
    hf.filter_low_hanging_walkable_obstacles(3)
    assert hf.export().column(2, 2, 5) == [(0, 5, 1), (6, 8, 1), (9, 15, 1)]


def _kat1(orc):
    hf = orc.Heightfield(3, 3, (0, 0, 0), (3, 3, 3), 1.0, 1.0)
    hf.add_span(0, 0, 0, 2, 1)
    hf.add_span(0, 0, 3, 5, 1)
    hf.add_span(1, 0, 0, 2, 1)
    hf.add_span(0, 1, 0, 2, 1)
    return hf


def test_compact_heightfield_creation(oracle):
    chf = oracle.CompactHeightfield.build_from_heightfield(_kat1(oracle))
    a = chf.export()
    assert chf.width == 3 and chf.height == 3
    assert a.cell_index.size == 9
    assert a.span_min.size == 4
    assert a.conn_ref.size > 0


def test_get_neighbor(oracle):
    hf = oracle.Heightfield(2, 1, (0, 0, 0), (2, 1, 1), 1.0, 1.0)
    hf.add_span(0, 0, 0, 1, 1)
    hf.add_span(1, 0, 0, 1, 1)
    chf = oracle.CompactHeightfield.build_from_heightfield(hf)
    assert chf.get_neighbor(0, DIR_E) == 1


# ---- hand-derived KATs; also re-derived by oracle/pymodel.py in test_oracle_vs_pymodel ----
def test_kat1_cells_connections(oracle):
    chf = oracle.CompactHeightfield.build_from_heightfield(_kat1(oracle))
    a = chf.export()
    assert list(a.cell_index[:4]) == [0, 2, NONE, 3] and list(a.cell_count[:4]) == [2, 1, 0, 1]
    nei = {s: {d: chf.get_neighbor(s, d) for d in range(8) if chf.get_neighbor(s, d) is not None} for s in range(4)}
    assert nei == {0: {DIR_E: 2, DIR_S: 3}, 1: {}, 2: {DIR_SW: 3, DIR_W: 0}, 3: {DIR_N: 0, DIR_NE: 2}}
    assert a.con.reshape(-1, 4).tolist() == [[63, 0, 0, 63], [63, 63, 63, 63], [0, 63, 63, 63], [63, 63, 63, 0]]
    conns = list(zip(a.conn_ref.tolist(), a.conn_dir.tolist(), a.conn_next.tolist()))
    assert conns == [(3, 5, NONE), (2, 3, 0), (0, 7, NONE), (3, 6, 2), (2, 2, NONE), (0, 1, 4)]
    assert a.first_conn.tolist() == [1, NONE, 3, 5]
    assert a.max_height == 5 and a.walkable_span_count == 4


def _flat(orc, n):
    hf = orc.Heightfield(n, n, (0, 0, 0), (n, 10, n), 1.0, 1.0)
    for z in range(n):
        for x in range(n):
            hf.add_span(x, z, 0, 5, 1)
    return hf


def _rows(a, n):
    return ["".join(str(int(v)) for v in a[z * n:(z + 1) * n]) for z in range(n)]


def test_kat2_distance_5x5(oracle):
    chf = oracle.CompactHeightfield.build_from_heightfield(_flat(oracle, 5))
    chf.build_distance_field()
    a = chf.export()
    assert a.conn_ref.size == 144
    assert a.max_distance == 3
    assert _rows(a.dist, 5) == ["00000", "02220", "02120", "02120", "00000"]


def test_kat3_distance_8x8(oracle):
    chf = oracle.CompactHeightfield.build_from_heightfield(_flat(oracle, 8))
    chf.build_distance_field()
    a = chf.export()
    assert a.max_distance == 6
    assert _rows(a.dist, 8) == ["00000000", "02222220", "02222220", "02333220", "02344320", "02343220", "02233220",
                                "00000000"]


def test_kat4_erode_8x8(oracle):
    chf = oracle.CompactHeightfield.build_from_heightfield(_flat(oracle, 8))
    chf.erode_walkable_area(2)
    a = chf.export()
    keep = {(x, z) for z in range(8) for x in range(8) if a.areas[z * 8 + x] == 1}
    assert keep == {(x, z) for x in (2, 3, 4) for z in (3, 4, 5, 6)}
    assert (a.span_area == 1).all()  # area.rs:229 only chf.areas is written


# ---- config.rs -------------------------------------------------------------------------------
def test_config_default_and_grid(oracle):
    c = oracle.config_default()
    assert (c.cs, c.ch) == (np.float32(0.3), np.float32(0.2))
    assert (c.walkable_height, c.walkable_climb, c.walkable_radius, c.border_size) == (2, 1, 1, 0)
    c.cs = 1.0
    g = oracle.calculate_grid_size(c, (0, 0, 0), (10, 5, 7.5))
    assert (g.width, g.height) == (11, 8)  # exact division -> +1 ; otherwise ceil (config.rs:96-106)


def test_config_validate(oracle):
    c = oracle.config_default()
    assert oracle.config_validate(c) == -1  # width/height 0
    c.width = c.height = 4
    assert oracle.config_validate(c) == 0
    c.walkable_slope_angle = 91.0
    assert oracle.config_validate(c) == -1
    c.walkable_slope_angle = 45.0
    c.max_vertices_per_polygon = 2
    assert oracle.config_validate(c) == -1
