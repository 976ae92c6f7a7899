"""Shared helpers for the GPU parity tests: byte-for-byte comparison of one tile's device output
against the CPU oracle (SURVEY.md §8(c) 'parity method')."""
import numpy as np


def assert_tile_equal(tv, hfa, chfa=None, what=""):
    """tv: recast_rs_b200.builder.TileVoxels; hfa/chfa: oracle.HfArrays / oracle.ChfArrays."""
    def eq(a, b, name):
        a, b = np.asarray(a), np.asarray(b)
        assert a.shape == b.shape, f"{what} {name}: shape {a.shape} != {b.shape}"
        if not np.array_equal(a, b):
            bad = np.flatnonzero(a != b)
            raise AssertionError(f"{what} {name}: {bad.size} mismatches, first at {bad[:8].tolist()}: "
                                 f"got {a[bad[:8]].tolist()} want {b[bad[:8]].tolist()}")

    eq(tv.hf_cell_offset, hfa.cell_offset, "hf_cell_offset")
    if chfa is None:
        eq(tv.span_min, hfa.span_min, "span_min")
        eq(tv.span_max, hfa.span_max, "span_max")
        eq(tv.span_area, hfa.span_area, "span_area")
        return
    eq(tv.span_min, chfa.span_min, "span_min")
    eq(tv.span_max, chfa.span_max, "span_max")
    eq(tv.span_area, chfa.span_area, "span_area")
    eq(tv.cell_index, chfa.cell_index, "cell_index")
    eq(tv.cell_count, chfa.cell_count, "cell_count")
    eq(tv.areas, chfa.areas, "areas")
    eq(tv.con, chfa.con, "con")
    eq(tv.first_conn, chfa.first_conn, "first_conn")
    eq(tv.conn_ref, chfa.conn_ref, "conn_ref")
    eq(tv.conn_dir, chfa.conn_dir, "conn_dir")
    eq(tv.conn_next, chfa.conn_next, "conn_next")
    eq(tv.dist, chfa.dist, "dist")
    assert tv.max_height == chfa.max_height, f"{what} max_height {tv.max_height} != {chfa.max_height}"
    assert tv.max_distance == chfa.max_distance, f"{what} max_distance {tv.max_distance} != {chfa.max_distance}"
    assert tv.walkable_span_count == chfa.walkable_span_count, f"{what} walkable_span_count"
    assert tv.status == 0, f"{what} status {tv.status}"


def random_soup(rng, ntri, lo, hi, ylo, yhi, small=0.5):
    """Triangles of mixed size; some degenerate, some vertical walls, some outside the bounds."""
    verts = []
    for _ in range(ntri):
        c = rng.uniform(lo - 1.0, hi + 1.0, size=2)
        kind = int(rng.integers(0, 6))
        ext = {0: small, 1: 2.0, 2: 6.0, 3: small, 4: 1.0, 5: 12.0}[kind]
        tri = [(c[0] + rng.uniform(-ext, ext), rng.uniform(ylo, yhi), c[1] + rng.uniform(-ext, ext)) for _k in range(3)]
        if kind == 3:
            tri[2] = tri[1]
        if kind == 4:
            tri[1] = (tri[0][0], tri[1][1], tri[1][2])
            tri[2] = (tri[0][0], tri[2][1], tri[2][2])
        verts += tri
    v = np.array(verts, dtype=np.float32).reshape(-1)
    return v, np.arange(ntri * 3, dtype=np.int32)


def poison_vertices(rng, verts, frac=0.06):
    """Some vertices get one non-finite (or barely finite) coordinate: NaN, +-inf, +-3e38.  The reference has no check for
    them: f32::min / max skip a NaN, `as i32` / `as i16` saturate and turn NaN into 0, comparisons with NaN are false -
    whatever comes out of that is what the GPU path has to produce too."""
    v = np.array(verts, dtype=np.float32).reshape(-1, 3).copy()
    rows = rng.choice(v.shape[0], size=max(1, int(frac * v.shape[0])), replace=False)
    vals = np.array([np.nan, np.inf, -np.inf, 3.0e38, -3.0e38], np.float32)
    for r in rows:
        v[r, int(rng.integers(0, 3))] = vals[int(rng.integers(0, vals.size))]
    return v.reshape(-1)
