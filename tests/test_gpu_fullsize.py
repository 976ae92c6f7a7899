"""BASELINE.json's full-size configurations on the GPU: size-independent properties over every tile,
bit-exact agreement of the two independent device back ends (tile-resident shared-memory kernels vs the
global-memory v1 pipeline), and oracle parity on a sample of tiles."""
import os

import numpy as np
import pytest

from tests.parity import assert_tile_equal

pytestmark = pytest.mark.gpu


def _context(force_global=False, band_cells=None):
    """band_cells: cut tiles that do not fit shared memory whole into bands of about that many cells (RCB_BAND_CELLS)."""
    import recast_rs_b200.builder as rb

    env = {"RCB_FORCE_GLOBAL": "1" if force_global else "0"}
    if band_cells is not None:
        env["RCB_BAND_CELLS"] = str(band_cells)
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        return rb.Context(0)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def _invariants(tv, what, eroded=False):
    """Properties every tile satisfies (SURVEY.md Appendix B invariants + layout rules of §8(a) row BC)."""
    C, S, K = tv.width * tv.height, tv.span_count, tv.conn_count
    off = tv.hf_cell_offset.astype(np.int64)
    assert off[0] == 0 and off[-1] == S and (np.diff(off) >= 0).all(), what
    cnt = np.diff(off)
    np.testing.assert_array_equal(tv.cell_count, cnt)
    np.testing.assert_array_equal(tv.cell_index[cnt > 0], off[:-1][cnt > 0])
    assert (tv.cell_index[cnt == 0] == 0xFFFFFFFF).all()
    if S:
        # columns ascending and strictly separated: span[k+1].min > span[k].max
        same_col = np.ones(S, bool)
        same_col[off[1:-1][off[1:-1] < S]] = False  # first span of each later column
        k = np.flatnonzero(same_col[1:]) + 1
        assert (tv.span_min[k].astype(int) > tv.span_max[k - 1].astype(int)).all(), what
        assert (tv.span_min >= 0).all()
        if not eroded:  # after erosion connections survive, so this no longer holds
            assert (tv.dist[tv.areas == 0] == 0).all()  # no connections => boundary => 0
        assert tv.walkable_span_count == int((tv.span_area != 0).sum())
        assert tv.max_height == int(tv.span_max.astype(np.uint16).max())
    if K:
        assert (tv.conn_ref < S).all() and (tv.conn_dir < 8).all()
        assert (tv.span_area[tv.conn_ref] != 0).all()  # connections only reach spans walkable at compaction
        nconn = np.zeros(S, np.int64)
        has = tv.first_conn != 0xFFFFFFFF
        # first_conn is the LAST entry of the span's block (blocks are laid out in span order)
        ends = tv.first_conn[has].astype(np.int64) + 1
        assert (np.diff(ends) > 0).all() and ends[-1] == K
        begins = np.concatenate([[0], ends[:-1]])
        nconn[has] = ends - begins
        assert (nconn[tv.span_area == 0] == 0).all()  # only walkable spans own connections
        # inside a block directions strictly descend (pushed in reverse order)
        blk = np.repeat(np.arange(ends.size), ends - begins)
        d = tv.conn_dir.astype(int)
        same = blk[1:] == blk[:-1]
        assert (d[1:][same] < d[:-1][same]).all()
        nxt = tv.conn_next.astype(np.int64)
        idx = np.arange(K)
        first_in_blk = np.isin(idx, begins)
        assert (nxt[first_in_blk] == 0xFFFFFFFF).all() and (nxt[~first_in_blk] == idx[~first_in_blk] - 1).all()


def _compare_results(ra, rb_, ntiles, names):
    for i in range(ntiles):
        a, b = ra.tile(i), rb_.tile(i)
        for k in names:
            x, y = getattr(a, k), getattr(b, k)
            if isinstance(x, np.ndarray):
                assert np.array_equal(x, y), f"tile {i}: {k} differs between back ends"
            else:
                assert x == y, f"tile {i}: {k} differs between back ends"


NAMES = ("hf_cell_offset", "span_min", "span_max", "span_area", "cell_index", "cell_count", "areas", "con", "first_conn",
         "conn_ref", "conn_dir", "conn_next", "dist", "span_count", "conn_count", "walkable_span_count", "max_height",
         "max_distance", "status")


def test_config2_full(oracle):
    """1M-triangle terrain, 2025 tiles of 64x64 cells (BASELINE config 2)."""
    import recast_rs_b200.builder as rb
    from recast_rs_b200 import synth

    v, t, cfgs = synth.config2(seed=1)
    assert t.size // 3 == 1002528 and len(cfgs) == 2025
    ct, cg = _context(False), _context(True)
    for c in (ct, cg):
        c.upload_mesh(v, t)
    with ct.build_tiles(cfgs, rb.STAGE_ALL_BUILD) as r1, cg.build_tiles(cfgs, rb.STAGE_ALL_BUILD) as r2:
        r1.download(), r2.download()
        t1, t2 = r1.totals(), r2.totals()
        t1.pop("result_bytes"), t2.pop("result_bytes")  # arena padding differs between the back ends
        assert t1 == t2 and t1["tiles"] == 2025 and t1["cells"] == sum(c.width * c.height for c in cfgs)
        _compare_results(r1, r2, len(cfgs), NAMES)
        for i in range(0, len(cfgs), 7):
            _invariants(r1.tile(i), f"config2 tile {i}")
        # corners, centre, the partial tiles of the last row / column and a spread of interior tiles against the oracle
        picks = sorted({0, 44, 1012, 1980, 2024, 22, 45 * 22, 45 * 22 + 44, 2002} | set(range(3, 2025, 127)))
        assert len(picks) >= 25
        for i in picks:
            hfa, chfa = oracle.build_tile(cfgs[i], v, t)
            assert_tile_equal(r1.tile(i), hfa, chfa, what=f"config2 tile {i}")
    ct.close(), cg.close()


def test_config1_full_single_tile(oracle):
    """100k-triangle terrain, ONE tile of ~896x896 cells (BASELINE config 1; global-memory back end)."""
    import recast_rs_b200.builder as rb
    from recast_rs_b200 import synth

    v, t, cfgs = synth.config1(seed=1)
    ctx = _context(False)
    ctx.upload_mesh(v, t)
    with ctx.build_tiles(cfgs, rb.STAGE_ALL_BUILD | rb.STAGE_ERODE, 1) as r:
        tv = r.download().tile(0)
        hfa, chfa = oracle.build_tile(cfgs[0], v, t, rb.STAGE_ALL_BUILD | rb.STAGE_ERODE, 1)
        assert_tile_equal(tv, hfa, chfa, what="config1")
        _invariants(tv, "config1", eroded=True)
    ctx.close()


def test_config4_full_interior(oracle):
    """1M-triangle 16-floor interior, 64x64 tiles, >= 16 spans per column (BASELINE config 4)."""
    import recast_rs_b200.builder as rb
    from recast_rs_b200 import synth

    v, t, cfgs = synth.config4(seed=1)
    assert t.size // 3 == 1002528
    ctx, cb = _context(False), _context(False, band_cells=512)  # global-memory back end / shared-memory bands of 8 rows
    for c in (ctx, cb):
        c.upload_mesh(v, t)
    cb.profile_enable(True)
    with ctx.build_tiles(cfgs, rb.STAGE_ALL_BUILD) as r, cb.build_tiles(cfgs, rb.STAGE_ALL_BUILD) as r2:
        r.download(), r2.download()
        tot, tot2 = r.totals(), r2.totals()
        assert tot["spans"] >= 15 * tot["cells"] * 0.9
        tot.pop("result_bytes"), tot2.pop("result_bytes")
        assert tot == tot2
        _compare_results(r, r2, len(cfgs), NAMES)  # the two back ends agree on every tile
        for i in range(0, len(cfgs), 5):
            _invariants(r.tile(i), f"config4 tile {i}")
        for i in (0, len(cfgs) // 2, len(cfgs) - 1):
            hfa, chfa = oracle.build_tile(cfgs[i], v, t)
            assert_tile_equal(r.tile(i), hfa, chfa, what=f"config4 tile {i}")
    assert "k_tile_compact" in cb.profile_read()  # the banded context really took the tile-resident kernels
    ctx.close(), cb.close()


def test_config3_scaled_open_world(oracle):
    """Open world with buildings, 128x128 tiles (BASELINE config 3 at 1/16 of the triangles so the oracle
    sample stays within seconds; the full 10M-triangle world is exercised by bench.py --config 3)."""
    import recast_rs_b200.builder as rb
    from recast_rs_b200 import synth

    v, t, cfgs = synth.config3(seed=1, nq=550, lattice=41)
    ct, cg = _context(False), _context(True)
    for c in (ct, cg):
        c.upload_mesh(v, t)
    with ct.build_tiles(cfgs, rb.STAGE_ALL_BUILD) as r1, cg.build_tiles(cfgs, rb.STAGE_ALL_BUILD) as r2:
        r1.download(), r2.download()
        ta, tb = r1.totals(), r2.totals()
        ta.pop("result_bytes"), tb.pop("result_bytes")
        assert ta == tb
        _compare_results(r1, r2, len(cfgs), NAMES)
        for i in range(0, len(cfgs), 11):
            _invariants(r1.tile(i), f"config3 tile {i}")
        for i in (0, len(cfgs) // 2, len(cfgs) - 1):
            hfa, chfa = oracle.build_tile(cfgs[i], v, t)
            assert_tile_equal(r1.tile(i), hfa, chfa, what=f"config3 tile {i}")
    ct.close(), cg.close()


def test_config3_full(oracle):
    """BASELINE config 3 at FULL size: the 10M-triangle open world, 4761 tiles of 128x128 cells.  Both back ends (global
    memory; shared-memory bands of 32 rows) must agree on every tile; invariants on a stride; corners, centre and the
    building-densest tiles against the oracle (which scans all 10M triangles per tile, as lib.rs:217 does)."""
    import recast_rs_b200.builder as rb
    from recast_rs_b200 import shard, synth

    v, t, cfgs = synth.config3(seed=1)
    assert t.size // 3 == 10002752 and len(cfgs) == 69 * 69
    cg, cb = _context(True), _context(False, band_cells=4608)
    for c in (cg, cb):
        c.upload_mesh(v, t)
    cb.profile_enable(True)
    with cg.build_tiles(cfgs, rb.STAGE_ALL_BUILD) as r1, cb.build_tiles(cfgs, rb.STAGE_ALL_BUILD) as r2:
        r1.download(), r2.download()
        ta, tb = r1.totals(), r2.totals()
        ta.pop("result_bytes"), tb.pop("result_bytes")
        assert ta == tb and ta["tiles"] == 4761 and ta["triangles"] == 10002752
        _compare_results(r1, r2, len(cfgs), NAMES)
        for i in range(0, len(cfgs), 37):
            _invariants(r1.tile(i), f"config3 tile {i}")
        dense = np.argsort(shard.tile_costs(v, t, cfgs))[-2:].tolist()  # the two tiles with the most building area
        for i in sorted({0, 68, 69 * 34 + 34, 69 * 68, 4760} | set(dense)):
            hfa, chfa = oracle.build_tile(cfgs[i], v, t)
            assert_tile_equal(r1.tile(i), hfa, chfa, what=f"config3 tile {i}")
    assert "k_tile_compact" in cb.profile_read()
    cg.close(), cb.close()


def test_determinism_repeated_runs():
    """Atomics decide where fragments land, never what the result is: repeated runs are bit-identical."""
    import recast_rs_b200.builder as rb
    from recast_rs_b200 import synth

    v, t, cfgs = synth.config2(seed=3, nq=200, tile=64)
    ctx = _context(False)
    ctx.upload_mesh(v, t)
    ref = None
    for _ in range(3):
        with ctx.build_tiles(cfgs, rb.STAGE_ALL_BUILD) as r:
            r.download()
            sig = [(r.tile(i).dist.tobytes(), r.tile(i).span_area.tobytes(), r.tile(i).conn_ref.tobytes()) for i in range(0, len(cfgs), 3)]
        assert ref is None or sig == ref
        ref = sig
    ctx.close()


def test_seventy_thousand_tiny_tiles_both_back_ends(oracle):
    """More tiles than one grid dimension holds (65535): 2x2-cell tiles over a small terrain, both back ends."""
    import recast_rs_b200.builder as rb
    from recast_rs_b200 import synth
    from recast_rs_b200.config import RecastConfig

    v, t = synth.terrain(160, 1.2, 4.0, 3)
    cfgs = synth.tile_grid(v, 2, RecastConfig(cs=0.3, ch=0.2))
    assert len(cfgs) > 70000
    res = []
    for fg in (False, True):
        ctx = _context(force_global=fg)
        ctx.upload_mesh(v, t)
        with ctx.build_tiles(cfgs, rb.STAGE_ALL_BUILD) as r:
            r.download()
            tot = r.totals()
            pick = [0, 1, 65534, 65535, 65536, len(cfgs) - 1, len(cfgs) // 2]
            res.append((tot, [r.tile(i) for i in pick], pick))
            res[-1] = (tot, [(tv.span_min.copy(), tv.span_area.copy(), tv.dist.copy(), tv.conn_ref.copy(), tv) for tv in res[-1][1]], pick)
        ctx.close()
    a, b = res
    a[0].pop("result_bytes"), b[0].pop("result_bytes")
    assert a[0] == b[0] and a[0]["tiles"] == len(cfgs) and a[0]["spans"] > 0
    for x, y in zip(a[1], b[1]):
        for k in range(4):
            np.testing.assert_array_equal(x[k], y[k])
    for i, x in zip(a[2], a[1]):
        ohf = oracle.Heightfield.from_config(cfgs[i])
        ohf.builder_rasterize(cfgs[i], v, t)
        oc = oracle.CompactHeightfield.build_from_heightfield(ohf)
        oc.build_distance_field()
        e = oc.export()
        np.testing.assert_array_equal(x[0], e.span_min)
        np.testing.assert_array_equal(x[2], e.dist)


@pytest.mark.parametrize("groups", [1, 3, 8, 5000])
def test_streamed_build_equals_plain_build(groups):
    """rcb_build_tiles_streamed (D2H of finished tile groups overlapped with the next groups) returns the same bytes."""
    import recast_rs_b200.builder as rb
    from recast_rs_b200 import errors, synth

    v, t, cfgs = synth.config2(nq=240)  # 256 tiles
    ctx = _context()
    ctx.upload_mesh(v, t)
    with ctx.build_tiles(cfgs, rb.STAGE_ALL_BUILD) as a, ctx.build_tiles_streamed(cfgs, rb.STAGE_ALL_BUILD, groups=groups) as b:
        a.download()
        ta, tb = a.totals(), b.totals()
        ta.pop("result_bytes"), tb.pop("result_bytes")
        assert ta == tb
        for i in range(len(cfgs)):
            x, y = a.tile(i), b.tile(i)
            for k in ("hf_cell_offset", "span_min", "span_max", "span_area", "cell_index", "cell_count", "areas", "con", "first_conn",
                      "conn_ref", "conn_dir", "conn_next", "dist"):
                assert np.array_equal(getattr(x, k), getattr(y, k)), (i, k)
            assert (x.max_distance, x.max_height, x.walkable_span_count) == (y.max_distance, y.max_height, y.walkable_span_count)
    with ctx.build_tiles_streamed([], rb.STAGE_ALL_BUILD) as e:
        assert e.totals()["tiles"] == 0
    bad = t.copy()
    bad[-1] = 10 ** 7
    ctx.upload_mesh(v, bad)
    with pytest.raises(errors.NavMeshGeneration):
        ctx.build_tiles_streamed(cfgs, rb.STAGE_ALL_BUILD)
    ctx.close()
