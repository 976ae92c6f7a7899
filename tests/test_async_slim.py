"""The build path without host round trips (batches issued from the previous build's sizes and verified on the
device) and the slim result layout (RCB_RESULT_SLIM): both must return exactly the bytes of the plain path."""
import os

import numpy as np
import pytest

from tests.parity import assert_tile_equal

pytestmark = pytest.mark.gpu

NAMES = ("hf_cell_offset", "span_min", "span_max", "span_area", "cell_index", "cell_count", "areas", "con", "first_conn",
         "conn_ref", "conn_dir", "conn_next", "dist", "span_count", "conn_count", "walkable_span_count", "max_height",
         "max_distance", "status")


def _same(a, b, what):
    for k in NAMES:
        x, y = getattr(a, k), getattr(b, k)
        if isinstance(x, np.ndarray):
            assert x.dtype == y.dtype and np.array_equal(x, y), f"{what}: {k} differs"
        else:
            assert x == y, f"{what}: {k} differs ({x} != {y})"


def _ctx(**env):
    import recast_rs_b200.builder as rb

    old = {k: os.environ.get(k) for k in env}
    os.environ.update({k: str(v) for k, v in env.items()})
    try:
        return rb.Context(0)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def test_second_build_is_speculative_and_identical(oracle):
    import recast_rs_b200.builder as rb
    from recast_rs_b200 import synth

    v, t, cfgs = synth.config2(seed=2, nq=150, tile=64)
    ctx = _ctx(RCB_NO_SPEC=0)
    ctx.upload_mesh(v, t)
    with ctx.build_tiles(cfgs, rb.STAGE_ALL_BUILD) as r0:
        r0.download()
        first = [r0.tile(i) for i in range(len(cfgs))]
        tot0 = r0.totals()
    assert ctx.stats()["speculative"] == 0
    for rep in range(3):
        with ctx.build_tiles(cfgs, rb.STAGE_ALL_BUILD) as r:
            r.download()
            assert r.totals() == tot0 or {k: v for k, v in r.totals().items() if k != "result_bytes"} == {k: v for k, v in tot0.items() if k != "result_bytes"}
            for i in range(len(cfgs)):
                _same(r.tile(i), first[i], f"rep {rep} tile {i}")
    st = ctx.stats()
    assert st["speculative"] == 3 and st["rebuilt"] == 0 and st["cached_batches"] == 1
    for i in (0, len(cfgs) // 2, len(cfgs) - 1):
        hfa, chfa = oracle.build_tile(cfgs[i], v, t)
        assert_tile_equal(first[i], hfa, chfa, what=f"tile {i}")
    ctx.close()


def test_speculation_that_falls_short_is_rebuilt(oracle):
    """Same configurations and mesh size, but the mesh is folded: more fragments and spans per tile than the first build
    measured.  The device raises the flag, the result is rebuilt behind rcb_result_download and equals a fresh build."""
    import recast_rs_b200.builder as rb
    from recast_rs_b200 import synth

    v, t, cfgs = synth.config2(seed=2, nq=150, tile=64)
    # the same triangles folded onto half the area, the folded half lifted by 8 units: twice the spans per column there,
    # so the largest tile outgrows the shared memory the first build measured
    v2 = v.copy().reshape(-1, 3)
    mid = 0.5 * float(v2[:, 0].max())
    right = v2[:, 0] > mid
    v2[right, 0] = 2.0 * mid - v2[right, 0]
    v2[right, 1] += 8.0
    v2 = np.ascontiguousarray(v2.reshape(-1), np.float32)
    for c in cfgs:  # the y range of the tiles must hold both meshes
        c.bmax = (c.bmax[0], float(v2.reshape(-1, 3)[:, 1].max()) + 1.0, c.bmax[2])
    ctx = _ctx(RCB_NO_SPEC=0)
    ctx.upload_mesh(v, t)
    ctx.build_tiles(cfgs, rb.STAGE_ALL_BUILD).free()
    ctx.upload_mesh(v2, t)
    with ctx.build_tiles(cfgs, rb.STAGE_ALL_BUILD) as r:
        r.download()
        got = [r.tile(i) for i in range(len(cfgs))]
        tot = r.totals()
    st = ctx.stats()
    assert st["speculative"] == 1 and st["rebuilt"] == 1, st
    fresh = _ctx(RCB_NO_SPEC=1)
    fresh.upload_mesh(v2, t)
    with fresh.build_tiles(cfgs, rb.STAGE_ALL_BUILD) as r:
        r.download()
        for i in range(len(cfgs)):
            _same(got[i], r.tile(i), f"tile {i}")
        tf = r.totals()
    assert {k: x for k, x in tot.items() if k != "result_bytes"} == {k: x for k, x in tf.items() if k != "result_bytes"}
    assert fresh.stats()["speculative"] == 0
    # and the next build of the steeper mesh is speculative again, from the new sizes
    with ctx.build_tiles(cfgs, rb.STAGE_ALL_BUILD) as r:
        r.download()
        _same(r.tile(3), got[3], "after rebuild")
    assert ctx.stats() == {"speculative": 2, "rebuilt": 1, "cached_batches": 1}
    hfa, chfa = oracle.build_tile(cfgs[5], v2, t)
    assert_tile_equal(got[5], hfa, chfa, what="steeper mesh tile 5")
    ctx.close(), fresh.close()


@pytest.mark.parametrize("force_global", [0, 1])
def test_slim_layout_expands_to_the_full_layout(force_global):
    import recast_rs_b200.builder as rb
    from recast_rs_b200 import synth

    v, t, cfgs = synth.config2(seed=3, nq=130, tile=64)
    ctx = _ctx(RCB_FORCE_GLOBAL=force_global)
    ctx.upload_mesh(v, t)
    for rep in range(2):  # the second pass of the tile-resident back end is speculative
        with ctx.build_tiles(cfgs, rb.STAGE_ALL_BUILD) as full, ctx.build_tiles(cfgs, rb.STAGE_ALL_BUILD | rb.RESULT_SLIM) as slim:
            full.download(), slim.download()
            bf, bs = full.totals()["result_bytes"], slim.totals()["result_bytes"]
            assert bs < 0.27 * bf, (bs, bf)
            for i in range(len(cfgs)):
                raw = slim.tile(i, expand=False)
                assert raw.hf_cell_offset is None and raw.first_conn is None and raw.conn_next is None and raw.conn_dir is None
                assert raw.conn_mask is not None and raw.conn_off8 is not None and raw.conn_ref16 is None and raw.conn_ref is None
                assert raw.cell_index is None and raw.cell_count is None and raw.cell_count16 is not None
                assert raw.con is None and raw.con8 is None and raw.areas is None  # areas == span_area: nothing wrote it
                _same(slim.tile(i), full.tile(i), f"rep {rep} tile {i}")
            dv = slim.tile_device_view(0)  # device views keep the full arrays
            assert dv.first_conn and dv.conn_next and dv.conn_ref and dv.hf_cell_offset
    ctx.close()


def test_slim_layout_with_a_tile_of_more_than_65535_spans():
    """A tile of 100 k spans: span references would not fit 16 bits, but the slim layout ships offsets inside the neighbour
    cell, which do."""
    import recast_rs_b200.builder as rb
    from recast_rs_b200 import synth

    v, t, cfgs = synth.config1(seed=1, nq=80)  # one tile of ~320x320 cells
    ctx = _ctx()
    ctx.upload_mesh(v, t)
    with ctx.build_tiles(cfgs, rb.STAGE_ALL_BUILD) as full, ctx.build_tiles(cfgs, rb.STAGE_ALL_BUILD | rb.RESULT_SLIM) as slim:
        full.download(), slim.download()
        assert full.tile(0).span_count > 65535
        raw = slim.tile(0, expand=False)
        assert raw.conn_off8 is not None and raw.conn_ref16 is None and raw.conn_ref is None and raw.conn_mask is not None
        _same(slim.tile(0), full.tile(0), "single big tile")
    ctx.close()


def test_streamed_slim_speculative_build_matches():
    import recast_rs_b200.builder as rb
    from recast_rs_b200 import synth

    v, t, cfgs = synth.config2(seed=1, nq=240, tile=64)  # 256 tiles
    ctx = _ctx()
    ctx.upload_mesh(v, t)
    with ctx.build_tiles(cfgs, rb.STAGE_ALL_BUILD) as ref:
        ref.download()
        want = [ref.tile(i) for i in range(len(cfgs))]
    for rep in range(3):
        with ctx.build_tiles_streamed(cfgs, rb.STAGE_ALL_BUILD | rb.RESULT_SLIM, groups=4) as r:
            for i in range(len(cfgs)):
                _same(r.tile(i), want[i], f"rep {rep} tile {i}")
    st = ctx.stats()
    assert st["speculative"] >= 8 and st["rebuilt"] == 0, st
    ctx.close()


def test_result_outlives_its_context():
    """`with Context() as ctx: r = ctx.build_tiles(...)` and r collected after the block."""
    import recast_rs_b200.builder as rb
    from recast_rs_b200 import errors, synth

    v, t, cfgs = synth.config2(seed=1, nq=48, tile=64)
    with rb.Context(0) as ctx:
        ctx.upload_mesh(v, t)
        r = ctx.build_tiles(cfgs, rb.STAGE_ALL_BUILD)
        r2 = ctx.build_tiles(cfgs, rb.STAGE_ALL_BUILD)
        r2.download()
    with pytest.raises((errors.RecastError, ValueError)):
        r.totals()
    r.free()
    r2.free()
    r.free()  # idempotent
    with rb.Context(0) as ctx:  # the library is still healthy
        ctx.upload_mesh(v, t)
        with ctx.build_tiles(cfgs, rb.STAGE_ALL_BUILD) as r3:
            assert r3.totals()["tiles"] == len(cfgs)


def test_bad_index_fails_every_build_until_the_mesh_is_replaced():
    import recast_rs_b200.builder as rb
    from recast_rs_b200 import errors, synth

    v, t, cfgs = synth.config2(seed=1, nq=48, tile=64)
    ctx = _ctx()
    ctx.upload_mesh(v, t)
    ctx.build_tiles(cfgs, rb.STAGE_ALL_BUILD).free()
    bad = t.copy()
    bad[7] = -1
    ctx.upload_mesh(v, bad)
    for _ in range(2):  # also when the batch would have been issued speculatively
        with pytest.raises(errors.NavMeshGeneration):
            ctx.build_tiles(cfgs, rb.STAGE_ALL_BUILD)
    with ctx.build_tiles([], rb.STAGE_ALL_BUILD) as r:  # no tile, no rasterize_triangles call, no error
        assert r.totals()["tiles"] == 0
    ctx.upload_mesh(v, t)
    with ctx.build_tiles(cfgs, rb.STAGE_ALL_BUILD) as r:
        assert r.totals()["spans"] > 0
    ctx.close()


def test_tile_of_more_than_2_pow_24_cells(oracle):
    """fragment records of the v2 clipper carry a 24-bit cell; a 4100x4100-cell tile must take the other clipper."""
    import recast_rs_b200.builder as rb
    from recast_rs_b200.config import RecastConfig

    W = 4100
    cs = 0.25
    # a few large triangles whose cells lie beyond index 2^24 (rows z >= 4093)
    v = np.array([[10.0, 1.0, 1020.0], [900.0, 2.0, 1021.0], [400.0, 5.0, 1024.9],
                  [1000.0, 3.0, 1023.0], [1024.0, 3.5, 1024.5], [1010.0, 9.0, 1000.0],
                  [5.0, 0.5, 5.0], [9.0, 0.7, 5.5], [6.0, 2.0, 9.0]], np.float32).reshape(-1)
    t = np.arange(9, dtype=np.int32)
    cfg = RecastConfig(width=W, height=W, cs=cs, ch=0.2, bmin=(0.0, 0.0, 0.0), bmax=(W * cs, 20.0, W * cs))
    ctx = _ctx()
    ctx.upload_mesh(v, t)
    with ctx.build_tiles([cfg], rb.STAGE_RASTER | rb.STAGE_FILTER) as r:
        tv = r.download().tile(0)
    ohf = oracle.Heightfield.from_config(cfg)
    ohf.builder_rasterize(cfg, v, t)
    e = ohf.export()
    assert tv.span_count == e.span_min.size and tv.span_count > 1000
    np.testing.assert_array_equal(tv.hf_cell_offset, e.cell_offset)
    np.testing.assert_array_equal(tv.span_min, e.span_min)
    np.testing.assert_array_equal(tv.span_max, e.span_max)
    np.testing.assert_array_equal(tv.span_area, e.span_area)
    assert int(np.flatnonzero(np.diff(e.cell_offset.astype(np.int64)))[-1]) > (1 << 24)
    ctx.close()


def test_bounds_build_reports(request):
    """RCB_DEBUG_SO=1 (librecast_b200_dbg.so, -DRCB_BOUNDS): the checked accessors must bite on a deliberate violation and
    must have seen none from the real kernels so far; the normal build says it is not instrumented."""
    import recast_rs_b200.builder as rb
    from recast_rs_b200 import synth

    ctx = _ctx()
    v, t, cfgs = synth.config2(seed=1, nq=64, tile=64)
    ctx.upload_mesh(v, t)
    with ctx.build_tiles(cfgs, rb.STAGE_ALL_BUILD | rb.STAGE_REGIONS) as r:
        r.download()
    rep = ctx.bounds_report()
    if os.environ.get("RCB_DEBUG_SO") == "1":
        assert rep["instrumented"] and rep["violations"] == 0, rep
        a = (__import__("ctypes").c_uint64 * 4)()
        from recast_rs_b200 import _ffi

        assert _ffi.lib().rcb_debug_bounds_report(ctx._h, a, 2) == 1
        assert a[0] == 1 and (a[1] >> 32) == 1 and a[2] == 8 and a[3] == 8, list(a)  # file 1 (k_tile.cu), index 8 of 8
        assert ctx.bounds_report()["violations"] == 0  # cleared: the session-end verdict counts real kernels only
    else:
        assert not rep["instrumented"] and rep["violations"] == 0
    ctx.close()


def test_slim_layout_wide_fallbacks_and_written_areas(oracle):
    """Columns of 300 spans: offsets inside a cell pass 255, so the slim download falls back to span references (16 bits:
    7500 spans; 32 bits for the 17x17 grid's 86700) and the 16-bit con; and with erosion CompactHeightfield.areas differs
    from CompactSpan.area, so it travels too."""
    import recast_rs_b200.builder as rb
    from recast_rs_b200.config import RecastConfig

    W = H = 5
    n = 300
    off = np.arange(W * H + 1, dtype=np.uint32) * n
    smin = np.tile(np.arange(n, dtype=np.int16) * 4, W * H)
    smax = (smin + 1).astype(np.int16)
    area = np.ones(W * H * n, np.uint8)
    cfg = RecastConfig(width=W, height=H, cs=1.0, ch=1.0, bmin=(0.0, 0.0, 0.0), bmax=(5.0, 2000.0, 5.0))
    ctx = _ctx()
    mask = rb.STAGE_COMPACT | rb.STAGE_DIST | rb.STAGE_ERODE
    with ctx.build_from_spans([cfg], off, smin, smax, area, mask, 1) as full, \
            ctx.build_from_spans([cfg], off, smin, smax, area, mask | rb.RESULT_SLIM, 1) as slim:
        full.download(), slim.download()
        raw = slim.tile(0, expand=False)
        assert raw.conn_off8 is None and raw.con8 is None and raw.con is not None and int(raw.con.max()) > 255  # the wide arrays instead
        assert raw.cell_count16 is not None and raw.conn_ref16 is not None and raw.conn_ref is None
        assert raw.areas is not None and (raw.areas != raw.span_area).any()  # eroded
        _same(slim.tile(0), full.tile(0), "deep columns")
    ohf = oracle.Heightfield.from_csr(W, H, cfg.bmin, cfg.bmax, cfg.cs, cfg.ch, off, smin, smax, area)
    oc = oracle.CompactHeightfield.build_from_heightfield(ohf)
    oc.erode_walkable_area(1)
    oc.build_distance_field()
    e = oc.export()
    np.testing.assert_array_equal(raw.areas, e.areas)
    np.testing.assert_array_equal(raw.con, e.con)
    # 17 x 17 columns of 300 spans: 86700 spans in the tile, references need 32 bits as well
    W2 = 17
    off2 = np.arange(W2 * W2 + 1, dtype=np.uint32) * n
    smin2 = np.tile(np.arange(n, dtype=np.int16) * 4, W2 * W2)
    cfg2 = RecastConfig(width=W2, height=W2, cs=1.0, ch=1.0, bmin=(0.0, 0.0, 0.0), bmax=(17.0, 2000.0, 17.0))
    a2 = np.ones(W2 * W2 * n, np.uint8)
    m2 = rb.STAGE_COMPACT | rb.STAGE_DIST
    with ctx.build_from_spans([cfg2], off2, smin2, (smin2 + 1).astype(np.int16), a2, m2) as full, \
            ctx.build_from_spans([cfg2], off2, smin2, (smin2 + 1).astype(np.int16), a2, m2 | rb.RESULT_SLIM) as slim:
        full.download(), slim.download()
        raw = slim.tile(0, expand=False)
        assert raw.conn_off8 is None and raw.conn_ref16 is None and raw.conn_ref is not None and int(raw.conn_ref.max()) > 65535
        assert raw.areas is None
        _same(slim.tile(0), full.tile(0), "deep columns, wide references")
    ctx.close()
