"""Tiles cut into bands of rows (k_tile.cu): every cut must give the bytes of the uncut tile.  RCB_BAND_CELLS forces
small bands onto small tiles, so the hand-over between bands (halo rows, forward sweep, blur, connection bases) is
exercised on inputs the oracle finishes in seconds; the full-size 128x128-cell case is tests/test_gpu_fullsize.py."""
import os

import numpy as np
import pytest

from tests.parity import assert_tile_equal, random_soup

pytestmark = pytest.mark.gpu

NAMES = ("hf_cell_offset", "span_min", "span_max", "span_area", "cell_index", "cell_count", "areas", "con", "first_conn",
         "conn_ref", "conn_dir", "conn_next", "dist", "span_count", "conn_count", "walkable_span_count", "max_height",
         "max_distance", "status")


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


def _same(a, b, what):
    for k in NAMES:
        x, y = getattr(a, k), getattr(b, k)
        if isinstance(x, np.ndarray):
            assert x.dtype == y.dtype and x.shape == y.shape, f"{what}: {k} shape"
            if not np.array_equal(x, y):
                bad = np.flatnonzero(x != y)
                raise AssertionError(f"{what}: {k} differs at {bad[:8].tolist()} ({bad.size}): {x[bad[:8]].tolist()} != {y[bad[:8]].tolist()}")
        else:
            assert x == y, f"{what}: {k} differs ({x} != {y})"


@pytest.mark.parametrize("band_cells", [128, 200, 330, 1100, 2048])
def test_terrain_tiles_cut_into_bands(oracle, band_cells):
    """64x64-cell terrain tiles (and the partial tiles of the last row / column) in bands of 2, 3, 5, 17 and 32 rows."""
    import recast_rs_b200.builder as rb
    from recast_rs_b200 import synth

    v, t, cfgs = synth.config2(seed=4, nq=100, tile=64)
    whole, cut = _ctx(RCB_BAND_CELLS=1 << 20), _ctx(RCB_BAND_CELLS=band_cells)
    for c in (whole, cut):
        c.upload_mesh(v, t)
    for rep in range(2):  # the second build of the cut context is issued speculatively
        with whole.build_tiles(cfgs, rb.STAGE_ALL_BUILD) as r1, cut.build_tiles(cfgs, rb.STAGE_ALL_BUILD) as r2:
            r1.download(), r2.download()
            a, b = r1.totals(), r2.totals()
            a.pop("result_bytes"), b.pop("result_bytes")
            assert a == b
            for i in range(len(cfgs)):
                _same(r2.tile(i), r1.tile(i), f"band_cells {band_cells} rep {rep} tile {i}")
            if rep == 0:
                for i in (0, len(cfgs) // 2, len(cfgs) - 1):
                    hfa, chfa = oracle.build_tile(cfgs[i], v, t)
                    assert_tile_equal(r2.tile(i), hfa, chfa, what=f"tile {i}")
    assert cut.stats()["speculative"] == 1 and cut.stats()["rebuilt"] == 0
    whole.close(), cut.close()


@pytest.mark.parametrize("band_cells,seed", [(96, 1), (150, 2), (400, 3)])
def test_random_soup_cut_into_bands(oracle, band_cells, seed):
    """Triangle soups (walls, slivers, degenerate and out-of-bounds triangles, several spans per column) on odd tile
    shapes: the ledge filter, the linking and the blur all reach across band borders."""
    import recast_rs_b200.builder as rb
    from recast_rs_b200.config import RecastConfig

    rng = np.random.default_rng(seed)
    v, t = random_soup(rng, 700, 0.0, 14.0, 0.0, 6.0)
    cfgs = [RecastConfig(width=47, height=39, cs=0.3, ch=0.2, bmin=(0.0, -0.5, 0.0), bmax=(14.1, 7.0, 11.7), walkable_height=3, walkable_climb=2),
            RecastConfig(width=30, height=61, cs=0.25, ch=0.15, bmin=(2.0, 0.0, -1.0), bmax=(9.5, 6.5, 14.25)),
            RecastConfig(width=5, height=40, cs=0.5, ch=0.2, bmin=(6.0, 0.0, 0.0), bmax=(8.5, 6.5, 20.0))]
    cut = _ctx(RCB_BAND_CELLS=band_cells)
    cut.upload_mesh(v, t)
    with cut.build_tiles(cfgs, rb.STAGE_ALL_BUILD) as r:
        r.download()
        for i, c in enumerate(cfgs):
            hfa, chfa = oracle.build_tile(c, v, t)
            assert_tile_equal(r.tile(i), hfa, chfa, what=f"soup tile {i} band_cells {band_cells}")
    cut.close()


def test_open_world_128_tiles_are_cut_by_default(oracle):
    """128x128-cell tiles take the tile-resident back end in four bands of 32 rows (BASELINE config 3 at 1/64 size)."""
    import recast_rs_b200.builder as rb
    from recast_rs_b200 import synth

    v, t, cfgs = synth.config3(seed=2, nq=275, lattice=20)
    ct, cg = _ctx(RCB_FORCE_GLOBAL=0, RCB_BAND_CELLS=4608), _ctx(RCB_FORCE_GLOBAL=1)
    for c in (ct, cg):
        c.upload_mesh(v, t)
    ct.profile_enable(True)
    with ct.build_tiles(cfgs, rb.STAGE_ALL_BUILD) as r1, cg.build_tiles(cfgs, rb.STAGE_ALL_BUILD) as r2:
        r1.download(), r2.download()
        for i in range(len(cfgs)):
            _same(r1.tile(i), r2.tile(i), f"tile {i}")
        for i in (0, len(cfgs) // 2 + 1, len(cfgs) - 1):
            hfa, chfa = oracle.build_tile(cfgs[i], v, t)
            assert_tile_equal(r1.tile(i), hfa, chfa, what=f"tile {i}")
    kern = ct.profile_read()
    assert "k_tile_compact" in kern and "k_link_spans2" not in kern, sorted(kern)  # the tile-resident kernels ran
    ct.close(), cg.close()


def test_erosion_keeps_tiles_whole(oracle):
    """ERODE is a second row-sequential sweep that bands do not hand over: the batch must still be right."""
    import recast_rs_b200.builder as rb
    from recast_rs_b200 import synth

    v, t, cfgs = synth.config2(seed=5, nq=60, tile=64)
    cut = _ctx(RCB_BAND_CELLS=256)
    cut.upload_mesh(v, t)
    with cut.build_tiles(cfgs, rb.STAGE_ALL_BUILD | rb.STAGE_ERODE, 2) as r:
        r.download()
        for i in (0, len(cfgs) - 1):
            hfa, chfa = oracle.build_tile(cfgs[i], v, t, rb.STAGE_ALL_BUILD | rb.STAGE_ERODE, 2)
            assert_tile_equal(r.tile(i), hfa, chfa, what=f"eroded tile {i}")
    cut.close()


def test_regions_on_banded_tiles(oracle):
    import recast_rs_b200.builder as rb
    from recast_rs_b200 import synth

    v, t, cfgs = synth.config2(seed=6, nq=60, tile=64)
    whole, cut = _ctx(RCB_BAND_CELLS=1 << 20), _ctx(RCB_BAND_CELLS=512)
    for c in (whole, cut):
        c.upload_mesh(v, t)
    with whole.build_tiles(cfgs, rb.STAGE_ALL_BUILD | rb.STAGE_REGIONS) as r1, cut.build_tiles(cfgs, rb.STAGE_ALL_BUILD | rb.STAGE_REGIONS) as r2:
        r1.download(), r2.download()
        for i in range(len(cfgs)):
            a, b = r1.tile(i), r2.tile(i)
            _same(b, a, f"tile {i}")
            assert np.array_equal(a.reg, b.reg) and a.max_regions == b.max_regions
    whole.close(), cut.close()
