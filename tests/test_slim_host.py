"""The binding layer's half of RCB_RESULT_SLIM on the CPU: `TileVoxels.expand_slim` (recast_rs_b200/builder.py) must rebuild
hf_cell_offset, cell_index, first_conn, conn_next, conn_dir, conn_ref, con and areas from what a slim download carries.
Here the slim arrays are derived with numpy from the ORACLE's full compact heightfield (a host restatement of k_slim_pack:
conn_mask = directions present, conn_off8 = reference minus the neighbour cell's first span, counts in 16 bits), so the
rebuild rule - compact_heightfield.rs:375-427: connections in span order, inside a span in descending direction, `next` ->
the one pushed before, `first_connection` = the last pushed - is held to the reference's layout without a GPU.  The device
side of the same contract is tests/test_async_slim.py."""
import numpy as np
import pytest

from tests.parity import random_soup

NONE = 0xFFFFFFFF
DX = np.array([-1, 0, 1, 1, 1, 0, -1, -1], np.int64)  # compact_heightfield.rs:109-111
DZ = np.array([-1, -1, -1, 0, 1, 1, 1, 0], np.int64)


def _slim_from_oracle(cfg, hfa, chfa, wide=None):
    """What a slim host view of this tile holds (numpy twin of k_slim_pack); `wide` forces one of the fallbacks."""
    from recast_rs_b200.builder import TileVoxels

    S, K, W = chfa.span_min.size, chfa.conn_ref.size, cfg.width
    cnt = chfa.cell_count.astype(np.int64)
    span_cell = np.repeat(np.arange(cnt.size, dtype=np.int64), cnt)
    # owner span of every connection: follow first_conn / conn_next
    owner = np.full(K, -1, np.int64)
    for s in range(S):
        k = int(chfa.first_conn[s])
        while k != NONE:
            owner[k] = s
            k = int(chfa.conn_next[k])
    assert (owner >= 0).all()
    mask = np.zeros(S, np.uint8)
    np.bitwise_or.at(mask, owner, (1 << chfa.conn_dir.astype(np.int64)).astype(np.uint8))
    d = chfa.conn_dir.astype(np.int64)
    ncell = span_cell[owner] + DZ[d] * W + DX[d]
    off = chfa.conn_ref.astype(np.int64) - chfa.cell_index[ncell].astype(np.int64)
    assert (off >= 0).all() and (off < cnt[ncell]).all()
    tv = TileVoxels(cfg.width, cfg.height, S, K, chfa.walkable_span_count, chfa.max_height, chfa.max_distance, 0, None,
                    chfa.span_min.copy(), chfa.span_max.copy(), chfa.span_area.copy())
    tv.dist = chfa.dist.copy()
    tv.conn_mask = mask
    tv.areas = None if np.array_equal(chfa.areas, chfa.span_area) else chfa.areas.copy()
    if wide == "cnt32":
        tv.cell_count = chfa.cell_count.astype(np.uint32)
    else:
        tv.cell_count16 = chfa.cell_count.astype(np.uint16)
    if wide in ("ref16", "ref32"):  # an offset passed 255: references and con travel in the narrowest form that holds them
        if wide == "ref16":
            tv.conn_ref16, tv.con8 = chfa.conn_ref.astype(np.uint16), chfa.con.astype(np.uint8)
        else:
            tv.conn_ref, tv.con = chfa.conn_ref.astype(np.uint32), chfa.con.astype(np.uint16)
    else:
        assert off.max(initial=0) <= 255
        tv.conn_off8 = off.astype(np.uint8)
    return tv


def _check(tv, hfa, chfa, what):
    tv.expand_slim()
    for name, want in (("hf_cell_offset", hfa.cell_offset), ("cell_index", chfa.cell_index), ("cell_count", chfa.cell_count),
                       ("areas", chfa.areas), ("con", chfa.con), ("first_conn", chfa.first_conn), ("conn_ref", chfa.conn_ref),
                       ("conn_dir", chfa.conn_dir), ("conn_next", chfa.conn_next)):
        got = np.asarray(getattr(tv, name))
        assert got.shape == np.asarray(want).shape and np.array_equal(got.astype(np.int64), np.asarray(want).astype(np.int64)), f"{what}: {name}"


@pytest.mark.parametrize("wide", [None, "cnt32", "ref16", "ref32"])
def test_expand_slim_rebuilds_the_reference_layout(oracle, wide):
    from recast_rs_b200 import synth

    v, t, cfgs = synth.config3(seed=4, nq=60, tile=32, lattice=5)  # terrain with buildings: two and three spans per column
    picked, second_level = 0, False
    for i, cfg in enumerate(cfgs):
        hfa, chfa = oracle.build_tile(cfg, v, t)
        if chfa.conn_ref.size == 0:
            continue
        tv = _slim_from_oracle(cfg, hfa, chfa, wide)
        second_level = second_level or (chfa.cell_count.max() >= 2 and (tv.conn_off8 is None or bool(tv.conn_off8.any())))
        _check(tv, hfa, chfa, f"tile {i} wide={wide}")
        picked += 1
        if picked >= 4 and second_level:
            break
    assert picked >= 2 and second_level  # some connection leads to a span that is not its cell's first


def test_expand_slim_on_soups_and_empty_tiles(oracle):
    from recast_rs_b200.config import RecastConfig

    rng = np.random.default_rng(11)
    v, t = random_soup(rng, 300, 0.0, 10.0, 0.0, 5.0)
    for cfg in (RecastConfig(width=33, height=21, cs=0.3, ch=0.2, bmin=(0.0, -0.5, 0.0), bmax=(9.9, 6.0, 6.3), walkable_height=3, walkable_climb=2),
                RecastConfig(width=7, height=40, cs=0.25, ch=0.1, bmin=(3.0, 0.0, 0.0), bmax=(4.75, 5.5, 10.0)),
                RecastConfig(width=5, height=5, cs=0.5, ch=0.2, bmin=(100.0, 0.0, 100.0), bmax=(102.5, 5.0, 102.5))):  # nothing there
        hfa, chfa = oracle.build_tile(cfg, v, t)
        _check(_slim_from_oracle(cfg, hfa, chfa), hfa, chfa, f"soup {cfg.width}x{cfg.height}")
