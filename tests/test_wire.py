"""Heightfield wire formats of detour-dynamic: VoxelTile::serialize_spans /
deserialize_spans (io/voxel_tile.rs:70-176, big-endian, lenient), DynamicTile::reconstruct_heightfield
(dynamic_tile.rs:147-257, little-endian, strict), DynamicTileCheckpoint::clone_heightfield (checkpoint.rs:29-60).
The reference has no tests or vectors for these functions ("parity unpinned"): on the CPU the oracle is
cross-checked against the independent Python restatement; on the GPU the CUDA path is compared with the
oracle byte for byte, including truncated and garbage streams, and through round trips at full size."""
import io

import numpy as np
import pytest

from oracle import pymodel as pm
from recast_rs_b200.config import RecastConfig


def _stream(rng, cells, big, max_spans=5, bad_span_rate=0.0, lo=-20, hi=40):
    order = "big" if big else "little"
    out = bytearray()
    for _ in range(cells):
        n = int(rng.integers(0, max_spans + 1))
        out += n.to_bytes(2, order)
        for _i in range(n):
            a = int(rng.integers(lo, hi))
            b = a + int(rng.integers(0, 12))
            if rng.random() < bad_span_rate:
                a, b = b + 1, a
            ar = int(rng.integers(0, 3))
            if rng.random() < 0.1:  # values that only survive through the `as i16` / `as u8` wraps
                a, b, ar = a + 0x10000 * int(rng.integers(1, 9)), b + 0x10000 * int(rng.integers(1, 9)), ar + 256 * int(rng.integers(1, 9))
            for v in (a, b, ar):
                out += (v & 0xFFFFFFFF).to_bytes(4, order)
    return bytes(out)


def _export_eq(a, b):
    assert np.array_equal(a.cell_offset, b[0]) and np.array_equal(a.span_min, b[1])
    assert np.array_equal(a.span_max, b[2]) and np.array_equal(a.span_area, b[3])


@pytest.mark.parametrize("seed", range(6))
def test_oracle_vs_pymodel_wire(oracle, seed):
    rng = np.random.default_rng(seed)
    for it in range(40):
        w, h = int(rng.integers(1, 7)), int(rng.integers(1, 7))
        big = bool(it % 2)
        s = _stream(rng, w * h, big, bad_span_rate=0.05 if it % 3 == 0 else 0.0)
        if it % 4 == 0:
            s = s[:int(rng.integers(0, len(s) + 1))]
        if it % 7 == 0:
            s = bytes(rng.integers(0, 3, size=len(s), dtype=np.uint8))  # noise: counts run past the buffer
        ohf = oracle.Heightfield(w, h, (0, 0, 0), (1, 1, 1), 0.3, 0.2)
        phf = pm.Heightfield(w, h, (0, 0, 0), (1, 1, 1), 0.3, 0.2)
        arr = np.frombuffer(s, np.uint8)
        if big:
            ohf.deserialize_spans(arr)
            pm.voxel_deserialize_spans(phf, s)
        else:
            e1 = e2 = None
            try:
                ohf.reconstruct_heightfield(arr)
            except oracle.OracleError as e:
                e1 = e.status
            try:
                pm.dynamic_reconstruct_heightfield(phf, s)
            except pm.WireError as e:
                e2 = {"recast": -6, "navmesh": -2}[e.kind]
            assert e1 == e2
            if e1:
                continue
        _export_eq(ohf.export(), phf.export())
        for be in (True, False):
            assert ohf.serialize_spans(be).tobytes() == pm.voxel_serialize_spans(phf, be)
        _export_eq(ohf.checkpoint_clone().export(), pm.checkpoint_clone_heightfield(phf).export())


def test_oracle_wire_round_trip_and_sign_extension(oracle):
    hf = oracle.Heightfield(2, 1, (0, 0, 0), (1, 1, 1), 0.5, 0.5)
    hf.add_span(0, 0, -3, -1, 1)
    hf.add_span(0, 0, 4, 9, 0)
    hf.add_span(1, 0, 2, 2, 63)
    be = hf.serialize_spans(True).tobytes()
    # u16 count, then `min as u32` (sign-extended), `max as u32`, `area as u32`, all big-endian (voxel_tile.rs:95-111)
    assert be[:2] == b"\x00\x02" and be[2:6] == b"\xff\xff\xff\xfd" and be[6:10] == b"\xff\xff\xff\xff" and be[10:14] == b"\x00\x00\x00\x01"
    assert len(be) == 2 * 2 + 12 * 3
    back = oracle.Heightfield(2, 1, (0, 0, 0), (1, 1, 1), 0.5, 0.5)
    back.deserialize_spans(np.frombuffer(be, np.uint8))
    _export_eq(back.export(), tuple(getattr(hf.export(), k) for k in ("cell_offset", "span_min", "span_max", "span_area")))
    # the reference's own pair disagrees: its big-endian bytes fed to the little-endian reader are not a heightfield
    le_reader = oracle.Heightfield(2, 1, (0, 0, 0), (1, 1, 1), 0.5, 0.5)
    with pytest.raises(oracle.OracleError):
        le_reader.reconstruct_heightfield(np.frombuffer(be, np.uint8))  # count 0x0200 = 512 spans: short buffer


def test_voxel_tile_file_header_round_trip():
    from recast_rs_b200.wire import VoxelTile

    t = VoxelTile(3, -2, 5, 7, (0.5, -1.0, 2.0), (3.0, 4.0, 5.5), 0.5, 0.25, 0, b"\x00\x01" + bytes(12) + bytes(2 * 34))
    f = io.BytesIO()
    t.write_to(f)
    raw = f.getvalue()
    assert len(raw) == 14 * 4 + len(t.span_data)  # voxel_tile.rs:178-199: 5 i32, 8 f32, u32 length
    assert raw[:4] == (3).to_bytes(4, "big") and raw[4:8] == (-2 & 0xFFFFFFFF).to_bytes(4, "big")
    assert raw[52:56] == len(t.span_data).to_bytes(4, "big")
    f.seek(0)
    assert VoxelTile.read_from(f) == t
    with pytest.raises(EOFError):
        VoxelTile.read_from(io.BytesIO(raw[:-1]))


# ---- GPU ------------------------------------------------------------------------------------------------------
def _cfgs(rng, n):
    out = []
    for _ in range(n):
        w, h = int(rng.integers(1, 23)), int(rng.integers(1, 19))
        out.append(RecastConfig(width=w, height=h, cs=0.5, ch=0.25, bmin=(0.0, 0.0, 0.0), bmax=(w * 0.5, 8.0, h * 0.5), walkable_height=3,
                                walkable_climb=1))
    return out


def _offsets(streams):
    off = np.zeros(len(streams) + 1, np.uint64)
    off[1:] = np.cumsum([len(s) for s in streams], dtype=np.uint64)
    return off


def _ohf(oracle, c):
    return oracle.Heightfield(c.width, c.height, c.bmin, c.bmax, c.cs, c.ch)


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(4))
def test_gpu_deserialize_big_endian_lenient(oracle, seed):
    import recast_rs_b200.builder as rb
    from tests.parity import assert_tile_equal

    rng = np.random.default_rng(100 + seed)
    cfgs = _cfgs(rng, 24)
    streams = []
    for i, c in enumerate(cfgs):
        s = _stream(rng, c.width * c.height, True, max_spans=6, bad_span_rate=0.05)
        if i % 3 == 1:
            s = s[:int(rng.integers(0, len(s) + 1))]  # odd and even cut points, in counts and in spans
        if i % 8 == 5:
            s = bytes(rng.integers(0, 2, size=len(s), dtype=np.uint8))
        if i % 11 == 7:
            s = b""
        streams.append(s)
    with rb.Context(0) as ctx:
        with ctx.build_from_span_data(cfgs, b"".join(streams), _offsets(streams), rb.WIRE_VOXEL_TILE, rb.STAGE_COMPACT | rb.STAGE_DIST) as r:
            r.download()
            for i, c in enumerate(cfgs):
                ohf = _ohf(oracle, c)
                ohf.deserialize_spans(np.frombuffer(streams[i], np.uint8))
                oc = oracle.CompactHeightfield.build_from_heightfield(ohf)
                oc.build_distance_field()
                assert_tile_equal(r.tile(i), ohf.export(), oc.export(), what=f"tile {i}")


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(3))
def test_gpu_reconstruct_little_endian_strict(oracle, seed):
    import recast_rs_b200.builder as rb
    from recast_rs_b200 import errors
    from tests.parity import assert_tile_equal

    rng = np.random.default_rng(200 + seed)
    cfgs = _cfgs(rng, 16)
    streams = [_stream(rng, c.width * c.height, False, max_spans=6) for c in cfgs]
    with rb.Context(0) as ctx:
        with ctx.build_from_span_data(cfgs, b"".join(streams), _offsets(streams), rb.WIRE_DYNAMIC_TILE, 0) as r:
            r.download()
            for i, c in enumerate(cfgs):
                ohf = _ohf(oracle, c)
                ohf.reconstruct_heightfield(np.frombuffer(streams[i], np.uint8))
                assert_tile_equal(r.tile(i), ohf.export(), what=f"tile {i}")
        # every way the strict reader fails, one tile at a time so the oracle names the error
        for k in range(40):
            c = cfgs[k % len(cfgs)]
            s = _stream(rng, c.width * c.height, False, max_spans=4, bad_span_rate=0.0 if k % 2 else 0.02)
            if k % 4 < 2:
                s = s[:int(rng.integers(0, len(s) + 1))]
            want = None
            try:
                _ohf(oracle, c).reconstruct_heightfield(np.frombuffer(s, np.uint8))
            except oracle.OracleError as e:
                want = e.status
            exc = {None: None, -6: errors.Recast, -2: errors.NavMeshGeneration}[want]
            if exc is None:
                ctx.build_from_span_data([c], s, [0, len(s)], rb.WIRE_DYNAMIC_TILE, 0).free()
            else:
                with pytest.raises(exc):
                    ctx.build_from_span_data([c], s, [0, len(s)], rb.WIRE_DYNAMIC_TILE, 0)


@pytest.mark.gpu
def test_gpu_strict_reader_reports_first_failure_in_reference_order(oracle):
    """A tile whose cell 0 holds a min > max span AND whose tail is short fails in add_span (cell order);
    in a batch the first failing tile decides."""
    import recast_rs_b200.builder as rb
    from recast_rs_b200 import errors

    c = RecastConfig(width=2, height=1, cs=0.5, ch=0.5, bmin=(0, 0, 0), bmax=(1, 1, 0.5))
    le = lambda *v: b"".join((x & 0xFFFFFFFF).to_bytes(4, "little") for x in v)  # noqa: E731
    bad_span_then_short = (1).to_bytes(2, "little") + le(5, 2, 1) + (3).to_bytes(2, "little")
    short_then_bad_span = (2).to_bytes(2, "little") + le(1, 2, 1)
    good = (0).to_bytes(2, "little") * 2
    with pytest.raises(oracle.OracleError) as e:
        _ohf(oracle, c).reconstruct_heightfield(np.frombuffer(bad_span_then_short, np.uint8))
    assert e.value.status == -2
    with rb.Context(0) as ctx:
        with pytest.raises(errors.NavMeshGeneration):
            ctx.build_from_span_data([c], bad_span_then_short, [0, len(bad_span_then_short)], rb.WIRE_DYNAMIC_TILE, 0)
        with pytest.raises(errors.Recast):
            ctx.build_from_span_data([c], short_then_bad_span, [0, len(short_then_bad_span)], rb.WIRE_DYNAMIC_TILE, 0)
        ss = [good, short_then_bad_span, bad_span_then_short]
        with pytest.raises(errors.Recast, match="tile 1"):
            ctx.build_from_span_data([c] * 3, b"".join(ss), _offsets(ss), rb.WIRE_DYNAMIC_TILE, 0)
        # the lenient reader takes all of them
        ctx.build_from_span_data([c] * 3, b"".join(ss), _offsets(ss), rb.WIRE_VOXEL_TILE, 0).free()


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(3))
def test_gpu_serialize_and_classes(oracle, seed):
    """Rasterized tiles -> VoxelTile bytes (both byte orders) against the oracle, then back through the
    reference-shaped classes."""
    import recast_rs_b200.builder as rb
    from recast_rs_b200 import synth, wire
    from tests.parity import random_soup

    rng = np.random.default_rng(300 + seed)
    base = RecastConfig(cs=0.5, ch=0.25, walkable_height=3, walkable_climb=1)
    v, idx = random_soup(rng, 300, 0.0, 12.0, -1.0, 5.0, small=0.5)
    cfgs = synth.tile_grid(v, 9, base)
    with rb.Context(0) as ctx:
        ctx.upload_mesh(v, idx)
        with ctx.build_tiles(cfgs, rb.STAGE_RASTER | rb.STAGE_FILTER) as r:
            be, off = r.serialize_spans(rb.WIRE_VOXEL_TILE)
            le, off2 = r.serialize_spans(rb.WIRE_DYNAMIC_TILE)
            assert np.array_equal(off, off2) and int(off[-1]) == r.span_data_size() == be.size
            tiles = wire.voxel_tiles_from_result(r, cfgs)
        ohfs = []
        for i, c in enumerate(cfgs):
            ohf = oracle.Heightfield.from_config(c)
            ohf.builder_rasterize(c, v, idx)
            ohfs.append(ohf)
            assert be[int(off[i]):int(off[i + 1])].tobytes() == ohf.serialize_spans(True).tobytes(), i
            assert le[int(off[i]):int(off[i + 1])].tobytes() == ohf.serialize_spans(False).tobytes(), i
            assert tiles[i].span_data == ohf.serialize_spans(True).tobytes()
        # reference-shaped round trip, one tile
        t = tiles[len(tiles) // 2]
        e = ohfs[len(tiles) // 2].export()
        hf = t.heightfield(ctx)
        for a, b in ((hf.cell_offset, e.cell_offset), (hf.span_min, e.span_min), (hf.span_max, e.span_max), (hf.span_area, e.span_area)):
            np.testing.assert_array_equal(a, b)
        assert wire.VoxelTile.from_heightfield(t.tile_x, t.tile_z, hf, ctx).span_data == t.span_data
        # the little-endian twin is what DynamicTile::reconstruct_heightfield reads
        t_le = wire.VoxelTile(**{**t.__dict__, "span_data": le[int(off[len(tiles) // 2]):int(off[len(tiles) // 2 + 1])].tobytes()})
        hf2 = wire.DynamicTile.reconstruct_heightfield(t_le, ctx)
        np.testing.assert_array_equal(hf2.span_min, e.span_min)
        np.testing.assert_array_equal(hf2.span_area, e.span_area)


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(3))
def test_gpu_checkpoint_clone(oracle, seed):
    """clone_heightfield re-adds spans with Heightfield::add_span: columns that are not separated change."""
    import recast_rs_b200.builder as rb
    from recast_rs_b200 import wire

    rng = np.random.default_rng(400 + seed)
    w, h = 13, 9
    off, mn, mx, ar = [0], [], [], []
    for _ in range(w * h):
        z = int(rng.integers(-5, 5))
        for _k in range(int(rng.integers(0, 6))):
            top = z + int(rng.integers(0, 5))
            mn.append(z), mx.append(top), ar.append(int(rng.integers(0, 2)))
            z = top + int(rng.integers(-1, 3))  # touching / overlapping / separated followers
        off.append(len(mn))
    ohf = oracle.Heightfield.from_csr(w, h, (0, 0, 0), (w * 0.5, 4, h * 0.5), 0.5, 0.5, off, mn, mx, ar)
    want = ohf.checkpoint_clone().export()
    assert want.span_min.size < len(mn)  # something merged
    with rb.Context(0) as ctx:
        hf = rb.Heightfield(w, h, (0, 0, 0), (w * 0.5, 4, h * 0.5), 0.5, 0.5, off, mn, mx, ar, ctx=ctx)
        cp = wire.DynamicTileCheckpoint.new(hf, {1, 2})
        got = cp.heightfield
        np.testing.assert_array_equal(got.cell_offset, want.cell_offset)
        np.testing.assert_array_equal(got.span_min, want.span_min)
        np.testing.assert_array_equal(got.span_max, want.span_max)
        np.testing.assert_array_equal(got.span_area, want.span_area)
        assert cp.is_valid_for({2}) and not cp.is_valid_for({3})


@pytest.mark.gpu
def test_gpu_wire_round_trip_full_size():
    """Config 2 (2025 tiles, 8M spans): build -> serialize -> parse -> later stages equals the direct build."""
    import recast_rs_b200.builder as rb
    from recast_rs_b200 import synth

    v, idx, cfgs = synth.config2()
    with rb.Context(0) as ctx:
        ctx.upload_mesh(v, idx)
        with ctx.build_tiles(cfgs, rb.STAGE_ALL_BUILD) as r:
            data, off = r.serialize_spans(rb.WIRE_VOXEL_TILE)
            data_le, _ = r.serialize_spans(rb.WIRE_DYNAMIC_TILE)
            r.download()
            tot = r.totals()
            ref = [r.tile(i) for i in (0, 777, 2024)]
            ref = [(t.span_min.copy(), t.span_area.copy(), t.dist.copy(), t.con.copy(), t.conn_ref.copy()) for t in ref]
        assert data.size == 2 * tot["cells"] + 12 * tot["spans"]
        for fmt, d in ((rb.WIRE_VOXEL_TILE, data), (rb.WIRE_DYNAMIC_TILE, data_le)):
            # the serialized areas are post-filter: no filters on the way back
            with ctx.build_from_span_data(cfgs, d, off, fmt, rb.STAGE_COMPACT | rb.STAGE_DIST) as r2:
                t2 = r2.totals()
                assert (t2["spans"], t2["connections"], t2["walkable_spans"]) == (tot["spans"], tot["connections"], tot["walkable_spans"])
                d2, off2 = r2.serialize_spans(fmt)
                assert np.array_equal(off, off2) and np.array_equal(d, d2)
                r2.download()
                for k, i in enumerate((0, 777, 2024)):
                    t = r2.tile(i)
                    for a, b in zip((t.span_min, t.span_area, t.dist, t.con, t.conn_ref), ref[k]):
                        np.testing.assert_array_equal(a, b)
