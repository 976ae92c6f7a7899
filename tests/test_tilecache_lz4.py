"""Tile-cache front half: TileCache::decompress_tile (= lz4_flex::decompress_size_prepended,
tile_cache.rs:854-872) + TileCacheLayer::from_bytes (tile_cache_data.rs:128-232, 281-320), batched on the GPU and
feeding the layer pipeline of config 5.

Pinning: lz4_flex is a registry dependency that is not vendored in the reference tree.  Valid streams have exactly
one decoding, so the oracle's decoder is pinned against liblz4 (pyarrow's "lz4_raw" codec compresses the test
data and decodes it back); what lz4_flex does with MALFORMED streams is restated from its source and only
cross-checked between the C++ oracle and the Python restatement ("parity unpinned" for those)."""
import numpy as np
import pytest

from oracle import pymodel as pm
from recast_rs_b200 import tilecache as tc
from tests.test_tilecache import _assert_tile, _random_case

pa = pytest.importorskip("pyarrow")


def _compress(data: bytes) -> bytes:
    """lz4_flex::compress_prepend_size's format, with liblz4 as the compressor."""
    return len(data).to_bytes(4, "little") + pa.Codec("lz4_raw").compress(data, asbytes=True)


def _payloads(rng, count, maxlen=6000):
    out = []
    for it in range(count):
        n = int(rng.integers(0, maxlen))
        k = it % 5
        if k == 0:
            d = rng.integers(0, 256, n, dtype=np.uint8).tobytes()  # incompressible: literal runs > 15, > 270
        elif k == 1:
            d = rng.integers(0, 3, n, dtype=np.uint8).tobytes()  # short matches
        elif k == 2:
            d = (rng.integers(0, 4, max(1, n // 50), dtype=np.uint8).tobytes() * 60)[:n]  # long periodic matches
        elif k == 3:
            d = bytes(n)  # one run: offset 1, match length in the thousands
        else:
            d = b"".join(rng.integers(0, 256, int(rng.integers(1, 40)), dtype=np.uint8).tobytes() * int(rng.integers(1, 30))
                         for _ in range(int(rng.integers(1, 12))))[:n]
        out.append(d)
    return out


def _corrupt(rng, z: bytes) -> bytes:
    zz = bytearray(z)
    if len(zz) > 6:
        for _ in range(int(rng.integers(1, 4))):
            zz[int(rng.integers(0, len(zz)))] = int(rng.integers(0, 256))
        if rng.random() < 0.4:
            zz = zz[:int(rng.integers(0, len(zz)))]
        if rng.random() < 0.2:  # size prefix smaller / larger than the block decodes to
            zz[0:4] = int(rng.integers(0, 8000)).to_bytes(4, "little")
    return bytes(zz)


def _oracle_dec(oracle, z):
    try:
        return oracle.decompress_tile(z)
    except oracle.OracleError as e:
        return e.status


def test_lz4_known_answers(oracle):
    # literal-only block; the last sequence carries no match
    assert oracle.decompress_tile(b"\x05\x00\x00\x00" + b"\x50hello") == b"hello"
    # run: 1 literal 'a', then offset 1, match length 4 + 15 + 3 = 22 -> 23 x 'a'; the block must END with a
    # literal-only sequence (here an empty one), a match as the last thing is "expected another byte"
    z = b"\x17\x00\x00\x00" + b"\x1fa\x01\x00\x03" + b"\x00"
    assert oracle.decompress_tile(z) == b"a" * 23 == pm.decompress_tile(z)
    # period-3 overlap: 'abc' then offset 3, length 4+2 -> 'abcabcabc'
    z = b"\x09\x00\x00\x00" + b"\x32abc\x03\x00" + b"\x00"
    assert oracle.decompress_tile(z) == b"abcabcabc" == pm.decompress_tile(z)
    assert _oracle_dec(oracle, z[:-1]) == -7
    # empty input is Ok(empty) (tile_cache.rs:860-862); fewer than 4 bytes, missing token, bad offset are errors
    assert oracle.decompress_tile(b"") == b"" == pm.decompress_tile(b"")
    for bad in (b"\x01\x02", b"\x00\x00\x00\x00", b"\x09\x00\x00\x00" + b"\x32abc\x04\x00\x00", b"\x02\x00\x00\x00" + b"\x50hello"):
        assert _oracle_dec(oracle, bad) == -7
        with pytest.raises(pm.Lz4Error):
            pm.decompress_tile(bad)
    # a block that ends early yields fewer bytes than the prefix says (lz4_flex truncates, no error)
    assert oracle.decompress_tile(b"\x40\x00\x00\x00" + b"\x50hello") == b"hello"
    # offset 0 re-reads the zero fill
    z = b"\x08\x00\x00\x00" + b"\x10x\x00\x00" + b"\x30abc"
    assert oracle.decompress_tile(z) == b"x" + bytes(4) + b"abc" == pm.decompress_tile(z)


@pytest.mark.parametrize("seed", range(4))
def test_lz4_oracle_pinned_against_liblz4_and_pymodel(oracle, seed):
    rng = np.random.default_rng(seed)
    for d in _payloads(rng, 60):
        z = _compress(d)
        assert pa.Codec("lz4_raw").decompress(z[4:], decompressed_size=len(d), asbytes=True) == d
        assert oracle.decompress_tile(z) == d
        assert pm.decompress_tile(z) == d
        bad = _corrupt(rng, z)
        try:
            want = pm.decompress_tile(bad)
        except pm.Lz4Error:
            want = -7
        assert _oracle_dec(oracle, bad) == want


def test_layer_from_bytes_oracle_pymodel_and_host_mirror(oracle):
    layers, _, _, _ = _random_case(3, nlayers=6)
    rng = np.random.default_rng(0)
    for l in layers:
        cons = rng.integers(0, 16, l.width * l.height, dtype=np.uint8)
        b = tc.layer_to_bytes(l, cons, tx=3, ty=-4, tlayer=1)
        assert len(b) == 54 + 12 + 3 * l.width * l.height  # tile_cache_data.rs:261-278
        o, p, h = oracle.tilecache_layer_from_bytes(b), pm.tilecache_layer_from_bytes(b), tc.layer_from_bytes(b)
        for x in (o, h):
            assert (x.width, x.height, x.hmin, x.hmax) == (l.width, l.height, l.hmin, l.hmax) == (p["width"], p["height"], p["hmin"], p["hmax"])
            assert np.array_equal(x.regons, l.regons) and np.array_equal(x.areas, l.areas)
            assert np.allclose(x.bmin, np.float32(l.bmin)) and np.allclose(x.bmax, np.float32(l.bmax))
        assert p["regons"] == l.regons.tobytes() and p["cons"] == cons.tobytes()
        for bad in (b[:65], b"XILE" + b[4:], b[:4] + b"\x02\x00\x00\x00" + b[8:], b[:-1], b[:54] + (10 ** 6).to_bytes(4, "little") + b[58:]):
            with pytest.raises(oracle.OracleError) as e:
                oracle.tilecache_layer_from_bytes(bad)
            assert e.value.status == -8
            with pytest.raises(ValueError):
                pm.tilecache_layer_from_bytes(bad)
            with pytest.raises(ValueError):
                tc.layer_from_bytes(bad)


# ---- GPU --------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(3))
def test_gpu_decompress_tiles(oracle, seed):
    import recast_rs_b200.builder as rb
    from recast_rs_b200 import errors

    rng = np.random.default_rng(50 + seed)
    data = _payloads(rng, 80, maxlen=20000) + [b"", b"x"]
    blobs = [_compress(d) for d in data]
    blobs[-2] = b""  # empty compressed tile -> empty result
    with rb.Context(0) as ctx:
        got = ctx.decompress_tiles(blobs)
        assert len(got) == len(data)
        for i, (g, d) in enumerate(zip(got, data)):
            assert g == d, i
        # malformed streams one at a time: same verdict and same bytes as the oracle
        nbad = 0
        for z in blobs[:60]:
            bad = _corrupt(rng, z)
            want = _oracle_dec(oracle, bad)
            if want == -7:
                nbad += 1
                with pytest.raises(errors.DetourFailure):
                    ctx.decompress_tiles([bad])
            else:
                assert ctx.decompress_tiles([bad])[0] == want
        assert nbad > 5
        # in a batch the first failing tile is named
        with pytest.raises(errors.DetourFailure, match="tile 2"):
            ctx.decompress_tiles([blobs[0], blobs[1], b"\x01\x02", b"\x00\x00\x00\x00"])
        assert ctx.decompress_tiles([]) == []


def _compressed_case(seed, nlayers):
    layers, obstacles, cs, ch = _random_case(seed, nlayers=nlayers)
    rng = np.random.default_rng(seed)
    blobs = [_compress(tc.layer_to_bytes(l, rng.integers(0, 16, l.width * l.height, dtype=np.uint8), tx=i, ty=2 * i)) for i, l in enumerate(layers)]
    ranges = [(l.first_obstacle, l.obstacle_count) for l in layers]
    return layers, obstacles, cs, ch, blobs, ranges


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(4))
def test_gpu_tilecache_compressed_random(oracle, seed):
    import recast_rs_b200.builder as rb

    layers, obstacles, cs, ch, blobs, ranges = _compressed_case(200 + seed, 14)
    want = oracle.tilecache_build_compressed(blobs, obstacles, ranges, cs, ch)
    direct = oracle.tilecache_build(layers, obstacles, cs, ch)
    with rb.Context(0) as ctx:
        with ctx.build_tilecache_compressed(blobs, obstacles, ranges, cs, ch, rb.STAGE_COMPACT | rb.STAGE_DIST) as r:
            r.download()
            for i, a in enumerate(want):
                assert not isinstance(a, int)
                assert np.array_equal(a.span_area, direct[i].span_area)  # the compressed route changes nothing
                tv = r.tile(i)
                for k in ("span_min", "span_max", "span_area", "cell_index", "cell_count", "areas", "con", "first_conn", "conn_ref",
                          "conn_dir", "conn_next"):
                    assert np.array_equal(getattr(tv, k), getattr(a, k)), (i, k)
        # no obstacles at all
        with ctx.build_tilecache_compressed(blobs, [], None, cs, ch) as r:
            r.download()
            want0 = oracle.tilecache_build_compressed(blobs, [], None, cs, ch)
            for i, a in enumerate(want0):
                _assert_tile(r.tile(i), a, f"layer {i}")


@pytest.mark.gpu
def test_gpu_tilecache_compressed_errors_in_reference_order(oracle):
    import recast_rs_b200.builder as rb
    from recast_rs_b200 import errors

    ok = tc.TileCacheLayer((0, 0, 0), (1, 1, 1), 10, 11, 2, 2, np.ones(4, np.uint8), np.full(4, 63, np.uint8))
    good = _compress(tc.layer_to_bytes(ok))
    raw = tc.layer_to_bytes(ok)
    cases = {
        "lz4": (b"\x09\x00\x00\x00" + b"\x32abc\x04\x00\x00", errors.DetourFailure, -7),
        "prefix": (b"\x01\x02", errors.DetourFailure, -7),
        "empty": (b"", errors.DetourInvalidParam, -8),  # Ok(empty) then from_bytes: len < 66
        "magic": (_compress(b"XILE" + raw[4:]), errors.DetourInvalidParam, -8),
        "lengths": (_compress(raw[:-1]), errors.DetourInvalidParam, -8),
        "regons_len": (_compress(tc.layer_to_bytes(tc.TileCacheLayer((0, 0, 0), (1, 1, 1), 10, 11, 2, 2, np.ones(3, np.uint8), np.ones(4, np.uint8)))),
                       errors.InvalidMesh, -1),
        "hmin": (_compress(tc.layer_to_bytes(tc.TileCacheLayer((0, 0, 0), (1, 1, 1), 32767, 0, 2, 2, np.ones(4, np.uint8), np.ones(4, np.uint8)))),
                 errors.NavMeshGeneration, -2),
    }
    with rb.Context(0) as ctx:
        for name, (blob, exc, code) in cases.items():
            assert oracle.tilecache_build_compressed([blob], [], None, 0.5, 0.5) == [code], name
            with pytest.raises(exc):
                ctx.build_tilecache_compressed([blob], [], None, 0.5, 0.5)
        # the first failing layer decides, whatever fails later
        with pytest.raises(errors.InvalidMesh, match="layer 1"):
            ctx.build_tilecache_compressed([good, cases["regons_len"][0], cases["lz4"][0]], [], None, 0.5, 0.5)
        with pytest.raises(errors.DetourFailure, match="layer 0"):
            ctx.build_tilecache_compressed([cases["lz4"][0], cases["hmin"][0]], [], None, 0.5, 0.5)
        # hmin == 32767 with no non-empty cell never calls add_span
        empty = tc.TileCacheLayer((0, 0, 0), (1, 1, 1), 32767, 0, 2, 2, np.zeros(4, np.uint8), np.ones(4, np.uint8))
        ctx.build_tilecache_compressed([_compress(tc.layer_to_bytes(empty))], [], None, 0.5, 0.5).free()


@pytest.mark.gpu
def test_gpu_tilecache_compressed_config5_matches_packed_route():
    """256 layers of 48x48 cells with varied content: the compressed route and the packed route give the same bytes."""
    import recast_rs_b200.builder as rb

    layers, obstacles = tc.config5_varied(256, 48)
    blobs = [_compress(tc.layer_to_bytes(l, l.cons)) for l in layers]
    ranges = [(l.first_obstacle, l.obstacle_count) for l in layers]
    assert sum(len(b) for b in blobs) < 0.7 * sum(66 + 3 * 48 * 48 for _ in layers)  # the data does compress
    with rb.Context(0) as ctx:
        with ctx.build_tilecache_layers(layers, obstacles, 0.3, 0.2) as a, ctx.build_tilecache_compressed(blobs, obstacles, ranges, 0.3, 0.2) as b:
            a.download(), b.download()
            assert a.totals() == b.totals()
            for i in range(256):
                x, y = a.tile(i), b.tile(i)
                for k in ("span_min", "span_area", "areas", "con", "conn_ref", "conn_dir"):
                    assert np.array_equal(getattr(x, k), getattr(y, k)), (i, k)


@pytest.mark.gpu
def test_gpu_decompress_huge_size_prefix_allocates_by_the_input(oracle):
    """a 4-byte size prefix is untrusted.  A valid small block behind a prefix of 256 MiB decodes to the block's bytes
    (lz4_flex allocates the prefix and returns the bytes produced); the device scratch must follow LZ4's 255x ceiling on the
    INPUT, not the prefix - 64 such blobs would otherwise ask for 16 GiB of device memory (and be refused)."""
    import recast_rs_b200.builder as rb

    rng = np.random.default_rng(7)
    payload = bytes(rng.integers(0, 4, 3000, dtype=np.uint8))
    z = _compress(payload)
    huge = (1 << 28).to_bytes(4, "little") + z[4:]
    want = _oracle_dec(oracle, huge)
    assert want == payload  # the oracle (lz4_flex's rule): fewer bytes than the prefix says, no error
    with rb.Context(0) as ctx:
        got = ctx.decompress_tiles([huge] * 64 + [z])
        assert all(g == payload for g in got)
