"""rcb_multi: several devices of one box behind one handle (tile groups dealt from a shared queue).  With one device it
is the plain build cut into groups; with two or more every tile must still be the bytes of the single-device build."""
import numpy as np
import pytest

from tests.parity import assert_tile_equal

pytestmark = pytest.mark.gpu

NAMES = ("hf_cell_offset", "span_min", "span_max", "span_area", "cell_index", "cell_count", "areas", "con", "first_conn",
         "conn_ref", "conn_dir", "conn_next", "dist", "span_count", "conn_count", "walkable_span_count", "max_height",
         "max_distance", "status")


def _ndev():
    import torch

    return torch.cuda.device_count()


def _same(a, b, what):
    for k in NAMES:
        x, y = getattr(a, k), getattr(b, k)
        if isinstance(x, np.ndarray):
            assert x.dtype == y.dtype and np.array_equal(x, y), f"{what}: {k} differs"
        else:
            assert x == y, f"{what}: {k} differs ({x} != {y})"


@pytest.mark.parametrize("ndev", [1, 2, 8])
def test_multi_device_build_equals_single_device(oracle, ndev):
    import recast_rs_b200.builder as rb
    from recast_rs_b200 import synth

    if _ndev() < ndev:
        pytest.skip(f"needs {ndev} CUDA devices")
    v, t, cfgs = synth.config3(seed=3, nq=330, lattice=25)  # uneven: buildings
    single = rb.Context(0)
    single.upload_mesh(v, t)
    with single.build_tiles(cfgs, rb.STAGE_ALL_BUILD) as r:
        r.download()
        want = [r.tile(i) for i in range(len(cfgs))]
        tot = r.totals()
    single.close()
    with rb.MultiContext(list(range(ndev))) as m:
        m.upload_mesh(v, t)
        for rep, (tpg, slim) in enumerate([(0, False), (5, True), (0, False)]):  # the third build reuses the first one's sizes
            r = m.build_tiles(cfgs, rb.STAGE_ALL_BUILD | (rb.RESULT_SLIM if slim else 0), tiles_per_group=tpg, download=rep != 0)
            if rep == 0:
                r.download()
            mt = r.totals()
            assert sum(mt["tiles_per_device"]) == len(cfgs) and len(mt["tiles_per_device"]) == ndev
            for k in ("tiles", "triangles", "vertices", "cells", "fragments", "spans", "connections", "walkable_spans"):
                assert mt[k] == tot[k], (k, mt[k], tot[k])
            for i in range(len(cfgs)):
                _same(r.tile(i), want[i], f"ndev {ndev} rep {rep} tile {i}")
            dev, view = r.tile_device(len(cfgs) - 1)
            assert 0 <= dev < ndev and view.span_min and view.conn_ref
            r.free()
    for i in (0, len(cfgs) - 1):
        hfa, chfa = oracle.build_tile(cfgs[i], v, t)
        assert_tile_equal(want[i], hfa, chfa, what=f"tile {i}")


def test_multi_errors():
    import recast_rs_b200.builder as rb
    from recast_rs_b200 import errors, synth

    with pytest.raises(errors.RecastError):
        rb.MultiContext([0, 4096])
    v, t, cfgs = synth.config2(seed=1, nq=48, tile=64)
    with rb.MultiContext([0]) as m:
        bad = t.copy()
        bad[5] = 10 ** 8
        m.upload_mesh(v, bad)
        with pytest.raises(errors.NavMeshGeneration):
            m.build_tiles(cfgs, rb.STAGE_ALL_BUILD)
        m.upload_mesh(v, t)
        with m.build_tiles([], rb.STAGE_ALL_BUILD) as r:
            assert r.totals()["tiles"] == 0
        cfgs[1].cs = -1.0
        with pytest.raises(errors.InvalidMesh):
            m.build_tiles(cfgs, rb.STAGE_ALL_BUILD, tiles_per_group=1)


def test_two_contexts_on_one_device_with_batches_of_different_size():
    """Two host threads, a context each, ONE device, batches whose tile kernels want different amounts of dynamic shared
    memory, issued back to back without host round trips (rcb_multi's two workers per device do exactly this).  The limit
    a kernel may ask for is one value per (kernel, device): set per launch by both threads it was a race - the thread with
    the larger batch could find the limit lowered under it (an intermittent illegal access).  Forty builds per thread,
    every one checked against the serial build of the same batch."""
    import threading

    import recast_rs_b200.builder as rb
    from recast_rs_b200 import synth

    v, t, cfgs = synth.config3(seed=5, nq=300, tile=64, lattice=14)  # 64x64 tiles (tile-resident back end) with buildings:
    # span counts, hence shared-memory sizes, differ by tile
    cfgs = list(cfgs)
    order = sorted(range(len(cfgs)), key=lambda i: (cfgs[i].bmin[0], cfgs[i].bmin[2]))
    halves = [[cfgs[i] for i in order[: len(order) // 3]], [cfgs[i] for i in order[len(order) // 3:]]]
    serial = rb.Context(0)
    serial.upload_mesh(v, t)
    want = []
    for h in halves:
        with serial.build_tiles(h, rb.STAGE_ALL_BUILD) as r:
            r.download()
            tot = r.totals()
            tot.pop("result_bytes")
            want.append((tot, r.tile(0), r.tile(len(h) - 1)))
    serial.close()
    errors = []

    def work(k):
        try:
            ctx = rb.Context(0)
            ctx.upload_mesh(v, t)
            for rep in range(40):
                with ctx.build_tiles(halves[k], rb.STAGE_ALL_BUILD) as r:
                    tot = r.totals()
                    tot.pop("result_bytes")
                    assert tot == want[k][0], (k, rep, tot, want[k][0])
                    if rep % 13 == 0:
                        r.download()
                        _same(r.tile(0), want[k][1], f"thread {k} rep {rep} first tile")
                        _same(r.tile(len(halves[k]) - 1), want[k][2], f"thread {k} rep {rep} last tile")
            assert ctx.stats()["speculative"] >= 30
            ctx.close()
        except BaseException as e:  # noqa: BLE001 - reported by the main thread
            errors.append(repr(e))

    th = [threading.Thread(target=work, args=(k,)) for k in range(2)]
    for x in th:
        x.start()
    for x in th:
        x.join()
    assert not errors, errors
