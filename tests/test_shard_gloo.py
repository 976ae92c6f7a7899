"""N>1 host path on CPU: world-size-2 gloo.  Each rank builds its shard of the tiles from a pre-binned
sub-mesh (the oracle stands in for the device builder here — this test is about the sharding, binning
and gather logic, which must reproduce the single-rank result exactly)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, mode, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist

    from oracle import oracle as orc
    from recast_rs_b200 import shard, synth

    dist.init_process_group("gloo", rank=rank, world_size=world)
    v, t, cfgs = synth.config2(seed=5, nq=40, tile=32)
    ids = shard.shard_tiles(len(cfgs), rank, world, mode)
    mine = [cfgs[i] for i in ids]
    sv, st = shard.prebin_mesh(v, t, mine)
    counts = []
    for c in mine:
        hfa, chfa = orc.build_tile(c, sv, st)
        counts.append([chfa.span_min.size, chfa.conn_ref.size, chfa.walkable_span_count, int(chfa.dist.astype(np.int64).sum())])
    table = shard.gather_tile_counts(ids, np.array(counts, dtype=np.int64).reshape(len(ids), 4), len(cfgs), world)
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, table, st.size // 3))


@pytest.mark.parametrize("mode", ["block_cyclic", "slab"])
def test_two_rank_sharding_matches_single_rank(oracle, mode):
    import torch.multiprocessing as mp

    from recast_rs_b200 import shard, synth

    v, t, cfgs = synth.config2(seed=5, nq=40, tile=32)
    want = []
    for c in cfgs:
        hfa, chfa = oracle.build_tile(c, v, t)
        want.append([chfa.span_min.size, chfa.conn_ref.size, chfa.walkable_span_count, int(chfa.dist.astype(np.int64).sum())])
    want = np.array(want, dtype=np.int64)

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, mode, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, table, ntri in got:
        np.testing.assert_array_equal(table, want)  # every rank ends with the full, identical table
        assert 0 < ntri <= t.size // 3
    if mode == "slab":  # a slab needs fewer triangles than the whole mesh
        assert all(ntri < t.size // 3 for _, _, ntri in got)


def test_shard_partitions_are_disjoint_and_complete():
    from recast_rs_b200 import shard

    for n in (1, 7, 64, 2025):
        for world in (1, 2, 3, 8):
            for mode in ("block_cyclic", "slab"):
                parts = [shard.shard_tiles(n, r, world, mode) for r in range(world)]
                allv = np.sort(np.concatenate(parts))
                np.testing.assert_array_equal(allv, np.arange(n))
                assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def test_prebin_keeps_order_and_results(oracle):
    from recast_rs_b200 import shard, synth

    v, t, cfgs = synth.config2(seed=2, nq=30, tile=32)
    sub = cfgs[:3]
    sv, st = shard.prebin_mesh(v, t, sub)
    assert sv.size < v.size  # only the vertices the slab's triangles use travel to the rank
    # subsequence of the original triangle list, same order, same coordinates (vertices are renumbered, not moved)
    tv = v.reshape(-1, 3)[t.reshape(-1, 3)]
    sc = sv.reshape(-1, 3)[st.reshape(-1, 3)]
    pos = 0
    for row in sc:
        while not np.array_equal(tv[pos], row):
            pos += 1
        pos += 1
    kv, kt = shard.prebin_mesh(v, t, sub, compact_vertices=False)
    assert kv is v and np.array_equal(v.reshape(-1, 3)[kt.reshape(-1, 3)], sc)
    for c in sub:
        a, ca = oracle.build_tile(c, v, t)
        b, cb = oracle.build_tile(c, sv, st)
        np.testing.assert_array_equal(a.span_area, b.span_area)
        np.testing.assert_array_equal(ca.dist, cb.dist)
        np.testing.assert_array_equal(ca.conn_ref, cb.conn_ref)


def test_weighted_slabs_cover_the_tiles_and_balance_cost():
    from recast_rs_b200 import shard, synth

    v, t, cfgs = synth.config3(seed=1, nq=300, lattice=22)
    cost = shard.tile_costs(v, t, cfgs)
    assert cost.shape == (len(cfgs),) and (cost > 0).all()
    for world in (1, 2, 4, 8):
        parts = [shard.shard_tiles_weighted(v, t, cfgs, r, world) for r in range(world)]
        np.testing.assert_array_equal(np.concatenate(parts), np.arange(len(cfgs)))  # contiguous, in order, complete
        loads = np.array([cost[p].sum() for p in parts])
        assert loads.max() <= loads.mean() + cost.max() * 1.01  # within one tile of the mean
