#!/usr/bin/env python
"""bench.py — tiled voxelization throughput (BASELINE.json metric) on N B200s.

A step = one pass of the hot path (rasterize -> filters -> compact heightfield + connections ->
distance field) over every tile of the workload.  One process per GPU; tiles shard across ranks with
no data-path collective (each rank owns a slab of the world: its tiles and the triangles over them),
so scaling is "weak".

  value : triangles voxelized per second, whole job, inputs resident in HBM, results left in HBM
  e2e   : same metric through the C ABI with HOST buffers: per step H2D of the mesh from pinned
          memory, the build, and D2H of every result array into pinned memory
  roofline: algorithmic bytes / device time of the step vs the measured HBM peak
  cpu_baseline: the C++ oracle (a port of the reference; the Rust reference cannot be built here)
          on the host cores, bounded sample

`--impl reference` times the oracle port with all host threads (the reference arm of the contract).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

_REAL_STDOUT = None


def emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is not None:
        os.write(_REAL_STDOUT, data)
    else:
        sys.stdout.write(data.decode())
        sys.stdout.flush()


METRIC = "input triangles voxelized/sec"
UNIT = "triangles/s"


def make_workload(config: int, seed: int):
    from recast_rs_b200 import synth

    if config == 1:
        v, t, cfgs = synth.config1(seed)
        name = "config1: 100k-triangle terrain(nq=224), single tile (cs=0.3, ch=0.2)"
    elif config == 2:
        v, t, cfgs = synth.config2(seed)
        name = "config2: 1M-triangle terrain(nq=708), 64x64-cell tiles (2025 tiles), cs=0.3 ch=0.2"
    elif config == 3:
        v, t, cfgs = synth.config3(seed)
        name = "config3: 10M-triangle open world (terrain nq=2200 + 26896 boxes), 128x128-cell tiles"
    elif config == 4:
        v, t, cfgs = synth.config4(seed)
        name = "config4: 1M-triangle 16-floor interior, 64x64-cell tiles"
    elif config == 0:  # tiny, for CPU-side tests of this script
        v, t, cfgs = synth.config2(seed, nq=64, tile=64)
        name = "tiny: 8k-triangle terrain, 64x64-cell tiles"
    else:
        raise SystemExit(f"unknown config {config}")
    return v, t, cfgs, name


def make_sharded_workload(config: int, seed: int, rank: int, world: int, mode: str = "slab"):
    """N > 1: ONE world with ~world x the triangles of the single-GPU workload (weak scaling); tiles are
    split into contiguous slabs of tile rows, one per rank, and each rank receives the triangles over its
    slab (order preserved).  No data moves between ranks on the build path.
    Returns (verts, tris, cfgs, name, world_triangles, world_tiles, (world_verts, world_tris)): the last pair is the
    UNSHARDED mesh, which the parity check feeds to the oracle."""
    from recast_rs_b200 import shard, synth

    if world == 1:
        v, t, cfgs, name = make_workload(config, seed)
        return v, t, cfgs, name, t.size // 3, len(cfgs), (v, t)
    if config == 3:  # BASELINE configs[2]: ONE 10M-triangle world, its tiles sharded over the ranks (strong scaling)
        v, t, all_cfgs, name = make_workload(3, seed)
        ids = shard.shard_tiles_weighted(v, t, all_cfgs, rank, world) if mode == "weighted" else shard.shard_tiles(len(all_cfgs), rank, world, mode)
        cfgs = [all_cfgs[i] for i in ids]
        sv, st = shard.prebin_mesh(v, t, cfgs)
        return sv, st, cfgs, name + f", one world {mode}-sharded over {world} ranks", t.size // 3, len(all_cfgs), (v, t)
    if config != 2:
        v, t, cfgs, name = make_workload(config, seed + rank)  # independent replicas of the named config
        return v, t, cfgs, name + f" x {world} independent slabs", (t.size // 3) * world, len(cfgs) * world, (v, t)
    nq = shard.weak_scaling_world(708, world)
    v, t = synth.terrain(nq, 1.2, 4.0, seed)
    all_cfgs = synth.tile_grid(v, 64)
    ids = shard.shard_tiles(len(all_cfgs), rank, world, "slab")
    cfgs = [all_cfgs[i] for i in ids]
    sv, st = shard.prebin_mesh(v, t, cfgs)
    name = (f"config2 scaled x{world}: {t.size // 3}-triangle terrain(nq={nq}), 64x64-cell tiles ({len(all_cfgs)} tiles), "
            f"slab-sharded over {world} ranks, cs=0.3 ch=0.2")
    return sv, st, cfgs, name, t.size // 3, len(all_cfgs), (v, t)


def alg_bytes(tot):
    """B_alg of SURVEY.md §8(d) / BASELINE.md §3."""
    T, V, C, S, K = tot["triangles"], tot["vertices"], tot["cells"], tot["spans"], tot["connections"]
    return 12 * T + 12 * V + 8 * C + 5 * S + 1 * S + 8 * S + 2 * S + 5 * K


def read_traffic():
    """dram__bytes_read+write per launch of the main kernels, from the committed ncu --set full capture."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        return json.load(open(p))
    except Exception:
        return None


def read_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clock / throttle-reason samples during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.idx)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_leg(v, t, cfgs, threads, budget_s, prebinned=False):
    """Oracle port on a bounded sample of tiles sized for ~budget_s; returns (triangles/s, sample text, secs)."""
    from oracle import oracle as orc

    ntiles, T = len(cfgs), t.size // 3
    probe = min(ntiles, max(threads * 2, 8))
    s0, _ = orc.build_tiles_timed(v, t, cfgs[:probe], nthreads=threads, prebinned=prebinned)
    per_tile = max(s0 / probe, 1e-6)
    m = int(min(ntiles, max(probe, budget_s / per_tile)))
    secs, _ = orc.build_tiles_timed(v, t, cfgs[:m], nthreads=threads, prebinned=prebinned)
    how = "each handed only the triangles over it" if prebinned else "each scans the whole mesh as lib.rs:217 does"
    return T * (m / ntiles) / secs, f"first {m} of {ntiles} tiles ({how})", secs


def run_reference(args, rank, world):
    """Reference arm: the oracle port (kind 'port') with all host threads; rank 0 only."""
    if rank != 0:
        return
    from oracle import oracle as orc

    # Always the single-GPU world of the named config: the reference scans the whole mesh for every tile (lib.rs:217), so its
    # triangles/s fall with the size of the world; timing it on the N-times larger world of the native arm's weak scaling
    # would measure that quadratic loop, not the host.  (The driver divides the native arm's value by this line's.)
    v, t, cfgs, name = make_workload(args.config, args.seed)
    threads = orc.hardware_concurrency()
    vals, ms = [], []
    sample = ""
    for i in range(args.warmup + args.steps):
        val, sample, secs = cpu_leg(v, t, cfgs, threads, budget_s=max(2.0, 60.0 / (args.warmup + args.steps)))
        if i >= args.warmup:
            vals.append(val)
            ms.append(secs * 1e3)
    value = float(np.mean(vals))
    # the same port when every tile is handed only the triangles over it (what a caller with a spatial index would do;
    # the reference has none): reported so the native/reference ratio can be read both ways
    pre_val, pre_sample, _ = cpu_leg(v, t, cfgs, threads, budget_s=6.0, prebinned=True)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": float(np.mean(ms)), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32 clip + i16/u16 spans", "data": "synthetic",
        "config": {"workload": name, "seed": args.seed},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                         "prebinned_value": pre_val, "prebinned_sample": pre_sample,
                         "note": "C++ restatement of recast-rs (oracle/), not the Rust build: no cargo/rustc in the image; at N > 1 "
                                 "still the single-GPU world of the config"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


def run_tilecache(args, rank, world, local_rank):
    """BASELINE config 5: batched rebuild of 256 dirty tile-cache layers per frame; metric = tiles/s."""
    import torch

    import recast_rs_b200.builder as rb
    from recast_rs_b200 import tilecache as tcm

    layers, obstacles = tcm.config5(256, 48, seed=args.seed + rank)
    name = "config5: 256 dirty tile-cache layers of 48x48 cells per frame, 1-3 cylinder/box/oriented-box obstacles each"
    comp = None
    if args.tc_compressed:
        # the layers as TileCache stores them: TileCacheLayer::to_bytes under lz4 (liblz4 via pyarrow is the compressor)
        import pyarrow as pa

        layers, obstacles = tcm.config5_varied(256, 48, seed=args.seed + rank)
        codec = pa.Codec("lz4_raw")
        raws = [tcm.layer_to_bytes(l, l.cons) for l in layers]
        comp = tcm.CompressedTiles([len(r).to_bytes(4, "little") + codec.compress(r, asbytes=True) for r in raws], obstacles,
                                   [(l.first_obstacle, l.obstacle_count) for l in layers])
        name = ("config5c: 256 dirty tile-cache layers of 48x48 cells per frame, stored LZ4-compressed (%d -> %d bytes), decompress + "
                "from_bytes + build + obstacles" % (sum(map(len, raws)), int(comp.off[-1])))
    if args.impl == "reference":
        if rank != 0:
            return
        from oracle import oracle as orc

        threads = orc.hardware_concurrency()
        ms = []
        for i in range(args.warmup + args.steps):
            secs, _ = (orc.tilecache_build_compressed_timed(comp, 0.3, 0.2, nthreads=threads) if comp is not None else
                       orc.tilecache_build_timed(layers, obstacles, 0.3, 0.2, nthreads=threads))
            if i >= args.warmup:
                ms.append(secs * 1e3)
        value = len(layers) / (np.mean(ms) * 1e-3)
        emit({"impl": "reference", "metric": "tile-cache tiles rebuilt/sec", "value": value, "unit": "tiles/s", "n_gpus": args.gpus,
              "steps": args.steps, "warmup": args.warmup, "ms_per_step": float(np.mean(ms)), "higher_is_better": True, "scaling": "weak",
              "vs_baseline": None, "dtype": "u8/i16 cells + f32 obstacle tests", "data": "synthetic", "config": {"workload": name},
              "cpu_baseline": {"value": value, "unit": "tiles/s", "cores": threads, "kind": "port", "sample": "all 256 layers"},
              "e2e": {"value": value, "unit": "tiles/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
        return
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod

        dist = dist_mod
        dist.init_process_group("nccl", device_id=dev)
    ctx = rb.Context(local_rank)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    packed = tcm.Packed(layers, obstacles)  # the C structs a native caller already holds
    if comp is not None:
        build = lambda: ctx.build_tilecache_compressed(comp, None, None, 0.3, 0.2)  # noqa: E731
    else:
        build = lambda: ctx.build_tilecache_layers(packed, None, 0.3, 0.2)  # noqa: E731
    tot = None
    for _ in range(max(args.warmup, 3)):
        with build() as r:
            tot = r.totals()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = ctx.launch_count()
    dev_ms, e2e_ms = [], []
    for _ in range(args.steps):  # the layers are HOST data in the reference too: every step is end to end
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        a.record(stream)
        r = build()
        b.record(stream)
        r.download()
        torch.cuda.synchronize()
        e2e_ms.append((time.perf_counter() - t0) * 1e3)
        dev_ms.append(a.elapsed_time(b))
        d2h = r.totals()["result_bytes"]
        r.free()
    launches = ctx.launch_count() - l0
    t_hold = time.perf_counter()
    while time.perf_counter() - t_hold < 0.6:  # same load, untimed, until nvidia-smi has sampled it
        build().free()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    clocks["window"] = "timed region + identical untimed steps for 0.6 s (nvidia-smi -lms 100)"
    agg = torch.tensor([float(sum(dev_ms)), float(sum(e2e_ms))], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(agg, op=dist.ReduceOp.MAX)
    if rank == 0:
        peak, peak_src = read_peaks()
        ntl = len(layers) * world
        C, S, K = tot["cells"], tot["spans"], tot["connections"]
        balg = 2 * C + 8 * C + 5 * S + S + 8 * S + 5 * K  # regons+areas in; cells, spans, areas, con, connections out
        ms_step = float(agg[0]) / args.steps
        line = {"metric": "tile-cache tiles rebuilt/sec", "value": ntl * args.steps / (float(agg[0]) * 1e-3), "unit": "tiles/s",
                "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "u8/i16 cells + f32 obstacle tests", "data": "synthetic",
                "config": {"workload": name, "totals_rank0": tot, "note": "value includes the H2D of the layer bytes (host data by definition)"},
                "e2e": {"value": ntl * args.steps / (float(agg[1]) * 1e-3), "unit": "tiles/s", "h2d_bytes_per_step": int(comp.off[-1]) if comp is not None else int(2 * C),
                        "d2h_bytes_per_step": int(d2h), "ms_per_step": float(agg[1]) / args.steps},
                "gpu_launches": int(launches), "clocks": clocks,
                "roofline": {"bound": "hbm", "achieved": balg / (ms_step * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                             "frac": balg / (ms_step * 1e-3) / 1e9 / peak, "traffic": None, "peak_source": peak_src,
                             "kernel": "whole step (launch-latency bound at this size)"}}
        if world == 1 and not args.no_cpu:
            from oracle import oracle as orc

            secs, _ = (orc.tilecache_build_compressed_timed(comp, 0.3, 0.2, nthreads=1) if comp is not None else
                       orc.tilecache_build_timed(layers, obstacles, 0.3, 0.2, nthreads=1))
            line["cpu_baseline"] = {"value": len(layers) / secs, "unit": "tiles/s", "cores": 1, "kind": "port", "sample": "all 256 layers"}
        emit(line)
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()


def config3_leg(args, rank, world, local_rank, ctx, stream, flush, dist, dev, barrier):
    """BASELINE configs[2] (the north_star's target): ONE 10M-triangle open world, 128x128-cell tiles, its tiles dealt to
    the ranks in cost-balanced slabs (strong scaling).  Returns the sub-record rank 0 prints (None elsewhere)."""
    import torch

    import recast_rs_b200.builder as rb
    from oracle import oracle as orc
    from recast_rs_b200.config import configs_to_array
    from tests.parity import assert_tile_equal

    from recast_rs_b200 import shard

    wv, wt, all_cfgs, name = make_workload(3, args.seed)
    T_world, tiles_world, whole = wt.size // 3, len(all_cfgs), (wv, wt)
    cost = shard.tile_costs(wv, wt, all_cfgs)
    rounds = []

    def load(cost):
        cuts = shard.cut_by_cost(cost, world)
        cfgs = all_cfgs[cuts[rank]:cuts[rank + 1]]
        v, t = shard.prebin_mesh(wv, wt, cfgs)
        arr = configs_to_array(cfgs)
        d_v, d_t = torch.from_numpy(v).to(dev), torch.from_numpy(t).to(dev)
        ctx.set_mesh_device(d_v.data_ptr(), d_v.numel(), d_t.data_ptr(), d_t.numel(), keep=(d_v, d_t))
        tot = None
        for _ in range(3):
            with ctx.build_tiles(arr, rb.STAGE_ALL_BUILD) as r:
                tot = r.totals()
        return cuts, cfgs, v, t, arr, tot

    def measure(arr, n):
        evs = []
        for _ in range(n):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            ctx.build_tiles(arr, rb.STAGE_ALL_BUILD).free()
            b.record(stream)
            evs.append((a, b))
        torch.cuda.synchronize()
        return float(np.median([a.elapsed_time(b) for a, b in evs]))

    # static slabs cut by ESTIMATED cost, then one round of feedback: every rank times its slab, the slabs are re-priced by
    # what they measured and cut again (shard.refine_costs) - untimed set-up, like the pre-binning itself
    cuts, cfgs, v, t, arr, tot = load(cost)
    for it in range(2):
        mine = torch.tensor([measure(arr, 3)], dtype=torch.float64, device=dev)
        allms = [torch.zeros_like(mine) for _ in range(world)]
        if dist is not None:
            dist.all_gather(allms, mine)
        else:
            allms = [mine]
        ms_all = [float(x[0]) for x in allms]
        rounds.append({"cuts": [int(c) for c in cuts], "ms_per_rank": [round(x, 3) for x in ms_all]})
        if it == 1 or world == 1:
            break
        cost = shard.refine_costs(cost, cuts, ms_all)
        cuts, cfgs, v, t, arr, tot = load(cost)
    name += f", one world sharded over {world} ranks (contiguous slabs cut by estimated cost, re-cut once by measured time)"
    steps = max(3, min(args.steps, 5))
    warm = [ctx.build_tiles(arr, rb.STAGE_ALL_BUILD) for _ in range(steps)]  # arena pool: one pair per timed step
    for r in warm:
        r.totals()
        r.free()
    barrier()
    evs, held = [], []
    for _ in range(steps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        held.append(ctx.build_tiles(arr, rb.STAGE_ALL_BUILD))
        b.record(stream)
        evs.append((a, b))
    barrier()
    for r in held:
        tt = r.totals()
        assert {k: x for k, x in tt.items() if k != "result_bytes"} == {k: x for k, x in tot.items() if k != "result_bytes"}, tt
        r.free()
    ms = float(sum(a.elapsed_time(b) for a, b in evs))
    with ctx.build_tiles(arr, rb.STAGE_ALL_BUILD) as r:  # parity: the slab's border tiles against the oracle on the whole world
        r.download()
        for i in sorted({0, len(cfgs) - 1}):
            hfa, chfa = orc.build_tile(cfgs[i], whole[0], whole[1])
            assert_tile_equal(r.tile(i), hfa, chfa, what=f"config3 rank {rank} tile {i}")
    agg = torch.tensor([ms], dtype=torch.float64, device=dev)
    mn = torch.tensor([ms], dtype=torch.float64, device=dev)
    cnt = torch.tensor([float(alg_bytes(tot)), float(len(cfgs)), float(v.nbytes + t.nbytes)], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(agg, op=dist.ReduceOp.MAX)
        dist.all_reduce(mn, op=dist.ReduceOp.MIN)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    if rank != 0:
        return None
    peak, _ = read_peaks()
    ms_step = float(agg[0]) / steps
    achieved = float(cnt[0]) / world / (ms_step * 1e-3) / 1e9
    return {"workload": name, "scaling": "strong", "n_gpus": world, "steps": steps, "ms_per_step": ms_step,
            "ms_per_step_fastest_rank": float(mn[0]) / steps, "value": T_world / (ms_step * 1e-3), "unit": UNIT,
            "tiles_per_s": tiles_world / (ms_step * 1e-3), "roofline_frac": achieved / peak, "achieved_gbs_per_gpu": achieved,
            "h2d_bytes_all_ranks": int(cnt[2]), "h2d_bytes_whole_world": int(whole[0].nbytes + whole[1].nbytes),
            "parity_check": {"tiles": 2 * world, "ok": True},
            "sharding": "contiguous tile slabs cut by estimated cost (shard.tile_costs), re-cut once by the time every rank measured (shard.refine_costs)",
            "sharding_rounds": rounds}


def run_native(args, rank, world, local_rank):
    import torch

    import recast_rs_b200.builder as rb

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; recast_rs_b200 has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod

        dist = dist_mod
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # every rank owns one slab of the world: its own tiles and the triangles over them
    v, t, cfgs, name, T_world, tiles_world, whole = make_sharded_workload(args.config, args.seed, rank, world, args.shard)
    from recast_rs_b200.config import configs_to_array

    cfg_arr = configs_to_array(cfgs)
    ctx = rb.Context(local_rank)
    # a dedicated (non-default) stream: the library launches on exactly this stream, so the events
    # recorded on it bracket every kernel of the step, including the ones after the last host sync
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    ctx.set_stream(stream.cuda_stream)
    d_v = torch.from_numpy(v).to(dev)
    d_t = torch.from_numpy(t).to(dev)
    ctx.set_mesh_device(d_v.data_ptr(), d_v.numel(), d_t.data_ptr(), d_t.numel(), keep=(d_v, d_t))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    stage = rb.STAGE_ALL_BUILD

    def step():
        r = ctx.build_tiles(cfg_arr, stage)
        return r

    tot = None
    for _ in range(max(args.warmup, 3)):
        r = step()
        tot = r.totals()
        r.free()
    # the timed steps keep their results until the region is over (they are checked then): let the context's arena pool
    # hold that many arenas now, so that no step allocates device memory inside the timed region
    warm = [step() for _ in range(args.steps)]
    for r in warm:
        r.totals()
        r.free()
    barrier()

    # ---- timed region: K steps, device time by CUDA events on the launching stream ----------------------
    sampler = ClockSampler(local_rank)
    sampler.start()
    evs = []
    l0 = ctx.launch_count()
    barrier()
    wall0 = time.perf_counter()
    held = []
    for _ in range(args.steps):
        flush.zero_()  # L2 flush between timed iterations (outside the event pair)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        r = step()
        b.record(stream)
        held.append(r)  # the host does not wait inside the step: what it produced is checked after the timed region
        evs.append((a, b))
    barrier()
    wall = time.perf_counter() - wall0
    launches = ctx.launch_count() - l0
    st0 = ctx.stats()
    for r in held:  # every timed step must have produced the full result without being rebuilt
        tt = r.totals()
        if {k: x for k, x in tt.items() if k != "result_bytes"} != {k: x for k, x in tot.items() if k != "result_bytes"}:
            raise SystemExit(f"bench.py: a timed step produced {tt}, expected {tot}")
        r.free()
    st1 = ctx.stats()
    if st1["rebuilt"] != st0["rebuilt"]:
        raise SystemExit(f"bench.py: {st1['rebuilt'] - st0['rebuilt']} timed steps had to be rebuilt (sizes assumed from the previous build were short)")
    # the timed region lasts tens of milliseconds, less than one nvidia-smi period: keep the same load running
    # (untimed) until the sampler has seen it, so the clocks line describes this workload and not an idle GPU
    while time.perf_counter() - wall0 < 0.7:
        flush.zero_()
        step().free()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    clocks["window"] = "timed region + identical untimed steps up to 0.7 s (nvidia-smi -lms 100)"
    dev_ms = [a.elapsed_time(b) for a, b in evs]
    total_ms = float(sum(dev_ms))

    # ---- per-kernel device time (separate, untimed pass) ---------------------------------------------------
    ctx.profile_enable(True)
    ctx.profile_reset()
    nprof = 3
    for _ in range(nprof):
        flush.zero_()
        step().free()
    kern = ctx.profile_read()
    ctx.profile_enable(False)
    kern_ms = {k: ms / nprof for k, (ms, _) in kern.items()}

    # ---- e2e: host buffers in, host buffers out ------------------------------------------------------------
    # The public host-to-host call: rcb_upload_mesh + rcb_build_tiles_streamed with the slim result layout
    # (RCB_RESULT_SLIM: the arrays the binding layer rebuilds from the layout rule stay on the device).  The D2H
    # of a finished tile group overlaps the kernels of the next ones.  The full-layout, one-group call is timed
    # beside it (`e2e.full_layout`): same metric, 1.7x the bytes.
    hv = torch.from_numpy(v).pin_memory()
    ht = torch.from_numpy(t).pin_memory()
    ctx2 = rb.Context(local_rank)  # a context that owns its mesh copy
    ctx2.set_stream(stream.cuda_stream)

    def e2e_leg(slim, groups, reps):
        ms, nbytes, check = [], 0, None
        mask = stage | (rb.RESULT_SLIM if slim else 0)
        for i in range(2 + reps):
            flush.zero_()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            ctx2.upload_mesh(hv.numpy(), ht.numpy())
            if groups > 1:
                r = ctx2.build_tiles_streamed(cfg_arr, mask, 0, groups)
            else:
                r = ctx2.build_tiles(cfg_arr, mask)
                r.download()
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) * 1e3
            tt = r.totals()
            nbytes = tt["result_bytes"]
            if {k: x for k, x in tt.items() if k != "result_bytes"} != {k: x for k, x in tot.items() if k != "result_bytes"}:
                raise SystemExit(f"bench.py: an end-to-end step produced {tt}, expected {tot}")
            if i == 2 + reps - 1:
                check = r.tile(0)  # host arrays of the last step (the parity check below reads them)
            r.free()
            if i >= 2:
                ms.append(dt)
        return ms, nbytes, check

    e2e_ms, d2h_bytes, e2e_tile0 = e2e_leg(True, max(args.e2e_groups, 1), args.steps)
    full_ms, full_bytes, _ = e2e_leg(False, 1, max(3, args.steps // 2))
    # the same call from two host threads, one context (and stream) each, taking alternate steps: the D2H of one step's
    # results overlaps the upload and the kernels of the next.  Reported beside the one-context number, never instead of it.
    pipe = None
    if world == 1:
        try:
            import threading
            nctx = 2
            pctx = [ctx2, rb.Context(local_rank)]
            pst = [stream, torch.cuda.Stream(device=dev)]
            pctx[1].set_stream(pst[1].cuda_stream)
            err = []

            def worker(j, n):
                try:
                    for _ in range(j, n, nctx):
                        pctx[j].upload_mesh(hv.numpy(), ht.numpy())
                        rr = pctx[j].build_tiles(cfg_arr, stage | rb.RESULT_SLIM)
                        rr.download()
                        rr.free()
                except Exception as e:  # noqa: BLE001
                    err.append(e)

            def run(n):
                th = [threading.Thread(target=worker, args=(j, n)) for j in range(nctx)]
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                for x in th:
                    x.start()
                for x in th:
                    x.join()
                torch.cuda.synchronize()
                return (time.perf_counter() - t0) * 1e3

            run(2 * nctx)  # warm-up: both contexts size their pools
            npipe = max(args.steps, 2 * nctx)
            pms = run(npipe)
            if not err:
                pipe = {"value": T_world * npipe / (pms * 1e-3), "unit": UNIT, "ms_per_step": pms / npipe, "steps": npipe,
                        "call": "%d host threads, one rcb_ctx each, alternate steps of rcb_upload_mesh + rcb_build_tiles(RCB_RESULT_SLIM) + rcb_result_download; no L2 flush between steps" % nctx}
            pctx[1].close()
        except Exception:  # noqa: BLE001 - an extra figure, never a reason to lose the line
            pipe = None
    barrier()

    # ---- parity check (outside every timed region): this rank's first and last tile - the borders of its slab -
    # against the oracle run on the UNSHARDED world; fails the run on a mismatch --------------------------------
    parity = None
    if not args.no_parity:
        from oracle import oracle as orc
        from tests.parity import assert_tile_equal

        wv, wt = whole
        with ctx.build_tiles(cfg_arr, stage) as r:
            r.download()
            picks = sorted({0, len(cfgs) - 1}) if len(cfgs) else []
            for i in picks:
                hfa, chfa = orc.build_tile(cfgs[i], wv, wt)
                assert_tile_equal(r.tile(i), hfa, chfa, what=f"rank {rank} tile {i}")
        if e2e_tile0 is not None and len(cfgs):  # the slim, streamed host-to-host path returns the same tile
            hfa, chfa = orc.build_tile(cfgs[0], wv, wt)
            assert_tile_equal(e2e_tile0, hfa, chfa, what=f"rank {rank} e2e tile 0")
        parity = len(picks) + (1 if e2e_tile0 is not None and len(cfgs) else 0)
    pc = torch.tensor([float(parity or 0)], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(pc, op=dist.ReduceOp.SUM)

    # ---- BASELINE configs[2] beside the headline at N > 1: ONE 10M-triangle world, tiles cost-balanced over the ranks -------
    c3 = None
    if world > 1 and args.config == 2 and not args.no_config3:
        c3 = config3_leg(args, rank, world, local_rank, ctx, stream, flush, dist, dev, barrier)

    # ---- aggregate over ranks ------------------------------------------------------------------------------
    agg = torch.tensor([total_ms, float(sum(e2e_ms)), float(sum(full_ms))], dtype=torch.float64, device=dev)
    cnt = torch.tensor([float(alg_bytes(tot))], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(agg, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    max_total_ms, max_e2e_ms, max_full_ms = float(agg[0]), float(agg[1]), float(agg[2])
    # triangles of the WORLD (a triangle on a slab border is rasterized by two ranks but counted once)
    T_all, tiles_all, balg_all = float(T_world), float(tiles_world), float(cnt[0])

    if rank == 0:
        peak, peak_src = read_peaks()
        ms_per_step = max_total_ms / args.steps
        value = T_all * args.steps / (max_total_ms * 1e-3)
        e2e_value = T_all * len(e2e_ms) / (max_e2e_ms * 1e-3)
        achieved = balg_all / world / (ms_per_step * 1e-3) / 1e9  # per GPU
        dom = max(kern_ms.items(), key=lambda kv: kv[1]) if kern_ms else ("", 0.0)
        ksum = sum(kern_ms.values())
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if (args.config == 3 and world > 1) else "weak",
            "vs_baseline": None,
            "dtype": "f32 clip + i16/u16 spans", "data": "synthetic",
            "config": {"workload": name,
                       "seed": args.seed, "stages": "raster+filters+compact+connections+distance_field",
                       "l2": "256 MiB flush between timed steps; per-step working set > L2",
                       "totals_rank0": tot},
            "tiles_per_s": tiles_all * args.steps / (max_total_ms * 1e-3),
            "wall_ms_per_step": wall * 1e3 / args.steps,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(v.nbytes + t.nbytes),
                    "d2h_bytes_per_step": int(d2h_bytes), "ms_per_step": max_e2e_ms / len(e2e_ms),
                    "call": ("rcb_upload_mesh + rcb_build_tiles_streamed(groups=%d, RCB_RESULT_SLIM)" % args.e2e_groups) if args.e2e_groups > 1
                    else "rcb_upload_mesh + rcb_build_tiles(RCB_RESULT_SLIM) + rcb_result_download",
                    "layout": "slim: conn_mask u8[S] + conn_ref16 u16[K] instead of first_conn, conn_next, conn_dir, conn_ref; "
                              "hf_cell_offset left out (prefix sum of cell_count); rebuilt by the binding layer (INTEGRATION.md)",
                    "full_layout": {"value": T_world * len(full_ms) / (max_full_ms * 1e-3), "unit": UNIT, "ms_per_step": max_full_ms / len(full_ms),
                                    "d2h_bytes_per_step": int(full_bytes), "call": "rcb_upload_mesh + rcb_build_tiles + rcb_result_download"},
                    **({"pipelined": pipe} if pipe else {})},
            "gpu_launches": int(launches),
            "parity_check": {"tiles": int(pc[0]), "ok": True, "against": "oracle on the unsharded world: first and last tile of every rank's slab, "
                             "plus tile 0 of the slim streamed end-to-end call"} if not args.no_parity else None,
            **({"config3": c3} if c3 else {}),
            "issue": {"steps_issued_without_host_round_trip": st1["speculative"], "steps_rebuilt": st1["rebuilt"]},
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": (read_traffic() or {}).get("main_kernels_sum") if args.config == 2 and world == 1 else None,
                         "traffic_per_kernel": (read_traffic() or {}).get("per_launch_dram_bytes") if args.config == 2 and world == 1 else None,
                         # the kernels are instruction-issue / latency bound, not HBM bound: the roofline that applies to them
                         "issue_roofline": (read_traffic() or {}).get("issue_roofline") if args.config == 2 and world == 1 else None,
                         "peak_source": peak_src, "kernel": "whole step (all launches)",
                         "alg_bytes_per_step_per_gpu": balg_all / world,
                         "alg_bytes_per_triangle": balg_all / T_all,
                         "dominant_kernel": {"name": dom[0], "ms": dom[1], "share_of_kernel_time": dom[1] / ksum if ksum else None},
                         "kernel_ms": {k: round(ms, 4) for k, ms in sorted(kern_ms.items(), key=lambda kv: -kv[1])},
                         "kernel_ms_sum": ksum},
        }
        if world == 1 and not args.no_cpu:
            try:
                val, sample, _ = cpu_leg(v, t, cfgs, 1, budget_s=args.cpu_budget)
                cb = {"value": val, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample,
                      "note": "C++ restatement of recast-rs (oracle/), single thread like the reference (its `parallel` feature is dead code)"}
                val2, _, _ = cpu_leg(v, t, cfgs, 1, budget_s=args.cpu_budget / 3, prebinned=True)
                cb["prebinned_value"] = val2
                line["cpu_baseline"] = cb
            except Exception as e:  # the oracle is test infrastructure; its absence must not kill the GPU number
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": f"failed: {e}"}
        emit(line)
    ctx2.close()
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--config", type=int, default=2)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--shard", default="weighted", choices=["weighted", "slab", "block_cyclic"], help="how config 3's tiles are dealt to the ranks")
    ap.add_argument("--e2e-groups", type=int, default=4, help="tile groups of the host-to-host call (1: build then download)")
    ap.add_argument("--tc-compressed", action="store_true", help="config 5 from LZ4-compressed tiles (SURVEY 8(f) rank 3)")
    ap.add_argument("--cpu-budget", type=float, default=12.0)
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle check of this rank's border tiles")
    ap.add_argument("--no-config3", action="store_true", help="N > 1: skip the strong-scaling sub-record of BASELINE configs[2]")
    args = ap.parse_args()
    # stdout carries exactly ONE JSON line: anything else written to fd 1 (NCCL's version banner,
    # library chatter) is rerouted to stderr
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.config == 5:
        run_tilecache(args, rank, world, local_rank)
    elif args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_native(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
