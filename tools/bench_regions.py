#!/usr/bin/env python
"""Region-building leg on config 2: device time of the path with and without RCB_STAGE_REGIONS,
the k_regions kernel time, and the oracle port of build_regions on one host core for a sample of tiles.
usage: python tools/bench_regions.py [--cpu-tiles N] > profiles/rNN_regions.json"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import recast_rs_b200.builder as rb
from recast_rs_b200 import synth
from recast_rs_b200.config import configs_to_array

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=7)
ap.add_argument("--cpu-tiles", type=int, default=48)
ap.add_argument("--config", type=int, default=2)
args = ap.parse_args()
v, t, cfgs = {2: synth.config2, 3: synth.config3, 4: synth.config4}[args.config]()
arr = configs_to_array(cfgs)
ctx = rb.Context(0)
st = torch.cuda.Stream(); torch.cuda.set_stream(st); ctx.set_stream(st.cuda_stream)
dv, dt = torch.from_numpy(v).cuda(), torch.from_numpy(t).cuda()
ctx.set_mesh_device(dv.data_ptr(), dv.numel(), dt.data_ptr(), dt.numel(), keep=(dv, dt))
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timed(mask):
    for _ in range(3): ctx.build_tiles(arr, mask).free()
    ms = []
    for _ in range(args.steps):
        flush.zero_(); a, b = torch.cuda.Event(True), torch.cuda.Event(True)
        a.record(st); r = ctx.build_tiles(arr, mask); b.record(st); r.free(); torch.cuda.synchronize(); ms.append(a.elapsed_time(b))
    ctx.profile_enable(True); ctx.profile_reset()
    for _ in range(3): flush.zero_(); ctx.build_tiles(arr, mask).free()
    k = ctx.profile_read(); ctx.profile_enable(False)
    return float(np.median(ms)), {n: round(m / 3, 4) for n, (m, c) in sorted(k.items(), key=lambda kv: -kv[1][0])[:6]}


base_ms, base_k = timed(rb.STAGE_ALL_BUILD)
reg_ms, reg_k = timed(rb.STAGE_ALL_BUILD | rb.STAGE_REGIONS)
with ctx.build_tiles(arr, rb.STAGE_ALL_BUILD | rb.STAGE_REGIONS) as r:
    tot = r.totals()
    r.download()
    nreg = sum(int(r.tile(i).reg.max(initial=0)) for i in range(len(cfgs)))
    sample = [(i, r.tile(i).reg.copy()) for i in range(min(args.cpu_tiles, len(cfgs)))]

from oracle import oracle as orc
chfs = []
for i, _ in sample:
    hf = orc.Heightfield.from_config(cfgs[i]); hf.builder_rasterize(cfgs[i], v, t)
    chfs.append(orc.CompactHeightfield.build_from_heightfield(hf))
t0 = time.perf_counter()
regs = [c.build_regions(cfgs[i].border_size, cfgs[i].min_region_area, cfgs[i].merge_region_area)[0] for (i, _), c in zip(sample, chfs)]
cpu = time.perf_counter() - t0
for (i, g), w in zip(sample, regs):
    assert np.array_equal(g, w), i
k_ms = reg_k.get("k_regions", 0.0)
print(json.dumps({
    "workload": f"config{args.config}: {len(cfgs)} tiles, {tot['spans']} spans, {nreg} regions after filtering",
    "device_ms_without_regions": base_ms, "device_ms_with_regions": reg_ms, "k_regions_ms": k_ms, "kernels_ms": reg_k,
    "tiles_per_s_regions_only": len(cfgs) / (k_ms * 1e-3) if k_ms else None,
    "cpu_baseline": {"kind": "port", "cores": 1, "sample": f"first {len(sample)} tiles, build_regions incl. its distance field",
                     "tiles_per_s": len(sample) / cpu, "ms_per_tile": cpu / len(sample) * 1e3},
}))
