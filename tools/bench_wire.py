#!/usr/bin/env python
"""Wire-format leg on config 2's 2025 heightfields: one JSON line.

  serialize : device heightfields -> VoxelTile span data in HOST memory (kernel + D2H)
  parse     : HOST span data -> device heightfields through Heightfield::add_span (H2D + walk + replay + gather)
  cpu       : the oracle port of the same two functions on one host core, on a sample of the tiles

usage: python tools/bench_wire.py [--steps K] [--fmt 0|1] > profiles/rNN_wire.json"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import recast_rs_b200.builder as rb
from recast_rs_b200 import synth
from recast_rs_b200.config import configs_to_array

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=7)
ap.add_argument("--fmt", type=int, default=0)
ap.add_argument("--cpu-tiles", type=int, default=64)
args = ap.parse_args()

v, t, cfgs = synth.config2()
arr = configs_to_array(cfgs)
ctx = rb.Context(0)
st = torch.cuda.Stream(); torch.cuda.set_stream(st); ctx.set_stream(st.cuda_stream)
ctx.upload_mesh(v, t)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
res = ctx.build_tiles(arr, rb.STAGE_RASTER | rb.STAGE_FILTER)
tot = res.totals()
nbytes = res.span_data_size()
pinned = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True).numpy()  # what a native caller would hold
data, off = res.serialize_spans(args.fmt, out=pinned)


def timed(fn):
    ms = []
    for i in range(3 + args.steps):
        flush.zero_(); torch.cuda.synchronize()
        t0 = time.perf_counter(); fn(); torch.cuda.synchronize()
        if i >= 3: ms.append((time.perf_counter() - t0) * 1e3)
    return float(np.median(ms))


def kernels(fn):
    ctx.profile_enable(True); ctx.profile_reset()
    for _ in range(3): flush.zero_(); fn()
    k = ctx.profile_read(); ctx.profile_enable(False)
    return {n: round(m / 3, 4) for n, (m, c) in sorted(k.items(), key=lambda kv: -kv[1][0])}


ser_ms = timed(lambda: res.serialize_spans(args.fmt, out=pinned))
ser_k = kernels(lambda: res.serialize_spans(args.fmt, out=pinned))
par_ms = timed(lambda: ctx.build_from_span_data(arr, data, off, args.fmt, 0).free())
par_k = kernels(lambda: ctx.build_from_span_data(arr, data, off, args.fmt, 0).free())

# CPU port, one core, first N tiles
from oracle import oracle as orc
n = min(args.cpu_tiles, len(cfgs))
res.download()
hfs = []
for i in range(n):
    tv, c = res.tile(i), cfgs[i]
    hfs.append(orc.Heightfield.from_csr(c.width, c.height, c.bmin, c.bmax, c.cs, c.ch, tv.hf_cell_offset, tv.span_min, tv.span_max, tv.span_area))
t0 = time.perf_counter(); blobs = [h.serialize_spans(args.fmt == 0) for h in hfs]; cpu_ser = time.perf_counter() - t0
t0 = time.perf_counter()
for i in range(n):
    c = cfgs[i]
    h = orc.Heightfield(c.width, c.height, c.bmin, c.bmax, c.cs, c.ch)
    (h.deserialize_spans if args.fmt == 0 else h.reconstruct_heightfield)(blobs[i])
cpu_par = time.perf_counter() - t0
for i in range(n):
    assert blobs[i].tobytes() == data[int(off[i]):int(off[i + 1])].tobytes()

peak = 6549.8
try:
    peak = float(json.load(open("MEASURED_PEAKS.json")).get("hbm_gbs", peak))
except Exception:
    pass
ntl = len(cfgs)
k_ser = ser_k.get("k_wire_serialize", 0.0)
k_par = sum(par_k.get(k, 0.0) for k in ("k_wire_walk", "k_wire_replay", "k_gather_spans"))
balg_ser = 9 * tot["spans"] + 4 * tot["cells"] + nbytes  # CSR in, bytes out
print(json.dumps({
    "workload": "config2 heightfields (2025 tiles of 64x64 cells) through the detour-dynamic span format",
    "format": ["VoxelTile big-endian (io/voxel_tile.rs:70-176)", "DynamicTile little-endian (dynamic_tile.rs:147-257)"][args.fmt],
    "tiles": ntl, "spans": tot["spans"], "span_data_bytes": nbytes,
    "serialize": {"e2e_ms": ser_ms, "tiles_per_s": ntl / (ser_ms * 1e-3), "kernels_ms": ser_k, "d2h_bytes": nbytes,
                  "roofline": {"bound": "hbm", "kernel": "k_wire_serialize", "achieved": balg_ser / (k_ser * 1e-3) / 1e9 if k_ser else None,
                               "peak": peak, "unit": "GB/s", "frac": balg_ser / (k_ser * 1e-3) / 1e9 / peak if k_ser else None}},
    "parse": {"e2e_ms": par_ms, "tiles_per_s": ntl / (par_ms * 1e-3), "kernels_ms": par_k, "h2d_bytes": nbytes, "device_ms_main_kernels": k_par},
    "cpu_baseline": {"kind": "port", "cores": 1, "sample": f"first {n} tiles", "serialize_tiles_per_s": n / cpu_ser, "parse_tiles_per_s": n / cpu_par},
}))
