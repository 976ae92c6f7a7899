#!/usr/bin/env python
"""Quick per-kernel timing of config 2 (no e2e, no CPU leg): prints ms/step and the kernel table."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import recast_rs_b200.builder as rb
from recast_rs_b200 import synth
from recast_rs_b200.config import configs_to_array
cfgno = int(sys.argv[1]) if len(sys.argv) > 1 else 2
v, t, cfgs = {1: synth.config1, 2: synth.config2, 3: synth.config3, 4: synth.config4}[cfgno]()
arr = configs_to_array(cfgs)
ctx = rb.Context(0)
st = torch.cuda.Stream(); torch.cuda.set_stream(st); assert st.cuda_stream != 0; ctx.set_stream(st.cuda_stream)
dv, dt = torch.from_numpy(v).cuda(), torch.from_numpy(t).cuda()
ctx.set_mesh_device(dv.data_ptr(), dv.numel(), dt.data_ptr(), dt.numel(), keep=(dv, dt))
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for _ in range(3):
    with ctx.build_tiles(arr) as r0: tot = r0.totals()
ms = []
for _ in range(7):
    flush.zero_(); a, b = torch.cuda.Event(True), torch.cuda.Event(True)
    a.record(st); r = ctx.build_tiles(arr); b.record(st); r.free(); torch.cuda.synchronize(); ms.append(a.elapsed_time(b))
ctx.profile_enable(True); ctx.profile_reset()
for _ in range(3): flush.zero_(); ctx.build_tiles(arr).free()
k = ctx.profile_read()
with ctx.build_tiles(arr) as r0: assert r0.totals()["spans"] == tot["spans"]
print(f"cfg{cfgno} band_cells={os.environ.get('RCB_BAND_CELLS','-')} {ctx.stats()} spans={tot['spans']} median {np.median(ms):.3f} ms  min {min(ms):.3f}  |", {n: round(m / 3, 3) for n, (m, c) in sorted(k.items(), key=lambda kv: -kv[1][0])[:int(os.environ.get("QB_TOP", "6"))]})
