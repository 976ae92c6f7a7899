#!/usr/bin/env python
"""Small end-to-end case for compute-sanitizer: tile mode, global mode, tile cache, area ops."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import recast_rs_b200.builder as rb
from recast_rs_b200 import synth, tilecache as tc

v, t, cfgs = synth.config2(seed=1, nq=40, tile=32)
for force in ("0", "1"):
    os.environ["RCB_FORCE_GLOBAL"] = force
    ctx = rb.Context(0)
    ctx.upload_mesh(v, t)
    with ctx.build_tiles(cfgs, rb.STAGE_ALL_BUILD | rb.STAGE_ERODE, 1) as r:
        r.download(); print("mode", force, r.totals())
    ctx.close()
os.environ["RCB_FORCE_GLOBAL"] = "0"
ctx = rb.Context(0)
L, O = tc.config5(8, 16)
with ctx.build_tilecache_layers(L, O, 0.3, 0.2) as r:
    r.download(); print("tilecache", r.totals()["spans"])
ctx.upload_mesh(v, t)
with ctx.build_tiles(cfgs[:2], rb.STAGE_RASTER | rb.STAGE_FILTER) as r:
    tv = r.download().tile(0)
with ctx.apply_area_ops([cfgs[0]], tv.hf_cell_offset, tv.span_min, tv.span_max, tv.span_area,
                        [("median",), ("box", (0, -9, 0), (5, 9, 5), 7), ("cylinder", (4, 0, 4), 2.0, 9.0, 3)], stage_mask=rb.STAGE_COMPACT | rb.STAGE_DIST) as r:
    print("area ops", int((r.download().tile(0).areas == 7).sum()))
ctx.close()
print("done")
