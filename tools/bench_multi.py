#!/usr/bin/env python
"""rcb_multi (one process, N devices, tile groups from a shared queue) on BASELINE configs[2]: ONE 10M-triangle open world,
128x128-cell tiles.  Wall-clock per build through the product entry (host threads, queue and all), results left on the devices.
usage: bench_multi.py [ndevices] [config] [tiles_per_group]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import recast_rs_b200.builder as rb
from recast_rs_b200 import synth
from recast_rs_b200.config import configs_to_array

ndev = int(sys.argv[1]) if len(sys.argv) > 1 else torch.cuda.device_count()
cfgno = int(sys.argv[2]) if len(sys.argv) > 2 else 3
tpg = int(sys.argv[3]) if len(sys.argv) > 3 else 0
v, t, cfgs = {2: synth.config2, 3: synth.config3, 4: synth.config4}[cfgno]()
arr = configs_to_array(cfgs)
with rb.MultiContext(list(range(ndev))) as m:
    t0 = time.perf_counter()
    m.upload_mesh(v, t)
    up_ms = (time.perf_counter() - t0) * 1e3
    tot = None
    for _ in range(3):  # the first build of a group on a device measures its sizes; later ones are issued without host round trips
        with m.build_tiles(arr, rb.STAGE_ALL_BUILD, tiles_per_group=tpg) as r:
            tot = r.totals()
    ms = []
    for _ in range(7):
        for d in range(ndev):
            torch.cuda.synchronize(d)
        t0 = time.perf_counter()
        r = m.build_tiles(arr, rb.STAGE_ALL_BUILD, tiles_per_group=tpg)
        tt = r.totals()  # waits for every group
        ms.append((time.perf_counter() - t0) * 1e3)
        assert {k: x for k, x in tt.items() if k not in ("result_bytes", "tiles_per_device")} == {k: x for k, x in tot.items() if k not in ("result_bytes", "tiles_per_device")}
        last = tt["tiles_per_device"]
        r.free()
    T = t.size // 3
    print(json.dumps({"entry": "rcb_multi_build_tiles", "config": cfgno, "n_devices": ndev, "tiles": len(cfgs), "triangles": T,
                      "tiles_per_group": tpg or "default (4 groups per device)", "upload_mesh_ms": up_ms, "ms_per_build_median": float(np.median(ms)),
                      "ms_per_build_min": float(min(ms)), "triangles_per_s": T / (float(np.median(ms)) * 1e-3), "tiles_per_device_last_build": last,
                      "timing": "host wall clock around rcb_multi_build_tiles + rcb_multi_result_totals (which waits for every group)"}))
