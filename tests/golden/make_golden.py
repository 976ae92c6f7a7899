#!/usr/bin/env python
"""Generates tests/golden/*.npz — small input/output vectors of the hot path.

Provenance: the Rust reference cannot run in this image (no cargo/rustc), so these vectors are produced
by the CPU oracle (oracle/oracle.cpp), AFTER it has been pinned against the reference's own unit-test
expectations (tests/test_oracle_reference_kats.py) and cross-checked against the independent Python
restatement (tests/test_oracle_vs_pymodel.py).  They freeze the oracle (regression pin, CPU test) and
give the GPU tests fixed vectors that do not depend on regenerating anything at run time.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import oracle as orc  # noqa: E402
from recast_rs_b200 import synth  # noqa: E402
from recast_rs_b200.config import RecastConfig  # noqa: E402

CFG_FIELDS = ["width", "height", "cs", "ch", "walkable_slope_angle", "walkable_height", "walkable_climb"]


def pack_cfgs(cfgs):
    return np.array([[getattr(c, f) for f in CFG_FIELDS] + list(c.bmin) + list(c.bmax) for c in cfgs], dtype=np.float64)


def unpack_cfgs(a):
    out = []
    for row in a:
        kw = {f: (int(row[i]) if f in ("width", "height", "walkable_height", "walkable_climb") else float(np.float32(row[i])))
              for i, f in enumerate(CFG_FIELDS)}
        n = len(CFG_FIELDS)
        out.append(RecastConfig(bmin=tuple(float(np.float32(x)) for x in row[n:n + 3]),
                                bmax=tuple(float(np.float32(x)) for x in row[n + 3:n + 6]), **kw))
    return out


def build_case(name, verts, tris, cfgs, stage_mask, erode_radius):
    data = {"verts": verts.astype(np.float32), "tris": tris.astype(np.int32), "cfgs": pack_cfgs(cfgs),
            "stage_mask": np.array([stage_mask]), "erode_radius": np.array([erode_radius])}
    for i, c in enumerate(cfgs):
        hfa, chfa = orc.build_tile(c, verts, tris, stage_mask, erode_radius)
        data[f"t{i}_hf_cell_offset"] = hfa.cell_offset
        for k in ("span_min", "span_max", "span_area", "cell_index", "cell_count", "areas", "con", "first_conn", "conn_ref",
                  "conn_dir", "conn_next", "dist"):
            data[f"t{i}_{k}"] = getattr(chfa, k)
        data[f"t{i}_scalars"] = np.array([chfa.walkable_span_count, chfa.max_height, chfa.max_distance])
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **data)
    print(name, "tiles", len(cfgs), "triangles", tris.size // 3)


def main():
    sys.path.insert(0, os.path.join(os.path.dirname(HERE)))
    from parity import random_soup

    v, t, cfgs = synth.config2(seed=1, nq=24, tile=32)
    build_case("terrain_nq24_tile32", v, t, cfgs, 1 | 2 | 4 | 16, 0)
    v, t, cfgs = synth.config4(seed=2, floors=4, nq=12, tile=24)
    build_case("interior_4floors_erode1", v, t, cfgs, 1 | 2 | 4 | 8 | 16, 1)
    rng = np.random.default_rng(20261019)
    v, t = random_soup(rng, 250, 0.0, 12.0, -1.0, 7.0, small=0.3)
    cfgs = synth.tile_grid(v, 16, RecastConfig(cs=0.3, ch=0.2, walkable_height=3, walkable_climb=2))
    build_case("soup_250_tile16", v, t, cfgs, 1 | 2 | 4 | 16, 0)


if __name__ == "__main__":
    main()
