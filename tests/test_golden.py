"""Committed golden vectors (tests/golden/*.npz, made by tests/golden/make_golden.py).
CPU: the oracle still reproduces them (regression pin).  GPU: the CUDA path reproduces them."""
import glob
import importlib.util
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
FILES = sorted(glob.glob(os.path.join(HERE, "golden", "*.npz")))
KEYS = ("span_min", "span_max", "span_area", "cell_index", "cell_count", "areas", "con", "first_conn", "conn_ref", "conn_dir",
        "conn_next", "dist")


def _mk():
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "golden", "make_golden.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_golden_files_present():
    assert len(FILES) >= 3


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f) for f in FILES])
def test_oracle_reproduces_golden(oracle, path):
    g = np.load(path)
    cfgs = _mk().unpack_cfgs(g["cfgs"])
    for i, c in enumerate(cfgs):
        hfa, chfa = oracle.build_tile(c, g["verts"], g["tris"], int(g["stage_mask"][0]), int(g["erode_radius"][0]))
        np.testing.assert_array_equal(hfa.cell_offset, g[f"t{i}_hf_cell_offset"])
        for k in KEYS:
            np.testing.assert_array_equal(getattr(chfa, k), g[f"t{i}_{k}"], err_msg=f"tile {i} {k}")
        assert [chfa.walkable_span_count, chfa.max_height, chfa.max_distance] == g[f"t{i}_scalars"].tolist()


@pytest.mark.gpu
@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f) for f in FILES])
def test_gpu_reproduces_golden(path):
    import recast_rs_b200.builder as rb

    g = np.load(path)
    cfgs = _mk().unpack_cfgs(g["cfgs"])
    ctx = rb.Context(0)
    ctx.upload_mesh(g["verts"], g["tris"])
    with ctx.build_tiles(cfgs, int(g["stage_mask"][0]), int(g["erode_radius"][0])) as r:
        r.download()
        for i in range(len(cfgs)):
            tv = r.tile(i)
            np.testing.assert_array_equal(tv.hf_cell_offset, g[f"t{i}_hf_cell_offset"])
            for k in KEYS:
                np.testing.assert_array_equal(getattr(tv, k), g[f"t{i}_{k}"], err_msg=f"tile {i} {k}")
            assert [tv.walkable_span_count, tv.max_height, tv.max_distance] == g[f"t{i}_scalars"].tolist()
            assert tv.status == 0
    ctx.close()
