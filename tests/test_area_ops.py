"""Area operators on CompactHeightfield.areas: median filter, box / cylinder /
convex-polygon marking (area.rs:238-437, 544-618).  CPU: oracle vs the independent Python restatement;
GPU: CUDA path vs the oracle, one operator at a time through the reference-shaped methods and as one
ordered batch through rcb_apply_area_ops."""
import numpy as np
import pytest

from oracle import pymodel as pm
from recast_rs_b200.config import RecastConfig
from tests.parity import random_soup


def _scene(seed):
    rng = np.random.default_rng(seed)
    w, h, cs, ch = 14, 11, 0.5, 0.25
    bmin = (1.0, -0.5, -2.0)
    cfg = RecastConfig(width=w, height=h, cs=cs, ch=ch, bmin=bmin, bmax=(bmin[0] + w * cs, 6.0, bmin[2] + h * cs), walkable_height=2,
                       walkable_climb=2)
    v, idx = random_soup(rng, 120, 0.0, 8.0, -0.5, 4.0, small=cs)
    ops = []
    for _ in range(7):
        k = int(rng.integers(0, 4))
        cx, cz, cy = rng.uniform(1, 8), rng.uniform(-2, 3.5), rng.uniform(-0.5, 3)
        aid = int(rng.choice([0, 2, 5, 9]))
        if k == 0:
            ops.append(("median",))
        elif k == 1:
            ops.append(("box", (cx - rng.uniform(0.3, 2), cy, cz - rng.uniform(0.3, 2)), (cx + rng.uniform(0.3, 2), cy + rng.uniform(0.3, 3), cz + rng.uniform(0.3, 2)), aid))
        elif k == 2:
            ops.append(("cylinder", (cx, cy, cz), float(rng.uniform(0.4, 2.5)), float(rng.uniform(0.3, 3)), aid))
        else:
            n = int(rng.integers(3, 7))
            ang = np.sort(rng.uniform(0, 2 * np.pi, n))
            rad = rng.uniform(0.8, 2.5)
            verts = np.stack([cx + rad * np.cos(ang), np.zeros(n), cz + rad * np.sin(ang)], axis=1).astype(np.float32)
            ops.append(("poly", verts, float(cy), float(cy + rng.uniform(0.5, 3)), aid))
    return cfg, v, idx, ops


def _apply_oracle(chf, op):
    if op[0] == "median":
        chf.median_filter_walkable_area()
    elif op[0] == "box":
        chf.mark_box_area(op[1], op[2], op[3])
    elif op[0] == "cylinder":
        chf.mark_cylinder_area(op[1], op[2], op[3], op[4])
    else:
        chf.mark_convex_poly_area(op[1], op[2], op[3], op[4])


def _apply_pm(c, op):
    if op[0] == "median":
        pm.median_filter_walkable_area(c)
    elif op[0] == "box":
        pm.mark_box_area(c, op[1], op[2], op[3])
    elif op[0] == "cylinder":
        pm.mark_cylinder_area(c, op[1], op[2], op[3], op[4])
    else:
        pm.mark_convex_poly_area(c, op[1], op[2], op[3], op[4])


@pytest.mark.parametrize("seed", range(4))
def test_oracle_vs_pymodel_area_ops(oracle, seed):
    cfg, v, idx, ops = _scene(seed)
    ohf = oracle.Heightfield.from_config(cfg)
    ohf.builder_rasterize(cfg, v, idx)
    phf = pm.Heightfield(cfg.width, cfg.height, cfg.bmin, cfg.bmax, cfg.cs, cfg.ch)
    pm.builder_rasterize(phf, cfg, v, idx)
    oc, pc = oracle.CompactHeightfield.build_from_heightfield(ohf), pm.Compact(phf)
    changed = 0
    for op in ops:
        before = oc.export().areas.copy()
        _apply_oracle(oc, op)
        _apply_pm(pc, op)
        after = oc.export().areas
        assert after.tolist() == pc.areas, op[0]
        changed += int((before != after).sum())
    assert changed > 0


def _poison_ops(rng, ops):
    """One parameter of every volume becomes NaN, +-inf, +-3e38 or 0: the reference validates none of them (area.rs:298-437,
    544-618), its casts and comparisons decide."""
    vals = [np.nan, np.inf, -np.inf, 3e38, -3e38, 0.0]
    pick = lambda: float(vals[int(rng.integers(0, len(vals)))])
    out = []
    for op in ops:
        op = list(op)
        if op[0] == "box":
            a, b = list(op[1]), list(op[2])
            (a if rng.random() < 0.5 else b)[int(rng.integers(0, 3))] = pick()
            op[1], op[2] = tuple(a), tuple(b)
        elif op[0] == "cylinder":
            k = int(rng.integers(0, 3))
            if k == 0:
                p = list(op[1])
                p[int(rng.integers(0, 3))] = pick()
                op[1] = tuple(p)
            else:
                op[1 + k] = pick()
        elif op[0] == "poly":
            k = int(rng.integers(0, 3))
            if k == 0:
                vv = op[1].copy()
                vv[int(rng.integers(0, len(vv))), int(rng.choice([0, 2]))] = pick()
                op[1] = vv
            else:
                op[1 + k] = pick()
        out.append(tuple(op))
    return out


@pytest.mark.parametrize("seed", range(12))
def test_oracle_vs_pymodel_area_ops_with_non_finite_parameters(oracle, seed):
    cfg, v, idx, ops = _scene(seed)
    ops = _poison_ops(np.random.default_rng(9000 + seed), ops)
    ohf = oracle.Heightfield.from_config(cfg)
    ohf.builder_rasterize(cfg, v, idx)
    phf = pm.Heightfield(cfg.width, cfg.height, cfg.bmin, cfg.bmax, cfg.cs, cfg.ch)
    with np.errstate(all="ignore"):
        pm.builder_rasterize(phf, cfg, v, idx)
        oc, pc = oracle.CompactHeightfield.build_from_heightfield(ohf), pm.Compact(phf)
        for op in ops:
            _apply_oracle(oc, op)
            _apply_pm(pc, op)
            assert oc.export().areas.tolist() == pc.areas, op[0]


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(4))
def test_gpu_area_ops(oracle, seed):
    import recast_rs_b200.builder as rb

    cfg, v, idx, ops = _scene(10 + seed)
    ohf = oracle.Heightfield.from_config(cfg)
    ohf.builder_rasterize(cfg, v, idx)
    e = ohf.export()
    ctx = rb.Context(0)
    hf = rb.Heightfield(cfg.width, cfg.height, cfg.bmin, cfg.bmax, cfg.cs, cfg.ch, e.cell_offset, e.span_min, e.span_max, e.span_area, ctx=ctx)
    chf = rb.CompactHeightfield.build_from_heightfield(hf)
    oc = oracle.CompactHeightfield.build_from_heightfield(ohf)
    for op in ops:  # one reference-shaped call per operator
        _apply_oracle(oc, op)
        if op[0] == "median":
            rb.median_filter_walkable_area(chf)
        elif op[0] == "box":
            rb.mark_box_area(chf, op[1], op[2], op[3])
        elif op[0] == "cylinder":
            rb.mark_cylinder_area(chf, op[1], op[2], op[3], op[4])
        else:
            rb.mark_convex_poly_area(chf, op[1], len(op[1]), op[2], op[3], op[4])
        np.testing.assert_array_equal(chf.areas, oc.export().areas, err_msg=op[0])
    # the same operators as ONE ordered batch, followed by the distance field on the marked areas
    with ctx.apply_area_ops([hf._cfg()], hf.cell_offset, hf.span_min, hf.span_max, hf.span_area, ops,
                            stage_mask=rb.STAGE_COMPACT | rb.STAGE_DIST) as r:
        tv = r.download().tile(0)
    oc.build_distance_field()
    want = oc.export()
    np.testing.assert_array_equal(tv.areas, want.areas)
    np.testing.assert_array_equal(tv.dist, want.dist)
    assert tv.max_distance == want.max_distance
    ctx.close()


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(3))
def test_gpu_area_ops_with_non_finite_parameters(oracle, seed):
    import recast_rs_b200.builder as rb

    cfg, v, idx, ops = _scene(30 + seed)
    ops = _poison_ops(np.random.default_rng(9100 + seed), ops)
    ohf = oracle.Heightfield.from_config(cfg)
    ohf.builder_rasterize(cfg, v, idx)
    e = ohf.export()
    ctx = rb.Context(0)
    hf = rb.Heightfield(cfg.width, cfg.height, cfg.bmin, cfg.bmax, cfg.cs, cfg.ch, e.cell_offset, e.span_min, e.span_max, e.span_area, ctx=ctx)
    oc = oracle.CompactHeightfield.build_from_heightfield(ohf)
    for op in ops:
        _apply_oracle(oc, op)
    with ctx.apply_area_ops([hf._cfg()], hf.cell_offset, hf.span_min, hf.span_max, hf.span_area, ops, stage_mask=rb.STAGE_COMPACT) as r:
        tv = r.download().tile(0)
    np.testing.assert_array_equal(tv.areas, oc.export().areas)
    ctx.close()
