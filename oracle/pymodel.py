"""Second, independent restatement of the reference hot path in plain Python (small cases only).

TEST INFRASTRUCTURE ONLY.  Written separately from oracle.cpp, with the reference's own data
structures (dict keyed by (x,z) of singly linked span objects; connection list walked by
get_neighbor) so that a differential test oracle.cpp <-> pymodel.py catches transcription slips
in either.  Float arithmetic uses numpy.float32 scalars so every operation rounds to f32 like
the Rust code.  Citations relative to /root/reference/crates/recast/src/.
"""
from __future__ import annotations

import math

import numpy as np

f32 = np.float32
NONE = None


class Span:
    __slots__ = ("min", "max", "area", "next")

    def __init__(self, mn, mx, area, nxt=None):
        self.min, self.max, self.area, self.next = mn, mx, area, nxt


class Heightfield:
    def __init__(self, width, height, bmin, bmax, cs, ch):  # heightfield.rs:80-99
        self.width, self.height = width, height
        self.bmin = [f32(v) for v in bmin]
        self.bmax = [f32(v) for v in bmax]
        self.cs, self.ch = f32(cs), f32(ch)
        self.spans = {(x, z): None for z in range(height) for x in range(width)}

    def column(self, x, z):
        out, s = [], self.spans[(x, z)]
        while s is not None:
            out.append((s.min, s.max, s.area))
            s = s.next
        return out

    def export(self):
        off, mn, mx, ar = [0], [], [], []
        for z in range(self.height):
            for x in range(self.width):
                for (a, b, c) in self.column(x, z):
                    mn.append(a), mx.append(b), ar.append(c)
                off.append(len(mn))
        return (np.array(off, np.uint32), np.array(mn, np.int16), np.array(mx, np.int16), np.array(ar, np.uint8))

    # heightfield.rs:102-222
    def add_span(self, x, z, mn, mx, area):
        if x < 0 or x >= self.width or z < 0 or z >= self.height:
            raise ValueError("oob")
        if mn > mx:
            raise ValueError("min>max")
        new = Span(mn, mx, area)
        first = self.spans[(x, z)]
        if first is None:
            self.spans[(x, z)] = new
            return
        prev, cur = None, first
        while True:
            if cur.area == area and cur.max >= mn and mx >= cur.min:
                cur.min, cur.max = min(cur.min, mn), max(cur.max, mx)
                nx = cur.next
                if nx is not None and nx.area == area and cur.max >= nx.min:
                    cur.max = max(cur.max, nx.max)
                    cur.next = nx.next
                return
            if mx < cur.min:
                new.next = cur
                if prev is not None:
                    prev.next = new
                else:
                    self.spans[(x, z)] = new
                return
            if cur.next is None:
                cur.next = new
                return
            prev, cur = cur, cur.next


# rasterization.rs:51-168
def add_span_internal(hf, x, z, smin, smax, area, thr):
    col = hf.spans.get((x, z))
    if col is None:
        hf.spans[(x, z)] = Span(smin, smax, area)
        return
    prev, cur = None, col
    while cur is not None:
        if smax < cur.min:
            new = Span(smin, smax, area, cur)
            if prev is not None:
                prev.next = new
            else:
                hf.spans[(x, z)] = new
            return
        if smin > cur.max:
            prev, cur = cur, cur.next
            continue
        mmin, mmax = min(smin, cur.min), max(smax, cur.max)
        marea = area
        if abs(mmax - cur.max) <= thr:
            marea = max(marea, cur.area)
        nx = cur.next
        while nx is not None:
            if nx.min > mmax:
                break
            mmax = max(mmax, nx.max)
            nx = nx.next
        cur.min, cur.max, cur.area, cur.next = mmin, mmax, marea, nx
        return
    if prev is not None:
        prev.next = Span(smin, smax, area)


def add_span(hf, x, z, smin, smax, area, thr):  # rasterization.rs:23-47
    if x < 0 or z < 0 or x >= hf.width or z >= hf.height:
        return
    add_span_internal(hf, x, z, smin, smax, area, thr)


def _fmin(a, b):
    """f32::min: the other operand when one is NaN (Python's min() keeps its first argument then)."""
    if a != a:
        return b
    if b != b:
        return a
    return a if a < b else b


def _fmax(a, b):
    """f32::max, as _fmin."""
    if a != a:
        return b
    if b != b:
        return a
    return a if a > b else b


def _as_i32(v):
    v = float(v)
    if math.isnan(v):
        return 0
    return int(math.trunc(max(-2147483648.0, min(2147483647.0, v))))  # saturating, infinities included (Rust `as i32`)


def _as_i16(v):
    v = float(v)
    if math.isnan(v):
        return 0
    return int(math.trunc(max(-32768.0, min(32767.0, v))))  # saturating (Rust `as i16`)


# rasterization.rs:172-242
def divide_poly(inv, offset, axis):
    out1, out2 = [], []
    n = len(inv)
    if n == 0:
        return out1, out2
    sides = [(-1 if v[axis] < offset else (1 if v[axis] > offset else 0)) for v in inv]
    for i in range(n):
        j = (i + 1) % n
        vi, vj, si, sj = inv[i], inv[j], sides[i], sides[j]
        if si == 0:
            out1.append(vi), out2.append(vi)
        elif si < 0:
            out1.append(vi)
            if sj > 0:
                t = f32(f32(offset - vi[axis]) / f32(vj[axis] - vi[axis]))
                p = tuple(f32(vi[k] + f32(t * f32(vj[k] - vi[k]))) for k in range(3))
                out1.append(p), out2.append(p)
        else:
            out2.append(vi)
            if sj < 0:
                t = f32(f32(offset - vi[axis]) / f32(vj[axis] - vi[axis]))
                p = tuple(f32(vi[k] + f32(t * f32(vj[k] - vi[k]))) for k in range(3))
                out1.append(p), out2.append(p)
    return out1, out2


# rasterization.rs:246-366
def rasterize_tri(v0, v1, v2, area, hf, thr):
    with np.errstate(all="ignore"):
        v0, v1, v2 = (tuple(f32(c) for c in v) for v in (v0, v1, v2))
        ics, ich = f32(f32(1.0) / hf.cs), f32(f32(1.0) / hf.ch)
        tmin = [_fmin(_fmin(v0[i], v1[i]), v2[i]) for i in range(3)]  # rasterization.rs:258-264: f32::min / max
        tmax = [_fmax(_fmax(v0[i], v1[i]), v2[i]) for i in range(3)]
        for i in range(3):
            if not (tmin[i] <= hf.bmax[i] and tmax[i] >= hf.bmin[i]):
                return
        x0 = _as_i32(f32(f32(tmin[0] - hf.bmin[0]) * ics))
        x1 = _as_i32(f32(f32(tmax[0] - hf.bmin[0]) * ics))
        z0 = _as_i32(f32(f32(tmin[2] - hf.bmin[2]) * ics))
        z1 = _as_i32(f32(f32(tmax[2] - hf.bmin[2]) * ics))
        x0, x1 = max(x0, 0), min(x1, hf.width - 1)
        z0, z1 = max(z0, -1), min(z1, hf.height - 1)
        p1 = [v0, v1, v2]
        for z in range(z0, z1 + 1):
            cz = f32(hf.bmin[2] + f32(f32(z + 1) * hf.cs))
            in_row, p1 = divide_poly(p1, cz, 2)
            if len(in_row) < 3 or z < 0:
                continue
            p2 = in_row
            for x in range(x0, x1 + 1):
                cx = f32(hf.bmin[0] + f32(f32(x + 1) * hf.cs))
                cell, p2 = divide_poly(p2, cx, 0)
                if len(cell) < 3:
                    continue
                ys = [v[1] for v in cell]
                mn, mx = ys[0], ys[0]
                for y in ys[1:]:
                    mn, mx = _fmin(mn, y), _fmax(mx, y)  # :333-339
                smin = _as_i16(np.floor(f32(f32(mn - hf.bmin[1]) * ich)))
                smax = _as_i16(np.ceil(f32(f32(mx - hf.bmin[1]) * ich)))
                smin = max(smin, 0)
                add_span_internal(hf, x, z, smin, smax, area, thr)


# lib.rs:252-270 with glam 0.29.3 scalar Vec3
def classify(v0, v1, v2, thr):
    with np.errstate(all="ignore"):
        v0, v1, v2 = (tuple(f32(c) for c in v) for v in (v0, v1, v2))
        e1 = tuple(f32(v1[i] - v0[i]) for i in range(3))
        e2 = tuple(f32(v2[i] - v0[i]) for i in range(3))
        cx = f32(f32(e1[1] * e2[2]) - f32(e2[1] * e1[2]))
        cy = f32(f32(e1[2] * e2[0]) - f32(e2[2] * e1[0]))
        cz = f32(f32(e1[0] * e2[1]) - f32(e2[0] * e1[1]))
        ln = f32(np.sqrt(f32(f32(f32(cx * cx) + f32(cy * cy)) + f32(cz * cz))))
        if ln < np.finfo(np.float32).eps:
            return None
        ny = f32(cy * f32(f32(1.0) / ln))
        return 1 if abs(ny) >= thr else 0


def _i16(v):
    v &= 0xFFFF
    return v - 0x10000 if v >= 0x8000 else v


# heightfield.rs:286-328
def filter_low_hanging(hf, climb):
    for z in range(hf.height):
        for x in range(hf.width):
            prev, prev_walk, prev_area = None, False, 0
            s = hf.spans[(x, z)]
            while s is not None:
                walk = s.area != 0
                if (not walk) and prev_walk and prev is not None:
                    if s.max - prev.max <= climb:
                        s.area = prev_area
                prev_walk, prev_area, prev = walk, s.area, s
                s = s.next


# heightfield.rs:332-488
def filter_ledge(hf, height, climb):
    MAXH = 0xFFFF
    negc = _i16(-climb)
    for z in range(hf.height):
        for x in range(hf.width):
            s = hf.spans[(x, z)]
            while s is not None:
                if s.area == 0:
                    s = s.next
                    continue
                floor = s.max
                ceiling = s.next.min if s.next is not None else MAXH
                lowest = MAXH
                lo = hi = s.max
                ledge = False
                for dx, dz in ((0, -1), (1, 0), (0, 1), (-1, 0)):
                    nx, nz = x + dx, z + dz
                    if nx < 0 or nz < 0 or nx >= hf.width or nz >= hf.height:
                        lowest, ledge = negc - 1, True
                        break
                    ns = hf.spans[(nx, nz)]
                    if ns is None:
                        lowest, ledge = negc - 1, True
                        break
                    nceil = ns.min
                    if min(ceiling, nceil) - floor >= height:
                        lowest, ledge = negc - 1, True
                        break
                    while ns is not None:
                        nfloor = ns.max
                        nceil = ns.next.min if ns.next is not None else MAXH
                        if min(ceiling, nceil) - max(floor, nfloor) < height:
                            ns = ns.next
                            continue
                        d = nfloor - floor
                        lowest = min(lowest, d)
                        if abs(d) <= climb:
                            lo, hi = min(lo, nfloor), max(hi, nfloor)
                        elif d < negc:
                            break
                        ns = ns.next
                if ledge or lowest < negc or (hi - lo) > climb:
                    s.area = 0
                s = s.next


# heightfield.rs:492-531
def filter_low_height(hf, height):
    for z in range(hf.height):
        for x in range(hf.width):
            s = hf.spans[(x, z)]
            while s is not None:
                ceiling = s.next.min if s.next is not None else 0xFFFF
                if ceiling - s.max < height:
                    s.area = 0
                s = s.next


# lib.rs:186-290
def builder_rasterize(hf, cfg, verts, indices, run_filters=True):
    verts = np.asarray(verts, np.float32).reshape(-1)
    indices = np.asarray(indices, np.int64).reshape(-1)
    if verts.size == 0 or indices.size == 0:
        return
    if verts.size % 3 or indices.size % 3:
        raise ValueError("NavMeshGeneration")
    with np.errstate(all="ignore"):
        thr = f32(math.cos(float(f32(f32(f32(cfg.walkable_slope_angle) * f32(math.pi)) / f32(180.0)))))  # cosf, correctly rounded
    nv = verts.size // 3
    for t in range(indices.size // 3):
        i0, i1, i2 = (int(indices[t * 3 + k]) for k in range(3))
        if not (0 <= i0 < nv and 0 <= i1 < nv and 0 <= i2 < nv):
            raise ValueError("NavMeshGeneration")
        v0, v1, v2 = (verts[i * 3:i * 3 + 3] for i in (i0, i1, i2))
        area = classify(v0, v1, v2, thr)
        if area is None:
            continue
        rasterize_tri(v0, v1, v2, area, hf, 1)
    if run_filters:
        filter_low_hanging(hf, _i16(cfg.walkable_climb))
        filter_ledge(hf, _i16(cfg.walkable_height), _i16(cfg.walkable_climb))
        filter_low_height(hf, _i16(cfg.walkable_height))


DIRX = [-1, 0, 1, 1, 1, 0, -1, -1]
DIRZ = [-1, -1, -1, 0, 1, 1, 1, 0]


class Compact:
    """compact_heightfield.rs:175-430, 495-727; area.rs:73-234 — literal, list-walking."""

    def __init__(self, hf):
        w, h = hf.width, hf.height
        self.width, self.height = w, h
        self.bmin, self.cs, self.ch = list(hf.bmin), hf.cs, hf.ch
        self.cells, self.spans, self.areas = [], [], []
        self.max_height, self.walkable_span_count = 0, 0
        for z in range(h):
            for x in range(w):
                col = hf.column(x, z)
                if not col:
                    self.cells.append((None, 0))
                    continue
                self.cells.append((len(self.spans), len(col)))
                for (mn, mx, ar) in col:
                    self.spans.append({"min": mn, "max": mx, "area": ar, "first": None, "con": [63] * 4})
                    self.areas.append(ar)
                    self.max_height = max(self.max_height, mx & 0xFFFF)
                    if ar != 0:
                        self.walkable_span_count += 1
        self.dist = [0] * len(self.spans)
        self.max_distance = 0
        self.connections = []
        per_span = [[] for _ in self.spans]
        for i, (idx, cnt) in enumerate(self.cells):
            if idx is None:
                continue
            x, z = i % w, i // w
            for s in range(idx, idx + cnt):
                if self.areas[s] == 0:
                    continue
                sp = self.spans[s]
                for d in range(8):
                    nx, nz = x + DIRX[d], z + DIRZ[d]
                    if nx < 0 or nz < 0 or nx >= w or nz >= h:
                        continue
                    nidx, ncnt = self.cells[nz * w + nx]
                    if nidx is None:
                        continue
                    best, bestd = None, 2 ** 31 - 1
                    for ns in range(nidx, nidx + ncnt):
                        if self.areas[ns] == 0:
                            continue
                        nsp = self.spans[ns]
                        if max(nsp["min"], sp["min"]) > min(nsp["max"], sp["max"]):
                            continue
                        hd = (abs(nsp["min"] - sp["min"]) + abs(nsp["max"] - sp["max"])) & 0xFFFF
                        if hd < bestd:
                            bestd, best = hd, ns
                    if best is not None:
                        per_span[s].append((best, d))
        for s, lst in enumerate(per_span):
            first = None
            for (n, d) in reversed(lst):
                self.connections.append((n, d, first))
                first = len(self.connections) - 1
            self.spans[s]["first"] = first
            for (n, d) in lst:
                d4 = {7: 0, 5: 1, 3: 2, 1: 3}.get(d)
                if d4 is None:
                    continue
                for (cidx, ccnt) in self.cells:  # :433-442 O(cells) scan
                    if cidx is not None and cidx <= n < cidx + ccnt:
                        self.spans[s]["con"][d4] = n - cidx
                        break

    def get_neighbor(self, s, d):
        c = self.spans[s]["first"]
        while c is not None:
            n, cd, nx = self.connections[c]
            if cd == d:
                return n
            c = nx
        return None

    def _iter(self, reverse=False):
        w, h = self.width, self.height
        ys = range(h - 1, -1, -1) if reverse else range(h)
        for y in ys:
            xs = range(w - 1, -1, -1) if reverse else range(w)
            for x in xs:
                idx, cnt = self.cells[y * w + x]
                if idx is None:
                    continue
                for s in range(idx, idx + cnt):
                    yield s

    def build_distance_field(self):
        S = len(self.spans)
        src = [0xFFFF] * S
        sat = lambda a, b: min(a + b, 0xFFFF)
        for s in self._iter():
            cnt = 0
            for d in (1, 3, 5, 7):
                n = self.get_neighbor(s, d)
                if n is not None and self.areas[n] == self.areas[s]:
                    cnt += 1
            if cnt != 4:
                src[s] = 0

        def relax(s, d1, d2):
            a = self.get_neighbor(s, d1)
            if a is None:
                return
            if sat(src[a], 2) < src[s]:
                src[s] = sat(src[a], 2)
            b = self.get_neighbor(a, d2)
            if b is not None and sat(src[b], 3) < src[s]:
                src[s] = sat(src[b], 3)

        for s in self._iter():
            relax(s, 0, 3)
            relax(s, 3, 2)
        for s in self._iter(True):
            relax(s, 2, 1)
            relax(s, 1, 0)
        self.raw_dist = list(src)
        self.max_distance = max(src) if src else 0
        dst = [0] * S
        for s in self._iter():
            cd = src[s]
            if cd <= 2:
                dst[s] = cd
                continue
            d, n = cd, 1
            for dr in (1, 3, 5, 7):
                nb = self.get_neighbor(s, dr)
                if nb is None:
                    continue
                d += src[nb]
                n += 1
                dg = self.get_neighbor(nb, (dr + 1) % 8)
                if dg is not None:
                    d += src[dg]
                    n += 1
            dst[s] = (d // n) & 0xFFFF
        self.dist = dst

    def erode_walkable_area(self, radius):
        S = len(self.spans)
        dist = [0xFF] * S
        sat = lambda a, b: min(a + b, 0xFF)
        for s in self._iter():
            if self.areas[s] == 0:
                dist[s] = 0
                continue
            cnt = 0
            for d in (7, 5, 3, 1):
                n = self.get_neighbor(s, d)
                if n is None or self.areas[n] == 0:
                    break
                cnt += 1
            if cnt != 4:
                dist[s] = 0

        def relax(s, d1, d2):
            a = self.get_neighbor(s, d1)
            if a is None:
                return
            if sat(dist[a], 2) < dist[s]:
                dist[s] = sat(dist[a], 2)
            b = self.get_neighbor(a, d2)
            if b is not None and sat(dist[b], 3) < dist[s]:
                dist[s] = sat(dist[b], 3)

        for s in self._iter():
            relax(s, 0, 3)
            relax(s, 3, 2)
        for s in self._iter(True):
            relax(s, 2, 1)
            relax(s, 1, 0)
        thr = (radius * 2) & 0xFF
        for s in range(S):
            if dist[s] < thr:
                self.areas[s] = 0


# ---- tile-cache layers (detour-tilecache/src/tile_cache_builder.rs) ------------------------------------------
def _fl(v):
    return _as_i32(np.floor(v))


def _ce(v):
    return _as_i32(np.ceil(v))


def tilecache_build(layer, obstacles, cs, ch):
    """layer_to_compact_heightfield (:100-126) + rasterize_obstacles (:175-251); returns a Compact whose
    spans[i]["area"] carries the obstacle marks (chf.areas is left alone, as in the reference)."""
    cs, ch = f32(cs), f32(ch)
    w, h = layer.width, layer.height
    if len(layer.regons) != w * h or len(layer.areas) != w * h:
        raise ValueError("InvalidMesh")
    hf = Heightfield(w, h, layer.bmin, layer.bmax, cs, ch)
    for y in range(h):
        for x in range(w):
            idx = y * w + x
            if layer.regons[idx] == 0:
                continue
            hf.add_span(x, y, _i16(int(layer.hmin)), _i16(int(layer.hmin) + 1), int(layer.areas[idx]))
    c = Compact(hf)
    bmin = [f32(v) for v in layer.bmin]
    bmax = [f32(v) for v in layer.bmax]
    with np.errstate(all="ignore"):
        for o in obstacles[layer.first_obstacle:layer.first_obstacle + layer.obstacle_count]:
            if o.state in (0, 3):
                continue
            a = [f32(v) for v in o.a]
            b = [f32(v) for v in o.b]
            if o.type == 0:
                omin, omax = (f32(a[0] - b[0]), f32(a[2] - b[0])), (f32(a[0] + b[0]), f32(a[2] + b[0]))
            elif o.type == 1:
                omin, omax = (a[0], a[2]), (b[0], b[2])
            else:
                rad = f32(np.sqrt(f32(f32(b[0] * b[0]) + f32(b[2] * b[2]))))
                omin, omax = (f32(a[0] - rad), f32(a[2] - rad)), (f32(a[0] + rad), f32(a[2] + rad))
            if omax[0] < bmin[0] or omin[0] > bmax[0] or omax[1] < bmin[2] or omin[1] > bmax[2]:
                continue
            if o.type == 1:
                x0, x1 = _fl(f32(f32(a[0] - bmin[0]) / cs)), _ce(f32(f32(b[0] - bmin[0]) / cs))
                z0, z1 = _fl(f32(f32(a[2] - bmin[2]) / cs)), _ce(f32(f32(b[2] - bmin[2]) / cs))
            else:
                r = b[0] if o.type == 0 else rad
                x0, x1 = _fl(f32(f32(f32(a[0] - r) - bmin[0]) / cs)), _ce(f32(f32(f32(a[0] + r) - bmin[0]) / cs))
                z0, z1 = _fl(f32(f32(f32(a[2] - r) - bmin[2]) / cs)), _ce(f32(f32(f32(a[2] + r) - bmin[2]) / cs))
            x0, x1, z0, z1 = max(x0, 0), min(x1, w - 1), max(z0, 0), min(z1, h - 1)
            for z in range(z0, z1 + 1):
                for x in range(x0, x1 + 1):
                    if o.type == 0:
                        cx_ = f32(bmin[0] + f32(f32(f32(x) + f32(0.5)) * cs))
                        cz_ = f32(bmin[2] + f32(f32(f32(z) + f32(0.5)) * cs))
                        dx, dz = f32(cx_ - a[0]), f32(cz_ - a[2])
                        inside = f32(f32(dx * dx) + f32(dz * dz)) <= f32(b[0] * b[0])
                        lo, hi = a[1], f32(a[1] + b[1])
                    elif o.type == 1:
                        inside, lo, hi = True, a[1], b[1]
                    else:
                        cx_ = f32(f32(f32(f32(x) * cs) + bmin[0]) + f32(cs * f32(0.5)))
                        cz_ = f32(f32(f32(f32(z) * cs) + bmin[2]) + f32(cs * f32(0.5)))
                        dx, dz = f32(cx_ - a[0]), f32(cz_ - a[2])
                        cos_a = f32(f32(f32(2.0) * f32(f32(o.rot_aux[1]) + f32(0.5))) - f32(1.0))
                        sin_a = f32(f32(2.0) * f32(o.rot_aux[0]))
                        lx = f32(f32(cos_a * dx) + f32(sin_a * dz))
                        lz = f32(f32(f32(-sin_a) * dx) + f32(cos_a * dz))
                        inside = abs(lx) <= b[0] and abs(lz) <= b[2]
                        lo, hi = f32(a[1] - b[1]), f32(a[1] + b[1])
                    if not inside:
                        continue
                    idx, cnt = c.cells[z * w + x]
                    if idx is None:
                        continue
                    for s in range(idx, idx + cnt):
                        sp = c.spans[s]
                        if o.type == 0:
                            y0 = f32(bmin[1] + f32(f32(sp["min"]) * ch))
                            y1 = f32(y0 + ch)
                            hit = y0 < hi and y1 > lo
                        else:
                            y0 = f32(f32(f32(sp["min"]) * ch) + bmin[1])
                            y1 = f32(f32(f32(sp["max"]) * ch) + bmin[1])
                            hit = y1 > lo and y0 < hi
                        if hit:
                            sp["area"] = 0
    return c


# ---- area operators (area.rs) -----------------------------------------------------------------------------------
def _trunc_div(a, b, d):
    with np.errstate(all="ignore"):
        return _as_i32(f32(f32(f32(a) - f32(b)) / f32(d)))


def _point_in_poly(verts, px, pz):  # area.rs:29-50
    inside = False
    n = len(verts)
    j = n - 1
    for i in range(n):
        vi, vj = verts[i], verts[j]
        if (vi[2] > pz) == (vj[2] > pz):
            j = i
            continue
        with np.errstate(all="ignore"):
            if px < f32(f32(f32(f32(vj[0] - vi[0]) * f32(pz - vi[2])) / f32(vj[2] - vi[2])) + vi[0]):
                inside = not inside
        j = i
    return inside


def _footprint(c, minx, maxx, minz, maxz):
    w, h = c.width, c.height
    if maxx < 0 or minx >= w or maxz < 0 or minz >= h:
        return
    for z in range(max(minz, 0), min(maxz, h - 1) + 1):
        for x in range(max(minx, 0), min(maxx, w - 1) + 1):
            idx, cnt = c.cells[z * w + x]
            if idx is None:
                continue
            for s in range(idx, idx + cnt):
                yield x, z, s


def median_filter_walkable_area(c):  # area.rs:238-294
    out = [0xFF] * len(c.spans)
    for s in range(len(c.spans)):
        if c.areas[s] == 0:
            out[s] = 0
            continue
        nei = [c.areas[s]] * 9
        for i, d in enumerate((1, 3, 5, 7)):
            a = c.get_neighbor(s, d)
            if a is None:
                continue
            if c.areas[a] != 0:
                nei[i * 2] = c.areas[a]
            b = c.get_neighbor(a, (d + 1) % 8)
            if b is not None and c.areas[b] != 0:
                nei[i * 2 + 1] = c.areas[b]
        out[s] = sorted(nei)[4]
    c.areas = out


def mark_box_area(c, bmin, bmax, area_id):  # area.rs:298-355
    minx, miny, minz = _trunc_div(bmin[0], c.bmin[0], c.cs), _trunc_div(bmin[1], c.bmin[1], c.ch), _trunc_div(bmin[2], c.bmin[2], c.cs)
    maxx, maxy, maxz = _trunc_div(bmax[0], c.bmin[0], c.cs), _trunc_div(bmax[1], c.bmin[1], c.ch), _trunc_div(bmax[2], c.bmin[2], c.cs)
    for _, _, s in _footprint(c, minx, maxx, minz, maxz):
        y = c.spans[s]["min"]
        if y < miny or y > maxy or c.areas[s] == 0:
            continue
        c.areas[s] = area_id


def mark_cylinder_area(c, pos, radius, height, area_id):  # area.rs:544-618
    pos = [f32(v) for v in pos]
    radius, height = f32(radius), f32(height)
    lo = [f32(pos[0] - radius), pos[1], f32(pos[2] - radius)]
    hi = [f32(pos[0] + radius), f32(pos[1] + height), f32(pos[2] + radius)]
    minx, miny, minz = _trunc_div(lo[0], c.bmin[0], c.cs), _trunc_div(lo[1], c.bmin[1], c.ch), _trunc_div(lo[2], c.bmin[2], c.cs)
    maxx, maxy, maxz = _trunc_div(hi[0], c.bmin[0], c.cs), _trunc_div(hi[1], c.bmin[1], c.ch), _trunc_div(hi[2], c.bmin[2], c.cs)
    r2 = f32(radius * radius)
    for x, z, s in _footprint(c, minx, maxx, minz, maxz):
        cx = f32(c.bmin[0] + f32(f32(f32(x) + f32(0.5)) * c.cs))
        cz = f32(c.bmin[2] + f32(f32(f32(z) + f32(0.5)) * c.cs))
        dx, dz = f32(cx - pos[0]), f32(cz - pos[2])
        if f32(f32(dx * dx) + f32(dz * dz)) >= r2 or c.areas[s] == 0:
            continue
        y = c.spans[s]["min"]
        if miny <= y <= maxy:
            c.areas[s] = area_id


def mark_convex_poly_area(c, verts, min_y, max_y, area_id):  # area.rs:359-437
    verts = [tuple(f32(x) for x in v) for v in np.asarray(verts, np.float32).reshape(-1, 3)]
    lo = [verts[0][0], f32(min_y), verts[0][2]]
    hi = [verts[0][0], f32(max_y), verts[0][2]]
    for v in verts[1:]:
        lo[0], lo[2], hi[0], hi[2] = _fmin(lo[0], v[0]), _fmin(lo[2], v[2]), _fmax(hi[0], v[0]), _fmax(hi[2], v[2])  # f32::min / max
    minx, miny, minz = _trunc_div(lo[0], c.bmin[0], c.cs), _trunc_div(lo[1], c.bmin[1], c.ch), _trunc_div(lo[2], c.bmin[2], c.cs)
    maxx, maxy, maxz = _trunc_div(hi[0], c.bmin[0], c.cs), _trunc_div(hi[1], c.bmin[1], c.ch), _trunc_div(hi[2], c.bmin[2], c.cs)
    for x, z, s in _footprint(c, minx, maxx, minz, maxz):
        if c.areas[s] == 0:
            continue
        y = c.spans[s]["min"]
        if y < miny or y > maxy:
            continue
        px = f32(c.bmin[0] + f32(f32(f32(x) + f32(0.5)) * c.cs))
        pz = f32(c.bmin[2] + f32(f32(f32(z) + f32(0.5)) * c.cs))
        if _point_in_poly(verts, px, pz):
            c.areas[s] = area_id


# ---- heightfield wire formats (detour-dynamic/src) --------------------------------------------------
class WireError(Exception):
    """Err(..) of the strict reader: kind = 'recast' (short buffer) or 'navmesh' (add_span)."""

    def __init__(self, kind):
        super().__init__(kind)
        self.kind = kind


def voxel_serialize_spans(hf, big_endian=True):  # io/voxel_tile.rs:70-118
    order = "big" if big_endian else "little"
    out = bytearray()
    for z in range(hf.height):
        for x in range(hf.width):
            col = hf.column(x, z)
            out += (len(col) & 0xFFFF).to_bytes(2, order)
            for (mn, mx, ar) in col:
                out += (mn & 0xFFFFFFFF).to_bytes(4, order)  # i16 as u32: sign extension
                out += (mx & 0xFFFFFFFF).to_bytes(4, order)
                out += int(ar).to_bytes(4, order)
    return bytes(out)


def _wrap_i16(v):
    v &= 0xFFFF
    return v - 0x10000 if v >= 0x8000 else v


def voxel_deserialize_spans(hf, data):  # io/voxel_tile.rs:120-176 (big-endian, lenient)
    data = bytes(data)
    pos = 0
    for z in range(hf.height):
        for x in range(hf.width):
            if pos + 2 > len(data):
                break
            count = int.from_bytes(data[pos:pos + 2], "big")
            pos += 2
            for _ in range(count):
                if pos + 12 > len(data):
                    break
                mn = _wrap_i16(int.from_bytes(data[pos:pos + 4], "big"))
                mx = _wrap_i16(int.from_bytes(data[pos + 4:pos + 8], "big"))
                ar = int.from_bytes(data[pos + 8:pos + 12], "big") & 0xFF
                pos += 12
                try:
                    hf.add_span(x, z, mn, mx, ar)
                except ValueError:
                    pass  # `let _ =`


def dynamic_reconstruct_heightfield(hf, data):  # dynamic_tile.rs:147-257 (little-endian, strict)
    data = bytes(data)
    if len(data) < hf.width * hf.height * 2:
        raise WireError("recast")
    pos = 0
    for z in range(hf.height):
        for x in range(hf.width):
            if pos + 2 > len(data):
                raise WireError("recast")
            count = int.from_bytes(data[pos:pos + 2], "little")
            pos += 2
            if count == 0:
                continue
            if pos + count * 12 > len(data):
                raise WireError("recast")
            for _ in range(count):
                mn = _wrap_i16(int.from_bytes(data[pos:pos + 4], "little"))
                mx = _wrap_i16(int.from_bytes(data[pos + 4:pos + 8], "little"))
                ar = int.from_bytes(data[pos + 8:pos + 12], "little") & 0xFF
                pos += 12
                try:
                    hf.add_span(x, z, mn, mx, ar)
                except ValueError:
                    raise WireError("navmesh")


def checkpoint_clone_heightfield(src):  # checkpoint.rs:29-60
    dst = Heightfield(src.width, src.height, src.bmin, src.bmax, src.cs, src.ch)
    for z in range(src.height):
        for x in range(src.width):
            for (mn, mx, ar) in src.column(x, z):
                try:
                    dst.add_span(x, z, mn, mx, ar)
                except ValueError:
                    pass
    return dst


# ---- tile-cache front half (detour-tilecache/src) ----------------------------------------------------------
class Lz4Error(Exception):
    pass


def lz4_decompress_size_prepended(blob):
    """lz4_flex::decompress_size_prepended (0.11) restated from the LZ4 block format; see oracle.cpp."""
    blob = bytes(blob)
    if len(blob) < 4:
        raise Lz4Error("size prefix")
    cap = int.from_bytes(blob[:4], "little")
    src = blob[4:]
    n, i = len(src), 0
    out = bytearray()

    def read_integer(v):
        nonlocal i
        while True:
            if i >= n:
                raise Lz4Error("length byte")
            b = src[i]
            i += 1
            v += b
            if b != 255:
                return v

    while True:
        if i >= n:
            raise Lz4Error("token")
        token = src[i]
        i += 1
        lit = token >> 4
        if lit == 15:
            lit = read_integer(lit)
        if lit > n - i or lit > cap - len(out):
            raise Lz4Error("literals")
        out += src[i:i + lit]
        i += lit
        if i >= n:
            break
        if n - i < 2:
            raise Lz4Error("offset")
        offset = src[i] | (src[i + 1] << 8)
        i += 2
        mlen = 4 + (token & 15)
        if (token & 15) == 15:
            mlen = read_integer(mlen)
        if offset > len(out) or mlen > cap - len(out):
            raise Lz4Error("match")
        for _ in range(mlen):
            out.append(out[-offset] if offset else 0)  # offset 0 re-reads the zero fill of vec![0; size]
    return bytes(out)


def decompress_tile(blob):  # tile_cache.rs:854-872
    return b"" if len(blob) == 0 else lz4_decompress_size_prepended(blob)


def tilecache_layer_from_bytes(data):  # tile_cache_data.rs:128-232, 281-320 -> dict, or ValueError
    data = bytes(data)
    if len(data) < 66:
        raise ValueError("InvalidParam")
    u32 = lambda p: int.from_bytes(data[p:p + 4], "little")  # noqa: E731
    if u32(0) != 0x4C494554 or u32(4) != 1:
        raise ValueError("InvalidParam")
    rl, al, cl = u32(54), u32(58), u32(62)
    if 66 + rl + al + cl > len(data):
        raise ValueError("InvalidParam")
    import struct
    return {"bmin": struct.unpack_from("<3f", data, 20), "bmax": struct.unpack_from("<3f", data, 32),
            "hmin": int.from_bytes(data[44:46], "little"), "hmax": int.from_bytes(data[46:48], "little"),
            "width": data[48], "height": data[49], "regons": data[66:66 + rl], "areas": data[66 + rl:66 + rl + al],
            "cons": data[66 + rl + al:66 + rl + al + cl]}


# ---- region building: watershed.rs:364-746 (written from the Rust, independently of oracle.cpp) ---------------
RC_BORDER_REG = 0x8000
_OX, _OY = (-1, 0, 1, 0), (0, 1, 0, -1)  # compact_heightfield.rs:748-767


def _ws_nei(c, x, y, i, d):
    """neighbour through span.con[dir] as flood/expand use it (watershed.rs:173-176); None if con == 63."""
    k = c.spans[i]["con"][d]
    if k == 63:
        return None
    ax, ay = x + _OX[d], y + _OY[d]
    return ax, ay, c.cells[ay * c.width + ax][0] + k


def build_regions_watershed(c, border_size, min_region_area, merge_region_area):
    """-> (src_reg list, max_regions).  `c` is a Compact; build_distance_field is run first (:375)."""
    w, h, S = c.width, c.height, len(c.spans)
    c.build_distance_field()
    src_reg, src_dist = [0] * S, [0] * S
    NB = 8
    lvl = [[] for _ in range(NB)]
    region_id = 1
    level = (c.max_distance + 1) & ~1

    def paint(minx, maxx, miny, maxy, reg):
        for y in range(miny, maxy):
            for x in range(minx, maxx):
                ci = y * w + x
                if 0 <= ci < len(c.cells) and c.cells[ci][0] is not None:
                    for i in range(c.cells[ci][0], c.cells[ci][0] + c.cells[ci][1]):
                        if c.areas[i] != 0:
                            src_reg[i] = reg

    if border_size > 0:
        bw, bh = min(w, border_size), min(h, border_size)
        for rect in ((0, bw, 0, h), (w - bw, w, 0, h), (0, w, 0, bh), (0, w, h - bh, h)):
            paint(*rect, region_id | RC_BORDER_REG)
            region_id += 1

    def sort_cells(start_level):
        start = start_level >> 1
        for s in lvl:
            s.clear()
        for y in range(h):
            for x in range(w):
                idx, cnt = c.cells[y * w + x]
                if idx is None:
                    continue
                for i in range(idx, idx + cnt):
                    if c.areas[i] == 0 or src_reg[i] != 0:
                        continue
                    sid = max(start - (c.dist[i] >> 1), 0)
                    if sid < NB:
                        lvl[sid].append([x, y, i])

    def expand(max_iter, level, stack, fill):
        if fill:
            stack.clear()
            for y in range(h):
                for x in range(w):
                    idx, cnt = c.cells[y * w + x]
                    if idx is None:
                        continue
                    for i in range(idx, idx + cnt):
                        if c.dist[i] >= level and src_reg[i] == 0 and c.areas[i] != 0:
                            stack.append([x, y, i])
        else:
            for e in stack:
                if src_reg[e[2]] != 0:
                    e[2] = None
        it = 0
        while stack:
            failed, dirty = 0, []
            for e in stack:
                x, y, i = e
                if i is None:
                    failed += 1
                    continue
                r, d2, area = src_reg[i], 0xFFFF, c.areas[i]
                for d in range(4):
                    n = _ws_nei(c, x, y, i, d)
                    if n is None or c.areas[n[2]] != area:
                        continue
                    ai = n[2]
                    if src_reg[ai] > 0 and not (src_reg[ai] & RC_BORDER_REG) and src_dist[ai] + 2 < d2:
                        r, d2 = src_reg[ai], src_dist[ai] + 2
                if r != 0:
                    dirty.append((e, i, r, d2))
                else:
                    failed += 1
            for (e, i, r, d2) in dirty:
                e[2] = None
            for (e, i, r, d2) in dirty:
                src_reg[i], src_dist[i] = r, d2
            stack[:] = [e for e in stack if e[2] is not None]
            if failed * 3 >= len(stack) * 2:
                break
            if level > 0:
                it += 1
                if it >= max_iter:
                    break

    def flood(x, y, i, level, r):
        area = c.areas[i]
        st = [(x, y, i)]
        src_reg[i], src_dist[i] = r, 0
        lev = max(level - 2, 0)
        count = 0
        while st:
            cx, cy, ci = st.pop()
            ar = 0
            for d in range(4):
                n = _ws_nei(c, cx, cy, ci, d)
                if n is None:
                    continue
                ax, ay, ai = n
                if c.areas[ai] != area:
                    continue
                nr = src_reg[ai]
                if nr & RC_BORDER_REG:
                    continue
                if nr != 0 and nr != r:
                    ar = nr
                    break
                n2 = _ws_nei(c, ax, ay, ai, (d + 1) & 3)
                if n2 is None or c.areas[n2[2]] != area:
                    continue
                nr2 = src_reg[n2[2]]
                if nr2 != 0 and nr2 != r:
                    ar = nr2
                    break
            if ar != 0:
                src_reg[ci] = 0
                continue
            count += 1
            for d in range(4):
                n = _ws_nei(c, cx, cy, ci, d)
                if n is None:
                    continue
                ax, ay, ai = n
                if c.areas[ai] != area:
                    continue
                if c.dist[ai] >= lev and src_reg[ai] == 0:
                    src_reg[ai], src_dist[ai] = r, 0
                    st.append((ax, ay, ai))
        return count > 0

    s_id = -1
    while level > 0:
        level = max(level - 2, 0)
        s_id = (s_id + 1) & (NB - 1)
        if s_id == 0:
            sort_cells(level)
        else:
            lvl[s_id].extend([list(e) for e in lvl[s_id - 1] if e[2] is not None and src_reg[e[2]] == 0])
        expand(8, level, lvl[s_id], False)
        for e in lvl[s_id]:
            x, y, i = e
            if i is not None and src_reg[i] == 0 and flood(x, y, i, level, region_id):
                if region_id == 0xFFFF:
                    raise OverflowError("Region ID overflow")
                region_id += 1
    expand(64, 0, [], True)
    max_regions = region_id

    # merge_and_filter_regions (:543-633)
    regs = [{"n": 0, "id": i, "border": False, "conn": []} for i in range(region_id)]
    for s in range(S):
        rid = src_reg[s]
        if rid == 0 or rid & RC_BORDER_REG:
            continue
        reg = regs[rid]
        reg["n"] += 1
        for d8 in (7, 5, 3, 1):
            n = c.get_neighbor(s, d8)
            if n is None:
                continue
            nr = src_reg[n]
            if nr == rid or nr == 0:
                continue
            if nr & RC_BORDER_REG:
                reg["border"] = True
                continue
            if nr not in reg["conn"]:
                reg["conn"].append(nr)
    removals = [(i, reg["id"]) for i, reg in enumerate(regs)
                if reg["id"] != 0 and reg["n"] != 0 and reg["n"] < min_region_area and not reg["border"]]
    for (i, old) in removals:
        regs[i]["id"] = 0
        for s in range(S):
            if src_reg[s] == old:
                src_reg[s] = 0
    if merge_region_area > 0:
        while True:
            merges = []
            for i, reg in enumerate(regs):
                if reg["id"] == 0 or reg["n"] == 0:
                    continue
                if reg["n"] < merge_region_area and not reg["border"]:
                    best, best_n = 0, 0
                    for nid in reg["conn"]:
                        if nid < len(regs) and regs[nid]["id"] != 0 and regs[nid]["n"] > best_n:
                            best, best_n = regs[nid]["id"], regs[nid]["n"]
                    if best > 0:
                        merges.append((i, reg["id"], best, reg["n"]))
            for (i, old, best, n) in merges:
                for s in range(S):
                    if src_reg[s] == old:
                        src_reg[s] = best
                regs[i]["id"] = 0
                for r2 in regs:
                    if r2["id"] == best:
                        r2["n"] += n
                        break
            if not merges:
                break
    remap, new_id = [0] * len(regs), 1
    for i, reg in enumerate(regs):
        if reg["id"] != 0:
            remap[i] = new_id
            new_id += 1
    for s in range(S):
        if src_reg[s] < len(remap):
            src_reg[s] = remap[src_reg[s]]
    return src_reg, max_regions


# ---- convex volumes: convex_volume.rs:130-164, 375-430 (independent restatement) ------------------------------------
def convex_contains_point(verts, min_h, max_h, p):
    if f32(p[1]) < f32(min_h) or f32(p[1]) > f32(max_h):
        return False
    x, z = f32(p[0]), f32(p[2])
    inside, n = False, len(verts)
    j = n - 1
    for i in range(n):
        vi, vj = verts[i], verts[j]
        j = i
        if (f32(vi[2]) > z) == (f32(vj[2]) > z):
            continue
        xi = f32(f32(f32(f32(vj[0]) - f32(vi[0])) * f32(z - f32(vi[2]))) / f32(f32(vj[2]) - f32(vi[2]))) + f32(vi[0])
        if x >= f32(xi):
            continue
        inside = not inside
    return inside


def apply_convex_volumes(hf, volumes):
    """volumes: list of (verts[n][3], min_height, max_height, area_type); later entries win."""
    for z in range(hf.height):
        for x in range(hf.width):
            wx = f32(hf.bmin[0] + f32(f32(f32(x) + f32(0.5)) * hf.cs))
            wz = f32(hf.bmin[2] + f32(f32(f32(z) + f32(0.5)) * hf.cs))
            s = hf.spans[(x, z)]
            while s is not None:
                wy = f32(hf.bmin[1] + f32(f32(s.max) * hf.ch))
                for (verts, mn, mx, area) in volumes:
                    if convex_contains_point(verts, mn, mx, (wx, wy, wz)):
                        s.area = area
                s = s.next
