// oracle.cpp — CPU restatement of the recast-rs voxelization front end.
//
// TEST INFRASTRUCTURE ONLY.  This file is the parity checker and the CPU baseline: only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.  The
// product (recast_rs_b200/csrc) never links, calls or falls back to anything in oracle/.
//
// What it is: a C++17 restatement ("port") of the Rust reference's algorithm for the hot path,
// written from the cited lines with flat arrays and index-linked span lists instead of
// HashMap<(i32,i32),Option<Rc<RefCell<Span>>>>.  The Rust reference cannot be compiled in this
// image (no cargo/rustc), so this is NOT the reference binary.
//
// Parity pinning: every expectation the reference's own unit tests hold for this path
// (rasterization.rs:484-559, heightfield.rs:1206-1250,1417-1452, compact_heightfield.rs:1028-1117)
// is asserted against this oracle in tests/test_oracle_reference_kats.py.  Functions the reference
// never tests (RecastBuilder, filter_ledge_spans, filter_walkable_low_height_spans,
// build_distance_field, box_blur, erode_walkable_area, the clipper's numerics) are pinned only by
// the source text => for those rows "parity unpinned" (see DESIGN.md).  The same holds for the
// detour-dynamic wire formats at the end of this file (no reference test touches them).  The LZ4 block
// decoder (lz4_flex 0.11, not vendored) is pinned against liblz4 for valid streams only.
//
// Third-party arithmetic: glam 0.29.3 (Cargo.lock) Vec3 sub/cross/length/normalize used at
// lib.rs:253-262 is restated from the published glam source (scalar Vec3):
//   cross = (a.y*b.z - b.y*a.z, a.z*b.x - b.z*a.x, a.x*b.y - b.x*a.y); dot = (x*x + y*y) + z*z;
//   length = sqrt(dot); normalize = v * (1.0/length).
//
// Build: g++ -O2 -std=c++17 -ffp-contract=off -fno-fast-math -fPIC -shared (oracle/Makefile).
// All paths cited are relative to /root/reference/crates/recast/src/ unless stated.

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <thread>
#include <vector>

#include "../include/recast_b200.h"

namespace {

constexpr uint32_t NONE = 0xFFFFFFFFu;

// ---- Rust `as` cast semantics (saturating float->int, NaN -> 0) -----------------------------
inline int32_t f32_as_i32(float v) {
    if (std::isnan(v)) return 0;
    if (v >= 2147483648.0f) return std::numeric_limits<int32_t>::max();
    if (v <= -2147483648.0f) return std::numeric_limits<int32_t>::min();
    return (int32_t)v;  // truncates toward zero
}
inline int16_t f32_as_i16(float v) {
    if (std::isnan(v)) return 0;
    if (v >= 32767.0f) return 32767;
    if (v <= -32768.0f) return -32768;
    return (int16_t)v;
}

// ---- Heightfield (heightfield.rs:23-32, 58-99) ------------------------------------------------
struct Span {
    int16_t smin, smax;
    uint8_t area;
    uint32_t next;
};

struct Heightfield {
    int32_t width = 0, height = 0;
    float bmin[3] = {0, 0, 0}, bmax[3] = {0, 0, 0};
    float cs = 0, ch = 0;
    std::vector<uint32_t> first;  // per cell z*width+x, NONE = empty column (heightfield.rs:86)
    std::vector<Span> pool;
    bool poly_overflow = false;

    uint32_t alloc(int16_t mn, int16_t mx, uint8_t a, uint32_t nx) {
        pool.push_back(Span{mn, mx, a, nx});
        return (uint32_t)pool.size() - 1;
    }
};

// rasterization.rs:51-168 add_span_internal — the rasterizer's ordered insert + merge.
void add_span_internal(Heightfield& hf, int32_t x, int32_t z, int16_t smin, int16_t smax, uint8_t area,
                       int32_t thr) {
    uint32_t& head = hf.first[(size_t)z * hf.width + x];
    if (head == NONE) {  // :65-74
        head = hf.alloc(smin, smax, area, NONE);
        return;
    }
    uint32_t prev = NONE;
    uint32_t cur = head;
    while (cur != NONE) {
        // NB: take copies, alloc() may grow the pool
        const int16_t cmin = hf.pool[cur].smin, cmax = hf.pool[cur].smax;
        if (smax < cmin) {  // :84-100 insert before current
            uint32_t n = hf.alloc(smin, smax, area, cur);
            if (prev != NONE)
                hf.pool[prev].next = n;
            else
                hf.first[(size_t)z * hf.width + x] = n;
            return;
        }
        if (smin > cmax) {  // :102-109
            prev = cur;
            cur = hf.pool[cur].next;
            continue;
        }
        // :111-153 overlap (closed intervals): merge into `cur`
        int16_t mmin = std::min(smin, cmin);
        int16_t mmax = std::max(smax, cmax);
        uint8_t marea = area;
        if (std::abs((int32_t)mmax - (int32_t)cmax) <= thr) marea = std::max(marea, hf.pool[cur].area);  // :125-128
        uint32_t nx = hf.pool[cur].next;
        while (nx != NONE) {  // :131-145 absorb followers, their areas are dropped
            if (hf.pool[nx].smin > mmax) break;
            mmax = std::max(mmax, hf.pool[nx].smax);
            nx = hf.pool[nx].next;
        }
        Span& c = hf.pool[cur];
        c.smin = mmin;
        c.smax = mmax;
        c.area = marea;
        c.next = nx;
        return;
    }
    // :156-165 append
    if (prev != NONE) {
        uint32_t n = hf.alloc(smin, smax, area, NONE);
        hf.pool[prev].next = n;
    }
}

// heightfield.rs:102-222 Heightfield::add_span + insert_span — the second, different merge rule.
int hf_add_span2(Heightfield& hf, int32_t x, int32_t z, int16_t mn, int16_t mx, uint8_t area) {
    if (x < 0 || x >= hf.width || z < 0 || z >= hf.height) return RCB_ENAVMESH_GEN;  // :103-108
    if (mn > mx) return RCB_ENAVMESH_GEN;                                             // :110-115
    const size_t cell = (size_t)z * hf.width + x;
    if (hf.first[cell] == NONE) {  // :122-125
        hf.first[cell] = hf.alloc(mn, mx, area, NONE);
        return 0;
    }
    uint32_t prev = NONE, cur = hf.first[cell];
    for (;;) {
        Span c = hf.pool[cur];
        if (c.area == area && c.smax >= mn && mx >= c.smin) {  // :160-192 same-area overlap
            int16_t mmin = std::min(c.smin, mn), mmax = std::max(c.smax, mx);
            hf.pool[cur].smin = mmin;
            hf.pool[cur].smax = mmax;
            uint32_t nx = c.next;
            if (nx != NONE) {  // :173-189 absorb at most ONE following span
                Span n = hf.pool[nx];
                if (n.area == area && mmax >= n.smin) {
                    hf.pool[cur].smax = std::max(mmax, n.smax);
                    hf.pool[cur].next = n.next;
                }
            }
            return 0;
        }
        if (mx < c.smin) {  // :195-207 insert before current
            uint32_t n = hf.alloc(mn, mx, area, cur);
            if (prev != NONE)
                hf.pool[prev].next = n;
            else
                hf.first[cell] = n;
            return 0;
        }
        if (c.next == NONE) {  // :210-214 append
            uint32_t n = hf.alloc(mn, mx, area, NONE);
            hf.pool[cur].next = n;
            return 0;
        }
        prev = cur;  // :217-218
        cur = c.next;
    }
}

// ---- clipper ------------------------------------------------------------------------------------
constexpr int MAXV = 16;  // reference uses growable Vec; geometric bound is 7, overflow is flagged
struct Poly {
    int n = 0;
    float v[MAXV * 3];
};

// rasterization.rs:172-242 divide_poly.  out1 = side < offset, out2 = side > offset.
bool divide_poly(const Poly& in, Poly& out1, Poly& out2, float offset, int axis) {
    out1.n = 0;
    out2.n = 0;
    const int n = in.n;
    if (n == 0) return true;  // :185-187
    int side[MAXV];
    for (int i = 0; i < n; ++i) {  // :190-200
        float c = in.v[i * 3 + axis];
        side[i] = c < offset ? -1 : (c > offset ? 1 : 0);
    }
    bool ok = true;
    auto push = [&ok](Poly& p, const float* q) {
        if (p.n >= MAXV) {
            ok = false;
            return;
        }
        p.v[p.n * 3 + 0] = q[0];
        p.v[p.n * 3 + 1] = q[1];
        p.v[p.n * 3 + 2] = q[2];
        p.n++;
    };
    for (int i = 0; i < n; ++i) {  // :203-241
        const int j = (i + 1) % n;
        const float* vi = &in.v[i * 3];
        const float* vj = &in.v[j * 3];
        const int si = side[i], sj = side[j];
        if (si == 0) {
            push(out1, vi);
            push(out2, vi);
        } else if (si < 0) {
            push(out1, vi);
            if (sj > 0) {
                float t = (offset - vi[axis]) / (vj[axis] - vi[axis]);
                float p[3];
                for (int k = 0; k < 3; ++k) p[k] = vi[k] + t * (vj[k] - vi[k]);
                push(out1, p);
                push(out2, p);
            }
        } else {
            push(out2, vi);
            if (sj < 0) {
                float t = (offset - vi[axis]) / (vj[axis] - vi[axis]);
                float p[3];
                for (int k = 0; k < 3; ++k) p[k] = vi[k] + t * (vj[k] - vi[k]);
                push(out1, p);
                push(out2, p);
            }
        }
    }
    return ok;
}

// rasterization.rs:246-366 rasterize_tri (+ :370-377 overlap_bounds).
// `sink(x,z,smin,smax)` receives every fragment in emission order.
template <class Sink>
void rasterize_tri(const float* v0, const float* v1, const float* v2, const Heightfield& hf, bool& overflow,
                   Sink&& sink) {
    const float ics = 1.0f / hf.cs;  // :254-255
    const float ich = 1.0f / hf.ch;
    float tmin[3], tmax[3];
    for (int i = 0; i < 3; ++i) {  // :258-264
        tmin[i] = std::fmin(std::fmin(v0[i], v1[i]), v2[i]);
        tmax[i] = std::fmax(std::fmax(v0[i], v1[i]), v2[i]);
    }
    // :270 overlap_bounds, inclusive on all three axes
    if (!(tmin[0] <= hf.bmax[0] && tmax[0] >= hf.bmin[0] && tmin[1] <= hf.bmax[1] && tmax[1] >= hf.bmin[1] &&
          tmin[2] <= hf.bmax[2] && tmax[2] >= hf.bmin[2]))
        return;
    int32_t x0 = f32_as_i32((tmin[0] - hf.bmin[0]) * ics);  // :275-278
    int32_t x1 = f32_as_i32((tmax[0] - hf.bmin[0]) * ics);
    int32_t z0 = f32_as_i32((tmin[2] - hf.bmin[2]) * ics);
    int32_t z1 = f32_as_i32((tmax[2] - hf.bmin[2]) * ics);
    x0 = std::max(x0, 0);  // :281-285
    x1 = std::min(x1, hf.width - 1);
    z0 = std::max(z0, -1);
    z1 = std::min(z1, hf.height - 1);

    Poly p1, in_row, rest, p2, cell;
    p1.n = 3;
    for (int k = 0; k < 3; ++k) {
        p1.v[k] = v0[k];
        p1.v[3 + k] = v1[k];
        p1.v[6 + k] = v2[k];
    }
    for (int32_t z = z0; z <= z1; ++z) {  // :300-363
        const float cz_max = hf.bmin[2] + (float)(z + 1) * hf.cs;
        if (!divide_poly(p1, in_row, rest, cz_max, 2)) overflow = true;
        p1 = rest;
        if (in_row.n < 3 || z < 0) continue;
        p2 = in_row;
        for (int32_t x = x0; x <= x1; ++x) {
            const float cx_max = hf.bmin[0] + (float)(x + 1) * hf.cs;
            if (!divide_poly(p2, cell, rest, cx_max, 0)) overflow = true;
            p2 = rest;
            if (cell.n < 3) continue;
            float min_y = cell.v[1], max_y = cell.v[1];
            for (int i = 1; i < cell.n; ++i) {
                min_y = std::fmin(min_y, cell.v[i * 3 + 1]);
                max_y = std::fmax(max_y, cell.v[i * 3 + 1]);
            }
            int16_t smin = f32_as_i16(std::floor((min_y - hf.bmin[1]) * ich));  // :346-347
            int16_t smax = f32_as_i16(std::ceil((max_y - hf.bmin[1]) * ich));
            smin = std::max<int16_t>(smin, 0);  // :350
            sink(x, z, smin, smax);
        }
    }
}

void rasterize_triangle(Heightfield& hf, const float* v0, const float* v1, const float* v2, uint8_t area,
                        int32_t thr) {
    bool ovf = false;
    rasterize_tri(v0, v1, v2, hf, ovf,
                  [&](int32_t x, int32_t z, int16_t a, int16_t b) { add_span_internal(hf, x, z, a, b, area, thr); });
    if (ovf) hf.poly_overflow = true;
}

// lib.rs:212-213 — threshold computed once per call in f32 (libm cosf).
inline float slope_threshold(float angle) { return std::cos(angle * 3.14159265358979323846f / 180.0f); }

// lib.rs:252-270 — returns false when the triangle is skipped (degenerate).
inline bool classify(const float* v0, const float* v1, const float* v2, float thr, uint8_t& area) {
    const float e1x = v1[0] - v0[0], e1y = v1[1] - v0[1], e1z = v1[2] - v0[2];
    const float e2x = v2[0] - v0[0], e2y = v2[1] - v0[1], e2z = v2[2] - v0[2];
    const float cx = e1y * e2z - e2y * e1z;
    const float cy = e1z * e2x - e2z * e1x;
    const float cz = e1x * e2y - e2x * e1y;
    const float len = std::sqrt((cx * cx + cy * cy) + cz * cz);
    if (len < std::numeric_limits<float>::epsilon()) return false;  // :258
    const float ny = cy * (1.0f / len);
    area = (std::fabs(ny) >= thr) ? 1 : 0;  // :266-270
    return true;
}

// heightfield.rs:286-328
void filter_low_hanging(Heightfield& hf, int16_t climb) {
    for (size_t c = 0; c < hf.first.size(); ++c) {
        bool prev_walkable = false;
        uint8_t prev_area = 0;
        uint32_t prev = NONE;
        for (uint32_t s = hf.first[c]; s != NONE; s = hf.pool[s].next) {
            const bool walkable = hf.pool[s].area != 0;
            if (!walkable && prev_walkable && prev != NONE) {
                if ((int32_t)hf.pool[s].smax - (int32_t)hf.pool[prev].smax <= (int32_t)climb) hf.pool[s].area = prev_area;
            }
            prev_walkable = walkable;
            prev_area = hf.pool[s].area;
            prev = s;
        }
    }
}

// heightfield.rs:332-488
void filter_ledge(Heightfield& hf, int16_t height, int16_t climb) {
    constexpr int32_t MAXH = 0xffff;
    const int32_t neg_climb = (int32_t)(int16_t)(-climb);  // `-walkable_climb as i32` = (-climb) as i32
    static const int dx[4] = {0, 1, 0, -1}, dz[4] = {-1, 0, 1, 0};
    for (int32_t z = 0; z < hf.height; ++z)
        for (int32_t x = 0; x < hf.width; ++x)
            for (uint32_t s = hf.first[(size_t)z * hf.width + x]; s != NONE; s = hf.pool[s].next) {
                if (hf.pool[s].area == 0) continue;
                const int32_t floor = hf.pool[s].smax;
                const int32_t ceiling = hf.pool[s].next != NONE ? (int32_t)hf.pool[hf.pool[s].next].smin : MAXH;
                int32_t lowest_diff = MAXH;
                int32_t lo = floor, hi = floor;
                bool ledge = false;
                for (int d = 0; d < 4; ++d) {
                    const int32_t nx = x + dx[d], nz = z + dz[d];
                    if (nx < 0 || nz < 0 || nx >= hf.width || nz >= hf.height) {  // :377-385
                        lowest_diff = neg_climb - 1;
                        ledge = true;
                        break;
                    }
                    uint32_t ns = hf.first[(size_t)nz * hf.width + nx];
                    if (ns == NONE) {  // :448-453 empty neighbour column => ledge
                        lowest_diff = neg_climb - 1;
                        ledge = true;
                        break;
                    }
                    int32_t nceil = hf.pool[ns].smin;  // :393-398
                    if (std::min(ceiling, nceil) - floor >= (int32_t)height) {  // :401-406
                        lowest_diff = neg_climb - 1;
                        ledge = true;
                        break;
                    }
                    for (; ns != NONE; ns = hf.pool[ns].next) {  // :409-447
                        const int32_t nfloor = hf.pool[ns].smax;
                        nceil = hf.pool[ns].next != NONE ? (int32_t)hf.pool[hf.pool[ns].next].smin : MAXH;
                        if (std::min(ceiling, nceil) - std::max(floor, nfloor) < (int32_t)height) continue;
                        const int32_t diff = nfloor - floor;
                        lowest_diff = std::min(lowest_diff, diff);
                        if (std::abs(diff) <= (int32_t)climb) {
                            lo = std::min(lo, nfloor);
                            hi = std::max(hi, nfloor);
                        } else if (diff < neg_climb) {
                            break;
                        }
                    }
                }
                if (ledge)
                    hf.pool[s].area = 0;
                else if (lowest_diff < neg_climb)
                    hf.pool[s].area = 0;
                else if (hi - lo > (int32_t)climb)
                    hf.pool[s].area = 0;
            }
}

// heightfield.rs:492-531
void filter_low_height(Heightfield& hf, int16_t height) {
    constexpr int32_t MAXH = 0xffff;
    for (size_t c = 0; c < hf.first.size(); ++c)
        for (uint32_t s = hf.first[c]; s != NONE; s = hf.pool[s].next) {
            const int32_t floor = hf.pool[s].smax;
            const int32_t ceiling = hf.pool[s].next != NONE ? (int32_t)hf.pool[hf.pool[s].next].smin : MAXH;
            if (ceiling - floor < (int32_t)height) hf.pool[s].area = 0;
        }
}

// lib.rs:186-290 RecastBuilder::rasterize_triangles.  Returns rcb_status.
int builder_rasterize(Heightfield& hf, const rcb_config& cfg, const float* verts, size_t nfloats,
                      const int32_t* idx, size_t nidx, bool run_filters) {
    if (nfloats == 0 || nidx == 0) return 0;           // :195-197
    if (nfloats % 3 != 0) return RCB_ENAVMESH_GEN;     // :199-203
    if (nidx % 3 != 0) return RCB_ENAVMESH_GEN;        // :205-209
    const float thr = slope_threshold(cfg.walkable_slope_angle);
    const size_t ntri = nidx / 3, max_idx = nfloats / 3;
    for (size_t i = 0; i < ntri; ++i) {
        // `as usize` of a negative i32 is huge => out of bounds (:219-233)
        const uint64_t i0 = (uint64_t)(int64_t)idx[i * 3], i1 = (uint64_t)(int64_t)idx[i * 3 + 1],
                       i2 = (uint64_t)(int64_t)idx[i * 3 + 2];
        if (i0 >= max_idx || i1 >= max_idx || i2 >= max_idx) return RCB_ENAVMESH_GEN;
        const float *v0 = verts + i0 * 3, *v1 = verts + i1 * 3, *v2 = verts + i2 * 3;
        uint8_t area;
        if (!classify(v0, v1, v2, thr, area)) continue;
        rasterize_triangle(hf, v0, v1, v2, area, 1);  // :273
    }
    if (run_filters) {
        filter_low_hanging(hf, (int16_t)cfg.walkable_climb);                              // :278
        filter_ledge(hf, (int16_t)cfg.walkable_height, (int16_t)cfg.walkable_climb);      // :281-284
        filter_low_height(hf, (int16_t)cfg.walkable_height);                              // :287
    }
    return 0;
}

// ---- CompactHeightfield (compact_heightfield.rs:129-170) ----------------------------------------
struct Compact {
    int32_t width = 0, height = 0;
    float bmin[3] = {0, 0, 0}, cs = 0, ch = 0;
    std::vector<uint32_t> cell_index, cell_count;
    std::vector<int16_t> smin, smax;
    std::vector<uint8_t> span_area, areas;
    std::vector<uint16_t> con;  // 4 per span
    std::vector<uint32_t> first_conn;
    std::vector<uint32_t> conn_ref, conn_next;
    std::vector<uint8_t> conn_dir;
    std::vector<uint16_t> dist;
    uint32_t walkable_span_count = 0;
    uint16_t max_height = 0, max_distance = 0;

    // compact_heightfield.rs:495-510 get_neighbor — literal list walk
    uint32_t get_neighbor(uint32_t s, uint8_t dir) const {
        uint32_t c = first_conn[s];
        while (c != NONE) {
            if (conn_dir[c] == dir) return conn_ref[c];
            c = conn_next[c];
        }
        return NONE;
    }
};

const int DIRX[8] = {-1, 0, 1, 1, 1, 0, -1, -1};  // compact_heightfield.rs:109-111
const int DIRZ[8] = {-1, -1, -1, 0, 1, 1, 1, 0};

// compact_heightfield.rs:175-281 build_from_heightfield + :284-430 build_connections
void compact_build(const Heightfield& hf, Compact& c) {
    const int32_t w = hf.width, h = hf.height;
    c.width = w;
    c.height = h;
    for (int i = 0; i < 3; ++i) c.bmin[i] = hf.bmin[i];
    c.cs = hf.cs;
    c.ch = hf.ch;
    const size_t ncell = (size_t)w * h;
    c.cell_index.assign(ncell, NONE);
    c.cell_count.assign(ncell, 0);
    c.walkable_span_count = 0;
    c.max_height = 0;
    for (size_t i = 0; i < ncell; ++i) {  // :203-250, row-major
        if (hf.first[i] == NONE) continue;
        c.cell_index[i] = (uint32_t)c.smin.size();
        uint32_t n = 0;
        for (uint32_t s = hf.first[i]; s != NONE; s = hf.pool[s].next) {
            const Span& sp = hf.pool[s];
            c.smin.push_back(sp.smin);
            c.smax.push_back(sp.smax);
            c.span_area.push_back(sp.area);
            c.areas.push_back(sp.area);
            c.max_height = std::max(c.max_height, (uint16_t)sp.smax);  // :240 `as u16` wraps
            if (sp.area != 0) c.walkable_span_count++;
            ++n;
        }
        c.cell_count[i] = n;
    }
    const size_t S = c.smin.size();
    c.dist.assign(S, 0);  // :253
    c.con.assign(S * 4, 63);
    c.first_conn.assign(S, NONE);

    // :297-372 per walkable span, 8 directions, best overlapping walkable neighbour
    std::vector<uint32_t> nei(S * 8, NONE);
    for (size_t i = 0; i < ncell; ++i) {
        if (c.cell_index[i] == NONE) continue;
        const int32_t x = (int32_t)(i % w), z = (int32_t)(i / w);
        for (uint32_t k = 0; k < c.cell_count[i]; ++k) {
            const uint32_t s = c.cell_index[i] + k;
            if (c.areas[s] == 0) continue;
            for (int dir = 0; dir < 8; ++dir) {
                const int32_t nx = x + DIRX[dir], nz = z + DIRZ[dir];
                if (nx < 0 || nz < 0 || nx >= w || nz >= h) continue;
                const size_t nc = (size_t)nz * w + nx;
                if (c.cell_index[nc] == NONE) continue;
                uint32_t best = NONE;
                int32_t best_diff = std::numeric_limits<int32_t>::max();
                for (uint32_t nk = 0; nk < c.cell_count[nc]; ++nk) {
                    const uint32_t ns = c.cell_index[nc] + nk;
                    if (c.areas[ns] == 0) continue;
                    const int16_t bot = std::max(c.smin[ns], c.smin[s]);
                    const int16_t top = std::min(c.smax[ns], c.smax[s]);
                    if (bot > top) continue;
                    // abs_diff on i16 -> u16; the u16 sum wraps in release builds (:352-354)
                    const uint16_t d0 = (uint16_t)std::abs((int32_t)c.smin[ns] - (int32_t)c.smin[s]);
                    const uint16_t d1 = (uint16_t)std::abs((int32_t)c.smax[ns] - (int32_t)c.smax[s]);
                    const int32_t diff = (int32_t)(uint16_t)(d0 + d1);
                    if (diff < best_diff) {
                        best_diff = diff;
                        best = ns;
                    }
                }
                nei[(size_t)s * 8 + dir] = best;
            }
        }
    }
    // :375-389 connection array: spans in index order, each span's connections pushed in REVERSE
    // direction order, `next` -> previously pushed, first_connection = last pushed (lowest dir)
    for (size_t s = 0; s < S; ++s) {
        uint32_t first = NONE;
        for (int dir = 7; dir >= 0; --dir) {
            const uint32_t n = nei[s * 8 + dir];
            if (n == NONE) continue;
            c.conn_ref.push_back(n);
            c.conn_dir.push_back((uint8_t)dir);
            c.conn_next.push_back(first);
            first = (uint32_t)c.conn_ref.size() - 1;
        }
        c.first_conn[s] = first;
    }
    // :392-427 con[4]: 8-dir 7->0 (W), 5->1 (S), 3->2 (E), 1->3 (N); value = offset inside the
    // neighbour's cell.  (The reference finds that cell with an O(cells) scan, :433-442; the
    // neighbour cell is known here.)
    static const int d8of4[4] = {7, 5, 3, 1};
    for (size_t i = 0; i < ncell; ++i) {
        if (c.cell_index[i] == NONE) continue;
        const int32_t x = (int32_t)(i % w), z = (int32_t)(i / w);
        for (uint32_t k = 0; k < c.cell_count[i]; ++k) {
            const uint32_t s = c.cell_index[i] + k;
            for (int d4 = 0; d4 < 4; ++d4) {
                const int dir = d8of4[d4];
                const uint32_t n = nei[(size_t)s * 8 + dir];
                if (n == NONE) continue;
                const size_t nc = (size_t)(z + DIRZ[dir]) * w + (x + DIRX[dir]);
                c.con[(size_t)s * 4 + d4] = (uint16_t)(n - c.cell_index[nc]);
            }
        }
    }
}

inline uint16_t sat_add16(uint16_t a, uint16_t b) {
    uint32_t r = (uint32_t)a + b;
    return r > 0xffff ? 0xffff : (uint16_t)r;
}
inline uint8_t sat_add8(uint8_t a, uint8_t b) {
    uint32_t r = (uint32_t)a + b;
    return r > 0xff ? 0xff : (uint8_t)r;
}

// compact_heightfield.rs:514-543 build_distance_field, :546-675 calculate_distance_field, :679-727 box_blur
void build_distance_field(Compact& c) {
    const size_t S = c.smin.size();
    const int32_t w = c.width, h = c.height;
    std::vector<uint16_t> src(S, 0xffff), dst(S, 0);
    auto for_spans = [&](bool reverse, auto&& fn) {
        if (!reverse) {
            for (int32_t y = 0; y < h; ++y)
                for (int32_t x = 0; x < w; ++x) {
                    const size_t ci = (size_t)y * w + x;
                    if (c.cell_index[ci] == NONE) continue;
                    for (uint32_t k = 0; k < c.cell_count[ci]; ++k) fn(c.cell_index[ci] + k);
                }
        } else {
            for (int32_t y = h - 1; y >= 0; --y)
                for (int32_t x = w - 1; x >= 0; --x) {
                    const size_t ci = (size_t)y * w + x;
                    if (c.cell_index[ci] == NONE) continue;
                    for (uint32_t k = 0; k < c.cell_count[ci]; ++k) fn(c.cell_index[ci] + k);
                }
        }
    };
    // :556-585 boundary
    for_spans(false, [&](uint32_t s) {
        int cnt = 0;
        for (uint8_t dir : {1, 3, 5, 7}) {
            uint32_t n = c.get_neighbor(s, dir);
            if (n != NONE && c.areas[n] == c.areas[s]) cnt++;
        }
        if (cnt != 4) src[s] = 0;
    });
    // :588-627 forward.  The dir numbers are 4-direction numbers applied to the 8-direction table.
    for_spans(false, [&](uint32_t s) {
        uint32_t a = c.get_neighbor(s, 0);
        if (a != NONE) {
            if (sat_add16(src[a], 2) < src[s]) src[s] = sat_add16(src[a], 2);
            uint32_t aa = c.get_neighbor(a, 3);
            if (aa != NONE && sat_add16(src[aa], 3) < src[s]) src[s] = sat_add16(src[aa], 3);
        }
        uint32_t b = c.get_neighbor(s, 3);
        if (b != NONE) {
            if (sat_add16(src[b], 2) < src[s]) src[s] = sat_add16(src[b], 2);
            uint32_t bb = c.get_neighbor(b, 2);
            if (bb != NONE && sat_add16(src[bb], 3) < src[s]) src[s] = sat_add16(src[bb], 3);
        }
    });
    // :630-669 backward
    for_spans(true, [&](uint32_t s) {
        uint32_t a = c.get_neighbor(s, 2);
        if (a != NONE) {
            if (sat_add16(src[a], 2) < src[s]) src[s] = sat_add16(src[a], 2);
            uint32_t aa = c.get_neighbor(a, 1);
            if (aa != NONE && sat_add16(src[aa], 3) < src[s]) src[s] = sat_add16(src[aa], 3);
        }
        uint32_t b = c.get_neighbor(s, 1);
        if (b != NONE) {
            if (sat_add16(src[b], 2) < src[s]) src[s] = sat_add16(src[b], 2);
            uint32_t bb = c.get_neighbor(b, 0);
            if (bb != NONE && sat_add16(src[bb], 3) < src[s]) src[s] = sat_add16(src[bb], 3);
        }
    });
    uint16_t max_dist = 0;  // :672
    for (size_t i = 0; i < S; ++i) max_dist = std::max(max_dist, src[i]);
    // :679-727 box_blur(threshold 1) -> thr = 2
    const uint16_t thr = 2;
    for_spans(false, [&](uint32_t s) {
        const uint16_t cd = src[s];
        if (cd <= thr) {
            dst[s] = cd;
            return;
        }
        int32_t d = cd, n = 1;
        for (uint8_t dir : {1, 3, 5, 7}) {
            uint32_t nb = c.get_neighbor(s, dir);
            if (nb == NONE) continue;
            d += src[nb];
            n++;
            uint32_t dg = c.get_neighbor(nb, (uint8_t)((dir + 1) % 8));
            if (dg != NONE) {
                d += src[dg];
                n++;
            }
        }
        dst[s] = (uint16_t)(d / n);
    });
    c.dist = dst;  // :531-534
    c.max_distance = max_dist;
}

// area.rs:73-234 erode_walkable_area
void erode_walkable_area(Compact& c, int32_t radius) {
    const size_t S = c.smin.size();
    const int32_t w = c.width, h = c.height;
    std::vector<uint8_t> dist(S, 0xff);
    static const uint8_t d8of4[4] = {7, 5, 3, 1};  // compact_heightfield.rs:732-745
    for (int32_t z = 0; z < h; ++z)
        for (int32_t x = 0; x < w; ++x) {
            const size_t ci = (size_t)z * w + x;
            if (c.cell_index[ci] == NONE) continue;
            for (uint32_t k = 0; k < c.cell_count[ci]; ++k) {
                const uint32_t s = c.cell_index[ci] + k;
                if (c.areas[s] == 0) {  // :91-94
                    dist[s] = 0;
                    continue;
                }
                int cnt = 0;
                for (int d = 0; d < 4; ++d) {  // :100-120
                    uint32_t n = c.get_neighbor(s, d8of4[d]);
                    if (n == NONE) break;
                    if (c.areas[n] == 0) break;
                    cnt++;
                }
                if (cnt != 4) dist[s] = 0;
            }
        }
    auto relax = [&](uint32_t s, uint8_t d1, uint8_t d2) {
        uint32_t a = c.get_neighbor(s, d1);
        if (a == NONE) return;
        uint8_t nd = sat_add8(dist[a], 2);
        if (nd < dist[s]) dist[s] = nd;
        uint32_t b = c.get_neighbor(a, d2);
        if (b != NONE) {
            nd = sat_add8(dist[b], 3);
            if (nd < dist[s]) dist[s] = nd;
        }
    };
    for (int32_t z = 0; z < h; ++z)  // :132-176
        for (int32_t x = 0; x < w; ++x) {
            const size_t ci = (size_t)z * w + x;
            if (c.cell_index[ci] == NONE) continue;
            for (uint32_t k = 0; k < c.cell_count[ci]; ++k) {
                relax(c.cell_index[ci] + k, 0, 3);
                relax(c.cell_index[ci] + k, 3, 2);
            }
        }
    for (int32_t z = h - 1; z >= 0; --z)  // :179-223
        for (int32_t x = w - 1; x >= 0; --x) {
            const size_t ci = (size_t)z * w + x;
            if (c.cell_index[ci] == NONE) continue;
            for (uint32_t k = 0; k < c.cell_count[ci]; ++k) {
                relax(c.cell_index[ci] + k, 2, 1);
                relax(c.cell_index[ci] + k, 1, 0);
            }
        }
    const uint8_t thr = (uint8_t)(uint32_t)(radius * 2);  // :226 `(radius*2) as u8`
    for (size_t s = 0; s < S; ++s)
        if (dist[s] < thr) c.areas[s] = 0;  // :227-231: spans[].area untouched
}

Heightfield* hf_new(int32_t w, int32_t h, const float* bmin, const float* bmax, float cs, float ch) {
    auto* hf = new Heightfield();
    hf->width = w;
    hf->height = h;
    for (int i = 0; i < 3; ++i) {
        hf->bmin[i] = bmin[i];
        hf->bmax[i] = bmax[i];
    }
    hf->cs = cs;
    hf->ch = ch;
    hf->first.assign((size_t)std::max(w, 0) * (size_t)std::max(h, 0), NONE);
    return hf;
}

}  // namespace

// =================================================================================================
// C API for ctypes (tests/, bench.py cpu_baseline, smoke)
// =================================================================================================
extern "C" {

typedef struct orc_hf orc_hf;
typedef struct orc_chf orc_chf;

// config.rs:53-76
void orc_config_default(rcb_config* c) {
    std::memset(c, 0, sizeof(*c));
    c->cs = 0.3f;
    c->ch = 0.2f;
    c->walkable_slope_angle = 45.0f;
    c->walkable_height = 2;
    c->walkable_climb = 1;
    c->walkable_radius = 1;
    c->max_edge_len = 12;
    c->max_simplification_error = 1.3f;
    c->min_region_area = 8;
    c->merge_region_area = 20;
    c->max_vertices_per_polygon = 6;
    c->detail_sample_dist = 6.0f;
    c->detail_sample_max_error = 1.0f;
    c->border_size = 0;
}

// config.rs:85-107
void orc_config_calculate_grid_size(rcb_config* c, const float* bmin, const float* bmax) {
    for (int i = 0; i < 3; ++i) {
        c->bmin[i] = bmin[i];
        c->bmax[i] = bmax[i];
    }
    const float wf = (bmax[0] - bmin[0]) / c->cs;
    const float hf = (bmax[2] - bmin[2]) / c->cs;
    // f32::fract() = self - self.trunc()
    c->width = (wf - std::trunc(wf)) == 0.0f ? f32_as_i32(wf) + 1 : f32_as_i32(std::ceil(wf));
    c->height = (hf - std::trunc(hf)) == 0.0f ? f32_as_i32(hf) + 1 : f32_as_i32(std::ceil(hf));
}

// config.rs:110-136
int orc_config_validate(const rcb_config* c) {
    if (c->width <= 0 || c->height <= 0) return RCB_EINVALID_MESH;
    if (c->cs <= 0.0f || c->ch <= 0.0f) return RCB_EINVALID_MESH;
    if (c->walkable_slope_angle < 0.0f || c->walkable_slope_angle > 90.0f) return RCB_EINVALID_MESH;
    if (c->max_vertices_per_polygon < 3) return RCB_EINVALID_MESH;
    return 0;
}

orc_hf* orc_hf_new(int32_t w, int32_t h, const float* bmin, const float* bmax, float cs, float ch) {
    return (orc_hf*)hf_new(w, h, bmin, bmax, cs, ch);
}
void orc_hf_free(orc_hf* p) { delete (Heightfield*)p; }

// rasterization.rs:23-47 add_span (out of bounds silently ignored)
int orc_add_span(orc_hf* p, int32_t x, int32_t z, int16_t smin, int16_t smax, uint8_t area, int32_t thr) {
    Heightfield& hf = *(Heightfield*)p;
    if (x < 0 || z < 0 || x >= hf.width || z >= hf.height) return 0;
    add_span_internal(hf, x, z, smin, smax, area, thr);
    return 0;
}
// heightfield.rs:102 Heightfield::add_span
int orc_hf_add_span(orc_hf* p, int32_t x, int32_t z, int16_t mn, int16_t mx, uint8_t area) {
    return hf_add_span2(*(Heightfield*)p, x, z, mn, mx, area);
}
// rasterization.rs:381-401
int orc_rasterize_triangle(orc_hf* p, const float* v0, const float* v1, const float* v2, uint8_t area, int32_t thr) {
    rasterize_triangle(*(Heightfield*)p, v0, v1, v2, area, thr);
    return 0;
}
// rasterization.rs:405-428 (indices are trusted there; bounds-checked here to stay memory-safe)
int orc_rasterize_triangles(orc_hf* p, const float* verts, size_t nfloats, const int32_t* tris,
                            const uint8_t* areas, size_t ntris, int32_t thr) {
    Heightfield& hf = *(Heightfield*)p;
    const size_t nv = nfloats / 3;
    for (size_t i = 0; i < ntris; ++i) {
        const uint64_t a = (uint64_t)(int64_t)tris[i * 3], b = (uint64_t)(int64_t)tris[i * 3 + 1],
                       c = (uint64_t)(int64_t)tris[i * 3 + 2];
        if (a >= nv || b >= nv || c >= nv) return RCB_ENAVMESH_GEN;
        rasterize_triangle(hf, verts + a * 3, verts + b * 3, verts + c * 3, areas[i], thr);
    }
    return 0;
}
// lib.rs:186-290
int orc_builder_rasterize(orc_hf* p, const rcb_config* cfg, const float* verts, size_t nfloats,
                          const int32_t* idx, size_t nidx, int run_filters) {
    return builder_rasterize(*(Heightfield*)p, *cfg, verts, nfloats, idx, nidx, run_filters != 0);
}
void orc_filter_low_hanging_walkable_obstacles(orc_hf* p, int16_t climb) { filter_low_hanging(*(Heightfield*)p, climb); }
void orc_filter_ledge_spans(orc_hf* p, int16_t height, int16_t climb) { filter_ledge(*(Heightfield*)p, height, climb); }
void orc_filter_walkable_low_height_spans(orc_hf* p, int16_t height) { filter_low_height(*(Heightfield*)p, height); }
int orc_hf_poly_overflow(const orc_hf* p) { return ((const Heightfield*)p)->poly_overflow ? 1 : 0; }

uint64_t orc_hf_span_count(const orc_hf* p) {
    const Heightfield& hf = *(const Heightfield*)p;
    uint64_t n = 0;
    for (uint32_t f : hf.first)
        for (uint32_t s = f; s != NONE; s = hf.pool[s].next) ++n;
    return n;
}
// CSR export: cell_offset[C+1], smin/smax/area[S]
void orc_hf_export(const orc_hf* p, uint32_t* cell_offset, int16_t* smin, int16_t* smax, uint8_t* area) {
    const Heightfield& hf = *(const Heightfield*)p;
    uint32_t n = 0;
    for (size_t c = 0; c < hf.first.size(); ++c) {
        cell_offset[c] = n;
        for (uint32_t s = hf.first[c]; s != NONE; s = hf.pool[s].next) {
            smin[n] = hf.pool[s].smin;
            smax[n] = hf.pool[s].smax;
            area[n] = hf.pool[s].area;
            ++n;
        }
    }
    cell_offset[hf.first.size()] = n;
}
// Build a heightfield directly from CSR columns (no merge logic: test fixtures / round trips).
orc_hf* orc_hf_from_csr(int32_t w, int32_t h, const float* bmin, const float* bmax, float cs, float ch,
                        const uint32_t* cell_offset, const int16_t* smin, const int16_t* smax, const uint8_t* area) {
    Heightfield* hf = hf_new(w, h, bmin, bmax, cs, ch);
    const size_t C = hf->first.size();
    for (size_t c = 0; c < C; ++c) {
        uint32_t prev = NONE;
        for (uint32_t i = cell_offset[c]; i < cell_offset[c + 1]; ++i) {
            uint32_t n = hf->alloc(smin[i], smax[i], area[i], NONE);
            if (prev == NONE)
                hf->first[c] = n;
            else
                hf->pool[prev].next = n;
            prev = n;
        }
    }
    return (orc_hf*)hf;
}

orc_chf* orc_chf_build(const orc_hf* p) {
    auto* c = new Compact();
    compact_build(*(const Heightfield*)p, *c);
    return (orc_chf*)c;
}
void orc_chf_free(orc_chf* p) { delete (Compact*)p; }
void orc_chf_build_distance_field(orc_chf* p) { build_distance_field(*(Compact*)p); }
void orc_chf_erode_walkable_area(orc_chf* p, int32_t radius) { erode_walkable_area(*(Compact*)p, radius); }
uint32_t orc_chf_get_neighbor(const orc_chf* p, uint32_t span, uint8_t dir) {
    return ((const Compact*)p)->get_neighbor(span, dir);
}
// sizes: out[0]=C out[1]=S out[2]=K out[3]=walkable_span_count out[4]=max_height out[5]=max_distance
void orc_chf_sizes(const orc_chf* p, uint64_t* out) {
    const Compact& c = *(const Compact*)p;
    out[0] = c.cell_index.size();
    out[1] = c.smin.size();
    out[2] = c.conn_ref.size();
    out[3] = c.walkable_span_count;
    out[4] = c.max_height;
    out[5] = c.max_distance;
}
void orc_chf_export(const orc_chf* p, uint32_t* cell_index, uint32_t* cell_count, int16_t* smin, int16_t* smax,
                    uint8_t* span_area, uint8_t* areas, uint16_t* con, uint32_t* first_conn, uint32_t* conn_ref,
                    uint8_t* conn_dir, uint32_t* conn_next, uint16_t* dist) {
    const Compact& c = *(const Compact*)p;
    auto cp = [](auto* dst, const auto& v) {
        if (dst && !v.empty()) std::memcpy(dst, v.data(), v.size() * sizeof(v[0]));
    };
    cp(cell_index, c.cell_index);
    cp(cell_count, c.cell_count);
    cp(smin, c.smin);
    cp(smax, c.smax);
    cp(span_area, c.span_area);
    cp(areas, c.areas);
    cp(con, c.con);
    cp(first_conn, c.first_conn);
    cp(conn_ref, c.conn_ref);
    cp(conn_dir, c.conn_dir);
    cp(conn_next, c.conn_next);
    cp(dist, c.dist);
}

// -------------------------------------------------------------------------------------------------
// Batched build for the CPU baseline: for every tile runs what the reference runs,
//   Heightfield::new -> RecastBuilder::rasterize_triangles (whole mesh scanned per tile, lib.rs:217)
//   -> CompactHeightfield::build_from_heightfield -> [erode] -> build_distance_field,
// tiles distributed over `nthreads` std::threads (the reference itself is single-threaded; its
// `parallel` cargo feature is dead).  `prebinned` != 0 scans only the triangles whose
// AABB touches the tile (a courtesy the reference does not have; results are identical because
// rasterize_tri rejects the others at rasterization.rs:270).
// totals: [0]=spans [1]=connections [2]=walkable [3]=xor-fold checksum of all outputs
// Returns elapsed seconds (wall) or a negative rcb_status.
// -------------------------------------------------------------------------------------------------
double orc_build_tiles(const float* verts, size_t nfloats, const int32_t* idx, size_t nidx, const rcb_config* cfgs,
                       int ntiles, uint32_t stage_mask, int erode_radius, int nthreads, int prebinned,
                       uint64_t* totals) {
    if (nfloats % 3 != 0 || nidx % 3 != 0) return RCB_ENAVMESH_GEN;
    const size_t ntri = nidx / 3, nv = nfloats / 3;
    for (size_t i = 0; i < nidx; ++i)
        if ((uint64_t)(int64_t)idx[i] >= nv) return RCB_ENAVMESH_GEN;
    if (nthreads < 1) nthreads = 1;
    std::atomic<int> next{0};
    std::vector<uint64_t> acc((size_t)nthreads * 4, 0);
    auto t0 = std::chrono::steady_clock::now();
    auto work = [&](int tid) {
        std::vector<int32_t> local_idx;
        for (;;) {
            const int t = next.fetch_add(1);
            if (t >= ntiles) break;
            const rcb_config& cfg = cfgs[t];
            Heightfield* hf = hf_new(cfg.width, cfg.height, cfg.bmin, cfg.bmax, cfg.cs, cfg.ch);
            const int32_t* use_idx = idx;
            size_t use_n = nidx;
            if (prebinned) {
                local_idx.clear();
                for (size_t i = 0; i < ntri; ++i) {
                    const float *a = verts + (size_t)idx[i * 3] * 3, *b = verts + (size_t)idx[i * 3 + 1] * 3,
                                *c = verts + (size_t)idx[i * 3 + 2] * 3;
                    const float x0 = std::fmin(std::fmin(a[0], b[0]), c[0]), x1 = std::fmax(std::fmax(a[0], b[0]), c[0]);
                    const float z0 = std::fmin(std::fmin(a[2], b[2]), c[2]), z1 = std::fmax(std::fmax(a[2], b[2]), c[2]);
                    if (x0 <= cfg.bmax[0] && x1 >= cfg.bmin[0] && z0 <= cfg.bmax[2] && z1 >= cfg.bmin[2]) {
                        local_idx.push_back(idx[i * 3]);
                        local_idx.push_back(idx[i * 3 + 1]);
                        local_idx.push_back(idx[i * 3 + 2]);
                    }
                }
                use_idx = local_idx.data();
                use_n = local_idx.size();
            }
            if (stage_mask & RCB_STAGE_RASTER)
                builder_rasterize(*hf, cfg, verts, nfloats, use_idx, use_n, (stage_mask & RCB_STAGE_FILTER) != 0);
            uint64_t* a = &acc[(size_t)tid * 4];
            if (stage_mask & RCB_STAGE_COMPACT) {
                Compact c;
                compact_build(*hf, c);
                if (stage_mask & RCB_STAGE_ERODE) erode_walkable_area(c, erode_radius);
                if (stage_mask & RCB_STAGE_DIST) build_distance_field(c);
                a[0] += c.smin.size();
                a[1] += c.conn_ref.size();
                a[2] += c.walkable_span_count;
                uint64_t x = 0;
                for (size_t i = 0; i < c.dist.size(); ++i) x = x * 1099511628211ull + c.dist[i] + c.areas[i];
                a[3] ^= x;
            } else {
                a[0] += hf->pool.size();
            }
            delete hf;
        }
    };
    if (nthreads == 1) {
        work(0);
    } else {
        std::vector<std::thread> th;
        for (int i = 0; i < nthreads; ++i) th.emplace_back(work, i);
        for (auto& t : th) t.join();
    }
    auto t1 = std::chrono::steady_clock::now();
    if (totals) {
        totals[0] = totals[1] = totals[2] = totals[3] = 0;
        for (int i = 0; i < nthreads; ++i) {
            totals[0] += acc[(size_t)i * 4];
            totals[1] += acc[(size_t)i * 4 + 1];
            totals[2] += acc[(size_t)i * 4 + 2];
            totals[3] ^= acc[(size_t)i * 4 + 3];
        }
    }
    return std::chrono::duration<double>(t1 - t0).count();
}

int orc_hardware_concurrency(void) { return (int)std::thread::hardware_concurrency(); }

}  // extern "C"

// =================================================================================================
// Tile-cache layers (detour-tilecache/src/tile_cache_builder.rs), SURVEY §8(a) row TC
// =================================================================================================
namespace {

// rasterize_obstacles (:175-251) and the three shape rasterizers, on a Compact built from the layer.
void tc_apply_obstacle(Compact& c, const float* hbmin, const float* hbmax, float cs, float ch, const rcb_tc_obstacle& o) {
    if (o.state == RCB_OBSTACLE_EMPTY || o.state == RCB_OBSTACLE_REMOVING) return;  // :183-185
    float omin[3], omax[3];
    if (o.type == RCB_OBSTACLE_CYLINDER) {  // :192-200
        omin[0] = o.a[0] - o.b[0], omin[1] = o.a[1], omin[2] = o.a[2] - o.b[0];
        omax[0] = o.a[0] + o.b[0], omax[1] = o.a[1] + o.b[1], omax[2] = o.a[2] + o.b[0];
    } else if (o.type == RCB_OBSTACLE_BOX) {  // :201-205
        for (int i = 0; i < 3; ++i) omin[i] = o.a[i], omax[i] = o.b[i];
    } else {  // :206-224
        const float radius = std::sqrt(o.b[0] * o.b[0] + o.b[2] * o.b[2]);
        omin[0] = o.a[0] - radius, omin[1] = o.a[1] - o.b[1], omin[2] = o.a[2] - radius;
        omax[0] = o.a[0] + radius, omax[1] = o.a[1] + o.b[1], omax[2] = o.a[2] + radius;
    }
    if (omax[0] < hbmin[0] || omin[0] > hbmax[0] || omax[2] < hbmin[2] || omin[2] > hbmax[2]) return;  // :228-234
    const int32_t W = c.width, H = c.height;
    auto clampx = [&](int32_t& lo, int32_t& hi, int32_t n) {
        lo = std::max(lo, 0);
        hi = std::min(hi, n - 1);
    };
    if (o.type == RCB_OBSTACLE_CYLINDER) {  // :254-320
        const float cx = o.a[0], cy = o.a[1], cz = o.a[2], radius = o.b[0], height = o.b[1];
        int32_t min_x = f32_as_i32(std::floor((cx - radius - hbmin[0]) / cs)), max_x = f32_as_i32(std::ceil((cx + radius - hbmin[0]) / cs));
        int32_t min_z = f32_as_i32(std::floor((cz - radius - hbmin[2]) / cs)), max_z = f32_as_i32(std::ceil((cz + radius - hbmin[2]) / cs));
        clampx(min_x, max_x, W);
        clampx(min_z, max_z, H);
        for (int32_t z = min_z; z <= max_z; ++z)
            for (int32_t x = min_x; x <= max_x; ++x) {
                const float cell_x = hbmin[0] + ((float)x + 0.5f) * cs;
                const float cell_z = hbmin[2] + ((float)z + 0.5f) * cs;
                const float dx = cell_x - cx, dz = cell_z - cz;
                if (dx * dx + dz * dz <= radius * radius) {
                    const size_t ci = (size_t)z * W + x;
                    if (c.cell_index[ci] == NONE) continue;
                    for (uint32_t k = 0; k < c.cell_count[ci]; ++k) {
                        const uint32_t s = c.cell_index[ci] + k;
                        const float y0 = hbmin[1] + (float)(int32_t)c.smin[s] * ch;  // span.y == span.min
                        const float y1 = y0 + ch;
                        if (y0 < cy + height && y1 > cy) c.span_area[s] = 0;  // :307-310, CompactSpan.area only
                    }
                }
            }
    } else if (o.type == RCB_OBSTACLE_BOX) {  // :422-467
        int32_t min_x = f32_as_i32(std::floor((o.a[0] - hbmin[0]) / cs)), max_x = f32_as_i32(std::ceil((o.b[0] - hbmin[0]) / cs));
        int32_t min_z = f32_as_i32(std::floor((o.a[2] - hbmin[2]) / cs)), max_z = f32_as_i32(std::ceil((o.b[2] - hbmin[2]) / cs));
        clampx(min_x, max_x, W);
        clampx(min_z, max_z, H);
        for (int32_t z = min_z; z <= max_z; ++z)
            for (int32_t x = min_x; x <= max_x; ++x) {
                const size_t ci = (size_t)x + (size_t)z * W;
                if (c.cell_index[ci] == NONE) continue;
                for (uint32_t s = c.cell_index[ci]; s < c.cell_index[ci] + c.cell_count[ci]; ++s) {
                    const float y0 = (float)c.smin[s] * ch + hbmin[1], y1 = (float)c.smax[s] * ch + hbmin[1];
                    if (y1 > o.a[1] && y0 < o.b[1]) c.span_area[s] = 0;
                }
            }
    } else {  // :470-549
        const float chc = o.rot_aux[1] + 0.5f, chs = o.rot_aux[0];
        const float radius = std::sqrt(o.b[0] * o.b[0] + o.b[2] * o.b[2]);
        int32_t min_x = f32_as_i32(std::floor((o.a[0] - radius - hbmin[0]) / cs)), max_x = f32_as_i32(std::ceil((o.a[0] + radius - hbmin[0]) / cs));
        int32_t min_z = f32_as_i32(std::floor((o.a[2] - radius - hbmin[2]) / cs)), max_z = f32_as_i32(std::ceil((o.a[2] + radius - hbmin[2]) / cs));
        clampx(min_x, max_x, W);
        clampx(min_z, max_z, H);
        for (int32_t z = min_z; z <= max_z; ++z)
            for (int32_t x = min_x; x <= max_x; ++x) {
                const float cell_x = (float)x * cs + hbmin[0] + cs * 0.5f;
                const float cell_z = (float)z * cs + hbmin[2] + cs * 0.5f;
                const float dx = cell_x - o.a[0], dz = cell_z - o.a[2];
                const float cos_a = 2.0f * chc - 1.0f, sin_a = 2.0f * chs;
                const float lx = cos_a * dx + sin_a * dz;
                const float lz = -sin_a * dx + cos_a * dz;
                if (std::fabs(lx) <= o.b[0] && std::fabs(lz) <= o.b[2]) {
                    const size_t ci = (size_t)x + (size_t)z * W;
                    if (c.cell_index[ci] == NONE) continue;
                    for (uint32_t s = c.cell_index[ci]; s < c.cell_index[ci] + c.cell_count[ci]; ++s) {
                        const float y0 = (float)c.smin[s] * ch + hbmin[1], y1 = (float)c.smax[s] * ch + hbmin[1];
                        if (y1 > o.a[1] - o.b[1] && y0 < o.a[1] + o.b[1]) c.span_area[s] = 0;
                    }
                }
            }
    }
}

}  // namespace

extern "C" {

// layer_to_compact_heightfield (:100-126) + rasterize_obstacles (:175-251).  Returns an orc_chf (free with
// orc_chf_free) or nullptr with *status set (RCB_EINVALID_MESH / RCB_ENAVMESH_GEN).
orc_chf* orc_tilecache_build(const rcb_tc_layer* layer, const rcb_tc_obstacle* obstacles, float cs, float ch, int* status) {
    *status = 0;
    const size_t n = (size_t)layer->width * layer->height;
    if (layer->regons_len != n || layer->areas_len != n) {  // :138-143
        *status = RCB_EINVALID_MESH;
        return nullptr;
    }
    Heightfield* hf = hf_new(layer->width, layer->height, layer->bmin, layer->bmax, cs, ch);
    for (size_t y = 0; y < layer->height; ++y)
        for (size_t x = 0; x < layer->width; ++x) {  // :146-169
            const size_t idx = y * layer->width + x;
            if (layer->regons[idx] == 0) continue;
            const int32_t cell_height = (int32_t)layer->hmin;
            const int st = hf_add_span2(*hf, (int32_t)x, (int32_t)y, (int16_t)cell_height, (int16_t)(cell_height + 1), layer->areas[idx]);
            if (st) {
                *status = st;
                delete hf;
                return nullptr;
            }
        }
    auto* c = new Compact();
    compact_build(*hf, *c);
    delete hf;
    for (uint32_t i = 0; i < layer->obstacle_count; ++i)
        tc_apply_obstacle(*c, layer->bmin, layer->bmax, cs, ch, obstacles[layer->first_obstacle + i]);
    return (orc_chf*)c;
}

// batched, timed (bench cpu_baseline for config 5); returns seconds
double orc_tilecache_build_batch(const rcb_tc_layer* layers, int nlayers, const rcb_tc_obstacle* obstacles, float cs, float ch,
                                 int nthreads, uint64_t* totals) {
    if (nthreads < 1) nthreads = 1;
    std::atomic<int> next{0};
    std::vector<uint64_t> acc((size_t)nthreads * 2, 0);
    auto t0 = std::chrono::steady_clock::now();
    auto work = [&](int tid) {
        for (;;) {
            const int i = next.fetch_add(1);
            if (i >= nlayers) break;
            int st = 0;
            orc_chf* c = orc_tilecache_build(&layers[i], obstacles, cs, ch, &st);
            if (c) {
                acc[(size_t)tid * 2] += ((Compact*)c)->smin.size();
                acc[(size_t)tid * 2 + 1] += ((Compact*)c)->conn_ref.size();
                orc_chf_free(c);
            }
        }
    };
    if (nthreads == 1)
        work(0);
    else {
        std::vector<std::thread> th;
        for (int i = 0; i < nthreads; ++i) th.emplace_back(work, i);
        for (auto& t : th) t.join();
    }
    auto t1 = std::chrono::steady_clock::now();
    if (totals) {
        totals[0] = totals[1] = 0;
        for (int i = 0; i < nthreads; ++i) totals[0] += acc[(size_t)i * 2], totals[1] += acc[(size_t)i * 2 + 1];
    }
    return std::chrono::duration<double>(t1 - t0).count();
}

}  // extern "C"


// =================================================================================================
// Area operators on CompactHeightfield.areas (area.rs), SURVEY §8(f) rank 1
// =================================================================================================
namespace {

// area.rs:29-50
bool point_in_poly(uint32_t nv, const float* verts, float px, float pz) {
    bool in = false;
    for (uint32_t i = 0, j = nv - 1; i < nv; j = i++) {
        const float* vi = verts + i * 3;
        const float* vj = verts + j * 3;
        if ((vi[2] > pz) == (vj[2] > pz)) continue;
        if (px < (vj[0] - vi[0]) * (pz - vi[2]) / (vj[2] - vi[2]) + vi[0]) in = !in;
    }
    return in;
}

template <class Fn>
void for_footprint(Compact& c, int32_t minx, int32_t maxx, int32_t minz, int32_t maxz, Fn&& fn) {
    const int32_t w = c.width, h = c.height;
    if (maxx < 0 || minx >= w || maxz < 0 || minz >= h) return;
    minx = std::max(minx, 0), maxx = std::min(maxx, w - 1), minz = std::max(minz, 0), maxz = std::min(maxz, h - 1);
    for (int32_t z = minz; z <= maxz; ++z)
        for (int32_t x = minx; x <= maxx; ++x) {
            const size_t ci = (size_t)z * w + x;
            if (c.cell_index[ci] == NONE) continue;
            for (uint32_t k = 0; k < c.cell_count[ci]; ++k) fn(x, z, c.cell_index[ci] + k);
        }
}

}  // namespace

extern "C" {

// area.rs:238-294
void orc_chf_median_filter_walkable_area(orc_chf* p) {
    Compact& c = *(Compact*)p;
    const size_t S = c.smin.size();
    std::vector<uint8_t> out(S, 0xff);
    for (uint32_t s = 0; s < S; ++s) {
        if (c.areas[s] == 0) {
            out[s] = 0;
            continue;
        }
        uint8_t nei[9];
        for (int i = 0; i < 9; ++i) nei[i] = c.areas[s];
        static const uint8_t dirs[4] = {1, 3, 5, 7};
        for (int i = 0; i < 4; ++i) {
            const uint32_t a = c.get_neighbor(s, dirs[i]);
            if (a == NONE) continue;
            if (c.areas[a] != 0) nei[i * 2] = c.areas[a];
            const uint32_t b = c.get_neighbor(a, (uint8_t)((dirs[i] + 1) % 8));
            if (b != NONE && c.areas[b] != 0) nei[i * 2 + 1] = c.areas[b];
        }
        for (int i = 1; i < 9; ++i) {  // insert_sort area.rs:13-25
            const uint8_t v = nei[i];
            int j = i - 1;
            while (j >= 0 && nei[j] > v) {
                nei[j + 1] = nei[j];
                --j;
            }
            nei[j + 1] = v;
        }
        out[s] = nei[4];
    }
    c.areas = out;
}

// area.rs:298-355
void orc_chf_mark_box_area(orc_chf* p, const float* bmin, const float* bmax, uint8_t area_id) {
    Compact& c = *(Compact*)p;
    const int32_t minx = f32_as_i32((bmin[0] - c.bmin[0]) / c.cs), miny = f32_as_i32((bmin[1] - c.bmin[1]) / c.ch),
                  minz = f32_as_i32((bmin[2] - c.bmin[2]) / c.cs), maxx = f32_as_i32((bmax[0] - c.bmin[0]) / c.cs),
                  maxy = f32_as_i32((bmax[1] - c.bmin[1]) / c.ch), maxz = f32_as_i32((bmax[2] - c.bmin[2]) / c.cs);
    for_footprint(c, minx, maxx, minz, maxz, [&](int32_t, int32_t, uint32_t s) {
        const int32_t y = c.smin[s];
        if (y < miny || y > maxy) return;
        if (c.areas[s] == 0) return;
        c.areas[s] = area_id;
    });
}

// area.rs:544-618
void orc_chf_mark_cylinder_area(orc_chf* p, const float* pos, float radius, float height, uint8_t area_id) {
    Compact& c = *(Compact*)p;
    const float bbmin[3] = {pos[0] - radius, pos[1], pos[2] - radius}, bbmax[3] = {pos[0] + radius, pos[1] + height, pos[2] + radius};
    const int32_t minx = f32_as_i32((bbmin[0] - c.bmin[0]) / c.cs), miny = f32_as_i32((bbmin[1] - c.bmin[1]) / c.ch),
                  minz = f32_as_i32((bbmin[2] - c.bmin[2]) / c.cs), maxx = f32_as_i32((bbmax[0] - c.bmin[0]) / c.cs),
                  maxy = f32_as_i32((bbmax[1] - c.bmin[1]) / c.ch), maxz = f32_as_i32((bbmax[2] - c.bmin[2]) / c.cs);
    const float r2 = radius * radius;
    for_footprint(c, minx, maxx, minz, maxz, [&](int32_t x, int32_t z, uint32_t s) {
        const float cx = c.bmin[0] + ((float)x + 0.5f) * c.cs, cz = c.bmin[2] + ((float)z + 0.5f) * c.cs;
        const float dx = cx - pos[0], dz = cz - pos[2];
        if (dx * dx + dz * dz >= r2) return;
        if (c.areas[s] == 0) return;
        const int32_t y = c.smin[s];
        if (y >= miny && y <= maxy) c.areas[s] = area_id;
    });
}

// area.rs:359-437
void orc_chf_mark_convex_poly_area(orc_chf* p, const float* verts, uint32_t nv, float min_y, float max_y, uint8_t area_id) {
    Compact& c = *(Compact*)p;
    float bmin[3] = {verts[0], min_y, verts[2]}, bmax[3] = {verts[0], max_y, verts[2]};
    for (uint32_t i = 1; i < nv; ++i) {
        bmin[0] = std::fmin(bmin[0], verts[i * 3]), bmin[2] = std::fmin(bmin[2], verts[i * 3 + 2]);
        bmax[0] = std::fmax(bmax[0], verts[i * 3]), bmax[2] = std::fmax(bmax[2], verts[i * 3 + 2]);
    }
    const int32_t minx = f32_as_i32((bmin[0] - c.bmin[0]) / c.cs), miny = f32_as_i32((bmin[1] - c.bmin[1]) / c.ch),
                  minz = f32_as_i32((bmin[2] - c.bmin[2]) / c.cs), maxx = f32_as_i32((bmax[0] - c.bmin[0]) / c.cs),
                  maxy = f32_as_i32((bmax[1] - c.bmin[1]) / c.ch), maxz = f32_as_i32((bmax[2] - c.bmin[2]) / c.cs);
    for_footprint(c, minx, maxx, minz, maxz, [&](int32_t x, int32_t z, uint32_t s) {
        if (c.areas[s] == 0) return;
        const int32_t y = c.smin[s];
        if (y < miny || y > maxy) return;
        const float px = c.bmin[0] + ((float)x + 0.5f) * c.cs, pz = c.bmin[2] + ((float)z + 0.5f) * c.cs;
        if (point_in_poly(nv, verts, px, pz)) c.areas[s] = area_id;
    });
}

void orc_chf_set_areas(orc_chf* p, const uint8_t* areas) {
    Compact& c = *(Compact*)p;
    std::memcpy(c.areas.data(), areas, c.areas.size());
}

}  // extern "C"

// ======================================================================================================
// Heightfield wire formats.  Paths relative to /root/reference/crates/detour-dynamic/src/.
// The reference holds no test or golden vector for these functions: "parity unpinned" — the restatement is
// cross-checked against oracle/pymodel.py and by serialize -> deserialize round trips.
// ======================================================================================================
extern "C" {

// io/voxel_tile.rs:70-118 VoxelTile::serialize_spans.  Per cell (z-major): u16 count, then per span
// `min as u32`, `max as u32` (i16 -> u32 sign-extends), `area as u32`, everything BIG-endian.
// big_endian == 0 writes the byte-swapped twin (no writer of that form exists in the reference; it is
// what DynamicTile::reconstruct_heightfield expects to read).  out holds 2*C + 12*S bytes.
void orc_voxel_serialize_spans(const orc_hf* p, int big_endian, uint8_t* out) {
    const Heightfield& hf = *(const Heightfield*)p;
    size_t o = 0;
    auto put16 = [&](uint16_t v) {
        if (big_endian) out[o++] = (uint8_t)(v >> 8), out[o++] = (uint8_t)v;
        else out[o++] = (uint8_t)v, out[o++] = (uint8_t)(v >> 8);
    };
    auto put32 = [&](uint32_t v) {
        for (int k = 0; k < 4; ++k) out[o++] = (uint8_t)(v >> (big_endian ? 24 - 8 * k : 8 * k));
    };
    for (size_t c = 0; c < hf.first.size(); ++c) {
        uint16_t n = 0;  // :72-87 counts[index] += 1 on a u16
        for (uint32_t s = hf.first[c]; s != NONE; s = hf.pool[s].next) ++n;
        put16(n);
        for (uint32_t s = hf.first[c]; s != NONE; s = hf.pool[s].next) {
            put32((uint32_t)(int32_t)hf.pool[s].smin);
            put32((uint32_t)(int32_t)hf.pool[s].smax);
            put32((uint32_t)hf.pool[s].area);
        }
    }
}

// io/voxel_tile.rs:120-176 VoxelTile::deserialize_spans: BIG-endian, lenient.  A short buffer `break`s the
// innermost loop only (the walk then goes on from wherever it stands), add_span errors are dropped (`let _ =`).
void orc_voxel_deserialize_spans(orc_hf* p, const uint8_t* d, uint64_t len) {
    Heightfield& hf = *(Heightfield*)p;
    uint64_t pos = 0;
    auto be32 = [&](uint64_t q) { return ((uint32_t)d[q] << 24) | ((uint32_t)d[q + 1] << 16) | ((uint32_t)d[q + 2] << 8) | d[q + 3]; };
    for (int32_t z = 0; z < hf.height; ++z) {
        for (int32_t x = 0; x < hf.width; ++x) {
            if (pos + 2 > len) break;  // :129-131
            const uint32_t count = ((uint32_t)d[pos] << 8) | d[pos + 1];
            pos += 2;
            for (uint32_t i = 0; i < count; ++i) {
                if (pos + 12 > len) break;  // :138-140
                const int16_t mn = (int16_t)be32(pos), mx = (int16_t)be32(pos + 4);
                const uint8_t ar = (uint8_t)be32(pos + 8);
                pos += 12;
                (void)hf_add_span2(hf, x, z, mn, mx, ar);  // :169
            }
        }
    }
}

// dynamic_tile.rs:147-257 DynamicTile::reconstruct_heightfield: LITTLE-endian, strict.
// Returns 0, RCB_ERECAST (-6, Error::Recast: short buffer) or RCB_ENAVMESH_GEN (add_span's error, `?`).
int orc_dynamic_reconstruct_heightfield(orc_hf* p, const uint8_t* d, uint64_t len) {
    Heightfield& hf = *(Heightfield*)p;
    const uint64_t cells = (uint64_t)((int64_t)hf.width * hf.height);
    if (len < cells * 2) return -6;  // :171-175
    uint64_t pos = 0;
    auto le32 = [&](uint64_t q) { return (uint32_t)d[q] | ((uint32_t)d[q + 1] << 8) | ((uint32_t)d[q + 2] << 16) | ((uint32_t)d[q + 3] << 24); };
    for (int32_t z = 0; z < hf.height; ++z) {
        for (int32_t x = 0; x < hf.width; ++x) {
            if (pos + 2 > len) return -6;  // :181-189
            const uint64_t count = (uint32_t)d[pos] | ((uint32_t)d[pos + 1] << 8);
            pos += 2;
            if (count == 0) continue;
            if (pos + count * 12 > len) return -6;  // :207-216
            for (uint64_t i = 0; i < count; ++i) {
                const int32_t mn = (int32_t)le32(pos), mx = (int32_t)le32(pos + 4), ar = (int32_t)le32(pos + 8);
                pos += 12;
                const int st = hf_add_span2(hf, x, z, (int16_t)mn, (int16_t)mx, (uint8_t)ar);  // :253
                if (st) return st;
            }
        }
    }
    return 0;
}

// checkpoint.rs:29-60 DynamicTileCheckpoint::clone_heightfield: every span re-added through
// Heightfield::add_span in column order (errors dropped), which may merge touching same-area spans.
orc_hf* orc_checkpoint_clone_heightfield(const orc_hf* p) {
    const Heightfield& src = *(const Heightfield*)p;
    Heightfield* dst = hf_new(src.width, src.height, src.bmin, src.bmax, src.cs, src.ch);
    for (int32_t z = 0; z < src.height; ++z)
        for (int32_t x = 0; x < src.width; ++x)
            for (uint32_t s = src.first[(size_t)z * src.width + x]; s != NONE; s = src.pool[s].next)
                (void)hf_add_span2(*dst, x, z, src.pool[s].smin, src.pool[s].smax, src.pool[s].area);
    return (orc_hf*)dst;
}

}  // extern "C"

// ======================================================================================================
// Tile-cache front half: TileCache::decompress_tile + TileCacheLayer::from_bytes.
// Paths relative to /root/reference/crates/detour-tilecache/src/.
//
// decompress_tile (tile_cache.rs:854-872) is `lz4_flex::decompress_size_prepended` — lz4_flex 0.11
// (workspace Cargo.toml:64; the crate is a registry dependency, absent from /root/reference).  Restated here
// from the published format and decoder: a 4-byte little-endian uncompressed size, then ONE LZ4 block:
//   sequence = token (hi nibble literal length, lo nibble match length - 4; 15 = continue with bytes that add
//   up until one is != 255), literals, u16 little-endian offset, match copied byte by byte (overlap = run).
//   The block ends after the literals of the last sequence (input exhausted).
// lz4_flex's safe decoder allocates `vec![0; size]`, fails on: missing token / length / offset byte, literal
// run past the input, literal or match past the output size, offset beyond the bytes produced so far; and it
// returns the bytes produced (truncate), which may be FEWER than the prefix says.  Valid streams are pinned
// against liblz4 (pyarrow "lz4_raw") in tests/test_tilecache_lz4.py; the verdict on malformed streams is
// restated from the crate's source as remembered and is "parity unpinned".
// ======================================================================================================
extern "C" {

// Returns the prefixed size, or -1 when fewer than 4 bytes are present (lz4_flex: Err).
int64_t orc_lz4_prepended_size(const uint8_t* in, uint64_t len) {
    if (len < 4) return -1;
    return (int64_t)((uint32_t)in[0] | ((uint32_t)in[1] << 8) | ((uint32_t)in[2] << 16) | ((uint32_t)in[3] << 24));
}

// Decodes in[4..len) into out[0..cap) (cap = the prefixed size, zero-filled by the caller).
// Returns 0 and *produced, or RCB_EDETOUR_FAILURE (-7).
int orc_lz4_decompress_size_prepended(const uint8_t* in, uint64_t len, uint8_t* out, uint64_t cap, uint64_t* produced) {
    *produced = 0;
    if (len < 4) return -7;
    const uint8_t* ip = in + 4;
    const uint64_t n = len - 4;
    uint64_t i = 0, o = 0;
    auto read_integer = [&](uint64_t& v) {  // bytes summed until one is not 255
        for (;;) {
            if (i >= n) return false;
            const uint8_t b = ip[i++];
            v += b;
            if (b != 255) return true;
        }
    };
    for (;;) {
        if (i >= n) return -7;  // a token is expected (also: empty block)
        const uint8_t token = ip[i++];
        uint64_t lit = token >> 4;
        if (lit == 15 && !read_integer(lit)) return -7;
        if (lit > n - i) return -7;    // LiteralOutOfBounds
        if (lit > cap - o) return -7;  // OutputTooSmall
        memcpy(out + o, ip + i, lit);
        i += lit;
        o += lit;
        if (i >= n) break;  // end of block
        if (n - i < 2) return -7;
        const uint64_t offset = (uint64_t)ip[i] | ((uint64_t)ip[i + 1] << 8);
        i += 2;
        uint64_t mlen = 4 + (token & 15u);
        if ((token & 15u) == 15 && !read_integer(mlen)) return -7;
        if (offset > o) return -7;      // OffsetOutOfBounds
        if (mlen > cap - o) return -7;  // OutputTooSmall
        for (uint64_t k = 0; k < mlen; ++k) out[o + k] = out[o + k - offset];  // offset 0 re-reads the zero fill
        o += mlen;
    }
    *produced = o;
    return 0;
}

// tile_cache_data.rs:128-232 TileCacheLayerHeader::from_bytes + :281-320 TileCacheLayer::from_bytes.
// Fills the header fields of *out and points regons/areas INTO data.  Returns 0 or RCB_EDETOUR_INVALID_PARAM (-8).
int orc_tilecache_layer_from_bytes(const uint8_t* data, uint64_t len, rcb_tc_layer* out, uint32_t* cons_len) {
    if (len < 66) return -8;  // :282-285 (header 54 + 3 lengths)
    auto u32 = [&](size_t p) { return (uint32_t)data[p] | ((uint32_t)data[p + 1] << 8) | ((uint32_t)data[p + 2] << 16) | ((uint32_t)data[p + 3] << 24); };
    auto f32 = [&](size_t p) {
        const uint32_t v = u32(p);
        float f;
        memcpy(&f, &v, 4);
        return f;
    };
    if (u32(0) != 0x4C494554u || u32(4) != 1u) return -8;  // validate(): magic 'TILE', version 1 (:83-91)
    memset(out, 0, sizeof *out);
    for (int k = 0; k < 3; ++k) out->bmin[k] = f32(20 + 4 * k), out->bmax[k] = f32(32 + 4 * k);
    out->hmin = (uint16_t)(data[44] | (data[45] << 8));
    out->hmax = (uint16_t)(data[46] | (data[47] << 8));
    out->width = data[48];
    out->height = data[49];
    const uint64_t rl = u32(54), al = u32(58), cl = u32(62);
    if (66 + rl + al + cl > len) return -8;  // :311-313
    out->regons = data + 66;
    out->areas = data + 66 + rl;
    out->regons_len = (uint32_t)rl;
    out->areas_len = (uint32_t)al;
    *cons_len = (uint32_t)cl;
    return 0;
}

}  // extern "C"

extern "C" {
orc_chf* orc_tilecache_build(const rcb_tc_layer* layer, const rcb_tc_obstacle* obstacles, float cs, float ch, int* status);
void orc_chf_free(orc_chf* p);

// batched, timed (CPU baseline of the compressed tile-cache leg): per tile decompress_tile -> from_bytes ->
// layer_to_compact_heightfield + rasterize_obstacles, as TileCache::build_nav_mesh_tile does (tile_cache.rs:704-712).
// Returns seconds; totals[0] = spans, totals[1] = connections, totals[2] = tiles that failed.
double orc_tilecache_build_compressed_batch(const uint8_t* blobs, const uint64_t* off, int n, const uint32_t* first_obstacle,
                                            const uint32_t* obstacle_count, const rcb_tc_obstacle* obstacles, float cs, float ch,
                                            int nthreads, uint64_t* totals) {
    if (nthreads < 1) nthreads = 1;
    std::atomic<int> next{0};
    std::vector<uint64_t> acc((size_t)nthreads * 3, 0);
    auto t0 = std::chrono::steady_clock::now();
    auto work = [&](int tid) {
        std::vector<uint8_t> raw;
        for (;;) {
            const int i = next.fetch_add(1);
            if (i >= n) break;
            const uint8_t* b = blobs + off[i];
            const uint64_t len = off[i + 1] - off[i];
            uint64_t produced = 0;
            if (len) {
                const int64_t size = orc_lz4_prepended_size(b, len);
                if (size < 0) { ++acc[(size_t)tid * 3 + 2]; continue; }
                raw.assign((size_t)size, 0);
                if (orc_lz4_decompress_size_prepended(b, len, raw.data(), (uint64_t)size, &produced)) { ++acc[(size_t)tid * 3 + 2]; continue; }
            }
            rcb_tc_layer layer;
            uint32_t cons = 0;
            if (orc_tilecache_layer_from_bytes(raw.data(), produced, &layer, &cons)) { ++acc[(size_t)tid * 3 + 2]; continue; }
            layer.first_obstacle = first_obstacle ? first_obstacle[i] : 0;
            layer.obstacle_count = obstacle_count ? obstacle_count[i] : 0;
            int st = 0;
            orc_chf* c = orc_tilecache_build(&layer, obstacles, cs, ch, &st);
            if (!c) { ++acc[(size_t)tid * 3 + 2]; continue; }
            acc[(size_t)tid * 3] += ((Compact*)c)->smin.size();
            acc[(size_t)tid * 3 + 1] += ((Compact*)c)->conn_ref.size();
            orc_chf_free(c);
        }
    };
    if (nthreads == 1)
        work(0);
    else {
        std::vector<std::thread> th;
        for (int i = 0; i < nthreads; ++i) th.emplace_back(work, i);
        for (auto& t : th) t.join();
    }
    auto t1 = std::chrono::steady_clock::now();
    if (totals) {
        totals[0] = totals[1] = totals[2] = 0;
        for (int i = 0; i < nthreads; ++i)
            for (int k = 0; k < 3; ++k) totals[k] += acc[(size_t)i * 3 + k];
    }
    return std::chrono::duration<double>(t1 - t0).count();
}
}  // extern "C"

// ======================================================================================================
// Region building: watershed::build_regions_watershed (watershed.rs:364-510), the
// function behind CompactHeightfield::build_regions (compact_heightfield.rs:770-778) that
// RecastBuilder::build_mesh calls next (lib.rs:83-87).  Paths relative to crates/recast/src/.
// The reference holds no test for watershed.rs: "parity unpinned"; cross-checked against oracle/pymodel.py.
// ======================================================================================================
namespace {

constexpr uint16_t BORDER_REG = 0x8000;  // watershed.rs:8
struct LevelEntry {
    int32_t x, y;
    uint32_t index;  // NONE = usize::MAX marker
};

// neighbour through span.con[dir] (watershed.rs:173-176): 63 means "not connected" even when it is a real offset
inline uint32_t ws_con_neighbor(const Compact& c, int32_t x, int32_t y, uint32_t i, int dir, int32_t& ax, int32_t& ay) {
    static const int OX[4] = {-1, 0, 1, 0}, OY[4] = {0, 1, 0, -1};  // compact_heightfield.rs:748-767
    const uint16_t k = c.con[(size_t)i * 4 + dir];
    if (k == 63) return NONE;
    ax = x + OX[dir];
    ay = y + OY[dir];
    return c.cell_index[(size_t)ay * c.width + ax] + k;
}

// watershed.rs:65-119
void ws_sort_cells_by_level(uint16_t start_level, const Compact& c, const std::vector<uint16_t>& src_reg,
                            std::vector<LevelEntry>* stacks, int nb_stacks, int log_levels) {
    start_level = (uint16_t)(start_level >> log_levels);
    for (int s = 0; s < nb_stacks; ++s) stacks[s].clear();
    for (int32_t y = 0; y < c.height; ++y)
        for (int32_t x = 0; x < c.width; ++x) {
            const size_t ci = (size_t)y * c.width + x;
            if (c.cell_index[ci] == NONE) continue;
            for (uint32_t s = 0; s < c.cell_count[ci]; ++s) {
                const uint32_t i = c.cell_index[ci] + s;
                if (c.areas[i] == 0 || src_reg[i] != 0) continue;
                const uint16_t level = (uint16_t)(c.dist[i] >> log_levels);
                const uint32_t sid = start_level > level ? (uint32_t)(start_level - level) : 0u;  // saturating_sub
                if ((int)sid < nb_stacks) stacks[sid].push_back(LevelEntry{x, y, i});
            }
        }
}

// watershed.rs:136-242
bool ws_flood_region(int32_t x, int32_t y, uint32_t i, uint16_t level, uint16_t r, const Compact& c, std::vector<uint16_t>& src_reg,
                     std::vector<uint16_t>& src_dist, std::vector<LevelEntry>& stack) {
    const uint8_t area = c.areas[i];
    stack.clear();
    stack.push_back(LevelEntry{x, y, i});
    src_reg[i] = r;
    src_dist[i] = 0;
    const uint16_t lev = level >= 2 ? (uint16_t)(level - 2) : 0;
    int count = 0;
    while (!stack.empty()) {
        const LevelEntry back = stack.back();
        stack.pop_back();
        const int32_t cx = back.x, cy = back.y;
        const uint32_t ci = back.index;
        uint16_t ar = 0;
        for (int dir = 0; dir < 4; ++dir) {
            int32_t ax, ay;
            const uint32_t ai = ws_con_neighbor(c, cx, cy, ci, dir, ax, ay);
            if (ai == NONE) continue;
            if (c.areas[ai] != area) continue;
            const uint16_t nr = src_reg[ai];
            if (nr & BORDER_REG) continue;
            if (nr != 0 && nr != r) {
                ar = nr;
                break;
            }
            const int dir2 = (dir + 1) & 3;
            int32_t ax2, ay2;
            const uint32_t ai2 = ws_con_neighbor(c, ax, ay, ai, dir2, ax2, ay2);
            if (ai2 == NONE) continue;
            if (c.areas[ai2] != area) continue;
            const uint16_t nr2 = src_reg[ai2];
            if (nr2 != 0 && nr2 != r) {
                ar = nr2;
                break;
            }
        }
        if (ar != 0) {
            src_reg[ci] = 0;
            continue;
        }
        ++count;
        for (int dir = 0; dir < 4; ++dir) {
            int32_t ax, ay;
            const uint32_t ai = ws_con_neighbor(c, cx, cy, ci, dir, ax, ay);
            if (ai == NONE) continue;
            if (c.areas[ai] != area) continue;
            if (c.dist[ai] >= lev && src_reg[ai] == 0) {
                src_reg[ai] = r;
                src_dist[ai] = 0;
                stack.push_back(LevelEntry{ax, ay, ai});
            }
        }
    }
    return count > 0;
}

// watershed.rs:245-361
void ws_expand_regions(int max_iter, uint16_t level, const Compact& c, std::vector<uint16_t>& src_reg, std::vector<uint16_t>& src_dist,
                       std::vector<LevelEntry>& stack, bool fill_stack) {
    if (fill_stack) {
        stack.clear();
        for (int32_t y = 0; y < c.height; ++y)
            for (int32_t x = 0; x < c.width; ++x) {
                const size_t ci = (size_t)y * c.width + x;
                if (c.cell_index[ci] == NONE) continue;
                for (uint32_t s = 0; s < c.cell_count[ci]; ++s) {
                    const uint32_t i = c.cell_index[ci] + s;
                    if (c.dist[i] >= level && src_reg[i] == 0 && c.areas[i] != 0) stack.push_back(LevelEntry{x, y, i});
                }
            }
    } else {
        for (auto& e : stack)
            if (src_reg[e.index] != 0) e.index = NONE;
    }
    int iter = 0;
    struct Dirty {
        uint32_t idx;
        uint16_t reg, dist;
    };
    std::vector<Dirty> dirty;
    std::vector<size_t> marks;
    while (!stack.empty()) {
        size_t failed = 0;
        dirty.clear();
        marks.clear();
        for (size_t j = 0; j < stack.size(); ++j) {
            const LevelEntry& e = stack[j];
            const uint32_t i = e.index;
            if (i == NONE) {
                ++failed;
                continue;
            }
            uint16_t r = src_reg[i];
            uint16_t d2 = 0xffff;
            const uint8_t area = c.areas[i];
            for (int dir = 0; dir < 4; ++dir) {
                int32_t ax, ay;
                const uint32_t ai = ws_con_neighbor(c, e.x, e.y, i, dir, ax, ay);
                if (ai == NONE) continue;
                if (c.areas[ai] != area) continue;
                if (src_reg[ai] > 0 && (src_reg[ai] & BORDER_REG) == 0 && (uint16_t)(src_dist[ai] + 2) < d2) {
                    r = src_reg[ai];
                    d2 = (uint16_t)(src_dist[ai] + 2);
                }
            }
            if (r != 0) {
                marks.push_back(j);
                dirty.push_back(Dirty{i, r, d2});
            } else {
                ++failed;
            }
        }
        for (size_t j : marks) stack[j].index = NONE;
        for (const Dirty& d : dirty) {
            src_reg[d.idx] = d.reg;
            src_dist[d.idx] = d.dist;
        }
        stack.erase(std::remove_if(stack.begin(), stack.end(), [](const LevelEntry& e) { return e.index == NONE; }), stack.end());
        if (failed * 3 >= stack.size() * 2) break;  // :350 — compared with the length AFTER the retain
        if (level > 0) {
            ++iter;
            if (iter >= max_iter) break;
        }
    }
}

void ws_paint_rect(int32_t minx, int32_t maxx, int32_t miny, int32_t maxy, uint16_t reg, const Compact& c, std::vector<uint16_t>& src_reg) {
    for (int32_t y = miny; y < maxy; ++y)
        for (int32_t x = minx; x < maxx; ++x) {
            const long long ci = (long long)y * c.width + x;  // watershed.rs:524-526: the bound is on the flat index
            if (ci < 0 || ci >= (long long)c.cell_index.size()) continue;
            if (c.cell_index[(size_t)ci] == NONE) continue;
            for (uint32_t s = 0; s < c.cell_count[(size_t)ci]; ++s) {
                const uint32_t i = c.cell_index[(size_t)ci] + s;
                if (c.areas[i] != 0) src_reg[i] = reg;
            }
        }
}

struct WsRegion {
    int32_t span_count = 0;
    uint16_t id = 0;
    bool connects_to_border = false;
    std::vector<uint32_t> connections;
};

// watershed.rs:543-633, 636-716, 719-746
void ws_merge_and_filter(int32_t min_region_area, int32_t merge_region_area, uint16_t max_region_id, const Compact& c,
                         std::vector<uint16_t>& src_reg) {
    static const uint8_t DIR8[4] = {7, 5, 3, 1};  // compact_heightfield.rs:737-743
    std::vector<WsRegion> regions(max_region_id);
    for (uint16_t i = 0; i < max_region_id; ++i) regions[i].id = i;
    const size_t S = c.smin.size();
    for (size_t s = 0; s < S; ++s) {
        const uint16_t rid = src_reg[s];
        if (rid == 0 || (rid & BORDER_REG)) continue;
        WsRegion& reg = regions[rid];
        reg.span_count++;
        for (int dir = 0; dir < 4; ++dir) {
            const uint32_t n = c.get_neighbor((uint32_t)s, DIR8[dir]);
            if (n == NONE) continue;
            const uint16_t nr = src_reg[n];
            if (nr == rid || nr == 0) continue;
            if (nr & BORDER_REG) {
                reg.connects_to_border = true;
                continue;
            }
            if (std::find(reg.connections.begin(), reg.connections.end(), (uint32_t)nr) == reg.connections.end())
                reg.connections.push_back(nr);
        }
    }
    // remove small regions (:598-622)
    for (uint16_t i = 0; i < max_region_id; ++i) {
        WsRegion& reg = regions[i];
        if (reg.id == 0 || (reg.id & BORDER_REG) || reg.span_count == 0) continue;
        if (reg.span_count < min_region_area && !reg.connects_to_border) {
            const uint16_t old = reg.id;
            reg.id = 0;
            for (auto& v : src_reg)
                if (v == old) v = 0;
        }
    }
    // merge small regions (:636-716): decisions of one round are taken before any of them is applied
    if (merge_region_area > 0) {
        for (;;) {
            struct Merge {
                size_t idx;
                uint16_t old_id, best;
                int32_t count;
            };
            std::vector<Merge> merges;
            for (size_t i = 0; i < regions.size(); ++i) {
                const WsRegion& reg = regions[i];
                if (reg.id == 0 || (reg.id & BORDER_REG) || reg.span_count == 0) continue;
                if (reg.span_count < merge_region_area && !reg.connects_to_border) {
                    uint16_t best = 0;
                    int32_t best_size = 0;
                    for (uint32_t nid : reg.connections) {
                        if (nid >= regions.size()) continue;
                        const WsRegion& nb = regions[nid];
                        if (nb.id == 0 || (nb.id & BORDER_REG)) continue;
                        if (nb.span_count > best_size) {
                            best = nb.id;
                            best_size = nb.span_count;
                        }
                    }
                    if (best > 0) merges.push_back(Merge{i, reg.id, best, reg.span_count});
                }
            }
            for (const Merge& m : merges) {
                for (auto& v : src_reg)
                    if (v == m.old_id) v = m.best;
                regions[m.idx].id = 0;
                for (auto& r : regions)
                    if (r.id == m.best) {
                        r.span_count += m.count;
                        break;
                    }
            }
            if (merges.empty()) break;
        }
    }
    // compact ids (:719-746)
    std::vector<uint16_t> remap(regions.size(), 0);
    uint16_t new_id = 1;
    for (size_t i = 0; i < regions.size(); ++i) {
        if (regions[i].id == 0) continue;
        if (regions[i].id & BORDER_REG)
            remap[i] = regions[i].id;
        else
            remap[i] = new_id++;
    }
    for (auto& v : src_reg)
        if (v < (uint16_t)remap.size()) v = remap[v];  // :739 — `remap.len() as u16`
}

}  // namespace

extern "C" {

// watershed.rs:364-510.  reg: [S] out.  Returns 0 or RCB_ERECAST (-6, "Region ID overflow"); *max_regions = chf.max_regions.
// Runs build_distance_field first, as the reference does (:375).
int orc_chf_build_regions(orc_chf* p, int32_t border_size, int32_t min_region_area, int32_t merge_region_area, uint16_t* reg,
                          uint16_t* max_regions) {
    Compact& c = *(Compact*)p;
    const int32_t w = c.width, h = c.height;
    const size_t S = c.smin.size();
    build_distance_field(c);
    std::vector<uint16_t> src_reg(S, 0), src_dist(S, 0);
    constexpr int LOG_NB_STACKS = 3, NB_STACKS = 1 << LOG_NB_STACKS;
    std::vector<LevelEntry> lvl[NB_STACKS], stack;
    uint16_t region_id = 1;
    uint16_t level = (uint16_t)((c.max_distance + 1) & ~1);
    const int expand_iters = 8;
    if (border_size > 0) {
        const int32_t bw = std::min(w, border_size), bh = std::min(h, border_size);
        ws_paint_rect(0, bw, 0, h, region_id | BORDER_REG, c, src_reg);
        region_id++;
        ws_paint_rect(w - bw, w, 0, h, region_id | BORDER_REG, c, src_reg);
        region_id++;
        ws_paint_rect(0, w, 0, bh, region_id | BORDER_REG, c, src_reg);
        region_id++;
        ws_paint_rect(0, w, h - bh, h, region_id | BORDER_REG, c, src_reg);
        region_id++;
    }
    int s_id = -1;
    while (level > 0) {
        level = level >= 2 ? (uint16_t)(level - 2) : 0;
        s_id = (s_id + 1) & (NB_STACKS - 1);
        if (s_id == 0) {
            ws_sort_cells_by_level(level, c, src_reg, lvl, NB_STACKS, 1);
        } else {
            for (const LevelEntry& e : lvl[s_id - 1])  // append_stacks (:122-133)
                if (e.index != NONE && e.index < S && src_reg[e.index] == 0) lvl[s_id].push_back(e);
        }
        ws_expand_regions(expand_iters, level, c, src_reg, src_dist, lvl[s_id], false);
        for (size_t j = 0; j < lvl[s_id].size(); ++j) {
            const LevelEntry cur = lvl[s_id][j];
            if (cur.index < S && src_reg[cur.index] == 0 &&
                ws_flood_region(cur.x, cur.y, cur.index, level, region_id, c, src_reg, src_dist, stack)) {
                if (region_id == 0xFFFF) return -6;
                region_id++;
            }
        }
    }
    ws_expand_regions(expand_iters * 8, 0, c, src_reg, src_dist, stack, true);
    if (max_regions) *max_regions = region_id;
    ws_merge_and_filter(min_region_area, merge_region_area, region_id, c, src_reg);
    if (reg && S) memcpy(reg, src_reg.data(), 2 * S);
    return 0;
}

}  // extern "C"

// ======================================================================================================
// Convex volumes (second half): ConvexVolume / ConvexVolumeSet of convex_volume.rs, as used by
// RecastBuilder::build_mesh_with_volumes / build_heightfield_with_volumes (lib.rs:293-361): after the filters,
// every span of the SOLID heightfield takes the area of the last volume that contains the point
// (cell centre, span top).  Pinned by the reference's own tests (convex_volume.rs:434-558) in
// tests/test_convex_volumes.py.
// ======================================================================================================
extern "C" {

// convex_volume.rs:166-199 is_convex (xz-plane cross products; zero crosses ignored).  verts: n xyz triples.
int orc_convex_is_convex(const float* v, uint32_t n) {
    if (n < 3) return 0;
    float sign = 0.0f;
    for (uint32_t i = 0; i < n; ++i) {
        const uint32_t j = (i + 1) % n, k = (i + 2) % n;
        const float dx1 = v[j * 3] - v[i * 3], dz1 = v[j * 3 + 2] - v[i * 3 + 2];
        const float dx2 = v[k * 3] - v[j * 3], dz2 = v[k * 3 + 2] - v[j * 3 + 2];
        const float cross = dx1 * dz2 - dz1 * dx2;
        if (cross != 0.0f) {
            if (sign == 0.0f)
                sign = cross;
            else if ((cross > 0.0f) != (sign > 0.0f))
                return 0;
        }
    }
    return 1;
}

// ConvexVolume::new (:28-68): 0 or RCB_EINVALID_MESH
int orc_convex_volume_validate(const float* v, uint32_t n, float min_h, float max_h) {
    if (n < 3 || n > 32 || min_h > max_h || !orc_convex_is_convex(v, n)) return RCB_EINVALID_MESH;
    return 0;
}

// contains_point (:130-138) + point_in_polygon_2d (:140-164)
int orc_convex_contains_point(const float* v, uint32_t n, float min_h, float max_h, const float* p) {
    if (p[1] < min_h || p[1] > max_h) return 0;
    const float x = p[0], z = p[2];
    bool inside = false;
    for (uint32_t i = 0, j = n - 1; i < n; j = i++) {
        const float* vi = v + i * 3;
        const float* vj = v + j * 3;
        if ((vi[2] > z) == (vj[2] > z)) continue;
        if (x >= (vj[0] - vi[0]) * (z - vi[2]) / (vj[2] - vi[2]) + vi[0]) continue;
        inside = !inside;
    }
    return inside ? 1 : 0;
}

// ConvexVolumeSet::apply_to_heightfield (:392-430) with get_area_at_point (:375-389): later volumes win.
void orc_hf_apply_convex_volumes(orc_hf* p, const rcb_convex_volume* vols, int nvols, const float* verts) {
    Heightfield& hf = *(Heightfield*)p;
    for (int32_t z = 0; z < hf.height; ++z)
        for (int32_t x = 0; x < hf.width; ++x) {
            const float wx = hf.bmin[0] + ((float)x + 0.5f) * hf.cs;
            const float wz = hf.bmin[2] + ((float)z + 0.5f) * hf.cs;
            for (uint32_t s = hf.first[(size_t)z * hf.width + x]; s != NONE; s = hf.pool[s].next) {
                const float pt[3] = {wx, hf.bmin[1] + (float)hf.pool[s].smax * hf.ch, wz};
                uint8_t area = hf.pool[s].area;
                for (int k = 0; k < nvols; ++k)
                    if (orc_convex_contains_point(verts + 3 * (size_t)vols[k].first_vert, vols[k].num_verts, vols[k].min_height,
                                                  vols[k].max_height, pt))
                        area = (uint8_t)vols[k].area_type;
                hf.pool[s].area = area;
            }
        }
}

}  // extern "C"
