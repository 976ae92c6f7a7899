// k_raster.cu — triangle -> span fragments -> per-column span lists.
//
// Replaces (paths under crates/recast/src/):
//   lib.rs:217-274              per-triangle loop, index check, slope classification (glam 0.29.3)
//   rasterization.rs:246-366    rasterize_tri  (+ :370-377 overlap_bounds)
//   rasterization.rs:172-242    divide_poly
//   rasterization.rs:51-168     add_span_internal (ordered insert + merge)
//
// Bit-exactness: every float expression keeps the reference's operand order and is written with
// round-to-nearest intrinsics (__fadd_rn/__fsub_rn/__fmul_rn/__fdiv_rn/__fsqrt_rn), which ptxas
// never contracts into FMA; the file is additionally compiled with --fmad=false --prec-div=true
// --prec-sqrt=true --ftz=false.  `as i32` = __float2int_rz (saturating, NaN -> 0), like Rust.
//
// Order dependence: span AREA depends on the order triangles hit a column.  Fragments
// are therefore keyed (cell, triangle index) and each column is replayed in ascending triangle
// index by one thread, running the reference's merge rule verbatim.
#include "rcb_common.cuh"

namespace {

struct Poly {
    int n;
    float v[RCB_MAX_POLY_VERTS * 3];
};

__device__ __forceinline__ void poly_push(Poly& p, float x, float y, float z, bool& ovf) {
    if (p.n >= RCB_MAX_POLY_VERTS) {
        ovf = true;
        return;
    }
    p.v[p.n * 3 + 0] = x;
    p.v[p.n * 3 + 1] = y;
    p.v[p.n * 3 + 2] = z;
    p.n++;
}

// rasterization.rs:172-242.  out1: side < offset, out2: side > offset.
__device__ void divide_poly(const Poly& in, Poly& out1, Poly& out2, float offset, int axis, bool& ovf) {
    out1.n = 0;
    out2.n = 0;
    const int n = in.n;
    if (n == 0) return;
    for (int i = 0; i < n; ++i) {
        const int j = (i + 1 == n) ? 0 : i + 1;
        const float vix = in.v[i * 3], viy = in.v[i * 3 + 1], viz = in.v[i * 3 + 2];
        const float vjx = in.v[j * 3], vjy = in.v[j * 3 + 1], vjz = in.v[j * 3 + 2];
        const float ci = axis == 0 ? vix : viz;
        const float cj = axis == 0 ? vjx : vjz;
        const int si = ci < offset ? -1 : (ci > offset ? 1 : 0);
        const int sj = cj < offset ? -1 : (cj > offset ? 1 : 0);
        if (si == 0) {
            poly_push(out1, vix, viy, viz, ovf);
            poly_push(out2, vix, viy, viz, ovf);
        } else {
            Poly& mine = si < 0 ? out1 : out2;
            poly_push(mine, vix, viy, viz, ovf);
            if ((si < 0 && sj > 0) || (si > 0 && sj < 0)) {
                // t = (offset - vi[a]) / (vj[a] - vi[a]);  p[k] = vi[k] + t * (vj[k] - vi[k])
                const float t = __fdiv_rn(__fsub_rn(offset, ci), __fsub_rn(cj, ci));
                const float px = __fadd_rn(vix, __fmul_rn(t, __fsub_rn(vjx, vix)));
                const float py = __fadd_rn(viy, __fmul_rn(t, __fsub_rn(vjy, viy)));
                const float pz = __fadd_rn(viz, __fmul_rn(t, __fsub_rn(vjz, viz)));
                poly_push(out1, px, py, pz, ovf);
                poly_push(out2, px, py, pz, ovf);
            }
        }
    }
}

struct Tri {
    float v0[3], v1[3], v2[3];
    float tmin[3], tmax[3];
};

__device__ __forceinline__ void tri_bounds(Tri& t) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {  // rasterization.rs:258-264
        t.tmin[i] = fminf(fminf(t.v0[i], t.v1[i]), t.v2[i]);
        t.tmax[i] = fmaxf(fmaxf(t.v0[i], t.v1[i]), t.v2[i]);
    }
}

// rasterization.rs:270 + :275-285.  Returns false when the triangle misses the tile.
__device__ __forceinline__ bool tile_footprint(const Tri& t, const TileDesc& td, int& x0, int& x1, int& z0, int& z1) {
    if (!(t.tmin[0] <= td.bmax[0] && t.tmax[0] >= td.bmin[0] && t.tmin[1] <= td.bmax[1] && t.tmax[1] >= td.bmin[1] &&
          t.tmin[2] <= td.bmax[2] && t.tmax[2] >= td.bmin[2]))
        return false;
    x0 = __float2int_rz(__fmul_rn(__fsub_rn(t.tmin[0], td.bmin[0]), td.ics));
    x1 = __float2int_rz(__fmul_rn(__fsub_rn(t.tmax[0], td.bmin[0]), td.ics));
    z0 = __float2int_rz(__fmul_rn(__fsub_rn(t.tmin[2], td.bmin[2]), td.ics));
    z1 = __float2int_rz(__fmul_rn(__fsub_rn(t.tmax[2], td.bmin[2]), td.ics));
    x0 = max(x0, 0);
    x1 = min(x1, td.width - 1);
    z0 = max(z0, -1);
    z1 = min(z1, td.height - 1);
    return true;
}

// lib.rs:252-270 (glam scalar Vec3).  Returns false for a skipped (degenerate) triangle.
__device__ __forceinline__ bool classify(const Tri& t, float thr, uint8_t& area) {
    const float e1x = __fsub_rn(t.v1[0], t.v0[0]), e1y = __fsub_rn(t.v1[1], t.v0[1]), e1z = __fsub_rn(t.v1[2], t.v0[2]);
    const float e2x = __fsub_rn(t.v2[0], t.v0[0]), e2y = __fsub_rn(t.v2[1], t.v0[1]), e2z = __fsub_rn(t.v2[2], t.v0[2]);
    const float cx = __fsub_rn(__fmul_rn(e1y, e2z), __fmul_rn(e2y, e1z));
    const float cy = __fsub_rn(__fmul_rn(e1z, e2x), __fmul_rn(e2z, e1x));
    const float cz = __fsub_rn(__fmul_rn(e1x, e2y), __fmul_rn(e2x, e1y));
    const float len = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(cx, cx), __fmul_rn(cy, cy)), __fmul_rn(cz, cz)));
    if (len < 1.1920928955078125e-07f) return false;  // f32::EPSILON
    const float ny = __fmul_rn(cy, __fdiv_rn(1.0f, len));
    area = (fabsf(ny) >= thr) ? 1 : 0;
    return true;
}

__device__ __forceinline__ bool load_triangle(const MeshView& m, uint32_t tri, Tri& t, bool& bad_index) {
    const int32_t i0 = m.tris[tri * 3ull], i1 = m.tris[tri * 3ull + 1], i2 = m.tris[tri * 3ull + 2];
    if ((uint32_t)i0 >= m.nverts || (uint32_t)i1 >= m.nverts || (uint32_t)i2 >= m.nverts) {  // lib.rs:224-233
        bad_index = true;
        return false;
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        t.v0[k] = m.verts[i0 * 3ull + k];
        t.v1[k] = m.verts[i1 * 3ull + k];
        t.v2[k] = m.verts[i2 * 3ull + k];
    }
    tri_bounds(t);
    return true;
}

// Enumerate the tiles whose AABB the triangle AABB touches (conservative via the locator grid, then
// the reference's exact inclusive test inside tile_footprint).  A tile registered in several bins is
// visited once: only in the bin (max(tri lower bin, tile lower bin)).
template <class Fn>
__device__ __forceinline__ void for_each_tile(const Tri& t, const Locator& loc, const TileDesc* tiles, Fn&& fn) {
    // NaN coordinates fail every comparison of overlap_bounds => no tile
    if (!(t.tmin[0] <= t.tmax[0]) || !(t.tmin[2] <= t.tmax[2])) return;
    const int bx0 = rcb_bin_coord(t.tmin[0], loc.gx0, loc.inv_bx, loc.nbx);
    const int bx1 = rcb_bin_coord(t.tmax[0], loc.gx0, loc.inv_bx, loc.nbx);
    const int bz0 = rcb_bin_coord(t.tmin[2], loc.gz0, loc.inv_bz, loc.nbz);
    const int bz1 = rcb_bin_coord(t.tmax[2], loc.gz0, loc.inv_bz, loc.nbz);
    for (int bz = bz0; bz <= bz1; ++bz)
        for (int bx = bx0; bx <= bx1; ++bx) {
            const uint32_t b = (uint32_t)bz * loc.nbx + bx;
            const uint32_t e = loc.bin_off[b + 1];
            for (uint32_t k = loc.bin_off[b]; k < e; ++k) {
                const uint32_t tile = loc.bin_tiles[k];
                const TileDesc& td = tiles[tile];
                if (bx != max(bx0, td.bin_x0) || bz != max(bz0, td.bin_z0)) continue;
                int x0, x1, z0, z1;
                if (!tile_footprint(t, td, x0, x1, z0, z1)) continue;
                fn(tile, td, x0, x1, z0, z1);
            }
        }
}

// ---- pass 1: validate indices, upper bound of the fragment count --------------------------------
__global__ void __launch_bounds__(256) k_count_fragments(RasterBuffers rb) {
    const uint32_t tri = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long ub = 0;
    bool bad = false;
    if (tri < rb.mesh.ntris) {
        Tri t;
        if (load_triangle(rb.mesh, tri, t, bad)) {
            for_each_tile(t, rb.loc, rb.tiles, [&](uint32_t, const TileDesc&, int x0, int x1, int z0, int z1) {
                const int zz0 = max(z0, 0);
                if (x1 >= x0 && z1 >= zz0) ub += (unsigned long long)(x1 - x0 + 1) * (unsigned long long)(z1 - zz0 + 1);
            });
        }
    }
    // block reduction -> one atomic per block
    __shared__ unsigned long long sh[8];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) ub += __shfl_down_sync(0xffffffffu, ub, d);
    if (lane_id() == 0) sh[threadIdx.x >> 5] = ub;
    const bool any_bad = __syncthreads_or(bad);
    if (threadIdx.x == 0) {
        unsigned long long s = 0;
        for (int i = 0; i < 8; ++i) s += sh[i];
        if (s) atomicAdd(&rb.counters[0], s);
        if (any_bad) atomicOr(&rb.counters[3], 1ull);
    }
}

// ---- pass 2: clip and emit fragments --------------------------------------------------------------
__device__ __forceinline__ void emit_fragment(const RasterBuffers& rb, uint32_t gcell, uint32_t tri, int smin, int smax,
                                              uint32_t area) {
    const uint32_t rank = atomicAdd(&rb.cell_cnt[gcell], 1u);
    // warp-aggregated append to the unsorted fragment list
    const unsigned m = __activemask();
    const int leader = __ffs(m) - 1;
    unsigned long long base = 0;
    if ((int)lane_id() == leader) base = atomicAdd(&rb.counters[1], (unsigned long long)__popc(m));
    base = __shfl_sync(m, base, leader);
    const unsigned long long slot = base + __popc(m & ((1u << lane_id()) - 1u));
    if (slot < rb.frag_cap)
        rb.frag_tmp[slot] = make_uint4(gcell, tri, ((uint32_t)smin & 0xffffu) | ((uint32_t)smax << 16),
                                       (rank & 0xffffffu) | (area << 24));
    else
        atomicOr(&rb.counters[2], 2ull);  // cannot happen: frag_cap >= upper bound
}

__global__ void __launch_bounds__(128) k_clip_emit(RasterBuffers rb, TileInfo* __restrict__ tinfo) {
    const uint32_t tri = blockIdx.x * blockDim.x + threadIdx.x;
    if (tri >= rb.mesh.ntris) return;
    Tri t;
    bool bad = false;
    if (!load_triangle(rb.mesh, tri, t, bad)) return;
    const bool explicit_area = rb.mesh.tri_areas != nullptr;
    const uint32_t given_area = explicit_area ? rb.mesh.tri_areas[tri] : 0;

    for_each_tile(t, rb.loc, rb.tiles, [&](uint32_t tile, const TileDesc& td, int x0, int x1, int z0, int z1) {
        uint32_t area = given_area;
        if (!explicit_area) {
            uint8_t a;
            if (!classify(t, td.slope_thr, a)) return;  // lib.rs:258-260
            area = a;
        }
        bool ovf = false;
        Poly bufA, bufB, in_row, cell, rest2;
        Poly* p1 = &bufA;
        Poly* rest = &bufB;
        p1->n = 3;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            p1->v[k] = t.v0[k];
            p1->v[3 + k] = t.v1[k];
            p1->v[6 + k] = t.v2[k];
        }
        for (int z = z0; z <= z1; ++z) {  // rasterization.rs:300-363
            const float cz_max = __fadd_rn(td.bmin[2], __fmul_rn((float)(z + 1), td.cs));
            divide_poly(*p1, in_row, *rest, cz_max, 2, ovf);
            Poly* sw = p1;
            p1 = rest;
            rest = sw;
            if (in_row.n < 3 || z < 0) continue;
            Poly* p2 = &in_row;
            Poly* r2 = &rest2;
            // in_row is consumed by the column loop; p2/r2 ping-pong between in_row and rest2
            for (int x = x0; x <= x1; ++x) {
                const float cx_max = __fadd_rn(td.bmin[0], __fmul_rn((float)(x + 1), td.cs));
                divide_poly(*p2, cell, *r2, cx_max, 0, ovf);
                Poly* s2 = p2;
                p2 = r2;
                r2 = s2;
                if (cell.n < 3) continue;
                float min_y = cell.v[1], max_y = cell.v[1];
                for (int i = 1; i < cell.n; ++i) {
                    min_y = fminf(min_y, cell.v[i * 3 + 1]);
                    max_y = fmaxf(max_y, cell.v[i * 3 + 1]);
                }
                // rasterization.rs:346-350: floor/ceil, saturating `as i16`, clamp smin to 0
                int smin = __float2int_rz(floorf(__fmul_rn(__fsub_rn(min_y, td.bmin[1]), td.ich)));
                int smax = __float2int_rz(ceilf(__fmul_rn(__fsub_rn(max_y, td.bmin[1]), td.ich)));
                smin = min(max(smin, -32768), 32767);
                smax = min(max(smax, -32768), 32767);
                smin = max(smin, 0);
                emit_fragment(rb, td.cell_base + (uint32_t)z * td.width + x, tri, smin, smax, area);
            }
        }
        if (ovf) {
            atomicOr(&tinfo[tile].status, 1u);
            atomicOr(&rb.counters[2], 1ull);
        }
    });
}

// Column segments hold one record per fragment {triangle, min | max << 16, area}: one scattered store per fragment, and
// the replay's sort moves a record at a time.  8 bytes while triangle indices fit 24 bits, 16 bytes beyond.
struct Rec8 {
    uint2 v;
    __device__ __forceinline__ static Rec8 make(uint32_t tri, uint32_t mm, uint32_t area) { return Rec8{make_uint2((tri << 8) | area, mm)}; }
    __device__ __forceinline__ uint32_t key() const { return v.x; }  // triangle index in the high bits
    __device__ __forceinline__ uint32_t mm() const { return v.y; }
    __device__ __forceinline__ uint32_t area() const { return v.x & 0xffu; }
};
struct Rec16 {
    uint4 v;
    __device__ __forceinline__ static Rec16 make(uint32_t tri, uint32_t mm, uint32_t area) { return Rec16{make_uint4(tri, mm, area, 0u)}; }
    __device__ __forceinline__ uint32_t key() const { return v.x; }
    __device__ __forceinline__ uint32_t mm() const { return v.y; }
    __device__ __forceinline__ uint32_t area() const { return v.z; }
};

// ---- pass 3: scatter fragments to their column segment --------------------------------------------
template <typename R>
__global__ void __launch_bounds__(256) k_scatter_fragments(const uint4* __restrict__ frag_tmp,
                                                           const unsigned long long* __restrict__ nfrag_dev,
                                                           const uint32_t* __restrict__ frag_off, R* __restrict__ rec) {
    const unsigned long long n = *nfrag_dev;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const uint4 r = frag_tmp[i];
        rec[frag_off[r.x] + (r.w & 0xffffffu)] = R::make(r.y, r.z, r.w >> 24);
    }
}

// ---- pass 4: per-column ordered replay of add_span_internal ----------------------------------------
// One thread per column.  The column's fragment records occupy [frag_off[c], frag_off[c+1]) in arbitrary order.  They
// are sorted by triangle index in place, then replayed; the span list is built in place over the consumed prefix of
// the same segment (#spans <= #fragments consumed).
template <typename R>
__global__ void __launch_bounds__(256) k_replay_columns(uint32_t ncells, const uint32_t* __restrict__ frag_off, R* __restrict__ rec,
                                                        uint32_t* __restrict__ span_cnt, int32_t thr, const uint32_t* __restrict__ ex_off) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncells) return;
    const uint32_t b = frag_off[c], e = frag_off[c + 1];
    const uint32_t n = e - b;
    // rasterizing onto a non-empty heightfield: the column's first nex entries are its existing spans, already a list
    const uint32_t nex = ex_off ? ex_off[c + 1] - ex_off[c] : 0u;
    if (n == 0) {
        span_cnt[c] = 0;
        return;
    }
    R* r = rec + b;
    // insertion sort by triangle index (typical n: 2-4 on terrain, ~40 in the 16-floor interior)
    for (uint32_t i = nex + 1; i < n; ++i) {
        const R key = r[i];
        uint32_t j = i;
        while (j > nex && r[j - 1].key() > key.key()) {
            r[j] = r[j - 1];
            --j;
        }
        if (j != i) r[j] = key;
    }
    // replay (rasterization.rs:51-168) on the array form of the list
    uint32_t ns = nex;
    for (uint32_t i = nex; i < n; ++i) {
        const R f = r[i];
        const uint32_t w = f.mm();
        const int smin = (int16_t)(w & 0xffffu), smax = (int16_t)(w >> 16);
        const uint32_t area = f.area();
        uint32_t p = 0;
        if (nex == 0 && ns > 4) {
            // a list this rule built from nothing is ordered and disjoint: every span that ends below smin is skipped
            // by the walk below, so start at the first that does not
            uint32_t hi = ns;
            while (p < hi) {
                const uint32_t mid = (p + hi) >> 1;
                if ((int)(int16_t)(r[mid].mm() >> 16) < smin)
                    p = mid + 1;
                else
                    hi = mid;
            }
        }
        bool done = false;
        while (p < ns) {
            const R cur = r[p];
            const uint32_t cw = cur.mm();
            const int cmin = (int16_t)(cw & 0xffffu), cmax = (int16_t)(cw >> 16);
            if (smax < cmin) break;  // insert before p
            if (smin > cmax) {
                ++p;
                continue;
            }
            // merge into p
            const int mmin = min(smin, cmin);
            int mmax = max(smax, cmax);
            uint32_t marea = area;
            if (abs(mmax - cmax) <= thr) marea = max(marea, cur.area());
            uint32_t q = p + 1;
            while (q < ns) {
                const uint32_t nw = r[q].mm();
                if ((int)(int16_t)(nw & 0xffffu) > mmax) break;
                mmax = max(mmax, (int)(int16_t)(nw >> 16));
                ++q;
            }
            r[p] = R::make(0u, ((uint32_t)mmin & 0xffffu) | ((uint32_t)mmax << 16), marea);
            if (q > p + 1) {  // drop absorbed spans (p, q)
                const uint32_t drop = q - (p + 1);
                for (uint32_t k = q; k < ns; ++k) r[k - drop] = r[k];
                ns -= drop;
            }
            done = true;
            break;
        }
        if (!done) {  // insert at p (p == ns: append)
            for (uint32_t k = ns; k > p; --k) r[k] = r[k - 1];
            if (p != i) r[p] = f;
            ++ns;
        }
    }
    span_cnt[c] = ns;
}

// ---- pass 5: compact the per-column span prefixes into the final heightfield CSR -------------------
__global__ void __launch_bounds__(256) k_gather_spans(uint32_t ncells, const uint32_t* __restrict__ frag_off,
                                                      const uint32_t* __restrict__ fs_mm,
                                                      const uint8_t* __restrict__ fs_area,
                                                      const uint32_t* __restrict__ hf_off, int16_t* __restrict__ smin,
                                                      int16_t* __restrict__ smax, uint8_t* __restrict__ sarea) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncells) return;
    const uint32_t src = frag_off[c];
    const uint32_t dst = hf_off[c], n = hf_off[c + 1] - dst;
    for (uint32_t k = 0; k < n; ++k) {
        const uint32_t w = fs_mm[src + k];
        smin[dst + k] = (int16_t)(w & 0xffffu);
        smax[dst + k] = (int16_t)(w >> 16);
        sarea[dst + k] = fs_area[src + k];
    }
}

template <typename R>
__global__ void __launch_bounds__(256) k_gather_records(uint32_t ncells, const uint32_t* __restrict__ frag_off, const R* __restrict__ rec,
                                                        const uint32_t* __restrict__ hf_off, int16_t* __restrict__ smin,
                                                        int16_t* __restrict__ smax, uint8_t* __restrict__ sarea) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncells) return;
    const uint32_t src = frag_off[c];
    const uint32_t dst = hf_off[c], n = hf_off[c + 1] - dst;
    for (uint32_t k = 0; k < n; ++k) {
        const R f = rec[src + k];
        smin[dst + k] = (int16_t)(f.mm() & 0xffffu);
        smax[dst + k] = (int16_t)(f.mm() >> 16);
        sarea[dst + k] = (uint8_t)f.area();
    }
}

// ---- per-tile fragment segments -> column order (global back end behind the v2 clipper) ------------------
// One CTA per tile.  Scattering tile by tile keeps the random writes inside the tile's own region
// (a few hundred KB, L2-resident); scattering the whole batch at once cost 240 B of DRAM traffic per
// fragment (13 % L2 hit rate, ncu on config 3).  The per-column counts come from the clipper itself (flag 0x200).
template <typename R>
__global__ void __launch_bounds__(512) k_seg_scatter(const TileDesc* __restrict__ tiles, int ntiles, const uint32_t* __restrict__ tile_frags,
                                                     const uint32_t* __restrict__ tile_base, const uint32_t* __restrict__ tile_cap,
                                                     const uint32_t* __restrict__ tile_cursor, const uint32_t* __restrict__ frag_off,
                                                     uint32_t* __restrict__ cell_cnt, R* __restrict__ rec, unsigned long long* __restrict__ counters) {
    const uint32_t tile = blockIdx.z * gridDim.y + blockIdx.y;
    if (tile >= (uint32_t)ntiles) return;
    const uint32_t F = min(tile_cursor[tile], tile_cap[tile]);
    const uint32_t* fr = tile_frags + 3ull * tile_base[tile];  // three arrays of `cap` words
    const uint32_t cap = tile_cap[tile];
    const uint32_t cb = tiles[tile].cell_base;
    if (blockIdx.x == 0 && threadIdx.x == 0 && F) atomicAdd(&counters[1], (unsigned long long)F);  // fragments of the batch
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < F; i += gridDim.x * blockDim.x) {
        const uint32_t r0 = fr[i], r1 = fr[cap + i], r2 = fr[2ull * cap + i];
        const uint32_t g = cb + (r0 & 0xffffffu);
        // the column's counter runs back to zero; slots fill in arrival order, which is close to triangle
        // order, so the replay's insertion sort stays near its best case
        const uint32_t pos = frag_off[g + 1] - atomicSub(&cell_cnt[g], 1u);
        rec[pos] = R::make(r2, r1, r0 >> 24);
    }
}

}  // namespace

// grid (x = slices of a tile's segment, y = tile): few big tiles still fill the machine
static dim3 seg_grid(int ntiles, uint32_t max_tile_frags) {
    uint32_t slices = (max_tile_frags + 8191u) / 8192u;
    const uint32_t want = (uint32_t)((148 * 4 + ntiles - 1) / ntiles);  // enough CTAs for every SM
    slices = slices < 1 ? 1 : slices;
    if (slices > want) slices = want > 1 ? want : 1;
    const unsigned gy = ntiles < 65535 ? (unsigned)ntiles : 65535u;  // tile = blockIdx.z * gridDim.y + blockIdx.y
    return dim3(slices, gy, ((unsigned)ntiles + gy - 1) / gy);
}

void launch_seg_scatter(const TileDesc* tiles, int ntiles, const uint32_t* tile_frags, const uint32_t* tile_base, const uint32_t* tile_cap,
                        const uint32_t* tile_cursor, const uint32_t* frag_off, uint32_t* cell_cnt, void* rec, bool wide,
                        unsigned long long* counters, uint32_t max_tile_frags, cudaStream_t s) {
    if (ntiles == 0) return;
    KScope _k("k_seg_scatter", s);
    if (wide)
        k_seg_scatter<Rec16><<<seg_grid(ntiles, max_tile_frags), 512, 0, s>>>(tiles, ntiles, tile_frags, tile_base, tile_cap, tile_cursor, frag_off, cell_cnt, (Rec16*)rec, counters);
    else
        k_seg_scatter<Rec8><<<seg_grid(ntiles, max_tile_frags), 512, 0, s>>>(tiles, ntiles, tile_frags, tile_base, tile_cap, tile_cursor, frag_off, cell_cnt, (Rec8*)rec, counters);
}

void launch_count_fragments(const RasterBuffers& rb, cudaStream_t s) {
    if (rb.mesh.ntris == 0) return;
    {
        KScope _k("k_count_fragments", s);
        k_count_fragments<<<(rb.mesh.ntris + 255) / 256, 256, 0, s>>>(rb);
    }
}

void launch_clip_emit(const RasterBuffers& rb, TileInfo* tinfo, cudaStream_t s) {
    if (rb.mesh.ntris == 0) return;
    {
        KScope _k("k_clip_emit", s);
        k_clip_emit<<<(rb.mesh.ntris + 127) / 128, 128, 0, s>>>(rb, tinfo);
    }
}

void launch_scatter_fragments(const uint4* frag_tmp, const unsigned long long* nfrag_dev, uint32_t nfrag_hint,
                              const uint32_t* frag_off, void* rec, bool wide, cudaStream_t s) {
    if (nfrag_hint == 0) return;
    unsigned blocks = (unsigned)((nfrag_hint + 255ull) / 256ull);
    if (blocks > 148u * 64u) blocks = 148u * 64u;
    {
        KScope _k("k_scatter_fragments", s);
        if (wide)
            k_scatter_fragments<Rec16><<<blocks, 256, 0, s>>>(frag_tmp, nfrag_dev, frag_off, (Rec16*)rec);
        else
            k_scatter_fragments<Rec8><<<blocks, 256, 0, s>>>(frag_tmp, nfrag_dev, frag_off, (Rec8*)rec);
    }
}

void launch_replay_columns(uint32_t ncells, const uint32_t* frag_off, void* rec, bool wide, uint32_t* span_cnt, int32_t merge_thr,
                           cudaStream_t s, const uint32_t* ex_off) {
    if (ncells == 0) return;
    {
        KScope _k("k_replay_columns", s);
        if (wide)
            k_replay_columns<Rec16><<<(ncells + 255) / 256, 256, 0, s>>>(ncells, frag_off, (Rec16*)rec, span_cnt, merge_thr, ex_off);
        else
            k_replay_columns<Rec8><<<(ncells + 255) / 256, 256, 0, s>>>(ncells, frag_off, (Rec8*)rec, span_cnt, merge_thr, ex_off);
    }
}

// Rasterizing onto a non-empty heightfield (rasterize_triangles on a Heightfield that already holds spans,
// rasterization.rs:405-428 / lib.rs:186-290): sign = +1 adds the existing span counts to the per-column fragment
// counts before the offsets are scanned; the second call (sign = -1) takes them out again, so that the scatter fills
// the END of each column's range, and copies the existing spans to its beginning.
template <typename R>
__global__ void __launch_bounds__(256) k_onto(uint32_t ncells, const uint32_t* __restrict__ ex_off, uint32_t* __restrict__ cell_cnt,
                                              int place, const uint32_t* __restrict__ frag_off, const int16_t* __restrict__ smin,
                                              const int16_t* __restrict__ smax, const uint8_t* __restrict__ sarea, R* __restrict__ rec) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncells) return;
    const uint32_t b = ex_off[c], n = ex_off[c + 1] - b;
    if (!place) {
        cell_cnt[c] += n;
        return;
    }
    cell_cnt[c] -= n;
    const uint32_t o = frag_off[c];
    for (uint32_t k = 0; k < n; ++k)
        rec[o + k] = R::make(0u, ((uint32_t)(uint16_t)smin[b + k]) | ((uint32_t)(uint16_t)smax[b + k] << 16), sarea[b + k]);
}

void launch_onto(uint32_t ncells, const uint32_t* ex_off, uint32_t* cell_cnt, int place, const uint32_t* frag_off, const int16_t* smin,
                 const int16_t* smax, const uint8_t* sarea, void* rec, bool wide, cudaStream_t s) {
    if (ncells == 0) return;
    KScope _k("k_onto", s);
    if (wide)
        k_onto<Rec16><<<(ncells + 255) / 256, 256, 0, s>>>(ncells, ex_off, cell_cnt, place, frag_off, smin, smax, sarea, (Rec16*)rec);
    else
        k_onto<Rec8><<<(ncells + 255) / 256, 256, 0, s>>>(ncells, ex_off, cell_cnt, place, frag_off, smin, smax, sarea, (Rec8*)rec);
}

void launch_gather_records(uint32_t ncells, const uint32_t* frag_off, const void* rec, bool wide, const uint32_t* hf_off, int16_t* smin,
                           int16_t* smax, uint8_t* sarea, cudaStream_t s) {
    if (ncells == 0) return;
    KScope _k("k_gather_spans", s);
    if (wide)
        k_gather_records<Rec16><<<(ncells + 255) / 256, 256, 0, s>>>(ncells, frag_off, (const Rec16*)rec, hf_off, smin, smax, sarea);
    else
        k_gather_records<Rec8><<<(ncells + 255) / 256, 256, 0, s>>>(ncells, frag_off, (const Rec8*)rec, hf_off, smin, smax, sarea);
}

void launch_gather_spans(uint32_t ncells, const uint32_t* frag_off, const uint32_t* fs_mm, const uint8_t* fs_area,
                         const uint32_t* hf_off, int16_t* smin, int16_t* smax, uint8_t* sarea, cudaStream_t s) {
    if (ncells == 0) return;
    {
        KScope _k("k_gather_spans", s);
        k_gather_spans<<<(ncells + 255) / 256, 256, 0, s>>>(ncells, frag_off, fs_mm, fs_area, hf_off, smin, smax, sarea);
    }
}
