// k_raster2.cu — second-generation clipper: (triangle, tile, row-chunk) work items, polygons in
// shared memory, row-granular load balance.
//
// Replaces the same reference code as k_raster.cu (lib.rs:217-274, rasterization.rs:172-366) and is
// bit-identical to it; what changes is the mapping to the machine.  ncu on the v1 clipper
// (profiles/r01_v1_*) showed 10.8 of 32 lanes active (threads walked whole triangles of very different
// row/column counts), 610 M local-memory sectors and 716 MB of DRAM writes from polygon spills.
//
// Structure (one CTA = CLIP_ITEMS work items, CLIP_THREADS threads):
//   phase 1  one of the first CLIP_ITEMS threads = item (tri, tile, chunk of up to CLIP_ROWS rows): z-chain.  The chain
//            of a triangle is sequential (each cut interpolates from the previous cut's rounded points), so a chunk that
//            starts at row r first replays the r cuts before it (one divide per row; triangles of up to CLIP_ROWS rows
//            are one chunk and replay nothing), then stores the row polygons of its rows — x,y only, z is never read
//            again — into shared-memory slots.
//   phase 2  every thread = row slot: x-chain over the columns of that row.  Only the count and the y-range of
//            the cell piece are needed, so the cell polygon is never stored.  Columns whose cut lies
//            strictly left of every vertex are skipped without a divide, ahead of the column loop (the divide would
//            return the polygon unchanged and an empty cell piece, so skipping is exact).
// Fragments go either to one global list with a per-column rank (v1 back end) or to per-tile segments
// (tile-resident back end, k_tile.cu).
#include <algorithm>

#include "rcb_common.cuh"
#define RCB_FILE_ID 2
#include "rcb_bounds.cuh"

namespace {

constexpr int CLIP_THREADS = 128;
constexpr int CLIP_ROWS = RCB_CLIP_ROWS;                  // rows per work item
// Capacities = the combinatorial bounds (a cut crosses a convex polygon's boundary twice): what is left of a triangle
// behind a row line has <= 4 vertices, a row strip of it <= 5, what is left of that behind a column line <= 6.  A
// polygon that outgrows its slot raises the overflow flag and the batch is redone by the 12-vertex clipper.  Shared
// memory per thread decides how many CTAs an SM holds, and this kernel's time follows its occupancy (5 -> 3 CTAs: +49 %).
constexpr int ZCAP = 4;                                   // z-chain polygon capacity
constexpr int RCAP = 5;                                   // row polygon capacity (slots)
constexpr int XCAP = 6;                                   // x-chain rest polygon capacity
constexpr int SLOT_WORDS = 1 + 2 * RCAP;                  // row word + verts (odd stride: no bank conflicts)
constexpr int HDR_WORDS = 3;                              // per work item: triangle, tile, x0 | x1 << 16 (shared by its rows)
static_assert(2 * XCAP * 2 <= 2 * ZCAP * 3, "phase-2 buffers alias the z-chain scratch");
// A CTA holds NSLOTS row polygons whatever the rows per work item: CLIP_ITEMS = NSLOTS / CLIP_ROWS items per CTA, whose
// z-chains (phase 1) run on the first CLIP_ITEMS threads while phase 2 spreads the rows over all CLIP_THREADS.  Rows per item
// trade the replay of the rows before a chunk (short chunks: a 5-row triangle costs 3 + (3 + 2) chain steps in chunks of 3)
// against idle threads in phase 1; shared memory, and with it the CTAs per SM this kernel's time follows, stays the same.
constexpr int SLOTS_PER_THREAD = 3;
constexpr int NSLOTS = CLIP_THREADS * SLOTS_PER_THREAD;
constexpr int CLIP_ITEMS = NSLOTS / CLIP_ROWS;            // work items per CTA
static_assert(CLIP_ITEMS * CLIP_ROWS == NSLOTS && CLIP_ITEMS <= CLIP_THREADS && CLIP_ITEMS >= 1, "rows per item must divide the CTA's slots");
constexpr int CLIP_CTAS = 7;                              // resident CTAs per SM the shared memory allows
constexpr int NBUCK = 16;  // phase-2 rows are grouped by (vertex count 3..6+, columns spanned 1..4+): a scheduling hint only
                           // (six column classes - 1, 2, 3, 4, 5-8, 9+ - were tried for config 3's roofs and walls: no change there,
                           // +1 % on config 2 for the extra barrier)

struct Tri {
    float v0[3], v1[3], v2[3];
    float tmin[3], tmax[3];
};

__device__ __forceinline__ bool load_triangle(const MeshView& m, uint32_t tri, Tri& t) {
    const int32_t i0 = m.tris[tri * 3ull], i1 = m.tris[tri * 3ull + 1], i2 = m.tris[tri * 3ull + 2];
    if ((uint32_t)i0 >= m.nverts || (uint32_t)i1 >= m.nverts || (uint32_t)i2 >= m.nverts) return false;  // lib.rs:224-233
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        t.v0[k] = m.verts[i0 * 3ull + k];
        t.v1[k] = m.verts[i1 * 3ull + k];
        t.v2[k] = m.verts[i2 * 3ull + k];
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) {  // rasterization.rs:258-264
        t.tmin[i] = fminf(fminf(t.v0[i], t.v1[i]), t.v2[i]);
        t.tmax[i] = fmaxf(fmaxf(t.v0[i], t.v1[i]), t.v2[i]);
    }
    return true;
}

// rasterization.rs:270 + :275-285
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

// lib.rs:252-270 (glam scalar Vec3)
__device__ __forceinline__ bool classify(const Tri& t, float thr, uint32_t& area) {
    const float e1x = __fsub_rn(t.v1[0], t.v0[0]), e1y = __fsub_rn(t.v1[1], t.v0[1]), e1z = __fsub_rn(t.v1[2], t.v0[2]);
    const float e2x = __fsub_rn(t.v2[0], t.v0[0]), e2y = __fsub_rn(t.v2[1], t.v0[1]), e2z = __fsub_rn(t.v2[2], t.v0[2]);
    const float cx = __fsub_rn(__fmul_rn(e1y, e2z), __fmul_rn(e2y, e1z));
    const float cy = __fsub_rn(__fmul_rn(e1z, e2x), __fmul_rn(e2z, e1x));
    const float cz = __fsub_rn(__fmul_rn(e1x, e2y), __fmul_rn(e2x, e1y));
    const float len = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(cx, cx), __fmul_rn(cy, cy)), __fmul_rn(cz, cz)));
    if (len < 1.1920928955078125e-07f) return false;
    const float ny = __fmul_rn(cy, __fdiv_rn(1.0f, len));
    area = (fabsf(ny) >= thr) ? 1u : 0u;
    return true;
}

template <class Fn>
__device__ __forceinline__ void for_each_tile(const Tri& t, const Locator& loc, const TileDesc* __restrict__ tiles, Fn&& fn) {
    if (!(t.tmin[0] <= t.tmax[0]) || !(t.tmin[2] <= t.tmax[2])) return;  // NaN: overlap_bounds is false everywhere
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

// ---- pass 1: validate indices; per triangle: number of work items; per tile: fragment upper bound ----
__global__ void __launch_bounds__(256) k_bin_count(Raster2 rb) {
    const uint32_t tri = blockIdx.x * blockDim.x + threadIdx.x;
    if (rb.mesh.blocks) {
        // A batch that covers part of the world (a tile group of a streamed or multi-GPU build) meets mostly triangles
        // that lie elsewhere: the 256 consecutive triangles of this CTA are skipped together when their common AABB misses
        // the batch's extent (inclusive comparison, as overlap_bounds, rasterization.rs:370-377; fminf / fmaxf ignore NaN
        // as f32::min / max do, so the block box holds every triangle box).
        const float4 b = rb.mesh.blocks[blockIdx.x];
        if (!(b.x <= rb.loc.ex1 && b.z >= rb.loc.ex0 && b.y <= rb.loc.ez1 && b.w >= rb.loc.ez0)) {
            if (tri < rb.mesh.ntris) rb.item_cnt[tri] = 0;
            return;
        }
    }
    unsigned long long ub = 0;
    uint32_t items = 0;
    bool bad = false;
    if (tri < rb.mesh.ntris) {
        Tri t;
        if (load_triangle(rb.mesh, tri, t)) {
            for_each_tile(t, rb.loc, rb.tiles, [&](uint32_t tile, const TileDesc& td, int x0, int x1, int z0, int z1) {
                const int zz0 = max(z0, 0);
                if (x1 < x0 || z1 < zz0) return;
                const uint32_t nrows = (uint32_t)(z1 - zz0 + 1), ncols = (uint32_t)(x1 - x0 + 1);
                items += (nrows + CLIP_ROWS - 1) / CLIP_ROWS;
                if (!rb.tile_ub) ub += ncols * nrows;
                if (rb.tile_ub) {
                    // per band: the rows of the footprint the band holds, its halo rows included (a border row counts
                    // for both bands, as emit_fragment writes it to both)
                    const int R = td.band_rows;
                    const uint32_t b0 = td.band_first;
                    const int bl = max(zz0 - 1, 0) / R, bh = min(z1 + 1, td.height - 1) / R;
                    for (int b = bl; b <= bh; ++b) {
                        const int hz0 = max(b * R - 1, 0), hz1 = min((b + 1) * R + 1, td.height);  // rows held: [hz0, hz1)
                        const int r0 = max(zz0, hz0), r1 = min(z1, hz1 - 1);
                        if (r1 < r0) continue;
                        const uint32_t f = ncols * (uint32_t)(r1 - r0 + 1);
                        ub += f;  // the segments' total: a border row is held by two bands
                        // neighbouring triangles mostly share the band: one atomic per (warp, band)
                        const unsigned peers = __match_any_sync(__activemask(), b0 + (uint32_t)b);
                        const uint32_t sum = __reduce_add_sync(peers, f);
                        if ((int)lane_id() == __ffs(peers) - 1) atomicAdd(&rb.tile_ub[b0 + (uint32_t)b], sum);
                    }
                }
            });
        } else {
            bad = true;
        }
        rb.item_cnt[tri] = items;
    }
    __shared__ unsigned long long sh_ub[8];
    __shared__ uint32_t sh_it[8];
    unsigned long long it64 = items;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        ub += __shfl_down_sync(0xffffffffu, ub, d);
        it64 += __shfl_down_sync(0xffffffffu, it64, d);
    }
    if (lane_id() == 0) {
        sh_ub[threadIdx.x >> 5] = ub;
        sh_it[threadIdx.x >> 5] = (uint32_t)it64;
    }
    const bool any_bad = __syncthreads_or(bad);
    if (threadIdx.x == 0) {
        unsigned long long s = 0, si = 0;
        for (int i = 0; i < 8; ++i) {
            s += sh_ub[i];
            si += sh_it[i];
        }
        if (s) atomicAdd(&rb.counters[0], s);
        if (si) atomicAdd(&rb.counters[4], si);
        if (any_bad) atomicOr(&rb.counters[3], 1ull);
    }
}

// lib.rs:224-233: any index outside [0, nverts) fails the build (a negative i32 becomes a huge usize)
__global__ void __launch_bounds__(256) k_check_indices(const int32_t* __restrict__ tris, size_t nidx, uint32_t nverts,
                                                       unsigned long long* __restrict__ flag) {
    bool bad = false;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nidx; i += (size_t)gridDim.x * blockDim.x)
        bad = bad || (uint32_t)tris[i] >= nverts;
    if (__syncthreads_or(bad) && threadIdx.x == 0) atomicOr(flag, 1ull);
}

// rasterize_triangles_u16 (rasterization.rs:432-455) indexes with `tris[i] as usize`: the indices widened, nothing else;
// rasterize_triangle_soup (:459-476) reads vertex 3 i + k of triangle i: the identity index list.  src == nullptr: identity.
__global__ void __launch_bounds__(256) k_widen_indices(const uint16_t* __restrict__ src, int32_t* __restrict__ dst, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        dst[i] = src ? (int32_t)src[i] : (int32_t)i;
}

// xz AABB per run of RCB_MESH_BLOCK consecutive triangles (k_bin_count's early reject); a run with an index out of range
// gets an infinite box (the build fails on it anyway, lib.rs:224-233)
__global__ void __launch_bounds__(RCB_MESH_BLOCK) k_mesh_blocks(const float* __restrict__ verts, const int32_t* __restrict__ tris, uint32_t ntris,
                                                                uint32_t nverts, float4* __restrict__ blocks) {
    const uint32_t tri = blockIdx.x * RCB_MESH_BLOCK + threadIdx.x;
    float x0 = INFINITY, z0 = INFINITY, x1 = -INFINITY, z1 = -INFINITY;
    if (tri < ntris) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const uint32_t i = (uint32_t)tris[tri * 3ull + k];
            if (i >= nverts) {
                x0 = z0 = -INFINITY, x1 = z1 = INFINITY;
                break;
            }
            const float x = verts[i * 3ull], z = verts[i * 3ull + 2];
            x0 = fminf(x0, x), x1 = fmaxf(x1, x), z0 = fminf(z0, z), z1 = fmaxf(z1, z);
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        x0 = fminf(x0, __shfl_down_sync(0xffffffffu, x0, d));
        z0 = fminf(z0, __shfl_down_sync(0xffffffffu, z0, d));
        x1 = fmaxf(x1, __shfl_down_sync(0xffffffffu, x1, d));
        z1 = fmaxf(z1, __shfl_down_sync(0xffffffffu, z1, d));
    }
    __shared__ float sh[4][RCB_MESH_BLOCK / 32];
    if ((threadIdx.x & 31) == 0) sh[0][threadIdx.x >> 5] = x0, sh[1][threadIdx.x >> 5] = z0, sh[2][threadIdx.x >> 5] = x1, sh[3][threadIdx.x >> 5] = z1;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < RCB_MESH_BLOCK / 32; ++w)
            x0 = fminf(x0, sh[0][w]), z0 = fminf(z0, sh[1][w]), x1 = fmaxf(x1, sh[2][w]), z1 = fmaxf(z1, sh[3][w]);
        blocks[blockIdx.x] = make_float4(x0, z0, x1, z1);
    }
}

// ---- pass 2: write the work items (tri, tile, chunk) in triangle order -------------------------------
__global__ void __launch_bounds__(256) k_fill_items(Raster2 rb, const uint32_t* __restrict__ item_off, uint2* __restrict__ items) {
    const uint32_t tri = blockIdx.x * blockDim.x + threadIdx.x;
    if (tri >= rb.mesh.ntris) return;
    if (rb.counters[2] & RCB_FLAG_ABORT) return;  // speculative issue: the item buffer is too small
    uint32_t o = item_off[tri];
    if (item_off[tri + 1] == o) return;
    Tri t;
    if (!load_triangle(rb.mesh, tri, t)) return;
    for_each_tile(t, rb.loc, rb.tiles, [&](uint32_t tile, const TileDesc&, int x0, int x1, int z0, int z1) {
        const int zz0 = max(z0, 0);
        if (x1 < x0 || z1 < zz0) return;
        const uint32_t nchunks = ((uint32_t)(z1 - zz0 + 1) + CLIP_ROWS - 1) / CLIP_ROWS;
        for (uint32_t c = 0; c < nchunks; ++c, ++o) items[RCB_IX(o, item_off[tri + 1])] = make_uint2(tri, (tile << 12) | c);
    });
}

// ---- clip --------------------------------------------------------------------------------------------
struct ClipSmem {
    float zbuf[2 * ZCAP * 3 * CLIP_THREADS];  // phase 1: [(buf*ZCAP + v)*3 + k][tid]; phase 2: two x/y buffers per thread
    uint32_t slots[NSLOTS * SLOT_WORDS];      // slot item*CLIP_ROWS + r belongs to item `item` of the CTA, row r of its chunk:
                                              // [0] = z | vertices << 16 (3 bits) | bucket << 19 (5 bits) | area << 24 (0: empty), then x,y pairs
    uint32_t hdr[CLIP_ITEMS * HDR_WORDS];     // what the rows of one work item have in common
    uint16_t list[NSLOTS];                    // compacted ids of the slots that hold a row polygon
    uint32_t nb[NBUCK], cur[NBUCK];           // rows per (vertex count, columns spanned) bucket and fill cursors
    uint32_t nlist;
};

#define ZB(buf, v, k) sm.zbuf[(((buf) * ZCAP + RCB_IX(v, ZCAP)) * 3 + (k)) * CLIP_THREADS + threadIdx.x]

// One record into the segment of band `seg`; `lcell` is the cell inside the band's held rows.
// (Tried: one atomic per lane with the record written an iteration later, so that the x-chain runs during the round trip:
// k_clip_items 0.57 -> 2.69 ms on config 2 - six thousand same-address atomics per tile serialize in L2.  The warp-aggregated
// form below stays: one atomic per (warp, band).)
__device__ __forceinline__ void emit_to_segment(const Raster2& rb, const TileDesc& td, uint32_t seg, uint32_t lcell, uint32_t tcell, uint32_t tri,
                                                uint32_t mm, uint32_t area) {
    // lanes of a warp mostly hit the same band
    const unsigned act = __activemask();
    const unsigned peers = __match_any_sync(act, seg);
    const int leader = __ffs(peers) - 1;
    uint32_t base = 0;
    if ((int)lane_id() == leader) base = atomicAdd(&rb.tile_cursor[seg], (uint32_t)__popc(peers));
    base = __shfl_sync(peers, base, leader);
    const uint32_t pos = base + __popc(peers & ((1u << lane_id()) - 1u));
    if (pos < rb.tile_cap[seg]) {
        // the segment holds three arrays of `cap` words (cell | area << 24, smin | smax << 16, triangle): consecutive slots
        // of a warp are consecutive words of each
        const uint32_t cap = rb.tile_cap[seg];
        RCB_CHECK(tcell < (uint32_t)(td.width * td.height) && (lcell & 0xff000000u) == 0 && (uint32_t)(td.band_first) <= seg);
        uint32_t* rec = rb.tile_frags + 3ull * rb.tile_base[seg] + pos;
        rec[0] = lcell | (area << 24);
        rec[cap] = mm;
        rec[2ull * cap] = tri;
        // global back end: the per-column counts its scatter needs (a pass over the fragment list otherwise)
        if (rb.flags & 0x200u) atomicAdd(&rb.cell_cnt[td.cell_base + tcell], 1u);
    } else {
        atomicOr(&rb.counters[2], 2ull);
    }
}

__device__ __forceinline__ void emit_fragment(const Raster2& rb, const TileDesc& td, uint32_t tile, int x, int z,
                                              uint32_t tri, int smin, int smax, uint32_t area) {
    const uint32_t mm = ((uint32_t)smin & 0xffffu) | ((uint32_t)smax << 16);
    const uint32_t tcell = (uint32_t)z * td.width + x;
    if (rb.tile_cursor) {
        // per-band segments (tile-resident back end, and the global back end's per-tile scatter with whole-tile bands).
        // The band that owns row z, and - a band keeps one halo row on either side - the band above when z is its
        // neighbour's first row, the band below when z is its neighbour's last row.
        const int R = td.band_rows;
        if (R >= td.height) {  // the tile is one band
            emit_to_segment(rb, td, td.band_first, tcell, tcell, tri, mm, area);
        } else {
            const int b = (int)rcb_div_w((uint32_t)z, td.bmagic);
            const uint32_t seg = td.band_first + (uint32_t)b;
            emit_to_segment(rb, td, seg, (uint32_t)(z - max(b * R - 1, 0)) * td.width + x, tcell, tri, mm, area);
            if (z == b * R && b > 0) emit_to_segment(rb, td, seg - 1u, (uint32_t)(z - max((b - 1) * R - 1, 0)) * td.width + x, tcell, tri, mm, area);
            if (z == (b + 1) * R - 1 && z + 1 < td.height) emit_to_segment(rb, td, seg + 1u, (uint32_t)x, tcell, tri, mm, area);
        }
    } else {
        const uint32_t gcell = td.cell_base + tcell;
        const uint32_t rank = atomicAdd(&rb.cell_cnt[gcell], 1u);
        const unsigned m = __activemask();
        const int leader = __ffs(m) - 1;
        unsigned long long base = 0;
        if ((int)lane_id() == leader) base = atomicAdd(&rb.counters[1], (unsigned long long)__popc(m));
        base = __shfl_sync(m, base, leader);
        const unsigned long long slot = base + __popc(m & ((1u << lane_id()) - 1u));
        if (slot < rb.frag_cap)
            rb.frag_tmp[slot] = make_uint4(gcell, tri, mm, (rank & 0xffffffu) | (area << 24));
        else
            atomicOr(&rb.counters[2], 2ull);
    }
}

__global__ void __launch_bounds__(CLIP_THREADS, CLIP_CTAS) k_clip_items(Raster2 rb, const uint2* __restrict__ items, uint32_t nitems,
                                                             TileInfo* __restrict__ tinfo) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    ClipSmem& sm = *reinterpret_cast<ClipSmem*>(smem_raw);
    if (rb.spec) {  // issued before the host knew the item count: the grid covers the capacity, the count is on the device
        if (rb.counters[2] & RCB_FLAG_ABORT) return;
        nitems = (uint32_t)min((unsigned long long)nitems, rb.counters[4]);
        if (blockIdx.x * CLIP_ITEMS >= nitems) return;
    }
    const bool timed = (rb.flags & 0x100u) != 0 && threadIdx.x == 0;  // RCB_PHASE_TIMERS: counters[40..43]
    long long tclk = timed ? clock64() : 0;
    if (threadIdx.x < NBUCK) sm.nb[threadIdx.x] = sm.cur[threadIdx.x] = 0;
    if (threadIdx.x == 0) sm.nlist = 0;
#pragma unroll
    for (int r = 0; r < SLOTS_PER_THREAD; ++r) sm.slots[(threadIdx.x * SLOTS_PER_THREAD + r) * SLOT_WORDS] = 0;  // slot empty
    __syncthreads();

    // ---------------- phase 1: z-chain of one work item ------------------------------------------------
    const uint32_t it = blockIdx.x * CLIP_ITEMS + threadIdx.x;
    if (threadIdx.x < CLIP_ITEMS && it < nitems) {
        const uint2 item = items[it];
        const uint32_t tri = RCB_IX(item.x, rb.mesh.ntris), tile = RCB_IX(item.y >> 12, rb.ntiles), chunk = item.y & 0xfffu;
        const TileDesc& td = rb.tiles[tile];
        Tri t;
        int x0 = 0, x1 = 0, z0 = 0, z1 = 0;
        uint32_t area = rb.mesh.tri_areas ? (uint32_t)rb.mesh.tri_areas[tri] : 0u;
        bool ok = load_triangle(rb.mesh, tri, t) && tile_footprint(t, td, x0, x1, z0, z1);
        if (ok && !rb.mesh.tri_areas) ok = classify(t, td.slope_thr, area);  // lib.rs:258-260: degenerate => skipped
        if (ok) {
            sm.hdr[threadIdx.x * HDR_WORDS] = tri;
            sm.hdr[threadIdx.x * HDR_WORDS + 1] = tile;
            sm.hdr[threadIdx.x * HDR_WORDS + 2] = (uint32_t)x0 | ((uint32_t)x1 << 16);
            const int zs = max(z0, 0) + (int)chunk * CLIP_ROWS;
            const int ze = min(zs + CLIP_ROWS - 1, z1);
            bool ovf = false;
            int cur = 0, n = 3;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                ZB(0, 0, k) = t.v0[k];
                ZB(0, 1, k) = t.v1[k];
                ZB(0, 2, k) = t.v2[k];
            }
            // The polygons of one thread are a column of `zbuf` (element stride CLIP_THREADS): the loops below walk them
            // with pointers and carry vertex j into the next iteration's vertex i, instead of indexing (buffer, vertex,
            // component) afresh for every access (three multiply-adds per address, three to four loads per vertex).
            constexpr int ZV = 3 * CLIP_THREADS, ZBUF = ZCAP * ZV;  // floats per vertex / per buffer
            float* const zb = sm.zbuf + threadIdx.x;
            const float tz0 = td.bmin[2], tcs = td.cs, tics = td.ics;
            for (int z = z0; z <= ze && n > 0; ++z) {  // rasterization.rs:300-311
                const float cz = __fadd_rn(tz0, __fmul_rn((float)(z + 1), tcs));
                const bool keep = z >= zs;  // rows before the chunk only advance the chain
                const uint32_t slot_id = threadIdx.x * CLIP_ROWS + (uint32_t)(keep ? z - zs : 0);
                uint32_t* const slot = sm.slots + slot_id * SLOT_WORDS;
                uint32_t* ps = slot + 1;
                int nrow = 0, nrest = 0;
                float rminx = 3.0e38f, rmaxx = -3.0e38f;  // x-extent of the row polygon (bucketing hint)
                const float* pi = zb + cur * ZBUF;
                float* po = zb + (cur ^ 1) * ZBUF;
                const float v0x = pi[0], v0y = pi[CLIP_THREADS], v0z = pi[2 * CLIP_THREADS];
                float vix = v0x, viy = v0y, viz = v0z;
                // divide_poly(p1 -> in_row (slot, x/y only), rest (the other buffer))  rasterization.rs:172-242
                for (int i = 0; i < n; ++i) {
                    float vjx = v0x, vjy = v0y, vjz = v0z;
                    if (i + 1 < n) {
                        pi += ZV;
                        RCB_CHECK(pi + 2 * CLIP_THREADS < zb + (cur + 1) * ZBUF);
                        vjx = pi[0], vjy = pi[CLIP_THREADS], vjz = pi[2 * CLIP_THREADS];
                    }
                    const bool il = viz < cz, ig = viz > cz;  // si = il ? -1 : ig ? 1 : 0 (NaN: 0)
                    if (!ig) {
                        if (keep) {
                            if (nrow < RCAP) {
                                RCB_CHECK(ps + 1 < slot + SLOT_WORDS);
                                ps[0] = __float_as_uint(vix);
                                ps[1] = __float_as_uint(viy);
                                ps += 2;
                            }
                            rminx = fminf(rminx, vix);
                            rmaxx = fmaxf(rmaxx, vix);
                        }
                        nrow++;
                    }
                    if (!il) {
                        if (nrest < ZCAP) {
                            RCB_CHECK(po >= zb + (cur ^ 1) * ZBUF && po + 2 * CLIP_THREADS < zb + ((cur ^ 1) + 1) * ZBUF);
                            po[0] = vix, po[CLIP_THREADS] = viy, po[2 * CLIP_THREADS] = viz;
                            po += ZV;
                        }
                        nrest++;
                    }
                    if ((il && vjz > cz) || (ig && vjz < cz)) {
                        const float tt = __fdiv_rn(__fsub_rn(cz, viz), __fsub_rn(vjz, viz));
                        const float px = __fadd_rn(vix, __fmul_rn(tt, __fsub_rn(vjx, vix)));
                        const float py = __fadd_rn(viy, __fmul_rn(tt, __fsub_rn(vjy, viy)));
                        const float pz = __fadd_rn(viz, __fmul_rn(tt, __fsub_rn(vjz, viz)));
                        if (keep) {
                            if (nrow < RCAP) {
                                RCB_CHECK(ps + 1 < slot + SLOT_WORDS);
                                ps[0] = __float_as_uint(px);
                                ps[1] = __float_as_uint(py);
                                ps += 2;
                            }
                            rminx = fminf(rminx, px);
                            rmaxx = fmaxf(rmaxx, px);
                        }
                        nrow++;
                        if (nrest < ZCAP) {
                            RCB_CHECK(po >= zb + (cur ^ 1) * ZBUF && po + 2 * CLIP_THREADS < zb + ((cur ^ 1) + 1) * ZBUF);
                            po[0] = px, po[CLIP_THREADS] = py, po[2 * CLIP_THREADS] = pz;
                            po += ZV;
                        }
                        nrest++;
                    }
                    vix = vjx, viy = vjy, viz = vjz;
                }
                cur ^= 1;
                ovf = ovf || nrest > ZCAP || (keep && nrow > RCAP);  // a polygon outgrew its slot: counted, not stored
                n = min(nrest, ZCAP);
                if (keep && nrow >= 3 && z >= 0) {  // rasterization.rs:308-311
                    const int ncols = min(max(__float2int_rz((rmaxx - rminx) * tics), 0), 3);
                    const int bk = min(min(nrow, RCAP) - 3, 3) * 4 + ncols;
                    slot[0] = (uint32_t)z | ((uint32_t)min(nrow, RCAP) << 16) | ((uint32_t)bk << 19) | (area << 24);
                    // bucket by vertex count so the lanes of a warp run vertex loops of equal length
                    atomicAdd(&sm.nb[bk], 1u);
                }
            }
            if (ovf) {
                atomicOr(&tinfo[tile].status, 1u);
                atomicOr(&rb.counters[2], 1ull);
            }
        }
    }
    __syncthreads();
    if (timed) {
        const long long n = clock64();
        atomicAdd(&rb.counters[40], (unsigned long long)(n - tclk));
        tclk = n;
    }

    // ---------------- phase 2: x-chain, one row polygon per thread at a time ------------------------------
    // every thread files its own row slots into `list`, grouped by bucket
    uint32_t boff[NBUCK], nlist = 0;
#pragma unroll
    for (int k = 0; k < NBUCK; ++k) {
        boff[k] = nlist;
        nlist += sm.nb[k];
    }
#pragma unroll
    for (int r = 0; r < SLOTS_PER_THREAD; ++r) {
        const uint32_t slot_id = threadIdx.x * SLOTS_PER_THREAD + r;
        const uint32_t w0 = sm.slots[slot_id * SLOT_WORDS];
        if (((w0 >> 16) & 0x7u) >= 3) {
            const uint32_t bk = (w0 >> 19) & 0x1fu;
            uint32_t o = 0;
#pragma unroll
            for (int k = 0; k < NBUCK; ++k) o = (k == (int)bk) ? boff[k] : o;
            sm.list[RCB_IX(o + atomicAdd(&sm.cur[RCB_IX(bk, NBUCK)], 1u), NSLOTS)] = (uint16_t)slot_id;
        }
    }
    __syncthreads();
    if (timed) {
        const long long n = clock64();
        atomicAdd(&rb.counters[41], (unsigned long long)(n - tclk));
        tclk = n;
    }
    for (uint32_t li = threadIdx.x; li < nlist; li += CLIP_THREADS) {
        const uint32_t sid = RCB_IX(sm.list[li], NSLOTS);
        const uint32_t* slot = sm.slots + sid * SLOT_WORDS;
        const uint32_t* hdr = sm.hdr + (sid / CLIP_ROWS) * HDR_WORDS;
        const uint32_t tri = hdr[0], tile = hdr[1];
        const uint32_t h2 = slot[0], h3 = hdr[2];
        const int z = (int)(h2 & 0xffffu);
        int n = (int)((h2 >> 16) & 0x7u);
        const uint32_t area = h2 >> 24;
        const int x0 = (int)(h3 & 0xffffu), x1 = (int)(h3 >> 16);
        const TileDesc& td = rb.tiles[tile];
        constexpr int XV = 2 * CLIP_THREADS, XBUF = XCAP * XV;  // floats per vertex / per buffer
        float* const xb = sm.zbuf + threadIdx.x;
        const float tx0 = td.bmin[0], tcs = td.cs, ty0 = td.bmin[1], tich = td.ich;
        float minx = 0.f;
        bool has_nan = false;
        for (int v = 0; v < n; ++v) {
            const float vx = __uint_as_float(slot[1 + 2 * v]);
            RCB_CHECK(v < XCAP);
            xb[v * XV] = vx;
            xb[v * XV + CLIP_THREADS] = __uint_as_float(slot[2 + 2 * v]);
            has_nan = has_nan || (vx != vx);
            minx = v ? fminf(minx, vx) : vx;
        }
        // Exact skip: while every vertex lies strictly right of the cut, divide_poly returns the polygon unchanged and an
        // empty cell piece.  Only the leading columns of the footprint can be skipped (what is left behind a cut starts at
        // that cut), so the skipping is done here, ahead of the loop: the lanes of a warp then spend every iteration of
        // the loop on a column their row really crosses, instead of idling through each other's leading columns.
        int x = x0;
        if (!has_nan)
            while (x <= x1 && minx > __fadd_rn(tx0, __fmul_rn((float)(x + 1), tcs))) ++x;
        bool ovf = false;
        int cur = 0;
        for (; x <= x1 && n > 0; ++x) {  // rasterization.rs:317-362
            const float cx = __fadd_rn(tx0, __fmul_rn((float)(x + 1), tcs));
            int ncell = 0, nrest = 0;
            // The reference takes the first y as is and folds the others in with f32::min / max, which return the other
            // operand when one is NaN - exactly what fminf / fmaxf do from a NaN start.
            float miny = __int_as_float(0x7fc00000), maxy = miny;
            const float* pi = xb + cur * XBUF;
            float* po = xb + (cur ^ 1) * XBUF;
            const float v0x = pi[0], v0y = pi[CLIP_THREADS];
            float vix = v0x, viy = v0y;
            for (int i = 0; i < n; ++i) {
                float vjx = v0x, vjy = v0y;
                if (i + 1 < n) {
                    pi += XV;
                    RCB_CHECK(pi + CLIP_THREADS < xb + (cur + 1) * XBUF);
                    vjx = pi[0], vjy = pi[CLIP_THREADS];
                }
                const bool il = vix < cx, ig = vix > cx;  // si = il ? -1 : ig ? 1 : 0 (NaN: 0)
                if (!ig) {
                    miny = fminf(miny, viy);
                    maxy = fmaxf(maxy, viy);
                    ncell++;
                }
                if (!il) {
                    if (nrest < XCAP) {
                        RCB_CHECK(po >= xb + (cur ^ 1) * XBUF && po + CLIP_THREADS < xb + ((cur ^ 1) + 1) * XBUF);
                        po[0] = vix, po[CLIP_THREADS] = viy;
                        po += XV;
                    }
                    nrest++;
                }
                if ((il && vjx > cx) || (ig && vjx < cx)) {
                    const float tt = __fdiv_rn(__fsub_rn(cx, vix), __fsub_rn(vjx, vix));
                    const float px = __fadd_rn(vix, __fmul_rn(tt, __fsub_rn(vjx, vix)));
                    const float py = __fadd_rn(viy, __fmul_rn(tt, __fsub_rn(vjy, viy)));
                    miny = fminf(miny, py);
                    maxy = fmaxf(maxy, py);
                    ncell++;
                    if (nrest < XCAP) {
                        RCB_CHECK(po >= xb + (cur ^ 1) * XBUF && po + CLIP_THREADS < xb + ((cur ^ 1) + 1) * XBUF);
                        po[0] = px, po[CLIP_THREADS] = py;
                        po += XV;
                    }
                    nrest++;
                }
                vix = vjx, viy = vjy;
            }
            cur ^= 1;
            ovf = ovf || nrest > XCAP;
            n = min(nrest, XCAP);
            if (ncell >= 3) {  // rasterization.rs:325-361
                int smin = __float2int_rz(floorf(__fmul_rn(__fsub_rn(miny, ty0), tich)));
                int smax = __float2int_rz(ceilf(__fmul_rn(__fsub_rn(maxy, ty0), tich)));
                smin = min(max(smin, -32768), 32767);
                smax = min(max(smax, -32768), 32767);
                smin = max(smin, 0);
                emit_fragment(rb, td, tile, x, z, tri, smin, smax, area);
            }
        }
        if (ovf) {
            atomicOr(&tinfo[tile].status, 1u);
            atomicOr(&rb.counters[2], 1ull);
        }
    }
    if (timed) {  // thread 0's own share of phase 2 (no barrier behind it: the CTA ends when its slowest thread does)
        atomicAdd(&rb.counters[42], (unsigned long long)(clock64() - tclk));
        atomicAdd(&rb.counters[43], 1ull);
    }
}

}  // namespace

RCB_BOUNDS_EXPORT(rcb_bounds_read_raster2)

void launch_bin_count(const Raster2& rb, cudaStream_t s) {
    if (rb.mesh.ntris == 0) return;
    KScope _k("k_bin_count", s);
    k_bin_count<<<(rb.mesh.ntris + 255) / 256, 256, 0, s>>>(rb);
}

void launch_check_indices(const int32_t* tris, size_t nidx, uint32_t nverts, unsigned long long* flag, cudaStream_t s) {
    if (nidx == 0) return;
    KScope _k("k_check_indices", s);
    const unsigned blocks = (unsigned)std::min<size_t>((nidx + 255) / 256, 148 * 16);
    k_check_indices<<<blocks, 256, 0, s>>>(tris, nidx, nverts, flag);
}

void launch_widen_indices(const uint16_t* src, int32_t* dst, size_t n, cudaStream_t s) {
    if (n == 0) return;
    KScope _k("k_widen_indices", s);
    k_widen_indices<<<(unsigned)std::min<size_t>((n + 255) / 256, 148 * 16), 256, 0, s>>>(src, dst, n);
}

void launch_mesh_blocks(const float* verts, const int32_t* tris, uint32_t ntris, uint32_t nverts, float4* blocks, cudaStream_t s) {
    if (ntris == 0) return;
    KScope _k("k_mesh_blocks", s);
    k_mesh_blocks<<<(ntris + RCB_MESH_BLOCK - 1) / RCB_MESH_BLOCK, RCB_MESH_BLOCK, 0, s>>>(verts, tris, ntris, nverts, blocks);
}

void launch_fill_items(const Raster2& rb, const uint32_t* item_off, uint2* items, cudaStream_t s) {
    if (rb.mesh.ntris == 0) return;
    KScope _k("k_fill_items", s);
    k_fill_items<<<(rb.mesh.ntris + 255) / 256, 256, 0, s>>>(rb, item_off, items);
}

void launch_clip_items(const Raster2& rb, const uint2* items, uint32_t nitems, TileInfo* tinfo, cudaStream_t s) {
    if (nitems == 0) return;
    rcb_raise_dynamic_smem((const void*)k_clip_items, sizeof(ClipSmem));  // per (kernel, device); a process may drive several
    KScope _k("k_clip_items", s);
    k_clip_items<<<(nitems + CLIP_ITEMS - 1) / CLIP_ITEMS, CLIP_THREADS, sizeof(ClipSmem), s>>>(rb, items, nitems, tinfo);
}
