// k_compact.cu — compact heightfield: cells, span attributes, 8-direction linking, connection list.
//
// Replaces (crates/recast/src/compact_heightfield.rs):
//   :175-281 build_from_heightfield  (cells row-major, ALL spans kept, areas copy, max_height,
//                                     walkable_span_count)
//   :284-430 build_connections       (best-overlap walkable neighbour in 8 directions, connection
//                                     array in reverse direction order per span, con[4])
//   :433-442 get_cell_index_for_span (O(cells) scan in the reference; the neighbour cell is known here)
// The heightfield CSR (hf_off) is already the compaction prefix sum, so "compaction" is a gather of
// offsets; linking is a per-span stencil over the 8 neighbour columns (one thread per span, walkable spans compacted
// per block, ordered columns bisected).
#include <algorithm>

#include "rcb_common.cuh"

namespace {

__constant__ int c_dirx[8] = {-1, 0, 1, 1, 1, 0, -1, -1};  // compact_heightfield.rs:109-111
__constant__ int c_dirz[8] = {-1, -1, -1, 0, 1, 1, 1, 0};

__global__ void k_tile_bases(const TileDesc* __restrict__ tiles, int ntiles, const uint32_t* __restrict__ hf_off,
                             const uint32_t* __restrict__ conn_off, TileInfo* __restrict__ tinfo) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ntiles) return;
    const uint32_t c0 = tiles[t].cell_base, c1 = c0 + (uint32_t)(tiles[t].width * tiles[t].height);
    if (hf_off) {
        tinfo[t].span_base = hf_off[c0];
        tinfo[t].span_count = hf_off[c1] - hf_off[c0];
    }
    if (conn_off) {
        tinfo[t].conn_base = conn_off[c0];
        tinfo[t].conn_count = conn_off[c1] - conn_off[c0];
    }
}

__global__ void __launch_bounds__(256) k_cells(const TileDesc* __restrict__ tiles, TileMap tm,
                                               const uint32_t* __restrict__ hf_off, uint32_t* __restrict__ hf_cell_offset,
                                               uint32_t* __restrict__ cell_index, uint32_t* __restrict__ cell_count,
                                               const TileInfo* __restrict__ tinfo) {
    const uint32_t tile = RCB_TILE_OF_BLOCK();
    if (tile >= tm.ntiles) return;
    const TileDesc& td = tiles[tile];
    const uint32_t C = (uint32_t)(td.width * td.height);
    const uint32_t lc = RCB_CELL_OF_THREAD();
    if (lc >= C) return;
    const uint32_t g = td.cell_base + lc;
    const uint32_t base = tinfo[tile].span_base;
    const uint32_t b = hf_off[g], e = hf_off[g + 1];
    uint32_t* off = hf_cell_offset + td.cell_base + tile;  // each tile owns C+1 entries
    off[lc] = b - base;
    if (lc == C - 1) off[C] = e - base;
    if (cell_index) {
        cell_index[g] = (e > b) ? (b - base) : RCB_NONE;  // compact_heightfield.rs:209-223
        cell_count[g] = e - b;
    }
}

// ---- span-parallel linking ---------------------------------------------------------------------------------
// One thread per SPAN instead of one per column: a column of a multi-level interior holds 16+ spans, each probing up to
// 8 x 16 neighbour spans, and a thread that walks its column alone leaves the machine mostly empty (config 4: 590 k
// columns, 8 M spans).  k_span_cells records (tile, cell) per span; the connection offsets are then a scan over spans.
// It also finds out whether every column is ordered (min and max both non-decreasing), which is what the rasterizer
// produces but Heightfield::add_span and a caller's CSR need not: span_cell[S].x != 0 = some column is not.  In an ordered
// column the spans that can overlap / face a given span are one contiguous run, found by bisection (k_link_spans2,
// k_filter_spans2) instead of a walk over the whole column.
__global__ void __launch_bounds__(256) k_span_cells(const TileDesc* __restrict__ tiles, TileMap tm, const uint32_t* __restrict__ hf_off,
                                                    const int16_t* __restrict__ smin, const int16_t* __restrict__ smax, size_t S,
                                                    uint2* __restrict__ span_cell) {
    const uint32_t tile = RCB_TILE_OF_BLOCK();
    if (tile >= tm.ntiles) return;
    const TileDesc& td = tiles[tile];
    const uint32_t lc = RCB_CELL_OF_THREAD();
    if (lc >= (uint32_t)(td.width * td.height)) return;
    const uint32_t g = td.cell_base + lc;
    const uint32_t b = hf_off[g], e = hf_off[g + 1];
    bool unordered = false;
    int pmn = -0x8000, pmx = -0x8000;
    for (uint32_t s = b; s < e; ++s) {
        span_cell[s] = make_uint2(tile, lc);
        if (e - b > 1) {
            const int mn = smin[s], mx = smax[s];
            unordered |= mn < pmn || mx < pmx;
            pmn = mn;
            pmx = mx;
        }
    }
    if (unordered) span_cell[S].x = 1u;
}

// Only walkable spans link (:262); on real geometry they are about half of all spans, scattered through every warp.  Each
// thread first settles its own span's trivial outputs, then the block's walkable spans are compacted so that the eight
// neighbour searches run in full warps.
__global__ void __launch_bounds__(256) k_link_spans2(const TileDesc* __restrict__ tiles, uint32_t S, const uint2* __restrict__ span_cell,
                                                     const uint32_t* __restrict__ hf_off, const int16_t* __restrict__ smin,
                                                     const int16_t* __restrict__ smax, const uint8_t* __restrict__ sarea,
                                                     uint8_t* __restrict__ areas, uint32_t* __restrict__ nid, uint16_t* __restrict__ con,
                                                     uint32_t* __restrict__ span_conn_cnt, TileInfo* __restrict__ tinfo) {
    __shared__ uint32_t sh_list[256], sh_wcnt[8];
    const uint32_t s0 = blockIdx.x * blockDim.x;
    bool walk = false;
    {
        const uint32_t s = s0 + threadIdx.x;
        if (s < S) {
            const uint32_t tile = span_cell[s].x;
            const uint32_t a = sarea[s];
            areas[s] = (uint8_t)a;  // :237
            // per-tile reductions, one atomic per (warp, tile): :240, :243-245
            const unsigned peers = __match_any_sync(__activemask(), tile);
            const uint32_t wsum = __reduce_add_sync(peers, a != 0 ? 1u : 0u);
            const uint32_t hmax = __reduce_max_sync(peers, (uint32_t)(uint16_t)smax[s]);
            if ((int)(threadIdx.x & 31) == __ffs(peers) - 1) {
                if (wsum) atomicAdd(&tinfo[tile].walkable_span_count, wsum);
                if (hmax) atomicMax(&tinfo[tile].max_height, hmax);
            }
            walk = a != 0;
            if (!walk) {
#pragma unroll
                for (int d = 0; d < 8; ++d) nid[(size_t)d * S + s] = RCB_NONE;
                *reinterpret_cast<uint2*>(con + 4 * (size_t)s) = make_uint2(63u | (63u << 16), 63u | (63u << 16));
                span_conn_cnt[s] = 0;
            }
        }
    }
    const uint32_t nwalk = compact_block256(walk, sh_list, sh_wcnt);
    if (threadIdx.x >= nwalk) return;
    const uint32_t s = s0 + sh_list[threadIdx.x];
    const uint2 sc = span_cell[s];
    const bool ordered = span_cell[S].x == 0;
    const uint32_t lc = sc.y;
    const TileDesc& td = tiles[sc.x];
    const int W = td.width, H = td.height;
    const int z = (int)(lc / (uint32_t)W), x = (int)lc - z * W;
    const int mn = smin[s], mx = smax[s];
    uint16_t c4[4] = {63, 63, 63, 63};
    uint32_t nconn = 0;
#pragma unroll
    for (int d = 0; d < 8; ++d) {
        const int nx = x + c_dirx[d], nz = z + c_dirz[d];
        uint32_t best = RCB_NEI_NONE, nb = 0;
        if (!(nx < 0 || nz < 0 || nx >= W || nz >= H)) {
            const uint32_t ng = td.cell_base + (uint32_t)nz * W + nx;
            nb = hf_off[ng];
            const uint32_t ne = hf_off[ng + 1];
            int best_diff = 0x7fffffff;
            uint32_t ns = nb;
            if (ordered && ne - nb > 2) {  // first span whose top reaches this one's bottom
                uint32_t hi = ne;
                while (ns < hi) {
                    const uint32_t mid = (ns + hi) >> 1;
                    if ((int)smax[mid] < mn)
                        ns = mid + 1;
                    else
                        hi = mid;
                }
            }
            for (; ns < ne; ++ns) {
                const int nmn = smin[ns], nmx = smax[ns];
                if (ordered && nmn > mx) break;  // nor can any later one overlap
                if (sarea[ns] == 0) continue;
                if (max(nmn, mn) > min(nmx, mx)) continue;
                const int diff = (int)(uint16_t)(abs(nmn - mn) + abs(nmx - mx));
                if (diff < best_diff) {
                    best_diff = diff;
                    best = ns - nb;
                }
            }
        }
        nid[(size_t)d * S + s] = best != RCB_NEI_NONE ? nb + best : RCB_NONE;  // absolute span id for every later stage
        if (best != RCB_NEI_NONE) {
            nconn++;
            if (d == 7) c4[0] = (uint16_t)best;  // :404-410
            if (d == 5) c4[1] = (uint16_t)best;
            if (d == 3) c4[2] = (uint16_t)best;
            if (d == 1) c4[3] = (uint16_t)best;
        }
    }
    *reinterpret_cast<uint2*>(con + 4 * (size_t)s) =
        make_uint2((uint32_t)c4[0] | ((uint32_t)c4[1] << 16), (uint32_t)c4[2] | ((uint32_t)c4[3] << 16));
    span_conn_cnt[s] = nconn;
}

__global__ void k_tile_conn_bases(int ntiles, const uint32_t* __restrict__ conn_off_span, TileInfo* __restrict__ tinfo) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ntiles) return;
    const uint32_t b = tinfo[t].span_base, e = b + tinfo[t].span_count;
    tinfo[t].conn_base = conn_off_span[b];
    tinfo[t].conn_count = conn_off_span[e] - conn_off_span[b];
}

// The connections of 256 consecutive spans are one contiguous range of the list: build it in shared memory, store it coalesced.
__global__ void __launch_bounds__(256) k_write_connections2(uint32_t S, const uint2* __restrict__ span_cell, const uint32_t* __restrict__ nid,
                                                            const uint32_t* __restrict__ conn_off_span, const TileInfo* __restrict__ tinfo,
                                                            uint32_t* __restrict__ first_conn, uint32_t* __restrict__ conn_ref,
                                                            uint8_t* __restrict__ conn_dir, uint32_t* __restrict__ conn_next) {
    __shared__ uint32_t sh_ref[2048], sh_next[2048];
    __shared__ uint8_t sh_dir[2048];
    const uint32_t s0 = blockIdx.x * 256u, s = s0 + threadIdx.x;
    const uint32_t k0 = conn_off_span[s0], k1 = conn_off_span[min(S, s0 + 256u)];
    if (s < S) {
        const uint32_t tile = span_cell[s].x;
        const uint32_t span_base = tinfo[tile].span_base, conn_base = tinfo[tile].conn_base;
        uint32_t k = conn_off_span[s];
        uint32_t first = RCB_NONE;
#pragma unroll
        for (int d = 7; d >= 0; --d) {
            const uint32_t n = nid[(size_t)d * S + s];
            if (n == RCB_NONE) continue;
            sh_ref[k - k0] = n - span_base;
            sh_dir[k - k0] = (uint8_t)d;
            sh_next[k - k0] = first;
            first = k - conn_base;
            ++k;
        }
        first_conn[s] = first;
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < k1 - k0; i += 256u) {
        conn_ref[k0 + i] = sh_ref[i];
        conn_next[k0 + i] = sh_next[i];
        conn_dir[k0 + i] = sh_dir[i];
    }
}

// Slim result layout (RCB_RESULT_SLIM): what the host cannot rebuild from the layout rule of compact_heightfield.rs:375-389
// is WHICH directions a span is connected in - one byte per span (conn_mask) - and, per connection, which span of the
// neighbour cell it reaches: its offset inside that cell, one byte (conn_off8; the cell is known from the direction, its
// first span from the cell counts).  first_connection, next, dir, ref and con all follow.  The wider forms (conn_ref16, con8)
// are derived as well and travel only when an offset passes 255; flag bits 8 / 1 / 2 / 4 say what did not fit
// (offset in 8 bits, reference in 16, con in 8, cell count in 16).
__global__ void __launch_bounds__(256) k_slim_pack(const TileDesc* __restrict__ tiles, const TileInfo* __restrict__ tinfo,
                                                   const uint32_t* __restrict__ first_conn, const uint32_t* __restrict__ conn_ref,
                                                   const uint8_t* __restrict__ conn_dir, const uint32_t* __restrict__ conn_next,
                                                   const uint32_t* __restrict__ cell_index, const uint32_t* __restrict__ cell_count,
                                                   const uint16_t* __restrict__ con, uint8_t* __restrict__ conn_mask,
                                                   uint16_t* __restrict__ conn_ref16, uint8_t* __restrict__ conn_off8,
                                                   uint16_t* __restrict__ cell_count16, uint8_t* __restrict__ con8,
                                                   unsigned long long* __restrict__ flag) {
    const TileInfo ti = tinfo[blockIdx.x];
    const TileDesc& td = tiles[blockIdx.x];
    const int W = td.width;
    const uint32_t C = (uint32_t)(W * td.height);
    unsigned wide = 0;
    for (uint32_t c = blockIdx.y * 256u + threadIdx.x; c < C; c += gridDim.y * 256u) {
        const uint32_t n = cell_count[td.cell_base + c];
        if (n > 0xffffu) wide |= 4u;
        cell_count16[td.cell_base + c] = (uint16_t)n;
        if (n == 0) continue;
        const uint32_t first = cell_index[td.cell_base + c];
        for (uint32_t s = first; s < first + n; ++s) {
            uint32_t mask = 0;
            for (uint32_t k = first_conn[ti.span_base + s]; k != RCB_NONE; k = conn_next[ti.conn_base + k]) {
                const uint32_t d = conn_dir[ti.conn_base + k];
                mask |= 1u << d;
                // the neighbour cell of direction d exists (there is a connection into it) and is not empty
                const uint32_t nc = (uint32_t)((int)c + c_dirz[d] * W + c_dirx[d]);
                const uint32_t r = conn_ref[ti.conn_base + k];
                const uint32_t off = r - cell_index[td.cell_base + nc];
                if (off > 0xffu) wide |= 8u;
                if (r > 0xffffu) wide |= 1u;
                conn_off8[ti.conn_base + k] = (uint8_t)off;
                conn_ref16[ti.conn_base + k] = (uint16_t)r;
            }
            conn_mask[ti.span_base + s] = (uint8_t)mask;
            const uint2 c4 = *reinterpret_cast<const uint2*>(con + 4ull * (ti.span_base + s));  // 4 x u16, 8-byte aligned
            const uint32_t c0 = c4.x & 0xffffu, c1 = c4.x >> 16, c2 = c4.y & 0xffffu, c3 = c4.y >> 16;
            if ((c0 | c1 | c2 | c3) > 0xffu) wide |= 2u;
            *reinterpret_cast<uint32_t*>(con8 + 4ull * (ti.span_base + s)) = (c0 & 0xffu) | ((c1 & 0xffu) << 8) | ((c2 & 0xffu) << 16) | (c3 << 24);
        }
    }
    wide = __reduce_or_sync(0xffffffffu, wide);
    if (wide && (threadIdx.x & 31) == 0) atomicOr(flag, (unsigned long long)wide);
}

}  // namespace

void launch_slim_pack(const TileDesc* tiles, const TileInfo* tinfo, int ntiles, uint32_t max_cells, const uint32_t* first_conn,
                      const uint32_t* conn_ref, const uint8_t* conn_dir, const uint32_t* conn_next, const uint32_t* cell_index,
                      const uint32_t* cell_count, const uint16_t* con, uint8_t* conn_mask, uint16_t* conn_ref16, uint8_t* conn_off8,
                      uint16_t* cell_count16, uint8_t* con8, unsigned long long* flag, cudaStream_t s) {
    if (ntiles == 0) return;
    KScope _k("k_slim_pack", s);
    const unsigned gy = std::max(1u, std::min(64u, (max_cells + 1023u) / 1024u));  // a block takes ~1 k cells
    for (int t0 = 0; t0 < ntiles; t0 += 65535) {
        const int n = std::min(65535, ntiles - t0);
        k_slim_pack<<<dim3((unsigned)n, gy), 256, 0, s>>>(tiles + t0, tinfo + t0, first_conn, conn_ref, conn_dir, conn_next, cell_index, cell_count,
                                                          con, conn_mask, conn_ref16, conn_off8, cell_count16, con8, flag);
    }
}

void launch_span_cells(const TileDesc* tiles, TileMap tm, const uint32_t* hf_off, const int16_t* smin, const int16_t* smax, size_t S,
                       uint2* span_cell, cudaStream_t s) {
    if (tm.nblocks == 0) return;
    cudaMemsetAsync(span_cell + S, 0, sizeof(uint2), s);  // the "some column is not ordered" flag
    KScope _k("k_span_cells", s);
    k_span_cells<<<tm.grid(), 256, 0, s>>>(tiles, tm, hf_off, smin, smax, S, span_cell);
}

void launch_link_spans2(const TileDesc* tiles, size_t S, const uint2* span_cell, const uint32_t* hf_off, const int16_t* smin,
                        const int16_t* smax, const uint8_t* sarea, uint8_t* areas, uint32_t* nid, uint16_t* con, uint32_t* span_conn_cnt,
                        TileInfo* tinfo, cudaStream_t s) {
    if (S == 0) return;
    KScope _k("k_link_spans2", s);
    k_link_spans2<<<(unsigned)((S + 255) / 256), 256, 0, s>>>(tiles, (uint32_t)S, span_cell, hf_off, smin, smax, sarea, areas, nid, con,
                                                              span_conn_cnt, tinfo);
}

void launch_tile_conn_bases(int ntiles, const uint32_t* conn_off_span, TileInfo* tinfo, cudaStream_t s) {
    if (ntiles == 0) return;
    KScope _k("k_tile_conn_bases", s);
    k_tile_conn_bases<<<(ntiles + 127) / 128, 128, 0, s>>>(ntiles, conn_off_span, tinfo);
}

void launch_write_connections2(size_t S, const uint2* span_cell, const uint32_t* nid, const uint32_t* conn_off_span, const TileInfo* tinfo, uint32_t* first_conn, uint32_t* conn_ref,
                               uint8_t* conn_dir, uint32_t* conn_next, cudaStream_t s) {
    if (S == 0) return;
    KScope _k("k_write_connections2", s);
    k_write_connections2<<<(unsigned)((S + 255) / 256), 256, 0, s>>>((uint32_t)S, span_cell, nid, conn_off_span, tinfo,
                                                                     first_conn, conn_ref, conn_dir, conn_next);
}

void launch_tile_bases(const TileDesc* tiles, int ntiles, const uint32_t* hf_off, const uint32_t* conn_off,
                       TileInfo* tinfo, cudaStream_t s) {
    if (ntiles == 0) return;
    {
        KScope _k("k_tile_bases", s);
        k_tile_bases<<<(ntiles + 127) / 128, 128, 0, s>>>(tiles, ntiles, hf_off, conn_off, tinfo);
    }
}

void launch_cells_and_offsets(const TileDesc* tiles, TileMap tm, const uint32_t* hf_off, uint32_t* hf_cell_offset,
                              uint32_t* cell_index, uint32_t* cell_count, TileInfo* tinfo, int, cudaStream_t s) {
    if (tm.nblocks == 0) return;
    {
        KScope _k("k_cells", s);
        k_cells<<<tm.grid(), 256, 0, s>>>(tiles, tm, hf_off, hf_cell_offset, cell_index, cell_count, tinfo);
    }
}
