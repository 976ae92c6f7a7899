// k_wire.cu — heightfield wire formats of detour-dynamic, batched over tiles.
//
//   VoxelTile::serialize_spans          io/voxel_tile.rs:70-118   CSR -> bytes (big-endian)
//   VoxelTile::deserialize_spans        io/voxel_tile.rs:120-176  bytes (big-endian, lenient) -> Heightfield::add_span
//   DynamicTile::reconstruct_heightfield dynamic_tile.rs:147-257  bytes (little-endian, strict) -> Heightfield::add_span
//   DynamicTileCheckpoint::clone_heightfield checkpoint.rs:29-60  spans -> Heightfield::add_span
//
// A tile's stream is `u16 count, count x (u32 min, u32 max, u32 area)` per cell, z-major.  Where a cell
// starts depends on every count before it, so reading is a pointer chase: `k_wire_walk` does that chase
// (one thread per tile, all tiles in flight) and leaves (position, count) per cell; everything after it is
// one thread per column: `k_wire_replay` re-inserts the column's spans with the rule of
// Heightfield::add_span (heightfield.rs:102-222 — same-area overlap merges and absorbs at most ONE
// follower; this is not the rasterizer's rule) on an array form of the list.
#include "rcb_common.cuh"

namespace {

__device__ __forceinline__ uint32_t rd16(const uint8_t* p, bool be) {
    return be ? ((uint32_t)p[0] << 8) | p[1] : (uint32_t)p[0] | ((uint32_t)p[1] << 8);
}
// streams are 2-byte aligned at best: assemble words from halfwords when the base allows, else bytes
__device__ __forceinline__ uint32_t rd32(const uint8_t* p, bool be) {
    const uint32_t b0 = p[0], b1 = p[1], b2 = p[2], b3 = p[3];
    return be ? (b0 << 24) | (b1 << 16) | (b2 << 8) | b3 : b0 | (b1 << 8) | (b2 << 16) | (b3 << 24);
}

// err: the first failure of the strict reader in (tile, cell, step) order, as tile << 33 | cell << 1 | kind;
// kind 0 = short buffer (checked before the cell's spans are added), 1 = add_span refused a span
//
// One warp per tile.  The warp stages the stream through shared memory in 4 KB pieces (coalesced 16-byte
// loads).  A serial chase (lane 0) costs ~270 cycles of dependent instructions per cell (0.58 ms for the 2025
// tiles of config 2; 1.5 ms when it chased through global memory), so the warp first tries a speculative
// walk, 32 cells at a time, and chases serially only across stretches where that fails.
constexpr int WALK_CHUNK = 4096;  // bytes staged per refill
constexpr int WALK_CELLS = 512;   // cells buffered before a flush
constexpr int WALK_SERIAL = 16;   // cells the serial chase crosses before the speculative walk gets another try

__global__ void __launch_bounds__(32) k_wire_walk(const TileDesc* __restrict__ tiles, int ntiles, const uint8_t* __restrict__ data,
                                                  const unsigned long long* __restrict__ tile_off, int strict_le,
                                                  uint32_t* __restrict__ cell_pos, uint32_t* __restrict__ cell_n,
                                                  unsigned long long* __restrict__ err) {
    __shared__ __align__(16) uint8_t buf[WALK_CHUNK + 16];
    __shared__ uint32_t s_pos[WALK_CELLS], s_n[WALK_CELLS];
    const int tile = blockIdx.x;
    const int lane = threadIdx.x;
    const TileDesc td = tiles[tile];
    const uint32_t C = (uint32_t)(td.width * td.height);
    const uint8_t* d = data + tile_off[tile];
    const uint32_t len = (uint32_t)(tile_off[tile + 1] - tile_off[tile]);  // < 4 GiB (checked by the host)
    uint32_t* cp = cell_pos + td.cell_base;
    uint32_t* cn = cell_n + td.cell_base;
    // 32-bit offsets throughout: the chase is a chain of dependent instructions, 64-bit ones double it.
    // Offsets are kept relative to `org`, the 16-byte line at or before the stream's first byte.
    const uint32_t skew = (uint32_t)((uintptr_t)d & 15);
    const uint8_t* org = d - skew;
    const uint32_t end = len + skew;  // one past the last byte, org-relative
    uint32_t pos = skew;              // lane 0's copy is the truth; broadcast after every walk
    uint32_t lo = 0, hi = 0;          // org-relative offsets held in buf: [lo, hi); empty at the start
    uint32_t c = 0;
    int state = (strict_le && len < 2ull * C) ? 2 : 0;  // 0 walking, 1 lenient reader ran dry, 2 strict reader failed (dynamic_tile.rs:171-175)
    uint32_t bad_cell = 0;
    uint32_t stride = 14;  // guess: one span per cell
    while (c < C && state == 0) {
        if (end - pos >= 2 && pos + 2 > hi) {  // refill from the 16-byte line holding `pos`
            lo = pos & ~15u;
            hi = end - lo > WALK_CHUNK ? lo + WALK_CHUNK : end;
            __syncwarp();
            for (uint32_t k = lane; k < WALK_CHUNK / 16; k += 32)
                if (lo + 16u * k < hi) ((uint4*)buf)[k] = __ldg((const uint4*)(org + lo) + k);  // the buffer carries 64 spare bytes
            __syncwarp();
        }
        // Speculative stretch.  Lane i guesses where cell c+i starts (the previous round's scan, else a constant
        // stride), reads the count there, and an exclusive scan of the cell lengths says where each cell really
        // starts IF every earlier lane guessed right: the longest prefix of confirmed lanes is accepted (lane 0
        // is always right), the rest re-guess from the scan.  Uniform columns go 32 cells a round; anything
        // that is not a plain cell (window edge, spans past the end of the stream) stops the stretch and is
        // left to the serial code below, which also owns the readers' end-of-data semantics.
        {
            uint32_t g = pos + (uint32_t)lane * stride;
            for (;;) {
                const bool inwin = c + lane < C && g >= lo && g + 2 <= hi && g + 2 >= g;
                uint32_t count = 0, ln = 0;
                bool plain = false;
                if (inwin) {
                    const uint32_t o = g - lo;
                    const uint32_t w0 = *(const uint32_t*)(buf + (o & ~3u)), w1 = *(const uint32_t*)(buf + (o & ~3u) + 4);
                    const uint32_t h = __funnelshift_r(w0, w1, (o & 3u) * 8u);
                    count = strict_le ? (h & 0xffffu) : (__byte_perm(h, 0, 0x4401));
                    plain = 12u * count <= end - (g + 2);
                    ln = 2u + 12u * count;
                }
                uint32_t incl = ln;
#pragma unroll
                for (int dd = 1; dd < 32; dd <<= 1) {
                    const uint32_t v = __shfl_up_sync(0xffffffffu, incl, dd);
                    if (lane >= dd) incl += v;
                }
                const uint32_t at = pos + incl - ln;  // where this cell starts if all earlier lanes were right
                const unsigned wrong = __ballot_sync(0xffffffffu, !(plain && g == at));
                const uint32_t nv = wrong ? (uint32_t)__ffs(wrong) - 1u : 32u;
                if ((uint32_t)lane < nv) {
                    cp[c + lane] = g + 2 - skew;
                    cn[c + lane] = count;
                }
                if (nv == 0) break;
                stride = __shfl_sync(0xffffffffu, ln, nv - 1);
                pos += __shfl_sync(0xffffffffu, incl, nv - 1);
                c += nv;
                if (nv < 4) break;  // irregular stretch: cross it serially
                if (nv == 32) {
                    g = pos + (uint32_t)lane * stride;
                } else {
                    const uint32_t shifted = __shfl_down_sync(0xffffffffu, at, nv);
                    const uint32_t last = __shfl_sync(0xffffffffu, at + ln, 31);
                    g = (uint32_t)lane < 32u - nv ? shifted : last + ((uint32_t)lane - (32u - nv)) * stride;
                }
            }
        }
        if (!(c < C)) break;
        uint32_t nb = 0;
        if (lane == 0) {
            while (c < C && nb < WALK_SERIAL) {
                if (end - pos < 2) {  // dynamic_tile.rs:181-189; the lenient reader just stops (voxel_tile.rs:129-131)
                    state = strict_le ? 2 : 1;
                    bad_cell = c;
                    break;
                }
                if (pos + 2 > hi) break;  // refill
                const uint32_t o = pos - lo;
                const uint32_t w0 = *(const uint32_t*)(buf + (o & ~3u)), w1 = *(const uint32_t*)(buf + (o & ~3u) + 4);
                const uint32_t h = __funnelshift_r(w0, w1, (o & 3u) * 8u);
                const uint32_t count = strict_le ? (h & 0xffffu) : (__byte_perm(h, 0, 0x4401));
                pos += 2;
                const uint32_t room = end - pos;
                uint32_t n = count;
                if (12u * count > room) {
                    if (strict_le) {  // dynamic_tile.rs:207-216
                        state = 2;
                        bad_cell = c;
                        break;
                    }
                    n = room / 12u;  // voxel_tile.rs:138-140
                }
                s_pos[nb] = pos - skew;
                s_n[nb] = n;
                ++nb;
                ++c;
                pos += 12u * n;
            }
        }
        const uint32_t c_new = __shfl_sync(0xffffffffu, c, 0);
        nb = __shfl_sync(0xffffffffu, nb, 0);
        pos = __shfl_sync(0xffffffffu, pos, 0);
        state = __shfl_sync(0xffffffffu, state, 0);
        bad_cell = __shfl_sync(0xffffffffu, bad_cell, 0);
        __syncwarp();
        const uint32_t c0 = c_new - nb;
        for (uint32_t k = lane; k < nb; k += 32) cp[c0 + k] = s_pos[k], cn[c0 + k] = s_n[k];
        c = c_new;
        __syncwarp();
    }
    for (uint32_t k = c + lane; k < C; k += 32) cp[k] = 0, cn[k] = 0;  // nothing is read past the point of failure
    if (state == 2 && lane == 0) atomicMin(err, ((unsigned long long)tile << 33) | ((unsigned long long)bad_cell << 1));
}

struct ReplaySrc {
    const uint8_t* data;  // byte streams (nullptr: CSR source)
    const unsigned long long* tile_off;
    const uint32_t* cell_pos;
    int strict_le;
    const int16_t* smin;  // CSR source, indexed like the input offsets
    const int16_t* smax;
    const uint8_t* sarea;
};

__global__ void __launch_bounds__(256) k_wire_replay(const TileDesc* __restrict__ tiles, TileMap tm, ReplaySrc src,
                                                     const uint32_t* __restrict__ in_off, uint32_t* __restrict__ mm,
                                                     uint8_t* __restrict__ ar, uint32_t* __restrict__ span_cnt,
                                                     unsigned long long* __restrict__ err) {
    const uint32_t tile = RCB_TILE_OF_BLOCK();
    if (tile >= tm.ntiles) return;
    const TileDesc& td = tiles[tile];
    const uint32_t cl = RCB_CELL_OF_THREAD();
    if (cl >= (uint32_t)(td.width * td.height)) return;
    const uint32_t c = td.cell_base + cl;
    const uint32_t b = in_off[c], n = in_off[c + 1] - b;
    const uint8_t* p = src.data ? src.data + src.tile_off[tile] + src.cell_pos[c] : nullptr;
    const bool be = !src.strict_le;
    uint32_t* m = mm + b;
    uint8_t* a = ar + b;
    uint32_t ns = 0;
    bool failed = false;
    for (uint32_t i = 0; i < n; ++i) {
        int mn, mx;
        uint32_t area;
        if (p) {  // `u32 as i16`, `u32 as u8`: wrapping
            mn = (int16_t)rd32(p + 12ull * i, be);
            mx = (int16_t)rd32(p + 12ull * i + 4, be);
            area = rd32(p + 12ull * i + 8, be) & 0xffu;
        } else {
            mn = src.smin[b + i];
            mx = src.smax[b + i];
            area = src.sarea[b + i];
        }
        if (mn > mx) {  // heightfield.rs:110-115: Err — dropped by the lenient callers, fatal for the strict one
            failed = true;
            if (src.strict_le) break;
            continue;
        }
        uint32_t q = 0;
        for (;;) {
            if (q == ns) {  // empty column (:122-125) or end of list (:210-214)
                m[ns] = ((uint32_t)mn & 0xffffu) | ((uint32_t)mx << 16);
                a[ns] = (uint8_t)area;
                ++ns;
                break;
            }
            const uint32_t cw = m[q];
            const int cmin = (int16_t)(cw & 0xffffu), cmax = (int16_t)(cw >> 16);
            if (a[q] == area && cmax >= mn && mx >= cmin) {  // :160-192
                const int mmin = min(cmin, mn);
                int mmax = max(cmax, mx);
                if (q + 1 < ns) {
                    const uint32_t nw = m[q + 1];
                    if (a[q + 1] == area && mmax >= (int)(int16_t)(nw & 0xffffu)) {
                        mmax = max(mmax, (int)(int16_t)(nw >> 16));
                        for (uint32_t k = q + 1; k + 1 < ns; ++k) m[k] = m[k + 1], a[k] = a[k + 1];
                        --ns;
                    }
                }
                m[q] = ((uint32_t)mmin & 0xffffu) | ((uint32_t)mmax << 16);
                break;
            }
            if (mx < cmin) {  // :195-207 insert before
                for (uint32_t k = ns; k > q; --k) m[k] = m[k - 1], a[k] = a[k - 1];
                m[q] = ((uint32_t)mn & 0xffffu) | ((uint32_t)mx << 16);
                a[q] = (uint8_t)area;
                ++ns;
                break;
            }
            ++q;
        }
    }
    span_cnt[c] = ns;
    if (failed && src.strict_le) atomicMin(err, ((unsigned long long)tile << 33) | ((unsigned long long)cl << 1) | 1ull);
}

__device__ __forceinline__ void wr16(uint8_t* p, uint32_t v, bool be) {
    // streams of different tiles start at even offsets of an even base: halfword stores are aligned
    *(uint16_t*)p = (uint16_t)(be ? __byte_perm(v, 0, 0x0001) & 0xffffu : v & 0xffffu);
}
__device__ __forceinline__ void wr32(uint8_t* p, uint32_t v, bool be) {
    if (be) v = __byte_perm(v, 0, 0x0123);
    *(uint16_t*)p = (uint16_t)(v & 0xffffu);
    *(uint16_t*)(p + 2) = (uint16_t)(v >> 16);
}

__global__ void __launch_bounds__(256) k_wire_serialize(const TileDesc* __restrict__ tiles, TileMap tm,
                                                        const uint32_t* __restrict__ hf_cell_offset, const TileInfo* __restrict__ tinfo,
                                                        const int16_t* __restrict__ smin, const int16_t* __restrict__ smax,
                                                        const uint8_t* __restrict__ sarea, uint8_t* __restrict__ out, int big_endian) {
    const uint32_t tile = RCB_TILE_OF_BLOCK();
    if (tile >= tm.ntiles) return;
    const TileDesc& td = tiles[tile];
    const uint32_t cl = RCB_CELL_OF_THREAD();
    if (cl >= (uint32_t)(td.width * td.height)) return;
    const uint32_t* co = hf_cell_offset + td.cell_base + tile;  // width*height+1 tile-relative offsets
    const uint32_t sb = tinfo[tile].span_base;
    const uint32_t o = co[cl], n = co[cl + 1] - o;
    uint8_t* p = out + 2ull * (td.cell_base + cl) + 12ull * ((unsigned long long)sb + o);
    const bool be = big_endian != 0;
    wr16(p, n & 0xffffu, be);  // voxel_tile.rs:72-87: the count is a u16
    p += 2;
    for (uint32_t k = 0; k < n; ++k, p += 12) {
        const size_t s = (size_t)sb + o + k;
        wr32(p, (uint32_t)(int32_t)smin[s], be);  // `min as u32`: sign extension
        wr32(p + 4, (uint32_t)(int32_t)smax[s], be);
        wr32(p + 8, (uint32_t)sarea[s], be);
    }
}

}  // namespace

void launch_wire_walk(const TileDesc* tiles, int ntiles, const uint8_t* data, const unsigned long long* tile_off, int strict_le,
                      uint32_t* cell_pos, uint32_t* cell_n, unsigned long long* err, cudaStream_t s) {
    if (ntiles == 0) return;
    KScope _k("k_wire_walk", s);
    k_wire_walk<<<ntiles, 32, 0, s>>>(tiles, ntiles, data, tile_off, strict_le, cell_pos, cell_n, err);
}

void launch_wire_replay(const TileDesc* tiles, TileMap tm, const uint8_t* data, const unsigned long long* tile_off,
                        const uint32_t* cell_pos, int strict_le, const int16_t* smin, const int16_t* smax, const uint8_t* sarea,
                        const uint32_t* in_off, uint32_t* mm, uint8_t* ar, uint32_t* span_cnt, unsigned long long* err, cudaStream_t s) {
    if (tm.nblocks == 0) return;
    KScope _k("k_wire_replay", s);
    ReplaySrc src{data, tile_off, cell_pos, strict_le, smin, smax, sarea};
    k_wire_replay<<<tm.grid(), 256, 0, s>>>(tiles, tm, src, in_off, mm, ar, span_cnt, err);
}

void launch_wire_serialize(const TileDesc* tiles, TileMap tm, const uint32_t* hf_cell_offset, const TileInfo* tinfo,
                           const int16_t* smin, const int16_t* smax, const uint8_t* sarea, uint8_t* out, int big_endian,
                           cudaStream_t s) {
    if (tm.nblocks == 0) return;
    KScope _k("k_wire_serialize", s);
    k_wire_serialize<<<tm.grid(), 256, 0, s>>>(tiles, tm, hf_cell_offset, tinfo, smin, smax, sarea, out, big_endian);
}
