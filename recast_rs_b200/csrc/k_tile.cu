// k_tile.cu — tile-resident back end: one CTA owns one tile and keeps its working set in shared memory.
//
//   k_tile_spans   (A)  fragments of the tile (per-tile segment written by k_clip_items)
//                       -> bucket by column (shared-memory histogram + scan)
//                       -> per-column replay of add_span_internal in triangle order (rasterization.rs:51-168)
//                       -> the three walkable filters fused (heightfield.rs:286-531, order of lib.rs:278-287)
//                       -> heightfield CSR offsets (final) + the tile's spans in tile-local compact order
//   k_tile_scan         span base of every tile (exclusive scan over tiles) + batch totals
//   k_tile_compact (B)  spans of the tile -> CompactCell / CompactSpan arrays, 8-direction linking and the
//                       connection list (compact_heightfield.rs:175-430), erosion (area.rs:73-234) and the
//                       distance field with box blur (compact_heightfield.rs:514-727), all on shared memory;
//                       the tile's first connection index comes from a decoupled look-back over tiles.
//
// HBM traffic per tile: fragments read twice (second pass hits L2), spans written once and read once,
// every result array written once.  Semantics are identical to the v1 kernels (k_raster.cu, k_filter.cu,
// k_compact.cu, k_dist.cu), which stay as the fallback for tiles whose working set exceeds shared memory.
#include "rcb_common.cuh"

namespace {

constexpr int TILE_THREADS = 512;
constexpr int MAXH = 0xffff;
__constant__ int t_dirx[8] = {-1, 0, 1, 1, 1, 0, -1, -1};  // compact_heightfield.rs:109-111
__constant__ int t_dirz[8] = {-1, -1, -1, 0, 1, 1, 1, 0};

// Optional per-phase cycle accounting (rcb_profile_* with RCB_PHASE_TIMERS=1): thread 0 of every CTA adds the
// cycles since the previous mark to phase slot `id`.  phase == nullptr (the default) costs one predicate.
struct PhaseClock {
    unsigned long long* acc;
    long long t;
    __device__ PhaseClock(unsigned long long* a) : acc(a), t(0) {
        if (acc && threadIdx.x == 0) t = clock64();
    }
    __device__ void mark(int id) {
        if (acc && threadIdx.x == 0) {
            const long long n = clock64();
            atomicAdd(&acc[id], (unsigned long long)(n - t));
            t = n;
        }
    }
};

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t n = __shfl_up_sync(0xffffffffu, v, d);
        if ((threadIdx.x & 31) >= d) v += n;
    }
    return v;
}

// Exclusive scan of cnt[0..n) into out[0..n], out[n] = total (cnt and out may alias).  Rounds of
// TILE_THREADS consecutive elements (thread t takes element r*TILE_THREADS + t: conflict-free), each
// round one block-wide scan with a running carry.  Uses sh_warp[TILE_THREADS/32 + 1].
__device__ uint32_t block_scan_array(const uint32_t* cnt, uint32_t* out, uint32_t n, uint32_t* sh_warp) {
    uint32_t carry = 0;
    for (uint32_t base = 0; base < n; base += TILE_THREADS) {
        const uint32_t i = base + threadIdx.x;
        const uint32_t v = i < n ? cnt[i] : 0;
        const uint32_t inc = warp_incl_scan(v);
        if ((threadIdx.x & 31) == 31) sh_warp[threadIdx.x >> 5] = inc;
        __syncthreads();
        if (threadIdx.x < 32) {
            const uint32_t w = threadIdx.x < TILE_THREADS / 32 ? sh_warp[threadIdx.x] : 0;
            const uint32_t winc = warp_incl_scan(w);
            if (threadIdx.x < TILE_THREADS / 32) sh_warp[threadIdx.x] = winc - w;
            if (threadIdx.x == 31) sh_warp[TILE_THREADS / 32] = winc;
        }
        __syncthreads();
        if (i < n) out[i] = carry + inc - v + sh_warp[threadIdx.x >> 5];
        carry += sh_warp[TILE_THREADS / 32];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[n] = carry;
    __syncthreads();
    return carry;
}

// ======================================================================================================
// A: fragments -> spans
// ======================================================================================================
__global__ void __launch_bounds__(TILE_THREADS, 2) k_tile_spans(TileSpansArgs a) {
    extern __shared__ __align__(16) uint32_t smem[];
    __shared__ uint32_t sh_warp[TILE_THREADS / 32 + 1];
    __shared__ uint32_t sh_red[2];
    const uint32_t tile = blockIdx.x;
    const TileDesc& td = a.tiles[tile];
    const int W = td.width, H = td.height;
    const uint32_t C = (uint32_t)(W * H);
    const unsigned long long wmagic = td.wmagic;
    const uint32_t F = min(a.tile_cursor[tile], a.tile_cap[tile]);
    // Two launches share this kernel: pass 0 runs with a shared-memory size that lets two CTAs share an SM
    // and defers the tiles that need more; pass 1 (large allocation) handles only the deferred tiles.
    constexpr uint32_t DEFERRED = 0x80000000u;
    if (a.pass == 1 && !(a.tinfo[tile].status & DEFERRED)) return;
    // layout: off[C+1] | cnt[C] | soff[C+1] | rec[F] (uint2: area << 24 | triangle, smin | smax << 16)
    const size_t need = 4ull * (3ull * C + 4) + 8ull * F + 16;
    if (need > a.smem_bytes) {
        if (threadIdx.x == 0) {
            if (a.pass == 0 && a.defer) {
                a.tinfo[tile].status |= DEFERRED;
            } else {
                atomicOr(&a.counters[2], 4ull);  // does not fit: the host reruns the batch on the v1 back end
                a.tinfo[tile].span_count = 0;
                a.tinfo[tile].walkable_span_count = 0;
            }
        }
        return;
    }
    uint32_t* off = smem;
    uint32_t* cnt = off + (C + 1);
    uint32_t* soff = cnt + C;
    uint2* rec = reinterpret_cast<uint2*>(reinterpret_cast<uintptr_t>(soff + (C + 1) + 1) & ~(uintptr_t)7);
    const uint32_t* frags = a.tile_frags + 3ull * a.tile_base[tile];

    PhaseClock pc(a.phase);
    for (uint32_t c = threadIdx.x; c < C; c += TILE_THREADS) cnt[c] = 0;
    if (threadIdx.x < 2) sh_red[threadIdx.x] = 0;
    __syncthreads();
    // 1. histogram by column
    for (uint32_t i = threadIdx.x; i < F; i += TILE_THREADS) atomicAdd(&cnt[frags[3ull * i] & 0xffffffu], 1u);
    __syncthreads();
    pc.mark(0);
    // 2. column offsets
    block_scan_array(cnt, off, C, sh_warp);
    for (uint32_t c = threadIdx.x; c < C; c += TILE_THREADS) cnt[c] = 0;
    __syncthreads();
    pc.mark(1);
    // 3. scatter into column order (order inside a column is arbitrary; the replay sorts by triangle)
    for (uint32_t i = threadIdx.x; i < F; i += TILE_THREADS) {
        const uint32_t r0 = frags[3ull * i], r1 = frags[3ull * i + 1], r2 = frags[3ull * i + 2];
        const uint32_t cell = r0 & 0xffffffu;
        const uint32_t pos = off[cell] + atomicAdd(&cnt[cell], 1u);
        rec[pos] = make_uint2((r0 & 0xff000000u) | (r2 & 0xffffffu), r1);
    }
    __syncthreads();
    pc.mark(2);
    // 4. per-column replay (rasterization.rs:51-168), spans built in place over the consumed prefix
    const int thr = a.merge_thr;
    for (uint32_t c = threadIdx.x; c < C; c += TILE_THREADS) {
        const uint32_t n = cnt[c];
        if (n <= 1) continue;  // empty column, or a single fragment which is its own span
        uint2* r = rec + off[c];
        for (uint32_t i = 1; i < n; ++i) {  // insertion sort by triangle index
            const uint2 key = r[i];
            const uint32_t kt = key.x & 0xffffffu;
            uint32_t j = i;
            while (j > 0 && (r[j - 1].x & 0xffffffu) > kt) {
                r[j] = r[j - 1];
                --j;
            }
            if (j != i) r[j] = key;
        }
        uint32_t ns = 1;  // the first fragment opens the column (:65-74)
        for (uint32_t i = 1; i < n; ++i) {
            const uint2 fr = r[i];
            const uint32_t w = fr.y;
            const int smin = (int16_t)(w & 0xffffu), smax = (int16_t)(w >> 16);
            const uint32_t area = fr.x >> 24;
            uint32_t p = 0;
            bool done = false;
            while (p < ns) {
                const uint2 cs = r[p];
                const int cmin = (int16_t)(cs.y & 0xffffu), cmax = (int16_t)(cs.y >> 16);
                if (smax < cmin) break;
                if (smin > cmax) {
                    ++p;
                    continue;
                }
                const int mmin = min(smin, cmin);
                int mmax = max(smax, cmax);
                uint32_t marea = area;
                if (abs(mmax - cmax) <= thr) marea = max(marea, cs.x >> 24);
                uint32_t q = p + 1;
                while (q < ns) {
                    const uint32_t nw = r[q].y;
                    if ((int)(int16_t)(nw & 0xffffu) > mmax) break;
                    mmax = max(mmax, (int)(int16_t)(nw >> 16));
                    ++q;
                }
                r[p] = make_uint2(marea << 24, ((uint32_t)mmin & 0xffffu) | ((uint32_t)mmax << 16));
                if (q > p + 1) {
                    const uint32_t drop = q - (p + 1);
                    for (uint32_t k = q; k < ns; ++k) r[k - drop] = r[k];
                    ns -= drop;
                }
                done = true;
                break;
            }
            if (!done) {
                for (uint32_t k = ns; k > p; --k) r[k] = r[k - 1];
                r[p] = fr;
                ++ns;
            }
        }
        cnt[c] = ns;
    }
    __syncthreads();
    pc.mark(3);
    // 5. span offsets (tile-local compact order)
    const uint32_t S = block_scan_array(cnt, soff, C, sh_warp);
    pc.mark(4);

    // 6. filters, fused (see k_filter.cu for why one bottom-up walk per column is exact)
    uint32_t walkable = 0, maxh = 0;
    const uint32_t fmask = a.fmask;
    const int height = td.walkable_height, climb = td.walkable_climb, neg_climb = td.neg_climb;
    for (uint32_t c = threadIdx.x; c < C; c += TILE_THREADS) {
        const uint32_t b = off[c], e = b + cnt[c];
        if (b == e) continue;
        const int z = (int)rcb_div_w(c, wmagic), x = (int)c - z * W;
        const int dx[4] = {0, 1, 0, -1}, dz[4] = {-1, 0, 1, 0};
        uint32_t nb[4], ne[4];
        bool oob[4];
#pragma unroll
        for (int d = 0; d < 4; ++d) {
            const int nx = x + dx[d], nz = z + dz[d];
            oob[d] = nx < 0 || nz < 0 || nx >= W || nz >= H;
            const uint32_t nc = oob[d] ? c : (uint32_t)(nz * W + nx);
            nb[d] = off[nc];
            ne[d] = oob[d] ? nb[d] : nb[d] + cnt[nc];
        }
        bool prev_walkable = false, have_prev = false;
        uint32_t prev_area_f1 = 0;
        int prev_max = 0;
        for (uint32_t s = b; s < e; ++s) {
            const uint2 sp = rec[s];
            const uint32_t w = sp.y;
            const int floor = (int16_t)(w >> 16);
            const int ceiling = (s + 1 < e) ? (int)(int16_t)(rec[s + 1].y & 0xffffu) : MAXH;
            const uint32_t orig = sp.x >> 24;
            const bool wk = orig != 0;
            uint32_t ar = orig;
            if ((fmask & 1u) && !wk && prev_walkable && have_prev && (floor - prev_max) <= climb) ar = prev_area_f1;
            prev_walkable = wk;
            prev_area_f1 = ar;
            prev_max = floor;
            have_prev = true;
            if ((fmask & 2u) && ar != 0) {
                int lowest = MAXH, lo = floor, hi = floor;
                bool ledge = false;
#pragma unroll
                for (int d = 0; d < 4 && !ledge; ++d) {
                    if (oob[d] || nb[d] == ne[d]) {
                        ledge = true;
                        break;
                    }
                    int nceil = (int16_t)(rec[nb[d]].y & 0xffffu);
                    if (min(ceiling, nceil) - floor >= height) {
                        ledge = true;
                        break;
                    }
                    for (uint32_t ns = nb[d]; ns < ne[d]; ++ns) {
                        const int nfloor = (int16_t)(rec[ns].y >> 16);
                        nceil = (ns + 1 < ne[d]) ? (int)(int16_t)(rec[ns + 1].y & 0xffffu) : MAXH;
                        if (min(ceiling, nceil) - max(floor, nfloor) < height) continue;
                        const int diff = nfloor - floor;
                        lowest = min(lowest, diff);
                        if (abs(diff) <= climb) {
                            lo = min(lo, nfloor);
                            hi = max(hi, nfloor);
                        } else if (diff < neg_climb) {
                            break;
                        }
                    }
                }
                if (ledge || lowest < neg_climb || (hi - lo) > climb) ar = 0;
            }
            if ((fmask & 4u) && ceiling - floor < height) ar = 0;
            if (ar != orig) rec[s].x = ar << 24;  // neighbours read .y only
            walkable += ar != 0;
            maxh = max(maxh, (uint32_t)(uint16_t)(w >> 16));
        }
    }
    __syncthreads();
    pc.mark(5);
    // 7. outputs: final heightfield CSR offsets, spans in tile-local compact order
    uint32_t* out_off = a.hf_cell_offset + td.cell_base + tile;
    const uint32_t tb = a.tile_base[tile];
    for (uint32_t c = threadIdx.x; c < C; c += TILE_THREADS) {
        const uint32_t so = soff[c], n = cnt[c], b = off[c];
        out_off[c] = so;
        for (uint32_t k = 0; k < n; ++k) {
            const uint2 sp = rec[b + k];
            a.ts_mm[tb + so + k] = sp.y;
            a.ts_area[tb + so + k] = (uint8_t)(sp.x >> 24);
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        walkable += __shfl_down_sync(0xffffffffu, walkable, d);
        maxh = max(maxh, __shfl_down_sync(0xffffffffu, maxh, d));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&sh_red[0], walkable);
        atomicMax(&sh_red[1], maxh);
    }
    __syncthreads();
    pc.mark(6);
    if (threadIdx.x == 0) {
        out_off[C] = S;
        if (F) atomicAdd(&a.counters[1], (unsigned long long)F);
        a.tinfo[tile].span_count = S;
        a.tinfo[tile].walkable_span_count = sh_red[0];
        a.tinfo[tile].max_height = sh_red[1];
        if (a.pass == 1) a.tinfo[tile].status &= ~DEFERRED;
    }
}

// span_base per tile + batch totals (counters[6] = S, counters[7] = walkable spans); one block
__global__ void __launch_bounds__(1024) k_tile_scan(TileInfo* __restrict__ tinfo, int ntiles, unsigned long long* counters) {
    __shared__ uint32_t sh[33];
    __shared__ unsigned long long sh_walk;
    if (threadIdx.x == 0) sh_walk = 0;
    __syncthreads();
    __shared__ uint32_t sh_maxs;
    if (threadIdx.x == 0) sh_maxs = 0;
    uint32_t carry = 0, maxs = 0;
    unsigned long long walk = 0;
    for (int base = 0; base < ntiles; base += 1024) {
        const int t = base + threadIdx.x;
        const uint32_t v = t < ntiles ? tinfo[t].span_count : 0;
        if (t < ntiles) walk += tinfo[t].walkable_span_count;
        maxs = max(maxs, v);
        const uint32_t inc = warp_incl_scan(v);
        if ((threadIdx.x & 31) == 31) sh[threadIdx.x >> 5] = inc;
        __syncthreads();
        if (threadIdx.x < 32) {
            const uint32_t w = sh[threadIdx.x];
            const uint32_t winc = warp_incl_scan(w);
            sh[threadIdx.x] = winc - w;
            if (threadIdx.x == 31) sh[32] = winc;
        }
        __syncthreads();
        if (t < ntiles) tinfo[t].span_base = carry + inc - v + sh[threadIdx.x >> 5];
        carry += sh[32];
        __syncthreads();
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {  // one shared-memory atomic per warp, not per thread
        walk += __shfl_down_sync(0xffffffffu, walk, d);
        maxs = max(maxs, __shfl_down_sync(0xffffffffu, maxs, d));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&sh_walk, walk);
        atomicMax(&sh_maxs, maxs);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        counters[6] = carry;
        counters[7] = sh_walk;
        counters[8] = sh_maxs;
    }
}

// tile_base = exclusive scan of the per-tile fragment bounds; counters[5] = largest bound; one block
__global__ void __launch_bounds__(1024) k_tile_segments(const uint32_t* __restrict__ tile_ub, uint32_t* __restrict__ tile_base,
                                                        int ntiles, unsigned long long* counters) {
    __shared__ uint32_t sh[33];
    __shared__ uint32_t sh_max;
    if (threadIdx.x == 0) sh_max = 0;
    __syncthreads();
    uint32_t carry = 0, mx = 0;
    for (int base = 0; base < ntiles; base += 1024) {
        const int t = base + threadIdx.x;
        const uint32_t v = t < ntiles ? tile_ub[t] : 0;
        mx = max(mx, v);
        const uint32_t inc = warp_incl_scan(v);
        if ((threadIdx.x & 31) == 31) sh[threadIdx.x >> 5] = inc;
        __syncthreads();
        if (threadIdx.x < 32) {
            const uint32_t w = sh[threadIdx.x];
            const uint32_t winc = warp_incl_scan(w);
            sh[threadIdx.x] = winc - w;
            if (threadIdx.x == 31) sh[32] = winc;
        }
        __syncthreads();
        if (t < ntiles) tile_base[t] = carry + inc - v + sh[threadIdx.x >> 5];
        carry += sh[32];
        __syncthreads();
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) mx = max(mx, __shfl_down_sync(0xffffffffu, mx, d));
    if ((threadIdx.x & 31) == 0) atomicMax(&sh_max, mx);
    __syncthreads();
    if (threadIdx.x == 0) {
        tile_base[ntiles] = carry;
        counters[5] = sh_max;
    }
}

// ======================================================================================================
// B: spans -> compact heightfield, connections, erosion, distance field
// ======================================================================================================
struct TileB {
    const uint32_t* soff;  // [C+1]
    const uint32_t* mm;    // [S] smin | smax << 16
    const uint8_t* area;   // [S] span area (post-filter)
    uint8_t* areas;        // [S] chf.areas (erosion writes it)
    const uint8_t* nei;    // [8][S] neighbour offset inside the neighbour cell, 0xff = none
    uint32_t S, S4;  // S4: S rounded up to 4 (row pitch of nei)
    int W, H;
    unsigned long long wmagic;
};

__device__ __forceinline__ uint32_t tnei(const TileB& g, uint32_t s, int x, int z, int dir) {
    const uint32_t o = g.nei[(uint32_t)dir * g.S4 + s];
    if (o == 0xffu) return RCB_NONE;
    return g.soff[(z + t_dirz[dir]) * g.W + (x + t_dirx[dir])] + o;
}

template <typename T>
__device__ __forceinline__ T sat_add(T v, unsigned b) {
    const unsigned r = (unsigned)v + b;
    const unsigned mx = (T)~(T)0;
    return (T)(r > mx ? mx : r);
}

// Forward-sweep gather table: for every span the tile-local ids of NW, NW.E, E, E.NE (0xffff = none).
// The chain nei -> cell offset -> nei -> cell offset does not depend on distances, so it is hoisted out
// of the H sequential row steps (it was 27 % of k_tile_compact's stall samples, profiles/r01_v3_*).
__device__ void build_forward_index(const TileB& g, uint2* fidx) {
    const int W = g.W;
    const uint32_t C = (uint32_t)(W * g.H);
    for (uint32_t c = threadIdx.x; c < C; c += TILE_THREADS) {
        const int z = (int)rcb_div_w(c, g.wmagic), x = (int)c - z * W;
        for (uint32_t s = g.soff[c]; s < g.soff[c + 1]; ++s) {
            uint32_t a0 = tnei(g, s, x, z, 0), a1 = RCB_NONE, b0 = tnei(g, s, x, z, 3), b1 = RCB_NONE;
            if (a0 != RCB_NONE) a1 = tnei(g, a0, x - 1, z - 1, 3);
            if (b0 != RCB_NONE) b1 = tnei(g, b0, x + 1, z, 2);
            fidx[s] = make_uint2((a0 & 0xffffu) | (a1 << 16), (b0 & 0xffffu) | (b1 << 16));
        }
    }
}

// boundary -> forward rows -> backward gather, on shared memory.  ERODE: thresholds areas; else D + max.
// fidx: optional forward gather table (nullptr: walk the neighbour chain inside the row loop).
template <typename T, bool ERODE>
__device__ void tile_sweep(const TileB& g, T* init, T* Fw, const uint2* fidx, unsigned thr, uint32_t* sh_max) {
    const int W = g.W, H = g.H;
    const uint32_t C = (uint32_t)(W * H);
    const unsigned long long wmagic = g.wmagic;
    const T inf = (T)~(T)0;
    for (uint32_t c = threadIdx.x; c < C; c += TILE_THREADS) {
        const int z = (int)rcb_div_w(c, wmagic), x = (int)c - z * W;
        for (uint32_t s = g.soff[c]; s < g.soff[c + 1]; ++s) {
            const uint8_t ar = g.areas[s];
            int cnt = 0;
            if (!ERODE || ar != 0) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint32_t n = tnei(g, s, x, z, 7 - 2 * k);
                    if (n == RCB_NONE) continue;
                    if (ERODE ? (g.areas[n] != 0) : (g.areas[n] == ar)) cnt++;
                }
            }
            init[s] = (cnt == 4) ? inf : (T)0;
        }
    }
    __syncthreads();
    // forward: row z needs row z-1 final and row z init (see k_dist.cu).  Only the warps that own a
    // column take part in the H row barriers (named barrier 1); the rest wait at the __syncthreads below.
    const int sweep_threads = min(TILE_THREADS, ((W + 31) / 32) * 32);
    if ((int)threadIdx.x < sweep_threads) {
        for (int z = 0; z < H; ++z) {
            for (int x = threadIdx.x; x < W; x += sweep_threads) {
                const uint32_t c = (uint32_t)(z * W + x);
                const uint32_t sb = g.soff[c], se = g.soff[c + 1];
                for (uint32_t s = sb; s < se; ++s) {
                    T v = init[s];
                    if (fidx) {
                        const uint2 ix = fidx[s];
                        const uint32_t a0 = ix.x & 0xffffu, a1 = ix.x >> 16, b0 = ix.y & 0xffffu, b1 = ix.y >> 16;
                        if (a0 != 0xffffu) v = min(v, sat_add<T>(Fw[a0], 2));
                        if (a1 != 0xffffu) v = min(v, sat_add<T>(Fw[a1], 3));
                        if (b0 != 0xffffu) v = min(v, sat_add<T>(init[b0], 2));
                        if (b1 != 0xffffu) v = min(v, sat_add<T>(Fw[b1], 3));
                    } else {
                        const uint32_t aa = tnei(g, s, x, z, 0);
                        if (aa != RCB_NONE) {
                            v = min(v, sat_add<T>(Fw[aa], 2));
                            const uint32_t ab = tnei(g, aa, x - 1, z - 1, 3);
                            if (ab != RCB_NONE) v = min(v, sat_add<T>(Fw[ab], 3));
                        }
                        const uint32_t ba = tnei(g, s, x, z, 3);
                        if (ba != RCB_NONE) {
                            v = min(v, sat_add<T>(init[ba], 2));
                            const uint32_t bb = tnei(g, ba, x + 1, z, 2);
                            if (bb != RCB_NONE) v = min(v, sat_add<T>(Fw[bb], 3));
                        }
                    }
                    Fw[s] = v;
                }
            }
            asm volatile("bar.sync 1, %0;" ::"r"(sweep_threads) : "memory");
        }
    }
    __syncthreads();
    // backward = gather from forward values; D overwrites init (init is dead)
    uint32_t vmax = 0;
    for (uint32_t c = threadIdx.x; c < C; c += TILE_THREADS) {
        const int z = (int)rcb_div_w(c, wmagic), x = (int)c - z * W;
        for (uint32_t s = g.soff[c]; s < g.soff[c + 1]; ++s) {
            T v = Fw[s];
            const uint32_t aa = tnei(g, s, x, z, 2);
            if (aa != RCB_NONE) {
                v = min(v, sat_add<T>(Fw[aa], 2));
                const uint32_t ab = tnei(g, aa, x + 1, z - 1, 1);
                if (ab != RCB_NONE) v = min(v, sat_add<T>(Fw[ab], 3));
            }
            const uint32_t ba = tnei(g, s, x, z, 1);
            if (ba != RCB_NONE) {
                v = min(v, sat_add<T>(Fw[ba], 2));
                const uint32_t bb = tnei(g, ba, x, z - 1, 0);
                if (bb != RCB_NONE) v = min(v, sat_add<T>(Fw[bb], 3));
            }
            init[s] = v;
            if (!ERODE) vmax = max(vmax, (uint32_t)v);
        }
    }
    if (!ERODE) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) vmax = max(vmax, __shfl_down_sync(0xffffffffu, vmax, d));
        if ((threadIdx.x & 31) == 0 && vmax) atomicMax(sh_max, vmax);
    }
    __syncthreads();
    if (ERODE) {
        for (uint32_t s = threadIdx.x; s < g.S; s += TILE_THREADS)
            if ((unsigned)init[s] < thr) g.areas[s] = 0;  // area.rs:226-231
        __syncthreads();
    }
}

// Exclusive prefix of `mine` over tiles by decoupled look-back (warp 0 calls this with all 32 lanes).
// look[t] = flag << 62 | value; flag 1 = tile aggregate, 2 = inclusive prefix.  Tiles obtain their ids
// from a ticket counter, so every predecessor is already running or done.
__device__ uint32_t lookback_exclusive(volatile unsigned long long* look, uint32_t tile, uint32_t mine) {
    const uint32_t lane = threadIdx.x & 31;
    if (lane == 0) look[tile] = ((tile == 0 ? 2ull : 1ull) << 62) | mine;
    uint32_t excl = 0;
    for (int t0 = (int)tile - 1; t0 >= 0; t0 -= 32) {
        const int t = t0 - (int)lane;
        unsigned long long v = 2ull << 62;  // before tile 0: inclusive prefix 0
        if (t >= 0) {
            do {
                v = look[t];
            } while ((v >> 62) == 0);
        }
        const unsigned incl = __ballot_sync(0xffffffffu, (v >> 62) == 2);
        const uint32_t val = (uint32_t)(v & 0xffffffffull);
        if (incl) {
            const int first = __ffs(incl) - 1;  // nearest predecessor with an inclusive prefix
            excl += __reduce_add_sync(0xffffffffu, (int)lane <= first ? val : 0u);
            break;
        }
        excl += __reduce_add_sync(0xffffffffu, val);
    }
    if (lane == 0 && tile != 0) look[tile] = (2ull << 62) | (unsigned long long)(excl + mine);
    return excl;
}

__global__ void __launch_bounds__(TILE_THREADS, 2) k_tile_compact(TileCompactArgs a) {
    extern __shared__ __align__(16) uint32_t smem[];
    __shared__ uint32_t sh_warp[TILE_THREADS / 32 + 1];
    __shared__ uint32_t sh_misc[4];  // [0] tile id, [1] conn base, [2] max distance, [3] overflow
    // dynamic tile id: tiles start in order, which the look-back below relies on
    if (threadIdx.x == 0) {
        sh_misc[0] = atomicAdd(a.tile_ticket, 1u);
        sh_misc[2] = 0;
        sh_misc[3] = 0;
    }
    __syncthreads();
    const uint32_t tile = sh_misc[0];
    const TileDesc& td = a.tiles[tile];
    const int W = td.width, H = td.height;
    const uint32_t C = (uint32_t)(W * H);
    const unsigned long long wmagic = td.wmagic;
    const uint32_t S = a.tinfo[tile].span_count;
    const uint32_t span_base = a.tinfo[tile].span_base;
    // layout: soff[C+1] | d16a[S] u16 | d16b[S] u16 | areas[S] | nei[8S] | coff[C+1] | mm[S] | area[S]
    // (coff, mm and area are dead once the connections are written; the forward gather table reuses them)
    const size_t need = 4ull * (2ull * C + 2) + 4ull * S + 4ull * S + 2ull * S + 8ull * S + 64;
    const bool fits = a.do_compact ? need <= a.smem_bytes : 4ull * (C + 1) <= a.smem_bytes;
    const uint32_t S2 = (S + 1) & ~1u, S4 = (S + 3) & ~3u;
    uint32_t* soff = smem;
    uint16_t* d16a = reinterpret_cast<uint16_t*>(soff + (C + 1));
    uint16_t* d16b = d16a + S2;
    uint8_t* areas = reinterpret_cast<uint8_t*>(d16b + S2);
    uint8_t* nei = areas + S4;
    uint32_t* coff = reinterpret_cast<uint32_t*>(nei + 8ull * S4);
    uint32_t* mm = coff + (C + 1);
    uint8_t* area = reinterpret_cast<uint8_t*>(mm + S);

    volatile unsigned long long* look = a.lookback;
    if (!fits) {
        // still take part in the look-back chain so later tiles do not wait forever
        if (threadIdx.x < 32) {
            const uint32_t excl = lookback_exclusive(look, tile, 0);
            if (threadIdx.x == 0) {
                atomicOr(&a.counters[2], 4ull);
                a.tinfo[tile].conn_base = excl;
                a.tinfo[tile].conn_count = 0;
            }
        }
        return;
    }
    PhaseClock pc(a.phase ? a.phase + 8 : nullptr);
    // 1. load the tile's spans
    const uint32_t* g_off = a.hf_cell_offset_in + td.cell_base + tile;
    uint32_t* g_off_out = a.hf_cell_offset + td.cell_base + tile;
    for (uint32_t c = threadIdx.x; c <= C; c += TILE_THREADS) {
        const uint32_t o = g_off[c];
        soff[c] = o;
        g_off_out[c] = o;
    }
    const uint32_t tb = a.tile_base[tile];
    for (uint32_t s = threadIdx.x; s < S; s += TILE_THREADS) {
        const uint32_t w = a.ts_mm[tb + s];
        const uint8_t ar = a.ts_area[tb + s];
        if (a.do_compact) {
            mm[s] = w;
            area[s] = ar;
            areas[s] = ar;
        }
        // final span attributes (compact_heightfield.rs:231-237)
        a.smin[span_base + s] = (int16_t)(w & 0xffffu);
        a.smax[span_base + s] = (int16_t)(w >> 16);
        a.sarea[span_base + s] = ar;
    }
    if (!a.do_compact) return;  // heightfield-only result
    __syncthreads();
    pc.mark(0);
    // 2. cells (compact_heightfield.rs:203-223)
    for (uint32_t c = threadIdx.x; c < C; c += TILE_THREADS) {
        const uint32_t b = soff[c], n = soff[c + 1] - b;
        a.cell_index[td.cell_base + c] = n ? b : RCB_NONE;
        a.cell_count[td.cell_base + c] = n;
        if (n > 254) sh_misc[3] = 1;  // neighbour offsets are stored in 8 bits here
    }
    // 3. link (compact_heightfield.rs:297-372) + con[4] (:392-427)
    for (uint32_t c = threadIdx.x; c < C; c += TILE_THREADS) {
        const uint32_t b = soff[c], e = soff[c + 1];
        uint32_t nconn = 0;
        if (e > b) {
            const int z = (int)rcb_div_w(c, wmagic), x = (int)c - z * W;
            uint32_t nb[8], ne[8];
#pragma unroll
            for (int d = 0; d < 8; ++d) {
                const int nx = x + t_dirx[d], nz = z + t_dirz[d];
                if (nx < 0 || nz < 0 || nx >= W || nz >= H) {
                    nb[d] = ne[d] = 0;
                } else {
                    nb[d] = soff[nz * W + nx];
                    ne[d] = soff[nz * W + nx + 1];
                }
            }
            for (uint32_t s = b; s < e; ++s) {
                const uint32_t w = mm[s];
                const int mn = (int16_t)(w & 0xffffu), mx = (int16_t)(w >> 16);
                uint32_t c4[4] = {63, 63, 63, 63};
                if (area[s] != 0) {
#pragma unroll
                    for (int d = 0; d < 8; ++d) {
                        uint32_t best = 0xffu;
                        int best_diff = 0x7fffffff;
                        for (uint32_t ns = nb[d]; ns < ne[d]; ++ns) {
                            if (area[ns] == 0) continue;
                            const uint32_t nw = mm[ns];
                            const int nmn = (int16_t)(nw & 0xffffu), nmx = (int16_t)(nw >> 16);
                            if (max(nmn, mn) > min(nmx, mx)) continue;
                            const int diff = (int)(uint16_t)(abs(nmn - mn) + abs(nmx - mx));
                            if (diff < best_diff) {
                                best_diff = diff;
                                best = ns - nb[d];
                            }
                        }
                        nei[(uint32_t)d * S4 + s] = (uint8_t)best;
                        if (best != 0xffu) {
                            nconn++;
                            if (d == 7) c4[0] = best;
                            if (d == 5) c4[1] = best;
                            if (d == 3) c4[2] = best;
                            if (d == 1) c4[3] = best;
                        }
                    }
                } else {
#pragma unroll
                    for (int d = 0; d < 8; ++d) nei[(uint32_t)d * S4 + s] = 0xffu;
                }
                *reinterpret_cast<uint2*>(a.con + 4ull * (span_base + s)) = make_uint2(c4[0] | (c4[1] << 16), c4[2] | (c4[3] << 16));
            }
        }
        coff[c] = nconn;
    }
    __syncthreads();
    pc.mark(1);
    // 4. connection offsets inside the tile, then the tile's base by decoupled look-back over tiles
    const uint32_t K = block_scan_array(coff, coff, C, sh_warp);
    if (threadIdx.x < 32) {
        const uint32_t excl = lookback_exclusive(look, tile, K);
        if (threadIdx.x == 0) {
            if (sh_misc[3]) atomicOr(&a.counters[2], 4ull);
            sh_misc[1] = excl;
            a.tinfo[tile].conn_base = excl;
            a.tinfo[tile].conn_count = K;
        }
    }
    __syncthreads();
    const uint32_t conn_base = sh_misc[1];
    pc.mark(2);
    // 5. connection list (compact_heightfield.rs:375-389): reverse direction order, next -> previous
    for (uint32_t c = threadIdx.x; c < C; c += TILE_THREADS) {
        const uint32_t b = soff[c], e = soff[c + 1];
        if (b == e) continue;
        const int z = (int)rcb_div_w(c, wmagic), x = (int)c - z * W;
        uint32_t k = coff[c];
        for (uint32_t s = b; s < e; ++s) {
            uint32_t first = RCB_NONE;
            for (int d = 7; d >= 0; --d) {
                const uint32_t o = nei[(uint32_t)d * S4 + s];
                if (o == 0xffu) continue;
                a.conn_ref[conn_base + k] = soff[(z + t_dirz[d]) * W + (x + t_dirx[d])] + o;
                a.conn_dir[conn_base + k] = (uint8_t)d;
                a.conn_next[conn_base + k] = first;
                first = k;
                ++k;
            }
            a.first_conn[span_base + s] = first;
        }
    }
    pc.mark(3);
    const TileB g{soff, mm, area, areas, nei, S, S4, W, H, wmagic};
    // coff and mm are dead once the connections are written: reuse them for the forward gather table
    uint2* fidx = nullptr;
    if ((a.do_erode || a.do_dist) && S < 0xffffu && 8ull * S + 8 <= 4ull * (C + 1) + 5ull * S) {
        __syncthreads();
        fidx = reinterpret_cast<uint2*>(reinterpret_cast<uintptr_t>(coff + 1) & ~(uintptr_t)7);
        build_forward_index(g, fidx);
        __syncthreads();
    }
    pc.mark(4);
    // 6. erosion (area.rs:73-234) — before the distance field when both are requested
    if (a.do_erode) {
        const unsigned thr = (unsigned)(uint8_t)(uint32_t)(a.erode_radius * 2);
        tile_sweep<uint8_t, true>(g, reinterpret_cast<uint8_t*>(d16a), reinterpret_cast<uint8_t*>(d16b), fidx, thr, nullptr);
    }
    for (uint32_t s = threadIdx.x; s < S; s += TILE_THREADS) a.areas[span_base + s] = areas[s];
    pc.mark(5);
    // 7. distance field (compact_heightfield.rs:514-727)
    if (a.do_dist) {
        tile_sweep<uint16_t, false>(g, d16a, d16b, fidx, 0, &sh_misc[2]);
        pc.mark(6);
        // box blur: d16a holds D; result to global
        for (uint32_t c = threadIdx.x; c < C; c += TILE_THREADS) {
            const int z = (int)rcb_div_w(c, wmagic), x = (int)c - z * W;
            for (uint32_t s = soff[c]; s < soff[c + 1]; ++s) {
                const uint32_t cd = d16a[s];
                uint32_t r = cd;
                if (cd > 2) {
                    int d = (int)cd, n = 1;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int dir = 1 + 2 * k;
                        const uint32_t nbr = tnei(g, s, x, z, dir);
                        if (nbr == RCB_NONE) continue;
                        d += d16a[nbr];
                        n++;
                        const uint32_t dg = tnei(g, nbr, x + t_dirx[dir], z + t_dirz[dir], (dir + 1) & 7);
                        if (dg != RCB_NONE) {
                            d += d16a[dg];
                            n++;
                        }
                    }
                    r = (uint32_t)(d / n);
                }
                a.dist[span_base + s] = (uint16_t)r;
            }
        }
        __syncthreads();
        pc.mark(7);
        if (threadIdx.x == 0) a.tinfo[tile].max_distance = sh_misc[2];
    } else {
        for (uint32_t s = threadIdx.x; s < S; s += TILE_THREADS) a.dist[span_base + s] = 0;  // compact_heightfield.rs:253
    }
}

}  // namespace

static void set_smem_attr(const void* fn, size_t bytes) {
    cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

void launch_tile_spans(const TileSpansArgs& a, int ntiles, cudaStream_t s) {
    if (ntiles == 0) return;
    set_smem_attr((const void*)k_tile_spans, a.smem_bytes);
    KScope _k("k_tile_spans", s);
    k_tile_spans<<<ntiles, TILE_THREADS, a.smem_bytes, s>>>(a);
}

void launch_tile_segments(const uint32_t* tile_ub, uint32_t* tile_base, int ntiles, unsigned long long* counters, cudaStream_t s) {
    KScope _k("k_tile_segments", s);
    k_tile_segments<<<1, 1024, 0, s>>>(tile_ub, tile_base, ntiles, counters);
}

void launch_tile_scan(TileInfo* tinfo, int ntiles, unsigned long long* counters, cudaStream_t s) {
    KScope _k("k_tile_scan", s);
    k_tile_scan<<<1, 1024, 0, s>>>(tinfo, ntiles, counters);
}

void launch_tile_compact(const TileCompactArgs& a, int ntiles, cudaStream_t s) {
    if (ntiles == 0) return;
    set_smem_attr((const void*)k_tile_compact, a.smem_bytes);
    KScope _k("k_tile_compact", s);
    k_tile_compact<<<ntiles, TILE_THREADS, a.smem_bytes, s>>>(a);
}
