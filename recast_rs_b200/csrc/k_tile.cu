// k_tile.cu — tile-resident back end: one CTA owns one BAND of a tile (all of it when the tile is small enough, else a run
// of rows, see BandDesc) and keeps its working set in shared memory.
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
#define RCB_FILE_ID 1
#include "rcb_bounds.cuh"

namespace {

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

// Exclusive scan of cnt[0..n) into out[0..n], out[n] = total (cnt and out may alias when they have the same type).
// Every warp owns one contiguous chunk: it first adds its chunk up, the chunk totals are scanned by warp 0, and a second
// pass over the chunk writes the offsets - two block barriers in all.  (Rounds of one element per thread with a block-wide
// scan each cost 2 barriers per TILE_THREADS elements: 8 k cycles for a 64x64 tile, three times per tile.)
// Uses sh_warp[TILE_THREADS/32 + 1].
template <int TILE_THREADS, typename OutT>
__device__ uint32_t block_scan_array(RCB_A(uint32_t) cnt, RCB_A(OutT) out, uint32_t n, uint32_t* sh_warp) {
    constexpr uint32_t NW = TILE_THREADS / 32;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t chunk = ((n + NW - 1) / NW + 31u) & ~31u;  // elements per warp, a multiple of 32
    const uint32_t c0 = min(warp * chunk, n), c1 = min(c0 + chunk, n);
    uint32_t sum = 0;
    for (uint32_t i = c0 + lane; i < c1; i += 32) sum += cnt[i];
    sum = __reduce_add_sync(0xffffffffu, sum);
    if (lane == 0) sh_warp[warp] = sum;
    __syncthreads();
    if (warp == 0) {
        const uint32_t w = lane < NW ? sh_warp[lane] : 0;
        const uint32_t winc = warp_incl_scan(w);
        if (lane < NW) sh_warp[lane] = winc - w;
        if (lane == NW - 1) sh_warp[NW] = winc;
    }
    __syncthreads();
    uint32_t carry = sh_warp[warp];
    const uint32_t total = sh_warp[NW];
    for (uint32_t i0 = c0; i0 < c1; i0 += 32) {
        const uint32_t i = i0 + lane;
        const uint32_t v = i < c1 ? cnt[i] : 0;
        const uint32_t inc = warp_incl_scan(v);
        if (i < c1) out[i] = (OutT)(carry + inc - v);
        carry += __shfl_sync(0xffffffffu, inc, 31);
    }
    if (threadIdx.x == 0) out[n] = (OutT)total;
    __syncthreads();
    return total;
}

// ======================================================================================================
// A: fragments -> spans
// ======================================================================================================
// TILE_THREADS = 512 (two CTAs per SM) for whole tiles; 256 (four per SM) for tiles cut into bands: a band waits for its
// predecessor's forward sweep, and more resident CTAs keep the SM busy meanwhile.
template <int TILE_THREADS>
__global__ void __launch_bounds__(TILE_THREADS, 1024 / TILE_THREADS) k_tile_spans(TileSpansArgs a) {
    extern __shared__ __align__(16) uint32_t smem[];
    __shared__ uint32_t sh_warp[TILE_THREADS / 32 + 1];
    __shared__ uint32_t sh_red[4];  // [0] walkable spans, [1] max height, [2] own fragments, [3] replay work-list length
    const uint32_t band = blockIdx.x;
    if (a.counters[2] & RCB_FLAG_ABORT) return;  // a speculative size was too small: the host redoes the batch
    const BandDesc bd = a.bands[band];
    const TileDesc& td = a.tiles[bd.tile];
    const int W = td.width, H = td.height;
    const uint32_t C = (uint32_t)(W * (bd.hz1 - bd.hz0));                                    // cells held: own rows + halo rows
    const uint32_t oc0 = (uint32_t)(W * (bd.z0 - bd.hz0)), oc1 = (uint32_t)(W * (bd.z1 - bd.hz0));  // own cells [oc0, oc1)
    const unsigned long long wmagic = td.wmagic;
    const uint32_t F = min(a.tile_cursor[band], a.tile_cap[band]);
    // Two launches share this kernel: pass 0 runs with a shared-memory size that lets two CTAs share an SM
    // and defers the bands that need more; pass 1 (large allocation) handles only the deferred bands.
    constexpr uint32_t DEFERRED = 0x80000000u;
    if (a.pass == 1 && !(a.binfo[band].status & DEFERRED)) return;
    // layout: cnt u32[C] | off u16[C+1] | soff u16[C+1] | rec[F] (uint2: area << 24 | triangle, smin | smax << 16)
    const size_t need = 8ull * (C + 4) + 8ull * F + 16;
    if (need > a.smem_bytes || F >= 0xffffu) {
        if (threadIdx.x == 0) {
            if (a.pass == 0 && a.defer) {
                a.binfo[band].status |= DEFERRED;
            } else {
                atomicOr(&a.counters[2], 4ull);  // does not fit: the host reruns the batch with smaller bands / the v1 back end
                a.binfo[band].span_count = 0;
                a.binfo[band].walkable = 0;
                a.binfo[band].top_spans = a.binfo[band].bot_spans = 0;
            }
        }
        return;
    }
    RCB_A(uint32_t) cnt = RCB_MK(uint32_t, smem, C);
    RCB_A(uint16_t) off = RCB_MK(uint16_t, smem + C, C + 1);
    RCB_A(uint16_t) soff = RCB_MK(uint16_t, reinterpret_cast<uint16_t*>(smem + C) + ((C + 2) & ~1u), C + 1);
    RCB_A(uint2) rec = RCB_MK(uint2, (reinterpret_cast<uintptr_t>(reinterpret_cast<uint16_t*>(smem + C) + 2 * ((C + 2) & ~1u)) + 7) & ~(uintptr_t)7, F);
    RCB_A(const uint32_t) frags = RCB_MK(const uint32_t, a.tile_frags + 3ull * a.tile_base[band], 3ull * a.tile_cap[band]);

    PhaseClock pc(a.phase);
    for (uint32_t c = threadIdx.x; c < C; c += TILE_THREADS) cnt[c] = 0;
    if (threadIdx.x < 3) sh_red[threadIdx.x] = 0;
    __syncthreads();
    // 1. histogram by column (the segment is three arrays of `cap` words: this pass reads the first; four loads in flight)
    const uint32_t cap = a.tile_cap[band];
    for (uint32_t i0 = threadIdx.x; i0 < F; i0 += 4 * TILE_THREADS) {
        uint32_t c4[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) c4[u] = (i0 + u * TILE_THREADS < F) ? frags[i0 + u * TILE_THREADS] : 0xffffffffu;
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (c4[u] != 0xffffffffu) atomicAdd(&cnt[c4[u] & 0xffffffu], 1u);
    }
    __syncthreads();
    pc.mark(0);
    // 2. column offsets
    block_scan_array<TILE_THREADS>(cnt, off, C, sh_warp);
    for (uint32_t c = threadIdx.x; c < C; c += TILE_THREADS) cnt[c] = 0;
    __syncthreads();
    pc.mark(1);
    // 3. scatter into column order (order inside a column is arbitrary; the replay sorts by triangle)
    uint32_t fown = 0;  // fragments of the band's own rows (halo rows are also some other band's own)
    for (uint32_t i0 = threadIdx.x; i0 < F; i0 += 2 * TILE_THREADS) {
        uint32_t r0[2], r1[2], r2[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const uint32_t i = i0 + u * TILE_THREADS;
            const bool in = i < F;
            r0[u] = in ? frags[i] : 0xffffffffu;
            r1[u] = in ? frags[cap + i] : 0u;
            r2[u] = in ? frags[2ull * cap + i] : 0u;
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            if (i0 + u * TILE_THREADS >= F) continue;
            const uint32_t cell = r0[u] & 0xffffffu;
            const uint32_t pos = off[cell] + atomicAdd(&cnt[cell], 1u);
            rec[pos] = make_uint2((r0[u] & 0xff000000u) | (r2[u] & 0xffffffu), r1[u]);
            fown += cell >= oc0 && cell < oc1;
        }
    }
    __syncthreads();
    pc.mark(2);
    // 4. per-column replay (rasterization.rs:51-168), spans built in place over the consumed prefix
    const int thr = a.merge_thr;
    // Only columns with two or more fragments need a replay (an empty column, or a single fragment which is its own
    // span, is done).  They are listed first (soff is free until the spans are counted), so that every lane of every
    // warp has a column to work on instead of two lanes in three idling next to the busy ones.
    if (threadIdx.x == 0) sh_red[3] = 0;
    __syncthreads();
    for (uint32_t c0 = 0; c0 < C; c0 += TILE_THREADS) {
        const uint32_t c = c0 + threadIdx.x;
        const bool w = c < C && cnt[c] >= 2;
        const unsigned m = __ballot_sync(0xffffffffu, w);
        uint32_t base = 0;
        if ((threadIdx.x & 31) == 0 && m) base = atomicAdd(&sh_red[3], (uint32_t)__popc(m));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (w) soff[base + __popc(m & ((1u << (threadIdx.x & 31)) - 1u))] = (uint16_t)c;
    }
    __syncthreads();
    const uint32_t nwork = sh_red[3];
    for (uint32_t k = threadIdx.x; k < nwork; k += TILE_THREADS) {
        const uint32_t c = soff[k];
        const uint32_t n = cnt[c];
        RCB_A(uint2) r = rec + off[c];
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
    // 5. span offsets (band-local compact order)
    block_scan_array<TILE_THREADS>(cnt, soff, C, sh_warp);
    pc.mark(4);

    // 6. filters, fused (see k_filter.cu for why one bottom-up walk per column is exact); own rows only - the halo rows are
    //    here for their geometry (the ledge filter reads the four neighbour columns' spans, never their areas)
    uint32_t walkable = 0, maxh = 0;
    const uint32_t fmask = a.fmask;
    const int height = td.walkable_height, climb = td.walkable_climb, neg_climb = td.neg_climb;
    for (uint32_t c = oc0 + threadIdx.x; c < oc1; c += TILE_THREADS) {
        const uint32_t b = off[c], e = b + cnt[c];
        if (b == e) continue;
        const int zl = (int)rcb_div_w(c, wmagic), x = (int)c - zl * W;
        const int z = zl + bd.hz0;  // row in the tile
        const int dx[4] = {0, 1, 0, -1}, dz[4] = {-1, 0, 1, 0};
        uint32_t nb[4], ne[4];
        bool oob[4];
#pragma unroll
        for (int d = 0; d < 4; ++d) {
            const int nx = x + dx[d], nz = z + dz[d];
            oob[d] = nx < 0 || nz < 0 || nx >= W || nz >= H;
            const uint32_t nc = oob[d] ? c : (uint32_t)((nz - bd.hz0) * W + nx);
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
    // 7. outputs for the own rows: heightfield CSR offsets (band-local; k_tile_compact makes them tile-relative), spans in
    //    band-local compact order
    RCB_A(uint32_t) out_off = RCB_MK(uint32_t, a.hf_cell_offset + td.cell_base + bd.tile + (uint32_t)(W * bd.z0), oc1 - oc0);
    RCB_A(uint32_t) ts_mm = RCB_MK(uint32_t, a.ts_mm + a.tile_base[band], a.tile_cap[band]);
    RCB_A(uint8_t) ts_area = RCB_MK(uint8_t, a.ts_area + a.tile_base[band], a.tile_cap[band]);
    const uint32_t so0 = soff[oc0];
    for (uint32_t c = oc0 + threadIdx.x; c < oc1; c += TILE_THREADS) {
        const uint32_t so = soff[c] - so0, n = cnt[c], b = off[c];
        out_off[c - oc0] = so;
        for (uint32_t k = 0; k < n; ++k) {
            const uint2 sp = rec[b + k];
            ts_mm[so + k] = sp.y;
            ts_area[so + k] = (uint8_t)(sp.x >> 24);
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        walkable += __shfl_down_sync(0xffffffffu, walkable, d);
        maxh = max(maxh, __shfl_down_sync(0xffffffffu, maxh, d));
        fown += __shfl_down_sync(0xffffffffu, fown, d);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&sh_red[0], walkable);
        atomicMax(&sh_red[1], maxh);
        atomicAdd(&sh_red[2], fown);
    }
    __syncthreads();
    pc.mark(6);
    if (threadIdx.x == 0) {
        if (sh_red[2]) atomicAdd(&a.counters[1], (unsigned long long)sh_red[2]);
        BandInfo& bi = a.binfo[band];
        bi.span_count = soff[oc1] - so0;
        bi.top_spans = soff[oc0 + (uint32_t)W] - so0;
        bi.bot_spans = soff[oc1] - soff[oc1 - (uint32_t)W];
        bi.walkable = sh_red[0];
        bi.max_height = sh_red[1];
        if (a.pass == 1) bi.status &= ~DEFERRED;
    }
}

// Shared memory k_tile_compact needs for a band that holds `held_cells` cells (own + halo rows) with `held_spans` spans, of
// which `top` / `bot` belong to the halo rows: coff u32[own_cells + 1] | mm u32 (aliased by the two u16 sweep buffers, which
// also hold a dummy span and the ghost values handed over by the neighbour bands) | nid u16[8][held + dummy] | soff u16 | areas u8.
__host__ __device__ inline size_t compact_smem_need(uint32_t own_cells, uint32_t held_cells, uint32_t held_spans, uint32_t top, uint32_t bot) {
    const size_t G2 = ((size_t)held_spans + 1 + 4ull * top + bot + 1) & ~(size_t)1;
    const size_t S2 = ((size_t)held_spans + 2) & ~(size_t)1;
    return 4ull * (own_cells + 1) + 4ull * G2 + 16ull * S2 + 2ull * (((size_t)held_cells + 2) & ~(size_t)1) + held_spans + 64;
}

// span_base / hand-over base per band (exclusive scans over bands), per-tile totals, batch totals; one block.
// counters[6] = S, [7] = walkable spans, [8] = largest compact_smem_need, [10] = hand-over bytes, [11] = most fragments in a band
__global__ void __launch_bounds__(1024) k_tile_scan(const BandDesc* __restrict__ bands, BandInfo* __restrict__ binfo, int nbands,
                                                    const TileDesc* __restrict__ tiles, const uint32_t* __restrict__ band_first,
                                                    TileInfo* __restrict__ tinfo, int ntiles, unsigned long long* counters,
                                                    unsigned long long span_cap, unsigned long long pub_cap, int compact,
                                                    const uint32_t* __restrict__ seg_cursor, const uint32_t* __restrict__ seg_cap) {
    __shared__ uint32_t sh[33], sh2[33];
    __shared__ unsigned long long sh_walk;
    __shared__ uint32_t sh_need, sh_fmax;
    if (threadIdx.x == 0) sh_walk = 0, sh_need = 0, sh_fmax = 0;
    __syncthreads();
    uint32_t carry = 0, carry2 = 0, need = 0, fmax = 0;
    unsigned long long walk = 0;
    for (int base = 0; base < nbands; base += 1024) {
        const int b = base + threadIdx.x;
        uint32_t v = 0, p = 0;
        if (b < nbands) {
            const BandInfo bi = binfo[b];
            const BandDesc bd = bands[b];
            v = bi.span_count;
            walk += bi.walkable;
            fmax = max(fmax, min(seg_cursor[b], seg_cap[b]));
            const bool has_top = bd.hz0 < bd.z0, has_bot = bd.hz1 > bd.z1;
            // what this band hands over: 16 B per span of its last row (to the next band) and of its first row (to the previous)
            p = (has_bot ? 16u * bi.bot_spans : 0u) + (has_top ? 16u * bi.top_spans : 0u);
            if (compact) {
                const uint32_t W = (uint32_t)tiles[bd.tile].width;
                const uint32_t top = has_top ? binfo[b - 1].bot_spans : 0u, bot = has_bot ? binfo[b + 1].top_spans : 0u;
                const size_t n = compact_smem_need(W * (uint32_t)(bd.z1 - bd.z0), W * (uint32_t)(bd.hz1 - bd.hz0), v + top + bot, top, bot);
                need = max(need, (uint32_t)min(n, (size_t)0xffffffffu));
                if ((size_t)v + 1 + 5ull * top + 2ull * bot >= 0xfff0u) need = 0xffffffffu;  // band-local span ids are 16 bit
            }
        }
        const uint32_t inc = warp_incl_scan(v), inc2 = warp_incl_scan(p);
        if ((threadIdx.x & 31) == 31) sh[threadIdx.x >> 5] = inc, sh2[threadIdx.x >> 5] = inc2;
        __syncthreads();
        if (threadIdx.x < 32) {
            const uint32_t w = sh[threadIdx.x], w2 = sh2[threadIdx.x];
            const uint32_t winc = warp_incl_scan(w), winc2 = warp_incl_scan(w2);
            sh[threadIdx.x] = winc - w, sh2[threadIdx.x] = winc2 - w2;
            if (threadIdx.x == 31) sh[32] = winc, sh2[32] = winc2;
        }
        __syncthreads();
        if (b < nbands) {
            binfo[b].span_base = carry + inc - v + sh[threadIdx.x >> 5];
            binfo[b].pub_base = carry2 + inc2 - p + sh2[threadIdx.x >> 5];
        }
        carry += sh[32];
        carry2 += sh2[32];
        __syncthreads();
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {  // one shared-memory atomic per warp, not per thread
        walk += __shfl_down_sync(0xffffffffu, walk, d);
        need = max(need, __shfl_down_sync(0xffffffffu, need, d));
        fmax = max(fmax, __shfl_down_sync(0xffffffffu, fmax, d));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&sh_walk, walk);
        atomicMax(&sh_need, need);
        atomicMax(&sh_fmax, fmax);
    }
    __syncthreads();  // also: every band's span_base is visible to the block
    // per-tile totals (TileInfo is what the C ABI reports; bands are an internal unit)
    for (int t = threadIdx.x; t < ntiles; t += 1024) {
        const uint32_t b0 = band_first[t], b1 = band_first[t + 1];
        uint32_t sc = 0, wk = 0, mh = 0;
        for (uint32_t b = b0; b < b1; ++b) {
            sc += binfo[b].span_count;
            wk += binfo[b].walkable;
            mh = max(mh, binfo[b].max_height);
        }
        tinfo[t].span_base = b1 > b0 ? binfo[b0].span_base : carry;
        tinfo[t].span_count = sc;
        tinfo[t].walkable_span_count = wk;
        tinfo[t].max_height = mh;
    }
    if (threadIdx.x == 0) {
        counters[6] = carry;
        counters[7] = sh_walk;
        counters[8] = sh_need;
        counters[10] = carry2;
        counters[11] = sh_fmax;  // the most fragments any band received (its bound, counters[5], is about twice that)
        // speculative issue: more spans than the arena was laid out for, or a band that did not fit shared memory
        if (carry > span_cap || carry2 > pub_cap || (span_cap != ~0ull && (counters[2] & 4ull))) atomicOr(&counters[2], RCB_FLAG_ABORT);
    }
}

// tile_base = exclusive scan of the per-tile fragment bounds; counters[5] = largest bound; one block
__global__ void __launch_bounds__(1024) k_tile_segments(const uint32_t* __restrict__ tile_ub, uint32_t* __restrict__ tile_base,
                                                        int ntiles, unsigned long long* counters, unsigned long long cap_frags,
                                                        unsigned long long cap_items) {
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
        if (carry > cap_frags || counters[4] > cap_items) atomicOr(&counters[2], RCB_FLAG_ABORT);
    }
}

// ======================================================================================================
// B: spans -> compact heightfield, connections, erosion, distance field
// ======================================================================================================
// Band-local span numbering: [0, nt) the previous band's last row (top halo), [own0, own1) the band's own spans,
// [own1, S) the next band's first row (bottom halo), S the dummy span, [G0, G0 + 4 nt) four ghost values per top-halo span,
// [G0 + 4 nt, +nbot) one per bottom-halo span.  The numbering is contiguous in the tile's span order, so
// tile-relative index = band_off + local - nt for own and halo spans alike.
struct TileB {
    RCB_A(uint16_t) soff;  // [held cells + 1] first span of every held cell (band-local ids)
    RCB_A(uint8_t) areas;  // [S] chf.areas (erosion writes it)
    RCB_A(uint16_t) nid;   // [8][S2] band-local id of the neighbour span per direction; S = none.  Entry S of
                           // every row is a dummy span whose neighbours are itself (branch-free sweeps)
    uint32_t S, S2;        // S2 >= S + 1: row pitch of nid
    uint32_t own0, own1;   // the band's own spans
    uint32_t nt, nbot, G0; // halo spans above / below, first ghost slot
    int W, rows, row0;     // own rows are the held rows [row0, row0 + rows)
    // hand-over between the bands of a tile (nullptr / 0 for a band without that neighbour)
    uint32_t band;
    bool has_top, has_bot;
    RCB_A(const uint4) in_top;   // previous band's records for its last row = my top halo
    RCB_A(const uint4) in_bot;   // next band's records for its first row = my bottom halo
    RCB_A(uint4) out_down;       // my last row, for the next band
    RCB_A(uint4) out_up;         // my first row, for the previous band (16 B records keep every region 16-byte aligned)
    uint32_t* flag_f;
    uint32_t* flag_d;
};

__device__ __forceinline__ uint32_t tnei(const TileB& g, uint32_t s, int dir) { return g.nid[(uint32_t)dir * g.S2 + s]; }

template <typename T>
__device__ __forceinline__ T sat_add(T v, unsigned b) {
    const unsigned r = (unsigned)v + b;
    const unsigned mx = (T)~(T)0;
    return (T)(r > mx ? mx : r);
}

__device__ __forceinline__ void wait_flag(const uint32_t* flag) {
    while (*reinterpret_cast<const volatile uint32_t*>(flag) == 0) __nanosleep(64);
    __threadfence();
}

// boundary -> forward rows -> backward gather, on shared memory.  ERODE: thresholds areas; else D + max.
// The forward sweep occupies only the warps that own a column; `idle_work(rank, count)` runs meanwhile on the
// others (work that touches neither the sweep buffers nor areas).  Returns whether idle_work ran.
// Bands (never with ERODE): the sweep starts from the previous band's last row and hands its own last row on; the
// gather and the blur read one / two rows beyond the band, which arrive as halo + ghost values (see TileB).
template <int TILE_THREADS, typename T, bool ERODE, class Work>
__device__ bool tile_sweep(const TileB& g, RCB_A(T) init, RCB_A(T) Fw, unsigned thr, uint32_t* sh_max, PhaseClock* pc, Work&& idle_work) {
    const int W = g.W, H = g.rows;
    const uint32_t TN = g.S;  // "no neighbour" == the dummy span
    const T inf = (T)~(T)0;
    // boundary marking (compact_heightfield.rs:556-585 / area.rs:91-125)
    for (uint32_t s = g.own0 + threadIdx.x; s < g.own1; s += TILE_THREADS) {
        const uint8_t ar = g.areas[s];
        int cnt = 0;
        if (!ERODE || ar != 0) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t n = tnei(g, s, 7 - 2 * k);
                if (n == TN) continue;
                if (ERODE ? (g.areas[n] != 0) : (g.areas[n] == ar)) cnt++;
            }
        }
        init[s] = (cnt == 4) ? inf : (T)0;
    }
    if (threadIdx.x == 0) init[g.S] = Fw[g.S] = inf;  // the dummy span never lowers a minimum
    __syncthreads();
    if (pc) pc->mark(8);
    // forward: row z needs row z-1 final and row z init (see k_dist.cu).  The chain is one warp-instruction
    // stream deep (H dependent steps), so the body is branch-free: missing neighbours point at the dummy
    // span (value inf), and a + 2 never needs clamping because the running minimum is already <= inf.
    // Only the warps that own a column take part in the H row barriers (named barrier 1).
    const int sweep_threads = min(TILE_THREADS, ((W + 31) / 32) * 32);
    if ((int)threadIdx.x < sweep_threads) {
        RCB_AR(uint16_t) soff = g.soff + g.row0 * W;  // own rows
        RCB_AR(uint16_t) nid0 = g.nid;               // NW
        RCB_AR(uint16_t) nid2 = g.nid + 2 * g.S2;    // NE
        RCB_AR(uint16_t) nid3 = g.nid + 3 * g.S2;    // E
        const uint32_t DUMMY = g.S;
        auto relax = [&](uint32_t s, uint32_t a0, uint32_t a1, uint32_t b0, uint32_t b1) {
            uint32_t v = init[s];
            v = min(v, (uint32_t)Fw[a0] + 2u);
            v = min(v, (uint32_t)Fw[a1] + 3u);
            v = min(v, (uint32_t)init[b0] + 2u);
            v = min(v, (uint32_t)Fw[b1] + 3u);
            return v;
        };
        auto row_barrier = [&]() {
            if (sweep_threads == 32)
                __syncwarp();
            else
                asm volatile("bar.sync 1, %0;" ::"r"(sweep_threads) : "memory");
        };
        if (!ERODE && g.has_top) {
            // the previous band's last row: its forward values, and through three ghost slots per span the values of that
            // span's E, N and NW neighbours (the only things the relaxation and the gather reach THROUGH a halo span)
            if (threadIdx.x == 0) {
                const long long t0 = pc && pc->acc ? clock64() : 0;
                wait_flag(&g.flag_f[g.band - 1]);
                if (pc && pc->acc) atomicAdd(&pc->acc[10], (unsigned long long)(clock64() - t0));
            }
            row_barrier();
            for (uint32_t j = threadIdx.x; j < g.nt; j += (uint32_t)sweep_threads) {
                const uint4 r = __ldcg(&g.in_top[j]);
                Fw[j] = (T)(r.x & 0xffffu);
                Fw[g.G0 + 4 * j] = (T)(r.x >> 16);
                Fw[g.G0 + 4 * j + 1] = (T)(r.y & 0xffffu);
                Fw[g.G0 + 4 * j + 2] = (T)(r.y >> 16);
            }
            row_barrier();
        }
        if (W <= sweep_threads) {
            // One column per thread, software-pipelined three rows deep: the neighbour ids of a span do not
            // depend on distances, so while row z is relaxed the ids of row z+1 (second hop), row z+2 (first
            // hop) and the cell offsets of row z+3 are already in flight.  The dependent chain per row
            // shrinks from five shared-memory round trips to one.
            const int x = threadIdx.x;
            const bool col = x < W;
            auto cell_span = [&](int z, uint32_t& sb, uint32_t& n) {  // first span + count of cell (x, z)
                sb = DUMMY;
                n = 0;
                if (col && z < H) {
                    const uint32_t b = soff[z * W + x], e = soff[z * W + x + 1];
                    n = e - b;
                    sb = n ? b : DUMMY;
                }
            };
            uint32_t s0, n0, s1, n1, s2, n2, s3, n3;
            cell_span(0, s0, n0);
            cell_span(1, s1, n1);
            cell_span(2, s2, n2);
            uint32_t a0 = nid0[s0], b0 = nid3[s0];
            uint32_t a1 = nid3[a0], b1 = nid2[b0];
            uint32_t a0n = nid0[s1], b0n = nid3[s1];
            for (int z = 0; z < H; ++z) {
                // everything below is independent of this row's result
                const uint32_t a1n = nid3[a0n], b1n = nid2[b0n];    // row z+1, second hop
                const uint32_t a0nn = nid0[s2], b0nn = nid3[s2];    // row z+2, first hop
                cell_span(z + 3, s3, n3);                           // row z+3, cell offsets
                if (n0) {
                    Fw[s0] = (T)relax(s0, a0, a1, b0, b1);
                    for (uint32_t s = s0 + 1; s < s0 + n0; ++s) {  // further spans of the column: not pipelined
                        const uint32_t c0 = nid0[s], d0 = nid3[s];
                        Fw[s] = (T)relax(s, c0, nid3[c0], d0, nid2[d0]);
                    }
                }
                row_barrier();
                s0 = s1, n0 = n1, a0 = a0n, b0 = b0n, a1 = a1n, b1 = b1n;
                s1 = s2, n1 = n2, a0n = a0nn, b0n = b0nn;
                s2 = s3, n2 = n3;
            }
        } else {
            for (int z = 0; z < H; ++z) {
                for (int x = threadIdx.x; x < W; x += sweep_threads) {
                    const uint32_t c = (uint32_t)(z * W + x);
                    for (uint32_t s = soff[c]; s < soff[c + 1]; ++s) {
                        const uint32_t c0 = nid0[s], d0 = nid3[s];  // NW (row z-1), E (row z: init value)
                        Fw[s] = (T)relax(s, c0, nid3[c0], d0, nid2[d0]);  // NW.E, E.NE (row z-1)
                    }
                }
                row_barrier();
            }
        }
        if (!ERODE && g.has_bot) {
            // hand the last own row's forward values to the next band (+ those of each span's E, N, NW neighbours) as soon
            // as the sweep is through: the next band's sweep waits for exactly this, not for this band's connection list
            const uint32_t lr0s = g.soff[(g.row0 + g.rows - 1) * W];
            for (uint32_t s = lr0s + threadIdx.x; s < g.own1; s += (uint32_t)sweep_threads) {
                uint4 r;
                r.x = (uint32_t)Fw[s] | ((uint32_t)Fw[tnei(g, s, 3)] << 16);
                r.y = (uint32_t)Fw[tnei(g, s, 1)] | ((uint32_t)Fw[tnei(g, s, 0)] << 16);
                r.z = r.w = 0;
                g.out_down[s - lr0s] = r;
            }
            __threadfence();
            row_barrier();
            if (threadIdx.x == 0) *reinterpret_cast<volatile uint32_t*>(&g.flag_f[g.band]) = 1u;
        }
        if (pc && pc->acc && threadIdx.x == 0) atomicAdd(&pc->acc[12], (unsigned long long)(clock64() - pc->t));  // sweep incl. its wait
    } else {
        const long long t0 = pc && pc->acc ? clock64() : 0;
        idle_work(threadIdx.x - (uint32_t)sweep_threads, (uint32_t)(TILE_THREADS - sweep_threads));
        if (pc && pc->acc && (int)threadIdx.x == sweep_threads) atomicAdd(&pc->acc[13], (unsigned long long)(clock64() - t0));  // connection list
    }
    __syncthreads();
    if (pc) pc->mark(9);
    const uint32_t lr0 = g.soff[(g.row0 + g.rows - 1) * W], fr1 = g.soff[(g.row0 + 1) * W];  // last own row = [lr0, own1), first = [own0, fr1)
    // backward = gather from forward values; D overwrites init (init is dead)
    uint32_t vmax = 0;
    for (uint32_t s = g.own0 + threadIdx.x; s < g.own1; s += TILE_THREADS) {
        const uint32_t a0 = tnei(g, s, 2), b0 = tnei(g, s, 1);    // NE, N
        const uint32_t a1 = tnei(g, a0, 1), b1 = tnei(g, b0, 0);  // NE.N, N.NW
        uint32_t v = Fw[s];
        v = min(v, (uint32_t)Fw[a0] + 2u);
        v = min(v, (uint32_t)Fw[a1] + 3u);
        v = min(v, (uint32_t)Fw[b0] + 2u);
        v = min(v, (uint32_t)Fw[b1] + 3u);
        init[s] = (T)v;
        if (!ERODE) vmax = max(vmax, v);
    }
    if (!ERODE) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) vmax = max(vmax, __shfl_down_sync(0xffffffffu, vmax, d));
        if ((threadIdx.x & 31) == 0 && vmax) atomicMax(sh_max, vmax);
    }
    __syncthreads();
    if (ERODE) {
        for (uint32_t s = g.own0 + threadIdx.x; s < g.own1; s += TILE_THREADS)
            if ((unsigned)init[s] < thr) g.areas[s] = 0;  // area.rs:226-231
        __syncthreads();
    }
    if (!ERODE && (g.has_top || g.has_bot)) {
        // The box blur reads D one row beyond the band, and through that row's spans one row further (N then NE, S then SW):
        // each band publishes D of its border rows together with D of the NE / SW neighbour (and whether there is one).
        if (g.has_bot)
            for (uint32_t s = lr0 + threadIdx.x; s < g.own1; s += TILE_THREADS) {
                const uint32_t ne = tnei(g, s, 2);
                uint4 r = g.out_down[s - lr0];
                r.z = (uint32_t)init[s] | ((uint32_t)(ne != TN ? init[ne] : (T)0) << 16);
                r.w = ne != TN ? 1u : 0u;
                g.out_down[s - lr0] = r;
            }
        if (g.has_top)
            for (uint32_t s = g.own0 + threadIdx.x; s < fr1; s += TILE_THREADS) {
                const uint32_t sw = tnei(g, s, 6);
                g.out_up[s - g.own0] = make_uint4((uint32_t)init[s] | ((uint32_t)(sw != TN ? init[sw] : (T)0) << 16), sw != TN ? 1u : 0u, 0u, 0u);
            }
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) {
            *reinterpret_cast<volatile uint32_t*>(&g.flag_d[g.band]) = 1u;
            const long long t0 = pc && pc->acc ? clock64() : 0;
            if (g.has_top) wait_flag(&g.flag_d[g.band - 1]);
            if (g.has_bot) wait_flag(&g.flag_d[g.band + 1]);
            if (pc && pc->acc) atomicAdd(&pc->acc[11], (unsigned long long)(clock64() - t0));
        }
        __syncthreads();
        RCB_A(uint16_t) nid2 = g.nid + 2 * g.S2;
        RCB_A(uint16_t) nid6 = g.nid + 6 * g.S2;
        for (uint32_t j = threadIdx.x; j < g.nt; j += TILE_THREADS) {
            const uint4 r = __ldcg(&g.in_top[j]);
            init[j] = (T)(r.z & 0xffffu);
            init[g.G0 + 4 * j + 3] = (T)(r.z >> 16);
            nid2[j] = (uint16_t)(r.w ? g.G0 + 4 * j + 3 : TN);
        }
        for (uint32_t j = threadIdx.x; j < g.nbot; j += TILE_THREADS) {
            const uint4 r = __ldcg(&g.in_bot[j]);
            init[g.own1 + j] = (T)(r.x & 0xffffu);
            init[g.G0 + 4 * g.nt + j] = (T)(r.x >> 16);
            nid6[g.own1 + j] = (uint16_t)(r.y ? g.G0 + 4 * g.nt + j : TN);
        }
        __syncthreads();
    }
    return sweep_threads < TILE_THREADS;
}

// Decoupled look-back over bands, split in two so that nobody waits: a band PUBLISHES its connection count
// as soon as it is known and RESOLVES its exclusive prefix only when it needs it (after the distance field),
// by which time every predecessor has published.  look[t] = flag << 62 | value; flag 1 = band aggregate,
// 2 = inclusive prefix.  Band ids come from a ticket counter, so predecessors are running or done.
__device__ __forceinline__ void lookback_publish(volatile unsigned long long* look, uint32_t tile, uint32_t mine) {
    look[tile] = ((tile == 0 ? 2ull : 1ull) << 62) | mine;
}
__device__ uint32_t lookback_resolve(volatile unsigned long long* look, uint32_t tile, uint32_t mine) {  // warp 0, all lanes
    const uint32_t lane = threadIdx.x & 31;
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
// inclusive prefix of band `t` (t < my band): it resolves without waiting for anything after it
__device__ __forceinline__ uint32_t lookback_inclusive(volatile unsigned long long* look, int t) {
    if (t < 0) return 0;
    unsigned long long v;
    do {
        v = look[t];
    } while ((v >> 62) != 2);
    return (uint32_t)(v & 0xffffffffull);
}

template <int TILE_THREADS>
__global__ void __launch_bounds__(TILE_THREADS, 1024 / TILE_THREADS) k_tile_compact(TileCompactArgs a) {
    extern __shared__ __align__(16) uint32_t smem[];
    __shared__ uint32_t sh_warp[TILE_THREADS / 32 + 1];
    __shared__ uint32_t sh_misc[6];  // [0] band id, [1] conn base, [2] max distance, [3] link work-list length, [4] tile conn base
    if (a.counters[2] & RCB_FLAG_ABORT) return;  // uniform over the grid: nobody is left waiting in the look-back
    // dynamic band id: bands start in order, which the look-back and the hand-over between bands rely on
    if (threadIdx.x == 0) {
        sh_misc[0] = atomicAdd(a.tile_ticket, 1u);
        sh_misc[2] = 0;
        sh_misc[3] = 0;
    }
    __syncthreads();
    const uint32_t band = sh_misc[0];
    const BandDesc bd = a.bands[band];
    const uint32_t tile = bd.tile;
    const TileDesc& td = a.tiles[tile];
    const int W = td.width, H = td.height;
    const int rows = bd.z1 - bd.z0, row0 = bd.z0 - bd.hz0;
    const uint32_t C = (uint32_t)(W * rows);                 // own cells
    const uint32_t Ch = (uint32_t)(W * (bd.hz1 - bd.hz0));  // held cells
    const uint32_t oc0 = (uint32_t)(W * row0);
    const unsigned long long wmagic = td.wmagic;
    const bool has_top = bd.hz0 < bd.z0, has_bot = bd.hz1 > bd.z1;
    const BandInfo bi = a.binfo[band];
    const uint32_t So = bi.span_count;        // own spans
    const uint32_t span_base = bi.span_base;  // batch-wide index of the first own span
    const uint32_t nt = has_top ? a.binfo[band - 1].bot_spans : 0u, nbot = has_bot ? a.binfo[band + 1].top_spans : 0u;
    const uint32_t S = nt + So + nbot;  // held spans
    const uint32_t own0 = nt, own1 = nt + So;
    const uint32_t tile_span_base = a.tinfo[tile].span_base;
    const uint32_t band_off = span_base - tile_span_base;  // tile-relative index of the first own span
    // layout: coff u32[C+1] | mm u32[G2] (aliased later by d16a,d16b u16[G2]) | nid u16[8][S2] | soff u16[Ch+1] | areas u8[S]
    const uint32_t S2 = (S + 2) & ~1u;  // >= S + 1: one dummy span behind the real ones
    const uint32_t TN = S;              // "no neighbour" is the dummy span's id
    const uint32_t G0 = S + 1;
    const uint32_t G2 = (G0 + 4 * nt + nbot + 1) & ~1u;
    const size_t need = compact_smem_need(C, Ch, S, nt, nbot);
    const bool fits = a.do_compact ? (need <= a.smem_bytes && G2 < 0xfff0u) : true;
    RCB_A(uint32_t) coff = RCB_MK(uint32_t, smem, C + 1);
    RCB_A(uint32_t) mm = RCB_MK(uint32_t, smem + (C + 1), G2);
    RCB_A(uint16_t) d16a = RCB_MK(uint16_t, smem + (C + 1), G2);
    RCB_A(uint16_t) d16b = RCB_MK(uint16_t, reinterpret_cast<uint16_t*>(smem + (C + 1)) + G2, G2);
    RCB_A(uint16_t) nid = RCB_MK(uint16_t, smem + (C + 1) + G2, 8ull * S2);
    RCB_A(uint16_t) soff = RCB_MK(uint16_t, reinterpret_cast<uint16_t*>(smem + (C + 1) + G2) + 8ull * S2, Ch + 1);
    RCB_A(uint8_t) areas = RCB_MK(uint8_t, reinterpret_cast<uint16_t*>(smem + (C + 1) + G2) + 8ull * S2 + ((Ch + 2) & ~1u), S);

    volatile unsigned long long* look = a.lookback;
    // band-local offsets written by k_tile_spans, tile cell order (W * H + 1 words per tile in both arrays)
    RCB_A(const uint32_t) g_off = RCB_MK(const uint32_t, a.hf_cell_offset_in + td.cell_base + tile, (uint32_t)(W * H) + 1u);
    RCB_A(uint32_t) g_off_out = RCB_MK(uint32_t, a.hf_cell_offset + td.cell_base + tile, (uint32_t)(W * H) + 1u);
    const uint32_t tb = a.tile_base[band];
    // this band's slices of the batch-wide arrays
    RCB_A(const uint32_t) my_mm = RCB_MK(const uint32_t, a.ts_mm + tb, So);
    RCB_A(const uint8_t) my_ar = RCB_MK(const uint8_t, a.ts_area + tb, So);
    RCB_A(int16_t) o_smin = RCB_MK(int16_t, a.smin + span_base, So);
    RCB_A(int16_t) o_smax = RCB_MK(int16_t, a.smax + span_base, So);
    RCB_A(uint8_t) o_sarea = RCB_MK(uint8_t, a.sarea + span_base, So);
    const uint32_t c_tile0 = (uint32_t)(W * bd.z0);  // first own cell in the tile
    if (!a.do_compact || !fits) {
        // heightfield-only result, or a band too large for shared memory (the host then reruns the batch with smaller
        // bands or on the global back end): copy the spans through, keep the look-back chain and the hand-over alive
        for (uint32_t c = threadIdx.x; c < C; c += TILE_THREADS) g_off_out[c_tile0 + c] = band_off + g_off[c_tile0 + c];
        if (bd.z1 == H && threadIdx.x == 0) g_off_out[(uint32_t)(W * H)] = a.tinfo[tile].span_count;
        for (uint32_t s = threadIdx.x; s < So; s += TILE_THREADS) {
            const uint32_t w = my_mm[s];
            o_smin[s] = (int16_t)(w & 0xffffu);
            o_smax[s] = (int16_t)(w >> 16);
            o_sarea[s] = my_ar[s];
        }
        if (a.do_compact && threadIdx.x < 32) {
            if (threadIdx.x == 0) {
                atomicOr(&a.counters[2], 4ull);
                lookback_publish(look, band, 0);
                *reinterpret_cast<volatile uint32_t*>(&a.flag_f[band]) = 1u;
                *reinterpret_cast<volatile uint32_t*>(&a.flag_d[band]) = 1u;
            }
            __syncwarp();
            const uint32_t excl = lookback_resolve(look, band, 0);
            if (threadIdx.x == 0 && bd.z0 == 0) {
                a.tinfo[tile].conn_base = excl;
                a.tinfo[tile].conn_count = 0;
            }
        }
        return;
    }
    PhaseClock pc(a.phase ? a.phase + 8 : nullptr);
    // 1. load the band's spans and its halo rows; final span attributes (compact_heightfield.rs:231-237).  Loads are issued in
    //    batches of four per thread so their latencies overlap.
    {
        RCB_AR(const uint32_t) in_off = g_off + c_tile0;
        RCB_AR(const uint32_t) in_mm = my_mm;
        RCB_AR(const uint8_t) in_ar = my_ar;
        RCB_AR(int16_t) o_min = o_smin;
        RCB_AR(int16_t) o_max = o_smax;
        RCB_AR(uint8_t) o_ar = o_sarea;
        RCB_AR(uint16_t) so_own = soff + oc0;
        for (uint32_t c0 = threadIdx.x; c0 < C; c0 += 4 * TILE_THREADS) {
            uint32_t o[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) o[u] = (c0 + u * TILE_THREADS < C) ? in_off[c0 + u * TILE_THREADS] : 0u;
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (c0 + u * TILE_THREADS < C) {
                    so_own[c0 + u * TILE_THREADS] = (uint16_t)(nt + o[u]);
                    g_off_out[c_tile0 + c0 + u * TILE_THREADS] = band_off + o[u];
                }
        }
        if (threadIdx.x == 0) {
            so_own[C] = (uint16_t)own1;  // == soff of the first bottom-halo cell, or of the end
            soff[Ch] = (uint16_t)S;
            if (bd.z1 == H) g_off_out[(uint32_t)(W * H)] = a.tinfo[tile].span_count;
        }
        for (uint32_t s0 = threadIdx.x; s0 < So; s0 += 4 * TILE_THREADS) {
            uint32_t w[4];
            uint8_t ar[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const uint32_t i = s0 + u * TILE_THREADS;
                w[u] = i < So ? in_mm[i] : 0u;
                ar[u] = i < So ? in_ar[i] : (uint8_t)0;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const uint32_t i = s0 + u * TILE_THREADS;
                if (i < So) {
                    mm[own0 + i] = w[u];
                    areas[own0 + i] = ar[u];  // chf.areas == span area until erosion runs
                    o_min[i] = (int16_t)(w[u] & 0xffffu);
                    o_max[i] = (int16_t)(w[u] >> 16);
                    o_ar[i] = ar[u];
                }
            }
        }
        if (has_top) {  // the previous band's last row: offsets relative to that row's first span, spans at the end of its segment
            RCB_AR(const uint32_t) p_off = g_off + (c_tile0 - (uint32_t)W);
            const uint32_t pb = a.tile_base[band - 1] + (a.binfo[band - 1].span_count - nt);
            RCB_A(const uint32_t) p_mm = RCB_MK(const uint32_t, a.ts_mm + pb, nt);
            RCB_A(const uint8_t) p_ar = RCB_MK(const uint8_t, a.ts_area + pb, nt);
            const uint32_t row_first = p_off[0];
            for (uint32_t x = threadIdx.x; x < (uint32_t)W; x += TILE_THREADS) soff[x] = (uint16_t)(p_off[x] - row_first);
            for (uint32_t i = threadIdx.x; i < nt; i += TILE_THREADS) {
                mm[i] = p_mm[i];
                areas[i] = p_ar[i];
            }
        }
        if (has_bot) {  // the next band's first row: the head of its segment
            RCB_AR(const uint32_t) n_off = g_off + (c_tile0 + C);
            const uint32_t nb0 = a.tile_base[band + 1];
            RCB_A(const uint32_t) n_mm = RCB_MK(const uint32_t, a.ts_mm + nb0, nbot);
            RCB_A(const uint8_t) n_ar = RCB_MK(const uint8_t, a.ts_area + nb0, nbot);
            for (uint32_t x = threadIdx.x; x < (uint32_t)W; x += TILE_THREADS) soff[oc0 + C + x] = (uint16_t)(own1 + n_off[x]);
            for (uint32_t i = threadIdx.x; i < nbot; i += TILE_THREADS) {
                mm[own1 + i] = n_mm[i];
                areas[own1 + i] = n_ar[i];
            }
        }
    }
    __syncthreads();
    pc.mark(0);
    // 2. cells (compact_heightfield.rs:203-223); spans that cannot link get their defaults; the walkable ones are listed.
    //    Only ~45 % of the spans are walkable: linking straight out of the cell loop left most lanes idle next to the
    //    busy ones.  The list lives in the band's fragment segment, which is dead since k_tile_spans.
    RCB_A(uint32_t) work = RCB_MK(uint32_t, a.work + 3ull * tb, 3ull * a.tile_cap[band]);
    // per-band slices of the compact arrays (cells of the own rows; spans; the connection arrays up to their capacity)
    RCB_A(uint32_t) o_cell_index = RCB_MK(uint32_t, a.cell_index + td.cell_base + c_tile0, C);
    RCB_A(uint32_t) o_cell_count = RCB_MK(uint32_t, a.cell_count + td.cell_base + c_tile0, C);
    RCB_A(uint16_t) o_con = RCB_MK(uint16_t, a.con + 4ull * span_base, 4ull * So);
    RCB_A(uint32_t) o_first_conn = RCB_MK(uint32_t, a.first_conn + span_base, So);
    RCB_A(uint8_t) o_areas = RCB_MK(uint8_t, a.areas + span_base, So);
    RCB_A(uint16_t) o_dist = RCB_MK(uint16_t, a.dist + span_base, So);
    RCB_A(uint32_t) o_conn_ref = RCB_MK(uint32_t, a.conn_ref, a.Kcap);
    RCB_A(uint8_t) o_conn_dir = RCB_MK(uint8_t, a.conn_dir, a.Kcap);
    RCB_A(uint32_t) o_conn_next = RCB_MK(uint32_t, a.conn_next, a.Kcap);
    if (threadIdx.x == 0) sh_misc[3] = 0;
    // halo spans take part as link targets only: no neighbours of their own (the ghost links are set by the sweeps)
    for (uint32_t s = threadIdx.x; s < nt + nbot; s += TILE_THREADS) {
        const uint32_t sp = s < nt ? s : own1 + (s - nt);
#pragma unroll
        for (int d = 0; d < 8; ++d) nid[(uint32_t)d * S2 + sp] = (uint16_t)TN;
    }
    if (has_top) {  // what the sweeps reach THROUGH a top-halo span: its E, N and NW neighbours' values, as ghost spans
        for (uint32_t j = threadIdx.x; j < nt; j += TILE_THREADS) {
            nid[3u * S2 + j] = (uint16_t)(G0 + 4 * j);
            nid[1u * S2 + j] = (uint16_t)(G0 + 4 * j + 1);
            nid[0u * S2 + j] = (uint16_t)(G0 + 4 * j + 2);
        }
    }
    __syncthreads();
    for (uint32_t c0 = 0; c0 < C; c0 += TILE_THREADS) {
        const uint32_t c = c0 + threadIdx.x;
        uint32_t b = 0, e = 0, nw = 0;
        if (c < C) {
            b = soff[oc0 + c], e = soff[oc0 + c + 1];
            o_cell_index[c] = e > b ? band_off + (b - nt) : RCB_NONE;
            o_cell_count[c] = e - b;
            coff[c] = 0;
            for (uint32_t sp = b; sp < e; ++sp) {
                if (areas[sp] != 0) {
                    ++nw;
                } else {
#pragma unroll
                    for (int d = 0; d < 8; ++d) nid[(uint32_t)d * S2 + sp] = (uint16_t)TN;
                    *reinterpret_cast<uint2*>(&o_con[4ull * (sp - nt)]) = make_uint2(63u | (63u << 16), 63u | (63u << 16));
                }
            }
        }
        uint32_t incl = nw;  // warp-aggregated append: one atomic per warp
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, incl, d);
            if ((threadIdx.x & 31) >= (uint32_t)d) incl += v;
        }
        uint32_t base = 0;
        const uint32_t tot = __shfl_sync(0xffffffffu, incl, 31);
        if ((threadIdx.x & 31) == 31 && tot) base = atomicAdd(&sh_misc[3], tot);
        base = __shfl_sync(0xffffffffu, base, 31) + incl - nw;
        for (uint32_t sp = b; sp < e; ++sp)
            if (areas[sp] != 0) work[base++] = ((oc0 + c) << 16) | sp;
    }
    __syncthreads();
    // 3. link (:297-372) + con[4] (:392-427), one listed span per thread
    const uint32_t nwork = sh_misc[3];
    for (uint32_t k = threadIdx.x; k < nwork; k += TILE_THREADS) {
        const uint32_t item = work[k];
        const uint32_t c = item >> 16, sp = item & 0xffffu;  // c: held-cell index
        const int zl = (int)rcb_div_w(c, wmagic), x = (int)c - zl * W;
        const int z = zl + bd.hz0;  // row in the tile
        const uint32_t w = mm[sp];
        const int mn = (int16_t)(w & 0xffffu), mx = (int16_t)(w >> 16);
        uint32_t c4[4] = {63, 63, 63, 63};
        uint32_t nconn = 0;
#pragma unroll
        for (int d = 0; d < 8; ++d) {
            const int nx = x + t_dirx[d], nz = z + t_dirz[d];
            uint32_t best = TN;
            if (!(nx < 0 || nz < 0 || nx >= W || nz >= H)) {
                const uint32_t ncell = (uint32_t)((nz - bd.hz0) * W + nx);
                const uint32_t nb = soff[ncell], ne = soff[ncell + 1];
                int best_diff = 0x7fffffff;
                for (uint32_t ns = nb; ns < ne; ++ns) {
                    if (areas[ns] == 0) continue;
                    const uint32_t nw = mm[ns];
                    const int nmn = (int16_t)(nw & 0xffffu), nmx = (int16_t)(nw >> 16);
                    if (max(nmn, mn) > min(nmx, mx)) continue;
                    const int diff = (int)(uint16_t)(abs(nmn - mn) + abs(nmx - mx));
                    if (diff < best_diff) {
                        best_diff = diff;
                        best = ns;
                    }
                }
                if (best != TN) {
                    nconn++;
                    const uint32_t o = best - nb;  // offset inside the neighbour cell (:418-421)
                    if (d == 7) c4[0] = o;
                    if (d == 5) c4[1] = o;
                    if (d == 3) c4[2] = o;
                    if (d == 1) c4[3] = o;
                }
            }
            nid[(uint32_t)d * S2 + sp] = (uint16_t)best;
        }
        *reinterpret_cast<uint2*>(&o_con[4ull * (sp - nt)]) =
            make_uint2((c4[0] & 0xffffu) | (c4[1] << 16), (c4[2] & 0xffffu) | (c4[3] << 16));
        if (nconn) atomicAdd(&coff[c - oc0], nconn);
    }
    if (threadIdx.x < 8) nid[threadIdx.x * S2 + S] = (uint16_t)S;  // the dummy span's neighbours are itself
    __syncthreads();
    pc.mark(1);
    // 4. connection offsets inside the band; publish the band's count (nobody waits here)
    const uint32_t K = block_scan_array<TILE_THREADS>(coff, coff, C, sh_warp);
    if (threadIdx.x == 0) lookback_publish(look, band, K);
    pc.mark(2);
    TileB g;
    g.soff = soff, g.areas = areas, g.nid = nid, g.S = S, g.S2 = S2, g.own0 = own0, g.own1 = own1, g.nt = nt, g.nbot = nbot, g.G0 = G0;
    g.W = W, g.rows = rows, g.row0 = row0, g.band = band, g.has_top = has_top, g.has_bot = has_bot;
    g.flag_f = a.flag_f, g.flag_d = a.flag_d;
    g.in_top = RCB_MK(const uint4, has_top ? a.pub + a.binfo[band - 1].pub_base : nullptr, nt);
    g.out_down = RCB_MK(uint4, a.pub + bi.pub_base, has_bot ? bi.bot_spans : 0u);
    g.out_up = RCB_MK(uint4, a.pub + bi.pub_base + (has_bot ? 16ull * bi.bot_spans : 0ull), has_top ? bi.top_spans : 0u);
    g.in_bot = RCB_MK(const uint4, nullptr, 0);
    if (has_bot) {
        const BandInfo nb_i = a.binfo[band + 1];  // its bottom hand-over (if any) comes first, then the records for me
        const bool n_has_bot = a.bands[band + 1].hz1 > a.bands[band + 1].z1;
        g.in_bot = RCB_MK(const uint4, a.pub + nb_i.pub_base + (n_has_bot ? 16ull * nb_i.bot_spans : 0ull), nbot);
    }
    // 7 (hoisted). The connection list (compact_heightfield.rs:375-389: reverse direction order, next -> previously
    //    pushed, first_connection = last pushed) needs only the neighbour table and the tile's first connection
    //    index, so the warps that sit out the row-sequential forward sweep write it meanwhile.
    auto conn_work = [&](uint32_t rank, uint32_t nworkers) {
        if (rank < 32) {
            const uint32_t excl = lookback_resolve(look, band, K);
            // first_connection / next are tile-relative: the tile's base is the prefix in front of its first band
            const uint32_t tile_base_k = bd.z0 == 0 ? excl : lookback_inclusive(look, (int)a.band_first[tile] - 1);
            if (rank == 0) {
                sh_misc[1] = excl;
                sh_misc[4] = tile_base_k;
                if (bd.z0 == 0) a.tinfo[tile].conn_base = excl;
                if (K) {
                    atomicAdd(&a.tinfo[tile].conn_count, K);
                    atomicAdd(&a.counters[9], (unsigned long long)K);
                }
                if ((unsigned long long)excl + K > a.Kcap) {
                    atomicOr(&a.counters[2], RCB_FLAG_CONN);
                    sh_misc[1] = 0xffffffffu;
                }
            }
        }
        if (nworkers == TILE_THREADS)
            __syncthreads();
        else
            asm volatile("bar.sync 2, %0;" ::"r"(nworkers) : "memory");
        const uint32_t conn_base = sh_misc[1];
        if (conn_base == 0xffffffffu) return;  // the connection arena is too small for this band: the host redoes the batch
        const uint32_t krel = conn_base - sh_misc[4];  // the band's first connection, tile-relative
        const uint32_t ref_off = band_off - nt;        // band-local span id -> tile-relative
        // Each worker warp takes a contiguous run of cells, i.e. of spans and of the list, 32 spans at a time: a lane
        // counts its span's connections, then the lanes turn to the run's list SLOTS - slot t belongs to the span whose
        // range holds t (bisection over the lanes' offsets) and is its j-th present direction from 7 downwards - so the
        // three stores of a round are coalesced.  (One span per thread wrote its own 5-8 slots: ~20 sectors per store.)
        const uint32_t lane = rank & 31u, wid = rank >> 5, nwarps = nworkers >> 5;
        const uint32_t cpw = (C + nwarps - 1) / nwarps;
        const uint32_t cw0 = min(wid * cpw, C), cw1 = min(cw0 + cpw, C);
        uint32_t kbase = coff[cw0];
        // coff is dead once every worker holds its kbase: its bytes become one 256-entry slot table per warp
        if (nworkers == TILE_THREADS)
            __syncthreads();
        else
            asm volatile("bar.sync 2, %0;" ::"r"(nworkers) : "memory");
        RCB_A(uint16_t) slot_tab = RCB_MK(uint16_t, reinterpret_cast<uint16_t*>(RCB_RAW(coff)) + 256u * wid, 256);  // <= 16 warps x 512 B <= 4 (C + 1) B for C >= 2047
        const bool tab_ok = 512u * nwarps <= 4u * (C + 1u);
        for (uint32_t s0 = soff[oc0 + cw0], s1 = soff[oc0 + cw1]; s0 < s1; s0 += 32) {
            const uint32_t s = s0 + lane;
            uint32_t mask = 0;
            if (s < s1) {
#pragma unroll
                for (int d = 0; d < 8; ++d) mask |= (nid[(uint32_t)d * S2 + s] != (uint16_t)TN ? 1u : 0u) << d;
            }
            const uint32_t cnt = __popc(mask);
            uint32_t incl = cnt;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t v = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= (uint32_t)d) incl += v;
            }
            const uint32_t excl = incl - cnt, T = __shfl_sync(0xffffffffu, incl, 31);
            if (s < s1) o_first_conn[s - nt] = cnt ? krel + kbase + incl - 1 : RCB_NONE;  // the last one pushed
            if (tab_ok) {
                // every lane files its span's slots: owner lane | direction << 5 | first-of-its-block << 8, directions from 7 down
                // (the order the list is pushed in); then the lanes turn to the SLOTS, so the three stores of a round are coalesced
                uint32_t o = excl;
#pragma unroll
                for (int d = 7; d >= 0; --d)
                    if (mask & (1u << d)) {
                        slot_tab[o] = (uint16_t)(lane | ((uint32_t)d << 5) | (o == excl ? 0x100u : 0u));
                        ++o;
                    }
                __syncwarp();
                for (uint32_t t = lane; t < T; t += 32) {
                    const uint32_t e = slot_tab[t];
                    const uint32_t d = (e >> 5) & 7u;
                    const uint32_t k = kbase + t;
                    o_conn_ref[conn_base + k] = ref_off + nid[d * S2 + s0 + (e & 31u)];
                    o_conn_dir[conn_base + k] = (uint8_t)d;
                    o_conn_next[conn_base + k] = (e & 0x100u) ? RCB_NONE : krel + k - 1;
                }
                __syncwarp();
            } else {
                // (tiny tiles: no room for the tables) slot t belongs to the span whose range holds t - bisection over the lanes'
                // offsets - and is its j-th present direction from 7 downwards
                const uint32_t rmask = __brev(mask) >> 24;  // bit i = direction 7 - i: the order the list is pushed in
                for (uint32_t t = lane; t < ((T + 31u) & ~31u); t += 32) {
                    uint32_t lo = 0, hi = 31;
#pragma unroll
                    for (int it = 0; it < 5; ++it) {
                        const uint32_t mid = (lo + hi + 1) >> 1;
                        const uint32_t em = __shfl_sync(0xffffffffu, excl, mid);
                        if (em <= t)
                            lo = mid;
                        else
                            hi = mid - 1;
                    }
                    const uint32_t oex = __shfl_sync(0xffffffffu, excl, lo), orm = __shfl_sync(0xffffffffu, rmask, lo);
                    if (t < T) {
                        const uint32_t j = t - oex;
                        const uint32_t d = 7u - (uint32_t)__fns(orm, 0, (int)j + 1);
                        const uint32_t k = kbase + t;
                        o_conn_ref[conn_base + k] = ref_off + nid[d * S2 + s0 + lo];
                        o_conn_dir[conn_base + k] = (uint8_t)d;
                        o_conn_next[conn_base + k] = j ? krel + k - 1 : RCB_NONE;
                    }
                }
            }
            kbase += T;
        }
    };
    auto no_work = [](uint32_t, uint32_t) {};
    bool conn_written = false;
    // 5. erosion (area.rs:73-234) — before the distance field when both are requested.  mm is dead: its
    //    bytes are the sweep buffers d16a / d16b from here on.  (Whole-tile bands only: the host never cuts a tile when
    //    erosion is asked for.)
    if (a.do_erode) {
        const unsigned thr = (unsigned)(uint8_t)(uint32_t)(a.erode_radius * 2);
        conn_written = tile_sweep<TILE_THREADS, uint8_t, true>(g, RCB_MK(uint8_t, RCB_RAW(d16a), 2ull * G2), RCB_MK(uint8_t, RCB_RAW(d16b), 2ull * G2),
                                                               thr, nullptr, nullptr, conn_work);
    }
    // (Tried: areas - the one result array that is complete in shared memory - as ONE cp.async.bulk per tile, shared -> global,
    // issued by one thread: SASS UBLKCP.G.S, no measurable change, profiles/r02_bulk_store_ab.txt.  It is 2 % of a tile's output;
    // everything else is produced in registers and there is no shared memory left to stage it in.)
    for (uint32_t s = threadIdx.x; s < So; s += TILE_THREADS) o_areas[s] = areas[own0 + s];
    pc.mark(3);
    // 6. distance field (compact_heightfield.rs:514-727)
    if (a.do_dist) {
        if (conn_written)
            tile_sweep<TILE_THREADS, uint16_t, false>(g, d16a, d16b, 0, &sh_misc[2], &pc, no_work);
        else
            conn_written = tile_sweep<TILE_THREADS, uint16_t, false>(g, d16a, d16b, 0, &sh_misc[2], &pc, conn_work);
        pc.mark(4);
        // box blur: d16a holds D; result to global
        for (uint32_t s = own0 + threadIdx.x; s < own1; s += TILE_THREADS) {
            const uint32_t cd = d16a[s];
            uint32_t r = cd;
            if (cd > 2) {
                int d = (int)cd, n = 1;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int dir = 1 + 2 * k;
                    const uint32_t nbr = tnei(g, s, dir);
                    if (nbr == TN) continue;
                    d += d16a[nbr];
                    n++;
                    const uint32_t dg = tnei(g, nbr, (dir + 1) & 7);
                    if (dg != TN) {
                        d += d16a[dg];
                        n++;
                    }
                }
                r = (uint32_t)(d / n);
            }
            o_dist[s - nt] = (uint16_t)r;
        }
        if (threadIdx.x == 0 && sh_misc[2]) atomicMax(&a.tinfo[tile].max_distance, sh_misc[2]);
        pc.mark(5);
    } else {
        for (uint32_t s = threadIdx.x; s < So; s += TILE_THREADS) o_dist[s] = 0;  // compact_heightfield.rs:253
    }
    pc.mark(6);
    if (!conn_written) conn_work(threadIdx.x, TILE_THREADS);  // no sweep ran, or the tile is as wide as the block
    pc.mark(7);
}

}  // namespace

RCB_BOUNDS_EXPORT(rcb_bounds_read_tile)

// proof that the checked accessors bite: one deliberate access one element past an array of eight (debug build only)
__global__ void k_bounds_selftest(uint32_t* p) {
    RCB_A(uint32_t) a = RCB_MK(uint32_t, p, 8);
    a[8 - (threadIdx.x == 1 ? 0 : 1)] = 1u;  // thread 1 writes a[8]
}
void rcb_bounds_selftest(uint32_t* scratch16, cudaStream_t s) {
#ifdef RCB_BOUNDS
    k_bounds_selftest<<<1, 2, 0, s>>>(scratch16);
#else
    (void)scratch16, (void)s;
#endif
}

static void set_smem_attr(const void* fn, size_t bytes) { rcb_raise_dynamic_smem(fn, bytes); }

void launch_tile_spans(const TileSpansArgs& a, int nbands, cudaStream_t s) {
    if (nbands == 0) return;
    KScope _k("k_tile_spans", s);
    if (a.threads == 256) {
        set_smem_attr((const void*)k_tile_spans<256>, a.smem_bytes);
        k_tile_spans<256><<<nbands, 256, a.smem_bytes, s>>>(a);
    } else {
        set_smem_attr((const void*)k_tile_spans<512>, a.smem_bytes);
        k_tile_spans<512><<<nbands, 512, a.smem_bytes, s>>>(a);
    }
}

void launch_tile_segments(const uint32_t* tile_ub, uint32_t* tile_base, int ntiles, unsigned long long* counters,
                          unsigned long long cap_frags, unsigned long long cap_items, cudaStream_t s) {
    KScope _k("k_tile_segments", s);
    k_tile_segments<<<1, 1024, 0, s>>>(tile_ub, tile_base, ntiles, counters, cap_frags, cap_items);
}

void launch_tile_scan(const BandDesc* bands, BandInfo* binfo, int nbands, const TileDesc* tiles, const uint32_t* band_first, TileInfo* tinfo,
                      int ntiles, unsigned long long* counters, unsigned long long span_cap, unsigned long long pub_cap, int compact,
                      const uint32_t* seg_cursor, const uint32_t* seg_cap, cudaStream_t s) {
    KScope _k("k_tile_scan", s);
    k_tile_scan<<<1, 1024, 0, s>>>(bands, binfo, nbands, tiles, band_first, tinfo, ntiles, counters, span_cap, pub_cap, compact, seg_cursor,
                                   seg_cap);
}

size_t tile_compact_smem_bytes(uint32_t own_cells, uint32_t held_cells, uint32_t held_spans, uint32_t top_spans, uint32_t bot_spans) {
    return compact_smem_need(own_cells, held_cells, held_spans, top_spans, bot_spans);
}

void launch_tile_compact(const TileCompactArgs& a, int nbands, cudaStream_t s) {
    if (nbands == 0) return;
    KScope _k("k_tile_compact", s);
    if (a.threads == 256) {
        set_smem_attr((const void*)k_tile_compact<256>, a.smem_bytes);
        k_tile_compact<256><<<nbands, 256, a.smem_bytes, s>>>(a);
    } else {
        set_smem_attr((const void*)k_tile_compact<512>, a.smem_bytes);
        k_tile_compact<512><<<nbands, 512, a.smem_bytes, s>>>(a);
    }
}
