// k_regions_impl.cuh — body of k_regions.cu, compiled once per CTA size (RG_BT threads, RG_MIN_CTAS CTAs per SM,
// namespace RG_NS).  See k_regions.cu for the design notes.
namespace RG_NS {


constexpr int BT = RG_BT;
constexpr int MERGE_Q = BT == 32 ? 1 : 4;  // spans per thread and step in the merge passes (the 32-thread build has no registers to spare)
constexpr uint32_t NONE32 = 0xFFFFFFFFu;
constexpr uint32_t BORDER = 0x8000u;

struct BlockSh {
    uint32_t warp[BT / 32 + 1];
    uint32_t misc[4];
};

// exclusive prefix of v over the block; *total = block sum.  Two barriers.
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t* total, BlockSh& sh) {
    const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t n = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= (uint32_t)d) inc += n;
    }
    if (lane == 31) sh.warp[w] = inc;
    __syncthreads();
    if (w == 0) {
        uint32_t t = lane < BT / 32 ? sh.warp[lane] : 0;
#pragma unroll
        for (int d = 1; d < BT / 32; d <<= 1) {
            const uint32_t n = __shfl_up_sync(0xffffffffu, t, d);
            if (lane >= (uint32_t)d) t += n;
        }
        if (lane < BT / 32) sh.warp[lane] = t;  // inclusive warp totals
    }
    __syncthreads();
    const uint32_t before = w ? sh.warp[w - 1] : 0;
    *total = sh.warp[BT / 32 - 1];
    const uint32_t r = before + inc - v;
    __syncthreads();  // sh.warp is reused by the next call
    return r;
}

__device__ __forceinline__ uint32_t block_sum(uint32_t v, BlockSh& sh) {
    uint32_t t;
    block_excl_scan(v, &t, sh);
    return t;
}

__device__ __forceinline__ uint32_t block_min(uint32_t v, BlockSh& sh) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v = min(v, __shfl_down_sync(0xffffffffu, v, d));
    if ((threadIdx.x & 31) == 0) sh.warp[threadIdx.x >> 5] = v;
    __syncthreads();
    uint32_t r = sh.warp[0];
#pragma unroll
    for (int k = 1; k < BT / 32; ++k) r = min(r, sh.warp[k]);
    __syncthreads();
    return r;
}

// ordered compaction of the indices i in [0, n) with pred(i) to emit(position, i); positions start at out_base.
// Source and destination must not alias.  Every thread owns a contiguous run of up to 64 indices (one scan per BT*64 indices instead of one per BT: the
// kernel's time is the number of barrier-separated phases times a memory round trip).
template <class Pred, class Emit>
__device__ uint32_t block_compact(uint32_t n, uint32_t out_base, BlockSh& sh, Pred pred, Emit emit) {
    for (uint32_t c0 = 0; c0 < n; c0 += BT * 64u) {
        const uint32_t m = min(n - c0, (uint32_t)BT * 64u);
        const uint32_t per = (m + BT - 1) / BT;  // <= 64
        const uint32_t lo = c0 + threadIdx.x * per, hi = min(lo + per, c0 + m);
        unsigned long long mask = 0;
        for (uint32_t i = lo; i < hi; ++i) {
            uint32_t payload;
            if (pred(i, payload)) mask |= 1ull << (i - lo);
        }
        uint32_t total;
        uint32_t pos = out_base + block_excl_scan((uint32_t)__popcll(mask), &total, sh);
        while (mask) {
            const uint32_t k = (uint32_t)__ffsll((long long)mask) - 1u;
            mask &= mask - 1;
            uint32_t payload = 0;
            pred(lo + k, payload);
            emit(pos++, payload);
        }
        out_base += total;
    }
    return out_base;
}

struct Tile {
    uint32_t S, C;
    int W, H;
    // every array is the tile's slice of a batch-wide one, behind a checked accessor in the -DRCB_BOUNDS build (rcb_bounds.cuh)
    RCB_A(const uint32_t) cidx;
    RCB_A(const uint32_t) ccnt;
    RCB_A(const uint8_t) areas;
    RCB_A(const uint16_t) con;
    RCB_A(const uint16_t) dist;
    RCB_A(const uint32_t) first_conn;
    RCB_A(const uint32_t) conn_ref;
    RCB_A(const uint8_t) conn_dir;
    RCB_A(const uint32_t) conn_next;
    RCB_A(uint16_t) reg;    // src_reg, also the result
    RCB_A(uint16_t) sdist;  // src_dist
    RCB_A(uint16_t) reg0;   // labels before merge_and_filter (its connection lists are built once, from these)
    RCB_A(uint32_t) nb[4];  // con-based neighbour (tile-relative span) or NONE32 — global-memory mode
    RCB_A(uint16_t) nb16[4];  // the same table with 16-bit entries (0xFFFF = none) when the tile has fewer than 65535 spans: half the
                        // footprint in L1/L2; nb16[0] == nullptr otherwise
    bool walk_lists;  // some cell holds 64+ spans: con == 63 may hide a real neighbour (merge_and_filter follows the list)
    RCB_A(uint32_t) stack[8];
    RCB_A(uint32_t) bufA;
    RCB_A(uint32_t) bufB;
};

__device__ __forceinline__ uint32_t nbr(const Tile& t, int d, uint32_t s) {
    if (RCB_RAW(t.nb16[0])) {
        const uint32_t v = __ldg(&t.nb16[d][s]);
        return v == 0xFFFFu ? NONE32 : v;
    }
    return __ldg(&t.nb[d][s]);
}

// does cell ci stay in region r?  (watershed.rs:167-216) — false when a 4-neighbour, or the neighbour reached from it by
// turning clockwise, of the same area carries another region.
__device__ __forceinline__ bool ws_kept(const Tile& t, uint32_t ci, uint32_t r, uint32_t area) {
    // The verdict is an AND over the four directions, so their order is immaterial: all first-level neighbours are
    // fetched together, then all second-level ones - two table round trips deep instead of up to eight (the reference's
    // early exits only skip work).
    uint32_t a1[4], a2[4], ar1[4], nr1[4];
#pragma unroll
    for (int d = 0; d < 4; ++d) a1[d] = nbr(t, d, ci);
#pragma unroll
    for (int d = 0; d < 4; ++d) {
        const uint32_t i = a1[d] != NONE32 ? a1[d] : ci;
        ar1[d] = t.areas[i];
        nr1[d] = *(volatile uint16_t*)&t.reg[i];
        a2[d] = nbr(t, (d + 1) & 3, i);
    }
    bool bad = false;
#pragma unroll
    for (int d = 0; d < 4; ++d) {
        const bool v1 = a1[d] != NONE32 && ar1[d] == area && !(nr1[d] & BORDER);  // a border neighbour skips its diagonal too
        bad |= v1 && nr1[d] != 0 && nr1[d] != r;
        const uint32_t i2 = a2[d] != NONE32 ? a2[d] : ci;
        const uint32_t ar2 = t.areas[i2];
        const uint32_t nr2 = *(volatile uint16_t*)&t.reg[i2];
        bad |= v1 && a2[d] != NONE32 && ar2 == area && nr2 != 0 && nr2 != r;  // no border test here (:205-210)
    }
    return !bad;
}

__device__ __forceinline__ uint32_t cas16(uint16_t* p, uint16_t cmp, uint16_t val) {
    return (uint32_t)atomicCAS((unsigned short*)p, (unsigned short)cmp, (unsigned short)val);
}

// watershed.rs:245-361.  `len` is the stack length (uniform); returns the new length.
// The stack is t.stack[sidx]; the retain step compacts it into t.bufB and the two buffers trade places.
__device__ uint32_t ws_expand(Tile& t, int max_iter, uint32_t level, int sidx, uint32_t len, bool fill, BlockSh& sh) {
    RCB_A(uint32_t) stack = t.stack[sidx];
    if (fill) {
        len = block_compact(t.S, 0, sh, [&](uint32_t s, uint32_t& pl) { pl = s; return t.dist[s] >= level && t.reg[s] == 0 && t.areas[s] != 0; },
                            [&](uint32_t pos, uint32_t s) { stack[pos] = s; });
    } else {
        for (uint32_t j = threadIdx.x; j < len; j += BT) {
            const uint32_t v = stack[j];
            if (v != NONE32 && t.reg[v] != 0) stack[j] = NONE32;
        }
    }
    __syncthreads();
    int iter = 0;
    while (len > 0) {
        uint32_t failed = 0;
        for (uint32_t j = threadIdx.x; j < len; j += BT) {
            const uint32_t i = stack[j];
            uint32_t res = 0;
            if (i == NONE32) {
                failed++;
            } else {
                uint32_t r = t.reg[i], d2 = 0xffffu;
                const uint32_t area = t.areas[i];
                // the four neighbours, then what hangs off them, fetched together; combined in direction order
                uint32_t an[4], aar[4], ara[4], asd[4];
#pragma unroll
                for (int d = 0; d < 4; ++d) an[d] = nbr(t, d, i);
#pragma unroll
                for (int d = 0; d < 4; ++d) {
                    const uint32_t q = an[d] != NONE32 ? an[d] : i;
                    aar[d] = t.areas[q];
                    ara[d] = t.reg[q];
                    asd[d] = t.sdist[q];
                }
#pragma unroll
                for (int d = 0; d < 4; ++d) {
                    if (an[d] == NONE32 || aar[d] != area) continue;
                    const uint32_t ra = ara[d];
                    const uint32_t da = (uint32_t)(uint16_t)(asd[d] + 2);
                    if (ra > 0 && (ra & BORDER) == 0 && da < d2) {
                        r = ra;
                        d2 = da;
                    }
                }
                if (r != 0)
                    res = r | (d2 << 16);
                else
                    failed++;
            }
            t.bufA[j] = res;
        }
        __syncthreads();  // every gather saw the labels of the previous pass
        for (uint32_t j = threadIdx.x; j < len; j += BT) {
            const uint32_t res = t.bufA[j];
            if (res != 0) {
                const uint32_t i = stack[j];
                t.reg[i] = (uint16_t)(res & 0xffffu);
                t.sdist[i] = (uint16_t)(res >> 16);
                stack[j] = NONE32;
            }
        }
        __syncthreads();
        const uint32_t nfailed = block_sum(failed, sh);
        {
            RCB_A(uint32_t) dst = t.bufB;
            len = block_compact(len, 0, sh, [&](uint32_t j, uint32_t& pl) { pl = stack[j]; return pl != NONE32; },
                                [&](uint32_t pos, uint32_t v) { dst[pos] = v; });
            t.bufB = stack;
            t.stack[sidx] = stack = dst;
        }
        __syncthreads();
        if ((unsigned long long)nfailed * 3ull >= (unsigned long long)len * 2ull) break;  // :350, length AFTER the retain
        if (level > 0) {
            if (++iter >= max_iter) break;
        }
    }
    return len;
}

// get_neighbor_connection (compact_heightfield.rs:732-745, 495-510): the connection list; it agrees with con[] unless
// a neighbour sits at offset 63 of its cell, which takes a cell of 64+ spans
__device__ __forceinline__ uint32_t list_neighbor(const Tile& t, uint32_t s, int d4) {
    const uint32_t viacon = nbr(t, d4, s);
    if (viacon != NONE32 || !t.walk_lists) return viacon;
    const uint32_t d8 = 7u - 2u * (uint32_t)d4;  // W, S, E, N
    uint32_t c = t.first_conn[s];
    while (c != NONE32) {
        if (t.conn_dir[c] == d8) return t.conn_ref[c];
        c = t.conn_next[c];
    }
    return NONE32;
}

struct RegClock {
    unsigned long long* acc;
    long long t, t0;
    __device__ RegClock(unsigned long long* a) : acc(a), t(0), t0(0) {
        if (acc && threadIdx.x == 0) t = t0 = clock64();
    }
    __device__ void finish() {  // slowest tile
        if (acc && threadIdx.x == 0) atomicMax(&acc[8], (unsigned long long)(clock64() - t0));
    }
    __device__ void mark(int id) {
        if (acc && threadIdx.x == 0) {
            const long long n = clock64();
            atomicAdd(&acc[id], (unsigned long long)(n - t));
            t = n;
        }
    }
};

// 14 CTAs per SM (<= 73 registers): the 2025 tiles of config 2 are all resident at once
__global__ void __launch_bounds__(BT, RG_MIN_CTAS) k_regions(RegionArgs a) {
    __shared__ BlockSh sh;
    RegClock pc(a.phase);
    __shared__ uint32_t sh_len[8];
    __shared__ uint32_t sh_next;
    const uint32_t tile = blockIdx.x;
    const TileDesc& td = a.tiles[tile];
    const TileInfo ti = a.tinfo[tile];
    Tile t;
    t.S = ti.span_count;
    t.W = td.width;
    t.H = td.height;
    t.C = (uint32_t)(t.W * t.H);
    const size_t sb = ti.span_base;
    t.cidx = RCB_MK(const uint32_t, a.cell_index + td.cell_base, t.C);
    t.ccnt = RCB_MK(const uint32_t, a.cell_count + td.cell_base, t.C);
    t.areas = RCB_MK(const uint8_t, a.areas + sb, t.S);
    t.con = RCB_MK(const uint16_t, a.con + 4 * sb, 4ull * t.S);
    t.dist = RCB_MK(const uint16_t, a.dist + sb, t.S);
    t.first_conn = RCB_MK(const uint32_t, a.first_conn + sb, t.S);
    t.conn_ref = RCB_MK(const uint32_t, a.conn_ref + ti.conn_base, ti.conn_count);
    t.conn_dir = RCB_MK(const uint8_t, a.conn_dir + ti.conn_base, ti.conn_count);
    t.conn_next = RCB_MK(const uint32_t, a.conn_next + ti.conn_base, ti.conn_count);
    t.reg = RCB_MK(uint16_t, a.reg + sb, t.S);
    t.sdist = RCB_MK(uint16_t, a.src_dist + sb, t.S);
    t.reg0 = RCB_MK(uint16_t, a.reg0 + sb, t.S);
#pragma unroll
    for (int d = 0; d < 4; ++d) t.nb[d] = RCB_MK(uint32_t, a.nb4 + (size_t)d * a.Stot + sb, t.S);
#pragma unroll
    for (int k = 0; k < 8; ++k) t.stack[k] = RCB_MK(uint32_t, a.stacks + (size_t)k * a.Stot + sb, t.S);
    t.bufA = RCB_MK(uint32_t, a.bufA + sb, t.S);
    t.bufB = RCB_MK(uint32_t, a.bufB + sb, t.S);
    const uint32_t S = t.S;
    const int W = t.W, H = t.H;

    // ---- 0. working arrays; neighbours through con[] -----------------------------------------------------
    // Tiles whose hot arrays fit keep them in shared memory for the whole kernel: the algorithm is a long
    // sequence of short dependent phases, and its time is (phases) x (latency of one access).
    extern __shared__ __align__(16) unsigned char rg_smem[];
    RCB_A(uint16_t) const g_reg = t.reg;
    const bool in_smem = a.smem_spans >= S && S < 0xFFFFu;
    const bool small = a.all_small != 0;  // S < 0xFFFF for every tile of the batch
#pragma unroll
    for (int d = 0; d < 4; ++d) t.nb16[d] = RCB_MK(uint16_t, small ? (uint16_t*)a.nb4 + (size_t)d * a.Stot + sb : nullptr, small ? t.S : 0u);
    if (in_smem) {
        // labels and the neighbour table (10 B per span) are what the floods hammer; the rest is read through L1
        t.reg = RCB_MK(uint16_t, rg_smem, S);
    }
    for (uint32_t s = threadIdx.x; s < S; s += BT) t.reg[s] = 0, t.sdist[s] = 0;
    uint32_t maxcnt = 0;
    for (uint32_t c = threadIdx.x; c < t.C; c += BT) {
        const uint32_t b = t.cidx[c];
        if (b == NONE32) continue;
        const uint32_t n = t.ccnt[c];
        maxcnt = max(maxcnt, n);
        const int y = (int)(c / (uint32_t)W), x = (int)c - y * W;
        for (uint32_t s = b; s < b + n; ++s) {
#pragma unroll
            for (int d = 0; d < 4; ++d) {
                const uint32_t k = t.con[4 * (size_t)s + d];
                const int ox = d == 0 ? -1 : (d == 2 ? 1 : 0), oy = d == 1 ? 1 : (d == 3 ? -1 : 0);  // compact_heightfield.rs:748-767
                const uint32_t v = k == 63u ? NONE32 : t.cidx[(uint32_t)((y + oy) * W + (x + ox))] + k;
                if (small)
                    t.nb16[d][s] = (uint16_t)v;  // NONE32 -> 0xFFFF
                else
                    t.nb[d][s] = v;
            }
        }
    }
    t.walk_lists = __syncthreads_or(maxcnt > 63u) != 0;
    pc.mark(0);
    // ---- border regions (:393-426) ----------------------------------------------------------------------------
    uint32_t region_id = 1;
    if (td.aux0 > 0) {
        const int bw = min(W, td.aux0), bh = min(H, td.aux0);
        const int rx0[4] = {0, W - bw, 0, 0}, rx1[4] = {bw, W, W, W}, ry0[4] = {0, 0, 0, H - bh}, ry1[4] = {H, H, bh, H};
        for (int k = 0; k < 4; ++k) {
            for (uint32_t c = threadIdx.x; c < t.C; c += BT) {
                const int y = (int)(c / (uint32_t)W), x = (int)c - y * W;
                if (x < rx0[k] || x >= rx1[k] || y < ry0[k] || y >= ry1[k]) continue;
                const uint32_t b = t.cidx[c];
                if (b == NONE32) continue;
                for (uint32_t s = b; s < b + t.ccnt[c]; ++s)
                    if (t.areas[s] != 0) t.reg[s] = (uint16_t)(region_id | BORDER);
            }
            region_id++;
            __syncthreads();
        }
    }

    // ---- watershed levels (:431-477) ---------------------------------------------------------------------------
    uint32_t level = ((uint32_t)ti.max_distance + 1u) & ~1u;
    int s_id = -1;
    bool overflow = false;
    while (level > 0 && !overflow) {
        level = level >= 2 ? level - 2 : 0;
        s_id = (s_id + 1) & 7;
        uint32_t len;
        if (s_id == 0) {  // sort_cells_by_level (:65-119), log_levels_per_stack = 1
            // all eight stacks in one sweep: every thread classifies a contiguous run of spans (4 bits each, kept in
            // registers), the eight per-thread counts are scanned two to a word, then the run is written out in order
            const uint32_t start = level >> 1;
            uint32_t base8[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) base8[k] = 0;
            constexpr uint32_t RUN = BT >= 1024 ? 32u : 64u;  // BT * RUN < 65536: 16 bits hold a prefix
            for (uint32_t c0 = 0; c0 < S; c0 += BT * RUN) {
                const uint32_t m = min(S - c0, (uint32_t)BT * RUN);
                const uint32_t per = (m + BT - 1) / BT;  // <= RUN
                const uint32_t lo = c0 + threadIdx.x * per, hi = min(lo + per, c0 + m);
                unsigned long long nib[4] = {0, 0, 0, 0};
                uint32_t cnt[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) cnt[k] = 0;
                for (uint32_t i = lo; i < hi; ++i) {
                    uint32_t sid = 15;
                    if (t.areas[i] != 0 && t.reg[i] == 0) {
                        const uint32_t lv = (uint32_t)t.dist[i] >> 1;
                        const uint32_t d = start > lv ? start - lv : 0u;  // saturating_sub (:97)
                        if (d < 8) sid = d;
                    }
                    const uint32_t q = i - lo;
                    nib[q >> 4] |= (unsigned long long)sid << ((q & 15u) * 4u);
#pragma unroll
                    for (int k = 0; k < 8; ++k) cnt[k] += sid == (uint32_t)k;
                }
                uint32_t ex[8];
#pragma unroll
                for (int k = 0; k < 8; k += 2) {  // two 16-bit counts to a word
                    uint32_t total;
                    const uint32_t e = block_excl_scan(cnt[k] | (cnt[k + 1] << 16), &total, sh);
                    ex[k] = base8[k] + (e & 0xffffu);
                    ex[k + 1] = base8[k + 1] + (e >> 16);
                    base8[k] += total & 0xffffu;
                    base8[k + 1] += total >> 16;
                }
                for (uint32_t i = lo; i < hi; ++i) {
                    const uint32_t q = i - lo;
                    const uint32_t sid = (uint32_t)(nib[q >> 4] >> ((q & 15u) * 4u)) & 15u;
#pragma unroll
                    for (int k = 0; k < 8; ++k)
                        if (sid == (uint32_t)k) t.stack[k][ex[k]++] = i;
                }
            }
            if (threadIdx.x < 8) {
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    if (threadIdx.x == (uint32_t)k) sh_len[k] = base8[k];
            }
            __syncthreads();
            len = sh_len[0];
            pc.mark(1);
        } else {  // append_stacks (:122-133)
            RCB_A(uint32_t) src = t.stack[s_id - 1];
            RCB_A(uint32_t) dst = t.stack[s_id];
            len = block_compact(sh_len[s_id - 1], sh_len[s_id], sh,
                                [&](uint32_t j, uint32_t& pl) { pl = src[j]; return pl != NONE32 && t.reg[pl] == 0; },
                                [&](uint32_t pos, uint32_t v) { dst[pos] = v; });
            __syncthreads();
            pc.mark(2);
        }
        len = ws_expand(t, 8, level, s_id, len, false, sh);
        pc.mark(3);
        RCB_A(uint32_t) stack = t.stack[s_id];
        if (threadIdx.x == 0) sh_len[s_id] = len;
        __syncthreads();
        // new regions (:454-476): entries in order; the first one that is unlabelled and survives its own test floods
        const uint32_t lev = level >= 2 ? level - 2 : 0;
        uint32_t cursor = 0;
        for (;;) {
            // the next seed: the first entry at or after the cursor that is still unlabelled and passes its own test.
            // One window of BT entries per step (one entry per thread): entries before the hit are done for good,
            // and nobody walks the whole stack on its own.
            uint32_t jmin = NONE32;
            while (cursor < len) {
                const uint32_t j = cursor + threadIdx.x;
                uint32_t mine = NONE32;
                if (j < len) {
                    const uint32_t v = stack[j];
                    if (t.reg[v] == 0 && ws_kept(t, v, region_id, t.areas[v])) mine = j;
                }
                jmin = block_min(mine, sh);
                if (jmin != NONE32) break;
                cursor += BT;
            }
            pc.mark(4);
            if (jmin == NONE32) break;
            const uint32_t seed = stack[jmin];
            const uint32_t area = t.areas[seed];
            if (threadIdx.x == 0) {
                t.reg[seed] = (uint16_t)region_id;
                t.sdist[seed] = 0;
                t.bufA[0] = seed;
            }
            uint32_t ncur = 1;
            RCB_A(uint32_t) cur = t.bufA;
            RCB_A(uint32_t) nxt = t.bufB;
            while (ncur > 0) {
                if (threadIdx.x == 0) sh_next = 0;
                __syncthreads();
                // four lanes per frontier cell, one per direction: a wave's latency is the instruction chain of one
                // lane, and a lane that handles one neighbour (and its clockwise diagonal) has a quarter of it
                for (uint32_t k0 = 0; k0 < ncur; k0 += BT / 4) {
                    const uint32_t k = k0 + (threadIdx.x >> 2);
                    const int d = (int)(threadIdx.x & 3u);
                    const bool valid = k < ncur;
                    uint32_t ci = 0, ai = NONE32, dai = 0;
                    bool bad = false, same = false;
                    if (valid) {
                        ci = cur[k];
                        ai = nbr(t, d, ci);
                        if (ai != NONE32) {
                            // everything that hangs off ai is fetched together: its area, label, distance (needed after
                            // the group vote) and its clockwise neighbour
                            dai = t.dist[ai];
                            const uint32_t ar = t.areas[ai];
                            const uint32_t nr = *(volatile uint16_t*)&t.reg[ai];
                            const uint32_t ai2 = nbr(t, (d + 1) & 3, ai);
                            if (ar == area) {
                                same = true;
                                if (!(nr & BORDER)) {  // a border neighbour skips its diagonal too (:183-186)
                                    if (nr != 0 && nr != region_id) {
                                        bad = true;
                                    } else if (ai2 != NONE32 && t.areas[ai2] == area) {
                                        const uint32_t nr2 = *(volatile uint16_t*)&t.reg[ai2];
                                        if (nr2 != 0 && nr2 != region_id) bad = true;
                                    }
                                }
                            }
                        }
                    }
                    const unsigned grp = (__ballot_sync(0xffffffffu, bad) >> ((threadIdx.x & 31u) & ~3u)) & 0xfu;
                    if (!valid) continue;
                    if (grp) {
                        if (d == 0) t.reg[ci] = 0;  // :218-221
                        continue;
                    }
                    if (same && dai >= lev && cas16(&t.reg[ai], 0, (uint16_t)region_id) == 0) {
                        t.sdist[ai] = 0;
                        nxt[atomicAdd(&sh_next, 1u)] = ai;
                    }
                }
                __syncthreads();
                ncur = sh_next;
                RCB_A(uint32_t) tmp = cur;
                cur = nxt;
                nxt = tmp;
                __syncthreads();
            }
            pc.mark(5);
            if (region_id == 0xFFFFu) {  // :469-471
                overflow = true;
                break;
            }
            region_id++;
            cursor = jmin + 1;
        }
    }
    if (overflow) {
        if (threadIdx.x == 0) atomicMin(a.err, tile);
        return;
    }
    // ---- final expansion (:480-488) ----------------------------------------------------------------------------
    ws_expand(t, 64, 0, 0, 0, true, sh);
    pc.mark(6);
    if (threadIdx.x == 0) a.tinfo[tile].max_regions = region_id;  // :491, before filtering

    // ---- merge_and_filter_regions (:543-633) ----------------------------------------------------------------
    const uint32_t R = region_id;
    const size_t rb = sb + 2 * (size_t)tile;  // R <= S + 1 entries per tile
    // per-region tables: the tile owns span_count + 2 entries of each
    RCB_A(int32_t) cnt = RCB_MK(int32_t, a.r_cnt + rb, S + 2);
    RCB_A(int32_t) cadd = RCB_MK(int32_t, a.r_cadd + rb, S + 2);
    RCB_A(uint16_t) alive = RCB_MK(uint16_t, a.r_alive + rb, S + 2);
    RCB_A(uint8_t) border = RCB_MK(uint8_t, a.r_border + rb, S + 2);
    RCB_A(unsigned long long) best = RCB_MK(unsigned long long, a.r_best + rb, S + 2);
    RCB_A(uint16_t) target = RCB_MK(uint16_t, a.r_target + rb, S + 2);
    RCB_A(uint16_t) fin = RCB_MK(uint16_t, a.r_fin + rb, S + 2);
    const int32_t min_area = td.aux1, merge_area = td.aux2;
    for (uint32_t i = threadIdx.x; i < R; i += BT) cnt[i] = 0, alive[i] = (uint16_t)i, border[i] = 0;
    for (uint32_t s = threadIdx.x; s < S; s += BT) t.reg0[s] = t.reg[s];
    __syncthreads();
    // the span passes below fetch a span's four neighbours, then their four labels, together (a chain two loads deep
    // however many directions matter)
    for (uint32_t s = threadIdx.x; s < S; s += BT) {
        const uint32_t rid = t.reg0[s];
        uint32_t nn[4], nr[4];
#pragma unroll
        for (int d = 0; d < 4; ++d) nn[d] = list_neighbor(t, s, d);
#pragma unroll
        for (int d = 0; d < 4; ++d) nr[d] = t.reg0[nn[d] != NONE32 ? nn[d] : s];
        if (rid == 0 || (rid & BORDER)) continue;
        atomicAdd(&cnt[rid], 1);
        bool at_border = false;
#pragma unroll
        for (int d = 0; d < 4; ++d) at_border |= nn[d] != NONE32 && (nr[d] & BORDER);
        if (at_border) border[rid] = 1;
    }
    __syncthreads();
    // small regions (:598-622)
    for (uint32_t i = threadIdx.x; i < R; i += BT) {
        uint16_t f = (uint16_t)i;
        if (alive[i] != 0 && cnt[i] != 0 && cnt[i] < min_area && !border[i]) {
            alive[i] = 0;
            f = 0;
        }
        fin[i] = f;
    }
    __syncthreads();
    for (uint32_t s = threadIdx.x; s < S; s += BT) {
        const uint32_t v = t.reg[s];
        if (v < R) t.reg[s] = fin[v];
    }
    __syncthreads();
    // merges (:636-716): a round's decisions are taken from the state before the round
    if (merge_area > 0) {
        for (;;) {
            uint32_t ncand = 0;  // regions that would merge if they find a neighbour: none -> no span pass needed
            for (uint32_t i = threadIdx.x; i < R; i += BT) {
                best[i] = 0ull, target[i] = 0, cadd[i] = 0;
                ncand += alive[i] != 0 && cnt[i] != 0 && cnt[i] < merge_area && !border[i];
            }
            if (block_sum(ncand, sh) == 0) break;
            for (int pass = 0; pass < 2; ++pass) {
                // MERGE_Q spans per thread and step: nearly all of them only find out that their region is no candidate, and
                // one span at a time that was two dependent loads per span, S / BT times in a row
                for (uint32_t s0 = threadIdx.x; s0 < S; s0 += MERGE_Q * BT) {
                    uint32_t rid4[MERGE_Q];
                    bool cand[MERGE_Q];
#pragma unroll
                    for (int q = 0; q < MERGE_Q; ++q) {
                        const uint32_t sq = s0 + (uint32_t)q * BT;
                        rid4[q] = sq < S ? (uint32_t)t.reg0[sq] : 0u;
                    }
#pragma unroll
                    for (int q = 0; q < MERGE_Q; ++q) {
                        const bool lab = rid4[q] != 0 && !(rid4[q] & BORDER);
                        const uint32_t ri = lab ? rid4[q] : 0u;  // entry 0 is always there
                        const uint32_t al = alive[ri], bo = border[ri];
                        const int32_t cn = cnt[ri];
                        cand[q] = lab && al != 0 && cn != 0 && cn < merge_area && !bo;
                    }
#pragma unroll
                    for (int q = 0; q < MERGE_Q; ++q) {
                        if (!cand[q]) continue;
                        const uint32_t s = s0 + (uint32_t)q * BT, rid = rid4[q];
                        uint32_t nn[4], nrs[4];
#pragma unroll
                        for (int d = 0; d < 4; ++d) nn[d] = list_neighbor(t, s, d);
#pragma unroll
                        for (int d = 0; d < 4; ++d) nrs[d] = t.reg0[nn[d] != NONE32 ? nn[d] : s];
#pragma unroll
                        for (int d = 0; d < 4; ++d) {
                            if (nn[d] == NONE32) continue;
                            const uint32_t nr = nrs[d];
                            if (nr == rid || nr == 0 || (nr & BORDER) || nr >= R) continue;
                            if (alive[nr] == 0 || cnt[nr] <= 0) continue;
                            // first neighbour in list order with the largest span count == max of (count, -first occurrence)
                            const unsigned long long key = ((unsigned long long)(uint32_t)cnt[nr] << 32) | (uint32_t)~(s * 4u + (uint32_t)d);
                            if (pass == 0)
                                atomicMax(&best[rid], key);
                            else if (key == best[rid])
                                target[rid] = (uint16_t)nr;
                        }
                    }
                }
                __syncthreads();
            }
            uint32_t nm = 0;
            for (uint32_t i = threadIdx.x; i < R; i += BT) nm += target[i] != 0;
            if (block_sum(nm, sh) == 0) break;
            // applied in index order: a label follows the merges that come after the one that produced it
            for (uint32_t v = threadIdx.x; v < R; v += BT) {
                uint32_t cur = v;
                int last = -1;
                while (target[cur] != 0 && (int)cur > last) {
                    last = (int)cur;
                    cur = target[cur];
                }
                fin[v] = (uint16_t)cur;
                const uint32_t b = target[v];
                if (b != 0 && !(target[b] != 0 && b < v)) atomicAdd(&cadd[b], cnt[v]);  // the target is still a region (:700-706)
            }
            __syncthreads();
            for (uint32_t i = threadIdx.x; i < R; i += BT) {
                if (target[i] != 0) alive[i] = 0;
                cnt[i] += cadd[i];
            }
            for (uint32_t s = threadIdx.x; s < S; s += BT) {
                const uint32_t v = t.reg[s];
                if (v < R) t.reg[s] = fin[v];
            }
            __syncthreads();
        }
    }
    // compact ids (:719-746)
    {
        uint32_t base = 1;
        for (uint32_t c0 = 0; c0 < R; c0 += BT) {
            const uint32_t i = c0 + threadIdx.x;
            const bool live = i < R && alive[i] != 0;
            uint32_t total;
            const uint32_t ex = block_excl_scan(live ? 1u : 0u, &total, sh);
            if (i < R) fin[i] = live ? (uint16_t)(base + ex) : (uint16_t)0;
            base += total;
        }
        __syncthreads();
        for (uint32_t s = threadIdx.x; s < S; s += BT) {
            const uint32_t v = t.reg[s];
            g_reg[s] = v < R ? fin[v] : (uint16_t)v;
        }
    }
    pc.mark(7);
    pc.finish();
}

}  // namespace RG_NS
