// k_dist.cu — distance field and walkable-area erosion on the compact heightfield.
//
// Replaces:
//   compact_heightfield.rs:514-543 build_distance_field, :546-675 calculate_distance_field,
//   :679-727 box_blur;  area.rs:73-234 erode_walkable_area.
//
// The reference's "two-pass chamfer" indexes the 8-direction connection table with 4-direction
// numbers.  Reading the loops literally:
//   forward  (z up,   x up)  : F[i] = min(init[i], F[NW]+2, F[NW.E]+3, init[E]+2, F[E.NE]+3)
//       NW, NW.E and E.NE lie in row z-1 (already final), E lies in row z and has not been visited
//       yet (still its init value)  =>  rows are sequential, spans inside a row are independent.
//   backward (z down, x down): D[i] = min(F[i], F[NE]+2, F[NE.N]+3, F[N]+2, F[N.NW]+3)
//       NE, N are in row z-1 and NE.N, N.NW in row z-2, all visited LATER by the descending sweep
//       =>  they still hold forward values: the backward pass is a pure gather.
// (x.y means "neighbour y of neighbour x"; all additions saturate in u16 / u8.)
// tests/test_oracle_* pin this equivalence against the literal sequential sweeps of the oracle.
//
// Kernels: init (boundary marking) -> gather ids + static part of the relaxation -> forward (one CTA per tile, H
// sequential row steps with __syncthreads) -> backward gather [+ max | + erosion threshold] -> blur.  All but the forward
// sweep are one thread per span over the absolute neighbour ids k_link_spans2 leaves behind.
#include "rcb_common.cuh"

namespace {

// neighbour `dir` of span `s`: global span index or RCB_NONE.  k_link_spans2 leaves the 8-direction table as absolute
// span ids (u32[8][S]), so nothing below needs to know which cell or tile a span belongs to: one thread per span.
struct Grid {
    const uint32_t* nid;
    size_t S;
};
__device__ __forceinline__ uint32_t nei(const Grid& g, uint32_t s, int dir) { return g.nid[(size_t)dir * g.S + s]; }

template <typename T>
__device__ __forceinline__ T sat_add(T a, unsigned b) {
    const unsigned r = (unsigned)a + b;
    const unsigned mx = (T)~(T)0;
    return (T)(r > mx ? mx : r);
}

// ---- boundary marking -----------------------------------------------------------------------------
// DIST : every span with fewer than 4 cardinal neighbours of EQUAL areas[] -> 0 (compact_heightfield.rs:556-585)
// ERODE: areas==0 -> 0; else 0 unless all of W,S,E,N exist with areas != 0        (area.rs:91-125)
template <typename T, bool ERODE>
__global__ void __launch_bounds__(256) k_boundary(const uint32_t* __restrict__ nid, size_t S, const uint8_t* __restrict__ areas,
                                                  T* __restrict__ init) {
    const size_t s = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    const Grid g{nid, S};
    const T inf = (T)~(T)0;
    const uint8_t a = areas[s];
    int cnt = 0;
    if (!ERODE || a != 0) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t n = nei(g, (uint32_t)s, 7 - 2 * k);  // 7,5,3,1
            if (n == RCB_NONE) continue;
            if (ERODE ? (areas[n] != 0) : (areas[n] == a)) cnt++;
        }
    }
    init[s] = (cnt == 4) ? inf : (T)0;
}

// ---- forward sweep ----------------------------------------------------------------------------------------
// The neighbour ids a row needs (NW, NW.E, E, E.NE) do not depend on distances: k_fwd_index gathers them
// once for every span, in parallel, as global span ids.  A missing neighbour becomes the dummy span S
// (value inf), which makes the sweep body branch-free, and "+2 / +3" needs no clamp because the running
// minimum never exceeds inf.  k_forward2 then walks the rows of a tile with one CTA: per row one 16-byte
// index load (prefetched one row ahead) and the distance loads.
template <typename T>
__global__ void __launch_bounds__(256) k_fwd_index(const uint32_t* __restrict__ nid, size_t S, const T* __restrict__ init,
                                                   uint4* __restrict__ fidx) {
    const size_t s = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    const Grid g{nid, S};
    const uint32_t dummy = (uint32_t)S;
    uint32_t a0 = nei(g, (uint32_t)s, 0), a1 = dummy, b0 = nei(g, (uint32_t)s, 3), b1 = dummy;
    if (a0 != RCB_NONE) {
        const uint32_t t = nei(g, a0, 3);
        if (t != RCB_NONE) a1 = t;
    } else
        a0 = dummy;
    // the part of the relaxation that does not depend on the sweep: the span's own initial value and its E neighbour's
    uint32_t base = init[s];
    if (b0 != RCB_NONE) {
        base = min(base, (uint32_t)init[b0] + 2u);
        const uint32_t t = nei(g, b0, 2);
        if (t != RCB_NONE) b1 = t;
    }
    fidx[s] = make_uint4(a0, a1, base, b1);
}

template <typename T>
__global__ void __launch_bounds__(1024) k_forward2(const TileDesc* __restrict__ tiles, const uint32_t* __restrict__ hf_off,
                                                   const uint4* __restrict__ fidx, size_t S, T* F) {
    const TileDesc& td = tiles[blockIdx.x];
    const int W = td.width, H = td.height;
    const T inf = (T)~(T)0;
    if (threadIdx.x == 0) F[S] = inf;  // dummy span; the same value from every CTA: benign
    __syncthreads();
    auto relax = [&](uint32_t, const uint4 ix) {
        uint32_t v = ix.z;  // min(init[s], init[E] + 2), folded in by k_fwd_index
        v = min(v, (uint32_t)F[ix.x] + 2u);
        v = min(v, (uint32_t)F[ix.y] + 3u);
        v = min(v, (uint32_t)F[ix.w] + 3u);
        return (T)v;
    };
    const uint32_t* __restrict__ off = hf_off + td.cell_base;
    if (W <= (int)blockDim.x) {
        // one column per thread; the next row's cell range and first index vector are fetched a row ahead
        const int x = threadIdx.x;
        const bool col = x < W;
        uint32_t sb = 0, se = 0, nb = 0, ne = 0;
        if (col) {
            sb = off[x];
            se = off[x + 1];
            if (H > 1) {
                nb = off[W + x];
                ne = off[W + x + 1];
            }
        }
        uint4 ix = make_uint4(0, 0, 0, 0);
        if (sb < se) ix = fidx[sb];
        for (int z = 0; z < H; ++z) {
            uint4 nix = make_uint4(0, 0, 0, 0);
            if (nb < ne) nix = fidx[nb];  // row z+1, independent of this row's result
            uint32_t nnb = 0, nne = 0;
            if (col && z + 2 < H) {
                nnb = off[(z + 2) * W + x];
                nne = off[(z + 2) * W + x + 1];
            }
            if (sb < se) {
                F[sb] = relax(sb, ix);
                for (uint32_t s = sb + 1; s < se; ++s) F[s] = relax(s, fidx[s]);
            }
            __syncthreads();
            sb = nb, se = ne, ix = nix, nb = nnb, ne = nne;
        }
    } else {
        for (int z = 0; z < H; ++z) {
            for (int x = threadIdx.x; x < W; x += blockDim.x)
                for (uint32_t s = off[z * W + x]; s < off[z * W + x + 1]; ++s) F[s] = relax(s, fidx[s]);
            __syncthreads();
        }
    }
}

// ---- backward gather ---------------------------------------------------------------------------------
// DIST : D = gather, per-tile max (before blur) -> tinfo.max_distance
// ERODE: D < thr  =>  areas = 0 (area.rs:226-231)
template <typename T, bool ERODE>
__global__ void __launch_bounds__(256) k_backward(const uint32_t* __restrict__ nid, size_t S, const uint2* __restrict__ span_cell,
                                                  const T* __restrict__ F, T* __restrict__ D, uint8_t* areas, unsigned thr,
                                                  TileInfo* __restrict__ tinfo) {
    const size_t s = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    const Grid g{nid, S};
    T v = F[s];
    const uint32_t a = nei(g, (uint32_t)s, 2);  // NE
    if (a != RCB_NONE) {
        v = min(v, sat_add<T>(F[a], 2));
        const uint32_t aa = nei(g, a, 1);  // NE.N
        if (aa != RCB_NONE) v = min(v, sat_add<T>(F[aa], 3));
    }
    const uint32_t bn = nei(g, (uint32_t)s, 1);  // N
    if (bn != RCB_NONE) {
        v = min(v, sat_add<T>(F[bn], 2));
        const uint32_t bb = nei(g, bn, 0);  // N.NW
        if (bb != RCB_NONE) v = min(v, sat_add<T>(F[bb], 3));
    }
    if (ERODE) {
        if ((unsigned)v < thr) areas[s] = 0;
    } else {
        D[s] = v;
        // per-tile maximum (before the blur): one atomic per (warp, tile)
        const uint32_t tile = span_cell[s].x;
        const unsigned peers = __match_any_sync(__activemask(), tile);
        const uint32_t vmax = __reduce_max_sync(peers, (uint32_t)v);
        if ((int)lane_id() == __ffs(peers) - 1 && vmax) atomicMax(&tinfo[tile].max_distance, vmax);
    }
}

// ---- box blur (compact_heightfield.rs:679-727, threshold 1 -> thr 2) ----------------------------------
__global__ void __launch_bounds__(256) k_blur(const uint32_t* __restrict__ nid, size_t S, const uint16_t* __restrict__ D,
                                              uint16_t* __restrict__ dist) {
    const size_t s = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    const Grid g{nid, S};
    const uint32_t cd = D[s];
    if (cd <= 2) {
        dist[s] = (uint16_t)cd;
        return;
    }
    int d = (int)cd, n = 1;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int dir = 1 + 2 * k;  // 1,3,5,7
        const uint32_t nb = nei(g, (uint32_t)s, dir);
        if (nb == RCB_NONE) continue;
        d += D[nb];
        n++;
        const uint32_t dg = nei(g, nb, (dir + 1) & 7);
        if (dg != RCB_NONE) {
            d += D[dg];
            n++;
        }
    }
    dist[s] = (uint16_t)(d / n);
}

}  // namespace

// one thread per column of the widest tile (<= 1024; wider tiles stride): a 1024-thread CTA on a 140-cell-wide tile is
// 86 % idle threads that still hold their SM's warp slots
static int forward_threads(int max_width) {
    const int t = (max_width + 31) & ~31;
    return t < 64 ? 64 : (t > 1024 ? 1024 : t);
}

void launch_distance_field(const TileDesc* tiles, int ntiles, const uint32_t* hf_off, const uint32_t* nid, const uint2* span_cell, size_t S,
                           const uint8_t* areas, uint16_t* tmp_init, uint16_t* tmp_f, uint16_t* tmp_d, uint16_t* dist, TileInfo* tinfo,
                           uint4* fidx, int max_width, cudaStream_t s) {
    if (S == 0 || ntiles == 0) return;
    const unsigned nb = (unsigned)((S + 255) / 256);
    {
        KScope _k("k_boundary", s);
        k_boundary<uint16_t, false><<<nb, 256, 0, s>>>(nid, S, areas, tmp_init);
    }
    {
        KScope _k("k_fwd_index", s);
        k_fwd_index<uint16_t><<<nb, 256, 0, s>>>(nid, S, tmp_init, fidx);
    }
    {
        KScope _k("k_forward2", s);
        k_forward2<uint16_t><<<ntiles, forward_threads(max_width), 0, s>>>(tiles, hf_off, fidx, S, tmp_f);
    }
    {
        KScope _k("k_backward", s);
        k_backward<uint16_t, false><<<nb, 256, 0, s>>>(nid, S, span_cell, tmp_f, tmp_d, nullptr, 0, tinfo);
    }
    {
        KScope _k("k_blur", s);
        k_blur<<<nb, 256, 0, s>>>(nid, S, tmp_d, dist);
    }
}

void launch_erode(const TileDesc* tiles, int ntiles, const uint32_t* hf_off, const uint32_t* nid, size_t S, uint8_t* areas,
                  uint8_t* tmp_init, uint8_t* tmp_f, int radius, uint4* fidx, int max_width, cudaStream_t s) {
    if (S == 0 || ntiles == 0) return;
    const unsigned nb = (unsigned)((S + 255) / 256);
    const unsigned thr = (unsigned)(uint8_t)(uint32_t)(radius * 2);  // area.rs:226 `(radius*2) as u8`
    {
        KScope _k("k_boundary", s);
        k_boundary<uint8_t, true><<<nb, 256, 0, s>>>(nid, S, areas, tmp_init);
    }
    {
        KScope _k("k_fwd_index", s);
        k_fwd_index<uint8_t><<<nb, 256, 0, s>>>(nid, S, tmp_init, fidx);
    }
    {
        KScope _k("k_forward2", s);
        k_forward2<uint8_t><<<ntiles, forward_threads(max_width), 0, s>>>(tiles, hf_off, fidx, S, tmp_f);
    }
    {
        KScope _k("k_backward", s);
        k_backward<uint8_t, true><<<nb, 256, 0, s>>>(nid, S, nullptr, tmp_f, nullptr, areas, thr, nullptr);
    }
}
