// k_filter.cu — the three walkable filters, fused into one pass over the heightfield CSR.
//
// Replaces (crates/recast/src/heightfield.rs), applied in the order of lib.rs:278-287:
//   :286-328 filter_low_hanging_walkable_obstacles   (F1, sequential bottom-up inside a column)
//   :332-488 filter_ledge_spans                       (F2, reads own area after F1 + neighbour GEOMETRY)
//   :492-531 filter_walkable_low_height_spans         (F3, geometry only)
// F2 never reads a neighbour's area and F1 only carries state inside its own column, so one thread
// per column can apply F1 -> F2 -> F3 to each span in a single bottom-up walk and write the final
// area in place: no other thread reads what it writes.  Pure integer work; HBM-bound (5 B read per
// span + 4 neighbour columns, mostly L1/L2 hits; 1 B written).
#include "rcb_common.cuh"

namespace {

constexpr int MAXH = 0xffff;

__global__ void __launch_bounds__(256) k_filter_spans(const TileDesc* __restrict__ tiles, TileMap tm,
                                                      const uint32_t* __restrict__ hf_off,
                                                      const int16_t* __restrict__ smin,
                                                      const int16_t* __restrict__ smax, uint8_t* __restrict__ sarea,
                                                      uint32_t fmask) {
    const uint32_t tile = RCB_TILE_OF_BLOCK();
    if (tile >= tm.ntiles) return;
    const TileDesc& td = tiles[tile];
    const int W = td.width, H = td.height;
    const uint32_t lc = RCB_CELL_OF_THREAD();
    if (lc >= (uint32_t)(W * H)) return;
    const int x = lc % W, z = lc / W;
    const uint32_t g = td.cell_base + lc;
    const uint32_t b = hf_off[g], e = hf_off[g + 1];
    if (b == e) return;
    const int height = td.walkable_height, climb = td.walkable_climb, neg_climb = td.neg_climb;

    // neighbour column ranges in the order (0,-1),(1,0),(0,1),(-1,0) (heightfield.rs:339)
    const int dx[4] = {0, 1, 0, -1}, dz[4] = {-1, 0, 1, 0};
    uint32_t nb[4], ne[4];
    bool oob[4];
#pragma unroll
    for (int d = 0; d < 4; ++d) {
        const int nx = x + dx[d], nz = z + dz[d];
        oob[d] = nx < 0 || nz < 0 || nx >= W || nz >= H;
        if (!oob[d]) {
            const uint32_t ng = td.cell_base + (uint32_t)nz * W + nx;
            nb[d] = hf_off[ng];
            ne[d] = hf_off[ng + 1];
        } else {
            nb[d] = ne[d] = 0;
        }
    }

    bool prev_walkable = false;
    uint32_t prev_area_f1 = 0;
    int prev_max = 0;
    bool have_prev = false;
    for (uint32_t s = b; s < e; ++s) {
        const int floor = smax[s];
        const int ceiling = (s + 1 < e) ? (int)smin[s + 1] : MAXH;
        const uint32_t orig = sarea[s];
        // F1 (:299-319)
        const bool walkable = orig != 0;
        uint32_t a = orig;
        if ((fmask & 1u) && !walkable && prev_walkable && have_prev && (floor - prev_max) <= climb) a = prev_area_f1;
        prev_walkable = walkable;
        prev_area_f1 = a;
        prev_max = floor;
        have_prev = true;
        // F2 (:351-474)
        if ((fmask & 2u) && a != 0) {
            int lowest = MAXH, lo = floor, hi = floor;
            bool ledge = false;
#pragma unroll
            for (int d = 0; d < 4 && !ledge; ++d) {
                if (oob[d] || nb[d] == ne[d]) {  // :377-385, :448-453
                    ledge = true;
                    break;
                }
                int nceil = smin[nb[d]];
                if (min(ceiling, nceil) - floor >= height) {  // :401-406
                    ledge = true;
                    break;
                }
                for (uint32_t ns = nb[d]; ns < ne[d]; ++ns) {  // :409-447
                    const int nfloor = smax[ns];
                    nceil = (ns + 1 < ne[d]) ? (int)smin[ns + 1] : MAXH;
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
            if (ledge || lowest < neg_climb || (hi - lo) > climb) a = 0;
        }
        // F3 (:507-517)
        if ((fmask & 4u) && ceiling - floor < height) a = 0;
        if (a != orig) sarea[s] = (uint8_t)a;
    }
}

}  // namespace

void launch_filter_spans(const TileDesc* tiles, TileMap tm, const uint32_t* hf_off, const int16_t* smin,
                         const int16_t* smax, uint8_t* sarea, uint32_t fmask, cudaStream_t s) {
    if (tm.nblocks == 0) return;
    {
        KScope _k("k_filter_spans", s);
        k_filter_spans<<<tm.grid(), 256, 0, s>>>(tiles, tm, hf_off, smin, smax, sarea, fmask);
    }
}
