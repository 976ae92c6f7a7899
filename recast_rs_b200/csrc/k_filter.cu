// k_filter.cu — the three walkable filters, fused into one pass over the heightfield CSR.
//
// Replaces (crates/recast/src/heightfield.rs), applied in the order of lib.rs:278-287:
//   :286-328 filter_low_hanging_walkable_obstacles   (F1, sequential bottom-up inside a column)
//   :332-488 filter_ledge_spans                       (F2, reads own area after F1 + neighbour GEOMETRY)
//   :492-531 filter_walkable_low_height_spans         (F3, geometry only)
// F2 never reads a neighbour's area and F1 only looks one span down its own column, so F1 -> F2 -> F3 of a span
// depend on original data only and every span can be filtered on its own.  Pure integer work (5 B read per span
// + 4 neighbour columns, mostly L1/L2 hits; 1 B written).
#include "rcb_common.cuh"

namespace {

constexpr int MAXH = 0xffff;

// One thread per SPAN (columns of a multi-level interior hold 16+ spans; a thread per column leaves the machine mostly
// empty).  F1's carry is local after all: it only passes on the area of a previous span that was
// walkable to begin with, and F1 never changes such a span, so span s needs the ORIGINAL area and max of span s-1 — which
// is why the result goes to a second buffer (area_out) instead of in place.
__global__ void __launch_bounds__(256) k_filter_spans2(const TileDesc* __restrict__ tiles, uint32_t S, const uint2* __restrict__ span_cell,
                                                       const uint32_t* __restrict__ hf_off, const int16_t* __restrict__ smin,
                                                       const int16_t* __restrict__ smax, const uint8_t* __restrict__ sarea,
                                                       uint8_t* __restrict__ area_out, uint32_t fmask) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    const uint2 sc = span_cell[s];
    const bool ordered = span_cell[S].x == 0;  // every column ordered (k_span_cells): bisect instead of walking the column
    const TileDesc& td = tiles[sc.x];
    const uint32_t lc = sc.y;
    const int W = td.width, H = td.height;
    const int z = (int)(lc / (uint32_t)W), x = (int)lc - z * W;
    const uint32_t g = td.cell_base + lc;
    const uint32_t b = hf_off[g], e = hf_off[g + 1];
    const int height = td.walkable_height, climb = td.walkable_climb, neg_climb = td.neg_climb;
    const int floor = smax[s];
    const int ceiling = (s + 1 < e) ? (int)smin[s + 1] : MAXH;
    const uint32_t orig = sarea[s];
    uint32_t a = orig;
    // F1 (:299-319)
    if ((fmask & 1u) && orig == 0 && s > b) {
        const uint32_t pa = sarea[s - 1];
        if (pa != 0 && (floor - (int)smax[s - 1]) <= climb) a = pa;
    }
    // F2 (:351-474)
    if ((fmask & 2u) && a != 0) {
        const int dx[4] = {0, 1, 0, -1}, dz[4] = {-1, 0, 1, 0};  // heightfield.rs:339
        int lowest = MAXH, lo = floor, hi = floor;
        bool ledge = false;
#pragma unroll
        for (int d = 0; d < 4 && !ledge; ++d) {
            const int nx = x + dx[d], nz = z + dz[d];
            if (nx < 0 || nz < 0 || nx >= W || nz >= H) {  // :377-385
                ledge = true;
                break;
            }
            const uint32_t ng = td.cell_base + (uint32_t)nz * W + nx;
            const uint32_t nb = hf_off[ng], ne = hf_off[ng + 1];
            if (nb == ne) {  // :448-453
                ledge = true;
                break;
            }
            int nceil = smin[nb];
            if (min(ceiling, nceil) - floor >= height) {  // :401-406
                ledge = true;
                break;
            }
            // :409-447.  The `break` below only skips work: once it is taken, lowest < neg_climb decides the span.  In an
            // ordered column the gaps that can pass the height test are one run: from the first whose ceiling clears
            // floor + height to the last whose floor stays under ceiling - height.
            uint32_t ns = nb;
            if (ordered && ne - nb > 2) {
                uint32_t top = ne - 1;  // the top gap's ceiling is MAXH
                while (ns < top) {
                    const uint32_t mid = (ns + top) >> 1;
                    if ((int)smin[mid + 1] - floor < height)
                        ns = mid + 1;
                    else
                        top = mid;
                }
            }
            for (; ns < ne; ++ns) {
                const int nfloor = smax[ns];
                if (ordered && ceiling - nfloor < height) break;
                nceil = (ns + 1 < ne) ? (int)smin[ns + 1] : MAXH;
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
    area_out[s] = (uint8_t)a;
}

}  // namespace

void launch_filter_spans2(const TileDesc* tiles, size_t S, const uint2* span_cell, const uint32_t* hf_off, const int16_t* smin,
                          const int16_t* smax, const uint8_t* sarea, uint8_t* area_out, uint32_t fmask, cudaStream_t s) {
    if (S == 0) return;
    KScope _k("k_filter_spans2", s);
    k_filter_spans2<<<(unsigned)((S + 255) / 256), 256, 0, s>>>(tiles, (uint32_t)S, span_cell, hf_off, smin, smax, sarea, area_out, fmask);
}
