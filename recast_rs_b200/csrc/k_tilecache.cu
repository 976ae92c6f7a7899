// k_tilecache.cu — tile-cache layers -> heightfield spans, and obstacle marking on the compact heightfield.
//
// Replaces (crates/detour-tilecache/src/tile_cache_builder.rs):
//   :129-172 decode_layer_to_heightfield   one span (hmin, hmin+1, areas[idx]) per cell with regons[idx] != 0
//   :175-251 rasterize_obstacles            state / XZ-AABB rejection, dispatch on the obstacle type
//   :254-320 rasterize_cylinder_obstacle    :422-467 rasterize_box_obstacle    :470-549 rasterize_oriented_box_obstacle
// Obstacles only ever write 0 into CompactSpan.area, so any order (and any parallel schedule) gives the
// reference's result.  Float expressions keep the reference's operand order with *_rn intrinsics.
#include "rcb_common.cuh"

namespace {

// per cell: 1 if the layer cell is non-empty
__global__ void __launch_bounds__(256) k_tc_count(const TileDesc* __restrict__ tiles, TileMap tm, const uint8_t* __restrict__ regons,
                                                  uint32_t* __restrict__ cell_cnt) {
    const uint32_t tile = RCB_TILE_OF_BLOCK();
    if (tile >= tm.ntiles) return;
    const TileDesc& td = tiles[tile];
    const uint32_t lc = RCB_CELL_OF_THREAD();
    if (lc >= (uint32_t)(td.width * td.height)) return;
    const uint32_t g = td.cell_base + lc;
    cell_cnt[g] = regons[g] != 0 ? 1u : 0u;
}

__global__ void __launch_bounds__(256) k_tc_spans(const TileDesc* __restrict__ tiles, TileMap tm, const uint8_t* __restrict__ regons,
                                                  const uint8_t* __restrict__ areas, const uint32_t* __restrict__ hf_off,
                                                  int16_t* __restrict__ smin, int16_t* __restrict__ smax, uint8_t* __restrict__ sarea) {
    const uint32_t tile = RCB_TILE_OF_BLOCK();
    if (tile >= tm.ntiles) return;
    const TileDesc& td = tiles[tile];
    const uint32_t lc = RCB_CELL_OF_THREAD();
    if (lc >= (uint32_t)(td.width * td.height)) return;
    const uint32_t g = td.cell_base + lc;
    if (regons[g] == 0) return;
    const uint32_t s = hf_off[g];
    const int32_t h = td.aux0;               // header.hmin as i32 (:158)
    smin[s] = (int16_t)(uint16_t)(uint32_t)h;        // `as i16` wraps (:164)
    smax[s] = (int16_t)(uint16_t)(uint32_t)(h + 1);  // :165
    sarea[s] = areas[g];
}

__device__ __forceinline__ int f2i_floor_div(float num, float cs) { return __float2int_rz(floorf(__fdiv_rn(num, cs))); }
__device__ __forceinline__ int f2i_ceil_div(float num, float cs) { return __float2int_rz(ceilf(__fdiv_rn(num, cs))); }

// One CTA per layer; obstacles of the layer in sequence, cells of the footprint in parallel.
__global__ void __launch_bounds__(256) k_tc_obstacles(const TileDesc* __restrict__ tiles, const rcb_tc_obstacle* __restrict__ obs,
                                                      const uint32_t* __restrict__ hf_off, const int16_t* __restrict__ smin,
                                                      const int16_t* __restrict__ smax, uint8_t* __restrict__ sarea) {
    const TileDesc& td = tiles[blockIdx.x];
    const int W = td.width, H = td.height;
    const float cs = td.cs, ch = td.ch;
    const float bx = td.bmin[0], by = td.bmin[1], bz = td.bmin[2];
    for (int oi = 0; oi < td.aux2; ++oi) {
        const rcb_tc_obstacle o = obs[td.aux1 + oi];
        if (o.state == RCB_OBSTACLE_EMPTY || o.state == RCB_OBSTACLE_REMOVING) continue;  // :183-185
        float omin_x, omax_x, omin_z, omax_z, radius = 0.f;
        if (o.type == RCB_OBSTACLE_CYLINDER) {
            omin_x = __fsub_rn(o.a[0], o.b[0]), omax_x = __fadd_rn(o.a[0], o.b[0]);
            omin_z = __fsub_rn(o.a[2], o.b[0]), omax_z = __fadd_rn(o.a[2], o.b[0]);
        } else if (o.type == RCB_OBSTACLE_BOX) {
            omin_x = o.a[0], omax_x = o.b[0], omin_z = o.a[2], omax_z = o.b[2];
        } else {
            radius = __fsqrt_rn(__fadd_rn(__fmul_rn(o.b[0], o.b[0]), __fmul_rn(o.b[2], o.b[2])));  // :212
            omin_x = __fsub_rn(o.a[0], radius), omax_x = __fadd_rn(o.a[0], radius);
            omin_z = __fsub_rn(o.a[2], radius), omax_z = __fadd_rn(o.a[2], radius);
        }
        if (omax_x < td.bmin[0] || omin_x > td.bmax[0] || omax_z < td.bmin[2] || omin_z > td.bmax[2]) continue;  // :228-234
        int min_x, max_x, min_z, max_z;
        if (o.type == RCB_OBSTACLE_BOX) {  // :433-436
            min_x = f2i_floor_div(__fsub_rn(o.a[0], bx), cs), max_x = f2i_ceil_div(__fsub_rn(o.b[0], bx), cs);
            min_z = f2i_floor_div(__fsub_rn(o.a[2], bz), cs), max_z = f2i_ceil_div(__fsub_rn(o.b[2], bz), cs);
        } else {  // :269-272, :494-497  (c - r - bmin) / cs
            const float r = o.type == RCB_OBSTACLE_CYLINDER ? o.b[0] : radius;
            min_x = f2i_floor_div(__fsub_rn(__fsub_rn(o.a[0], r), bx), cs), max_x = f2i_ceil_div(__fsub_rn(__fadd_rn(o.a[0], r), bx), cs);
            min_z = f2i_floor_div(__fsub_rn(__fsub_rn(o.a[2], r), bz), cs), max_z = f2i_ceil_div(__fsub_rn(__fadd_rn(o.a[2], r), bz), cs);
        }
        min_x = max(min_x, 0), max_x = min(max_x, W - 1), min_z = max(min_z, 0), max_z = min(max_z, H - 1);
        if (max_x < min_x || max_z < min_z) continue;
        const int fw = max_x - min_x + 1, fn = fw * (max_z - min_z + 1);
        for (int i = threadIdx.x; i < fn; i += blockDim.x) {
            const int x = min_x + i % fw, z = min_z + i / fw;
            bool inside;
            float y_lo, y_hi;  // the obstacle's height interval
            if (o.type == RCB_OBSTACLE_CYLINDER) {  // :284-292
                const float cell_x = __fadd_rn(bx, __fmul_rn(__fadd_rn((float)x, 0.5f), cs));
                const float cell_z = __fadd_rn(bz, __fmul_rn(__fadd_rn((float)z, 0.5f), cs));
                const float dx = __fsub_rn(cell_x, o.a[0]), dz = __fsub_rn(cell_z, o.a[2]);
                inside = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dz, dz)) <= __fmul_rn(o.b[0], o.b[0]);
                y_lo = o.a[1];
                y_hi = __fadd_rn(o.a[1], o.b[1]);
            } else if (o.type == RCB_OBSTACLE_BOX) {
                inside = true;
                y_lo = o.a[1];
                y_hi = o.b[1];
            } else {  // :509-525
                const float cell_x = __fadd_rn(__fadd_rn(__fmul_rn((float)x, cs), bx), __fmul_rn(cs, 0.5f));
                const float cell_z = __fadd_rn(__fadd_rn(__fmul_rn((float)z, cs), bz), __fmul_rn(cs, 0.5f));
                const float dx = __fsub_rn(cell_x, o.a[0]), dz = __fsub_rn(cell_z, o.a[2]);
                const float cos_a = __fsub_rn(__fmul_rn(2.0f, __fadd_rn(o.rot_aux[1], 0.5f)), 1.0f);
                const float sin_a = __fmul_rn(2.0f, o.rot_aux[0]);
                const float lx = __fadd_rn(__fmul_rn(cos_a, dx), __fmul_rn(sin_a, dz));
                const float lz = __fadd_rn(__fmul_rn(-sin_a, dx), __fmul_rn(cos_a, dz));
                inside = fabsf(lx) <= o.b[0] && fabsf(lz) <= o.b[2];
                y_lo = __fsub_rn(o.a[1], o.b[1]);
                y_hi = __fadd_rn(o.a[1], o.b[1]);
            }
            if (!inside) continue;
            const uint32_t g = td.cell_base + (uint32_t)z * W + x;
            for (uint32_t s = hf_off[g]; s < hf_off[g + 1]; ++s) {
                float y0, y1;
                if (o.type == RCB_OBSTACLE_CYLINDER) {  // :304-305: span.y (== min) and one cell of height
                    y0 = __fadd_rn(by, __fmul_rn((float)(int)smin[s], ch));
                    y1 = __fadd_rn(y0, ch);
                } else {  // :454-455, :533-534
                    y0 = __fadd_rn(__fmul_rn((float)smin[s], ch), by);
                    y1 = __fadd_rn(__fmul_rn((float)smax[s], ch), by);
                }
                if (o.type == RCB_OBSTACLE_CYLINDER ? (y0 < y_hi && y1 > y_lo) : (y1 > y_lo && y0 < y_hi)) sarea[s] = 0;
            }
        }
    }
}

}  // namespace

void launch_tc_count(const TileDesc* tiles, TileMap tm, const uint8_t* regons, uint32_t* cell_cnt, cudaStream_t s) {
    if (tm.nblocks == 0) return;
    KScope _k("k_tc_count", s);
    k_tc_count<<<tm.grid(), 256, 0, s>>>(tiles, tm, regons, cell_cnt);
}

void launch_tc_spans(const TileDesc* tiles, TileMap tm, const uint8_t* regons, const uint8_t* areas, const uint32_t* hf_off,
                     int16_t* smin, int16_t* smax, uint8_t* sarea, cudaStream_t s) {
    if (tm.nblocks == 0) return;
    KScope _k("k_tc_spans", s);
    k_tc_spans<<<tm.grid(), 256, 0, s>>>(tiles, tm, regons, areas, hf_off, smin, smax, sarea);
}

void launch_tc_obstacles(const TileDesc* tiles, int ntiles, const rcb_tc_obstacle* obs, const uint32_t* hf_off, const int16_t* smin,
                         const int16_t* smax, uint8_t* sarea, cudaStream_t s) {
    if (ntiles == 0) return;
    KScope _k("k_tc_obstacles", s);
    k_tc_obstacles<<<ntiles, 256, 0, s>>>(tiles, obs, hf_off, smin, smax, sarea);
}
