// k_area.cu — area operators on CompactHeightfield.areas.
//
// Replaces (crates/recast/src/area.rs): median_filter_walkable_area :238-294, mark_box_area :298-355,
// mark_convex_poly_area :359-437 (+ point_in_poly :29-50), mark_cylinder_area :544-618.
// The mark operators touch each span independently (a span's new area depends on its own area and
// geometry only), so a run of consecutive marks is ONE pass in which every span replays the run in
// order; the median filter gathers neighbour areas through the connection table into a second buffer.
#include "rcb_common.cuh"

namespace {

__device__ __forceinline__ int trunc_div(float v, float origin, float cell) {  // ((v - origin) / cell) as i32
    return __float2int_rz(__fdiv_rn(__fsub_rn(v, origin), cell));
}

__device__ bool point_in_poly(uint32_t nv, const float* __restrict__ verts, float px, float pz) {  // area.rs:29-50
    bool in = false;
    for (uint32_t i = 0, j = nv - 1; i < nv; j = i++) {
        const float vix = verts[i * 3], viz = verts[i * 3 + 2], vjx = verts[j * 3], vjz = verts[j * 3 + 2];
        if ((viz > pz) == (vjz > pz)) continue;
        const float e = __fadd_rn(__fdiv_rn(__fmul_rn(__fsub_rn(vjx, vix), __fsub_rn(pz, viz)), __fsub_rn(vjz, viz)), vix);
        if (px < e) in = !in;
    }
    return in;
}

__global__ void __launch_bounds__(256) k_area_marks(const TileDesc* __restrict__ tiles, TileMap tm, const uint32_t* __restrict__ hf_off,
                                                    const int16_t* __restrict__ smin, uint8_t* __restrict__ areas,
                                                    const rcb_area_op* __restrict__ ops, int op0, int op1,
                                                    const float* __restrict__ poly_verts) {
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
    const float cell_x = __fadd_rn(td.bmin[0], __fmul_rn(__fadd_rn((float)x, 0.5f), td.cs));
    const float cell_z = __fadd_rn(td.bmin[2], __fmul_rn(__fadd_rn((float)z, 0.5f), td.cs));
    for (int oi = op0; oi < op1; ++oi) {
        const rcb_area_op op = ops[oi];
        float lo[3], hi[3];
        if (op.type == RCB_AREA_OP_BOX) {
            for (int k = 0; k < 3; ++k) lo[k] = op.a[k], hi[k] = op.b[k];
        } else if (op.type == RCB_AREA_OP_CYLINDER) {  // :555-560
            lo[0] = __fsub_rn(op.a[0], op.b[0]), lo[1] = op.a[1], lo[2] = __fsub_rn(op.a[2], op.b[0]);
            hi[0] = __fadd_rn(op.a[0], op.b[0]), hi[1] = __fadd_rn(op.a[1], op.b[1]), hi[2] = __fadd_rn(op.a[2], op.b[0]);
        } else {  // :371-381
            const float* v = poly_verts + 3ull * op.first_vert;
            lo[0] = hi[0] = v[0];
            lo[2] = hi[2] = v[2];
            lo[1] = op.a[1];
            hi[1] = op.b[1];
            for (uint32_t i = 1; i < op.num_verts; ++i) {
                lo[0] = fminf(lo[0], v[i * 3]), lo[2] = fminf(lo[2], v[i * 3 + 2]);
                hi[0] = fmaxf(hi[0], v[i * 3]), hi[2] = fmaxf(hi[2], v[i * 3 + 2]);
            }
        }
        int minx = trunc_div(lo[0], td.bmin[0], td.cs), miny = trunc_div(lo[1], td.bmin[1], td.ch), minz = trunc_div(lo[2], td.bmin[2], td.cs);
        int maxx = trunc_div(hi[0], td.bmin[0], td.cs), maxy = trunc_div(hi[1], td.bmin[1], td.ch), maxz = trunc_div(hi[2], td.bmin[2], td.cs);
        if (maxx < 0 || minx >= W || maxz < 0 || minz >= H) continue;  // early-out
        minx = max(minx, 0), maxx = min(maxx, W - 1), minz = max(minz, 0), maxz = min(maxz, H - 1);
        if (x < minx || x > maxx || z < minz || z > maxz) continue;
        bool column_in = true;
        if (op.type == RCB_AREA_OP_CYLINDER) {  // :589-597
            const float dx = __fsub_rn(cell_x, op.a[0]), dz = __fsub_rn(cell_z, op.a[2]);
            column_in = !(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dz, dz)) >= __fmul_rn(op.b[0], op.b[0]));
        }
        if (!column_in) continue;
        bool poly_known = false, poly_in = false;
        for (uint32_t s = b; s < e; ++s) {
            if (areas[s] == 0) continue;
            const int y = smin[s];  // span.y
            if (y < miny || y > maxy) continue;
            if (op.type == RCB_AREA_OP_CONVEX_POLY) {
                if (!poly_known) {
                    poly_in = point_in_poly(op.num_verts, poly_verts + 3ull * op.first_vert, cell_x, cell_z);
                    poly_known = true;
                }
                if (!poly_in) continue;
            }
            areas[s] = (uint8_t)op.area_id;
        }
    }
}

__global__ void __launch_bounds__(256) k_area_median(const uint32_t* __restrict__ nid, size_t S, const uint8_t* __restrict__ areas,
                                                     uint8_t* __restrict__ out) {
    const size_t s = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    {
        const uint8_t own = areas[s];
        if (own == 0) {  // :256-259
            out[s] = 0;
            return;
        }
        uint8_t n[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) n[i] = own;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int dir = 1 + 2 * i;
            const uint32_t a = nid[(size_t)dir * S + s];
            if (a == RCB_NONE) continue;
            if (areas[a] != 0) n[i * 2] = areas[a];
            const uint32_t bb = nid[(size_t)((dir + 1) & 7) * S + a];
            if (bb != RCB_NONE && areas[bb] != 0) n[i * 2 + 1] = areas[bb];
        }
        // median of nine: rank by counting (same value as the reference's insertion sort, :283-284)
        uint8_t med = own;
#pragma unroll
        for (int i = 0; i < 9; ++i) {
            int less = 0, leq = 0;
#pragma unroll
            for (int j = 0; j < 9; ++j) {
                less += n[j] < n[i];
                leq += n[j] <= n[i];
            }
            if (less <= 4 && 4 < leq) med = n[i];
        }
        out[s] = med;
    }
}

// ConvexVolumeSet::apply_to_heightfield (convex_volume.rs:392-430) on the SOLID heightfield: one thread per column;
// every span takes the area of the last volume that contains (cell centre, span top).
__global__ void __launch_bounds__(256) k_convex_volumes(const TileDesc* __restrict__ tiles, TileMap tm, const uint32_t* __restrict__ hf_off,
                                                        const int16_t* __restrict__ smax, uint8_t* __restrict__ sarea,
                                                        const rcb_convex_volume* __restrict__ vols, int nvols,
                                                        const float* __restrict__ verts) {
    const uint32_t tile = RCB_TILE_OF_BLOCK();
    if (tile >= tm.ntiles) return;
    const TileDesc& td = tiles[tile];
    const uint32_t lc = RCB_CELL_OF_THREAD();
    if (lc >= (uint32_t)(td.width * td.height)) return;
    const uint32_t g = td.cell_base + lc;
    const uint32_t b = hf_off[g], e = hf_off[g + 1];
    if (b == e) return;
    const int z = (int)(lc / (uint32_t)td.width), x = (int)lc - z * td.width;
    const float wx = __fadd_rn(td.bmin[0], __fmul_rn(__fadd_rn((float)x, 0.5f), td.cs));  // :405-406
    const float wz = __fadd_rn(td.bmin[2], __fmul_rn(__fadd_rn((float)z, 0.5f), td.cs));
    for (uint32_t s = b; s < e; ++s) {
        const float wy = __fadd_rn(td.bmin[1], __fmul_rn((float)smax[s], td.ch));  // :417
        uint32_t area = sarea[s];
        for (int k = 0; k < nvols; ++k) {
            const rcb_convex_volume v = vols[k];
            if (wy < v.min_height || wy > v.max_height) continue;  // :132-134
            const float* p = verts + 3ull * v.first_vert;
            bool inside = false;
            for (uint32_t i = 0, j = v.num_verts - 1; i < v.num_verts; j = i++) {  // :140-164
                const float vix = p[3 * i], viz = p[3 * i + 2], vjx = p[3 * j], vjz = p[3 * j + 2];
                if ((viz > wz) == (vjz > wz)) continue;
                const float xi = __fadd_rn(__fdiv_rn(__fmul_rn(__fsub_rn(vjx, vix), __fsub_rn(wz, viz)), __fsub_rn(vjz, viz)), vix);
                if (wx >= xi) continue;
                inside = !inside;
            }
            if (inside) area = v.area_type & 0xffu;  // later volumes win (:383-387)
        }
        sarea[s] = (uint8_t)area;
    }
}

}  // namespace

void launch_convex_volumes(const TileDesc* tiles, TileMap tm, const uint32_t* hf_off, const int16_t* smax, uint8_t* sarea,
                           const rcb_convex_volume* vols, int nvols, const float* verts, cudaStream_t s) {
    if (tm.nblocks == 0 || nvols <= 0) return;
    KScope _k("k_convex_volumes", s);
    k_convex_volumes<<<tm.grid(), 256, 0, s>>>(tiles, tm, hf_off, smax, sarea, vols, nvols, verts);
}

void launch_area_marks(const TileDesc* tiles, TileMap tm, const uint32_t* hf_off, const int16_t* smin, uint8_t* areas,
                       const rcb_area_op* ops, int op0, int op1, const float* poly_verts, cudaStream_t s) {
    if (tm.nblocks == 0 || op1 <= op0) return;
    KScope _k("k_area_marks", s);
    k_area_marks<<<tm.grid(), 256, 0, s>>>(tiles, tm, hf_off, smin, areas, ops, op0, op1, poly_verts);
}

void launch_area_median(const uint32_t* nid, size_t S, const uint8_t* areas, uint8_t* out, cudaStream_t s) {
    if (S == 0) return;
    KScope _k("k_area_median", s);
    k_area_median<<<(unsigned)((S + 255) / 256), 256, 0, s>>>(nid, S, areas, out);
}
