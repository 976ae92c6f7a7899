// rcb_common.cuh — shared device/host declarations for the B200 voxelization kernels.
//
// Data layout in HBM for one batch of tiles (DESIGN.md "Data layout"):
//   * cells of all tiles are concatenated in tile order; tile t owns global cells
//     [cell_base[t], cell_base[t] + W_t*H_t), row-major z*W + x inside the tile;
//   * spans of all tiles are concatenated in global cell order (so also tile order); every
//     index handed back through the C ABI is made tile-relative by subtracting the tile's base;
//   * connections likewise.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/recast_b200.h"

#define RCB_MAX_POLY_VERTS 12  // clip polygon capacity (geometric bound is 7); overflow sets status bit0
#define RCB_NEI_NONE 0xFFFFu
#ifndef RCB_CLIP_ROWS
#define RCB_CLIP_ROWS 6  // rows of one (triangle, tile) pair handled by one clip work item (384 / rows items per CTA, k_raster2.cu)
#endif

// Per-tile constants, uploaded once per batch.
struct TileDesc {
    int32_t width, height;
    float bmin[3], bmax[3];
    float cs, ch;
    float ics, ich;          // 1/cs, 1/ch (rasterization.rs:254-255)
    float slope_thr;         // cos(angle*PI/180) computed on the host with cosf (lib.rs:212-213)
    int32_t walkable_height; // already cast `as i16` then widened (lib.rs:278-287)
    int32_t walkable_climb;
    int32_t neg_climb;       // (-climb as i16) as i32, heightfield.rs:382
    uint32_t cell_base;      // first global cell
    int32_t bin_x0, bin_z0;  // lower corner of the tile's range in the locator grid (dedup rule)
    int32_t aux0, aux1, aux2;   // tile-cache layers: hmin, first obstacle, obstacle count
    unsigned long long wmagic;  // ceil(2^40 / width): c / width == (c * wmagic) >> 40 for c < 2^24, width < 2^16
    // how the tile is cut into bands for the tile-resident kernels (BandDesc below); whole tile: band_rows == height
    int32_t band_rows;
    uint32_t band_first;        // index of the tile's first band (the bands of a tile are consecutive)
    unsigned long long bmagic;  // ceil(2^40 / band_rows): z / band_rows == (z * bmagic) >> 40
};

__host__ __device__ inline uint32_t rcb_div_w(uint32_t c, unsigned long long wmagic) {
    return (uint32_t)(((unsigned long long)c * wmagic) >> 40);
}

// Per-tile results (device-written, downloaded with the arena).
struct TileInfo {
    uint32_t span_base, span_count;
    uint32_t conn_base, conn_count;
    uint32_t walkable_span_count;
    uint32_t max_height;    // u16 value
    uint32_t max_distance;  // u16 value
    uint32_t status;
    uint32_t max_regions;   // chf.max_regions (watershed.rs:491), 0 unless REGIONS ran
};

// A band = rows [z0, z1) of one tile, the unit of work of the tile-resident kernels (k_tile.cu).  A tile whose working set
// fits shared memory is one band; a larger one (128x128 cells, deep columns) is cut into bands of equal height that run as
// independent CTAs.  What a band needs from its neighbours is one halo row on either side, [hz0, hz1) = [z0-1, z1+1) clamped
// to the tile: the ledge filter and the 8-direction linking look one cell away, and the distance field's row-sequential
// sweep hands its last row to the next band through global memory.
struct BandDesc {
    uint32_t tile;
    int32_t z0, z1;    // own rows
    int32_t hz0, hz1;  // rows held in shared memory (own + halo)
};
struct BandInfo {  // device-written per band
    uint32_t span_base;   // first own span, batch-wide (k_tile_scan)
    uint32_t span_count;  // own spans
    uint32_t top_spans;   // spans of the first own row (what the previous band keeps as its bottom halo)
    uint32_t bot_spans;   // spans of the last own row
    uint32_t walkable, max_height;
    uint32_t pub_base;    // byte offset of this band's hand-over records (k_tile_scan)
    uint32_t status;
};

// Uniform locator grid over the tile AABBs: maps a triangle AABB to candidate tiles.
struct Locator {
    float gx0, gz0;      // origin
    float inv_bx, inv_bz;  // 1 / bin size
    int32_t nbx, nbz;
    const uint32_t* bin_off;    // [nbx*nbz + 1]
    const uint32_t* bin_tiles;  // tile ids
    float ex0, ez0, ex1, ez1;  // xz extent of the batch's tiles (infinite when a tile bound is not finite)
};

// Device-side bin coordinate; monotone in v, identical on host and device (sub, mul, floor, clamp).
__host__ __device__ inline int32_t rcb_bin_coord(float v, float g0, float inv, int32_t nb) {
    float f = floorf((v - g0) * inv);
    if (!(f > 0.0f)) return 0;  // also NaN
    if (f >= (float)(nb - 1)) return nb - 1;
    return (int32_t)f;
}

struct MeshView {
    const float* verts;
    const int32_t* tris;
    uint32_t ntris;
    uint32_t nverts;
    const uint8_t* tri_areas;  // explicit areas (rcb_rasterize) or nullptr (classify by slope)
    int32_t merge_thr;         // flag_merge_threshold
    const float4* blocks;      // nullable: xz AABB (min x, min z, max x, max z) of every run of 256 consecutive triangles
};
#define RCB_MESH_BLOCK 256
void launch_widen_indices(const uint16_t* src_or_null, int32_t* dst, size_t n, cudaStream_t s);
void launch_mesh_blocks(const float* verts, const int32_t* tris, uint32_t ntris, uint32_t nverts, float4* blocks, cudaStream_t s);

#ifdef __CUDACC__
__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }
#endif

// ---- launchers (one per .cu file) ---------------------------------------------------------------
struct RasterBuffers {
    // inputs
    const TileDesc* tiles;
    int ntiles;
    Locator loc;
    MeshView mesh;
    // scratch
    unsigned long long* counters;  // [0] fragment upper bound, [1] fragment cursor, [2] overflow flags
    uint32_t* cell_cnt;            // [Ctot] fragment count per cell, later span count per cell
    uint4* frag_tmp;               // [frag capacity] unsorted fragment records
    uint32_t frag_cap;
};

void launch_count_fragments(const RasterBuffers& rb, cudaStream_t s);
void launch_clip_emit(const RasterBuffers& rb, TileInfo* tinfo, cudaStream_t s);
// column segments hold fragment records `rec` = {triangle, min | max << 16, area}: 8 bytes each, 16 (`wide`) when triangle
// indices need more than 24 bits
void launch_scatter_fragments(const uint4* frag_tmp, const unsigned long long* nfrag_dev, uint32_t nfrag_hint,
                              const uint32_t* frag_off, void* rec, bool wide, cudaStream_t s);
void launch_replay_columns(uint32_t ncells, const uint32_t* frag_off, void* rec, bool wide, uint32_t* span_cnt, int32_t merge_thr,
                           cudaStream_t s, const uint32_t* ex_off = nullptr);
void launch_onto(uint32_t ncells, const uint32_t* ex_off, uint32_t* cell_cnt, int place, const uint32_t* frag_off, const int16_t* smin,
                 const int16_t* smax, const uint8_t* sarea, void* rec, bool wide, cudaStream_t s);
void launch_gather_records(uint32_t ncells, const uint32_t* frag_off, const void* rec, bool wide, const uint32_t* hf_off, int16_t* smin,
                           int16_t* smax, uint8_t* sarea, cudaStream_t s);
void launch_gather_spans(uint32_t ncells, const uint32_t* frag_off, const uint32_t* fs_mm, const uint8_t* fs_area,
                         const uint32_t* hf_off, int16_t* smin, int16_t* smax, uint8_t* sarea, cudaStream_t s);

// second-generation front end (k_raster2.cu)
struct Raster2 {
    const TileDesc* tiles;
    int ntiles;
    Locator loc;
    MeshView mesh;
    unsigned long long* counters;  // [0] frag upper bound [1] frag cursor [2] flags (1 poly overflow, 2 pool overflow)
                                   // [3] bad index [4] work items
    uint32_t* item_cnt;            // [ntris + 1]
    uint32_t* tile_ub;             // [nbands] per-band fragment upper bound (nullable)
    // global-list mode (v1 back end)
    uint32_t* cell_cnt;
    uint4* frag_tmp;
    uint32_t frag_cap;
    // per-band segment mode (tile-resident back end); enabled when tile_cursor != nullptr
    uint32_t* tile_cursor;     // [nbands]
    const uint32_t* tile_base; // [nbands] first record of the band's segment
    const uint32_t* tile_cap;  // [nbands]
    uint32_t* tile_frags;      // per segment three arrays of tile_cap words: cell | area<<24, smin | smax<<16, triangle
    uint32_t flags;            // 0x100: phase timers (RCB_PHASE_TIMERS); 0x200: count fragments per column while emitting (global back end)
    uint32_t spec;             // 1: issued speculatively - the work-item count is read from counters[4] on the device
};
// counters[2] flag bits: 1 clip polygon overflow, 2 fragment pool overflow, 4 a tile does not fit shared memory,
// 8 a speculative size was too small (every later kernel of the batch returns at once), 16 connection arena too small
#define RCB_FLAG_ABORT 8ull
#define RCB_FLAG_CONN 16ull
void launch_check_indices(const int32_t* tris, size_t nidx, uint32_t nverts, unsigned long long* flag, cudaStream_t s);
void launch_bin_count(const Raster2& rb, cudaStream_t s);
void launch_fill_items(const Raster2& rb, const uint32_t* item_off, uint2* items, cudaStream_t s);
void launch_clip_items(const Raster2& rb, const uint2* items, uint32_t nitems, TileInfo* tinfo, cudaStream_t s);

// Rows per band of a tile of `width` x `height` cells when tiles are cut into bands of about `band_cells` cells: the
// smallest number of bands that keeps a band within band_cells, all bands equal (the last may be shorter), never fewer
// than two rows (a band hands its last TWO rows' values to the next one).  Same function on host and device.
__host__ __device__ inline int32_t rcb_band_rows(int32_t width, int32_t height, uint32_t band_cells) {
    if ((unsigned long long)width * (unsigned long long)height <= band_cells) return height;
    long long r = (long long)(band_cells / (uint32_t)width);
    if (r < 2) r = 2;
    const long long nb = (height + r - 1) / r;
    r = (height + nb - 1) / nb;
    if (r < 2) r = 2;
    return (int32_t)r;
}

// tile-resident back end (k_tile.cu)
struct TileSpansArgs {
    const TileDesc* tiles;
    const BandDesc* bands;
    BandInfo* binfo;
    const uint32_t* tile_frags;   // per-band segments: three arrays of tile_cap words each (3 words per record in all)
    const uint32_t* tile_base;    // [ntiles] first record of the segment (also the base of ts_* below)
    const uint32_t* tile_cap;     // [ntiles]
    const uint32_t* tile_cursor;  // [ntiles] records written
    uint32_t* hf_cell_offset;     // [Ctot + ntiles] tile-relative CSR offsets (C+1 per tile)
    uint32_t* ts_mm;              // tile-local compact spans: smin | smax << 16
    uint8_t* ts_area;
    unsigned long long* counters;
    uint32_t fmask;
    int32_t merge_thr;
    uint32_t smem_bytes;
    unsigned long long* phase;  // optional [16] per-phase cycle accumulators (nullptr = off)
    int threads;                // CTA size: 512 (whole tiles) or 256 (tiles cut into bands)
    int pass;                   // 0: first launch, 1: second launch for tiles deferred by pass 0
    int defer;                  // pass 0 may defer tiles that need more shared memory (a pass 1 follows)
};
struct TileCompactArgs {
    const TileDesc* tiles;
    const BandDesc* bands;
    const BandInfo* binfo;
    const uint32_t* band_first;  // [ntiles + 1]
    uint32_t* work;              // per-band scratch lists (the bands' fragment segments, dead by now): 3 words per segment record
    unsigned char* pub;          // hand-over records between the bands of a tile (BandInfo::pub_base)
    uint32_t* flag_f;            // [nbands] zeroed: band's forward sweep published
    uint32_t* flag_d;            // [nbands] zeroed: band's distance values published
    const uint32_t* tile_cap;    // [nbands] records a band's fragment segment holds (bounds of `work`)
    const uint32_t* tile_base;
    const uint32_t* hf_cell_offset_in;  // from k_tile_spans
    const uint32_t* ts_mm;
    const uint8_t* ts_area;
    TileInfo* tinfo;
    unsigned long long* counters;
    unsigned long long* lookback;  // [ntiles], zeroed
    uint32_t* tile_ticket;         // zeroed
    // outputs (final arrays)
    uint32_t* hf_cell_offset;
    int16_t* smin;
    int16_t* smax;
    uint8_t* sarea;
    uint32_t* cell_index;
    uint32_t* cell_count;
    uint8_t* areas;
    uint16_t* con;
    uint32_t* first_conn;
    uint16_t* dist;
    uint32_t* conn_ref;
    uint8_t* conn_dir;
    uint32_t* conn_next;
    int do_compact, do_erode, do_dist, erode_radius;
    int threads;                // CTA size: 512 (whole tiles) or 256 (tiles cut into bands)
    uint32_t smem_bytes;
    unsigned long long Kcap;    // capacity of the connection arrays (a tile that would pass it sets RCB_FLAG_CONN)
    unsigned long long* phase;  // optional [16]: slots 8.. are this kernel's phases
};
void launch_tile_spans(const TileSpansArgs& a, int nbands, cudaStream_t s);
// cap_frags / cap_items: what the fragment and work-item buffers hold (~0: no check); beyond -> RCB_FLAG_ABORT
void launch_tile_segments(const uint32_t* tile_ub, uint32_t* tile_base, int ntiles, unsigned long long* counters,
                          unsigned long long cap_frags, unsigned long long cap_items, cudaStream_t s);
// span_cap: what the result arena was laid out for (~0: no check); beyond -> RCB_FLAG_ABORT
// counters[6] spans, [7] walkable spans, [8] shared-memory bytes the largest band needs in k_tile_compact, [10] hand-over bytes
void launch_tile_scan(const BandDesc* bands, BandInfo* binfo, int nbands, const TileDesc* tiles, const uint32_t* band_first, TileInfo* tinfo,
                      int ntiles, unsigned long long* counters, unsigned long long span_cap, unsigned long long pub_cap, int compact,
                      const uint32_t* seg_cursor, const uint32_t* seg_cap, cudaStream_t s);
// slim result layout: conn_mask[s] = the span's directions, conn_off8[k] = the connection's offset inside the neighbour cell;
// conn_ref16 / cell_count16 / con8 = narrow forms of conn_ref, cell_count, con; *flag |= 8 / 1 / 4 / 2 where a value does not fit
void launch_slim_pack(const TileDesc* tiles, const TileInfo* tinfo, int ntiles, uint32_t max_cells, const uint32_t* first_conn,
                      const uint32_t* conn_ref, const uint8_t* conn_dir, const uint32_t* conn_next, const uint32_t* cell_index,
                      const uint32_t* cell_count, const uint16_t* con, uint8_t* conn_mask, uint16_t* conn_ref16, uint8_t* conn_off8,
                      uint16_t* cell_count16, uint8_t* con8, unsigned long long* flag, cudaStream_t s);
void launch_tile_compact(const TileCompactArgs& a, int nbands, cudaStream_t s);
size_t tile_compact_smem_bytes(uint32_t own_cells, uint32_t held_cells, uint32_t held_spans, uint32_t top_spans, uint32_t bot_spans);

void launch_seg_scatter(const TileDesc* tiles, int ntiles, const uint32_t* tile_frags, const uint32_t* tile_base, const uint32_t* tile_cap,
                        const uint32_t* tile_cursor, const uint32_t* frag_off, uint32_t* cell_cnt, void* rec, bool wide,
                        unsigned long long* counters, uint32_t max_tile_frags, cudaStream_t s);

// exclusive scan of n u32 values, out has n+1 entries (out[n] = total).  tmp: >= scan_tmp_elems(n) u32.
size_t scan_tmp_elems(size_t n);
void launch_exclusive_scan(const uint32_t* in, uint32_t* out, size_t n, uint32_t* tmp, cudaStream_t s);

// Per-column kernels run on a 3-D grid: blockIdx.x = 256-column block inside a tile,
// tile = blockIdx.z * gridDim.y + blockIdx.y.  Blocks past a (partial) tile's last cell exit.
struct TileMap {
    uint32_t ntiles;
    uint32_t blocks_per_tile;  // ceil(max cells per tile / 256)
    uint32_t nblocks;          // 0 when there is nothing to launch
#ifdef __CUDACC__
    dim3 grid() const {
        const uint32_t gy = ntiles < 65535u ? ntiles : 65535u;
        return dim3(blocks_per_tile, gy, (ntiles + gy - 1) / gy);
    }
#endif
};
#ifdef __CUDACC__
#define RCB_TILE_OF_BLOCK() (blockIdx.z * gridDim.y + blockIdx.y)
#define RCB_CELL_OF_THREAD() (blockIdx.x * blockDim.x + threadIdx.x)
#endif

// k_regions.cu — watershed region building.  All span arrays are batch-wide (indexed by the tile's span_base);
// tiles[].aux0/1/2 carry border_size / min_region_area / merge_region_area.
struct RegionArgs {
    const TileDesc* tiles;
    TileInfo* tinfo;
    const uint32_t* cell_index;
    const uint32_t* cell_count;
    const uint8_t* areas;
    const uint16_t* con;
    const uint16_t* dist;
    const uint32_t* first_conn;
    const uint32_t* conn_ref;
    const uint8_t* conn_dir;
    const uint32_t* conn_next;
    uint16_t* reg;  // out
    size_t Stot;
    // scratch
    uint16_t* src_dist;
    uint16_t* reg0;
    uint32_t* nb4;     // [4][Stot]
    uint32_t* stacks;  // [8][Stot]
    uint32_t* bufA;
    uint32_t* bufB;
    int32_t* r_cnt;    // per-region tables: tile t owns entries [span_base + 2t, +span_count + 2)
    int32_t* r_cadd;
    uint16_t* r_alive;
    uint8_t* r_border;
    unsigned long long* r_best;
    uint16_t* r_target;
    uint16_t* r_fin;
    unsigned int* err;  // first tile whose region ids overflowed (init 0xffffffff)
    unsigned long long* phase;  // RCB_PHASE_TIMERS: cycle sums per phase (thread 0 of every tile), or nullptr
    uint32_t smem_spans;        // tiles with at most this many spans keep their arrays in shared memory (0: none do)
    uint32_t all_small;         // every tile of the batch has fewer than 65535 spans: 16-bit neighbour table (one layout per batch:
                                // 16- and 32-bit planes of different tiles would overlap)
    int narrow;                 // 0: 64 threads; 1: the 32-thread build (32 CTAs per SM), for batches that would not be resident
                                // otherwise; 2: the 256-thread build, for a few very large tiles
};
size_t regions_smem_bytes(uint32_t max_spans);
void launch_regions(const RegionArgs& a, int ntiles, cudaStream_t s);

// k_lz4.cu — tile-cache front half
struct RcbTcParsed {  // one layer after TileCacheLayer::from_bytes; status = the rcb_status of the first failing check
    int32_t status;
    uint32_t width, height, hmin;
    float bmin[3], bmax[3];
    uint32_t regons_off, areas_off;  // inside the decoded bytes
};
void launch_lz4_decompress(const uint8_t* blobs, const unsigned long long* blob_off, int nlayers, uint8_t* out,
                           const unsigned long long* out_off, uint32_t* status, uint32_t* produced, cudaStream_t s);
void launch_tc_parse(int nlayers, const uint8_t* raw, const unsigned long long* raw_off, const uint32_t* lz_status,
                     const uint32_t* produced, RcbTcParsed* outp, cudaStream_t s);
void launch_tc_gather(const TileDesc* tiles, TileMap tm, const uint8_t* raw, const unsigned long long* raw_off,
                      const RcbTcParsed* parsed, uint8_t* regons, uint8_t* areas, cudaStream_t s);

// k_wire.cu — heightfield wire formats
void launch_wire_walk(const TileDesc* tiles, int ntiles, const uint8_t* data, const unsigned long long* tile_off, int strict_le,
                      uint32_t* cell_pos, uint32_t* cell_n, unsigned long long* err, cudaStream_t s);
void launch_wire_serialize(const TileDesc* tiles, TileMap tm, const uint32_t* hf_cell_offset, const TileInfo* tinfo,
                           const int16_t* smin, const int16_t* smax, const uint8_t* sarea, uint8_t* out, int big_endian,
                           cudaStream_t s);
void launch_wire_replay(const TileDesc* tiles, TileMap tm, const uint8_t* data, const unsigned long long* tile_off,
                        const uint32_t* cell_pos, int strict_le, const int16_t* smin, const int16_t* smax, const uint8_t* sarea,
                        const uint32_t* in_off, uint32_t* mm, uint8_t* ar, uint32_t* span_cnt, unsigned long long* err, cudaStream_t s);

#ifdef __CUDACC__
// Block-ordered compaction of the threads that have work: returns how many, sh_list[i] = local index of the i-th.
__device__ __forceinline__ uint32_t compact_block256(bool pred, uint32_t* sh_list, uint32_t* sh_wcnt) {
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const unsigned m = __ballot_sync(0xffffffffu, pred);
    if (lane == 0) sh_wcnt[warp] = __popc(m);
    __syncthreads();
    uint32_t base = 0, total = 0;
#pragma unroll
    for (uint32_t w = 0; w < 8; ++w) {
        const uint32_t c = sh_wcnt[w];
        if (w < warp) base += c;
        total += c;
    }
    if (pred) sh_list[base + __popc(m & ((1u << lane) - 1u))] = threadIdx.x;
    __syncthreads();
    return total;
}
#endif

void launch_cells_and_offsets(const TileDesc* tiles, TileMap tm, const uint32_t* hf_off, uint32_t* hf_cell_offset,
                              uint32_t* cell_index, uint32_t* cell_count, TileInfo* tinfo, int ntiles,
                              cudaStream_t s);
void launch_filter_spans2(const TileDesc* tiles, size_t S, const uint2* span_cell, const uint32_t* hf_off, const int16_t* smin,
                          const int16_t* smax, const uint8_t* sarea, uint8_t* area_out, uint32_t fmask, cudaStream_t s);
void launch_span_cells(const TileDesc* tiles, TileMap tm, const uint32_t* hf_off, const int16_t* smin, const int16_t* smax, size_t S,
                       uint2* span_cell, cudaStream_t s);  // span_cell[S].x = "some column is not ordered"
void launch_link_spans2(const TileDesc* tiles, size_t S, const uint2* span_cell, const uint32_t* hf_off, const int16_t* smin,
                        const int16_t* smax, const uint8_t* sarea, uint8_t* areas, uint32_t* nid, uint16_t* con, uint32_t* span_conn_cnt,
                        TileInfo* tinfo, cudaStream_t s);
void launch_tile_conn_bases(int ntiles, const uint32_t* conn_off_span, TileInfo* tinfo, cudaStream_t s);
void launch_write_connections2(size_t S, const uint2* span_cell, const uint32_t* nid, const uint32_t* conn_off_span, const TileInfo* tinfo,
                               uint32_t* first_conn, uint32_t* conn_ref, uint8_t* conn_dir, uint32_t* conn_next, cudaStream_t s);
void launch_tile_bases(const TileDesc* tiles, int ntiles, const uint32_t* hf_off, const uint32_t* conn_off,
                       TileInfo* tinfo, cudaStream_t s);

// `nid` = u32[8][S] absolute neighbour span ids (k_link_spans2), `span_cell` = (tile, local cell) per span
void launch_distance_field(const TileDesc* tiles, int ntiles, const uint32_t* hf_off, const uint32_t* nid, const uint2* span_cell, size_t S,
                           const uint8_t* areas, uint16_t* tmp_init, uint16_t* tmp_f, uint16_t* tmp_d, uint16_t* dist, TileInfo* tinfo,
                           uint4* fidx, int max_width, cudaStream_t s);
void launch_erode(const TileDesc* tiles, int ntiles, const uint32_t* hf_off, const uint32_t* nid, size_t S, uint8_t* areas,
                  uint8_t* tmp_init, uint8_t* tmp_f, int radius, uint4* fidx, int max_width, cudaStream_t s);

// area operators (k_area.cu)
void launch_convex_volumes(const TileDesc* tiles, TileMap tm, const uint32_t* hf_off, const int16_t* smax, uint8_t* sarea,
                           const rcb_convex_volume* vols, int nvols, const float* verts, cudaStream_t s);
void launch_area_marks(const TileDesc* tiles, TileMap tm, const uint32_t* hf_off, const int16_t* smin, uint8_t* areas,
                       const rcb_area_op* ops, int op0, int op1, const float* poly_verts, cudaStream_t s);
void launch_area_median(const uint32_t* nid, size_t S, const uint8_t* areas, uint8_t* out, cudaStream_t s);

// tile-cache layers (k_tilecache.cu)
void launch_tc_count(const TileDesc* tiles, TileMap tm, const uint8_t* regons, uint32_t* cell_cnt, cudaStream_t s);
void launch_tc_spans(const TileDesc* tiles, TileMap tm, const uint8_t* regons, const uint8_t* areas, const uint32_t* hf_off,
                     int16_t* smin, int16_t* smax, uint8_t* sarea, cudaStream_t s);
void launch_tc_obstacles(const TileDesc* tiles, int ntiles, const rcb_tc_obstacle* obs, const uint32_t* hf_off, const int16_t* smin,
                         const int16_t* smax, uint8_t* sarea, cudaStream_t s);

// -DRCB_BOUNDS build: per translation unit, the violations its checked accessors saw (rcb_bounds.cuh); zeros otherwise
void rcb_bounds_read_tile(unsigned long long out[4], int reset);
void rcb_bounds_read_raster2(unsigned long long out[4], int reset);
void rcb_bounds_read_regions(unsigned long long out[4], int reset);
void rcb_bounds_selftest(uint32_t* scratch16, cudaStream_t s);  // one deliberate violation (debug build; no-op otherwise)

// Per-kernel hooks of the context that is running a pipeline on THIS host thread (one rcb_ctx per host thread): the launch
// counter behind rcb_launch_count and, when rcb_profile_enable is on, an event pair around every kernel so device time is
// attributed per kernel name.  Thread-local: contexts on other threads neither see nor clear them.
struct RcbProfHook {
    void (*mark)(void* user, const char* name, int begin, cudaStream_t s);
    void* user;
    unsigned long long* launches;
};
extern thread_local RcbProfHook g_rcb_prof;

// cudaFuncAttributeMaxDynamicSharedMemorySize is one value per (kernel, device), shared by every host thread that drives
// the device: two contexts on one device (rcb_multi's two workers, a pipelined caller) that set it to their own batch's
// size and then launch can undo each other - the launch with the larger size then fails, and the kernels behind it read
// a stale intermediate.  The limit is therefore only ever RAISED, under a lock; a launch may use less than the limit.
void rcb_raise_dynamic_smem(const void* kernel, size_t bytes);
bool rcb_sync_check();                                         // RCB_SYNC_CHECK set in the environment (read once)
void rcb_sync_check_kernel(const char* name, cudaStream_t s);  // launch error / stream synchronize -> stderr
struct KScope {
    const char* name;
    cudaStream_t s;
    KScope(const char* n, cudaStream_t st) : name(n), s(st) {
        if (g_rcb_prof.mark) g_rcb_prof.mark(g_rcb_prof.user, name, 1, s);
    }
    ~KScope() {
        if (g_rcb_prof.launches) ++*g_rcb_prof.launches;
        if (g_rcb_prof.mark) g_rcb_prof.mark(g_rcb_prof.user, name, 0, s);
        if (rcb_sync_check()) rcb_sync_check_kernel(name, s);  // RCB_SYNC_CHECK=1: wait for the kernel, name it if it failed
    }
};
