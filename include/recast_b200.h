/*
 * recast_b200.h — C ABI of the B200-native tiled voxelization front end for recast-rs.
 *
 * This is the drop-in boundary: a Rust `-sys` crate, a C++ caller or Python
 * ctypes binds exactly these symbols.  Plain pointers and sizes only; no C++/torch types.
 * Every entry point names the reference interface it replaces (paths relative to the
 * recast-rs workspace, `crates/recast/src/...`).
 *
 * Conventions
 *   - all functions return 0 (RCB_OK) or a negative rcb_status; a human-readable message for the
 *     last failure on a context is available from rcb_last_error().  Nothing aborts or throws
 *     across this boundary.  Error kinds mirror recast_common::Error (recast-common/src/lib.rs:19-43).
 *   - one rcb_ctx per host thread and per device; calls on one context are serialised by the
 *     caller (the reference types are !Send/!Sync: heightfield.rs:75 Rc<RefCell<Span>>).
 *   - inputs are borrowed for the duration of the call; outputs are owned by an rcb_result and
 *     stay valid until rcb_result_free().
 *   - "NONE" for an Option<usize> in the reference is 0xFFFFFFFF here (RCB_NONE).
 *   - span / connection indices inside a tile view are TILE-RELATIVE, exactly the values the
 *     reference stores in a single CompactHeightfield.
 *   - there is no CPU fallback: without a CUDA device rcb_create fails with RCB_ECUDA.
 */
#ifndef RECAST_B200_H
#define RECAST_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RCB_NONE 0xFFFFFFFFu
#define RCB_NOT_CONNECTED 63u /* compact_heightfield.rs:61 — sentinel in con[4] (collides with a real offset 63, replicated) */

typedef enum rcb_status {
    RCB_OK = 0,
    RCB_EINVALID_MESH = -1, /* Error::InvalidMesh      (config.rs:113-133, tile_cache_builder.rs:138-143) */
    RCB_ENAVMESH_GEN = -2,  /* Error::NavMeshGeneration (lib.rs:199-233, heightfield.rs:103-115) */
    RCB_ECUDA = -3,         /* CUDA runtime failure / no device */
    RCB_EOVERFLOW = -4,     /* an internal pool overflowed (clip polygon > RCB_MAX_POLY_VERTS) */
    RCB_EARG = -5,          /* bad argument (null pointer, bad tile index, wrong stage order) */
    RCB_ERECAST = -6,       /* Error::Recast (detour-dynamic/src/dynamic_tile.rs:171-216: short span data) */
    RCB_EDETOUR_FAILURE = -7,      /* Error::Detour(Status::Failure): LZ4 decode failed (detour-tilecache/src/tile_cache.rs:865-870) */
    RCB_EDETOUR_INVALID_PARAM = -8 /* Error::Detour(Status::InvalidParam): TileCacheLayer::from_bytes (tile_cache_data.rs:83-91, 129, 282-313) */
} rcb_status;

/* Stage mask for rcb_build_tiles / rcb_build_from_spans. */
enum {
    RCB_STAGE_RASTER = 1u,  /* rasterization.rs:246-366 rasterize_tri + :51-168 add_span_internal */
    RCB_STAGE_FILTER = 2u,  /* heightfield.rs:286-531 the three walkable filters (lib.rs:278-287) */
    RCB_STAGE_COMPACT = 4u, /* compact_heightfield.rs:175-430 build_from_heightfield + build_connections */
    RCB_STAGE_ERODE = 8u,   /* area.rs:73-234 erode_walkable_area (runs before DIST when both set) */
    RCB_STAGE_DIST = 16u,   /* compact_heightfield.rs:514-727 build_distance_field */
    RCB_STAGE_REGIONS = 0x40u, /* CompactHeightfield::build_regions -> watershed::build_regions_watershed (watershed.rs:364-510)
                                  with the tile's border_size / min_region_area / merge_region_area; implies COMPACT | DIST */
    /* individual filters (Heightfield::filter_* called one at a time); RCB_STAGE_FILTER = all three */
    RCB_FILTER_LOW_HANGING = 0x100u, /* heightfield.rs:286 */
    RCB_FILTER_LEDGE = 0x200u,       /* heightfield.rs:332 */
    RCB_FILTER_LOW_HEIGHT = 0x400u,  /* heightfield.rs:492 */
    RCB_STAGE_ALL_BUILD = 1u | 2u | 4u | 16u, /* what RecastBuilder::build_mesh reaches (lib.rs:77-87 → watershed.rs:375) */
    /* Result layout flag (not a stage): rcb_result_download copies only what the host cannot rebuild, in the narrowest form
     * that holds it - the copy, not the kernels, is what an end-to-end build waits for (PCIe).  The HOST views then carry
     * hf_cell_offset = cell_index = first_conn = conn_next = conn_dir = NULL and
     *   conn_mask[S]      bit d set <=> the span has a connection in direction d.  Connections are laid out in span order and,
     *                     inside a span, in DESCENDING direction (compact_heightfield.rs:375-389 pushes dir 7..0 with
     *                     next -> the one pushed before, first_connection = the last pushed), so
     *                     first_conn[s] = (connections of spans <= s) - 1 (or NONE), conn_dir = the mask's bits from 7 down,
     *                     conn_next[k] = k - 1 inside a span's block, NONE at its start;
     *   conn_off8[K]      per connection, the OFFSET of the span it reaches inside that span's cell (u8).  The cell is the
     *                     neighbour cell of the connection's direction (DIR_OFFSET_X/Z, compact_heightfield.rs:109-111), its
     *                     first span is cell_index, so conn_ref[k] = cell_index[neighbour cell] + conn_off8[k]; and con[d4] of
     *                     the span is the same offset for the four cardinal directions (7 -> 0, 5 -> 1, 3 -> 2, 1 -> 3,
     *                     :397-427), 63 where the mask has none.  conn_ref = conn_ref16 = con = con8 = NULL then.
     *                     Should an offset pass 255 (a cell with 257+ spans) conn_off8 = NULL and the references and con are
     *                     there instead: conn_ref16 (u16; conn_ref u32 if a tile holds 65536+ spans), con8 (u8; con if an
     *                     offset passes 255 there too);
     *   cell_count16[C]   cell_count_arr as u16 (cell_count_arr = NULL; the wide one instead should a cell hold 65536+ spans);
     *                     hf_cell_offset is its exclusive prefix sum, cell_index the same with NONE for empty cells;
     *   areas             NULL while CompactHeightfield.areas still equals CompactSpan.area (span_area): nothing that writes
     *                     it ran (no ERODE, no area operator, no areas_override).
     * INTEGRATION.md has the rebuild loop.  Device views (rcb_result_tile_device) always carry the full arrays. */
    RCB_RESULT_SLIM = 0x1000u
};

/* Field-for-field RecastConfig (config.rs:7-51). */
typedef struct rcb_config {
    int32_t width;  /* cells along x */
    int32_t height; /* cells along z */
    float cs;
    float ch;
    float bmin[3];
    float bmax[3];
    float walkable_slope_angle;
    int32_t walkable_height;
    int32_t walkable_climb;
    int32_t walkable_radius;
    int32_t max_edge_len;
    float max_simplification_error;
    int32_t min_region_area;
    int32_t merge_region_area;
    int32_t max_vertices_per_polygon;
    float detail_sample_dist;
    float detail_sample_max_error;
    int32_t border_size;
} rcb_config;

typedef struct rcb_ctx rcb_ctx;
typedef struct rcb_result rcb_result;

/* Totals over a batch (the T,V,C,S,K of SURVEY.md §8(d)); filled by rcb_result_totals. */
typedef struct rcb_totals {
    uint64_t tiles;
    uint64_t triangles;
    uint64_t vertices;
    uint64_t cells;
    uint64_t fragments; /* (triangle,cell) fragments emitted by the clipper; 0 when RASTER did not run */
    uint64_t spans;
    uint64_t connections;
    uint64_t walkable_spans;
    uint64_t result_bytes; /* bytes of the result arena (what rcb_result_download copies D2H) */
} rcb_totals;

/*
 * One tile's result, flat SoA.  Pointers are host pointers (after rcb_result_download) or
 * device pointers (rcb_result_tile_device); arrays that a skipped stage would have produced
 * are NULL.  Layouts replace, in order:
 *   Heightfield  (heightfield.rs:58-76)           -> CSR: hf_cell_offset[C+1], span_min/max/area[S]
 *   CompactCell  (compact_heightfield.rs:15-20)   -> cell_index[C] (RCB_NONE = None), cell_count[C]
 *   CompactSpan  (compact_heightfield.rs:31-48)   -> span_min/max[S], span_area[S], con[4*S], first_conn[S]
 *                                                    (y == min, h == (u8)(max-min), reg == 0: derivable)
 *   CompactConnection (:68-75)                    -> conn_ref[K], conn_dir[K], conn_next[K]
 *   CompactHeightfield.{areas,dist} (:152-154)    -> areas[S], dist[S]
 */
typedef struct rcb_tile_view {
    int32_t width, height;
    uint32_t cell_count;  /* C = width*height */
    uint32_t span_count;  /* S */
    uint32_t conn_count;  /* K */
    uint32_t walkable_span_count;
    uint16_t max_height;   /* compact_heightfield.rs:240 */
    uint16_t max_distance; /* compact_heightfield.rs:540,672 — max BEFORE the blur */
    uint32_t status;       /* 0 ok; bit0: clip polygon overflow */
    const uint32_t* hf_cell_offset; /* [C+1], tile-relative, ascending */
    const int16_t* span_min;        /* [S] */
    const int16_t* span_max;        /* [S] */
    const uint8_t* span_area;       /* [S]  Span.area / CompactSpan.area (post-filter when FILTER ran) */
    const uint32_t* cell_index;     /* [C]  Some(first span) or RCB_NONE */
    const uint32_t* cell_count_arr; /* [C] */
    const uint8_t* areas;           /* [S]  CompactHeightfield.areas (post-erosion when ERODE ran) */
    const uint16_t* con;            /* [4*S] span-major: con[4*s + d4], d4: 0 W,1 S,2 E,3 N; 63 = none */
    const uint32_t* first_conn;     /* [S]  RCB_NONE = None */
    const uint32_t* conn_ref;       /* [K]  tile-relative span index */
    const uint8_t* conn_dir;        /* [K]  0..7 = NW,N,NE,E,SE,S,SW,W */
    const uint32_t* conn_next;      /* [K]  RCB_NONE = None */
    const uint16_t* dist;           /* [S]  zeros unless DIST ran (compact_heightfield.rs:253) */
    const uint16_t* reg;            /* [S]  CompactSpan.reg (watershed.rs:505-507); NULL unless REGIONS ran */
    uint32_t max_regions;           /* chf.max_regions (watershed.rs:491: the id counter BEFORE filtering); 0 unless REGIONS ran */
    uint32_t _pad;
    const uint8_t* conn_mask;       /* [S]  RCB_RESULT_SLIM only: directions in which the span is connected (bit d) */
    const uint16_t* conn_ref16;     /* [K]  RCB_RESULT_SLIM only: conn_ref in 16 bits (NULL when a tile has 65536+ spans) */
    const uint16_t* cell_count16;   /* [C]  RCB_RESULT_SLIM only: cell_count_arr in 16 bits */
    const uint8_t* con8;            /* [4*S] RCB_RESULT_SLIM only: con in 8 bits */
    const uint8_t* conn_off8;       /* [K]  RCB_RESULT_SLIM only: offset of the connected span inside its (neighbour) cell */
} rcb_tile_view;

/* ---- context ------------------------------------------------------------------------------ */
int rcb_create(int device, rcb_ctx** out);
void rcb_destroy(rcb_ctx* ctx);
const char* rcb_last_error(const rcb_ctx* ctx);
/* Launch every kernel and copy of this context on `cuda_stream` (a cudaStream_t, used verbatim: NULL is
 * the legacy default stream).  Lets a caller order and time the work with its own events (bench.py passes
 * a torch stream).  A new context runs on a private non-blocking stream; rcb_use_own_stream returns to it. */
int rcb_set_stream(rcb_ctx* ctx, void* cuda_stream);
int rcb_use_own_stream(rcb_ctx* ctx);
/* Number of kernel launches issued by this context since creation (bench `gpu_launches`). */
uint64_t rcb_launch_count(const rcb_ctx* ctx);
/* Diagnostics: out[0] = batches issued without a host round trip (sizes taken from the previous build of the same
 * configurations and verified on the device), out[1] = how many of those had to be rebuilt because a size was short,
 * out[2] = configuration batches cached by the context.  The reference has no counterpart (it has no device). */
int rcb_ctx_stats(const rcb_ctx* ctx, uint64_t out[4]);
/* Memory-safety stand-in for compute-sanitizer (closed on this GPU pool): librecast_b200_dbg.so is the same library built
 * with -DRCB_BOUNDS, in which every index into the hand-laid-out shared-memory regions and per-tile arena slices of the
 * clipper, the tile-resident kernels and the region builder goes through a checked accessor.  Returns 1 in that build (0 in
 * the normal one, where the accessors are raw pointers) and fills out[0] = violations since the last reset on this
 * context's device, out[1] = file id << 32 | line of the first, out[2] = its index, out[3] = the array's size.
 * reset: 0 keep the counters, 1 clear them, 2 first commit one deliberate violation (self-test of the checked build), then as 1. */
int rcb_debug_bounds_report(rcb_ctx* ctx, uint64_t out[4], int reset);

/* Per-stage device timers (CUDA events around every kernel launch), the device-side counterpart of
 * RecastContext::start_timer/stop_timer with TimerCategory::{Rasterization, Filtering,
 * CompactHeightfield, ...} (context.rs:25-54,223-251).  Off by default.  rcb_profile_read
 * synchronises the stream, writes up to `cap` entries accumulated since the last reset and
 * returns the number of entries available. */
typedef struct rcb_stage_time {
    char name[40];
    float ms;       /* summed device time of the stage's launches */
    uint32_t calls; /* launches accumulated */
} rcb_stage_time;
int rcb_profile_enable(rcb_ctx* ctx, int on);
int rcb_profile_read(rcb_ctx* ctx, rcb_stage_time* out, int cap);
int rcb_profile_reset(rcb_ctx* ctx);

/* ---- config (config.rs) ------------------------------------------------------------------ */
void rcb_config_default(rcb_config* cfg);                                        /* config.rs:53-76 */
void rcb_config_calculate_grid_size(rcb_config* cfg, const float bmin[3], const float bmax[3]); /* config.rs:85-107 */
int rcb_config_validate(const rcb_config* cfg);                                  /* config.rs:110-136; RCB_EINVALID_MESH */

/* ---- mesh ---------------------------------------------------------------------------------
 * The (&[f32] vertices, &[i32] indices) pair every RecastBuilder entry point takes (lib.rs:68,116,186).
 * rcb_upload_mesh copies host -> device; rcb_set_mesh_device borrows device memory (zero-copy). */
int rcb_upload_mesh(rcb_ctx* ctx, const float* verts, size_t nfloats, const int32_t* tris, size_t nidx);
int rcb_set_mesh_device(rcb_ctx* ctx, const void* d_verts, size_t nfloats, const void* d_tris, size_t nidx);
/* The other two index forms of the reference's free functions; rcb_rasterize then is
 *   rasterize_triangles_u16 (rasterization.rs:432-455)   rcb_upload_mesh_u16: `tris[i] as usize`, widened on the device;
 *   rasterize_triangle_soup (rasterization.rs:459-476)   rcb_upload_triangle_soup: 9 floats per triangle, no index list.
 * (rcb_build_tiles works on them too.) */
int rcb_upload_mesh_u16(rcb_ctx* ctx, const float* verts, size_t nfloats, const uint16_t* tris, size_t nidx);
int rcb_upload_triangle_soup(rcb_ctx* ctx, const float* verts, size_t ntris);

/* ---- batched tile build ---------------------------------------------------------------------
 * rcb_build_tiles: for every tile config runs RecastBuilder::build_heightfield (lib.rs:116-135 ->
 * rasterize_triangles lib.rs:186-290: slope classification, rasterize_triangle with merge
 * threshold 1, then the three filters when FILTER is set), then the stages named by stage_mask.
 * Errors follow lib.rs:195-233: empty verts/indices => Ok with empty heightfields; len%3 != 0 or an
 * index >= nverts => RCB_ENAVMESH_GEN (validated up front; the reference drops its partial
 * heightfield on Err so the observable result is the same).
 * This is the multi-tile entry the reference lacks: each tile is an independent
 * Heightfield, as in DynamicTile::build (detour-dynamic/src/dynamic_tile.rs:444-483). */
int rcb_build_tiles(rcb_ctx* ctx, const rcb_config* cfgs, int ntiles, uint32_t stage_mask,
                    int erode_radius, rcb_result** out);

/* rcb_build_tiles_streamed: rcb_build_tiles + rcb_result_download for callers that want the voxel data in HOST
 * memory: the tiles are processed in `ngroups` consecutive groups (<= 0: 8) and the device-to-host copy of a
 * finished group runs on a second stream while the next groups compute, so the call costs about
 * max(compute, PCIe copy) instead of their sum.  The result is already downloaded; rcb_result_tile /
 * _tile_device / _totals work as usual (tile indices are those of `cfgs`), rcb_result_serialize_spans does not
 * (serialize from an rcb_build_tiles result).  Same errors as rcb_build_tiles (the first failing group's). */
int rcb_build_tiles_streamed(rcb_ctx* ctx, const rcb_config* cfgs, int ntiles, uint32_t stage_mask, int erode_radius,
                             int ngroups, rcb_result** out);

/* rcb_rasterize: the free function rasterize_triangles(verts, tris, tri_area_ids, n, hf, thr)
 * (rasterization.rs:405-428) per tile, with caller-provided per-triangle areas and merge
 * threshold; no slope classification, no degenerate-triangle skip.  Later stages per stage_mask. */
int rcb_rasterize(rcb_ctx* ctx, const rcb_config* cfgs, int ntiles, const uint8_t* tri_areas,
                  int32_t flag_merge_threshold, uint32_t stage_mask, int erode_radius, rcb_result** out);

/* rcb_rasterize_onto: rasterize the uploaded mesh onto heightfields that ALREADY hold spans (host CSR as for
 * rcb_build_from_spans): what rasterize_triangles (rasterization.rs:405-428) and RecastBuilder::rasterize_triangles
 * (lib.rs:186-290) do when the &mut Heightfield they are given is not empty — detour-dynamic re-applies colliders to a
 * restored checkpoint this way (dynamic_tile.rs).  The existing spans are the initial state of every column's list;
 * the new fragments are merged into it in triangle order with add_span_internal's rule.  tri_areas == NULL: areas by
 * slope as RecastBuilder does (flag_merge_threshold should then be 1, lib.rs:273); else the given areas.
 * Later stages per stage_mask (filters run over the whole heightfield, as lib.rs:276-287 does). */
int rcb_rasterize_onto(rcb_ctx* ctx, const rcb_config* cfgs, int ntiles, const uint32_t* cell_offset, const int16_t* span_min,
                       const int16_t* span_max, const uint8_t* span_area, const uint8_t* tri_areas, int32_t flag_merge_threshold,
                       uint32_t stage_mask, int erode_radius, rcb_result** out);

/* rcb_build_from_spans: start from existing heightfields (host CSR, one per tile, concatenated):
 * Heightfield::filter_* (heightfield.rs:286,332,492), CompactHeightfield::build_from_heightfield
 * (compact_heightfield.rs:175), build_distance_field (:514), erode_walkable_area (area.rs:73).
 * cell_offset holds, tile after tile, width*height+1 tile-relative offsets; span arrays are
 * concatenated in tile order.  RCB_STAGE_RASTER must not be set.
 * areas_override (nullable, [S]): CompactHeightfield.areas as the caller last left them (after an
 * earlier erode / area marking); it replaces the areas copied from span_area AFTER the connections
 * are built and BEFORE ERODE / DIST run, which is the state the reference's operators see when they
 * are called one at a time on a live CompactHeightfield (connections are never rebuilt). */
int rcb_build_from_spans(rcb_ctx* ctx, const rcb_config* cfgs, int ntiles, const uint32_t* cell_offset,
                         const int16_t* span_min, const int16_t* span_max, const uint8_t* span_area,
                         const uint8_t* areas_override, uint32_t stage_mask, int erode_radius, rcb_result** out);

/* ---- area operators on the compact heightfield ----------------------------------------
 * rcb_apply_area_ops: the reference's stand-alone CompactHeightfield.areas operators, applied IN ORDER to
 * every tile of the batch (volumes are given in world coordinates):
 *   RCB_AREA_OP_MEDIAN      median_filter_walkable_area (area.rs:238-294)
 *   RCB_AREA_OP_BOX         mark_box_area               (area.rs:298-355)  a = bmin, b = bmax
 *   RCB_AREA_OP_CYLINDER    mark_cylinder_area          (area.rs:544-618)  a = position, b[0] = radius, b[1] = height
 *   RCB_AREA_OP_CONVEX_POLY mark_convex_poly_area       (area.rs:359-437)  a[1] = min_y, b[1] = max_y,
 *                           vertices poly_verts[3*first_vert .. 3*(first_vert+num_verts))
 * Input is an existing heightfield (CSR, as rcb_build_from_spans) plus the caller's current areas
 * (areas_override, nullable); the connections are rebuilt from span_area exactly as
 * CompactHeightfield::build_from_heightfield does, then the operators run, then ERODE / DIST when asked. */
enum { RCB_AREA_OP_MEDIAN = 0, RCB_AREA_OP_BOX = 1, RCB_AREA_OP_CYLINDER = 2, RCB_AREA_OP_CONVEX_POLY = 3 };
typedef struct rcb_area_op {
    int32_t type;
    uint32_t area_id;
    float a[3];
    float b[3];
    uint32_t first_vert, num_verts;
} rcb_area_op;
int rcb_apply_area_ops(rcb_ctx* ctx, const rcb_config* cfgs, int ntiles, const uint32_t* cell_offset, const int16_t* span_min,
                       const int16_t* span_max, const uint8_t* span_area, const uint8_t* areas_override, const rcb_area_op* ops,
                       int nops, const float* poly_verts, size_t npoly_floats, uint32_t stage_mask, int erode_radius,
                       rcb_result** out);

/* ---- convex volumes (ConvexVolumeSet::apply_to_heightfield, convex_volume.rs:392-430) -------------
 * rcb_build_tiles_with_volumes: RecastBuilder::build_heightfield_with_volumes / build_mesh_with_volumes
 * (lib.rs:293-361): rcb_build_tiles with the volumes applied to the SOLID heightfield after the filters and
 * before compaction.  Every span takes the area_type of the LAST volume that contains the point
 * (tile bmin.x + (x + 0.5) * cs, bmin.y + span.max * ch, bmin.z + (z + 0.5) * cs) — get_area_at_point (:375-389),
 * contains_point (:130-164: min_height <= y <= max_height and an odd number of crossings in xz).
 * A volume is what ConvexVolume::new (:28-68) accepts: 3..32 vertices (xyz triples at poly_verts[3*first_vert ..]),
 * min_height <= max_height, convex in xz (is_convex :166-199); otherwise RCB_EINVALID_MESH.
 * rcb_apply_convex_volumes: the same on existing heightfields (CSR as rcb_build_from_spans). */
typedef struct rcb_convex_volume {
    uint32_t first_vert, num_verts;
    float min_height, max_height;
    uint32_t area_type; /* u8 */
} rcb_convex_volume;
int rcb_build_tiles_with_volumes(rcb_ctx* ctx, const rcb_config* cfgs, int ntiles, uint32_t stage_mask, int erode_radius,
                                 const rcb_convex_volume* volumes, int nvolumes, const float* poly_verts, size_t npoly_floats,
                                 rcb_result** out);
int rcb_apply_convex_volumes(rcb_ctx* ctx, const rcb_config* cfgs, int ntiles, const uint32_t* cell_offset, const int16_t* span_min,
                             const int16_t* span_max, const uint8_t* span_area, const rcb_convex_volume* volumes, int nvolumes,
                             const float* poly_verts, size_t npoly_floats, uint32_t stage_mask, int erode_radius, rcb_result** out);

/* ---- heightfield wire formats of detour-dynamic ----------------------------------------
 * A tile's span data is, per cell in z-major order, `u16 count` then count x (`u32 min, u32 max, u32 area`).
 *   RCB_WIRE_VOXEL_TILE    big-endian.  Writer: VoxelTile::serialize_spans (io/voxel_tile.rs:70-118, `min as u32`
 *                          sign-extends the i16).  Reader: VoxelTile::deserialize_spans (:120-176) — lenient: a
 *                          short buffer ends the span loop of the cell it hits and the walk goes on from
 *                          there; add_span errors (min > max) drop the span.
 *   RCB_WIRE_DYNAMIC_TILE  little-endian.  Reader: DynamicTile::reconstruct_heightfield (dynamic_tile.rs:147-257)
 *                          — strict: short buffer -> RCB_ERECAST, add_span error -> RCB_ENAVMESH_GEN.  The
 *                          reference has no writer for this byte order (its own writer/reader pair disagree,
 *                          SURVEY §8(f)); the writer here is the byte-swapped twin of serialize_spans.
 * Both readers insert every span through Heightfield::add_span (heightfield.rs:102-222), whose merge rule
 * (same-area overlap, at most one follower absorbed) is NOT the rasterizer's.
 *
 * rcb_result_span_data_size: bytes of all tiles' streams = 2 * cells + 12 * spans.
 * rcb_result_serialize_spans: writes the streams of every tile of a result (heightfield as it stands after
 *   the result's filters), tile after tile, into `out` (host, >= rcb_result_span_data_size bytes);
 *   tile_offsets (nullable) receives ntiles + 1 byte offsets.
 * rcb_build_from_span_data: parses ntiles streams (data + tile_offsets[i] .. tile_offsets[i+1]) into
 *   heightfields on the device and runs the later stages of stage_mask on them (RCB_STAGE_RASTER must not be
 *   set), exactly like rcb_build_from_spans.
 * RCB_STAGE_READD (for rcb_build_from_spans): re-insert the given spans through Heightfield::add_span, in
 *   column order, before anything else — DynamicTileCheckpoint::clone_heightfield (checkpoint.rs:29-60). */
typedef enum rcb_wire_format { RCB_WIRE_VOXEL_TILE = 0, RCB_WIRE_DYNAMIC_TILE = 1 } rcb_wire_format;
#define RCB_STAGE_READD 0x20u
uint64_t rcb_result_span_data_size(const rcb_result* res);
int rcb_result_serialize_spans(rcb_result* res, int format, uint8_t* out, uint64_t capacity, uint64_t* tile_offsets);
int rcb_build_from_span_data(rcb_ctx* ctx, const rcb_config* cfgs, int ntiles, const uint8_t* data, const uint64_t* tile_offsets,
                             int format, uint32_t stage_mask, int erode_radius, rcb_result** out);

/* ---- tile-cache layers (SURVEY §8(a) row TC, BASELINE config 5) ---------------------------------------
 * rcb_build_tilecache_layers: TileCacheBuilder::layer_to_compact_heightfield
 * (detour-tilecache/src/tile_cache_builder.rs:100-126: decode_layer_to_heightfield :129-172, then
 * CompactHeightfield::build_from_heightfield) followed by rasterize_obstacles (:175-251 with
 * rasterize_cylinder/box/oriented_box_obstacle :254-320, :422-467, :470-549) for a batch of layers.
 * Every non-empty cell (regons != 0) becomes ONE span (hmin, hmin+1, areas[idx]); obstacles are applied
 * AFTER the connections are built and clear CompactSpan.area only (never CompactHeightfield.areas), as
 * the reference does.  Errors: regons_len/areas_len != width*height -> RCB_EINVALID_MESH (:138-143);
 * hmin == 32767 with a non-empty cell -> RCB_ENAVMESH_GEN (Heightfield::add_span min > max,
 * heightfield.rs:110-115).  stage_mask: RCB_STAGE_COMPACT is implied; RCB_STAGE_DIST / _ERODE optional. */
typedef struct rcb_tc_layer {
    float bmin[3], bmax[3]; /* TileCacheLayerHeader.bmin/bmax (tile_cache_data.rs:33-35) */
    uint16_t hmin, hmax;    /* :37-39 */
    uint8_t width, height;  /* :41-43 */
    uint16_t _pad;
    const uint8_t* regons;  /* [width*height], 0 = empty cell (TileCacheLayer.regons :242) */
    const uint8_t* areas;   /* [width*height] (:244) */
    uint32_t regons_len, areas_len;
    uint32_t first_obstacle, obstacle_count; /* the obstacles passed to build_tile_from_layer for this layer */
} rcb_tc_layer;

enum { RCB_OBSTACLE_CYLINDER = 0, RCB_OBSTACLE_BOX = 1, RCB_OBSTACLE_ORIENTED_BOX = 2 };
enum { RCB_OBSTACLE_EMPTY = 0, RCB_OBSTACLE_PROCESSING = 1, RCB_OBSTACLE_PROCESSED = 2, RCB_OBSTACLE_REMOVING = 3 };

/* ObstacleData (tile_cache.rs:126-152) + ObstacleState (:109-118), flattened:
 *   cylinder:     a = pos,    b[0] = radius, b[1] = height
 *   box:          a = bmin,   b = bmax
 *   oriented box: a = center, b = half_extents, rot_aux = [cos(a/2)sin(-a/2), cos(a/2)cos(a/2) - 0.5] */
typedef struct rcb_tc_obstacle {
    int32_t type;
    int32_t state;
    float a[3];
    float b[3];
    float rot_aux[2];
} rcb_tc_obstacle;

int rcb_build_tilecache_layers(rcb_ctx* ctx, const rcb_tc_layer* layers, int nlayers, const rcb_tc_obstacle* obstacles,
                               int nobstacles, float cs, float ch, uint32_t stage_mask, int erode_radius,
                               rcb_result** out);

/* ---- tile-cache front half --------------------------------------------------------
 * Compressed tiles as TileCache stores them (tile_cache.rs:34, 848-851): lz4_flex::compress_prepend_size =
 * 4-byte little-endian uncompressed size + one LZ4 block; decoded they are TileCacheLayer::to_bytes
 * (tile_cache_data.rs:261-278: 54-byte header, three u32 lengths, regons, areas, cons).
 *
 * rcb_decompress_tiles: TileCache::decompress_tile (tile_cache.rs:854-872) for n blobs
 *   (blobs + blob_offsets[i] .. blob_offsets[i+1]).  out_offsets[n+1] (output) are the positions of the decoded
 *   tiles in `out` (host; capacity out_cap bytes; pass out = NULL to only size it: out_offsets[n] is the total).
 *   An empty blob decodes to nothing; a blob lz4_flex rejects -> RCB_EDETOUR_FAILURE (first such tile).
 *   As lz4_flex does, a block that ends early yields FEWER bytes than its size prefix says.
 * rcb_build_tilecache_compressed: per layer decompress_tile -> TileCacheLayer::from_bytes ->
 *   rcb_build_tilecache_layers' pipeline, everything on the device (one small D2H of the parsed headers).
 *   first_obstacle / obstacle_count (nullable: no obstacles): the range of `obstacles` that applies to layer i.
 *   Errors are those of the first layer that fails, in the order the reference meets them:
 *   RCB_EDETOUR_FAILURE, RCB_EDETOUR_INVALID_PARAM, RCB_EINVALID_MESH, RCB_ENAVMESH_GEN. */
int rcb_decompress_tiles(rcb_ctx* ctx, const uint8_t* blobs, const uint64_t* blob_offsets, int n, uint8_t* out, uint64_t out_cap,
                         uint64_t* out_offsets);
int rcb_build_tilecache_compressed(rcb_ctx* ctx, const uint8_t* blobs, const uint64_t* blob_offsets, int nlayers,
                                   const uint32_t* first_obstacle, const uint32_t* obstacle_count, const rcb_tc_obstacle* obstacles,
                                   int nobstacles, float cs, float ch, uint32_t stage_mask, int erode_radius, rcb_result** out);

/* ---- results ------------------------------------------------------------------------------- */
int rcb_result_totals(const rcb_result* res, rcb_totals* out);
int rcb_result_download(rcb_result* res);                       /* one D2H of the result arena into pinned memory */
int rcb_result_tile(const rcb_result* res, int tile, rcb_tile_view* out);        /* host pointers; needs download */
int rcb_result_tile_device(const rcb_result* res, int tile, rcb_tile_view* out); /* device pointers; scalars need download */
void rcb_result_free(rcb_result* res);

/* ---- several devices of one box --------------------------------------------------------------------
 * The multi-tile loop of detour-dynamic (DynamicNavMesh::build / build_async, crates/detour-dynamic/src/dynamic_navmesh.rs:
 * 558-684: `for key in tile_keys` -> DynamicTile::build) over the GPUs of one box, from ONE host thread of the caller:
 *   rcb_multi_create        two rcb_ctx (and, during a build, two host threads) per listed device.
 *   rcb_multi_upload_mesh   the mesh goes host -> devices[0] once and from there to the other devices (peer copy over NVLink);
 *                           indices are validated as in rcb_upload_mesh.
 *   rcb_multi_build_tiles   tiles are independent (no halo, border_size 0 on every multi-tile path of the reference), so the
 *                           configurations are cut into groups of `tiles_per_group` consecutive tiles (<= 0: about four
 *                           groups per device); two internal workers per device (a context and a host thread each, so that
 *                           two groups run side by side on a GPU) work through their own shares of the groups, then steal
 *                           from the worker with the most left.  No data moves
 *                           between devices.  download != 0: every group's result is also copied to pinned host memory by
 *                           its device's thread (rcb_multi_result_tile then works at once).
 *   rcb_multi_result_*      as rcb_result_*: tile indices are those of `cfgs`; rcb_multi_result_tile_device also reports the
 *                           device that holds the tile.  tiles_per_device (nullable, ndevices entries): how the queue dealt.
 * Errors: the first failing group's status (same kinds as rcb_build_tiles). */
typedef struct rcb_multi rcb_multi;
typedef struct rcb_multi_result rcb_multi_result;
int rcb_multi_create(const int* devices, int ndevices, rcb_multi** out);
void rcb_multi_destroy(rcb_multi* m);
const char* rcb_multi_last_error(const rcb_multi* m);
int rcb_multi_device_count(const rcb_multi* m);
int rcb_multi_upload_mesh(rcb_multi* m, const float* verts, size_t nfloats, const int32_t* tris, size_t nidx);
int rcb_multi_build_tiles(rcb_multi* m, const rcb_config* cfgs, int ntiles, uint32_t stage_mask, int erode_radius,
                          int tiles_per_group, int download, rcb_multi_result** out);
int rcb_multi_result_download(rcb_multi_result* res);
int rcb_multi_result_totals(const rcb_multi_result* res, rcb_totals* out, int* tiles_per_device, int ndevices);
int rcb_multi_result_tile(const rcb_multi_result* res, int tile, rcb_tile_view* out);
int rcb_multi_result_tile_device(const rcb_multi_result* res, int tile, rcb_tile_view* out, int* device);
void rcb_multi_result_free(rcb_multi_result* res);

#ifdef __cplusplus
}
#endif
#endif /* RECAST_B200_H */
