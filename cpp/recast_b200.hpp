// recast_b200.hpp — header-only C++ host wrapper over the C ABI (include/recast_b200.h).
//
// Mirrors the reference's builder interface for the voxelization half
// (crates/recast/src/lib.rs:51-359: RecastBuilder::new / config / build_heightfield / build_mesh's
// voxel stages) so a C++ caller reads like the reference's Rust:
//
//     recast_b200::RecastConfig cfg;                       // RecastConfig::default(), config.rs:53-76
//     cfg.calculate_grid_size(bmin, bmax);                 // config.rs:85-107
//     recast_b200::Context ctx(0);
//     recast_b200::RecastBuilder builder(cfg, ctx);
//     auto batch = builder.build_tiles(verts, nfloats, tris, nidx, tiles.data(), tiles.size());
//     rcb_tile_view v = batch.tile(0);                     // flat SoA of Heightfield + CompactHeightfield
//
// Round-2 additions: build_tiles_streamed (host-to-host, copy overlapped with compute, optionally the slim layout with
// expand_slim() as the binding layer's rebuild), rasterize_triangles_u16 / _soup entry forms, Context::stats(), and
// MultiContext (several GPUs of one box behind one handle, the loop of DynamicNavMesh::build).
//
// Errors: recast_common::Error::{InvalidMesh, NavMeshGeneration} map to the exceptions below.
#pragma once
#include <cstring>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../include/recast_b200.h"

namespace recast_b200 {

struct Error : std::runtime_error {
    int status;
    Error(int st, const std::string& msg) : std::runtime_error(msg), status(st) {}
};
struct InvalidMesh : Error {
    using Error::Error;
};
struct NavMeshGeneration : Error {
    using Error::Error;
};

inline void check(int st, const rcb_ctx* ctx, const char* what) {
    if (st == RCB_OK) return;
    std::string msg = ctx ? rcb_last_error(ctx) : "";
    if (msg.empty()) msg = what;
    if (st == RCB_EINVALID_MESH) throw InvalidMesh(st, msg);
    if (st == RCB_ENAVMESH_GEN) throw NavMeshGeneration(st, msg);
    throw Error(st, msg);
}

struct RecastConfig : rcb_config {
    RecastConfig() { rcb_config_default(this); }
    void calculate_grid_size(const float bmin_[3], const float bmax_[3]) { rcb_config_calculate_grid_size(this, bmin_, bmax_); }
    void validate() const { check(rcb_config_validate(this), nullptr, "Invalid configuration"); }
};

class Context {
   public:
    explicit Context(int device = 0) { check(rcb_create(device, &ctx_), nullptr, "rcb_create failed: no CUDA device (no CPU fallback)"); }
    ~Context() { rcb_destroy(ctx_); }
    Context(const Context&) = delete;
    Context& operator=(const Context&) = delete;
    rcb_ctx* raw() const { return ctx_; }
    void set_stream(void* cuda_stream) { check(rcb_set_stream(ctx_, cuda_stream), ctx_, "set_stream"); }  // verbatim; nullptr = legacy default stream
    // The (&[f32] vertices, &[i32] indices) pair every RecastBuilder entry point takes (lib.rs:68,116,186), kept on the device
    // for the builds that follow; indices are validated here (lib.rs:224-233 -> NavMeshGeneration).
    void upload_mesh(const float* verts, size_t nfloats, const int32_t* tris, size_t nidx) {
        check(rcb_upload_mesh(ctx_, verts, nfloats, tris, nidx), ctx_, "upload_mesh");
    }
    // A mesh that already lives on this context's device (borrowed until the next mesh call; must not change meanwhile).
    void set_mesh_device(const void* d_verts, size_t nfloats, const void* d_tris, size_t nidx) {
        check(rcb_set_mesh_device(ctx_, d_verts, nfloats, d_tris, nidx), ctx_, "set_mesh_device");
    }
    void use_own_stream() { check(rcb_use_own_stream(ctx_), ctx_, "use_own_stream"); }
    uint64_t launch_count() const { return rcb_launch_count(ctx_); }
    // RecastContext::start_timer / stop_timer per TimerCategory (context.rs:25-54,223-251), device side: summed kernel time per name.
    void profile_enable(bool on) { check(rcb_profile_enable(ctx_, on ? 1 : 0), ctx_, "profile_enable"); }
    void profile_reset() { check(rcb_profile_reset(ctx_), ctx_, "profile_reset"); }
    std::vector<rcb_stage_time> profile_read() {
        std::vector<rcb_stage_time> out(128);
        const int n = rcb_profile_read(ctx_, out.data(), (int)out.size());
        if (n < 0) check(n, ctx_, "profile_read");
        out.resize(n < (int)out.size() ? n : (int)out.size());
        return out;
    }
    // -DRCB_BOUNDS build only (librecast_b200_dbg.so): violations the checked accessors saw; false in the normal build.
    bool bounds_report(uint64_t out[4], int reset = 0) { return rcb_debug_bounds_report(ctx_, out, reset) == 1; }
    struct Stats {
        uint64_t issued_without_round_trip, rebuilt, cached_batches;
    };
    // builds issued with sizes assumed from the previous build of the same configurations (verified on the device), how
    // many of those had to be redone, configuration batches cached by the context
    Stats stats() const {
        uint64_t o[4] = {0, 0, 0, 0};
        check(rcb_ctx_stats(ctx_, o), ctx_, "stats");
        return Stats{o[0], o[1], o[2]};
    }

   private:
    rcb_ctx* ctx_ = nullptr;
};

class Batch;
// VoxelTile::deserialize_spans (io/voxel_tile.rs:120-176, RCB_WIRE_VOXEL_TILE) / DynamicTile::reconstruct_heightfield
// (dynamic_tile.rs:147-257, RCB_WIRE_DYNAMIC_TILE) for a batch of span streams, later stages per stage_mask.
inline Batch build_from_span_data(Context& ctx, const rcb_config* tiles, size_t ntiles, const uint8_t* data, const uint64_t* tile_offsets,
                                  int format, uint32_t stage_mask = 0, int erode_radius = 0);
// TileCache::decompress_tile + TileCacheLayer::from_bytes + layer_to_compact_heightfield + rasterize_obstacles
// (detour-tilecache/src/tile_cache.rs:854-872, tile_cache_data.rs:281-320, tile_cache_builder.rs:100-251) for a batch of compressed tiles.
inline Batch build_tilecache_compressed(Context& ctx, const uint8_t* blobs, const uint64_t* blob_offsets, size_t nlayers,
                                        const uint32_t* first_obstacle, const uint32_t* obstacle_count, const rcb_tc_obstacle* obstacles,
                                        size_t nobstacles, float cs, float ch, uint32_t stage_mask = RCB_STAGE_COMPACT);

class Batch {
   public:
    Batch(rcb_ctx* ctx, rcb_result* res) : ctx_(ctx), res_(res) {}
    Batch(Batch&& o) noexcept : ctx_(o.ctx_), res_(std::exchange(o.res_, nullptr)) {}
    Batch(const Batch&) = delete;
    ~Batch() { rcb_result_free(res_); }
    rcb_totals totals() const {
        rcb_totals t{};
        check(rcb_result_totals(res_, &t), ctx_, "totals");
        return t;
    }
    void download() { check(rcb_result_download(res_), ctx_, "download"); }
    rcb_tile_view tile(int i) const {  // host pointers, valid while the Batch lives; call download() first
        rcb_tile_view v{};
        check(rcb_result_tile(res_, i, &v), ctx_, "tile");
        return v;
    }
    rcb_tile_view tile_device(int i) const {
        rcb_tile_view v{};
        check(rcb_result_tile_device(res_, i, &v), ctx_, "tile_device");
        return v;
    }
    // VoxelTile::serialize_spans (detour-dynamic/src/io/voxel_tile.rs:70-118) for every tile; tile_offsets gets ntiles + 1 entries.
    std::vector<uint8_t> serialize_spans(int format = RCB_WIRE_VOXEL_TILE, std::vector<uint64_t>* tile_offsets = nullptr) const {
        std::vector<uint8_t> out(rcb_result_span_data_size(res_));
        if (tile_offsets) tile_offsets->assign(totals().tiles + 1, 0);
        check(rcb_result_serialize_spans(res_, format, out.data(), out.size(), tile_offsets ? tile_offsets->data() : nullptr), ctx_,
              "serialize_spans");
        return out;
    }

   private:
    rcb_ctx* ctx_;
    rcb_result* res_;
};

// What an RCB_RESULT_SLIM download leaves to the binding layer, rebuilt from the layout rule of
// compact_heightfield.rs:375-389: connections lie in span order and, inside a span, in descending direction (directions 7..0
// are pushed with `next` -> the one pushed before; `first_connection` = the last pushed).  The same loop a Rust shim runs
// while it fills Vec<CompactSpan>; held to byte equality with the full layout by cpp/example.cpp on a GPU.
struct SlimTile {
    std::vector<uint32_t> hf_cell_offset, cell_index, cell_count, first_conn, conn_ref, conn_next;
    std::vector<uint16_t> con;
    std::vector<uint8_t> conn_dir, areas;
};

inline SlimTile expand_slim(const rcb_tile_view& v) {
    if (!v.conn_mask) throw Error(RCB_EARG, "expand_slim: not a slim host view (download an RCB_RESULT_SLIM result first)");
    static const int dx[8] = {-1, 0, 1, 1, 1, 0, -1, -1}, dz[8] = {-1, -1, -1, 0, 1, 1, 1, 0};  // compact_heightfield.rs:109-111
    static const int d4[8] = {-1, 3, -1, 2, -1, 1, -1, 0};  // 8-direction -> con slot (:397-427); -1: diagonal
    const uint32_t C = v.cell_count, S = v.span_count, K = v.conn_count;
    SlimTile t;
    t.cell_count.resize(C);
    t.hf_cell_offset.resize((size_t)C + 1);
    t.cell_index.resize(C);
    uint32_t run = 0;
    for (uint32_t c = 0; c < C; ++c) {
        const uint32_t n = v.cell_count16 ? v.cell_count16[c] : v.cell_count_arr[c];
        t.cell_count[c] = n;
        t.hf_cell_offset[c] = run;
        t.cell_index[c] = n ? run : RCB_NONE;
        run += n;
    }
    t.hf_cell_offset[C] = run;
    if (run != S) throw Error(RCB_EARG, "expand_slim: cell counts do not add up to the span count");
    const uint8_t* areas = v.areas ? v.areas : v.span_area;  // nothing wrote chf.areas: still the copy of CompactSpan.area it starts as
    t.areas.assign(areas, areas + S);
    t.first_conn.assign(S, RCB_NONE);
    t.conn_ref.resize(K);
    t.conn_next.resize(K);
    t.conn_dir.resize(K);
    t.con.assign((size_t)4 * S, 63);
    if (!v.conn_off8) {  // a connection's offset passed 255 somewhere: con travelled after all, in 8 or 16 bits
        for (size_t i = 0; i < (size_t)4 * S; ++i) t.con[i] = v.con8 ? (uint16_t)v.con8[i] : v.con[i];
    }
    uint32_t k = 0;
    for (uint32_t c = 0; c < C; ++c)
        for (uint32_t sp = t.hf_cell_offset[c]; sp < t.hf_cell_offset[c + 1]; ++sp) {
            const uint32_t mask = v.conn_mask[sp];
            bool first = true;
            for (int d = 7; d >= 0; --d) {
                if (!(mask & (1u << d))) continue;
                if (k >= K) throw Error(RCB_EARG, "expand_slim: more mask bits than connections");
                t.conn_dir[k] = (uint8_t)d;
                t.conn_next[k] = first ? RCB_NONE : k - 1;
                first = false;
                if (v.conn_off8) {
                    const uint32_t ncell = (uint32_t)((int)c + dz[d] * v.width + dx[d]);  // in the tile: the mask says a neighbour exists
                    t.conn_ref[k] = t.cell_index[ncell] + v.conn_off8[k];
                    if (d4[d] >= 0) t.con[(size_t)4 * sp + d4[d]] = v.conn_off8[k];
                } else {
                    t.conn_ref[k] = v.conn_ref16 ? (uint32_t)v.conn_ref16[k] : v.conn_ref[k];
                }
                t.first_conn[sp] = k++;
            }
        }
    if (k != K) throw Error(RCB_EARG, "expand_slim: mask bits do not add up to the connection count");
    return t;
}

class RecastBuilder {
   public:
    RecastBuilder(const RecastConfig& cfg, Context& ctx) : cfg_(cfg), ctx_(ctx) {}
    const RecastConfig& config() const { return cfg_; }

    // RecastBuilder::build_heightfield (lib.rs:116-135): one tile, rasterize + the three filters.
    Batch build_heightfield(const float* verts, size_t nfloats, const int32_t* tris, size_t nidx) {
        return run(verts, nfloats, tris, nidx, &cfg_, 1, RCB_STAGE_RASTER | RCB_STAGE_FILTER, 0);
    }
    // Voxel half of RecastBuilder::build_mesh (lib.rs:68-87 -> watershed.rs:375) for many tiles at once.
    Batch build_tiles(const float* verts, size_t nfloats, const int32_t* tris, size_t nidx, const rcb_config* tiles, size_t ntiles,
                      uint32_t stage_mask = RCB_STAGE_ALL_BUILD, int erode_radius = 0) {
        return run(verts, nfloats, tris, nidx, tiles, ntiles, stage_mask, erode_radius);
    }
    // The host-to-host form: tiles in `ngroups` consecutive groups, a finished group's copy to pinned host memory overlapping
    // the next groups' kernels; slim = RCB_RESULT_SLIM (a quarter of the bytes; expand_slim() rebuilds the rest per tile).
    Batch build_tiles_streamed(const float* verts, size_t nfloats, const int32_t* tris, size_t nidx, const rcb_config* tiles, size_t ntiles,
                               int ngroups = 4, bool slim = true, uint32_t stage_mask = RCB_STAGE_ALL_BUILD, int erode_radius = 0) {
        check(rcb_upload_mesh(ctx_.raw(), verts, nfloats, tris, nidx), ctx_.raw(), "upload_mesh");
        rcb_result* res = nullptr;
        check(rcb_build_tiles_streamed(ctx_.raw(), tiles, (int)ntiles, stage_mask | (slim ? (uint32_t)RCB_RESULT_SLIM : 0u), erode_radius, ngroups,
                                       &res),
              ctx_.raw(), "build_tiles_streamed");
        return Batch(ctx_.raw(), res);  // already downloaded
    }
    // rasterize_triangles_u16 (rasterization.rs:432-455): 16-bit indices, widened on the device.
    Batch build_tiles_u16(const float* verts, size_t nfloats, const uint16_t* tris, size_t nidx, const rcb_config* tiles, size_t ntiles,
                          uint32_t stage_mask = RCB_STAGE_ALL_BUILD, int erode_radius = 0) {
        check(rcb_upload_mesh_u16(ctx_.raw(), verts, nfloats, tris, nidx), ctx_.raw(), "upload_mesh_u16");
        return built(tiles, ntiles, stage_mask, erode_radius);
    }
    // rasterize_triangle_soup (rasterization.rs:459-476): nine floats per triangle, no index list.
    Batch build_tiles_soup(const float* verts, size_t ntris, const rcb_config* tiles, size_t ntiles, uint32_t stage_mask = RCB_STAGE_ALL_BUILD,
                           int erode_radius = 0) {
        check(rcb_upload_triangle_soup(ctx_.raw(), verts, ntris), ctx_.raw(), "upload_triangle_soup");
        return built(tiles, ntiles, stage_mask, erode_radius);
    }

   private:
    Batch run(const float* verts, size_t nfloats, const int32_t* tris, size_t nidx, const rcb_config* tiles, size_t ntiles,
              uint32_t stage_mask, int erode_radius) {
        check(rcb_upload_mesh(ctx_.raw(), verts, nfloats, tris, nidx), ctx_.raw(), "upload_mesh");
        return built(tiles, ntiles, stage_mask, erode_radius);
    }
    Batch built(const rcb_config* tiles, size_t ntiles, uint32_t stage_mask, int erode_radius) {
        rcb_result* res = nullptr;
        check(rcb_build_tiles(ctx_.raw(), tiles, (int)ntiles, stage_mask, erode_radius, &res), ctx_.raw(), "build_tiles");
        Batch b(ctx_.raw(), res);
        b.download();
        return b;
    }
    RecastConfig cfg_;
    Context& ctx_;
};

// Several GPUs of one box behind one handle: the multi-tile loop of DynamicNavMesh::build / build_async
// (detour-dynamic/src/dynamic_navmesh.rs:558-684) from ONE caller thread.  The tiles are dealt in groups to two internal
// workers per device (own share first, then stealing); results stay on the device that built them unless download is asked for.
class MultiBatch {
   public:
    MultiBatch(rcb_multi* m, rcb_multi_result* res) : m_(m), res_(res) {}
    MultiBatch(MultiBatch&& o) noexcept : m_(o.m_), res_(std::exchange(o.res_, nullptr)) {}
    MultiBatch(const MultiBatch&) = delete;
    ~MultiBatch() { rcb_multi_result_free(res_); }
    rcb_totals totals(std::vector<int>* tiles_per_device = nullptr) const {
        rcb_totals t{};
        const int n = rcb_multi_device_count(m_);
        if (tiles_per_device) tiles_per_device->assign(n, 0);
        mcheck(rcb_multi_result_totals(res_, &t, tiles_per_device ? tiles_per_device->data() : nullptr, tiles_per_device ? n : 0), "totals");
        return t;
    }
    void download() { mcheck(rcb_multi_result_download(res_), "download"); }
    rcb_tile_view tile(int i) const {  // host pointers; call download() first (or build with download = true)
        rcb_tile_view v{};
        mcheck(rcb_multi_result_tile(res_, i, &v), "tile");
        return v;
    }
    rcb_tile_view tile_device(int i, int* device = nullptr) const {
        rcb_tile_view v{};
        mcheck(rcb_multi_result_tile_device(res_, i, &v, device), "tile_device");
        return v;
    }

   private:
    void mcheck(int st, const char* what) const {
        if (st == RCB_OK) return;
        std::string msg = rcb_multi_last_error(m_);
        if (msg.empty()) msg = what;
        if (st == RCB_EINVALID_MESH) throw InvalidMesh(st, msg);
        if (st == RCB_ENAVMESH_GEN) throw NavMeshGeneration(st, msg);
        throw Error(st, msg);
    }
    rcb_multi* m_;
    rcb_multi_result* res_;
};

class MultiContext {
   public:
    explicit MultiContext(const std::vector<int>& devices) {
        check(rcb_multi_create(devices.data(), (int)devices.size(), &m_), nullptr, "rcb_multi_create failed: no CUDA device (no CPU fallback)");
    }
    ~MultiContext() { rcb_multi_destroy(m_); }
    MultiContext(const MultiContext&) = delete;
    MultiContext& operator=(const MultiContext&) = delete;
    int device_count() const { return rcb_multi_device_count(m_); }
    void upload_mesh(const float* verts, size_t nfloats, const int32_t* tris, size_t nidx) {
        const int st = rcb_multi_upload_mesh(m_, verts, nfloats, tris, nidx);
        if (st != RCB_OK) raise(st, "upload_mesh");
    }
    MultiBatch build_tiles(const rcb_config* tiles, size_t ntiles, uint32_t stage_mask = RCB_STAGE_ALL_BUILD, int erode_radius = 0,
                           int tiles_per_group = 0, bool download = false) {
        rcb_multi_result* res = nullptr;
        const int st = rcb_multi_build_tiles(m_, tiles, (int)ntiles, stage_mask, erode_radius, tiles_per_group, download ? 1 : 0, &res);
        if (st != RCB_OK) raise(st, "build_tiles");
        return MultiBatch(m_, res);
    }

   private:
    [[noreturn]] void raise(int st, const char* what) const {
        std::string msg = rcb_multi_last_error(m_);
        if (msg.empty()) msg = what;
        if (st == RCB_EINVALID_MESH) throw InvalidMesh(st, msg);
        if (st == RCB_ENAVMESH_GEN) throw NavMeshGeneration(st, msg);
        throw Error(st, msg);
    }
    rcb_multi* m_ = nullptr;
};

inline Batch build_from_span_data(Context& ctx, const rcb_config* tiles, size_t ntiles, const uint8_t* data, const uint64_t* tile_offsets,
                                  int format, uint32_t stage_mask, int erode_radius) {
    rcb_result* res = nullptr;
    check(rcb_build_from_span_data(ctx.raw(), tiles, (int)ntiles, data, tile_offsets, format, stage_mask, erode_radius, &res), ctx.raw(),
          "build_from_span_data");
    return Batch(ctx.raw(), res);
}

inline Batch build_tilecache_compressed(Context& ctx, const uint8_t* blobs, const uint64_t* blob_offsets, size_t nlayers,
                                        const uint32_t* first_obstacle, const uint32_t* obstacle_count, const rcb_tc_obstacle* obstacles,
                                        size_t nobstacles, float cs, float ch, uint32_t stage_mask) {
    rcb_result* res = nullptr;
    check(rcb_build_tilecache_compressed(ctx.raw(), blobs, blob_offsets, (int)nlayers, first_obstacle, obstacle_count, obstacles,
                                         (int)nobstacles, cs, ch, stage_mask, 0, &res),
          ctx.raw(), "build_tilecache_compressed");
    return Batch(ctx.raw(), res);
}

// An existing heightfield per tile as host CSR (cell_offset: width*height+1 tile-relative offsets per tile, concatenated;
// spans concatenated in tile order) - the argument of every entry below that starts from spans instead of triangles.
struct SpanInput {
    const uint32_t* cell_offset;
    const int16_t* span_min;
    const int16_t* span_max;
    const uint8_t* span_area;
};

// Heightfield::filter_* (heightfield.rs:286,332,492), CompactHeightfield::build_from_heightfield (compact_heightfield.rs:175),
// erode_walkable_area (area.rs:73), build_distance_field (:514) on existing heightfields; RCB_STAGE_READD re-inserts the spans
// through Heightfield::add_span first (DynamicTileCheckpoint::clone_heightfield, checkpoint.rs:29-60).  areas_override:
// CompactHeightfield.areas as the caller last left them (nullable).
inline Batch build_from_spans(Context& ctx, const rcb_config* tiles, size_t ntiles, const SpanInput& in, const uint8_t* areas_override,
                              uint32_t stage_mask, int erode_radius = 0) {
    rcb_result* res = nullptr;
    check(rcb_build_from_spans(ctx.raw(), tiles, (int)ntiles, in.cell_offset, in.span_min, in.span_max, in.span_area, areas_override, stage_mask,
                               erode_radius, &res),
          ctx.raw(), "build_from_spans");
    return Batch(ctx.raw(), res);
}

// The free function rasterize_triangles(verts, tris, tri_area_ids, n, hf, flag_merge_threshold) (rasterization.rs:405-428) on the
// context's mesh: caller-given triangle areas, no slope classification, no degenerate-triangle skip.
inline Batch rasterize(Context& ctx, const rcb_config* tiles, size_t ntiles, const uint8_t* tri_areas, int32_t flag_merge_threshold,
                       uint32_t stage_mask = RCB_STAGE_RASTER, int erode_radius = 0) {
    rcb_result* res = nullptr;
    check(rcb_rasterize(ctx.raw(), tiles, (int)ntiles, tri_areas, flag_merge_threshold, stage_mask, erode_radius, &res), ctx.raw(), "rasterize");
    return Batch(ctx.raw(), res);
}

// The same onto heightfields that already hold spans (a restored checkpoint that gets its colliders re-applied, dynamic_tile.rs);
// tri_areas == nullptr: areas by slope as RecastBuilder::rasterize_triangles does (lib.rs:186-290; merge threshold 1, :273).
inline Batch rasterize_onto(Context& ctx, const rcb_config* tiles, size_t ntiles, const SpanInput& existing, const uint8_t* tri_areas,
                            int32_t flag_merge_threshold = 1, uint32_t stage_mask = RCB_STAGE_RASTER, int erode_radius = 0) {
    rcb_result* res = nullptr;
    check(rcb_rasterize_onto(ctx.raw(), tiles, (int)ntiles, existing.cell_offset, existing.span_min, existing.span_max, existing.span_area,
                             tri_areas, flag_merge_threshold, stage_mask, erode_radius, &res),
          ctx.raw(), "rasterize_onto");
    return Batch(ctx.raw(), res);
}

// median_filter_walkable_area / mark_box_area / mark_cylinder_area / mark_convex_poly_area (area.rs:238-437, 544-618), applied in
// order to every tile (rcb_area_op; polygon vertices as xyz triples in poly_verts).
inline Batch apply_area_ops(Context& ctx, const rcb_config* tiles, size_t ntiles, const SpanInput& in, const uint8_t* areas_override,
                            const std::vector<rcb_area_op>& ops, const std::vector<float>& poly_verts = {},
                            uint32_t stage_mask = RCB_STAGE_COMPACT, int erode_radius = 0) {
    rcb_result* res = nullptr;
    check(rcb_apply_area_ops(ctx.raw(), tiles, (int)ntiles, in.cell_offset, in.span_min, in.span_max, in.span_area, areas_override, ops.data(),
                             (int)ops.size(), poly_verts.data(), poly_verts.size(), stage_mask, erode_radius, &res),
          ctx.raw(), "apply_area_ops");
    return Batch(ctx.raw(), res);
}

// RecastBuilder::build_heightfield_with_volumes / build_mesh_with_volumes (lib.rs:293-361) on the context's mesh:
// ConvexVolumeSet::apply_to_heightfield (convex_volume.rs:392-430) after the filters, before compaction.
inline Batch build_tiles_with_volumes(Context& ctx, const rcb_config* tiles, size_t ntiles, const std::vector<rcb_convex_volume>& volumes,
                                      const std::vector<float>& poly_verts, uint32_t stage_mask = RCB_STAGE_ALL_BUILD, int erode_radius = 0) {
    rcb_result* res = nullptr;
    check(rcb_build_tiles_with_volumes(ctx.raw(), tiles, (int)ntiles, stage_mask, erode_radius, volumes.data(), (int)volumes.size(),
                                       poly_verts.data(), poly_verts.size(), &res),
          ctx.raw(), "build_tiles_with_volumes");
    return Batch(ctx.raw(), res);
}

// ConvexVolumeSet::apply_to_heightfield on existing heightfields.
inline Batch apply_convex_volumes(Context& ctx, const rcb_config* tiles, size_t ntiles, const SpanInput& in,
                                  const std::vector<rcb_convex_volume>& volumes, const std::vector<float>& poly_verts, uint32_t stage_mask = 0,
                                  int erode_radius = 0) {
    rcb_result* res = nullptr;
    check(rcb_apply_convex_volumes(ctx.raw(), tiles, (int)ntiles, in.cell_offset, in.span_min, in.span_max, in.span_area, volumes.data(),
                                   (int)volumes.size(), poly_verts.data(), poly_verts.size(), stage_mask, erode_radius, &res),
          ctx.raw(), "apply_convex_volumes");
    return Batch(ctx.raw(), res);
}

// TileCacheBuilder::layer_to_compact_heightfield + rasterize_obstacles (detour-tilecache/src/tile_cache_builder.rs:100-251) for a
// batch of decoded layers.
inline Batch build_tilecache_layers(Context& ctx, const std::vector<rcb_tc_layer>& layers, const std::vector<rcb_tc_obstacle>& obstacles, float cs,
                                    float ch, uint32_t stage_mask = RCB_STAGE_COMPACT, int erode_radius = 0) {
    rcb_result* res = nullptr;
    check(rcb_build_tilecache_layers(ctx.raw(), layers.data(), (int)layers.size(), obstacles.data(), (int)obstacles.size(), cs, ch, stage_mask,
                                     erode_radius, &res),
          ctx.raw(), "build_tilecache_layers");
    return Batch(ctx.raw(), res);
}

// TileCache::decompress_tile (tile_cache.rs:854-872) for n compressed tiles; returns the decoded bytes, tile i at
// [offsets[i], offsets[i + 1]).
inline std::vector<uint8_t> decompress_tiles(Context& ctx, const uint8_t* blobs, const uint64_t* blob_offsets, size_t n,
                                             std::vector<uint64_t>* offsets = nullptr) {
    std::vector<uint64_t> off(n + 1, 0);
    check(rcb_decompress_tiles(ctx.raw(), blobs, blob_offsets, (int)n, nullptr, 0, off.data()), ctx.raw(), "decompress_tiles (sizing)");
    std::vector<uint8_t> out(off[n]);
    check(rcb_decompress_tiles(ctx.raw(), blobs, blob_offsets, (int)n, out.data(), out.size(), off.data()), ctx.raw(), "decompress_tiles");
    out.resize(off[n]);  // a block that ends early yields fewer bytes than its size prefix says, as lz4_flex does
    if (offsets) *offsets = off;
    return out;
}

}  // namespace recast_b200
