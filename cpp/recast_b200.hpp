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
// Errors: recast_common::Error::{InvalidMesh, NavMeshGeneration} map to the exceptions below.
#pragma once
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

   private:
    Batch run(const float* verts, size_t nfloats, const int32_t* tris, size_t nidx, const rcb_config* tiles, size_t ntiles,
              uint32_t stage_mask, int erode_radius) {
        check(rcb_upload_mesh(ctx_.raw(), verts, nfloats, tris, nidx), ctx_.raw(), "upload_mesh");
        rcb_result* res = nullptr;
        check(rcb_build_tiles(ctx_.raw(), tiles, (int)ntiles, stage_mask, erode_radius, &res), ctx_.raw(), "build_tiles");
        Batch b(ctx_.raw(), res);
        b.download();
        return b;
    }
    RecastConfig cfg_;
    Context& ctx_;
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

}  // namespace recast_b200
