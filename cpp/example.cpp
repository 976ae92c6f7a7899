// Minimal C++ caller of the drop-in boundary (builds a flat 2-triangle floor into one tile).
// g++ -std=c++17 cpp/example.cpp -Lrecast_rs_b200 -lrecast_b200 -Wl,-rpath,$PWD/recast_rs_b200 -o /tmp/example
#include <cstdio>

#include "recast_b200.hpp"

int main() {
    using namespace recast_b200;
    const float verts[] = {0, 1, 0, 10, 1, 0, 10, 1, 10, 0, 1, 10};
    const int32_t tris[] = {0, 2, 1, 0, 3, 2};
    RecastConfig cfg;
    const float bmin[3] = {0, 0, 0}, bmax[3] = {10, 5, 10};
    cfg.cs = 0.5f;
    cfg.calculate_grid_size(bmin, bmax);
    try {
        cfg.validate();
        Context ctx(0);
        RecastBuilder builder(cfg, ctx);
        Batch b = builder.build_tiles(verts, 12, tris, 6, &cfg, 1);
        rcb_tile_view v = b.tile(0);
        std::printf("cells %u spans %u connections %u max_distance %u\n", v.cell_count, v.span_count, v.conn_count, v.max_distance);
        // the same tile with regions (CompactHeightfield::build_regions) and its VoxelTile span data
        Batch r = builder.build_tiles(verts, 12, tris, 6, &cfg, 1, RCB_STAGE_ALL_BUILD | RCB_STAGE_REGIONS);
        rcb_tile_view w = r.tile(0);
        unsigned nreg = 0;
        for (uint32_t i = 0; i < w.span_count; ++i) nreg = w.reg[i] > nreg ? w.reg[i] : nreg;
        std::printf("regions %u max_regions %u span_data %zu\n", nreg, w.max_regions, r.serialize_spans().size());
    } catch (const Error& e) {
        std::printf("error %d: %s\n", e.status, e.what());  // RCB_ECUDA on a box without a GPU
        return e.status == RCB_ECUDA ? 0 : 1;
    }
    return 0;
}
