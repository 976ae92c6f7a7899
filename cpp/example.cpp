// Minimal C++ caller of the drop-in boundary (builds a flat 2-triangle floor into one tile).
// g++ -std=c++17 cpp/example.cpp -Lrecast_rs_b200 -lrecast_b200 -Wl,-rpath,$PWD/recast_rs_b200 -o /tmp/example
#include <cstdio>
#include <cstring>
#include <vector>

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
        // A floor with a platform over part of it (two spans per column there), cut into 3 x 3 tiles: the host-to-host call
        // with the slim layout, its arrays rebuilt by expand_slim(), must be the bytes of the full layout.
        const float verts2[] = {0, 1, 0, 10, 1, 0, 10, 1, 10, 0, 1, 10, 2, 4, 2, 6, 4, 2, 6, 4, 6, 2, 4, 6};
        const int32_t tris2[] = {0, 2, 1, 0, 3, 2, 4, 6, 5, 4, 7, 6};
        std::vector<rcb_config> tiles;
        for (int tz = 0; tz < 3; ++tz)
            for (int tx = 0; tx < 3; ++tx) {
                RecastConfig c;
                c.cs = 0.25f;
                const float lo[3] = {tx * 3.5f, 0, tz * 3.5f}, hi[3] = {tx * 3.5f + 3.5f, 6, tz * 3.5f + 3.5f};
                c.calculate_grid_size(lo, hi);
                tiles.push_back(c);
            }
        Batch full = builder.build_tiles(verts2, 24, tris2, 12, tiles.data(), tiles.size());
        Batch slim = builder.build_tiles_streamed(verts2, 24, tris2, 12, tiles.data(), tiles.size(), 2, true);
        unsigned long long S = 0, K = 0;
        bool same = true, any_offset = false;
        for (int i = 0; i < (int)tiles.size(); ++i) {
            const rcb_tile_view a = full.tile(i), b = slim.tile(i);
            const SlimTile t = expand_slim(b);
            same = same && a.span_count == b.span_count && a.conn_count == b.conn_count && a.cell_count == b.cell_count;
            if (!same) break;
            const size_t C = a.cell_count, s = a.span_count, k = a.conn_count;
            same = same && !std::memcmp(a.hf_cell_offset, t.hf_cell_offset.data(), 4 * (C + 1)) && !std::memcmp(a.cell_index, t.cell_index.data(), 4 * C) &&
                   !std::memcmp(a.cell_count_arr, t.cell_count.data(), 4 * C) && !std::memcmp(a.first_conn, t.first_conn.data(), 4 * s) &&
                   !std::memcmp(a.conn_ref, t.conn_ref.data(), 4 * k) && !std::memcmp(a.conn_next, t.conn_next.data(), 4 * k) &&
                   !std::memcmp(a.conn_dir, t.conn_dir.data(), k) && !std::memcmp(a.con, t.con.data(), 8 * s) && !std::memcmp(a.areas, t.areas.data(), s) &&
                   !std::memcmp(a.span_min, b.span_min, 2 * s) && !std::memcmp(a.span_max, b.span_max, 2 * s) && !std::memcmp(a.dist, b.dist, 2 * s);
            for (size_t j = 0; b.conn_off8 && j < k; ++j) any_offset = any_offset || b.conn_off8[j] != 0;
            S += s, K += k;
        }
        std::printf("slim expand %s: tiles %zu spans %llu connections %llu second-level offsets %s\n", same ? "ok" : "MISMATCH", tiles.size(), S, K,
                    any_offset ? "yes" : "no");
        // the centre tile again, this time from its filtered spans (Heightfield -> CompactHeightfield -> distance field): same bytes
        {
            const rcb_tile_view a = full.tile(4);
            const SpanInput in{a.hf_cell_offset, a.span_min, a.span_max, a.span_area};
            ctx.profile_enable(true);
            Batch again = build_from_spans(ctx, &tiles[4], 1, in, nullptr, RCB_STAGE_COMPACT | RCB_STAGE_DIST);
            again.download();
            const rcb_tile_view b = again.tile(0);
            const bool ok = a.span_count == b.span_count && a.conn_count == b.conn_count && !std::memcmp(a.conn_ref, b.conn_ref, 4 * (size_t)a.conn_count) &&
                            !std::memcmp(a.dist, b.dist, 2 * (size_t)a.span_count) && !std::memcmp(a.con, b.con, 8 * (size_t)a.span_count);
            std::printf("from spans %s: spans %u connections %u kernels timed %zu\n", ok ? "ok" : "MISMATCH", b.span_count, b.conn_count,
                        ctx.profile_read().size());
            ctx.profile_enable(false);
            same = same && ok;
        }
        // the same tiles through the several-GPU entry, on however many devices the box has listed as {0}
        MultiContext multi({0});
        multi.upload_mesh(verts2, 24, tris2, 12);
        MultiBatch mb = multi.build_tiles(tiles.data(), tiles.size(), RCB_STAGE_ALL_BUILD, 0, 2, true);
        const rcb_totals mt = mb.totals();
        std::printf("multi %s: devices %d tiles %llu spans %llu connections %llu issued without a round trip %llu\n",
                    mt.spans == S && mt.connections == K ? "ok" : "MISMATCH", multi.device_count(), (unsigned long long)mt.tiles,
                    (unsigned long long)mt.spans, (unsigned long long)mt.connections, (unsigned long long)ctx.stats().issued_without_round_trip);
        if (!same || mt.spans != S || mt.connections != K) return 1;
    } catch (const Error& e) {
        std::printf("error %d: %s\n", e.status, e.what());  // RCB_ECUDA on a box without a GPU
        return e.status == RCB_ECUDA ? 0 : 1;
    }
    return 0;
}
