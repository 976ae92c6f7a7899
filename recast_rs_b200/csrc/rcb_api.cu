// rcb_api.cu — C ABI (include/recast_b200.h), context, device memory pools and the stage pipeline.
//
// Host-side counterpart of the reference's RecastBuilder (crates/recast/src/lib.rs:56-359):
//   rcb_build_tiles      = for each tile: build_heightfield (lib.rs:116) [+ filters] + build_from_heightfield
//                          (compact_heightfield.rs:175) [+ erode (area.rs:73)] + build_distance_field (:514)
//   rcb_rasterize        = rasterization.rs:405 rasterize_triangles with explicit areas
//   rcb_build_from_spans = the same stages starting from existing Heightfields
// No CPU fallback exists in this file: every stage is a CUDA kernel launch.
#include <cmath>
#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "rcb_common.cuh"

// ---------------------------------------------------------------------------------------------------
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 4 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) {
            e = cudaMalloc(&p, bytes);
            want = bytes;
        }
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <class T>
    T* as() const {
        return (T*)p;
    }
};

struct Arena {
    char* p = nullptr;
    size_t cap = 0;
};

struct ArenaLayout {  // byte offsets inside arena A (cells + spans) and arena B (connections)
    size_t tinfo, hf_cell_offset, cell_index, cell_count, span_min, span_max, span_area, areas, con, first_conn, dist, reg, conn_mask;
    size_t cell_count16, con8;  // slim layout: the narrow twins of cell_count (u16) and con (4 x u8)
    bool areas_in_slim;         // chf.areas differs from CompactSpan.area (erosion, area operators): part of the slim download
    size_t bytesA;
    size_t slimA;  // the arrays a slim download copies are the first slimA bytes of arena A
    size_t conn_ref, conn_dir, conn_next, conn_ref16, conn_off8;
    size_t bytesB;
    size_t slimB;
    uint64_t Scap, Kcap;  // the span / connection counts the offsets were computed for (>= the real ones)
};

struct HostLocator {
    Locator loc{};
    std::vector<uint32_t> bin_off, bin_tiles;
};

// What the previous build of the same batch found out, so that the next one can be issued without waiting for the
// device: buffer sizes, shared-memory sizes and arena layouts are chosen from these numbers (plus a margin), the
// kernels check them on the device, and a failed check makes the host redo the batch the slow way (BatchHint::valid
// then holds the new numbers).
struct BatchHint {
    bool valid = false;
    uint64_t ntris = 0, nverts = 0;
    uint32_t stage_mask = 0;
    int32_t merge_thr = 0;
    uint64_t ub = 0, nitems = 0, ub_max = 0, S = 0, walk = 0, K = 0;
    uint64_t f_max = 0;      // most fragments one band received (ub_max is the bound it was given room for)
    uint64_t need_b = 0;     // shared memory of k_tile_compact's largest band
    uint64_t pub = 0;        // bytes handed over between the bands of a tile
    uint32_t band_cells = 0;  // how the tiles were cut (the sizes above belong to this cut)
};

// Everything about a batch that depends only on its configurations: validated tile descriptors, the locator grid over
// the tile bounds and their device copies.  Cached per context (keyed on the configuration bytes), so a caller that
// rebuilds the same tiles - every frame of a dynamic nav mesh, every step of a benchmark - does not pay for them again.
struct BatchEntry {
    std::vector<rcb_config> cfgs;  // the key
    std::vector<TileDesc> tiles;
    HostLocator L;
    uint64_t Ctot = 0;
    uint32_t max_cells = 0;
    DevBuf d_tiles, d_bin_off, d_bin_tiles;
    bool have_locator = false, uploaded_tiles = false, uploaded_locator = false;
    // how the tiles are cut into bands for the tile-resident kernels (BandDesc), for the current band_cells
    uint32_t band_cells = 0;
    std::vector<BandDesc> bands;
    std::vector<uint32_t> band_first;  // [ntiles + 1]
    uint32_t max_held_cells = 0, max_bands_per_tile = 0;
    DevBuf d_bands, d_band_first;
    BatchHint hint;
    unsigned long long last_use = 0;
    ~BatchEntry() {
        d_tiles.release();
        d_bin_off.release();
        d_bin_tiles.release();
        d_bands.release();
        d_band_first.release();
    }
};

struct rcb_ctx {
    int device = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    cudaStream_t copy_stream = nullptr;  // D2H of finished tile groups while later groups compute (rcb_build_tiles_streamed)
    std::string err;
    unsigned long long launches = 0;  // kernel launches issued by this context
    // mesh
    DevBuf d_verts, d_tris, d_tri_areas, d_mesh_blocks;
    const float4* mesh_blocks = nullptr;  // xz AABB per 256 consecutive triangles of the current mesh
    const float* verts = nullptr;
    const int32_t* tris = nullptr;
    size_t nfloats = 0, nidx = 0;
    bool mesh_bad = false;  // an index >= nverts (checked on the device when the mesh is set: lib.rs:224-233)
    // per batch
    DevBuf d_tiles, d_bin_off, d_bin_tiles;
    std::vector<std::shared_ptr<BatchEntry>> batch_cache;  // most recent configurations (tile tables, locator, hints)
    unsigned long long use_clock = 0;
    bool no_spec = false;     // RCB_NO_SPEC=1: never issue a batch speculatively (always the synchronous path)
    uint64_t spec_runs = 0, spec_fallbacks = 0;
    std::vector<rcb_result*> live;  // results not yet freed (rcb_destroy releases their arenas)
    // scratch
    DevBuf d_counters, d_cell_cnt, d_frag_off, d_frag_tmp, d_frag_rec, d_fs_mm, d_fs_area, d_scan_tmp, d_hf_off, d_conn_cnt,
        d_conn_off, d_nid, d_t16a, d_t16b, d_t16c, d_t8a, d_t8b, d_item_cnt, d_item_off, d_items, d_tile_ub, d_tile_base, d_tile_cursor,
        d_tile_frags, d_ts_mm, d_ts_area, d_celloff_tmp, d_lookback, d_tc_regons, d_tc_areas, d_tc_obs, d_fidx, d_span_cell, d_wire, d_wire_off,
        d_wire_pos, d_wire_inoff, d_wire_err, d_tc_raw, d_tc_parsed, d_rg_a, d_rg_b, d_binfo, d_pub;
    int tile_threads = 0;        // RCB_TILE_THREADS: 256 / 512 (0: by whether the tiles are cut)
    uint32_t band_cells = 0x7fffffffu;  // RCB_BAND_CELLS: cut tiles into bands of about this many cells for the tile-resident kernels.
                                        // Default: never cut - a tile that does not fit shared memory whole takes the global back
                                        // end, which measured faster than bands on BASELINE configs 3 and 4
    bool phase_timers = false;  // RCB_PHASE_TIMERS=1: per-phase cycle sums of the tile kernels on stderr
    bool force_global = false;  // RCB_FORCE_GLOBAL=1: never use the tile-resident back end (k_tile.cu)
    bool force_v1 = false;  // RCB_FORCE_V1=1: first-generation clipper (k_raster.cu), kept as the wide-tile fallback
    // pinned staging
    char* h_stage = nullptr;
    size_t h_stage_cap = 0;
    unsigned long long* h_counters = nullptr;  // pinned, 32 entries
    unsigned long long* h_slots = nullptr;     // pinned, kSlots x 16: counters of results issued speculatively
    std::vector<int> free_slots;
    // pools
    std::vector<Arena> free_dev, free_host;
    // per-stage timers
    bool prof_on = false;
    struct ProfEv {
        const char* name;
        cudaEvent_t a, b;
    };
    std::vector<ProfEv> prof_pending;
    std::vector<cudaEvent_t> prof_free;
    std::vector<rcb_stage_time> prof_acc;
};

struct rcb_result {
    rcb_ctx* ctx = nullptr;
    int ntiles = 0;
    uint32_t stage_mask = 0;
    std::shared_ptr<BatchEntry> entry;  // run_pipeline results: the tile table lives in the cached batch entry
    std::vector<TileDesc> own_tiles;    // tile-cache results: their own table
    const std::vector<TileDesc>& tiles() const { return entry ? entry->tiles : own_tiles; }
    // a batch issued speculatively (sizes from BatchHint): nothing is known on the host until `done` has passed
    bool pending = false;
    int slot = -1;
    cudaEvent_t done = nullptr;
    int redo_erode = 0;  // what a failed speculation needs to run again (the configurations are entry->cfgs)
    bool slim = false;         // RCB_RESULT_SLIM: conn_mask / conn_ref16 exist, the download copies the slim prefix only
    bool slim_off8 = false;    // conn_off8 is valid (every connection's offset inside the neighbour cell fits 8 bits): the
                               // host copy then holds neither conn_ref16 nor con8
    bool slim_ref16 = false;   // conn_ref16 is valid (every tile-relative span index fits 16 bits)
    bool slim_con8 = false;    // con8 is valid (every offset inside a neighbour cell fits 8 bits)
    bool slim_cnt16 = false;   // cell_count16 is valid (no cell holds 65536+ spans)
    bool slim_host = false;    // the host copy holds only the slim arrays
    bool slim_packed = false;  // k_slim_pack has run on this result
    bool spec = false;         // issued speculatively (pending: the device may still raise a flag)
    int failed = 0;            // a rebuild after a failed speculation failed with this status
    unsigned long long* slim_flag = nullptr;  // pinned: k_slim_pack met a span reference that needs more than 16 bits
    uint64_t copied_bytes = 0;                // bytes the download moved
    bool copy_in_flight = false;
    uint64_t Ctot = 0, S = 0, K = 0, F = 0, ntris = 0, nverts = 0;
    bool have_K = false;
    ArenaLayout lay{};
    Arena devA, devB, hostA, hostB;
    bool downloaded = false, have_tinfo = false;
    bool tile_mode = false;  // connections arena sized by an upper bound (Kcap); exact K comes from tinfo
    uint64_t Kcap = 0;
    std::vector<TileInfo> tinfo;
    // composite result of rcb_build_tiles_streamed: tiles [first[g], first[g+1]) live in children[g]
    std::vector<rcb_result*> children;
    std::vector<int> child_first;
};

namespace {

int fail(rcb_ctx* ctx, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (ctx) ctx->err = buf;
    return code;
}

cudaEvent_t prof_event(rcb_ctx* ctx) {
    cudaEvent_t e = nullptr;
    if (!ctx->prof_free.empty()) {
        e = ctx->prof_free.back();
        ctx->prof_free.pop_back();
    } else {
        cudaEventCreate(&e);
    }
    return e;
}
// g_rcb_prof callback: event pair around every kernel launch
void prof_mark(void* user, const char* name, int begin, cudaStream_t s) {
    rcb_ctx* ctx = (rcb_ctx*)user;
    cudaEvent_t e = prof_event(ctx);
    cudaEventRecord(e, s);
    if (begin)
        ctx->prof_pending.push_back({name, e, nullptr});
    else if (!ctx->prof_pending.empty())
        ctx->prof_pending.back().b = e;
}
struct ProfInstall {  // installs this context's hooks on the calling thread for the duration of one pipeline run
    RcbProfHook prev;
    explicit ProfInstall(rcb_ctx* ctx) : prev(g_rcb_prof) {
        g_rcb_prof = RcbProfHook{ctx->prof_on ? prof_mark : nullptr, ctx, &ctx->launches};
    }
    ~ProfInstall() { g_rcb_prof = prev; }
};

void prof_collect(rcb_ctx* ctx) {
    for (auto& p : ctx->prof_pending) {
        float ms = 0;
        if (!p.b) continue;
        cudaEventSynchronize(p.b);
        cudaEventElapsedTime(&ms, p.a, p.b);
        bool found = false;
        for (auto& acc : ctx->prof_acc)
            if (!strcmp(acc.name, p.name)) {
                acc.ms += ms;
                acc.calls++;
                found = true;
                break;
            }
        if (!found) {
            rcb_stage_time st{};
            strncpy(st.name, p.name, sizeof(st.name) - 1);
            st.ms = ms;
            st.calls = 1;
            ctx->prof_acc.push_back(st);
        }
        ctx->prof_free.push_back(p.a);
        ctx->prof_free.push_back(p.b);
    }
    ctx->prof_pending.clear();
}

#define CU(expr)                                                                                           \
    do {                                                                                                   \
        cudaError_t _e = (expr);                                                                           \
        if (_e != cudaSuccess)                                                                             \
            return fail(ctx, RCB_ECUDA, "CUDA error %s at %s:%d (%s)", cudaGetErrorString(_e), __FILE__, __LINE__, #expr); \
    } while (0)

constexpr uint64_t kTileSmem = 226 * 1024;  // dynamic shared memory budget of the tile-resident kernels

size_t align_up(size_t v, size_t a = 256) { return (v + a - 1) / a * a; }

cudaError_t pool_get(std::vector<Arena>& pool, size_t bytes, bool host, Arena& out) {
    int best = -1;
    for (size_t i = 0; i < pool.size(); ++i)
        if (pool[i].cap >= bytes && pool[i].cap <= 2 * bytes + (1u << 20) && (best < 0 || pool[i].cap < pool[best].cap))
            best = (int)i;
    if (best >= 0) {
        out = pool[best];
        pool.erase(pool.begin() + best);
        return cudaSuccess;
    }
    out.cap = align_up(bytes + bytes / 8 + 256);
    cudaError_t e = host ? cudaMallocHost((void**)&out.p, out.cap) : cudaMalloc((void**)&out.p, out.cap);
    if (e != cudaSuccess) {
        // drop the pool and retry with the exact size
        for (auto& a : pool) host ? cudaFreeHost(a.p) : cudaFree(a.p);
        pool.clear();
        out.cap = align_up(bytes + 256);
        e = host ? cudaMallocHost((void**)&out.p, out.cap) : cudaMalloc((void**)&out.p, out.cap);
    }
    if (e != cudaSuccess) out = Arena{};
    return e;
}

void pool_put(std::vector<Arena>& pool, Arena& a, bool host) {
    if (!a.p) return;
    if (pool.size() >= 64) {  // bound the pool (a streamed build holds 2 arenas per tile group): release the smallest
        size_t m = 0;
        for (size_t i = 1; i < pool.size(); ++i)
            if (pool[i].cap < pool[m].cap) m = i;
        host ? cudaFreeHost(pool[m].p) : cudaFree(pool[m].p);
        pool.erase(pool.begin() + m);
    }
    pool.push_back(a);
    a = Arena{};
}

int stage_upload(rcb_ctx* ctx, DevBuf& dst, const void* src, size_t bytes, size_t& stage_off) {
    if (bytes == 0) return RCB_OK;
    CU(dst.ensure(bytes));
    memcpy(ctx->h_stage + stage_off, src, bytes);
    CU(cudaMemcpyAsync(dst.p, ctx->h_stage + stage_off, bytes, cudaMemcpyHostToDevice, ctx->stream));
    stage_off += align_up(bytes);
    return RCB_OK;
}

int ensure_stage(rcb_ctx* ctx, size_t bytes) {
    if (bytes <= ctx->h_stage_cap) return RCB_OK;
    if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
    ctx->h_stage = nullptr;
    ctx->h_stage_cap = 0;
    CU(cudaMallocHost((void**)&ctx->h_stage, bytes + bytes / 2));
    ctx->h_stage_cap = bytes + bytes / 2;
    return RCB_OK;
}

// `as i16` of an i32 config value (lib.rs:278-287): wraps
inline int32_t as_i16(int32_t v) { return (int32_t)(int16_t)(uint16_t)(uint32_t)v; }

void build_locator(std::vector<TileDesc>& tiles, HostLocator& L) {
    const int n = (int)tiles.size();
    float gx0 = INFINITY, gz0 = INFINITY, gx1 = -INFINITY, gz1 = -INFINITY, bsx = 0, bsz = 0;
    for (auto& t : tiles) {
        gx0 = fminf(gx0, t.bmin[0]);
        gz0 = fminf(gz0, t.bmin[2]);
        gx1 = fmaxf(gx1, t.bmax[0]);
        gz1 = fmaxf(gz1, t.bmax[2]);
        bsx = fmaxf(bsx, t.bmax[0] - t.bmin[0]);
        bsz = fmaxf(bsz, t.bmax[2] - t.bmin[2]);
    }
    // extent for the early reject of whole triangle blocks (k_bin_count): exact tile bounds, or everything
    L.loc.ex0 = gx0, L.loc.ez0 = gz0, L.loc.ex1 = gx1, L.loc.ez1 = gz1;
    if (!std::isfinite(gx0) || !std::isfinite(gz0) || !std::isfinite(gx1) || !std::isfinite(gz1) || n == 0)
        L.loc.ex0 = L.loc.ez0 = -INFINITY, L.loc.ex1 = L.loc.ez1 = INFINITY;
    if (!(bsx > 0) || !std::isfinite(bsx)) bsx = 1.0f;
    if (!(bsz > 0) || !std::isfinite(bsz)) bsz = 1.0f;
    if (!std::isfinite(gx0)) gx0 = 0;
    if (!std::isfinite(gz0)) gz0 = 0;
    double ex = (double)gx1 - gx0, ez = (double)gz1 - gz0;
    if (!std::isfinite(ex) || ex < 0) ex = 0;
    if (!std::isfinite(ez) || ez < 0) ez = 0;
    int nbx = (int)fmin(1024.0, floor(ex / bsx) + 1.0), nbz = (int)fmin(1024.0, floor(ez / bsz) + 1.0);
    if (nbx < 1) nbx = 1;
    if (nbz < 1) nbz = 1;
    // if the cap was hit, enlarge the bins so the grid still spans the extent
    if ((double)nbx * bsx < ex) bsx = (float)(ex / nbx) * 1.0001f;
    if ((double)nbz * bsz < ez) bsz = (float)(ez / nbz) * 1.0001f;
    L.loc.gx0 = gx0;
    L.loc.gz0 = gz0;
    L.loc.inv_bx = 1.0f / bsx;
    L.loc.inv_bz = 1.0f / bsz;
    L.loc.nbx = nbx;
    L.loc.nbz = nbz;
    std::vector<uint32_t> cnt((size_t)nbx * nbz + 1, 0);
    std::vector<int> r(4 * (size_t)n);
    for (int i = 0; i < n; ++i) {
        TileDesc& t = tiles[i];
        int x0 = rcb_bin_coord(t.bmin[0], gx0, L.loc.inv_bx, nbx), x1 = rcb_bin_coord(t.bmax[0], gx0, L.loc.inv_bx, nbx);
        int z0 = rcb_bin_coord(t.bmin[2], gz0, L.loc.inv_bz, nbz), z1 = rcb_bin_coord(t.bmax[2], gz0, L.loc.inv_bz, nbz);
        t.bin_x0 = x0;
        t.bin_z0 = z0;
        r[4 * i] = x0, r[4 * i + 1] = x1, r[4 * i + 2] = z0, r[4 * i + 3] = z1;
        for (int z = z0; z <= z1; ++z)
            for (int x = x0; x <= x1; ++x) cnt[(size_t)z * nbx + x + 1]++;
    }
    for (size_t i = 1; i < cnt.size(); ++i) cnt[i] += cnt[i - 1];
    L.bin_off = cnt;
    L.bin_tiles.assign(cnt.back(), 0);
    std::vector<uint32_t> cur(cnt.begin(), cnt.end() - 1);
    for (int i = 0; i < n; ++i)
        for (int z = r[4 * i + 2]; z <= r[4 * i + 3]; ++z)
            for (int x = r[4 * i]; x <= r[4 * i + 1]; ++x) L.bin_tiles[cur[(size_t)z * nbx + x]++] = (uint32_t)i;
}

// Finds (or creates) the cached entry of a batch of configurations: validation (lib.rs:74 / config.rs:110-136), tile
// descriptors, cell bases.  The key is the configuration bytes themselves (compared in full, most recent first).
int get_batch_entry(rcb_ctx* ctx, const rcb_config* cfgs, int ntiles, std::shared_ptr<BatchEntry>& out) {
    ctx->use_clock++;
    for (auto& e : ctx->batch_cache)
        if ((int)e->cfgs.size() == ntiles && (ntiles == 0 || memcmp(e->cfgs.data(), cfgs, sizeof(rcb_config) * (size_t)ntiles) == 0)) {
            e->last_use = ctx->use_clock;
            out = e;
            return RCB_OK;
        }
    auto e = std::make_shared<BatchEntry>();
    e->tiles.resize(ntiles);
    uint64_t Ctot = 0;
    uint32_t max_cells = 0;
    for (int i = 0; i < ntiles; ++i) {
        const rcb_config& c = cfgs[i];
        const int v = rcb_config_validate(&c);
        if (v != RCB_OK)
            return fail(ctx, v, "tile %d: invalid configuration (width/height <= 0, cs/ch <= 0, slope outside [0,90] or max_verts < 3)", i);
        TileDesc& t = e->tiles[i];
        memset(&t, 0, sizeof t);
        t.width = c.width;
        t.height = c.height;
        for (int k = 0; k < 3; ++k) t.bmin[k] = c.bmin[k], t.bmax[k] = c.bmax[k];
        t.cs = c.cs;
        t.ch = c.ch;
        t.ics = 1.0f / c.cs;
        t.ich = 1.0f / c.ch;
        t.slope_thr = cosf(c.walkable_slope_angle * 3.14159265358979323846f / 180.0f);  // lib.rs:212-213
        t.walkable_height = as_i16(c.walkable_height);
        t.walkable_climb = as_i16(c.walkable_climb);
        t.neg_climb = (int32_t)(int16_t)(-(int16_t)t.walkable_climb);
        t.aux0 = c.border_size, t.aux1 = c.min_region_area, t.aux2 = c.merge_region_area;  // k_regions
        const uint64_t cells = (uint64_t)c.width * (uint64_t)c.height;
        if (cells > 0x7fffffffull || Ctot + cells + ntiles >= 0xffffffffull) return fail(ctx, RCB_EARG, "batch too large: more than 2^32 cells");
        t.cell_base = (uint32_t)Ctot;
        t.wmagic = ((1ull << 40) / (unsigned long long)c.width) + 1;
        Ctot += cells;
        if (cells > max_cells) max_cells = (uint32_t)cells;
    }
    e->cfgs.assign(cfgs, cfgs + ntiles);
    e->Ctot = Ctot;
    e->max_cells = max_cells;
    e->last_use = ctx->use_clock;
    if (ctx->batch_cache.size() >= 96) {  // bounded: drop the least recently used (results that still use it keep it alive)
        size_t m = 0;
        for (size_t i = 1; i < ctx->batch_cache.size(); ++i)
            if (ctx->batch_cache[i]->last_use < ctx->batch_cache[m]->last_use) m = i;
        ctx->batch_cache.erase(ctx->batch_cache.begin() + m);
    }
    ctx->batch_cache.insert(ctx->batch_cache.begin(), e);
    out = e;
    return RCB_OK;
}

// Device copies of an entry's tables (once per entry; the locator only when a rasterizing call needs it).
int upload_batch_entry(rcb_ctx* ctx, BatchEntry& e, bool need_locator) {
    const size_t ntiles = e.tiles.size();
    if (need_locator && !e.have_locator) {
        build_locator(e.tiles, e.L);  // also sets bin_x0 / bin_z0 of the tile descriptors
        e.have_locator = true;
        e.uploaded_tiles = false;
    }
    const bool up_tiles = !e.uploaded_tiles, up_loc = need_locator && !e.uploaded_locator;
    if (!up_tiles && !up_loc) return RCB_OK;
    // these copies read the entry's own host vectors, which outlive the copy (the entry is kept by the result)
    if (up_tiles) {
        CU(e.d_tiles.ensure(sizeof(TileDesc) * ntiles + 64));
        if (ntiles) CU(cudaMemcpyAsync(e.d_tiles.p, e.tiles.data(), sizeof(TileDesc) * ntiles, cudaMemcpyHostToDevice, ctx->stream));
        e.uploaded_tiles = true;
    }
    if (up_loc) {
        CU(e.d_bin_off.ensure(4 * e.L.bin_off.size() + 64));
        CU(e.d_bin_tiles.ensure(4 * e.L.bin_tiles.size() + 64));
        CU(cudaMemcpyAsync(e.d_bin_off.p, e.L.bin_off.data(), 4 * e.L.bin_off.size(), cudaMemcpyHostToDevice, ctx->stream));
        if (!e.L.bin_tiles.empty())
            CU(cudaMemcpyAsync(e.d_bin_tiles.p, e.L.bin_tiles.data(), 4 * e.L.bin_tiles.size(), cudaMemcpyHostToDevice, ctx->stream));
        e.uploaded_locator = true;
    }
    CU(cudaStreamSynchronize(ctx->stream));  // pageable sources; once per entry
    return RCB_OK;
}

// Cuts every tile of the entry into bands of about `band_cells` cells (rcb_band_rows: whole tiles when they are small enough)
// and puts the tables on the device; kept with the entry until another cut is asked for.
int make_bands(rcb_ctx* ctx, BatchEntry& e, uint32_t band_cells) {
    if (e.band_cells == band_cells && !e.band_first.empty()) return RCB_OK;
    const size_t ntiles = e.tiles.size();
    e.bands.clear();
    e.band_first.assign(ntiles + 1, 0);
    e.max_held_cells = 0;
    e.max_bands_per_tile = 0;
    for (size_t t = 0; t < ntiles; ++t) {
        TileDesc& td = e.tiles[t];
        const int R = rcb_band_rows(td.width, td.height, band_cells);
        e.band_first[t] = (uint32_t)e.bands.size();
        td.band_rows = R;
        td.band_first = (uint32_t)e.bands.size();
        td.bmagic = ((1ull << 40) / (unsigned long long)R) + 1;
        uint32_t nb = 0;
        for (int z0 = 0; z0 < td.height; z0 += R, ++nb) {
            BandDesc b;
            b.tile = (uint32_t)t;
            b.z0 = z0;
            b.z1 = std::min(z0 + R, td.height);
            b.hz0 = std::max(z0 - 1, 0);
            b.hz1 = std::min(b.z1 + 1, td.height);
            e.bands.push_back(b);
            e.max_held_cells = std::max(e.max_held_cells, (uint32_t)(td.width * (b.hz1 - b.hz0)));
        }
        e.max_bands_per_tile = std::max(e.max_bands_per_tile, nb);
    }
    e.band_first[ntiles] = (uint32_t)e.bands.size();
    e.band_cells = band_cells;
    CU(e.d_tiles.ensure(sizeof(TileDesc) * ntiles + 64));  // the tile descriptors carry their cut
    if (ntiles) CU(cudaMemcpyAsync(e.d_tiles.p, e.tiles.data(), sizeof(TileDesc) * ntiles, cudaMemcpyHostToDevice, ctx->stream));
    CU(e.d_bands.ensure(sizeof(BandDesc) * e.bands.size() + 64));
    CU(e.d_band_first.ensure(4 * (ntiles + 1) + 64));
    if (!e.bands.empty()) CU(cudaMemcpyAsync(e.d_bands.p, e.bands.data(), sizeof(BandDesc) * e.bands.size(), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(e.d_band_first.p, e.band_first.data(), 4 * (ntiles + 1), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));  // pageable sources; once per cut
    return RCB_OK;
}

// Arena A: per-tile info, cells, spans.  The arrays a slim download copies come first (one copy of a prefix): span geometry
// and areas, dist, reg, and the narrow forms k_slim_pack derives - conn_mask u8[S], cell_count16 u16[C], con8 u8[4S].  What the
// binding layer rebuilds (hf_cell_offset, cell_index, first_conn) or gets in the narrow form (cell_count, con) comes last,
// as does chf.areas while it still equals CompactSpan.area (nothing ran that writes it).
void make_layout(ArenaLayout& l, int ntiles, uint64_t C, uint64_t S, uint32_t stage_mask, bool areas_written = true) {
    size_t o = 0;
    auto put = [&o](size_t bytes) {
        size_t r = o;
        o = align_up(o + bytes);
        return r;
    };
    const bool compact = stage_mask & RCB_STAGE_COMPACT;
    const bool slim = (stage_mask & RCB_RESULT_SLIM) && compact;
    l.areas_in_slim = !slim || areas_written || (stage_mask & RCB_STAGE_ERODE);
    l.tinfo = put(sizeof(TileInfo) * (size_t)ntiles);
    l.span_min = put(2 * S);
    l.span_max = put(2 * S);
    l.span_area = put(S);
    if (l.areas_in_slim) l.areas = put(compact ? S : 0);
    l.dist = put(compact ? 2 * S : 0);
    l.reg = put((stage_mask & RCB_STAGE_REGIONS) ? 2 * S : 0);
    l.conn_mask = put(slim ? S : 0);
    l.cell_count16 = put(slim ? 2 * C : 0);
    l.slimA = o;
    l.con8 = put(slim ? 4 * S : 0);  // travels only when a connection's offset does not fit conn_off8
    if (!l.areas_in_slim) l.areas = put(compact ? S : 0);
    l.cell_index = put(compact ? 4 * C : 0);
    l.cell_count = put(compact ? 4 * C : 0);
    l.con = put(compact ? 8 * S : 0);
    l.hf_cell_offset = put(4 * (C + ntiles));
    l.first_conn = put(compact ? 4 * S : 0);
    l.bytesA = o;
    l.Scap = S;
}
void make_layout_conn(ArenaLayout& l, uint64_t K, bool slim) {
    size_t o = 0;
    auto put = [&o](size_t bytes) {
        size_t r = o;
        o = align_up(o + bytes);
        return r;
    };
    l.conn_off8 = put(slim ? K : 0);
    l.slimB = o;
    l.conn_ref16 = put(slim ? 2 * K : 0);  // travels only when an offset does not fit conn_off8
    l.conn_ref = put(4 * K);
    l.conn_next = put(4 * K);
    l.conn_dir = put(K);
    l.bytesB = o;
    l.Kcap = K;
}

struct Source {  // where the heightfield comes from
    bool raster = false;
    const uint8_t* tri_areas = nullptr;  // host, explicit areas (rcb_rasterize)
    int32_t merge_thr = 1;
    const uint32_t* cell_offset = nullptr;  // host CSR (rcb_build_from_spans)
    const int16_t* smin = nullptr;
    const int16_t* smax = nullptr;
    const uint8_t* sarea = nullptr;
    const uint8_t* areas_override = nullptr;  // host, optional
    bool no_spec = false;                     // the synchronous path, whatever the entry's hint says
    uint32_t band_cells = 0;                  // cut the tiles into bands of about this many cells (0: the default / the hint's)
    bool no_tile_mode = false;                // set when the tile-resident back end reported an overflow
    bool wide_clip = false;                   // set when the 6-vertex clipper overflowed: use the 12-vertex one
    const rcb_area_op* ops = nullptr;         // host, area operators applied after linking (rcb_apply_area_ops)
    int nops = 0;
    const float* poly_verts = nullptr;
    size_t npoly_floats = 0;
    const rcb_convex_volume* vols = nullptr;  // host, convex volumes applied after the filters (rcb_*_with_volumes)
    int nvols = 0;
    const float* vol_verts = nullptr;
    size_t nvol_floats = 0;
    const uint8_t* span_data = nullptr;  // host, serialized heightfields (rcb_build_from_span_data)
    const uint64_t* span_data_off = nullptr;
    int wire_format = 0;
};

// Global-memory back end (v1 kernels) from a heightfield CSR that is already in the result arena:
// filters -> cells/offsets -> link + connection list -> erosion -> distance field.
int global_back_end(rcb_ctx* ctx, rcb_result* res, const TileDesc* d_tiles, TileMap tm, int ntiles, uint64_t Ctot, uint64_t S,
                    uint32_t fmask, uint32_t stage_mask, int erode_radius, const uint8_t* areas_override,
                    const rcb_area_op* ops = nullptr, int nops = 0, const float* poly_verts = nullptr, size_t npoly_floats = 0,
                    const rcb_convex_volume* vols = nullptr, int nvols = 0, const float* vol_verts = nullptr, size_t nvol_floats = 0) {
    cudaStream_t st = ctx->stream;
    uint32_t* d_hf_off = ctx->d_hf_off.as<uint32_t>();
    char* A = res->devA.p;
    TileInfo* d_tinfo = (TileInfo*)(A + res->lay.tinfo);
    int16_t* d_smin = (int16_t*)(A + res->lay.span_min);
    int16_t* d_smax = (int16_t*)(A + res->lay.span_max);
    uint8_t* d_sarea = (uint8_t*)(A + res->lay.span_area);

    // ---- FILTER ------------------------------------------------------------------------------------------
    const bool need_span_cells = S > 0 && (fmask || (stage_mask & RCB_STAGE_COMPACT));
    if (need_span_cells) {
        CU(ctx->d_span_cell.ensure(8 * (S + 1)));  // (tile, cell) per span: every later stage is one thread per span
        launch_span_cells(d_tiles, tm, d_hf_off, d_smin, d_smax, S, ctx->d_span_cell.as<uint2>(), st);
    }
    if (fmask && S > 0) {  // span-parallel, out of place (F1 reads the previous span's original area)
        CU(ctx->d_t8a.ensure(S + 16));
        launch_filter_spans2(d_tiles, S, ctx->d_span_cell.as<uint2>(), d_hf_off, d_smin, d_smax, d_sarea, ctx->d_t8a.as<uint8_t>(), fmask, st);
        CU(cudaMemcpyAsync(d_sarea, ctx->d_t8a.p, S, cudaMemcpyDeviceToDevice, st));
    }

    // ---- convex volumes on the solid heightfield (lib.rs:306, convex_volume.rs:392) -------------------------------
    if (nvols > 0 && S > 0) {
        const size_t vb = align_up(sizeof(rcb_convex_volume) * (size_t)nvols);
        CU(ctx->d_rg_b.ensure(vb + 4 * nvol_floats + 64));
        char* d = (char*)ctx->d_rg_b.p;
        CU(cudaMemcpyAsync(d, vols, sizeof(rcb_convex_volume) * (size_t)nvols, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(d + vb, vol_verts, 4 * nvol_floats, cudaMemcpyHostToDevice, st));
        CU(cudaStreamSynchronize(st));  // pageable host sources
        launch_convex_volumes(d_tiles, tm, d_hf_off, d_smax, d_sarea, (const rcb_convex_volume*)d, nvols, (const float*)(d + vb), st);
    }

    // ---- cells / offsets ---------------------------------------------------------------------------------
    const bool compact = stage_mask & RCB_STAGE_COMPACT;
    launch_tile_bases(d_tiles, ntiles, d_hf_off, nullptr, d_tinfo, st);
    launch_cells_and_offsets(d_tiles, tm, d_hf_off, (uint32_t*)(A + res->lay.hf_cell_offset),
                             compact ? (uint32_t*)(A + res->lay.cell_index) : nullptr,
                             compact ? (uint32_t*)(A + res->lay.cell_count) : nullptr, d_tinfo, ntiles, st);

    // ---- COMPACT: link, connection list ----------------------------------------------------------------------
    if (compact) {
        // span-parallel linking (k_link_spans2): counts and offsets are per SPAN
        CU(ctx->d_conn_cnt.ensure(4 * (S + 1)));
        CU(ctx->d_conn_off.ensure(4 * (S + 1)));
        CU(ctx->d_scan_tmp.ensure(4 * scan_tmp_elems(std::max<uint64_t>(S, Ctot) + 1)));
        CU(ctx->d_nid.ensure(32 * S + 256));  // u32[8][S] absolute neighbour ids
        uint32_t* d_conn_cnt = ctx->d_conn_cnt.as<uint32_t>();
        uint32_t* d_conn_off = ctx->d_conn_off.as<uint32_t>();
        uint32_t* d_nid = ctx->d_nid.as<uint32_t>();
        uint2* d_span_cell = ctx->d_span_cell.as<uint2>();
        uint8_t* d_areas = (uint8_t*)(A + res->lay.areas);
        launch_link_spans2(d_tiles, S, d_span_cell, d_hf_off, d_smin, d_smax, d_sarea, d_areas, d_nid, (uint16_t*)(A + res->lay.con),
                           d_conn_cnt, d_tinfo, st);
        ctx->h_counters[4] = 0;
        if (S > 0) {
            launch_exclusive_scan(d_conn_cnt, d_conn_off, S, ctx->d_scan_tmp.as<uint32_t>(), st);
            CU(cudaMemcpyAsync(&ctx->h_counters[4], d_conn_off + S, 4, cudaMemcpyDeviceToHost, st));
        } else {
            CU(cudaMemsetAsync(d_conn_off, 0, 4, st));
        }
        CU(cudaStreamSynchronize(st));
        const uint64_t K = (uint32_t)ctx->h_counters[4];
        res->K = K;
        res->have_K = true;
        make_layout_conn(res->lay, K, res->slim);
        CU(pool_get(ctx->free_dev, res->lay.bytesB, false, res->devB));
        char* B = res->devB.p;
        launch_tile_conn_bases(ntiles, d_conn_off, d_tinfo, st);
        if (S > 0)
            launch_write_connections2(S, d_span_cell, d_nid, d_conn_off, d_tinfo, (uint32_t*)(A + res->lay.first_conn),
                                      (uint32_t*)(B + res->lay.conn_ref), (uint8_t*)(B + res->lay.conn_dir),
                                      (uint32_t*)(B + res->lay.conn_next), st);
        CU(cudaMemsetAsync(A + res->lay.dist, 0, 2 * S, st));  // compact_heightfield.rs:253
        if (areas_override && S > 0) {
            CU(cudaMemcpyAsync(d_areas, areas_override, S, cudaMemcpyHostToDevice, st));
            CU(cudaStreamSynchronize(st));
        }
        if (nops > 0 && S > 0) {
            // area operators in order: runs of marks are one pass each, a median is a gather into a second buffer
            CU(ctx->d_tc_obs.ensure(sizeof(rcb_area_op) * (size_t)nops + 4 * npoly_floats + 64));
            rcb_area_op* d_ops = ctx->d_tc_obs.as<rcb_area_op>();
            float* d_poly = (float*)(d_ops + nops);
            CU(cudaMemcpyAsync(d_ops, ops, sizeof(rcb_area_op) * (size_t)nops, cudaMemcpyHostToDevice, st));
            if (npoly_floats) CU(cudaMemcpyAsync(d_poly, poly_verts, 4 * npoly_floats, cudaMemcpyHostToDevice, st));
            CU(cudaStreamSynchronize(st));
            CU(ctx->d_t8a.ensure(S));
            int i = 0;
            while (i < nops) {
                if (ops[i].type == RCB_AREA_OP_MEDIAN) {
                    launch_area_median(d_nid, S, d_areas, ctx->d_t8a.as<uint8_t>(), st);
                    CU(cudaMemcpyAsync(d_areas, ctx->d_t8a.p, S, cudaMemcpyDeviceToDevice, st));  // area.rs:291
                    ++i;
                } else {
                    int j = i;
                    while (j < nops && ops[j].type != RCB_AREA_OP_MEDIAN) ++j;
                    launch_area_marks(d_tiles, tm, d_hf_off, d_smin, d_areas, d_ops, i, j, d_poly, st);
                    i = j;
                }
            }
        }
        int max_width = 0;  // the forward sweep runs one thread per column of a tile
        for (const TileDesc& t : res->tiles()) max_width = std::max(max_width, (int)t.width);
        if ((stage_mask & RCB_STAGE_ERODE) && S > 0) {
            CU(ctx->d_t8a.ensure(S + 16));
            CU(ctx->d_t8b.ensure(S + 16));
            CU(ctx->d_fidx.ensure(16 * (S + 1)));
            launch_erode(d_tiles, ntiles, d_hf_off, d_nid, S, d_areas, ctx->d_t8a.as<uint8_t>(), ctx->d_t8b.as<uint8_t>(),
                         erode_radius, ctx->d_fidx.as<uint4>(), max_width, st);
        }
        if ((stage_mask & RCB_STAGE_DIST) && S > 0) {
            CU(ctx->d_t16a.ensure(2 * S + 16));
            CU(ctx->d_t16b.ensure(2 * S + 16));
            CU(ctx->d_t16c.ensure(2 * S + 16));
            CU(ctx->d_fidx.ensure(16 * (S + 1)));
            launch_distance_field(d_tiles, ntiles, d_hf_off, d_nid, d_span_cell, S, d_areas, ctx->d_t16a.as<uint16_t>(),
                                  ctx->d_t16b.as<uint16_t>(), ctx->d_t16c.as<uint16_t>(), (uint16_t*)(A + res->lay.dist), d_tinfo,
                                  ctx->d_fidx.as<uint4>(), max_width, st);
        }
    }
    return RCB_OK;
}

int fetch_tinfo(rcb_result* res, cudaStream_t on = nullptr, bool use_on = false);
int post_counters(rcb_ctx* ctx, rcb_result* res, bool spec);
int resolve_result(rcb_result* res);

// CompactHeightfield::build_regions on a finished result (both back ends): k_regions.cu.
int build_regions_stage(rcb_ctx* ctx, rcb_result* res, const TileDesc* d_tiles, int ntiles, uint64_t S) {
    if (ntiles == 0) return RCB_OK;
    cudaStream_t st = ctx->stream;
    char* A = res->devA.p;
    char* B = res->devB.p;
    const size_t Sp = (size_t)S + 2 * (size_t)ntiles + 8;  // per-region tables: span_count + 2 entries per tile
    size_t o = 0;
    auto put = [&o](size_t bytes) {
        size_t r = o;
        o = align_up(o + bytes);
        return r;
    };
    const size_t o_sdist = put(2 * S + 16), o_reg0 = put(2 * S + 16), o_nb4 = put(16 * S + 16), o_stacks = put(32 * S + 16), o_a = put(4 * S + 16),
                 o_b = put(4 * S + 16);
    CU(ctx->d_rg_a.ensure(o + 256));
    o = 0;
    const size_t o_cnt = put(4 * Sp), o_cadd = put(4 * Sp), o_alive = put(2 * Sp), o_border = put(Sp), o_best = put(8 * Sp), o_target = put(2 * Sp),
                 o_fin = put(2 * Sp), o_err = put(16), o_phase = put(8 * 16);
    CU(ctx->d_rg_b.ensure(o + 256));
    char* a = (char*)ctx->d_rg_a.p;
    char* b = (char*)ctx->d_rg_b.p;
    CU(cudaMemsetAsync(b + o_err, 0xff, 16, st));
    CU(cudaMemsetAsync(b + o_phase, 0, 8 * 16, st));
    RegionArgs ra{};
    ra.tiles = d_tiles;
    ra.tinfo = (TileInfo*)(A + res->lay.tinfo);
    ra.cell_index = (const uint32_t*)(A + res->lay.cell_index);
    ra.cell_count = (const uint32_t*)(A + res->lay.cell_count);
    ra.areas = (const uint8_t*)(A + res->lay.areas);
    ra.con = (const uint16_t*)(A + res->lay.con);
    ra.dist = (const uint16_t*)(A + res->lay.dist);
    ra.first_conn = (const uint32_t*)(A + res->lay.first_conn);
    ra.conn_ref = (const uint32_t*)(B + res->lay.conn_ref);
    ra.conn_dir = (const uint8_t*)(B + res->lay.conn_dir);
    ra.conn_next = (const uint32_t*)(B + res->lay.conn_next);
    ra.reg = (uint16_t*)(A + res->lay.reg);
    ra.Stot = (size_t)S;
    ra.src_dist = (uint16_t*)(a + o_sdist);
    ra.reg0 = (uint16_t*)(a + o_reg0);
    ra.nb4 = (uint32_t*)(a + o_nb4);
    ra.stacks = (uint32_t*)(a + o_stacks);
    ra.bufA = (uint32_t*)(a + o_a);
    ra.bufB = (uint32_t*)(a + o_b);
    ra.r_cnt = (int32_t*)(b + o_cnt);
    ra.r_cadd = (int32_t*)(b + o_cadd);
    ra.r_alive = (uint16_t*)(b + o_alive);
    ra.r_border = (uint8_t*)(b + o_border);
    ra.r_best = (unsigned long long*)(b + o_best);
    ra.r_target = (uint16_t*)(b + o_target);
    ra.r_fin = (uint16_t*)(b + o_fin);
    ra.err = (unsigned int*)(b + o_err);
    ra.phase = ctx->phase_timers ? (unsigned long long*)(b + o_phase) : nullptr;
    {
        // shared-memory residency is decided per launch: it takes the largest tile of the batch to fit
        int r = fetch_tinfo(res, nullptr, false);
        if (r) return r;
        uint32_t smax = 0;
        for (auto& ti : res->tinfo) smax = std::max(smax, ti.span_count);
        // Labels in shared memory make a tile ~1.6x faster, but the kernel ends with its slowest tile and a tile is a long
        // dependent chain: what counts is how many waves of CTAs the batch needs (14 CTAs per SM by registers).
        const size_t bytes = regions_smem_bytes(smax);
        const bool fits = smax < 0xFFFFu && bytes <= 100 * 1024;
        const size_t per_sm_smem = fits ? std::min<size_t>(14, (227 * 1024) / (bytes + 1024)) : 0;
        // relative cost of one tile: 1.0 with labels in shared memory, ~1.6 with labels in global memory (either CTA size)
        const double t_smem = per_sm_smem ? std::ceil((double)ntiles / (148.0 * (double)per_sm_smem)) : 1e30;
        const double t_wide = 1.6 * std::ceil((double)ntiles / (148.0 * 14.0));
        const double t_narrow = 1.6 * std::ceil((double)ntiles / (148.0 * 32.0));
        int mode = t_smem <= t_wide && t_smem <= t_narrow ? 0 : (t_wide <= t_narrow ? 1 : 2);
        if (const char* e = getenv("RCB_REGIONS_MODE")) mode = atoi(e) == 0 && !fits ? 1 : atoi(e);  // 0 smem, 1 wide, 2 narrow
        ra.all_small = smax < 0xFFFFu ? 1u : 0u;
        ra.smem_spans = mode == 0 ? std::max(smax, 1u) : 0u;
        ra.narrow = mode == 2 ? 1 : 0;
        if (mode != 2 && ntiles <= 148 * 2 && smax >= 16384) ra.narrow = 2;  // few, very large tiles: 256 threads each
        if (const char* e = getenv("RCB_REGIONS_WIDE256")) ra.narrow = atoi(e) ? 2 : ra.narrow == 2 ? 0 : ra.narrow;
        if (const char* e = getenv("RCB_REGIONS_BT")) ra.narrow = atoi(e) == 1024 ? 3 : atoi(e) == 256 ? 2 : atoi(e) == 32 ? 1 : 0;
    }
    launch_regions(ra, ntiles, st);
    CU(cudaMemcpyAsync(&ctx->h_counters[7], b + o_err, 4, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    CU(cudaGetLastError());
    if (ctx->phase_timers) {
        unsigned long long ph[9];
        cudaMemcpy(ph, b + o_phase, sizeof ph, cudaMemcpyDeviceToHost);
        fprintf(stderr, "[rcb phase cycles/tile] regions: init %llu sort %llu append %llu expand %llu seed-search %llu flood %llu final-expand %llu merge %llu | slowest tile %llu\n",
                ph[0] / ntiles, ph[1] / ntiles, ph[2] / ntiles, ph[3] / ntiles, ph[4] / ntiles, ph[5] / ntiles, ph[6] / ntiles, ph[7] / ntiles, ph[8]);
    }
    const uint32_t bad = (uint32_t)ctx->h_counters[7];
    if (bad != 0xffffffffu) return fail(ctx, RCB_ERECAST, "tile %u: Region ID overflow", bad);
    return RCB_OK;
}

int run_pipeline(rcb_ctx* ctx, const rcb_config* cfgs, int ntiles, uint32_t stage_mask, int erode_radius,
                 const Source& src, rcb_result** out) {
    if (!ctx || !out || (ntiles > 0 && !cfgs) || ntiles < 0) return fail(ctx, RCB_EARG, "null argument");
    *out = nullptr;
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    ProfInstall prof_install(ctx);
    if (stage_mask & RCB_STAGE_REGIONS) stage_mask |= RCB_STAGE_COMPACT | RCB_STAGE_DIST;  // watershed.rs:375 builds the distance field itself
    if ((stage_mask & (RCB_STAGE_ERODE | RCB_STAGE_DIST)) && !(stage_mask & RCB_STAGE_COMPACT))
        return fail(ctx, RCB_EARG, "ERODE/DIST need RCB_STAGE_COMPACT");
    uint32_t fmask = 0;
    if (stage_mask & RCB_STAGE_FILTER) fmask = 7;
    if (stage_mask & RCB_FILTER_LOW_HANGING) fmask |= 1;
    if (stage_mask & RCB_FILTER_LEDGE) fmask |= 2;
    if (stage_mask & RCB_FILTER_LOW_HEIGHT) fmask |= 4;

    // ---- configs -> tile descriptors (cached per context: SURVEY R2 callers rebuild the same tiles again and again) ----
    std::shared_ptr<BatchEntry> entry;
    {
        const int r = get_batch_entry(ctx, cfgs, ntiles, entry);
        if (r) return r;
    }
    auto* res = new rcb_result();
    res->ctx = ctx;
    res->ntiles = ntiles;
    res->stage_mask = stage_mask;
    res->entry = entry;
    res->slim = (stage_mask & RCB_RESULT_SLIM) && (stage_mask & RCB_STAGE_COMPACT);
    ctx->live.push_back(res);
    const uint64_t Ctot = entry->Ctot;
    const uint32_t max_cells = entry->max_cells;
    res->Ctot = Ctot;
    TileMap tm{(uint32_t)ntiles, (max_cells + 255) / 256, 0};
    tm.nblocks = tm.blocks_per_tile * tm.ntiles;

    auto bail = [&](int code) {
        rcb_result_free(res);
        return code;
    };

    // ---- mesh checks (lib.rs:195-209) ----------------------------------------------------------------
    bool do_raster = src.raster;
    size_t ntris = 0, nverts = 0;
    if (src.raster) {
        if (ctx->nfloats == 0 || ctx->nidx == 0) {
            do_raster = false;  // Ok(()) with untouched heightfields
            if (src.cell_offset) fmask = 0;  // lib.rs:195 returns before the filters: an existing heightfield stays as it is
        } else {
            if (ctx->nfloats % 3 != 0) return bail(fail(ctx, RCB_ENAVMESH_GEN, "Vertex array length must be a multiple of 3"));
            if (ctx->nidx % 3 != 0) return bail(fail(ctx, RCB_ENAVMESH_GEN, "Index array length must be a multiple of 3"));
            ntris = ctx->nidx / 3;
            nverts = ctx->nfloats / 3;
            if (ntris > 0xffffffffull / 3 || nverts > 0x7fffffffull) return bail(fail(ctx, RCB_EARG, "mesh too large"));
        }
    }
    res->ntris = ntris;
    res->nverts = nverts;

    if (do_raster && ntiles > 0 && ctx->mesh_bad)  // lib.rs:224-233 (the indices were checked on the device when the mesh was set)
        return bail(fail(ctx, RCB_ENAVMESH_GEN, "Triangle index out of bounds (max: %zu)", nverts - 1));

    // ---- per-batch tables on the device (cached with the entry) -----------------------------------------
    {
        int r = upload_batch_entry(ctx, *entry, do_raster);
        if (r) return bail(r);
        if (do_raster && src.tri_areas) {
            if ((r = ensure_stage(ctx, align_up(ntris) + 1024))) return bail(r);
            size_t so = 0;
            if ((r = stage_upload(ctx, ctx->d_tri_areas, src.tri_areas, ntris, so))) return bail(r);
        }
    }
    const HostLocator& L = entry->L;
    const TileDesc* d_tiles = entry->d_tiles.as<TileDesc>();
    const std::vector<TileDesc>& tiles = entry->tiles;

#define CUB(expr)                                                                                            \
    do {                                                                                                     \
        cudaError_t _e = (expr);                                                                             \
        if (_e != cudaSuccess)                                                                               \
            return bail(fail(ctx, RCB_ECUDA, "CUDA error %s at %s:%d (%s)", cudaGetErrorString(_e), __FILE__, __LINE__, #expr)); \
    } while (0)

    CUB(ctx->d_hf_off.ensure(4 * (Ctot + 1)));
    CUB(ctx->d_scan_tmp.ensure(4 * scan_tmp_elems(Ctot + 1)));
    uint32_t* d_hf_off = ctx->d_hf_off.as<uint32_t>();
    uint64_t S = 0, F = 0;

    if (src.raster) {
        // ---- RASTER: count -> clip/emit -> scan -> scatter -> replay -> scan -------------------------
        CUB(ctx->d_counters.ensure(48 * sizeof(unsigned long long)));
        CUB(ctx->d_cell_cnt.ensure(4 * (Ctot + 1)));
        CUB(ctx->d_frag_off.ensure(4 * (Ctot + 1)));
        CUB(cudaMemsetAsync(ctx->d_counters.p, 0, 48 * sizeof(unsigned long long), st));
        CUB(cudaMemsetAsync(ctx->d_cell_cnt.p, 0, 4 * (Ctot + 1), st));
        // tile info lives in the arena, which is sized after the span count is known: keep the clip
        // status in scratch until then
        CUB(ctx->d_t16a.ensure(sizeof(TileInfo) * (size_t)ntiles + 256));
        TileInfo* d_tinfo_tmp = ctx->d_t16a.as<TileInfo>();
        CUB(cudaMemsetAsync(d_tinfo_tmp, 0, sizeof(TileInfo) * (size_t)ntiles, st));

        // v2 front end (k_raster2.cu) needs 16-bit cell coordinates and the (tile, chunk) item encoding
        bool v2 = ntiles <= (1 << 20) && !ctx->force_v1 && !src.wide_clip;
        for (int i = 0; i < ntiles && v2; ++i)
            if (tiles[i].width > 65535 || tiles[i].height > 4096 * RCB_CLIP_ROWS ||
                (uint64_t)tiles[i].width * (uint64_t)tiles[i].height > (1ull << 24))  // fragment records carry a 24-bit cell
                v2 = false;

        RasterBuffers rb{};
        rb.tiles = d_tiles;
        rb.ntiles = ntiles;
        rb.loc = L.loc;
        rb.loc.bin_off = entry->d_bin_off.as<uint32_t>();
        rb.loc.bin_tiles = entry->d_bin_tiles.as<uint32_t>();
        rb.mesh.verts = ctx->verts;
        rb.mesh.tris = ctx->tris;
        rb.mesh.ntris = (uint32_t)ntris;
        rb.mesh.nverts = (uint32_t)nverts;
        rb.mesh.tri_areas = src.tri_areas ? ctx->d_tri_areas.as<uint8_t>() : nullptr;
        rb.mesh.merge_thr = src.merge_thr;
        rb.mesh.blocks = ctx->mesh_blocks;
        rb.counters = ctx->d_counters.as<unsigned long long>();
        rb.cell_cnt = ctx->d_cell_cnt.as<uint32_t>();
        Raster2 r2{};
        r2.tiles = rb.tiles;
        r2.ntiles = ntiles;
        r2.loc = rb.loc;
        r2.mesh = rb.mesh;
        r2.counters = rb.counters;
        r2.cell_cnt = rb.cell_cnt;
        r2.flags = ctx->phase_timers ? 0x100u : 0u;
        uint64_t ub = 0, nitems = 0, ub_max = 0;
        bool want_tile = false;
        // tile-resident back end candidates: a band of two rows (plus its halo rows) must fit shared memory; fragment
        // records carry a 24-bit triangle index
        int max_width = 0;
        for (const TileDesc& t : tiles) max_width = std::max(max_width, (int)t.width);
        const bool tile_ok = v2 && !ctx->force_global && !src.no_tile_mode && src.nvols == 0 && !src.cell_offset &&
                             8ull * 4ull * (uint64_t)max_width + 4096 <= kTileSmem / 2 && ntris < (1u << 24);
        // Speculative issue: the previous build of this batch left its sizes in the entry's hint.  With them every
        // buffer, shared-memory size and arena layout can be chosen now, so the whole step is enqueued without a single
        // host round trip; the kernels verify the sizes on the device and rcb_result_* redo the batch if one was short.
        const BatchHint h = entry->hint;
        const bool spec = do_raster && ntiles > 0 && tile_ok && !ctx->no_spec && !src.no_spec && !src.tri_areas && h.valid && h.ub > 0 &&
                          h.ntris == ntris && h.nverts == nverts && h.stage_mask == stage_mask && h.merge_thr == src.merge_thr &&
                          !(stage_mask & RCB_STAGE_REGIONS) && !ctx->phase_timers && (src.band_cells == 0 || src.band_cells == h.band_cells);
        auto margin = [](uint64_t v) { return v + v / 64 + 256; };
        const bool compact = (stage_mask & RCB_STAGE_COMPACT) != 0;
        // How the tiles are cut into bands: whole tiles when they fit (4608 cells: a 64x64 tile and a bit), else equal runs of
        // rows; erosion is a second row-sequential sweep that the bands do not hand over, so it keeps tiles whole.
        uint32_t band_cells = src.band_cells ? src.band_cells : (h.valid && h.band_cells ? h.band_cells : ctx->band_cells);
        if ((stage_mask & RCB_STAGE_ERODE) || !tile_ok) band_cells = 0x7fffffffu;
        int nbands = ntiles;
        auto retry_with = [&](uint64_t need, uint64_t budget) -> int {  // a band needs `need` bytes of `budget`: cut finer, or give up
            Source s2 = src;
            s2.no_spec = true;
            const uint64_t cur = std::min<uint64_t>(band_cells, entry->max_held_cells);
            uint64_t next = need ? cur * budget / need * 7 / 8 : 0;
            if (next >= cur) next = cur * 3 / 4;
            const bool cutting = ctx->band_cells != 0x7fffffffu;  // RCB_BAND_CELLS opted in
            if (!cutting || (stage_mask & RCB_STAGE_ERODE) || next < 4ull * (uint64_t)max_width || (uint64_t)max_width * 2 > next)
                s2.no_tile_mode = true;  // bands of two rows would still not fit: the global-memory back end
            else
                s2.band_cells = (uint32_t)next;
            rcb_result_free(res);
            return run_pipeline(ctx, cfgs, ntiles, stage_mask, erode_radius, s2, out);
        };
        if (do_raster && ntiles > 0) {
            if (v2) {
                {
                    const int r = make_bands(ctx, *entry, band_cells);
                    if (r) return bail(r);
                }
                nbands = (int)entry->bands.size();
                CUB(ctx->d_item_cnt.ensure(4 * (ntris + 2)));
                CUB(ctx->d_item_off.ensure(4 * (ntris + 2)));
                CUB(ctx->d_scan_tmp.ensure(4 * scan_tmp_elems(ntris + 1)));
                r2.item_cnt = ctx->d_item_cnt.as<uint32_t>();
                // per-band fragment bounds -> per-band segments (both back ends consume fragments segment by segment;
                // the global one always with whole-tile bands)
                CUB(ctx->d_tile_ub.ensure(4 * ((size_t)nbands + 1)));
                CUB(ctx->d_tile_base.ensure(4 * ((size_t)nbands + 2)));
                CUB(ctx->d_tile_cursor.ensure(4 * ((size_t)nbands + 1)));
                CUB(cudaMemsetAsync(ctx->d_tile_ub.p, 0, 4 * ((size_t)nbands + 1), st));
                CUB(cudaMemsetAsync(ctx->d_tile_cursor.p, 0, 4 * ((size_t)nbands + 1), st));
                r2.tile_ub = ctx->d_tile_ub.as<uint32_t>();
                launch_bin_count(r2, st);
                // speculative: the fragment / work-item totals must stay within what the buffers were sized for
                launch_tile_segments(r2.tile_ub, ctx->d_tile_base.as<uint32_t>(), nbands, r2.counters, spec ? margin(h.ub) : ~0ull,
                                     spec ? margin(h.nitems) : ~0ull, st);
            } else {
                launch_count_fragments(rb, st);
            }
            if (spec) {
                // shared-memory sizes decide how many CTAs an SM holds: no margin on the per-band maxima (a band that
                // outgrows them raises its flag and the batch is rebuilt with the new maxima)
                ub = margin(h.ub), nitems = margin(h.nitems), ub_max = h.ub_max;
                want_tile = true;
            } else {
                CUB(cudaMemcpyAsync(ctx->h_counters, ctx->d_counters.p, 6 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
                CUB(cudaStreamSynchronize(st));
                if (ctx->h_counters[3]) return bail(fail(ctx, RCB_ENAVMESH_GEN, "Triangle index out of bounds (max: %zu)", nverts - 1));
                ub = ctx->h_counters[0];
                nitems = ctx->h_counters[4];
                ub_max = ctx->h_counters[5];
                if (ub >= 0xfffffff0ull || nitems >= 0xfffffff0ull)
                    return bail(fail(ctx, RCB_EARG, "batch too large: fragment bound %llu", (unsigned long long)ub));
                if (tile_ok) {
                    // the bound counts whole footprints; real fragment counts are about half of it on triangle meshes
                    const uint64_t needA = 8ull * entry->max_held_cells + 8ull * (ub_max / 3) + 4096;
                    want_tile = needA <= kTileSmem && entry->max_bands_per_tile <= 32;
                    if (!want_tile && band_cells != 0x7fffffffu && entry->max_bands_per_tile <= 32) return retry_with(needA, kTileSmem);
                    if (!want_tile && entry->bands.size() != (size_t)ntiles) {  // the global back end wants whole-tile segments
                        Source s2 = src;
                        s2.no_tile_mode = true;
                        rcb_result_free(res);
                        return run_pipeline(ctx, cfgs, ntiles, stage_mask, erode_radius, s2, out);
                    }
                }
            }
        }
        if (want_tile && ub > 0) {
            // ================= tile-resident back end (k_tile.cu) ==============================================
            const uint32_t max_held = entry->max_held_cells;
            const BandDesc* d_bands = entry->d_bands.as<BandDesc>();
            CUB(ctx->d_tile_frags.ensure(12 * ub + 64));
            CUB(ctx->d_ts_mm.ensure(4 * ub + 64));
            CUB(ctx->d_ts_area.ensure(ub + 64));
            CUB(ctx->d_celloff_tmp.ensure(4 * (Ctot + ntiles) + 64));
            // look-back words, the ticket, and the two hand-over flags per band
            const size_t look_bytes = 8 * ((size_t)nbands + 2), flags_bytes = 4 * (size_t)nbands;
            CUB(ctx->d_lookback.ensure(look_bytes + 2 * flags_bytes + 64));
            CUB(ctx->d_binfo.ensure(sizeof(BandInfo) * (size_t)nbands + 64));
            CUB(ctx->d_items.ensure(8 * nitems + 64));
            CUB(cudaMemsetAsync(ctx->d_lookback.p, 0, look_bytes + 2 * flags_bytes, st));
            CUB(cudaMemsetAsync(ctx->d_binfo.p, 0, sizeof(BandInfo) * (size_t)nbands, st));
            BandInfo* d_binfo = ctx->d_binfo.as<BandInfo>();
            r2.tile_cursor = ctx->d_tile_cursor.as<uint32_t>();
            r2.tile_base = ctx->d_tile_base.as<uint32_t>();
            r2.tile_cap = ctx->d_tile_ub.as<uint32_t>();
            r2.tile_frags = ctx->d_tile_frags.as<uint32_t>();
            r2.spec = spec ? 1u : 0u;
            launch_exclusive_scan(r2.item_cnt, ctx->d_item_off.as<uint32_t>(), ntris, ctx->d_scan_tmp.as<uint32_t>(), st);
            launch_fill_items(r2, ctx->d_item_off.as<uint32_t>(), ctx->d_items.as<uint2>(), st);
            launch_clip_items(r2, ctx->d_items.as<uint2>(), (uint32_t)nitems, d_tinfo_tmp, st);
            TileSpansArgs ta{};
            ta.tiles = d_tiles;
            ta.bands = d_bands;
            ta.binfo = d_binfo;
            ta.tile_frags = r2.tile_frags;
            ta.tile_base = r2.tile_base;
            ta.tile_cap = r2.tile_cap;
            ta.tile_cursor = r2.tile_cursor;
            ta.hf_cell_offset = ctx->d_celloff_tmp.as<uint32_t>();
            ta.ts_mm = ctx->d_ts_mm.as<uint32_t>();
            ta.ts_area = ctx->d_ts_area.as<uint8_t>();
            ta.counters = r2.counters;
            ta.fmask = fmask;
            ta.merge_thr = src.merge_thr;
            ta.phase = ctx->phase_timers ? r2.counters + 16 : nullptr;
            // 512 threads, two CTAs per SM (RCB_TILE_THREADS=256: four per SM, for A/B runs)
            const int threads = ctx->tile_threads ? ctx->tile_threads : 512;  // measured: 256 threads / 4 CTAs per SM is slower on both
            const uint64_t ctas_per_sm = 1024 / threads;
            ta.threads = threads;
            if (spec && h.f_max) {
                // the previous build measured the fullest band: its size exactly, one launch
                ta.pass = 0;
                ta.defer = 0;
                ta.smem_bytes = (uint32_t)std::min<uint64_t>(8ull * (max_held + 4) + 8ull * h.f_max + 64, kTileSmem);
                launch_tile_spans(ta, nbands, st);
            } else {
                // exact bound: never overflows when it fits.  Above the budget that lets the SM hold its full complement of
                // CTAs the first launch runs small and defers the few bands that need more to a second, large launch.
                const uint64_t exact = 8ull * (max_held + 4) + 8ull * ub_max + 64;
                const uint64_t small = kTileSmem / ctas_per_sm - 2048;
                ta.pass = 0;
                ta.defer = exact > small;
                ta.smem_bytes = (uint32_t)(exact <= small ? exact : small);
                launch_tile_spans(ta, nbands, st);
                if (ta.defer) {
                    ta.pass = 1;
                    ta.defer = 0;
                    ta.smem_bytes = (uint32_t)(exact <= kTileSmem ? exact : kTileSmem);
                    launch_tile_spans(ta, nbands, st);
                }
            }
            uint64_t walk = 0, needB = 64, Kcap = 0, pub = 0;
            if (spec) {
                S = margin(h.S), walk = margin(h.walk), needB = h.need_b, pub = margin(h.pub);
                Kcap = compact ? (h.K ? margin(h.K) : 8ull * walk) : 0;
                launch_tile_scan(d_bands, d_binfo, nbands, d_tiles, entry->d_band_first.as<uint32_t>(), d_tinfo_tmp, ntiles, r2.counters, S, pub, compact ? 1 : 0, r2.tile_cursor, r2.tile_cap, st);
            } else {
                launch_tile_scan(d_bands, d_binfo, nbands, d_tiles, entry->d_band_first.as<uint32_t>(), d_tinfo_tmp, ntiles, r2.counters, ~0ull, ~0ull, compact ? 1 : 0, r2.tile_cursor, r2.tile_cap, st);
                CUB(cudaMemcpyAsync(ctx->h_counters, ctx->d_counters.p, 12 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
                CUB(cudaStreamSynchronize(st));
                const bool overflow = (ctx->h_counters[2] & 6ull) != 0;
                const bool poly_overflow = (ctx->h_counters[2] & 1ull) != 0;  // a clip polygon outgrew the 6-vertex slots
                S = ctx->h_counters[6];
                walk = ctx->h_counters[7];
                needB = compact ? ctx->h_counters[8] : 64;
                pub = ctx->h_counters[10];
                if (overflow || poly_overflow) {  // redo the batch on the global-memory back end (and the wide clipper)
                    Source s2 = src;
                    s2.no_tile_mode = true;
                    s2.wide_clip = src.wide_clip || poly_overflow;
                    rcb_result_free(res);
                    return run_pipeline(ctx, cfgs, ntiles, stage_mask, erode_radius, s2, out);
                }
                if (needB > kTileSmem) return retry_with(needB, kTileSmem);  // the largest band does not fit: cut finer
                Kcap = compact ? 8ull * walk : 0;
                res->S = S;
                res->F = ctx->h_counters[1];
                // what the next build of this batch may assume (BatchHint); K follows when the result is resolved
                BatchHint nh;
                nh.valid = true;
                nh.ntris = ntris, nh.nverts = nverts, nh.stage_mask = stage_mask, nh.merge_thr = src.merge_thr;
                nh.ub = ub, nh.nitems = nitems, nh.ub_max = ub_max, nh.S = S, nh.walk = walk, nh.need_b = needB, nh.pub = pub;
                nh.f_max = ctx->h_counters[11];
                nh.band_cells = band_cells;
                nh.K = (h.valid && h.S == S && h.walk == walk && h.band_cells == band_cells) ? h.K : 0;
                entry->hint = nh;
            }
            CUB(ctx->d_pub.ensure(pub + 64));
            make_layout(res->lay, ntiles, Ctot, S, stage_mask, src.areas_override != nullptr || src.nops > 0);
            CUB(pool_get(ctx->free_dev, res->lay.bytesA, false, res->devA));
            char* A = res->devA.p;
            TileInfo* d_tinfo = (TileInfo*)(A + res->lay.tinfo);
            CUB(cudaMemcpyAsync(d_tinfo, d_tinfo_tmp, sizeof(TileInfo) * (size_t)ntiles, cudaMemcpyDeviceToDevice, st));
            make_layout_conn(res->lay, Kcap, res->slim);
            if (compact) CUB(pool_get(ctx->free_dev, res->lay.bytesB, false, res->devB));
            char* B = res->devB.p;
            TileCompactArgs tc{};
            tc.tiles = d_tiles;
            tc.bands = d_bands;
            tc.binfo = d_binfo;
            tc.band_first = entry->d_band_first.as<uint32_t>();
            tc.work = r2.tile_frags;
            tc.pub = (unsigned char*)ctx->d_pub.p;
            tc.flag_f = (uint32_t*)((char*)ctx->d_lookback.p + look_bytes);
            tc.flag_d = tc.flag_f + nbands;
            tc.tile_base = r2.tile_base;
            tc.tile_cap = r2.tile_cap;
            tc.hf_cell_offset_in = ctx->d_celloff_tmp.as<uint32_t>();
            tc.ts_mm = ctx->d_ts_mm.as<uint32_t>();
            tc.ts_area = ctx->d_ts_area.as<uint8_t>();
            tc.tinfo = d_tinfo;
            tc.counters = r2.counters;
            tc.lookback = ctx->d_lookback.as<unsigned long long>();
            tc.tile_ticket = (uint32_t*)(ctx->d_lookback.as<unsigned long long>() + nbands);
            tc.hf_cell_offset = (uint32_t*)(A + res->lay.hf_cell_offset);
            tc.smin = (int16_t*)(A + res->lay.span_min);
            tc.smax = (int16_t*)(A + res->lay.span_max);
            tc.sarea = (uint8_t*)(A + res->lay.span_area);
            if (compact) {
                tc.cell_index = (uint32_t*)(A + res->lay.cell_index);
                tc.cell_count = (uint32_t*)(A + res->lay.cell_count);
                tc.areas = (uint8_t*)(A + res->lay.areas);
                tc.con = (uint16_t*)(A + res->lay.con);
                tc.first_conn = (uint32_t*)(A + res->lay.first_conn);
                tc.dist = (uint16_t*)(A + res->lay.dist);
                tc.conn_ref = (uint32_t*)(B + res->lay.conn_ref);
                tc.conn_dir = (uint8_t*)(B + res->lay.conn_dir);
                tc.conn_next = (uint32_t*)(B + res->lay.conn_next);
            }
            tc.Kcap = Kcap;
            tc.do_compact = compact;
            tc.do_erode = (stage_mask & RCB_STAGE_ERODE) != 0;
            tc.do_dist = (stage_mask & RCB_STAGE_DIST) != 0;
            tc.erode_radius = erode_radius;
            tc.phase = ctx->phase_timers ? r2.counters + 16 : nullptr;
            tc.threads = threads;
            tc.smem_bytes = (uint32_t)std::min<uint64_t>(std::max<uint64_t>(needB, 64), kTileSmem);
            launch_tile_compact(tc, nbands, st);
            if (ctx->phase_timers) {
                unsigned long long ph[32];
                cudaMemcpyAsync(ph, r2.counters + 16, sizeof ph, cudaMemcpyDeviceToHost, st);
                cudaStreamSynchronize(st);
                fprintf(stderr, "[rcb phase cycles/CTA] clip: z-chain %llu list %llu x-chain(thread 0) %llu | ", ph[24] / (ph[27] ? ph[27] : 1),
                        ph[25] / (ph[27] ? ph[27] : 1), ph[26] / (ph[27] ? ph[27] : 1));
                fprintf(stderr, "[rcb phase cycles/band] (%d bands) spans:", nbands);
                for (int i = 0; i < 7; ++i) fprintf(stderr, " %llu", ph[i] / (unsigned long long)nbands);
                fprintf(stderr, " | compact:");
                for (int i = 8; i < 22; ++i) fprintf(stderr, " %llu", ph[i] / (unsigned long long)nbands);
                fprintf(stderr, " (load link scan areas dist blur - conn-at-end boundary forward | wait-F wait-D sweep conn-list)");
                fprintf(stderr, "\n");
            }
            res->tile_mode = true;
            res->Kcap = Kcap;
            res->redo_erode = erode_radius;
            CUB(cudaGetLastError());
            // the batch's counters follow the work into a pinned slot; nobody waits here (resolve_result does, later)
            {
                const int r = post_counters(ctx, res, spec);
                if (r) return bail(r);
            }
            if (stage_mask & RCB_STAGE_REGIONS) {
                const int r = build_regions_stage(ctx, res, d_tiles, ntiles, S);
                if (r) return bail(r);
            }
            *out = res;
            return RCB_OK;
        }
        uint32_t* d_frag_off = ctx->d_frag_off.as<uint32_t>();
        // ---- rasterizing onto heightfields that already hold spans (rasterization.rs:405 / lib.rs:186 on a non-empty hf) ----
        const bool onto = src.cell_offset != nullptr;
        uint64_t Sex = 0;
        const uint32_t* d_ex_off = nullptr;
        const int16_t *d_ex_min = nullptr, *d_ex_max = nullptr;
        const uint8_t* d_ex_area = nullptr;
        std::vector<uint32_t> gex;
        if (onto) {
            if (!v2 && do_raster) return bail(fail(ctx, RCB_EARG, "rasterizing onto a non-empty heightfield needs the v2 clipper (tile too wide, or a clip polygon overflowed)"));
            gex.resize(Ctot + 1);
            uint64_t pos = 0;
            for (int i = 0; i < ntiles; ++i) {
                const uint64_t C = (uint64_t)tiles[i].width * tiles[i].height;
                const uint32_t* co = src.cell_offset + pos;
                if (co[0] != 0) return bail(fail(ctx, RCB_EARG, "tile %d: cell_offset[0] != 0", i));
                for (uint64_t c = 0; c < C; ++c) {
                    if (co[c + 1] < co[c]) return bail(fail(ctx, RCB_EARG, "tile %d: cell_offset not ascending", i));
                    gex[tiles[i].cell_base + c] = (uint32_t)(Sex + co[c]);
                }
                Sex += co[C];
                pos += C + 1;
                if (Sex + ub >= 0xfffffff0ull) return bail(fail(ctx, RCB_EARG, "too many spans"));
            }
            gex[Ctot] = (uint32_t)Sex;
            const size_t o_max = align_up(2 * Sex + 16), o_area = 2 * o_max;
            CUB(ctx->d_wire.ensure(o_area + Sex + 64));
            CUB(ctx->d_wire_inoff.ensure(4 * (Ctot + 1)));
            CUB(cudaMemcpyAsync(ctx->d_wire_inoff.p, gex.data(), 4 * (Ctot + 1), cudaMemcpyHostToDevice, st));
            if (Sex) {
                CUB(cudaMemcpyAsync(ctx->d_wire.p, src.smin, 2 * Sex, cudaMemcpyHostToDevice, st));
                CUB(cudaMemcpyAsync((char*)ctx->d_wire.p + o_max, src.smax, 2 * Sex, cudaMemcpyHostToDevice, st));
                CUB(cudaMemcpyAsync((char*)ctx->d_wire.p + o_area, src.sarea, Sex, cudaMemcpyHostToDevice, st));
            }
            CUB(cudaStreamSynchronize(st));  // pageable host sources
            d_ex_off = ctx->d_wire_inoff.as<uint32_t>();
            d_ex_min = (const int16_t*)ctx->d_wire.p;
            d_ex_max = (const int16_t*)((const char*)ctx->d_wire.p + o_max);
            d_ex_area = (const uint8_t*)((const char*)ctx->d_wire.p + o_area);
        }
        const uint64_t cap = ub + Sex;
        if (cap > 0) {
            CUB(ctx->d_frag_rec.ensure(16 * cap));  // fragment records
        }
        void* d_rec = ctx->d_frag_rec.p;
        const bool wide_rec = ntris > (1u << 24) || getenv("RCB_WIDE_RECORDS") != nullptr;  // 8-byte records hold 24-bit triangle indices
        if (ub > 0) {
            if (v2) {
                // fragments go to per-tile segments; each tile then counts and scatters its own columns
                CUB(ctx->d_tile_frags.ensure(12 * ub + 64));
                CUB(ctx->d_items.ensure(8 * nitems + 64));
                r2.tile_cursor = ctx->d_tile_cursor.as<uint32_t>();
                r2.tile_base = ctx->d_tile_base.as<uint32_t>();
                r2.tile_cap = ctx->d_tile_ub.as<uint32_t>();
                r2.tile_frags = ctx->d_tile_frags.as<uint32_t>();
                launch_exclusive_scan(r2.item_cnt, ctx->d_item_off.as<uint32_t>(), ntris, ctx->d_scan_tmp.as<uint32_t>(), st);
                launch_fill_items(r2, ctx->d_item_off.as<uint32_t>(), ctx->d_items.as<uint2>(), st);
                r2.flags |= 0x200u;  // the clipper counts fragments per column as it emits them
                launch_clip_items(r2, ctx->d_items.as<uint2>(), (uint32_t)nitems, d_tinfo_tmp, st);
            } else {
                CUB(ctx->d_frag_tmp.ensure(16 * ub));
                rb.frag_tmp = ctx->d_frag_tmp.as<uint4>();
                rb.frag_cap = (uint32_t)ub;
                launch_clip_emit(rb, d_tinfo_tmp, st);
            }
        }
        if (onto) launch_onto((uint32_t)Ctot, d_ex_off, rb.cell_cnt, 0, nullptr, nullptr, nullptr, nullptr, nullptr, false, st);
        launch_exclusive_scan(rb.cell_cnt, d_frag_off, Ctot, ctx->d_scan_tmp.as<uint32_t>(), st);
        if (onto)
            launch_onto((uint32_t)Ctot, d_ex_off, rb.cell_cnt, 1, d_frag_off, d_ex_min, d_ex_max, d_ex_area, d_rec, wide_rec, st);
        if (ub > 0) {
            if (v2)
                launch_seg_scatter(d_tiles, ntiles, r2.tile_frags, r2.tile_base, r2.tile_cap, r2.tile_cursor, d_frag_off, rb.cell_cnt, d_rec,
                                   wide_rec, rb.counters, (uint32_t)ctx->h_counters[5], st);
            else
                launch_scatter_fragments(rb.frag_tmp, rb.counters + 1, (uint32_t)ub, d_frag_off, d_rec, wide_rec, st);
        }
        if (cap > 0)
            launch_replay_columns((uint32_t)Ctot, d_frag_off, d_rec, wide_rec, rb.cell_cnt, src.merge_thr, st, d_ex_off);
        launch_exclusive_scan(rb.cell_cnt, d_hf_off, Ctot, ctx->d_scan_tmp.as<uint32_t>(), st);
        CUB(cudaMemcpyAsync(&ctx->h_counters[4], d_hf_off + Ctot, 4, cudaMemcpyDeviceToHost, st));
        CUB(cudaMemcpyAsync(&ctx->h_counters[5], ctx->d_counters.as<unsigned long long>() + 1, 8, cudaMemcpyDeviceToHost, st));
        CUB(cudaMemcpyAsync(&ctx->h_counters[6], ctx->d_counters.as<unsigned long long>() + 2, 8, cudaMemcpyDeviceToHost, st));
        CUB(cudaStreamSynchronize(st));
        S = (uint32_t)ctx->h_counters[4];
        F = ctx->h_counters[5];
        if (ctx->h_counters[6] & 2ull) return bail(fail(ctx, RCB_EOVERFLOW, "fragment pool overflow"));
        if ((ctx->h_counters[6] & 1ull) && v2) {  // the fast clipper's 6-vertex polygons overflowed: 12-vertex clipper
            Source s2 = src;
            s2.no_tile_mode = true;
            s2.wide_clip = true;
            rcb_result_free(res);
            return run_pipeline(ctx, cfgs, ntiles, stage_mask, erode_radius, s2, out);
        }
        if (ctx->h_counters[6] & 1ull)
            return bail(fail(ctx, RCB_EOVERFLOW, "clip polygon exceeded %d vertices", RCB_MAX_POLY_VERTS));
        res->F = F;
        res->S = S;
        make_layout(res->lay, ntiles, Ctot, S, stage_mask, src.areas_override != nullptr || src.nops > 0);
        CUB(pool_get(ctx->free_dev, res->lay.bytesA, false, res->devA));
        char* A = res->devA.p;
        CUB(cudaMemcpyAsync(A + res->lay.tinfo, d_tinfo_tmp, sizeof(TileInfo) * (size_t)ntiles, cudaMemcpyDeviceToDevice, st));
        if (S > 0)
            launch_gather_records((uint32_t)Ctot, d_frag_off, d_rec, wide_rec, d_hf_off, (int16_t*)(A + res->lay.span_min),
                                  (int16_t*)(A + res->lay.span_max), (uint8_t*)(A + res->lay.span_area), st);
    } else if (src.span_data || (stage_mask & RCB_STAGE_READD)) {
        // ---- heightfields rebuilt through Heightfield::add_span: wire formats / checkpoint clone (k_wire.cu) ------
        const bool bytes_in = src.span_data != nullptr;
        const int strict = bytes_in && src.wire_format == RCB_WIRE_DYNAMIC_TILE;
        uint64_t Sin = 0, nbytes = 0;
        CUB(ctx->d_wire_inoff.ensure(4 * (Ctot + 1)));
        CUB(ctx->d_wire_err.ensure(64));
        CUB(ctx->d_cell_cnt.ensure(4 * (Ctot + 1)));
        CUB(cudaMemsetAsync(ctx->d_wire_err.p, 0xff, 64, st));
        uint32_t* d_in_off = ctx->d_wire_inoff.as<uint32_t>();
        uint32_t* d_cnt = ctx->d_cell_cnt.as<uint32_t>();
        unsigned long long* d_err = ctx->d_wire_err.as<unsigned long long>();
        std::vector<uint32_t> goff;
        const int16_t *d_in_min = nullptr, *d_in_max = nullptr;
        const uint8_t* d_in_area = nullptr;
        if (bytes_in) {
            for (int i = 0; i < ntiles; ++i)
                if (src.span_data_off[i + 1] < src.span_data_off[i] || src.span_data_off[i + 1] - src.span_data_off[i] >= 0xffff0000ull)
                    return bail(fail(ctx, RCB_EARG, "tile %d: tile_offsets not ascending (or a stream of 4 GiB)", i));
            nbytes = ntiles > 0 ? src.span_data_off[ntiles] : 0;
            Sin = nbytes / 12 + 1;  // no span takes fewer than 12 bytes
            if (Sin >= 0xfffffff0ull) return bail(fail(ctx, RCB_EARG, "too many spans"));
            CUB(ctx->d_wire.ensure(nbytes + 64));
            CUB(ctx->d_wire_off.ensure(8 * ((size_t)ntiles + 1)));
            CUB(ctx->d_wire_pos.ensure(4 * (Ctot + 1)));
            if (nbytes) CUB(cudaMemcpyAsync(ctx->d_wire.p, src.span_data, nbytes, cudaMemcpyHostToDevice, st));
            CUB(cudaMemcpyAsync(ctx->d_wire_off.p, src.span_data_off, 8 * ((size_t)ntiles + 1), cudaMemcpyHostToDevice, st));
            launch_wire_walk(d_tiles, ntiles, ctx->d_wire.as<uint8_t>(), ctx->d_wire_off.as<unsigned long long>(), strict,
                             ctx->d_wire_pos.as<uint32_t>(), d_cnt, d_err, st);
            launch_exclusive_scan(d_cnt, d_in_off, Ctot, ctx->d_scan_tmp.as<uint32_t>(), st);
        } else {
            goff.resize(Ctot + 1);
            uint64_t sbase = 0, pos = 0;
            for (int i = 0; i < ntiles; ++i) {
                const uint64_t C = (uint64_t)tiles[i].width * tiles[i].height;
                const uint32_t* co = src.cell_offset + pos;
                if (co[0] != 0) return bail(fail(ctx, RCB_EARG, "tile %d: cell_offset[0] != 0", i));
                for (uint64_t c = 0; c < C; ++c) {
                    if (co[c + 1] < co[c]) return bail(fail(ctx, RCB_EARG, "tile %d: cell_offset not ascending", i));
                    goff[tiles[i].cell_base + c] = (uint32_t)(sbase + co[c]);
                }
                sbase += co[C];
                pos += C + 1;
                if (sbase >= 0xfffffff0ull) return bail(fail(ctx, RCB_EARG, "too many spans"));
            }
            goff[Ctot] = (uint32_t)sbase;
            Sin = sbase;
            const size_t o_max = align_up(2 * Sin), o_area = 2 * align_up(2 * Sin);
            CUB(ctx->d_wire.ensure(o_area + Sin + 64));
            CUB(cudaMemcpyAsync(d_in_off, goff.data(), 4 * (Ctot + 1), cudaMemcpyHostToDevice, st));
            if (Sin) {
                CUB(cudaMemcpyAsync(ctx->d_wire.p, src.smin, 2 * Sin, cudaMemcpyHostToDevice, st));
                CUB(cudaMemcpyAsync((char*)ctx->d_wire.p + o_max, src.smax, 2 * Sin, cudaMemcpyHostToDevice, st));
                CUB(cudaMemcpyAsync((char*)ctx->d_wire.p + o_area, src.sarea, Sin, cudaMemcpyHostToDevice, st));
            }
            d_in_min = (const int16_t*)ctx->d_wire.p;
            d_in_max = (const int16_t*)((const char*)ctx->d_wire.p + o_max);
            d_in_area = (const uint8_t*)((const char*)ctx->d_wire.p + o_area);
        }
        CUB(ctx->d_fs_mm.ensure(4 * Sin + 64));
        CUB(ctx->d_fs_area.ensure(Sin + 64));
        launch_wire_replay(d_tiles, tm, bytes_in ? ctx->d_wire.as<uint8_t>() : nullptr, ctx->d_wire_off.as<unsigned long long>(),
                           ctx->d_wire_pos.as<uint32_t>(), strict, d_in_min, d_in_max, d_in_area, d_in_off, ctx->d_fs_mm.as<uint32_t>(),
                           ctx->d_fs_area.as<uint8_t>(), d_cnt, d_err, st);
        launch_exclusive_scan(d_cnt, d_hf_off, Ctot, ctx->d_scan_tmp.as<uint32_t>(), st);
        CUB(cudaMemcpyAsync(&ctx->h_counters[0], d_hf_off + Ctot, 4, cudaMemcpyDeviceToHost, st));
        CUB(cudaMemcpyAsync(&ctx->h_counters[1], d_err, 8, cudaMemcpyDeviceToHost, st));
        CUB(cudaStreamSynchronize(st));  // also covers the pageable host sources above
        if (ctx->h_counters[1] != ~0ull) {
            // first failure in the order the reference meets them: tiles one at a time, cells in z-major order,
            // a cell's length checks before its spans are added
            const unsigned long long e = ctx->h_counters[1];
            const unsigned t = (unsigned)(e >> 33), c = (unsigned)((e >> 1) & 0xffffffffu);
            if (e & 1ull) return bail(fail(ctx, RCB_ENAVMESH_GEN, "tile %u cell %u: Invalid span height: min > max", t, c));
            return bail(fail(ctx, RCB_ERECAST, "tile %u cell %u: Invalid span data: insufficient data", t, c));
        }
        S = (uint32_t)ctx->h_counters[0];
        res->S = S;
        make_layout(res->lay, ntiles, Ctot, S, stage_mask, src.areas_override != nullptr || src.nops > 0);
        CUB(pool_get(ctx->free_dev, res->lay.bytesA, false, res->devA));
        char* A = res->devA.p;
        CUB(cudaMemsetAsync(A + res->lay.tinfo, 0, sizeof(TileInfo) * (size_t)ntiles, st));
        if (S > 0)
            launch_gather_spans((uint32_t)Ctot, d_in_off, ctx->d_fs_mm.as<uint32_t>(), ctx->d_fs_area.as<uint8_t>(), d_hf_off,
                                (int16_t*)(A + res->lay.span_min), (int16_t*)(A + res->lay.span_max),
                                (uint8_t*)(A + res->lay.span_area), st);
    } else {
        // ---- existing heightfields (host CSR) -------------------------------------------------------------
        std::vector<uint32_t> goff(Ctot + 1);
        uint64_t sbase = 0, pos = 0;
        for (int i = 0; i < ntiles; ++i) {
            const uint64_t C = (uint64_t)tiles[i].width * tiles[i].height;
            const uint32_t* co = src.cell_offset + pos;
            for (uint64_t c = 0; c < C; ++c) {
                if (co[c + 1] < co[c]) return bail(fail(ctx, RCB_EARG, "tile %d: cell_offset not ascending", i));
                goff[tiles[i].cell_base + c] = (uint32_t)(sbase + co[c]);
            }
            if (co[0] != 0) return bail(fail(ctx, RCB_EARG, "tile %d: cell_offset[0] != 0", i));
            sbase += co[C];
            pos += C + 1;
            if (sbase >= 0xffffffffull) return bail(fail(ctx, RCB_EARG, "too many spans"));
        }
        goff[Ctot] = (uint32_t)sbase;
        S = sbase;
        res->S = S;
        make_layout(res->lay, ntiles, Ctot, S, stage_mask, src.areas_override != nullptr || src.nops > 0);
        CUB(pool_get(ctx->free_dev, res->lay.bytesA, false, res->devA));
        char* A = res->devA.p;
        CUB(cudaMemsetAsync(A + res->lay.tinfo, 0, sizeof(TileInfo) * (size_t)ntiles, st));
        CUB(cudaMemcpyAsync(d_hf_off, goff.data(), 4 * (Ctot + 1), cudaMemcpyHostToDevice, st));
        if (S > 0) {
            CUB(cudaMemcpyAsync(A + res->lay.span_min, src.smin, 2 * S, cudaMemcpyHostToDevice, st));
            CUB(cudaMemcpyAsync(A + res->lay.span_max, src.smax, 2 * S, cudaMemcpyHostToDevice, st));
            CUB(cudaMemcpyAsync(A + res->lay.span_area, src.sarea, S, cudaMemcpyHostToDevice, st));
        }
        CUB(cudaStreamSynchronize(st));  // goff is a stack-owned pageable buffer
    }

    {
        const int r = global_back_end(ctx, res, d_tiles, tm, ntiles, Ctot, S, fmask, stage_mask, erode_radius, src.areas_override, src.ops,
                                      src.nops, src.poly_verts, src.npoly_floats, src.vols, src.nvols, src.vol_verts, src.nvol_floats);
        if (r) return bail(r);
    }
    CUB(cudaGetLastError());
    if (stage_mask & RCB_STAGE_REGIONS) {
        const int r = build_regions_stage(ctx, res, d_tiles, ntiles, S);
        if (r) return bail(r);
    }
    *out = res;
    return RCB_OK;
#undef CUB
}

constexpr int kSlots = 256;     // results whose counters may be in flight at once
constexpr int kSlotWords = 16;  // unsigned long long per slot

// The batch's device counters follow the kernels into a pinned slot of their own, with an event behind them.  Nobody
// waits here: resolve_result does when the host first needs a number.
int post_counters(rcb_ctx* ctx, rcb_result* res, bool spec) {
    if (ctx->free_slots.empty()) {  // more results in flight than slots: wait for this one now
        CU(cudaMemcpyAsync(ctx->h_counters + 16, ctx->d_counters.p, kSlotWords * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        res->slot = -2;  // values are in h_counters + 16 right now
        res->pending = true;
        res->spec = spec;
        return resolve_result(res);
    }
    res->slot = ctx->free_slots.back();
    ctx->free_slots.pop_back();
    CU(cudaMemcpyAsync(ctx->h_slots + (size_t)res->slot * kSlotWords, ctx->d_counters.p, kSlotWords * sizeof(unsigned long long),
                       cudaMemcpyDeviceToHost, ctx->stream));
    if (!res->done) CU(cudaEventCreateWithFlags(&res->done, cudaEventDisableTiming));
    CU(cudaEventRecord(res->done, ctx->stream));
    res->pending = true;
    res->spec = spec;
    if (spec) ctx->spec_runs++;
    return RCB_OK;
}

void release_slot(rcb_result* res) {
    if (res->slot >= 0 && res->ctx) res->ctx->free_slots.push_back(res->slot);
    res->slot = -1;
}

// Takes over everything `from` holds (a rebuilt batch) and leaves it empty.
void adopt_result(rcb_result* res, rcb_result* from) {
    rcb_ctx* ctx = res->ctx;
    pool_put(ctx->free_dev, res->devA, false);
    pool_put(ctx->free_dev, res->devB, false);
    pool_put(ctx->free_host, res->hostA, true);
    pool_put(ctx->free_host, res->hostB, true);
    res->Ctot = from->Ctot, res->S = from->S, res->K = from->K, res->F = from->F, res->ntris = from->ntris, res->nverts = from->nverts;
    res->have_K = from->have_K;
    res->lay = from->lay;
    res->devA = from->devA, res->devB = from->devB, res->hostA = from->hostA, res->hostB = from->hostB;
    from->devA = from->devB = from->hostA = from->hostB = Arena{};
    res->downloaded = from->downloaded, res->have_tinfo = from->have_tinfo, res->tile_mode = from->tile_mode;
    res->Kcap = from->Kcap;
    res->tinfo.swap(from->tinfo);
    res->entry = from->entry;
    res->slim_ref16 = from->slim_ref16, res->slim_host = from->slim_host, res->slim_packed = from->slim_packed;
    res->slim_con8 = from->slim_con8, res->slim_cnt16 = from->slim_cnt16, res->slim_off8 = from->slim_off8;
    res->pending = false;
    res->spec = false;
}

// First host use of a result whose counters are still in flight: wait for them; a speculative batch whose sizes
// turned out too small (any flag raised on the device) is rebuilt here the synchronous way, in place.
int resolve_result(rcb_result* res) {
    if (!res->pending) return res->failed;
    rcb_ctx* ctx = res->ctx;
    const unsigned long long* c = ctx->h_counters + 16;
    if (res->slot >= 0) {
        CU(cudaEventSynchronize(res->done));
        c = ctx->h_slots + (size_t)res->slot * kSlotWords;
    }
    const unsigned long long flags = c[2];
    res->pending = false;
    if (!res->spec || flags == 0) {
        res->S = c[6], res->F = c[1], res->K = c[9];
        res->have_K = true;
        if (res->entry && res->entry->hint.valid && res->entry->hint.stage_mask == res->stage_mask) {
            BatchHint& h = res->entry->hint;  // what the next build of this batch may assume
            if (res->spec) {
                h.ub = c[0], h.nitems = c[4], h.S = c[6], h.walk = c[7];
                h.ub_max = std::max<uint64_t>(h.ub_max, c[5]);
                h.need_b = std::max<uint64_t>(h.need_b, c[8]);
                h.f_max = std::max<uint64_t>(h.f_max, c[11]);
                h.pub = c[10];
                h.K = c[9];
            } else if (h.S == c[6] && h.walk == c[7]) {
                h.K = c[9];
            }
        }
        release_slot(res);
        return RCB_OK;
    }
    release_slot(res);
    ctx->spec_fallbacks++;
    if (res->entry) res->entry->hint.valid = false;
    Source s;
    s.raster = true;
    s.merge_thr = 1;  // lib.rs:273
    s.no_spec = true;
    s.wide_clip = (flags & 1ull) != 0;
    s.no_tile_mode = s.wide_clip;
    rcb_result* again = nullptr;
    const std::shared_ptr<BatchEntry> entry = res->entry;
    const int r = run_pipeline(ctx, entry->cfgs.data(), res->ntiles, res->stage_mask, res->redo_erode, s, &again);
    if (r) {
        res->failed = r;
        return r;
    }
    int r2 = resolve_result(again);
    if (r2) {
        rcb_result_free(again);
        res->failed = r2;
        return r2;
    }
    release_slot(again);
    adopt_result(res, again);
    rcb_result_free(again);
    return RCB_OK;
}

int fetch_tinfo(rcb_result* res, cudaStream_t on, bool use_on) {
    rcb_ctx* ctx = res->ctx;
    {
        const int r = resolve_result(res);
        if (r) return r;
    }
    if (res->have_tinfo) return RCB_OK;
    const cudaStream_t st = use_on ? on : ctx->stream;
    res->tinfo.resize(res->ntiles);
    if (res->ntiles > 0) {
        CU(cudaMemcpyAsync(res->tinfo.data(), res->devA.p + res->lay.tinfo, sizeof(TileInfo) * (size_t)res->ntiles,
                           cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
    }
    res->have_tinfo = true;
    if (res->tile_mode && !res->have_K) {
        uint64_t k = 0;
        for (auto& t : res->tinfo) k += t.conn_count;
        res->K = k;
        res->have_K = true;
    }
    return RCB_OK;
}

}  // namespace

// ===================================================================================================
extern "C" {

int rcb_create(int device, rcb_ctx** out) {
    if (!out) return RCB_EARG;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device < 0 || device >= n) return RCB_ECUDA;
    if (cudaSetDevice(device) != cudaSuccess) return RCB_ECUDA;
    auto* ctx = new rcb_ctx();
    ctx->device = device;
    if (cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaMallocHost((void**)&ctx->h_counters, 32 * sizeof(unsigned long long)) != cudaSuccess ||
        cudaMallocHost((void**)&ctx->h_slots, (size_t)kSlots * kSlotWords * sizeof(unsigned long long)) != cudaSuccess) {
        delete ctx;
        return RCB_ECUDA;
    }
    for (int i = kSlots - 1; i >= 0; --i) ctx->free_slots.push_back(i);
    ctx->stream = ctx->own_stream;
    if (const char* e = getenv("RCB_FORCE_V1")) ctx->force_v1 = e[0] == '1';
    if (const char* e = getenv("RCB_FORCE_GLOBAL")) ctx->force_global = e[0] == '1';
    if (const char* e = getenv("RCB_PHASE_TIMERS")) ctx->phase_timers = e[0] == '1';
    if (const char* e = getenv("RCB_NO_SPEC")) ctx->no_spec = e[0] == '1';
    if (const char* e = getenv("RCB_BAND_CELLS")) ctx->band_cells = (uint32_t)std::max(1, atoi(e));
    if (const char* e = getenv("RCB_TILE_THREADS")) ctx->tile_threads = atoi(e) == 256 ? 256 : atoi(e) == 512 ? 512 : 0;
    *out = ctx;
    return RCB_OK;
}

void rcb_destroy(rcb_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    // results the caller still holds: their arenas go with the context, the handles stay valid for rcb_result_free only
    for (rcb_result* r : ctx->live) {
        for (Arena* a : {&r->devA, &r->devB})
            if (a->p) cudaFree(a->p), *a = Arena{};
        for (Arena* a : {&r->hostA, &r->hostB})
            if (a->p) cudaFreeHost(a->p), *a = Arena{};
        if (r->done) cudaEventDestroy(r->done), r->done = nullptr;
        r->entry.reset();
        r->ctx = nullptr;
    }
    ctx->live.clear();
    ctx->batch_cache.clear();
    DevBuf* bufs[] = {&ctx->d_verts, &ctx->d_tris, &ctx->d_tri_areas, &ctx->d_mesh_blocks, &ctx->d_tiles, &ctx->d_bin_off, &ctx->d_bin_tiles,
                      &ctx->d_counters, &ctx->d_cell_cnt, &ctx->d_frag_off, &ctx->d_frag_tmp, &ctx->d_frag_rec, &ctx->d_fs_mm,
                      &ctx->d_fs_area, &ctx->d_scan_tmp, &ctx->d_hf_off, &ctx->d_conn_cnt, &ctx->d_conn_off, &ctx->d_nid,
                      &ctx->d_t16a, &ctx->d_t16b, &ctx->d_t16c, &ctx->d_t8a, &ctx->d_t8b, &ctx->d_item_cnt, &ctx->d_item_off, &ctx->d_items, &ctx->d_tile_ub, &ctx->d_tile_base,
                      &ctx->d_tile_cursor, &ctx->d_tile_frags, &ctx->d_ts_mm, &ctx->d_ts_area, &ctx->d_celloff_tmp, &ctx->d_lookback, &ctx->d_tc_regons, &ctx->d_tc_areas,
                      &ctx->d_tc_obs, &ctx->d_fidx, &ctx->d_span_cell, &ctx->d_wire, &ctx->d_wire_off, &ctx->d_wire_pos, &ctx->d_wire_inoff, &ctx->d_wire_err,
                      &ctx->d_tc_raw, &ctx->d_tc_parsed, &ctx->d_rg_a, &ctx->d_rg_b, &ctx->d_binfo, &ctx->d_pub};
    for (auto* b : bufs) b->release();
    for (auto& a : ctx->free_dev) cudaFree(a.p);
    for (auto& a : ctx->free_host) cudaFreeHost(a.p);
    if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
    if (ctx->h_counters) cudaFreeHost(ctx->h_counters);
    if (ctx->h_slots) cudaFreeHost(ctx->h_slots);
    prof_collect(ctx);
    for (auto e : ctx->prof_free) cudaEventDestroy(e);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    delete ctx;
}

const char* rcb_last_error(const rcb_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int rcb_set_stream(rcb_ctx* ctx, void* cuda_stream) {
    if (!ctx) return RCB_EARG;
    ctx->stream = (cudaStream_t)cuda_stream;  // verbatim: NULL is the legacy default stream
    return RCB_OK;
}

int rcb_use_own_stream(rcb_ctx* ctx) {
    if (!ctx) return RCB_EARG;
    ctx->stream = ctx->own_stream;
    return RCB_OK;
}

int rcb_profile_enable(rcb_ctx* ctx, int on) {
    if (!ctx) return RCB_EARG;
    ctx->prof_on = on != 0;
    return RCB_OK;
}

int rcb_profile_read(rcb_ctx* ctx, rcb_stage_time* out, int cap) {
    if (!ctx) return RCB_EARG;
    cudaSetDevice(ctx->device);
    prof_collect(ctx);
    const int n = (int)ctx->prof_acc.size();
    for (int i = 0; i < n && i < cap && out; ++i) out[i] = ctx->prof_acc[i];
    return n;
}

int rcb_profile_reset(rcb_ctx* ctx) {
    if (!ctx) return RCB_EARG;
    prof_collect(ctx);
    ctx->prof_acc.clear();
    return RCB_OK;
}

uint64_t rcb_launch_count(const rcb_ctx* ctx) { return ctx ? ctx->launches : 0; }

int rcb_debug_bounds_report(rcb_ctx* ctx, uint64_t out[4], int reset) {
    if (!ctx || !out) return RCB_EARG;
    cudaSetDevice(ctx->device);
    if (reset == 2) {  // self-test: trip one violation on purpose, then report as with reset == 1
        if (ctx->d_counters.ensure(48 * sizeof(unsigned long long)) != cudaSuccess) return RCB_ECUDA;
        rcb_bounds_selftest(ctx->d_counters.as<uint32_t>() + 64, ctx->stream);
        reset = 1;
    }
    cudaDeviceSynchronize();
    out[0] = out[1] = out[2] = out[3] = 0;
    void (*readers[])(unsigned long long*, int) = {rcb_bounds_read_tile, rcb_bounds_read_raster2, rcb_bounds_read_regions};
    for (auto rd : readers) {
        unsigned long long v[4] = {0, 0, 0, 0};
        rd(v, reset);
        if (v[0] && !out[0]) out[1] = v[1], out[2] = v[2], out[3] = v[3];
        out[0] += v[0];
    }
#ifdef RCB_BOUNDS
    return 1;
#else
    return 0;
#endif
}

/* out[0] batches issued speculatively, out[1] of those rebuilt because a size was short, out[2] cached batch entries */
int rcb_ctx_stats(const rcb_ctx* ctx, uint64_t out[4]) {
    if (!ctx || !out) return RCB_EARG;
    out[0] = ctx->spec_runs, out[1] = ctx->spec_fallbacks, out[2] = ctx->batch_cache.size(), out[3] = 0;
    return RCB_OK;
}

// config.rs:53-76
void rcb_config_default(rcb_config* c) {
    memset(c, 0, sizeof *c);
    c->cs = 0.3f;
    c->ch = 0.2f;
    c->walkable_slope_angle = 45.0f;
    c->walkable_height = 2;
    c->walkable_climb = 1;
    c->walkable_radius = 1;
    c->max_edge_len = 12;
    c->max_simplification_error = 1.3f;
    c->min_region_area = 8;
    c->merge_region_area = 20;
    c->max_vertices_per_polygon = 6;
    c->detail_sample_dist = 6.0f;
    c->detail_sample_max_error = 1.0f;
    c->border_size = 0;
}

static int32_t rust_f32_as_i32(float v) {
    if (v != v) return 0;
    if (v >= 2147483648.0f) return INT32_MAX;
    if (v <= -2147483648.0f) return INT32_MIN;
    return (int32_t)v;
}

// config.rs:85-107: fract()==0 -> trunc+1 else ceil
void rcb_config_calculate_grid_size(rcb_config* c, const float bmin[3], const float bmax[3]) {
    for (int i = 0; i < 3; ++i) c->bmin[i] = bmin[i], c->bmax[i] = bmax[i];
    const float wf = (bmax[0] - bmin[0]) / c->cs;
    const float hf = (bmax[2] - bmin[2]) / c->cs;
    c->width = (wf - truncf(wf) == 0.0f) ? rust_f32_as_i32(wf) + 1 : rust_f32_as_i32(ceilf(wf));
    c->height = (hf - truncf(hf) == 0.0f) ? rust_f32_as_i32(hf) + 1 : rust_f32_as_i32(ceilf(hf));
}

// config.rs:110-136
int rcb_config_validate(const rcb_config* c) {
    if (!c) return RCB_EARG;
    if (c->width <= 0 || c->height <= 0) return RCB_EINVALID_MESH;
    if (c->cs <= 0.0f || c->ch <= 0.0f) return RCB_EINVALID_MESH;
    if (c->walkable_slope_angle < 0.0f || c->walkable_slope_angle > 90.0f) return RCB_EINVALID_MESH;
    if (c->max_vertices_per_polygon < 3) return RCB_EINVALID_MESH;
    return RCB_OK;
}

// lib.rs:224-233: the reference fails the build at the first triangle with an index >= vertices.len()/3.  The indices are
// looked at once, here, when the mesh is set (the builds then start without a host round trip for this check); a build over
// a mesh with a bad index fails with RCB_ENAVMESH_GEN before anything is rasterized, which is what the reference's caller sees.
static int check_mesh_indices(rcb_ctx* ctx) {
    ctx->mesh_bad = false;
    ctx->mesh_blocks = nullptr;
    ctx->h_counters[31] = 0;
    if (ctx->nidx && ctx->nfloats && ctx->nidx % 3 == 0 && ctx->nfloats % 3 == 0 && ctx->nidx / 3 <= 0xffffffffull / 3) {
        const uint32_t ntris = (uint32_t)(ctx->nidx / 3);
        const size_t nblocks = ((size_t)ntris + RCB_MESH_BLOCK - 1) / RCB_MESH_BLOCK;
        CU(ctx->d_mesh_blocks.ensure(sizeof(float4) * nblocks + 64));
        launch_mesh_blocks(ctx->verts, ctx->tris, ntris, (uint32_t)std::min<size_t>(ctx->nfloats / 3, 0x7fffffffu), ctx->d_mesh_blocks.as<float4>(),
                           ctx->stream);
        ctx->mesh_blocks = ctx->d_mesh_blocks.as<float4>();
    }
    if (ctx->nidx && ctx->nfloats) {
        CU(ctx->d_counters.ensure(48 * sizeof(unsigned long long)));
        unsigned long long* flag = ctx->d_counters.as<unsigned long long>() + 47;
        CU(cudaMemsetAsync(flag, 0, sizeof(unsigned long long), ctx->stream));
        launch_check_indices(ctx->tris, ctx->nidx - ctx->nidx % 3, (uint32_t)std::min<size_t>(ctx->nfloats / 3, 0x7fffffffu), flag, ctx->stream);
        CU(cudaMemcpyAsync(&ctx->h_counters[31], flag, sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
    }
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaGetLastError());
    ctx->mesh_bad = ctx->h_counters[31] != 0;
    return RCB_OK;
}

int rcb_upload_mesh(rcb_ctx* ctx, const float* verts, size_t nfloats, const int32_t* tris, size_t nidx) {
    if (!ctx || (nfloats && !verts) || (nidx && !tris)) return fail(ctx, RCB_EARG, "null mesh pointer");
    CU(cudaSetDevice(ctx->device));
    CU(ctx->d_verts.ensure(4 * nfloats + 16));
    CU(ctx->d_tris.ensure(4 * nidx + 16));
    if (nfloats) CU(cudaMemcpyAsync(ctx->d_verts.p, verts, 4 * nfloats, cudaMemcpyHostToDevice, ctx->stream));
    if (nidx) CU(cudaMemcpyAsync(ctx->d_tris.p, tris, 4 * nidx, cudaMemcpyHostToDevice, ctx->stream));
    ctx->verts = ctx->d_verts.as<float>();
    ctx->tris = ctx->d_tris.as<int32_t>();
    ctx->nfloats = nfloats;
    ctx->nidx = nidx;
    return check_mesh_indices(ctx);  // synchronizes: the caller's buffers are only borrowed for this call
}

// rasterize_triangles_u16 (rasterization.rs:432): the same mesh with `u16` indices, widened on the device
int rcb_upload_mesh_u16(rcb_ctx* ctx, const float* verts, size_t nfloats, const uint16_t* tris, size_t nidx) {
    if (!ctx || (nfloats && !verts) || (nidx && !tris)) return fail(ctx, RCB_EARG, "null mesh pointer");
    CU(cudaSetDevice(ctx->device));
    CU(ctx->d_verts.ensure(4 * nfloats + 16));
    CU(ctx->d_tris.ensure(4 * nidx + 16));
    CU(ctx->d_tri_areas.ensure(2 * nidx + 16));  // staging of the 16-bit indices
    if (nfloats) CU(cudaMemcpyAsync(ctx->d_verts.p, verts, 4 * nfloats, cudaMemcpyHostToDevice, ctx->stream));
    if (nidx) CU(cudaMemcpyAsync(ctx->d_tri_areas.p, tris, 2 * nidx, cudaMemcpyHostToDevice, ctx->stream));
    launch_widen_indices(ctx->d_tri_areas.as<uint16_t>(), ctx->d_tris.as<int32_t>(), nidx, ctx->stream);
    ctx->verts = ctx->d_verts.as<float>();
    ctx->tris = ctx->d_tris.as<int32_t>();
    ctx->nfloats = nfloats;
    ctx->nidx = nidx;
    return check_mesh_indices(ctx);
}

// rasterize_triangle_soup (rasterization.rs:459): sequential vertices, triangle i = vertices 3 i, 3 i + 1, 3 i + 2
int rcb_upload_triangle_soup(rcb_ctx* ctx, const float* verts, size_t ntris) {
    if (!ctx || (ntris && !verts)) return fail(ctx, RCB_EARG, "null mesh pointer");
    if (ntris > 0x7fffffffull / 3) return fail(ctx, RCB_EARG, "mesh too large");
    CU(cudaSetDevice(ctx->device));
    const size_t nfloats = 9 * ntris, nidx = 3 * ntris;
    CU(ctx->d_verts.ensure(4 * nfloats + 16));
    CU(ctx->d_tris.ensure(4 * nidx + 16));
    if (nfloats) CU(cudaMemcpyAsync(ctx->d_verts.p, verts, 4 * nfloats, cudaMemcpyHostToDevice, ctx->stream));
    launch_widen_indices(nullptr, ctx->d_tris.as<int32_t>(), nidx, ctx->stream);
    ctx->verts = ctx->d_verts.as<float>();
    ctx->tris = ctx->d_tris.as<int32_t>();
    ctx->nfloats = nfloats;
    ctx->nidx = nidx;
    return check_mesh_indices(ctx);
}

int rcb_set_mesh_device(rcb_ctx* ctx, const void* d_verts, size_t nfloats, const void* d_tris, size_t nidx) {
    if (!ctx || (nfloats && !d_verts) || (nidx && !d_tris)) return fail(ctx, RCB_EARG, "null mesh pointer");
    ctx->verts = (const float*)d_verts;
    ctx->tris = (const int32_t*)d_tris;
    ctx->nfloats = nfloats;
    ctx->nidx = nidx;
    CU(cudaSetDevice(ctx->device));
    return check_mesh_indices(ctx);
}

int rcb_build_tiles(rcb_ctx* ctx, const rcb_config* cfgs, int ntiles, uint32_t stage_mask, int erode_radius,
                    rcb_result** out) {
    if (!(stage_mask & RCB_STAGE_RASTER)) return fail(ctx, RCB_EARG, "rcb_build_tiles needs RCB_STAGE_RASTER");
    Source s;
    s.raster = true;
    s.merge_thr = 1;  // lib.rs:273
    return run_pipeline(ctx, cfgs, ntiles, stage_mask, erode_radius, s, out);
}

static int download_on(rcb_result* res, cudaStream_t st, bool wait);
static int finish_download(rcb_result* res, cudaStream_t st);
static int slim_pack(rcb_result* res, cudaStream_t st);
static uint64_t download_bytes(const rcb_result* res);

int rcb_rasterize_onto(rcb_ctx* ctx, const rcb_config* cfgs, int ntiles, const uint32_t* cell_offset, const int16_t* span_min,
                       const int16_t* span_max, const uint8_t* span_area, const uint8_t* tri_areas, int32_t flag_merge_threshold,
                       uint32_t stage_mask, int erode_radius, rcb_result** out) {
    if (ntiles > 0 && !cell_offset) return fail(ctx, RCB_EARG, "cell_offset is null");
    Source s;
    s.raster = true;
    s.tri_areas = tri_areas;
    s.merge_thr = flag_merge_threshold;
    s.cell_offset = cell_offset;
    s.smin = span_min;
    s.smax = span_max;
    s.sarea = span_area;
    return run_pipeline(ctx, cfgs, ntiles, stage_mask | RCB_STAGE_RASTER, erode_radius, s, out);
}

int rcb_build_tiles_streamed(rcb_ctx* ctx, const rcb_config* cfgs, int ntiles, uint32_t stage_mask, int erode_radius, int ngroups,
                             rcb_result** out) {
    if (!ctx || !out || ntiles < 0 || (ntiles > 0 && !cfgs)) return fail(ctx, RCB_EARG, "null argument");
    *out = nullptr;
    if (!(stage_mask & RCB_STAGE_RASTER)) return fail(ctx, RCB_EARG, "rcb_build_tiles_streamed needs RCB_STAGE_RASTER");
    if (ngroups < 1) ngroups = 8;
    if (ngroups > ntiles) ngroups = ntiles > 0 ? ntiles : 1;
    CU(cudaSetDevice(ctx->device));
    if (!ctx->copy_stream) CU(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    Source s;
    s.raster = true;
    s.merge_thr = 1;  // lib.rs:273
    auto* res = new rcb_result();
    res->ctx = ctx;
    res->ntiles = ntiles;
    res->stage_mask = stage_mask;
    ctx->live.push_back(res);
    std::vector<cudaEvent_t> done;
    auto cleanup = [&](int code) {
        cudaStreamSynchronize(ctx->copy_stream);
        for (auto e : done) cudaEventDestroy(e);
        if (code != RCB_OK) {
            rcb_result_free(res);
            res = nullptr;
        }
        return code;
    };
    // Group g computes on the context's stream while the copy stream drains group g-1.  When the groups are issued
    // speculatively (BatchHint) the host never waits inside the loop: every group's kernels and copies are simply queued.
    for (int g = 0; g < ngroups; ++g) {
        const int first = (int)((long long)ntiles * g / ngroups), last = (int)((long long)ntiles * (g + 1) / ngroups);
        rcb_result* child = nullptr;
        int r = run_pipeline(ctx, cfgs + first, last - first, stage_mask, erode_radius, s, &child);
        if (r) return cleanup(r);
        res->children.push_back(child);
        res->child_first.push_back(first);
        // (the slim layout's derived arrays are packed on the COPY stream, by download_on: packed on the compute stream,
        // which already carries H2D + every group's kernels, the build measured 3.2 ms instead of 3.0 on config 2)
        cudaEvent_t e = nullptr;
        if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess || cudaEventRecord(e, ctx->stream) != cudaSuccess)
            return cleanup(fail(ctx, RCB_ECUDA, "CUDA event failure"));
        done.push_back(e);
        if (cudaStreamWaitEvent(ctx->copy_stream, e, 0) != cudaSuccess) return cleanup(fail(ctx, RCB_ECUDA, "cudaStreamWaitEvent"));
        r = download_on(child, ctx->copy_stream, false);
        if (r) return cleanup(r);
    }
    if (cudaStreamSynchronize(ctx->copy_stream) != cudaSuccess) return cleanup(fail(ctx, RCB_ECUDA, "copy stream failed"));
    for (auto* c : res->children) {
        const int r = finish_download(c, ctx->copy_stream);
        if (r) return cleanup(r);
    }
    res->Ctot = 0, res->S = 0;  // sizes live in the children (rcb_result_span_data_size sums them)
    res->downloaded = true;
    res->have_tinfo = true;
    cleanup(RCB_OK);
    *out = res;
    return RCB_OK;
}

int rcb_rasterize(rcb_ctx* ctx, const rcb_config* cfgs, int ntiles, const uint8_t* tri_areas, int32_t flag_merge_threshold,
                  uint32_t stage_mask, int erode_radius, rcb_result** out) {
    if (!tri_areas && ctx && ctx->nidx) return fail(ctx, RCB_EARG, "tri_areas is null");
    Source s;
    s.raster = true;
    s.tri_areas = tri_areas;
    s.merge_thr = flag_merge_threshold;
    return run_pipeline(ctx, cfgs, ntiles, stage_mask | RCB_STAGE_RASTER, erode_radius, s, out);
}

int rcb_build_from_spans(rcb_ctx* ctx, const rcb_config* cfgs, int ntiles, const uint32_t* cell_offset, const int16_t* span_min,
                         const int16_t* span_max, const uint8_t* span_area, const uint8_t* areas_override, uint32_t stage_mask,
                         int erode_radius, rcb_result** out) {
    if (stage_mask & RCB_STAGE_RASTER) return fail(ctx, RCB_EARG, "rcb_build_from_spans: RCB_STAGE_RASTER must not be set");
    if (ntiles > 0 && !cell_offset) return fail(ctx, RCB_EARG, "cell_offset is null");
    Source s;
    s.cell_offset = cell_offset;
    s.smin = span_min;
    s.smax = span_max;
    s.sarea = span_area;
    s.areas_override = areas_override;
    return run_pipeline(ctx, cfgs, ntiles, stage_mask, erode_radius, s, out);
}

// ConvexVolume::new (convex_volume.rs:28-68) for every volume of a set
static int check_volumes(rcb_ctx* ctx, const rcb_convex_volume* vols, int nvols, const float* verts, size_t nfloats) {
    if (nvols < 0 || (nvols > 0 && (!vols || !verts))) return fail(ctx, RCB_EARG, "null argument");
    for (int k = 0; k < nvols; ++k) {
        const rcb_convex_volume& v = vols[k];
        if (3ull * ((uint64_t)v.first_vert + v.num_verts) > nfloats) return fail(ctx, RCB_EARG, "volume %d: vertices out of range", k);
        if (v.num_verts < 3) return fail(ctx, RCB_EINVALID_MESH, "Convex volume requires at least 3 vertices");
        if (v.num_verts > 32) return fail(ctx, RCB_EINVALID_MESH, "Convex volume has too many vertices: %u (max: 32)", v.num_verts);
        if (v.min_height > v.max_height) return fail(ctx, RCB_EINVALID_MESH, "Convex volume min_height > max_height");
        const float* p = verts + 3ull * v.first_vert;
        float sign = 0.0f;  // is_convex (:166-199)
        for (uint32_t i = 0; i < v.num_verts; ++i) {
            const uint32_t j = (i + 1) % v.num_verts, l = (i + 2) % v.num_verts;
            const float dx1 = p[3 * j] - p[3 * i], dz1 = p[3 * j + 2] - p[3 * i + 2];
            const float dx2 = p[3 * l] - p[3 * j], dz2 = p[3 * l + 2] - p[3 * j + 2];
            const volatile float m1 = dx1 * dz2, m2 = dz1 * dx2;  // no contraction: two products, one subtraction
            const float cross = m1 - m2;
            if (cross != 0.0f) {
                if (sign == 0.0f)
                    sign = cross;
                else if ((cross > 0.0f) != (sign > 0.0f))
                    return fail(ctx, RCB_EINVALID_MESH, "Vertices do not form a convex polygon");
            }
        }
    }
    return RCB_OK;
}

int rcb_build_tiles_with_volumes(rcb_ctx* ctx, const rcb_config* cfgs, int ntiles, uint32_t stage_mask, int erode_radius,
                                 const rcb_convex_volume* volumes, int nvolumes, const float* poly_verts, size_t npoly_floats,
                                 rcb_result** out) {
    if (!(stage_mask & RCB_STAGE_RASTER)) return fail(ctx, RCB_EARG, "rcb_build_tiles_with_volumes needs RCB_STAGE_RASTER");
    int r = check_volumes(ctx, volumes, nvolumes, poly_verts, npoly_floats);
    if (r) return r;
    Source s;
    s.raster = true;
    s.merge_thr = 1;  // lib.rs:273
    s.vols = volumes, s.nvols = nvolumes, s.vol_verts = poly_verts, s.nvol_floats = npoly_floats;
    return run_pipeline(ctx, cfgs, ntiles, stage_mask, erode_radius, s, out);
}

int rcb_apply_convex_volumes(rcb_ctx* ctx, const rcb_config* cfgs, int ntiles, const uint32_t* cell_offset, const int16_t* span_min,
                             const int16_t* span_max, const uint8_t* span_area, const rcb_convex_volume* volumes, int nvolumes,
                             const float* poly_verts, size_t npoly_floats, uint32_t stage_mask, int erode_radius, rcb_result** out) {
    if (stage_mask & RCB_STAGE_RASTER) return fail(ctx, RCB_EARG, "rcb_apply_convex_volumes: RCB_STAGE_RASTER must not be set");
    if (ntiles > 0 && !cell_offset) return fail(ctx, RCB_EARG, "cell_offset is null");
    int r = check_volumes(ctx, volumes, nvolumes, poly_verts, npoly_floats);
    if (r) return r;
    Source s;
    s.cell_offset = cell_offset;
    s.smin = span_min;
    s.smax = span_max;
    s.sarea = span_area;
    s.vols = volumes, s.nvols = nvolumes, s.vol_verts = poly_verts, s.nvol_floats = npoly_floats;
    return run_pipeline(ctx, cfgs, ntiles, stage_mask, erode_radius, s, out);
}

int rcb_build_from_span_data(rcb_ctx* ctx, const rcb_config* cfgs, int ntiles, const uint8_t* data, const uint64_t* tile_offsets,
                             int format, uint32_t stage_mask, int erode_radius, rcb_result** out) {
    if (stage_mask & RCB_STAGE_RASTER) return fail(ctx, RCB_EARG, "rcb_build_from_span_data: RCB_STAGE_RASTER must not be set");
    if (format != RCB_WIRE_VOXEL_TILE && format != RCB_WIRE_DYNAMIC_TILE) return fail(ctx, RCB_EARG, "unknown wire format %d", format);
    if (!tile_offsets || (ntiles > 0 && tile_offsets[ntiles] > 0 && !data)) return fail(ctx, RCB_EARG, "null argument");
    static const uint8_t empty = 0;
    Source s;
    s.span_data = data ? data : &empty;
    s.span_data_off = tile_offsets;
    s.wire_format = format;
    return run_pipeline(ctx, cfgs, ntiles, stage_mask & ~(uint32_t)RCB_STAGE_READD, erode_radius, s, out);
}

uint64_t rcb_result_span_data_size(const rcb_result* res) {
    if (!res) return 0;
    if (res->pending && res->ctx) resolve_result(const_cast<rcb_result*>(res));  // the span count of an issued build is on the device
    uint64_t n = 2 * res->Ctot + 12 * res->S;
    for (auto* c : res->children) n += rcb_result_span_data_size(c);
    return n;
}

int rcb_result_serialize_spans(rcb_result* res, int format, uint8_t* out, uint64_t capacity, uint64_t* tile_offsets) {
    if (!res) return RCB_EARG;
    rcb_ctx* ctx = res->ctx;
    if (format != RCB_WIRE_VOXEL_TILE && format != RCB_WIRE_DYNAMIC_TILE) return fail(ctx, RCB_EARG, "unknown wire format %d", format);
    if (!res->children.empty()) return fail(ctx, RCB_EARG, "rcb_result_serialize_spans: not available on a streamed result");
    const uint64_t total = rcb_result_span_data_size(res);
    if (!out || capacity < total) return fail(ctx, RCB_EARG, "rcb_result_serialize_spans: need %llu bytes", (unsigned long long)total);
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    ProfInstall prof_install(ctx);
    int r = fetch_tinfo(res);
    if (r) return r;
    if (tile_offsets) {
        for (int i = 0; i < res->ntiles; ++i) tile_offsets[i] = 2ull * res->tiles()[i].cell_base + 12ull * res->tinfo[i].span_base;
        tile_offsets[res->ntiles] = total;
    }
    if (total == 0) return RCB_OK;
    CU(ctx->d_wire.ensure(total + 64));
    TileMap tm{(uint32_t)res->ntiles, 0, 0};
    uint32_t max_cells = 0;
    for (auto& t : res->tiles()) max_cells = std::max(max_cells, (uint32_t)(t.width * t.height));
    tm.blocks_per_tile = (max_cells + 255) / 256;
    tm.nblocks = tm.blocks_per_tile * tm.ntiles;
    // the tile table of this result may no longer be the one in ctx->d_tiles
    CU(ctx->d_tiles.ensure(sizeof(TileDesc) * (size_t)res->ntiles + 64));
    CU(cudaMemcpyAsync(ctx->d_tiles.p, res->tiles().data(), sizeof(TileDesc) * (size_t)res->ntiles, cudaMemcpyHostToDevice, st));
    const char* A = res->devA.p;
    launch_wire_serialize(ctx->d_tiles.as<TileDesc>(), tm, (const uint32_t*)(A + res->lay.hf_cell_offset),
                          (const TileInfo*)(A + res->lay.tinfo), (const int16_t*)(A + res->lay.span_min),
                          (const int16_t*)(A + res->lay.span_max), (const uint8_t*)(A + res->lay.span_area), ctx->d_wire.as<uint8_t>(),
                          format == RCB_WIRE_VOXEL_TILE, st);
    CU(cudaMemcpyAsync(out, ctx->d_wire.p, total, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    CU(cudaGetLastError());
    return RCB_OK;
}

// Tile-cache layers whose regons / areas / tile table / obstacles are already on the device (d_tc_regons,
// d_tc_areas, d_tiles, d_tc_obs): spans, connections, later stages, obstacles.  Frees res on failure.
static int tilecache_back_end(rcb_ctx* ctx, rcb_result* res, TileMap tm, int nlayers, uint64_t Ctot, int nobstacles, uint32_t stage_mask,
                              int erode_radius, rcb_result** out) {
    cudaStream_t st = ctx->stream;
    auto bail = [&](int code) {
        rcb_result_free(res);
        return code;
    };
#define CUB(expr)                                                                                            \
    do {                                                                                                     \
        cudaError_t _e = (expr);                                                                             \
        if (_e != cudaSuccess)                                                                               \
            return bail(fail(ctx, RCB_ECUDA, "CUDA error %s at %s:%d (%s)", cudaGetErrorString(_e), __FILE__, __LINE__, #expr)); \
    } while (0)
    const TileDesc* d_tiles = ctx->d_tiles.as<TileDesc>();
    CUB(ctx->d_hf_off.ensure(4 * (Ctot + 1)));
    CUB(ctx->d_cell_cnt.ensure(4 * (Ctot + 1)));
    CUB(ctx->d_scan_tmp.ensure(4 * scan_tmp_elems(Ctot + 1)));
    uint32_t* d_hf_off = ctx->d_hf_off.as<uint32_t>();
    launch_tc_count(d_tiles, tm, ctx->d_tc_regons.as<uint8_t>(), ctx->d_cell_cnt.as<uint32_t>(), st);
    launch_exclusive_scan(ctx->d_cell_cnt.as<uint32_t>(), d_hf_off, Ctot, ctx->d_scan_tmp.as<uint32_t>(), st);
    CUB(cudaMemcpyAsync(&ctx->h_counters[4], d_hf_off + Ctot, 4, cudaMemcpyDeviceToHost, st));
    CUB(cudaStreamSynchronize(st));
    const uint64_t S = (uint32_t)ctx->h_counters[4];
    res->S = S;
    make_layout(res->lay, nlayers, Ctot, S, stage_mask);
    CUB(pool_get(ctx->free_dev, res->lay.bytesA, false, res->devA));
    char* A = res->devA.p;
    CUB(cudaMemsetAsync(A + res->lay.tinfo, 0, sizeof(TileInfo) * (size_t)nlayers, st));
    int16_t* d_smin = (int16_t*)(A + res->lay.span_min);
    int16_t* d_smax = (int16_t*)(A + res->lay.span_max);
    uint8_t* d_sarea = (uint8_t*)(A + res->lay.span_area);
    if (S > 0)
        launch_tc_spans(d_tiles, tm, ctx->d_tc_regons.as<uint8_t>(), ctx->d_tc_areas.as<uint8_t>(), d_hf_off, d_smin, d_smax, d_sarea, st);
    {
        const int r = global_back_end(ctx, res, d_tiles, tm, nlayers, Ctot, S, 0, stage_mask, erode_radius, nullptr);
        if (r) return bail(r);
    }
    // obstacles come AFTER the connections exist and only clear CompactSpan.area (:65-68, :309)
    if (nobstacles > 0 && S > 0)
        launch_tc_obstacles(d_tiles, nlayers, ctx->d_tc_obs.as<rcb_tc_obstacle>(), d_hf_off, d_smin, d_smax, d_sarea, st);
    CUB(cudaGetLastError());
    *out = res;
    return RCB_OK;
#undef CUB
}

int rcb_build_tilecache_layers(rcb_ctx* ctx, const rcb_tc_layer* layers, int nlayers, const rcb_tc_obstacle* obstacles, int nobstacles,
                               float cs, float ch, uint32_t stage_mask, int erode_radius, rcb_result** out) {
    if (!ctx || !out || nlayers < 0 || (nlayers > 0 && !layers) || nobstacles < 0 || (nobstacles > 0 && !obstacles))
        return fail(ctx, RCB_EARG, "null argument");
    *out = nullptr;
    if (!(cs > 0.0f) || !(ch > 0.0f)) return fail(ctx, RCB_EINVALID_MESH, "Invalid cell size or height");
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    ProfInstall prof_install(ctx);
    stage_mask = (stage_mask | RCB_STAGE_COMPACT) & ~(uint32_t)(RCB_STAGE_RASTER | RCB_STAGE_FILTER | 0x700u);
    auto* res = new rcb_result();
    res->ctx = ctx;
    res->ntiles = nlayers;
    res->stage_mask = stage_mask;
    res->own_tiles.resize(nlayers);
    uint64_t Ctot = 0;
    uint32_t max_cells = 0;
    auto bail = [&](int code) {
        rcb_result_free(res);
        return code;
    };
    for (int i = 0; i < nlayers; ++i) {
        const rcb_tc_layer& l = layers[i];
        const uint32_t cells = (uint32_t)l.width * l.height;
        if (l.regons_len != cells) return bail(fail(ctx, RCB_EINVALID_MESH, "Invalid region data size"));  // :138-140
        if (l.areas_len != cells) return bail(fail(ctx, RCB_EINVALID_MESH, "Invalid area data size"));     // :141-143
        if (cells && (!l.regons || !l.areas)) return bail(fail(ctx, RCB_EARG, "layer %d: null data", i));
        if ((uint64_t)l.first_obstacle + l.obstacle_count > (uint64_t)nobstacles) return bail(fail(ctx, RCB_EARG, "layer %d: obstacle range", i));
        if (l.hmin == 32767) {  // (32767 as i16, 32768 as i16 = -32768): Heightfield::add_span rejects min > max
            for (uint32_t c = 0; c < cells; ++c)
                if (l.regons[c] != 0) return bail(fail(ctx, RCB_ENAVMESH_GEN, "Invalid span height: min (32767) > max (-32768)"));
        }
        TileDesc& t = res->own_tiles[i];
        memset(&t, 0, sizeof t);
        t.width = l.width;
        t.height = l.height;
        for (int k = 0; k < 3; ++k) t.bmin[k] = l.bmin[k], t.bmax[k] = l.bmax[k];
        t.cs = cs;
        t.ch = ch;
        t.ics = 1.0f / cs;
        t.ich = 1.0f / ch;
        t.aux0 = (int32_t)l.hmin;
        t.aux1 = (int32_t)l.first_obstacle;
        t.aux2 = (int32_t)l.obstacle_count;
        t.cell_base = (uint32_t)Ctot;
        t.wmagic = l.width ? ((1ull << 40) / (unsigned long long)l.width) + 1 : 0;
        Ctot += cells;
        if (cells > max_cells) max_cells = cells;
    }
    res->Ctot = Ctot;
    TileMap tm{(uint32_t)nlayers, (max_cells + 255) / 256, 0};
    tm.nblocks = tm.blocks_per_tile * tm.ntiles;
#define CUB(expr)                                                                                            \
    do {                                                                                                     \
        cudaError_t _e = (expr);                                                                             \
        if (_e != cudaSuccess)                                                                               \
            return bail(fail(ctx, RCB_ECUDA, "CUDA error %s at %s:%d (%s)", cudaGetErrorString(_e), __FILE__, __LINE__, #expr)); \
    } while (0)
    // upload: tile descriptors, cell data (concatenated in layer order), obstacles
    const size_t need = align_up(sizeof(TileDesc) * (size_t)nlayers) + 2 * align_up(Ctot) + align_up(sizeof(rcb_tc_obstacle) * (size_t)nobstacles) + 1024;
    {
        int r = ensure_stage(ctx, need);
        if (r) return bail(r);
    }
    CUB(ctx->d_tiles.ensure(sizeof(TileDesc) * (size_t)nlayers + 16));
    CUB(ctx->d_tc_regons.ensure(Ctot + 16));
    CUB(ctx->d_tc_areas.ensure(Ctot + 16));
    CUB(ctx->d_tc_obs.ensure(sizeof(rcb_tc_obstacle) * (size_t)nobstacles + 16));
    {
        char* h = ctx->h_stage;
        size_t o = 0;
        memcpy(h + o, res->own_tiles.data(), sizeof(TileDesc) * (size_t)nlayers);
        if (nlayers) CUB(cudaMemcpyAsync(ctx->d_tiles.p, h + o, sizeof(TileDesc) * (size_t)nlayers, cudaMemcpyHostToDevice, st));
        o += align_up(sizeof(TileDesc) * (size_t)nlayers);
        char* hr = h + o;
        o += align_up(Ctot);
        char* ha = h + o;
        o += align_up(Ctot);
        for (int i = 0; i < nlayers; ++i) {
            const uint32_t cells = (uint32_t)layers[i].width * layers[i].height;
            if (!cells) continue;
            memcpy(hr + res->own_tiles[i].cell_base, layers[i].regons, cells);
            memcpy(ha + res->own_tiles[i].cell_base, layers[i].areas, cells);
        }
        if (Ctot) {
            CUB(cudaMemcpyAsync(ctx->d_tc_regons.p, hr, Ctot, cudaMemcpyHostToDevice, st));
            CUB(cudaMemcpyAsync(ctx->d_tc_areas.p, ha, Ctot, cudaMemcpyHostToDevice, st));
        }
        if (nobstacles) {
            memcpy(h + o, obstacles, sizeof(rcb_tc_obstacle) * (size_t)nobstacles);
            CUB(cudaMemcpyAsync(ctx->d_tc_obs.p, h + o, sizeof(rcb_tc_obstacle) * (size_t)nobstacles, cudaMemcpyHostToDevice, st));
        }
    }
    return tilecache_back_end(ctx, res, tm, nlayers, Ctot, nobstacles, stage_mask, erode_radius, out);
#undef CUB
}

int rcb_decompress_tiles(rcb_ctx* ctx, const uint8_t* blobs, const uint64_t* blob_offsets, int n, uint8_t* out, uint64_t out_cap,
                         uint64_t* out_offsets) {
    if (!ctx || n < 0 || !blob_offsets || !out_offsets || (n > 0 && blob_offsets[n] > 0 && !blobs)) return fail(ctx, RCB_EARG, "null argument");
    std::vector<unsigned long long> cap_off((size_t)n + 1, 0);
    for (int i = 0; i < n; ++i) {
        if (blob_offsets[i + 1] < blob_offsets[i]) return fail(ctx, RCB_EARG, "tile %d: blob_offsets not ascending", i);
        const uint64_t len = blob_offsets[i + 1] - blob_offsets[i];
        if (len >= 0xffff0000ull) return fail(ctx, RCB_EARG, "tile %d: blob of 4 GiB", i);
        const uint8_t* b = blobs + blob_offsets[i];
        const uint64_t size = len >= 4 ? ((uint64_t)b[0] | ((uint64_t)b[1] << 8) | ((uint64_t)b[2] << 16) | ((uint64_t)b[3] << 24)) : 0;
        // an LZ4 block cannot produce more than 255 bytes per input byte: a larger size prefix only means a short output
        // (lz4_flex allocates the prefix; here the allocation is device memory that stays with the context)
        cap_off[i + 1] = cap_off[i] + std::min<uint64_t>(size, 255ull * len + 1024);
    }
    const uint64_t total_cap = cap_off[n];
    if (total_cap > (8ull << 30)) return fail(ctx, RCB_EARG, "size prefixes add up to %llu bytes", (unsigned long long)total_cap);
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    ProfInstall prof_install(ctx);
    const uint64_t nbytes = n > 0 ? blob_offsets[n] : 0;
    CU(ctx->d_wire.ensure(nbytes + 64));
    CU(ctx->d_wire_off.ensure(16 * ((size_t)n + 1)));
    CU(ctx->d_tc_raw.ensure(total_cap + 64));
    CU(ctx->d_wire_err.ensure(8 * ((size_t)n + 1)));
    unsigned long long* d_boff = ctx->d_wire_off.as<unsigned long long>();
    unsigned long long* d_ooff = d_boff + n + 1;
    uint32_t* d_status = ctx->d_wire_err.as<uint32_t>();
    uint32_t* d_prod = d_status + n + 1;
    if (nbytes) CU(cudaMemcpyAsync(ctx->d_wire.p, blobs, nbytes, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d_boff, blob_offsets, 8 * ((size_t)n + 1), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d_ooff, cap_off.data(), 8 * ((size_t)n + 1), cudaMemcpyHostToDevice, st));
    launch_lz4_decompress(ctx->d_wire.as<uint8_t>(), d_boff, n, ctx->d_tc_raw.as<uint8_t>(), d_ooff, d_status, d_prod, st);
    std::vector<uint32_t> hs(2 * ((size_t)n + 1));
    CU(cudaMemcpyAsync(hs.data(), d_status, 8 * ((size_t)n + 1), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    CU(cudaGetLastError());
    out_offsets[0] = 0;
    for (int i = 0; i < n; ++i) {
        if (hs[i]) return fail(ctx, RCB_EDETOUR_FAILURE, "tile %d: LZ4 decompression failed", i);
        out_offsets[i + 1] = out_offsets[i] + hs[(size_t)n + 1 + i];
    }
    if (!out) return RCB_OK;
    if (out_cap < out_offsets[n]) return fail(ctx, RCB_EARG, "rcb_decompress_tiles: need %llu bytes", (unsigned long long)out_offsets[n]);
    for (int i = 0; i < n; ++i)  // produced <= size prefix: compact while copying
        if (out_offsets[i + 1] > out_offsets[i])
            CU(cudaMemcpyAsync(out + out_offsets[i], (const char*)ctx->d_tc_raw.p + cap_off[i], out_offsets[i + 1] - out_offsets[i], cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return RCB_OK;
}

int rcb_build_tilecache_compressed(rcb_ctx* ctx, const uint8_t* blobs, const uint64_t* blob_offsets, int nlayers,
                                   const uint32_t* first_obstacle, const uint32_t* obstacle_count, const rcb_tc_obstacle* obstacles,
                                   int nobstacles, float cs, float ch, uint32_t stage_mask, int erode_radius, rcb_result** out) {
    if (!ctx || !out || nlayers < 0 || !blob_offsets || (nlayers > 0 && blob_offsets[nlayers] > 0 && !blobs) || nobstacles < 0 ||
        (nobstacles > 0 && (!obstacles || !first_obstacle || !obstacle_count)))
        return fail(ctx, RCB_EARG, "null argument");
    *out = nullptr;
    if (!(cs > 0.0f) || !(ch > 0.0f)) return fail(ctx, RCB_EINVALID_MESH, "Invalid cell size or height");
    std::vector<unsigned long long> cap_off((size_t)nlayers + 1, 0);
    for (int i = 0; i < nlayers; ++i) {
        if (blob_offsets[i + 1] < blob_offsets[i]) return fail(ctx, RCB_EARG, "layer %d: blob_offsets not ascending", i);
        const uint64_t len = blob_offsets[i + 1] - blob_offsets[i];
        if (len >= 0xffff0000ull) return fail(ctx, RCB_EARG, "layer %d: blob of 4 GiB", i);
        const uint8_t* b = blobs + blob_offsets[i];
        const uint64_t size = len >= 4 ? ((uint64_t)b[0] | ((uint64_t)b[1] << 8) | ((uint64_t)b[2] << 16) | ((uint64_t)b[3] << 24)) : 0;
        cap_off[i + 1] = cap_off[i] + std::min<uint64_t>(size, 255ull * len + 1024);  // the decoder refuses to write past it (255 B of
                                                                                       // output per input byte is LZ4's ceiling)
        if (first_obstacle && obstacle_count && (uint64_t)first_obstacle[i] + obstacle_count[i] > (uint64_t)nobstacles)
            return fail(ctx, RCB_EARG, "layer %d: obstacle range", i);
    }
    const uint64_t total_cap = cap_off[nlayers];
    if (total_cap > (8ull << 30)) return fail(ctx, RCB_EARG, "size prefixes add up to %llu bytes", (unsigned long long)total_cap);
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    ProfInstall prof_install(ctx);
    stage_mask = (stage_mask | RCB_STAGE_COMPACT) & ~(uint32_t)(RCB_STAGE_RASTER | RCB_STAGE_FILTER | 0x700u);
    // ---- decode + parse on the device ---------------------------------------------------------------------
    const uint64_t nbytes = nlayers > 0 ? blob_offsets[nlayers] : 0;
    CU(ctx->d_wire.ensure(nbytes + 64));
    CU(ctx->d_wire_off.ensure(16 * ((size_t)nlayers + 1)));
    CU(ctx->d_tc_raw.ensure(total_cap + 64));
    CU(ctx->d_wire_err.ensure(8 * ((size_t)nlayers + 1)));
    CU(ctx->d_tc_parsed.ensure(sizeof(RcbTcParsed) * ((size_t)nlayers + 1)));
    unsigned long long* d_boff = ctx->d_wire_off.as<unsigned long long>();
    unsigned long long* d_ooff = d_boff + nlayers + 1;
    uint32_t* d_status = ctx->d_wire_err.as<uint32_t>();
    uint32_t* d_prod = d_status + nlayers + 1;
    if (nbytes) CU(cudaMemcpyAsync(ctx->d_wire.p, blobs, nbytes, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d_boff, blob_offsets, 8 * ((size_t)nlayers + 1), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d_ooff, cap_off.data(), 8 * ((size_t)nlayers + 1), cudaMemcpyHostToDevice, st));
    launch_lz4_decompress(ctx->d_wire.as<uint8_t>(), d_boff, nlayers, ctx->d_tc_raw.as<uint8_t>(), d_ooff, d_status, d_prod, st);
    launch_tc_parse(nlayers, ctx->d_tc_raw.as<uint8_t>(), d_ooff, d_status, d_prod, ctx->d_tc_parsed.as<RcbTcParsed>(), st);
    std::vector<RcbTcParsed> parsed((size_t)nlayers);
    if (nlayers) CU(cudaMemcpyAsync(parsed.data(), ctx->d_tc_parsed.p, sizeof(RcbTcParsed) * (size_t)nlayers, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    CU(cudaGetLastError());
    for (int i = 0; i < nlayers; ++i) {
        switch (parsed[i].status) {
            case 0: break;
            case RCB_EDETOUR_FAILURE: return fail(ctx, RCB_EDETOUR_FAILURE, "layer %d: LZ4 decompression failed", i);
            case RCB_EDETOUR_INVALID_PARAM: return fail(ctx, RCB_EDETOUR_INVALID_PARAM, "layer %d: invalid tile-cache layer data", i);
            case RCB_EINVALID_MESH: return fail(ctx, RCB_EINVALID_MESH, "layer %d: Invalid region / area data size", i);
            default: return fail(ctx, RCB_ENAVMESH_GEN, "layer %d: Invalid span height: min (32767) > max (-32768)", i);
        }
    }
    // ---- tile table ------------------------------------------------------------------------------------------
    auto* res = new rcb_result();
    res->ctx = ctx;
    res->ntiles = nlayers;
    res->stage_mask = stage_mask;
    res->own_tiles.resize(nlayers);
    uint64_t Ctot = 0;
    uint32_t max_cells = 0;
    for (int i = 0; i < nlayers; ++i) {
        const RcbTcParsed& l = parsed[i];
        TileDesc& t = res->own_tiles[i];
        memset(&t, 0, sizeof t);
        t.width = (int32_t)l.width;
        t.height = (int32_t)l.height;
        for (int k = 0; k < 3; ++k) t.bmin[k] = l.bmin[k], t.bmax[k] = l.bmax[k];
        t.cs = cs;
        t.ch = ch;
        t.ics = 1.0f / cs;
        t.ich = 1.0f / ch;
        t.aux0 = (int32_t)l.hmin;
        t.aux1 = first_obstacle ? (int32_t)first_obstacle[i] : 0;
        t.aux2 = obstacle_count ? (int32_t)obstacle_count[i] : 0;
        t.cell_base = (uint32_t)Ctot;
        t.wmagic = l.width ? ((1ull << 40) / (unsigned long long)l.width) + 1 : 0;
        const uint32_t cells = l.width * l.height;
        Ctot += cells;
        if (cells > max_cells) max_cells = cells;
    }
    res->Ctot = Ctot;
    TileMap tm{(uint32_t)nlayers, (max_cells + 255) / 256, 0};
    tm.nblocks = tm.blocks_per_tile * tm.ntiles;
    auto bail = [&](int code) {
        rcb_result_free(res);
        return code;
    };
    const size_t need = align_up(sizeof(TileDesc) * (size_t)nlayers) + align_up(sizeof(rcb_tc_obstacle) * (size_t)nobstacles) + 1024;
    {
        int r = ensure_stage(ctx, need);
        if (r) return bail(r);
    }
    cudaError_t e = cudaSuccess;
    auto ok = [&](cudaError_t x) { return (e = x) == cudaSuccess; };
    if (!ok(ctx->d_tiles.ensure(sizeof(TileDesc) * (size_t)nlayers + 16)) || !ok(ctx->d_tc_regons.ensure(Ctot + 16)) ||
        !ok(ctx->d_tc_areas.ensure(Ctot + 16)) || !ok(ctx->d_tc_obs.ensure(sizeof(rcb_tc_obstacle) * (size_t)nobstacles + 16)))
        return bail(fail(ctx, RCB_ECUDA, "CUDA error %s (tile-cache buffers)", cudaGetErrorString(e)));
    {
        char* h = ctx->h_stage;
        memcpy(h, res->own_tiles.data(), sizeof(TileDesc) * (size_t)nlayers);
        if (nlayers && !ok(cudaMemcpyAsync(ctx->d_tiles.p, h, sizeof(TileDesc) * (size_t)nlayers, cudaMemcpyHostToDevice, st)))
            return bail(fail(ctx, RCB_ECUDA, "CUDA error %s (tile table)", cudaGetErrorString(e)));
        if (nobstacles) {
            char* ho = h + align_up(sizeof(TileDesc) * (size_t)nlayers);
            memcpy(ho, obstacles, sizeof(rcb_tc_obstacle) * (size_t)nobstacles);
            if (!ok(cudaMemcpyAsync(ctx->d_tc_obs.p, ho, sizeof(rcb_tc_obstacle) * (size_t)nobstacles, cudaMemcpyHostToDevice, st)))
                return bail(fail(ctx, RCB_ECUDA, "CUDA error %s (obstacles)", cudaGetErrorString(e)));
        }
    }
    launch_tc_gather(ctx->d_tiles.as<TileDesc>(), tm, ctx->d_tc_raw.as<uint8_t>(), d_ooff, ctx->d_tc_parsed.as<RcbTcParsed>(),
                     ctx->d_tc_regons.as<uint8_t>(), ctx->d_tc_areas.as<uint8_t>(), st);
    return tilecache_back_end(ctx, res, tm, nlayers, Ctot, nobstacles, stage_mask, erode_radius, out);
}

int rcb_apply_area_ops(rcb_ctx* ctx, const rcb_config* cfgs, int ntiles, const uint32_t* cell_offset, const int16_t* span_min,
                       const int16_t* span_max, const uint8_t* span_area, const uint8_t* areas_override, const rcb_area_op* ops, int nops,
                       const float* poly_verts, size_t npoly_floats, uint32_t stage_mask, int erode_radius, rcb_result** out) {
    if (stage_mask & RCB_STAGE_RASTER) return fail(ctx, RCB_EARG, "rcb_apply_area_ops: RCB_STAGE_RASTER must not be set");
    if (ntiles > 0 && !cell_offset) return fail(ctx, RCB_EARG, "cell_offset is null");
    if (nops < 0 || (nops > 0 && !ops)) return fail(ctx, RCB_EARG, "ops is null");
    for (int i = 0; i < nops; ++i) {
        if (ops[i].type < RCB_AREA_OP_MEDIAN || ops[i].type > RCB_AREA_OP_CONVEX_POLY) return fail(ctx, RCB_EARG, "op %d: unknown type", i);
        if (ops[i].type == RCB_AREA_OP_CONVEX_POLY &&
            (ops[i].num_verts == 0 || !poly_verts || 3ull * ((uint64_t)ops[i].first_vert + ops[i].num_verts) > npoly_floats))
            return fail(ctx, RCB_EARG, "op %d: polygon vertices out of range", i);
    }
    Source s;
    s.cell_offset = cell_offset;
    s.smin = span_min;
    s.smax = span_max;
    s.sarea = span_area;
    s.areas_override = areas_override;
    s.ops = ops;
    s.nops = nops;
    s.poly_verts = poly_verts;
    s.npoly_floats = npoly_floats;
    return run_pipeline(ctx, cfgs, ntiles, stage_mask | RCB_STAGE_COMPACT, erode_radius, s, out);
}

int rcb_result_totals(const rcb_result* cres, rcb_totals* out) {
    auto* res = const_cast<rcb_result*>(cres);
    if (!res || !out) return RCB_EARG;
    if (!res->ctx) return RCB_EARG;  // the context is gone (rcb_destroy): only rcb_result_free is left
    if (!res->children.empty()) {
        memset(out, 0, sizeof *out);
        for (auto* c : res->children) {
            rcb_totals t;
            int r = rcb_result_totals(c, &t);
            if (r) return r;
            out->tiles += t.tiles, out->cells += t.cells, out->fragments += t.fragments, out->spans += t.spans;
            out->connections += t.connections, out->walkable_spans += t.walkable_spans, out->result_bytes += t.result_bytes;
            out->triangles = t.triangles, out->vertices = t.vertices;
        }
        return RCB_OK;
    }
    int r = fetch_tinfo(res);
    if (r) return r;
    memset(out, 0, sizeof *out);
    out->tiles = res->ntiles;
    out->triangles = res->ntris;
    out->vertices = res->nverts;
    out->cells = res->Ctot;
    out->fragments = res->F;
    out->spans = res->S;
    out->connections = res->K;
    for (auto& t : res->tinfo) out->walkable_spans += t.walkable_span_count;
    out->result_bytes = res->copied_bytes ? res->copied_bytes : download_bytes(res);
    return RCB_OK;
}

// Bytes a download of this result moves (what rcb_totals.result_bytes reports before the download has happened).
static uint64_t download_bytes(const rcb_result* res) {
    const ArenaLayout& l = res->lay;
    if (res->slim) return l.slimA + l.slimB;
    if (res->tile_mode && res->have_K && res->lay.bytesB) return l.bytesA + 9 * res->K;
    return l.bytesA + l.bytesB;
}

// Slim layout: conn_mask / conn_ref16 are derived on the device from the full arrays right before the first download.
static int slim_pack(rcb_result* res, cudaStream_t st) {
    rcb_ctx* ctx = res->ctx;
    if (!res->slim || res->slim_packed) return RCB_OK;
    res->slim_packed = true;
    if (res->ntiles == 0 || !res->entry) return RCB_OK;
    char* A = res->devA.p;
    char* B = res->devB.p;
    CU(ctx->d_wire_err.ensure(64));
    if (!res->slim_flag) CU(cudaMallocHost((void**)&res->slim_flag, sizeof(unsigned long long)));
    *res->slim_flag = 0;
    CU(cudaMemsetAsync(ctx->d_wire_err.p, 0, 8, st));
    launch_slim_pack(res->entry->d_tiles.as<TileDesc>(), (const TileInfo*)(A + res->lay.tinfo), res->ntiles, res->entry->max_cells,
                     (const uint32_t*)(A + res->lay.first_conn), (const uint32_t*)(B + res->lay.conn_ref),
                     (const uint8_t*)(B + res->lay.conn_dir), (const uint32_t*)(B + res->lay.conn_next), (const uint32_t*)(A + res->lay.cell_index),
                     (const uint32_t*)(A + res->lay.cell_count), (const uint16_t*)(A + res->lay.con), (uint8_t*)(A + res->lay.conn_mask),
                     (uint16_t*)(B + res->lay.conn_ref16), (uint8_t*)(B + res->lay.conn_off8), (uint16_t*)(A + res->lay.cell_count16),
                     (uint8_t*)(A + res->lay.con8), ctx->d_wire_err.as<unsigned long long>(), st);
    CU(cudaMemcpyAsync(res->slim_flag, ctx->d_wire_err.p, 8, cudaMemcpyDeviceToHost, st));
    return RCB_OK;
}

// D2H of one result's arenas on `st`.  A result that is still pending is copied by its CAPACITIES (the layout was made for
// Scap spans / Kcap connections, tight bounds of a speculative batch), so nothing waits for the device here; with
// wait == false the copies are left in flight and finish_download completes the bookkeeping once `st` has been synchronized.
static int download_on(rcb_result* res, cudaStream_t st, bool wait) {
    rcb_ctx* ctx = res->ctx;
    if (res->downloaded) return RCB_OK;
    ProfInstall prof_install(ctx);
    const bool on_own = st == ctx->stream;
    if (!(res->pending && res->spec)) {  // sizes are known (or cheap to know): copy exactly what exists
        int r = on_own ? fetch_tinfo(res) : fetch_tinfo(res, st, true);
        if (r) return r;
    }
    if (res->slim) {
        int r = slim_pack(res, st);
        if (r) return r;
    }
    const ArenaLayout& l = res->lay;
    CU(pool_get(ctx->free_host, l.bytesA, true, res->hostA));  // slim: only the prefix is copied; the rest is room for a wide fallback
    uint64_t copied = 0;
    if (l.bytesB) {
        if (res->slim) {
            if (l.slimB) {
                CU(pool_get(ctx->free_host, l.bytesB, true, res->hostB));  // room for conn_ref should 16 bits not do
                CU(cudaMemcpyAsync(res->hostB.p, res->devB.p, l.slimB, cudaMemcpyDeviceToHost, st));
                copied += l.slimB;
            }
        } else {
            CU(pool_get(ctx->free_host, l.bytesB, true, res->hostB));
            if (res->tile_mode && !res->pending) {  // the arena is sized by a bound on K: copy the used prefix of each array
                const size_t K = res->K;
                if (K) {
                    CU(cudaMemcpyAsync(res->hostB.p + l.conn_ref, res->devB.p + l.conn_ref, 4 * K, cudaMemcpyDeviceToHost, st));
                    CU(cudaMemcpyAsync(res->hostB.p + l.conn_next, res->devB.p + l.conn_next, 4 * K, cudaMemcpyDeviceToHost, st));
                    CU(cudaMemcpyAsync(res->hostB.p + l.conn_dir, res->devB.p + l.conn_dir, K, cudaMemcpyDeviceToHost, st));
                    copied += 9 * K;
                }
            } else {
                CU(cudaMemcpyAsync(res->hostB.p, res->devB.p, l.bytesB, cudaMemcpyDeviceToHost, st));
                copied += l.bytesB;
            }
        }
    }
    const size_t nA = res->slim ? l.slimA : l.bytesA;
    CU(cudaMemcpyAsync(res->hostA.p, res->devA.p, nA, cudaMemcpyDeviceToHost, st));
    copied += nA;
    res->copied_bytes = copied;
    res->copy_in_flight = true;
    if (wait) {
        CU(cudaStreamSynchronize(st));
        return finish_download(res, st);
    }
    return RCB_OK;
}

// After the copies of download_on have completed: host-side bookkeeping; a speculative batch that failed on the device
// is rebuilt (resolve_result) and downloaded again, synchronously.
static int finish_download(rcb_result* res, cudaStream_t st) {
    rcb_ctx* ctx = res->ctx;
    if (res->downloaded || !res->copy_in_flight) return RCB_OK;
    res->copy_in_flight = false;
    const bool was_spec = res->pending && res->spec;
    int r = resolve_result(res);
    if (r) return r;
    if (was_spec && !res->hostA.p) {  // rebuilt: the arenas copied above are gone
        res->copied_bytes = 0;
        return download_on(res, ctx->stream, true);
    }
    if (res->slim) {
        const unsigned long long f = res->slim_flag ? *res->slim_flag : (res->ntiles == 0 ? 0ull : 15ull);
        res->slim_ref16 = !(f & 1ull), res->slim_con8 = !(f & 2ull), res->slim_cnt16 = !(f & 4ull), res->slim_off8 = !(f & 8ull);
        bool more = false;
        if (!res->slim_off8) {
            // a connection whose offset inside the neighbour cell passes 255 (a cell with 257+ spans): the references and
            // con travel after all, each in the narrowest form that holds it
            if (res->slim_ref16 && res->K) {
                CU(cudaMemcpyAsync(res->hostB.p + res->lay.conn_ref16, res->devB.p + res->lay.conn_ref16, 2 * res->K, cudaMemcpyDeviceToHost, st));
                res->copied_bytes += 2 * res->K, more = true;
            } else if (res->K) {
                CU(cudaMemcpyAsync(res->hostB.p + res->lay.conn_ref, res->devB.p + res->lay.conn_ref, 4 * res->K, cudaMemcpyDeviceToHost, st));
                res->copied_bytes += 4 * res->K, more = true;
            }
            if (res->S) {
                const size_t off = res->slim_con8 ? res->lay.con8 : res->lay.con, n = (res->slim_con8 ? 4 : 8) * res->S;
                CU(cudaMemcpyAsync(res->hostA.p + off, res->devA.p + off, n, cudaMemcpyDeviceToHost, st));
                res->copied_bytes += n, more = true;
            }
        }
        if (!res->slim_cnt16 && res->Ctot) {
            CU(cudaMemcpyAsync(res->hostA.p + res->lay.cell_count, res->devA.p + res->lay.cell_count, 4 * res->Ctot, cudaMemcpyDeviceToHost, st));
            res->copied_bytes += 4 * res->Ctot, more = true;
        }
        if (more) CU(cudaStreamSynchronize(st));
    }
    res->slim_host = res->slim;
    res->tinfo.assign((TileInfo*)(res->hostA.p + res->lay.tinfo), (TileInfo*)(res->hostA.p + res->lay.tinfo) + res->ntiles);
    res->have_tinfo = true;
    if (res->tile_mode && !res->have_K) {
        uint64_t k = 0;
        for (auto& t : res->tinfo) k += t.conn_count;
        res->K = k, res->have_K = true;
    }
    res->downloaded = true;
    return RCB_OK;
}

int rcb_result_download(rcb_result* res) {
    if (!res) return RCB_EARG;
    rcb_ctx* ctx = res->ctx;
    if (!ctx) return RCB_EARG;
    if (res->downloaded) return RCB_OK;
    CU(cudaSetDevice(ctx->device));
    if (!res->children.empty()) {
        for (auto* c : res->children) {
            int r = rcb_result_download(c);
            if (r) return r;
        }
        res->downloaded = true;
        return RCB_OK;
    }
    return download_on(res, ctx->stream, true);
}

static int fill_view(const rcb_result* res, int tile, rcb_tile_view* v, const char* A, const char* B, bool host) {
    const TileDesc& td = res->tiles()[tile];
    const TileInfo& ti = res->tinfo[tile];
    const ArenaLayout& l = res->lay;
    const bool slim = host && res->slim_host;  // the host copy of a slim result lacks what the host can rebuild
    memset(v, 0, sizeof *v);
    v->width = td.width;
    v->height = td.height;
    v->cell_count = (uint32_t)(td.width * td.height);
    v->span_count = ti.span_count;
    v->conn_count = ti.conn_count;
    v->walkable_span_count = ti.walkable_span_count;
    v->max_height = (uint16_t)ti.max_height;
    v->max_distance = (uint16_t)ti.max_distance;
    v->status = ti.status;
    if (!slim) v->hf_cell_offset = (const uint32_t*)(A + l.hf_cell_offset) + td.cell_base + tile;
    v->span_min = (const int16_t*)(A + l.span_min) + ti.span_base;
    v->span_max = (const int16_t*)(A + l.span_max) + ti.span_base;
    v->span_area = (const uint8_t*)(A + l.span_area) + ti.span_base;
    if (res->stage_mask & RCB_STAGE_COMPACT) {
        v->dist = (const uint16_t*)(A + l.dist) + ti.span_base;
        if (!slim) {
            v->cell_index = (const uint32_t*)(A + l.cell_index) + td.cell_base;
            v->cell_count_arr = (const uint32_t*)(A + l.cell_count) + td.cell_base;
            v->areas = (const uint8_t*)(A + l.areas) + ti.span_base;
            v->con = (const uint16_t*)(A + l.con) + 4 * (size_t)ti.span_base;
            v->first_conn = (const uint32_t*)(A + l.first_conn) + ti.span_base;
            v->conn_ref = (const uint32_t*)(B + l.conn_ref) + ti.conn_base;
            v->conn_dir = (const uint8_t*)(B + l.conn_dir) + ti.conn_base;
            v->conn_next = (const uint32_t*)(B + l.conn_next) + ti.conn_base;
        } else {  // the host copy of a slim result: narrow forms, or the wide array where a narrow one did not do
            if (l.areas_in_slim) v->areas = (const uint8_t*)(A + l.areas) + ti.span_base;  // else NULL: areas == span_area
            if (res->slim_cnt16)
                v->cell_count16 = (const uint16_t*)(A + l.cell_count16) + td.cell_base;
            else
                v->cell_count_arr = (const uint32_t*)(A + l.cell_count) + td.cell_base;
            if (!res->slim_off8) {
                if (res->slim_con8)
                    v->con8 = (const uint8_t*)(A + l.con8) + 4 * (size_t)ti.span_base;
                else
                    v->con = (const uint16_t*)(A + l.con) + 4 * (size_t)ti.span_base;
            }
        }
        if (res->slim && res->slim_packed) {
            v->conn_mask = (const uint8_t*)(A + l.conn_mask) + ti.span_base;
            if (!host) {
                v->conn_off8 = (const uint8_t*)(B + l.conn_off8) + ti.conn_base;
                v->conn_ref16 = (const uint16_t*)(B + l.conn_ref16) + ti.conn_base;
                v->cell_count16 = (const uint16_t*)(A + l.cell_count16) + td.cell_base;
                v->con8 = (const uint8_t*)(A + l.con8) + 4 * (size_t)ti.span_base;
            } else if (res->slim_off8) {
                v->conn_off8 = (const uint8_t*)(B + l.conn_off8) + ti.conn_base;
            } else if (res->slim_ref16) {
                v->conn_ref16 = (const uint16_t*)(B + l.conn_ref16) + ti.conn_base;
            } else {
                v->conn_ref = (const uint32_t*)(B + l.conn_ref) + ti.conn_base;
            }
        }
        if (res->stage_mask & RCB_STAGE_REGIONS) {
            v->reg = (const uint16_t*)(A + l.reg) + ti.span_base;
            v->max_regions = ti.max_regions;
        }
    }
    return RCB_OK;
}

static const rcb_result* child_of(const rcb_result* res, int& tile) {
    const size_t g = (size_t)(std::upper_bound(res->child_first.begin(), res->child_first.end(), tile) - res->child_first.begin()) - 1;
    tile -= res->child_first[g];
    return res->children[g];
}

int rcb_result_tile(const rcb_result* res, int tile, rcb_tile_view* out) {
    if (!res || !out || tile < 0 || tile >= res->ntiles) return RCB_EARG;
    if (!res->children.empty()) {
        const rcb_result* c = child_of(res, tile);
        return rcb_result_tile(c, tile, out);
    }
    if (!res->ctx) return RCB_EARG;
    if (!res->downloaded) return fail(res->ctx, RCB_EARG, "rcb_result_tile: call rcb_result_download first");
    return fill_view(res, tile, out, res->hostA.p, res->hostB.p, true);
}

int rcb_result_tile_device(const rcb_result* cres, int tile, rcb_tile_view* out) {
    if (cres && !cres->children.empty()) {
        if (!out || tile < 0 || tile >= cres->ntiles) return RCB_EARG;
        const rcb_result* c = child_of(cres, tile);
        return rcb_result_tile_device(c, tile, out);
    }
    auto* res = const_cast<rcb_result*>(cres);
    if (!res || !out || tile < 0 || tile >= res->ntiles || !res->ctx) return RCB_EARG;
    int r = fetch_tinfo(res);
    if (r) return r;
    return fill_view(res, tile, out, res->devA.p, res->devB.p, false);
}

void rcb_result_free(rcb_result* res) {
    if (!res) return;
    for (auto* c : res->children) rcb_result_free(c);
    rcb_ctx* ctx = res->ctx;
    if (ctx) {
        cudaSetDevice(ctx->device);
        for (size_t i = 0; i < ctx->live.size(); ++i)
            if (ctx->live[i] == res) {
                ctx->live[i] = ctx->live.back();
                ctx->live.pop_back();
                break;
            }
        if (res->copy_in_flight) cudaDeviceSynchronize();  // a host arena must not return to the pool under a running copy
        release_slot(res);
        if (res->done) cudaEventDestroy(res->done);
        // arenas go back to the pool; work that still reads them is ordered on the same stream
        pool_put(ctx->free_dev, res->devA, false);
        pool_put(ctx->free_dev, res->devB, false);
        pool_put(ctx->free_host, res->hostA, true);
        pool_put(ctx->free_host, res->hostB, true);
    }
    if (res->slim_flag) cudaFreeHost(res->slim_flag);
    delete res;
}

}  // extern "C"
