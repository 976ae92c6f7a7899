// rcb_multi.cu — one process, several devices: the multi-tile build loop of detour-dynamic
// (DynamicNavMesh::build / build_async, crates/detour-dynamic/src/dynamic_navmesh.rs:558-684: `for key in tile_keys`, serial,
// batches of 10 with yield_now) spread over the GPUs of one box.
//
// Tiles are independent units (no halo, border_size 0), so nothing but the mesh and the results ever moves:
//   * the mesh is uploaded ONCE (host -> first device) and handed to the other devices over NVLink (cudaMemcpyPeer);
//   * the tiles are cut into groups of consecutive tiles; two workers per device (a context and a host thread each) work
//     through their own shares of the groups and then steal from the worker with the most left (the static slabs of round 1
//     lost 24 % on BASELINE config 3, whose buildings are unevenly spread), running the ordinary single-device pipeline
//     (rcb_build_tiles) on each group, two groups in flight per worker;
//   * a group meets mostly triangles that lie elsewhere; k_bin_count skips them 256 at a time (MeshView::blocks).
// Built on the public single-device entry points of rcb_api.cu; no collective is involved.
#include <atomic>
#include <mutex>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/recast_b200.h"

// Two contexts (streams, scratch buffers, host threads) per device: the groups a device builds are small enough that the
// latency-bound stretches of one (the row-sequential sweeps, the read-backs of the global back end) leave the GPU idle;
// a second group running beside it fills them.  Context w belongs to device w / kLanes.
constexpr int kLanes = 2;

struct rcb_multi {
    std::vector<int> devices;
    std::vector<rcb_ctx*> ctx;  // [devices * kLanes]
    std::vector<void*> d_verts, d_tris;  // per device copies of the mesh
    std::vector<size_t> cap_verts, cap_tris;
    size_t nfloats = 0, nidx = 0;
    std::string err;
};

struct rcb_multi_result {
    rcb_multi* m = nullptr;
    int ntiles = 0;
    std::vector<int> group_first;       // [ngroups + 1]
    std::vector<rcb_result*> group_res;  // [ngroups]
    std::vector<int> group_dev;          // the context (worker) that built the group: device = m->devices[w / kLanes]
    std::vector<double> dev_ms;          // per device: host time spent building its groups
};

namespace {

int mfail(rcb_multi* m, int code, const std::string& msg) {
    if (m) m->err = msg;
    return code;
}

}  // namespace

extern "C" {

int rcb_multi_create(const int* devices, int ndevices, rcb_multi** out) {
    if (!out || ndevices <= 0 || !devices) return RCB_EARG;
    *out = nullptr;
    auto* m = new rcb_multi();
    for (int i = 0; i < ndevices; ++i) {
        for (int l = 0; l < kLanes; ++l) {
            rcb_ctx* c = nullptr;
            const int r = rcb_create(devices[i], &c);
            if (r) {
                for (auto* x : m->ctx) rcb_destroy(x);
                delete m;
                return r;
            }
            m->ctx.push_back(c);
        }
        m->devices.push_back(devices[i]);
    }
    m->d_verts.assign(ndevices, nullptr);
    m->d_tris.assign(ndevices, nullptr);
    m->cap_verts.assign(ndevices, 0);
    m->cap_tris.assign(ndevices, 0);
    // peer access where the topology has it (NVLink / NVSwitch): the mesh hand-over then goes device to device
    for (int i = 1; i < ndevices; ++i) {
        int can = 0;
        if (cudaDeviceCanAccessPeer(&can, devices[i], devices[0]) == cudaSuccess && can) {
            cudaSetDevice(devices[i]);
            const cudaError_t e = cudaDeviceEnablePeerAccess(devices[0], 0);
            if (e != cudaSuccess) cudaGetLastError();  // already enabled, or not supported: cudaMemcpyPeer stages through the host
        }
    }
    *out = m;
    return RCB_OK;
}

void rcb_multi_destroy(rcb_multi* m) {
    if (!m) return;
    for (auto* c : m->ctx) rcb_destroy(c);
    for (size_t i = 0; i < m->devices.size(); ++i) {
        cudaSetDevice(m->devices[i]);
        if (m->d_verts[i]) cudaFree(m->d_verts[i]);
        if (m->d_tris[i]) cudaFree(m->d_tris[i]);
    }
    delete m;
}

const char* rcb_multi_last_error(const rcb_multi* m) { return m ? m->err.c_str() : "null context"; }

int rcb_multi_device_count(const rcb_multi* m) { return m ? (int)m->devices.size() : 0; }

int rcb_multi_upload_mesh(rcb_multi* m, const float* verts, size_t nfloats, const int32_t* tris, size_t nidx) {
    if (!m || (nfloats && !verts) || (nidx && !tris)) return mfail(m, RCB_EARG, "null mesh pointer");
    const int n = (int)m->devices.size();
    for (int i = 0; i < n; ++i) {
        if (cudaSetDevice(m->devices[i]) != cudaSuccess) return mfail(m, RCB_ECUDA, "cudaSetDevice failed");
        if (4 * nfloats + 16 > m->cap_verts[i]) {
            if (m->d_verts[i]) cudaFree(m->d_verts[i]);
            m->d_verts[i] = nullptr;
            if (cudaMalloc(&m->d_verts[i], 4 * nfloats + 16) != cudaSuccess) return mfail(m, RCB_ECUDA, "out of device memory (vertices)");
            m->cap_verts[i] = 4 * nfloats + 16;
        }
        if (4 * nidx + 16 > m->cap_tris[i]) {
            if (m->d_tris[i]) cudaFree(m->d_tris[i]);
            m->d_tris[i] = nullptr;
            if (cudaMalloc(&m->d_tris[i], 4 * nidx + 16) != cudaSuccess) return mfail(m, RCB_ECUDA, "out of device memory (indices)");
            m->cap_tris[i] = 4 * nidx + 16;
        }
    }
    // host -> first device, once; first device -> every other device
    cudaSetDevice(m->devices[0]);
    if (nfloats && cudaMemcpy(m->d_verts[0], verts, 4 * nfloats, cudaMemcpyHostToDevice) != cudaSuccess) return mfail(m, RCB_ECUDA, "H2D failed");
    if (nidx && cudaMemcpy(m->d_tris[0], tris, 4 * nidx, cudaMemcpyHostToDevice) != cudaSuccess) return mfail(m, RCB_ECUDA, "H2D failed");
    for (int i = 1; i < n; ++i) {
        if (nfloats && cudaMemcpyPeer(m->d_verts[i], m->devices[i], m->d_verts[0], m->devices[0], 4 * nfloats) != cudaSuccess)
            return mfail(m, RCB_ECUDA, "peer copy failed (vertices)");
        if (nidx && cudaMemcpyPeer(m->d_tris[i], m->devices[i], m->d_tris[0], m->devices[0], 4 * nidx) != cudaSuccess)
            return mfail(m, RCB_ECUDA, "peer copy failed (indices)");
    }
    for (int i = 0; i < n; ++i) {
        cudaSetDevice(m->devices[i]);
        cudaDeviceSynchronize();
        for (int l = 0; l < kLanes; ++l) {  // both contexts of the device borrow the same copy (and check its indices)
            const int r = rcb_set_mesh_device(m->ctx[i * kLanes + l], m->d_verts[i], nfloats, m->d_tris[i], nidx);
            if (r) return mfail(m, r, rcb_last_error(m->ctx[i * kLanes + l]));
        }
    }
    m->nfloats = nfloats;
    m->nidx = nidx;
    return RCB_OK;
}

int rcb_multi_build_tiles(rcb_multi* m, const rcb_config* cfgs, int ntiles, uint32_t stage_mask, int erode_radius, int tiles_per_group,
                          int download, rcb_multi_result** out) {
    if (!m || !out || ntiles < 0 || (ntiles > 0 && !cfgs)) return mfail(m, RCB_EARG, "null argument");
    *out = nullptr;
    const int ndev = (int)m->devices.size();
    if (tiles_per_group <= 0) {
        // about four groups per device: a straggler costs at most a quarter of a device's share before the others step in,
        // and the fixed cost of a build (a dozen small launches) stays a few per cent of a group
        tiles_per_group = (ntiles + 4 * ndev - 1) / (4 * ndev);
        if (tiles_per_group < 1) tiles_per_group = 1;
    }
    auto* res = new rcb_multi_result();
    res->m = m;
    res->ntiles = ntiles;
    for (int f = 0; f < ntiles; f += tiles_per_group) res->group_first.push_back(f);
    const int ngroups = (int)res->group_first.size();
    res->group_first.push_back(ntiles);
    res->group_res.assign(ngroups, nullptr);
    res->group_dev.assign(ngroups, -1);
    res->dev_ms.assign(ndev, 0.0);
    // Every worker (a context and its host thread; kLanes per device) starts on its own contiguous share of the groups - the
    // same (group, context) pairs build after build, so each context already knows the group's sizes and issues it without a
    // host round trip - and, when that is used up, steals from the back of the share with the most groups left.
    const int nw = ndev * kLanes;
    std::mutex qm;
    std::vector<int> lo(nw), hi(nw);
    for (int w = 0; w < nw; ++w) lo[w] = (int)((long long)ngroups * w / nw), hi[w] = (int)((long long)ngroups * (w + 1) / nw);
    auto take = [&](int w) -> int {
        std::lock_guard<std::mutex> lk(qm);
        if (lo[w] < hi[w]) return lo[w]++;
        int victim = -1;
        for (int k = 0; k < nw; ++k)
            if (hi[k] - lo[k] > (victim < 0 ? 0 : hi[victim] - lo[victim])) victim = k;
        return victim < 0 ? -1 : --hi[victim];
    };
    std::vector<int> status(nw, RCB_OK);
    std::vector<std::string> msg(nw);
    auto worker = [&](int d) {  // d: worker index
        rcb_ctx* ctx = m->ctx[d];
        cudaSetDevice(m->devices[d / kLanes]);
        int prev = -1;
        auto settle = [&](int g) {  // the first host use waits for the group (and rebuilds it, should a size have been short)
            rcb_totals t;
            int r = rcb_result_totals(res->group_res[g], &t);
            if (r == RCB_OK && download) r = rcb_result_download(res->group_res[g]);
            if (r != RCB_OK && status[d] == RCB_OK) status[d] = r, msg[d] = rcb_last_error(ctx);
        };
        for (;;) {
            const int g = status[d] == RCB_OK ? take(d) : -1;
            if (g < 0) break;
            const int first = res->group_first[g], n = res->group_first[g + 1] - first;
            rcb_result* r = nullptr;
            const int rc = rcb_build_tiles(ctx, cfgs + first, n, stage_mask, erode_radius, &r);
            if (rc != RCB_OK) {
                status[d] = rc, msg[d] = rcb_last_error(ctx);
                break;
            }
            res->group_res[g] = r;
            res->group_dev[g] = d;
            // two groups in flight per device: this one is queued behind the previous one, which is waited for now - so a
            // device takes its next group only when it is about to run dry, and faster devices take more of them
            if (prev >= 0) settle(prev);
            prev = g;
        }
        if (prev >= 0) settle(prev);
    };
    std::vector<std::thread> th;
    for (int w = 0; w < nw; ++w) th.emplace_back(worker, w);
    for (auto& t : th) t.join();
    for (int d = 0; d < nw; ++d)
        if (status[d] != RCB_OK) {
            const int code = status[d];
            m->err = msg[d];
            rcb_multi_result_free(res);
            return code;
        }
    *out = res;
    return RCB_OK;
}

static int group_of(const rcb_multi_result* res, int tile) {
    int lo = 0, hi = (int)res->group_res.size() - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) / 2;
        if (res->group_first[mid] <= tile)
            lo = mid;
        else
            hi = mid - 1;
    }
    return lo;
}

int rcb_multi_result_tile(const rcb_multi_result* res, int tile, rcb_tile_view* out) {
    if (!res || !out || tile < 0 || tile >= res->ntiles) return RCB_EARG;
    const int g = group_of(res, tile);
    return rcb_result_tile(res->group_res[g], tile - res->group_first[g], out);
}

int rcb_multi_result_tile_device(const rcb_multi_result* res, int tile, rcb_tile_view* out, int* device) {
    if (!res || !out || tile < 0 || tile >= res->ntiles) return RCB_EARG;
    const int g = group_of(res, tile);
    if (device) *device = res->m->devices[res->group_dev[g] / kLanes];
    return rcb_result_tile_device(res->group_res[g], tile - res->group_first[g], out);
}

int rcb_multi_result_download(rcb_multi_result* res) {
    if (!res) return RCB_EARG;
    const int ndev = (int)res->m->ctx.size();  // one thread per context: a context is driven by one thread at a time
    std::vector<int> status(ndev, RCB_OK);
    std::vector<std::thread> th;
    for (int d = 0; d < ndev; ++d)
        th.emplace_back([&, d]() {
            cudaSetDevice(res->m->devices[d / kLanes]);
            for (size_t g = 0; g < res->group_res.size(); ++g)
                if (res->group_dev[g] == d) {
                    const int r = rcb_result_download(res->group_res[g]);
                    if (r && status[d] == RCB_OK) status[d] = r;
                }
        });
    for (auto& t : th) t.join();
    for (int d = 0; d < ndev; ++d)
        if (status[d]) return mfail(res->m, status[d], rcb_last_error(res->m->ctx[d]));
    return RCB_OK;
}

int rcb_multi_result_totals(const rcb_multi_result* res, rcb_totals* out, int* tiles_per_device, int ndevices) {
    if (!res || !out) return RCB_EARG;
    memset(out, 0, sizeof *out);
    for (int d = 0; tiles_per_device && d < ndevices; ++d) tiles_per_device[d] = 0;
    for (size_t g = 0; g < res->group_res.size(); ++g) {
        rcb_totals t;
        const int r = rcb_result_totals(res->group_res[g], &t);
        if (r) return r;
        out->tiles += t.tiles, out->cells += t.cells, out->fragments += t.fragments, out->spans += t.spans;
        out->connections += t.connections, out->walkable_spans += t.walkable_spans, out->result_bytes += t.result_bytes;
        out->triangles = t.triangles, out->vertices = t.vertices;
        if (tiles_per_device && res->group_dev[g] / kLanes < ndevices) tiles_per_device[res->group_dev[g] / kLanes] += (int)t.tiles;
    }
    return RCB_OK;
}

void rcb_multi_result_free(rcb_multi_result* res) {
    if (!res) return;
    for (size_t g = 0; g < res->group_res.size(); ++g)
        if (res->group_res[g]) {
            if (res->group_dev[g] >= 0) cudaSetDevice(res->m->devices[res->group_dev[g] / kLanes]);
            rcb_result_free(res->group_res[g]);
        }
    delete res;
}

}  // extern "C"
