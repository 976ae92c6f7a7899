// k_lz4.cu — tile-cache front half, batched over layers:
//   TileCache::decompress_tile       detour-tilecache/src/tile_cache.rs:854-872  (lz4_flex::decompress_size_prepended)
//   TileCacheLayer::from_bytes       detour-tilecache/src/tile_cache_data.rs:128-232, 281-320
//
// LZ4 block format (lz4_flex 0.11, Cargo.toml:64): after a 4-byte little-endian size, sequences of
// `token | literal-length bytes | literals | u16 offset | match-length bytes`; a match copies bytes the block
// already produced, byte by byte, so offset < length repeats a period.  One warp decodes one layer: the
// sequence headers are read by every lane from the same addresses (one broadcast transaction), literal and
// match bytes are copied 32 at a time; a match with overlap is written as out[o+i] = out[o-offset + i % offset],
// which only reads bytes produced before the match.  Decoding is serial in the sequences of one block and
// parallel over the layers of the batch.
#include "rcb_common.cuh"

namespace {

// status per layer: 0 ok, 1 = lz4_flex's Err (Error::Detour(Failure))
//
// Every step of a sequence depends on the one before (token -> lengths -> literals -> offset -> match), so the
// decode time of a layer is (sequences) x (latency of that chain).  With input and output in global memory
// the chain cost ~1000 cycles per sequence; blocks that fit are therefore staged through shared memory
// (input <= LZ_IN_CAP, size prefix <= LZ_OUT_CAP; anything larger takes the global-memory path) and the
// decoded bytes leave in one coalesced copy.
constexpr uint32_t LZ_IN_CAP = 16 * 1024, LZ_OUT_CAP = 32 * 1024;

__global__ void __launch_bounds__(32) k_lz4_decompress(const uint8_t* __restrict__ blobs, const unsigned long long* __restrict__ blob_off,
                                                       uint8_t* out, const unsigned long long* __restrict__ out_off,
                                                       uint32_t* __restrict__ status, uint32_t* __restrict__ produced) {
    extern __shared__ __align__(16) uint8_t lz_smem[];
    const int layer = blockIdx.x;
    const uint32_t lane = threadIdx.x;
    const unsigned long long b0 = blob_off[layer], blen = blob_off[layer + 1] - b0;
    if (blen < 4) {  // empty input: Ok(empty) (tile_cache.rs:860-862); 1-3 bytes: the size prefix is missing
        if (lane == 0) status[layer] = blen == 0 ? 0u : 1u, produced[layer] = 0;
        return;
    }
    const uint32_t n = (uint32_t)(blen - 4);
    const uint32_t cap = (uint32_t)(out_off[layer + 1] - out_off[layer]);  // the size prefix, exactly
    const uint8_t* ip = blobs + b0 + 4;
    uint8_t* gout = out + out_off[layer];
    uint8_t* op = gout;
    if (n <= LZ_IN_CAP) {  // stage the block (bytewise: blobs are packed back to back, no alignment to rely on)
        for (uint32_t k = lane; k < n; k += 32) lz_smem[k] = ip[k];
        ip = lz_smem;
    }
    const bool out_staged = cap <= LZ_OUT_CAP;
    if (out_staged) op = lz_smem + LZ_IN_CAP;
    __syncwarp();
    uint32_t i = 0, o = 0;
    uint32_t st = 0;
    for (;;) {
        if (i >= n) { st = 1; break; }  // a token is expected
        const uint32_t token = ip[i++];
        uint32_t lit = token >> 4;
        if (lit == 15) {
            for (;;) {
                if (i >= n) { st = 1; break; }
                const uint32_t b = ip[i++];
                lit += b;
                if (b != 255) break;
            }
            if (st) break;
        }
        if (lit > n - i || lit > cap - o) { st = 1; break; }  // LiteralOutOfBounds / OutputTooSmall
        for (uint32_t k = lane; k < lit; k += 32) op[o + k] = ip[i + k];
        i += lit;
        o += lit;
        if (i >= n) break;  // end of block
        if (n - i < 2) { st = 1; break; }
        const uint32_t offset = (uint32_t)ip[i] | ((uint32_t)ip[i + 1] << 8);
        i += 2;
        uint32_t mlen = 4 + (token & 15u);
        if ((token & 15u) == 15u) {
            for (;;) {
                if (i >= n) { st = 1; break; }
                const uint32_t b = ip[i++];
                mlen += b;
                if (b != 255) break;
            }
            if (st) break;
        }
        if (offset > o || mlen > cap - o) { st = 1; break; }  // OffsetOutOfBounds / OutputTooSmall
        __syncwarp();  // the literals just written may be the match source
        if (offset == 0) {  // lz4_flex decodes into vec![0; size]: offset 0 re-reads bytes not written yet
            for (uint32_t k = lane; k < mlen; k += 32) op[o + k] = 0;
        } else {
            const uint8_t* src = op + o - offset;
            if (offset >= mlen) {
                for (uint32_t k = lane; k < mlen; k += 32) op[o + k] = src[k];
            } else if (offset >= 32) {  // every lane of a 32-byte step reads bytes produced before the step
                for (uint32_t k0 = 0; k0 < mlen; k0 += 32) {
                    if (k0 + lane < mlen) op[o + k0 + lane] = op[o + k0 + lane - offset];
                    __syncwarp();
                }
            } else {
                for (uint32_t k = lane; k < mlen; k += 32) op[o + k] = src[k % offset];
            }
        }
        o += mlen;
        __syncwarp();
    }
    if (out_staged && !st) {
        __syncwarp();
        for (uint32_t k = lane; k < o; k += 32) gout[k] = op[k];
    }
    if (lane == 0) status[layer] = st, produced[layer] = st ? 0u : o;
}

__device__ __forceinline__ uint32_t ld32(const uint8_t* p) {
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}

// One thread per layer: TileCacheLayer::from_bytes on the decoded bytes, then the checks
// decode_layer_to_heightfield makes before it touches a cell (tile_cache_builder.rs:138-143) and the
// add_span failure of hmin == 32767 (heightfield.rs:110-115), so that the host sees, per layer, the FIRST
// error the reference would return.
__global__ void __launch_bounds__(64) k_tc_parse(int nlayers, const uint8_t* __restrict__ raw, const unsigned long long* __restrict__ raw_off,
                                                 const uint32_t* __restrict__ lz_status, const uint32_t* __restrict__ produced,
                                                 RcbTcParsed* __restrict__ outp) {
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= nlayers) return;
    RcbTcParsed p{};
    const uint8_t* d = raw + raw_off[l];
    const uint32_t len = produced[l];
    if (lz_status[l]) {
        p.status = RCB_EDETOUR_FAILURE;
    } else if (len < 66 || ld32(d) != 0x4C494554u || ld32(d + 4) != 1u) {  // tile_cache_data.rs:282-285, 83-91
        p.status = RCB_EDETOUR_INVALID_PARAM;
    } else {
        const unsigned long long rl = ld32(d + 54), al = ld32(d + 58), cl = ld32(d + 62);
        if (66ull + rl + al + cl > len) {  // :311-313
            p.status = RCB_EDETOUR_INVALID_PARAM;
        } else {
            for (int k = 0; k < 3; ++k) p.bmin[k] = __uint_as_float(ld32(d + 20 + 4 * k)), p.bmax[k] = __uint_as_float(ld32(d + 32 + 4 * k));
            p.hmin = (uint32_t)d[44] | ((uint32_t)d[45] << 8);
            p.width = d[48];
            p.height = d[49];
            p.regons_off = 66;
            p.areas_off = 66 + (uint32_t)rl;
            const uint32_t cells = p.width * p.height;
            if (rl != cells || al != cells) {
                p.status = RCB_EINVALID_MESH;  // "Invalid region data size" / "Invalid area data size"
            } else if (p.hmin == 32767) {      // (32767 as i16, 32768 as i16): min > max on the first non-empty cell
                for (uint32_t c = 0; c < cells; ++c)
                    if (d[66 + c] != 0) {
                        p.status = RCB_ENAVMESH_GEN;
                        break;
                    }
            }
        }
    }
    outp[l] = p;
}

__global__ void __launch_bounds__(256) k_tc_gather(const TileDesc* __restrict__ tiles, TileMap tm, const uint8_t* __restrict__ raw,
                                                   const unsigned long long* __restrict__ raw_off, const RcbTcParsed* __restrict__ parsed,
                                                   uint8_t* __restrict__ regons, uint8_t* __restrict__ areas) {
    const uint32_t tile = RCB_TILE_OF_BLOCK();
    if (tile >= tm.ntiles) return;
    const TileDesc& td = tiles[tile];
    const uint32_t c = RCB_CELL_OF_THREAD();
    if (c >= (uint32_t)(td.width * td.height)) return;
    const uint8_t* d = raw + raw_off[tile];
    regons[td.cell_base + c] = d[parsed[tile].regons_off + c];
    areas[td.cell_base + c] = d[parsed[tile].areas_off + c];
}

}  // namespace

void launch_lz4_decompress(const uint8_t* blobs, const unsigned long long* blob_off, int nlayers, uint8_t* out,
                           const unsigned long long* out_off, uint32_t* status, uint32_t* produced, cudaStream_t s) {
    if (nlayers == 0) return;
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(k_lz4_decompress, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(LZ_IN_CAP + LZ_OUT_CAP));
        attr_set = true;
    }
    KScope _k("k_lz4_decompress", s);
    k_lz4_decompress<<<nlayers, 32, LZ_IN_CAP + LZ_OUT_CAP, s>>>(blobs, blob_off, out, out_off, status, produced);
}

void launch_tc_parse(int nlayers, const uint8_t* raw, const unsigned long long* raw_off, const uint32_t* lz_status,
                     const uint32_t* produced, RcbTcParsed* outp, cudaStream_t s) {
    if (nlayers == 0) return;
    KScope _k("k_tc_parse", s);
    k_tc_parse<<<(nlayers + 63) / 64, 64, 0, s>>>(nlayers, raw, raw_off, lz_status, produced, outp);
}

void launch_tc_gather(const TileDesc* tiles, TileMap tm, const uint8_t* raw, const unsigned long long* raw_off,
                      const RcbTcParsed* parsed, uint8_t* regons, uint8_t* areas, cudaStream_t s) {
    if (tm.nblocks == 0) return;
    KScope _k("k_tc_gather", s);
    k_tc_gather<<<tm.grid(), 256, 0, s>>>(tiles, tm, raw, raw_off, parsed, regons, areas);
}
