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

// ---- one sequence header -----------------------------------------------------------------------------------
struct LzSeq {
    uint32_t lit_start, lit, off, mlen, next;
    int kind;  // 0 literals + match, 1 last sequence (literals only, input exhausted), 2 lz4_flex's Err while parsing
};
// max_chain bounds the 255-continuation loops (speculative parses of arbitrary bytes must stay cheap);
// kind 3 = "longer than that, parse me for real".
__device__ __forceinline__ LzSeq lz_parse(const uint8_t* ip, uint32_t n, uint32_t p, uint32_t max_chain) {
    LzSeq s{0, 0, 0, 0, 0, 2};
    if (p >= n) return s;  // a token is expected
    const uint32_t token = ip[p];
    uint32_t q = p + 1, lit = token >> 4;
    if (lit == 15) {
        for (uint32_t c = 0;; ++c) {
            if (q >= n) return s;
            if (c >= max_chain) { s.kind = 3; return s; }
            const uint32_t b = ip[q++];
            lit += b;
            if (b != 255) break;
        }
    }
    if (lit > n - q) return s;  // LiteralOutOfBounds
    s.lit_start = q;
    s.lit = lit;
    uint32_t r = q + lit;
    if (r >= n) { s.kind = 1; return s; }  // end of block
    if (n - r < 2) return s;
    s.off = (uint32_t)ip[r] | ((uint32_t)ip[r + 1] << 8);
    r += 2;
    uint32_t mlen = 4 + (token & 15u);
    if ((token & 15u) == 15u) {
        for (uint32_t c = 0;; ++c) {
            if (r >= n) return s;
            if (c >= max_chain) { s.kind = 3; return s; }
            const uint32_t b = ip[r++];
            mlen += b;
            if (b != 255) break;
        }
    }
    s.mlen = mlen;
    s.next = r;
    s.kind = 0;
    return s;
}

// copies of one sequence by the whole warp; `op` may alias nothing but itself
__device__ __forceinline__ void lz_copy(const uint8_t* ip, uint8_t* op, uint32_t lane, uint32_t o, uint32_t lit_start, uint32_t lit,
                                        bool has_match, uint32_t offset, uint32_t mlen) {
    if (lit) {
        for (uint32_t k = lane; k < lit; k += 32) op[o + k] = ip[lit_start + k];
        __syncwarp();  // the literals just written may be the match source
    }
    if (!has_match) return;
    o += lit;
    if (offset == 0) {  // lz4_flex decodes into vec![0; size]: offset 0 re-reads bytes not written yet
        for (uint32_t k = lane; k < mlen; k += 32) op[o + k] = 0;
    } else if (offset >= mlen) {
        for (uint32_t k = lane; k < mlen; k += 32) op[o + k] = op[o - offset + k];
    } else if (offset >= 32) {  // every lane of a 32-byte step reads bytes produced before the step
        for (uint32_t k0 = 0; k0 < mlen; k0 += 32) {
            if (k0 + lane < mlen) op[o + k0 + lane] = op[o + k0 + lane - offset];
            __syncwarp();
        }
    } else {
        for (uint32_t k = lane; k < mlen; k += 32) op[o + k] = op[o - offset + k % offset];
    }
    __syncwarp();
}

// status per layer: 0 ok, 1 = lz4_flex's Err (Error::Detour(Failure))
//
// Every step of a sequence depends on the one before (token -> lengths -> literals -> offset -> match), and one
// warp runs that chain at ~800 cycles per sequence when it is written as one loop.  Blocks that fit shared
// memory (input <= LZ_IN_CAP, size prefix <= LZ_OUT_CAP) are therefore decoded in phases:
//   A  every byte position is parsed as if a sequence started there (parallel): next[p]
//   B  the real sequence starts are the chain 0, next[0], next[next[0]], ... (one shared-memory load per step)
//   C  per sequence (parallel): lengths -> output positions by a warp scan; all of lz4_flex's bound checks
//   D  copies in sequence order, headers decoded 32 sequences at a time and handed round by shuffles
// Larger blocks take the one-loop path in global memory.
constexpr uint32_t LZ_IN_CAP = 8 * 1024, LZ_OUT_CAP = 32 * 1024, LZ_MAX_SEQ = LZ_IN_CAP / 3 + 2;
constexpr uint32_t LZ_SMEM = LZ_IN_CAP + LZ_OUT_CAP + 2 * (LZ_IN_CAP + 8) + 2 * 2 * ((LZ_MAX_SEQ + 7) & ~7u);

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
    const uint32_t cap = (uint32_t)(out_off[layer + 1] - out_off[layer]);  // the size prefix (or LZ4's 255x ceiling on this input, if that is smaller)
    const uint8_t* gin = blobs + b0 + 4;
    uint8_t* gout = out + out_off[layer];
    uint32_t st = 0, o = 0;
    if (n <= LZ_IN_CAP && cap <= LZ_OUT_CAP) {
        uint8_t* ip = lz_smem;
        uint8_t* op = lz_smem + LZ_IN_CAP;
        uint16_t* next = (uint16_t*)(lz_smem + LZ_IN_CAP + LZ_OUT_CAP);
        uint16_t* seq_pos = next + LZ_IN_CAP + 8;
        uint16_t* out_pos = seq_pos + ((LZ_MAX_SEQ + 7) & ~7u);
        for (uint32_t k = lane; k < n; k += 32) ip[k] = gin[k];  // bytewise: blobs are packed back to back
        __syncwarp();
        // A
        for (uint32_t p = lane; p < n; p += 32) {
            const LzSeq s = lz_parse(ip, n, p, 8);
            next[p] = s.kind == 0 ? (uint16_t)s.next : (uint16_t)(0xFFFC + s.kind);  // 0xFFFD last, 0xFFFE error, 0xFFFF long chain
        }
        __syncwarp();
        // B (uniform: every lane walks the same chain; next[n] is the "token expected" error)
        if (lane == 0) next[n] = 0xFFFE;
        __syncwarp();
        uint32_t ns = 0;
        for (uint32_t p = 0;;) {
            uint32_t nx = next[p];
            if (nx >= 0xFFFD) {
                if (nx == 0xFFFF) {
                    const LzSeq s = lz_parse(ip, n, p, 0xffffffffu);
                    nx = s.kind == 0 ? s.next : 0xFFFCu + s.kind;
                }
                if (nx == 0xFFFE) { st = 1; break; }
                if (nx == 0xFFFD) {
                    if (lane == 0) seq_pos[ns] = (uint16_t)p;
                    ++ns;
                    break;
                }
            }
            if (lane == 0) seq_pos[ns] = (uint16_t)p;
            ++ns;
            p = nx;
        }
        __syncwarp();
        if (!st) {
            // C
            uint32_t carry = 0;
            bool bad = false;
            for (uint32_t base = 0; base < ns; base += 32) {
                const uint32_t k = base + lane;
                LzSeq s{0, 0, 0, 0, 0, 1};
                if (k < ns) s = lz_parse(ip, n, seq_pos[k], 0xffffffffu);
                const uint32_t lit = min(s.lit, cap + 1), ml = s.kind == 0 ? min(s.mlen, cap + 1) : 0;
                uint32_t incl = k < ns ? lit + ml : 0;
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t v = __shfl_up_sync(0xffffffffu, incl, d);
                    if (lane >= (uint32_t)d) incl += v;
                }
                const uint32_t ok = carry + incl - (k < ns ? lit + ml : 0);  // output position of this sequence
                if (k < ns) {
                    // LiteralOutOfBounds was checked by the parse; OutputTooSmall, OffsetOutOfBounds:
                    if (ok > cap || lit > cap - ok) bad = true;
                    else if (s.kind == 0 && (s.off > ok + lit || ml > cap - (ok + lit))) bad = true;
                    out_pos[k] = (uint16_t)min(ok, 0xffffu);
                }
                carry += __shfl_sync(0xffffffffu, incl, 31);
            }
            if (__any_sync(0xffffffffu, bad)) st = 1;
            o = carry;
            __syncwarp();
        }
        if (!st) {
            // D: the copies are the serial part (each may read what the one before wrote), so the loop body is
            // kept short: three packed words per sequence, fetched one sequence ahead of the copy that uses them
            for (uint32_t base = 0; base < ns; base += 32) {
                const uint32_t k = base + lane;
                LzSeq s{0, 0, 0, 0, 0, 1};
                uint32_t my_o = 0;
                if (k < ns) s = lz_parse(ip, n, seq_pos[k], 0xffffffffu), my_o = out_pos[k];
                // two packed words per sequence: output position | offset, and match length | fast flag | magic.
                // fast = no literals, a match of 1..32 bytes, offset != 0; magic M = 1024/off + 1 turns
                // lane % off into lane - off * ((lane * M) >> 10) for off, lane < 32.
                const bool fast = s.lit == 0 && s.kind == 0 && s.mlen <= 32 && s.off != 0;
                const uint32_t magic = (s.off != 0 && s.off < 32) ? 1024u / s.off + 1u : 0u;  // 0: lane % off == lane
                const uint32_t w0 = my_o | (s.off << 16), w1 = min(s.mlen, 0xffffu) | ((fast ? 1u : 0u) << 16) | (magic << 17);
                const uint32_t cnt = min(32u, ns - base);
                uint32_t n0 = __shfl_sync(0xffffffffu, w0, 0), n1 = __shfl_sync(0xffffffffu, w1, 0);
                for (uint32_t j = 0; j < cnt; ++j) {
                    const uint32_t c0 = n0, c1 = n1;
                    const uint32_t jn = min(j + 1, cnt - 1);
                    n0 = __shfl_sync(0xffffffffu, w0, jn), n1 = __shfl_sync(0xffffffffu, w1, jn);
                    const uint32_t so = c0 & 0xffffu, off = c0 >> 16;
                    if (c1 & 0x10000u) {  // the common case: one short match
                        const uint32_t d = lane - off * ((lane * (c1 >> 17)) >> 10);
                        if (lane < (c1 & 0xffffu)) op[so + lane] = op[so - off + d];
                        __syncwarp();
                    } else {
                        const uint32_t lit_start = __shfl_sync(0xffffffffu, s.lit_start, j), lit = __shfl_sync(0xffffffffu, s.lit, j);
                        const int kind = __shfl_sync(0xffffffffu, s.kind, j);
                        lz_copy(ip, op, lane, so, lit_start, lit, kind == 0, off, c1 & 0xffffu);
                    }
                }
            }
            __syncwarp();
            for (uint32_t k = lane; k < o; k += 32) gout[k] = op[k];
        }
    } else {
        // one-loop decoder in global memory (headers read by every lane from the same address)
        uint32_t p = 0;
        for (;;) {
            const LzSeq s = lz_parse(gin, n, p, 0xffffffffu);
            if (s.kind == 2) { st = 1; break; }
            if (s.lit > cap - o) { st = 1; break; }  // OutputTooSmall
            if (s.kind == 0 && (s.off > o + s.lit || s.mlen > cap - (o + s.lit))) { st = 1; break; }  // OffsetOutOfBounds / OutputTooSmall
            __syncwarp();
            lz_copy(gin, gout, lane, o, s.lit_start, s.lit, s.kind == 0, s.off, s.mlen);
            o += s.lit + (s.kind == 0 ? s.mlen : 0);
            if (s.kind == 1) break;
            p = s.next;
        }
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
    rcb_raise_dynamic_smem((const void*)k_lz4_decompress, LZ_SMEM);  // per (kernel, device); a process may drive several
    KScope _k("k_lz4_decompress", s);
    k_lz4_decompress<<<nlayers, 32, LZ_SMEM, s>>>(blobs, blob_off, out, out_off, status, produced);
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
