// k_regions.cu — region building: watershed::build_regions_watershed
// (crates/recast/src/watershed.rs:364-510), the function behind CompactHeightfield::build_regions
// (compact_heightfield.rs:770-778) that RecastBuilder::build_mesh runs right after the distance field
// (lib.rs:83-87).  One CTA per tile, all tiles of the batch in one launch.
//
// The reference is a sequential program; what makes a parallel restatement possible, step by step:
//   sort_cells_by_level / append_stacks (:65-133)  ordered filters of span lists -> block-wide ordered compaction
//   expand_regions (:245-361)                      gathers from the labels of the PREVIOUS pass (the `dirty_entries`
//                                                  list is applied after the loop): every entry in parallel
//   flood_region (:136-242)                        whether a cell stays in the region depends only on labels of
//                                                  OTHER regions, which do not change during the flood, so the DFS
//                                                  order is immaterial: breadth-first waves, cells claimed with a
//                                                  16-bit compare-and-swap.  Seeds are tried in stack order; a seed
//                                                  that fails now fails for good (other regions only grow), so the
//                                                  next flood starts from the first entry that is still unlabelled
//                                                  and passes the test.
//   merge_and_filter_regions (:543-746)            per-region counts by atomics; the connection LIST of a region only
//                                                  matters through "first neighbour with the largest span count",
//                                                  i.e. the maximum of (count, -first occurrence) — one 64-bit
//                                                  atomicMax per region; merges of a round are decided together and
//                                                  applied in index order, which a per-region pointer chase reproduces.
// flood / expand reach neighbours through span.con[] (63 = not connected, even when it is a real offset);
// merge_and_filter goes through get_neighbor_connection -> the connection list, as the reference does.
#include "rcb_common.cuh"
#define RCB_FILE_ID 3
#include "rcb_bounds.cuh"

// Two builds of the same kernel.  A tile is one long dependent chain and the kernel ends with its slowest tile, so what
// matters is that every tile of the batch is resident at once:
//   64 threads, 14 CTAs per SM (<= 73 registers) — labels in shared memory; the 2025 tiles of config 2 in one wave
//   32 threads, 32 CTAs per SM (64 registers)    — labels in global memory; the 4761 128x128 tiles of config 3 in two
//   256 / 1024 threads                            — a few very large tiles (config 4: 144 tiles of 56 k spans): wider phases
#define RG_BT 64
#define RG_MIN_CTAS 14
#define RG_NS rg64
#include "k_regions_impl.cuh"
#undef RG_BT
#undef RG_MIN_CTAS
#undef RG_NS
#define RG_BT 32
#define RG_MIN_CTAS 32
#define RG_NS rg32
#include "k_regions_impl.cuh"
#undef RG_BT
#undef RG_MIN_CTAS
#undef RG_NS
#define RG_BT 256
#define RG_MIN_CTAS 2
#define RG_NS rg256
#include "k_regions_impl.cuh"
#undef RG_BT
#undef RG_MIN_CTAS
#undef RG_NS
#define RG_BT 1024
#define RG_MIN_CTAS 1
#define RG_NS rg1024
#include "k_regions_impl.cuh"
#undef RG_BT
#undef RG_MIN_CTAS
#undef RG_NS

size_t regions_smem_bytes(uint32_t max_spans) { return 2 * (size_t)((max_spans + 7u) & ~7u) + 16; }

void launch_regions(const RegionArgs& a, int ntiles, cudaStream_t s) {
    if (ntiles == 0) return;
    const size_t smem = a.smem_spans ? regions_smem_bytes(a.smem_spans) : 0;
    KScope _k("k_regions", s);
    if (a.narrow == 1) {
        rg32::k_regions<<<ntiles, 32, smem, s>>>(a);
    } else if (a.narrow == 3) {
        rcb_raise_dynamic_smem((const void*)rg1024::k_regions, smem);
        rg1024::k_regions<<<ntiles, 1024, smem, s>>>(a);
    } else if (a.narrow == 2) {
        rcb_raise_dynamic_smem((const void*)rg256::k_regions, smem);
        rg256::k_regions<<<ntiles, 256, smem, s>>>(a);
    } else {
        rcb_raise_dynamic_smem((const void*)rg64::k_regions, smem);
        rg64::k_regions<<<ntiles, 64, smem, s>>>(a);
    }
}

RCB_BOUNDS_EXPORT(rcb_bounds_read_regions)
