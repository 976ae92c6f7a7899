"""recast_rs_b200 — B200-native tiled voxelization front end for recast-rs.

`RecastBuilder` / `RecastConfig` / `Heightfield` / `CompactHeightfield` mirror the reference's API
(crates/recast/src/lib.rs) over the C ABI of include/recast_b200.h; the compute path is hand-written
CUDA for sm_100a in recast_rs_b200/csrc.  There is no CPU fallback.
"""
from .config import RecastConfig, RcbConfig  # noqa: F401
from .errors import RecastError, InvalidMesh, NavMeshGeneration, CudaError, PoolOverflow  # noqa: F401


def __getattr__(name):  # lazy: importing the package must not need the native library
    if name in ("RecastBuilder", "Heightfield", "CompactHeightfield", "Context", "BatchResult", "TileVoxels",
                "erode_walkable_area", "STAGE_RASTER", "STAGE_FILTER", "STAGE_COMPACT", "STAGE_ERODE", "STAGE_DIST",
                "STAGE_ALL_BUILD", "FILTER_LOW_HANGING", "FILTER_LEDGE", "FILTER_LOW_HEIGHT", "STAGE_READD", "STAGE_REGIONS", "WIRE_VOXEL_TILE",
                "WIRE_DYNAMIC_TILE"):
        from . import builder

        return getattr(builder, name)
    raise AttributeError(name)
