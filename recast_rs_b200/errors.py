"""Error mapping: recast_common::Error (crates/recast-common/src/lib.rs:19-43) <-> rcb_status."""
from __future__ import annotations

RCB_OK = 0
RCB_EINVALID_MESH = -1
RCB_ENAVMESH_GEN = -2
RCB_ECUDA = -3
RCB_EOVERFLOW = -4
RCB_EARG = -5
RCB_ERECAST = -6
RCB_EDETOUR_FAILURE = -7
RCB_EDETOUR_INVALID_PARAM = -8


class RecastError(Exception):
    """Base of every error raised by this package (reference: recast_common::Error)."""


class InvalidMesh(RecastError):
    """Error::InvalidMesh — config.rs:113-133."""


class NavMeshGeneration(RecastError):
    """Error::NavMeshGeneration — lib.rs:199-233, heightfield.rs:103-115."""


class Recast(RecastError):
    """Error::Recast — detour-dynamic/src/dynamic_tile.rs:171-216 (short span data)."""


class DetourFailure(RecastError):
    """Error::Detour(Status::Failure) — LZ4 decode failed, detour-tilecache/src/tile_cache.rs:865-870."""


class DetourInvalidParam(RecastError):
    """Error::Detour(Status::InvalidParam) — TileCacheLayer::from_bytes, tile_cache_data.rs:83-91, 129, 282-313."""


class CudaError(RecastError):
    """CUDA runtime failure or missing device/extension; there is no CPU fallback."""


class PoolOverflow(RecastError):
    """An internal device pool overflowed (clip polygon vertex capacity)."""


_MAP = {
    RCB_EINVALID_MESH: InvalidMesh,
    RCB_ENAVMESH_GEN: NavMeshGeneration,
    RCB_ECUDA: CudaError,
    RCB_EOVERFLOW: PoolOverflow,
    RCB_EARG: ValueError,
    RCB_ERECAST: Recast,
    RCB_EDETOUR_FAILURE: DetourFailure,
    RCB_EDETOUR_INVALID_PARAM: DetourInvalidParam,
}


def check(status: int, ctx=None, what: str = "") -> None:
    """Raise the exception matching a negative rcb_status."""
    if status == RCB_OK:
        return
    msg = what
    if ctx is not None:
        from . import _ffi

        raw = _ffi.lib().rcb_last_error(ctx)
        if raw:
            msg = raw.decode("utf-8", "replace")
    raise _MAP.get(status, RecastError)(msg or f"rcb_status {status}")
