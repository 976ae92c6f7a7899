"""RecastConfig — field-for-field mirror of the reference's `RecastConfig`
(crates/recast/src/config.rs:7-51) and its C-ABI twin `rcb_config` (include/recast_b200.h).

Pure data: importing this module needs neither CUDA nor the native library.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass, field
from typing import Sequence


class RcbConfig(ctypes.Structure):
    """`rcb_config` of include/recast_b200.h (same order as config.rs:7-51)."""

    _fields_ = [
        ("width", ctypes.c_int32),
        ("height", ctypes.c_int32),
        ("cs", ctypes.c_float),
        ("ch", ctypes.c_float),
        ("bmin", ctypes.c_float * 3),
        ("bmax", ctypes.c_float * 3),
        ("walkable_slope_angle", ctypes.c_float),
        ("walkable_height", ctypes.c_int32),
        ("walkable_climb", ctypes.c_int32),
        ("walkable_radius", ctypes.c_int32),
        ("max_edge_len", ctypes.c_int32),
        ("max_simplification_error", ctypes.c_float),
        ("min_region_area", ctypes.c_int32),
        ("merge_region_area", ctypes.c_int32),
        ("max_vertices_per_polygon", ctypes.c_int32),
        ("detail_sample_dist", ctypes.c_float),
        ("detail_sample_max_error", ctypes.c_float),
        ("border_size", ctypes.c_int32),
    ]


@dataclass
class RecastConfig:
    """Mirror of `RecastConfig` (config.rs:7-51) with the defaults of config.rs:53-76."""

    width: int = 0
    height: int = 0
    cs: float = 0.3
    ch: float = 0.2
    bmin: Sequence[float] = field(default_factory=lambda: (0.0, 0.0, 0.0))
    bmax: Sequence[float] = field(default_factory=lambda: (0.0, 0.0, 0.0))
    walkable_slope_angle: float = 45.0
    walkable_height: int = 2
    walkable_climb: int = 1
    walkable_radius: int = 1
    max_edge_len: int = 12
    max_simplification_error: float = 1.3
    min_region_area: int = 8
    merge_region_area: int = 20
    max_vertices_per_polygon: int = 6
    detail_sample_dist: float = 6.0
    detail_sample_max_error: float = 1.0
    border_size: int = 0

    @classmethod
    def new(cls) -> "RecastConfig":
        return cls()

    # -- C struct round trip ---------------------------------------------------------------------
    def to_c(self) -> RcbConfig:
        c = RcbConfig()
        for name, _ in RcbConfig._fields_:
            v = getattr(self, name)
            if name in ("bmin", "bmax"):
                setattr(c, name, (ctypes.c_float * 3)(*[float(x) for x in v]))
            else:
                setattr(c, name, v)
        return c

    @classmethod
    def from_c(cls, c: RcbConfig) -> "RecastConfig":
        kw = {}
        for name, _ in RcbConfig._fields_:
            v = getattr(c, name)
            kw[name] = tuple(v) if name in ("bmin", "bmax") else v
        return cls(**kw)

    # -- config.rs:85-107 / :110-136: the arithmetic lives in the native library (f32, libm) -----
    def calculate_grid_size(self, bmin: Sequence[float], bmax: Sequence[float]) -> None:
        from . import _ffi

        c = self.to_c()
        _ffi.lib().rcb_config_calculate_grid_size(
            ctypes.byref(c), (ctypes.c_float * 3)(*bmin), (ctypes.c_float * 3)(*bmax)
        )
        self.bmin, self.bmax = tuple(c.bmin), tuple(c.bmax)
        self.width, self.height = c.width, c.height

    def validate(self) -> None:
        from . import _ffi
        from .errors import check

        c = self.to_c()
        check(_ffi.lib().rcb_config_validate(ctypes.byref(c)), None, "Invalid configuration")


def configs_to_array(cfgs: Sequence[RecastConfig | RcbConfig]):
    """Pack a list of configs into a contiguous `rcb_config[]`."""
    arr = (RcbConfig * len(cfgs))()
    for i, c in enumerate(cfgs):
        arr[i] = c if isinstance(c, RcbConfig) else c.to_c()
    return arr
