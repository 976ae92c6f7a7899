//! Raw bindings to include/recast_b200.h (hand-written; keep in sync with the header).
//! Source only — never compiled in the build image (no Rust toolchain there).
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

pub const RCB_NONE: u32 = 0xFFFF_FFFF;
pub const RCB_OK: c_int = 0;
pub const RCB_EINVALID_MESH: c_int = -1;
pub const RCB_ENAVMESH_GEN: c_int = -2;
pub const RCB_ECUDA: c_int = -3;
pub const RCB_EOVERFLOW: c_int = -4;
pub const RCB_EARG: c_int = -5;
pub const RCB_ERECAST: c_int = -6;
pub const RCB_EDETOUR_FAILURE: c_int = -7;
pub const RCB_EDETOUR_INVALID_PARAM: c_int = -8;
pub const RCB_STAGE_RASTER: u32 = 1;
pub const RCB_STAGE_FILTER: u32 = 2;
pub const RCB_STAGE_COMPACT: u32 = 4;
pub const RCB_STAGE_ERODE: u32 = 8;
pub const RCB_STAGE_DIST: u32 = 16;
pub const RCB_STAGE_READD: u32 = 0x20;
pub const RCB_STAGE_REGIONS: u32 = 0x40;
pub const RCB_WIRE_VOXEL_TILE: c_int = 0;
pub const RCB_WIRE_DYNAMIC_TILE: c_int = 1;
pub const RCB_FILTER_LOW_HANGING: u32 = 0x100;
pub const RCB_FILTER_LEDGE: u32 = 0x200;
pub const RCB_FILTER_LOW_HEIGHT: u32 = 0x400;
pub const RCB_RESULT_SLIM: u32 = 0x1000;

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct rcb_config {
    pub width: i32,
    pub height: i32,
    pub cs: f32,
    pub ch: f32,
    pub bmin: [f32; 3],
    pub bmax: [f32; 3],
    pub walkable_slope_angle: f32,
    pub walkable_height: i32,
    pub walkable_climb: i32,
    pub walkable_radius: i32,
    pub max_edge_len: i32,
    pub max_simplification_error: f32,
    pub min_region_area: i32,
    pub merge_region_area: i32,
    pub max_vertices_per_polygon: i32,
    pub detail_sample_dist: f32,
    pub detail_sample_max_error: f32,
    pub border_size: i32,
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct rcb_totals {
    pub tiles: u64,
    pub triangles: u64,
    pub vertices: u64,
    pub cells: u64,
    pub fragments: u64,
    pub spans: u64,
    pub connections: u64,
    pub walkable_spans: u64,
    pub result_bytes: u64,
}

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct rcb_tile_view {
    pub width: i32,
    pub height: i32,
    pub cell_count: u32,
    pub span_count: u32,
    pub conn_count: u32,
    pub walkable_span_count: u32,
    pub max_height: u16,
    pub max_distance: u16,
    pub status: u32,
    pub hf_cell_offset: *const u32,
    pub span_min: *const i16,
    pub span_max: *const i16,
    pub span_area: *const u8,
    pub cell_index: *const u32,
    pub cell_count_arr: *const u32,
    pub areas: *const u8,
    pub con: *const u16,
    pub first_conn: *const u32,
    pub conn_ref: *const u32,
    pub conn_dir: *const u8,
    pub conn_next: *const u32,
    pub dist: *const u16,
    pub reg: *const u16,
    pub max_regions: u32,
    pub _pad: u32,
    pub conn_mask: *const u8,
    pub conn_ref16: *const u16,
    pub cell_count16: *const u16,
    pub con8: *const u8,
    pub conn_off8: *const u8,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct rcb_stage_time {
    pub name: [c_char; 40],
    pub ms: f32,
    pub calls: u32,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct rcb_area_op {
    pub type_: i32,
    pub area_id: u32,
    pub a: [f32; 3],
    pub b: [f32; 3],
    pub first_vert: u32,
    pub num_verts: u32,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct rcb_tc_layer {
    pub bmin: [f32; 3],
    pub bmax: [f32; 3],
    pub hmin: u16,
    pub hmax: u16,
    pub width: u8,
    pub height: u8,
    pub _pad: u16,
    pub regons: *const u8,
    pub areas: *const u8,
    pub regons_len: u32,
    pub areas_len: u32,
    pub first_obstacle: u32,
    pub obstacle_count: u32,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct rcb_tc_obstacle {
    pub type_: i32,
    pub state: i32,
    pub a: [f32; 3],
    pub b: [f32; 3],
    pub rot_aux: [f32; 2],
}

#[repr(C)]
pub struct rcb_ctx {
    _private: [u8; 0],
}
#[repr(C)]
pub struct rcb_result {
    _private: [u8; 0],
}
#[repr(C)]
pub struct rcb_multi {
    _private: [u8; 0],
}
#[repr(C)]
pub struct rcb_multi_result {
    _private: [u8; 0],
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct rcb_convex_volume {
    pub first_vert: u32,
    pub num_verts: u32,
    pub min_height: f32,
    pub max_height: f32,
    pub area_type: u32,
}

extern "C" {
    pub fn rcb_create(device: c_int, out: *mut *mut rcb_ctx) -> c_int;
    pub fn rcb_destroy(ctx: *mut rcb_ctx);
    pub fn rcb_last_error(ctx: *const rcb_ctx) -> *const c_char;
    pub fn rcb_set_stream(ctx: *mut rcb_ctx, cuda_stream: *mut c_void) -> c_int;
    pub fn rcb_use_own_stream(ctx: *mut rcb_ctx) -> c_int;
    pub fn rcb_launch_count(ctx: *const rcb_ctx) -> u64;
    pub fn rcb_ctx_stats(ctx: *const rcb_ctx, out: *mut u64) -> c_int;
    pub fn rcb_debug_bounds_report(ctx: *mut rcb_ctx, out: *mut u64, reset: c_int) -> c_int;
    pub fn rcb_profile_enable(ctx: *mut rcb_ctx, on: c_int) -> c_int;
    pub fn rcb_profile_read(ctx: *mut rcb_ctx, out: *mut rcb_stage_time, cap: c_int) -> c_int;
    pub fn rcb_profile_reset(ctx: *mut rcb_ctx) -> c_int;
    pub fn rcb_config_default(cfg: *mut rcb_config);
    pub fn rcb_config_calculate_grid_size(cfg: *mut rcb_config, bmin: *const f32, bmax: *const f32);
    pub fn rcb_config_validate(cfg: *const rcb_config) -> c_int;
    pub fn rcb_upload_mesh(ctx: *mut rcb_ctx, verts: *const f32, nfloats: usize, tris: *const i32, nidx: usize) -> c_int;
    pub fn rcb_upload_mesh_u16(ctx: *mut rcb_ctx, verts: *const f32, nfloats: usize, tris: *const u16, nidx: usize) -> c_int;
    pub fn rcb_upload_triangle_soup(ctx: *mut rcb_ctx, verts: *const f32, ntris: usize) -> c_int;
    pub fn rcb_set_mesh_device(ctx: *mut rcb_ctx, d_verts: *const c_void, nfloats: usize, d_tris: *const c_void, nidx: usize) -> c_int;
    pub fn rcb_build_tiles(ctx: *mut rcb_ctx, cfgs: *const rcb_config, ntiles: c_int, stage_mask: u32, erode_radius: c_int, out: *mut *mut rcb_result) -> c_int;
    pub fn rcb_build_tiles_streamed(ctx: *mut rcb_ctx, cfgs: *const rcb_config, ntiles: c_int, stage_mask: u32, erode_radius: c_int, ngroups: c_int, out: *mut *mut rcb_result) -> c_int;
    pub fn rcb_rasterize(ctx: *mut rcb_ctx, cfgs: *const rcb_config, ntiles: c_int, tri_areas: *const u8, flag_merge_threshold: i32, stage_mask: u32, erode_radius: c_int, out: *mut *mut rcb_result) -> c_int;
    pub fn rcb_rasterize_onto(ctx: *mut rcb_ctx, cfgs: *const rcb_config, ntiles: c_int, cell_offset: *const u32, span_min: *const i16, span_max: *const i16, span_area: *const u8, tri_areas: *const u8, flag_merge_threshold: i32, stage_mask: u32, erode_radius: c_int, out: *mut *mut rcb_result) -> c_int;
    pub fn rcb_build_from_spans(ctx: *mut rcb_ctx, cfgs: *const rcb_config, ntiles: c_int, cell_offset: *const u32, span_min: *const i16, span_max: *const i16, span_area: *const u8, areas_override: *const u8, stage_mask: u32, erode_radius: c_int, out: *mut *mut rcb_result) -> c_int;
    pub fn rcb_apply_area_ops(ctx: *mut rcb_ctx, cfgs: *const rcb_config, ntiles: c_int, cell_offset: *const u32, span_min: *const i16, span_max: *const i16, span_area: *const u8, areas_override: *const u8, ops: *const rcb_area_op, nops: c_int, poly_verts: *const f32, npoly_floats: usize, stage_mask: u32, erode_radius: c_int, out: *mut *mut rcb_result) -> c_int;
    pub fn rcb_build_tilecache_layers(ctx: *mut rcb_ctx, layers: *const rcb_tc_layer, nlayers: c_int, obstacles: *const rcb_tc_obstacle, nobstacles: c_int, cs: f32, ch: f32, stage_mask: u32, erode_radius: c_int, out: *mut *mut rcb_result) -> c_int;
    pub fn rcb_build_tiles_with_volumes(ctx: *mut rcb_ctx, cfgs: *const rcb_config, ntiles: c_int, stage_mask: u32, erode_radius: c_int, volumes: *const rcb_convex_volume, nvolumes: c_int, poly_verts: *const f32, npoly_floats: usize, out: *mut *mut rcb_result) -> c_int;
    pub fn rcb_apply_convex_volumes(ctx: *mut rcb_ctx, cfgs: *const rcb_config, ntiles: c_int, cell_offset: *const u32, span_min: *const i16, span_max: *const i16, span_area: *const u8, volumes: *const rcb_convex_volume, nvolumes: c_int, poly_verts: *const f32, npoly_floats: usize, stage_mask: u32, erode_radius: c_int, out: *mut *mut rcb_result) -> c_int;
    pub fn rcb_build_from_span_data(ctx: *mut rcb_ctx, cfgs: *const rcb_config, ntiles: c_int, data: *const u8, tile_offsets: *const u64, format: c_int, stage_mask: u32, erode_radius: c_int, out: *mut *mut rcb_result) -> c_int;
    pub fn rcb_result_span_data_size(res: *const rcb_result) -> u64;
    pub fn rcb_result_serialize_spans(res: *mut rcb_result, format: c_int, out: *mut u8, capacity: u64, tile_offsets: *mut u64) -> c_int;
    pub fn rcb_decompress_tiles(ctx: *mut rcb_ctx, blobs: *const u8, blob_offsets: *const u64, n: c_int, out: *mut u8, out_cap: u64, out_offsets: *mut u64) -> c_int;
    pub fn rcb_build_tilecache_compressed(ctx: *mut rcb_ctx, blobs: *const u8, blob_offsets: *const u64, nlayers: c_int, first_obstacle: *const u32, obstacle_count: *const u32, obstacles: *const rcb_tc_obstacle, nobstacles: c_int, cs: f32, ch: f32, stage_mask: u32, erode_radius: c_int, out: *mut *mut rcb_result) -> c_int;
    pub fn rcb_result_totals(res: *const rcb_result, out: *mut rcb_totals) -> c_int;
    pub fn rcb_result_download(res: *mut rcb_result) -> c_int;
    pub fn rcb_result_tile(res: *const rcb_result, tile: c_int, out: *mut rcb_tile_view) -> c_int;
    pub fn rcb_result_tile_device(res: *const rcb_result, tile: c_int, out: *mut rcb_tile_view) -> c_int;
    pub fn rcb_result_free(res: *mut rcb_result);
    pub fn rcb_multi_create(devices: *const c_int, ndevices: c_int, out: *mut *mut rcb_multi) -> c_int;
    pub fn rcb_multi_destroy(m: *mut rcb_multi);
    pub fn rcb_multi_last_error(m: *const rcb_multi) -> *const c_char;
    pub fn rcb_multi_device_count(m: *const rcb_multi) -> c_int;
    pub fn rcb_multi_upload_mesh(m: *mut rcb_multi, verts: *const f32, nfloats: usize, tris: *const i32, nidx: usize) -> c_int;
    pub fn rcb_multi_build_tiles(m: *mut rcb_multi, cfgs: *const rcb_config, ntiles: c_int, stage_mask: u32, erode_radius: c_int, tiles_per_group: c_int, download: c_int, out: *mut *mut rcb_multi_result) -> c_int;
    pub fn rcb_multi_result_download(res: *mut rcb_multi_result) -> c_int;
    pub fn rcb_multi_result_totals(res: *const rcb_multi_result, out: *mut rcb_totals, tiles_per_device: *mut c_int, ndevices: c_int) -> c_int;
    pub fn rcb_multi_result_tile(res: *const rcb_multi_result, tile: c_int, out: *mut rcb_tile_view) -> c_int;
    pub fn rcb_multi_result_tile_device(res: *const rcb_multi_result, tile: c_int, out: *mut rcb_tile_view, device: *mut c_int) -> c_int;
    pub fn rcb_multi_result_free(res: *mut rcb_multi_result);
}
