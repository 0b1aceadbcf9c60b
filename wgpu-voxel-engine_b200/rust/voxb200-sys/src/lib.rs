//! Raw bindings to `libvoxb200.so`; one declaration per symbol of `include/voxb200.h`.
//! NOT COMPILED IN THIS REPOSITORY'S CI (the build image has no rustc); kept in
//! lock-step with the header by tests/test_abi_cpu.py::test_rust_bindings_cover_the_header.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

pub const VX_OK: c_int = 0;
pub const VX_ERR_BAD_ARG: c_int = -1;
pub const VX_ERR_CAPACITY: c_int = -2;
pub const VX_ERR_CUDA: c_int = -3;
pub const VX_ERR_NO_DEVICE: c_int = -4;
pub const VX_ERR_OOM: c_int = -5;
pub const VX_ERR_BAD_BLOCK: c_int = -6;
pub const VX_VOXELS: usize = 32768;
pub const VX_MAX_QUADS_PER_CHUNK: usize = 98304;
pub const VX_K_COUNT_KINDS: usize = 8;

#[repr(C)]
pub struct vx_ctx {
    _private: [u8; 0],
}

/// perlin_noise.rs:14-18
#[repr(C)]
#[derive(Copy, Clone, Debug)]
pub struct vx_noise_params {
    pub seed: i32,
    pub freq: f32,
    pub gain: f32,
    pub lacunarity: f32,
    pub octaves: i32,
    /// factor on every `simplex_3d` value; 32.69587493801679 (default) or 1.0 (raw sum), see voxb200.h
    pub simplex_scale: f32,
}

/// extension: vertical neighbours + DIRT rule, see voxb200.h
#[repr(C)]
#[derive(Copy, Clone, Debug, Default)]
pub struct vx_terrain_opts {
    pub vertical_neighbours: u32,
    pub dirt_rule: u32,
}

/// renderer/instance.rs:1-6 `InstanceRaw` (68 bytes, `#[repr(C)]`) -- same layout, so a
/// `&[vx_instance]` can be `bytemuck::cast_slice`d straight into `create_buffer_init`.
#[repr(C)]
#[derive(Copy, Clone, Debug)]
pub struct vx_instance {
    pub t_matrix: [[f32; 4]; 4],
    pub texture_index: u32,
}

extern "C" {
    pub fn vx_version() -> c_int;
    pub fn vx_last_error() -> *const c_char;
    pub fn vx_device_count() -> c_int;
    pub fn vx_create(devices: *const c_int, n_devices: c_int, out: *mut *mut vx_ctx) -> c_int;
    pub fn vx_destroy(ctx: *mut vx_ctx);
    pub fn vx_num_devices(ctx: *const vx_ctx) -> c_int;
    pub fn vx_default_noise_params(p: *mut vx_noise_params);
    pub fn vx_launch_count(ctx: *const vx_ctx) -> u64;
    pub fn vx_host_alloc(out: *mut *mut c_void, bytes: usize) -> c_int;
    pub fn vx_host_free(p: *mut c_void) -> c_int;
    pub fn vx_fill(ctx: *mut vx_ctx, coords: *const i32, n: usize, params: *const vx_noise_params,
                   blocks_out: *mut u8, noise_out: *mut f32, occupancy_out: *mut u32,
                   near_zero_out: *mut u64) -> c_int;
    pub fn vx_mesh(ctx: *mut vx_ctx, coords: *const i32, n: usize, blocks: *const u8, nbr_idx: *const i32,
                   mesh_mask: *const u8, counts: *mut u32, offsets: *mut u64, instances: *mut vx_instance,
                   capacity: usize, total_out: *mut u64) -> c_int;
    pub fn vx_gen_mesh(ctx: *mut vx_ctx, coords: *const i32, n: usize, params: *const vx_noise_params,
                       nbr_mask: *const u8, counts: *mut u32, offsets: *mut u64, instances: *mut vx_instance,
                       capacity: usize, total_out: *mut u64, blocks_out: *mut u8) -> c_int;
    pub fn vx_gen_mesh_packed(ctx: *mut vx_ctx, coords: *const i32, n: usize, params: *const vx_noise_params,
                              nbr_mask: *const u8, counts: *mut u32, offsets: *mut u64, quads: *mut u32,
                              capacity: usize, total_out: *mut u64) -> c_int;
    pub fn vx_gen_mesh_packed_device(ctx: *mut vx_ctx, coords: *const i32, n: usize,
                                     params: *const vx_noise_params, nbr_mask: *const u8, counts: *mut u32,
                                     offsets: *mut u64, quads: *mut u32, capacity: usize, status: *mut u32,
                                     stream: *mut c_void) -> c_int;
    pub fn vx_planes_from_matrix(m: *const f32, planes: *mut f32) -> c_int;
    pub fn vx_frustum_cull(ctx: *mut vx_ctx, planes: *const f32, n_frames: usize, coords: *const i32, n: usize,
                           visible: *mut u8) -> c_int;
    pub fn vx_frustum_cull_bits(ctx: *mut vx_ctx, planes: *const f32, n_frames: usize, coords: *const i32,
                                n: usize, bits: *mut u32) -> c_int;
    pub fn vx_fill_device(ctx: *mut vx_ctx, coords: *const i32, n: usize, params: *const vx_noise_params,
                          blocks_out: *mut u8, noise_out: *mut f32, occupancy_out: *mut u32,
                          near_zero_out: *mut u64, stream: *mut c_void) -> c_int;
    pub fn vx_mesh_device(ctx: *mut vx_ctx, coords: *const i32, n: usize, blocks: *const u8,
                          nbr_idx: *const i32, mesh_mask: *const u8, counts: *mut u32, offsets: *mut u64,
                          instances: *mut vx_instance, capacity: usize, status: *mut u32,
                          stream: *mut c_void) -> c_int;
    pub fn vx_gen_mesh_device(ctx: *mut vx_ctx, coords: *const i32, n: usize, params: *const vx_noise_params,
                              nbr_mask: *const u8, counts: *mut u32, offsets: *mut u64,
                              instances: *mut vx_instance, capacity: usize, status: *mut u32,
                              blocks_out: *mut u8, stream: *mut c_void) -> c_int;
    pub fn vx_frustum_cull_device(ctx: *mut vx_ctx, planes: *const f32, n_frames: usize, coords: *const i32,
                                  n: usize, visible: *mut u8, bits: *mut u32, stream: *mut c_void) -> c_int;
    pub fn vx_gen_mesh_ex(ctx: *mut vx_ctx, coords: *const i32, n: usize, params: *const vx_noise_params,
                          opts: *const vx_terrain_opts, nbr_mask: *const u8, counts: *mut u32, offsets: *mut u64,
                          instances: *mut vx_instance, capacity: usize, total_out: *mut u64,
                          blocks_out: *mut u8) -> c_int;
    pub fn vx_gen_mesh_ex_device(ctx: *mut vx_ctx, coords: *const i32, n: usize, params: *const vx_noise_params,
                                 opts: *const vx_terrain_opts, nbr_mask: *const u8, counts: *mut u32,
                                 offsets: *mut u64, instances: *mut vx_instance, capacity: usize,
                                 status: *mut u32, blocks_out: *mut u8, stream: *mut c_void) -> c_int;
    pub fn vx_emit_again_device(ctx: *mut vx_ctx, coords: *const i32, n: usize, counts: *mut u32,
                                offsets: *mut u64, instances: *mut vx_instance, capacity: usize,
                                status: *mut u32, stream: *mut c_void) -> c_int;
    pub fn vx_sync(ctx: *mut vx_ctx) -> c_int;
    pub fn vx_measure_peak(ctx: *mut vx_ctx, kind: c_int, tera_ops: *mut f64) -> c_int;
    pub fn vx_set_timing(ctx: *mut vx_ctx, on: c_int) -> c_int;
    pub fn vx_get_timing(ctx: *mut vx_ctx, ms: *mut f64, launches: *mut u64) -> c_int;
}
