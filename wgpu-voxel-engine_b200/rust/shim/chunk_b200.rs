//! Drop-in replacements for the hot-path bodies of the reference's `src/chunk.rs`,
//! `src/perlin_noise.rs` and `src/frustum_culling.rs`, forwarding to libvoxb200 through
//! `voxb200-sys`.  Signatures are the reference's (chunk.rs:21, :48-52; frustum_culling.rs:16);
//! see INTEGRATION.md for the three-line patch at each call site.
//! NOT COMPILED HERE (no rustc in the build image).
use crate::coordinate::ChunkCoord3D;
use crate::cube::{Cube, CubeType};
use crate::renderer::instance::InstanceRaw;
use std::ffi::CStr;
use voxb200_sys as ffi;

/// One GPU context per PROCESS, created on first use: the reference's signatures
/// (`Chunk::new(position)`, `create_mesh(&self, device, adjacent_chunks)`, `Frustum::contains(&self, &pos)`,
/// chunk.rs:21,48-52, frustum_culling.rs:16) stay exactly as they are -- nothing has to be threaded
/// through the worker closures of chunk_builder.rs:80-84,105-108,125-129.  libvoxb200 is thread-safe:
/// one-chunk calls from the uvth pool threads lease scratch lanes inside the context and overlap.
pub struct B200 {
    ctx: *mut ffi::vx_ctx,
}
unsafe impl Send for B200 {}
unsafe impl Sync for B200 {}

fn check(rc: i32) {
    if rc != ffi::VX_OK {
        // the reference has no error path here: it panics (chunk_builder.rs:83,107,128)
        let msg = unsafe { CStr::from_ptr(ffi::vx_last_error()) }.to_string_lossy().into_owned();
        panic!("voxb200 error {}: {}", rc, msg);
    }
}

static GPU: std::sync::OnceLock<B200> = std::sync::OnceLock::new();

/// The process-global context (`vx_create(NULL, 0)`: every GPU of the box).  Never destroyed: it lives as
/// long as the process, like `graphics.device`.
pub fn gpu() -> &'static B200 {
    GPU.get_or_init(|| {
        let mut ctx = std::ptr::null_mut();
        check(unsafe { ffi::vx_create(std::ptr::null(), 0, &mut ctx) });
        B200 { ctx }
    })
}

fn cube_from_id(id: u8) -> Cube {
    Cube::new(match id {
        0 => CubeType::GRASS,
        1 => CubeType::DIRT,
        _ => CubeType::AIR,
    })
}

/// Body of `Chunk::generate_terrain` (chunk.rs:27-46): replaces the simdnoise call and the
/// threshold loop.  `cubes` ends up in the same idx order (x + 32 z + 1024 y).
pub fn generate_terrain(cubes: &mut Vec<Cube>, pos: ChunkCoord3D) {
    let gpu = gpu();
    let coords = [pos.x, pos.y, pos.z];
    let mut ids = vec![0u8; ffi::VX_VOXELS];
    check(unsafe {
        ffi::vx_fill(gpu.ctx, coords.as_ptr(), 1, std::ptr::null(), ids.as_mut_ptr(),
                     std::ptr::null_mut(), std::ptr::null_mut(), std::ptr::null_mut())
    });
    cubes.extend(ids.into_iter().map(cube_from_id));
}

/// Body of `Chunk::create_mesh` up to the wgpu upload (chunk.rs:48-57, 254-260): returns the
/// `Vec<InstanceRaw>` that `ChunkMesh::new` hands to `create_buffer_init`.
/// `own` / `adjacent` are the CubeType bytes of the chunk and of its -x,+x,-z,+z neighbours.
pub fn create_mesh_instances(pos: ChunkCoord3D, own: &[u8],
                             adjacent: [Option<(ChunkCoord3D, &[u8])>; 4]) -> Vec<InstanceRaw> {
    let gpu = gpu();
    let mut coords = vec![pos.x, pos.y, pos.z];
    let mut blocks = own.to_vec();
    let mut nbr = [-1i32; 4];
    for (k, a) in adjacent.iter().enumerate() {
        if let Some((p, b)) = a {
            nbr[k] = (coords.len() / 3) as i32;
            coords.extend_from_slice(&[p.x, p.y, p.z]);
            blocks.extend_from_slice(b);
        }
    }
    let n = coords.len() / 3;
    let mut nbr_idx = vec![-1i32; n * 4];
    nbr_idx[..4].copy_from_slice(&nbr);
    let mut mesh_mask = vec![0u8; n];
    mesh_mask[0] = 1;
    let (mut counts, mut offsets, mut total) = (vec![0u32; n], vec![0u64; n + 1], 0u64);
    let mut inst: Vec<InstanceRaw> = Vec::with_capacity(ffi::VX_MAX_QUADS_PER_CHUNK);
    check(unsafe {
        ffi::vx_mesh(gpu.ctx, coords.as_ptr(), n, blocks.as_ptr(), nbr_idx.as_ptr(), mesh_mask.as_ptr(),
                     counts.as_mut_ptr(), offsets.as_mut_ptr(), inst.as_mut_ptr() as *mut ffi::vx_instance,
                     ffi::VX_MAX_QUADS_PER_CHUNK, &mut total)
    });
    unsafe { inst.set_len(total as usize) }; // InstanceRaw and vx_instance share one repr(C) layout
    inst
}

/// Pinned (page-locked) host memory from `vx_host_alloc`.  A download into it runs at link speed (~50 GB/s); a
/// download into a fresh `Vec` is bound by the first-touch page faults of the destination (~10 GB/s even with the
/// library's bounce buffers and copy threads).  One arena per thread, grown on demand.
struct PinnedArena {
    ptr: *mut u8,
    bytes: usize,
}
impl PinnedArena {
    const fn new() -> Self {
        PinnedArena { ptr: std::ptr::null_mut(), bytes: 0 }
    }
    fn ensure(&mut self, bytes: usize) -> *mut u8 {
        if bytes > self.bytes {
            if !self.ptr.is_null() {
                check(unsafe { ffi::vx_host_free(self.ptr as *mut std::ffi::c_void) });
                self.ptr = std::ptr::null_mut();
                self.bytes = 0;
            }
            let want = bytes + bytes / 4;
            let mut p: *mut std::ffi::c_void = std::ptr::null_mut();
            check(unsafe { ffi::vx_host_alloc(&mut p, want) });
            self.ptr = p as *mut u8;
            self.bytes = want;
        }
        self.ptr
    }
}
thread_local! {
    static ARENA: std::cell::RefCell<PinnedArena> = std::cell::RefCell::new(PinnedArena::new());
}

/// The fast path: the worker closure of chunk_builder.rs:80-84 for a whole coordinate list in
/// one call (noise + meshing on the GPU, instances and cube ids straight into this thread's pinned arena).
/// Returns per-chunk (cube ids, instances).  `nbr_mask[i]` bit k = neighbour k counts as loaded.
/// ONE pass: the instance buffer is sized from the terrain's typical density and the call is repeated
/// only if the library answers VX_ERR_CAPACITY (counts, offsets and the exact total are valid then).
pub fn build_chunks(positions: &[ChunkCoord3D], nbr_mask: Option<&[u8]>) -> Vec<(Vec<u8>, Vec<InstanceRaw>)> {
    let gpu = gpu();
    let n = positions.len();
    let coords: Vec<i32> = positions.iter().flat_map(|p| vec![p.x, p.y, p.z]).collect();
    let mask_ptr = nbr_mask.map_or(std::ptr::null(), |m| m.as_ptr());
    let (mut counts, mut offsets, mut total) = (vec![0u32; n], vec![0u64; n + 1], 0u64);
    ARENA.with(|arena| {
        let mut arena = arena.borrow_mut();
        // default terrain: ~4.6 k quads per chunk with neighbours, at most ~7.9 k without
        let mut cap = n * 8192;
        loop {
            let inst_bytes = cap * std::mem::size_of::<InstanceRaw>();
            let base = arena.ensure(inst_bytes + n * ffi::VX_VOXELS);        // [instances][cube ids]
            let blk_ptr = unsafe { base.add(inst_bytes) };
            let rc = unsafe {
                ffi::vx_gen_mesh(gpu.ctx, coords.as_ptr(), n, std::ptr::null(), mask_ptr, counts.as_mut_ptr(),
                                 offsets.as_mut_ptr(), base as *mut ffi::vx_instance, cap, &mut total, blk_ptr)
            };
            if rc == ffi::VX_ERR_CAPACITY {
                cap = total as usize; // exact now
                continue;
            }
            check(rc);
            // the arena is page aligned and InstanceRaw is 17 x 4 bytes, align 4 (renderer/instance.rs:1-6)
            let inst = unsafe { std::slice::from_raw_parts(base as *const InstanceRaw, total as usize) };
            let blocks = unsafe { std::slice::from_raw_parts(blk_ptr as *const u8, n * ffi::VX_VOXELS) };
            return (0..n).map(|i| {
                (blocks[i * ffi::VX_VOXELS..(i + 1) * ffi::VX_VOXELS].to_vec(),
                 inst[offsets[i] as usize..offsets[i + 1] as usize].to_vec())
            }).collect();
        }
    })
}

/// `Frustum::matrix_to_planes` (frustum_culling.rs:47-85): `m` = `matrix.data.0` flattened.
pub fn matrix_to_planes(m: &[f32; 16]) -> [f32; 24] {
    let mut planes = [0f32; 24];
    check(unsafe { ffi::vx_planes_from_matrix(m.as_ptr(), planes.as_mut_ptr()) });
    planes
}

/// `Frustum::contains` (frustum_culling.rs:16-40) for a whole list in one launch; use it at
/// chunk_builder.rs:217-255 (enqueue gate) and world.rs:34-37 (draw gate).
pub fn frustum_contains_many(planes: &[f32; 24], positions: &[ChunkCoord3D]) -> Vec<bool> {
    let gpu = gpu();
    let coords: Vec<i32> = positions.iter().flat_map(|p| vec![p.x, p.y, p.z]).collect();
    let mut vis = vec![0u8; positions.len()];
    check(unsafe {
        ffi::vx_frustum_cull(gpu.ctx, planes.as_ptr(), 1, coords.as_ptr(), positions.len(), vis.as_mut_ptr())
    });
    vis.into_iter().map(|v| v != 0).collect()
}
