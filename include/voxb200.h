/*
 * voxb200.h -- C ABI of libvoxb200.so: the B200-native chunk-production path
 * for Blatko1/wgpu-voxel-engine (noise terrain fill -> hidden-face culling /
 * instance emission -> per-chunk frustum AABB test).
 *
 * This is the drop-in boundary.  Each entry point names the reference
 * interface it replaces (paths relative to the reference's src/).  The Rust
 * FFI crate that binds these symbols is in wgpu-voxel-engine_b200/rust/ and is
 * walked through in INTEGRATION.md.
 *
 * Conventions
 *   - every call returns an int status: 0 = VX_OK, negative = error; the text
 *     of the last error on the calling thread is vx_last_error().  The
 *     reference has no error path here (it panics, chunk_builder.rs:83,107,128);
 *     the Rust shim panics on non-zero.
 *   - all outputs are caller-allocated; the library never frees caller memory.
 *   - there is NO CPU fallback: without a CUDA device every compute call fails
 *     with VX_ERR_NO_DEVICE.
 *   - every entry point is thread-safe.  Batch calls (more than 32 chunks) and all *_device calls on one
 *     context are serialised by the context lock; short vx_fill / vx_mesh / vx_gen_mesh calls (<= 32
 *     chunks, the reference's one-chunk-per-call shape) instead lease one of up to 8 scratch lanes (own
 *     buffers and streams on the context's first GPU) and run concurrently -- a ninth concurrent caller
 *     waits for a lane.  vx_sync covers the context's own streams, not the lanes (a lane call has
 *     finished when it returns).  Use one context per GPU, or one multi-GPU context.
 *   - host output buffers may be ordinary pageable memory (a Vec, a numpy array) or pinned memory from
 *     vx_host_alloc.  Pinned buffers are written at link speed (~50 GB/s); large downloads into pageable
 *     memory go through the library's own pinned bounce buffers and a few copy threads (default: half
 *     the hardware threads, at most 8; environment variable VX_COPY_THREADS overrides), 10-25 GB/s
 *     depending on whether the pages have been touched before.
 *   - chunk edge is fixed at 32 (world.rs:21); voxel index inside a chunk is
 *     idx = x + 32*z + 1024*y (chunk.rs:38,70); chunk -> world is a wrapping
 *     i32 multiply by 32 (coordinate.rs:60-65, release-build semantics).
 */
#ifndef VOXB200_H
#define VOXB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VX_VERSION 110          /* 0.1.1: vx_noise_params.simplex_scale */
#define VX_CHUNK 32
#define VX_VOXELS 32768
#define VX_ROWS 1024            /* occupancy words per chunk: row = y*32+z, bit = x */
#define VX_MAX_QUADS_PER_CHUNK 98304
#define VX_MAX_OCTAVES 8

enum {
    VX_OK = 0,
    VX_ERR_BAD_ARG = -1,
    VX_ERR_CAPACITY = -2,       /* instance buffer too small; counts/offsets/total are valid */
    VX_ERR_CUDA = -3,
    VX_ERR_NO_DEVICE = -4,
    VX_ERR_OOM = -5,
    VX_ERR_BAD_BLOCK = -6       /* a block id outside {0,1,2} (cube.rs:16-21) */
};

/* cube.rs:16-21 CubeType */
enum { VX_GRASS = 0, VX_DIRT = 1, VX_AIR = 2 };
/* per-cube emission order of chunk.rs:77-240 */
enum { VX_LEFT = 0, VX_RIGHT = 1, VX_BACK = 2, VX_FRONT = 3, VX_DOWN = 4, VX_UP = 5 };

/* perlin_noise.rs:14-18 -- the reference hard-codes these; the defaults
 * (vx_default_noise_params) reproduce them. */
typedef struct vx_noise_params {
    int32_t seed;        /* 1    */
    float freq;          /* 0.09 */
    float gain;          /* 2.5  */
    float lacunarity;    /* 0.5  */
    int32_t octaves;     /* 3 (1..VX_MAX_OCTAVES) */
    /* Factor on every simplex_3d value (one f32 multiply per octave sample, before the FBM combination).
     * simdnoise normalises simplex_3d to [-1, 1] with 32.69587493801679 in its later 3.1.x sources; the
     * FastNoiseSIMD-lineage code it grew from returns the raw sum (= 1.0f here, bit for bit).  The crate
     * is not vendored with the reference (Cargo.lock:1802-1808), so which of the two 3.1.6 ships is
     * settled by tools/simdnoise_golden/ on a machine with cargo; both are parity-tested against the
     * oracle.  0 selects the default. */
    float simplex_scale; /* 32.69587493801679 */
} vx_noise_params;
#define VX_SIMPLEX_SCALE_DEFAULT 32.69587493801679f
#define VX_SIMPLEX_SCALE_RAW 1.0f

/* renderer/instance.rs:1-6 InstanceRaw, #[repr(C)]: 16 f32 + u32 = 68 bytes.
 * t_matrix[c] is column c of Translation(pos) * Rotation(face) (quad.rs:47-54). */
typedef struct vx_instance {
    float t_matrix[4][4];
    uint32_t texture_index;
} vx_instance;

typedef struct vx_ctx vx_ctx;

/* ---- library / context -------------------------------------------------- */
int vx_version(void);
const char *vx_last_error(void);
/* number of CUDA devices, or a negative status */
int vx_device_count(void);
/* devices == NULL: use n_devices devices starting at 0 (n_devices <= 0: all).
 * With several devices every batch call shards its list by contiguous index
 * ranges, one stream per GPU, no inter-GPU traffic (halo slabs are recomputed).
 * A device may be named more than once: every entry is a slot with its own streams and scratch. */
int vx_create(const int *devices, int n_devices, vx_ctx **out);
void vx_destroy(vx_ctx *ctx);
int vx_num_devices(const vx_ctx *ctx);
void vx_default_noise_params(vx_noise_params *p);
/* kernels launched by this context so far (bench.py's gpu_launches) */
uint64_t vx_launch_count(const vx_ctx *ctx);
/* pinned host memory for the host-buffer entry points (optional but faster) */
int vx_host_alloc(void **out, size_t bytes);
int vx_host_free(void *p);

/* ---- terrain fill -------------------------------------------------------
 * Replaces Chunk::new / Chunk::generate_terrain (chunk.rs:21-46) and
 * perlin_noise::perlin_3d_block_{avx2,sse41} (perlin_noise.rs:5-41).
 * coords: int32[n][3] chunk coordinates (ChunkCoord3D, coordinate.rs:41-46).
 * Any output may be NULL:
 *   blocks_out     uint8 [n][32768]  CubeType per voxel (Vec<Cube>, chunk.rs:17)
 *   noise_out      float [n][32768]  the f32 noise field (perlin_noise.rs:20)
 *   occupancy_out  uint32[n][1024]   bit x of word y*32+z set <=> voxel != AIR
 *   near_zero_out  uint64[1]         voxels with |noise| <= 1e-6 (tolerance census) */
int vx_fill(vx_ctx *ctx, const int32_t *coords, size_t n, const vx_noise_params *params,
            uint8_t *blocks_out, float *noise_out, uint32_t *occupancy_out,
            uint64_t *near_zero_out);

/* ---- meshing of caller-provided block grids -----------------------------
 * Replaces Chunk::create_mesh / cull_unseen_triangles / ChunkMesh::new
 * (chunk.rs:48-57, 59-245, 254-266) minus the wgpu upload.
 * blocks: uint8[n][32768].  nbr_idx: int32[n][4] = index into this list of the
 * -x,+x,-z,+z neighbour (order of chunk_builder.rs:134-169) or -1 for None;
 * NULL = look neighbours up by coordinate inside the list.  mesh_mask: uint8[n],
 * 0 = the chunk only serves as a neighbour (count 0); NULL = mesh all.
 * Outputs: counts uint32[n], offsets uint64[n+1] (exclusive scan of counts),
 * instances[capacity] in list order, each chunk's quads in the reference's
 * emission order ((y,z,x)-major, then LEFT,RIGHT,BACK,FRONT,DOWN,UP).
 * instances may be NULL (count only).  total_out receives offsets[n].
 * An nbr_idx entry >= n is VX_ERR_BAD_ARG (checked before anything is enqueued).
 * With several GPUs in the context, lists of more than 32 chunks are split by contiguous index range; a
 * device also receives the out-of-range chunks its own chunks name as neighbours (mesh-less ghosts). */
int vx_mesh(vx_ctx *ctx, const int32_t *coords, size_t n, const uint8_t *blocks,
            const int32_t *nbr_idx, const uint8_t *mesh_mask,
            uint32_t *counts, uint64_t *offsets,
            vx_instance *instances, size_t capacity, uint64_t *total_out);

/* ---- fused generate + mesh ----------------------------------------------
 * The worker closure of chunk_builder.rs:80-84 (Chunk::new then create_mesh)
 * over a coordinate list.  nbr_mask: uint8[n], bit k set = neighbour k is
 * present (its blocks come from the same noise, wherever it is); NULL =
 * present iff the neighbour coordinate is in the list.  blocks_out optional.
 * Lists of up to 32 chunks (the reference's one-chunk-per-call shape) take a one-synchronisation path
 * and run concurrently when several threads call on the same context (up to 8 at a time);
 * if `instances` is pinned host memory (vx_host_alloc) the kernels write it directly.  On
 * VX_ERR_CAPACITY nothing is written to `instances`; counts, offsets and total_out are valid. */
int vx_gen_mesh(vx_ctx *ctx, const int32_t *coords, size_t n, const vx_noise_params *params,
                const uint8_t *nbr_mask,
                uint32_t *counts, uint64_t *offsets,
                vx_instance *instances, size_t capacity, uint64_t *total_out,
                uint8_t *blocks_out);

/* ---- extension: packed quads -------------------------------
 * NOT the reference's wire format.  Same as vx_gen_mesh, but each quad is one
 * 32-bit record instead of a 68-byte InstanceRaw, for renderers that expand in
 * the vertex shader (17x less PCIe traffic):
 *   bits 0-4 x | 5-9 z | 10-14 y (voxel idx inside the chunk) | 15-17 face
 *   (VX_LEFT..VX_UP) | 18 block type (0 GRASS, 1 DIRT).
 * Quads of chunk i are quads[offsets[i] .. offsets[i+1]), in the same order as
 * the instances; InstanceRaw = translate(chunk*32 + (x,y,z)) * rotation(face),
 * texture = TEXTURE_INDEX_LIST[type][slot(face)]. */
int vx_gen_mesh_packed(vx_ctx *ctx, const int32_t *coords, size_t n, const vx_noise_params *params,
                       const uint8_t *nbr_mask, uint32_t *counts, uint64_t *offsets,
                       uint32_t *quads, size_t capacity, uint64_t *total_out);
int vx_gen_mesh_packed_device(vx_ctx *ctx, const int32_t *coords, size_t n, const vx_noise_params *params,
                              const uint8_t *nbr_mask, uint32_t *counts, uint64_t *offsets,
                              uint32_t *quads, size_t capacity, uint32_t *status, void *stream);

/* ---- extension: vertical neighbours and a DIRT rule -----------
 * NOT reference behaviour: the reference never consults a chunk's -y / +y neighbours (every face on a
 * y border is emitted, chunk.rs:205-240) and never generates DIRT (cube.rs:19, texture.rs:242).
 *   vertical_neighbours  nbr_mask bits 4 (-y) and 5 (+y) are honoured: a present vertical neighbour hides
 *                        the border faces it covers, exactly like the x / z rule of chunk.rs:89-108; an
 *                        absent one leaves them emitted.  With nbr_mask == NULL a vertical neighbour is
 *                        present iff it is in the list.
 *   dirt_rule            a solid voxel whose +y neighbour voxel is solid is DIRT (texture layer 1 on all six
 *                        faces, texture.rs:242), else GRASS; the voxel above y = 31 is looked up in the +y
 *                        chunk when vertical_neighbours says it is present, else it counts as air.
 * opts == NULL, or both fields 0, is vx_gen_mesh.  blocks_out then carries DIRT (1) as well. */
typedef struct vx_terrain_opts {
    uint32_t vertical_neighbours;
    uint32_t dirt_rule;
} vx_terrain_opts;
int vx_gen_mesh_ex(vx_ctx *ctx, const int32_t *coords, size_t n, const vx_noise_params *params,
                   const vx_terrain_opts *opts, const uint8_t *nbr_mask,
                   uint32_t *counts, uint64_t *offsets, vx_instance *instances, size_t capacity,
                   uint64_t *total_out, uint8_t *blocks_out);
int vx_gen_mesh_ex_device(vx_ctx *ctx, const int32_t *coords, size_t n, const vx_noise_params *params,
                          const vx_terrain_opts *opts, const uint8_t *nbr_mask,
                          uint32_t *counts, uint64_t *offsets, vx_instance *instances, size_t capacity,
                          uint32_t *status, uint8_t *blocks_out, void *stream);

/* ---- frustum -------------------------------------------------------------
 * vx_planes_from_matrix replaces Frustum::matrix_to_planes
 * (frustum_culling.rs:47-85): m = column-major Matrix4<f32> (m[c*4+r]),
 * planes[6][4] in the stored order near, far, left, right, top, bottom.
 * vx_frustum_cull replaces Frustum::contains (frustum_culling.rs:16-40) for
 * n chunks x n_frames plane sets: visible[f*n + i] = 1 iff ANY of the 8 AABB
 * corners of chunk i is inside ALL six planes of frame f. */
int vx_planes_from_matrix(const float m[16], float planes[24]);
int vx_frustum_cull(vx_ctx *ctx, const float *planes, size_t n_frames,
                    const int32_t *coords, size_t n, uint8_t *visible);
/* bit-packed result: bits[f*ceil(n/32) + i/32] bit (i%32) */
int vx_frustum_cull_bits(vx_ctx *ctx, const float *planes, size_t n_frames,
                         const int32_t *coords, size_t n, uint32_t *bits);

/* ---- device-resident variants --------------------------------------------
 * Same semantics; every pointer is DEVICE memory on the context's first GPU,
 * work is enqueued on `stream` (a cudaStream_t; NULL = the context's own
 * stream -- pass cudaStreamLegacy / cudaStreamPerThread to name a default
 * stream explicitly) and the call returns without synchronising.  Capacity overflow cannot be
 * reported synchronously: status[0] (device uint32, optional) gets bit 0 set,
 * bit 1 = bad block id, bit 2 = an nbr_idx entry >= n (treated as None).  Scratch lives in the context and is valid until the
 * next call on it. */
int vx_fill_device(vx_ctx *ctx, const int32_t *coords, size_t n, const vx_noise_params *params,
                   uint8_t *blocks_out, float *noise_out, uint32_t *occupancy_out,
                   uint64_t *near_zero_out, void *stream);
int vx_mesh_device(vx_ctx *ctx, const int32_t *coords, size_t n, const uint8_t *blocks,
                   const int32_t *nbr_idx, const uint8_t *mesh_mask,
                   uint32_t *counts, uint64_t *offsets,
                   vx_instance *instances, size_t capacity, uint32_t *status, void *stream);
int vx_gen_mesh_device(vx_ctx *ctx, const int32_t *coords, size_t n, const vx_noise_params *params,
                       const uint8_t *nbr_mask,
                       uint32_t *counts, uint64_t *offsets,
                       vx_instance *instances, size_t capacity, uint32_t *status,
                       uint8_t *blocks_out, void *stream);
int vx_frustum_cull_device(vx_ctx *ctx, const float *planes, size_t n_frames,
                           const int32_t *coords, size_t n, uint8_t *visible, uint32_t *bits,
                           void *stream);
/* Emits the instances of the LAST vx_gen_mesh_device / vx_mesh_device batch of this context again (its
 * occupancy planes and neighbour table are still in the context's scratch; counts / offsets are the
 * arrays that call filled).  For a renderer that lost or re-sized its instance buffer: ChunkMesh::new
 * (chunk.rs:254-266) without a new Chunk::new.  bench.py also uses it to time the emit kernel over
 * back-to-back launches. */
int vx_emit_again_device(vx_ctx *ctx, const int32_t *coords, size_t n, uint32_t *counts, uint64_t *offsets,
                         vx_instance *instances, size_t capacity, uint32_t *status, void *stream);
/* block until everything enqueued by this context has finished */
int vx_sync(vx_ctx *ctx);

/* ---- measurement helpers (bench.py) ---------------------------------------
 * Measured non-tensor issue peaks of the current device, in Tera-ops/s
 * (FMA counts 2): kind 0 = FP32 FFMA, 1 = FP32 FMUL+FADD (non-fused),
 * 2 = INT32 IMAD+LOP3 mix. */
int vx_measure_peak(vx_ctx *ctx, int kind, double *tera_ops);

/* Per-kernel-class device time, measured with CUDA events recorded on the
 * launching stream around every launch while timing is on.  vx_get_timing
 * waits for the recorded work, returns the summed milliseconds and launch
 * counts per class since the last call, and resets them. */
enum {
    VX_K_FILL = 0,      /* noise fill of full chunks            */
    VX_K_HALO = 1,      /* noise fill of halo slabs             */
    VX_K_RESOLVE = 2,   /* neighbour hash insert + lookup       */
    VX_K_COUNT = 3,     /* face count per chunk + exclusive scan (run by its last CTA) */
    VX_K_SCAN = 4,      /* (unused: the scan is fused into VX_K_COUNT) */
    VX_K_EMIT = 5,      /* instance emission                    */
    VX_K_PACK = 6,      /* block ids -> bit planes              */
    VX_K_FRUSTUM = 7,   /* AABB x plane test                    */
    VX_K_COUNT_KINDS = 8
};
int vx_set_timing(vx_ctx *ctx, int on);
int vx_get_timing(vx_ctx *ctx, double ms[VX_K_COUNT_KINDS], uint64_t launches[VX_K_COUNT_KINDS]);

#ifdef __cplusplus
}
#endif
#endif
