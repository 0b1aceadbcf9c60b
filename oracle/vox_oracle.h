/*
 * vox_oracle.h -- CPU oracle for the chunk-production hot path of
 * Blatko1/wgpu-voxel-engine.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this library.
 * The product (libvoxb200.so) never links, loads or calls anything in here.
 *
 * PARITY STATUS
 *   - threshold, mesher, texture slots, instance packing, plane extraction and
 *     the 8-corner frustum test are restated from reference SOURCE and are
 *     deterministic integer / IEEE-754 logic: pinned by source.
 *   - the noise arithmetic lives in the un-vendored crate simdnoise 3.1.6
 *     (Cargo.lock:1802-1808) on simdeez 1.0.7 (Cargo.lock:1792-1799).  It is
 *     restated from the crate's published algorithm.  The
 *     reference holds no tests / golden vectors and cannot be built here, so
 *     NOISE PARITY IS UNPINNED against the real crate; it is pinned only
 *     against an independent numpy restatement (tests/golden/make_golden.py)
 *     and the SURVEY App. A self-check vectors.
 */
#ifndef VOX_ORACLE_H
#define VOX_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VXO_CHUNK 32            /* world.rs:21 CHUNK_USIZE */
#define VXO_VOXELS 32768        /* 32^3 */

/* cube.rs:16-21 */
enum { VXO_GRASS = 0, VXO_DIRT = 1, VXO_AIR = 2 };

/* Face ids in the reference's per-cube emission order, chunk.rs:77-240 */
enum { VXO_LEFT = 0, VXO_RIGHT = 1, VXO_BACK = 2, VXO_FRONT = 3, VXO_DOWN = 4, VXO_UP = 5 };

/* perlin_noise.rs:14-18 */
typedef struct {
    int32_t seed;       /* 1 */
    float freq;         /* 0.09 */
    float gain;         /* 2.5 */
    float lacunarity;   /* 0.5 */
    int32_t octaves;    /* 3 */
    /* Factor applied to every simplex_3d value (one f32 multiply per octave sample) before the FBM
     * combination.  simdnoise's later 3.1.x sources normalise simplex_3d to [-1, 1] with
     * 32.69587493801679 ("scaling factor found by numerical optimization"); the FastNoiseSIMD-lineage
     * sources it started from return the raw sum (scale 1).  Which of the two 3.1.6 ships cannot be
     * checked here (crate not vendored): the default is the normalised form, 1.0f selects the raw
     * one, and tools/simdnoise_golden/ is the recipe that settles it on a machine with cargo. */
    float simplex_scale; /* 32.69587493801679 */
} vxo_noise_params;

/* quad.rs:7-11 before to_raw(): integer world position, face, texture layer */
typedef struct {
    int32_t x, y, z;
    uint32_t face;
    uint32_t tex;
} vxo_quad;

/* renderer/instance.rs:1-6, #[repr(C)], 68 bytes */
typedef struct {
    float t_matrix[4][4];   /* column arrays: t_matrix[c][r] */
    uint32_t texture_index;
} vxo_instance_raw;

void vxo_default_params(vxo_noise_params *p);

/* simdnoise scalar restatement; exposed for unit tests */
float vxo_simplex_3d(float x, float y, float z, int32_t seed);           /* raw sum, before the scale */
float vxo_fbm_3d(float x, float y, float z, float lac, float gain, int32_t octaves, int32_t seed,
                 float simplex_scale);

/* perlin_noise.rs:5-22 -> simdnoise get_3d_noise; out[x + 32*z + 1024*y] */
void vxo_noise_chunk(const int32_t coord[3], const vxo_noise_params *p, float *out);
/* same values, 8-wide AVX2 lanes like simdnoise::avx2 (falls back to scalar) */
void vxo_noise_chunk_simd(const int32_t coord[3], const vxo_noise_params *p, float *out);
int vxo_have_avx2(void);

/* chunk.rs:27-46 */
void vxo_terrain_from_noise(const float *noise, uint8_t *cubes);
void vxo_generate_terrain(const int32_t coord[3], const vxo_noise_params *p, uint8_t *cubes);

/* chunk.rs:59-245; adj[k] (k: 0=-x 1=+x 2=-z 3=+z, chunk_builder.rs:134-169)
 * may be NULL (= None).  Returns the number of quads; writes at most cap. */
size_t vxo_cull_unseen_triangles(const int32_t coord[3], const uint8_t *cubes,
                                 const uint8_t *const adj[4], vxo_quad *out, size_t cap);

/* quad.rs:14-54 (per-quad sin/cos + 4x4 product, like the reference) */
void vxo_quad_to_raw(const vxo_quad *q, vxo_instance_raw *out);
/* the six rotations, R[face][row][col], evaluated like quad.rs:16-40 */
void vxo_face_rotation(uint32_t face, float R[3][3]);

/* chunk.rs:48-57 + ChunkMesh::new chunk.rs:254-266 minus the wgpu upload */
size_t vxo_create_mesh(const int32_t coord[3], const uint8_t *cubes,
                       const uint8_t *const adj[4], vxo_instance_raw *out, size_t cap);

/* frustum_culling.rs:47-85; m = column-major 4x4 (m[c*4+r]); planes[6][4] in
 * the stored order near, far, left, right, top, bottom (:84) */
void vxo_matrix_to_planes(const float m[16], float planes[24]);
/* frustum_culling.rs:16-40 */
int vxo_frustum_contains(const float planes[24], const int32_t coord[3]);
void vxo_frustum_cull(const float *planes, size_t n_frames, const int32_t *coords, size_t n,
                      uint8_t *out /* [n_frames][n] */);

/*
 * Multithreaded work function = the reference's worker closure
 * (chunk_builder.rs:80-84) over a coordinate list: Chunk::new for every chunk,
 * then create_mesh with the neighbours found in the same list and allowed by
 * nbr_mask (bit k = neighbour k present; NULL = every in-list neighbour).
 * counts[n] receives quads per chunk.  If inst != NULL the instances are also
 * stored contiguously in list order (capacity cap quads, offsets[n+1] out).
 * Returns total quads, or (size_t)-1 on capacity overflow / alloc failure.
 * use_simd: 1 = AVX2 lanes for the noise like simdnoise::avx2.
 */
size_t vxo_gen_mesh_batch(const int32_t *coords, size_t n, const vxo_noise_params *p,
                          const uint8_t *nbr_mask, int nthreads, int use_simd,
                          uint32_t *counts, uint64_t *offsets,
                          vxo_instance_raw *inst, size_t cap, uint8_t *cubes_out);

/* ---- EXTENSION -- not a restatement: the reference never consults vertical neighbours
 * (chunk.rs:205-240) and never generates DIRT (cube.rs:19).  These define the extension's semantics by analogy
 * with the reference's own x / z border rule so that the CUDA path has something to be checked against. */
size_t vxo_cull_unseen_triangles_ex(const int32_t coord[3], const uint8_t *cubes,
                                    const uint8_t *const adj[6] /* -x,+x,-z,+z,-y,+y */, vxo_quad *out, size_t cap);
size_t vxo_create_mesh_ex(const int32_t coord[3], const uint8_t *cubes,
                          const uint8_t *const adj[6], vxo_instance_raw *out, size_t cap);
void vxo_dirt_rule(uint8_t *cubes, const uint8_t *above /* the +y chunk, or NULL */);
size_t vxo_gen_mesh_batch_ex(const int32_t *coords, size_t n, const vxo_noise_params *p,
                             const uint8_t *nbr_mask /* bits 0-5 */, int vertical, int dirt_rule,
                             int nthreads, int use_simd, uint32_t *counts, uint64_t *offsets,
                             vxo_instance_raw *inst, size_t cap, uint8_t *cubes_out);

/* mesher-only variant over caller-provided block grids (config 4) */
size_t vxo_mesh_batch(const int32_t *coords, size_t n, const uint8_t *cubes,
                      const int32_t *nbr_idx /* [n][4], -1 = None */, int nthreads,
                      uint32_t *counts, uint64_t *offsets,
                      vxo_instance_raw *inst, size_t cap);

#ifdef __cplusplus
}
#endif
#endif
