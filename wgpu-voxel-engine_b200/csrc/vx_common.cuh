// vx_common.cuh -- shared device/host declarations for libvoxb200 (sm_100a).
#pragma once
#include <atomic>
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/voxb200.h"

#define VX_FULL 0xffffffffu

// kernel-launch counter of a device / lane: bumped by whichever host thread launches, read by vx_launch_count
typedef std::atomic<uint64_t> vx_counter;

// status word bits written by device code (vx_*_device `status`)
#define VX_ST_OVERFLOW 1u
#define VX_ST_BAD_BLOCK 2u
#define VX_ST_BAD_NBR 4u        /* a caller-provided nbr_idx entry >= n (treated as None) */

// Neighbour reference encoding used between the resolve and mesh kernels:
//   >= 0  : index of a full chunk in the occupancy array
//   -1    : None (border faces are emitted, chunk.rs:102-108 ...)
//   <= -2 : halo slab number -(v + 2) in the halo array (32 words per slab)
#define VX_NBR_NONE (-1)
__host__ __device__ inline int vx_halo_ref(int slot) { return -(slot + 2); }
__host__ __device__ inline int vx_halo_slot(int ref) { return -ref - 2; }

struct vx_halo_item {      // one 32x32 border slab of a chunk that is not in the batch
    int32_t cx, cy, cz;
    int32_t kind;          // 0: x = 31 slab (serves a -x lookup), 1: x = 0, 2: z = 31, 3: z = 0;
                           // extension (vertical neighbours): 4: y = 31 plane (serves a -y lookup), 5: y = 0
};

// ---- launchers implemented in the kernel translation units ----------------
struct vx_fill_args {
    const int32_t *coords;            // [n][3]           (full-chunk mode)
    const vx_halo_item *halo_items;   // [*halo_count]    (halo mode)
    const uint32_t *halo_count;       // device counter   (halo mode)
    uint32_t max_units;               // grid bound: n*32 or max halo items
    vx_noise_params p;
    uint32_t *occ;                    // [n][1024] or halo [slots][32]
    uint8_t *blocks;                  // optional [n][32768]
    float *noise;                     // optional [n][32768]
    unsigned long long *near_zero;    // optional counter
    float thr1, thr2;                 // sign-only shortcuts (set by the launcher)
    float err_margin;                 // certified |fused - exact| bound (set by the launcher)
};
cudaError_t vx_launch_fill(const vx_fill_args &a, bool halo_mode, cudaStream_t s, vx_counter *launches);

struct vx_mesh_args {
    const uint32_t *occ;        // [items][1024]
    const uint32_t *typ;        // [items][1024] DIRT plane or nullptr (all GRASS)
    const uint32_t *halo;       // [slots][32]
    const int32_t *coords;      // [n][3]
    const int32_t *nbr;         // [n][4] encoded neighbour refs
    const int32_t *nbr_y;       // extension: [n][2] refs of the -y / +y neighbours, nullptr = reference behaviour
                                // (y-border faces always emitted, chunk.rs:216-222,234-240)
    int dirt_rule;              // extension: a solid voxel with a solid voxel above it is DIRT (typ == nullptr only)
    const uint8_t *mesh_mask;   // [n] or nullptr
    uint32_t *counts;           // [n]
    uint64_t *offsets;          // [n+1]
    vx_instance *out;           // (uint32_t * when packed)
    int packed;                 // emit 32-bit quad records instead of 68-byte instances
    unsigned long long capacity;
    uint32_t *status;           // optional
    uint32_t *ticket;           // count kernel: zero-initialised CTA ticket (last CTA runs the scan)
    uint32_t n;
    uint32_t m0;                // emit: first chunk of this launch
    unsigned long long quad_base;   // emit: offsets value that maps to out[0]
};
cudaError_t vx_launch_mesh_count(const vx_mesh_args &a, cudaStream_t s, vx_counter *launches);
cudaError_t vx_launch_mesh_emit(const vx_mesh_args &a, uint32_t len, cudaStream_t s, vx_counter *launches);
// extension: CubeType bytes from the occupancy planes (+ the DIRT rule); needs the mesh args' nbr_y / halo
cudaError_t vx_launch_blocks_ex(const vx_mesh_args &a, uint8_t *blocks, cudaStream_t s, vx_counter *launches);
cudaError_t vx_launch_pack(const uint8_t *blocks, uint32_t n, uint32_t *occ, uint32_t *typ,
                           uint32_t *status, cudaStream_t s, vx_counter *launches);

struct vx_resolve_args {
    const int32_t *coords;      // [n][3]
    const uint8_t *nbr_mask;    // [n] or nullptr
    uint32_t n;                 // whole list (hashed)
    uint32_t lo, hi;            // this device meshes chunks [lo, hi); refs are relative to lo
    int32_t *table;             // [table_size] hash slots (-1 = empty)
    uint32_t table_size;        // power of two
    int32_t *nbr;               // out [hi-lo][4]
    int32_t *nbr_y;             // out [hi-lo][2] (-y, +y; mask bits 4, 5) or nullptr
    vx_halo_item *halo_items;   // out [4(hi-lo)]  (nullptr: neighbours outside are None)
    uint32_t *halo_count;       // out counter
};
cudaError_t vx_launch_resolve(const vx_resolve_args &a, cudaStream_t s, vx_counter *launches);
cudaError_t vx_launch_remap_nbr(const int32_t *nbr_in, int32_t *nbr_out, uint32_t n, uint32_t *status, cudaStream_t s,
                                vx_counter *launches);

cudaError_t vx_launch_frustum(const float *planes, uint32_t n_frames, const int32_t *coords, uint32_t n,
                              uint8_t *visible, uint32_t *bits, cudaStream_t s, vx_counter *launches);

cudaError_t vx_init_mesh_kernels();

cudaError_t vx_run_peak(int kind, int n_sm, double *tera_ops, cudaStream_t s, vx_counter *launches);
