// Host build of the DEVICE mesher core (csrc/vx_mesh_rows.cuh: staged chunk layout, face masks of a row, instance
// tables) for CPU testing.  The staging below fills ChunkSmem the way stage_chunk (kernels_mesh.cu) does -- from block
// ids instead of bit planes -- and the expansion walks rows / voxels / faces in the reference's order (chunk.rs:59-245);
// the face tests and the table words are the device source itself.  Test infrastructure only.
#include <cstdint>
#include <cstring>

#include "../../include/voxb200.h"
#define __device__
#define __forceinline__ inline
#define __constant__ static const
#define __align__(n) alignas(n)
static inline uint32_t __popc(uint32_t v) { return (uint32_t)__builtin_popcount(v); }
struct alignas(16) uint4 { uint32_t x, y, z, w; };
#include "../../wgpu-voxel-engine_b200/csrc/vx_mesh_rows.cuh"

static uint32_t row_word(const uint8_t *blocks, uint32_t y, uint32_t z)      // bit x set <=> voxel (x, y, z) != AIR (cube.rs:16-21)
{
    uint32_t w = 0;
    for (uint32_t x = 0; x < 32; ++x) w |= (blocks[x + 32 * z + 1024 * y] != 2 ? 1u : 0u) << x;
    return w;
}

// blocks: the chunk; nbr[k]: block grid of neighbour k (-x, +x, -z, +z; chunk_builder.rs:134-169) or NULL = None.
// Writes the instances in the reference's order and returns their number; `use4` takes the four-rows-at-a-time form.
extern "C" uint32_t shim_mesh_chunk(const int32_t *coord, const uint8_t *blocks, const uint8_t *const *nbr, int use4,
                                    vx_instance *out)
{
    static ChunkSmem s;
    memset(&s, 0, sizeof s);                                  // y borders: zero rows = always emit (chunk.rs:216-222,234-240)
    for (uint32_t r = 0; r < VX_ROWS; ++r) s.occ[32 + r] = row_word(blocks, r >> 5, r & 31u);
    for (uint32_t y = 0; y < 32; ++y) {
        if (nbr[2]) s.back[y] = row_word(nbr[2], y, 31);      // -z neighbour: its z = 31 rows
        if (nbr[3]) s.front[y] = row_word(nbr[3], y, 0);
        for (uint32_t z = 0; z < 32; ++z) {
            if (nbr[0]) s.xm[y] |= ((row_word(nbr[0], y, z) >> 31) & 1u) << z;      // -x neighbour: its x = 31 voxels
            if (nbr[1]) s.xp[y] |= (row_word(nbr[1], y, z) & 1u) << z;
        }
    }
    uint32_t n = 0;
    for (uint32_t t = 0; t < VX_ROWS / 4; ++t) {
        RowFaces f4[4];
        if (use4) row_faces4(s, t, f4);
        for (uint32_t j = 0; j < 4; ++j) {
            const uint32_t r = 4 * t + j, y = r >> 5, z = r & 31u;
            const RowFaces f = use4 ? f4[j] : row_faces(s, r);
            uint32_t cnt = 0;
            for (uint32_t x = 0; x < 32; ++x)
                for (uint32_t g = 0; g < 6; ++g) {
                    if (!((f.m[g] >> x) & 1u)) continue;
                    const uint32_t type = blocks[x + 32 * z + 1024 * y];      // 0 GRASS, 1 DIRT
                    vx_instance q;
                    uint32_t w[17];
                    for (int k = 0; k < 12; ++k) w[k] = c_rot[g][k];
                    const float px = (float)(int32_t)((uint32_t)coord[0] * 32u + x), py = (float)(int32_t)((uint32_t)coord[1] * 32u + y),
                                pz = (float)(int32_t)((uint32_t)coord[2] * 32u + z), one = 1.0f;
                    memcpy(&w[12], &px, 4); memcpy(&w[13], &py, 4); memcpy(&w[14], &pz, 4); memcpy(&w[15], &one, 4);
                    w[16] = c_tex[type][g];
                    memcpy(&q, w, sizeof q);
                    out[n++] = q;
                    ++cnt;
                }
            if (cnt != row_count(f)) return 0xffffffffu;
        }
    }
    return n;
}
