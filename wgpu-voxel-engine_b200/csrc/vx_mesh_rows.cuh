// vx_mesh_rows.cuh -- device-side core of the mesher (kernels_mesh.cu): the staged chunk layout, the face masks of a
// row (chunk.rs:59-245 as shift / AND tests on 32-voxel rows), and the instance tables (quad.rs:16-40,
// texture.rs:238-243).  Kept apart from the kernels so that tests/test_device_mesh_on_host.py can compile exactly this
// source for the host and compare the quads it implies with the oracle without a GPU.
#pragma once

struct RowFaces {
    uint32_t m[6];      // LEFT, RIGHT, BACK, FRONT, DOWN, UP
};

struct __align__(16) ChunkSmem {
    uint32_t occ[VX_ROWS + 64];   // rows -32..-1 and 1024..1055 are zero (y borders always emit)
    uint32_t back[32];            // -z neighbour: row (y, 31) per y, or 0
    uint32_t front[32];           // +z neighbour: row (y, 0) per y, or 0
    uint32_t xm[32];              // -x neighbour: word y, bit z = its voxel (31, y, z)
    uint32_t xp[32];              // +x neighbour: word y, bit z = its voxel (0, y, z)
};

__device__ __forceinline__ RowFaces row_faces(const ChunkSmem &s, uint32_t r)
{
    const uint32_t y = r >> 5, z = r & 31u;
    const uint32_t o = s.occ[32 + r];
    RowFaces f;
    f.m[VX_LEFT] = o & ~((o << 1) | ((s.xm[y] >> z) & 1u));
    f.m[VX_RIGHT] = o & ~((o >> 1) | (((s.xp[y] >> z) & 1u) << 31));
    const uint32_t bw = (z > 0) ? s.occ[32 + r - 1] : s.back[y];
    const uint32_t fw = (z < 31) ? s.occ[32 + r + 1] : s.front[y];
    f.m[VX_BACK] = o & ~bw;
    f.m[VX_FRONT] = o & ~fw;
    f.m[VX_DOWN] = o & ~s.occ[32 + r - 32];
    f.m[VX_UP] = o & ~s.occ[32 + r + 32];
    return f;
}

// Face masks of the four consecutive rows 4t .. 4t+3 (same y, z = z0 .. z0+3) for the per-chunk row scan: the rows
// themselves and the rows below / above arrive as three 128-bit shared loads instead of 4 x 7 scalar loads at a
// stride of four words (a 4-way bank conflict: the 12 % of conflicting wavefronts ncu still showed in round 2).
__device__ __forceinline__ void row_faces4(const ChunkSmem &s, uint32_t t, RowFaces (&f)[4])
{
    const uint32_t r0 = 4u * t, y = r0 >> 5, z0 = r0 & 31u;
    const uint4 o4 = *reinterpret_cast<const uint4 *>(&s.occ[32 + r0]);
    const uint4 d4 = *reinterpret_cast<const uint4 *>(&s.occ[r0]);              // rows r - 32
    const uint4 u4 = *reinterpret_cast<const uint4 *>(&s.occ[64 + r0]);         // rows r + 32
    const uint32_t o[4] = {o4.x, o4.y, o4.z, o4.w}, dn[4] = {d4.x, d4.y, d4.z, d4.w}, up[4] = {u4.x, u4.y, u4.z, u4.w};
    const uint32_t before = (z0 > 0u) ? s.occ[32 + r0 - 1] : s.back[y];
    const uint32_t after = (z0 < 28u) ? s.occ[32 + r0 + 4] : s.front[y];
    const uint32_t xm = s.xm[y] >> z0, xp = s.xp[y] >> z0;
#pragma unroll
    for (uint32_t j = 0; j < 4; ++j) {
        const uint32_t oj = o[j], bw = j ? o[j - 1] : before, fw = (j < 3u) ? o[j + 1] : after;
        f[j].m[VX_LEFT] = oj & ~((oj << 1) | ((xm >> j) & 1u));
        f[j].m[VX_RIGHT] = oj & ~((oj >> 1) | (((xp >> j) & 1u) << 31));
        f[j].m[VX_BACK] = oj & ~bw;
        f[j].m[VX_FRONT] = oj & ~fw;
        f[j].m[VX_DOWN] = oj & ~dn[j];
        f[j].m[VX_UP] = oj & ~up[j];
    }
}

__device__ __forceinline__ uint32_t row_count(const RowFaces &f)
{
    return __popc(f.m[0]) + __popc(f.m[1]) + __popc(f.m[2]) + __popc(f.m[3]) + __popc(f.m[4]) + __popc(f.m[5]);
}

// quad.rs:16-40 evaluated in f32: the six rotations as
// InstanceRaw column words 0..11, with e = cos(pi/2 as f32) = -4.371139e-8 and
// d = -sin(pi as f32) = 8.742278e-8.  Signed zeros are written as +0.
#define VX_E 0xb33bbd2eu     /* -4.371139e-8 */
#define VX_NE 0x333bbd2eu
#define VX_D 0x33bbbd2eu     /*  8.742278e-8 */
#define VX_ND 0xb3bbbd2eu
#define VX_1 0x3f800000u
#define VX_N1 0xbf800000u
__constant__ uint32_t c_rot[6][12] = {
    // column 0 (r0,r1,r2,0), column 1, column 2  -- R[row][col] from App. B
    /* LEFT  rows (e,0,-1)(0,1,0)(1,0,e)   */ {VX_E, 0, VX_1, 0, 0, VX_1, 0, 0, VX_N1, 0, VX_E, 0},
    /* RIGHT rows (e,0,1)(0,1,0)(-1,0,e)   */ {VX_E, 0, VX_N1, 0, 0, VX_1, 0, 0, VX_1, 0, VX_E, 0},
    /* BACK  rows (-1,0,-d)(0,1,0)(d,0,-1) */ {VX_N1, 0, VX_D, 0, 0, VX_1, 0, 0, VX_ND, 0, VX_N1, 0},
    /* FRONT identity                      */ {VX_1, 0, 0, 0, 0, VX_1, 0, 0, 0, 0, VX_1, 0},
    /* DOWN  rows (1,0,0)(0,e,-1)(0,1,e)   */ {VX_1, 0, 0, 0, 0, VX_E, VX_1, 0, 0, VX_N1, VX_E, 0},
    /* UP    rows (1,0,0)(0,e,1)(0,-1,e)   */ {VX_1, 0, 0, 0, 0, VX_E, VX_N1, 0, 0, VX_1, VX_E, 0},
};
// texture.rs:238-243 TEXTURE_INDEX_LIST and the slot each face reads
// (chunk.rs:77-240: LEFT [0], RIGHT [1], BACK [2], FRONT [3], DOWN [5], UP [4])
__constant__ uint32_t c_tex[2][6] = {
    /* GRASS [0,0,0,0,2,1] */ {0, 0, 0, 0, 1, 2},
    /* DIRT  [1,1,1,1,1,1] */ {1, 1, 1, 1, 1, 1},
};

