// kernels_mesh.cu -- hidden-face culling and instance emission on occupancy
// bitmasks.  Replaces Chunk::cull_unseen_triangles (chunk.rs:59-245),
// Quad::new / Quad::to_raw (quad.rs:14-54) and the Vec<InstanceRaw> build of
// ChunkMesh::new (chunk.rs:254-260).
//
// Data model: a chunk is 1024 row words, row r = y*32 + z, bit x set <=> the
// voxel is not AIR (cube.rs:16-21); an optional second plane marks DIRT.
// The six face tests of one row are shifts/ANDs of row words:
//   LEFT  = o & ~((o << 1) | bit31 of the -x neighbour row)     chunk.rs:77-109
//   RIGHT = o & ~((o >> 1) | bit0 of the +x neighbour row)      chunk.rs:110-140
//   BACK  = o & ~row(y, z-1)   FRONT = o & ~row(y, z+1)         chunk.rs:141-204
//   DOWN  = o & ~row(y-1, z)   UP    = o & ~row(y+1, z)         chunk.rs:205-240
// with "neighbour absent" = all-air on x/z borders (border faces emitted) and
// y borders always emitted.  Quads leave in the reference's order: cubes in
// (y,z,x) order, faces LEFT,RIGHT,BACK,FRONT,DOWN,UP -- an exclusive scan of
// per-row counts keeps that order, so parity is a plain compare.
// The emit kernel is HBM-write bound (68 B per quad).
#include "vx_common.cuh"

namespace {

constexpr int kMeshThreads = 256;           // 4 consecutive rows per thread
constexpr int kWindow = 4096;               // quad descriptors staged per pass

// ---------------------------------------------------------------------------
// bytes -> bit planes (vx_mesh with caller-provided Vec<Cube>)
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t gather_msb4(uint32_t m)
{
    // bit k of the result = bit 7 of byte k of m
    return ((((m >> 7) & 0x01010101u) * 0x01020408u) >> 24) & 15u;
}

__global__ void __launch_bounds__(256)
vx_pack_kernel(const uint8_t *__restrict__ blocks, uint32_t n, uint32_t *__restrict__ occ,
               uint32_t *__restrict__ typ, uint32_t *status)
{
    // one thread per 16 voxels (uint4); a lane pair covers one 32-voxel row
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t total = (size_t)n * (VX_VOXELS / 16);
    uint32_t o16 = 0, d16 = 0, bad = 0;
    if (t < total) {
        const uint4 v = __ldg(reinterpret_cast<const uint4 *>(blocks) + t);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const uint32_t air = __vcmpeq4(w[i], 0x02020202u);     // 0xff where AIR
            const uint32_t dirt = __vcmpeq4(w[i], 0x01010101u);
            const uint32_t big = __vcmpgtu4(w[i], 0x02020202u);
            o16 |= gather_msb4(~air) << (4 * i);
            d16 |= gather_msb4(dirt) << (4 * i);
            bad |= big;
        }
    }
    const uint32_t o_hi = __shfl_down_sync(VX_FULL, o16, 1);
    const uint32_t d_hi = __shfl_down_sync(VX_FULL, d16, 1);
    if (t < total && (threadIdx.x & 1) == 0) {
        occ[t >> 1] = o16 | (o_hi << 16);
        typ[t >> 1] = d16 | (d_hi << 16);
    }
    if (bad && status) atomicOr(status, VX_ST_BAD_BLOCK);
}

// ---------------------------------------------------------------------------
// neighbour resolution on the device: open-addressing hash of the coordinate
// list (chunk_builder.rs:134-169 looks the four neighbours up in World.chunks)
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t coord_hash(int32_t x, int32_t y, int32_t z)
{
    uint32_t h = (uint32_t)x * 0x9E3779B1u;
    h ^= ((uint32_t)z + 0x7F4A7C15u) * 0x85EBCA77u;
    h ^= ((uint32_t)y + 0x165667B1u) * 0xC2B2AE3Du;
    h ^= h >> 15;
    h *= 0x27D4EB2Fu;
    h ^= h >> 13;
    return h;
}

__global__ void vx_ht_insert_kernel(const int32_t *__restrict__ coords, uint32_t n,
                                    int32_t *table, uint32_t mask)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int32_t x = coords[3 * i], y = coords[3 * i + 1], z = coords[3 * i + 2];
    uint32_t s = coord_hash(x, y, z) & mask;
    for (;;) {
        const int32_t prev = atomicCAS(&table[s], -1, (int32_t)i);
        if (prev == -1) return;
        // duplicate coordinate: keep the smallest index so lookups are deterministic
        if (coords[3 * prev] == x && coords[3 * prev + 1] == y && coords[3 * prev + 2] == z) {
            atomicMin(&table[s], (int32_t)i);
            return;
        }
        s = (s + 1) & mask;
    }
}

__global__ void vx_resolve_kernel(const vx_resolve_args a)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (a.hi - a.lo) * 4u) return;
    const uint32_t i = a.lo + (t >> 2), k = t & 3u;
    int32_t ref = VX_NBR_NONE;
    bool present = false;
    const bool wanted = a.nbr_mask ? ((a.nbr_mask[i] >> k) & 1u) : true;
    if (wanted) {
        // chunk_builder.rs:134-169: k = 0:-x 1:+x 2:-z 3:+z
        const int32_t x = (int32_t)((uint32_t)a.coords[3 * i] + (k == 0 ? 0xffffffffu : (k == 1 ? 1u : 0u)));
        const int32_t y = a.coords[3 * i + 1];
        const int32_t z = (int32_t)((uint32_t)a.coords[3 * i + 2] + (k == 2 ? 0xffffffffu : (k == 3 ? 1u : 0u)));
        const uint32_t mask = a.table_size - 1u;
        uint32_t s = coord_hash(x, y, z) & mask;
        for (;;) {
            const int32_t j = a.table[s];
            if (j < 0) break;
            if (a.coords[3 * j] == x && a.coords[3 * j + 1] == y && a.coords[3 * j + 2] == z) {
                present = true;
                if ((uint32_t)j >= a.lo && (uint32_t)j < a.hi) ref = j - (int32_t)a.lo;
                break;
            }
            s = (s + 1) & mask;
        }
        if (a.nbr_mask) present = true;          // the mask says so, wherever the chunk is
        if (ref < 0 && present && a.halo_items) {
            // present but not resident on this device (outside the list, or in
            // another GPU's shard): its border slab is regenerated from noise
            const uint32_t slot = atomicAdd(a.halo_count, 1u);
            vx_halo_item it;
            it.cx = x; it.cy = y; it.cz = z; it.kind = (int32_t)k;
            a.halo_items[slot] = it;
            ref = vx_halo_ref((int)slot);
        }
    }
    a.nbr[t] = ref;
}

// caller-provided nbr_idx (>= 0 index, anything negative = None)
__global__ void vx_remap_nbr_kernel(const int32_t *__restrict__ in, int32_t *__restrict__ out, uint32_t n4)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n4) out[t] = in[t] >= 0 ? in[t] : VX_NBR_NONE;
}

// ---------------------------------------------------------------------------
// face masks
// ---------------------------------------------------------------------------
struct RowFaces {
    uint32_t m[6];      // LEFT, RIGHT, BACK, FRONT, DOWN, UP
};

struct ChunkSmem {
    uint32_t occ[VX_ROWS + 64];   // rows -32..-1 and 1024..1055 are zero (y borders always emit)
    uint32_t back[32];            // -z neighbour: row (y, 31) per y, or 0
    uint32_t front[32];           // +z neighbour: row (y, 0) per y, or 0
    uint32_t xm[32];              // -x neighbour: word y, bit z = its voxel (31, y, z)
    uint32_t xp[32];              // +x neighbour: word y, bit z = its voxel (0, y, z)
};

// Stage one chunk and its one-voxel halo.  Must be called by all threads.
template <int THREADS>
__device__ __forceinline__ void stage_chunk(ChunkSmem &s, const vx_mesh_args &a, uint32_t m)
{
    const uint32_t tid = threadIdx.x;
    const int32_t nb0 = a.nbr[4 * m + 0], nb1 = a.nbr[4 * m + 1];
    const int32_t nb2 = a.nbr[4 * m + 2], nb3 = a.nbr[4 * m + 3];
    const uint32_t *own = a.occ + (size_t)m * VX_ROWS;
    for (uint32_t r = tid; r < 32; r += THREADS) {
        s.occ[r] = 0;
        s.occ[32 + VX_ROWS + r] = 0;
    }
    for (uint32_t r = tid; r < VX_ROWS; r += THREADS) s.occ[32 + r] = __ldg(own + r);
    if (tid < 32) {
        const uint32_t y = tid;
        uint32_t b = 0, f = 0;
        if (nb2 >= 0) b = __ldg(a.occ + (size_t)nb2 * VX_ROWS + y * 32 + 31);
        else if (nb2 <= -2) b = __ldg(a.halo + (size_t)vx_halo_slot(nb2) * 32 + y);
        if (nb3 >= 0) f = __ldg(a.occ + (size_t)nb3 * VX_ROWS + y * 32 + 0);
        else if (nb3 <= -2) f = __ldg(a.halo + (size_t)vx_halo_slot(nb3) * 32 + y);
        s.back[y] = b;
        s.front[y] = f;
        if (nb0 <= -2) s.xm[y] = __ldg(a.halo + (size_t)vx_halo_slot(nb0) * 32 + y);
        else if (nb0 == VX_NBR_NONE) s.xm[y] = 0;
        if (nb1 <= -2) s.xp[y] = __ldg(a.halo + (size_t)vx_halo_slot(nb1) * 32 + y);
        else if (nb1 == VX_NBR_NONE) s.xp[y] = 0;
    }
    // in-batch x neighbours: transpose bit 31 / bit 0 of their 1024 rows into
    // 32 words (word y, bit z) with one ballot per warp-row-group
    if (nb0 >= 0 || nb1 >= 0) {
        const uint32_t lane = tid & 31u;
        for (uint32_t r0 = (tid & ~31u); r0 < VX_ROWS; r0 += THREADS) {
            const uint32_t r = r0 + lane;           // y = r0 / 32, z = lane
            if (nb0 >= 0) {
                const uint32_t w = __ldg(a.occ + (size_t)nb0 * VX_ROWS + r);
                const uint32_t bal = __ballot_sync(VX_FULL, (w >> 31) & 1u);
                if (lane == 0) s.xm[r0 >> 5] = bal;
            }
            if (nb1 >= 0) {
                const uint32_t w = __ldg(a.occ + (size_t)nb1 * VX_ROWS + r);
                const uint32_t bal = __ballot_sync(VX_FULL, w & 1u);
                if (lane == 0) s.xp[r0 >> 5] = bal;
            }
        }
    }
    __syncthreads();
}

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

__device__ __forceinline__ uint32_t row_count(const RowFaces &f)
{
    return __popc(f.m[0]) + __popc(f.m[1]) + __popc(f.m[2]) + __popc(f.m[3]) + __popc(f.m[4]) + __popc(f.m[5]);
}

// ---------------------------------------------------------------------------
// pass 1: quads per chunk
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kMeshThreads)
vx_mesh_count_kernel(const vx_mesh_args a)
{
    __shared__ ChunkSmem s;
    __shared__ uint32_t warp_sums[kMeshThreads / 32];
    const uint32_t m = blockIdx.x;
    if (a.mesh_mask && a.mesh_mask[m] == 0) {
        if (threadIdx.x == 0) a.counts[m] = 0;
        return;
    }
    stage_chunk<kMeshThreads>(s, a, m);
    uint32_t c = 0;
#pragma unroll
    for (uint32_t j = 0; j < 4; ++j) c += row_count(row_faces(s, threadIdx.x * 4 + j));
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(VX_FULL, c, d);
    if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t tot = 0;
#pragma unroll
        for (int w = 0; w < kMeshThreads / 32; ++w) tot += warp_sums[w];
        a.counts[m] = tot;
    }
}

// ---------------------------------------------------------------------------
// exclusive scan of the counts -> uint64 offsets[n+1]   (single CTA)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
vx_scan_kernel(const uint32_t *__restrict__ counts, uint64_t *__restrict__ offsets, uint32_t n)
{
    __shared__ unsigned long long warp_tot[32];
    __shared__ unsigned long long carry_s;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, wid = tid >> 5;
    if (tid == 0) carry_s = 0;
    __syncthreads();
    for (uint32_t base = 0; base < n; base += 1024) {
        const uint32_t i = base + tid;
        const unsigned long long v = (i < n) ? counts[i] : 0ull;
        unsigned long long inc = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned long long t = __shfl_up_sync(VX_FULL, inc, d);
            if (lane >= (uint32_t)d) inc += t;
        }
        if (lane == 31) warp_tot[wid] = inc;
        __syncthreads();
        if (wid == 0) {
            unsigned long long w = warp_tot[lane];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const unsigned long long t = __shfl_up_sync(VX_FULL, w, d);
                if (lane >= (uint32_t)d) w += t;
            }
            warp_tot[lane] = w;       // inclusive
        }
        __syncthreads();
        const unsigned long long carry = carry_s;
        const unsigned long long before = carry + (wid ? warp_tot[wid - 1] : 0ull) + (inc - v);
        if (i < n) offsets[i] = before;
        __syncthreads();
        if (tid == 1023) carry_s = before + v;
        __syncthreads();
    }
    if (tid == 0) offsets[n] = carry_s;
}

// ---------------------------------------------------------------------------
// pass 2: emission
// ---------------------------------------------------------------------------
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

struct EmitSmem {
    ChunkSmem c;
    uint32_t typ[VX_ROWS];
    uint32_t desc[kWindow];
    uint32_t lut[16][17];         // [type*8 + face][word]
    uint32_t pos[3][32];          // f32 bits of world x / y / z per local coordinate
    uint32_t warp_tot[kMeshThreads / 32];
};

__global__ void __launch_bounds__(kMeshThreads)
vx_mesh_emit_kernel(const vx_mesh_args a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    EmitSmem &s = *reinterpret_cast<EmitSmem *>(smem_raw);
    const uint32_t m = a.m0 + blockIdx.x, tid = threadIdx.x, lane = tid & 31u, wid = tid >> 5;

    // this launch covers chunks [m0, m0 + gridDim.x); its quads land at
    // out[offsets[m] - quad_base]
    const unsigned long long launch_quads = a.offsets[a.m0 + gridDim.x] - a.quad_base;
    if (launch_quads > a.capacity) {
        if (blockIdx.x == 0 && tid == 0 && a.status) atomicOr(a.status, VX_ST_OVERFLOW);
        return;
    }
    const uint32_t count = a.counts[m];
    if (count == 0) return;                      // also covers mesh_mask == 0

    // constant tables -> shared (conflict-free per-lane lookups later)
    for (uint32_t i = tid; i < 16 * 17; i += kMeshThreads) {
        const uint32_t tf = i / 17, w = i - tf * 17, type = tf >> 3, face = tf & 7u;
        uint32_t v = 0;
        if (face < 6) {
            if (w < 12) v = c_rot[face][w];
            else if (w == 15) v = VX_1;
            else if (w == 16) v = c_tex[type][face];
        }
        s.lut[tf][w] = v;
    }
    if (tid < 96) {
        // coordinate.rs:60-65 + chunk.rs:66,68,76 + quad.rs:42: (chunk*32 + local) as f32
        const uint32_t axis = tid >> 5, v = tid & 31u;
        const uint32_t c = (uint32_t)a.coords[3 * m + (axis == 0 ? 0 : (axis == 1 ? 1 : 2))];
        s.pos[axis][v] = __float_as_uint((float)(int32_t)(c * 32u + v));
    }
    if (a.typ) {
        for (uint32_t r = tid; r < VX_ROWS; r += kMeshThreads) s.typ[r] = __ldg(a.typ + (size_t)m * VX_ROWS + r);
    }
    stage_chunk<kMeshThreads>(s.c, a, m);        // ends with __syncthreads()

    // per-thread: 4 consecutive rows, exclusive scan of row counts in row order
    RowFaces rf[4];
    uint32_t cnt[4], mine = 0;
#pragma unroll
    for (uint32_t j = 0; j < 4; ++j) {
        rf[j] = row_faces(s.c, tid * 4 + j);
        cnt[j] = row_count(rf[j]);
        mine += cnt[j];
    }
    uint32_t inc = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(VX_FULL, inc, d);
        if (lane >= (uint32_t)d) inc += t;
    }
    if (lane == 31) s.warp_tot[wid] = inc;
    __syncthreads();
    uint32_t row_base = inc - mine;
#pragma unroll
    for (uint32_t w = 0; w < kMeshThreads / 32; ++w)
        if (w < wid) row_base += s.warp_tot[w];

    uint32_t *out32 = reinterpret_cast<uint32_t *>(a.out) + (size_t)(a.offsets[m] - a.quad_base) * 17u;

    for (uint32_t w0 = 0; w0 < count; w0 += kWindow) {
        // --- descriptors of quads [w0, w0 + kWindow) in emission order -------------
        uint32_t k = row_base;
#pragma unroll
        for (uint32_t j = 0; j < 4; ++j) {
            const uint32_t r = tid * 4 + j;
            if (k + cnt[j] > w0 && k < w0 + kWindow) {
                uint32_t any = rf[j].m[0] | rf[j].m[1] | rf[j].m[2] | rf[j].m[3] | rf[j].m[4] | rf[j].m[5];
                const uint32_t tw = a.typ ? s.typ[r] : 0u;
                uint32_t kk = k;
                while (any) {
                    const uint32_t x = __ffs(any) - 1;
                    any &= any - 1;
                    // descriptor: bits 0-14 voxel idx (x + 32 z + 1024 y), 15-17 face, 18 type
                    const uint32_t base = (r << 5) | x | (((tw >> x) & 1u) << 18);
#pragma unroll
                    for (uint32_t f = 0; f < 6; ++f) {
                        if ((rf[j].m[f] >> x) & 1u) {
                            const uint32_t slot = kk - w0;
                            if (slot < (uint32_t)kWindow) s.desc[slot] = base | (f << 15);
                            ++kk;
                        }
                    }
                }
            }
            k += cnt[j];
        }
        __syncthreads();
        // --- expansion: 17 words per quad, fully coalesced 32-bit stores ---------
        const uint32_t nq = min((uint32_t)kWindow, count - w0);
        const uint32_t nwords = nq * 17u;
        uint32_t *dst = out32 + (size_t)w0 * 17u;
        for (uint32_t i = tid; i < nwords; i += kMeshThreads) {
            const uint32_t q = i / 17u, c = i - q * 17u;
            const uint32_t d = s.desc[q];
            uint32_t v = s.lut[(d >> 15) & 15u][c];
            const uint32_t ax = c - 12u;
            if (ax < 3u) {
                // word 12: x (bits 0-4), 13: y (bits 10-14), 14: z (bits 5-9)
                const uint32_t sh = (0x050a00u >> (8u * ax)) & 0xffu;
                v = s.pos[ax][(d >> sh) & 31u];
            }
            dst[i] = v;
        }
        __syncthreads();
    }
}

}  // namespace

// ---------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------
cudaError_t vx_launch_pack(const uint8_t *blocks, uint32_t n, uint32_t *occ, uint32_t *typ,
                           uint32_t *status, cudaStream_t s, uint64_t *launches)
{
    if (n == 0) return cudaSuccess;
    const size_t total = (size_t)n * (VX_VOXELS / 16);
    vx_pack_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(blocks, n, occ, typ, status);
    if (launches) ++*launches;
    return cudaGetLastError();
}

cudaError_t vx_launch_resolve(const vx_resolve_args &a, cudaStream_t s, uint64_t *launches)
{
    if (a.n == 0 || a.hi <= a.lo) return cudaSuccess;
    cudaError_t e = cudaMemsetAsync(a.table, 0xff, sizeof(int32_t) * a.table_size, s);
    if (e != cudaSuccess) return e;
    if (a.halo_count) {
        e = cudaMemsetAsync(a.halo_count, 0, sizeof(uint32_t), s);
        if (e != cudaSuccess) return e;
    }
    vx_ht_insert_kernel<<<(a.n + 255) / 256, 256, 0, s>>>(a.coords, a.n, a.table, a.table_size - 1);
    vx_resolve_kernel<<<((a.hi - a.lo) * 4 + 255) / 256, 256, 0, s>>>(a);
    if (launches) *launches += 2;
    return cudaGetLastError();
}

cudaError_t vx_launch_remap_nbr(const int32_t *nbr_in, int32_t *nbr_out, uint32_t n, cudaStream_t s,
                                uint64_t *launches)
{
    if (n == 0) return cudaSuccess;
    vx_remap_nbr_kernel<<<(n * 4 + 255) / 256, 256, 0, s>>>(nbr_in, nbr_out, n * 4);
    if (launches) ++*launches;
    return cudaGetLastError();
}

cudaError_t vx_launch_mesh_count(const vx_mesh_args &a, cudaStream_t s, uint64_t *launches)
{
    if (a.n == 0) return cudaSuccess;
    vx_mesh_count_kernel<<<a.n, kMeshThreads, 0, s>>>(a);
    if (launches) ++*launches;
    return cudaGetLastError();
}

cudaError_t vx_launch_scan(const uint32_t *counts, uint64_t *offsets, uint32_t n, cudaStream_t s,
                           uint64_t *launches)
{
    vx_scan_kernel<<<1, 1024, 0, s>>>(counts, offsets, n);
    if (launches) ++*launches;
    return cudaGetLastError();
}

cudaError_t vx_launch_mesh_emit(const vx_mesh_args &a, uint32_t len, cudaStream_t s, uint64_t *launches)
{
    if (len == 0) return cudaSuccess;
    static_assert(sizeof(EmitSmem) <= 48 * 1024, "emit kernel must fit the default dynamic smem limit");
    vx_mesh_emit_kernel<<<len, kMeshThreads, sizeof(EmitSmem), s>>>(a);
    if (launches) ++*launches;
    return cudaGetLastError();
}
