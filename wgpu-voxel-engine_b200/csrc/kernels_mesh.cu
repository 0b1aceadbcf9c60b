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
#include <cstddef>

#include "vx_common.cuh"

namespace {

constexpr int kMeshThreads = 256;           // 4 consecutive rows per thread

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
    const uint32_t K = a.nbr_y ? 6u : 4u;                    // neighbour slots per chunk
    if (t >= (a.hi - a.lo) * K) return;
    const uint32_t li = t / K, i = a.lo + li, k = t - li * K;
    int32_t ref = VX_NBR_NONE;
    bool present = false;
    const bool wanted = a.nbr_mask ? ((a.nbr_mask[i] >> k) & 1u) : true;
    if (wanted) {
        // chunk_builder.rs:134-169: k = 0:-x 1:+x 2:-z 3:+z   (extension: 4:-y 5:+y)
        const int32_t x = (int32_t)((uint32_t)a.coords[3 * i] + (k == 0 ? 0xffffffffu : (k == 1 ? 1u : 0u)));
        const int32_t y = (int32_t)((uint32_t)a.coords[3 * i + 1] + (k == 4 ? 0xffffffffu : (k == 5 ? 1u : 0u)));
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
    if (k < 4u) a.nbr[4u * li + k] = ref;
    else a.nbr_y[2u * li + (k - 4u)] = ref;
}

// Small lists (<= kSmallResolve chunks): one CTA builds the hash table in shared
// memory and answers the lookups in the same launch (no memsets, no second kernel).
constexpr uint32_t kSmallResolve = 4096;

__device__ __forceinline__ int32_t lookup_in(const int32_t *table, uint32_t mask, const int32_t *coords,
                                             int32_t x, int32_t y, int32_t z)
{
    uint32_t s = coord_hash(x, y, z) & mask;
    for (;;) {
        const int32_t j = table[s];
        if (j < 0) return -1;
        if (coords[3 * j] == x && coords[3 * j + 1] == y && coords[3 * j + 2] == z) return j;
        s = (s + 1) & mask;
    }
}

__global__ void __launch_bounds__(1024) vx_resolve_small_kernel(const vx_resolve_args a)
{
    // dynamic shared memory: int32 table[table_size] | int32 coords[3 * n]
    extern __shared__ int32_t rs_smem[];
    __shared__ uint32_t halo_n;
    int32_t *table = rs_smem;
    int32_t *co = rs_smem + a.table_size;
    const uint32_t mask = a.table_size - 1u;
    for (uint32_t i = threadIdx.x; i <= mask; i += blockDim.x) table[i] = -1;
    for (uint32_t i = threadIdx.x; i < 3u * a.n; i += blockDim.x) co[i] = __ldg(a.coords + i);
    if (threadIdx.x == 0) halo_n = 0;
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < a.n; i += blockDim.x) {
        const int32_t x = co[3 * i], y = co[3 * i + 1], z = co[3 * i + 2];
        uint32_t s = coord_hash(x, y, z) & mask;
        for (;;) {
            const int32_t prev = atomicCAS(&table[s], -1, (int32_t)i);
            if (prev == -1) break;
            if (co[3 * prev] == x && co[3 * prev + 1] == y && co[3 * prev + 2] == z) {
                atomicMin(&table[s], (int32_t)i);      // duplicates: the smallest index answers lookups
                break;
            }
            s = (s + 1) & mask;
        }
    }
    __syncthreads();
    const uint32_t K = a.nbr_y ? 6u : 4u;
    for (uint32_t t = threadIdx.x; t < (a.hi - a.lo) * K; t += blockDim.x) {
        const uint32_t li = t / K, i = a.lo + li, k = t - li * K;
        int32_t ref = VX_NBR_NONE;
        const bool wanted = a.nbr_mask ? ((a.nbr_mask[i] >> k) & 1u) : true;
        if (wanted) {
            const int32_t x = (int32_t)((uint32_t)co[3 * i] + (k == 0 ? 0xffffffffu : (k == 1 ? 1u : 0u)));
            const int32_t y = (int32_t)((uint32_t)co[3 * i + 1] + (k == 4 ? 0xffffffffu : (k == 5 ? 1u : 0u)));
            const int32_t z = (int32_t)((uint32_t)co[3 * i + 2] + (k == 2 ? 0xffffffffu : (k == 3 ? 1u : 0u)));
            const int32_t j = lookup_in(table, mask, co, x, y, z);
            bool present = a.nbr_mask != nullptr || j >= 0;
            if (j >= 0 && (uint32_t)j >= a.lo && (uint32_t)j < a.hi) ref = j - (int32_t)a.lo;
            if (ref < 0 && present && a.halo_items) {
                const uint32_t slot = atomicAdd(&halo_n, 1u);
                vx_halo_item it;
                it.cx = x; it.cy = y; it.cz = z; it.kind = (int32_t)k;
                a.halo_items[slot] = it;
                ref = vx_halo_ref((int)slot);
            }
        }
        if (k < 4u) a.nbr[4u * li + k] = ref;
        else a.nbr_y[2u * li + (k - 4u)] = ref;
    }
    __syncthreads();
    if (threadIdx.x == 0 && a.halo_count) *a.halo_count = halo_n;
}

// caller-provided nbr_idx (>= 0 index, anything negative = None)
// an index >= n would make the mesher read another allocation: it becomes None and raises a status bit
__global__ void vx_remap_nbr_kernel(const int32_t *__restrict__ in, int32_t *__restrict__ out, uint32_t n4, uint32_t n,
                                    uint32_t *status)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n4) return;
    const int32_t v = in[t];
    if (v >= 0 && (uint32_t)v >= n) {
        out[t] = VX_NBR_NONE;
        if (status) atomicOr(status, VX_ST_BAD_NBR);
    } else {
        out[t] = v >= 0 ? v : VX_NBR_NONE;
    }
}

// ---------------------------------------------------------------------------
// face masks
// ---------------------------------------------------------------------------
#include "vx_mesh_rows.cuh"

// Stage one chunk and its one-voxel halo.  Must be called by all threads.
template <int THREADS>
__device__ __forceinline__ void stage_chunk(ChunkSmem &s, const vx_mesh_args &a, uint32_t m)
{
    const uint32_t tid = threadIdx.x;
    const int4 nb = __ldg(reinterpret_cast<const int4 *>(a.nbr) + m);      // (the table is the library's own: aligned)
    const int32_t nb0 = nb.x, nb1 = nb.y, nb2 = nb.z, nb3 = nb.w;
    const uint32_t *own = a.occ + (size_t)m * VX_ROWS;
    // y borders: zero rows = the reference's "always emit" (chunk.rs:216-222,234-240).  Extension: a present
    // vertical neighbour supplies its facing plane (rows (31, z) of the chunk below, rows (0, z) of the one above).
    const int32_t nby0 = a.nbr_y ? a.nbr_y[2 * m + 0] : VX_NBR_NONE, nby1 = a.nbr_y ? a.nbr_y[2 * m + 1] : VX_NBR_NONE;
    for (uint32_t r = tid; r < 32; r += THREADS) {
        uint32_t below = 0, above = 0;
        if (nby0 >= 0) below = __ldg(a.occ + (size_t)nby0 * VX_ROWS + 31u * 32u + r);
        else if (nby0 <= -2) below = __ldg(a.halo + (size_t)vx_halo_slot(nby0) * 32 + r);
        if (nby1 >= 0) above = __ldg(a.occ + (size_t)nby1 * VX_ROWS + r);
        else if (nby1 <= -2) above = __ldg(a.halo + (size_t)vx_halo_slot(nby1) * 32 + r);
        s.occ[r] = below;
        s.occ[32 + VX_ROWS + r] = above;
    }
    if (THREADS * 4 == VX_ROWS) {
        // one 128-bit load and store per thread (the occupancy array is the library's own: 4 KiB per chunk, aligned)
        *reinterpret_cast<uint4 *>(&s.occ[32 + 4 * tid]) = __ldg(reinterpret_cast<const uint4 *>(own) + tid);
    } else {
        for (uint32_t r = tid; r < VX_ROWS; r += THREADS) s.occ[32 + r] = __ldg(own + r);
    }
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
        // all loads of a thread are issued before the first ballot (the count kernel is bound by this latency)
        constexpr uint32_t kRounds = VX_ROWS / THREADS;
        static_assert(VX_ROWS % THREADS == 0 && THREADS % 32 == 0, "whole warp-row-groups per round");
        const uint32_t lane = tid & 31u;
        uint32_t wm[kRounds], wp[kRounds];
#pragma unroll
        for (uint32_t k = 0; k < kRounds; ++k) {
            const uint32_t r = (tid & ~31u) + k * THREADS + lane;          // y = r / 32, z = lane
            wm[k] = nb0 >= 0 ? __ldg(a.occ + (size_t)nb0 * VX_ROWS + r) : 0u;
            wp[k] = nb1 >= 0 ? __ldg(a.occ + (size_t)nb1 * VX_ROWS + r) : 0u;
        }
#pragma unroll
        for (uint32_t k = 0; k < kRounds; ++k) {
            const uint32_t y = ((tid & ~31u) + k * THREADS) >> 5;
            if (nb0 >= 0) {
                const uint32_t bal = __ballot_sync(VX_FULL, (wm[k] >> 31) & 1u);
                if (lane == 0) s.xm[y] = bal;
            }
            if (nb1 >= 0) {
                const uint32_t bal = __ballot_sync(VX_FULL, wp[k] & 1u);
                if (lane == 0) s.xp[y] = bal;
            }
        }
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------
// exclusive scan of the counts -> uint64 offsets[n+1], run by ONE CTA of
// kMeshThreads threads: the last CTA of the count kernel to finish (ticket).
// ---------------------------------------------------------------------------
__device__ void scan_counts_cta(const uint32_t *counts, uint64_t *offsets, uint32_t n)
{
    // Exclusive scan of n counts by ONE CTA (the last one of the count kernel, so it is a serial tail of the launch):
    // every thread owns a contiguous run of ceil(n / threads) counts, sums it with independent loads in flight, the 256
    // sums are scanned across the CTA once, and the thread writes its run.  (A loop of CTA-wide scans over 256 counts
    // at a time costs a load latency and three barriers per 256 counts: ~25 us for 8192 chunks instead of ~4.)
    __shared__ unsigned long long warp_tot[kMeshThreads / 32];
    const uint32_t tid = threadIdx.x, lane = tid & 31u, wid = tid >> 5;
    const uint32_t per = (n + kMeshThreads - 1) / kMeshThreads;
    const uint32_t lo = min(n, tid * per), hi = min(n, lo + per);
    // whole runs of 4 counts / 2 offsets move as 128-bit accesses when the run length and the pointers allow it
    const bool vec = (per & 3u) == 0u && ((reinterpret_cast<uintptr_t>(counts) | reinterpret_cast<uintptr_t>(offsets)) & 15u) == 0u &&
                     hi - lo == per;
    unsigned long long sum = 0;
    if (vec) {
#pragma unroll 4
        for (uint32_t i = lo; i < hi; i += 4u) {
            const uint4 v = __ldcg(reinterpret_cast<const uint4 *>(counts + i));
            sum += (unsigned long long)v.x + v.y + v.z + v.w;
        }
    } else {
#pragma unroll 8
        for (uint32_t i = lo; i < hi; ++i) sum += __ldcg(counts + i);
    }
    unsigned long long inc = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned long long t = __shfl_up_sync(VX_FULL, inc, d);
        if (lane >= (uint32_t)d) inc += t;
    }
    if (lane == 31) warp_tot[wid] = inc;
    __syncthreads();
    unsigned long long run = inc - sum, total = 0;
#pragma unroll
    for (uint32_t w = 0; w < kMeshThreads / 32; ++w) {
        if (w < wid) run += warp_tot[w];
        total += warp_tot[w];
    }
    if (vec) {
#pragma unroll 4
        for (uint32_t i = lo; i < hi; i += 4u) {
            const uint4 v = __ldcg(reinterpret_cast<const uint4 *>(counts + i));
            const unsigned long long o1 = run + v.x, o2 = o1 + v.y, o3 = o2 + v.z;
            *reinterpret_cast<ulonglong2 *>(offsets + i) = make_ulonglong2(run, o1);
            *reinterpret_cast<ulonglong2 *>(offsets + i + 2u) = make_ulonglong2(o2, o3);
            run = o3 + v.w;
        }
    } else {
#pragma unroll 8
        for (uint32_t i = lo; i < hi; ++i) {
            offsets[i] = run;
            run += __ldcg(counts + i);
        }
    }
    if (tid == 0) offsets[n] = total;
}

// Called by every CTA of the count kernel after its count is written.
__device__ __forceinline__ void finish_with_scan(const vx_mesh_args &a)
{
    __shared__ uint32_t is_last;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();                                   // counts[m] visible before the ticket
        const uint32_t ticket = atomicAdd(a.ticket, 1u);
        is_last = (ticket == gridDim.x - 1u);
        if (is_last) *a.ticket = 0;                        // re-armed for the next launch
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
        scan_counts_cta(a.counts, a.offsets, a.n);
    }
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
        finish_with_scan(a);
        return;
    }
    stage_chunk<kMeshThreads>(s, a, m);
    uint32_t c = 0;
    RowFaces f4[4];
    row_faces4(s, threadIdx.x, f4);
#pragma unroll
    for (uint32_t j = 0; j < 4; ++j) c += row_count(f4[j]);
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
    finish_with_scan(a);
}

// ---------------------------------------------------------------------------
// extension: CubeType bytes of a batch from its occupancy planes, with the DIRT rule (a solid voxel whose +y
// neighbour voxel is solid is DIRT; the +y chunk's bottom plane comes through nbr_y like the mesher's)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kMeshThreads)
vx_blocks_ex_kernel(const vx_mesh_args a, uint8_t *__restrict__ blocks)
{
    __shared__ ChunkSmem s;
    const uint32_t m = blockIdx.x;
    stage_chunk<kMeshThreads>(s, a, m);
#pragma unroll
    for (uint32_t j = 0; j < 4; ++j) {
        const uint32_t r = threadIdx.x + j * kMeshThreads;      // row y*32 + z
        const uint32_t o = s.occ[32 + r], d = a.dirt_rule ? (o & s.occ[32 + r + 32]) : 0u;
        uint4 *dst = reinterpret_cast<uint4 *>(blocks + (size_t)m * VX_VOXELS + (size_t)r * 32u);
#pragma unroll
        for (uint32_t h = 0; h < 2u; ++h) {
            uint32_t w[4];
#pragma unroll
            for (uint32_t q = 0; q < 4u; ++q) {
                const uint32_t sh = 16u * h + 4u * q;
                const uint32_t air = ((~o >> sh) & 15u) * 0x00204081u & 0x01010101u;       // byte k = !occ bit
                const uint32_t drt = ((d >> sh) & 15u) * 0x00204081u & 0x01010101u;
                w[q] = (air << 1) | drt;                         // AIR = 2, DIRT = 1, GRASS = 0 (cube.rs:16-21)
            }
            dst[h] = make_uint4(w[0], w[1], w[2], w[3]);
        }
    }
}

// ---------------------------------------------------------------------------
// pass 2: emission
// ---------------------------------------------------------------------------
constexpr int kEmitThreads = 256;
constexpr int kEmitWarps = kEmitThreads / 32;
// CTAs per chunk: 1 for batches that fill the GPU (2 and 4 were measured slower there: every CTA repeats
// the chunk set-up), kEmitSplitSmall for short lists, where one CTA per chunk would leave most SMs idle
// and the call is latency-bound (the reference's one-chunk-per-call site, chunk_builder.rs:80-84)
constexpr int kEmitSplitSmall = 4;
constexpr uint32_t kEmitSmallList = 148;     // lists up to this many chunks use kEmitSplitSmall
constexpr int kLutStride = 20;               // words per LUT row (80 B: 16-B aligned, bank-spread)
constexpr int kTileWords = 32 * 17 + 4;      // one warp tile (32 instances) + alignment lead
#ifndef VX_STAGE_BUFS
#define VX_STAGE_BUFS 2
#endif
constexpr int kStageBufs = VX_STAGE_BUFS;   // staging tiles per warp

struct EmitSmem {
    ChunkSmem c;
    uint32_t nz_base[VX_ROWS + 1];           // first quad of each NON-EMPTY row (+ sentinel = count)
    uint16_t nz_row[VX_ROWS];                // its row index
    uint32_t nz_coarse[32];                  // nz_base[32 * i]: the coarse probe reads these at stride 1 (the
                                             // stride-32 gather from nz_base itself is a 32-way bank conflict)
    __align__(16) uint32_t lut[16 * kLutStride];   // [type*8 + face]: 12 rotation words, [16] = texture layer
    uint32_t pos[3][32];                     // f32 bits of world x / y / z per local coordinate
    uint32_t face_sel[64];                   // [6 face bits of a cube] -> its set faces in order, 3 bits each
    uint32_t warp_tot[kEmitWarps];
    __align__(16) uint32_t stage[kEmitWarps][kStageBufs][kTileWords];   // per-warp staging tiles
};

__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}

// TMA 1-D bulk store shared -> global (sm_90+): 16-B aligned addresses, size % 16 == 0
__device__ __forceinline__ void bulk_store(void *gdst, const void *ssrc, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 :: "l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}

// Emission.  After the per-chunk setup (stage + row scan) there is no CTA-wide
// barrier: each warp owns every kEmitWarps-th tile of 32 consecutive quads and
// each lane decodes ITS quad independently --
//   row   = binary search of the quad index in the row scan,
//   cube  = bit-parallel select over 3-bit vertical counters of the six masks,
//   face  = k-th set face bit of that cube,
// builds the 68-byte InstanceRaw in the warp's staging tile (stride-17 words,
// bank-conflict free) and the tile leaves through one TMA bulk store.
//
// PACKED = true writes one 32-bit quad record instead of the 68-byte instance (extension for
// renderers that expand in the vertex shader): bits 0-4 x, 5-9 z, 10-14 y (= voxel idx),
// 15-17 face (VX_LEFT..VX_UP), 18 block type (0 GRASS, 1 DIRT); the chunk is implied by offsets[].
// TYPES: where a quad's block type comes from -- 0: every block is GRASS (the reference's terrain), 1: the caller's
// DIRT plane (vx_mesh), 2: the DIRT rule of the extension (the voxel above is solid).  A template parameter so that
// the reference path carries no per-quad test for the other two.
#ifndef VX_EMIT_TILE_UNROLL
#define VX_EMIT_TILE_UNROLL 1
#endif
constexpr int kEmitTileUnroll = VX_EMIT_TILE_UNROLL;      // A/B aid: tiles per iteration of a warp's tile loop
template <bool PACKED, int kEmitSplit, int TYPES>
__global__ void __launch_bounds__(kEmitThreads)
vx_mesh_emit_kernel(const vx_mesh_args a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    EmitSmem &s = *reinterpret_cast<EmitSmem *>(smem_raw);
    // kEmitSplit CTAs share one chunk (interleaved tiles): more, shorter CTAs -> smaller tail
    const uint32_t part = blockIdx.x % kEmitSplit, n_chunks = gridDim.x / kEmitSplit;
    const uint32_t m = a.m0 + blockIdx.x / kEmitSplit, tid = threadIdx.x, lane = tid & 31u, wid = tid >> 5;

    // this launch covers chunks [m0, m0 + n_chunks); its quads land at
    // out[offsets[m] - quad_base]
    const unsigned long long launch_quads = a.offsets[a.m0 + n_chunks] - a.quad_base;
    if (launch_quads > a.capacity) {
        if (blockIdx.x == 0 && tid == 0 && a.status) atomicOr(a.status, VX_ST_OVERFLOW);
        return;
    }
    const uint32_t count = a.counts[m];
    if (count == 0) return;                      // also covers mesh_mask == 0

    // constant tables -> shared
    for (uint32_t i = tid; i < 16 * kLutStride; i += kEmitThreads) {
        const uint32_t tf = i / kLutStride, w = i - tf * kLutStride, type = tf >> 3, face = tf & 7u;
        uint32_t v = 0;
        if (face < 6) {
            if (w < 12) v = c_rot[face][w];
            else if (w == 16) v = c_tex[type][face];
        }
        s.lut[i] = v;
    }
    if (tid >= 96 && tid < 128) s.nz_coarse[tid - 96] = 0;        // entries past the last non-empty row are never used
    if (tid >= 128 && tid < 192) {
        // k-th set bit of the 6-bit face pattern fb, for all k at once: faces LEFT..UP in emission order
        const uint32_t fb = tid - 128u;
        uint32_t packed = 0, k = 0;
#pragma unroll
        for (uint32_t g = 0; g < 6; ++g)
            if ((fb >> g) & 1u) packed |= g << (3u * k++);
        s.face_sel[fb] = packed;
    }
    if (tid < 96) {
        // coordinate.rs:60-65 + chunk.rs:66,68,76 + quad.rs:42: (chunk*32 + local) as f32
        const uint32_t axis = tid >> 5, v = tid & 31u;
        const uint32_t c = (uint32_t)a.coords[3 * m + axis];
        s.pos[axis][v] = __float_as_uint((float)(int32_t)(c * 32u + v));
    }
    stage_chunk<kEmitThreads>(s.c, a, m);        // ends with __syncthreads()

    // ---- per-row quad counts -> exclusive scan in row order, compacted to the non-empty rows
    //      (one packed scan: quads in bits 0-19, non-empty rows in bits 20-30) ------------------
    uint32_t nzc;
    {
        uint32_t cnt[4], mine = 0;
        RowFaces f4[4];
        row_faces4(s.c, tid, f4);
#pragma unroll
        for (uint32_t j = 0; j < 4; ++j) {
            cnt[j] = row_count(f4[j]);
            mine += cnt[j] + (cnt[j] ? (1u << 20) : 0u);
        }
        uint32_t inc = mine;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(VX_FULL, inc, d);
            if (lane >= (uint32_t)d) inc += t;
        }
        if (lane == 31) s.warp_tot[wid] = inc;
        __syncthreads();
        uint32_t base = inc - mine, total = 0;
#pragma unroll
        for (uint32_t w = 0; w < (uint32_t)kEmitWarps; ++w) {
            if (w < wid) base += s.warp_tot[w];
            total += s.warp_tot[w];
        }
        nzc = total >> 20;
#pragma unroll
        for (uint32_t j = 0; j < 4; ++j) {
            if (cnt[j]) {
                s.nz_base[base >> 20] = base & 0xfffffu;
                if (((base >> 20) & 31u) == 0u) s.nz_coarse[base >> 25] = base & 0xfffffu;
                s.nz_row[base >> 20] = (uint16_t)(tid * 4 + j);
                base += cnt[j] + (1u << 20);
            }
        }
        if (tid == 0) s.nz_base[nzc] = count;
        __syncthreads();                         // last CTA-wide barrier
    }

    uint32_t *out32 = reinterpret_cast<uint32_t *>(a.out);
    const unsigned long long chunk_word0 = (a.offsets[m] - a.quad_base) * 17ull;   // first word of this chunk
    const bool out_aligned = (reinterpret_cast<unsigned long long>(out32) & 15ull) == 0;
    // A 32-quad tile is 544 words = 136 x 16 bytes, so every tile of the chunk starts at the same offset inside a
    // 16-byte line as the chunk itself: the alignment arithmetic is per chunk, and tile addresses are 32-bit offsets
    // from the chunk's base.  stage word i <-> chunk word (tile * 544 - lead + i).
    uint32_t *const out_chunk = out32 + chunk_word0;
    const uint32_t lead = (uint32_t)(chunk_word0 & 3ull);
    const uint32_t body0 = lead ? 4u : 0u;                  // first 16-byte aligned stage word at or after `lead`
    const uint32_t *typ_rows = a.typ ? a.typ + (size_t)m * VX_ROWS : nullptr;
    const uint32_t n_tiles = (count + 31u) >> 5;
    uint32_t it = 0;

#pragma unroll(kEmitTileUnroll)
    for (uint32_t tile = wid + part * kEmitWarps; tile < n_tiles; tile += kEmitWarps * kEmitSplit, ++it) {
        uint32_t *stg = PACKED ? nullptr : s.stage[wid][it % kStageBufs];
        if (!PACKED) {
            // the bulk store that last read this buffer (two tiles ago) must be done reading it
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read %0;" :: "n"(kStageBufs - 1) : "memory");
            __syncwarp();
        }

        const uint32_t q0 = tile << 5;
        const uint32_t nq = min(32u, count - q0);
        // Row of every quad of the tile, cooperatively: idx0 = the non-empty row holding q0 (two
        // 32-way ballots over nz_base), then each later row start inside the tile sets one bit.
        const uint32_t c_probe = s.nz_coarse[lane];
        const uint32_t cblk = __popc(__ballot_sync(VX_FULL, lane * 32u < nzc && c_probe <= q0)) - 1u;
        const uint32_t f_probe = s.nz_base[min(cblk * 32u + lane, nzc)];
        const uint32_t idx0 = cblk * 32u + __popc(__ballot_sync(VX_FULL, cblk * 32u + lane < nzc && f_probe <= q0)) - 1u;
        const uint32_t nxt = s.nz_base[min(idx0 + 1u + lane, nzc)] - q0;       // >= 1
        const uint32_t starts = __reduce_or_sync(VX_FULL, (idx0 + 1u + lane < nzc && nxt < 32u) ? (1u << nxt) : 0u);
        if (lane < nq) {
            const uint32_t q = q0 + lane;
            const uint32_t idx = idx0 + __popc(starts & ((2u << lane) - 1u));
            const uint32_t r = s.nz_row[idx];
            uint32_t k = q - s.nz_base[idx];
            const RowFaces f = row_faces(s.c, r);
            // 3-bit vertical counters: per x, c(x) = number of faces = s0 + 2 s1 + 4 s2
            const uint32_t a0 = f.m[0] ^ f.m[1] ^ f.m[2], a1 = (f.m[0] & f.m[1]) | (f.m[2] & (f.m[0] ^ f.m[1]));
            const uint32_t b0 = f.m[3] ^ f.m[4] ^ f.m[5], b1 = (f.m[3] & f.m[4]) | (f.m[5] & (f.m[3] ^ f.m[4]));
            const uint32_t s0 = a0 ^ b0, c0 = a0 & b0;
            const uint32_t s1 = a1 ^ b1 ^ c0, s2 = (a1 & b1) | (c0 & (a1 ^ b1));
            // cube: largest x with (# quads of cubes < x) <= k
            uint32_t x = 0, before = 0;
#pragma unroll
            for (uint32_t step = 16; step > 0; step >>= 1) {
                const uint32_t lt = (1u << (x + step)) - 1u;
                const uint32_t c = __popc(s0 & lt) + 2u * __popc(s1 & lt) + 4u * __popc(s2 & lt);
                if (c <= k) { x += step; before = c; }
            }
            k -= before;
            // face: the k-th set face bit of cube x, in LEFT,RIGHT,BACK,FRONT,DOWN,UP order (table look-up)
            uint32_t fb = 0;
#pragma unroll
            for (uint32_t g = 0; g < 6; ++g) fb |= ((f.m[g] >> x) & 1u) << g;
            const uint32_t face = (s.face_sel[fb] >> (3u * k)) & 7u;
            const uint32_t type = TYPES == 1 ? ((__ldg(typ_rows + r) >> x) & 1u)
                                : TYPES == 2 ? ((s.c.occ[32 + r + 32] >> x) & 1u) : 0u;
            if (PACKED) {
                // 128 B per warp, coalesced
                out32[(a.offsets[m] - a.quad_base) + q] = (r << 5) | x | (face << 15) | (type << 18);
                continue;
            }
            const uint32_t tf = type * 8u + face;
            const uint4 *L = reinterpret_cast<const uint4 *>(&s.lut[tf * kLutStride]);
            const uint4 c0v = L[0], c1v = L[1], c2v = L[2];
            uint32_t *w = stg + lead + lane * 17u;             // stride 17 words: conflict-free
            w[0] = c0v.x; w[1] = c0v.y; w[2] = c0v.z; w[3] = c0v.w;
            w[4] = c1v.x; w[5] = c1v.y; w[6] = c1v.z; w[7] = c1v.w;
            w[8] = c2v.x; w[9] = c2v.y; w[10] = c2v.z; w[11] = c2v.w;
            w[12] = s.pos[0][x];
            w[13] = s.pos[1][r >> 5];
            w[14] = s.pos[2][r & 31u];
            w[15] = VX_1;
            w[16] = s.lut[tf * kLutStride + 16];
        }
        if (PACKED) continue;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic writes -> async proxy
        __syncwarp();
        // Stage words [lead, lead + nwords) hold the tile.  The 16-byte aligned body [body0, body1) leaves as one bulk
        // store; the <= 3-word head and tail (or everything, if the caller's buffer is not 16-byte aligned) as
        // scalar stores.
        const uint32_t end = lead + nq * 17u, body1 = end & ~3u;
        uint32_t *const gdst = out_chunk + ((ptrdiff_t)(tile * 544u) - (ptrdiff_t)lead);   // stage word i -> gdst[i]
        if (out_aligned && body1 > body0) {
            if (lane == 0) bulk_store(gdst + body0, stg + body0, (body1 - body0) * 4u);
            // at most 3 head + 3 tail words: one predicated store
            const uint32_t head = body0 - lead, n_edge = head + (end - body1);
            if (lane < n_edge) {
                const uint32_t wi = (lane < head) ? lead + lane : body1 + (lane - head);
                gdst[wi] = stg[wi];
            }
        } else {
            for (uint32_t i = lead + lane; i < end; i += 32u) gdst[i] = stg[i];
        }
    }
    if (!PACKED && lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // smem must outlive the reads
}

}  // namespace

// ---------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------
cudaError_t vx_launch_pack(const uint8_t *blocks, uint32_t n, uint32_t *occ, uint32_t *typ,
                           uint32_t *status, cudaStream_t s, vx_counter *launches)
{
    if (n == 0) return cudaSuccess;
    const size_t total = (size_t)n * (VX_VOXELS / 16);
    vx_pack_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(blocks, n, occ, typ, status);
    if (launches) launches->fetch_add(1, std::memory_order_relaxed);
    return cudaGetLastError();
}

cudaError_t vx_launch_resolve(const vx_resolve_args &a, cudaStream_t s, vx_counter *launches)
{
    if (a.n == 0 || a.hi <= a.lo) return cudaSuccess;
    if (a.n <= kSmallResolve) {
        vx_resolve_args b = a;
        b.table_size = 64;
        while (b.table_size < 2 * a.n) b.table_size <<= 1;
        const size_t smem = (size_t)b.table_size * 4 + (size_t)a.n * 12;      // <= 32 KB + 48 KB (vx_init_mesh_kernels)
        vx_resolve_small_kernel<<<1, 1024, smem, s>>>(b);
        if (launches) launches->fetch_add(1, std::memory_order_relaxed);
        return cudaGetLastError();
    }
    cudaError_t e = cudaMemsetAsync(a.table, 0xff, sizeof(int32_t) * a.table_size, s);
    if (e != cudaSuccess) return e;
    if (a.halo_count) {
        e = cudaMemsetAsync(a.halo_count, 0, sizeof(uint32_t), s);
        if (e != cudaSuccess) return e;
    }
    vx_ht_insert_kernel<<<(a.n + 255) / 256, 256, 0, s>>>(a.coords, a.n, a.table, a.table_size - 1);
    vx_resolve_kernel<<<((a.hi - a.lo) * (a.nbr_y ? 6 : 4) + 255) / 256, 256, 0, s>>>(a);
    if (launches) launches->fetch_add(2, std::memory_order_relaxed);
    return cudaGetLastError();
}

cudaError_t vx_launch_remap_nbr(const int32_t *nbr_in, int32_t *nbr_out, uint32_t n, uint32_t *status, cudaStream_t s,
                                vx_counter *launches)
{
    if (n == 0) return cudaSuccess;
    vx_remap_nbr_kernel<<<(n * 4 + 255) / 256, 256, 0, s>>>(nbr_in, nbr_out, n * 4, n, status);
    if (launches) launches->fetch_add(1, std::memory_order_relaxed);
    return cudaGetLastError();
}

cudaError_t vx_launch_blocks_ex(const vx_mesh_args &a, uint8_t *blocks, cudaStream_t s, vx_counter *launches)
{
    if (a.n == 0) return cudaSuccess;
    vx_blocks_ex_kernel<<<a.n, kMeshThreads, 0, s>>>(a, blocks);
    if (launches) launches->fetch_add(1, std::memory_order_relaxed);
    return cudaGetLastError();
}

cudaError_t vx_launch_mesh_count(const vx_mesh_args &a, cudaStream_t s, vx_counter *launches)
{
    if (a.n == 0) return cudaSuccess;
    vx_mesh_count_kernel<<<a.n, kMeshThreads, 0, s>>>(a);
    if (launches) launches->fetch_add(1, std::memory_order_relaxed);
    return cudaGetLastError();
}

template <bool PACKED, int SPLIT>
static void launch_emit_types(const vx_mesh_args &a, uint32_t len, size_t smem, cudaStream_t s)
{
    if (a.typ) vx_mesh_emit_kernel<PACKED, SPLIT, 1><<<len * SPLIT, kEmitThreads, smem, s>>>(a);
    else if (a.dirt_rule) vx_mesh_emit_kernel<PACKED, SPLIT, 2><<<len * SPLIT, kEmitThreads, smem, s>>>(a);
    else vx_mesh_emit_kernel<PACKED, SPLIT, 0><<<len * SPLIT, kEmitThreads, smem, s>>>(a);
}

cudaError_t vx_launch_mesh_emit(const vx_mesh_args &a, uint32_t len, cudaStream_t s, vx_counter *launches)
{
    if (len == 0) return cudaSuccess;
    static_assert(sizeof(EmitSmem) <= 64 * 1024, "emit kernel shared memory budget");
    if (len <= kEmitSmallList) {
        if (a.packed) launch_emit_types<true, kEmitSplitSmall>(a, len, offsetof(EmitSmem, stage), s);
        else launch_emit_types<false, kEmitSplitSmall>(a, len, sizeof(EmitSmem), s);
    } else {
        if (a.packed) launch_emit_types<true, 1>(a, len, offsetof(EmitSmem, stage), s);
        else launch_emit_types<false, 1>(a, len, sizeof(EmitSmem), s);
    }
    if (launches) launches->fetch_add(1, std::memory_order_relaxed);
    return cudaGetLastError();
}

// Per-device kernel attributes (dynamic shared memory above 48 KB); called by vx_create with
// the device current.
template <int SPLIT, int TYPES>
static cudaError_t allow_emit_smem()
{
    return cudaFuncSetAttribute(vx_mesh_emit_kernel<false, SPLIT, TYPES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)sizeof(EmitSmem));
}

cudaError_t vx_init_mesh_kernels()
{
    cudaError_t e = allow_emit_smem<1, 0>();
    if (e == cudaSuccess) e = allow_emit_smem<1, 1>();
    if (e == cudaSuccess) e = allow_emit_smem<1, 2>();
    if (e == cudaSuccess) e = allow_emit_smem<kEmitSplitSmall, 0>();
    if (e == cudaSuccess) e = allow_emit_smem<kEmitSplitSmall, 1>();
    if (e == cudaSuccess) e = allow_emit_smem<kEmitSplitSmall, 2>();
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(vx_resolve_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)(2 * kSmallResolve * 4 + kSmallResolve * 12));
    return e;
}
