// kernels_frustum.cu -- Frustum::contains (frustum_culling.rs:16-40) for
// n chunk AABBs x F plane sets per launch.
//
// One lane owns one chunk and keeps its six AABB coordinates in registers; a
// CTA covers 256 chunks x up to kFramesPerCta frames whose planes are staged
// in shared memory (warp-uniform broadcast reads).  The distance is evaluated
// exactly as the reference does, ((a*x + b*y) + c*z) + d with every product
// and sum rounded separately (frustum_culling.rs:29-32), and a corner is
// rejected iff dist < 0 (NaN passes, :33-35).  The result is "any corner
// inside all six planes" (:37) -- a pure function, so planes may be skipped
// once no lane of the warp has a live corner.
#include "vx_common.cuh"

namespace {

// One plane against the eight corners of one AABB.  The reference evaluates ((a*x + b*y) + c*z) + d with every
// product and sum rounded separately (frustum_culling.rs:29-32); the six products are shared by the corners, and
// the 20 additions run as 10 packed FADD2 (sm_100: two IEEE-rn sums per issue slot) on the (x0, x1) corner pairs.
// Every product stays a SCALAR __fmul_rn: ptxas contracts a packed product into a packed add that consumes it
// (FFMA2) even where contraction is off, and does not fuse across the scalar / packed boundary.
//   m[0..3] = running minima of the pairs (000,100) (001,101) (010,110) (011,111)   [xyz bits as in :19-26]
__device__ __forceinline__ void plane_dist(const float4 pl, float x0, float x1, float y0, float y1, float z0, float z1,
                                           float2 (&t)[4])
{
    const float2 ax = make_float2(__fmul_rn(pl.x, x0), __fmul_rn(pl.x, x1));
    const float by0 = __fmul_rn(pl.y, y0), by1 = __fmul_rn(pl.y, y1);
    const float cz0 = __fmul_rn(pl.z, z0), cz1 = __fmul_rn(pl.z, z1);
    const float2 s0 = __fadd2_rn(ax, make_float2(by0, by0)), s1 = __fadd2_rn(ax, make_float2(by1, by1));
    const float2 d = make_float2(pl.w, pl.w);
    const float2 t00 = __fadd2_rn(__fadd2_rn(s0, make_float2(cz0, cz0)), d);      // corners (x, y0, z0)
    const float2 t01 = __fadd2_rn(__fadd2_rn(s0, make_float2(cz1, cz1)), d);      // (x, y0, z1)
    const float2 t10 = __fadd2_rn(__fadd2_rn(s1, make_float2(cz0, cz0)), d);      // (x, y1, z0)
    const float2 t11 = __fadd2_rn(__fadd2_rn(s1, make_float2(cz1, cz1)), d);      // (x, y1, z1)
    t[0] = t00; t[1] = t01; t[2] = t10; t[3] = t11;
}
// two planes per step: min(min(m, a), b) is one three-input FMNMX3 on sm_100.  fminf ignores a NaN operand, which
// is the reference's behaviour (a NaN distance never rejects, :33)
__device__ __forceinline__ void plane_pair_step(const float4 pa, const float4 pb, float x0, float x1, float y0, float y1,
                                                float z0, float z1, float2 (&m)[4])
{
    float2 ta[4], tb[4];
    plane_dist(pa, x0, x1, y0, y1, z0, z1, ta);
    plane_dist(pb, x0, x1, y0, y1, z0, z1, tb);
#pragma unroll
    for (int c = 0; c < 4; ++c)
        m[c] = make_float2(fminf(fminf(m[c].x, ta[c].x), tb[c].x), fminf(fminf(m[c].y, ta[c].y), tb[c].y));
}
__device__ __forceinline__ float best_corner(const float2 (&m)[4])
{
    return fmaxf(fmaxf(fmaxf(m[0].x, m[0].y), fmaxf(m[1].x, m[1].y)), fmaxf(fmaxf(m[2].x, m[2].y), fmaxf(m[3].x, m[3].y)));
}

constexpr int kThreads = 256;
constexpr int kFramesPerCta = 64;

template <bool WANT_BYTES, bool WANT_BITS>
__global__ void __launch_bounds__(kThreads)
vx_frustum_kernel(const float *__restrict__ planes, uint32_t n_frames,
                  const int32_t *__restrict__ coords, uint32_t n,
                  uint8_t *__restrict__ visible, uint32_t *__restrict__ bits, uint32_t n_words)
{
    __shared__ float4 s_pl[kFramesPerCta * 6];
    const uint32_t f0 = blockIdx.y * kFramesPerCta;
    const uint32_t nf = min((uint32_t)kFramesPerCta, n_frames - f0);
    for (uint32_t i = threadIdx.x; i < nf * 6; i += kThreads)
        s_pl[i] = __ldg(reinterpret_cast<const float4 *>(planes) + (size_t)f0 * 6 + i);
    __syncthreads();

    const uint32_t i = blockIdx.x * kThreads + threadIdx.x;
    const bool live = i < n;
    // coordinate.rs:53-58: (chunk * 32) as f32 (wrapping i32 product);
    // frustum_culling.rs:19-26: the far corners are p + 32.
    float x0 = 0.f, y0 = 0.f, z0 = 0.f;
    if (live) {
        x0 = (float)(int32_t)((uint32_t)__ldg(coords + 3 * (size_t)i + 0) * 32u);
        y0 = (float)(int32_t)((uint32_t)__ldg(coords + 3 * (size_t)i + 1) * 32u);
        z0 = (float)(int32_t)((uint32_t)__ldg(coords + 3 * (size_t)i + 2) * 32u);
    }
    const float x1 = __fadd_rn(x0, 32.0f), y1 = __fadd_rn(y0, 32.0f), z1 = __fadd_rn(z0, 32.0f);

    for (uint32_t f = 0; f < nf; ++f) {
        // Per corner, the minimum over the planes of its distance: the corner is inside all six
        // planes iff that minimum is not < 0.  fminf ignores a NaN operand, which is the
        // reference's behaviour (a NaN distance never rejects, frustum_culling.rs:33).
        const float inf = __int_as_float(0x7f800000);
        float2 m[4] = {make_float2(inf, inf), make_float2(inf, inf), make_float2(inf, inf), make_float2(inf, inf)};
        bool dead = false;
#pragma unroll
        for (uint32_t p = 1; p < 6; p += 2) {
            plane_pair_step(s_pl[f * 6 + p - 1], s_pl[f * 6 + p], x0, x1, y0, y1, z0, z1, m);
            if (p == 1 || p == 3) {
                // the result is a pure function, so the remaining planes can be skipped once no
                // lane of the warp has a live corner
                if (!__any_sync(VX_FULL, live && !(best_corner(m) < 0.0f))) {
                    dead = true;
                    break;
                }
            }
        }
        const bool vis = live && !dead && !(best_corner(m) < 0.0f);
        const uint32_t word = __ballot_sync(VX_FULL, vis);
        if (WANT_BYTES && live) visible[(size_t)(f0 + f) * n + i] = vis ? 1 : 0;
        if (WANT_BITS && (threadIdx.x & 31) == 0 && live) bits[(size_t)(f0 + f) * n_words + (i >> 5)] = word;
    }
}

// ---------------------------------------------------------------------------
// One plane set per launch (the reference's own call shape: Frustum::contains against the current
// camera, chunk_builder.rs:251 / world.rs:35, for every loaded chunk).  13 bytes per test, so the launch
// is latency-bound: there is no shared-memory staging and no CTA barrier.  A thread owns FOUR consecutive
// chunks whose 12 coordinates arrive as three 128-bit loads (a warp reads 1536 contiguous bytes); the
// six planes are six uniform 128-bit loads issued beside them.  Same arithmetic and the same warp-level
// plane skipping as the general kernel.
// ---------------------------------------------------------------------------
__device__ __forceinline__ bool aabb_visible(const float4 (&pl)[6], int32_t cx, int32_t cy, int32_t cz, bool live)
{
    const float x0 = (float)(int32_t)((uint32_t)cx * 32u), y0 = (float)(int32_t)((uint32_t)cy * 32u),
                z0 = (float)(int32_t)((uint32_t)cz * 32u);
    const float x1 = __fadd_rn(x0, 32.0f), y1 = __fadd_rn(y0, 32.0f), z1 = __fadd_rn(z0, 32.0f);
    const float inf = __int_as_float(0x7f800000);
    float2 m[4] = {make_float2(inf, inf), make_float2(inf, inf), make_float2(inf, inf), make_float2(inf, inf)};
#pragma unroll
    for (uint32_t p = 1; p < 6; p += 2) {
        plane_pair_step(pl[p - 1], pl[p], x0, x1, y0, y1, z0, z1, m);
        if (p == 1 || p == 3) {
            if (!__any_sync(VX_FULL, live && !(best_corner(m) < 0.0f))) return false;
        }
    }
    return live && !(best_corner(m) < 0.0f);
}

// G groups of four chunks per thread (group g of round r = r * threads + thread: every round is a
// coalesced sweep).  The launcher picks G so that the whole grid is resident at once -- one wave, all
// loads of a thread issued up front.
template <bool WANT_BYTES, bool WANT_BITS, int G>
__global__ void __launch_bounds__(kThreads)
vx_frustum_one_kernel(const float *__restrict__ planes, const int32_t *__restrict__ coords, uint32_t n,
                      uint8_t *__restrict__ visible, uint32_t *__restrict__ bits)
{
    const uint32_t lane = threadIdx.x & 31u, stride = gridDim.x * kThreads;
    const uint32_t g0 = blockIdx.x * kThreads + threadIdx.x;     // first group of four chunks
    int32_t c[G][12];
#pragma unroll
    for (int r = 0; r < G; ++r) {
        const uint32_t g = g0 + r * stride, i0 = g * 4u;
        if (i0 + 4u <= n) {
            const uint4 *p = reinterpret_cast<const uint4 *>(coords) + (size_t)g * 3;
            const uint4 a = __ldg(p), b = __ldg(p + 1), d = __ldg(p + 2);
            c[r][0] = a.x; c[r][1] = a.y; c[r][2] = a.z; c[r][3] = a.w; c[r][4] = b.x; c[r][5] = b.y;
            c[r][6] = b.z; c[r][7] = b.w; c[r][8] = d.x; c[r][9] = d.y; c[r][10] = d.z; c[r][11] = d.w;
        } else {
#pragma unroll
            for (uint32_t k = 0; k < 12; ++k) c[r][k] = (i0 + k / 3u < n) ? __ldg(coords + 3 * (size_t)i0 + k) : 0;
        }
    }
    float4 pl[6];
#pragma unroll
    for (uint32_t p = 0; p < 6; ++p) pl[p] = __ldg(reinterpret_cast<const float4 *>(planes) + p);
#pragma unroll
    for (int r = 0; r < G; ++r) {
        const uint32_t g = g0 + r * stride, i0 = g * 4u;
        if ((g & ~31u) * 4u >= n) break;                         // whole warp past the end (warp-uniform)
        uint32_t nib = 0;
#pragma unroll
        for (uint32_t k = 0; k < 4; ++k)
            nib |= (aabb_visible(pl, c[r][3 * k], c[r][3 * k + 1], c[r][3 * k + 2], i0 + k < n) ? 1u : 0u) << k;
        if (WANT_BYTES && i0 < n) {
            const uint32_t w = (nib & 1u) | ((nib & 2u) << 7) | ((nib & 4u) << 14) | ((nib & 8u) << 21);
            if (i0 + 4u <= n) *reinterpret_cast<uint32_t *>(visible + i0) = w;       // visible is 4-byte aligned (launcher)
            else for (uint32_t k = 0; i0 + k < n; ++k) visible[i0 + k] = (uint8_t)((nib >> k) & 1u);
        }
        if (WANT_BITS) {
            // eight lanes hold one 32-chunk word
            uint32_t v = nib << (4u * (lane & 7u));
            v |= __shfl_xor_sync(VX_FULL, v, 1);
            v |= __shfl_xor_sync(VX_FULL, v, 2);
            v |= __shfl_xor_sync(VX_FULL, v, 4);
            if ((lane & 7u) == 0u && i0 < n) bits[i0 >> 5] = v;
        }
    }
}

template <int G>
void launch_one(const float *planes, const int32_t *coords, uint32_t n, uint8_t *visible, uint32_t *bits, cudaStream_t s)
{
    const uint32_t groups = (n + 3u) / 4u, threads = (groups + G - 1) / G, g1 = (threads + kThreads - 1) / kThreads;
    if (visible && bits) vx_frustum_one_kernel<true, true, G><<<g1, kThreads, 0, s>>>(planes, coords, n, visible, bits);
    else if (visible) vx_frustum_one_kernel<true, false, G><<<g1, kThreads, 0, s>>>(planes, coords, n, visible, bits);
    else vx_frustum_one_kernel<false, true, G><<<g1, kThreads, 0, s>>>(planes, coords, n, visible, bits);
}

}  // namespace

cudaError_t vx_launch_frustum(const float *planes, uint32_t n_frames, const int32_t *coords, uint32_t n,
                              uint8_t *visible, uint32_t *bits, cudaStream_t s, vx_counter *launches)
{
    if (n == 0 || n_frames == 0) return cudaSuccess;
    if (n_frames == 1 && (reinterpret_cast<uintptr_t>(coords) & 15u) == 0 && (reinterpret_cast<uintptr_t>(planes) & 15u) == 0 &&
        (!visible || (reinterpret_cast<uintptr_t>(visible) & 3u) == 0)) {
        // one wave if possible: 148 SMs x 4 resident CTAs x 256 threads x 4 chunks = 606 k chunks with G = 1
        if (n <= 148u * 4u * kThreads * 4u) launch_one<1>(planes, coords, n, visible, bits, s);
        else launch_one<2>(planes, coords, n, visible, bits, s);
        if (launches) launches->fetch_add(1, std::memory_order_relaxed);
        return cudaGetLastError();
    }
    const dim3 grid((n + kThreads - 1) / kThreads, (n_frames + kFramesPerCta - 1) / kFramesPerCta);
    const uint32_t n_words = (n + 31) / 32;
    if (visible && bits)
        vx_frustum_kernel<true, true><<<grid, kThreads, 0, s>>>(planes, n_frames, coords, n, visible, bits, n_words);
    else if (visible)
        vx_frustum_kernel<true, false><<<grid, kThreads, 0, s>>>(planes, n_frames, coords, n, visible, bits, n_words);
    else
        vx_frustum_kernel<false, true><<<grid, kThreads, 0, s>>>(planes, n_frames, coords, n, visible, bits, n_words);
    if (launches) launches->fetch_add(1, std::memory_order_relaxed);
    return cudaGetLastError();
}
