// kernels_frustum.cu -- Frustum::contains (frustum_culling.rs:16-40) for
// n chunk AABBs x F plane sets per launch.
//
// The reference evaluates, per corner and plane, ((a*x + b*y) + c*z) + d with every product and sum rounded
// separately (:29-32), rejects a corner iff dist < 0 (a NaN distance passes, :33-35) and answers "some corner is
// inside all six planes" (:37).  Two kernels produce exactly that answer:
//   vx_frustum_kernel      n chunks x up to 64 plane sets per CTA, planes and their sign classes in shared memory,
//                          a lane carries four chunks (their 24 corner coordinates) through all frames;
//   vx_frustum_one_kernel  one plane set per launch (the engine's call shape), four consecutive chunks per thread.
// Both decide a chunk from the TWO corners per plane that the coefficient signs pick (the distance is a monotone
// function of each coordinate, see "Extreme-corner classification" below) and run the reference's corner-by-corner
// test only on chunks that straddle two planes or more -- or on every chunk when a coefficient is NaN, infinite or
// larger than 2^90, where the ordering argument does not hold.  Planes are skipped once every chunk of a warp is
// rejected; the side planes go first because they reject the most.
#include "vx_common.cuh"

namespace {

#include "vx_frustum.cuh"

constexpr int kThreads = 256;
constexpr int kFramesPerCta = 64;
// chunk AABBs a lane tests together: the per-plane bookkeeping (plane load, class dispatch, loop) is paid once for all
#ifndef VX_FR_FRAME_BOXES
#define VX_FR_FRAME_BOXES 4
#endif
#ifndef VX_FR_FRAME_UNROLL
#define VX_FR_FRAME_UNROLL 6
#endif
#ifndef VX_FR_ONE_UNROLL
#define VX_FR_ONE_UNROLL 6
#endif
#ifndef VX_FR_ONE_THREADS
#define VX_FR_ONE_THREADS 128
#endif
constexpr int kFrameBoxes = VX_FR_FRAME_BOXES;       // general kernel
// planes per iteration of the class-dispatch loop (both kernels measured fastest with the loop unrolled completely:
// 334 -> 378 G tests/s at 1000 frames per launch, 5.7 -> 5.3 us for one frame)
constexpr int kFrameUnroll = VX_FR_FRAME_UNROLL, kOneUnroll = VX_FR_ONE_UNROLL;
constexpr int kOneThreads = VX_FR_ONE_THREADS;       // CTA size of the single-frame kernel

// n chunks x up to kFramesPerCta plane sets per CTA.  A warp covers 32 * kFrameBoxes consecutive chunks, lane l the
// chunks base + 32 k + l, so every store of a frame is a coalesced row and a ballot is one word of the bit output.
template <bool WANT_BYTES, bool WANT_BITS>
__global__ void __launch_bounds__(kThreads)
vx_frustum_kernel(const float *__restrict__ planes, uint32_t n_frames,
                  const int32_t *__restrict__ coords, uint32_t n,
                  uint8_t *__restrict__ visible, uint32_t *__restrict__ bits, uint32_t n_words)
{
    __shared__ float4 s_pl[kFramesPerCta * 6];
    __shared__ uint32_t s_cls[kFramesPerCta];
    const uint32_t f0 = blockIdx.y * kFramesPerCta;
    const uint32_t nf = min((uint32_t)kFramesPerCta, n_frames - f0);
    const float4 *gp = reinterpret_cast<const float4 *>(planes) + (size_t)f0 * 6;
    for (uint32_t i = threadIdx.x; i < nf * 6; i += kThreads) s_pl[i] = __ldg(gp + i);
    if (threadIdx.x < nf) {
        uint32_t cls = 0;
#pragma unroll
        for (uint32_t p = 0; p < 6; ++p) cls |= plane_class(__ldg(gp + threadIdx.x * 6 + p)) << (4u * p);
        s_cls[threadIdx.x] = cls;
    }
    __syncthreads();

    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t wbase = (blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5)) * (32u * kFrameBoxes);
    if (wbase >= n) return;
    Boxes<kFrameBoxes> b;
    uint32_t live = 0;
#pragma unroll
    for (int k = 0; k < kFrameBoxes; ++k) {
        const uint32_t i = wbase + 32u * k + lane;
        int32_t cx = 0, cy = 0, cz = 0;
        if (i < n) {
            cx = __ldg(coords + 3 * (size_t)i + 0);
            cy = __ldg(coords + 3 * (size_t)i + 1);
            cz = __ldg(coords + 3 * (size_t)i + 2);
            live |= 1u << k;
        }
        set_box(b, k, cx, cy, cz);
    }

    for (uint32_t f = 0; f < nf; ++f) {
        const uint32_t vis = frame_test<kFrameUnroll>(s_pl + f * 6, s_cls[f], b, live);
        uint32_t mine = 0;
#pragma unroll
        for (int k = 0; k < kFrameBoxes; ++k) {
            const uint32_t word = __ballot_sync(VX_FULL, (vis >> k) & 1u);
            if (lane == (uint32_t)k) mine = word;
            if (WANT_BYTES && ((live >> k) & 1u)) visible[(size_t)(f0 + f) * n + wbase + 32u * k + lane] = (vis >> k) & 1u;
        }
        if (WANT_BITS && lane < (uint32_t)kFrameBoxes && wbase + 32u * lane < n)
            bits[(size_t)(f0 + f) * n_words + (wbase >> 5) + lane] = mine;
    }
}

// ---------------------------------------------------------------------------
// One plane set per launch (the reference's own call shape: Frustum::contains against the current
// camera, chunk_builder.rs:251 / world.rs:35, for every loaded chunk).  13 bytes per test, so the launch
// is latency-bound.  A thread owns FOUR consecutive chunks whose 12 coordinates arrive as three 128-bit
// loads (a warp reads 1536 contiguous bytes), issued before anything else; warp 0 stages the six planes
// and their classes in shared memory meanwhile.  Same classification as the general kernel.  Measured on
// 10^6 AABBs (tools/probes/frustum_ab.py): small CTAs and one group per thread (1.65 waves that balance
// over the SMs) beat two groups per thread in one wave, 5.3 against 6.3 us.
// ---------------------------------------------------------------------------
template <bool WANT_BYTES, bool WANT_BITS>
__global__ void __launch_bounds__(kOneThreads)
vx_frustum_one_kernel(const float *__restrict__ planes, const int32_t *__restrict__ coords, uint32_t n,
                      uint8_t *__restrict__ visible, uint32_t *__restrict__ bits)
{
    __shared__ float4 s_pl[6];
    __shared__ uint32_t s_cls;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t g = blockIdx.x * kOneThreads + threadIdx.x, i0 = g * 4u;     // group of four chunks
    int32_t c[12];
    if (i0 + 4u <= n) {
        const uint4 *p = reinterpret_cast<const uint4 *>(coords) + (size_t)g * 3;
        const uint4 a = __ldg(p), b = __ldg(p + 1), d = __ldg(p + 2);
        c[0] = a.x; c[1] = a.y; c[2] = a.z; c[3] = a.w; c[4] = b.x; c[5] = b.y;
        c[6] = b.z; c[7] = b.w; c[8] = d.x; c[9] = d.y; c[10] = d.z; c[11] = d.w;
    } else {
#pragma unroll
        for (uint32_t k = 0; k < 12; ++k) c[k] = (i0 + k / 3u < n) ? __ldg(coords + 3 * (size_t)i0 + k) : 0;
    }
    if (threadIdx.x < 32u) {
        float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
        if (lane < 6u) {
            q = __ldg(reinterpret_cast<const float4 *>(planes) + lane);
            s_pl[lane] = q;
        }
        const uint32_t cls = __reduce_or_sync(VX_FULL, lane < 6u ? plane_class(q) << (4u * lane) : 0u);
        if (lane == 0u) s_cls = cls;
    }
    __syncthreads();
    if ((g & ~31u) * 4u >= n) return;                             // whole warp past the end (warp-uniform)
    Boxes<4> b;
    uint32_t live = 0;
#pragma unroll
    for (uint32_t k = 0; k < 4; ++k) {
        set_box(b, k, c[3 * k], c[3 * k + 1], c[3 * k + 2]);
        live |= (i0 + k < n ? 1u : 0u) << k;
    }
    const uint32_t nib = frame_test<kOneUnroll>(s_pl, s_cls, b, live);
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

void launch_one(const float *planes, const int32_t *coords, uint32_t n, uint8_t *visible, uint32_t *bits, cudaStream_t s)
{
    const uint32_t groups = (n + 3u) / 4u, g1 = (groups + kOneThreads - 1) / kOneThreads;
    if (visible && bits) vx_frustum_one_kernel<true, true><<<g1, kOneThreads, 0, s>>>(planes, coords, n, visible, bits);
    else if (visible) vx_frustum_one_kernel<true, false><<<g1, kOneThreads, 0, s>>>(planes, coords, n, visible, bits);
    else vx_frustum_one_kernel<false, true><<<g1, kOneThreads, 0, s>>>(planes, coords, n, visible, bits);
}

}  // namespace

cudaError_t vx_launch_frustum(const float *planes, uint32_t n_frames, const int32_t *coords, uint32_t n,
                              uint8_t *visible, uint32_t *bits, cudaStream_t s, vx_counter *launches)
{
    if (n == 0 || n_frames == 0) return cudaSuccess;
    if (n_frames == 1 && (reinterpret_cast<uintptr_t>(coords) & 15u) == 0 && (reinterpret_cast<uintptr_t>(planes) & 15u) == 0 &&
        (!visible || (reinterpret_cast<uintptr_t>(visible) & 3u) == 0)) {
        launch_one(planes, coords, n, visible, bits, s);
        if (launches) launches->fetch_add(1, std::memory_order_relaxed);
        return cudaGetLastError();
    }
    const uint32_t per_cta = kThreads * kFrameBoxes;
    const dim3 grid((n + per_cta - 1) / per_cta, (n_frames + kFramesPerCta - 1) / kFramesPerCta);
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
