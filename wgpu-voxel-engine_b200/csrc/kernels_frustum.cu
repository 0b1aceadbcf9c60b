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
__device__ __forceinline__ float best_corner(const float2 (&m)[4])
{
    return fmaxf(fmaxf(fmaxf(m[0].x, m[0].y), fmaxf(m[1].x, m[1].y)), fmaxf(fmaxf(m[2].x, m[2].y), fmaxf(m[3].x, m[3].y)));
}

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
// evaluate the side planes (stored third and fourth, frustum_culling.rs:84) before near / far: the result does not
// depend on the order, and in an open world the side planes reject the most chunks, so more warps are decided after
// the first pair (378 -> 406 G tests/s on the camera sweep of BASELINE configs[4])
#ifndef VX_FR_SIDES_FIRST
#define VX_FR_SIDES_FIRST 1
#endif
constexpr int kFrameBoxes = VX_FR_FRAME_BOXES;       // general kernel
// planes per iteration of the class-dispatch loop (both kernels measured fastest with the loop unrolled completely:
// 334 -> 378 G tests/s at 1000 frames per launch, 5.7 -> 5.3 us for one frame)
constexpr int kFrameUnroll = VX_FR_FRAME_UNROLL, kOneUnroll = VX_FR_ONE_UNROLL;
constexpr int kOneThreads = VX_FR_ONE_THREADS;       // CTA size of the single-frame kernel

// ---------------------------------------------------------------------------
// Extreme-corner classification.  With finite coefficients the distance
//   dist(x, y, z) = fl(fl(fl(fl(a*x) + fl(b*y)) + fl(c*z)) + d)            (frustum_culling.rs:29-32)
// is a monotone function of each coordinate: IEEE round-to-nearest is monotone, so x <= x' gives fl(a*x) <= fl(a*x')
// for a >= 0 (the reverse for a < 0), and p <= p' gives fl(p + q) <= fl(p' + q).  Over the eight corners of an AABB
// the smallest distance to a plane is therefore the one at the corner chosen by the coefficient signs (x0 where
// a >= 0, x1 where a < 0, ...: the "n corner") and the largest the one at the opposite ("p") corner -- of the very
// same rounded expression the reference evaluates, not of its real-number idealisation.  Hence
//   * some plane has dist(p corner) < 0    =>  that plane rejects all eight corners        =>  contains() = false,
//   * no plane has dist(n corner) < 0      =>  every corner is inside every plane           =>  contains() = true,
//   * exactly one plane k has dist(n corner) < 0 (and dist(p corner) >= 0): every corner is inside the other five
//     planes, and the p corner of plane k is a corner of the AABB that is inside plane k too  =>  contains() = true,
// and only an AABB that straddles two or more planes (the edges and the apex of the frustum) needs the reference's
// corner-by-corner test (:27-38), which it then gets unchanged.  Two corners per plane instead of eight: 6 products,
// 3 packed additions on the (n, p) pair, one minimum and one sign count.  (A distance of -0 is counted as a
// straddle although it is not < 0: that only sends the chunk to the corner test more often than needed.)
// The argument needs sums free of NaN and of inf - inf: coordinates are at most 2^31 in magnitude, so coefficients
// up to 2^90 keep every product and partial sum finite (< 2^124); a plane set with a larger, infinite or NaN
// coefficient (plane_class bit 3) takes the corner-by-corner test for every chunk.
// ---------------------------------------------------------------------------
template <int C> struct Boxes { float x0[C], x1[C], y0[C], y1[C], z0[C], z1[C]; };

__device__ __forceinline__ uint32_t plane_class(const float4 pl)
{
    const float lim = 1.2379400392853803e27f;      // 2^90
    const bool ok = fabsf(pl.x) <= lim && fabsf(pl.y) <= lim && fabsf(pl.z) <= lim && fabsf(pl.w) <= lim;   // false for NaN
    return (pl.x < 0.0f ? 1u : 0u) | (pl.y < 0.0f ? 2u : 0u) | (pl.z < 0.0f ? 4u : 0u) | (ok ? 0u : 8u);
}

// pm[c] = running minimum over the planes of the distance at the p corner, ns[c] = number of planes whose distance
// at the n corner has its sign bit set; K = sign class of the plane
template <uint32_t K, int C>
__device__ __forceinline__ void extreme_step(const float4 pl, const Boxes<C> &b, float (&pm)[C], uint32_t (&ns)[C])
{
#pragma unroll
    for (int c = 0; c < C; ++c) {
        const float nx = (K & 1u) ? b.x1[c] : b.x0[c], px = (K & 1u) ? b.x0[c] : b.x1[c];
        const float ny = (K & 2u) ? b.y1[c] : b.y0[c], py = (K & 2u) ? b.y0[c] : b.y1[c];
        const float nz = (K & 4u) ? b.z1[c] : b.z0[c], pz = (K & 4u) ? b.z0[c] : b.z1[c];
        // scalar products (see plane_dist), packed sums in the reference's association
        float2 s = __fadd2_rn(make_float2(__fmul_rn(pl.x, nx), __fmul_rn(pl.x, px)),
                              make_float2(__fmul_rn(pl.y, ny), __fmul_rn(pl.y, py)));
        s = __fadd2_rn(s, make_float2(__fmul_rn(pl.z, nz), __fmul_rn(pl.z, pz)));
        s = __fadd2_rn(s, make_float2(pl.w, pl.w));
        pm[c] = fminf(pm[c], s.y);
        ns[c] += __float_as_uint(s.x) >> 31;
    }
}

// the reference's test of one AABB, corner by corner (order independent: plane by plane here).  Only chunks that
// straddle a plane come here, so it is a rolled loop that stays small in code and registers.
__device__ __forceinline__ bool corner_test(const float4 *pl, float x0, float x1, float y0, float y1, float z0, float z1)
{
    const float inf = __int_as_float(0x7f800000);
    float2 m[4] = {make_float2(inf, inf), make_float2(inf, inf), make_float2(inf, inf), make_float2(inf, inf)};
#pragma unroll 1
    for (uint32_t p = 0; p < 6; ++p) {
        float2 t[4];
        plane_dist(pl[p], x0, x1, y0, y1, z0, z1, t);
        // fminf ignores a NaN operand, which is the reference's behaviour (a NaN distance never rejects, :33)
#pragma unroll
        for (int c = 0; c < 4; ++c) m[c] = make_float2(fminf(m[c].x, t[c].x), fminf(m[c].y, t[c].y));
    }
    return !(best_corner(m) < 0.0f);
}

// Visibility bits of a lane's C chunks against one plane set.  pl: six planes (shared memory), cls: their
// classes, four bits per plane.  Must be called by the whole warp (planes are skipped once every lane is decided).
template <int UNROLL, int C>
__device__ __forceinline__ uint32_t frame_test(const float4 *pl, uint32_t cls, const Boxes<C> &b, uint32_t live)
{
    uint32_t vis = 0, todo = live;
    if (!(cls & 0x888888u)) {
        const float inf = __int_as_float(0x7f800000);
        float pm[C];
        uint32_t ns[C];
#pragma unroll
        for (int c = 0; c < C; ++c) { pm[c] = inf; ns[c] = 0; }
#pragma unroll(UNROLL)
        for (uint32_t i = 0; i < 6; ++i) {
            // the result does not depend on the order of the planes
            const uint32_t p = (VX_FR_SIDES_FIRST && i < 4u) ? (i ^ 2u) : i;
            const float4 q = pl[p];
            switch ((cls >> (4u * p)) & 7u) {                   // warp-uniform
            case 0: extreme_step<0>(q, b, pm, ns); break;
            case 1: extreme_step<1>(q, b, pm, ns); break;
            case 2: extreme_step<2>(q, b, pm, ns); break;
            case 3: extreme_step<3>(q, b, pm, ns); break;
            case 4: extreme_step<4>(q, b, pm, ns); break;
            case 5: extreme_step<5>(q, b, pm, ns); break;
            case 6: extreme_step<6>(q, b, pm, ns); break;
            default: extreme_step<7>(q, b, pm, ns); break;
            }
            if (i == 1 || i == 3) {
                // every chunk of the warp rejected already?  (slots past the end of the list hold chunk (0, 0, 0): they
                // can only keep a tail warp going, never change a result)
                float open = pm[0];
#pragma unroll
                for (int c = 1; c < C; ++c) open = fmaxf(open, pm[c]);
                if (!__any_sync(VX_FULL, !(open < 0.0f))) return 0;
            }
        }
        todo = 0;
#pragma unroll
        for (int c = 0; c < C; ++c) {
            if (((live >> c) & 1u) && !(pm[c] < 0.0f)) {
                if (ns[c] <= 1u) vis |= 1u << c;
                else todo |= 1u << c;                             // straddles two planes or more
            }
        }
    }
    while (todo) {
        const uint32_t c = __ffs(todo) - 1;
        todo &= todo - 1;
        float x0 = b.x0[0], x1 = b.x1[0], y0 = b.y0[0], y1 = b.y1[0], z0 = b.z0[0], z1 = b.z1[0];
#pragma unroll
        for (int k = 1; k < C; ++k)
            if (c == (uint32_t)k) { x0 = b.x0[k]; x1 = b.x1[k]; y0 = b.y0[k]; y1 = b.y1[k]; z0 = b.z0[k]; z1 = b.z1[k]; }
        vis |= (corner_test(pl, x0, x1, y0, y1, z0, z1) ? 1u : 0u) << c;
    }
    return vis;
}

// coordinate.rs:53-58: (chunk * 32) as f32 (wrapping i32 product); frustum_culling.rs:19-26: the far corners are p + 32
template <int C>
__device__ __forceinline__ void set_box(Boxes<C> &b, int c, int32_t cx, int32_t cy, int32_t cz)
{
    b.x0[c] = (float)(int32_t)((uint32_t)cx * 32u);
    b.y0[c] = (float)(int32_t)((uint32_t)cy * 32u);
    b.z0[c] = (float)(int32_t)((uint32_t)cz * 32u);
    b.x1[c] = __fadd_rn(b.x0[c], 32.0f);
    b.y1[c] = __fadd_rn(b.y0[c], 32.0f);
    b.z1[c] = __fadd_rn(b.z0[c], 32.0f);
}

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
