// vx_frustum.cuh -- device-side logic of the frustum kernels (kernels_frustum.cu): the reference's plane distances,
// the extreme-corner classification and the corner-by-corner test.  Kept apart from the kernels so that
// tests/test_device_frustum_on_host.py can compile exactly this source for the host (intrinsic shim, a warp of one
// lane) and compare it with the oracle without a GPU.
#pragma once

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

// evaluate the side planes (stored third and fourth, frustum_culling.rs:84) before near / far: the result does not
// depend on the order, and in an open world the side planes reject the most chunks, so more warps are decided after
// the first pair (378 -> 406 G tests/s on the camera sweep of BASELINE configs[4])
#ifndef VX_FR_SIDES_FIRST
#define VX_FR_SIDES_FIRST 1
#endif
// two planes per iteration of the classification loop: their p-corner distances join the running minimum in one
// three-input minimum (FMNMX3): 434 -> 450 G tests/s at 1000 frames per launch, 5.21 -> 5.11 us for one frame
#ifndef VX_FR_PAIR_MIN
#define VX_FR_PAIR_MIN 1
#endif
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
template <uint32_t K, int C, bool ACCUMULATE = true>
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
        pm[c] = ACCUMULATE ? fminf(pm[c], s.y) : s.y;
        ns[c] += __float_as_uint(s.x) >> 31;
    }
}

// one plane through the body specialised for its sign class (a warp-uniform branch)
template <int C, bool ACCUMULATE>
__device__ __forceinline__ void extreme_dispatch(const float4 q, uint32_t k, const Boxes<C> &b, float (&pm)[C], uint32_t (&ns)[C])
{
    switch (k) {
    case 0: extreme_step<0, C, ACCUMULATE>(q, b, pm, ns); break;
    case 1: extreme_step<1, C, ACCUMULATE>(q, b, pm, ns); break;
    case 2: extreme_step<2, C, ACCUMULATE>(q, b, pm, ns); break;
    case 3: extreme_step<3, C, ACCUMULATE>(q, b, pm, ns); break;
    case 4: extreme_step<4, C, ACCUMULATE>(q, b, pm, ns); break;
    case 5: extreme_step<5, C, ACCUMULATE>(q, b, pm, ns); break;
    case 6: extreme_step<6, C, ACCUMULATE>(q, b, pm, ns); break;
    default: extreme_step<7, C, ACCUMULATE>(q, b, pm, ns); break;
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
#if VX_FR_PAIR_MIN
        // two planes per iteration: their p-corner distances join the running minimum in one three-input minimum
#pragma unroll(UNROLL >= 2 ? UNROLL / 2 : 1)
        for (uint32_t i = 0; i < 6; i += 2) {
            // the result does not depend on the order of the planes
            const uint32_t p0 = (VX_FR_SIDES_FIRST && i < 4u) ? (i ^ 2u) : i, p1 = p0 + 1u;
            float pa[C], pb[C];
            extreme_dispatch<C, false>(pl[p0], (cls >> (4u * p0)) & 7u, b, pa, ns);
            extreme_dispatch<C, false>(pl[p1], (cls >> (4u * p1)) & 7u, b, pb, ns);
#pragma unroll
            for (int c = 0; c < C; ++c) pm[c] = fminf(fminf(pm[c], pa[c]), pb[c]);
            if (i < 4u) {
                // every chunk of the warp rejected already?  (slots past the end of the list hold chunk (0, 0, 0): they
                // can only keep a tail warp going, never change a result)
                float open = pm[0];
#pragma unroll
                for (int c = 1; c < C; ++c) open = fmaxf(open, pm[c]);
                if (!__any_sync(VX_FULL, !(open < 0.0f))) return 0;
            }
        }
#else
#pragma unroll(UNROLL)
        for (uint32_t i = 0; i < 6; ++i) {
            // the result does not depend on the order of the planes
            const uint32_t p = (VX_FR_SIDES_FIRST && i < 4u) ? (i ^ 2u) : i;
            extreme_dispatch<C, true>(pl[p], (cls >> (4u * p)) & 7u, b, pm, ns);
            if (i == 1 || i == 3) {
                // every chunk of the warp rejected already?  (slots past the end of the list hold chunk (0, 0, 0): they
                // can only keep a tail warp going, never change a result)
                float open = pm[0];
#pragma unroll
                for (int c = 1; c < C; ++c) open = fmaxf(open, pm[c]);
                if (!__any_sync(VX_FULL, !(open < 0.0f))) return 0;
            }
        }
#endif
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

