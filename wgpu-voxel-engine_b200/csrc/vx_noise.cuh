// vx_noise.cuh -- simdnoise 3.1.6 `fbm_3d` over hashed `simplex_3d`, one lane
// per voxel.  Follows the reference's call shape perlin_noise.rs:5-22 and the
// crate's published algorithm.  Every f32 operation is a
// separately rounded IEEE op (__fmul_rn/__fadd_rn are never contracted into
// FMA), integer arithmetic wraps mod 2^32 and the hash shift is arithmetic,
// so results are bit-identical to the scalar/AVX2 CPU evaluation.
#pragma once
#include <cstdint>

#define VXN_F3 0x1.555556p-2f      /* (f32)(1/3) */
#define VXN_G3 0x1.555556p-3f      /* (f32)(1/6) */
#define VXN_G33 (-0.5f)            /* 3/6 - 1    */
#define VXN_XP 1619u
#define VXN_YP 31337u
#define VXN_ZP 6971u

__device__ __forceinline__ float vxn_grad(uint32_t h, float x, float y, float z)
{
    // h = ((i ^ seed) ^ j) ^ k  (xor order is irrelevant)
    h = ((h * h) * 60493u) * h;
    h = (uint32_t)(((int32_t)h) >> 13) ^ h;
    // h13 = h & 13:  h13 < 8  <=> bit 3 clear;  h13 < 2 <=> bits 2,3 clear;
    // h13 == 12 <=> bits 2,3 set and bit 0 clear
    const float u = (h & 8u) ? y : x;
    const float v = ((h & 12u) == 0u) ? y : ((h & 13u) == 12u ? x : z);
    const float su = __uint_as_float(__float_as_uint(u) ^ (h << 31));
    const float sv = __uint_as_float(__float_as_uint(v) ^ ((h << 30) & 0x80000000u));
    return __fadd_rn(su, sv);
}

__device__ __forceinline__ float vxn_weight(float a, float b, float c)
{
    float t = __fsub_rn(__fsub_rn(__fsub_rn(0.6f, __fmul_rn(a, a)), __fmul_rn(b, b)), __fmul_rn(c, c));
    t = fmaxf(t, 0.0f);              // cmpge mask: negatives (and NaN) -> 0; one FMNMX
    t = __fmul_rn(t, t);
    return __fmul_rn(t, t);
}

// xy = x + y is passed in so callers can hoist it when x and y are loop
// invariant; it must be exactly __fadd_rn(x, y).
__device__ __forceinline__ float vxn_simplex3(float x, float y, float z, float xy, uint32_t seed)
{
    const float f = __fmul_rn(VXN_F3, __fadd_rn(xy, z));
    const float xf = floorf(__fadd_rn(x, f));
    const float yf = floorf(__fadd_rn(y, f));
    const float zf = floorf(__fadd_rn(z, f));

    // Lattice hashes.  Each axis only ever takes two values over the four
    // simplex corners (base, base + PRIME), so both are formed once and the
    // middle corners select between them; the seed is folded into the x term.
    const uint32_t ib0 = (uint32_t)__float2int_rz(xf) * VXN_XP;
    const uint32_t ia = ib0 ^ seed, ib = (ib0 + VXN_XP) ^ seed;
    const uint32_t ja = (uint32_t)__float2int_rz(yf) * VXN_YP, jb = ja + VXN_YP;
    const uint32_t ka = (uint32_t)__float2int_rz(zf) * VXN_ZP, kb = ka + VXN_ZP;

    const float g = __fmul_rn(VXN_G3, __fadd_rn(__fadd_rn(xf, yf), zf));
    const float x0 = __fsub_rn(x, __fsub_rn(xf, g));
    const float y0 = __fsub_rn(y, __fsub_rn(yf, g));
    const float z0 = __fsub_rn(z, __fsub_rn(zf, g));

    const bool xy_ = x0 >= y0, yz_ = y0 >= z0, xz_ = x0 >= z0;
    const bool i1 = xy_ && xz_, j1 = !xy_ && yz_, k1 = !xz_ && !yz_;
    const bool i2 = xy_ || xz_, j2 = !xy_ || yz_, k2 = !(xz_ && yz_);

    // (x0 + (m ? -1 : 0)) + C: the -1 sum is shared by corners 1 and 2.  The
    // "+ 0" arm only differs from x0 for -0, which the following + C absorbs.
    const float xm = __fadd_rn(x0, -1.0f), ym = __fadd_rn(y0, -1.0f), zm = __fadd_rn(z0, -1.0f);
    const float x1 = __fadd_rn(i1 ? xm : x0, VXN_G3);
    const float y1 = __fadd_rn(j1 ? ym : y0, VXN_G3);
    const float z1 = __fadd_rn(k1 ? zm : z0, VXN_G3);
    const float x2 = __fadd_rn(i2 ? xm : x0, VXN_F3);
    const float y2 = __fadd_rn(j2 ? ym : y0, VXN_F3);
    const float z2 = __fadd_rn(k2 ? zm : z0, VXN_F3);
    const float x3 = __fadd_rn(x0, VXN_G33);
    const float y3 = __fadd_rn(y0, VXN_G33);
    const float z3 = __fadd_rn(z0, VXN_G33);

    const float v0 = __fmul_rn(vxn_weight(x0, y0, z0), vxn_grad(ia ^ ja ^ ka, x0, y0, z0));
    const float v1 = __fmul_rn(vxn_weight(x1, y1, z1),
                               vxn_grad((i1 ? ib : ia) ^ (j1 ? jb : ja) ^ (k1 ? kb : ka), x1, y1, z1));
    const float v2 = __fmul_rn(vxn_weight(x2, y2, z2),
                               vxn_grad((i2 ? ib : ia) ^ (j2 ? jb : ja) ^ (k2 ? kb : ka), x2, y2, z2));
    const float v3 = __fmul_rn(vxn_weight(x3, y3, z3), vxn_grad(ib ^ jb ^ kb, x3, y3, z3));
    return __fadd_rn(__fadd_rn(__fadd_rn(v3, v2), v1), v0);
}

// ---------------------------------------------------------------------------
// Sign-only variant for the fast fill kernel.  Identical arithmetic except the
// gradient: instead of selecting u, v and XOR-ing sign bits (ALU pipe: 3 FSEL +
// 2 SHF + 5 LOP3/ISETP per corner) the 16 possible (h & 15) outcomes index a
// table of coefficients (cx, cy, cz) in {-1, 0, +1} and
//     grad = (cx*x + cy*y) + cz*z                       (FMA pipe + one LDS.128)
// Products by +-1 / 0 are exact and adding a +-0 term is exact, so the value
// equals (+-u) + (+-v) except possibly the SIGN OF A ZERO result -- which can
// never change `noise < 0` (a zero stays a zero through every later product
// and sum).  Not used when the f32 noise field itself is returned.
//
// Table layout: entry e for lane l sits at float4 index e*8 + (l & 7): the 8
// lanes of one LDS.128 wavefront always hit 8 distinct 16-byte bank groups.
// ---------------------------------------------------------------------------
#define VXN_LUT_FLOAT4S 128

__device__ __forceinline__ float4 vxn_grad_coeffs(uint32_t e)
{
    // perm bits: b0 sign of u, b1 sign of v, (b3,b2): 00 u=x v=y | 01 u=x v=z | 10 u=y v=z |
    // 11 u=y v=(b0 ? z : x)        [h13 < 8 ? x : y ;  h13 < 2 ? y : (h13 == 12 ? x : z)]
    const float su = (e & 1u) ? -1.0f : 1.0f, sv = (e & 2u) ? -1.0f : 1.0f;
    float4 c = make_float4(0.f, 0.f, 0.f, 0.f);
    const uint32_t sel = (e >> 2) & 3u;
    if (sel == 0u) { c.x = su; c.y = sv; }
    else if (sel == 1u) { c.x = su; c.z = sv; }
    else if (sel == 2u) { c.y = su; c.z = sv; }
    else if (e & 1u) { c.y = su; c.z = sv; }
    else { c.y = su; c.x = sv; }
    return c;
}

#if defined(__CUDACC__)
// lut_lane = 32-bit shared-window address of this lane's column of a table whose base is
// 2048-byte aligned IN THE SHARED WINDOW (the kernel carves it out of a 4 KB buffer at run
// time; __align__ on a static __shared__ array does not guarantee it), so "base + e*128"
// is a bit-field insert: ((t << 7) & 0x780) | base -> SHF + one LOP3.
typedef uint32_t vxn_lut_ref;
__device__ __forceinline__ float4 vxn_lut_load(vxn_lut_ref lut_lane, uint32_t t)
{
    const uint32_t addr = ((t << 7) & 0x780u) | lut_lane;
    float4 c;
    asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(c.x), "=f"(c.y), "=f"(c.z), "=f"(c.w) : "r"(addr));
    return c;
}
#else
typedef const float4 *vxn_lut_ref;          // host build (tests/host_shim): plain pointer
static inline float4 vxn_lut_load(vxn_lut_ref lut_lane, uint32_t t) { return lut_lane[(t & 15u) * 8u]; }
#endif

// FUSED = false: separately rounded ops, bit-equal to the exact noise up to the sign of zero.
// FUSED = true : same cell selection, but weights / gradient dots / sums use FMA (fewer
//                instructions, different roundings).  |fused - exact| <= VXN_FUSED_ERR_ULPS * 2^-24
//                per simplex_3d value (bound derived in DESIGN.md 4.1), so a fused result is only
//                trusted for a SIGN decision when it clears that margin; otherwise the caller
//                re-evaluates with FUSED = false.
#define VXN_FUSED_ERR_ULPS 5.4f

template <bool FUSED>
__device__ __forceinline__ float vxn_grad_lut(uint32_t h, float x, float y, float z, vxn_lut_ref lut_lane)
{
    h = ((h * h) * 60493u) * h;
    const float4 c = vxn_lut_load(lut_lane, (uint32_t)(((int32_t)h) >> 13) ^ h);    // entry (.. & 15)
    if (FUSED) return __fmaf_rn(c.z, z, __fmaf_rn(c.y, y, __fmul_rn(c.x, x)));
    return __fadd_rn(__fadd_rn(__fmul_rn(c.x, x), __fmul_rn(c.y, y)), __fmul_rn(c.z, z));
}

template <bool FUSED>
__device__ __forceinline__ float vxn_weight_t(float a, float b, float c)
{
    float t;
    if (FUSED) t = __fmaf_rn(-c, c, __fmaf_rn(-b, b, __fmaf_rn(-a, a, 0.6f)));
    else t = __fsub_rn(__fsub_rn(__fsub_rn(0.6f, __fmul_rn(a, a)), __fmul_rn(b, b)), __fmul_rn(c, c));
    t = fmaxf(t, 0.0f);
    t = __fmul_rn(t, t);
    return __fmul_rn(t, t);
}

template <bool FUSED>
__device__ __forceinline__ float vxn_simplex3_sign(float x, float y, float z, float xy, uint32_t seed,
                                                   vxn_lut_ref lut_lane)
{
    // ---- cell selection: ALWAYS the reference's exact operation sequence, so the fused and the
    //      exact evaluation pick the same simplex and the same corner order ----------------------
    const float f = __fmul_rn(VXN_F3, __fadd_rn(xy, z));
    const float xf = floorf(__fadd_rn(x, f));
    const float yf = floorf(__fadd_rn(y, f));
    const float zf = floorf(__fadd_rn(z, f));

    const uint32_t ib0 = (uint32_t)__float2int_rz(xf) * VXN_XP;
    const uint32_t ia = ib0 ^ seed, ib = (ib0 + VXN_XP) ^ seed;
    const uint32_t ja = (uint32_t)__float2int_rz(yf) * VXN_YP, jb = ja + VXN_YP;
    const uint32_t ka = (uint32_t)__float2int_rz(zf) * VXN_ZP, kb = ka + VXN_ZP;

    const float g = __fmul_rn(VXN_G3, __fadd_rn(__fadd_rn(xf, yf), zf));
    const float x0 = __fsub_rn(x, __fsub_rn(xf, g));
    const float y0 = __fsub_rn(y, __fsub_rn(yf, g));
    const float z0 = __fsub_rn(z, __fsub_rn(zf, g));

    const bool xy_ = x0 >= y0, yz_ = y0 >= z0, xz_ = x0 >= z0;
    const bool i1 = xy_ && xz_, j1 = !xy_ && yz_, k1 = !xz_ && !yz_;
    const bool i2 = xy_ || xz_, j2 = !xy_ || yz_, k2 = !(xz_ && yz_);

    const float xm = __fadd_rn(x0, -1.0f), ym = __fadd_rn(y0, -1.0f), zm = __fadd_rn(z0, -1.0f);
    const float x1 = __fadd_rn(i1 ? xm : x0, VXN_G3);
    const float y1 = __fadd_rn(j1 ? ym : y0, VXN_G3);
    const float z1 = __fadd_rn(k1 ? zm : z0, VXN_G3);
    const float x2 = __fadd_rn(i2 ? xm : x0, VXN_F3);
    const float y2 = __fadd_rn(j2 ? ym : y0, VXN_F3);
    const float z2 = __fadd_rn(k2 ? zm : z0, VXN_F3);
    const float x3 = __fadd_rn(x0, VXN_G33);
    const float y3 = __fadd_rn(y0, VXN_G33);
    const float z3 = __fadd_rn(z0, VXN_G33);

    // ---- smooth part --------------------------------------------------------------------------
    const float w0 = vxn_weight_t<FUSED>(x0, y0, z0), g0 = vxn_grad_lut<FUSED>(ia ^ ja ^ ka, x0, y0, z0, lut_lane);
    const float w1 = vxn_weight_t<FUSED>(x1, y1, z1),
                g1 = vxn_grad_lut<FUSED>((i1 ? ib : ia) ^ (j1 ? jb : ja) ^ (k1 ? kb : ka), x1, y1, z1, lut_lane);
    const float w2 = vxn_weight_t<FUSED>(x2, y2, z2),
                g2 = vxn_grad_lut<FUSED>((i2 ? ib : ia) ^ (j2 ? jb : ja) ^ (k2 ? kb : ka), x2, y2, z2, lut_lane);
    const float w3 = vxn_weight_t<FUSED>(x3, y3, z3), g3 = vxn_grad_lut<FUSED>(ib ^ jb ^ kb, x3, y3, z3, lut_lane);
    if (FUSED) return __fmaf_rn(w0, g0, __fmaf_rn(w1, g1, __fmaf_rn(w2, g2, __fmul_rn(w3, g3))));
    const float v0 = __fmul_rn(w0, g0), v1 = __fmul_rn(w1, g1), v2 = __fmul_rn(w2, g2), v3 = __fmul_rn(w3, g3);
    return __fadd_rn(__fadd_rn(__fadd_rn(v3, v2), v1), v0);
}
