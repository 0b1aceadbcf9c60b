// Host build of the DEVICE noise header (csrc/vx_noise.cuh) for CPU testing:
// the CUDA intrinsics are mapped to plain separately-rounded f32 ops
// (compile with -ffp-contract=off).  Test infrastructure only.
#include <cmath>
#include <cstdint>
#include <cstring>
#define __device__
#define __forceinline__ inline
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
static inline int __float2int_rz(float f) { return (int)f; }
static inline float __uint_as_float(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline uint32_t __float_as_uint(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
struct float4 { float x, y, z, w; };
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
// packed f32x2 intrinsics of sm_100: two independent IEEE-rn results
struct float2 { float x, y; };
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline float2 __fadd2_rn(float2 a, float2 b) { return float2{a.x + b.x, a.y + b.y}; }
static inline float2 __fmul2_rn(float2 a, float2 b) { return float2{a.x * b.x, a.y * b.y}; }
static inline float2 __ffma2_rn(float2 a, float2 b, float2 c) { return float2{fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)}; }
static inline float __saturatef(float v) { return v != v ? 0.0f : fminf(fmaxf(v, 0.0f), 1.0f); }
#include "../../wgpu-voxel-engine_b200/csrc/vx_noise.cuh"

// same coordinate construction and hoisting as vx_fill_kernel (kernels_fill.cu)
// simplex_scale of the following calls (vx_noise_params.simplex_scale); 1 = raw sum
static float g_scale = 1.0f;
extern "C" void shim_set_scale(float s) { g_scale = s; }

extern "C" void shim_noise_chunk(const int32_t *c, int32_t seed, float freq, float gain, float lac, int oct,
                                 float *out)
{
    const float scale = g_scale;
    const float ox = (float)(int32_t)((uint32_t)c[0] * 32u);
    const float oy = (float)(int32_t)((uint32_t)c[2] * 32u);
    const float oz = (float)(int32_t)((uint32_t)c[1] * 32u);
    for (int zz = 0; zz < 32; ++zz)
        for (int lane = 0; lane < 32; ++lane) {
            const float sx = ox + (float)lane, sy = oy + (float)zz;
            for (int y = 0; y < 32; ++y) {
                float x = sx * freq, yy = sy * freq, z = (oz + (float)y) * freq, amp = 1.0f;
                float r = vxn_simplex3(x, yy, z, x + yy, (uint32_t)seed) * scale;
                for (int o = 1; o < oct; ++o) {
                    x *= lac; yy *= lac; z *= lac; amp *= gain;
                    r = (vxn_simplex3(x, yy, z, x + yy, (uint32_t)seed) * scale) * amp + r;
                }
                out[lane + 32 * zz + 1024 * y] = r;
            }
        }
}

// the sign-only (coefficient-table) variant used by vx_fill_brick_kernel
template <bool FUSED>
static void noise_chunk_sign(const int32_t *c, int32_t seed, float freq, float gain, float lac, float *out)
{
    float4 lut[VXN_LUT_FLOAT4S];
    for (uint32_t i = 0; i < VXN_LUT_FLOAT4S; ++i) lut[i] = vxn_grad_coeffs(i >> 3);
    const float ox = (float)(int32_t)((uint32_t)c[0] * 32u);
    const float oy = (float)(int32_t)((uint32_t)c[2] * 32u);
    const float oz = (float)(int32_t)((uint32_t)c[1] * 32u);
    for (int zz = 0; zz < 32; ++zz)
        for (int lane = 0; lane < 32; ++lane) {
            const float sx = ox + (float)lane, sy = oy + (float)zz;
            const float4 *ll = lut + (lane & 7);
            for (int y = 0; y < 32; ++y) {
                float x = sx * freq, yy = sy * freq, z = (oz + (float)y) * freq, amp = 1.0f;
                float sv[3], av[3];
                for (int o = 0; o < 3; ++o) {
                    if (o) { x *= lac; yy *= lac; z *= lac; amp *= gain; }
                    sv[o] = vxn_simplex3_sign<FUSED>(x, yy, z, x + yy, (uint32_t)seed, ll);
                    av[o] = amp;
                }
                float r;
                if (FUSED)      // the kernels' fused combination: one product by fl(scale * amp) per octave
                    r = fmaf(sv[0], g_scale, fmaf(sv[1], g_scale * av[1], sv[2] * (g_scale * av[2])));
                else            // the reference's order: r = s0'; r = s1' * amp + r; ...   with s' = s * scale
                    r = ((sv[2] * g_scale) * av[2]) + (((sv[1] * g_scale) * av[1]) + (sv[0] * g_scale));
                out[lane + 32 * zz + 1024 * y] = r;
            }
        }
}

extern "C" void shim_noise_chunk_sign(const int32_t *c, int32_t seed, float freq, float gain, float lac, float *out)
{
    noise_chunk_sign<false>(c, seed, freq, gain, lac, out);
}
// fused (FMA) evaluation of the smooth part: NOT bit-equal, bounded difference
extern "C" void shim_noise_chunk_fused(const int32_t *c, int32_t seed, float freq, float gain, float lac, float *out)
{
    noise_chunk_sign<true>(c, seed, freq, gain, lac, out);
}
extern "C" float shim_fused_err_ulps(void) { return VXN_FUSED_ERR_ULPS; }

// the two-voxels-per-thread packed (fused-only) variant, paired as vx_fill_brick2_kernel pairs them: (y, y + 2)
static void noise_chunk_x2(const int32_t *c, int32_t seed, float freq, float gain, float lac, float *out)
{
    float4 lut[VXN_LUT_FLOAT4S];
    for (uint32_t i = 0; i < VXN_LUT_FLOAT4S; ++i) lut[i] = vxn_grad_coeffs(i >> 3);
    const float ox = (float)(int32_t)((uint32_t)c[0] * 32u);
    const float oy = (float)(int32_t)((uint32_t)c[2] * 32u);
    const float oz = (float)(int32_t)((uint32_t)c[1] * 32u);
    const float amp1 = 1.0f * gain, amp2 = amp1 * gain, ka1 = g_scale * amp1, ka2 = g_scale * amp2;
    for (int zz = 0; zz < 32; ++zz)
        for (int lane = 0; lane < 32; ++lane) {
            const float sx = ox + (float)lane, sy = oy + (float)zz;
            const float4 *ll = lut + (lane & 7);
            float hx[3], hy[3], hxy[3];
            hx[0] = sx * freq; hy[0] = sy * freq; hxy[0] = hx[0] + hy[0];
            for (int o = 1; o < 3; ++o) { hx[o] = hx[o - 1] * lac; hy[o] = hy[o - 1] * lac; hxy[o] = hx[o] + hy[o]; }
            for (int it = 0; it < 8; ++it)
                for (int yy = 0; yy < 2; ++yy) {
                    const int ya = 4 * it + yy, yb = ya + 2;
                    const float za0 = (oz + (float)ya) * freq, zb0 = (oz + (float)yb) * freq;
                    const float za1 = za0 * lac, zb1 = zb0 * lac, za2 = za1 * lac, zb2 = zb1 * lac;
                    const float2 s3 = vxn_simplex3_fused_x2(VX2(hx[2]), VX2(hy[2]), make_float2(za2, zb2), VX2(hxy[2]), (uint32_t)seed, ll);
                    const float2 s2 = vxn_simplex3_fused_x2(VX2(hx[1]), VX2(hy[1]), make_float2(za1, zb1), VX2(hxy[1]), (uint32_t)seed, ll);
                    const float2 s1 = vxn_simplex3_fused_x2(VX2(hx[0]), VX2(hy[0]), make_float2(za0, zb0), VX2(hxy[0]), (uint32_t)seed, ll);
                    const float2 r = __ffma2_rn(s1, VX2(g_scale), __ffma2_rn(s2, VX2(ka1), __fmul2_rn(s3, VX2(ka2))));   // kernel's fused order
                    out[lane + 32 * zz + 1024 * ya] = r.x;
                    out[lane + 32 * zz + 1024 * yb] = r.y;
                }
        }
}
// the lattice-table variant as vx_fill_lattice_kernel runs it: per (chunk, 4-wide z block) index boxes from
// the 8 box corners, one table per octave, then the packed fused evaluation.  Returns the number of z
// blocks whose boxes did NOT fit the tables (their voxels are left untouched).
static int noise_chunk_lattice(const int32_t *c, int32_t seed, float freq, float gain, float lac, float *out)
{
    const float ox = (float)(int32_t)((uint32_t)c[0] * 32u);
    const float oy = (float)(int32_t)((uint32_t)c[2] * 32u);
    const float oz = (float)(int32_t)((uint32_t)c[1] * 32u);
    const float amp1 = 1.0f * gain, amp2 = amp1 * gain, ka1 = g_scale * amp1, ka2 = g_scale * amp2;
    static char tab[3][VXL_TAB_BYTES];
    int unfit = 0;
    for (int zblk = 0; zblk < 8; ++zblk) {
        float lo[3][3], hi[3][3];
        bool fit = true;
        float4 og[3];
        for (int o = 0; o < 3; ++o) {
            for (int cn = 0; cn < 8; ++cn) {
                float fx = (ox + ((cn & 1) ? 31.0f : 0.0f)) * freq;
                float fy = (oy + (float)(zblk * 4 + ((cn & 2) ? 3 : 0))) * freq;
                float fz = (oz + ((cn & 4) ? 31.0f : 0.0f)) * freq;
                for (int q = 0; q < o; ++q) { fx *= lac; fy *= lac; fz *= lac; }
                float v[3];
                vxl_cell_floor(fx, fy, fz, fx + fy, &v[0], &v[1], &v[2]);
                for (int d = 0; d < 3; ++d) {
                    if (cn == 0 || v[d] < lo[o][d]) lo[o][d] = v[d];
                    if (cn == 0 || v[d] > hi[o][d]) hi[o][d] = v[d];
                }
            }
            for (int d = 0; d < 3; ++d) fit = fit && vxl_axis_fits(lo[o][d], hi[o][d]);
            og[o] = make_float4(lo[o][0], lo[o][1], lo[o][2], VXL_MAGIC);
        }
        if (!fit) { ++unfit; continue; }
        for (int o = 0; o < 3; ++o) {
            memset(tab[o], 0xff, VXL_TAB_BYTES);          // NaN pattern: reading an unbuilt entry poisons the result
            const int32_t i0 = (int32_t)lo[o][0], j0 = (int32_t)lo[o][1], k0 = (int32_t)lo[o][2];
            for (int32_t i = i0; i <= (int32_t)hi[o][0] + 1; ++i)
                for (int32_t j = j0; j <= (int32_t)hi[o][1] + 1; ++j)
                    for (int32_t k = k0; k <= (int32_t)hi[o][2] + 1; ++k) {
                        const float4 e = vxl_entry((uint32_t)seed, i, j, k);
                        memcpy(tab[o] + (uint32_t)(i - i0) * VXL_SI + (uint32_t)(j - j0) * VXL_SJ +
                                   (uint32_t)(k - k0) * VXL_SK, &e, 16);
                    }
        }
        for (int zz = zblk * 4; zz < zblk * 4 + 4; ++zz)
            for (int lane = 0; lane < 32; ++lane) {
                const float sx = ox + (float)lane, sy = oy + (float)zz;
                float hx[3], hy[3], hxy[3];
                hx[0] = sx * freq; hy[0] = sy * freq; hxy[0] = hx[0] + hy[0];
                for (int o = 1; o < 3; ++o) { hx[o] = hx[o - 1] * lac; hy[o] = hy[o - 1] * lac; hxy[o] = hx[o] + hy[o]; }
                for (int it = 0; it < 8; ++it)
                    for (int yy = 0; yy < 2; ++yy) {
                        const int ya = 4 * it + yy, yb = ya + 2;
                        const float za0 = (oz + (float)ya) * freq, zb0 = (oz + (float)yb) * freq;
                        const float za1 = za0 * lac, zb1 = zb0 * lac, za2 = za1 * lac, zb2 = zb1 * lac;
                        const float2 s3 = vxn_simplex3_lattice_x2(VX2(hx[2]), VX2(hy[2]), make_float2(za2, zb2), VX2(hxy[2]), og[2], tab[2]);
                        const float2 s2 = vxn_simplex3_lattice_x2(VX2(hx[1]), VX2(hy[1]), make_float2(za1, zb1), VX2(hxy[1]), og[1], tab[1]);
                        const float2 s1 = vxn_simplex3_lattice_x2(VX2(hx[0]), VX2(hy[0]), make_float2(za0, zb0), VX2(hxy[0]), og[0], tab[0]);
                        const float2 r = __ffma2_rn(s1, VX2(g_scale), __ffma2_rn(s2, VX2(ka1), __fmul2_rn(s3, VX2(ka2))));
                        out[lane + 32 * zz + 1024 * ya] = r.x;
                        out[lane + 32 * zz + 1024 * yb] = r.y;
                    }
            }
    }
    return unfit;
}
extern "C" int shim_noise_chunk_lattice(const int32_t *c, int32_t seed, float freq, float gain, float lac, float *out)
{
    return noise_chunk_lattice(c, seed, freq, gain, lac, out);
}

// scalar evaluation in the same two combination orders, for an element-wise bit comparison
template <bool FUSED>
static void noise_chunk_x1(const int32_t *c, int32_t seed, float freq, float gain, float lac, float *out)
{
    float4 lut[VXN_LUT_FLOAT4S];
    for (uint32_t i = 0; i < VXN_LUT_FLOAT4S; ++i) lut[i] = vxn_grad_coeffs(i >> 3);
    const float ox = (float)(int32_t)((uint32_t)c[0] * 32u);
    const float oy = (float)(int32_t)((uint32_t)c[2] * 32u);
    const float oz = (float)(int32_t)((uint32_t)c[1] * 32u);
    const float amp1 = 1.0f * gain, amp2 = amp1 * gain;
    for (int zz = 0; zz < 32; ++zz)
        for (int lane = 0; lane < 32; ++lane) {
            const float sx = ox + (float)lane, sy = oy + (float)zz;
            const float4 *ll = lut + (lane & 7);
            const float x0 = sx * freq, y0 = sy * freq, x1 = x0 * lac, y1 = y0 * lac, x2 = x1 * lac, y2 = y1 * lac;
            for (int y = 0; y < 32; ++y) {
                const float z0 = (oz + (float)y) * freq, z1 = z0 * lac, z2 = z1 * lac;
                const float s3 = vxn_simplex3_sign<FUSED>(x2, y2, z2, x2 + y2, (uint32_t)seed, ll);
                const float s2 = vxn_simplex3_sign<FUSED>(x1, y1, z1, x1 + y1, (uint32_t)seed, ll);
                const float s1 = vxn_simplex3_sign<FUSED>(x0, y0, z0, x0 + y0, (uint32_t)seed, ll);
                out[lane + 32 * zz + 1024 * y] = FUSED ? fmaf(s1, g_scale, fmaf(s2, g_scale * amp1, s3 * (g_scale * amp2)))
                                                       : ((s3 * g_scale) * amp2 + ((s2 * g_scale) * amp1 + (s1 * g_scale)));
            }
        }
}
extern "C" void shim_noise_chunk_x2(const int32_t *c, int32_t seed, float freq, float gain, float lac, int fused, float *out)
{
    (void)fused;
    noise_chunk_x2(c, seed, freq, gain, lac, out);
}
extern "C" void shim_noise_chunk_x1(const int32_t *c, int32_t seed, float freq, float gain, float lac, int fused, float *out)
{
    if (fused) noise_chunk_x1<true>(c, seed, freq, gain, lac, out);
    else noise_chunk_x1<false>(c, seed, freq, gain, lac, out);
}
