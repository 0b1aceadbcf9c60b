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
#include "../../wgpu-voxel-engine_b200/csrc/vx_noise.cuh"

// same coordinate construction and hoisting as vx_fill_kernel (kernels_fill.cu)
extern "C" void shim_noise_chunk(const int32_t *c, int32_t seed, float freq, float gain, float lac, int oct,
                                 float *out)
{
    const float ox = (float)(int32_t)((uint32_t)c[0] * 32u);
    const float oy = (float)(int32_t)((uint32_t)c[2] * 32u);
    const float oz = (float)(int32_t)((uint32_t)c[1] * 32u);
    for (int zz = 0; zz < 32; ++zz)
        for (int lane = 0; lane < 32; ++lane) {
            const float sx = ox + (float)lane, sy = oy + (float)zz;
            for (int y = 0; y < 32; ++y) {
                float x = sx * freq, yy = sy * freq, z = (oz + (float)y) * freq, amp = 1.0f;
                float r = vxn_simplex3(x, yy, z, x + yy, (uint32_t)seed);
                for (int o = 1; o < oct; ++o) {
                    x *= lac; yy *= lac; z *= lac; amp *= gain;
                    r = vxn_simplex3(x, yy, z, x + yy, (uint32_t)seed) * amp + r;
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
                float r = vxn_simplex3_sign<FUSED>(x, yy, z, x + yy, (uint32_t)seed, ll);
                for (int o = 1; o < 3; ++o) {
                    x *= lac; yy *= lac; z *= lac; amp *= gain;
                    r = vxn_simplex3_sign<FUSED>(x, yy, z, x + yy, (uint32_t)seed, ll) * amp + r;
                }
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
