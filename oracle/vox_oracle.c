/*
 * vox_oracle.c -- CPU restatement of the reference's chunk-production path.
 * TEST INFRASTRUCTURE ONLY (see vox_oracle.h).  Build: oracle/Makefile
 * (gcc -O2 -ffp-contract=off: Rust never contracts a*b+c, neither may we).
 *
 * Every function cites the reference file:line it follows
 * (paths relative to /root/reference/src).
 */
#define _GNU_SOURCE
#include "vox_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdatomic.h>
#include <stdlib.h>
#include <string.h>

#if defined(__x86_64__)
#include <immintrin.h>
#define VXO_X86 1
#else
#define VXO_X86 0
#endif

/* ------------------------------------------------------------------------ */
/* Noise: simdnoise 3.1.6 `fbm_3d` over `simplex_3d` (UN-VENDORED third-party
 * arithmetic; restated from the crate's published algorithm).
 * All f32 ops are separately rounded; integer ops wrap mod 2^32; the hash
 * shift is arithmetic.                                                      */
/* ------------------------------------------------------------------------ */

static const float F3 = 1.0f / 3.0f;          /* skew   */
static const float G3 = 1.0f / 6.0f;          /* unskew */
static const float G33 = 3.0f / 6.0f - 1.0f;  /* -0.5   */
#define X_PRIME 1619u
#define Y_PRIME 31337u
#define Z_PRIME 6971u
/* (f32) 32.69587493801679: simplex_3d's normaliser in the later 3.1.x sources (vox_oracle.h) */
#define VXO_SIMPLEX_SCALE 32.69587493801679f

void vxo_default_params(vxo_noise_params *p)
{
    /* perlin_noise.rs:14-18 */
    p->seed = 1;
    p->freq = 0.09f;
    p->gain = 2.5f;
    p->lacunarity = 0.5f;
    p->octaves = 3;
    p->simplex_scale = VXO_SIMPLEX_SCALE;
}

static inline uint32_t hash3d(int32_t seed, uint32_t i, uint32_t j, uint32_t k)
{
    uint32_t h = ((i ^ (uint32_t)seed) ^ j) ^ k;
    h = ((h * h) * 60493u) * h;
    return (uint32_t)(((int32_t)h) >> 13) ^ h;
}

static inline float f_from_bits(uint32_t b) { float f; memcpy(&f, &b, 4); return f; }
static inline uint32_t bits_from_f(float f) { uint32_t b; memcpy(&b, &f, 4); return b; }

static inline float grad3d(int32_t seed, uint32_t i, uint32_t j, uint32_t k,
                           float x, float y, float z)
{
    uint32_t h = hash3d(seed, i, j, k);
    uint32_t h13 = h & 13u;
    float u = (h13 < 8u) ? x : y;
    float v = (h13 < 2u) ? y : ((h13 == 12u) ? x : z);
    /* sign flips are XORs of the sign bit: (h&1)->u, (h&2)->v */
    float su = f_from_bits(bits_from_f(u) ^ (h << 31));
    float sv = f_from_bits(bits_from_f(v) ^ ((h & 2u) << 30));
    return su + sv;
}

float vxo_simplex_3d(float x, float y, float z, int32_t seed)
{
    float f = F3 * ((x + y) + z);
    float x0 = floorf(x + f);
    float y0 = floorf(y + f);
    float z0 = floorf(z + f);

    uint32_t i = (uint32_t)(int32_t)x0 * X_PRIME;
    uint32_t j = (uint32_t)(int32_t)y0 * Y_PRIME;
    uint32_t k = (uint32_t)(int32_t)z0 * Z_PRIME;

    float g = G3 * ((x0 + y0) + z0);
    x0 = x - (x0 - g);
    y0 = y - (y0 - g);
    z0 = z - (z0 - g);

    int xy = x0 >= y0, yz = y0 >= z0, xz = x0 >= z0;
    int i1 = xy & xz, j1 = (!xy) & yz, k1 = (!xz) & (!yz);
    int i2 = xy | xz, j2 = (!xy) | yz, k2 = !(xz & yz);

    float x1 = (x0 + (i1 ? -1.0f : 0.0f)) + G3;
    float y1 = (y0 + (j1 ? -1.0f : 0.0f)) + G3;
    float z1 = (z0 + (k1 ? -1.0f : 0.0f)) + G3;
    float x2 = (x0 + (i2 ? -1.0f : 0.0f)) + F3;
    float y2 = (y0 + (j2 ? -1.0f : 0.0f)) + F3;
    float z2 = (z0 + (k2 ? -1.0f : 0.0f)) + F3;
    float x3 = x0 + G33;
    float y3 = y0 + G33;
    float z3 = z0 + G33;

    float t0 = ((0.6f - x0 * x0) - y0 * y0) - z0 * z0;
    float t1 = ((0.6f - x1 * x1) - y1 * y1) - z1 * z1;
    float t2 = ((0.6f - x2 * x2) - y2 * y2) - z2 * z2;
    float t3 = ((0.6f - x3 * x3) - y3 * y3) - z3 * z3;

    /* negative weights are masked to +0 (cmpge mask AND) */
    t0 = (t0 >= 0.0f) ? t0 : 0.0f;
    t1 = (t1 >= 0.0f) ? t1 : 0.0f;
    t2 = (t2 >= 0.0f) ? t2 : 0.0f;
    t3 = (t3 >= 0.0f) ? t3 : 0.0f;

    t0 = t0 * t0; t1 = t1 * t1; t2 = t2 * t2; t3 = t3 * t3;
    t0 = t0 * t0; t1 = t1 * t1; t2 = t2 * t2; t3 = t3 * t3;

    float v0 = t0 * grad3d(seed, i, j, k, x0, y0, z0);
    float v1 = t1 * grad3d(seed, i + (i1 ? X_PRIME : 0u), j + (j1 ? Y_PRIME : 0u),
                           k + (k1 ? Z_PRIME : 0u), x1, y1, z1);
    float v2 = t2 * grad3d(seed, i + (i2 ? X_PRIME : 0u), j + (j2 ? Y_PRIME : 0u),
                           k + (k2 ? Z_PRIME : 0u), x2, y2, z2);
    float v3 = t3 * grad3d(seed, i + X_PRIME, j + Y_PRIME, k + Z_PRIME, x3, y3, z3);

    float p1 = v3 + v2;
    float p2 = p1 + v1;
    return p2 + v0;
}

float vxo_fbm_3d(float x, float y, float z, float lac, float gain, int32_t octaves, int32_t seed,
                 float scale)
{
    /* simplex_3d returns `sum * scale` (scale = 1 leaves the sum untouched bit for bit) */
    float amp = 1.0f;
    float result = vxo_simplex_3d(x, y, z, seed) * scale;
    for (int32_t o = 1; o < octaves; ++o) {
        x = x * lac;
        y = y * lac;
        z = z * lac;
        amp = amp * gain;
        result = ((vxo_simplex_3d(x, y, z, seed) * scale) * amp) + result;
    }
    return result;
}

/* perlin_noise.rs:5-22: builder axes x <- pos.x*32, y <- pos.z*32, z <- pos.y*32
 * (wrapping i32 product as in a release build), so simdnoise's (x,y,z) is the
 * engine's (x,z,y) and its output index x + 32*y + 1024*z is the engine idx. */
static void chunk_origin(const int32_t coord[3], float *sx, float *sy, float *sz)
{
    *sx = (float)(int32_t)((uint32_t)coord[0] * 32u);
    *sy = (float)(int32_t)((uint32_t)coord[2] * 32u);
    *sz = (float)(int32_t)((uint32_t)coord[1] * 32u);
}

void vxo_noise_chunk(const int32_t coord[3], const vxo_noise_params *p, float *out)
{
    float sx, sy, sz;
    chunk_origin(coord, &sx, &sy, &sz);
    size_t idx = 0;
    for (int zz = 0; zz < VXO_CHUNK; ++zz) {          /* simdnoise z = engine y */
        float z = sz + (float)zz;
        for (int yy = 0; yy < VXO_CHUNK; ++yy) {      /* simdnoise y = engine z */
            float y = sy + (float)yy;
            for (int xx = 0; xx < VXO_CHUNK; ++xx) {
                float x = sx + (float)xx;
                out[idx++] = vxo_fbm_3d(x * p->freq, y * p->freq, z * p->freq,
                                        p->lacunarity, p->gain, p->octaves, p->seed, p->simplex_scale);
            }
        }
    }
}

/* ---- 8-wide lanes, same operation sequence per lane (simdnoise::avx2) ---- */
#if VXO_X86
#define AVX2 __attribute__((target("avx2")))

AVX2 static inline __m256i hash3d_v(int32_t seed, __m256i i, __m256i j, __m256i k)
{
    __m256i h = _mm256_xor_si256(i, _mm256_set1_epi32(seed));
    h = _mm256_xor_si256(j, h);
    h = _mm256_xor_si256(k, h);
    h = _mm256_mullo_epi32(_mm256_mullo_epi32(_mm256_mullo_epi32(h, h), _mm256_set1_epi32(60493)), h);
    return _mm256_xor_si256(_mm256_srai_epi32(h, 13), h);
}

AVX2 static inline __m256 grad3d_v(int32_t seed, __m256i i, __m256i j, __m256i k,
                                   __m256 x, __m256 y, __m256 z)
{
    __m256i h = hash3d_v(seed, i, j, k);
    __m256i h13 = _mm256_and_si256(h, _mm256_set1_epi32(13));
    __m256 l8 = _mm256_castsi256_ps(_mm256_cmpgt_epi32(_mm256_set1_epi32(8), h13));
    __m256 u = _mm256_blendv_ps(y, x, l8);
    __m256 l2 = _mm256_castsi256_ps(_mm256_cmpgt_epi32(_mm256_set1_epi32(2), h13));
    __m256 h12 = _mm256_castsi256_ps(_mm256_cmpeq_epi32(_mm256_set1_epi32(12), h13));
    __m256 v = _mm256_blendv_ps(_mm256_blendv_ps(z, x, h12), y, l2);
    __m256 h1 = _mm256_castsi256_ps(_mm256_slli_epi32(h, 31));
    __m256 h2 = _mm256_castsi256_ps(_mm256_slli_epi32(_mm256_and_si256(h, _mm256_set1_epi32(2)), 30));
    return _mm256_add_ps(_mm256_xor_ps(u, h1), _mm256_xor_ps(v, h2));
}

AVX2 static inline __m256 simplex_3d_v(__m256 x, __m256 y, __m256 z, int32_t seed)
{
    const __m256 vF3 = _mm256_set1_ps(F3), vG3 = _mm256_set1_ps(G3), vG33 = _mm256_set1_ps(G33);
    const __m256 neg1 = _mm256_set1_ps(-1.0f), zero = _mm256_setzero_ps(), c06 = _mm256_set1_ps(0.6f);
    const __m256i XP = _mm256_set1_epi32((int)X_PRIME), YP = _mm256_set1_epi32((int)Y_PRIME),
                  ZP = _mm256_set1_epi32((int)Z_PRIME);
    const __m256 ones = _mm256_castsi256_ps(_mm256_set1_epi32(-1));

    __m256 f = _mm256_mul_ps(vF3, _mm256_add_ps(_mm256_add_ps(x, y), z));
    __m256 x0 = _mm256_floor_ps(_mm256_add_ps(x, f));
    __m256 y0 = _mm256_floor_ps(_mm256_add_ps(y, f));
    __m256 z0 = _mm256_floor_ps(_mm256_add_ps(z, f));

    __m256i i = _mm256_mullo_epi32(_mm256_cvtps_epi32(x0), XP);
    __m256i j = _mm256_mullo_epi32(_mm256_cvtps_epi32(y0), YP);
    __m256i k = _mm256_mullo_epi32(_mm256_cvtps_epi32(z0), ZP);

    __m256 g = _mm256_mul_ps(vG3, _mm256_add_ps(_mm256_add_ps(x0, y0), z0));
    x0 = _mm256_sub_ps(x, _mm256_sub_ps(x0, g));
    y0 = _mm256_sub_ps(y, _mm256_sub_ps(y0, g));
    z0 = _mm256_sub_ps(z, _mm256_sub_ps(z0, g));

    __m256 xy = _mm256_cmp_ps(x0, y0, _CMP_GE_OQ);
    __m256 yz = _mm256_cmp_ps(y0, z0, _CMP_GE_OQ);
    __m256 xz = _mm256_cmp_ps(x0, z0, _CMP_GE_OQ);

    __m256 i1 = _mm256_and_ps(xy, xz);
    __m256 j1 = _mm256_andnot_ps(xy, yz);
    __m256 k1 = _mm256_andnot_ps(xz, _mm256_andnot_ps(yz, ones));
    __m256 i2 = _mm256_or_ps(xy, xz);
    __m256 j2 = _mm256_or_ps(_mm256_andnot_ps(xy, ones), yz);
    __m256 k2 = _mm256_andnot_ps(_mm256_and_ps(xz, yz), ones);

    __m256 x1 = _mm256_add_ps(_mm256_add_ps(x0, _mm256_and_ps(neg1, i1)), vG3);
    __m256 y1 = _mm256_add_ps(_mm256_add_ps(y0, _mm256_and_ps(neg1, j1)), vG3);
    __m256 z1 = _mm256_add_ps(_mm256_add_ps(z0, _mm256_and_ps(neg1, k1)), vG3);
    __m256 x2 = _mm256_add_ps(_mm256_add_ps(x0, _mm256_and_ps(neg1, i2)), vF3);
    __m256 y2 = _mm256_add_ps(_mm256_add_ps(y0, _mm256_and_ps(neg1, j2)), vF3);
    __m256 z2 = _mm256_add_ps(_mm256_add_ps(z0, _mm256_and_ps(neg1, k2)), vF3);
    __m256 x3 = _mm256_add_ps(x0, vG33);
    __m256 y3 = _mm256_add_ps(y0, vG33);
    __m256 z3 = _mm256_add_ps(z0, vG33);

#define TN(a, b, c) _mm256_sub_ps(_mm256_sub_ps(_mm256_sub_ps(c06, _mm256_mul_ps(a, a)), _mm256_mul_ps(b, b)), _mm256_mul_ps(c, c))
    __m256 t0 = TN(x0, y0, z0), t1 = TN(x1, y1, z1), t2 = TN(x2, y2, z2), t3 = TN(x3, y3, z3);
#undef TN
    t0 = _mm256_and_ps(t0, _mm256_cmp_ps(t0, zero, _CMP_GE_OQ));
    t1 = _mm256_and_ps(t1, _mm256_cmp_ps(t1, zero, _CMP_GE_OQ));
    t2 = _mm256_and_ps(t2, _mm256_cmp_ps(t2, zero, _CMP_GE_OQ));
    t3 = _mm256_and_ps(t3, _mm256_cmp_ps(t3, zero, _CMP_GE_OQ));
    t0 = _mm256_mul_ps(t0, t0); t1 = _mm256_mul_ps(t1, t1); t2 = _mm256_mul_ps(t2, t2); t3 = _mm256_mul_ps(t3, t3);
    t0 = _mm256_mul_ps(t0, t0); t1 = _mm256_mul_ps(t1, t1); t2 = _mm256_mul_ps(t2, t2); t3 = _mm256_mul_ps(t3, t3);

    __m256 v0 = _mm256_mul_ps(t0, grad3d_v(seed, i, j, k, x0, y0, z0));
    __m256i i1i = _mm256_add_epi32(i, _mm256_and_si256(_mm256_castps_si256(i1), XP));
    __m256i j1i = _mm256_add_epi32(j, _mm256_and_si256(_mm256_castps_si256(j1), YP));
    __m256i k1i = _mm256_add_epi32(k, _mm256_and_si256(_mm256_castps_si256(k1), ZP));
    __m256 v1 = _mm256_mul_ps(t1, grad3d_v(seed, i1i, j1i, k1i, x1, y1, z1));
    __m256i i2i = _mm256_add_epi32(i, _mm256_and_si256(_mm256_castps_si256(i2), XP));
    __m256i j2i = _mm256_add_epi32(j, _mm256_and_si256(_mm256_castps_si256(j2), YP));
    __m256i k2i = _mm256_add_epi32(k, _mm256_and_si256(_mm256_castps_si256(k2), ZP));
    __m256 v2 = _mm256_mul_ps(t2, grad3d_v(seed, i2i, j2i, k2i, x2, y2, z2));
    __m256 v3 = _mm256_mul_ps(t3, grad3d_v(seed, _mm256_add_epi32(i, XP), _mm256_add_epi32(j, YP),
                                           _mm256_add_epi32(k, ZP), x3, y3, z3));
    __m256 p1 = _mm256_add_ps(v3, v2);
    __m256 p2 = _mm256_add_ps(p1, v1);
    return _mm256_add_ps(p2, v0);
}

AVX2 static void noise_chunk_avx2(const int32_t coord[3], const vxo_noise_params *p, float *out)
{
    float sx, sy, sz;
    chunk_origin(coord, &sx, &sy, &sz);
    const __m256 freq = _mm256_set1_ps(p->freq), lac = _mm256_set1_ps(p->lacunarity),
                 gain = _mm256_set1_ps(p->gain), scale = _mm256_set1_ps(p->simplex_scale);
    float xa[8];
    for (int i = 0; i < 8; ++i) xa[i] = sx + (float)i;
    size_t idx = 0;
    __m256 z = _mm256_set1_ps(sz);
    for (int zz = 0; zz < VXO_CHUNK; ++zz) {
        __m256 y = _mm256_set1_ps(sy);
        for (int yy = 0; yy < VXO_CHUNK; ++yy) {
            __m256 x = _mm256_loadu_ps(xa);
            for (int xx = 0; xx < VXO_CHUNK / 8; ++xx) {
                __m256 fx = _mm256_mul_ps(x, freq), fy = _mm256_mul_ps(y, freq), fz = _mm256_mul_ps(z, freq);
                __m256 amp = _mm256_set1_ps(1.0f);
                __m256 r = _mm256_mul_ps(simplex_3d_v(fx, fy, fz, p->seed), scale);
                for (int32_t o = 1; o < p->octaves; ++o) {
                    fx = _mm256_mul_ps(fx, lac);
                    fy = _mm256_mul_ps(fy, lac);
                    fz = _mm256_mul_ps(fz, lac);
                    amp = _mm256_mul_ps(amp, gain);
                    r = _mm256_add_ps(_mm256_mul_ps(_mm256_mul_ps(simplex_3d_v(fx, fy, fz, p->seed), scale), amp), r);
                }
                _mm256_storeu_ps(out + idx, r);
                idx += 8;
                x = _mm256_add_ps(x, _mm256_set1_ps(8.0f));
            }
            y = _mm256_add_ps(y, _mm256_set1_ps(1.0f));
        }
        z = _mm256_add_ps(z, _mm256_set1_ps(1.0f));
    }
}
#endif

int vxo_have_avx2(void)
{
#if VXO_X86
    return __builtin_cpu_supports("avx2") ? 1 : 0;
#else
    return 0;
#endif
}

void vxo_noise_chunk_simd(const int32_t coord[3], const vxo_noise_params *p, float *out)
{
#if VXO_X86
    /* chunk.rs:29-33 picks avx2 when detected (sse41 otherwise; same lanes) */
    if (vxo_have_avx2()) {
        noise_chunk_avx2(coord, p, out);
        return;
    }
#endif
    vxo_noise_chunk(coord, p, out);
}

/* ------------------------------------------------------------------------ */
/* Terrain threshold, chunk.rs:27-46                                         */
/* ------------------------------------------------------------------------ */
void vxo_terrain_from_noise(const float *noise, uint8_t *cubes)
{
    /* chunk.rs:35-45: noise[idx] < 0. -> GRASS else AIR, pushed in idx order */
    for (size_t i = 0; i < VXO_VOXELS; ++i)
        cubes[i] = (noise[i] < 0.0f) ? VXO_GRASS : VXO_AIR;
}

void vxo_generate_terrain(const int32_t coord[3], const vxo_noise_params *p, uint8_t *cubes)
{
    float *noise = (float *)malloc(sizeof(float) * VXO_VOXELS);
    vxo_noise_chunk_simd(coord, p, noise);
    vxo_terrain_from_noise(noise, cubes);
    free(noise);
}

/* ------------------------------------------------------------------------ */
/* Mesher, chunk.rs:59-245                                                   */
/* ------------------------------------------------------------------------ */
/* texture.rs:238-243 */
static const uint32_t TEXTURE_INDEX_LIST[2][6] = {
    {0, 0, 0, 0, 2, 1},
    {1, 1, 1, 1, 1, 1},
};

#define CIDX(x, y, z) ((size_t)(x) + 32u * (size_t)(z) + 1024u * (size_t)(y))

/* adj[0..3] as in the reference; adj[4] (-y) / adj[5] (+y) are the EXTENSION of SURVEY 8(f4): NULL (always, for
 * the reference's own behaviour) emits every y-border face like chunk.rs:216-222,234-240 do. */
static size_t cull_impl(const int32_t coord[3], const uint8_t *cubes,
                        const uint8_t *const adj[6], vxo_quad *out, size_t cap)
{
    /* coordinate.rs:60-65 */
    const int32_t wx = (int32_t)((uint32_t)coord[0] * 32u);
    const int32_t wy = (int32_t)((uint32_t)coord[1] * 32u);
    const int32_t wz = (int32_t)((uint32_t)coord[2] * 32u);
    size_t n = 0;
#define PUSH(face_, slot_)                                                          \
    do {                                                                          \
        if (n < cap) {                                                            \
            out[n].x = px; out[n].y = py; out[n].z = pz;                          \
            out[n].face = (face_); out[n].tex = tex[(slot_)];                       \
        }                                                                         \
        ++n;                                                                      \
    } while (0)

    for (int y = 0; y < 32; ++y) {
        const int32_t py = y + wy;
        for (int z = 0; z < 32; ++z) {
            const int32_t pz = z + wz;
            for (int x = 0; x < 32; ++x) {
                const uint8_t c = cubes[CIDX(x, y, z)];
                if (c == VXO_AIR) continue;                      /* :71-73 */
                const uint32_t *tex = TEXTURE_INDEX_LIST[c];     /* :74-75 */
                const int32_t px = x + wx;
                /* LEFT :77-109 */
                if (x > 0) {
                    if (cubes[CIDX(x - 1, y, z)] == VXO_AIR) PUSH(VXO_LEFT, 0);
                } else if (adj[0]) {
                    if (adj[0][CIDX(31, y, z)] == VXO_AIR) PUSH(VXO_LEFT, 0);
                } else PUSH(VXO_LEFT, 0);
                /* RIGHT :110-140 */
                if (x < 31) {
                    if (cubes[CIDX(x + 1, y, z)] == VXO_AIR) PUSH(VXO_RIGHT, 1);
                } else if (adj[1]) {
                    if (adj[1][CIDX(0, y, z)] == VXO_AIR) PUSH(VXO_RIGHT, 1);
                } else PUSH(VXO_RIGHT, 1);
                /* BACK :141-173 */
                if (z > 0) {
                    if (cubes[CIDX(x, y, z - 1)] == VXO_AIR) PUSH(VXO_BACK, 2);
                } else if (adj[2]) {
                    if (adj[2][CIDX(x, y, 31)] == VXO_AIR) PUSH(VXO_BACK, 2);
                } else PUSH(VXO_BACK, 2);
                /* FRONT :174-204 */
                if (z < 31) {
                    if (cubes[CIDX(x, y, z + 1)] == VXO_AIR) PUSH(VXO_FRONT, 3);
                } else if (adj[3]) {
                    if (adj[3][CIDX(x, y, 0)] == VXO_AIR) PUSH(VXO_FRONT, 3);
                } else PUSH(VXO_FRONT, 3);
                /* DOWN :205-222 -- tex slot 5; y border always emits (adj[4] is NULL for the reference) */
                if (y > 0) {
                    if (cubes[CIDX(x, y - 1, z)] == VXO_AIR) PUSH(VXO_DOWN, 5);
                } else if (adj[4]) {
                    if (adj[4][CIDX(x, 31, z)] == VXO_AIR) PUSH(VXO_DOWN, 5);     /* extension: same rule as x / z */
                } else PUSH(VXO_DOWN, 5);
                /* UP :223-240 -- tex slot 4; y border always emits (adj[5] is NULL for the reference) */
                if (y < 31) {
                    if (cubes[CIDX(x, y + 1, z)] == VXO_AIR) PUSH(VXO_UP, 4);
                } else if (adj[5]) {
                    if (adj[5][CIDX(x, 0, z)] == VXO_AIR) PUSH(VXO_UP, 4);        /* extension */
                } else PUSH(VXO_UP, 4);
            }
        }
    }
#undef PUSH
    return n;
}

size_t vxo_cull_unseen_triangles(const int32_t coord[3], const uint8_t *cubes,
                                 const uint8_t *const adj[4], vxo_quad *out, size_t cap)
{
    const uint8_t *a6[6] = {adj[0], adj[1], adj[2], adj[3], NULL, NULL};   /* chunk.rs:205-240: never consulted */
    return cull_impl(coord, cubes, a6, out, cap);
}

/* EXTENSION (no reference behaviour to restate): vertical neighbours consulted like x / z ones */
size_t vxo_cull_unseen_triangles_ex(const int32_t coord[3], const uint8_t *cubes,
                                    const uint8_t *const adj[6], vxo_quad *out, size_t cap)
{
    return cull_impl(coord, cubes, adj, out, cap);
}

/* EXTENSION: the DIRT rule -- a solid voxel whose +y neighbour voxel is solid becomes DIRT (cube.rs:19 defines the
 * type, texture.rs:242 its textures; nothing in the reference generates it).  `above` = the +y chunk or NULL. */
void vxo_dirt_rule(uint8_t *cubes, const uint8_t *above)
{
    for (int y = 0; y < 32; ++y)
        for (int z = 0; z < 32; ++z)
            for (int x = 0; x < 32; ++x) {
                if (cubes[CIDX(x, y, z)] == VXO_AIR) continue;
                const int covered = (y < 31) ? cubes[CIDX(x, y + 1, z)] != VXO_AIR
                                             : (above && above[CIDX(x, 0, z)] != VXO_AIR);
                cubes[CIDX(x, y, z)] = covered ? VXO_DIRT : VXO_GRASS;
            }
}

/* ------------------------------------------------------------------------ */
/* Quad -> InstanceRaw, quad.rs:14-54                                        */
/* ------------------------------------------------------------------------ */
/* nalgebra 0.29 Rotation3::from_euler_angles(roll, pitch, yaw) = Rz*Ry*Rx */
static void from_euler_angles(float roll, float pitch, float yaw, float R[3][3])
{
    float sr = sinf(roll), cr = cosf(roll);
    float sp = sinf(pitch), cp = cosf(pitch);
    float sy = sinf(yaw), cy = cosf(yaw);
    R[0][0] = cy * cp; R[0][1] = cy * sp * sr - sy * cr; R[0][2] = cy * sp * cr + sy * sr;
    R[1][0] = sy * cp; R[1][1] = sy * sp * sr + cy * cr; R[1][2] = sy * sp * cr - cy * sr;
    R[2][0] = -sp;     R[2][1] = cp * sr;                R[2][2] = cp * cr;
}

/* f32::to_radians: deg * (PI_f32 / 180_f32) */
static float to_radians(float deg)
{
    const float pi = 3.14159265358979323846264338327950288f;
    return deg * (pi / 180.0f);
}

void vxo_face_rotation(uint32_t face, float R[3][3])
{
    switch (face) {
    case VXO_UP:    from_euler_angles(to_radians(-90.0f), 0.0f, 0.0f, R); break;   /* :17-20 */
    case VXO_DOWN:  from_euler_angles(to_radians(90.0f), 0.0f, 0.0f, R); break;    /* :21-24 */
    case VXO_LEFT:  from_euler_angles(0.0f, to_radians(-90.0f), 0.0f, R); break;   /* :25-28 */
    case VXO_RIGHT: from_euler_angles(0.0f, to_radians(90.0f), 0.0f, R); break;    /* :29-32 */
    case VXO_FRONT: from_euler_angles(0.0f, 0.0f, 0.0f, R); break;                 /* :33-35 */
    default:        from_euler_angles(0.0f, to_radians(180.0f), 0.0f, R); break;   /* BACK :36-39 */
    }
}

void vxo_quad_to_raw(const vxo_quad *q, vxo_instance_raw *out)
{
    /* quad.rs:48-49: position.to_homogeneous() * rotation.to_homogeneous(),
     * `.into()` yields column arrays.  Signed zeros are canonicalised to +0
     * (they depend on nalgebra's gemm summation order and are not pinned). */
    float R[3][3];
    vxo_face_rotation(q->face, R);
    float T[4][4] = {{1, 0, 0, (float)q->x}, {0, 1, 0, (float)q->y}, {0, 0, 1, (float)q->z}, {0, 0, 0, 1}};
    float H[4][4] = {{R[0][0], R[0][1], R[0][2], 0}, {R[1][0], R[1][1], R[1][2], 0},
                     {R[2][0], R[2][1], R[2][2], 0}, {0, 0, 0, 1}};
    for (int c = 0; c < 4; ++c)
        for (int r = 0; r < 4; ++r) {
            float s = T[r][0] * H[0][c];
            s = s + T[r][1] * H[1][c];
            s = s + T[r][2] * H[2][c];
            s = s + T[r][3] * H[3][c];
            out->t_matrix[c][r] = s + 0.0f;
        }
    out->texture_index = q->tex;
}

static size_t create_mesh_impl(const int32_t coord[3], const uint8_t *cubes,
                               const uint8_t *const adj[6], vxo_instance_raw *out, size_t cap)
{
    vxo_quad *faces = (vxo_quad *)malloc(sizeof(vxo_quad) * 98304);
    size_t n = cull_impl(coord, cubes, adj, faces, 98304);
    size_t m = n < cap ? n : cap;
    for (size_t i = 0; i < m; ++i) vxo_quad_to_raw(&faces[i], &out[i]);   /* chunk.rs:255 */
    free(faces);
    return n;
}

size_t vxo_create_mesh(const int32_t coord[3], const uint8_t *cubes,
                       const uint8_t *const adj[4], vxo_instance_raw *out, size_t cap)
{
    const uint8_t *a6[6] = {adj[0], adj[1], adj[2], adj[3], NULL, NULL};
    return create_mesh_impl(coord, cubes, a6, out, cap);
}

size_t vxo_create_mesh_ex(const int32_t coord[3], const uint8_t *cubes,
                          const uint8_t *const adj[6], vxo_instance_raw *out, size_t cap)
{
    return create_mesh_impl(coord, cubes, adj, out, cap);
}

/* ------------------------------------------------------------------------ */
/* Frustum, frustum_culling.rs                                               */
/* ------------------------------------------------------------------------ */
void vxo_matrix_to_planes(const float m[16], float planes[24])
{
    /* frustum_culling.rs:47-85; data[c][r] = m[c*4 + r] */
#define D(c, r) m[(c) * 4 + (r)]
    float left[4], right[4], top[4], bottom[4], nearp[4], farp[4];
    for (int c = 0; c < 4; ++c) {
        left[c]   = D(c, 3) + D(c, 0);   /* :55-58 */
        right[c]  = D(c, 3) - D(c, 0);   /* :60-63 */
        top[c]    = D(c, 3) - D(c, 1);   /* :65-68 */
        bottom[c] = D(c, 3) + D(c, 1);   /* :70-73 */
        nearp[c]  = D(c, 3) + D(c, 2);   /* :75-78 */
        farp[c]   = D(c, 3) - D(c, 2);   /* :80-83 */
    }
#undef D
    const float *order[6] = {nearp, farp, left, right, top, bottom};   /* :84 */
    for (int i = 0; i < 6; ++i) memcpy(planes + 4 * i, order[i], 16);
}

int vxo_frustum_contains(const float planes[24], const int32_t coord[3])
{
    /* coordinate.rs:53-58 */
    const float px = (float)(int32_t)((uint32_t)coord[0] * 32u);
    const float py = (float)(int32_t)((uint32_t)coord[1] * 32u);
    const float pz = (float)(int32_t)((uint32_t)coord[2] * 32u);
    /* frustum_culling.rs:19-26 corner order */
    const float ex[8] = {px, px + 32.f, px, px + 32.f, px, px + 32.f, px, px + 32.f};
    const float ey[8] = {py, py, py, py, py + 32.f, py + 32.f, py + 32.f, py + 32.f};
    const float ez[8] = {pz, pz, pz + 32.f, pz + 32.f, pz, pz, pz + 32.f, pz + 32.f};
    for (int p = 0; p < 8; ++p) {
        int inside = 1;
        for (int i = 0; i < 6; ++i) {
            const float *pl = planes + 4 * i;
            float dist = pl[0] * ex[p] + pl[1] * ey[p] + pl[2] * ez[p] + pl[3];   /* :29-32 */
            if (dist < 0.0f) { inside = 0; break; }                               /* :33-35 */
        }
        if (inside) return 1;                                                     /* :37 */
    }
    return 0;
}

void vxo_frustum_cull(const float *planes, size_t n_frames, const int32_t *coords, size_t n,
                      uint8_t *out)
{
    for (size_t f = 0; f < n_frames; ++f)
        for (size_t i = 0; i < n; ++i)
            out[f * n + i] = (uint8_t)vxo_frustum_contains(planes + 24 * f, coords + 3 * i);
}

/* ------------------------------------------------------------------------ */
/* Multithreaded batch driver = the worker closure of chunk_builder.rs:80-84 */
/* run over a coordinate list by a pool of nthreads (uvth default num_cpus). */
/* ------------------------------------------------------------------------ */
typedef struct {
    const int32_t *coords;
    size_t n;
    const vxo_noise_params *p;
    const uint8_t *nbr_mask;
    const int32_t *nbr_idx;      /* resolved [n][NB] */
    int nb;                      /* 4 (reference) or 6 (extension: -y, +y appended) */
    int dirt_rule;               /* extension */
    int use_simd;
    uint8_t *cubes;              /* [n][32768] */
    uint32_t *counts;
    vxo_instance_raw **per_chunk;   /* malloc'd Vec<InstanceRaw> per chunk, or NULL */
    int keep;
    int do_terrain;
    atomic_size_t next_a, next_b;
    pthread_barrier_t bar;
    atomic_int failed;
} batch_t;

static void *batch_worker(void *arg)
{
    batch_t *b = (batch_t *)arg;
    float *noise = (float *)malloc(sizeof(float) * VXO_VOXELS);
    vxo_instance_raw *tmp = (vxo_instance_raw *)malloc(sizeof(vxo_instance_raw) * 98304);
    if (!noise || !tmp) atomic_store(&b->failed, 1);
    /* phase A: Chunk::new (chunk.rs:21-25) */
    if (b->do_terrain && !atomic_load(&b->failed)) {
        for (;;) {
            size_t i = atomic_fetch_add(&b->next_a, 1);
            if (i >= b->n) break;
            if (b->use_simd) vxo_noise_chunk_simd(b->coords + 3 * i, b->p, noise);
            else vxo_noise_chunk(b->coords + 3 * i, b->p, noise);
            vxo_terrain_from_noise(noise, b->cubes + i * VXO_VOXELS);
        }
    }
    int serial = pthread_barrier_wait(&b->bar);
    if (b->dirt_rule) {
        /* extension: types from solidity (which no thread changes), by one thread */
        if (serial == PTHREAD_BARRIER_SERIAL_THREAD && !atomic_load(&b->failed))
            for (size_t i = 0; i < b->n; ++i) {
                int32_t j = (b->nb == 6) ? b->nbr_idx[6 * i + 5] : -1;
                vxo_dirt_rule(b->cubes + i * VXO_VOXELS, j >= 0 ? b->cubes + (size_t)j * VXO_VOXELS : NULL);
            }
        pthread_barrier_wait(&b->bar);
    }
    /* phase B: create_mesh (chunk.rs:48-57) */
    if (!atomic_load(&b->failed)) {
        for (;;) {
            size_t i = atomic_fetch_add(&b->next_b, 1);
            if (i >= b->n) break;
            const uint8_t *adj[6] = {NULL, NULL, NULL, NULL, NULL, NULL};
            for (int k = 0; k < b->nb; ++k) {
                int32_t j = b->nbr_idx[(size_t)b->nb * i + k];
                adj[k] = (j >= 0) ? b->cubes + (size_t)j * VXO_VOXELS : NULL;
            }
            size_t q = create_mesh_impl(b->coords + 3 * i, b->cubes + i * VXO_VOXELS, adj, tmp, 98304);
            b->counts[i] = (uint32_t)q;
            if (b->keep) {
                vxo_instance_raw *v = (vxo_instance_raw *)malloc(sizeof(vxo_instance_raw) * (q ? q : 1));
                if (!v) { atomic_store(&b->failed, 1); break; }
                memcpy(v, tmp, sizeof(vxo_instance_raw) * q);
                b->per_chunk[i] = v;
            }
        }
    }
    free(noise);
    free(tmp);
    return NULL;
}

/* open-addressing map coord -> first index */
static uint64_t coord_hash(const int32_t *c)
{
    uint64_t h = (uint64_t)(uint32_t)c[0] * 0x9E3779B97F4A7C15ull;
    h ^= ((uint64_t)(uint32_t)c[1] + 0x7F4A7C15ull) * 0xC2B2AE3D27D4EB4Full;
    h ^= ((uint64_t)(uint32_t)c[2] + 0x165667B1ull) * 0xD6E8FEB86659FD93ull;
    h ^= h >> 29;
    return h;
}

static int32_t *resolve_neighbours(const int32_t *coords, size_t n, const uint8_t *nbr_mask, int nb)
{
    size_t cap = 16;
    while (cap < 2 * n + 2) cap <<= 1;
    int32_t *table = (int32_t *)malloc(sizeof(int32_t) * cap);
    int32_t *idx = (int32_t *)malloc(sizeof(int32_t) * (size_t)nb * (n ? n : 1));
    if (!table || !idx) { free(table); free(idx); return NULL; }
    for (size_t i = 0; i < cap; ++i) table[i] = -1;
    for (size_t i = 0; i < n; ++i) {
        size_t s = coord_hash(coords + 3 * i) & (cap - 1);
        for (;;) {
            if (table[s] < 0) { table[s] = (int32_t)i; break; }
            if (memcmp(coords + 3 * table[s], coords + 3 * i, 12) == 0) break;
            s = (s + 1) & (cap - 1);
        }
    }
    static const int dx[6] = {-1, 1, 0, 0, 0, 0}, dz[6] = {0, 0, -1, 1, 0, 0};   /* chunk_builder.rs:134-169 */
    static const int dy[6] = {0, 0, 0, 0, -1, 1};                                /* extension slots 4, 5 */
    for (size_t i = 0; i < n; ++i)
        for (int k = 0; k < nb; ++k) {
            int32_t q[3] = {(int32_t)((uint32_t)coords[3 * i] + (uint32_t)dx[k]),
                            (int32_t)((uint32_t)coords[3 * i + 1] + (uint32_t)dy[k]),
                            (int32_t)((uint32_t)coords[3 * i + 2] + (uint32_t)dz[k])};
            int32_t found = -1;
            if (!nbr_mask || (nbr_mask[i] >> k) & 1) {
                size_t s = coord_hash(q) & (cap - 1);
                while (table[s] >= 0) {
                    if (memcmp(coords + 3 * table[s], q, 12) == 0) { found = table[s]; break; }
                    s = (s + 1) & (cap - 1);
                }
            }
            idx[(size_t)nb * i + k] = found;
        }
    free(table);
    return idx;
}

static size_t run_batch(batch_t *b, int nthreads, uint64_t *offsets, vxo_instance_raw *inst, size_t cap)
{
    if (nthreads < 1) nthreads = 1;
    if ((size_t)nthreads > b->n && b->n > 0) nthreads = (int)b->n;
    atomic_init(&b->next_a, 0);
    atomic_init(&b->next_b, 0);
    atomic_init(&b->failed, 0);
    pthread_barrier_init(&b->bar, NULL, (unsigned)nthreads);
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)nthreads);
    for (int t = 0; t < nthreads; ++t) pthread_create(&th[t], NULL, batch_worker, b);
    for (int t = 0; t < nthreads; ++t) pthread_join(th[t], NULL);
    pthread_barrier_destroy(&b->bar);
    free(th);
    if (atomic_load(&b->failed)) return (size_t)-1;
    size_t total = 0;
    for (size_t i = 0; i < b->n; ++i) {
        if (offsets) offsets[i] = total;
        total += b->counts[i];
    }
    if (offsets) offsets[b->n] = total;
    size_t ret = total;
    if (b->keep) {
        if (inst && total > cap) ret = (size_t)-1;
        size_t off = 0;
        for (size_t i = 0; i < b->n; ++i) {
            if (inst && ret != (size_t)-1)
                memcpy(inst + off, b->per_chunk[i], sizeof(vxo_instance_raw) * b->counts[i]);
            off += b->counts[i];
            free(b->per_chunk[i]);
        }
    }
    return ret;
}

size_t vxo_gen_mesh_batch(const int32_t *coords, size_t n, const vxo_noise_params *p,
                          const uint8_t *nbr_mask, int nthreads, int use_simd,
                          uint32_t *counts, uint64_t *offsets,
                          vxo_instance_raw *inst, size_t cap, uint8_t *cubes_out)
{
    batch_t b;
    memset(&b, 0, sizeof b);
    b.coords = coords; b.n = n; b.p = p; b.nbr_mask = nbr_mask; b.use_simd = use_simd;
    b.counts = counts; b.do_terrain = 1; b.keep = 1; b.nb = 4;
    uint8_t *cubes = cubes_out ? cubes_out : (uint8_t *)malloc((size_t)VXO_VOXELS * (n ? n : 1));
    int32_t *idx = resolve_neighbours(coords, n, nbr_mask, 4);
    vxo_instance_raw **pc = (vxo_instance_raw **)calloc(n ? n : 1, sizeof(*pc));
    size_t ret = (size_t)-1;
    if (cubes && idx && pc) {
        b.cubes = cubes; b.nbr_idx = idx; b.per_chunk = pc;
        ret = run_batch(&b, nthreads, offsets, inst, cap);
    }
    if (!cubes_out) free(cubes);
    free(idx);
    free(pc);
    return ret;
}

/* EXTENSION: the same work function with vertical neighbours (mask bits 4 = -y, 5 = +y) and / or
 * the DIRT rule.  vertical = 0 and dirt_rule = 0 is vxo_gen_mesh_batch. */
size_t vxo_gen_mesh_batch_ex(const int32_t *coords, size_t n, const vxo_noise_params *p,
                             const uint8_t *nbr_mask, int vertical, int dirt_rule, int nthreads, int use_simd,
                             uint32_t *counts, uint64_t *offsets,
                             vxo_instance_raw *inst, size_t cap, uint8_t *cubes_out)
{
    batch_t b;
    memset(&b, 0, sizeof b);
    b.coords = coords; b.n = n; b.p = p; b.nbr_mask = nbr_mask; b.use_simd = use_simd;
    b.counts = counts; b.do_terrain = 1; b.keep = 1; b.nb = vertical ? 6 : 4; b.dirt_rule = dirt_rule;
    uint8_t *cubes = cubes_out ? cubes_out : (uint8_t *)malloc((size_t)VXO_VOXELS * (n ? n : 1));
    int32_t *idx = resolve_neighbours(coords, n, nbr_mask, b.nb);
    vxo_instance_raw **pc = (vxo_instance_raw **)calloc(n ? n : 1, sizeof(*pc));
    size_t ret = (size_t)-1;
    if (cubes && idx && pc) {
        b.cubes = cubes; b.nbr_idx = idx; b.per_chunk = pc;
        ret = run_batch(&b, nthreads, offsets, inst, cap);
    }
    if (!cubes_out) free(cubes);
    free(idx);
    free(pc);
    return ret;
}

size_t vxo_mesh_batch(const int32_t *coords, size_t n, const uint8_t *cubes,
                      const int32_t *nbr_idx, int nthreads,
                      uint32_t *counts, uint64_t *offsets,
                      vxo_instance_raw *inst, size_t cap)
{
    batch_t b;
    memset(&b, 0, sizeof b);
    b.coords = coords; b.n = n; b.counts = counts; b.do_terrain = 0; b.keep = 1; b.nb = 4;
    b.cubes = (uint8_t *)cubes; b.nbr_idx = nbr_idx;
    vxo_instance_raw **pc = (vxo_instance_raw **)calloc(n ? n : 1, sizeof(*pc));
    if (!pc) return (size_t)-1;
    b.per_chunk = pc;
    size_t ret = run_batch(&b, nthreads, offsets, inst, cap);
    free(pc);
    return ret;
}
