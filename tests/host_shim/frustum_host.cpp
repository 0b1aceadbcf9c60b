// Host build of the DEVICE frustum logic (csrc/vx_frustum.cuh) for CPU testing: the CUDA intrinsics are mapped to
// plain separately-rounded f32 operations (compile with -ffp-contract=off) and a warp has ONE lane, so a vote is the
// lane's own predicate.  Test infrastructure only.
#include <cmath>
#include <cstdint>
#include <cstring>
#define __device__
#define __forceinline__ inline
#define VX_FULL 0xffffffffu
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }
static inline uint32_t __float_as_uint(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline int __ffs(uint32_t v) { return v ? __builtin_ctz(v) + 1 : 0; }
static inline bool __any_sync(uint32_t, bool p) { return p; }
struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline float2 __fadd2_rn(float2 a, float2 b) { return float2{a.x + b.x, a.y + b.y}; }
// device fminf / fmaxf ignore a NaN operand, like the C library's
#include "../../wgpu-voxel-engine_b200/csrc/vx_frustum.cuh"

// visible[i] for n chunks against one plane set, through frame_test<UNROLL, 4> on groups of four chunks (the last
// group is padded with dead slots, as in the kernels)
template <int UNROLL>
static void run(const float *planes, const int32_t *coords, int n, uint8_t *visible)
{
    float4 pl[6];
    uint32_t cls = 0;
    for (int p = 0; p < 6; ++p) {
        pl[p] = float4{planes[4 * p], planes[4 * p + 1], planes[4 * p + 2], planes[4 * p + 3]};
        cls |= plane_class(pl[p]) << (4 * p);
    }
    for (int g = 0; g < n; g += 4) {
        Boxes<4> b;
        uint32_t live = 0;
        for (int k = 0; k < 4; ++k) {
            const bool in = g + k < n;
            set_box(b, k, in ? coords[3 * (g + k)] : 0, in ? coords[3 * (g + k) + 1] : 0, in ? coords[3 * (g + k) + 2] : 0);
            live |= (in ? 1u : 0u) << k;
        }
        const uint32_t vis = frame_test<UNROLL>(pl, cls, b, live);
        for (int k = 0; k < 4 && g + k < n; ++k) visible[g + k] = (vis >> k) & 1u;
    }
}

extern "C" void shim_frustum(const float *planes, const int32_t *coords, int n, uint8_t *visible, int unroll)
{
    if (unroll == 6) run<6>(planes, coords, n, visible);
    else if (unroll == 2) run<2>(planes, coords, n, visible);
    else run<1>(planes, coords, n, visible);
}

// the corner-by-corner test alone (what a straddling chunk gets)
extern "C" void shim_corner_test(const float *planes, const int32_t *coords, int n, uint8_t *visible)
{
    float4 pl[6];
    for (int p = 0; p < 6; ++p) pl[p] = float4{planes[4 * p], planes[4 * p + 1], planes[4 * p + 2], planes[4 * p + 3]};
    for (int i = 0; i < n; ++i) {
        Boxes<1> b;
        set_box(b, 0, coords[3 * i], coords[3 * i + 1], coords[3 * i + 2]);
        visible[i] = corner_test(pl, b.x0[0], b.x1[0], b.y0[0], b.y1[0], b.z0[0], b.z1[0]) ? 1 : 0;
    }
}

// share of the chunks the extreme corners decide without the corner test: the class of every plane set must be usable
extern "C" uint32_t shim_plane_classes(const float *planes)
{
    uint32_t cls = 0;
    for (int p = 0; p < 6; ++p)
        cls |= plane_class(float4{planes[4 * p], planes[4 * p + 1], planes[4 * p + 2], planes[4 * p + 3]}) << (4 * p);
    return cls;
}
