// kernels_peaks.cu -- measured non-tensor issue peaks for the roofline
// denominators (MEASURED_PEAKS.json has no FP32 figure): FP32 FFMA, FP32
// non-fused FMUL+FADD (what a bit-exact noise kernel may use), INT32 IMAD+LOP3.
#include "vx_common.cuh"

namespace {

constexpr int kIters = 4096;
constexpr int kChains = 16;

template <int KIND>
__global__ void __launch_bounds__(256) vx_peak_kernel(float *sink, float seedf, uint32_t seedi)
{
    float a[kChains];
    uint32_t h[kChains];
#pragma unroll
    for (int c = 0; c < kChains; ++c) {
        a[c] = seedf + (float)(threadIdx.x + c);
        h[c] = seedi + threadIdx.x * 8u + c;
    }
    const float m = 1.0000001f, b = seedf * 1e-7f;
#pragma unroll 1
    for (int it = 0; it < kIters; ++it) {
#pragma unroll
        for (int c = 0; c < kChains; ++c) {
            if (KIND == 0) {
                // immediate multiplicand (the 1-cycle FFMA form) + a register addend
                a[c] = __fmaf_rn(a[c], 1.0000001f, b);
                a[c] = __fmaf_rn(a[c], 0.9999999f, b);
            } else if (KIND == 1) {
                a[c] = __fmul_rn(a[c], m);
                a[c] = __fadd_rn(a[c], b);
            } else {
                h[c] = h[c] * 60493u + seedi;
                h[c] = (h[c] >> 13) ^ h[c];      // SHF + LOP3
            }
        }
    }
    float acc = 0.f;
#pragma unroll
    for (int c = 0; c < kChains; ++c) acc += a[c] + (float)h[c];
    if (acc == 12345.678f) sink[0] = acc;       // never true; keeps the chains alive
}

}  // namespace

cudaError_t vx_run_peak(int kind, int n_sm, double *tera_ops, cudaStream_t s, vx_counter *launches)
{
    float *sink = nullptr;
    cudaError_t e = cudaMalloc(&sink, 4);
    if (e != cudaSuccess) return e;
    cudaEvent_t t0, t1;
    cudaEventCreate(&t0);
    cudaEventCreate(&t1);
    const int grid = n_sm * 8, block = 256;
    double best = 0.0;
    for (int rep = 0; rep < 6; ++rep) {
        cudaEventRecord(t0, s);
        if (kind == 0) vx_peak_kernel<0><<<grid, block, 0, s>>>(sink, 1.0f, 1u);
        else if (kind == 1) vx_peak_kernel<1><<<grid, block, 0, s>>>(sink, 1.0f, 1u);
        else vx_peak_kernel<2><<<grid, block, 0, s>>>(sink, 1.0f, 1u);
        cudaEventRecord(t1, s);
        e = cudaEventSynchronize(t1);
        if (e != cudaSuccess) break;
        if (launches) launches->fetch_add(1, std::memory_order_relaxed);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, t0, t1);
        // per chain step: kind 0 = 2 FFMA = 4 flop; kind 1 = FMUL + FADD = 2 flop;
        // kind 2 = IMAD + SHF + LOP3 = 3 integer ops
        const double per = (kind == 0) ? 4.0 : (kind == 1 ? 2.0 : 3.0);
        const double ops = (double)grid * block * (double)kIters * kChains * per;
        if (rep > 0 && ms > 0.f) best = fmax(best, ops / (ms * 1e-3) / 1e12);
    }
    cudaEventDestroy(t0);
    cudaEventDestroy(t1);
    cudaFree(sink);
    *tera_ops = best;
    return e;
}
