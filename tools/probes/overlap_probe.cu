// Does a small, late kernel get SM slots while a long grid that saturates the SMs is still dispatching?
//   A: many CTAs, each spins `spin_us`; occupies the SM through registers (64 x 256 x 4) or shared memory.
//   B: one CTA per SM, launched 1 ms later on another stream (same or high priority).
// Prints when B's first CTA started / its last CTA ended relative to A's first start and last end.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned long long gtime()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__global__ void __launch_bounds__(256, 4) kernA(unsigned long long *stamp, int spin_us, float *sink)
{
    extern __shared__ float sm[];
    const unsigned long long t0 = gtime();
    if (threadIdx.x == 0) atomicMin(stamp + 0, t0);
    float acc[52];
#pragma unroll
    for (int i = 0; i < 52; ++i) acc[i] = threadIdx.x * 0.001f + i;
    while (gtime() - t0 < (unsigned long long)spin_us * 1000ull) {
#pragma unroll
        for (int i = 0; i < 52; ++i) acc[i] = acc[i] * 1.0001f + 0.5f;
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 52; ++i) s += acc[i];
    if (s == 12345.678f) sink[0] = s + sm[0];
    if (threadIdx.x == 0) atomicMax(stamp + 1, gtime());
}

__global__ void __launch_bounds__(256) kernB(unsigned long long *stamp, int spin_us, float *sink)
{
    extern __shared__ float sm[];
    const unsigned long long t0 = gtime();
    if (threadIdx.x == 0) atomicMin(stamp + 2, t0);
    float b[34];
#pragma unroll
    for (int i = 0; i < 34; ++i) b[i] = threadIdx.x * 0.01f + i;
    while (gtime() - t0 < (unsigned long long)spin_us * 1000ull) {
#pragma unroll
        for (int i = 0; i < 34; ++i) b[i] = b[i] * 1.0001f + 0.5f;
    }
    float a = 0.f;
#pragma unroll
    for (int i = 0; i < 34; ++i) a += b[i];
    if (a == 12345.678f) sink[0] = a + sm[0];
    if (threadIdx.x == 0) atomicMax(stamp + 3, gtime());
}

int main(int argc, char **argv)
{
    const int smemA = argc > 1 ? atoi(argv[1]) : 30 * 1024;      // bytes per CTA of A
    const int smemB = argc > 2 ? atoi(argv[2]) : 48 * 1024;
    const int hi = argc > 3 ? atoi(argv[3]) : 1;                  // B on a high-priority stream
    const int carve = argc > 4 ? atoi(argv[4]) : 0;               // max shared carve-out for both
    unsigned long long *stamp, *stamp2, h[4];
    const int nB = argc > 5 ? atoi(argv[5]) : 8;
    float *sink;
    cudaMalloc(&stamp, 32);
    cudaMalloc(&stamp2, 32);
    cudaMalloc(&sink, 4);
    cudaFuncSetAttribute(kernA, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(kernB, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (carve) {
        cudaFuncSetAttribute(kernA, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        cudaFuncSetAttribute(kernB, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    }
    int lo_p, hi_p;
    cudaDeviceGetStreamPriorityRange(&lo_p, &hi_p);
    cudaStream_t sa, sb;
    cudaStreamCreateWithPriority(&sa, cudaStreamNonBlocking, lo_p);
    cudaStreamCreateWithPriority(&sb, cudaStreamNonBlocking, hi ? hi_p : lo_p);
    cudaFuncAttributes fa, fb;
    cudaFuncGetAttributes(&fa, kernA);
    cudaFuncGetAttributes(&fb, kernB);
    for (int rep = 0; rep < 2; ++rep) {
        h[0] = h[2] = ~0ull;
        h[1] = h[3] = 0;
        cudaMemcpy(stamp, h, 32, cudaMemcpyHostToDevice);
        kernA<<<148 * 4 * 50, 256, smemA, sa>>>(stamp, 20, sink);          // ~50 waves x 20 us = 1 ms
        // B becomes ready ~300 us later
        cudaEvent_t ev;
        cudaEventCreate(&ev);
        kernB<<<1, 32, 0, sb>>>(stamp2, 300, sink);                        // delay on B's stream
        kernB<<<148 * nB, 256, smemB, sb>>>(stamp, 20, sink);
        cudaDeviceSynchronize();
        cudaMemcpy(h, stamp, 32, cudaMemcpyDeviceToHost);
        if (rep)
            printf("A regs %d smem %d | B regs %d x%d CTAs/SM smem %d prio %s carve %d | A ran %.0f us; B first CTA start +%.0f us, last end +%.0f us after A's start; A end at +%.0f\n",
                   fa.numRegs, smemA, fb.numRegs, nB, smemB, hi ? "high" : "same", carve, (h[1] - h[0]) / 1e3, (double)(h[2] - h[0]) / 1e3,
                   (double)(h[3] - h[0]) / 1e3, (double)(h[1] - h[0]) / 1e3);
    }
    return 0;
}
