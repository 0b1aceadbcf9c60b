// Micro-probe: issue rates of scalar vs packed (f32x2) FP32 instructions on sm_100a, alone and
// interleaved with ALU-pipe work.  Prints warp-instructions per clock per SM sub-partition.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o f32x2_probe f32x2_probe.cu && ./f32x2_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CHAINS 8
#define ITERS 2048

template <int MODE>
__global__ void __launch_bounds__(256) probe(float *out, float seed, long long *clk)
{
    float a[2 * CHAINS];
    uint32_t u[CHAINS];
#pragma unroll
    for (int i = 0; i < 2 * CHAINS; ++i) a[i] = seed + threadIdx.x * 1e-3f + i;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) u[i] = threadIdx.x * 2654435761u + i;
    const float m = seed * 0.999f, c = seed * 1e-3f;
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) {
            if (MODE == 0) {                 // scalar FFMA, 3 register operands: 2 per chain
                a[2 * i] = __fmaf_rn(a[2 * i], m, c);
                a[2 * i + 1] = __fmaf_rn(a[2 * i + 1], m, c);
            } else if (MODE == 1) {          // packed FFMA2: 1 per chain (same flops as MODE 0)
                float2 v = __ffma2_rn(make_float2(a[2 * i], a[2 * i + 1]), make_float2(m, m), make_float2(c, c));
                a[2 * i] = v.x; a[2 * i + 1] = v.y;
            } else if (MODE == 2) {          // scalar FADD + FMUL
                a[2 * i] = __fadd_rn(a[2 * i], c);
                a[2 * i + 1] = __fmul_rn(a[2 * i + 1], m);
            } else if (MODE == 3) {          // packed FADD2 then FMUL2 (2x the flops of MODE 2)
                float2 v = __fadd2_rn(make_float2(a[2 * i], a[2 * i + 1]), make_float2(c, c));
                v = __fmul2_rn(v, make_float2(m, m));
                a[2 * i] = v.x; a[2 * i + 1] = v.y;
            } else if (MODE == 4) {          // ALU only: 2 LOP3 per chain
                u[i] = (u[i] ^ (u[i] >> 7)) ;
                u[i] = (u[i] & 0x7ffffffu) | (u[i] << 3);
            } else if (MODE == 5) {          // 2 scalar FFMA + 2 ALU per chain
                a[2 * i] = __fmaf_rn(a[2 * i], m, c);
                a[2 * i + 1] = __fmaf_rn(a[2 * i + 1], m, c);
                u[i] = (u[i] ^ (u[i] >> 7));
                u[i] = (u[i] & 0x7ffffffu) | (u[i] << 3);
            } else if (MODE == 6) {          // 1 FFMA2 + 2 ALU per chain
                float2 v = __ffma2_rn(make_float2(a[2 * i], a[2 * i + 1]), make_float2(m, m), make_float2(c, c));
                a[2 * i] = v.x; a[2 * i + 1] = v.y;
                u[i] = (u[i] ^ (u[i] >> 7));
                u[i] = (u[i] & 0x7ffffffu) | (u[i] << 3);
            } else if (MODE == 7) {          // IMAD only: 2 per chain
                u[i] = u[i] * 60493u + 17u;
                u[i] = u[i] * u[i] + 3u;
            }
        }
    }
    const long long t1 = clock64();
    float s = 0.f; uint32_t x = 0;
#pragma unroll
    for (int i = 0; i < 2 * CHAINS; ++i) s += a[i];
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) x ^= u[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + (float)x;
    if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
}

template <int MODE>
void run(const char *name, int instr_per_chain, float *out, long long *clk)
{
    // 4 CTAs of 256 threads per SM -> 8 warps per sub-partition
    probe<MODE><<<148 * 4, 256>>>(out, 1.0001f, clk);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    probe<MODE><<<148 * 4, 256>>>(out, 1.0001f, clk);
    cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long h; cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
    const double winstr = 8.0 * ITERS * CHAINS * instr_per_chain;     // warp-instructions per sub-partition
    printf("%-44s %7.3f warp-instr/clk/SMSP (clock64)   %8.3f ms\n", name, winstr / (double)h, ms);
}

int main()
{
    float *out; long long *clk;
    cudaMalloc(&out, 148 * 4 * 256 * 4); cudaMalloc(&clk, 8);
    run<0>("FFMA x2 (scalar, 3-reg)", 2, out, clk);
    run<1>("FFMA2 x1 (packed)", 1, out, clk);
    run<2>("FADD + FMUL (scalar)", 2, out, clk);
    run<3>("FADD2 + FMUL2 (packed)", 2, out, clk);
    run<4>("LOP3/SHF x2 (ALU)", 2, out, clk);
    run<5>("2 FFMA + 2 ALU", 4, out, clk);
    run<6>("1 FFMA2 + 2 ALU", 3, out, clk);
    run<7>("IMAD x2", 2, out, clk);
    return 0;
}
