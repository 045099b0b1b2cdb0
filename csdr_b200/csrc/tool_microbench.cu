// tool_microbench.cu -- standalone probe of the FP32 issue rates that bound the FIR bank kernel on B200:
// scalar FFMA vs packed FFMA2 (vector and uniform-register operand), with and without an LDS.128 stream.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/microbench csdr_b200/csrc/tool_microbench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c)
{
    uint64_t ra = *reinterpret_cast<uint64_t*>(&a), rb = *reinterpret_cast<uint64_t*>(&b), rc = *reinterpret_cast<uint64_t*>(&c), rd;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
    return *reinterpret_cast<float2*>(&rd);
}
struct P { float2 h[32]; };

template <int MODE>
__global__ void __launch_bounds__(256) k(float2* out, int iters, const __grid_constant__ P p, long long* cyc)
{
    __shared__ float4 sm[1024];
    for (int i = threadIdx.x; i < 1024; i += 256) sm[i] = make_float4(1e-3f * i, 1.f, 0.5f, 0.25f);
    __syncthreads();
    float2 acc[16];
#pragma unroll
    for (int r = 0; r < 16; r++) acc[r] = make_float2(threadIdx.x * 1e-3f + r, 1.f);
    float2 a = make_float2(1.0001f, 0.9999f), b = make_float2(0.5f, 0.25f);
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
        if (MODE == 0) {            // scalar FFMA, 32 per step
#pragma unroll
            for (int r = 0; r < 16; r++) { acc[r].x = fmaf(acc[r].x, a.x, b.x); acc[r].y = fmaf(acc[r].y, a.y, b.y); }
        } else if (MODE == 1) {     // FFMA2 vector operands
#pragma unroll
            for (int r = 0; r < 16; r++) acc[r] = ffma2(acc[r], a, b);
        } else if (MODE == 2) {     // FFMA2 with uniform-register tap: acc = s*h + acc
#pragma unroll
            for (int r = 0; r < 16; r++) acc[r] = ffma2(a, p.h[r], acc[r]);
        } else if (MODE == 3) {     // FIR-like: 1 LDS.128 per 16 FFMA2 (two samples x 8 taps)
            float4 w = sm[(threadIdx.x * 5 + it) & 1023];
#pragma unroll
            for (int r = 0; r < 8; r++) { acc[r] = ffma2(make_float2(w.x, w.y), p.h[r], acc[r]); acc[r + 8] = ffma2(make_float2(w.z, w.w), p.h[r + 8], acc[r + 8]); }
        } else if (MODE == 4) {     // 1 LDS.128 per 8 FFMA2
            float4 w = sm[(threadIdx.x * 5 + it) & 1023];
#pragma unroll
            for (int r = 0; r < 4; r++) { acc[r] = ffma2(make_float2(w.x, w.y), p.h[r], acc[r]); acc[r + 8] = ffma2(make_float2(w.z, w.w), p.h[r + 8], acc[r + 8]); }
        }
    }
    long long t1 = clock64();
    float2 s = make_float2(0, 0);
#pragma unroll
    for (int r = 0; r < 16; r++) { s.x += acc[r].x; s.y += acc[r].y; }
    out[blockIdx.x * 256 + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int MODE>
void run(const char* name, int ctas_per_sm, double fma_per_iter_per_thread)
{
    int iters = 20000; P p; for (int i = 0; i < 32; i++) p.h[i] = make_float2(1.f / (i + 2), 1.f / (i + 2));
    float2* out; long long* cyc; cudaMalloc(&out, 148 * 8 * 256 * 8); cudaMalloc(&cyc, 8);
    int grid = 148 * ctas_per_sm;
    k<MODE><<<grid, 256>>>(out, 100, p, cyc); cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0); k<MODE><<<grid, 256>>>(out, iters, p, cyc); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    double fma = fma_per_iter_per_thread * iters * 256.0 * grid;
    printf("%-34s ctas/SM=%d  %.3f ms  %.1f TFLOP/s  %.1f FMA/clk/SM  (SM clk ~%.0f MHz)\n", name, ctas_per_sm, ms, 2 * fma / ms / 1e9,
           fma_per_iter_per_thread * iters * 256.0 * ctas_per_sm / (double)c, c / (ms * 1e3));
    cudaFree(out); cudaFree(cyc);
}

int main()
{
    for (int c = 1; c <= 4; c *= 2) {
        run<0>("FFMA scalar (32/step)", c, 32);
        run<1>("FFMA2 vector operands (16/step)", c, 32);
        run<2>("FFMA2 uniform-reg tap (16/step)", c, 32);
        run<3>("FFMA2 UR + 1 LDS.128 per 16", c, 32);
        run<4>("FFMA2 UR + 1 LDS.128 per 8", c, 16);
    }
    return 0;
}
