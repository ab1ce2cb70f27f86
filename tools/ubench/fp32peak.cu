// FP32 pipe microbenchmark: FFMA vs FFMA2 (packed f32x2, sm_100) throughput, alone and mixed with ALU work.
#include <cstdio>
#include <cuda_runtime.h>
#define ITER 4096
#define CHAINS 8
__global__ void k_ffma(float* out, float a, float b) {
    float x[CHAINS];
    for (int i = 0; i < CHAINS; i++) x[i] = threadIdx.x*0.001f+i;
    for (int it = 0; it < ITER; it++)
#pragma unroll
        for (int i = 0; i < CHAINS; i++) x[i] = fmaf(x[i], a, b);
    float s = 0; for (int i = 0; i < CHAINS; i++) s += x[i];
    if (s == 123.456f) out[0] = s;
}
__global__ void k_ffma2(float* out, float a, float b) {
    float2 x[CHAINS]; const float2 A = make_float2(a, a), B = make_float2(b, b);
    for (int i = 0; i < CHAINS; i++) x[i] = make_float2(threadIdx.x*0.001f+i, threadIdx.x*0.002f+i);
    for (int it = 0; it < ITER; it++)
#pragma unroll
        for (int i = 0; i < CHAINS; i++) x[i] = __ffma2_rn(x[i], A, B);
    float s = 0; for (int i = 0; i < CHAINS; i++) s += x[i].x+x[i].y;
    if (s == 123.456f) out[0] = s;
}
// mixed: per iteration CHAINS fma(-pairs) + CHAINS integer ALU ops (LOP3/IADD3 chain)
__global__ void k_mix1(float* out, float a, float b, int m) {
    float x[CHAINS]; unsigned u[CHAINS];
    for (int i = 0; i < CHAINS; i++) { x[i] = threadIdx.x*0.001f+i; u[i] = threadIdx.x+i; }
    for (int it = 0; it < ITER; it++)
#pragma unroll
        for (int i = 0; i < CHAINS; i++) { x[i] = fmaf(x[i], a, b); u[i] = (u[i]^m)+it; }
    float s = 0; unsigned t = 0; for (int i = 0; i < CHAINS; i++) { s += x[i]; t += u[i]; }
    if (s == 123.456f || t == 0x12345u) out[0] = s+t;
}
__global__ void k_mix2(float* out, float a, float b, int m) {
    float2 x[CHAINS/2]; unsigned u[CHAINS]; const float2 A = make_float2(a, a), B = make_float2(b, b);
    for (int i = 0; i < CHAINS/2; i++) x[i] = make_float2(threadIdx.x*0.001f+i, threadIdx.x*0.002f+i);
    for (int i = 0; i < CHAINS; i++) u[i] = threadIdx.x+i;
    for (int it = 0; it < ITER; it++) {
#pragma unroll
        for (int i = 0; i < CHAINS/2; i++) x[i] = __ffma2_rn(x[i], A, B);
#pragma unroll
        for (int i = 0; i < CHAINS; i++) u[i] = (u[i]^m)+it;
    }
    float s = 0; unsigned t = 0; for (int i = 0; i < CHAINS/2; i++) s += x[i].x+x[i].y; for (int i = 0; i < CHAINS; i++) t += u[i];
    if (s == 123.456f || t == 0x12345u) out[0] = s+t;
}
template <typename F> float timeIt(F f) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; r++) { cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); best = ms < best ? ms : best; }
    return best;
}
int main() {
    float* out; cudaMalloc(&out, 4);
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int grid = p.multiProcessorCount*8, block = 256;
    const double threads = (double) grid*block;
    float t1 = timeIt([&] { k_ffma<<<grid, block>>>(out, 1.0001f, 0.5f); });
    float t2 = timeIt([&] { k_ffma2<<<grid, block>>>(out, 1.0001f, 0.5f); });
    float t3 = timeIt([&] { k_mix1<<<grid, block>>>(out, 1.0001f, 0.5f, 5); });
    float t4 = timeIt([&] { k_mix2<<<grid, block>>>(out, 1.0001f, 0.5f, 5); });
    printf("{\"sms\": %d, \"ffma_tflops\": %.2f, \"ffma2_tflops\": %.2f, \"mix_ffma_ms\": %.4f, \"mix_ffma2_ms\": %.4f, \"ffma_ms\": %.4f, \"ffma2_ms\": %.4f, "
           "\"note\": \"ffma: 8 chains x 4096 FFMA/thread; ffma2: 8 chains x 4096 FFMA2 (2 FMA each); mix: 8 FMA (as 8 FFMA or 4 FFMA2) + 8x(LOP3+IADD3) per iteration\"}\n",
           p.multiProcessorCount, threads*ITER*CHAINS*2/(t1*1e-3)/1e12, threads*ITER*CHAINS*4/(t2*1e-3)/1e12, t3, t4, t1, t2);
    return 0;
}
