// Microbenchmark: scalar FADD/FFMA vs packed add.f32x2 / fma.f32x2 throughput on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
#define ITER 4096
__global__ void k_scalar_add(float* out, float a0) {
    float r[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) r[i] = a0 + i;
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) r[i] = r[i] + r[(i + 5) & 15];
    }
    float s = 0; for (int i = 0; i < 16; ++i) s += r[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_packed_add(float* out, float a0) {
    unsigned long long r[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { float lo = a0 + 2 * i, hi = a0 + 2 * i + 1; asm("mov.b64 %0, {%1, %2};" : "=l"(r[i]) : "f"(lo), "f"(hi)); }
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(r[i]) : "l"(r[i]), "l"(r[(i + 3) & 7]));
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(r[i])); s += lo + hi; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_scalar_fma(float* out, float a0) {
    float r[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) r[i] = a0 + i;
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) r[i] = fmaf(r[i], r[(i + 5) & 15], r[(i + 9) & 15]);
    }
    float s = 0; for (int i = 0; i < 16; ++i) s += r[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_packed_fma(float* out, float a0) {
    unsigned long long r[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { float lo = a0 + 2 * i, hi = a0 + 2 * i + 1; asm("mov.b64 %0, {%1, %2};" : "=l"(r[i]) : "f"(lo), "f"(hi)); }
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r[i]) : "l"(r[i]), "l"(r[(i + 3) & 7]), "l"(r[(i + 5) & 7]));
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(r[i])); s += lo + hi; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <class K> float timeit(K k, float* out, const char* name, double flop_per_thread) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<<<148 * 4, 512>>>(out, 1.0f);
    cudaEventRecord(e0);
    for (int i = 0; i < 10; ++i) k<<<148 * 4, 512>>>(out, 1.0f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 10;
    double tf = flop_per_thread * 148 * 4 * 512 / (ms * 1e-3) / 1e12;
    printf("%-14s %.3f ms  %.1f Tflop-equivalent/s (adds or fmas counted as issued per fp32 lane)\n", name, ms, tf);
    return ms;
}
int main() {
    float* out; cudaMalloc(&out, 148 * 4 * 512 * 4);
    timeit(k_scalar_add, out, "scalar FADD", 16.0 * ITER);
    timeit(k_packed_add, out, "add.f32x2", 16.0 * ITER);
    timeit(k_scalar_fma, out, "scalar FFMA", 2 * 16.0 * ITER);
    timeit(k_packed_fma, out, "fma.f32x2", 2 * 16.0 * ITER);
    return 0;
}
