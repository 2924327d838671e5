// Microbenchmark: per-SM bandwidth of an L2-resident register <-> global exchange on sm_100a.
//
// Question it answers: the on-chip 256^2 form has to transpose the wave once per 2-D FFT between the CTAs that share a
// tile.  tools/ubench/dsmem_bw.cu measured 12-14 B/clk/SM each way for that exchange through distributed shared memory.  The
// alternative keeps the exchange buffers in global memory, small enough to stay in the 126 MB L2 (tiles in flight x 512 KB): every
// thread stores its registers with 16-byte coalesced stores and reads its new elements back with 16-byte ld.global.cg loads.  How many
// bytes per clock and SM does that path move when every SM of the chip does it at once?
//
//   wr     st.global.v4 only                  rd     ld.global.cg.v4 only           wr+rd  store own region, barrier, load partner's
// CTAS_PER_SM x THREADS threads per SM, KB per CTA and exchange as given.  Build:
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o l2_bw tools/ubench/l2_bw.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <vector>

constexpr int ITER = 200;

template <int MODE, int NV> __global__ void k_x(float4* buf, size_t region4, long long* cycles, float* sink) {
    float4 v[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = make_float4(threadIdx.x, i, blockIdx.x, 1.f);
    float4* mine = buf + (size_t)blockIdx.x * region4;
    const float4* other = buf + (size_t)(blockIdx.x ^ 1) * region4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < ITER; ++it) {
        if (MODE != 1) {
#pragma unroll
            for (int i = 0; i < NV; ++i) mine[i * blockDim.x + threadIdx.x] = v[i];
        }
        if (MODE == 2) __syncthreads();
        if (MODE != 0) {
            float4 r[NV];
#pragma unroll
            for (int i = 0; i < NV; ++i) r[i] = __ldcg(other + i * blockDim.x + threadIdx.x);
#pragma unroll
            for (int i = 0; i < NV; ++i) { acc.x += r[i].x; acc.y += r[i].w; }
        }
#pragma unroll
        for (int i = 0; i < NV; ++i) v[i].x += acc.y;
    }
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    if (acc.x == 123456.f) sink[0] = acc.x + v[0].x;
}

template <int MODE, int NV> void run(const char* name, int threads, int ctas_per_sm, int sms) {
    const int grid = sms * ctas_per_sm;
    const size_t region4 = (size_t)NV * threads;
    float4* buf; long long* d_cyc; float* d_sink;
    cudaMalloc(&buf, region4 * 16 * grid); cudaMalloc(&d_cyc, grid * 8); cudaMalloc(&d_sink, 4);
    cudaMemset(buf, 0, region4 * 16 * grid);
    for (int rep = 0; rep < 2; ++rep) k_x<MODE, NV><<<grid, threads>>>(buf, region4, d_cyc, d_sink);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
    std::vector<long long> h(grid);
    cudaMemcpy(h.data(), d_cyc, grid * 8, cudaMemcpyDeviceToHost);
    long long mx = 0; double mean = 0;
    for (long long c : h) { mx = c > mx ? c : mx; mean += (double)c / grid; }
    const double kb = region4 * 16 / 1024.0, per = (double)mx / ITER;
    const double bytes_sm = region4 * 16.0 * ctas_per_sm * (MODE == 2 ? 2 : 1);
    printf("%-6s %4d thr x %d CTA/SM, %5.0f KB per CTA%s: %8.0f cycles / exchange (max; mean %.0f) -> %5.1f B/clk/SM (footprint %.0f MB)\n", name, threads,
           ctas_per_sm, kb, MODE == 2 ? " out + same in" : "", per, mean / ITER, bytes_sm / per, region4 * 16.0 * grid / 1e6);
    cudaFree(buf); cudaFree(d_cyc); cudaFree(d_sink);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    printf("%s, %d SMs, clock %d MHz\n", p.name, p.multiProcessorCount, p.clockRate / 1000);
    const int sms = p.multiProcessorCount;
    run<0, 16>("wr", 256, 2, sms);
    run<1, 16>("rd", 256, 2, sms);
    run<2, 16>("wr+rd", 256, 2, sms);
    run<0, 16>("wr", 512, 1, sms);
    run<1, 16>("rd", 512, 1, sms);
    run<2, 16>("wr+rd", 512, 1, sms);
    run<2, 8>("wr+rd", 256, 2, sms);
    run<2, 16>("wr+rd", 256, 1, sms);
    return 0;
}
