// Microbenchmark + numerical check of the cluster-resident 256x256 2-D FFT of ptyrad_b200/csrc/fft256c.cuh on sm_100a.
//
// Question it answers (north-star item 1): what does a 2-D FFT pair (forward, pointwise table multiply, inverse) of a
// 256^2 wave cost when the wave stays in the registers of an 8-CTA cluster and crosses CTAs once per FFT, against the general
// row/column path's ~21 k cycles per SM and quarter-tile FFT (which includes its pointwise physics) -- and is the crossing cheaper
// through L2 (TX_L2) or through distributed shared memory (TX_DSMEM)?
//
// Each resident cluster loads one wave (natural row-major) into layout R and runs ITER pairs; clock64 per CTA, max / mean over CTAs.
// "cycles per SM and quarter-tile FFT" = cycles per pair / resident CTAs per SM (a CTA holds an eighth of the wave, a pair is two FFTs).
// Check: forward spectrum against a float64 host DFT, and the round trip against the input.
// Build:  nvcc -O3 -std=c++17 --expt-relaxed-constexpr -gencode arch=compute_100a,code=sm_100a -o fft256_cluster tools/ubench/fft256_cluster.cu
#include "../../ptyrad_b200/csrc/fft256c.cuh"
#include <cmath>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <vector>
using namespace ptyb;
using namespace ptyb::fused256;

constexpr size_t SMEM = sizeof(float2) * (E_ELEMS + 256);

// mode 0: forward only, spectrum written in natural order; mode 1: ITER x (forward, * HF, inverse), result written in natural order
template <int TX> __global__ void __launch_bounds__(FT, 2) k_fft(const float2* __restrict__ in, float2* __restrict__ out,
                                                                 const float4* __restrict__ HF, float4* scratch, int ntiles, int iters,
                                                                 int mode, long long* cycles, int delay, int* smids) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* E = reinterpret_cast<float2*>(smem_raw);
    float2* tw = E + E_ELEMS;
    const Geo g;
    load_twiddles(tw);
    __syncthreads();
    const int cid = blockIdx.x / CL, ncl = gridDim.x / CL;
    Tx tx{scratch + (size_t)cid * TILE, 0};
    long long total = 0;
    if (threadIdx.x == 0) { uint32_t sm; asm volatile("mov.u32 %0, %%smid;" : "=r"(sm)); smids[blockIdx.x] = (int)sm; }
    if (delay) {                                           // de-phase the two CTAs of an SM: the second to arrive idles for `delay` cycles first
        __shared__ int s_slot;
        if (threadIdx.x == 0) { uint32_t sm; asm volatile("mov.u32 %0, %%smid;" : "=r"(sm)); s_slot = atomicAdd(smids + gridDim.x + sm, 1); }
        __syncthreads();
        if (s_slot & 1) {
            const long long t0 = clock64();
            while (clock64() - t0 < delay) {}
        }
    }
    for (int tile = cid; tile < ntiles; tile += ncl) {
        float2 v[32];
        const float2* src = in + (size_t)tile * TILE + g.w * 256 + 32 * g.rank + g.l;
#pragma unroll
        for (int k = 0; k < 32; ++k) v[k] = src[(size_t)k * 8 * 256];
        const long long t0 = clock64();
        if (mode == 0) {
            fft2_R_to_F<TX>(v, E, tw, g, tx);
            float2* dst = out + (size_t)tile * TILE + (size_t)(32 * g.rank + g.a) * 256 + g.xl;
#pragma unroll
            for (int u = 0; u < 32; ++u) dst[8 * u] = v[u];
        } else {
            const float4* hf = HF + (size_t)g.rank * (SLAB / 2) + g.t;
            for (int it = 0; it < iters; ++it) {
                fft2_R_to_F<TX>(v, E, tw, g, tx);
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const float4 h = __ldg(hf + j * FT);
                    v[2 * j] = cmul(v[2 * j], make_float2(h.x, h.y));
                    v[2 * j + 1] = cmul(v[2 * j + 1], make_float2(h.z, h.w));
                }
                fft2_F_to_R<TX>(v, E, tw, g, tx, [] {});
            }
            float2* dst = out + (size_t)tile * TILE + g.w * 256 + 32 * g.rank + g.l;
#pragma unroll
            for (int k = 0; k < 32; ++k) dst[(size_t)k * 8 * 256] = v[k];
        }
        total += clock64() - t0;
    }
    if (threadIdx.x == 0) cycles[blockIdx.x] = total;
}

// Ping-pong form: ONE CTA of 512 threads per SM, two independent 256-thread groups, each holding an eighth of its own wave (the
// cluster of 8 CTAs carries two waves); the groups synchronise separately (named barriers inside the CTA, one mbarrier per group
// across the cluster: TX_L2G), so one group's crossing latency is covered by the other group's DFTs.
__global__ void __launch_bounds__(2 * FT, 1) k_fft_pp(const float2* __restrict__ in, float2* __restrict__ out, const float4* __restrict__ HF,
                                                      float4* scratch, int ntiles, int iters, long long* cycles, int delay) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const Geo g;
    float2* E = reinterpret_cast<float2*>(smem_raw) + (size_t)g.grp * E_ELEMS;
    float2* tw = reinterpret_cast<float2*>(smem_raw) + 2 * E_ELEMS;
    uint64_t* mb = reinterpret_cast<uint64_t*>(tw + 256) + g.grp;
    load_twiddles(tw);
    if (g.t == 0) mbar_init(mb, CL);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    cluster_arrive();
    cluster_wait();
    const int cid = blockIdx.x / CL, ncl = gridDim.x / CL;
    Tx tx{scratch + (size_t)(2 * cid + g.grp) * TILE, 0, mb, 0};
    long long total = 0;
    if (delay && g.grp) {
        const long long t0 = clock64();
        while (clock64() - t0 < delay) {}
    }
    for (int tile = 2 * cid + g.grp; tile < ntiles; tile += 2 * ncl) {
        float2 v[32];
        const float2* src = in + (size_t)tile * TILE + g.w * 256 + 32 * g.rank + g.l;
#pragma unroll
        for (int k = 0; k < 32; ++k) v[k] = src[(size_t)k * 8 * 256];
        const long long t0 = clock64();
        const float4* hf = HF + (size_t)g.rank * (SLAB / 2) + g.t;
        for (int it = 0; it < iters; ++it) {
            fft2_R_to_F<TX_L2G, 2>(v, E, tw, g, tx);
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const float4 h = __ldg(hf + j * FT);
                v[2 * j] = cmul(v[2 * j], make_float2(h.x, h.y));
                v[2 * j + 1] = cmul(v[2 * j + 1], make_float2(h.z, h.w));
            }
            fft2_F_to_R<TX_L2G, 2>(v, E, tw, g, tx, [] {});
        }
        float2* dst = out + (size_t)tile * TILE + g.w * 256 + 32 * g.rank + g.l;
#pragma unroll
        for (int k = 0; k < 32; ++k) dst[(size_t)k * 8 * 256] = v[k];
        total += clock64() - t0;
    }
    if (g.t == 0) cycles[2 * blockIdx.x + g.grp] = total;
    // no CTA may exit while a peer can still arrive on its mbarriers
    cluster_arrive();
    cluster_wait();
}

int run_pp(int delay) {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const size_t smem = sizeof(float2) * (2 * E_ELEMS + 256) + 16;
    cudaFuncSetAttribute(k_fft_pp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(2 * FT);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cfg.gridDim = dim3(CL * 64);
    int ncl = 0;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&ncl, k_fft_pp, &cfg);
    if (e != cudaSuccess || ncl < 1) { printf("pingpong: cudaOccupancyMaxActiveClusters: %s (%d)\n", cudaGetErrorString(e), ncl); return 1; }
    cfg.gridDim = dim3(CL * ncl);
    const int ntiles = 2 * ncl;
    std::vector<std::complex<float>> h_in((size_t)ntiles * TILE), h_out((size_t)ntiles * TILE);
    srand(1);
    for (auto& z : h_in) z = {float(rand()) / RAND_MAX - 0.5f, float(rand()) / RAND_MAX - 0.5f};
    float2 *d_in, *d_out; float4 *d_hf, *d_scr; long long* d_cyc;
    cudaMalloc(&d_in, h_in.size() * 8); cudaMalloc(&d_out, h_in.size() * 8);
    cudaMalloc(&d_hf, TILE * 8); cudaMalloc(&d_scr, (size_t)2 * ncl * TILE * 16); cudaMalloc(&d_cyc, 2 * CL * ncl * 8);
    cudaMemcpy(d_in, h_in.data(), h_in.size() * 8, cudaMemcpyHostToDevice);
    std::vector<float2> hf(TILE, make_float2(1.0f / TILE, 0.f));
    cudaMemcpy(d_hf, hf.data(), TILE * 8, cudaMemcpyHostToDevice);
    auto launch = [&](int iters) {
        cudaError_t e2 = cudaLaunchKernelEx(&cfg, k_fft_pp, (const float2*)d_in, d_out, (const float4*)d_hf, d_scr, ntiles, iters, d_cyc, delay);
        if (e2 == cudaSuccess) e2 = cudaDeviceSynchronize();
        if (e2 != cudaSuccess) { printf("pingpong: %s\n", cudaGetErrorString(e2)); exit(1); }
    };
    launch(3);
    cudaMemcpy(h_out.data(), d_out, h_out.size() * 8, cudaMemcpyDeviceToHost);
    double err = 0, nrm = 0;
    for (size_t i = 0; i < h_in.size(); ++i) { err += std::norm(std::complex<double>(h_out[i]) - std::complex<double>(h_in[i])); nrm += std::norm(std::complex<double>(h_in[i])); }
    printf("pingpong 3 round trips, all %d tiles: rel l2 error %.2e\n", ntiles, std::sqrt(err / nrm));
    const int ITER = 200;
    launch(ITER);
    std::vector<long long> h(2 * CL * ncl);
    cudaMemcpy(h.data(), d_cyc, h.size() * 8, cudaMemcpyDeviceToHost);
    long long mx = 0; double mean = 0;
    for (long long c : h) { mx = c > mx ? c : mx; mean += (double)c / h.size(); }
    const double occ = 2.0 * CL * ncl / p.multiProcessorCount;      // wave eighths per SM, chip average
    printf("pingpong delay %d: %d resident clusters x 2 waves (%.2f eighths/SM): %8.0f cycles / FFT pair per group (max; mean %.0f) -> %6.0f cycles per SM and quarter-tile FFT\n",
           delay, ncl, occ, (double)mx / ITER, mean / ITER, (double)mx / ITER / occ);
    cudaFree(d_in); cudaFree(d_out); cudaFree(d_hf); cudaFree(d_scr); cudaFree(d_cyc);
    return 0;
}

template <int TX> int run(const char* name, int occ_target, int delay = 0) {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    cudaFuncSetAttribute(k_fft<TX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM);
    // occ_target 1: pad the dynamic shared memory so that only one CTA fits per SM
    const size_t smem = occ_target == 1 ? 120 * 1024 : SMEM;
    cudaFuncSetAttribute(k_fft<TX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(FT);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cfg.gridDim = dim3(CL * 64);
    int ncl = 0;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&ncl, k_fft<TX>, &cfg);
    if (e != cudaSuccess || ncl < 1) { printf("%s: cudaOccupancyMaxActiveClusters: %s (%d)\n", name, cudaGetErrorString(e), ncl); return 1; }
    cfg.gridDim = dim3(CL * ncl);
    const int ntiles = ncl;
    std::vector<std::complex<float>> h_in((size_t)ntiles * TILE), h_out((size_t)ntiles * TILE);
    srand(1);
    for (auto& z : h_in) z = {float(rand()) / RAND_MAX - 0.5f, float(rand()) / RAND_MAX - 0.5f};
    float2 *d_in, *d_out; float4 *d_hf, *d_scr; long long* d_cyc;
    cudaMalloc(&d_in, h_in.size() * 8); cudaMalloc(&d_out, h_in.size() * 8);
    cudaMalloc(&d_hf, TILE * 8); cudaMalloc(&d_scr, (size_t)ncl * TILE * 16); cudaMalloc(&d_cyc, CL * ncl * 8); int* d_sm; cudaMalloc(&d_sm, (CL * ncl + 256) * 4);
    cudaMemcpy(d_in, h_in.data(), h_in.size() * 8, cudaMemcpyHostToDevice);
    std::vector<float2> hf(TILE, make_float2(1.0f / TILE, 0.f));
    cudaMemcpy(d_hf, hf.data(), TILE * 8, cudaMemcpyHostToDevice);
    auto launch = [&](int iters, int mode) {
        cudaMemset(d_sm, 0, (CL * ncl + 256) * 4);
        cudaError_t e2 = cudaLaunchKernelEx(&cfg, k_fft<TX>, (const float2*)d_in, d_out, (const float4*)d_hf, d_scr, ntiles, iters, mode, d_cyc, mode ? delay : 0, d_sm);
        if (e2 == cudaSuccess) e2 = cudaDeviceSynchronize();
        if (e2 != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e2)); exit(1); }
    };
    // numerical check, tile 0: forward spectrum vs float64 host DFT
    launch(1, 0);
    cudaMemcpy(h_out.data(), d_out, h_out.size() * 8, cudaMemcpyDeviceToHost);
    {
        const int N = 256;
        std::vector<std::complex<double>> W(N), tmp((size_t)N * N), ref((size_t)N * N);
        for (int n = 0; n < N; ++n) W[n] = std::polar(1.0, -2.0 * M_PI * n / N);
        for (int y = 0; y < N; ++y)
            for (int kx = 0; kx < N; ++kx) {
                std::complex<double> s = 0;
                for (int x = 0; x < N; ++x) s += std::complex<double>(h_in[(size_t)y * N + x]) * W[(x * kx) % N];
                tmp[(size_t)y * N + kx] = s;
            }
        double err = 0, nrm = 0;
        for (int ky = 0; ky < N; ++ky)
            for (int kx = 0; kx < N; ++kx) {
                std::complex<double> s = 0;
                for (int y = 0; y < N; ++y) s += tmp[(size_t)y * N + kx] * W[(y * ky) % N];
                const std::complex<double> d = std::complex<double>(h_out[(size_t)ky * N + kx]) - s;
                err += std::norm(d); nrm += std::norm(s);
            }
        printf("%-8s forward vs float64 DFT: rel l2 error %.2e\n", name, std::sqrt(err / nrm));
    }
    const int ITER = 200;
    launch(3, 1);
    cudaMemcpy(h_out.data(), d_out, h_out.size() * 8, cudaMemcpyDeviceToHost);
    {
        double err = 0, nrm = 0;
        for (size_t i = 0; i < h_in.size(); ++i) { err += std::norm(std::complex<double>(h_out[i]) - std::complex<double>(h_in[i])); nrm += std::norm(std::complex<double>(h_in[i])); }
        printf("%-8s 3 round trips, all %d tiles: rel l2 error %.2e\n", name, ntiles, std::sqrt(err / nrm));
    }
    launch(ITER, 1);
    std::vector<long long> h(CL * ncl);
    cudaMemcpy(h.data(), d_cyc, h.size() * 8, cudaMemcpyDeviceToHost);
    long long mx = 0; double mean = 0;
    for (long long c : h) { mx = c > mx ? c : mx; mean += (double)c / h.size(); }
    const double occ = (double)CL * ncl / p.multiProcessorCount;
    {
        std::vector<int> sm(CL * ncl);
        cudaMemcpy(sm.data(), d_sm, sm.size() * 4, cudaMemcpyDeviceToHost);
        int same = 0, pairs = 0;
        for (size_t i = 0; i < sm.size(); ++i)
            for (size_t j = i + 1; j < sm.size(); ++j)
                if (sm[i] == sm[j]) { ++pairs; same += (i / CL == j / CL); }
        printf("%-8s co-resident CTA pairs: %d, of which in the same cluster: %d; cluster 0 on SMs", name, pairs, same);
        for (int i = 0; i < CL; ++i) printf(" %d", sm[i]);
        printf("\n");
        if (delay == 1) for (int c = 0; c < ncl; ++c) { printf("  cluster %2d:", c); for (int i = 0; i < CL; ++i) printf(" %3d", sm[c * CL + i]); printf("\n"); }
    }
    printf("%-8s delay %d: %d resident clusters (%.2f CTAs/SM): %8.0f cycles / FFT pair per CTA (max; mean %.0f) -> %6.0f cycles per SM and quarter-tile FFT\n",
           name, delay, ncl, occ, (double)mx / ITER, mean / ITER, (double)mx / ITER / occ);
    cudaFree(d_in); cudaFree(d_out); cudaFree(d_hf); cudaFree(d_scr); cudaFree(d_cyc);
    return 0;
}

int main(int argc, char** argv) {
    if (argc > 1 && argv[1][0] == 'p') return run_pp(argc > 2 ? atoi(argv[2]) : 0);
    if (argc > 1) return argv[1][0] == 'd' ? run<TX_DSMEM>("DSMEM", 2) : run<TX_L2>("L2", 2);      // one variant (for ncu)
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    printf("%s, %d SMs, clock %d MHz\n", p.name, p.multiProcessorCount, p.clockRate / 1000);
    run<TX_L2>("L2", 2);
    run_pp(0);
    run_pp(6000);
    run_pp(12000);
    run<TX_DSMEM>("DSMEM", 2);
    run<TX_NONE>("none", 2);
    run<TX_BARRIER>("barrier", 2);
    run<TX_L2>("L2/occ1", 1);
    run<TX_DSMEM>("DSM/occ1", 1);
    run<TX_NONE>("none/occ1", 1);
    run<TX_BARRIER>("barr/occ1", 1);
    return 0;
}
