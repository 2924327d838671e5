// Microbenchmark: distributed-shared-memory (DSMEM) all-to-all inside a thread-block cluster on sm_100a.
//
// Question it answers (north-star item 1): a 256x256 complex64 wave is 512 KB = four SMs' worth of registers.  The
// on-chip form that moves the fewest bytes keeps a 64x256 row slab per CTA of a 4-CTA cluster and transposes the tile ONCE per 2-D FFT:
// every CTA sends 3/4 of its slab (96 KB) to its three peers and receives 96 KB.  How many cycles does that exchange cost, against
// the ~11 k cycles the whole register-resident 128^2 FFT takes per SM?
//
// Variants (each: cluster of CL CTAs x 512 threads, one CTA per SM, every SM of the chip busy, ITER exchanges, cycles per exchange
// from clock64 on every CTA, max over CTAs):
//   st     remote 16-byte stores through mapa'd generic pointers (st.shared::cluster), then barrier.cluster
//   ld     remote 16-byte loads (ld.shared::cluster) into local shared memory, then barrier.cluster
//   bulk   one thread per CTA issues cp.async.bulk.shared::cluster.shared::cta (TMA engine) of 32 KB to each peer, completion on the
//          receiver's mbarrier (complete_tx), plus one barrier.cluster per exchange for the write-after-read hazard
//   local  the same bytes st.shared + ld.shared inside the CTA with __syncthreads (what the 128^2 kernel's E1 exchange does)
//   sync   barrier.cluster alone
// Build:  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o dsmem_bw tools/ubench/dsmem_bw.cu
#include <cooperative_groups.h>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <vector>
namespace cg = cooperative_groups;

constexpr int THREADS = 512;
constexpr int ITER = 200;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void cluster_sync_() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// MODE 0: st, 1: ld, 2: bulk, 3: local, 4: sync only.   chunk = bytes sent to EACH peer per exchange
template <int MODE> __global__ void __launch_bounds__(THREADS, 1) k_a2a(int CL, int chunk, long long* cycles, float* sink) {
    extern __shared__ __align__(128) unsigned char smem[];
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    float4* src = reinterpret_cast<float4*>(smem);                                   // CL * chunk bytes (slot r: data for peer r)
    float4* dst = reinterpret_cast<float4*>(smem + (size_t)CL * chunk);              // CL * chunk bytes (slot r: data from peer r)
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + (size_t)2 * CL * chunk);
    const int n4 = chunk / 16;                                                       // float4 per peer chunk
    for (int i = threadIdx.x; i < CL * n4; i += THREADS) src[i] = make_float4(rank, i, 1.f, 2.f);
    if (MODE == 2 && threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    cluster_sync_();
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const long long t0 = clock64();
    for (int it = 0; it < ITER; ++it) {
        if (MODE == 0) {
            for (int p = 1; p < CL; ++p) {
                const int peer = (rank + p) % CL;
                float4* rdst = cluster.map_shared_rank(dst, peer) + (size_t)rank * n4;
                const float4* s = src + (size_t)peer * n4;
                for (int i = threadIdx.x; i < n4; i += THREADS) rdst[i] = s[i];
            }
            cluster_sync_();
        } else if (MODE == 1) {
            for (int p = 1; p < CL; ++p) {
                const int peer = (rank + p) % CL;
                const float4* rsrc = cluster.map_shared_rank(src, peer) + (size_t)rank * n4;
                float4* d = dst + (size_t)peer * n4;
                for (int i = threadIdx.x; i < n4; i += THREADS) d[i] = rsrc[i];
            }
            cluster_sync_();
        } else if (MODE == 2) {
            if (threadIdx.x == 0) {
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"((uint32_t)((CL - 1) * chunk)) : "memory");
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                for (int p = 1; p < CL; ++p) {
                    const uint32_t peer = (uint32_t)((rank + p) % CL);
                    const uint32_t rdst = mapa(smem_u32(dst) + (uint32_t)rank * chunk, peer);
                    const uint32_t rbar = mapa(smem_u32(bar), peer);
                    asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                 ::"r"(rdst), "r"(smem_u32(src) + peer * chunk), "r"((uint32_t)chunk), "r"(rbar) : "memory");
                }
            }
            // everybody waits for the incoming bytes of this exchange (phase parity = it & 1)
            uint32_t done = 0;
            while (!done) {
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                             : "=r"(done) : "r"(smem_u32(bar)), "r"((uint32_t)(it & 1)) : "memory");
            }
            cluster_sync_();                                                        // peers may overwrite our dst only after we consumed it
        } else if (MODE == 3) {
            for (int p = 1; p < CL; ++p) {
                const int peer = (rank + p) % CL;
                float4* d = dst + (size_t)rank * n4;
                const float4* s = src + (size_t)peer * n4;
                for (int i = threadIdx.x; i < n4; i += THREADS) d[i] = s[i];
            }
            __syncthreads();
        } else {
            cluster_sync_();
        }
        acc.x += dst[(threadIdx.x + it) % (CL * n4)].x;
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    if (acc.x == 123456.f) sink[0] = acc.x;
}

template <int MODE> void run(const char* name, int CL, int chunk, int sms) {
    const size_t smem = (size_t)2 * CL * chunk + 64;
    cudaFuncSetAttribute(k_a2a<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (CL > 8) cudaFuncSetAttribute(k_a2a<MODE>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    const int grid = sms / CL * CL;
    long long* d_cyc;
    float* d_sink;
    cudaMalloc(&d_cyc, grid * sizeof(long long));
    cudaMalloc(&d_sink, 4);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(THREADS);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    for (int rep = 0; rep < 2; ++rep) {
        cudaError_t e = cudaLaunchKernelEx(&cfg, k_a2a<MODE>, CL, chunk, d_cyc, d_sink);
        if (e != cudaSuccess) { printf("%-6s CL=%d chunk=%d: launch failed: %s\n", name, CL, chunk, cudaGetErrorString(e)); cudaGetLastError(); return; }
        e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%-6s CL=%d chunk=%d: %s\n", name, CL, chunk, cudaGetErrorString(e)); exit(1); }
    }
    std::vector<long long> h(grid);
    cudaMemcpy(h.data(), d_cyc, grid * sizeof(long long), cudaMemcpyDeviceToHost);
    long long mx = 0; double mean = 0;
    for (long long c : h) { mx = c > mx ? c : mx; mean += (double)c / grid; }
    const double per = (double)mx / ITER, bytes = (double)(CL - 1) * chunk;
    printf("%-6s cluster=%2d  %6.1f KB out + %6.1f KB in per CTA: %8.0f cycles / exchange (max CTA; mean %.0f)  -> %5.1f B/clk/SM each way\n",
           name, CL, bytes / 1024, bytes / 1024, per, mean / ITER, MODE == 4 ? 0.0 : bytes / per);
    cudaFree(d_cyc); cudaFree(d_sink);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    printf("%s, %d SMs, clock %d MHz\n", p.name, p.multiProcessorCount, p.clockRate / 1000);
    const int sms = p.multiProcessorCount;
    // the 256^2 transposition: cluster of 4, 32 KB to each of 3 peers (96 KB out, 96 KB in)
    for (int CL : {2, 4, 8}) {
        // src + dst regions of CL chunks each must fit 227 KB: the 4-CTA case moves 3 x 24 KB per CTA (the 256^2 transposition moves
        // 3 x 32 KB; cycles scale with the bytes, the B/clk figure is what carries over)
        const int chunk = CL == 2 ? 49152 : (CL == 4 ? 24576 : 12288);
        run<4>("sync", CL, chunk, sms);
        run<0>("st", CL, chunk, sms);
        run<1>("ld", CL, chunk, sms);
        run<2>("bulk", CL, chunk, sms);
        run<3>("local", CL, chunk, sms);
    }
    return 0;
}
