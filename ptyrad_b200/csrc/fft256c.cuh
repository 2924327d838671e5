// 256x256 complex 2-D FFT shared by the 8 CTAs of a thread-block cluster, wave in REGISTERS (sm_100a).
//
// A 256^2 complex64 wave is 512 KB: more than one SM holds.  Eight CTAs of 256 threads keep 32 complex values per thread (64 KB per
// CTA, two CTAs per SM, so a whole wave occupies four SMs' worth of registers) and carry it through every slice like fused128.cuh does
// for 128^2.  tools/proto_fused256.py is the NumPy model of this index algebra.
//
//   layout R (real space)  CTA c, thread t: x = 32c + (t & 31), yl = t >> 5;      v[k] = psi[yl + 8k][x]        (rows of 256 B per warp)
//   layout M (middle)      CTA c, thread t: a = t >> 3, xl = t & 7;               v[4b + i] = Y[x = 32c + xl + 8i][ky = a + 32b]
//   layout F1              CTA c', thread t: a = t >> 3, xl = t & 7;              v[k] = Y[x = xl + 8k][ky = a + 32c']
//   layout F (Fourier)     CTA c', thread t: ky = 32c' + (t >> 3), lx = t & 7;    v[u] = X[ky][kx = lx + 8u]
//
//   forward 2-D FFT (R -> F): DFT32 over k in registers (y, stride 8) -> exchange Ea through shared memory (CTA wide)
//                             twiddle W256^(yl a), DFT8 over yl (rest of y)                                   = layout M
//                             transposition T between the CTAs of the cluster (register pairs as 16-byte words) = layout F1
//                             DFT32 over k in registers (x, stride 8)  -> exchange Eb (warp local, same shared memory)
//                             twiddle W256^(xl a2), DFT8 over xl (rest of x)                                  = layout F
//   inverse (F -> R): the same stages backwards with conjugate twiddles; T is its own mirror (same store / load index functions).
//
// T, the one step that crosses CTAs, has two implementations (tools/ubench/fft256_cluster.cu measures both):
//   TX_L2    every thread stores its 16 register pairs to a cluster-private scratch tile in global memory (coalesced 512 B per warp,
//            small enough to stay in L2: 2 x 512 KB per resident cluster), barrier.cluster (release / acquire), 16 ld.global.cg.v4
//            straight into the registers of the new layout.  No shared memory involved.
//   TX_DSMEM the pairs go to the destination CTA's shared memory with st.shared::cluster.v4 (the receive buffer aliases the exchange
//            buffer, so a second, split-phase cluster barrier orders it against the peers' earlier reads), then LDS.128.
#pragma once
#include "dft_regs.cuh"
#include <stdint.h>

namespace ptyb {
namespace fused256 {

constexpr int FN = 256;
constexpr int CL = 8;                    // CTAs per cluster = per wave
constexpr int FT = 256;                  // threads per CTA
constexpr int TILE = FN * FN;            // 65536
constexpr int SLAB = TILE / CL;          // 8192 elements per CTA, 32 per thread
constexpr int SA = 264;                  // Ea: float2 stride between the 32 a-planes (8 yl x 32 x + 8 pad: conflict-free DFT8-stage reads)
constexpr int SB = 33;                   // Eb: per-warp [a2][lane] rows padded to 33
constexpr int E_ELEMS = 32 * SA;         // 8448 float2 = 67584 B (Eb: 8 warps x 32 x 33 = the same 8448)

enum { TX_L2 = 0, TX_DSMEM = 1, TX_NONE = 2, TX_BARRIER = 3, TX_L2G = 4 };   // 2, 3 (no exchange / barrier only) exist for the microbenchmark

struct Geo {
    int t, w, l, a, xl, al, rank, grp;
    __device__ __forceinline__ Geo() {
        grp = threadIdx.x >> 8;                      // two independent 256-thread groups per CTA in the ping-pong kernels, else 0
        t = threadIdx.x & 255; w = t >> 5; l = t & 31; a = t >> 3; xl = t & 7; al = l >> 3;
        uint32_t r;
        asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
        rank = (int)r;
    }
};

__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t mapa_rank(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_cluster_f4(uint32_t addr, float2 a, float2 b) {
    asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y) : "memory");
}

// Transposition state of one wave: the two alternating L2 scratch tiles (TX_L2 / TX_L2G) and which one is next; TX_L2G: the
// group's mbarrier in this CTA's shared memory (8 arrivals per phase, one from the same group of every CTA of the cluster)
struct Tx {
    float4* scr;        // this wave's 2 x (TILE/2) float4
    int phase;
    uint64_t* mbar;
    uint32_t parity;
};

// barrier over the 256 threads that share a wave slab: the whole CTA, or one of the two groups of a ping-pong CTA (named barrier)
template <int GROUPS> __device__ __forceinline__ void group_sync(const Geo& g) {
    if (GROUPS == 1) __syncthreads();
    else asm volatile("bar.sync %0, 256;" ::"r"(1 + g.grp) : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* mb, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(mb)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* mb, uint32_t rank) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(mapa_rank(smem_addr(mb), rank)) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* mb, uint32_t parity) {
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n\t@!p bra WAIT_%=;\n\t}"
                 ::"r"(smem_addr(mb)), "r"(parity) : "memory");
}

// pair q of this thread goes to CTA q >> 1, slot [2 rank + (q & 1)][t]; this thread's new pair q is slot [q][t] of its own CTA.
// TX_DSMEM: `armed` says that the split-phase "receive buffer is free" arrive has been issued (after the caller's last read of E).
template <int TX, int GROUPS>
__device__ __forceinline__ void transpose_T(float2 (&v)[32], Tx& tx, float2* E, const Geo& g) {
    if (TX == TX_NONE) {
        group_sync<GROUPS>(g);
    } else if (TX == TX_L2G) {
        // like TX_L2, but the crossing is synchronised per GROUP: the group's stores are ordered by its named barrier, eight lanes
        // release-arrive on the same group's mbarrier of every CTA of the cluster, everybody acquires on its own.  The other group of
        // the CTA keeps computing meanwhile (barrier.cluster would stop both).
        float4* buf = tx.scr + (size_t)(tx.phase & 1) * (TILE / 2);
        tx.phase ^= 1;
        float4* dst = buf + (2 * g.rank) * FT + g.t;
#pragma unroll
        for (int q = 0; q < 16; ++q)
            dst[((q >> 1) * 16 + (q & 1)) * FT] = make_float4(v[2 * q].x, v[2 * q].y, v[2 * q + 1].x, v[2 * q + 1].y);
        group_sync<GROUPS>(g);
        if (g.t < CL) mbar_arrive_remote(tx.mbar, (uint32_t)g.t);
        mbar_wait_cluster(tx.mbar, tx.parity);
        tx.parity ^= 1;
        const float4* src = buf + (size_t)(g.rank * 16) * FT + g.t;
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            const float4 r = __ldcg(src + q * FT);
            v[2 * q] = make_float2(r.x, r.y);
            v[2 * q + 1] = make_float2(r.z, r.w);
        }
    } else if (TX == TX_BARRIER) {
        cluster_arrive();
        cluster_wait();
    } else if (TX == TX_L2) {
        float4* buf = tx.scr + (size_t)(tx.phase & 1) * (TILE / 2);
        tx.phase ^= 1;
        float4* dst = buf + (2 * g.rank) * FT + g.t;
#pragma unroll
        for (int q = 0; q < 16; ++q)
            dst[((q >> 1) * 16 + (q & 1)) * FT] = make_float4(v[2 * q].x, v[2 * q].y, v[2 * q + 1].x, v[2 * q + 1].y);
        cluster_arrive();
        cluster_wait();
        const float4* src = buf + (size_t)(g.rank * 16) * FT + g.t;
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            const float4 r = __ldcg(src + q * FT);
            v[2 * q] = make_float2(r.x, r.y);
            v[2 * q + 1] = make_float2(r.z, r.w);
        }
    } else {
        cluster_wait();                                     // every peer has finished reading its exchange buffer (arrive: see callers)
        const uint32_t slot = smem_addr(E) + (uint32_t)((2 * g.rank) * FT + g.t) * 16u;
#pragma unroll
        for (int q = 0; q < 16; ++q)
            st_cluster_f4(mapa_rank(slot + (uint32_t)((q & 1) * FT * 16), (uint32_t)(q >> 1)), v[2 * q], v[2 * q + 1]);
        cluster_arrive();
        cluster_wait();
        const float4* src = reinterpret_cast<const float4*>(E) + g.t;
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            const float4 r = src[q * FT];
            v[2 * q] = make_float2(r.x, r.y);
            v[2 * q + 1] = make_float2(r.z, r.w);
        }
        group_sync<GROUPS>(g);                              // the next exchange writes other warps' slots of the same buffer
    }
}

// tw[n] = exp(-2 pi i n / 256), n < 256 (shared memory)
template <int TX, int GROUPS = 1>
__device__ __forceinline__ void fft2_R_to_F(float2 (&v)[32], float2* E, const float2* tw, const Geo& g, Tx& tx) {
    Dft<32, -1>::run(v);
    group_sync<GROUPS>(g);                                  // earlier readers of E are done
    {
        float2* p = E + g.w * 32 + g.l;
#pragma unroll
        for (int a = 0; a < 32; ++a) p[a * SA] = v[a];
    }
    group_sync<GROUPS>(g);
    {
        const float2* q = E + g.a * SA + g.xl;
        float2 c[4][8];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int y = 0; y < 8; ++y) c[i][y] = q[y * 32 + 8 * i];
        if (TX == TX_DSMEM) cluster_arrive();               // this CTA's exchange buffer may now receive the peers' pairs
#pragma unroll
        for (int y = 1; y < 8; ++y) {
            const float2 wv = tw[y * g.a];
#pragma unroll
            for (int i = 0; i < 4; ++i) c[i][y] = cmul(c[i][y], wv);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            Dft<8, -1>::run(c[i]);
#pragma unroll
            for (int b = 0; b < 8; ++b) v[4 * b + i] = c[i][b];
        }
    }
    transpose_T<TX, GROUPS>(v, tx, E, g);
    Dft<32, -1>::run(v);
    float2* eb = E + g.w * (32 * SB);
    {
        float2* p = eb + g.l;
#pragma unroll
        for (int a2 = 0; a2 < 32; ++a2) p[a2 * SB] = v[a2];
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int a2 = g.xl + 8 * i;
        const float2* r = eb + a2 * SB + g.al * 8;
        float2 c[8];
#pragma unroll
        for (int x = 0; x < 8; ++x) c[x] = r[x];
#pragma unroll
        for (int x = 1; x < 8; ++x) c[x] = cmul(c[x], tw[x * a2]);
        Dft<8, -1>::run(c);
#pragma unroll
        for (int b = 0; b < 8; ++b) v[i + 4 * b] = c[b];
    }
}

// `pre` runs between the last shared-memory read and the last register DFT: from there on the 32 slots this thread has just read
// (E[a*SA + w*32 + l]) belong to it alone until the next forward FFT's first barrier (the kernels park the asynchronous copy of the
// next slice's object ROI there, like fused128.cuh).
template <int TX, int GROUPS = 1, class Pre>
__device__ __forceinline__ void fft2_F_to_R(float2 (&v)[32], float2* E, const float2* tw, const Geo& g, Tx& tx, Pre pre) {
    float2* eb = E + g.w * (32 * SB);
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int a2 = g.xl + 8 * i;
        float2* r = eb + a2 * SB + g.al * 8;
        float2 c[8];
#pragma unroll
        for (int b = 0; b < 8; ++b) c[b] = v[i + 4 * b];
        Dft<8, +1>::run(c);
        r[0] = c[0];
#pragma unroll
        for (int x = 1; x < 8; ++x) r[x] = cmulc(c[x], tw[x * a2]);
    }
    __syncwarp();
    {
        const float2* p = eb + g.l;
#pragma unroll
        for (int a2 = 0; a2 < 32; ++a2) v[a2] = p[a2 * SB];
    }
    if (TX == TX_DSMEM) cluster_arrive();                   // (aligned: every thread of the CTA arrives after its own last read)
    Dft<32, +1>::run(v);
    transpose_T<TX, GROUPS>(v, tx, E, g);
    {
        float2* q = E + g.a * SA + g.xl;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float2 c[8];
#pragma unroll
            for (int b = 0; b < 8; ++b) c[b] = v[4 * b + i];
            Dft<8, +1>::run(c);
            q[8 * i] = c[0];
#pragma unroll
            for (int y = 1; y < 8; ++y) q[y * 32 + 8 * i] = cmulc(c[y], tw[y * g.a]);
        }
    }
    group_sync<GROUPS>(g);
    {
        const float2* p = E + g.w * 32 + g.l;
#pragma unroll
        for (int a = 0; a < 32; ++a) v[a] = p[a * SA];
    }
    pre();
    Dft<32, +1>::run(v);
}

__device__ __forceinline__ void load_twiddles(float2* tw) {
    for (int n = threadIdx.x; n < 256; n += blockDim.x) {
        float sn, cs;
        sincospif(-2.0f * float(n) / 256.0f, &sn, &cs);
        tw[n] = make_float2(cs, sn);
    }
}

}  // namespace fused256
}  // namespace ptyb
