// Fused on-chip multislice kernels for N = 128 (sm_100a).
//
// One CTA of 512 threads keeps one 128x128 complex wave entirely in REGISTERS (32 complex values per thread,
// 128 KB = half of the SM's register file) and carries it through every slice: transmission multiply, 2-D FFT,
// propagator multiply, inverse 2-D FFT -- the wave never returns to HBM except for the one stash store per slice
// that the adjoint needs.  Shared memory (132 KB) is only the exchange medium of the FFT:
//
//   layout R (real space)   thread t: x = t & 127, yl = t >> 7;          v[k] = psi[yl + 4k][x]
//   layout F (Fourier)      thread t: w2 = t >> 5, l = t & 31, rsel = l >> 4, q = (l >> 2) & 3, vv = l & 3;
//                                                                         v[u] = X[w2 + 16 rsel + 32 q][vv + 4u]
//   forward 2-D FFT (R -> F): DFT32 over k in registers (y, stride 4)        -> exchange E1 through smem (CTA wide)
//                             y-twiddle, 4x4 DFT (rest of y and of x), x-twiddle -> exchange E2 (warp local, same smem)
//                             DFT32 over j in registers (x)
//   inverse (F -> R): the same stages backwards with conjugate twiddles.
//
// So a 2-D FFT costs two shared-memory round trips (4 x 128 KB of smem traffic) instead of the eight of a row/column
// slab FFT, the DFT32 twiddles are compile-time immediates, every global access of layout R is a coalesced row
// segment (ROI gather, gradient scatter; 512 B per warp with the packed layouts below), and layout-F side tables (propagator, probe spectrum, dL/dI)
// are pre-permuted so that they are read thread-privately and coalesced.  tools/proto_fused128.py is the NumPy
// model of this index algebra.
//
// Forward: grid (P, M, B), one wave per CTA; psi_z is stashed per slice through per-warp shared-memory staging blocks and
// TMA bulk stores (cp.async.bulk), the far-field spectrum is kept for the adjoint.  Adjoint: one CTA per (sample, object
// mode, probe mode); conj(psi_z) gphi_z is scattered straight into the packed, L2-resident dense object gradient with
// red.global.add.v4.f32.  Every buffer this path owns uses 16-byte "pair" layouts so that a thread's two consecutive elements
// are one 128-bit access.
#pragma once
#include "general_kernels.cuh"
#include "../../include/ptyrad_b200.h"
#include <atomic>
#include <functional>
#include <string>
#include <cstring>

namespace ptyb {
namespace fused128 {

constexpr int FN = 128;
constexpr int FT = 512;                 // threads per CTA
constexpr int CH = 528;                 // elements per chunk region (512 used by E1, 16 rows x 33 by E2)
constexpr int E_ELEMS = 32 * CH;        // 16896 float2 = 135168 B
constexpr int TILE = FN * FN;           // 16384
// exchange buffer + twiddle/ramp tables + reduction scratch (+ 16 per-warp 4 KB TMA staging blocks where used)
constexpr size_t SMEM_BYTES_BWD = sizeof(float2) * (E_ELEMS + 128 + 4 * 128) + 128 * sizeof(float);
constexpr size_t SMEM_BYTES_FWD = sizeof(float2) * (E_ELEMS + 128 + 4 * 128) + 128 * sizeof(float) + 16 * 4096;   // + stash staging

struct Args {
    FwdArgs f;
    const float2* HF;       // (TILE) layout F, pre-scaled by 1/N^2
    const float2* PhatF;    // (P, TILE) layout F, pre-scaled by 1/N^2
    float2* farF;           // (B, M, P, TILE) layout F far-field spectra F2(psi_{Z-1} O_{Z-1}) (unnormalised), kept for the adjoint
    float4* gOpack;         // (M,Z,Noy,Nox) packed dense object-gradient scratch
    float2* phisF;          // (B,P,M,Z-1,TILE) layout F or null
    // adjoint only
    const float* G;
    float2* gPhatF;         // (P, TILE) layout F
    float2* gprobe;         // (P, N, N) natural (unshifted probes)
    float* gprop;
    float* gshift;
    float dx, k0;
    int shift, need_obj, need_probe, need_shift, need_prop, units;
};

struct Geo {
    int t, x, yl, w2, lane, rsel, e16, ky, vv;
    __device__ __forceinline__ Geo() {
        t = threadIdx.x; x = t & 127; yl = t >> 7; w2 = t >> 5; lane = t & 31;
        rsel = lane >> 4; e16 = lane & 15; vv = lane & 3;
        ky = w2 + 16 * rsel + 32 * ((lane >> 2) & 3);
    }
    __device__ __forceinline__ int kx(int u) const { return vv + 4 * u; }
};

// ---- the 2-D FFT over registers + shared memory -------------------------------------------------------------
// tw[n] = exp(-2 pi i n / 128), n < 128 (shared memory)
template <int DIR> __device__ __forceinline__ float2 twd(const float2* tw, int e) {
    float2 w = tw[e];
    if (DIR > 0) w.y = -w.y;
    return w;
}

__device__ __forceinline__ void fft2_R_to_F(float2 (&v)[32], float2* E, const float2* tw, const Geo& g) {
    Dft<32, -1>::run(v);
    // No CTA barrier before these stores: thread t writes exactly the slots E[r*CH + yl*N + x] that it alone has read (layout-R reads
    // at the end of the previous inverse FFT, behind that FFT's CTA barrier) and that it alone filled with the ROI (cp.async, waited
    // for by t); every other use of E by the kernels is fenced by its own barriers.  Measured: forward 1.170 -> 1.138 ms, adjoint
    // 1.267 -> 1.250 ms at C2, S64 +1.3 % (profiles/r02/ab_first_barrier_and_early_stash.txt).
    {
        float2* p = E + g.yl * 128 + g.x;
#pragma unroll
        for (int r = 0; r < 32; ++r) p[r * CH] = v[r];
    }
    __syncthreads();
    const int j = g.lane;
    float2 xt[4];
#pragma unroll
    for (int c = 1; c < 4; ++c) xt[c] = tw[j * c];
#pragma unroll
    for (int rs = 0; rs < 2; ++rs) {
        const int r = g.w2 + 16 * rs;
        float2* ch = E + r * CH;
        float2 a[4][4];                                // [yl][s]
#pragma unroll
        for (int y = 0; y < 4; ++y)
#pragma unroll
            for (int s = 0; s < 4; ++s) a[y][s] = ch[y * 128 + j + 32 * s];
        __syncwarp();
#pragma unroll
        for (int y = 1; y < 4; ++y) {
            const float2 w = tw[y * r];
#pragma unroll
            for (int s = 0; s < 4; ++s) a[y][s] = cmul(a[y][s], w);
        }
#pragma unroll
        for (int s = 0; s < 4; ++s) {                  // DFT4 over yl -> q
            float2 c[4] = {a[0][s], a[1][s], a[2][s], a[3][s]};
            Dft<4, -1>::run(c);
            a[0][s] = c[0]; a[1][s] = c[1]; a[2][s] = c[2]; a[3][s] = c[3];
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {                  // DFT4 over s -> vv, x-twiddle, store transposed
            Dft<4, -1>::run(a[q]);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                float2 o = c ? cmul(a[q][c], xt[c]) : a[q][c];
                ch[(q * 4 + c) * 33 + j] = o;
            }
        }
    }
    __syncwarp();
    {
        const float2* p = E + (g.w2 + 16 * g.rsel) * CH + g.e16 * 33;
#pragma unroll
        for (int jj = 0; jj < 32; ++jj) v[jj] = p[jj];
    }
    Dft<32, -1>::run(v);
}

// `pre` runs between the last shared-memory read and the last register DFT: from there on the 32 slots this thread has just read
// (E[r*CH + yl*128 + x]) belong to it alone until the next forward FFT's first barrier, which is where the kernels park the
// asynchronous copy of the next slice's object ROI (prefetch_roi_to_E).
template <class Pre>
__device__ __forceinline__ void fft2_F_to_R(float2 (&v)[32], float2* E, const float2* tw, const Geo& g, Pre pre) {
    Dft<32, +1>::run(v);
    // no CTA barrier here: this warp only writes its OWN two chunks, whose only foreign readers are the layout-R reads at
    // the end of an earlier inverse FFT, and a forward FFT (whose CTA barrier every thread passes after those reads) always runs between two inverse FFTs
    __syncwarp();
    {
        float2* p = E + (g.w2 + 16 * g.rsel) * CH + g.e16 * 33;
#pragma unroll
        for (int jj = 0; jj < 32; ++jj) p[jj] = v[jj];
    }
    __syncwarp();
    const int j = g.lane;
    float2 xt[4];
#pragma unroll
    for (int c = 1; c < 4; ++c) xt[c] = cconj(tw[j * c]);
#pragma unroll
    for (int rs = 0; rs < 2; ++rs) {
        const int r = g.w2 + 16 * rs;
        float2* ch = E + r * CH;
        float2 a[4][4];                                // [q][vv]
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                float2 o = ch[(q * 4 + c) * 33 + j];
                a[q][c] = c ? cmul(o, xt[c]) : o;
            }
        __syncwarp();
#pragma unroll
        for (int q = 0; q < 4; ++q) Dft<4, +1>::run(a[q]);      // over vv -> s
        float2 yt[4];                                           // loaded once: the stores below would stop the compiler from reusing them
#pragma unroll
        for (int y = 1; y < 4; ++y) yt[y] = tw[y * r];
#pragma unroll
        for (int s = 0; s < 4; ++s) {                           // over q -> yl
            float2 c[4] = {a[0][s], a[1][s], a[2][s], a[3][s]};
            Dft<4, +1>::run(c);
#pragma unroll
            for (int y = 0; y < 4; ++y) {
                float2 o = y ? cmulc(c[y], yt[y]) : c[y];       // conj(W^{yl r})
                ch[y * 128 + j + 32 * s] = o;
            }
        }
    }
    __syncthreads();
    {
        const float2* p = E + g.yl * 128 + g.x;
#pragma unroll
        for (int r = 0; r < 32; ++r) v[r] = p[r * CH];
    }
    pre();
    Dft<32, +1>::run(v);
}

// ---- shared memory carve ----------------------------------------------------------------------------------------
struct Smem {
    float2* E;      // exchange buffer
    float2* tw;     // 128
    float2* wy;     // 128 shift ramp (y), wx, tilt ramps ey, ex
    float2* wx;
    float2* ey;
    float2* ex;
    float* fl;      // forward only, 64 KB: per-warp TMA staging blocks of the stash stores (16 warps x 4 KB)
    float* red;     // 128 floats (block_sum scratch)
};
__device__ __forceinline__ Smem carve_smem(unsigned char* raw) {
    Smem s;
    s.E = reinterpret_cast<float2*>(raw);
    s.tw = s.E + E_ELEMS;
    s.wy = s.tw + 128; s.wx = s.wy + 128; s.ey = s.wx + 128; s.ex = s.ey + 128;
    s.red = reinterpret_cast<float*>(s.ex + 128);     // 128 floats of block_sum scratch
    s.fl = s.red + 128;                               // forward only (SMEM_BYTES_FWD); the adjoint launches without it
    return s;
}

__device__ __forceinline__ void load_tables(const Smem& s, const Args& a, int b) {
    for (int n = threadIdx.x; n < 128; n += blockDim.x) {
        float sn, cs;
        sincospif(-2.0f * float(n) / 128.0f, &sn, &cs);
        s.tw[n] = make_float2(cs, sn);
        if (a.f.wvec) {
            s.wy[n] = a.f.wvec[((size_t)b * 2 + 0) * 128 + n];
            s.wx[n] = a.f.wvec[((size_t)b * 2 + 1) * 128 + n];
        }
        if (a.f.tvec) {
            s.ey[n] = a.f.tvec[((size_t)b * 2 + 0) * 128 + n];
            s.ex[n] = a.f.tvec[((size_t)b * 2 + 1) * 128 + n];
        }
    }
}

// ---- packed ("pair") layouts ----------------------------------------------------------------------------------------
// Every buffer this path owns is laid out so that a thread's two consecutive register elements (2j, 2j+1) are ONE 16-byte
// word: float4 index j*512 + t inside a tile (t = yl*128 + x in layout R, t = threadIdx in layout F).  The global phases are
// bound by (LG instruction-queue slots) / (L2 latency), so halving the instruction count matters more than the bytes.
//   stash, farF, phisF, HF, PhatF, gPhatF : pairs of complex
//   gOpack[m][z][Y][X] = (contribution to gO[Y][X], contribution to gO[Y+4][X])  -- one red.global.add.v4.f32
__device__ __forceinline__ int p2_index(int u, int t) { return (((u >> 1) * 512 + t) << 1) + (u & 1); }   // float2 index
__device__ __forceinline__ int p4_index(int u, int t) { return (((u >> 2) * 512 + t) << 2) + (u & 3); }   // float index

__device__ __forceinline__ void f_coords(int i, int& ky, int& kx, int& u, int& t) {   // i = u*512 + t
    u = i >> 9; t = i & 511;
    const int w2 = t >> 5, lane = t & 31;
    ky = w2 + 16 * (lane >> 4) + 32 * ((lane >> 2) & 3);
    kx = (lane & 3) + 4 * u;
}
// srcT is [kx][ky] (the general path's transposed spectra) -> pair layout F, scaled
__global__ void k_permute_to_F(const float2* __restrict__ srcT, float2* __restrict__ dstF, float scale) {
    const int c = blockIdx.y;
    int ky, kx, u, t;
    f_coords(blockIdx.x * blockDim.x + threadIdx.x, ky, kx, u, t);
    dstF[(size_t)c * TILE + p2_index(u, t)] = cscale(srcT[(size_t)c * TILE + kx * 128 + ky], scale);
}
__global__ void k_unpermute_from_F(const float2* __restrict__ srcF, float2* __restrict__ dstT) {
    const int c = blockIdx.y;
    int ky, kx, u, t;
    f_coords(blockIdx.x * blockDim.x + threadIdx.x, ky, kx, u, t);
    dstT[(size_t)c * TILE + kx * 128 + ky] = srcF[(size_t)c * TILE + p2_index(u, t)];
}

// dp <- eps, arrival counters <- 0: the forward CTAs ADD their mode's intensities into dp (red.global.add.v4.f32)
__global__ void k_dp_init(float4* __restrict__ dp4, size_t n4, float eps, int* __restrict__ counter, int B) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < n4) dp4[i] = make_float4(eps, eps, eps, eps);
    if (i < (size_t)B) counter[i] = 0;
}

// g_a = Re(gO e^{-i phi}), g_phi = Im(gO conj(O)) with gO[Y][X] = gOpack[Y][X].xy + gOpack[Y-4][X].zw
__global__ void k_obj_finish_pack(const float4* __restrict__ gOp, const float* __restrict__ a, const float* __restrict__ ph,
                                  float* __restrict__ ga, float* __restrict__ gp, int Noy, int Nox, size_t n, const float* __restrict__ scale, int add) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float sc = scale ? scale[0] : 1.0f;           // batch-level factor of dL/dI (chunked steps), else 1
    const int Y = int((i / Nox) % Noy);
    const float4 lo = gOp[i];
    float2 g = make_float2(lo.x, lo.y);
    if (Y >= 4) { const float4 up = gOp[i - (size_t)4 * Nox]; g.x += up.z; g.y += up.w; }
    float s, c;
    sincosf(ph[i], &s, &c);
    const float va = sc * (g.x * c + g.y * s), vp = sc * a[i] * (g.y * c - g.x * s);
    if (add) { ga[i] += va; gp[i] += vp; }      // the gradient arrays already hold other terms (loss_sparse, written early on a side stream)
    else { ga[i] = va; gp[i] = vp; }
}

// ---- memory helpers -----------------------------------------------------------------------------------------------
// fire-and-forget vector reductions; no "memory" clobber: nothing in these kernels reads the target back, and a clobber
// would stop the compiler from hoisting the next loads above it
__device__ __forceinline__ void red_f2(float2* addr, float2 v) {
    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(addr), "f"(v.x), "f"(v.y));
}
__device__ __forceinline__ void red_f4(float4* addr, float2 a, float2 b) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
}
// L2 prefetch of a contiguous range (TMA bulk prefetch: one instruction, no registers, no smem)
__device__ __forceinline__ void l2_prefetch(const void* p, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes));
}
// whole 128 KB tile: 32 lanes of warp 0 x 4 KB
__device__ __forceinline__ void l2_prefetch_tile(const float4* tile) {
    if (threadIdx.x < 32) l2_prefetch(tile + threadIdx.x * 256, 4096);
}
// TMA bulk copy shared -> global (asynchronous, issued by one lane; no LSU store traffic, no register reads at drain time)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bulk_store(void* gdst, const void* ssrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// stash tile layout: [warp][j][lane] of 16-byte pairs -> each warp's data of 8 consecutive j is one contiguous 4 KB block
__device__ __forceinline__ int stash_index(int t, int j) { return (((t >> 5) * 16 + j) << 5) + (t & 31); }

__device__ __forceinline__ float2 lo2(float4 q) { return make_float2(q.x, q.y); }
__device__ __forceinline__ float2 hi2(float4 q) { return make_float2(q.z, q.w); }
__device__ __forceinline__ float4 pack2(float2 a, float2 b) { return make_float4(a.x, a.y, b.x, b.y); }

// The object ROI of the NEXT pointwise phase is copied global -> shared memory asynchronously (cp.async: no registers, no wait)
// while the last register DFT of the inverse FFT runs, into the 32 exchange-buffer slots the thread has just read.  (16-byte copies
// from a pair-packed object copy, two lanes sharing their slots, were measured 15 % slower: profiles/r02/ab_opair16.txt.)
__device__ __forceinline__ void cp_async8(uint32_t saddr, const void* g) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(saddr), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
// O_z ROI -> this thread's own slots of E: slot k holds O_z[cy + yl + 4k][cx + x]  (Oz points at row cy + yl, column cx + x)
__device__ __forceinline__ void prefetch_roi_to_E(float2* E, const Geo& g, const float2* __restrict__ Oz, int Nox) {
    const uint32_t s0 = smem_u32(E + g.yl * 128 + g.x);
#pragma unroll
    for (int k = 0; k < 32; ++k) cp_async8(s0 + k * (CH * 8), Oz + (size_t)(4 * k) * Nox);
    cp_async_commit();
}

#ifndef F128_CHK
#define F128_CHK 4
#endif
constexpr int CH2 = F128_CHK;     // pointwise phases load CH2 16-byte words per stream ahead of use

// ---- forward ------------------------------------------------------------------------------------------------------
// TILT: per-sample tilt ramps multiply the propagator; PHIS: the Fourier-domain waves are kept for the tilt / thickness gradients.
// Compile-time switches: as runtime flags inside the unrolled pointwise loops they cost a branch + convergence barrier per element
// (BSSY/BSYNC/BRA/UMOV were 8 % of the forward's and 14 % of the adjoint's stall samples).
template <bool TILT, bool PHIS>
__global__ void __launch_bounds__(FT, 1) k_forward(Args a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const Smem s = carve_smem(smem_raw);
    const Geo g;
    const Dims& d = a.f.d;
    const int p = blockIdx.x, m = blockIdx.y, b = blockIdx.z;
    int cy, cx;
    roi_origin(d, a.f.crop, a.f.idx, b, cy, cx);
    const size_t plane = (size_t)d.Noy * d.Nox;
    const float2* Oroi = a.f.O + (size_t)obj_mode(d, b, m) * d.Z * plane + (size_t)(cy + g.yl) * d.Nox + cx + g.x;   // slice 0, this thread's first row
    load_tables(s, a, b);
    __syncthreads();
    const size_t tile = ((size_t)b * d.P + p) * d.M + m;
    const int tR = g.yl * 128 + g.x;
    const float2 eyv = TILT ? s.ey[g.ky] : make_float2(1.f, 0.f);
    float2 v[32];
    if (a.shift) {
        const float4* __restrict__ ph = reinterpret_cast<const float4*>(a.PhatF) + (size_t)p * (TILE / 2) + g.t;
        const float2 wyv = s.wy[g.ky];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const float4 q = __ldg(ph + j * 512);
            v[2 * j] = cmul(lo2(q), cmul(wyv, s.wx[g.kx(2 * j)]));
            v[2 * j + 1] = cmul(hi2(q), cmul(wyv, s.wx[g.kx(2 * j + 1)]));
        }
    } else {
        const float2* __restrict__ pr = a.f.probe + (size_t)p * TILE + tR;
#pragma unroll
        for (int k = 0; k < 32; ++k) v[k] = __ldg(pr + k * 512);
    }
    if (!a.shift) prefetch_roi_to_E(s.E, g, Oroi, d.Nox);        // no inverse FFT precedes slice 0: fetch its ROI now
    for (int z = a.shift ? -1 : 0; z < d.Z; ++z) {
        if (z >= 0) {
            float4* st = reinterpret_cast<float4*>(a.f.stash) + (tile * d.Z + z) * (TILE / 2);
            cp_async_wait_all();                                  // O_z sits in this thread's own slots of E
            const float2* __restrict__ Os = s.E + tR;
            // psi_z -> stash through a per-warp 4 KB staging block and one TMA bulk store per half (8 pairs): the stores
            // leave the LSU path, so the FFT's shared-memory traffic is not queued behind a 128 KB store burst
            float4* sw = reinterpret_cast<float4*>(s.fl) + (g.w2 * 8) * 32 + g.lane;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                if (g.lane == 0) bulk_wait_read0();
                __syncwarp();
#pragma unroll
                for (int j = h * 8; j < h * 8 + 8; ++j) {
                    const int k = 2 * j;
                    sw[(j - h * 8) * 32] = pack2(v[k], v[k + 1]);
                    v[k] = cmul(v[k], Os[k * CH]);
                    v[k + 1] = cmul(v[k + 1], Os[(k + 1) * CH]);
                }
                fence_async_smem();
                __syncwarp();
                if (g.lane == 0) {
                    bulk_store(st + stash_index(g.t & ~31, h * 8), sw, 4096);
                    bulk_commit();
                }
            }
            fft2_R_to_F(v, s.E, s.tw, g);
            if (z == d.Z - 1) break;
            const float4* __restrict__ hf = reinterpret_cast<const float4*>(a.HF) + g.t;
            float4* __restrict__ ph = PHIS ? reinterpret_cast<float4*>(a.phisF) + (tile * (d.Z - 1) + z) * (TILE / 2) + g.t : nullptr;
#pragma unroll
            for (int j0 = 0; j0 < 16; j0 += CH2) {
                float4 h[CH2];
#pragma unroll
                for (int i = 0; i < CH2; ++i) h[i] = __ldg(hf + (j0 + i) * 512);
#pragma unroll
                for (int i = 0; i < CH2; ++i) {
                    const int u = 2 * (j0 + i);
                    if (PHIS) ph[(j0 + i) * 512] = pack2(v[u], v[u + 1]);
                    float2 h0 = lo2(h[i]), h1 = hi2(h[i]);
                    if (TILT) { h0 = cmul(h0, cmul(eyv, s.ex[g.kx(u)])); h1 = cmul(h1, cmul(eyv, s.ex[g.kx(u + 1)])); }
                    v[u] = cmul(v[u], h0);
                    v[u + 1] = cmul(v[u + 1], h1);
                }
            }
        }
        {
            const float2* On = Oroi + (size_t)(z + 1) * plane;    // the ROI the pointwise phase after this inverse FFT multiplies
            fft2_F_to_R(v, s.E, s.tw, g, [&] { prefetch_roi_to_E(s.E, g, On, d.Nox); });
        }
    }
    // far field: the spectrum itself is kept for the adjoint; this mode's intensity occu_m |Psi|^2 / N^2 is ADDED into dp (pre-set to
    // eps by k_dp_init) -- the mode reduction of forward.py:79 happens in L2, no partial-intensity buffer, no reduction launch
    const float oc = a.f.occu[m] * (1.0f / (128.0f * 128.0f));
    const size_t ft = ((size_t)b * d.M + m) * d.P + p;
    float4* __restrict__ ff = reinterpret_cast<float4*>(a.farF) + ft * (TILE / 2) + g.t;
#pragma unroll
    for (int j = 0; j < 16; ++j) ff[j * 512] = pack2(v[2 * j], v[2 * j + 1]);
    // layout F -> natural order through the (now idle) exchange buffer, so that the reds are whole 16-byte words of a row:
    // float at [ky][kx ^ swz(ky)], swz flips the bank bits that the lanes' different ky would otherwise share
    float* Ef = reinterpret_cast<float*>(s.E);
    __syncthreads();                                   // every warp is done with its E2 reads of the last forward FFT
    {
        const int swz = (((g.ky >> 5) & 3) << 2) ^ (((g.ky >> 4) & 1) << 4);
        float* row = Ef + g.ky * 128;
#pragma unroll
        for (int u = 0; u < 32; ++u) row[g.kx(u) ^ swz] = oc * cabs2(v[u]);
    }
    __syncthreads();
    float* dpb = a.f.dp + (size_t)b * TILE;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int i4 = g.t + 512 * j, ky = i4 >> 5, kx0 = (i4 & 31) << 2;
        const int swz = (((ky >> 5) & 3) << 2) ^ (((ky >> 4) & 1) << 4);
        const float4 q = *reinterpret_cast<const float4*>(Ef + ky * 128 + (kx0 ^ swz));
        red_f4(reinterpret_cast<float4*>(dpb + ((ky + 64) & 127) * 128 + ((kx0 + 64) & 127)), make_float2(q.x, q.y), make_float2(q.z, q.w));
    }
    if (a.f.lf.on) {
        // fused loss (north star item 3): the LAST of this pattern's M*P CTAs to arrive finds the finished intensities in L2 and adds
        // the pattern's contribution to the batch sums of the data losses (what k_loss_partial does in the unfused sequence)
        __threadfence();                               // this CTA's reds are performed before its arrival is counted
        __syncthreads();
        __shared__ int s_last;
        if (threadIdx.x == 0) s_last = atomicAdd(a.f.lf.counter + b, 1) == d.M * d.P - 1;
        __syncthreads();
        if (s_last) {
            __threadfence();
            const LossFuse& lf = a.f.lf;
            const float* M_ = lf.mv.meas + (size_t)lf.rows[b] * lf.mv.Hs * lf.mv.Ws;
            float acc5[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
            if (meas_plain(lf.mv)) {
                // cold code (one CTA per pattern) kept ROLLED: unrolled, its pow / log expansions made the kernel 482 KB and the forward
                // 3.5 % slower (1.215 -> 1.173 ms at C2); out of line (__noinline__) it costs a 480-byte frame and is slower again (1.224)
#pragma unroll 1
                for (int j = 0; j < 8; ++j) {
                    const int q = g.t + 512 * j;
                    const float4 i4 = __ldcg(reinterpret_cast<const float4*>(dpb) + q);
                    const float4 m4 = __ldg(reinterpret_cast<const float4*>(M_) + q);
                    loss_pixel(lf.k, i4.x, m4.x, 4 * q, TILE, lf.pac, acc5); loss_pixel(lf.k, i4.y, m4.y, 4 * q + 1, TILE, lf.pac, acc5);
                    loss_pixel(lf.k, i4.z, m4.z, 4 * q + 2, TILE, lf.pac, acc5); loss_pixel(lf.k, i4.w, m4.w, 4 * q + 3, TILE, lf.pac, acc5);
                }
            } else {
                for (int pix = g.t; pix < TILE; pix += FT)
                    loss_pixel(lf.k, __ldcg(dpb + pix), meas_at(lf.mv, M_, pix >> 7, pix & 127), pix, TILE, lf.pac, acc5);
            }
            float* red5 = Ef;                          // 5 x 32 floats of scratch: the exchange buffer is idle
            __syncthreads();
            block_sum<5>(acc5, red5);
            if (threadIdx.x == 0) loss_stats_commit(lf.k, acc5, lf.stats);
        }
    }
    if (g.lane == 0) bulk_wait_all();       // the staging blocks must outlive the TMA reads; writes complete before exit
}

// ---- adjoint --------------------------------------------------------------------------------------------------------
// pointwise phase after the inverse FFT.  MODE 3: scatter conj(psi_z) gphi_z into the dense gradient with vector reds (handing the
// same data to the TMA engine as cp.reduce.async.bulk from staged shared memory was measured slower: 1.85 vs 1.67 ms per C2 batch);
// MODE 4: object gradient not wanted.
// O_z comes from this thread's own slots of E (prefetched during the inverse FFT); only the stash streams through registers.
// (Staging half of the stash tile in shared memory by TMA a whole FFT pair ahead -- cp.async.bulk + one mbarrier per warp -- was
// measured 16 % SLOWER, profiles/r02/ab_tma_stash_adjoint.txt: the kernel sits at the 128-register limit and the extra live state
// doubles the spills inside the DFTs.)
template <int MODE>
__device__ __forceinline__ void accum_phase_E(float2 (&v)[32], const float4* __restrict__ st, const float2* __restrict__ Os, size_t ostr,
                                              float4* __restrict__ gOz) {
    constexpr int CS = 8;
#pragma unroll
    for (int j0 = 0; j0 < 16; j0 += CS) {
        float4 ps[CS];
        if (MODE != 4) {
#pragma unroll
            for (int i = 0; i < CS; ++i) ps[i] = __ldg(st + (j0 + i) * 32);
        }
#pragma unroll
        for (int i = 0; i < CS; ++i) {
            const int k = 2 * (j0 + i);
            if (MODE != 4) {
                const float2 c0 = cmulc(v[k], lo2(ps[i])), c1 = cmulc(v[k + 1], hi2(ps[i]));     // conj(psi) * gphi
                red_f4(gOz + (j0 + i) * ostr, c0, c1);
            }
            v[k] = cmulc(v[k], Os[k * CH]);                        // gpsi_z = conj(O_z) gphi_z
            v[k + 1] = cmulc(v[k + 1], Os[(k + 1) * CH]);
        }
    }
}

// One CTA per unit (sample, object mode, probe mode); every mode scatters its conj(psi_z) gphi_z straight into the dense (L2-resident)
// object gradient with red.global.add.v4.f32 -- no accumulator traffic, all units independent.  (Accumulating over the probe modes
// in a CTA-private scratch first was measured slower -- 2.1 vs 1.63 ms at C2, the 148 MB of accumulators do not stay in L2 -- and
// was removed.)
// Per unit the loop runs "steps" s = Z .. 0 with ONE forward/inverse FFT call site:
//   s = Z     : (stashed) F2(psi_{Z-1} O_{Z-1}) * (2 occu G~ / N^2)     -> inverse -> gphi_{Z-1}
//   s = z >= 1: F2(gpsi_z) * conj(H_n)/N^2 [+ propagator-gradient sums] -> inverse -> gphi_{z-1}
//   s = 0     : F2(gpsi_0) -> probe-spectrum and shift gradients (only with shifted probes)
// TILT / PROP (tilt ramps on the propagator / tilt + thickness gradient sums) are compile-time, see k_forward.
template <bool TILT, bool PROP>
__global__ void __launch_bounds__(FT, 1) k_backward(Args a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const Smem s = carve_smem(smem_raw);
    const Geo g;
    const Dims& d = a.f.d;
    const int tR = g.yl * 128 + g.x;
    const bool want_probe_fft = a.shift && (a.need_probe || a.need_shift);
    const size_t plane = (size_t)d.Noy * d.Nox;
    for (int unit = blockIdx.x; unit < a.units; unit += gridDim.x) {
        const int p = unit % d.P, bm = unit / d.P, b = bm / d.M, m = bm % d.M;
        const int64_t n0 = a.f.idx[b];
        int cy, cx;
        roi_origin(d, a.f.crop, a.f.idx, b, cy, cx);
        const float2* Oroi = a.f.O + (size_t)obj_mode(d, b, m) * d.Z * plane + (size_t)(cy + g.yl) * d.Nox + cx + g.x;
        __syncthreads();
        load_tables(s, a, b);
        // dL/dI in layout F, scaled 2 occu_m G~ / N^2, is gathered from global memory in the (single) start phase per mode
        const float gsc = 2.0f * a.f.occu[m] * (1.0f / (128.0f * 128.0f));
        const float* __restrict__ Grow = a.G + (size_t)b * TILE + ((g.ky + 64) & 127) * 128;
        __syncthreads();
        const float2 eyv = TILT ? s.ey[g.ky] : make_float2(1.f, 0.f);
        const size_t ostr = (size_t)8 * d.Nox;
        const size_t roi0 = (size_t)(cy + g.yl) * d.Nox + cx + g.x;
        float s3[3] = {0.f, 0.f, 0.f};                  // Ky S, Kx S, (Kz-k0) S
        const size_t tile = ((size_t)b * d.P + p) * d.M + m;
        const float4* stash_t = reinterpret_cast<const float4*>(a.f.stash) + tile * d.Z * (TILE / 2);
        float2 v[32];
        for (int st_i = d.Z; st_i >= 0; --st_i) {
            if (st_i == 0 && !want_probe_fft) break;
            // prefetch what the pointwise phase after the inverse FFT will read: slice zn = (st_i == Z ? Z-1 : st_i-1)
            const int zn = st_i == d.Z ? d.Z - 1 : st_i - 1;
            if (st_i > 0 && a.need_obj) l2_prefetch_tile(stash_t + (size_t)zn * (TILE / 2));
            if (st_i < d.Z) fft2_R_to_F(v, s.E, s.tw, g);
            if (st_i == d.Z) {
                const float4* __restrict__ ff = reinterpret_cast<const float4*>(a.farF) + (((size_t)b * d.M + m) * d.P + p) * (TILE / 2) + g.t;
#pragma unroll
                for (int j0 = 0; j0 < 16; j0 += CH2) {
                    float4 f[CH2];
#pragma unroll
                    for (int i = 0; i < CH2; ++i) f[i] = __ldg(ff + (j0 + i) * 512);
#pragma unroll
                    for (int i = 0; i < CH2; ++i) {
                        const int u = 2 * (j0 + i);
                        v[u] = cscale(lo2(f[i]), gsc * __ldg(Grow + ((g.kx(u) + 64) & 127)));
                        v[u + 1] = cscale(hi2(f[i]), gsc * __ldg(Grow + ((g.kx(u + 1) + 64) & 127)));
                    }
                }
            } else if (st_i >= 1) {
                const float4* __restrict__ hf = reinterpret_cast<const float4*>(a.HF) + g.t;
                const float4* __restrict__ ph = PROP ? reinterpret_cast<const float4*>(a.phisF) + (tile * (d.Z - 1) + (st_i - 1)) * (TILE / 2) + g.t : nullptr;
                const float Ky = PROP ? kgrid(g.ky, 128, a.dx) : 0.f;
#pragma unroll
                for (int j0 = 0; j0 < 16; j0 += CH2) {
                    float4 h[CH2], phi[CH2];
#pragma unroll
                    for (int i = 0; i < CH2; ++i) { h[i] = __ldg(hf + (j0 + i) * 512); if (PROP) phi[i] = __ldg(ph + (j0 + i) * 512); }
#pragma unroll
                    for (int i = 0; i < CH2; ++i) {
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const int u = 2 * (j0 + i) + e;
                            float2 hh = e ? hi2(h[i]) : lo2(h[i]);
                            if (TILT) hh = cmul(hh, cmul(eyv, s.ex[g.kx(u)]));
                            v[u] = cmulc(v[u], hh);              // conj(H)/N^2 * F2(gpsi)
                            if (PROP) {
                                const float2 pz = e ? hi2(phi[i]) : lo2(phi[i]);
                                const float sv = pz.x * v[u].y - pz.y * v[u].x;
                                const float Kx = kgrid(g.kx(u), 128, a.dx);
                                const float k2 = Kx * Kx + Ky * Ky;
                                s3[0] += Ky * sv; s3[1] += Kx * sv; s3[2] += -k2 / (sqrtf(a.k0 * a.k0 - k2) + a.k0) * sv;
                            }
                        }
                    }
                }
            } else {
                // st_i == 0: v = N^2 T of gpsi_0 (shifted probes): probe-spectrum and shift gradients
                const float4* __restrict__ phf = reinterpret_cast<const float4*>(a.PhatF) + (size_t)p * (TILE / 2) + g.t;   // Phat / N^2
                float4* __restrict__ gp = reinterpret_cast<float4*>(a.gPhatF) + (size_t)p * (TILE / 2) + g.t;
                const float2 wyv = s.wy[g.ky];
                const float kapy = float((g.ky + 64) & 127) * (1.0f / 128.0f);
                const float invN2 = 1.0f / (128.0f * 128.0f);
                float r2[2] = {0.f, 0.f};
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const float4 pq = __ldg(phf + j * 512);
                    float2 cw[2];
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int u = 2 * j + e;
                        const float2 w = cmul(wyv, s.wx[g.kx(u)]);
                        cw[e] = cmulc(v[u], w);                              // conj(w') * N^2 T
                        const float2 pv = e ? hi2(pq) : lo2(pq);
                        const float qv = cw[e].y * pv.x - cw[e].x * pv.y;    // Im(conj(w') T conj(Phat))
                        r2[0] += kapy * qv;
                        r2[1] += float((g.kx(u) + 64) & 127) * (1.0f / 128.0f) * qv;
                    }
                    if (a.need_probe) red_f4(gp + j * 512, cscale(cw[0], invN2), cscale(cw[1], invN2));
                }
                if (a.need_shift) {
                    block_sum<2>(r2, s.red);
                    if (threadIdx.x == 0) {
                        atomicAdd(a.gshift + 2 * n0 + 0, -6.283185307179586f * r2[0]);
                        atomicAdd(a.gshift + 2 * n0 + 1, -6.283185307179586f * r2[1]);
                    }
                }
                break;
            }
            {
                const float4* st = stash_t + (size_t)zn * (TILE / 2) + stash_index(g.t, 0);
                float4* gOz = a.gOpack + ((size_t)obj_mode(d, b, m) * d.Z + zn) * plane + roi0;
                const float2* On = Oroi + (size_t)zn * plane;
                fft2_F_to_R(v, s.E, s.tw, g, [&] { prefetch_roi_to_E(s.E, g, On, d.Nox); });     // gphi_{zn}
                cp_async_wait_all();
                if (a.need_obj) accum_phase_E<3>(v, st, s.E + tR, ostr, gOz);
                else accum_phase_E<4>(v, st, s.E + tR, ostr, gOz);
            }
        }
        if (!a.shift && a.need_probe) {               // unshifted probes: g_probe += gpsi_0 (natural layout)
            float2* gp = a.gprobe + (size_t)p * TILE + tR;
#pragma unroll
            for (int k = 0; k < 32; ++k) red_f2(gp + k * 512, v[k]);
        }
        if (PROP) {
            block_sum<3>(s3, s.red);
            if (threadIdx.x == 0) {
                atomicAdd(a.gprop + 3 * b + 0, s3[0]);
                atomicAdd(a.gprop + 3 * b + 1, s3[1]);
                atomicAdd(a.gprop + 3 * b + 2, s3[2]);
            }
        }
    }
}

// ---- host side --------------------------------------------------------------------------------------------------------
struct Scratch {
    float2 *HF, *PhatF, *gPhatF, *farF;
    float4* gOpack;
    int* counter;           // (B) arrival counters of the fused loss
    size_t total;
};
inline Scratch carve_scratch(const ptyb200_cfg& c, int B, unsigned char* base) {
    Scratch s;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~size_t(255); return base + o; };
    const size_t obj = (size_t)((c.reserved[1] & 1) ? B : 1) * c.M * c.Z * c.Noy * c.Nox;
    s.HF = (float2*)take((size_t)TILE * 8);
    s.PhatF = (float2*)take((size_t)c.P * TILE * 8);
    s.gPhatF = (float2*)take((size_t)c.P * TILE * 8);
    s.gOpack = (float4*)take(obj * 16);
    s.counter = (int*)take((size_t)B * 4);
    s.farF = (float2*)take((size_t)B * c.M * c.P * TILE * 8);
    s.total = off;
    return s;
}
inline bool covers(const ptyb200_cfg& c) { return c.N == 128; }
inline size_t scratch_bytes(const ptyb200_cfg& c, int B) { return covers(c) ? carve_scratch(c, B, nullptr).total : 0; }

inline int fail(std::string& err, const char* what, cudaError_t e) {
    err = std::string("fused128: ") + what + ": " + cudaGetErrorString(e);
    return 1;
}
#define F128_CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail(err, #call, e_); } while (0)

inline Args make_args(const ptyb200_cfg& c, const FwdArgs& f, const Scratch& sc, float2* phis) {
    Args a;
    memset(&a, 0, sizeof a);
    a.f = f; a.HF = sc.HF; a.PhatF = sc.PhatF; a.farF = sc.farF; a.phisF = f.phis ? phis : nullptr;
    a.gOpack = sc.gOpack;
    a.gPhatF = sc.gPhatF; a.shift = c.shift_probes;
    return a;
}

// f.HT = transposed propagator [kx][ky]; f.PhatT = probe spectrum [kx][ky] (both made by setup_common)
// sH / sP: the branches on which the transposed propagator / the probe spectrum are being made (api.cu: setup_common); their permuted
// copies are made there too, and `join` brings both back into `st` just before the wave kernel
template <class Join>
inline int forward(const ptyb200_cfg& c, int B, FwdArgs f, const float* obja, const float* objp, unsigned char* scratch, cudaStream_t st,
                   std::string& err, std::atomic<long long>* launches, cudaStream_t sH, cudaStream_t sP, Join join) {
    Scratch sc = carve_scratch(c, B, scratch);
    Args a = make_args(c, f, sc, f.phis);
    const float inv = 1.0f / (128.0f * 128.0f);
    const size_t obj = (size_t)((c.reserved[1] & 1) ? B : 1) * c.M * c.Z * c.Noy * c.Nox;
    k_permute_to_F<<<dim3(TILE / 256, 1), 256, 0, sH>>>(f.HT, sc.HF, inv);
    F128_CK(cudaGetLastError()); ++*launches;
    if (c.shift_probes) {
        k_permute_to_F<<<dim3(TILE / 256, c.P), 256, 0, sP>>>(f.PhatT, sc.PhatF, inv);
        F128_CK(cudaGetLastError()); ++*launches;
    }
    a.f.lf.counter = sc.counter;
    {
        const size_t n4 = (size_t)B * TILE / 4;
        k_dp_init<<<(unsigned)((n4 + 255) / 256), 256, 0, st>>>(reinterpret_cast<float4*>(f.dp), n4, c.eps, sc.counter, B);
        F128_CK(cudaGetLastError()); ++*launches;
    }
    if (int r = join()) return r;
    const dim3 grid(c.P, c.M, B);
    const bool tilt = f.tvec != nullptr, phis = a.phisF != nullptr;
#define F128_LAUNCH_FWD(T, PH)                                                                                                     \
    do {                                                                                                                           \
        F128_CK(cudaFuncSetAttribute(k_forward<T, PH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES_FWD));         \
        k_forward<T, PH><<<grid, FT, SMEM_BYTES_FWD, st>>>(a);                                                                     \
    } while (0)
    if (tilt) { if (phis) F128_LAUNCH_FWD(true, true); else F128_LAUNCH_FWD(true, false); }
    else      { if (phis) F128_LAUNCH_FWD(false, true); else F128_LAUNCH_FWD(false, false); }
#undef F128_LAUNCH_FWD
    F128_CK(cudaGetLastError()); ++*launches;
    return 0;
}

// adjoint incl. the polar backward of the object gradient; the caller zeroes gPhatT / gprop / gshift and runs the
// probe-spectrum inverse FFT
inline int backward(const ptyb200_cfg& c, int B, const BwdArgs& bw, const float* obja, const float* objp, float* g_obja, float* g_objp,
                    unsigned char* scratch, float2* g_probe, float2* gPhatT, cudaStream_t st, std::string& err, std::atomic<long long>* launches,
                    int acc_flags = 0, const float* scale = nullptr, bool finish_only = false,
                    const std::function<cudaStream_t()>& fork_fin = nullptr) {
    // acc_flags (PTYB200_ACC_*): KEEP_GRADS = the accumulators already hold earlier chunks of the batch; NO_FINISH = leave them raw.
    // finish_only: no adjoint, only the completion of the accumulators (with the batch-level `scale` of an unscaled loss gradient).
    Scratch sc = carve_scratch(c, B, scratch);
    Args a = make_args(c, bw.f, sc, bw.f.phis);
    a.G = bw.G; a.gprop = bw.gprop; a.gshift = bw.gshift; a.gprobe = g_probe;
    a.dx = bw.dx; a.k0 = bw.k0;
    a.need_obj = bw.need_obj; a.need_probe = bw.need_probe; a.need_shift = bw.need_shift; a.need_prop = bw.need_prop;
    a.units = B * c.M * c.P;
    const size_t obj = (size_t)((c.reserved[1] & 1) ? B : 1) * c.M * c.Z * c.Noy * c.Nox;
    if (!finish_only && !(acc_flags & PTYB200_ACC_KEEP_GRADS)) {
        if (a.need_obj) F128_CK(cudaMemsetAsync(sc.gOpack, 0, obj * 16, st));
        if (a.need_probe) {
            if (c.shift_probes) F128_CK(cudaMemsetAsync(sc.gPhatF, 0, (size_t)c.P * TILE * 8, st));
            else F128_CK(cudaMemsetAsync(g_probe, 0, (size_t)c.P * TILE * 8, st));
        }
    }
    const bool tilt = bw.f.tvec != nullptr, prop = a.need_prop != 0;
#define F128_LAUNCH_BWD(T, PR)                                                                                                     \
    do {                                                                                                                           \
        F128_CK(cudaFuncSetAttribute(k_backward<T, PR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES_BWD));        \
        k_backward<T, PR><<<a.units, FT, SMEM_BYTES_BWD, st>>>(a);                                                                 \
    } while (0)
    if (!finish_only) {
        if (tilt) { if (prop) F128_LAUNCH_BWD(true, true); else F128_LAUNCH_BWD(true, false); }
        else      { if (prop) F128_LAUNCH_BWD(false, true); else F128_LAUNCH_BWD(false, false); }
        F128_CK(cudaGetLastError()); ++*launches;
    }
#undef F128_LAUNCH_BWD
    const cudaStream_t sp = fork_fin ? fork_fin() : st;        // branch of the probe-gradient chain (api.cu continues it and joins)
    if (acc_flags & PTYB200_ACC_NO_FINISH) return 0;
    if (a.need_probe && c.shift_probes) {
        k_unpermute_from_F<<<dim3(TILE / 256, c.P), 256, 0, sp>>>(sc.gPhatF, gPhatT);
        F128_CK(cudaGetLastError()); ++*launches;
    }
    if (a.need_obj) {
        k_obj_finish_pack<<<(unsigned)((obj + 255) / 256), 256, 0, st>>>(sc.gOpack, obja, objp, g_obja, g_objp, c.Noy, c.Nox, obj, scale, (acc_flags & PTYB200_ACC_ADD_OBJ) ? 1 : 0);
        F128_CK(cudaGetLastError()); ++*launches;
    }
    return 0;
}

}  // namespace fused128
}  // namespace ptyb
