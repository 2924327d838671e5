// Fused on-chip multislice kernels for N = 64 (sm_100a): the register-resident design of fused128.cuh with 16 values per thread.
//
// One CTA of 256 threads keeps one 64x64 complex wave in registers (16 complex values per thread) through every slice; shared memory
// (34 KB) is the exchange medium of the FFT.  With 32 KB of registers and ~68 KB of shared memory per CTA THREE CTAs share an SM, so
// one tile's exchange / global phases overlap the other tiles' DFTs -- the overlap that the 128^2 kernel (one CTA per SM) cannot have.
//
//   N = 4 R, R = 16;  T = 4 N = 256 threads;  NW = T / 32 = R / 2 = 8 warps
//   layout R (real space)   thread t: x = t % N, yl = t / N;              v[k] = psi[yl + 4k][x],  k < R
//   layout F (Fourier)      thread t: w2 = t >> 5, l = t & 31, rsel = l >> 4, q = (l >> 2) & 3, vv = l & 3;
//                                                                         v[u] = X[w2 + NW rsel + R q][vv + 4u],  u < R
//   forward 2-D FFT (R -> F): DFT_R over k in registers (y, stride 4)     -> exchange E1 (CTA wide): chunk r = [yl][x], R chunks
//                             y-twiddle, 4x4 DFT (rest of y and of x), x-twiddle: each half-warp takes one chunk, lane j = l & 15
//                             holds x = j + R s -> exchange E2 (half-warp local, same chunk, 16 rows padded to R + 1)
//                             DFT_R over j in registers (x)
//   inverse: the same stages backwards.  Everything else (packed pair layouts, TMA-staged stash stores, cp.async ROI prefetch into the
//   exchange slots, red.global.add.v4 gradient scatter, mode reduction of |Psi|^2 in L2 with the fused loss sums) follows fused128.cuh.
#pragma once
#include "fused128.cuh"

namespace ptyb {
namespace fused64 {

using fused128::Args;
using fused128::red_f2;
using fused128::red_f4;
using fused128::l2_prefetch;
using fused128::smem_u32;
using fused128::bulk_store;
using fused128::bulk_commit;
using fused128::bulk_wait_read0;
using fused128::bulk_wait_all;
using fused128::fence_async_smem;
using fused128::lo2;
using fused128::hi2;
using fused128::pack2;
using fused128::cp_async8;
using fused128::cp_async_commit;
using fused128::cp_async_wait_all;

constexpr int R = 16;
constexpr int FN = 4 * R;               // 64
constexpr int FT = 4 * FN;              // 256 threads per CTA
constexpr int NW = FT / 32;             // 8 warps
constexpr int CH = 4 * FN + 16;         // 272 elements per chunk (256 used by E1, 16 rows x 17 by E2)
constexpr int E_ELEMS = R * CH;         // 4352 float2 = 34816 B
constexpr int TILE = FN * FN;           // 4096
constexpr int HN = FN / 2;
constexpr int MINB = 3;                 // CTAs per SM the kernels are compiled for (85 registers per thread)
// exchange buffer + twiddle / ramp tables + reduction scratch (+ one 4 KB TMA staging block per warp in the forward)
constexpr size_t SMEM_BYTES_BWD = sizeof(float2) * (E_ELEMS + FN + 4 * FN) + 128 * sizeof(float);
constexpr size_t SMEM_BYTES_FWD = SMEM_BYTES_BWD + NW * 4096;

struct Geo {
    int t, x, yl, w2, lane, rsel, e16, ky, vv;
    __device__ __forceinline__ Geo() {
        t = threadIdx.x; x = t & (FN - 1); yl = t / FN; w2 = t >> 5; lane = t & 31;
        rsel = lane >> 4; e16 = lane & 15; vv = lane & 3;
        ky = w2 + NW * rsel + R * ((lane >> 2) & 3);
    }
    __device__ __forceinline__ int kx(int u) const { return vv + 4 * u; }
};

// tw[n] = exp(-2 pi i n / N), n < N (shared memory)
__device__ __forceinline__ void fft2_R_to_F(float2 (&v)[R], float2* E, const float2* tw, const Geo& g) {
    Dft<R, -1>::run(v);
    // No CTA barrier before these stores: thread t writes exactly the slots E[r*CH + yl*N + x] that it alone has read (layout-R reads
    // at the end of the previous inverse FFT, behind that FFT's CTA barrier) and that it alone filled with the ROI (cp.async, waited
    // for by t); every other use of E by the kernels is fenced by its own barriers.  Measured: forward 1.170 -> 1.138 ms, adjoint
    // 1.267 -> 1.250 ms at C2, S64 +1.3 % (profiles/r02/ab_first_barrier_and_early_stash.txt).
    {
        float2* p = E + g.yl * FN + g.x;
#pragma unroll
        for (int r = 0; r < R; ++r) p[r * CH] = v[r];
    }
    __syncthreads();
    const int j = g.e16, r = g.w2 + NW * g.rsel;       // this half-warp's chunk
    float2* ch = E + r * CH;
    float2 xt[4];
#pragma unroll
    for (int c = 1; c < 4; ++c) xt[c] = tw[j * c];
    {
        float2 a[4][4];                                // [yl][s]
#pragma unroll
        for (int y = 0; y < 4; ++y)
#pragma unroll
            for (int s = 0; s < 4; ++s) a[y][s] = ch[y * FN + j + R * s];
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
        for (int q = 0; q < 4; ++q) {                  // DFT4 over s -> vv, x-twiddle W_N^(j c), store transposed
            Dft<4, -1>::run(a[q]);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                float2 o = c ? cmul(a[q][c], xt[c]) : a[q][c];
                ch[(q * 4 + c) * (R + 1) + j] = o;
            }
        }
    }
    __syncwarp();
    {
        const float2* p = ch + g.e16 * (R + 1);
#pragma unroll
        for (int jj = 0; jj < R; ++jj) v[jj] = p[jj];
    }
    Dft<R, -1>::run(v);
}

// `pre` runs between the last shared-memory read and the last register DFT: from there on the R slots this thread has just read
// (E[r*CH + yl*N + x]) belong to it alone until the next forward FFT's first barrier (the ROI prefetch is parked there).
template <class Pre>
__device__ __forceinline__ void fft2_F_to_R(float2 (&v)[R], float2* E, const float2* tw, const Geo& g, Pre pre) {
    Dft<R, +1>::run(v);
    const int j = g.e16, r = g.w2 + NW * g.rsel;
    float2* ch = E + r * CH;
    // no CTA barrier here: this half-warp only writes its OWN chunk, whose only foreign readers are the layout-R reads at the end of
    // an earlier inverse FFT, and a forward FFT (whose CTA barrier every thread passes after those reads) always runs between two inverse FFTs
    __syncwarp();
    {
        float2* p = ch + g.e16 * (R + 1);
#pragma unroll
        for (int jj = 0; jj < R; ++jj) p[jj] = v[jj];
    }
    __syncwarp();
    float2 xt[4];
#pragma unroll
    for (int c = 1; c < 4; ++c) xt[c] = tw[j * c];
    {
        float2 a[4][4];                                // [q][vv]
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                float2 o = ch[(q * 4 + c) * (R + 1) + j];
                a[q][c] = c ? cmulc(o, xt[c]) : o;
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
                ch[y * FN + j + R * s] = o;
            }
        }
    }
    __syncthreads();
    {
        const float2* p = E + g.yl * FN + g.x;
#pragma unroll
        for (int rr = 0; rr < R; ++rr) v[rr] = p[rr * CH];
    }
    pre();
    Dft<R, +1>::run(v);
}

// ---- shared memory carve ----------------------------------------------------------------------------------------
struct Smem {
    float2 *E, *tw, *wy, *wx, *ey, *ex;
    float* red;     // 128 floats (block_sum scratch)
    float* fl;      // forward only: one 4 KB TMA staging block per warp
};
__device__ __forceinline__ Smem carve_smem(unsigned char* raw) {
    Smem s;
    s.E = reinterpret_cast<float2*>(raw);
    s.tw = s.E + E_ELEMS;
    s.wy = s.tw + FN; s.wx = s.wy + FN; s.ey = s.wx + FN; s.ex = s.ey + FN;
    s.red = reinterpret_cast<float*>(s.ex + FN);
    s.fl = s.red + 128;
    return s;
}
__device__ __forceinline__ void load_tables(const Smem& s, const Args& a, int b) {
    for (int n = threadIdx.x; n < FN; n += blockDim.x) {
        float sn, cs;
        sincospif(-2.0f * float(n) / float(FN), &sn, &cs);
        s.tw[n] = make_float2(cs, sn);
        if (a.f.wvec) {
            s.wy[n] = a.f.wvec[((size_t)b * 2 + 0) * FN + n];
            s.wx[n] = a.f.wvec[((size_t)b * 2 + 1) * FN + n];
        }
        if (a.f.tvec) {
            s.ey[n] = a.f.tvec[((size_t)b * 2 + 0) * FN + n];
            s.ex[n] = a.f.tvec[((size_t)b * 2 + 1) * FN + n];
        }
    }
}

// ---- packed pair layouts: float4 index j*T + t holds a thread's register elements (2j, 2j+1) ----------------------------------
__device__ __forceinline__ int p2_index(int u, int t) { return (((u >> 1) * FT + t) << 1) + (u & 1); }   // float2 index
__device__ __forceinline__ void f_coords(int i, int& ky, int& kx, int& u, int& t) {   // i = u*T + t
    u = i / FT; t = i % FT;
    const int w2 = t >> 5, lane = t & 31;
    ky = w2 + NW * (lane >> 4) + R * ((lane >> 2) & 3);
    kx = (lane & 3) + 4 * u;
}
// srcT is [kx][ky] (the general path's transposed spectra) -> pair layout F, scaled
__global__ void k_permute_to_F(const float2* __restrict__ srcT, float2* __restrict__ dstF, float scale) {
    const int c = blockIdx.y;
    int ky, kx, u, t;
    f_coords(blockIdx.x * blockDim.x + threadIdx.x, ky, kx, u, t);
    dstF[(size_t)c * TILE + p2_index(u, t)] = cscale(srcT[(size_t)c * TILE + kx * FN + ky], scale);
}
__global__ void k_unpermute_from_F(const float2* __restrict__ srcF, float2* __restrict__ dstT) {
    const int c = blockIdx.y;
    int ky, kx, u, t;
    f_coords(blockIdx.x * blockDim.x + threadIdx.x, ky, kx, u, t);
    dstT[(size_t)c * TILE + kx * FN + ky] = srcF[(size_t)c * TILE + p2_index(u, t)];
}

// whole 32 KB tile: 8 lanes of warp 0 x 4 KB
__device__ __forceinline__ void l2_prefetch_tile(const float4* tile) {
    if (threadIdx.x < TILE / 512) l2_prefetch(tile + threadIdx.x * 256, 4096);
}
// stash tile layout: [warp][j][lane] of 16-byte pairs -> each warp's R/2 = 8 pairs are one contiguous 4 KB block
__device__ __forceinline__ int stash_index(int t, int j) { return (((t >> 5) * (R / 2) + j) << 5) + (t & 31); }

// O_z ROI -> this thread's own slots of E: slot k holds O_z[cy + yl + 4k][cx + x]  (Oz points at row cy + yl, column cx + x)
__device__ __forceinline__ void prefetch_roi_to_E(float2* E, const Geo& g, const float2* __restrict__ Oz, int Nox) {
    const uint32_t s0 = smem_u32(E + g.yl * FN + g.x);
#pragma unroll
    for (int k = 0; k < R; ++k) cp_async8(s0 + k * (CH * 8), Oz + (size_t)(4 * k) * Nox);
    cp_async_commit();
}

// ---- forward: grid (P, M, B) -------------------------------------------------------------------------------------------
template <bool TILT, bool PHIS>
__global__ void __launch_bounds__(FT, MINB) k_forward(Args a) {
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
    const int tR = g.yl * FN + g.x;
    const float2 eyv = TILT ? s.ey[g.ky] : make_float2(1.f, 0.f);
    float2 v[R];
    if (a.shift) {
        const float4* __restrict__ ph = reinterpret_cast<const float4*>(a.PhatF) + (size_t)p * (TILE / 2) + g.t;
        const float2 wyv = s.wy[g.ky];
#pragma unroll
        for (int j = 0; j < R / 2; ++j) {
            const float4 q = __ldg(ph + j * FT);
            v[2 * j] = cmul(lo2(q), cmul(wyv, s.wx[g.kx(2 * j)]));
            v[2 * j + 1] = cmul(hi2(q), cmul(wyv, s.wx[g.kx(2 * j + 1)]));
        }
    } else {
        const float2* __restrict__ pr = a.f.probe + (size_t)p * TILE + tR;
#pragma unroll
        for (int k = 0; k < R; ++k) v[k] = __ldg(pr + k * FT);
    }
    if (!a.shift) prefetch_roi_to_E(s.E, g, Oroi, d.Nox);        // no inverse FFT precedes slice 0: fetch its ROI now
    for (int z = a.shift ? -1 : 0; z < d.Z; ++z) {
        if (z >= 0) {
            float4* st = reinterpret_cast<float4*>(a.f.stash) + (tile * d.Z + z) * (TILE / 2);
            cp_async_wait_all();                                  // O_z sits in this thread's own slots of E
            const float2* __restrict__ Os = s.E + tR;
            // psi_z -> stash through the warp's 4 KB staging block and ONE TMA bulk store (8 pairs per thread)
            float4* sw = reinterpret_cast<float4*>(s.fl) + (g.w2 * 8) * 32 + g.lane;
            if (g.lane == 0) bulk_wait_read0();
            __syncwarp();
#pragma unroll
            for (int j = 0; j < R / 2; ++j) {
                const int k = 2 * j;
                sw[j * 32] = pack2(v[k], v[k + 1]);
                v[k] = cmul(v[k], Os[k * CH]);
                v[k + 1] = cmul(v[k + 1], Os[(k + 1) * CH]);
            }
            fence_async_smem();
            __syncwarp();
            if (g.lane == 0) {
                bulk_store(st + stash_index(g.t & ~31, 0), sw, 4096);
                bulk_commit();
            }
            fft2_R_to_F(v, s.E, s.tw, g);
            if (z == d.Z - 1) break;
            const float4* __restrict__ hf = reinterpret_cast<const float4*>(a.HF) + g.t;
            float4* __restrict__ ph = PHIS ? reinterpret_cast<float4*>(a.phisF) + (tile * (d.Z - 1) + z) * (TILE / 2) + g.t : nullptr;
            float4 h[R / 2];
#pragma unroll
            for (int j = 0; j < R / 2; ++j) h[j] = __ldg(hf + j * FT);
#pragma unroll
            for (int j = 0; j < R / 2; ++j) {
                const int u = 2 * j;
                if (PHIS) ph[j * FT] = pack2(v[u], v[u + 1]);
                float2 h0 = lo2(h[j]), h1 = hi2(h[j]);
                if (TILT) { h0 = cmul(h0, cmul(eyv, s.ex[g.kx(u)])); h1 = cmul(h1, cmul(eyv, s.ex[g.kx(u + 1)])); }
                v[u] = cmul(v[u], h0);
                v[u + 1] = cmul(v[u + 1], h1);
            }
        }
        {
            const float2* On = Oroi + (size_t)(z + 1) * plane;    // the ROI the pointwise phase after this inverse FFT multiplies
            fft2_F_to_R(v, s.E, s.tw, g, [&] { prefetch_roi_to_E(s.E, g, On, d.Nox); });
        }
    }
    // far field: the spectrum is kept for the adjoint; this mode's intensity occu_m |Psi|^2 / N^2 is ADDED into dp (pre-set to eps by
    // k_dp_init): the mode reduction of forward.py:79 happens in L2
    const float oc = a.f.occu[m] * (1.0f / float(TILE));
    const size_t ft = ((size_t)b * d.M + m) * d.P + p;
    float4* __restrict__ ff = reinterpret_cast<float4*>(a.farF) + ft * (TILE / 2) + g.t;
#pragma unroll
    for (int j = 0; j < R / 2; ++j) ff[j * FT] = pack2(v[2 * j], v[2 * j + 1]);
    // layout F -> natural order through the (now idle) exchange buffer (rows padded to N + 4 floats), then whole 16-byte words of a
    // row: float at [ky][kx ^ swz(ky)], swz flips the bank bits that the lanes' different ky (q, rsel) would otherwise share
    float* Ef = reinterpret_cast<float*>(s.E);
    __syncthreads();                                   // every warp is done with its E2 reads of the last forward FFT
    {
        const int swz = (((g.ky >> 4) & 3) << 2) ^ (((g.ky >> 3) & 1) << 4);
        float* row = Ef + g.ky * (FN + 4);
#pragma unroll
        for (int u = 0; u < R; ++u) row[g.kx(u) ^ swz] = oc * cabs2(v[u]);
    }
    __syncthreads();
    float* dpb = a.f.dp + (size_t)b * TILE;
#pragma unroll
    for (int j = 0; j < TILE / 4 / FT; ++j) {
        const int i4 = g.t + FT * j, ky = i4 / (FN / 4), kx0 = (i4 % (FN / 4)) << 2;
        const int swz = (((ky >> 4) & 3) << 2) ^ (((ky >> 3) & 1) << 4);
        const float4 q = *reinterpret_cast<const float4*>(Ef + ky * (FN + 4) + (kx0 ^ swz));
        red_f4(reinterpret_cast<float4*>(dpb + ((ky + HN) & (FN - 1)) * FN + ((kx0 + HN) & (FN - 1))), make_float2(q.x, q.y), make_float2(q.z, q.w));
    }
    if (a.f.lf.on) {
        // fused loss: the LAST of this pattern's M*P CTAs to arrive finds the finished intensities in L2 and adds the pattern's
        // contribution to the batch sums of the data losses
        __threadfence();
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
#pragma unroll 1                                       // cold code (one CTA per pattern): kept rolled, see fused128.cuh
                for (int j = 0; j < TILE / 4 / FT; ++j) {
                    const int q = g.t + FT * j;
                    const float4 i4 = __ldcg(reinterpret_cast<const float4*>(dpb) + q);
                    const float4 m4 = __ldg(reinterpret_cast<const float4*>(M_) + q);
                    loss_pixel(lf.k, i4.x, m4.x, 4 * q, TILE, lf.pac, acc5); loss_pixel(lf.k, i4.y, m4.y, 4 * q + 1, TILE, lf.pac, acc5);
                    loss_pixel(lf.k, i4.z, m4.z, 4 * q + 2, TILE, lf.pac, acc5); loss_pixel(lf.k, i4.w, m4.w, 4 * q + 3, TILE, lf.pac, acc5);
                }
            } else {
                for (int pix = g.t; pix < TILE; pix += FT)
                    loss_pixel(lf.k, __ldcg(dpb + pix), meas_at(lf.mv, M_, pix / FN, pix % FN), pix, TILE, lf.pac, acc5);
            }
            float* red5 = Ef;                          // 5 x 32 floats of scratch: the exchange buffer is idle
            __syncthreads();
            block_sum<5>(acc5, red5);
            if (threadIdx.x == 0) loss_stats_commit(lf.k, acc5, lf.stats);
        }
    }
    if (g.lane == 0) bulk_wait_all();       // the staging blocks must outlive the TMA reads; writes complete before exit
}

// ---- adjoint: one CTA per unit (sample, object mode, probe mode), see fused128.cuh -----------------------------------------
template <int MODE>
__device__ __forceinline__ void accum_phase_E(float2 (&v)[R], const float4* __restrict__ st, const float2* __restrict__ Os, size_t ostr,
                                              float4* __restrict__ gOz) {
    float4 ps[R / 2];
    if (MODE != 4) {
#pragma unroll
        for (int i = 0; i < R / 2; ++i) ps[i] = __ldg(st + i * 32);
    }
#pragma unroll
    for (int i = 0; i < R / 2; ++i) {
        const int k = 2 * i;
        if (MODE != 4) {
            const float2 c0 = cmulc(v[k], lo2(ps[i])), c1 = cmulc(v[k + 1], hi2(ps[i]));     // conj(psi) * gphi
            red_f4(gOz + i * ostr, c0, c1);
        }
        v[k] = cmulc(v[k], Os[k * CH]);                        // gpsi_z = conj(O_z) gphi_z
        v[k + 1] = cmulc(v[k + 1], Os[(k + 1) * CH]);
    }
}

template <bool TILT, bool PROP>
__global__ void __launch_bounds__(FT, MINB) k_backward(Args a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const Smem s = carve_smem(smem_raw);
    const Geo g;
    const Dims& d = a.f.d;
    const int tR = g.yl * FN + g.x;
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
        const float gsc = 2.0f * a.f.occu[m] * (1.0f / float(TILE));
        const float* __restrict__ Grow = a.G + (size_t)b * TILE + ((g.ky + HN) & (FN - 1)) * FN;
        __syncthreads();
        const float2 eyv = TILT ? s.ey[g.ky] : make_float2(1.f, 0.f);
        const size_t ostr = (size_t)8 * d.Nox;          // pair j = rows yl + 8j and yl + 8j + 4 of the ROI
        const size_t roi0 = (size_t)(cy + g.yl) * d.Nox + cx + g.x;
        float s3[3] = {0.f, 0.f, 0.f};                  // Ky S, Kx S, (Kz-k0) S
        const size_t tile = ((size_t)b * d.P + p) * d.M + m;
        const float4* stash_t = reinterpret_cast<const float4*>(a.f.stash) + tile * d.Z * (TILE / 2);
        float2 v[R];
        for (int st_i = d.Z; st_i >= 0; --st_i) {
            if (st_i == 0 && !want_probe_fft) break;
            const int zn = st_i == d.Z ? d.Z - 1 : st_i - 1;
            if (st_i > 0 && a.need_obj) l2_prefetch_tile(stash_t + (size_t)zn * (TILE / 2));
            if (st_i < d.Z) fft2_R_to_F(v, s.E, s.tw, g);
            if (st_i == d.Z) {
                const float4* __restrict__ ff = reinterpret_cast<const float4*>(a.farF) + (((size_t)b * d.M + m) * d.P + p) * (TILE / 2) + g.t;
                float4 f[R / 2];
#pragma unroll
                for (int j = 0; j < R / 2; ++j) f[j] = __ldg(ff + j * FT);
#pragma unroll
                for (int j = 0; j < R / 2; ++j) {
                    const int u = 2 * j;
                    v[u] = cscale(lo2(f[j]), gsc * __ldg(Grow + ((g.kx(u) + HN) & (FN - 1))));
                    v[u + 1] = cscale(hi2(f[j]), gsc * __ldg(Grow + ((g.kx(u + 1) + HN) & (FN - 1))));
                }
            } else if (st_i >= 1) {
                const float4* __restrict__ hf = reinterpret_cast<const float4*>(a.HF) + g.t;
                const float4* __restrict__ ph = PROP ? reinterpret_cast<const float4*>(a.phisF) + (tile * (d.Z - 1) + (st_i - 1)) * (TILE / 2) + g.t : nullptr;
                const float Ky = PROP ? kgrid(g.ky, FN, a.dx) : 0.f;
#pragma unroll
                for (int j = 0; j < R / 2; ++j) {
                    const float4 h = __ldg(hf + j * FT);
                    float4 phi = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (PROP) phi = __ldg(ph + j * FT);
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int u = 2 * j + e;
                        float2 hh = e ? hi2(h) : lo2(h);
                        if (TILT) hh = cmul(hh, cmul(eyv, s.ex[g.kx(u)]));
                        v[u] = cmulc(v[u], hh);              // conj(H)/N^2 * F2(gpsi)
                        if (PROP) {
                            const float2 pz = e ? hi2(phi) : lo2(phi);
                            const float sv = pz.x * v[u].y - pz.y * v[u].x;
                            const float Kx = kgrid(g.kx(u), FN, a.dx);
                            const float k2 = Kx * Kx + Ky * Ky;
                            s3[0] += Ky * sv; s3[1] += Kx * sv; s3[2] += -k2 / (sqrtf(a.k0 * a.k0 - k2) + a.k0) * sv;
                        }
                    }
                }
            } else {
                // st_i == 0: v = N^2 T of gpsi_0 (shifted probes): probe-spectrum and shift gradients
                const float4* __restrict__ phf = reinterpret_cast<const float4*>(a.PhatF) + (size_t)p * (TILE / 2) + g.t;   // Phat / N^2
                float4* __restrict__ gp = reinterpret_cast<float4*>(a.gPhatF) + (size_t)p * (TILE / 2) + g.t;
                const float2 wyv = s.wy[g.ky];
                const float kapy = float((g.ky + HN) & (FN - 1)) * (1.0f / float(FN));
                const float invN2 = 1.0f / float(TILE);
                float r2[2] = {0.f, 0.f};
#pragma unroll
                for (int j = 0; j < R / 2; ++j) {
                    const float4 pq = __ldg(phf + j * FT);
                    float2 cw[2];
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int u = 2 * j + e;
                        const float2 w = cmul(wyv, s.wx[g.kx(u)]);
                        cw[e] = cmulc(v[u], w);                              // conj(w') * N^2 T
                        const float2 pv = e ? hi2(pq) : lo2(pq);
                        const float qv = cw[e].y * pv.x - cw[e].x * pv.y;    // Im(conj(w') T conj(Phat))
                        r2[0] += kapy * qv;
                        r2[1] += float((g.kx(u) + HN) & (FN - 1)) * (1.0f / float(FN)) * qv;
                    }
                    if (a.need_probe) red_f4(gp + j * FT, cscale(cw[0], invN2), cscale(cw[1], invN2));
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
            for (int k = 0; k < R; ++k) red_f2(gp + k * FT, v[k]);
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

// ---- host side: same scratch layout and call contract as fused128 ------------------------------------------------------------
using fused128::Scratch;
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
inline bool covers(const ptyb200_cfg& c) { return c.N == FN; }
inline size_t scratch_bytes(const ptyb200_cfg& c, int B) { return covers(c) ? carve_scratch(c, B, nullptr).total : 0; }

#define F64_CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { err = std::string("fused64: ") + #call + ": " + cudaGetErrorString(e_); return 1; } } while (0)

inline Args make_args(const ptyb200_cfg& c, const FwdArgs& f, const Scratch& sc, float2* phis) {
    Args a;
    memset(&a, 0, sizeof a);
    a.f = f; a.HF = sc.HF; a.PhatF = sc.PhatF; a.farF = sc.farF; a.phisF = f.phis ? phis : nullptr;
    a.gOpack = sc.gOpack;
    a.gPhatF = sc.gPhatF; a.shift = c.shift_probes;
    return a;
}

// sH / sP: the branches on which the transposed propagator / the probe spectrum are being made (api.cu: setup_common); their permuted
// copies are made there too, and `join` brings both back into `st` just before the wave kernel
template <class Join>
inline int forward(const ptyb200_cfg& c, int B, FwdArgs f, const float* obja, const float* objp, unsigned char* scratch, cudaStream_t st,
                   std::string& err, std::atomic<long long>* launches, cudaStream_t sH, cudaStream_t sP, Join join) {
    Scratch sc = carve_scratch(c, B, scratch);
    Args a = fused64::make_args(c, f, sc, f.phis);
    const float inv = 1.0f / float(TILE);
    k_permute_to_F<<<dim3(TILE / 256, 1), 256, 0, sH>>>(f.HT, sc.HF, inv);
    F64_CK(cudaGetLastError()); ++*launches;
    if (c.shift_probes) {
        k_permute_to_F<<<dim3(TILE / 256, c.P), 256, 0, sP>>>(f.PhatT, sc.PhatF, inv);
        F64_CK(cudaGetLastError()); ++*launches;
    }
    a.f.lf.counter = sc.counter;
    {
        const size_t n4 = (size_t)B * TILE / 4;
        fused128::k_dp_init<<<(unsigned)((n4 + 255) / 256), 256, 0, st>>>(reinterpret_cast<float4*>(f.dp), n4, c.eps, sc.counter, B);
        F64_CK(cudaGetLastError()); ++*launches;
    }
    if (int r = join()) return r;
    const dim3 grid(c.P, c.M, B);
    const bool tilt = f.tvec != nullptr, phis = a.phisF != nullptr;
#define F64_LAUNCH_FWD(T, PH)                                                                                                      \
    do {                                                                                                                           \
        F64_CK(cudaFuncSetAttribute(fused64::k_forward<T, PH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES_FWD));          \
        fused64::k_forward<T, PH><<<grid, FT, SMEM_BYTES_FWD, st>>>(a);                                                                     \
    } while (0)
    if (tilt) { if (phis) F64_LAUNCH_FWD(true, true); else F64_LAUNCH_FWD(true, false); }
    else      { if (phis) F64_LAUNCH_FWD(false, true); else F64_LAUNCH_FWD(false, false); }
#undef F64_LAUNCH_FWD
    F64_CK(cudaGetLastError()); ++*launches;
    return 0;
}

inline int backward(const ptyb200_cfg& c, int B, const BwdArgs& bw, const float* obja, const float* objp, float* g_obja, float* g_objp,
                    unsigned char* scratch, float2* g_probe, float2* gPhatT, cudaStream_t st, std::string& err, std::atomic<long long>* launches,
                    int acc_flags = 0, const float* scale = nullptr, bool finish_only = false,
                    const std::function<cudaStream_t()>& fork_fin = nullptr) {
    // acc_flags (PTYB200_ACC_*): KEEP_GRADS = the accumulators already hold earlier chunks of the batch; NO_FINISH = leave them raw.
    // finish_only: no adjoint, only the completion of the accumulators (with the batch-level `scale` of an unscaled loss gradient).
    Scratch sc = carve_scratch(c, B, scratch);
    Args a = fused64::make_args(c, bw.f, sc, bw.f.phis);
    a.G = bw.G; a.gprop = bw.gprop; a.gshift = bw.gshift; a.gprobe = g_probe;
    a.dx = bw.dx; a.k0 = bw.k0;
    a.need_obj = bw.need_obj; a.need_probe = bw.need_probe; a.need_shift = bw.need_shift; a.need_prop = bw.need_prop;
    a.units = B * c.M * c.P;
    const size_t obj = (size_t)((c.reserved[1] & 1) ? B : 1) * c.M * c.Z * c.Noy * c.Nox;
    if (!finish_only && !(acc_flags & PTYB200_ACC_KEEP_GRADS)) {
        if (a.need_obj) F64_CK(cudaMemsetAsync(sc.gOpack, 0, obj * 16, st));
        if (a.need_probe) {
            if (c.shift_probes) F64_CK(cudaMemsetAsync(sc.gPhatF, 0, (size_t)c.P * TILE * 8, st));
            else F64_CK(cudaMemsetAsync(g_probe, 0, (size_t)c.P * TILE * 8, st));
        }
    }
    const bool tilt = bw.f.tvec != nullptr, prop = a.need_prop != 0;
#define F64_LAUNCH_BWD(T, PR)                                                                                                      \
    do {                                                                                                                           \
        F64_CK(cudaFuncSetAttribute(fused64::k_backward<T, PR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES_BWD));         \
        fused64::k_backward<T, PR><<<a.units, FT, SMEM_BYTES_BWD, st>>>(a);                                                                 \
    } while (0)
    if (!finish_only) {
        if (tilt) { if (prop) F64_LAUNCH_BWD(true, true); else F64_LAUNCH_BWD(true, false); }
        else      { if (prop) F64_LAUNCH_BWD(false, true); else F64_LAUNCH_BWD(false, false); }
        F64_CK(cudaGetLastError()); ++*launches;
    }
#undef F64_LAUNCH_BWD
    const cudaStream_t sp = fork_fin ? fork_fin() : st;        // branch of the probe-gradient chain (api.cu continues it and joins)
    if (acc_flags & PTYB200_ACC_NO_FINISH) return 0;
    if (a.need_probe && c.shift_probes) {
        k_unpermute_from_F<<<dim3(TILE / 256, c.P), 256, 0, sp>>>(sc.gPhatF, gPhatT);
        F64_CK(cudaGetLastError()); ++*launches;
    }
    if (a.need_obj) {
        fused128::k_obj_finish_pack<<<(unsigned)((obj + 255) / 256), 256, 0, st>>>(sc.gOpack, obja, objp, g_obja, g_objp, c.Noy, c.Nox, obj, scale, (acc_flags & PTYB200_ACC_ADD_OBJ) ? 1 : 0);
        F64_CK(cudaGetLastError()); ++*launches;
    }
    return 0;
}

}  // namespace fused64
}  // namespace ptyb
