// General multislice path: every 2-D FFT is two row passes over ROWS-row slabs (8 rows) held in shared memory, each
// pass writing its result transposed, with the pointwise physics fused into the passes:
//
//   forward, per slice z      DA: (inverse-x) -> psi_z -> stash, *O_z (ROI gather), forward-x  -> G1 (transposed)
//                             BC: forward-y, *H_n, inverse-y                                    -> G2
//   far field                 FINAL: forward-y, |.|^2 * occu summed over probe/object modes     -> dp (fftshifted)
//   adjoint, per slice z      START: forward-y, *2 occu G, inverse-y                            -> G2
//                             DA^H: inverse-x -> gphi_z; gO += conj(psi_z) gphi (summed over probe modes in
//                                   registers, one red.global per pixel), gpsi = conj(O_z) gphi, forward-x -> G1
//                             BC^H: forward-y, *conj(H_n) [+ tilt/thickness sums], inverse-y     -> G2
//
// A CTA owns (sample, object mode, probe-mode group, ROWS-row slab) and loops over the probe modes of its group, so the O_z ROI,
// the propagator values and the object-gradient accumulators live in registers across the loop.
// The host can run the slice sequence on CHUNKS of the batch (api.cu: gen_plan; default one chunk): the pass buffers G1/G2 hold only
// the chunk's tiles.
// Works for N = N1*N2 with N1,N2 <= 16 (see dispatch in api.cu).  The N = 128 on-chip kernels are in fused128.cuh.
#pragma once
#include "rowfft.cuh"
#include <stdint.h>

namespace ptyb {

#ifndef PTYB_ROWS
#define PTYB_ROWS 8      // 8-row slabs, 128-thread CTAs, 4 (6 for N <= 192) CTAs per SM: finer barriers; measured +3-4 % over 16 rows x 256 threads
#endif
constexpr int ROWS = PTYB_ROWS;        // slab height (rows per CTA)
constexpr int NT = 16 * PTYB_ROWS;     // threads per CTA (one register-stage work item per thread at N = 256)
#ifndef GEN_MINB
#define GEN_MINB (2 * 16 / PTYB_ROWS)   // minimum resident CTAs per SM requested from the compiler (register cap 128 per thread)
#endif
// the kernels with one cached table (k_fwd_da, k_fwd_bc, k_bwd_bc) run at 80 registers (768 threads per SM instead of 512) without
// spilling up to N = 192 (measured: C5 +14 %); at N = 256 the 80-register build spills and is slower (C4 -3 %, C3 -13 %)
#ifndef GEN_MINB_LIGHT
#define GEN_MINB_LIGHT(F) (((F::N) <= 192 ? 3 : 2) * 16 / PTYB_ROWS)
#endif

struct Dims {
    int N, P, M, Z, Noy, Nox, B;
    int patch;      // 1: the "object" arrays are per-sample patches (B,M,Z,N,N) (pre-blurred ROIs); crop offsets are zero
    int b0;         // first sample of the chunk this launch works on (grid z = sample inside the chunk)
    int pg;         // probe modes per CTA (grid y = object mode + M * probe-mode group)
};
// chunk-local decomposition of a (nb, M*groups, chunk) grid
struct Unit { int m, b, bl, p_lo, p_hi; };
__device__ __forceinline__ Unit unit_of(const Dims& d) {
    Unit u;
    u.m = blockIdx.y % d.M;
    u.p_lo = (blockIdx.y / d.M) * d.pg;
    u.p_hi = min(d.P, u.p_lo + d.pg);
    u.bl = blockIdx.z;
    u.b = d.b0 + u.bl;
    return u;
}
// accesses of the write-once / read-once streams (stash, far-field tiles, Fourier stash).  Streaming (evict-first, .cs) hints were
// measured on B200 and rejected: no gain on the stores, and ld.global.cs on the stash / Fourier-stash reads made the C3 adjoint 27 %
// slower (it defeats the L2 prefetch issued one probe mode ahead).  -DPTYB_STREAM_HINTS rebuilds that variant.
#ifdef PTYB_STREAM_HINTS
__device__ __forceinline__ void st_stream(float2* p, float2 v) { __stcs(p, v); }
__device__ __forceinline__ float2 ld_stream(const float2* p) { return __ldcs(p); }
#else
__device__ __forceinline__ void st_stream(float2* p, float2 v) { *p = v; }
__device__ __forceinline__ float2 ld_stream(const float2* p) { return *p; }
#endif
// object plane index and ROI offset of sample b, object mode m
__device__ __forceinline__ int obj_mode(const Dims& d, int b, int m) { return d.patch ? b * d.M + m : m; }
__device__ __forceinline__ void roi_origin(const Dims& d, const int32_t* crop, const int64_t* idx, int b, int& cy, int& cx) {
    if (d.patch) { cy = 0; cx = 0; return; }
    const int64_t n0 = idx[b];
    cy = crop[2 * n0]; cx = crop[2 * n0 + 1];
}

__device__ __forceinline__ int shift_idx(int k, int N) {  // (k + N/2) mod N  (fftshift == ifftshift for even N)
    int h = N >> 1;
    k += h;
    return k >= N ? k - N : k;
}

// K[n]: ifftshift'ed, half-bin shifted angular frequency grid (models.py:164-171)
__device__ __forceinline__ float kgrid(int n, int N, float dx) {
    int j = shift_idx(n, N);
    return 6.283185307179586f * ((float(j - (N >> 1)) + 0.5f) / float(N)) / dx;
}

__device__ __forceinline__ void red_add_f2(float2* addr, float2 v) {
#if __CUDA_ARCH__ >= 900
    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(addr), "f"(v.x), "f"(v.y) : "memory");
#else
    atomicAdd(&addr->x, v.x);
    atomicAdd(&addr->y, v.y);
#endif
}

// L2 prefetch of a contiguous range (TMA bulk prefetch; address and size multiples of 16 bytes)
__device__ __forceinline__ void l2_prefetch_range(const void* p, unsigned bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes));
}
// every slab this path reads is 16 contiguous rows of a tile: warm the NEXT probe mode's slab while this one is processed
template <int N> __device__ __forceinline__ void prefetch_slab(const float2* tile, int row0) {
    if (threadIdx.x < ROWS) l2_prefetch_range(tile + (size_t)(row0 + threadIdx.x) * N, N * 8);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// block-wide sum of NV floats; result valid in thread 0. red = smem scratch of NV*32 floats.
template <int NV> __device__ __forceinline__ void block_sum(float (&v)[NV], float* red) {
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = warp_sum(v[i]);
    __syncthreads();
    if (lane == 0)
#pragma unroll
        for (int i = 0; i < NV; ++i) red[i * 32 + w] = v[i];
    __syncthreads();
    if (w == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            float t = lane < nw ? red[i * 32 + lane] : 0.f;
            v[i] = warp_sum(t);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// small setup kernels
// ------------------------------------------------------------------------------------------------

// O = a * exp(i*phi) for the whole object (torch.polar, forward.py:53), evaluated once per step
__global__ void k_obj_polar(const float* __restrict__ a, const float* __restrict__ ph, float2* __restrict__ O, size_t n) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    float s, c;
    sincosf(ph[i], &s, &c);
    float av = a[i];
    O[i] = make_float2(av * c, av * s);
}

// per-sample shift ramps, separable: w'[ky,kx] = wy[ky]*wx[kx], w.[n] = exp(-2 pi i s kappa_n),
// kappa_n = ((n + N/2) mod N)/N   (image_proc.py:531-532 with the non-centred grid of models.py:179)
__global__ void k_shift_vectors(const float* __restrict__ shifts, const int64_t* __restrict__ idx, int B, int N,
                                float2* __restrict__ wvec) {
    int b = blockIdx.x;
    int64_t n0 = idx[b];
    float sy = shifts[2 * n0], sx = shifts[2 * n0 + 1];
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
        float kap = float(shift_idx(n, N)) / float(N);
        float s, c;
        sincospif(-2.0f * sy * kap, &s, &c);
        wvec[((size_t)b * 2 + 0) * N + n] = make_float2(c, s);
        sincospif(-2.0f * sx * kap, &s, &c);
        wvec[((size_t)b * 2 + 1) * N + n] = make_float2(c, s);
    }
}

// per-sample tilt ramps, separable: exp(i dz (Ky tan ty + Kx tan tx)) = ey[ky]*ex[kx]  (models.py:336-347)
__global__ void k_tilt_vectors(const float* __restrict__ tilts, int tilt_mode, const int64_t* __restrict__ idx, int B, int N,
                               float dx, const float* __restrict__ dz, float2* __restrict__ tvec) {
    int b = blockIdx.x;
    int64_t n0 = tilt_mode == 2 ? idx[b] : 0;
    float ty = tanf(tilts[2 * n0] / 1e3f), tx = tanf(tilts[2 * n0 + 1] / 1e3f);
    float d = dz[0];
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
        float K = kgrid(n, N, dx);
        float s, c;
        sincosf(d * K * ty, &s, &c);
        tvec[((size_t)b * 2 + 0) * N + n] = make_float2(c, s);
        sincosf(d * K * tx, &s, &c);
        tvec[((size_t)b * 2 + 1) * N + n] = make_float2(c, s);
    }
}

__global__ void k_transpose(const float2* __restrict__ in, float2* __restrict__ out, int N) {
    __shared__ float2 t[32][33];
    int x = blockIdx.x * 32 + threadIdx.x, y0 = blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y)
        if (x < N && y0 + j < N) t[j][threadIdx.x] = in[(size_t)(y0 + j) * N + x];
    __syncthreads();
    int ox = blockIdx.y * 32 + threadIdx.x, oy0 = blockIdx.x * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y)
        if (ox < N && oy0 + j < N) out[(size_t)(oy0 + j) * N + ox] = t[threadIdx.x][j];
}

// exp(i*dz*Kz) in float64 (models.py:222-223,341,355)
__global__ void k_propagator(int N, float dx, float lambd, const float* __restrict__ dz, float2* __restrict__ H) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N * N) return;
    int ky = i / N, kx = i % N;
    const double twopi = 6.283185307179586476925286766559;
    double Ky = twopi * ((double(shift_idx(ky, N) - (N >> 1)) + 0.5) / double(N)) / double(dx);
    double Kx = twopi * ((double(shift_idx(kx, N) - (N >> 1)) + 0.5) / double(N)) / double(dx);
    double k0 = twopi / double(lambd);
    double ph = double(dz[0]) * sqrt(k0 * k0 - Kx * Kx - Ky * Ky);
    double s, c;
    sincos(ph, &s, &c);
    H[i] = make_float2(float(c), float(s));
}

// ROI gather, bit-exact copy (models.py:251-265)
__global__ void k_gather_patches(Dims d, const int64_t* __restrict__ idx, const float* __restrict__ obja,
                                 const float* __restrict__ objp, const int32_t* __restrict__ crop, float* __restrict__ out) {
    int b = blockIdx.z, mz = blockIdx.y;
    int64_t n0 = idx[b];
    int32_t cy = crop[2 * n0], cx = crop[2 * n0 + 1];
    size_t plane = (size_t)mz * d.Noy * d.Nox;
    float2* o = reinterpret_cast<float2*>(out) + ((size_t)b * d.M * d.Z + mz) * d.N * d.N;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < d.N * d.N; e += gridDim.x * blockDim.x) {
        int y = e / d.N, x = e % d.N;
        size_t src = plane + (size_t)(cy + y) * d.Nox + (cx + x);
        o[e] = make_float2(obja[src], objp[src]);
    }
}

// ------------------------------------------------------------------------------------------------
// slab helpers
// ------------------------------------------------------------------------------------------------
template <class F> struct Slab {
    static constexpr int N = F::N;
    static constexpr int EPT = ROWS * N / NT;  // elements per thread in the pointwise phases
    static_assert(ROWS * N % NT == 0, "slab size must be a multiple of the block size");
    static constexpr int SLAB_ELEMS = ROWS * F::RS;
    // natural ownership: lanes run along the row (coalesced for row-major global arrays)
    template <class Fn> __device__ __forceinline__ static void nat(Fn f) {
#pragma unroll
        for (int i = 0; i < EPT; ++i) {
            int e = threadIdx.x + i * NT;
            f(i, e / N, e % N);
        }
    }
    // transposed ownership: lanes run across the ROWS rows (coalesced for arrays indexed [q][row])
    template <class Fn> __device__ __forceinline__ static void tr(Fn f) {
#pragma unroll
        for (int i = 0; i < EPT; ++i) {
            int e = threadIdx.x + i * NT;
            f(i, e % ROWS, e / ROWS);
        }
    }
    static constexpr size_t smem_bytes() { return sizeof(float2) * (SLAB_ELEMS + N) + sizeof(float) * (ROWS * (N + 1) + 8 * 32); }
};

#define PTYB_SMEM_CARVE(F)                                                      \
    extern __shared__ __align__(16) unsigned char smem_raw[];                   \
    float2* slab = reinterpret_cast<float2*>(smem_raw);                          \
    float2* twN = slab + Slab<F>::SLAB_ELEMS;                                    \
    float* fbuf = reinterpret_cast<float*>(twN + F::N);                          \
    float* red = fbuf + ROWS * (F::N + 1);                                       \
    (void)fbuf; (void)red;                                                       \
    F::fill_twiddles(twN);                                                       \
    __syncthreads();

// ------------------------------------------------------------------------------------------------
// plain pass: FFT along the rows of `count` tiles, output transposed or not.  grid (N/ROWS, count)
//   DIR=-1: natural in, frequency-ordered out;  DIR=+1: frequency-ordered (natural index) in, natural out
// ------------------------------------------------------------------------------------------------
template <class F, int DIR, bool TOUT> __global__ void __launch_bounds__(NT, GEN_MINB) k_pass(const float2* __restrict__ in, float2* __restrict__ out, float scale) {
    PTYB_SMEM_CARVE(F)
    constexpr int N = F::N;
    const int r0 = blockIdx.x * ROWS;
    const float2* src = in + (size_t)blockIdx.y * N * N;
    float2* dst = out + (size_t)blockIdx.y * N * N;
    Slab<F>::nat([&](int, int r, int n) {
        float2 v = src[(size_t)(r0 + r) * N + n];
        slab[r * F::RS + (DIR < 0 ? F::addr(n) : F::apos(n))] = v;
    });
    __syncthreads();
    if (DIR < 0) F::forward(slab, ROWS, twN); else F::inverse(slab, ROWS, twN);
    if (TOUT) {
        Slab<F>::tr([&](int, int r, int q) {
            float2 v = slab[r * F::RS + (DIR < 0 ? F::apos(q) : F::addr(q))];
            dst[(size_t)q * N + r0 + r] = cscale(v, scale);
        });
    } else {
        Slab<F>::nat([&](int, int r, int q) {
            float2 v = slab[r * F::RS + (DIR < 0 ? F::apos(q) : F::addr(q))];
            dst[(size_t)(r0 + r) * N + q] = cscale(v, scale);
        });
    }
}

// ------------------------------------------------------------------------------------------------
// loss terms and the measurement view (losses.py:36-104, models.py:384-416): used by the loss kernels further down and by the
// kernels that fuse the mode reduction with the loss
// ------------------------------------------------------------------------------------------------
struct LossK {
    int s_on, p_on, b_on;
    float s_w, s_p, p_w, p_p, p_eps, b_w, b_p;
};

// x^p with the exponents the loss terms actually use evaluated exactly-rounded and cheaply (dp_pow = 0.5 and 1 and their
// derivative exponents -0.5 and 0; powf costs ~10x more and is kept for everything else)
__device__ __forceinline__ float powp(float x, float p) {
    if (p == 0.5f) return sqrtf(x);
    if (p == 1.0f) return x;
    if (p == -0.5f) return 1.0f / sqrtf(x);
    if (p == 0.0f) return 1.0f;
    return powf(x, p);
}

// Measured pattern of sample b as the loss sees it (models.py:384-416): row idx[b] of the stored (Ntot,Hs,Ws) array, optionally
// pasted into a padded background canvas (Hp,Wp) at [h1:h2, w1:w2] ("on-the-fly" padding, models.py:401-405) and optionally
// resampled bilinearly by fixed scale factors and divided by their product (models.py:407-409; the arithmetic follows ATen's
// upsample_bilinear2d with align_corners = false and a given scale_factor: src = (dst + 0.5) / scale - 0.5, clamped at 0).
// Evaluated on the fly inside the loss kernels: no gathered / padded / resampled copy is ever materialised.
struct MeasView {
    const float* meas;      // (Ntot,Hs,Ws)
    const float* padded;    // (Hp,Wp) or null
    int Hs, Ws, Hp, Wp, h1, w1, h2, w2;
    int resample;           // 0: output pixel = source pixel
    int vec;                // 1: plain layout and 16-byte aligned rows -> 128-bit reads
    float ry, rx;           // source pixels per output pixel (1 / scale_factor)
    float scale;            // 1 / prod(scale_factor)
};
__device__ __forceinline__ float meas_src(const MeasView& v, const float* __restrict__ Mrow, int y, int x) {
    if (!v.padded) return Mrow[(size_t)y * v.Ws + x];
    if (y >= v.h1 && y < v.h2 && x >= v.w1 && x < v.w2) return Mrow[(size_t)(y - v.h1) * v.Ws + (x - v.w1)];
    return v.padded[(size_t)y * v.Wp + x];
}
__device__ __forceinline__ float meas_at(const MeasView& v, const float* __restrict__ Mrow, int Y, int X) {
    if (!v.resample) return meas_src(v, Mrow, Y, X);
    const int H = v.padded ? v.Hp : v.Hs, W = v.padded ? v.Wp : v.Ws;
    const float sy = fmaxf(v.ry * (float(Y) + 0.5f) - 0.5f, 0.f), sx = fmaxf(v.rx * (float(X) + 0.5f) - 0.5f, 0.f);
    const int y0 = min(int(sy), H - 1), x0 = min(int(sx), W - 1);
    const int yp = y0 < H - 1 ? 1 : 0, xp = x0 < W - 1 ? 1 : 0;
    const float ly = sy - float(y0), lx = sx - float(x0), hy = 1.f - ly, hx = 1.f - lx;
    const float val = hy * (hx * meas_src(v, Mrow, y0, x0) + lx * meas_src(v, Mrow, y0, x0 + xp)) +
                      ly * (hx * meas_src(v, Mrow, y0 + yp, x0) + lx * meas_src(v, Mrow, y0 + yp, x0 + xp));
    return val * v.scale;
}
__device__ __forceinline__ bool meas_plain(const MeasView& v) { return v.vec != 0; }

// fused mode reduction + loss (north star item 3): the kernel that completes the intensities of a pattern also forms that pattern's
// contribution to the batch sums of the data losses, so dp is not re-read by a separate reduction launch
struct LossFuse {
    int on;
    LossK k;
    MeasView mv;
    double* stats;          // 8 doubles, zeroed (see k_loss_partial)
    float* pac;             // 2*N*N floats, zeroed (PACBED sums) or null
    int* counter;           // fused path: one arrival counter per pattern, zeroed
    const int64_t* rows;    // (B) row of `mv.meas` holding the pattern of sample b
};
// contribution of one pixel to the five running sums (a[0..4]) of k_loss_partial
__device__ __forceinline__ void loss_pixel(const LossK& k, float I, float Mv, int pix, int NN, float* pac, float (&a)[5]) {
    if (k.s_on) { float mp = powp(Mv, k.s_p), df = powp(I, k.s_p) - mp; a[0] += df * df; a[1] += mp; }
    if (k.p_on) { float mq = powp(Mv, k.p_p), iq = powp(I, k.p_p); a[2] += mq * logf(iq + k.p_eps) - iq; a[3] += mq; }
    if (k.b_on) { a[4] += powp(Mv, k.b_p); atomicAdd(pac + pix, I); atomicAdd(pac + NN + pix, Mv); }
}
__device__ __forceinline__ void loss_stats_commit(const LossK& k, const float (&v)[5], double* stats) {
    if (k.s_on) { atomicAdd(stats + 0, (double)v[0]); atomicAdd(stats + 1, (double)v[1]); }
    if (k.p_on) { atomicAdd(stats + 2, (double)v[2]); atomicAdd(stats + 3, (double)v[3]); }
    if (k.b_on) atomicAdd(stats + 4, (double)v[4]);
}

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
struct FwdArgs {
    Dims d;
    const int64_t* idx;
    const int32_t* crop;
    const float2* O;        // (M,Z,Noy,Nox)
    const float2* probe;    // (P,N,N)
    const float2* PhatT;    // (P,N,N) [kx][ky]
    const float2* HT;       // (N,N)   [kx][ky]
    const float2* wvec;     // (B,2,N)  shift ramps or null
    const float2* tvec;     // (B,2,N)  tilt ramps or null
    const float* occu;      // (M)
    float2* stash;          // (B,P,M,Z,N,N)
    float2* phis;           // (B,P,M,Z-1,N,N) [kx][ky] or null
    float2* G1;             // (B,P,M,N,N) [kx][y]
    float2* G2;             // (B,P,M,N,N) [y][kx]
    float2* farT;           // (B,P,M,N,N) [kx][y]
    float* dp;              // (B,N,N)
    float eps;
    LossFuse lf;            // lf.on: also accumulate the data-loss sums while the intensities are complete in registers
};

// psi0 half-shifted: G2[b,p,0][y][kx] = (1/N) * inverse-y( PhatT[p][kx][ky] * wy[ky] * wx[kx] ).  grid (N/ROWS, groups, chunk)
template <class F> __global__ void __launch_bounds__(NT, GEN_MINB) k_init_shift(FwdArgs a) {
    PTYB_SMEM_CARVE(F)
    constexpr int N = F::N, N1 = F::N1, N2 = F::N2;
    const int kx0 = blockIdx.x * ROWS, bl = blockIdx.z, b = a.d.b0 + bl;
    const int p_lo = blockIdx.y * a.d.pg, p_hi = min(a.d.P, p_lo + a.d.pg);
    // two register stages: A item (row r, k1): Phat[kx][k1 + N1 k2] * w' straight from global, inverse DFT over k2 -> smem;
    //                      C item (row rr, j): inverse DFT over k1 -> y = j + N2 k, stored straight to the transposed tile
    const bool itemA = threadIdx.x < ROWS * N1, itemC = threadIdx.x < ROWS * N2;
    const int rA = threadIdx.x / N1, k1A = threadIdx.x % N1;
    const int rC = threadIdx.x % ROWS, jC = threadIdx.x / ROWS;
    float2 wreg[N2];
    if (itemA) {
        const float2* wy = a.wvec + ((size_t)b * 2 + 0) * N;
        const float2 wxv = a.wvec[((size_t)b * 2 + 1) * N + kx0 + rA];
#pragma unroll
        for (int k2 = 0; k2 < N2; ++k2) wreg[k2] = cmul(wy[k1A + N1 * k2], wxv);
    }
    for (int p = p_lo; p < p_hi; ++p) {
        if (itemA) {
            const float2* sp = a.PhatT + (size_t)p * N * N + (size_t)(kx0 + rA) * N + k1A;
            float2 v[N2];
#pragma unroll
            for (int k2 = 0; k2 < N2; ++k2) v[k2] = cmul(sp[N1 * k2], wreg[k2]);
            F::inv_stage2_regs(slab + rA * F::RS, k1A, v, twN);
        }
        __syncthreads();
        if (itemC) {
            float2 v[N1];
            F::inv_stage1_regs(slab + rC * F::RS, jC, v);
            float2* dst = a.G2 + ((size_t)bl * a.d.P + p) * a.d.M * N * N + (size_t)jC * N + kx0 + rC;   // m = 0 slot
#pragma unroll
            for (int k = 0; k < N1; ++k) dst[(size_t)N2 * k * N] = cscale(v[k], 1.0f / N);
        }
        __syncthreads();
    }
}

// grid (N/ROWS, M*groups, chunk).  src_mode: 0 = psi_z comes from G2 (inverse-x), 1 = psi_0 is the unshifted probe.
// Three register stages per probe mode, two shared-memory exchanges between them (4 smem passes instead of 8):
//   A  item (row r, k1):  X[k1 + N1 k2] straight from global (128-byte segments), inverse DFT over k2, twiddle        -> smem
//   B  item (row r, j):   inverse DFT over k1 -> psi_z[x = j + N2 k] -> stash, *O_z, forward DFT over k, twiddle     -> smem
//   C  item (row rr, k1): forward DFT over j -> frequency q = k1 + N1 k2, stored straight to the transposed tile: the lanes run
//      over the ROWS rows of the slab, so each store instruction writes 8*ROWS-byte segments and the odd row stride keeps the
//      shared-memory reads conflict free.
template <class F> __global__ void __launch_bounds__(NT, GEN_MINB_LIGHT(F)) k_fwd_da(FwdArgs a, int z, int src_mode, int last) {
    PTYB_SMEM_CARVE(F)
    constexpr int N = F::N, N1 = F::N1, N2 = F::N2;
    static_assert(ROWS * N2 <= NT && ROWS * N1 <= NT, "one fused work item per thread");
    const Dims& d = a.d;
    const Unit un = unit_of(d);
    const int y0 = blockIdx.x * ROWS, m = un.m, b = un.b;
    int cy, cx;
    roi_origin(d, a.crop, a.idx, b, cy, cx);
    const float2* Oz = a.O + ((size_t)obj_mode(d, b, m) * d.Z + z) * d.Noy * d.Nox;
    const bool itemA = threadIdx.x < ROWS * N1, itemB = threadIdx.x < ROWS * N2;
    const int rA = threadIdx.x / N1, k1A = threadIdx.x % N1;          // stage A
    const int r = threadIdx.x / N2, j = threadIdx.x % N2;             // stage B
    const int rC = threadIdx.x % ROWS, k1C = threadIdx.x / ROWS;      // stage C (itemA range)
    float2 Oreg[N1];
    if (itemB) {
#pragma unroll
        for (int k = 0; k < N1; ++k) Oreg[k] = Oz[(size_t)(cy + y0 + r) * d.Nox + cx + j + N2 * k];
    }
    const int msrc = (z == 0) ? 0 : m;
    for (int p = un.p_lo; p < un.p_hi; ++p) {
        const size_t tile = ((size_t)b * d.P + p) * d.M;           // batch-wide streams (stash, far-field tiles)
        const size_t ltile = ((size_t)un.bl * d.P + p) * d.M;      // chunk-local pass buffers G1 / G2
        if (src_mode == 0) {
            const float2* src = a.G2 + (ltile + msrc) * N * N;
            if (p + 1 < un.p_hi) prefetch_slab<N>(src + (size_t)d.M * N * N, y0);
            if (itemA) {
                float2 v[N2];
                const float2* sp = src + (size_t)(y0 + rA) * N + k1A;
#pragma unroll
                for (int k2 = 0; k2 < N2; ++k2) v[k2] = sp[N1 * k2];
                F::inv_stage2_regs(slab + rA * F::RS, k1A, v, twN);
            }
            __syncthreads();
        }
        if (itemB) {
            float2* row = slab + r * F::RS;
            float2* st = a.stash + ((tile + m) * d.Z + z) * N * N + (size_t)(y0 + r) * N + j;
            float2 v[N1];
            if (src_mode == 0) {
#pragma unroll
                for (int k1 = 0; k1 < N1; ++k1) v[k1] = row[F::addr(j + N2 * k1)];
                Dft<N1, +1, false>::run(v);
#pragma unroll
                for (int k = 0; k < N1; ++k) v[k] = cscale(v[k], 1.0f / N);
            } else {
                const float2* pr = a.probe + (size_t)p * N * N + (size_t)(y0 + r) * N + j;
#pragma unroll
                for (int k = 0; k < N1; ++k) v[k] = pr[N2 * k];
            }
#pragma unroll
            for (int k = 0; k < N1; ++k) {
                st_stream(st + N2 * k, v[k]);
                v[k] = cmul(v[k], Oreg[k]);
            }
            Dft<N1, -1, false>::run(v);
#pragma unroll
            for (int k1 = 0; k1 < N1; ++k1) row[F::addr(j + N2 * k1)] = k1 ? cmul(v[k1], twN[j * k1]) : v[k1];
        }
        __syncthreads();
        if (itemA) {
            float2 v[N2];
            F::fwd_stage2_regs(slab + rC * F::RS, k1C, v);
            float2* dst = (last ? a.farT + (tile + m) * N * N : a.G1 + (ltile + m) * N * N) + (size_t)k1C * N + y0 + rC;
            if (last) {
#pragma unroll
                for (int k2 = 0; k2 < N2; ++k2) st_stream(dst + (size_t)N1 * k2 * N, v[k2]);
            } else {
#pragma unroll
                for (int k2 = 0; k2 < N2; ++k2) dst[(size_t)N1 * k2 * N] = v[k2];
            }
        }
        __syncthreads();
    }
}

// propagator values of one fused work item (row kx0 + r, k1): ky = k1 + N1*k2, k2 < N2
template <class F> __device__ __forceinline__ void load_prop_item(const FwdArgs& a, int b, int kx, int k1, float2 (&Hreg)[F::N2]) {
    constexpr int N = F::N, N1 = F::N1, N2 = F::N2;
#pragma unroll
    for (int k2 = 0; k2 < N2; ++k2) Hreg[k2] = a.HT[(size_t)kx * N + k1 + N1 * k2];
    if (a.tvec) {
        const float2* ey = a.tvec + ((size_t)b * 2 + 0) * N;
        const float2 exv = a.tvec[((size_t)b * 2 + 1) * N + kx];
#pragma unroll
        for (int k2 = 0; k2 < N2; ++k2) Hreg[k2] = cmul(Hreg[k2], cmul(ey[k1 + N1 * k2], exv));
    }
}

// grid (N/ROWS, M*groups, chunk): G1[kx][y] -> forward-y -> *H -> inverse-y -> G2[y][kx].  Same three-stage structure as k_fwd_da:
//   A  item (row r, j):   y = j + N2 k straight from global, forward DFT over k, twiddle                           -> smem
//   B  item (row r, k1):  forward DFT over j -> ky = k1 + N1 k2, *H_n, inverse DFT over k2, conjugate twiddle      -> smem
//   C  item (row rr, j):  inverse DFT over k1 -> y = j + N2 k, stored straight to the transposed tile (lanes over the slab rows)
template <class F> __global__ void __launch_bounds__(NT, GEN_MINB_LIGHT(F)) k_fwd_bc(FwdArgs a, int z) {
    PTYB_SMEM_CARVE(F)
    constexpr int N = F::N, N1 = F::N1, N2 = F::N2;
    const Dims& d = a.d;
    const Unit un = unit_of(d);
    const int kx0 = blockIdx.x * ROWS, m = un.m, b = un.b;
    const bool itemA = threadIdx.x < ROWS * N2, itemB = threadIdx.x < ROWS * N1;
    const int rA = threadIdx.x / N2, jA = threadIdx.x % N2;
    const int r = threadIdx.x / N1, k1 = threadIdx.x % N1;
    const int rC = threadIdx.x % ROWS, jC = threadIdx.x / ROWS;
    float2 Hreg[N2];
    if (itemB) load_prop_item<F>(a, b, kx0 + r, k1, Hreg);
    for (int p = un.p_lo; p < un.p_hi; ++p) {
        const size_t tile = ((size_t)b * d.P + p) * d.M + m;
        const size_t ltile = ((size_t)un.bl * d.P + p) * d.M + m;
        const float2* src = a.G1 + ltile * N * N;
        if (p + 1 < un.p_hi) prefetch_slab<N>(src + (size_t)d.M * N * N, kx0);
        if (itemA) {
            float2 v[N1];
            const float2* sp = src + (size_t)(kx0 + rA) * N + jA;
#pragma unroll
            for (int k = 0; k < N1; ++k) v[k] = sp[N2 * k];
            F::fwd_stage1_regs(slab + rA * F::RS, jA, v, twN);
        }
        __syncthreads();
        if (itemB) {
            float2* row = slab + r * F::RS;
            float2* ph = a.phis ? a.phis + (tile * (d.Z - 1) + z) * N * N + (size_t)(kx0 + r) * N + k1 : nullptr;
            float2 v[N2];
#pragma unroll
            for (int j = 0; j < N2; ++j) v[j] = row[F::addr(j + N2 * k1)];
            Dft<N2, -1, false>::run(v);
#pragma unroll
            for (int k2 = 0; k2 < N2; ++k2) {
                if (ph) st_stream(ph + N1 * k2, v[k2]);
                v[k2] = cmul(v[k2], Hreg[k2]);
            }
            Dft<N2, +1, false>::run(v);
#pragma unroll
            for (int j = 0; j < N2; ++j) row[F::addr(j + N2 * k1)] = k1 ? cmulc(v[j], twN[j * k1]) : v[j];
        }
        __syncthreads();
        if (itemA) {
            float2 v[N1];
            F::inv_stage1_regs(slab + rC * F::RS, jC, v);
            float2* dst = a.G2 + ltile * N * N + (size_t)jC * N + kx0 + rC;
#pragma unroll
            for (int k = 0; k < N1; ++k) dst[(size_t)N2 * k * N] = cscale(v[k], 1.0f / N);
        }
        __syncthreads();
    }
}

// grid (N/ROWS, B): dp[b][sh(ky)][sh(kx)] = eps + sum_{m,p} occu_m |forward-y(farT)|^2 / N^2
template <class F> __global__ void __launch_bounds__(NT, GEN_MINB) k_fwd_final(FwdArgs a) {
    PTYB_SMEM_CARVE(F)
    constexpr int N = F::N, N1 = F::N1, N2 = F::N2;
    const Dims& d = a.d;
    const int kx0 = blockIdx.x * ROWS, b = blockIdx.y;
    // two register stages per tile: A item (row r, j): y = j + N2 k straight from global, forward DFT over k -> smem;
    // B item (row r, k1): forward DFT over j -> ky = k1 + N1 k2, |.|^2 accumulated in registers over the modes
    const bool itemA = threadIdx.x < ROWS * N2, itemB = threadIdx.x < ROWS * N1;
    const int rA = threadIdx.x / N2, jA = threadIdx.x % N2;
    const int r = threadIdx.x / N1, k1 = threadIdx.x % N1;
    float acc[N2];
#pragma unroll
    for (int k2 = 0; k2 < N2; ++k2) acc[k2] = 0.f;
    for (int m = 0; m < d.M; ++m) {
        const float oc = a.occu[m];
        for (int p = 0; p < d.P; ++p) {
            if (itemA) {
                const float2* sp = a.farT + (((size_t)b * d.P + p) * d.M + m) * N * N + (size_t)(kx0 + rA) * N + jA;
                float2 v[N1];
#pragma unroll
                for (int k = 0; k < N1; ++k) v[k] = sp[N2 * k];
                F::fwd_stage1_regs(slab + rA * F::RS, jA, v, twN);
            }
            __syncthreads();
            if (itemB) {
                float2 v[N2];
                F::fwd_stage2_regs(slab + r * F::RS, k1, v);
#pragma unroll
                for (int k2 = 0; k2 < N2; ++k2) acc[k2] += oc * cabs2(v[k2]);
            }
            __syncthreads();
        }
    }
    const float inv = 1.0f / (float(N) * float(N));
    if (itemB) {
#pragma unroll
        for (int k2 = 0; k2 < N2; ++k2) fbuf[r * (N + 1) + k1 + N1 * k2] = acc[k2] * inv + a.eps;
    }
    __syncthreads();
    float* dp = a.dp + (size_t)b * N * N;
    Slab<F>::tr([&](int, int rr, int ky) { dp[(size_t)shift_idx(ky, N) * N + shift_idx(kx0 + rr, N)] = fbuf[rr * (N + 1) + ky]; });
    if (a.lf.on) {                                   // this CTA holds the finished intensities of its slab: add its loss sums
        const float* M_ = a.lf.mv.meas + (size_t)a.lf.rows[b] * a.lf.mv.Hs * a.lf.mv.Ws;
        float acc5[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
        Slab<F>::tr([&](int, int rr, int ky) {
            const int Y = shift_idx(ky, N), X = shift_idx(kx0 + rr, N);
            loss_pixel(a.lf.k, fbuf[rr * (N + 1) + ky], meas_at(a.lf.mv, M_, Y, X), Y * N + X, N * N, a.lf.pac, acc5);
        });
        block_sum<5>(acc5, red);
        if (threadIdx.x == 0) loss_stats_commit(a.lf.k, acc5, a.lf.stats);
    }
}

// ------------------------------------------------------------------------------------------------
// adjoint
// ------------------------------------------------------------------------------------------------
struct BwdArgs {
    FwdArgs f;
    const float* G;         // (B,N,N) dL/d(dp)
    float2* gO;             // (M,Z,Noy,Nox) scratch, zeroed
    float2* gPhatT;         // (P,N,N) [kx][ky] scratch, zeroed
    float* gprop;           // (B,3) sums  K_y S, K_x S, (Kz-k0) S ; zeroed
    float* gshift;          // (Ntot,2) dense output, zeroed
    float dx, k0;
    int need_obj, need_probe, need_shift, need_prop;
};

// grid (N/ROWS, M*groups, chunk): farT -> forward-y -> * 2 occu G~ -> inverse-y -> G2
template <class F> __global__ void __launch_bounds__(NT, GEN_MINB) k_bwd_start(BwdArgs a) {
    PTYB_SMEM_CARVE(F)
    constexpr int N = F::N, N1 = F::N1, N2 = F::N2;
    const Dims& d = a.f.d;
    const Unit un = unit_of(d);
    const int kx0 = blockIdx.x * ROWS, m = un.m, b = un.b;
    // three register stages like k_fwd_bc, with 2 occu_m dL/dI (fftshifted) in the place of the propagator
    const bool itemA = threadIdx.x < ROWS * N2, itemB = threadIdx.x < ROWS * N1;
    const int rA = threadIdx.x / N2, jA = threadIdx.x % N2;
    const int r = threadIdx.x / N1, k1 = threadIdx.x % N1;
    const int rC = threadIdx.x % ROWS, jC = threadIdx.x / ROWS;
    const float* G = a.G + (size_t)b * N * N;
    Slab<F>::tr([&](int, int rr, int ky) { fbuf[rr * (N + 1) + ky] = G[(size_t)shift_idx(ky, N) * N + shift_idx(kx0 + rr, N)]; });
    __syncthreads();
    const float oc2 = 2.0f * a.f.occu[m];
    float Greg[N2];
#pragma unroll
    for (int k2 = 0; k2 < N2; ++k2) Greg[k2] = itemB ? oc2 * fbuf[r * (N + 1) + k1 + N1 * k2] : 0.f;
    for (int p = un.p_lo; p < un.p_hi; ++p) {
        const size_t tile = ((size_t)b * d.P + p) * d.M + m;
        const size_t ltile = ((size_t)un.bl * d.P + p) * d.M + m;
        if (itemA) {
            const float2* sp = a.f.farT + tile * N * N + (size_t)(kx0 + rA) * N + jA;
            float2 v[N1];
#pragma unroll
            for (int k = 0; k < N1; ++k) v[k] = ld_stream(sp + N2 * k);
            F::fwd_stage1_regs(slab + rA * F::RS, jA, v, twN);
        }
        __syncthreads();
        if (itemB) {
            float2* row = slab + r * F::RS;
            float2 v[N2];
            F::fwd_stage2_regs(row, k1, v);
#pragma unroll
            for (int k2 = 0; k2 < N2; ++k2) v[k2] = cscale(v[k2], Greg[k2]);
            F::inv_stage2_regs(row, k1, v, twN);
        }
        __syncthreads();
        if (itemA) {
            float2 v[N1];
            F::inv_stage1_regs(slab + rC * F::RS, jC, v);
            float2* dst = a.f.G2 + ltile * N * N + (size_t)jC * N + kx0 + rC;
#pragma unroll
            for (int k = 0; k < N1; ++k) dst[(size_t)N2 * k * N] = cscale(v[k], 1.0f / N);
        }
        __syncthreads();
    }
}

// grid (N/ROWS, M*groups, chunk).  out_mode: 0 = forward-x and store transposed to G1 (z>0, or z==0 with shifted probes),
//                                 1 = store gpsi_0 untransformed (natural) to G1 (z==0, unshifted probes), 2 = nothing.
// Three register stages like k_fwd_da: A inverse DFT over k2 from global; B inverse DFT over k1 -> gphi_z, gO accumulation over the
// probe modes, conj(O_z) multiply, forward DFT over k; C forward DFT over j -> transposed store.
template <class F> __global__ void __launch_bounds__(NT, GEN_MINB) k_bwd_da(BwdArgs a, int z, int out_mode) {
    PTYB_SMEM_CARVE(F)
    constexpr int N = F::N, N1 = F::N1, N2 = F::N2;
    const Dims& d = a.f.d;
    const Unit un = unit_of(d);
    const int y0 = blockIdx.x * ROWS, m = un.m, b = un.b;
    int cy, cx;
    roi_origin(d, a.f.crop, a.f.idx, b, cy, cx);
    const float2* Oz = a.f.O + ((size_t)obj_mode(d, b, m) * d.Z + z) * d.Noy * d.Nox;
    const bool itemA = threadIdx.x < ROWS * N1, itemB = threadIdx.x < ROWS * N2;
    const int rA = threadIdx.x / N1, k1A = threadIdx.x % N1;
    const int r = threadIdx.x / N2, j = threadIdx.x % N2;
    const int rC = threadIdx.x % ROWS, k1C = threadIdx.x / ROWS;
    const size_t roi = (size_t)(cy + y0 + r) * d.Nox + cx + j;
    // register budget of stage B (128): accO + the stash values requested ahead of the DFT + the DFT's own array; O_z is re-read
    // per probe mode instead of being cached in registers -- its 32 KB slab stays in L1 because the streaming tiles bypass L1 (ld.cg)
    float2 accO[N1];
#pragma unroll
    for (int k = 0; k < N1; ++k) accO[k] = make_float2(0.f, 0.f);
    for (int p = un.p_lo; p < un.p_hi; ++p) {
        const size_t tile = ((size_t)b * d.P + p) * d.M + m;
        const size_t ltile = ((size_t)un.bl * d.P + p) * d.M + m;
        const float2* src = a.f.G2 + ltile * N * N;
        const float2* stt = a.f.stash + (tile * d.Z + z) * N * N;
        if (p + 1 < un.p_hi) {
            prefetch_slab<N>(src + (size_t)d.M * N * N, y0);
            prefetch_slab<N>(stt + (size_t)d.M * d.Z * N * N, y0);
        }
        if (itemA) {
            float2 v[N2];
            const float2* sp = src + (size_t)(y0 + rA) * N + k1A;
#pragma unroll
            for (int k2 = 0; k2 < N2; ++k2) v[k2] = __ldcg(sp + N1 * k2);
            F::inv_stage2_regs(slab + rA * F::RS, k1A, v, twN);
        }
        float2 psi[N1];
        if (itemB) {                                           // requested ahead of the barrier and the DFT: latency hidden
            const float2* st = stt + (size_t)(y0 + r) * N + j;
#pragma unroll
            for (int k = 0; k < N1; ++k) psi[k] = __ldcg(st + N2 * k);
        }
        __syncthreads();
        float2* dst = a.f.G1 + ltile * N * N;
        if (itemB) {
            float2* row = slab + r * F::RS;
            float2 v[N1];
#pragma unroll
            for (int k1 = 0; k1 < N1; ++k1) v[k1] = row[F::addr(j + N2 * k1)];
            Dft<N1, +1, false>::run(v);
#pragma unroll
            for (int k = 0; k < N1; ++k) {
                const float2 gphi = cscale(v[k], 1.0f / N);
                accO[k] = cadd(accO[k], cmulc(gphi, psi[k]));               // conj(psi) * gphi
                v[k] = cmulc(gphi, __ldg(Oz + roi + N2 * k));               // conj(O) * gphi
            }
            if (out_mode == 0) {
                Dft<N1, -1, false>::run(v);
#pragma unroll
                for (int k1 = 0; k1 < N1; ++k1) row[F::addr(j + N2 * k1)] = k1 ? cmul(v[k1], twN[j * k1]) : v[k1];
            } else if (out_mode == 1) {
#pragma unroll
                for (int k = 0; k < N1; ++k) dst[(size_t)(y0 + r) * N + j + N2 * k] = v[k];
            }
        }
        __syncthreads();
        if (out_mode == 0) {
            if (itemA) {
                float2 v[N2];
                F::fwd_stage2_regs(slab + rC * F::RS, k1C, v);
                float2* dp = dst + (size_t)k1C * N + y0 + rC;
#pragma unroll
                for (int k2 = 0; k2 < N2; ++k2) dp[(size_t)N1 * k2 * N] = v[k2];
            }
            __syncthreads();
        }
    }
    if (a.need_obj && itemB) {
        float2* gOz = a.gO + ((size_t)obj_mode(d, b, m) * d.Z + z) * d.Noy * d.Nox + roi;
#pragma unroll
        for (int k = 0; k < N1; ++k) red_add_f2(gOz + N2 * k, accO[k]);
    }
}

// grid (N/ROWS, M*groups, chunk), z >= 1: G1 -> forward-y -> *conj(H_n) [+ propagator-gradient sums vs Phi_{z-1}] -> inverse-y -> G2
// (three register stages like k_fwd_bc; PROP is compile-time so that the plain adjoint carries no per-element branches)
template <class F, bool PROP> __global__ void __launch_bounds__(NT, GEN_MINB_LIGHT(F)) k_bwd_bc(BwdArgs a, int z) {
    PTYB_SMEM_CARVE(F)
    constexpr int N = F::N, N1 = F::N1, N2 = F::N2;
    const Dims& d = a.f.d;
    const Unit un = unit_of(d);
    const int kx0 = blockIdx.x * ROWS, m = un.m, b = un.b;
    const bool itemA = threadIdx.x < ROWS * N2, itemB = threadIdx.x < ROWS * N1;
    const int rA = threadIdx.x / N2, jA = threadIdx.x % N2;
    const int r = threadIdx.x / N1, k1 = threadIdx.x % N1;
    const int rC = threadIdx.x % ROWS, jC = threadIdx.x / ROWS;
    float2 Hreg[N2];
    if (itemB) load_prop_item<F>(a.f, b, kx0 + r, k1, Hreg);
    float s3[3] = {0.f, 0.f, 0.f};
    const float Kx = kgrid(kx0 + r, N, a.dx);
    const float invN2 = 1.0f / (float(N) * float(N));
    float* kyS = fbuf;                                       // Ky[ky] (propagator gradients only); Kz - k0 is formed on the fly so
    if (PROP) {                                              // that the registers can hold the Fourier stash loaded ahead of the DFT
        for (int n = threadIdx.x; n < N; n += NT) kyS[n] = kgrid(n, N, a.dx);
        __syncthreads();
    }
    for (int p = un.p_lo; p < un.p_hi; ++p) {
        const size_t tile = ((size_t)b * d.P + p) * d.M + m;
        const size_t ltile = ((size_t)un.bl * d.P + p) * d.M + m;
        const float2* src = a.f.G1 + ltile * N * N;
        if (p + 1 < un.p_hi) {
            prefetch_slab<N>(src + (size_t)d.M * N * N, kx0);
            if (PROP) prefetch_slab<N>(a.f.phis + ((tile + d.M) * (d.Z - 1) + (z - 1)) * N * N, kx0);
        }
        if (itemA) {
            float2 v[N1];
            const float2* sp = src + (size_t)(kx0 + rA) * N + jA;
#pragma unroll
            for (int k = 0; k < N1; ++k) v[k] = sp[N2 * k];
            F::fwd_stage1_regs(slab + rA * F::RS, jA, v, twN);
        }
        __syncthreads();
        if (itemB) {
            float2* row = slab + r * F::RS;
            const float2* ph = PROP ? a.f.phis + (tile * (d.Z - 1) + (z - 1)) * N * N + (size_t)(kx0 + r) * N + k1 : nullptr;
            float2 v[N2], phi[N2];
            if (PROP) {                                                      // issued before the DFT: latency hidden behind it
#pragma unroll
                for (int k2 = 0; k2 < N2; ++k2) phi[k2] = ld_stream(ph + N1 * k2);
            }
#pragma unroll
            for (int jj = 0; jj < N2; ++jj) v[jj] = row[F::addr(jj + N2 * k1)];
            Dft<N2, -1, false>::run(v);
#pragma unroll
            for (int k2 = 0; k2 < N2; ++k2) {
                v[k2] = cmulc(v[k2], Hreg[k2]);                              // conj(H) * F2(gpsi)
                if (PROP) {
                    const float sv = (phi[k2].x * v[k2].y - phi[k2].y * v[k2].x) * invN2;   // Im(conj(Phi) * v) / N^2
                    const float Ky = kyS[k1 + N1 * k2];
                    const float kk = Kx * Kx + Ky * Ky;
                    s3[0] += Ky * sv; s3[1] += Kx * sv;
                    s3[2] += -kk / (sqrtf(a.k0 * a.k0 - kk) + a.k0) * sv;    // Kz - k0, cancellation-free
                }
            }
            Dft<N2, +1, false>::run(v);
#pragma unroll
            for (int jj = 0; jj < N2; ++jj) row[F::addr(jj + N2 * k1)] = k1 ? cmulc(v[jj], twN[jj * k1]) : v[jj];
        }
        __syncthreads();
        if (itemA) {
            float2 v[N1];
            F::inv_stage1_regs(slab + rC * F::RS, jC, v);
            float2* dst = a.f.G2 + ltile * N * N + (size_t)jC * N + kx0 + rC;
#pragma unroll
            for (int k = 0; k < N1; ++k) dst[(size_t)N2 * k * N] = cscale(v[k], 1.0f / N);
        }
        __syncthreads();
    }
    if (PROP) {
        block_sum<3>(s3, red);
        if (threadIdx.x == 0) {
            atomicAdd(a.gprop + 3 * b + 0, s3[0]);
            atomicAdd(a.gprop + 3 * b + 1, s3[1]);
            atomicAdd(a.gprop + 3 * b + 2, s3[2]);
        }
    }
}

// shifted probes: grid (N/ROWS, P, nsub), samples bl_lo..bl_hi of the chunk per CTA.  T = forward-y(sum_m G1[bl,p,m]) / N^2 ;
// gPhatT += conj(w') T ; shift gradients -2 pi sum kappa Im(conj(w') conj(Phat) T).
// Two register stages per sample: A item (row r, j): y = j + N2 k straight from global (summed over the object modes), forward DFT
// over k, twiddle -> smem;  B item (row r, k1): forward DFT over j -> ky = k1 + N1 k2, and all the pointwise work on those N2
// frequencies in registers (the accumulators of the probe-spectrum gradient stay in registers across the samples).
template <class F> __global__ void __launch_bounds__(NT, GEN_MINB) k_bwd_probe(BwdArgs a, int nchunk_b, int bsub) {
    PTYB_SMEM_CARVE(F)
    constexpr int N = F::N, N1 = F::N1, N2 = F::N2;
    const Dims& d = a.f.d;
    const int kx0 = blockIdx.x * ROWS, p = blockIdx.y;
    const int b_lo = blockIdx.z * bsub, b_hi = min(nchunk_b, b_lo + bsub);
    const bool itemA = threadIdx.x < ROWS * N2, itemB = threadIdx.x < ROWS * N1;
    const int rA = threadIdx.x / N2, jA = threadIdx.x % N2;
    const int r = threadIdx.x / N1, k1 = threadIdx.x % N1;
    float2 acc[N2], Ph[N2];
#pragma unroll
    for (int k2 = 0; k2 < N2; ++k2) {
        acc[k2] = make_float2(0.f, 0.f);
        Ph[k2] = itemB ? a.f.PhatT[((size_t)p * N + kx0 + r) * N + k1 + N1 * k2] : make_float2(0.f, 0.f);
    }
    const float kapx = float(shift_idx(kx0 + r, N)) / float(N);
    const float invN2 = 1.0f / (float(N) * float(N));
    for (int bl = b_lo; bl < b_hi; ++bl) {
        const int b = d.b0 + bl;
        const size_t tile0 = ((size_t)bl * d.P + p) * d.M;
        if (itemA) {
            float2 v[N1];
            const float2* sp = a.f.G1 + tile0 * N * N + (size_t)(kx0 + rA) * N + jA;
#pragma unroll
            for (int k = 0; k < N1; ++k) v[k] = __ldcg(sp + N2 * k);
            for (int m = 1; m < d.M; ++m) {
#pragma unroll
                for (int k = 0; k < N1; ++k) v[k] = cadd(v[k], __ldcg(sp + (size_t)m * N * N + N2 * k));
            }
            F::fwd_stage1_regs(slab + rA * F::RS, jA, v, twN);
        }
        __syncthreads();
        float s2[2] = {0.f, 0.f};
        if (itemB) {
            const float2* wy = a.f.wvec + ((size_t)b * 2 + 0) * N;
            const float2 wxv = a.f.wvec[((size_t)b * 2 + 1) * N + kx0 + r];
            float2 v[N2];
            F::fwd_stage2_regs(slab + r * F::RS, k1, v);
#pragma unroll
            for (int k2 = 0; k2 < N2; ++k2) {
                const int ky = k1 + N1 * k2;
                const float2 T = cscale(v[k2], invN2);
                const float2 cwT = cmulc(T, cmul(wy[ky], wxv));              // conj(w') * T
                acc[k2] = cadd(acc[k2], cwT);
                const float q = cwT.y * Ph[k2].x - cwT.x * Ph[k2].y;         // Im(conj(w') T conj(Phat))
                s2[0] += float(shift_idx(ky, N)) / float(N) * q;
                s2[1] += kapx * q;
            }
        }
        if (a.need_shift) {
            block_sum<2>(s2, red);                                          // its barriers also order the slab reuse
            if (threadIdx.x == 0) {
                int64_t n0 = a.f.idx[b];
                atomicAdd(a.gshift + 2 * n0 + 0, -6.283185307179586f * s2[0]);
                atomicAdd(a.gshift + 2 * n0 + 1, -6.283185307179586f * s2[1]);
            }
        }
        __syncthreads();
    }
    if (a.need_probe && itemB) {
        float2* dst = a.gPhatT + ((size_t)p * N + kx0 + r) * N + k1;
#pragma unroll
        for (int k2 = 0; k2 < N2; ++k2) red_add_f2(dst + N1 * k2, acc[k2]);
    }
}

// unshifted probes: g_probe[p][y][x] (+)= sum_{bl,m} G1[bl,p,m][y][x] (natural layout) over the nb samples of one chunk; chunks run
// one after the other on the stream, the first one stores.  grid (ceil(N*N/256), P)
__global__ void k_bwd_probe_noshift(Dims d, int nb, int first, const float2* __restrict__ G1, float2* __restrict__ gprobe) {
    int e = blockIdx.x * blockDim.x + threadIdx.x, p = blockIdx.y;
    if (e >= d.N * d.N) return;
    float2* dst = gprobe + (size_t)p * d.N * d.N + e;
    float2 acc = first ? make_float2(0.f, 0.f) : *dst;
    for (int b = 0; b < nb; ++b)
        for (int m = 0; m < d.M; ++m) acc = cadd(acc, G1[(((size_t)b * d.P + p) * d.M + m) * d.N * d.N + e]);
    *dst = acc;
}

// g_a = Re(gO e^{-i phi}), g_phi = Im(gO conj(O))   (polar backward, forward.py:53)
// `scale` (device scalar or null): the batch-level factor of dL/dI when the adjoint ran on the unscaled loss gradient (chunked steps)
__global__ void k_obj_finish(const float2* __restrict__ gO, const float* __restrict__ a, const float* __restrict__ ph,
                             float* __restrict__ ga, float* __restrict__ gp, size_t n, const float* __restrict__ scale, int add) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float sc = scale ? scale[0] : 1.0f;
    float s, c;
    sincosf(ph[i], &s, &c);
    float2 g = gO[i];
    const float va = sc * (g.x * c + g.y * s), vp = sc * a[i] * (g.y * c - g.x * s);
    if (add) { ga[i] += va; gp[i] += vp; }      // the gradient arrays already hold other terms (loss_sparse, written early on a side stream)
    else { ga[i] = va; gp[i] = vp; }
}
__global__ void k_add_into(float4* __restrict__ dst, const float4* __restrict__ src, size_t n4) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < n4) { float4 a = dst[i]; const float4 b = src[i]; a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; dst[i] = a; }
}
__global__ void k_scale(float* __restrict__ x, size_t n, const float* __restrict__ scale) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < n) x[i] *= scale[0];
}

// tilt / thickness chain rule from the per-sample sums (models.py:336-356).  single block.
__global__ void k_prop_finish(const float* __restrict__ gprop, const float* __restrict__ tilts, int tilt_mode,
                              const int64_t* __restrict__ idx, int B, const float* __restrict__ dz, float* g_tilts,
                              float* g_dz) {
    float acc_dz = 0.f, acc_ty = 0.f, acc_tx = 0.f;
    const float d = dz[0];
    for (int b = threadIdx.x; b < B; b += blockDim.x) {
        float Ay = gprop[3 * b], Ax = gprop[3 * b + 1], Az = gprop[3 * b + 2];
        float thy = 0.f, thx = 0.f;
        int64_t n0 = 0;
        if (tilt_mode) {
            n0 = tilt_mode == 2 ? idx[b] : 0;
            thy = tilts[2 * n0] / 1e3f; thx = tilts[2 * n0 + 1] / 1e3f;
        }
        float cy = cosf(thy), cx = cosf(thx);
        float gy = d * Ay / (cy * cy) / 1e3f, gx = d * Ax / (cx * cx) / 1e3f;
        if (g_tilts) {
            if (tilt_mode == 2) { atomicAdd(g_tilts + 2 * n0, gy); atomicAdd(g_tilts + 2 * n0 + 1, gx); }
            else { acc_ty += gy; acc_tx += gx; }
        }
        acc_dz += Az + tanf(thy) * Ay + tanf(thx) * Ax;
    }
    __shared__ float red[3 * 32];
    float v[3] = {acc_dz, acc_ty, acc_tx};
    block_sum<3>(v, red);
    if (threadIdx.x == 0) {
        if (g_dz) g_dz[0] = v[0];
        if (g_tilts && tilt_mode == 1) { g_tilts[0] = v[1]; g_tilts[1] = v[2]; }
    }
}

// ------------------------------------------------------------------------------------------------
// losses (losses.py:36-104)
// ------------------------------------------------------------------------------------------------
// the transformed patterns as a tensor (PtychoAD.get_measurements(indices), models.py:384-416).  grid (chunks, B)
__global__ void k_meas_gather(MeasView mv, const int64_t* __restrict__ idx, int N, float* __restrict__ out) {
    const int NN = N * N, b = blockIdx.y;
    const float* __restrict__ M_ = mv.meas + (size_t)idx[b] * mv.Hs * mv.Ws;
    for (int pix = blockIdx.x * blockDim.x + threadIdx.x; pix < NN; pix += gridDim.x * blockDim.x)
        out[(size_t)b * NN + pix] = meas_at(mv, M_, pix / N, pix % N);
}

// stats: [0] sum (I^p-M^p)^2  [1] sum M^p  [2] sum (M^q log(I^q+e) - I^q)  [3] sum M^q  [4] sum M^r  [5] sum (Ibar^r - Mbar^r)^2
// grid (chunks, B): block (c, b) handles a contiguous chunk of pattern b
__global__ void k_loss_partial(LossK k, const float* __restrict__ dp, MeasView mv, const int64_t* __restrict__ idx,
                               int B, int N, double* stats, float* pac) {
    const int NN = N * N, b = blockIdx.y;
    const float* __restrict__ I_ = dp + (size_t)b * NN;
    const float* __restrict__ M_ = mv.meas + (size_t)idx[b] * mv.Hs * mv.Ws;
    float v[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
    auto term = [&](int pix, float I, float Mv) { loss_pixel(k, I, Mv, pix, NN, pac, v); };
    if (meas_plain(mv)) {                                     // 128-bit reads of both streams (N*N is a multiple of 4)
        const float4* __restrict__ I4 = reinterpret_cast<const float4*>(I_);
        const float4* __restrict__ M4 = reinterpret_cast<const float4*>(M_);
        for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < NN / 4; q += gridDim.x * blockDim.x) {
            const float4 i4 = I4[q], m4 = __ldg(M4 + q);
            term(4 * q, i4.x, m4.x); term(4 * q + 1, i4.y, m4.y); term(4 * q + 2, i4.z, m4.z); term(4 * q + 3, i4.w, m4.w);
        }
    } else {
        for (int pix = blockIdx.x * blockDim.x + threadIdx.x; pix < NN; pix += gridDim.x * blockDim.x)
            term(pix, I_[pix], meas_at(mv, M_, pix / N, pix % N));
    }
    __shared__ float red[5 * 32];
    block_sum<5>(v, red);
    if (threadIdx.x == 0) loss_stats_commit(k, v, stats);
}

__global__ void k_loss_final(LossK k, int B, int N, double* stats, const float* __restrict__ pac, float* losses3) {
    const size_t NN = (size_t)N * N;
    __shared__ double sh[32];
    double acc = 0;
    if (k.b_on) {
        for (size_t i = threadIdx.x; i < NN; i += blockDim.x) {
            float ib = pac[i] / B, mb = pac[NN + i] / B;
            float df = powp(ib, k.b_p) - powp(mb, k.b_p);
            acc += (double)df * df;
        }
        acc = warp_sum_d(acc);
        if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
        __syncthreads();
        if (threadIdx.x == 0) { double t = 0; for (int i = 0; i < (blockDim.x + 31) / 32; ++i) t += sh[i]; stats[5] = t; }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const double nel = (double)B * NN;
        losses3[0] = k.s_on ? float(k.s_w * sqrt(stats[0] / nel) / (stats[1] / nel)) : 0.f;
        losses3[1] = k.p_on ? float(-k.p_w * (stats[2] / nel) / (stats[3] / nel)) : 0.f;
        losses3[2] = k.b_on ? float(k.b_w * sqrt(stats[5] / NN) / (stats[4] / nel)) : 0.f;
    }
}

__global__ void k_loss_grad(LossK k, const float* __restrict__ dp, MeasView mv, const int64_t* __restrict__ idx,
                            int B, int N, const double* __restrict__ stats, const float* __restrict__ pac,
                            const float* __restrict__ up, float* __restrict__ G) {
    const int NN = N * N, b = blockIdx.y;
    const double nel = (double)B * NN;
    float cs = 0.f, cp = 0.f, cb = 0.f;
    if (!up) {
        // unscaled form (chunked steps): the factor that depends on the sums over the WHOLE batch is left out and applied to the
        // finished gradients instead (k_loss_scale); the adjoint is linear in G.  One data term only (checked by the caller).
        if (k.s_on) cs = k.s_w * k.s_p;
        if (k.p_on) cp = -k.p_w * k.p_p;
    } else {
        if (k.s_on) cs = float(up[0] * k.s_w * k.s_p / (nel * sqrt(stats[0] / nel) * (stats[1] / nel)));
        if (k.p_on) cp = float(-up[1] * k.p_w * k.p_p / (nel * (stats[3] / nel)));
        if (k.b_on) cb = float(up[2] * k.b_w * k.b_p / (nel * sqrt(stats[5] / NN) * (stats[4] / nel)));
    }
    const float* __restrict__ I_ = dp + (size_t)b * NN;
    const float* __restrict__ M_ = mv.meas + (size_t)idx[b] * mv.Hs * mv.Ws;
    float* __restrict__ G_ = G + (size_t)b * NN;
    auto grad = [&](int pix, float I, float Mv) {
        float g = 0.f;
        if (k.s_on) g += cs * (powp(I, k.s_p) - powp(Mv, k.s_p)) * powp(I, k.s_p - 1.0f);
        if (k.p_on) g += cp * (powp(Mv, k.p_p) / (powp(I, k.p_p) + k.p_eps) - 1.0f) * powp(I, k.p_p - 1.0f);
        if (k.b_on) {
            const float ib = pac[pix] / B, mb = pac[NN + pix] / B;
            g += cb * (powp(ib, k.b_p) - powp(mb, k.b_p)) * powp(ib, k.b_p - 1.0f);
        }
        return g;
    };
    const bool need_m = k.s_on || k.p_on;
    if (meas_plain(mv)) {
        const float4* __restrict__ I4 = reinterpret_cast<const float4*>(I_);
        const float4* __restrict__ M4 = reinterpret_cast<const float4*>(M_);
        for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < NN / 4; q += gridDim.x * blockDim.x) {
            const float4 i4 = I4[q];
            const float4 m4 = need_m ? __ldg(M4 + q) : make_float4(0.f, 0.f, 0.f, 0.f);
            reinterpret_cast<float4*>(G_)[q] = make_float4(grad(4 * q, i4.x, m4.x), grad(4 * q + 1, i4.y, m4.y), grad(4 * q + 2, i4.z, m4.z),
                                                           grad(4 * q + 3, i4.w, m4.w));
        }
    } else {
        for (int pix = blockIdx.x * blockDim.x + threadIdx.x; pix < NN; pix += gridDim.x * blockDim.x)
            G_[pix] = grad(pix, I_[pix], need_m ? meas_at(mv, M_, pix / N, pix % N) : 0.f);
    }
}

// the batch-level factor that k_loss_grad's unscaled form leaves out: nel = elements of the WHOLE batch
__global__ void k_loss_scale(LossK k, double nel, const double* __restrict__ stats, const float* __restrict__ up, float* __restrict__ scale) {
    if (threadIdx.x || blockIdx.x) return;
    if (k.s_on) scale[0] = float(up[0] / (nel * sqrt(stats[0] / nel) * (stats[1] / nel)));
    else scale[0] = float(up[1] / (nel * (stats[3] / nel)));
}

// sparse: sum over the batch ROIs of |phi|^n == sum over object pixels of cover[px] * |phi[px]|^n, where cover counts
// how many ROIs of the batch contain the pixel (k_cover).  grid (chunks, M*Z)
__global__ void k_sparse_partial(Dims d, float order, const float* __restrict__ objp, const int32_t* __restrict__ cover, double* Ssum) {
    const int mz = blockIdx.y;
    const size_t plane = (size_t)d.Noy * d.Nox;
    const float* pl = objp + (size_t)mz * plane;
    float acc = 0.f;
    for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < plane; e += (size_t)gridDim.x * blockDim.x) {
        const int c = cover[e];
        if (c) {
            const float v = fabsf(pl[e]);
            acc += float(c) * (order == 1.0f ? v : (order == 2.0f ? v * v : powf(v, order)));
        }
    }
    __shared__ float red[32];
    float r1[1] = {acc};
    block_sum<1>(r1, red);
    if (threadIdx.x == 0) atomicAdd(Ssum + mz / d.Z, (double)r1[0]);
}

__global__ void k_sparse_final(Dims d, float weight, float order, const float* __restrict__ occu, const double* __restrict__ Ssum,
                               float* loss) {
    if (threadIdx.x || blockIdx.x) return;
    double cnt = (double)d.B * d.Z * d.N * d.N, t = 0;
    for (int m = 0; m < d.M; ++m) t += occu[m] * pow(Ssum[m] / cnt, 1.0 / order);
    loss[0] = float(weight * t);
}

__global__ void k_cover(Dims d, const int32_t* __restrict__ crop, const int64_t* __restrict__ idx, int32_t* cover) {
    int b = blockIdx.y;
    int64_t n0 = idx[b];
    int cy = crop[2 * n0], cx = crop[2 * n0 + 1];
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < d.N * d.N; e += gridDim.x * blockDim.x)
        atomicAdd(cover + (size_t)(cy + e / d.N) * d.Nox + cx + e % d.N, 1);
}

__global__ void k_sparse_grad(Dims d, float weight, float order, const float* __restrict__ objp, const float* __restrict__ occu,
                              const double* __restrict__ Ssum, const float* __restrict__ up, const int32_t* __restrict__ cover,
                              float* __restrict__ g_objp) {
    size_t plane = (size_t)d.Noy * d.Nox, n = plane * d.M * d.Z;
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    int m = int(i / (plane * d.Z));
    int c = cover[i % plane];
    if (!c) return;
    double cnt = (double)d.B * d.Z * d.N * d.N;
    float S = float(Ssum[m] / cnt);
    float coef = up[0] * weight * occu[m] * powf(S, 1.0f / order - 1.0f) / float(cnt);
    float v = objp[i], av = fabsf(v);
    float sg = v > 0.f ? 1.f : (v < 0.f ? -1.f : 0.f);
    float pw = order == 1.0f ? 1.0f : (order == 2.0f ? av : powf(av, order - 1.0f));
    g_objp[i] += coef * pw * sg * float(c);
}


// ------------------------------------------------------------------------------------------------
// 5-tap Gaussian blur with reflect padding along one axis, and its adjoint (torchvision gaussian_blur(kernel_size=5), used by
// models.py:275-284 (object pre-blur), models.py:379-380 (detector blur) and losses.py:125,134).
//   forward:  out[n] = sum_m W[n][m] in[m],   adjoint:  out[m] = sum_n W[n][m] in[n],   W[n][m] = sum_i k_i [refl(n+i) == m]
// refl(t) = -t (t < 0), 2(S-1)-t (t > S-1): every non-zero W[n][m] has |n-m| <= 2, so both directions are 5-point gathers.
// ------------------------------------------------------------------------------------------------
struct Blur5 { float k[5]; };
__device__ __forceinline__ int refl_idx(int t, int S) { return t < 0 ? -t : (t > S - 1 ? 2 * (S - 1) - t : t); }
__device__ __forceinline__ float blur_w(const Blur5& b, int n, int m, int S) {       // W[n][m]
    float w = 0.f;
#pragma unroll
    for (int i = -2; i <= 2; ++i) w += (refl_idx(n + i, S) == m) ? b.k[i + 2] : 0.f;
    return w;
}
// AXIS 0: along W (contiguous), 1: along H.  grid: ceil(planes*H*W / 256)
template <int AXIS, bool ADJ> __global__ void k_blur5(Blur5 b, const float* __restrict__ in, float* __restrict__ out, long long total, int H, int W) {
    const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (e >= total) return;
    const int x = int(e % W), y = int((e / W) % H);
    const int S = AXIS == 0 ? W : H, c = AXIS == 0 ? x : y;
    const long long stride = AXIS == 0 ? 1 : W;
    const float* base = in + (e - (long long)c * stride);
    float acc = 0.f;
#pragma unroll
    for (int dlt = -2; dlt <= 2; ++dlt) {
        const int o = c + dlt;
        if (o < 0 || o >= S) continue;
        const float w = ADJ ? blur_w(b, o, c, S) : blur_w(b, c, o, S);
        acc += w * base[(long long)o * stride];
    }
    out[e] = acc;
}

// Object pre-blur (models.py:267-284) without the reference's gather tensor: pass 1 reads the ROI of every (sample, object mode, slice)
// straight from the dense object and blurs it along x (reflect INSIDE the patch, as torchvision pads the gathered patch), pass 2 is
// k_blur5<1> along y.  BLUR = false: the plain ROI planes (loss_simlar without blur).  grid (ceil(B*M*Z*N*N / 256), 2 planes: amp, phase)
template <bool BLUR> __global__ void k_roi_blurx(Dims d, Blur5 bl, const int64_t* __restrict__ idx, const int32_t* __restrict__ crop,
                                                 const float* __restrict__ obja, const float* __restrict__ objp, float* __restrict__ outa,
                                                 float* __restrict__ outp) {
    const long long NN = (long long)d.N * d.N, total = (long long)d.B * d.M * d.Z * NN;
    const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (e >= total) return;
    const int x = int(e % d.N), y = int((e / d.N) % d.N), mz = int((e / NN) % (d.M * d.Z)), b = int(e / (NN * d.M * d.Z));
    const int64_t n0 = idx[b];
    const int cy = crop[2 * n0], cx = crop[2 * n0 + 1];
    const float* src = (blockIdx.y ? objp : obja) + ((size_t)mz * d.Noy + cy + y) * d.Nox + cx;
    float acc;
    if (BLUR) {
        acc = 0.f;
#pragma unroll
        for (int dlt = -2; dlt <= 2; ++dlt) {
            const int o = x + dlt;
            if (o < 0 || o >= d.N) continue;
            acc += blur_w(bl, x, o, d.N) * src[o];
        }
    } else acc = src[x];
    (blockIdx.y ? outp : outa)[e] = acc;
}
// adjoint of the pair above for patch gradients that were already taken through the adjoint of the y pass (k_blur5<1, true>): adjoint
// along x (5-point gather) and scatter-ADD into the dense object gradient.  Null plane pointers are skipped.
template <bool BLUR> __global__ void k_roi_blurx_adj_scatter(Dims d, Blur5 bl, const int64_t* __restrict__ idx, const int32_t* __restrict__ crop,
                                                             const float* __restrict__ ga_patch, const float* __restrict__ gp_patch,
                                                             float* __restrict__ g_obja, float* __restrict__ g_objp) {
    const float* in = blockIdx.y ? gp_patch : ga_patch;
    float* out = blockIdx.y ? g_objp : g_obja;
    if (!in || !out) return;
    const long long NN = (long long)d.N * d.N, total = (long long)d.B * d.M * d.Z * NN;
    const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (e >= total) return;
    const int x = int(e % d.N), y = int((e / d.N) % d.N), mz = int((e / NN) % (d.M * d.Z)), b = int(e / (NN * d.M * d.Z));
    const int64_t n0 = idx[b];
    const int cy = crop[2 * n0], cx = crop[2 * n0 + 1];
    float acc;
    if (BLUR) {
        const float* row = in + (e - x);
        acc = 0.f;
#pragma unroll
        for (int dlt = -2; dlt <= 2; ++dlt) {
            const int o = x + dlt;
            if (o < 0 || o >= d.N) continue;
            acc += blur_w(bl, o, x, d.N) * row[o];
        }
    } else acc = in[e];
    atomicAdd(out + ((size_t)mz * d.Noy + cy + y) * d.Nox + cx + x, acc);
}

// loss_simlar (losses.py:106-141) on ROI planes (B,M,Z,N,N) (from k_roi_blurx / k_blur5: blurred or plain): area interpolation =
// adaptive average pooling to (Zo,Yo,Xo) (window of output i over an axis of length I: [floor(i I / O), ceil((i+1) I / O)) ), times
// occu_m, unbiased std over the object modes, mean over (B,Zo,Yo,Xo).  One thread per output cell; M <= 8.
struct SimlarDims { int B, M, Z, N, Zo, Yo, Xo; };
__device__ __forceinline__ void pool_window(int i, int I, int O, int& lo, int& hi) {
    lo = (int)(((long long)i * I) / O);
    hi = (int)((((long long)(i + 1)) * I + O - 1) / O);
}
__device__ __forceinline__ float simlar_cell(const SimlarDims& d, const float* __restrict__ plane, const float* __restrict__ occu, long long cell,
                                             float (&v)[8], float& mean, int& z0, int& z1, int& y0, int& y1, int& x0, int& x1, int& b) {
    const int xo = int(cell % d.Xo), yo = int((cell / d.Xo) % d.Yo), zo = int((cell / ((long long)d.Xo * d.Yo)) % d.Zo);
    b = int(cell / ((long long)d.Xo * d.Yo * d.Zo));
    pool_window(zo, d.Z, d.Zo, z0, z1); pool_window(yo, d.N, d.Yo, y0, y1); pool_window(xo, d.N, d.Xo, x0, x1);
    const float inv = 1.0f / float((z1 - z0) * (y1 - y0) * (x1 - x0));
    mean = 0.f;
    for (int m = 0; m < d.M; ++m) {
        const float* pl = plane + ((size_t)b * d.M + m) * d.Z * d.N * d.N;
        float acc = 0.f;
        for (int z = z0; z < z1; ++z)
            for (int y = y0; y < y1; ++y)
                for (int x = x0; x < x1; ++x) acc += pl[((size_t)z * d.N + y) * d.N + x];
        v[m] = acc * inv * occu[m];
        mean += v[m];
    }
    mean /= float(d.M);
    float var = 0.f;
    for (int m = 0; m < d.M; ++m) var += (v[m] - mean) * (v[m] - mean);
    return sqrtf(var / float(d.M - 1));
}
// sum_out[0] += scale * sum over cells of std  (scale = weight / number of cells)
__global__ void k_simlar_fwd(SimlarDims d, const float* __restrict__ plane, const float* __restrict__ occu, double scale, double* sum_out) {
    const long long cells = (long long)d.B * d.Zo * d.Yo * d.Xo;
    float acc[1] = {0.f};
    for (long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x; c < cells; c += (long long)gridDim.x * blockDim.x) {
        float v[8], mean; int z0, z1, y0, y1, x0, x1, b;
        acc[0] += simlar_cell(d, plane, occu, c, v, mean, z0, z1, y0, y1, x0, x1, b);
    }
    __shared__ float red[32];
    block_sum<1>(acc, red);
    if (threadIdx.x == 0) atomicAdd(sum_out, scale * (double)acc[0]);
}
// g_plane (zeroed by the caller) += upstream * d(loss)/d(plane); windows may overlap, hence atomics
__global__ void k_simlar_bwd(SimlarDims d, const float* __restrict__ plane, const float* __restrict__ occu, float scale, const float* __restrict__ up,
                             float* __restrict__ g_plane) {
    const long long cells = (long long)d.B * d.Zo * d.Yo * d.Xo;
    const float u = up[0] * scale;
    for (long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x; c < cells; c += (long long)gridDim.x * blockDim.x) {
        float v[8], mean; int z0, z1, y0, y1, x0, x1, b;
        const float sd = simlar_cell(d, plane, occu, c, v, mean, z0, z1, y0, y1, x0, x1, b);
        if (!(sd > 0.f)) continue;
        const float inv = 1.0f / float((z1 - z0) * (y1 - y0) * (x1 - x0));
        for (int m = 0; m < d.M; ++m) {
            const float g = u * (v[m] - mean) / (float(d.M - 1) * sd) * occu[m] * inv;
            float* pl = g_plane + ((size_t)b * d.M + m) * d.Z * d.N * d.N;
            for (int z = z0; z < z1; ++z)
                for (int y = y0; y < y1; ++y)
                    for (int x = x0; x < x1; ++x) atomicAdd(pl + ((size_t)z * d.N + y) * d.N + x, g);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// per-iteration object constraints (constraints.py:83-114, 165-208), SURVEY 8f rank 3: streaming passes over the object, in place
// ------------------------------------------------------------------------------------------------
// 1-D Gaussian blur along one axis of a dense array viewed as (outer, L, inner): obj_rblur (x then y, reflect padding: torchvision
// gaussian_blur, constraints.py:94-97) and obj_zblur (z, replicate padding: Conv1d(padding_mode='replicate'), image_proc.py:443-455)
struct BlurTaps { float k[15]; int n; };
template <int PAD /*0 reflect, 1 replicate*/>
__global__ void k_blur_axis(BlurTaps t, const float* __restrict__ in, float* __restrict__ out, long long total, int L, long long inner) {
    const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (e >= total) return;
    const int l = int((e / inner) % L);
    const float* base = in + (e - (long long)l * inner);
    const int h = t.n >> 1;
    float acc = 0.f;
    for (int i = 0; i < t.n; ++i) {
        int o = l + i - h;
        if (PAD == 0) { if (o < 0) o = -o; if (o > L - 1) o = 2 * (L - 1) - o; o = max(0, min(L - 1, o)); }
        else o = max(0, min(L - 1, o));
        acc += t.k[i] * base[(long long)o * inner];
    }
    out[e] = acc;
}

struct ObjConstraints {
    int mirrored_on; float mirrored_relax, mirrored_scale, mirrored_power;
    int thresh_on;   float thresh_relax, thresh_lo, thresh_hi;
    int postiv_on;   float postiv_relax; int postiv_subtract_min;
};
__global__ void k_obj_min(const float* __restrict__ p, long long n, float* out /* pre-set to +inf */) {
    float m = INFINITY;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) m = fminf(m, p[i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fminf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) {            // float min through the ordered-int trick (valid for any sign)
        int* oi = reinterpret_cast<int*>(out);
        if (m >= 0.f) atomicMin(oi, __float_as_int(m)); else atomicMax(reinterpret_cast<unsigned*>(oi), __float_as_uint(m));
    }
}
// mirrored_amp -> obja_thresh -> objp_postiv in the reference's order (constraints.py:240-243), one read and one write per voxel
__global__ void k_obj_voxel_constraints(ObjConstraints c, float* __restrict__ a, float* __restrict__ p, long long n, const float* __restrict__ pmin) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    float av = a[i], pv = p[i];
    if (c.mirrored_on) {
        const float vp = powf(fmaxf(pv, 0.f), c.mirrored_power);
        av = c.mirrored_relax * av + (1.f - c.mirrored_relax) * (1.f - c.mirrored_scale * vp);
    }
    if (c.thresh_on) av = c.thresh_relax * av + (1.f - c.thresh_relax) * fminf(fmaxf(av, c.thresh_lo), c.thresh_hi);
    if (c.postiv_on) {
        const float mod = c.postiv_subtract_min ? pv - pmin[0] : fmaxf(pv, 0.f);
        pv = c.postiv_relax * pv + (1.f - c.postiv_relax) * mod;
    }
    if (c.mirrored_on || c.thresh_on) a[i] = av;
    if (c.postiv_on) p[i] = pv;
}

// ------------------------------------------------------------------------------------------------
// fused multi-tensor Adam (torch.optim.Adam semantics: no amsgrad, no weight decay, no maximize), reconstruction.py:759
// ------------------------------------------------------------------------------------------------
struct AdamTensors {
    float* p[8];
    const float* g[8];
    float* m[8];
    float* v[8];
    float* step[8];         // per-tensor step counters (device float32 scalars: torch.optim.Adam's state['step'])
    float lr[8];
    long long n[8];
    int count;
    float beta1, beta2, eps;
};
__global__ void k_adam_advance(AdamTensors a) {
    if (threadIdx.x < a.count) a.step[threadIdx.x][0] += 1.0f;
}
// every tensor carries its OWN step (torch.optim.Adam keeps one per parameter: a tensor whose start_iter comes later starts its
// bias correction at 1); the counters were already advanced for this step by k_adam_advance
__global__ void k_adam(AdamTensors a) {
    const int ti = blockIdx.y;
    if (ti >= a.count) return;
    const double t = (double)a.step[ti][0];
    const float bc1 = float(1.0 - pow((double)a.beta1, t));
    const float bc2s = float(sqrt(1.0 - pow((double)a.beta2, t)));
    const float step_size = a.lr[ti] / bc1;
    float* __restrict__ p = a.p[ti];
    const float* __restrict__ g = a.g[ti];
    float* __restrict__ m = a.m[ti];
    float* __restrict__ v = a.v[ti];
    const long long n = a.n[ti];
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float gi = g[i];
        const float mi = a.beta1 * m[i] + (1.0f - a.beta1) * gi;
        const float vi = a.beta2 * v[i] + (1.0f - a.beta2) * gi * gi;
        m[i] = mi; v[i] = vi;
        const float denom = sqrtf(vi) / bc2s + a.eps;
        p[i] -= step_size * (mi / denom);
    }
}

}  // namespace ptyb
