// Two-stage in-place FFT of length N = N1*N2 along the rows of a shared-memory slab.
//
//   forward  (DIR=-1): natural order -> "scrambled" order, frequency q lands at pos(q)
//   inverse  (DIR=+1): scrambled -> natural (unnormalised; callers fold 1/N into their own scale)
//
// Each work item is one small register DFT (dft_regs.cuh):
//   stage 1 item (row, j<N2):  v[k] = x[j + N2*k], DFT<N1> over k -> k1, times W_N^{j*k1}, written back at j + N2*k1
//   stage 2 item (row, k1<N1): v[j] = x[j + N2*k1], DFT<N2> over j -> k2, written back at k2 + N2*k1   (= X[k1 + N1*k2])
// so no item ever writes a location another item of the same stage reads, and one barrier between the
// stages is all that is needed.  Rows are padded by one element per N2 group (addr()) which makes the
// stride-N2 accesses of the frequency-ordered phases bank-conflict free for 64-bit words; the row stride RS
// takes care of the accesses that run across the rows of a slab.
//
// The item functions are __host__ __device__: tests/csrc_host/test_fft_host.cu runs them on the CPU.
#pragma once
#include "dft_regs.cuh"

namespace ptyb {

template <int N1_, int N2_> struct RowFFT {
    static constexpr int N1 = N1_, N2 = N2_, N = N1_ * N2_;
    static constexpr int ROW_ELEMS = N + N1;                                  // N + N/N2 padding elements
#ifndef PTYB_RS_MOD16
#define PTYB_RS_MOD16 2
#endif
    // row stride == 2 (mod 16) elements, i.e. 4 banks: the transposed accesses (lanes over the 8 rows of a slab, the two
    // half-warp halves on neighbouring work items, which sit 2 banks apart) are then conflict free for 64-bit words
    static constexpr int RS = (ROW_ELEMS + 15 - PTYB_RS_MOD16) / 16 * 16 + PTYB_RS_MOD16;

    PTYB_HD static int addr(int n) { return n + n / N2; }
    PTYB_HD static int pos(int q) { return (q / N1) + N2 * (q % N1); }        // where frequency q sits after fwd
    PTYB_HD static int apos(int q) { return addr(pos(q)); }

    // twN[n] = exp(-2*pi*i*n/N), n in [0,N)
    template <int DIR> PTYB_HD static float2 tw(const float2* twN, int e) {
        float2 w = twN[e];
        if (DIR > 0) w.y = -w.y;
        return w;
    }

    PTYB_HD static void fwd_stage1(float2* row, int j, const float2* twN) {
        float2 v[N1];
#pragma unroll
        for (int k = 0; k < N1; ++k) v[k] = row[addr(j + N2 * k)];
        Dft<N1, -1, false>::run(v);
#pragma unroll
        for (int k1 = 0; k1 < N1; ++k1) {
            float2 r = (k1 == 0) ? v[k1] : cmul(v[k1], tw<-1>(twN, j * k1));
            row[addr(j + N2 * k1)] = r;
        }
    }
    PTYB_HD static void fwd_stage2(float2* row, int k1) {
        float2 v[N2];
#pragma unroll
        for (int j = 0; j < N2; ++j) v[j] = row[addr(j + N2 * k1)];
        Dft<N2, -1, false>::run(v);
#pragma unroll
        for (int k2 = 0; k2 < N2; ++k2) row[addr(k2 + N2 * k1)] = v[k2];
    }
    PTYB_HD static void inv_stage2(float2* row, int k1, const float2* twN) {
        float2 v[N2];
#pragma unroll
        for (int k2 = 0; k2 < N2; ++k2) v[k2] = row[addr(k2 + N2 * k1)];
        Dft<N2, +1, false>::run(v);
#pragma unroll
        for (int j = 0; j < N2; ++j) {
            float2 r = (k1 == 0) ? v[j] : cmul(v[j], tw<+1>(twN, j * k1));
            row[addr(j + N2 * k1)] = r;
        }
    }
    PTYB_HD static void inv_stage1(float2* row, int j) {
        float2 v[N1];
#pragma unroll
        for (int k1 = 0; k1 < N1; ++k1) v[k1] = row[addr(j + N2 * k1)];
        Dft<N1, +1, false>::run(v);
#pragma unroll
        for (int k = 0; k < N1; ++k) row[addr(j + N2 * k)] = v[k];
    }

    // Register-fed / register-draining variants of the outer stages, for kernels that load the first stage's operands straight
    // from global memory and store the last stage's results straight to global memory (general_kernels.cuh, stages A and C):
    //   fwd_stage1_regs: v[k] = x[j + N2*k] in registers           -> row (as fwd_stage1 leaves it)
    //   fwd_stage2_regs: row (as fwd_stage1 leaves it)             -> v[k2] = X[k1 + N1*k2] in registers
    //   inv_stage2_regs: v[k2] = X[k1 + N1*k2] in registers        -> row (as inv_stage2 leaves it)
    //   inv_stage1_regs: row (as inv_stage2 leaves it)             -> v[k] = x[j + N2*k] in registers (unnormalised)
    PTYB_HD static void fwd_stage1_regs(float2* row, int j, float2 (&v)[N1], const float2* twN) {
        Dft<N1, -1, false>::run(v);
#pragma unroll
        for (int k1 = 0; k1 < N1; ++k1) row[addr(j + N2 * k1)] = (k1 == 0) ? v[k1] : cmul(v[k1], tw<-1>(twN, j * k1));
    }
    PTYB_HD static void fwd_stage2_regs(const float2* row, int k1, float2 (&v)[N2]) {
#pragma unroll
        for (int j = 0; j < N2; ++j) v[j] = row[addr(j + N2 * k1)];
        Dft<N2, -1, false>::run(v);
    }
    PTYB_HD static void inv_stage2_regs(float2* row, int k1, float2 (&v)[N2], const float2* twN) {
        Dft<N2, +1, false>::run(v);
#pragma unroll
        for (int j = 0; j < N2; ++j) row[addr(j + N2 * k1)] = (k1 == 0) ? v[j] : cmul(v[j], tw<+1>(twN, j * k1));
    }
    PTYB_HD static void inv_stage1_regs(const float2* row, int j, float2 (&v)[N1]) {
#pragma unroll
        for (int k1 = 0; k1 < N1; ++k1) v[k1] = row[addr(j + N2 * k1)];
        Dft<N1, +1, false>::run(v);
    }

#ifdef __CUDACC__
    // Block-wide drivers over `rows` rows of a slab with row stride RS.  All threads of the block must call.
    __device__ static void forward(float2* slab, int rows, const float2* twN) {
        for (int it = threadIdx.x; it < rows * N2; it += blockDim.x) fwd_stage1(slab + (it / N2) * RS, it % N2, twN);
        __syncthreads();
        for (int it = threadIdx.x; it < rows * N1; it += blockDim.x) fwd_stage2(slab + (it / N1) * RS, it % N1);
        __syncthreads();
    }
    __device__ static void inverse(float2* slab, int rows, const float2* twN) {
        for (int it = threadIdx.x; it < rows * N1; it += blockDim.x) inv_stage2(slab + (it / N1) * RS, it % N1, twN);
        __syncthreads();
        for (int it = threadIdx.x; it < rows * N2; it += blockDim.x) inv_stage1(slab + (it / N2) * RS, it % N2);
        __syncthreads();
    }
    // half transforms, for kernels that fuse their pointwise work into the register stage in the middle:
    //   inverse_first  : inverse stage over k2 (all rows), then barrier -> the caller runs inverse stage over k1 itself
    //   forward_last   : forward stage over j (all rows), then barrier  -> after the caller ran the forward stage over k
    //   forward_first / inverse_last likewise for the other nesting
    __device__ static void inverse_first(float2* slab, int rows, const float2* twN) {
        for (int it = threadIdx.x; it < rows * N1; it += blockDim.x) inv_stage2(slab + (it / N1) * RS, it % N1, twN);
        __syncthreads();
    }
    __device__ static void forward_last(float2* slab, int rows) {
        for (int it = threadIdx.x; it < rows * N1; it += blockDim.x) fwd_stage2(slab + (it / N1) * RS, it % N1);
        __syncthreads();
    }
    __device__ static void forward_first(float2* slab, int rows, const float2* twN) {
        for (int it = threadIdx.x; it < rows * N2; it += blockDim.x) fwd_stage1(slab + (it / N2) * RS, it % N2, twN);
        __syncthreads();
    }
    __device__ static void inverse_last(float2* slab, int rows) {
        for (int it = threadIdx.x; it < rows * N2; it += blockDim.x) inv_stage1(slab + (it / N2) * RS, it % N2);
        __syncthreads();
    }
    // twN table fill (exact via sincospif); call before first use, followed by __syncthreads()
    __device__ static void fill_twiddles(float2* twN) {
        for (int n = threadIdx.x; n < N; n += blockDim.x) {
            float s, c;
            sincospif(-2.0f * float(n) / float(N), &s, &c);
            twN[n] = make_float2(c, s);
        }
    }
#endif
};

}  // namespace ptyb
