// 'sparse' grouping of scan positions on the GPU (reference make_batches, reconstruction.py:540-587: the greedy loop that the
// reference documents as "more than 10 min on a CPU" for a 256x256 scan).
//
// The algorithm is strictly sequential over the points -- point i joins the group whose NEAREST member is FARTHEST from it, given
// every assignment made before -- so one CTA of 1024 threads walks the points in order.  Per point: every thread takes a stride of
// the already assigned points j < i, computes |pos_j - pos_i|^2 in float64 (the reference compares float64 cdist distances; the
// square root is monotonic) and folds it into the per-group minimum in shared memory (atomicMin on the bit pattern of the
// non-negative double); then the block picks the group with the largest minimum, first index on ties like np.argmax.
// Work N^2 / 2 distance evaluations: 65536 points -> 2.1 G, well under a second on one SM, against minutes for the Python loop.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ptyb {

constexpr int GROUP_THREADS = 1024;

// pos: (n) double2, the G seeds first (group g's seed at slot g), then the points in assignment order.  label: (n) out.
__global__ void __launch_bounds__(GROUP_THREADS, 1) k_sparse_groups(const double2* __restrict__ pos, int n, int G, int* label) {
    extern __shared__ unsigned long long gmin[];           // G per-group minima + 32 (value, index) pairs of reduction scratch
    unsigned long long* rv = gmin + G;
    int* ri = reinterpret_cast<int*>(rv + 32);
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    for (int g = tid; g < G && g < n; g += GROUP_THREADS) label[g] = g;
    __syncthreads();
    for (int i = G; i < n; ++i) {
        for (int g = tid; g < G; g += GROUP_THREADS) gmin[g] = ~0ull;
        __syncthreads();
        const double2 p = pos[i];
        for (int j = tid; j < i; j += GROUP_THREADS) {
            const double2 q = pos[j];
            const double dx = q.x - p.x, dy = q.y - p.y;
            atomicMin(gmin + __ldcg(label + j), (unsigned long long)__double_as_longlong(dx * dx + dy * dy));
        }
        __syncthreads();
        // argmax over the groups, first index on ties
        unsigned long long bv = 0;
        int bi = 0x7fffffff;
        for (int g = tid; g < G; g += GROUP_THREADS) {
            const unsigned long long v = gmin[g];
            if (v > bv || (v == bv && g < bi)) { bv = v; bi = g; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        if (lane == 0) { rv[wid] = bv; ri[wid] = bi; }
        __syncthreads();
        if (wid == 0) {
            bv = rv[lane]; bi = ri[lane];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const unsigned long long ov = __shfl_xor_sync(0xffffffffu, bv, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
            }
            if (lane == 0) label[i] = bi;
        }
        __syncthreads();                                   // label[i] is visible to the whole block before the next point reads it
    }
}

}  // namespace ptyb
