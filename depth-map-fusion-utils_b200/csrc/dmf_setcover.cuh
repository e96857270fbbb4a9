// dmf_setcover.cuh -- K5: bitwise-OR combine of per-rank bitsets, K6: greedy set cover over visibility bitsets
// (Algorithms::greedySetCover, reference include/Algorithms.hpp:38-86).
#pragma once
#include "dmf_device.cuh"

// dst[w] |= OR_r src[r][w]
__global__ void k_or_reduce(u64* __restrict__ dst, const u64* __restrict__ src, int n_src, size_t words) {
    for (size_t w = blockIdx.x * (size_t)blockDim.x + threadIdx.x; w < words; w += (size_t)gridDim.x * blockDim.x) {
        u64 acc = dst[w];
        for (int r = 0; r < n_src; r++) acc |= __ldg(src + (size_t)r * words + w);
        dst[w] = acc;
    }
}

// gain[s] = |set_s \ covered| = popcount(bits[s] & ~covered)   (the set_difference of Algorithms.hpp:57), one block per set;
// sets already selected get gain 0 (they are erased from set_ids, :83)
// (stride = u64 words between consecutive sets: == words for a dense array, larger for the gathered buffer of a sharded sweep)
__global__ void __launch_bounds__(256) k_cover_gain(const u64* __restrict__ bits, const u64* __restrict__ covered, const int* __restrict__ taken,
                                                    size_t words, size_t stride, unsigned* __restrict__ gain) {
    __shared__ unsigned s_warp[8];
    const int s = blockIdx.x;
    unsigned cnt = 0;
    if (!taken[s]) {
        const u64* b = bits + (size_t)s * stride;
        for (size_t w = threadIdx.x; w < words; w += blockDim.x) cnt += __popcll(__ldg(b + w) & ~covered[w]);
    }
    for (int o = 16; o; o >>= 1) cnt += __shfl_down_sync(0xffffffffu, cnt, o);
    if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) { unsigned t = 0; for (int i = 0; i < 8; i++) t += s_warp[i]; gain[s] = t; }
}

// argmax with the reference's tie-break: strict '>' while scanning ascending set ids => lowest index among the maxima (:60).
// result[0] = selected (-1 if every gain is 0, :71), result[1] = max gain.  If the pick is accepted (gain >= 5, :73)
// the set is marked taken and OR-ed into covered by k_cover_apply.
__global__ void __launch_bounds__(1024) k_cover_pick(const unsigned* __restrict__ gain, int n_sets, int* __restrict__ result) {
    __shared__ unsigned long long s_best[32];
    // pack (gain, ~index) so that max() prefers larger gain, then smaller index
    unsigned long long best = 0;
    for (int s = threadIdx.x; s < n_sets; s += blockDim.x) {
        unsigned long long key = ((unsigned long long)gain[s] << 32) | (unsigned)(0xFFFFFFFFu - (unsigned)s);
        best = max(best, key);
    }
    for (int o = 16; o; o >>= 1) best = max(best, __shfl_down_sync(0xffffffffu, best, o));
    if ((threadIdx.x & 31) == 0) s_best[threadIdx.x >> 5] = best;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < (int)(blockDim.x >> 5); i++) best = max(best, s_best[i]);
        unsigned g = (unsigned)(best >> 32);
        result[0] = g ? (int)(0xFFFFFFFFu - (unsigned)(best & 0xFFFFFFFFu)) : -1;
        result[1] = (int)g;
    }
}

__global__ void k_cover_apply(const u64* __restrict__ bits, u64* __restrict__ covered, int* __restrict__ taken, size_t words, size_t stride, const int* __restrict__ result) {
    const int sel = result[0];
    if (sel < 0 || result[1] < 5) return;
    const u64* b = bits + (size_t)sel * stride;
    for (size_t w = blockIdx.x * (size_t)blockDim.x + threadIdx.x; w < words; w += (size_t)gridDim.x * blockDim.x) covered[w] |= b[w];
    if (blockIdx.x == 0 && threadIdx.x == 0) taken[sel] = 1;
}
