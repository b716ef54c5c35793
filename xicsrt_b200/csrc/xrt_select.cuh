// xrt_select.cuh -- found / lost selection on the device (reference _sort_raytrace, xicsrt_raytrace.py:229-278).
//
// The fused kernel marks found rays (and the candidates of the lost sample) as bits of a bitmap indexed by ray id.
// These kernels turn a bitmap into the ascending list of its ids -- the reference's found rays are in ray order -- by
// count / scan / emit, and pick the lost sample as the m candidates with the smallest 64-bit Philox keys (a uniform
// random subset, as the reference's shuffle gives) by a radix select, again emitted in ascending id order.  No sort
// and no library call; everything is deterministic for a given seed.
#pragma once
#include "xrt_trace.cuh"

namespace xrt {

constexpr int kSelBlock = 256;
constexpr int kSelWordsPerThread = 8;
constexpr int kSelWordsPerBlock = kSelBlock * kSelWordsPerThread;    // 2048 words = 65536 ids per block

// bits set in each block's chunk of the bitmap
__global__ void __launch_bounds__(kSelBlock) k_bits_count(const uint32_t *__restrict__ bits, uint64_t n_words,
                                                          uint32_t *__restrict__ block_sums) {
    const uint64_t w0 = (uint64_t)blockIdx.x * kSelWordsPerBlock + (uint64_t)threadIdx.x * kSelWordsPerThread;
    uint32_t c = 0;
#pragma unroll
    for (int j = 0; j < kSelWordsPerThread; ++j)
        if (w0 + j < n_words) c += __popc(__ldg(bits + w0 + j));
    __shared__ uint32_t s_warp[kSelBlock / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_down_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int i = 0; i < kSelBlock / 32; ++i) t += s_warp[i];
        block_sums[blockIdx.x] = t;
    }
}

// exclusive prefix sum of the block sums (one block; 64-bit running total) and the grand total
__global__ void __launch_bounds__(1024) k_bits_scan(const uint32_t *__restrict__ block_sums, uint32_t n_blocks,
                                                    unsigned long long *__restrict__ block_offsets,
                                                    unsigned long long *__restrict__ total) {
    __shared__ unsigned long long s_part[1024];
    const uint32_t per = (n_blocks + 1023u) / 1024u;
    const uint32_t b0 = threadIdx.x * per;
    unsigned long long c = 0;
    for (uint32_t i = 0; i < per; ++i)
        if (b0 + i < n_blocks) c += block_sums[b0 + i];
    s_part[threadIdx.x] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long run = 0;
        for (int i = 0; i < 1024; ++i) { const unsigned long long v = s_part[i]; s_part[i] = run; run += v; }
        if (total) *total = run;
    }
    __syncthreads();
    unsigned long long run = s_part[threadIdx.x];
    for (uint32_t i = 0; i < per; ++i)
        if (b0 + i < n_blocks) { block_offsets[b0 + i] = run; run += block_sums[b0 + i]; }
}

// ids of the set bits, ascending
__global__ void __launch_bounds__(kSelBlock) k_bits_emit(const uint32_t *__restrict__ bits, uint64_t n_words, uint64_t id_begin,
                                                         const unsigned long long *__restrict__ block_offsets,
                                                         uint64_t *__restrict__ ids_out, uint64_t capacity) {
    const uint64_t w0 = (uint64_t)blockIdx.x * kSelWordsPerBlock + (uint64_t)threadIdx.x * kSelWordsPerThread;
    uint32_t w[kSelWordsPerThread];
    uint32_t c = 0;
#pragma unroll
    for (int j = 0; j < kSelWordsPerThread; ++j) {
        w[j] = (w0 + j < n_words) ? __ldg(bits + w0 + j) : 0u;
        c += __popc(w[j]);
    }
    // exclusive scan of the per-thread counts over the block
    const unsigned lane = threadIdx.x & 31u;
    uint32_t inc = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= (unsigned)o) inc += v;
    }
    __shared__ uint32_t s_warp[kSelBlock / 32];
    if (lane == 31) s_warp[threadIdx.x >> 5] = inc;
    __syncthreads();
    uint32_t warp_off = 0;
    for (int i = 0; i < (int)(threadIdx.x >> 5); ++i) warp_off += s_warp[i];
    unsigned long long slot = block_offsets[blockIdx.x] + warp_off + (inc - c);
#pragma unroll
    for (int j = 0; j < kSelWordsPerThread; ++j) {
        uint32_t v = w[j];
        while (v) {
            const int b = __ffs(v) - 1;
            v &= v - 1u;
            if (slot < capacity) ids_out[slot] = id_begin + ((w0 + j) << 5) + (uint64_t)b;
            ++slot;
        }
    }
}

// 64-bit sampling keys of a list of ray ids (the key the fused kernel compared with the threshold)
__global__ void __launch_bounds__(256) k_lost_keys(const __grid_constant__ PhiloxKeys pk, uint64_t stream_id,
                                                   const uint64_t *__restrict__ ids, uint64_t n, uint64_t *__restrict__ keys) {
    for (uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (uint64_t)gridDim.x * 256) {
        PhiloxDraws dr;
        dr.init(pk, stream_id, ids[i], 0);
        keys[i] = dr.lost_key();
    }
}

// The m entries with the smallest keys of (ids, keys)[0..n), written to out in their input (= ascending id) order.
// One block: MSB-first radix select of the m-th smallest key, 8 bits per pass, then an ordered emit; ties at the
// threshold key (probability 2^-64 per pair) are broken by position.  out_count receives min(m, n).
__global__ void __launch_bounds__(1024) k_select_smallest(const uint64_t *__restrict__ ids, const uint64_t *__restrict__ keys,
                                                          uint64_t n, uint64_t m, uint64_t *__restrict__ out,
                                                          unsigned long long *__restrict__ out_count) {
    __shared__ unsigned int s_hist[256];
    __shared__ unsigned long long s_prefix, s_mask, s_need, s_run, s_ties_left;
    __shared__ unsigned int s_warp[32];
    __shared__ unsigned char s_tie_keep[1024], s_tie[1024];
    if (m >= n) {       // everything is kept
        for (uint64_t i = threadIdx.x; i < n; i += 1024) out[i] = ids[i];
        if (threadIdx.x == 0) *out_count = n;
        return;
    }
    if (threadIdx.x == 0) { s_prefix = 0ull; s_mask = 0ull; s_need = m; }
    __syncthreads();
    for (int shift = 56; shift >= 0; shift -= 8) {
        for (int i = threadIdx.x; i < 256; i += 1024) s_hist[i] = 0u;
        __syncthreads();
        const unsigned long long prefix = s_prefix, mask = s_mask;
        for (uint64_t i = threadIdx.x; i < n; i += 1024) {
            const unsigned long long k = keys[i];
            if ((k & mask) == prefix) atomicAdd(&s_hist[(k >> shift) & 0xffull], 1u);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned long long need = s_need;        // rank (1-based) of the wanted key among the keys matching the prefix
            int d = 0;
            for (; d < 256; ++d) {
                if (s_hist[d] >= need) break;
                need -= s_hist[d];
            }
            s_prefix = prefix | ((unsigned long long)d << shift);
            s_mask = mask | (0xffull << shift);
            s_need = need;
        }
        __syncthreads();
    }
    const unsigned long long kth = s_prefix;         // the m-th smallest key; s_need of the entries equal to it are kept
    if (threadIdx.x == 0) { s_run = 0ull; s_ties_left = s_need; }
    __syncthreads();
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    for (uint64_t base = 0; base < n; base += 1024) {
        const uint64_t i = base + threadIdx.x;
        const unsigned long long k = i < n ? keys[i] : ~0ull;
        const bool below = i < n && k < kth;
        bool tie = i < n && k == kth;
        // ties: keep the first s_ties_left of them in position order (rare: resolved by a serial pass of thread 0)
        s_tie_keep[threadIdx.x] = 0;
        __syncthreads();
        if (__syncthreads_or(tie)) {
            s_tie[threadIdx.x] = tie ? 1 : 0;
            __syncthreads();
            if (threadIdx.x == 0) {
                for (int t = 0; t < 1024; ++t)
                    if (s_tie[t] && s_ties_left > 0) { s_tie_keep[t] = 1; --s_ties_left; }
            }
            __syncthreads();
        }
        const bool keep = below || (tie && s_tie_keep[threadIdx.x]);
        const unsigned mk = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) s_warp[warp] = __popc(mk);
        __syncthreads();
        unsigned off = 0, tot = 0;
        for (unsigned wv = 0; wv < 32; ++wv) { if (wv < warp) off += s_warp[wv]; tot += s_warp[wv]; }
        if (keep) out[s_run + off + __popc(mk & ((1u << lane) - 1u))] = ids[i];
        __syncthreads();
        if (threadIdx.x == 0) s_run += tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) *out_count = s_run;
}

}  // namespace xrt
