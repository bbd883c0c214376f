// matrix.cuh — pairwise consistency matrix and greedy consensus growth (integer / bit kernels).
//
// K3 of SURVEY.md §2.2: the interval-overlap rule of computeIndependentSubgraph
// (/root/reference/src/consensus.cpp:157-159) decides which pairs need a solve; the greedy set growth is the
// row-AND + popcount form of "candidate k joins iff it agrees with every member" (BASELINE.json north_star).
// All HBM-bound bit work: rows are written as packed 32-bit words, one warp ballot per word (coalesced).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ipcb {

__device__ __forceinline__ bool intervals_overlap(int a0, int b0, int a1, int b1) { return (min(b0, b1) - max(a0, a1)) > 0; }

// per row j (time order): number of earlier candidates i < j whose interval overlaps positively. One CTA per row.
__global__ void overlap_count(const int* __restrict__ lo, const int* __restrict__ hi, int n, int* __restrict__ counts) {
    const int j = blockIdx.x;
    const int a = lo[j], b = hi[j];
    int c = 0;
    for (int i = threadIdx.x; i < j; i += blockDim.x) c += intervals_overlap(lo[i], hi[i], a, b) ? 1 : 0;
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    __shared__ int s[32];
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) { int t = 0; for (int w = 0; w < (blockDim.x + 31) / 32; ++w) t += s[w]; counts[j] = t; }
}
// exclusive scan of the row counts (n <= a few 10^4: one CTA, sequential over chunks of blockDim)
__global__ void row_offsets(const int* __restrict__ counts, int n, int base, int* __restrict__ rowptr, int* __restrict__ total) {
    __shared__ int s[1024];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = base;
    __syncthreads();
    for (int c0 = 0; c0 < n; c0 += blockDim.x) {
        const int i = c0 + threadIdx.x;
        const int v = i < n ? counts[i] : 0;
        s[threadIdx.x] = v;
        __syncthreads();
        for (int o = 1; o < blockDim.x; o <<= 1) {
            int t = threadIdx.x >= o ? s[threadIdx.x - o] : 0;
            __syncthreads();
            s[threadIdx.x] += t;
            __syncthreads();
        }
        if (i < n) rowptr[i] = carry + s[threadIdx.x] - v;
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) carry += s[threadIdx.x];
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}
// fill the check list: checks [0, n) are the diagonal (fast) checks in time order, then for every row j its overlapping
// pairs (member = order[i], cand = order[j]) in increasing i. One CTA per row, ballot compaction keeps the order.
__global__ void overlap_fill(const int* __restrict__ lo, const int* __restrict__ hi, const int* __restrict__ order, int n,
                             const int* __restrict__ rowptr, int* __restrict__ member, int* __restrict__ cand, int* __restrict__ pair_i,
                             int* __restrict__ pair_j) {
    const int j = blockIdx.x;
    if (threadIdx.x == 0) { member[j] = -1; cand[j] = order[j]; }
    const int a = lo[j], b = hi[j];
    __shared__ int s_base;
    __shared__ int s_w[32];
    if (threadIdx.x == 0) s_base = rowptr[j];
    __syncthreads();
    const int nw = (blockDim.x + 31) / 32, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int i0 = 0; i0 < j; i0 += blockDim.x) {
        const int i = i0 + threadIdx.x;
        const bool ov = i < j && intervals_overlap(lo[i], hi[i], a, b);
        const unsigned m = __ballot_sync(0xffffffffu, ov);
        if (lane == 0) s_w[w] = __popc(m);
        __syncthreads();
        int off = s_base;
        for (int q = 0; q < w; ++q) off += s_w[q];
        if (ov) {
            const int p = off + __popc(m & ((1u << lane) - 1));
            member[p] = order[i]; cand[p] = order[j]; pair_i[p] = i; pair_j[p] = j;
        }
        __syncthreads();
        if (threadIdx.x == 0) { int t = 0; for (int q = 0; q < nw; ++q) t += s_w[q]; s_base += t; }
        __syncthreads();
    }
}
// matrix without the solved pairs: diagonal = fast verdict, non-overlapping (r, c) = diag[r] & diag[c], overlapping = 0 (set later).
// One warp per output word: lane = column, ballot = word, coalesced row-major stores.
__global__ void matrix_init(const int* __restrict__ lo, const int* __restrict__ hi, const unsigned char* __restrict__ diag, int n, int words,
                            uint32_t* __restrict__ rows) {
    const int lane = threadIdx.x & 31;
    const long long gw = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (gw >= (long long)n * words) return;
    const int r = (int)(gw / words), wd = (int)(gw % words);
    const int c = wd * 32 + lane;
    bool bit = false;
    if (c < n) {
        if (c == r) bit = diag[r] != 0;
        else bit = diag[r] && diag[c] && !intervals_overlap(lo[r], hi[r], lo[c], hi[c]);
    }
    const unsigned m = __ballot_sync(0xffffffffu, bit);
    if (lane == 0) rows[(size_t)r * words + wd] = m;
}
// solved pair verdicts into both triangles
__global__ void matrix_scatter(const unsigned char* __restrict__ verdict, const int* __restrict__ pair_i, const int* __restrict__ pair_j, int first,
                               int n_checks, int words, uint32_t* __restrict__ rows) {
    const int p = first + blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_checks || !verdict[p]) return;
    const int i = pair_i[p], j = pair_j[p];
    atomicOr(&rows[(size_t)j * words + (i >> 5)], 1u << (i & 31));
    atomicOr(&rows[(size_t)i * words + (j >> 5)], 1u << (j & 31));
}
// greedy consensus growth: candidate k (matrix order) joins iff it is consistent with every current member:
// (S & ~row_k) == 0 over all words, i.e. popcount(S & row_k) == popcount(S). One CTA walks the candidates.
__global__ void greedy_consensus(const uint32_t* __restrict__ rows, int n, int words, uint32_t* __restrict__ S, unsigned char* __restrict__ in_set) {
    for (int w = threadIdx.x; w < words; w += blockDim.x) S[w] = 0;
    __syncthreads();
    for (int k = 0; k < n; ++k) {
        const uint32_t* row = rows + (size_t)k * words;
        int bad = 0;
        for (int w = threadIdx.x; w < words; w += blockDim.x) bad |= (S[w] & ~row[w]) != 0;
        const bool self = (row[k >> 5] >> (k & 31)) & 1;      // its own fast check
        const int any_bad = __syncthreads_or(bad);
        if (threadIdx.x == 0) {
            const bool join = self && !any_bad;
            in_set[k] = join ? 1 : 0;
            if (join) S[k >> 5] |= 1u << (k & 31);
        }
        __syncthreads();
    }
}

}  // namespace ipcb
