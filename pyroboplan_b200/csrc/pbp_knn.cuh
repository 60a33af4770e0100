// Brute-force k nearest neighbours in joint space for batched PRM construction.
//
// Restates Graph.get_nearest_neighbors (/root/reference/src/pyroboplan/planning/graph.py:271-301)
// as used by PRMPlanner.connect_node (planning/prm.py:186-187): all nodes with
// configuration_distance(q_i, q_j) <= radius, sorted by distance (stable: ties keep index order),
// first k.  "Predecessor" mode mirrors the reference's incremental construction, where node i only
// sees the nodes inserted before it (j < i).
//
// KNN_SPLIT threads share one query node: candidates stream through shared memory in tiles and
// thread s of a query takes candidates s, s + KNN_SPLIT, ... of every tile, keeping its own sorted
// top-k in thread-interleaved shared memory (almost every candidate fails the first comparison
// against the current k-th best); the KNN_SPLIT lists are merged at the end, ties resolved by
// index.  (One thread per query left a 20 k-node roadmap with ~1 block per SM: 4.2 ms; splitting
// gives the SMs 8x the threads.)  Distances accumulate in FP64: the nodes are float32, hence exact
// in FP64, and the reference orders float64 norms -- 20 k x 20 k x 9 DFMA is ~3.6 GFLOP, noise for
// the chip.
#pragma once
#include <stdint.h>

namespace pbp {

constexpr int KNN_THREADS = 128;
constexpr int KNN_SPLIT = 8;                          // threads per query
constexpr int KNN_QUERIES = KNN_THREADS / KNN_SPLIT;  // queries per block
constexpr int KNN_TILE = 128;
constexpr int KNN_MAX_K = 64;
constexpr int KNN_MAX_NQ = 64;

inline size_t knn_smem_bytes(int nq, int k) {
    return (size_t)KNN_TILE * nq * 4 + (size_t)k * KNN_THREADS * (8 + 4) + 16;
}

// NQ > 0: query coordinates in registers; NQ == 0: any nq <= KNN_MAX_NQ (query re-read from global memory)
template <int NQ>
__global__ void __launch_bounds__(KNN_THREADS)
k_knn(const float *__restrict__ nodes, int64_t n, int nq_dyn, int k, double radius2, int predecessors_only,
      int64_t block_first, int64_t block_stride, int32_t *__restrict__ idx_out, float *__restrict__ dist_out) {
    extern __shared__ __align__(16) unsigned char knn_smem[];
    const int nq = NQ > 0 ? NQ : nq_dyn;
    double *sd = reinterpret_cast<double *>(knn_smem);                                   // [k][KNN_THREADS]
    int32_t *si = reinterpret_cast<int32_t *>(knn_smem + (size_t)k * KNN_THREADS * 8);     // [k][KNN_THREADS]
    float *tile = reinterpret_cast<float *>(knn_smem + (size_t)k * KNN_THREADS * 12);      // [KNN_TILE][nq]
    const int tid = threadIdx.x;
    const int sub = tid % KNN_SPLIT;                      // which slice of every tile this thread scans
    // this block's KNN_QUERIES query nodes: query block (block_first + blockIdx.x * block_stride) of the node list --
    // ranks of a multi-GPU job interleave the query blocks (the work of a predecessor query grows with its index);
    // results land in the caller's LOCAL rows blockIdx.x * KNN_QUERIES ...
    const int64_t i0 = (block_first + (int64_t)blockIdx.x * block_stride) * KNN_QUERIES;
    const int64_t i = i0 + tid / KNN_SPLIT;
    const int64_t row = (int64_t)blockIdx.x * KNN_QUERIES + tid / KNN_SPLIT;
    const bool live = i < n;
    float qi[NQ > 0 ? NQ : 1];
    if (NQ > 0 && live) {
#pragma unroll
        for (int c = 0; c < NQ; c++) qi[c] = nodes[i * NQ + c];
    }
    for (int s = 0; s < k; s++) { sd[s * KNN_THREADS + tid] = INFINITY; si[s * KNN_THREADS + tid] = -1; }
    double worst = INFINITY;   // this thread's k-th best so far
    const int64_t j_end = predecessors_only ? min(n, i0 + KNN_QUERIES) : n;
    for (int64_t j0 = 0; j0 < j_end; j0 += KNN_TILE) {
        const int cnt = (int)min((int64_t)KNN_TILE, j_end - j0);
        __syncthreads();
        for (int t = tid; t < cnt * nq; t += KNN_THREADS) tile[t] = nodes[j0 * nq + t];
        __syncthreads();
        if (!live) continue;
        const int lim = predecessors_only ? (int)min((int64_t)cnt, i - j0) : cnt;   // j < i only
        for (int jj = sub; jj < lim; jj += KNN_SPLIT) {
            const int64_t j = j0 + jj;
            if (j == i) continue;
            double d2 = 0.0;
            if (NQ > 0) {
#pragma unroll
                for (int c = 0; c < NQ; c++) {
                    const double d = (double)qi[c] - (double)tile[jj * NQ + c];
                    d2 = fma(d, d, d2);
                }
            } else {
                for (int c = 0; c < nq; c++) {
                    const double d = (double)nodes[i * nq + c] - (double)tile[jj * nq + c];
                    d2 = fma(d, d, d2);
                }
            }
            if (!(d2 <= radius2) || !(d2 < worst)) continue;
            // insert, keeping the list sorted (strict <: an equal distance stays behind earlier indices)
            int s = k - 1;
            while (s > 0 && sd[(s - 1) * KNN_THREADS + tid] > d2) {
                sd[s * KNN_THREADS + tid] = sd[(s - 1) * KNN_THREADS + tid];
                si[s * KNN_THREADS + tid] = si[(s - 1) * KNN_THREADS + tid];
                s--;
            }
            sd[s * KNN_THREADS + tid] = d2;
            si[s * KNN_THREADS + tid] = (int32_t)j;
            worst = sd[(k - 1) * KNN_THREADS + tid];
        }
    }
    __syncthreads();
    if (!live || sub != 0) return;
    // merge the KNN_SPLIT sorted lists of this query: smallest (distance, index) first
    int head[KNN_SPLIT];
#pragma unroll
    for (int t = 0; t < KNN_SPLIT; t++) head[t] = 0;
    for (int s = 0; s < k; s++) {
        double bd = INFINITY;
        int32_t bi = -1;
        int bt = -1;
#pragma unroll
        for (int t = 0; t < KNN_SPLIT; t++) {
            if (head[t] >= k) continue;
            const double d = sd[head[t] * KNN_THREADS + tid + t];
            const int32_t ix = si[head[t] * KNN_THREADS + tid + t];
            if (ix >= 0 && (d < bd || (d == bd && ix < bi))) { bd = d; bi = ix; bt = t; }
        }
        if (bt >= 0) {
#pragma unroll
            for (int t = 0; t < KNN_SPLIT; t++)
                if (t == bt) head[t]++;
        }
        idx_out[row * k + s] = bi;
        dist_out[row * k + s] = bi >= 0 ? (float)sqrt(bd) : INFINITY;
    }
}

}  // namespace pbp
