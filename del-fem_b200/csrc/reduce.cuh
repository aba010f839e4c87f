// reduce.cuh — deterministic two-stage reductions: fixed grid, fixed tree, finalised by the last block to finish.
// The result is bit-identical from run to run regardless of block scheduling (the partials are summed in a fixed
// order by whichever block happens to be last).
#pragma once
#include <cuda_runtime.h>

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}

// Block-level sum of K values per thread (fixed tree). Result valid in thread 0.
template <int K>
__device__ __forceinline__ void block_sum(double (&v)[K], double *s_buf /* [K*32] shared */) {
  const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int k = 0; k < K; ++k) v[k] = warp_sum(v[k]);
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < K; ++k) s_buf[k * 32 + wid] = v[k];
  }
  __syncthreads();
  if (wid == 0) {
#pragma unroll
    for (int k = 0; k < K; ++k) {
      double t = (lane < nw) ? s_buf[k * 32 + lane] : 0.0;
      v[k] = warp_sum(t);
    }
  }
  __syncthreads();
}

// Grid-level finalisation. Every block calls this with its per-thread partial sums. Returns true in ALL threads of
// the one block that finished last; in that block, thread 0 holds the grid totals in v[].
// partials: [gridDim.x * K] ; ticket: one counter (self-resetting).
template <int K>
__device__ __forceinline__ bool grid_sum_last_block(double (&v)[K], double *partials, unsigned int *ticket,
                                                    double *s_buf /* [K*32] */) {
  __shared__ bool s_last;
  block_sum<K>(v, s_buf);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < K; ++k) partials[(size_t)blockIdx.x * K + k] = v[k];
    __threadfence();
    const unsigned t = atomicAdd(ticket, 1u);
    s_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!s_last) return false;
  __threadfence();
  double acc[K];
#pragma unroll
  for (int k = 0; k < K; ++k) acc[k] = 0.0;
  for (unsigned b = threadIdx.x; b < gridDim.x; b += blockDim.x) {
#pragma unroll
    for (int k = 0; k < K; ++k) acc[k] += __ldcg(&partials[(size_t)b * K + k]);
  }
  block_sum<K>(acc, s_buf);
#pragma unroll
  for (int k = 0; k < K; ++k) v[k] = acc[k];
  if (threadIdx.x == 0) *ticket = 0u;
  return true;
}
