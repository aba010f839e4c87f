// micro-benchmark: latency of a dependent load that hits L2, for the load flavours the persistent kernels use, on data that
// (a) nobody has written since the launch began and (b) another SM wrote just before a grid barrier.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o l2lat l2lat.cu && ./l2lat
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void grid_barrier(unsigned *ctr, unsigned epoch) {
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
    unsigned v;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
    } while (v < epoch * gridDim.x);
  }
  __syncthreads();
}

template <int V>
__device__ __forceinline__ unsigned long long ld(const unsigned long long *p) {
  unsigned long long v;
  if (V == 0) asm volatile("ld.global.ca.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  if (V == 1) asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  if (V == 2) asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  if (V == 3) asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  if (V == 4) asm volatile("ld.global.nc.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// chain[i] holds the index of the next element; each CTA chases inside the slice that its neighbour CTA owns (and, when
// `fresh`, rewrote just before the barrier)
template <int V>
__global__ void k_lat(unsigned long long *chain, unsigned slice, unsigned hops, int fresh, unsigned *ctr, unsigned iters, long long *out) {
  const unsigned nb = (blockIdx.x + 1) % gridDim.x;
  long long total = 0;
  for (unsigned it = 1; it <= iters; ++it) {
    if (fresh) {   // rewrite the own slice (same values): every line becomes recently-written-by-this-SM
      for (unsigned i = threadIdx.x; i < slice; i += blockDim.x) {
        unsigned long long *p = chain + (size_t)blockIdx.x * slice + i;
        *p = (size_t)blockIdx.x * slice + (i * 97u + 31u) % slice;
      }
    }
    grid_barrier(ctr, it);
    if (threadIdx.x == 0) {
      unsigned long long idx = (size_t)nb * slice + (it * 13u) % slice;
      const long long t0 = clock64();
      for (unsigned h = 0; h < hops; ++h) idx = ld<V>(chain + idx);
      const long long t1 = clock64();
      if (idx == 0xFFFFFFFFFFFFull) out[1] = 1;
      total += t1 - t0;
    }
  }
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = total;
}

template <int V>
void run(const char *name, int fresh, int threads) {
  const unsigned grid = 148, slice = 8192, hops = 16, iters = 200;   // 148 x 64 KB = 9.7 MB
  unsigned long long *chain, *h = new unsigned long long[(size_t)grid * slice];
  for (unsigned b = 0; b < grid; ++b)
    for (unsigned i = 0; i < slice; ++i) h[(size_t)b * slice + i] = (size_t)b * slice + (i * 97u + 31u) % slice;
  cudaMalloc(&chain, sizeof(unsigned long long) * grid * slice);
  cudaMemcpy(chain, h, sizeof(unsigned long long) * grid * slice, cudaMemcpyHostToDevice);
  unsigned *ctr;
  long long *out, hout[2];
  cudaMalloc(&ctr, 4);
  cudaMalloc(&out, 16);
  cudaMemset(ctr, 0, 4);
  cudaMemset(out, 0, 16);
  k_lat<V><<<grid, threads>>>(chain, slice, hops, fresh, ctr, iters, out);
  cudaDeviceSynchronize();
  cudaMemcpy(hout, out, 16, cudaMemcpyDeviceToHost);
  printf("%-28s %-34s threads %4d : %7.1f cycles per dependent load  (%s)\n", name, fresh ? "written by another SM before the barrier" : "not written during the launch", threads,
         (double)hout[0] / ((double)hops * iters), cudaGetErrorString(cudaGetLastError()));
  cudaFree(chain);
  cudaFree(ctr);
  cudaFree(out);
  delete[] h;
}

int main() {
  for (int fresh = 0; fresh < 2; ++fresh) {
    run<0>("ld.global.ca", fresh, 128);
    run<1>("ld.global.cg (__ldcg)", fresh, 128);
    run<2>("ld.relaxed.gpu", fresh, 128);
    run<3>("ld.volatile", fresh, 128);
    run<4>("ld.global.nc (__ldg)", fresh, 128);
  }
  run<1>("ld.global.cg (__ldcg)", 1, 512);
  return 0;
}
