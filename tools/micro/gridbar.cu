// micro-benchmark: cost of one grid-wide barrier (148 resident CTAs) between two dependent phases, for the barrier flavours the
// ILU sweeps could use.  Each iteration: every CTA stores a value, barrier, every thread 0 loads its neighbour CTA's value (ld.cg)
// and checks it.   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gridbar gridbar.cu && ./gridbar
#include <cooperative_groups.h>
#include <cstdio>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;

template <int V>
__device__ __forceinline__ void barrier(unsigned *ctr, unsigned *flags, unsigned epoch) {
  if (V == 0 || V == 1) {   // fence + atomicAdd on 4 (V=0) / 1 (V=1) counters, relaxed poll, fence
    constexpr int W = V == 0 ? 4 : 1;
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      atomicAdd(ctr + (blockIdx.x % W), 1u);
      unsigned want[4];
      for (int w = 0; w < 4; ++w) want[w] = w < W ? epoch * ((gridDim.x + W - 1 - w) / W) : 0u;
      for (;;) {
        unsigned v0, v1, v2, v3;
        asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v0), "=r"(v1), "=r"(v2), "=r"(v3) : "l"(ctr) : "memory");
        if (v0 >= want[0] && v1 >= want[1] && v2 >= want[2] && v3 >= want[3]) break;
      }
      __threadfence();
    }
    __syncthreads();
  } else if (V == 2) {   // red.release + ld.acquire
    __syncthreads();
    if (threadIdx.x == 0) {
      asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
      unsigned v;
      do {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
      } while (v < epoch * gridDim.x);
    }
    __syncthreads();
  } else if (V == 3) {   // one flag per CTA (st.release), warp 0 polls all flags
    __syncthreads();
    if (threadIdx.x < 32) {
      if (threadIdx.x == 0) asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(flags + blockIdx.x), "r"(epoch) : "memory");
      for (;;) {
        bool ok = true;
        for (unsigned j = threadIdx.x; j < gridDim.x; j += 32) {
          unsigned v;
          asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(flags + j) : "memory");
          ok = ok && v >= epoch;
        }
        if (__all_sync(0xFFFFFFFFu, ok)) break;
      }
      __threadfence();
    }
    __syncthreads();
  } else if (V == 4) {
    cg::this_grid().sync();
  } else if (V == 5) {   // like 0 but without the fences (NOT a correct barrier for data; lower bound of the signalling alone)
    __syncthreads();
    if (threadIdx.x == 0) {
      atomicAdd(ctr + (blockIdx.x % 4), 1u);
      unsigned want[4];
      for (int w = 0; w < 4; ++w) want[w] = epoch * ((gridDim.x + 3 - w) / 4);
      for (;;) {
        unsigned v0, v1, v2, v3;
        asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v0), "=r"(v1), "=r"(v2), "=r"(v3) : "l"(ctr) : "memory");
        if (v0 >= want[0] && v1 >= want[1] && v2 >= want[2] && v3 >= want[3]) break;
      }
    }
    __syncthreads();
  } else if (V >= 7 && V <= 9) {   // two-level tree: G group counters on separate 128-byte lines -> root -> G release flags
    constexpr unsigned G = V == 7 ? 4u : V == 8 ? 8u : 16u;
    __syncthreads();
    if (threadIdx.x == 0) {
      const unsigned g = blockIdx.x % G, members = (gridDim.x + G - 1 - g) / G;
      unsigned old;
      asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], 1;" : "=r"(old) : "l"(ctr + 32 * (1 + g)) : "memory");
      if (old + 1 == epoch * members) {   // last of the group
        asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], 1;" : "=r"(old) : "l"(ctr) : "memory");
        if (old + 1 == epoch * G) {       // last group: release everybody
          for (unsigned q = 0; q < G; ++q) asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(flags + 32 * q), "r"(epoch) : "memory");
        }
      }
      unsigned v;
      do {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(flags + 32 * g) : "memory");
      } while (v < epoch);
    }
    __syncthreads();
  } else if (V == 10) {   // one counter, but the release is a separate flag line written by the last arriver
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned old;
      asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], 1;" : "=r"(old) : "l"(ctr) : "memory");
      if (old + 1 == epoch * gridDim.x) asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(flags), "r"(epoch) : "memory");
      unsigned v;
      do {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(flags) : "memory");
      } while (v < epoch);
    }
    __syncthreads();
  } else if (V == 11) {   // like 8, relaxed polling + one fence
    constexpr unsigned G = 8u;
    __syncthreads();
    if (threadIdx.x == 0) {
      const unsigned g = blockIdx.x % G, members = (gridDim.x + G - 1 - g) / G;
      unsigned old;
      asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], 1;" : "=r"(old) : "l"(ctr + 32 * (1 + g)) : "memory");
      if (old + 1 == epoch * members) {
        asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], 1;" : "=r"(old) : "l"(ctr) : "memory");
        if (old + 1 == epoch * G) {
          for (unsigned q = 0; q < G; ++q) asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(flags + 32 * q), "r"(epoch) : "memory");
        }
      }
      unsigned v;
      do {
        asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(flags + 32 * g) : "memory");
      } while (v < epoch);
      __threadfence();
    }
    __syncthreads();
  } else if (V == 6) {   // every warp's lane 0 polls (no second __syncthreads)
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      atomicAdd(ctr + (blockIdx.x % 4), 1u);
    }
    if ((threadIdx.x & 31) == 0) {
      unsigned want[4];
      for (int w = 0; w < 4; ++w) want[w] = epoch * ((gridDim.x + 3 - w) / 4);
      for (;;) {
        unsigned v0, v1, v2, v3;
        asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v0), "=r"(v1), "=r"(v2), "=r"(v3) : "l"(ctr) : "memory");
        if (v0 >= want[0] && v1 >= want[1] && v2 >= want[2] && v3 >= want[3]) break;
      }
      __threadfence();
    }
    __syncwarp();
  }
}

template <int V>
__global__ void k_bar(unsigned *ctr, unsigned *flags, double *vec, unsigned iters, unsigned *err) {
  for (unsigned it = 1; it <= iters; ++it) {
    if (threadIdx.x == 0) vec[blockIdx.x * 16] = (double)(it * 1000u + blockIdx.x);
    barrier<V>(ctr, flags, it);
    if ((threadIdx.x & 31) == 0) {
      const unsigned nb = (blockIdx.x + 1 + (threadIdx.x >> 5)) % gridDim.x;
      const double v = __ldcg(vec + nb * 16);
      // the neighbour may already be one iteration ahead (it passed the barrier and stored again)
      if (v != (double)(it * 1000u + nb) && v != (double)((it + 1) * 1000u + nb)) atomicAdd(err, 1u);
    }
  }
}

template <int V>
void run(const char *name, int threads, int grid) {
  unsigned *ctr, *flags, *err;
  double *vec;
  cudaMalloc(&ctr, 128 * 40);
  cudaMalloc(&flags, 4096);
  cudaMalloc(&err, 4);
  cudaMalloc(&vec, 4096 * 16 * 8);
  const unsigned iters = 2000;
  float best = 1e30f;
  unsigned herr = 0;
  for (int rep = 0; rep < 3; ++rep) {
    cudaMemset(ctr, 0, 128 * 40);
    cudaMemset(flags, 0, 4096);
    cudaMemset(err, 0, 4);
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    cudaEventRecord(a);
    if (V == 4) {
      unsigned it = iters;
      void *args[] = {&ctr, &flags, &vec, &it, &err};
      cudaLaunchCooperativeKernel((void *)k_bar<V>, dim3(grid), dim3(threads), args, 0, 0);
    } else {
      k_bar<V><<<grid, threads>>>(ctr, flags, vec, iters, err);
    }
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    if (ms < best) best = ms;
    cudaMemcpy(&herr, err, 4, cudaMemcpyDeviceToHost);
  }
  printf("%-44s threads %4d grid %3d : %.3f us per barrier  (errors %u, %s)\n", name, threads, grid, best * 1000.0f / iters, herr,
         cudaGetErrorString(cudaGetLastError()));
  cudaFree(ctr);
  cudaFree(flags);
  cudaFree(err);
  cudaFree(vec);
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  const int g = p.multiProcessorCount;
  for (int threads : {128, 512}) {
    run<0>("fence + atomicAdd (4 counters) + poll + fence", threads, g);
    run<1>("fence + atomicAdd (1 counter) + poll + fence", threads, g);
    run<2>("red.release + ld.acquire poll", threads, g);
    run<3>("flag per CTA, warp 0 polls all", threads, g);
    run<4>("cooperative groups grid.sync()", threads, g);
    run<5>("atomicAdd + poll, NO fences (signalling only)", threads, g);
    run<6>("4 counters, every warp polls", threads, g);
    run<7>("tree: 4 group lines -> root -> 4 flag lines", threads, g);
    run<8>("tree: 8 group lines -> root -> 8 flag lines", threads, g);
    run<9>("tree: 16 group lines -> root -> 16 flag lines", threads, g);
    run<10>("one counter, separate release flag line", threads, g);
    run<11>("tree 8, relaxed flags + fence", threads, g);
  }
  run<0>("fence + atomicAdd (4 counters), 74 CTAs", 128, g / 2);
  run<0>("fence + atomicAdd (4 counters), 37 CTAs", 128, g / 4);
  return 0;
}
