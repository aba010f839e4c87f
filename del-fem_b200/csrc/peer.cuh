// peer.cuh — scalar all-reduce fused into the producing / consuming kernels over NVLink peer memory.
// Every rank owns a small "mailbox" in its own HBM, IPC-mapped into all peers. The last block of a reducing kernel writes
// its rank's partial sums straight into every peer's mailbox (plain NVLink stores) and then releases a sequence flag;
// the consuming kernel's blocks acquire the R flags of their LOCAL mailbox and add the R partials in rank order, so every
// rank obtains the bit-identical total without a collective launch (NCCL all-reduce: ~2 x 25 us per PCG iteration).
// Spins are bounded (about 2 s): a lost peer raises an error flag instead of hanging the GPU.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

static const int DFB_MAX_RANKS = 8;
static const int DFB_MBOX_SLOTS = 4;

struct DfbMailbox {
  double val[DFB_MBOX_SLOTS][DFB_MAX_RANKS][4];
  unsigned long long flag[DFB_MBOX_SLOTS][DFB_MAX_RANKS];
  // halo push: peer r stores the sequence number of the ghost exchange it has completed INTO THIS rank's landing zone
  unsigned long long halo_flag[DFB_MAX_RANKS];
  unsigned long long pad[8];
  // the landing zone (2 parity buffers of ctx->peer_halo_bytes each) follows the struct in the same IPC allocation
};
static_assert(sizeof(DfbMailbox) % 128 == 0, "landing zone must start on a 128-byte boundary");

// receiver side of the halo push, handed to the SpMV: ghost columns (>= n_own) are read from the landing zone once every
// peer in `wait_mask` has posted the exchange id the solver state names (hbase + 1 + iteration)
struct DfbHaloWait {
  const double *xg[2];                   // landing bases (parity 0/1) pre-offset by -n_own*N (xg + col*N addresses ghost `col`); xg[0] == nullptr: off
  uint32_t n_own;
  uint32_t wait_mask;                    // bit r: wait for halo_flag[r]
  const unsigned long long *halo_flag;   // LOCAL mailbox flags
  uint32_t tile_lo_end, tile_hi_begin;   // tiles [0, tile_lo_end) and [tile_hi_begin, ...) hold boundary rows: processed last
};

// called by one lane; returns false on timeout
__device__ __forceinline__ bool halo_wait(const DfbHaloWait &hw, unsigned long long seq) {
  const long long t0 = clock64();
  for (int r = 0; r < DFB_MAX_RANKS; ++r) {
    if (!((hw.wait_mask >> r) & 1u)) continue;
    for (;;) {
      unsigned long long v;
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(hw.halo_flag + r) : "memory");
      if (v >= seq) break;
      if (clock64() - t0 > 4000000000ll) return false;
    }
  }
  return true;
}

struct DfbPeerView {
  int rank;
  int nranks;                      // <= 1: peer path disabled
  int post;                        // this launch publishes its reduction to the peers (the exchange id comes from the solver state)
  DfbMailbox *mbox[DFB_MAX_RANKS];  // mbox[rank] is local
};

// called by ALL threads of the (last) block after thread 0 left the K totals in shared memory `vals`: publish this rank's
// partial sums of exchange `seq` to every rank (including itself). Thread r serves peer r, so the NVLink round trips of the
// R peers overlap instead of queueing behind one thread.
template <int K>
__device__ __noinline__ void peer_post_block(const DfbPeerView &pv, unsigned long long seq, const double *vals) {
  __syncthreads();
  const int r = (int)threadIdx.x;
  if (r < pv.nranks) {
    const int slot = (int)(seq % DFB_MBOX_SLOTS);
    volatile double *dst = pv.mbox[r]->val[slot][pv.rank];
#pragma unroll
    for (int k = 0; k < K; ++k) dst[k] = vals[k];
    __threadfence_system();
    volatile unsigned long long *f = &pv.mbox[r]->flag[slot][pv.rank];
    *f = seq;
  }
}

// called by ALL threads of a block (blockDim.x >= nranks): wait for the R partials of exchange `seq` in the local mailbox
// and return their sum in rank order (identical on every rank). *err is set to 2 on timeout.
template <int K>
__device__ __forceinline__ void peer_gather(const DfbPeerView &pv, unsigned long long seq, double (&out)[K], int *err) {
  const int slot = (int)(seq % DFB_MBOX_SLOTS);
  DfbMailbox *mb = pv.mbox[pv.rank];
  if ((int)threadIdx.x < pv.nranks) {
    const unsigned long long *f = &mb->flag[slot][threadIdx.x];
    unsigned long long v = 0;
    const long long t0 = clock64();
    for (;;) {
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(f) : "memory");  // local HBM: L2 is the coherence point
      if (v == seq) break;
      if (clock64() - t0 > 4000000000ll) {
        *err = 2;
        break;
      }
    }
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < K; ++k) out[k] = 0.0;
  for (int r = 0; r < pv.nranks; ++r) {
#pragma unroll
    for (int k = 0; k < K; ++k) {
      double v;
      asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(&mb->val[slot][r][k]) : "memory");
      out[k] += v;
    }
  }
}

// All-reduce consumer folded into a multi-block kernel: the FIRST block to arrive (ticket; it is resident by construction, so
// nobody can wait for a block that was never scheduled) gathers the R partials of exchange `seq` from the local mailbox, stores
// the rank-ordered total in dst[0..K) and releases a local flag; one thread of every other block polls that local flag.
// (Letting every block spin on the mailbox line itself starved the incoming NVLink writes and measured slower than NCCL.)
template <int K>
__device__ __noinline__ void peer_gather_grid(const DfbPeerView &pv, unsigned long long seq, double *dst, unsigned long long *gflag,
                                                 unsigned int *ticket, int *err, double (&out)[K]) {
  __shared__ int s_elected;
  if (threadIdx.x == 0) s_elected = (atomicInc(ticket, gridDim.x - 1) == 0u) ? 1 : 0;
  __syncthreads();
  if (s_elected) {
    peer_gather<K>(pv, seq, out, err);
    if (threadIdx.x == 0) {
#pragma unroll
      for (int k = 0; k < K; ++k) dst[k] = out[k];
      asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(gflag), "l"(seq) : "memory");
    }
  } else {
    if (threadIdx.x == 0) {
      const long long t0 = clock64();
      for (;;) {
        unsigned long long v;
        asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(gflag) : "memory");
        if (v == seq) break;
        if (clock64() - t0 > 4000000000ll) {
          *err = 2;
          break;
        }
      }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < K; ++k) out[k] = __ldcg(dst + k);
  }
}
