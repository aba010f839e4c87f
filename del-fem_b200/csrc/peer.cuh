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
// peer in `wait_mask` has posted `seq`
struct DfbHaloWait {
  const double *xg;                      // landing base pre-offset by -n_own*N (so xg + col*N addresses ghost `col`); nullptr: off
  uint32_t n_own;
  uint32_t wait_mask;                    // bit r: wait for halo_flag[r]
  unsigned long long seq;
  const unsigned long long *halo_flag;   // LOCAL mailbox flags
  uint32_t tile_lo_end, tile_hi_begin;   // tiles [0, tile_lo_end) and [tile_hi_begin, ...) hold boundary rows: processed last
};

// called by one lane; returns false on timeout
__device__ __forceinline__ bool halo_wait(const DfbHaloWait &hw) {
  const long long t0 = clock64();
  for (int r = 0; r < DFB_MAX_RANKS; ++r) {
    if (!((hw.wait_mask >> r) & 1u)) continue;
    for (;;) {
      unsigned long long v;
      asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(hw.halo_flag + r) : "memory");
      if (v >= hw.seq) break;
      if (clock64() - t0 > 4000000000ll) return false;
    }
  }
  __threadfence_system();
  return true;
}

struct DfbPeerView {
  int rank;
  int nranks;                      // <= 1: peer path disabled
  unsigned long long seq;          // sequence number of THIS exchange (0: none)
  DfbMailbox *mbox[DFB_MAX_RANKS];  // mbox[rank] is local
};

// called by ONE thread: publish K partial sums of this rank to every rank (including itself)
template <int K>
__device__ __forceinline__ void peer_post(const DfbPeerView &pv, const double *vals) {
  const int slot = (int)(pv.seq % DFB_MBOX_SLOTS);
  // all values to all peers (fire-and-forget NVLink stores), ONE system fence, then the flags
  for (int r = 0; r < pv.nranks; ++r) {
    volatile double *dst = pv.mbox[r]->val[slot][pv.rank];
#pragma unroll
    for (int k = 0; k < K; ++k) dst[k] = vals[k];
  }
  __threadfence_system();
  for (int r = 0; r < pv.nranks; ++r) {
    volatile unsigned long long *f = &pv.mbox[r]->flag[slot][pv.rank];
    *f = pv.seq;
  }
}

// parallel form: called by ALL threads of the (last) block after thread 0 left the K totals in shared memory `vals`;
// thread r serves peer r, so the NVLink round trips of the R peers overlap instead of queueing behind one thread
template <int K>
__device__ __forceinline__ void peer_post_block(const DfbPeerView &pv, const double *vals) {
  __syncthreads();
  const int r = (int)threadIdx.x;
  if (r < pv.nranks) {
    const int slot = (int)(pv.seq % DFB_MBOX_SLOTS);
    volatile double *dst = pv.mbox[r]->val[slot][pv.rank];
#pragma unroll
    for (int k = 0; k < K; ++k) dst[k] = vals[k];
    __threadfence_system();
    volatile unsigned long long *f = &pv.mbox[r]->flag[slot][pv.rank];
    *f = pv.seq;
  }
}

// called by ALL threads of a block (blockDim.x >= nranks): wait for the R partials of exchange pv.seq in the local mailbox
// and return their sum in rank order (identical on every rank). *err is set to 2 on timeout.
template <int K>
__device__ __forceinline__ void peer_gather(const DfbPeerView &pv, double (&out)[K], int *err) {
  const int slot = (int)(pv.seq % DFB_MBOX_SLOTS);
  DfbMailbox *mb = pv.mbox[pv.rank];
  if ((int)threadIdx.x < pv.nranks) {
    const unsigned long long *f = &mb->flag[slot][threadIdx.x];
    unsigned long long v = 0;
    const long long t0 = clock64();
    for (;;) {
      asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(f) : "memory");  // local HBM: L2 is the coherence point
      if (v == pv.seq) break;
      if (clock64() - t0 > 4000000000ll) {
        *err = 2;
        break;
      }
    }
    __threadfence_system();  // acquire: the values below were written before the flag
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < K; ++k) out[k] = 0.0;
  for (int r = 0; r < pv.nranks; ++r) {
#pragma unroll
    for (int k = 0; k < K; ++k) {
      double v;
      asm volatile("ld.volatile.global.f64 %0, [%1];" : "=d"(v) : "l"(&mb->val[slot][r][k]) : "memory");
      out[k] += v;
    }
  }
}
