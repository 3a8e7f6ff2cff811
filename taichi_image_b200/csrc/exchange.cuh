// Device side of the shared-exposure mailbox exchange (see exchange.cu for the protocol and the host API):
// layout, post and wait as inline functions, so that the metering kernels can run the exchange INSIDE their
// last-block epilogue instead of as separate 1-warp kernels (profiles/r01: 8 launches per shared metering update
// cost the 0.2 ms step 26 us; the fused form needs 2).
#pragma once
#include "common.cuh"

namespace isp {

constexpr int kMaxRanks = 64;
constexpr int kXchgRanks = 32;          // one lane per rank

// layout in 32-bit words, per kind k (0: record 1, 1: record 2) and parity p:
//   data [k][p][world][8]      flags [k][p][world]      then  seq[2] (this rank's own counters), err
struct MailboxLayout {
  int world;
  __host__ __device__ int data(int k, int p, int r) const { return ((k * 2 + p) * world + r) * 8; }
  __host__ __device__ int flag(int k, int p, int r) const { return 4 * world * 8 + (k * 2 + p) * world + r; }
  __host__ __device__ int seq(int k) const { return 4 * world * 8 + 4 * world + k; }
  __host__ __device__ int err() const { return 4 * world * 8 + 4 * world + 2; }
  __host__ __device__ int words() const { return 4 * world * 8 + 4 * world + 3; }
};

struct PeerPtrs { uint32_t* p[kMaxRanks]; };

// the ranks' mailboxes as a kernel argument of the metering kernels; world == 0: no exchange
struct PeerXchg {
  uint32_t* p[kXchgRanks];
  int world, rank;
};

// ONE warp (all 32 lanes converged): lane r < world writes this rank's record of `nf` floats into rank r's mailbox,
// fences system-wide and publishes the sequence number next to it; returns the sequence number used.
__device__ __forceinline__ uint32_t mailbox_post_warp(uint32_t* const* peers, int world, int rank, int kind, const float* rec, int nf) {
  const MailboxLayout L{world};
  const int lane = threadIdx.x & 31;
  uint32_t* mine = peers[rank];
  // this rank's sequence number of this kind: advanced here, read back by the wait that follows in program / stream order
  const uint32_t seq = mine[L.seq(kind)] + 1u;
  const int par = (int)(seq & 1u);
  if (lane < world) {
    uint32_t* dst = peers[lane];
    for (int i = 0; i < nf; ++i) dst[L.data(kind, par, rank) + i] = __float_as_uint(rec[i]);
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(dst + L.flag(kind, par, rank)), "r"(seq) : "memory");
  }
  __syncwarp();
  if (lane == 0) mine[L.seq(kind)] = seq;
  __syncwarp();
  return seq;
}

// ONE warp: lane r < world spins (acquire, bounded ~2 s) until rank r's record with sequence number `seq` has arrived in
// this rank's own mailbox and copies it to gathered[r * nf ...] (global or shared memory).  A rank that never posts is
// delivered as NaN and the mailbox's error word is set (PeerExchange.check()).
__device__ __forceinline__ void mailbox_wait_warp(uint32_t* mine, int world, int kind, uint32_t seq, float* gathered, int nf) {
  const MailboxLayout L{world};
  const int lane = threadIdx.x & 31;
  const int par = (int)(seq & 1u);
  if (lane < world) {
    const uint32_t* f = mine + L.flag(kind, par, lane);
    uint32_t v = 0;
    long long spins = 0;
    bool late = false;
    while (true) {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
      if ((int)(v - seq) >= 0) break;
      // a probe = one system-scope load of local memory (~1 us) + the sleep: 2^21 probes are roughly 2 s
      if (++spins > (1LL << 21)) { atomicExch(mine + L.err(), 1u); late = true; break; }
      __nanosleep(128);
    }
    for (int i = 0; i < nf; ++i)
      gathered[lane * nf + i] = late ? __int_as_float(0x7fc00000) : __uint_as_float(__ldcg(mine + L.data(kind, par, lane) + i));
  }
  __syncwarp();
}

}  // namespace isp
