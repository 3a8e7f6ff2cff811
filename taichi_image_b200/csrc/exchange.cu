// Peer-memory exchange of the shared-exposure records over NVLink / NVSwitch.
//
// The joint metering of a rig needs two tiny all-gathers per time step (8 and 32 bytes per rank).  Going
// through NCCL costs two host-side collective enqueues per step, which makes a 0.2 ms step host-bound; here the
// exchange is two 1-warp kernels per gather on the metering stream, with no host involvement and capturable in
// a CUDA graph:
//   post : the rank writes its record straight into EVERY rank's mailbox (peer-mapped device memory, plain
//          st.global over NVLink), fences system-wide, then publishes a sequence number next to it;
//   wait : the rank spins (ld.acquire.sys) on the sequence numbers of all ranks in its OWN mailbox and copies
//          the records to a local buffer in rank order -- every rank folds them identically afterwards.
// Mailboxes are double-buffered on the parity of the sequence number: a rank cannot run more than one step ahead
// of the slowest rank (it needs that rank's record to finish its own step), so a slot is never overwritten while
// a peer still reads it.  The spin is bounded (~2 s): on expiry the error word is set, the late rank's record is
// delivered as NaN and the wait returns.
//
// One process per GPU: the mailbox is cudaMalloc'ed by its owner, exported with cudaIpcGetMemHandle and opened by
// the peers (cudaIpcOpenMemHandle enables peer access lazily); the handles travel once through torch.distributed.
#include <string.h>
#include "exchange.cuh"

namespace isp {

__global__ void mailbox_post_kernel(const float* __restrict__ rec, int kind, const PeerPtrs peers, int world, int rank) {
  mailbox_post_warp(peers.p, world, rank, kind, rec, kind == 0 ? B200ISP_REC1 : B200ISP_REC2);
}

__global__ void mailbox_wait_kernel(uint32_t* __restrict__ mine, int kind, int world, float* __restrict__ gathered) {
  const MailboxLayout L{world};
  mailbox_wait_warp(mine, world, kind, mine[L.seq(kind)], gathered, kind == 0 ? B200ISP_REC1 : B200ISP_REC2);
}

}  // namespace isp

using namespace isp;

extern "C" size_t b200isp_mailbox_bytes(int world) {
  if (world < 1 || world > kMaxRanks) return 0;
  return sizeof(uint32_t) * (size_t)MailboxLayout{world}.words();
}

extern "C" int b200isp_mailbox_create(int world, void** mailbox_out, void* ipc_handle_out64) {
  ISP_REQUIRE(world >= 1 && world <= kMaxRanks && mailbox_out && ipc_handle_out64, B200ISP_E_ARG, "mailbox_create: bad argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  void* p = nullptr;
  int st = cuda_status(cudaMalloc(&p, b200isp_mailbox_bytes(world)), "cudaMalloc(mailbox)");
  if (st) return st;
  st = cuda_status(cudaMemset(p, 0, b200isp_mailbox_bytes(world)), "cudaMemset(mailbox)");
  if (st) return st;
  cudaIpcMemHandle_t h;
  st = cuda_status(cudaIpcGetMemHandle(&h, p), "cudaIpcGetMemHandle");
  if (st) return st;
  memcpy(ipc_handle_out64, &h, sizeof(h));
  *mailbox_out = p;
  return B200ISP_OK;
}

extern "C" int b200isp_mailbox_open(const void* ipc_handle64, void** mailbox_out) {
  ISP_REQUIRE(ipc_handle64 && mailbox_out, B200ISP_E_ARG, "mailbox_open: null pointer");
  cudaIpcMemHandle_t h;
  memcpy(&h, ipc_handle64, sizeof(h));
  return cuda_status(cudaIpcOpenMemHandle(mailbox_out, h, cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle");
}

extern "C" int b200isp_mailbox_close(void* mailbox, int is_owner) {
  if (!mailbox) return B200ISP_OK;
  return cuda_status(is_owner ? cudaFree(mailbox) : cudaIpcCloseMemHandle(mailbox), "mailbox_close");
}

static int peer_table(const char* what, void* const* peers_host, int world, int rank, PeerPtrs& pp) {
  ISP_REQUIRE(peers_host, B200ISP_E_ARG, "%s: null mailbox table", what);
  ISP_REQUIRE(world >= 1 && world <= 32 && rank >= 0 && rank < world, B200ISP_E_ARG, "%s: world %d rank %d", what, world, rank);
  for (int r = 0; r < world; ++r) {
    ISP_REQUIRE(peers_host[r], B200ISP_E_ARG, "%s: null mailbox of rank %d", what, r);
    pp.p[r] = (uint32_t*)peers_host[r];
  }
  return B200ISP_OK;
}

// rec: this rank's record on the device (kind 1: 2 floats, kind 2: 8 floats); peers_host: world mailbox pointers
// (index = rank; entry `rank` is the local mailbox).
extern "C" int b200isp_mailbox_post(const float* rec, int kind, void* const* peers_host, int world, int rank,
                                    b200isp_stream stream) {
  ISP_REQUIRE(rec && (kind == 1 || kind == 2), B200ISP_E_ARG, "mailbox_post: bad argument");
  PeerPtrs pp;
  const int st = peer_table("mailbox_post", peers_host, world, rank, pp);
  if (st) return st;
  mailbox_post_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(rec, kind - 1, pp, world, rank);
  ISP_LAUNCH_CHECK("mailbox_post_kernel");
  return B200ISP_OK;
}

// gathered: [world][2 | 8] floats on the device, rank order
extern "C" int b200isp_mailbox_wait(void* mailbox, int kind, int world, float* gathered, b200isp_stream stream) {
  ISP_REQUIRE(mailbox && gathered && (kind == 1 || kind == 2) && world >= 1 && world <= 32, B200ISP_E_ARG, "mailbox_wait: bad argument");
  mailbox_wait_kernel<<<1, 32, 0, (cudaStream_t)stream>>>((uint32_t*)mailbox, kind - 1, world, gathered);
  ISP_LAUNCH_CHECK("mailbox_wait_kernel");
  return B200ISP_OK;
}

extern "C" int b200isp_mailbox_exchange(const float* rec, int kind, void* const* peers_host, int world, int rank,
                                        float* gathered, b200isp_stream stream) {
  const int st = b200isp_mailbox_post(rec, kind, peers_host, world, rank, stream);
  if (st) return st;
  return b200isp_mailbox_wait(peers_host[rank], kind, world, gathered, stream);
}

// 1 if a bounded wait expired on this rank's mailbox (a peer never posted); synchronises the stream
extern "C" int b200isp_mailbox_error(const void* mailbox, int world, b200isp_stream stream) {
  ISP_REQUIRE(mailbox && world >= 1 && world <= kMaxRanks, B200ISP_E_ARG, "mailbox_error: bad argument");
  uint32_t e = 0;
  int st = cuda_status(cudaMemcpyAsync(&e, (const uint32_t*)mailbox + MailboxLayout{world}.err(), sizeof(e), cudaMemcpyDeviceToHost,
                                       (cudaStream_t)stream), "mailbox_error copy");
  if (st) return st;
  st = cuda_status(cudaStreamSynchronize((cudaStream_t)stream), "mailbox_error sync");
  if (st) return st;
  return e ? 1 : 0;
}
