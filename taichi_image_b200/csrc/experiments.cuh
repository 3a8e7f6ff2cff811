// EXPERIMENTS (not part of the shipped build): K_CORE row rings in shared memory for the pair-engine sweep.
//   -DISP_S2_RING=1   ring filled by per-lane cp.async (LDGSTS), round 1
//   -DISP_S2_RING=2   ring filled by TMA bulk copies (cp.async.bulk + mbarrier, one elected lane), round 2
// Build a variant with  B200ISP_NVCC_EXTRA="-DISP_S2_RING=2" python build.py --out variants/libb200isp_tma.so  and select
// it at run time with B200ISP_LIB=<path>.  Results: DESIGN.md 6c, profiles/r01_ring_experiment.txt, profiles/r02_tma_ring.txt.
#pragma once

namespace isp {

// ---------------------------------------------------------------- K_CORE with a shared-memory row ring (cp.async)
// The register-fetch loop above issues a row's global loads one step (two rows) before it decodes them; at 80 % of
// HBM peak the loaded DRAM latency is longer than that and the decode's first shift is the kernel's top stall
// (profiles/r01_stream2_kernel_ncu.txt: long scoreboard 2.2 cycles per issue).  Loaders that define kRing let the
// hot kind stage its rows through a per-warp ring of kRingStages row PAIRS filled by cp.async (LDGSTS: no
// registers are held while the data is in flight), so a row pair is requested four steps before it is decoded and
// the two raw-row register sets of the fetch-ahead disappear.
// MEASURED (cfg2, profiles/r01_ring_experiment.txt): the ring does remove the memory stall (long scoreboard 2.2 ->
// 0.2 cycles per issue) but the kernel then becomes instruction-fetch bound with the 3x-unrolled 22 KB loop body
// (no_instruction 0.45 -> 2.5; 229 us), and with a single-step 8 KB body the register moves of the sliding window
// cost as much as the stall saved (35.4 instructions per pixel; 185 us) -- both slower than the register-fetch loop
// (172 us).  Kept as a build option (ISP_S2_RING=1, ISP_S2_RING_UNROLL=1|3) for the next round: it needs a window
// rotation that costs neither code size nor moves.  Slot layout of one row: slot word 3 + s = packed
// word (3 * tcol0 - 1 + s), s = 0..97 (halo word, 96 own words, halo word), so the 96 own words start 16-byte
// aligned and are copied as 24 x 16 bytes (when the frame pitch is 16-byte aligned; 4-byte copies otherwise);
// lane t reads slot words 3t+3 .. 3t+7.
constexpr int kRingStages = 4;
constexpr int kRowSlotWords = 104;        // 101 used, 416 bytes
// per-warp ring: kRingStages row pairs + (TMA form) one 8-byte mbarrier per stage
constexpr int kRingWarpWords = kRingStages * 2 * kRowSlotWords + 2 * kRingStages;

// ---- TMA form (ISP_S2_RING == 2): one bulk copy per row, completion counted in bytes on the stage's mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, int bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "W_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@!p bra W_%=;\n\t}"
      ::"r"(a), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t* smem_dst, const void* gmem_src, int bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src), "r"(bytes),
                 "r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}

__device__ __forceinline__ void cp_async16(uint32_t* smem_dst, const uint32_t* gmem_src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}

__device__ __forceinline__ void cp_async4(uint32_t* smem_dst, const uint32_t* gmem_src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <bool BROW0, bool GFIRST0, class Loader, class Epi>
__device__ __forceinline__ void stream2_rows_ring(const Loader& ld, const Epi& epi, typename Epi::State& st, const StreamGeom& g,
                                                  int frame, int tcol, int r0, int rend, uint32_t* ring /* this warp's */) {
  constexpr int D = kRingStages;
  constexpr uint32_t kMask = Loader::kRowMask;
  const int lane = threadIdx.x & 31;
  typename Loader::Cursor cur;
  ld.template open<K_CORE>(cur, frame, tcol, g);
  const ptrdiff_t pitch = ld.pitch();
  const uint32_t* fbase = cur.p - 3 * lane - 1;                  // word 3 * tcol0 - 1 of row 0
  const bool vec16 = ((pitch & 3) == 0) && ((reinterpret_cast<uintptr_t>(fbase + 1) & 15u) == 0);   // kernel-uniform

#if ISP_S2_RING == 2
  // ---- TMA form: lane 0 arms the stage's mbarrier with the byte count and issues ONE bulk copy per row (416 bytes from
  // the 16-byte chunk in front of the strip); the lanes wait on the barrier's phase parity and read their five words.
  if (!vec16) {                                                   // bulk copies need 16-byte aligned rows: register-fetch loop
    stream2_rows<BROW0, GFIRST0, K_CORE, false>(ld, epi, st, g, frame, tcol, r0, rend);
    return;
  }
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring + D * 2 * kRowSlotWords);
  if (lane == 0) {
#pragma unroll
    for (int q = 0; q < D; ++q) mbar_init(&bars[q], 1);
  }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncwarp();
  auto issue = [&](int q) {
    const int rr = r0 - 2 + 2 * q;
    if (rr >= rend + 2) return;                                   // past this task's halo: never waited on
    if (lane == 0) {
      uint64_t* bar = &bars[q & (D - 1)];
      const uint32_t* src = fbase - 3 + (ptrdiff_t)rr * pitch;
      uint32_t* dst = ring + (q & (D - 1)) * (2 * kRowSlotWords);
      mbar_expect_tx(bar, 2 * 4 * kRowSlotWords);
      bulk_g2s(dst, src, 4 * kRowSlotWords, bar);
      bulk_g2s(dst + kRowSlotWords, src + pitch, 4 * kRowSlotWords, bar);
    }
  };
  auto take = [&](int q, typename Loader::Raw& a, typename Loader::Raw& b) {
    mbar_wait(&bars[q & (D - 1)], (unsigned)(q / D) & 1u);        // pair q has landed
    const uint32_t* s = ring + (q & (D - 1)) * (2 * kRowSlotWords) + 3 + 3 * lane;
#pragma unroll
    for (int i = 0; i < 5; ++i) { a.w[i] = s[i]; b.w[i] = s[kRowSlotWords + i]; }
    __syncwarp();                                                 // every lane has read the slot: refill it
    issue(q + D);
  };
#else
  // pair q = rows (r0 - 2 + 2q, r0 - 1 + 2q); pairs past the image are clamped to the last one (never decoded)
  auto issue = [&](int q) {
    const int rr = r0 - 2 + 2 * q;
    if (rr >= rend + 2) { cp_async_commit(); return; }           // past this task's halo: an empty group keeps the count
    const uint32_t* src = fbase + (ptrdiff_t)rr * pitch;
    uint32_t* dst = ring + (q & (D - 1)) * (2 * kRowSlotWords);
    if (vec16) {
#pragma unroll
      for (int rw = 0; rw < 2; ++rw) {
        if (lane < 24) cp_async16(dst + rw * kRowSlotWords + 4 + 4 * lane, src + rw * pitch + 1 + 4 * lane);
        else if (lane < 26) cp_async4(dst + rw * kRowSlotWords + (lane == 24 ? 3 : 100), src + rw * pitch + (lane == 24 ? 0 : 97));
      }
    } else {
#pragma unroll
      for (int rw = 0; rw < 2; ++rw) {
#pragma unroll
        for (int i = 0; i < 3; ++i) cp_async4(dst + rw * kRowSlotWords + 3 + lane + 32 * i, src + rw * pitch + lane + 32 * i);
        if (lane < 2) cp_async4(dst + rw * kRowSlotWords + 99 + lane, src + rw * pitch + 96 + lane);
      }
    }
    cp_async_commit();
  };
  auto take = [&](int q, typename Loader::Raw& a, typename Loader::Raw& b) {
    cp_async_wait<D - 1>();                                       // pair q has landed (at most D-1 younger groups pending)
    __syncwarp();
    const uint32_t* s = ring + (q & (D - 1)) * (2 * kRowSlotWords) + 3 + 3 * lane;
#pragma unroll
    for (int i = 0; i < 5; ++i) { a.w[i] = s[i]; b.w[i] = s[kRowSlotWords + i]; }
    __syncwarp();                                                 // every lane has read the slot: refill it
    issue(q + D);
  };
#endif

  f2 W[6][8];
  typename Loader::Raw ra, rb;
#pragma unroll
  for (int q = 0; q < D; ++q) issue(q);
  take(0, ra, rb);
  ld.template decode<K_CORE>(cur, ra, kMask, W[0]);
  ld.template decode<K_CORE>(cur, rb, kMask, W[1]);
  take(1, ra, rb);
  ld.template decode<K_CORE>(cur, ra, kMask, W[2]);
  ld.template decode<K_CORE>(cur, rb, kMask, W[3]);

#define ISP_RSTEP(U, ROW)                                                                                       \
  {                                                                                                             \
    const int row_ = (ROW);                                                                                     \
    take(((row_ - r0) >> 1) + 2, ra, rb);                          /* rows row+2, row+3 */                      \
    ld.template decode<K_CORE>(cur, ra, kMask, W[(2 * (U) + 4) % 6]);                                           \
    f2 R_[4], G_[4], B_[4];                                                                                     \
    malvar_row2<BROW0, GFIRST0>(W[(2 * (U)) % 6], W[(2 * (U) + 1) % 6], W[(2 * (U) + 2) % 6],                   \
                                W[(2 * (U) + 3) % 6], W[(2 * (U) + 4) % 6], R_, G_, B_);                        \
    epi.template emit<BROW0, GFIRST0, K_CORE>(st, row_, R_, G_, B_);                                            \
    ld.template decode<K_CORE>(cur, rb, kMask, W[(2 * (U) + 5) % 6]);                                           \
    malvar_row2<!BROW0, !GFIRST0>(W[(2 * (U) + 1) % 6], W[(2 * (U) + 2) % 6], W[(2 * (U) + 3) % 6],             \
                                  W[(2 * (U) + 4) % 6], W[(2 * (U) + 5) % 6], R_, G_, B_);                      \
    epi.template emit<!BROW0, !GFIRST0, K_CORE>(st, row_ + 1, R_, G_, B_);                                      \
  }

#ifndef ISP_S2_RING_UNROLL
#define ISP_S2_RING_UNROLL 1
#endif
#if ISP_S2_RING_UNROLL == 3
#pragma unroll 1
  for (int row = r0; row < rend; row += 6) {
    ISP_RSTEP(0, row);
    if (row + 2 >= rend) break;
    ISP_RSTEP(1, row + 2);
    if (row + 4 >= rend) break;
    ISP_RSTEP(2, row + 4);
  }
#else
  // one copy of the step (7 KB of SASS instead of 22 KB: with the memory stall gone the 3x-unrolled body was
  // instruction-fetch bound, no_instruction 2.5 cycles per issue); the window slides by register moves
#pragma unroll 1
  for (int row = r0; row < rend; row += 2) {
    ISP_RSTEP(0, row);
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
      for (int i = 0; i < 8; ++i) W[k][i] = W[k + 2][i];
  }
#endif
#undef ISP_RSTEP
#if ISP_S2_RING != 2
  cp_async_wait<0>();                                             // nothing of this task may land after the warp moves on
#endif
}


}  // namespace isp
