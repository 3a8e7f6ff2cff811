// Register-window streaming engine for the Malvar-He-Cutler demosaic (reference: bayer.py:114-177).
//
// B200 mapping: one thread owns 8 consecutive pixel columns and walks DOWN the image keeping a
// 6-row window of the CFA in registers (8 own columns + 2 halo columns per side), so
// every CFA sample is fetched from memory once per row-chunk and decoded once; no shared memory, no
// block barriers.  A warp covers a 256-pixel strip (coalesced 384-byte packed12 rows / 256..1024
// byte typed rows), a "task" is (frame, row-chunk, strip) and tasks are laid out so that the warps
// of a block sit on adjacent strips of the same rows (halo words hit L1).  The grid has one warp
// per task; rows_per_task sizes it to several waves over the 148 SMs.  Rows are fetched one step
// (two rows) ahead of their use, so the global-load latency is covered by a whole step of math.
//
// The 13-tap filters are evaluated through shared partial sums:
//   NS(c) = v[r-1][c]+v[r+1][c]   EW = v[r][c-1]+v[r][c+1]   NNSS = v[r-2][c]+v[r+2][c]
//   EEWW = v[r][c-2]+v[r][c+2]    D = NS(c-1)+NS(c+1)
//   R/B site:  own = 16C   G = 8C+4(NS+EW)-2(NNSS+EEWW)   opposite = 12C+4D-3(NNSS+EEWW)
//   G site:    G = 16C     colour with horizontal neighbours = 10C+8EW-2D-2EEWW+NNSS
//                          colour with vertical neighbours   = 10C+8NS-2D-2NNSS+EEWW
// which are the tables of bayer.py:30-55 (x16).  To save multiplies the engine hands the epilogue
// SCALED sums (own/16, G/2, opposite/4, G-site colours/2; see SiteScale) -- the epilogue folds the
// power-of-two factor into the normalisation constant it multiplies with anyway.  The sums are exact
// for integer-valued inputs (|sum| < 2^24), so the evaluation order does not matter there.
// Out-of-image taps read as 0 and the streaming kernel normalises every pixel by 16; the 2-pixel
// image frame, where the reference renormalises by the in-bounds weight sum (bayer.py:145-151), is
// then rewritten by a per-pixel border kernel (pixel_ops.cuh) launched right after on the same stream.
#pragma once
#include "common.cuh"


namespace isp {

struct StreamGeom {
  int H, W;
  int ntcols;          // ceil(W / 8): thread columns per row
  int warps_per_row;   // ceil(ntcols / 32)
  int rows_per_task;   // even
  int nchunks;         // ceil(H / rows_per_task)
  int nframes;
  long long total_tasks;   // nframes * nchunks * warps_per_row (warp tasks)
};

inline StreamGeom make_geom(int H, int W, int nframes, int rows_per_task) {
  StreamGeom g;
  g.H = H; g.W = W;
  g.ntcols = (W + 7) / 8;
  g.warps_per_row = (g.ntcols + 31) / 32;
  if (rows_per_task <= 0) {
    // aim at >= ~6 waves of 16 warps on 148 SMs, keep the 4 halo rows <= ~12 % of the loads
    rows_per_task = 48;
    while (rows_per_task > 12 &&
           (long long)nframes * ((H + rows_per_task - 1) / rows_per_task) * g.warps_per_row < 6LL * 16 * kNumSMs)
      rows_per_task -= 12;
  }
  rows_per_task += rows_per_task & 1;
  g.rows_per_task = rows_per_task;
  g.nchunks = (H + rows_per_task - 1) / rows_per_task;
  g.nframes = nframes;
  g.total_tasks = (long long)nframes * g.nchunks * g.warps_per_row;
  return g;
}

// ---------------------------------------------------------------- per-site filter formulas (scaled)
// R/B site: own colour = C (x16), G = 4C + 2(NS+EW) - (NNSS+EEWW) (x2), opposite = 3C + D - 0.75(NNSS+EEWW) (x4)
__device__ __forceinline__ void malvar_csite(float C, float NS, float EW, float NNSS, float EEWW, float D,
                                             float& g2, float& opp4) {
  const float A = NS + EW, Bq = NNSS + EEWW;
  g2 = fmaf(4.f, C, fmaf(2.f, A, -Bq));
  opp4 = fmaf(-0.75f, Bq, fmaf(3.f, C, D));
}
// G site: G = C (x16), horizontal-neighbour colour = 5C - D + 4EW - EEWW + 0.5NNSS (x2), vertical likewise
__device__ __forceinline__ void malvar_gsite(float C, float NS, float EW, float NNSS, float EEWW, float D,
                                             float& h2, float& v2) {
  const float T = fmaf(5.f, C, -D);
  h2 = fmaf(0.5f, NNSS, fmaf(4.f, EW, T) - EEWW);
  v2 = fmaf(0.5f, EEWW, fmaf(4.f, NS, T) - NNSS);
}

// scale that turns the engine's value for (row type, pixel j, channel) back into the x16 filter sum
template <bool BROW, bool GFIRST>
struct SiteScale {
  static __host__ __device__ constexpr bool gsite(int j) { return ((j & 1) == 0) == GFIRST; }
  static __host__ __device__ constexpr float r(int j) { return gsite(j) ? 2.f : (BROW ? 4.f : 16.f); }
  static __host__ __device__ constexpr float g(int j) { return gsite(j) ? 16.f : 2.f; }
  static __host__ __device__ constexpr float b(int j) { return gsite(j) ? 2.f : (BROW ? 16.f : 4.f); }
};

// One output row of 8 pixels.  z / p1 / p2 are full 12-wide rows (column c0-2+i at index i); the two rows
// above are passed as pointers with a compile-time column offset because the engine only carries the
// columns that are still needed of them: m2[(j+2) + M2OFF] is column j+2 (NN tap), m1[(k+1) + M1OFF]
// column k+1 (N / NW / NE taps).
template <bool BROW, bool GFIRST, int M2OFF, int M1OFF>
__device__ __forceinline__ void malvar_row(const float* __restrict__ m2, const float* __restrict__ m1,
                                           const float (&z)[12], const float (&p1)[12], const float (&p2)[12],
                                           float (&R)[8], float (&G)[8], float (&B)[8]) {
  float ns[10];
#pragma unroll
  for (int k = 0; k < 10; ++k) ns[k] = m1[k + 1 + M1OFF] + p1[k + 1];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float C = z[j + 2];
    const float EW = z[j + 1] + z[j + 3];
    const float EEWW = z[j] + z[j + 4];
    const float NS = ns[j + 1];
    const float D = ns[j] + ns[j + 2];
    const float NNSS = m2[j + 2 + M2OFF] + p2[j + 2];
    if (!SiteScale<BROW, GFIRST>::gsite(j)) {
      float g2, opp4;
      malvar_csite(C, NS, EW, NNSS, EEWW, D, g2, opp4);
      G[j] = g2;
      R[j] = BROW ? opp4 : C;
      B[j] = BROW ? C : opp4;
    } else {
      float h2, v2;
      malvar_gsite(C, NS, EW, NNSS, EEWW, D, h2, v2);
      G[j] = C;
      R[j] = BROW ? v2 : h2;
      B[j] = BROW ? h2 : v2;
    }
  }
}

// One step = two output rows (row, row+1).  cA/cB are rows row / row+1, nA/nB receive rows row+2 / row+3
// (fetched one step earlier), oA / oB hold what is still needed of rows row-2 (columns 2..9) and row-1
// (columns 1..10).  The rows for the NEXT step are requested before the math so their latency is covered.
template <bool BROW0, bool GFIRST0, class Loader, class Epi>
__device__ __forceinline__ void stream_step(const Loader& ld, const Epi& epi, typename Epi::State& st,
                                            const typename Loader::Cursor& cur, const StreamGeom& g, int row, int rend,
                                            typename Loader::Raw& raw0, typename Loader::Raw& raw1,
                                            float (&oA)[8], float (&oB)[10], const float (&cA)[12], const float (&cB)[12],
                                            float (&nA)[12], float (&nB)[12]) {
  ld.decode(raw0, nA);
  ld.decode(raw1, nB);
  // rows past the chunk halo or the image come back as zeros and are never used
  ld.fetch(cur, row + 4 < rend + 2 ? row + 4 : -1, g, raw0);
  ld.fetch(cur, row + 5 < rend + 2 ? row + 5 : -1, g, raw1);
  ld.prefetch(cur, row + 8, g);
  ld.prefetch(cur, row + 9, g);
  float R[8], G[8], B[8];
  malvar_row<BROW0, GFIRST0, -2, -1>(oA, oB, cA, cB, nA, R, G, B);
  epi.template emit<BROW0, GFIRST0>(st, row, R, G, B);
  malvar_row<!BROW0, !GFIRST0, -1, 0>(oB, cA, cB, nA, nB, R, G, B);
  epi.template emit<!BROW0, !GFIRST0>(st, row + 1, R, G, B);
#pragma unroll
  for (int j = 0; j < 8; ++j) oA[j] = cA[j + 2];
#pragma unroll
  for (int j = 0; j < 10; ++j) oB[j] = cB[j + 1];
}

// ---------------------------------------------------------------- warp-cooperative row store
// A lane owns 8 pixels x 3 channels = NW 32-bit words of an output row, i.e. a 4*NW-byte segment with a
// 4*NW-byte lane stride: stored straight from registers, every STG.128 would touch 32 scattered 16-byte
// pieces (half sectors) of twelve 128-byte lines.  Measured on B200 with the sweep's 1.5 B : 6 B traffic
// mix (profiles/r01_membench.txt) that shape tops out at ~3.9 TB/s, the same bytes written as whole
// contiguous 512-byte runs reach ~5.8 TB/s.  So each warp transposes the row through a private
// shared-memory stage (conflict-free for NW = 6 and 12) and writes it as consecutive 16-byte (8-byte when
// the row pitch is only 8-byte aligned: RGB8) chunks; only __syncwarp, no block barrier.
struct WarpCtx {
  int lane;
  int tcol0;            // first thread column of the warp's strip
  int nvalid;           // lanes of the strip that map to real thread columns (the others compute a clamped copy)
  uint32_t* stage;      // this warp's staging buffer, Epi::kStageWords words
};

// Output rows are written once and never re-read by the ISP: store them with the evict-first policy so that
// they do not push the packed input rows (which the metering pass has just pulled into the 126 MB L2 and the
// sweep is about to read) out of L2.
#ifndef ISP_STREAMING_STORES
#define ISP_STREAMING_STORES 1
#endif
__device__ __forceinline__ void st_out(uint4* p, const uint4& v) {
#if ISP_STREAMING_STORES
  __stcs(p, v);
#else
  *p = v;
#endif
}
__device__ __forceinline__ void st_out(uint2* p, const uint2& v) {
#if ISP_STREAMING_STORES
  __stcs(p, v);
#else
  *p = v;
#endif
}

template <int NW, bool FULL = false>       // FULL: all 32 lanes map to real thread columns (no store predicates)
__device__ __forceinline__ void warp_store_row(const WarpCtx& wc, void* row_dst /* warp's first byte of the row */,
                                               const uint32_t (&w)[NW]) {
  constexpr int CW = (NW % 4 == 0) ? 4 : 2;     // words per chunk
  constexpr int PER_LANE = NW / CW;
  const int nchunks = wc.nvalid * PER_LANE;
  __syncwarp();
  if constexpr (CW == 4) {
    uint4* s = reinterpret_cast<uint4*>(wc.stage);
#pragma unroll
    for (int i = 0; i < PER_LANE; ++i) s[PER_LANE * wc.lane + i] = make_uint4(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
    __syncwarp();
    uint4* d = reinterpret_cast<uint4*>(row_dst);
#pragma unroll
    for (int i = 0; i < PER_LANE; ++i) {
      const int idx = i * 32 + wc.lane;
      if (FULL || idx < nchunks) st_out(d + idx, s[idx]);
    }
  } else {
    uint2* s = reinterpret_cast<uint2*>(wc.stage);
#pragma unroll
    for (int i = 0; i < PER_LANE; ++i) s[PER_LANE * wc.lane + i] = make_uint2(w[2 * i], w[2 * i + 1]);
    __syncwarp();
    uint2* d = reinterpret_cast<uint2*>(row_dst);
#pragma unroll
    for (int i = 0; i < PER_LANE; ++i) {
      const int idx = i * 32 + wc.lane;
      if (FULL || idx < nchunks) st_out(d + idx, s[idx]);
    }
  }
}

// Loader concept:
//   struct Raw;  struct Cursor;
//   void open(Cursor&, int frame, int tcol, const StreamGeom&)           per-task base pointer / edge flags
//   void fetch(const Cursor&, int row, const StreamGeom&, Raw&)          issue the global loads of one row (0 outside)
//   void prefetch(const Cursor&, int row, const StreamGeom&)             optional L2 prefetch of a later row
//   void decode(const Raw&, float (&v)[12])                              v[j] = CFA at column 8*tcol-2+j
// Epilogue concept:
//   static constexpr int kStageWords                                      per-warp staging words (0: stores nothing)
//   struct State;  void init(State&, int frame, int tcol, const WarpCtx&);  void finish(State&, int frame, int lane, bool task_ok)
//   template <bool BROW, bool GFIRST> void emit(State&, int row, R, G, B)   scaled filter sums, see SiteScale
template <int PATTERN, class Loader, class Epi>
__global__ void __launch_bounds__(256, 2) stream_kernel(const Loader ld, const Epi epi, const StreamGeom g) {
  constexpr bool BROW0 = (PATTERN == B200ISP_GBRG || PATTERN == B200ISP_BGGR);
  constexpr bool GFIRST0 = (PATTERN == B200ISP_GRBG || PATTERN == B200ISP_GBRG);
  const int lane = threadIdx.x & 31;
  long long task = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const bool task_ok = task < g.total_tasks;
  if (!task_ok) task = 0;
  const int strip = (int)(task % g.warps_per_row);
  const long long t2 = task / g.warps_per_row;
  const int chunk = (int)(t2 % g.nchunks);
  const int frame = (int)(t2 / g.nchunks);
  const int tcol = min(strip * 32 + lane, g.ntcols - 1);   // lanes past the last column recompute it (never stored)

  __shared__ __align__(16) uint32_t stage[8][Epi::kStageWords > 0 ? Epi::kStageWords : 1];
  WarpCtx wc;
  wc.lane = lane;
  wc.tcol0 = strip * 32;
  wc.nvalid = min(32, g.ntcols - strip * 32);
  wc.stage = stage[threadIdx.x >> 5];

  typename Epi::State st;
  epi.init(st, frame, tcol, wc);

  if (task_ok) {
    const int r0 = chunk * g.rows_per_task;
    const int rend = min(r0 + g.rows_per_task, g.H);
    typename Loader::Cursor cur;
    ld.open(cur, frame, tcol, g);

    // The window: two 2-row buffers that alternate between "centre rows" and "incoming rows" (the loop is
    // unrolled twice so no row is ever copied) plus the 18 values still needed of the two rows above.
    float b0A[12], b0B[12], b1A[12], b1B[12], oA[8], oB[10];
    typename Loader::Raw raw0, raw1;
    {
      float t[12];
      ld.fetch(cur, r0 - 2, g, raw0);
      ld.fetch(cur, r0 - 1, g, raw1);
      ld.decode(raw0, t);
#pragma unroll
      for (int j = 0; j < 8; ++j) oA[j] = t[j + 2];
      ld.decode(raw1, t);
#pragma unroll
      for (int j = 0; j < 10; ++j) oB[j] = t[j + 1];
    }
    ld.fetch(cur, r0, g, raw0);
    ld.fetch(cur, r0 + 1, g, raw1);
    ld.decode(raw0, b0A);
    ld.decode(raw1, b0B);
    ld.fetch(cur, r0 + 2, g, raw0);
    ld.fetch(cur, r0 + 3, g, raw1);
    ld.prefetch(cur, r0 + 4, g); ld.prefetch(cur, r0 + 5, g);
    ld.prefetch(cur, r0 + 6, g); ld.prefetch(cur, r0 + 7, g);

#pragma unroll 1
    for (int row = r0; row < rend; row += 4) {
      stream_step<BROW0, GFIRST0>(ld, epi, st, cur, g, row, rend, raw0, raw1, oA, oB, b0A, b0B, b1A, b1B);
      if (row + 2 < rend)
        stream_step<BROW0, GFIRST0>(ld, epi, st, cur, g, row + 2, rend, raw0, raw1, oA, oB, b1A, b1B, b0A, b0B);
      else
        break;
    }
  }
  epi.finish(st, frame, lane, task_ok);
}

template <int PATTERN, class Loader, class Epi>
inline int launch_stream(const Loader& ld, const Epi& epi, const StreamGeom& g, cudaStream_t s, const char* what) {
  if (g.total_tasks == 0) return B200ISP_OK;
  const long long blocks = (g.total_tasks + 7) / 8;
  stream_kernel<PATTERN, Loader, Epi><<<(unsigned)blocks, 256, 0, s>>>(ld, epi, g);
  return cuda_status(cudaPeekAtLastError(), what);
}

#define ISP_DISPATCH_PATTERN(p, P, ...)                                        \
  switch (p) {                                                                 \
    case B200ISP_RGGB: { constexpr int P = B200ISP_RGGB; __VA_ARGS__; break; } \
    case B200ISP_GRBG: { constexpr int P = B200ISP_GRBG; __VA_ARGS__; break; } \
    case B200ISP_GBRG: { constexpr int P = B200ISP_GBRG; __VA_ARGS__; break; } \
    case B200ISP_BGGR: { constexpr int P = B200ISP_BGGR; __VA_ARGS__; break; } \
    default: isp::set_error("unknown bayer pattern %d", (int)(p)); return B200ISP_E_ARG; \
  }

}  // namespace isp
