// Camera32 Reinhard with ONE demosaic sweep (the metric's own path: packed12 -> Reinhard -> RGB8).
//
// The reference needs the frame-global maximum of the mapped values before it can write a pixel (camera_isp.py:213-218).
// Recomputing the map in a second full sweep costs the decode + Malvar work twice (r02_reinhard_diet.txt: 106
// instructions per pixel over both sweeps, issue-bound at 55-62 %).  Camera16 avoids that with the f16 scratch the
// reference writes anyway; for Camera32 an f16 (or any float16) scratch is not exact and an f32 one costs 12 + 12 B/px.
//
// But without colour correction the demosaiced Camera32 value IS an integer in disguise:  the x16 filter sum of 12-bit
// samples is an integer I, the sweep's value is fl(I * f32(1/4095) / 16) (one rounding of an exact product) and the clamp
// to [0, 1] is a clamp of I to [0, 65 520].  So pass A (the max sweep) also stores clamp(I, 0, 65 535) as THREE u16 PER
// PIXEL -- bit-exact, 6 B/px -- and pass B is an element-wise kernel that rebuilds the f32 RGB from it, maps, normalises
// by the maximum, applies gamma and quantises.  The pixels of the 2-pixel image frame are renormalised by a division
// (border_fix.cuh) and do not have that form: pass A writes their f32 RGB to a small side table (4 W + 4 (H - 4) pixels per
// frame), pass B reads it back.  Requirements: Camera32, color_adapt == 0, no CCM; everything else keeps the two sweeps.
#pragma once
#include "fused_isp.cuh"

namespace isp {

// index of a frame pixel in the side table (inverse of border_coord, pixel_ops.cuh); H, W >= 4
__device__ __forceinline__ int border_index(int row, int col, int H, int W) {
  if (row < 2) return row * W + col;
  if (row >= H - 2) return (row - (H - 4)) * W + col;
  return 4 * W + (row - 2) * 4 + (col < 2 ? col : col - (W - 4));
}

constexpr float kI2Rgb = kInv4095 * 0.0625f;          // rgb = fl(I * f32(1/4095) / 16), see above

struct U16Scratch {
  uint16_t* map[B200ISP_MAX_FRAMES];      // (H, W, 3) u16: clamp(I, 0, 65535)
  float* frame[B200ISP_MAX_FRAMES];       // border_count(H, W) x 3 f32: RGB of the frame pixels
};

template <bool CA0>
struct EpiReinhardMaxU16 {      // pass A
  U16Scratch sc;
  IspConsts k;
  static constexpr int kStageWords = 32 * 12;
  static constexpr bool kSplitEdge = false;
  static constexpr bool kCompactLoop = true;
  struct State { ReinhardConsts c; float mx; int edge, tcol; uint16_t* out; float* side; WarpCtx wc; };
  __device__ __forceinline__ void init(State& st, int frame, int tcol, const WarpCtx& wc) const {
    st.c = reinhard_consts(k, frame, false);
    st.mx = 0.f;
    st.edge = edge_bits(tcol, k.W);
    st.tcol = tcol;
    st.wc = wc;
    st.out = sc.map[k.frame0 + frame] + 24 * wc.tcol0;
    st.side = sc.frame[k.frame0 + frame];
  }
  __device__ __forceinline__ bool fast_kinds_ok(const State&) const { return true; }

  template <bool BROW, bool GFIRST, int KIND>
  __device__ __forceinline__ void emit(State& st, int row, const f2 (&R)[4], const f2 (&G)[4], const f2 (&B)[4]) const {
    using SS = SiteScale2<BROW, GFIRST>;
    float mx = st.mx;
    if constexpr (KIND == K_GENERAL) {
      // border rows: renormalised values -> side table; the u16 map of these rows is never read
      Vals24 x;
      raw_with_frame<false, BROW, GFIRST, KIND>(R, G, B, row, k.H, st.edge, k.kbase, x);
      const bool live = st.wc.lane < st.wc.nvalid;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        float rgb[3], p[3];
        raw_to_rgb<false>(k, &x.v[3 * q], rgb);
        reinhard_p<false, CA0>(st.c, rgb, p);
        mx = fmaxf(mx, fmaxf(p[0], fmaxf(p[1], p[2])));
        if (live) {
          float* d = st.side + 3 * (size_t)border_index(row, 8 * st.tcol + q, k.H, k.W);
          d[0] = rgb[0]; d[1] = rgb[1]; d[2] = rgb[2];
        }
      }
    } else {
      // the exact integers: I = (S * scale - 16) * 4096, clamped to u16 by the saturating pack
      constexpr float magic = 12582912.f - 65536.f;
      uint32_t w[12];
      {
        int n[24];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const f2 q[3] = {fma2_rz(R[j], bc(SS::r(j) * 4096.f), bc(magic)), fma2_rz(G[j], bc(SS::g(j) * 4096.f), bc(magic)),
                           fma2_rz(B[j], bc(SS::b(j) * 4096.f), bc(magic))};
#pragma unroll
          for (int ch = 0; ch < 3; ++ch) {
            float a, b;
            upk(q[ch], a, b);
            n[3 * j + ch] = __float_as_int(a) - 0x4B400000;
            n[3 * (j + 4) + ch] = __float_as_int(b) - 0x4B400000;
          }
        }
#pragma unroll
        for (int i = 0; i < 12; ++i) asm("cvt.pack.sat.u16.s32 %0, %1, %2;" : "=r"(w[i]) : "r"(n[2 * i + 1]), "r"(n[2 * i]));
      }
      warp_store_row<12, false>(st.wc, st.out + (size_t)((unsigned)row * (unsigned)(3 * k.W)), w);
      // the maximum of the map (largest channel; exact fallback when a denominator is negative, see reinhard_pmax2)
      f2 X[4][3];
      pairs_to_raw2<false, BROW, GFIRST>(R, G, B, X);
      if (KIND == K_EDGE && st.edge) {
        patch_cols_pairs<BROW, GFIRST>(X, st.edge, k.kbase);
        if (st.wc.lane < st.wc.nvalid) {              // frame columns of an interior row -> side table
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            if (q >= 2 && q < 6) continue;
            if ((q < 2 && (st.edge & 1)) || (q >= 6 && (st.edge & 2))) {
              float* d = st.side + 3 * (size_t)border_index(row, 8 * st.tcol + q, k.H, k.W);
#pragma unroll
              for (int ch = 0; ch < 3; ++ch) d[ch] = clamp01(q < 4 ? lo_of(X[q & 3][ch]) : hi_of(X[q & 3][ch]));
            }
          }
        }
      }
      float dmin = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        f2 rgb[3];
        raw2_to_rgb2<false>(k, X[j], rgb);
        float lo, hi;
        if constexpr (CA0) {
          upk(reinhard_pmax2(st.c, rgb, dmin), lo, hi);
          mx = fmaxf(mx, fmaxf(lo, hi));
        } else {
          dmin = -1.f;
        }
      }
      if (dmin < 0.f) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          f2 rgb[3];
          raw2_to_rgb2<false>(k, X[j], rgb);
#pragma unroll
          for (int l = 0; l < 2; ++l) {
            const float c3[3] = {l ? hi_of(rgb[0]) : lo_of(rgb[0]), l ? hi_of(rgb[1]) : lo_of(rgb[1]), l ? hi_of(rgb[2]) : lo_of(rgb[2])};
            float p[3];
            reinhard_p<false, CA0>(st.c, c3, p);
            mx = fmaxf(mx, fmaxf(p[0], fmaxf(p[1], p[2])));
          }
        }
      }
    }
    st.mx = mx;
  }
  __device__ __forceinline__ void finish(State& st, int frame, int lane, bool task_ok) const {
    const float m = warp_max(st.mx);
    if (lane == 0 && task_ok && m > 0.f)
      atomicMax(reinterpret_cast<unsigned int*>(&k.ws->frame_max[k.frame0 + frame]), __float_as_uint(m));
  }
};

// pass B: 8 pixels of one row per thread
template <typename OutT, bool CA0, bool GAMMA>
__global__ void __launch_bounds__(128) reinhard_u16_out_kernel(const U16Scratch sc, const FramePtrs fp, const IspConsts k) {
  const int gx = blockIdx.x * blockDim.x + threadIdx.x, row = blockIdx.y, frame = blockIdx.z;
  const int ntcols = k.W >> 3;
  if (gx >= ntcols) return;
  const ReinhardConsts c = reinhard_consts(k, frame, true);
  alignas(16) uint32_t I2[12];
  ld_bytes<48>(sc.map[frame] + ((size_t)row * k.W + 8 * gx) * 3, I2);
  float rgbf[24];
#pragma unroll
  for (int i = 0; i < 12; ++i) {        // u16 -> f32 without I2F: the half-word dropped into the mantissa of 2^23, minus 2^23
    const float lo = __uint_as_float(__byte_perm(I2[i], 0x4B000000u, 0x7610)) - 8388608.f;
    const float hi = __uint_as_float(__byte_perm(I2[i], 0x4B000000u, 0x7632)) - 8388608.f;
    rgbf[2 * i] = fminf(__fmul_rn(lo, kI2Rgb), 1.0f);
    rgbf[2 * i + 1] = fminf(__fmul_rn(hi, kI2Rgb), 1.0f);
  }
  if (row < 2 || row >= k.H - 2 || gx == 0 || gx == ntcols - 1) {           // image frame: the renormalised values of pass A
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int col = 8 * gx + q;
      if (row < 2 || row >= k.H - 2 || col < 2 || col >= k.W - 2) {
        const float* s = sc.frame[frame] + 3 * (size_t)border_index(row, col, k.H, k.W);
        rgbf[3 * q] = s[0]; rgbf[3 * q + 1] = s[1]; rgbf[3 * q + 2] = s[2];
      }
    }
  }
  uint32_t v[24];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float y[2][3];
    if constexpr (CA0) {
      const f2 rgb[3] = {pk(rgbf[3 * j], rgbf[3 * (j + 4)]), pk(rgbf[3 * j + 1], rgbf[3 * (j + 4) + 1]), pk(rgbf[3 * j + 2], rgbf[3 * (j + 4) + 2])};
      f2 n[3], r;
      reinhard_nr2(c, rgb, bc(c.max_out), n, r);          // q = p / max_out straight from the shared reciprocal
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) {
        y[0][ch] = __saturatef(lo_of(n[ch]) * lo_of(r));
        y[1][ch] = __saturatef(hi_of(n[ch]) * hi_of(r));
      }
    } else {
#pragma unroll
      for (int l = 0; l < 2; ++l) {
        const float c3[3] = {rgbf[3 * (j + 4 * l)], rgbf[3 * (j + 4 * l) + 1], rgbf[3 * (j + 4 * l) + 2]};
        float p[3];
        reinhard_p<false, false>(c, c3, p);
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) y[l][ch] = __saturatef(p[ch] * c.out_scale_inv_max);
      }
    }
#pragma unroll
    for (int l = 0; l < 2; ++l)
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) {
        float q = y[l][ch];
        if constexpr (GAMMA) q = fast_pow(q, c.inv_gamma);
        v[3 * (j + 4 * l) + ch] = Quant<OutT>::q(q);
      }
  }
  constexpr int NW = Quant<OutT>::kWords;
  alignas(16) uint32_t w[NW];
  Quant<OutT>::pack(v, w);
  OutT* dst = reinterpret_cast<OutT*>(fp.out[frame]) + (size_t)row * k.orow + 24 * gx;
  st_bytes<4 * NW>(dst, w);
}

// host: frames [0, n_frames); scratch_bytes_per_frame = H W 6 (map) + border_count 12 (side table), 16-byte aligned
inline size_t reinhard_u16_frame_bytes_impl(int H, int W) {
  const size_t a = (size_t)H * W * 6, b = (size_t)border_count(H, W) * 12;
  return ((a + 15) & ~(size_t)15) + ((b + 15) & ~(size_t)15);
}

}  // namespace isp
