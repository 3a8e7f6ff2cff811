// Fused camera-ISP sweep: packed12 Bayer -> [metering] -> demosaic -> WB/CCM -> tone map -> RGB8/16/f16
// without materialising the CFA or the float RGB image.
//
// Reference path replaced (camera_isp.py): load_packed12 :333-340 (decode12 packed.py:91-131 +
// bayer_to_rgb bayer.py:114-177), update_metering :376-385 (metering_kernel :142-166),
// tonemap_linear :405-413 (tonemap.py:11-17), tonemap_reinhard :394-403 (reinhard_kernel :177-218).
// Rounding points of the ISP dtype are reproduced: Camera16 rounds the CFA,
// the demosaiced RGB and the Reinhard intermediate through f16; Camera32 keeps f32.
//
// Launch plan per call (all on one stream, no host sync):
//   [metering]  meter_phase1 -> meter_phase2   sparse per-pixel sampler straight from the packed bytes
//   linear      stream<EpiLinear>  + border kernel
//   reinhard    stream<EpiReinhardMax> + border   (frame-global max of the mapped values)
//               stream<EpiReinhard>    + border   (second sweep re-reads the packed frame from L2)
//   none        stream<EpiRgb> + border           (load_packed12 only: float RGB out)
#pragma once
#include "stream_engine.cuh"
#include "pixel_ops.cuh"
#include "metering.cuh"
#include "reinhard.cuh"

namespace isp {

constexpr float kInv4095 = (float)(1.0 / 4095.0);      // packed.py:99 with scale 1.0

struct FramePtrs {
  const uint8_t* in[B200ISP_MAX_FRAMES];
  void* out[B200ISP_MAX_FRAMES];
};

// ---------------------------------------------------------------- packed12 row loader (standard layout)
// Thread column tcol owns pixels 8*tcol..8*tcol+7 = bytes 12*tcol..12*tcol+11 = words 3*tcol..3*tcol+2;
// the halo pixels live in the top 3 bytes of word 3*tcol-1 and the low 3 bytes of word 3*tcol+3.
// Camera32: samples are decoded as b = 1 + v/4096 (12 bits dropped into the mantissa of 1.0f: one
// shift + one LOP3 per pixel, no int->float conversion).  All filter weights sum to 16, so the sums
// come out as 16 + S/4096 exactly and the bias folds into the epilogue's FMA.
// Camera16: the reference stores cfa = f16(v * f32(1/4095)); the same value is produced with one FMA
// (exact product, single rounding) and a packed f32->f16->f32 round trip.
// (x & 0x007FF800) | 0x3F800000 in ONE LOP3 (the compiler splits the two immediates into two)
__device__ __forceinline__ float biased_from_shifted(uint32_t shifted, uint32_t mask, uint32_t one) {
  uint32_t r;
  asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(r) : "r"(shifted), "r"(mask), "r"(one));
  return __uint_as_float(r);
}

template <bool CAM16>
struct Packed12Loader {
  FramePtrs fp;
  int pitch_words;       // W * 3 / 8
  int frame0;
  struct Raw { uint32_t w[5]; };
  struct Cursor { const uint32_t* p; bool left, right; };

  __device__ __forceinline__ void open(Cursor& c, int frame, int tcol, const StreamGeom& g) const {
    c.p = reinterpret_cast<const uint32_t*>(fp.in[frame0 + frame]) + 3 * tcol;
    c.left = tcol > 0;
    c.right = tcol < g.ntcols - 1;
  }

  __device__ __forceinline__ void fetch(const Cursor& c, int row, const StreamGeom& g, Raw& raw) const {
    const bool rv = (unsigned)row < (unsigned)g.H;
    const uint32_t* p = c.p + (rv ? (unsigned)row * (unsigned)pitch_words : 0u);
    raw.w[0] = (rv && c.left) ? __ldg(p - 1) : 0u;
    raw.w[1] = rv ? __ldg(p) : 0u;
    raw.w[2] = rv ? __ldg(p + 1) : 0u;
    raw.w[3] = rv ? __ldg(p + 2) : 0u;
    raw.w[4] = (rv && c.right) ? __ldg(p + 3) : 0u;
  }

  // L2 prefetch of a row several steps ahead: the 32 lanes' 12-byte segments form one contiguous 384-byte
  // run, so every 10th lane touching its own address covers all of its 128-byte lines.
  __device__ __forceinline__ void prefetch(const Cursor& c, int row, const StreamGeom& g) const {
    if ((unsigned)row < (unsigned)g.H && (threadIdx.x & 31) % 10 == 0)
      asm volatile("prefetch.global.L2 [%0];" ::"l"(c.p + (unsigned)row * (unsigned)pitch_words));
  }

  __device__ __forceinline__ void decode(const Raw& raw, float (&v)[12]) const {
    const uint32_t w0 = raw.w[0], w1 = raw.w[1], w2 = raw.w[2], w3 = raw.w[3], w4 = raw.w[4];
    const uint32_t mask = 0x007FF800u, one = 0x3F800000u;
#define ISP_B(x) biased_from_shifted((x), mask, one)
    v[0] = ISP_B(w0 << 3);                          // pixel -2: bits 8..19 of w0
    v[1] = ISP_B(w0 >> 9);                          // pixel -1: bits 20..31 of w0
    v[2] = ISP_B(w1 << 11);                         // pixel 0 : bits 0..11 of w1
    v[3] = ISP_B(w1 >> 1);                          // pixel 1 : bits 12..23
    v[4] = ISP_B(__funnelshift_r(w1, w2, 13));      // pixel 2 : bits 24..35 of (w2:w1)
    v[5] = ISP_B(w2 << 7);                          // pixel 3 : bits 4..15 of w2
    v[6] = ISP_B(w2 >> 5);                          // pixel 4 : bits 16..27 of w2
    v[7] = ISP_B(__funnelshift_r(w2, w3, 17));      // pixel 5 : bits 28..39 of (w3:w2)
    v[8] = ISP_B(w3 << 3);                          // pixel 6 : bits 8..19 of w3
    v[9] = ISP_B(w3 >> 9);                          // pixel 7 : bits 20..31 of w3
    v[10] = ISP_B(w4 << 11);                        // pixel 8
    v[11] = ISP_B(w4 >> 1);                         // pixel 9
#undef ISP_B
    if constexpr (CAM16) {
      constexpr float k = 4096.f * kInv4095;         // (b - 1) * 4096 * f32(1/4095), one rounding
#pragma unroll
      for (int j = 0; j < 12; j += 2) {
        const __half2 h = __floats2half2_rn(fmaf(v[j], k, -k), fmaf(v[j + 1], k, -k));
        v[j] = __low2float(h);
        v[j + 1] = __high2float(h);
      }
    }
  }
};

// per-pixel CFA sample, literal packed.py:23-31 + :98-100 (used by the samplers and border kernels)
template <bool CAM16>
struct Packed12Src {
  FramePtrs fp;
  int pitch;             // bytes per packed row
  __device__ __forceinline__ float at(int frame, int r, int c) const {
    const uint8_t* p = fp.in[frame] + (size_t)r * pitch + 3 * (c >> 1);
    const uint32_t b1 = p[1];
    const uint32_t v = (c & 1) ? ((uint32_t)p[2] << 4) | (b1 >> 4) : ((b1 & 0xFu) << 8) | p[0];
    return round_isp<CAM16>(__fmul_rn((float)v, kInv4095));
  }
};

// ---------------------------------------------------------------- shared front end: sums -> ISP RGB
struct IspConsts {
  int H, W, pattern;
  int ccm;
  float m[9];
  // tone map
  float gamma, intensity, la, ca;
  const float* metrics;      // device, 9 floats
  Workspace* ws;
  int frame0;
};

// literal front end for one pixel (bayer.py:137-155 + ISP dtype rounding), used off the hot path
template <bool CAM16>
__device__ __forceinline__ void isp_rgb_pixel(const Packed12Src<CAM16>& src, const IspConsts& k, int frame, int row, int col,
                                              float (&rgb)[3]) {
  float c[3], t[3];
  malvar_pixel(src, frame, k.pattern, row, col, k.H, k.W, c, t);
  float r = __fdiv_rn(c[0], t[0]), g = __fdiv_rn(c[1], t[1]), b = __fdiv_rn(c[2], t[2]);   // in_scale = 1.0
  if (k.ccm) ccm_apply(k.m, r, g, b);
  rgb[0] = round_isp<CAM16>(clamp01(r));
  rgb[1] = round_isp<CAM16>(clamp01(g));
  rgb[2] = round_isp<CAM16>(clamp01(b));
}

// hot-path front end: scaled filter sums (value * scale = x16 sum; biased by 16 for Camera32) -> ISP RGB in [0,1]
// cr/cg/cb are the compile-time SiteScale factors of this pixel; they fold into the constants.
template <bool CAM16, bool CCM>
__device__ __forceinline__ void isp_rgb_fast(const IspConsts& k, float sr, float sg, float sb, float cr, float cg, float cb,
                                             float (&rgb)[3]) {
  float r, g, b;
  if constexpr (CAM16) {
    r = sr * (cr * 0.0625f); g = sg * (cg * 0.0625f); b = sb * (cb * 0.0625f);
  } else {
    constexpr float kn = 256.f * kInv4095;           // 4096 * f32(1/4095) / 16
    r = fmaf(sr, cr * kn, -16.f * kn); g = fmaf(sg, cg * kn, -16.f * kn); b = fmaf(sb, cb * kn, -16.f * kn);
  }
  if constexpr (CCM) {
    const float x = fmaf(b, k.m[2], fmaf(g, k.m[1], r * k.m[0]));
    const float y = fmaf(b, k.m[5], fmaf(g, k.m[4], r * k.m[3]));
    const float z = fmaf(b, k.m[8], fmaf(g, k.m[7], r * k.m[6]));
    r = x; g = y; b = z;
  }
  r = clamp01(r); g = clamp01(g); b = clamp01(b);
  if constexpr (CAM16) {
    const __half2 h = __floats2half2_rn(r, g);
    rgb[0] = __low2float(h); rgb[1] = __high2float(h); rgb[2] = __half2float(__float2half_rn(b));
  } else {
    rgb[0] = r; rgb[1] = g; rgb[2] = b;
  }
}

// ---------------------------------------------------------------- quantisers
// trunc(y * scale) for y in [0, ~1]: one FMA in round-toward-zero mode against 2^23 leaves the integer
// in the low mantissa bits (no F2I); differs from trunc(rn(y*scale)) only when y*scale rounds up to an
// integer within half an ulp, i.e. by at most 1 LSB in ~1e-7 of the cases.
template <typename OutT> struct Quant;
template <> struct Quant<uint8_t> {
  static constexpr int kWords = 6;     // 24 bytes per 8 pixels
  static __device__ __forceinline__ uint32_t q(float y) { return __float_as_uint(__fmaf_rz(y, 255.f, 8388608.f)); }
  static __device__ __forceinline__ void pack(const uint32_t (&v)[24], uint32_t (&w)[6]) {
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      const uint32_t lo = __byte_perm(v[4 * i], v[4 * i + 1], 0x0040);
      const uint32_t hi = __byte_perm(v[4 * i + 2], v[4 * i + 3], 0x0040);
      w[i] = __byte_perm(lo, hi, 0x5410);
    }
  }
};
template <> struct Quant<uint16_t> {
  static constexpr int kWords = 12;
  static __device__ __forceinline__ uint32_t q(float y) { return __float_as_uint(__fmaf_rz(y, 65535.f, 8388608.f)); }
  static __device__ __forceinline__ void pack(const uint32_t (&v)[24], uint32_t (&w)[12]) {
#pragma unroll
    for (int i = 0; i < 12; ++i) w[i] = __byte_perm(v[2 * i], v[2 * i + 1], 0x5410);
  }
};
template <> struct Quant<__half> {
  static constexpr int kWords = 12;
  static __device__ __forceinline__ uint32_t q(float y) { return __float_as_uint(y); }
  static __device__ __forceinline__ void pack(const uint32_t (&v)[24], uint32_t (&w)[12]) {
#pragma unroll
    for (int i = 0; i < 12; ++i) {
      const __half2 h = __floats2half2_rn(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1]));
      w[i] = *reinterpret_cast<const uint32_t*>(&h);
    }
  }
};
template <> struct Quant<float> {
  static constexpr int kWords = 24;
  static __device__ __forceinline__ uint32_t q(float y) { return __float_as_uint(y); }
  static __device__ __forceinline__ void pack(const uint32_t (&v)[24], uint32_t (&w)[24]) {
#pragma unroll
    for (int i = 0; i < 24; ++i) w[i] = v[i];
  }
};

// quantised row of 8 pixels -> packed words -> warp-cooperative contiguous store (stream_engine.cuh)
template <typename OutT>
__device__ __forceinline__ void store_row8(const WarpCtx& wc, OutT* warp_out /* frame + 24 * tcol0 */, int W, int row,
                                           const uint32_t (&v)[24]) {
  constexpr int NW = Quant<OutT>::kWords;
  uint32_t w[NW];
  Quant<OutT>::pack(v, w);
  warp_store_row<NW>(wc, warp_out + (size_t)((unsigned)row * (unsigned)W) * 3, w);
}

template <typename OutT> __device__ __forceinline__ void store_px(void* frame_out, int W, int row, int col, const float (&y)[3]) {
  OutT* dst = reinterpret_cast<OutT*>(frame_out) + ((size_t)row * W + col) * 3;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    if constexpr (DT<OutT>::is_int) dst[k] = (OutT)(Quant<OutT>::q(y[k]) & (uint32_t)DT<OutT>::scale);
    else dst[k] = cast_from_f32<OutT>(y[k]);
  }
}

// ---------------------------------------------------------------- tone-map stages (per pixel, shared by hot + border)
struct LinearConsts { float a, bmin, inv_gamma; int has_gamma; };

__device__ __forceinline__ LinearConsts linear_consts(const float* __restrict__ metrics, float gamma) {
  LinearConsts c;
  const float bmin = metrics[0], bmax = metrics[1];
  c.a = __fdiv_rn(1.0f, __fsub_rn(bmax, bmin));        // tonemap.py:12
  c.bmin = bmin;
  c.inv_gamma = __fdiv_rn(1.0f, gamma);
  c.has_gamma = gamma != 1.0f;
  return c;
}

// tonemap.py:15-16: clamp(((x - min) * inv_range)^(1/gamma), 0, 1).  The subtraction and the product
// are rounded separately, exactly like the reference expression: saturated pixels (x == max) then hit
// the same side of the truncating quantiser as the reference (an FMA would not).
template <bool GAMMA>
__device__ __forceinline__ void linear_px(const LinearConsts& c, const float (&rgb)[3], float (&y)[3]) {
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    float v = __saturatef(__fmul_rn(__fsub_rn(rgb[k], c.bmin), c.a));
    if constexpr (GAMMA) v = __saturatef(fast_pow(v, c.inv_gamma));
    y[k] = v;
  }
}

struct ReinhardConsts { ReinhardParams p; float b; float out_scale_inv_max; float inv_gamma; int has_gamma; int ca0; };

template <bool CAM16, bool CA0>
__device__ __forceinline__ void reinhard_p(const ReinhardConsts& c, const float (&rgb)[3], float (&p)[3]) {
  float s[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) s[k] = fmaf(rgb[k], c.p.inv_range, c.b);
  reinhard_map_fast<CA0>(c.p, s, p);
}

// camera_isp.py:211-218: stored = cast_T(p); out = trunc(scale * (stored / max_out)^(1/gamma))
template <bool CAM16, bool GAMMA>
__device__ __forceinline__ void reinhard_out(const ReinhardConsts& c, const float (&p)[3], float (&y)[3]) {
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    float q = fmaxf(round_isp<CAM16>(p[k]) * c.out_scale_inv_max, 0.f);     // NaN / negative -> 0
    if constexpr (GAMMA) q = fast_pow(q, c.inv_gamma);
    y[k] = fminf(q, 1.0f);     // the reference does not clamp (q <= 1 + one f16 ulp); saturate for the RZ-FMA quantiser
  }
}

// ---------------------------------------------------------------- hot-path epilogues
// The per-pixel code is branch-free: the runtime options (CCM on/off, gamma != 1, color_adapt == 0) are
// template flags of emit_t and are selected once per 8-pixel row by a warp-uniform switch, so the
// compiler can interleave the eight independent pixel chains of a row.
#define ISP_FLAG_DISPATCH2(F0, F1, CALL)                               \
  do {                                                                 \
    if (F0) { if (F1) { CALL(true, true); } else { CALL(true, false); } } \
    else    { if (F1) { CALL(false, true); } else { CALL(false, false); } } \
  } while (0)

template <bool CAM16, typename OutT>
struct EpiRgb {       // load_packed12: ISP-dtype float RGB out
  FramePtrs fp;
  IspConsts k;
  static constexpr int kStageWords = 32 * Quant<OutT>::kWords;
  struct State { OutT* out; WarpCtx wc; };
  __device__ __forceinline__ void init(State& st, int frame, int, const WarpCtx& wc) const {
    st.out = reinterpret_cast<OutT*>(fp.out[k.frame0 + frame]) + 24 * wc.tcol0;
    st.wc = wc;
  }
  __device__ __forceinline__ void finish(State&, int, int, bool) const {}
  template <bool BROW, bool GFIRST, bool CCM>
  __device__ __forceinline__ void emit_t(const State& st, int row, const float (&R)[8], const float (&G)[8], const float (&B)[8]) const {
    using SS = SiteScale<BROW, GFIRST>;
    uint32_t v[24];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float rgb[3];
      isp_rgb_fast<CAM16, CCM>(k, R[j], G[j], B[j], SS::r(j), SS::g(j), SS::b(j), rgb);
      v[3 * j] = __float_as_uint(rgb[0]); v[3 * j + 1] = __float_as_uint(rgb[1]); v[3 * j + 2] = __float_as_uint(rgb[2]);
    }
    store_row8<OutT>(st.wc, st.out, k.W, row, v);
  }
  template <bool BROW, bool GFIRST>
  __device__ __forceinline__ void emit(State& st, int row, const float (&R)[8], const float (&G)[8], const float (&B)[8]) const {
    if (k.ccm) emit_t<BROW, GFIRST, true>(st, row, R, G, B);
    else emit_t<BROW, GFIRST, false>(st, row, R, G, B);
  }
};

template <bool CAM16, typename OutT>
struct EpiLinear {
  FramePtrs fp;
  IspConsts k;
  static constexpr int kStageWords = 32 * Quant<OutT>::kWords;
  struct State { LinearConsts c; OutT* out; WarpCtx wc; };
  __device__ __forceinline__ void init(State& st, int frame, int, const WarpCtx& wc) const {
    st.c = linear_consts(k.metrics, k.gamma);
    st.out = reinterpret_cast<OutT*>(fp.out[k.frame0 + frame]) + 24 * wc.tcol0;
    st.wc = wc;
  }
  __device__ __forceinline__ void finish(State&, int, int, bool) const {}
  template <bool BROW, bool GFIRST, bool CCM, bool GAMMA>
  __device__ __forceinline__ void emit_t(const State& st, int row, const float (&R)[8], const float (&G)[8], const float (&B)[8]) const {
    using SS = SiteScale<BROW, GFIRST>;
    uint32_t v[24];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float rgb[3], y[3];
      isp_rgb_fast<CAM16, CCM>(k, R[j], G[j], B[j], SS::r(j), SS::g(j), SS::b(j), rgb);
      linear_px<GAMMA>(st.c, rgb, y);
      v[3 * j] = Quant<OutT>::q(y[0]); v[3 * j + 1] = Quant<OutT>::q(y[1]); v[3 * j + 2] = Quant<OutT>::q(y[2]);
    }
    store_row8<OutT>(st.wc, st.out, k.W, row, v);
  }
  template <bool BROW, bool GFIRST>
  __device__ __forceinline__ void emit(State& st, int row, const float (&R)[8], const float (&G)[8], const float (&B)[8]) const {
#define ISP_CALL(A, B_) emit_t<BROW, GFIRST, A, B_>(st, row, R, G, B)
    ISP_FLAG_DISPATCH2(k.ccm, st.c.has_gamma, ISP_CALL);
#undef ISP_CALL
  }
};

__device__ __forceinline__ ReinhardConsts reinhard_consts(const IspConsts& k, int frame, bool with_max) {
  ReinhardConsts c;
  c.p = reinhard_params(k.metrics, k.intensity, k.la, k.ca);
  c.b = -c.p.bmin * c.p.inv_range;
  c.ca0 = k.ca == 0.f;
  c.inv_gamma = (float)(1.0 / (double)k.gamma);
  c.has_gamma = k.gamma != 1.0f;
  c.out_scale_inv_max = 0.f;
  if (with_max) c.out_scale_inv_max = __fdiv_rn(1.0f, fmaxf(1e-6f, __ldcg(&k.ws->frame_max[k.frame0 + frame])));
  return c;
}

template <bool CAM16>
struct EpiReinhardMax {      // pass 1 without the write-back: frame-global max of the mapped values
  IspConsts k;
  static constexpr int kStageWords = 0;
  struct State { ReinhardConsts c; float mx; bool first, last; };
  __device__ __forceinline__ void init(State& st, int frame, int tcol, const WarpCtx&) const {
    st.c = reinhard_consts(k, frame, false);
    st.mx = 0.f;
    st.first = tcol == 0;
    st.last = tcol == (k.W >> 3) - 1;
  }
  template <bool BROW, bool GFIRST, bool CCM, bool CA0>
  __device__ __forceinline__ void emit_t(State& st, const float (&R)[8], const float (&G)[8], const float (&B)[8]) const {
    using SS = SiteScale<BROW, GFIRST>;
    float mx = st.mx;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float rgb[3], p[3];
      isp_rgb_fast<CAM16, CCM>(k, R[j], G[j], B[j], SS::r(j), SS::g(j), SS::b(j), rgb);
      reinhard_p<CAM16, CA0>(st.c, rgb, p);
      float m = fmaxf(p[0], fmaxf(p[1], p[2]));
      if ((j < 2 && st.first) || (j >= 6 && st.last)) m = 0.f;     // image frame: border kernel
      mx = fmaxf(mx, m);
    }
    st.mx = mx;
  }
  template <bool BROW, bool GFIRST>
  __device__ __forceinline__ void emit(State& st, int row, const float (&R)[8], const float (&G)[8], const float (&B)[8]) const {
    // the 2-pixel image frame is handled (with the exact border normalisation) by the border kernel
    if (row < 2 || row >= k.H - 2) return;
#define ISP_CALL(A, B_) emit_t<BROW, GFIRST, A, B_>(st, R, G, B)
    ISP_FLAG_DISPATCH2(k.ccm, st.c.ca0, ISP_CALL);
#undef ISP_CALL
  }
  __device__ __forceinline__ void finish(State& st, int frame, int lane, bool task_ok) const {
    const float m = warp_max(st.mx);
    if (lane == 0 && task_ok && m > 0.f)
      atomicMax(reinterpret_cast<unsigned int*>(&k.ws->frame_max[k.frame0 + frame]), __float_as_uint(m));
  }
};

template <bool CAM16, typename OutT>
struct EpiReinhard {         // pass 2 recomputed from the packed frame: map, normalise by the max, gamma, quantise
  FramePtrs fp;
  IspConsts k;
  static constexpr int kStageWords = 32 * Quant<OutT>::kWords;
  struct State { ReinhardConsts c; OutT* out; WarpCtx wc; };
  __device__ __forceinline__ void init(State& st, int frame, int, const WarpCtx& wc) const {
    st.c = reinhard_consts(k, frame, true);
    st.out = reinterpret_cast<OutT*>(fp.out[k.frame0 + frame]) + 24 * wc.tcol0;
    st.wc = wc;
  }
  __device__ __forceinline__ void finish(State&, int, int, bool) const {}
  template <bool BROW, bool GFIRST, bool CCM, bool CA0, bool GAMMA>
  __device__ __forceinline__ void emit_t(const State& st, int row, const float (&R)[8], const float (&G)[8], const float (&B)[8]) const {
    using SS = SiteScale<BROW, GFIRST>;
    uint32_t v[24];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float rgb[3], p[3], y[3];
      isp_rgb_fast<CAM16, CCM>(k, R[j], G[j], B[j], SS::r(j), SS::g(j), SS::b(j), rgb);
      reinhard_p<CAM16, CA0>(st.c, rgb, p);
      reinhard_out<CAM16, GAMMA>(st.c, p, y);
      v[3 * j] = Quant<OutT>::q(y[0]); v[3 * j + 1] = Quant<OutT>::q(y[1]); v[3 * j + 2] = Quant<OutT>::q(y[2]);
    }
    store_row8<OutT>(st.wc, st.out, k.W, row, v);
  }
  template <bool BROW, bool GFIRST>
  __device__ __forceinline__ void emit(State& st, int row, const float (&R)[8], const float (&G)[8], const float (&B)[8]) const {
    if (st.c.has_gamma) {
#define ISP_CALL(A, B_) emit_t<BROW, GFIRST, A, B_, true>(st, row, R, G, B)
      ISP_FLAG_DISPATCH2(k.ccm, st.c.ca0, ISP_CALL);
#undef ISP_CALL
    } else {
#define ISP_CALL(A, B_) emit_t<BROW, GFIRST, A, B_, false>(st, row, R, G, B)
      ISP_FLAG_DISPATCH2(k.ccm, st.c.ca0, ISP_CALL);
#undef ISP_CALL
    }
  }
};

// ---------------------------------------------------------------- border kernel (2-pixel frame, exact normalisation)
enum { MODE_RGB = 0, MODE_LINEAR = 1, MODE_RMAX = 2, MODE_REINHARD = 3 };

template <bool CAM16, int MODE, typename OutT>
__global__ void __launch_bounds__(256) isp_border_kernel(const Packed12Src<CAM16> src, const FramePtrs fp, const IspConsts k,
                                                         int nframes, long long per_frame) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool ok = idx < per_frame * nframes;
  float mx = 0.f;
  int frame = 0;
  if (ok) {
    frame = (int)(idx / per_frame);
    int row, col;
    border_coord(idx % per_frame, k.H, k.W, row, col);
    float rgb[3];
    isp_rgb_pixel<CAM16>(src, k, k.frame0 + frame, row, col, rgb);
    void* out = fp.out[k.frame0 + frame];
    if constexpr (MODE == MODE_RGB) {
      store_px<OutT>(out, k.W, row, col, rgb);
    } else if constexpr (MODE == MODE_LINEAR) {
      const LinearConsts c = linear_consts(k.metrics, k.gamma);
      float y[3];
      if (c.has_gamma) linear_px<true>(c, rgb, y); else linear_px<false>(c, rgb, y);
      store_px<OutT>(out, k.W, row, col, y);
    } else {
      const ReinhardConsts c = reinhard_consts(k, frame, MODE == MODE_REINHARD);
      float p[3];
      if (c.ca0) reinhard_p<CAM16, true>(c, rgb, p); else reinhard_p<CAM16, false>(c, rgb, p);
      if constexpr (MODE == MODE_RMAX) {
        mx = fmaxf(p[0], fmaxf(p[1], p[2]));
      } else {
        float y[3];
        if (c.has_gamma) reinhard_out<CAM16, true>(c, p, y); else reinhard_out<CAM16, false>(c, p, y);
        store_px<OutT>(out, k.W, row, col, y);
      }
    }
  }
  if constexpr (MODE == MODE_RMAX) {
    // a warp may straddle two frames only at a frame boundary; keep it simple: one atomic per thread with m > 0
    // is avoided by a warp reduction when the whole warp sits in one frame.
    const int f0 = __shfl_sync(0xffffffffu, frame, 0);
    const bool uniform = __all_sync(0xffffffffu, frame == f0);
    if (uniform) {
      const float m = warp_max(mx);
      if ((threadIdx.x & 31) == 0 && m > 0.f)
        atomicMax(reinterpret_cast<unsigned int*>(&k.ws->frame_max[k.frame0 + f0]), __float_as_uint(m));
    } else if (ok && mx > 0.f) {
      atomicMax(reinterpret_cast<unsigned int*>(&k.ws->frame_max[k.frame0 + frame]), __float_as_uint(mx));
    }
  }
}

// ---------------------------------------------------------------- metering samplers straight from packed12
// generic: any stride, literal per-pixel demosaic
template <bool CAM16>
struct Packed12Sampler {
  Packed12Src<CAM16> src;
  IspConsts k;
  int stride, hs, ws_;
  __device__ __forceinline__ void sample(long long idx, float (&rgb)[3]) const {
    const int j = (int)(idx % ws_);
    const long long q = idx / ws_;
    const int i = (int)(q % hs);
    const int f = (int)(q / hs);
    isp_rgb_pixel<CAM16>(src, k, f, i * stride, j * stride, rgb);
  }
};

// stride % 8 == 0 (the default 8): a sample is pixel 0 of thread column col/8 of the streaming kernel.
// It reads nine 32-bit words (1/2/3/2/1 over the five rows; consecutive samples of a row are 12 bytes
// apart, so a warp's loads are contiguous) and evaluates exactly the partial-sum formulas of
// malvar_row + isp_rgb_fast, i.e. the very value the sweep will produce for that pixel.  Samples on the
// 2-pixel image frame (row 0, column 0) take the literal border path.
template <bool CAM16>
struct Packed12FastSampler {
  Packed12Src<CAM16> src;
  IspConsts k;
  int stride, hs, ws_;
  int pitch_words;

  static __device__ __forceinline__ float dec(uint32_t shifted) {
    const float b = __uint_as_float((shifted & 0x007FF800u) | 0x3F800000u);      // 1 + v/4096
    if constexpr (CAM16) {
      constexpr float kk = 4096.f * kInv4095;
      return __half2float(__float2half_rn(fmaf(b, kk, -kk)));
    } else {
      return b;
    }
  }

  __device__ __forceinline__ void sample(long long idx, float (&rgb)[3]) const {
    const int j = (int)(idx % ws_);
    const long long q = idx / ws_;
    const int i = (int)(q % hs);
    const int f = (int)(q / hs);
    const int row = i * stride, col = j * stride;
    if (row < 2 || row >= k.H - 2 || col < 2 || col >= k.W - 2) {
      isp_rgb_pixel<CAM16>(src, k, f, row, col, rgb);
      return;
    }
    const uint32_t* p = reinterpret_cast<const uint32_t*>(src.fp.in[f]) + (size_t)row * pitch_words + 3 * (col >> 3);
    const uint32_t a0 = __ldg(p - 2 * pitch_words);                                   // row-2: col
    const uint32_t bm = __ldg(p - pitch_words - 1), b0 = __ldg(p - pitch_words);      // row-1: col-1..col+1
    const uint32_t cm = __ldg(p - 1), c0 = __ldg(p), c1 = __ldg(p + 1);               // row  : col-2..col+2
    const uint32_t dm = __ldg(p + pitch_words - 1), d0 = __ldg(p + pitch_words);      // row+1
    const uint32_t e0 = __ldg(p + 2 * pitch_words);                                   // row+2
    // pixel col-2 = bits 8..19 of word -1, col-1 = bits 20..31; col = bits 0..11 of word 0, col+1 = bits 12..23,
    // col+2 = bits 24..35 of (word 1 : word 0)
    const float C = dec(c0 << 11);
    const float EW = dec(cm >> 9) + dec(c0 >> 1);
    const float EEWW = dec(cm << 3) + dec(__funnelshift_r(c0, c1, 13));
    const float NS = dec(b0 << 11) + dec(d0 << 11);
    const float D = (dec(bm >> 9) + dec(dm >> 9)) + (dec(b0 >> 1) + dec(d0 >> 1));
    const float NNSS = dec(a0 << 11) + dec(e0 << 11);
    const bool brow0 = (k.pattern == B200ISP_GBRG || k.pattern == B200ISP_BGGR);
    const bool gfirst0 = (k.pattern == B200ISP_GRBG || k.pattern == B200ISP_GBRG);
    const bool brow = brow0 != ((row & 1) != 0), gsite = gfirst0 != ((row & 1) != 0);
    float R, G, B, cr, cg, cb;
    if (!gsite) {
      float g2, opp4;
      malvar_csite(C, NS, EW, NNSS, EEWW, D, g2, opp4);
      G = g2; cg = 2.f;
      R = brow ? opp4 : C; cr = brow ? 4.f : 16.f;
      B = brow ? C : opp4; cb = brow ? 16.f : 4.f;
    } else {
      float h2, v2;
      malvar_gsite(C, NS, EW, NNSS, EEWW, D, h2, v2);
      G = C; cg = 16.f;
      R = brow ? v2 : h2; cr = 2.f;
      B = brow ? h2 : v2; cb = 2.f;
    }
    if (k.ccm) isp_rgb_fast<CAM16, true>(k, R, G, B, cr, cg, cb, rgb);
    else isp_rgb_fast<CAM16, false>(k, R, G, B, cr, cg, cb, rgb);
  }
};

// ---------------------------------------------------------------- host orchestration
template <bool CAM16, int MODE, typename OutT>
static int run_pass(const FramePtrs& fp, IspConsts k, int frame0, int nframes, int rows_per_task, cudaStream_t s,
                    void* ev_start = nullptr, void* ev_stop = nullptr) {
  k.frame0 = frame0;
  const StreamGeom g = make_geom(k.H, k.W, nframes, rows_per_task);
  Packed12Loader<CAM16> ld;
  ld.fp = fp; ld.pitch_words = k.W * 3 / 8; ld.frame0 = frame0;
  int st = B200ISP_OK;
  if (ev_start) cudaEventRecord((cudaEvent_t)ev_start, s);
  ISP_DISPATCH_PATTERN(k.pattern, P, {
    if constexpr (MODE == MODE_RGB) { EpiRgb<CAM16, OutT> e{fp, k}; st = launch_stream<P>(ld, e, g, s, "isp_stream<rgb>"); }
    else if constexpr (MODE == MODE_LINEAR) { EpiLinear<CAM16, OutT> e{fp, k}; st = launch_stream<P>(ld, e, g, s, "isp_stream<linear>"); }
    else if constexpr (MODE == MODE_RMAX) { EpiReinhardMax<CAM16> e{k}; st = launch_stream<P>(ld, e, g, s, "isp_stream<reinhard_max>"); }
    else { EpiReinhard<CAM16, OutT> e{fp, k}; st = launch_stream<P>(ld, e, g, s, "isp_stream<reinhard>"); }
  });
  if (ev_stop) cudaEventRecord((cudaEvent_t)ev_stop, s);
  if (st) return st;
  Packed12Src<CAM16> src{fp, k.W * 3 / 2};
  const long long per_frame = border_count(k.H, k.W);
  const long long total = per_frame * nframes;
  isp_border_kernel<CAM16, MODE, OutT><<<(unsigned)((total + 255) / 256), 256, 0, s>>>(src, fp, k, nframes, per_frame);
  return cuda_status(cudaPeekAtLastError(), "isp_border_kernel");
}

// frame-global Reinhard max for frames [frame0, frame0 + nframes): independent of the output dtype,
// instantiated once per ISP dtype (fused_inst.cu with ISP_INST_RMAX)
template <bool CAM16>
int run_rmax(const FramePtrs& fp, IspConsts k, int frame0, int nframes, int rows_per_task, cudaStream_t s) {
  return run_pass<CAM16, MODE_RMAX, uint8_t>(fp, k, frame0, nframes, rows_per_task, s);
}

template <bool CAM16, typename OutT>
int run_fused(const FramePtrs& fp, int n_frames, const b200isp_fused_params& p, IspConsts k, cudaStream_t s) {
  const int rpt = p.rows_per_task;
  constexpr bool kIspOut = (CAM16 && std::is_same<OutT, __half>::value) || (!CAM16 && std::is_same<OutT, float>::value);
  if (p.tonemap == B200ISP_TM_NONE) {
    if constexpr (kIspOut) return run_pass<CAM16, MODE_RGB, OutT>(fp, k, 0, n_frames, rpt, s, p.profile_start, p.profile_stop);
    else { set_error("process_packed12: TM_NONE writes the ISP dtype"); return B200ISP_E_DTYPE; }
  }
  if constexpr (std::is_same<OutT, float>::value) {
    set_error("process_packed12: tone-mapped output must be u8, u16 or f16");
    return B200ISP_E_DTYPE;
  } else {
    if (p.tonemap == B200ISP_TM_LINEAR) return run_pass<CAM16, MODE_LINEAR, OutT>(fp, k, 0, n_frames, rpt, s, p.profile_start, p.profile_stop);
    // Reinhard: the second sweep should find the packed frames in L2 -> interleave max / write passes
    // per group of frames whose packed bytes stay well inside the 126 MB L2.
    int st = cuda_status(cudaMemsetAsync(k.ws->frame_max, 0, sizeof(float) * B200ISP_MAX_FRAMES, s), "memset frame_max");
    if (st) return st;
    const long long frame_bytes = (long long)k.H * k.W * 3 / 2;
    int group = (int)((48LL << 20) / (frame_bytes > 0 ? frame_bytes : 1));
    if (group < 1) group = 1;
    for (int f = 0; f < n_frames; f += group) {
      const int n = (n_frames - f < group) ? n_frames - f : group;
      st = run_rmax<CAM16>(fp, k, f, n, rpt, s);
      if (st) return st;
      st = run_pass<CAM16, MODE_REINHARD, OutT>(fp, k, f, n, rpt, s, f == 0 ? p.profile_start : nullptr, f == 0 ? p.profile_stop : nullptr);
      if (st) return st;
    }
    return B200ISP_OK;
  }
}

extern template int run_rmax<true>(const FramePtrs&, IspConsts, int, int, int, cudaStream_t);
extern template int run_rmax<false>(const FramePtrs&, IspConsts, int, int, int, cudaStream_t);

}  // namespace isp
