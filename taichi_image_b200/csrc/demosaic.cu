// Stand-alone CFA mosaic / Malvar-He-Cutler demosaic on typed planes
// (reference: bayer.py:101-112 rgb_to_bayer_kernel, bayer.py:114-190 bayer_to_rgb_kernel).
#include "stream_engine.cuh"
#include "pixel_ops.cuh"

namespace isp {

// ---------------------------------------------------------------- typed row loaders
// ALIGNED: W % 8 == 0 and 16-byte aligned base -> word/vector loads; else per-element predicated loads.
template <typename T, bool ALIGNED> struct PlaneLoader;

template <typename T> struct PlaneCursor { const T* p; bool left, right; };

template <> struct PlaneLoader<uint8_t, true> {
  const uint8_t* base;
  struct Raw { uint32_t w[4]; };
  using Cursor = PlaneCursor<uint8_t>;
  __device__ __forceinline__ void open(Cursor& c, int, int tcol, const StreamGeom& g) const {
    c.p = base + 8 * tcol; c.left = tcol > 0; c.right = tcol < g.ntcols - 1;
  }
  __device__ __forceinline__ void fetch(const Cursor& c, int row, const StreamGeom& g, Raw& raw) const {
    const bool rv = (unsigned)row < (unsigned)g.H;
    const uint32_t* p = reinterpret_cast<const uint32_t*>(c.p + (rv ? (unsigned)row * (unsigned)g.W : 0u));
    raw.w[0] = (rv && c.left) ? __ldg(p - 1) : 0u;
    raw.w[1] = rv ? __ldg(p) : 0u;
    raw.w[2] = rv ? __ldg(p + 1) : 0u;
    raw.w[3] = (rv && c.right) ? __ldg(p + 2) : 0u;
  }
  __device__ __forceinline__ void prefetch(const Cursor&, int, const StreamGeom&) const {}
  __device__ __forceinline__ void decode(const Raw& raw, float (&v)[12]) const {
#pragma unroll
    for (int j = 0; j < 12; ++j) v[j] = (float)((raw.w[(j + 2) >> 2] >> (8 * ((j + 2) & 3))) & 0xFFu);
  }
};

template <typename T> struct PlaneLoader16 {   // u16 / i16 / f16
  const T* base;
  struct Raw { uint32_t w[6]; };
  using Cursor = PlaneCursor<T>;
  __device__ __forceinline__ void open(Cursor& c, int, int tcol, const StreamGeom& g) const {
    c.p = base + 8 * tcol; c.left = tcol > 0; c.right = tcol < g.ntcols - 1;
  }
  __device__ __forceinline__ void fetch(const Cursor& c, int row, const StreamGeom& g, Raw& raw) const {
    const bool rv = (unsigned)row < (unsigned)g.H;
    const uint32_t* p = reinterpret_cast<const uint32_t*>(c.p + (rv ? (unsigned)row * (unsigned)g.W : 0u));
    raw.w[0] = (rv && c.left) ? __ldg(p - 1) : 0u;
    uint4 q = make_uint4(0u, 0u, 0u, 0u);
    if (rv) q = __ldg(reinterpret_cast<const uint4*>(p));
    raw.w[1] = q.x; raw.w[2] = q.y; raw.w[3] = q.z; raw.w[4] = q.w;
    raw.w[5] = (rv && c.right) ? __ldg(p + 4) : 0u;
  }
  __device__ __forceinline__ void prefetch(const Cursor&, int, const StreamGeom&) const {}
  __device__ __forceinline__ void decode(const Raw& raw, float (&v)[12]) const {
#pragma unroll
    for (int j = 0; j < 12; ++j) {
      const uint16_t h = (uint16_t)(raw.w[j >> 1] >> (16 * (j & 1)));
      if constexpr (sizeof(T) == 2 && !DT<T>::is_int) v[j] = __half2float(__ushort_as_half(h));
      else if constexpr (std::is_same<T, int16_t>::value) v[j] = (float)(int16_t)h;
      else v[j] = (float)h;
    }
  }
};
template <> struct PlaneLoader<uint16_t, true> : PlaneLoader16<uint16_t> {};
template <> struct PlaneLoader<int16_t, true> : PlaneLoader16<int16_t> {};
template <> struct PlaneLoader<__half, true> : PlaneLoader16<__half> {};

template <> struct PlaneLoader<float, true> {
  const float* base;
  struct Raw { float v[12]; };
  using Cursor = PlaneCursor<float>;
  __device__ __forceinline__ void open(Cursor& c, int, int tcol, const StreamGeom& g) const {
    c.p = base + 8 * tcol; c.left = tcol > 0; c.right = tcol < g.ntcols - 1;
  }
  __device__ __forceinline__ void fetch(const Cursor& c, int row, const StreamGeom& g, Raw& raw) const {
    const bool rv = (unsigned)row < (unsigned)g.H;
    const float* p = c.p + (rv ? (unsigned)row * (unsigned)g.W : 0u);
    float2 l = make_float2(0.f, 0.f), r = make_float2(0.f, 0.f);
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
    if (rv && c.left) l = __ldg(reinterpret_cast<const float2*>(p - 2));
    if (rv) { a = __ldg(reinterpret_cast<const float4*>(p)); b = __ldg(reinterpret_cast<const float4*>(p + 4)); }
    if (rv && c.right) r = __ldg(reinterpret_cast<const float2*>(p + 8));
    raw.v[0] = l.x; raw.v[1] = l.y;
    raw.v[2] = a.x; raw.v[3] = a.y; raw.v[4] = a.z; raw.v[5] = a.w;
    raw.v[6] = b.x; raw.v[7] = b.y; raw.v[8] = b.z; raw.v[9] = b.w;
    raw.v[10] = r.x; raw.v[11] = r.y;
  }
  __device__ __forceinline__ void prefetch(const Cursor&, int, const StreamGeom&) const {}
  __device__ __forceinline__ void decode(const Raw& raw, float (&v)[12]) const {
#pragma unroll
    for (int j = 0; j < 12; ++j) v[j] = raw.v[j];
  }
};

// ---------------------------------------------------------------- 8-pixel RGB store
template <typename OutT, bool ALIGNED>
__device__ __forceinline__ void store_px8(OutT* dst /* at (row, 8*tcol, 0) */, const OutT (&o)[24], int ncols) {
  if constexpr (ALIGNED) {
    constexpr int kBytes = 24 * (int)sizeof(OutT);
    if constexpr (kBytes % 16 == 0) {
      const uint4* s = reinterpret_cast<const uint4*>(o);
      uint4* d = reinterpret_cast<uint4*>(dst);
#pragma unroll
      for (int i = 0; i < kBytes / 16; ++i) d[i] = s[i];
    } else {
      const uint2* s = reinterpret_cast<const uint2*>(o);
      uint2* d = reinterpret_cast<uint2*>(dst);
#pragma unroll
      for (int i = 0; i < kBytes / 8; ++i) d[i] = s[i];
    }
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (j < ncols) {
        dst[3 * j] = o[3 * j]; dst[3 * j + 1] = o[3 * j + 1]; dst[3 * j + 2] = o[3 * j + 2];
      }
  }
}

// ---------------------------------------------------------------- demosaic epilogue
// bayer.py:150-155, :132-134:  c = sum / (in_scale * 16); [c = M c]; clamp(c,0,1); cast(c*out_scale).
// Integer planes without CCM take clamp(floor(sum/16), 0, scale), which equals the float chain for
// every reachable sum (exhaustively checked in tests/test_host_cpu.py::test_integer_demosaic_equals_floor_division).
template <typename T>
struct EpiDemosaic {
  T* out;
  int W;
  int ccm;          // runtime (warp-uniform) flag
  float m[9];
  static constexpr int kRowWords = 24 * (int)sizeof(T) / 4;       // 32-bit words of one thread's 8 output pixels
  static constexpr int kStageWords = 32 * kRowWords;
  struct State { T* out; WarpCtx wc; };
  __device__ __forceinline__ void init(State& st, int, int, const WarpCtx& wc) const { st.out = out + 24 * wc.tcol0; st.wc = wc; }
  __device__ __forceinline__ void finish(State&, int, int, bool) const {}

  __device__ __forceinline__ void finish_px(float cr, float cg, float cb, T* o) const {
    if (ccm) ccm_apply(m, cr, cg, cb);
    constexpr float os = DT<T>::scale;
    o[0] = cast_from_f32<T>(__fmul_rn(clamp01(cr), os));
    o[1] = cast_from_f32<T>(__fmul_rn(clamp01(cg), os));
    o[2] = cast_from_f32<T>(__fmul_rn(clamp01(cb), os));
  }

  template <bool BROW, bool GFIRST>
  __device__ __forceinline__ void emit(State& st, int row, const float (&R)[8], const float (&G)[8], const float (&B)[8]) const {
    using SS = SiteScale<BROW, GFIRST>;
    constexpr float is = DT<T>::scale;
    alignas(16) T o[24];
    if (DT<T>::is_int && !ccm) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {      // value * scale / 16 is an exact power-of-two scaling of the integer sum
        o[3 * j]     = (T)(int)fminf(fmaxf(R[j] * (SS::r(j) * 0.0625f), 0.f), is);
        o[3 * j + 1] = (T)(int)fminf(fmaxf(G[j] * (SS::g(j) * 0.0625f), 0.f), is);
        o[3 * j + 2] = (T)(int)fminf(fmaxf(B[j] * (SS::b(j) * 0.0625f), 0.f), is);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        finish_px(__fdiv_rn(R[j] * SS::r(j), is * 16.f), __fdiv_rn(G[j] * SS::g(j), is * 16.f),
                  __fdiv_rn(B[j] * SS::b(j), is * 16.f), o + 3 * j);
    }
    uint32_t w[kRowWords];
#pragma unroll
    for (int i = 0; i < kRowWords; ++i) w[i] = reinterpret_cast<const uint32_t*>(o)[i];
    warp_store_row<kRowWords>(st.wc, st.out + (size_t)((unsigned)row * (unsigned)W) * 3, w);
  }
};

// ---------------------------------------------------------------- per-pixel kernel
// Literal bayer.py:137-155 (see pixel_ops.cuh): whole image (mixed dtypes, widths that are not a
// multiple of 8, unaligned views) or only the 2-pixel frame after a streaming launch.
template <typename T> struct PlaneSrc {
  const T* base; int W;
  __device__ __forceinline__ float at(int, int r, int c) const { return to_f32(base[(size_t)r * W + c]); }
};

template <typename InT, typename OutT>
__global__ void __launch_bounds__(256) demosaic_pixel_kernel(const InT* __restrict__ bayer, OutT* __restrict__ out,
                                                             int H, int W, int pattern, int ccm, const float9 m,
                                                             int border_only, long long count) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= count) return;
  int row, col;
  if (border_only) border_coord(idx, H, W, row, col);
  else { row = (int)(idx / W); col = (int)(idx % W); }
  const PlaneSrc<InT> src{bayer, W};
  float c[3], t[3];
  malvar_pixel(src, 0, pattern, row, col, H, W, c, t);
  constexpr float is = DT<InT>::scale, os = DT<OutT>::scale;
  float cr = __fdiv_rn(c[0], __fmul_rn(is, t[0])), cg = __fdiv_rn(c[1], __fmul_rn(is, t[1])), cb = __fdiv_rn(c[2], __fmul_rn(is, t[2]));
  if (ccm) ccm_apply(m.v, cr, cg, cb);
  OutT* o = out + ((size_t)row * W + col) * 3;
  o[0] = cast_from_f32<OutT>(__fmul_rn(clamp01(cr), os));
  o[1] = cast_from_f32<OutT>(__fmul_rn(clamp01(cg), os));
  o[2] = cast_from_f32<OutT>(__fmul_rn(clamp01(cb), os));
}

template <typename InT, typename OutT>
static int run_demosaic_pixel(const void* bayer, void* rgb, int H, int W, int pattern, const float* ccm,
                              bool border_only, cudaStream_t s) {
  float9 m;
  for (int i = 0; i < 9; ++i) m.v[i] = ccm ? ccm[i] : 0.f;
  const long long count = border_only ? border_count(H, W) : (long long)H * W;
  demosaic_pixel_kernel<InT, OutT><<<(unsigned)((count + 255) / 256), 256, 0, s>>>(
      (const InT*)bayer, (OutT*)rgb, H, W, pattern, ccm != nullptr, m, border_only ? 1 : 0, count);
  return cuda_status(cudaPeekAtLastError(), "demosaic_pixel_kernel");
}

template <typename T>
static int run_demosaic_stream(const void* bayer, void* rgb, int H, int W, int pattern, const float* ccm, cudaStream_t s) {
  PlaneLoader<T, true> ld;
  ld.base = (const T*)bayer;
  EpiDemosaic<T> epi;
  epi.out = (T*)rgb; epi.W = W; epi.ccm = ccm != nullptr;
  for (int i = 0; i < 9; ++i) epi.m[i] = ccm ? ccm[i] : 0.f;
  const StreamGeom g = make_geom(H, W, 1, 0);
  ISP_DISPATCH_PATTERN(pattern, P, {
    const int st = launch_stream<P>(ld, epi, g, s, "bayer_to_rgb");
    if (st) return st;
  });
  return run_demosaic_pixel<T, T>(bayer, rgb, H, W, pattern, ccm, true, s);
}

// ---------------------------------------------------------------- bilinear demosaic (EXTENSION)
// north_star names "Malvar-He-Cutler / bilinear demosaic"; the reference has only Malvar.  Defined as the
// classic 3x3 bilinear CFA interpolation evaluated with the SAME rule as bayer.py:137-155: per channel
// c = sum of in-bounds (v * w), t = sum of in-bounds w, c / (in_scale * t), [CCM], clamp, cast(c * out_scale) --
// i.e. the mean of the in-bounds neighbours of that colour.  Site kernels (x4; taps in row-major 3x3 order):
//   K0 R site: R = centre, G = N,S,E,W, B = 4 diagonals       K3 B site: K0 with R <-> B
//   K1 G site with R above/below: R = N,S, B = E,W            K2 G site with R left/right: R = E,W, B = N,S
// accumulate tap i (weight W) into channel CH exactly like the generic loop: c += v * w (separately rounded), t += w
#define ISP_BL_TAP(CH, I, Wt) { c[CH] = __fadd_rn(c[CH], __fmul_rn(v[I], Wt)); t[CH] += m[I] * Wt; }

template <typename InT, typename OutT>
__global__ void __launch_bounds__(256) demosaic_bilinear_kernel(const InT* __restrict__ bayer, OutT* __restrict__ out,
                                                                int H, int W, int pattern, int ccm, const float9 m9) {
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  const int row = blockIdx.y;
  if (col >= W || row >= H) return;
  const int K = site_kernel(pattern, row, col);
  // 3x3 neighbourhood, 0 outside the image (m = in-bounds indicator); taps in row-major order
  float v[9], m[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) {
    const int rr = row + i / 3 - 1, cc = col + i % 3 - 1;
    const bool in = rr >= 0 && rr < H && cc >= 0 && cc < W;
    v[i] = in ? to_f32(bayer[(size_t)rr * W + cc]) : 0.f;
    m[i] = in ? 1.f : 0.f;
  }
  float c[3] = {0.f, 0.f, 0.f}, t[3] = {0.f, 0.f, 0.f};
  // c_bilinear[K] with compile-time weights (zero-weight taps add exactly 0 and are skipped); same tap order
  if (K == 0) {              // R site: R centre, G cross, B diagonals
    ISP_BL_TAP(2, 0, 1.f) ISP_BL_TAP(1, 1, 1.f) ISP_BL_TAP(2, 2, 1.f) ISP_BL_TAP(1, 3, 1.f) ISP_BL_TAP(0, 4, 4.f)
    ISP_BL_TAP(1, 5, 1.f) ISP_BL_TAP(2, 6, 1.f) ISP_BL_TAP(1, 7, 1.f) ISP_BL_TAP(2, 8, 1.f)
  } else if (K == 3) {       // B site: the same with R <-> B
    ISP_BL_TAP(0, 0, 1.f) ISP_BL_TAP(1, 1, 1.f) ISP_BL_TAP(0, 2, 1.f) ISP_BL_TAP(1, 3, 1.f) ISP_BL_TAP(2, 4, 4.f)
    ISP_BL_TAP(1, 5, 1.f) ISP_BL_TAP(0, 6, 1.f) ISP_BL_TAP(1, 7, 1.f) ISP_BL_TAP(0, 8, 1.f)
  } else if (K == 1) {       // G site, R above / below, B left / right
    ISP_BL_TAP(0, 1, 2.f) ISP_BL_TAP(2, 3, 2.f) ISP_BL_TAP(1, 4, 4.f) ISP_BL_TAP(2, 5, 2.f) ISP_BL_TAP(0, 7, 2.f)
  } else {                   // G site, R left / right, B above / below
    ISP_BL_TAP(2, 1, 2.f) ISP_BL_TAP(0, 3, 2.f) ISP_BL_TAP(1, 4, 4.f) ISP_BL_TAP(0, 5, 2.f) ISP_BL_TAP(2, 7, 2.f)
  }
  constexpr float is = DT<InT>::scale, os = DT<OutT>::scale;
  // a 2-pixel-wide image can leave a colour without any in-bounds neighbour (t = 0): define it as 0
  float cr = t[0] > 0.f ? __fdiv_rn(c[0], __fmul_rn(is, t[0])) : 0.f;
  float cg = t[1] > 0.f ? __fdiv_rn(c[1], __fmul_rn(is, t[1])) : 0.f;
  float cb = t[2] > 0.f ? __fdiv_rn(c[2], __fmul_rn(is, t[2])) : 0.f;
  if (ccm) ccm_apply(m9.v, cr, cg, cb);
  OutT* o = out + ((size_t)row * W + col) * 3;
  o[0] = cast_from_f32<OutT>(__fmul_rn(clamp01(cr), os));
  o[1] = cast_from_f32<OutT>(__fmul_rn(clamp01(cg), os));
  o[2] = cast_from_f32<OutT>(__fmul_rn(clamp01(cb), os));
}
#undef ISP_BL_TAP

// ---------------------------------------------------------------- rgb_to_bayer (bayer.py:101-112)
// pixel_orders (bayer.py:85-90): channel at ((r0,c0),(r0,c1),(r1,c0),(r1,c1)), 2 bits each
template <typename T>
__global__ void __launch_bounds__(256) rgb_to_bayer_kernel(const T* __restrict__ rgb, T* __restrict__ bayer,
                                                           int H, int W, unsigned order) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int r = blockIdx.y;
  if (c >= (W & ~1) || r >= (H & ~1)) return;
  const int ch = (order >> (2 * ((r & 1) * 2 + (c & 1)))) & 3;
  bayer[(size_t)r * W + c] = rgb[((size_t)r * W + c) * 3 + ch];
}

// eight pixels per thread with 8 / 16-byte accesses (W % 8 == 0, 16-byte aligned bases)
template <typename T>
__global__ void __launch_bounds__(128) rgb_to_bayer_vec_kernel(const T* __restrict__ rgb, T* __restrict__ bayer,
                                                               int H, int W, unsigned order) {
  const int gx = blockIdx.x * blockDim.x + threadIdx.x;
  const int r = blockIdx.y;
  if (gx >= W / 8) return;
  alignas(16) T px[24];
  alignas(16) T out[8];
  ld_bytes<24 * sizeof(T)>(rgb + ((size_t)r * W + 8 * gx) * 3, px);
  const int ch0 = (order >> (2 * ((r & 1) * 2))) & 3, ch1 = (order >> (2 * ((r & 1) * 2 + 1))) & 3;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const int ch = (c & 1) ? ch1 : ch0;
    out[c] = ch == 0 ? px[3 * c] : (ch == 1 ? px[3 * c + 1] : px[3 * c + 2]);
  }
  st_bytes<8 * sizeof(T)>(bayer + (size_t)r * W + 8 * gx, out);
}

}  // namespace isp

using namespace isp;

extern "C" int b200isp_rgb_to_bayer(const void* rgb, void* bayer, int dtype, int height, int width,
                                    int pattern, b200isp_stream stream) {
  ISP_REQUIRE(height >= 0 && width >= 0, B200ISP_E_SHAPE, "rgb_to_bayer: negative size");
  ISP_REQUIRE(pattern >= 0 && pattern <= 3, B200ISP_E_ARG, "rgb_to_bayer: unknown pattern %d", pattern);
  if (height < 2 || width < 2) return B200ISP_OK;
  ISP_REQUIRE(rgb && bayer, B200ISP_E_ARG, "rgb_to_bayer: null pointer");
  // RGGB (0,1,1,2) GRBG (1,0,2,1) GBRG (1,2,0,1) BGGR (2,1,1,0)
  const unsigned orders[4] = {0u | (1u << 2) | (1u << 4) | (2u << 6), 1u | (0u << 2) | (2u << 4) | (1u << 6),
                              1u | (2u << 2) | (0u << 4) | (1u << 6), 2u | (1u << 2) | (1u << 4) | (0u << 6)};
  cudaStream_t s = (cudaStream_t)stream;
  if (width % 8 == 0 && height % 2 == 0 && ((reinterpret_cast<uintptr_t>(rgb) | reinterpret_cast<uintptr_t>(bayer)) & 15u) == 0) {
    const dim3 grid((width / 8 + 127) / 128, height);
    ISP_DISPATCH_DTYPE(dtype, T, (rgb_to_bayer_vec_kernel<T><<<grid, 128, 0, s>>>((const T*)rgb, (T*)bayer, height, width, orders[pattern])));
  } else {
    const dim3 grid((width + 255) / 256, height);
    ISP_DISPATCH_DTYPE(dtype, T, (rgb_to_bayer_kernel<T><<<grid, 256, 0, s>>>((const T*)rgb, (T*)bayer, height, width, orders[pattern])));
  }
  ISP_LAUNCH_CHECK("rgb_to_bayer_kernel");
  return B200ISP_OK;
}

extern "C" int b200isp_bayer_to_rgb(const void* bayer, int in_dtype, void* rgb, int out_dtype,
                                    int height, int width, int pattern, const float* ccm9_host,
                                    b200isp_stream stream) {
  ISP_REQUIRE(height >= 0 && width >= 0 && height % 2 == 0 && width % 2 == 0, B200ISP_E_SHAPE,
              "bayer_to_rgb: image must be even size, got %dx%d", height, width);
  ISP_REQUIRE(pattern >= 0 && pattern <= 3, B200ISP_E_ARG, "bayer_to_rgb: unknown pattern %d", pattern);
  ISP_REQUIRE(valid_dtype(in_dtype) && valid_dtype(out_dtype), B200ISP_E_DTYPE, "bayer_to_rgb: bad dtype");
  if (height == 0 || width == 0) return B200ISP_OK;
  ISP_REQUIRE(bayer && rgb, B200ISP_E_ARG, "bayer_to_rgb: null pointer");
  cudaStream_t s = (cudaStream_t)stream;
  const bool aligned = (width % 8 == 0) && ((reinterpret_cast<uintptr_t>(bayer) | reinterpret_cast<uintptr_t>(rgb)) & 15u) == 0;
  if (aligned && in_dtype == out_dtype && height >= 4 && width >= 8) {
    ISP_DISPATCH_DTYPE(in_dtype, T, return (run_demosaic_stream<T>(bayer, rgb, height, width, pattern, ccm9_host, s)));
  }
  ISP_DISPATCH_DTYPE(in_dtype, InT, {
    ISP_DISPATCH_DTYPE(out_dtype, OutT, return (run_demosaic_pixel<InT, OutT>(bayer, rgb, height, width, pattern, ccm9_host, false, s)));
  });
  return B200ISP_OK;
}

extern "C" int b200isp_bayer_to_rgb_bilinear(const void* bayer, int in_dtype, void* rgb, int out_dtype,
                                             int height, int width, int pattern, const float* ccm9_host,
                                             b200isp_stream stream) {
  ISP_REQUIRE(height >= 0 && width >= 0 && height % 2 == 0 && width % 2 == 0, B200ISP_E_SHAPE,
              "bayer_to_rgb_bilinear: image must be even size, got %dx%d", height, width);
  ISP_REQUIRE(pattern >= 0 && pattern <= 3, B200ISP_E_ARG, "bayer_to_rgb_bilinear: unknown pattern %d", pattern);
  ISP_REQUIRE(valid_dtype(in_dtype) && valid_dtype(out_dtype), B200ISP_E_DTYPE, "bayer_to_rgb_bilinear: bad dtype");
  if (height == 0 || width == 0) return B200ISP_OK;
  ISP_REQUIRE(bayer && rgb, B200ISP_E_ARG, "bayer_to_rgb_bilinear: null pointer");
  float9 m;
  for (int i = 0; i < 9; ++i) m.v[i] = ccm9_host ? ccm9_host[i] : 0.f;
  const dim3 grid((width + 255) / 256, height);
  cudaStream_t s = (cudaStream_t)stream;
  ISP_DISPATCH_DTYPE(in_dtype, InT, {
    ISP_DISPATCH_DTYPE(out_dtype, OutT, (demosaic_bilinear_kernel<InT, OutT><<<grid, 256, 0, s>>>(
        (const InT*)bayer, (OutT*)rgb, height, width, pattern, ccm9_host != nullptr, m)));
  });
  ISP_LAUNCH_CHECK("demosaic_bilinear_kernel");
  return B200ISP_OK;
}
