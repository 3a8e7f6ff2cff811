// Planar YUV 4:2:0 encode / decode (reference: color/yuv_420.py:12-90).
//
// Layout (yuv_420.py:95-118): one (3H/2, W) plane of the output dtype = H rows of Y followed by two (H/2, W/2)
// chroma planes; the SECOND matrix row goes to chroma plane 1 and the third to plane 0 (yuv_420.py:62-63).
// Reference quirks reproduced as they are: the BT.601 matrix is applied to the BGR-swizzled pixel
// (Y = 0.299 B + 0.587 G + 0.114 R, yuv_420.py:24-26) and `tm.clamp(0, 1, v)` evaluates to min(1, v) because Taichi's
// signature is clamp(x, xmin, xmax) (no lower clamp; negative values saturate to 0 in the integer casts here).
// Encode: one thread per 2x2 quad (4 pixels in, 4 Y + 2 chroma out); decode: one thread per 2 horizontally
// adjacent pixels (one chroma pair in, 6 values out).  Both are HBM-bound: 3 s_in + 1.5 s_out bytes per pixel.
#include <type_traits>
#include "common.cuh"

namespace isp {

struct Mat3 { float m[9]; };

__device__ __forceinline__ void mat_vec(const Mat3& M, float x0, float x1, float x2, float (&o)[3]) {
#pragma unroll
  for (int r = 0; r < 3; ++r)      // row dot products left to right, separately rounded (Taichi mat @ vec)
    o[r] = __fadd_rn(__fadd_rn(__fmul_rn(M.m[3 * r], x0), __fmul_rn(M.m[3 * r + 1], x1)), __fmul_rn(M.m[3 * r + 2], x2));
}

template <typename InT, typename OutT>
__global__ void __launch_bounds__(256) rgb_yuv420_kernel(const InT* __restrict__ rgb, OutT* __restrict__ yuv, int H, int W, Mat3 M) {
  const int qx = blockIdx.x * blockDim.x + threadIdx.x;     // quad column
  const int qy = blockIdx.y;                                // quad row
  if (qx >= W / 2 || qy >= H / 2) return;
  constexpr float is = DT<InT>::scale, os = DT<OutT>::scale;
  float u = 0.f, v = 0.f;
  bool first = true;
#pragma unroll
  for (int dy = 0; dy < 2; ++dy)
#pragma unroll
    for (int dx = 0; dx < 2; ++dx) {                        // ti.ndrange(2, 2): (0,0), (0,1), (1,0), (1,1)
      const int r = 2 * qy + dy, c = 2 * qx + dx;
      const InT* p = rgb + ((size_t)r * W + c) * 3;
      const float R = __fdiv_rn(to_f32(p[0]), is), G = __fdiv_rn(to_f32(p[1]), is), B = __fdiv_rn(to_f32(p[2]), is);
      float o[3];
      mat_vec(M, B, G, R, o);                               // applied to .bgr
      const float y = o[0], cu = __fadd_rn(o[1], 0.5f), cv = __fadd_rn(o[2], 0.5f);
      yuv[(size_t)r * W + c] = cast_from_f32<OutT>(__fmul_rn(fminf(1.0f, y), os));
      if (first) { u = cu; v = cv; first = false; } else { u = __fadd_rn(u, cu); v = __fadd_rn(v, cv); }
    }
  OutT* planes = yuv + (size_t)H * W;
  const size_t plane = (size_t)(H / 2) * (W / 2), idx = (size_t)qy * (W / 2) + qx;
  planes[plane + idx] = cast_from_f32<OutT>(__fmul_rn(fminf(1.0f, __fdiv_rn(u, 4.0f)), os));     // plane 1 <- second row
  planes[idx] = cast_from_f32<OutT>(__fmul_rn(fminf(1.0f, __fdiv_rn(v, 4.0f)), os));             // plane 0 <- third row
}

template <typename InT, typename OutT>
__global__ void __launch_bounds__(256) yuv420_rgb_kernel(const InT* __restrict__ yuv, OutT* __restrict__ rgb, int H, int W, Mat3 Minv) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int r = blockIdx.y;
  if (c >= W || r >= H) return;
  constexpr float is = DT<InT>::scale, os = DT<OutT>::scale;
  const InT* planes = yuv + (size_t)H * W;
  const size_t plane = (size_t)(H / 2) * (W / 2), idx = (size_t)(r / 2) * (W / 2) + (c / 2);
  const float y = __fdiv_rn(to_f32(yuv[(size_t)r * W + c]), is);
  const float cu = __fsub_rn(__fdiv_rn(to_f32(planes[plane + idx]), is), 0.5f);
  const float cv = __fsub_rn(__fdiv_rn(to_f32(planes[idx]), is), 0.5f);
  float o[3];
  mat_vec(Minv, y, cu, cv, o);                              // = bgr
  OutT* dst = rgb + ((size_t)r * W + c) * 3;
  dst[0] = cast_from_f32<OutT>(__fmul_rn(fminf(1.0f, o[2]), os));
  dst[1] = cast_from_f32<OutT>(__fmul_rn(fminf(1.0f, o[1]), os));
  dst[2] = cast_from_f32<OutT>(__fmul_rn(fminf(1.0f, o[0]), os));
}

// ---------------------------------------------------------------- vector forms (W % 8 == 0, 16-byte aligned bases)
// The same arithmetic in the same order, eight pixels (four quads) per thread and row with 8 / 16-byte accesses:
// the scalar kernels above issue one 1..4-byte request per element and are request-bound (13-15 % of HBM peak).
// to_f32(x) / in_scale (yuv_420.py:50, :80).  For u8 there are only 256 inputs: a shared-memory table of the correctly
// rounded quotients replaces the IEEE division (the kernels are instruction-bound, not HBM-bound, with it).
template <typename InT> struct UnitScale {
  static constexpr bool kTable = std::is_same<InT, uint8_t>::value;
  float* lut;
  __device__ __forceinline__ void init(float* smem) {
    lut = smem;
    if constexpr (kTable) {
      for (int i = threadIdx.x; i < 256; i += blockDim.x) smem[i] = __fdiv_rn((float)i, 255.0f);
      __syncthreads();
    }
  }
  __device__ __forceinline__ float operator()(InT x) const {
    if constexpr (kTable) return lut[x];
    else return __fdiv_rn(to_f32(x), DT<InT>::scale);
  }
};

template <typename InT, typename OutT>
__global__ void __launch_bounds__(256) rgb_yuv420_vec_kernel(const InT* __restrict__ rgb, OutT* __restrict__ yuv, int H, int W, Mat3 M) {
  __shared__ float smem_lut[UnitScale<InT>::kTable ? 256 : 1];
  UnitScale<InT> unit;
  unit.init(smem_lut);
  const int gx = blockIdx.x * blockDim.x + threadIdx.x;     // group of four quads = pixel columns 8 gx .. 8 gx + 7
  const int qy = blockIdx.y;
  if (gx >= W / 8) return;
  constexpr float os = DT<OutT>::scale;
  alignas(16) InT px[2][24];
  alignas(16) OutT yrow[2][8];
  alignas(16) OutT cu4[4], cv4[4];
#pragma unroll
  for (int dy = 0; dy < 2; ++dy) ld_bytes<24 * sizeof(InT)>(rgb + ((size_t)(2 * qy + dy) * W + 8 * gx) * 3, px[dy]);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    float u = 0.f, v = 0.f;
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {                      // (0,0), (0,1), (1,0), (1,1) like the scalar kernel
        const InT* p = &px[dy][3 * (2 * q + dx)];
        const float R = unit(p[0]), G = unit(p[1]), B = unit(p[2]);
        float o[3];
        mat_vec(M, B, G, R, o);
        const float cu = __fadd_rn(o[1], 0.5f), cv = __fadd_rn(o[2], 0.5f);
        yrow[dy][2 * q + dx] = cast_from_f32<OutT>(__fmul_rn(fminf(1.0f, o[0]), os));
        if (dy == 0 && dx == 0) { u = cu; v = cv; } else { u = __fadd_rn(u, cu); v = __fadd_rn(v, cv); }
      }
    // u / 4.0: power-of-two divisor, the product with 0.25 is the same correctly rounded value
    cu4[q] = cast_from_f32<OutT>(__fmul_rn(fminf(1.0f, __fmul_rn(u, 0.25f)), os));
    cv4[q] = cast_from_f32<OutT>(__fmul_rn(fminf(1.0f, __fmul_rn(v, 0.25f)), os));
  }
#pragma unroll
  for (int dy = 0; dy < 2; ++dy) st_bytes<8 * sizeof(OutT)>(yuv + (size_t)(2 * qy + dy) * W + 8 * gx, yrow[dy]);
  OutT* planes = yuv + (size_t)H * W;
  const size_t plane = (size_t)(H / 2) * (W / 2), idx = (size_t)qy * (W / 2) + 4 * gx;
  st_bytes<4 * sizeof(OutT)>(planes + plane + idx, cu4);      // plane 1 <- second matrix row
  st_bytes<4 * sizeof(OutT)>(planes + idx, cv4);              // plane 0 <- third matrix row
}

template <typename InT, typename OutT>
__global__ void __launch_bounds__(256) yuv420_rgb_vec_kernel(const InT* __restrict__ yuv, OutT* __restrict__ rgb, int H, int W, Mat3 Minv) {
  __shared__ float smem_lut[UnitScale<InT>::kTable ? 256 : 1];
  UnitScale<InT> unit;
  unit.init(smem_lut);
  const int gx = blockIdx.x * blockDim.x + threadIdx.x;     // pixel columns 8 gx .. 8 gx + 7 of row r
  const int r = blockIdx.y;
  if (gx >= W / 8) return;
  constexpr float os = DT<OutT>::scale;
  const InT* planes = yuv + (size_t)H * W;
  const size_t plane = (size_t)(H / 2) * (W / 2), idx = (size_t)(r / 2) * (W / 2) + 4 * gx;
  alignas(16) InT y8[8];
  alignas(16) InT cu4[4], cv4[4];
  alignas(16) OutT out[24];
  ld_bytes<8 * sizeof(InT)>(yuv + (size_t)r * W + 8 * gx, y8);
  ld_bytes<4 * sizeof(InT)>(planes + plane + idx, cu4);
  ld_bytes<4 * sizeof(InT)>(planes + idx, cv4);
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const float y = unit(y8[c]);
    const float cu = __fsub_rn(unit(cu4[c >> 1]), 0.5f);
    const float cv = __fsub_rn(unit(cv4[c >> 1]), 0.5f);
    float o[3];
    mat_vec(Minv, y, cu, cv, o);                            // = bgr
    out[3 * c] = cast_from_f32<OutT>(__fmul_rn(fminf(1.0f, o[2]), os));
    out[3 * c + 1] = cast_from_f32<OutT>(__fmul_rn(fminf(1.0f, o[1]), os));
    out[3 * c + 2] = cast_from_f32<OutT>(__fmul_rn(fminf(1.0f, o[0]), os));
  }
  st_bytes<24 * sizeof(OutT)>(rgb + ((size_t)r * W + 8 * gx) * 3, out);
}

static inline bool vec_ok(const void* a, const void* b, int width) {
  return width % 8 == 0 && ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b)) & 15u) == 0;
}

}  // namespace isp

using namespace isp;

// matrix9_host: 9 row-major floats on the host: YCrCb_T_bgr for the encoder, its inverse for the decoder (the
// Python side owns the constants, like the reference's module-level matrices yuv_420.py:12-18)
extern "C" int b200isp_rgb_yuv420(const void* rgb, int in_dtype, void* yuv, int out_dtype, int height, int width,
                                  const float* matrix9_host, b200isp_stream stream) {
  ISP_REQUIRE(rgb && yuv && matrix9_host, B200ISP_E_ARG, "rgb_yuv420: null pointer");
  ISP_REQUIRE(height >= 2 && width >= 2 && height % 2 == 0 && width % 2 == 0, B200ISP_E_SHAPE, "rgb_yuv420: even size needed, got %dx%d", height, width);
  ISP_REQUIRE(valid_dtype(in_dtype) && valid_dtype(out_dtype), B200ISP_E_DTYPE, "rgb_yuv420: bad dtype");
  Mat3 M;
  for (int i = 0; i < 9; ++i) M.m[i] = matrix9_host[i];
  cudaStream_t s = (cudaStream_t)stream;
  if (vec_ok(rgb, yuv, width)) {
    const dim3 grid((width / 8 + 127) / 128, height / 2);
    ISP_DISPATCH_DTYPE(in_dtype, InT, {
      ISP_DISPATCH_DTYPE(out_dtype, OutT, (rgb_yuv420_vec_kernel<InT, OutT><<<grid, 128, 0, s>>>((const InT*)rgb, (OutT*)yuv, height, width, M)));
    });
  } else {
    const dim3 grid((width / 2 + 255) / 256, height / 2);
    ISP_DISPATCH_DTYPE(in_dtype, InT, {
      ISP_DISPATCH_DTYPE(out_dtype, OutT, (rgb_yuv420_kernel<InT, OutT><<<grid, 256, 0, s>>>((const InT*)rgb, (OutT*)yuv, height, width, M)));
    });
  }
  ISP_LAUNCH_CHECK("rgb_yuv420_kernel");
  return B200ISP_OK;
}

extern "C" int b200isp_yuv420_rgb(const void* yuv, int in_dtype, void* rgb, int out_dtype, int height, int width,
                                  const float* matrix9_host, b200isp_stream stream) {
  ISP_REQUIRE(rgb && yuv && matrix9_host, B200ISP_E_ARG, "yuv420_rgb: null pointer");
  ISP_REQUIRE(height >= 2 && width >= 2 && height % 2 == 0 && width % 2 == 0, B200ISP_E_SHAPE, "yuv420_rgb: even size needed, got %dx%d", height, width);
  ISP_REQUIRE(valid_dtype(in_dtype) && valid_dtype(out_dtype), B200ISP_E_DTYPE, "yuv420_rgb: bad dtype");
  Mat3 M;
  for (int i = 0; i < 9; ++i) M.m[i] = matrix9_host[i];
  cudaStream_t s = (cudaStream_t)stream;
  if (vec_ok(rgb, yuv, width)) {
    const dim3 grid((width / 8 + 127) / 128, height);
    ISP_DISPATCH_DTYPE(in_dtype, InT, {
      ISP_DISPATCH_DTYPE(out_dtype, OutT, (yuv420_rgb_vec_kernel<InT, OutT><<<grid, 128, 0, s>>>((const InT*)yuv, (OutT*)rgb, height, width, M)));
    });
  } else {
    const dim3 grid((width + 255) / 256, height);
    ISP_DISPATCH_DTYPE(in_dtype, InT, {
      ISP_DISPATCH_DTYPE(out_dtype, OutT, (yuv420_rgb_kernel<InT, OutT><<<grid, 256, 0, s>>>((const InT*)yuv, (OutT*)rgb, height, width, M)));
    });
  }
  ISP_LAUNCH_CHECK("yuv420_rgb_kernel");
  return B200ISP_OK;
}
