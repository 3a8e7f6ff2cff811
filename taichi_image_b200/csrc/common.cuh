// Shared device/host helpers for the B200 (sm_100a) camera-ISP kernels.
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>
#include <stdlib.h>
#include <stddef.h>
#include <type_traits>
#include "../../include/b200isp.h"

namespace isp {

// ------------------------------------------------------------------ errors
void set_error(const char* fmt, ...);
int cuda_status(cudaError_t e, const char* what);

#define ISP_REQUIRE(cond, code, ...)            \
  do {                                          \
    if (!(cond)) {                              \
      isp::set_error(__VA_ARGS__);              \
      return (code);                            \
    }                                           \
  } while (0)

#define ISP_LAUNCH_CHECK(what)                                         \
  do {                                                                 \
    int _s = isp::cuda_status(cudaPeekAtLastError(), what);            \
    if (_s) return _s;                                                 \
  } while (0)

constexpr int kNumSMs = 148;   // B200

struct float9 { float v[9]; };

// ------------------------------------------------------------------ workspace layout
// Caller-owned, zero-initialised once; every kernel that uses a counter/flag restores it to 0.
struct Workspace {
  unsigned int counter[8];        // last-block-done tickets
  float bounds[2];                // phase-1 min/max (blended) for the metering kernels
  float frame_max[B200ISP_MAX_FRAMES];   // Reinhard per-frame max (bit pattern of a non-negative float)
  float frame_max2[B200ISP_MAX_FRAMES];  // the same for the frames the one-sweep u16 map declined (exact two-sweep fallback)
  float scratch[32];
  float partials[1];              // [kMaxPartialBlocks][kPartialStride] follows
};
constexpr int kMaxPartialBlocks = 2048;
constexpr int kPartialStride = 8;
constexpr size_t kWorkspaceBytes = sizeof(Workspace) + sizeof(float) * kMaxPartialBlocks * kPartialStride;

// ------------------------------------------------------------------ dtype traits (types.py:12-18)
template <typename T> struct DT;
template <> struct DT<uint8_t>  { static constexpr float scale = 255.f;   static constexpr bool is_int = true;  static constexpr int id = B200ISP_U8; };
template <> struct DT<uint16_t> { static constexpr float scale = 65535.f; static constexpr bool is_int = true;  static constexpr int id = B200ISP_U16; };
template <> struct DT<int16_t>  { static constexpr float scale = 32767.f; static constexpr bool is_int = true;  static constexpr int id = B200ISP_I16; };
template <> struct DT<__half>   { static constexpr float scale = 1.f;     static constexpr bool is_int = false; static constexpr int id = B200ISP_F16; };
template <> struct DT<float>    { static constexpr float scale = 1.f;     static constexpr bool is_int = false; static constexpr int id = B200ISP_F32; };

__device__ __forceinline__ float to_f32(uint8_t v)  { return (float)v; }
__device__ __forceinline__ float to_f32(uint16_t v) { return (float)v; }
__device__ __forceinline__ float to_f32(int16_t v)  { return (float)v; }
__device__ __forceinline__ float to_f32(__half v)   { return __half2float(v); }
__device__ __forceinline__ float to_f32(float v)    { return v; }

// ti.cast(float, T): truncation toward zero for integers (saturating where the reference is UB,
// NaN -> 0), round-to-nearest-even for f16.
template <typename T> __device__ __forceinline__ T cast_from_f32(float x);
template <> __device__ __forceinline__ uint8_t  cast_from_f32<uint8_t>(float x)  { return (uint8_t)min(__float2uint_rz(x), 255u); }
template <> __device__ __forceinline__ uint16_t cast_from_f32<uint16_t>(float x) { return (uint16_t)min(__float2uint_rz(x), 65535u); }
template <> __device__ __forceinline__ int16_t  cast_from_f32<int16_t>(float x)  { return (int16_t)max(min(__float2int_rz(x), 32767), -32768); }
template <> __device__ __forceinline__ __half   cast_from_f32<__half>(float x)   { return __float2half_rn(x); }
template <> __device__ __forceinline__ float    cast_from_f32<float>(float x)    { return x; }

// round a float through the ISP dtype (Camera16: f16 store + reload; Camera32: identity)
template <bool CAM16> __device__ __forceinline__ float round_isp(float x) {
  if constexpr (CAM16) return __half2float(__float2half_rn(x));
  else return x;
}

__device__ __forceinline__ float clamp01(float x) { return __saturatef(x); }

// color/__init__.py:7-10 (separate roundings, left to right, like the oracle)
__device__ __forceinline__ float rgb_gray(float r, float g, float b) {
  return __fadd_rn(__fadd_rn(__fmul_rn(r, 0.299f), __fmul_rn(g, 0.587f)), __fmul_rn(b, 0.114f));
}

// ------------------------------------------------------------------ reductions
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// (shifted & mask) | one in ONE LOP3 (the compiler splits the two immediates into two): drops an integer field into the
// mantissa of the float `one` -- the conversion-free sample decode of the sweeps
__device__ __forceinline__ float biased_from_shifted(uint32_t shifted, uint32_t mask, uint32_t one) {
  uint32_t r;
  asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(r) : "r"(shifted), "r"(mask), "r"(one));
  return __uint_as_float(r);
}

// dispatch helpers -------------------------------------------------------------------------
#define ISP_DISPATCH_DTYPE(dt, T, ...)                                   \
  switch (dt) {                                                          \
    case B200ISP_U8:  { using T = uint8_t;  __VA_ARGS__; break; }        \
    case B200ISP_U16: { using T = uint16_t; __VA_ARGS__; break; }        \
    case B200ISP_I16: { using T = int16_t;  __VA_ARGS__; break; }        \
    case B200ISP_F16: { using T = __half;   __VA_ARGS__; break; }        \
    case B200ISP_F32: { using T = float;    __VA_ARGS__; break; }        \
    default: isp::set_error("unsupported dtype %d", (int)(dt)); return B200ISP_E_DTYPE; \
  }

inline bool valid_dtype(int dt) { return dt >= B200ISP_U8 && dt <= B200ISP_F32; }
inline size_t dtype_size(int dt) { return dt == B200ISP_U8 ? 1 : (dt == B200ISP_F32 ? 4 : 2); }

// NB bytes between global memory and a 16-byte aligned local buffer with the widest access NB allows (16 / 8 / 4)
template <int NB> __device__ __forceinline__ void ld_bytes(const void* g, void* local) {
  if constexpr (NB % 16 == 0) {
#pragma unroll
    for (int i = 0; i < NB / 16; ++i) reinterpret_cast<uint4*>(local)[i] = __ldg(reinterpret_cast<const uint4*>(g) + i);
  } else if constexpr (NB % 8 == 0) {
#pragma unroll
    for (int i = 0; i < NB / 8; ++i) reinterpret_cast<uint2*>(local)[i] = __ldg(reinterpret_cast<const uint2*>(g) + i);
  } else {
#pragma unroll
    for (int i = 0; i < NB / 4; ++i) reinterpret_cast<uint32_t*>(local)[i] = __ldg(reinterpret_cast<const uint32_t*>(g) + i);
  }
}
template <int NB> __device__ __forceinline__ void st_bytes(void* g, const void* local) {
  if constexpr (NB % 16 == 0) {
#pragma unroll
    for (int i = 0; i < NB / 16; ++i) reinterpret_cast<uint4*>(g)[i] = reinterpret_cast<const uint4*>(local)[i];
  } else if constexpr (NB % 8 == 0) {
#pragma unroll
    for (int i = 0; i < NB / 8; ++i) reinterpret_cast<uint2*>(g)[i] = reinterpret_cast<const uint2*>(local)[i];
  } else {
#pragma unroll
    for (int i = 0; i < NB / 4; ++i) reinterpret_cast<uint32_t*>(g)[i] = reinterpret_cast<const uint32_t*>(local)[i];
  }
}

}  // namespace isp
