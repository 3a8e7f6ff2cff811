// 12-bit pack / unpack and raw-16 decode (reference: packed.py:12-210).
//
// Layouts (3 bytes <-> 2 pixels), little-endian 24-bit group x = b0 | b1<<8 | b2<<16:
//   standard (packed.py:12-31): p0 = x & 0xFFF,              p1 = x >> 12
//   IDS      (packed.py:36-55): p0 = b0<<4 | (b2 & 0xF),     p1 = b1<<4 | b2>>4   (decode)
//                               b0 = p0>>4, b1 = p1>>4, b2 = (p0&0xF)<<4 | (p1&0xF) (encode; as
//                               written the reference's IDS encode is NOT the inverse of its decode)
// Vector path: one thread converts kPackNpx pixels (8: 12 packed bytes, three words).
#include "common.cuh"

namespace isp {

template <bool IDS> __device__ __forceinline__ void decode_pair(uint32_t x, uint32_t& p0, uint32_t& p1) {
  if constexpr (!IDS) {
    p0 = x & 0xFFFu;
    p1 = (x >> 12) & 0xFFFu;
  } else {
    const uint32_t b0 = x & 0xFFu, b1 = (x >> 8) & 0xFFu, b2 = (x >> 16) & 0xFFu;
    p0 = (b0 << 4) | (b2 & 0xFu);
    p1 = (b1 << 4) | (b2 >> 4);
  }
}

template <bool IDS> __device__ __forceinline__ uint32_t encode_pair(uint32_t p0, uint32_t p1) {
  uint32_t b0, b1, b2;
  if constexpr (!IDS) {
    b0 = p0 & 0xFFu;
    b1 = (((p1 & 0xFu) << 4) | (p0 >> 8)) & 0xFFu;
    b2 = (p1 >> 4) & 0xFFu;
  } else {
    b0 = (p0 >> 4) & 0xFFu;
    b1 = (p1 >> 4) & 0xFFu;
    b2 = (((p0 & 0xFu) << 4) | (p1 & 0xFu)) & 0xFFu;
  }
  return b0 | (b1 << 8) | (b2 << 16);
}

// packed.py:98-104 write_value_{scaled,direct}
template <typename T, bool SCALED> __device__ __forceinline__ T decoded_value(uint32_t v, float k) {
  if constexpr (SCALED) return cast_from_f32<T>(__fmul_rn((float)v, k));
  else if constexpr (DT<T>::is_int) return (T)v;
  else return cast_from_f32<T>((float)v);
}

// packed.py:66-73 read_value_{scaled,direct}
template <typename T, bool SCALED> __device__ __forceinline__ uint32_t value_to_u12(T v, float k) {
  if constexpr (SCALED) {
    const float r = roundf(__fmul_rn(to_f32(v), k));      // ti.round: half away from zero
    return (uint32_t)(uint16_t)(int)r;
  } else if constexpr (DT<T>::is_int) {
    return (uint32_t)(uint16_t)v;
  } else {
    return (uint32_t)(uint16_t)(int)to_f32(v);
  }
}

// ---------------------------------------------------------------- decode12
// NPX pixels per thread (ISP_PACK_NPX, a multiple of 8): 8 keeps a warp's stores contiguous (one STG.128 per thread
// for u16 / f16 outputs) -- measured against 32 (three 16-byte loads, but stores 64..128 bytes apart per thread).
#ifndef ISP_PACK_NPX
#define ISP_PACK_NPX 8
#endif
constexpr int kPackNpx = ISP_PACK_NPX;
static_assert(kPackNpx % 8 == 0, "whole 12-byte groups");

template <typename T, bool SCALED, bool IDS>
__global__ void __launch_bounds__(256) decode12_vec_kernel(const uint8_t* __restrict__ enc, T* __restrict__ out,
                                                           int64_t n_groups, float k) {
  constexpr int NW = kPackNpx * 3 / 8;
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n_groups) return;
  alignas(16) uint32_t w[NW];
  ld_bytes<4 * NW>(enc + g * (4 * NW), w);
  alignas(16) T v[kPackNpx];
#pragma unroll
  for (int i = 0; i < kPackNpx / 2; ++i) {
    const int k0 = (3 * i) / 4, off = (24 * i) % 32;
    const uint32_t x = (off == 0 ? w[k0] : (off == 8 ? (w[k0] >> 8) : __funnelshift_r(w[k0], w[k0 + 1 < NW ? k0 + 1 : k0], off))) & 0xFFFFFFu;
    uint32_t p0, p1;
    decode_pair<IDS>(x, p0, p1);
    v[2 * i] = decoded_value<T, SCALED>(p0, k);
    v[2 * i + 1] = decoded_value<T, SCALED>(p1, k);
  }
  st_bytes<kPackNpx * (int)sizeof(T)>(out + g * kPackNpx, v);
}

template <typename T, bool SCALED, bool IDS>
__global__ void __launch_bounds__(256) decode12_pair_kernel(const uint8_t* __restrict__ enc, T* __restrict__ out,
                                                            int64_t pair_begin, int64_t n_pairs, float k) {
  const int64_t i = pair_begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_pairs) return;
  const uint8_t* b = enc + 3 * i;
  const uint32_t x = (uint32_t)b[0] | ((uint32_t)b[1] << 8) | ((uint32_t)b[2] << 16);
  uint32_t p0, p1;
  decode_pair<IDS>(x, p0, p1);
  out[2 * i] = decoded_value<T, SCALED>(p0, k);
  out[2 * i + 1] = decoded_value<T, SCALED>(p1, k);
}

// ---------------------------------------------------------------- encode12
template <typename T, bool SCALED, bool IDS>
__global__ void __launch_bounds__(256) encode12_vec_kernel(const T* __restrict__ values, uint8_t* __restrict__ enc,
                                                           int64_t n_groups, float k) {
  constexpr int NW = kPackNpx * 3 / 8;
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n_groups) return;
  alignas(16) T v[kPackNpx];
  ld_bytes<kPackNpx * (int)sizeof(T)>(values + g * kPackNpx, v);
  alignas(16) uint32_t w[NW];
#pragma unroll
  for (int i = 0; i < NW; ++i) w[i] = 0u;
#pragma unroll
  for (int i = 0; i < kPackNpx / 2; ++i) {
    const uint32_t x = encode_pair<IDS>(value_to_u12<T, SCALED>(v[2 * i], k), value_to_u12<T, SCALED>(v[2 * i + 1], k));
    const int k0 = (3 * i) / 4, off = (24 * i) % 32;
    w[k0] |= x << off;
    if (off > 8) w[k0 + 1 < NW ? k0 + 1 : k0] |= x >> (32 - off);
  }
  st_bytes<4 * NW>(enc + g * (4 * NW), w);
}

template <typename T, bool SCALED, bool IDS>
__global__ void __launch_bounds__(256) encode12_pair_kernel(const T* __restrict__ values, uint8_t* __restrict__ enc,
                                                            int64_t pair_begin, int64_t n_pairs, float k) {
  const int64_t i = pair_begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_pairs) return;
  const uint32_t x = encode_pair<IDS>(value_to_u12<T, SCALED>(values[2 * i], k), value_to_u12<T, SCALED>(values[2 * i + 1], k));
  enc[3 * i] = (uint8_t)x;
  enc[3 * i + 1] = (uint8_t)(x >> 8);
  enc[3 * i + 2] = (uint8_t)(x >> 16);
}

// ---------------------------------------------------------------- decode16 (packed.py:149-157)
template <typename T, bool SCALED>
__global__ void __launch_bounds__(256) decode16_kernel(const uint8_t* __restrict__ enc, T* __restrict__ out,
                                                       int64_t n, float k, bool vec) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (vec) {     // 8 values per thread, 16-byte load
    const int64_t i = t * 8;
    if (i + 8 <= n) {
      const uint4 q = __ldg(reinterpret_cast<const uint4*>(enc) + t);
      const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const uint32_t v = (w[j >> 1] >> (16 * (j & 1))) & 0xFFFFu;   // little-endian byte pair
        out[i + j] = decoded_value<T, SCALED>(v, k);
      }
      return;
    }
    for (int64_t j = i; j < n; ++j) {
      const uint32_t v = (uint32_t)enc[2 * j] | ((uint32_t)enc[2 * j + 1] << 8);
      out[j] = decoded_value<T, SCALED>(v, k);
    }
  } else if (t < n) {
    const uint32_t v = (uint32_t)enc[2 * t] | ((uint32_t)enc[2 * t + 1] << 8);
    out[t] = decoded_value<T, SCALED>(v, k);
  }
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

template <typename T, bool SCALED, bool IDS>
static int launch_decode12(const uint8_t* enc, T* out, int64_t n_values, float k, cudaStream_t s) {
  const int64_t n_pairs = n_values / 2;
  int64_t n_groups = (aligned16(enc) && aligned16(out)) ? n_values / kPackNpx : 0;
  if (n_groups > 0) {
    decode12_vec_kernel<T, SCALED, IDS><<<(unsigned)((n_groups + 255) / 256), 256, 0, s>>>(enc, out, n_groups, k);
    ISP_LAUNCH_CHECK("decode12_vec_kernel");
  }
  const int64_t rest = n_pairs - n_groups * (kPackNpx / 2);
  if (rest > 0) {
    decode12_pair_kernel<T, SCALED, IDS><<<(unsigned)((rest + 255) / 256), 256, 0, s>>>(enc, out, n_groups * (kPackNpx / 2), n_pairs, k);
    ISP_LAUNCH_CHECK("decode12_pair_kernel");
  }
  return B200ISP_OK;
}

template <typename T, bool SCALED, bool IDS>
static int launch_encode12(const T* values, uint8_t* enc, int64_t n_values, float k, cudaStream_t s) {
  const int64_t n_pairs = n_values / 2;
  int64_t n_groups = (aligned16(enc) && aligned16(values)) ? n_values / kPackNpx : 0;
  if (n_groups > 0) {
    encode12_vec_kernel<T, SCALED, IDS><<<(unsigned)((n_groups + 255) / 256), 256, 0, s>>>(values, enc, n_groups, k);
    ISP_LAUNCH_CHECK("encode12_vec_kernel");
  }
  const int64_t rest = n_pairs - n_groups * (kPackNpx / 2);
  if (rest > 0) {
    encode12_pair_kernel<T, SCALED, IDS><<<(unsigned)((rest + 255) / 256), 256, 0, s>>>(values, enc, n_groups * (kPackNpx / 2), n_pairs, k);
    ISP_LAUNCH_CHECK("encode12_pair_kernel");
  }
  return B200ISP_OK;
}

}  // namespace isp

namespace isp {
// IDS 12-bit layout -> standard layout, byte stream to byte stream (packed.py:36-44 followed by :12-20 without
// leaving registers): lets IDS frames take the fused sweep, whose row loader reads the standard layout.
//   IDS:       p0 = b0 << 4 | (b2 & 0xF),  p1 = b1 << 4 | b2 >> 4
//   standard:  c0 = p0 & 0xFF,  c1 = (p1 & 0xF) << 4 | p0 >> 8,  c2 = p1 >> 4
__device__ __forceinline__ void ids_to_std3(uint32_t b0, uint32_t b1, uint32_t b2, uint32_t& c0, uint32_t& c1, uint32_t& c2) {
  const uint32_t p0 = (b0 << 4) | (b2 & 0xFu), p1 = (b1 << 4) | (b2 >> 4);
  c0 = p0 & 0xFFu; c1 = ((p1 & 0xFu) << 4) | (p0 >> 8); c2 = p1 >> 4;
}

// 12 bytes (4 triplets = 8 pixels) per thread as three aligned words; the tail triplets byte-wise
__global__ void __launch_bounds__(256) repack12_ids_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst,
                                                           long long n_triplets, bool words) {
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nth = (long long)gridDim.x * blockDim.x;
  const long long n4 = words ? n_triplets / 4 : 0;
  for (long long i = tid; i < n4; i += nth) {
    const uint32_t* p = reinterpret_cast<const uint32_t*>(src) + 3 * i;
    const uint32_t w[3] = {__ldg(p), __ldg(p + 1), __ldg(p + 2)};
    uint32_t o[3] = {0u, 0u, 0u};
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      uint32_t b[3], c[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) b[k] = (w[(3 * t + k) >> 2] >> (8 * ((3 * t + k) & 3))) & 0xFFu;
      ids_to_std3(b[0], b[1], b[2], c[0], c[1], c[2]);
#pragma unroll
      for (int k = 0; k < 3; ++k) o[(3 * t + k) >> 2] |= c[k] << (8 * ((3 * t + k) & 3));
    }
    uint32_t* q = reinterpret_cast<uint32_t*>(dst) + 3 * i;
    q[0] = o[0]; q[1] = o[1]; q[2] = o[2];
  }
  for (long long i = 4 * n4 + tid; i < n_triplets; i += nth) {
    uint32_t c0, c1, c2;
    ids_to_std3(src[3 * i], src[3 * i + 1], src[3 * i + 2], c0, c1, c2);
    dst[3 * i] = (uint8_t)c0; dst[3 * i + 1] = (uint8_t)c1; dst[3 * i + 2] = (uint8_t)c2;
  }
}
}  // namespace isp

using namespace isp;

extern "C" int b200isp_repack12_ids(const uint8_t* ids, uint8_t* standard, int64_t n_bytes, b200isp_stream stream) {
  ISP_REQUIRE(n_bytes >= 0 && n_bytes % 3 == 0, B200ISP_E_SHAPE, "repack12_ids: byte count must be a multiple of 3, got %lld", (long long)n_bytes);
  if (n_bytes == 0) return B200ISP_OK;
  ISP_REQUIRE(ids && standard, B200ISP_E_ARG, "repack12_ids: null pointer");
  const long long n = n_bytes / 3;
  const bool words = ((reinterpret_cast<uintptr_t>(ids) | reinterpret_cast<uintptr_t>(standard)) & 3u) == 0;
  long long blocks = (n / 4 + 255) / 256;
  if (blocks < 1) blocks = 1;
  if (blocks > 8 * kNumSMs) blocks = 8 * kNumSMs;
  repack12_ids_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(ids, standard, n, words);
  ISP_LAUNCH_CHECK("repack12_ids_kernel");
  return B200ISP_OK;
}

extern "C" int b200isp_decode12(const uint8_t* encoded, int64_t n_values, void* out, int out_dtype,
                                int scaled, int ids_format, b200isp_stream stream) {
  ISP_REQUIRE(n_values >= 0 && n_values % 2 == 0, B200ISP_E_SHAPE, "decode12: n_values must be even, got %lld", (long long)n_values);
  if (n_values == 0) return B200ISP_OK;
  ISP_REQUIRE(encoded && out, B200ISP_E_ARG, "decode12: null pointer");
  cudaStream_t s = (cudaStream_t)stream;
  ISP_DISPATCH_DTYPE(out_dtype, T, {
    const float k = (float)((double)DT<T>::scale / 4095.0);   // packed.py:99 (python double -> f32 constant)
    T* o = (T*)out;
    if (scaled) return ids_format ? launch_decode12<T, true, true>(encoded, o, n_values, k, s)
                                  : launch_decode12<T, true, false>(encoded, o, n_values, k, s);
    return ids_format ? launch_decode12<T, false, true>(encoded, o, n_values, k, s)
                      : launch_decode12<T, false, false>(encoded, o, n_values, k, s);
  });
  return B200ISP_OK;
}

extern "C" int b200isp_encode12(const void* values, int in_dtype, int64_t n_values, uint8_t* encoded,
                                int scaled, int ids_format, b200isp_stream stream) {
  ISP_REQUIRE(n_values >= 0 && n_values % 2 == 0, B200ISP_E_SHAPE, "encode12: n_values must be even, got %lld", (long long)n_values);
  if (n_values == 0) return B200ISP_OK;
  ISP_REQUIRE(encoded && values, B200ISP_E_ARG, "encode12: null pointer");
  cudaStream_t s = (cudaStream_t)stream;
  ISP_DISPATCH_DTYPE(in_dtype, T, {
    const float k = (float)(4095.0 / (double)DT<T>::scale);   // packed.py:67
    const T* v = (const T*)values;
    if (scaled) return ids_format ? launch_encode12<T, true, true>(v, encoded, n_values, k, s)
                                  : launch_encode12<T, true, false>(v, encoded, n_values, k, s);
    return ids_format ? launch_encode12<T, false, true>(v, encoded, n_values, k, s)
                      : launch_encode12<T, false, false>(v, encoded, n_values, k, s);
  });
  return B200ISP_OK;
}

extern "C" int b200isp_decode16(const uint8_t* encoded, int64_t n_values, void* out, int out_dtype,
                                int scaled, b200isp_stream stream) {
  ISP_REQUIRE(n_values >= 0, B200ISP_E_SHAPE, "decode16: negative size");
  if (n_values == 0) return B200ISP_OK;
  ISP_REQUIRE(encoded && out, B200ISP_E_ARG, "decode16: null pointer");
  cudaStream_t s = (cudaStream_t)stream;
  const bool vec = aligned16(encoded);
  const int64_t threads = vec ? (n_values + 7) / 8 : n_values;
  const unsigned blocks = (unsigned)((threads + 255) / 256);
  ISP_DISPATCH_DTYPE(out_dtype, T, {
    const float k = (float)((double)DT<T>::scale / 65535.0);  // packed.py:140
    if (scaled) decode16_kernel<T, true><<<blocks, 256, 0, s>>>(encoded, (T*)out, n_values, k, vec);
    else        decode16_kernel<T, false><<<blocks, 256, 0, s>>>(encoded, (T*)out, n_values, k, vec);
  });
  ISP_LAUNCH_CHECK("decode16_kernel");
  return B200ISP_OK;
}

// ================================================================ 10-bit packed (EXTENSION; no reference counterpart)
// MIPI CSI-2 RAW10: 5 bytes <-> 4 pixels; bytes 0..3 hold bits 9..2 of pixels 0..3, byte 4 their bits 1..0 (pixel 0 in the
// lowest bit pair).  Value conventions are those of the 12-bit functions (packed.py:66-73, :98-104) with 1023 for 4095:
// scaled decode = v * (scale_T / 1023), scaled encode = round(v * (1023 / scale_T)); the code is truncated to 10 bits.
// Vector path: one thread converts 16 pixels = 20 packed bytes (five words; inputs 4-byte, outputs 16-byte aligned).
namespace isp {

__device__ __forceinline__ uint32_t byte_of(const uint32_t (&w)[5], int b) { return (w[b >> 2] >> (8 * (b & 3))) & 0xFFu; }

template <typename T, bool SCALED>
__global__ void __launch_bounds__(256) decode10_vec_kernel(const uint8_t* __restrict__ enc, T* __restrict__ out, int64_t n_groups, float k) {
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n_groups) return;
  uint32_t w[5];
#pragma unroll
  for (int i = 0; i < 5; ++i) w[i] = __ldg(reinterpret_cast<const uint32_t*>(enc + g * 20) + i);
  alignas(16) T v[16];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const uint32_t lsb = byte_of(w, 5 * q + 4);
#pragma unroll
    for (int j = 0; j < 4; ++j) v[4 * q + j] = decoded_value<T, SCALED>((byte_of(w, 5 * q + j) << 2) | ((lsb >> (2 * j)) & 3u), k);
  }
  st_bytes<16 * (int)sizeof(T)>(out + g * 16, v);
}

template <typename T, bool SCALED>
__global__ void __launch_bounds__(256) decode10_quad_kernel(const uint8_t* __restrict__ enc, T* __restrict__ out, int64_t quad_begin,
                                                            int64_t n_quads, float k) {
  const int64_t i = quad_begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_quads) return;
  const uint8_t* b = enc + 5 * i;
  const uint32_t lsb = b[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) out[4 * i + j] = decoded_value<T, SCALED>(((uint32_t)b[j] << 2) | ((lsb >> (2 * j)) & 3u), k);
}

template <typename T, bool SCALED>
__global__ void __launch_bounds__(256) encode10_vec_kernel(const T* __restrict__ values, uint8_t* __restrict__ enc, int64_t n_groups, float k) {
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n_groups) return;
  alignas(16) T v[16];
  ld_bytes<16 * (int)sizeof(T)>(values + g * 16, v);
  uint32_t w[5] = {0u, 0u, 0u, 0u, 0u};
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    uint32_t lsb = 0u;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint32_t p = value_to_u12<T, SCALED>(v[4 * q + j], k) & 0x3FFu;
      const int b = 5 * q + j;
      w[b >> 2] |= (p >> 2) << (8 * (b & 3));
      lsb |= (p & 3u) << (2 * j);
    }
    const int b = 5 * q + 4;
    w[b >> 2] |= lsb << (8 * (b & 3));
  }
#pragma unroll
  for (int i = 0; i < 5; ++i) reinterpret_cast<uint32_t*>(enc + g * 20)[i] = w[i];
}

template <typename T, bool SCALED>
__global__ void __launch_bounds__(256) encode10_quad_kernel(const T* __restrict__ values, uint8_t* __restrict__ enc, int64_t quad_begin,
                                                            int64_t n_quads, float k) {
  const int64_t i = quad_begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_quads) return;
  uint32_t lsb = 0u;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint32_t p = value_to_u12<T, SCALED>(values[4 * i + j], k) & 0x3FFu;
    enc[5 * i + j] = (uint8_t)(p >> 2);
    lsb |= (p & 3u) << (2 * j);
  }
  enc[5 * i + 4] = (uint8_t)lsb;
}

template <typename T, bool SCALED>
static int launch_decode10(const uint8_t* enc, T* out, int64_t n_values, float k, cudaStream_t s) {
  const int64_t n_quads = n_values / 4;
  const int64_t n_groups = ((reinterpret_cast<uintptr_t>(enc) & 3u) == 0 && aligned16(out)) ? n_values / 16 : 0;
  if (n_groups > 0) {
    decode10_vec_kernel<T, SCALED><<<(unsigned)((n_groups + 255) / 256), 256, 0, s>>>(enc, out, n_groups, k);
    ISP_LAUNCH_CHECK("decode10_vec_kernel");
  }
  const int64_t rest = n_quads - 4 * n_groups;
  if (rest > 0) {
    decode10_quad_kernel<T, SCALED><<<(unsigned)((rest + 255) / 256), 256, 0, s>>>(enc, out, 4 * n_groups, n_quads, k);
    ISP_LAUNCH_CHECK("decode10_quad_kernel");
  }
  return B200ISP_OK;
}

template <typename T, bool SCALED>
static int launch_encode10(const T* values, uint8_t* enc, int64_t n_values, float k, cudaStream_t s) {
  const int64_t n_quads = n_values / 4;
  const int64_t n_groups = ((reinterpret_cast<uintptr_t>(enc) & 3u) == 0 && aligned16(values)) ? n_values / 16 : 0;
  if (n_groups > 0) {
    encode10_vec_kernel<T, SCALED><<<(unsigned)((n_groups + 255) / 256), 256, 0, s>>>(values, enc, n_groups, k);
    ISP_LAUNCH_CHECK("encode10_vec_kernel");
  }
  const int64_t rest = n_quads - 4 * n_groups;
  if (rest > 0) {
    encode10_quad_kernel<T, SCALED><<<(unsigned)((rest + 255) / 256), 256, 0, s>>>(values, enc, 4 * n_groups, n_quads, k);
    ISP_LAUNCH_CHECK("encode10_quad_kernel");
  }
  return B200ISP_OK;
}

}  // namespace isp

extern "C" int b200isp_decode10(const uint8_t* encoded, int64_t n_values, void* out, int out_dtype, int scaled, b200isp_stream stream) {
  ISP_REQUIRE(n_values >= 0 && n_values % 4 == 0, B200ISP_E_SHAPE, "decode10: n_values must be a multiple of 4, got %lld", (long long)n_values);
  if (n_values == 0) return B200ISP_OK;
  ISP_REQUIRE(encoded && out, B200ISP_E_ARG, "decode10: null pointer");
  cudaStream_t s = (cudaStream_t)stream;
  ISP_DISPATCH_DTYPE(out_dtype, T, {
    const float k = (float)((double)DT<T>::scale / 1023.0);
    if (scaled) return isp::launch_decode10<T, true>(encoded, (T*)out, n_values, k, s);
    return isp::launch_decode10<T, false>(encoded, (T*)out, n_values, k, s);
  });
  return B200ISP_OK;
}

extern "C" int b200isp_encode10(const void* values, int in_dtype, int64_t n_values, uint8_t* encoded, int scaled, b200isp_stream stream) {
  ISP_REQUIRE(n_values >= 0 && n_values % 4 == 0, B200ISP_E_SHAPE, "encode10: n_values must be a multiple of 4, got %lld", (long long)n_values);
  if (n_values == 0) return B200ISP_OK;
  ISP_REQUIRE(encoded && values, B200ISP_E_ARG, "encode10: null pointer");
  cudaStream_t s = (cudaStream_t)stream;
  ISP_DISPATCH_DTYPE(in_dtype, T, {
    const float k = (float)(1023.0 / (double)DT<T>::scale);
    if (scaled) return isp::launch_encode10<T, true>((const T*)values, encoded, n_values, k, s);
    return isp::launch_encode10<T, false>((const T*)values, encoded, n_values, k, s);
  });
  return B200ISP_OK;
}
