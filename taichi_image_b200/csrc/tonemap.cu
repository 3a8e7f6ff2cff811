// Eager (per-stage) tone-mapping kernels on materialised float RGB images.
//   b200isp_bounds              util.py:49-60 bounds_func
//   b200isp_linear              tonemap.py:11-17 linear_func (tonemap.linear_kernel and ISP linear_kernel)
//   b200isp_reinhard_standalone tonemap.py:134-168 (five dependent passes, 4 launches, no host sync)
//   b200isp_metering_update     camera_isp.py:142-175
//   b200isp_isp_reinhard        camera_isp.py:177-218
//   b200isp_load_convert        camera_isp.py:82-99
#include "metering.cuh"
#include "reinhard.cuh"

namespace isp {

static __host__ __device__ inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ---------------------------------------------------------------- 4-element vector access
template <typename T> __device__ __forceinline__ void load4(const T* p, float (&v)[4]) {
  if constexpr (sizeof(T) == 4) {
    const float4 q = *reinterpret_cast<const float4*>(p);
    v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
  } else if constexpr (sizeof(T) == 2) {
    const uint2 q = *reinterpret_cast<const uint2*>(p);
    const T* t = reinterpret_cast<const T*>(&q);
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = to_f32(t[i]);
  } else {
    const uint32_t q = *reinterpret_cast<const uint32_t*>(p);
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = (float)((q >> (8 * i)) & 0xFFu);
  }
}
template <typename T> __device__ __forceinline__ void store4(T* p, const T (&v)[4]) {
  if constexpr (sizeof(T) == 4) *reinterpret_cast<uint4*>(p) = *reinterpret_cast<const uint4*>(v);
  else if constexpr (sizeof(T) == 2) *reinterpret_cast<uint2*>(p) = *reinterpret_cast<const uint2*>(v);
  else *reinterpret_cast<uint32_t*>(p) = *reinterpret_cast<const uint32_t*>(v);
}

// ---------------------------------------------------------------- bounds (util.py:49-60)
template <typename T>
__global__ void __launch_bounds__(256) bounds_kernel(const T* __restrict__ x, long long n, bool vec, float* out, Workspace* ws) {
  __shared__ float smem[8 * 2];
  float v[2] = {INFINITY, -INFINITY};
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nth = (long long)gridDim.x * blockDim.x;
  if (vec) {
    const long long n4 = n / 4;
    for (long long i = tid; i < n4; i += nth) {
      float a[4];
      load4<T>(x + 4 * i, a);
      v[0] = fminf(v[0], fminf(fminf(a[0], a[1]), fminf(a[2], a[3])));
      v[1] = fmaxf(v[1], fmaxf(fmaxf(a[0], a[1]), fmaxf(a[2], a[3])));
    }
    for (long long i = n4 * 4 + tid; i < n; i += nth) { const float a = to_f32(x[i]); v[0] = fminf(v[0], a); v[1] = fmaxf(v[1], a); }
  } else {
    for (long long i = tid; i < n; i += nth) { const float a = to_f32(x[i]); v[0] = fminf(v[0], a); v[1] = fmaxf(v[1], a); }
  }
  const int op[2] = {0, 1};
  block_fold<2>(v, op, smem);
  if (threadIdx.x == 0) {
    ws->partials[blockIdx.x * kPartialStride + 0] = v[0];
    ws->partials[blockIdx.x * kPartialStride + 1] = v[1];
  }
  if (last_block_ticket(&ws->counter[0])) {
    float f[2] = {INFINITY, -INFINITY};
    for (int b = threadIdx.x; b < (int)gridDim.x; b += blockDim.x) {
      f[0] = fminf(f[0], __ldcg(&ws->partials[b * kPartialStride + 0]));
      f[1] = fmaxf(f[1], __ldcg(&ws->partials[b * kPartialStride + 1]));
    }
    block_fold<2>(f, op, smem);
    if (threadIdx.x == 0) { out[0] = f[0]; out[1] = f[1]; }
  }
}

static inline int reduce_grid(long long n_items_per_thread_units) {
  long long b = (n_items_per_thread_units + 256 * 8 - 1) / (256 * 8);
  if (b < 1) b = 1;
  if (b > 8 * kNumSMs) b = 8 * kNumSMs;
  return (int)b;
}

template <typename T>
static int launch_bounds_kernel(const T* x, long long n, float* out, Workspace* ws, cudaStream_t s) {
  const bool vec = aligned16(x);
  bounds_kernel<T><<<reduce_grid(vec ? n / 4 : n), 256, 0, s>>>(x, n, vec, out, ws);
  return cuda_status(cudaPeekAtLastError(), "bounds_kernel");
}

// ---------------------------------------------------------------- linear_func (tonemap.py:11-17)
template <typename InT, typename OutT>
__device__ __forceinline__ OutT linear_value(float x, float bmin, float inv_range, float inv_gamma, bool has_gamma) {
  float y = __fmul_rn(__fsub_rn(x, bmin), inv_range);
  if (has_gamma) y = powf(y, inv_gamma);
  return cast_from_f32<OutT>(__fmul_rn(clamp01(y), DT<OutT>::scale));   // NaN -> 0 through the saturating clamp
}

template <typename InT, typename OutT>
__global__ void __launch_bounds__(256) linear_kernel(const InT* __restrict__ src, OutT* __restrict__ dst, long long n,
                                                     const float* __restrict__ bounds, float gamma, bool vec) {
  const float bmin = bounds[0], bmax = bounds[1];
  const float inv_range = __fdiv_rn(1.0f, __fsub_rn(bmax, bmin));
  const float inv_gamma = __fdiv_rn(1.0f, gamma);
  const bool has_gamma = gamma != 1.0f;
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (vec) {
    const long long i = tid * 4;
    if (i + 4 <= n) {
      float a[4];
      load4<InT>(src + i, a);
      alignas(16) OutT o[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) o[k] = linear_value<InT, OutT>(a[k], bmin, inv_range, inv_gamma, has_gamma);
      store4<OutT>(dst + i, o);
    } else {
      for (long long j = i; j < n; ++j) dst[j] = linear_value<InT, OutT>(to_f32(src[j]), bmin, inv_range, inv_gamma, has_gamma);
    }
  } else if (tid < n) {
    dst[tid] = linear_value<InT, OutT>(to_f32(src[tid]), bmin, inv_range, inv_gamma, has_gamma);
  }
}

template <typename InT, typename OutT>
static int launch_linear(const InT* src, OutT* dst, long long n, const float* bounds, float gamma, cudaStream_t s) {
  const bool vec = aligned16(src) && aligned16(dst);
  const long long threads = vec ? (n + 3) / 4 : n;
  linear_kernel<InT, OutT><<<(unsigned)((threads + 255) / 256), 256, 0, s>>>(src, dst, n, bounds, gamma, vec);
  return cuda_status(cudaPeekAtLastError(), "linear_kernel");
}

// 24 consecutive elements (8 RGB pixels) with 16-byte accesses (8-byte accesses for 1-byte types)
template <typename T> __device__ __forceinline__ void load24(const T* p, float (&v)[24]) {
  if constexpr (sizeof(T) == 1) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const uint2 w = reinterpret_cast<const uint2*>(p)[i];
      const T* t = reinterpret_cast<const T*>(&w);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[i * 8 + j] = to_f32(t[j]);
    }
  } else {
    constexpr int PER = 16 / (int)sizeof(T);               // elements per 16 bytes
    const uint4* q = reinterpret_cast<const uint4*>(p);
#pragma unroll
    for (int i = 0; i < 24 / PER; ++i) {
      const uint4 w = q[i];
      const T* t = reinterpret_cast<const T*>(&w);
#pragma unroll
      for (int j = 0; j < PER; ++j) v[i * PER + j] = to_f32(t[j]);
    }
  }
}
template <typename T> __device__ __forceinline__ void store24(T* p, const float (&v)[24]) {
  constexpr int PER = 16 / (int)sizeof(T);
  uint4* q = reinterpret_cast<uint4*>(p);
#pragma unroll
  for (int i = 0; i < 24 / PER; ++i) {
    alignas(16) T t[PER];
#pragma unroll
    for (int j = 0; j < PER; ++j) t[j] = cast_from_f32<T>(v[i * PER + j]);
    q[i] = *reinterpret_cast<const uint4*>(t);
  }
}

// ---------------------------------------------------------------- stand-alone Reinhard (tonemap.py:134-168)
// The reference runs five dependent passes over an f32 temp image (normalise -> meter -> map in place -> bounds -> linear).
// Every pass after the first needs a whole-image reduction of the one before, so the source must be read four times -- but
// the temp image need not exist: each pass RECOMPUTES the normalised value (two roundings) and, from pass C on, the map
// from the source dtype.  Same arithmetic per value as the temp form (the temp held exactly these f32 values), 4 reads of
// the source + 1 write of the result instead of 1 + 4 x 12 B/px of f32 traffic (f32 -> u8: 51 instead of 75 B/px, u8 ->
// u8: 15 instead of 57 B/px).
template <typename T> __device__ __forceinline__ void store24_raw(T* p, const T (&v)[24]) {
  if constexpr (sizeof(T) == 1) {
#pragma unroll
    for (int i = 0; i < 3; ++i) reinterpret_cast<uint2*>(p)[i] = reinterpret_cast<const uint2*>(v)[i];
  } else {
#pragma unroll
    for (int i = 0; i < 24 * (int)sizeof(T) / 16; ++i) reinterpret_cast<uint4*>(p)[i] = reinterpret_cast<const uint4*>(v)[i];
  }
}

// pass A: metering_func(clamp((src - min) / range, 0, 1), Bounds(0,1)) (tonemap.py:77-103, :147-149): log-gray min/max, sums.
template <typename InT>
__global__ void __launch_bounds__(256) sa_normalise_meter_kernel(const InT* __restrict__ src, long long n_px, const float* __restrict__ b,
                                                                 Workspace* ws) {
  __shared__ float smem[8 * 7];
  const float bmin = b[0];
  const float inv_range = __fdiv_rn(1.0f, __fsub_rn(b[1], bmin));
  float v[7] = {INFINITY, -INFINITY, 0.f, 0.f, 0.f, 0.f, 0.f};
  auto accum = [&](const float* s) {
    const float gray = rgb_gray(s[0], s[1], s[2]);     // (temp - 0) / (1 - 0) == temp
    const float lg = logf(fmaxf(gray, 1e-4f));
    v[0] = fminf(v[0], lg); v[1] = fmaxf(v[1], lg);
    v[2] += lg; v[3] += gray; v[4] += s[0]; v[5] += s[1]; v[6] += s[2];
  };
  // eight pixels per thread with 16-byte accesses when the image allows it, the tail (and unaligned images) per pixel
  const long long n8 = aligned16(src) ? n_px / 8 : 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    float x[24];
    load24(src + 24 * i, x);
#pragma unroll
    for (int e = 0; e < 24; ++e) x[e] = clamp01(__fmul_rn(__fsub_rn(x[e], bmin), inv_range));
#pragma unroll
    for (int q = 0; q < 8; ++q) accum(x + 3 * q);
  }
  for (long long i = 8 * n8 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_px; i += (long long)gridDim.x * blockDim.x) {
    float s[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) s[k] = clamp01(__fmul_rn(__fsub_rn(to_f32(src[3 * i + k]), bmin), inv_range));
    accum(s);
  }
  const int op[7] = {0, 1, 2, 2, 2, 2, 2};
  block_fold<7>(v, op, smem);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < 7; ++k) ws->partials[blockIdx.x * kPartialStride + k] = v[k];
  }
  if (last_block_ticket(&ws->counter[0])) {
    float f[7] = {INFINITY, -INFINITY, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int bk = threadIdx.x; bk < (int)gridDim.x; bk += blockDim.x) {
      f[0] = fminf(f[0], __ldcg(&ws->partials[bk * kPartialStride + 0]));
      f[1] = fmaxf(f[1], __ldcg(&ws->partials[bk * kPartialStride + 1]));
#pragma unroll
      for (int k = 2; k < 7; ++k) f[k] += __ldcg(&ws->partials[bk * kPartialStride + k]);
    }
    block_fold<7>(f, op, smem);
    if (threadIdx.x == 0) {
      const float fn = (float)n_px;
      // 9-float record in the ISP layout so reinhard_params() can be shared:
      // [bounds.min=0, bounds.max=1, log_b.min, log_b.max, log_mean, mean, rgb_mean]
      // with log_bounds = Bounds(log_min, -log_max) exactly as written at tonemap.py:102.
      float* m = ws->scratch + 8;
      m[0] = 0.f; m[1] = 1.f; m[2] = f[0]; m[3] = -f[1];
      m[4] = __fdiv_rn(f[2], fn); m[5] = __fdiv_rn(f[3], fn);
      m[6] = __fdiv_rn(f[4], fn); m[7] = __fdiv_rn(f[5], fn); m[8] = __fdiv_rn(f[6], fn);
    }
  }
}

// pass B (WRITE = false): bounds_func of the Reinhard map (tonemap.py:107-131, :146) -- min / max only, nothing stored.
// pass C (WRITE = true):  dst = linear_func(map, those bounds, gamma) (:154); `temp`, when the caller wants the reference's
//                         side effect, receives the un-normalised map.
template <typename InT, typename OutT, bool WRITE>
__global__ void __launch_bounds__(256) sa_reinhard_kernel(const InT* __restrict__ src, OutT* __restrict__ dst, float* __restrict__ temp,
                                                          long long n_px, const float* __restrict__ b, float intensity, float la, float ca,
                                                          float gamma, Workspace* ws) {
  __shared__ float smem[8 * 2];
  const ReinhardParams p = reinhard_params(ws->scratch + 8, intensity, la, ca);
  const float bmin = b[0];
  const float inv_range = __fdiv_rn(1.0f, __fsub_rn(b[1], bmin));
  const float omin = WRITE ? b[2] : 0.f;
  const float oinv = WRITE ? __fdiv_rn(1.0f, __fsub_rn(b[3], omin)) : 0.f;
  const float inv_gamma = __fdiv_rn(1.0f, gamma);
  const bool has_gamma = gamma != 1.0f;
  float v[2] = {INFINITY, -INFINITY};
  const bool ca0 = ca == 0.f;
  // bmin = 0, range = 1: the scaled value is the normalised value itself.  MUFU evaluation of the map like the ISP kernels
  // below (relative error ~1e-6, inside the <= 1 LSB contract)
  auto map_px = [&](const float* x, float* o) {
    float sc[3], r[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) sc[k] = clamp01(__fmul_rn(__fsub_rn(x[k], bmin), inv_range));
    if (ca0) reinhard_map_fast<true>(p, sc, r); else reinhard_map_fast<false>(p, sc, r);
#pragma unroll
    for (int k = 0; k < 3; ++k) { o[k] = r[k]; v[0] = fminf(v[0], r[k]); v[1] = fmaxf(v[1], r[k]); }
  };
  // linear_func of the map: the map already carries the ~1e-6 relative error of its MUFU evaluation, so the gamma power uses
  // the same lg2 / ex2 pair (powf made this pass issue-bound: 122 us of the call's 205 us on 4096x3000 f32 -> u8)
  auto out_value = [&](float r) -> OutT {
    float y = __fmul_rn(__fsub_rn(r, omin), oinv);
    if (has_gamma) y = fast_pow(y, inv_gamma);
    return cast_from_f32<OutT>(__fmul_rn(clamp01(y), DT<OutT>::scale));      // NaN -> 0 through the saturating clamp
  };
  bool vec = aligned16(src);
  if (WRITE) vec = vec && aligned16(dst) && (!temp || aligned16(temp));
  const long long n8 = vec ? n_px / 8 : 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    float x[24], o[24];
    load24(src + 24 * i, x);
#pragma unroll
    for (int q = 0; q < 8; ++q) map_px(x + 3 * q, o + 3 * q);
    if constexpr (WRITE) {
      if (temp) store24(temp + 24 * i, o);
      alignas(16) OutT y[24];
#pragma unroll
      for (int e = 0; e < 24; ++e) y[e] = out_value(o[e]);
      store24_raw(dst + 24 * i, y);
    }
  }
  for (long long i = 8 * n8 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_px; i += (long long)gridDim.x * blockDim.x) {
    const float x[3] = {to_f32(src[3 * i]), to_f32(src[3 * i + 1]), to_f32(src[3 * i + 2])};
    float o[3];
    map_px(x, o);
    if constexpr (WRITE) {
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        if (temp) temp[3 * i + k] = o[k];
        dst[3 * i + k] = out_value(o[k]);
      }
    }
  }
  if constexpr (!WRITE) {
    const int op[2] = {0, 1};
    block_fold<2>(v, op, smem);
    if (threadIdx.x == 0) {
      ws->partials[blockIdx.x * kPartialStride + 0] = v[0];
      ws->partials[blockIdx.x * kPartialStride + 1] = v[1];
    }
    if (last_block_ticket(&ws->counter[0])) {
      float f[2] = {INFINITY, -INFINITY};
      for (int bk = threadIdx.x; bk < (int)gridDim.x; bk += blockDim.x) {
        f[0] = fminf(f[0], __ldcg(&ws->partials[bk * kPartialStride + 0]));
        f[1] = fmaxf(f[1], __ldcg(&ws->partials[bk * kPartialStride + 1]));
      }
      block_fold<2>(f, op, smem);
      if (threadIdx.x == 0) { ws->scratch[2] = f[0]; ws->scratch[3] = f[1]; }
    }
  }
}

// ---------------------------------------------------------------- ISP Reinhard (camera_isp.py:177-218)
// pass 1: p -> image (ISP dtype), max over f32 p.  pass 2: (image / max_out)^(1/gamma) * scale -> out.
// camera_isp.py:200-213 (pass 1): the un-normalised map written back in the ISP dtype + its frame-global max.
// Eight pixels per thread with 16-byte accesses; the map uses the MUFU evaluation of the fused sweep (relative error
// ~1e-6, inside the <= 1 LSB contract; the literal IEEE order costs ~10x more instructions for the same output).
// blockIdx.y = image of the batch (one launch per pass for all frames of a time step; frame_max[blockIdx.y])
struct ImagePtrs { void* image[B200ISP_MAX_FRAMES]; void* out[B200ISP_MAX_FRAMES]; };

template <typename T>
__global__ void __launch_bounds__(256) isp_reinhard_pass1_kernel(const ImagePtrs ptrs, long long n_px,
                                                                 const float* __restrict__ metrics, float intensity,
                                                                 float la, float ca, Workspace* ws) {
  __shared__ float smem[8];
  T* __restrict__ image = reinterpret_cast<T*>(ptrs.image[blockIdx.y]);
  const ReinhardParams p = reinhard_params(metrics, intensity, la, ca);
  const float b = -p.bmin * p.inv_range;
  const bool ca0 = ca == 0.f;
  const bool vec = (reinterpret_cast<uintptr_t>(image) & 15u) == 0;
  const long long n8 = vec ? n_px / 8 : 0;
  float mx = 0.f;
  auto map_px = [&](const float* x, float* o) {
    const float sc[3] = {fmaf(x[0], p.inv_range, b), fmaf(x[1], p.inv_range, b), fmaf(x[2], p.inv_range, b)};
    float r[3];
    if (ca0) reinhard_map_fast<true>(p, sc, r); else reinhard_map_fast<false>(p, sc, r);
    o[0] = r[0]; o[1] = r[1]; o[2] = r[2];
    mx = fmaxf(mx, fmaxf(r[0], fmaxf(r[1], r[2])));
  };
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    float v[24], o[24];
    load24(image + 24 * i, v);
#pragma unroll
    for (int q = 0; q < 8; ++q) map_px(v + 3 * q, o + 3 * q);
    store24(image + 24 * i, o);
  }
  for (long long i = 8 * n8 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_px; i += (long long)gridDim.x * blockDim.x) {
    const float x[3] = {to_f32(image[3 * i]), to_f32(image[3 * i + 1]), to_f32(image[3 * i + 2])};
    float o[3];
    map_px(x, o);
#pragma unroll
    for (int k = 0; k < 3; ++k) image[3 * i + k] = cast_from_f32<T>(o[k]);
  }
  mx = warp_max(mx);
  if ((threadIdx.x & 31) == 0) smem[threadIdx.x >> 5] = mx;
  __syncthreads();
  if (threadIdx.x < 32) {
    mx = threadIdx.x < (blockDim.x >> 5) ? smem[threadIdx.x] : 0.f;
    mx = warp_max(mx);
    if (threadIdx.x == 0 && mx > 0.f) atomicMax(reinterpret_cast<unsigned int*>(&ws->frame_max[blockIdx.y]), __float_as_uint(mx));
  }
}

// camera_isp.py:215-218 (pass 2): out = cast(scale * (x' / max_out)^(1/gamma)), x' = the stored map; no clamp in the
// reference (values <= 1 + one f16 ulp) -- the integer casts saturate.
template <typename T, typename OutT>
__global__ void __launch_bounds__(256) isp_reinhard_pass2_kernel(const ImagePtrs ptrs, long long n_elems, float gamma, Workspace* ws) {
  const T* __restrict__ image = reinterpret_cast<const T*>(ptrs.image[blockIdx.y]);
  OutT* __restrict__ out = reinterpret_cast<OutT*>(ptrs.out[blockIdx.y]);
  const float inv_max = __fdiv_rn(1.0f, fmaxf(1e-6f, __ldcg(&ws->frame_max[blockIdx.y])));
  const float inv_gamma = (float)(1.0 / (double)gamma);     // python double 1.0 / gamma -> f32 constant (:217)
  const bool has_gamma = gamma != 1.0f;
  constexpr int PER = 16 / (int)sizeof(T);
  const bool vec = (reinterpret_cast<uintptr_t>(image) & 15u) == 0 &&
                   (reinterpret_cast<uintptr_t>(out) & (PER * sizeof(OutT) - 1)) == 0;
  const long long nv = vec ? n_elems / PER : 0;
  auto tone = [&](float x) {
    float q = x * inv_max;
    if (has_gamma) q = fast_pow(fmaxf(q, 0.f), inv_gamma);
    return cast_from_f32<OutT>(__fmul_rn(DT<OutT>::scale, q));
  };
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += (long long)gridDim.x * blockDim.x) {
    const uint4 w = reinterpret_cast<const uint4*>(image)[i];
    const T* t = reinterpret_cast<const T*>(&w);
    alignas(16) OutT o[PER];
#pragma unroll
    for (int j = 0; j < PER; ++j) o[j] = tone(to_f32(t[j]));
    if constexpr (PER * sizeof(OutT) == 16) reinterpret_cast<uint4*>(out)[i] = *reinterpret_cast<const uint4*>(o);
    else if constexpr (PER * sizeof(OutT) == 8) reinterpret_cast<uint2*>(out)[i] = *reinterpret_cast<const uint2*>(o);
    else if constexpr (PER * sizeof(OutT) == 4) reinterpret_cast<uint32_t*>(out)[i] = *reinterpret_cast<const uint32_t*>(o);
    else {
#pragma unroll
      for (int j = 0; j < PER; ++j) out[PER * i + j] = o[j];
    }
  }
  for (long long i = nv * PER + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_elems; i += (long long)gridDim.x * blockDim.x)
    out[i] = tone(to_f32(image[i]));
}

// ---------------------------------------------------------------- loaders (camera_isp.py:82-99)
template <typename InT, typename OutT, int MODE>
__global__ void __launch_bounds__(256) load_convert_kernel(const InT* __restrict__ src, OutT* __restrict__ dst, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float x = to_f32(src[i]);
  if (MODE == 0) x = __fdiv_rn(x, 65535.0f);
  dst[i] = cast_from_f32<OutT>(x);
}

}  // namespace isp

using namespace isp;

extern "C" int b200isp_bounds(const void* src, int dtype, int64_t n_elems, float* bounds_out, void* workspace,
                              b200isp_stream stream) {
  ISP_REQUIRE(src && bounds_out && workspace && n_elems > 0, B200ISP_E_ARG, "bounds: bad argument");
  ISP_DISPATCH_DTYPE(dtype, T, return (launch_bounds_kernel<T>((const T*)src, n_elems, bounds_out, (Workspace*)workspace, (cudaStream_t)stream)));
  return B200ISP_OK;
}

extern "C" int b200isp_linear(const void* src, int in_dtype, void* dst, int out_dtype, int64_t n_elems,
                              const float* bounds, float gamma, b200isp_stream stream) {
  ISP_REQUIRE(n_elems >= 0, B200ISP_E_SHAPE, "linear: negative size");
  if (n_elems == 0) return B200ISP_OK;
  ISP_REQUIRE(src && dst && bounds, B200ISP_E_ARG, "linear: null pointer");
  ISP_REQUIRE(gamma > 0.f, B200ISP_E_ARG, "linear: gamma must be positive");
  ISP_DISPATCH_DTYPE(in_dtype, InT, {
    ISP_DISPATCH_DTYPE(out_dtype, OutT, return (launch_linear<InT, OutT>((const InT*)src, (OutT*)dst, n_elems, bounds, gamma, (cudaStream_t)stream)));
  });
  return B200ISP_OK;
}

extern "C" int b200isp_reinhard_standalone(const void* src, int in_dtype, float* temp, void* dst, int out_dtype,
                                           int64_t n_pixels, float gamma, float intensity, float light_adapt,
                                           float color_adapt, void* workspace, b200isp_stream stream) {
  ISP_REQUIRE(n_pixels > 0 && src && dst && workspace, B200ISP_E_ARG, "reinhard_standalone: bad argument");
  ISP_REQUIRE(gamma > 0.f, B200ISP_E_ARG, "reinhard_standalone: gamma must be positive");
  Workspace* ws = (Workspace*)workspace;
  cudaStream_t s = (cudaStream_t)stream;
  float* ws_bounds = reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + offsetof(Workspace, scratch));   // [0:2] source, [2:4] map
  const int grid = meter_grid(n_pixels);
  // pass C writes: 8 pixels per thread, enough CTAs for every pixel (no grid-stride tail of idle SMs)
  const unsigned wgrid = (unsigned)std::max<long long>(1, std::min<long long>((n_pixels / 8 + 255) / 256, 1 << 30));
  ISP_DISPATCH_DTYPE(in_dtype, InT, {
    int st = launch_bounds_kernel<InT>((const InT*)src, n_pixels * 3, ws_bounds, ws, s);       // tonemap.py:146
    if (st) return st;
    sa_normalise_meter_kernel<InT><<<grid, 256, 0, s>>>((const InT*)src, n_pixels, ws_bounds, ws);   // :147-149
    ISP_LAUNCH_CHECK("sa_normalise_meter_kernel");
    sa_reinhard_kernel<InT, uint8_t, false><<<grid, 256, 0, s>>>((const InT*)src, nullptr, nullptr, n_pixels, ws_bounds, intensity,
                                                                  light_adapt, color_adapt, gamma, ws);   // :150-153
    ISP_LAUNCH_CHECK("sa_reinhard_kernel<bounds>");
    ISP_DISPATCH_DTYPE(out_dtype, OutT, {
      sa_reinhard_kernel<InT, OutT, true><<<wgrid, 256, 0, s>>>((const InT*)src, (OutT*)dst, temp, n_pixels, ws_bounds, intensity,
                                                                 light_adapt, color_adapt, gamma, ws);     // :154
    });
    ISP_LAUNCH_CHECK("sa_reinhard_kernel<write>");
  });
  return B200ISP_OK;
}

extern "C" int b200isp_metering_update(const void* const* images_host, int n_images, int dtype, int height, int width,
                                       int stride, float alpha, float* metrics, void* workspace, b200isp_stream stream) {
  ISP_REQUIRE(images_host && metrics && workspace, B200ISP_E_ARG, "metering_update: null pointer");
  ISP_REQUIRE(n_images >= 1 && n_images <= B200ISP_MAX_FRAMES, B200ISP_E_FRAMES,
              "metering_update: %d images (1..%d supported per call)", n_images, B200ISP_MAX_FRAMES);
  ISP_REQUIRE(height > 0 && width > 0 && stride > 0, B200ISP_E_SHAPE, "metering_update: bad shape");
  ISP_REQUIRE(dtype == B200ISP_F16 || dtype == B200ISP_F32, B200ISP_E_DTYPE, "metering_update: ISP dtype must be f16 or f32");
  const int hs = (height + stride - 1) / stride, wsamp = (width + stride - 1) / stride;
  const long long n = (long long)n_images * hs * wsamp;
  ISP_DISPATCH_DTYPE(dtype, T, {
    ImageSampler<T> smp;
    for (int i = 0; i < n_images; ++i) smp.img[i] = (const T*)images_host[i];
    smp.W = width; smp.stride = stride; smp.hs = hs; smp.ws_ = wsamp;
    return launch_metering(smp, n, alpha, metrics, metrics, (Workspace*)workspace, (cudaStream_t)stream);
  });
  return B200ISP_OK;
}

// shared-exposure halves of b200isp_metering_update (see include/b200isp.h)
template <class F>
static int with_image_sampler(const char* what, const void* const* images_host, int n_images, int dtype, int height, int width,
                              int stride, F f) {
  ISP_REQUIRE(images_host, B200ISP_E_ARG, "%s: null pointer", what);
  ISP_REQUIRE(n_images >= 1 && n_images <= B200ISP_MAX_FRAMES, B200ISP_E_FRAMES,
              "%s: %d images (1..%d supported per call)", what, n_images, B200ISP_MAX_FRAMES);
  ISP_REQUIRE(height > 0 && width > 0 && stride > 0, B200ISP_E_SHAPE, "%s: bad shape", what);
  ISP_REQUIRE(dtype == B200ISP_F16 || dtype == B200ISP_F32, B200ISP_E_DTYPE, "%s: ISP dtype must be f16 or f32", what);
  const int hs = (height + stride - 1) / stride, wsamp = (width + stride - 1) / stride;
  const long long n = (long long)n_images * hs * wsamp;
  ISP_DISPATCH_DTYPE(dtype, T, {
    ImageSampler<T> smp;
    for (int i = 0; i < n_images; ++i) smp.img[i] = (const T*)images_host[i];
    smp.W = width; smp.stride = stride; smp.hs = hs; smp.ws_ = wsamp;
    return f(smp, n);
  });
  return B200ISP_OK;
}

extern "C" int b200isp_metering_phase1(const void* const* images_host, int n_images, int dtype, int height, int width,
                                       int stride, float* rec1, void* workspace, b200isp_stream stream) {
  ISP_REQUIRE(rec1 && workspace, B200ISP_E_ARG, "metering_phase1: null pointer");
  return with_image_sampler("metering_phase1", images_host, n_images, dtype, height, width, stride, [&](const auto& smp, long long n) {
    return launch_metering_phase1(smp, n, (Workspace*)workspace, (cudaStream_t)stream, nullptr, rec1);
  });
}

extern "C" int b200isp_metering_phase2(const void* const* images_host, int n_images, int dtype, int height, int width,
                                       int stride, const float* gathered1, int world, float alpha, const float* metrics_prev,
                                       float* rec2, void* workspace, b200isp_stream stream) {
  ISP_REQUIRE(gathered1 && metrics_prev && rec2 && workspace && world >= 1, B200ISP_E_ARG, "metering_phase2: bad argument");
  return with_image_sampler("metering_phase2", images_host, n_images, dtype, height, width, stride, [&](const auto& smp, long long n) {
    return launch_metering_phase2(smp, n, gathered1, world, alpha, metrics_prev, (Workspace*)workspace, (cudaStream_t)stream,
                                  nullptr, rec2);
  });
}

// camera_isp.py:394-403 for a list of same-size images: one launch per pass for the whole list
extern "C" int b200isp_isp_reinhard_batch(void* const* images_host, void* const* outputs_host, int n_images, int dtype,
                                          int out_dtype, int64_t n_pixels, const float* metrics, float gamma, float intensity,
                                          float light_adapt, float color_adapt, void* workspace, b200isp_stream stream) {
  ISP_REQUIRE(images_host && outputs_host && metrics && workspace && n_pixels > 0, B200ISP_E_ARG, "isp_reinhard: bad argument");
  ISP_REQUIRE(n_images >= 1 && n_images <= B200ISP_MAX_FRAMES, B200ISP_E_FRAMES, "isp_reinhard: %d images (1..%d per call)",
              n_images, B200ISP_MAX_FRAMES);
  ISP_REQUIRE(dtype == B200ISP_F16 || dtype == B200ISP_F32, B200ISP_E_DTYPE, "isp_reinhard: ISP dtype must be f16 or f32");
  ISP_REQUIRE(gamma > 0.f, B200ISP_E_ARG, "isp_reinhard: gamma must be positive");
  ImagePtrs ptrs;
  for (int i = 0; i < n_images; ++i) {
    ISP_REQUIRE(images_host[i] && outputs_host[i], B200ISP_E_ARG, "isp_reinhard: null image %d", i);
    ptrs.image[i] = images_host[i];
    ptrs.out[i] = outputs_host[i];
  }
  Workspace* ws = (Workspace*)workspace;
  cudaStream_t s = (cudaStream_t)stream;
  const dim3 grid((unsigned)meter_grid(n_pixels), (unsigned)n_images);
  {
    const int st = cuda_status(cudaMemsetAsync(&ws->frame_max[0], 0, sizeof(float) * n_images, s), "memset frame_max");
    if (st) return st;
  }
  ISP_DISPATCH_DTYPE(dtype, T, {
    isp_reinhard_pass1_kernel<T><<<grid, 256, 0, s>>>(ptrs, n_pixels, metrics, intensity, light_adapt, color_adapt, ws);
    ISP_LAUNCH_CHECK("isp_reinhard_pass1_kernel");
    ISP_DISPATCH_DTYPE(out_dtype, OutT, (isp_reinhard_pass2_kernel<T, OutT><<<grid, 256, 0, s>>>(ptrs, n_pixels * 3, gamma, ws)));
  });
  ISP_LAUNCH_CHECK("isp_reinhard_pass2_kernel");
  return B200ISP_OK;
}

extern "C" int b200isp_isp_reinhard(void* image, int dtype, void* output, int out_dtype, int64_t n_pixels,
                                    const float* metrics, float gamma, float intensity, float light_adapt,
                                    float color_adapt, void* workspace, b200isp_stream stream) {
  ISP_REQUIRE(image && output, B200ISP_E_ARG, "isp_reinhard: bad argument");
  return b200isp_isp_reinhard_batch(&image, &output, 1, dtype, out_dtype, n_pixels, metrics, gamma, intensity, light_adapt,
                                    color_adapt, workspace, stream);
}

extern "C" int b200isp_load_convert(const void* src, void* dst, int out_dtype, int64_t n_elems, int mode,
                                    b200isp_stream stream) {
  ISP_REQUIRE(n_elems >= 0 && mode >= 0 && mode <= 2, B200ISP_E_ARG, "load_convert: bad argument");
  if (n_elems == 0) return B200ISP_OK;
  ISP_REQUIRE(src && dst, B200ISP_E_ARG, "load_convert: null pointer");
  ISP_REQUIRE(out_dtype == B200ISP_F16 || out_dtype == B200ISP_F32, B200ISP_E_DTYPE, "load_convert: ISP dtype must be f16 or f32");
  cudaStream_t s = (cudaStream_t)stream;
  const unsigned blocks = (unsigned)((n_elems + 255) / 256);
  ISP_DISPATCH_DTYPE(out_dtype, OutT, {
    if (mode == 0) load_convert_kernel<uint16_t, OutT, 0><<<blocks, 256, 0, s>>>((const uint16_t*)src, (OutT*)dst, n_elems);
    else if (mode == 1) load_convert_kernel<float, OutT, 1><<<blocks, 256, 0, s>>>((const float*)src, (OutT*)dst, n_elems);
    else load_convert_kernel<uint16_t, OutT, 2><<<blocks, 256, 0, s>>>((const uint16_t*)src, (OutT*)dst, n_elems);
  });
  ISP_LAUNCH_CHECK("load_convert_kernel");
  return B200ISP_OK;
}

// ---------------------------------------------------------------- luminance histogram + percentiles (EXTENSION)
// north_star asks for a "percentile histogram" next to the reference's statistics; the reference has none
// (SURVEY 2.4).  Defined on the SAME samples the metering uses (the [::stride, ::stride] RGB of all frames, as cached
// by the metering pass): bin = min(bins-1, trunc(rgb_gray(sample) * bins)), luminance in [0,1].  Integer counts:
// deterministic.  Percentile p = smallest bin edge (b+1)/bins whose cumulative count reaches p % of the samples.
namespace isp {
__global__ void __launch_bounds__(256) sample_histogram_kernel(const float* __restrict__ samples, long long n, int bins,
                                                               unsigned int* __restrict__ hist) {
  extern __shared__ unsigned int sh[];
  for (int b = threadIdx.x; b < bins; b += blockDim.x) sh[b] = 0u;
  __syncthreads();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float g = rgb_gray(__ldcg(samples + 3 * i), __ldcg(samples + 3 * i + 1), __ldcg(samples + 3 * i + 2));
    const int b = min(bins - 1, max(0, (int)__float2int_rz(__fmul_rn(g, (float)bins))));
    atomicAdd(&sh[b], 1u);
  }
  __syncthreads();
  for (int b = threadIdx.x; b < bins; b += blockDim.x)
    if (sh[b]) atomicAdd(&hist[b], sh[b]);
}

__global__ void histogram_percentiles_kernel(const unsigned int* __restrict__ hist, int bins, const float* __restrict__ pct, int npct,
                                             float* __restrict__ out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  unsigned long long total = 0;
  for (int b = 0; b < bins; ++b) total += hist[b];
  for (int k = 0; k < npct; ++k) {
    const double need = (double)pct[k] * 0.01 * (double)total;
    unsigned long long cum = 0;
    int b = 0;
    for (; b < bins; ++b) { cum += hist[b]; if ((double)cum >= need) break; }
    out[k] = (float)(min(b, bins - 1) + 1) / (float)bins;
  }
}
}  // namespace isp

extern "C" int b200isp_sample_histogram(const float* samples, int64_t n_samples, int bins, uint32_t* hist, b200isp_stream stream) {
  ISP_REQUIRE(samples && hist && n_samples >= 0 && bins >= 1 && bins <= 4096, B200ISP_E_ARG, "sample_histogram: bad argument");
  cudaStream_t s = (cudaStream_t)stream;
  int st = cuda_status(cudaMemsetAsync(hist, 0, sizeof(uint32_t) * bins, s), "memset hist");
  if (st) return st;
  if (n_samples == 0) return B200ISP_OK;
  sample_histogram_kernel<<<meter_grid(n_samples), 256, sizeof(unsigned int) * bins, s>>>(samples, n_samples, bins, hist);
  ISP_LAUNCH_CHECK("sample_histogram_kernel");
  return B200ISP_OK;
}

extern "C" int b200isp_histogram_percentiles(const uint32_t* hist, int bins, const float* percents, int n_percents, float* out,
                                             b200isp_stream stream) {
  ISP_REQUIRE(hist && percents && out && bins >= 1 && n_percents >= 1, B200ISP_E_ARG, "histogram_percentiles: bad argument");
  histogram_percentiles_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(hist, bins, percents, n_percents, out);
  ISP_LAUNCH_CHECK("histogram_percentiles_kernel");
  return B200ISP_OK;
}
