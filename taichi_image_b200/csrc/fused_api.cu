// C entry point of the fused packed12 sweep + the sparse metering launch (see fused_isp.cuh).
#include "fused_isp.cuh"

namespace isp {
extern template int run_fused<true, uint8_t>(const FramePtrs&, int, const b200isp_fused_params&, IspConsts, cudaStream_t);
extern template int run_fused<true, uint16_t>(const FramePtrs&, int, const b200isp_fused_params&, IspConsts, cudaStream_t);
extern template int run_fused<true, __half>(const FramePtrs&, int, const b200isp_fused_params&, IspConsts, cudaStream_t);
extern template int run_fused<false, uint8_t>(const FramePtrs&, int, const b200isp_fused_params&, IspConsts, cudaStream_t);
extern template int run_fused<false, uint16_t>(const FramePtrs&, int, const b200isp_fused_params&, IspConsts, cudaStream_t);
extern template int run_fused<false, __half>(const FramePtrs&, int, const b200isp_fused_params&, IspConsts, cudaStream_t);
extern template int run_fused<false, float>(const FramePtrs&, int, const b200isp_fused_params&, IspConsts, cudaStream_t);
}  // namespace isp

using namespace isp;

extern "C" int b200isp_process_packed12(const uint8_t* const* packed_host, void* const* out_host, int n_frames,
                                        const b200isp_fused_params* params, float* metrics, void* workspace,
                                        b200isp_stream stream) {
  ISP_REQUIRE(packed_host && params && workspace, B200ISP_E_ARG, "process_packed12: null pointer");
  const b200isp_fused_params& p = *params;
  ISP_REQUIRE(n_frames >= 1 && n_frames <= B200ISP_MAX_FRAMES, B200ISP_E_FRAMES,
              "process_packed12: %d frames (1..%d supported per call)", n_frames, B200ISP_MAX_FRAMES);
  ISP_REQUIRE(p.height >= 4 && p.width >= 8 && p.height % 2 == 0 && p.width % 8 == 0, B200ISP_E_SHAPE,
              "process_packed12: fused path needs even height >= 4 and width %% 8 == 0, got %dx%d", p.height, p.width);
  ISP_REQUIRE(p.pattern >= 0 && p.pattern <= 3, B200ISP_E_ARG, "process_packed12: unknown pattern %d", p.pattern);
  ISP_REQUIRE(p.isp_dtype == B200ISP_F16 || p.isp_dtype == B200ISP_F32, B200ISP_E_DTYPE, "process_packed12: isp_dtype must be f16/f32");
  ISP_REQUIRE(p.tonemap >= B200ISP_TM_LINEAR && p.tonemap <= B200ISP_TM_NONE, B200ISP_E_ARG, "process_packed12: unknown tonemap %d", p.tonemap);
  const bool needs_out = true;
  ISP_REQUIRE(!needs_out || out_host, B200ISP_E_ARG, "process_packed12: null output list");
  ISP_REQUIRE(p.tonemap == B200ISP_TM_NONE || metrics, B200ISP_E_ARG, "process_packed12: metrics required for tone mapping");
  ISP_REQUIRE(p.tonemap == B200ISP_TM_NONE || p.gamma > 0.f, B200ISP_E_ARG, "process_packed12: gamma must be positive");
  if (p.tonemap == B200ISP_TM_NONE)
    ISP_REQUIRE(p.out_dtype == p.isp_dtype, B200ISP_E_DTYPE, "process_packed12: TM_NONE writes the ISP dtype");
  else
    ISP_REQUIRE(p.out_dtype == B200ISP_U8 || p.out_dtype == B200ISP_U16 || p.out_dtype == B200ISP_F16, B200ISP_E_DTYPE,
                "process_packed12: tone-mapped output must be u8, u16 or f16");

  FramePtrs fp;
  for (int i = 0; i < n_frames; ++i) {
    fp.in[i] = packed_host[i];
    fp.out[i] = out_host[i];
    ISP_REQUIRE(fp.in[i] && fp.out[i], B200ISP_E_ARG, "process_packed12: null frame pointer %d", i);
    ISP_REQUIRE(((reinterpret_cast<uintptr_t>(fp.in[i]) & 3u) == 0) && ((reinterpret_cast<uintptr_t>(fp.out[i]) & 15u) == 0),
                B200ISP_E_ALIGN, "process_packed12: frame %d needs 4-byte aligned input and 16-byte aligned output", i);
  }
  IspConsts k;
  k.H = p.height; k.W = p.width; k.pattern = p.pattern;
  k.ccm = p.has_ccm ? 1 : 0;
  for (int i = 0; i < 9; ++i) k.m[i] = p.ccm[i];
  k.gamma = p.gamma; k.intensity = p.intensity; k.la = p.light_adapt; k.ca = p.color_adapt;
  k.metrics = metrics; k.ws = (Workspace*)workspace; k.frame0 = 0;
  cudaStream_t s = (cudaStream_t)stream;
  const bool cam16 = p.isp_dtype == B200ISP_F16;

  if (p.update_metering && p.tonemap != B200ISP_TM_NONE) {
    const int stride = p.metering_stride > 0 ? p.metering_stride : 8;
    const int hs = (p.height + stride - 1) / stride, wsamp = (p.width + stride - 1) / stride;
    const long long n = (long long)n_frames * hs * wsamp;
    float* cache = (p.meter_cache && p.meter_cache_bytes >= (size_t)n * 3 * sizeof(float)) ? (float*)p.meter_cache : nullptr;
    int st;
    if (stride % 8 == 0) {
      if (cam16) {
        Packed12FastSampler<true> smp{Packed12Src<true>{fp, p.width * 3 / 2}, k, stride, hs, wsamp, p.width * 3 / 8};
        st = launch_metering(smp, n, p.alpha, metrics, k.ws, s, cache);
      } else {
        Packed12FastSampler<false> smp{Packed12Src<false>{fp, p.width * 3 / 2}, k, stride, hs, wsamp, p.width * 3 / 8};
        st = launch_metering(smp, n, p.alpha, metrics, k.ws, s, cache);
      }
    } else if (cam16) {
      Packed12Sampler<true> smp{Packed12Src<true>{fp, p.width * 3 / 2}, k, stride, hs, wsamp};
      st = launch_metering(smp, n, p.alpha, metrics, k.ws, s, cache);
    } else {
      Packed12Sampler<false> smp{Packed12Src<false>{fp, p.width * 3 / 2}, k, stride, hs, wsamp};
      st = launch_metering(smp, n, p.alpha, metrics, k.ws, s, cache);
    }
    if (st) return st;
  }

#define RUN(CAM, T) return run_fused<CAM, T>(fp, n_frames, p, k, s)
  if (cam16) {
    switch (p.out_dtype) {
      case B200ISP_U8: RUN(true, uint8_t);
      case B200ISP_U16: RUN(true, uint16_t);
      case B200ISP_F16: RUN(true, __half);
      default: break;
    }
  } else {
    switch (p.out_dtype) {
      case B200ISP_U8: RUN(false, uint8_t);
      case B200ISP_U16: RUN(false, uint16_t);
      case B200ISP_F16: RUN(false, __half);
      case B200ISP_F32: RUN(false, float);
      default: break;
    }
  }
#undef RUN
  isp::set_error("process_packed12: unsupported isp/out dtype combination %d/%d", p.isp_dtype, p.out_dtype);
  return B200ISP_E_DTYPE;
}
