// Host side + instantiations of the one-sweep Camera32 Reinhard path (reinhard_u16.cuh).
#include "reinhard_u16.cuh"

namespace isp {

size_t reinhard_u16_frame_bytes(int H, int W) { return reinhard_u16_frame_bytes_impl(H, W); }

static int run_pass_a(const U16Scratch& sc, const FramePtrs& fp, const IspConsts& k, int n_frames, int rpt, cudaStream_t s) {
  const Stream2Geom g = make_geom2(k.H, k.W, n_frames, rpt);
  Packed12Loader2<false> ld;
  ld.fp = fp; ld.pitch_words = k.W * 3 / 8; ld.frame0 = 0; ld.ids = k.ids;
  const EpiReinhardMaxU16<true> e{sc, k};
  int st = B200ISP_OK;
  ISP_DISPATCH_PATTERN(k.pattern, P, { st = launch_stream2<P>(ld, e, g, s, "isp_stream<reinhard_u16>", k.kbase != 0); });
  return st;
}

template <typename OutT>
static int run_impl(const FramePtrs& fp, int n_frames, const b200isp_fused_params& p, IspConsts k, cudaStream_t s) {
  k.frame0 = 0;
  U16Scratch sc;
  const size_t per = reinhard_u16_frame_bytes(k.H, k.W), map_bytes = (((size_t)k.H * k.W * 6) + 15) & ~(size_t)15;
  for (int f = 0; f < n_frames; ++f) {
    sc.map[f] = reinterpret_cast<uint16_t*>((char*)p.reinhard_scratch + per * f);
    sc.frame[f] = reinterpret_cast<float*>((char*)p.reinhard_scratch + per * f + map_bytes);
  }
  if (p.profile_start) record_profile_event(p.profile_start, s);
  int st = run_pass_a(sc, fp, k, n_frames, p.rows_per_task, s);
  if (p.profile_stop) record_profile_event(p.profile_stop, s);
  if (st) return st;
  const dim3 grid((unsigned)((k.W / 8 + 127) / 128), (unsigned)k.H, (unsigned)n_frames);
  if (k.gamma != 1.0f) reinhard_u16_out_kernel<OutT, true, true><<<grid, 128, 0, s>>>(sc, fp, k);
  else reinhard_u16_out_kernel<OutT, true, false><<<grid, 128, 0, s>>>(sc, fp, k);
  return cuda_status(cudaPeekAtLastError(), "reinhard_u16_out_kernel");
}

template <> int run_reinhard_u16<uint8_t>(const FramePtrs& fp, int n, const b200isp_fused_params& p, IspConsts k, cudaStream_t s) { return run_impl<uint8_t>(fp, n, p, k, s); }
template <> int run_reinhard_u16<uint16_t>(const FramePtrs& fp, int n, const b200isp_fused_params& p, IspConsts k, cudaStream_t s) { return run_impl<uint16_t>(fp, n, p, k, s); }
template <> int run_reinhard_u16<__half>(const FramePtrs& fp, int n, const b200isp_fused_params& p, IspConsts k, cudaStream_t s) { return run_impl<__half>(fp, n, p, k, s); }

}  // namespace isp
