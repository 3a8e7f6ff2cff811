// Explicit instantiations of the resizing sweep, one translation unit per (ISP dtype, mode) so that they compile in
// parallel.  Build with -DISP_RZ_CAM16=0|1 -DISP_RZ_MODE=0|1|2 (RZ_RGB | RZ_LINEAR | RZ_RSTORE).
#define ISP_RZ_INST 1
#include "fused_resize.cu"

namespace isp {
#define ISP_RZ_DEF(T)                                                                                                    \
  template <> int run_resize_pass<ISP_RZ_CAM16 != 0, ISP_RZ_MODE, T>(const FramePtrs& io, const IspConsts& k,            \
                                                                      const b200isp_fused_params& p, int n, cudaStream_t s) { \
    return run_resize_pass_impl<ISP_RZ_CAM16 != 0, ISP_RZ_MODE, T>(io, k, p, n, s);                                     \
  }
#if ISP_RZ_MODE == 1
ISP_RZ_DEF(uint8_t) ISP_RZ_DEF(uint16_t) ISP_RZ_DEF(__half)
#elif ISP_RZ_CAM16
ISP_RZ_DEF(__half)
#else
ISP_RZ_DEF(float)
#endif
}  // namespace isp
