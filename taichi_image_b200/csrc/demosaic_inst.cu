// Explicit instantiation of the stand-alone bayer_to_rgb sweep (plane_sweep.cuh), one translation unit per dtype.
// Build with -DISP_PLANE_T=<type>
#include "plane_sweep.cuh"

namespace isp {
template <> int run_demosaic_sweep<ISP_PLANE_T>(const void* bayer, void* rgb, int H, int W, int pattern, const float* ccm, bool bilinear,
                                                cudaStream_t s) {
  return run_demosaic_sweep_impl<ISP_PLANE_T>(bayer, rgb, H, W, pattern, ccm, bilinear, s);
}
}  // namespace isp
