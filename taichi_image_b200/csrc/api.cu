// Library-level entry points: version, thread-local error string, workspace size.
#include <stdarg.h>
#include <string.h>
#include "common.cuh"

namespace isp {

static thread_local char g_error[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

int cuda_status(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return B200ISP_OK;
  set_error("%s: CUDA error %d (%s)", what, (int)e, cudaGetErrorString(e));
  return B200ISP_E_CUDA;
}

}  // namespace isp

extern "C" int b200isp_version(void) { return B200ISP_VERSION; }
extern "C" const char* b200isp_last_error(void) { return isp::g_error; }
extern "C" size_t b200isp_workspace_bytes(void) { return isp::kWorkspaceBytes; }
