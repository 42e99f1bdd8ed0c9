// Error plumbing and device checks shared by every C-ABI entry point (include/ld_b200.h).
#include <stdarg.h>

#include "common.cuh"

namespace ldb {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return 0;
  set_error("%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
  return (int)e;
}

}  // namespace ldb

extern "C" const char* ldb200_last_error(void) { return ldb::g_err; }
extern "C" int ldb200_version(void) { return 100; }

extern "C" int ldb200_check_device(void) {
  int dev = 0;
  LDB_CUDA(cudaGetDevice(&dev));
  int major = 0, minor = 0;
  LDB_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  LDB_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
  LDB_REQUIRE(major == 10, LDB_EARCH, "ld_b200 kernels are built for sm_100a only; device is sm_%d%d", major, minor);
  return 0;
}
