// Version / error plumbing of the C ABI (include/nerf_b200.h).
#include <string.h>

#include "common.cuh"

namespace nb200 {

thread_local char g_last_cuda_error[256] = "";

int record_cuda_error(cudaError_t e, const char* what) {
  snprintf(g_last_cuda_error, sizeof(g_last_cuda_error), "%s: %s (%s)", what, cudaGetErrorName(e),
           cudaGetErrorString(e));
  return NB200_ERR_CUDA;
}

int sm_count() {
  static int cached[64];   // per device; racing writers store the same value
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev >= 0 && dev < 64 && cached[dev] > 0) return cached[dev];
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
    return 148;
  if (dev >= 0 && dev < 64) cached[dev] = n;
  return n;
}

}  // namespace nb200

extern "C" {

int nb200_version(void) { return 100; }

int nb200_compiled_arch(void) { return 100; }

int nb200_device_arch(void) {
  int dev = 0, major = 0, minor = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return NB200_ERR_CUDA;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess)
    return NB200_ERR_CUDA;
  if (cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev) != cudaSuccess)
    return NB200_ERR_CUDA;
  return major * 10 + minor;
}

const char* nb200_error_string(int code) {
  switch (code) {
    case NB200_OK: return "ok";
    case NB200_ERR_ARG: return "invalid argument";
    case NB200_ERR_UNSUPPORTED: return "unsupported shape or precision";
    case NB200_ERR_CUDA: return "CUDA runtime error";
    case NB200_ERR_ARCH: return "device is not sm_100 (B200)";
    case NB200_ERR_WORKSPACE: return "workspace too small";
    case NB200_ERR_KERNEL: return "kernel-internal failure";
    default: return "unknown error";
  }
}

const char* nb200_last_cuda_error(void) { return nb200::g_last_cuda_error; }

}  // extern "C"
