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

// ---- peer-shareable device buffers (CUDA IPC): the only entry points that allocate.  A buffer is cudaMalloc'ed by
// its owner, exported as a 64-byte handle, and opened by the other ranks of the node WITH THEIR OWN DEVICE CURRENT, which
// is what maps it for that device's kernels over NVLink (cudaIpcMemLazyEnablePeerAccess).
int nb200_p2p_alloc(size_t bytes, void** dev_ptr, void* handle64) {
  if (!dev_ptr || !handle64 || bytes == 0) return NB200_ERR_ARG;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  void* p = nullptr;
  NB_CUDA_CHECK(cudaMalloc(&p, bytes));
  NB_CUDA_CHECK(cudaMemset(p, 0, bytes));
  NB_CUDA_CHECK(cudaDeviceSynchronize());
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) { cudaFree(p); return nb200::record_cuda_error(e, "cudaIpcGetMemHandle"); }
  memcpy(handle64, &h, 64);
  *dev_ptr = p;
  return NB200_OK;
}
int nb200_p2p_open(const void* handle64, void** dev_ptr) {
  if (!handle64 || !dev_ptr) return NB200_ERR_ARG;
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  NB_CUDA_CHECK(cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return NB200_OK;
}
int nb200_p2p_close(void* dev_ptr) {
  if (!dev_ptr) return NB200_ERR_ARG;
  NB_CUDA_CHECK(cudaIpcCloseMemHandle(dev_ptr));
  return NB200_OK;
}
int nb200_p2p_free(void* dev_ptr) {
  if (!dev_ptr) return NB200_ERR_ARG;
  NB_CUDA_CHECK(cudaFree(dev_ptr));
  return NB200_OK;
}

}  // extern "C"
