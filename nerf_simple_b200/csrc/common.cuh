// Shared helpers for libnerf_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/nerf_b200.h"

namespace nb200 {

extern thread_local char g_last_cuda_error[256];

int record_cuda_error(cudaError_t e, const char* what);

#define NB_CUDA_CHECK(expr)                                                        \
  do {                                                                             \
    cudaError_t _e = (expr);                                                       \
    if (_e != cudaSuccess) return ::nb200::record_cuda_error(_e, #expr);           \
  } while (0)

#define NB_LAUNCH_CHECK(name)                                                      \
  do {                                                                             \
    cudaError_t _e = cudaGetLastError();                                           \
    if (_e != cudaSuccess) return ::nb200::record_cuda_error(_e, name);            \
  } while (0)

#define NB_TRY_RC(expr)                 \
  do {                                  \
    int _rc = (expr);                   \
    if (_rc != NB200_OK) return _rc;    \
  } while (0)

static inline cudaStream_t as_stream(nb200_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// Number of SMs of the current device (cached).
int sm_count();

// The 12 Linear layers in state_dict order: index into the 24-pointer params array is 2*l
// (weight) and 2*l+1 (bias).  utils/nets.py:16-32.
enum Layer {
  L0_0 = 0, L0_1, L0_2, L0_3, L0_4,  // layers_0.{0,2,4,6,8}
  L_SKIP,                            // skip_conn_layer.0   (in = 256 + 63)
  L1_0, L1_1,                        // layers_1.{0,2}
  L_SIGMA,                           // sigma_fc.0          (256 -> 1)
  L_2,                               // layers_2            (256 -> 256, no activation)
  L_C0,                              // color_fc.0          (256 + 27 -> 128)
  L_C1,                              // color_fc.2          (128 -> 3)
  NUM_LAYERS
};

constexpr int kHidden = 256;
constexpr int kLp = 10, kLd = 4;
constexpr int kPosX = 3 + 6 * kLp;  // 63
constexpr int kPosD = 3 + 6 * kLd;  // 27
constexpr int kPosXPad = 64, kPosDPad = 32;

}  // namespace nb200
