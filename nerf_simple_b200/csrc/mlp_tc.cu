// NB200_BF16 precision: fused posenc + MLP on tcgen05 / TMEM (placeholder until the kernel lands).
#include "common.cuh"

namespace nb200 {
size_t tc_packed_bytes() { return 0; }
size_t tc_saved_bytes(int64_t) { return 0; }
size_t tc_scratch_bytes(int64_t, int) { return 0; }
int tc_pack_weights(const float* const*, void*, cudaStream_t) { return NB200_ERR_UNSUPPORTED; }
int tc_forward(int, const float*, const float*, int64_t, int, const void*, float*, void*, void*, size_t,
               cudaStream_t) { return NB200_ERR_UNSUPPORTED; }
int tc_backward(int, const float*, const float*, int64_t, int, const void*, const float*, const void*,
                float* const*, void*, size_t, cudaStream_t) { return NB200_ERR_UNSUPPORTED; }
}  // namespace nb200
