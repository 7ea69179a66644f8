// Probe: tcgen05.ld throughput per SM as a function of the number of warps and the load shape.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tmem_ld_bw tmem_ld_bw.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ void ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]) : "r"(taddr) : "memory");
}
template <int MODE>   // 0: x16 one in flight, 1: x16 two in flight, 2: x32 one in flight
__global__ void probe(unsigned long long* out, uint32_t* sink, int iters) {
  __shared__ uint32_t tm;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&tm)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = tm + (((uint32_t)(warp & 3) * 32u) << 16) + (uint32_t)((warp >> 2) & 3) * 128u;
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  if (MODE == 0) {
    uint32_t a[16];
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int q = 0; q < 8; ++q) { ld16(base + q * 16, a); asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); acc += a[0] ^ a[15]; }
    }
  } else if (MODE == 1) {
    uint32_t a[16], b[16];
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int q = 0; q < 8; q += 2) {
        ld16(base + q * 16, a); ld16(base + q * 16 + 16, b);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); acc += a[0] ^ a[15] ^ b[3] ^ b[12];
      }
    }
  } else {
    uint32_t a[32];
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int q = 0; q < 4; ++q) { ld32(base + q * 32, a); asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); acc += a[0] ^ a[31]; }
    }
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) out[0] = (unsigned long long)(t1 - t0);
  sink[threadIdx.x] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512) : "memory");
}
int main() {
  unsigned long long* d; uint32_t* sink;
  cudaMalloc(&d, 8); cudaMalloc(&sink, 4096);
  const int iters = 2000;
  for (int mode = 0; mode < 3; ++mode)
    for (int warps : {1, 4, 8, 16}) {
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) probe<0><<<1, warps * 32>>>(d, sink, iters);
        if (mode == 1) probe<1><<<1, warps * 32>>>(d, sink, iters);
        if (mode == 2) probe<2><<<1, warps * 32>>>(d, sink, iters);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      }
      unsigned long long c; cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
      const double bytes = (double)warps * iters * 128.0 * 32 * 4;   // 128 columns x 32 lanes x 4 B per warp-iteration
      printf("mode %d (%s) warps %2d: %llu clk, %.1f B/clk per SM, %.1f clk per 128-column pass\n", mode,
             mode == 0 ? "x16, 1 in flight" : mode == 1 ? "x16, 2 in flight" : "x32, 1 in flight", warps, c, bytes / c, (double)c / iters);
    }
  return 0;
}
