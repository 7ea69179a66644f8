// NB200_BF16 precision: fused positional-encoding + NeRF MLP on tcgen05 / TMEM (sm_100a).
//   forward  <- utils/xyz.py:16-36 + utils/nets.py:34-43 (+ utils/rendering.py:31-40 in rays mode)
//
// One persistent CTA per SM.  A CTA keeps TWO 128-sample tiles in flight (slots 0/1): while the
// epilogue warps of one slot turn the fp32 accumulator of layer l (TMEM) into the bf16 A operand
// of layer l+1 (shared memory, never HBM), the single MMA-issuing thread runs layer l of the other
// slot.  Weights are streamed per layer from the L2-resident packed image (pre-swizzled bf16
// UMMA operand slabs) by a TMA bulk-copy producer warp through a 2-stage ring.
//
//   warps 0-3 : encoder + epilogue of slot 0   (thread i <-> sample row i <-> TMEM lane i)
//   warps 4-7 : encoder + epilogue of slot 1
//   warp  8   : weight producer (cp.async.bulk -> mbarrier complete_tx)
//   warp  9   : TMEM allocator + MMA issuer (tcgen05.mma, cta_group::1, M=128, N=256/128, K=16)
//
// Shared memory (bytes, 1024-aligned):  A[2] 2x64 KB (128 x 256 bf16, 4 K-blocks of 128 B rows,
// SWIZZLE_128B) | E[2] 2x16 KB (encoded input: posx 63->64, later posd 27->32) | W ring 2x32 KB.
// TMEM: 512 columns = two 128x256 fp32 accumulators.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace nb200 {
using namespace tc;

// ------------------------------------------------------------------ packed weight image
constexpr int kNumMmaLayers = 10;  // L0_0..L0_4, skip, L1_0, L1_1, layers_2, color_fc.0
constexpr int kMaxSlabs = 48;

struct SlabDesc {
  uint32_t off;     // byte offset of the slab image in the packed buffer
  uint32_t bytes;   // n * 128
  uint16_t n;       // operand rows (output features): 256 or 128
  uint16_t wcol0;   // first weight column (input feature) covered
  uint16_t kvalid;  // valid columns (rest zero padded up to 64)
  uint16_t ldw;     // row pitch of the fp32 weight
  uint8_t ml;       // MMA layer index 0..9
  uint8_t layer;    // Layer enum (index into the 24 parameter pointers / 2)
  uint8_t src;      // A operand: 0 = activation buffer K-block `kb`, 1 = encoding buffer
  uint8_t kb;
  uint8_t ksteps;   // number of K=16 MMAs
  uint8_t first;    // first slab of its layer (accumulator is overwritten)
  uint8_t last;     // last slab of its layer (commit to acc_full)
  uint8_t transposed;  // image(n,k) = W[(wcol0+k)*ldw + n] (dgrad operand) instead of W[n*ldw + wcol0+k]
};

struct PackedLayout {
  int num_fwd;
  SlabDesc fwd[kMaxSlabs];
  int num_bwd;
  SlabDesc bwd[kMaxSlabs];  // dgrad chain: transposed weight images, 9 MMA layers
  uint32_t f32_off;  // fp32 section: bias[10][256] | wsig[256] | bsig(+pad 4) | wc1[3][128] | bc1[3](+pad)
  uint32_t total_bytes;
};
constexpr int kF32Bias = 0, kF32WSig = 2560, kF32BSig = 2816, kF32WC1 = 2820, kF32BC1 = 3204,
              kF32Floats = 3264;

__constant__ PackedLayout c_layout;
static PackedLayout h_layout;
static bool h_layout_ready = false;

__host__ __device__ constexpr int mma_layer_of(int ml) {
  return ml <= 4 ? L0_0 + ml : (ml == 5 ? L_SKIP : (ml <= 7 ? L1_0 + (ml - 6) : (ml == 8 ? L_2 : L_C0)));
}

struct LayoutBuilder {
  SlabDesc* arr;
  uint32_t off;
  int s;
  void add(int ml, int layer, int transposed, int n, int wcol0, int kvalid, int ldw, int src, int kb, int ksteps) {
    SlabDesc& d = arr[s++];
    d.off = off; d.bytes = (uint32_t)n * 128u; d.n = (uint16_t)n; d.wcol0 = (uint16_t)wcol0;
    d.kvalid = (uint16_t)kvalid; d.ldw = (uint16_t)ldw; d.ml = (uint8_t)ml; d.layer = (uint8_t)layer;
    d.src = (uint8_t)src; d.kb = (uint8_t)kb; d.ksteps = (uint8_t)ksteps; d.first = 0; d.last = 0;
    d.transposed = (uint8_t)transposed;
    off += d.bytes;
  }
};

// dgrad chain (mlp_bwd.cuh): MMA layer bl = 1..9 computes delta_in = delta_out @ W (W un-transposed
// [out,in]); its B operand image is W^T: image(n = in-feature, k = out-feature).
__host__ __device__ constexpr int bwd_layer_of(int bl) {
  return bl == 1 ? L_C0 : (bl == 2 ? L_2 : (bl == 3 ? L1_1 : (bl == 4 ? L1_0 : (bl == 5 ? L_SKIP : L0_4 - (bl - 6)))));
}

static void build_layout() {
  memset(&h_layout, 0, sizeof(h_layout));
  LayoutBuilder b;
  b.arr = h_layout.fwd; b.off = 0; b.s = 0;
  for (int ml = 0; ml < kNumMmaLayers; ++ml) {
    const int first = b.s, ly = mma_layer_of(ml);
    if (ml == 0) {
      b.add(ml, ly, 0, 256, 0, kPosX, kPosX, 1, 0, 4);
    } else if (ml == 5) {  // cat([h, posx]) (utils/nets.py:38)
      for (int kb = 0; kb < 4; ++kb) b.add(ml, ly, 0, 256, kb * 64, 64, kHidden + kPosX, 0, kb, 4);
      b.add(ml, ly, 0, 256, 256, kPosX, kHidden + kPosX, 1, 0, 4);
    } else if (ml == 9) {  // cat([g, posd]) (utils/nets.py:42), 128 outputs
      for (int kb = 0; kb < 4; ++kb) b.add(ml, ly, 0, 128, kb * 64, 64, kHidden + kPosD, 0, kb, 4);
      b.add(ml, ly, 0, 128, 256, kPosD, kHidden + kPosD, 1, 0, 2);
    } else {
      for (int kb = 0; kb < 4; ++kb) b.add(ml, ly, 0, 256, kb * 64, 64, kHidden, 0, kb, 4);
    }
    h_layout.fwd[first].first = 1;
    h_layout.fwd[b.s - 1].last = 1;
  }
  h_layout.num_fwd = b.s;
  // dgrad slabs: n = 256 input features (rows of the image), k-blocks over the output features
  LayoutBuilder t;
  t.arr = h_layout.bwd; t.off = b.off; t.s = 0;
  for (int bl = 1; bl <= 9; ++bl) {
    const int first = t.s, ly = bwd_layer_of(bl);
    const int ldw = (ly == L_C0) ? kHidden + kPosD : (ly == L_SKIP ? kHidden + kPosX : kHidden);
    const int nkb = (bl == 1) ? 2 : 4;  // color_fc.0 has 128 outputs
    for (int kb = 0; kb < nkb; ++kb) t.add(bl, ly, 1, 256, kb * 64, 64, ldw, 0, kb, 4);
    h_layout.bwd[first].first = 1;
    h_layout.bwd[t.s - 1].last = 1;
  }
  h_layout.num_bwd = t.s;
  h_layout.f32_off = t.off;
  h_layout.total_bytes = t.off + kF32Floats * (uint32_t)sizeof(float);
}

static int ensure_layout() {
  if (h_layout_ready) return NB200_OK;
  build_layout();
  NB_CUDA_CHECK(cudaMemcpyToSymbol(c_layout, &h_layout, sizeof(PackedLayout)));
  h_layout_ready = true;
  return NB200_OK;
}

struct ParamPtrs { const float* p[24]; };

// One block per slab: fp32 weight -> bf16 SWIZZLE_128B operand image [n rows x 64 k-columns].
__global__ void __launch_bounds__(256) pack_slabs_kernel(ParamPtrs P, uint8_t* __restrict__ packed) {
  const bool is_bwd = (int)blockIdx.x >= c_layout.num_fwd;
  const SlabDesc d = is_bwd ? c_layout.bwd[blockIdx.x - c_layout.num_fwd] : c_layout.fwd[blockIdx.x];
  const float* W = P.p[2 * d.layer];
  for (int item = threadIdx.x; item < d.n * 8; item += blockDim.x) {
    const int n = item >> 3, j = item & 7;
    uint32_t w[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float v[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int k = j * 8 + 2 * e + h;
        v[h] = 0.f;
        if (k < d.kvalid)
          v[h] = d.transposed ? __ldg(W + (size_t)(d.wcol0 + k) * d.ldw + n) : __ldg(W + (size_t)n * d.ldw + d.wcol0 + k);
      }
      w[e] = pack_bf16x2(v[0], v[1]);
    }
    *reinterpret_cast<uint4*>(packed + d.off + sw128_off(n, j)) = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

__global__ void __launch_bounds__(256) pack_f32_kernel(ParamPtrs P, float* __restrict__ f) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= kF32Floats) return;
  float v = 0.f;
  if (i < kF32WSig) {
    const int ml = i >> 8, j = i & 255;
    const int layer = mma_layer_of(ml);
    const int n = (ml == 9) ? 128 : 256;
    v = (j < n) ? P.p[2 * layer + 1][j] : 0.f;
  } else if (i < kF32BSig) {
    v = P.p[2 * L_SIGMA][i - kF32WSig];
  } else if (i == kF32BSig) {
    v = P.p[2 * L_SIGMA + 1][0];
  } else if (i >= kF32WC1 && i < kF32WC1 + 384) {
    v = P.p[2 * L_C1][i - kF32WC1];
  } else if (i >= kF32BC1 && i < kF32BC1 + 3) {
    v = P.p[2 * L_C1 + 1][i - kF32BC1];
  }
  f[i] = v;
}

// ------------------------------------------------------------------------ forward kernel
constexpr int kTileM = 128;
constexpr uint32_t kABytes = 65536, kEBytes = 16384, kWStageBytes = 32768;
constexpr int kNumWStages = 2;
constexpr uint32_t kSmemA = 0, kSmemE = 2 * kABytes, kSmemW = kSmemE + 2 * kEBytes,
                   kSmemBar = kSmemW + kNumWStages * kWStageBytes;  // 229376
constexpr uint32_t kSmemTotal = kSmemBar + 128;
constexpr uint32_t kSmemLaunch = kSmemTotal + 1024;  // slack for manual 1024 B alignment
constexpr int kFwdThreads = 320;

// saved activations (training): tensors 0..7 = h0..h7, 8 = g (256 cols, 64 KB per tile), 9 = c1
// (128 cols, 32 KB per tile), 10 = posx (64 cols, 16 KB), 11 = posd (32 of 64 cols, 16 KB).  Every
// tile is stored as [K-block][128 rows x 128 B SWIZZLE_128B], i.e. exactly the UMMA operand image
// the backward kernels bulk-copy back into shared memory (K-major for dgrad, MN-major for wgrad).
constexpr size_t kSavedTileBytes = 9 * 65536 + 32768 + 16384 + 16384;
__host__ __device__ __forceinline__ size_t saved_tensor_off(int t, int64_t num_tiles) {
  const size_t per_tile = t < 9 ? (size_t)t * 65536 : (t == 9 ? 9 * 65536 : (t == 10 ? 9 * 65536 + 32768 : 9 * 65536 + 49152));
  return per_tile * (size_t)num_tiles;
}
__host__ __device__ __forceinline__ size_t saved_tile_bytes(int t) {
  return t < 9 ? 65536 : (t == 9 ? 32768 : 16384);
}

struct FwdParams {
  int in_mode;
  const float* in0;
  const float* in1;
  int64_t M;
  int N;
  const uint8_t* packed;
  float* out;
  uint8_t* saved;  // null for inference
  int64_t num_tiles;
};

__device__ __forceinline__ void load_query_tc(const FwdParams& p, int64_t m, float v[6]) {
  if (p.in_mode == NB200_IN_POINTS) {
    const float2* q = reinterpret_cast<const float2*>(p.in0 + m * 6);
    const float2 a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y;
  } else {
    const int64_t ray = m / p.N;
    const float2* q = reinterpret_cast<const float2*>(p.in0 + ray * 6);
    const float2 a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2);
    const float t = __ldg(p.in1 + m);
    const float dx = b.y, dy = c.x, dz = c.y;
    v[0] = __fadd_rn(a.x, __fmul_rn(dx, t));   // utils/rendering.py:34-36 (d un-normalised)
    v[1] = __fadd_rn(a.y, __fmul_rn(dy, t));
    v[2] = __fadd_rn(b.x, __fmul_rn(dz, t));
    const float inv = 1.0f / sqrtf(fmaf(dz, dz, fmaf(dy, dy, dx * dx)));   // :37
    v[3] = dx * inv; v[4] = dy * inv; v[5] = dz * inv;
  }
}

// Encode 3 coordinates with L levels into the bf16 operand row `r` of a SWIZZLE_128B image:
// cols [x0,x1,x2, per coordinate: sin(2^i x), cos(2^i x) ...], zero padded to NCH*8 columns.
// Level 0 uses the accurate sincosf; higher levels the double-angle recurrence (abs. error
// <= 2^i * 1e-7, far below bf16 resolution).
template <int L, int J0, int J1>
__device__ __forceinline__ void encode_row(const float* x, uint32_t img_base, uint32_t r, uint8_t* gsave) {
  // 64 columns = 8 chunks of 16 bytes; this call stores chunks [J0, J1) (columns beyond 3+6L are zero)
  float f[64];
#pragma unroll
  for (int i = 0; i < 64; ++i) f[i] = 0.f;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    f[c] = x[c];
    float s, co;
    sincosf(x[c], &s, &co);
#pragma unroll
    for (int i = 0; i < L; ++i) {
      f[3 + c * 2 * L + 2 * i] = s;
      f[3 + c * 2 * L + 2 * i + 1] = co;
      const float s2 = 2.f * s * co;
      co = fmaf(-2.f * s, s, 1.f);
      s = s2;
    }
  }
#pragma unroll
  for (int j = J0; j < J1; ++j) {
    const uint32_t w0 = pack_bf16x2(f[8 * j], f[8 * j + 1]), w1 = pack_bf16x2(f[8 * j + 2], f[8 * j + 3]),
                   w2 = pack_bf16x2(f[8 * j + 4], f[8 * j + 5]), w3 = pack_bf16x2(f[8 * j + 6], f[8 * j + 7]);
    st_shared_v4(img_base + sw128_off(r, j), w0, w1, w2, w3);
    if (gsave) *reinterpret_cast<uint4*>(gsave + sw128_off(r, j)) = make_uint4(w0, w1, w2, w3);
  }
}

template <bool kSave>
__global__ void __launch_bounds__(kFwdThreads, 1) mlp_fwd_tc_kernel(const FwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + kSmemBar;
  // barriers: w_full[2] @0, w_empty[2] @16, act_ready[2] @32, acc_full[2] @48, tmem ptr @64
  const uint32_t bar_wfull = bar_base, bar_wempty = bar_base + 16, bar_act = bar_base + 32,
                 bar_acc = bar_base + 48, tmem_slot = bar_base + 64;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kNumWStages; ++i) {
      mbar_init(bar_wfull + 8 * i, 1);
      mbar_init(bar_wempty + 8 * i, 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar_act + 8 * s, 128);
      mbar_init(bar_acc + 8 * s, 1);
    }
    fence_mbar_init();
  }
  if (warp == 9) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  const int64_t T = p.num_tiles;
  const int64_t G = gridDim.x;
  const int64_t my_tiles = (blockIdx.x < T) ? (T - blockIdx.x + G - 1) / G : 0;
  const float* f32sec = reinterpret_cast<const float*>(p.packed + c_layout.f32_off);

  if (warp < 8) {
    // ===================== encoder + epilogue warpgroup of one slot =====================
    const int slot = warp >> 2;
    const uint32_t r = threadIdx.x & 127;  // row in tile == TMEM lane
    const uint32_t a_img = smem_base + kSmemA + slot * kABytes;
    const uint32_t e_img = smem_base + kSmemE + slot * kEBytes;
    const uint32_t t_lane = tmem_base + (((uint32_t)(warp & 3) * 32u) << 16) + (uint32_t)slot * 256u;
    uint32_t acc_parity = 0;
    for (int64_t k = slot; k < my_tiles; k += 2) {
      const int64_t tile = blockIdx.x + k * G;
      const int64_t m_raw = tile * kTileM + r;
      const bool row_valid = m_raw < p.M;
      const int64_t m = row_valid ? m_raw : p.M - 1;
      float v[6];
      load_query_tc(p, m, v);
      encode_row<kLp, 0, 8>(v, e_img, r,   // posx -> E[slot], K = 64
                         kSave ? p.saved + saved_tensor_off(10, T) + (size_t)tile * 16384 : nullptr);
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(bar_act + 8 * slot);
      float sigma = 0.f;
      for (int ml = 0; ml < kNumMmaLayers; ++ml) {
        mbar_wait(bar_acc + 8 * slot, acc_parity, 100 + ml);
        acc_parity ^= 1;
        tc_fence_after();
        const float* bias = f32sec + kF32Bias + ml * 256;
        if (ml == 5) {
          // posx has been consumed by the skip layer: the encoding buffer now carries posd
          encode_row<kLd, 0, 8>(v + 3, e_img, r,   // cols 27..63 zero: wgrad reads the image with N = 64
                             kSave ? p.saved + saved_tensor_off(11, T) + (size_t)tile * 16384 : nullptr);
        }
        if (ml < 9) {
          const bool relu = (ml != 8);  // layers_2 has no activation (utils/nets.py:28,41)
          uint8_t* gsave = nullptr;
          if (kSave) gsave = p.saved + saved_tensor_off(ml, T) + (size_t)tile * 65536;
#pragma unroll 1
          for (int c = 0; c < 8; ++c) {  // 8 chunks of 32 accumulator columns
            uint32_t acc[32];
            tmem_ld32(t_lane + c * 32, acc);
            tmem_ld_wait();
            float x[32];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + c * 32) + q);
              x[4 * q] = __uint_as_float(acc[4 * q]) + b4.x;
              x[4 * q + 1] = __uint_as_float(acc[4 * q + 1]) + b4.y;
              x[4 * q + 2] = __uint_as_float(acc[4 * q + 2]) + b4.z;
              x[4 * q + 3] = __uint_as_float(acc[4 * q + 3]) + b4.w;
            }
            if (relu) {
#pragma unroll
              for (int i = 0; i < 32; ++i) x[i] = fmaxf(x[i], 0.f);
            }
            if (ml == 7) {  // sigma head reads the layers_1 output (utils/nets.py:40)
              const float* ws = f32sec + kF32WSig + c * 32;
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                const float4 w4 = __ldg(reinterpret_cast<const float4*>(ws) + q);
                sigma = fmaf(x[4 * q], w4.x, sigma);
                sigma = fmaf(x[4 * q + 1], w4.y, sigma);
                sigma = fmaf(x[4 * q + 2], w4.z, sigma);
                sigma = fmaf(x[4 * q + 3], w4.w, sigma);
              }
            }
            const uint32_t kb = c >> 1;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const uint32_t j = (c & 1) * 4 + q;
              const uint32_t w0 = pack_bf16x2(x[8 * q], x[8 * q + 1]), w1 = pack_bf16x2(x[8 * q + 2], x[8 * q + 3]),
                             w2 = pack_bf16x2(x[8 * q + 4], x[8 * q + 5]), w3 = pack_bf16x2(x[8 * q + 6], x[8 * q + 7]);
              const uint32_t o = kb * 16384u + sw128_off(r, j);
              st_shared_v4(a_img + o, w0, w1, w2, w3);
              if (kSave) *reinterpret_cast<uint4*>(gsave + o) = make_uint4(w0, w1, w2, w3);
            }
          }
          fence_proxy_async_smem();
          tc_fence_before();
          mbar_arrive(bar_act + 8 * slot);
        } else {
          // color_fc.0 epilogue (128 columns) + color_fc.2 (128 -> 3) on CUDA cores
          uint8_t* gsave = nullptr;
          if (kSave) gsave = p.saved + saved_tensor_off(9, T) + (size_t)tile * 32768;
          float rgb[3] = {0.f, 0.f, 0.f};
          const float* wc1 = f32sec + kF32WC1;
#pragma unroll 1
          for (int c = 0; c < 4; ++c) {
            uint32_t acc[32];
            tmem_ld32(t_lane + c * 32, acc);
            tmem_ld_wait();
            float x[32];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + c * 32) + q);
              x[4 * q] = fmaxf(__uint_as_float(acc[4 * q]) + b4.x, 0.f);
              x[4 * q + 1] = fmaxf(__uint_as_float(acc[4 * q + 1]) + b4.y, 0.f);
              x[4 * q + 2] = fmaxf(__uint_as_float(acc[4 * q + 2]) + b4.z, 0.f);
              x[4 * q + 3] = fmaxf(__uint_as_float(acc[4 * q + 3]) + b4.w, 0.f);
            }
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) {
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                const float4 w4 = __ldg(reinterpret_cast<const float4*>(wc1 + ch * 128 + c * 32) + q);
                rgb[ch] = fmaf(x[4 * q], w4.x, rgb[ch]);
                rgb[ch] = fmaf(x[4 * q + 1], w4.y, rgb[ch]);
                rgb[ch] = fmaf(x[4 * q + 2], w4.z, rgb[ch]);
                rgb[ch] = fmaf(x[4 * q + 3], w4.w, rgb[ch]);
              }
            }
            if (kSave) {
              const uint32_t kb = c >> 1;
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const uint32_t j = (c & 1) * 4 + q;
                *reinterpret_cast<uint4*>(gsave + kb * 16384u + sw128_off(r, j)) =
                    make_uint4(pack_bf16x2(x[8 * q], x[8 * q + 1]), pack_bf16x2(x[8 * q + 2], x[8 * q + 3]),
                               pack_bf16x2(x[8 * q + 4], x[8 * q + 5]), pack_bf16x2(x[8 * q + 6], x[8 * q + 7]));
              }
            }
          }
          if (row_valid) {
            const float bs = __ldg(f32sec + kF32BSig);
            reinterpret_cast<float4*>(p.out)[m_raw] =
                make_float4(rgb[0] + __ldg(f32sec + kF32BC1), rgb[1] + __ldg(f32sec + kF32BC1 + 1),
                            rgb[2] + __ldg(f32sec + kF32BC1 + 2), sigma + bs);  // (r,g,b,sigma) :43
          }
          tc_fence_before();  // TMEM reads of this tile are ordered before the next act_ready arrive
        }
      }
    }
  } else if (warp == 8) {
    // ================================ weight producer ================================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      const int nslabs = c_layout.num_fwd;
      for (int64_t pr = 0; pr * 2 < my_tiles; ++pr) {
        const int nslots = (my_tiles - 2 * pr >= 2) ? 2 : 1;
        int s0 = 0;
        for (int ml = 0; ml < kNumMmaLayers; ++ml) {
          int s1 = s0;
          while (!c_layout.fwd[s1].last) ++s1;
          for (int slot = 0; slot < nslots; ++slot) {
            for (int s = s0; s <= s1; ++s) {
              const uint32_t off = c_layout.fwd[s].off, bytes = c_layout.fwd[s].bytes;
              mbar_wait(bar_wempty + 8 * stage, phase ^ 1, 200);
              mbar_arrive_expect_tx(bar_wfull + 8 * stage, bytes);
              tma_bulk_g2s(smem_base + kSmemW + stage * kWStageBytes, p.packed + off, bytes,
                           bar_wfull + 8 * stage);
              if (++stage == kNumWStages) { stage = 0; phase ^= 1; }
            }
          }
          s0 = s1 + 1;
        }
      }
      (void)nslabs;
    }
  } else {
    // ================================== MMA issuer ==================================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      uint32_t act_parity[2] = {0, 0};
      const uint32_t idesc256 = umma_idesc_bf16(128, 256, 0, 0), idesc128 = umma_idesc_bf16(128, 128, 0, 0);
      for (int64_t pr = 0; pr * 2 < my_tiles; ++pr) {
        const int nslots = (my_tiles - 2 * pr >= 2) ? 2 : 1;
        int s0 = 0;
        for (int ml = 0; ml < kNumMmaLayers; ++ml) {
          int s1 = s0;
          while (!c_layout.fwd[s1].last) ++s1;
          for (int slot = 0; slot < nslots; ++slot) {
            mbar_wait(bar_act + 8 * slot, act_parity[slot], 300 + ml);
            act_parity[slot] ^= 1;
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + (uint32_t)slot * 256u;
            for (int s = s0; s <= s1; ++s) {
              const SlabDesc& d = c_layout.fwd[s];
              mbar_wait(bar_wfull + 8 * stage, phase, 400);
              tc_fence_after();
              const uint32_t a_addr = d.src ? (smem_base + kSmemE + slot * kEBytes)
                                            : (smem_base + kSmemA + slot * kABytes + d.kb * 16384u);
              const uint32_t b_addr = smem_base + kSmemW + stage * kWStageBytes;
              const uint32_t idesc = (d.n == 256) ? idesc256 : idesc128;
              const int ksteps = d.ksteps;
              for (int kk = 0; kk < ksteps; ++kk) {
                umma_bf16(d_tmem, umma_smem_desc(a_addr + kk * 32, 16, 1024),
                          umma_smem_desc(b_addr + kk * 32, 16, 1024), idesc, (d.first && kk == 0) ? 0u : 1u);
              }
              umma_commit(bar_wempty + 8 * stage);  // slab may be overwritten once these MMAs retire
              if (++stage == kNumWStages) { stage = 0; phase ^= 1; }
            }
            umma_commit(bar_acc + 8 * slot);  // accumulator of (slot, layer) complete
          }
          s0 = s1 + 1;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

#include "mlp_tc_bwd.cuh"
#include "mlp_chain.cuh"

// ------------------------------------------------------------------------------ host API
size_t tc_packed_bytes() {
  if (!h_layout_ready) build_layout();  // layout is host-computable without a device
  return h_layout.total_bytes;
}
size_t tc_saved_bytes(int64_t M) { return (size_t)ceil_div64(M, kTileM) * kSavedTileBytes; }
size_t tc_scratch_bytes(int64_t M, int train) { return train ? (size_t)ceil_div64(M, kTileM) * kDeltaTileBytes : 0; }

int tc_pack_weights(const float* const* P, void* packed, cudaStream_t s) {
  NB_TRY_RC(ensure_layout());
  ParamPtrs pp;
  for (int i = 0; i < 24; ++i) pp.p[i] = P[i];
  pack_slabs_kernel<<<h_layout.num_fwd + h_layout.num_bwd, 256, 0, s>>>(pp, reinterpret_cast<uint8_t*>(packed));
  NB_LAUNCH_CHECK("pack_slabs_kernel");
  pack_f32_kernel<<<(kF32Floats + 255) / 256, 256, 0, s>>>(
      pp, reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(packed) + h_layout.f32_off));
  NB_LAUNCH_CHECK("pack_f32_kernel");
  return NB200_OK;
}

static int check_arch() {
  static int arch = 0;
  if (arch == 0) arch = nb200_device_arch();
  if (arch < 0) return NB200_ERR_CUDA;
  return (arch / 10 == 10) ? NB200_OK : NB200_ERR_ARCH;
}

static bool use_v1() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("NB200_TC_V1"); v = (e && e[0] == '1') ? 1 : 0; }
  return v == 1;
}

// biases / head weights of the net being run go to the constant bank (stream-ordered D2D copy)
static int upload_consts(const void* packed, cudaStream_t s) {
  NB_CUDA_CHECK(cudaMemcpyToSymbolAsync(c_f32, reinterpret_cast<const uint8_t*>(packed) + h_layout.f32_off,
                                        kF32Floats * sizeof(float), 0, cudaMemcpyDeviceToDevice, s));
  return NB200_OK;
}

// Tensor maps over the packed weight image (plain [rows x 64 bf16] view, no TMA swizzle: the image
// is pre-swizzled), cached per packed buffer.  cuTensorMapEncodeTiled is fetched through the runtime
// so the library does not link against libcuda.
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
struct TmapPair { const void* packed; CUtensorMap m128, m64; };
static int get_tmaps(const void* packed, const TmapPair** out) {
  static TmapPair cache[8];
  static int used = 0, next = 0;
  for (int i = 0; i < used; ++i)
    if (cache[i].packed == packed) { *out = &cache[i]; return NB200_OK; }
  static EncodeTiledFn encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    NB_CUDA_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess) return NB200_ERR_CUDA;
    encode = reinterpret_cast<EncodeTiledFn>(fn);
  }
  TmapPair& t = cache[next];
  next = (next + 1) % 8;
  if (used < 8) ++used;
  t.packed = nullptr;
  const cuuint64_t gdim[2] = {64, (cuuint64_t)(h_layout.f32_off / 128)};
  const cuuint64_t gstride[1] = {128};
  const cuuint32_t estride[2] = {1, 1};
  for (int k = 0; k < 2; ++k) {
    const cuuint32_t box[2] = {64, k == 0 ? 128u : 64u};
    CUresult r = encode(k == 0 ? &t.m128 : &t.m64, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(packed), gdim, gstride,
                        box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      snprintf(g_last_cuda_error, sizeof(g_last_cuda_error), "cuTensorMapEncodeTiled failed (%d)", (int)r);
      return NB200_ERR_CUDA;
    }
  }
  t.packed = packed;
  *out = &t;
  return NB200_OK;
}

static int chain_grid(int64_t T) {
  const int64_t PT = (T + 1) / 2;         // 256-row pair-tiles
  const int64_t want = (PT + 1) / 2;      // two pair-tiles per cluster keep the ping-pong busy
  const int64_t maxc = sm_count() / 2;
  const int64_t clusters = want < maxc ? (want > 0 ? want : 1) : maxc;
  return (int)(2 * clusters);
}

int tc_forward(int in_mode, const float* in0, const float* in1, int64_t M, int N, const void* packed,
               float* out, void* saved, void*, size_t, cudaStream_t s) {
  NB_TRY_RC(check_arch());
  NB_TRY_RC(ensure_layout());
  static bool attr_set = false;
  if (!attr_set) {
    NB_CUDA_CHECK(cudaFuncSetAttribute(mlp_fwd_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)kSmemLaunch));
    NB_CUDA_CHECK(cudaFuncSetAttribute(mlp_fwd_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)kSmemLaunch));
    NB_CUDA_CHECK(cudaFuncSetAttribute(chain_kernel<FwdEpi<false>>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)kCSmemLaunch));
    NB_CUDA_CHECK(cudaFuncSetAttribute(chain_kernel<FwdEpi<true>>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)kCSmemLaunch));
    attr_set = true;
  }
  const int64_t T = ceil_div64(M, kTileM);
  if (use_v1()) {
    FwdParams p;
    p.in_mode = in_mode; p.in0 = in0; p.in1 = in1; p.M = M; p.N = N;
    p.packed = reinterpret_cast<const uint8_t*>(packed);
    p.out = out; p.saved = reinterpret_cast<uint8_t*>(saved);
    p.num_tiles = T;
    const int64_t want = (p.num_tiles + 1) / 2;  // two tiles per CTA keep the ping-pong busy
    const int grid = (int)(want < sm_count() ? (want > 0 ? want : 1) : sm_count());
    if (saved)
      mlp_fwd_tc_kernel<true><<<grid, kFwdThreads, kSmemLaunch, s>>>(p);
    else
      mlp_fwd_tc_kernel<false><<<grid, kFwdThreads, kSmemLaunch, s>>>(p);
    NB_LAUNCH_CHECK("mlp_fwd_tc_kernel");
    return NB200_OK;
  }
  NB_TRY_RC(upload_consts(packed, s));
  const TmapPair* tm = nullptr;
  NB_TRY_RC(get_tmaps(packed, &tm));
  FwdEpiParams p;
  p.tmap128 = tm->m128; p.tmap64 = tm->m64;
  { const char* e = getenv("NB200_DBG"); p.dbg = e ? atoi(e) : 0; }
  p.dbg_counters = nullptr;
  if (p.dbg & 8) {
    static unsigned long long* ctr = nullptr;
    if (!ctr) { cudaMalloc(&ctr, 64); }
    unsigned long long h[4];
    cudaMemcpy(h, ctr, 32, cudaMemcpyDeviceToHost);   // counters of the previous launch (debug only; syncs)
    printf("nb200 dbg: mma-warp cycles wait_act=%llu wait_wfull=%llu wait_wpeer=%llu total=%llu\n", h[0], h[1], h[2], h[3]);
    cudaMemset(ctr, 0, 64);
    p.dbg_counters = ctr;
  }
  p.in_mode = in_mode; p.in0 = in0; p.in1 = in1; p.M = M; p.N = N;
  p.packed = reinterpret_cast<const uint8_t*>(packed);
  p.out = out; p.saved = reinterpret_cast<uint8_t*>(saved);
  p.num_tiles = T;
  const int grid = chain_grid(T);
  if (saved)
    chain_kernel<FwdEpi<true>><<<grid, kCThreads, kCSmemLaunch, s>>>(p);
  else
    chain_kernel<FwdEpi<false>><<<grid, kCThreads, kCSmemLaunch, s>>>(p);
  NB_LAUNCH_CHECK("chain_kernel<FwdEpi>");
  return NB200_OK;
}

int tc_backward(int, const float*, const float*, int64_t M, int, const void* packed, const float* d_out,
                const void* saved, float* const* G, void* scratch, size_t scratch_bytes, cudaStream_t s) {
  NB_TRY_RC(check_arch());
  NB_TRY_RC(ensure_layout());
  if (!scratch || scratch_bytes < tc_scratch_bytes(M, 1)) return NB200_ERR_WORKSPACE;
  static bool attr_set = false;
  if (!attr_set) {
    NB_CUDA_CHECK(cudaFuncSetAttribute(mlp_dgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemLaunch));
    NB_CUDA_CHECK(cudaFuncSetAttribute(mlp_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kWgSmemLaunch));
    NB_CUDA_CHECK(cudaFuncSetAttribute(chain_kernel<DgradEpi>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kCSmemLaunch));
    attr_set = true;
  }
  const int64_t T = ceil_div64(M, kTileM);
  const uint8_t* sv = reinterpret_cast<const uint8_t*>(saved);
  uint8_t* ds = reinterpret_cast<uint8_t*>(scratch);
  // 1. fused delta chain
  BwdParams bp;
  bp.dbg = 0;
  bp.M = M; bp.num_tiles = T; bp.packed = reinterpret_cast<const uint8_t*>(packed); bp.saved = sv;
  bp.d_out = d_out; bp.dscr = ds;
  if (use_v1()) {
    const int64_t want = (T + 1) / 2;
    const int grid = (int)(want < sm_count() ? (want > 0 ? want : 1) : sm_count());
    mlp_dgrad_tc_kernel<<<grid, kFwdThreads, kSmemLaunch, s>>>(bp);
    NB_LAUNCH_CHECK("mlp_dgrad_tc_kernel");
  } else {
    NB_TRY_RC(upload_consts(packed, s));
    const TmapPair* tm = nullptr;
    NB_TRY_RC(get_tmaps(packed, &tm));
    bp.tmap128 = tm->m128; bp.tmap64 = tm->m64;
    chain_kernel<DgradEpi><<<chain_grid(T), kCThreads, kCSmemLaunch, s>>>(bp);
    NB_LAUNCH_CHECK("chain_kernel<DgradEpi>");
  }
  // 2. weight gradients: (delta tensor, input tensor) pairs
  WgradParams wp;
  memset(&wp, 0, sizeof(wp));
  wp.T = T;
  int n = 0;
  auto item = [&](int dt, int st, int layer, int ldw, int col0, int ncols, int nrows, bool bias) {
    WItem& w = wp.items[n++];
    w.a_ptr = ds + delta_tensor_off(dt, T);
    w.a_tile_bytes = dt == 0 ? 32768u : 65536u;
    w.a_chunks = dt == 0 ? 2 : 4;
    w.b_ptr = sv + saved_tensor_off(st, T);
    w.b_tile_bytes = (uint32_t)saved_tile_bytes(st);
    w.b_chunks = st >= 10 ? 1 : 4;
    w.n_mma = st >= 10 ? 64 : 256;
    w.dW = G[2 * layer]; w.db = bias ? G[2 * layer + 1] : nullptr;
    w.ldw = ldw; w.col0 = col0; w.ncols = ncols; w.nrows = nrows;
    w.cost = (w.a_chunks + w.b_chunks) * 16;
  };
  item(0, 8, L_C0, kHidden + kPosD, 0, kHidden, kHidden / 2, true);        // color_fc.0 <- g
  item(0, 11, L_C0, kHidden + kPosD, kHidden, kPosD, kHidden / 2, false);  // color_fc.0 <- posd
  item(1, 7, L_2, kHidden, 0, kHidden, kHidden, true);                      // layers_2   <- h7
  item(2, 6, L1_1, kHidden, 0, kHidden, kHidden, true);                     // layers_1.2 <- h6
  item(3, 5, L1_0, kHidden, 0, kHidden, kHidden, true);                     // layers_1.0 <- h5
  item(4, 4, L_SKIP, kHidden + kPosX, 0, kHidden, kHidden, true);           // skip       <- h4
  item(4, 10, L_SKIP, kHidden + kPosX, kHidden, kPosX, kHidden, false);     // skip       <- posx
  item(5, 3, L0_4, kHidden, 0, kHidden, kHidden, true);                     // layers_0.8 <- h3
  item(6, 2, L0_3, kHidden, 0, kHidden, kHidden, true);
  item(7, 1, L0_2, kHidden, 0, kHidden, kHidden, true);
  item(8, 0, L0_1, kHidden, 0, kHidden, kHidden, true);
  item(9, 10, L0_0, kPosX, 0, kPosX, kHidden, true);                        // layers_0.0 <- posx
  wp.num_items = n;
  mlp_wgrad_tc_kernel<<<sm_count(), kWgThreads, kWgSmemLaunch, s>>>(wp);
  NB_LAUNCH_CHECK("mlp_wgrad_tc_kernel");
  // 3. sigma / colour heads
  {
    const int64_t want = ceil_div64(T * 8, 8 * 4);  // ~4 (tile, row-group) items per warp
    const int64_t cap = (int64_t)sm_count() * 4;
    mlp_head_grads_kernel<<<(unsigned)(want < cap ? (want > 0 ? want : 1) : cap), 256, 0, s>>>(
        sv, d_out, M, T, G[2 * L_SIGMA], G[2 * L_SIGMA + 1], G[2 * L_C1], G[2 * L_C1 + 1]);
    NB_LAUNCH_CHECK("mlp_head_grads_kernel");
  }
  return NB200_OK;
}

}  // namespace nb200
