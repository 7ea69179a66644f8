// Alpha compositing along rays, forward and analytic backward (HBM-bound).
//   nb200_composite_forward  <- utils/rendering.py:60-85 (volume_render)
//   nb200_composite_backward <- autograd of the same lines
//
// Layout: one warp per ray; lane l owns the S = ceil(N/32) CONSECUTIVE samples l*S .. l*S+S-1
// (16*S contiguous bytes of (r,g,b,sigma) and 4*S of ts per lane, vector loads).  The exclusive
// cumprod transmittance is S-1 serial multiplies per lane plus ONE 5-step warp-shuffle product
// scan per ray; t[i+1] is a register except at the lane boundary.  N <= 256 keeps all per-sample
// state in registers (the backward needs a second, reverse sweep).  Larger N falls back to a
// thread-per-ray kernel.
//
// Algorithmic bytes: fwd 20 B/sample read + 20 B/ray written (+8 B/sample when
// alpha/weights are requested); bwd 36 B/sample (20 read + 16 written) + 12..20 B/ray.
#include "common.cuh"
#include "stream_common.cuh"

namespace nb200 {

constexpr int kWarpsPerBlock = 8;
// forward tuning knobs (rays per warp iteration, resident blocks per SM) for N = 64 / 128
#ifndef NB_FWD_R64
#define NB_FWD_R64 3
#endif
#ifndef NB_FWD_MB64
#define NB_FWD_MB64 4
#endif
#ifndef NB_FWD_R128
#define NB_FWD_R128 2
#endif
#ifndef NB_FWD_MB128
#define NB_FWD_MB128 3
#endif
#ifndef NB_BWD_R64
#define NB_BWD_R64 2
#endif
#ifndef NB_BWD_MB64
#define NB_BWD_MB64 4
#endif
#ifndef NB_BWD_R128
#define NB_BWD_R128 1
#endif
#ifndef NB_BWD_MB128
#define NB_BWD_MB128 4
#endif

// The kernels are HBM-bound only if the per-sample math stays around 100 instructions, so the
// transcendental functions use the MUFU units (ex2/lg2.approx, <= 2^-21 relative error; the
// compositing outputs stay within 5e-6 of the reference's libm-based fp32 math).
__device__ __forceinline__ float fast_exp(float x) { return __expf(x); }
__device__ __forceinline__ float softplus_ref(float x) {
  // F.softplus(beta=1, threshold=20): utils/rendering.py:67
  // log1p(e) needs RELATIVE accuracy for tiny e: the last sample multiplies it by delta = 1e10 (:61),
  // so below 1e-2 use the series e - e^2/2 + e^3/3 instead of log(1 + e).
  const float e = __expf(x);
  const float series = e * fmaf(e, fmaf(e, 0.33333334f, -0.5f), 1.f);
  return x > 20.f ? x : (e < 1e-2f ? series : __logf(1.f + e));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  return v;
}

// |dirs| as the reference sees it (utils/rendering.py:62).  dirs_mode 0: dirs is [B,3] exactly as
// handed to volume_render.  dirs_mode 1: dirs is the rays tensor [B,6]; the direction is first
// normalised like render_nerf does (:37) and the norm of THAT (~1.0) is used.
__device__ __forceinline__ float dir_norm(const float* __restrict__ dirs, int64_t ray, int dirs_mode) {
  float dx, dy, dz;
  if (dirs_mode == 0) {
    dx = __ldg(dirs + ray * 3); dy = __ldg(dirs + ray * 3 + 1); dz = __ldg(dirs + ray * 3 + 2);
  } else {
    const float2* q = reinterpret_cast<const float2*>(dirs + ray * 6);
    const float2 b = __ldg(q + 1), c = __ldg(q + 2);
    const float ax = b.y, ay = c.x, az = c.y;
    const float inv = rsqrtf(fmaf(az, az, fmaf(ay, ay, ax * ax)));   // d/|d| to ~1 ulp (:37)
    dx = ax * inv; dy = ay * inv; dz = az * inv;
  }
  return sqrtf(fmaf(dz, dz, fmaf(dy, dy, dx * dx)));
}

struct SampleState {
  float4 o;     // r,g,b,sigma
  float t;      // sample depth
  float delta;  // (t[i+1]-t[i]) * |dir|, last = 1e10*|dir|
  float e;      // exp(-softplus(sigma)*delta)
  float T;      // exclusive transmittance
};

// Loads sample `idx` of `ray` and evaluates alpha; invalid lanes produce alpha=0, factor=1.
__device__ __forceinline__ void load_sample(const float* __restrict__ outs, const float* __restrict__ ts,
                                            int64_t ray, int N, int idx, float norm, SampleState& s,
                                            float& alpha, float& fac) {
  if (idx < N) {
    const int64_t g = ray * N + idx;
    s.o = __ldg(reinterpret_cast<const float4*>(outs) + g);
    s.t = __ldg(ts + g);
    const float d = (idx == N - 1) ? 1e10f : __fsub_rn(__ldg(ts + g + 1), s.t);  // :60-61
    s.delta = __fmul_rn(d, norm);                                               // :62
    s.e = fast_exp(__fmul_rn(-softplus_ref(s.o.w), s.delta));                        // :67
    alpha = __fsub_rn(1.f, s.e);
    fac = __fadd_rn(__fsub_rn(1.f, alpha), 1e-10f);                              // :68
  } else {
    s.o = make_float4(0.f, 0.f, 0.f, 0.f);
    s.t = 0.f; s.delta = 0.f; s.e = 1.f;
    alpha = 0.f; fac = 1.f;
  }
}

// Warp-wide exclusive product scan of `fac` with a running carry (updated).
__device__ __forceinline__ float excl_cumprod(float fac, float& carry, int lane) {
  float p = fac;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const float v = __shfl_up_sync(0xffffffffu, p, d);
    if (lane >= d) p *= v;
  }
  float ex = __shfl_up_sync(0xffffffffu, p, 1);
  if (lane == 0) ex = 1.f;
  const float T = carry * ex;
  carry *= __shfl_sync(0xffffffffu, p, 31);
  return T;
}

// Sum 16 per-lane values across the warp with 16 shuffles (a butterfly that halves the number of
// live values at every step); afterwards lane L holds the total of value L >> 1.
__device__ __forceinline__ float warp_multi_sum16(float (&v)[16], int lane) {
#pragma unroll
  for (int d = 16, n = 8; d >= 2; d >>= 1, n >>= 1) {
    const bool up = (lane & d) != 0;
#pragma unroll
    for (int i = 0; i < n; ++i) {
      const float keep = up ? v[i + n] : v[i];
      const float send = up ? v[i] : v[i + n];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, d);
    }
  }
  return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 1);
}

template <int S>
__device__ __forceinline__ void load_ts_vec(const float* __restrict__ p, float (&t)[S]) {
  if constexpr (S % 4 == 0) {
#pragma unroll
    for (int j = 0; j < S; j += 4) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(p + j));
      t[j] = v.x; t[j + 1] = v.y; t[j + 2] = v.z; t[j + 3] = v.w;
    }
  } else if constexpr (S % 2 == 0) {
#pragma unroll
    for (int j = 0; j < S; j += 2) {
      const float2 v = __ldg(reinterpret_cast<const float2*>(p + j));
      t[j] = v.x; t[j + 1] = v.y;
    }
  } else {
#pragma unroll
    for (int j = 0; j < S; ++j) t[j] = __ldg(p + j);
  }
}

template <int S>
__device__ __forceinline__ void store_vec(float* __restrict__ p, const float (&v)[S]) {
  if constexpr (S % 4 == 0) {
#pragma unroll
    for (int j = 0; j < S; j += 4) *reinterpret_cast<float4*>(p + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
  } else if constexpr (S % 2 == 0) {
#pragma unroll
    for (int j = 0; j < S; j += 2) *reinterpret_cast<float2*>(p + j) = make_float2(v[j], v[j + 1]);
  } else {
#pragma unroll
    for (int j = 0; j < S; ++j) p[j] = v[j];
  }
}

// |dirs| for the forward kernel: one lane per ray of the iteration, MUFU rsqrt/sqrt (<= 2 ulp).
template <int kDirsMode>
__device__ __forceinline__ float dir_norm_fast(const float* __restrict__ dirs, int64_t ray) {
  float dx, dy, dz;
  if constexpr (kDirsMode == 0) {
    dx = __ldg(dirs + ray * 3); dy = __ldg(dirs + ray * 3 + 1); dz = __ldg(dirs + ray * 3 + 2);
  } else {
    const float2* q = reinterpret_cast<const float2*>(dirs + ray * 6);
    const float2 b = __ldg(q + 1), c = __ldg(q + 2);
    return unit_dir_norm(b.y, c.x, c.y);
  }
  float n;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(n) : "f"(fmaf(dz, dz, fmaf(dy, dy, dx * dx))));              // :62
  return n;
}

// Forward.  One warp per ray; lane l owns the S CONSECUTIVE samples l*S .. l*S+S-1 (N <= 32*S), so
// the exclusive transmittance is S-1 serial multiplies plus ONE 5-step warp product scan per ray
// (instead of one scan per 32 samples) and t[i+1] is a register except at the lane boundary.
// R rays per warp iteration with every load issued before any math; the 5 per-ray sums of the R
// rays share one 16-value butterfly reduction; the R direction norms are computed by R lanes at
// once.  kFull (N == 32*S, 16-byte aligned tensors): compile-time strides, vector ts loads /
// alpha, weights stores and no bounds logic.
template <int S, int R, int MB, bool kFull, bool kAW, int kDirsMode>
__global__ void __launch_bounds__(kWarpsPerBlock * 32, MB)
composite_fwd_kernel(const float* __restrict__ outs, const float* __restrict__ ts,
                     const float* __restrict__ dirs, int64_t B, int Nrt, float* __restrict__ rgb,
                     float* __restrict__ disp, float* __restrict__ acc, float* __restrict__ alpha_out,
                     float* __restrict__ w_out) {
  static_assert(R * 5 <= 16, "per-iteration sums must fit the 16-value warp reduction");
  const int N = kFull ? 32 * S : Nrt;
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * kWarpsPerBlock;
  const int i0 = lane * S;
  for (int64_t ray0 = warp0 * R; ray0 < B; ray0 += nwarps * R) {
    float4 o[R][S];
    float t[R][S];
    const int64_t g0 = ray0 * N + i0;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const bool ok = ray0 + r < B;
      if constexpr (kFull) {
        if (ok) {
#pragma unroll
          for (int j = 0; j < S; ++j) o[r][j] = __ldg(reinterpret_cast<const float4*>(outs) + g0 + r * N + j);
          load_ts_vec<S>(ts + g0 + r * N, t[r]);
        } else {
#pragma unroll
          for (int j = 0; j < S; ++j) { o[r][j] = make_float4(0.f, 0.f, 0.f, 0.f); t[r][j] = 0.f; }
        }
      } else {
#pragma unroll
        for (int j = 0; j < S; ++j) {
          if (ok && i0 + j < N) {
            o[r][j] = __ldg(reinterpret_cast<const float4*>(outs) + g0 + r * N + j);
            t[r][j] = __ldg(ts + g0 + r * N + j);
          } else {
            o[r][j] = make_float4(0.f, 0.f, 0.f, 0.f); t[r][j] = 0.f;
          }
        }
      }
    }
    // lane r < R: |dir| of ray r of this iteration
    const float my_norm = dir_norm_fast<kDirsMode>(dirs, ray0 + lane < B && lane < R ? ray0 + lane : ray0);
    float sums[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) sums[i] = 0.f;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const float norm = __shfl_sync(0xffffffffu, my_norm, r);
      const float t_next_lane = __shfl_down_sync(0xffffffffu, t[r][0], 1);
      float a[S], pre[S];
      float run = 1.f;
#pragma unroll
      for (int j = 0; j < S; ++j) {
        const int idx = i0 + j;
        const bool valid = kFull || idx < N;
        const bool last = kFull ? (j == S - 1 && lane == 31) : (idx == N - 1);
        const float tn = (j == S - 1) ? t_next_lane : t[r][j + 1 < S ? j + 1 : j];
        const float d = last ? 1e10f : __fsub_rn(tn, t[r][j]);                        // :60-61
        const float delta = __fmul_rn(d, norm);                                       // :62
        const float e = transmit_factor(o[r][j].w, delta);                            // :67
        a[j] = valid ? __fsub_rn(1.f, e) : 0.f;
        const float fac = valid ? __fadd_rn(__fsub_rn(1.f, a[j]), 1e-10f) : 1.f;      // :68
        pre[j] = run;
        run *= fac;
      }
      // exclusive product scan of the per-lane products
      float p = run;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const float v = __shfl_up_sync(0xffffffffu, p, d);
        p *= (lane >= d) ? v : 1.f;
      }
      float ex = __shfl_up_sync(0xffffffffu, p, 1);
      ex = lane == 0 ? 1.f : ex;
      float sr = 0.f, sg = 0.f, sb = 0.f, sd = 0.f, sa = 0.f;
      float w[S];
#pragma unroll
      for (int j = 0; j < S; ++j) {
        w[j] = a[j] * (ex * pre[j]);
        sr = fmaf(w[j], o[r][j].x, sr); sg = fmaf(w[j], o[r][j].y, sg); sb = fmaf(w[j], o[r][j].z, sb);
        sd = fmaf(w[j], t[r][j], sd);
        sa += w[j];
      }
      if constexpr (kAW) {
        if (ray0 + r < B) {
          if constexpr (kFull) {
            if (alpha_out) store_vec<S>(alpha_out + g0 + r * N, a);
            if (w_out) store_vec<S>(w_out + g0 + r * N, w);
          } else {
#pragma unroll
            for (int j = 0; j < S; ++j)
              if (i0 + j < N) {
                if (alpha_out) alpha_out[g0 + r * N + j] = a[j];
                if (w_out) w_out[g0 + r * N + j] = w[j];
              }
          }
        }
      }
      sums[r * 5] = sr; sums[r * 5 + 1] = sg; sums[r * 5 + 2] = sb; sums[r * 5 + 3] = sd; sums[r * 5 + 4] = sa;
    }
    // lane L now owns value L>>1 = 5*r + k: k = 0..2 rgb, 3 depth (needs acc: lane of value 5r+4), 4 acc
    const float tot = warp_multi_sum16(sums, lane);
    const int vi = lane >> 1, r = vi / 5, k = vi - 5 * r;
    const float acc_r = __shfl_sync(0xffffffffu, tot, (5 * (r < R ? r : 0) + 4) * 2);
    if (!(lane & 1) && r < R && ray0 + r < B) {
      const int64_t ray = ray0 + r;
      if (k < 3) rgb[ray * 3 + k] = tot;
      else if (k == 3) disp[ray] = disparity(tot, acc_r);
      else acc[ray] = tot;
    }
  }
}

// Backward: same lane-owns-S-consecutive-samples layout.  alpha / transmittance are recomputed in
// registers (forward product scan), then sum_{i>j} g_i w_i is a per-lane serial suffix plus one warp
// suffix-sum scan (PyTorch's zero-free cumprod backward: reverse_cumsum(grad*out)/input).  kExtra:
// any of the d_disp / d_acc / d_alpha / d_w cotangents is present (never in train.py, which only
// differentiates rgb).
template <int S, int R, int MB, bool kFull, bool kExtra, int kDirsMode>
__global__ void __launch_bounds__(kWarpsPerBlock * 32, MB)
composite_bwd_kernel(const float* __restrict__ outs, const float* __restrict__ ts,
                     const float* __restrict__ dirs, const float* __restrict__ d_rgb,
                     const float* __restrict__ d_disp, const float* __restrict__ d_acc,
                     const float* __restrict__ d_alpha, const float* __restrict__ d_w, int64_t B, int Nrt,
                     float* __restrict__ d_outs) {
  const int N = kFull ? 32 * S : Nrt;
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * kWarpsPerBlock;
  const int i0 = lane * S;
  for (int64_t ray0 = warp0 * R; ray0 < B; ray0 += nwarps * R) {
    float4 o[R][S];
    float t[R][S];
    const int64_t g0 = ray0 * N + i0;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const bool ok = ray0 + r < B;
      if constexpr (kFull) {
        if (ok) {
#pragma unroll
          for (int j = 0; j < S; ++j) o[r][j] = __ldg(reinterpret_cast<const float4*>(outs) + g0 + r * N + j);
          load_ts_vec<S>(ts + g0 + r * N, t[r]);
        } else {
#pragma unroll
          for (int j = 0; j < S; ++j) { o[r][j] = make_float4(0.f, 0.f, 0.f, 0.f); t[r][j] = 0.f; }
        }
      } else {
#pragma unroll
        for (int j = 0; j < S; ++j) {
          if (ok && i0 + j < N) {
            o[r][j] = __ldg(reinterpret_cast<const float4*>(outs) + g0 + r * N + j);
            t[r][j] = __ldg(ts + g0 + r * N + j);
          } else {
            o[r][j] = make_float4(0.f, 0.f, 0.f, 0.f); t[r][j] = 0.f;
          }
        }
      }
    }
    // lane r < R: |dir| and the rgb cotangent of ray r of this iteration
    const int64_t my_ray = (lane < R && ray0 + lane < B) ? ray0 + lane : ray0;
    const float my_norm = dir_norm_fast<kDirsMode>(dirs, my_ray);
    const float my_gr = __ldg(d_rgb + my_ray * 3), my_gg = __ldg(d_rgb + my_ray * 3 + 1), my_gb = __ldg(d_rgb + my_ray * 3 + 2);
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int64_t ray = ray0 + r;
      const bool ok = ray < B;
      const float norm = __shfl_sync(0xffffffffu, my_norm, r);
      const float gr = __shfl_sync(0xffffffffu, my_gr, r), gg = __shfl_sync(0xffffffffu, my_gg, r),
                  gb = __shfl_sync(0xffffffffu, my_gb, r);
      const float t_next_lane = __shfl_down_sync(0xffffffffu, t[r][0], 1);
      float e[S], delta[S], fac[S], pre[S];
      float run = 1.f;
#pragma unroll
      for (int j = 0; j < S; ++j) {
        const int idx = i0 + j;
        const bool valid = kFull || idx < N;
        const bool last = kFull ? (j == S - 1 && lane == 31) : (idx == N - 1);
        const float tn = (j == S - 1) ? t_next_lane : t[r][j + 1 < S ? j + 1 : j];
        const float d = last ? 1e10f : __fsub_rn(tn, t[r][j]);                        // :60-61
        delta[j] = __fmul_rn(d, norm);                                                // :62
        e[j] = valid ? transmit_factor(o[r][j].w, delta[j]) : 1.f;                    // :67
        const float a = __fsub_rn(1.f, e[j]);
        fac[j] = valid ? __fadd_rn(__fsub_rn(1.f, a), 1e-10f) : 1.f;                  // :68
        pre[j] = run;
        run *= fac[j];
      }
      float p = run;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const float v = __shfl_up_sync(0xffffffffu, p, d);
        p *= (lane >= d) ? v : 1.f;
      }
      float ex = __shfl_up_sync(0xffffffffu, p, 1);
      ex = lane == 0 ? 1.f : ex;
      float T[S], w[S];
#pragma unroll
      for (int j = 0; j < S; ++j) { T[j] = ex * pre[j]; w[j] = (1.f - e[j]) * T[j]; }
      // disp = 1/max(1e-10, depth/acc)   (:82-83); only when those cotangents exist
      float g_depth = 0.f, g_acc = 0.f;
      if constexpr (kExtra) {
        if (d_acc) g_acc = __ldg(d_acc + (ok ? ray : ray0));
        if (d_disp) {
          float sd = 0.f, sa = 0.f;
#pragma unroll
          for (int j = 0; j < S; ++j) { sd = fmaf(w[j], t[r][j], sd); sa += w[j]; }
          const float depth = warp_sum(sd), acc = warp_sum(sa);
          const float q = depth / acc;
          const float m = fmaxf(1e-10f, q);
          const float g_q = (q > 1e-10f) ? -__ldg(d_disp + (ok ? ray : ray0)) / (m * m) : 0.f;
          g_depth = g_q / acc;
          g_acc -= g_q * depth / (acc * acc);
        }
      }
      float gw[S], suf[S];
      float tail = 0.f;       // sum of gw_i w_i over the later samples of this lane
#pragma unroll
      for (int j = S - 1; j >= 0; --j) {
        const bool valid = kFull || i0 + j < N;
        gw[j] = fmaf(gr, o[r][j].x, fmaf(gg, o[r][j].y, gb * o[r][j].z));
        if constexpr (kExtra) {
          gw[j] += g_depth * t[r][j] + g_acc;
          if (d_w) gw[j] += (valid && ok) ? __ldg(d_w + g0 + r * N + j) : 0.f;
        }
        suf[j] = tail;
        tail += valid ? gw[j] * w[j] : 0.f;
      }
      float q = tail;         // inclusive suffix sum over lanes >= lane
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const float v = __shfl_down_sync(0xffffffffu, q, d);
        q += (lane + d < 32) ? v : 0.f;
      }
      const float later = q - tail;     // lanes > lane
      float4 go[S];
#pragma unroll
      for (int j = 0; j < S; ++j) {
        const float Sj = later + suf[j];                         // sum_{i>j} gw_i w_i
        float g_a = gw[j] * T[j] - __fdividef(Sj, fac[j]);
        if constexpr (kExtra) {
          if (d_alpha) g_a += ((kFull || i0 + j < N) && ok) ? __ldg(d_alpha + g0 + r * N + j) : 0.f;
        }
        const float g_sp = (g_a * e[j]) * delta[j];              // this order keeps 0*1e10 == 0
        const float z = exp2f_approx(fminf(o[r][j].w, 20.f) * 1.4426950408889634f);
        const float g_sigma = (o[r][j].w > 20.f) ? g_sp : g_sp * __fdividef(z, z + 1.f);
        go[j] = make_float4(w[j] * gr, w[j] * gg, w[j] * gb, g_sigma);
      }
      if (ok) {
#pragma unroll
        for (int j = 0; j < S; ++j)
          if (kFull || i0 + j < N) reinterpret_cast<float4*>(d_outs)[g0 + r * N + j] = go[j];
      }
    }
  }
}

// ---- thread-per-ray fallback for N > 256 (correct, not tuned) -------------------------------
__global__ void composite_fwd_serial_kernel(const float* __restrict__ outs, const float* __restrict__ ts,
                                            const float* __restrict__ dirs, int dirs_mode, int64_t B, int N,
                                            float* __restrict__ rgb, float* __restrict__ disp,
                                            float* __restrict__ acc, float* __restrict__ alpha_out,
                                            float* __restrict__ w_out) {
  const int64_t ray = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (ray >= B) return;
  const float norm = dir_norm(dirs, ray, dirs_mode);
  float T = 1.f, sr = 0.f, sg = 0.f, sb = 0.f, sd = 0.f, sa = 0.f;
  for (int i = 0; i < N; ++i) {
    SampleState s; float a, fac;
    load_sample(outs, ts, ray, N, i, norm, s, a, fac);
    const float w = a * T;
    sr = fmaf(w, s.o.x, sr); sg = fmaf(w, s.o.y, sg); sb = fmaf(w, s.o.z, sb);
    sd = fmaf(w, s.t, sd); sa += w;
    if (alpha_out) alpha_out[ray * N + i] = a;
    if (w_out) w_out[ray * N + i] = w;
    T *= fac;
  }
  rgb[ray * 3] = sr; rgb[ray * 3 + 1] = sg; rgb[ray * 3 + 2] = sb;
  disp[ray] = disparity(sd, sa);
  acc[ray] = sa;
}

__global__ void composite_bwd_serial_kernel(const float* __restrict__ outs, const float* __restrict__ ts,
                                            const float* __restrict__ dirs, int dirs_mode, const float* __restrict__ d_rgb,
                                            const float* __restrict__ d_disp, const float* __restrict__ d_acc,
                                            const float* __restrict__ d_alpha, const float* __restrict__ d_w,
                                            int64_t B, int N, float* __restrict__ d_outs) {
  const int64_t ray = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (ray >= B) return;
  const float norm = dir_norm(dirs, ray, dirs_mode);
  const float gr = d_rgb[ray * 3], gg = d_rgb[ray * 3 + 1], gb = d_rgb[ray * 3 + 2];
  // pass 1: depth, acc
  float T = 1.f, depth = 0.f, acc = 0.f;
  for (int i = 0; i < N; ++i) {
    SampleState s; float a, fac;
    load_sample(outs, ts, ray, N, i, norm, s, a, fac);
    depth = fmaf(a * T, s.t, depth); acc += a * T; T *= fac;
  }
  float g_depth = 0.f, g_acc = d_acc ? d_acc[ray] : 0.f;
  if (d_disp) {
    const float q = depth / acc, m = fmaxf(1e-10f, q);
    const float g_q = (q > 1e-10f) ? -d_disp[ray] / (m * m) : 0.f;
    g_depth = g_q / acc;
    g_acc -= g_q * depth / (acc * acc);
  }
  // pass 2: total of gw_i*w_i
  float total = 0.f; T = 1.f;
  for (int i = 0; i < N; ++i) {
    SampleState s; float a, fac;
    load_sample(outs, ts, ray, N, i, norm, s, a, fac);
    float gw = fmaf(gr, s.o.x, fmaf(gg, s.o.y, gb * s.o.z)) + g_depth * s.t + g_acc;
    if (d_w) gw += d_w[ray * N + i];
    total += gw * a * T; T *= fac;
  }
  // pass 3: gradients, S_j = total - prefix_inclusive_j
  float prefix = 0.f; T = 1.f;
  for (int i = 0; i < N; ++i) {
    SampleState s; float a, fac;
    load_sample(outs, ts, ray, N, i, norm, s, a, fac);
    const float w = a * T;
    float gw = fmaf(gr, s.o.x, fmaf(gg, s.o.y, gb * s.o.z)) + g_depth * s.t + g_acc;
    if (d_w) gw += d_w[ray * N + i];
    prefix += gw * w;
    float g_a = gw * T - (total - prefix) / fac;
    if (d_alpha) g_a += d_alpha[ray * N + i];
    const float g_sp = (g_a * s.e) * s.delta;
    float g_sigma = g_sp;
    if (!(s.o.w > 20.f)) { const float z = fast_exp(s.o.w); g_sigma = g_sp * z / (z + 1.f); }
    reinterpret_cast<float4*>(d_outs)[ray * N + i] = make_float4(w * gr, w * gg, w * gb, g_sigma);
    T *= fac;
  }
}

static int warp_grid(int64_t B) {
  const int64_t blocks = ceil_div64(B, kWarpsPerBlock);
  const int64_t cap = (int64_t)sm_count() * 32;  // 8 resident blocks/SM x 4 waves, grid-stride beyond
  return (int)(blocks < cap ? blocks : cap);
}

}  // namespace nb200

extern "C" {

int nb200_composite_forward(const float* outs, const float* ts, const float* dirs, int dirs_mode, int64_t B, int N,
                            float* rgb, float* disp, float* acc, float* alpha, float* weights,
                            nb200_stream_t stream) {
  using namespace nb200;
  if (B < 0 || N < 2 || (dirs_mode != 0 && dirs_mode != 1)) return NB200_ERR_ARG;
  if (B == 0) return NB200_OK;  // empty batch: pointers may be null
  if (!outs || !ts || !dirs || !rgb || !disp || !acc) return NB200_ERR_ARG;
  if (((uintptr_t)outs & 15) || (dirs_mode == 1 && ((uintptr_t)dirs & 7))) return NB200_ERR_ARG;  // float4 / float2 rows
  cudaStream_t s = as_stream(stream);
  const bool want_aw = alpha != nullptr || weights != nullptr;
  const bool aligned = (((uintptr_t)outs | (uintptr_t)ts | (uintptr_t)alpha | (uintptr_t)weights) & 15) == 0;
#define NB_FWD_L(S, R, MB, FULL, AW, DM)                                                                     \
  composite_fwd_kernel<S, R, MB, FULL, AW, DM><<<warp_grid((B + R - 1) / R), kWarpsPerBlock * 32, 0, s>>>(   \
      outs, ts, dirs, B, N, rgb, disp, acc, alpha, weights)
#define NB_FWD_D(S, R, MB, FULL, AW)                                                                         \
  do {                                                                                                       \
    if (dirs_mode) NB_FWD_L(S, R, MB, FULL, AW, 1); else NB_FWD_L(S, R, MB, FULL, AW, 0);                    \
  } while (0)
#define NB_FWD(S, R, MB)                                                                                     \
  do {                                                                                                       \
    const bool full = (N == S * 32) && aligned;                                                              \
    if (full && want_aw) NB_FWD_D(S, R, MB, true, true);                                                     \
    else if (full) NB_FWD_D(S, R, MB, true, false);                                                          \
    else if (want_aw) NB_FWD_D(S, R, MB, false, true);                                                       \
    else NB_FWD_D(S, R, MB, false, false);                                                                   \
  } while (0)
  if (N <= 32) NB_FWD(1, 3, 4);
  else if (N <= 64) NB_FWD(2, NB_FWD_R64, NB_FWD_MB64);
  else if (N <= 96) NB_FWD(3, 2, 4);
  else if (N <= 128) NB_FWD(4, NB_FWD_R128, NB_FWD_MB128);
  else if (N <= 192) NB_FWD(6, 1, 3);
  else if (N <= 256) NB_FWD(8, 1, 3);
  else {
    composite_fwd_serial_kernel<<<(unsigned)ceil_div64(B, 128), 128, 0, as_stream(stream)>>>(
        outs, ts, dirs, dirs_mode, B, N, rgb, disp, acc, alpha, weights);
  }
#undef NB_FWD
#undef NB_FWD_D
#undef NB_FWD_L
  NB_LAUNCH_CHECK("composite_fwd_kernel");
  return NB200_OK;
}

int nb200_composite_backward(const float* outs, const float* ts, const float* dirs, int dirs_mode,
                             const float* d_rgb, const float* d_disp, const float* d_acc,
                             const float* d_alpha, const float* d_w, int64_t B, int N,
                             float* d_outs, nb200_stream_t stream) {
  using namespace nb200;
  if (B < 0 || N < 2 || (dirs_mode != 0 && dirs_mode != 1)) return NB200_ERR_ARG;
  if (B == 0) return NB200_OK;
  if (!outs || !ts || !dirs || !d_rgb || !d_outs) return NB200_ERR_ARG;
  if ((((uintptr_t)outs | (uintptr_t)d_outs) & 15) || (dirs_mode == 1 && ((uintptr_t)dirs & 7))) return NB200_ERR_ARG;
  cudaStream_t s = as_stream(stream);
  const bool extra = d_disp || d_acc || d_alpha || d_w;
  const bool aligned = (((uintptr_t)outs | (uintptr_t)ts | (uintptr_t)d_outs) & 15) == 0;
#define NB_BWD_L(S, R, MB, FULL, EX, DM)                                                                     \
  composite_bwd_kernel<S, R, MB, FULL, EX, DM><<<warp_grid((B + R - 1) / R), kWarpsPerBlock * 32, 0, s>>>(   \
      outs, ts, dirs, d_rgb, d_disp, d_acc, d_alpha, d_w, B, N, d_outs)
#define NB_BWD_D(S, R, MB, FULL, EX)                                                                         \
  do {                                                                                                       \
    if (dirs_mode) NB_BWD_L(S, R, MB, FULL, EX, 1); else NB_BWD_L(S, R, MB, FULL, EX, 0);                    \
  } while (0)
#define NB_BWD(S, R, MB)                                                                                     \
  do {                                                                                                       \
    const bool full = (N == S * 32) && aligned;                                                              \
    if (full && !extra) NB_BWD_D(S, R, MB, true, false);                                                     \
    else if (full) NB_BWD_D(S, R, MB, true, true);                                                           \
    else if (!extra) NB_BWD_D(S, R, MB, false, false);                                                       \
    else NB_BWD_D(S, R, MB, false, true);                                                                    \
  } while (0)
  if (N <= 32) NB_BWD(1, 2, 4);
  else if (N <= 64) NB_BWD(2, NB_BWD_R64, NB_BWD_MB64);
  else if (N <= 96) NB_BWD(3, 1, 4);
  else if (N <= 128) NB_BWD(4, NB_BWD_R128, NB_BWD_MB128);
  else if (N <= 192) NB_BWD(6, 1, 2);
  else if (N <= 256) NB_BWD(8, 1, 2);
  else
    composite_bwd_serial_kernel<<<(unsigned)ceil_div64(B, 128), 128, 0, s>>>(
        outs, ts, dirs, dirs_mode, d_rgb, d_disp, d_acc, d_alpha, d_w, B, N, d_outs);
#undef NB_BWD
#undef NB_BWD_D
#undef NB_BWD_L
  NB_LAUNCH_CHECK("composite_bwd_kernel");
  return NB200_OK;
}

}  // extern "C"
