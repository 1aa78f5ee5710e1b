// bias_act.cu -- bias + (noise) + leaky-ReLU + gain family, forward and backward (sm_100a).
//
// Replaces reference models/stylegan2/op/fused_bias_act_kernel.cu:18-99 (one scalar element per
// thread-iteration, 128-thread blocks, int div/mod per element) with 128-bit streaming accesses,
// four independent vectors in flight per thread and one channel lookup per vector.  HBM-bound:
// algorithmic bytes = s*(2*numel) forward, s*(3*numel) backward (s = bytes/element).
#include "common.cuh"

namespace gg {
namespace {

constexpr int kThreads = 256;
constexpr int kUnroll = 4;

__device__ __forceinline__ float act_apply(float x, float ref, int act, int grad, float alpha) {
  // table of fused_bias_act_kernel.cu:28-47 (act*10+grad)
  if (grad == 2) return 0.f;
  if (act == 3) {
    const float gate = (grad == 0) ? x : ref;
    return gate > 0.f ? x : x * alpha;
  }
  return x;
}

// Flat kernel, exact reference semantics: bias index = (i / step_b) % size_b.
// VEC elements per access (16 bytes when VEC = 16/sizeof(T), or VEC = 1 scalar fallback).
template <typename T, int VEC, typename Index>
__global__ void __launch_bounds__(kThreads)
bias_act_flat_kernel(T* __restrict__ out, const T* __restrict__ x, const T* __restrict__ bias,
                     const T* __restrict__ ref, int act, int grad, float alpha, float scale,
                     Index n_vec, Index step_b, Index size_b) {
  const Index base = (static_cast<Index>(blockIdx.x) * kUnroll) * kThreads + threadIdx.x;
  if constexpr (VEC > 1) {
    Vec16<T> xv[kUnroll], rv[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const Index v = base + static_cast<Index>(u) * kThreads;
      if (v < n_vec) {
        xv[u] = ld_vec_stream(x + v * VEC);
        if (ref) rv[u] = ld_vec_stream(ref + v * VEC);
      }
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const Index v = base + static_cast<Index>(u) * kThreads;
      if (v < n_vec) {
        float b = 0.f;
        if (bias) b = Cvt<T>::to_f(bias[((v * VEC) / step_b) % size_b]);
        Vec16<T> o;
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
          const float r = ref ? Cvt<T>::to_f(rv[u].v[k]) : 0.f;
          o.v[k] = Cvt<T>::from_f(act_apply(Cvt<T>::to_f(xv[u].v[k]) + b, r, act, grad, alpha) * scale);
        }
        st_vec_stream(out + v * VEC, o);
      }
    }
  } else {
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const Index i = base + static_cast<Index>(u) * kThreads;
      if (i < n_vec) {
        float b = 0.f;
        if (bias) b = Cvt<T>::to_f(bias[(i / step_b) % size_b]);
        const float r = ref ? Cvt<T>::to_f(ref[i]) : 0.f;
        out[i] = Cvt<T>::from_f(act_apply(Cvt<T>::to_f(x[i]) + b, r, act, grad, alpha) * scale);
      }
    }
  }
}

// (N, C, HW) kernel with per-sample noise plane: out = lrelu(x + nw*noise[n,p] + bias[c]) * scale.
template <typename T, int VEC>
__global__ void __launch_bounds__(kThreads)
noise_bias_act_kernel(T* __restrict__ out, const T* __restrict__ x, const T* __restrict__ noise,
                      const float* __restrict__ noise_weight, const float* __restrict__ bias,
                      const float* __restrict__ row_scale, float alpha, float scale, int64_t n_vec, int64_t C,
                      int64_t HW) {
  const float nw = noise ? (noise_weight ? __ldg(noise_weight) : 1.f) : 0.f;
  const int64_t base = (static_cast<int64_t>(blockIdx.x) * kUnroll) * kThreads + threadIdx.x;
  Vec16<T> xv[kUnroll], nv[kUnroll];
  float bv[kUnroll], rv[kUnroll];
#pragma unroll
  for (int u = 0; u < kUnroll; ++u) {
    const int64_t v = base + static_cast<int64_t>(u) * kThreads;
    if (v < n_vec) {
      const int64_t i = v * VEC;
      xv[u] = ld_vec_stream(x + i);
      const int64_t row = i / HW;  // n*C + c
      const int64_t p = i - row * HW;
      const int64_t n = row / C;
      const int64_t c = row - n * C;
      bv[u] = bias ? __ldg(bias + c) : 0.f;
      rv[u] = row_scale ? __ldg(row_scale + row) : 1.f;
      if (noise) nv[u] = *reinterpret_cast<const Vec16<T>*>(noise + n * HW + p);  // re-read per channel: keep in L1/L2
    }
  }
#pragma unroll
  for (int u = 0; u < kUnroll; ++u) {
    const int64_t v = base + static_cast<int64_t>(u) * kThreads;
    if (v < n_vec) {
      Vec16<T> o;
#pragma unroll
      for (int k = 0; k < VEC; ++k) {
        float t = fmaf(Cvt<T>::to_f(xv[u].v[k]), rv[u], bv[u]);
        if (noise) t = fmaf(nw, Cvt<T>::to_f(nv[u].v[k]), t);
        o.v[k] = Cvt<T>::from_f((t > 0.f ? t : t * alpha) * scale);
      }
      st_vec_stream(out + v * VEC, o);
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
noise_bias_act_scalar_kernel(T* __restrict__ out, const T* __restrict__ x, const T* __restrict__ noise,
                             const float* __restrict__ noise_weight, const float* __restrict__ bias,
                             const float* __restrict__ row_scale, float alpha, float scale, int64_t numel, int64_t C,
                             int64_t HW) {
  const float nw = noise ? (noise_weight ? __ldg(noise_weight) : 1.f) : 0.f;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x; i < numel;
       i += static_cast<int64_t>(gridDim.x) * kThreads) {
    const int64_t row = i / HW, p = i - row * HW, n = row / C, c = row - n * C;
    float t = Cvt<T>::to_f(x[i]) * (row_scale ? __ldg(row_scale + row) : 1.f) + (bias ? __ldg(bias + c) : 0.f);
    if (noise) t = fmaf(nw, Cvt<T>::to_f(noise[n * HW + p]), t);
    out[i] = Cvt<T>::from_f((t > 0.f ? t : t * alpha) * scale);
  }
}

// Backward: gx = (out > 0 ? g : alpha*g) * scale, plus per-(row, chunk) partial sums of gx.
// One CTA owns `chunk` consecutive elements of one (n,c) row (chunk == HW when HW is small and a CTA
// then owns kRowsSmall... see launch code).  Partials are reduced by bias_grad_finish_kernel in a
// fixed order -> bit-reproducible grad_bias (the reference's grad_input.sum() is not).
template <typename T, int VEC>
__global__ void __launch_bounds__(kThreads)
bias_act_bwd_kernel(T* __restrict__ gx, float* __restrict__ partial, const T* __restrict__ g,
                    const T* __restrict__ out, float alpha, float scale, int64_t HW, int64_t chunk,
                    int chunks_per_row) {
  // blockIdx.x = row * chunks_per_row + chunk_id
  const int64_t row = blockIdx.x / chunks_per_row;
  const int ck = blockIdx.x - row * chunks_per_row;
  const int64_t p0 = static_cast<int64_t>(ck) * chunk;
  const int64_t p1 = min(p0 + chunk, HW);
  const int64_t off = row * HW;
  float acc = 0.f;
  if constexpr (VEC > 1) {
    const int64_t nv = (p1 - p0) / VEC;  // chunk and HW are multiples of VEC on this path
    for (int64_t v0 = 0; v0 < nv; v0 += static_cast<int64_t>(kThreads) * kUnroll) {
      Vec16<T> gv[kUnroll], ov[kUnroll];
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const int64_t v = v0 + u * kThreads + threadIdx.x;
        if (v < nv) {
          gv[u] = ld_vec_stream(g + off + p0 + v * VEC);
          ov[u] = ld_vec_stream(out + off + p0 + v * VEC);
        }
      }
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const int64_t v = v0 + u * kThreads + threadIdx.x;
        if (v < nv) {
          Vec16<T> r;
#pragma unroll
          for (int k = 0; k < VEC; ++k) {
            const float gg_ = Cvt<T>::to_f(gv[u].v[k]);
            const float y = (Cvt<T>::to_f(ov[u].v[k]) > 0.f ? gg_ : gg_ * alpha) * scale;
            r.v[k] = Cvt<T>::from_f(y);
            acc += Cvt<T>::to_f(r.v[k]);  // sum what was stored (reference sums the stored grad_input)
          }
          st_vec_stream(gx + off + p0 + v * VEC, r);
        }
      }
    }
  } else {
    for (int64_t p = p0 + threadIdx.x; p < p1; p += kThreads) {
      const float gg_ = Cvt<T>::to_f(g[off + p]);
      const T y = Cvt<T>::from_f((Cvt<T>::to_f(out[off + p]) > 0.f ? gg_ : gg_ * alpha) * scale);
      gx[off + p] = y;
      acc += Cvt<T>::to_f(y);
    }
  }
  if (partial) {
    __shared__ float wsum[kThreads / 32];
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < kThreads / 32; ++w) s += wsum[w];
      partial[blockIdx.x] = s;
    }
  }
}

// Small-row variant: one warp per (n,c) row (HW < 1024), 8 rows per CTA.
template <typename T>
__global__ void __launch_bounds__(kThreads)
bias_act_bwd_rows_kernel(T* __restrict__ gx, float* __restrict__ partial, const T* __restrict__ g,
                         const T* __restrict__ out, float alpha, float scale, int64_t rows, int64_t HW) {
  const int64_t row = static_cast<int64_t>(blockIdx.x) * (kThreads / 32) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const int64_t off = row * HW;
  float acc = 0.f;
  for (int64_t p = lane; p < HW; p += 32) {
    const float gg_ = Cvt<T>::to_f(g[off + p]);
    const T y = Cvt<T>::from_f((Cvt<T>::to_f(out[off + p]) > 0.f ? gg_ : gg_ * alpha) * scale);
    gx[off + p] = y;
    acc += Cvt<T>::to_f(y);
  }
  if (partial) {
    acc = warp_sum(acc);
    if (lane == 0) partial[row] = acc;
  }
}

// grad_bias[c] = sum_n sum_k partial[(n*C + c)*K + k]   (fixed order; one warp per channel)
__global__ void bias_grad_finish_kernel(float* __restrict__ grad_bias, const float* __restrict__ partial,
                                        int64_t N, int64_t C, int K) {
  const int64_t c = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (c >= C) return;
  const int lane = threadIdx.x & 31;
  float acc = 0.f;
  const int64_t total = N * K;
  for (int64_t j = lane; j < total; j += 32) {
    const int64_t n = j / K;
    const int k = static_cast<int>(j - n * K);
    acc += partial[(n * C + c) * K + k];
  }
  acc = warp_sum(acc);
  if (lane == 0) grad_bias[c] = acc;
}

// out[n,c,p] = x[n,c,p] * s[n*C + c]   (modulating the INPUT of a weight-shared convolution, networks.py:236/253
// rewritten as conv(W, x*s)); with `y` given, also row_dot[row] = sum_p x[row,p]*y[row,p] (the gradient w.r.t. s).
template <typename T, int VEC>
__global__ void __launch_bounds__(kThreads)
channel_scale_kernel(T* __restrict__ out, float* __restrict__ partial, const T* __restrict__ x, const T* __restrict__ y,
                     const float* __restrict__ s, int64_t HW, int64_t chunk, int chunks_per_row) {
  const int64_t row = blockIdx.x / chunks_per_row;
  const int ck = blockIdx.x - row * chunks_per_row;
  const int64_t p0 = static_cast<int64_t>(ck) * chunk;
  const int64_t p1 = min(p0 + chunk, HW);
  const int64_t off = row * HW;
  const float sv = __ldg(s + row);
  float acc = 0.f;
  if constexpr (VEC > 1) {
    const int64_t nv = (p1 - p0) / VEC;
    for (int64_t v0 = 0; v0 < nv; v0 += static_cast<int64_t>(kThreads) * kUnroll) {
      Vec16<T> xv[kUnroll], yv[kUnroll];
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const int64_t v = v0 + u * kThreads + threadIdx.x;
        if (v < nv) {
          xv[u] = ld_vec_stream(x + off + p0 + v * VEC);
          if (y) yv[u] = ld_vec_stream(y + off + p0 + v * VEC);
        }
      }
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const int64_t v = v0 + u * kThreads + threadIdx.x;
        if (v < nv) {
          Vec16<T> r;
#pragma unroll
          for (int k = 0; k < VEC; ++k) {
            const float xf = Cvt<T>::to_f(xv[u].v[k]);
            r.v[k] = Cvt<T>::from_f(xf * sv);
            if (y) acc = fmaf(xf, Cvt<T>::to_f(yv[u].v[k]), acc);
          }
          st_vec_stream(out + off + p0 + v * VEC, r);
        }
      }
    }
  } else {
    for (int64_t p = p0 + threadIdx.x; p < p1; p += kThreads) {
      const float xf = Cvt<T>::to_f(x[off + p]);
      out[off + p] = Cvt<T>::from_f(xf * sv);
      if (y) acc = fmaf(xf, Cvt<T>::to_f(y[off + p]), acc);
    }
  }
  if (partial) {
    __shared__ float wsum[kThreads / 32];
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < kThreads / 32; ++w) t += wsum[w];
      partial[blockIdx.x] = t;
    }
  }
}

// small rows (HW < 1024): one warp per row, 8 rows per CTA
template <typename T>
__global__ void __launch_bounds__(kThreads)
channel_scale_rows_kernel(T* __restrict__ out, float* __restrict__ row_dot, const T* __restrict__ x,
                          const T* __restrict__ y, const float* __restrict__ s, int64_t rows, int64_t HW) {
  const int64_t row = static_cast<int64_t>(blockIdx.x) * (kThreads / 32) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const int64_t off = row * HW;
  const float sv = __ldg(s + row);
  float acc = 0.f;
  for (int64_t p = lane; p < HW; p += 32) {
    const float xf = Cvt<T>::to_f(x[off + p]);
    out[off + p] = Cvt<T>::from_f(xf * sv);
    if (y) acc = fmaf(xf, Cvt<T>::to_f(y[off + p]), acc);
  }
  if (row_dot) {
    acc = warp_sum(acc);
    if (lane == 0) row_dot[row] = acc;
  }
}

// row_dot[row] = sum_k partial[row*K + k]
__global__ void row_finish_kernel(float* __restrict__ row_dot, const float* __restrict__ partial, int64_t rows, int K) {
  const int64_t r = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  float acc = 0.f;
  for (int k = 0; k < K; ++k) acc += partial[r * K + k];
  row_dot[r] = acc;
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

template <typename T>
int launch_flat(void* out, const void* x, const void* bias, const void* ref, int act, int grad,
                float alpha, float scale, int64_t size_x, int64_t step_b, int64_t size_b,
                cudaStream_t st) {
  constexpr int V = 16 / sizeof(T);
  const bool vec = (step_b % V == 0 || bias == nullptr) && (size_x % V == 0) && aligned16(out) &&
                   aligned16(x) && (ref == nullptr || aligned16(ref));
  const int64_t n_vec = vec ? size_x / V : size_x;
  const int64_t per_cta = static_cast<int64_t>(kThreads) * kUnroll;
  const int64_t grid = (n_vec + per_cta - 1) / per_cta;
  if (grid > 0x7fffffffLL) return fail(GG_ERR_BAD_ARG, "fused_bias_act: tensor too large");
  const bool small = size_x < (1LL << 31);
  auto* o = static_cast<T*>(out);
  auto* xi = static_cast<const T*>(x);
  auto* b = static_cast<const T*>(bias);
  auto* r = static_cast<const T*>(ref);
  if (size_b == 0) size_b = 1;
  if (step_b == 0) step_b = 1;
#define GG_LAUNCH(VEC_, IDX_)                                                                      \
  bias_act_flat_kernel<T, VEC_, IDX_><<<static_cast<unsigned>(grid), kThreads, 0, st>>>(           \
      o, xi, b, r, act, grad, alpha, scale, static_cast<IDX_>(n_vec), static_cast<IDX_>(step_b),   \
      static_cast<IDX_>(size_b))
  if (vec) {
    if (small) GG_LAUNCH(V, uint32_t); else GG_LAUNCH(V, int64_t);
  } else {
    if (small) GG_LAUNCH(1, uint32_t); else GG_LAUNCH(1, int64_t);
  }
#undef GG_LAUNCH
  GG_CHECK_LAUNCH("fused_bias_act launch");
  return GG_OK;
}

template <typename T>
int launch_noise(void* out, const void* x, const void* noise, const float* nw, const float* bias,
                 const float* row_scale, float alpha, float scale, int64_t N, int64_t C, int64_t HW, cudaStream_t st) {
  constexpr int V = 16 / sizeof(T);
  const int64_t numel = N * C * HW;
  const bool vec = (HW % V == 0) && aligned16(out) && aligned16(x) && (!noise || aligned16(noise));
  if (vec) {
    const int64_t n_vec = numel / V;
    const int64_t per_cta = static_cast<int64_t>(kThreads) * kUnroll;
    const int64_t grid = (n_vec + per_cta - 1) / per_cta;
    if (grid > 0x7fffffffLL) return fail(GG_ERR_BAD_ARG, "noise_bias_act: tensor too large");
    noise_bias_act_kernel<T, V><<<static_cast<unsigned>(grid), kThreads, 0, st>>>(
        static_cast<T*>(out), static_cast<const T*>(x), static_cast<const T*>(noise), nw, bias, row_scale, alpha,
        scale, n_vec, C, HW);
  } else {
    int64_t grid = (numel + kThreads - 1) / kThreads;
    if (grid > 148 * 16) grid = 148 * 16;
    noise_bias_act_scalar_kernel<T><<<static_cast<unsigned>(grid), kThreads, 0, st>>>(
        static_cast<T*>(out), static_cast<const T*>(x), static_cast<const T*>(noise), nw, bias, row_scale, alpha,
        scale, numel, C, HW);
  }
  GG_CHECK_LAUNCH("noise_bias_act launch");
  return GG_OK;
}

// geometry of the backward reduction, shared by the workspace query and the launch
struct BwdGeom {
  bool small_rows;      // one warp per row
  int64_t chunk;        // elements per CTA within a row
  int chunks_per_row;   // K
};
inline BwdGeom bwd_geom(int64_t HW) {
  BwdGeom g;
  g.small_rows = HW < 1024;
  if (g.small_rows) {
    g.chunk = HW;
    g.chunks_per_row = 1;
  } else {
    g.chunk = 8192;  // 32 KB of fp32 per stream per CTA
    g.chunks_per_row = static_cast<int>((HW + g.chunk - 1) / g.chunk);
  }
  return g;
}

template <typename T>
int launch_bwd(void* gx, float* grad_bias, void* workspace, const void* g, const void* out, float alpha,
               float scale, int64_t N, int64_t C, int64_t HW, cudaStream_t st) {
  constexpr int V = 16 / sizeof(T);
  const int64_t rows = N * C;
  const BwdGeom geo = bwd_geom(HW);
  float* partial = grad_bias ? static_cast<float*>(workspace) : nullptr;
  if (geo.small_rows) {
    const int64_t grid = (rows + (kThreads / 32) - 1) / (kThreads / 32);
    if (grid > 0x7fffffffLL) return fail(GG_ERR_BAD_ARG, "bias_act_backward: too many rows");
    bias_act_bwd_rows_kernel<T><<<static_cast<unsigned>(grid), kThreads, 0, st>>>(
        static_cast<T*>(gx), partial, static_cast<const T*>(g), static_cast<const T*>(out), alpha, scale,
        rows, HW);
  } else {
    const int64_t grid = rows * geo.chunks_per_row;
    if (grid > 0x7fffffffLL) return fail(GG_ERR_BAD_ARG, "bias_act_backward: too many rows");
    const bool vec = (HW % V == 0) && (geo.chunk % V == 0) && aligned16(gx) && aligned16(g) && aligned16(out);
    if (vec)
      bias_act_bwd_kernel<T, V><<<static_cast<unsigned>(grid), kThreads, 0, st>>>(
          static_cast<T*>(gx), partial, static_cast<const T*>(g), static_cast<const T*>(out), alpha,
          scale, HW, geo.chunk, geo.chunks_per_row);
    else
      bias_act_bwd_kernel<T, 1><<<static_cast<unsigned>(grid), kThreads, 0, st>>>(
          static_cast<T*>(gx), partial, static_cast<const T*>(g), static_cast<const T*>(out), alpha,
          scale, HW, geo.chunk, geo.chunks_per_row);
  }
  GG_CHECK_LAUNCH("bias_act_backward launch");
  if (grad_bias) {
    const int warps = 4;
    const int64_t grid = (C + warps - 1) / warps;
    bias_grad_finish_kernel<<<static_cast<unsigned>(grid), warps * 32, 0, st>>>(grad_bias, partial, N, C,
                                                                                geo.chunks_per_row);
    GG_CHECK_LAUNCH("bias_grad_finish launch");
  }
  return GG_OK;
}

}  // namespace
}  // namespace gg

using namespace gg;

extern "C" {

int gg_fused_bias_act(void* out, const void* x, const void* bias, const void* ref, int dtype, int act,
                      int grad, float alpha, float scale, int64_t size_x, int64_t step_b,
                      int64_t size_b, void* stream) {
  if (size_x < 0 || step_b < 0 || size_b < 0) return fail(GG_ERR_BAD_ARG, "fused_bias_act: negative size");
  if (size_x == 0) return GG_OK;
  if (!out || !x) return fail(GG_ERR_BAD_ARG, "fused_bias_act: null tensor");
  if (act != 1 && act != 3) return fail(GG_ERR_UNSUPPORTED, "fused_bias_act: act must be 1 (linear) or 3 (lrelu)");
  if (grad < 0 || grad > 2) return fail(GG_ERR_BAD_ARG, "fused_bias_act: grad must be 0, 1 or 2");
  if (bias && (size_b <= 0 || step_b <= 0)) return fail(GG_ERR_BAD_ARG, "fused_bias_act: bias given with size_b/step_b <= 0");
  auto st = static_cast<cudaStream_t>(stream);
  switch (dtype) {
    case GG_F32: return launch_flat<float>(out, x, bias, ref, act, grad, alpha, scale, size_x, step_b, size_b, st);
    case GG_F16: return launch_flat<__half>(out, x, bias, ref, act, grad, alpha, scale, size_x, step_b, size_b, st);
    case GG_BF16: return launch_flat<__nv_bfloat16>(out, x, bias, ref, act, grad, alpha, scale, size_x, step_b, size_b, st);
    default: return fail(GG_ERR_UNSUPPORTED, "fused_bias_act: dtype %d not supported (f32/f16/bf16)", dtype);
  }
}

int gg_noise_bias_act(void* out, const void* x, const void* noise, const float* noise_weight,
                      const float* bias, const float* row_scale, int dtype, float alpha, float scale, int64_t N,
                      int64_t C, int64_t HW, void* stream) {
  if (N < 0 || C < 0 || HW < 0) return fail(GG_ERR_BAD_ARG, "noise_bias_act: negative size");
  if (N * C * HW == 0) return GG_OK;
  if (!out || !x) return fail(GG_ERR_BAD_ARG, "noise_bias_act: null tensor");
  auto st = static_cast<cudaStream_t>(stream);
  switch (dtype) {
    case GG_F32: return launch_noise<float>(out, x, noise, noise_weight, bias, row_scale, alpha, scale, N, C, HW, st);
    case GG_F16: return launch_noise<__half>(out, x, noise, noise_weight, bias, row_scale, alpha, scale, N, C, HW, st);
    case GG_BF16: return launch_noise<__nv_bfloat16>(out, x, noise, noise_weight, bias, row_scale, alpha, scale, N, C, HW, st);
    default: return fail(GG_ERR_UNSUPPORTED, "noise_bias_act: dtype %d not supported", dtype);
  }
}

int64_t gg_channel_scale_workspace(int64_t rows, int64_t HW) {
  if (rows <= 0 || HW <= 0) return 0;
  const int64_t chunk = 16384;
  return rows * ((HW + chunk - 1) / chunk) * static_cast<int64_t>(sizeof(float));
}

int gg_channel_scale(void* out, float* row_dot, void* workspace, const void* x, const void* y, const float* s, int dtype,
                     int64_t rows, int64_t HW, void* stream) {
  if (rows < 0 || HW < 0) return fail(GG_ERR_BAD_ARG, "channel_scale: negative size");
  if (rows * HW == 0) return GG_OK;
  if (!out || !x || !s) return fail(GG_ERR_BAD_ARG, "channel_scale: null tensor");
  if (row_dot && (!y || !workspace)) return fail(GG_ERR_BAD_ARG, "channel_scale: row_dot needs y and a workspace");
  const int64_t chunk = 16384;
  const int K = static_cast<int>((HW + chunk - 1) / chunk);
  const int64_t grid = rows * K;
  if (grid > 0x7fffffffLL) return fail(GG_ERR_BAD_ARG, "channel_scale: too many rows");
  auto st = static_cast<cudaStream_t>(stream);
  float* partial = row_dot ? static_cast<float*>(workspace) : nullptr;
  const void* yy = row_dot ? y : nullptr;
  if (HW < 1024) {
    const unsigned g = static_cast<unsigned>((rows + (kThreads / 32) - 1) / (kThreads / 32));
    switch (dtype) {
      case GG_F32: channel_scale_rows_kernel<float><<<g, kThreads, 0, st>>>(static_cast<float*>(out), row_dot, static_cast<const float*>(x), static_cast<const float*>(yy), s, rows, HW); break;
      case GG_F16: channel_scale_rows_kernel<__half><<<g, kThreads, 0, st>>>(static_cast<__half*>(out), row_dot, static_cast<const __half*>(x), static_cast<const __half*>(yy), s, rows, HW); break;
      case GG_BF16: channel_scale_rows_kernel<__nv_bfloat16><<<g, kThreads, 0, st>>>(static_cast<__nv_bfloat16*>(out), row_dot, static_cast<const __nv_bfloat16*>(x), static_cast<const __nv_bfloat16*>(yy), s, rows, HW); break;
      default: return fail(GG_ERR_UNSUPPORTED, "channel_scale: dtype %d not supported", dtype);
    }
    GG_CHECK_LAUNCH("channel_scale_rows launch");
    return GG_OK;
  }
#define GG_CS(T_)                                                                                                   \
  do {                                                                                                              \
    constexpr int V = 16 / sizeof(T_);                                                                              \
    const bool vec = (HW % V == 0) && aligned16(out) && aligned16(x) && (!yy || aligned16(yy));                     \
    if (vec)                                                                                                        \
      channel_scale_kernel<T_, V><<<static_cast<unsigned>(grid), kThreads, 0, st>>>(                                \
          static_cast<T_*>(out), partial, static_cast<const T_*>(x), static_cast<const T_*>(yy), s, HW, chunk, K);  \
    else                                                                                                            \
      channel_scale_kernel<T_, 1><<<static_cast<unsigned>(grid), kThreads, 0, st>>>(                                \
          static_cast<T_*>(out), partial, static_cast<const T_*>(x), static_cast<const T_*>(yy), s, HW, chunk, K);  \
  } while (0)
  switch (dtype) {
    case GG_F32: GG_CS(float); break;
    case GG_F16: GG_CS(__half); break;
    case GG_BF16: GG_CS(__nv_bfloat16); break;
    default: return fail(GG_ERR_UNSUPPORTED, "channel_scale: dtype %d not supported", dtype);
  }
#undef GG_CS
  GG_CHECK_LAUNCH("channel_scale launch");
  if (row_dot) {
    row_finish_kernel<<<static_cast<unsigned>((rows + 255) / 256), 256, 0, st>>>(row_dot, partial, rows, K);
    GG_CHECK_LAUNCH("row_finish launch");
  }
  return GG_OK;
}

int64_t gg_bias_act_backward_workspace(int64_t N, int64_t C, int64_t HW) {
  if (N <= 0 || C <= 0 || HW <= 0) return 0;
  return N * C * gg::bwd_geom(HW).chunks_per_row * static_cast<int64_t>(sizeof(float));
}

int gg_bias_act_backward(void* gx, float* grad_bias, void* workspace, const void* g, const void* out,
                         int dtype, float alpha, float scale, int64_t N, int64_t C, int64_t HW,
                         void* stream) {
  if (N < 0 || C < 0 || HW < 0) return fail(GG_ERR_BAD_ARG, "bias_act_backward: negative size");
  auto st = static_cast<cudaStream_t>(stream);
  if (N * C * HW == 0) {
    if (grad_bias && C > 0) {
      cudaError_t e = cudaMemsetAsync(grad_bias, 0, C * sizeof(float), st);
      if (e != cudaSuccess) return cuda_fail(e, "bias_act_backward memset");
    }
    return GG_OK;
  }
  if (!gx || !g || !out) return fail(GG_ERR_BAD_ARG, "bias_act_backward: null tensor");
  if (grad_bias && !workspace) return fail(GG_ERR_BAD_ARG, "bias_act_backward: grad_bias needs a workspace");
  switch (dtype) {
    case GG_F32: return launch_bwd<float>(gx, grad_bias, workspace, g, out, alpha, scale, N, C, HW, st);
    case GG_F16: return launch_bwd<__half>(gx, grad_bias, workspace, g, out, alpha, scale, N, C, HW, st);
    case GG_BF16: return launch_bwd<__nv_bfloat16>(gx, grad_bias, workspace, g, out, alpha, scale, N, C, HW, st);
    default: return fail(GG_ERR_UNSUPPORTED, "bias_act_backward: dtype %d not supported", dtype);
  }
}

}  // extern "C"
