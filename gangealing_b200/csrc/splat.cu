// splat.cu -- Gaussian forward splatting of points into an image (sm_100a).
//
// Replaces reference utils/splat2d_cuda/src/splat_gpu_impl.cu:41-96 (one 32-thread block per 32 points,
// (C+1) scalar float atomics per footprint pixel) and the five ATen passes of splat_gpu.c:20-41
// (zeros, clone, clamp, add, divide).
//
//  * accumulators are interleaved per pixel -- slot 0 = sum of alpha, slots 1..C = sum of alpha*value[c],
//    padded to a multiple of 4 floats -- so one footprint pixel receives ONE 16-byte vector reduction
//    (red.global.add.v4.f32, sm_90+) per group of 4 slots instead of C+1 scalar atomics;
//  * warp aggregation: consecutive points are spatial neighbours (callers splat rasterised masks), so lanes
//    that target the same pixel in the same footprint step are combined with a segmented shuffle scan and
//    only the last lane of each run issues the reduction;
//  * normalisation (input + sum) / (alpha [clamped >= 1 if soft] + 1e-8) and the NCHW re-layout are one
//    fused pass.
// Float atomics make the summation order (hence the last bits) run-to-run dependent, exactly as in the
// reference; the SET of touched pixels is deterministic.
// HBM: algorithmic bytes 4*N*(P*(2+C) + (C+1)*H*W + 2*C*H*W); the scatter itself is L2-atomic-bound.
#include "common.cuh"

namespace gg {
namespace {

struct SplatParams {
  int64_t n;
  int64_t points;   // P
  int c, h, w;
  int slots;        // (C + 1) rounded up to a multiple of 4
};

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// One lane per point.  All lanes of a warp walk the same (dy, dx) footprint schedule (the warp maximum), with a
// per-lane validity flag, so that same-pixel lanes can be merged.
template <int GROUPS>  // slots / 4 handled with compile-time unrolling for GROUPS <= 2; generic loop otherwise
__global__ void __launch_bounds__(256)
splat_scatter_kernel(float* __restrict__ acc, const float* __restrict__ coords, const float* __restrict__ values,
                     const float* __restrict__ sigma, SplatParams p, int64_t total) {
  const int lane = threadIdx.x & 31;
  const int64_t warp_base = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) - lane;
  for (int64_t base = warp_base; base < total; base += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t index = base + lane;
    const bool live = index < total;
    int64_t n = 0;
    float x = -1.f, y = -1.f, norm = 0.f;
    int t = 0, b = -1, l = 0, r = -1;
    const float* val = values;
    if (live) {
      n = index / p.points;
      const float2 xy = *reinterpret_cast<const float2*>(coords + index * 2);
      x = xy.x; y = xy.y;
      const float sd = __ldg(sigma + n);
      const float len = 2.f * sd;
      norm = -1.f / (2.f * sd * sd);
      val = values + index * p.c;
      // points outside the image are ignored (splat_gpu_impl.cu:76); bounds: :78-81
      if (x >= 0.f && x < static_cast<float>(p.w) && y >= 0.f && y < static_cast<float>(p.h)) {
        t = static_cast<int>(fmaxf(0.f, floorf(y - len)));
        b = static_cast<int>(fminf(static_cast<float>(p.h - 1), ceilf(y + len)));
        l = static_cast<int>(fmaxf(0.f, floorf(x - len)));
        r = static_cast<int>(fminf(static_cast<float>(p.w - 1), ceilf(x + len)));
      }
    }
    const int nh = b - t + 1, nwid = r - l + 1;  // <= 0 for dead lanes
    int max_h = nh, max_w = nwid;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      max_h = max(max_h, __shfl_xor_sync(0xffffffffu, max_h, o));
      max_w = max(max_w, __shfl_xor_sync(0xffffffffu, max_w, o));
    }
    float v[GROUPS * 4];
#pragma unroll
    for (int s = 0; s < GROUPS * 4; ++s) v[s] = (live && s >= 1 && s <= p.c && nh > 0) ? __ldg(val + s - 1) : 0.f;
    float* acc_n = acc + n * p.h * static_cast<int64_t>(p.w) * p.slots;
    for (int dy = 0; dy < max_h; ++dy) {
      for (int dx = 0; dx < max_w; ++dx) {
        const bool ok = dy < nh && dx < nwid;
        const int py = t + dy, px = l + dx;
        // key: unique negative for idle lanes so they never merge
        const int64_t key = ok ? (n * p.h + py) * static_cast<int64_t>(p.w) + px : -1 - lane;
        float alpha = 0.f;
        if (ok) {
          const float ddx = static_cast<float>(px) - x, ddy = static_cast<float>(py) - y;
          alpha = expf(norm * (ddx * ddx + ddy * ddy));
        }
        float s[GROUPS * 4];
        s[0] = alpha;
#pragma unroll
        for (int q = 1; q < GROUPS * 4; ++q) s[q] = alpha * v[q];
        // Merge CONTIGUOUS runs of lanes that hit the same pixel (rasterised point sets put duplicates next to each
        // other): segmented inclusive scan bounded by the first lane of the run, then the run's last lane issues
        // the reduction.  Non-adjacent duplicates simply issue their own reductions.
        const unsigned peers = __match_any_sync(0xffffffffu, key);
        const unsigned below = ~peers & ((1u << lane) - 1u);          // lanes below me that are NOT my pixel
        const int run_start = below ? 32 - __clz(below) : 0;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const bool take = lane - d >= run_start;
#pragma unroll
          for (int q = 0; q < GROUPS * 4; ++q) {
            const float up = __shfl_up_sync(0xffffffffu, s[q], d);
            if (take) s[q] += up;
          }
        }
        const bool tail = ok && (lane == 31 || !((peers >> (lane + 1)) & 1u));
        if (tail) {
          float* dst = acc_n + (static_cast<int64_t>(py) * p.w + px) * p.slots;
#pragma unroll
          for (int g = 0; g < GROUPS; ++g) red_add_v4(dst + g * 4, s[g * 4], s[g * 4 + 1], s[g * 4 + 2], s[g * 4 + 3]);
        }
      }
    }
  }
}

// generic channel count: scalar atomics per slot (C > 7); still interleaved accumulators
__global__ void __launch_bounds__(256)
splat_scatter_generic_kernel(float* __restrict__ acc, const float* __restrict__ coords, const float* __restrict__ values,
                             const float* __restrict__ sigma, SplatParams p, int64_t total) {
  for (int64_t index = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; index < total;
       index += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t n = index / p.points;
    const float x = coords[index * 2], y = coords[index * 2 + 1];
    if (!(x >= 0.f && x < static_cast<float>(p.w) && y >= 0.f && y < static_cast<float>(p.h))) continue;
    const float sd = sigma[n], len = 2.f * sd, norm = -1.f / (2.f * sd * sd);
    const int t = static_cast<int>(fmaxf(0.f, floorf(y - len)));
    const int b = static_cast<int>(fminf(static_cast<float>(p.h - 1), ceilf(y + len)));
    const int l = static_cast<int>(fmaxf(0.f, floorf(x - len)));
    const int r = static_cast<int>(fminf(static_cast<float>(p.w - 1), ceilf(x + len)));
    const float* val = values + index * p.c;
    float* acc_n = acc + n * p.h * static_cast<int64_t>(p.w) * p.slots;
    for (int py = t; py <= b; ++py)
      for (int px = l; px <= r; ++px) {
        const float ddx = static_cast<float>(px) - x, ddy = static_cast<float>(py) - y;
        const float alpha = expf(norm * (ddx * ddx + ddy * ddy));
        float* dst = acc_n + (static_cast<int64_t>(py) * p.w + px) * p.slots;
        atomicAdd(dst, alpha);
        for (int c = 0; c < p.c; ++c) atomicAdd(dst + 1 + c, alpha * val[c]);
      }
  }
}

// out[n,c,y,x] = (input[n,c,y,x] + acc[n,y,x,1+c]) / (alpha' + 1e-8)     (splat_gpu.c:36-41)
__global__ void __launch_bounds__(256)
splat_normalize_kernel(float* __restrict__ out, const float* __restrict__ input, const float* __restrict__ acc,
                       SplatParams p, int soft, int64_t total) {
  for (int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t hw = static_cast<int64_t>(p.h) * p.w;
    const int64_t pix = idx % hw;
    const int64_t nc = idx / hw;
    const int c = static_cast<int>(nc % p.c);
    const int64_t n = nc / p.c;
    const float* a = acc + (n * hw + pix) * p.slots;
    float alpha = a[0];
    if (soft) alpha = fmaxf(alpha, 1.f);
    out[idx] = (input[idx] + a[1 + c]) / (alpha + 1e-8f);
  }
}

inline int splat_grid(int64_t total, int threads) {
  int64_t g = (total + threads - 1) / threads;
  const int64_t cap = static_cast<int64_t>(sm_count()) * 8;
  return static_cast<int>(g < cap ? (g > 0 ? g : 1) : cap);
}

}  // namespace
}  // namespace gg

using namespace gg;

extern "C" {

int64_t gg_splat2d_workspace(int64_t N, int C, int H, int W) {
  if (N < 0 || C < 0 || H < 0 || W < 0) return -1;
  const int slots = ((C + 1) + 3) / 4 * 4;
  return N * H * static_cast<int64_t>(W) * slots * static_cast<int64_t>(sizeof(float));
}

int gg_splat2d_forward(float* out, void* workspace, const float* input, const float* coordinates, const float* values,
                       const float* sigma, int64_t N, int64_t P, int C, int H, int W, int soft_normalize,
                       void* stream) {
  if (N < 0 || P < 0 || C < 0 || H < 0 || W < 0) return fail(GG_ERR_BAD_ARG, "splat2d: negative size");
  const int64_t numel = N * C * H * static_cast<int64_t>(W);
  if (numel == 0) return GG_OK;  // reference returns the (empty) clone (splat_gpu.c:23-26)
  if (!out || !input || !workspace || !sigma) return fail(GG_ERR_BAD_ARG, "splat2d: null tensor");
  if (P > 0 && (!coordinates || !values)) return fail(GG_ERR_BAD_ARG, "splat2d: null points");
  auto st = static_cast<cudaStream_t>(stream);
  SplatParams p;
  p.n = N; p.points = P; p.c = C; p.h = H; p.w = W;
  p.slots = ((C + 1) + 3) / 4 * 4;
  cudaError_t e = cudaMemsetAsync(workspace, 0, static_cast<size_t>(gg_splat2d_workspace(N, C, H, W)), st);
  if (e != cudaSuccess) return cuda_fail(e, "splat2d workspace memset");
  float* acc = static_cast<float*>(workspace);
  const int64_t total = N * P;
  if (total > 0) {
    const int grid = splat_grid(total, 256);
    if (p.slots == 4)
      splat_scatter_kernel<1><<<grid, 256, 0, st>>>(acc, coordinates, values, sigma, p, total);
    else if (p.slots == 8)
      splat_scatter_kernel<2><<<grid, 256, 0, st>>>(acc, coordinates, values, sigma, p, total);
    else
      splat_scatter_generic_kernel<<<grid, 256, 0, st>>>(acc, coordinates, values, sigma, p, total);
    GG_CHECK_LAUNCH("splat_scatter launch");
  }
  splat_normalize_kernel<<<splat_grid(numel, 256), 256, 0, st>>>(out, input, acc, p, soft_normalize ? 1 : 0, numel);
  GG_CHECK_LAUNCH("splat_normalize launch");
  return GG_OK;
}

}  // extern "C"
