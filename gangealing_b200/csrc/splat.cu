// splat.cu -- Gaussian forward splatting of points into an image (sm_100a).
//
// Replaces reference utils/splat2d_cuda/src/splat_gpu_impl.cu:41-96 (one 32-thread block per 32 points,
// (C+1) scalar float atomics per footprint pixel) and the five ATen passes of splat_gpu.c:20-41
// (zeros, clone, clamp, add, divide).
//
//  * accumulators are interleaved per pixel -- slot 0 = sum of alpha, slots 1..C = sum of alpha*value[c],
//    padded to a multiple of 4 floats -- so one footprint pixel receives ONE 16-byte vector reduction
//    (red.global.add.v4.f32, sm_90+) per group of 4 slots instead of C+1 scalar atomics: a quarter of the reference's
//    atomic traffic for RGB colours (C = 3) or a mask (C = 1);
//  * one thread per point, 64-thread CTAs (a dense mask of 4e5 points fills the machine, a sparse one of 2.5e4 still
//    spreads over ~400 CTAs), fire-and-forget reductions: no return value, no ordering between them;
//  * normalisation (input + sum) / (alpha [clamped >= 1 if soft] + 1e-8) and the NCHW re-layout are one fused pass.
// Cross-point aggregation was built twice and measured slower BOTH times, so it is not here (numbers under profiles/):
//   - round 1: one lane per point, same-pixel lanes merged per footprint step with match.any + a segmented shuffle scan --
//     0.55-0.75x the reference at sigma 1.3 (r02_opbench_vs_reference_b32_before.txt): the merge cost more than it saved;
//   - round 2a: a "torus" of per-lane register accumulators walking raster-ordered points (one flush per pixel when the
//     footprint window moves off it) -- 0.72-0.85x the reference in 3 of 4 config-4 cases (r02_opbench_vs_reference_b32.json):
//     all 32 lanes execute bookkeeping for every point while a sigma-0.3 footprint has 9 pixels;
//   - round 2b (this kernel): 1.3-1.7x the reference in all four cases (r02_splat_modes.txt).  At the dense end (4e5 points,
//     7x7 footprints: 19.8 M vector reductions in 81 us = 244 G/s) the scatter runs at the L2's reduction rate; the v4 form
//     is what moved the needle, not merging.
// Float atomics make the summation order (hence the last bits) run-to-run dependent, exactly as in the
// reference; the SET of touched pixels is deterministic.
// HBM: algorithmic bytes 4*N*(P*(2+C) + (C+1)*H*W + 2*C*H*W); the scatter itself is L2-reduction-bound.
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

// LOOKUP (reference spatial_transformer.py:141-157 `uncongeal_points` + helpers.py:178-187): the
// points arrive as QUERY coordinates in the congealed frame; their image positions are looked up in the STN's sampling
// grid -- F.grid_sample(grid as a 2-channel image, query, 'border', align_corners=False) -- and un-normalised to pixels
// (spatial_transformer.py:621-623) as the points are loaded, instead of a grid_sample launch + 4 elementwise launches.
struct LookupParams {
  const float* grid;     // (N, gh, gw, 2)
  int gh, gw;
  float k, m;            // unnormalize: ((g / k) / 2 + 0.5) * m,  k = (res-1)/res, m = out_res - 1
  float* points_out;     // (N, P, 2) or null: the looked-up pixel coordinates
};

__device__ __forceinline__ float2 lookup_point(const LookupParams& lk, int64_t n, float qx, float qy) {
  // ATen grid_sampler_2d, bilinear, padding_mode=border, align_corners=False
  float ix = ((qx + 1.f) * lk.gw - 1.f) / 2.f, iy = ((qy + 1.f) * lk.gh - 1.f) / 2.f;
  ix = fminf(fmaxf(ix, 0.f), static_cast<float>(lk.gw - 1));
  iy = fminf(fmaxf(iy, 0.f), static_cast<float>(lk.gh - 1));
  const float fx = floorf(ix), fy = floorf(iy);
  const int x0 = static_cast<int>(fx), y0 = static_cast<int>(fy);
  const float wx1 = ix - fx, wx0 = (fx + 1.f) - ix, wy1 = iy - fy, wy0 = (fy + 1.f) - iy;
  const float* g = lk.grid + n * lk.gh * static_cast<int64_t>(lk.gw) * 2;
  float ox = 0.f, oy = 0.f;
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      const int yy = y0 + a, xx = x0 + b;
      if (yy >= 0 && yy < lk.gh && xx >= 0 && xx < lk.gw) {
        const float2 v = __ldg(reinterpret_cast<const float2*>(g + (static_cast<int64_t>(yy) * lk.gw + xx) * 2));
        const float w = (a ? wy1 : wy0) * (b ? wx1 : wx0);
        ox = fmaf(v.x, w, ox); oy = fmaf(v.y, w, oy);
      }
    }
  return make_float2(((ox / lk.k) / 2.f + 0.5f) * lk.m, ((oy / lk.k) / 2.f + 0.5f) * lk.m);
}

// One thread per point, one 16-byte reduction per footprint pixel and group of 4 accumulator slots (GROUPS = 1: C <= 3,
// GROUPS = 2: C <= 7).  Same footprint schedule as the reference kernel (splat_gpu_impl.cu:60-94: rows t..b, columns l..r).
template <int GROUPS, bool LOOKUP>
__global__ void __launch_bounds__(64)
splat_direct_kernel(float* __restrict__ acc, const float* __restrict__ coords, const float* __restrict__ values,
                    const float* __restrict__ sigma, SplatParams p, int64_t total, LookupParams lk) {
  const int64_t index = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (index >= total) return;
  const int64_t n = index / p.points;
  float2 xy = __ldg(reinterpret_cast<const float2*>(coords + index * 2));
  if (LOOKUP) {
    xy = lookup_point(lk, n, xy.x, xy.y);
    if (lk.points_out) *reinterpret_cast<float2*>(lk.points_out + index * 2) = xy;
  }
  const float x = xy.x, y = xy.y;
  // points outside the image are ignored (splat_gpu_impl.cu:76); bounds: :78-81
  if (!(x >= 0.f && x < static_cast<float>(p.w) && y >= 0.f && y < static_cast<float>(p.h))) return;
  const float sd = __ldg(sigma + n), len = 2.f * sd, norm = -1.f / (2.f * sd * sd);
  const int t = static_cast<int>(fmaxf(0.f, floorf(y - len))), b = static_cast<int>(fminf(static_cast<float>(p.h - 1), ceilf(y + len)));
  const int l = static_cast<int>(fmaxf(0.f, floorf(x - len))), r = static_cast<int>(fminf(static_cast<float>(p.w - 1), ceilf(x + len)));
  const float* val = values + index * p.c;
  float v[GROUPS * 4];
  v[0] = 1.f;                                             // slot 0 accumulates alpha itself
#pragma unroll
  for (int q = 1; q < GROUPS * 4; ++q) v[q] = (q <= p.c) ? __ldg(val + q - 1) : 0.f;
  float* acc_n = acc + n * p.h * static_cast<int64_t>(p.w) * (GROUPS * 4);
  for (int py = t; py <= b; ++py) {
    const float ddy = static_cast<float>(py) - y;
    float* row = acc_n + static_cast<int64_t>(py) * p.w * (GROUPS * 4);
    for (int px = l; px <= r; ++px) {
      const float ddx = static_cast<float>(px) - x;
      const float a = expf(norm * (ddx * ddx + ddy * ddy));
      float* dst = row + px * (GROUPS * 4);
#pragma unroll
      for (int g = 0; g < GROUPS; ++g) red_add_v4(dst + 4 * g, a * v[4 * g], a * v[4 * g + 1], a * v[4 * g + 2], a * v[4 * g + 3]);
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

static int splat_impl(float* out, void* workspace, const float* input, const float* coordinates, const float* values,
                      const float* sigma, int64_t N, int64_t P, int C, int H, int W, int soft_normalize, const LookupParams* lk,
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
  if (lk && (p.slots != 4 || static_cast<int64_t>(H) * W >= 0x7fffffffLL))
    return fail(GG_ERR_UNSUPPORTED, "splat2d_lookup: the fused lookup serves C <= 3 (the call sites splat RGB colours or a 1-channel mask)");
  cudaError_t e = cudaMemsetAsync(workspace, 0, static_cast<size_t>(gg_splat2d_workspace(N, C, H, W)), st);
  if (e != cudaSuccess) return cuda_fail(e, "splat2d workspace memset");
  float* acc = static_cast<float*>(workspace);
  const int64_t total = N * P;
  if (total > 0) {
    if (p.slots <= 8 && static_cast<int64_t>(H) * W * p.slots < 0x7fffffffLL) {
      const int64_t blocks = (total + 63) / 64;
      if (blocks > 0x7fffffffLL) return fail(GG_ERR_BAD_ARG, "splat2d: too many points");
      const unsigned g = static_cast<unsigned>(blocks);
      if (lk) splat_direct_kernel<1, true><<<g, 64, 0, st>>>(acc, coordinates, values, sigma, p, total, *lk);
      else if (p.slots == 4) splat_direct_kernel<1, false><<<g, 64, 0, st>>>(acc, coordinates, values, sigma, p, total, LookupParams{});
      else splat_direct_kernel<2, false><<<g, 64, 0, st>>>(acc, coordinates, values, sigma, p, total, LookupParams{});
    } else {
      splat_scatter_generic_kernel<<<splat_grid(total, 256), 256, 0, st>>>(acc, coordinates, values, sigma, p, total);
    }
    GG_CHECK_LAUNCH("splat_scatter launch");
  }
  splat_normalize_kernel<<<splat_grid(numel, 256), 256, 0, st>>>(out, input, acc, p, soft_normalize ? 1 : 0, numel);
  GG_CHECK_LAUNCH("splat_normalize launch");
  return GG_OK;
}

int gg_splat2d_forward(float* out, void* workspace, const float* input, const float* coordinates, const float* values,
                       const float* sigma, int64_t N, int64_t P, int C, int H, int W, int soft_normalize,
                       void* stream) {
  return splat_impl(out, workspace, input, coordinates, values, sigma, N, P, C, H, W, soft_normalize, nullptr, stream);
}

int gg_splat2d_lookup_forward(float* out, float* points_out, void* workspace, const float* input, const float* grid,
                              const float* query, const float* values, const float* sigma, int64_t N, int64_t P, int C,
                              int H, int W, int grid_h, int grid_w, float unnorm_k, float unnorm_m, int soft_normalize,
                              void* stream) {
  if (grid_h < 1 || grid_w < 1 || !(unnorm_k != 0.f)) return fail(GG_ERR_BAD_ARG, "splat2d_lookup: bad grid geometry");
  if (N * P > 0 && !grid) return fail(GG_ERR_BAD_ARG, "splat2d_lookup: null grid");
  LookupParams lk;
  lk.grid = grid; lk.gh = grid_h; lk.gw = grid_w; lk.k = unnorm_k; lk.m = unnorm_m; lk.points_out = points_out;
  return splat_impl(out, workspace, input, query, values, sigma, N, P, C, H, W, soft_normalize, &lk, stream);
}

}  // extern "C"
