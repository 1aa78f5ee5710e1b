// splat.cu -- Gaussian forward splatting of points into an image (sm_100a).
//
// Replaces reference utils/splat2d_cuda/src/splat_gpu_impl.cu:41-96 (one 32-thread block per 32 points,
// (C+1) scalar float atomics per footprint pixel) and the five ATen passes of splat_gpu.c:20-41
// (zeros, clone, clamp, add, divide).
//
//  * accumulators are interleaved per pixel -- slot 0 = sum of alpha, slots 1..C = sum of alpha*value[c],
//    padded to a multiple of 4 floats -- so one footprint pixel receives ONE 16-byte vector reduction
//    (red.global.add.v4.f32, sm_90+) per group of 4 slots instead of C+1 scalar atomics;
//  * warp aggregation ACROSS points (splat_torus_kernel, C <= 3): a warp walks a contiguous chunk of points, its lanes
//    own the slots of a T x T torus of pixels (slot = (py mod T, px mod T), T >= the footprint extent) and keep the
//    running sums of "their" pixel in registers while consecutive points keep hitting it -- callers splat rasterised
//    masks in raster order, so a pixel collects all of its ~(footprint width x points per pixel) contributions in ONE
//    lane and is flushed with ONE 16-byte reduction when the footprint window moves off it.  No shuffles, no
//    match.any: the round-1 kernel (one lane per point, same-pixel lanes merged with a segmented shuffle scan per
//    footprint step) spent more in the merge than it saved and was slower than the reference at sigma 1.3
//    (profiles/r02_opbench_vs_reference_b32_before.txt); it is kept for C in 4..7;
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

// ---------------------------------------------------------------- torus accumulation (C <= 3: 4 accumulator slots per pixel)
// T = torus side (4, 8 or 16), chosen per warp from sigma[n]: T >= 2*ceil(2*sigma) + 2 >= the footprint extent, so a
// footprint never wraps onto itself.  A lane owns slots lane, lane + 32, ... (T*T/32 of them; 16 lanes idle when T == 4).
constexpr int kMaxTorusSlots = 8;   // T = 16

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

template <bool LOOKUP>
__global__ void __launch_bounds__(128)
splat_torus_kernel(float* __restrict__ acc, const float* __restrict__ coords, const float* __restrict__ values,
                   const float* __restrict__ sigma, SplatParams p, int chunk, int chunks_per_sample, LookupParams lk) {
  const int lane = threadIdx.x & 31;
  const int64_t warp_id = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t n = warp_id / chunks_per_sample;
  if (n >= p.n) return;
  const int64_t q0 = (warp_id - n * chunks_per_sample) * chunk;
  const int64_t q1 = min(q0 + static_cast<int64_t>(chunk), p.points);
  const float sd = __ldg(sigma + n);
  const float len = 2.f * sd;
  const float norm = -1.f / (2.f * sd * sd);
  const int extent = static_cast<int>(floorf(2.f * len + 3.f));   // >= (b - t + 1) for every point
  float* acc_n = acc + n * p.h * static_cast<int64_t>(p.w) * 4;
  const float* cpt = coords + n * p.points * 2;
  const float* vpt = values + n * p.points * p.c;
  if (extent > 16) {
    // very wide footprints: lanes stride over the window of each point, one reduction per (point, pixel)
    for (int64_t q = q0; q < q1; ++q) {
      float x = __ldg(cpt + q * 2), y = __ldg(cpt + q * 2 + 1);
      if (LOOKUP) {
        const float2 pp = lookup_point(lk, n, x, y);
        x = pp.x; y = pp.y;
        if (lk.points_out && lane == 0) *reinterpret_cast<float2*>(lk.points_out + (n * p.points + q) * 2) = pp;
      }
      if (!(x >= 0.f && x < static_cast<float>(p.w) && y >= 0.f && y < static_cast<float>(p.h))) continue;
      const int t = static_cast<int>(fmaxf(0.f, floorf(y - len))), b = static_cast<int>(fminf(static_cast<float>(p.h - 1), ceilf(y + len)));
      const int l = static_cast<int>(fmaxf(0.f, floorf(x - len))), r = static_cast<int>(fminf(static_cast<float>(p.w - 1), ceilf(x + len)));
      const int wd = r - l + 1, cnt = wd * (b - t + 1);
      const float v0 = p.c > 0 ? __ldg(vpt + q * p.c) : 0.f, v1 = p.c > 1 ? __ldg(vpt + q * p.c + 1) : 0.f;
      const float v2 = p.c > 2 ? __ldg(vpt + q * p.c + 2) : 0.f;
      for (int e = lane; e < cnt; e += 32) {
        const int py = t + e / wd, px = l + e % wd;
        const float ddx = static_cast<float>(px) - x, ddy = static_cast<float>(py) - y;
        const float a = expf(norm * (ddx * ddx + ddy * ddy));
        red_add_v4(acc_n + (static_cast<int64_t>(py) * p.w + px) * 4, a, a * v0, a * v1, a * v2);
      }
    }
    return;
  }
  const int T = extent <= 4 ? 4 : (extent <= 8 ? 8 : 16);
  const int tmask = T - 1, tshift = (T == 4) ? 2 : (T == 8 ? 3 : 4);
  // T == 4: 16 slots -- the two half-warps run two independent tori on alternate points (a pixel held by both is simply
  // flushed twice); otherwise T*T/32 slots per lane
  const int npar = (T == 4) ? 2 : 1;
  const int sub = (T == 4) ? (lane >> 4) : 0;
  const int lane_slot = (T == 4) ? (lane & 15) : lane;
  const int nslots = (T == 4) ? 1 : (T * T) / 32;
  int hid[kMaxTorusSlots];
  float a0[kMaxTorusSlots], a1[kMaxTorusSlots], a2[kMaxTorusSlots], a3[kMaxTorusSlots];
#pragma unroll
  for (int k = 0; k < kMaxTorusSlots; ++k) { hid[k] = -1; a0[k] = a1[k] = a2[k] = a3[k] = 0.f; }
  for (int64_t qb = q0; qb < q1; qb += 32) {
    // one coalesced load per lane fetches 32 points; they are then broadcast with shuffles (no per-point load latency)
    const int64_t ql = qb + lane;
    float lx = -1.f, ly = -1.f, lv0 = 0.f, lv1 = 0.f, lv2 = 0.f;
    if (ql < q1) {
      float2 xy = __ldg(reinterpret_cast<const float2*>(cpt + ql * 2));
      if (LOOKUP) {
        xy = lookup_point(lk, n, xy.x, xy.y);
        if (lk.points_out) *reinterpret_cast<float2*>(lk.points_out + (n * p.points + ql) * 2) = xy;
      }
      lx = xy.x; ly = xy.y;
      if (p.c > 0) lv0 = __ldg(vpt + ql * p.c);
      if (p.c > 1) lv1 = __ldg(vpt + ql * p.c + 1);
      if (p.c > 2) lv2 = __ldg(vpt + ql * p.c + 2);
    }
    const int cnt = static_cast<int>(min(static_cast<int64_t>(32), q1 - qb));
    for (int i = 0; i < cnt; i += npar) {
      const int src = min(i + sub, 31);
      const float x = __shfl_sync(0xffffffffu, lx, src), y = __shfl_sync(0xffffffffu, ly, src);
      const float v0 = __shfl_sync(0xffffffffu, lv0, src), v1 = __shfl_sync(0xffffffffu, lv1, src);
      const float v2 = __shfl_sync(0xffffffffu, lv2, src);
      // points outside the image are ignored (splat_gpu_impl.cu:76); bounds: :78-81
      if (i + sub >= cnt || !(x >= 0.f && x < static_cast<float>(p.w) && y >= 0.f && y < static_cast<float>(p.h))) continue;
      const int t = static_cast<int>(fmaxf(0.f, floorf(y - len))), b = static_cast<int>(fminf(static_cast<float>(p.h - 1), ceilf(y + len)));
      const int l = static_cast<int>(fmaxf(0.f, floorf(x - len))), r = static_cast<int>(fminf(static_cast<float>(p.w - 1), ceilf(x + len)));
#pragma unroll
      for (int k = 0; k < kMaxTorusSlots; ++k) {
        if (k < nslots) {
          const int slot = lane_slot + 32 * k;
          const int sy = slot >> tshift, sx = slot & tmask;
          const int px = l + ((sx - l) & tmask), py = t + ((sy - t) & tmask);   // the pixel of the window congruent to this slot
          if (px <= r && py <= b) {
            const int id = py * p.w + px;
            if (id != hid[k]) {
              if (hid[k] >= 0) red_add_v4(acc_n + static_cast<int64_t>(hid[k]) * 4, a0[k], a1[k], a2[k], a3[k]);
              hid[k] = id; a0[k] = a1[k] = a2[k] = a3[k] = 0.f;
            }
            const float ddx = static_cast<float>(px) - x, ddy = static_cast<float>(py) - y;
            const float a = expf(norm * (ddx * ddx + ddy * ddy));
            a0[k] += a; a1[k] = fmaf(a, v0, a1[k]); a2[k] = fmaf(a, v1, a2[k]); a3[k] = fmaf(a, v2, a3[k]);
          }
        }
      }
    }
  }
#pragma unroll
  for (int k = 0; k < kMaxTorusSlots; ++k)
    if (k < nslots && hid[k] >= 0) red_add_v4(acc_n + static_cast<int64_t>(hid[k]) * 4, a0[k], a1[k], a2[k], a3[k]);
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
    const int grid = splat_grid(total, 256);
    if (p.slots == 4 && static_cast<int64_t>(H) * W < 0x7fffffffLL) {
      // points per warp: long enough to aggregate (a raster row of a dense mask revisits a pixel ~footprint x density
      // times), short enough to fill the machine (~16 warps per SM)
      int64_t chunk = (P + 16LL * sm_count() - 1) / (16LL * sm_count());
      chunk = chunk < 64 ? 64 : (chunk > 512 ? 512 : chunk);
      const int64_t cps = (P + chunk - 1) / chunk;
      const int64_t warps = N * cps;
      const int64_t blocks = (warps + 3) / 4;
      if (blocks > 0x7fffffffLL || cps > 0x7fffffffLL) return fail(GG_ERR_BAD_ARG, "splat2d: too many points");
      if (lk)
        splat_torus_kernel<true><<<static_cast<unsigned>(blocks), 128, 0, st>>>(acc, coordinates, values, sigma, p, static_cast<int>(chunk),
                                                                               static_cast<int>(cps), *lk);
      else
        splat_torus_kernel<false><<<static_cast<unsigned>(blocks), 128, 0, st>>>(acc, coordinates, values, sigma, p, static_cast<int>(chunk),
                                                                                static_cast<int>(cps), LookupParams{});
    } else if (p.slots == 4)
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
