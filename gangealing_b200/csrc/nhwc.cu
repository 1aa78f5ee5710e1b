// nhwc.cu -- channels-last (N, H, W, C) variants of the StyledConv tail family (sm_100a).
//
// Why: cuDNN's tensor-core convolution kernels are NHWC-native; with NCHW activations every convolution of the step
// is bracketed by nchwToNhwc / nhwcToNchw conversion kernels (22 % of the step in profiles/r01_step_launches_b8).
// Keeping the generator's activations channels-last end to end removes them -- provided the hand-written kernels
// between the convolutions speak NHWC too.  In this layout a pixel's channels are contiguous, so with C % 4 == 0
// EVERYTHING is 16-byte aligned: the blur can use a real 4-D TMA tensor map (cp.async.bulk.tensor, SASS UTMALDG)
// whose out-of-bounds zero fill implements the padding of upfirdn2d for free.
//
// Same math / reference citations as the NCHW kernels (bias_act.cu, upfirdn2d.cu):
//   gg_noise_bias_act_nhwc       lrelu(rs[n,c]*x + nw*noise[n,p] + b[c])*gain                networks.py:291-298,346-348
//   gg_bias_act_backward_nhwc    gx = (out>0 ? g : a*g)*gain ; grad_bias[c] = sum gx          op/fused_act.py:20-38
//   gg_channel_scale_nhwc        x*s[n,c] (+ row_dot[n,c] = sum_p x*y)                        networks.py:236,243
//   gg_blur_nhwc                 upfirdn2d(up=down=1, <=4x4 separable or not) [+ fused tail]   networks.py:266 (+346-348)
#include <cuda.h>

#include "common.cuh"

namespace gg {
namespace {

constexpr int kT = 256;

// ------------------------------------------------------------------------------------------------ elementwise
__global__ void __launch_bounds__(kT)
noise_bias_act_nhwc_kernel(float* __restrict__ out, const float* __restrict__ x, const float* __restrict__ noise,
                           const float* __restrict__ noise_weight, const float* __restrict__ bias,
                           const float* __restrict__ row_scale, float alpha, float gain, int64_t n_vec, int c4,
                           int64_t hw) {
  const float nw = noise ? (noise_weight ? __ldg(noise_weight) : 1.f) : 0.f;
  const int64_t base = (static_cast<int64_t>(blockIdx.x) * 4) * kT + threadIdx.x;
  Vec16<float> xv[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int64_t v = base + static_cast<int64_t>(u) * kT;
    if (v < n_vec) xv[u] = ld_vec_stream(x + v * 4);
  }
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int64_t v = base + static_cast<int64_t>(u) * kT;
    if (v < n_vec) {
      const int64_t pix = v / c4;                      // n*hw + p
      const int cq = static_cast<int>(v - pix * c4);
      const int64_t n = pix / hw;
      const float4 b = bias ? __ldg(reinterpret_cast<const float4*>(bias) + cq) : make_float4(0.f, 0.f, 0.f, 0.f);
      const float4 r = row_scale ? __ldg(reinterpret_cast<const float4*>(row_scale + n * c4 * 4) + cq)
                                 : make_float4(1.f, 1.f, 1.f, 1.f);
      const float nz = noise ? nw * __ldg(noise + pix) : 0.f;
      const float bb[4] = {b.x, b.y, b.z, b.w}, rr[4] = {r.x, r.y, r.z, r.w};
      Vec16<float> o;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float t = fmaf(xv[u].v[k], rr[k], bb[k]) + nz;
        o.v[k] = (t > 0.f ? t : t * alpha) * gain;
      }
      st_vec_stream(out + v * 4, o);
    }
  }
}

// One CTA = `rows` consecutive pixels of one sample x all channels.  Thread = (channel quad, pixel lane); per-channel
// sums are reduced across the CTA's pixel lanes in shared memory and written as one partial row per CTA.
// MODE 0: channel_scale (out = x*s, dot = sum x*y)   MODE 1: bias_act backward (out = act'(ref)*x*gain, dot = sum out)
template <int MODE>
__global__ void __launch_bounds__(kT)
rowwise_nhwc_kernel(float* __restrict__ out, float* __restrict__ partial, const float* __restrict__ x,
                    const float* __restrict__ y, const float* __restrict__ s, float alpha, float gain, int c4,
                    int64_t hw, int chunk, int chunks_per_sample) {
  extern __shared__ float red[];                       // [pixel lanes][C] partial sums
  const int64_t n = blockIdx.x / chunks_per_sample;
  const int ck = blockIdx.x - n * chunks_per_sample;
  const int64_t p0 = static_cast<int64_t>(ck) * chunk, p1 = min(p0 + chunk, hw);
  const int lanes_p = kT / c4 > 0 ? kT / c4 : 1;       // pixel lanes when C/4 <= 256
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  // a thread owns channel quads cq = tid % c4 (+ k*kT when c4 > kT is not supported: C <= 1024)
  const int cq = threadIdx.x % c4;
  const int pl = threadIdx.x / c4;
  if (pl < lanes_p) {
    float4 sv = make_float4(1.f, 1.f, 1.f, 1.f);
    if (MODE == 0) sv = __ldg(reinterpret_cast<const float4*>(s + n * c4 * 4) + cq);
    auto body = [&](const float4 xv, const float4 yv, int64_t off) {
      float4 o;
      if (MODE == 0) {
        o = make_float4(xv.x * sv.x, xv.y * sv.y, xv.z * sv.z, xv.w * sv.w);
        if (y) {
          acc.x = fmaf(xv.x, yv.x, acc.x); acc.y = fmaf(xv.y, yv.y, acc.y);
          acc.z = fmaf(xv.z, yv.z, acc.z); acc.w = fmaf(xv.w, yv.w, acc.w);
        }
      } else {                                 // y = saved forward output
        o.x = (yv.x > 0.f ? xv.x : xv.x * alpha) * gain; o.y = (yv.y > 0.f ? xv.y : xv.y * alpha) * gain;
        o.z = (yv.z > 0.f ? xv.z : xv.z * alpha) * gain; o.w = (yv.w > 0.f ? xv.w : xv.w * alpha) * gain;
        acc.x += o.x; acc.y += o.y; acc.z += o.z; acc.w += o.w;
      }
      *reinterpret_cast<float4*>(out + off) = o;
    };
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    const bool has_y = (MODE == 1) || y != nullptr;
    int64_t p = p0 + pl;
    // 4 pixels per trip: all loads issued before the first dependent store (memory-level parallelism)
    for (; p + 3 * lanes_p < p1; p += 4 * lanes_p) {
      float4 xv[4], yv[4];
      int64_t off[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        off[u] = ((n * hw + p + u * lanes_p) * c4 + cq) * 4;
        xv[u] = __ldcs(reinterpret_cast<const float4*>(x + off[u]));
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) yv[u] = has_y ? __ldcs(reinterpret_cast<const float4*>(y + off[u])) : zero;
#pragma unroll
      for (int u = 0; u < 4; ++u) body(xv[u], yv[u], off[u]);
    }
    for (; p < p1; p += lanes_p) {
      const int64_t off = ((n * hw + p) * c4 + cq) * 4;
      const float4 xv = __ldcs(reinterpret_cast<const float4*>(x + off));
      const float4 yv = has_y ? __ldcs(reinterpret_cast<const float4*>(y + off)) : zero;
      body(xv, yv, off);
    }
  }
  if (partial) {
    float4* r4 = reinterpret_cast<float4*>(red);
    if (pl < lanes_p) r4[pl * c4 + cq] = acc;
    __syncthreads();
    if (threadIdx.x < c4) {
      float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int l = 0; l < lanes_p; ++l) {
        const float4 v = r4[l * c4 + threadIdx.x];
        t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
      }
      reinterpret_cast<float4*>(partial + static_cast<int64_t>(blockIdx.x) * c4 * 4)[threadIdx.x] = t;
    }
  }
}

// dst[r][c] = sum_k partial[(r*K + k)][c]    (r = sample for channel_scale; a single row for grad_bias)
// CTA = 32 channels x 32 k-lanes: each lane sums every 32nd partial row (4 independent loads per trip), then the
// 32 lane sums are combined through shared memory in a fixed order (deterministic).
__global__ void __launch_bounds__(1024)
nhwc_finish_kernel(float* __restrict__ dst, const float* __restrict__ partial, int64_t rows, int K, int C) {
  __shared__ float red[32][33];
  const int cblocks = (C + 31) / 32;
  const int64_t r = blockIdx.x / cblocks;
  const int c = (blockIdx.x - r * cblocks) * 32 + threadIdx.x;
  const int ky = threadIdx.y;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  if (c < C) {
    const float* base = partial + r * K * C + c;
    int k = ky;
    for (; k + 96 < K; k += 128) {
      a0 += base[static_cast<int64_t>(k) * C];
      a1 += base[static_cast<int64_t>(k + 32) * C];
      a2 += base[static_cast<int64_t>(k + 64) * C];
      a3 += base[static_cast<int64_t>(k + 96) * C];
    }
    for (; k < K; k += 32) a0 += base[static_cast<int64_t>(k) * C];
  }
  red[ky][threadIdx.x] = (a0 + a1) + (a2 + a3);
  __syncthreads();
  if (ky == 0 && c < C) {
    float t = 0.f;
#pragma unroll
    for (int q = 0; q < 32; ++q) t += red[q][threadIdx.x];
    dst[r * C + c] = t;
  }
}

// ------------------------------------------------------------------------------------------------ to-RGB (1x1, 3 outputs)
// out[n,o,p] = sum_i wm[n,o,i] * x[n,p,i] + bias[o] + skip[n,o,p]   x: NHWC, out/skip: planar (N, 3, HW)
// One pass over x (the only large operand).  A group of 8 lanes owns 4 consecutive pixels: every lane streams its
// channel quads (l + 8j) of the 4 pixels (4 independent 128-bit loads per trip), the 3 x C modulated filter sits in
// shared memory, and the 12 partial dot products are combined with 3 butterfly steps.
__global__ void __launch_bounds__(kT)
to_rgb_nhwc_fwd_kernel(float* __restrict__ out, const float* __restrict__ x, const float* __restrict__ wm,
                       const float* __restrict__ bias, const float* __restrict__ skip, int c4, int64_t hw, int chunk,
                       int chunks_per_sample) {
  extern __shared__ __align__(16) float wsm[];          // [3][C]
  const int64_t n = blockIdx.x / chunks_per_sample;
  const int ck = blockIdx.x - n * chunks_per_sample;
  const int64_t p0 = static_cast<int64_t>(ck) * chunk, p1 = min(p0 + chunk, hw);
  {
    const float4* src = reinterpret_cast<const float4*>(wm + n * 3 * c4 * 4);
    float4* dst = reinterpret_cast<float4*>(wsm);
    for (int i = threadIdx.x; i < 3 * c4; i += kT) dst[i] = __ldg(src + i);
  }
  __syncthreads();
  const float4* w4 = reinterpret_cast<const float4*>(wsm);
  const int l = threadIdx.x & 7, grp = threadIdx.x >> 3;
  const bool vec_ok = (hw & 3) == 0;
  const unsigned gmask = 0xffu << (threadIdx.x & 24);     // groups of one warp may leave the loop at different trips
  for (int64_t pb = p0 + grp * 4; pb < p1; pb += (kT / 8) * 4) {
    float acc[4][3];
#pragma unroll
    for (int u = 0; u < 4; ++u) acc[u][0] = acc[u][1] = acc[u][2] = 0.f;
    const float4* xp[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t pu = min(pb + u, p1 - 1);             // clamped: a tail pixel is recomputed, never stored
      xp[u] = reinterpret_cast<const float4*>(x + (n * hw + pu) * c4 * 4);
    }
    for (int q = l; q < c4; q += 8) {
      float4 xv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) xv[u] = __ldcs(xp[u] + q);
      const float4 w0 = w4[q], w1 = w4[c4 + q], w2 = w4[2 * c4 + q];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        acc[u][0] = fmaf(xv[u].x, w0.x, fmaf(xv[u].y, w0.y, fmaf(xv[u].z, w0.z, fmaf(xv[u].w, w0.w, acc[u][0]))));
        acc[u][1] = fmaf(xv[u].x, w1.x, fmaf(xv[u].y, w1.y, fmaf(xv[u].z, w1.z, fmaf(xv[u].w, w1.w, acc[u][1]))));
        acc[u][2] = fmaf(xv[u].x, w2.x, fmaf(xv[u].y, w2.y, fmaf(xv[u].z, w2.z, fmaf(xv[u].w, w2.w, acc[u][2]))));
      }
    }
#pragma unroll
    for (int m = 4; m >= 1; m >>= 1)
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int o = 0; o < 3; ++o) acc[u][o] += __shfl_xor_sync(gmask, acc[u][o], m);
    if (l < 3) {                                           // lane o of the group stores output plane o
      const int o = l;
      float r[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) r[u] = (o == 0 ? acc[u][0] : o == 1 ? acc[u][1] : acc[u][2]) + (bias ? __ldg(bias + o) : 0.f);
      const int64_t off = (n * 3 + o) * hw + pb;
      if (vec_ok && pb + 3 < p1) {
        if (skip) {
          const float4 sk = __ldg(reinterpret_cast<const float4*>(skip + off));
          r[0] += sk.x; r[1] += sk.y; r[2] += sk.z; r[3] += sk.w;
        }
        *reinterpret_cast<float4*>(out + off) = make_float4(r[0], r[1], r[2], r[3]);
      } else {
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (pb + u < p1) out[off + u] = r[u] + (skip ? __ldg(skip + off + u) : 0.f);
      }
    }
  }
}

// Backward: gx[n,p,i] = sum_o wm[n,o,i] g[n,o,p]  and  gwm[n,o,i] = sum_p g[n,o,p] x[n,p,i], one pass over x / gx.
// Thread = (channel quad, pixel lane) as in rowwise_nhwc_kernel; the three g planes are warp-broadcast loads.
__global__ void __launch_bounds__(kT)
to_rgb_nhwc_bwd_kernel(float* __restrict__ gx, float* __restrict__ partial, const float* __restrict__ g,
                       const float* __restrict__ x, const float* __restrict__ wm, int c4, int64_t hw, int chunk,
                       int chunks_per_sample) {
  extern __shared__ float red[];                          // [pixel lanes][3][C]
  const int64_t n = blockIdx.x / chunks_per_sample;
  const int ck = blockIdx.x - n * chunks_per_sample;
  const int64_t p0 = static_cast<int64_t>(ck) * chunk, p1 = min(p0 + chunk, hw);
  const int lanes_p = kT / c4 > 0 ? kT / c4 : 1;
  const int cq = threadIdx.x % c4, pl = threadIdx.x / c4;
  float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0, a2 = a0;
  if (pl < lanes_p) {
    const float4* wp = reinterpret_cast<const float4*>(wm + n * 3 * c4 * 4);
    const float4 w0 = __ldg(wp + cq), w1 = __ldg(wp + c4 + cq), w2 = __ldg(wp + 2 * c4 + cq);
    const float* g0 = g + n * 3 * hw;
    auto body = [&](int64_t p, const float4 xv, float s0, float s1, float s2) {
      float4 o;
      o.x = fmaf(w2.x, s2, fmaf(w1.x, s1, w0.x * s0)); o.y = fmaf(w2.y, s2, fmaf(w1.y, s1, w0.y * s0));
      o.z = fmaf(w2.z, s2, fmaf(w1.z, s1, w0.z * s0)); o.w = fmaf(w2.w, s2, fmaf(w1.w, s1, w0.w * s0));
      if (gx) *reinterpret_cast<float4*>(gx + ((n * hw + p) * c4 + cq) * 4) = o;
      a0.x = fmaf(s0, xv.x, a0.x); a0.y = fmaf(s0, xv.y, a0.y); a0.z = fmaf(s0, xv.z, a0.z); a0.w = fmaf(s0, xv.w, a0.w);
      a1.x = fmaf(s1, xv.x, a1.x); a1.y = fmaf(s1, xv.y, a1.y); a1.z = fmaf(s1, xv.z, a1.z); a1.w = fmaf(s1, xv.w, a1.w);
      a2.x = fmaf(s2, xv.x, a2.x); a2.y = fmaf(s2, xv.y, a2.y); a2.z = fmaf(s2, xv.z, a2.z); a2.w = fmaf(s2, xv.w, a2.w);
    };
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    int64_t p = p0 + pl;
    for (; p + 3 * lanes_p < p1; p += 4 * lanes_p) {
      float4 xv[4];
      float s[4][3];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int64_t pu = p + u * lanes_p;
        xv[u] = partial ? __ldcs(reinterpret_cast<const float4*>(x + ((n * hw + pu) * c4 + cq) * 4)) : zero;
        s[u][0] = __ldg(g0 + pu); s[u][1] = __ldg(g0 + hw + pu); s[u][2] = __ldg(g0 + 2 * hw + pu);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) body(p + u * lanes_p, xv[u], s[u][0], s[u][1], s[u][2]);
    }
    for (; p < p1; p += lanes_p) {
      const float4 xv = partial ? __ldcs(reinterpret_cast<const float4*>(x + ((n * hw + p) * c4 + cq) * 4)) : zero;
      body(p, xv, __ldg(g0 + p), __ldg(g0 + hw + p), __ldg(g0 + 2 * hw + p));
    }
  }
  if (partial) {
    float4* r4 = reinterpret_cast<float4*>(red);
    if (pl < lanes_p) {
      r4[(pl * 3 + 0) * c4 + cq] = a0; r4[(pl * 3 + 1) * c4 + cq] = a1; r4[(pl * 3 + 2) * c4 + cq] = a2;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 3 * c4; i += kT) {
      float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int q = 0; q < lanes_p; ++q) {
        const float4 v = r4[q * 3 * c4 + i];
        t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
      }
      reinterpret_cast<float4*>(partial + static_cast<int64_t>(blockIdx.x) * 3 * c4 * 4)[i] = t;
    }
  }
}

// ------------------------------------------------------------------------------------------------ blur (TMA tiled)
constexpr int kCB = 32;     // channels per CTA (128 B per pixel in the tile)
constexpr int kBX = 64;     // output columns per CTA; a thread owns 2 adjacent columns x 4 channels
constexpr int kRY = 4;      // input rows per pipeline stage
constexpr int kNS = 3;      // stages
constexpr int kTileW = kBX + 3;
constexpr int kStageFloats = kRY * kTileW * kCB;

struct BlurNhwcParams {
  int n, c, in_h, in_w, out_h, out_w;
  int pad_x0, pad_y0;
  int seg_rows;             // output rows per CTA (grid.y segments)
  int act;                  // FUSED: 1 linear, 3 lrelu
  float alpha, gain;
};

__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* tmap, int c0, int x0, int y0, int n0,
                                            uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
      ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(c0), "r"(x0), "r"(y0), "r"(n0), "r"(smem_u32(bar))
      : "memory");
}

// CTA = (sample n, 64-column block, 32-channel chunk, row segment).  Input rows stream through a 3-stage ring of
// {32 ch x 67 px x 4 rows} TMA boxes (zero-filled outside the image = upfirdn2d's padding); each thread slides a
// 4-row window of horizontal results for its 2 columns x 4 channels down the whole segment, so a row is read from
// shared memory once and from HBM once (+3 halo rows per segment, +3/64 halo columns).
template <bool FUSED, bool SEP>
__global__ void __launch_bounds__(kT, 2)
blur_nhwc_kernel(float* __restrict__ out, const __grid_constant__ CUtensorMap tmap, const float* __restrict__ filt,
                 int kh, int kw, const float* __restrict__ noise, const float* __restrict__ noise_weight,
                 const float* __restrict__ bias, const float* __restrict__ row_scale, BlurNhwcParams p) {
  extern __shared__ __align__(128) float tiles[];
  __shared__ uint64_t full_bar[kNS];
  const int tid = threadIdx.x;
  const int cq = tid & 7;                 // channel quad within the 32-channel chunk
  const int xg = tid >> 3;                // 0..31 -> columns 2*xg, 2*xg+1 of the block
  const int chunks = p.c / kCB;
  const int bx = blockIdx.x / chunks, cc = blockIdx.x - bx * chunks;
  const int n = blockIdx.z;
  const int oy0 = blockIdx.y * p.seg_rows;
  const int rows_out = min(p.seg_rows, p.out_h - oy0);
  const int x_out0 = bx * kBX;            // first output column of the block
  const int c0 = cc * kCB;
  // input row/col of tap (0,0) for output (oy0, x_out0)
  const int iy0 = oy0 - p.pad_y0, ix0 = x_out0 - p.pad_x0;
  const int rows_in = rows_out + 3;
  const int n_stage_iters = (rows_in + kRY - 1) / kRY;

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < kNS; ++s) mbar_init(&full_bar[s], 1);
    mbar_fence_init();
  }
  __syncthreads();
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < kNS - 1; ++s)
      if (s < n_stage_iters) {
        mbar_expect_tx(&full_bar[s], kStageFloats * 4);
        tma_load_4d(tiles + s * kStageFloats, &tmap, c0, ix0, iy0 + s * kRY, n, &full_bar[s]);
      }
  }

  // taps (flipped: true convolution), rank-1 factorisation when possible
  float kf[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b)
      kf[a][b] = (a < kh && b < kw) ? __ldg(filt + (kh - 1 - a) * kw + (kw - 1 - b)) : 0.f;
  float ku[4], kv[4];
  {
    int a0 = 0, b0 = 0;
    float big = 0.f;
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b)
        if (fabsf(kf[a][b]) > big) { big = fabsf(kf[a][b]); a0 = a; b0 = b; }
    float piv = 1.f;
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b)
        if (a == a0 && b == b0) piv = kf[a][b];
    const float inv = big > 0.f ? 1.f / piv : 0.f;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      float col = 0.f;
#pragma unroll
      for (int b = 0; b < 4; ++b) if (b == b0) col = kf[a][b];
      ku[a] = col * inv;
    }
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      float row = 0.f;
#pragma unroll
      for (int a = 0; a < 4; ++a) if (a == a0) row = kf[a][b];
      kv[b] = row;
    }
  }

  // per-thread channel constants (a thread keeps its 4 channels for the whole segment)
  float4 bq = make_float4(0.f, 0.f, 0.f, 0.f), rq = make_float4(1.f, 1.f, 1.f, 1.f);
  float nw = 0.f;
  if (FUSED) {
    if (bias) bq = __ldg(reinterpret_cast<const float4*>(bias + c0) + cq);
    if (row_scale) rq = __ldg(reinterpret_cast<const float4*>(row_scale + static_cast<int64_t>(n) * p.c + c0) + cq);
    nw = noise ? (noise_weight ? __ldg(noise_weight) : 1.f) : 0.f;
  }
  // lrelu(t)*gain == max(T, T*slope) with T = gain*t when gain > 0 and 0 <= slope <= 1: the gain is folded into the
  // row scale, the bias and the noise weight, the row scale into the vertical taps (2 epilogue ops per output)
  const bool fast = FUSED && p.gain > 0.f && ((p.act == 3 && p.alpha >= 0.f && p.alpha <= 1.f) || p.act == 1);
  const float neg = (p.act == 3) ? p.alpha : 1.f;
  if (fast) {
    rq.x *= p.gain; rq.y *= p.gain; rq.z *= p.gain; rq.w *= p.gain;
    bq.x *= p.gain; bq.y *= p.gain; bq.z *= p.gain; bq.w *= p.gain;
    nw *= p.gain;
  }
  float4 kur[4];                           // vertical taps x row scale (separable fused path)
#pragma unroll
  for (int a = 0; a < 4; ++a) kur[a] = make_float4(ku[a] * rq.x, ku[a] * rq.y, ku[a] * rq.z, ku[a] * rq.w);
  const int xo = x_out0 + 2 * xg;         // first of this thread's two output columns
  const bool ok0 = xo < p.out_w, ok1 = xo + 1 < p.out_w;
  // noise of the rows a stage completes is fetched one stage ahead (its latency hides behind the previous stage)
  float nzn[kRY][2];
#pragma unroll
  for (int rr = 0; rr < kRY; ++rr) nzn[rr][0] = nzn[rr][1] = 0.f;
  auto fetch_noise = [&](int it_) {
#pragma unroll
    for (int rr = 0; rr < kRY; ++rr) {
      const int ro = it_ * kRY + rr - 3;
      nzn[rr][0] = nzn[rr][1] = 0.f;
      if (ro >= 0 && ro < rows_out) {
        const float* np_ = noise + (static_cast<int64_t>(n) * p.out_h + oy0 + ro) * p.out_w + xo;
        if (ok0) nzn[rr][0] = __ldg(np_);
        if (ok1) nzn[rr][1] = __ldg(np_ + 1);
      }
    }
  };
  if (FUSED && noise) fetch_noise(0);

  // window: sep -> horizontal results hw[4 rows][2 cols] (float4 over channels); else raw inputs rw[4 rows][5 cols]
  float4 hw[SEP ? 4 : 1][2];
  float4 rw[SEP ? 1 : 4][5];
  int r_in = 0;                            // input rows consumed so far (relative to iy0)
  for (int it = 0; it < n_stage_iters; ++it) {
    const int stage = it % kNS;
    if (tid == 0) {
      const int nxt = it + kNS - 1;
      if (nxt < n_stage_iters) {
        const int ns = nxt % kNS;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx(&full_bar[ns], kStageFloats * 4);
        tma_load_4d(tiles + ns * kStageFloats, &tmap, c0, ix0, iy0 + nxt * kRY, n, &full_bar[ns]);
      }
    }
    float nzc[kRY][2];
#pragma unroll
    for (int rr = 0; rr < kRY; ++rr) { nzc[rr][0] = nzn[rr][0]; nzc[rr][1] = nzn[rr][1]; }
    if (FUSED && noise && it + 1 < n_stage_iters) fetch_noise(it + 1);
    mbar_wait(&full_bar[stage], static_cast<uint32_t>((it / kNS) & 1));
    const float* st = tiles + stage * kStageFloats;
#pragma unroll
    for (int rr = 0; rr < kRY; ++rr, ++r_in) {
      // 5 input pixels (columns 2xg .. 2xg+4 of the tile) x 4 channels of this thread
      const float4* rowp = reinterpret_cast<const float4*>(st + (rr * kTileW + 2 * xg) * kCB) + cq;
      float4 q[5];
#pragma unroll
      for (int i = 0; i < 5; ++i) q[i] = rowp[i * (kCB / 4)];
      if constexpr (SEP) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          float4 h;
          h.x = fmaf(kv[3], q[j + 3].x, fmaf(kv[2], q[j + 2].x, fmaf(kv[1], q[j + 1].x, kv[0] * q[j].x)));
          h.y = fmaf(kv[3], q[j + 3].y, fmaf(kv[2], q[j + 2].y, fmaf(kv[1], q[j + 1].y, kv[0] * q[j].y)));
          h.z = fmaf(kv[3], q[j + 3].z, fmaf(kv[2], q[j + 2].z, fmaf(kv[1], q[j + 1].z, kv[0] * q[j].z)));
          h.w = fmaf(kv[3], q[j + 3].w, fmaf(kv[2], q[j + 2].w, fmaf(kv[1], q[j + 1].w, kv[0] * q[j].w)));
          hw[rr][j] = h;                   // kRY == 4: slot rr == r_in & 3
        }
      } else {
#pragma unroll
        for (int i = 0; i < 5; ++i) rw[rr][i] = q[i];
      }
      const int ro = r_in - 3;             // output row (relative to oy0) completed by this input row
      if (ro >= 0 && ro < rows_out) {
        const int oy = oy0 + ro;
        float4 acc[2];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          float4 a4 = make_float4(0.f, 0.f, 0.f, 0.f);
          if constexpr (SEP) {
            if (FUSED) {                   // accumulate straight into row_scale*t + bias + noise (all x gain if `fast`)
              const float nzj = nw * nzc[rr][j];
              a4 = make_float4(bq.x + nzj, bq.y + nzj, bq.z + nzj, bq.w + nzj);
            }
#pragma unroll
            for (int a = 0; a < 4; ++a) {
              const float4 h = hw[(rr + 1 + a) & 3][j];   // rows r_in-3 .. r_in in order
              if (FUSED) {
                a4.x = fmaf(kur[a].x, h.x, a4.x); a4.y = fmaf(kur[a].y, h.y, a4.y);
                a4.z = fmaf(kur[a].z, h.z, a4.z); a4.w = fmaf(kur[a].w, h.w, a4.w);
              } else {
                a4.x = fmaf(ku[a], h.x, a4.x); a4.y = fmaf(ku[a], h.y, a4.y);
                a4.z = fmaf(ku[a], h.z, a4.z); a4.w = fmaf(ku[a], h.w, a4.w);
              }
            }
          } else {
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
              for (int b = 0; b < 4; ++b) {
                const float4 v = rw[(rr + 1 + a) & 3][j + b];
                a4.x = fmaf(kf[a][b], v.x, a4.x); a4.y = fmaf(kf[a][b], v.y, a4.y);
                a4.z = fmaf(kf[a][b], v.z, a4.z); a4.w = fmaf(kf[a][b], v.w, a4.w);
              }
            if (FUSED) {
              const float nzj = nw * nzc[rr][j];
              a4.x = fmaf(a4.x, rq.x, bq.x + nzj); a4.y = fmaf(a4.y, rq.y, bq.y + nzj);
              a4.z = fmaf(a4.z, rq.z, bq.z + nzj); a4.w = fmaf(a4.w, rq.w, bq.w + nzj);
            }
          }
          if (FUSED) {
            if (fast) {
              a4.x = fmaxf(a4.x, a4.x * neg); a4.y = fmaxf(a4.y, a4.y * neg);
              a4.z = fmaxf(a4.z, a4.z * neg); a4.w = fmaxf(a4.w, a4.w * neg);
            } else {
              a4.x = (a4.x > 0.f ? a4.x : a4.x * neg) * p.gain; a4.y = (a4.y > 0.f ? a4.y : a4.y * neg) * p.gain;
              a4.z = (a4.z > 0.f ? a4.z : a4.z * neg) * p.gain; a4.w = (a4.w > 0.f ? a4.w : a4.w * neg) * p.gain;
            }
          }
          acc[j] = a4;
        }
        float* op = out + (((static_cast<int64_t>(n) * p.out_h + oy) * p.out_w + xo) * p.c + c0) + cq * 4;
        if (ok0) *reinterpret_cast<float4*>(op) = acc[0];
        if (ok1) *reinterpret_cast<float4*>(op + p.c) = acc[1];
      }
    }
    __syncthreads();   // the stage is free for the producer
  }
}

// ---- host: tensor map through the driver entry point (no link-time libcuda dependency)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

inline int grid1(int64_t total, int per_cta) {
  int64_t g = (total + per_cta - 1) / per_cta;
  return static_cast<int>(g > 0 ? g : 1);
}

}  // namespace
}  // namespace gg

using namespace gg;

extern "C" {

int gg_noise_bias_act_nhwc(float* out, const float* x, const float* noise, const float* noise_weight, const float* bias,
                           const float* row_scale, float alpha, float scale, int64_t N, int C, int64_t HW, void* stream) {
  if (N < 0 || C < 0 || HW < 0) return fail(GG_ERR_BAD_ARG, "noise_bias_act_nhwc: negative size");
  const int64_t numel = N * HW * C;
  if (numel == 0) return GG_OK;
  if (C % 4 != 0) return fail(GG_ERR_UNSUPPORTED, "noise_bias_act_nhwc: C must be a multiple of 4");
  if (!out || !x) return fail(GG_ERR_BAD_ARG, "noise_bias_act_nhwc: null tensor");
  const int64_t n_vec = numel / 4;
  const int64_t grid = (n_vec + 4 * kT - 1) / (4 * kT);
  if (grid > 0x7fffffffLL) return fail(GG_ERR_BAD_ARG, "noise_bias_act_nhwc: tensor too large");
  noise_bias_act_nhwc_kernel<<<static_cast<unsigned>(grid), kT, 0, static_cast<cudaStream_t>(stream)>>>(
      out, x, noise, noise_weight, bias, row_scale, alpha, scale, n_vec, C / 4, HW);
  GG_CHECK_LAUNCH("noise_bias_act_nhwc launch");
  return GG_OK;
}

// Pixels per CTA: enough CTAs to fill the machine ~8x over, at least 4 trips of the CTA's pixel lanes each.
static int64_t rowwise_chunk(int64_t N, int C, int64_t HW) {
  const int c4 = C / 4;
  const int lanes_p = kT / c4 > 0 ? kT / c4 : 1;
  const int64_t target = 8LL * sm_count();
  int64_t k = (target + N - 1) / N;
  const int64_t kmax = (HW + 4 * lanes_p - 1) / (4 * lanes_p);
  if (k > kmax) k = kmax;
  if (k < 1) k = 1;
  return (HW + k - 1) / k;
}

int64_t gg_nhwc_rowwise_workspace(int64_t N, int C, int64_t HW) {
  if (N <= 0 || C <= 0 || HW <= 0 || C % 4 != 0) return 0;
  const int64_t chunk = rowwise_chunk(N, C, HW);
  return N * ((HW + chunk - 1) / chunk) * C * static_cast<int64_t>(sizeof(float));
}

static int launch_rowwise(int mode, float* out, float* dst, void* workspace, const float* x, const float* y, const float* s,
                          float alpha, float gain, int64_t N, int C, int64_t HW, bool per_sample, void* stream) {
  if (N < 0 || C < 0 || HW < 0) return fail(GG_ERR_BAD_ARG, "nhwc rowwise: negative size");
  if (N * HW * C == 0) return GG_OK;
  if (C % 4 != 0 || C > 1024) return fail(GG_ERR_UNSUPPORTED, "nhwc rowwise: C must be a multiple of 4 and <= 1024");
  if (!out || !x) return fail(GG_ERR_BAD_ARG, "nhwc rowwise: null tensor");
  if (dst && !workspace) return fail(GG_ERR_BAD_ARG, "nhwc rowwise: reduction needs a workspace");
  const int c4 = C / 4;
  const int64_t chunk64 = rowwise_chunk(N, C, HW);
  if (chunk64 > 0x7fffffffLL) return fail(GG_ERR_BAD_ARG, "nhwc rowwise: plane too large");
  const int chunk = static_cast<int>(chunk64);
  const int K = static_cast<int>((HW + chunk - 1) / chunk);
  const int64_t grid = N * K;
  if (grid > 0x7fffffffLL) return fail(GG_ERR_BAD_ARG, "nhwc rowwise: too many CTAs");
  const int lanes_p = kT / c4 > 0 ? kT / c4 : 1;
  const size_t smem = static_cast<size_t>(lanes_p) * C * sizeof(float);
  float* partial = dst ? static_cast<float*>(workspace) : nullptr;
  auto st = static_cast<cudaStream_t>(stream);
  if (mode == 0)
    rowwise_nhwc_kernel<0><<<static_cast<unsigned>(grid), kT, smem, st>>>(out, partial, x, dst ? y : nullptr, s, alpha, gain,
                                                                         c4, HW, chunk, K);
  else
    rowwise_nhwc_kernel<1><<<static_cast<unsigned>(grid), kT, smem, st>>>(out, partial, x, y, s, alpha, gain, c4, HW, chunk, K);
  GG_CHECK_LAUNCH("nhwc rowwise launch");
  if (dst) {
    const int64_t rows = per_sample ? N : 1;
    const int kk = per_sample ? K : static_cast<int>(N * K);
    nhwc_finish_kernel<<<static_cast<unsigned>(rows * ((C + 31) / 32)), dim3(32, 32), 0, st>>>(dst, partial, rows, kk, C);
    GG_CHECK_LAUNCH("nhwc finish launch");
  }
  return GG_OK;
}

int gg_channel_scale_nhwc(float* out, float* row_dot, void* workspace, const float* x, const float* y, const float* s,
                          int64_t N, int C, int64_t HW, void* stream) {
  if (!s) return fail(GG_ERR_BAD_ARG, "channel_scale_nhwc: null scale");
  if (row_dot && !y) return fail(GG_ERR_BAD_ARG, "channel_scale_nhwc: row_dot needs y");
  return launch_rowwise(0, out, row_dot, workspace, x, y, s, 0.f, 1.f, N, C, HW, true, stream);
}

int gg_bias_act_backward_nhwc(float* gx, float* grad_bias, void* workspace, const float* g, const float* out_saved,
                              float alpha, float scale, int64_t N, int C, int64_t HW, void* stream) {
  if (!out_saved) return fail(GG_ERR_BAD_ARG, "bias_act_backward_nhwc: null saved output");
  return launch_rowwise(1, gx, grad_bias, workspace, g, out_saved, nullptr, alpha, scale, N, C, HW, false, stream);
}

int64_t gg_to_rgb_nhwc_workspace(int64_t N, int C, int64_t HW) { return 3 * gg_nhwc_rowwise_workspace(N, C, HW); }

int gg_to_rgb_nhwc_forward(float* out, const float* x, const float* wm, const float* bias, const float* skip, int64_t N,
                           int C, int64_t HW, void* stream) {
  if (N < 0 || C < 0 || HW < 0) return fail(GG_ERR_BAD_ARG, "to_rgb_nhwc: negative size");
  if (N * HW == 0) return GG_OK;
  if (C < 32 || C % 32 != 0 || C > 1024) return fail(GG_ERR_UNSUPPORTED, "to_rgb_nhwc: C must be a multiple of 32, <= 1024");
  if (!out || !x || !wm) return fail(GG_ERR_BAD_ARG, "to_rgb_nhwc: null tensor");
  // pixels per CTA: a multiple of the 128 pixels one trip covers, ~8 CTAs per SM over the whole batch
  const int64_t target = 8LL * sm_count();
  int64_t k = (target + N - 1) / N;
  const int64_t kmax = (HW + 127) / 128;
  if (k > kmax) k = kmax;
  if (k < 1) k = 1;
  int64_t chunk = ((HW + k - 1) / k + 127) / 128 * 128;
  if (chunk > 0x7fffff00LL) return fail(GG_ERR_BAD_ARG, "to_rgb_nhwc: plane too large");
  const int K = static_cast<int>((HW + chunk - 1) / chunk);
  const int64_t grid = N * K;
  if (grid > 0x7fffffffLL) return fail(GG_ERR_BAD_ARG, "to_rgb_nhwc: too many CTAs");
  to_rgb_nhwc_fwd_kernel<<<static_cast<unsigned>(grid), kT, static_cast<size_t>(3) * C * sizeof(float),
                           static_cast<cudaStream_t>(stream)>>>(out, x, wm, bias, skip, C / 4, HW, static_cast<int>(chunk), K);
  GG_CHECK_LAUNCH("to_rgb_nhwc forward launch");
  return GG_OK;
}

int gg_to_rgb_nhwc_backward(float* gx, float* gwm, void* workspace, const float* g, const float* x, const float* wm,
                            int64_t N, int C, int64_t HW, void* stream) {
  if (N < 0 || C < 0 || HW < 0) return fail(GG_ERR_BAD_ARG, "to_rgb_nhwc backward: negative size");
  if (N * HW == 0 || C == 0) return GG_OK;
  if (C % 4 != 0 || C > 1024) return fail(GG_ERR_UNSUPPORTED, "to_rgb_nhwc backward: C must be a multiple of 4, <= 1024");
  if (!g || !wm || (!gx && !gwm)) return fail(GG_ERR_BAD_ARG, "to_rgb_nhwc backward: null tensor");
  if (gwm && (!x || !workspace)) return fail(GG_ERR_BAD_ARG, "to_rgb_nhwc backward: gwm needs x and a workspace");
  const int c4 = C / 4;
  const int64_t chunk64 = rowwise_chunk(N, C, HW);
  if (chunk64 > 0x7fffffffLL) return fail(GG_ERR_BAD_ARG, "to_rgb_nhwc backward: plane too large");
  const int chunk = static_cast<int>(chunk64);
  const int K = static_cast<int>((HW + chunk - 1) / chunk);
  const int64_t grid = N * K;
  if (grid > 0x7fffffffLL) return fail(GG_ERR_BAD_ARG, "to_rgb_nhwc backward: too many CTAs");
  const int lanes_p = kT / c4 > 0 ? kT / c4 : 1;
  const size_t smem = static_cast<size_t>(lanes_p) * 3 * C * sizeof(float);
  auto st = static_cast<cudaStream_t>(stream);
  float* partial = gwm ? static_cast<float*>(workspace) : nullptr;
  to_rgb_nhwc_bwd_kernel<<<static_cast<unsigned>(grid), kT, smem, st>>>(gx, partial, g, x, wm, c4, HW, chunk, K);
  GG_CHECK_LAUNCH("to_rgb_nhwc backward launch");
  if (gwm) {
    nhwc_finish_kernel<<<static_cast<unsigned>(N * ((3 * C + 31) / 32)), dim3(32, 32), 0, st>>>(gwm, partial, N, K, 3 * C);
    GG_CHECK_LAUNCH("to_rgb_nhwc finish launch");
  }
  return GG_OK;
}

int gg_blur_nhwc(float* out, const float* in, const float* kernel, const float* noise, const float* noise_weight,
                 const float* bias, const float* row_scale, int64_t N, int C, int in_h, int in_w, int kernel_h,
                 int kernel_w, int separable, int pad_x0, int pad_x1, int pad_y0, int pad_y1, int fused, int act,
                 float alpha, float scale, void* stream) {
  if (N < 0 || C < 0 || in_h < 1 || in_w < 1) return fail(GG_ERR_BAD_ARG, "blur_nhwc: bad shape");
  if (kernel_h < 1 || kernel_w < 1 || kernel_h > 4 || kernel_w > 4) return fail(GG_ERR_UNSUPPORTED, "blur_nhwc: filter must be <= 4x4");
  if (C % kCB != 0) return fail(GG_ERR_UNSUPPORTED, "blur_nhwc: C must be a multiple of %d", kCB);
  if (act != 1 && act != 3) return fail(GG_ERR_UNSUPPORTED, "blur_nhwc: act must be 1 or 3");
  const int out_h = in_h + pad_y0 + pad_y1 - kernel_h + 1;
  const int out_w = in_w + pad_x0 + pad_x1 - kernel_w + 1;
  if (out_h < 1 || out_w < 1) return fail(GG_ERR_BAD_ARG, "blur_nhwc: empty output");
  if (N == 0 || C == 0) return GG_OK;
  if (!out || !in || !kernel) return fail(GG_ERR_BAD_ARG, "blur_nhwc: null tensor");
  if (N > 65535) return fail(GG_ERR_UNSUPPORTED, "blur_nhwc: batch > 65535");
  EncodeTiledFn enc = encode_fn();
  if (!enc) return fail(GG_ERR_CUDA, "blur_nhwc: cuTensorMapEncodeTiled is not available from this driver");
  CUtensorMap tmap;
  const cuuint64_t gdim[4] = {static_cast<cuuint64_t>(C), static_cast<cuuint64_t>(in_w), static_cast<cuuint64_t>(in_h),
                              static_cast<cuuint64_t>(N)};
  const cuuint64_t gstr[3] = {static_cast<cuuint64_t>(C) * 4, static_cast<cuuint64_t>(in_w) * C * 4,
                              static_cast<cuuint64_t>(in_h) * in_w * C * 4};
  const cuuint32_t box[4] = {kCB, kTileW, kRY, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  const CUresult r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(in), gdim, gstr, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(GG_ERR_CUDA, "blur_nhwc: cuTensorMapEncodeTiled failed (%d)", static_cast<int>(r));
  BlurNhwcParams p;
  p.n = static_cast<int>(N); p.c = C; p.in_h = in_h; p.in_w = in_w; p.out_h = out_h; p.out_w = out_w;
  p.pad_x0 = pad_x0; p.pad_y0 = pad_y0;
  p.act = act; p.alpha = alpha; p.gain = scale;
  // row segments: enough CTAs to fill the machine twice, at least 16 rows each (3 halo rows per segment)
  const int xblocks = (out_w + kBX - 1) / kBX;
  const int64_t base_ctas = static_cast<int64_t>(xblocks) * (C / kCB) * N;
  int segs = static_cast<int>((2LL * 2 * sm_count() + base_ctas - 1) / base_ctas);
  int seg_rows = (out_h + segs - 1) / segs;
  if (seg_rows < 16) seg_rows = out_h < 16 ? out_h : 16;
  seg_rows = (seg_rows + 3) / 4 * 4;
  p.seg_rows = seg_rows;
  const dim3 grid(static_cast<unsigned>(xblocks * (C / kCB)), static_cast<unsigned>((out_h + seg_rows - 1) / seg_rows),
                  static_cast<unsigned>(N));
  const size_t smem = static_cast<size_t>(kNS) * kStageFloats * sizeof(float);
  static DeviceOnce configured;
  if (configured.needed()) {
    cudaError_t e = cudaSuccess;
    const void* kernels[4] = {reinterpret_cast<const void*>(blur_nhwc_kernel<true, true>),
                              reinterpret_cast<const void*>(blur_nhwc_kernel<true, false>),
                              reinterpret_cast<const void*>(blur_nhwc_kernel<false, true>),
                              reinterpret_cast<const void*>(blur_nhwc_kernel<false, false>)};
    for (int i = 0; i < 4 && e == cudaSuccess; ++i)
      e = cudaFuncSetAttribute(kernels[i], cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return cuda_fail(e, "blur_nhwc smem opt-in");
    configured.done();
  }
  auto st = static_cast<cudaStream_t>(stream);
#define GG_BLUR(F_, S_) blur_nhwc_kernel<F_, S_><<<grid, kT, smem, st>>>(out, tmap, kernel, kernel_h, kernel_w, noise, noise_weight, bias, row_scale, p)
  if (fused) { if (separable) GG_BLUR(true, true); else GG_BLUR(true, false); }
  else { noise = nullptr; noise_weight = nullptr; bias = nullptr; row_scale = nullptr;
         if (separable) GG_BLUR(false, true); else GG_BLUR(false, false); }
#undef GG_BLUR
  GG_CHECK_LAUNCH("blur_nhwc launch");
  return GG_OK;
}

}  // extern "C"
