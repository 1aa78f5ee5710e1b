// nhwc_reduce.cuh -- second stage of the deterministic two-stage per-channel reductions of the channels-last kernels.
#pragma once
#include "common.cuh"

namespace gg {

// dst[r][c] = sum_k partial[(r*K + k)][c]    (r = sample for channel_scale; a single row for grad_bias)
// CTA = 32 channels x 32 k-lanes: each lane sums every 32nd partial row (4 independent loads per trip), then the
// 32 lane sums are combined through shared memory in a fixed order (deterministic).
static __global__ void __launch_bounds__(1024)
nhwc_finish_kernel(float* __restrict__ dst, const float* __restrict__ partial, int64_t rows, int K, int C, int pitch = 0) {
  if (pitch == 0) pitch = C;     // pitch: floats between consecutive partial rows (> C when the rows hold several sums)
  __shared__ float red[32][33];
  const int cblocks = (C + 31) / 32;
  const int64_t r = blockIdx.x / cblocks;
  const int c = (blockIdx.x - r * cblocks) * 32 + threadIdx.x;
  const int ky = threadIdx.y;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  if (c < C) {
    const float* base = partial + r * K * pitch + c;
    int k = ky;
    for (; k + 96 < K; k += 128) {
      a0 += base[static_cast<int64_t>(k) * pitch];
      a1 += base[static_cast<int64_t>(k + 32) * pitch];
      a2 += base[static_cast<int64_t>(k + 64) * pitch];
      a3 += base[static_cast<int64_t>(k + 96) * pitch];
    }
    for (; k < K; k += 32) a0 += base[static_cast<int64_t>(k) * pitch];
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


}  // namespace gg
