// resample.cu -- BilinearDownsample as ONE kernel (sm_100a), SURVEY.md 8(f) rank 1.
//
// reference: models/spatial_transformers/antialiased_sampling.py:241-256 -- ReflectionPad2d(stride/2), a depthwise
// 1 x 2s convolution with stride (1, s), then a depthwise 2s x 1 convolution with stride (s, 1) (cross-correlation,
// tent taps `kernel_horz` / `kernel_vert`, one copy per channel).  It sits between the generator's output and the
// STN's input on every step (train.py:62, spatial_transformer.py:579-582): three launches and a padded + a
// half-filtered temporary there, one gather here:
//     out[m, oy, ox] = sum_i sum_j kv[c][i] kh[c][j] in[m, R(oy s + i - p), R(ox s + j - p)],  p = s/2, R = reflect
// The backward is the exact adjoint in gather form (deterministic): an input pixel collects from the <= 3 padded
// positions that reflect onto it per axis, each covered by <= 2 output windows.
// Images are 3-channel and small (25 MB at batch 32): HBM/L2-bound streaming, algorithmic bytes 4 M (H W + OH OW).
#include "common.cuh"

namespace gg {
namespace {

constexpr int kT = 256;
constexpr int kMaxTaps = 32;   // stride <= 16

__device__ __forceinline__ int reflect_index(int t, int n) {   // ReflectionPad2d: no edge repeat; requires |pad| < n
  t = t < 0 ? -t : t;
  return t >= n ? 2 * (n - 1) - t : t;
}

__global__ void __launch_bounds__(kT)
tent_down_fwd_kernel(float* __restrict__ out, const float* __restrict__ in, const float* __restrict__ taps_h,
                     const float* __restrict__ taps_v, int C, int in_h, int in_w, int out_h, int out_w, int s,
                     int64_t total) {
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * kT + threadIdx.x;
  if (idx >= total) return;
  const int ox = static_cast<int>(idx % out_w);
  const int64_t r = idx / out_w;
  const int oy = static_cast<int>(r % out_h);
  const int64_t m = r / out_h;
  const int c = static_cast<int>(m % C);
  const int p = s / 2, taps = 2 * s;
  const float* kh = taps_h + c * taps;
  const float* kv = taps_v + c * taps;
  const float* plane = in + m * in_h * static_cast<int64_t>(in_w);
  float acc = 0.f;
  for (int i = 0; i < taps; ++i) {
    const int y = reflect_index(oy * s + i - p, in_h);
    const float* row = plane + static_cast<int64_t>(y) * in_w;
    float h = 0.f;                       // horizontal pass first, like the reference (rounding order)
    for (int j = 0; j < taps; ++j) h = fmaf(__ldg(kh + j), __ldg(row + reflect_index(ox * s + j - p, in_w)), h);
    acc = fmaf(__ldg(kv + i), h, acc);
  }
  out[idx] = acc;
}

// padded positions u (0 <= u < n + 2p) whose reflection is input index y: y + p, p - y (1 <= y <= p),
// 2(n-1) - y + p (n-1-p <= y <= n-2)
__device__ __forceinline__ int padded_positions(int y, int n, int p, int (&u)[3]) {
  int k = 0;
  u[k++] = y + p;
  if (y >= 1 && y <= p) u[k++] = p - y;
  if (y >= n - 1 - p && y <= n - 2) u[k++] = 2 * (n - 1) - y + p;
  return k;
}

__global__ void __launch_bounds__(kT)
tent_down_bwd_kernel(float* __restrict__ gin, const float* __restrict__ gout, const float* __restrict__ taps_h,
                     const float* __restrict__ taps_v, int C, int in_h, int in_w, int out_h, int out_w, int s,
                     int64_t total) {
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * kT + threadIdx.x;
  if (idx >= total) return;
  const int x = static_cast<int>(idx % in_w);
  const int64_t r = idx / in_w;
  const int y = static_cast<int>(r % in_h);
  const int64_t m = r / in_h;
  const int c = static_cast<int>(m % C);
  const int p = s / 2, taps = 2 * s;
  const float* kh = taps_h + c * taps;
  const float* kv = taps_v + c * taps;
  const float* plane = gout + m * out_h * static_cast<int64_t>(out_w);
  int uy[3], ux[3];
  const int ny = padded_positions(y, in_h, p, uy), nx = padded_positions(x, in_w, p, ux);
  float acc = 0.f;
  for (int a = 0; a < ny; ++a) {
    const int u = uy[a];
    int oy_lo = (u - taps + 1 + s - 1);             // ceil((u - 2s + 1) / s) for possibly negative numerators
    oy_lo = oy_lo >= 0 ? oy_lo / s : -((-oy_lo + s - 1) / s);
    oy_lo = max(oy_lo, 0);
    const int oy_hi = min(u / s, out_h - 1);
    for (int oy = oy_lo; oy <= oy_hi; ++oy) {
      const float wv = __ldg(kv + (u - oy * s));
      const float* row = plane + static_cast<int64_t>(oy) * out_w;
      for (int b = 0; b < nx; ++b) {
        const int v = ux[b];
        int ox_lo = (v - taps + 1 + s - 1);
        ox_lo = ox_lo >= 0 ? ox_lo / s : -((-ox_lo + s - 1) / s);
        ox_lo = max(ox_lo, 0);
        const int ox_hi = min(v / s, out_w - 1);
        for (int ox = ox_lo; ox <= ox_hi; ++ox) acc = fmaf(wv * __ldg(kh + (v - ox * s)), __ldg(row + ox), acc);
      }
    }
  }
  gin[idx] = acc;
}

int check_tent(const char* who, int64_t N, int C, int in_h, int in_w, int stride, int* out_h, int* out_w) {
  if (N < 0 || C < 0 || in_h < 0 || in_w < 0) return fail(GG_ERR_BAD_ARG, "%s: negative size", who);
  if (stride < 1 || 2 * stride > kMaxTaps) return fail(GG_ERR_UNSUPPORTED, "%s: stride must be in 1..%d", who, kMaxTaps / 2);
  const int p = stride / 2;
  if (N * C > 0 && (in_h <= p || in_w <= p)) return fail(GG_ERR_BAD_ARG, "%s: reflection padding needs a plane larger than stride/2", who);
  *out_h = (in_h + 2 * p - 2 * stride) / stride + 1;
  *out_w = (in_w + 2 * p - 2 * stride) / stride + 1;
  if (N * C > 0 && (in_h + 2 * p < 2 * stride || in_w + 2 * p < 2 * stride)) return fail(GG_ERR_BAD_ARG, "%s: plane smaller than the filter", who);
  return GG_OK;
}

}  // namespace
}  // namespace gg

using namespace gg;

extern "C" {

int gg_tent_downsample_forward(float* out, const float* in, const float* taps_h, const float* taps_v, int64_t N, int C,
                               int in_h, int in_w, int stride, void* stream) {
  int oh = 0, ow = 0;
  int rc = check_tent("tent_downsample", N, C, in_h, in_w, stride, &oh, &ow);
  if (rc != GG_OK) return rc;
  const int64_t total = N * C * oh * static_cast<int64_t>(ow);
  if (total == 0) return GG_OK;
  if (!out || !in || !taps_h || !taps_v) return fail(GG_ERR_BAD_ARG, "tent_downsample: null tensor");
  const int64_t grid = (total + kT - 1) / kT;
  if (grid > 0x7fffffffLL) return fail(GG_ERR_BAD_ARG, "tent_downsample: tensor too large");
  tent_down_fwd_kernel<<<static_cast<unsigned>(grid), kT, 0, static_cast<cudaStream_t>(stream)>>>(
      out, in, taps_h, taps_v, C, in_h, in_w, oh, ow, stride, total);
  GG_CHECK_LAUNCH("tent_downsample forward launch");
  return GG_OK;
}

int gg_tent_downsample_backward(float* grad_in, const float* grad_out, const float* taps_h, const float* taps_v,
                                int64_t N, int C, int in_h, int in_w, int stride, void* stream) {
  int oh = 0, ow = 0;
  int rc = check_tent("tent_downsample backward", N, C, in_h, in_w, stride, &oh, &ow);
  if (rc != GG_OK) return rc;
  const int64_t total = N * C * in_h * static_cast<int64_t>(in_w);
  if (total == 0) return GG_OK;
  if (!grad_in || !grad_out || !taps_h || !taps_v) return fail(GG_ERR_BAD_ARG, "tent_downsample backward: null tensor");
  const int64_t grid = (total + kT - 1) / kT;
  if (grid > 0x7fffffffLL) return fail(GG_ERR_BAD_ARG, "tent_downsample backward: tensor too large");
  tent_down_bwd_kernel<<<static_cast<unsigned>(grid), kT, 0, static_cast<cudaStream_t>(stream)>>>(
      grad_in, grad_out, taps_h, taps_v, C, in_h, in_w, oh, ow, stride, total);
  GG_CHECK_LAUNCH("tent_downsample backward launch");
  return GG_OK;
}

}  // extern "C"
