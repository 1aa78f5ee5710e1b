// upfirdn2d.cu -- FIR resampling (upsample / pad / filter / decimate) for sm_100a.
//
// Two kernels:
//  * fir4_band_kernel: up = down = 1, filter <= 4x4 (every Blur on GANgealing's hot path,
//    reference upfirdn2d_kernel.cu "mode 1/2", :250-262).  A persistent CTA walks (plane, row-band)
//    work items; the input rows of a band are ONE contiguous span of global memory (full-width rows
//    of a dense plane), which a single elected thread moves with a 1-D bulk-TMA copy
//    (cp.async.bulk, SASS UBLKCP) into a 3-stage shared-memory ring signalled by mbarriers.  Rows of
//    these tensors are (2H+1)*4 bytes long, i.e. never 16-byte aligned, so the copy moves the
//    16-byte-aligned superset of the span and the consumers index through the residual shift.
//    Consumers keep a 4x4 register window per output column and slide it down the band, so every
//    input element is read from shared memory 4x (not 16x) and from HBM once (+3 halo rows/band).
//    The epilogue optionally fuses NoiseInjection + bias + leaky-ReLU + gain (+ per-plane scale):
//    the "fused upfirdn2d+bias-act path" -- one read and one write of the activation instead of
//    three of each (blur, noise add, bias-act) in the reference.
//    Algorithmic bytes/launch = s*M*(Hin*Win + Hout*Wout) [+ s*N*Hout*Wout noise + 4*(C+1+16)].
//  * upfirdn2d_generic_kernel: any up/down/pad/filter size (RGB skip up x2, its backward down x2,
//    5x5 test filters ...): one thread per output, polyphase tap walk, fp32 accumulate.
#include "common.cuh"

namespace gg {
namespace {

// ------------------------------------------------------------------------------------------------
// generic gather kernel
// ------------------------------------------------------------------------------------------------
struct GenericParams {
  int in_h, in_w, out_h, out_w;
  int kh, kw;
  int up_x, up_y, down_x, down_y;
  int pad_x0, pad_y0;
};

__device__ __forceinline__ int floordiv(int a, int b) {  // b > 0
  int q = a / b;
  return (q * b > a) ? q - 1 : q;
}
__device__ __forceinline__ int ceildiv_s(int a, int b) { return -floordiv(-a, b); }

template <typename T>
__global__ void __launch_bounds__(256)
upfirdn2d_generic_kernel(T* __restrict__ out, const T* __restrict__ in, const float* __restrict__ taps,
                         GenericParams p, int64_t total) {
  // out[m, oy, ox] = sum_{ky,kx} U[oy*dy + ky, ox*dx + kx] * taps[kh-1-ky][kw-1-kx]
  // U = zero-inserted, padded input: U[y, x] = in[(y-pad_y0)/up_y, (x-pad_x0)/up_x] when divisible & in range.
  for (int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int ox = static_cast<int>(idx % p.out_w);
    const int64_t t = idx / p.out_w;
    const int oy = static_cast<int>(t % p.out_h);
    const int64_t m = t / p.out_h;
    const int y0 = oy * p.down_y - p.pad_y0;  // U-row of ky = 0, in input*up coordinates
    const int x0 = ox * p.down_x - p.pad_x0;
    const int iy_lo = max(ceildiv_s(y0, p.up_y), 0);
    const int iy_hi = min(floordiv(y0 + p.kh - 1, p.up_y), p.in_h - 1);
    const int ix_lo = max(ceildiv_s(x0, p.up_x), 0);
    const int ix_hi = min(floordiv(x0 + p.kw - 1, p.up_x), p.in_w - 1);
    const T* plane = in + m * p.in_h * static_cast<int64_t>(p.in_w);
    float acc = 0.f;
    for (int iy = iy_lo; iy <= iy_hi; ++iy) {
      const int ky = iy * p.up_y - y0;
      const float* trow = taps + (p.kh - 1 - ky) * p.kw;
      const T* irow = plane + static_cast<int64_t>(iy) * p.in_w;
      for (int ix = ix_lo; ix <= ix_hi; ++ix) {
        const int kx = ix * p.up_x - x0;
        acc = fmaf(Cvt<T>::to_f(irow[ix]), __ldg(trow + (p.kw - 1 - kx)), acc);
      }
    }
    out[idx] = Cvt<T>::from_f(acc);
  }
}

// ------------------------------------------------------------------------------------------------
// band kernel (up = down = 1, <= 4x4 taps)
// ------------------------------------------------------------------------------------------------
constexpr int kBandThreads = 256;
constexpr int kBandWarps = kBandThreads / 32;
constexpr int kStages = 3;
constexpr int kRS = 8;  // output rows per warp task (register window slides over kRS + 3 input rows)

struct BandParams {
  int64_t planes;        // M = N*C
  int in_h, in_w, out_h, out_w;
  int pad_x0, pad_y0;
  int band_rows;         // R: output rows per work item (multiple of kRS)
  int bands;             // ceil(out_h / R)
  int stage_elems;       // shared-memory elements per stage (>= (R+3)*in_w + 2*16/sizeof(T))
  // fused epilogue
  int C;                 // channels (plane m -> n = m / C, c = m % C)
  int act;               // 1 linear, 3 lrelu
  float alpha, scale;
};

template <typename T>
struct Span {            // contiguous input span of one work item
  const T* src;          // 16-byte aligned start
  uint32_t bytes;        // multiple of 16 (0: band sees only padding)
  int shift;             // elements between src and the first needed element
  int iy_lo;             // first staged input row
};

template <typename T>
__device__ __forceinline__ Span<T> item_span(const T* in, const BandParams& p, int64_t item) {
  Span<T> s;
  const int64_t m = item / p.bands;
  const int band = static_cast<int>(item - m * p.bands);
  const int oy0 = band * p.band_rows;
  const int rows = min(p.band_rows, p.out_h - oy0);
  const int lo = max(oy0 - p.pad_y0, 0);
  const int hi = min(oy0 + rows - 1 + 3 - p.pad_y0, p.in_h - 1);
  s.iy_lo = lo;
  if (hi < lo) {
    s.src = in; s.bytes = 0; s.shift = 0;
    return s;
  }
  const T* first = in + (m * p.in_h + lo) * static_cast<int64_t>(p.in_w);
  const T* last = in + (m * p.in_h + hi + 1) * static_cast<int64_t>(p.in_w);  // one past
  const uintptr_t a0 = reinterpret_cast<uintptr_t>(first) & ~static_cast<uintptr_t>(15);
  const uintptr_t a1 = (reinterpret_cast<uintptr_t>(last) + 15) & ~static_cast<uintptr_t>(15);
  s.src = reinterpret_cast<const T*>(a0);
  s.bytes = static_cast<uint32_t>(a1 - a0);
  s.shift = static_cast<int>((reinterpret_cast<uintptr_t>(first) - a0) / sizeof(T));
  return s;
}

// CO = output columns per lane (interleaved by 32: conflict-free shared-memory reads at any shift)
template <typename T, int CO, bool FUSED>
__global__ void __launch_bounds__(kBandThreads, 2)
fir4_band_kernel(T* __restrict__ out, const T* __restrict__ in, const float* __restrict__ filt, int kh,
                 int kw, const T* __restrict__ noise, const float* __restrict__ noise_weight,
                 const float* __restrict__ bias, const float* __restrict__ row_scale, BandParams p,
                 int64_t n_items) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ uint64_t full_bar[kStages];
  T* stage_base = reinterpret_cast<T*>(smem_raw);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < kStages; ++s) mbar_init(&full_bar[s], 1);
    mbar_fence_init();
  }
  __syncthreads();

  // flipped 4x4 taps in registers: kf[a][b] multiplies input (oy + a - pad_y0, ox + b - pad_x0)
  float kf[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b)
      kf[a][b] = (a < kh && b < kw) ? __ldg(filt + (kh - 1 - a) * kw + (kw - 1 - b)) : 0.f;  // flipped: true convolution

  float nw = 0.f;
  if (FUSED) nw = noise ? (noise_weight ? __ldg(noise_weight) : 1.f) : 0.f;

  const int64_t stride = gridDim.x;
  // prologue: fill kStages-1 stages
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < kStages - 1; ++s) {
      const int64_t it = blockIdx.x + s * stride;
      if (it < n_items) {
        const Span<T> sp = item_span(in, p, it);
        if (sp.bytes) {
          mbar_expect_tx(&full_bar[s], sp.bytes);
          tma_bulk_g2s(stage_base + static_cast<int64_t>(s) * p.stage_elems, sp.src, sp.bytes, &full_bar[s]);
        }
      }
    }
  }

  const int strips_x = (p.out_w + 32 * CO - 1) / (32 * CO);

  uint32_t phase_bits = 0;  // bit s: parity the next wait on stage s must observe
  int k = 0;                // local iteration counter
  for (int64_t item = blockIdx.x; item < n_items; item += stride, ++k) {
    const int stage = k % kStages;
    // prefetch the item kStages-1 ahead into the stage freed by the previous iteration
    if (tid == 0) {
      const int64_t nxt = item + (kStages - 1) * stride;
      if (nxt < n_items) {
        const int ns = (k + kStages - 1) % kStages;
        const Span<T> sp = item_span(in, p, nxt);
        if (sp.bytes) {
          mbar_expect_tx(&full_bar[ns], sp.bytes);
          tma_bulk_g2s(stage_base + static_cast<int64_t>(ns) * p.stage_elems, sp.src, sp.bytes, &full_bar[ns]);
        }
      }
    }
    const Span<T> sp = item_span(in, p, item);
    const int64_t m = item / p.bands;
    const int band = static_cast<int>(item - m * p.bands);
    const int oy0 = band * p.band_rows;
    const int rows = min(p.band_rows, p.out_h - oy0);
    if (sp.bytes) {  // a padding-only band issues no transfer, so its stage's phase does not advance
      mbar_wait(&full_bar[stage], (phase_bits >> stage) & 1u);
      phase_bits ^= 1u << stage;
    }
    const T* tile = stage_base + static_cast<int64_t>(stage) * p.stage_elems + sp.shift;

    float rs = 1.f, bc = 0.f;
    int64_t n = 0;
    if (FUSED) {
      n = m / p.C;
      const int c = static_cast<int>(m - n * p.C);
      if (row_scale) rs = __ldg(row_scale + m);
      if (bias) bc = __ldg(bias + c);
    }

    const int strips_y = (rows + kRS - 1) / kRS;
    const int n_tasks = strips_x * strips_y;
    for (int task = warp; task < n_tasks; task += kBandWarps) {
      const int sy = task / strips_x;
      const int sx = task - sy * strips_x;
      const int oys = oy0 + sy * kRS;                 // first output row of the strip
      const int nrow = min(kRS, oy0 + rows - oys);    // valid output rows in the strip
      const int iys = oys - p.pad_y0;                 // input row feeding tap row a = 0 of output oys

      int col[CO];        // tile column of tap b = 0
      unsigned msk[CO];   // bit b: tap column valid
      bool cok[CO];
#pragma unroll
      for (int j = 0; j < CO; ++j) {
        const int ox = sx * 32 * CO + lane + 32 * j;
        cok[j] = ox < p.out_w;
        col[j] = ox - p.pad_x0;
        unsigned mk = 0;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const int ix = col[j] + b;
          if (cok[j] && ix >= 0 && ix < p.in_w) mk |= 1u << b;
        }
        msk[j] = mk;
      }

      float win[4][CO][4];
#pragma unroll
      for (int r = 0; r < kRS + 3; ++r) {
        // load input row iys + r into window slot r & 3
        const int iy = iys + r;
        const bool row_ok = (iy >= 0) && (iy < p.in_h) && (r < nrow + 3);
        const T* trow = tile + static_cast<int64_t>(iy - sp.iy_lo) * p.in_w;
#pragma unroll
        for (int j = 0; j < CO; ++j)
#pragma unroll
          for (int b = 0; b < 4; ++b)
            win[r & 3][j][b] = (row_ok && ((msk[j] >> b) & 1u)) ? Cvt<T>::to_f(trow[col[j] + b]) : 0.f;
        if (r >= 3) {
          const int ro = r - 3;  // output row within the strip
          if (ro < nrow) {
            const int oy = oys + ro;
            T* orow = out + (m * p.out_h + oy) * static_cast<int64_t>(p.out_w);
            const T* nrowp = nullptr;
            if (FUSED && noise) nrowp = noise + (n * p.out_h + oy) * static_cast<int64_t>(p.out_w);
#pragma unroll
            for (int j = 0; j < CO; ++j) {
              float acc = 0.f;
#pragma unroll
              for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) acc = fmaf(win[(ro + a) & 3][j][b], kf[a][b], acc);
              if (cok[j]) {
                const int ox = sx * 32 * CO + lane + 32 * j;
                if (FUSED) {
                  float t = acc * rs;
                  if (nrowp) t = t + nw * Cvt<T>::to_f(nrowp[ox]);
                  t += bc;
                  if (p.act == 3) t = t > 0.f ? t : t * p.alpha;
                  acc = t * p.scale;
                }
                orow[ox] = Cvt<T>::from_f(acc);
              }
            }
          }
        }
      }
    }
    __syncthreads();  // every warp is done with `stage` before it is refilled next iteration
  }
}

inline int dtype_size(int dtype) { return dtype == GG_F32 ? 4 : 2; }

// host-side geometry for the band kernel; returns false when the shape does not fit shared memory
struct BandPlan {
  BandParams p;
  int co;
  size_t smem_bytes;
  int64_t n_items;
  int grid;
};

inline bool plan_band(int dtype, int64_t planes, int in_h, int in_w, int out_h, int out_w, int pad_x0,
                      int pad_y0, BandPlan* plan) {
  const int es = dtype_size(dtype);
  const int slack = 32 / es;  // alignment shift (<16 B) + tail round-up (<16 B)
  const int co = out_w > 64 ? 4 : (out_w > 32 ? 2 : 1);
  // warp tasks per band: strips_x * R/kRS; aim for >= 8 tasks and <= ~40 KB per stage
  const int strips_x = (out_w + 32 * co - 1) / (32 * co);
  int r = kRS * ((kBandWarps + strips_x - 1) / strips_x);
  const int64_t budget = 40 * 1024;
  while (r > kRS && static_cast<int64_t>(r + 3) * in_w * es > budget) r -= kRS;
  // grow small-plane bands up to the budget (fewer items, less halo)
  while (r < out_h && static_cast<int64_t>(r + kRS + 3) * in_w * es <= budget / 2) r += kRS;
  const int out_rounded = ((out_h + kRS - 1) / kRS) * kRS;
  if (r > out_rounded) r = out_rounded;
  if (r < kRS) r = kRS;
  const int64_t stage_elems = (static_cast<int64_t>(r + 3) * in_w + slack + 15) / 16 * 16;
  const size_t smem = static_cast<size_t>(stage_elems) * es * kStages;
  if (smem > 200 * 1024) return false;
  BandParams& p = plan->p;
  p.planes = planes; p.in_h = in_h; p.in_w = in_w; p.out_h = out_h; p.out_w = out_w;
  p.pad_x0 = pad_x0; p.pad_y0 = pad_y0;
  p.band_rows = r; p.bands = (out_h + r - 1) / r;
  p.stage_elems = static_cast<int>(stage_elems);
  p.C = 1; p.act = 1; p.alpha = 0.f; p.scale = 1.f;
  plan->co = co;
  plan->smem_bytes = smem;
  plan->n_items = planes * p.bands;
  const int ctas_per_sm = smem * 2 <= 220 * 1024 ? 2 : 1;
  const int64_t max_grid = static_cast<int64_t>(sm_count()) * ctas_per_sm;
  plan->grid = static_cast<int>(plan->n_items < max_grid ? plan->n_items : max_grid);
  return true;
}

template <typename T, int CO, bool FUSED>
int launch_band_t(const BandPlan& pl, void* out, const void* in, const float* filt, int kh, int kw, const void* noise,
                  const float* nw, const float* bias, const float* row_scale, cudaStream_t st) {
  auto kern = fir4_band_kernel<T, CO, FUSED>;
  static thread_local size_t configured = 0;  // per instantiation, per thread: max smem opted in so far
  if (pl.smem_bytes > configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(200 * 1024));
    if (e != cudaSuccess) return cuda_fail(e, "fir4_band smem opt-in");
    configured = 200 * 1024;
  }
  kern<<<pl.grid, kBandThreads, pl.smem_bytes, st>>>(
      static_cast<T*>(out), static_cast<const T*>(in), filt, kh, kw, static_cast<const T*>(noise), nw, bias,
      row_scale, pl.p, pl.n_items);
  GG_CHECK_LAUNCH("fir4_band launch");
  return GG_OK;
}

template <typename T, bool FUSED>
int launch_band_co(const BandPlan& pl, void* out, const void* in, const float* filt, int kh, int kw, const void* noise,
                   const float* nw, const float* bias, const float* row_scale, cudaStream_t st) {
  switch (pl.co) {
    case 4: return launch_band_t<T, 4, FUSED>(pl, out, in, filt, kh, kw, noise, nw, bias, row_scale, st);
    case 2: return launch_band_t<T, 2, FUSED>(pl, out, in, filt, kh, kw, noise, nw, bias, row_scale, st);
    default: return launch_band_t<T, 1, FUSED>(pl, out, in, filt, kh, kw, noise, nw, bias, row_scale, st);
  }
}

template <bool FUSED>
int launch_band(int dtype, const BandPlan& pl, void* out, const void* in, const float* filt, int kh, int kw,
                const void* noise, const float* nw, const float* bias, const float* row_scale,
                cudaStream_t st) {
  switch (dtype) {
    case GG_F32: return launch_band_co<float, FUSED>(pl, out, in, filt, kh, kw, noise, nw, bias, row_scale, st);
    case GG_F16: return launch_band_co<__half, FUSED>(pl, out, in, filt, kh, kw, noise, nw, bias, row_scale, st);
    case GG_BF16: return launch_band_co<__nv_bfloat16, FUSED>(pl, out, in, filt, kh, kw, noise, nw, bias, row_scale, st);
    default: return fail(GG_ERR_UNSUPPORTED, "upfirdn2d: dtype %d not supported (f32/f16/bf16)", dtype);
  }
}

template <typename T>
int launch_generic_t(void* out, const void* in, const float* filt, const GenericParams& gp, int64_t total,
                     cudaStream_t st) {
  int64_t grid = (total + 255) / 256;
  const int64_t cap = static_cast<int64_t>(sm_count()) * 32;
  if (grid > cap) grid = cap;
  upfirdn2d_generic_kernel<T><<<static_cast<unsigned>(grid), 256, 0, st>>>(
      static_cast<T*>(out), static_cast<const T*>(in), filt, gp, total);
  GG_CHECK_LAUNCH("upfirdn2d_generic launch");
  return GG_OK;
}

inline int check_common(const char* who, const void* out, const void* in, const float* kernel, int dtype,
                        int64_t major, int in_h, int in_w, int kh, int kw, int up_x, int up_y, int down_x,
                        int down_y, int out_h, int out_w) {
  if (major < 0 || in_h < 0 || in_w < 0) return fail(GG_ERR_BAD_ARG, "%s: negative size", who);
  if (kh < 1 || kw < 1) return fail(GG_ERR_BAD_ARG, "%s: empty filter", who);
  if (up_x < 1 || up_y < 1 || down_x < 1 || down_y < 1) return fail(GG_ERR_BAD_ARG, "%s: up/down must be >= 1", who);
  if (out_h < 1 || out_w < 1) return fail(GG_ERR_BAD_ARG, "%s: output would be empty (%d x %d)", who, out_h, out_w);
  if (dtype != GG_F32 && dtype != GG_F16 && dtype != GG_BF16)
    return fail(GG_ERR_UNSUPPORTED, "%s: dtype %d not supported (f32/f16/bf16)", who, dtype);
  if (major > 0 && (!out || !in || !kernel)) return fail(GG_ERR_BAD_ARG, "%s: null tensor", who);
  return GG_OK;
}

}  // namespace
}  // namespace gg

using namespace gg;

extern "C" {

int gg_upfirdn2d(void* out, const void* in, const float* kernel, int dtype, int64_t major, int in_h,
                 int in_w, int kernel_h, int kernel_w, int up_x, int up_y, int down_x, int down_y,
                 int pad_x0, int pad_x1, int pad_y0, int pad_y1, void* stream) {
  // out size: reference upfirdn2d.py:103-104 / upfirdn2d_kernel.cu:236-239
  const int out_h = (up_y >= 1 && down_y >= 1 && kernel_h >= 1) ? (in_h * up_y + pad_y0 + pad_y1 - kernel_h) / down_y + 1 : 0;
  const int out_w = (up_x >= 1 && down_x >= 1 && kernel_w >= 1) ? (in_w * up_x + pad_x0 + pad_x1 - kernel_w) / down_x + 1 : 0;
  if (in_h * up_y + pad_y0 + pad_y1 - kernel_h < 0 || in_w * up_x + pad_x0 + pad_x1 - kernel_w < 0)
    return fail(GG_ERR_BAD_ARG, "upfirdn2d: filter larger than padded input");
  int rc = check_common("upfirdn2d", out, in, kernel, dtype, major, in_h, in_w, kernel_h, kernel_w, up_x, up_y,
                        down_x, down_y, out_h, out_w);
  if (rc != GG_OK) return rc;
  if (major == 0) return GG_OK;
  auto st = static_cast<cudaStream_t>(stream);
  if (up_x == 1 && up_y == 1 && down_x == 1 && down_y == 1 && kernel_h <= 4 && kernel_w <= 4 && in_h > 0 &&
      in_w > 0) {
    BandPlan pl;
    if (plan_band(dtype, major, in_h, in_w, out_h, out_w, pad_x0, pad_y0, &pl))
      return launch_band<false>(dtype, pl, out, in, kernel, kernel_h, kernel_w, nullptr, nullptr, nullptr,
                                nullptr, st);
  }
  GenericParams gp{in_h, in_w, out_h, out_w, kernel_h, kernel_w, up_x, up_y, down_x, down_y, pad_x0, pad_y0};
  const int64_t total = major * out_h * static_cast<int64_t>(out_w);
  switch (dtype) {
    case GG_F32: return launch_generic_t<float>(out, in, kernel, gp, total, st);
    case GG_F16: return launch_generic_t<__half>(out, in, kernel, gp, total, st);
    default: return launch_generic_t<__nv_bfloat16>(out, in, kernel, gp, total, st);
  }
}

int gg_blur_noise_bias_act(void* out, const void* in, const float* kernel, const void* noise,
                           const float* noise_weight, const float* bias, const float* row_scale,
                           int dtype, int64_t N, int64_t C, int in_h, int in_w, int kernel_h,
                           int kernel_w, int pad_x0, int pad_x1, int pad_y0, int pad_y1, int act,
                           float alpha, float scale, void* stream) {
  if (N < 0 || C < 0) return fail(GG_ERR_BAD_ARG, "blur_noise_bias_act: negative size");
  if (act != 1 && act != 3) return fail(GG_ERR_UNSUPPORTED, "blur_noise_bias_act: act must be 1 or 3");
  if (kernel_h > 4 || kernel_w > 4) return fail(GG_ERR_UNSUPPORTED, "blur_noise_bias_act: filter larger than 4x4");
  const int out_h = in_h + pad_y0 + pad_y1 - kernel_h + 1;
  const int out_w = in_w + pad_x0 + pad_x1 - kernel_w + 1;
  const int64_t major = N * C;
  int rc = check_common("blur_noise_bias_act", out, in, kernel, dtype, major, in_h, in_w, kernel_h, kernel_w,
                        1, 1, 1, 1, out_h, out_w);
  if (rc != GG_OK) return rc;
  if (major == 0) return GG_OK;
  if (in_h == 0 || in_w == 0) return fail(GG_ERR_BAD_ARG, "blur_noise_bias_act: empty input plane");
  BandPlan pl;
  if (!plan_band(dtype, major, in_h, in_w, out_h, out_w, pad_x0, pad_y0, &pl))
    return fail(GG_ERR_UNSUPPORTED, "blur_noise_bias_act: rows of %d elements do not fit the staging ring", in_w);
  pl.p.C = static_cast<int>(C);
  pl.p.act = act;
  pl.p.alpha = alpha;
  pl.p.scale = scale;
  return launch_band<true>(dtype, pl, out, in, kernel, kernel_h, kernel_w, noise, noise_weight, bias,
                           row_scale, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
