// modconv.cu -- weight modulation / demodulation of StyleGAN2's ModulatedConv2d (sm_100a).
//
// reference: models/stylegan2/networks.py:233-253
//     weight = scale * W * style[b, i]                          (B, O, I, k, k)
//     demod  = rsqrt(sum_{i,kh,kw} weight^2 + 1e-8)             (B, O)
//     weight = weight * demod                                   -> grouped-conv filters
// i.e. ~6 ATen launches that read/write three (B, O, I, k, k) temporaries.  Here:
//   1. gg_modconv_wsq      Wsq[o, i] = sum_{kh,kw} W[o,i,kh,kw]^2            (once per frozen filter bank)
//   2. gg_modconv_demod    demod[b, o] = rsqrt(scale^2 * sum_i Wsq[o,i] * style[b,i]^2 + eps)
//        -- THE dense contraction of the hot path: [O x I] . [I x B] on the 5th-gen tensor cores
//        (tcgen05.mma kind::tf32, accumulator in TMEM).  Operands are split hi/lo (3 MMAs per k-step) so the
//        result carries ~fp32 accuracy although each MMA rounds its inputs to TF32.
//   3. gg_modconv_modulate out = scale * W * style[b,i] * demod[b,o], written ONCE, directly in the layout the
//        grouped convolution wants ((B*O, I, k, k), or (B*I, O, k, k) for the transposed up-convolution).
// Bytes: the modulate pass writes 4*B*O*I*k*k (its reads of W are L2 hits); FLOPs of the GEMM: 2*B*O*I (tiny).
#include "common.cuh"

namespace gg {
namespace {

// ------------------------------------------------------------------------------------------------ 1. Wsq
__global__ void __launch_bounds__(256)
wsq_kernel(float* __restrict__ wsq, const float* __restrict__ w, int64_t oi, int kk) {
  for (int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; idx < oi;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float* p = w + idx * kk;
    float acc = 0.f;
    for (int k = 0; k < kk; ++k) acc = fmaf(p[k], p[k], acc);
    wsq[idx] = acc;
  }
}

// ------------------------------------------------------------------------------------------------ 2. demod GEMM
// D[o, b] = sum_i Wsq[o, i] * s2[b, i]: M = 128 rows of O per CTA, N = B padded to 16.., K = I in blocks of 32 fp32
// (= one 128-byte swizzle atom).  A and B are K-major in shared memory, SWIZZLE_128B canonical layout:
// row r, 16-byte chunk c  ->  byte offset r*128 + ((c ^ (r & 7)) << 4); 8-row groups are 1024 B apart (SBO).
constexpr int kDemodThreads = 256;   // 2 threads per A row while staging; warps 0-3 read the accumulator
constexpr int kBlockK = 32;      // fp32 elements per k-block (128 B)
constexpr int kUmmaK = 8;        // tf32: 32 B per MMA

__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3ffff) >> 4);            // start address, 16-byte units
  d |= static_cast<uint64_t>(1) << 16;                               // leading byte offset (unused for swizzled K-major)
  d |= static_cast<uint64_t>(1024 >> 4) << 32;                       // stride byte offset: 8 rows * 128 B
  d |= static_cast<uint64_t>(1) << 46;                               // descriptor version (sm_100)
  d |= static_cast<uint64_t>(2) << 61;                               // SWIZZLE_128B
  return d;
}

__device__ __forceinline__ uint32_t make_idesc_tf32(int m, int n) {
  uint32_t d = 0;
  d |= 1u << 4;                          // D format: f32
  d |= 2u << 7;                          // A format: tf32
  d |= 2u << 10;                         // B format: tf32
  // bits 15 / 16: A / B major = 0 (K-major)
  d |= static_cast<uint32_t>(n >> 3) << 17;
  d |= static_cast<uint32_t>(m >> 4) << 24;
  return d;
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}

__device__ __forceinline__ bool mbar_wait_bounded(uint64_t* bar, uint32_t parity) {
  // a wrong descriptor must not hang the GPU: give up after ~20 ms and let the caller flag the failure
  const long long t0 = clock64();
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (!done && clock64() - t0 > 40000000LL) return false;
  }
  return true;
}

__device__ __forceinline__ void split_tf32(float v, float& hi, float& lo) {
  hi = __uint_as_float(__float_as_uint(v) & 0xffffe000u);   // exactly representable in TF32 (10-bit mantissa)
  lo = v - hi;                                              // exact in fp32; itself rounded to TF32 by the MMA
}

// store 4 consecutive k-values (one 16-byte chunk) of row r into a SWIZZLE_128B K-major tile
__device__ __forceinline__ void st_tile_chunk(float* tile, int r, int c, float4 v) {
  const uint32_t off = static_cast<uint32_t>(r) * 128u + static_cast<uint32_t>((c ^ (r & 7)) << 4);
  *reinterpret_cast<float4*>(reinterpret_cast<unsigned char*>(tile) + off) = v;
}

// NA = swizzle atoms (32-float k-blocks) staged per pipeline step: 2 when the B tiles are small enough, so that
// 16 independent 16-byte loads per thread are in flight while the previous step's MMAs run.
// One launch serves a BATCH of layers (the generator's 13 modulated convolutions are known up front: their demodulation
// coefficients are independent small GEMMs, 4 CTAs each -- one 27 us launch per layer was pure latency): a CTA looks up
// its layer by block index; every layer shares the batch size B.
constexpr int kMaxDemodLayers = 32;
struct DemodBatch {
  const float* wsq[kMaxDemodLayers];
  const float* style[kMaxDemodLayers];
  float* out[kMaxDemodLayers];
  float scale2[kMaxDemodLayers];
  int O[kMaxDemodLayers], I[kMaxDemodLayers];
  int first_block[kMaxDemodLayers + 1];
  int layers;
};

template <int NA>
__global__ void __launch_bounds__(kDemodThreads)
demod_umma_kernel(const __grid_constant__ DemodBatch batch, float eps, int B, int n_pad, int tmem_cols) {
  int layer = 0;
  while (layer + 1 < batch.layers && static_cast<int>(blockIdx.x) >= batch.first_block[layer + 1]) ++layer;
  float* __restrict__ demod = batch.out[layer];
  const float* __restrict__ wsq = batch.wsq[layer];
  const float* __restrict__ style = batch.style[layer];
  const float scale2 = batch.scale2[layer];
  const int O = batch.O[layer], I = batch.I[layer];
  extern __shared__ __align__(1024) unsigned char smem[];
  // per atom: A_hi, A_lo (128 x 32 fp32 = 16 KB each), B_hi, B_lo (n_pad x 32 fp32 each)
  unsigned char* base = smem + ((1024u - (smem_u32(smem) & 1023u)) & 1023u);   // SWIZZLE_128B tiles: 1024-byte aligned
  const int a_elems = 128 * kBlockK, b_elems = n_pad * kBlockK;
  float* tiles = reinterpret_cast<float*>(base);
  auto a_hi = [&](int a) { return tiles + a * (2 * a_elems + 2 * b_elems); };
  auto a_lo = [&](int a) { return a_hi(a) + a_elems; };
  auto b_hi = [&](int a) { return a_lo(a) + a_elems; };
  auto b_lo = [&](int a) { return b_hi(a) + b_elems; };
  __shared__ uint64_t mma_bar;
  __shared__ uint32_t tmem_base_smem;
  __shared__ int failed;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int o0 = (static_cast<int>(blockIdx.x) - batch.first_block[layer]) * 128;

  if (tid == 0) {
    mbar_init(&mma_bar, 1);
    mbar_fence_init();
    failed = 0;
  }
  if (warp == 0) {  // TMEM allocation: one warp, power-of-two columns >= 32
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)),
                 "r"(static_cast<uint32_t>(tmem_cols))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = tmem_base_smem;

  const uint32_t idesc = make_idesc_tf32(128, n_pad);
  const bool vec_ok = (I % 4 == 0) && ((reinterpret_cast<uintptr_t>(wsq) & 15) == 0);
  const int a_r = tid & 127;            // A-tile row staged by this thread
  const int a_c0 = (tid >> 7) * 4;      // ... and its first of 4 chunks (two threads share a row)
  const int o_row = o0 + a_r;
  const float* a_row = wsq + static_cast<int64_t>(min(o_row, O - 1)) * I;

  // register staging of this thread's half A row: NA atoms x 4 chunks of 4 floats
  float4 ra[NA][4];
  auto load_a = [&](int kb0) {
#pragma unroll
    for (int a = 0; a < NA; ++a)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int i = (kb0 + a) * kBlockK + (a_c0 + c) * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (o_row < O) {
          if (vec_ok && i + 3 < I) {
            v = __ldg(reinterpret_cast<const float4*>(a_row + i));
          } else {
            if (i + 0 < I) v.x = __ldg(a_row + i + 0);
            if (i + 1 < I) v.y = __ldg(a_row + i + 1);
            if (i + 2 < I) v.z = __ldg(a_row + i + 2);
            if (i + 3 < I) v.w = __ldg(a_row + i + 3);
          }
        }
        ra[a][c] = v;
      }
  };

  uint32_t parity = 0;
  bool ok = true;
  const int k_blocks = (I + kBlockK - 1) / kBlockK;
  load_a(0);
  for (int kb = 0; kb < k_blocks && ok; kb += NA) {
    // ---- registers -> swizzled tiles (A), global -> tiles (B: squared styles, small)
#pragma unroll
    for (int a = 0; a < NA; ++a) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float v[4] = {ra[a][c].x, ra[a][c].y, ra[a][c].z, ra[a][c].w};
        float hi[4], lo[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) split_tf32(v[j], hi[j], lo[j]);
        st_tile_chunk(a_hi(a), a_r, a_c0 + c, make_float4(hi[0], hi[1], hi[2], hi[3]));
        st_tile_chunk(a_lo(a), a_r, a_c0 + c, make_float4(lo[0], lo[1], lo[2], lo[3]));
      }
      for (int rc = tid; rc < n_pad * 8; rc += kDemodThreads) {   // (row, chunk) pairs: coalesced over chunks
        const int r = rc >> 3, c = rc & 7;
        float hi[4], lo[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int i = (kb + a) * kBlockK + c * 4 + j;
          const float sv = (r < B && i < I) ? __ldg(style + static_cast<int64_t>(r) * I + i) : 0.f;
          split_tf32(sv * sv, hi[j], lo[j]);
        }
        st_tile_chunk(b_hi(a), r, c, make_float4(hi[0], hi[1], hi[2], hi[3]));
        st_tile_chunk(b_lo(a), r, c, make_float4(lo[0], lo[1], lo[2], lo[3]));
      }
    }
    // generic-proxy writes -> visible to the tensor-core (async) proxy
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
      for (int a = 0; a < NA; ++a) {
        if (kb + a < k_blocks) {
          const uint64_t dah = make_smem_desc(smem_u32(a_hi(a))), dal = make_smem_desc(smem_u32(a_lo(a)));
          const uint64_t dbh = make_smem_desc(smem_u32(b_hi(a))), dbl = make_smem_desc(smem_u32(b_lo(a)));
#pragma unroll
          for (int k = 0; k < kBlockK / kUmmaK; ++k) {
            const uint64_t adv = static_cast<uint64_t>((k * kUmmaK * 4) >> 4);   // +32 B per k-step inside the atom
            umma_tf32(tmem_d, dah + adv, dbh + adv, idesc, (kb | a | k) ? 1u : 0u);
            umma_tf32(tmem_d, dah + adv, dbl + adv, idesc, 1u);
            umma_tf32(tmem_d, dal + adv, dbh + adv, idesc, 1u);
          }
        }
      }
      // completion of all MMAs issued so far -> mbarrier (implies tcgen05.fence::before_thread_sync)
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mma_bar))
                   : "memory");
    }
    if (kb + NA < k_blocks) load_a(kb + NA);   // next step's loads fly while the tensor core works
    ok = mbar_wait_bounded(&mma_bar, parity);  // tiles may be overwritten / accumulator read after this
    parity ^= 1u;
  }
  if (!ok) failed = 1;
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  __syncthreads();

  // ---- epilogue: warp w (< 4) owns TMEM lanes [32w, 32w+32) = rows o0 + 32w + lane; 8 columns (batch entries) per load
  const int o = o0 + warp * 32 + lane;
  for (int n0 = 0; n0 < (warp < 4 ? n_pad : 0); n0 += 8) {
    uint32_t v[8];
    const uint32_t taddr = tmem_d + (static_cast<uint32_t>(warp * 32) << 16) + static_cast<uint32_t>(n0);
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int b = n0 + j;
      if (o < O && b < B) {
        const float d = failed ? __int_as_float(0x7fc00000) : rsqrtf(fmaf(scale2, __uint_as_float(v[j]), eps));
        demod[static_cast<int64_t>(b) * O + o] = d;
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(static_cast<uint32_t>(tmem_cols))
                 : "memory");
}

// ------------------------------------------------------------------------------------------------ 3. modulate
// flat over the output in 16-byte vectors.  ROW = (b, o) [plain] or (b, i) [transposed]; INNER = I*kk or O*kk.
// src is W in the matching layout: (O, I, kk) plain, (I, O, kk) transposed (pre-transposed once by the caller).
template <bool TRANSPOSED>
__global__ void __launch_bounds__(256)
modulate_kernel(float* __restrict__ out, const float* __restrict__ src, const float* __restrict__ style,
                const float* __restrict__ demod, float scale, int B, int O, int I, int kk, int64_t total_vec) {
  const int rows_per_b = TRANSPOSED ? I : O;
  const int inner = (TRANSPOSED ? O : I) * kk;       // multiple of 4 checked on the host
  const int inner_vec = inner >> 2;
  for (int64_t v = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; v < total_vec;
       v += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t row = v / inner_vec;
    const int e0 = static_cast<int>(v - row * inner_vec) << 2;
    const int b = static_cast<int>(row / rows_per_b);
    const int r = static_cast<int>(row - static_cast<int64_t>(b) * rows_per_b);
    const float4 w = __ldg(reinterpret_cast<const float4*>(src + static_cast<int64_t>(r) * inner + e0));
    const float wv[4] = {w.x, w.y, w.z, w.w};
    float o4[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int q = (e0 + j) / kk;                   // i (plain) or o (transposed)
      float f;
      if (TRANSPOSED) f = __ldg(style + static_cast<int64_t>(b) * I + r) * (demod ? __ldg(demod + static_cast<int64_t>(b) * O + q) : 1.f);
      else            f = __ldg(style + static_cast<int64_t>(b) * I + q) * (demod ? __ldg(demod + static_cast<int64_t>(b) * O + r) : 1.f);
      o4[j] = scale * wv[j] * f;
    }
    st_vec_stream(out + v * 4, *reinterpret_cast<const Vec16<float>*>(o4));
  }
}

inline int grid_cap(int64_t total, int threads, int per_sm) {
  int64_t g = (total + threads - 1) / threads;
  const int64_t cap = static_cast<int64_t>(sm_count()) * per_sm;
  return static_cast<int>(g < cap ? (g > 0 ? g : 1) : cap);
}

}  // namespace
}  // namespace gg

using namespace gg;

extern "C" {

int gg_modconv_wsq(float* wsq, const float* weight, int O, int I, int kk, void* stream) {
  if (O < 0 || I < 0 || kk < 1) return fail(GG_ERR_BAD_ARG, "modconv_wsq: bad shape");
  const int64_t oi = static_cast<int64_t>(O) * I;
  if (oi == 0) return GG_OK;
  if (!wsq || !weight) return fail(GG_ERR_BAD_ARG, "modconv_wsq: null tensor");
  wsq_kernel<<<grid_cap(oi, 256, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(wsq, weight, oi, kk);
  GG_CHECK_LAUNCH("wsq launch");
  return GG_OK;
}

static int launch_demod(const DemodBatch& batch, int blocks, int min_i, float eps, int B, cudaStream_t st) {
  const int n_pad = (B + 15) / 16 * 16;
  int tmem_cols = 32;
  while (tmem_cols < n_pad) tmem_cols <<= 1;
  const int na = (n_pad <= 32 && min_i >= 4 * kBlockK) ? 4 : ((n_pad <= 64 && min_i > kBlockK) ? 2 : 1);
  const size_t smem = static_cast<size_t>(na) * (2 * 128 + 2 * n_pad) * kBlockK * sizeof(float) + 1024;
  static DeviceOnce configured;
  if (configured.needed()) {
    cudaError_t e = cudaFuncSetAttribute(demod_umma_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(demod_umma_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(demod_umma_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 170 * 1024);
    if (e != cudaSuccess) return cuda_fail(e, "modconv_demod smem opt-in");
    configured.done();
  }
  if (na == 4) demod_umma_kernel<4><<<blocks, kDemodThreads, smem, st>>>(batch, eps, B, n_pad, tmem_cols);
  else if (na == 2) demod_umma_kernel<2><<<blocks, kDemodThreads, smem, st>>>(batch, eps, B, n_pad, tmem_cols);
  else demod_umma_kernel<1><<<blocks, kDemodThreads, smem, st>>>(batch, eps, B, n_pad, tmem_cols);
  GG_CHECK_LAUNCH("demod_umma launch");
  return GG_OK;
}

int gg_modconv_demod(float* demod, const float* wsq, const float* style, float scale, float eps, int B, int O, int I,
                     void* stream) {
  if (B < 0 || O < 0 || I < 0) return fail(GG_ERR_BAD_ARG, "modconv_demod: bad shape");
  if (B == 0 || O == 0) return GG_OK;
  if (!demod || !wsq || !style) return fail(GG_ERR_BAD_ARG, "modconv_demod: null tensor");
  if (B > 256) return fail(GG_ERR_UNSUPPORTED, "modconv_demod: batch %d > 256 (split the call)", B);
  DemodBatch batch;
  batch.layers = 1;
  batch.wsq[0] = wsq; batch.style[0] = style; batch.out[0] = demod; batch.scale2[0] = scale * scale;
  batch.O[0] = O; batch.I[0] = I; batch.first_block[0] = 0; batch.first_block[1] = (O + 127) / 128;
  return launch_demod(batch, (O + 127) / 128, I, eps, B, static_cast<cudaStream_t>(stream));
}

int gg_modconv_demod_batched(int layers, float* const* demod, const float* const* wsq, const float* const* style,
                             const float* scale, const int* O, const int* I, float eps, int B, void* stream) {
  if (layers < 0 || B < 0) return fail(GG_ERR_BAD_ARG, "modconv_demod_batched: bad shape");
  if (layers == 0 || B == 0) return GG_OK;
  if (layers > kMaxDemodLayers) return fail(GG_ERR_UNSUPPORTED, "modconv_demod_batched: more than %d layers", kMaxDemodLayers);
  if (!demod || !wsq || !style || !scale || !O || !I) return fail(GG_ERR_BAD_ARG, "modconv_demod_batched: null table");
  if (B > 256) return fail(GG_ERR_UNSUPPORTED, "modconv_demod_batched: batch %d > 256 (split the call)", B);
  DemodBatch batch;
  batch.layers = layers;
  int blocks = 0, min_i = 1 << 30;
  for (int l = 0; l < layers; ++l) {
    if (O[l] < 1 || I[l] < 1) return fail(GG_ERR_BAD_ARG, "modconv_demod_batched: layer %d has an empty shape", l);
    if (!demod[l] || !wsq[l] || !style[l]) return fail(GG_ERR_BAD_ARG, "modconv_demod_batched: layer %d has a null tensor", l);
    batch.wsq[l] = wsq[l]; batch.style[l] = style[l]; batch.out[l] = demod[l]; batch.scale2[l] = scale[l] * scale[l];
    batch.O[l] = O[l]; batch.I[l] = I[l]; batch.first_block[l] = blocks;
    blocks += (O[l] + 127) / 128;
    if (I[l] < min_i) min_i = I[l];
  }
  batch.first_block[layers] = blocks;
  return launch_demod(batch, blocks, min_i, eps, B, static_cast<cudaStream_t>(stream));
}

int gg_modconv_modulate(float* out, const float* weight, const float* style, const float* demod, float scale, int B,
                        int O, int I, int kk, int transposed, void* stream) {
  if (B < 0 || O < 0 || I < 0 || kk < 1) return fail(GG_ERR_BAD_ARG, "modconv_modulate: bad shape");
  const int64_t total = static_cast<int64_t>(B) * O * I * kk;
  if (total == 0) return GG_OK;
  if (!out || !weight || !style) return fail(GG_ERR_BAD_ARG, "modconv_modulate: null tensor");
  const int inner = (transposed ? O : I) * kk;
  if (inner % 4 != 0) return fail(GG_ERR_UNSUPPORTED, "modconv_modulate: inner extent %d is not a multiple of 4", inner);
  const int64_t total_vec = total / 4;
  auto st = static_cast<cudaStream_t>(stream);
  if (transposed)
    modulate_kernel<true><<<grid_cap(total_vec, 256, 16), 256, 0, st>>>(out, weight, style, demod, scale, B, O, I, kk, total_vec);
  else
    modulate_kernel<false><<<grid_cap(total_vec, 256, 16), 256, 0, st>>>(out, weight, style, demod, scale, B, O, I, kk, total_vec);
  GG_CHECK_LAUNCH("modulate launch");
  return GG_OK;
}

}  // extern "C"
