/*
 * gg_b200.h -- C ABI of libgg_b200.so: the sm_100a kernels behind GANgealing's op-level hot path.
 *
 * Every entry point replaces one native binding (or one cluster of ATen launches) of the reference
 * wpeebles/gangealing; the reference location is cited on each declaration (paths relative to the
 * reference checkout).  Conventions, identical for all entry points:
 *
 *   - plain pointers and sizes only, no torch types; all pointers are DEVICE pointers unless the
 *     parameter name ends in `_host`;
 *   - the caller owns every buffer (inputs, outputs, workspaces); the library allocates nothing,
 *     keeps no mutable global state and never synchronises the device;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *   - return value: 0 = ok, negative = error (GG_ERR_*); gg_last_error() returns a thread-local
 *     human-readable message for the last failing call on this thread.  Never calls exit();
 *   - re-entrant: safe to call concurrently from the Python main thread and PyTorch's autograd thread;
 *   - tensors are dense ("contiguous") in the layout stated per function; element type is selected
 *     by a gg_dtype code; accumulation is always fp32.
 *   - staged (bulk-TMA) kernels may READ, never write, up to 15 bytes before/after an input buffer so
 *     that transfers are 16-byte aligned; those bytes never influence results.  (Any CUDA allocation
 *     is at least 256-byte granular, so the enclosing 16-byte window is always mapped.)
 */
#ifndef GG_B200_H_
#define GG_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GG_API __attribute__((visibility("default")))

/* element types */
enum { GG_F32 = 0, GG_F16 = 1, GG_BF16 = 2, GG_F64 = 3 };

/* error codes */
enum {
  GG_OK = 0,
  GG_ERR_BAD_ARG = -1,      /* null pointer, negative size, inconsistent shape            */
  GG_ERR_UNSUPPORTED = -2,  /* dtype / mode not implemented by this entry point           */
  GG_ERR_CUDA = -3          /* a CUDA runtime call failed; message holds cudaGetErrorString */
};

/* padding modes of the samplers (torch.nn.functional.grid_sample's padding_mode) */
enum { GG_PAD_ZEROS = 0, GG_PAD_BORDER = 1, GG_PAD_REFLECTION = 2 };

GG_API int gg_version(void);                 /* ABI version, bumped on any signature change */
GG_API const char* gg_last_error(void);      /* thread-local, never NULL */
GG_API int gg_sm_count(void);                /* multiProcessorCount of the current device (cached) */

/* ------------------------------------------------------------------------------------------------
 * fused_bias_act -- replaces `fused.fused_bias_act(input, bias, refer, act, grad, alpha, scale)`
 *   reference: models/stylegan2/op/fused_bias_act.cpp:11-21, fused_bias_act_kernel.cu:18-99
 *   x'   = x + bias[(i / step_b) % size_b]            (bias == NULL: no bias)
 *   act=1: y = x'                       (grad 0/1), 0 (grad 2)
 *   act=3: y = x'   > 0 ? x' : alpha*x' (grad 0)
 *          y = ref  > 0 ? x' : alpha*x' (grad 1; ref = saved forward OUTPUT), 0 (grad 2)
 *   out = y * scale
 *   x/out/ref: `size_x` elements of `dtype`; bias: `size_b` elements of `dtype`.
 * ---------------------------------------------------------------------------------------------- */
GG_API int gg_fused_bias_act(void* out, const void* x, const void* bias, const void* ref, int dtype,
                             int act, int grad, float alpha, float scale, int64_t size_x,
                             int64_t step_b, int64_t size_b, void* stream);

/* ------------------------------------------------------------------------------------------------
 * gg_noise_bias_act -- NoiseInjection + FusedLeakyReLU of a StyledConv in one pass
 *   reference: models/stylegan2/networks.py:291-298 (noise) + :344-350 + op/fused_act.py:52-58
 *   out[n,c,p] = lrelu(x[n,c,p] + noise_weight[0]*noise[n,p] + bias[c], alpha) * scale
 *   x/out: (N, C, HW) `dtype`;  noise: (N, HW) `dtype` or NULL;  noise_weight: 1 fp32 (device) or
 *   NULL (=1);  bias: C fp32 or NULL.
 * ---------------------------------------------------------------------------------------------- */
GG_API int gg_noise_bias_act(void* out, const void* x, const void* noise, const float* noise_weight,
                             const float* bias, int dtype, float alpha, float scale, int64_t N,
                             int64_t C, int64_t HW, void* stream);

/* ------------------------------------------------------------------------------------------------
 * gg_bias_act_backward -- FusedLeakyReLUFunctionBackward in one pass
 *   reference: models/stylegan2/op/fused_act.py:20-38 (kernel call act=3,grad=1 + grad_input.sum(dims))
 *   gx[n,c,p] = (out[n,c,p] > 0 ? g : alpha*g) * scale
 *   grad_bias[c] = sum_{n,p} gx[n,c,p]      (fp32; skipped when grad_bias == NULL)
 *   Deterministic two-stage reduction; `workspace` must hold gg_bias_act_backward_workspace() bytes
 *   (may be NULL when grad_bias is NULL).
 * ---------------------------------------------------------------------------------------------- */
GG_API int64_t gg_bias_act_backward_workspace(int64_t N, int64_t C, int64_t HW);
GG_API int gg_bias_act_backward(void* gx, float* grad_bias, void* workspace, const void* g,
                                const void* out, int dtype, float alpha, float scale, int64_t N,
                                int64_t C, int64_t HW, void* stream);

/* ------------------------------------------------------------------------------------------------
 * gg_upfirdn2d -- replaces `upfirdn2d_op.upfirdn2d(input, kernel, up_x, up_y, down_x, down_y,
 *                                                  pad_x0, pad_x1, pad_y0, pad_y1)`
 *   reference: models/stylegan2/op/upfirdn2d.cpp:12-23, upfirdn2d_kernel.cu:209-369
 *   (semantics: upfirdn2d.py:159-200 -- zero-insert upsample, pad/crop, TRUE convolution with
 *    `kernel`, decimate).  in: (major, in_h, in_w) dense; out: (major, out_h, out_w) with
 *    out_h = (in_h*up_y + pad_y0 + pad_y1 - kernel_h) / down_y + 1 (same for w); the reference's
 *    trailing `minor` dimension is always 1 in GANgealing and is not modelled.
 *   kernel: kernel_h*kernel_w fp32 taps (device), NOT flipped (the library flips, as the reference).
 *   Dispatch: up=down=1 and kernel <= 4x4 -> bulk-TMA staged band kernel; otherwise generic gather.
 * ---------------------------------------------------------------------------------------------- */
GG_API int gg_upfirdn2d(void* out, const void* in, const float* kernel, int dtype, int64_t major,
                        int in_h, int in_w, int kernel_h, int kernel_w, int up_x, int up_y,
                        int down_x, int down_y, int pad_x0, int pad_x1, int pad_y0, int pad_y1,
                        void* stream);

/* ------------------------------------------------------------------------------------------------
 * gg_blur_noise_bias_act -- the fused StyledConv(upsample) tail: Blur -> NoiseInjection ->
 *   FusedLeakyReLU in ONE pass over the activation (the "fused upfirdn2d+bias-act path").
 *   reference: models/stylegan2/networks.py:266 (blur) + :346-348 (noise, activate);
 *              op/upfirdn2d_kernel.cu:107-207 + op/fused_bias_act_kernel.cu:18-49
 *   t = upfirdn2d(in, kernel, up=1, down=1, pad)               (kernel <= 4x4)
 *   out[n,c,y,x] = lrelu(row_scale[n*C+c]*t + noise_weight[0]*noise[n,y,x] + bias[c], alpha) * scale
 *   in: (N*C, in_h, in_w) dtype; out: (N*C, out_h, out_w); noise: (N, out_h, out_w) dtype or NULL;
 *   noise_weight: 1 fp32 or NULL(=1); bias: C fp32 or NULL; row_scale: N*C fp32 or NULL (=1; lets a
 *   caller fold the per-sample demodulation of a weight-shared conv into the tail);
 *   act: 1 linear, 3 leaky-relu.
 * ---------------------------------------------------------------------------------------------- */
GG_API int gg_blur_noise_bias_act(void* out, const void* in, const float* kernel, const void* noise,
                                  const float* noise_weight, const float* bias,
                                  const float* row_scale, int dtype, int64_t N, int64_t C, int in_h,
                                  int in_w, int kernel_h, int kernel_w, int pad_x0, int pad_x1,
                                  int pad_y0, int pad_y1, int act, float alpha, float scale,
                                  void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GG_B200_H_ */
