/* pmfb.h — C-ABI of libpmf_b200.so, the B200 (sm_100a) kernels behind the PMF hot path.
 *
 * The reference (ICEORY/PMF) is pure PyTorch and has no FFI of its own; the
 * operations below are the device work that the reference's Python modules dispatch to ATen/cuDNN:
 *   - pc_processor/models/pmf_net.py:10-36   ResidualBasedFusionBlock   -> pmfb_conv_fwd (+ epilogue)
 *   - pc_processor/models/salsanext.py:9-164 SalsaNext conv stacks      -> pmfb_conv_fwd / pmfb_conv_wgrad
 *   - pc_processor/models/pmf_net.py:41-100  ResNet encoder (torchvision)-> pmfb_conv_fwd / pmfb_stem_*
 *   - pc_processor/postproc/knn.py:55-143    KNN.forward                -> pmfb_knn_vote
 *   - pc_processor/dataset/semantic_kitti/parser.py:209-227 +
 *     pc_processor/dataset/perspective_view_loader.py:87-131            -> pmfb_project_scatter
 *
 * Conventions
 *   - plain C, raw device pointers and sizes; no torch types.
 *   - every call only ENQUEUES work on `stream` (a cudaStream_t passed as void*); it never
 *     synchronises the device, never allocates or frees device memory.
 *   - returns 0 on success, a negative pmfb_status otherwise; pmfb_last_error() gives the text
 *     (thread-local).  There is no CPU fallback: without a CUDA device every compute entry fails.
 *   - activations are fp32 NHWC ("channels-last") views: channel stride 1, arbitrary element
 *     strides for n / y / x, so a view may be a channel slice of a wider concat buffer or a
 *     parity sub-grid of a larger image.
 */
#ifndef PMFB_H_
#define PMFB_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PMFB_ABI_VERSION 6
#define PMFB_MAX_TAPS 9

typedef enum {
  PMFB_OK = 0,
  PMFB_ERR_INVALID = -1,   /* bad argument / unsupported geometry */
  PMFB_ERR_CUDA = -2,      /* CUDA runtime / driver error        */
  PMFB_ERR_NO_DEVICE = -3  /* no sm_100 device available          */
} pmfb_status;

/* Operand storage of the implicit-GEMM kernels.  F32: fp32 values (pre-rounded to tf32), kind::tf32 UMMAs.  F16 / BF16:
 * 16-bit "shadow" copies of the same NHWC views (see pmfb_convert16), kind::f16 UMMAs with fp32 accumulation: the same
 * 32 bytes per operand row and instruction cover K = 16 channels instead of 8. */
#define PMFB_DT_F32 0
#define PMFB_DT_F16 1
#define PMFB_DT_BF16 2
#define PMFB_DT_F16_BF16 3 /* A operand (activations) fp16, B operand (weights / output gradients) bf16 */

typedef enum { PMFB_ACT_NONE = 0, PMFB_ACT_RELU = 1, PMFB_ACT_LEAKY = 2, PMFB_ACT_SIGMOID = 3 } pmfb_act;

/* NHWC view, channel stride 1.  ptr addresses (n=0,y=0,x=0,c=0); strides in ELEMENTS. */
typedef struct {
  const float* ptr;
  int64_t sn, sy, sx;
} pmfb_view;

/* Fused per-element epilogue, shared by the conv kernels and pmfb_pointwise:
 *   v = alpha1[c]*v + beta1[c];  v += r1;  v = act(v);  v = alpha2[c]*v + beta2[c];
 *   v *= mul;  v += r2;  (optionally) v = round_to_tf32(v)
 * NULL pointers skip the corresponding step. */
typedef struct {
  const float* alpha1;
  const float* beta1;
  const float* alpha2;
  const float* beta2;
  pmfb_view r1;
  pmfb_view mul;
  pmfb_view r2;
  int32_t act;       /* pmfb_act */
  int32_t round_out; /* 1: store values rounded (RN) to tf32 so a following tensor-core conv is exact on them */
} pmfb_epilogue;

/* A gather source for the implicit GEMM: up to 5 TMA dimensions ordered
 * (channel, w, parity, h, n).  dims[] in elements, strides[] in BYTES for dims 1..4
 * (dim 0 is contiguous).  Stride-1 convs use parity dim = 1; stride-2 convs view the
 * input as (2C, W/2, 2, H/2, N) so that every tap is a plain box. */
typedef struct {
  const float* ptr;
  uint64_t dims[5];
  uint64_t strides[4];
} pmfb_tma_src;

/* Implicit-GEMM convolution on tcgen05 (kind::tf32, fp32 accumulate in TMEM).
 *   out[n,y,x,co] = epilogue( sum_{tap,ci} x[n, y+dh[tap], x+dw[tap], (parity dp[tap]), dc[tap]+ci] * w[tap][co][ci] )
 * Also used for dgrad (taps negated, weights packed [tap][ci][co]). */
typedef struct {
  pmfb_tma_src x;
  const float* w; /* packed [n_taps][c_out][c_in], c_in contiguous, values already tf32-rounded */
  int32_t c_in, c_out, n_taps;
  int32_t tap_dc[PMFB_MAX_TAPS], tap_dw[PMFB_MAX_TAPS], tap_dp[PMFB_MAX_TAPS], tap_dh[PMFB_MAX_TAPS];
  int32_t tap_wi[PMFB_MAX_TAPS]; /* weight slab used by tap t (only if use_tap_wi; else slab t): lets the     */
  int32_t use_tap_wi;            /* stride-2 dgrad run a SUBSET of the taps per input parity class            */
  int32_t n_batch, out_h, out_w; /* logical output grid                        */
  int32_t tile_w, tile_h;        /* tile_w*tile_h == 128 output pixels per CTA */
  int32_t n_tile;                /* output channels per CTA: multiple of 16, <= 256 */
  float* out;
  int64_t o_sn, o_sy, o_sx; /* element strides of the output view (channel stride 1) */
  pmfb_epilogue epi;
  /* Optional fused BatchNorm statistics of the epilogue result (nn.BatchNorm2d training mode, which follows every
   * conv of salsanext.py / pmf_net.py / torchvision's blocks): bn_stats[0:c_out] += sum over all output pixels,
   * bn_stats[c_out:2*c_out] += sum of squares (double accumulators, caller zeroes; same contract as pmfb_bn_stats).
   * Only where pmfb_conv_fused_stats_ok(desc) returns 1; NULL = off. */
  double* bn_stats;
  /* PMFB_DT_*: element type of x.ptr and w (x.strides stay in BYTES).  16-bit operands: stride-1 layers on the halo kernel
   * only (pmfb_conv16_ok), c_in % 8 == 0. */
  int32_t dtype;
  /* 1: `out` addresses an fp16 buffer (o_sn / o_sy / o_sx stay ELEMENT strides, c_out and o_sx multiples of 8): the result
   * is stored rounded to fp16 and the fused statistics are those of the stored values.  Only with the fused-statistics
   * epilogue of the halo kernel (training-mode conv -> [LeakyReLU] -> BatchNorm: the pre-BN activation is read twice more by
   * the BatchNorm passes and never by a convolution). */
  int32_t out_half;
} pmfb_conv_desc;

/* Weight gradient on tcgen05: dw[tap][ci][co] += sum_{n,y,x} x[n,y+dh,x+dw,..,dc+ci] * dy[n,y,x,co]
 * (both operands MN-major, pixels are the GEMM K dimension, split-K with fp32 red.add).
 * dw must be zeroed by the caller before the first call of a step. */
typedef struct {
  pmfb_tma_src x;
  pmfb_tma_src dy; /* (c_out, W, 1, H, N) */
  int32_t c_in, c_out, n_taps;
  int32_t tap_dc[PMFB_MAX_TAPS], tap_dw[PMFB_MAX_TAPS], tap_dp[PMFB_MAX_TAPS], tap_dh[PMFB_MAX_TAPS];
  int32_t n_batch, out_h, out_w; /* pixel grid of dy */
  int32_t ptile_w, ptile_h;      /* ptile_w*ptile_h == 32 pixels per pipeline stage */
  int32_t n_tile;                /* c_out columns per CTA: multiple of 32, <= 256 */
  int32_t ksplit;                /* number of pixel-range splits (grid.z) */
  float* dw;                     /* packed [n_taps][c_in][c_out] */
  /* PMFB_DT_F32 (fp32 x and dy, kind::tf32) or PMFB_DT_BF16 (bf16 shadows of both, kind::f16; stride-1 layers on the halo
   * kernel only, c_in and c_out multiples of 8; see pmfb_wgrad16_ok).  x.strides / dy.strides stay in BYTES. */
  int32_t dtype;
  int32_t reserved;
} pmfb_wgrad_desc;

int pmfb_abi_version(void);
/* Multiprocessor count of the current device (148 on a B200): what the library sizes its grids from. */
int pmfb_sm_count(void);
const char* pmfb_last_error(void);
/* 0 if an sm_100 device is present and the driver entry points resolve. */
int pmfb_init(void);

int pmfb_conv_fwd(const pmfb_conv_desc* d, void* stream);
/* 1 if pmfb_conv_fwd can accumulate desc->bn_stats inside its epilogue for this geometry / epilogue (stride-1 layers on
 * the halo kernel with a compile-time epilogue variant), else 0 (the caller then runs pmfb_bn_stats on the output). */
int pmfb_conv_fused_stats_ok(const pmfb_conv_desc* d);
int pmfb_conv_wgrad(const pmfb_wgrad_desc* d, void* stream);
/* 1 if pmfb_conv_fwd accepts 16-bit operands (dtype F16 / BF16) for this geometry, else 0. */
int pmfb_conv16_ok(const pmfb_conv_desc* d);
int pmfb_wgrad16_ok(const pmfb_wgrad_desc* d);

/* ---------------------------------------------------------------------------------------------
 * Layout / packing kernels
 * ------------------------------------------------------------------------------------------- */

/* cudaMemsetAsync(ptr, 0, bytes) on the stream: zeroes the fp64 reduction scratch and the packed wgrad buffers. */
int pmfb_memset_zero(void* ptr, size_t bytes, void* stream);

/* NCHW-ish strided source (element strides s_n,s_c,s_h,s_w; e.g. the channel-slice views that
 * tasks/pmf/trainer.py:296-297 passes) -> NHWC with c_dst channels at a pixel stride of dst_pix_stride
 * elements (= c_dst for a dense tensor; larger when dst is a channel slice of a concat buffer):
 *   dst[n,y,x, s*C + c] = src[n,c,y, x + s - n_shift/2]   (0 outside the row), s in [0,n_shift)
 *   channels >= n_shift*C are zero.  n_shift=1: plain NCHW->NHWC (+channel zero-pad);
 *   n_shift=7: the horizontally-unrolled input of the 7x7 stem (pmf_net.py:69-70), which turns the
 *   7x7x3 convolution into a 7-tap (vertical) convolution over 21(+11 zero) channels on tcgen05.
 * round_out: store tf32-rounded. */
int pmfb_pack_input(const float* src, int64_t s_n, int64_t s_c, int64_t s_h, int64_t s_w, int32_t n, int32_t c,
                    int32_t h, int32_t w, int32_t n_shift, float* dst, int32_t c_dst, int64_t dst_pix_stride,
                    int32_t round_out, void* stream);

/* dense NHWC view -> dense NCHW (module outputs, reference layout). */
int pmfb_nhwc_to_nchw(const pmfb_view* src, int32_t n, int32_t h, int32_t w, int32_t c, float* dst, void* stream);

/* OIHW conv weight -> packed tf32-rounded operands of the implicit GEMM, zero-padded to (c_out_p, c_in_p)
 * (both multiples of 4, so 3/5/17-channel tensors become TMA-legal):
 *   fwd  [taps][c_out_p][c_in_p]   (pmfb_conv_fwd forward)     if fwd   != NULL
 *   dgrad[taps][c_in_p][c_out_p]   (pmfb_conv_fwd as dgrad)    if dgrad != NULL
 * stem=1: the 7x7x3 stem; taps = kh and the packed input channel is kw_i*c_in + c (c_in_p >= kw*c_in).
 * round_out: 1 = store tf32-rounded values (the default kind::tf32 mode); 0 = keep fp32 (input of pmfb_split_tf32). */
int pmfb_pack_weight(const float* w, int32_t c_out, int32_t c_in, int32_t kh, int32_t kw, int32_t stem,
                     int32_t c_out_p, int32_t c_in_p, float* fwd, float* dgrad, int32_t round_out, void* stream);

/* Precise mode (3xTF32): split an NHWC fp32 view into tf32-exact parts  hi = rna_tf32(x), lo = rna_tf32(x - hi)  so that
 *   x*w ~= hi_x*hi_w + hi_x*lo_w + lo_x*hi_w   (three kind::tf32 UMMAs per K step into the same TMEM accumulator; every
 * product is exact in fp32) recovers the reference's fp32 convolution arithmetic (pmf_net.py / salsanext.py nn.Conv2d) to
 * ~2^-21 per operand.  The three products are laid out along the GEMM K dimension (channels), zero-padded per part to a
 * multiple of 32 channels cp = roundup(c, 32), and the implicit-GEMM kernels run unchanged over c_in = 3*cp:
 *   mode 0: out[..., 3*cp] = [hi | hi | lo]   (activations / gradients: the A operand of pmfb_conv_fwd)
 *   mode 1: out[..., 3*cp] = [hi | lo | hi]   (packed weights viewed as (1, taps, c_out, c_in))
 *   mode 2: out[..., c] = hi     mode 3: out[..., c] = lo   (the x / dy operands of the three pmfb_conv_wgrad launches)
 * out: channel stride 1, element strides o_sn / o_sy / o_sx. */
int pmfb_split_tf32(const pmfb_view* in, int32_t n, int32_t h, int32_t w, int32_t c, float* out, int64_t o_sn, int64_t o_sy,
                    int64_t o_sx, int32_t mode, void* stream);

/* packed wgrad [taps][c_in_p][c_out_p] -> OIHW gradient (grad = or += depending on accumulate). */
int pmfb_unpack_wgrad(const float* packed, int32_t c_out, int32_t c_in, int32_t kh, int32_t kw, int32_t stem,
                      int32_t c_out_p, int32_t c_in_p, float* grad, int32_t accumulate, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Elementwise / normalisation kernels (NHWC views, C % 4 == 0)
 * ------------------------------------------------------------------------------------------- */

/* out = epilogue(in) elementwise; `in` may be NULL-ptr view meaning zeros.  Strides of 0 broadcast. */
int pmfb_pointwise(const pmfb_view* in, float* out, int64_t o_sn, int64_t o_sy, int64_t o_sx, int32_t n, int32_t h,
                   int32_t w, int32_t c, const pmfb_epilogue* epi, void* stream);

/* pmfb_pointwise that ALSO stores the result as 16-bit shadows (same element strides and channel offset as out): out16 in
 * dtype16 = PMFB_DT_F16 or PMFB_DT_BF16 (fp16 values saturate to the fp16 range) and / or out16_bf16 in bf16.  The BN-apply
 * pass that produces a conv input writes the operands the kind::f16 convolution (fp16) and weight gradient (bf16) will read
 * in the same pass.  Both NULL: identical to pmfb_pointwise.
 * in_half: the input view addresses an fp16 buffer (the pre-BatchNorm activation a convolution stored with
 * pmfb_conv_desc.out_half) with the given ELEMENT strides. */
int pmfb_pointwise16(const pmfb_view* in, float* out, int64_t o_sn, int64_t o_sy, int64_t o_sx, int32_t n, int32_t h,
                     int32_t w, int32_t c, const pmfb_epilogue* epi, void* out16, int32_t dtype16, void* out16_bf16,
                     int32_t in_half, void* stream);

/* Training-mode BatchNorm finalisation fused into the BN-apply pass (pmfb_pointwise16_bn): from the fp64 channel sums the
 * convolution's epilogue accumulated (sums[0:C] = sum x, sums[C:2C] = sum x^2 over `count` pixels) every thread derives the
 * affine of its own channels, alpha = gamma * invstd, beta = bias - mean * alpha, and uses it as the epilogue's
 * alpha1 / beta1 (which must be NULL); one thread per channel also stores alpha / beta / mean / invstd (read by the backward
 * passes) and updates the running statistics -- exactly pmfb_bn_finalize's arithmetic, without its launch. */
typedef struct {
  const double* sums;
  int64_t count;
  const float* gamma; /* may be NULL (= 1) */
  const float* beta;  /* may be NULL (= 0) */
  float* running_mean; /* may be NULL */
  float* running_var;  /* may be NULL */
  float momentum, eps;
  float *alpha_out, *beta_out, *mean_out, *invstd_out; /* C floats each, all required */
} pmfb_bn_fuse;
/* in_half must be 1 (the fused form exists for the fp16 pre-BatchNorm activations of the "f16" mode). */
int pmfb_pointwise16_bn(const pmfb_view* in, float* out, int64_t o_sn, int64_t o_sy, int64_t o_sx, int32_t n, int32_t h,
                        int32_t w, int32_t c, const pmfb_epilogue* epi, void* out16, int32_t dtype16, void* out16_bf16,
                        int32_t in_half, const pmfb_bn_fuse* bn, void* stream);

/* 16-bit shadow of an fp32 NHWC view (c % 8 == 0; out16 has element strides o_sn / o_sy / o_sx, channel stride 1). */
int pmfb_convert16(const pmfb_view* in, int32_t n, int32_t h, int32_t w, int32_t c, void* out16, int64_t o_sn, int64_t o_sy,
                   int64_t o_sx, int32_t dtype16, void* stream);

/* sums[0:C] += sum_x, sums[C:2C] += sum_x^2 over all pixels (double accumulators, caller zeroes). */
int pmfb_bn_stats(const pmfb_view* x, int32_t n, int32_t h, int32_t w, int32_t c, double* sums, void* stream);

/* BatchNorm2d bookkeeping (nn.BatchNorm2d, eps 1e-5, momentum 0.1).
 * train (sums != NULL): mean/var(biased) from sums over `count` elements; running stats updated in
 *   place with the unbiased variance; eval (sums == NULL): uses running stats.
 * Writes alpha = gamma*invstd, beta_out = beta - mean*alpha, and mean / invstd (for backward). */
int pmfb_bn_finalize(const double* sums, int64_t count, int32_t c, const float* gamma, const float* beta,
                     float* running_mean, float* running_var, float momentum, float eps, float* alpha,
                     float* beta_out, float* mean_out, float* invstd_out, void* stream);

/* BatchNorm / activation backward.  Common input gradient
 *   g = dy * (mul.ptr ? mul : 1) * act'(z)        act_z in {NONE, RELU, LEAKY, SIGMOID}, derivative taken at the
 *   OUTPUT z of the activation; if act_z != NONE and z.ptr == NULL, z is recomputed as act(alpha*x + beta)
 *   (the sigmoid attention gate of pmf_net.py:20-35 is never stored).
 * pass 1: red[0:C] += sum g, red[C:2C] += sum g*xhat, xhat = (x-mean)*invstd  (fp64, caller zeroes). */
int pmfb_bn_bwd_reduce(const pmfb_view* dy, const pmfb_view* mul, const pmfb_view* z, int32_t act_z,
                       const pmfb_view* x, const float* mean, const float* invstd, const float* alpha,
                       const float* beta, int32_t n, int32_t h, int32_t w, int32_t c, double* red, void* stream);

/* pass 2:
 *   mean != NULL:  dx = gamma*invstd*(g - S1/M - xhat*S2/M)       (BatchNorm backward, M = n*h*w)
 *   mean == NULL:  dx = g                                          (plain activation backward)
 *   leaky_x: dx *= (x > 0 ? 1 : 0.01)   (the conv->LeakyReLU->BN order of salsanext.py / pmf_net.py:13-18: x is
 *            the stored activation output, whose sign is the pre-activation's sign)
 *   dx (may be NULL) optionally tf32-rounded; dgamma = S2, dbeta = S1; colsum[0:C] += sum dx (bias gradient of
 *   the preceding conv; fp64, caller zeroes) if colsum != NULL; if g_out != NULL, g is written (or accumulated)
 *   there: the identity-branch gradient of a residual add. */
int pmfb_bn_bwd_apply(const pmfb_view* dy, const pmfb_view* mul, const pmfb_view* z, int32_t act_z,
                      const pmfb_view* x, const float* mean, const float* invstd, const float* alpha,
                      const float* beta, const float* gamma, const double* red, int32_t leaky_x, int32_t n,
                      int32_t h, int32_t w, int32_t c, float* dx, int64_t d_sn, int64_t d_sy, int64_t d_sx,
                      int32_t round_out, float* dgamma, float* dbeta, double* colsum, float* g_out, int64_t g_sn,
                      int64_t g_sy, int64_t g_sx, int32_t g_accumulate, void* stream);

/* pmfb_bn_bwd_apply that also stores dx as bf16 (dx16: same strides d_sn / d_sy / d_sx as dx): the operand of the
 * kind::f16 dgrad and weight gradient.  dx may then be NULL (only the bf16 copy is stored). */
int pmfb_bn_bwd_apply16(const pmfb_view* dy, const pmfb_view* mul, const pmfb_view* z, int32_t act_z,
                        const pmfb_view* x, const float* mean, const float* invstd, const float* alpha,
                        const float* beta, const float* gamma, const double* red, int32_t leaky_x, int32_t n,
                        int32_t h, int32_t w, int32_t c, float* dx, int64_t d_sn, int64_t d_sy, int64_t d_sx,
                        int32_t round_out, float* dgamma, float* dbeta, double* colsum, float* g_out, int64_t g_sn,
                        int64_t g_sy, int64_t g_sx, int32_t g_accumulate, void* dx16, int32_t x_half, void* stream);
/* x_half (here and in pmfb_bn_bwd_apply16): the view x addresses an fp16 buffer (element strides), see pmfb_conv_desc.out_half. */
int pmfb_bn_bwd_reduce16(const pmfb_view* dy, const pmfb_view* mul, const pmfb_view* z, int32_t act_z, const pmfb_view* x,
                         const float* mean, const float* invstd, const float* alpha, const float* beta, int32_t n, int32_t h,
                         int32_t w, int32_t c, double* red, int32_t x_half, void* stream);

/* out[i*C + c] (+)= sum over pixels (per image if per_image) of x; double accumulators, caller zeroes. */
int pmfb_colsum(const pmfb_view* x, int32_t n, int32_t h, int32_t w, int32_t c, int32_t per_image, double* out,
                void* stream);
/* EPMF sparse-conv mask ops (pc_processor/models/epmf_net.py).
 * pmfb_pixel_mask:   mask[n,y,x] = (sum_c |x[n,y,x,c]| != 0) ? 1 : 0                      (epmf_net.py:67)
 * pmfb_mask_maxpool: mask_out = MaxPool2d(k, stride, padding=0, dilation)(F.pad(mask_in, pad)), dense (N,H,W) fp32 maps,
 *                    output (h + 2 pad - dilation (k-1) - 1) / stride + 1 per side                 (epmf_net.py:43-44)
 * pmfb_pixel_scale:  out = (act(in * pre_mask[p]) * alpha[c] + beta[c] + r) * post_mask[p], optional tf32 rounding;
 *                    NULL pointers skip a step: the x*mask / LeakyReLU / eval BatchNorm / +shortcut / *mask chain
 *                    around SparseVariantConv                                                      (epmf_net.py:31,49,69-82) */
int pmfb_pixel_mask(const pmfb_view* x, int32_t n, int32_t h, int32_t w, int32_t c, float* mask, void* stream);
int pmfb_mask_maxpool(const float* mask_in, int32_t n, int32_t h, int32_t w, int32_t k, int32_t stride, int32_t dilation,
                      int32_t pad, float* mask_out, void* stream);
int pmfb_pixel_scale(const pmfb_view* in, int32_t n, int32_t h, int32_t w, int32_t c, const float* pre_mask, int32_t act,
                     const float* alpha, const float* beta, const pmfb_view* r, const float* post_mask, float* out,
                     int64_t o_sn, int64_t o_sy, int64_t o_sx, int32_t round_out, void* stream);

/* Batched weight layout jobs: ONE launch performs pmfb_pack_weight (unpack = 0: src = OIHW weight, dst = forward
 * packing or NULL, dst2 = dgrad packing or NULL) or pmfb_unpack_wgrad (unpack = 1: src = packed gradient, dst = OIHW
 * gradient, dst2 unused) for every job of a DEVICE-resident table.  job.start is the exclusive prefix sum of the jobs'
 * work sizes (pack: taps*c_out_p*c_in_p, unpack: c_out*c_in*kh*kw); total_work is the grand total.  Replaces the
 * per-layer calls of the reference-shaped module tree (one nn.Conv2d each: salsanext.py:12-61, pmf_net.py:13-29,
 * torchvision resnet) inside a captured training step. */
typedef struct {
  const float* src;
  float* dst;
  float* dst2;
  int32_t c_out, c_in, kh, kw, stem, c_out_p, c_in_p, accumulate;
  int32_t no_round; /* pack only: 1 = keep fp32 (precise mode, see pmfb_split_tf32) */
  int32_t reserved;
  int64_t start;
} pmfb_weight_job;
int pmfb_weight_jobs(int32_t unpack, const pmfb_weight_job* jobs_device, int32_t n_jobs, int64_t total_work,
                     void* stream);

/* float dst = (or +=) (float)src * scale, n elements, optionally tf32-rounded. */
int pmfb_d2f(const double* src, float* dst, int64_t n, float scale, int32_t accumulate, int32_t round_out,
             void* stream);

/* 3x3 stride-2 pad-1 pooling.  kind 0: AvgPool2d(count_include_pad) (salsanext.py:65);
 * kind 1: MaxPool2d (torchvision stem), idx (uint8, dense [n,ho,wo,c]) records the arg-max tap 0..8.
 * chan_scale (may be NULL): per-(n,c) Dropout2d scale applied to the pooled value (salsanext.py:92-96:
 * pool(dropout(x)) == dropout-scale * pool(x) because the mask is constant over a plane). */
/* out16 / out16_bf16 (here, in pmfb_pixel_shuffle and in pmfb_upsample2x; either may be NULL): fp16 / bf16 shadows of the
 * result with the same ELEMENT strides as out -- the operands of the "f16" mode's kind::f16 convolutions and weight
 * gradients, written by the producing pass instead of a separate pmfb_convert16. */
int pmfb_pool3s2(int32_t kind, const pmfb_view* x, int32_t n, int32_t h, int32_t w, int32_t c,
                 const float* chan_scale, float* out, int64_t o_sn, int64_t o_sy, int64_t o_sx, uint8_t* idx,
                 int32_t round_out, void* out16, void* out16_bf16, void* stream);
int pmfb_pool3s2_bwd(int32_t kind, const pmfb_view* dy, int32_t n, int32_t h, int32_t w, int32_t c,
                     const float* chan_scale, float* dx, int64_t d_sn, int64_t d_sy, int64_t d_sx,
                     const uint8_t* idx, int32_t accumulate, void* stream);

/* PixelShuffle(2) (salsanext.py:137): out[n,2y+i,2x+j,c] = x[n,y,x,4c+2i+j] * (chan_scale ? chan_scale[n*C+c] : 1).
 * (h,w,c) are the OUTPUT-side channel count c and INPUT spatial dims h,w.  bwd is the inverse gather. */
int pmfb_pixel_shuffle(const pmfb_view* x, int32_t n, int32_t h, int32_t w, int32_t c, const float* chan_scale,
                       float* out, int64_t o_sn, int64_t o_sy, int64_t o_sx, int32_t round_out, void* out16, void* out16_bf16,
                       void* stream);
int pmfb_pixel_shuffle_bwd(const pmfb_view* dy, int32_t n, int32_t h, int32_t w, int32_t c,
                           const float* chan_scale, float* dx, int64_t d_sn, int64_t d_sy, int64_t d_sx,
                           int32_t accumulate, int32_t round_out, void* stream);

/* nn.Upsample(scale_factor=2, mode="bilinear"), align_corners=False (pmf_net.py:191-210). (h,w) = input dims. */
int pmfb_upsample2x(const pmfb_view* x, int32_t n, int32_t h, int32_t w, int32_t c, float* out, int64_t o_sn,
                    int64_t o_sy, int64_t o_sx, int32_t round_out, void* out16, void* out16_bf16, void* stream);
int pmfb_upsample2x_bwd(const pmfb_view* dy, int32_t n, int32_t h, int32_t w, int32_t c, float* dx, int64_t d_sn,
                        int64_t d_sy, int64_t d_sx, int32_t accumulate, void* stream);

/* F.softmax(dim=1) of NHWC logits, written as dense NCHW probabilities (pmf_net.py:177-178,221);
 * bwd: dz[n,y,x,c] = p*(dp - sum_c dp*p) from dense NCHW p and dp.  c = number of classes (<= 64); the NHWC
 * view holds c rounded up to a multiple of 4 channels, pad channels are ignored (fwd) / written as 0 (bwd). */
int pmfb_softmax_nchw(const pmfb_view* logits, int32_t n, int32_t h, int32_t w, int32_t c, float* out, void* stream);
int pmfb_softmax_nchw_bwd(const float* p, const float* dp, int32_t n, int32_t h, int32_t w, int32_t c, float* dz,
                          int64_t d_sn, int64_t d_sy, int64_t d_sx, int32_t round_out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Fused loss head: the loss block tasks/pmf/trainer.py:305-332 applies to the two softmax maps
 * ------------------------------------------------------------------------------------------- */

/* One pass over the two dense NCHW probability maps (n, c, h, w) and the int64 labels (n, h, w):
 *   focal:       -(1 - p_t)^focal_gamma * log(clamp(p_t, 1e-6)) * alpha[t] summed over the pixels with label > 0
 *                (pc_processor/loss/focal_softmax.py:28-62 with softmax=False, mask = label > 0), per head;
 *   perception:  entropy -> confidence = 1 - H(p)/log c per head; guide weights 1[d>0]|d|1[conf>=tau] with
 *                d = conf_lidar - conf_camera; KLDivLoss(reduction="none")(log p_lidar, p_camera) * w_camera and the
 *                reverse direction, summed over all elements (trainer.py:231-252, 305-319).
 * sums (8 doubles, caller zeroes): [0] focal lidar, [1] focal camera, [2] number of labelled pixels, [3] KL numerator of
 * loss_per_pcd, [4] of loss_per_img, [5] sum of lidar entropies, [6] sum of camera entropies.  The loss value is
 *   w_focal * (sums[0] + sums[1]) / sums[2] + w_per * (sums[3] + sums[4]) / (n*c*h*w).
 * d_lidar / d_camera (dense NCHW, may both be NULL): WRITTEN with the gradient of exactly that value with respect to the
 * two maps (through the logs, the entropies and the guide weights, like autograd on the reference's expressions). */
int pmfb_loss_head(const float* p_lidar, const float* p_camera, const int64_t* label, int32_t n, int32_t c, int32_t h,
                   int32_t w, const float* alpha, float focal_gamma, float tau, float w_focal, float w_per, float* d_lidar,
                   float* d_camera, double* sums, void* stream);

/* Lovasz-softmax (pc_processor/loss/lovasz_softmax.py:55-145; classes="present", per_image=False, pixels whose label ==
 * ignore dropped) of one or two heads (probs1 may be NULL) in one pipeline: compaction of the labelled pixels, one LSD
 * radix sort of (head, class, descending |fg - p|, pixel) keys, foreground scan, Jaccard gradient, per-class losses.
 * loss_out[head] += mean over the present classes of the class losses (doubles, caller zeroes).
 * d_probs0 / d_probs1 (dense NCHW, may be NULL): ACCUMULATED (+=) with grad_scale * d loss / d probs.
 * workspace: pmfb_lovasz_workspace_bytes(n*h*w, c, heads) bytes, 256-byte aligned, contents undefined on entry; nothing is
 * read back by the host (grids are sized for the worst case and take the labelled-pixel count from device memory). */
size_t pmfb_lovasz_workspace_bytes(int64_t n_pixels, int32_t c, int32_t n_heads);
int pmfb_lovasz(const float* probs0, const float* probs1, const int64_t* label, int32_t n, int32_t c, int32_t h, int32_t w,
                int32_t ignore, float grad_scale, float* d_probs0, float* d_probs1, double* loss_out, void* workspace,
                size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Post-processing / pre-processing
 * ------------------------------------------------------------------------------------------- */

/* KNN.forward (pc_processor/postproc/knn.py:55-143): per point, S x S window of the range image
 * (zero padded), |range - r| * (1 - gaussian) weights, k smallest (ties -> lower window index),
 * cutoff -> class nclasses, vote over classes 1..nclasses-1, first max wins.
 * inv_gauss: S*S fp32 table (1 - normalised gaussian), computed by the host exactly as knn.py:12-34.
 * proj_argmax / px / py / out are int64 as the reference passes them. px = column, py = row. */
int pmfb_knn_vote(const float* proj_range, const int64_t* proj_argmax, int32_t h, int32_t w,
                  const float* unproj_range, const int64_t* px, const int64_t* py, int64_t n_points,
                  const float* inv_gauss, int32_t search, int32_t knn, float cutoff, int32_t nclasses,
                  int64_t* out, void* stream);

/* The same vote for ANY NUMBER OF FRAMES in one launch (the reference is un-batched, knn.py:56-59): the
 * range / argmax images are (n_frames, h, w), the point arrays are the frames' points concatenated, point_offsets
 * (n_frames + 1 int64, device memory) gives each frame's slice.  One thread per point, top-k in registers (knn <= 8). */
int pmfb_knn_vote_batched(const float* proj_range, const int64_t* proj_argmax, int32_t n_frames, int32_t h, int32_t w,
                          const float* unproj_range, const int64_t* px, const int64_t* py, const int64_t* point_offsets,
                          int64_t n_points, const float* inv_gauss, int32_t search, int32_t knn, float cutoff,
                          int32_t nclasses, int64_t* out, void* stream);

/* Inference tail on the device (tasks/pmf_eval_semantickitti/infer.py:107-146, tasks/pmf_eval_nuscenes/infer.py:18-38):
 * pmfb_argmax_nchw:   label[b,y,x] = argmax_c probs[b,c,y0+y,x0+x] (first maximum, like torch.argmax) over the crop
 *                     window (y0, x0, out_h, out_w) of a dense NCHW map; conf (may be NULL) = the winning probability.
 * pmfb_lut_remap:     out[i] = lut[labels[i]] (class_map_lut_inv: training ids -> dataset ids; out-of-range -> 0).
 * pmfb_merge_cameras: per point the prediction of the camera with the highest confidence (ties: lowest camera; no
 *                     camera: -1).  Inputs are the cameras' per-point arrays concatenated (point index int64, confidence
 *                     f32, prediction int64, camera id int32 in [0,8)); scratch: pc_size uint64. */
int pmfb_argmax_nchw(const float* probs, int32_t n, int32_t c, int32_t h, int32_t w, int32_t y0, int32_t x0, int32_t out_h,
                     int32_t out_w, int64_t* label, float* conf, void* stream);
int pmfb_lut_remap(const int64_t* labels, int64_t n, const int32_t* lut, int32_t lut_size, int32_t* out, void* stream);
int pmfb_merge_cameras(const int64_t* point_idx, const float* conf, const int64_t* argmax, const int32_t* cam,
                       int64_t n_entries, int64_t pc_size, uint64_t* scratch, int64_t* merged, void* stream);

/* IOUEval.addBatch (pc_processor/metrics/iou_eval.py:31-57) on the device: conf[pred[i], target[i]] += 1 (row = prediction,
 * column = ground truth; int64 nclasses x nclasses matrix, accumulated; entries outside [0, nclasses) are skipped). */
int pmfb_confusion_add(const int64_t* pred, const int64_t* target, int64_t n, int32_t nclasses, int64_t* conf, void* stream);

/* Perspective projection + scatter (parser.py:209-227, perspective_view_loader.py:87-131):
 * q = M*[x y z 1]^T in float64, keep x>0.5 and 0<u<W, 0<v<H, truncate to (row,col); per pixel the point
 * with the HIGHEST index wins (numpy fancy-assignment order).  Two passes over `winner` (int32 H*W,
 * caller provides; filled with -1 by this call).  feat: dense (5,H,W) [depth,x,y,z,i]; mask (H,W) f32;
 * label_img (H,W) f32; rows/cols (N) int32 (-1 for dropped points); depth (N) f32 = |xyz|.
 * points: device (N,4) fp32, 16-byte aligned; labels: device int32 (may be NULL);
 * proj_matrix: HOST pointer to the 3x4 row-major float64 P2*Tr (it is passed by value to the kernel). */
int pmfb_project_scatter(const float* points, const int32_t* labels, int64_t n_points, const double* proj_matrix,
                         int32_t h, int32_t w, int32_t* winner, float* feat, float* mask, float* label_img,
                         int32_t* rows, int32_t* cols, float* depth, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PMFB_H_ */
