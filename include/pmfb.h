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

#define PMFB_ABI_VERSION 1
#define PMFB_MAX_TAPS 9

typedef enum {
  PMFB_OK = 0,
  PMFB_ERR_INVALID = -1,   /* bad argument / unsupported geometry */
  PMFB_ERR_CUDA = -2,      /* CUDA runtime / driver error        */
  PMFB_ERR_NO_DEVICE = -3  /* no sm_100 device available          */
} pmfb_status;

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
  int32_t n_batch, out_h, out_w; /* logical output grid                        */
  int32_t tile_w, tile_h;        /* tile_w*tile_h == 128 output pixels per CTA */
  int32_t n_tile;                /* output channels per CTA: multiple of 16, <= 256 */
  float* out;
  int64_t o_sn, o_sy, o_sx; /* element strides of the output view (channel stride 1) */
  pmfb_epilogue epi;
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
} pmfb_wgrad_desc;

int pmfb_abi_version(void);
const char* pmfb_last_error(void);
/* 0 if an sm_100 device is present and the driver entry points resolve. */
int pmfb_init(void);

int pmfb_conv_fwd(const pmfb_conv_desc* d, void* stream);
int pmfb_conv_wgrad(const pmfb_wgrad_desc* d, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PMFB_H_ */
