// Layout, packing and elementwise kernels of libpmf_b200.so (all HBM-bound; fp32 NHWC views, float4 per thread).
//
// Reference semantics being replaced (ATen kernels on the reference side):
//   pointwise        : the BN-apply / activation / residual / gate chains of salsanext.py:23-36,69-104,136-164,
//                      pmf_net.py:31-36 and torchvision BasicBlock
//   pool3s2          : nn.AvgPool2d(3,2,1) salsanext.py:65 and nn.MaxPool2d(3,2,1) of the torchvision stem (pmf_net.py:95)
//   pixel_shuffle    : nn.PixelShuffle(2) salsanext.py:137
//   upsample2x       : nn.Upsample(scale_factor=2, mode="bilinear") pmf_net.py:191-210
//   softmax_nchw     : F.softmax(dim=1) pmf_net.py:177-178,221
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.h"
#include "epilogue.cuh"

namespace pmfb {

// four fp32 values -> four 16-bit values (8 bytes).  fp16 saturates to +-65504 (conv inputs are O(1..100); an overflow to
// inf would poison a whole accumulator row).
__device__ __forceinline__ uint2 pack16(float4 v, int dt) {
  uint2 r;
  if (dt == PMFB_DT_BF16) {
    const __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
    r.x = *reinterpret_cast<const unsigned int*>(&a);
    r.y = *reinterpret_cast<const unsigned int*>(&b);
  } else {
    const float m = 65504.f;
    const __half2 a = __floats2half2_rn(fminf(fmaxf(v.x, -m), m), fminf(fmaxf(v.y, -m), m));
    const __half2 b = __floats2half2_rn(fminf(fmaxf(v.z, -m), m), fminf(fmaxf(v.w, -m), m));
    r.x = *reinterpret_cast<const unsigned int*>(&a);
    r.y = *reinterpret_cast<const unsigned int*>(&b);
  }
  return r;
}

static inline int grid_for(long long work, int threads, int max_blocks = 0) {
  if (max_blocks <= 0) max_blocks = sm_count() * 16;
  long long b = (work + threads - 1) / threads;
  if (b < 1) b = 1;
  if (b > max_blocks) b = max_blocks;
  return (int)b;
}

static inline bool view_ok(const pmfb_view* v) {
  return !v->ptr || ((((v->sn | v->sy | v->sx) % 4) == 0) && ((reinterpret_cast<uintptr_t>(v->ptr) & 15) == 0));
}
static inline bool out_ok(const float* p, int64_t sn, int64_t sy, int64_t sx) {
  return p && (((sn | sy | sx) % 4) == 0) && ((reinterpret_cast<uintptr_t>(p) & 15) == 0);
}

__device__ __forceinline__ float4 rnd4(float4 v) {
  v.x = round_tf32(v.x); v.y = round_tf32(v.y); v.z = round_tf32(v.z); v.w = round_tf32(v.w);
  return v;
}

// ------------------------------------------------------------------------------------ pointwise
// Thread = (fixed 4-channel group, pixel lane); four pixels per trip with every operand load issued before the first
// use.  Pixel-linear views (offset = pixel * sx) avoid all index divisions; broadcast / parity views decode.
struct PWV {
  const float* p;
  long long sn, sy, sx;
  int linear;
};
__device__ __forceinline__ const float* pw_at(const PWV& v, unsigned pix, unsigned hw, unsigned w, int c) {
  if (v.linear) return v.p + (long long)pix * v.sx + c;
  const unsigned n = pix / hw, q = pix - n * hw, y = q / w, x = q - y * w;
  return v.p + (long long)n * v.sn + (long long)y * v.sy + (long long)x * v.sx + c;
}
struct PWParams {
  PWV in, out, r1, mul, r2;
  const float* alpha1;
  const float* beta1;
  const float* alpha2;
  const float* beta2;
  int act, round_out;
  unsigned short* out16;  // optional 16-bit shadow of the result (same element offsets as out)
  int dt16;
  unsigned short* out16b;  // optional second shadow, always bf16 (the wgrad operand)
  pmfb_bn_fuse bn;         // BNF kernels: BatchNorm finalisation from the fused channel sums (see include/pmfb.h)
};

// OPS (bit 0: r1, bit 1: mul, bit 2: r2) is a compile-time mask of the per-pixel operands: the common launches (BN apply,
// concat copies) carry none or one, and with the unused operand arrays compiled out the kernel fits 5-6 blocks per SM
// instead of 2 (ncu: 104 registers, 24 % occupancy, 4.0 TB/s): more 16-byte loads in flight per SM.
// IH: the input view is an fp16 buffer (the pre-BatchNorm activation stored by the conv epilogue, pmfb_conv_desc.out_half);
// its four values travel packed (8 bytes) and twice as many pixels are kept in flight.
// BNF (with IH): alpha1 / beta1 are derived here from the convolution's fused channel sums (pmfb_bn_fuse) instead of being
// read from the vectors a separate bn_finalize launch would have written.
template <int OPS, bool IH, bool BNF = false>
__global__ void __launch_bounds__(256)
pointwise_kernel(PWParams P, unsigned npix, unsigned hw, unsigned w, int c4, int G) {
  constexpr bool kR1 = (OPS & 1) != 0, kMul = (OPS & 2) != 0, kR2 = (OPS & 4) != 0;
  constexpr int U = (IH && OPS == 0) ? 8 : 4;  // BN apply without extra operand streams: twice the pixels in flight
  const int L = 256 / G;
  const int gl = threadIdx.x % G, pl = threadIdx.x / G;
  const int cg = blockIdx.y * G + gl;
  if (pl >= L || cg >= c4) return;
  const int c = cg * 4;
  const float4 one = make_float4(1.f, 1.f, 1.f, 1.f), zero = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 a1 = one, b1 = zero;
  if constexpr (BNF) {  // pmfb_bn_finalize's arithmetic for this thread's four channels
    const int C = c4 * 4;
    const double cnt = (double)P.bn.count;
    float al[4], be[4], mu[4], is[4];
    const bool owner = blockIdx.x == 0 && pl == 0;  // one thread per channel group stores the vectors / running statistics
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const double m = P.bn.sums[c + j] / cnt;
      double v = P.bn.sums[C + c + j] / cnt - m * m;
      if (v < 0.0) v = 0.0;
      const float mean = (float)m, var = (float)v;
      const float invstd = 1.f / sqrtf(var + P.bn.eps);
      const float g = P.bn.gamma ? P.bn.gamma[c + j] : 1.f;
      const float b = P.bn.beta ? P.bn.beta[c + j] : 0.f;
      al[j] = g * invstd;
      be[j] = b - mean * al[j];
      mu[j] = mean;
      is[j] = invstd;
      if (owner) {
        if (P.bn.running_mean) P.bn.running_mean[c + j] = (1.f - P.bn.momentum) * P.bn.running_mean[c + j] + P.bn.momentum * mean;
        if (P.bn.running_var) {
          const float unbiased = cnt > 1.0 ? (float)(v * cnt / (cnt - 1.0)) : var;
          P.bn.running_var[c + j] = (1.f - P.bn.momentum) * P.bn.running_var[c + j] + P.bn.momentum * unbiased;
        }
      }
    }
    a1 = make_float4(al[0], al[1], al[2], al[3]);
    b1 = make_float4(be[0], be[1], be[2], be[3]);
    if (owner) {
      *reinterpret_cast<float4*>(P.bn.alpha_out + c) = a1;
      *reinterpret_cast<float4*>(P.bn.beta_out + c) = b1;
      *reinterpret_cast<float4*>(P.bn.mean_out + c) = make_float4(mu[0], mu[1], mu[2], mu[3]);
      *reinterpret_cast<float4*>(P.bn.invstd_out + c) = make_float4(is[0], is[1], is[2], is[3]);
    }
  } else {
    if (P.alpha1) a1 = ld4(P.alpha1 + c);
    if (P.beta1) b1 = ld4(P.beta1 + c);
  }
  const float4 a2 = P.alpha2 ? ld4(P.alpha2 + c) : one, b2 = P.beta2 ? ld4(P.beta2 + c) : zero;
  const unsigned stride = gridDim.x * L;
  for (unsigned p0 = blockIdx.x * L + pl; p0 < npix; p0 += U * stride) {
    float4 v[IH ? 1 : U], r1[kR1 ? U : 1], mu[kMul ? U : 1], r2[kR2 ? U : 1];
    uint2 vh[IH ? U : 1];
#pragma unroll
    for (int k = 0; k < U; ++k) {
      const unsigned p = p0 + k * stride;
      if (p < npix) {
        if constexpr (IH)
          vh[k] = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const unsigned short*>(P.in.p) + ((long long)p * P.in.sx + c)));
        else
          v[k] = P.in.p ? ld4(pw_at(P.in, p, hw, w, c)) : zero;
        if constexpr (kR1) r1[k] = ld4(IH ? P.r1.p + (long long)p * P.r1.sx + c : pw_at(P.r1, p, hw, w, c));
        if constexpr (kMul) mu[k] = ld4(pw_at(P.mul, p, hw, w, c));
        if constexpr (kR2) r2[k] = ld4(IH ? P.r2.p + (long long)p * P.r2.sx + c : pw_at(P.r2, p, hw, w, c));
      }
    }
#pragma unroll
    for (int k = 0; k < U; ++k) {
      const unsigned p = p0 + k * stride;
      if (p < npix) {
        float4 o;
        if constexpr (IH) {
          const float2 fa = __half22float2(*reinterpret_cast<const __half2*>(&vh[k].x));
          const float2 fb = __half22float2(*reinterpret_cast<const __half2*>(&vh[k].y));
          o = make_float4(fa.x, fa.y, fb.x, fb.y);
        } else {
          o = v[k];
        }
        o.x = o.x * a1.x + b1.x; o.y = o.y * a1.y + b1.y; o.z = o.z * a1.z + b1.z; o.w = o.w * a1.w + b1.w;
        if constexpr (kR1) { o.x += r1[k].x; o.y += r1[k].y; o.z += r1[k].z; o.w += r1[k].w; }
        if (P.act) { o.x = epi_act(P.act, o.x); o.y = epi_act(P.act, o.y); o.z = epi_act(P.act, o.z); o.w = epi_act(P.act, o.w); }
        o.x = o.x * a2.x + b2.x; o.y = o.y * a2.y + b2.y; o.z = o.z * a2.z + b2.z; o.w = o.w * a2.w + b2.w;
        if constexpr (kMul) { o.x *= mu[k].x; o.y *= mu[k].y; o.z *= mu[k].z; o.w *= mu[k].w; }
        if constexpr (kR2) { o.x += r2[k].x; o.y += r2[k].y; o.z += r2[k].z; o.w += r2[k].w; }
        if (P.round_out) o = rnd4(o);
        float* op = const_cast<float*>(IH ? P.out.p + (long long)p * P.out.sx + c : pw_at(P.out, p, hw, w, c));
        *reinterpret_cast<float4*>(op) = o;
        if (P.out16) *reinterpret_cast<uint2*>(P.out16 + (op - P.out.p)) = pack16(o, P.dt16);
        if (P.out16b) *reinterpret_cast<uint2*>(P.out16b + (op - P.out.p)) = pack16(o, PMFB_DT_BF16);
      }
    }
  }
}

template <int OPS, bool IH = false, bool BNF = false>
static int launch_pointwise_t(const PWParams& P, long long npix, int h, int w, int c4, cudaStream_t stream) {
  const int G = c4 < 256 ? c4 : 256;
  const int L = 256 / G;
  const int gy = (c4 + G - 1) / G;
  static int per_sm = 0;
  if (per_sm == 0 && (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pointwise_kernel<OPS, IH, BNF>, 256, 0) != cudaSuccess || per_sm < 1))
    per_sm = 4;
  long long gx = (npix + (long long)L * 8 - 1) / ((long long)L * 8);
  long long cap = ((long long)sm_count() * per_sm) / gy;  // one resident wave
  if (cap < 1) cap = 1;
  if (gx > cap) gx = cap;
  if (gx < 1) gx = 1;
  pointwise_kernel<OPS, IH, BNF><<<dim3((unsigned)gx, (unsigned)gy), 256, 0, stream>>>(P, (unsigned)npix, (unsigned)(h * w), (unsigned)w, c4, G);
  PMFB_LAUNCH_CHECK("pointwise_kernel");
  return PMFB_OK;
}

// ------------------------------------------------------------------------------------ pack_input
// one thread per destination pixel; reads are coalesced along x per source channel.
__global__ void __launch_bounds__(256)
pack_input_kernel(const float* __restrict__ src, long long s_n, long long s_c, long long s_h, long long s_w, int n, int c,
                  int h, int w, int n_shift, float* __restrict__ dst, int c_dst, long long dst_pix_stride, int round_out) {
  const long long total = (long long)n * h * w;
  const int half = n_shift / 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long p = i;
    const int x = (int)(p % w);
    p /= w;
    const int y = (int)(p % h);
    const int ni = (int)(p / h);
    const float* s0 = src + (long long)ni * s_n + (long long)y * s_h;
    float* d = dst + i * dst_pix_stride;
    for (int j0 = 0; j0 < c_dst; j0 += 4) {
      float v[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int j = j0 + k;
        float val = 0.f;
        if (j < n_shift * c) {
          const int s = j / c, ch = j - s * c;
          const int xs = x + s - half;
          if (xs >= 0 && xs < w) val = __ldg(s0 + (long long)ch * s_c + (long long)xs * s_w);
        }
        v[k] = round_out ? round_tf32(val) : val;
      }
      *reinterpret_cast<float4*>(d + j0) = make_float4(v[0], v[1], v[2], v[3]);
    }
  }
}

// ------------------------------------------------------------------------------------ NHWC -> NCHW
// 32 pixels x 32 channels tile through shared memory.
__global__ void __launch_bounds__(256)
nhwc_to_nchw_kernel(EpiView src, int n, int h, int w, int c, float* __restrict__ dst) {
  __shared__ float tile[32][33];
  const long long hw = (long long)h * w;
  const long long ptiles = (hw + 31) / 32;
  const int ctiles = (c + 31) / 32;
  const long long total = (long long)n * ptiles * ctiles;
  for (long long t = blockIdx.x; t < total; t += gridDim.x) {
    const int ct = (int)(t % ctiles);
    long long r = t / ctiles;
    const long long pt = r % ptiles;
    const int ni = (int)(r / ptiles);
    for (int k = threadIdx.x; k < 1024; k += 256) {
      const int pl = k >> 5, cl = k & 31;
      const long long p = pt * 32 + pl;
      const int ch = ct * 32 + cl;
      float v = 0.f;
      if (p < hw && ch < c) {
        const int y = (int)(p / w), x = (int)(p % w);
        v = __ldg(src.p + (long long)ni * src.sn + (long long)y * src.sy + (long long)x * src.sx + ch);
      }
      tile[pl][cl] = v;
    }
    __syncthreads();
    for (int k = threadIdx.x; k < 1024; k += 256) {
      const int cl = k >> 5, pl = k & 31;
      const long long p = pt * 32 + pl;
      const int ch = ct * 32 + cl;
      if (p < hw && ch < c) dst[((long long)ni * c + ch) * hw + p] = tile[pl][cl];
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------ weights
// OIHW -> packed [taps][c_out_p][c_in_p] (fwd) and [taps][c_in_p][c_out_p] (dgrad); pad entries are zero.
// stem: taps = kh, packed input channel j = kw_i*c_in + c (the horizontally unrolled 7x7 stem).
__global__ void pack_weight_kernel(const float* __restrict__ w, int c_out, int c_in, int kh, int kw, int stem, int c_out_p,
                                   int c_in_p, float* __restrict__ fwd, float* __restrict__ dgrad, int round_out) {
  const int taps = stem ? kh : kh * kw;
  const long long total = (long long)taps * c_out_p * c_in_p;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(i % c_in_p);
    long long r = i / c_in_p;
    const int co = (int)(r % c_out_p);
    const int t = (int)(r / c_out_p);
    float v = 0.f;
    if (co < c_out) {
      if (stem) {
        if (j < kw * c_in) {
          const int kj = j / c_in, ci = j - kj * c_in;
          v = w[(((long long)co * c_in + ci) * kh + t) * kw + kj];
        }
      } else if (j < c_in) {
        v = w[((long long)co * c_in + j) * taps + t];
      }
    }
    if (round_out) v = round_tf32(v);
    if (fwd) fwd[i] = v;
    if (dgrad) dgrad[((long long)t * c_in_p + j) * c_out_p + co] = v;
  }
}

__global__ void unpack_wgrad_kernel(const float* __restrict__ packed, int c_out, int c_in, int kh, int kw, int stem,
                                    int c_out_p, int c_in_p, float* __restrict__ grad, int accumulate) {
  const int taps = kh * kw;
  const long long total = (long long)c_out * c_in * taps;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int t = (int)(i % taps);
    long long r = i / taps;
    const int ci = (int)(r % c_in);
    const int co = (int)(r / c_in);
    float v;
    if (stem) {
      const int ki = t / kw, kj = t - ki * kw;
      v = packed[((long long)ki * c_in_p + (kj * c_in + ci)) * c_out_p + co];
    } else {
      v = packed[((long long)t * c_in_p + ci) * c_out_p + co];
    }
    grad[i] = accumulate ? grad[i] + v : v;
  }
}

__global__ void d2f_kernel(const double* __restrict__ src, float* __restrict__ dst, long long n, float scale, int accumulate,
                           int round_out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float v = (float)(src[i] * (double)scale);
    if (accumulate) v += dst[i];
    dst[i] = round_out ? round_tf32(v) : v;
  }
}

// ------------------------------------------------------------------------------------ 3x3 s2 p1 pooling
__global__ void __launch_bounds__(256)
pool3s2_kernel(int kind, EpiView xin, int n, int h, int w, int c4, const float* __restrict__ chan_scale, float* out,
               long long o_sn, long long o_sy, long long o_sx, uint8_t* idx, int round_out, unsigned short* out16,
               unsigned short* out16b) {
  const int ho = h / 2, wo = w / 2;
  const long long total = (long long)n * ho * wo * c4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int cg = (int)(i % c4);
    long long p = i / c4;
    const int xo = (int)(p % wo);
    p /= wo;
    const int yo = (int)(p % ho);
    const int ni = (int)(p / ho);
    const int c = cg * 4;
    float4 acc;
    uchar4 am = make_uchar4(0, 0, 0, 0);
    if (kind == 0) acc = make_float4(0.f, 0.f, 0.f, 0.f);
    else acc = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const int y = 2 * yo - 1 + t / 3, x = 2 * xo - 1 + t % 3;
      if (y < 0 || y >= h || x < 0 || x >= w) continue;
      const float4 v = ld4(xin.p + (long long)ni * xin.sn + (long long)y * xin.sy + (long long)x * xin.sx + c);
      if (kind == 0) {
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      } else {
        if (v.x > acc.x) { acc.x = v.x; am.x = t; }
        if (v.y > acc.y) { acc.y = v.y; am.y = t; }
        if (v.z > acc.z) { acc.z = v.z; am.z = t; }
        if (v.w > acc.w) { acc.w = v.w; am.w = t; }
      }
    }
    if (kind == 0) {
      acc.x /= 9.f; acc.y /= 9.f; acc.z /= 9.f; acc.w /= 9.f;
    } else if (idx) {
      *reinterpret_cast<uchar4*>(idx + (((long long)ni * ho + yo) * wo + xo) * (c4 * 4) + c) = am;
    }
    if (chan_scale) {
      const float4 sc = ld4(chan_scale + (long long)ni * (c4 * 4) + c);
      acc.x *= sc.x; acc.y *= sc.y; acc.z *= sc.z; acc.w *= sc.w;
    }
    if (round_out) acc = rnd4(acc);
    const long long off = (long long)ni * o_sn + (long long)yo * o_sy + (long long)xo * o_sx + c;
    *reinterpret_cast<float4*>(out + off) = acc;
    if (out16) *reinterpret_cast<uint2*>(out16 + off) = pack16(acc, PMFB_DT_F16);
    if (out16b) *reinterpret_cast<uint2*>(out16b + off) = pack16(acc, PMFB_DT_BF16);
  }
}

// gather form: every input pixel sums the (at most 4) output windows that cover it.  32-bit index arithmetic and all
// four candidate loads (plus the accumulate read) issued before the first use: the kernel is a pure HBM stream.
template <int KIND>
__global__ void __launch_bounds__(256, KIND == 0 ? 3 : 2)
pool3s2_bwd_kernel(EpiView dy, int n, int h, int w, int c4, const float* __restrict__ chan_scale, float* dx,
                   long long d_sn, long long d_sy, long long d_sx, const uint8_t* __restrict__ idx, int accumulate) {
  // Two (pixel, 4-channel group) items per thread and trip, every load of both issued before the first use: the kernel is a
  // gather (<= 4 pooled gradients per pixel, mostly L2 hits) plus one full-resolution read-modify-write stream, and with one
  // item per trip it ran latency-bound at 1.9 TB/s.
  constexpr int U = 2;
  const unsigned ho = h / 2, wo = w / 2;
  const unsigned total = (unsigned)n * h * w * c4;  // host guarantees < 2^32
  const unsigned step = gridDim.x * blockDim.x;
  for (unsigned i0 = blockIdx.x * blockDim.x + threadIdx.x; i0 < total; i0 += U * step) {
    float4 g[U][4], old[U], sc[U];
    uchar4 am[KIND != 0 ? U : 1][4];
    unsigned tk[KIND != 0 ? U : 1][4];
    float* dptr[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const unsigned i = i0 + u * step;
      dptr[u] = nullptr;
      if (i >= total) continue;
      const unsigned cg = i % (unsigned)c4;
      unsigned p = i / (unsigned)c4;
      const unsigned x = p % (unsigned)w;
      p /= (unsigned)w;
      const unsigned y = p % (unsigned)h;
      const unsigned ni = p / (unsigned)h;
      const int c = cg * 4;
      // windows: 2*yo-1 <= y <= 2*yo+1  ->  yo in {y/2, (y+1)/2}
      const unsigned yo[2] = {y / 2, (y + 1) / 2}, xo[2] = {x / 2, (x + 1) / 2};
      const bool vy[2] = {true, (y & 1u) && yo[1] < ho}, vx[2] = {true, (x & 1u) && xo[1] < wo};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int a = k >> 1, b = k & 1;
        g[u][k] = make_float4(0.f, 0.f, 0.f, 0.f);
        if constexpr (KIND != 0) {
          am[u][k] = make_uchar4(255, 255, 255, 255);
          tk[u][k] = (y - (2 * yo[a] - 1)) * 3 + (x - (2 * xo[b] - 1));
        }
        if (vy[a] && vx[b]) {
          g[u][k] = ld4(dy.p + (long long)ni * dy.sn + (long long)yo[a] * dy.sy + (long long)xo[b] * dy.sx + c);
          if constexpr (KIND != 0)
            am[u][k] = *reinterpret_cast<const uchar4*>(idx + (((long long)ni * ho + yo[a]) * wo + xo[b]) * (c4 * 4) + c);
        }
      }
      dptr[u] = dx + (long long)ni * d_sn + (long long)y * d_sy + (long long)x * d_sx + c;
      old[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (accumulate) old[u] = *reinterpret_cast<const float4*>(dptr[u]);
      sc[u] = make_float4(1.f, 1.f, 1.f, 1.f);
      if (chan_scale) sc[u] = ld4(chan_scale + (long long)ni * (c4 * 4) + c);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (!dptr[u]) continue;
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if constexpr (KIND == 0) {
          acc.x += g[u][k].x; acc.y += g[u][k].y; acc.z += g[u][k].z; acc.w += g[u][k].w;
        } else {
          const unsigned t = tk[u][k];
          if (am[u][k].x == t) acc.x += g[u][k].x;
          if (am[u][k].y == t) acc.y += g[u][k].y;
          if (am[u][k].z == t) acc.z += g[u][k].z;
          if (am[u][k].w == t) acc.w += g[u][k].w;
        }
      }
      if constexpr (KIND == 0) {
        acc.x /= 9.f; acc.y /= 9.f; acc.z /= 9.f; acc.w /= 9.f;
      }
      acc.x = acc.x * sc[u].x + old[u].x; acc.y = acc.y * sc[u].y + old[u].y;
      acc.z = acc.z * sc[u].z + old[u].z; acc.w = acc.w * sc[u].w + old[u].w;
      *reinterpret_cast<float4*>(dptr[u]) = acc;
    }
  }
}

// ------------------------------------------------------------------------------------ PixelShuffle(2)
// thread = (input pixel, group of 4 OUTPUT channels): reads 16 contiguous input channels, writes 4 pixels x float4.
__global__ void __launch_bounds__(256)
pixel_shuffle_kernel(EpiView xin, int n, int h, int w, int c4, const float* __restrict__ chan_scale, float* out,
                     long long o_sn, long long o_sy, long long o_sx, int round_out, unsigned short* out16, unsigned short* out16b) {
  const long long total = (long long)n * h * w * c4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int cg = (int)(i % c4);
    long long p = i / c4;
    const int x = (int)(p % w);
    p /= w;
    const int y = (int)(p % h);
    const int ni = (int)(p / h);
    const float* s = xin.p + (long long)ni * xin.sn + (long long)y * xin.sy + (long long)x * xin.sx + cg * 16;
    const float4 a = ld4(s), b = ld4(s + 4), cc = ld4(s + 8), d = ld4(s + 12);  // channel 4*(4cg+k) + 2i+j
    float4 sc = make_float4(1.f, 1.f, 1.f, 1.f);
    if (chan_scale) sc = ld4(chan_scale + (long long)ni * (c4 * 4) + cg * 4);
    float4 o[4];
    o[0] = make_float4(a.x * sc.x, b.x * sc.y, cc.x * sc.z, d.x * sc.w);
    o[1] = make_float4(a.y * sc.x, b.y * sc.y, cc.y * sc.z, d.y * sc.w);
    o[2] = make_float4(a.z * sc.x, b.z * sc.y, cc.z * sc.z, d.z * sc.w);
    o[3] = make_float4(a.w * sc.x, b.w * sc.y, cc.w * sc.z, d.w * sc.w);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int ii = q >> 1, jj = q & 1;
      float4 v = round_out ? rnd4(o[q]) : o[q];
      const long long off = (long long)ni * o_sn + (long long)(2 * y + ii) * o_sy + (long long)(2 * x + jj) * o_sx + cg * 4;
      *reinterpret_cast<float4*>(out + off) = v;
      if (out16) *reinterpret_cast<uint2*>(out16 + off) = pack16(v, PMFB_DT_F16);
      if (out16b) *reinterpret_cast<uint2*>(out16b + off) = pack16(v, PMFB_DT_BF16);
    }
  }
}

__global__ void __launch_bounds__(256)
pixel_shuffle_bwd_kernel(EpiView dy, int n, int h, int w, int c4, const float* __restrict__ chan_scale, float* dx,
                         long long d_sn, long long d_sy, long long d_sx, int accumulate, int round_out) {
  const long long total = (long long)n * h * w * c4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int cg = (int)(i % c4);
    long long p = i / c4;
    const int x = (int)(p % w);
    p /= w;
    const int y = (int)(p % h);
    const int ni = (int)(p / h);
    float4 g[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int ii = q >> 1, jj = q & 1;
      g[q] = ld4(dy.p + (long long)ni * dy.sn + (long long)(2 * y + ii) * dy.sy + (long long)(2 * x + jj) * dy.sx + cg * 4);
    }
    float4 sc = make_float4(1.f, 1.f, 1.f, 1.f);
    if (chan_scale) sc = ld4(chan_scale + (long long)ni * (c4 * 4) + cg * 4);
    float4 o[4];
    o[0] = make_float4(g[0].x * sc.x, g[1].x * sc.x, g[2].x * sc.x, g[3].x * sc.x);
    o[1] = make_float4(g[0].y * sc.y, g[1].y * sc.y, g[2].y * sc.y, g[3].y * sc.y);
    o[2] = make_float4(g[0].z * sc.z, g[1].z * sc.z, g[2].z * sc.z, g[3].z * sc.z);
    o[3] = make_float4(g[0].w * sc.w, g[1].w * sc.w, g[2].w * sc.w, g[3].w * sc.w);
    float* d = dx + (long long)ni * d_sn + (long long)y * d_sy + (long long)x * d_sx + cg * 16;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float4 v = o[k];
      if (accumulate) {
        const float4 e = *reinterpret_cast<const float4*>(d + 4 * k);
        v.x += e.x; v.y += e.y; v.z += e.z; v.w += e.w;
      }
      if (round_out) v = rnd4(v);
      *reinterpret_cast<float4*>(d + 4 * k) = v;
    }
  }
}

// ------------------------------------------------------------------------------------ bilinear x2 (align_corners=False)
__device__ __forceinline__ void up2_src(int o, int size, int& i0, int& i1, float& l1) {
  float s = 0.5f * (o + 0.5f) - 0.5f;
  if (s < 0.f) s = 0.f;
  i0 = (int)s;
  i1 = i0 + (i0 < size - 1 ? 1 : 0);
  l1 = s - (float)i0;
}

__global__ void __launch_bounds__(256)
upsample2x_kernel(EpiView xin, int n, int h, int w, int c4, unsigned short* out16, unsigned short* out16b, float* out, long long o_sn, long long o_sy, long long o_sx,
                  int round_out) {
  const int ho = 2 * h, wo = 2 * w;
  const long long total = (long long)n * ho * wo * c4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int cg = (int)(i % c4);
    long long p = i / c4;
    const int xo = (int)(p % wo);
    p /= wo;
    const int yo = (int)(p % ho);
    const int ni = (int)(p / ho);
    int y0, y1, x0, x1;
    float ly, lx;
    up2_src(yo, h, y0, y1, ly);
    up2_src(xo, w, x0, x1, lx);
    const float hy = 1.f - ly, hx = 1.f - lx;
    const float* b = xin.p + (long long)ni * xin.sn + cg * 4;
    const float4 v00 = ld4(b + (long long)y0 * xin.sy + (long long)x0 * xin.sx);
    const float4 v01 = ld4(b + (long long)y0 * xin.sy + (long long)x1 * xin.sx);
    const float4 v10 = ld4(b + (long long)y1 * xin.sy + (long long)x0 * xin.sx);
    const float4 v11 = ld4(b + (long long)y1 * xin.sy + (long long)x1 * xin.sx);
    float4 o;
    o.x = hy * (hx * v00.x + lx * v01.x) + ly * (hx * v10.x + lx * v11.x);
    o.y = hy * (hx * v00.y + lx * v01.y) + ly * (hx * v10.y + lx * v11.y);
    o.z = hy * (hx * v00.z + lx * v01.z) + ly * (hx * v10.z + lx * v11.z);
    o.w = hy * (hx * v00.w + lx * v01.w) + ly * (hx * v10.w + lx * v11.w);
    if (round_out) o = rnd4(o);
    const long long off = (long long)ni * o_sn + (long long)yo * o_sy + (long long)xo * o_sx + cg * 4;
    *reinterpret_cast<float4*>(out + off) = o;
    if (out16) *reinterpret_cast<uint2*>(out16 + off) = pack16(o, PMFB_DT_F16);
    if (out16b) *reinterpret_cast<uint2*>(out16b + off) = pack16(o, PMFB_DT_BF16);
  }
}

__device__ __forceinline__ float up2_weight(int o, int size, int i) {
  int i0, i1;
  float l1;
  up2_src(o, size, i0, i1, l1);
  float wgt = 0.f;
  if (i0 == i) wgt += 1.f - l1;
  if (i1 == i) wgt += l1;
  return wgt;
}

__global__ void __launch_bounds__(256)
upsample2x_bwd_kernel(EpiView dy, int n, int h, int w, int c4, float* dx, long long d_sn, long long d_sy, long long d_sx,
                      int accumulate) {
  const int ho = 2 * h, wo = 2 * w;
  const long long total = (long long)n * h * w * c4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int cg = (int)(i % c4);
    long long p = i / c4;
    const int x = (int)(p % w);
    p /= w;
    const int y = (int)(p % h);
    const int ni = (int)(p / h);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int yo = 2 * y - 1; yo <= 2 * y + 2; ++yo) {
      if (yo < 0 || yo >= ho) continue;
      const float wy = up2_weight(yo, h, y);
      if (wy == 0.f) continue;
      for (int xo = 2 * x - 1; xo <= 2 * x + 2; ++xo) {
        if (xo < 0 || xo >= wo) continue;
        const float wgt = wy * up2_weight(xo, w, x);
        if (wgt == 0.f) continue;
        const float4 g = ld4(dy.p + (long long)ni * dy.sn + (long long)yo * dy.sy + (long long)xo * dy.sx + cg * 4);
        acc.x += wgt * g.x; acc.y += wgt * g.y; acc.z += wgt * g.z; acc.w += wgt * g.w;
      }
    }
    float* d = dx + (long long)ni * d_sn + (long long)y * d_sy + (long long)x * d_sx + cg * 4;
    if (accumulate) {
      const float4 e = *reinterpret_cast<const float4*>(d);
      acc.x += e.x; acc.y += e.y; acc.z += e.z; acc.w += e.w;
    }
    *reinterpret_cast<float4*>(d) = acc;
  }
}

// ------------------------------------------------------------------------------------ softmax over channels
// one thread per pixel; NHWC logits in, dense NCHW probabilities out (plane writes are coalesced across the warp).
constexpr int kMaxSoftmaxC = 64;

__global__ void __launch_bounds__(256)
softmax_nchw_kernel(EpiView lg, int n, int h, int w, int c, float* __restrict__ out) {
  const long long hw = (long long)h * w;
  const long long total = (long long)n * hw;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ni = (int)(i / hw);
    const long long p = i - (long long)ni * hw;
    const int y = (int)(p / w), x = (int)(p % w);
    const float* s = lg.p + (long long)ni * lg.sn + (long long)y * lg.sy + (long long)x * lg.sx;
    float v[kMaxSoftmaxC];
    float m = -INFINITY;
#pragma unroll
    for (int j = 0; j < kMaxSoftmaxC; j += 4) {
      if (j < c) {  // the view is padded to a multiple of 4 channels; pad lanes are ignored
        const float4 t = ld4(s + j);
        v[j] = t.x; v[j + 1] = t.y; v[j + 2] = t.z; v[j + 3] = t.w;
      }
    }
#pragma unroll
    for (int j = 0; j < kMaxSoftmaxC; ++j)
      if (j < c) m = fmaxf(m, v[j]);
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < kMaxSoftmaxC; ++j) {
      if (j < c) {
        v[j] = expf(v[j] - m);
        sum += v[j];
      }
    }
    const float inv = 1.f / sum;
    float* o = out + (long long)ni * c * hw + p;
#pragma unroll
    for (int j = 0; j < kMaxSoftmaxC; ++j)
      if (j < c) o[(long long)j * hw] = v[j] * inv;
  }
}

// One thread per pixel: the NCHW reads are coalesced across the warp (consecutive pixels of one channel plane).  The
// NHWC result of a 256-pixel block is one contiguous run when the destination is dense (d_sx == cpad): it is staged
// through shared memory and written with fully coalesced stores (a per-thread float4 write at an 80-byte stride ran
// at 0.7 TB/s).  `dense` = 0 keeps the direct strided write for arbitrary views.
template <int CMAX>
__global__ void __launch_bounds__(256)
softmax_nchw_bwd_kernel(const float* __restrict__ pr, const float* __restrict__ dp, int n, int h, int w, int c, int cpad, float* dz,
                        long long d_sn, long long d_sy, long long d_sx, int round_out, int dense) {
  extern __shared__ float s_out[];  // 256 x (cpad | 1)
  const int pitch = cpad | 1;
  const long long hw = (long long)h * w;
  const long long total = (long long)n * hw;
  for (long long base = blockIdx.x * 256ll; base < total; base += (long long)gridDim.x * 256ll) {
    const long long i = base + threadIdx.x;
    if (i < total) {
      const int ni = (int)(i / hw);
      const long long p = i - (long long)ni * hw;
      const float* pp = pr + (long long)ni * c * hw + p;
      const float* gp = dp + (long long)ni * c * hw + p;
      float pv[CMAX], gv[CMAX];
      float dot = 0.f;
#pragma unroll
      for (int j = 0; j < CMAX; ++j) {
        if (j < c) {
          pv[j] = __ldg(pp + (long long)j * hw);
          gv[j] = __ldg(gp + (long long)j * hw);
          dot += pv[j] * gv[j];
        }
      }
      if (dense) {
#pragma unroll
        for (int j = 0; j < CMAX; ++j) {
          if (j < cpad) {
            float o = j < c ? pv[j] * (gv[j] - dot) : 0.f;
            if (round_out) o = round_tf32(o);
            s_out[threadIdx.x * pitch + j] = o;
          }
        }
      } else {
        const int y = (int)(p / w), x = (int)(p % w);
        float* d = dz + (long long)ni * d_sn + (long long)y * d_sy + (long long)x * d_sx;
#pragma unroll
        for (int j = 0; j < CMAX; j += 4) {
          if (j < c) {
            float4 o = make_float4(pv[j] * (gv[j] - dot), j + 1 < c ? pv[j + 1] * (gv[j + 1] - dot) : 0.f,
                                   j + 2 < c ? pv[j + 2] * (gv[j + 2] - dot) : 0.f,
                                   j + 3 < c ? pv[j + 3] * (gv[j + 3] - dot) : 0.f);
            if (round_out) o = rnd4(o);
            *reinterpret_cast<float4*>(d + j) = o;
          }
        }
      }
    }
    if (dense) {
      __syncthreads();
      const long long left = total - base;
      const int npx = left < 256 ? (int)left : 256;
      float* d = dz + base * cpad;
      for (int k = threadIdx.x; k < npx * cpad; k += 256) {
        const int px = k / cpad, j = k - px * cpad;
        d[k] = s_out[px * pitch + j];
      }
      __syncthreads();
    }
  }
}

static inline EpiView ev(const pmfb_view* v) {
  EpiView e;
  e.p = v ? v->ptr : nullptr;
  e.sn = v ? v->sn : 0;
  e.sy = v ? v->sy : 0;
  e.sx = v ? v->sx : 0;
  return e;
}

}  // namespace pmfb

using namespace pmfb;

#define REQ(cond, ...) \
  do {                 \
    if (!(cond)) return fail(PMFB_ERR_INVALID, __VA_ARGS__); \
  } while (0)

extern "C" int pmfb_memset_zero(void* ptr, size_t bytes, void* stream) {
  REQ(ptr || bytes == 0, "memset_zero: null pointer");
  if (bytes == 0) return PMFB_OK;
  PMFB_CUDA_CHECK(cudaMemsetAsync(ptr, 0, bytes, (cudaStream_t)stream));
  return PMFB_OK;
}

static inline PWV pwv(const float* p, long long sn, long long sy, long long sx, int h, int w) {
  PWV o;
  o.p = p;
  o.sn = sn;
  o.sy = sy;
  o.sx = sx;
  o.linear = (p && sy == (long long)w * sx && sn == (long long)h * sy) ? 1 : 0;
  return o;
}

extern "C" int pmfb_pointwise(const pmfb_view* in, float* out, int64_t o_sn, int64_t o_sy, int64_t o_sx, int32_t n,
                              int32_t h, int32_t w, int32_t c, const pmfb_epilogue* epi, void* stream) {
  return pmfb_pointwise16(in, out, o_sn, o_sy, o_sx, n, h, w, c, epi, nullptr, PMFB_DT_F16, nullptr, 0, stream);
}

extern "C" int pmfb_pointwise16(const pmfb_view* in, float* out, int64_t o_sn, int64_t o_sy, int64_t o_sx, int32_t n,
                                int32_t h, int32_t w, int32_t c, const pmfb_epilogue* epi, void* out16, int32_t dtype16,
                                void* out16_bf16, int32_t in_half, void* stream) {
  return pmfb_pointwise16_bn(in, out, o_sn, o_sy, o_sx, n, h, w, c, epi, out16, dtype16, out16_bf16, in_half, nullptr, stream);
}

extern "C" int pmfb_pointwise16_bn(const pmfb_view* in, float* out, int64_t o_sn, int64_t o_sy, int64_t o_sx, int32_t n,
                                   int32_t h, int32_t w, int32_t c, const pmfb_epilogue* epi, void* out16, int32_t dtype16,
                                   void* out16_bf16, int32_t in_half, const pmfb_bn_fuse* bn, void* stream) {
  REQ(!bn || (in_half && bn->sums && bn->count > 0 && bn->alpha_out && bn->beta_out && bn->mean_out && bn->invstd_out && epi &&
              !epi->alpha1 && !epi->beta1),
      "pointwise16_bn: the fused finalisation needs in_half, sums, count > 0, the four output vectors and NULL alpha1 / beta1");
  REQ(!bn || ((reinterpret_cast<uintptr_t>(bn->alpha_out) | reinterpret_cast<uintptr_t>(bn->beta_out) |
               reinterpret_cast<uintptr_t>(bn->mean_out) | reinterpret_cast<uintptr_t>(bn->invstd_out)) & 15) == 0,
      "pointwise16_bn: output vectors must be 16-byte aligned");
  REQ(!in_half || (in && in->ptr), "pointwise16: in_half needs an input view");
  REQ(!out16_bf16 || (reinterpret_cast<uintptr_t>(out16_bf16) & 7) == 0, "pointwise16: the bf16 output must be 8-byte aligned");
  REQ(epi && c > 0 && c % 4 == 0, "pointwise: c=%d must be a positive multiple of 4", c);
  REQ(!out16 || ((dtype16 == PMFB_DT_F16 || dtype16 == PMFB_DT_BF16) && (reinterpret_cast<uintptr_t>(out16) & 7) == 0),
      "pointwise16: the 16-bit output must be 8-byte aligned, dtype F16 or BF16");
  REQ(out_ok(out, o_sn, o_sy, o_sx), "pointwise: bad output view");
  REQ(!in || view_ok(in), "pointwise: bad input view");
  EpiParams E;
  int rc = epi_from_c(epi, &E);
  if (rc) return rc;
  const long long npix = (long long)n * h * w;
  if (npix == 0) return PMFB_OK;
  REQ(npix < (1ll << 31), "pointwise: too many pixels");
  PWParams P;
  P.bn = pmfb_bn_fuse{};
  P.in = in ? pwv(in->ptr, in->sn, in->sy, in->sx, h, w) : pwv(nullptr, 0, 0, 0, h, w);
  P.out = pwv(out, o_sn, o_sy, o_sx, h, w);
  P.r1 = pwv(E.r1.p, E.r1.sn, E.r1.sy, E.r1.sx, h, w);
  P.mul = pwv(E.mul.p, E.mul.sn, E.mul.sy, E.mul.sx, h, w);
  P.r2 = pwv(E.r2.p, E.r2.sn, E.r2.sy, E.r2.sx, h, w);
  P.alpha1 = E.alpha1;
  P.beta1 = E.beta1;
  P.alpha2 = E.alpha2;
  P.beta2 = E.beta2;
  P.act = E.act;
  P.round_out = E.round_out;
  P.out16 = static_cast<unsigned short*>(out16);
  P.dt16 = dtype16;
  P.out16b = static_cast<unsigned short*>(out16_bf16);
  const int ops = (P.r1.p ? 1 : 0) | (P.mul.p ? 2 : 0) | (P.r2.p ? 4 : 0);
  const cudaStream_t st = (cudaStream_t)stream;
  REQ(!in_half || (P.in.linear && P.out.linear && (!P.r1.p || P.r1.linear) && (!P.r2.p || P.r2.linear)),
      "pointwise16: in_half needs dense (pixel-linear) in / out / r1 / r2 views");
  if (bn) {
    P.bn = *bn;
#define PMFB_PWB(o) case o: return launch_pointwise_t<o, true, true>(P, npix, h, w, c / 4, st);
    switch (ops) {
      PMFB_PWB(0) PMFB_PWB(1) PMFB_PWB(2) PMFB_PWB(3) PMFB_PWB(4) PMFB_PWB(5) PMFB_PWB(6)
      default: return launch_pointwise_t<7, true, true>(P, npix, h, w, c / 4, st);
    }
#undef PMFB_PWB
  }
#define PMFB_PW(o) case o: return in_half ? launch_pointwise_t<o, true>(P, npix, h, w, c / 4, st) : launch_pointwise_t<o, false>(P, npix, h, w, c / 4, st);
  switch (ops) {
    PMFB_PW(0) PMFB_PW(1) PMFB_PW(2) PMFB_PW(3) PMFB_PW(4) PMFB_PW(5) PMFB_PW(6)
    default: return in_half ? launch_pointwise_t<7, true>(P, npix, h, w, c / 4, st) : launch_pointwise_t<7, false>(P, npix, h, w, c / 4, st);
  }
#undef PMFB_PW
}

extern "C" int pmfb_pack_input(const float* src, int64_t s_n, int64_t s_c, int64_t s_h, int64_t s_w, int32_t n, int32_t c,
                               int32_t h, int32_t w, int32_t n_shift, float* dst, int32_t c_dst, int64_t dst_pix_stride,
                               int32_t round_out, void* stream) {
  REQ(src && dst && c > 0 && n_shift >= 1 && (n_shift & 1), "pack_input: bad arguments");
  REQ(c_dst % 4 == 0 && c_dst >= n_shift * c, "pack_input: c_dst=%d must be a multiple of 4 and >= %d", c_dst, n_shift * c);
  REQ((reinterpret_cast<uintptr_t>(dst) & 15) == 0 && dst_pix_stride >= c_dst && dst_pix_stride % 4 == 0,
      "pack_input: dst must be 16-byte aligned with a pixel stride >= c_dst that is a multiple of 4");
  const long long total = (long long)n * h * w;
  if (total == 0) return PMFB_OK;
  pack_input_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(src, s_n, s_c, s_h, s_w, n, c, h, w, n_shift, dst,
                                                                            c_dst, dst_pix_stride, round_out);
  PMFB_LAUNCH_CHECK("pack_input_kernel");
  return PMFB_OK;
}

extern "C" int pmfb_nhwc_to_nchw(const pmfb_view* src, int32_t n, int32_t h, int32_t w, int32_t c, float* dst, void* stream) {
  REQ(src && src->ptr && dst && c > 0, "nhwc_to_nchw: bad arguments");
  const long long tiles = (long long)n * (((long long)h * w + 31) / 32) * ((c + 31) / 32);
  if (tiles == 0) return PMFB_OK;
  nhwc_to_nchw_kernel<<<(int)(tiles < sm_count() * 16 ? tiles : sm_count() * 16), 256, 0, (cudaStream_t)stream>>>(ev(src), n, h, w, c, dst);
  PMFB_LAUNCH_CHECK("nhwc_to_nchw_kernel");
  return PMFB_OK;
}

extern "C" int pmfb_pack_weight(const float* w, int32_t c_out, int32_t c_in, int32_t kh, int32_t kw, int32_t stem,
                                int32_t c_out_p, int32_t c_in_p, float* fwd, float* dgrad, int32_t round_out, void* stream) {
  REQ(w && (fwd || dgrad) && c_out > 0 && c_in > 0 && kh > 0 && kw > 0, "pack_weight: bad arguments");
  REQ(c_out_p >= c_out && c_out_p % 4 == 0 && c_in_p % 4 == 0 && c_in_p >= (stem ? kw * c_in : c_in),
      "pack_weight: padded dims (%d,%d) too small or not multiples of 4", c_out_p, c_in_p);
  const long long total = (long long)(stem ? kh : kh * kw) * c_out_p * c_in_p;
  pack_weight_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(w, c_out, c_in, kh, kw, stem, c_out_p, c_in_p, fwd,
                                                                             dgrad, round_out);
  PMFB_LAUNCH_CHECK("pack_weight_kernel");
  return PMFB_OK;
}

extern "C" int pmfb_unpack_wgrad(const float* packed, int32_t c_out, int32_t c_in, int32_t kh, int32_t kw, int32_t stem,
                                 int32_t c_out_p, int32_t c_in_p, float* grad, int32_t accumulate, void* stream) {
  REQ(packed && grad && c_out > 0 && c_in > 0, "unpack_wgrad: bad arguments");
  REQ(c_out_p >= c_out && c_in_p >= (stem ? kw * c_in : c_in), "unpack_wgrad: padded dims too small");
  const long long total = (long long)c_out * c_in * kh * kw;
  unpack_wgrad_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(packed, c_out, c_in, kh, kw, stem, c_out_p, c_in_p,
                                                                              grad, accumulate);
  PMFB_LAUNCH_CHECK("unpack_wgrad_kernel");
  return PMFB_OK;
}

// ------------------------------------------------------------------------------------ 16-bit shadows
namespace pmfb {
__global__ void __launch_bounds__(256)
convert16_kernel(EpiView x, int n, int h, int w, int c4, unsigned short* __restrict__ out, long long o_sn, long long o_sy,
                 long long o_sx, int dt) {
  const long long total = (long long)n * h * w * c4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int q = (int)(i % c4);
    long long r = i / c4;
    const int xx = (int)(r % w);
    r /= w;
    const int yy = (int)(r % h);
    const int nn = (int)(r / h);
    const float4 v = ld4(x.p + (long long)nn * x.sn + (long long)yy * x.sy + (long long)xx * x.sx + 4 * q);
    *reinterpret_cast<uint2*>(out + (long long)nn * o_sn + (long long)yy * o_sy + (long long)xx * o_sx + 4 * q) = pack16(v, dt);
  }
}
}  // namespace pmfb

extern "C" int pmfb_convert16(const pmfb_view* in, int32_t n, int32_t h, int32_t w, int32_t c, void* out16, int64_t o_sn,
                              int64_t o_sy, int64_t o_sx, int32_t dtype16, void* stream) {
  REQ(in && in->ptr && out16 && c > 0 && c % 4 == 0, "convert16: bad arguments");
  REQ(dtype16 == PMFB_DT_F16 || dtype16 == PMFB_DT_BF16, "convert16: dtype must be F16 or BF16");
  REQ(view_ok(in) && ((o_sn | o_sy | o_sx) % 4) == 0 && (reinterpret_cast<uintptr_t>(out16) & 7) == 0, "convert16: bad views");
  const long long total = (long long)n * h * w * (c / 4);
  if (total == 0) return PMFB_OK;
  convert16_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(ev(in), n, h, w, c / 4, static_cast<unsigned short*>(out16),
                                                                          o_sn, o_sy, o_sx, dtype16);
  PMFB_LAUNCH_CHECK("convert16_kernel");
  return PMFB_OK;
}

// ------------------------------------------------------------------------------------ 3xTF32 operand split
// Precise mode (PMFB_PRECISION=3xtf32): x = hi + lo with hi = rna_tf32(x), lo = rna_tf32(x - hi) (|x - hi - lo| <= 2^-22 |x|).
// A tf32 x tf32 product is exact in the fp32 accumulator, so  x*w ~= hi_x*hi_w + hi_x*lo_w + lo_x*hi_w  restores ~21 bits
// of operand mantissa with THREE kind::tf32 UMMAs per K step into the same TMEM accumulator.  The three products are laid
// out along the GEMM K dimension (channels): an activation becomes [hi | hi | lo], a weight [hi | lo | hi], each part
// zero-padded to a multiple of 32 channels (one TMA slab), and the unmodified implicit-GEMM kernels walk 3x the slabs.
namespace pmfb {
__global__ void split_tf32_kernel(EpiView x, int n, int h, int w, int c, int cp, int mode, float* __restrict__ out, long long o_sn,
                                  long long o_sy, long long o_sx) {
  const int parts = mode < 2 ? 3 : 1;
  const int cq = (mode < 2 ? cp : c) >> 2;  // float4 groups per part
  const long long total = (long long)n * h * w * parts * cq;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int q = (int)(i % cq);
    long long r = i / cq;
    const int part = (int)(r % parts);
    r /= parts;
    const int xx = (int)(r % w);
    r /= w;
    const int yy = (int)(r % h);
    const int nn = (int)(r / h);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (4 * q < c) v = ld4(x.p + (long long)nn * x.sn + (long long)yy * x.sy + (long long)xx * x.sx + 4 * q);
    // which half this part holds: mode 0 [hi|hi|lo], mode 1 [hi|lo|hi], mode 2 hi, mode 3 lo
    const bool lo = (mode == 0 && part == 2) || (mode == 1 && part == 1) || mode == 3;
    float4 hi4 = make_float4(round_tf32(v.x), round_tf32(v.y), round_tf32(v.z), round_tf32(v.w));
    float4 o = hi4;
    if (lo) o = make_float4(round_tf32(v.x - hi4.x), round_tf32(v.y - hi4.y), round_tf32(v.z - hi4.z), round_tf32(v.w - hi4.w));
    *reinterpret_cast<float4*>(out + (long long)nn * o_sn + (long long)yy * o_sy + (long long)xx * o_sx + (long long)part * cp + 4 * q) = o;
  }
}
}  // namespace pmfb

extern "C" int pmfb_split_tf32(const pmfb_view* in, int32_t n, int32_t h, int32_t w, int32_t c, float* out, int64_t o_sn,
                               int64_t o_sy, int64_t o_sx, int32_t mode, void* stream) {
  REQ(in && in->ptr && out && c > 0 && c % 4 == 0 && mode >= 0 && mode <= 3, "split_tf32: bad arguments");
  REQ(((in->sn | in->sy | in->sx | o_sn | o_sy | o_sx) % 4) == 0 && (reinterpret_cast<uintptr_t>(in->ptr) & 15) == 0 &&
          (reinterpret_cast<uintptr_t>(out) & 15) == 0, "split_tf32: views must be 16-byte aligned with strides multiple of 4");
  const int cp = (c + 31) & ~31;
  const long long total = (long long)n * h * w * (mode < 2 ? 3 * (cp / 4) : c / 4);
  if (total == 0) return PMFB_OK;
  split_tf32_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(ev(in), n, h, w, c, cp, mode, out, o_sn, o_sy, o_sx);
  PMFB_LAUNCH_CHECK("split_tf32_kernel");
  return PMFB_OK;
}

// ------------------------------------------------------------------------------------ EPMF sparse-conv mask ops
// epmf_net.py:67  mask = x.abs().sum(1).ne(0)            -> pixel_mask_kernel
// epmf_net.py:43  mask = MaxPool2d(k, stride, 0, dil)(F.pad(mask, pad))  (zero padding)  -> mask_maxpool_kernel
// epmf_net.py:31,49,69-82  x*mask, LeakyReLU, BatchNorm (eval affine), +shortcut, *mask  -> pixel_scale_kernel
namespace pmfb {

__global__ void __launch_bounds__(256)
pixel_mask_kernel(EpiView x, int n, int h, int w, int c4, float* __restrict__ mask) {
  const long long total = (long long)n * h * w;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long p = i;
    const int xx = (int)(p % w);
    p /= w;
    const int yy = (int)(p % h);
    const int ni = (int)(p / h);
    const float* s = x.p + (long long)ni * x.sn + (long long)yy * x.sy + (long long)xx * x.sx;
    float acc = 0.f;
    for (int g = 0; g < c4; ++g) {
      const float4 v = ld4(s + 4 * g);
      acc += fabsf(v.x) + fabsf(v.y) + fabsf(v.z) + fabsf(v.w);
    }
    mask[i] = acc != 0.f ? 1.f : 0.f;
  }
}

__global__ void __launch_bounds__(256)
mask_maxpool_kernel(const float* __restrict__ in, int n, int h, int w, int k, int stride, int dil, int pad, int oh, int ow,
                    float* __restrict__ out) {
  const long long total = (long long)n * oh * ow;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long p = i;
    const int ox = (int)(p % ow);
    p /= ow;
    const int oy = (int)(p % oh);
    const int ni = (int)(p / oh);
    float m = 0.f;  // the zero padding takes part in the maximum; masks are >= 0
    bool any = false;
    for (int a = 0; a < k; ++a) {
      const int yy = oy * stride - pad + a * dil;
      for (int b = 0; b < k; ++b) {
        const int xx = ox * stride - pad + b * dil;
        const float v = (yy >= 0 && yy < h && xx >= 0 && xx < w) ? in[((long long)ni * h + yy) * w + xx] : 0.f;
        m = any ? fmaxf(m, v) : v;
        any = true;
      }
    }
    out[i] = m;
  }
}

// out = (act(in * pre[p]) * alpha[c] + beta[c] + r) * post[p], optional tf32 rounding; thread = (pixel, float4 group).
// 32-bit indexing; pixel-linear views (offset = pixel * sx: dense tensors and channel slices of concat buffers) skip the
// (n, y, x) decode.
__device__ __forceinline__ long long px_off(const EpiView& v, int linear, unsigned pix, unsigned hw, unsigned w) {
  if (linear) return (long long)pix * v.sx;
  const unsigned n = pix / hw, q = pix - n * hw, y = q / w, x = q - y * w;
  return (long long)n * v.sn + (long long)y * v.sy + (long long)x * v.sx;
}

__global__ void __launch_bounds__(256)
pixel_scale_kernel(EpiView in, int lin_in, unsigned npix, unsigned hw, unsigned w, unsigned c4, const float* __restrict__ pre, int act,
                   const float* __restrict__ alpha, const float* __restrict__ beta, EpiView r, int lin_r,
                   const float* __restrict__ post, EpiView out, int lin_out, int round_out) {
  const unsigned total = npix * c4;  // host guarantees < 2^32
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const unsigned pix = i / c4;
    const int c = 4 * (int)(i - pix * c4);
    float4 v = ld4(in.p + px_off(in, lin_in, pix, hw, w) + c);
    float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r.p) q = ld4(r.p + px_off(r, lin_r, pix, hw, w) + c);
    if (pre) { const float m = __ldg(pre + pix); v.x *= m; v.y *= m; v.z *= m; v.w *= m; }
    if (act) { v.x = epi_act(act, v.x); v.y = epi_act(act, v.y); v.z = epi_act(act, v.z); v.w = epi_act(act, v.w); }
    if (alpha) { const float4 a = ld4(alpha + c); v.x *= a.x; v.y *= a.y; v.z *= a.z; v.w *= a.w; }
    if (beta) { const float4 b = ld4(beta + c); v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w; }
    v.x += q.x; v.y += q.y; v.z += q.z; v.w += q.w;
    if (post) { const float m = __ldg(post + pix); v.x *= m; v.y *= m; v.z *= m; v.w *= m; }
    if (round_out) v = rnd4(v);
    *reinterpret_cast<float4*>(const_cast<float*>(out.p) + px_off(out, lin_out, pix, hw, w) + c) = v;
  }
}

}  // namespace pmfb

extern "C" int pmfb_pixel_mask(const pmfb_view* x, int32_t n, int32_t h, int32_t w, int32_t c, float* mask, void* stream) {
  REQ(x && x->ptr && view_ok(x) && mask && c > 0 && c % 4 == 0, "pixel_mask: bad arguments (c=%d)", c);
  const long long total = (long long)n * h * w;
  if (total == 0) return PMFB_OK;
  pixel_mask_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(ev(x), n, h, w, c / 4, mask);
  PMFB_LAUNCH_CHECK("pixel_mask_kernel");
  return PMFB_OK;
}

extern "C" int pmfb_mask_maxpool(const float* mask_in, int32_t n, int32_t h, int32_t w, int32_t k, int32_t stride, int32_t dilation,
                                 int32_t pad, float* mask_out, void* stream) {
  REQ(mask_in && mask_out && k >= 1 && stride >= 1 && dilation >= 1 && pad >= 0, "mask_maxpool: bad arguments");
  const int oh = (h + 2 * pad - dilation * (k - 1) - 1) / stride + 1, ow = (w + 2 * pad - dilation * (k - 1) - 1) / stride + 1;
  REQ(oh > 0 && ow > 0, "mask_maxpool: empty output");
  const long long total = (long long)n * oh * ow;
  if (total == 0) return PMFB_OK;
  mask_maxpool_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(mask_in, n, h, w, k, stride, dilation, pad, oh, ow,
                                                                              mask_out);
  PMFB_LAUNCH_CHECK("mask_maxpool_kernel");
  return PMFB_OK;
}

extern "C" int pmfb_pixel_scale(const pmfb_view* in, int32_t n, int32_t h, int32_t w, int32_t c, const float* pre_mask, int32_t act,
                                const float* alpha, const float* beta, const pmfb_view* r, const float* post_mask, float* out,
                                int64_t o_sn, int64_t o_sy, int64_t o_sx, int32_t round_out, void* stream) {
  REQ(in && in->ptr && view_ok(in) && out_ok(out, o_sn, o_sy, o_sx) && c > 0 && c % 4 == 0, "pixel_scale: bad arguments (c=%d)", c);
  REQ(!r || view_ok(r), "pixel_scale: bad residual view");
  REQ(act >= 0 && act <= 3, "pixel_scale: act=%d", act);
  const long long total = (long long)n * h * w * (c / 4);
  if (total == 0) return PMFB_OK;
  REQ(total < (1ll << 32), "pixel_scale: tensor too large for 32-bit indexing");
  auto lin = [&](long long sn, long long sy, long long sx) { return (sy == (long long)w * sx && sn == (long long)h * sy) ? 1 : 0; };
  EpiView vo;
  vo.p = out;
  vo.sn = o_sn;
  vo.sy = o_sy;
  vo.sx = o_sx;
  const EpiView vr = (r && r->ptr) ? ev(r) : ev(nullptr);
  pixel_scale_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(
      ev(in), lin(in->sn, in->sy, in->sx), (unsigned)((long long)n * h * w), (unsigned)(h * w), (unsigned)w, (unsigned)(c / 4), pre_mask,
      act, alpha, beta, vr, vr.p ? lin(vr.sn, vr.sy, vr.sx) : 0, post_mask, vo, lin(o_sn, o_sy, o_sx), round_out);
  PMFB_LAUNCH_CHECK("pixel_scale_kernel");
  return PMFB_OK;
}

// ------------------------------------------------------------------------------------ batched weight jobs
// One launch packs (or unpacks) EVERY convolution weight of the network: a device-resident table of jobs, each owning
// the global work range [start, next start).  A training step issued 110 pack + 110 unpack + 110 memset launches of a
// few microseconds each (~3 ms of launch-bound time per step inside the CUDA graph).
namespace pmfb {

__device__ __forceinline__ int find_job(const pmfb_weight_job* __restrict__ jobs, int n, long long i, int hint) {
  int j = hint;
  while (j + 1 < n && i >= jobs[j + 1].start) ++j;
  return j;
}

template <bool UNPACK>
__global__ void __launch_bounds__(256)
weight_jobs_kernel(const pmfb_weight_job* __restrict__ jobs, int n_jobs, long long total) {
  constexpr int kChunk = 256 * 8;
  __shared__ int s_first;
  for (long long base = (long long)blockIdx.x * kChunk; base < total; base += (long long)gridDim.x * kChunk) {
    if (threadIdx.x == 0) {  // binary search: last job with start <= base
      int lo = 0, hi = n_jobs - 1;
      while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (jobs[mid].start <= base) lo = mid; else hi = mid - 1;
      }
      s_first = lo;
    }
    __syncthreads();
    int j = s_first, cur = -1;
    pmfb_weight_job J;
    for (int k = 0; k < 8; ++k) {
      const long long g = base + k * 256 + threadIdx.x;
      if (g >= total) break;
      j = find_job(jobs, n_jobs, g, j);
      if (j != cur) {
        J = jobs[j];
        cur = j;
      }
      const long long i = g - J.start;
      const int taps_p = J.stem ? J.kh : J.kh * J.kw;
      if (!UNPACK) {
        const int jj = (int)(i % J.c_in_p);
        long long r = i / J.c_in_p;
        const int co = (int)(r % J.c_out_p);
        const int t = (int)(r / J.c_out_p);
        float v = 0.f;
        if (co < J.c_out) {
          if (J.stem) {
            if (jj < J.kw * J.c_in) {
              const int kj = jj / J.c_in, ci = jj - kj * J.c_in;
              v = J.src[(((long long)co * J.c_in + ci) * J.kh + t) * J.kw + kj];
            }
          } else if (jj < J.c_in) {
            v = J.src[((long long)co * J.c_in + jj) * taps_p + t];
          }
        }
        if (!J.no_round) v = round_tf32(v);
        if (J.dst) J.dst[i] = v;
        if (J.dst2) J.dst2[((long long)t * J.c_in_p + jj) * J.c_out_p + co] = v;
      } else {
        const int taps = J.kh * J.kw;
        const int t = (int)(i % taps);
        long long r = i / taps;
        const int ci = (int)(r % J.c_in);
        const int co = (int)(r / J.c_in);
        float v;
        if (J.stem) {
          const int ki = t / J.kw, kj = t - ki * J.kw;
          v = J.src[((long long)ki * J.c_in_p + (kj * J.c_in + ci)) * J.c_out_p + co];
        } else {
          v = J.src[((long long)t * J.c_in_p + ci) * J.c_out_p + co];
        }
        J.dst[i] = J.accumulate ? J.dst[i] + v : v;
      }
    }
    __syncthreads();
  }
}

}  // namespace pmfb

extern "C" int pmfb_weight_jobs(int32_t unpack, const pmfb_weight_job* jobs_device, int32_t n_jobs, int64_t total_work,
                                void* stream) {
  REQ(jobs_device && n_jobs > 0 && total_work >= 0, "weight_jobs: bad arguments");
  if (total_work == 0) return PMFB_OK;
  const int grid = grid_for((total_work + 7) / 8, 256);
  if (unpack)
    weight_jobs_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(jobs_device, n_jobs, total_work);
  else
    weight_jobs_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(jobs_device, n_jobs, total_work);
  PMFB_LAUNCH_CHECK("weight_jobs_kernel");
  return PMFB_OK;
}

extern "C" int pmfb_d2f(const double* src, float* dst, int64_t n, float scale, int32_t accumulate, int32_t round_out,
                        void* stream) {
  REQ(src && dst && n >= 0, "d2f: bad arguments");
  if (n == 0) return PMFB_OK;
  d2f_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(src, dst, n, scale, accumulate, round_out);
  PMFB_LAUNCH_CHECK("d2f_kernel");
  return PMFB_OK;
}

extern "C" int pmfb_pool3s2(int32_t kind, const pmfb_view* x, int32_t n, int32_t h, int32_t w, int32_t c,
                            const float* chan_scale, float* out, int64_t o_sn, int64_t o_sy, int64_t o_sx, uint8_t* idx,
                            int32_t round_out, void* out16, void* out16_bf16, void* stream) {
  REQ(x && x->ptr && view_ok(x) && out_ok(out, o_sn, o_sy, o_sx), "pool3s2: bad views");
  REQ(((reinterpret_cast<uintptr_t>(out16) | reinterpret_cast<uintptr_t>(out16_bf16)) & 7) == 0, "pool3s2: 16-bit outputs must be 8-byte aligned");
  REQ(c % 4 == 0 && h % 2 == 0 && w % 2 == 0 && (kind == 0 || kind == 1), "pool3s2: c%%4, even h/w, kind in {0,1}");
  const long long total = (long long)n * (h / 2) * (w / 2) * (c / 4);
  if (total == 0) return PMFB_OK;
  pool3s2_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(kind, ev(x), n, h, w, c / 4, chan_scale, out, o_sn, o_sy,
                                                                         o_sx, idx, round_out, static_cast<unsigned short*>(out16),
                                                                         static_cast<unsigned short*>(out16_bf16));
  PMFB_LAUNCH_CHECK("pool3s2_kernel");
  return PMFB_OK;
}

extern "C" int pmfb_pool3s2_bwd(int32_t kind, const pmfb_view* dy, int32_t n, int32_t h, int32_t w, int32_t c,
                                const float* chan_scale, float* dx, int64_t d_sn, int64_t d_sy, int64_t d_sx,
                                const uint8_t* idx, int32_t accumulate, void* stream) {
  REQ(dy && dy->ptr && view_ok(dy) && out_ok(dx, d_sn, d_sy, d_sx), "pool3s2_bwd: bad views");
  REQ(c % 4 == 0 && h % 2 == 0 && w % 2 == 0 && (kind == 0 || (kind == 1 && idx)), "pool3s2_bwd: bad arguments");
  const long long total = (long long)n * h * w * (c / 4);
  if (total == 0) return PMFB_OK;
  REQ(total < (1ll << 32), "pool3s2_bwd: tensor too large for 32-bit indexing");
  if (kind == 0)
    pool3s2_bwd_kernel<0><<<grid_for(total, 512), 256, 0, (cudaStream_t)stream>>>(ev(dy), n, h, w, c / 4, chan_scale, dx, d_sn, d_sy, d_sx,
                                                                                  idx, accumulate);
  else
    pool3s2_bwd_kernel<1><<<grid_for(total, 512), 256, 0, (cudaStream_t)stream>>>(ev(dy), n, h, w, c / 4, chan_scale, dx, d_sn, d_sy, d_sx,
                                                                                  idx, accumulate);
  PMFB_LAUNCH_CHECK("pool3s2_bwd_kernel");
  return PMFB_OK;
}

extern "C" int pmfb_pixel_shuffle(const pmfb_view* x, int32_t n, int32_t h, int32_t w, int32_t c, const float* chan_scale,
                                  float* out, int64_t o_sn, int64_t o_sy, int64_t o_sx, int32_t round_out, void* out16,
                                  void* out16_bf16, void* stream) {
  REQ(x && x->ptr && view_ok(x) && out_ok(out, o_sn, o_sy, o_sx) && c % 4 == 0, "pixel_shuffle: bad arguments");
  REQ(((reinterpret_cast<uintptr_t>(out16) | reinterpret_cast<uintptr_t>(out16_bf16)) & 7) == 0, "pixel_shuffle: 16-bit outputs must be 8-byte aligned");
  const long long total = (long long)n * h * w * (c / 4);
  if (total == 0) return PMFB_OK;
  pixel_shuffle_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(ev(x), n, h, w, c / 4, chan_scale, out, o_sn, o_sy,
                                                                               o_sx, round_out, static_cast<unsigned short*>(out16),
                                                                               static_cast<unsigned short*>(out16_bf16));
  PMFB_LAUNCH_CHECK("pixel_shuffle_kernel");
  return PMFB_OK;
}

extern "C" int pmfb_pixel_shuffle_bwd(const pmfb_view* dy, int32_t n, int32_t h, int32_t w, int32_t c,
                                      const float* chan_scale, float* dx, int64_t d_sn, int64_t d_sy, int64_t d_sx,
                                      int32_t accumulate, int32_t round_out, void* stream) {
  REQ(dy && dy->ptr && view_ok(dy) && out_ok(dx, d_sn, d_sy, d_sx) && c % 4 == 0, "pixel_shuffle_bwd: bad arguments");
  const long long total = (long long)n * h * w * (c / 4);
  if (total == 0) return PMFB_OK;
  pixel_shuffle_bwd_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(ev(dy), n, h, w, c / 4, chan_scale, dx, d_sn,
                                                                                   d_sy, d_sx, accumulate, round_out);
  PMFB_LAUNCH_CHECK("pixel_shuffle_bwd_kernel");
  return PMFB_OK;
}

extern "C" int pmfb_upsample2x(const pmfb_view* x, int32_t n, int32_t h, int32_t w, int32_t c, float* out, int64_t o_sn,
                               int64_t o_sy, int64_t o_sx, int32_t round_out, void* out16, void* out16_bf16, void* stream) {
  REQ(x && x->ptr && view_ok(x) && out_ok(out, o_sn, o_sy, o_sx) && c % 4 == 0, "upsample2x: bad arguments");
  REQ(((reinterpret_cast<uintptr_t>(out16) | reinterpret_cast<uintptr_t>(out16_bf16)) & 7) == 0, "upsample2x: 16-bit outputs must be 8-byte aligned");
  const long long total = (long long)n * (2 * h) * (2 * w) * (c / 4);
  if (total == 0) return PMFB_OK;
  upsample2x_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(ev(x), n, h, w, c / 4, static_cast<unsigned short*>(out16),
                                                                            static_cast<unsigned short*>(out16_bf16), out, o_sn, o_sy, o_sx,
                                                                            round_out);
  PMFB_LAUNCH_CHECK("upsample2x_kernel");
  return PMFB_OK;
}

extern "C" int pmfb_upsample2x_bwd(const pmfb_view* dy, int32_t n, int32_t h, int32_t w, int32_t c, float* dx, int64_t d_sn,
                                   int64_t d_sy, int64_t d_sx, int32_t accumulate, void* stream) {
  REQ(dy && dy->ptr && view_ok(dy) && out_ok(dx, d_sn, d_sy, d_sx) && c % 4 == 0, "upsample2x_bwd: bad arguments");
  const long long total = (long long)n * h * w * (c / 4);
  if (total == 0) return PMFB_OK;
  upsample2x_bwd_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(ev(dy), n, h, w, c / 4, dx, d_sn, d_sy, d_sx,
                                                                                accumulate);
  PMFB_LAUNCH_CHECK("upsample2x_bwd_kernel");
  return PMFB_OK;
}

extern "C" int pmfb_softmax_nchw(const pmfb_view* logits, int32_t n, int32_t h, int32_t w, int32_t c, float* out,
                                 void* stream) {
  REQ(logits && logits->ptr && view_ok(logits) && out, "softmax_nchw: bad views");
  REQ(c > 0 && c <= kMaxSoftmaxC, "softmax_nchw: c=%d must be in [1,%d]", c, kMaxSoftmaxC);
  const long long total = (long long)n * h * w;
  if (total == 0) return PMFB_OK;
  softmax_nchw_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(ev(logits), n, h, w, c, out);
  PMFB_LAUNCH_CHECK("softmax_nchw_kernel");
  return PMFB_OK;
}

extern "C" int pmfb_softmax_nchw_bwd(const float* p, const float* dp, int32_t n, int32_t h, int32_t w, int32_t c, float* dz,
                                     int64_t d_sn, int64_t d_sy, int64_t d_sx, int32_t round_out, void* stream) {
  REQ(p && dp && out_ok(dz, d_sn, d_sy, d_sx), "softmax_nchw_bwd: bad views");
  REQ(c > 0 && c <= kMaxSoftmaxC, "softmax_nchw_bwd: c=%d must be in [1,%d]", c, kMaxSoftmaxC);
  const long long total = (long long)n * h * w;
  if (total == 0) return PMFB_OK;
  // dense destination: the row of pixel i starts at i * d_sx for every image/row, and d_sx is the padded channel count
  const int cpad = (int)d_sx;
  const int dense = (d_sy == (long long)w * d_sx && d_sn == (long long)h * d_sy && d_sx >= c && d_sx <= kMaxSoftmaxC) ? 1 : 0;
  const size_t smem = dense ? 256 * (size_t)(cpad | 1) * sizeof(float) : 0;
  const int grid = grid_for((total + 255) / 256 * 256, 256);
  if ((dense ? cpad : c) <= 24)  // the 20-class (KITTI) / 17-class (nuScenes) heads: a third of the registers
    softmax_nchw_bwd_kernel<24><<<grid, 256, smem, (cudaStream_t)stream>>>(p, dp, n, h, w, c, dense ? cpad : c, dz, d_sn, d_sy, d_sx,
                                                                           round_out, dense);
  else
    softmax_nchw_bwd_kernel<kMaxSoftmaxC><<<grid, 256, smem, (cudaStream_t)stream>>>(p, dp, n, h, w, c, dense ? cpad : c, dz, d_sn,
                                                                                     d_sy, d_sx, round_out, dense);
  PMFB_LAUNCH_CHECK("softmax_nchw_bwd_kernel");
  return PMFB_OK;
}
