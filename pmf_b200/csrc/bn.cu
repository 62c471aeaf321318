// BatchNorm2d statistics / backward and the other per-channel reductions of libpmf_b200.so.
//
// Reference semantics: nn.BatchNorm2d (eps 1e-5, momentum 0.1) as used at salsanext.py:18-21,48-61,120-132,
// pmf_net.py:17,23,28,190-209 and in torchvision's BasicBlock/Bottleneck; train mode normalises with the biased
// batch variance and updates running_var with the unbiased one.
//
// All reductions share one skeleton: a CTA owns G = min(C/4, 256) float4 channel groups and 256/G pixel lanes,
// strides over the pixels, keeps fp32 partials that are flushed into fp64 every 64 pixels, reduces the pixel
// lanes through shared memory and issues one fp64 atomicAdd per channel per CTA.
#include <stdlib.h>

#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.h"
#include "epilogue.cuh"

namespace pmfb {

constexpr int kRedThreads = 256;

// Pixel-indexed view.  Every activation the executor creates is "pixel-linear" (offset = pixel * sx: dense NHWC or a
// channel slice of it), which keeps 64-bit divisions out of the inner loops; broadcast (Dropout2d scale) and parity
// sub-grid views take the general decode.
struct PV {
  const float* p;
  long long sn, sy, sx;
  int linear;
};

__device__ __forceinline__ const float* pv_at(const PV& v, unsigned pix, unsigned hw, unsigned w, int c) {
  if (v.linear) return v.p + (long long)pix * v.sx + c;
  const unsigned n = pix / hw, q = pix - n * hw, y = q / w, x = q - y * w;
  return v.p + (long long)n * v.sn + (long long)y * v.sy + (long long)x * v.sx + c;
}

// LIN: the view is known (checked on the host) to be pixel-linear: no runtime branch, no divisions in the unrolled loops
template <bool LIN>
__device__ __forceinline__ const float* pv_at_t(const PV& v, unsigned pix, unsigned hw, unsigned w, int c) {
  if constexpr (LIN) return v.p + (long long)pix * v.sx + c;
  else return pv_at(v, pix, hw, w, c);
}

static inline PV pv(const pmfb_view* v, int h, int w) {
  PV o;
  o.p = v ? v->ptr : nullptr;
  o.sn = v ? v->sn : 0;
  o.sy = v ? v->sy : 0;
  o.sx = v ? v->sx : 0;
  o.linear = (o.p && o.sy == (long long)w * o.sx && o.sn == (long long)h * o.sy) ? 1 : 0;
  return o;
}
static inline PV pv_out(float* p, long long sn, long long sy, long long sx, int h, int w) {
  PV o;
  o.p = p;
  o.sn = sn;
  o.sy = sy;
  o.sx = sx;
  o.linear = (p && sy == (long long)w * sx && sn == (long long)h * sy) ? 1 : 0;
  return o;
}

// Per-thread partial sums stay in fp32: the grid is sized so that a thread adds at most a few hundred values
// (>= 16 pixels per thread, <= npix / (grid.x * lanes)); the cross-thread and cross-CTA reduction is fp64.
template <int NRED>
struct RedAcc {
  float4 f[NRED];
  __device__ __forceinline__ void init() {
#pragma unroll
    for (int r = 0; r < NRED; ++r) f[r] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  __device__ __forceinline__ void add(int r, float4 v) {
    f[r].x += v.x; f[r].y += v.y; f[r].z += v.z; f[r].w += v.w;
  }
  __device__ __forceinline__ double get(int r, int k) const {
    return k == 0 ? (double)f[r].x : k == 1 ? (double)f[r].y : k == 2 ? (double)f[r].z : (double)f[r].w;
  }
};

// F::load(pix, c, L&) issues the global loads of one (pixel, 4-channel group); F::consume(pix, c, L&, acc) uses them.
// Four pixels are loaded before the first is consumed, so every thread keeps >= 4 independent 16-byte loads per
// operand in flight (these kernels are pure HBM streams).
template <int NRED, class F>
__global__ void __launch_bounds__(kRedThreads, 3)
chan_reduce_kernel(F f, unsigned npix, unsigned hw, unsigned w, int c4, int G, int per_image, double* __restrict__ out) {
  __shared__ float sm[NRED * 4 * kRedThreads];
  const int L = kRedThreads / G;
  const int gl = threadIdx.x % G, pl = threadIdx.x / G;
  const int cg = blockIdx.y * G + gl;
  const bool active = (pl < L) && (cg < c4);
  const unsigned img0 = per_image ? blockIdx.z : 0;
  const unsigned base = img0 * hw;  // per_image: npix == hw pixels of image blockIdx.z
  RedAcc<NRED> acc;
  acc.init();
  if (active) {
    const unsigned stride = gridDim.x * L;
    const int c = cg * 4;
    unsigned p = blockIdx.x * L + pl;
    typename F::Consts k;
    f.prep(c, k, blockIdx.x == 0 && pl == 0 && blockIdx.z == 0);  // per-channel constants of this thread's four channels, loaded once
    constexpr int U = F::kUnroll;
    for (; p + (U - 1) * stride < npix; p += U * stride) {
      typename F::Loaded l[U];
#pragma unroll
      for (int u = 0; u < U; ++u) f.load(base + p + u * stride, hw, w, c, l[u]);
#pragma unroll
      for (int u = 0; u < U; ++u) f.consume(base + p + u * stride, hw, w, c, k, l[u], acc);
    }
    for (; p < npix; p += stride) {
      typename F::Loaded l0;
      f.load(base + p, hw, w, c, l0);
      f.consume(base + p, hw, w, c, k, l0, acc);
    }
  }
  // Block reduction: every thread parks its NRED*4 fp32 partials in shared memory; G*NRED*4 threads then each sum one
  // (channel, quantity) over the L pixel lanes in fp64 and issue ONE atomic.  A block owns at most 32 channels (G <= 8),
  // so a launch issues (pixel blocks) x C x NRED atomics in total instead of (all blocks) x C x NRED: ncu showed the
  // atomic drain, not the streaming loop, dominating the small-tensor launches.
#pragma unroll
  for (int r = 0; r < NRED; ++r) {
    sm[(r * 4 + 0) * kRedThreads + threadIdx.x] = active ? acc.f[r].x : 0.f;
    sm[(r * 4 + 1) * kRedThreads + threadIdx.x] = active ? acc.f[r].y : 0.f;
    sm[(r * 4 + 2) * kRedThreads + threadIdx.x] = active ? acc.f[r].z : 0.f;
    sm[(r * 4 + 3) * kRedThreads + threadIdx.x] = active ? acc.f[r].w : 0.f;
  }
  __syncthreads();
  if (out && (int)threadIdx.x < G * NRED * 4) {
    const int g2 = threadIdx.x % G, q = threadIdx.x / G;  // q = r*4 + k
    const int cg2 = blockIdx.y * G + g2;
    if (cg2 < c4) {
      double s = 0.0;
      for (int l = 0; l < L; ++l) s += (double)sm[q * kRedThreads + l * G + g2];
      const int C = c4 * 4;
      atomicAdd(out + ((long long)img0 * NRED + (q >> 2)) * C + cg2 * 4 + (q & 3), s);
    }
  }
}

struct StatsF {
  static constexpr int kUnroll = 8;
  PV x;
  struct Loaded { float4 v; };
  struct Consts {};
  __device__ __forceinline__ void prep(int, Consts&, bool) const {}
  __device__ __forceinline__ void load(unsigned pix, unsigned hw, unsigned w, int c, Loaded& l) const {
    l.v = ld4(pv_at(x, pix, hw, w, c));
  }
  __device__ __forceinline__ void consume(unsigned, unsigned, unsigned, int, const Consts&, const Loaded& l, RedAcc<2>& a) const {
    a.add(0, l.v);
    a.add(1, make_float4(l.v.x * l.v.x, l.v.y * l.v.y, l.v.z * l.v.z, l.v.w * l.v.w));
  }
};

struct ColsumF {
  static constexpr int kUnroll = 8;
  PV x;
  struct Loaded { float4 v; };
  struct Consts {};
  __device__ __forceinline__ void prep(int, Consts&, bool) const {}
  __device__ __forceinline__ void load(unsigned pix, unsigned hw, unsigned w, int c, Loaded& l) const {
    l.v = ld4(pv_at(x, pix, hw, w, c));
  }
  __device__ __forceinline__ void consume(unsigned, unsigned, unsigned, int, const Consts&, const Loaded& l, RedAcc<1>& a) const {
    a.add(0, l.v);
  }
};

__device__ __forceinline__ float act_grad(int act, float z) {
  switch (act) {
    case PMFB_ACT_RELU: return z > 0.f ? 1.f : 0.f;
    case PMFB_ACT_LEAKY: return z > 0.f ? 1.f : 0.01f;
    case PMFB_ACT_SIGMOID: return z * (1.f - z);
    default: return 1.f;
  }
}

// g = dy * mul * act'(z); z is either stored (zv) or recomputed as act(alpha*x + beta) (sigmoid gate).
// MUL / ZST (stored z) are compile-time: a launch without them (the plain conv -> LeakyReLU -> BN backward, most of the
// bytes) carries two operand streams, keeps four pixels of loads in flight per thread and has no dead registers.
// four packed fp16 values (8 bytes) -> float4
__device__ __forceinline__ float4 unpack_h4(uint2 r) {
  const __half2 a = *reinterpret_cast<const __half2*>(&r.x), b = *reinterpret_cast<const __half2*>(&r.y);
  const float2 fa = __half22float2(a), fb = __half22float2(b);
  return make_float4(fa.x, fa.y, fb.x, fb.y);
}

// XH: x (the stored pre-BatchNorm activation) is an fp16 buffer with the same ELEMENT strides; its four values travel
// packed (8 bytes) until they are used, and the kernels keep twice as many pixels in flight (same bytes in flight per thread).
template <bool MUL, bool ZST, bool XH>
struct GradIn {
  PV dy, mul, z, x;
  const float* mean;
  const float* invstd;
  const float* alpha;
  const float* beta;
  int act_z;
  struct Loaded { float4 g, xv[XH ? 0 : 1], m[MUL ? 1 : 0], zz[ZST ? 1 : 0]; uint2 xh[XH ? 1 : 0]; };
  // XH launches always carry BatchNorm statistics (host-checked) and recompute z = act(alpha*x + beta) only in the gated
  // (MUL) form: the plain conv -> LeakyReLU -> BN backward then keeps 8 fewer constant registers per thread.
  static constexpr bool kRecomp = !ZST && (MUL || !XH);
  struct Consts { float4 mu, is, a[kRecomp ? 1 : 0], b[kRecomp ? 1 : 0]; };
  __device__ __forceinline__ bool has_bn() const { return XH || mean != nullptr; }
  __device__ __forceinline__ void prep(int c, Consts& k, bool = false) const {
    k.mu = k.is = make_float4(0.f, 0.f, 0.f, 0.f);
    if (has_bn()) { k.mu = ld4(mean + c); k.is = ld4(invstd + c); }
    if constexpr (kRecomp) {
      k.a[0] = k.b[0] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (act_z) { k.a[0] = ld4(alpha + c); k.b[0] = ld4(beta + c); }
    }
  }
  __device__ __forceinline__ void load(unsigned pix, unsigned hw, unsigned w, int c, Loaded& l) const {
    l.g = ld4(pv_at_t<XH>(dy, pix, hw, w, c));
    if constexpr (MUL) l.m[0] = ld4(pv_at(mul, pix, hw, w, c));
    if constexpr (XH) {  // x present and pixel-linear (host-checked)
      l.xh[0] = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const unsigned short*>(x.p) + ((long long)pix * x.sx + c)));
    } else {
      if (x.p) l.xv[0] = ld4(pv_at(x, pix, hw, w, c));
    }
    if constexpr (ZST) l.zz[0] = ld4(pv_at_t<XH>(z, pix, hw, w, c));
  }
  __device__ __forceinline__ void finish(const Consts& k, const Loaded& l, float4& g, float4& xhat, float4& xv) const {
    g = l.g;
    if constexpr (MUL) { g.x *= l.m[0].x; g.y *= l.m[0].y; g.z *= l.m[0].z; g.w *= l.m[0].w; }
    if constexpr (XH) xv = unpack_h4(l.xh[0]);
    else xv = x.p ? l.xv[0] : make_float4(0.f, 0.f, 0.f, 0.f);
    if constexpr (ZST || kRecomp) {
      if (act_z) {
        float4 zz;
        if constexpr (ZST) {
          zz = l.zz[0];
        } else {
          const float4 a = k.a[0], b = k.b[0];
          zz = make_float4(epi_act(act_z, a.x * xv.x + b.x), epi_act(act_z, a.y * xv.y + b.y),
                           epi_act(act_z, a.z * xv.z + b.z), epi_act(act_z, a.w * xv.w + b.w));
        }
        g.x *= act_grad(act_z, zz.x); g.y *= act_grad(act_z, zz.y);
        g.z *= act_grad(act_z, zz.z); g.w *= act_grad(act_z, zz.w);
      }
    }
    if (has_bn()) {
      const float4 mu = k.mu, is = k.is;
      xhat = make_float4((xv.x - mu.x) * is.x, (xv.y - mu.y) * is.y, (xv.z - mu.z) * is.z, (xv.w - mu.w) * is.w);
    } else {
      xhat = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
};

template <bool MUL, bool ZST, bool XH = false>
struct BnBwdReduceF {
  static constexpr int kUnroll = XH ? ((MUL || ZST) ? 3 : 6) : ((MUL || ZST) ? 2 : 4);
  GradIn<MUL, ZST, XH> in;
  typedef typename GradIn<MUL, ZST, XH>::Loaded Loaded;
  typedef typename GradIn<MUL, ZST, XH>::Consts Consts;
  __device__ __forceinline__ void prep(int c, Consts& k, bool) const { in.prep(c, k); }
  __device__ __forceinline__ void load(unsigned pix, unsigned hw, unsigned w, int c, Loaded& l) const { in.load(pix, hw, w, c, l); }
  __device__ __forceinline__ void consume(unsigned, unsigned, unsigned, int, const Consts& k, const Loaded& l, RedAcc<2>& a) const {
    float4 g, xh, xv;
    in.finish(k, l, g, xh, xv);
    a.add(0, g);
    a.add(1, make_float4(g.x * xh.x, g.y * xh.y, g.z * xh.z, g.w * xh.w));
  }
};

// With BN (in.mean != NULL): dx = gamma*invstd*(g - S1/M - xhat*S2/M) [* leaky'(x)].
// Without BN: dx = g [* leaky'(x)]   (plain activation backward, e.g. the conv->LeakyReLU shortcuts).
// GACC: g_out is accumulated into (its old value is one more operand stream).
template <bool MUL, bool ZST, bool GACC, bool XH = false>
struct BnBwdApplyF {
  static constexpr int kUnroll = XH ? ((MUL || ZST || GACC) ? 3 : 6) : ((MUL || ZST || GACC) ? 2 : 4);
  GradIn<MUL, ZST, XH> in;
  const float* gamma;
  const double* red;
  float* dgamma;  // BatchNorm parameter gradients, stored by the owner threads (16-byte aligned) or NULL
  float* dbeta;
  double inv_count;
  int C, leaky_x, round_out;
  PV dx;
  PV g_out;
  unsigned short* dx16;  // optional bf16 copy of dx (same element offsets): the operand of the kind::f16 dgrad / wgrad
  int write_dx32;        // 0: only the bf16 copy is stored (dx.p is then a placeholder base used for offsets only)
  struct Loaded { typename GradIn<MUL, ZST, XH>::Loaded i; float4 e[GACC ? 1 : 0]; };
  struct Consts { typename GradIn<MUL, ZST, XH>::Consts i; float4 gi, m1, m2; };
  // owner: exactly one thread per 4-channel group of the launch -- it also stores the BatchNorm parameter gradients
  // (dbeta = sum g, dgamma = sum g * xhat: the reduce pass's sums), which used to be a separate tiny launch
  __device__ __forceinline__ void prep(int c, Consts& k, bool owner) const {
    in.prep(c, k.i);
    k.gi = k.m1 = k.m2 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (in.has_bn()) {
      if (owner) {
        if (dbeta) *reinterpret_cast<float4*>(dbeta + c) = make_float4((float)red[c], (float)red[c + 1], (float)red[c + 2], (float)red[c + 3]);
        if (dgamma)
          *reinterpret_cast<float4*>(dgamma + c) = make_float4((float)red[C + c], (float)red[C + c + 1], (float)red[C + c + 2], (float)red[C + c + 3]);
      }
      const float4 gm = ld4(gamma + c);
      k.gi = make_float4(gm.x * k.i.is.x, gm.y * k.i.is.y, gm.z * k.i.is.z, gm.w * k.i.is.w);
      k.m1 = make_float4((float)(red[c] * inv_count), (float)(red[c + 1] * inv_count), (float)(red[c + 2] * inv_count),
                         (float)(red[c + 3] * inv_count));
      k.m2 = make_float4((float)(red[C + c] * inv_count), (float)(red[C + c + 1] * inv_count),
                         (float)(red[C + c + 2] * inv_count), (float)(red[C + c + 3] * inv_count));
    }
  }
  __device__ __forceinline__ void load(unsigned pix, unsigned hw, unsigned w, int c, Loaded& l) const {
    in.load(pix, hw, w, c, l.i);
    if constexpr (GACC) l.e[0] = ld4(pv_at_t<XH>(g_out, pix, hw, w, c));
  }
  __device__ __forceinline__ void consume(unsigned pix, unsigned hw, unsigned w, int c, const Consts& k, const Loaded& l,
                                          RedAcc<1>& a) const {
    float4 g, xh, xv;
    in.finish(k.i, l.i, g, xh, xv);
    if (g_out.p) {
      float4 o = g;
      if constexpr (GACC) { o.x += l.e[0].x; o.y += l.e[0].y; o.z += l.e[0].z; o.w += l.e[0].w; }
      *reinterpret_cast<float4*>(const_cast<float*>(pv_at_t<XH>(g_out, pix, hw, w, c))) = o;
    }
    float4 d = g;
    if (in.has_bn()) {
      d.x = k.gi.x * (g.x - k.m1.x - xh.x * k.m2.x);
      d.y = k.gi.y * (g.y - k.m1.y - xh.y * k.m2.y);
      d.z = k.gi.z * (g.z - k.m1.z - xh.z * k.m2.z);
      d.w = k.gi.w * (g.w - k.m1.w - xh.w * k.m2.w);
    }
    if (leaky_x) {
      d.x *= xv.x > 0.f ? 1.f : 0.01f; d.y *= xv.y > 0.f ? 1.f : 0.01f;
      d.z *= xv.z > 0.f ? 1.f : 0.01f; d.w *= xv.w > 0.f ? 1.f : 0.01f;
    }
    a.add(0, d);  // bias gradient of the preceding conv: summed before the tf32 rounding of the stored value
    if (round_out) {
      d.x = round_tf32(d.x); d.y = round_tf32(d.y); d.z = round_tf32(d.z); d.w = round_tf32(d.w);
    }
    if (dx.p) {
      float* dp = const_cast<float*>(pv_at_t<XH>(dx, pix, hw, w, c));
      if (write_dx32) *reinterpret_cast<float4*>(dp) = d;
      if (dx16) {
        const __nv_bfloat162 a = __floats2bfloat162_rn(d.x, d.y), b = __floats2bfloat162_rn(d.z, d.w);
        uint2 r;
        r.x = *reinterpret_cast<const unsigned int*>(&a);
        r.y = *reinterpret_cast<const unsigned int*>(&b);
        *reinterpret_cast<uint2*>(dx16 + (dp - dx.p)) = r;
      }
    }
  }
};

__global__ void bn_finalize_kernel(const double* __restrict__ sums, double count, int c, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float* running_mean, float* running_var, float momentum,
                                   float eps, float* alpha, float* beta_out, float* mean_out, float* invstd_out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= c) return;
  float mean, var;
  if (sums) {
    const double m = sums[i] / count;
    double v = sums[c + i] / count - m * m;
    if (v < 0.0) v = 0.0;
    mean = (float)m;
    var = (float)v;
    if (running_mean) running_mean[i] = (1.f - momentum) * running_mean[i] + momentum * mean;
    if (running_var) {
      const float unbiased = count > 1.0 ? (float)(v * count / (count - 1.0)) : var;
      running_var[i] = (1.f - momentum) * running_var[i] + momentum * unbiased;
    }
  } else {
    mean = running_mean[i];
    var = running_var[i];
  }
  const float invstd = 1.f / sqrtf(var + eps);
  const float g = gamma ? gamma[i] : 1.f;
  const float b = beta ? beta[i] : 0.f;
  const float a = g * invstd;
  if (alpha) alpha[i] = a;
  if (beta_out) beta_out[i] = b - mean * a;
  if (mean_out) mean_out[i] = mean;
  if (invstd_out) invstd_out[i] = invstd;
}

__global__ void red_to_params_kernel(const double* __restrict__ red, int c, float* dgamma, float* dbeta) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= c) return;
  if (dbeta) dbeta[i] = (float)red[i];
  if (dgamma) dgamma[i] = (float)red[c + i];
}

static inline bool vok(const pmfb_view* v) {
  return !v || !v->ptr || ((((v->sn | v->sy | v->sx) % 4) == 0) && ((reinterpret_cast<uintptr_t>(v->ptr) & 15) == 0));
}

struct RedGrid {
  dim3 grid;
  int G;
};
// One resident wave: grid.x * grid.y * grid.z == SMs * (blocks of this kernel an SM can hold), so no tail wave.
template <class K>
static inline RedGrid red_grid(K kernel, long long npix, int c4, int images) {
  static int per_sm = 0, sms = 0;  // one static pair per kernel instantiation
  if (per_sm == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kRedThreads, 0) != cudaSuccess || per_sm < 1) per_sm = 2;
  }
  RedGrid r;
  static int g_cap = 0;
  if (g_cap == 0) {
    const char* e = getenv("PMFB_RED_G");
    g_cap = e ? atoi(e) : 8;
    if (g_cap < 1 || g_cap > 64 || (g_cap & (g_cap - 1))) g_cap = 8;
  }
  r.G = c4 < g_cap ? c4 : g_cap;  // <= 32 channels per block by default: see the block-reduction note in chan_reduce_kernel
  const int L = kRedThreads / r.G;
  const int gy = (c4 + r.G - 1) / r.G;
  long long gx = (npix + (long long)L * 16 - 1) / ((long long)L * 16);  // >= 16 pixels per thread
  long long cap = ((long long)sms * per_sm) / ((long long)gy * images);
  if (cap < 1) cap = 1;
  if (gx > cap) gx = cap;
  if (gx < 1) gx = 1;
  r.grid = dim3((unsigned)gx, (unsigned)gy, (unsigned)images);
  return r;
}

}  // namespace pmfb

using namespace pmfb;

#define REQ(cond, ...) \
  do {                 \
    if (!(cond)) return fail(PMFB_ERR_INVALID, __VA_ARGS__); \
  } while (0)

extern "C" int pmfb_bn_stats(const pmfb_view* x, int32_t n, int32_t h, int32_t w, int32_t c, double* sums, void* stream) {
  REQ(x && x->ptr && vok(x) && sums && c > 0 && c % 4 == 0, "bn_stats: bad arguments (c=%d)", c);
  const long long npix = (long long)n * h * w;
  if (npix == 0) return PMFB_OK;
  RedGrid g = red_grid(chan_reduce_kernel<2, StatsF>, npix, c / 4, 1);
  StatsF f{pv(x, h, w)};
  chan_reduce_kernel<2, StatsF><<<g.grid, kRedThreads, 0, (cudaStream_t)stream>>>(f, (unsigned)npix, (unsigned)(h * w), (unsigned)w,
                                                                                  c / 4, g.G, 0, sums);
  PMFB_LAUNCH_CHECK("bn_stats");
  return PMFB_OK;
}

extern "C" int pmfb_colsum(const pmfb_view* x, int32_t n, int32_t h, int32_t w, int32_t c, int32_t per_image, double* out,
                           void* stream) {
  REQ(x && x->ptr && vok(x) && out && c > 0 && c % 4 == 0, "colsum: bad arguments (c=%d)", c);
  const long long npix = per_image ? (long long)h * w : (long long)n * h * w;
  if (npix == 0 || n == 0) return PMFB_OK;
  RedGrid g = red_grid(chan_reduce_kernel<1, ColsumF>, npix, c / 4, per_image ? n : 1);
  ColsumF f{pv(x, h, w)};
  chan_reduce_kernel<1, ColsumF><<<g.grid, kRedThreads, 0, (cudaStream_t)stream>>>(f, (unsigned)npix, (unsigned)(h * w), (unsigned)w,
                                                                                   c / 4, g.G, per_image, out);
  PMFB_LAUNCH_CHECK("colsum");
  return PMFB_OK;
}

extern "C" int pmfb_bn_finalize(const double* sums, int64_t count, int32_t c, const float* gamma, const float* beta,
                                float* running_mean, float* running_var, float momentum, float eps, float* alpha,
                                float* beta_out, float* mean_out, float* invstd_out, void* stream) {
  REQ(c > 0 && (sums || (running_mean && running_var)), "bn_finalize: need sums (train) or running stats (eval)");
  REQ(!sums || count > 0, "bn_finalize: count must be positive");
  bn_finalize_kernel<<<(c + 127) / 128, 128, 0, (cudaStream_t)stream>>>(sums, (double)count, c, gamma, beta, running_mean,
                                                                        running_var, momentum, eps, alpha, beta_out,
                                                                        mean_out, invstd_out);
  PMFB_LAUNCH_CHECK("bn_finalize");
  return PMFB_OK;
}

template <class GI>
static int fill_grad_in(GI* gi, const pmfb_view* dy, const pmfb_view* mul, const pmfb_view* z, int act_z,
                        const pmfb_view* x, const float* mean, const float* invstd, const float* alpha, const float* beta,
                        int h, int w) {
  if (!dy || !dy->ptr || !vok(dy) || !vok(mul) || !vok(z) || !vok(x)) return fail(PMFB_ERR_INVALID, "bn_bwd: bad views");
  if (act_z < 0 || act_z > 3) return fail(PMFB_ERR_INVALID, "bn_bwd: act_z=%d", act_z);
  if (act_z && !(z && z->ptr) && !(alpha && beta && x && x->ptr))
    return fail(PMFB_ERR_INVALID, "bn_bwd: act_z without z needs x, alpha and beta to recompute it");
  if ((mean != nullptr) != (invstd != nullptr)) return fail(PMFB_ERR_INVALID, "bn_bwd: mean and invstd go together");
  if (mean && !(x && x->ptr)) return fail(PMFB_ERR_INVALID, "bn_bwd: BN backward needs x");
  gi->dy = pv(dy, h, w);
  gi->mul = pv(mul, h, w);
  gi->z = pv(z, h, w);
  gi->x = pv(x, h, w);
  gi->mean = mean;
  gi->invstd = invstd;
  gi->alpha = alpha;
  gi->beta = beta;
  gi->act_z = act_z;
  return PMFB_OK;
}

template <bool MUL, bool ZST, bool XH>
static int bn_bwd_reduce_t(const pmfb_view* dy, const pmfb_view* mul, const pmfb_view* z, int32_t act_z, const pmfb_view* x,
                           const float* mean, const float* invstd, const float* alpha, const float* beta, int32_t n, int32_t h,
                           int32_t w, int32_t c, double* red, void* stream) {
  typedef BnBwdReduceF<MUL, ZST, XH> F;
  F f;
  int rc = fill_grad_in(&f.in, dy, mul, z, act_z, x, mean, invstd, alpha, beta, h, w);
  if (rc) return rc;
  if (XH && !(f.in.x.p && f.in.x.linear && f.in.dy.linear && (!ZST || f.in.z.linear)))
    return fail(PMFB_ERR_INVALID, "bn_bwd_reduce16: x_half needs dense (pixel-linear) dy / x / z views");
  if (XH && !MUL && !ZST && act_z) return fail(PMFB_ERR_INVALID, "bn_bwd_reduce16: x_half recomputes z only in the gated (mul) form");
  const long long npix = (long long)n * h * w;
  if (npix == 0) return PMFB_OK;
  RedGrid g = red_grid(chan_reduce_kernel<2, F>, npix, c / 4, 1);
  chan_reduce_kernel<2, F><<<g.grid, kRedThreads, 0, (cudaStream_t)stream>>>(f, (unsigned)npix, (unsigned)(h * w), (unsigned)w, c / 4,
                                                                             g.G, 0, red);
  PMFB_LAUNCH_CHECK("bn_bwd_reduce");
  return PMFB_OK;
}

struct ApplyArgs {
  const pmfb_view *dy, *mul, *z, *x;
  int act_z;
  const float *mean, *invstd, *alpha, *beta, *gamma;
  const double* red;
  int leaky_x, n, h, w, c;
  float* dx;
  long long d_sn, d_sy, d_sx;
  int round_out;
  float *dgamma, *dbeta;
  double* colsum;
  float* g_out;
  long long g_sn, g_sy, g_sx;
  void* stream;
  void* dx16;
};

template <bool MUL, bool ZST, bool GACC, bool XH>
static int bn_bwd_apply_t(const ApplyArgs& A) {
  typedef BnBwdApplyF<MUL, ZST, GACC, XH> F;
  F f;
  int rc = fill_grad_in(&f.in, A.dy, A.mul, A.z, A.act_z, A.x, A.mean, A.invstd, A.alpha, A.beta, A.h, A.w);
  if (rc) return rc;
  const long long npix = (long long)A.n * A.h * A.w;
  if (npix == 0) return PMFB_OK;
  f.gamma = A.gamma;
  f.red = A.red;
  const bool vec_ok = ((reinterpret_cast<uintptr_t>(A.dgamma) | reinterpret_cast<uintptr_t>(A.dbeta)) & 15) == 0;
  f.dgamma = (A.mean && vec_ok) ? A.dgamma : nullptr;
  f.dbeta = (A.mean && vec_ok) ? A.dbeta : nullptr;
  f.inv_count = 1.0 / (double)npix;
  f.C = A.c;
  f.leaky_x = A.leaky_x;
  f.round_out = A.round_out;
  // dx == NULL with dx16 set: only the bf16 copy is wanted; a placeholder base keeps the offset arithmetic
  f.dx = pv_out(A.dx ? A.dx : (A.dx16 ? reinterpret_cast<float*>(uintptr_t(1024)) : nullptr), A.d_sn, A.d_sy, A.d_sx, A.h, A.w);
  f.dx16 = static_cast<unsigned short*>(A.dx16);
  f.write_dx32 = A.dx ? 1 : 0;
  f.g_out = pv_out(A.g_out, A.g_sn, A.g_sy, A.g_sx, A.h, A.w);
  if (XH && !(f.in.x.p && f.in.x.linear && f.in.dy.linear && (!ZST || f.in.z.linear) && (!f.dx.p || f.dx.linear) &&
              (!f.g_out.p || f.g_out.linear)))
    return fail(PMFB_ERR_INVALID, "bn_bwd_apply16: x_half needs dense (pixel-linear) dy / x / z / dx / g_out views");
  if (XH && (!A.mean || (!MUL && !ZST && A.act_z)))
    return fail(PMFB_ERR_INVALID, "bn_bwd_apply16: x_half needs BatchNorm statistics and recomputes z only in the gated (mul) form");
  RedGrid g = red_grid(chan_reduce_kernel<1, F>, npix, A.c / 4, 1);
  chan_reduce_kernel<1, F><<<g.grid, kRedThreads, 0, (cudaStream_t)A.stream>>>(f, (unsigned)npix, (unsigned)(A.h * A.w), (unsigned)A.w,
                                                                               A.c / 4, g.G, 0, A.colsum);
  PMFB_LAUNCH_CHECK("bn_bwd_apply");
  if (A.mean && (A.dgamma || A.dbeta) && !vec_ok) {  // unaligned parameter-gradient storage: the separate copy
    red_to_params_kernel<<<(A.c + 127) / 128, 128, 0, (cudaStream_t)A.stream>>>(A.red, A.c, A.dgamma, A.dbeta);
    PMFB_LAUNCH_CHECK("red_to_params");
  }
  return PMFB_OK;
}

extern "C" int pmfb_bn_bwd_reduce16(const pmfb_view* dy, const pmfb_view* mul, const pmfb_view* z, int32_t act_z,
                                    const pmfb_view* x, const float* mean, const float* invstd, const float* alpha,
                                    const float* beta, int32_t n, int32_t h, int32_t w, int32_t c, double* red, int32_t x_half,
                                    void* stream) {
  REQ(red && mean && c > 0 && c % 4 == 0, "bn_bwd_reduce: bad arguments");
  const bool has_mul = mul && mul->ptr, has_z = act_z && z && z->ptr;
  const int key = (has_mul ? 4 : 0) | (has_z ? 2 : 0) | (x_half ? 1 : 0);
#define PMFB_RED(m, zz, xh) return bn_bwd_reduce_t<m, zz, xh>(dy, mul, z, act_z, x, mean, invstd, alpha, beta, n, h, w, c, red, stream)
  switch (key) {
    case 0: PMFB_RED(false, false, false);
    case 1: PMFB_RED(false, false, true);
    case 2: PMFB_RED(false, true, false);
    case 3: PMFB_RED(false, true, true);
    case 4: PMFB_RED(true, false, false);
    case 5: PMFB_RED(true, false, true);
    case 6: PMFB_RED(true, true, false);
    default: PMFB_RED(true, true, true);
  }
#undef PMFB_RED
}

extern "C" int pmfb_bn_bwd_reduce(const pmfb_view* dy, const pmfb_view* mul, const pmfb_view* z, int32_t act_z,
                                  const pmfb_view* x, const float* mean, const float* invstd, const float* alpha,
                                  const float* beta, int32_t n, int32_t h, int32_t w, int32_t c, double* red,
                                  void* stream) {
  return pmfb_bn_bwd_reduce16(dy, mul, z, act_z, x, mean, invstd, alpha, beta, n, h, w, c, red, 0, stream);
}

extern "C" int pmfb_bn_bwd_apply(const pmfb_view* dy, const pmfb_view* mul, const pmfb_view* z, int32_t act_z,
                                 const pmfb_view* x, const float* mean, const float* invstd, const float* alpha,
                                 const float* beta, const float* gamma, const double* red, int32_t leaky_x, int32_t n,
                                 int32_t h, int32_t w, int32_t c, float* dx, int64_t d_sn, int64_t d_sy, int64_t d_sx,
                                 int32_t round_out, float* dgamma, float* dbeta, double* colsum, float* g_out, int64_t g_sn,
                                 int64_t g_sy, int64_t g_sx, int32_t g_accumulate, void* stream) {
  return pmfb_bn_bwd_apply16(dy, mul, z, act_z, x, mean, invstd, alpha, beta, gamma, red, leaky_x, n, h, w, c, dx, d_sn, d_sy, d_sx,
                             round_out, dgamma, dbeta, colsum, g_out, g_sn, g_sy, g_sx, g_accumulate, nullptr, 0, stream);
}

extern "C" int pmfb_bn_bwd_apply16(const pmfb_view* dy, const pmfb_view* mul, const pmfb_view* z, int32_t act_z,
                                   const pmfb_view* x, const float* mean, const float* invstd, const float* alpha,
                                   const float* beta, const float* gamma, const double* red, int32_t leaky_x, int32_t n,
                                   int32_t h, int32_t w, int32_t c, float* dx, int64_t d_sn, int64_t d_sy, int64_t d_sx,
                                   int32_t round_out, float* dgamma, float* dbeta, double* colsum, float* g_out, int64_t g_sn,
                                   int64_t g_sy, int64_t g_sx, int32_t g_accumulate, void* dx16, int32_t x_half, void* stream) {
  REQ(!dx16 || (reinterpret_cast<uintptr_t>(dx16) & 7) == 0, "bn_bwd_apply16: dx16 must be 8-byte aligned");
  REQ(dx || !dx16 || (((d_sn | d_sy | d_sx) % 4) == 0), "bn_bwd_apply16: bad dx16 strides");
  REQ(c > 0 && c % 4 == 0, "bn_bwd_apply: c=%d", c);
  REQ(!mean || (gamma && red), "bn_bwd_apply: BN backward needs gamma and red");
  REQ(!leaky_x || (x && x->ptr), "bn_bwd_apply: leaky_x needs x");
  REQ(!dx || ((((d_sn | d_sy | d_sx) % 4) == 0) && ((reinterpret_cast<uintptr_t>(dx) & 15) == 0)), "bn_bwd_apply: bad dx view");
  REQ(!g_out || ((((g_sn | g_sy | g_sx) % 4) == 0) && ((reinterpret_cast<uintptr_t>(g_out) & 15) == 0)),
      "bn_bwd_apply: bad g_out view");
  ApplyArgs A{dy, mul, z, x, act_z, mean, invstd, alpha, beta, gamma, red, leaky_x, n, h, w, c, dx, d_sn, d_sy, d_sx, round_out,
              dgamma, dbeta, colsum, g_out, g_sn, g_sy, g_sx, stream, dx16};
  const int key = ((mul && mul->ptr) ? 4 : 0) | ((act_z && z && z->ptr) ? 2 : 0) | ((g_out && g_accumulate) ? 1 : 0);
  if (x_half) {
    REQ(x && x->ptr, "bn_bwd_apply16: x_half needs x");
    switch (key) {
      case 0: return bn_bwd_apply_t<false, false, false, true>(A);
      case 1: return bn_bwd_apply_t<false, false, true, true>(A);
      case 2: return bn_bwd_apply_t<false, true, false, true>(A);
      case 3: return bn_bwd_apply_t<false, true, true, true>(A);
      case 4: return bn_bwd_apply_t<true, false, false, true>(A);
      case 5: return bn_bwd_apply_t<true, false, true, true>(A);
      case 6: return bn_bwd_apply_t<true, true, false, true>(A);
      default: return bn_bwd_apply_t<true, true, true, true>(A);
    }
  }
  switch (key) {
    case 0: return bn_bwd_apply_t<false, false, false, false>(A);
    case 1: return bn_bwd_apply_t<false, false, true, false>(A);
    case 2: return bn_bwd_apply_t<false, true, false, false>(A);
    case 3: return bn_bwd_apply_t<false, true, true, false>(A);
    case 4: return bn_bwd_apply_t<true, false, false, false>(A);
    case 5: return bn_bwd_apply_t<true, false, true, false>(A);
    case 6: return bn_bwd_apply_t<true, true, false, false>(A);
    default: return bn_bwd_apply_t<true, true, true, false>(A);
  }
}
