// BatchNorm2d statistics / backward and the other per-channel reductions of libpmf_b200.so.
//
// Reference semantics: nn.BatchNorm2d (eps 1e-5, momentum 0.1) as used at salsanext.py:18-21,48-61,120-132,
// pmf_net.py:17,23,28,190-209 and in torchvision's BasicBlock/Bottleneck; train mode normalises with the biased
// batch variance and updates running_var with the unbiased one.
//
// All reductions share one skeleton: a CTA owns G = min(C/4, 256) float4 channel groups and 256/G pixel lanes,
// strides over the pixels, keeps fp32 partials that are flushed into fp64 every 64 pixels, reduces the pixel
// lanes through shared memory and issues one fp64 atomicAdd per channel per CTA.
#include "common.h"
#include "epilogue.cuh"

namespace pmfb {

constexpr int kRedThreads = 256;

template <int NRED>
struct RedAcc {
  float4 f[NRED];
  double d[NRED][4];
  int pending;
  __device__ __forceinline__ void init() {
#pragma unroll
    for (int r = 0; r < NRED; ++r) {
      f[r] = make_float4(0.f, 0.f, 0.f, 0.f);
      d[r][0] = d[r][1] = d[r][2] = d[r][3] = 0.0;
    }
    pending = 0;
  }
  __device__ __forceinline__ void flush() {
#pragma unroll
    for (int r = 0; r < NRED; ++r) {
      d[r][0] += f[r].x; d[r][1] += f[r].y; d[r][2] += f[r].z; d[r][3] += f[r].w;
      f[r] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    pending = 0;
  }
  __device__ __forceinline__ void add(int r, float4 v) {
    f[r].x += v.x; f[r].y += v.y; f[r].z += v.z; f[r].w += v.w;
  }
  __device__ __forceinline__ void step() {
    if (++pending == 64) flush();
  }
};

// F::operator()(ni, y, x, c, acc) consumes one (pixel, 4-channel group).
template <int NRED, class F>
__global__ void __launch_bounds__(kRedThreads)
chan_reduce_kernel(F f, int n, int h, int w, int c4, int G, int per_image, double* __restrict__ out) {
  __shared__ double sm[NRED * 4][kRedThreads];
  const int L = kRedThreads / G;
  const int gl = threadIdx.x % G, pl = threadIdx.x / G;
  const int cg = blockIdx.y * G + gl;
  const bool active = (pl < L) && (cg < c4);
  const long long hw = (long long)h * w;
  const long long npix = per_image ? hw : (long long)n * hw;
  const int img0 = per_image ? blockIdx.z : 0;
  RedAcc<NRED> acc;
  acc.init();
  if (active) {
    for (long long p = (long long)blockIdx.x * L + pl; p < npix; p += (long long)gridDim.x * L) {
      const int ni = per_image ? img0 : (int)(p / hw);
      const long long q = per_image ? p : p - (long long)ni * hw;
      const int y = (int)(q / w), x = (int)(q % w);
      f(ni, y, x, cg * 4, acc);
      acc.step();
    }
    acc.flush();
  }
#pragma unroll
  for (int r = 0; r < NRED; ++r)
#pragma unroll
    for (int k = 0; k < 4; ++k) sm[r * 4 + k][threadIdx.x] = active ? acc.d[r][k] : 0.0;
  __syncthreads();
  if (out && pl == 0 && cg < c4) {
    const int C = c4 * 4;
#pragma unroll
    for (int r = 0; r < NRED; ++r)
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        double s = 0.0;
        for (int l = 0; l < L; ++l) s += sm[r * 4 + k][l * G + gl];
        atomicAdd(out + ((long long)img0 * NRED + r) * C + cg * 4 + k, s);
      }
  }
}

struct StatsF {
  EpiView x;
  __device__ __forceinline__ void operator()(int ni, int y, int xx, int c, RedAcc<2>& a) const {
    const float4 v = ld4(x.p + (long long)ni * x.sn + (long long)y * x.sy + (long long)xx * x.sx + c);
    a.add(0, v);
    a.add(1, make_float4(v.x * v.x, v.y * v.y, v.z * v.z, v.w * v.w));
  }
};

struct ColsumF {
  EpiView x;
  __device__ __forceinline__ void operator()(int ni, int y, int xx, int c, RedAcc<1>& a) const {
    a.add(0, ld4(x.p + (long long)ni * x.sn + (long long)y * x.sy + (long long)xx * x.sx + c));
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
struct GradIn {
  EpiView dy, mul, z, x;
  const float* mean;
  const float* invstd;
  const float* alpha;
  const float* beta;
  int act_z;
  __device__ __forceinline__ void load(int ni, int y, int xx, int c, float4& g, float4& xhat, float4& xv) const {
    g = ld4(dy.p + (long long)ni * dy.sn + (long long)y * dy.sy + (long long)xx * dy.sx + c);
    if (mul.p) {
      const float4 m = ld4(mul.p + (long long)ni * mul.sn + (long long)y * mul.sy + (long long)xx * mul.sx + c);
      g.x *= m.x; g.y *= m.y; g.z *= m.z; g.w *= m.w;
    }
    xv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (x.p) xv = ld4(x.p + (long long)ni * x.sn + (long long)y * x.sy + (long long)xx * x.sx + c);
    if (act_z) {
      float4 zz;
      if (z.p) {
        zz = ld4(z.p + (long long)ni * z.sn + (long long)y * z.sy + (long long)xx * z.sx + c);
      } else {
        const float4 a = ld4(alpha + c), b = ld4(beta + c);
        zz = make_float4(epi_act(act_z, a.x * xv.x + b.x), epi_act(act_z, a.y * xv.y + b.y),
                         epi_act(act_z, a.z * xv.z + b.z), epi_act(act_z, a.w * xv.w + b.w));
      }
      g.x *= act_grad(act_z, zz.x); g.y *= act_grad(act_z, zz.y);
      g.z *= act_grad(act_z, zz.z); g.w *= act_grad(act_z, zz.w);
    }
    if (mean) {
      const float4 mu = ld4(mean + c), is = ld4(invstd + c);
      xhat = make_float4((xv.x - mu.x) * is.x, (xv.y - mu.y) * is.y, (xv.z - mu.z) * is.z, (xv.w - mu.w) * is.w);
    } else {
      xhat = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
};

struct BnBwdReduceF {
  GradIn in;
  __device__ __forceinline__ void operator()(int ni, int y, int xx, int c, RedAcc<2>& a) const {
    float4 g, xh, xv;
    in.load(ni, y, xx, c, g, xh, xv);
    a.add(0, g);
    a.add(1, make_float4(g.x * xh.x, g.y * xh.y, g.z * xh.z, g.w * xh.w));
  }
};

// With BN (in.mean != NULL): dx = gamma*invstd*(g - S1/M - xhat*S2/M) [* leaky'(x)].
// Without BN: dx = g [* leaky'(x)]   (plain activation backward, e.g. the conv->LeakyReLU shortcuts).
struct BnBwdApplyF {
  GradIn in;
  const float* gamma;
  const double* red;
  double inv_count;
  int C, leaky_x, round_out;
  float* dx;
  long long d_sn, d_sy, d_sx;
  float* g_out;
  long long g_sn, g_sy, g_sx;
  int g_accumulate;
  __device__ __forceinline__ void operator()(int ni, int y, int xx, int c, RedAcc<1>& a) const {
    float4 g, xh, xv;
    in.load(ni, y, xx, c, g, xh, xv);
    if (g_out) {
      float* gp = g_out + (long long)ni * g_sn + (long long)y * g_sy + (long long)xx * g_sx + c;
      float4 o = g;
      if (g_accumulate) {
        const float4 e = *reinterpret_cast<const float4*>(gp);
        o.x += e.x; o.y += e.y; o.z += e.z; o.w += e.w;
      }
      *reinterpret_cast<float4*>(gp) = o;
    }
    float4 d = g;
    if (in.mean) {
      const float4 gm = ld4(gamma + c), is = ld4(in.invstd + c);
      const float m1[4] = {(float)(red[c] * inv_count), (float)(red[c + 1] * inv_count), (float)(red[c + 2] * inv_count),
                           (float)(red[c + 3] * inv_count)};
      const float m2[4] = {(float)(red[C + c] * inv_count), (float)(red[C + c + 1] * inv_count),
                           (float)(red[C + c + 2] * inv_count), (float)(red[C + c + 3] * inv_count)};
      d.x = gm.x * is.x * (g.x - m1[0] - xh.x * m2[0]);
      d.y = gm.y * is.y * (g.y - m1[1] - xh.y * m2[1]);
      d.z = gm.z * is.z * (g.z - m1[2] - xh.z * m2[2]);
      d.w = gm.w * is.w * (g.w - m1[3] - xh.w * m2[3]);
    }
    if (leaky_x) {
      d.x *= xv.x > 0.f ? 1.f : 0.01f; d.y *= xv.y > 0.f ? 1.f : 0.01f;
      d.z *= xv.z > 0.f ? 1.f : 0.01f; d.w *= xv.w > 0.f ? 1.f : 0.01f;
    }
    if (round_out) {
      d.x = round_tf32(d.x); d.y = round_tf32(d.y); d.z = round_tf32(d.z); d.w = round_tf32(d.w);
    }
    if (dx) *reinterpret_cast<float4*>(dx + (long long)ni * d_sn + (long long)y * d_sy + (long long)xx * d_sx + c) = d;
    a.add(0, d);
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

static inline EpiView ev(const pmfb_view* v) {
  EpiView e;
  e.p = v ? v->ptr : nullptr;
  e.sn = v ? v->sn : 0;
  e.sy = v ? v->sy : 0;
  e.sx = v ? v->sx : 0;
  return e;
}

static inline bool vok(const pmfb_view* v) {
  return !v || !v->ptr || ((((v->sn | v->sy | v->sx) % 4) == 0) && ((reinterpret_cast<uintptr_t>(v->ptr) & 15) == 0));
}

struct RedGrid {
  dim3 grid;
  int G;
};
static inline RedGrid red_grid(long long npix, int c4, int images) {
  RedGrid r;
  r.G = c4 < kRedThreads ? c4 : kRedThreads;
  const int L = kRedThreads / r.G;
  const int gy = (c4 + r.G - 1) / r.G;
  long long gx = (npix + (long long)L * 8 - 1) / ((long long)L * 8);  // >= 8 pixels per thread
  long long cap = (148 * 8) / ((long long)gy * images);
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
  RedGrid g = red_grid(npix, c / 4, 1);
  StatsF f{ev(x)};
  chan_reduce_kernel<2, StatsF><<<g.grid, kRedThreads, 0, (cudaStream_t)stream>>>(f, n, h, w, c / 4, g.G, 0, sums);
  PMFB_LAUNCH_CHECK("bn_stats");
  return PMFB_OK;
}

extern "C" int pmfb_colsum(const pmfb_view* x, int32_t n, int32_t h, int32_t w, int32_t c, int32_t per_image, double* out,
                           void* stream) {
  REQ(x && x->ptr && vok(x) && out && c > 0 && c % 4 == 0, "colsum: bad arguments (c=%d)", c);
  const long long npix = per_image ? (long long)h * w : (long long)n * h * w;
  if (npix == 0 || n == 0) return PMFB_OK;
  RedGrid g = red_grid(npix, c / 4, per_image ? n : 1);
  ColsumF f{ev(x)};
  chan_reduce_kernel<1, ColsumF><<<g.grid, kRedThreads, 0, (cudaStream_t)stream>>>(f, n, h, w, c / 4, g.G, per_image, out);
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

static int fill_grad_in(GradIn* gi, const pmfb_view* dy, const pmfb_view* mul, const pmfb_view* z, int act_z,
                        const pmfb_view* x, const float* mean, const float* invstd, const float* alpha, const float* beta) {
  if (!dy || !dy->ptr || !vok(dy) || !vok(mul) || !vok(z) || !vok(x)) return fail(PMFB_ERR_INVALID, "bn_bwd: bad views");
  if (act_z < 0 || act_z > 3) return fail(PMFB_ERR_INVALID, "bn_bwd: act_z=%d", act_z);
  if (act_z && !(z && z->ptr) && !(alpha && beta && x && x->ptr))
    return fail(PMFB_ERR_INVALID, "bn_bwd: act_z without z needs x, alpha and beta to recompute it");
  if ((mean != nullptr) != (invstd != nullptr)) return fail(PMFB_ERR_INVALID, "bn_bwd: mean and invstd go together");
  if (mean && !(x && x->ptr)) return fail(PMFB_ERR_INVALID, "bn_bwd: BN backward needs x");
  gi->dy = ev(dy);
  gi->mul = ev(mul);
  gi->z = ev(z);
  gi->x = ev(x);
  gi->mean = mean;
  gi->invstd = invstd;
  gi->alpha = alpha;
  gi->beta = beta;
  gi->act_z = act_z;
  return PMFB_OK;
}

extern "C" int pmfb_bn_bwd_reduce(const pmfb_view* dy, const pmfb_view* mul, const pmfb_view* z, int32_t act_z,
                                  const pmfb_view* x, const float* mean, const float* invstd, const float* alpha,
                                  const float* beta, int32_t n, int32_t h, int32_t w, int32_t c, double* red,
                                  void* stream) {
  REQ(red && mean && c > 0 && c % 4 == 0, "bn_bwd_reduce: bad arguments");
  BnBwdReduceF f;
  int rc = fill_grad_in(&f.in, dy, mul, z, act_z, x, mean, invstd, alpha, beta);
  if (rc) return rc;
  const long long npix = (long long)n * h * w;
  if (npix == 0) return PMFB_OK;
  RedGrid g = red_grid(npix, c / 4, 1);
  chan_reduce_kernel<2, BnBwdReduceF><<<g.grid, kRedThreads, 0, (cudaStream_t)stream>>>(f, n, h, w, c / 4, g.G, 0, red);
  PMFB_LAUNCH_CHECK("bn_bwd_reduce");
  return PMFB_OK;
}

extern "C" int pmfb_bn_bwd_apply(const pmfb_view* dy, const pmfb_view* mul, const pmfb_view* z, int32_t act_z,
                                 const pmfb_view* x, const float* mean, const float* invstd, const float* alpha,
                                 const float* beta, const float* gamma, const double* red, int32_t leaky_x, int32_t n,
                                 int32_t h, int32_t w, int32_t c, float* dx, int64_t d_sn, int64_t d_sy, int64_t d_sx,
                                 int32_t round_out, float* dgamma, float* dbeta, double* colsum, float* g_out, int64_t g_sn,
                                 int64_t g_sy, int64_t g_sx, int32_t g_accumulate, void* stream) {
  REQ(c > 0 && c % 4 == 0, "bn_bwd_apply: c=%d", c);
  REQ(!mean || (gamma && red), "bn_bwd_apply: BN backward needs gamma and red");
  REQ(!leaky_x || (x && x->ptr), "bn_bwd_apply: leaky_x needs x");
  REQ(!dx || ((((d_sn | d_sy | d_sx) % 4) == 0) && ((reinterpret_cast<uintptr_t>(dx) & 15) == 0)), "bn_bwd_apply: bad dx view");
  REQ(!g_out || ((((g_sn | g_sy | g_sx) % 4) == 0) && ((reinterpret_cast<uintptr_t>(g_out) & 15) == 0)),
      "bn_bwd_apply: bad g_out view");
  BnBwdApplyF f;
  int rc = fill_grad_in(&f.in, dy, mul, z, act_z, x, mean, invstd, alpha, beta);
  if (rc) return rc;
  const long long npix = (long long)n * h * w;
  if (npix == 0) return PMFB_OK;
  f.gamma = gamma;
  f.red = red;
  f.inv_count = 1.0 / (double)npix;
  f.C = c;
  f.leaky_x = leaky_x;
  f.round_out = round_out;
  f.dx = dx;
  f.d_sn = d_sn; f.d_sy = d_sy; f.d_sx = d_sx;
  f.g_out = g_out;
  f.g_sn = g_sn; f.g_sy = g_sy; f.g_sx = g_sx;
  f.g_accumulate = g_accumulate;
  RedGrid g = red_grid(npix, c / 4, 1);
  chan_reduce_kernel<1, BnBwdApplyF><<<g.grid, kRedThreads, 0, (cudaStream_t)stream>>>(f, n, h, w, c / 4, g.G, 0, colsum);
  PMFB_LAUNCH_CHECK("bn_bwd_apply");
  if (mean && (dgamma || dbeta)) {
    red_to_params_kernel<<<(c + 127) / 128, 128, 0, (cudaStream_t)stream>>>(red, c, dgamma, dbeta);
    PMFB_LAUNCH_CHECK("red_to_params");
  }
  return PMFB_OK;
}
