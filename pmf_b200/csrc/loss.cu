// Fused loss head of the PMF trainer: what tasks/pmf/trainer.py:305-332 computes on the two softmax maps
// with ~40 elementwise ATen kernels and 2 x 19 full-tensor sorts per step.
//
//   pmfb_loss_head : ONE pass over the two (B, C, H, W) probability maps and the labels: log / entropy / confidence, the
//                    focal term of both heads (pc_processor/loss/focal_softmax.py:28-62), both KL directions of the
//                    perception-aware loss with their guide weights (trainer.py:231-252), forward sums AND the gradient
//                    with respect to both maps (written, dense NCHW).
//   pmfb_lovasz    : Lovasz-softmax of both heads (pc_processor/loss/lovasz_softmax.py:55-145, classes="present",
//                    ignore label dropped): compaction of the labelled pixels, ONE radix sort of
//                    (head, class, descending error, pixel) 64-bit keys, a scan of the foreground flags, the Jaccard
//                    gradient per sorted position, the class losses and the gradient scattered (accumulated) into the
//                    same gradient maps.  No host synchronisation: grids are sized for the worst case and read the
//                    number of labelled pixels from device memory.
//
// All kernels are HBM-bound integer / elementwise work: coalesced channel-plane reads of the NCHW maps, shared-memory
// histograms and warp match/ballot ranking for the sort; grids are multiples of the SM count.
#include <stdint.h>

#include "common.h"

namespace pmfb {

static inline int lgrid(long long work, int threads, int per_sm = 8) {
  long long b = (work + threads - 1) / threads;
  if (b < 1) b = 1;
  const long long cap = (long long)sm_count() * per_sm;
  return (int)(b > cap ? cap : b);
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}

// block-wide sum of NV doubles per thread into global accumulators (one atomic per value per block)
template <int NV>
__device__ __forceinline__ void block_accumulate(double (&v)[NV], double* __restrict__ dst) {
  __shared__ double sm[NV][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const double s = warp_sum(v[k]);
    if (lane == 0) sm[k][warp] = s;
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      double s = lane < nw ? sm[k][lane] : 0.0;
      s = warp_sum(s);
      if (lane == 0 && s != 0.0) atomicAdd(dst + k, s);
    }
  }
}

// ------------------------------------------------------------------------------------------------ labelled-pixel count
__global__ void loss_count_kernel(const long long* __restrict__ label, long long n_pix, double* __restrict__ sums) {
  double v[1] = {0.0};
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n_pix; i += (long long)gridDim.x * blockDim.x)
    v[0] += label[i] > 0 ? 1.0 : 0.0;
  block_accumulate<1>(v, sums + 2);
}

// ------------------------------------------------------------------------------------------------ fused head
// sums: [0] focal lidar (un-normalised sum), [1] focal camera, [2] labelled pixels (label > 0), [3] sum of w_img * KL(p_cam || p_lidar)
//       terms (loss_per_pcd numerator), [4] loss_per_img numerator, [5] sum of lidar entropies, [6] camera entropies.
template <int CMAX>
__global__ void __launch_bounds__(256)
loss_head_kernel(const float* __restrict__ pl, const float* __restrict__ pc, const long long* __restrict__ label, int n, int c,
                 long long hw, const float* __restrict__ alpha, float focal_gamma, float tau, float w_per, float w_focal,
                 float* __restrict__ dl, float* __restrict__ dc, double* __restrict__ sums) {
  const long long n_pix = (long long)n * hw;
  const float inv_logc = 1.f / logf((float)c);
  const double n_lab = sums[2];  // written by loss_count_kernel (stream order)
  const float focal_scale = n_lab > 0.0 ? (float)((double)w_focal / n_lab) : 0.f;
  const float per_scale = w_per / (float)((double)n_pix * (double)c);  // mean over ALL B*C*H*W elements (trainer.py:247-250)
  double acc[7] = {0, 0, 0, 0, 0, 0, 0};
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n_pix; i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / hw, r = i - b * hw;
    const float* xl = pl + b * c * hw + r;
    const float* xc = pc + b * c * hw + r;
    float vl[CMAX], vc[CMAX];
#pragma unroll
    for (int k = 0; k < CMAX; ++k)
      if (k < c) {
        vl[k] = __ldg(xl + (long long)k * hw);
        vc[k] = __ldg(xc + (long long)k * hw);
      }
    // pass 1: entropies (trainer.py:305-319) and the two KL sums (trainer.py:247-250; KLDivLoss(reduction="none")(log a, b) =
    // xlogy(b, b) - b * log a)
    float sl = 0.f, sc = 0.f, A = 0.f, Bk = 0.f;
    float lgl[CMAX], lgc[CMAX];  // log(clamp(p, 1e-8)); equals log p unless p < 1e-8 (then xlogy's log p is taken separately)
#pragma unroll
    for (int k = 0; k < CMAX; ++k)
      if (k < c) {
        const float ll = logf(fmaxf(vl[k], 1e-8f)), lc = logf(fmaxf(vc[k], 1e-8f));
        lgl[k] = ll;
        lgc[k] = lc;
        sl += vl[k] * ll;
        sc += vc[k] * lc;
        const float rl = vl[k] >= 1e-8f ? ll : (vl[k] > 0.f ? logf(vl[k]) : 0.f);  // log p of xlogy(p, p)
        const float rc = vc[k] >= 1e-8f ? lc : (vc[k] > 0.f ? logf(vc[k]) : 0.f);
        A += vc[k] * rc - vc[k] * ll;   // KL term pulling the LiDAR head towards the camera head
        Bk += vl[k] * rl - vl[k] * lc;
      }
    const float conf_l = 1.f + sl * inv_logc, conf_c = 1.f + sc * inv_logc;  // 1 - entropy / log C
    const float imp = conf_l - conf_c;
    const bool on_l = imp > 0.f && conf_l >= tau;   // pcd_guide_weight active: weights the camera head's KL (loss_per_img)
    const bool on_c = imp < 0.f && conf_c >= tau;   // img_guide_weight active: weights the LiDAR head's KL (loss_per_pcd)
    const float w_l = on_l ? fabsf(imp) : 0.f, w_c = on_c ? fabsf(imp) : 0.f;
    acc[3] += (double)(w_c * A);
    acc[4] += (double)(w_l * Bk);
    acc[5] += (double)(-sl * inv_logc);
    acc[6] += (double)(-sc * inv_logc);
    // d(w_c A + w_l Bk) / d imp: w_l = imp (imp > 0), w_c = -imp (imp < 0) where active
    const float dimp = (on_l ? Bk : 0.f) - (on_c ? A : 0.f);
    const long long t = label[i];
    const bool lab = t > 0;
    float al = 0.f;
    if (lab) al = __ldg(alpha + t);
    if (dl != nullptr) {
      float* gl = dl + b * c * hw + r;
      float* gc = dc + b * c * hw + r;
#pragma unroll
      for (int k = 0; k < CMAX; ++k)
        if (k < c) {
          const float p = vl[k], q = vc[k];
          const float ll = lgl[k], lc = lgc[k];
          const float cl_p = p >= 1e-8f ? 1.f : 0.f, cl_q = q >= 1e-8f ? 1.f : 0.f;  // d log(clamp(x)) / dx = [x >= 1e-8] / x
          const float rl1 = p >= 1e-8f ? ll + 1.f : (p > 0.f ? logf(p) + 1.f : 0.f);  // d xlogy(p, p) / dp = log p + 1
          const float rc1 = q >= 1e-8f ? lc + 1.f : (q > 0.f ? logf(q) + 1.f : 0.f);
          // perception-aware terms
          float g_l = w_c * (-q * cl_p / fmaxf(p, 1e-8f)) + w_l * (rl1 - lc) + dimp * (ll + cl_p) * inv_logc;
          float g_c = w_l * (-p * cl_q / fmaxf(q, 1e-8f)) + w_c * (rc1 - ll) - dimp * (lc + cl_q) * inv_logc;
          g_l *= per_scale;
          g_c *= per_scale;
          if (lab && k == (int)t) {  // focal: -(1 - p)^g * log(clamp(p, 1e-6)) * alpha[t], mean over the labelled pixels
            const float pt_l = p, pt_c = q;
            const float lg_l = pt_l >= 1e-6f ? ll : logf(1e-6f), lg_c = pt_c >= 1e-6f ? lc : logf(1e-6f);
            const float om_l = 1.f - pt_l, om_c = 1.f - pt_c;
            const float pw_l = focal_gamma == 2.f ? om_l * om_l : powf(om_l, focal_gamma);
            const float pw_c = focal_gamma == 2.f ? om_c * om_c : powf(om_c, focal_gamma);
            const float dpw_l = focal_gamma == 2.f ? 2.f * om_l : focal_gamma * powf(om_l, focal_gamma - 1.f);
            const float dpw_c = focal_gamma == 2.f ? 2.f * om_c : focal_gamma * powf(om_c, focal_gamma - 1.f);
            acc[0] += (double)(-pw_l * lg_l * al);
            acc[1] += (double)(-pw_c * lg_c * al);
            g_l += focal_scale * al * (dpw_l * lg_l - (pt_l >= 1e-6f ? pw_l / pt_l : 0.f));
            g_c += focal_scale * al * (dpw_c * lg_c - (pt_c >= 1e-6f ? pw_c / pt_c : 0.f));
          }
          gl[(long long)k * hw] = g_l;
          gc[(long long)k * hw] = g_c;
        }
    } else if (lab) {
      const int k = (int)t;
      const float pt_l = __ldg(xl + (long long)k * hw), pt_c = __ldg(xc + (long long)k * hw);
      const float om_l = 1.f - pt_l, om_c = 1.f - pt_c;
      acc[0] += (double)(-(focal_gamma == 2.f ? om_l * om_l : powf(om_l, focal_gamma)) * logf(fmaxf(pt_l, 1e-6f)) * al);
      acc[1] += (double)(-(focal_gamma == 2.f ? om_c * om_c : powf(om_c, focal_gamma)) * logf(fmaxf(pt_c, 1e-6f)) * al);
    }
  }
  acc[2] = 0.0;
  block_accumulate<7>(acc, sums);
}

// ------------------------------------------------------------------------------------------------ Lovasz pipeline
constexpr int kPayBits = 26;                 // payload: (compact pixel index << 1) | foreground; index < 2^25
constexpr int kErrShift = kPayBits;          // 32 bits of inverted error
constexpr int kSegShift = kPayBits + 32;     // 6 bits: head * C + class
constexpr int kTile = 4096;                  // keys per block and radix pass
constexpr int kRadixThreads = 256;

struct LovCtl {          // device-resident control block (zeroed by the caller's memset at the start of the call)
  unsigned int n_valid;          // labelled (non-ignored) pixels P
  unsigned int n_present;        // classes present among them
  unsigned int class_count[32];  // foreground pixels per class
  double class_loss[64];         // [head * 32 + class]
};

// K1: compaction of the pixels whose label != ignore (order inside a block preserved, blocks in arrival order: the order
// only breaks exact error ties, which leaves the loss value unchanged) + class histogram.
__global__ void __launch_bounds__(256)
lov_compact_kernel(const long long* __restrict__ label, long long n_pix, int ignore, int c, LovCtl* __restrict__ ctl,
                   int* __restrict__ valid_idx) {
  __shared__ unsigned int s_hist[32];
  __shared__ unsigned int s_warp[8];
  __shared__ unsigned int s_base;
  if (threadIdx.x < 32) s_hist[threadIdx.x] = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (long long base = (long long)blockIdx.x * 256; base < n_pix; base += (long long)gridDim.x * 256) {
    const long long i = base + threadIdx.x;
    long long t = ignore;
    if (i < n_pix) t = label[i];
    const bool ok = (i < n_pix) && t != ignore && t >= 0 && t < c;
    const unsigned int m = __ballot_sync(0xffffffffu, ok);
    if (lane == 0) s_warp[warp] = __popc(m);
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned int tot = 0;
      for (int w = 0; w < 8; ++w) {
        const unsigned int v = s_warp[w];
        s_warp[w] = tot;
        tot += v;
      }
      s_base = tot ? atomicAdd(&ctl->n_valid, tot) : 0u;
    }
    __syncthreads();
    if (ok) {
      valid_idx[s_base + s_warp[warp] + __popc(m & ((1u << lane) - 1u))] = (int)i;
      atomicAdd(&s_hist[(int)t], 1u);
    }
    __syncthreads();
  }
  if (threadIdx.x < 32 && s_hist[threadIdx.x]) atomicAdd(&ctl->class_count[threadIdx.x], s_hist[threadIdx.x]);
}

__global__ void lov_present_kernel(LovCtl* __restrict__ ctl, int c) {
  if (threadIdx.x == 0) {
    unsigned int n = 0;
    for (int k = 0; k < c; ++k) n += ctl->class_count[k] ? 1u : 0u;
    ctl->n_present = n;
  }
}

// K2: keys.  Element (head, class k, compact pixel j) -> (seg = head*C + k) << 58 | ~bits(|fg - p|) << 26 | j << 1 | fg.
// Ascending order of the key = segments in order, errors DESCENDING inside a segment (lovasz_softmax.py:114).
__global__ void __launch_bounds__(256)
lov_keys_kernel(const float* __restrict__ p0, const float* __restrict__ p1, const long long* __restrict__ label, int n_heads, int c,
                long long hw, const LovCtl* __restrict__ ctl, const int* __restrict__ valid_idx, unsigned long long* __restrict__ keys) {
  const long long P = ctl->n_valid;
  const long long total = P * c * n_heads;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long j = e % P;
    const int seg = (int)(e / P);
    const int head = seg / c, k = seg - head * c;
    const long long pix = valid_idx[j];
    const long long b = pix / hw, r = pix - b * hw;
    const float* p = head ? p1 : p0;
    const float v = __ldg(p + (b * c + k) * hw + r);
    const unsigned int fg = label[pix] == k ? 1u : 0u;
    const float err = fabsf((float)fg - v);
    const unsigned long long inv = (unsigned long long)(~__float_as_uint(err));
    keys[e] = ((unsigned long long)seg << kSegShift) | (inv << kErrShift) | ((unsigned long long)j << 1) | fg;
  }
}

// K3: LSD radix sort, 8-bit digits.  hist[digit * nblk_max + block].
__global__ void __launch_bounds__(kRadixThreads)
radix_hist_kernel(const unsigned long long* __restrict__ keys, const LovCtl* __restrict__ ctl, int per_pixel, int shift,
                  unsigned int* __restrict__ hist, int nblk_max) {
  const long long N = (long long)ctl->n_valid * per_pixel;
  const long long start = (long long)blockIdx.x * kTile;
  if (start >= N) return;
  __shared__ unsigned int s_h[256];
  s_h[threadIdx.x] = 0;
  __syncthreads();
  const long long end = start + kTile < N ? start + kTile : N;
  for (long long i = start + threadIdx.x; i < end; i += kRadixThreads)
    atomicAdd(&s_h[(unsigned int)(keys[i] >> shift) & 255u], 1u);
  __syncthreads();
  hist[(long long)threadIdx.x * nblk_max + blockIdx.x] = s_h[threadIdx.x];
}

// exclusive scan of every digit's row of the (digit-major) histogram over the blocks in use: one block of 1024 threads per
// digit walks its row with a running carry and leaves the row total in digit_total[digit]; the 256 totals are scanned in
// the prologue of the scatter kernel.
__global__ void __launch_bounds__(1024)
radix_scan_kernel(unsigned int* __restrict__ hist, const LovCtl* __restrict__ ctl, int per_pixel, int nblk_max,
                  unsigned int* __restrict__ digit_total) {
  const long long N = (long long)ctl->n_valid * per_pixel;
  const int nblk = (int)((N + kTile - 1) / kTile);
  unsigned int* row = hist + (long long)blockIdx.x * nblk_max;
  __shared__ unsigned int s_w[32];
  __shared__ unsigned int s_carry;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int base = 0; base < nblk; base += 1024) {
    const int e = base + threadIdx.x;
    const unsigned int v = e < nblk ? row[e] : 0u;
    unsigned int x = v;  // inclusive warp scan
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned int y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) s_w[warp] = x;
    __syncthreads();
    if (warp == 0) {
      unsigned int w = s_w[lane], z = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned int y = __shfl_up_sync(0xffffffffu, z, o);
        if (lane >= o) z += y;
      }
      s_w[lane] = z - w;  // exclusive prefix of the warp totals
    }
    __syncthreads();
    const unsigned int carry = s_carry;
    if (e < nblk) row[e] = carry + s_w[warp] + x - v;
    __syncthreads();
    if (threadIdx.x == 1023) s_carry = carry + s_w[31] + x;
    __syncthreads();
  }
  if (threadIdx.x == 0) digit_total[blockIdx.x] = s_carry;
}

// stable scatter: 16 rounds of 256 keys; per round the lanes of a warp holding the same digit are ranked with match_any,
// the warps are chained through shared memory, and the block's running digit bases advance.
__global__ void __launch_bounds__(kRadixThreads)
radix_scatter_kernel(const unsigned long long* __restrict__ in, unsigned long long* __restrict__ out, const LovCtl* __restrict__ ctl,
                     int per_pixel, int shift, const unsigned int* __restrict__ hist, int nblk_max,
                     const unsigned int* __restrict__ digit_total) {
  const long long N = (long long)ctl->n_valid * per_pixel;
  const long long start = (long long)blockIdx.x * kTile;
  if (start >= N) return;
  __shared__ unsigned int s_base[256];
  __shared__ unsigned int s_cnt[8][256];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  {  // exclusive scan of the 256 digit totals (thread = digit) + this block's offset inside the digit's row
    __shared__ unsigned int s_tw[8];
    const unsigned int v = digit_total[threadIdx.x];
    unsigned int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned int y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) s_tw[warp] = x;
    __syncthreads();
    unsigned int pre = 0;
    for (int w = 0; w < warp; ++w) pre += s_tw[w];
    s_base[threadIdx.x] = pre + x - v + hist[(long long)threadIdx.x * nblk_max + blockIdx.x];
  }
  for (int r = 0; r < kTile / kRadixThreads; ++r) {
#pragma unroll
    for (int w = 0; w < 8; ++w) s_cnt[w][threadIdx.x] = 0;
    __syncthreads();
    const long long i = start + (long long)r * kRadixThreads + threadIdx.x;
    const bool ok = i < N;
    unsigned long long key = 0;
    unsigned int d = 256u + (unsigned int)lane;  // a value no valid digit takes and no other lane shares
    if (ok) {
      key = in[i];
      d = (unsigned int)(key >> shift) & 255u;
    }
    const unsigned int m = __match_any_sync(0xffffffffu, d);
    const unsigned int rank = __popc(m & ((1u << lane) - 1u));
    if (ok && rank == 0) s_cnt[warp][d] = __popc(m);
    __syncthreads();
    {  // thread = digit: exclusive prefix over the 8 warps, then advance the running base
      unsigned int run = s_base[threadIdx.x];
#pragma unroll
      for (int w = 0; w < 8; ++w) {
        const unsigned int v = s_cnt[w][threadIdx.x];
        s_cnt[w][threadIdx.x] = run;
        run += v;
      }
      s_base[threadIdx.x] = run;
    }
    __syncthreads();
    if (ok) out[s_cnt[warp][d] + rank] = key;
    __syncthreads();
  }
}

// K4: inclusive scan of the foreground flags over the sorted keys (three phases: tile sums, scan of tile sums, apply).
__global__ void __launch_bounds__(256)
fgscan_tile_kernel(const unsigned long long* __restrict__ keys, const LovCtl* __restrict__ ctl, int per_pixel,
                   unsigned int* __restrict__ tile_sum) {
  const long long N = (long long)ctl->n_valid * per_pixel;
  const long long start = (long long)blockIdx.x * kTile;
  if (start >= N) return;
  const long long end = start + kTile < N ? start + kTile : N;
  unsigned int s = 0;
  for (long long i = start + threadIdx.x; i < end; i += 256) s += (unsigned int)(keys[i] & 1ull);
  __shared__ unsigned int sm[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int t = 0;
    for (int w = 0; w < 8; ++w) t += sm[w];
    tile_sum[blockIdx.x] = t;
  }
}

__global__ void __launch_bounds__(1024)
fgscan_sums_kernel(unsigned int* __restrict__ tile_sum, const LovCtl* __restrict__ ctl, int per_pixel) {
  const long long N = (long long)ctl->n_valid * per_pixel;
  const int nblk = (int)((N + kTile - 1) / kTile);
  __shared__ unsigned int s_w[32];
  __shared__ unsigned int s_carry;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int base = 0; base < nblk; base += 1024) {
    const int e = base + threadIdx.x;
    const unsigned int v = e < nblk ? tile_sum[e] : 0u;
    unsigned int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned int y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) s_w[warp] = x;
    __syncthreads();
    if (warp == 0) {
      unsigned int w = s_w[lane], z = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned int y = __shfl_up_sync(0xffffffffu, z, o);
        if (lane >= o) z += y;
      }
      s_w[lane] = z - w;
    }
    __syncthreads();
    const unsigned int carry = s_carry;
    if (e < nblk) tile_sum[e] = carry + s_w[warp] + x - v;  // exclusive
    __syncthreads();
    if (threadIdx.x == 1023) s_carry = carry + s_w[31] + x;
    __syncthreads();
  }
}

// K5: per sorted position: foreground count so far inside its (head, class) segment -> Jaccard gradient
// (lovasz_softmax.py:55-66), class loss contribution e * g, and the gradient scattered into d_probs:
//   d loss / d p[pixel, class] = scale / n_present * g * (fg ? -1 : +1)       (errors = |fg - p|, 0 <= p <= 1)
__global__ void __launch_bounds__(256)
lov_apply_kernel(const unsigned long long* __restrict__ keys, const unsigned int* __restrict__ tile_excl, LovCtl* __restrict__ ctl,
                 int n_heads, int c, long long hw, const int* __restrict__ valid_idx, float scale, float* __restrict__ d0,
                 float* __restrict__ d1) {
  const long long P = ctl->n_valid;
  const long long N = P * c * n_heads;
  const long long start = (long long)blockIdx.x * kTile;
  if (start >= N) return;
  __shared__ unsigned int s_w[8];
  __shared__ unsigned int s_run;
  __shared__ double s_loss[64];
  __shared__ unsigned int s_before[64];  // foreground flags held by the segments in front of segment s
  if (threadIdx.x < 64) {
    s_loss[threadIdx.x] = 0.0;
    unsigned int before = 0;
    for (int s = 0; s < (int)threadIdx.x && s < c * n_heads; ++s) before += ctl->class_count[s % c];
    s_before[threadIdx.x] = before;
  }
  if (threadIdx.x == 0) s_run = tile_excl[blockIdx.x];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float gscale = ctl->n_present ? scale / (float)ctl->n_present : 0.f;
  double my_loss = 0.0;  // this thread's loss contributions to segment my_seg (a tile spans one or two segments)
  int my_seg = -1;
  for (int r = 0; r < kTile / 256; ++r) {
    const long long i = start + (long long)r * 256 + threadIdx.x;
    const bool ok = i < N;
    unsigned long long key = 0;
    if (ok) key = keys[i];
    const unsigned int fg = (unsigned int)(key & 1ull);
    unsigned int x = ok ? fg : 0u;  // inclusive block scan of fg
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned int y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) s_w[warp] = x;
    __syncthreads();
    unsigned int pre = 0;
    for (int w = 0; w < warp; ++w) pre += s_w[w];
    const unsigned int run = s_run;
    const unsigned int cum = run + pre + x;  // foreground flags in sorted positions [0, i], over ALL segments
    __syncthreads();
    if (threadIdx.x == 255) s_run = cum;
    if (ok) {
      const int seg = (int)(key >> kSegShift);
      const int head = seg / c, k = seg - head * c;
      const unsigned int G = ctl->class_count[k];
      if (G) {  // classes="present"
        // every earlier segment holds exactly its class's G foreground flags: subtract them
        const float cf = (float)(cum - s_before[seg]);    // inclusive foreground cumsum inside the segment
        const long long pos = i - (long long)seg * P;     // 0-based position inside the segment
        const float gts = (float)G;
        const float jac = 1.f - (gts - cf) / (gts + ((float)(pos + 1) - cf));
        float g = jac;
        if (pos > 0) {
          const float cfp = cf - (float)fg;
          g = jac - (1.f - (gts - cfp) / (gts + ((float)pos - cfp)));
        }
        const float err = __uint_as_float(~(unsigned int)(key >> kErrShift));
        if (seg != my_seg) {
          if (my_seg >= 0 && my_loss != 0.0) atomicAdd(&s_loss[(my_seg / c) * 32 + (my_seg % c)], my_loss);
          my_seg = seg;
          my_loss = 0.0;
        }
        my_loss += (double)err * (double)g;
        float* d = head ? d1 : d0;
        if (d != nullptr) {
          const long long j = (long long)((key & ((1ull << kPayBits) - 1ull)) >> 1);
          const long long pix = valid_idx[j];
          const long long b = pix / hw, rr = pix - b * hw;
          d[(b * c + k) * hw + rr] += gscale * (fg ? -g : g);
        }
      }
    }
    __syncthreads();
  }
  if (my_seg >= 0 && my_loss != 0.0) atomicAdd(&s_loss[(my_seg / c) * 32 + (my_seg % c)], my_loss);
  __syncthreads();
  if (threadIdx.x < 64 && s_loss[threadIdx.x] != 0.0) atomicAdd(&ctl->class_loss[threadIdx.x], s_loss[threadIdx.x]);
}

__global__ void lov_finalize_kernel(const LovCtl* __restrict__ ctl, int n_heads, int c, double* __restrict__ loss_out) {
  if (threadIdx.x < n_heads) {
    double s = 0.0;
    for (int k = 0; k < c; ++k)
      if (ctl->class_count[k]) s += ctl->class_loss[threadIdx.x * 32 + k];
    loss_out[threadIdx.x] += ctl->n_present ? s / (double)ctl->n_present : 0.0;
  }
}

struct LovLayout {
  size_t ctl, valid, keys_a, keys_b, hist, tiles, dtot, total;
  int nblk_max;
};

static LovLayout lov_layout(long long n_pix, int c, int n_heads) {
  LovLayout L;
  const long long n_max = n_pix * c * n_heads;
  L.nblk_max = (int)((n_max + kTile - 1) / kTile);
  if (L.nblk_max < 1) L.nblk_max = 1;
  size_t o = 0;
  auto take = [&](size_t bytes) {
    const size_t at = o;
    o += (bytes + 255) & ~(size_t)255;
    return at;
  };
  L.ctl = take(sizeof(LovCtl));
  L.valid = take((size_t)n_pix * 4);
  L.keys_a = take((size_t)n_max * 8);
  L.keys_b = take((size_t)n_max * 8);
  L.hist = take((size_t)256 * L.nblk_max * 4);
  L.tiles = take((size_t)L.nblk_max * 4);
  L.dtot = take(256 * 4);
  L.total = o;
  return L;
}

}  // namespace pmfb

using namespace pmfb;

#define LREQ(cond, ...) \
  do {                  \
    if (!(cond)) return fail(PMFB_ERR_INVALID, __VA_ARGS__); \
  } while (0)

extern "C" int pmfb_loss_head(const float* p_lidar, const float* p_camera, const int64_t* label, int32_t n, int32_t c, int32_t h,
                              int32_t w, const float* alpha, float focal_gamma, float tau, float w_focal, float w_per,
                              float* d_lidar, float* d_camera, double* sums, void* stream) {
  LREQ(p_lidar && p_camera && label && alpha && sums, "loss_head: null argument");
  LREQ(n > 0 && c >= 2 && c <= 32 && h > 0 && w > 0, "loss_head: bad shape (n=%d, c=%d, h=%d, w=%d); c must be in [2, 32]", n, c, h, w);
  LREQ((d_lidar == nullptr) == (d_camera == nullptr), "loss_head: both gradient maps or none");
  const long long hw = (long long)h * w, n_pix = (long long)n * hw;
  cudaStream_t st = (cudaStream_t)stream;
  loss_count_kernel<<<lgrid(n_pix, 256, 4), 256, 0, st>>>(reinterpret_cast<const long long*>(label), n_pix, sums);
  PMFB_LAUNCH_CHECK("loss_count_kernel");
  const int grid = lgrid(n_pix, 256, 8);
  if (c <= 20)
    loss_head_kernel<20><<<grid, 256, 0, st>>>(p_lidar, p_camera, reinterpret_cast<const long long*>(label), n, c, hw, alpha, focal_gamma,
                                               tau, w_per, w_focal, d_lidar, d_camera, sums);
  else
    loss_head_kernel<32><<<grid, 256, 0, st>>>(p_lidar, p_camera, reinterpret_cast<const long long*>(label), n, c, hw, alpha, focal_gamma,
                                               tau, w_per, w_focal, d_lidar, d_camera, sums);
  PMFB_LAUNCH_CHECK("loss_head_kernel");
  return PMFB_OK;
}

extern "C" size_t pmfb_lovasz_workspace_bytes(int64_t n_pixels, int32_t c, int32_t n_heads) {
  if (n_pixels <= 0 || c <= 0 || n_heads <= 0) return 0;
  return lov_layout(n_pixels, c, n_heads).total;
}

extern "C" int pmfb_lovasz(const float* probs0, const float* probs1, const int64_t* label, int32_t n, int32_t c, int32_t h, int32_t w,
                           int32_t ignore, float grad_scale, float* d_probs0, float* d_probs1, double* loss_out, void* workspace,
                           size_t workspace_bytes, void* stream) {
  LREQ(probs0 && label && loss_out && workspace, "lovasz: null argument");
  const int n_heads = probs1 ? 2 : 1;
  LREQ(n > 0 && c >= 1 && c <= 32 && h > 0 && w > 0, "lovasz: bad shape; c must be <= 32");
  const long long hw = (long long)h * w, n_pix = (long long)n * hw;
  LREQ(n_pix < (1LL << (kPayBits - 1)), "lovasz: %lld pixels exceed the 2^25 the sort key's payload can index", n_pix);
  LREQ((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "lovasz: workspace must be 256-byte aligned");
  const LovLayout L = lov_layout(n_pix, c, n_heads);
  LREQ(workspace_bytes >= L.total, "lovasz: workspace of %zu bytes < %zu needed (pmfb_lovasz_workspace_bytes)", workspace_bytes, L.total);
  cudaStream_t st = (cudaStream_t)stream;
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  LovCtl* ctl = reinterpret_cast<LovCtl*>(ws + L.ctl);
  int* valid = reinterpret_cast<int*>(ws + L.valid);
  unsigned long long* ka = reinterpret_cast<unsigned long long*>(ws + L.keys_a);
  unsigned long long* kb = reinterpret_cast<unsigned long long*>(ws + L.keys_b);
  unsigned int* hist = reinterpret_cast<unsigned int*>(ws + L.hist);
  unsigned int* tiles = reinterpret_cast<unsigned int*>(ws + L.tiles);
  unsigned int* dtot = reinterpret_cast<unsigned int*>(ws + L.dtot);
  const int per_pixel = c * n_heads;
  const long long* lab = reinterpret_cast<const long long*>(label);

  PMFB_CUDA_CHECK(cudaMemsetAsync(ctl, 0, sizeof(LovCtl), st));
  lov_compact_kernel<<<lgrid(n_pix, 256, 8), 256, 0, st>>>(lab, n_pix, ignore, c, ctl, valid);
  PMFB_LAUNCH_CHECK("lov_compact_kernel");
  lov_present_kernel<<<1, 32, 0, st>>>(ctl, c);
  PMFB_LAUNCH_CHECK("lov_present_kernel");
  lov_keys_kernel<<<lgrid(n_pix * per_pixel, 256, 16), 256, 0, st>>>(probs0, probs1, lab, n_heads, c, hw, ctl, valid, ka);
  PMFB_LAUNCH_CHECK("lov_keys_kernel");
  // sort on key bits [26, 64): errors (32 bits) then segment (6 bits) -> five 8-bit passes
  unsigned long long* src = ka;
  unsigned long long* dst = kb;
  for (int pass = 0; pass < 5; ++pass) {
    const int shift = kErrShift + 8 * pass;
    radix_hist_kernel<<<L.nblk_max, kRadixThreads, 0, st>>>(src, ctl, per_pixel, shift, hist, L.nblk_max);
    PMFB_LAUNCH_CHECK("radix_hist_kernel");
    radix_scan_kernel<<<256, 1024, 0, st>>>(hist, ctl, per_pixel, L.nblk_max, dtot);
    PMFB_LAUNCH_CHECK("radix_scan_kernel");
    radix_scatter_kernel<<<L.nblk_max, kRadixThreads, 0, st>>>(src, dst, ctl, per_pixel, shift, hist, L.nblk_max, dtot);
    PMFB_LAUNCH_CHECK("radix_scatter_kernel");
    unsigned long long* t = src;
    src = dst;
    dst = t;
  }
  fgscan_tile_kernel<<<L.nblk_max, 256, 0, st>>>(src, ctl, per_pixel, tiles);
  PMFB_LAUNCH_CHECK("fgscan_tile_kernel");
  fgscan_sums_kernel<<<1, 1024, 0, st>>>(tiles, ctl, per_pixel);
  PMFB_LAUNCH_CHECK("fgscan_sums_kernel");
  lov_apply_kernel<<<L.nblk_max, 256, 0, st>>>(src, tiles, ctl, n_heads, c, hw, valid, grad_scale, d_probs0, d_probs1);
  PMFB_LAUNCH_CHECK("lov_apply_kernel");
  lov_finalize_kernel<<<1, 32, 0, st>>>(ctl, n_heads, c, loss_out);
  PMFB_LAUNCH_CHECK("lov_finalize_kernel");
  return PMFB_OK;
}
