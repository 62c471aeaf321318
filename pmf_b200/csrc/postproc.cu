// KNN back-projection and perspective projection kernels (HBM / latency bound gather-scatter work).
//
//   knn_vote        : pc_processor/postproc/knn.py:55-143 (KNN.forward).  One warp per point: lanes stride over the
//                     S*S window (gather of range + label, zero padded like F.unfold), k rounds of warp-wide
//                     lexicographic (distance, window index) arg-min by shuffles, then a match/ballot vote.
//   project_scatter : pc_processor/dataset/semantic_kitti/parser.py:209-227 (mapLidar2Camera) +
//                     pc_processor/dataset/perspective_view_loader.py:87-131 (scatter).  Pass 1: one thread per
//                     point, fp64 projection, atomicMax of the point index per pixel ("last writer wins" of numpy
//                     fancy assignment == highest index wins).  Pass 2: one thread per pixel gathers the winner.
#include "common.h"

namespace pmfb {

constexpr int kKnnMaxSearch = 15;  // S*S <= 225 candidates -> <= 8 per lane
constexpr int kKnnMaxPerLane = (kKnnMaxSearch * kKnnMaxSearch + 31) / 32;
constexpr int kKnnMaxK = 32;

__global__ void __launch_bounds__(256)
knn_vote_kernel(const float* __restrict__ proj_range, const long long* __restrict__ proj_argmax, int h, int w,
                const float* __restrict__ unproj_range, const long long* __restrict__ px, const long long* __restrict__ py,
                long long n_points, const float* __restrict__ inv_gauss, int search, int knn, float cutoff, int nclasses,
                long long* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const int s2 = search * search;
  const int pad = (search - 1) / 2;
  const int center = (s2 - 1) / 2;
  for (long long p = warp; p < n_points; p += nwarps) {
    const int cx = (int)px[p], cy = (int)py[p];
    const float r = unproj_range[p];
    float d[kKnnMaxPerLane];
    int lab[kKnnMaxPerLane];
#pragma unroll
    for (int j = 0; j < kKnnMaxPerLane; ++j) {
      const int idx = lane + 32 * j;
      d[j] = INFINITY;
      lab[j] = 0;
      if (idx < s2) {
        const int i = idx / search, jj = idx - i * search;
        const int y = cy + i - pad, x = cx + jj - pad;
        float nr = 0.f;  // F.unfold zero-pads both images (knn.py:80-82,115-117)
        int nl = 0;
        if (y >= 0 && y < h && x >= 0 && x < w) {
          nr = __ldg(proj_range + (long long)y * w + x);
          nl = (int)__ldg(proj_argmax + (long long)y * w + x);
        }
        if (nr < 0.f) nr = INFINITY;       // knn.py:91
        if (idx == center) nr = r;         // knn.py:94-95
        d[j] = __fmul_rn(fabsf(__fsub_rn(nr, r)), __ldg(inv_gauss + idx));  // knn.py:98-108
        lab[j] = nl;
      }
    }
    // k rounds of warp-wide (distance, index) arg-min; lane `round` keeps the winner of that round.
    // NaN distances order last (numpy/torch sort convention): map them to +inf with the highest indices.
    unsigned taken = 0;  // bit j set: this lane's candidate j already selected
    int my_label = -1;   // label (after cutoff) selected in round == lane
    for (int round = 0; round < knn; ++round) {
      float bd = INFINITY;
      int bi = 0x7fffffff;
#pragma unroll
      for (int j = 0; j < kKnnMaxPerLane; ++j) {
        const int idx = lane + 32 * j;
        if (idx < s2 && !((taken >> j) & 1u)) {
          const float dj = d[j] != d[j] ? INFINITY : d[j];
          if (dj < bd || (dj == bd && idx < bi)) {
            bd = dj;
            bi = idx;
          }
        }
      }
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        const float od = __shfl_xor_sync(0xffffffffu, bd, off);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
        if (od < bd || (od == bd && oi < bi)) {
          bd = od;
          bi = oi;
        }
      }
      // bi is the winning window index on every lane (0x7fffffff only if knn > s2)
      int sel_label = nclasses;
      float sel_d = INFINITY;
      const int owner = bi & 31, slot = bi >> 5;
      if (bi != 0x7fffffff) {
        int l = 0;
        float dd = 0.f;
        if (lane == owner) {
#pragma unroll
          for (int j = 0; j < kKnnMaxPerLane; ++j)
            if (j == slot) {
              l = lab[j];
              dd = d[j];
              taken |= 1u << j;
            }
        }
        sel_label = __shfl_sync(0xffffffffu, l, owner);
        sel_d = __shfl_sync(0xffffffffu, dd, owner);
      }
      if (cutoff > 0.f && sel_d > cutoff) sel_label = nclasses;  // knn.py:125-128
      if (bi == 0x7fffffff) sel_label = -1;
      if (lane == round) my_label = sel_label;
    }
    // vote over classes 1..nclasses-1, first maximum wins, all-zero -> class 1 (knn.py:132-138)
    const bool voter = my_label >= 1 && my_label < nclasses;
    const unsigned peers = __match_any_sync(0xffffffffu, voter ? my_label : -1 - lane);
    int cnt = voter ? __popc(peers) : 0;
    int cls = voter ? my_label : 0x7fffffff;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      const int oc = __shfl_xor_sync(0xffffffffu, cnt, off);
      const int ol = __shfl_xor_sync(0xffffffffu, cls, off);
      if (oc > cnt || (oc == cnt && ol < cls)) {
        cnt = oc;
        cls = ol;
      }
    }
    if (lane == 0) out[p] = cnt > 0 ? (long long)cls : 1;
  }
}

// ---------------------------------------------------------------------------------------------- batched KNN
// One THREAD per point, any number of frames per launch (the reference is un-batched, knn.py:56-59, and a
// single frame is launch-latency bound).  The S*S gathers of a point are independent (memory-level parallelism instead of
// the warp kernel's 5-deep shuffle chains), the k best (distance, window index) pairs live in a sorted register array
// (strict insertion => the lower window index wins a tie, the oracle's rule), the cutoff is applied when a candidate is
// read, and the vote is k^2 register compares.  Same fp32 operations as knn.py:91-108 (no FMA contraction).
constexpr int kKnnRegK = 8;

template <int KMAX>
__global__ void __launch_bounds__(128)
knn_vote_points_kernel(const float* __restrict__ proj_range, const long long* __restrict__ proj_argmax, int n_frames, int h, int w,
                       const float* __restrict__ unproj_range, const long long* __restrict__ px, const long long* __restrict__ py,
                       const long long* __restrict__ offsets, long long n_points, const float* __restrict__ inv_gauss, int search,
                       int knn, float cutoff, int nclasses, long long* __restrict__ out) {
  const int s2 = search * search;
  const int pad = (search - 1) / 2;
  const int center = (s2 - 1) / 2;
  const long long hw = (long long)h * w;
  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < n_points; p += (long long)gridDim.x * blockDim.x) {
    int f = 0;
    if (offsets != nullptr && n_frames > 1) {  // frame of this point: last f with offsets[f] <= p
      int lo = 0, hi = n_frames - 1;
      while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (__ldg(offsets + mid) <= p) lo = mid; else hi = mid - 1;
      }
      f = lo;
    }
    const float* rng = proj_range + f * hw;
    const long long* amx = proj_argmax + f * hw;
    const int cx = (int)px[p], cy = (int)py[p];
    const float r = unproj_range[p];
    float bd[KMAX];
    int bl[KMAX];
#pragma unroll
    for (int j = 0; j < KMAX; ++j) {
      bd[j] = INFINITY;
      bl[j] = -1;
    }
    int filled = 0;
    for (int i = 0; i < search; ++i) {
      const int y = cy + i - pad;
      const bool yin = y >= 0 && y < h;
      for (int jj = 0; jj < search; ++jj) {
        const int idx = i * search + jj;
        const int x = cx + jj - pad;
        float nr = 0.f;  // F.unfold zero-pads both images (knn.py:80-82,115-117)
        int nl = 0;
        if (yin && x >= 0 && x < w) {
          nr = __ldg(rng + (long long)y * w + x);
          nl = (int)__ldg(amx + (long long)y * w + x);
        }
        if (nr < 0.f) nr = INFINITY;   // knn.py:91
        if (idx == center) nr = r;     // knn.py:94-95
        const float d = __fmul_rn(fabsf(__fsub_rn(nr, r)), __ldg(inv_gauss + idx));  // knn.py:98-108
        if (cutoff > 0.f && d > cutoff) nl = nclasses;                               // knn.py:125-128
        float cd = d != d ? INFINITY : d;  // NaN orders last
        int cl = nl;
        bool shifting = false;
#pragma unroll
        for (int j = 0; j < KMAX; ++j) {
          if (j < knn) {
            if (shifting || j >= filled || cd < bd[j]) {
              const float td = bd[j];
              const int tl = bl[j];
              bd[j] = cd;
              bl[j] = cl;
              cd = td;
              cl = tl;
              shifting = true;
            }
          }
        }
        filled = filled < knn ? filled + 1 : knn;
      }
    }
    // vote over classes 1..nclasses-1, first (lowest) maximum wins, no vote -> class 1 (knn.py:132-138)
    int best_cnt = 0, best_cls = 1;
#pragma unroll
    for (int j = 0; j < KMAX; ++j) {
      if (j < knn && bl[j] >= 1 && bl[j] < nclasses) {
        int cnt = 0;
#pragma unroll
        for (int i = 0; i < KMAX; ++i) cnt += (i < knn && bl[i] == bl[j]) ? 1 : 0;
        if (cnt > best_cnt || (cnt == best_cnt && bl[j] < best_cls)) {
          best_cnt = cnt;
          best_cls = bl[j];
        }
      }
    }
    out[p] = (long long)best_cls;
  }
}

// ---------------------------------------------------------------------------------------------- inference tail
// argmax over the class planes of a dense NCHW probability map inside a crop window (infer.py:107-110: crop back to the
// un-padded frame, then argmax; first maximum wins like torch.argmax) + the winning probability (the per-point
// "confidence" the nuScenes merge compares, pmf_eval_nuscenes/infer.py:139-150).
__global__ void __launch_bounds__(256)
argmax_nchw_kernel(const float* __restrict__ probs, int n, int c, int h, int w, int y0, int x0, int oh, int ow,
                   long long* __restrict__ label, float* __restrict__ conf) {
  const long long total = (long long)n * oh * ow;
  const long long hw = (long long)h * w;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % ow);
    long long r = i / ow;
    const int y = (int)(r % oh);
    const int b = (int)(r / oh);
    const float* p = probs + (long long)b * c * hw + (long long)(y + y0) * w + (x + x0);
    float best = __ldg(p);
    int arg = 0;
    for (int k = 1; k < c; ++k) {
      const float v = __ldg(p + (long long)k * hw);
      if (v > best) {
        best = v;
        arg = k;
      }
    }
    label[i] = arg;
    if (conf) conf[i] = best;
  }
}

// labels -> original dataset ids through class_map_lut_inv (infer.py:129)
__global__ void lut_remap_kernel(const long long* __restrict__ labels, long long n, const int* __restrict__ lut, int lut_size,
                                 int* __restrict__ out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long l = labels[i];
    out[i] = (l >= 0 && l < lut_size) ? __ldg(lut + l) : 0;
  }
}

// 6-camera merge (pmf_eval_nuscenes/infer.py:18-38): per point the prediction of the camera with the highest confidence;
// ties -> the lowest camera (torch.argmax), points no camera saw -> -1.  Pass 1: 64-bit atomicMax of
// (confidence bits | inverted camera | entry) per point; pass 2: decode.
__global__ void merge_cameras_pass1(const long long* __restrict__ point_idx, const float* __restrict__ conf, const int* __restrict__ cam,
                                    long long n_entries, long long pc_size, unsigned long long* __restrict__ best) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n_entries; i += (long long)gridDim.x * blockDim.x) {
    const long long j = point_idx[i];
    if (j < 0 || j >= pc_size) continue;
    const float c = conf[i];
    // merge_conf starts at 0: a camera wins with a positive confidence; when every row is 0 torch.argmax picks row 0, i.e.
    // camera 0's entry if it saw the point (even with confidence 0), else -1
    if (!(c > 0.f) && !(c == 0.f && cam[i] == 0)) continue;
    const unsigned long long key = ((unsigned long long)__float_as_uint(c) << 32) | ((unsigned long long)(7 - cam[i]) << 28) |
                                   (unsigned long long)(i & 0x0FFFFFFF);
    atomicMax(best + j, key);
  }
}

__global__ void merge_cameras_pass2(const unsigned long long* __restrict__ best, const long long* __restrict__ argmax,
                                    long long pc_size, long long* __restrict__ merged) {
  for (long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x; j < pc_size; j += (long long)gridDim.x * blockDim.x) {
    const unsigned long long key = best[j];
    merged[j] = key != 0ull ? argmax[(long long)(key & 0x0FFFFFFFull)] : -1;  // low 28 bits: the winning entry (cam bits make key != 0)
  }
}

// ---------------------------------------------------------------------------------------------- confusion matrix
// IOUEval.addBatch (pc_processor/metrics/iou_eval.py:31-57): conf[pred, target] += 1 for every pixel.  Per-block histogram
// in shared memory (c*c <= 4096 counters), one global atomic per touched cell and block.
__global__ void __launch_bounds__(256)
confusion_add_kernel(const long long* __restrict__ pred, const long long* __restrict__ target, long long n, int c,
                     unsigned long long* __restrict__ conf) {
  extern __shared__ unsigned int s_conf[];
  const int cc = c * c;
  for (int i = threadIdx.x; i < cc; i += blockDim.x) s_conf[i] = 0;
  __syncthreads();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long p = pred[i], t = target[i];
    if (p >= 0 && p < c && t >= 0 && t < c) atomicAdd(&s_conf[(int)p * c + (int)t], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < cc; i += blockDim.x)
    if (s_conf[i]) atomicAdd(conf + i, (unsigned long long)s_conf[i]);
}

// ---------------------------------------------------------------------------------------------- projection
__global__ void fill_i32_kernel(int* p, long long n, int v) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) p[i] = v;
}

struct ProjM {
  double m[12];
};

__global__ void __launch_bounds__(256)
project_points_kernel(const float* __restrict__ points, long long n_points, ProjM M, int h, int w, int* __restrict__ winner,
                      int* __restrict__ rows, int* __restrict__ cols, float* __restrict__ depth) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n_points; i += (long long)gridDim.x * blockDim.x) {
    const float4 pt = __ldg(reinterpret_cast<const float4*>(points) + i);
    // numpy.linalg.norm(float32, axis=1): sqrt((x*x + y*y) + z*z) in fp32, no FMA contraction
    const float ss = __fadd_rn(__fadd_rn(__fmul_rn(pt.x, pt.x), __fmul_rn(pt.y, pt.y)), __fmul_rn(pt.z, pt.z));
    if (depth) depth[i] = __fsqrt_rn(ss);
    int row = -1, col = -1;
    if (pt.x > 0.5f) {  // parser.py:216
      const double x = pt.x, y = pt.y, z = pt.z;
      const double q0 = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(M.m[0], x), __dmul_rn(M.m[1], y)), __dmul_rn(M.m[2], z)), M.m[3]);
      const double q1 = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(M.m[4], x), __dmul_rn(M.m[5], y)), __dmul_rn(M.m[6], z)), M.m[7]);
      const double q2 = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(M.m[8], x), __dmul_rn(M.m[9], y)), __dmul_rn(M.m[10], z)), M.m[11]);
      const double u = q0 / q2, v = q1 / q2;
      if (u > 0.0 && u < (double)w && v > 0.0 && v < (double)h) {  // parser.py:222-223 (strict)
        row = (int)v;  // astype(np.int32): truncation (perspective_view_loader.py:92-93)
        col = (int)u;
        atomicMax(winner + (long long)row * w + col, (int)i);
      }
    }
    if (rows) rows[i] = row;
    if (cols) cols[i] = col;
  }
}

__global__ void __launch_bounds__(256)
project_gather_kernel(const float* __restrict__ points, const int* __restrict__ labels, const float* __restrict__ depth_pts,
                      int h, int w, const int* __restrict__ winner, float* __restrict__ feat, float* __restrict__ mask,
                      float* __restrict__ label_img) {
  const long long hw = (long long)h * w;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < hw; i += (long long)gridDim.x * blockDim.x) {
    const int k = winner[i];
    float4 pt = make_float4(0.f, 0.f, 0.f, 0.f);
    float dep = 0.f, m = 0.f, lb = 0.f;
    if (k >= 0) {
      pt = __ldg(reinterpret_cast<const float4*>(points) + k);
      dep = depth_pts ? depth_pts[k]
                      : __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(pt.x, pt.x), __fmul_rn(pt.y, pt.y)), __fmul_rn(pt.z, pt.z)));
      m = 1.f;
      lb = labels ? (float)labels[k] : 0.f;
    }
    feat[i] = dep;
    feat[hw + i] = pt.x;
    feat[2 * hw + i] = pt.y;
    feat[3 * hw + i] = pt.z;
    feat[4 * hw + i] = pt.w;
    if (mask) mask[i] = m;
    if (label_img) label_img[i] = lb;
  }
}

}  // namespace pmfb

using namespace pmfb;

#define REQ(cond, ...) \
  do {                 \
    if (!(cond)) return fail(PMFB_ERR_INVALID, __VA_ARGS__); \
  } while (0)

extern "C" int pmfb_knn_vote(const float* proj_range, const int64_t* proj_argmax, int32_t h, int32_t w,
                             const float* unproj_range, const int64_t* px, const int64_t* py, int64_t n_points,
                             const float* inv_gauss, int32_t search, int32_t knn, float cutoff, int32_t nclasses,
                             int64_t* out, void* stream) {
  REQ(search % 2 == 1, "Nearest neighbor kernel must be odd number");  // knn.py:73-74
  REQ(search >= 1 && search <= kKnnMaxSearch, "knn_vote: search=%d must be in [1,%d]", search, kKnnMaxSearch);
  REQ(knn >= 1 && knn <= kKnnMaxK && knn <= search * search, "knn_vote: knn=%d must be in [1,min(%d,search^2)]", knn, kKnnMaxK);
  REQ(proj_range && proj_argmax && inv_gauss && h > 0 && w > 0 && nclasses >= 2, "knn_vote: bad arguments");
  if (n_points == 0) return PMFB_OK;
  REQ(unproj_range && px && py && out && n_points > 0, "knn_vote: bad point arrays");
  if (knn <= kKnnRegK) {  // thread-per-point kernel (top-k in registers); the warp kernel keeps k > 8
    long long tb = (n_points + 127) / 128;
    if (tb > sm_count() * 16) tb = sm_count() * 16;
    knn_vote_points_kernel<kKnnRegK><<<(int)tb, 128, 0, (cudaStream_t)stream>>>(
        proj_range, reinterpret_cast<const long long*>(proj_argmax), 1, h, w, unproj_range, reinterpret_cast<const long long*>(px),
        reinterpret_cast<const long long*>(py), nullptr, n_points, inv_gauss, search, knn, cutoff, nclasses,
        reinterpret_cast<long long*>(out));
    PMFB_LAUNCH_CHECK("knn_vote_points_kernel");
    return PMFB_OK;
  }
  long long blocks = (n_points + 7) / 8;  // 8 warps per CTA, one point per warp per iteration
  if (blocks > sm_count() * 32) blocks = sm_count() * 32;
  knn_vote_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(
      proj_range, reinterpret_cast<const long long*>(proj_argmax), h, w, unproj_range, reinterpret_cast<const long long*>(px),
      reinterpret_cast<const long long*>(py), n_points, inv_gauss, search, knn, cutoff, nclasses,
      reinterpret_cast<long long*>(out));
  PMFB_LAUNCH_CHECK("knn_vote_kernel");
  return PMFB_OK;
}

extern "C" int pmfb_knn_vote_batched(const float* proj_range, const int64_t* proj_argmax, int32_t n_frames, int32_t h, int32_t w,
                                     const float* unproj_range, const int64_t* px, const int64_t* py, const int64_t* point_offsets,
                                     int64_t n_points, const float* inv_gauss, int32_t search, int32_t knn, float cutoff,
                                     int32_t nclasses, int64_t* out, void* stream) {
  REQ(search % 2 == 1, "Nearest neighbor kernel must be odd number");  // knn.py:73-74
  REQ(search >= 1 && search <= kKnnMaxSearch, "knn_vote_batched: search=%d must be in [1,%d]", search, kKnnMaxSearch);
  REQ(knn >= 1 && knn <= kKnnRegK && knn <= search * search, "knn_vote_batched: knn=%d must be in [1,min(%d,search^2)]", knn, kKnnRegK);
  REQ(proj_range && proj_argmax && inv_gauss && n_frames >= 1 && h > 0 && w > 0 && nclasses >= 2, "knn_vote_batched: bad arguments");
  REQ(n_frames == 1 || point_offsets, "knn_vote_batched: point_offsets (n_frames + 1 entries, device) required");
  if (n_points == 0) return PMFB_OK;
  REQ(unproj_range && px && py && out && n_points > 0, "knn_vote_batched: bad point arrays");
  long long tb = (n_points + 127) / 128;
  if (tb > sm_count() * 16) tb = sm_count() * 16;
  knn_vote_points_kernel<kKnnRegK><<<(int)tb, 128, 0, (cudaStream_t)stream>>>(
      proj_range, reinterpret_cast<const long long*>(proj_argmax), n_frames, h, w, unproj_range, reinterpret_cast<const long long*>(px),
      reinterpret_cast<const long long*>(py), reinterpret_cast<const long long*>(point_offsets), n_points, inv_gauss, search, knn,
      cutoff, nclasses, reinterpret_cast<long long*>(out));
  PMFB_LAUNCH_CHECK("knn_vote_points_kernel");
  return PMFB_OK;
}

extern "C" int pmfb_argmax_nchw(const float* probs, int32_t n, int32_t c, int32_t h, int32_t w, int32_t y0, int32_t x0, int32_t out_h,
                                int32_t out_w, int64_t* label, float* conf, void* stream) {
  REQ(probs && label && n > 0 && c > 0 && h > 0 && w > 0, "argmax_nchw: bad arguments");
  REQ(y0 >= 0 && x0 >= 0 && out_h > 0 && out_w > 0 && y0 + out_h <= h && x0 + out_w <= w, "argmax_nchw: crop window outside the map");
  const long long total = (long long)n * out_h * out_w;
  long long b = (total + 255) / 256;
  if (b > sm_count() * 16) b = sm_count() * 16;
  argmax_nchw_kernel<<<(int)b, 256, 0, (cudaStream_t)stream>>>(probs, n, c, h, w, y0, x0, out_h, out_w,
                                                              reinterpret_cast<long long*>(label), conf);
  PMFB_LAUNCH_CHECK("argmax_nchw_kernel");
  return PMFB_OK;
}

extern "C" int pmfb_lut_remap(const int64_t* labels, int64_t n, const int32_t* lut, int32_t lut_size, int32_t* out, void* stream) {
  REQ(n >= 0 && lut && lut_size > 0, "lut_remap: bad arguments");
  if (n == 0) return PMFB_OK;
  REQ(labels && out, "lut_remap: null arrays");
  long long b = (n + 255) / 256;
  if (b > sm_count() * 8) b = sm_count() * 8;
  lut_remap_kernel<<<(int)b, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const long long*>(labels), n, lut, lut_size, out);
  PMFB_LAUNCH_CHECK("lut_remap_kernel");
  return PMFB_OK;
}

extern "C" int pmfb_merge_cameras(const int64_t* point_idx, const float* conf, const int64_t* argmax, const int32_t* cam,
                                  int64_t n_entries, int64_t pc_size, uint64_t* scratch, int64_t* merged, void* stream) {
  REQ(pc_size >= 0 && n_entries >= 0 && n_entries < (1ll << 28), "merge_cameras: bad sizes (at most 2^28 entries)");
  if (pc_size == 0) return PMFB_OK;
  REQ(scratch && merged, "merge_cameras: null outputs");
  cudaStream_t st = (cudaStream_t)stream;
  PMFB_CUDA_CHECK(cudaMemsetAsync(scratch, 0, (size_t)pc_size * 8, st));
  if (n_entries > 0) {
    REQ(point_idx && conf && argmax && cam, "merge_cameras: null inputs");
    long long b = (n_entries + 255) / 256;
    if (b > sm_count() * 8) b = sm_count() * 8;
    merge_cameras_pass1<<<(int)b, 256, 0, st>>>(reinterpret_cast<const long long*>(point_idx), conf, cam, n_entries, pc_size,
                                               reinterpret_cast<unsigned long long*>(scratch));
    PMFB_LAUNCH_CHECK("merge_cameras_pass1");
  }
  long long b2 = (pc_size + 255) / 256;
  if (b2 > sm_count() * 8) b2 = sm_count() * 8;
  merge_cameras_pass2<<<(int)b2, 256, 0, st>>>(reinterpret_cast<const unsigned long long*>(scratch),
                                              reinterpret_cast<const long long*>(argmax), pc_size,
                                              reinterpret_cast<long long*>(merged));
  PMFB_LAUNCH_CHECK("merge_cameras_pass2");
  return PMFB_OK;
}

extern "C" int pmfb_confusion_add(const int64_t* pred, const int64_t* target, int64_t n, int32_t nclasses, int64_t* conf,
                                  void* stream) {
  REQ(nclasses >= 1 && nclasses <= 64 && conf && n >= 0, "confusion_add: bad arguments (nclasses <= 64)");
  if (n == 0) return PMFB_OK;
  REQ(pred && target, "confusion_add: null inputs");
  long long b = (n + 255) / 256;
  if (b > sm_count() * 4) b = sm_count() * 4;
  confusion_add_kernel<<<(int)b, 256, (size_t)nclasses * nclasses * 4, (cudaStream_t)stream>>>(
      reinterpret_cast<const long long*>(pred), reinterpret_cast<const long long*>(target), n, nclasses,
      reinterpret_cast<unsigned long long*>(conf));
  PMFB_LAUNCH_CHECK("confusion_add_kernel");
  return PMFB_OK;
}

extern "C" int pmfb_project_scatter(const float* points, const int32_t* labels, int64_t n_points, const double* proj_matrix,
                                    int32_t h, int32_t w, int32_t* winner, float* feat, float* mask, float* label_img,
                                    int32_t* rows, int32_t* cols, float* depth, void* stream) {
  REQ(proj_matrix && winner && feat && h > 0 && w > 0 && n_points >= 0, "project_scatter: bad arguments");
  REQ(n_points == 0 || (points && (reinterpret_cast<uintptr_t>(points) & 15) == 0), "project_scatter: points must be 16-byte aligned (N,4) fp32");
  REQ(n_points < (1ll << 31), "project_scatter: too many points");
  ProjM M;
  for (int i = 0; i < 12; ++i) M.m[i] = proj_matrix[i];  // host pointer: 3x4 row-major float64 (P2 @ Tr)
  cudaStream_t st = (cudaStream_t)stream;
  const long long hw = (long long)h * w;
  int blocks = (int)((hw + 255) / 256 < sm_count() * 8 ? (hw + 255) / 256 : sm_count() * 8);
  fill_i32_kernel<<<blocks, 256, 0, st>>>(winner, hw, -1);
  PMFB_LAUNCH_CHECK("fill_i32_kernel");
  if (n_points > 0) {
    int pb = (int)((n_points + 255) / 256 < sm_count() * 8 ? (n_points + 255) / 256 : sm_count() * 8);
    project_points_kernel<<<pb, 256, 0, st>>>(points, n_points, M, h, w, winner, rows, cols, depth);
    PMFB_LAUNCH_CHECK("project_points_kernel");
  }
  project_gather_kernel<<<blocks, 256, 0, st>>>(points, labels, depth, h, w, winner, feat, mask, label_img);
  PMFB_LAUNCH_CHECK("project_gather_kernel");
  return PMFB_OK;
}
