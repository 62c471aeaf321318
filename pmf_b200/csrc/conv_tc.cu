// Implicit-GEMM convolution kernels on the Blackwell tensor cores (tcgen05, kind::tf32).
//
//   conv_fwd_tc  : out[pixel][co] = epi( sum_{tap,ci} x[pixel+tap][ci] * w[tap][co][ci] )
//                  M = 128 output pixels (a tile_w x tile_h patch of one image), N = n_tile output
//                  channels, K = taps x C_in walked in 32-channel (128-byte) slabs.
//                  A-operand slabs are fetched by ONE tiled TMA box {32ch, tile_w, 1, tile_h, 1}
//                  per (tap, slab): the box lands in shared memory as 128 rows x 128 B with the
//                  128B swizzle, i.e. exactly the K-major SWIZZLE_128B UMMA operand; zero padding,
//                  dilation and image borders come from TMA out-of-bounds zero fill.
//                  Used for forward and (with transposed weights and negated taps) for dgrad.
//   conv_wgrad_tc: dw[tap][ci][co] += sum_pixels x[pixel+tap][ci] * dy[pixel][co]
//                  both operands MN-major (channels contiguous, pixels = GEMM K), split-K over
//                  pixel ranges with vector fp32 reductions into the packed gradient.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer,
// warps 2..5 = epilogue (TMEM lane quadrant = warp_idx % 4).
//
// Reference semantics being replaced: every nn.Conv2d call of pc_processor/models/pmf_net.py and
// salsanext.py (cuDNN on the reference side); see SURVEY.md Appendix C for the layer list.
#include <stdlib.h>

#include "common.h"
#include "epilogue.cuh"
#include "ptx.cuh"

namespace pmfb {

constexpr int kThreads = 192;
constexpr int kTileM = 128;
constexpr int kSlabBytes = kTileM * 128;  // 128 rows x 32 fp32
constexpr int kCtrlBytes = 1024;
constexpr int kMaxStages = 8;

struct Ctrl {
  uint64_t full[kMaxStages];
  uint64_t empty[kMaxStages];
  uint64_t tmem_full;
  uint32_t tmem_base;
};

struct ConvFwdK {
  int n_taps, kc_per_tap;
  int tap_dc[PMFB_MAX_TAPS], tap_dw[PMFB_MAX_TAPS], tap_dp[PMFB_MAX_TAPS], tap_dh[PMFB_MAX_TAPS], tap_wi[PMFB_MAX_TAPS];
  int tile_w, tile_h, tiles_x, tiles_y;
  int out_h, out_w, c_out, n_tile;
  int stages, tmem_cols;
  float* out;
  long long o_sn, o_sy, o_sx;
  EpiParams epi;
};

__global__ void __launch_bounds__(kThreads)
conv_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmx, const __grid_constant__ CUtensorMap tmw,
                   const __grid_constant__ ConvFwdK P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  Ctrl* ctrl = reinterpret_cast<Ctrl*>(smem);
  uint8_t* tiles = smem + kCtrlBytes;
  const int stage_b_bytes = P.n_tile * 128;
  const int stage_bytes = kSlabBytes + stage_b_bytes;

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);  // provably warp-uniform
  const int lane = threadIdx.x & 31;

  int bx = blockIdx.x;
  const int tx = bx % P.tiles_x;
  bx /= P.tiles_x;
  const int ty = bx % P.tiles_y;
  const int n_img = bx / P.tiles_y;
  const int x0 = tx * P.tile_w, y0 = ty * P.tile_h;
  const int n0 = blockIdx.y * P.n_tile;
  const int iters = P.n_taps * P.kc_per_tap;

  if (threadIdx.x == 0) {
    for (int s = 0; s < P.stages; ++s) {
      mbar_init(&ctrl->full[s], 1);
      mbar_init(&ctrl->empty[s], 1);
    }
    mbar_init(&ctrl->tmem_full, 1);
    fence_mbar_init();
    fence_proxy_async();
    tma_prefetch_desc(&tmx);
    tma_prefetch_desc(&tmw);
  }
  if (warp == 1) {
    tmem_alloc(&ctrl->tmem_base, (uint32_t)P.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = ctrl->tmem_base;

  if (warp == 0) {
    if (lane == 0) {
      for (int it = 0; it < iters; ++it) {
        const int tap = it / P.kc_per_tap;
        const int kc = it - tap * P.kc_per_tap;
        const int s = it % P.stages;
        const uint32_t ph = (uint32_t)(it / P.stages) & 1u;
        mbar_wait(&ctrl->empty[s], ph ^ 1u);
        uint8_t* a_s = tiles + (size_t)s * stage_bytes;
        uint8_t* b_s = a_s + kSlabBytes;
        mbar_expect_tx(&ctrl->full[s], (uint32_t)stage_bytes);
        tma_load_5d(a_s, &tmx, &ctrl->full[s], P.tap_dc[tap] + kc * 32, x0 + P.tap_dw[tap],
                    P.tap_dp[tap], y0 + P.tap_dh[tap], n_img);
        tma_load_3d(b_s, &tmw, &ctrl->full[s], kc * 32, n0, P.tap_wi[tap]);
      }
    }
  } else if (warp == 1) {
    // whole warp walks the loop with uniform descriptors; one elected lane issues (see conv_halo.cu)
    const uint32_t idesc = make_idesc_tf32(kTileM, (uint32_t)P.n_tile, 0, 0);
    const uint32_t hi = ((1024u >> 4) & 0x3FFFu) | (1u << 14) | (2u << 29);
    const uint32_t lbo_lo = (16u >> 4) << 16;
    uint32_t ph = 0;
    for (int it = 0, s = 0; it < iters; ++it) {
      mbar_wait(&ctrl->full[s], ph);
      tc_fence_after();
      const uint32_t a_addr = smem_u32(tiles + (size_t)s * stage_bytes);
      const uint32_t a_lo = ((a_addr & 0x3FFFFu) >> 4) | lbo_lo;
      const uint32_t b_lo = (((a_addr + kSlabBytes) & 0x3FFFFu) >> 4) | lbo_lo;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        umma_issue<false>(tmem_base, a_lo + 2u * k, hi, b_lo + 2u * k, hi, idesc, (uint32_t)(it | k));
      }
      umma_commit_warp(&ctrl->empty[s]);
      if (++s == P.stages) {
        s = 0;
        ph ^= 1u;
      }
    }
    umma_commit_warp(&ctrl->tmem_full);
  } else {
    // ---------------- epilogue: TMEM -> registers -> fused pointwise -> global (NHWC)
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int th = row / P.tile_w;
    const int tw = row - th * P.tile_w;
    const int y = y0 + th, x = x0 + tw;
    const bool valid = (y < P.out_h) && (x < P.out_w);
    mbar_wait(&ctrl->tmem_full, 0);
    tc_fence_after();
    const long long opix = (long long)n_img * P.o_sn + (long long)y * P.o_sy + (long long)x * P.o_sx;
    EpiPixel ep = epi_pixel(P.epi, n_img, y, x);
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
    for (int cb = 0; cb < P.n_tile; cb += 16) {
      float v[16];
      tmem_ld16(taddr + (uint32_t)cb, v);
      tmem_ld_wait();
      if (valid) {
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          const int c = n0 + cb + j;
          if (c < P.c_out) {
            float4 o = epi_apply4(P.epi, ep, c, make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]));
            *reinterpret_cast<float4*>(P.out + opix + c) = o;
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, (uint32_t)P.tmem_cols);
}

// =====================================================================================
// wgrad
// =====================================================================================
struct WgradK {
  int n_taps, cb_per_tap;
  int tap_dc[PMFB_MAX_TAPS], tap_dw[PMFB_MAX_TAPS], tap_dp[PMFB_MAX_TAPS], tap_dh[PMFB_MAX_TAPS];
  int ptile_w, ptile_h, ptiles_x, ptiles_y, n_batch;
  int c_in, c_out, n_tile, ksplit;
  int stages, tmem_cols;
  float* dw;
};

constexpr int kBoxBytes = 32 * 128;  // 32 pixels x 32 channels fp32

__global__ void __launch_bounds__(kThreads)
conv_wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmx, const __grid_constant__ CUtensorMap tmdy,
                     const __grid_constant__ WgradK P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  Ctrl* ctrl = reinterpret_cast<Ctrl*>(smem);
  uint8_t* tiles = smem + kCtrlBytes;
  const int nb_boxes = P.n_tile / 32;
  const int stage_bytes = (4 + nb_boxes) * kBoxBytes;

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);  // provably warp-uniform
  const int lane = threadIdx.x & 31;

  const int total_rb = P.n_taps * P.cb_per_tap;
  const int rb0 = blockIdx.x * 4;
  const int n_rb = min(4, total_rb - rb0);
  const int n0 = blockIdx.y * P.n_tile;
  const int total_pt = P.ptiles_x * P.ptiles_y * P.n_batch;
  const int pt_begin = (int)(((long long)total_pt * blockIdx.z) / P.ksplit);
  const int pt_end = (int)(((long long)total_pt * (blockIdx.z + 1)) / P.ksplit);
  const int iters = pt_end - pt_begin;

  if (threadIdx.x == 0) {
    for (int s = 0; s < P.stages; ++s) {
      mbar_init(&ctrl->full[s], 1);
      mbar_init(&ctrl->empty[s], 1);
    }
    mbar_init(&ctrl->tmem_full, 1);
    fence_mbar_init();
    fence_proxy_async();
    tma_prefetch_desc(&tmx);
    tma_prefetch_desc(&tmdy);
  }
  if (warp == 1) {
    tmem_alloc(&ctrl->tmem_base, (uint32_t)P.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = ctrl->tmem_base;

  if (iters > 0) {
    if (warp == 0) {
      if (lane == 0) {
        for (int it = 0; it < iters; ++it) {
          int pt = pt_begin + it;
          const int px = pt % P.ptiles_x;
          pt /= P.ptiles_x;
          const int py = pt % P.ptiles_y;
          const int n_img = pt / P.ptiles_y;
          const int x0 = px * P.ptile_w, y0 = py * P.ptile_h;
          const int s = it % P.stages;
          const uint32_t ph = (uint32_t)(it / P.stages) & 1u;
          mbar_wait(&ctrl->empty[s], ph ^ 1u);
          uint8_t* a_s = tiles + (size_t)s * stage_bytes;
          uint8_t* b_s = a_s + 4 * kBoxBytes;
          mbar_expect_tx(&ctrl->full[s], (uint32_t)((n_rb + nb_boxes) * kBoxBytes));
          for (int r = 0; r < n_rb; ++r) {
            const int rb = rb0 + r;
            const int tap = rb / P.cb_per_tap;
            const int cb = rb - tap * P.cb_per_tap;
            tma_load_5d(a_s + r * kBoxBytes, &tmx, &ctrl->full[s], P.tap_dc[tap] + cb * 32,
                        x0 + P.tap_dw[tap], P.tap_dp[tap], y0 + P.tap_dh[tap], n_img);
          }
          for (int j = 0; j < nb_boxes; ++j)
            tma_load_5d(b_s + j * kBoxBytes, &tmdy, &ctrl->full[s], n0 + j * 32, x0, 0, y0, n_img);
        }
      }
    } else if (warp == 1) {
      const uint32_t idesc = make_idesc_tf32(kTileM, (uint32_t)P.n_tile, 1, 1);
      // MN-major tf32 operands must use the SWIZZLE_128B_BASE32B layout (layout type 1): 32 channels (128 B)
      // contiguous per pixel row, 32B chunks XOR-swizzled over 4-row (512 B) atoms.  One K=8 MMA step spans two
      // atoms (SBO = 512 B); LBO = distance between 32-channel blocks; a K step advances the start by 1024 B.
      const uint32_t hi = ((512u >> 4) & 0x3FFFu) | (1u << 14) | (1u << 29);
      const uint32_t lbo_lo = (((uint32_t)kBoxBytes >> 4) & 0x3FFFu) << 16;
      uint32_t ph = 0;
      for (int it = 0, s = 0; it < iters; ++it) {
        mbar_wait(&ctrl->full[s], ph);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(tiles + (size_t)s * stage_bytes);
        const uint32_t a_lo = ((a_addr & 0x3FFFFu) >> 4) | lbo_lo;
        const uint32_t b_lo = (((a_addr + 4 * kBoxBytes) & 0x3FFFFu) >> 4) | lbo_lo;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          umma_issue<false>(tmem_base, a_lo + 64u * k, hi, b_lo + 64u * k, hi, idesc, (uint32_t)(it | k));
        }
        umma_commit_warp(&ctrl->empty[s]);
        if (++s == P.stages) {
          s = 0;
          ph ^= 1u;
        }
      }
      umma_commit_warp(&ctrl->tmem_full);
    } else {
      const int q = warp & 3;  // accumulator rows q*32..q*32+31 == row-block q of this CTA
      const int rb = rb0 + q;
      const int tap = rb / P.cb_per_tap;
      const int cb = rb - tap * P.cb_per_tap;
      const int ci = cb * 32 + lane;
      const bool valid = (q < n_rb) && (ci < P.c_in);
      mbar_wait(&ctrl->tmem_full, 0);
      tc_fence_after();
      float* dst = P.dw + ((long long)tap * P.c_in + ci) * P.c_out;
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
      for (int c0 = 0; c0 < P.n_tile; c0 += 16) {
        float v[16];
        tmem_ld16(taddr + (uint32_t)c0, v);
        tmem_ld_wait();
        if (valid) {
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            const int co = n0 + c0 + j;
            if (co < P.c_out)
              atomicAdd(reinterpret_cast<float4*>(dst + co), make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]));
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, (uint32_t)P.tmem_cols);
}

static int pow2_cols(int n) {
  int c = 32;
  while (c < n) c <<= 1;
  return c;
}

static const int kSmemBudget = 200 * 1024;

}  // namespace pmfb

namespace pmfb {
int halo_eligible(const pmfb_conv_desc* d);
int launch_conv_halo(const pmfb_conv_desc* d, void* stream);
int halo_fused_stats_ok(const pmfb_conv_desc* d);
int wgrad_halo_eligible(const pmfb_wgrad_desc* d);
int launch_wgrad_halo(const pmfb_wgrad_desc* d, void* stream);
}  // namespace pmfb

using namespace pmfb;

extern "C" int pmfb_conv_fused_stats_ok(const pmfb_conv_desc* d) {
  if (!d || d->n_taps < 1 || d->n_taps > PMFB_MAX_TAPS) return 0;
  const char* e = getenv("PMFB_CONV_V1");
  if (e && atoi(e)) return 0;
  return (halo_eligible(d) && halo_fused_stats_ok(d)) ? 1 : 0;
}

extern "C" int pmfb_conv16_ok(const pmfb_conv_desc* d) {
  if (!d || d->n_taps < 1 || d->n_taps > PMFB_MAX_TAPS || d->c_in % 8) return 0;
  const char* e = getenv("PMFB_CONV_V1");
  if (e && atoi(e)) return 0;
  return halo_eligible(d) ? 1 : 0;
}

extern "C" int pmfb_conv_fwd(const pmfb_conv_desc* d, void* stream) {
  if (!d) return fail(PMFB_ERR_INVALID, "null desc");
  if (d->n_taps < 1 || d->n_taps > PMFB_MAX_TAPS) return fail(PMFB_ERR_INVALID, "n_taps=%d", d->n_taps);
  if (d->c_out % 4 || d->c_in % 4) return fail(PMFB_ERR_INVALID, "c_in/c_out must be multiples of 4");
  if ((d->o_sn | d->o_sy | d->o_sx) % 4 || (reinterpret_cast<uintptr_t>(d->out) & 15))
    return fail(PMFB_ERR_INVALID, "output view must be 16-byte aligned with strides multiple of 4");
  {
    // stride-1 layers with a small tap halo run on the persistent halo-tile kernel (conv_halo.cu); the
    // tap-per-TMA kernel below keeps the stride-2 (parity-layout) and wide-dilation (ASPP) layers.
    static int force_v1 = -1;
    if (force_v1 < 0) {
      const char* e = getenv("PMFB_CONV_V1");
      force_v1 = (e && atoi(e)) ? 1 : 0;
    }
    if (!force_v1 && halo_eligible(d)) return launch_conv_halo(d, stream);
  }
  if (d->dtype != PMFB_DT_F32) return fail(PMFB_ERR_INVALID, "conv_fwd: 16-bit operands are only supported on the stride-1 halo kernel (query pmfb_conv16_ok)");
  if (d->bn_stats) return fail(PMFB_ERR_INVALID, "conv_fwd: fused BN statistics are not available for this layer (query pmfb_conv_fused_stats_ok)");
  if (d->out_half) return fail(PMFB_ERR_INVALID, "conv_fwd: fp16 output exists only with the fused-statistics epilogue of the halo kernel");
  if (d->tile_w * d->tile_h != kTileM) return fail(PMFB_ERR_INVALID, "tile_w*tile_h must be 128");
  if (d->n_tile < 16 || d->n_tile > 256 || d->n_tile % 16)
    return fail(PMFB_ERR_INVALID, "n_tile=%d must be a multiple of 16 in [16,256]", d->n_tile);
  if (d->n_taps < 1 || d->n_taps > PMFB_MAX_TAPS) return fail(PMFB_ERR_INVALID, "n_taps=%d", d->n_taps);
  if (d->c_out % 4 || d->c_in % 4) return fail(PMFB_ERR_INVALID, "c_in/c_out must be multiples of 4");
  if ((d->o_sn | d->o_sy | d->o_sx) % 4 || (reinterpret_cast<uintptr_t>(d->out) & 15))
    return fail(PMFB_ERR_INVALID, "output view must be 16-byte aligned with strides multiple of 4");

  CUtensorMap tmx, tmw;
  uint32_t boxx[5] = {32, (uint32_t)d->tile_w, 1, (uint32_t)d->tile_h, 1};
  int rc = make_tmap_f32(&tmx, d->x.ptr, 5, d->x.dims, d->x.strides, boxx);
  if (rc) return rc;
  int n_slabs = d->n_taps;
  if (d->use_tap_wi)
    for (int i = 0; i < d->n_taps; ++i) {
      if (d->tap_wi[i] < 0) return fail(PMFB_ERR_INVALID, "tap_wi[%d]=%d", i, d->tap_wi[i]);
      if (d->tap_wi[i] + 1 > n_slabs) n_slabs = d->tap_wi[i] + 1;
    }
  uint64_t wdims[3] = {(uint64_t)d->c_in, (uint64_t)d->c_out, (uint64_t)n_slabs};
  uint64_t wstr[2] = {(uint64_t)d->c_in * 4, (uint64_t)d->c_in * d->c_out * 4};
  uint32_t boxw[3] = {32, (uint32_t)d->n_tile, 1};
  rc = make_tmap_f32(&tmw, d->w, 3, wdims, wstr, boxw);
  if (rc) return rc;

  ConvFwdK P;
  P.n_taps = d->n_taps;
  P.kc_per_tap = (d->c_in + 31) / 32;
  for (int i = 0; i < PMFB_MAX_TAPS; ++i) {
    P.tap_dc[i] = d->tap_dc[i];
    P.tap_dw[i] = d->tap_dw[i];
    P.tap_dp[i] = d->tap_dp[i];
    P.tap_dh[i] = d->tap_dh[i];
    P.tap_wi[i] = d->use_tap_wi ? d->tap_wi[i] : i;
  }
  P.tile_w = d->tile_w;
  P.tile_h = d->tile_h;
  P.tiles_x = (d->out_w + d->tile_w - 1) / d->tile_w;
  P.tiles_y = (d->out_h + d->tile_h - 1) / d->tile_h;
  P.out_h = d->out_h;
  P.out_w = d->out_w;
  P.c_out = d->c_out;
  P.n_tile = d->n_tile;
  const int stage_bytes = kSlabBytes + d->n_tile * 128;
  int stages = (96 * 1024 - kCtrlBytes) / stage_bytes;  // aim for 2 CTAs / SM
  if (stages < 3) stages = (kSmemBudget - kCtrlBytes) / stage_bytes < 4 ? (kSmemBudget - kCtrlBytes) / stage_bytes : 4;
  if (stages > kMaxStages) stages = kMaxStages;
  const int iters = P.n_taps * P.kc_per_tap;
  if (stages > iters) stages = iters < 1 ? 1 : iters;
  P.stages = stages;
  P.tmem_cols = pow2_cols(d->n_tile);
  P.out = d->out;
  P.o_sn = d->o_sn;
  P.o_sy = d->o_sy;
  P.o_sx = d->o_sx;
  rc = epi_from_c(&d->epi, &P.epi);
  if (rc) return rc;

  const size_t smem = (size_t)kCtrlBytes + (size_t)stages * stage_bytes + 1024;
  static size_t smem_set = 0;
  if (smem > smem_set) {
    PMFB_CUDA_CHECK(cudaFuncSetAttribute(conv_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kSmemBudget + 4096)));
    smem_set = kSmemBudget + 4096;
  }
  dim3 grid((unsigned)(P.tiles_x * P.tiles_y * d->n_batch), (unsigned)((d->c_out + d->n_tile - 1) / d->n_tile), 1);
  conv_fwd_tc_kernel<<<grid, kThreads, smem, (cudaStream_t)stream>>>(tmx, tmw, P);
  PMFB_LAUNCH_CHECK("conv_fwd_tc_kernel");
  return PMFB_OK;
}

extern "C" int pmfb_wgrad16_ok(const pmfb_wgrad_desc* d) {
  if (!d || d->n_taps < 1 || d->n_taps > PMFB_MAX_TAPS || d->c_in % 8 || d->c_out % 8) return 0;
  const char* e = getenv("PMFB_WGRAD_V1");
  if (e && atoi(e)) return 0;
  return wgrad_halo_eligible(d) ? 1 : 0;
}

extern "C" int pmfb_conv_wgrad(const pmfb_wgrad_desc* d, void* stream) {
  if (!d) return fail(PMFB_ERR_INVALID, "null desc");
  if (d->n_taps < 1 || d->n_taps > PMFB_MAX_TAPS) return fail(PMFB_ERR_INVALID, "n_taps=%d", d->n_taps);
  if (d->c_out % 4 || d->c_in % 4) return fail(PMFB_ERR_INVALID, "c_in/c_out must be multiples of 4");
  {
    static int force_v1 = -1;
    if (force_v1 < 0) {
      const char* e = getenv("PMFB_WGRAD_V1");
      force_v1 = (e && atoi(e)) ? 1 : 0;
    }
    if (!force_v1 && wgrad_halo_eligible(d)) return launch_wgrad_halo(d, stream);
  }
  if (d->dtype != PMFB_DT_F32) return fail(PMFB_ERR_INVALID, "conv_wgrad: bf16 operands are only supported on the stride-1 halo kernel (query pmfb_wgrad16_ok)");
  if (d->ptile_w * d->ptile_h != 32) return fail(PMFB_ERR_INVALID, "ptile_w*ptile_h must be 32");
  if (d->n_tile < 32 || d->n_tile > 256 || d->n_tile % 32)
    return fail(PMFB_ERR_INVALID, "n_tile=%d must be a multiple of 32 in [32,256]", d->n_tile);
  if (d->n_taps < 1 || d->n_taps > PMFB_MAX_TAPS) return fail(PMFB_ERR_INVALID, "n_taps=%d", d->n_taps);
  if (d->c_out % 4 || d->c_in % 4) return fail(PMFB_ERR_INVALID, "c_in/c_out must be multiples of 4");
  if (d->ksplit < 1) return fail(PMFB_ERR_INVALID, "ksplit");

  CUtensorMap tmx, tmdy;
  uint32_t box[5] = {32, (uint32_t)d->ptile_w, 1, (uint32_t)d->ptile_h, 1};
  int rc = make_tmap_f32(&tmx, d->x.ptr, 5, d->x.dims, d->x.strides, box, true);
  if (rc) return rc;
  rc = make_tmap_f32(&tmdy, d->dy.ptr, 5, d->dy.dims, d->dy.strides, box, true);
  if (rc) return rc;

  WgradK P;
  P.n_taps = d->n_taps;
  P.cb_per_tap = (d->c_in + 31) / 32;
  for (int i = 0; i < PMFB_MAX_TAPS; ++i) {
    P.tap_dc[i] = d->tap_dc[i];
    P.tap_dw[i] = d->tap_dw[i];
    P.tap_dp[i] = d->tap_dp[i];
    P.tap_dh[i] = d->tap_dh[i];
  }
  P.ptile_w = d->ptile_w;
  P.ptile_h = d->ptile_h;
  P.ptiles_x = (d->out_w + d->ptile_w - 1) / d->ptile_w;
  P.ptiles_y = (d->out_h + d->ptile_h - 1) / d->ptile_h;
  P.n_batch = d->n_batch;
  P.c_in = d->c_in;
  P.c_out = d->c_out;
  P.n_tile = d->n_tile;
  const int total_pt = P.ptiles_x * P.ptiles_y * P.n_batch;
  P.ksplit = d->ksplit > total_pt ? total_pt : d->ksplit;
  const int stage_bytes = (4 + d->n_tile / 32) * kBoxBytes;
  int stages = (96 * 1024 - kCtrlBytes) / stage_bytes;
  if (stages < 3) stages = (kSmemBudget - kCtrlBytes) / stage_bytes < 4 ? (kSmemBudget - kCtrlBytes) / stage_bytes : 4;
  if (stages > kMaxStages) stages = kMaxStages;
  P.stages = stages;
  P.tmem_cols = pow2_cols(d->n_tile);
  P.dw = d->dw;

  const size_t smem = (size_t)kCtrlBytes + (size_t)stages * stage_bytes + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    PMFB_CUDA_CHECK(cudaFuncSetAttribute(conv_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kSmemBudget + 4096)));
    attr_set = true;
  }
  const int total_rb = P.n_taps * P.cb_per_tap;
  dim3 grid((unsigned)((total_rb + 3) / 4), (unsigned)((d->c_out + d->n_tile - 1) / d->n_tile), (unsigned)P.ksplit);
  conv_wgrad_tc_kernel<<<grid, kThreads, smem, (cudaStream_t)stream>>>(tmx, tmdy, P);
  PMFB_LAUNCH_CHECK("conv_wgrad_tc_kernel");
  return PMFB_OK;
}
