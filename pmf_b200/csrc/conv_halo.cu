// conv_fwd_halo_kernel — second-generation implicit-GEMM convolution for the stride-1 layers (forward and dgrad).
//
// Why: in the tap-per-TMA kernel (conv_tc.cu) every tap re-fetches the whole 128-pixel input box from L2, so a
// 3x3 layer moves 9x its input over the L2->SM fabric; ncu showed the 32/64-channel full-resolution layers
// latency/L2 bound (tensor pipe 2-7 %, DRAM 9-25 %).  Here a CTA loads, per 32-channel slab, ONE halo tile
//   (16*MT + 2*hy) x (8 + 2*hx) pixels x 32 ch   (one TMA box, 128B-swizzled, zero-filled outside the image)
// and forms the A operand of every tap by POINTING the UMMA shared-memory descriptor at a shifted window of that
// tile: with an 8-pixel-wide output tile every 8-row swizzle atom is one output row, so consecutive atoms are a
// constant (8 + 2*hx)*128 bytes apart (the descriptor's stride-byte-offset) for every tap.  Only the weights stream
// per tap.  The kernel is persistent (one CTA per SM walks the tile list), keeps up to two accumulator sets in TMEM so
// the epilogue of tile i overlaps the MMAs of tile i+1, and processes MT stacked 128-pixel tiles per weight fetch.
//
// Warp roles (320 threads): warp 0 TMA producer, warp 1 TMEM allocator + MMA issuer, warps 2..9 epilogue (two per TMEM
// lane quadrant).
#include <stdlib.h>

#include <cuda_fp16.h>

#include "common.h"
#include "epilogue.cuh"
#include "ptx.cuh"

namespace pmfb {

constexpr int kHThreads = 320;  // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue (two per TMEM lane quadrant)
constexpr int kHEpiWarps = 8;
constexpr int kHMaxC = 512;       // per-channel epilogue vectors staged in shared memory
constexpr int kHCtrlBytes = 1024 + 4 * kHMaxC * 4;
constexpr int kHMaxB = 8;
constexpr int kHSmemBudget = 208 * 1024;
constexpr int kHStageBytes = 4096;  // per epilogue warp: 32 pixels x 32 channels, the source box of one TMA store

struct HCtrl {
  uint64_t full_a[2], empty_a[2];
  uint64_t full_b[kHMaxB], empty_b[kHMaxB];
  uint64_t tmem_full[2], tmem_empty[2];
  uint32_t tmem_base;
};

struct HaloK {
  int n_taps, ks;
  int tap_dw[PMFB_MAX_TAPS], tap_dh[PMFB_MAX_TAPS], tap_wi[PMFB_MAX_TAPS];
  int tap_off[PMFB_MAX_TAPS];  // (descriptor units of 16 B) offset of tap t's window inside the halo tile
  int hx, hy, mt;
  int tiles_x, tiles_y, n_batch, n_blocks;
  int out_h, out_w, c_out, n_tile;
  int nsb, a_bytes, a_box_bytes, nacc, tmem_cols, use_base_off, tma_store;
  int dtype, kslab;  // operand type (PMFB_DT_*) and channels per 128-byte slab (32 fp32 / 64 16-bit)
  int klast;         // 32-byte K steps of the LAST slab (a 32-channel layer fills half a 16-bit slab: 2 steps, not 4)
  int row_bytes;     // operand row in shared memory: 128 (SWIZZLE_128B) or, for 16-bit layers of <= 32 channels, 64 (SWIZZLE_64B)
  float* out;
  long long o_sn, o_sy, o_sx;
  EpiParams epi;
  double* bn_stats;
};

__device__ __forceinline__ uint64_t desc_with_base(uint32_t saddr, uint32_t lbo, uint32_t sbo, int use_base_off) {
  uint64_t d = make_smem_desc(saddr, lbo, sbo, 2u);
  if (use_base_off) d |= static_cast<uint64_t>((saddr >> 7) & 7u) << 49;
  return d;
}

// EPI selects the epilogue at compile time.  kEpiGeneric evaluates the whole pmfb_epilogue functional from runtime
// flags (eval-mode fusions: BN affine, residual, gate).  The fast variants cover what a TRAINING step launches --
// [+bias] [LeakyReLU] [tf32 round] and [+= r1 (gradient accumulation)] [tf32 round] -- with everything else compiled
// out: ncu showed the generic epilogue executing ~1350 warp instructions per 32x32 unit (8 warps per SM, issue and
// dependency bound at 1.4 TB/s of stores) and every thin full-resolution layer pinned to that rate.
constexpr int kEpiGeneric = -1;
constexpr int kEpiB1 = 1, kEpiR1 = 2, kEpiRnd = 4, kEpiLeaky = 8;
// kEpiStats: BatchNorm batch statistics of the epilogue result, fused.  After a warp has parked its 32-pixel x 32-channel
// unit in the staging tile, lane c re-reads COLUMN c (32 conflict-free 4-byte loads: the 128B swizzle permutes the
// 16-byte chunks of a row, so the 32 lanes of one row read hit 32 different banks), adds the 32 values and their squares
// and issues two shared-memory float atomics into per-CTA channel sums; the CTA flushes them once, at its end, with
// fp64 atomics (pmfb_bn_stats contract).  Saves the separate statistics pass over every pre-BN activation.
constexpr int kEpiStats = 16;
// Eval-mode fusions (inference: BASELINE config 5): alpha1 scale (eval BN folded in front of the activation: torchvision
// conv -> BN -> ReLU), ReLU, the post-activation affine alpha2/beta2 (SalsaNext conv -> LeakyReLU -> BN) and the second
// residual r2 (ResContextBlock / ResBlock shortcut).  Only the combinations listed in launch_halo_variant exist.
constexpr int kEpiA1 = 32, kEpiRelu = 64, kEpiA2B2 = 128, kEpiR2 = 256;
// kEpiHalf (only with kEpiStats): the result is stored as fp16 -- the pre-BatchNorm activation of a training-mode
// conv -> [LeakyReLU] -> BN layer, which only the BatchNorm passes read (never a convolution).  A warp's unit is then a
// 32-pixel x 32-channel tile of 64-byte rows (SWIZZLE_64B: 16-byte chunk ^= (row >> 1) & 3, conflict-free 16-byte stores),
// stored by one bulk tensor store; the fused statistics are those of the ROUNDED values (what BatchNorm will normalise).
constexpr int kEpiHalf = 512;
constexpr int kHStatsC = 256;  // fused statistics: c_out <= 256 (2 x 256 fp64 accumulators in the unused alpha2/beta2 slots)

// MMA issue loop of conv_fwd_halo_kernel (warp 1), specialised at compile time on the operand kind so that the issuing
// warp's inner loop carries no per-instruction branch (instruction issue of this warp paces the tensor pipe).
// MMA issue loop of conv_fwd_halo_kernel (warp 1), specialised at compile time on the operand kind.  Instruction issue of
// this ONE warp paces the tensor pipe on the thin layers (ncu, 32 -> 32 channels at 480x640: ~145 warp instructions per
// tap for 4 UMMAs of 80 cycles each), so the per-tap body is kept minimal: ring slot / phase / descriptor words advance
// by adds (no modulo), the tap's window offset comes from a host-computed table, everything stays warp-uniform.
template <bool F16>
__device__ __forceinline__ void mma_issue_loop(const HaloK& P, HCtrl* ctrl, uint8_t* a_buf, uint8_t* b_buf, const int b_bytes,
                                               const uint32_t tmem_base_in, const int total, const int pitch) {
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, tmem_base_in, 0);  // provably warp-uniform
    const uint32_t idesc = F16 ? make_idesc_f16(128, (uint32_t)P.n_tile, P.dtype == PMFB_DT_BF16 ? 1u : 0u,
                                                (P.dtype == PMFB_DT_BF16 || P.dtype == PMFB_DT_F16_BF16) ? 1u : 0u, 0, 0)
                               : make_idesc_tf32(128, (uint32_t)P.n_tile, 0, 0);
    const uint32_t rb = (uint32_t)P.row_bytes;
    const uint32_t lay = rb == 128u ? (2u << 29) : (4u << 29);  // layout type: SWIZZLE_128B = 2, SWIZZLE_64B = 4
    const uint32_t sbo = (uint32_t)pitch * rb;
    const uint32_t hi_a = ((sbo >> 4) & 0x3FFFu) | (1u << 14) | lay;   // bits 32..63 of the A descriptor
    const uint32_t hi_b = (((8u * rb) >> 4) & 0x3FFFu) | (1u << 14) | lay;
    const uint32_t lbo_lo = (16u >> 4) << 16;
    const uint32_t j_step = ((uint32_t)(16 * pitch) * rb) >> 4;
    const int ks = P.ks, n_taps = P.n_taps, klast = P.klast;
    const bool two = P.mt == 2;
    const uint32_t n_tile = (uint32_t)P.n_tile, nsb = (uint32_t)P.nsb, nacc = (uint32_t)P.nacc;
    const uint32_t acc_cols = (uint32_t)(P.mt * P.n_tile);
    const uint32_t a_lo_buf0 = ((smem_u32(a_buf) & 0x3FFFFu) >> 4) | lbo_lo;
    const uint32_t a_lo_buf1 = ((smem_u32(a_buf + (size_t)P.a_bytes) & 0x3FFFFu) >> 4) | lbo_lo;
    const uint32_t b_lo_first = ((smem_u32(b_buf) & 0x3FFFFu) >> 4) | lbo_lo;
    const uint32_t b_step = (uint32_t)b_bytes >> 4;
    uint32_t a_it = 0;
    uint32_t st = 0, st_ph = 0, b_lo = b_lo_first;  // weight ring: slot, its phase, descriptor low word of the slot
    uint32_t buf = 0, buf_ph = 0;                   // accumulator set and its phase
    for (int w = blockIdx.x; w < total; w += gridDim.x) {
      mbar_wait(&ctrl->tmem_empty[buf], buf_ph ^ 1u);
      tc_fence_after();
      const uint32_t d_base = tmem_base + buf * acc_cols;
      uint32_t accumulate = 0;
      for (int s = 0; s < ks; ++s) {
        const uint32_t ab = a_it & 1u;
        mbar_wait(&ctrl->full_a[ab], (a_it >> 1) & 1u);
        tc_fence_after();
        const uint32_t a_lo0 = ab ? a_lo_buf1 : a_lo_buf0;
        const int nk = (s == ks - 1) ? klast : 4;
        for (int t = 0; t < n_taps; ++t) {
          mbar_wait(&ctrl->full_b[st], st_ph);
          tc_fence_after();
          const uint32_t a_lo = a_lo0 + (uint32_t)P.tap_off[t];
          if (nk == 4) {
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_issue<F16>(d_base, a_lo + 2u * k, hi_a, b_lo + 2u * k, hi_b, idesc, accumulate | (uint32_t)k);
            if (two) {
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_issue<F16>(d_base + n_tile, a_lo + j_step + 2u * k, hi_a, b_lo + 2u * k, hi_b, idesc, accumulate | (uint32_t)k);
            }
          } else if (nk == 2) {  // 32 channels in a 16-bit slab (64-byte operand rows)
            umma_issue<F16>(d_base, a_lo, hi_a, b_lo, hi_b, idesc, accumulate);
            umma_issue<F16>(d_base, a_lo + 2u, hi_a, b_lo + 2u, hi_b, idesc, 1u);
            if (two) {
              umma_issue<F16>(d_base + n_tile, a_lo + j_step, hi_a, b_lo, hi_b, idesc, accumulate);
              umma_issue<F16>(d_base + n_tile, a_lo + j_step + 2u, hi_a, b_lo + 2u, hi_b, idesc, 1u);
            }
          } else {  // other partial last slabs
            for (int k = 0; k < nk; ++k) umma_issue<F16>(d_base, a_lo + 2u * k, hi_a, b_lo + 2u * k, hi_b, idesc, accumulate | (uint32_t)k);
            if (two)
              for (int k = 0; k < nk; ++k)
                umma_issue<F16>(d_base + n_tile, a_lo + j_step + 2u * k, hi_a, b_lo + 2u * k, hi_b, idesc, accumulate | (uint32_t)k);
          }
          accumulate = 1;
          umma_commit_warp(&ctrl->empty_b[st]);
          if (++st == nsb) {
            st = 0;
            st_ph ^= 1u;
            b_lo = b_lo_first;
          } else {
            b_lo += b_step;
          }
        }
        umma_commit_warp(&ctrl->empty_a[ab]);
        ++a_it;
      }
      umma_commit_warp(&ctrl->tmem_full[buf]);
      if (++buf == nacc) {
        buf = 0;
        buf_ph ^= 1u;
      }
    }
}

template <int EPI>
__global__ void __launch_bounds__(kHThreads, 1)
conv_fwd_halo_kernel(const __grid_constant__ CUtensorMap tmx, const __grid_constant__ CUtensorMap tmw,
                     const __grid_constant__ CUtensorMap tmo, const __grid_constant__ HaloK P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  HCtrl* ctrl = reinterpret_cast<HCtrl*>(smem);
  uint8_t* a_buf = smem + kHCtrlBytes;
  uint8_t* b_buf = a_buf + 2 * (size_t)P.a_bytes;
  const int b_bytes = P.n_tile * P.row_bytes;

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);  // provably warp-uniform -> uniform datapath
  const int lane = threadIdx.x & 31;
  const int total = P.tiles_x * P.tiles_y * P.n_batch * P.n_blocks;
  const int th = 16 * P.mt;          // CTA tile height in pixels
  const int pitch = 8 + 2 * P.hx;    // halo tile row pitch in pixels

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&ctrl->full_a[i], 1);
      mbar_init(&ctrl->empty_a[i], 1);
      mbar_init(&ctrl->tmem_full[i], 1);
      mbar_init(&ctrl->tmem_empty[i], kHEpiWarps);
    }
    for (int i = 0; i < P.nsb; ++i) {
      mbar_init(&ctrl->full_b[i], 1);
      mbar_init(&ctrl->empty_b[i], 1);
    }
    fence_mbar_init();
    fence_proxy_async();
    tma_prefetch_desc(&tmx);
    tma_prefetch_desc(&tmw);
    if (P.tma_store) tma_prefetch_desc(&tmo);
  }
  if (warp == 1) {
    tmem_alloc(&ctrl->tmem_base, (uint32_t)P.tmem_cols);
    tmem_relinquish();
  }
  {
    float* sv = reinterpret_cast<float*>(smem + 1024);
    const float* src[4] = {P.epi.alpha1, P.epi.beta1, P.epi.alpha2, P.epi.beta2};
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (src[k])
        for (int c = threadIdx.x; c < P.c_out; c += kHThreads) sv[k * kHMaxC + c] = __ldg(src[k] + c);
  }
  if constexpr (EPI != kEpiGeneric && (EPI & kEpiStats) != 0) {  // per-CTA channel sums live in the unused alpha2/beta2 slots
    double* ds = reinterpret_cast<double*>(smem + 1024 + 2 * kHMaxC * 4);  // [0:256) sums, [256:512) sums of squares
    for (int c = threadIdx.x; c < 2 * kHStatsC; c += kHThreads) ds[c] = 0.0;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = ctrl->tmem_base;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t a_it = 0, st = 0, st_ph = 1;  // st_ph: parity to wait for on the slot's EMPTY barrier (first pass is free)
      const uint32_t nsb = (uint32_t)P.nsb;
      for (int w = blockIdx.x; w < total; w += gridDim.x) {
        int r = w;
        const int nb = r % P.n_blocks; r /= P.n_blocks;
        const int tx = r % P.tiles_x; r /= P.tiles_x;
        const int ty = r % P.tiles_y;
        const int n_img = r / P.tiles_y;
        const int x0 = tx * 8, y0 = ty * th, n0 = nb * P.n_tile;
        for (int s = 0; s < P.ks; ++s) {
          const uint32_t ab = a_it & 1u;
          mbar_wait(&ctrl->empty_a[ab], ((a_it >> 1) & 1u) ^ 1u);
          mbar_expect_tx(&ctrl->full_a[ab], (uint32_t)P.a_box_bytes);
          tma_load_5d(a_buf + (size_t)ab * P.a_bytes, &tmx, &ctrl->full_a[ab], s * P.kslab, x0 - P.hx, 0, y0 - P.hy, n_img);
          ++a_it;
          for (int t = 0; t < P.n_taps; ++t) {
            mbar_wait(&ctrl->empty_b[st], st_ph);
            mbar_expect_tx(&ctrl->full_b[st], (uint32_t)b_bytes);
            tma_load_3d(b_buf + (size_t)st * b_bytes, &tmw, &ctrl->full_b[st], s * P.kslab, n0, P.tap_wi[t]);
            if (++st == nsb) {
              st = 0;
              st_ph ^= 1u;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------- MMA issuer.  The WHOLE warp walks the loops so that every descriptor lives in uniform registers
    // (ncu: with a single divergent lane the ~20 integer instructions that rebuilt the two 64-bit descriptors for each
    // UTCHMMA made instruction issue, not the tensor pipe, the limiter); only the tcgen05 instructions are predicated
    // on one lane.  Descriptors are (constant high word | low word), and the low word advances by plain adds:
    // +2 (32 B >> 4) per K step, +(tap row/column offset) per tap, +(16 rows) per stacked tile.
    if (P.dtype != PMFB_DT_F32) mma_issue_loop<true>(P, ctrl, a_buf, b_buf, b_bytes, tmem_base, total, pitch);
    else mma_issue_loop<false>(P, ctrl, a_buf, b_buf, b_bytes, tmem_base, total, pitch);
  } else {
    // ------------- epilogue: 8 warps; warp pair (q, half) shares TMEM lane quadrant q and alternates 32-column units.
    // Every unit issues its TMEM load and ALL of its residual global loads before the first use (memory-level
    // parallelism is what bounds the 32/64-channel full-resolution layers), then applies the fused functional with the
    // per-channel vectors staged in shared memory, and writes one full 128-byte line per thread.
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const float* sv = reinterpret_cast<const float*>(smem + 1024);  // [alpha1 | beta1 | alpha2 | beta2] x kHMaxC
    // Output path: each warp parks its 32-pixel x 32-channel unit in a private 4 KB staging tile (128B-swizzled rows,
    // conflict-free 16-byte stores) and one lane issues a bulk tensor store of the (32 ch, 8 x, 4 y) box: full-line
    // writes, image-edge and channel-tail clipping by the TMA unit.  (Per-thread float4 stores at a 128-byte stride
    // sustained only ~2 TB/s of stores and doubled the L1->L2 write sectors.)
    uint8_t* stage = b_buf + (size_t)P.nsb * b_bytes + (size_t)(warp - 2) * kHStageBytes;
    const uint32_t st_row = smem_u32(stage) + (uint32_t)lane * 128u;
    const uint32_t st_x = (uint32_t)(lane & 7);
    const bool tma_store = P.tma_store != 0;
    const bool hA1 = P.epi.alpha1 != nullptr, hB1 = P.epi.beta1 != nullptr, hA2 = P.epi.alpha2 != nullptr,
               hB2 = P.epi.beta2 != nullptr;
    const int chunks = (P.n_tile + 31) >> 5;
    uint32_t acc_it = 0;
    for (int w = blockIdx.x; w < total; w += gridDim.x, ++acc_it) {
      int r = w;
      const int nb = r % P.n_blocks; r /= P.n_blocks;
      const int tx = r % P.tiles_x; r /= P.tiles_x;
      const int ty = r % P.tiles_y;
      const int n_img = r / P.tiles_y;
      const int x0 = tx * 8, y0 = ty * th, n0 = nb * P.n_tile;
      const uint32_t buf = acc_it % (uint32_t)P.nacc;
      mbar_wait_sleep(&ctrl->tmem_full[buf], (acc_it / (uint32_t)P.nacc) & 1u);
      tc_fence_after();
      const int row = q * 32 + lane;
      const int units = P.mt * chunks;
      if constexpr (EPI != kEpiGeneric) {
        for (int u = half; u < units; u += 2) {
          const int j = u / chunks, ch = u - j * chunks;
          const int y = y0 + 16 * j + (row >> 3), x = x0 + (row & 7);
          const bool valid = (y < P.out_h) && (x < P.out_w);
          const int c0 = n0 + ch * 32;
          const int ncol = min(32, P.n_tile - ch * 32);
          const int nq = min(ncol, P.c_out - c0) >> 2;  // float4 groups of this unit (warp-uniform)
          const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * (uint32_t)(P.mt * P.n_tile) +
                                 (uint32_t)(j * P.n_tile + ch * 32);
          float v[32];
          tmem_ld16(taddr, v);
          if (ncol > 16) tmem_ld16(taddr + 16, v + 16);
          float* optr = P.out + ((long long)n_img * P.o_sn + (long long)y * P.o_sy + (long long)x * P.o_sx) + c0;
          float4 r1v[8];
          if constexpr ((EPI & kEpiR1) != 0) {
            const float* r1p = P.epi.r1.p + ((long long)n_img * P.epi.r1.sn + (long long)y * P.epi.r1.sy + (long long)x * P.epi.r1.sx) + c0;
            if (valid) {
#pragma unroll
              for (int i = 0; i < 8; ++i)
                if (i < nq) r1v[i] = ld4(r1p + 4 * i);
            }
          }
          if constexpr ((EPI & kEpiR2) != 0) {  // never together with r1: the same registers hold it
            const float* r2p = P.epi.r2.p + ((long long)n_img * P.epi.r2.sn + (long long)y * P.epi.r2.sy + (long long)x * P.epi.r2.sx) + c0;
            if (valid) {
#pragma unroll
              for (int i = 0; i < 8; ++i)
                if (i < nq) r1v[i] = ld4(r2p + 4 * i);
            }
          }
          // bias vector of this unit: fetched from shared memory while the TMEM load is in flight
          float4 bv[8];
          if constexpr ((EPI & kEpiB1) != 0) {
#pragma unroll
            for (int i = 0; i < 8; ++i)
              if (i < nq) bv[i] = *reinterpret_cast<const float4*>(sv + kHMaxC + c0 + 4 * i);
          }
          tmem_ld_wait();
          if (tma_store) {  // the previous unit's store must have finished reading the staging tile
            if (lane == 0) bulk_wait_group_read0();
            __syncwarp();
          }
          if constexpr ((EPI & kEpiHalf) != 0) {  // [+bias] [LeakyReLU] -> fp16, 8 channels per 16-byte store
            if (valid) {
              const uint32_t hrow = smem_u32(stage) + (uint32_t)lane * 64u;
              const uint32_t hsw = (uint32_t)(lane >> 1) & 3u;
#pragma unroll
              for (int i2 = 0; i2 < 4; ++i2) {
                if (2 * i2 < nq) {
                  uint32_t pk[4];
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    float a = v[8 * i2 + 2 * e], b = v[8 * i2 + 2 * e + 1];
                    if constexpr ((EPI & kEpiB1) != 0) {
                      const float4 bb = bv[2 * i2 + (e >> 1)];
                      a += (e & 1) ? bb.z : bb.x;
                      b += (e & 1) ? bb.w : bb.y;
                    }
                    if constexpr ((EPI & kEpiLeaky) != 0) {
                      a = fmaxf(a, 0.01f * a);
                      b = fmaxf(b, 0.01f * b);
                    }
                    const float m = 65504.f;  // saturate: an inf would poison the batch statistics
                    const __half2 h = __floats2half2_rn(fminf(fmaxf(a, -m), m), fminf(fmaxf(b, -m), m));
                    pk[e] = *reinterpret_cast<const uint32_t*>(&h);
                  }
                  st_shared_v4(hrow + ((((uint32_t)i2) ^ hsw) << 4),
                               make_float4(__uint_as_float(pk[0]), __uint_as_float(pk[1]), __uint_as_float(pk[2]), __uint_as_float(pk[3])));
                }
              }
            }
          } else
          if (valid) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              if (i < nq) {
                float4 o = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
                if constexpr ((EPI & kEpiA1) != 0) {
                  const float4 a = *reinterpret_cast<const float4*>(sv + c0 + 4 * i);
                  o.x *= a.x; o.y *= a.y; o.z *= a.z; o.w *= a.w;
                }
                if constexpr ((EPI & kEpiB1) != 0) { o.x += bv[i].x; o.y += bv[i].y; o.z += bv[i].z; o.w += bv[i].w; }
                if constexpr ((EPI & kEpiR1) != 0) { o.x += r1v[i].x; o.y += r1v[i].y; o.z += r1v[i].z; o.w += r1v[i].w; }
                if constexpr ((EPI & kEpiLeaky) != 0) {
                  o.x = fmaxf(o.x, 0.01f * o.x); o.y = fmaxf(o.y, 0.01f * o.y);
                  o.z = fmaxf(o.z, 0.01f * o.z); o.w = fmaxf(o.w, 0.01f * o.w);
                }
                if constexpr ((EPI & kEpiRelu) != 0) {  // v > 0 ? v : 0, as epi_act (NaN -> 0)
                  o.x = o.x > 0.f ? o.x : 0.f; o.y = o.y > 0.f ? o.y : 0.f; o.z = o.z > 0.f ? o.z : 0.f; o.w = o.w > 0.f ? o.w : 0.f;
                }
                if constexpr ((EPI & kEpiA2B2) != 0) {
                  const float4 a = *reinterpret_cast<const float4*>(sv + 2 * kHMaxC + c0 + 4 * i);
                  const float4 b = *reinterpret_cast<const float4*>(sv + 3 * kHMaxC + c0 + 4 * i);
                  o.x = o.x * a.x + b.x; o.y = o.y * a.y + b.y; o.z = o.z * a.z + b.z; o.w = o.w * a.w + b.w;
                }
                if constexpr ((EPI & kEpiR2) != 0) { o.x += r1v[i].x; o.y += r1v[i].y; o.z += r1v[i].z; o.w += r1v[i].w; }
                if constexpr ((EPI & kEpiRnd) != 0) { o.x = round_tf32(o.x); o.y = round_tf32(o.y); o.z = round_tf32(o.z); o.w = round_tf32(o.w); }
                if (tma_store) st_shared_v4(st_row + ((((uint32_t)i) ^ st_x) << 4), o);
                else *reinterpret_cast<float4*>(optr + 4 * i) = o;
              }
            }
          }
          if (tma_store) {
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
              tma_store_4d(&tmo, stage, c0, x0, y0 + 16 * j + 4 * q, n_img);
              bulk_commit_group();
            }
          }
          if constexpr ((EPI & kEpiStats) != 0) {
            const uint32_t vmask = __ballot_sync(0xffffffffu, valid);
            if (lane < 4 * nq) {
              float s1 = 0.f, s2 = 0.f;
              if constexpr ((EPI & kEpiHalf) != 0) {  // column `lane` of the fp16 tile: 2-byte loads, 16 banks per row
                const uint32_t col = smem_u32(stage) + (uint32_t)((lane & 7) << 1);
                const uint32_t chunk = (uint32_t)(lane >> 3);
#pragma unroll 8
                for (int r = 0; r < 32; ++r) {
                  if ((vmask >> r) & 1u) {
                    const float t = __half2float(__ushort_as_half(
                        ld_shared_u16(col + (uint32_t)r * 64u + ((chunk ^ ((uint32_t)(r >> 1) & 3u)) << 4))));
                    s1 += t;
                    s2 += t * t;
                  }
                }
              } else {
              const uint32_t col = smem_u32(stage) + (uint32_t)((lane & 3) << 2);
              const uint32_t chunk = (uint32_t)(lane >> 2);
#pragma unroll 8
              for (int r = 0; r < 32; ++r) {
                if ((vmask >> r) & 1u) {
                  const float t = ld_shared_f32(col + (uint32_t)r * 128u + ((chunk ^ (uint32_t)(r & 7)) << 4));
                  s1 += t;
                  s2 += t * t;
                }
              }
              }
              // fp64 shared accumulators: the order of the atomics must not show (a replayed step has to reproduce
              // an eager one, and near-constant channels turn fp32 ordering noise into percent-level BN changes)
              double* acc = reinterpret_cast<double*>(smem + 1024 + 2 * kHMaxC * 4) + c0 + lane;
              atomicAdd(acc, (double)s1);
              atomicAdd(acc + kHStatsC, (double)s2);
            }
          }
        }
      } else
      for (int u = half; u < units; u += 2) {
        const int j = u / chunks, ch = u - j * chunks;
        const int y = y0 + 16 * j + (row >> 3), x = x0 + (row & 7);
        const bool valid = (y < P.out_h) && (x < P.out_w);
        const int c0 = n0 + ch * 32;                       // first channel of this unit
        const int ncol = min(32, P.n_tile - ch * 32);      // 32 or 16
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * (uint32_t)(P.mt * P.n_tile) +
                               (uint32_t)(j * P.n_tile + ch * 32);
        float v[32];
        tmem_ld16(taddr, v);
        if (ncol > 16) tmem_ld16(taddr + 16, v + 16);
        const long long pix_o = (long long)n_img * P.o_sn + (long long)y * P.o_sy + (long long)x * P.o_sx;
        float4 r1v[8], mulv[8], r2v[8];
        const EpiPixel ep = epi_pixel(P.epi, n_img, y, x);
        if (valid) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int c = c0 + 4 * i;
            const bool on = (4 * i < ncol) && (c < P.c_out);
            if (ep.r1 && on) r1v[i] = ld4(ep.r1 + c);
            if (ep.mul && on) mulv[i] = ld4(ep.mul + c);
            if (ep.r2 && on) r2v[i] = ld4(ep.r2 + c);
          }
        }
        tmem_ld_wait();
        if (tma_store) {
          if (lane == 0) bulk_wait_group_read0();
          __syncwarp();
        }
        if (valid) {
          float* optr = P.out + pix_o;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int c = c0 + 4 * i;
            if ((4 * i < ncol) && (c < P.c_out)) {
              float4 o = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
              if (hA1) { const float4 a = *reinterpret_cast<const float4*>(sv + c); o.x *= a.x; o.y *= a.y; o.z *= a.z; o.w *= a.w; }
              if (hB1) { const float4 b = *reinterpret_cast<const float4*>(sv + kHMaxC + c); o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w; }
              if (ep.r1) { o.x += r1v[i].x; o.y += r1v[i].y; o.z += r1v[i].z; o.w += r1v[i].w; }
              if (P.epi.act) { o.x = epi_act(P.epi.act, o.x); o.y = epi_act(P.epi.act, o.y); o.z = epi_act(P.epi.act, o.z); o.w = epi_act(P.epi.act, o.w); }
              if (hA2) { const float4 a = *reinterpret_cast<const float4*>(sv + 2 * kHMaxC + c); o.x *= a.x; o.y *= a.y; o.z *= a.z; o.w *= a.w; }
              if (hB2) { const float4 b = *reinterpret_cast<const float4*>(sv + 3 * kHMaxC + c); o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w; }
              if (ep.mul) { o.x *= mulv[i].x; o.y *= mulv[i].y; o.z *= mulv[i].z; o.w *= mulv[i].w; }
              if (ep.r2) { o.x += r2v[i].x; o.y += r2v[i].y; o.z += r2v[i].z; o.w += r2v[i].w; }
              if (P.epi.round_out) { o.x = round_tf32(o.x); o.y = round_tf32(o.y); o.z = round_tf32(o.z); o.w = round_tf32(o.w); }
              if (tma_store) st_shared_v4(st_row + ((((uint32_t)i) ^ st_x) << 4), o);
              else *reinterpret_cast<float4*>(optr + c) = o;
            }
          }
        }
        if (tma_store) {
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            tma_store_4d(&tmo, stage, c0, x0, y0 + 16 * j + 4 * q, n_img);
            bulk_commit_group();
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&ctrl->tmem_empty[buf]);
    }
    if (tma_store && lane == 0) bulk_wait_group0();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, (uint32_t)P.tmem_cols);
  if constexpr (EPI != kEpiGeneric && (EPI & kEpiStats) != 0) {
    const double* ds = reinterpret_cast<const double*>(smem + 1024 + 2 * kHMaxC * 4);
    for (int c = threadIdx.x; c < P.c_out; c += kHThreads) {
      atomicAdd(P.bn_stats + c, ds[c]);
      atomicAdd(P.bn_stats + P.c_out + c, ds[kHStatsC + c]);
    }
  }
}

static int pow2_cols_h(int n) {
  int c = 32;
  while (c < n) c <<= 1;
  return c;
}

// Returns 1 if the descriptor can run on the halo kernel (stride-1 source, taps inside a small halo).
int halo_eligible(const pmfb_conv_desc* d) {
  if (d->x.dims[2] != 1) return 0;
  int hx = 0, hy = 0;
  for (int i = 0; i < d->n_taps; ++i) {
    if (d->tap_dp[i] != 0 || d->tap_dc[i] != 0) return 0;
    const int ax = d->tap_dw[i] < 0 ? -d->tap_dw[i] : d->tap_dw[i];
    const int ay = d->tap_dh[i] < 0 ? -d->tap_dh[i] : d->tap_dh[i];
    if (ax > hx) hx = ax;
    if (ay > hy) hy = ay;
  }
  return (hx <= 2 && hy <= 3 && d->c_out <= kHMaxC) ? 1 : 0;
}

template <int EPI>
static int launch_halo_t(int grid, size_t smem, cudaStream_t stream, const CUtensorMap& tmx, const CUtensorMap& tmw,
                         const CUtensorMap& tmo, const HaloK& P) {
  static bool attr_set = false;
  if (!attr_set) {
    PMFB_CUDA_CHECK(cudaFuncSetAttribute(conv_fwd_halo_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, kHSmemBudget + 2048));
    attr_set = true;
  }
  conv_fwd_halo_kernel<EPI><<<grid, kHThreads, smem, stream>>>(tmx, tmw, tmo, P);
  PMFB_LAUNCH_CHECK("conv_fwd_halo_kernel");
  return PMFB_OK;
}

static int launch_halo_variant(int epi, int grid, size_t smem, cudaStream_t stream, const CUtensorMap& tmx, const CUtensorMap& tmw,
                               const CUtensorMap& tmo, const HaloK& P) {
  switch (epi) {
#define PMFB_HV(e) case e: return launch_halo_t<e>(grid, smem, stream, tmx, tmw, tmo, P);
    PMFB_HV(0) PMFB_HV(1) PMFB_HV(2) PMFB_HV(3) PMFB_HV(4) PMFB_HV(5) PMFB_HV(6) PMFB_HV(7)
    PMFB_HV(8) PMFB_HV(9) PMFB_HV(12) PMFB_HV(13)
    PMFB_HV(16) PMFB_HV(17) PMFB_HV(24) PMFB_HV(25)
    PMFB_HV(16 | kEpiHalf) PMFB_HV(17 | kEpiHalf) PMFB_HV(24 | kEpiHalf) PMFB_HV(25 | kEpiHalf)
    PMFB_HV(kEpiB1 | kEpiLeaky | kEpiA2B2 | kEpiRnd) PMFB_HV(kEpiB1 | kEpiLeaky | kEpiA2B2 | kEpiRnd | kEpiR2)
    PMFB_HV(kEpiA1 | kEpiB1 | kEpiRelu | kEpiRnd) PMFB_HV(kEpiA1 | kEpiB1 | kEpiRelu | kEpiRnd | kEpiR1)
    PMFB_HV(kEpiA1 | kEpiB1 | kEpiRnd)
#undef PMFB_HV
    default: return launch_halo_t<kEpiGeneric>(grid, smem, stream, tmx, tmw, tmo, P);
  }
}

static int halo_fast_epi(const pmfb_conv_desc* d) {
  const pmfb_epilogue& E = d->epi;
  if (E.mul.ptr || E.act == PMFB_ACT_SIGMOID || (E.alpha2 != nullptr) != (E.beta2 != nullptr)) return kEpiGeneric;
  int m = (E.alpha1 ? kEpiA1 : 0) | (E.beta1 ? kEpiB1 : 0) | (E.r1.ptr ? kEpiR1 : 0) | (E.round_out ? kEpiRnd : 0) |
          (E.act == PMFB_ACT_LEAKY ? kEpiLeaky : 0) | (E.act == PMFB_ACT_RELU ? kEpiRelu : 0) | (E.alpha2 ? kEpiA2B2 : 0) |
          (E.r2.ptr ? kEpiR2 : 0);
  // note: alpha2/beta2 in the fast path are one fused multiply-add, the generic path multiplies then adds (same fp32
  // results up to the fma's single rounding; within the 1e-3 parity bar, eval mode only)
  switch (m) {
    // training step: [+bias][LeakyReLU][round], [+= r1][round]
    case 0: case 1: case 4: case 5: case 8: case 9: case 12: case 13: case 2: case 6:
    // inference: conv -> LeakyReLU -> BN(eval) [+ shortcut]; conv -> BN(eval) [+ identity] -> ReLU; conv -> BN(eval)
    case kEpiB1 | kEpiLeaky | kEpiA2B2 | kEpiRnd:
    case kEpiB1 | kEpiLeaky | kEpiA2B2 | kEpiRnd | kEpiR2:
    case kEpiA1 | kEpiB1 | kEpiRelu | kEpiRnd:
    case kEpiA1 | kEpiB1 | kEpiRelu | kEpiRnd | kEpiR1:
    case kEpiA1 | kEpiB1 | kEpiRnd:
      return m;
    default:
      return kEpiGeneric;
  }
}

// Fused BN statistics: stride-1 halo layers whose epilogue is [+bias] [LeakyReLU] without rounding / accumulation, with the
// bulk-store epilogue (its staging tile is what the statistics read) and blocks that are multiples of 32 channels.
int halo_fused_stats_ok(const pmfb_conv_desc* d) {
  const char* e = getenv("PMFB_HALO_TMA_STORE");
  if (e && !atoi(e)) return 0;
  const char* f = getenv("PMFB_HALO_FAST_EPI");
  if (f && !atoi(f)) return 0;
  const char* g = getenv("PMFB_FUSED_BN_STATS");
  if (g && !atoi(g)) return 0;
  const int epi = halo_fast_epi(d);
  if (epi == kEpiGeneric || (epi & ~(kEpiB1 | kEpiLeaky))) return 0;
  if ((reinterpret_cast<uintptr_t>(d->out) & 15) || d->o_sx % 4 || d->o_sy % 4 || d->o_sn % 4) return 0;
  if (d->out_half && (d->c_out % 8 || d->o_sx % 8 || d->o_sy % 8 || d->o_sn % 8)) return 0;
  return d->c_out <= kHStatsC ? 1 : 0;
}

int launch_conv_halo(const pmfb_conv_desc* d, void* stream) {
  static int sm_count = 0;
  static int base_off_mode = -1;
  if (sm_count == 0) {
    int dev = 0;
    PMFB_CUDA_CHECK(cudaGetDevice(&dev));
    PMFB_CUDA_CHECK(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));
    const char* e = getenv("PMFB_HALO_BASEOFF");
    base_off_mode = e ? atoi(e) : 0;  // measured on B200: the 128B swizzle is a pure function of the smem address, no base offset needed
  }
  HaloK P;
  P.n_taps = d->n_taps;
  P.dtype = d->dtype;
  // 16-bit layers with <= 32 input channels use 64-byte operand rows (SWIZZLE_64B): a 128-byte row would be half zero fill
  const bool sw64 = d->dtype != PMFB_DT_F32 && d->c_in <= 32;
  P.row_bytes = sw64 ? 64 : 128;
  P.kslab = d->dtype == PMFB_DT_F32 ? 32 : (sw64 ? 32 : 64);
  if (d->dtype != PMFB_DT_F32 && d->c_in % 8) return fail(PMFB_ERR_INVALID, "conv halo: 16-bit operands need c_in %% 8 == 0 (c_in=%d)", d->c_in);
  P.ks = (d->c_in + P.kslab - 1) / P.kslab;
  {
    const int per_step = d->dtype == PMFB_DT_F32 ? 8 : 16;  // channels per 32-byte K step
    const int rem = d->c_in - (P.ks - 1) * P.kslab;
    P.klast = (rem + per_step - 1) / per_step;
  }
  P.hx = P.hy = 0;
  for (int i = 0; i < PMFB_MAX_TAPS; ++i) {
    P.tap_dw[i] = d->tap_dw[i];
    P.tap_dh[i] = d->tap_dh[i];
    P.tap_wi[i] = d->use_tap_wi ? d->tap_wi[i] : i;
    if (i < d->n_taps) {
      const int ax = d->tap_dw[i] < 0 ? -d->tap_dw[i] : d->tap_dw[i];
      const int ay = d->tap_dh[i] < 0 ? -d->tap_dh[i] : d->tap_dh[i];
      if (ax > P.hx) P.hx = ax;
      if (ay > P.hy) P.hy = ay;
    }
  }
  P.tiles_x = (d->out_w + 7) / 8;
  // Tile shape: (MT stacked 128-pixel tiles) x (n_tile output channels) per work item, chosen by a cost model fitted
  // to measurements on B200 (profiles/r1_halo_model.md):
  //   * one kind::tf32 UMMA (M=128, N=n, K=8, both operands in shared memory) occupies the tensor pipe for
  //     (128 + n) / 2 cycles: the operand fetch (128 + n rows x 32 B at 64 B/clk), not the math (n / 2), is the limiter,
  //     so the widest n_tile always wins on MMA time and stacking tiles (MT) only amortises the weight fetch;
  //   * an item also needs its L2->SM fill (~43 B/clk) and its epilogue (TMEM -> staging -> bulk store); with two
  //     TMEM accumulator sets (2*MT*n_tile <= 512 columns) these overlap the next item's MMAs, otherwise the
  //     epilogue is serialised behind them;
  //   * persistent CTAs walk the item list round-robin: time = ceil(items / SMs) * item time.
  const int c_out16 = ((d->c_out + 15) / 16) * 16;
  int n_tile = c_out16 < 256 ? c_out16 : 256;
  int mt = 2;
  {
    double best = 1e30;
    const int n_full = n_tile;
    // split blocks stay multiples of 32 channels: the epilogue's store boxes are 32 channels wide and only the tensor's
    // own channel count clips them
    const int n_cands[3] = {n_full, (n_full % 64 == 0 && n_full >= 128) ? n_full / 2 : 0,
                            (n_full % 128 == 0 && n_full >= 256) ? n_full / 4 : 0};
    const double ksteps = (4.0 * (P.ks - 1) + P.klast) * d->n_taps;
    for (int mi = 2; mi >= 1; --mi)
      for (int ni = 0; ni < 3; ++ni) {
        const int nt = n_cands[ni];
        if (nt == 0 || mi * nt > 512) continue;
        // shared memory: two halo tiles + >= 2 weight stages + the epilogue staging tiles must fit
        const int a_b = ((16 * mi + 2 * P.hy) * (8 + 2 * P.hx) * P.row_bytes + 1023) & ~1023;
        const int nb_fit = (kHSmemBudget - kHCtrlBytes - 2 * a_b - kHEpiWarps * kHStageBytes) / (nt * P.row_bytes);
        if (nb_fit < 2) continue;
        const long long items = (long long)P.tiles_x * ((d->out_h + 16 * mi - 1) / (16 * mi)) * d->n_batch * ((d->c_out + nt - 1) / nt);
        const long long waves = (items + sm_count - 1) / sm_count;
        // + ~200 cycles per (slab, tap): barrier round trip and descriptor arithmetic of the issuing warp that the
        // tensor pipe does not hide (fitted: 64-channel layers lose 30 % with MT=1, 256-channel ones gain 25 %)
        const double mma = ksteps * mi * (128.0 + nt) * 0.5 + 200.0 * P.ks * d->n_taps;
        const double bytes = ksteps * 8.0 * 4.0 * nt + (double)P.ks * (16 * mi + 2 * P.hy) * (8 + 2 * P.hx) * 128.0;
        const double fill = bytes / 43.0;
        const double epi = 400.0 + mi * 128.0 * nt * 4.0 / 24.0;
        const double body = mma > fill ? mma : fill;
        const double item = (2 * mi * nt <= 512 ? (body > epi ? body : epi) : body + epi) + 600.0;
        const double cost = (double)waves * item * (nb_fit < 4 ? 1.1 : 1.0);  // a 2-3 deep weight ring exposes L2 latency
        if (cost < best) {
          best = cost;
          mt = mi;
          n_tile = nt;
        }
      }
  }
  P.n_tile = n_tile;
  P.n_blocks = (d->c_out + n_tile - 1) / n_tile;
  P.c_out = d->c_out;
  P.out_h = d->out_h;
  P.out_w = d->out_w;
  P.n_batch = d->n_batch;
  P.mt = mt;
  P.tiles_y = (d->out_h + 16 * mt - 1) / (16 * mt);
  P.nacc = (2 * mt * n_tile <= 512) ? 2 : 1;
  P.tmem_cols = pow2_cols_h(P.nacc * mt * n_tile);
  const int rows = 16 * mt + 2 * P.hy, pitch = 8 + 2 * P.hx;
  for (int i = 0; i < PMFB_MAX_TAPS; ++i)
    P.tap_off[i] = i < d->n_taps ? (((P.tap_dh[i] + P.hy) * pitch + P.tap_dw[i] + P.hx) * P.row_bytes) >> 4 : 0;
  P.a_box_bytes = rows * pitch * P.row_bytes;
  P.a_bytes = (P.a_box_bytes + 1023) & ~1023;
  const int b_bytes = n_tile * P.row_bytes;
  static int tma_store_mode = -1;
  if (tma_store_mode < 0) {
    const char* e = getenv("PMFB_HALO_TMA_STORE");
    tma_store_mode = e ? atoi(e) : 1;
  }
  // the bulk tensor store needs a 16-byte aligned base and strides; every view the executor builds satisfies this
  P.tma_store = (tma_store_mode && (reinterpret_cast<uintptr_t>(d->out) & 15) == 0 && d->o_sx % 4 == 0 && d->o_sy % 4 == 0 &&
                 d->o_sn % 4 == 0 && (P.n_blocks == 1 || n_tile % 32 == 0)) ? 1 : 0;
  const int stage_total = P.tma_store ? kHEpiWarps * kHStageBytes : 0;
  int nsb = (kHSmemBudget - kHCtrlBytes - 2 * P.a_bytes - stage_total) / b_bytes;
  if (nsb > kHMaxB) nsb = kHMaxB;
  if (nsb < 2) return fail(PMFB_ERR_INVALID, "conv halo: shared memory budget exceeded (n_tile=%d)", n_tile);
  if (d->c_out > kHMaxC) return fail(PMFB_ERR_INVALID, "conv halo: c_out=%d > %d", d->c_out, kHMaxC);
  P.nsb = nsb;
  P.use_base_off = base_off_mode;
  P.out = d->out;
  P.o_sn = d->o_sn;
  P.o_sy = d->o_sy;
  P.o_sx = d->o_sx;
  int rc = epi_from_c(&d->epi, &P.epi);
  if (rc) return rc;

  CUtensorMap tmx, tmw;
  const bool h16 = d->dtype != PMFB_DT_F32;
  const uint64_t esz = h16 ? 2 : 4;
  uint32_t boxx[5] = {(uint32_t)P.kslab, (uint32_t)pitch, 1, (uint32_t)rows, 1};
  rc = h16 ? make_tmap_16(&tmx, d->x.ptr, 5, d->x.dims, d->x.strides, boxx, d->dtype == PMFB_DT_BF16, sw64)
           : make_tmap_f32(&tmx, d->x.ptr, 5, d->x.dims, d->x.strides, boxx);
  if (rc) return rc;
  int n_slabs = d->n_taps;
  if (d->use_tap_wi)
    for (int i = 0; i < d->n_taps; ++i)
      if (d->tap_wi[i] + 1 > n_slabs) n_slabs = d->tap_wi[i] + 1;
  uint64_t wdims[3] = {(uint64_t)d->c_in, (uint64_t)d->c_out, (uint64_t)n_slabs};
  uint64_t wstr[2] = {(uint64_t)d->c_in * esz, (uint64_t)d->c_in * d->c_out * esz};
  uint32_t boxw[3] = {(uint32_t)P.kslab, (uint32_t)n_tile, 1};
  rc = h16 ? make_tmap_16(&tmw, d->w, 3, wdims, wstr, boxw, d->dtype != PMFB_DT_F16, sw64) : make_tmap_f32(&tmw, d->w, 3, wdims, wstr, boxw);
  if (rc) return rc;

  CUtensorMap tmo = tmw;  // placeholder when the direct-store path is used
  if (P.tma_store) {
    uint64_t odims[4] = {(uint64_t)d->c_out, (uint64_t)d->out_w, (uint64_t)d->out_h, (uint64_t)d->n_batch};
    const uint64_t osz = d->out_half ? 2 : 4;
    uint64_t ostr[3] = {(uint64_t)d->o_sx * osz, (uint64_t)d->o_sy * osz, (uint64_t)d->o_sn * osz};
    uint32_t boxo[4] = {32, 8, 4, 1};
    rc = d->out_half ? make_tmap_16(&tmo, d->out, 4, odims, ostr, boxo, false, true) : make_tmap_f32(&tmo, d->out, 4, odims, ostr, boxo);
    if (rc) return rc;
  }
  const size_t smem = (size_t)kHCtrlBytes + 2 * (size_t)P.a_bytes + (size_t)nsb * b_bytes + stage_total + 1024;
  const long long total = (long long)P.tiles_x * P.tiles_y * P.n_batch * P.n_blocks;
  const int grid = (int)(total < sm_count ? total : sm_count);
  // epilogue variant
  int epi = kEpiGeneric;
  static int fast_mode = -1;
  if (fast_mode < 0) {
    const char* e = getenv("PMFB_HALO_FAST_EPI");
    fast_mode = e ? atoi(e) : 1;
  }
  if (fast_mode) epi = halo_fast_epi(d);
  P.bn_stats = d->bn_stats;
  if (d->bn_stats) {
    if (!halo_fused_stats_ok(d) || !P.tma_store || epi == kEpiGeneric)
      return fail(PMFB_ERR_INVALID, "conv halo: fused BN statistics are not available for this layer");
    epi |= kEpiStats;
    if (d->out_half) epi |= kEpiHalf;
  } else if (d->out_half) {
    return fail(PMFB_ERR_INVALID, "conv halo: fp16 output exists only with the fused-statistics epilogue");
  }
  return launch_halo_variant(epi, grid, smem, (cudaStream_t)stream, tmx, tmw, tmo, P);
}

}  // namespace pmfb
