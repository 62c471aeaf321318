// conv_wgrad_halo_kernel — weight gradient of the stride-1 layers with ALL taps accumulated in TMEM.
//
//   dw[tap][ci][co] += sum_pixels x[pixel + tap][ci] * dy[pixel][co]
//
// The first wgrad kernel (conv_tc.cu) gives every CTA four (tap, 32-channel) row blocks, so x is re-read once per tap
// and dy once per row-block group: at full resolution (tensors far larger than L2) a 3x3 layer pulled ~12x its
// operands from HBM and ran at 4.7 TB/s of redundant traffic.  Here a CTA owns one (128-input-channel group,
// output-channel block) pair and a strided share of the 8x8-pixel tiles; per tile it loads the x halo tile
// ((8+2hy) x (8+2hx) pixels, one TMA box per 32 channels) and the dy tile ONCE, and issues, for every tile row (K = 8
// pixels) and every tap, one UMMA whose A descriptor points at the tap-shifted window of the halo tile
// (both operands MN-major, SWIZZLE_128B_BASE32B, exactly as in conv_tc.cu).  Tap t accumulates in TMEM columns
// [t*n_tile, (t+1)*n_tile) — taps*n_tile <= 512 — for the CTA's whole pixel share; one fp32 vector-atomic flush into
// the packed gradient at the end (split-K across the CTAs of a group).
//
// Warp roles (192 threads): warp 0 TMA producer, warp 1 TMEM allocator + MMA issuer, warps 2..5 flush.
#include <stdlib.h>

#include "common.h"
#include "ptx.cuh"

namespace pmfb {

constexpr int kWThreads = 192;
constexpr int kWCtrlBytes = 1024;
constexpr int kWMaxStages = 6;
constexpr int kWSmemBudget = 208 * 1024;

struct WCtrl {
  uint64_t full[kWMaxStages], empty[kWMaxStages];
  uint64_t tmem_full;
  uint32_t tmem_base;
};

struct WHaloK {
  int n_taps;
  int tap_dw[PMFB_MAX_TAPS], tap_dh[PMFB_MAX_TAPS];
  int hx, hy;
  int tiles_x, tiles_y, n_batch;
  int c_in, c_out, n_tile, n_blocks, c_groups, split;
  int nblk_b;                 // 32-channel dy boxes per stage
  // Tap packing (c_in <= 32): the four 32-row blocks of the M=128 A operand are up to four TAPS of the same 32-channel
  // block instead of four channel blocks -- the descriptor's leading byte offset is the (constant) distance between the
  // taps of a group inside the halo tile -- so a 3x3 layer issues 3 UMMAs per tile row instead of 9.
  int packed, n_groups;
  int gps, tsplits;                // tap groups per CTA and the number of such tap ranges (TMEM holds gps * n_tile columns)
  int grp_off[PMFB_MAX_TAPS];      // first tap's window offset inside the halo tile, in pixels (128-byte rows)
  int grp_lbo[PMFB_MAX_TAPS];      // distance between consecutive taps of the group, in pixels
  int grp_cnt[PMFB_MAX_TAPS];
  int grp_tap[PMFB_MAX_TAPS][4];   // weight-gradient slab of each packed tap
  int a_blk_bytes, a_span, b_blk_bytes, stage_bytes, stages, tmem_cols;
  int dtype, kslab;  // PMFB_DT_F32 (tf32, 32-channel boxes) or PMFB_DT_BF16 (kind::f16, 64-channel boxes)
  float* dw;
};

__global__ void __launch_bounds__(kWThreads, 1)
conv_wgrad_halo_kernel(const __grid_constant__ CUtensorMap tmx, const __grid_constant__ CUtensorMap tmdy,
                       const __grid_constant__ WHaloK P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  WCtrl* ctrl = reinterpret_cast<WCtrl*>(smem);
  uint8_t* tiles = smem + kWCtrlBytes;

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const int pitch = 8 + 2 * P.hx;

  // CTA -> (group, split index); group -> (channel group, output block)
  const int grp0 = blockIdx.x / P.split, sidx = blockIdx.x - grp0 * P.split;
  const int grp = grp0 / P.tsplits, ts = grp0 - grp * P.tsplits;
  const int cg = grp / P.n_blocks, nb = grp - cg * P.n_blocks;
  const int g_begin = ts * P.gps, g_end = min(g_begin + P.gps, P.n_groups);
  const int ci0 = cg * 128, n0 = nb * P.n_tile;
  int nblk_a = (P.c_in - ci0 + P.kslab - 1) / P.kslab;
  if (nblk_a > 128 / P.kslab) nblk_a = 128 / P.kslab;
  const int total_tiles = P.tiles_x * P.tiles_y * P.n_batch;
  const int iters = (total_tiles - sidx + P.split - 1) / P.split;  // tiles sidx, sidx+split, ...

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
        const uint32_t tx_bytes = (uint32_t)(nblk_a * ((8 + 2 * P.hy) * pitch * 128) + P.nblk_b * P.b_blk_bytes);
        for (int it = 0; it < iters; ++it) {
          int r = sidx + it * P.split;
          const int tx = r % P.tiles_x; r /= P.tiles_x;
          const int ty = r % P.tiles_y;
          const int n_img = r / P.tiles_y;
          const int x0 = tx * 8, y0 = ty * 8;
          const int s = it % P.stages;
          mbar_wait(&ctrl->empty[s], ((uint32_t)(it / P.stages) & 1u) ^ 1u);
          uint8_t* a_s = tiles + (size_t)s * P.stage_bytes;
          uint8_t* b_s = a_s + (size_t)P.a_span;
          mbar_expect_tx(&ctrl->full[s], tx_bytes);
          for (int i = 0; i < nblk_a; ++i)
            tma_load_5d(a_s + (size_t)i * P.a_blk_bytes, &tmx, &ctrl->full[s], ci0 + i * P.kslab, x0 - P.hx, 0, y0 - P.hy, n_img);
          for (int j = 0; j < P.nblk_b; ++j)
            tma_load_5d(b_s + (size_t)j * P.b_blk_bytes, &tmdy, &ctrl->full[s], n0 + j * P.kslab, x0, 0, y0, n_img);
        }
      }
    } else if (warp == 1 && P.dtype != PMFB_DT_F32) {
      // ---- 16-bit operands (bf16 x bf16, kind::f16, K = 16 pixels per UMMA).  Both operands MN-major with the plain
      // 128B swizzle: a pixel row holds 64 channels (128 B), a swizzle atom is 8 pixel rows (1024 B).  One K step spans two
      // atoms: for dy (dense 8x8 tile) they are 1024 B apart, for the x halo tile two consecutive 8-pixel tile rows, i.e.
      // pitch*128 B apart -- the stride-byte-offset; the leading-byte-offset steps over 64-channel blocks.
      const uint32_t idesc = make_idesc_f16(128, (uint32_t)P.n_tile, 1, 1, 1, 1);
      const uint32_t hi_a = ((((uint32_t)pitch * 128u) >> 4) & 0x3FFFu) | (1u << 14) | (2u << 29);
      const uint32_t hi_b = ((1024u >> 4) & 0x3FFFu) | (1u << 14) | (2u << 29);
      const uint32_t lbo_a = (((uint32_t)P.a_blk_bytes >> 4) & 0x3FFFu) << 16;
      const uint32_t lbo_b = (((uint32_t)P.b_blk_bytes >> 4) & 0x3FFFu) << 16;
      const uint32_t a_kstep = (uint32_t)(2 * pitch * 128) >> 4;
      for (int it = 0, s = 0, s_ph = 0; it < iters; ++it) {
        mbar_wait(&ctrl->full[s], (uint32_t)s_ph);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(tiles + (size_t)s * P.stage_bytes);
        const uint32_t a_lo00 = ((a_addr & 0x3FFFFu) >> 4) | lbo_a;
        const uint32_t b_lo0 = (((a_addr + (uint32_t)P.a_span) & 0x3FFFFu) >> 4) | lbo_b;
        for (int t = g_begin; t < g_end; ++t) {
          const uint32_t a_tap = a_lo00 + (uint32_t)((P.grp_off[t] * 128) >> 4);
          const uint32_t d_col = tmem_base + (uint32_t)((t - g_begin) * P.n_tile);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {  // two tile rows (16 pixels) per K step
            umma_issue<true>(d_col, a_tap + (uint32_t)ks * a_kstep, hi_a, b_lo0 + 128u * ks, hi_b, idesc, (uint32_t)(it | ks));
          }
        }
        umma_commit_warp(&ctrl->empty[s]);
        if (++s == P.stages) {
          s = 0;
          s_ph ^= 1;
        }
      }
      umma_commit_warp(&ctrl->tmem_full);
    } else if (warp == 1) {
      // whole warp, uniform descriptors, one elected lane issues (see conv_halo.cu)
      const uint32_t idesc = make_idesc_tf32(128, (uint32_t)P.n_tile, 1, 1);
      const uint32_t hi = ((512u >> 4) & 0x3FFFu) | (1u << 14) | (1u << 29);  // SBO 512 B, SWIZZLE_128B_BASE32B
      const uint32_t lbo_a = (((uint32_t)P.a_blk_bytes >> 4) & 0x3FFFu) << 16;
      const uint32_t lbo_b = (((uint32_t)P.b_blk_bytes >> 4) & 0x3FFFu) << 16;
      for (int it = 0, s = 0, s_ph = 0; it < iters; ++it) {
        mbar_wait(&ctrl->full[s], (uint32_t)s_ph);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(tiles + (size_t)s * P.stage_bytes);
        const uint32_t a_lo00 = (a_addr & 0x3FFFFu) >> 4;
        const uint32_t b_lo0 = (((a_addr + (uint32_t)P.a_span) & 0x3FFFFu) >> 4) | lbo_b;
        for (int t = g_begin; t < g_end; ++t) {
          const uint32_t lbo = P.packed ? ((((uint32_t)P.grp_lbo[t] * 128u) >> 4) & 0x3FFFu) << 16 : lbo_a;
          const uint32_t a_tap = a_lo00 + lbo + (uint32_t)((P.grp_off[t] * 128) >> 4);
          const uint32_t d_col = tmem_base + (uint32_t)((t - g_begin) * P.n_tile);
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {  // one tile row (8 pixels) per K step
            umma_issue<false>(d_col, a_tap + (uint32_t)(ks * pitch * 8), hi, b_lo0 + 64u * ks, hi, idesc, (uint32_t)(it | ks));
          }
        }
        umma_commit_warp(&ctrl->empty[s]);
        if (++s == P.stages) {
          s = 0;
          s_ph ^= 1;
        }
      }
      umma_commit_warp(&ctrl->tmem_full);
    } else {
      const int q = warp & 3;
      const int ci = P.packed ? lane : ci0 + q * 32 + lane;
      const bool valid = ci < P.c_in;
      mbar_wait_sleep(&ctrl->tmem_full, 0);
      tc_fence_after();
      for (int t = g_begin; t < g_end; ++t) {
        if (P.packed && q >= P.grp_cnt[t]) continue;  // warp-uniform
        const int tap = P.packed ? P.grp_tap[t][q] : t;
        float* dst = P.dw + ((long long)tap * P.c_in + ci) * P.c_out;
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((t - g_begin) * P.n_tile);
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
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, (uint32_t)P.tmem_cols);
}

int wgrad_halo_eligible(const pmfb_wgrad_desc* d) {
  if (d->x.dims[2] != 1) return 0;
  int hx = 0, hy = 0;
  for (int i = 0; i < d->n_taps; ++i) {
    if (d->tap_dp[i] != 0 || d->tap_dc[i] != 0) return 0;
    const int ax = d->tap_dw[i] < 0 ? -d->tap_dw[i] : d->tap_dw[i];
    const int ay = d->tap_dh[i] < 0 ? -d->tap_dh[i] : d->tap_dh[i];
    if (ax > hx) hx = ax;
    if (ay > hy) hy = ay;
  }
  return (hx <= 2 && hy <= 3) ? 1 : 0;
}

int launch_wgrad_halo(const pmfb_wgrad_desc* d, void* stream) {
  static int sm_count = 0;
  if (sm_count == 0) {
    int dev = 0;
    PMFB_CUDA_CHECK(cudaGetDevice(&dev));
    PMFB_CUDA_CHECK(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));
  }
  WHaloK P;
  P.n_taps = d->n_taps;
  P.dtype = d->dtype;
  const bool h16 = d->dtype != PMFB_DT_F32;
  if (h16 && d->dtype != PMFB_DT_BF16) return fail(PMFB_ERR_INVALID, "wgrad halo: 16-bit operands must be bf16");
  if (h16 && (d->c_in % 8 || d->c_out % 8)) return fail(PMFB_ERR_INVALID, "wgrad halo: bf16 operands need c_in, c_out multiples of 8");
  P.kslab = h16 ? 64 : 32;
  P.hx = P.hy = 0;
  for (int i = 0; i < PMFB_MAX_TAPS; ++i) {
    P.tap_dw[i] = d->tap_dw[i];
    P.tap_dh[i] = d->tap_dh[i];
    if (i < d->n_taps) {
      const int ax = d->tap_dw[i] < 0 ? -d->tap_dw[i] : d->tap_dw[i];
      const int ay = d->tap_dh[i] < 0 ? -d->tap_dh[i] : d->tap_dh[i];
      if (ax > P.hx) P.hx = ax;
      if (ay > P.hy) P.hy = ay;
    }
  }
  P.c_in = d->c_in;
  P.c_out = d->c_out;
  P.n_batch = d->n_batch;
  P.tiles_x = (d->out_w + 7) / 8;
  P.tiles_y = (d->out_h + 7) / 8;
  const int pitch = 8 + 2 * P.hx, rows = 8 + 2 * P.hy;
  // ---- tap groups.  Unpacked: one group per tap.  Packed (c_in <= 32): greedy arithmetic progressions of up to four
  // window offsets (e.g. the three taps of one 3x3 kernel row: distance = dilation pixels; the 7 vertical taps of the
  // unrolled stem: distance = one tile row).
  static int pack_mode = -1;
  if (pack_mode < 0) {
    const char* e = getenv("PMFB_WGRAD_PACK");
    pack_mode = e ? atoi(e) : 1;
  }
  P.packed = (pack_mode && d->c_in <= 32 && d->n_taps > 1 && !h16) ? 1 : 0;
  P.n_groups = 0;
  {
    int off[PMFB_MAX_TAPS], used[PMFB_MAX_TAPS];
    for (int i = 0; i < d->n_taps; ++i) {
      off[i] = (d->tap_dh[i] + P.hy) * pitch + d->tap_dw[i] + P.hx;
      used[i] = 0;
    }
    for (;;) {
      int first = -1;
      for (int i = 0; i < d->n_taps; ++i)
        if (!used[i] && (first < 0 || off[i] < off[first])) first = i;
      if (first < 0) break;
      const int g = P.n_groups++;
      used[first] = 1;
      P.grp_off[g] = off[first];
      P.grp_cnt[g] = 1;
      P.grp_lbo[g] = 0;
      P.grp_tap[g][0] = first;
      if (!P.packed) continue;
      int next = -1;
      for (int i = 0; i < d->n_taps; ++i)
        if (!used[i] && (next < 0 || off[i] < off[next])) next = i;
      if (next < 0) continue;
      const int delta = off[next] - off[first];
      if (delta <= 0 || delta * 8 > 0x3FFF) continue;
      int last = off[first];
      while (P.grp_cnt[g] < 4) {
        int hit = -1;
        for (int i = 0; i < d->n_taps; ++i)
          if (!used[i] && off[i] == last + delta) hit = i;
        if (hit < 0) break;
        used[hit] = 1;
        P.grp_tap[g][P.grp_cnt[g]++] = hit;
        last += delta;
      }
      P.grp_lbo[g] = delta;
    }
  }
  // Every tap group a CTA owns lives in TMEM: gps * n_tile <= 512 columns.  One UMMA (M=128, N=n, K=8) costs about
  // (128 + n) / 2 cycles of operand fetch, i.e. (128 + n) / (2 n) cycles per accumulator column: with all nine taps of a
  // 3x3 kernel in one CTA n_tile is 48 (1.8 cycles/column); giving a CTA one kernel row (gps = 3) allows n_tile = 128-160
  // (1.0) at the price of fetching the x / dy tiles once per tap range.  Pick the split with the smallest
  // max(MMA, L2->SM fill) time.
  const int c_out16 = (d->c_out + 15) & ~15;
  const int c_groups = (d->c_in + 127) / 128;
  int nblk_a_max = (d->c_in + P.kslab - 1) / P.kslab;
  if (nblk_a_max > 128 / P.kslab) nblk_a_max = 128 / P.kslab;
  const int blocks_m = 128 / P.kslab;  // operand blocks the M = 128 descriptor spans
  const int a_blk = (rows * pitch * 128 + 1023) & ~1023;
  int n_tile = 0, gps = P.n_groups;
  {
    double best = 1e30;
    const int cands[4] = {P.n_groups, 3, 2, 1};
    for (int ci = 0; ci < 4; ++ci) {
      const int g = cands[ci];
      if (g > P.n_groups || (ci > 0 && g == P.n_groups)) continue;
      int nt = (512 / g) & ~15;
      if (nt > 256) nt = 256;
      if (nt > c_out16) nt = c_out16;
      if (nt < 16) continue;
      const int nbk = (d->c_out + nt - 1) / nt;
      const int bal = (((d->c_out + nbk - 1) / nbk) + 15) & ~15;  // balance the output blocks (128 -> 3 x 48, not 48+48+32)
      if (bal < nt) nt = bal;
      const int tsp = (P.n_groups + g - 1) / g;
      // shared memory: at least two (x halo blocks + dy blocks) stages, three preferred
      const int stage_b = nblk_a_max * a_blk + ((nt + P.kslab - 1) / P.kslab) * 8192;
      const int st_fit = (kWSmemBudget - kWCtrlBytes - (blocks_m - nblk_a_max) * a_blk) / stage_b;
      if (st_fit < 2) continue;
      const double mma = (double)c_groups * nbk * P.n_groups * (h16 ? 4.0 : 8.0) * (128.0 + nt) * 0.5;
      const double fill = (double)c_groups * nbk * tsp * (nblk_a_max * (double)a_blk + ((nt + P.kslab - 1) / P.kslab) * 8192.0) / 43.0;
      const double cost = ((mma > fill ? mma : fill) + 0.25 * (mma > fill ? fill : mma)) * (st_fit < 3 ? 1.15 : 1.0);
      if (cost < best) {
        best = cost;
        n_tile = nt;
        gps = g;
      }
    }
  }
  if (n_tile < 16) return fail(PMFB_ERR_INVALID, "wgrad halo: too many taps (%d)", d->n_taps);
  static int gps_override = -1;
  if (gps_override < 0) {
    const char* e = getenv("PMFB_WGRAD_GPS");
    gps_override = e ? atoi(e) : 0;
  }
  if (gps_override == 99) {  // the previous behaviour: every tap group in one CTA
    gps = P.n_groups;
    n_tile = (512 / P.n_groups) & ~15;
    if (n_tile > 256) n_tile = 256;
    if (n_tile > c_out16) n_tile = c_out16;
    const int nbk = (d->c_out + n_tile - 1) / n_tile;
    const int bal = (((d->c_out + nbk - 1) / nbk) + 15) & ~15;
    if (bal < n_tile) n_tile = bal;
  }
  P.gps = gps;
  P.tsplits = (P.n_groups + gps - 1) / gps;
  P.n_tile = n_tile;
  P.n_blocks = (d->c_out + n_tile - 1) / n_tile;
  P.c_groups = (d->c_in + 127) / 128;
  const int groups = P.n_blocks * P.c_groups * P.tsplits;
  const long long total_tiles = (long long)P.tiles_x * P.tiles_y * P.n_batch;
  int split = sm_count / groups;
  if (split < 1) split = 1;
  if (split > total_tiles) split = (int)total_tiles;
  P.split = split;
  P.nblk_b = (n_tile + P.kslab - 1) / P.kslab;
  P.a_blk_bytes = (rows * pitch * 128 + 1023) & ~1023;
  P.b_blk_bytes = 64 * 128;
  // Only the channel blocks that exist are loaded and given shared memory; the A descriptor still spans four blocks
  // (LBO = a_blk_bytes), so for c_in < 128 its upper rows read whatever follows (dy tile / next stage / tail pad):
  // those accumulator rows are never flushed, and every D row depends on its own A row only.
  P.a_span = nblk_a_max * P.a_blk_bytes;
  P.stage_bytes = P.a_span + P.nblk_b * P.b_blk_bytes;
  const int tail_pad = (blocks_m - nblk_a_max) * P.a_blk_bytes;
  int stages = (kWSmemBudget - kWCtrlBytes - tail_pad) / P.stage_bytes;
  if (stages > kWMaxStages) stages = kWMaxStages;
  if (stages < 2) return fail(PMFB_ERR_INVALID, "wgrad halo: stage of %d bytes does not fit twice", P.stage_bytes);
  P.stages = stages;
  int cols = 32;
  while (cols < P.gps * n_tile) cols <<= 1;
  P.tmem_cols = cols;
  P.dw = d->dw;

  CUtensorMap tmx, tmdy;
  uint32_t boxx[5] = {(uint32_t)P.kslab, (uint32_t)pitch, 1, (uint32_t)rows, 1};
  int rc = h16 ? make_tmap_16(&tmx, d->x.ptr, 5, d->x.dims, d->x.strides, boxx, true)
               : make_tmap_f32(&tmx, d->x.ptr, 5, d->x.dims, d->x.strides, boxx, true);
  if (rc) return rc;
  uint32_t boxy[5] = {(uint32_t)P.kslab, 8, 1, 8, 1};
  rc = h16 ? make_tmap_16(&tmdy, d->dy.ptr, 5, d->dy.dims, d->dy.strides, boxy, true)
           : make_tmap_f32(&tmdy, d->dy.ptr, 5, d->dy.dims, d->dy.strides, boxy, true);
  if (rc) return rc;

  const size_t smem = (size_t)kWCtrlBytes + (size_t)stages * P.stage_bytes + tail_pad + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    PMFB_CUDA_CHECK(cudaFuncSetAttribute(conv_wgrad_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kWSmemBudget + 2048));
    attr_set = true;
  }
  conv_wgrad_halo_kernel<<<groups * split, kWThreads, smem, (cudaStream_t)stream>>>(tmx, tmdy, P);
  PMFB_LAUNCH_CHECK("conv_wgrad_halo_kernel");
  return PMFB_OK;
}

}  // namespace pmfb
