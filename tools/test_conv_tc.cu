// Standalone bring-up test for the tcgen05 conv kernels: calls the C-ABI of libpmf_b200.so and checks
// against a double-precision CPU loop (a restatement of nn.Conv2d's cross-correlation with zero padding).
// Build: see tools/Makefile.  Run on the GPU box: tools/test_conv_tc [filter]
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "../include/pmfb.h"

static uint32_t g_seed = 12345;
static float frand() {
  g_seed = g_seed * 1664525u + 1013904223u;
  return ((g_seed >> 8) & 0xFFFF) / 65536.0f - 0.5f;
}
static float tf32_round(float x) {
  uint32_t u;
  memcpy(&u, &x, 4);
  u += 0x1000u;  // RN (ties away), matches cvt.rna.tf32
  u &= 0xFFFFE000u;
  memcpy(&x, &u, 4);
  return x;
}
#define CK(x)                                                                  \
  do {                                                                         \
    cudaError_t e = (x);                                                       \
    if (e != cudaSuccess) {                                                    \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
      exit(2);                                                                 \
    }                                                                          \
  } while (0)

struct Case {
  const char* name;
  int N, H, W, Cin, Cout, kh, kw, dil, pad, stride;
  int tile_w, tile_h, n_tile;
  int epi;  // 0 none, 1 leaky+affine2+r2, 2 affine1+r1+relu, 3 sigmoid*mul+r2, 4 round_out
};

static float act_ref(int act, float v) {
  if (act == 1) return v > 0 ? v : 0;
  if (act == 2) return v > 0 ? v : 0.01f * v;
  if (act == 3) return 1.f / (1.f + expf(-v));
  return v;
}

static int run_fwd(const Case& c) {
  const int Ho = (c.H + 2 * c.pad - c.dil * (c.kh - 1) - 1) / c.stride + 1;
  const int Wo = (c.W + 2 * c.pad - c.dil * (c.kw - 1) - 1) / c.stride + 1;
  const int T = c.kh * c.kw;
  std::vector<float> x((size_t)c.N * c.H * c.W * c.Cin), w((size_t)T * c.Cout * c.Cin);
  for (auto& v : x) v = tf32_round(frand());
  for (auto& v : w) v = tf32_round(frand() * 0.25f);
  const size_t on = (size_t)c.N * Ho * Wo * c.Cout;
  std::vector<float> a1(c.Cout), b1(c.Cout), a2(c.Cout), b2(c.Cout), r1(on), mul(on), r2(on);
  for (int i = 0; i < c.Cout; ++i) {
    a1[i] = 1.f + frand();
    b1[i] = frand();
    a2[i] = 1.f + frand();
    b2[i] = frand();
  }
  for (size_t i = 0; i < on; ++i) {
    r1[i] = frand();
    mul[i] = frand();
    r2[i] = frand();
  }

  float *dx, *dw, *dout, *da1, *db1, *da2, *db2, *dr1, *dmul, *dr2;
  CK(cudaMalloc(&dx, x.size() * 4));
  CK(cudaMalloc(&dw, w.size() * 4));
  CK(cudaMalloc(&dout, on * 4));
  CK(cudaMalloc(&da1, c.Cout * 4));
  CK(cudaMalloc(&db1, c.Cout * 4));
  CK(cudaMalloc(&da2, c.Cout * 4));
  CK(cudaMalloc(&db2, c.Cout * 4));
  CK(cudaMalloc(&dr1, on * 4));
  CK(cudaMalloc(&dmul, on * 4));
  CK(cudaMalloc(&dr2, on * 4));
  CK(cudaMemcpy(dx, x.data(), x.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dw, w.data(), w.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(da1, a1.data(), c.Cout * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(db1, b1.data(), c.Cout * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(da2, a2.data(), c.Cout * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(db2, b2.data(), c.Cout * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dr1, r1.data(), on * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dmul, mul.data(), on * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dr2, r2.data(), on * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(dout, 0xFF, on * 4));

  pmfb_conv_desc d;
  memset(&d, 0, sizeof(d));
  d.x.ptr = dx;
  if (c.stride == 1) {
    d.x.dims[0] = c.Cin; d.x.dims[1] = c.W; d.x.dims[2] = 1; d.x.dims[3] = c.H; d.x.dims[4] = c.N;
    d.x.strides[0] = (uint64_t)c.Cin * 4;
    d.x.strides[1] = (uint64_t)c.W * c.Cin * 4;  // parity dim (size 1): any legal stride
    d.x.strides[2] = (uint64_t)c.W * c.Cin * 4;
    d.x.strides[3] = (uint64_t)c.H * c.W * c.Cin * 4;
  } else {
    d.x.dims[0] = 2 * c.Cin; d.x.dims[1] = c.W / 2; d.x.dims[2] = 2; d.x.dims[3] = c.H / 2; d.x.dims[4] = c.N;
    d.x.strides[0] = (uint64_t)2 * c.Cin * 4;
    d.x.strides[1] = (uint64_t)c.W * c.Cin * 4;
    d.x.strides[2] = (uint64_t)2 * c.W * c.Cin * 4;
    d.x.strides[3] = (uint64_t)c.H * c.W * c.Cin * 4;
  }
  d.w = dw;
  d.c_in = c.Cin; d.c_out = c.Cout; d.n_taps = T;
  std::vector<int> tdh(T), tdw(T);
  for (int i = 0; i < c.kh; ++i)
    for (int j = 0; j < c.kw; ++j) {
      const int t = i * c.kw + j;
      const int dh = i * c.dil - c.pad, dwv = j * c.dil - c.pad;
      tdh[t] = dh; tdw[t] = dwv;
      if (c.stride == 1) {
        d.tap_dc[t] = 0; d.tap_dw[t] = dwv; d.tap_dp[t] = 0; d.tap_dh[t] = dh;
      } else {
        const int ph = ((dh % 2) + 2) % 2, pw = ((dwv % 2) + 2) % 2;
        d.tap_dc[t] = pw * c.Cin; d.tap_dw[t] = (dwv - pw) / 2; d.tap_dp[t] = ph; d.tap_dh[t] = (dh - ph) / 2;
      }
    }
  d.n_batch = c.N; d.out_h = Ho; d.out_w = Wo;
  d.tile_w = c.tile_w; d.tile_h = c.tile_h; d.n_tile = c.n_tile;
  d.out = dout;
  d.o_sx = c.Cout; d.o_sy = (int64_t)Wo * c.Cout; d.o_sn = (int64_t)Ho * Wo * c.Cout;
  pmfb_view ov;
  ov.sn = d.o_sn; ov.sy = d.o_sy; ov.sx = d.o_sx;
  int act = 0;
  if (c.epi == 1) {
    act = 2; d.epi.beta1 = db1; d.epi.alpha2 = da2; d.epi.beta2 = db2; ov.ptr = dr2; d.epi.r2 = ov;
  } else if (c.epi == 2) {
    act = 1; d.epi.alpha1 = da1; d.epi.beta1 = db1; ov.ptr = dr1; d.epi.r1 = ov;
  } else if (c.epi == 3) {
    act = 3; d.epi.alpha1 = da1; d.epi.beta1 = db1; ov.ptr = dmul; d.epi.mul = ov; ov.ptr = dr2; d.epi.r2 = ov;
  } else if (c.epi == 4) {
    d.epi.round_out = 1;
  }
  d.epi.act = act;

  int rc = pmfb_conv_fwd(&d, 0);
  if (rc) {
    printf("[%s] pmfb_conv_fwd rc=%d: %s\n", c.name, rc, pmfb_last_error());
    return 1;
  }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("[%s] kernel error: %s\n", c.name, cudaGetErrorString(e));
    return 2;
  }
  std::vector<float> out(on);
  CK(cudaMemcpy(out.data(), dout, on * 4, cudaMemcpyDeviceToHost));

  double maxerr = 0, maxref = 0;
  long bad = 0;
  int shown = 0;
  for (int n = 0; n < c.N; ++n)
    for (int oy = 0; oy < Ho; ++oy)
      for (int ox = 0; ox < Wo; ++ox)
        for (int co = 0; co < c.Cout; ++co) {
          double acc = 0;
          for (int t = 0; t < T; ++t) {
            const int iy = oy * c.stride + tdh[t], ix = ox * c.stride + tdw[t];
            if (iy < 0 || iy >= c.H || ix < 0 || ix >= c.W) continue;
            const float* xp = &x[(((size_t)n * c.H + iy) * c.W + ix) * c.Cin];
            const float* wp = &w[((size_t)t * c.Cout + co) * c.Cin];
            for (int ci = 0; ci < c.Cin; ++ci) acc += (double)xp[ci] * wp[ci];
          }
          const size_t oi = (((size_t)n * Ho + oy) * Wo + ox) * c.Cout + co;
          float v = (float)acc;
          if (c.epi == 1) { v = v + b1[co]; v = act_ref(2, v); v = a2[co] * v + b2[co]; v += r2[oi]; }
          if (c.epi == 2) { v = a1[co] * v + b1[co]; v += r1[oi]; v = act_ref(1, v); }
          if (c.epi == 3) { v = a1[co] * v + b1[co]; v = act_ref(3, v); v *= mul[oi]; v += r2[oi]; }
          if (c.epi == 4) v = tf32_round(v);
          const double err = fabs((double)out[oi] - v);
          if (fabs(v) > maxref) maxref = fabs(v);
          if (!(err <= 2e-3 * (1.0 + fabs(v))) ) {
            ++bad;
            if (shown < 6) {
              printf("   mismatch n=%d y=%d x=%d co=%d got=%g ref=%g\n", n, oy, ox, co, out[oi], v);
              ++shown;
            }
          }
          if (err > maxerr) maxerr = err;
        }
  printf("[%s] %s  Ho=%d Wo=%d maxerr=%.3g maxref=%.3g bad=%ld/%zu\n", c.name, bad ? "FAIL" : "ok", Ho, Wo,
         maxerr, maxref, bad, on);
  cudaFree(dx); cudaFree(dw); cudaFree(dout); cudaFree(da1); cudaFree(db1); cudaFree(da2); cudaFree(db2);
  cudaFree(dr1); cudaFree(dmul); cudaFree(dr2);
  return bad ? 1 : 0;
}

struct WCase {
  const char* name;
  int N, H, W, Cin, Cout, kh, kw, dil, pad;
  int ptile_w, ptile_h, n_tile, ksplit;
};

static int run_wgrad(const WCase& c) {
  const int Ho = c.H + 2 * c.pad - c.dil * (c.kh - 1);
  const int Wo = c.W + 2 * c.pad - c.dil * (c.kw - 1);
  const int T = c.kh * c.kw;
  std::vector<float> x((size_t)c.N * c.H * c.W * c.Cin), dy((size_t)c.N * Ho * Wo * c.Cout);
  for (auto& v : x) v = tf32_round(frand());
  for (auto& v : dy) v = tf32_round(frand());
  const size_t wn = (size_t)T * c.Cin * c.Cout;
  float *dx, *ddy, *ddw;
  CK(cudaMalloc(&dx, x.size() * 4));
  CK(cudaMalloc(&ddy, dy.size() * 4));
  CK(cudaMalloc(&ddw, wn * 4));
  CK(cudaMemcpy(dx, x.data(), x.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(ddy, dy.data(), dy.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(ddw, 0, wn * 4));
  pmfb_wgrad_desc d;
  memset(&d, 0, sizeof(d));
  d.x.ptr = dx;
  d.x.dims[0] = c.Cin; d.x.dims[1] = c.W; d.x.dims[2] = 1; d.x.dims[3] = c.H; d.x.dims[4] = c.N;
  d.x.strides[0] = (uint64_t)c.Cin * 4; d.x.strides[1] = (uint64_t)c.W * c.Cin * 4;
  d.x.strides[2] = (uint64_t)c.W * c.Cin * 4; d.x.strides[3] = (uint64_t)c.H * c.W * c.Cin * 4;
  d.dy.ptr = ddy;
  d.dy.dims[0] = c.Cout; d.dy.dims[1] = Wo; d.dy.dims[2] = 1; d.dy.dims[3] = Ho; d.dy.dims[4] = c.N;
  d.dy.strides[0] = (uint64_t)c.Cout * 4; d.dy.strides[1] = (uint64_t)Wo * c.Cout * 4;
  d.dy.strides[2] = (uint64_t)Wo * c.Cout * 4; d.dy.strides[3] = (uint64_t)Ho * Wo * c.Cout * 4;
  d.c_in = c.Cin; d.c_out = c.Cout; d.n_taps = T;
  std::vector<int> tdh(T), tdw(T);
  for (int i = 0; i < c.kh; ++i)
    for (int j = 0; j < c.kw; ++j) {
      const int t = i * c.kw + j;
      tdh[t] = i * c.dil - c.pad; tdw[t] = j * c.dil - c.pad;
      d.tap_dh[t] = tdh[t]; d.tap_dw[t] = tdw[t];
    }
  d.n_batch = c.N; d.out_h = Ho; d.out_w = Wo;
  d.ptile_w = c.ptile_w; d.ptile_h = c.ptile_h; d.n_tile = c.n_tile; d.ksplit = c.ksplit;
  d.dw = ddw;
  int rc = pmfb_conv_wgrad(&d, 0);
  if (rc) {
    printf("[%s] pmfb_conv_wgrad rc=%d: %s\n", c.name, rc, pmfb_last_error());
    return 1;
  }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("[%s] kernel error: %s\n", c.name, cudaGetErrorString(e));
    return 2;
  }
  std::vector<float> got(wn);
  CK(cudaMemcpy(got.data(), ddw, wn * 4, cudaMemcpyDeviceToHost));
  std::vector<double> ref(wn, 0.0);
  for (int n = 0; n < c.N; ++n)
    for (int oy = 0; oy < Ho; ++oy)
      for (int ox = 0; ox < Wo; ++ox) {
        const float* dyp = &dy[(((size_t)n * Ho + oy) * Wo + ox) * c.Cout];
        for (int t = 0; t < T; ++t) {
          const int iy = oy + tdh[t], ix = ox + tdw[t];
          if (iy < 0 || iy >= c.H || ix < 0 || ix >= c.W) continue;
          const float* xp = &x[(((size_t)n * c.H + iy) * c.W + ix) * c.Cin];
          for (int ci = 0; ci < c.Cin; ++ci) {
            double* rp = &ref[((size_t)t * c.Cin + ci) * c.Cout];
            const double xv = xp[ci];
            for (int co = 0; co < c.Cout; ++co) rp[co] += xv * dyp[co];
          }
        }
      }
  double maxerr = 0, maxref = 0;
  long bad = 0;
  int shown = 0;
  for (size_t i = 0; i < wn; ++i) {
    const double err = fabs(got[i] - ref[i]);
    if (fabs(ref[i]) > maxref) maxref = fabs(ref[i]);
    if (err > maxerr) maxerr = err;
    if (!(err <= 2e-3 * (1.0 + fabs(ref[i])))) {
      ++bad;
      if (shown < 6) {
        const int co = i % c.Cout, ci = (i / c.Cout) % c.Cin, t = i / ((size_t)c.Cout * c.Cin);
        printf("   mismatch tap=%d ci=%d co=%d got=%g ref=%g\n", t, ci, co, got[i], ref[i]);
        ++shown;
      }
    }
  }
  printf("[%s] %s maxerr=%.3g maxref=%.3g bad=%ld/%zu\n", c.name, bad ? "FAIL" : "ok", maxerr, maxref, bad, wn);
  cudaFree(dx); cudaFree(ddy); cudaFree(ddw);
  return bad ? 1 : 0;
}

int main(int argc, char** argv) {
  const char* filt = argc > 1 ? argv[1] : "";
  const bool list_only = argc > 1 && strcmp(argv[1], "--list") == 0;
  int rc = list_only ? 0 : pmfb_init();
  if (!list_only) printf("pmfb_init rc=%d %s\n", rc, rc ? pmfb_last_error() : "");
  if (rc) return 3;
  const Case cases[] = {
      // name                N  H   W   Cin Cout kh kw dil pad s  tw  th  nt  epi
      {"f_1x1_c32_n128",     1, 8,  16, 32, 128, 1, 1, 1,  0,  1, 16, 8,  128, 0},
      {"f_1x1_c128_n64",     1, 8,  16, 128, 64, 1, 1, 1,  0,  1, 16, 8,  64, 0},
      {"f_1x1_c256_n256",    2, 8,  32, 256, 256, 1, 1, 1, 0,  1, 32, 4,  256, 0},
      {"f_3x3_c64_n64",      2, 16, 32, 64, 64, 3, 3, 1,  1,  1, 32, 4,  64, 0},
      {"f_3x3d2_c32_n32",    1, 16, 128, 32, 32, 3, 3, 2, 2,  1, 128, 1, 32, 1},
      {"f_2x2d2_c64_n64",    1, 16, 64, 64, 64, 2, 2, 2,  1,  1, 64, 2,  64, 2},
      {"f_3x3_c80_n32",      1, 16, 32, 80, 32, 3, 3, 1,  1,  1, 16, 8,  32, 3},
      {"f_1x1_c32_n20",      1, 16, 32, 32, 20, 1, 1, 1,  0,  1, 16, 8,  32, 4},
      {"f_3x3_odd_40x30",    2, 30, 40, 64, 128, 3, 3, 1, 1,  1, 8,  16, 128, 1},
      {"f_3x3d6_c64",        1, 8,  40, 64, 64, 3, 3, 6,  6,  1, 8,  16, 64, 0},
      {"f_3x3s2_c64_n128",   2, 16, 32, 64, 128, 3, 3, 1, 1,  2, 16, 8,  128, 0},
      {"f_1x1s2_c64_n128",   1, 16, 32, 64, 128, 1, 1, 1, 0,  2, 16, 8,  128, 0},
      {"f_3x3_c512_n256",    1, 8,  32, 512, 512, 3, 3, 1, 1, 1, 32, 4,  256, 0},
      {"f_3x3_c16_n20",      1, 16, 32, 16, 20, 3, 3, 1,  1,  1, 32, 4,  32, 0},
  };
  const WCase wcases[] = {
      // name              N  H   W   Cin Cout kh kw dil pad pw ph nt  ks
      {"w_1x1_c32_o32",    1, 8,  16, 32, 32,  1, 1, 1,  0,  16, 2, 32, 1},
      {"w_1x1_c128_o128",  1, 8,  16, 128, 128, 1, 1, 1, 0,  16, 2, 128, 1},
      {"w_3x3_c64_o64",    2, 16, 32, 64, 64,  3, 3, 1,  1,  32, 1, 64, 3},
      {"w_3x3d2_c32_o32",  1, 16, 64, 32, 32,  3, 3, 2,  2,  8,  4, 32, 2},
      {"w_2x2d2_c64_o128", 1, 16, 32, 64, 128, 2, 2, 2,  1,  16, 2, 128, 4},
      {"w_3x3_c80_o32",    1, 16, 32, 80, 32,  3, 3, 1,  1,  32, 1, 32, 2},
      {"w_3x3_odd_o256",   2, 30, 40, 128, 256, 3, 3, 1, 1,  8,  4, 256, 5},
      {"w_1x1_c32_o20",    1, 16, 32, 32, 20,  1, 1, 1,  0,  32, 1, 32, 2},
      {"w_3x3_c16_o20",    1, 16, 32, 16, 20,  3, 3, 1,  1,  32, 1, 32, 1},
  };
  if (list_only) {
    for (const Case& c : cases) printf("%s\n", c.name);
    for (const WCase& c : wcases) printf("%s\n", c.name);
    return 0;
  }
  int fails = 0, ran = 0;
  for (const Case& c : cases) {
    if (strstr(c.name, filt) == nullptr) continue;
    ++ran;
    int r = run_fwd(c);
    if (r == 2) { printf("aborting after sticky CUDA error\n"); return 2; }
    fails += r;
  }
  for (const WCase& c : wcases) {
    if (strstr(c.name, filt) == nullptr) continue;
    ++ran;
    int r = run_wgrad(c);
    if (r == 2) { printf("aborting after sticky CUDA error\n"); return 2; }
    fails += r;
  }
  printf("SUMMARY ran=%d failed=%d\n", ran, fails);
  return fails ? 1 : 0;
}
