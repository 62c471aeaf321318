// The fused per-element functional shared by the conv epilogue and the standalone pointwise kernel
// (pmfb_epilogue in include/pmfb.h):
//   v = alpha1[c]*v + beta1[c];  v += r1;  v = act(v);  v = alpha2[c]*v + beta2[c];  v *= mul;  v += r2
// It expresses, in one pass, every elementwise pattern around the reference's convolutions:
//   SalsaNext / fusion conv->LeakyReLU->BN(+shortcut)   (salsanext.py:27-36, pmf_net.py:13-18)
//   ResNet conv->BN(+identity)->ReLU                     (torchvision BasicBlock)
//   attention gate sigmoid(BN(conv)) * fuse + pcd        (pmf_net.py:20-35)
#pragma once
#include "common.h"
#include "ptx.cuh"

namespace pmfb {

struct EpiView {
  const float* p;
  long long sn, sy, sx;
};

struct EpiParams {
  const float* alpha1;
  const float* beta1;
  const float* alpha2;
  const float* beta2;
  EpiView r1, mul, r2;
  int act;
  int round_out;
};

struct EpiPixel {
  const float* r1;
  const float* mul;
  const float* r2;
};

static inline int epi_from_c(const pmfb_epilogue* e, EpiParams* o) {
  o->alpha1 = e->alpha1;
  o->beta1 = e->beta1;
  o->alpha2 = e->alpha2;
  o->beta2 = e->beta2;
  const pmfb_view* vs[3] = {&e->r1, &e->mul, &e->r2};
  EpiView* os[3] = {&o->r1, &o->mul, &o->r2};
  for (int i = 0; i < 3; ++i) {
    os[i]->p = vs[i]->ptr;
    os[i]->sn = vs[i]->sn;
    os[i]->sy = vs[i]->sy;
    os[i]->sx = vs[i]->sx;
    if (vs[i]->ptr && (((vs[i]->sn | vs[i]->sy | vs[i]->sx) % 4) || (reinterpret_cast<uintptr_t>(vs[i]->ptr) & 15)))
      return fail(PMFB_ERR_INVALID, "epilogue operand %d must be 16-byte aligned with strides multiple of 4", i);
  }
  if (e->act < 0 || e->act > 3) return fail(PMFB_ERR_INVALID, "act=%d", e->act);
  o->act = e->act;
  o->round_out = e->round_out;
  return PMFB_OK;
}

__device__ __forceinline__ EpiPixel epi_pixel(const EpiParams& E, int n, int y, int x) {
  EpiPixel p;
  p.r1 = E.r1.p ? E.r1.p + (long long)n * E.r1.sn + (long long)y * E.r1.sy + (long long)x * E.r1.sx : nullptr;
  p.mul = E.mul.p ? E.mul.p + (long long)n * E.mul.sn + (long long)y * E.mul.sy + (long long)x * E.mul.sx : nullptr;
  p.r2 = E.r2.p ? E.r2.p + (long long)n * E.r2.sn + (long long)y * E.r2.sy + (long long)x * E.r2.sx : nullptr;
  return p;
}

__device__ __forceinline__ float epi_act(int act, float v) {
  switch (act) {
    case PMFB_ACT_RELU: return v > 0.f ? v : 0.f;
    case PMFB_ACT_LEAKY: return v > 0.f ? v : 0.01f * v;
    case PMFB_ACT_SIGMOID: return 1.f / (1.f + expf(-v));
    default: return v;
  }
}

__device__ __forceinline__ float4 ld4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

__device__ __forceinline__ float4 epi_apply4(const EpiParams& E, const EpiPixel& px, int c, float4 v) {
  if (E.alpha1) {
    const float4 a = ld4(E.alpha1 + c);
    v.x *= a.x; v.y *= a.y; v.z *= a.z; v.w *= a.w;
  }
  if (E.beta1) {
    const float4 b = ld4(E.beta1 + c);
    v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
  }
  if (px.r1) {
    const float4 r = ld4(px.r1 + c);
    v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
  }
  if (E.act) {
    v.x = epi_act(E.act, v.x); v.y = epi_act(E.act, v.y);
    v.z = epi_act(E.act, v.z); v.w = epi_act(E.act, v.w);
  }
  if (E.alpha2) {
    const float4 a = ld4(E.alpha2 + c);
    v.x *= a.x; v.y *= a.y; v.z *= a.z; v.w *= a.w;
  }
  if (E.beta2) {
    const float4 b = ld4(E.beta2 + c);
    v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
  }
  if (px.mul) {
    const float4 m = ld4(px.mul + c);
    v.x *= m.x; v.y *= m.y; v.z *= m.z; v.w *= m.w;
  }
  if (px.r2) {
    const float4 r = ld4(px.r2 + c);
    v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
  }
  if (E.round_out) {
    v.x = round_tf32(v.x); v.y = round_tf32(v.y); v.z = round_tf32(v.z); v.w = round_tf32(v.w);
  }
  return v;
}

}  // namespace pmfb
