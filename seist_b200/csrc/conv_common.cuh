// Helpers shared by the convolution kernels: evaluation of a consumer view at a conv-input coordinate.
#pragma once
#include "common.cuh"

namespace seist {

// ------------------------------------------------------------------------------------------------
// value of the conv input channel row at conv-input coordinate p (after pool / up-sampling, before pad)
// ------------------------------------------------------------------------------------------------
struct RowSrc {
  const float* x;   // channel row of the source view (length Lsrc)
  float sc, sh;
  int act;
};

__device__ __forceinline__ float row_u(const RowSrc& r, int i) {
  float u = fmaf(r.sc, r.x[i], r.sh);
  return r.act == SEIST_ACT_GELU ? gelu_f(u) : u;
}

__device__ __forceinline__ void upsample_coords(int p, float ratio, int Lsrc, int& i0, int& i1, float& lam) {
  float src = ratio * ((float)p + 0.5f) - 0.5f;
  src = src < 0.f ? 0.f : src;
  i0 = (int)src;
  i0 = i0 < Lsrc - 1 ? i0 : Lsrc - 1;
  i1 = i0 < Lsrc - 1 ? i0 + 1 : i0;
  lam = src - (float)i0;
  lam = lam < 0.f ? 0.f : (lam > 1.f ? 1.f : lam);
}

__device__ __forceinline__ float conv_input_at(const SeistOp& op, const RowSrc& r, int p, int Lsrc, float ratio) {
  if (p < 0 || p >= op.L_in) return 0.f;
  if (op.pool > 1) {
    const int s0 = p * op.pool;
    const int cnt = min(op.pool, Lsrc - s0);
    float sum = 0.f, mx = -INFINITY;
    for (int i = 0; i < cnt; ++i) {
      const float u = row_u(r, s0 + i);
      sum += u;
      mx = fmaxf(mx, u);
    }
    return sum / (float)cnt + mx;
  }
  if (op.up_src_L > 0) {
    int i0, i1;
    float lam;
    upsample_coords(p, ratio, Lsrc, i0, i1, lam);
    return (1.f - lam) * row_u(r, i0) + lam * row_u(r, i1);
  }
  return row_u(r, p);
}

__device__ __forceinline__ RowSrc make_row(const SeistOp& op, int n, int ci) {
  int cv;
  const int vi = resolve_view(op, ci, cv);
  const SeistView& v = op.in[vi];
  RowSrc r;
  r.x = view_row(v, n, cv);
  view_coef(op, v, cv, r.sc, r.sh);
  r.act = v.act;
  return r;
}



}  // namespace seist
