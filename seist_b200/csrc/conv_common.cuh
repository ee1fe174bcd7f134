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




// Two-phase staging of linearly up-sampled rows (reference F.interpolate(mode="linear"), models/seist.py:566):
// phase 1 evaluates BN/GELU once per SOURCE sample into `src_s`, phase 2 interpolates from shared memory, so
// the activation is not re-evaluated for both neighbours of every up-sampled sample.  Must be called by all
// threads of the CTA (contains a barrier).  dst[r*pitch + pos] <-> conv-input coordinate p_base + pos.
__device__ __forceinline__ void stage_upsampled_rows(const SeistOp& op, int n, int ci0, int nrows, float* dst, int pitch,
                                                     int width, int p_base, float* src_s, int spitch, int Lsrc,
                                                     float ratio) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int p_lo = max(p_base, 0), p_hi = min(p_base + width - 1, op.L_in - 1);
  int i_lo = 0, count = 0;
  if (p_hi >= p_lo) {
    int a0, a1, b0, b1;
    float lam;
    upsample_coords(p_lo, ratio, Lsrc, a0, a1, lam);
    upsample_coords(p_hi, ratio, Lsrc, b0, b1, lam);
    i_lo = a0;
    count = min(b1 - a0 + 1, spitch);
  }
  for (int r = warp; r < nrows; r += nwarps) {
    const RowSrc rs = make_row(op, n, ci0 + r);
    for (int i = lane; i < count; i += 32) src_s[r * spitch + i] = row_u(rs, i_lo + i);
  }
  __syncthreads();
  for (int r = warp; r < nrows; r += nwarps) {
    const float* sr = src_s + r * spitch;
    float* d = dst + r * pitch;
    for (int pos = lane; pos < width; pos += 32) {
      const int p = p_base + pos;
      float v = 0.f;
      if (p >= 0 && p < op.L_in) {
        int i0, i1;
        float lam;
        upsample_coords(p, ratio, Lsrc, i0, i1, lam);
        v = (1.f - lam) * sr[i0 - i_lo] + lam * sr[i1 - i_lo];
      }
      d[pos] = v;
    }
  }
}

}  // namespace seist
