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




// BN-apply / GELU of a view on the row elements lane, lane + 32, ... inside [lo, hi), in place in shared memory, two
// elements per packed instruction
__device__ __forceinline__ void view_inplace(float* d, int lane, int lo, int hi, const RowSrc& rs) {
  const float2 sc = dup2(rs.sc), sh = dup2(rs.sh);
  for (int pos = lane; pos < hi; pos += 64) {
    const int p1 = pos + 32;
    const bool ok0 = pos >= lo, ok1 = p1 >= lo && p1 < hi;
    float2 v = make_float2(ok0 ? d[pos] : 0.f, ok1 ? d[p1] : 0.f);
    v = fma2(sc, v, sh);
    if (rs.act == SEIST_ACT_GELU) v = gelu2(v);
    if (ok0) d[pos] = v.x;
    if (ok1) d[p1] = v.y;
  }
}

// ------------------------------------------------------------------------------------------------
// Asynchronous staging of conv-input rows.  dst[r*pitch + pos] <-> conv-input coordinate p_base + pos of channel
// ci0 + r.  Element ownership (row = warp + k*nwarps, pos = lane + 32*u) is the same in the issue and the transform
// function, so a thread only ever touches elements it copied itself.
// ------------------------------------------------------------------------------------------------
// raw copies of the valid range, zeros in the padding and in rows >= nrows (up to nrows_pad)
__device__ __forceinline__ void rows_issue_plain(const SeistOp& op, int n, int ci0, int nrows, int nrows_pad, float* dst,
                                                 int pitch, int width, int p_base) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int r = warp; r < nrows_pad; r += nwarps) {
    float* d = dst + r * pitch;
    if (r >= nrows) {
      for (int pos = lane; pos < width; pos += 32) d[pos] = 0.f;
      continue;
    }
    int cv;
    const int vi = resolve_view(op, ci0 + r, cv);
    const float* xr = view_row(op.in[vi], n, cv);
    const uint32_t da = smem_addr(d);
    for (int pos = lane; pos < width; pos += 32) {
      const int p = p_base + pos;
      if (p >= 0 && p < op.L_in) cp_async4(da + 4 * pos, xr + p);
      else d[pos] = 0.f;
    }
  }
}
// BN-apply / GELU of the consumer view, in place, after cp_async_wait<0>()
__device__ __forceinline__ void rows_transform_plain(const SeistOp& op, int n, int ci0, int nrows, float* dst, int pitch,
                                                     int width, int p_base) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int pos_lo = max(0, -p_base), pos_hi = min(width, op.L_in - p_base);   // valid positions [pos_lo, pos_hi)
  for (int r = warp; r < nrows; r += nwarps) {
    const RowSrc rs = make_row(op, n, ci0 + r);
    if (rs.act == SEIST_ACT_NONE && rs.sc == 1.f && rs.sh == 0.f) continue;
    float* d = dst + r * pitch;
    view_inplace(d, lane, pos_lo, pos_hi, rs);
  }
}

// source window of the linearly up-sampled rows: conv-input coordinates [p_base, p_base + width) read the source
// samples [i_lo, i_lo + count)
__device__ __forceinline__ void upsample_window(const SeistOp& op, int p_base, int width, int spitch, int Lsrc, float ratio,
                                                int& i_lo, int& count) {
  const int p_lo = max(p_base, 0), p_hi = min(p_base + width - 1, op.L_in - 1);
  i_lo = 0;
  count = 0;
  if (p_hi >= p_lo) {
    int a0, a1, b0, b1;
    float lam;
    upsample_coords(p_lo, ratio, Lsrc, a0, a1, lam);
    upsample_coords(p_hi, ratio, Lsrc, b0, b1, lam);
    i_lo = a0;
    count = min(b1 - a0 + 1, spitch);
  }
}
__device__ __forceinline__ void src_issue(const SeistOp& op, int n, int ci0, int nrows, float* src_s, int spitch, int i_lo,
                                          int count) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int r = warp; r < nrows; r += nwarps) {
    int cv;
    const int vi = resolve_view(op, ci0 + r, cv);
    const float* xr = view_row(op.in[vi], n, cv) + i_lo;
    const uint32_t da = smem_addr(src_s + r * spitch);
    for (int i = lane; i < count; i += 32) cp_async4(da + 4 * i, xr + i);
  }
}
__device__ __forceinline__ void src_transform(const SeistOp& op, int n, int ci0, int nrows, float* src_s, int spitch, int count) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int r = warp; r < nrows; r += nwarps) {
    const RowSrc rs = make_row(op, n, ci0 + r);
    float* d = src_s + r * spitch;
    if (rs.act != SEIST_ACT_NONE || rs.sc != 1.f || rs.sh != 0.f) view_inplace(d, lane, 0, count, rs);
  }
}
__device__ __forceinline__ void rows_interpolate(const SeistOp& op, int nrows, float* dst, int pitch, int width, int p_base,
                                                 const float* src_s, int spitch, int i_lo, int Lsrc, float ratio) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
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

// ------------------------------------------------------------------------------------------------
// Asynchronous staging of the output-gradient ("gacc") rows of the weight-gradient kernels.
// g_s holds channel-PAIR rows: the 8 floats at g_s[pr*gpitch + 8*q] are {g[2pr][4q..4q+3] interleaved with
// g[2pr+1][4q..4q+3]}.  The raw operands of (pr, q) land in exactly those 32 bytes (row 2pr first, then row 2pr+1) and, as
// far as the op needs them, at the same offset of the planes x_s / d_s; the owning thread (idx -> (pr, q), the same
// mapping in both functions) later combines them in place: BN backward, sigmoid', drop factors, pair interleave.
//   plane 0 (g_s): du when the output carries a BatchNorm gradient, else dxd (zeros if absent)
//   x_s: the forward output (BN backward / sigmoid');  d_s: dxd next to a BatchNorm gradient
// ------------------------------------------------------------------------------------------------
template <int NT>
__device__ __forceinline__ void gacc_issue(const SeistOp& op, int n, int l0, int L, int co_base, int co_hi, int CO_B, int QPR,
                                           int gpitch, float* g_s, float* x_s, float* d_s, bool has_bn, bool need_x) {
  const float* src0 = has_bn ? op.out.g : op.out_dxd;
  const float* src2 = has_bn ? op.out_dxd : nullptr;
  for (int idx = threadIdx.x; idx < (CO_B / 2) * QPR; idx += NT) {
    const int pr = idx / QPR, q = idx - pr * QPR;
    const int l = l0 + 4 * q;
    const int o8 = pr * gpitch + 8 * q;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int co = co_base + 2 * pr + h;
      float* d = g_s + o8 + 4 * h;
      if (co < co_hi && l < L) {
        const size_t off = ((size_t)n * op.out.Ct + op.out.c0 + co) * (size_t)L + l;
        if (src0) cp_async16(smem_addr(d), src0 + off);
        else *reinterpret_cast<float4*>(d) = make_float4(0.f, 0.f, 0.f, 0.f);
        if (need_x) cp_async16(smem_addr(x_s + o8 + 4 * h), op.out.x + off);
        if (src2) cp_async16(smem_addr(d_s + o8 + 4 * h), src2 + off);
      } else {
        *reinterpret_cast<float4*>(d) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
  }
}
template <int NT, typename OC>
__device__ __forceinline__ void gacc_combine(const SeistOp& op, int n, int l0, int L, int co_base, int co_hi, int CO_B, int QPR,
                                             int gpitch, float* g_s, const float* x_s, const float* d_s, const OC* oc_s,
                                             bool has_bn, bool need_x, float pf, uint64_t seed, int Cout_all) {
  const bool has_d = has_bn && op.out_dxd != nullptr;
  for (int idx = threadIdx.x; idx < (CO_B / 2) * QPR; idx += NT) {
    const int pr = idx / QPR, q = idx - pr * QPR;
    const int l = l0 + 4 * q;
    const int o8 = pr * gpitch + 8 * q;
    float4 gh[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int row = 2 * pr + h, co = co_base + row;
      float4 gv = *reinterpret_cast<const float4*>(g_s + o8 + 4 * h);
      if (co < co_hi && l < L) {
        if (need_x) {
          const float4 x = *reinterpret_cast<const float4*>(x_s + o8 + 4 * h);
          if (has_bn) {
            const OC o = oc_s[row];
            gv = fma4(splat4(o.A), gv, fma4(splat4(o.Bx), x, splat4(o.Cc)));
            if (has_d) gv = add4(gv, *reinterpret_cast<const float4*>(d_s + o8 + 4 * h));
          }
          if (op.out_act == SEIST_OUT_SIGMOID) gv = mul4(gv, mul4(x, fma4(x, splat4(-1.f), splat4(1.f))));
        }
        gv = scale4(gv, pf);
        if (op.p_elem > 0.f) gv = mul4(gv, keep4(op.p_elem, seed, op.seed_elem, ((uint64_t)n * Cout_all + co) * (uint64_t)L + l));
      }
      gh[h] = gv;
    }
    *reinterpret_cast<float4*>(g_s + o8) = make_float4(gh[0].x, gh[1].x, gh[0].y, gh[1].y);
    *reinterpret_cast<float4*>(g_s + o8 + 4) = make_float4(gh[0].z, gh[1].z, gh[0].w, gh[1].w);
  }
}

// Two-phase staging of linearly up-sampled rows (reference F.interpolate(mode="linear"), models/seist.py:566):
// phase 1 brings the SOURCE samples in (asynchronous copies, all in flight at once) and evaluates BN/GELU once per
// source sample in place, phase 2 interpolates from shared memory, so the activation is not re-evaluated for both
// neighbours of every up-sampled sample.  Must be called by all threads of the CTA (contains a barrier).
__device__ __forceinline__ void stage_upsampled_rows(const SeistOp& op, int n, int ci0, int nrows, float* dst, int pitch,
                                                     int width, int p_base, float* src_s, int spitch, int Lsrc,
                                                     float ratio) {
  int i_lo, count;
  upsample_window(op, p_base, width, spitch, Lsrc, ratio, i_lo, count);
  src_issue(op, n, ci0, nrows, src_s, spitch, i_lo, count);
  cp_async_commit();
  cp_async_wait<0>();
  src_transform(op, n, ci0, nrows, src_s, spitch, count);
  __syncthreads();
  rows_interpolate(op, nrows, dst, pitch, width, p_base, src_s, spitch, i_lo, Lsrc, ratio);
}

}  // namespace seist
