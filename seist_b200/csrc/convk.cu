// Dense (groups == 1) k-tap convolutions: the composed stem paths (k = 5..19, stride 1/2, reference
// models/seist.py:124-155) and the up-sampling head (k = 7/11 behind F.interpolate(linear), :536,:566).
//
// Tile: 256 threads = WC channel-warps x WP sample-warps; a thread owns 4 CONSECUTIVE output samples x 8
// output channels.  The (channel, sample) input tile with halo is staged once per 8-channel chunk with
// BN-apply / GELU / linear up-sampling / zero padding evaluated on the way in; per reduction channel a
// thread pulls its sliding window (3*S + K samples) with 16-byte shared loads and reuses every window
// element for up to K taps x 8 channels (K and S are template parameters so the window lives in
// registers).  The same engine run with flipped/transposed weights over the BN-backward-combined output
// gradient is the data gradient (stride 1).  The weight gradient keeps lanes on the sample axis: a warp
// owns a (4 co) x (TCI ci) x K tile of dW in registers and reduces it over lanes once per CTA lifetime.
#include "common.cuh"
#include "conv_common.cuh"

namespace seist {

constexpr int CK_NT = 256;
constexpr int CK_CIC = 8;

__device__ __forceinline__ float4 lds4(const float* p) { return *reinterpret_cast<const float4*>(p); }

template <int K, int S>
struct CkWin {
  static constexpr int WIN = 3 * S + K;
  static constexpr int NV = (WIN + 3) / 4;
};

// acc[c][j] += sum_t w[t][c] * win[j*S + t] for one reduction channel
// channel pairs ride the two lanes of FFMA2; ck_acc reads channel c, sample j of a [4][4] pair tile
__device__ __forceinline__ float ck_acc(const float2 (&acc)[4][4], int c, int j) { return (c & 1) ? acc[c >> 1][j].y : acc[c >> 1][j].x; }

template <int K, int S>
__device__ __forceinline__ void ck_accumulate(const float* irow, const float* wrow, int wstride, float2 (&acc)[4][4]) {
  constexpr int NV = CkWin<K, S>::NV;
  float win[NV * 4];
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const float4 q = lds4(irow + 4 * v);
    win[4 * v] = q.x;
    win[4 * v + 1] = q.y;
    win[4 * v + 2] = q.z;
    win[4 * v + 3] = q.w;
  }
#pragma unroll
  for (int t = 0; t < K; ++t) {
    const float4 w0 = lds4(wrow + t * wstride), w1 = lds4(wrow + t * wstride + 4);
    const float2 w[4] = {make_float2(w0.x, w0.y), make_float2(w0.z, w0.w), make_float2(w1.x, w1.y), make_float2(w1.z, w1.w)};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 v = dup2(win[j * S + t]);
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[c][j] = fma2(w[c], v, acc[c][j]);
    }
  }
}

// stage rows of the consumer view (conv-input coordinates p_base .. p_base + width) — 8 loads in flight
__device__ __forceinline__ void ck_stage_input(const SeistOp& op, int n, int ci0, int cic, float* in_s, int pitch,
                                               int width, int p_base, int Lsrc, float ratio) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool plain = op.up_src_L == 0;
  for (int r = warp; r < CK_CIC; r += CK_NT / 32) {
    float* dst = in_s + r * pitch;
    if (r >= cic) {
      for (int pos = lane; pos < width; pos += 32) dst[pos] = 0.f;
      continue;
    }
    const RowSrc rs = make_row(op, n, ci0 + r);
    if (plain) {
      for (int pos0 = lane; pos0 < width; pos0 += 32 * 8) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int p = p_base + pos0 + 32 * u;
          v[u] = (pos0 + 32 * u < width && p >= 0 && p < op.L_in) ? rs.x[p] : 0.f;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int pos = pos0 + 32 * u, p = p_base + pos;
          if (pos < width) {
            float t = fmaf(rs.sc, v[u], rs.sh);
            if (rs.act == SEIST_ACT_GELU) t = gelu_f(t);
            dst[pos] = (p >= 0 && p < op.L_in) ? t : 0.f;
          }
        }
      }
    } else {
      for (int pos = lane; pos < width; pos += 32) dst[pos] = conv_input_at(op, rs, p_base + pos, Lsrc, ratio);
    }
  }
}

// ================================================================================================
// forward
// ================================================================================================
template <int K, int S>
__global__ void __launch_bounds__(CK_NT, 3) convk_fwd_kernel(const __grid_constant__ SeistOp op, const int WC) {
  extern __shared__ __align__(16) float ck_smem[];
  const int WP = 8 / WC, CO_B = 8 * WC, TLo = 128 * WP;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wc = warp % WC, wp = warp / WC;
  const int gs_in = op.Cin / op.groups, gs_out = op.Cout / op.groups;
  const int tpg = (gs_out + CO_B - 1) / CO_B;          // output-channel tiles per group
  const int grp = blockIdx.z / tpg;
  const int n = blockIdx.y, l0 = blockIdx.x * TLo, co_base = grp * gs_out + (blockIdx.z - grp * tpg) * CO_B;
  const int co_end = (grp + 1) * gs_out;               // channels of this group only
  const int ci_grp = grp * gs_in;
  const int width = TLo * S + K - S;
  const int pitch = ((width + 3) & ~3) + 4;
  float* in_s = ck_smem;                               // [CIC][pitch]
  float* w_s = ck_smem + CK_CIC * pitch;               // [CIC][K][CO_B]
  float* red_s = w_s + CK_CIC * K * CO_B;              // [8 warps][16]
  float* src_s = red_s + 8 * 16;                       // [CIC][width+4] (up-sampled input only)
  const int Lsrc = op.in[0].L;
  const float ratio = op.up_src_L > 0 ? (float)Lsrc / (float)op.L_in : 1.f;
  const int p_base = l0 * S - op.pad_left;

  float2 acc[4][4];
#pragma unroll
  for (int c = 0; c < 4; ++c)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[c][j] = make_float2(0.f, 0.f);

  for (int ci0 = 0; ci0 < gs_in; ci0 += CK_CIC) {
    const int cic = min(CK_CIC, gs_in - ci0);
    if (op.up_src_L > 0) {
      stage_upsampled_rows(op, n, ci_grp + ci0, cic, in_s, pitch, width, p_base, src_s, width + 4, Lsrc, ratio);
      for (int r = cic + warp; r < CK_CIC; r += CK_NT / 32)
        for (int pos = lane; pos < width; pos += 32) in_s[r * pitch + pos] = 0.f;
    } else {
      ck_stage_input(op, n, ci_grp + ci0, cic, in_s, pitch, width, p_base, Lsrc, ratio);
    }
    for (int idx = tid; idx < CK_CIC * K * CO_B; idx += CK_NT) {
      const int col = idx % CO_B, rest = idx / CO_B;
      const int t = rest % K, r = rest / K;
      const int co = co_base + col;
      w_s[idx] = (r < cic && co < co_end) ? op.W[((size_t)co * gs_in + ci0 + r) * K + t] : 0.f;
    }
    __syncthreads();
    const float* ib = in_s + (wp * 128 + 4 * lane) * S;
    const float* wb = w_s + wc * 8;
    for (int r = 0; r < cic; ++r) ck_accumulate<K, S>(ib + r * pitch, wb + r * K * CO_B, CO_B, acc);
    __syncthreads();
  }

  // ---- epilogue ----------------------------------------------------------------------------------
  const uint64_t seed = load_seed(op.step_seed);
  const float pf = path_factor(op, seed, n), af = alpha_factor(op, seed, n);
  const bool stats = (op.out.bn >= 0) && op.bn_table[op.out.bn >= 0 ? op.out.bn : 0].use_batch;
  const int lq = l0 + wp * 128 + 4 * lane;
  const bool vec = (op.L_out & 3) == 0;
  float st[16];
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const int co = co_base + wc * 8 + c;
    float s1 = 0.f, s2 = 0.f;
    if (co < co_end) {
      const float b = op.bias ? op.bias[co] : 0.f;
      float asc = 1.f, ash = 0.f, bsc = 1.f, bsh = 0.f;
      const float *ra = nullptr, *rb = nullptr;
      if (op.res_a.C > 0) {
        view_coef(op, op.res_a, co, asc, ash);
        ra = view_row(op.res_a, n, co);
      }
      if (op.res_b.C > 0) {
        view_coef(op, op.res_b, co, bsc, bsh);
        rb = view_row(op.res_b, n, co);
      }
      float* orow = op.out.x + ((size_t)n * op.out.Ct + op.out.c0 + co) * (size_t)op.L_out;
      float r[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int l = lq + j;
        float v = 0.f;
        if (l < op.L_out) {
          v = (ck_acc(acc, c, j) + b) * pf * elem_factor(op, seed, n, co, l);
          if (ra) v += fmaf(asc, ra[l], ash);
          v *= af;
          if (rb) v += fmaf(bsc, rb[l], bsh);
          if (op.out_act == SEIST_OUT_SIGMOID) v = sigmoid_f(v);
          s1 += v;
          s2 = fmaf(v, v, s2);
        }
        r[j] = v;
      }
      if (vec && lq + 3 < op.L_out) {
        *reinterpret_cast<float4*>(orow + lq) = make_float4(r[0], r[1], r[2], r[3]);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (lq + j < op.L_out) orow[lq + j] = r[j];
      }
    }
    st[2 * c] = s1;
    st[2 * c + 1] = s2;
  }
  if (stats) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float s = warp_sum(st[i]);
      if (lane == 0) red_s[warp * 16 + i] = s;
    }
    __syncthreads();
    for (int idx = tid; idx < WC * 16; idx += CK_NT) {
      const int w0 = idx / 16, i = idx % 16;
      float s = 0.f;
      for (int p = 0; p < WP; ++p) s += red_s[(p * WC + w0) * 16 + i];
      const int co = co_base + w0 * 8 + (i >> 1);
      if (co < co_end) {
        const SeistBN& e = op.bn_table[op.out.bn];
        atomicAdd(&e.stat_acc[(i & 1) * e.C + op.out.bn_c0 + co], (double)s);
      }
    }
  }
}

// ================================================================================================
// backward (data):  d in[ci][p] = sum_{co,t} W[co][ci][t] gacc[co][(p + pad_left - t) / S]
//   S = 1: the forward engine over the combined output gradient with flipped/transposed weights.
//   S = 2 (pad_left even): input positions split by parity, p = 2u + r.  Parity r only meets the taps
//     t = r + 2s, so each parity is a stride-1 correlation of gacc with a sub-filter of KE = (K+1)/2 taps
//     (the odd one padded with a zero tap so that both share one window): the engine runs with 8 "virtual"
//     channels per warp = 4 input channels x 2 parities, and a thread that owns 4 consecutive u writes 8
//     consecutive input samples per channel.
// ================================================================================================
template <int K, int S>
__global__ void __launch_bounds__(CK_NT, 3) convk_bwd_data_kernel(const __grid_constant__ SeistOp op, const int WC) {
  extern __shared__ __align__(16) float ck_smem[];
  constexpr int KE = S == 2 ? (K + 1) / 2 : K;          // taps of the stride-1 engine
  constexpr int CPW = S == 2 ? 4 : 8;                   // real input channels per warp
  const int WP = 8 / WC, CI_B = CPW * WC, VC_B = 8 * WC, TLo = 128 * WP;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wc = warp % WC, wp = warp / WC;
  const int gs_in = op.Cin / op.groups, gs_out = op.Cout / op.groups;
  const int tpg = (gs_in + CI_B - 1) / CI_B;
  const int grp = blockIdx.z / tpg;
  const int n = blockIdx.y, p0 = blockIdx.x * TLo, ci_base = grp * gs_in + (blockIdx.z - grp * tpg) * CI_B;
  const int ci_end = (grp + 1) * gs_in;
  const int co_grp = grp * gs_out;
  const int width = TLo + KE - 1;
  const int pitch = ((width + 3) & ~3) + 4;
  float* z_s = ck_smem;                                // [CIC][pitch]
  float* w_s = ck_smem + CK_CIC * pitch;               // [CIC][KE][VC_B]
  float* red_s = w_s + CK_CIC * KE * VC_B;             // [8][16]
  const uint64_t seed = load_seed(op.step_seed);
  const float pf = path_factor(op, seed, n) * alpha_factor(op, seed, n);
  const int m_base = p0 + op.pad_left / S - (KE - 1);  // output-sample coordinate of z_s[.][0]

  float2 acc[4][4];
#pragma unroll
  for (int c = 0; c < 4; ++c)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[c][j] = make_float2(0.f, 0.f);

  for (int co0 = 0; co0 < gs_out; co0 += CK_CIC) {
    const int coc = min(CK_CIC, gs_out - co0);
    for (int r = warp; r < CK_CIC; r += CK_NT / 32) {
      float* dst = z_s + r * pitch;
      if (r >= coc) {
        for (int pos = lane; pos < width; pos += 32) dst[pos] = 0.f;
        continue;
      }
      const int co = co_grp + co0 + r;
      const OutGradCoef kc = out_grad_coef(op, co);
      for (int pos0 = lane; pos0 < width; pos0 += 32 * 4) {
        float v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int pos = pos0 + 32 * u, m = m_base + pos;
          v[u] = (pos < width && m >= 0 && m < op.L_out) ? out_grad_at(op, kc, n, co, m) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int pos = pos0 + 32 * u, m = m_base + pos;
          if (pos < width) {
            float t = v[u] * pf;
            if (op.p_elem > 0.f && m >= 0 && m < op.L_out) t *= elem_factor(op, seed, n, co, m);
            dst[pos] = t;
          }
        }
      }
    }
    // flipped + transposed weights: w_s[(r*KE + tf)*VC_B + col] = W[co0+r][ci(col)][t(tf, parity(col))]
    for (int idx = tid; idx < CK_CIC * KE * VC_B; idx += CK_NT) {
      const int col = idx % VC_B, rest = idx / VC_B;
      const int tf = rest % KE, r = rest / KE;
      const int ci = ci_base + (S == 2 ? (col >> 1) : col);
      const int t = S == 2 ? (col & 1) + 2 * (KE - 1 - tf) : (K - 1 - tf);
      w_s[idx] = (r < coc && ci < ci_end && t < K) ? op.W[((size_t)(co_grp + co0 + r) * gs_in + (ci - grp * gs_in)) * K + t] : 0.f;
    }
    __syncthreads();
    const float* zb = z_s + wp * 128 + 4 * lane;
    const float* wb = w_s + wc * 8;
    for (int r = 0; r < coc; ++r) ck_accumulate<KE, 1>(zb + r * pitch, wb + r * KE * VC_B, VC_B, acc);
    __syncthreads();
  }

  // ---- route to the source view ----------------------------------------------------------------------
  const int Lsrc = op.in[0].L;
  const float ratio = op.up_src_L > 0 ? (float)Lsrc / (float)op.L_in : 1.f;
  const int pq = p0 + wp * 128 + 4 * lane;
  float st[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) st[i] = 0.f;
  // fast path (whole quads inside the row, no up-sampling): 16-byte accesses, and the loads of a batch of
  // channels (x for khat / GELU', the old gradient when accumulating) are all issued before its first store
  const int pos0 = S == 2 ? 2 * pq : pq;
  const bool fast = op.up_src_L == 0 && (op.L_in & 3) == 0 && pos0 + 4 * S - 1 < op.L_in;
  if (fast) {
    const SeistView& v = op.in[0];          // convk ops have a single input view
    if (v.g != nullptr) {
      constexpr int NV = S == 2 ? 2 : 1;    // float4 per real channel
      constexpr int CB = S == 2 ? 2 : 4;    // real channels per batch (4 x-loads + 4 old-loads in flight)
      const bool need_x = v.act == SEIST_ACT_GELU || v.bn >= 0;
#pragma unroll
      for (int kb = 0; kb < CPW; kb += CB) {
        float4 xv[CB * NV], ov[CB * NV];
#pragma unroll
        for (int u = 0; u < CB; ++u) {
          const int ci = min(ci_base + wc * CPW + kb + u, ci_end - 1);
          const size_t off = ((size_t)n * v.Ct + v.c0 + ci) * (size_t)v.L + pos0;
#pragma unroll
          for (int q = 0; q < NV; ++q) {
            xv[u * NV + q] = need_x ? __ldg(reinterpret_cast<const float4*>(v.x + off) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
            ov[u * NV + q] = v.accum ? *(reinterpret_cast<const float4*>(v.g + off) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
#pragma unroll
        for (int u = 0; u < CB; ++u) {
          const int k = kb + u;
          const int ci = ci_base + wc * CPW + k;
          if (ci >= ci_end) continue;
          float sc, sh, mu = 0.f, istd = 0.f;
          view_coef(op, v, ci, sc, sh);
          if (v.bn >= 0) view_khat(op, v, ci, mu, istd);
          const size_t off = ((size_t)n * v.Ct + v.c0 + ci) * (size_t)v.L + pos0;
#pragma unroll
          for (int q = 0; q < NV; ++q) {
            const float4 x4 = xv[u * NV + q], o4 = ov[u * NV + q];
            const float xs[4] = {x4.x, x4.y, x4.z, x4.w};
            const float os[4] = {o4.x, o4.y, o4.z, o4.w};
            float g[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              // S = 1: element e is sample j = e of channel k; S = 2: element 4q+e is (j, parity) = ((4q+e)>>1, e&1)
              const int c = S == 2 ? 2 * k + (e & 1) : k;
              const int j = S == 2 ? (4 * q + e) >> 1 : e;
              float gv = ck_acc(acc, c, j);
              if (v.act == SEIST_ACT_GELU) gv *= gelu_grad_f(fmaf(sc, xs[e], sh));
              st[2 * c] += gv;
              st[2 * c + 1] = fmaf(gv, (xs[e] - mu) * istd, st[2 * c + 1]);
              g[e] = gv + os[e];
            }
            *(reinterpret_cast<float4*>(v.g + off) + q) = make_float4(g[0], g[1], g[2], g[3]);
          }
        }
      }
    }
  } else if (S == 1 && op.up_src_L > 0 && 2 * Lsrc == op.L_in && (op.L_in & 3) == 0) {
    // exact x2 linear up-sampling (the dpk head, reference models/seist.py:566): the quad p = 4m .. 4m+3 only touches
    // the sources 2m-1 .. 2m+2 with the fixed weights (.25 | .75 .75 .25 | .25 .75 .75 | .25).  Every lane owns the
    // sources 2m, 2m+1: the two outer contributions travel to the neighbouring lanes by shuffle, the activation
    // derivative and the BN sums are evaluated once per SOURCE sample, and 2 (+1 at a warp edge) atomics per
    // channel replace 8 (the target is zero-initialised by the preceding ZERO op).
    const SeistView& v = op.in[0];
    if (v.g != nullptr) {
      const bool inside = pq < op.L_in;
      const int m2 = pq >> 1;                               // source index 2m
      const bool first = pq == 0, last = pq + 4 >= op.L_in;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const int ci = ci_base + wc * 8 + c;
        const bool live = ci < ci_end;                      // warp-uniform
        const float d0 = inside && live ? ck_acc(acc, c, 0) : 0.f, d1 = inside && live ? ck_acc(acc, c, 1) : 0.f;
        const float d2 = inside && live ? ck_acc(acc, c, 2) : 0.f, d3 = inside && live ? ck_acc(acc, c, 3) : 0.f;
        float ga = 0.25f * d0, ge = 0.25f * d3;             // to 2m-1 / 2m+2
        float gb = 0.75f * (d0 + d1) + 0.25f * d2, gc = 0.25f * d1 + 0.75f * (d2 + d3);
        if (first) {
          gb += ga;
          ga = 0.f;
        }
        if (last) {
          gc += ge;
          ge = 0.f;
        }
        const float from_prev = __shfl_up_sync(0xffffffffu, ge, 1), from_next = __shfl_down_sync(0xffffffffu, ga, 1);
        if (lane > 0) gb += from_prev;
        if (lane < 31) gc += from_next;
        if (!live || !inside) continue;
        float sc, sh, mu = 0.f, istd = 0.f;
        view_coef(op, v, ci, sc, sh);
        if (v.bn >= 0) view_khat(op, v, ci, mu, istd);
        const float* xr = view_row(v, n, ci);
        float* gr = view_grad_row(v, n, ci);
        const float2 x2 = *reinterpret_cast<const float2*>(xr + m2);
        if (v.act == SEIST_ACT_GELU) {
          gb *= gelu_grad_f(fmaf(sc, x2.x, sh));
          gc *= gelu_grad_f(fmaf(sc, x2.y, sh));
        }
        atomicAdd(&gr[m2], gb);
        atomicAdd(&gr[m2 + 1], gc);
        float s1 = gb + gc;
        float s2 = fmaf(gb, (x2.x - mu) * istd, gc * ((x2.y - mu) * istd));
        if (lane == 0 && !first) {                           // 2m-1 belongs to the last lane of another warp
          const float xa = xr[m2 - 1];
          if (v.act == SEIST_ACT_GELU) ga *= gelu_grad_f(fmaf(sc, xa, sh));
          atomicAdd(&gr[m2 - 1], ga);
          s1 += ga;
          s2 = fmaf(ga, (xa - mu) * istd, s2);
        }
        if (lane == 31 && !last) {                           // 2m+2 belongs to the first lane of another warp
          const float xe = xr[m2 + 2];
          if (v.act == SEIST_ACT_GELU) ge *= gelu_grad_f(fmaf(sc, xe, sh));
          atomicAdd(&gr[m2 + 2], ge);
          s1 += ge;
          s2 = fmaf(ge, (xe - mu) * istd, s2);
        }
        st[2 * c] = s1;
        st[2 * c + 1] = s2;
      }
    }
  } else
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const int ci = ci_base + wc * CPW + (S == 2 ? (c >> 1) : c);
    float s1 = 0.f, s2 = 0.f;
    if (ci < ci_end) {
      int cv;
      const int vi = resolve_view(op, ci, cv);
      const SeistView& v = op.in[vi];
      if (v.g != nullptr) {
        float sc, sh, mu = 0.f, istd = 0.f;
        view_coef(op, v, cv, sc, sh);
        if (v.bn >= 0) view_khat(op, v, cv, mu, istd);
        const float* xr = view_row(v, n, cv);
        float* gr = view_grad_row(v, n, cv);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int p = S == 2 ? 2 * (pq + j) + (c & 1) : pq + j;
          if (p >= op.L_in) continue;
          const float d = ck_acc(acc, c, j);
          if (op.up_src_L > 0) {
            int i0, i1;
            float lam;
            upsample_coords(p, ratio, Lsrc, i0, i1, lam);
            const float x0 = xr[i0], x1 = xr[i1];
            float g0 = d * (1.f - lam), g1 = d * lam;
            if (v.act == SEIST_ACT_GELU) {
              g0 *= gelu_grad_f(fmaf(sc, x0, sh));
              g1 *= gelu_grad_f(fmaf(sc, x1, sh));
            }
            atomicAdd(&gr[i0], g0);
            atomicAdd(&gr[i1], g1);
            s1 += g0 + g1;
            s2 = fmaf(g0, (x0 - mu) * istd, s2);
            s2 = fmaf(g1, (x1 - mu) * istd, s2);
          } else {
            const float x = xr[p];
            float g = d;
            if (v.act == SEIST_ACT_GELU) g *= gelu_grad_f(fmaf(sc, x, sh));
            if (v.accum) gr[p] += g; else gr[p] = g;
            s1 += g;
            s2 = fmaf(g, (x - mu) * istd, s2);
          }
        }
      }
    }
    st[2 * c] = s1;
    st[2 * c + 1] = s2;
  }
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const float s = warp_sum(st[i]);
    if (lane == 0) red_s[warp * 16 + i] = s;
  }
  __syncthreads();
  for (int idx = tid; idx < WC * 16; idx += CK_NT) {
    const int w0 = idx / 16, i = idx % 16;
    float s = 0.f;
    for (int p = 0; p < WP; ++p) s += red_s[(p * WC + w0) * 16 + i];
    const int ci = ci_base + w0 * CPW + (S == 2 ? (i >> 2) : (i >> 1));
    if (ci < ci_end) {
      int cv;
      const int vi = resolve_view(op, ci, cv);
      const SeistView& v = op.in[vi];
      if (v.g != nullptr && v.bn >= 0) {
        const SeistBN& e = op.bn_table[v.bn];
        atomicAdd(&e.gstat_acc[(i & 1) * e.C + v.bn_c0 + cv], (double)s);
      }
    }
  }
}

// ================================================================================================
// launchers
// ================================================================================================
bool convk_eligible(const SeistOp& op) {
  if (op.pool > 1 || op.n_in != 1) return false;
  if (op.groups > 1 && ((op.Cin / op.groups) < 8 || (op.Cout / op.groups) < 8)) return false;
  if (op.stride != 1 && op.stride != 2) return false;
  switch (op.k) {
    case 3: case 5: case 7: case 9: case 11: case 13: case 15: case 19: break;
    default: return false;
  }
  if (op.stride == 2 && !(op.k == 7 || op.k == 11 || op.k == 15 || op.k == 19)) return false;
  return true;
}

template <typename Kf>
static int ck_set_smem(Kf kernel, size_t bytes) {
  if (bytes > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return (int)e;
  }
  return 0;
}

static int pick_wc(int channels) { return channels > 32 ? 8 : (channels > 16 ? 4 : (channels > 8 ? 2 : 1)); }

template <int K, int S>
static int launch_fwd_ks(const SeistOp& op, cudaStream_t s) {
  const int gs_out = op.Cout / op.groups;
  const int WC = pick_wc(gs_out), WP = 8 / WC, CO_B = 8 * WC, TLo = 128 * WP;
  const int width = TLo * S + K - S, pitch = ((width + 3) & ~3) + 4;
  const size_t smem = sizeof(float) * ((size_t)CK_CIC * pitch + (size_t)CK_CIC * K * CO_B + 8 * 16 +
                                       (op.up_src_L > 0 ? (size_t)CK_CIC * (width + 4) : 0));
  dim3 grid((op.L_out + TLo - 1) / TLo, op.N, op.groups * ((gs_out + CO_B - 1) / CO_B));
  int rc = ck_set_smem(convk_fwd_kernel<K, S>, smem);
  if (rc) return rc;
  convk_fwd_kernel<K, S><<<grid, CK_NT, smem, s>>>(op, WC);
  note_launch();
  return check_launch("convk_fwd");
}

template <int K, int S>
static int launch_bwdd_k(const SeistOp& op, cudaStream_t s) {
  constexpr int KE = S == 2 ? (K + 1) / 2 : K;
  constexpr int CPW = S == 2 ? 4 : 8;
  const int gs_in = op.Cin / op.groups;
  const int WC = pick_wc(gs_in * (8 / CPW)), WP = 8 / WC, CI_B = CPW * WC, TLo = 128 * WP;
  const int width = TLo + KE - 1, pitch = ((width + 3) & ~3) + 4;
  const size_t smem = sizeof(float) * ((size_t)CK_CIC * pitch + (size_t)CK_CIC * KE * 8 * WC + 8 * 16);
  dim3 grid(((op.L_in + S - 1) / S + TLo - 1) / TLo, op.N, op.groups * ((gs_in + CI_B - 1) / CI_B));
  int rc = ck_set_smem(convk_bwd_data_kernel<K, S>, smem);
  if (rc) return rc;
  convk_bwd_data_kernel<K, S><<<grid, CK_NT, smem, s>>>(op, WC);
  note_launch();
  return check_launch("convk_bwd_data");
}

#define CK_SWITCH_K(FN, ...)              \
  switch (op.k) {                         \
    case 3: return FN(3, __VA_ARGS__);    \
    case 5: return FN(5, __VA_ARGS__);    \
    case 7: return FN(7, __VA_ARGS__);    \
    case 9: return FN(9, __VA_ARGS__);    \
    case 11: return FN(11, __VA_ARGS__);  \
    case 13: return FN(13, __VA_ARGS__);  \
    case 15: return FN(15, __VA_ARGS__);  \
    default: return FN(19, __VA_ARGS__);  \
  }

int launch_convk_fwd(const SeistOp& op, cudaStream_t s) {
  if (op.stride == 2) {
    switch (op.k) {
      case 7: return launch_fwd_ks<7, 2>(op, s);
      case 11: return launch_fwd_ks<11, 2>(op, s);
      case 15: return launch_fwd_ks<15, 2>(op, s);
      default: return launch_fwd_ks<19, 2>(op, s);
    }
  }
#define FWD1(KK, dummy) launch_fwd_ks<KK, 1>(op, s)
  CK_SWITCH_K(FWD1, 0)
#undef FWD1
}

// stride 2 needs an even left pad and no up-sampling (true for every strided conv of the model family)
bool convk_bwd_data_eligible(const SeistOp& op) {
  if (!convk_eligible(op)) return false;
  if (op.stride == 1) return true;
  return (op.pad_left & 1) == 0 && op.up_src_L == 0;
}

int launch_convk_bwd_data(const SeistOp& op, cudaStream_t s) {
  if (op.stride == 2) {
    switch (op.k) {
      case 7: return launch_bwdd_k<7, 2>(op, s);
      case 11: return launch_bwdd_k<11, 2>(op, s);
      case 15: return launch_bwdd_k<15, 2>(op, s);
      default: return launch_bwdd_k<19, 2>(op, s);
    }
  }
#define BD(KK, dummy) launch_bwdd_k<KK, 1>(op, s)
  CK_SWITCH_K(BD, 0)
#undef BD
}

}  // namespace seist
