// Generalised 1-D convolution family (SEIST_OP_CONV_FWD / _BWD_DATA / _BWD_W / RES_BWD).
//
// Layout: tensors are (N, C, L) fp32 with the sample axis contiguous.  A CTA owns TL = 128 consecutive
// output samples of one waveform and a tile of output channels; lanes run along the sample axis so every
// global access is a 128-byte coalesced row segment, warps run along channels.  The (channel, sample)
// input tile — with BatchNorm-apply, GELU, pooling / linear up-sampling and zero padding evaluated while
// it is staged — lives in shared memory, weights are staged k-major so a thread reads its COT output
// channels as one broadcast vector.  BatchNorm statistics of the result are reduced with warp shuffles
// and added to the BN table in the epilogue, so BN never costs its own pass.
#include "common.cuh"
#include "conv_common.cuh"

namespace seist {

constexpr int TL = 128;   // samples per CTA
constexpr int NT = 128;   // threads per CTA: 4 warps (channel rows) x 32 lanes (samples)
constexpr int ROWS = 4;
constexpr int CIC = 16;   // reduction channels staged per chunk

// Stage `nrows` conv-input rows of `width` samples into shared memory (row r at dst + r*pitch).  A warp
// keeps four rows in flight per pass so four independent global loads are outstanding per lane.
template <typename ChanFn>
__device__ __forceinline__ void stage_rows(const SeistOp& op, int n, float* dst, int pitch, int width, int nrows,
                                           int p_base, int Lsrc, float ratio, ChanFn&& chan_of) {
  const int lane = threadIdx.x & 31, row = threadIdx.x >> 5;
  const bool plain = op.pool <= 1 && op.up_src_L == 0;
  for (int r0 = row; r0 < nrows; r0 += 4 * ROWS) {
    RowSrc rs[4];
    bool valid[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int r = r0 + u * ROWS;
      valid[u] = false;
      int ci = 0;
      if (r < nrows) ci = chan_of(r, valid[u]);
      rs[u] = make_row(op, n, valid[u] ? ci : 0);
    }
    for (int pos = lane; pos < width; pos += 32) {
      const int p = p_base + pos;
      float v[4];
      if (plain) {
        const bool inb = p >= 0 && p < op.L_in;
        const int pc = inb ? p : 0;
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = rs[u].x[pc];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          float t = fmaf(rs[u].sc, v[u], rs[u].sh);
          if (rs[u].act == SEIST_ACT_GELU) t = gelu_f(t);
          v[u] = (inb && valid[u]) ? t : 0.f;
        }
      } else {
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = valid[u] ? conv_input_at(op, rs[u], p, Lsrc, ratio) : 0.f;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int r = r0 + u * ROWS;
        if (r < nrows) dst[r * pitch + pos] = v[u];
      }
    }
  }
}

// ================================================================================================
// forward
// ================================================================================================
template <int COT>
__global__ void __launch_bounds__(NT) conv_fwd_kernel(const __grid_constant__ SeistOp op) {
  extern __shared__ float smem[];
  constexpr int CO_TILE = ROWS * COT;
  const int lane = threadIdx.x & 31, row = threadIdx.x >> 5;
  const int n = blockIdx.y;
  const int l0 = blockIdx.x * TL;
  const int co_base = blockIdx.z * CO_TILE;
  const int gs_in = op.Cin / op.groups, gs_out = op.Cout / op.groups;
  const int g_lo = co_base / gs_out;
  const int g_hi = (min(co_base + CO_TILE, op.Cout) - 1) / gs_out;
  const int ng = g_hi - g_lo + 1;
  const int qc = min(CIC, gs_in);
  const int k = op.k, stride = op.stride;
  const int TLin = (TL - 1) * stride + k;
  float* in_s = smem;                       // [ng][qc][TLin]
  float* w_s = smem + ng * qc * TLin;       // [qc][k][CO_TILE]
  const int Lsrc = op.in[0].L;
  const float ratio = op.up_src_L > 0 ? (float)Lsrc / (float)op.L_in : 1.f;

  const int co0 = co_base + row * COT;
  const int gl = (co0 < op.Cout ? co0 / gs_out : g_lo) - g_lo;

  float acc[COT][4];
#pragma unroll
  for (int c = 0; c < COT; ++c)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[c][j] = 0.f;

  const int p_base = l0 * stride - op.pad_left;
  for (int q0 = 0; q0 < gs_in; q0 += qc) {
    // ---- stage the input tile: one warp per channel row, lanes along samples -------------------
    stage_rows(op, n, in_s, TLin, TLin, ng * qc, p_base, Lsrc, ratio, [&](int r, bool& valid) {
      const int g = r / qc, qq = r - g * qc;
      valid = q0 + qq < gs_in;
      return (g_lo + g) * gs_in + q0 + qq;
    });
    // ---- stage weights k-major: w_s[(qq*k + t)*CO_TILE + col] ----------------------------------
    for (int idx = threadIdx.x; idx < qc * k * CO_TILE; idx += NT) {
      const int col = idx % CO_TILE;
      const int rest = idx / CO_TILE;
      const int t = rest % k, qq = rest / k;
      const int co = co_base + col;
      float w = 0.f;
      if (co < op.Cout && q0 + qq < gs_in) w = op.W[((size_t)co * gs_in + q0 + qq) * k + t];
      w_s[idx] = w;
    }
    __syncthreads();
    // ---- compute -------------------------------------------------------------------------------
    if (co0 < op.Cout) {
      const float* ib = in_s + gl * qc * TLin + lane * stride;
      for (int qq = 0; qq < qc; ++qq) {
        const float* irow = ib + qq * TLin;
        const float* wrow = w_s + (qq * k) * CO_TILE + row * COT;
        for (int t = 0; t < k; ++t) {
          float w[COT];
#pragma unroll
          for (int c = 0; c < COT; ++c) w[c] = wrow[t * CO_TILE + c];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float v = irow[32 * j * stride + t];
#pragma unroll
            for (int c = 0; c < COT; ++c) acc[c][j] = fmaf(w[c], v, acc[c][j]);
          }
        }
      }
    }
    __syncthreads();
  }

  // ---- epilogue ----------------------------------------------------------------------------------
  const uint64_t seed = load_seed(op.step_seed);
  const float pf = path_factor(op, seed, n), af = alpha_factor(op, seed, n);
  const bool stats = (op.out.bn >= 0) && op.bn_table[op.out.bn >= 0 ? op.out.bn : 0].use_batch;
#pragma unroll
  for (int c = 0; c < COT; ++c) {
    const int co = co0 + c;
    if (co >= op.Cout) break;   // warp-uniform
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
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int l = l0 + lane + 32 * j;
      if (l < op.L_out) {
        float v = (acc[c][j] + b) * pf * elem_factor(op, seed, n, co, l);
        if (ra) v += fmaf(asc, ra[l], ash);
        v *= af;
        if (rb) v += fmaf(bsc, rb[l], bsh);
        if (op.out_act == SEIST_OUT_SIGMOID) v = sigmoid_f(v);
        orow[l] = v;
        s1 += v;
        s2 = fmaf(v, v, s2);
      }
    }
    if (stats) {
      s1 = warp_sum(s1);
      s2 = warp_sum(s2);
      if (lane == 0) {
        const SeistBN& e = op.bn_table[op.out.bn];
        atomicAdd(&e.stat_acc[op.out.bn_c0 + co], (double)s1);
        atomicAdd(&e.stat_acc[e.C + op.out.bn_c0 + co], (double)s2);
      }
    }
  }
}

// ================================================================================================
// backward: data.  d(conv input)[ci][p] = sum_{co in group, t} W[co][q][t] * z[co][p + pad_left - t],
// z = gacc placed on the stride grid.  Same tile skeleton as the forward with the roles of the
// channel axes swapped and the taps flipped; the result tile goes through shared memory so the
// activation derivative / pooling argmax / up-sampling transpose can route it to the source view.
// ================================================================================================
template <int COT>
__global__ void __launch_bounds__(NT) conv_bwd_data_kernel(const __grid_constant__ SeistOp op) {
  extern __shared__ float smem[];
  constexpr int CI_TILE = ROWS * COT;
  const int lane = threadIdx.x & 31, row = threadIdx.x >> 5;
  const int n = blockIdx.y;
  const int p0 = blockIdx.x * TL;           // conv-input coordinates
  const int ci_base = blockIdx.z * CI_TILE;
  const int gs_in = op.Cin / op.groups, gs_out = op.Cout / op.groups;
  const int g_lo = ci_base / gs_in;
  const int g_hi = (min(ci_base + CI_TILE, op.Cin) - 1) / gs_in;
  const int ng = g_hi - g_lo + 1;
  const int oc = min(CIC, gs_out);
  const int k = op.k, stride = op.stride;
  const int TLz = TL + k - 1;
  float* z_s = smem;                          // [ng][oc][TLz]
  float* w_s = z_s + ng * oc * TLz;           // [oc][k][CI_TILE]
  float* d_s = w_s + oc * k * CI_TILE;        // [CI_TILE][TL]
  const uint64_t seed = load_seed(op.step_seed);
  const float pf = path_factor(op, seed, n) * alpha_factor(op, seed, n);

  const int ci0 = ci_base + row * COT;
  const int gl = (ci0 < op.Cin ? ci0 / gs_in : g_lo) - g_lo;

  float acc[COT][4];
#pragma unroll
  for (int c = 0; c < COT; ++c)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[c][j] = 0.f;

  const int m_base = p0 + op.pad_left - (k - 1);
  for (int o0 = 0; o0 < gs_out; o0 += oc) {
    for (int r0 = row; r0 < ng * oc; r0 += 4 * ROWS) {
      OutGradCoef kc[4];
      int cos[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int r = r0 + u * ROWS;
        const int g = r / oc, oo = r - g * oc;
        cos[u] = (r < ng * oc && o0 + oo < gs_out) ? (g_lo + g) * gs_out + o0 + oo : -1;
        kc[u] = out_grad_coef(op, cos[u] >= 0 ? cos[u] : 0);
      }
      for (int pos = lane; pos < TLz; pos += 32) {
        const int m = m_base + pos;
        const int l = m >= 0 ? m / stride : 0;
        const bool inb = m >= 0 && l * stride == m && l < op.L_out;
        float v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = (inb && cos[u] >= 0) ? out_grad_at(op, kc[u], n, cos[u], l) : 0.f;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int r = r0 + u * ROWS;
          if (r < ng * oc) {
            float t = v[u] * pf;
            if (op.p_elem > 0.f && inb && cos[u] >= 0) t *= elem_factor(op, seed, n, cos[u], l);
            z_s[r * TLz + pos] = t;
          }
        }
      }
    }
    // flipped, transposed weights: w_s[(oo*k + tf)*CI_TILE + col] = W[co][q][k-1-tf]
    for (int idx = threadIdx.x; idx < oc * k * CI_TILE; idx += NT) {
      const int col = idx % CI_TILE;
      const int rest = idx / CI_TILE;
      const int tf = rest % k, oo = rest / k;
      const int ci = ci_base + col;
      float w = 0.f;
      if (ci < op.Cin && o0 + oo < gs_out) {
        const int g = ci / gs_in, q = ci - g * gs_in;
        const int co = g * gs_out + o0 + oo;
        w = op.W[((size_t)co * gs_in + q) * k + (k - 1 - tf)];
      }
      w_s[idx] = w;
    }
    __syncthreads();
    if (ci0 < op.Cin) {
      const float* zb = z_s + gl * oc * TLz + lane;
      for (int oo = 0; oo < oc; ++oo) {
        const float* zrow = zb + oo * TLz;
        const float* wrow = w_s + (oo * k) * CI_TILE + row * COT;
        for (int t = 0; t < k; ++t) {
          float w[COT];
#pragma unroll
          for (int c = 0; c < COT; ++c) w[c] = wrow[t * CI_TILE + c];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float v = zrow[32 * j + t];
#pragma unroll
            for (int c = 0; c < COT; ++c) acc[c][j] = fmaf(w[c], v, acc[c][j]);
          }
        }
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int c = 0; c < COT; ++c)
#pragma unroll
    for (int j = 0; j < 4; ++j) d_s[(row * COT + c) * TL + lane + 32 * j] = acc[c][j];
  __syncthreads();

  // ---- route to the source views: one warp per channel row --------------------------------------
  const int Lsrc = op.in[0].L;
  const float ratio = op.up_src_L > 0 ? (float)Lsrc / (float)op.L_in : 1.f;
  for (int r = row; r < CI_TILE; r += ROWS) {
    const int ci = ci_base + r;
    if (ci >= op.Cin) break;
    int cv;
    const int vi = resolve_view(op, ci, cv);
    const SeistView& v = op.in[vi];
    if (v.g == nullptr) continue;   // no gradient wanted for this view (warp-uniform)
    float sc, sh, mu = 0.f, istd = 0.f;
    view_coef(op, v, cv, sc, sh);
    const bool has_bn = v.bn >= 0;
    if (has_bn) view_khat(op, v, cv, mu, istd);
    const float* xr = view_row(v, n, cv);
    float* gr = view_grad_row(v, n, cv);
    const float* drow = d_s + r * TL;
    float s1 = 0.f, s2 = 0.f;
    for (int j = 0; j < 4; ++j) {
      const int pos = lane + 32 * j;
      const int p = p0 + pos;
      if (p >= op.L_in) continue;
      const float d = drow[pos];
      if (op.pool > 1) {
        const int s0 = p * op.pool;
        const int cnt = min(op.pool, Lsrc - s0);
        int am = 0;
        float mx = -INFINITY;
        for (int i = 0; i < cnt; ++i) {
          const float u = fmaf(sc, xr[s0 + i], sh);
          if (u > mx) {
            mx = u;
            am = i;
          }
        }
        const float inv = 1.f / (float)cnt;
        for (int i = 0; i < cnt; ++i) {
          const float x = xr[s0 + i];
          float g = d * (inv + (i == am ? 1.f : 0.f));
          if (v.act == SEIST_ACT_GELU) g *= gelu_grad_f(fmaf(sc, x, sh));
          if (v.accum) gr[s0 + i] += g; else gr[s0 + i] = g;
          s1 += g;
          s2 = fmaf(g, (x - mu) * istd, s2);
        }
      } else if (op.up_src_L > 0) {
        int i0, i1;
        float lam;
        upsample_coords(p, ratio, Lsrc, i0, i1, lam);
        const float x0 = xr[i0], x1 = xr[i1];
        float g0 = d * (1.f - lam), g1 = d * lam;
        if (v.act == SEIST_ACT_GELU) {
          g0 *= gelu_grad_f(fmaf(sc, x0, sh));
          g1 *= gelu_grad_f(fmaf(sc, x1, sh));
        }
        atomicAdd(&gr[i0], g0);   // target zero-initialised by a preceding ZERO op
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
    if (has_bn) {
      s1 = warp_sum(s1);
      s2 = warp_sum(s2);
      if (lane == 0) gstat_add(op, v, cv, s1, s2);
    }
  }
}

// ================================================================================================
// backward: weights.  dW[co][q][t] = sum_{n,l} gacc[co][n,l] * convin[ci][n, l*stride + t - pad_left]
// A CTA owns a (32 co) x (128 (q,t)) tile of dW and a strided share of all (n, sample-chunk) tiles;
// lanes run along the (q,t) axis, the sample axis is the reduction.  Partial tiles are merged with
// float atomics (dW is zero-initialised by the host at the start of every backward pass).
// ================================================================================================
constexpr int PC = 32;      // samples per reduction chunk
constexpr int RT = 128;     // (q,t) pairs per CTA

template <int COT>
__global__ void __launch_bounds__(NT) conv_bwd_w_kernel(const __grid_constant__ SeistOp op) {
  extern __shared__ float smem[];
  constexpr int CO_TILE = ROWS * COT;
  const int lane = threadIdx.x & 31, row = threadIdx.x >> 5;
  const int co_base = blockIdx.y * CO_TILE;
  const int r_base = blockIdx.z * RT;
  const int gs_in = op.Cin / op.groups, gs_out = op.Cout / op.groups;
  const int k = op.k, stride = op.stride;
  const int R = gs_in * k;
  const int g_lo = co_base / gs_out;
  const int g_hi = (min(co_base + CO_TILE, op.Cout) - 1) / gs_out;
  const int ng = g_hi - g_lo + 1;
  const int q_lo = r_base / k;
  const int q_hi = min((r_base + RT - 1) / k, gs_in - 1);
  const int nq = q_hi - q_lo + 1;
  const int TLin = (PC - 1) * stride + k;
  const int TLp = TLin | 1;                  // odd row pitch: lanes on different rows hit different banks
  constexpr int GP = CO_TILE + 4;            // row pitch of g_s: spreads the transposed stores over banks
  float* g_s = smem;                         // [PC][GP]
  float* in_s = smem + PC * GP;              // [ng][nq][TLp]
  const uint64_t seed = load_seed(op.step_seed);
  const int Lsrc = op.in[0].L;
  const float ratio = op.up_src_L > 0 ? (float)Lsrc / (float)op.L_in : 1.f;

  const int co0 = co_base + row * COT;
  const int gl = (co0 < op.Cout ? co0 / gs_out : g_lo) - g_lo;
  int ioff[4];
  bool rvalid[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int r = r_base + lane + 32 * j;
    rvalid[j] = r < R;
    const int q = rvalid[j] ? r / k : q_lo;
    const int t = rvalid[j] ? r - q * k : 0;
    ioff[j] = (gl * nq + (q - q_lo)) * TLp + t;
  }
  float acc[COT][4];
  float bacc[COT];
#pragma unroll
  for (int c = 0; c < COT; ++c) {
    bacc[c] = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[c][j] = 0.f;
  }

  const int chunks_per_n = (op.L_out + PC - 1) / PC;
  const int total = op.N * chunks_per_n;
  for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
    const int n = tile / chunks_per_n;
    const int l0 = (tile - n * chunks_per_n) * PC;
    const float pf = path_factor(op, seed, n) * alpha_factor(op, seed, n);
    // gacc tile, transposed: g_s[p][col]
    for (int c0 = row; c0 < CO_TILE; c0 += 4 * ROWS) {
      const int l = l0 + lane;   // PC == 32
      float v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int co = co_base + c0 + u * ROWS;
        v[u] = 0.f;
        if (c0 + u * ROWS < CO_TILE && co < op.Cout && l < op.L_out) {
          const OutGradCoef kc = out_grad_coef(op, co);
          v[u] = out_grad_at(op, kc, n, co, l);
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int col = c0 + u * ROWS, co = co_base + col;
        if (col < CO_TILE) {
          float t = v[u] * pf;
          if (op.p_elem > 0.f && co < op.Cout && l < op.L_out) t *= elem_factor(op, seed, n, co, l);
          g_s[lane * GP + col] = t;
        }
      }
    }
    // conv-input rows
    const int p_base = l0 * stride - op.pad_left;
    stage_rows(op, n, in_s, TLp, TLin, ng * nq, p_base, Lsrc, ratio, [&](int r, bool& valid) {
      const int g = r / nq, qq = r - g * nq;
      valid = true;
      return (g_lo + g) * gs_in + q_lo + qq;
    });
    __syncthreads();
    if (co0 < op.Cout) {
      for (int p = 0; p < PC; ++p) {
        float g[COT];
#pragma unroll
        for (int c = 0; c < COT; ++c) g[c] = g_s[p * GP + row * COT + c];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float v = in_s[ioff[j] + p * stride];
#pragma unroll
          for (int c = 0; c < COT; ++c) acc[c][j] = fmaf(g[c], v, acc[c][j]);
        }
      }
      if (blockIdx.z == 0 && op.dbias != nullptr) {
#pragma unroll
        for (int c = 0; c < COT; ++c) bacc[c] += g_s[lane * GP + row * COT + c];
      }
    }
    __syncthreads();
  }
  if (co0 < op.Cout) {
#pragma unroll
    for (int c = 0; c < COT; ++c) {
      const int co = co0 + c;
      if (co >= op.Cout) break;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int r = r_base + lane + 32 * j;
        if (rvalid[j]) atomicAdd(&op.dW[(size_t)co * R + r], acc[c][j]);
      }
      if (blockIdx.z == 0 && op.dbias != nullptr) {
        const float s = warp_sum(bacc[c]);
        if (lane == 0) atomicAdd(&op.dbias[co], s);
      }
    }
  }
}

// ================================================================================================
// backward: residual pass-through.  res_a gets alpha(n) * dOut, res_b gets dOut.
// grid (sample tiles, Cout, N), one warp-row per block row segment.
// ================================================================================================
__global__ void __launch_bounds__(NT) res_bwd_kernel(const __grid_constant__ SeistOp op) {
  const int n = blockIdx.z, co = blockIdx.y;
  const int l = blockIdx.x * NT + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const uint64_t seed = load_seed(op.step_seed);
  const float af = alpha_factor(op, seed, n);
  const OutGradCoef kc = out_grad_coef(op, co);
  const bool ok = l < op.L_out;
  const float g = ok ? out_grad_at(op, kc, n, co, l) : 0.f;
#pragma unroll
  for (int which = 0; which < 2; ++which) {
    const SeistView& v = which == 0 ? op.res_a : op.res_b;
    if (v.C == 0 || v.g == nullptr) continue;
    const float gv = which == 0 ? g * af : g;
    float s1 = 0.f, s2 = 0.f;
    if (ok) {
      float* gr = view_grad_row(v, n, co);
      if (v.accum) gr[l] += gv; else gr[l] = gv;
    }
    if (v.bn >= 0) {
      float mu, istd;
      view_khat(op, v, co, mu, istd);
      if (ok) {
        const float x = view_row(v, n, co)[l];
        s1 = gv;
        s2 = gv * (x - mu) * istd;
      }
      s1 = warp_sum(s1);
      s2 = warp_sum(s2);
      if (lane == 0) gstat_add(op, v, co, s1, s2);
    }
  }
}

// ================================================================================================
// host-side launchers
// ================================================================================================
static int pick_cot(int per_group_out, int total_out, int groups) {
  // threads own COT output channels of ONE group; 4 warp rows per CTA
  int cot;
  if (groups > 1) {
    cot = per_group_out >= 8 ? 8 : (per_group_out >= 4 ? 4 : (per_group_out >= 2 ? 2 : 1));
    while (per_group_out % cot) cot >>= 1;
  } else {
    cot = total_out > 32 ? 16 : (total_out > 16 ? 8 : (total_out > 8 ? 4 : (total_out > 4 ? 2 : 1)));
  }
  return cot;
}

template <typename K>
static int set_smem(K kernel, size_t bytes) {
  if (bytes > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return (int)e;
  }
  return 0;
}

#define DISPATCH_COT(cot, CALL) \
  switch (cot) {                \
    case 1: CALL(1); break;     \
    case 2: CALL(2); break;     \
    case 4: CALL(4); break;     \
    case 8: CALL(8); break;     \
    default: CALL(16); break;   \
  }

int launch_conv_fwd(const SeistOp& op, cudaStream_t s) {
  const int gs_in = op.Cin / op.groups, gs_out = op.Cout / op.groups;
  const int cot = pick_cot(gs_out, op.Cout, op.groups);
  const int co_tile = ROWS * cot;
  const int qc = gs_in < CIC ? gs_in : CIC;
  const int ng_max = op.groups > 1 ? (co_tile + gs_out - 1) / gs_out + (co_tile % gs_out ? 1 : 0) : 1;
  const int TLin = (TL - 1) * op.stride + op.k;
  const size_t smem = sizeof(float) * ((size_t)ng_max * qc * TLin + (size_t)qc * op.k * co_tile);
  dim3 grid((op.L_out + TL - 1) / TL, op.N, (op.Cout + co_tile - 1) / co_tile);
  int rc = 0;
#define CALL(C)                                                         \
  rc = set_smem(conv_fwd_kernel<C>, smem);                              \
  if (!rc) conv_fwd_kernel<C><<<grid, NT, smem, s>>>(op);
  DISPATCH_COT(cot, CALL)
#undef CALL
  if (rc) return rc;
  note_launch();
  return check_launch("conv_fwd");
}

int launch_conv_bwd_data(const SeistOp& op, cudaStream_t s) {
  const int gs_in = op.Cin / op.groups, gs_out = op.Cout / op.groups;
  const int cot = pick_cot(gs_in, op.Cin, op.groups);
  const int ci_tile = ROWS * cot;
  const int oc = gs_out < CIC ? gs_out : CIC;
  const int ng_max = op.groups > 1 ? (ci_tile + gs_in - 1) / gs_in + (ci_tile % gs_in ? 1 : 0) : 1;
  const int TLz = TL + op.k - 1;
  const size_t smem = sizeof(float) * ((size_t)ng_max * oc * TLz + (size_t)oc * op.k * ci_tile + (size_t)ci_tile * TL);
  dim3 grid((op.L_in + TL - 1) / TL, op.N, (op.Cin + ci_tile - 1) / ci_tile);
  int rc = 0;
#define CALL(C)                                                         \
  rc = set_smem(conv_bwd_data_kernel<C>, smem);                         \
  if (!rc) conv_bwd_data_kernel<C><<<grid, NT, smem, s>>>(op);
  DISPATCH_COT(cot, CALL)
#undef CALL
  if (rc) return rc;
  note_launch();
  return check_launch("conv_bwd_data");
}

int launch_conv_bwd_w(const SeistOp& op, cudaStream_t s, int sm_count) {
  const int gs_in = op.Cin / op.groups, gs_out = op.Cout / op.groups;
  int cot = pick_cot(gs_out, op.Cout, op.groups);
  if (cot > 8) cot = 8;
  const int co_tile = ROWS * cot;
  const int R = gs_in * op.k;
  const int ng_max = op.groups > 1 ? (co_tile + gs_out - 1) / gs_out + (co_tile % gs_out ? 1 : 0) : 1;
  int nq_max = (RT + op.k - 1) / op.k + 1;
  if (nq_max > gs_in) nq_max = gs_in;
  const int TLin = (PC - 1) * op.stride + op.k;
  const int TLp = TLin | 1;
  const size_t smem = sizeof(float) * ((size_t)PC * (co_tile + 4) + (size_t)ng_max * nq_max * TLp);
  const int gy = (op.Cout + co_tile - 1) / co_tile, gz = (R + RT - 1) / RT;
  const long tiles = (long)op.N * ((op.L_out + PC - 1) / PC);
  long gx = (4L * sm_count + gy * gz - 1) / (gy * gz);
  if (gx > tiles) gx = tiles;
  if (gx < 1) gx = 1;
  dim3 grid((unsigned)gx, gy, gz);
  int rc = 0;
#define CALL(C)                                                         \
  rc = set_smem(conv_bwd_w_kernel<C>, smem);                            \
  if (!rc) conv_bwd_w_kernel<C><<<grid, NT, smem, s>>>(op);
  switch (cot) {
    case 1: CALL(1); break;
    case 2: CALL(2); break;
    case 4: CALL(4); break;
    default: CALL(8); break;
  }
#undef CALL
  if (rc) return rc;
  note_launch();
  return check_launch("conv_bwd_w");
}

int launch_res_bwd(const SeistOp& op, cudaStream_t s) {
  dim3 grid((op.L_out + NT - 1) / NT, op.Cout, op.N);
  res_bwd_kernel<<<grid, NT, 0, s>>>(op);
  note_launch();
  return check_launch("res_bwd");
}

}  // namespace seist
