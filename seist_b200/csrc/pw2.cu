// Pointwise (k = 1) convolutions, second generation: transform once, contract from shared memory.
//
// A CTA (4 warps) owns tiles of 128 consecutive samples.  The reduction operand tile — the consumer view
// with BatchNorm-apply / GELU evaluated (forward), or the output gradient with the BatchNorm-backward
// prologue, sigmoid' and dropout mask folded in (backward) — is loaded with 16-byte global loads (16 per
// thread in flight), transformed ONCE and parked in shared memory as (channel, sample) rows; every
// output-channel pass then contracts it against k-major weights in shared memory.  Lane = sample quad,
// warp = 16 output channels: per reduction channel one conflict-free LDS.128 (activations) + 4 broadcast
// LDS.128 (weights) feed 64 FMAs.  The expensive part of the prologue (erf-GELU, RNG) is therefore paid once
// per element instead of once per 16 output channels, and the input is read from HBM/L2 once per CTA.
// BatchNorm statistics of the result are reduced warp -> shared (float) -> one fp64 atomic per channel per CTA.
#include "common.cuh"
#include "conv_common.cuh"

namespace seist {

constexpr int P2_NT = 128;
constexpr int P2_TL = 128;     // samples per tile
constexpr int P2_KC = 64;      // reduction channels per chunk
constexpr int P2_NC = 64;      // output channels per pass
constexpr int P2_PITCH = P2_TL + 4;

__device__ __forceinline__ float4 p2_ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 p2_lds4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void p2_st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

struct P2Chan {          // a channel of a view, resolved once per CTA
  const float* x;        // row base for n = 0
  float* g;              // gradient row base for n = 0 or nullptr
  long long nstride;
  float sc, sh, mu, istd;
  int act, bn, bnc, accum;
};

__device__ __forceinline__ P2Chan p2_make_chan(const SeistOp& op, int ci, bool want_khat) {
  int cv;
  const int vi = resolve_view(op, ci, cv);
  const SeistView& v = op.in[vi];
  P2Chan c;
  c.x = v.x + (size_t)(v.c0 + cv) * v.L;
  c.g = v.g ? v.g + (size_t)(v.c0 + cv) * v.L : nullptr;
  c.nstride = (long long)v.Ct * v.L;
  view_coef(op, v, cv, c.sc, c.sh);
  c.mu = 0.f;
  c.istd = 0.f;
  if (want_khat && v.bn >= 0) view_khat(op, v, cv, c.mu, c.istd);
  c.act = v.act;
  c.bn = v.bn;
  c.bnc = v.bn_c0 + cv;
  c.accum = v.accum;
  return c;
}

// acc[c][.] += sum_kk w_s[kk][wbase + c] * tile[kk][quad]   (the shared-memory contraction core)
__device__ __forceinline__ void p2_contract(const float* tile, const float* w_s, int kc, int lane, int wbase,
                                            float4 (&acc)[16]) {
  const float* tp = tile + 4 * lane;
  const float* wp = w_s + wbase;
#pragma unroll 4
  for (int kk = 0; kk < kc; ++kk) {
    const float4 a = p2_lds4(tp + kk * P2_PITCH);
    const float4 w0 = p2_lds4(wp + kk * P2_NC), w1 = p2_lds4(wp + kk * P2_NC + 4);
    const float4 w2 = p2_lds4(wp + kk * P2_NC + 8), w3 = p2_lds4(wp + kk * P2_NC + 12);
    const float w[16] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w, w2.x, w2.y, w2.z, w2.w, w3.x, w3.y, w3.z, w3.w};
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      acc[c].x = fmaf(w[c], a.x, acc[c].x);
      acc[c].y = fmaf(w[c], a.y, acc[c].y);
      acc[c].z = fmaf(w[c], a.z, acc[c].z);
      acc[c].w = fmaf(w[c], a.w, acc[c].w);
    }
  }
}

// ================================================================================================
// forward
// ================================================================================================
__global__ void __launch_bounds__(P2_NT) pw2_fwd_kernel(const __grid_constant__ SeistOp op) {
  extern __shared__ __align__(16) unsigned char p2_raw[];
  const int Cin = op.Cin, Cout = op.Cout, L = op.L_out;
  float* in_s = reinterpret_cast<float*>(p2_raw);                 // [KC][PITCH]
  float* w_s = in_s + P2_KC * P2_PITCH;                           // [KC][NC]
  float* red_s = w_s + P2_KC * P2_NC;                             // [2*Cout] stats (float partials of this CTA)
  float* ep_s = red_s + 2 * Cout;                                 // bias, a_sc, a_sh, b_sc, b_sh  [5][Cout]
  P2Chan* ch_s = reinterpret_cast<P2Chan*>(ep_s + 5 * Cout + ((5 * Cout + 2 * Cout) & 1));   // [Cin], 8-byte aligned
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  for (int ci = tid; ci < Cin; ci += P2_NT) ch_s[ci] = p2_make_chan(op, ci, false);
  for (int co = tid; co < Cout; co += P2_NT) {
    float b = 0.f, asc = 1.f, ash = 0.f, bsc = 1.f, bsh = 0.f;
    if (op.bias) b = op.bias[co];
    if (op.res_a.C > 0) view_coef(op, op.res_a, co, asc, ash);
    if (op.res_b.C > 0) view_coef(op, op.res_b, co, bsc, bsh);
    ep_s[co] = b;
    ep_s[Cout + co] = asc;
    ep_s[2 * Cout + co] = ash;
    ep_s[3 * Cout + co] = bsc;
    ep_s[4 * Cout + co] = bsh;
    red_s[2 * co] = 0.f;
    red_s[2 * co + 1] = 0.f;
  }
  const bool w_resident = Cin <= P2_KC && Cout <= P2_NC;      // one weight tile: stage it once per CTA
  if (w_resident) {
    for (int idx = tid; idx < P2_KC * P2_NC; idx += P2_NT) {
      const int kk = idx / P2_NC, c = idx - kk * P2_NC;
      w_s[idx] = (kk < Cin && c < Cout) ? op.W[(size_t)c * Cin + kk] : 0.f;
    }
  }
  __syncthreads();

  const uint64_t seed = load_seed(op.step_seed);
  const bool stats = (op.out.bn >= 0) && op.bn_table[op.out.bn >= 0 ? op.out.bn : 0].use_batch;
  const int tiles_per_n = (L + P2_TL - 1) / P2_TL;
  const int total = op.N * tiles_per_n;
  const int Lsrc = op.in[0].L;

  for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
    const int n = tile / tiles_per_n;
    const int l0 = (tile - n * tiles_per_n) * P2_TL;
    const int l = l0 + 4 * lane;
    const bool ok = l < L;
    const float pf = path_factor(op, seed, n), af = alpha_factor(op, seed, n);
    for (int co0 = 0; co0 < Cout; co0 += P2_NC) {
      float4 acc[16];
#pragma unroll
      for (int c = 0; c < 16; ++c) acc[c] = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int k0 = 0; k0 < Cin; k0 += P2_KC) {
        const int kc = min(P2_KC, Cin - k0);
        const bool restage_in = !(co0 > 0 && Cin <= P2_KC);
        if (restage_in) {
          if (op.pool > 1) {
            // pooled input (LocalAwareAggregationBlock, models/seist.py:80-81,93): avg + max over `pool` samples
            for (int idx = tid; idx < kc * P2_TL; idx += P2_NT) {
              const int row = idx / P2_TL, pos = idx - row * P2_TL;
              const P2Chan& c = ch_s[k0 + row];
              const int p = l0 + pos;
              float v = 0.f;
              if (p < L) {
                const float* xr = c.x + (long long)n * c.nstride;
                const int s0 = p * op.pool, cnt = min(op.pool, Lsrc - s0);
                float sum = 0.f, mx = -INFINITY;
                for (int i = 0; i < cnt; ++i) {
                  float u = fmaf(c.sc, xr[s0 + i], c.sh);
                  if (c.act == SEIST_ACT_GELU) u = gelu_f(u);
                  sum += u;
                  mx = fmaxf(mx, u);
                }
                v = sum / (float)cnt + mx;
              }
              in_s[row * P2_PITCH + pos] = v;
            }
          } else {
            for (int idx0 = tid; idx0 < kc * 32; idx0 += P2_NT * 4) {
              float4 v[4];
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                const int idx = idx0 + u * P2_NT;
                const int row = idx >> 5, q = idx & 31;
                v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (idx < kc * 32 && l0 + 4 * q < L) {
                  const P2Chan& c = ch_s[k0 + row];
                  v[u] = p2_ldg4(c.x + (long long)n * c.nstride + l0 + 4 * q);
                }
              }
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                const int idx = idx0 + u * P2_NT;
                if (idx < kc * 32) {
                  const int row = idx >> 5, q = idx & 31;
                  const P2Chan& c = ch_s[k0 + row];
                  float4 t = v[u];
                  t.x = fmaf(c.sc, t.x, c.sh);
                  t.y = fmaf(c.sc, t.y, c.sh);
                  t.z = fmaf(c.sc, t.z, c.sh);
                  t.w = fmaf(c.sc, t.w, c.sh);
                  if (c.act == SEIST_ACT_GELU) {
                    t.x = gelu_f(t.x);
                    t.y = gelu_f(t.y);
                    t.z = gelu_f(t.z);
                    t.w = gelu_f(t.w);
                  }
                  if (l0 + 4 * q >= L) t = make_float4(0.f, 0.f, 0.f, 0.f);
                  p2_st4(in_s + row * P2_PITCH + 4 * q, t);
                }
              }
            }
          }
        }
        if (!w_resident) {
          for (int idx = tid; idx < P2_KC * P2_NC; idx += P2_NT) {
            const int kk = idx / P2_NC, c = idx - kk * P2_NC;     // conflict-free shared stores; W is cache-resident
            const int co = co0 + c;
            w_s[idx] = (kk < kc && co < Cout) ? op.W[(size_t)co * Cin + k0 + kk] : 0.f;
          }
        }
        __syncthreads();
        p2_contract(in_s, w_s, kc, lane, warp * 16, acc);
        __syncthreads();
      }
      // ---- epilogue of this pass -----------------------------------------------------------------
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        const int co = co0 + warp * 16 + c;
        if (co >= Cout) break;          // warp-uniform
        float s1 = 0.f, s2 = 0.f;
        if (ok) {
          float4 r = acc[c];
          const float b = ep_s[co];
          r.x = (r.x + b) * pf;
          r.y = (r.y + b) * pf;
          r.z = (r.z + b) * pf;
          r.w = (r.w + b) * pf;
          if (op.p_elem > 0.f) {
            const uint64_t e = ((uint64_t)n * Cout + co) * (uint64_t)L + l;
            r.x *= keep_scale(op.p_elem, seed, op.seed_elem, e);
            r.y *= keep_scale(op.p_elem, seed, op.seed_elem, e + 1);
            r.z *= keep_scale(op.p_elem, seed, op.seed_elem, e + 2);
            r.w *= keep_scale(op.p_elem, seed, op.seed_elem, e + 3);
          }
          if (op.res_a.C > 0) {
            const float4 a = p2_ldg4(op.res_a.x + ((size_t)n * op.res_a.Ct + op.res_a.c0 + co) * (size_t)L + l);
            const float sc = ep_s[Cout + co], sh = ep_s[2 * Cout + co];
            r.x += fmaf(sc, a.x, sh);
            r.y += fmaf(sc, a.y, sh);
            r.z += fmaf(sc, a.z, sh);
            r.w += fmaf(sc, a.w, sh);
          }
          r.x *= af;
          r.y *= af;
          r.z *= af;
          r.w *= af;
          if (op.res_b.C > 0) {
            const float4 a = p2_ldg4(op.res_b.x + ((size_t)n * op.res_b.Ct + op.res_b.c0 + co) * (size_t)L + l);
            const float sc = ep_s[3 * Cout + co], sh = ep_s[4 * Cout + co];
            r.x += fmaf(sc, a.x, sh);
            r.y += fmaf(sc, a.y, sh);
            r.z += fmaf(sc, a.z, sh);
            r.w += fmaf(sc, a.w, sh);
          }
          if (op.out_act == SEIST_OUT_SIGMOID) {
            r.x = sigmoid_f(r.x);
            r.y = sigmoid_f(r.y);
            r.z = sigmoid_f(r.z);
            r.w = sigmoid_f(r.w);
          }
          p2_st4(op.out.x + ((size_t)n * op.out.Ct + op.out.c0 + co) * (size_t)L + l, r);
          s1 = (r.x + r.y) + (r.z + r.w);
          s2 = fmaf(r.x, r.x, r.y * r.y) + fmaf(r.z, r.z, r.w * r.w);
        }
        if (stats) {
          s1 = warp_sum(s1);
          s2 = warp_sum(s2);
          if (lane == 0) {           // this warp is the only writer of channel `co` inside the CTA
            red_s[2 * co] += s1;
            red_s[2 * co + 1] += s2;
          }
        }
      }
    }
  }
  if (stats) {
    __syncthreads();
    const SeistBN& e = op.bn_table[op.out.bn];
    for (int i = tid; i < 2 * Cout; i += P2_NT)
      atomicAdd(&e.stat[(i & 1) * e.C + op.out.bn_c0 + (i >> 1)], (double)red_s[i]);
  }
}

// ================================================================================================
// backward (data): d in[ci] = sum_co W[co][ci] gacc[co]
// ================================================================================================
struct P2Out {
  float A, Bx, Cc;
};

__global__ void __launch_bounds__(P2_NT) pw2_bwd_data_kernel(const __grid_constant__ SeistOp op) {
  extern __shared__ __align__(16) unsigned char p2_raw[];
  const int Cin = op.Cin, Cout = op.Cout, L = op.L_out;
  float* g_s = reinterpret_cast<float*>(p2_raw);                  // [KC][PITCH]  (gacc rows)
  float* w_s = g_s + P2_KC * P2_PITCH;                            // [KC][NC]     (W^T chunk: [co][ci])
  float* red_s = w_s + P2_KC * P2_NC;                             // [2*Cin]
  P2Out* oc_s = reinterpret_cast<P2Out*>(red_s + 2 * Cin);        // [Cout]
  P2Chan* ch_s = reinterpret_cast<P2Chan*>(oc_s + Cout + ((3 * Cout + 2 * Cin) & 1));   // [Cin] targets
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  for (int ci = tid; ci < Cin; ci += P2_NT) {
    ch_s[ci] = p2_make_chan(op, ci, true);
    red_s[2 * ci] = 0.f;
    red_s[2 * ci + 1] = 0.f;
  }
  for (int co = tid; co < Cout; co += P2_NT) {
    const OutGradCoef k = out_grad_coef(op, co);
    P2Out o;
    o.A = k.A;
    o.Bx = k.Bx;
    o.Cc = k.Cc;
    oc_s[co] = o;
  }
  const bool w_resident = Cout <= P2_KC && Cin <= P2_NC;
  if (w_resident) {
    for (int idx = tid; idx < P2_KC * P2_NC; idx += P2_NT) {
      const int kk = idx / P2_NC, c = idx - kk * P2_NC;     // kk = co, c = ci (coalesced along W's rows)
      w_s[idx] = (kk < Cout && c < Cin) ? op.W[(size_t)kk * Cin + c] : 0.f;
    }
  }
  __syncthreads();

  const uint64_t seed = load_seed(op.step_seed);
  const bool has_bn = (op.out.bn >= 0 && op.out.g != nullptr);
  const bool need_x = has_bn || op.out_act == SEIST_OUT_SIGMOID;
  const int tiles_per_n = (L + P2_TL - 1) / P2_TL;
  const int total = op.N * tiles_per_n;

  for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
    const int n = tile / tiles_per_n;
    const int l0 = (tile - n * tiles_per_n) * P2_TL;
    const int l = l0 + 4 * lane;
    const bool ok = l < L;
    const float pf = path_factor(op, seed, n) * alpha_factor(op, seed, n);
    for (int ci0 = 0; ci0 < Cin; ci0 += P2_NC) {
      float4 acc[16];
#pragma unroll
      for (int c = 0; c < 16; ++c) acc[c] = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int k0 = 0; k0 < Cout; k0 += P2_KC) {
        const int kc = min(P2_KC, Cout - k0);
        const bool restage = !(ci0 > 0 && Cout <= P2_KC);
        if (restage) {
          for (int idx0 = tid; idx0 < kc * 32; idx0 += P2_NT * 2) {
            float4 dx[2], du[2], xo[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              const int idx = idx0 + u * P2_NT;
              const int row = idx >> 5, q = idx & 31;
              dx[u] = du[u] = xo[u] = make_float4(0.f, 0.f, 0.f, 0.f);
              if (idx < kc * 32 && l0 + 4 * q < L) {
                const size_t off = ((size_t)n * op.out.Ct + op.out.c0 + k0 + row) * (size_t)L + l0 + 4 * q;
                if (op.out_dxd) dx[u] = p2_ldg4(op.out_dxd + off);
                if (has_bn) du[u] = p2_ldg4(op.out.g + off);
                if (need_x) xo[u] = p2_ldg4(op.out.x + off);
              }
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              const int idx = idx0 + u * P2_NT;
              if (idx < kc * 32) {
                const int row = idx >> 5, q = idx & 31;
                const int co = k0 + row;
                const P2Out o = oc_s[co];
                float4 gv;
                gv.x = dx[u].x + fmaf(o.A, du[u].x, fmaf(o.Bx, xo[u].x, o.Cc));
                gv.y = dx[u].y + fmaf(o.A, du[u].y, fmaf(o.Bx, xo[u].y, o.Cc));
                gv.z = dx[u].z + fmaf(o.A, du[u].z, fmaf(o.Bx, xo[u].z, o.Cc));
                gv.w = dx[u].w + fmaf(o.A, du[u].w, fmaf(o.Bx, xo[u].w, o.Cc));
                if (op.out_act == SEIST_OUT_SIGMOID) {
                  gv.x *= xo[u].x * (1.f - xo[u].x);
                  gv.y *= xo[u].y * (1.f - xo[u].y);
                  gv.z *= xo[u].z * (1.f - xo[u].z);
                  gv.w *= xo[u].w * (1.f - xo[u].w);
                }
                gv.x *= pf;
                gv.y *= pf;
                gv.z *= pf;
                gv.w *= pf;
                const int lq = l0 + 4 * q;
                if (op.p_elem > 0.f && lq < L) {
                  const uint64_t e = ((uint64_t)n * Cout + co) * (uint64_t)L + lq;
                  gv.x *= keep_scale(op.p_elem, seed, op.seed_elem, e);
                  gv.y *= keep_scale(op.p_elem, seed, op.seed_elem, e + 1);
                  gv.z *= keep_scale(op.p_elem, seed, op.seed_elem, e + 2);
                  gv.w *= keep_scale(op.p_elem, seed, op.seed_elem, e + 3);
                }
                if (lq >= L) gv = make_float4(0.f, 0.f, 0.f, 0.f);
                p2_st4(g_s + row * P2_PITCH + 4 * q, gv);
              }
            }
          }
        }
        if (!w_resident) {
          for (int idx = tid; idx < P2_KC * P2_NC; idx += P2_NT) {
            const int kk = idx / P2_NC, c = idx - kk * P2_NC;
            const int co = k0 + kk, ci = ci0 + c;
            w_s[idx] = (kk < kc && ci < Cin) ? op.W[(size_t)co * Cin + ci] : 0.f;
          }
        }
        __syncthreads();
        p2_contract(g_s, w_s, kc, lane, warp * 16, acc);
        __syncthreads();
      }
      // ---- deposit ---------------------------------------------------------------------------------
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        const int ci = ci0 + warp * 16 + c;
        if (ci >= Cin) break;           // warp-uniform
        const P2Chan& ch = ch_s[ci];
        if (ch.g == nullptr) continue;  // warp-uniform
        float s1 = 0.f, s2 = 0.f;
        if (ok) {
          float4 gg = acc[c];
          const long long off = (long long)n * ch.nstride + l;
          float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
          if (ch.act == SEIST_ACT_GELU || ch.bn >= 0) x = p2_ldg4(ch.x + off);
          if (ch.act == SEIST_ACT_GELU) {
            gg.x *= gelu_grad_f(fmaf(ch.sc, x.x, ch.sh));
            gg.y *= gelu_grad_f(fmaf(ch.sc, x.y, ch.sh));
            gg.z *= gelu_grad_f(fmaf(ch.sc, x.z, ch.sh));
            gg.w *= gelu_grad_f(fmaf(ch.sc, x.w, ch.sh));
          }
          if (ch.bn >= 0) {
            s1 = (gg.x + gg.y) + (gg.z + gg.w);
            s2 = fmaf(gg.x, (x.x - ch.mu) * ch.istd, gg.y * ((x.y - ch.mu) * ch.istd)) +
                 fmaf(gg.z, (x.z - ch.mu) * ch.istd, gg.w * ((x.w - ch.mu) * ch.istd));
          }
          float* gp = ch.g + off;
          if (ch.accum) {
            const float4 old = *reinterpret_cast<const float4*>(gp);
            gg.x += old.x;
            gg.y += old.y;
            gg.z += old.z;
            gg.w += old.w;
          }
          p2_st4(gp, gg);
        }
        if (ch.bn >= 0) {
          s1 = warp_sum(s1);
          s2 = warp_sum(s2);
          if (lane == 0) {
            red_s[2 * ci] += s1;
            red_s[2 * ci + 1] += s2;
          }
        }
      }
    }
  }
  __syncthreads();
  for (int i = tid; i < 2 * Cin; i += P2_NT) {
    const P2Chan& ch = ch_s[i >> 1];
    if (ch.g != nullptr && ch.bn >= 0) {
      const SeistBN& e = op.bn_table[ch.bn];
      atomicAdd(&e.gstat[(i & 1) * e.C + ch.bnc], (double)red_s[i]);
    }
  }
}

// ================================================================================================
// launchers
// ================================================================================================
bool pw2_fwd_eligible(const SeistOp& op) {
  if (op.k != 1 || op.stride != 1 || op.groups != 1 || op.up_src_L > 0) return false;
  if (op.L_out & 3) return false;
  if (op.pool > 1 && op.n_in != 1) return false;
  return true;
}
bool pw2_bwd_eligible(const SeistOp& op) {
  return op.k == 1 && op.stride == 1 && op.groups == 1 && op.up_src_L == 0 && op.pool <= 1 && (op.L_out & 3) == 0;
}

static int p2_grid(const SeistOp& op, int sm_count) {
  const long tiles = (long)op.N * ((op.L_out + P2_TL - 1) / P2_TL);
  long g = 4L * sm_count;
  if (g > tiles) g = tiles;
  return (int)(g < 1 ? 1 : g);
}

int launch_pw2_fwd(const SeistOp& op, cudaStream_t s, int sm_count) {
  const size_t smem = sizeof(float) * (P2_KC * P2_PITCH + P2_KC * P2_NC + 7 * (size_t)op.Cout + 2) +
                      sizeof(P2Chan) * (size_t)op.Cin + 16;
  static size_t max_set = 0;
  if (smem > 48 * 1024 && smem > max_set) {
    cudaError_t e = cudaFuncSetAttribute(pw2_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    max_set = smem;
  }
  pw2_fwd_kernel<<<p2_grid(op, sm_count), P2_NT, smem, s>>>(op);
  note_launch();
  return check_launch("pw2_fwd");
}

int launch_pw2_bwd_data(const SeistOp& op, cudaStream_t s, int sm_count) {
  const size_t smem = sizeof(float) * (P2_KC * P2_PITCH + P2_KC * P2_NC + 2 * (size_t)op.Cin + 2) +
                      sizeof(P2Out) * (size_t)op.Cout + sizeof(P2Chan) * (size_t)op.Cin + 16;
  static size_t max_set = 0;
  if (smem > 48 * 1024 && smem > max_set) {
    cudaError_t e = cudaFuncSetAttribute(pw2_bwd_data_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    max_set = smem;
  }
  pw2_bwd_data_kernel<<<p2_grid(op, sm_count), P2_NT, smem, s>>>(op);
  note_launch();
  return check_launch("pw2_bwd_data");
}

}  // namespace seist
