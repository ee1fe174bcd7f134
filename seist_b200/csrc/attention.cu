// Attention core of AttentionBlock (reference models/seist.py:381-388):
//   out[n, h*E+e, l] = sum_j softmax_j( (q[n,h,:,l]/sqrt(E)) . k[n,h,:,j] ) * v[n,h,e,j]
// q is (N, C, Lq), k/v are (N, C, Lk) with Lk = Lq / attn_aggr_ratio (128 for L = 8192 at every stage),
// so K and V of one (waveform, head) — at most 2 x 32 x 128 floats — sit in shared memory and each
// thread owns one query column in registers (flash-style online softmax, no (Lq x Lk) matrix in HBM).
// Lanes run along the sample axis: every global access is a coalesced channel-row segment.
#include "common.cuh"

namespace seist {

constexpr int ATT_NT = 128;
constexpr int KV_CHUNK = 128;

__device__ __forceinline__ float attn_keep(const SeistOp& op, uint64_t seed, int n, int h, int l, int j) {
  if (op.p_attn <= 0.f) return 1.f;
  const uint64_t idx = (((uint64_t)n * op.heads + h) * (uint64_t)op.L_out + l) * (uint64_t)op.L_in + j;
  return keep_scale(op.p_attn, seed, op.seed_attn, idx);
}

// ---- forward: grid (ceil(Lq/128), N*heads) --------------------------------------------------------
template <int E>
__global__ void __launch_bounds__(ATT_NT) att_fwd_kernel(const __grid_constant__ SeistOp op) {
  __shared__ float k_s[E][KV_CHUNK];
  __shared__ float v_s[E][KV_CHUNK];
  const int n = blockIdx.y / op.heads, h = blockIdx.y % op.heads;
  const int l = blockIdx.x * ATT_NT + threadIdx.x;
  const int Lq = op.L_out, Lk = op.L_in;
  const bool ok = l < Lq;
  const uint64_t seed = load_seed(op.step_seed);
  const float scale = rsqrtf((float)E);
  const SeistView &qv = op.in[0], &kv = op.in[1], &vv = op.in[2];

  float q[E], o[E];
#pragma unroll
  for (int e = 0; e < E; ++e) {
    q[e] = ok ? view_row(qv, n, h * E + e)[l] * scale : 0.f;
    o[e] = 0.f;
  }
  float m = -INFINITY, den = 0.f;
  for (int j0 = 0; j0 < Lk; j0 += KV_CHUNK) {
    const int jn = min(KV_CHUNK, Lk - j0);
    __syncthreads();
    for (int idx = threadIdx.x; idx < E * KV_CHUNK; idx += ATT_NT) {
      const int e = idx / KV_CHUNK, j = idx % KV_CHUNK;
      k_s[e][j] = j < jn ? view_row(kv, n, h * E + e)[j0 + j] : 0.f;
      v_s[e][j] = j < jn ? view_row(vv, n, h * E + e)[j0 + j] : 0.f;
    }
    __syncthreads();
    for (int j = 0; j < jn; ++j) {
      float s = 0.f;
#pragma unroll
      for (int e = 0; e < E; ++e) s = fmaf(q[e], k_s[e][j], s);
      if (s > m) {   // rescale only when the running maximum moves
        const float r = __expf(m - s);
        den *= r;
#pragma unroll
        for (int e = 0; e < E; ++e) o[e] *= r;
        m = s;
      }
      const float p = __expf(s - m);
      den += p;
      const float pd = p * (ok ? attn_keep(op, seed, n, h, l, j0 + j) : 1.f);
#pragma unroll
      for (int e = 0; e < E; ++e) o[e] = fmaf(pd, v_s[e][j], o[e]);
    }
  }
  if (ok) {
    const float inv = 1.f / den;
#pragma unroll
    for (int e = 0; e < E; ++e)
      op.out.x[((size_t)n * op.out.Ct + op.out.c0 + h * E + e) * (size_t)Lq + l] = o[e] * inv;
    if (op.lse) op.lse[((size_t)n * op.heads + h) * Lq + l] = m + __logf(den);
  }
}

// ---- backward w.r.t. q: grid (ceil(Lq/128), N*heads); also writes delta[l] = sum_e dO*O -------------
template <int E>
__global__ void __launch_bounds__(ATT_NT) att_bwd_q_kernel(const __grid_constant__ SeistOp op) {
  __shared__ float k_s[E][KV_CHUNK];
  __shared__ float v_s[E][KV_CHUNK];
  const int n = blockIdx.y / op.heads, h = blockIdx.y % op.heads;
  const int l = blockIdx.x * ATT_NT + threadIdx.x;
  const int Lq = op.L_out, Lk = op.L_in;
  const bool ok = l < Lq;
  const uint64_t seed = load_seed(op.step_seed);
  const float scale = rsqrtf((float)E);
  const SeistView &qv = op.in[0], &kv = op.in[1], &vv = op.in[2];

  float q[E], dO[E], dq[E];
  float delta = 0.f;
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const size_t off = ((size_t)n * op.out.Ct + op.out.c0 + h * E + e) * (size_t)Lq + l;
    q[e] = ok ? view_row(qv, n, h * E + e)[l] * scale : 0.f;
    dO[e] = ok ? op.out_dxd[off] : 0.f;
    delta = fmaf(dO[e], ok ? op.out.x[off] : 0.f, delta);
    dq[e] = 0.f;
  }
  const float lse = ok ? op.lse[((size_t)n * op.heads + h) * Lq + l] : 0.f;
  if (ok) op.delta[((size_t)n * op.heads + h) * Lq + l] = delta;
  for (int j0 = 0; j0 < Lk; j0 += KV_CHUNK) {
    const int jn = min(KV_CHUNK, Lk - j0);
    __syncthreads();
    for (int idx = threadIdx.x; idx < E * KV_CHUNK; idx += ATT_NT) {
      const int e = idx / KV_CHUNK, j = idx % KV_CHUNK;
      k_s[e][j] = j < jn ? view_row(kv, n, h * E + e)[j0 + j] : 0.f;
      v_s[e][j] = j < jn ? view_row(vv, n, h * E + e)[j0 + j] : 0.f;
    }
    __syncthreads();
    for (int j = 0; j < jn; ++j) {
      float s = 0.f, dp = 0.f;
#pragma unroll
      for (int e = 0; e < E; ++e) {
        s = fmaf(q[e], k_s[e][j], s);
        dp = fmaf(dO[e], v_s[e][j], dp);
      }
      const float p = __expf(s - lse);
      const float keep = ok ? attn_keep(op, seed, n, h, l, j0 + j) : 1.f;
      const float ds = p * (dp * keep - delta);
#pragma unroll
      for (int e = 0; e < E; ++e) dq[e] = fmaf(ds, k_s[e][j], dq[e]);
    }
  }
  if (ok && qv.g != nullptr) {
#pragma unroll
    for (int e = 0; e < E; ++e) {
      float* g = view_grad_row(qv, n, h * E + e) + l;
      const float val = dq[e] * scale;
      if (qv.accum) *g += val; else *g = val;
    }
  }
}

// ---- backward w.r.t. k, v: grid (ceil(Lk/128), N*heads); one thread per key, loop over queries ------
template <int E>
__global__ void __launch_bounds__(ATT_NT) att_bwd_kv_kernel(const __grid_constant__ SeistOp op) {
  constexpr int QC = 64;
  __shared__ float q_s[E][QC];
  __shared__ float do_s[E][QC];
  __shared__ float lse_s[QC];
  __shared__ float dl_s[QC];
  const int n = blockIdx.y / op.heads, h = blockIdx.y % op.heads;
  const int j = blockIdx.x * ATT_NT + threadIdx.x;
  const int Lq = op.L_out, Lk = op.L_in;
  const bool ok = j < Lk;
  const uint64_t seed = load_seed(op.step_seed);
  const float scale = rsqrtf((float)E);
  const SeistView &qv = op.in[0], &kv = op.in[1], &vv = op.in[2];

  float kk[E], vj[E], dk[E], dv[E];
#pragma unroll
  for (int e = 0; e < E; ++e) {
    kk[e] = ok ? view_row(kv, n, h * E + e)[j] : 0.f;
    vj[e] = ok ? view_row(vv, n, h * E + e)[j] : 0.f;
    dk[e] = 0.f;
    dv[e] = 0.f;
  }
  for (int l0 = 0; l0 < Lq; l0 += QC) {
    const int ln = min(QC, Lq - l0);
    __syncthreads();
    for (int idx = threadIdx.x; idx < E * QC; idx += ATT_NT) {
      const int e = idx / QC, i = idx % QC;
      const size_t off = ((size_t)n * op.out.Ct + op.out.c0 + h * E + e) * (size_t)Lq + l0 + i;
      q_s[e][i] = i < ln ? view_row(qv, n, h * E + e)[l0 + i] * scale : 0.f;
      do_s[e][i] = i < ln ? op.out_dxd[off] : 0.f;
    }
    for (int i = threadIdx.x; i < QC; i += ATT_NT) {
      lse_s[i] = i < ln ? op.lse[((size_t)n * op.heads + h) * Lq + l0 + i] : 0.f;
      dl_s[i] = i < ln ? op.delta[((size_t)n * op.heads + h) * Lq + l0 + i] : 0.f;
    }
    __syncthreads();
    for (int i = 0; i < ln; ++i) {
      float s = 0.f, dp = 0.f;
#pragma unroll
      for (int e = 0; e < E; ++e) {
        s = fmaf(q_s[e][i], kk[e], s);
        dp = fmaf(do_s[e][i], vj[e], dp);
      }
      const float p = __expf(s - lse_s[i]);
      const float keep = ok ? attn_keep(op, seed, n, h, l0 + i, j) : 1.f;
      const float pd = p * keep;
      const float ds = p * (dp * keep - dl_s[i]);
#pragma unroll
      for (int e = 0; e < E; ++e) {
        dv[e] = fmaf(pd, do_s[e][i], dv[e]);
        dk[e] = fmaf(ds, q_s[e][i], dk[e]);   // q_s already carries 1/sqrt(E)
      }
    }
  }
  if (ok) {
#pragma unroll
    for (int e = 0; e < E; ++e) {
      if (kv.g != nullptr) {
        float* g = view_grad_row(kv, n, h * E + e) + j;
        if (kv.accum) *g += dk[e]; else *g = dk[e];
      }
      if (vv.g != nullptr) {
        float* g = view_grad_row(vv, n, h * E + e) + j;
        if (vv.accum) *g += dv[e]; else *g = dv[e];
      }
    }
  }
}

#define ATT_DISPATCH(KERNEL, GRIDX)                                         \
  {                                                                         \
    const int E = op.Cout / op.heads;                                       \
    dim3 grid((GRIDX + ATT_NT - 1) / ATT_NT, op.N * op.heads);              \
    switch (E) {                                                            \
      case 8: KERNEL<8><<<grid, ATT_NT, 0, s>>>(op); break;                 \
      case 16: KERNEL<16><<<grid, ATT_NT, 0, s>>>(op); break;               \
      case 32: KERNEL<32><<<grid, ATT_NT, 0, s>>>(op); break;               \
      default: set_error("attention: head_dim must be 8, 16 or 32"); return -3; \
    }                                                                       \
    note_launch();                                                          \
  }

int launch_att_fwd(const SeistOp& op, cudaStream_t s) {
  ATT_DISPATCH(att_fwd_kernel, op.L_out)
  return check_launch("att_fwd");
}
int launch_att_bwd_q(const SeistOp& op, cudaStream_t s) {
  ATT_DISPATCH(att_bwd_q_kernel, op.L_out)
  return check_launch("att_bwd_q");
}
int launch_att_bwd_kv(const SeistOp& op, cudaStream_t s) {
  ATT_DISPATCH(att_bwd_kv_kernel, op.L_in)
  return check_launch("att_bwd_kv");
}

}  // namespace seist
