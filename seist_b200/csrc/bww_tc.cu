// Weight gradient of 1x1 convolutions on the 5th-generation tensor cores (tcgen05 / TMEM), sm_100a only.
//
//   dW[co][ci] = sum_{n,l} gacc[co][n,l] * f(in[ci])[n,l]          (gradient of nn.Conv1d k=1 weights)
//
// GEMM view: D[M = 128 input channels (a tile of ci)][N = Cout] += A[M][K] * B[N][K]^T with K = SAMPLES.  In the
// (N, C, L) layout both operands are K-major as they are: a channel row is contiguous along the reduction
// axis, so a thread that holds a float4 of four consecutive samples writes exactly one 16-byte element of the
// UMMA canonical K-major layout.  The accumulator lives in TMEM for the whole persistent loop over
// (waveform, 32-sample chunk) tiles - there is NO per-tile epilogue: D is read once at the end and merged into
// dW with one float atomic per element per CTA.  Operands are split hi + lo (3 x kind::tf32 MMAs per K-step)
// so the result is fp32-accurate.  Two shared-memory stages: the global loads of chunk i+1 are in flight and
// the MMAs of chunk i run while chunk i+1 is transformed (BN-backward prologue / BN-apply + GELU) and stored.
#include "common.cuh"
#include "conv_common.cuh"

namespace seist {

constexpr int BT_NT = 256;
constexpr int BT_KC = 32;                 // samples per chunk = 4 UMMA K-steps
constexpr int BT_M = 128;                 // ci rows per CTA (UMMA M)
constexpr int BT_A_PART = (BT_KC / 8) * BT_M * 32;      // bytes of one precision part of the A stage (16 KB)
constexpr int BT_MAXQ = 4;                // max float4 work items per thread per operand per chunk

__device__ __forceinline__ uint32_t bt_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t bt_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFFu);
  d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  return d;     // layout_type 0: no swizzle, K-major canonical [k/8][row/8][(k%8)/4][row%8][k%4]
}
__device__ __forceinline__ void bt_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void bt_split4(const float4 v, float4& hi, float4& lo) {
  hi.x = __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u);
  hi.y = __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
  hi.z = __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u);
  hi.w = __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
  lo.x = v.x - hi.x;
  lo.y = v.y - hi.y;
  lo.z = v.z - hi.z;
  lo.w = v.w - hi.w;
}
__device__ __forceinline__ bool bt_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (int it = 0; it < (1 << 20) && !done; ++it) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}\n"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  }
  return done != 0;
}

__device__ int g_bt_err_dev = 0;

struct BtChan {
  const float* x;
  long long nstride;
  float sc, sh;
  int act, pad;
};
struct BtOut {
  float A, Bx, Cc, pad;
};

__device__ __noinline__ float4 bt_gelu4(float4 v) {
  v.x = gelu_f(v.x);
  v.y = gelu_f(v.y);
  v.z = gelu_f(v.z);
  v.w = gelu_f(v.w);
  return v;
}

// grid (persistent CTAs over sample chunks, ceil(Cin / 128))
__global__ void __launch_bounds__(BT_NT) bww_tc_kernel(const __grid_constant__ SeistOp op, const int N_pad, const int tmem_cols) {
  extern __shared__ __align__(16) unsigned char bt_raw[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int Cin = op.Cin, Cout = op.Cout, L = op.L_out;
  const int ci_base = blockIdx.y * BT_M;
  const int rows_a = min(BT_M, Cin - ci_base);
  const int b_part = (BT_KC / 8) * N_pad * 32;            // bytes of one precision part of the B stage
  const int stage_bytes = 2 * BT_A_PART + 2 * b_part;
  unsigned char* st0 = bt_raw + ((128u - (bt_smem_u32(bt_raw) & 127u)) & 127u);
  float* bias_s = reinterpret_cast<float*>(st0 + 2 * stage_bytes);           // [Cout] dbias partials
  uint64_t* bars = reinterpret_cast<uint64_t*>(bias_s + Cout + (Cout & 1));  // [2] stage-free barriers
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
  BtOut* oc_s = reinterpret_cast<BtOut*>(bars + 3);                           // [Cout]
  BtChan* ch_s = reinterpret_cast<BtChan*>(oc_s + Cout);                      // [rows_a]

  for (int i = tid; i < (2 * stage_bytes) / 16; i += BT_NT) reinterpret_cast<float4*>(st0)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int co = tid; co < Cout; co += BT_NT) {
    const OutGradCoef k = out_grad_coef(op, co);
    BtOut o;
    o.A = k.A;
    o.Bx = k.Bx;
    o.Cc = k.Cc;
    o.pad = 0.f;
    oc_s[co] = o;
    bias_s[co] = 0.f;
  }
  for (int r = tid; r < rows_a; r += BT_NT) {
    int cv;
    const int vi = resolve_view(op, ci_base + r, cv);
    const SeistView& vw = op.in[vi];
    BtChan c;
    c.x = vw.x + (size_t)(vw.c0 + cv) * vw.L;
    c.nstride = (long long)vw.Ct * vw.L;
    view_coef(op, vw, cv, c.sc, c.sh);
    c.act = vw.act;
    c.pad = 0;
    ch_s[r] = c;
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(bt_smem_u32(tmem_slot)),
                 "r"((uint32_t)tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bt_smem_u32(&bars[0])) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bt_smem_u32(&bars[1])) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N_pad >> 3) << 17) | ((uint32_t)(BT_M >> 4) << 24);

  const uint64_t seed = load_seed(op.step_seed);
  const bool has_bn = (op.out.bn >= 0 && op.out.g != nullptr);
  const bool need_x = has_bn || op.out_act == SEIST_OUT_SIGMOID;
  const int chunks_per_n = (L + BT_KC - 1) / BT_KC;
  const int total = op.N * chunks_per_n;
  const int items_a = rows_a * (BT_KC / 4), items_b = Cout * (BT_KC / 4);   // float4 work items per chunk

  // ---- register prefetch of one chunk -------------------------------------------------------------
  float4 pa[BT_MAXQ], pdx[BT_MAXQ], pdu[BT_MAXQ], px[BT_MAXQ];
  auto prefetch = [&](int tile) {
    const int n = tile / chunks_per_n;
    const int l0 = (tile - n * chunks_per_n) * BT_KC;
#pragma unroll
    for (int u = 0; u < BT_MAXQ; ++u) {
      const int idx = tid + u * BT_NT;
      const int row = idx >> 3, q = idx & 7;
      pa[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (idx < items_a && l0 + 4 * q < L) {
        const BtChan& c = ch_s[row];
        pa[u] = __ldg(reinterpret_cast<const float4*>(c.x + (long long)n * c.nstride + l0 + 4 * q));
      }
      pdx[u] = pdu[u] = px[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (idx < items_b && l0 + 4 * q < L) {
        const size_t off = ((size_t)n * op.out.Ct + op.out.c0 + row) * (size_t)L + l0 + 4 * q;
        if (op.out_dxd) pdx[u] = __ldg(reinterpret_cast<const float4*>(op.out_dxd + off));
        if (has_bn) pdu[u] = __ldg(reinterpret_cast<const float4*>(op.out.g + off));
        if (need_x) px[u] = __ldg(reinterpret_cast<const float4*>(op.out.x + off));
      }
    }
  };

  uint32_t parity[2] = {0u, 0u};
  int used[2] = {0, 0};                  // commits issued on each stage barrier
  bool failed = false, first = true;
  int it = 0;
  int tile = blockIdx.x;
  if (tile < total) prefetch(tile);
  for (; tile < total; tile += gridDim.x, ++it) {
    const int s = it & 1;
    const int n = tile / chunks_per_n;
    const int l0 = (tile - n * chunks_per_n) * BT_KC;
    unsigned char* a_hi = st0 + s * stage_bytes;
    unsigned char* a_lo = a_hi + BT_A_PART;
    unsigned char* b_hi = a_lo + BT_A_PART;
    unsigned char* b_lo = b_hi + b_part;
    if (used[s] > 0) {               // the MMAs that last read this stage must have completed
      if (!bt_wait(bt_smem_u32(&bars[s]), parity[s])) failed = true;
      parity[s] ^= 1;
    }
    const float pf = path_factor(op, seed, n) * alpha_factor(op, seed, n);
#pragma unroll
    for (int u = 0; u < BT_MAXQ; ++u) {
      const int idx = tid + u * BT_NT;
      const int row = idx >> 3, q = idx & 7;
      if (idx < items_a) {
        const BtChan& c = ch_s[row];
        float4 t = pa[u];
        t.x = fmaf(c.sc, t.x, c.sh);
        t.y = fmaf(c.sc, t.y, c.sh);
        t.z = fmaf(c.sc, t.z, c.sh);
        t.w = fmaf(c.sc, t.w, c.sh);
        if (c.act == SEIST_ACT_GELU) t = bt_gelu4(t);
        if (l0 + 4 * q >= L) t = make_float4(0.f, 0.f, 0.f, 0.f);
        float4 hi, lo;
        bt_split4(t, hi, lo);
        // element (row m, samples 4q..4q+3): K-block q/2, 8-row group m/8, 16-byte k-chunk q%2, row m%8
        const int o = (q >> 1) * (BT_M * 32) + (row >> 3) * 256 + (q & 1) * 128 + (row & 7) * 16;
        *reinterpret_cast<float4*>(a_hi + o) = hi;
        *reinterpret_cast<float4*>(a_lo + o) = lo;
      }
      if (idx < items_b) {
        const BtOut oc = oc_s[row];
        float4 g;
        g.x = pdx[u].x + fmaf(oc.A, pdu[u].x, fmaf(oc.Bx, px[u].x, oc.Cc));
        g.y = pdx[u].y + fmaf(oc.A, pdu[u].y, fmaf(oc.Bx, px[u].y, oc.Cc));
        g.z = pdx[u].z + fmaf(oc.A, pdu[u].z, fmaf(oc.Bx, px[u].z, oc.Cc));
        g.w = pdx[u].w + fmaf(oc.A, pdu[u].w, fmaf(oc.Bx, px[u].w, oc.Cc));
        if (op.out_act == SEIST_OUT_SIGMOID) {
          g.x *= px[u].x * (1.f - px[u].x);
          g.y *= px[u].y * (1.f - px[u].y);
          g.z *= px[u].z * (1.f - px[u].z);
          g.w *= px[u].w * (1.f - px[u].w);
        }
        g.x *= pf;
        g.y *= pf;
        g.z *= pf;
        g.w *= pf;
        const int lq = l0 + 4 * q;
        if (op.p_elem > 0.f && lq < L) {
          const uint64_t e = ((uint64_t)n * Cout + row) * (uint64_t)L + lq;
          const float4 kp = keep4(op.p_elem, seed, op.seed_elem, e);
          g.x *= kp.x;
          g.y *= kp.y;
          g.z *= kp.z;
          g.w *= kp.w;
        }
        if (lq >= L) g = make_float4(0.f, 0.f, 0.f, 0.f);
        if (blockIdx.y == 0 && op.dbias != nullptr) atomicAdd(&bias_s[row], (g.x + g.y) + (g.z + g.w));
        float4 hi, lo;
        bt_split4(g, hi, lo);
        const int o = (q >> 1) * (N_pad * 32) + (row >> 3) * 256 + (q & 1) * 128 + (row & 7) * 16;
        *reinterpret_cast<float4*>(b_hi + o) = hi;
        *reinterpret_cast<float4*>(b_lo + o) = lo;
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t ah = bt_smem_u32(a_hi), al = bt_smem_u32(a_lo), bh = bt_smem_u32(b_hi), bl = bt_smem_u32(b_lo);
#pragma unroll
      for (int kb = 0; kb < BT_KC / 8; ++kb) {
        const uint64_t dah = bt_desc(ah + kb * BT_M * 32, 128, 256), dal = bt_desc(al + kb * BT_M * 32, 128, 256);
        const uint64_t dbh = bt_desc(bh + kb * N_pad * 32, 128, 256), dbl = bt_desc(bl + kb * N_pad * 32, 128, 256);
        bt_mma(tmem_base, dah, dbh, idesc, (first && kb == 0) ? 0u : 1u);
        bt_mma(tmem_base, dal, dbh, idesc, 1u);
        bt_mma(tmem_base, dah, dbl, idesc, 1u);
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bt_smem_u32(&bars[s]))
                   : "memory");
    }
    first = false;
    used[s] += 1;
    if (tile + (int)gridDim.x < total) prefetch(tile + gridDim.x);
  }
  // ---- drain: all committed MMA batches must have completed before TMEM is read -------------------------
  for (int s = 0; s < 2; ++s) {
    if (used[s] > 0) {
      if (!bt_wait(bt_smem_u32(&bars[s]), parity[s])) failed = true;
    }
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (it > 0 && warp < 4) {
    // lane = ci row (TMEM lane 32*warp + lane); 16 output channels per tcgen05.ld
    const int r = 32 * warp + lane;
    const int ci = ci_base + r;
    for (int c0 = 0; c0 < N_pad; c0 += 16) {
      uint32_t rr[16];
      const uint32_t taddr = tmem_base + ((uint32_t)(32 * warp) << 16) + (uint32_t)c0;
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
          : "=r"(rr[0]), "=r"(rr[1]), "=r"(rr[2]), "=r"(rr[3]), "=r"(rr[4]), "=r"(rr[5]), "=r"(rr[6]), "=r"(rr[7]),
            "=r"(rr[8]), "=r"(rr[9]), "=r"(rr[10]), "=r"(rr[11]), "=r"(rr[12]), "=r"(rr[13]), "=r"(rr[14]), "=r"(rr[15])
          : "r"(taddr)
          : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        const int co = c0 + c;
        if (co < Cout && ci < Cin) atomicAdd(&op.dW[(size_t)co * Cin + ci], __uint_as_float(rr[c]));
      }
    }
  }
  __syncthreads();
  if (blockIdx.y == 0 && op.dbias != nullptr && it > 0)
    for (int co = tid; co < Cout; co += BT_NT) atomicAdd(&op.dbias[co], bias_s[co]);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)tmem_cols) : "memory");
  }
  if (failed) atomicExch(&g_bt_err_dev, 1);
}

bool bww_tc_eligible(const SeistOp& op) {
  if (op.k != 1 || op.stride != 1 || op.groups != 1 || op.pool > 1 || op.up_src_L > 0) return false;
  if ((op.L_out & 3) || op.Cout > 128 || op.Cout < 8 || op.Cin < 8) return false;
  for (int i = 0; i < op.n_in; ++i)
    if (op.in[i].L != op.L_out) return false;
  // register prefetch budget: BT_MAXQ float4 items per thread per operand per chunk
  if (op.Cout * (BT_KC / 4) > BT_MAXQ * BT_NT) return false;
  return true;
}

int launch_bww_tc(const SeistOp& op, cudaStream_t s, int sm_count) {
  const int N_pad = (op.Cout + 15) & ~15;
  int cols = 32;
  while (cols < N_pad) cols <<= 1;
  const size_t b_part = (size_t)(BT_KC / 8) * N_pad * 32;
  const size_t stage = 2 * (size_t)BT_A_PART + 2 * b_part;
  const int rows = op.Cin < BT_M ? op.Cin : BT_M;
  const size_t smem = 2 * stage + sizeof(float) * (op.Cout + 2) + 64 + sizeof(BtOut) * op.Cout + sizeof(BtChan) * rows + 256;
  static size_t max_set = 0;
  if (smem > max_set) {
    cudaError_t e = cudaFuncSetAttribute(bww_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smem < 49152 ? 49152 : smem));
    if (e != cudaSuccess) return (int)e;
    max_set = smem;
  }
  const int gy = (op.Cin + BT_M - 1) / BT_M;
  const long tiles = (long)op.N * ((op.L_out + BT_KC - 1) / BT_KC);
  long gx = (2L * sm_count + gy - 1) / gy;
  if (gx > tiles) gx = tiles;
  if (gx < 1) gx = 1;
  bww_tc_kernel<<<dim3((unsigned)gx, gy), BT_NT, smem, s>>>(op, N_pad, cols);
  note_launch();
  return check_launch("bww_tc");
}

int bww_tc_error_flag() {
  int v = 0;
  cudaMemcpyFromSymbol(&v, g_bt_err_dev, sizeof(int));
  return v;
}

}  // namespace seist
