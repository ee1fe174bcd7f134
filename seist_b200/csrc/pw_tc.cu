// 1x1 convolution forward on the 5th-generation tensor cores (tcgen05 / TMEM), sm_100a only.
//
//   out[co][p] = sum_ci W[co][ci] * f(in[ci][p])            (reference nn.Conv1d k=1, models/seist.py:86,107,...)
//
// GEMM view per 128-sample tile:  D[M = 128 samples][N = Cout] += A[M][K = Cin] * B[N][K]^T
//   A = the consumer view (BatchNorm-apply / GELU evaluated ONCE per element by the staging threads): each
//       thread owns one sample, reads its 32 channels with coalesced loads and writes them 4 channels at a
//       time (16-byte, bank-conflict-free stores) into the UMMA canonical K-major (no-swizzle) layout;
//   B = the weights, same K-major canonical layout   (both layouts verified word-by-word on hardware with
//       tools/tc_probe.cu);
//   D = fp32 accumulators in tensor memory (TMEM): lane = sample, column = output channel.
// Precision: kind::tf32 keeps 10 mantissa bits, the parity bar is 1e-3 against an fp32 CPU forward through
// ~50 layers, so every operand is split hi + lo (hi = top 19 bits, lo = x - hi) and three MMAs are issued:
// Ahi*Bhi + Alo*Bhi + Ahi*Blo (error ~2^-21).  One elected thread issues the MMAs and commits them to an
// mbarrier; the four warps then read their 32 TMEM lanes with tcgen05.ld and run the usual epilogue (bias,
// dropout, residuals, BatchNorm statistics) with one sample per lane, i.e. 128-byte coalesced stores.
#include "common.cuh"
#include "conv_common.cuh"

namespace seist {

constexpr int TC_NT = 128;
constexpr int TC_M = 128;            // samples per tile == UMMA M
constexpr int TC_KC = 32;            // reduction channels per chunk (4 UMMA K-steps of 8 tf32)
constexpr int TC_A_BYTES = (TC_KC / 8) * 4096;           // one precision part of the A chunk

__device__ __forceinline__ uint32_t tc_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// UMMA shared-memory matrix descriptor (cute/arch/mma_sm100_desc.hpp::SmemDescriptor)
__device__ __forceinline__ uint64_t tc_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFFu);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;                 // descriptor version (Blackwell)
  d |= (uint64_t)layout << 61;            // 0 = no swizzle, 2 = SWIZZLE_128B
  return d;
}

__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}

__device__ __forceinline__ void tc_split(float x, float& hi, float& lo) {
  hi = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
  lo = x - hi;
}

__device__ __forceinline__ bool tc_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (int it = 0; it < (1 << 20) && !done; ++it) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}\n"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  }
  return done != 0;
}

__device__ int g_tc_err_dev = 0;     // set if an mbarrier wait ever timed out (bounded spin: never hangs the GPU)

struct TcChan {          // reduction channel of the consumer view, resolved once per CTA
  const float* x;        // row base for n = 0
  long long nstride;
  float sc, sh;
  int act, pad;
};

// 16 values per lane -> lane l returns the warp-wide sum of value (l & 15)  (16 shuffles instead of 80)
__device__ __forceinline__ float tc_reduce16(float (&v)[16], int lane) {
#pragma unroll
  for (int half = 8, off = 16; half >= 1; half >>= 1, off >>= 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (i < half) {
        const float send = up ? v[i] : v[i + half];
        const float keep = up ? v[i + half] : v[i];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
      }
    }
  }
  return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 1);
}
// which of the 16 values lane `lane` holds after tc_reduce16
__device__ __forceinline__ int tc_reduce16_index(int lane) {
  return ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
}

__device__ __noinline__ float tc_gelu(float x) { return gelu_f(x); }     // shared body: keeps the kernel I-cache sized

// F_ELEM: element dropout in the epilogue; F_RES: residual views; F_GELU: some input view applies GELU
template <bool F_ELEM, bool F_RES, bool F_GELU>
__global__ void __launch_bounds__(TC_NT) pw_tc_fwd_kernel(const __grid_constant__ SeistOp op, const int N_pad,
                                                          const int tmem_cols, const int b_resident) {
  extern __shared__ __align__(16) unsigned char tc_raw[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int Cin = op.Cin, Cout = op.Cout, L = op.L_out;
  const int nchunks = (Cin + TC_KC - 1) / TC_KC;
  const int b_chunk_bytes = (TC_KC / 8) * N_pad * 32;          // one precision part of one K-chunk of B
  // carve-up
  unsigned char* a_hi = tc_raw + ((1024u - (tc_smem_u32(tc_raw) & 1023u)) & 1023u);
  unsigned char* a_lo = a_hi + TC_A_BYTES;
  unsigned char* b_hi = a_lo + TC_A_BYTES;
  const int b_parts = b_resident ? nchunks : 1;
  unsigned char* b_lo = b_hi + b_parts * b_chunk_bytes;
  float* ep_s = reinterpret_cast<float*>(b_lo + b_parts * b_chunk_bytes);   // bias, a_sc, a_sh, b_sc, b_sh [5][Cout]
  float* red_s = ep_s + 5 * Cout;                                            // [4 warps][2*Cout]
  uint64_t* bar = reinterpret_cast<uint64_t*>(red_s + 8 * Cout + ((13 * Cout) & 1));
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  TcChan* ch_s = reinterpret_cast<TcChan*>(bar + 2);                         // [Cin]

  for (int ci = tid; ci < Cin; ci += TC_NT) {
    int cv;
    const int vi = resolve_view(op, ci, cv);
    const SeistView& vw = op.in[vi];
    TcChan c;
    c.x = vw.x + (size_t)(vw.c0 + cv) * vw.L;
    c.nstride = (long long)vw.Ct * vw.L;
    view_coef(op, vw, cv, c.sc, c.sh);
    c.act = vw.act;
    c.pad = 0;
    ch_s[ci] = c;
  }
  for (int co = tid; co < Cout; co += TC_NT) {
    float b = 0.f, asc = 1.f, ash = 0.f, bsc = 1.f, bsh = 0.f;
    if (op.bias) b = op.bias[co];
    if (op.res_a.C > 0) view_coef(op, op.res_a, co, asc, ash);
    if (op.res_b.C > 0) view_coef(op, op.res_b, co, bsc, bsh);
    ep_s[co] = b;
    ep_s[Cout + co] = asc;
    ep_s[2 * Cout + co] = ash;
    ep_s[3 * Cout + co] = bsc;
    ep_s[4 * Cout + co] = bsh;
  }
  for (int i = tid; i < 8 * Cout; i += TC_NT) red_s[i] = 0.f;

  // weights -> canonical K-major, hi / lo parts: [chunk][k/8][n/8][(k%8)/4][n%8][k%4]
  auto stage_b = [&](int chunk, int slot) {
    const int k0 = chunk * TC_KC;
    for (int idx = tid; idx < N_pad * TC_KC; idx += TC_NT) {
      const int k = idx % TC_KC, nn = idx / TC_KC;
      float w = 0.f;
      if (nn < Cout && k0 + k < Cin) w = op.W[(size_t)nn * Cin + k0 + k];
      float hi, lo;
      tc_split(w, hi, lo);
      const int off = slot * b_chunk_bytes + (k >> 3) * (N_pad * 32) + (nn >> 3) * 256 + ((k & 7) >> 2) * 128 + (nn & 7) * 16 +
                      (k & 3) * 4;
      *reinterpret_cast<float*>(b_hi + off) = hi;
      *reinterpret_cast<float*>(b_lo + off) = lo;
    }
  };
  if (b_resident)
    for (int c = 0; c < nchunks; ++c) stage_b(c, c);

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc_smem_u32(tmem_slot)),
                 "r"((uint32_t)tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(tc_smem_u32(bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t bar_addr = tc_smem_u32(bar);
  uint32_t parity = 0;
  bool pending = false;         // an MMA batch has been committed and not yet waited for
  bool failed = false;

  // instruction descriptor: D=f32, A=B=tf32, A and B K-major, N = N_pad, M = 128
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N_pad >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);

  const uint64_t seed = load_seed(op.step_seed);
  const bool stats = (op.out.bn >= 0) && op.bn_table[op.out.bn >= 0 ? op.out.bn : 0].use_batch;
  const int tiles_per_n = (L + TC_M - 1) / TC_M;
  const int total = op.N * tiles_per_n;
  const int m = tid;                                   // this thread's sample inside the tile

  // prefetch of one (tile, chunk) step: raw channel values of sample m into registers
  float v[TC_KC];
  auto prefetch = [&](int tile, int chunk) {
    const int n = tile / tiles_per_n;
    const int lm = (tile - n * tiles_per_n) * TC_M + m;
    const int k0 = chunk * TC_KC;
#pragma unroll
    for (int r = 0; r < TC_KC; ++r) {
      v[r] = 0.f;
      if (k0 + r < Cin && lm < L) {
        const TcChan& c = ch_s[k0 + r];
        v[r] = __ldg(c.x + (long long)n * c.nstride + lm);
      }
    }
  };

  int tile = blockIdx.x, chunk = 0;
  if (tile < total) prefetch(tile, 0);
  while (tile < total) {
    const int n = tile / tiles_per_n;
    const int l0 = (tile - n * tiles_per_n) * TC_M;
    const int k0 = chunk * TC_KC;
    // the previous MMA batch reads the A (and streamed B) buffers: wait for it before overwriting them
    if (pending) {
      if (!tc_wait(bar_addr, parity)) failed = true;
      parity ^= 1;
      pending = false;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    // ---- transform + split + store this thread's sample: canonical K-major [k/8][m/8][(k%8)/4][m%8][k%4] ----
    {
      const bool inb = (l0 + m) < L;
#pragma unroll
      for (int r4 = 0; r4 < TC_KC; r4 += 4) {
        float t[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int r = r4 + j;
          t[j] = 0.f;
          if (k0 + r < Cin && inb) {
            const TcChan& c = ch_s[k0 + r];
            float u = fmaf(c.sc, v[r], c.sh);
            if (F_GELU && c.act == SEIST_ACT_GELU) u = tc_gelu(u);
            t[j] = u;
          }
        }
        float4 hi, lo;
        tc_split(t[0], hi.x, lo.x);
        tc_split(t[1], hi.y, lo.y);
        tc_split(t[2], hi.z, lo.z);
        tc_split(t[3], hi.w, lo.w);
        const int off = (r4 >> 3) * 4096 + (m >> 3) * 256 + ((r4 & 7) >> 2) * 128 + (m & 7) * 16;
        *reinterpret_cast<float4*>(a_hi + off) = hi;
        *reinterpret_cast<float4*>(a_lo + off) = lo;
      }
    }
    if (!b_resident) stage_b(chunk, 0);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> async proxy (tensor core)
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t a_hi_u = tc_smem_u32(a_hi), a_lo_u = tc_smem_u32(a_lo);
      const int slot = b_resident ? chunk : 0;
      const uint32_t b_hi_u = tc_smem_u32(b_hi) + slot * b_chunk_bytes, b_lo_u = tc_smem_u32(b_lo) + slot * b_chunk_bytes;
#pragma unroll
      for (int kb = 0; kb < TC_KC / 8; ++kb) {
        if (k0 + kb * 8 < Cin) {
          const uint64_t ah = tc_desc(a_hi_u + kb * 4096, 128, 256, 0);
          const uint64_t al = tc_desc(a_lo_u + kb * 4096, 128, 256, 0);
          const uint64_t bh = tc_desc(b_hi_u + kb * N_pad * 32, 128, 256, 0);
          const uint64_t bl = tc_desc(b_lo_u + kb * N_pad * 32, 128, 256, 0);
          tc_mma_tf32(tmem_base, ah, bh, idesc, (chunk > 0 || kb > 0) ? 1u : 0u);
          tc_mma_tf32(tmem_base, al, bh, idesc, 1u);
          tc_mma_tf32(tmem_base, ah, bl, idesc, 1u);
        }
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_addr) : "memory");
    }
    pending = true;
    // ---- next step: issue its global loads now so they fly while the tensor core works -------------------
    const bool last_chunk = (chunk + 1 == nchunks);
    const int next_tile = last_chunk ? tile + gridDim.x : tile;
    const int next_chunk = last_chunk ? 0 : chunk + 1;
    if (next_tile < total) prefetch(next_tile, next_chunk);

    if (last_chunk) {
      if (!tc_wait(bar_addr, parity)) failed = true;
      parity ^= 1;
      pending = false;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      // ---- epilogue: lane = sample (TMEM lane 32*warp + lane), 16 output channels per tcgen05.ld ----------
      const int l = l0 + m;
      const bool ok = l < L;
      const float pf = path_factor(op, seed, n), af = alpha_factor(op, seed, n);
      float* optr = op.out.x + ((size_t)n * op.out.Ct + op.out.c0) * (size_t)L + l;
      const float* ra = F_RES && op.res_a.C > 0 ? op.res_a.x + ((size_t)n * op.res_a.Ct + op.res_a.c0) * (size_t)L + l : nullptr;
      const float* rb = F_RES && op.res_b.C > 0 ? op.res_b.x + ((size_t)n * op.res_b.Ct + op.res_b.c0) * (size_t)L + l : nullptr;
#pragma unroll 1
      for (int c0 = 0; c0 < Cout; c0 += 16) {
        uint32_t rr[16];
        const uint32_t taddr = tmem_base + ((uint32_t)(32 * warp) << 16) + (uint32_t)c0;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
            : "=r"(rr[0]), "=r"(rr[1]), "=r"(rr[2]), "=r"(rr[3]), "=r"(rr[4]), "=r"(rr[5]), "=r"(rr[6]), "=r"(rr[7]),
              "=r"(rr[8]), "=r"(rr[9]), "=r"(rr[10]), "=r"(rr[11]), "=r"(rr[12]), "=r"(rr[13]), "=r"(rr[14]), "=r"(rr[15])
            : "r"(taddr)
            : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        float s1[16], s2[16];
        // element dropout: the 4 lanes of a sample quad share one hash per channel (common.cuh::keep4), so each
        // lane hashes ONE channel of every group of 4 and the group exchanges the results with shuffles
        uint64_t hq[4] = {0ull, 0ull, 0ull, 0ull};
        if (F_ELEM) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const int ck = min(c0 + 4 * g + (lane & 3), Cout - 1);
            hq[g] = rng_u64(seed, op.seed_elem, (((uint64_t)n * Cout + ck) * (uint64_t)L + (ok ? l : 0)) >> 2);
          }
        }
        const uint32_t thr = drop_threshold(op.p_elem);
        const float keep_s = 1.0f / (1.0f - op.p_elem);
#pragma unroll
        for (int c = 0; c < 16; ++c) {
          const int co = c0 + c;
          float val = 0.f;
          float kf = 1.f;
          if (F_ELEM) {
            const uint64_t hh = __shfl_sync(0xffffffffu, hq[c >> 2], (lane & ~3) | (c & 3));
            kf = ((uint32_t)(hh >> (16 * (l & 3))) & 0xFFFFu) >= thr ? keep_s : 0.f;
          }
          if (ok && co < Cout) {
            val = (__uint_as_float(rr[c]) + ep_s[co]) * pf;
            if (F_ELEM) val *= kf;
            if (F_RES) {
              if (ra) val += fmaf(ep_s[Cout + co], ra[(size_t)co * L], ep_s[2 * Cout + co]);
              val *= af;
              if (rb) val += fmaf(ep_s[3 * Cout + co], rb[(size_t)co * L], ep_s[4 * Cout + co]);
            } else {
              val *= af;
            }
            if (op.out_act == SEIST_OUT_SIGMOID) val = sigmoid_f(val);
            optr[(size_t)co * L] = val;
          }
          s1[c] = val;
          s2[c] = val * val;
        }
        if (stats) {
          const float t1 = tc_reduce16(s1, lane), t2 = tc_reduce16(s2, lane);
          const int co = c0 + tc_reduce16_index(lane);
          if ((lane & 1) == 0 && co < Cout) {          // single writer per (warp, channel): deterministic
            red_s[warp * 2 * Cout + 2 * co] += t1;
            red_s[warp * 2 * Cout + 2 * co + 1] += t2;
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncthreads();       // all TMEM reads of this tile are done before the next tile's MMAs overwrite it
    }
    tile = next_tile;
    chunk = next_chunk;
  }

  if (stats) {
    __syncthreads();
    const SeistBN& e = op.bn_table[op.out.bn];
    for (int i = tid; i < 2 * Cout; i += TC_NT) {
      const float s = (red_s[i] + red_s[2 * Cout + i]) + (red_s[4 * Cout + i] + red_s[6 * Cout + i]);
      atomicAdd(&e.stat_acc[(i & 1) * e.C + op.out.bn_c0 + (i >> 1)], (double)s);
    }
  }
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)tmem_cols) : "memory");
  }
  if (failed) atomicExch(&g_tc_err_dev, 1);
}

// ------------------------------------------------------------------------------------------------
bool pw_tc_eligible(const SeistOp& op) {
  if (op.k != 1 || op.stride != 1 || op.groups != 1 || op.pool > 1 || op.up_src_L > 0) return false;
  if ((op.L_out & 3) || op.Cout > 128 || op.Cout < 8) return false;
  for (int i = 0; i < op.n_in; ++i)
    if (op.in[i].L != op.L_out) return false;
  return true;
}

int launch_pw_tc_fwd(const SeistOp& op, cudaStream_t s, int sm_count) {
  const int N_pad = (op.Cout + 15) & ~15;
  int cols = 32;
  while (cols < N_pad) cols <<= 1;
  const int nchunks = (op.Cin + TC_KC - 1) / TC_KC;
  const size_t b_chunk = (size_t)(TC_KC / 8) * N_pad * 32;
  const int b_resident = (2 * nchunks * b_chunk <= 64 * 1024) ? 1 : 0;
  const size_t smem = 2 * (size_t)TC_A_BYTES + 2 * (b_resident ? nchunks : 1) * b_chunk + sizeof(float) * (13 * (size_t)op.Cout + 2) +
                      32 + sizeof(TcChan) * (size_t)op.Cin + 1024;
  const long tiles = (long)op.N * ((op.L_out + TC_M - 1) / TC_M);
  long g = 4L * sm_count;
  if (g > tiles) g = tiles;
  const unsigned grid = (unsigned)(g < 1 ? 1 : g);
  const bool f_elem = op.p_elem > 0.f, f_res = op.res_a.C > 0 || op.res_b.C > 0;
  bool f_gelu = false;
  for (int i = 0; i < op.n_in; ++i) f_gelu = f_gelu || op.in[i].act == SEIST_ACT_GELU;
  int rc = 0;
#define TC_LAUNCH(E, R, G)                                                                                           \
  {                                                                                                                   \
    static size_t max_set = 0;                                                                                        \
    if (smem > max_set) {                                                                                             \
      cudaError_t e = cudaFuncSetAttribute(pw_tc_fwd_kernel<E, R, G>, cudaFuncAttributeMaxDynamicSharedMemorySize,    \
                                           (int)(smem < 49152 ? 49152 : smem));                                       \
      if (e != cudaSuccess) rc = (int)e;                                                                              \
      max_set = smem;                                                                                                 \
    }                                                                                                                 \
    if (!rc) pw_tc_fwd_kernel<E, R, G><<<grid, TC_NT, smem, s>>>(op, N_pad, cols, b_resident);                        \
  }
  const int sel = (f_elem ? 4 : 0) | (f_res ? 2 : 0) | (f_gelu ? 1 : 0);
  switch (sel) {
    case 0: TC_LAUNCH(false, false, false) break;
    case 1: TC_LAUNCH(false, false, true) break;
    case 2: TC_LAUNCH(false, true, false) break;
    case 3: TC_LAUNCH(false, true, true) break;
    case 4: TC_LAUNCH(true, false, false) break;
    case 5: TC_LAUNCH(true, false, true) break;
    case 6: TC_LAUNCH(true, true, false) break;
    default: TC_LAUNCH(true, true, true) break;
  }
#undef TC_LAUNCH
  if (rc) return rc;
  note_launch();
  return check_launch("pw_tc_fwd");
}

int pw_tc_error_flag() {
  int v = 0;
  cudaMemcpyFromSymbol(&v, g_tc_err_dev, sizeof(int));
  return v;
}

}  // namespace seist
