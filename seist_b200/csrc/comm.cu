// Data-parallel exchange over NVLink peer memory (SeistComm, include/seist_b200.h).
//
// The reference converts every BatchNorm to torch.nn.SyncBatchNorm under DDP (training/train.py:374): one small NCCL
// collective per BatchNorm per direction (230 per seist_m_dpk step) plus the bucketed gradient all-reduce (:369).  Here
// every rank maps every peer's statistic / gradient buffers (symmetric memory) and
//   * the BN_PREPARE kernel itself sums the peers' partial statistics (one-shot all-reduce: <= 2 KB read per peer)
//     right where the coefficient table is computed - no collective call, no host involvement, graph capturable;
//   * the gradient all-reduce is one kernel reading every peer's flat gradient buffer over NVLink.
// Barrier: a per-lane epoch counter; rank r publishes epoch e with a system-scope release store into word
// [lane][r] of every peer's signal pad and spins (bounded) with acquire loads on its own pad.
#include "common.cuh"

namespace seist {

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ double ld_relaxed_sys_f64(const double* p) {
  double v;
  asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ float4 ld_relaxed_sys_f32x4(const float* p) {
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}

// all threads of the CTA call this; returns after every rank has entered the same exchange of `lane`
__device__ void comm_barrier(const SeistComm* comm, int lane) {
  __shared__ uint32_t epoch_s;
  const int tid = threadIdx.x;
  __syncthreads();
  if (tid == 0) {
    const uint32_t e = comm->epoch[lane] + 1u;
    comm->epoch[lane] = e;
    epoch_s = e;
    __threadfence_system();            // this rank's partial sums (earlier kernels of the stream) are visible to the peers
    for (int p = 0; p < comm->world; ++p)
      if (p != comm->rank) st_release_sys(comm->sig_peer[p] + lane * SEIST_MAX_WORLD + comm->rank, e);
  }
  __syncthreads();
  if (tid < comm->world && tid != comm->rank) {
    const uint32_t e = epoch_s;
    const uint32_t* w = comm->sig_peer[comm->rank] + lane * SEIST_MAX_WORLD + tid;
    bool ok = false;
    for (long it = 0; it < (1L << 23); ++it) {          // bounded (~10 s): a dead peer must not hang the GPU
      if ((int32_t)(ld_acquire_sys(w) - e) >= 0) { ok = true; break; }
      if ((it & 255) == 255) __nanosleep(200);
    }
    if (!ok) *comm->err = 1;
  }
  __syncthreads();
}

// BN_PREPARE with the cross-rank statistic sum fused in: entries [bn_lo, bn_lo + n_bn), ONE CTA
__global__ void __launch_bounds__(256) bn_prepare_xchg_kernel(const SeistBN* tab, int bn_lo, int n_bn, const SeistComm* comm, int fwd) {
  comm_barrier(comm, fwd ? 0 : 1);
  const int world = comm->world;
  for (int b = 0; b < n_bn; ++b) {
    const SeistBN& e = tab[bn_lo + b];
    if (e.is_chained || !e.use_batch) continue;
    const double* acc = fwd ? e.stat_acc : e.gstat_acc;
    double* red = fwd ? e.stat : e.gstat;
    const ptrdiff_t off = acc - (fwd ? comm->stat_peer[comm->rank] : comm->gstat_peer[comm->rank]);
    for (int i = threadIdx.x; i < 2 * e.C; i += blockDim.x) {
      // all peer loads in flight at once (one NVLink round trip, not world - 1), then a fixed-order sum: bit-identical on
      // every rank
      double v[SEIST_MAX_WORLD];
#pragma unroll
      for (int p = 0; p < SEIST_MAX_WORLD; ++p)
        v[p] = p < world ? ld_relaxed_sys_f64((fwd ? comm->stat_peer[p] : comm->gstat_peer[p]) + off + i) : 0.0;
      double s = 0.0;
#pragma unroll
      for (int p = 0; p < SEIST_MAX_WORLD; ++p) s += v[p];
      red[i] = s;
    }
  }
  __syncthreads();
  for (int b = 0; b < n_bn; ++b) {
    const int bn = bn_lo + b;
    const SeistBN& e = tab[bn];
    if (e.is_chained) continue;
    for (int c = threadIdx.x; c < e.C; c += blockDim.x) {
      float* k = e.coef + 8 * (size_t)c;
      if (fwd) {
        bn_fwd_coef(tab, bn, c, k[0], k[1]);
        bn_khat_coef(tab, bn, c, k[2], k[3]);
      } else {
        bn_bwd_coef(tab, bn, c, k[4], k[5], k[6]);
      }
    }
  }
}

int launch_bn_prepare_xchg(const SeistOp& op, bool fwd, cudaStream_t s) {
  bn_prepare_xchg_kernel<<<1, 256, 0, s>>>(op.bn_table, op.bn_lo, op.n_bn, op.comm, fwd ? 1 : 0);
  note_launch();
  return check_launch("bn_prepare_xchg");
}

__global__ void __launch_bounds__(32) comm_barrier_kernel(const SeistComm* comm, int lane) { comm_barrier(comm, lane); }

__global__ void __launch_bounds__(256) comm_allreduce_kernel(const SeistComm* comm, float* __restrict__ out, int64_t numel) {
  const int world = comm->world;
  const int64_t nq = numel >> 2;
  for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < nq; q += (int64_t)gridDim.x * blockDim.x) {
    float4 v[SEIST_MAX_WORLD];                             // every peer's load in flight before the first use
#pragma unroll
    for (int p = 0; p < SEIST_MAX_WORLD; ++p)
      v[p] = p < world ? ld_relaxed_sys_f32x4(comm->grad_peer[p] + 4 * q) : make_float4(0.f, 0.f, 0.f, 0.f);
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int p = 0; p < SEIST_MAX_WORLD; ++p) {            // fixed order: every rank computes the same bits
      s.x += v[p].x; s.y += v[p].y; s.z += v[p].z; s.w += v[p].w;
    }
    *reinterpret_cast<float4*>(out + 4 * q) = s;
  }
}

}  // namespace seist

using namespace seist;

extern "C" {

uint64_t seist_sizeof_comm(void) { return sizeof(SeistComm); }

int seist_comm_barrier(const SeistComm* comm, int32_t lane, void* stream) {
  if (comm == nullptr || lane < 0 || lane >= SEIST_SIG_LANES) { set_error("comm_barrier: bad arguments"); return -1; }
  comm_barrier_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(comm, lane);
  note_launch();
  return check_launch("comm_barrier");
}

int seist_comm_allreduce(const SeistComm* comm, int32_t world, float* out, int64_t numel, void* stream) {
  if (comm == nullptr || out == nullptr || numel <= 0 || (numel & 3) || world < 1 || world > SEIST_MAX_WORLD) {
    set_error("comm_allreduce: bad arguments (numel must be a multiple of 4)");
    return -1;
  }
  cudaStream_t s = (cudaStream_t)stream;
  comm_barrier_kernel<<<1, 32, 0, s>>>(comm, 2);           // every rank's gradients are complete
  note_launch();
  long g = (numel / 4 + 255) / 256;
  if (g > 296) g = 296;
  comm_allreduce_kernel<<<(unsigned)g, 256, 0, s>>>(comm, out, numel);
  note_launch();
  comm_barrier_kernel<<<1, 32, 0, s>>>(comm, 3);           // every peer has finished reading this rank's buffer
  note_launch();
  return check_launch("comm_allreduce");
}

}  // extern "C"
