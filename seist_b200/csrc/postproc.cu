// GPU post-processing of the detection / picking probability traces: replaces the per-waveform CPU loop of the
// reference's training/postprocess.py (`_pick_phase` -> `_detect_peaks` :15-111,161-193; `_detect_event` -> obspy
// trigger_onset :114-158), which runs on the host for every waveform of every step (training/train.py:141), and the counter
// part of utils/metrics.py:141-193.  One CTA per (waveform, channel) row; the row (L <= 16384 samples) is staged in shared
// memory once; integer outputs are bit-identical to the numpy oracle (oracle/postprocess_ref.py).
#include "common.cuh"

namespace seist {

constexpr int PP_NT = 256;
constexpr int PP_MAXK = 8;

struct PPKey {          // (value, index): larger value first, equal values: larger index first (oracle tie rule)
  float v;
  int i;
};
__device__ __forceinline__ bool pp_better(PPKey a, PPKey b) { return a.v > b.v || (a.v == b.v && a.i > b.i); }

__device__ PPKey pp_block_argmax(PPKey k, PPKey* red_s) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    PPKey t;
    t.v = __shfl_xor_sync(0xffffffffu, k.v, o);
    t.i = __shfl_xor_sync(0xffffffffu, k.i, o);
    if (pp_better(t, k)) k = t;
  }
  __syncthreads();
  if (lane == 0) red_s[warp] = k;
  __syncthreads();
  PPKey r = red_s[0];
  for (int w = 1; w < PP_NT / 32; ++w)
    if (pp_better(red_s[w], r)) r = red_s[w];
  return r;
}

// phases: out[row][topk] int64 (sorted by sample index, padded with pad_value)
__global__ void __launch_bounds__(PP_NT) pick_phase_kernel(const float* __restrict__ prob, long long row_stride, long long n_stride,
                                                           int L, float mph, int mpd, int topk, long long pad_value,
                                                           long long* __restrict__ out) {
  extern __shared__ float pp_x[];                 // [L]
  __shared__ PPKey red_s[PP_NT / 32];
  __shared__ int sel_s[PP_MAXK];
  const int n = blockIdx.x, tid = threadIdx.x;
  const float* row = prob + (long long)n * n_stride + row_stride;
  for (int i = tid; i < L; i += PP_NT) pp_x[i] = row[i];
  __syncthreads();
  int nsel = 0;
  for (int r = 0; r < topk; ++r) {                 // the topk highest rising-edge peaks >= mph (postprocess.py:67-97)
    PPKey best;
    best.v = -INFINITY;
    best.i = -1;
    for (int i = tid + 1; i < L - 1; i += PP_NT) { // first / last samples cannot be peaks (:82-85)
      const float x = pp_x[i];
      const float prv = x - pp_x[i - 1], nxt = pp_x[i + 1] - x;
      if (nxt <= 0.f && prv > 0.f && x >= mph) {
        bool taken = false;
        for (int s = 0; s < nsel; ++s) taken = taken || sel_s[s] == i;
        PPKey k;
        k.v = x;
        k.i = i;
        if (!taken && pp_better(k, best)) best = k;
      }
    }
    best = pp_block_argmax(best, red_s);
    if (best.i < 0) break;
    if (tid == 0) sel_s[nsel] = best.i;
    ++nsel;
    __syncthreads();
  }
  if (tid == 0) {
    // greedy suppression in height order (:98-105), then back to index order (:107)
    bool keep[PP_MAXK];
    for (int i = 0; i < nsel; ++i) keep[i] = true;
    for (int i = 0; i < nsel; ++i) {
      if (!keep[i]) continue;
      for (int j = 0; j < nsel; ++j)
        if (j != i && sel_s[j] >= sel_s[i] - mpd && sel_s[j] <= sel_s[i] + mpd) keep[j] = false;
    }
    int kept[PP_MAXK], nk = 0;
    for (int i = 0; i < nsel; ++i)
      if (keep[i]) kept[nk++] = sel_s[i];
    for (int i = 1; i < nk; ++i) {                  // insertion sort by index
      const int v = kept[i];
      int j = i - 1;
      while (j >= 0 && kept[j] > v) { kept[j + 1] = kept[j]; --j; }
      kept[j + 1] = v;
    }
    for (int i = 0; i < topk; ++i) out[(long long)n * topk + i] = i < nk ? (long long)kept[i] : pad_value;
  }
}

// detections: out[row][2*topk] int64: the topk longest maximal runs of prob > thr as inclusive [on, off], longest first
// (equal lengths: the earlier run), padded with [1, 0]
__global__ void __launch_bounds__(PP_NT) detect_event_kernel(const float* __restrict__ prob, long long row_stride, long long n_stride,
                                                             int L, float thr, int topk, long long* __restrict__ out) {
  extern __shared__ float pp_x[];                 // [L] then int run_end[L] (end of the run starting at i, or -1)
  __shared__ PPKey red_s[PP_NT / 32];
  __shared__ int sel_s[PP_MAXK];
  const int n = blockIdx.x, tid = threadIdx.x;
  int* len_s = reinterpret_cast<int*>(pp_x + L);  // run length at run starts, 0 elsewhere
  const float* row = prob + (long long)n * n_stride + row_stride;
  for (int i = tid; i < L; i += PP_NT) pp_x[i] = row[i];
  __syncthreads();
  for (int i = tid; i < L; i += PP_NT) {
    int len = 0;
    if (pp_x[i] > thr && (i == 0 || !(pp_x[i - 1] > thr))) {
      int e = i;
      while (e + 1 < L && pp_x[e + 1] > thr) ++e;
      len = e - i + 1;
    }
    len_s[i] = len;
  }
  __syncthreads();
  int nsel = 0;
  for (int r = 0; r < topk; ++r) {
    // key: (run length, -start): longer first; equal lengths: the EARLIER run (python's stable sort keeps input order).
    // "nothing found" is length 0 (real runs have length >= 1)
    PPKey best;
    best.v = 0.f;
    best.i = -(1 << 30);
    for (int i = tid; i < L; i += PP_NT) {
      const int len = len_s[i];
      if (len > 0) {
        bool taken = false;
        for (int s = 0; s < nsel; ++s) taken = taken || sel_s[s] == i;
        PPKey k;
        k.v = (float)len;
        k.i = -i;
        if (!taken && pp_better(k, best)) best = k;
      }
    }
    best = pp_block_argmax(best, red_s);
    if (best.v <= 0.f) break;
    if (tid == 0) sel_s[nsel] = -best.i;
    ++nsel;
    __syncthreads();
  }
  if (tid == 0) {
    for (int i = 0; i < topk; ++i) {
      long long on = 1, off = 0;
      if (i < nsel) { on = sel_s[i]; off = on + len_s[sel_s[i]] - 1; }
      out[((long long)n * topk + i) * 2] = on;
      out[((long long)n * topk + i) * 2 + 1] = off;
    }
  }
}

// counters (utils/metrics.py:152-193,205-232), accumulated with atomics into a double vector:
//   pick: [0] data_size [1] tp [2] predp [3] possp [4] sum_res [5] sum_squ_res [6] sum_abs_res
//   det : [0] data_size [1] tp [2] predp [3] possp
__global__ void pick_counters_kernel(const long long* __restrict__ tgt, const long long* __restrict__ pred, int n, int num_samples,
                                     int t_thres, double* __restrict__ acc) {
  double v[7] = {0, 0, 0, 0, 0, 0, 0};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const long long t = tgt[i], p = pred[i];
    const bool pb = p >= 0 && p < num_samples, tb = t >= 0 && t < num_samples;
    const long long d = t - p;
    const bool tp = pb && tb && (d < 0 ? -d : d) <= t_thres;
    v[0] += 1;
    v[1] += tp;
    v[2] += pb;
    v[3] += tb;
    if (tp) { v[4] += (double)d; v[5] += (double)d * (double)d; v[6] += (double)(d < 0 ? -d : d); }
  }
  for (int k = 0; k < 7; ++k) {
    double s = v[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0 && s != 0.0) atomicAdd(&acc[k], s);
  }
}
__global__ void __launch_bounds__(PP_NT) det_counters_kernel(const long long* __restrict__ tgt, const long long* __restrict__ pred,
                                                             int kt, int kp, int num_samples, double* __restrict__ acc) {
  const int n = blockIdx.x;
  int tp = 0, pp = 0, tt = 0;
  for (int i = threadIdx.x; i < num_samples; i += PP_NT) {
    bool tb = false, pb = false;
    for (int j = 0; j < kt; ++j) tb = tb || (tgt[((long long)n * kt + j) * 2] <= i && i <= tgt[((long long)n * kt + j) * 2 + 1]);
    for (int j = 0; j < kp; ++j) pb = pb || (pred[((long long)n * kp + j) * 2] <= i && i <= pred[((long long)n * kp + j) * 2 + 1]);
    tp += tb && pb;
    pp += pb;
    tt += tb;
  }
  double v[3] = {(double)tp, (double)pp, (double)tt};
  for (int k = 0; k < 3; ++k) {
    double s = v[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0 && s != 0.0) atomicAdd(&acc[1 + k], s);
  }
  if (threadIdx.x == 0) atomicAdd(&acc[0], 1.0);
}

}  // namespace seist

using namespace seist;

extern "C" {

int seist_pick_phase(const float* prob, int64_t N, int32_t C, int32_t channel, int32_t L, float threshold, int32_t min_peak_dist,
                     int32_t topk, int64_t pad_value, int64_t* out, void* stream) {
  if (!prob || !out || N <= 0 || L < 3 || L > 16384 || topk < 1 || topk > PP_MAXK || min_peak_dist <= 1 || channel < 0 || channel >= C) {
    set_error("pick_phase: bad arguments (3 <= L <= 16384, 1 <= topk <= 8, min_peak_dist > 1)");
    return -1;
  }
  const size_t smem = sizeof(float) * (size_t)L;
  static bool attr = false;
  if (!attr) { cudaFuncSetAttribute(pick_phase_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024); attr = true; }
  pick_phase_kernel<<<(unsigned)N, PP_NT, smem, (cudaStream_t)stream>>>(prob, (long long)channel * L, (long long)C * L, L, threshold,
                                                                       min_peak_dist, topk, (long long)pad_value, (long long*)out);
  note_launch();
  return check_launch("pick_phase");
}

int seist_detect_event(const float* prob, int64_t N, int32_t C, int32_t channel, int32_t L, float threshold, int32_t topk,
                       int64_t* out, void* stream) {
  if (!prob || !out || N <= 0 || L < 1 || L > 16384 || topk < 1 || topk > PP_MAXK || channel < 0 || channel >= C) {
    set_error("detect_event: bad arguments (L <= 16384, 1 <= topk <= 8)");
    return -1;
  }
  const size_t smem = 2 * sizeof(float) * (size_t)L;
  static bool attr = false;
  if (!attr) { cudaFuncSetAttribute(detect_event_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024); attr = true; }
  detect_event_kernel<<<(unsigned)N, PP_NT, smem, (cudaStream_t)stream>>>(prob, (long long)channel * L, (long long)C * L, L, threshold,
                                                                         topk, (long long*)out);
  note_launch();
  return check_launch("detect_event");
}

int seist_pick_counters(const int64_t* targets, const int64_t* preds, int64_t n, int32_t num_samples, int32_t t_thres, double* acc,
                        void* stream) {
  if (!targets || !preds || !acc || n <= 0) { set_error("pick_counters: bad arguments"); return -1; }
  long g = (n + 255) / 256;
  if (g > 296) g = 296;
  pick_counters_kernel<<<(unsigned)g, 256, 0, (cudaStream_t)stream>>>((const long long*)targets, (const long long*)preds, (int)n,
                                                                     num_samples, t_thres, acc);
  note_launch();
  return check_launch("pick_counters");
}

int seist_det_counters(const int64_t* targets, const int64_t* preds, int64_t N, int32_t k_targets, int32_t k_preds,
                       int32_t num_samples, double* acc, void* stream) {
  if (!targets || !preds || !acc || N <= 0 || k_targets < 1 || k_preds < 1) { set_error("det_counters: bad arguments"); return -1; }
  det_counters_kernel<<<(unsigned)N, PP_NT, 0, (cudaStream_t)stream>>>((const long long*)targets, (const long long*)preds, k_targets,
                                                                      k_preds, num_samples, acc);
  note_launch();
  return check_launch("det_counters");
}

}  // extern "C"
