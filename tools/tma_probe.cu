// TMA (cp.async.bulk.tensor.3d) probe for the tcconv raw staging: which ways of passing the tensor map work, and do
// unaligned / negative / far-out-of-bounds coordinates behave as zero-filled boxes?
// build: nvcc -gencode arch=compute_100a,code=sm_100a -o tools/tma_probe tools/tma_probe.cu ; run: ./tools/tma_probe
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <cuda.h>
#include <cuda_runtime.h>

struct Maps { CUtensorMap m[3]; };
struct Pad { int a[151]; };

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ void do_load(const CUtensorMap* map, int c0, int c1, int c2, float* out, int rows) {
  extern __shared__ __align__(128) unsigned char raw[];
  unsigned char* base = raw + ((128u - (s32(raw) & 127u)) & 127u);
  float* dst = reinterpret_cast<float*>(base);
  uint64_t* bar = reinterpret_cast<uint64_t*>(base + 8 * rows * 4);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(bar)), "r"(8 * rows * 4) : "memory");
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(s32(dst)), "l"((uint64_t)map), "r"(c0), "r"(c1), "r"(c2), "r"(s32(bar)) : "memory");
  }
  uint32_t done = 0;
  for (int it = 0; it < (1 << 20) && !done; ++it)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(done) : "r"(s32(bar)), "r"(0u) : "memory");
  __syncthreads();
  for (int i = threadIdx.x; i < 8 * rows; i += blockDim.x) out[i] = done ? dst[i] : -12345.f;
}

__global__ void k_direct(const __grid_constant__ CUtensorMap map, int c0, int c1, int c2, float* out, int rows) { do_load(&map, c0, c1, c2, out, rows); }
__global__ void k_struct(const __grid_constant__ Pad pad, const __grid_constant__ Maps maps, int which, int c0, int c1, int c2, float* out, int rows) {
  do_load(&maps.m[which], c0, c1, c2, out, rows);
}
__global__ void k_global(const CUtensorMap* map, int c0, int c1, int c2, float* out, int rows) { do_load(map, c0, c1, c2, out, rows); }

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static void report(const char* title, const float* h, int rows, int L, int Ct, int c0, int c1, int n) {
  int bad = 0;
  for (int c = 0; c < 8; ++c)
    for (int r = 0; r < rows; ++r) {
      const int p = c0 + r, ch = c1 + c;
      const float want = (p >= 0 && p < L && ch >= 0 && ch < Ct) ? (float)(n * 1000000 + ch * 10000 + p) : 0.f;
      if (h[c * rows + r] != want) { if (bad < 4) printf("   mismatch c=%d r=%d got %g want %g\n", c, r, h[c * rows + r], want); ++bad; }
    }
  printf("%-58s : %s (%d mismatches)\n", title, bad ? "FAIL" : "ok", bad);
}

int main() {
  const int L = 1024, Ct = 48, N = 2, rows = 136;
  float* x; cudaMallocManaged(&x, sizeof(float) * L * Ct * N);
  for (int n = 0; n < N; ++n) for (int c = 0; c < Ct; ++c) for (int p = 0; p < L; ++p) x[((size_t)n * Ct + c) * L + p] = (float)(n * 1000000 + c * 10000 + p);
  float* out; cudaMallocManaged(&out, sizeof(float) * 8 * rows);
  void* p = nullptr; cudaDriverEntryPointQueryResult q;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  printf("entry point: err=%d q=%d ptr=%p\n", (int)e, (int)q, p);
  EncodeTiledFn fn = (EncodeTiledFn)p;
  Maps maps; memset(&maps, 0, sizeof(maps));
  const cuuint64_t dims[3] = {L, Ct, N};
  const cuuint64_t strides[2] = {(cuuint64_t)L * 4, (cuuint64_t)L * Ct * 4};
  const cuuint32_t box[3] = {rows, 8, 1};
  const cuuint32_t es[3] = {1, 1, 1};
  for (int i = 0; i < 3; ++i) {
    CUresult r = fn(&maps.m[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, x, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode %d -> %d\n", i, (int)r);
  }
  const size_t smem = 8 * rows * 4 + 64 + 256;
  struct { int c0, c1, n; const char* what; } cases[] = {
      {0, 0, 0, "aligned origin"}, {128, 8, 1, "aligned interior"}, {-3, 16, 1, "negative unaligned start"},
      {5, 40, 0, "unaligned start"}, {1000, 40, 1, "tail out of bounds"}, {0, 44, 1, "channel tail out of bounds"},
      {0, 1 << 20, 0, "channel far out of bounds"}};
  Pad pad; memset(&pad, 0, sizeof(pad));
  CUtensorMap* dmap; cudaMalloc(&dmap, sizeof(CUtensorMap)); cudaMemcpy(dmap, &maps.m[0], sizeof(CUtensorMap), cudaMemcpyHostToDevice);
  for (auto& cs : cases) {
    char t[128];
    k_direct<<<1, 64, smem>>>(maps.m[0], cs.c0, cs.c1, cs.n, out, rows);
    e = cudaDeviceSynchronize(); snprintf(t, sizeof(t), "direct param  / %s", cs.what);
    if (e != cudaSuccess) { printf("%s: CUDA error %s\n", t, cudaGetErrorString(e)); return 1; }
    report(t, out, rows, L, Ct, cs.c0, cs.c1, cs.n);
    k_struct<<<1, 64, smem>>>(pad, maps, 2, cs.c0, cs.c1, cs.n, out, rows);
    e = cudaDeviceSynchronize(); snprintf(t, sizeof(t), "struct param[2] / %s", cs.what);
    if (e != cudaSuccess) { printf("%s: CUDA error %s\n", t, cudaGetErrorString(e)); return 1; }
    report(t, out, rows, L, Ct, cs.c0, cs.c1, cs.n);
    k_global<<<1, 64, smem>>>(dmap, cs.c0, cs.c1, cs.n, out, rows);
    e = cudaDeviceSynchronize(); snprintf(t, sizeof(t), "global memory / %s", cs.what);
    if (e != cudaSuccess) { printf("%s: CUDA error %s\n", t, cudaGetErrorString(e)); return 1; }
    report(t, out, rows, L, Ct, cs.c0, cs.c1, cs.n);
  }
  return 0;
}
