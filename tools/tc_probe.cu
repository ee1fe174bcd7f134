// Layout discovery for tcgen05.mma kind::tf32 operands: which raw shared-memory word feeds which (row, k)?
// One CTA per probed word: A-probe (B = all ones): D[m][*] lights the row m the word belongs to;
// B-probe (A = all ones): D[*][n] lights the column; K-probe: one-hot A word x one-hot B word.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -o tc_probe tools/tc_probe.cu ; run: ./tc_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int N = 32;                 // UMMA N
constexpr int A_WORDS = 1024;         // 128 x 8 tf32
constexpr int B_WORDS = N * 8;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t mk_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFFu);
  d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;
  return d;
}

struct Cfg { uint32_t a_lbo, a_sbo, a_layout, a_major, b_lbo, b_sbo, b_layout, b_major; };

// mode 0: A one-hot at word `blockIdx.x`, B all ones -> out[blk] = bitmask summary of lit rows (first lit row, count)
// mode 1: B one-hot at word blk, A all ones -> first lit column, count
// mode 2: A one-hot at a_list[blk / nb], B one-hot at b_list[blk % nb] -> D[m0][n0]
__global__ void __launch_bounds__(128) probe(Cfg cfg, int mode, const int* a_list, const int* b_list, int nb, int m0, int n0,
                                             int* out_first, int* out_count, float* out_val) {
  extern __shared__ __align__(16) unsigned char raw[];
  unsigned char* base = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
  float* A = reinterpret_cast<float*>(base);
  float* B = reinterpret_cast<float*>(base + 4096);
  uint64_t* bar = reinterpret_cast<uint64_t*>(base + 4096 + 1024);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int blk = blockIdx.x;
  int a_hot = -1, b_hot = -1;
  if (mode == 0) a_hot = blk;
  if (mode == 1) b_hot = blk;
  if (mode == 2) { a_hot = a_list[blk / nb]; b_hot = b_list[blk % nb]; }
  for (int i = tid; i < A_WORDS; i += 128) A[i] = (mode == 1) ? 1.f : (i == a_hot ? 1.f : 0.f);
  for (int i = tid; i < B_WORDS; i += 128) B[i] = (mode == 0) ? 1.f : (i == b_hot ? 1.f : 0.f);
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(32u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *slot;
  if (tid == 0) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (cfg.a_major << 15) | (cfg.b_major << 16) | ((uint32_t)(N >> 3) << 17) | (8u << 24);
    const uint64_t ad = mk_desc(smem_u32(A), cfg.a_lbo, cfg.a_sbo, cfg.a_layout);
    const uint64_t bd = mk_desc(smem_u32(B), cfg.b_lbo, cfg.b_sbo, cfg.b_layout);
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
                 ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(0u) : "memory");
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
  }
  uint32_t done = 0;
  for (int it = 0; it < (1 << 18) && !done; ++it)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(done) : "r"(smem_u32(bar)), "r"(0u) : "memory");
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  __shared__ float D[128][N + 1];
  for (int c0 = 0; c0 < N; c0 += 16) {
    uint32_t r[16];
    const uint32_t ta = tmem + ((uint32_t)(32 * warp) << 16) + c0;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(ta) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int c = 0; c < 16; ++c) D[32 * warp + lane][c0 + c] = __uint_as_float(r[c]);
  }
  __syncthreads();
  if (tid == 0) {
    int first = -1, count = 0;
    if (!done) first = -99;
    else if (mode == 0) { for (int m = 0; m < 128; ++m) if (D[m][0] != 0.f) { if (first < 0) first = m; ++count; } }
    else if (mode == 1) { for (int n = 0; n < N; ++n) if (D[0][n] != 0.f) { if (first < 0) first = n; ++count; } }
    out_first[blk] = first;
    out_count[blk] = count;
    out_val[blk] = (mode == 2) ? D[m0][n0] : D[0][0];
  }
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(32u) : "memory");
}

static void run_cfg(const char* title, Cfg cfg) {
  printf("==== %s : A(lbo=%u sbo=%u layout=%u major=%u)  B(lbo=%u sbo=%u layout=%u major=%u)\n", title, cfg.a_lbo, cfg.a_sbo,
         cfg.a_layout, cfg.a_major, cfg.b_lbo, cfg.b_sbo, cfg.b_layout, cfg.b_major);
  int *f, *c; float* v;
  cudaMallocManaged(&f, 4096 * 4); cudaMallocManaged(&c, 4096 * 4); cudaMallocManaged(&v, 4096 * 4);
  const size_t smem = 4096 + 1024 + 64 + 1024;
  // A probe
  probe<<<A_WORDS, 128, smem>>>(cfg, 0, nullptr, nullptr, 1, 0, 0, f, c, v);
  if (cudaDeviceSynchronize() != cudaSuccess) { printf("A probe failed: %s\n", cudaGetErrorString(cudaGetLastError())); return; }
  printf("A word -> row m (count)   [word index = float offset in the 4096-byte tile]\n");
  for (int i = 0; i < A_WORDS; ++i) { if (i % 16 == 0) printf("\n%4d:", i); printf(" %3d/%d", f[i], c[i]); }
  printf("\n");
  static int a_row0[64]; int na = 0;
  for (int i = 0; i < A_WORDS && na < 64; ++i) if (f[i] == 0 && c[i] >= 1) a_row0[na++] = i;
  // B probe
  probe<<<B_WORDS, 128, smem>>>(cfg, 1, nullptr, nullptr, 1, 0, 0, f, c, v);
  cudaDeviceSynchronize();
  printf("B word -> column n (count)\n");
  for (int i = 0; i < B_WORDS; ++i) { if (i % 16 == 0) printf("\n%4d:", i); printf(" %3d/%d", f[i], c[i]); }
  printf("\n");
  static int b_col0[64]; int nbb = 0;
  for (int i = 0; i < B_WORDS && nbb < 64; ++i) if (f[i] == 0 && c[i] >= 1) b_col0[nbb++] = i;
  printf("A words of row 0:"); for (int i = 0; i < na; ++i) printf(" %d", a_row0[i]); printf("\n");
  printf("B words of col 0:"); for (int i = 0; i < nbb; ++i) printf(" %d", b_col0[i]); printf("\n");
  if (na > 0 && nbb > 0 && na <= 16 && nbb <= 16) {
    int *al, *bl; cudaMallocManaged(&al, 64 * 4); cudaMallocManaged(&bl, 64 * 4);
    for (int i = 0; i < na; ++i) al[i] = a_row0[i];
    for (int i = 0; i < nbb; ++i) bl[i] = b_col0[i];
    probe<<<na * nbb, 128, smem>>>(cfg, 2, al, bl, nbb, 0, 0, f, c, v);
    cudaDeviceSynchronize();
    printf("K pairing (rows: A words of row 0, cols: B words of col 0) D[0][0]:\n");
    for (int i = 0; i < na; ++i) { printf("  A%4d:", a_row0[i]); for (int j = 0; j < nbb; ++j) printf(" %g", v[i * nbb + j]); printf("\n"); }
  }
}

int main() {
  run_cfg("cfg1: A MN-major SW128, B K-major none", Cfg{1024, 4096, 2, 1, 128, 256, 0, 0});
  run_cfg("cfg2: A MN-major no-swizzle (lbo=128? sbo=...)", Cfg{128, 1024, 0, 1, 128, 256, 0, 0});
  run_cfg("cfg3: A K-major none, B K-major none", Cfg{128, 256, 0, 0, 128, 256, 0, 0});
  run_cfg("cfg4: A MN-major SW128 swapped lbo/sbo", Cfg{4096, 1024, 2, 1, 256, 128, 0, 0});
  return 0;
}
