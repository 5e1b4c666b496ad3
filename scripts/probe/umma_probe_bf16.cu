// umma_probe_bf16.cu -- standalone probe of tcgen05 kind::f16 (BF16) shared-memory operand layouts
// (K-major SW128 and the MN-major view of the SAME bytes).
// Validates, on the GPU, which (major, swizzle mode, byte layout, LBO/SBO) combinations the
// hardware accepts for TF32 operands, before the SharedMLP backward kernel relies on them.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_probe umma_probe.cu && ./umma_probe
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include <functional>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

struct Params {
  uint32_t a_bytes, b_bytes;       // image sizes (multiples of 16)
  uint32_t a_lbo, a_sbo, a_type, a_kstep;   // descriptor fields, byte advance per K=8 step
  uint32_t b_lbo, b_sbo, b_type, b_kstep;
  uint32_t idesc, ksteps, N;
};

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t type) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fffu);
  d |= (uint64_t)((lbo >> 4) & 0x3fffu) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3fffu) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)type << 61;
  return d;
}

__global__ void __launch_bounds__(128, 1) probe_kernel(const uint8_t *a_img, const uint8_t *b_img, Params p, float *out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *base = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t *sa = base, *sb = base + ((p.a_bytes + 1023) & ~1023u);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (uint32_t i = tid * 16; i < p.a_bytes; i += 128 * 16) *(uint4 *)(sa + i) = *(const uint4 *)(a_img + i);
  for (uint32_t i = tid * 16; i < p.b_bytes; i += 128 * 16) *(uint4 *)(sb + i) = *(const uint4 *)(b_img + i);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  if (tid == 0) {
    for (uint32_t ks = 0; ks < p.ksteps; ++ks) {
      uint64_t da = make_desc(smem_u32(sa) + ks * p.a_kstep, p.a_lbo, p.a_sbo, p.a_type);
      uint64_t db = make_desc(smem_u32(sb) + ks * p.b_kstep, p.b_lbo, p.b_sbo, p.b_type);
      uint32_t acc = ks > 0;
      asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %4, 0;\n\t"
                   "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, q;\n\t}" ::"r"(tmem), "l"(da), "l"(db), "r"(p.idesc), "r"(acc) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  asm volatile("{\n\t.reg .pred q;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%0], 0;\n\t@q bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" ::"r"(smem_u32(&bar)) : "memory");
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // D: lane = m (128), column = n
  for (uint32_t c0 = 0; c0 < p.N; c0 += 8) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(tmem + ((uint32_t)(warp * 32) << 16) + c0));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int i = 0; i < 8; ++i) out[(size_t)tid * p.N + c0 + i] = __uint_as_float(r[i]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u) : "memory");
}

// ---- host-side layout functions: byte offset of element (row r of the MN extent, column k) ----
#include <cuda_bf16.h>
typedef std::function<uint32_t(int, int)> Off;
static uint32_t idesc(int n, int a_mn, int b_mn) {   // kind::f16: BF16 x BF16 -> F32
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24);
}
struct Operand { const char *name; int mn_major; uint32_t type, lbo, sbo, kstep, bytes; Off off; };

// K-major SW128, bf16: atom = 8 rows x 64 elements (128 B); K atoms (64 el) rows/8*1024 apart
static Operand k_major_sw128(int rows, int K) {
  Operand o; o.name = "K-major SW128"; o.mn_major = 0; o.type = 2; o.lbo = 16; o.sbo = 1024; o.kstep = 32;
  o.bytes = rows * 128 * (K / 64);
  o.off = [=](int r, int k) { return (uint32_t)((k >> 6) * (rows >> 3) * 1024 + (r >> 3) * 1024 + (r & 7) * 128 + (((((k & 63) >> 3)) ^ (r & 7)) << 4) + (k & 7) * 2); };
  return o;
}
// MN-major view of a K-major SW128 tile of the TRANSPOSED matrix: tile rows = k (K extent), 64 mn contiguous
static Operand mn_major_dual(int rows, int K, bool swap) {
  Operand o; o.name = swap ? "MN-major SW128 dual-view, LBO/SBO swapped" : "MN-major SW128 dual-view"; o.mn_major = 1; o.type = 2;
  uint32_t lbo = (K / 8) * 1024, sbo = 1024;     // 64-wide mn atoms, 8-deep k groups
  o.lbo = swap ? sbo : lbo; o.sbo = swap ? lbo : sbo; o.kstep = 2048; o.bytes = (rows / 64) * lbo;
  o.off = [=](int r, int k) { return (uint32_t)((r >> 6) * lbo + (k >> 3) * 1024 + (k & 7) * 128 + (((((r & 63) >> 3)) ^ (k & 7)) << 4) + (r & 7) * 2); };
  return o;
}

static double run(const Operand &A, const Operand &B, int N, int K) {
  std::vector<float> a(128 * K), b(N * K);
  srand(1);
  for (auto &v : a) v = (float)((rand() % 17) - 8);
  for (auto &v : b) v = (float)((rand() % 13) - 6);
  std::vector<uint8_t> ai(A.bytes + 1024, 0), bi(B.bytes + 1024, 0);
  for (int r = 0; r < 128; ++r) for (int k = 0; k < K; ++k) { __nv_bfloat16 h = __float2bfloat16(a[r * K + k]); memcpy(&ai[A.off(r, k)], &h, 2); }
  for (int r = 0; r < N; ++r) for (int k = 0; k < K; ++k) { __nv_bfloat16 h = __float2bfloat16(b[r * K + k]); memcpy(&bi[B.off(r, k)], &h, 2); }
  uint8_t *da, *db; float *dout;
  uint32_t ab = (A.bytes + 15) & ~15u, bb = (B.bytes + 15) & ~15u;
  CK(cudaMalloc(&da, ab)); CK(cudaMalloc(&db, bb)); CK(cudaMalloc(&dout, 128 * N * 4));
  CK(cudaMemcpy(da, ai.data(), ab, cudaMemcpyHostToDevice)); CK(cudaMemcpy(db, bi.data(), bb, cudaMemcpyHostToDevice));
  CK(cudaMemset(dout, 0xff, 128 * N * 4));
  Params p; p.a_bytes = ab; p.b_bytes = bb;
  p.a_lbo = A.lbo; p.a_sbo = A.sbo; p.a_type = A.type; p.a_kstep = A.kstep;
  p.b_lbo = B.lbo; p.b_sbo = B.sbo; p.b_type = B.type; p.b_kstep = B.kstep;
  p.idesc = idesc(N, A.mn_major, B.mn_major); p.ksteps = K / 16; p.N = N;
  size_t smem = ((ab + 1023) & ~1023u) + bb + 2048;
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  probe_kernel<<<1, 128, smem>>>(da, db, p, dout);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("   kernel error: %s\n", cudaGetErrorString(e)); exit(2); }
  std::vector<float> out(128 * N);
  CK(cudaMemcpy(out.data(), dout, 128 * N * 4, cudaMemcpyDeviceToHost));
  double maxerr = 0; int bad = 0;
  for (int m = 0; m < 128; ++m) for (int n = 0; n < N; ++n) {
    double w = 0; for (int k = 0; k < K; ++k) w += (double)a[m * K + k] * b[n * K + k];
    double d = fabs(w - out[m * N + n]); if (!(d <= 1e-3)) bad++; if (d > maxerr || d != d) maxerr = d;
  }
  cudaFree(da); cudaFree(db); cudaFree(dout);
  printf("A: %-44s B: %-44s  N=%d K=%d mismatches %d / %d  maxerr %g\n", A.name, B.name, N, K, bad, 128 * N, maxerr);
  return maxerr;
}

int main(int argc, char **argv) {
  const int K = 64, N = 64;
  const int v = argc > 1 ? atoi(argv[1]) : 0;
  const int swap = v >= 4 ? 1 : 0;
  switch (v % 4) {
    case 0: run(k_major_sw128(128, K), k_major_sw128(N, K), N, K); break;
    case 1: run(mn_major_dual(128, K, swap), k_major_sw128(N, K), N, K); break;
    case 2: run(k_major_sw128(128, K), mn_major_dual(N, K, swap), N, K); break;
    case 3: run(mn_major_dual(128, K, swap), mn_major_dual(N, K, swap), N, K); break;
  }
  return 0;
}
