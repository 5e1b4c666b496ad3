// lat_probe.cu -- dependent-chain latencies of the primitives FPS's per-iteration reduction is
// built from (sm_100a): redux.sync, shfl.sync, vote.ballot, LDS round trip, bar.sync with 16 warps,
// st.async -> mbarrier round trip between two CTAs of a cluster.  Prints cycles per dependent op.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#define REP 256

__global__ void k_redux(uint32_t *out, long long *cyc) {
  uint32_t v = threadIdx.x * 2654435761u;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < REP; ++i) v = __reduce_max_sync(0xffffffffu, v ^ (uint32_t)i) + threadIdx.x;
  long long t1 = clock64();
  out[threadIdx.x] = v;
  if (threadIdx.x == 0) cyc[0] = (t1 - t0) / REP;
}
__global__ void k_shfl(uint32_t *out, long long *cyc) {
  uint32_t v = threadIdx.x * 2654435761u;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < REP; ++i) v = __shfl_xor_sync(0xffffffffu, v, 1 + (i & 15)) + 1;
  long long t1 = clock64();
  out[threadIdx.x] = v;
  if (threadIdx.x == 0) cyc[0] = (t1 - t0) / REP;
}
__global__ void k_ballot(uint32_t *out, long long *cyc) {
  uint32_t v = threadIdx.x * 2654435761u;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < REP; ++i) v = __ballot_sync(0xffffffffu, (v >> (i & 7)) & 1) + threadIdx.x;
  long long t1 = clock64();
  out[threadIdx.x] = v;
  if (threadIdx.x == 0) cyc[0] = (t1 - t0) / REP;
}
__global__ void k_lds(uint32_t *out, long long *cyc) {
  __shared__ uint32_t s[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) s[i] = (i * 7 + 3) & 1023;
  __syncthreads();
  uint32_t v = threadIdx.x;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < REP; ++i) v = s[v];
  long long t1 = clock64();
  out[threadIdx.x] = v;
  if (threadIdx.x == 0) cyc[0] = (t1 - t0) / REP;
}
// STS by one lane -> bar.sync (all warps) -> LDS by everyone, repeated: the CTA-level round
__global__ void k_bar(uint32_t *out, long long *cyc) {
  __shared__ uint32_t s[2][32];
  uint32_t v = threadIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < REP; ++i) {
    if (lane == 0) s[i & 1][warp] = v + i;
    __syncthreads();
    v = s[i & 1][(lane + i) & ((blockDim.x >> 5) - 1)];
  }
  long long t1 = clock64();
  out[threadIdx.x] = v;
  if (threadIdx.x == 0) cyc[0] = (t1 - t0) / REP;
}
// cluster of 2: ping-pong with st.async + mbarrier complete_tx
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void __cluster_dims__(2, 1, 1) k_dsmem(uint32_t *out, long long *cyc) {
  __shared__ __align__(8) uint64_t bar[2];
  __shared__ uint32_t slot[2];
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const uint32_t b0 = smem_u32(&bar[0]);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b0));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b0 + 8));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], 4;" ::"r"(b0));
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], 4;" ::"r"(b0 + 8));
  }
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  uint32_t rslot[2], rbar[2];
  for (int p = 0; p < 2; ++p) {
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rslot[p]) : "r"(smem_u32(&slot[p])), "r"(rank ^ 1));
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rbar[p]) : "r"(b0 + 8 * p), "r"(rank ^ 1));
  }
  uint32_t v = 1;
  long long t0 = clock64();
  if (threadIdx.x == 0) {
#pragma unroll 1
    for (int i = 0; i < REP; ++i) {
      const int p = i & 1;
      // both CTAs send, both wait: one exchange per iteration (what FPS does)
      asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(rslot[p]), "r"(v), "r"(rbar[p]) : "memory");
      asm volatile(
          "{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" ::"r"(b0 + 8 * p), "r"((uint32_t)(i >> 1) & 1u) : "memory");
      v = slot[p] + 1;
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], 4;" ::"r"(b0 + 8 * p) : "memory");
    }
  }
  long long t1 = clock64();
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (threadIdx.x == 0 && rank == 0) { cyc[0] = (t1 - t0) / REP; out[0] = v; }
}

int main() {
  uint32_t *out; long long *cyc, h;
  cudaMalloc(&out, 4096 * 4); cudaMalloc(&cyc, 8);
#define RUN(name, kern, threads)                                              \
  kern<<<1, threads>>>(out, cyc); cudaDeviceSynchronize();                   \
  kern<<<1, threads>>>(out, cyc); cudaDeviceSynchronize();                   \
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);                            \
  printf("%-34s %4lld cycles per dependent op  (%s)\n", name, h, cudaGetErrorString(cudaGetLastError()));
  RUN("redux.sync.max.u32 (1 warp)", k_redux, 32)
  RUN("redux.sync.max.u32 (16 warps)", k_redux, 512)
  RUN("shfl.sync.bfly (1 warp)", k_shfl, 32)
  RUN("shfl.sync.bfly (16 warps)", k_shfl, 512)
  RUN("vote.ballot (1 warp)", k_ballot, 32)
  RUN("LDS dependent (1 warp)", k_lds, 32)
  RUN("STS+bar.sync+LDS round, 4 warps", k_bar, 128)
  RUN("STS+bar.sync+LDS round, 16 warps", k_bar, 512)
  RUN("STS+bar.sync+LDS round, 32 warps", k_bar, 1024)
  k_dsmem<<<2, 32>>>(out, cyc); cudaDeviceSynchronize();
  k_dsmem<<<2, 32>>>(out, cyc); cudaDeviceSynchronize();
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-34s %4lld cycles per exchange  (%s)\n", "st.async+mbarrier exchange, 2 CTAs", h, cudaGetErrorString(cudaGetLastError()));
  return 0;
}
