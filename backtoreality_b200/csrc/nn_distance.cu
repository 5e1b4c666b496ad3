// nn_distance.cu -- nearest-neighbour (chamfer) matching of the detection losses (sm_100a).
//
// Replaces the index part of nn_distance (reference utils/nn_distance.py:34-61), which tiles
// pc1 to (B,N,M,C) and pc2 to (B,N,M,C) with .repeat(), subtracts, reduces to a (B,N,M) cost
// matrix and takes torch.min along both axes: four full (B,N,M,C) materialisations for two index
// vectors.  Here each (b, n) / (b, m) row scans the other cloud once from shared memory and keeps
// its arg-min; nothing of size N*M ever touches HBM.  The distances themselves are recomputed by
// the host mirror from the matched pairs with torch ops (so autograd flows exactly where
// torch.min's backward would send it: to the arg-min pair).
//
// cost modes (nn_distance.py:52-57): 0 = sum d^2, 1 = sum |d| (l1), 2 = sum huber(d, delta)
// (l1smooth).  Terms are accumulated channel by channel in ascending order with separate
// multiply / add (no fma), the order a sequential fp32 reduction over the last dim uses.  Ties:
// the lowest index wins.
#include "common.cuh"

namespace b2r {
namespace {

constexpr int kNdThreads = 128;
constexpr int kNdTile = 512;   // points of the scanned cloud staged per tile (C <= 4 channels)
constexpr int kNdMaxC = 4;

__device__ __forceinline__ float nd_term(float d, int mode, float delta) {
  if (mode == 0) return __fmul_rn(d, d);
  const float a = fabsf(d);
  if (mode == 1) return a;
  const float q = fminf(a, delta);           // huber_loss (nn_distance.py:16-32)
  const float lin = __fsub_rn(a, q);
  return __fadd_rn(__fmul_rn(__fmul_rn(0.5f, q), q), __fmul_rn(delta, lin));
}

// for every row of `a` (B,Na,C): arg-min over the rows of `b` (B,Nb,C) of the cost
__global__ void __launch_bounds__(kNdThreads)
    nn_argmin_kernel(const float *__restrict__ a, const float *__restrict__ b, int Na, int Nb,
                     int C, int mode, float delta, int flip, long long *__restrict__ idx) {
  __shared__ float s_b[kNdTile * kNdMaxC];
  const int bi = blockIdx.y;
  const int i = blockIdx.x * kNdThreads + threadIdx.x;
  a += (size_t)bi * Na * C;
  b += (size_t)bi * Nb * C;
  float p[kNdMaxC] = {0.f, 0.f, 0.f, 0.f};
  if (i < Na)
    for (int c = 0; c < C; ++c) p[c] = a[(size_t)i * C + c];
  float best = INFINITY;
  int bj = 0;
  for (int base = 0; base < Nb; base += kNdTile) {
    const int n = min(kNdTile, Nb - base);
    __syncthreads();
    for (int e = threadIdx.x; e < n * C; e += kNdThreads) s_b[e] = b[(size_t)base * C + e];
    __syncthreads();
    if (i < Na)
      for (int j = 0; j < n; ++j) {
        float cost = 0.f;
        for (int c = 0; c < C; ++c) {
          // the reference subtracts pc1 - pc2; with flip the roles of a and b are swapped
          const float d = flip ? __fsub_rn(s_b[j * C + c], p[c]) : __fsub_rn(p[c], s_b[j * C + c]);
          cost = __fadd_rn(cost, nd_term(d, mode, delta));
        }
        if (cost < best) {   // strict: the lowest index wins ties
          best = cost;
          bj = base + j;
        }
      }
  }
  if (i < Na) idx[(size_t)bi * Na + i] = bj;
}

}  // namespace
}  // namespace b2r

extern "C" int b2r_nn_argmin(const float *pc1, const float *pc2, int B, int N, int M, int C,
                             int mode, float delta, long long *idx1, long long *idx2,
                             void *stream) {
  B2R_REQUIRE(B >= 0 && N >= 0 && M >= 0, "b2r_nn_argmin: negative size");
  B2R_REQUIRE(C >= 1 && C <= b2r::kNdMaxC, "b2r_nn_argmin: supports 1..4 channels (got %d)", C);
  B2R_REQUIRE(mode >= 0 && mode <= 2, "b2r_nn_argmin: mode must be 0 (l2), 1 (l1) or 2 (l1smooth)");
  if (B == 0) return B2R_OK;
  B2R_REQUIRE(N > 0 && M > 0, "b2r_nn_argmin: empty point set (torch.min over an empty dim raises)");
  B2R_REQUIRE(pc1 && pc2 && (idx1 || idx2), "b2r_nn_argmin: null pointer");
  B2R_REQUIRE(B <= 65535, "b2r_nn_argmin: B=%d exceeds gridDim.y", B);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (idx1) {
    dim3 g(b2r::ceil_div(N, b2r::kNdThreads), B, 1);
    b2r::nn_argmin_kernel<<<g, b2r::kNdThreads, 0, st>>>(pc1, pc2, N, M, C, mode, delta, 0, idx1);
  }
  if (idx2) {
    dim3 g(b2r::ceil_div(M, b2r::kNdThreads), B, 1);
    b2r::nn_argmin_kernel<<<g, b2r::kNdThreads, 0, st>>>(pc2, pc1, M, N, C, mode, delta, 1, idx2);
  }
  B2R_CHECK_LAUNCH();
  return B2R_OK;
}
