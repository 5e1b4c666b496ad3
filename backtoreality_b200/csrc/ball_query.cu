// ball_query.cu -- radius neighbour search with the reference's order semantics (sm_100a).
//
// Replaces query_ball_point_kernel (reference _ext_src/src/ball_query_gpu.cu:14-49): one THREAD
// per centre scanning all N points serially with strided 12-byte loads, B CTAs in total.
//
// Here: one WARP per centre, kWarps centres per CTA.  The CTA stages tiles of points in shared
// memory as structure-of-arrays; each lane tests 4 consecutive points per step with three
// 128-bit shared loads (float4-vectorised), so a warp covers 128 candidates per step, in index
// order.  Hits are rare (tens out of N), so the common path is pure FADD/FFMA work and ONE
// warp vote per 128 candidates; only on a hit are the four ballots ranked to find the output
// slots.  A warp stops as soon as it has `nsample` hits, a CTA as soon as all its warps have.
//
// Bit-exactness: d2 = fma(dz,dz, fma(dx,dx, dy*dy)) with
// d* = centre - point, compared with `<` against radius*radius computed in fp32.
// Output: first `nsample` hits in ascending index order, remaining slots = first hit,
// empty ball = zeros (the reference's torch::zeros init, ball_query.cpp:24-26).
#include "common.cuh"

namespace b2r {
namespace {

constexpr int kBqWarps = 8;       // centres per CTA
constexpr int kBqTile = 1024;     // points per shared-memory tile
constexpr int kBqPad = 12;        // array skew (keeps 16-B alignment, spreads banks on fill)
constexpr int kBqStride = kBqTile + kBqPad;

__global__ void __launch_bounds__(kBqWarps * 32)
    ball_query_kernel(const float *__restrict__ new_xyz, const float *__restrict__ xyz, int N,
                      int M, float radius, int nsample, int *__restrict__ idx) {
  __shared__ __align__(16) float s_p[3 * kBqStride];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.y;
  const int j = blockIdx.x * kBqWarps + warp;  // this warp's centre
  const bool active = j < M;
  xyz += (size_t)b * N * 3;

  float cx = 0.f, cy = 0.f, cz = 0.f;
  if (active) {
    const float *q = new_xyz + ((size_t)b * M + j) * 3;
    cx = q[0]; cy = q[1]; cz = q[2];
  }
  int *out = idx + ((size_t)b * M + (active ? j : 0)) * nsample;
  const float r2 = __fmul_rn(radius, radius);

  int cnt = 0;      // hits so far (warp-uniform)
  int first = 0;    // index of the first hit (warp-uniform)
  bool done = !active || nsample <= 0;

  for (int base = 0; base < N; base += kBqTile) {
    const int tile_n = min(kBqTile, N - base);
    __syncthreads();  // previous tile fully consumed
    // coalesced fill: AoS (x,y,z) in global -> SoA in shared
    for (int e = tid; e < tile_n * 3; e += kBqWarps * 32) {
      const int p = e / 3, c = e - p * 3;
      s_p[c * kBqStride + p] = xyz[(size_t)base * 3 + e];
    }
    // pad the tail of the last tile to a multiple of 4 with never-matching points
    if (tile_n & 3) {
      const int padded = (tile_n + 3) & ~3;
      if (tid < (padded - tile_n) * 3) {
        const int p = tile_n + tid / 3, c = tid % 3;
        s_p[c * kBqStride + p] = __int_as_float(0x7fc00000);  // NaN: d2 < r2 is false
      }
    }
    if (__syncthreads_and(done ? 1 : 0)) break;  // tile visible; all centres satisfied?

    if (!done) {
      for (int step = 0; step < tile_n; step += 128) {  // warp-uniform trip count
        const int off = step + lane * 4;
        const bool in = off < tile_n;  // lanes past the tile end test nothing
        const float4 X = *reinterpret_cast<const float4 *>(&s_p[0 * kBqStride + off]);
        const float4 Y = *reinterpret_cast<const float4 *>(&s_p[1 * kBqStride + off]);
        const float4 Z = *reinterpret_cast<const float4 *>(&s_p[2 * kBqStride + off]);
        const bool h0 =
            in && sumsq_ref(__fsub_rn(cx, X.x), __fsub_rn(cy, Y.x), __fsub_rn(cz, Z.x)) < r2;
        const bool h1 =
            in && sumsq_ref(__fsub_rn(cx, X.y), __fsub_rn(cy, Y.y), __fsub_rn(cz, Z.y)) < r2;
        const bool h2 =
            in && sumsq_ref(__fsub_rn(cx, X.z), __fsub_rn(cy, Y.z), __fsub_rn(cz, Z.z)) < r2;
        const bool h3 =
            in && sumsq_ref(__fsub_rn(cx, X.w), __fsub_rn(cy, Y.w), __fsub_rn(cz, Z.w)) < r2;
        if (__any_sync(0xffffffffu, h0 | h1 | h2 | h3)) {
          const unsigned m0 = __ballot_sync(0xffffffffu, h0), m1 = __ballot_sync(0xffffffffu, h1);
          const unsigned m2 = __ballot_sync(0xffffffffu, h2), m3 = __ballot_sync(0xffffffffu, h3);
          const unsigned lt = (1u << lane) - 1u;
          // hits of lower lanes come first (their points have lower indices), then own j'<j
          int pos = cnt + __popc(m0 & lt) + __popc(m1 & lt) + __popc(m2 & lt) + __popc(m3 & lt);
          const int k0 = base + off;
          if (cnt == 0) {  // warp-uniform: remember the very first hit
            const unsigned any = m0 | m1 | m2 | m3;
            const int fl = __ffs(any) - 1;  // lowest lane with a hit
            const int fj = ((m0 >> fl) & 1) ? 0 : ((m1 >> fl) & 1) ? 1 : ((m2 >> fl) & 1) ? 2 : 3;
            first = base + step + fl * 4 + fj;
          }
          if (h0) { if (pos < nsample) out[pos] = k0 + 0; ++pos; }
          if (h1) { if (pos < nsample) out[pos] = k0 + 1; ++pos; }
          if (h2) { if (pos < nsample) out[pos] = k0 + 2; ++pos; }
          if (h3) { if (pos < nsample) out[pos] = k0 + 3; ++pos; }
          cnt += __popc(m0) + __popc(m1) + __popc(m2) + __popc(m3);
          if (cnt >= nsample) { done = true; break; }
        }
      }
    }
  }

  // slots past the hit count repeat the first hit; an empty ball is all zeros
  if (active) {
    const int filled = min(cnt, nsample);
    const int fill = cnt > 0 ? first : 0;
    for (int l = filled + lane; l < nsample; l += 32) out[l] = fill;
  }
}

}  // namespace
}  // namespace b2r

extern "C" int b2r_ball_query(const float *new_xyz, const float *xyz, int B, int N, int M,
                              float radius, int nsample, int *idx, void *stream) {
  B2R_REQUIRE(B >= 0 && N >= 0 && M >= 0 && nsample >= 0,
              "b2r_ball_query: negative size (B=%d N=%d M=%d nsample=%d)", B, N, M, nsample);
  if (B == 0 || M == 0 || nsample == 0) return B2R_OK;
  B2R_REQUIRE(new_xyz && idx && (xyz || N == 0), "b2r_ball_query: null pointer");
  B2R_REQUIRE(B <= 65535, "b2r_ball_query: B=%d exceeds gridDim.y", B);
  dim3 grid(b2r::ceil_div(M, b2r::kBqWarps), B, 1);
  b2r::ball_query_kernel<<<grid, b2r::kBqWarps * 32, 0, static_cast<cudaStream_t>(stream)>>>(
      new_xyz, xyz, N, M, radius, nsample, idx);
  B2R_CHECK_LAUNCH();
  return B2R_OK;
}
