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

// =================================================================================================
// Grid-accelerated ball query (same results, bit for bit).
//
// The brute-force kernel above tests every (centre, point) pair: 82 M tests per 40k-point scene at
// SA1, of which ~25 per centre hit.  Here the points of a scene are first bucketed into a hashed
// uniform grid of cell size radius*(1+1e-3); a centre then only visits the 27 cells around it.
// The reference's semantics "the first `nsample` hits in ascending index order, remaining slots =
// the first hit" (ball_query_gpu.cu:14-49) equal "the nsample SMALLEST hit indices, ascending,
// padded with the smallest": the warp collects the hits of the 27 cells unordered, sorts them
// (bitonic, in shared memory) and writes the head.  The distance test is the same instruction
// sequence on the same operands, so the hit SET is identical; every point with d < radius lies in
// one of the 27 cells because |dx| < radius < cell size (the 1e-3 margin dominates the rounding of
// x * (1/cell) for |x| < 8000 cells; scenes with larger coordinates are detected by the bucketing
// pass and scanned in index order instead, see kSafeCells).  Hash aliasing (64 x 64 x 32 cells, coordinates wrapped)
// only adds candidates that the distance test rejects.
namespace b2r {
namespace {

constexpr int kGx = 64, kGy = 64, kGz = 32;          // hashed grid (wrap-around)
constexpr int kGCells = kGx * kGy * kGz;             // 131072
constexpr int kHitCap = 512;                         // per-warp candidate buffer
constexpr int kGqWarps = 8;

// Cell coordinate of x: floor(x / cell), clamped to +-2^30 before the int cast (a float outside
// the int range, or NaN, is undefined behaviour in the cast; such points can never pass the
// distance test, any cell will do for them).
__device__ __forceinline__ int cell_coord(float x, float inv) {
  return (int)fminf(fmaxf(floorf(x * inv), -1073741824.f), 1073741824.f);
}
__device__ __forceinline__ int grid_cell(float x, float y, float z, float inv) {
  const int ix = cell_coord(x, inv) & (kGx - 1);
  const int iy = cell_coord(y, inv) & (kGy - 1);
  const int iz = cell_coord(z, inv) & (kGz - 1);
  return (iz * kGy + iy) * kGx + ix;
}
// The 27-cell argument needs the fp32 rounding of x * (1/cell) to stay below the 1e-3 cell margin:
// |x / cell| * 2^-24 < 1e-3 / 2, i.e. |x| < ~8000 cells.  A scene with a finite coordinate beyond
// kSafeCells sets its flag and every centre of that scene takes the index-order scan instead.
constexpr float kSafeCells = 7900.f;
__device__ __forceinline__ bool beyond_safe(float x, float y, float z, float inv) {
  const float m = fmaxf(fabsf(x * inv), fmaxf(fabsf(y * inv), fabsf(z * inv)));
  return m > kSafeCells && m < INFINITY;   // NaN / inf points never hit: no need to fall back
}

// counts per cell; cell id of every point kept for the fill pass
__global__ void grid_count_kernel(const float *__restrict__ xyz, int N, float inv,
                                  int *__restrict__ count, int *__restrict__ cell_of,
                                  int *__restrict__ far_flag) {
  const int b = blockIdx.y;
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= N) return;
  const float *p = xyz + ((size_t)b * N + k) * 3;
  if (beyond_safe(p[0], p[1], p[2], inv)) far_flag[b] = 1;
  const int c = grid_cell(p[0], p[1], p[2], inv);
  cell_of[(size_t)b * N + k] = c;
  atomicAdd(count + (size_t)b * (kGCells + 1) + c, 1);
}

// Offsets of the cells.  Any assignment of disjoint ranges of [0, N) to the cells is valid (the
// query sorts its hits), so each CTA scans a chunk of 8192 cells locally and claims its base
// range with ONE atomicAdd on the scene's running total (kept in count[kGCells], zeroed with the
// counts): no second pass, kScanChunks CTAs per scene instead of one.  count[] is replaced by
// start[]; the cursor array is zeroed for the fill pass and ends up holding the cell
// populations, so a cell's range is [start, start + cursor).
constexpr int kScanChunks = kGCells / 8192;   // 16
__global__ void __launch_bounds__(1024) grid_scan_kernel(int *__restrict__ count_start,
                                                         int *__restrict__ cursor) {
  __shared__ int s_part[1024];
  __shared__ int s_base;
  const int b = blockIdx.y, t = threadIdx.x;
  int *cs = count_start + (size_t)b * (kGCells + 1) + (size_t)blockIdx.x * 8192;
  int *cur = cursor + (size_t)b * kGCells + (size_t)blockIdx.x * 8192;
  int c[8], sum = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {   // thread t takes cells t, t+1024, ...: coalesced
    c[i] = cs[i * 1024 + t];
    sum += c[i];
  }
  s_part[t] = sum;
  __syncthreads();
  for (int off = 1; off < 1024; off <<= 1) {   // Hillis-Steele inclusive scan of the partials
    const int v = t >= off ? s_part[t - off] : 0;
    __syncthreads();
    s_part[t] += v;
    __syncthreads();
  }
  if (t == 1023) s_base = atomicAdd(count_start + (size_t)b * (kGCells + 1) + kGCells, s_part[t]);
  __syncthreads();
  int run = s_base + s_part[t] - sum;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    cs[i * 1024 + t] = run;
    cur[i * 1024 + t] = 0;
    run += c[i];
  }
}

__global__ void grid_fill_kernel(const int *__restrict__ cell_of, const int *__restrict__ start,
                                 int *__restrict__ cursor, int N, int *__restrict__ sorted) {
  const int b = blockIdx.y;
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= N) return;
  const int c = cell_of[(size_t)b * N + k];
  const int pos = atomicAdd(cursor + (size_t)b * kGCells + c, 1);
  sorted[(size_t)b * N + start[(size_t)b * (kGCells + 1) + c] + pos] = k;
}

// ascending bitonic sort of n (power of two, <= kHitCap) ints in shared memory by one warp
__device__ __forceinline__ void warp_bitonic_sort(int *v, int n, int lane) {
  for (int k = 2; k <= n; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = lane; i < n; i += 32) {
        const int l = i ^ j;
        if (l > i) {
          const int a = v[i], c = v[l];
          const bool up = (i & k) == 0;
          if ((a > c) == up) { v[i] = c; v[l] = a; }
        }
      }
      __syncwarp();
    }
}

__global__ void __launch_bounds__(kGqWarps * 32)
    grid_query_kernel(const float *__restrict__ new_xyz, const float *__restrict__ xyz,
                      const int *__restrict__ start, const int *__restrict__ pop,
                      const int *__restrict__ sorted, int N, int M, float radius, float inv,
                      int nsample, int *__restrict__ idx, const int *__restrict__ far_flag) {
  __shared__ int s_hits[kGqWarps][kHitCap];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.y;
  const int j = blockIdx.x * kGqWarps + warp;
  if (j >= M) return;   // whole warp
  xyz += (size_t)b * N * 3;
  start += (size_t)b * (kGCells + 1);
  pop += (size_t)b * kGCells;
  sorted += (size_t)b * N;
  int *hits = s_hits[warp];
  const float *q = new_xyz + ((size_t)b * M + j) * 3;
  const float cx = q[0], cy = q[1], cz = q[2];
  const float r2 = __fmul_rn(radius, radius);
  int *out = idx + ((size_t)b * M + j) * nsample;
  if (far_flag[b] != 0 || beyond_safe(cx, cy, cz, inv)) {
    // coordinates too large for the grid's rounding margin: the reference's scan in index order
    // (ball_query_gpu.cu:31-47), 32 points per step
    int cnt = 0, first = 0;
    for (int k0 = 0; k0 < N && cnt < nsample; k0 += 32) {
      const int k = k0 + lane;
      bool hit = false;
      if (k < N) {
        const float *p = xyz + (size_t)k * 3;
        hit = sumsq_ref(__fsub_rn(cx, p[0]), __fsub_rn(cy, p[1]), __fsub_rn(cz, p[2])) < r2;
      }
      const unsigned m = __ballot_sync(0xffffffffu, hit);
      if (m) {
        if (cnt == 0) first = k0 + __ffs(m) - 1;
        const int pos = cnt + __popc(m & ((1u << lane) - 1u));
        if (hit && pos < nsample) out[pos] = k;
        cnt += __popc(m);
      }
    }
    for (int l = min(cnt, nsample) + lane; l < nsample; l += 32) out[l] = cnt > 0 ? first : 0;
    return;
  }
  const int ix = cell_coord(cx, inv), iy = cell_coord(cy, inv), iz = cell_coord(cz, inv);
  int n = 0;   // hits buffered (warp-uniform)

  // sort what is buffered and keep the `keep` smallest: they are the only ones that can matter
  auto compact = [&](int keep) {
    int np2 = 32;
    while (np2 < n) np2 <<= 1;
    for (int i = n + lane; i < np2; i += 32) hits[i] = 0x7fffffff;
    __syncwarp();
    warp_bitonic_sort(hits, np2, lane);
    if (n > keep) n = keep;
  };

  for (int dz = -1; dz <= 1; ++dz)
    for (int dy = -1; dy <= 1; ++dy)
      for (int dx = -1; dx <= 1; ++dx) {
        const int c = ((((iz + dz) & (kGz - 1)) * kGy + ((iy + dy) & (kGy - 1))) * kGx) +
                      ((ix + dx) & (kGx - 1));
        const int s0 = start[c], s1 = s0 + pop[c];
        for (int e0 = s0; e0 < s1; e0 += 32) {   // warp-uniform trip count
          const int e = e0 + lane;
          bool hit = false;
          int k = 0;
          if (e < s1) {
            k = sorted[e];
            const float *p = xyz + (size_t)k * 3;
            hit = sumsq_ref(__fsub_rn(cx, p[0]), __fsub_rn(cy, p[1]), __fsub_rn(cz, p[2])) < r2;
          }
          const unsigned m = __ballot_sync(0xffffffffu, hit);
          if (m) {
            if (n + 32 > kHitCap) {   // make room: only the nsample smallest can be emitted
              compact(nsample < kHitCap - 32 ? nsample : kHitCap - 32);
              __syncwarp();
            }
            if (hit) hits[n + __popc(m & ((1u << lane) - 1u))] = k;
            n += __popc(m);
            __syncwarp();
          }
        }
      }
  if (n == 0) {
    for (int l = lane; l < nsample; l += 32) out[l] = 0;
    return;
  }
  compact(nsample);
  __syncwarp();
  const int first = hits[0];
  for (int l = lane; l < nsample; l += 32) out[l] = l < n ? hits[l] : first;
}

inline size_t grid_ws_bytes(int B, int N) {
  // per scene: count/start (kGCells + 1), cursor (kGCells), cell_of (N), sorted (N) ints and
  // one "coordinates beyond the grid's safe range" flag
  return (size_t)B * ((size_t)(kGCells + 1) + kGCells + 2 * (size_t)N + 1) * sizeof(int);
}

}  // namespace
}  // namespace b2r

extern "C" long long b2r_ball_query_workspace_bytes(int B, int N) {
  if (B <= 0 || N <= 0) return 0;
  return (long long)b2r::grid_ws_bytes(B, N);
}

extern "C" int b2r_ball_query_grid(const float *new_xyz, const float *xyz, int B, int N, int M,
                                   float radius, int nsample, int *idx, void *workspace,
                                   long long workspace_bytes, void *stream) {
  B2R_REQUIRE(B >= 0 && N >= 0 && M >= 0 && nsample >= 0,
              "b2r_ball_query_grid: negative size (B=%d N=%d M=%d nsample=%d)", B, N, M, nsample);
  if (B == 0 || M == 0 || nsample == 0) return B2R_OK;
  // degenerate radii / tiny scenes / oversized nsample: the brute-force kernel is the right tool
  if (N < 1024 || !(radius > 0.f) || nsample > b2r::kHitCap - 32)
    return b2r_ball_query(new_xyz, xyz, B, N, M, radius, nsample, idx, stream);
  B2R_REQUIRE(new_xyz && xyz && idx && workspace, "b2r_ball_query_grid: null pointer");
  B2R_REQUIRE(workspace_bytes >= (long long)b2r::grid_ws_bytes(B, N),
              "b2r_ball_query_grid: workspace too small (%lld < %lld bytes)", workspace_bytes,
              (long long)b2r::grid_ws_bytes(B, N));
  B2R_REQUIRE(B <= 65535, "b2r_ball_query_grid: B=%d exceeds gridDim.y", B);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int *count = static_cast<int *>(workspace);
  int *cursor = count + (size_t)B * (b2r::kGCells + 1);
  int *cell_of = cursor + (size_t)B * b2r::kGCells;
  int *sorted = cell_of + (size_t)B * N;
  int *far_flag = sorted + (size_t)B * N;
  const float inv = 1.0f / (radius * 1.001f);
  B2R_CUDA(cudaMemsetAsync(count, 0, sizeof(int) * (size_t)B * (b2r::kGCells + 1), st));
  B2R_CUDA(cudaMemsetAsync(far_flag, 0, sizeof(int) * (size_t)B, st));
  dim3 gp(b2r::ceil_div(N, 256), B, 1);
  b2r::grid_count_kernel<<<gp, 256, 0, st>>>(xyz, N, inv, count, cell_of, far_flag);
  b2r::grid_scan_kernel<<<dim3(b2r::kScanChunks, B, 1), 1024, 0, st>>>(count, cursor);
  b2r::grid_fill_kernel<<<gp, 256, 0, st>>>(cell_of, count, cursor, N, sorted);
  dim3 gq(b2r::ceil_div(M, b2r::kGqWarps), B, 1);
  b2r::grid_query_kernel<<<gq, b2r::kGqWarps * 32, 0, st>>>(new_xyz, xyz, count, cursor, sorted, N,
                                                            M, radius, inv, nsample, idx,
                                                            far_flag);
  B2R_CHECK_LAUNCH();
  return B2R_OK;
}
