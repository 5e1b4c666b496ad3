// movers_staged.cu -- shared-memory staged index movers (sm_100a): grouping / three_interpolate
// forward with the source row held in shared memory, and their backward as a DETERMINISTIC,
// atomic-free gather over an inverse index ("scatter plan").
//
// Why (profiles/r02/movers_roofline_before.log): group.cu / interp.cu's forward kernels read the
// source row with one random 4-byte global load per output element -- 32 L1 wavefronts per warp
// instruction, ~1 element/cycle/SM, 0.1-0.45 of the HBM peak; their backward kernels issue one
// RED.ADD.F32 per element, and the L2 atomic units cap that at ~150 G adds/s = 0.03-0.1 of the
// peak no matter how the loads are arranged.  Both are channel-major (B,C,*) movers whose rows
// are independent, so:
//   * forward: a CTA stages the source row f[b,c,:] (N floats) in shared memory with coalesced
//     128-bit loads and gathers from there (random LDS: a few bank conflicts instead of 32
//     wavefronts), writing 128-bit coalesced streaming stores;
//   * backward: the index is inverted ONCE per index tensor (b2r_scatter_plan: counting sort of
//     the entries by (source tile, target), lists sorted by entry id), then every (b,c) row is
//     one CTA that stages tiles of grad_out[b,c,:] in shared memory and, per target, sums its
//     list from shared memory in a fixed order and writes grad_features[b,c,:] coalesced.  No
//     atomics, no memset, run-to-run bit-identical (the reference's atomicAdd order is
//     unspecified, group_points_gpu.cu:48-69, interpolate_gpu.cu:121-148: any order conforms).
#include <stdlib.h>

#include "common.cuh"

namespace b2r {
namespace {

constexpr int kStageThreads = 512;
constexpr int kSmemBudget = 200 * 1024;
constexpr int kMcMaxTargets = 4096;   // multi-channel gather: targets per row it keeps in registers
constexpr int kMcTile = 4096;         // ... and sources per staged tile
constexpr int kMcMaxRows = 8;         // ... and (b,c) rows per CTA (template CC: 4 or 8)

// ------------------------------------------------------------------------- group forward --
// out[b,c,e] = f[b,c,idx[b,e]];  grid (tiles of e, C chunks, B); smem: one source row
__global__ void __launch_bounds__(kStageThreads)
    group_fwd_staged_kernel(const float *__restrict__ f, const int *__restrict__ idx, int C, int N,
                            long long L, int cpb, long long tile, float *__restrict__ out) {
  extern __shared__ __align__(16) float s_row[];
  const int tid = threadIdx.x, b = blockIdx.z;
  const int c0 = blockIdx.y * cpb, c1 = min(C, c0 + cpb);
  const long long e0 = (long long)blockIdx.x * tile, e1 = min(L, e0 + tile);
  const int *ib = idx + (size_t)b * L;
  for (int c = c0; c < c1; ++c) {
    const float *src = f + ((size_t)b * C + c) * N;
    for (int i = tid * 4; i < N; i += kStageThreads * 4)   // N % 4 == 0 (checked by the host)
      *reinterpret_cast<float4 *>(s_row + i) = __ldg(reinterpret_cast<const float4 *>(src + i));
    __syncthreads();
    float *dst = out + ((size_t)b * C + c) * L;
    for (long long e = e0 + tid * 4; e < e1; e += kStageThreads * 4) {
      const int4 a = __ldg(reinterpret_cast<const int4 *>(ib + e));
      float4 v;
      v.x = s_row[a.x]; v.y = s_row[a.y]; v.z = s_row[a.z]; v.w = s_row[a.w];
      stg_stream_v4(reinterpret_cast<float4 *>(dst + e), v);
    }
    __syncthreads();
  }
}

// ----------------------------------------------------------------- three_interpolate fwd --
// out[b,c,j] = fma(p3,w3, fma(p1,w1, p2*w2)) (the reference's contraction, interp.cu); each thread
// keeps the indices / weights of kIpJ outputs in registers; `rows` source rows per smem stage
constexpr int kIpJ = 4;
constexpr int kIpT = 256;
__global__ void __launch_bounds__(kIpT)
    interp_fwd_staged_kernel(const float *__restrict__ f, const int *__restrict__ idx,
                             const float *__restrict__ w, int C, int m, int n, int cpb, int rows,
                             float *__restrict__ out) {
  extern __shared__ __align__(16) float s_rows[];  // [rows][m]
  const int tid = threadIdx.x, b = blockIdx.z;
  const int j0 = blockIdx.x * (kIpT * kIpJ) + tid;
  int a1[kIpJ], a2[kIpJ], a3[kIpJ];
  float w1[kIpJ], w2[kIpJ], w3[kIpJ];
#pragma unroll
  for (int u = 0; u < kIpJ; ++u) {
    const int j = j0 + u * kIpT;
    a1[u] = a2[u] = a3[u] = 0;
    w1[u] = w2[u] = w3[u] = 0.f;
    if (j < n) {
      const int *ip = idx + ((size_t)b * n + j) * 3;
      const float *wp = w + ((size_t)b * n + j) * 3;
      a1[u] = ip[0]; a2[u] = ip[1]; a3[u] = ip[2];
      w1[u] = wp[0]; w2[u] = wp[1]; w3[u] = wp[2];
    }
  }
  const int c0 = blockIdx.y * cpb, c1 = min(C, c0 + cpb);
  for (int cb = c0; cb < c1; cb += rows) {
    const int nr = min(rows, c1 - cb);
    const float *src = f + ((size_t)b * C + cb) * m;   // nr consecutive rows are contiguous
    const int total = nr * m;
    if ((m & 3) == 0) {
      for (int i = tid * 4; i < total; i += kIpT * 4)
        *reinterpret_cast<float4 *>(s_rows + i) = __ldg(reinterpret_cast<const float4 *>(src + i));
    } else {
      for (int i = tid; i < total; i += kIpT) s_rows[i] = __ldg(src + i);
    }
    __syncthreads();
    for (int r = 0; r < nr; ++r) {
      const float *row = s_rows + r * m;
      float *dst = out + ((size_t)b * C + cb + r) * n;
#pragma unroll
      for (int u = 0; u < kIpJ; ++u) {
        const int j = j0 + u * kIpT;
        if (j < n)
          dst[j] = __fmaf_rn(row[a3[u]], w3[u], __fmaf_rn(row[a1[u]], w1[u], __fmul_rn(row[a2[u]], w2[u])));
      }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------- scatter plan ---
// E entries per scene; entry e reads source position e / K (K = 1 grouping, 3 interpolation) and
// adds into target idx[e] in [0,N).  Sources are processed in tiles of TS positions (one smem
// stage); key(e) = tile * N + target.  plan = { int start[B][T*N + 1]; int pos[B][E];
// float wperm[B][E] (weighted only); int cursor[B][T*N] (scratch) }.
// Measured on B200 (profiles/r02/movers_roofline.log): interpolation (3 entries per source: the list
// walk dominates) gains from sharing it between rows at every size; grouping only for short rows.
inline bool use_mc(int N, int K) { return K == 3 ? N <= kMcMaxTargets : N <= 1024; }

struct PlanDims {
  long long E;
  int N, K, S, TS, T;
  long long TN;
  size_t off_pos, off_w, off_cur, bytes;
};

PlanDims plan_dims(int B, long long E, int N, int K, bool weighted) {
  PlanDims d;
  d.E = E; d.N = N; d.K = K;
  d.S = (int)(E / K);
  // multi-channel gather (4 rows of 4096 staged sources share one walk over the lists, sums in
  // registers) where it measured faster, else one row of 16384 sources per CTA
  d.TS = use_mc(N, K) ? kMcTile : 16384;
  d.T = d.S > 0 ? (d.S + d.TS - 1) / d.TS : 1;
  d.TN = (long long)d.T * N;
  size_t o = 0;
  o += sizeof(int) * (size_t)B * (d.TN + 1); o = (o + 255) & ~(size_t)255;
  d.off_pos = o; o += sizeof(int) * (size_t)B * E; o = (o + 255) & ~(size_t)255;
  d.off_w = o; if (weighted) { o += sizeof(float) * (size_t)B * E; o = (o + 255) & ~(size_t)255; }
  d.off_cur = o; o += sizeof(int) * (size_t)B * d.TN; o = (o + 255) & ~(size_t)255;
  d.bytes = o;
  return d;
}

__global__ void __launch_bounds__(256)
    plan_count_kernel(const int *__restrict__ idx, long long E, int N, int K, int TS, long long TN,
                      int *__restrict__ cnt) {
  const long long e = (long long)blockIdx.x * 256 + threadIdx.x;
  if (e >= E) return;
  const int b = blockIdx.y;
  const int t = (int)((e / K) / TS);
  atomicAdd(cnt + (size_t)b * TN + (size_t)t * N + idx[(size_t)b * E + e], 1);
}

// exclusive scan of cnt[b][0..TN) -> start[b][0..TN], start[b][TN] = E; cnt becomes the cursor
__global__ void __launch_bounds__(1024, 1)
    plan_scan_kernel(int *__restrict__ cnt, long long TN, long long E, int *__restrict__ start) {
  __shared__ long long s_wsum[32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  cnt += (size_t)blockIdx.x * TN;
  start += (size_t)blockIdx.x * (TN + 1);
  const long long seg = ((TN + 31) / 32 + 31) / 32 * 32;   // per warp, a multiple of 32
  const long long w0 = (long long)warp * seg;
  long long total = 0;
  for (long long r = 0; r < seg; r += 32) {
    const long long i = w0 + r + lane;
    const int v = i < TN ? cnt[i] : 0;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int nb = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += nb;
    }
    if (i < TN) start[i] = (int)(inc - v + total);   // warp-local exclusive prefix for now
    total += __shfl_sync(0xffffffffu, inc, 31);
  }
  if (lane == 0) s_wsum[warp] = total;
  __syncthreads();
  if (warp == 0) {
    const long long v = s_wsum[lane];
    long long inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const long long nb = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += nb;
    }
    s_wsum[lane] = inc - v;
  }
  __syncthreads();
  const int off = (int)s_wsum[warp];
  for (long long r = 0; r < seg; r += 32) {
    const long long i = w0 + r + lane;
    if (i < TN) {
      const int s = start[i] + off;
      start[i] = s;
      cnt[i] = s;
    }
  }
  if (tid == 0) start[TN] = (int)E;
}

__global__ void __launch_bounds__(256)
    plan_fill_kernel(const int *__restrict__ idx, long long E, int N, int K, int TS, long long TN,
                     int *__restrict__ cur, int *__restrict__ pos) {
  const long long e = (long long)blockIdx.x * 256 + threadIdx.x;
  if (e >= E) return;
  const int b = blockIdx.y;
  const int t = (int)((e / K) / TS);
  const int q = atomicAdd(cur + (size_t)b * TN + (size_t)t * N + idx[(size_t)b * E + e], 1);
  pos[(size_t)b * E + q] = (int)e;
}

// every list sorted by entry id (the fill order is whatever the atomics produced): the sums of
// scatter_gather_kernel are then evaluated in one fixed order, run after run
__global__ void __launch_bounds__(256)
    plan_sort_kernel(const int *__restrict__ start, long long TN, long long E, int *__restrict__ pos,
                     const float *__restrict__ w, float *__restrict__ wperm) {
  const long long k = (long long)blockIdx.x * 256 + threadIdx.x;
  if (k >= TN) return;
  const int b = blockIdx.y;
  const int *st = start + (size_t)b * (TN + 1);
  int *p = pos + (size_t)b * E;
  const int s0 = st[k], s1 = st[k + 1];
  for (int i = s0 + 1; i < s1; ++i) {
    const int v = p[i];
    int j = i - 1;
    while (j >= s0 && p[j] > v) { p[j + 1] = p[j]; --j; }
    p[j + 1] = v;
  }
  if (wperm != nullptr)
    for (int i = s0; i < s1; ++i) wperm[(size_t)b * E + i] = w[(size_t)b * E + p[i]];
}

// one CTA per (b,c) row: gf[b,c,n] = sum over the entries of n of g[b,c,entry / K] (* weight).
// G lanes share a target: they read G consecutive entries of its list (coalesced pos / weight
// loads -- one thread per list would touch 32 different lines per warp instruction and make the
// kernel LSU-bound), gather from the staged tile and combine with a fixed shuffle tree.
template <int K, bool WEIGHTED, int G>
__global__ void __launch_bounds__(kStageThreads)
    scatter_gather_kernel(const float *__restrict__ g, const int *__restrict__ start,
                          const int *__restrict__ pos, const float *__restrict__ wperm, int C, int N,
                          long long E, int S, int TS, int T, float *__restrict__ gf) {
  extern __shared__ __align__(16) float s_g[];
  const int tid = threadIdx.x, b = blockIdx.y, c = blockIdx.x;
  const float *row = g + ((size_t)b * C + c) * S;
  float *orow = gf + ((size_t)b * C + c) * N;
  const long long TN = (long long)T * N;
  start += (size_t)b * (TN + 1);
  pos += (size_t)b * E;
  if (WEIGHTED) wperm += (size_t)b * E;
  const bool vec = ((S & 3) == 0) && ((reinterpret_cast<uintptr_t>(row) & 15) == 0);
  const int sub = tid % G, grp = tid / G;
  constexpr int kGroups = kStageThreads / G;
  const int Nr = (N + kGroups - 1) / kGroups * kGroups;   // whole warps stay in the shuffle tree
  for (int t = 0; t < T; ++t) {
    const int base = t * TS, len = min(TS, S - base);
    if (vec) {
      for (int i = tid * 4; i < len; i += kStageThreads * 4)
        *reinterpret_cast<float4 *>(s_g + i) =
            __ldcs(reinterpret_cast<const float4 *>(row + base + i));
    } else {
      for (int i = tid; i < len; i += kStageThreads) s_g[i] = __ldcs(row + base + i);
    }
    __syncthreads();
    const int *st = start + (size_t)t * N;
    for (int n = grp; n < Nr; n += kGroups) {
      int s0 = 0, s1 = 0;
      if (n < N) { s0 = st[n]; s1 = st[n + 1]; }
      float acc = 0.f;
      for (int q = s0 + sub; q < s1; q += G) {
        float v = s_g[pos[q] / K - base];
        if (WEIGHTED) v = __fmul_rn(v, wperm[q]);
        acc += v;
      }
#pragma unroll
      for (int o = G / 2; o >= 1; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if (sub == 0 && n < N) {
        if (t == 0) orow[n] = acc;
        else if (s1 > s0) orow[n] += acc;
      }
    }
    __syncthreads();
  }
}

// Multi-channel variant for rows of <= 4096 targets (every level of the detectors but SA1): a CTA
// owns kMcRows consecutive channel rows of one scene.  The walk over a target's list -- the
// uncoalesced part: one lane per list -- is done ONCE for all rows (8x less index traffic and LSU
// pressure per byte of gradient), sums stay in registers across the source tiles (KT targets per
// thread), and the rows are written once at the end.  Same fixed summation order as above.
template <int K, bool WEIGHTED, int KT, int kMcRows>
__global__ void __launch_bounds__(kStageThreads)
    scatter_gather_mc_kernel(const float *__restrict__ g, const int *__restrict__ start,
                             const int *__restrict__ pos, const float *__restrict__ wperm, int C, int N,
                             long long E, int S, int T, float *__restrict__ gf) {
  extern __shared__ __align__(16) float s_g[];   // [kMcRows][kMcTile]
  const int tid = threadIdx.x, b = blockIdx.y, c0 = blockIdx.x * kMcRows;
  const int nr = min(kMcRows, C - c0);
  const float *rows = g + ((size_t)b * C + c0) * S;
  const long long TN = (long long)T * N;
  start += (size_t)b * (TN + 1);
  pos += (size_t)b * E;
  if (WEIGHTED) wperm += (size_t)b * E;
  const bool vec = ((S & 3) == 0) && ((reinterpret_cast<uintptr_t>(rows) & 15) == 0);
  float acc[KT][kMcRows];
#pragma unroll
  for (int k = 0; k < KT; ++k)
#pragma unroll
    for (int r = 0; r < kMcRows; ++r) acc[k][r] = 0.f;
  for (int t = 0; t < T; ++t) {
    const int base = t * kMcTile, len = min(kMcTile, S - base);
    for (int r = 0; r < nr; ++r) {
      const float *src = rows + (size_t)r * S + base;
      float *dst = s_g + r * kMcTile;
      if (vec) {
        for (int i = tid * 4; i < len; i += kStageThreads * 4)
          *reinterpret_cast<float4 *>(dst + i) = __ldcs(reinterpret_cast<const float4 *>(src + i));
      } else {
        for (int i = tid; i < len; i += kStageThreads) dst[i] = __ldcs(src + i);
      }
    }
    __syncthreads();
    const int *st = start + (size_t)t * N;
#pragma unroll
    for (int k = 0; k < KT; ++k) {
      const int n = tid + k * kStageThreads;
      if (n < N) {
        const int s0 = st[n], s1 = st[n + 1];
        for (int q = s0; q < s1; ++q) {
          const int p = pos[q] / K - base;
          float w = 1.f;
          if (WEIGHTED) w = wperm[q];
#pragma unroll
          for (int r = 0; r < kMcRows; ++r) {
            float v = s_g[r * kMcTile + p];      // rows >= nr hold stale data: never stored
            if (WEIGHTED) v = __fmul_rn(v, w);
            acc[k][r] += v;
          }
        }
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int k = 0; k < KT; ++k) {
    const int n = tid + k * kStageThreads;
    if (n < N)
#pragma unroll
      for (int r = 0; r < kMcRows; ++r)
        if (r < nr) gf[((size_t)b * C + c0 + r) * N + n] = acc[k][r];
  }
}

bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

// called by b2r_group_fwd (group.cu): true if the staged kernel took the call
bool group_fwd_staged(const float *f, const int *idx, int B, int C, int N, long long L, float *out,
                      cudaStream_t st, cudaError_t *err) {
  *err = cudaSuccess;
  if ((N & 3) || (L & 3) || (size_t)N * 4 > (size_t)kSmemBudget || !aligned16(f) || !aligned16(idx) ||
      !aligned16(out))
    return false;
  int cpb = 8;
  while (cpb > 1 && (long long)B * ((C + cpb - 1) / cpb) < 2LL * kNumSMs) cpb >>= 1;
  const long long rows = (long long)B * ((C + cpb - 1) / cpb);
  long long T = (2LL * kNumSMs + rows - 1) / rows;
  const long long tmax = (2 * L) / (N > 0 ? N : 1);   // a tile writes at least half a source row
  if (T > tmax) T = tmax;
  if (T < 1) T = 1;
  const long long unit = 4LL * kStageThreads;
  long long tile = ((L + T - 1) / T + unit - 1) / unit * unit;
  T = (L + tile - 1) / tile;
  static bool attr_done[64] = {};
  int dev = 0;
  if ((*err = cudaGetDevice(&dev)) != cudaSuccess) return true;
  if (dev < 0 || dev >= 64 || !attr_done[dev]) {
    *err = cudaFuncSetAttribute(group_fwd_staged_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                kSmemBudget);
    if (*err != cudaSuccess) return true;
    if (dev >= 0 && dev < 64) attr_done[dev] = true;
  }
  dim3 grid((unsigned)T, (unsigned)((C + cpb - 1) / cpb), (unsigned)B);
  group_fwd_staged_kernel<<<grid, kStageThreads, (size_t)N * 4, st>>>(f, idx, C, N, L, cpb, tile, out);
  *err = cudaGetLastError();
  return true;
}

// called by b2r_three_interp_fwd (interp.cu)
bool interp_fwd_staged(const float *f, const int *idx, const float *w, int B, int C, int m, int n,
                       float *out, cudaStream_t st, cudaError_t *err) {
  *err = cudaSuccess;
  if ((size_t)m * 4 > (size_t)kSmemBudget || m <= 0 || !aligned16(f)) return false;
  int rows = (48 * 1024) / (m * 4);
  rows = rows < 1 ? 1 : (rows > 8 ? 8 : rows);
  const int xb = (n + kIpT * kIpJ - 1) / (kIpT * kIpJ);
  int cpb = 32;
  while (cpb > rows && (long long)xb * ((C + cpb - 1) / cpb) * B < 2LL * kNumSMs) cpb >>= 1;
  if (cpb < rows) cpb = rows;
  static bool attr_done[64] = {};
  int dev = 0;
  if ((*err = cudaGetDevice(&dev)) != cudaSuccess) return true;
  if (dev < 0 || dev >= 64 || !attr_done[dev]) {
    *err = cudaFuncSetAttribute(interp_fwd_staged_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                kSmemBudget);
    if (*err != cudaSuccess) return true;
    if (dev >= 0 && dev < 64) attr_done[dev] = true;
  }
  dim3 grid((unsigned)xb, (unsigned)((C + cpb - 1) / cpb), (unsigned)B);
  interp_fwd_staged_kernel<<<grid, kIpT, (size_t)rows * m * 4, st>>>(f, idx, w, C, m, n, cpb, rows, out);
  *err = cudaGetLastError();
  return true;
}

}  // namespace b2r

using namespace b2r;

extern "C" long long b2r_scatter_plan_bytes(int B, long long entries, int N, int entries_per_source,
                                            int weighted) {
  if (B <= 0 || entries <= 0 || N <= 0 || entries_per_source <= 0) return 256;
  return (long long)plan_dims(B, entries, N, entries_per_source, weighted != 0).bytes;
}

extern "C" int b2r_scatter_plan(const int *idx, const float *weight, int B, long long entries, int N,
                                int entries_per_source, void *plan, long long plan_bytes,
                                void *stream) {
  B2R_REQUIRE(B >= 0 && entries >= 0 && N >= 0, "b2r_scatter_plan: negative size");
  B2R_REQUIRE(entries_per_source == 1 || entries_per_source == 3,
              "b2r_scatter_plan: entries_per_source=%d (1: grouping, 3: three_interpolate)",
              entries_per_source);
  if (B == 0 || entries == 0 || N == 0) return B2R_OK;
  B2R_REQUIRE(idx && plan, "b2r_scatter_plan: null pointer");
  B2R_REQUIRE(B <= 65535, "b2r_scatter_plan: B too large");
  B2R_REQUIRE(entries % entries_per_source == 0 && entries < (1LL << 31),
              "b2r_scatter_plan: entries=%lld", entries);
  const PlanDims d = plan_dims(B, entries, N, entries_per_source, weight != nullptr);
  B2R_REQUIRE(plan_bytes >= (long long)d.bytes, "b2r_scatter_plan: plan of %lld bytes, %lld needed",
              plan_bytes, (long long)d.bytes);
  B2R_REQUIRE(d.TN < (1LL << 31), "b2r_scatter_plan: too many (tile, target) keys");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  char *base = static_cast<char *>(plan);
  int *start = reinterpret_cast<int *>(base);
  int *pos = reinterpret_cast<int *>(base + d.off_pos);
  float *wperm = weight ? reinterpret_cast<float *>(base + d.off_w) : nullptr;
  int *cur = reinterpret_cast<int *>(base + d.off_cur);
  B2R_CUDA(cudaMemsetAsync(cur, 0, sizeof(int) * (size_t)B * d.TN, st));
  dim3 ge((unsigned)((entries + 255) / 256), (unsigned)B);
  plan_count_kernel<<<ge, 256, 0, st>>>(idx, entries, N, d.K, d.TS, d.TN, cur);
  B2R_CHECK_LAUNCH();
  plan_scan_kernel<<<B, 1024, 0, st>>>(cur, d.TN, entries, start);
  B2R_CHECK_LAUNCH();
  plan_fill_kernel<<<ge, 256, 0, st>>>(idx, entries, N, d.K, d.TS, d.TN, cur, pos);
  B2R_CHECK_LAUNCH();
  dim3 gk((unsigned)((d.TN + 255) / 256), (unsigned)B);
  plan_sort_kernel<<<gk, 256, 0, st>>>(start, d.TN, entries, pos, weight, wperm);
  B2R_CHECK_LAUNCH();
  return B2R_OK;
}

namespace {
template <int K, bool WEIGHTED, int G>
int launch_gather_g(const PlanDims &d, const float *g, const void *plan, int B, int C, int N,
                    long long E, float *gf, cudaStream_t st) {
  const char *base = static_cast<const char *>(plan);
  const int *start = reinterpret_cast<const int *>(base);
  const int *pos = reinterpret_cast<const int *>(base + d.off_pos);
  const float *wperm = WEIGHTED ? reinterpret_cast<const float *>(base + d.off_w) : nullptr;
  auto kern = scatter_gather_kernel<K, WEIGHTED, G>;
  static bool attr_done[64] = {};
  int dev = 0;
  B2R_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || !attr_done[dev]) {
    B2R_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, d.TS * 4));
    if (dev >= 0 && dev < 64) attr_done[dev] = true;
  }
  const int ts = d.S < d.TS ? (d.S + 3) / 4 * 4 : d.TS;
  dim3 grid((unsigned)C, (unsigned)B);
  kern<<<grid, kStageThreads, (size_t)ts * 4, st>>>(g, start, pos, wperm, C, N, E, d.S, d.TS, d.T, gf);
  B2R_CHECK_LAUNCH();
  return B2R_OK;
}

template <int K, bool WEIGHTED, int KT, int kMcRows>
int launch_gather_mc(const PlanDims &d, const float *g, const void *plan, int B, int C, int N,
                     long long E, float *gf, cudaStream_t st) {
  const char *base = static_cast<const char *>(plan);
  const int *start = reinterpret_cast<const int *>(base);
  const int *pos = reinterpret_cast<const int *>(base + d.off_pos);
  const float *wperm = WEIGHTED ? reinterpret_cast<const float *>(base + d.off_w) : nullptr;
  auto kern = scatter_gather_mc_kernel<K, WEIGHTED, KT, kMcRows>;
  constexpr int smem = kMcRows * kMcTile * 4;
  static bool attr_done[64] = {};
  int dev = 0;
  B2R_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || !attr_done[dev]) {
    B2R_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    if (dev >= 0 && dev < 64) attr_done[dev] = true;
  }
  dim3 grid((unsigned)((C + kMcRows - 1) / kMcRows), (unsigned)B);
  kern<<<grid, kStageThreads, smem, st>>>(g, start, pos, wperm, C, N, E, d.S, d.T, gf);
  B2R_CHECK_LAUNCH();
  return B2R_OK;
}

// One lane per target.  (Measured, profiles/r02/movers_roofline.log: 8- and 32-lane groups with a
// shuffle tree read pos / weight coalesced but leave 8-32x fewer independent chains in flight;
// at the detectors' shapes they were 1.5-2x SLOWER than one lane per target.)
template <int K, bool WEIGHTED>
int launch_gather(const float *g, const void *plan, int B, int C, int N, long long E, float *gf,
                  cudaStream_t st) {
  const PlanDims d = plan_dims(B, E, N, K, WEIGHTED);
  if (use_mc(N, K)) {
    // rows per CTA: 4 keep three CTAs resident per SM (8 rows, one CTA per SM, measured 25-40 %
    // slower although they halve the index traffic again)
    const int kt = N <= 512 ? 1 : N <= 1024 ? 2 : N <= 2048 ? 4 : 8;
    if (kt == 1) return launch_gather_mc<K, WEIGHTED, 1, 4>(d, g, plan, B, C, N, E, gf, st);
    if (kt == 2) return launch_gather_mc<K, WEIGHTED, 2, 4>(d, g, plan, B, C, N, E, gf, st);
    if (kt == 4) return launch_gather_mc<K, WEIGHTED, 4, 4>(d, g, plan, B, C, N, E, gf, st);
    return launch_gather_mc<K, WEIGHTED, 8, 4>(d, g, plan, B, C, N, E, gf, st);
  }
  return launch_gather_g<K, WEIGHTED, 1>(d, g, plan, B, C, N, E, gf, st);
}
}  // namespace

extern "C" int b2r_group_bwd_plan(const float *grad_out, const void *plan, int B, int C, int N, int NP,
                                  int NS, float *grad_features, void *stream) {
  B2R_REQUIRE(B >= 0 && C >= 0 && N >= 0 && NP >= 0 && NS >= 0, "b2r_group_bwd_plan: negative size");
  const long long L = (long long)NP * NS;
  if (B == 0 || C == 0 || N == 0) return B2R_OK;
  B2R_REQUIRE(grad_features, "b2r_group_bwd_plan: null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (L == 0) {
    B2R_CUDA(cudaMemsetAsync(grad_features, 0, sizeof(float) * (size_t)B * C * N, st));
    return B2R_OK;
  }
  B2R_REQUIRE(grad_out && plan, "b2r_group_bwd_plan: null pointer");
  B2R_REQUIRE(B <= 65535, "b2r_group_bwd_plan: B too large");
  return launch_gather<1, false>(grad_out, plan, B, C, N, L, grad_features, st);
}

extern "C" int b2r_three_interp_bwd_plan(const float *grad_out, const void *plan, int B, int C, int n,
                                         int m, float *grad_features, void *stream) {
  B2R_REQUIRE(B >= 0 && C >= 0 && m >= 0 && n >= 0, "b2r_three_interp_bwd_plan: negative size");
  if (B == 0 || C == 0 || m == 0) return B2R_OK;
  B2R_REQUIRE(grad_features, "b2r_three_interp_bwd_plan: null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (n == 0) {
    B2R_CUDA(cudaMemsetAsync(grad_features, 0, sizeof(float) * (size_t)B * C * m, st));
    return B2R_OK;
  }
  B2R_REQUIRE(grad_out && plan, "b2r_three_interp_bwd_plan: null pointer");
  B2R_REQUIRE(B <= 65535, "b2r_three_interp_bwd_plan: B too large");
  return launch_gather<3, true>(grad_out, plan, B, C, m, 3LL * n, grad_features, st);
}
