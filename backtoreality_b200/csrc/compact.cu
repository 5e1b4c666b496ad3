// compact.cu -- pad-free position space for the fused set-abstraction block.
//
// The reference's ball query (src/ball_query_gpu.cu:14-49) fills the slots of a centre that it
// found no neighbour for with COPIES of the first hit ("pad-with-first"), and everything after it
// -- QueryAndGroup, three Conv2d/BatchNorm/ReLU layers, max_pool2d (pointnet2_modules.py:245-267)
// -- computes those copies like any other sample.  On ScanNet-shaped scenes that is most of the
// work: SA1 (r = 0.2, nsample 64) finds ~35 neighbours, SA2 (r = 0.4, nsample 32) ~9.
//
// A copy of sample 0 has the same input row as sample 0, hence the same activations in every
// layer: it can never win the max-pool against sample 0 (ties go to the first) and it enters the
// BatchNorm batch statistics -- and, in backward, the BatchNorm-backward sums and the weight
// gradients -- exactly like sample 0 again.  So a centre with `cnt` distinct leading samples is
// computed on its first u = 8/16/32/64 >= cnt samples only, and sample 0 carries the weight
// 1 + (nsample - u) wherever a sum over positions is taken.  Results are those of the padded
// computation up to fp32 summation order.
//
// This file builds that position space once per ball query ("plan"):
//   * cnt   = 1 + last s with idx[s] != idx[0]   (valid for ANY idx, ball query or not)
//   * class = smallest of 8/16/32/64 >= cnt; centres are ordered by (class, centre id): every
//     class is one contiguous range of positions, padded with dead positions to a multiple of
//     128, so a tile of 32/64/128 positions never straddles a class or splits a centre
//     unevenly, and the centre-local sample index of a position is  position & (class - 1)
//   * cidx[p] = global source row (b*N + idx), ccen[p] = global centre id (b*NP + j), -1 dead
//   * meta    = class ends, live ends, total (read by the kernels: tile counts are data-dependent
//     and live on the device, so the whole block stays CUDA-graph capturable)
#include "common.cuh"

namespace b2r {
namespace {

constexpr int kScanThreads = 1024;

__device__ __forceinline__ int class_of(int cnt) {
  return cnt <= 8 ? 0 : cnt <= 16 ? 1 : cnt <= 32 ? 2 : 3;
}

// one warp per centre: number of leading samples that are not copies of sample 0
__global__ void compact_count_kernel(const int *__restrict__ idx, int G, int NS,
                                     unsigned char *__restrict__ cls) {
  const int w = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (w >= G) return;
  const int *row = idx + (size_t)w * NS;
  const int first = __ldg(row);
  int last = 0;
  for (int s = lane; s < NS; s += 32)
    if (__ldg(row + s) != first) last = s;   // s ascends per lane: the last hit is the largest
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) last = max(last, __shfl_xor_sync(0xffffffffu, last, o));
  if (lane == 0) cls[w] = (unsigned char)class_of(last + 1);
}

// one CTA: rank of every centre inside its class (ordered by centre id) -> first position
__global__ void __launch_bounds__(kScanThreads) compact_scan_kernel(
    const unsigned char *__restrict__ cls, int G, int *__restrict__ base, int *__restrict__ meta) {
  __shared__ int s_warp[4][kScanThreads / 32];
  __shared__ int s_start[4];
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int per = (G + kScanThreads - 1) / kScanThreads;
  const int g0 = min(G, t * per), g1 = min(G, g0 + per);
  int cnt[4] = {0, 0, 0, 0};
  for (int g = g0; g < g1; ++g) {
    const int c = cls[g];
#pragma unroll
    for (int k = 0; k < 4; ++k) cnt[k] += (c == k);
  }
  int excl[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {   // inclusive warp scan, then the warp totals
    int v = cnt[k];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int u = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += u;
    }
    excl[k] = v - cnt[k];
    if (lane == 31) s_warp[k][warp] = v;
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int own = s_warp[k][lane];
      int v = own;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += u;
      }
      s_warp[k][lane] = v - own;   // exclusive prefix of the warp totals
      if (lane == 31) {
        // total of class k is v; positions of the class, padded to 128
        s_start[k] = v;
      }
    }
  }
  __syncthreads();
  if (t == 0) {
    int start = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int n = s_start[k];
      const int live = start + n * (8 << k);
      const int end = (live + 127) & ~127;
      meta[k] = end;
      meta[4 + k] = live;
      meta[10 + k] = n;
      s_start[k] = start;
      start = end;
    }
    meta[8] = start;
    meta[9] = 0;
    meta[14] = meta[15] = 0;
  }
  __syncthreads();
  int run[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) run[k] = s_start[k] + (s_warp[k][warp] + excl[k]) * (8 << k);
  for (int g = g0; g < g1; ++g) {
    const int c = cls[g];
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (c == k) {
        base[g] = run[k];
        run[k] += 8 << k;
      }
  }
}

// one warp per centre writes its class-size run; four more warps write the dead tails
__global__ void compact_fill_kernel(const int *__restrict__ idx,
                                    const unsigned char *__restrict__ cls,
                                    const int *__restrict__ base, const int *__restrict__ meta,
                                    int G, int N, int NP, int NS, int *__restrict__ cidx,
                                    int *__restrict__ ccen) {
  const int w = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (w < G) {
    const int ns = 8 << cls[w];
    const int p0 = base[w];
    const int row0 = (w / NP) * N;
    const int *row = idx + (size_t)w * NS;
    for (int s = lane; s < ns; s += 32) {
      // s < NS always: a class never exceeds nsample (cnt <= NS, NS in {16,32,64})
      cidx[p0 + s] = row0 + __ldg(row + s);
      ccen[p0 + s] = w;
    }
  } else if (w < G + 4) {
    const int k = w - G;
    for (int p = meta[4 + k] + lane; p < meta[k]; p += 32) {
      cidx[p] = 0;
      ccen[p] = -1;
    }
  }
}

}  // namespace
}  // namespace b2r

using namespace b2r;

extern "C" long long b2r_compact_capacity(int B, int NP, int NS) {
  if (B <= 0 || NP <= 0 || NS <= 0) return 0;
  const long long M = (long long)B * NP * NS;
  return ((M + 127) & ~127ll) + 4 * 128;   // every class range is padded to 128 positions
}

extern "C" long long b2r_compact_workspace_bytes(int B, int NP) {
  if (B <= 0 || NP <= 0) return 0;
  const long long G = (long long)B * NP;
  return ((G + 15) & ~15ll) + 4 * G;   // class per centre (bytes), first position per centre
}

extern "C" int b2r_compact_plan(const int *idx, int B, int N, int NP, int NS, int *cidx, int *ccen,
                                int *meta, void *workspace, void *stream) {
  B2R_REQUIRE(idx && cidx && ccen && meta && workspace, "b2r_compact_plan: null pointer");
  B2R_REQUIRE(B > 0 && N > 0 && NP > 0, "b2r_compact_plan: non-positive size");
  if (!(NS == 16 || NS == 32 || NS == 64)) {
    set_error("b2r_compact_plan: nsample must be 16, 32 or 64 (got %d)", NS);
    return B2R_ERR_UNSUPPORTED;
  }
  const long long G = (long long)B * NP;
  B2R_REQUIRE(G * NS + 640 < 2147483647ll && (long long)B * N < 2147483647ll,
              "b2r_compact_plan: position / row index exceeds int32");
  unsigned char *cls = static_cast<unsigned char *>(workspace);
  int *base = reinterpret_cast<int *>(cls + ((G + 15) & ~15ll));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  compact_count_kernel<<<ceil_div(G * 32, 256), 256, 0, st>>>(idx, (int)G, NS, cls);
  B2R_CHECK_LAUNCH();
  compact_scan_kernel<<<1, kScanThreads, 0, st>>>(cls, (int)G, base, meta);
  B2R_CHECK_LAUNCH();
  compact_fill_kernel<<<ceil_div((G + 4) * 32, 256), 256, 0, st>>>(idx, cls, base, meta, (int)G,
                                                                  N, NP, NS, cidx, ccen);
  B2R_CHECK_LAUNCH();
  return B2R_OK;
}
