// mlp_thin.cu -- the FIRST SharedMLP layer of a set-abstraction block when its input is thin
// (Cin <= 8: relative xyz + at most five feature channels -- SA1 of both detectors: [dx,dy,dz,
// height] -> 64 for VoteNet, [dx,dy,dz] -> 64 for GroupFree3D).
//
// With K <= 8 the layer is not a GEMM worth a tensor core: per position it is Cout x K <= 512
// FMAs against 4*Cout bytes of HBM traffic written (forward) or 8*Cout bytes read (backward), so
// it is a pure streaming kernel.  Running it through the tcgen05 pipeline of mlp.cu / mlp_bwd.cu
// (smem operand tiles, TMEM round trip, 4 epilogue warps doing all the global traffic) measured
// 0.35 (forward) and 0.20 (backward) of the HBM roofline; here every warp streams whole 128-bit
// rows itself:
//   * one lane gathers one position's input row (ball-query index -> xyz - centre (/radius),
//     features), rounds it to TF32 (forward: the same operand values the tensor-core path uses)
//     and the row is broadcast by shuffles to the lanes that own that position's channels;
//   * LP = Cout/4 lanes own one position (4 channels each), so one warp instruction moves
//     32/LP complete rows = 512 contiguous bytes;
//   * forward: z = W x in registers, streamed out with st.global.v4, BatchNorm sums kept per
//     lane (fp32 per 32-position chunk, then double) and reduced once per CTA;
//   * backward (no input gradient needed -- xyz / height are leaves): dz = a*gr + b*z + c from
//     two 128-bit streams, dW += dz x^T accumulated in registers, reduced once per CTA.
// Replaces the same reference code as mlp.cu (pointnet2_utils.py:347-366 QueryAndGroup tail +
// pytorch_utils.py:11-36 first Conv2d) for these shapes; b2r_sa_layer_fwd / b2r_sa_layer_bwd
// dispatch here, the C ABI does not change.  B2R_NO_THIN=1 keeps the tensor-core path (A/B tests).
#include <stdlib.h>

#include "mlp_common.cuh"

namespace b2r {
namespace thin {
namespace {

using mlp::sw128_off;
using mlp::to_tf32;

constexpr int kThreads = 256, kWarps = kThreads / 32;
constexpr int kCtasPerSM = 2;   // ~128 registers x 256 threads
constexpr int kMaxK = 8;

struct ThinArgs {
  int N, NP, NS, C, K, Cout;            // C = feature channels, K = 3 + C
  long long M, per_scene;               // positions, positions per scene
  const float *xyz, *new_xyz, *feat_t;
  const int *idx;
  float radius;
  int normalize_xyz;
  // forward
  const float *w_image;                 // TF32 image of mlp.cu's pack_weight_kernel
  int Cout_pad, Cf4;
  float *z;
  double *stats;
  // backward
  const float *gr, *zin, *coef_a, *coef_b, *coef_c;
  float *dW;
  // compacted position space (csrc/compact.cu) or NULL; Mcap = its capacity (sizes the grid)
  const int *cidx, *ccen, *cmeta;
};

// one lane = one position of the warp's 32-position chunk: its input row in the reference's
// channel order [dx, dy, dz, f0 .. fC-1]
template <int KM, bool CMP>
__device__ __forceinline__ void gather_row(const ThinArgs &a, long long pos, float (&x)[KM]) {
  size_t prow;        // global source row
  long long centre;   // global centre
  if constexpr (CMP) {
    prow = (size_t)__ldg(a.cidx + pos);
    centre = max(__ldg(a.ccen + pos), 0);   // dead padding rows: any valid centre
  } else {
    const int b = (int)(pos / a.per_scene);
    centre = pos / a.NS;
    prow = (size_t)b * a.N + __ldg(a.idx + pos);
  }
  const float *pp = a.xyz + prow * 3;
  const float *qq = a.new_xyz + (size_t)centre * 3;
  float d0 = __fsub_rn(__ldg(pp), __ldg(qq));
  float d1 = __fsub_rn(__ldg(pp + 1), __ldg(qq + 1));
  float d2 = __fsub_rn(__ldg(pp + 2), __ldg(qq + 2));
  if (a.normalize_xyz) {
    d0 = __fdiv_rn(d0, a.radius);
    d1 = __fdiv_rn(d1, a.radius);
    d2 = __fdiv_rn(d2, a.radius);
  }
  x[0] = d0; x[1] = d1; x[2] = d2;
  const float *ff = a.feat_t + prow * a.C;
#pragma unroll
  for (int c = 0; c < KM - 3; ++c) x[3 + c] = c < a.C ? __ldg(ff + c) : 0.f;
}

// LP lanes per position (Cout = 4 * LP); PPW = 32 / LP positions per warp instruction;
// KM = register-array extent of the input row (4 for Cin <= 4, else 8)
template <int LP, int KM, bool CMP>
__global__ void __launch_bounds__(kThreads, kCtasPerSM) thin_fwd_kernel(const ThinArgs a) {
  constexpr int PPW = 32 / LP, ITERS = 32 / PPW;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane / LP, c4 = (lane % LP) * 4;   // position within the instruction, channel
  __shared__ double s_red[kWarps][2][LP * 4];
  __shared__ int s_meta[8];
  long long M = a.M;
  if constexpr (CMP) {
    M = __ldg(a.cmeta + 8);
    if (threadIdx.x < 8) s_meta[threadIdx.x] = __ldg(a.cmeta + threadIdx.x);
    __syncthreads();
  }

  // this lane's 4 x K weights, from the TF32 image (packed K order: features, pad, dx dy dz)
  float w[4][KM];
#pragma unroll
  for (int e = 0; e < 4; ++e)
#pragma unroll
    for (int k = 0; k < KM; ++k) {
      float v = 0.f;
      if (k < a.K) {
        const int kp = k < 3 ? a.Cf4 + k : k - 3;
        v = a.w_image[(sw128_off(c4 + e, kp >> 2, a.Cout_pad) >> 2) + (kp & 3)];
      }
      w[e][k] = v;
    }

  double acc_s[4] = {0, 0, 0, 0}, acc_ss[4] = {0, 0, 0, 0};
  const long long nchunks = M >> 5, cstride = (long long)gridDim.x * kWarps;
  long long chunk = (long long)blockIdx.x * kWarps + warp;
  float xn[KM];   // the NEXT chunk's row: its dependent index -> point loads overlap this chunk
  if (chunk < nchunks) gather_row<KM, CMP>(a, (chunk << 5) + lane, xn);
  for (; chunk < nchunks; chunk += cstride) {
    const long long pos0 = chunk << 5;
    float x[KM];
#pragma unroll
    for (int k = 0; k < KM; ++k) x[k] = __uint_as_float(to_tf32(xn[k]));
    if (chunk + cstride < nchunks) gather_row<KM, CMP>(a, ((chunk + cstride) << 5) + lane, xn);
    mlp::TileClass tc{};
    int p64 = 0;
    if constexpr (CMP) {
      tc = mlp::tile_class(s_meta, pos0, a.NS);
      p64 = (int)(pos0 & 63);
    }
    float ts[4] = {0, 0, 0, 0}, tss[4] = {0, 0, 0, 0};
    float *zrow = a.z + (size_t)pos0 * a.Cout + c4;
#pragma unroll 4
    for (int it = 0; it < ITERS; ++it) {
      const int src = it * PPW + sub;
      float xr[KM];
#pragma unroll
      for (int k = 0; k < KM; ++k) xr[k] = __shfl_sync(0xffffffffu, x[k], src);
      float zz[4];
      float wgt = 1.f;   // CMP: multiplicity of this position in sums over the padded positions
      if constexpr (CMP)
        wgt = src < tc.live ? ((((p64 + src) & (tc.ns - 1)) == 0) ? 1.f + tc.wx : 1.f) : 0.f;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < KM; ++k)
          if (k < a.K) s = fmaf(w[e][k], xr[k], s);
        zz[e] = s;
        if constexpr (CMP) {
          ts[e] = fmaf(wgt, s, ts[e]);
          tss[e] = fmaf(wgt * s, s, tss[e]);
        } else {
          ts[e] += s;
          tss[e] = fmaf(s, s, tss[e]);
        }
      }
      stg_stream_v4(reinterpret_cast<float4 *>(zrow + (size_t)src * a.Cout),
                    make_float4(zz[0], zz[1], zz[2], zz[3]));
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      acc_s[e] += (double)ts[e];
      acc_ss[e] += (double)tss[e];
    }
  }
  if (a.stats == nullptr) return;
  // lanes sub = 0..PPW-1 hold the same channels: fold them, then the warps, then one atomic per
  // channel per CTA
#pragma unroll
  for (int e = 0; e < 4; ++e) {
#pragma unroll
    for (int o = LP; o < 32; o <<= 1) {
      acc_s[e] += __shfl_xor_sync(0xffffffffu, acc_s[e], o);
      acc_ss[e] += __shfl_xor_sync(0xffffffffu, acc_ss[e], o);
    }
    if (sub == 0) {
      s_red[warp][0][c4 + e] = acc_s[e];
      s_red[warp][1][c4 + e] = acc_ss[e];
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * LP * 4; i += kThreads) {
    const int which = i / (LP * 4), c = i % (LP * 4);
    double t = 0.0;
#pragma unroll
    for (int wv = 0; wv < kWarps; ++wv) t += s_red[wv][which][c];
    atomicAdd(a.stats + (size_t)which * a.Cout + c, t);
  }
}

template <int LP, int KM, bool CMP>
__global__ void __launch_bounds__(kThreads, kCtasPerSM) thin_bwd_kernel(const ThinArgs a) {
  constexpr int PPW = 32 / LP, ITERS = 32 / PPW, UN = 4;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane / LP, c4 = (lane % LP) * 4;
  __shared__ float s_red[kWarps][LP * 4][KM];
  __shared__ int s_meta[8];
  long long M = a.M;
  if constexpr (CMP) {
    M = __ldg(a.cmeta + 8);
    if (threadIdx.x < 8) s_meta[threadIdx.x] = __ldg(a.cmeta + threadIdx.x);
    __syncthreads();
  }

  float ca[4], cb[4], cc[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    ca[e] = a.coef_a[c4 + e];
    cb[e] = a.coef_b[c4 + e];
    cc[e] = a.coef_c[c4 + e];
  }
  float acc[4][KM];
#pragma unroll
  for (int e = 0; e < 4; ++e)
#pragma unroll
    for (int k = 0; k < KM; ++k) acc[e][k] = 0.f;

  const long long nchunks = M >> 5, cstride = (long long)gridDim.x * kWarps;
  long long chunk = (long long)blockIdx.x * kWarps + warp;
  float xn[KM];
  if (chunk < nchunks) gather_row<KM, CMP>(a, (chunk << 5) + lane, xn);
  for (; chunk < nchunks; chunk += cstride) {
    const long long pos0 = chunk << 5;
    float x[KM];
#pragma unroll
    for (int k = 0; k < KM; ++k) x[k] = xn[k];
    if (chunk + cstride < nchunks) gather_row<KM, CMP>(a, ((chunk + cstride) << 5) + lane, xn);
    mlp::TileClass tc{};
    int p64 = 0;
    if constexpr (CMP) {
      tc = mlp::tile_class(s_meta, pos0, a.NS);
      p64 = (int)(pos0 & 63);
    }
    const int4 *grow = reinterpret_cast<const int4 *>(a.gr + (size_t)pos0 * a.Cout + c4);
    const int4 *zrow = reinterpret_cast<const int4 *>(a.zin + (size_t)pos0 * a.Cout + c4);
    const size_t rstride = (size_t)a.Cout >> 2;   // row stride in int4
#pragma unroll 1
    for (int it0 = 0; it0 < ITERS; it0 += UN) {
      int4 g[UN], zz[UN];
#pragma unroll
      for (int u = 0; u < UN; ++u) {   // 2*UN independent 128-bit streams in flight per lane
        const size_t r = (size_t)((it0 + u) * PPW + sub) * rstride;
        g[u] = ldg_stream_v4(grow + r);
        zz[u] = ldg_stream_v4(zrow + r);
      }
#pragma unroll
      for (int u = 0; u < UN; ++u) {
        const int src = (it0 + u) * PPW + sub;
        float xr[KM];
#pragma unroll
        for (int k = 0; k < KM; ++k) xr[k] = __shfl_sync(0xffffffffu, x[k], src);
        const float gv[4] = {__int_as_float(g[u].x), __int_as_float(g[u].y),
                             __int_as_float(g[u].z), __int_as_float(g[u].w)};
        const float zv[4] = {__int_as_float(zz[u].x), __int_as_float(zz[u].y),
                             __int_as_float(zz[u].z), __int_as_float(zz[u].w)};
        // CMP: gr carries the position's multiplicity already (it came through the weighted dz
        // of the layers above); the dense BatchNorm-backward term is once per padded position
        float wgt = 1.f;
        if constexpr (CMP)
          wgt = src < tc.live ? ((((p64 + src) & (tc.ns - 1)) == 0) ? 1.f + tc.wx : 1.f) : 0.f;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float dz = CMP ? fmaf(ca[e], gv[e], wgt * fmaf(cb[e], zv[e], cc[e]))
                               : fmaf(ca[e], gv[e], fmaf(cb[e], zv[e], cc[e]));
#pragma unroll
          for (int k = 0; k < KM; ++k)
            if (k < a.K) acc[e][k] = fmaf(dz, xr[k], acc[e][k]);
        }
      }
    }
  }
#pragma unroll
  for (int e = 0; e < 4; ++e)
#pragma unroll
    for (int k = 0; k < KM; ++k) {
#pragma unroll
      for (int o = LP; o < 32; o <<= 1) acc[e][k] += __shfl_xor_sync(0xffffffffu, acc[e][k], o);
      if (sub == 0) s_red[warp][c4 + e][k] = acc[e][k];
    }
  __syncthreads();
  for (int i = threadIdx.x; i < LP * 4 * a.K; i += kThreads) {
    const int c = i / a.K, k = i % a.K;
    float t = 0.f;
#pragma unroll
    for (int wv = 0; wv < kWarps; ++wv) t += s_red[wv][c][k];
    atomicAdd(a.dW + (size_t)c * a.K + k, t);   // (Cout, Cin) nn.Conv2d layout, accumulated
  }
}

bool disabled() {   // read per call so that tests can compare the two paths in one process
  const char *v = getenv("B2R_NO_THIN");
  return v && *v && *v != '0';
}

bool shape_ok(int Cin, int Cout, long long M, long long per_scene) {
  if (disabled()) return false;
  if (Cin < 3 || Cin > kMaxK) return false;
  if (!(Cout == 32 || Cout == 64 || Cout == 128)) return false;
  return M > 0 && (per_scene % 32) == 0;
}

int grid_for(long long M, int sm_limit) {
  int sms = kNumSMs;
  if (sm_limit > 0 && sm_limit < kNumSMs) sms = sm_limit;
  const long long want = ((M >> 5) + kWarps - 1) / kWarps;
  const long long cap = (long long)sms * kCtasPerSM;
  return (int)(want < cap ? want : cap);
}

}  // namespace

bool fwd_applicable(const b2r_sa_layer *d) {
  return d->mode == 0 && d->epilogue == 0 &&
         shape_ok(d->Cin, d->Cout, (long long)d->B * d->NP * d->NS, (long long)d->NP * d->NS);
}

int fwd_launch(const b2r_sa_layer *d, void *stream) {
  ThinArgs a = {};
  a.N = d->N; a.NP = d->NP; a.NS = d->NS; a.C = d->Cin - 3; a.K = d->Cin; a.Cout = d->Cout;
  a.per_scene = (long long)d->NP * d->NS;
  a.M = a.per_scene * d->B;
  a.xyz = d->xyz; a.new_xyz = d->new_xyz; a.feat_t = d->feat_t; a.idx = d->idx;
  a.radius = d->radius; a.normalize_xyz = d->normalize_xyz;
  a.w_image = d->w_image;
  a.Cout_pad = (d->Cout + 127) & ~127;
  a.Cf4 = (a.C + 3) & ~3;
  a.z = d->z; a.stats = d->stats;
  a.cidx = d->cidx; a.ccen = d->ccen; a.cmeta = d->cmeta;
  const bool cmp = d->cmeta != nullptr;
  const int grid = grid_for(cmp ? b2r_compact_capacity(d->B, d->NP, d->NS) : a.M, d->sm_limit);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
#define B2R_THIN(LPV)                                                                  \
  do {                                                                               \
    if (cmp) {                                                                       \
      if (a.K <= 4) thin_fwd_kernel<LPV, 4, true><<<grid, kThreads, 0, st>>>(a);     \
      else thin_fwd_kernel<LPV, 8, true><<<grid, kThreads, 0, st>>>(a);              \
    } else {                                                                         \
      if (a.K <= 4) thin_fwd_kernel<LPV, 4, false><<<grid, kThreads, 0, st>>>(a);    \
      else thin_fwd_kernel<LPV, 8, false><<<grid, kThreads, 0, st>>>(a);             \
    }                                                                                \
  } while (0)
  switch (d->Cout) {
    case 32: B2R_THIN(8); break;
    case 64: B2R_THIN(16); break;
    default: B2R_THIN(32); break;
  }
#undef B2R_THIN
  B2R_CHECK_LAUNCH();
  return B2R_OK;
}

bool bwd_applicable(const b2r_sa_layer_bwd_desc *d) {
  const bool dgrad = (d->g_feat_t != nullptr && d->Cin > 3) || d->g_xyz != nullptr ||
                     d->g_new_xyz != nullptr;
  return d->mode == 0 && !dgrad && d->dz == nullptr && d->dysel == nullptr && d->gr && d->z &&
         d->coef_a && d->coef_b && d->coef_c &&
         shape_ok(d->Cin, d->Cout, (long long)d->B * d->NP * d->NS, (long long)d->NP * d->NS);
}

int bwd_launch(const b2r_sa_layer_bwd_desc *d, void *stream) {
  ThinArgs a = {};
  a.N = d->N; a.NP = d->NP; a.NS = d->NS; a.C = d->Cin - 3; a.K = d->Cin; a.Cout = d->Cout;
  a.per_scene = (long long)d->NP * d->NS;
  a.M = a.per_scene * d->B;
  a.xyz = d->xyz; a.new_xyz = d->new_xyz; a.feat_t = d->feat_t; a.idx = d->idx;
  a.radius = d->radius; a.normalize_xyz = d->normalize_xyz;
  a.gr = d->gr; a.zin = d->z;
  a.coef_a = d->coef_a; a.coef_b = d->coef_b; a.coef_c = d->coef_c;
  a.dW = d->dW;
  a.cidx = d->cidx; a.ccen = d->ccen; a.cmeta = d->cmeta;
  const bool cmp = d->cmeta != nullptr;
  const int grid = grid_for(cmp ? b2r_compact_capacity(d->B, d->NP, d->NS) : a.M, d->sm_limit);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
#define B2R_THIN(LPV)                                                                  \
  do {                                                                               \
    if (cmp) {                                                                       \
      if (a.K <= 4) thin_bwd_kernel<LPV, 4, true><<<grid, kThreads, 0, st>>>(a);     \
      else thin_bwd_kernel<LPV, 8, true><<<grid, kThreads, 0, st>>>(a);              \
    } else {                                                                         \
      if (a.K <= 4) thin_bwd_kernel<LPV, 4, false><<<grid, kThreads, 0, st>>>(a);    \
      else thin_bwd_kernel<LPV, 8, false><<<grid, kThreads, 0, st>>>(a);             \
    }                                                                                \
  } while (0)
  switch (d->Cout) {
    case 32: B2R_THIN(8); break;
    case 64: B2R_THIN(16); break;
    default: B2R_THIN(32); break;
  }
#undef B2R_THIN
  B2R_CHECK_LAUNCH();
  return B2R_OK;
}

}  // namespace thin
}  // namespace b2r
