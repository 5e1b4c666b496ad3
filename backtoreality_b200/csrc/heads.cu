// heads.cu -- the memory-bound glue around the wide dense layers of csrc/dense.cu, on POINT-major
// tensors (one warp per point, 128-bit coalesced rows):
//
//   interp_cat   PointnetFPModule's  three_interpolate -> torch.cat([interpolated, skip], dim=1)
//                (reference pointnet2_modules.py:492-504; kernel semantics interpolate_gpu.cu:77-106
//                with the same FMA contraction fma(p3,w3, fma(p1,w1, p2*w2))) in one pass, and its
//                backward (interpolate_gpu.cu:121-148: three weighted scatter-adds + the slice).
//   vote_tail    VotingModule's offset / residual split and VoteNet's feature normalisation
//                (models/voting_module.py:56-64, models/votenet.py:93-94):
//                vote_xyz = seed_xyz + net[:, 0:3];  v = seed_features + net[:, 3:];  out = v / |v|_2
//                and its backward.
#include "common.cuh"

namespace b2r {
namespace {

constexpr int kHWarps = 8;

// X0 (B*n, C2 + C1): [sum_k w_k known[b, idx_k, :], skip[b, j, :]]
__global__ void interp_cat_fwd_kernel(const float *__restrict__ known, const float *__restrict__ skip,
                                      const int *__restrict__ idx, const float *__restrict__ weight,
                                      int n, int m, int C2, int C1, long long rows,
                                      float *__restrict__ out) {
  const long long row = (long long)blockIdx.x * kHWarps + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const int b = (int)(row / n);
  const int i1 = idx[row * 3], i2 = idx[row * 3 + 1], i3 = idx[row * 3 + 2];
  const float w1 = weight[row * 3], w2 = weight[row * 3 + 1], w3 = weight[row * 3 + 2];
  const float *k1 = known + ((size_t)b * m + i1) * C2;
  const float *k2 = known + ((size_t)b * m + i2) * C2;
  const float *k3 = known + ((size_t)b * m + i3) * C2;
  float *o = out + (size_t)row * (C2 + C1);
  for (int c = lane * 4; c < C2; c += 128) {
    const float4 p1 = __ldg(reinterpret_cast<const float4 *>(k1 + c));
    const float4 p2 = __ldg(reinterpret_cast<const float4 *>(k2 + c));
    const float4 p3 = __ldg(reinterpret_cast<const float4 *>(k3 + c));
    float4 r;
    r.x = __fmaf_rn(p3.x, w3, __fmaf_rn(p1.x, w1, __fmul_rn(p2.x, w2)));
    r.y = __fmaf_rn(p3.y, w3, __fmaf_rn(p1.y, w1, __fmul_rn(p2.y, w2)));
    r.z = __fmaf_rn(p3.z, w3, __fmaf_rn(p1.z, w1, __fmul_rn(p2.z, w2)));
    r.w = __fmaf_rn(p3.w, w3, __fmaf_rn(p1.w, w1, __fmul_rn(p2.w, w2)));
    *reinterpret_cast<float4 *>(o + c) = r;
  }
  if (C1 > 0) {
    const float *s = skip + (size_t)row * C1;
    for (int c = lane * 4; c < C1; c += 128)
      *reinterpret_cast<float4 *>(o + C2 + c) = __ldg(reinterpret_cast<const float4 *>(s + c));
  }
}

// g (B*n, ld_g) -> g_known (B,m,C2) += w_k g[:, :C2] (zeroed by the caller), g_skip (B*n, C1) = g[:, C2:]
__global__ void interp_cat_bwd_kernel(const float *__restrict__ g, int ld_g,
                                      const int *__restrict__ idx, const float *__restrict__ weight,
                                      int n, int m, int C2, int C1, long long rows,
                                      float *__restrict__ g_known, float *__restrict__ g_skip) {
  const long long row = (long long)blockIdx.x * kHWarps + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const int b = (int)(row / n);
  const float *gr = g + (size_t)row * ld_g;
  if (g_known != nullptr) {
    const int i1 = idx[row * 3], i2 = idx[row * 3 + 1], i3 = idx[row * 3 + 2];
    const float w1 = weight[row * 3], w2 = weight[row * 3 + 1], w3 = weight[row * 3 + 2];
    float *k1 = g_known + ((size_t)b * m + i1) * C2;
    float *k2 = g_known + ((size_t)b * m + i2) * C2;
    float *k3 = g_known + ((size_t)b * m + i3) * C2;
    for (int c = lane; c < C2; c += 32) {   // a warp instruction adds 32 consecutive channels
      const float v = gr[c];
      atomicAdd(k1 + c, v * w1);
      atomicAdd(k2 + c, v * w2);
      atomicAdd(k3 + c, v * w3);
    }
  }
  if (g_skip != nullptr && C1 > 0) {
    float *s = g_skip + (size_t)row * C1;
    for (int c = lane * 4; c < C1; c += 128)
      *reinterpret_cast<float4 *>(s + c) = *reinterpret_cast<const float4 *>(gr + C2 + c);
  }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// net (M, ld_net): [offset(3), residual(C)] -> vote_xyz (M,3), out (M,C) = v / |v|, norm (M)
__global__ void vote_tail_fwd_kernel(const float *__restrict__ net, int ld_net,
                                     const float *__restrict__ seed_xyz,
                                     const float *__restrict__ seed_feat, int C, long long rows,
                                     float *__restrict__ vote_xyz, float *__restrict__ out,
                                     float *__restrict__ norm) {
  const long long row = (long long)blockIdx.x * kHWarps + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float *nr = net + (size_t)row * ld_net;
  const float *sf = seed_feat + (size_t)row * C;
  if (lane < 3) vote_xyz[row * 3 + lane] = seed_xyz[row * 3 + lane] + nr[lane];
  float ss = 0.f;
  for (int c = lane; c < C; c += 32) {
    const float v = sf[c] + nr[3 + c];
    ss = fmaf(v, v, ss);
  }
  ss = warp_sum(ss);
  const float nrm = sqrtf(ss);
  if (lane == 0) norm[row] = nrm;
  float *o = out + (size_t)row * C;
  for (int c = lane; c < C; c += 32) o[c] = __fdiv_rn(sf[c] + nr[3 + c], nrm);
}

// g_out (M,C) (gradient of out), g_vxyz (M,3) or NULL -> g_net (M, ld_net), g_seed_feat (M,C)
// (both = (g - out (g . out)) / norm on the feature part); g_net[:, 0:3] = g_vxyz
__global__ void vote_tail_bwd_kernel(const float *__restrict__ g_out, const float *__restrict__ g_vxyz,
                                     const float *__restrict__ out, const float *__restrict__ norm,
                                     int C, int ld_net, long long rows, float *__restrict__ g_net,
                                     float *__restrict__ g_seed_feat) {
  const long long row = (long long)blockIdx.x * kHWarps + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  float *gn = g_net + (size_t)row * ld_net;
  if (lane < 3) gn[lane] = g_vxyz != nullptr ? g_vxyz[row * 3 + lane] : 0.f;
  for (int c = 3 + C + lane; c < ld_net; c += 32) gn[c] = 0.f;
  const float *o = out + (size_t)row * C;
  float dot = 0.f;
  if (g_out != nullptr) {
    const float *go = g_out + (size_t)row * C;
    for (int c = lane; c < C; c += 32) dot = fmaf(go[c], o[c], dot);
  }
  dot = warp_sum(dot);
  const float inv = 1.f / norm[row];
  for (int c = lane; c < C; c += 32) {
    const float gv = g_out != nullptr ? (g_out[(size_t)row * C + c] - o[c] * dot) * inv : 0.f;
    gn[3 + c] = gv;
    if (g_seed_feat != nullptr) g_seed_feat[(size_t)row * C + c] = gv;
  }
}

}  // namespace
}  // namespace b2r

using namespace b2r;

extern "C" int b2r_interp_cat_fwd(const float *known, const float *skip, const int *idx,
                                  const float *weight, int B, int n, int m, int C2, int C1,
                                  float *out, void *stream) {
  B2R_REQUIRE(B >= 0 && n >= 0 && m > 0 && C2 > 0 && C1 >= 0, "b2r_interp_cat_fwd: bad size");
  B2R_REQUIRE((C2 % 4) == 0 && (C1 % 4) == 0, "b2r_interp_cat_fwd: channel counts must be multiples of 4");
  if (B == 0 || n == 0) return B2R_OK;
  B2R_REQUIRE(known && idx && weight && out && (skip || C1 == 0), "b2r_interp_cat_fwd: null pointer");
  const long long rows = (long long)B * n;
  interp_cat_fwd_kernel<<<ceil_div(rows, kHWarps), kHWarps * 32, 0, static_cast<cudaStream_t>(stream)>>>(
      known, skip, idx, weight, n, m, C2, C1, rows, out);
  B2R_CHECK_LAUNCH();
  return B2R_OK;
}

extern "C" int b2r_interp_cat_bwd(const float *g, int ld_g, const int *idx, const float *weight, int B,
                                  int n, int m, int C2, int C1, float *g_known, float *g_skip,
                                  void *stream) {
  B2R_REQUIRE(B >= 0 && n >= 0 && m > 0 && C2 > 0 && C1 >= 0 && ld_g >= C2 + C1,
              "b2r_interp_cat_bwd: bad size");
  B2R_REQUIRE((C2 % 4) == 0 && (C1 % 4) == 0 && (ld_g % 4) == 0,
              "b2r_interp_cat_bwd: channel counts / pitch must be multiples of 4");
  if (B == 0 || n == 0) return B2R_OK;
  B2R_REQUIRE(g && idx && weight, "b2r_interp_cat_bwd: null pointer");
  const long long rows = (long long)B * n;
  interp_cat_bwd_kernel<<<ceil_div(rows, kHWarps), kHWarps * 32, 0, static_cast<cudaStream_t>(stream)>>>(
      g, ld_g, idx, weight, n, m, C2, C1, rows, g_known, g_skip);
  B2R_CHECK_LAUNCH();
  return B2R_OK;
}

extern "C" int b2r_vote_tail_fwd(const float *net, int ld_net, const float *seed_xyz,
                                 const float *seed_feat, long long M, int C, float *vote_xyz,
                                 float *out, float *norm, void *stream) {
  B2R_REQUIRE(M >= 0 && C > 0 && ld_net >= 3 + C, "b2r_vote_tail_fwd: bad size");
  if (M == 0) return B2R_OK;
  B2R_REQUIRE(net && seed_xyz && seed_feat && vote_xyz && out && norm, "b2r_vote_tail_fwd: null pointer");
  vote_tail_fwd_kernel<<<ceil_div(M, kHWarps), kHWarps * 32, 0, static_cast<cudaStream_t>(stream)>>>(
      net, ld_net, seed_xyz, seed_feat, C, M, vote_xyz, out, norm);
  B2R_CHECK_LAUNCH();
  return B2R_OK;
}

extern "C" int b2r_vote_tail_bwd(const float *g_out, const float *g_vote_xyz, const float *out,
                                 const float *norm, long long M, int C, int ld_net, float *g_net,
                                 float *g_seed_feat, void *stream) {
  B2R_REQUIRE(M >= 0 && C > 0 && ld_net >= 3 + C, "b2r_vote_tail_bwd: bad size");
  if (M == 0) return B2R_OK;
  B2R_REQUIRE(out && norm && g_net, "b2r_vote_tail_bwd: null pointer");
  vote_tail_bwd_kernel<<<ceil_div(M, kHWarps), kHWarps * 32, 0, static_cast<cudaStream_t>(stream)>>>(
      g_out, g_vote_xyz, out, norm, C, ld_net, M, g_net, g_seed_feat);
  B2R_CHECK_LAUNCH();
  return B2R_OK;
}
