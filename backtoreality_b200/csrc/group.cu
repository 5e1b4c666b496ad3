// group.cu -- index movers: gather, grouping and the fused QueryAndGroup tail (sm_100a).
//
// Replaces gather_points(_grad)_kernel (reference _ext_src/src/sampling_gpu.cu:13-52) and
// group_points(_grad)_kernel (src/group_points_gpu.cu:13-69), which launch only B CTAs (one
// per scene) and walk the output with a per-thread serial loop over nsample.
//
// All of these are HBM/L2-bound byte movers, so the design rules are: fill the machine
// (grid over positions x channel chunks x scenes), coalesced 128-bit stores on the large
// contiguous side, index vectors loaded once (int4) and reused across a chunk of channels,
// random reads confined to one feature row (<= 4*N bytes, L1/L2 resident).
// Backward passes zero-fill the output with cudaMemsetAsync and accumulate with fire-and-forget
// RED.ADD.F32 (same unordered-sum contract as the reference's atomicAdd).
#include <stdlib.h>

#include "common.cuh"

namespace b2r {
// movers_staged.cu: source row staged in shared memory (true if it took the call)
bool group_fwd_staged(const float *f, const int *idx, int B, int C, int N, long long L, float *out,
                      cudaStream_t st, cudaError_t *err);
bool movers_legacy() {   // B2R_MOVERS_LEGACY=1: round 1's direct-from-global kernels (A/B timing)
  static const bool v = []() {
    const char *e = getenv("B2R_MOVERS_LEGACY");
    return e && *e && *e != '0';
  }();
  return v;
}
namespace {

constexpr int kThreads = 256;

// ---------------------------------------------------------------- gather ------------------
// out[b,c,j] = f[b,c,idx[b,j]]   grid (ceil(M/256), ceil(C/cpb), B)
__global__ void __launch_bounds__(kThreads)
    gather_fwd_kernel(const float *__restrict__ f, const int *__restrict__ idx, int C, int N, int M,
                      int cpb, float *__restrict__ out) {
  const int j = blockIdx.x * kThreads + threadIdx.x;
  if (j >= M) return;
  const int b = blockIdx.z;
  const int a = idx[(size_t)b * M + j];
  const int c0 = blockIdx.y * cpb, c1 = min(C, c0 + cpb);
  const float *src = f + ((size_t)b * C + c0) * N + a;
  float *dst = out + ((size_t)b * C + c0) * M + j;
  for (int c = c0; c < c1; ++c, src += N, dst += M) *dst = __ldg(src);
}

__global__ void __launch_bounds__(kThreads)
    gather_bwd_kernel(const float *__restrict__ g, const int *__restrict__ idx, int C, int N, int M,
                      int cpb, float *__restrict__ gf) {
  const int j = blockIdx.x * kThreads + threadIdx.x;
  if (j >= M) return;
  const int b = blockIdx.z;
  const int a = idx[(size_t)b * M + j];
  const int c0 = blockIdx.y * cpb, c1 = min(C, c0 + cpb);
  const float *src = g + ((size_t)b * C + c0) * M + j;
  float *dst = gf + ((size_t)b * C + c0) * N + a;
  for (int c = c0; c < c1; ++c, src += M, dst += N) atomicAdd(dst, *src);
}

// ---------------------------------------------------------------- group -------------------
// out[b,c,e] = f[b,c,idx[b,e]], e in [0, L = NP*NS).  VEC=4: each thread owns 4 consecutive e
// (int4 index load, float4 store); VEC=1 is the scalar fallback when L % 4 != 0.
template <int VEC>
__global__ void __launch_bounds__(kThreads)
    group_fwd_kernel(const float *__restrict__ f, const int *__restrict__ idx, int C, int N,
                     long long L, int cpb, float *__restrict__ out) {
  const long long e = ((long long)blockIdx.x * kThreads + threadIdx.x) * VEC;
  if (e >= L) return;
  const int b = blockIdx.z;
  const int c0 = blockIdx.y * cpb, c1 = min(C, c0 + cpb);
  const float *src = f + ((size_t)b * C + c0) * N;
  float *dst = out + ((size_t)b * C + c0) * L + e;
  if constexpr (VEC == 4) {
    const int4 a = *reinterpret_cast<const int4 *>(idx + (size_t)b * L + e);
#pragma unroll 4
    for (int c = c0; c < c1; ++c, src += N, dst += L) {
      float4 v;
      v.x = __ldg(src + a.x); v.y = __ldg(src + a.y);
      v.z = __ldg(src + a.z); v.w = __ldg(src + a.w);
      stg_stream_v4(reinterpret_cast<float4 *>(dst), v);
    }
  } else {
    const int a = idx[(size_t)b * L + e];
    for (int c = c0; c < c1; ++c, src += N, dst += L) *dst = __ldg(src + a);
  }
}

template <int VEC>
__global__ void __launch_bounds__(kThreads)
    group_bwd_kernel(const float *__restrict__ g, const int *__restrict__ idx, int C, int N,
                     long long L, int cpb, float *__restrict__ gf) {
  const long long e = ((long long)blockIdx.x * kThreads + threadIdx.x) * VEC;
  if (e >= L) return;
  const int b = blockIdx.z;
  const int c0 = blockIdx.y * cpb, c1 = min(C, c0 + cpb);
  const float *src = g + ((size_t)b * C + c0) * L + e;
  float *dst = gf + ((size_t)b * C + c0) * N;
  if constexpr (VEC == 4) {
    const int4 a = *reinterpret_cast<const int4 *>(idx + (size_t)b * L + e);
#pragma unroll 4
    for (int c = c0; c < c1; ++c, src += L, dst += N) {
      const int4 raw = ldg_stream_v4(reinterpret_cast<const int4 *>(src));
      atomicAdd(dst + a.x, __int_as_float(raw.x));
      atomicAdd(dst + a.y, __int_as_float(raw.y));
      atomicAdd(dst + a.z, __int_as_float(raw.z));
      atomicAdd(dst + a.w, __int_as_float(raw.w));
    }
  } else {
    const int a = idx[(size_t)b * L + e];
    for (int c = c0; c < c1; ++c, src += L, dst += N) atomicAdd(dst + a, *src);
  }
}

// ------------------------------------------------- fused QueryAndGroup tail ----------------
// blockIdx.y == 0 writes the 3 relative-xyz channels, blockIdx.y >= 1 a chunk of feature
// channels.  out (B, 3+C, NP, NS).
template <int VEC>
__global__ void __launch_bounds__(kThreads)
    query_group_fwd_kernel(const float *__restrict__ xyz, const float *__restrict__ new_xyz,
                           const float *__restrict__ f, const int *__restrict__ idx, int C, int N,
                           int NP, int NS, int cpb, float radius, int normalize,
                           float *__restrict__ out) {
  const long long L = (long long)NP * NS;
  const long long e = ((long long)blockIdx.x * kThreads + threadIdx.x) * VEC;
  if (e >= L) return;
  const int b = blockIdx.z;
  int a[VEC];
  if constexpr (VEC == 4) {
    const int4 t = *reinterpret_cast<const int4 *>(idx + (size_t)b * L + e);
    a[0] = t.x; a[1] = t.y; a[2] = t.z; a[3] = t.w;
  } else {
    a[0] = idx[(size_t)b * L + e];
  }
  float *obase = out + (size_t)b * (3 + C) * L + e;
  if (blockIdx.y == 0) {
    const float *p = xyz + (size_t)b * N * 3;
    const float *q = new_xyz + (size_t)b * NP * 3;
    float r[3][VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      const int j = (int)((e + v) / NS);  // centre of this slot (VEC slots may straddle centres)
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float d = __fsub_rn(__ldg(p + (size_t)a[v] * 3 + c), __ldg(q + (size_t)j * 3 + c));
        if (normalize) d = __fdiv_rn(d, radius);  // true division, as `grouped_xyz /= radius`
        r[c][v] = d;
      }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      if constexpr (VEC == 4)
        stg_stream_v4(reinterpret_cast<float4 *>(obase + c * L),
                      make_float4(r[c][0], r[c][1], r[c][2], r[c][3]));
      else
        obase[c * L] = r[c][0];
    }
  } else {
    const int c0 = (blockIdx.y - 1) * cpb, c1 = min(C, c0 + cpb);
    const float *src = f + ((size_t)b * C + c0) * N;
    float *dst = obase + (size_t)(3 + c0) * L;
#pragma unroll 4
    for (int c = c0; c < c1; ++c, src += N, dst += L) {
      if constexpr (VEC == 4) {
        float4 v;
        v.x = __ldg(src + a[0]); v.y = __ldg(src + a[1]);
        v.z = __ldg(src + a[2]); v.w = __ldg(src + a[3]);
        stg_stream_v4(reinterpret_cast<float4 *>(dst), v);
      } else {
        *dst = __ldg(src + a[0]);
      }
    }
  }
}

// Backward: blockIdx.y == 0 handles the xyz channels (scatter to grad_xyz (B,N,3), and the
// negated sum over the ball to grad_new_xyz (B,NP,3)); blockIdx.y >= 1 feature channel chunks.
__global__ void __launch_bounds__(kThreads)
    query_group_bwd_kernel(const float *__restrict__ g, const int *__restrict__ idx, int C, int N,
                           int NP, int NS, int cpb, float radius, int normalize,
                           float *__restrict__ gxyz, float *__restrict__ gnew,
                           float *__restrict__ gf) {
  const long long L = (long long)NP * NS;
  const long long e = (long long)blockIdx.x * kThreads + threadIdx.x;
  if (e >= L) return;
  const int b = blockIdx.z;
  const int a = idx[(size_t)b * L + e];
  const float *gbase = g + (size_t)b * (3 + C) * L + e;
  if (blockIdx.y == 0) {
    if (gxyz == nullptr && gnew == nullptr) return;
    const int j = (int)(e / NS);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float v = gbase[c * L];
      if (normalize) v = __fdiv_rn(v, radius);
      if (gxyz) atomicAdd(gxyz + ((size_t)b * N + a) * 3 + c, v);
      if (gnew) atomicAdd(gnew + ((size_t)b * NP + j) * 3 + c, -v);
    }
  } else {
    if (gf == nullptr) return;
    const int c0 = (blockIdx.y - 1) * cpb, c1 = min(C, c0 + cpb);
    const float *src = gbase + (size_t)(3 + c0) * L;
    float *dst = gf + ((size_t)b * C + c0) * N + a;
#pragma unroll 4
    for (int c = c0; c < c1; ++c, src += L, dst += N) atomicAdd(dst, *src);
  }
}

// channels per CTA: enough CTAs to fill 148 SMs a few times over, but long enough loops to
// amortise the index load
int pick_cpb(long long pos_blocks, int C, int B) {
  int cpb = 16;
  while (cpb > 1 && pos_blocks * ((C + cpb - 1) / cpb) * B < 4LL * kNumSMs) cpb >>= 1;
  return cpb;
}

bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace
}  // namespace b2r

using namespace b2r;

extern "C" int b2r_gather_fwd(const float *features, const int *idx, int B, int C, int N, int M,
                              float *out, void *stream) {
  B2R_REQUIRE(B >= 0 && C >= 0 && N >= 0 && M >= 0, "b2r_gather_fwd: negative size");
  if (B == 0 || C == 0 || M == 0) return B2R_OK;
  B2R_REQUIRE(features && idx && out, "b2r_gather_fwd: null pointer");
  B2R_REQUIRE(B <= 65535, "b2r_gather_fwd: B too large");
  const int xb = ceil_div(M, kThreads);
  const int cpb = pick_cpb(xb, C, B);
  dim3 grid(xb, ceil_div(C, cpb), B);
  gather_fwd_kernel<<<grid, kThreads, 0, static_cast<cudaStream_t>(stream)>>>(features, idx, C, N,
                                                                               M, cpb, out);
  B2R_CHECK_LAUNCH();
  return B2R_OK;
}

extern "C" int b2r_gather_bwd(const float *grad_out, const int *idx, int B, int C, int N, int M,
                              float *grad_features, void *stream) {
  B2R_REQUIRE(B >= 0 && C >= 0 && N >= 0 && M >= 0, "b2r_gather_bwd: negative size");
  if (B == 0 || C == 0 || N == 0) return B2R_OK;
  B2R_REQUIRE(grad_features, "b2r_gather_bwd: null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  B2R_CUDA(cudaMemsetAsync(grad_features, 0, sizeof(float) * (size_t)B * C * N, st));
  if (M == 0) return B2R_OK;
  B2R_REQUIRE(grad_out && idx, "b2r_gather_bwd: null pointer");
  B2R_REQUIRE(B <= 65535, "b2r_gather_bwd: B too large");
  const int xb = ceil_div(M, kThreads);
  const int cpb = pick_cpb(xb, C, B);
  dim3 grid(xb, ceil_div(C, cpb), B);
  gather_bwd_kernel<<<grid, kThreads, 0, st>>>(grad_out, idx, C, N, M, cpb, grad_features);
  B2R_CHECK_LAUNCH();
  return B2R_OK;
}

extern "C" int b2r_group_fwd(const float *features, const int *idx, int B, int C, int N, int NP,
                             int NS, float *out, void *stream) {
  B2R_REQUIRE(B >= 0 && C >= 0 && N >= 0 && NP >= 0 && NS >= 0, "b2r_group_fwd: negative size");
  const long long L = (long long)NP * NS;
  if (B == 0 || C == 0 || L == 0) return B2R_OK;
  B2R_REQUIRE(features && idx && out, "b2r_group_fwd: null pointer");
  B2R_REQUIRE(B <= 65535, "b2r_group_fwd: B too large");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (!movers_legacy()) {
    cudaError_t e = cudaSuccess;
    if (group_fwd_staged(features, idx, B, C, N, L, out, st, &e)) {
      B2R_CUDA(e);
      return B2R_OK;
    }
  }
  const bool vec = (L % 4 == 0) && aligned16(idx) && aligned16(out);
  const int xb = ceil_div(L, (long long)kThreads * (vec ? 4 : 1));
  const int cpb = pick_cpb(xb, C, B);
  dim3 grid(xb, ceil_div(C, cpb), B);
  if (vec)
    group_fwd_kernel<4><<<grid, kThreads, 0, st>>>(features, idx, C, N, L, cpb, out);
  else
    group_fwd_kernel<1><<<grid, kThreads, 0, st>>>(features, idx, C, N, L, cpb, out);
  B2R_CHECK_LAUNCH();
  return B2R_OK;
}

extern "C" int b2r_group_bwd(const float *grad_out, const int *idx, int B, int C, int N, int NP,
                             int NS, float *grad_features, void *stream) {
  B2R_REQUIRE(B >= 0 && C >= 0 && N >= 0 && NP >= 0 && NS >= 0, "b2r_group_bwd: negative size");
  const long long L = (long long)NP * NS;
  if (B == 0 || C == 0 || N == 0) return B2R_OK;
  B2R_REQUIRE(grad_features, "b2r_group_bwd: null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  B2R_CUDA(cudaMemsetAsync(grad_features, 0, sizeof(float) * (size_t)B * C * N, st));
  if (L == 0) return B2R_OK;
  B2R_REQUIRE(grad_out && idx, "b2r_group_bwd: null pointer");
  B2R_REQUIRE(B <= 65535, "b2r_group_bwd: B too large");
  const bool vec = (L % 4 == 0) && aligned16(idx) && aligned16(grad_out);
  const int xb = ceil_div(L, (long long)kThreads * (vec ? 4 : 1));
  const int cpb = pick_cpb(xb, C, B);
  dim3 grid(xb, ceil_div(C, cpb), B);
  if (vec)
    group_bwd_kernel<4><<<grid, kThreads, 0, st>>>(grad_out, idx, C, N, L, cpb, grad_features);
  else
    group_bwd_kernel<1><<<grid, kThreads, 0, st>>>(grad_out, idx, C, N, L, cpb, grad_features);
  B2R_CHECK_LAUNCH();
  return B2R_OK;
}

extern "C" int b2r_query_group_fwd(const float *xyz, const float *new_xyz, const float *features,
                                   const int *idx, int B, int C, int N, int NP, int NS,
                                   float radius, int normalize_xyz, float *out, void *stream) {
  B2R_REQUIRE(B >= 0 && C >= 0 && N >= 0 && NP >= 0 && NS >= 0,
              "b2r_query_group_fwd: negative size");
  const long long L = (long long)NP * NS;
  if (B == 0 || L == 0) return B2R_OK;
  B2R_REQUIRE(xyz && new_xyz && idx && out && (features || C == 0),
              "b2r_query_group_fwd: null pointer");
  B2R_REQUIRE(B <= 65535, "b2r_query_group_fwd: B too large");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool vec = (L % 4 == 0) && aligned16(idx) && aligned16(out);
  const int xb = ceil_div(L, (long long)kThreads * (vec ? 4 : 1));
  const int cpb = C > 0 ? pick_cpb(xb, C, B) : 1;
  dim3 grid(xb, 1 + (C > 0 ? ceil_div(C, cpb) : 0), B);
  if (vec)
    query_group_fwd_kernel<4><<<grid, kThreads, 0, st>>>(xyz, new_xyz, features, idx, C, N, NP, NS,
                                                         cpb, radius, normalize_xyz, out);
  else
    query_group_fwd_kernel<1><<<grid, kThreads, 0, st>>>(xyz, new_xyz, features, idx, C, N, NP, NS,
                                                         cpb, radius, normalize_xyz, out);
  B2R_CHECK_LAUNCH();
  return B2R_OK;
}

extern "C" int b2r_query_group_bwd(const float *grad_out, const int *idx, int B, int C, int N,
                                   int NP, int NS, float radius, int normalize_xyz,
                                   float *grad_xyz, float *grad_new_xyz, float *grad_features,
                                   void *stream) {
  B2R_REQUIRE(B >= 0 && C >= 0 && N >= 0 && NP >= 0 && NS >= 0,
              "b2r_query_group_bwd: negative size");
  const long long L = (long long)NP * NS;
  if (B == 0) return B2R_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (grad_xyz && N > 0)
    B2R_CUDA(cudaMemsetAsync(grad_xyz, 0, sizeof(float) * (size_t)B * N * 3, st));
  if (grad_new_xyz && NP > 0)
    B2R_CUDA(cudaMemsetAsync(grad_new_xyz, 0, sizeof(float) * (size_t)B * NP * 3, st));
  if (grad_features && C > 0 && N > 0)
    B2R_CUDA(cudaMemsetAsync(grad_features, 0, sizeof(float) * (size_t)B * C * N, st));
  if (L == 0) return B2R_OK;
  B2R_REQUIRE(grad_out && idx, "b2r_query_group_bwd: null pointer");
  B2R_REQUIRE(B <= 65535, "b2r_query_group_bwd: B too large");
  const int xb = ceil_div(L, kThreads);
  const bool want_f = grad_features && C > 0;
  const int cpb = want_f ? pick_cpb(xb, C, B) : 1;
  dim3 grid(xb, 1 + (want_f ? ceil_div(C, cpb) : 0), B);
  query_group_bwd_kernel<<<grid, kThreads, 0, st>>>(grad_out, idx, C, N, NP, NS, cpb, radius,
                                                    normalize_xyz, grad_xyz, grad_new_xyz,
                                                    want_f ? grad_features : nullptr);
  B2R_CHECK_LAUNCH();
  return B2R_OK;
}
