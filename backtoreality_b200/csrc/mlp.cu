// mlp.cu -- the SharedMLP block of a set-abstraction layer on tcgen05 tensor cores (sm_100a).
//
// Replaces, for PointnetSAModuleVotes (reference pointnet2_modules.py:245-267), the chain
//   QueryAndGroup (group xyz, -= centre, /= radius, group features, cat)      5 HBM passes
//   3 x [cuDNN 1x1 conv -> BatchNorm2d -> ReLU]                                 ~9 HBM passes
//   F.max_pool2d over nsample                                                   1 HBM pass
// which materialises every (B, C, npoint, nsample) intermediate.
//
// One kernel, `sa_layer_fwd_kernel`, runs ONE conv layer as a fused GEMM
//        D[Cout x positions] = W[Cout x K] * X[positions x K]^T          (TF32 in, FP32 accumulate)
// with the neighbouring memory-bound work folded into its prologue / epilogue:
//   prologue  mode 0: X rows are GATHERED on the fly (ball-query idx -> point-major features +
//                     relative, radius-normalised xyz): the grouped tensor never exists in HBM.
//             mode 1: X rows are the previous layer's raw conv output with that layer's
//                     BatchNorm scale/shift + ReLU applied while loading.
//   epilogue  0: per-channel sum / sum-of-squares (BatchNorm batch statistics) + store raw z.
//             1: statistics + max AND min over the `nsample` positions of every centre (+ arg
//                indices).  BN+ReLU are monotone per channel, so max-pool(relu(bn(z))) is
//                relu(bn(max z)) for a non-negative BN scale and relu(bn(min z)) otherwise:
//                the (B,C,npoint,nsample) activation of the last layer is never stored.
// Orientation "channels on TMEM lanes": the accumulator row (TMEM lane) is an output channel and
// the column a position, so BN statistics, BN scale/shift and the max-pool over `nsample` are all
// per-THREAD register work in the epilogue (one thread owns one channel), no shuffles.
//
// Blackwell specifics: tcgen05.mma (kind::tf32, cta_group::1, M=128, N=64/128, K=8) issued by one
// elected thread, operands in 128B-swizzled K-major shared memory, accumulators in TMEM read back
// with tcgen05.ld.32x32b.x32, MMA completion signalled through tcgen05.commit -> mbarrier.
// The weight operand is staged by the TMA unit: one cp.async.bulk (UBLKCP) of a pre-swizzled
// weight image built by `pack_weight_kernel`.
#include "mlp_common.cuh"

namespace b2r {
using namespace mlp;
namespace {

// W (Cout, Cin) row-major fp32 (a 1x1 conv weight) -> TF32-rounded shared-memory image of the A
// operand: Cout padded to 128 rows per M tile, K padded to 32-element atoms, SW128 K-major.
__global__ void pack_weight_kernel(const float *__restrict__ w, int Cout, int Cin, int gather,
                                   int Kp, int Cout_pad, float *__restrict__ image) {
  const int KA = (Kp + 31) >> 5;
  const long long total = (long long)Cout_pad * KA * 32;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const int m = (int)(e / (KA * 32)), kp = (int)(e % (KA * 32));
    float v = 0.f;
    if (m < Cout && kp < Kp) {
      int k = -1;  // source column in the reference order [dx,dy,dz, f0..fC-1]
      if (gather) {
        const int C = Cin - 3, Cf4 = (C + 3) & ~3;
        if (kp < C) k = 3 + kp;
        else if (kp >= Cf4 && kp < Cf4 + 3) k = kp - Cf4;
      } else if (kp < Cin) {
        k = kp;
      }
      if (k >= 0) v = __uint_as_float(to_tf32(w[(size_t)m * Cin + k]));
    }
    const uint32_t off = sw128_off(m, kp >> 2, Cout_pad) + (kp & 3) * 4;
    image[off >> 2] = v;
  }
}

struct LayerArgs {
  int B, N, NP, NS, Cin, Cout, mode, epilogue;
  const float *xyz, *new_xyz, *feat_t;
  const int *idx;
  float radius;
  int normalize_xyz;
  const float *z_prev, *scale_prev, *shift_prev;
  const float *w_image;
  float *z;
  double *stats;
  float *zmax, *zmin;
  int *amax, *amin;
  // epilogue 2 (backward helper): dz of the pooled top layer
  const float *dysel;
  const int *asel;
  const float *bw_k1, *bw_k2, *bw_mean, *bw_invstd, *bw_gs;
  float *dz;
  int Kp, Cout_pad, num_tiles, chf_shift;
};

// max/min + arg over groups of NS columns held in registers; writes (centre, channel) entries
template <int NS>
__device__ __forceinline__ void pool_groups(const float (&v)[64], long long pos0, int c, int Cout,
                                            float *__restrict__ zmax, float *__restrict__ zmin,
                                            int *__restrict__ amax, int *__restrict__ amin) {
#pragma unroll
  for (int g = 0; g < 64 / NS; ++g) {
    float mx = v[g * NS], mn = v[g * NS];
    int ax = 0, an = 0;
#pragma unroll
    for (int s = 1; s < NS; ++s) {
      const float x = v[g * NS + s];
      if (x > mx) { mx = x; ax = s; }   // strict: the first maximum wins, like max_pool2d
      if (x < mn) { mn = x; an = s; }
    }
    const long long centre = (pos0 + g * NS) / NS;
    const size_t o = (size_t)centre * Cout + c;
    zmax[o] = mx; zmin[o] = mn; amax[o] = ax; amin[o] = an;
  }
}

// Backward helper (epilogue 2): with z of the pooled top layer back in registers, emit
//   dz = gamma*invstd * (dy - mean(dy) - xhat * mean(dy*xhat)),   xhat = (z - mean) * invstd
// where dy is the max-pool / ReLU routed output gradient: non-zero only at the selected sample.
template <int NS>
__device__ __forceinline__ void dz_groups(const float (&v)[64], long long pos0, int c, int Cout,
                                          const float *__restrict__ dysel,
                                          const int *__restrict__ asel, float k1, float k2,
                                          float mean, float invstd, float gs,
                                          float *__restrict__ dz) {
#pragma unroll
  for (int g = 0; g < 64 / NS; ++g) {
    const long long centre = (pos0 + g * NS) / NS;
    const float dyv = dysel[(size_t)centre * Cout + c];
    const int as = asel[(size_t)centre * Cout + c];
    float *o = dz + (size_t)(pos0 + g * NS) * Cout + c;
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      const float dy = (s == as) ? dyv : 0.f;
      const float xh = (v[g * NS + s] - mean) * invstd;
      o[(size_t)s * Cout] = gs * (dy - k1 - xh * k2);
    }
  }
}

// NT = positions per tile (MMA N).  MT = Cout_pad / 128 M-tiles.
//   MT == 1: warps 0-3 take columns [0,NT/2), warps 4-7 columns [NT/2,NT) of the single M tile
//   MT == 2: warps 0-3 take M tile 0, warps 4-7 M tile 1, all NT columns (NT must be 64)
template <int NT>
__global__ void __launch_bounds__(kMlpThreads, 1) sa_layer_fwd_kernel(const LayerArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve-up (all 1024-byte aligned where the swizzle needs it)
  const int KA = (a.Kp + 31) >> 5;
  const uint32_t w_bytes = (uint32_t)a.Cout_pad * KA * 128;
  const uint32_t x_bytes = (uint32_t)NT * KA * 128;
  uint8_t *base = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                              ~(uintptr_t)1023);
  uint8_t *s_w = base;
  uint8_t *s_x = s_w + w_bytes;
  float *s_scale = reinterpret_cast<float *>(s_x + x_bytes);
  float *s_shift = s_scale + a.Kp;
  int *s_idx = reinterpret_cast<int *>(s_shift + a.Kp);
  uint64_t *s_bar = reinterpret_cast<uint64_t *>(
      (reinterpret_cast<uintptr_t>(s_idx + NT) + 15) & ~(uintptr_t)15);  // [0]: weights, [1]: MMA
  uint32_t *s_tmem = reinterpret_cast<uint32_t *>(s_bar + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int MT = a.Cout_pad >> 7;
  const uint32_t bar_w = smem_u32(&s_bar[0]), bar_mma = smem_u32(&s_bar[1]);
  constexpr uint32_t kTmemCols = 128;

  // ---- one-time setup ----------------------------------------------------------------------
  if (tid == 0) {
    mbar_init(bar_w, 1);
    mbar_init(bar_mma, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(smem_u32(s_tmem), kTmemCols);
  // zero the X tile once: padding chunks are never written again
  for (uint32_t i = tid * 16; i < x_bytes; i += kMlpThreads * 16)
    *reinterpret_cast<uint4 *>(s_x + i) = make_uint4(0, 0, 0, 0);
  if (a.mode == 1)
    for (int i = tid; i < a.Cin; i += kMlpThreads) {
      s_scale[i] = a.scale_prev[i];
      s_shift[i] = a.shift_prev[i];
    }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;
  if (tid == 0) {  // TMA unit stages the whole pre-swizzled weight image
    mbar_expect_tx(bar_w, w_bytes);
    bulk_g2s(smem_u32(s_w), a.w_image, w_bytes, bar_w);
  }

  // epilogue role of this thread
  const int q = warp & 3, h = warp >> 2;
  const int mt = (MT == 2) ? h : 0;
  const int col0 = (MT == 2) ? 0 : h * (NT / 2);
  constexpr int kColsMax = (NT == 128) ? 64 : 64;
  const int ncols = (MT == 2) ? NT : NT / 2;  // 64, or 32 when (MT == 1, NT == 64)
  const int c = mt * 128 + q * 32 + lane;     // output channel owned by this thread
  const bool c_ok = c < a.Cout;
  double acc_s = 0.0, acc_ss = 0.0;
  (void)kColsMax;
  float e2_k1 = 0.f, e2_k2 = 0.f, e2_mean = 0.f, e2_invstd = 0.f, e2_gs = 0.f;
  if (a.epilogue == 2 && c_ok) {
    e2_k1 = a.bw_k1[c]; e2_k2 = a.bw_k2[c]; e2_mean = a.bw_mean[c];
    e2_invstd = a.bw_invstd[c]; e2_gs = a.bw_gs[c];
  }

  const int KS = (a.Kp + 7) >> 3;  // K = 8 slices actually issued
  const uint32_t idesc = idesc_tf32(NT);
  const long long per_scene = (long long)a.NP * a.NS;
  GatherSrc gsrc;
  gsrc.xyz = a.xyz; gsrc.new_xyz = a.new_xyz; gsrc.feat_t = a.feat_t;
  gsrc.N = a.N; gsrc.NP = a.NP; gsrc.NS = a.NS; gsrc.C = a.Cin - 3; gsrc.Cf4 = (a.Cin - 3 + 3) & ~3;
  gsrc.chf_shift = a.chf_shift; gsrc.radius = a.radius; gsrc.normalize_xyz = a.normalize_xyz;
  uint32_t mma_parity = 0;
  bool w_ready = false;

  for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
    const long long pos0 = (long long)tile * NT;

    // ---- prologue: build the X tile (B operand) in swizzled shared memory ------------------
    if (a.mode == 0) {
      if (tid < NT) s_idx[tid] = a.idx[pos0 + tid];
      __syncthreads();
      const int b = (int)(pos0 / per_scene);
      const int in_scene0 = (int)(pos0 - (long long)b * per_scene);
      build_x_gather<NT>(gsrc, b, in_scene0, s_idx, s_x, tid,
                         [](int row, int ch) { return sw128_off(row, ch, NT); });
    } else {
      // dense layer: the tile is one contiguous block of z_prev.  Loads are issued in batches of
      // 8 independent 16-byte requests per thread (memory-level parallelism), and the NEXT tile
      // of this CTA is pulled into L2 by a single bulk-prefetch instruction meanwhile.
      const int CH = a.Cin >> 2;
      const int total = NT * CH;
      if (tid == 0 && tile + (int)gridDim.x < a.num_tiles)
        prefetch_l2(a.z_prev + (size_t)(pos0 + (long long)gridDim.x * NT) * a.Cin,
                    (uint32_t)total * 16u);
      const float4 *src = reinterpret_cast<const float4 *>(a.z_prev + (size_t)pos0 * a.Cin);
      for (int i0 = tid; i0 < total; i0 += kMlpThreads * 8) {
        float4 t[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int i = i0 + u * kMlpThreads;
          if (i < total) t[u] = __ldg(src + i);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int i = i0 + u * kMlpThreads;
          if (i < total) {
            const int row = i / CH, ch = i - row * CH;
            const float4 sc = *reinterpret_cast<const float4 *>(s_scale + ch * 4);
            const float4 sh = *reinterpret_cast<const float4 *>(s_shift + ch * 4);
            uint4 out;
            out.x = to_tf32(fmaxf(fmaf(t[u].x, sc.x, sh.x), 0.f));
            out.y = to_tf32(fmaxf(fmaf(t[u].y, sc.y, sh.y), 0.f));
            out.z = to_tf32(fmaxf(fmaf(t[u].z, sc.z, sh.z), 0.f));
            out.w = to_tf32(fmaxf(fmaf(t[u].w, sc.w, sh.w), 0.f));
            *reinterpret_cast<uint4 *>(s_x + sw128_off(row, ch, NT)) = out;
          }
        }
      }
    }
    fence_async_smem();   // generic-proxy writes -> visible to the tensor core (async proxy)
    tc_fence_before();
    __syncthreads();

    // ---- MMA: one elected thread issues MT x KS tcgen05.mma, then commits to the mbarrier --
    if (tid == 0) {
      if (!w_ready) {
        mbar_wait(bar_w, 0);
        w_ready = true;
      }
      tc_fence_after();
      const uint32_t xa = smem_u32(s_x), wa = smem_u32(s_w);
      for (int m = 0; m < MT; ++m) {
        for (int ks = 0; ks < KS; ++ks) {
          const uint32_t koff = (uint32_t)(ks >> 2) * 1024u;  // K atom index (x rows/8 below)
          const uint64_t da = smem_desc_sw128(wa + koff * (uint32_t)(a.Cout_pad >> 3) +
                                              (uint32_t)m * 16u * 1024u + (uint32_t)(ks & 3) * 32u);
          const uint64_t db =
              smem_desc_sw128(xa + koff * (uint32_t)(NT >> 3) + (uint32_t)(ks & 3) * 32u);
          umma_tf32(tmem_base + (uint32_t)m * NT, da, db, idesc, ks > 0 ? 1u : 0u);
        }
      }
      umma_commit(bar_mma);
    }

    // ---- epilogue: TMEM -> registers; statistics, store / pool ------------------------------
    mbar_wait(bar_mma, mma_parity);
    mma_parity ^= 1u;
    tc_fence_after();
    float v[64];
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(mt * NT + col0);
    {
      uint32_t r[32];
      cuda::ptx::tcgen05_ld_32x32b(r, taddr);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
    }
    if (ncols == 64) {
      uint32_t r[32];
      cuda::ptx::tcgen05_ld_32x32b(r, taddr + 32);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int i = 0; i < 32; ++i) v[32 + i] = __uint_as_float(r[i]);
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i) v[32 + i] = 0.f;
    }
    tc_fence_before();  // TMEM reads done before the next tile's MMAs (ordered by __syncthreads)

    if (c_ok) {
      float ts = 0.f, tss = 0.f;
#pragma unroll
      for (int i = 0; i < 64; ++i) {
        ts += v[i];
        tss = fmaf(v[i], v[i], tss);
      }
      acc_s += (double)ts;
      acc_ss += (double)tss;
      if (a.epilogue == 0) {
        float *zp = a.z + (size_t)(pos0 + col0) * a.Cout + c;
#pragma unroll
        for (int i = 0; i < 64; ++i)
          if (i < ncols) zp[(size_t)i * a.Cout] = v[i];  // a warp writes 32 channels = 128 B
      } else if (a.epilogue == 1) {
        const long long p0 = pos0 + col0;
        if (a.NS == 16) pool_groups<16>(v, p0, c, a.Cout, a.zmax, a.zmin, a.amax, a.amin);
        else if (a.NS == 32) pool_groups<32>(v, p0, c, a.Cout, a.zmax, a.zmin, a.amax, a.amin);
        else pool_groups<64>(v, p0, c, a.Cout, a.zmax, a.zmin, a.amax, a.amin);
      } else {
        const long long p0 = pos0 + col0;
        if (a.NS == 16)
          dz_groups<16>(v, p0, c, a.Cout, a.dysel, a.asel, e2_k1, e2_k2, e2_mean, e2_invstd, e2_gs, a.dz);
        else if (a.NS == 32)
          dz_groups<32>(v, p0, c, a.Cout, a.dysel, a.asel, e2_k1, e2_k2, e2_mean, e2_invstd, e2_gs, a.dz);
        else
          dz_groups<64>(v, p0, c, a.Cout, a.dysel, a.asel, e2_k1, e2_k2, e2_mean, e2_invstd, e2_gs, a.dz);
      }
    }
    // the next iteration's prologue __syncthreads orders these TMEM loads before its MMAs and
    // this tile's MMA smem reads (complete: we waited on the commit) before the X overwrite
  }

  if (c_ok && a.stats != nullptr) {
    atomicAdd(a.stats + c, acc_s);
    atomicAdd(a.stats + a.Cout + c, acc_ss);
  }
  if (tid == 0 && !w_ready) mbar_wait(bar_w, 0);  // never leave a bulk copy in flight
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, kTmemCols);
}

// BatchNorm bookkeeping from the accumulated statistics (one thread per channel):
//   scale = gamma / sqrt(var_biased + eps), shift = beta - mean * scale        (training)
//   running_mean/var updated with `momentum` (unbiased variance), like nn.BatchNorm2d
__global__ void bn_finalize_kernel(const double *__restrict__ stats, int Cch, double count,
                                   const float *__restrict__ gamma, const float *__restrict__ beta,
                                   float eps, float momentum, float *__restrict__ running_mean,
                                   float *__restrict__ running_var, float *__restrict__ scale,
                                   float *__restrict__ shift, float *__restrict__ mean_out,
                                   float *__restrict__ invstd_out,
                                   long long *__restrict__ num_batches_tracked) {
  const int ch = blockIdx.x * blockDim.x + threadIdx.x;
  if (ch == 0 && num_batches_tracked != nullptr) *num_batches_tracked += 1;
  if (ch >= Cch) return;
  const double mean = stats[ch] / count;
  double var = stats[Cch + ch] / count - mean * mean;
  if (var < 0.0) var = 0.0;
  const float invstd = (float)(1.0 / sqrt(var + (double)eps));
  const float g = gamma ? gamma[ch] : 1.f, bta = beta ? beta[ch] : 0.f;
  scale[ch] = g * invstd;
  shift[ch] = bta - (float)mean * g * invstd;
  if (mean_out) mean_out[ch] = (float)mean;
  if (invstd_out) invstd_out[ch] = invstd;
  if (running_mean) {
    const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
    running_mean[ch] = (1.f - momentum) * running_mean[ch] + momentum * (float)mean;
    running_var[ch] = (1.f - momentum) * running_var[ch] + momentum * (float)unbiased;
  }
}

// out = relu(scale * (scale >= 0 ? zmax : zmin) + shift), written channel-major (B,C,NP) for the
// reference API and point-major (B,NP,C) for the next layer's gather.  One thread per (centre, c).
__global__ void pool_finalize_kernel(const float *__restrict__ zmax, const float *__restrict__ zmin,
                                     const float *__restrict__ scale,
                                     const float *__restrict__ shift, int B, int NP, int Cch,
                                     float *__restrict__ out_cm, float *__restrict__ out_pm) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)B * NP * Cch;
  if (e >= total) return;
  const int ch = (int)(e % Cch);
  const long long centre = e / Cch;
  const float s = scale[ch];
  const float z = s >= 0.f ? zmax[e] : zmin[e];
  const float y = fmaxf(fmaf(z, s, shift[ch]), 0.f);
  if (out_pm) out_pm[e] = y;
  if (out_cm) {
    const int b = (int)(centre / NP), j = (int)(centre % NP);
    out_cm[((size_t)b * Cch + ch) * NP + j] = y;
  }
}

// (B,C,N) channel-major -> (B,N,C) point-major (tile transpose through shared memory)
__global__ void to_point_major_kernel(const float *__restrict__ in, int Cch, int N,
                                      float *__restrict__ out) {
  __shared__ float t[32][33];
  const int b = blockIdx.z;
  const int n0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  in += (size_t)b * Cch * N;
  out += (size_t)b * Cch * N;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int cc = c0 + i, n = n0 + threadIdx.x;
    t[i][threadIdx.x] = (cc < Cch && n < N) ? in[(size_t)cc * N + n] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int n = n0 + i, cc = c0 + threadIdx.x;
    if (n < N && cc < Cch) out[(size_t)n * Cch + cc] = t[threadIdx.x][i];
  }
}

size_t layer_smem_bytes(int Kp, int Cout_pad, int NT) {
  const int KA = (Kp + 31) >> 5;
  return 1024 + (size_t)Cout_pad * KA * 128 + (size_t)NT * KA * 128 + 2 * (size_t)Kp * 4 +
         (size_t)NT * 4 + 16 + 2 * 8 + 16;
}

}  // namespace
}  // namespace b2r

using namespace b2r;

extern "C" long long b2r_mlp_weight_image_bytes(int Cout, int Cin, int gather) {
  if (Cout <= 0 || Cin <= 0) return 0;
  const int Kp = packed_k(Cin, gather);
  const int KA = (Kp + 31) >> 5;
  const int Cout_pad = (Cout + 127) & ~127;
  return (long long)Cout_pad * KA * 128;
}

extern "C" int b2r_mlp_pack_weight(const float *w, int Cout, int Cin, int gather, float *image,
                                   void *stream) {
  B2R_REQUIRE(w && image && Cout > 0 && Cin > 0, "b2r_mlp_pack_weight: bad argument");
  B2R_REQUIRE(!gather || Cin >= 3, "b2r_mlp_pack_weight: gather layers need Cin >= 3");
  const int Kp = packed_k(Cin, gather);
  const int KA = (Kp + 31) >> 5;
  const int Cout_pad = (Cout + 127) & ~127;
  const long long total = (long long)Cout_pad * KA * 32;
  pack_weight_kernel<<<ceil_div(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      w, Cout, Cin, gather, Kp, Cout_pad, image);
  B2R_CHECK_LAUNCH();
  return B2R_OK;
}

extern "C" int b2r_sa_layer_fwd(const b2r_sa_layer *d, void *stream) {
  B2R_REQUIRE(d != nullptr, "b2r_sa_layer_fwd: null descriptor");
  B2R_REQUIRE(d->B > 0 && d->NP > 0 && d->NS > 0 && d->Cin > 0 && d->Cout > 0,
              "b2r_sa_layer_fwd: non-positive size");
  B2R_REQUIRE(d->mode == 0 || d->mode == 1, "b2r_sa_layer_fwd: mode must be 0 or 1");
  B2R_REQUIRE(d->epilogue >= 0 && d->epilogue <= 2, "b2r_sa_layer_fwd: epilogue must be 0, 1 or 2");
  B2R_REQUIRE(d->w_image != nullptr, "b2r_sa_layer_fwd: null weight image");
  LayerArgs a;
  a.B = d->B; a.N = d->N; a.NP = d->NP; a.NS = d->NS; a.Cin = d->Cin; a.Cout = d->Cout;
  a.mode = d->mode; a.epilogue = d->epilogue;
  a.xyz = d->xyz; a.new_xyz = d->new_xyz; a.feat_t = d->feat_t; a.idx = d->idx;
  a.radius = d->radius; a.normalize_xyz = d->normalize_xyz;
  a.z_prev = d->z_prev; a.scale_prev = d->scale_prev; a.shift_prev = d->shift_prev;
  a.w_image = d->w_image; a.z = d->z; a.stats = d->stats;
  a.zmax = d->zmax; a.zmin = d->zmin; a.amax = d->amax; a.amin = d->amin;
  a.dysel = d->dysel; a.asel = d->asel; a.bw_k1 = d->bw_k1; a.bw_k2 = d->bw_k2;
  a.bw_mean = d->bw_mean; a.bw_invstd = d->bw_invstd; a.bw_gs = d->bw_gs; a.dz = d->dz;
  a.Kp = packed_k(d->Cin, d->mode == 0);
  a.Cout_pad = (d->Cout + 127) & ~127;
  a.chf_shift = pow2_shift(((d->Cin - 3 + 3) & ~3) >> 2);
  const long long M = (long long)d->B * d->NP * d->NS;
  if (d->mode == 0) {
    B2R_REQUIRE(d->Cin >= 3 && d->xyz && d->new_xyz && d->idx && (d->feat_t || d->Cin == 3),
                "b2r_sa_layer_fwd: gather mode needs xyz, new_xyz, idx (and feat_t when Cin > 3)");
    B2R_REQUIRE(d->Cin == 3 || ((reinterpret_cast<uintptr_t>(d->feat_t) & 15u) == 0),
                "b2r_sa_layer_fwd: feat_t must be 16-byte aligned");
  } else {
    B2R_REQUIRE(d->z_prev && d->scale_prev && d->shift_prev && (d->Cin % 4) == 0,
                "b2r_sa_layer_fwd: dense mode needs z_prev/scale/shift and Cin %% 4 == 0");
  }
  if (d->epilogue == 0) {
    B2R_REQUIRE(d->z != nullptr, "b2r_sa_layer_fwd: epilogue 0 needs z");
  } else {
    if (d->epilogue == 1)
      B2R_REQUIRE(d->zmax && d->zmin && d->amax && d->amin,
                  "b2r_sa_layer_fwd: epilogue 1 needs pool outputs");
    else
      B2R_REQUIRE(d->dysel && d->asel && d->bw_k1 && d->bw_k2 && d->bw_mean && d->bw_invstd &&
                      d->bw_gs && d->dz,
                  "b2r_sa_layer_fwd: epilogue 2 needs dysel/asel/k1/k2/mean/invstd/gs/dz");
    if (!(d->NS == 16 || d->NS == 32 || d->NS == 64)) {
      set_error("b2r_sa_layer_fwd: pooling supports nsample 16/32/64 (got %d)", d->NS);
      return B2R_ERR_UNSUPPORTED;
    }
  }
  if (a.Cout_pad > 256 || (M % 128) != 0 ||
      (d->mode == 0 && ((long long)d->NP * d->NS) % 128 != 0)) {
    set_error("b2r_sa_layer_fwd: needs Cout <= 256, B*NP*NS %% 128 == 0 and (gather layers) "
              "NP*NS %% 128 == 0 (Cout=%d, M=%lld, NP*NS=%lld)", d->Cout, M,
              (long long)d->NP * d->NS);
    return B2R_ERR_UNSUPPORTED;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int MT = a.Cout_pad >> 7;
  // tile width: 128 positions when one M tile and the operands fit, else 64
  int NT = 64;
  if (MT == 1 && layer_smem_bytes(a.Kp, a.Cout_pad, 128) <= 227 * 1024) NT = 128;
  if (MT == 1 && NT == 64 && d->epilogue >= 1) {
    set_error("b2r_sa_layer_fwd: pooling layer with Cout<=128 needs K small enough for 128-wide tiles");
    return B2R_ERR_UNSUPPORTED;
  }
  const size_t smem = layer_smem_bytes(a.Kp, a.Cout_pad, NT);
  if (smem > 227 * 1024) {
    set_error("b2r_sa_layer_fwd: operands need %zu bytes of shared memory (Cin=%d Cout=%d)", smem,
              d->Cin, d->Cout);
    return B2R_ERR_UNSUPPORTED;
  }
  a.num_tiles = (int)(M / NT);
  const int grid = a.num_tiles < kNumSMs ? a.num_tiles : kNumSMs;
  if (NT == 128) {
    B2R_CUDA(cudaFuncSetAttribute(sa_layer_fwd_kernel<128>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    sa_layer_fwd_kernel<128><<<grid, kMlpThreads, smem, st>>>(a);
  } else {
    B2R_CUDA(cudaFuncSetAttribute(sa_layer_fwd_kernel<64>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    sa_layer_fwd_kernel<64><<<grid, kMlpThreads, smem, st>>>(a);
  }
  B2R_CHECK_LAUNCH();
  return B2R_OK;
}

extern "C" int b2r_bn_finalize(const double *stats, int C, double count, const float *gamma,
                               const float *beta, float eps, float momentum, float *running_mean,
                               float *running_var, float *scale, float *shift, float *mean_out,
                               float *invstd_out, long long *num_batches_tracked, void *stream) {
  B2R_REQUIRE(stats && scale && shift && C > 0 && count > 0, "b2r_bn_finalize: bad argument");
  bn_finalize_kernel<<<ceil_div(C, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
      stats, C, count, gamma, beta, eps, momentum, running_mean, running_var, scale, shift,
      mean_out, invstd_out, num_batches_tracked);
  B2R_CHECK_LAUNCH();
  return B2R_OK;
}

extern "C" int b2r_pool_finalize(const float *zmax, const float *zmin, const float *scale,
                                 const float *shift, int B, int NP, int C, float *out_cm,
                                 float *out_pm, void *stream) {
  B2R_REQUIRE(zmax && zmin && scale && shift && B > 0 && NP > 0 && C > 0,
              "b2r_pool_finalize: bad argument");
  const long long total = (long long)B * NP * C;
  pool_finalize_kernel<<<ceil_div(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      zmax, zmin, scale, shift, B, NP, C, out_cm, out_pm);
  B2R_CHECK_LAUNCH();
  return B2R_OK;
}

extern "C" int b2r_to_point_major(const float *in, int B, int C, int N, float *out, void *stream) {
  B2R_REQUIRE(B >= 0 && C >= 0 && N >= 0, "b2r_to_point_major: negative size");
  if (B == 0 || C == 0 || N == 0) return B2R_OK;
  B2R_REQUIRE(in && out, "b2r_to_point_major: null pointer");
  dim3 grid(ceil_div(N, 32), ceil_div(C, 32), B), block(32, 8);
  to_point_major_kernel<<<grid, block, 0, static_cast<cudaStream_t>(stream)>>>(in, C, N, out);
  B2R_CHECK_LAUNCH();
  return B2R_OK;
}
