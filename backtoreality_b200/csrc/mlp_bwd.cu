// mlp_bwd.cu -- backward of one SharedMLP layer of a set-abstraction block on tcgen05 (sm_100a).
//
// Replaces, for one [1x1 conv -> BatchNorm2d -> ReLU] layer l inside PointnetSAModuleVotes
// (reference pointnet2_modules.py:245-267 / pytorch_utils.py:11-36,67-120), what autograd runs as
//   threshold_backward (ReLU) -> cudnn bn_bw -> cuDNN dgrad conv -> cuDNN wgrad conv
//   (+ for layer 0: cat backward, div/sub backward, two group_points_grad scatter kernels)
// i.e. ~8 full passes over (B, C, npoint, nsample) tensors, by ONE persistent kernel:
//
//   prologue  per tile of NT positions, build in 128B-swizzled shared memory
//               DZ[pos, co] = a[co]*gr[pos,co] + b[co]*z[pos,co] + c[co]        (BatchNorm backward:
//                             a = gamma*invstd, b = -a*k2*invstd, c = -a*k1 - b*mean, from
//                             b2r_bn_bwd_finalize; or a precomputed dz for the pooled top layer)
//               X[pos, k]   = the layer's input, RECOMPUTED exactly like the forward prologue
//                             (gathered rows, or relu(bn(z_prev)))
//   MMA       dgrad  D2[k, pos]  = sum_co W[co,k] * DZ[pos,co]      fresh per tile
//             wgrad  D3[co, k]  += sum_pos DZ[pos,co] * X[pos,k]    accumulated in TMEM over ALL tiles
//             wgrad contracts over POSITIONS, i.e. it needs both tiles "transposed".  tcgen05 can
//             read an operand MN-major straight from shared memory, but for 32-bit (TF32) data
//             only in the SWIZZLE_128B_BASE32B layout (128 B of MN x 4 K rows per atom, 32-byte
//             chunks XOR-swizzled with the row) -- measured with scripts/probe/umma_probe.cu,
//             profiles/r01/umma_probe_tf32_operand_layouts.log: the plain SW128 / no-swizzle
//             MN-major views of TF32 data give wrong products.  So the prologue writes DZ twice
//             (K-major SW128 for dgrad, BASE32B for wgrad) and X once (BASE32B); no data is ever
//             transposed through registers or re-read from HBM.  dgrad's A operand is a K-major
//             image of W^T packed once per step by b2r_mlp_pack_weight_t.
//   epilogue  dense layers: gr_prev = D2 * [relu(bn(z_prev)) > 0] stored position-major, plus the
//             per-channel sums  sum(gr_prev), sum(gr_prev * z_prev)  the next (lower) layer's
//             BatchNorm backward needs -- fused so gr_prev is written once and never re-read for
//             statistics;  gather layer: D2 rows are scatter-added (red.global.add.f32) straight
//             into the POINT-major feature gradient (B,N,C) / xyz gradients: the
//             (B,3+C,npoint,nsample) gradient tensor never exists.
//   end       D3 -> atomicAdd into dW (Cout x Cin, the nn.Conv2d layout).
#include "mlp_common.cuh"

namespace b2r {
using namespace mlp;
namespace {

struct BwdArgs {
  int B, N, NP, NS, Cin, Cout, mode;
  const float *xyz, *new_xyz, *feat_t;
  const int *idx;
  float radius;
  int normalize_xyz;
  const float *z_prev, *scale_prev, *shift_prev;
  const float *w_image;
  const float *dz;                              // direct dz (M,Cout), or NULL and:
  const float *gr, *z, *coef_a, *coef_b, *coef_c;
  float *dW;
  float *gr_prev;
  double *stats_prev;
  float *g_feat_t, *g_xyz, *g_new_xyz;
  int do_dgrad, do_wgrad;
  int Kp, KA, KAl, Cout_pad, num_tiles;   // KA: 32-wide atoms of packed K, KAl: of Cout
  int chf_shift;
};

// Shared-memory carve-up (host + device agree through this one function).
struct BwdSmem {
  uint32_t w_off, x_off, dzk_off, dz32_off, coef_off, scale_off, idx_off, bar_off, total;
  uint32_t w_bytes, x_bytes, dzk_bytes, dz32_bytes;
};
__host__ __device__ inline BwdSmem bwd_smem_layout(int Kp, int KA, int Cout, int Cout_pad, int NT,
                                                   int do_dgrad, int do_wgrad, int has_coef) {
  BwdSmem s;
  const uint32_t KAl = (uint32_t)(Cout + 31) >> 5;
  const uint32_t mt_p = (uint32_t)(KA + 3) >> 2;
  s.w_bytes = do_dgrad ? mt_p * 128u * KAl * 128u : 0u;        // W^T image: (packed K rows, Cout)
  s.x_bytes = do_wgrad ? (uint32_t)NT * KA * 128u : 0u;         // X, BASE32B
  s.dzk_bytes = do_dgrad ? (uint32_t)NT * KAl * 128u : 0u;      // DZ, K-major SW128
  s.dz32_bytes = do_wgrad ? (uint32_t)NT * (Cout_pad >> 5) * 128u : 0u;   // DZ, BASE32B
  s.w_off = 0;
  s.x_off = s.w_off + s.w_bytes;
  s.dzk_off = s.x_off + s.x_bytes;
  s.dz32_off = s.dzk_off + s.dzk_bytes;
  s.coef_off = s.dz32_off + s.dz32_bytes;
  s.scale_off = s.coef_off + (has_coef ? 3u * Cout * 4u : 0u);
  s.idx_off = s.scale_off + 2u * Kp * 4u;
  s.bar_off = (s.idx_off + 2u * (uint32_t)NT * 4u + 15u) & ~15u;   // idx double-buffered
  s.total = s.bar_off + 2 * 8 + 16 + 1024;  // + alignment slack
  return s;
}

// byte offset of 16-byte chunk `chunk` (4 consecutive channels) of position `pos` inside an
// MN-major SWIZZLE_128B_BASE32B tile of NT positions: 32-channel atoms NT*128 B apart; inside
// an atom 4-position groups of 512 B, one 128 B row per position, 32 B chunk index XOR (pos & 3)
__device__ __forceinline__ uint32_t t32_off(int pos, int chunk, int NT) {
  const int at = chunk >> 3, c8 = (chunk >> 1) & 3, half = chunk & 1;
  return (uint32_t)(at * NT * 128 + ((pos >> 2) << 9) + ((pos & 3) << 7) + ((c8 ^ (pos & 3)) << 5) +
                    (half << 4));
}
__device__ __forceinline__ uint64_t smem_desc_t32(uint32_t saddr, uint32_t lbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fffu);
  d |= (uint64_t)((lbo >> 4) & 0x3fffu) << 16;   // stride between 32-channel MN atoms
  d |= (uint64_t)(512u >> 4) << 32;              // stride between 4-position K groups
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;                        // SWIZZLE_128B_BASE32B
  return d;
}

// W (Cout, Cin) -> K-major SW128 image of W^T: rows = packed K (gather order, padded to whole
// 128-row M tiles), contraction dimension = Cout (padded to 32-element atoms), TF32-rounded
__global__ void pack_weight_t_kernel(const float *__restrict__ w, int Cout, int Cin, int gather,
                                     int Kp, int rows_t, int KAl, float *__restrict__ image) {
  const long long total = (long long)rows_t * KAl * 32;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const int kp = (int)(e / (KAl * 32)), co = (int)(e % (KAl * 32));
    float v = 0.f;
    if (kp < Kp && co < Cout) {
      int k = -1;
      if (gather) {
        const int C = Cin - 3, Cf4 = (C + 3) & ~3;
        if (kp < C) k = 3 + kp;
        else if (kp >= Cf4 && kp < Cf4 + 3) k = kp - Cf4;
      } else if (kp < Cin) {
        k = kp;
      }
      if (k >= 0) v = __uint_as_float(to_tf32(w[(size_t)co * Cin + k]));
    }
    image[(sw128_off(kp, co >> 2, rows_t) + (co & 3) * 4) >> 2] = v;
  }
}

template <int NT>
__global__ void __launch_bounds__(kMlpThreads, 1) sa_layer_bwd_kernel(const BwdArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *base = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                              ~(uintptr_t)1023);
  const bool has_coef = a.dz == nullptr;
  const BwdSmem L = bwd_smem_layout(a.Kp, a.KA, a.Cout, a.Cout_pad, NT, a.do_dgrad, a.do_wgrad,
                                    has_coef);
  uint8_t *s_w = base + L.w_off;
  uint8_t *s_x = base + L.x_off;
  uint8_t *s_dzk = base + L.dzk_off;
  uint8_t *s_dz32 = base + L.dz32_off;
  float *s_ca = reinterpret_cast<float *>(base + L.coef_off);
  float *s_cb = s_ca + a.Cout;
  float *s_cc = s_cb + a.Cout;
  float *s_scale = reinterpret_cast<float *>(base + L.scale_off);
  float *s_shift = s_scale + a.Kp;
  int *s_idx2 = reinterpret_cast<int *>(base + L.idx_off);
  uint64_t *s_bar = reinterpret_cast<uint64_t *>(base + L.bar_off);
  uint32_t *s_tmem = reinterpret_cast<uint32_t *>(s_bar + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q = warp & 3, h = warp >> 2;
  const int KA = a.KA, KAl = a.KAl;
  const int MTp = (KA + 3) >> 2;          // 128-row M tiles of the dgrad output (packed K rows)
  const int MTl = a.Cout_pad >> 7;        // 128-row M tiles of the wgrad output (Cout rows)
  constexpr int NCH = NT / 32;            // 32-column chunks per dgrad M tile
  const uint32_t bar_w = smem_u32(&s_bar[0]), bar_mma = smem_u32(&s_bar[1]);
  const uint32_t d3_col0 = a.do_dgrad ? (uint32_t)MTp * NT : 0u;
  constexpr uint32_t kTmemCols = 512;

  if (tid == 0) {
    mbar_init(bar_w, 1);
    mbar_init(bar_mma, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(smem_u32(s_tmem), kTmemCols);
  {  // zero the three tiles once (adjacent regions): padding chunks are never written again
    const uint32_t zb = L.x_bytes + L.dzk_bytes + L.dz32_bytes;
    for (uint32_t i = tid * 16; i < zb; i += kMlpThreads * 16)
      *reinterpret_cast<uint4 *>(s_x + i) = make_uint4(0, 0, 0, 0);
  }
  if (a.mode == 1)
    for (int i = tid; i < a.Cin; i += kMlpThreads) {
      s_scale[i] = a.scale_prev[i];
      s_shift[i] = a.shift_prev[i];
    }
  if (has_coef)
    for (int i = tid; i < a.Cout; i += kMlpThreads) {
      s_ca[i] = a.coef_a[i];
      s_cb[i] = a.coef_b[i];
      s_cc[i] = a.coef_c[i];
    }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;
  bool w_ready = !a.do_dgrad;
  if (tid == 0 && a.do_dgrad) {
    mbar_expect_tx(bar_w, L.w_bytes);
    bulk_g2s(smem_u32(s_w), a.w_image, L.w_bytes, bar_w);
  }

  const long long per_scene = (long long)a.NP * a.NS;
  const int C = a.Cin - 3, Cf4 = (C + 3) & ~3;   // gather mode: feature channels
  GatherSrc gsrc;
  gsrc.xyz = a.xyz; gsrc.new_xyz = a.new_xyz; gsrc.feat_t = a.feat_t;
  gsrc.N = a.N; gsrc.NP = a.NP; gsrc.NS = a.NS; gsrc.C = C; gsrc.Cf4 = Cf4;
  gsrc.chf_shift = a.chf_shift; gsrc.radius = a.radius; gsrc.normalize_xyz = a.normalize_xyz;
  const int CHl = a.Cout >> 2;                   // 16-byte chunks per DZ row
  const int KSl = (a.Cout + 7) >> 3;             // dgrad K steps (8 output channels each)
  const uint32_t idesc_dgrad = idesc_tf32(NT);   // A = W^T image, B = DZ, both K-major
  const uint32_t rows_t = (uint32_t)MTp * 128u;
  const uint32_t lbo_t = (uint32_t)NT * 128u;    // BASE32B tiles: 32-channel atoms
  uint32_t mma_parity = 0;
  float s1[3] = {0.f, 0.f, 0.f}, s2[3] = {0.f, 0.f, 0.f};   // per owned channel (by M tile)
  double d1[3] = {0.0, 0.0, 0.0}, d2[3] = {0.0, 0.0, 0.0};
  int tile_iter = 0;

  for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x, ++tile_iter) {
    const long long pos0 = (long long)tile * NT;
    // ball-query indices of this tile, double-buffered by tile parity: the scatter epilogue of
    // the previous tile may still be reading its copy while fast warps start this prologue
    int *s_idx = s_idx2 + (tile_iter & 1) * NT;

    // ---- prologue 1: DZ tile, written in both operand layouts ----------------------------------
    // (loads batched 4-deep per thread; the next tile of this CTA is bulk-prefetched into L2)
    {
      const int total = NT * CHl;
      const size_t o0 = (size_t)pos0 * a.Cout;
      if (tid == 0 && tile + (int)gridDim.x < a.num_tiles) {
        const size_t on = (size_t)(pos0 + (long long)gridDim.x * NT) * a.Cout;
        if (!has_coef) {
          prefetch_l2(a.dz + on, (uint32_t)total * 16u);
        } else {
          prefetch_l2(a.gr + on, (uint32_t)total * 16u);
          prefetch_l2(a.z + on, (uint32_t)total * 16u);
        }
        if (a.mode == 1)
          prefetch_l2(a.z_prev + (size_t)(pos0 + (long long)gridDim.x * NT) * a.Cin,
                      (uint32_t)NT * a.Cin * 4u);
      }
      for (int i0 = tid; i0 < total; i0 += kMlpThreads * 4) {
        float4 g[4], zz[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int i = i0 + u * kMlpThreads;
          if (i < total) {
            if (!has_coef) {
              g[u] = __ldg(reinterpret_cast<const float4 *>(a.dz + o0) + i);
            } else {
              g[u] = __ldg(reinterpret_cast<const float4 *>(a.gr + o0) + i);
              zz[u] = __ldg(reinterpret_cast<const float4 *>(a.z + o0) + i);
            }
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int i = i0 + u * kMlpThreads;
          if (i < total) {
            const int row = i / CHl, ch = i - row * CHl;
            float4 v = g[u];
            if (has_coef) {
              const float4 ca = *reinterpret_cast<const float4 *>(s_ca + ch * 4);
              const float4 cb = *reinterpret_cast<const float4 *>(s_cb + ch * 4);
              const float4 cc = *reinterpret_cast<const float4 *>(s_cc + ch * 4);
              v.x = fmaf(ca.x, g[u].x, fmaf(cb.x, zz[u].x, cc.x));
              v.y = fmaf(ca.y, g[u].y, fmaf(cb.y, zz[u].y, cc.y));
              v.z = fmaf(ca.z, g[u].z, fmaf(cb.z, zz[u].z, cc.z));
              v.w = fmaf(ca.w, g[u].w, fmaf(cb.w, zz[u].w, cc.w));
            }
            const uint4 out = make_uint4(to_tf32(v.x), to_tf32(v.y), to_tf32(v.z), to_tf32(v.w));
            if (a.do_dgrad) *reinterpret_cast<uint4 *>(s_dzk + sw128_off(row, ch, NT)) = out;
            if (a.do_wgrad) *reinterpret_cast<uint4 *>(s_dz32 + t32_off(row, ch, NT)) = out;
          }
        }
      }
    }
    // ---- prologue 2: X tile (same recomputation as the forward kernel), BASE32B layout ---------
    int tile_b = 0, in_scene0 = 0;   // gather layers: the tile's scene, first position in it
    if (a.mode == 0) {
      if (tid < NT) s_idx[tid] = a.idx[pos0 + tid];
      __syncthreads();
      tile_b = (int)(pos0 / per_scene);
      in_scene0 = (int)(pos0 - (long long)tile_b * per_scene);
    }
    if (a.do_wgrad) {
      if (a.mode == 0) {
        build_x_gather<NT>(gsrc, tile_b, in_scene0, s_idx, s_x, tid,
                           [](int row, int ch) { return t32_off(row, ch, NT); });
      } else {
        const int CH = a.Cin >> 2;
        const int total = NT * CH;
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
              *reinterpret_cast<uint4 *>(s_x + t32_off(row, ch, NT)) = out;
            }
          }
        }
      }
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();

    // ---- MMA issue (one thread) ---------------------------------------------------------------
    if (tid == 0) {
      if (!w_ready) {
        mbar_wait(bar_w, 0);
        w_ready = true;
      }
      tc_fence_after();
      const uint32_t xa = smem_u32(s_x), wa = smem_u32(s_w);
      const uint32_t dka = smem_u32(s_dzk), d32a = smem_u32(s_dz32);
      if (a.do_dgrad) {
        // D2[m-tile of packed K rows, NT positions] = W^T image (K-major) * DZ (K-major)
        for (int m = 0; m < MTp; ++m)
          for (int ks = 0; ks < KSl; ++ks) {
            const uint64_t adesc =
                smem_desc_sw128(wa + (uint32_t)(ks >> 2) * 1024u * (rows_t >> 3) +
                                (uint32_t)m * 16u * 1024u + (uint32_t)(ks & 3) * 32u);
            const uint64_t bdesc = smem_desc_sw128(dka + (uint32_t)(ks >> 2) * 1024u * (NT >> 3) +
                                                   (uint32_t)(ks & 3) * 32u);
            umma_tf32(tmem_base + (uint32_t)m * NT, adesc, bdesc, idesc_dgrad, ks > 0 ? 1u : 0u);
          }
      }
      if (a.do_wgrad) {
        // D3[m-tile of Cout rows, packed K columns] += DZ^T * X  (both MN-major, K = positions)
        for (int ml = 0; ml < MTl; ++ml)
          for (int a0 = 0; a0 < KA; a0 += 8) {
            const int na = (KA - a0) < 8 ? (KA - a0) : 8;
            const uint32_t idesc_w = idesc_tf32_ex(na * 32, 1, 1);
            for (int k8 = 0; k8 < NT / 8; ++k8) {
              const uint64_t adesc = smem_desc_t32(
                  d32a + (uint32_t)(4 * ml) * lbo_t + (uint32_t)k8 * 1024u, lbo_t);
              const uint64_t bdesc =
                  smem_desc_t32(xa + (uint32_t)a0 * lbo_t + (uint32_t)k8 * 1024u, lbo_t);
              umma_tf32(tmem_base + d3_col0 + (uint32_t)(ml * KA + a0) * 32u, adesc, bdesc,
                        idesc_w, (tile_iter > 0 || k8 > 0) ? 1u : 0u);
            }
          }
      }
      umma_commit(bar_mma);
    }

    // ---- epilogue: dgrad accumulator -> masked gradient + statistics, or scatter ---------------
    mbar_wait(bar_mma, mma_parity);
    mma_parity ^= 1u;
    tc_fence_after();
    if (a.do_dgrad) {
#pragma unroll
      for (int m = 0; m < 3; ++m) {
#pragma unroll
        for (int cc = 0; cc < NCH; ++cc) {
          if (m >= MTp || ((m * NCH + cc) & 1) != h) continue;
          uint32_t r[32];
          cuda::ptx::tcgen05_ld_32x32b(
              r, tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(m * NT + cc * 32));
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          const int kp = m * 128 + q * 32 + lane;   // packed K row owned by this thread
          const long long p0 = pos0 + cc * 32;
          if (a.mode == 1) {
            if (kp < a.Cin) {
              const float sc = s_scale[kp], sh = s_shift[kp];
              const float *zp = a.z_prev + (size_t)p0 * a.Cin + kp;
              float *gp = a.gr_prev + (size_t)p0 * a.Cin + kp;
              float t1 = 0.f, t2 = 0.f;
              float zv[32];   // all 32 (L2-resident) loads in flight before the first store
#pragma unroll
              for (int i = 0; i < 32; ++i) zv[i] = __ldg(zp + (size_t)i * a.Cin);
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                const float g = fmaf(zv[i], sc, sh) > 0.f ? __uint_as_float(r[i]) : 0.f;
                gp[(size_t)i * a.Cin] = g;
                t1 += g;
                t2 = fmaf(g, zv[i], t2);
              }
              s1[m] += t1;
              s2[m] += t2;
            }
          } else {
            // gather layer: scatter-add into the point-major feature / xyz gradients
            const bool is_feat = kp < C;
            const int e = kp - Cf4;                 // 0..2 for dx,dy,dz rows
            const bool is_xyz = (e >= 0 && e < 3);
            if (is_feat && a.g_feat_t != nullptr) {
              float *gb = a.g_feat_t + (size_t)tile_b * a.N * C + kp;
#pragma unroll
              for (int i = 0; i < 32; ++i)
                atomicAdd(gb + (size_t)s_idx[cc * 32 + i] * C, __uint_as_float(r[i]));
            } else if (is_xyz && (a.g_xyz != nullptr || a.g_new_xyz != nullptr)) {
              float run = 0.f;
              float *gx = a.g_xyz != nullptr ? a.g_xyz + (size_t)tile_b * a.N * 3 + e : nullptr;
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                const int ins = in_scene0 + cc * 32 + i;   // position inside the scene
                float v = __uint_as_float(r[i]);
                if (a.normalize_xyz) v = __fdiv_rn(v, a.radius);
                if (gx != nullptr) atomicAdd(gx + (size_t)s_idx[cc * 32 + i] * 3, v);
                run += v;
                if (((ins + 1) % a.NS) == 0 || i == 31) {   // end of this centre's run
                  if (a.g_new_xyz != nullptr)
                    atomicAdd(a.g_new_xyz + ((size_t)tile_b * a.NP + ins / a.NS) * 3 + e, -run);
                  run = 0.f;
                }
              }
            }
          }
        }
      }
      tc_fence_before();
      if (a.mode == 1) {
#pragma unroll
        for (int m = 0; m < 3; ++m) {
          d1[m] += (double)s1[m];
          d2[m] += (double)s2[m];
          s1[m] = 0.f;
          s2[m] = 0.f;
        }
      }
    }
  }

  if (a.do_dgrad && a.mode == 1 && a.stats_prev != nullptr) {
#pragma unroll
    for (int m = 0; m < 3; ++m) {
      const int kp = m * 128 + q * 32 + lane;
      // this thread's warp quad handled a chunk of M tile m iff NCH == 2 or the parity matches
      const bool mine = (NCH == 2) || ((m & 1) == h);
      if (m < MTp && mine && kp < a.Cin) {
        atomicAdd(a.stats_prev + kp, d1[m]);
        atomicAdd(a.stats_prev + a.Cin + kp, d2[m]);
      }
    }
  }

  // ---- wgrad accumulator -> dW (Cout, Cin) ------------------------------------------------------
  if (a.do_wgrad && tile_iter > 0) {
    // the last tile's commit (already waited on above) covers every wgrad MMA
    tc_fence_after();
    for (int u = h; u < MTl * KA; u += 2) {
      const int ml = u / KA, at = u - ml * KA;
      uint32_t r[32];
      cuda::ptx::tcgen05_ld_32x32b(r, tmem_base + ((uint32_t)(q * 32) << 16) + d3_col0 +
                                          (uint32_t)(ml * KA + at) * 32u);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      const int co = ml * 128 + q * 32 + lane;
      if (co < a.Cout) {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const int kp = at * 32 + i;
          int k = -1;
          if (a.mode == 0) {
            if (kp < C) k = 3 + kp;
            else if (kp >= Cf4 && kp < Cf4 + 3) k = kp - Cf4;
          } else if (kp < a.Cin) {
            k = kp;
          }
          if (k >= 0) atomicAdd(a.dW + (size_t)co * a.Cin + k, __uint_as_float(r[i]));
        }
      }
    }
  }
  (void)KAl;
  if (tid == 0 && !w_ready) mbar_wait(bar_w, 0);
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, kTmemCols);
}

// ---- max-pool / ReLU / BatchNorm backward of the pooled top layer: the sparse part ------------
// Forward kept, per (centre, channel): zmax, zmin and their sample indices.  The pooled output
// was y = relu(scale*zsel + shift) with zsel = (scale >= 0 ? zmax : zmin).  Its gradient reaches
// exactly ONE sample per (centre, channel):
//   dysel = dout * [y > 0],  asel = arg of zsel
// and the BatchNorm-backward sums over all positions reduce to sums over centres:
//   sum(gr) = sum(dysel),   sum(gr * z) = sum(dysel * zsel).
__global__ void pool_bwd_prep_kernel(const float *__restrict__ dout_cm,
                                     const float *__restrict__ dout_pm,
                                     const float *__restrict__ zmax, const float *__restrict__ zmin,
                                     const int *__restrict__ amax, const int *__restrict__ amin,
                                     const float *__restrict__ scale,
                                     const float *__restrict__ shift, int B, int NP, int Cch,
                                     float *__restrict__ dysel, int *__restrict__ asel,
                                     double *__restrict__ stats) {
  // thread <-> fixed channel, strided over centres: coalesced on the (centre, c) arrays
  const int c = threadIdx.x % Cch;
  const int lanes_per_block = blockDim.x / Cch;        // centres handled side by side
  const int sub = threadIdx.x / Cch;
  if (sub >= lanes_per_block) return;
  const long long ncentres = (long long)B * NP;
  const float s = scale[c], sh = shift[c];
  double a1 = 0.0, a2 = 0.0;
  for (long long ce = (long long)blockIdx.x * lanes_per_block + sub; ce < ncentres;
       ce += (long long)gridDim.x * lanes_per_block) {
    const size_t o = (size_t)ce * Cch + c;
    const bool pos = s >= 0.f;
    const float zs = pos ? zmax[o] : zmin[o];
    const int as = pos ? amax[o] : amin[o];
    const float y = fmaf(zs, s, sh);
    float g = 0.f;
    if (dout_cm != nullptr) {
      const int b = (int)(ce / NP), j = (int)(ce % NP);
      g += dout_cm[((size_t)b * Cch + c) * NP + j];
    }
    if (dout_pm != nullptr) g += dout_pm[o];
    g = y > 0.f ? g : 0.f;
    dysel[o] = g;
    asel[o] = as;
    a1 += (double)g;
    a2 += (double)g * (double)zs;
  }
  atomicAdd(stats + c, a1);
  atomicAdd(stats + Cch + c, a2);
}

// BatchNorm backward bookkeeping of one layer from  S1 = sum(gr), S2 = sum(gr*z)  (one thread per
// channel).  training: dz = a*gr + b*z + c with
//    gs = gamma*invstd, k1 = S1/M, k2 = (S2 - mean*S1)*invstd/M, a = gs, b = -gs*k2*invstd,
//    c = -gs*k1 - b*mean;       eval (running statistics): dz = gs*gr.
// dgamma = (S2 - mean*S1)*invstd, dbeta = S1.  Also emits the epilogue-2 form (k1,k2,gs).
__global__ void bn_bwd_finalize_kernel(const double *__restrict__ stats, int Cch, double count,
                                       const float *__restrict__ gamma,
                                       const float *__restrict__ mean,
                                       const float *__restrict__ invstd, int training,
                                       float *__restrict__ coef_a, float *__restrict__ coef_b,
                                       float *__restrict__ coef_c, float *__restrict__ k1_out,
                                       float *__restrict__ k2_out, float *__restrict__ gs_out,
                                       float *__restrict__ dgamma, float *__restrict__ dbeta) {
  const int ch = blockIdx.x * blockDim.x + threadIdx.x;
  if (ch >= Cch) return;
  const double S1 = stats[ch], S2 = stats[Cch + ch];
  const double mu = (double)mean[ch], is = (double)invstd[ch];
  const double g = gamma ? (double)gamma[ch] : 1.0;
  const double sx = (S2 - mu * S1) * is;   // sum(gr * xhat)
  const double gs = g * is;
  double k1 = 0.0, k2 = 0.0;
  if (training) {
    k1 = S1 / count;
    k2 = sx / count;
  }
  const double b = -gs * k2 * is;
  if (coef_a) coef_a[ch] = (float)gs;
  if (coef_b) coef_b[ch] = (float)b;
  if (coef_c) coef_c[ch] = (float)(-gs * k1 - b * mu);
  if (k1_out) k1_out[ch] = (float)k1;
  if (k2_out) k2_out[ch] = (float)k2;
  if (gs_out) gs_out[ch] = (float)gs;
  if (dgamma) dgamma[ch] = (float)sx;
  if (dbeta) dbeta[ch] = (float)S1;
}

}  // namespace
}  // namespace b2r

using namespace b2r;

extern "C" long long b2r_mlp_weight_t_image_bytes(int Cout, int Cin, int gather) {
  if (Cout <= 0 || Cin <= 0) return 0;
  const int Kp = packed_k(Cin, gather);
  const int KA = (Kp + 31) >> 5, KAl = (Cout + 31) >> 5;
  return (long long)((KA + 3) >> 2) * 128 * KAl * 128;
}

extern "C" int b2r_mlp_pack_weight_t(const float *w, int Cout, int Cin, int gather, float *image,
                                     void *stream) {
  B2R_REQUIRE(w && image && Cout > 0 && Cin > 0, "b2r_mlp_pack_weight_t: bad argument");
  B2R_REQUIRE(!gather || Cin >= 3, "b2r_mlp_pack_weight_t: gather layers need Cin >= 3");
  const int Kp = packed_k(Cin, gather);
  const int KA = (Kp + 31) >> 5, KAl = (Cout + 31) >> 5;
  const int rows_t = ((KA + 3) >> 2) * 128;
  const long long total = (long long)rows_t * KAl * 32;
  pack_weight_t_kernel<<<ceil_div(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      w, Cout, Cin, gather, Kp, rows_t, KAl, image);
  B2R_CHECK_LAUNCH();
  return B2R_OK;
}

namespace {
// one launch with the given work split; returns B2R_ERR_UNSUPPORTED when it does not fit
int launch_bwd(BwdArgs a, int do_dgrad, int do_wgrad, long long M, cudaStream_t st) {
  a.do_dgrad = do_dgrad;
  a.do_wgrad = do_wgrad;
  const int MTp = (a.KA + 3) >> 2, MTl = a.Cout_pad >> 7;
  const int has_coef = a.dz == nullptr;
  int NT = 0;
  for (int nt : {64, 32}) {
    if (M % nt) continue;
    const BwdSmem L = bwd_smem_layout(a.Kp, a.KA, a.Cout, a.Cout_pad, nt, do_dgrad, do_wgrad,
                                      has_coef);
    const int cols = (do_dgrad ? MTp * nt : 0) + (do_wgrad ? MTl * a.KA * 32 : 0);
    if (L.total <= 227u * 1024u && cols <= 512 && MTp <= 3) {
      NT = nt;
      break;
    }
  }
  if (NT == 0) return B2R_ERR_UNSUPPORTED;
  const BwdSmem L = bwd_smem_layout(a.Kp, a.KA, a.Cout, a.Cout_pad, NT, do_dgrad, do_wgrad,
                                    has_coef);
  a.num_tiles = (int)(M / NT);
  const int grid = a.num_tiles < kNumSMs ? a.num_tiles : kNumSMs;
  if (NT == 64) {
    B2R_CUDA(cudaFuncSetAttribute(sa_layer_bwd_kernel<64>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total));
    sa_layer_bwd_kernel<64><<<grid, kMlpThreads, L.total, st>>>(a);
  } else {
    B2R_CUDA(cudaFuncSetAttribute(sa_layer_bwd_kernel<32>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total));
    sa_layer_bwd_kernel<32><<<grid, kMlpThreads, L.total, st>>>(a);
  }
  B2R_CHECK_LAUNCH();
  return B2R_OK;
}
}  // namespace

extern "C" int b2r_sa_layer_bwd(const b2r_sa_layer_bwd_desc *d, void *stream) {
  B2R_REQUIRE(d != nullptr, "b2r_sa_layer_bwd: null descriptor");
  B2R_REQUIRE(d->B > 0 && d->NP > 0 && d->NS > 0 && d->Cin > 0 && d->Cout > 0,
              "b2r_sa_layer_bwd: non-positive size");
  B2R_REQUIRE(d->mode == 0 || d->mode == 1, "b2r_sa_layer_bwd: mode must be 0 or 1");
  B2R_REQUIRE(d->dW != nullptr, "b2r_sa_layer_bwd: null dW");
  B2R_REQUIRE(d->dz != nullptr ||
                  (d->gr && d->z && d->coef_a && d->coef_b && d->coef_c),
              "b2r_sa_layer_bwd: needs dz, or gr + z + coef_a/b/c");
  BwdArgs a;
  a.B = d->B; a.N = d->N; a.NP = d->NP; a.NS = d->NS; a.Cin = d->Cin; a.Cout = d->Cout;
  a.mode = d->mode;
  a.xyz = d->xyz; a.new_xyz = d->new_xyz; a.feat_t = d->feat_t; a.idx = d->idx;
  a.radius = d->radius; a.normalize_xyz = d->normalize_xyz;
  a.z_prev = d->z_prev; a.scale_prev = d->scale_prev; a.shift_prev = d->shift_prev;
  a.w_image = d->w_image_t;
  a.dz = d->dz; a.gr = d->gr; a.z = d->z;
  a.coef_a = d->coef_a; a.coef_b = d->coef_b; a.coef_c = d->coef_c;
  a.dW = d->dW; a.gr_prev = d->gr_prev; a.stats_prev = d->stats_prev;
  a.g_feat_t = d->g_feat_t; a.g_xyz = d->g_xyz; a.g_new_xyz = d->g_new_xyz;
  a.Kp = packed_k(d->Cin, d->mode == 0);
  a.KA = (a.Kp + 31) >> 5;
  a.KAl = (d->Cout + 31) >> 5;
  a.Cout_pad = (d->Cout + 127) & ~127;
  a.chf_shift = pow2_shift(((d->Cin - 3 + 3) & ~3) >> 2);
  a.do_dgrad = a.do_wgrad = 1;
  a.num_tiles = 0;
  int need_dgrad;
  if (d->mode == 0) {
    B2R_REQUIRE(d->Cin >= 3 && d->xyz && d->new_xyz && d->idx && (d->feat_t || d->Cin == 3),
                "b2r_sa_layer_bwd: gather mode needs xyz, new_xyz, idx (and feat_t when Cin > 3)");
    need_dgrad = (d->g_feat_t != nullptr && d->Cin > 3) || d->g_xyz != nullptr ||
                 d->g_new_xyz != nullptr;
  } else {
    B2R_REQUIRE(d->z_prev && d->scale_prev && d->shift_prev && (d->Cin % 4) == 0,
                "b2r_sa_layer_bwd: dense mode needs z_prev/scale/shift and Cin %% 4 == 0");
    B2R_REQUIRE(d->gr_prev != nullptr, "b2r_sa_layer_bwd: dense mode needs gr_prev");
    need_dgrad = 1;
  }
  B2R_REQUIRE(!need_dgrad || d->w_image_t != nullptr,
              "b2r_sa_layer_bwd: null transposed weight image");
  const long long M = (long long)d->B * d->NP * d->NS;
  if (d->mode == 0 && ((long long)d->NP * d->NS) % 64 != 0) {
    set_error("b2r_sa_layer_bwd: gather layers need NP*NS %% 64 == 0 (got %lld)",
              (long long)d->NP * d->NS);
    return B2R_ERR_UNSUPPORTED;
  }
  if (a.Cout_pad > 256 || (d->Cout % 8) != 0 || (M % 32) != 0) {
    set_error("b2r_sa_layer_bwd: needs Cout %% 8 == 0, Cout <= 256, B*NP*NS %% 32 == 0 "
              "(Cout=%d, M=%lld)", d->Cout, M);
    return B2R_ERR_UNSUPPORTED;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // one fused launch when both GEMMs' operands fit in shared memory + TMEM, else two launches
  int rc = launch_bwd(a, need_dgrad, 1, M, st);
  if (rc == B2R_ERR_UNSUPPORTED && need_dgrad) {
    rc = launch_bwd(a, 0, 1, M, st);
    if (rc == B2R_OK) rc = launch_bwd(a, 1, 0, M, st);
  }
  if (rc == B2R_ERR_UNSUPPORTED)
    set_error("b2r_sa_layer_bwd: layer Cin=%d Cout=%d does not fit shared memory / TMEM", d->Cin,
              d->Cout);
  return rc;
}

extern "C" int b2r_pool_bwd_prep(const float *dout_cm, const float *dout_pm, const float *zmax,
                                 const float *zmin, const int *amax, const int *amin,
                                 const float *scale, const float *shift, int B, int NP, int C,
                                 float *dysel, int *asel, double *stats, void *stream) {
  B2R_REQUIRE((dout_cm || dout_pm) && zmax && zmin && amax && amin && scale && shift && dysel &&
                  asel && stats && B > 0 && NP > 0 && C > 0 && C <= 1024,
              "b2r_pool_bwd_prep: bad argument");
  const int threads = C <= 256 ? 256 : 1024;
  const int per_block = threads / C;
  int grid = ceil_div((long long)B * NP, per_block);
  if (grid > kNumSMs * 4) grid = kNumSMs * 4;
  pool_bwd_prep_kernel<<<grid, threads, 0, static_cast<cudaStream_t>(stream)>>>(
      dout_cm, dout_pm, zmax, zmin, amax, amin, scale, shift, B, NP, C, dysel, asel, stats);
  B2R_CHECK_LAUNCH();
  return B2R_OK;
}

extern "C" int b2r_bn_bwd_finalize(const double *stats, int C, double count, const float *gamma,
                                   const float *mean, const float *invstd, int training,
                                   float *coef_a, float *coef_b, float *coef_c, float *k1,
                                   float *k2, float *gs, float *dgamma, float *dbeta,
                                   void *stream) {
  B2R_REQUIRE(stats && mean && invstd && C > 0 && count > 0, "b2r_bn_bwd_finalize: bad argument");
  bn_bwd_finalize_kernel<<<ceil_div(C, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
      stats, C, count, gamma, mean, invstd, training, coef_a, coef_b, coef_c, k1, k2, gs, dgamma,
      dbeta);
  B2R_CHECK_LAUNCH();
  return B2R_OK;
}
