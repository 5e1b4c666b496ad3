// dense.cu -- wide 1x1-conv layers on POINT-major tensors (M positions x C channels) on tcgen05:
// the SharedMLP of PointnetFPModule (reference pointnet2_modules.py:505-514: cat -> unsqueeze ->
// SharedMLP [512,256,256]), VotingModule's conv1..3 + bn1..2 (models/voting_module.py:38-65) and
// ProposalModule's conv1..3 + bn1..2 (models/proposal_module.py:115-119), forward and backward.
//
// These layers are too wide for the resident-weight design of mlp.cu (a 256x256 TF32 weight is
// 256 KB, a 512-channel input row 2 KB) and too small to care about it (2k-8k positions): every
// CTA computes ONE 128 x 128 output tile and STREAMS both operands through a 4-stage ring of
// 32-deep K chunks.  One kernel skeleton, three uses (all TF32 operands, FP32 accumulate):
//
//   FWD    z[pos, co]  = sum_k W[co,k] x[pos,k] (+ bias)      A = W image (TMA), B = x rows
//          x = in, or relu(in * scale + shift) applied while loading (the previous layer's
//          BatchNorm + ReLU); epilogue: per-channel sum / sum of squares + store z.
//   DGRAD  gin[pos, k] = sum_co W[co,k] dz[pos,co]            A = W^T image (TMA), B = dz rows
//          dz = ca*g + cb*z + cc (BatchNorm backward, coefficients from b2r_bn_bwd_finalize) or
//          g itself; epilogue: * [relu(bn(in)) > 0], the next BatchNorm-backward sums, store.
//   WGRAD  dW[co, k]  += sum_pos dz[pos,co] x[pos,k]          A = dz^T, B = x^T, both built by
//          the producers (positions are the contraction dimension, so the tiles are written
//          transposed into the K-major swizzled layout); split over position ranges (grid.z),
//          partial tiles added to dW with coalesced atomics.
//
// "Channels on TMEM lanes" as in mlp.cu: the accumulator row is a channel, so statistics and the
// BatchNorm-backward sums are per-thread register work.  Warp roles: 0-3 epilogue, 4 MMA issue,
// 5-11 producers (global loads -> transform -> TF32 -> swizzled smem); weights arrive by
// cp.async.bulk from pre-swizzled images (dense_pack_kernel), 16 KB per (row block, K chunk).
#include "mlp_common.cuh"

namespace b2r {
using namespace mlp;
namespace {

constexpr int kDT = 128;                       // tile: 128 A rows x 128 B rows
constexpr int kDK = 32;                        // K chunk: one 128-byte swizzle atom of TF32
constexpr uint32_t kDOperand = kDT * kDK * 4;  // 16 KB per operand per stage
// 12 warps: warp 4 issues the MMAs; the other 11 build operand tiles during the main loop
// (global loads -> transform -> TF32 -> swizzled smem), warps 0-3 then run the epilogue
constexpr int kDEpi = 128, kDProd = 352, kDThreads = 384;
constexpr int kDSlots = 3;                     // ceil(1024 float4 items / 352 threads)
// Per mode: operand stages S (swizzled TF32 tiles the tensor core reads), raw operands R per chunk
// (fp32 rows as they lie in global memory: x | g, z | g, z, x) and the depth D of the raw ring.
// Raw rows arrive by cp.async D-1 chunks ahead of their use, 16 bytes per thread into slots the
// SAME thread reads back (no barrier, no registers held across the global-memory latency).
// (The weight tile of a stage arrives by TMA once the stage is free, i.e. S chunks ahead: S also
// sets how much of the TMA latency is hidden.)
__host__ __device__ constexpr int dn_stages(int mode) { return mode == 0 ? 4 : mode == 1 ? 3 : 2; }
__host__ __device__ constexpr int dn_raws(int mode) { return mode == 0 ? 1 : mode == 1 ? 2 : 3; }
__host__ __device__ constexpr int dn_depth(int mode) { return mode == 0 ? 4 : 3; }
constexpr uint32_t kDRawSlot = kDSlots * kDProd * 16;   // bytes of one raw operand of one chunk
// per-channel coefficient vectors staged in shared memory (they are read once per item):
// [0, kDCoefX) floats: scale | shift of the x prologue (up to 512 + pad channels each),
// then ca | cb | cc of the dz prologue (up to 384 channels each)
constexpr int kDMaxCin = 512, kDMaxCoutBn = 512;
constexpr uint32_t kDCoefBytes = (2 * kDMaxCin + 3 * kDMaxCoutBn) * 4;
__host__ __device__ constexpr uint32_t dn_smem(int mode) {
  return (uint32_t)dn_stages(mode) * 2 * kDOperand +
         (uint32_t)dn_depth(mode) * dn_raws(mode) * kDRawSlot + kDCoefBytes +
         (2 * dn_stages(mode) + 1) * 8 + 16 + 1024;
}

struct DenseArgs {
  int M, Cin, Cout;
  const float *in;        // (M, ld_in): the layer input before its prologue
  int ld_in;
  const float *sc_in, *sh_in;   // NULL: x = in; else x = relu(in * sc + sh)
  // FWD
  const float *w_img;
  const float *bias;
  float *z;
  int ld_z;
  double *stats;
  // DGRAD / WGRAD
  const float *g, *zz;    // (M, ld_g) each; zz unused when ca == NULL
  int ld_g;
  const float *ca, *cb, *cc;
  const float *wt_img;
  float *gin;
  int ld_gin;
  double *stats_in;
  float *dW;              // (Cout, Cin)
  int split_len;          // WGRAD: positions per grid.z slice (multiple of 32)
};

// W (R x Kc, element (r,k) at w[r*ldw + k], or w[k*ldw + r] when transposed) -> TF32 image:
// blocks of 128 rows x 32 K (16 KB, SW128 K-major), block (rb, ka) at (rb * KA + ka) * 4096 floats
__global__ void dense_pack_kernel(const float *__restrict__ w, int R, int Kc, int ldw, int transposed,
                                  float *__restrict__ img) {
  const int KA = (Kc + kDK - 1) / kDK;
  const int Rp = (R + kDT - 1) / kDT * kDT;
  const long long total = (long long)Rp * KA * kDK;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(e / (KA * kDK)), k = (int)(e % (KA * kDK));
    float v = 0.f;
    if (r < R && k < Kc)
      v = __uint_as_float(to_tf32(transposed ? w[(size_t)k * ldw + r] : w[(size_t)r * ldw + k]));
    const size_t blk = (size_t)((r >> 7) * KA + (k >> 5)) * (kDT * kDK);
    img[blk + (sw128_off(r & 127, (k & 31) >> 2, kDT) >> 2) + (k & 3)] = v;
  }
}

__device__ __forceinline__ void dn_bar_prod() { asm volatile("bar.sync 2, %0;" ::"n"(kDProd)); }
__device__ __forceinline__ void dn_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// ---- raw load of one producer item (4 consecutive channels of one position): 16 bytes by
// ---- cp.async, zero-filled outside the tensor (src-size operand)
__device__ __forceinline__ void cp16_guard(uint32_t dst, const float *base, int ld, int pos, int ch,
                                           int M, int C) {
  const float *src = base;
  uint32_t bytes = 0;
  if (pos < M && ch < C) {
    src = base + (size_t)pos * ld + ch;
    bytes = (uint32_t)min(4, C - ch) * 4u;
  }
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(bytes)
               : "memory");
}
// x = in, or relu(in * sc + sh); zeros outside the tensor (coefficient vectors are padded to 4)
// (s_sc / s_sh: shared-memory copies of the coefficient vectors, indexed by ch - c0)
__device__ __forceinline__ uint4 make_x4(const DenseArgs &a, float4 v, int pos, int ch,
                                         const float *s_sc, const float *s_sh, int c0) {
  if (a.sc_in != nullptr && pos < a.M && ch < a.Cin) {
    const float4 sc = *reinterpret_cast<const float4 *>(s_sc + (ch - c0));
    const float4 sh = *reinterpret_cast<const float4 *>(s_sh + (ch - c0));
    v.x = fmaxf(fmaf(v.x, sc.x, sh.x), 0.f);
    v.y = ch + 1 < a.Cin ? fmaxf(fmaf(v.y, sc.y, sh.y), 0.f) : 0.f;
    v.z = ch + 2 < a.Cin ? fmaxf(fmaf(v.z, sc.z, sh.z), 0.f) : 0.f;
    v.w = ch + 3 < a.Cin ? fmaxf(fmaf(v.w, sc.w, sh.w), 0.f) : 0.f;
  }
  return make_uint4(to_tf32(v.x), to_tf32(v.y), to_tf32(v.z), to_tf32(v.w));
}
// dz = g, or ca*g + cb*z + cc; zeros outside the tensor
__device__ __forceinline__ uint4 make_dz4(const DenseArgs &a, float4 g, float4 z, int pos, int co,
                                          const float *s_ca, const float *s_cb, const float *s_cc,
                                          int c0) {
  if (a.ca != nullptr && pos < a.M && co < a.Cout) {
    const float4 A = *reinterpret_cast<const float4 *>(s_ca + (co - c0));
    const float4 B = *reinterpret_cast<const float4 *>(s_cb + (co - c0));
    const float4 C = *reinterpret_cast<const float4 *>(s_cc + (co - c0));
    g.x = fmaf(A.x, g.x, fmaf(B.x, z.x, C.x));
    g.y = co + 1 < a.Cout ? fmaf(A.y, g.y, fmaf(B.y, z.y, C.y)) : 0.f;
    g.z = co + 2 < a.Cout ? fmaf(A.z, g.z, fmaf(B.z, z.z, C.z)) : 0.f;
    g.w = co + 3 < a.Cout ? fmaf(A.w, g.w, fmaf(B.w, z.w, C.w)) : 0.f;
  }
  return make_uint4(to_tf32(g.x), to_tf32(g.y), to_tf32(g.z), to_tf32(g.w));
}
// MN-major TF32 operand tile (SWIZZLE_128B_BASE32B, validated by scripts/probe/umma_probe.cu):
// byte offset of element (r = MN index 0..127, k = 0..31); 4 consecutive r (r % 4 == 0) are 16
// contiguous bytes, so a position-major float4 is stored with one vector store
__device__ __forceinline__ uint32_t mn32_off(int r, int k) {
  return (uint32_t)((r >> 5) * 4096 + (k >> 2) * 512 + (k & 3) * 128 +
                    ((((r & 31) >> 3) ^ (k & 3)) << 5) + (r & 7) * 4);
}
__device__ __forceinline__ uint64_t smem_desc_mn32(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fffu);
  d |= (uint64_t)(4096u >> 4) << 16;   // LBO: next 32 MN elements
  d |= (uint64_t)(512u >> 4) << 32;    // SBO: next 4 K rows
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;              // SWIZZLE_128B_BASE32B
  return d;
}

// MODE 0 FWD, 1 DGRAD, 2 WGRAD
template <int MODE>
__global__ void __launch_bounds__(kDThreads, 1) dense_gemm_kernel(const DenseArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *base = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                              ~(uintptr_t)1023);
  constexpr int kDStages = dn_stages(MODE), kRaws = dn_raws(MODE), kDepth = dn_depth(MODE);
  uint8_t *s_raw = base + kDStages * 2 * kDOperand;
  float *s_sc = reinterpret_cast<float *>(s_raw + (size_t)kDepth * kRaws * kDRawSlot);
  float *s_sh = s_sc + kDMaxCin;
  float *s_ca = s_sh + kDMaxCin;
  float *s_cb = s_ca + kDMaxCoutBn;
  float *s_cc = s_cb + kDMaxCoutBn;
  uint64_t *s_bar = reinterpret_cast<uint64_t *>(s_cc + kDMaxCoutBn);
  uint32_t *s_tmem = reinterpret_cast<uint32_t *>(s_bar + 2 * kDStages + 1);
  // mbarriers: [0,S) full  [S,2S) empty  [2S] done
  auto bar = [&](int i) { return smem_u32(&s_bar[i]); };
  auto stage_a = [&](int s) { return base + (size_t)s * 2 * kDOperand; };
  auto stage_b = [&](int s) { return base + (size_t)s * 2 * kDOperand + kDOperand; };
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int bt = blockIdx.x;    // B tile: positions (FWD, DGRAD) / input channels (WGRAD)
  const int rb = blockIdx.y;    // A row block: output channels (FWD, WGRAD) / input channels (DGRAD)
  // contraction: FWD over Cin, DGRAD over Cout, WGRAD over this slice's positions
  int k_begin = 0, k_len = MODE == 0 ? a.Cin : a.Cout;
  if (MODE == 2) {
    k_begin = blockIdx.z * a.split_len;
    k_len = min(a.split_len, a.M - k_begin);
  }
  const int nch = (k_len + kDK - 1) / kDK;   // K chunks (FWD/DGRAD: also per row block of the image)

  if (tid == 0) {
    // full: the producers' arrival, plus (FWD/DGRAD) the arrive.expect_tx that precedes the TMA
    for (int i = 0; i < kDStages; ++i) mbar_init(bar(i), MODE == 2 ? 1 : 2);
    for (int i = kDStages; i <= 2 * kDStages; ++i) mbar_init(bar(i), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) tmem_alloc(smem_u32(s_tmem), 128);
  // coefficient vectors -> shared memory: the x prologue's (FWD: every input channel; WGRAD: this
  // CTA's 128-channel block) and the dz prologue's (DGRAD: every output channel; WGRAD: block)
  const int xc0 = MODE == 2 ? bt * kDT : 0, xcn = MODE == 2 ? kDT : ((a.Cin + 3) & ~3);
  const int zc0 = MODE == 2 ? rb * kDT : 0, zcn = MODE == 2 ? kDT : ((a.Cout + 3) & ~3);
  if (MODE != 1 && a.sc_in != nullptr)
    for (int i = tid; i < xcn; i += kDThreads) {
      const bool in = xc0 + i < ((a.Cin + 3) & ~3);
      s_sc[i] = in ? a.sc_in[xc0 + i] : 0.f;
      s_sh[i] = in ? a.sh_in[xc0 + i] : 0.f;
    }
  if (MODE != 0 && a.ca != nullptr)
    for (int i = tid; i < zcn; i += kDThreads) {
      const bool in = zc0 + i < ((a.Cout + 3) & ~3);
      s_ca[i] = in ? a.ca[zc0 + i] : 0.f;
      s_cb[i] = in ? a.cb[zc0 + i] : 0.f;
      s_cc[i] = in ? a.cc[zc0 + i] : 0.f;
    }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;

  if (warp != 4) {
    // =============================== PRODUCERS (11 warps) =======================================
    const int ptid = tid < 128 ? tid : tid - 32;
    // item i (0..1023) of a chunk: FWD/DGRAD: position row i >> 3, channel group i & 7 of the
    // chunk's 32 channels; WGRAD: position i >> 5 of the chunk's 32, channel group i & 31 of 128
    // this thread's 16-byte slot of raw operand `op` of chunk j, item u
    auto raw_slot = [&](int j, int op, int u) {
      return s_raw + (size_t)((j % kDepth) * kRaws + op) * kDRawSlot + (size_t)(u * kDProd + ptid) * 16;
    };
    auto issue_chunk = [&](int j) {   // one (possibly empty) cp.async group per call
      if (j < nch) {
#pragma unroll
        for (int u = 0; u < kDSlots; ++u) {
          const int i = ptid + u * kDProd;
          if (i < kDT * 8) {
            if (MODE == 0) {
              cp16_guard(smem_u32(raw_slot(j, 0, u)), a.in, a.ld_in, bt * kDT + (i >> 3),
                         j * kDK + (i & 7) * 4, a.M, a.Cin);
            } else if (MODE == 1) {
              const int pos = bt * kDT + (i >> 3), co = j * kDK + (i & 7) * 4;
              cp16_guard(smem_u32(raw_slot(j, 0, u)), a.g, a.ld_g, pos, co, a.M, a.Cout);
              if (a.ca != nullptr)
                cp16_guard(smem_u32(raw_slot(j, 1, u)), a.zz, a.ld_g, pos, co, a.M, a.Cout);
            } else {
              const int pl = i >> 5;
              const int pos = pl < k_len - j * kDK ? k_begin + j * kDK + pl : a.M;   // a.M: zeros
              const int co = rb * kDT + (i & 31) * 4, ch = bt * kDT + (i & 31) * 4;
              cp16_guard(smem_u32(raw_slot(j, 0, u)), a.g, a.ld_g, pos, co, a.M, a.Cout);
              if (a.ca != nullptr)
                cp16_guard(smem_u32(raw_slot(j, 1, u)), a.zz, a.ld_g, pos, co, a.M, a.Cout);
              cp16_guard(smem_u32(raw_slot(j, 2, u)), a.in, a.ld_in, pos, ch, a.M, a.Cin);
            }
          }
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    auto store_chunk = [&](int j) {
      const int s = j % kDStages, n = j / kDStages;
      mbar_wait(bar(kDStages + s), (uint32_t)((n & 1) ^ 1));   // MMAs of chunk j - S left stage s
      uint8_t *sa = stage_a(s), *sb = stage_b(s);
      if (MODE != 2 && ptid == 0) {   // weights: one 16 KB block of the pre-swizzled image
        const float *img = (MODE == 0 ? a.w_img : a.wt_img) + (size_t)(rb * nch + j) * (kDT * kDK);
        mbar_expect_tx(bar(s), kDOperand);
        bulk_g2s(smem_u32(sa), img, kDOperand, bar(s));
      }
#pragma unroll
      for (int u = 0; u < kDSlots; ++u) {
        const int i = ptid + u * kDProd;
        if (i < kDT * 8) {
          const float4 r0 = *reinterpret_cast<const float4 *>(raw_slot(j, 0, u));
          if (MODE == 0) {
            *reinterpret_cast<uint4 *>(sb + sw128_off(i >> 3, i & 7, kDT)) =
                make_x4(a, r0, bt * kDT + (i >> 3), j * kDK + (i & 7) * 4, s_sc, s_sh, 0);
          } else if (MODE == 1) {
            float4 r1 = r0;
            if (a.ca != nullptr) r1 = *reinterpret_cast<const float4 *>(raw_slot(j, 1, u));
            *reinterpret_cast<uint4 *>(sb + sw128_off(i >> 3, i & 7, kDT)) =
                make_dz4(a, r0, r1, bt * kDT + (i >> 3), j * kDK + (i & 7) * 4, s_ca, s_cb, s_cc, 0);
          } else {
            const int pl = i >> 5, c4 = i & 31;
            const int pos = pl < k_len - j * kDK ? k_begin + j * kDK + pl : a.M;
            float4 r1 = r0;
            if (a.ca != nullptr) r1 = *reinterpret_cast<const float4 *>(raw_slot(j, 1, u));
            const float4 r2 = *reinterpret_cast<const float4 *>(raw_slot(j, 2, u));
            *reinterpret_cast<uint4 *>(sa + mn32_off(c4 * 4, pl)) =
                make_dz4(a, r0, r1, pos, rb * kDT + c4 * 4, s_ca, s_cb, s_cc, zc0);
            *reinterpret_cast<uint4 *>(sb + mn32_off(c4 * 4, pl)) =
                make_x4(a, r2, pos, bt * kDT + c4 * 4, s_sc, s_sh, xc0);
          }
        }
      }
      fence_async_smem();
      dn_bar_prod();
      if (ptid == 0) dn_arrive(bar(s));
    };
    for (int j = 0; j < kDepth - 1; ++j) issue_chunk(j);
    for (int j = 0; j < nch; ++j) {
      issue_chunk(j + kDepth - 1);
      asm volatile("cp.async.wait_group %0;" ::"n"(kDepth - 1) : "memory");   // chunk j has landed
      store_chunk(j);
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
  } else {
    // =============================== MMA ISSUE ==================================================
    if (lane == 0) {
      const uint32_t idesc = MODE == 2 ? (idesc_tf32(kDT) | (1u << 15) | (1u << 16)) : idesc_tf32(kDT);
      for (int j = 0; j < nch; ++j) {
        const int s = j % kDStages, n = j / kDStages;
        mbar_wait(bar(s), (uint32_t)(n & 1));
        tc_fence_after();
        const uint32_t aa = smem_u32(stage_a(s)), ba = smem_u32(stage_b(s));
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          if (MODE == 2)
            umma_tf32(tmem_base, smem_desc_mn32(aa + ks * 1024u), smem_desc_mn32(ba + ks * 1024u),
                      idesc, (j > 0 || ks > 0) ? 1u : 0u);
          else
            umma_tf32(tmem_base, smem_desc_sw128(aa + ks * 32u), smem_desc_sw128(ba + ks * 32u),
                      idesc, (j > 0 || ks > 0) ? 1u : 0u);
        }
        umma_commit(bar(kDStages + s));
      }
      umma_commit(bar(2 * kDStages));
    }
    __syncwarp();
  }

  if (warp < 4 || warp >= 8) {
    // =============================== EPILOGUE (8 warps) =========================================
    // two warps per TMEM lane quadrant (warp & 3), each takes two of the four 32-column chunks
    const int half = warp >> 3;
    const int q = warp & 3;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const int arow = rb * kDT + q * 32 + lane;   // this thread's channel
    if (nch > 0) {
      mbar_wait(bar(2 * kDStages), 0);
      tc_fence_after();
    }
    if (MODE == 0) {
      const bool ok = arow < a.Cout;
      const float bias = (ok && a.bias != nullptr) ? a.bias[arow] : 0.f;
      double acc_s = 0.0, acc_ss = 0.0;
#pragma unroll 1
      for (int ch = half * 2; ch < half * 2 + 2; ++ch) {
        uint32_t r[32];
        cuda::ptx::tcgen05_ld_32x32b(r, tmem_base + lane_addr + (uint32_t)(ch * 32));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        const int pos0 = bt * kDT + ch * 32;
        if (ok) {
          float ts = 0.f, tss = 0.f;
          float *zp = a.z + (size_t)pos0 * a.ld_z + arow;
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            if (pos0 + i < a.M) {
              const float x = (nch > 0 ? __uint_as_float(r[i]) : 0.f) + bias;
              ts += x;
              tss = fmaf(x, x, tss);
              zp[(size_t)i * a.ld_z] = x;
            }
          }
          acc_s += (double)ts;
          acc_ss += (double)tss;
        }
      }
      if (ok && a.stats != nullptr) {
        atomicAdd(a.stats + arow, acc_s);
        atomicAdd(a.stats + a.Cout + arow, acc_ss);
      }
    } else if (MODE == 1) {
      const bool ok = arow < a.Cin;
      const bool masked = a.sc_in != nullptr;
      const float sc = (ok && masked) ? a.sc_in[arow] : 0.f, sh = (ok && masked) ? a.sh_in[arow] : 0.f;
      double d1 = 0.0, d2 = 0.0;
#pragma unroll 1
      for (int ch = half * 2; ch < half * 2 + 2; ++ch) {
        uint32_t r[32];
        cuda::ptx::tcgen05_ld_32x32b(r, tmem_base + lane_addr + (uint32_t)(ch * 32));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        const int pos0 = bt * kDT + ch * 32;
        if (ok) {
          float t1 = 0.f, t2 = 0.f;
          const float *ip = a.in + (size_t)pos0 * a.ld_in + arow;
          float *gp = a.gin + (size_t)pos0 * a.ld_gin + arow;
          if (masked) {
            float zv[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) zv[i] = pos0 + i < a.M ? __ldg(ip + (size_t)i * a.ld_in) : 0.f;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              if (pos0 + i < a.M) {
                const float gv = fmaf(zv[i], sc, sh) > 0.f ? __uint_as_float(r[i]) : 0.f;
                gp[(size_t)i * a.ld_gin] = gv;
                t1 += gv;
                t2 = fmaf(gv, zv[i], t2);
              }
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (pos0 + i < a.M) gp[(size_t)i * a.ld_gin] = __uint_as_float(r[i]);
          }
          d1 += (double)t1;
          d2 += (double)t2;
        }
      }
      if (ok && masked && a.stats_in != nullptr) {
        atomicAdd(a.stats_in + arow, d1);
        atomicAdd(a.stats_in + a.Cin + arow, d2);
      }
    } else {
      // partial dW tile -> dW (Cout, Cin): transposed through smem (stage 0 is free: every MMA has
      // completed) so that one atomic instruction covers 32 consecutive input channels of a row
      float *stg = reinterpret_cast<float *>(base) + (half * 4 + q) * (32 * 33);
      const int co0 = rb * kDT + q * 32;
#pragma unroll 1
      for (int ch = half * 2; ch < half * 2 + 2; ++ch) {
        uint32_t r[32];
        cuda::ptx::tcgen05_ld_32x32b(r, tmem_base + lane_addr + (uint32_t)(ch * 32));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int i = 0; i < 32; ++i) stg[lane * 33 + i] = nch > 0 ? __uint_as_float(r[i]) : 0.f;
        __syncwarp();
        const int k = bt * kDT + ch * 32 + lane;
        if (k < a.Cin) {
          const int rows = min(32, a.Cout - co0);
          for (int rr = 0; rr < rows; ++rr)
            atomicAdd(a.dW + (size_t)(co0 + rr) * a.Cin + k, stg[rr * 33 + lane]);
        }
        __syncwarp();
      }
    }
    tc_fence_before();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem_base, 128);
}


template <int MODE>
int launch_dense(const DenseArgs &a, dim3 grid, cudaStream_t st) {
  B2R_CUDA(cudaFuncSetAttribute(dense_gemm_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)dn_smem(MODE)));
  dense_gemm_kernel<MODE><<<grid, kDThreads, dn_smem(MODE), st>>>(a);
  B2R_CHECK_LAUNCH();
  return B2R_OK;
}

}  // namespace
}  // namespace b2r

using namespace b2r;

extern "C" long long b2r_dense_image_bytes(int rows, int k) {
  if (rows <= 0 || k <= 0) return 0;
  return (long long)((rows + kDT - 1) / kDT) * ((k + kDK - 1) / kDK) * kDOperand;
}

extern "C" int b2r_dense_pack(const float *w, int Cout, int Cin, float *w_img, float *wt_img,
                              void *stream) {
  B2R_REQUIRE(w && Cout > 0 && Cin > 0 && (w_img || wt_img), "b2r_dense_pack: bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (w_img != nullptr) {
    const long long n = b2r_dense_image_bytes(Cout, Cin) / 4;
    dense_pack_kernel<<<ceil_div(n, 256), 256, 0, st>>>(w, Cout, Cin, Cin, 0, w_img);
    B2R_CHECK_LAUNCH();
  }
  if (wt_img != nullptr) {
    const long long n = b2r_dense_image_bytes(Cin, Cout) / 4;
    dense_pack_kernel<<<ceil_div(n, 256), 256, 0, st>>>(w, Cin, Cout, Cin, 1, wt_img);
    B2R_CHECK_LAUNCH();
  }
  return B2R_OK;
}

namespace {
bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
}

extern "C" int b2r_dense_fwd(const b2r_dense_layer *d, void *stream) {
  B2R_REQUIRE(d != nullptr, "b2r_dense_fwd: null descriptor");
  B2R_REQUIRE(d->M > 0 && d->Cin > 0 && d->Cout > 0, "b2r_dense_fwd: non-positive size");
  B2R_REQUIRE(d->in && d->w_img && d->z, "b2r_dense_fwd: needs in, w_img, z");
  B2R_REQUIRE((d->sc_in == nullptr) == (d->sh_in == nullptr), "b2r_dense_fwd: scale and shift go together");
  B2R_REQUIRE((d->ld_in % 4) == 0 && aligned16(d->in) && aligned16(d->sc_in) && aligned16(d->sh_in),
              "b2r_dense_fwd: input rows / coefficient vectors must be 16-byte aligned");
  if (d->sc_in != nullptr && d->Cin > kDMaxCin) {
    set_error("b2r_dense_fwd: a BatchNorm+ReLU prologue supports Cin <= %d (got %d)", kDMaxCin, d->Cin);
    return B2R_ERR_UNSUPPORTED;
  }
  DenseArgs a = {};
  a.M = d->M; a.Cin = d->Cin; a.Cout = d->Cout;
  a.in = d->in; a.ld_in = d->ld_in; a.sc_in = d->sc_in; a.sh_in = d->sh_in;
  a.w_img = d->w_img; a.bias = d->bias; a.z = d->z; a.ld_z = d->ld_z; a.stats = d->stats;
  dim3 grid(ceil_div(d->M, kDT), ceil_div(d->Cout, kDT), 1);
  return launch_dense<0>(a, grid, static_cast<cudaStream_t>(stream));
}

extern "C" int b2r_dense_bwd(const b2r_dense_layer_bwd *d, void *stream) {
  B2R_REQUIRE(d != nullptr, "b2r_dense_bwd: null descriptor");
  B2R_REQUIRE(d->M > 0 && d->Cin > 0 && d->Cout > 0, "b2r_dense_bwd: non-positive size");
  B2R_REQUIRE(d->g != nullptr, "b2r_dense_bwd: needs g");
  B2R_REQUIRE(d->ca == nullptr || (d->cb && d->cc && d->zz),
              "b2r_dense_bwd: BatchNorm-backward form needs ca, cb, cc and z");
  B2R_REQUIRE((d->ld_g % 4) == 0 && aligned16(d->g) && aligned16(d->zz) && aligned16(d->ca) &&
                  aligned16(d->cb) && aligned16(d->cc),
              "b2r_dense_bwd: gradient rows / coefficient vectors must be 16-byte aligned");
  B2R_REQUIRE((d->sc_in == nullptr) == (d->sh_in == nullptr), "b2r_dense_bwd: scale and shift go together");
  if (d->ca != nullptr && d->Cout > kDMaxCoutBn) {
    set_error("b2r_dense_bwd: the BatchNorm-backward form supports Cout <= %d (got %d)", kDMaxCoutBn,
              d->Cout);
    return B2R_ERR_UNSUPPORTED;
  }
  DenseArgs a = {};
  a.M = d->M; a.Cin = d->Cin; a.Cout = d->Cout;
  a.in = d->in; a.ld_in = d->ld_in; a.sc_in = d->sc_in; a.sh_in = d->sh_in;
  a.g = d->g; a.zz = d->zz; a.ld_g = d->ld_g; a.ca = d->ca; a.cb = d->cb; a.cc = d->cc;
  a.wt_img = d->wt_img; a.gin = d->gin; a.ld_gin = d->ld_gin; a.stats_in = d->stats_in;
  a.dW = d->dW;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (d->gin != nullptr) {
    B2R_REQUIRE(d->wt_img != nullptr, "b2r_dense_bwd: the input gradient needs wt_img");
    B2R_REQUIRE(d->sc_in == nullptr || d->in != nullptr, "b2r_dense_bwd: the ReLU mask needs in");
    dim3 grid(ceil_div(d->M, kDT), ceil_div(d->Cin, kDT), 1);
    const int rc = launch_dense<1>(a, grid, st);
    if (rc != B2R_OK) return rc;
  }
  if (d->dW != nullptr) {
    B2R_REQUIRE(d->in != nullptr && (d->ld_in % 4) == 0 && aligned16(d->in) && aligned16(d->sc_in) &&
                    aligned16(d->sh_in),
                "b2r_dense_bwd: the weight gradient needs the (16-byte aligned) layer input");
    const int tiles = ceil_div(d->Cin, kDT) * ceil_div(d->Cout, kDT);
    int splits = kNumSMs / tiles;
    const int max_splits = (d->M + 63) / 64;
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    a.split_len = ((d->M + splits - 1) / splits + kDK - 1) / kDK * kDK;
    splits = (d->M + a.split_len - 1) / a.split_len;
    dim3 grid(ceil_div(d->Cin, kDT), ceil_div(d->Cout, kDT), splits);
    const int rc = launch_dense<2>(a, grid, st);
    if (rc != B2R_OK) return rc;
  }
  return B2R_OK;
}
