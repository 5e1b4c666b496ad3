// fps.cu -- furthest point sampling held across a thread-block cluster (sm_100a).
//
// Replaces furthest_point_sampling_kernel (reference _ext_src/src/sampling_gpu.cu:74-178), which
// runs ONE 512-thread CTA per scene and, per iteration, streams xyz + the running min-distance
// array (20 B/point) through L2 and reduces with a 9-round __syncthreads tree.
//
// Design (B200-first):
//   * One scene = one thread-block cluster of C CTAs (C in {1,2,4,8,16}; 16 is the non-portable
//     size, one cluster per GPC).  Every thread keeps P points (x,y,z,min-dist) in REGISTERS for
//     the whole kernel: after the initial load no global or L2 traffic remains on the critical
//     path.  Algorithmic HBM bytes are 12*N + 4*npoint per scene.
//   * Per iteration: P fused distance updates per thread, a 2-instruction warp arg-max
//     (redux.sync on a packed (distance, tie-rank) key), one shared-memory round per CTA, then ONE
//     DSMEM hop: warp 0 pushes the CTA's candidate (key + coordinates, 20 B) into every peer's
//     slot with st.async, which completes a transaction mbarrier in the destination CTA.  No
//     cluster-wide barrier.sync inside the loop (~380 cycles + L1 flush each, B300_MICROARCH.md).
//   * The winner's coordinates travel with the key, so the next iteration starts without a
//     dependent global load.
//
// Bit-exactness with the reference:
//   * distance = fma(dz,dz, fma(dx,dx, dy*dy)) with d* = p - p_old; validity test
//     !((double)|p|^2 <= 1e-3); running min via fminf; all spelled with _rn intrinsics.
//   * Ties: the reference's per-lane strided scan keeps the FIRST strictly larger value and its
//     shared-memory tree keeps the LOWER slot, so among equal distances the winner minimises
//     rank(k) = (bitrev_L(k mod bs), k div bs), bs = 2^L = b2r_ref_block_threads(N).
//     Here thread T owns points that all share k mod bs and visits them in ascending k, so a
//     strict '>' scan inside the thread reproduces the in-lane rule, and cross-thread reductions
//     maximise the 64-bit key (float_bits(dist) , ~rank) -- distances are >= +0 so their bit
//     patterns order like the floats.
#include "common.cuh"

namespace b2r {
namespace {

struct __align__(16) FpsRec {
  uint32_t hi, lo;  // key: distance bits, ~rank (0,0 = "no valid candidate")
  float x, y;
  float z;
  uint32_t pad0, pad1, pad2;
};
static_assert(sizeof(FpsRec) == 32, "FpsRec must be 32 bytes");

constexpr int kMaxCluster = 16;
constexpr uint32_t kMsgBytes = 20;  // v4.b32 {hi,lo,x,y} + b32 {z}

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::
                   : "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "B2R_WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra B2R_DONE_%=;\n\t"
      "bra B2R_WAIT_%=;\n\t"
      "B2R_DONE_%=:\n\t"
      "}" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void st_async_v4(uint32_t raddr, uint32_t a, uint32_t b, uint32_t c,
                                            uint32_t d, uint32_t rbar) {
  asm volatile(
      "st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1,%2,%3,%4}, [%5];" ::
          "r"(raddr),
      "r"(a), "r"(b), "r"(c), "r"(d), "r"(rbar)
      : "memory");
}
__device__ __forceinline__ void st_async_b32(uint32_t raddr, uint32_t a, uint32_t rbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::
                   "r"(raddr),
               "r"(a), "r"(rbar)
               : "memory");
}

// arg-max of the 64-bit key (hi,lo) over a full warp in two redux.sync instructions
__device__ __forceinline__ void warp_argmax(uint32_t hi, uint32_t lo, uint32_t &whi,
                                            uint32_t &wlo) {
  whi = __reduce_max_sync(0xffffffffu, hi);
  wlo = __reduce_max_sync(0xffffffffu, hi == whi ? lo : 0u);
}

// P points per thread in registers, NT threads per CTA, `csize` CTAs per scene.
// grid = (csize, B), cluster = (csize,1,1).  L = log2(b2r_ref_block_threads(N)).
template <int P, int NT>
__global__ void __launch_bounds__(NT, 1)
    fps_cluster_kernel(const float *__restrict__ xyz, int N, int npoint, int *__restrict__ idx,
                       int L, int csize) {
  extern __shared__ __align__(16) float s_pts[];  // [3][P*NT]: coordinates for winner lookup
  __shared__ FpsRec s_rec[2][32];                 // per-warp candidates, double-buffered
  __shared__ FpsRec s_slot[2][kMaxCluster];       // per-CTA candidates from the whole cluster
  __shared__ __align__(8) uint64_t s_bar[2];

  constexpr int NW = NT / 32;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t crank = blockIdx.x;  // cluster = (csize,1,1) and gridDim.x == csize
  const int scene = blockIdx.y;
  xyz += (size_t)scene * N * 3;
  idx += (size_t)scene * npoint;

  if (npoint <= 0) return;
  if (npoint == 1 || N <= 0) {  // uniform over the cluster: nobody touches a barrier
    if (crank == 0 && tid == 0) idx[0] = 0;
    return;
  }

  // ---- point ownership: thread T <-> reference lane t, sub-slice s ------------------------
  const int bs_mask = (1 << L) - 1;
  const int T = (int)crank * NT + tid;
  const int t = T & bs_mask;
  const int s = T >> L;
  const int S = (csize * NT) >> L;
  const uint32_t rank_base = (L ? (__brev((uint32_t)t) >> (32 - L)) : 0u) << 22;

  float px[P], py[P], pz[P], pt[P];
#pragma unroll
  for (int i = 0; i < P; ++i) {
    const long long k = (long long)t + ((long long)(s + S * i) << L);
    float x = __int_as_float(0x7fc00000), y = x, z = x, d = -1.0f;  // NaN coords never win
    if (k < N) {
      const float gx = xyz[k * 3 + 0], gy = xyz[k * 3 + 1], gz = xyz[k * 3 + 2];
      s_pts[(0 * P + i) * NT + tid] = gx;
      s_pts[(1 * P + i) * NT + tid] = gy;
      s_pts[(2 * P + i) * NT + tid] = gz;
      const float mag = sumsq_ref(gx, gy, gz);
      if (!((double)mag <= 1e-3)) {  // reference: `if (mag <= 1e-3) continue;` in double
        x = gx; y = gy; z = gz;
        d = 1e10f;  // reference scratch fill (sampling.cpp:78-80)
      }
    }
    px[i] = x; py[i] = y; pz[i] = z; pt[i] = d;
  }

  const uint32_t bar0 = smem_u32(&s_bar[0]);
  uint32_t r_slot0 = 0, r_slot1 = 0, r_bar0 = 0, r_bar1 = 0;
  if (csize > 1) {
    if (tid == 0) {
      mbar_init(bar0, 1);
      mbar_init(bar0 + 8, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      mbar_arrive_expect_tx(bar0, kMsgBytes * csize);
      mbar_arrive_expect_tx(bar0 + 8, kMsgBytes * csize);
    }
    if (warp == 0 && lane < csize) {
      r_slot0 = mapa(smem_u32(&s_slot[0][crank]), lane);
      r_slot1 = mapa(smem_u32(&s_slot[1][crank]), lane);
      r_bar0 = mapa(bar0, lane);
      r_bar1 = mapa(bar0 + 8, lane);
    }
    cluster_sync_all();  // every CTA's barriers are initialised before anyone sends
  }

  float ox = xyz[0], oy = xyz[1], oz = xyz[2];  // idx[0] = 0, valid or not (sampling_gpu.cu:92-93)
  if (crank == 0 && tid == 0) idx[0] = 0;

  for (int it = 0; it < npoint - 1; ++it) {
    const int par = it & 1;

    // ---- P register-resident distance updates + in-thread arg-max -------------------------
    float best = -1.0f;
    int bi = 0;
#pragma unroll
    for (int i = 0; i < P; ++i) {
      const float dx = __fsub_rn(px[i], ox), dy = __fsub_rn(py[i], oy), dz = __fsub_rn(pz[i], oz);
      const float d = sumsq_ref(dx, dy, dz);
      const float m = fminf(d, pt[i]);
      pt[i] = m;
      if (m > best) {  // strict: earliest (lowest-rank) point of this thread wins ties
        best = m;
        bi = i;
      }
    }
    const bool has = best >= 0.0f;
    const uint32_t hi = has ? __float_as_uint(best) : 0u;
    const uint32_t lo = has ? ~(rank_base | (uint32_t)(s + S * bi)) : 0u;

    uint32_t whi, wlo;
    warp_argmax(hi, lo, whi, wlo);
    const bool win = (wlo == 0u) ? (lane == 0) : (hi == whi && lo == wlo);
    if (win) {
      FpsRec r;
      r.hi = whi; r.lo = wlo;
      r.x = s_pts[(0 * P + bi) * NT + tid];
      r.y = s_pts[(1 * P + bi) * NT + tid];
      r.z = s_pts[(2 * P + bi) * NT + tid];
      r.pad0 = r.pad1 = r.pad2 = 0;
      *reinterpret_cast<uint4 *>(&s_rec[par][warp]) = *reinterpret_cast<uint4 *>(&r);
      s_rec[par][warp].z = r.z;
    }

    uint32_t khi, klo;  // the scene-wide winning key
    if (csize == 1) {
      __syncthreads();
      uint32_t h = 0, l = 0;
      if (lane < NW) { h = s_rec[par][lane].hi; l = s_rec[par][lane].lo; }
      warp_argmax(h, l, khi, klo);
      const int wl = __ffs(__ballot_sync(0xffffffffu, h == khi && l == klo)) - 1;
      ox = s_rec[par][wl].x; oy = s_rec[par][wl].y; oz = s_rec[par][wl].z;
    } else {
      if (warp == 0) {
        asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory");
        uint32_t h = 0, l = 0;
        if (lane < NW) { h = s_rec[par][lane].hi; l = s_rec[par][lane].lo; }
        uint32_t chi, clo;
        warp_argmax(h, l, chi, clo);
        const int wl = __ffs(__ballot_sync(0xffffffffu, h == chi && l == clo)) - 1;
        if (lane < csize) {  // one DSMEM hop: data + mbarrier completion in the peer CTA
          const float cx = s_rec[par][wl].x, cy = s_rec[par][wl].y, cz = s_rec[par][wl].z;
          const uint32_t dst = par ? r_slot1 : r_slot0, dbar = par ? r_bar1 : r_bar0;
          st_async_v4(dst, chi, clo, __float_as_uint(cx), __float_as_uint(cy), dbar);
          st_async_b32(dst + 16, __float_as_uint(cz), dbar);
        }
      } else {
        asm volatile("bar.arrive 1, %0;" ::"n"(NT) : "memory");
      }
      mbar_wait(bar0 + 8 * par, (uint32_t)(it >> 1) & 1u);
      uint32_t h = 0, l = 0;
      if (lane < csize) { h = s_slot[par][lane].hi; l = s_slot[par][lane].lo; }
      warp_argmax(h, l, khi, klo);
      const int wl = __ffs(__ballot_sync(0xffffffffu, h == khi && l == klo)) - 1;
      ox = s_slot[par][wl].x; oy = s_slot[par][wl].y; oz = s_slot[par][wl].z;
      // re-arm this barrier for iteration it+2 (safe: no peer can send for it+2 before it has
      // seen this CTA's it+1 message, which warp 0 issues after this point in program order)
      if (tid == 0) mbar_arrive_expect_tx(bar0 + 8 * par, kMsgBytes * csize);
    }

    if (crank == 0 && tid == 0) {
      int old = 0;
      if (klo != 0u) {
        const uint32_t rank = ~klo;
        const uint32_t tb = rank >> 22;
        const uint32_t tt = L ? (__brev(tb) >> (32 - L)) : 0u;
        old = (int)(tt + ((rank & 0x3fffffu) << L));
      }
      idx[it + 1] = old;
    }
  }

  if (csize > 1) cluster_sync_all();  // nobody exits while a peer may still address its smem
}

// ------------------------------------------------------------------------------ host side --
struct FpsPlan {
  int L, csize, NT, P, smem;
};

const int kPList[] = {1, 2, 3, 4, 5, 6, 7, 8, 10, 12, 14, 16, 20, 24};
constexpr int kPMax = 24;

int round_up_P(int p) {
  for (int v : kPList)
    if (v >= p) return v;
  return -1;
}

template <int P, int NT>
cudaError_t launch_inst(const FpsPlan &pl, const float *xyz, int B, int N, int npoint, int *idx,
                        cudaStream_t stream) {
  auto kern = fps_cluster_kernel<P, NT>;
  // Function attributes are PER DEVICE: one process may drive several GPUs (nn.DataParallel
  // replica threads, reference train_Votenet_FSB.py:164-168), so the "already set" flag is kept
  // per device ordinal.  Idempotent; a benign race at worst repeats the calls.
  static bool attr_done[64] = {};
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev < 0 || dev >= 64 || !attr_done[dev]) {
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             3 * P * NT * (int)sizeof(float));
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < 64) attr_done[dev] = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(pl.csize, B, 1);
  cfg.blockDim = dim3(NT, 1, 1);
  cfg.dynamicSmemBytes = pl.smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = pl.csize;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, xyz, N, npoint, idx, pl.L, pl.csize);
}

template <int NT>
cudaError_t launch_nt(const FpsPlan &pl, const float *xyz, int B, int N, int npoint, int *idx,
                      cudaStream_t st) {
  if constexpr (NT < 512) {
    switch (pl.P) {
      case 1: return launch_inst<1, NT>(pl, xyz, B, N, npoint, idx, st);
      case 2: return launch_inst<2, NT>(pl, xyz, B, N, npoint, idx, st);
      default: return cudaErrorInvalidValue;
    }
  } else {
    switch (pl.P) {
      case 1: return launch_inst<1, NT>(pl, xyz, B, N, npoint, idx, st);
      case 2: return launch_inst<2, NT>(pl, xyz, B, N, npoint, idx, st);
      case 3: return launch_inst<3, NT>(pl, xyz, B, N, npoint, idx, st);
      case 4: return launch_inst<4, NT>(pl, xyz, B, N, npoint, idx, st);
      case 5: return launch_inst<5, NT>(pl, xyz, B, N, npoint, idx, st);
      case 6: return launch_inst<6, NT>(pl, xyz, B, N, npoint, idx, st);
      case 7: return launch_inst<7, NT>(pl, xyz, B, N, npoint, idx, st);
      case 8: return launch_inst<8, NT>(pl, xyz, B, N, npoint, idx, st);
      case 10: return launch_inst<10, NT>(pl, xyz, B, N, npoint, idx, st);
      case 12: return launch_inst<12, NT>(pl, xyz, B, N, npoint, idx, st);
      case 14: return launch_inst<14, NT>(pl, xyz, B, N, npoint, idx, st);
      case 16: return launch_inst<16, NT>(pl, xyz, B, N, npoint, idx, st);
      case 20: return launch_inst<20, NT>(pl, xyz, B, N, npoint, idx, st);
      case 24: return launch_inst<24, NT>(pl, xyz, B, N, npoint, idx, st);
      default: return cudaErrorInvalidValue;
    }
  }
}

int env_int(const char *name, int dflt) {
  const char *v = getenv(name);
  return (v && *v) ? atoi(v) : dflt;
}

// Largest cluster size make_plan may pick.  Starts at 16 (non-portable, one cluster per GPC) or
// B2R_FPS_MAX_CLUSTER; b2r_fps lowers it to 8 if the driver refuses a 16-CTA cluster launch.
int g_max_cluster = []() {
  const int v = env_int("B2R_FPS_MAX_CLUSTER", kMaxCluster);
  return (v == 1 || v == 2 || v == 4 || v == 8 || v == 16) ? v : kMaxCluster;
}();

// Chooses cluster size / threads / points-per-thread for (B, N).  Returns false if N exceeds what
// 16 CTAs x 512 threads x 24 register-resident points can hold.
bool make_plan(int B, int N, FpsPlan *pl, int cluster_hint = 0) {
  const int bs = b2r_ref_block_threads(N > 0 ? N : 1);
  int L = 0;
  while ((1 << L) < bs) ++L;
  pl->L = L;
  if (N < 512) {
    pl->csize = 1;
    pl->NT = bs < 32 ? 32 : bs;
    const int rows = (N + bs - 1) / bs;           // 1 or 2
    const int S = pl->NT >> L;                    // >= 1
    pl->P = round_up_P((rows + S - 1) / S);
  } else {
    const int rows = (N + 511) / 512;
    const int max_c = g_max_cluster;
    int c;
    if (rows <= 8) {
      c = 1;  // <= 4096 points: one CTA, no DSMEM hop
    } else {
      // largest cluster that still leaves every scene of the batch resident at once
      c = 1;
      while (c * 2 <= max_c && (long long)B * (c * 2) <= kNumSMs && rows / (c * 2) >= 2) c *= 2;
      while (c < max_c && (rows + c - 1) / c > kPMax) c *= 2;  // capacity wins over residency
      // Measured on B200 (scripts/fps_sweep.py, profiles/r01/fps_cluster_sweep.log), ns per
      // dependent iteration at N = 40000 / 50000:  c=8: 739 / 823,  c=10: 707 / 746,
      // c=12: 949 / 1500,  c=16: 853 / 944.  The per-iteration cost is the DSMEM fan-out (one
      // st.async pair per peer) plus the mbarrier round trip, not the register-resident distance
      // updates, so fewer, fatter CTAs win until the per-thread work (> 10 points) takes over;
      // 10 CTAs x 512 threads is the sweet spot (cluster sizes need not be powers of two: thread
      // ownership only needs c*512 to be a multiple of the reference's 512-lane stride).
      if (c >= 8 && max_c >= 10 && (long long)B * 10 <= kNumSMs && (rows + 9) / 10 <= 10) c = 10;
      else if (c == 16 && (rows + 7) / 8 <= 10) c = 8;
    }
    // any size: ownership only needs c*512 % 2^L == 0.  A caller's hint (b2r_fps_ex) narrows the
    // cluster when FPS runs beside other kernels and latency matters less than the SMs it holds.
    int forced = env_int("B2R_FPS_CLUSTER", 0);
    if (cluster_hint >= 1 && cluster_hint <= max_c && rows > 8) {
      forced = cluster_hint;
      while (forced < max_c && (rows + forced - 1) / forced > kPMax) ++forced;  // capacity first
    }
    if (forced >= 1 && forced <= 16) c = forced;
    pl->csize = c;
    pl->NT = 512;
    const int p = (rows + c - 1) / c;
    if (p > kPMax) return false;
    pl->P = round_up_P(p);
  }
  pl->smem = 3 * pl->P * pl->NT * (int)sizeof(float);
  return pl->P > 0;
}

}  // namespace
}  // namespace b2r

extern "C" int b2r_fps_plan(int B, int N, int *cluster_size, int *threads, int *points_per_thread,
                            int *smem_bytes) {
  B2R_REQUIRE(B >= 0 && N >= 0, "b2r_fps_plan: negative size (B=%d N=%d)", B, N);
  b2r::FpsPlan pl;
  if (!b2r::make_plan(B, N, &pl)) {
    b2r::set_error("b2r_fps: N=%d exceeds the register-resident capacity (%d points)", N,
                   b2r::kMaxCluster * 512 * b2r::kPMax);
    return B2R_ERR_UNSUPPORTED;
  }
  if (cluster_size) *cluster_size = pl.csize;
  if (threads) *threads = pl.NT;
  if (points_per_thread) *points_per_thread = pl.P;
  if (smem_bytes) *smem_bytes = pl.smem;
  return B2R_OK;
}

extern "C" int b2r_fps(const float *xyz, int B, int N, int npoint, int *idx, void *stream) {
  return b2r_fps_ex(xyz, B, N, npoint, idx, 0, stream);
}

extern "C" int b2r_fps_ex(const float *xyz, int B, int N, int npoint, int *idx, int cluster_hint,
                          void *stream) {
  B2R_REQUIRE(B >= 0 && N >= 0 && npoint >= 0, "b2r_fps: negative size (B=%d N=%d npoint=%d)", B, N,
              npoint);
  B2R_REQUIRE(cluster_hint >= 0 && cluster_hint <= 16, "b2r_fps_ex: cluster_hint=%d not in [0,16]",
              cluster_hint);
  if (B == 0 || npoint == 0) return B2R_OK;
  B2R_REQUIRE(xyz != nullptr && idx != nullptr, "b2r_fps: null pointer");
  B2R_REQUIRE(B <= 65535, "b2r_fps: B=%d exceeds gridDim.y", B);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (N == 0) {  // nothing to sample from: the reference would emit index 0 everywhere
    B2R_CUDA(cudaMemsetAsync(idx, 0, sizeof(int) * (size_t)B * npoint, st));
    return B2R_OK;
  }
  b2r::FpsPlan pl;
  if (!b2r::make_plan(B, N, &pl, cluster_hint)) {
    b2r::set_error("b2r_fps: N=%d exceeds the register-resident capacity (%d points)", N,
                   b2r::kMaxCluster * 512 * b2r::kPMax);
    return B2R_ERR_UNSUPPORTED;
  }
  cudaError_t e;
  for (int attempt = 0; attempt < 2; ++attempt) {
    switch (pl.NT) {
      case 32: e = b2r::launch_nt<32>(pl, xyz, B, N, npoint, idx, st); break;
      case 64: e = b2r::launch_nt<64>(pl, xyz, B, N, npoint, idx, st); break;
      case 128: e = b2r::launch_nt<128>(pl, xyz, B, N, npoint, idx, st); break;
      case 256: e = b2r::launch_nt<256>(pl, xyz, B, N, npoint, idx, st); break;
      default: e = b2r::launch_nt<512>(pl, xyz, B, N, npoint, idx, st); break;
    }
    if (e == cudaSuccess || pl.csize <= 8) break;
    // a 16-CTA (non-portable) cluster was refused: fall back to the portable maximum, once
    (void)cudaGetLastError();
    b2r::g_max_cluster = 8;
    if (!b2r::make_plan(B, N, &pl, cluster_hint)) {
      b2r::set_error("b2r_fps: N=%d needs a 16-CTA cluster, which this device refused", N);
      return B2R_ERR_UNSUPPORTED;
    }
  }
  if (e != cudaSuccess) {
    b2r::set_error("b2r_fps launch (cluster=%d threads=%d P=%d smem=%d) failed: %s", pl.csize,
                   pl.NT, pl.P, pl.smem, cudaGetErrorString(e));
    return B2R_ERR_CUDA;
  }
  return B2R_OK;
}
