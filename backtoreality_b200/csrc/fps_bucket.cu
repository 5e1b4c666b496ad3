// fps_bucket.cu -- furthest point sampling over spatially sorted buckets, one scene per
// thread-block cluster, candidates exchanged warp -> every CTA in ONE DSMEM hop (sm_100a).
//
// Replaces furthest_point_sampling_kernel (reference _ext_src/src/sampling_gpu.cu:74-178) like
// fps.cu does, with the same bit-exact result, but attacks the two things
// that bound fps.cu's dependent iteration (profiles/r02/fps_phase_probe.txt, lat_probe_sm100a.txt):
//
//   * work: after j samples a new sample can only lower the running min-distance of points closer
//     to it than their current value, i.e. of a ~1/j fraction of the scene.  Points are therefore
//     sorted by a Morton cell code first (fps_sort_kernel, one counting sort per scene in shared
//     memory), every WARP owns a contiguous run of 32*P sorted points -- a spatially compact
//     bucket -- and keeps its bounding box.  An iteration starts with a warp-uniform test: if the
//     squared distance from the new sample to the box (computed with the reference's own rounding
//     sequence, which is monotone, so it is a true lower bound of every d the reference would
//     compute) is >= the bucket's largest running min-distance, fminf(d, temp) changes nothing and
//     the warp keeps last iteration's candidate.  On ScanNet-shaped rooms ~8 of 63 buckets are
//     touched per iteration (scripts/fps_bucket_sim.py).  The result is IDENTICAL: only updates
//     that cannot change any value are skipped.
//   * latency: a touched warp is usually the only busy warp of its scheduler, so its update loop
//     is written for instruction-level parallelism (four points side by side, a compare tree per
//     group); untouched warps only carry their candidate record over.  The exchange is fps.cu's:
//     per-warp records in shared memory, one named-barrier round, ONE st.async push of the CTA's
//     candidate (key + coordinates, 20 B) into every peer completing a transaction mbarrier, no
//     cluster barrier in the loop.  The winner's table slot rides in the low bits of the key, so
//     no ballot is needed to find whose coordinates to read.
//     (Measured and dropped, profiles/r02/fps_table_protocol.txt: every warp pushing its record
//     straight into a table replicated in all CTAs -- one level less, but 32x the st.async
//     completions per mbarrier, which serialise at ~6 cycles each in the receiving SM.)
//
// Tie rule: key = (float bits of the min-distance, lo) maximised, lo = 1<<31 | (~rank << 8) |
// table slot, rank(k) = (bitrev_L(k mod bs), k div bs) exactly as in fps.cu; every thread keeps
// its P points in ascending rank so a strict '>' scan is the in-lane rule.
#include <stdlib.h>

#include "common.cuh"

namespace b2r {
namespace {

[[maybe_unused]] constexpr int kTable = 256;  // max warps (buckets) per scene: 16 CTAs x 16 warps
constexpr int kMaxCta = 16;

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
      "B2R_BWAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra B2R_BDONE_%=;\n\t"
      "bra B2R_BWAIT_%=;\n\t"
      "B2R_BDONE_%=:\n\t"
      "}" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void st_async_b32(uint32_t raddr, uint32_t a, uint32_t rbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::
                   "r"(raddr),
               "r"(a), "r"(rbar)
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

// ------------------------------------------------------------------ Morton-cell counting sort --
__device__ __forceinline__ uint32_t spread3(uint32_t v) {  // 10 bits -> every third bit
  v = (v | (v << 16)) & 0x030000FFu;
  v = (v | (v << 8)) & 0x0300F00Fu;
  v = (v | (v << 4)) & 0x030C30C3u;
  v = (v | (v << 2)) & 0x09249249u;
  return v;
}
__device__ __forceinline__ bool finite3(float x, float y, float z) {
  return isfinite(x) && isfinite(y) && isfinite(z);
}

// One CTA per scene: bounding box -> cell histogram -> exclusive scan -> fill, all in shared
// memory.  perm[b][pos] = original index of the point at sorted position pos.  The order inside a
// cell is whatever the atomics produce; FPS's result does not depend on it (ranks travel with
// the points).  bits per axis: 3..5 (512 .. 32768 cells).
__global__ void __launch_bounds__(1024, 1)
    fps_sort_kernel(const float *__restrict__ xyz, int N, int bits, int *__restrict__ perm) {
  extern __shared__ int s_cnt[];
  __shared__ float s_red[6][32];
  __shared__ int s_wsum[32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  xyz += (size_t)blockIdx.x * N * 3;
  perm += (size_t)blockIdx.x * N;
  const int ncell = 1 << (3 * bits);

  float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
  for (int k = tid; k < N; k += 1024) {
    const float x = xyz[k * 3 + 0], y = xyz[k * 3 + 1], z = xyz[k * 3 + 2];
    if (finite3(x, y, z)) {
      lo[0] = fminf(lo[0], x); hi[0] = fmaxf(hi[0], x);
      lo[1] = fminf(lo[1], y); hi[1] = fmaxf(hi[1], y);
      lo[2] = fminf(lo[2], z); hi[2] = fmaxf(hi[2], z);
    }
  }
#pragma unroll
  for (int d = 0; d < 3; ++d) {
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      lo[d] = fminf(lo[d], __shfl_xor_sync(0xffffffffu, lo[d], o));
      hi[d] = fmaxf(hi[d], __shfl_xor_sync(0xffffffffu, hi[d], o));
    }
    if (lane == 0) { s_red[d][warp] = lo[d]; s_red[3 + d][warp] = hi[d]; }
  }
  for (int i = tid; i < ncell; i += 1024) s_cnt[i] = 0;
  __syncthreads();
  float scale[3];
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    float l = s_red[d][lane], h = s_red[3 + d][lane];
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      l = fminf(l, __shfl_xor_sync(0xffffffffu, l, o));
      h = fmaxf(h, __shfl_xor_sync(0xffffffffu, h, o));
    }
    lo[d] = l;
    scale[d] = (h > l) ? (float)(1 << bits) / (h - l) : 0.0f;
  }
  const int qmax = (1 << bits) - 1;
  auto cell_of = [&](int k) -> int {
    const float x = xyz[k * 3 + 0], y = xyz[k * 3 + 1], z = xyz[k * 3 + 2];
    if (!finite3(x, y, z)) return 0;
    const int qx = min(max((int)((x - lo[0]) * scale[0]), 0), qmax);
    const int qy = min(max((int)((y - lo[1]) * scale[1]), 0), qmax);
    const int qz = min(max((int)((z - lo[2]) * scale[2]), 0), qmax);
    return (int)(spread3((uint32_t)qx) | (spread3((uint32_t)qy) << 1) | (spread3((uint32_t)qz) << 2));
  };
  for (int k = tid; k < N; k += 1024) atomicAdd(&s_cnt[cell_of(k)], 1);
  __syncthreads();

  // exclusive scan of s_cnt[0..ncell): each warp owns a contiguous segment, rows of 32
  const int seg = ncell >= 1024 ? ncell / 32 : 32;
  const int used = ncell / seg;
  int total = 0;
  if (warp < used) {
    for (int r = 0; r < seg; r += 32) {
      const int i = warp * seg + r + lane;
      const int v = s_cnt[i];
      int inc = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int n = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += n;
      }
      s_cnt[i] = inc - v + total;
      total += __shfl_sync(0xffffffffu, inc, 31);
    }
    if (lane == 0) s_wsum[warp] = total;
  }
  __syncthreads();
  if (warp == 0) {
    const int v = lane < used ? s_wsum[lane] : 0;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int n = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += n;
    }
    s_wsum[lane] = inc - v;
  }
  __syncthreads();
  if (warp < used) {
    const int off = s_wsum[warp];
    for (int r = 0; r < seg; r += 32) s_cnt[warp * seg + r + lane] += off;
  }
  __syncthreads();
  for (int k = tid; k < N; k += 1024) perm[atomicAdd(&s_cnt[cell_of(k)], 1)] = k;
}

// ------------------------------------------------------------------------------ FPS kernel --
#ifdef B2R_FPS_PROFILE   // scripts/probe/fps_phase_probe.cu: per-warp cycle sums of the loop's phases
__device__ unsigned long long g_prof[kTable][8];
#define PROF_T(v) const long long v = clock64()
#else
#define PROF_T(v)
#endif

struct __align__(16) Rec {
  uint32_t hi, lo;  // key: distance bits, 1<<31 | ~rank << 8 | table slot  (0,0 = no candidate)
  float x, y;
  float z;
  uint32_t pad0, pad1, pad2;
};
static_assert(sizeof(Rec) == 32, "Rec must be 32 bytes");
constexpr uint32_t kMsgBytes = 20;  // v4.b32 {hi,lo,x,y} + b32 {z}

__device__ __forceinline__ void warp_argmax(uint32_t hi, uint32_t lo, uint32_t &whi,
                                            uint32_t &wlo) {
  whi = __reduce_max_sync(0xffffffffu, hi);
  wlo = __reduce_max_sync(0xffffffffu, hi == whi ? lo : 0u);
}

// P points per thread in registers, NT threads per CTA, `csize` CTAs per scene; grid =
// (csize, B), cluster = (csize,1,1).  L = log2(b2r_ref_block_threads(N)), SB = bits of k >> L.
template <int P, int NT>
__global__ void __launch_bounds__(NT, 1)
    fps_bucket_kernel(const float *__restrict__ xyz, const int *__restrict__ perm, int N, int npoint,
                      int *__restrict__ idx, int L, int SB, int csize) {
  extern __shared__ __align__(16) unsigned char s_dyn[];
  float *s_pts = reinterpret_cast<float *>(s_dyn);                   // [3][P*NT]
  uint32_t *s_lo = reinterpret_cast<uint32_t *>(s_pts + 3 * P * NT);  // [P*NT]
  __shared__ Rec s_rec[2][32];         // per-warp candidates, double-buffered
  __shared__ Rec s_slot[2][kMaxCta];   // per-CTA candidates from the whole cluster
  __shared__ __align__(8) uint64_t s_bar[2];

  constexpr int NW = NT / 32;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t crank = blockIdx.x;  // cluster = (csize,1,1) and gridDim.x == csize
  const int scene = blockIdx.y;
  xyz += (size_t)scene * N * 3;
  perm += (size_t)scene * N;
  idx += (size_t)scene * npoint;

  if (npoint <= 0) return;
  if (npoint == 1 || N <= 0) {  // uniform over the cluster: nobody touches a barrier
    if (crank == 0 && tid == 0) idx[0] = 0;
    return;
  }

  const int gw = (int)crank * NW + warp;  // this warp's bucket; also its slot id in the key
  const uint32_t rmask = (1u << (L + SB)) - 1u;

  // ---- load this thread's P points (sorted positions gw*32P + i*32 + lane) into smem ---------
  for (int i = 0; i < P; ++i) {
    const long long pos = (long long)gw * (32 * P) + i * 32 + lane;
    float x = __int_as_float(0x7fc00000), y = x, z = x;
    uint32_t lo = 0u;
    if (pos < N) {
      const int k = perm[pos];
      const float gx = xyz[k * 3 + 0], gy = xyz[k * 3 + 1], gz = xyz[k * 3 + 2];
      const float mag = sumsq_ref(gx, gy, gz);
      if (!((double)mag <= 1e-3)) {  // reference: `if (mag <= 1e-3) continue;` in double
        x = gx; y = gy; z = gz;
        const uint32_t t = (uint32_t)k & ((1u << L) - 1u);
        const uint32_t rank = ((L ? (__brev(t) >> (32 - L)) : 0u) << SB) | ((uint32_t)k >> L);
        lo = 0x80000000u | (((~rank) & rmask) << 8) | (uint32_t)gw;
      }
    }
    s_pts[(0 * P + i) * NT + tid] = x;
    s_pts[(1 * P + i) * NT + tid] = y;
    s_pts[(2 * P + i) * NT + tid] = z;
    s_lo[i * NT + tid] = lo;
  }
  // insertion sort of the thread's own slots by descending lo (= ascending rank, invalid last):
  // "first maximum wins" below is then the reference's in-lane rule (lowest rank among equals)
  for (int i = 1; i < P; ++i) {
    const uint32_t lo = s_lo[i * NT + tid];
    const float x = s_pts[(0 * P + i) * NT + tid], y = s_pts[(1 * P + i) * NT + tid],
                z = s_pts[(2 * P + i) * NT + tid];
    int j = i - 1;
    while (j >= 0 && s_lo[j * NT + tid] < lo) {
      s_lo[(j + 1) * NT + tid] = s_lo[j * NT + tid];
      s_pts[(0 * P + j + 1) * NT + tid] = s_pts[(0 * P + j) * NT + tid];
      s_pts[(1 * P + j + 1) * NT + tid] = s_pts[(1 * P + j) * NT + tid];
      s_pts[(2 * P + j + 1) * NT + tid] = s_pts[(2 * P + j) * NT + tid];
      --j;
    }
    s_lo[(j + 1) * NT + tid] = lo;
    s_pts[(0 * P + j + 1) * NT + tid] = x;
    s_pts[(1 * P + j + 1) * NT + tid] = y;
    s_pts[(2 * P + j + 1) * NT + tid] = z;
  }

  float px[P], py[P], pz[P], pt[P];
  float blx = INFINITY, bly = INFINITY, blz = INFINITY, bhx = -INFINITY, bhy = -INFINITY,
        bhz = -INFINITY;
#pragma unroll
  for (int i = 0; i < P; ++i) {
    px[i] = s_pts[(0 * P + i) * NT + tid];
    py[i] = s_pts[(1 * P + i) * NT + tid];
    pz[i] = s_pts[(2 * P + i) * NT + tid];
    const bool valid = s_lo[i * NT + tid] != 0u;
    pt[i] = valid ? 1e10f : -1.0f;  // reference scratch fill (sampling.cpp:78-80)
    if (valid) {                    // fminf/fmaxf drop NaN coordinates (their d is NaN: no update)
      blx = fminf(blx, px[i]); bhx = fmaxf(bhx, px[i]);
      bly = fminf(bly, py[i]); bhy = fmaxf(bhy, py[i]);
      blz = fminf(blz, pz[i]); bhz = fmaxf(bhz, pz[i]);
    }
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    blx = fminf(blx, __shfl_xor_sync(0xffffffffu, blx, o));
    bly = fminf(bly, __shfl_xor_sync(0xffffffffu, bly, o));
    blz = fminf(blz, __shfl_xor_sync(0xffffffffu, blz, o));
    bhx = fmaxf(bhx, __shfl_xor_sync(0xffffffffu, bhx, o));
    bhy = fmaxf(bhy, __shfl_xor_sync(0xffffffffu, bhy, o));
    bhz = fmaxf(bhz, __shfl_xor_sync(0xffffffffu, bhz, o));
  }

  const uint32_t bar0 = smem_u32(&s_bar[0]);
  uint32_t r_slot0 = 0, r_slot1 = 0, r_bar0 = 0, r_bar1 = 0;
  if (lane == 0) {
    Rec z = {0u, 0u, 0.f, 0.f, 0.f, 0u, 0u, 0u};
    s_rec[0][warp] = z;
    s_rec[1][warp] = z;
  }
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
  } else {
    __syncthreads();
  }

  float ox = xyz[0], oy = xyz[1], oz = xyz[2];  // idx[0] = 0, valid or not (sampling_gpu.cu:92-93)
  if (crank == 0 && tid == 0) idx[0] = 0;
  uint32_t ck_hi = 0u, ck_lo = 0u;  // this warp's candidate (uniform over the warp)
#ifdef B2R_FPS_PROFILE
  unsigned pacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#endif

  for (int it = 0; it < npoint - 1; ++it) {
    const int par = it & 1;
    PROF_T(t0);
#ifdef B2R_FPS_PROFILE
    bool touched = false;
#endif

    // ---- can the new sample lower any running min-distance of this bucket? ------------------
    // g* <= |fl(p - o)| for every p in the box (rounding is monotone), so lb <= every d below
    const float gx = fmaxf(fmaxf(__fsub_rn(blx, ox), __fsub_rn(ox, bhx)), 0.0f);
    const float gy = fmaxf(fmaxf(__fsub_rn(bly, oy), __fsub_rn(oy, bhy)), 0.0f);
    const float gz = fmaxf(fmaxf(__fsub_rn(blz, oz), __fsub_rn(oz, bhz)), 0.0f);
    const float lb = sumsq_ref(gx, gy, gz);
    const float wmax = ck_lo ? __uint_as_float(ck_hi) : -1.0f;
    if (it == 0 || lb < wmax) {
#ifdef B2R_FPS_PROFILE
      touched = true;
#endif
      // A touched warp is usually the only busy warp of its scheduler: nothing hides ALU latency
      // but its own instruction-level parallelism.  Groups of four points are updated side by
      // side and reduced by a small tree ("first maximum wins" at every node: the left operand
      // has the lower slot = lower rank); only one compare-select per group is a serial chain.
      float best = -1.0f;
      int bi = 0;
#pragma unroll
      for (int g = 0; g < P; g += 4) {
        float m[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (g + u < P) {
            const float dx = __fsub_rn(px[g + u], ox), dy = __fsub_rn(py[g + u], oy),
                        dz = __fsub_rn(pz[g + u], oz);
            m[u] = fminf(sumsq_ref(dx, dy, dz), pt[g + u]);
            pt[g + u] = m[u];
          } else {
            m[u] = -2.0f;
          }
        }
        const bool s01 = m[1] > m[0], s23 = m[3] > m[2];
        const float v01 = s01 ? m[1] : m[0], v23 = s23 ? m[3] : m[2];
        const int i01 = s01 ? g + 1 : g, i23 = s23 ? g + 3 : g + 2;
        const bool sg = v23 > v01;
        const float vg = sg ? v23 : v01;
        const int ig = sg ? i23 : i01;
        if (vg > best) {
          best = vg;
          bi = ig;
        }
      }
      const bool has = best >= 0.0f;
      const uint32_t hi = has ? __float_as_uint(best) : 0u;
      const uint32_t lo = has ? s_lo[bi * NT + tid] : 0u;
      warp_argmax(hi, lo, ck_hi, ck_lo);
      if (ck_lo != 0u && hi == ck_hi && lo == ck_lo) {
        Rec r;
        r.hi = ck_hi; r.lo = ck_lo;
        r.x = s_pts[(0 * P + bi) * NT + tid];
        r.y = s_pts[(1 * P + bi) * NT + tid];
        r.z = s_pts[(2 * P + bi) * NT + tid];
        r.pad0 = r.pad1 = r.pad2 = 0;
        *reinterpret_cast<uint4 *>(&s_rec[par][warp]) = *reinterpret_cast<uint4 *>(&r);
        s_rec[par][warp].z = r.z;
      }
    } else if (lane < 2) {  // untouched: last iteration's candidate stands, carried to this buffer
      reinterpret_cast<uint4 *>(&s_rec[par][warp])[lane] =
          reinterpret_cast<const uint4 *>(&s_rec[par ^ 1][warp])[lane];
    }
    __syncwarp();  // the record is read by other lanes of this warp next iteration (carry above)
    PROF_T(t1);

    uint32_t khi, klo;  // the scene-wide winning key
    if (csize == 1) {
      __syncthreads();
      uint32_t h = 0, l = 0;
      if (lane < NW) { h = s_rec[par][lane].hi; l = s_rec[par][lane].lo; }
      warp_argmax(h, l, khi, klo);
      if (klo != 0u) {
        const Rec *w = &s_rec[par][klo & 0xffu];
        ox = w->x; oy = w->y; oz = w->z;
      }
    } else {
      if (warp == 0) {
        asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory");
        uint32_t h = 0, l = 0;
        if (lane < NW) { h = s_rec[par][lane].hi; l = s_rec[par][lane].lo; }
        uint32_t chi, clo;
        warp_argmax(h, l, chi, clo);
        if (lane < csize) {  // one DSMEM hop: data + mbarrier completion in the peer CTA
          const Rec *w = &s_rec[par][clo ? (int)(clo & 0xffu) - (int)crank * NW : 0];
          const float cx = w->x, cy = w->y, cz = w->z;
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
      if (klo != 0u) {
        const Rec *w = &s_slot[par][(klo & 0xffu) / NW];
        ox = w->x; oy = w->y; oz = w->z;
      }
      // re-arm this barrier for iteration it+2 (safe: no peer can send for it+2 before it has
      // seen this CTA's it+1 message, which warp 0 issues after this point in program order)
      if (tid == 0) mbar_arrive_expect_tx(bar0 + 8 * par, kMsgBytes * csize);
    }

    if (crank == 0 && tid == 0) {
      int old = 0;
      if (klo != 0u) {
        const uint32_t rank = (~(klo >> 8)) & rmask;
        const uint32_t tb = rank >> SB;
        const uint32_t tt = L ? (__brev(tb) >> (32 - L)) : 0u;
        old = (int)(tt + ((rank & ((1u << SB) - 1u)) << L));
      }
      idx[it + 1] = old;
    }
#ifdef B2R_FPS_PROFILE
    {
      const long long t4 = clock64();
      pacc[touched ? 0 : 4] += 1;
      pacc[touched ? 1 : 5] += (unsigned)(t1 - t0);   // test (+ update + warp argmax + record)
      pacc[touched ? 3 : 6] += (unsigned)(t4 - t1);   // CTA round + cluster exchange + final reduce
    }
#endif
  }
#ifdef B2R_FPS_PROFILE
  if (lane == 0 && scene == 0)
    for (int q = 0; q < 8; ++q) g_prof[gw][q] = pacc[q];
#endif

  if (csize > 1) cluster_sync_all();  // nobody exits while a peer may still address its smem
}

// ------------------------------------------------------------------ one CTA, few fat warps --
// Scenes of <= 4096 points (the later SA levels: 2048 -> 1024 -> 512 -> 256, vote aggregation):
// one CTA holds the scene and the iteration is a pure latency chain.  fps.cu runs it on 16 warps
// (update, 2 redux, record, bar.sync over 16 warps, 16-record redux round, ballot, lookup: ~775
// cycles).  Here 4 (8 above 2048 points) warps own 4x the points each -- the update loop is
// ILP-scheduled, so a fatter thread costs little -- the barrier has 4 arrivals, and every thread
// reduces the 4-8 warp records itself with plain compares (no second redux round, no ballot; the
// winner's record slot rides in the key).  Points are not sorted and no bucket is skipped: with
// 4 buckets per scene there is nothing to skip.
template <int P, int NT>
__global__ void __launch_bounds__(NT, 1)
    fps_small_kernel(const float *__restrict__ xyz, int N, int npoint, int *__restrict__ idx, int L,
                     int SB) {
  extern __shared__ __align__(16) unsigned char s_dyn[];
  float(*s_pts)[P * NT] = reinterpret_cast<float(*)[P * NT]>(s_dyn);        // [3][P*NT]
  uint32_t *s_lo = reinterpret_cast<uint32_t *>(s_dyn) + 3 * P * NT;        // [P*NT]
  __shared__ Rec s_rec[2][NT / 32];
  constexpr int NW = NT / 32;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  xyz += (size_t)blockIdx.x * N * 3;
  idx += (size_t)blockIdx.x * npoint;
  if (npoint <= 0) return;
  if (npoint == 1 || N <= 0) {
    if (tid == 0) idx[0] = 0;
    return;
  }
  const uint32_t rmask = (1u << (L + SB)) - 1u;
  for (int i = 0; i < P; ++i) {
    const int k = tid + i * NT;
    float x = __int_as_float(0x7fc00000), y = x, z = x;
    uint32_t lo = 0u;
    if (k < N) {
      const float gx = xyz[k * 3 + 0], gy = xyz[k * 3 + 1], gz = xyz[k * 3 + 2];
      const float mag = sumsq_ref(gx, gy, gz);
      if (!((double)mag <= 1e-3)) {  // reference: `if (mag <= 1e-3) continue;` in double
        x = gx; y = gy; z = gz;
        const uint32_t t = (uint32_t)k & ((1u << L) - 1u);
        const uint32_t rank = ((L ? (__brev(t) >> (32 - L)) : 0u) << SB) | ((uint32_t)k >> L);
        lo = 0x80000000u | (((~rank) & rmask) << 8) | (uint32_t)warp;
      }
    }
    s_pts[0][i * NT + tid] = x;
    s_pts[1][i * NT + tid] = y;
    s_pts[2][i * NT + tid] = z;
    s_lo[i * NT + tid] = lo;
  }
  // the thread's slots in ascending rank (descending lo, invalid last): see fps_bucket_kernel
  for (int i = 1; i < P; ++i) {
    const uint32_t lo = s_lo[i * NT + tid];
    const float x = s_pts[0][i * NT + tid], y = s_pts[1][i * NT + tid], z = s_pts[2][i * NT + tid];
    int j = i - 1;
    while (j >= 0 && s_lo[j * NT + tid] < lo) {
      s_lo[(j + 1) * NT + tid] = s_lo[j * NT + tid];
      s_pts[0][(j + 1) * NT + tid] = s_pts[0][j * NT + tid];
      s_pts[1][(j + 1) * NT + tid] = s_pts[1][j * NT + tid];
      s_pts[2][(j + 1) * NT + tid] = s_pts[2][j * NT + tid];
      --j;
    }
    s_lo[(j + 1) * NT + tid] = lo;
    s_pts[0][(j + 1) * NT + tid] = x;
    s_pts[1][(j + 1) * NT + tid] = y;
    s_pts[2][(j + 1) * NT + tid] = z;
  }
  float px[P], py[P], pz[P], pt[P];
#pragma unroll
  for (int i = 0; i < P; ++i) {
    px[i] = s_pts[0][i * NT + tid];
    py[i] = s_pts[1][i * NT + tid];
    pz[i] = s_pts[2][i * NT + tid];
    pt[i] = s_lo[i * NT + tid] != 0u ? 1e10f : -1.0f;  // reference scratch fill (sampling.cpp:78-80)
  }
  float ox = xyz[0], oy = xyz[1], oz = xyz[2];  // idx[0] = 0, valid or not (sampling_gpu.cu:92-93)
  if (tid == 0) idx[0] = 0;

  for (int it = 0; it < npoint - 1; ++it) {
    const int par = it & 1;
    float best = -1.0f;
    int bi = 0;
#pragma unroll
    for (int g = 0; g < P; g += 4) {   // four points side by side + a compare tree (ILP)
      float m[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (g + u < P) {
          const float dx = __fsub_rn(px[g + u], ox), dy = __fsub_rn(py[g + u], oy),
                      dz = __fsub_rn(pz[g + u], oz);
          m[u] = fminf(sumsq_ref(dx, dy, dz), pt[g + u]);
          pt[g + u] = m[u];
        } else {
          m[u] = -2.0f;
        }
      }
      const bool s01 = m[1] > m[0], s23 = m[3] > m[2];
      const float v01 = s01 ? m[1] : m[0], v23 = s23 ? m[3] : m[2];
      const int i01 = s01 ? g + 1 : g, i23 = s23 ? g + 3 : g + 2;
      const bool sg = v23 > v01;
      const float vg = sg ? v23 : v01;
      const int ig = sg ? i23 : i01;
      if (vg > best) {
        best = vg;
        bi = ig;
      }
    }
    const bool has = best >= 0.0f;
    const uint32_t hi = has ? __float_as_uint(best) : 0u;
    const uint32_t lo = has ? s_lo[bi * NT + tid] : 0u;
    uint32_t whi, wlo;
    warp_argmax(hi, lo, whi, wlo);
    if (wlo == 0u ? lane == 0 : (hi == whi && lo == wlo)) {
      Rec r;
      r.hi = whi; r.lo = wlo;
      r.x = s_pts[0][bi * NT + tid];
      r.y = s_pts[1][bi * NT + tid];
      r.z = s_pts[2][bi * NT + tid];
      r.pad0 = r.pad1 = r.pad2 = 0;
      *reinterpret_cast<uint4 *>(&s_rec[par][warp]) = *reinterpret_cast<uint4 *>(&r);
      s_rec[par][warp].z = r.z;
    }
    __syncthreads();
    uint32_t khi = 0u, klo = 0u;
#pragma unroll
    for (int w = 0; w < NW; ++w) {
      const uint2 k = *reinterpret_cast<const uint2 *>(&s_rec[par][w]);
      if (k.x > khi || (k.x == khi && k.y > klo)) { khi = k.x; klo = k.y; }
    }
    if (klo != 0u) {
      const Rec *w = &s_rec[par][klo & 0xffu];
      ox = w->x; oy = w->y; oz = w->z;
    }
    if (tid == 0) {
      int old = 0;
      if (klo != 0u) {
        const uint32_t rank = (~(klo >> 8)) & rmask;
        const uint32_t tb = rank >> SB;
        const uint32_t tt = L ? (__brev(tb) >> (32 - L)) : 0u;
        old = (int)(tt + ((rank & ((1u << SB) - 1u)) << L));
      }
      idx[it + 1] = old;
    }
  }
}

// `exclusive`: the launch asks for kExclusiveSmem bytes of shared memory it does not use, so that
// no CTA of the step's persistent MLP kernels (167-198 KB each) can become co-resident on its SM.
// An FPS iteration is a dependent chain of a few hundred cycles on 4 warps; sharing the SM's issue
// slots with 16 busy warps stretched the 2048-point level from 0.36 to 0.91 ms inside the
// pipelined step (profiles/r02/cupti_trace_before_exclusive_fps.txt).  The MLP grids are capped
// to leave these SMs free (train_step.PipelinedTrainStep.default_caps).
constexpr int kExclusiveSmem = 100 * 1024;

template <int P, int NT>
cudaError_t launch_small(const float *xyz, int B, int N, int npoint, int *idx, int L, int SB,
                         bool exclusive, cudaStream_t st) {
  auto kern = fps_small_kernel<P, NT>;
  constexpr int own = 4 * P * NT * (int)sizeof(float);
  constexpr int attr = own > kExclusiveSmem ? own : kExclusiveSmem;
  static bool attr_done[64] = {};
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev < 0 || dev >= 64 || !attr_done[dev]) {
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, attr);
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < 64) attr_done[dev] = true;
  }
  const int smem = exclusive ? attr : own;
  kern<<<B, NT, smem, st>>>(xyz, N, npoint, idx, L, SB);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------ host side --
struct Plan {
  int L, SB, csize, NT, P, smem, bits;
};

const int kPList[] = {1, 2, 3, 4, 5, 6, 7, 8, 10, 12, 14, 16, 20, 24};
constexpr int kPMax = 24;

int round_up_P(int p) {
  for (int v : kPList)
    if (v >= p) return v;
  return -1;
}

int g_max_cluster = 16;
constexpr int kSingleCtaMaxN = 4096;
// B2R_FPS_BUCKET_SMALL=1: run the bucket kernel for single-CTA scenes too (tests, A/B timing)
const bool g_bucket_small = []() {
  const char *v = getenv("B2R_FPS_BUCKET_SMALL");
  return v && *v && *v != '0';
}();

bool make_plan(int B, int N, int cluster_hint, Plan *pl) {
  const int bs = b2r_ref_block_threads(N > 0 ? N : 1);
  int L = 0;
  while ((1 << L) < bs) ++L;
  pl->L = L;
  const int rows = (N + bs - 1) / bs;  // k >> L < rows
  int SB = 1;
  while ((1 << SB) < rows) ++SB;
  pl->SB = SB;
  if (L + SB + 8 > 31) return false;
  pl->bits = N > 8192 ? 5 : (N > 1024 ? 4 : 3);
  if (N <= 4096) {
    pl->csize = 1;
    int nt = 32;
    while (nt < N && nt < 512) nt <<= 1;
    pl->NT = nt;
    pl->P = round_up_P((N + nt - 1) / nt);
  } else {
    const int rows512 = (N + 511) / 512;
    // lowest latency: ~10 points per thread (the per-iteration update of a touched bucket and the
    // table reduction over 16*c entries balance there); a hint narrows the cluster
    int c = (rows512 + 9) / 10;
    if (c < 2) c = 2;
    while (c > 2 && (long long)B * c > kNumSMs) --c;
    if (cluster_hint >= 1) c = cluster_hint;
    while (c < g_max_cluster && (rows512 + c - 1) / c > kPMax) ++c;
    if (c > g_max_cluster) c = g_max_cluster;
    pl->csize = c;
    // 16 warps per CTA.  (Measured: 32 warps per CTA with half the points per thread -- 4 CTAs x 1024
    // threads x 10 points -- take 1218 ns per iteration against 824: the 32-warp barrier round costs
    // more than the shorter update saves; the single-CTA kernel shows the same trend down to 4 warps.)
    pl->NT = 512;
    const int p = (rows512 + c - 1) / c;
    if (p > kPMax) return false;
    pl->P = round_up_P(p);
  }
  pl->smem = 4 * pl->P * pl->NT * (int)sizeof(float);
  return pl->P > 0;
}

template <int P, int NT>
cudaError_t launch_inst(const Plan &pl, const float *xyz, const int *perm, int B, int N, int npoint,
                        int *idx, cudaStream_t stream) {
  auto kern = fps_bucket_kernel<P, NT>;
  static bool attr_done[64] = {};  // function attributes are per device
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev < 0 || dev >= 64 || !attr_done[dev]) {
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             4 * P * NT * (int)sizeof(float));
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
  return cudaLaunchKernelEx(&cfg, kern, xyz, perm, N, npoint, idx, pl.L, pl.SB, pl.csize);
}

template <int NT>
cudaError_t launch_nt(const Plan &pl, const float *xyz, const int *perm, int B, int N, int npoint,
                      int *idx, cudaStream_t st) {
  if constexpr (NT < 512) {
    if (pl.P == 1) return launch_inst<1, NT>(pl, xyz, perm, B, N, npoint, idx, st);
    return cudaErrorInvalidValue;
  } else {
    switch (pl.P) {
      case 1: return launch_inst<1, NT>(pl, xyz, perm, B, N, npoint, idx, st);
      case 2: return launch_inst<2, NT>(pl, xyz, perm, B, N, npoint, idx, st);
      case 3: return launch_inst<3, NT>(pl, xyz, perm, B, N, npoint, idx, st);
      case 4: return launch_inst<4, NT>(pl, xyz, perm, B, N, npoint, idx, st);
      case 5: return launch_inst<5, NT>(pl, xyz, perm, B, N, npoint, idx, st);
      case 6: return launch_inst<6, NT>(pl, xyz, perm, B, N, npoint, idx, st);
      case 7: return launch_inst<7, NT>(pl, xyz, perm, B, N, npoint, idx, st);
      case 8: return launch_inst<8, NT>(pl, xyz, perm, B, N, npoint, idx, st);
      case 10: return launch_inst<10, NT>(pl, xyz, perm, B, N, npoint, idx, st);
      case 12: return launch_inst<12, NT>(pl, xyz, perm, B, N, npoint, idx, st);
      case 14: return launch_inst<14, NT>(pl, xyz, perm, B, N, npoint, idx, st);
      case 16: return launch_inst<16, NT>(pl, xyz, perm, B, N, npoint, idx, st);
      case 20: return launch_inst<20, NT>(pl, xyz, perm, B, N, npoint, idx, st);
      case 24: return launch_inst<24, NT>(pl, xyz, perm, B, N, npoint, idx, st);
      default: return cudaErrorInvalidValue;
    }
  }
}

}  // namespace
}  // namespace b2r

extern "C" long long b2r_fps_workspace_bytes(int B, int N) {
  if (B <= 0 || N <= 0) return 0;
  return (long long)B * N * (long long)sizeof(int);
}

namespace {
int fps_sort_launch(const float *xyz, int B, int N, int bits, int *perm, cudaStream_t st) {
  static bool attr_done[64] = {};
  int dev = 0;
  B2R_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || !attr_done[dev]) {
    B2R_CUDA(cudaFuncSetAttribute(b2r::fps_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (1 << 15) * (int)sizeof(int)));
    if (dev >= 0 && dev < 64) attr_done[dev] = true;
  }
  b2r::fps_sort_kernel<<<B, 1024, (size_t)(1 << (3 * bits)) * sizeof(int), st>>>(xyz, N, bits, perm);
  B2R_CHECK_LAUNCH();
  return B2R_OK;
}
int fps_ws_impl(const float *xyz, int B, int N, int npoint, int *idx, int cluster_hint, void *workspace,
                long long workspace_bytes, bool presorted, void *stream);
}  // namespace

extern "C" int b2r_fps_sort(const float *xyz, int B, int N, void *workspace, long long workspace_bytes,
                            void *stream) {
  B2R_REQUIRE(B >= 0 && N >= 0, "b2r_fps_sort: negative size");
  if (B == 0 || N <= b2r::kSingleCtaMaxN) return B2R_OK;   // single-CTA scenes are not sorted
  B2R_REQUIRE(xyz != nullptr && workspace != nullptr && workspace_bytes >= b2r_fps_workspace_bytes(B, N),
              "b2r_fps_sort: null pointer or workspace of %lld bytes, %lld needed", workspace_bytes,
              b2r_fps_workspace_bytes(B, N));
  B2R_REQUIRE(B <= 65535, "b2r_fps_sort: B too large");
  b2r::Plan pl;
  if (!b2r::make_plan(B, N, 0, &pl)) {
    b2r::set_error("b2r_fps_sort: N=%d exceeds the register-resident capacity", N);
    return B2R_ERR_UNSUPPORTED;
  }
  return fps_sort_launch(xyz, B, N, pl.bits, static_cast<int *>(workspace), static_cast<cudaStream_t>(stream));
}

extern "C" int b2r_fps_ws(const float *xyz, int B, int N, int npoint, int *idx, int cluster_hint,
                          void *workspace, long long workspace_bytes, void *stream) {
  return fps_ws_impl(xyz, B, N, npoint, idx, cluster_hint, workspace, workspace_bytes, false, stream);
}

extern "C" int b2r_fps_ws_presorted(const float *xyz, int B, int N, int npoint, int *idx, int cluster_hint,
                                    void *workspace, long long workspace_bytes, void *stream) {
  return fps_ws_impl(xyz, B, N, npoint, idx, cluster_hint, workspace, workspace_bytes, true, stream);
}

namespace {
int fps_ws_impl(const float *xyz, int B, int N, int npoint, int *idx, int cluster_hint, void *workspace,
                long long workspace_bytes, bool presorted, void *stream) {
  B2R_REQUIRE(B >= 0 && N >= 0 && npoint >= 0, "b2r_fps_ws: negative size (B=%d N=%d npoint=%d)", B,
              N, npoint);
  B2R_REQUIRE(cluster_hint >= 0 && cluster_hint <= 16, "b2r_fps_ws: cluster_hint=%d not in [0,16]",
              cluster_hint);
  if (B == 0 || npoint == 0) return B2R_OK;
  B2R_REQUIRE(xyz != nullptr && idx != nullptr, "b2r_fps_ws: null pointer");
  B2R_REQUIRE(B <= 65535, "b2r_fps_ws: B=%d exceeds gridDim.y", B);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (N == 0) {  // nothing to sample from: the reference would emit index 0 everywhere
    B2R_CUDA(cudaMemsetAsync(idx, 0, sizeof(int) * (size_t)B * npoint, st));
    return B2R_OK;
  }
  // one CTA holds the scene: no DSMEM hop to shorten and little to skip -- the few-fat-warps
  // kernel is the shortest chain there (scripts/fps_sweep.py, profiles/r02/fps_table_protocol.txt)
  if (N <= b2r::kSingleCtaMaxN && !b2r::g_bucket_small) {
    b2r::Plan sp;
    if (!b2r::make_plan(B, N, 0, &sp)) return b2r_fps_ex(xyz, B, N, npoint, idx, 0, stream);
    const bool excl = cluster_hint > 0;   // "runs beside other kernels" (b2r_fps_ex's hint)
    cudaError_t e;
    // 4 warps per scene is the measured optimum (2048 / 1024 / 512 points: 351 / 298 / 284 ns per
    // iteration; 8 warps 394 / 362 / 340, 2 warps 423 / 338 / 283; 16 warps = fps.cu 394 / 364 / 332)
    if (N <= 128) e = b2r::launch_small<1, 128>(xyz, B, N, npoint, idx, sp.L, sp.SB, excl, st);
    else if (N <= 256) e = b2r::launch_small<2, 128>(xyz, B, N, npoint, idx, sp.L, sp.SB, excl, st);
    else if (N <= 512) e = b2r::launch_small<4, 128>(xyz, B, N, npoint, idx, sp.L, sp.SB, excl, st);
    else if (N <= 1024) e = b2r::launch_small<8, 128>(xyz, B, N, npoint, idx, sp.L, sp.SB, excl, st);
    else if (N <= 2048) e = b2r::launch_small<16, 128>(xyz, B, N, npoint, idx, sp.L, sp.SB, excl, st);
    else e = b2r::launch_small<16, 256>(xyz, B, N, npoint, idx, sp.L, sp.SB, excl, st);
    if (e != cudaSuccess) {
      b2r::set_error("b2r_fps_ws (single-CTA kernel, N=%d) launch failed: %s", N, cudaGetErrorString(e));
      return B2R_ERR_CUDA;
    }
    return B2R_OK;
  }
  B2R_REQUIRE(workspace != nullptr && workspace_bytes >= b2r_fps_workspace_bytes(B, N),
              "b2r_fps_ws: workspace of %lld bytes, %lld needed", workspace_bytes,
              b2r_fps_workspace_bytes(B, N));
  b2r::Plan pl;
  if (!b2r::make_plan(B, N, cluster_hint, &pl)) {
    b2r::set_error("b2r_fps_ws: N=%d exceeds the register-resident capacity (%d points)", N,
                   16 * 512 * b2r::kPMax);
    return B2R_ERR_UNSUPPORTED;
  }
  int *perm = static_cast<int *>(workspace);
  if (!presorted) {
    const int rc = fps_sort_launch(xyz, B, N, pl.bits, perm, st);
    if (rc != B2R_OK) return rc;
  }
  cudaError_t e = cudaSuccess;
  for (int attempt = 0; attempt < 2; ++attempt) {
    switch (pl.NT) {
      case 32: e = b2r::launch_nt<32>(pl, xyz, perm, B, N, npoint, idx, st); break;
      case 64: e = b2r::launch_nt<64>(pl, xyz, perm, B, N, npoint, idx, st); break;
      case 128: e = b2r::launch_nt<128>(pl, xyz, perm, B, N, npoint, idx, st); break;
      case 256: e = b2r::launch_nt<256>(pl, xyz, perm, B, N, npoint, idx, st); break;
      default: e = b2r::launch_nt<512>(pl, xyz, perm, B, N, npoint, idx, st); break;
    }
    if (e == cudaSuccess || pl.csize <= 8) break;
    // a non-portable (> 8 CTA) cluster was refused: fall back to the portable maximum, once
    (void)cudaGetLastError();
    b2r::g_max_cluster = 8;
    if (!b2r::make_plan(B, N, cluster_hint > 8 ? 8 : cluster_hint, &pl)) {
      b2r::set_error("b2r_fps_ws: N=%d needs a cluster of more than 8 CTAs, which this device refused", N);
      return B2R_ERR_UNSUPPORTED;
    }
  }
  if (e != cudaSuccess) {
    b2r::set_error("b2r_fps_ws launch (cluster=%d threads=%d P=%d smem=%d) failed: %s", pl.csize,
                   pl.NT, pl.P, pl.smem, cudaGetErrorString(e));
    return B2R_ERR_CUDA;
  }
  return B2R_OK;
}
}  // namespace
