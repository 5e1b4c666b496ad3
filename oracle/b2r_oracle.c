/*
 * b2r_oracle.c -- CPU restatement of the reference's nine `pointnet2._ext` ops.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product path:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this library, and only as the checker / CPU baseline.
 *
 * Each function restates, in plain C, the arithmetic of one reference CUDA kernel
 * (paths relative to /root/reference/detection/Votenet/pointnet2/_ext_src/):
 *
 *   orc_fps              src/sampling_gpu.cu:64-178 (+ scratch init src/sampling.cpp:78-80,
 *                        block size rule include/cuda_utils.h:20-24)
 *   orc_gather(_grad)    src/sampling_gpu.cu:13-52
 *   orc_ball_query       src/ball_query_gpu.cu:14-49 (+ zero init src/ball_query.cpp:24-26)
 *   orc_group(_grad)     src/group_points_gpu.cu:13-69
 *   orc_three_nn         src/interpolate_gpu.cu:14-64
 *   orc_interp(_grad)    src/interpolate_gpu.cu:77-148
 *
 * Float contraction: nvcc 12.9 -O2 (-fmad=true) compiles a*a + b*b + c*c in these
 * kernels to  FMUL b*b ; FFMA a,a ; FFMA c,c  (SURVEY.md appendix A).  This file
 * must be built with -ffp-contract=off and spells that contraction out with fmaf().
 *
 * Parity status: pinned on the GPU box against the reference's own kernels compiled
 * for sm_100a (oracle/_ref, see oracle/Makefile and tests/test_ref_pin_gpu.py) and
 * against the reference's only known-answer test (pointnet2_test.py:18-30).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

/* sum of squares with the reference SASS contraction: y-term rounded, x then z fused */
static inline float sumsq(float a, float b, float c) {
  return fmaf(c, c, fmaf(a, a, b * b));
}

/* include/cuda_utils.h:20-24 -- power-of-two thread count, capped at 512.
 * Uses the same double log()/log(2.0) quotient + truncation as the reference. */
int orc_block_threads(int work) {
  int p = (int)(log((double)work) / log(2.0));
  int t = 1 << p;
  if (t > 512) t = 512;
  if (t < 1) t = 1;
  return t;
}

int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* ---------------------------------------------------------------- FPS ----- */
/* One scene.  Emulates the CTA literally: `bs` lanes, each scanning k = t, t+bs, ...
 * keeping the first strictly larger running-min distance, then the shared-memory
 * tree whose merge keeps the lower slot unless the upper one is strictly larger. */
static void fps_scene(const float *p, int n, int m, int bs, float *tmp, int *out,
                      float *lane_v, int *lane_i) {
  if (m <= 0) return;
  for (int k = 0; k < n; ++k) tmp[k] = 1e10f; /* sampling.cpp:78-80 */
  int old = 0;
  out[0] = 0;
  for (int j = 1; j < m; ++j) {
    const float ox = p[old * 3 + 0], oy = p[old * 3 + 1], oz = p[old * 3 + 2];
    for (int t = 0; t < bs; ++t) {
      float best = -1.0f;
      int besti = 0;
      for (int k = t; k < n; k += bs) {
        const float x = p[k * 3 + 0], y = p[k * 3 + 1], z = p[k * 3 + 2];
        const float mag = sumsq(x, y, z);
        if ((double)mag <= 1e-3) continue; /* sampling_gpu.cu:105-106, double compare */
        const float d = sumsq(x - ox, y - oy, z - oz);
        const float d2 = fminf(d, tmp[k]);
        tmp[k] = d2;
        if (d2 > best) {
          best = d2;
          besti = k;
        }
      }
      lane_v[t] = best;
      lane_i[t] = besti;
    }
    for (int s = bs >> 1; s >= 1; s >>= 1) {
      for (int t = 0; t < s; ++t) {
        const float v1 = lane_v[t], v2 = lane_v[t + s];
        const int i1 = lane_i[t], i2 = lane_i[t + s];
        lane_v[t] = v1 > v2 ? v1 : (v2 > v1 ? v2 : v1); /* max(v1,v2) */
        lane_i[t] = v2 > v1 ? i2 : i1;
      }
    }
    old = lane_i[0];
    out[j] = old;
  }
}

/* xyz (B,N,3) f32 -> idx (B,npoint) i32 */
void orc_fps(const float *xyz, int B, int N, int npoint, int *idx) {
  const int bs = orc_block_threads(N);
#pragma omp parallel for schedule(dynamic, 1)
  for (int b = 0; b < B; ++b) {
    float *tmp = (float *)malloc(sizeof(float) * (size_t)(N > 0 ? N : 1));
    float *lv = (float *)malloc(sizeof(float) * 512);
    int *li = (int *)malloc(sizeof(int) * 512);
    fps_scene(xyz + (size_t)b * N * 3, N, npoint, bs, tmp, idx + (size_t)b * npoint, lv, li);
    free(tmp);
    free(lv);
    free(li);
  }
}

/* -------------------------------------------------------- gather / group -- */
/* features (B,C,N), idx (B,M) -> out (B,C,M) */
void orc_gather(const float *f, const int *idx, int B, int C, int N, int M, float *out) {
#pragma omp parallel for collapse(2)
  for (int b = 0; b < B; ++b)
    for (int c = 0; c < C; ++c) {
      const float *src = f + ((size_t)b * C + c) * N;
      float *dst = out + ((size_t)b * C + c) * M;
      const int *ix = idx + (size_t)b * M;
      for (int j = 0; j < M; ++j) dst[j] = src[ix[j]];
    }
}

/* grad_out (B,C,M), idx (B,M) -> grad_features (B,C,N), scatter-add into zeros */
void orc_gather_grad(const float *g, const int *idx, int B, int C, int N, int M, float *out) {
  memset(out, 0, sizeof(float) * (size_t)B * C * N);
#pragma omp parallel for collapse(2)
  for (int b = 0; b < B; ++b)
    for (int c = 0; c < C; ++c) {
      const float *src = g + ((size_t)b * C + c) * M;
      float *dst = out + ((size_t)b * C + c) * N;
      const int *ix = idx + (size_t)b * M;
      for (int j = 0; j < M; ++j) dst[ix[j]] += src[j];
    }
}

/* features (B,C,N), idx (B,NP,NS) -> out (B,C,NP,NS) */
void orc_group(const float *f, const int *idx, int B, int C, int N, int NP, int NS, float *out) {
  const size_t L = (size_t)NP * NS;
#pragma omp parallel for collapse(2)
  for (int b = 0; b < B; ++b)
    for (int c = 0; c < C; ++c) {
      const float *src = f + ((size_t)b * C + c) * N;
      float *dst = out + ((size_t)b * C + c) * L;
      const int *ix = idx + (size_t)b * L;
      for (size_t e = 0; e < L; ++e) dst[e] = src[ix[e]];
    }
}

/* grad_out (B,C,NP,NS), idx (B,NP,NS) -> grad_features (B,C,N) */
void orc_group_grad(const float *g, const int *idx, int B, int C, int N, int NP, int NS,
                    float *out) {
  const size_t L = (size_t)NP * NS;
  memset(out, 0, sizeof(float) * (size_t)B * C * N);
#pragma omp parallel for collapse(2)
  for (int b = 0; b < B; ++b)
    for (int c = 0; c < C; ++c) {
      const float *src = g + ((size_t)b * C + c) * L;
      float *dst = out + ((size_t)b * C + c) * N;
      const int *ix = idx + (size_t)b * L;
      for (size_t e = 0; e < L; ++e) dst[ix[e]] += src[e];
    }
}

/* ----------------------------------------------------------- ball query --- */
/* new_xyz (B,M,3), xyz (B,N,3) -> idx (B,M,NS); first NS hits in index order,
 * every slot pre-filled with the first hit, zeros when the ball is empty. */
void orc_ball_query(const float *new_xyz, const float *xyz, int B, int N, int M, float radius,
                    int NS, int *idx) {
  const float r2 = radius * radius;
  memset(idx, 0, sizeof(int) * (size_t)B * M * NS);
#pragma omp parallel for collapse(2) schedule(dynamic, 16)
  for (int b = 0; b < B; ++b)
    for (int j = 0; j < M; ++j) {
      const float *q = new_xyz + ((size_t)b * M + j) * 3;
      const float *p = xyz + (size_t)b * N * 3;
      int *o = idx + ((size_t)b * M + j) * NS;
      const float cx = q[0], cy = q[1], cz = q[2];
      int cnt = 0;
      for (int k = 0; k < N && cnt < NS; ++k) {
        const float d2 = sumsq(cx - p[k * 3 + 0], cy - p[k * 3 + 1], cz - p[k * 3 + 2]);
        if (d2 < r2) {
          if (cnt == 0)
            for (int l = 0; l < NS; ++l) o[l] = k;
          o[cnt++] = k;
        }
      }
    }
}

/* ------------------------------------------------------------- three_nn --- */
/* unknown (B,n,3), known (B,m,3) -> dist2 (B,n,3) f32, idx (B,n,3) i32 */
void orc_three_nn(const float *unknown, const float *known, int B, int n, int m, float *dist2,
                  int *idx) {
#pragma omp parallel for collapse(2) schedule(static)
  for (int b = 0; b < B; ++b)
    for (int j = 0; j < n; ++j) {
      const float *u = unknown + ((size_t)b * n + j) * 3;
      const float *p = known + (size_t)b * m * 3;
      const float ux = u[0], uy = u[1], uz = u[2];
      double b1 = 1e40, b2 = 1e40, b3 = 1e40; /* interpolate_gpu.cu:32 */
      int i1 = 0, i2 = 0, i3 = 0;
      for (int k = 0; k < m; ++k) {
        const float d = sumsq(ux - p[k * 3 + 0], uy - p[k * 3 + 1], uz - p[k * 3 + 2]);
        if (d < b1) {
          b3 = b2; i3 = i2;
          b2 = b1; i2 = i1;
          b1 = d;  i1 = k;
        } else if (d < b2) {
          b3 = b2; i3 = i2;
          b2 = d;  i2 = k;
        } else if (d < b3) {
          b3 = d;  i3 = k;
        }
      }
      float *od = dist2 + ((size_t)b * n + j) * 3;
      int *oi = idx + ((size_t)b * n + j) * 3;
      od[0] = (float)b1; od[1] = (float)b2; od[2] = (float)b3;
      oi[0] = i1; oi[1] = i2; oi[2] = i3;
    }
}

/* ---------------------------------------------------- three_interpolate --- */
/* features (B,C,m), idx (B,n,3), weight (B,n,3) -> out (B,C,n) */
void orc_interp(const float *f, const int *idx, const float *w, int B, int C, int m, int n,
                float *out) {
#pragma omp parallel for collapse(2)
  for (int b = 0; b < B; ++b)
    for (int c = 0; c < C; ++c) {
      const float *src = f + ((size_t)b * C + c) * m;
      float *dst = out + ((size_t)b * C + c) * n;
      const int *ix = idx + (size_t)b * n * 3;
      const float *ww = w + (size_t)b * n * 3;
      for (int j = 0; j < n; ++j) {
        const float p1 = src[ix[j * 3 + 0]], p2 = src[ix[j * 3 + 1]], p3 = src[ix[j * 3 + 2]];
        dst[j] = fmaf(p3, ww[j * 3 + 2], fmaf(p1, ww[j * 3 + 0], p2 * ww[j * 3 + 1]));
      }
    }
}

/* grad_out (B,C,n), idx, weight (B,n,3) -> grad_features (B,C,m) */
void orc_interp_grad(const float *g, const int *idx, const float *w, int B, int C, int n, int m,
                     float *out) {
  memset(out, 0, sizeof(float) * (size_t)B * C * m);
#pragma omp parallel for collapse(2)
  for (int b = 0; b < B; ++b)
    for (int c = 0; c < C; ++c) {
      const float *src = g + ((size_t)b * C + c) * n;
      float *dst = out + ((size_t)b * C + c) * m;
      const int *ix = idx + (size_t)b * n * 3;
      const float *ww = w + (size_t)b * n * 3;
      for (int j = 0; j < n; ++j) {
        dst[ix[j * 3 + 0]] += src[j] * ww[j * 3 + 0];
        dst[ix[j * 3 + 1]] += src[j] * ww[j * 3 + 1];
        dst[ix[j * 3 + 2]] += src[j] * ww[j * 3 + 2];
      }
    }
}
