// fps_phase_probe.cu -- where does an FPS iteration of csrc/fps_bucket.cu spend its cycles?
// Compiles the product kernel with B2R_FPS_PROFILE (per-warp clock64 sums per phase) and runs it
// on synthetic room-like scenes.  usage: fps_phase_probe N npoint cluster_hint
#define B2R_FPS_PROFILE 1
#include <cstdarg>
#include <cstdlib>
#include <vector>
#include "../../backtoreality_b200/csrc/fps_bucket.cu"
namespace b2r { void set_error(const char *fmt, ...) { va_list ap; va_start(ap, fmt); vfprintf(stderr, fmt, ap); va_end(ap); fputc('\n', stderr); } }
extern "C" int b2r_fps_ex(const float *, int, int, int, int *, int, void *) { fprintf(stderr, "set B2R_FPS_BUCKET_SMALL=1\n"); return -3; }
extern "C" int b2r_ref_block_threads(int n) { int p = 1; while (p * 2 <= n && p < 512) p *= 2; return p; }

int main(int argc, char **argv) {
  const int N = argc > 1 ? atoi(argv[1]) : 40000, np = argc > 2 ? atoi(argv[2]) : 2048,
            hint = argc > 3 ? atoi(argv[3]) : 0, B = 8;
  std::vector<float> h((size_t)B * N * 3);
  srand(1);
  auto rnd = []() { return (float)rand() / RAND_MAX; };
  for (int b = 0; b < B; ++b)
    for (int k = 0; k < N; ++k) {   // floor + two walls + boxes: points on surfaces of an 8x6x3 room
      float *p = &h[((size_t)b * N + k) * 3];
      const int s = rand() % 5;
      if (s < 2) { p[0] = rnd() * 8; p[1] = rnd() * 6; p[2] = 0.01f * rnd(); }
      else if (s == 2) { p[0] = rnd() * 8; p[1] = 0.01f * rnd(); p[2] = rnd() * 3; }
      else if (s == 3) { p[0] = 0.01f * rnd(); p[1] = rnd() * 6; p[2] = rnd() * 3; }
      else { p[0] = 2 + rnd(); p[1] = 2 + rnd() * 2; p[2] = 0.8f; }
      p[0] += 1; p[1] += 1; p[2] += 1;
    }
  float *xyz; int *idx, *ws;
  cudaMalloc(&xyz, h.size() * 4); cudaMalloc(&idx, (size_t)B * np * 4); cudaMalloc(&ws, (size_t)B * N * 4);
  cudaMemcpy(xyz, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  for (int rep = 0; rep < 2; ++rep) {
    unsigned long long zero[b2r::kTable][8] = {};
    cudaMemcpyToSymbol(b2r::g_prof, zero, sizeof(zero));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    int rc = b2r_fps_ws(xyz, B, N, np, idx, hint, ws, (long long)B * N * 4, nullptr);
    cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (rep == 0) continue;
    unsigned long long pr[b2r::kTable][8];
    cudaMemcpyFromSymbol(pr, b2r::g_prof, sizeof(pr));
    b2r::Plan pl; b2r::make_plan(B, N, hint, &pl);
    const int W = pl.csize * pl.NT / 32;
    printf("rc=%d N=%d np=%d cluster=%d NT=%d P=%d W=%d: %.3f ms, %.0f ns/iteration (%s)\n", rc, N, np, pl.csize, pl.NT, pl.P, W, ms,
           ms * 1e6 / (np - 1), cudaGetErrorString(cudaGetLastError()));
    double touched = 0, c_t = 0, c_s = 0, w_t = 0, w_s = 0, its = 0;
    for (int w = 0; w < W; ++w) {
      touched += pr[w][0]; its += pr[w][0] + pr[w][4];
      c_t += pr[w][1]; c_s += pr[w][5]; w_t += pr[w][3]; w_s += pr[w][6];
    }
    const double skipped = its - touched;
    printf("  buckets touched per iteration: %.1f of %d\n", touched / (np - 1), W);
    printf("  touched warp : test+update+argmax+record %.0f cyc, exchange+final reduce %.0f cyc\n", c_t / touched, w_t / touched);
    printf("  skipped warp : test+carry %.0f cyc, exchange+final reduce %.0f cyc\n", skipped ? c_s / skipped : 0, skipped ? w_s / skipped : 0);
    printf("  iteration %.0f cyc (includes ~3 clock reads)\n", (c_t + c_s + w_t + w_s) / its);
  }
  return 0;
}
