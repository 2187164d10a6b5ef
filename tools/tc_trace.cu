// Developer tool (not part of the product): runs the tcgen05 forward kernel standalone on the C2 shape,
// times it with CUDA events and prints the in-kernel clock64 trace of CTA 0 (steps 16..19).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I include -I kws_b200/csrc -o tools/tc_trace tools/tc_trace.cu
#ifndef NO_TRACE
#define FGRNN_TC_TRACE
#endif
#include "../kws_b200/csrc/fgrnn_tc.cu"

#include <cstdarg>
#include <cmath>
#include <cstdlib>
#include <vector>

namespace fgrnn {
void set_error_detail(const char* fmt, ...) { va_list ap; va_start(ap, fmt); vfprintf(stderr, fmt, ap); va_end(ap); fprintf(stderr, "\n"); }
void count_launch(int) {}
}  // namespace fgrnn

int main(int argc, char** argv) {
  using namespace fgrnn;
  const int B = argc > 1 ? atoi(argv[1]) : 8192, T = argc > 2 ? atoi(argv[2]) : 99, I = 32, H = 128;
  const bool with_out = argc > 3 ? atoi(argv[3]) != 0 : true;
  std::vector<float> hx((size_t)B * T * I), hW(I * H), hU(H * H), hb(H, 1.0f);
  srand(1);
  auto rnd = []() { double u = 0; for (int i = 0; i < 12; ++i) u += rand() / (double)RAND_MAX; return (float)(u - 6.0); };
  for (auto& v : hx) v = rnd();
  for (auto& v : hW) v = 0.1f * rnd();
  for (auto& v : hU) v = 0.1f * rnd();
  float *x, *W, *U, *bg, *bu, *zeta, *nu, *out, *hl;
  cudaMalloc(&x, hx.size() * 4); cudaMalloc(&W, hW.size() * 4); cudaMalloc(&U, hU.size() * 4);
  cudaMalloc(&bg, H * 4); cudaMalloc(&bu, H * 4); cudaMalloc(&zeta, 4); cudaMalloc(&nu, 4);
  cudaMalloc(&out, (size_t)B * T * H * 4); cudaMalloc(&hl, (size_t)B * H * 4);
  cudaMemcpy(x, hx.data(), hx.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(W, hW.data(), hW.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(U, hU.data(), hU.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(bg, hb.data(), H * 4, cudaMemcpyHostToDevice); cudaMemcpy(bu, hb.data(), H * 4, cudaMemcpyHostToDevice);
  const float z0 = 1.0f, n0 = -4.0f;
  cudaMemcpy(zeta, &z0, 4, cudaMemcpyHostToDevice); cudaMemcpy(nu, &n0, 4, cudaMemcpyHostToDevice);
  SmemFwdArgs a{};
  a.d = Dims{B, T, I, H, 0, 0, FGRNN_NL_SIGMOID, FGRNN_NL_TANH, FGRNN_F32};
  a.layout = FGRNN_LAYOUT_IH; a.W = W; a.U = U; a.bias_gate = bg; a.bias_update = bu; a.zeta = zeta; a.nu = nu;
  a.x = x; a.xsb = (int64_t)T * I; a.xst = I; a.h0 = nullptr;
  a.out = with_out ? out : nullptr; a.osb = (int64_t)T * H; a.ost = H; a.h_last = hl;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 3; ++i) launch_tc_fwd(a, 0);
  cudaDeviceSynchronize();
  const int reps = 50;
  cudaEventRecord(e0);
  for (int i = 0; i < reps; ++i) launch_tc_fwd(a, 0);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  printf("B=%d T=%d out=%d : %.1f us per launch, %.2f M seq/s, %.0f cycles/step @1.965GHz  (%s)\n", B, T, (int)with_out, ms / reps * 1e3,
         B / (ms / reps * 1e-3) * 1e-6, ms / reps * 1e-3 / T * 1.965e9, cudaGetErrorString(cudaGetLastError()));
  {  // accuracy of the first 64 rows against an fp64 evaluation of the recurrence, in units of the tolerance
    const int R = B < 64 ? B : 64;
    std::vector<float> ho((size_t)R * T * H);
    if (with_out) {
      cudaMemcpy(ho.data(), out, ho.size() * 4, cudaMemcpyDeviceToHost);
      std::vector<double> h(H), hn(H);
      const double sz = 1.0 / (1.0 + exp(-1.0)), sn = 1.0 / (1.0 + exp(4.0));
      double worst = 0;
      for (int r = 0; r < R; ++r) {
        for (auto& v : h) v = 0;
        for (int t = 0; t < T; ++t) {
          for (int n = 0; n < H; ++n) {
            double pre = 0;
            for (int k = 0; k < I; ++k) pre += (double)hx[((size_t)r * T + t) * I + k] * hW[k * H + n];
            for (int k = 0; k < H; ++k) pre += h[k] * hU[k * H + n];
            const double z = 1.0 / (1.0 + exp(-(pre + 1.0))), c = tanh(pre + 1.0);
            hn[n] = z * h[n] + (sz * (1 - z) + sn) * c;
          }
          h = hn;
          for (int n = 0; n < H; ++n) {
            const double e = fabs((double)ho[((size_t)r * T + t) * H + n] - h[n]) / (1e-6 + 1e-5 * fabs(h[n]));
            if (e > worst) worst = e;
          }
        }
      }
      printf("accuracy vs fp64 over %d rows: max |err| / (1e-6 + 1e-5 |h|) = %.3f\n", R, worst);
    }
  }
#ifdef FGRNN_TC_TRACE
  {
    const int nc = (B + 63) / 64 < 1024 ? (B + 63) / 64 : 1024;
    std::vector<unsigned long long> ct(1024 * 4);
    cudaMemcpyFromSymbol(ct.data(), g_tc_cta_time, ct.size() * 8);
    unsigned long long t0 = ~0ull, t1 = 0; double pro = 0, loop_min = 1e30, loop_max = 0, loop_sum = 0;
    for (int c = 0; c < nc; ++c) {
      if (ct[c * 4] < t0) t0 = ct[c * 4];
      if (ct[c * 4 + 2] > t1) t1 = ct[c * 4 + 2];
      pro += (double)(ct[c * 4 + 1] - ct[c * 4]);
      const double lp = (double)(ct[c * 4 + 2] - ct[c * 4 + 1]);
      loop_sum += lp; if (lp < loop_min) loop_min = lp; if (lp > loop_max) loop_max = lp;
    }
    printf("CTAs: first start -> last end %.1f us | prologue avg %.1f us | main loop min %.1f avg %.1f max %.1f us | start spread %.1f us\n",
           (t1 - t0) * 1e-3, pro / nc * 1e-3, loop_min * 1e-3, loop_sum / nc * 1e-3, loop_max * 1e-3,
           0.0);
  }
  long long tr[4 * 2 * 16];
  cudaMemcpyFromSymbol(tr, g_tc_trace, sizeof(tr));
  const long long base = tr[8];
  const char* names[16] = {"epi:wait_d", "epi:d_ready", "epi:ld_done", "epi:math+sts", "epi:arrived", "epi:stg_done", "", "",
                           "mma:wait_h", "mma:h_ready", "mma:issued", "", "cnv:wait_raw", "cnv:raw_ready", "cnv:x_empty", "cnv:done"};
  for (int t = 0; t < 4; ++t)
    for (int s = 0; s < 2; ++s) {
      printf("t=%d s=%d |", 16 + t, s);
      for (int k = 0; k < 16; ++k) if (names[k][0] && tr[(t * 2 + s) * 16 + k]) printf(" %s %lld", names[k], tr[(t * 2 + s) * 16 + k] - base);
      printf("\n");
    }
#endif
  return 0;
}
