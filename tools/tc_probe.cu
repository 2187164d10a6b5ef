// Developer tool (not part of the product): tcgen05 accumulation-rounding and timing probe on sm_100a.
//   accuracy : D = C0 + A.B with fp16 operands, K split into K/16 MMAs chained through the TMEM
//              accumulator; the result is compared bit-for-bit with four host models
//              (round-to-nearest / truncation, once at the end or once per MMA).
//   timing   : cycles per tcgen05.mma for several (M, N), tcgen05.ld cost, commit->wait latency.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tc_probe tools/tc_probe.cu
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  for (uint32_t spins = 0; !ok; ++spins) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (spins > (1u << 24)) __trap();
  }
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
               ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// K-major, no swizzle: 8 rows x 16 B core matrices, LBO (next K chunk) = 128 B, SBO (next 8 rows) = (K/8)*128 B
__host__ __device__ inline uint32_t op_offset(int row, int k, int K) { return (uint32_t)((row >> 3) * (K >> 3) * 128 + (k >> 3) * 128 + (row & 7) * 16 + (k & 7) * 2); }
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, int K) {
  const uint64_t lbo = 128 >> 4, sbo = (uint64_t)((K >> 3) * 128) >> 4;
  return (uint64_t)((smem_addr >> 4) & 0x3fff) | (lbo << 16) | (sbo << 32) | (1ull << 46);
}
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) { return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }

// ---- accuracy --------------------------------------------------------------------------------
// A[128][K], B[128][K] fp16 row-major (K contiguous); C0[128][128] fp32 or null; D[128][128]
__global__ void __launch_bounds__(128, 1) acc_kernel(const __half* A, const __half* B, const float* C0, float* D, int K) {
  extern __shared__ __align__(128) unsigned char sm[];
  unsigned char* As = sm;
  unsigned char* Bs = sm + 128 * K * 2;
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_s)), "r"(128) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  for (int e = tid; e < 128 * K; e += 128) {
    const int r = e / K, k = e - r * K;
    *reinterpret_cast<__half*>(As + op_offset(r, k, K)) = A[e];
    *reinterpret_cast<__half*>(Bs + op_offset(r, k, K)) = B[e];
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_s;
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
  if (C0) {
    for (int c0 = 0; c0 < 128; c0 += 16) {
      uint32_t v[16];
      for (int j = 0; j < 16; ++j) v[j] = __float_as_uint(C0[(warp * 32 + lane) * 128 + c0 + j]);
      tmem_st16(tmem + lane_base + c0, v);
    }
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (tid == 0) {
    const uint64_t dA = make_desc(smem_u32(As), K), dB = make_desc(smem_u32(Bs), K);
    for (int ks = 0; ks < K / 16; ++ks) {
      const uint64_t adv = (uint64_t)((ks * 256) >> 4);
      umma_f16(tmem, dA + adv, dB + adv, make_idesc(128, 128), (C0 != nullptr) || ks > 0);
    }
    umma_commit(smem_u32(&bar));
  }
  mbar_wait(smem_u32(&bar), 0);
  tc_fence_after();
  for (int c0 = 0; c0 < 128; c0 += 16) {
    uint32_t v[16];
    tmem_ld16(tmem + lane_base + c0, v);
    tmem_ld_wait();
    for (int j = 0; j < 16; ++j) D[(warp * 32 + lane) * 128 + c0 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128) : "memory");
}

// ---- timing ----------------------------------------------------------------------------------
// out[0] = cycles for n_mma MMAs (issue -> commit -> wait); out[1] = cycles for the issue loop alone;
// out[2] = cycles for 4 warps each reading `ldcols` columns of their quadrant with tcgen05.ld.x16
__global__ void __launch_bounds__(128, 1) time_kernel(long long* out, int M, int N, int n_mma, int ldcols, int a_stride_smem, int unrolled) {
  extern __shared__ __align__(128) unsigned char sm[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_s)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  for (int e = tid; e < 48 * 1024; e += 128) reinterpret_cast<uint32_t*>(sm)[e] = 0;   // 192 KB of zeros
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_s;
  long long t0 = 0, t1 = 0, t2 = 0;
  if (tid == 0) {
    const uint64_t dA = make_desc(smem_u32(sm), 128), dB = make_desc(smem_u32(sm + 64 * 1024), 128);
    const uint32_t idesc = make_idesc(M, N);
    uint64_t da[8], db[8];
    for (int i = 0; i < 8; ++i) { da[i] = dA + (uint64_t)((i * a_stride_smem) >> 4); db[i] = dB + (uint64_t)((i * a_stride_smem) >> 4); }
    t0 = clock64();
    if (unrolled) {
      for (int i = 0; i < n_mma; i += 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) umma_f16(tmem + (uint32_t)((j & 1) * 256), da[j], db[j], idesc, 1);
      }
    } else {
      for (int i = 0; i < n_mma; ++i) {
        const uint64_t adv = (uint64_t)(((i & 7) * a_stride_smem) >> 4);
        umma_f16(tmem + (uint32_t)((i & 1) * 256), dA + adv, dB + adv, idesc, 1);
      }
    }
    umma_commit(smem_u32(&bar));
    t1 = clock64();
  }
  mbar_wait(smem_u32(&bar), 0);
  tc_fence_after();
  if (tid == 0) { t2 = clock64(); out[0] = t2 - t0; out[1] = t1 - t0; }
  __syncthreads();
  const long long t3 = clock64();
  uint32_t accx = 0;
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
  for (int c0 = 0; c0 < ldcols; c0 += 16) {
    uint32_t v[16];
    tmem_ld16(tmem + lane_base + c0, v);
    tmem_ld_wait();
    for (int j = 0; j < 16; ++j) accx ^= v[j];
  }
  __syncthreads();
  const long long t4 = clock64();
  if (tid == 0) out[2] = t4 - t3;
  if (accx == 0x12345) out[3] = 1;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

// ---- host models -----------------------------------------------------------------------------
static float rz_from_double(double v) {          // truncate to fp32 (toward zero)
  float f = (float)v;                            // RN
  if (std::fabs((double)f) > std::fabs(v)) f = std::nextafterf(f, 0.0f);
  return f;
}
static double ulp_of(float f) { int e; std::frexp(f, &e); return std::ldexp(1.0, e - 24); }

static void run_accuracy(int K, bool with_c0, float c0_scale, float a_scale, float b_scale, const char* dump = nullptr) {
  std::vector<__half> A(128 * K), B(128 * K);
  std::vector<float> C0(128 * 128), D(128 * 128);
  auto rnd = []() { double u = 0; for (int i = 0; i < 12; ++i) u += rand() / (double)RAND_MAX; return u - 6.0; };
  for (auto& v : A) v = __float2half((float)(rnd() * a_scale));
  for (auto& v : B) v = __float2half((float)(rnd() * b_scale));
  for (auto& v : C0) v = (float)(rnd() * c0_scale);
  __half *dA, *dB; float *dC, *dD;
  CK(cudaMalloc(&dA, A.size() * 2)); CK(cudaMalloc(&dB, B.size() * 2)); CK(cudaMalloc(&dC, C0.size() * 4)); CK(cudaMalloc(&dD, D.size() * 4));
  CK(cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dC, C0.data(), C0.size() * 4, cudaMemcpyHostToDevice));
  const size_t smem = 2 * 128 * K * 2;
  CK(cudaFuncSetAttribute(acc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  acc_kernel<<<1, 128, smem>>>(dA, dB, with_c0 ? dC : nullptr, dD, K);
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
  if (dump) {
    FILE* f = fopen(dump, "wb");
    if (f) {
      int hdr[2] = {K, (int)with_c0};
      fwrite(hdr, 4, 2, f); fwrite(A.data(), 2, A.size(), f); fwrite(B.data(), 2, B.size(), f);
      fwrite(C0.data(), 4, C0.size(), f); fwrite(D.data(), 4, D.size(), f); fclose(f);
    }
  }
  long match_rn_end = 0, match_rz_end = 0, match_rn_step = 0, match_rz_step = 0;
  double sum_err = 0, sum_sq = 0, max_err = 0;
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < 128; ++n) {
      double exact = with_c0 ? (double)C0[m * 128 + n] : 0.0;
      float acc_rn = with_c0 ? C0[m * 128 + n] : 0.f, acc_rz = acc_rn;
      for (int ks = 0; ks < K / 16; ++ks) {
        double blk = 0;
        for (int k = ks * 16; k < ks * 16 + 16; ++k) blk += (double)__half2float(A[m * K + k]) * (double)__half2float(B[n * K + k]);
        exact += blk;
        acc_rn = (float)((double)acc_rn + blk);
        acc_rz = rz_from_double((double)acc_rz + blk);
      }
      const float got = D[m * 128 + n];
      match_rn_end += got == (float)exact;
      match_rz_end += got == rz_from_double(exact);
      match_rn_step += got == acc_rn;
      match_rz_step += got == acc_rz;
      const double e = ((double)got - exact) / ulp_of((float)exact) * (exact < 0 ? -1.0 : 1.0);   // >0: magnitude too large
      sum_err += e; sum_sq += e * e; max_err = std::fmax(max_err, std::fabs(e));
    }
  const double n = 128.0 * 128.0;
  printf("ACC K=%3d c0=%d(scale %g) a=%g b=%g | match: rn_end %5.1f%% rz_end %5.1f%% rn_step %5.1f%% rz_step %5.1f%% | err/ulp: mean(signed,toward larger |.|) %+.3f rms %.3f max %.2f\n",
         K, (int)with_c0, c0_scale, a_scale, b_scale, 100 * match_rn_end / n, 100 * match_rz_end / n, 100 * match_rn_step / n, 100 * match_rz_step / n,
         sum_err / n, std::sqrt(sum_sq / n), max_err);
  cudaFree(dA); cudaFree(dB); cudaFree(dC); cudaFree(dD);
}

static void run_timing(int M, int N, int n_mma, int ldcols, int unrolled) {
  long long* d; long long h[4] = {0, 0, 0, 0};
  CK(cudaMalloc(&d, 32)); CK(cudaMemset(d, 0, 32));
  const size_t smem = 192 * 1024;
  CK(cudaFuncSetAttribute(time_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  for (int rep = 0; rep < 2; ++rep) { time_kernel<<<1, 128, smem>>>(d, M, N, n_mma, ldcols, 256, unrolled); CK(cudaDeviceSynchronize()); }
  CK(cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost));
  printf("TIME unrolled=%d M=%3d N=%3d n_mma=%4d : total %6lld cyc (%.1f/mma)  issue-only %6lld cyc (%.1f/mma) | tcgen05.ld %3d cols x 4 warps: %lld cyc\n",
         unrolled, M, N, n_mma, h[0], (double)h[0] / n_mma, h[1], (double)h[1] / n_mma, ldcols, h[2]);
  cudaFree(d);
}

int main() {
  srand(1234);
  run_accuracy(16, false, 0.f, 1.f, 1.f, "gpurun_out/mma_k16.bin");
  run_accuracy(16, true, 8.f, 1.f, 1.f, "gpurun_out/mma_k16_c0.bin");
  run_accuracy(16, true, 1.f, 1.f, 1.f / 2048, "gpurun_out/mma_k16_c0_small.bin");
  run_accuracy(64, true, 1.f, 1.f, 1.f, "gpurun_out/mma_k64_c0.bin");
  run_accuracy(16, true, 1.f, 1.f, 1.f);
  run_accuracy(16, true, 64.f, 1.f, 1.f);
  run_accuracy(16, true, 1.f, 1.f, 1.f / 2048);     // products ~2^-11 below the accumulator (the lo.hi terms)
  run_accuracy(32, false, 0.f, 1.f, 1.f);
  run_accuracy(128, false, 0.f, 1.f, 1.f);
  run_accuracy(128, true, 4.f, 1.f, 1.f);
  run_accuracy(128, false, 0.f, 1.f, 1.f / 2048);
  const int shapes[][2] = {{128, 128}, {128, 64}, {128, 32}, {128, 16}, {128, 256}, {64, 128}, {64, 64}, {64, 256}};
  for (auto& s : shapes) {
    run_timing(s[0], s[1], 8, 64, 1);
    run_timing(s[0], s[1], 64, 128, 1);
    run_timing(s[0], s[1], 256, 16, 1);
  }
  return 0;
}
