// Developer tool (not part of the product): tcgen05.mma with the A operand in TMEM ("ts" form) and an
// MN-major B operand in shared memory -- layout check and issue-rate measurement on sm_100a.
//   D[m][n] = sum_k A[m][k] * B[n][k],  A: [128][K] fp16 written to TMEM with tcgen05.st (lane = m,
//   32-bit column c = {k=2c, k=2c+1}),  B: [N][K] fp16 stored MN-major: core matrix = 8 k-rows of
//   16 bytes (8 consecutive n), offset(n,k) = (k/8)*LBO + (n/8)*SBO + (k%8)*16 + (n%8)*2.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tc_probe2 tools/tc_probe2.cu
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
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
// A from TMEM, B from shared memory; issued by the calling thread
__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// whole warp executes; the instruction is predicated on an elected lane
__device__ __forceinline__ void umma_ts_elect(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p, q;\n\telect.sync _|q, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit_elect(uint32_t bar) {
  asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// MN-major, no swizzle
__host__ __device__ inline uint32_t b_offset(int n, int k, int N) { return (uint32_t)((k >> 3) * (N >> 3) * 128 + (n >> 3) * 128 + (k & 7) * 16 + (n & 7) * 2); }
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t smem_addr, int N) {
  const uint64_t sbo = 128 >> 4, lbo = (uint64_t)((N >> 3) * 128) >> 4;
  return (uint64_t)((smem_addr >> 4) & 0x3fff) | (lbo << 16) | (sbo << 32) | (1ull << 46);
}
// kind::f16: D fp32, A/B fp16, A K-major (TMEM), B MN-major
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) { return (1u << 4) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }

constexpr int A_COL = 256;     // TMEM column where the A operand starts (D at column 0)

// mode 0: correctness (n_mma = K/16 chained);  mode 1/2: timing with issue style (1 = single thread, 2 = elect)
__global__ void __launch_bounds__(128, 1) ts_kernel(const __half* A, const __half* B, float* D, long long* tout, int K, int N, int mode, int n_mma, int nacc) {
  extern __shared__ __align__(128) unsigned char sm[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_s)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  for (int e = tid; e < N * K; e += 128) {
    const int n = e / K, k = e - n * K;
    *reinterpret_cast<__half*>(sm + b_offset(n, k, N)) = B[e];
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_s;
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
  // A -> TMEM: thread = row m, 8 columns (16 k) per store
  for (int c0 = 0; c0 < K / 2; c0 += 8) {
    uint32_t v[8];
    for (int j = 0; j < 8; ++j) {
      const __half lo = A[tid * K + 2 * (c0 + j)], hi = A[tid * K + 2 * (c0 + j) + 1];
      v[j] = (uint32_t)__half_as_ushort(lo) | ((uint32_t)__half_as_ushort(hi) << 16);
    }
    tmem_st8(tmem + lane_base + A_COL + c0, v);
  }
  tmem_st_wait();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint64_t dB = make_desc_mn(smem_u32(sm), N);
  const uint32_t idesc = make_idesc(128, N);
  const uint64_t kstep = (uint64_t)((2 * (N >> 3) * 128) >> 4);      // 16 k = 2 k-groups
  long long t0 = 0, t1 = 0;
  if (mode == 0) {
    if (tid == 0) {
      for (int ks = 0; ks < K / 16; ++ks) umma_ts(tmem, tmem + A_COL + ks * 8, dB + ks * kstep, idesc, ks > 0);
      umma_commit(smem_u32(&bar));
    }
  } else if (mode == 1) {
    if (tid == 0) {
      t0 = clock64();
      for (int i = 0; i < n_mma; i += 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) umma_ts(tmem + (uint32_t)((j & 1) * 128), tmem + A_COL + j * 8, dB + j * kstep, idesc, 1);
      }
      umma_commit(smem_u32(&bar));
      t1 = clock64();
    }
  } else {
    if (warp == 0) {
      t0 = clock64();
      for (int i = 0; i < n_mma; i += 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) umma_ts_elect(tmem + (uint32_t)((j % nacc) * 32), tmem + A_COL + j * 8, dB + j * kstep, idesc, 1);
      }
      if (lane == 0) umma_commit(smem_u32(&bar));
      __syncwarp();
      t1 = clock64();
    }
  }
  if (mode == 3) {
    // bursts of n_mma MMAs separated by an idle gap of `nacc` cycles: issue+complete time of each burst
    if (warp == 0) {
      for (int rep = 0; rep < 6; ++rep) {
        const long long b0 = clock64();
#pragma unroll 1
        for (int i = 0; i < n_mma; i += 8) {
#pragma unroll
          for (int j = 0; j < 8; ++j) umma_ts_elect(tmem + (uint32_t)((j & 3) * 32), tmem + A_COL + j * 8, dB + j * kstep, idesc, 1);
        }
        const long long b1 = clock64();
        umma_commit_elect(smem_u32(&bar));
        __syncwarp();
        mbar_wait(smem_u32(&bar), rep & 1);
        const long long b2 = clock64();
        if (lane == 0) { tout[2 * rep] = b1 - b0; tout[2 * rep + 1] = b2 - b0; }
        while (clock64() - b2 < nacc) { }
      }
    }
    __syncthreads();
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
    return;
  }
  mbar_wait(smem_u32(&bar), 0);
  tc_fence_after();
  if (tid == 0 && mode != 0) { tout[0] = clock64() - t0; tout[1] = t1 - t0; }
  if (mode == 0) {
    for (int c0 = 0; c0 < N; c0 += 16) {
      uint32_t v[16];
      tmem_ld16(tmem + lane_base + c0, v);
      tmem_ld_wait();
      for (int j = 0; j < 16; ++j) D[tid * N + c0 + j] = __uint_as_float(v[j]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

int main() {
  srand(7);
  const int K = 128;
  for (int N : {16, 32}) {
    std::vector<__half> A(128 * K), B(N * K);
    std::vector<float> D(128 * N);
    for (auto& v : A) v = __float2half((float)(rand() % 2001 - 1000) / 1000.f);
    for (auto& v : B) v = __float2half((float)(rand() % 2001 - 1000) / 1000.f);
    __half *dA, *dB; float* dD; long long* dT;
    CK(cudaMalloc(&dA, A.size() * 2)); CK(cudaMalloc(&dB, B.size() * 2)); CK(cudaMalloc(&dD, D.size() * 4)); CK(cudaMalloc(&dT, 128));
    CK(cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice));
    const size_t smem = 64 * 1024;
    CK(cudaFuncSetAttribute(ts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ts_kernel<<<1, 128, smem>>>(dA, dB, dD, dT, K, N, 0, 0, 1);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
    double maxerr = 0;
    for (int m = 0; m < 128; ++m)
      for (int n = 0; n < N; ++n) {
        double s = 0;
        for (int k = 0; k < K; ++k) s += (double)__half2float(A[m * K + k]) * (double)__half2float(B[n * K + k]);
        maxerr = std::fmax(maxerr, std::fabs(s - (double)D[m * N + n]));
      }
    printf("TS N=%3d K=%d : max |D - exact| = %.3e  (%s)\n", N, K, maxerr, maxerr < 1e-4 ? "layout OK" : "LAYOUT WRONG");
    for (int nacc : {1, 2, 4})
      for (int n_mma : {8, 32, 256}) {
        long long h[2];
        for (int rep = 0; rep < 2; ++rep) { ts_kernel<<<1, 128, smem>>>(dA, dB, dD, dT, K, N, 2, n_mma, nacc); CK(cudaDeviceSynchronize()); }
        CK(cudaMemcpy(h, dT, 16, cudaMemcpyDeviceToHost));
        printf("TS-TIME N=%3d accumulators=%d n_mma=%3d : total %6lld cyc (%.1f/mma)  issue %6lld cyc (%.1f/mma)\n", N, nacc, n_mma, h[0], (double)h[0] / n_mma, h[1], (double)h[1] / n_mma);
      }
    if (false) {
      for (int gap : {0, 200, 500, 1000, 2000, 5000}) {
        long long h[12];
        ts_kernel<<<1, 128, smem>>>(dA, dB, dD, dT, K, N, 3, 32, gap); CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(h, dT, 96, cudaMemcpyDeviceToHost));
        printf("TS-GAP N=32 burst=32 idle gap %5d cyc : issue/complete per burst:", gap);
        for (int r = 0; r < 6; ++r) printf(" %lld/%lld", h[2 * r], h[2 * r + 1]);
        printf("\n");
      }
    }
    cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(dT);
  }
  return 0;
}
