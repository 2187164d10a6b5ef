// Developer tool (not part of the product): shared-memory wavefronts per LDS for lane->address
// patterns, read from ncu counters (l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld /
// smsp__sass_inst_executed_op_shared_ld).  Loads are volatile inline PTX so nothing is merged.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ float lds128(unsigned a) { float x, y, z, w; asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(x), "=f"(y), "=f"(z), "=f"(w) : "r"(a)); return x + y + z + w; }
__device__ __forceinline__ float lds64(unsigned a) { float x, y; asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(x), "=f"(y) : "r"(a)); return x + y; }
__device__ __forceinline__ float lds32(unsigned a) { float x; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x) : "r"(a)); return x; }

template <int VEC, int PAT>
__global__ void k(float* out, int iters) {
  __shared__ __align__(128) float s[8192];
  const int lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 8192; i += blockDim.x) s[i] = i;
  __syncthreads();
  const int q = lane >> 3, l8 = lane & 7;
  int chunk;   // index of the VEC-float element this lane reads
  switch (PAT) {
    case 0: chunk = 0; break;                       // 1 distinct (full broadcast)
    case 1: chunk = q; break;                       // 1 per quarter, 4 total contiguous
    case 2: chunk = q * 33; break;                  // 1 per quarter, 4 total, padded rows (stride 132 floats for VEC=4)
    case 3: chunk = l8; break;                      // 8 per quarter, same across quarters
    case 4: chunk = lane; break;                    // 32 distinct contiguous
    case 5: chunk = l8 & 3; break;                  // 4 per quarter, same across quarters
    case 6: chunk = l8 & 1; break;                  // 2 per quarter, same across quarters
    case 7: chunk = (l8 & 1) + 2 * q; break;        // 2 per quarter, 8 total contiguous
    case 8: chunk = lane & 15; break;               // 16 distinct, halves same
    case 9: chunk = lane >> 1; break;               // 16 distinct, lane pairs share
    case 10: chunk = lane >> 2; break;              // 8 distinct: 2 per quarter
    case 11: chunk = (l8 & 3) + 4 * (q & 1); break; // 4 per quarter, 8 total, quarters 0/2 and 1/3 same
    case 12: chunk = (l8 >> 1) + 4 * q; break;      // 4 per quarter (pairs), 16 total
    case 13: chunk = (l8 & 1) * 33 + 2 * 33 * q; break;  // 2 padded rows per quarter, 8 rows total
    case 14: chunk = (lane & 1) * 33; break;        // 2 distinct padded rows across whole warp
    default: chunk = 0;
  }
  const unsigned base = (unsigned)__cvta_generic_to_shared(s) + chunk * VEC * 4;
  float acc = 0.f;
  for (int it = 0; it < iters; ++it) {
    const unsigned a = base + (it & 3) * 8192;
    if (VEC == 4) acc += lds128(a); else if (VEC == 2) acc += lds64(a); else acc += lds32(a);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int VEC, int PAT>
void run() {
  float* out;
  cudaMalloc(&out, 4096 * 4);
  k<VEC, PAT><<<1, 128>>>(out, 1000);
  cudaDeviceSynchronize();
  cudaFree(out);
}

int main() {
  run<4, 0>(); run<4, 1>(); run<4, 2>(); run<4, 3>(); run<4, 4>(); run<4, 5>(); run<4, 6>(); run<4, 7>();
  run<4, 8>(); run<4, 9>(); run<4, 10>(); run<4, 11>(); run<4, 12>(); run<4, 13>(); run<4, 14>();
  run<2, 0>(); run<2, 3>(); run<2, 4>(); run<2, 8>(); run<2, 9>(); run<2, 10>(); run<2, 5>();
  run<1, 0>(); run<1, 3>(); run<1, 4>(); run<1, 10>();
  printf("done\n");
  return 0;
}
