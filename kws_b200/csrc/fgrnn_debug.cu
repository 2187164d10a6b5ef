// Diagnostic entry point of libfastgrnn_b200.so: poison the on-chip state that survives between kernel launches.
//
// Tensor memory and shared memory are not cleared between launches.  A kernel that (through a missing fence or
// ordering edge) consumed an operand before it was written would therefore be saved by whatever its predecessor left
// in place -- with identical weights, the correct values -- on every launch except the first one on an SM.  The
// first-launch tests (tests/test_gpu_first_launch.py, tools/first_launch_probe.cu) call this between launches so that
// EVERY launch starts from a NaN pattern: a stale read then shows up as NaN in the output instead of hiding.
#include "fgrnn_kernels.cuh"
#include "fgrnn_tc_common.cuh"

namespace fgrnn {

constexpr uint32_t kPoisonWord = 0x7fc07fc0u;     // NaN as an fp16 pair, a bf16 pair and an fp32 value

__global__ void __launch_bounds__(128, 1) poison_onchip_kernel(int smem_words) {
  extern __shared__ uint32_t poison_sm[];
  __shared__ uint32_t tmem_base;
  for (int i = threadIdx.x; i < smem_words; i += blockDim.x) poison_sm[i] = kPoisonWord;
  if (threadIdx.x < 32) tmem_alloc(smem_u32(&tmem_base), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t base = tmem_base + ((uint32_t)((threadIdx.x >> 5) * 32) << 16);
  uint32_t v[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = kPoisonWord;
  for (int c = 0; c < 512; c += 8) tmem_st8(base + c, v);
  tmem_st_wait();
  __nanosleep(20000);                              // hold the SM so that the grid spreads over all of them
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem_base, 512);
}

}  // namespace fgrnn

extern "C" int fgrnn_debug_poison_onchip(int device, void* stream) {
  using namespace fgrnn;
  int prev = -1, sms = 0;
  FGRNN_CUDA_TRY(cudaGetDevice(&prev));
  FGRNN_CUDA_TRY(cudaSetDevice(device));
  FGRNN_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
  const int smem = 227 * 1024 - 64;                // with the static word: one CTA per SM
  FGRNN_CUDA_TRY(cudaFuncSetAttribute(poison_onchip_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  poison_onchip_kernel<<<2 * sms, 128, smem, static_cast<cudaStream_t>(stream)>>>(smem / 4);
  FGRNN_LAUNCH_CHECK("poison_onchip_kernel");
  if (prev >= 0) cudaSetDevice(prev);
  return FGRNN_OK;
}
