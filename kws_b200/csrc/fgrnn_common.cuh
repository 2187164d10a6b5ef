// Shared device helpers and host-side plumbing for the FastGRNN engine (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

#include "fastgrnn_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "kws_b200 kernels are written for sm_100a only"
#endif

namespace fgrnn {

// ---------------------------------------------------------------------------------------------
// host-side error plumbing
// ---------------------------------------------------------------------------------------------
void set_error_detail(const char* fmt, ...);
void count_launch(int n = 1);

// Tuning / test overrides of the launchers.  Each key takes its default from the environment variable of the same name,
// read ONCE per process (no getenv on the launch path); fgrnn_debug_set_tuning changes it afterwards (tests).
enum TuneKey { TUNE_TC_NS = 0, TUNE_TC_NT, TUNE_TC_BR_NS, TUNE_TC_WIDE, TUNE_FAST_NL, TUNE_SMEM_CFG, TUNE_TC_LR, TUNE_TC_ALT, TUNE_TC_ACC2, TUNE_TC_BWD_FUSED, TUNE_TC_VR, TUNE_COUNT };
constexpr int TUNE_UNSET = -2147483647 - 1;
int tuning(TuneKey key);          // TUNE_UNSET when neither the environment nor a test set it

#define FGRNN_CUDA_TRY(expr)                                                              \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess) {                                                              \
      ::fgrnn::set_error_detail("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),   \
                                __FILE__, __LINE__);                                      \
      return FGRNN_ERR_CUDA;                                                              \
    }                                                                                     \
  } while (0)

#define FGRNN_LAUNCH_CHECK(name)                                                          \
  do {                                                                                    \
    cudaError_t _e = cudaGetLastError();                                                  \
    if (_e != cudaSuccess) {                                                              \
      ::fgrnn::set_error_detail("launch of %s failed: %s", name, cudaGetErrorString(_e)); \
      return FGRNN_ERR_CUDA;                                                              \
    }                                                                                     \
    ::fgrnn::count_launch();                                                              \
  } while (0)

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Bump allocator over the caller-provided workspace (sizes computed identically by the
// *_workspace_bytes queries, which run it with base == nullptr).
struct Carver {
  char* base;
  size_t off;
  explicit Carver(void* b) : base(static_cast<char*>(b)), off(0) {}
  template <typename T>
  T* take(size_t count) {
    off = align_up(off, 256);
    T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += count * sizeof(T);
    return p;
  }
  size_t total() const { return align_up(off, 256); }
};

// ---------------------------------------------------------------------------------------------
// nonlinearities (rnn.py:40-67) and their derivatives expressed on the OUTPUT value
// (cuda/fastgrnn_cuda_kernel.cu:27-40 for sigmoid/relu/tanh; clamps: slope inside the window)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float sigmoid_f(float a) { return 1.0f / (1.0f + expf(-a)); }

// Fast variants for the persistent kernels' default (sigmoid gate, tanh update) epilogue:
// MUFU.EX2 / MUFU.RCP based, <= ~2 ulp; tanh switches to an odd minimax polynomial (max rel
// error 1.1e-7) below 0.55 where 1 - 2/(1+e^{2x}) would cancel.  Margins vs the oracle are
// asserted by the GPU parity tests.
__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
// clamps that keep NaN (fmaxf / fminf return the other operand): a NaN pre-activation must stay NaN, as in the reference
__device__ __forceinline__ float fmax_nan(float a, float b) { float y; asm("max.NaN.f32 %0, %1, %2;" : "=f"(y) : "f"(a), "f"(b)); return y; }
__device__ __forceinline__ float fmin_nan(float a, float b) { float y; asm("min.NaN.f32 %0, %1, %2;" : "=f"(y) : "f"(a), "f"(b)); return y; }
__device__ __forceinline__ float sigmoid_fast(float a) { return rcp_approx(1.0f + ex2_approx(-1.4426950408889634f * a)); }
__device__ __forceinline__ float tanh_fast(float b) {
  const float e = ex2_approx(2.8853900817779268f * b);
  const float big = fmaf(-2.0f, rcp_approx(1.0f + e), 1.0f);
  const float s = b * b;
  float t = fmaf(s, 0.016458360478281975f, -0.05268390104174614f);
  t = fmaf(t, s, 0.13320937752723694f);
  t = fmaf(t, s, -0.33332955837249756f);
  const float small = fmaf(t * s, b, b);
  return fabsf(b) < 0.55f ? small : big;
}

template <int NL>
__device__ __forceinline__ float act(float a) {
  if (NL == FGRNN_NL_SIGMOID) return sigmoid_f(a);
  if (NL == FGRNN_NL_TANH) return tanhf(a);
  if (NL == FGRNN_NL_RELU) return fmaxf(a, 0.0f);
  if (NL == FGRNN_NL_QUANT_TANH) return fmaxf(fminf(a, 1.0f), -1.0f);
  if (NL == FGRNN_NL_QUANT_SIGM) return fmaxf(fminf((a + 1.0f) / 2.0f, 1.0f), 0.0f);
  if (NL == FGRNN_NL_QUANT_SIGM4) return fmaxf(fminf((a + 2.0f) / 4.0f, 1.0f), 0.0f);
  return a;
}

__device__ __forceinline__ float act_rt(int nl, float a) {
  switch (nl) {
    case FGRNN_NL_SIGMOID: return act<FGRNN_NL_SIGMOID>(a);
    case FGRNN_NL_TANH: return act<FGRNN_NL_TANH>(a);
    case FGRNN_NL_RELU: return act<FGRNN_NL_RELU>(a);
    case FGRNN_NL_QUANT_TANH: return act<FGRNN_NL_QUANT_TANH>(a);
    case FGRNN_NL_QUANT_SIGM: return act<FGRNN_NL_QUANT_SIGM>(a);
    default: return act<FGRNN_NL_QUANT_SIGM4>(a);
  }
}

// derivative d act / d a as a function of y = act(a)
__device__ __forceinline__ float dact_rt(int nl, float y) {
  switch (nl) {
    case FGRNN_NL_SIGMOID: return y * (1.0f - y);
    case FGRNN_NL_TANH: return 1.0f - y * y;
    case FGRNN_NL_RELU: return y > 0.0f ? 1.0f : 0.0f;
    case FGRNN_NL_QUANT_TANH: return (y > -1.0f && y < 1.0f) ? 1.0f : 0.0f;
    case FGRNN_NL_QUANT_SIGM: return (y > 0.0f && y < 1.0f) ? 0.5f : 0.0f;
    default: return (y > 0.0f && y < 1.0f) ? 0.25f : 0.0f;
  }
}

__device__ __forceinline__ float load_x(const void* x, int64_t idx, int dtype) {
  if (dtype == FGRNN_BF16) return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(x)[idx]);
  return reinterpret_cast<const float*>(x)[idx];
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---------------------------------------------------------------------------------------------
// canonical problem view used by every kernel family.  "Canonical" weights are row-major
// [K][N] in the oracle orientation (pre = A . Wc), independent of FgrnnProblem.weight_layout.
// ---------------------------------------------------------------------------------------------
struct Dims {
  int B, T, I, H, rW, rU;
  int gate_nl, update_nl, x_dtype;
};

}  // namespace fgrnn
