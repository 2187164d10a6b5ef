// Classifier head of the keyword spotter, forward + loss + backward in ONE launch (model.py:227-231:
// hidden2keyword = Linear(H -> classes) on the LAST hidden state, log_softmax, NLLLoss as trainClassifier.py:236 applies it):
//     logits = h_T . W^T + b;   logp = log_softmax(logits);   loss = -mean_b logp[b, label_b]
//     dlogits = (softmax - onehot) / B;   dW = dlogits^T . h_T;   db = sum_b dlogits;   dh_T = dlogits . W
// The trainer's torch version of this is ~15 small launches per step (sgemm + epilogue, softmax, nll, their backward,
// fills, reductions: ~25 % of the C3 step); here it is one kernel of B/64 CTAs whose per-CTA partial sums are added in
// CTA order by the last CTA to finish (deterministic, no floating-point atomics).
#include "fgrnn_kernels.cuh"

namespace fgrnn {

constexpr int HD_ROWS = 64, HD_THREADS = 256, HD_MAXC = 16;

struct HeadArgs {
  const float* h; int64_t h_stride;          // [B][H] rows (the last time step of the hidden states)
  const float* W; const float* b;            // nn.Linear: W [C][H], b [C]
  const int64_t* labels;                     // [B]
  float* loss; float* dW; float* db;         // outputs: scalar, [C][H], [C]
  float* dh; int64_t dh_stride;              // [B][H] rows: gradient w.r.t. h_T (already scaled by 1/B)
  float* logp;                               // optional [B][C]
  float* partial; unsigned* counter;         // workspace: [nCTA][C*H + C + 1], one counter (zero on entry, zero on exit)
  int B, H, C;
  float inv_b;
};

// Every phase keeps several short, independent FMA chains per thread (the first version had one 128- / 64-step chain per
// thread and phase and took 23 us for a single CTA: latency bound with eight warps per SM).
__global__ void __launch_bounds__(HD_THREADS) head_nll_kernel(const HeadArgs a) {
  extern __shared__ float hsm[];
  const int H = a.H, C = a.C, Hp = H + 4;          // padded pitch: rows (and classes) land in different banks
  float* Ws = hsm;                                 // [C][Hp]
  float* hs = Ws + C * Hp;                         // [ROWS][Hp]
  float* dl = hs + HD_ROWS * Hp;                   // [ROWS][HD_MAXC]: logits, then dlogits
  float* red = dl + HD_ROWS * HD_MAXC;             // [ROWS] loss terms
  __shared__ bool last;
  const int tid = threadIdx.x, row0 = blockIdx.x * HD_ROWS;
  const int H4 = H >> 2;
  for (int i = tid; i < C * H; i += HD_THREADS) {       // scalar loads: W may sit at any 4-byte offset of a flat parameter buffer
    const int c = i / H, k = i - c * H;
    Ws[c * Hp + k] = __ldg(a.W + i);
  }
  for (int i = tid; i < HD_ROWS * H4; i += HD_THREADS) {
    const int r = i / H4, k4 = i - r * H4;
    const float4 v = row0 + r < a.B ? __ldg(reinterpret_cast<const float4*>(a.h + (int64_t)(row0 + r) * a.h_stride) + k4) : make_float4(0.f, 0.f, 0.f, 0.f);
    *reinterpret_cast<float4*>(hs + r * Hp + k4 * 4) = v;
  }
  __syncthreads();
  // ---- logits: item = (row, class); up to four items per thread, their chains interleaved
  {
    constexpr int NI = (HD_ROWS * HD_MAXC + HD_THREADS - 1) / HD_THREADS;     // 4
    const int nitem = HD_ROWS * C;
    const float* hp[NI]; const float* wp[NI]; float acc[NI]; int rr[NI], cc[NI];
#pragma unroll
    for (int j = 0; j < NI; ++j) {
      const int it = tid + j * HD_THREADS, live = it < nitem;
      rr[j] = live ? it / C : 0; cc[j] = live ? it - rr[j] * C : 0;
      hp[j] = hs + rr[j] * Hp; wp[j] = Ws + cc[j] * Hp;
      acc[j] = live ? __ldg(a.b + cc[j]) : 0.f;
    }
    for (int k = 0; k < H; k += 4) {
#pragma unroll
      for (int j = 0; j < NI; ++j) {
        const float4 hv = *reinterpret_cast<const float4*>(hp[j] + k), wv = *reinterpret_cast<const float4*>(wp[j] + k);
        acc[j] = fmaf(hv.x, wv.x, acc[j]); acc[j] = fmaf(hv.y, wv.y, acc[j]); acc[j] = fmaf(hv.z, wv.z, acc[j]); acc[j] = fmaf(hv.w, wv.w, acc[j]);
      }
    }
#pragma unroll
    for (int j = 0; j < NI; ++j)
      if (tid + j * HD_THREADS < nitem) dl[rr[j] * HD_MAXC + cc[j]] = acc[j];
  }
  __syncthreads();
  // ---- log-softmax, loss term and dlogits: one thread per row (C <= 16)
  if (tid < HD_ROWS) {
    const int r = tid;
    const bool live = row0 + r < a.B;
    const int lab = live ? (int)a.labels[row0 + r] : -1;
    float lg[HD_MAXC], mx = -INFINITY;
#pragma unroll
    for (int c = 0; c < HD_MAXC; ++c) { lg[c] = c < C ? dl[r * HD_MAXC + c] : -INFINITY; mx = fmaxf(mx, lg[c]); }
    float se = 0.f;
#pragma unroll
    for (int c = 0; c < HD_MAXC; ++c) se += c < C ? expf(lg[c] - mx) : 0.f;
    const float lse = mx + logf(se);
    float lt = 0.f;
#pragma unroll
    for (int c = 0; c < HD_MAXC; ++c) {
      if (c < C) {
        const float lp = lg[c] - lse;
        if (a.logp && live) a.logp[(size_t)(row0 + r) * C + c] = lp;
        dl[r * HD_MAXC + c] = live ? (expf(lp) - (c == lab ? 1.f : 0.f)) * a.inv_b : 0.f;
        if (c == lab) lt = -lp;
      } else {
        dl[r * HD_MAXC + c] = 0.f;                         // classes past C: read (and discarded) by the dW loop below
      }
    }
    red[r] = live ? lt : 0.f;
  }
  __syncthreads();
  // ---- dh = dlogits . W: thread = one 4-unit column group, eight rows at once
  if (a.dh) {
    for (int k4 = tid % 32; k4 < H4; k4 += 32) {
      for (int r0 = tid / 32; r0 < HD_ROWS; r0 += HD_THREADS / 32) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int c = 0; c < C; ++c) {
          const float dv = dl[r0 * HD_MAXC + c];
          const float4 wv = *reinterpret_cast<const float4*>(Ws + c * Hp + k4 * 4);
          acc.x = fmaf(dv, wv.x, acc.x); acc.y = fmaf(dv, wv.y, acc.y); acc.z = fmaf(dv, wv.z, acc.z); acc.w = fmaf(dv, wv.w, acc.w);
        }
        if (row0 + r0 < a.B) *(reinterpret_cast<float4*>(a.dh + (int64_t)(row0 + r0) * a.dh_stride) + k4) = acc;
      }
    }
  }
  // ---- per-CTA partials: dW[c][k] = sum_r dl[r][c] * h[r][k] (thread = unit k x a class parity, all its classes at once), db[c], loss
  float* part = a.partial + (size_t)blockIdx.x * (C * H + C + 1);
  for (int k = tid % 128; k < H; k += 128) {
    const int c0 = tid / 128;                          // 0 or 1: this thread takes classes c0, c0 + 2, ...
    float acc[HD_MAXC / 2];
#pragma unroll
    for (int j = 0; j < HD_MAXC / 2; ++j) acc[j] = 0.f;
    for (int r = 0; r < HD_ROWS; ++r) {
      const float hv = hs[r * Hp + k];
#pragma unroll
      for (int j = 0; j < HD_MAXC / 2; ++j) acc[j] = fmaf(dl[r * HD_MAXC + c0 + 2 * j], hv, acc[j]);
    }
#pragma unroll
    for (int j = 0; j < HD_MAXC / 2; ++j)
      if (c0 + 2 * j < C) part[(c0 + 2 * j) * H + k] = acc[j];
  }
  if (tid < C) {
    float s = 0.f;
    for (int r = 0; r < HD_ROWS; ++r) s += dl[r * HD_MAXC + tid];
    part[C * H + tid] = s;
  }
  if (tid == 0) {
    float s = 0.f;
    for (int r = 0; r < HD_ROWS; ++r) s += red[r];
    part[C * H + C] = s * a.inv_b;
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) last = atomicAdd(a.counter, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!last) return;
  __threadfence();
  const int n = C * H + C + 1;
  for (int i = tid; i < n; i += HD_THREADS) {
    // fixed order (run-to-run identical), sixteen independent loads in flight per thread
    float s = 0.f;
    for (unsigned g0 = 0; g0 < gridDim.x; g0 += 16) {
      float v[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = g0 + j < gridDim.x ? __ldcg(a.partial + (size_t)(g0 + j) * n + i) : 0.f;
#pragma unroll
      for (int j = 0; j < 16; ++j) s += v[j];
    }
    if (i < C * H) a.dW[i] = s;
    else if (i < C * H + C) a.db[i - C * H] = s;
    else a.loss[0] = s;
  }
  if (tid == 0) *a.counter = 0u;
}

// Plain SGD over ONE flat parameter buffer and ONE flat gradient bucket (trainClassifier.py:240 optimizer.step() on
// torch.optim.SGD without momentum): p <- p - lr * (grad_scale * g); grad_scale = 1 / world size after the all-reduce.
__global__ void sgd_flat_kernel(float* __restrict__ p, const float* __restrict__ g, int64_t n, float lr, float grad_scale) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = fmaf(-lr, g[i] * grad_scale, p[i]);
}

}  // namespace fgrnn

using namespace fgrnn;

extern "C" int fgrnn_sgd_flat(float* params, const float* grads, int64_t n, float lr, float grad_scale, int32_t device, void* stream) {
  if (!params || !grads) { set_error_detail("sgd: params / grads is NULL"); return FGRNN_ERR_NULL; }
  if (n <= 0) return FGRNN_OK;
  int prev = -1;
  FGRNN_CUDA_TRY(cudaGetDevice(&prev));
  FGRNN_CUDA_TRY(cudaSetDevice(device));
  sgd_flat_kernel<<<(unsigned)((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(params, grads, n, lr, grad_scale);
  FGRNN_LAUNCH_CHECK("sgd_flat_kernel");
  if (prev >= 0) cudaSetDevice(prev);
  return FGRNN_OK;
}

extern "C" size_t fgrnn_head_workspace_bytes(int32_t B, int32_t H, int32_t C) {
  const size_t nb = (size_t)((B + HD_ROWS - 1) / HD_ROWS);
  return 256 + nb * ((size_t)C * H + C + 1) * sizeof(float);
}

extern "C" int fgrnn_head_nll(const float* h, int64_t h_stride, const float* W, const float* b, const int64_t* labels,
                              float* loss, float* dW, float* db, float* dh, int64_t dh_stride, float* logp,
                              void* workspace, size_t workspace_bytes, int32_t B, int32_t H, int32_t C, int32_t device, void* stream) {
  if (!h || !W || !b || !labels || !loss || !dW || !db || !workspace) { set_error_detail("head: a required pointer is NULL"); return FGRNN_ERR_NULL; }
  if (B < 1 || H < 4 || (H & 3) || C < 1 || C > HD_MAXC) { set_error_detail("head: B=%d H=%d (multiple of 4) C=%d (<= %d)", B, H, C, HD_MAXC); return FGRNN_ERR_SHAPE; }
  if (workspace_bytes < fgrnn_head_workspace_bytes(B, H, C)) { set_error_detail("head: workspace too small"); return FGRNN_ERR_WORKSPACE; }
  if ((reinterpret_cast<uintptr_t>(h) & 15) || (h_stride & 3) || (dh && ((reinterpret_cast<uintptr_t>(dh) & 15) || (dh_stride & 3)))) {
    set_error_detail("head: h / dh rows must be 16-byte aligned"); return FGRNN_ERR_ALIGN;
  }
  int prev = -1;
  FGRNN_CUDA_TRY(cudaGetDevice(&prev));
  FGRNN_CUDA_TRY(cudaSetDevice(device));
  HeadArgs a{};
  a.h = h; a.h_stride = h_stride; a.W = W; a.b = b; a.labels = labels; a.loss = loss; a.dW = dW; a.db = db;
  a.dh = dh; a.dh_stride = dh_stride; a.logp = logp;
  a.counter = static_cast<unsigned*>(workspace);                      // first 256 bytes: the CTA counter (zero between launches)
  a.partial = reinterpret_cast<float*>(static_cast<char*>(workspace) + 256);
  a.B = B; a.H = H; a.C = C; a.inv_b = 1.0f / (float)B;
  const size_t smem = ((size_t)C * (H + 4) + (size_t)HD_ROWS * (H + 4) + HD_ROWS * HD_MAXC + HD_ROWS) * sizeof(float);
  FGRNN_CUDA_TRY(cudaFuncSetAttribute(head_nll_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  head_nll_kernel<<<(unsigned)((B + HD_ROWS - 1) / HD_ROWS), HD_THREADS, smem, static_cast<cudaStream_t>(stream)>>>(a);
  FGRNN_LAUNCH_CHECK("head_nll_kernel");
  if (prev >= 0) cudaSetDevice(prev);
  return FGRNN_OK;
}
