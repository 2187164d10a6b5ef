/*
 * fastgrnn_b200.h -- C ABI of the B200-native FastGRNN recurrence engine.
 *
 * This is the drop-in boundary for the reference's `fastgrnn_cuda` extension:
 * the four pybind11 entry points of /root/reference/cuda/fastgrnn_cuda.cpp:235-240
 *
 *     forward          cuda/fastgrnn_cuda.cpp:73-107   -> fgrnn_forward  (T == 1)
 *     backward         cuda/fastgrnn_cuda.cpp:109-145  -> fgrnn_backward (T == 1)
 *     forward_unroll   cuda/fastgrnn_cuda.cpp:147-180  -> fgrnn_forward
 *     backward_unroll  cuda/fastgrnn_cuda.cpp:182-232  -> fgrnn_backward
 *
 * and for the Python unroller they sit under (rnn.py:574-668 BaseRNN.forward,
 * rnn.py:273-297 FastGRNNCell.forward).  The reference passes torch::Tensor
 * objects and allocates its outputs; this ABI is plain C: raw device pointers,
 * sizes, element strides and enums, an explicit CUDA stream, integer return
 * codes.  The caller allocates every output and the workspace; the library
 * allocates nothing, keeps no per-call state and never synchronises the host.
 *
 * Conventions
 *   - all matrices are fp32; `x` may be fp32 or bf16 (x_dtype); state is fp32.
 *   - strides are in ELEMENTS; the innermost (feature / hidden) stride is 1.
 *     (T,B,F), (B,T,F) and permuted views are all expressed by (stride_b, stride_t).
 *   - weight_layout selects the reference's two parameter layouts:
 *       FGRNN_LAYOUT_IH : FastGRNNCell   (rnn.py:246-256)  W[I,H]  U[H,H]  W1[I,rW] W2[rW,H] U1[H,rU] U2[rU,H],  pre = x.W  + h.U
 *       FGRNN_LAYOUT_HI : FastGRNNCUDA   (rnn.py:782-805)  W[H,I]  U[H,H]  W1[rW,I] W2[H,rW] U1[rU,H] U2[H,rU],  pre = x.W^T + h.U^T
 *     rW == 0 / rU == 0 selects the full-rank W / U (the reference tests w1.size(0)==0,
 *     cuda/fastgrnn_cuda.cpp:88-99).  Gradients are written in the same layout.
 *   - nonlinearity enums 0..2 are the reference's {"sigmoid":0,"relu":1,"tanh":2}
 *     (rnn.py:478, rnn.py:751); 3..5 are the remaining gen_nonlinearity names (rnn.py:53-60).
 *   - math (rnn.py:289-295):  pre = wComp + uComp;  z = gate(pre + bias_gate);
 *     c = update(pre + bias_update);  h' = z*h + (sigmoid(zeta)*(1-z) + sigmoid(nu))*c.
 *     Backward follows cuda/fastgrnn_cuda_kernel.cu:109-118,537-556 with the correct
 *     tanh-gate derivative (the reference's unrolled kernel uses d_sigmoid there, cu:519-521).
 */
#ifndef FASTGRNN_B200_H_
#define FASTGRNN_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FGRNN_ABI_VERSION 3

#if defined(__GNUC__)
#define FGRNN_API __attribute__((visibility("default")))
#else
#define FGRNN_API
#endif

/* return codes */
enum {
  FGRNN_OK = 0,
  FGRNN_ERR_NULL = 1,        /* a required pointer is NULL */
  FGRNN_ERR_SHAPE = 2,       /* a dimension is out of the supported range */
  FGRNN_ERR_ENUM = 3,        /* unknown nonlinearity / layout / dtype */
  FGRNN_ERR_ALIGN = 4,       /* pointer or stride alignment requirement violated */
  FGRNN_ERR_WORKSPACE = 5,   /* workspace missing or too small */
  FGRNN_ERR_CUDA = 6,        /* a CUDA runtime call or kernel launch failed */
  FGRNN_ERR_DEVICE = 7,      /* device is not an sm_100 part */
  FGRNN_ERR_VERSION = 8      /* desc.abi_version mismatch */
};

/* nonlinearities: 0..2 as rnn.py:478 / rnn.py:751, 3..5 rnn.py:53-60 */
enum {
  FGRNN_NL_SIGMOID = 0,
  FGRNN_NL_RELU = 1,
  FGRNN_NL_TANH = 2,
  FGRNN_NL_QUANT_TANH = 3,
  FGRNN_NL_QUANT_SIGM = 4,
  FGRNN_NL_QUANT_SIGM4 = 5
};

enum { FGRNN_LAYOUT_IH = 0, FGRNN_LAYOUT_HI = 1 };
enum { FGRNN_F32 = 0, FGRNN_BF16 = 1 };

/* kernel families (fgrnn_*_plan reports which one a descriptor selects) */
enum {
  FGRNN_PATH_GENERIC = 0,    /* any shape: weights streamed through L1/L2 */
  FGRNN_PATH_SMEM = 1,       /* persistent FFMA kernel, weights resident in shared memory */
  FGRNN_PATH_TCGEN05 = 2,    /* tcgen05/TMEM kernels: fused (H = 128, I <= 64) or hoisted x.W GEMM + recurrence (H = 128 / 256, I <= 256) */
  FGRNN_PATH_LOWRANK = 3     /* forward only: persistent FFMA kernel for W1.W2 / U1.U2 with H = 256 (backward: generic) */
};

/* Problem description shared by forward and backward. */
typedef struct FgrnnProblem {
  int32_t abi_version;       /* FGRNN_ABI_VERSION */
  int32_t device;            /* CUDA device ordinal the pointers live on */
  int32_t B, T, I, H;        /* batch, steps, input features, hidden units */
  int32_t rW, rU;            /* low-rank ranks, 0 = full rank */
  int32_t gate_nl;           /* FGRNN_NL_* for z  (rnn.py:290) */
  int32_t update_nl;         /* FGRNN_NL_* for c  (rnn.py:292); the reference CUDA path fixes tanh (cu:57) */
  int32_t weight_layout;     /* FGRNN_LAYOUT_* */
  int32_t x_dtype;           /* FGRNN_F32 | FGRNN_BF16 */
  int32_t force_path;        /* -1 = auto, else FGRNN_PATH_* (tests / benchmarks) */
  int32_t reserved0;
  /* parameters (device pointers; the unused rank variant may be NULL) */
  const float* W;  const float* U;
  const float* W1; const float* W2;
  const float* U1; const float* U2;
  const float* bias_gate;    /* [1,H] */
  const float* bias_update;  /* [1,H] */
  const float* zeta;         /* [1,1] raw (pre-sigmoid) */
  const float* nu;           /* [1,1] raw (pre-sigmoid) */
  /* input sequence */
  const void* x;  int64_t x_stride_b, x_stride_t;          /* [B,T,I] by strides */
  const float* h0;           /* [B,H] contiguous, NULL = zeros (rnn.py:588-591, rnn.py:816-818) */
  /* optional per-unit factors on the two pre-activations, [1,H] each, NULL = 1:
       z = gate(gate_scale * pre + bias_gate),  c = update(update_scale * pre + bias_update).
     This is what the reference's FastGRNNBatchNormCell computes in eval mode (rnn.py:377-408) once its four
     BatchNorm1d layers are folded: bn_w / bn_u scale the columns of W / U, bn_gate / bn_update become these
     factors plus shifts of the two biases (kws_b200/rnn.py FastGRNNBatchNorm does the folding).  Forward only. */
  const float* gate_scale;
  const float* update_scale;
} FgrnnProblem;

/* forward (replaces forward / forward_unroll, cuda/fastgrnn_cuda.cpp:73,147) */
typedef struct FgrnnForward {
  FgrnnProblem p;
  float* out;  int64_t out_stride_b, out_stride_t;         /* all T hidden states; may be NULL if h_last is set */
  float* h_last;             /* optional [B,H]: h_T, for chunked streaming / last-state-only callers */
  float* save_z;             /* optional [T,B,H] contiguous: z_s       (cu:340, returned by forward_unroll) */
  float* save_c;             /* optional [T,B,H] contiguous: h_prime_s (cu:341) */
  void* workspace; size_t workspace_bytes;
} FgrnnForward;

/* backward (replaces backward / backward_unroll, cuda/fastgrnn_cuda.cpp:109,182) */
typedef struct FgrnnBackward {
  FgrnnProblem p;
  const float* grad_h; int64_t grad_stride_b, grad_stride_t;   /* dL/dh_t, all T */
  const float* hs;     int64_t hs_stride_b, hs_stride_t;       /* hidden states from forward */
  const float* z_s;          /* [T,B,H] contiguous, saved by forward (required) */
  const float* c_s;          /* [T,B,H] contiguous, saved by forward (required) */
  /* outputs; every pointer is optional except where noted. Parameter grads are OVERWRITTEN
     (not accumulated), in p.weight_layout, shapes as the parameters. */
  float* d_x;  int64_t dx_stride_b, dx_stride_t;           /* fp32, NULL = skip (layer-0 input) */
  float* d_W;  float* d_U;  float* d_W1; float* d_W2; float* d_U1; float* d_U2;
  float* d_bias_gate; float* d_bias_update;                /* [1,H] */
  float* d_zeta; float* d_nu;                              /* [1,1] */
  float* d_h0;                                             /* [B,H] */
  void* workspace; size_t workspace_bytes;
  /* ABI 3.  First time step that has an upstream gradient: grad_h holds T - grad_t0 steps, entry (b, t - grad_t0) is
     dL/dh_t for t >= grad_t0, and dL/dh_t = 0 before.  0 = a gradient for every step (the reference's contract);
     T - 1 = only the LAST state is consumed (model.py:227-231: hidden2keyword(out[-1])), grad_h is then one [B,H]
     matrix and the dense, all-zero [T,B,H] gradient autograd would build for out[-1] is never written nor read. */
  int32_t grad_t0;
  int32_t reserved1;
} FgrnnBackward;

/* Bytes of device workspace the call needs (0 is possible). 256-byte aligned base required. */
FGRNN_API size_t fgrnn_forward_workspace_bytes(const FgrnnForward* d);
FGRNN_API size_t fgrnn_backward_workspace_bytes(const FgrnnBackward* d);

/* Which kernel family the descriptor selects (FGRNN_PATH_*), or -1 on an invalid descriptor. */
FGRNN_API int fgrnn_forward_plan(const FgrnnForward* d);
FGRNN_API int fgrnn_backward_plan(const FgrnnBackward* d);

/* Enqueue the work on `stream` (a cudaStream_t passed as void*; NULL = legacy default stream).
   Asynchronous: returns after the launches are queued. Returns FGRNN_OK or an error code. */
FGRNN_API int fgrnn_forward(const FgrnnForward* d, void* stream);
FGRNN_API int fgrnn_backward(const FgrnnBackward* d, void* stream);

/* Flat-bucket helper for data-parallel training: sums nothing, moves nothing across GPUs;
   it only reports how the parameter gradients are laid out when the caller points the d_*
   outputs into ONE contiguous fp32 bucket (so a single ncclAllReduce covers them).
   Writes element offsets in the fixed order {W|W1,W2, U|U1,U2, bias_gate, bias_update, zeta, nu}
   into offsets[8] (unused slots = -1) and returns the bucket length in floats. */
FGRNN_API int64_t fgrnn_grad_bucket_layout(const FgrnnProblem* p, int64_t offsets[8]);

FGRNN_API const char* fgrnn_strerror(int code);
/* Thread-local detail of the last non-OK return on this thread ("" if none). */
FGRNN_API const char* fgrnn_last_error_detail(void);
FGRNN_API int fgrnn_abi_version(void);
/* Cumulative number of kernels this library has launched in this process. */
FGRNN_API uint64_t fgrnn_launch_count(void);

/* Layout ingest for the reference trainer's native batches: x[b][f][t] by element strides (the loaders yield (B,F,T)
   with T innermost; trainClassifier.py:203-204 only permutes the VIEW) -> dst [B][T][F] contiguous fp32, one pass at
   HBM speed (32 x 32 tiles through shared memory).  The recurrence then reads dst with (stride_b, stride_t) = (T*F, F).
   mean / std (optional, [F] each, both or neither): the loaders' per-feature standardisation (x - mean) / std of
   preprocessing.py:60-76 applied on the way, with the reference's two correctly rounded operations. */
FGRNN_API int fgrnn_ingest_bft(const float* src, int64_t stride_b, int64_t stride_f, int64_t stride_t, float* dst,
                               const float* mean, const float* stdev, int32_t B, int32_t F, int32_t T, int32_t device, void* stream);

/* Classifier head of the keyword spotter on the LAST hidden state, forward + loss + backward in one launch
   (model.py:227-231 hidden2keyword + log_softmax, NLLLoss of trainClassifier.py:236; the torch version is ~15 launches):
     logits = h . W^T + b; loss = -mean_b log_softmax(logits)[b, labels[b]];
     dW [C,H], db [C], dh [B,H] (optional) = gradients of the MEAN loss; logp [B,C] optional.
   h / dh are given by row stride (e.g. the last time step of the hidden states / of a zeroed grad_h buffer).
   workspace: fgrnn_head_workspace_bytes(B,H,C) bytes, its first 4 bytes ZERO on entry (the kernel leaves them zero).
   C <= 16, H % 4 == 0.  Per-CTA partial sums are added in a fixed order: results are run-to-run identical. */
FGRNN_API size_t fgrnn_head_workspace_bytes(int32_t B, int32_t H, int32_t C);
FGRNN_API int fgrnn_head_nll(const float* h, int64_t h_stride, const float* W, const float* b, const int64_t* labels,
                             float* loss, float* dW, float* db, float* dh, int64_t dh_stride, float* logp,
                             void* workspace, size_t workspace_bytes, int32_t B, int32_t H, int32_t C, int32_t device, void* stream);

/* Plain SGD step over one flat parameter buffer and one flat gradient bucket (torch.optim.SGD without momentum /
   weight decay, trainClassifier.py:240): params[i] -= lr * (grad_scale * grads[i]); grad_scale = 1 / world size when the
   bucket holds the all-reduced SUM of the ranks' gradients. */
FGRNN_API int fgrnn_sgd_flat(float* params, const float* grads, int64_t n, float lr, float grad_scale, int32_t device, void* stream);

/* Data-parallel training step: gradient all-reduce FUSED with the SGD update over NVLink peer memory -- what replaces
   `all_reduce(flat bucket)` followed by `optimizer.step()` (trainClassifier.py:239-240 under one process per GPU; the north
   star's data-parallel configuration).  Every rank owns a region obtained from fgrnn_peer_alloc (zero-filled), sends its
   64-byte CUDA IPC handle to its peers (any transport: torch.distributed's object collectives, MPI, a file) and opens theirs
   with fgrnn_peer_open.  Region layout: [gradient bucket, n floats, padded to a multiple of 256 bytes][receive area,
   fgrnn_peer_recv_bytes(n, world) bytes].  Ranks must be in ONE node with P2P access (NVLink / NVSwitch), world <= 8.
     fgrnn_sgd_allreduce_peer: every rank pushes its bucket into the peers' receive areas (16-byte lines that carry the step
                               counter as their flag), then sum = bucket of rank 0 + rank 1 + ... (rank order, identical bits
                               on every rank); params -= lr * (grad_scale * sum); reduced (optional) = sum.
   Every rank must call it the same number of times with the same n (it is a collective); the launch is asynchronous on
   `stream` and may be captured in a CUDA graph (its step counters live in `state`: fgrnn_peer_state_bytes() of ZEROED device
   memory, private to the rank).  After a synchronisation, state[32] != 0 means a peer's lines did not arrive within 4 s. */
#define FGRNN_PEER_MAX_RANKS 8
#define FGRNN_PEER_HANDLE_BYTES 64
typedef struct FgrnnPeerStep {
  int32_t abi_version;       /* FGRNN_ABI_VERSION */
  int32_t device;
  int32_t world, rank;
  float* params;             /* [n] this rank's flat parameters (updated) */
  float* reduced;            /* [n] or NULL: the summed gradients (what p.grad holds after the all-reduce) */
  const float* bucket;       /* [n] this rank's gradients (the start of its own region) */
  void* recv[FGRNN_PEER_MAX_RANKS];        /* every rank's receive area as mapped in THIS process; [rank] = the local one */
  int32_t* state;            /* local, fgrnn_peer_state_bytes(), zero before the first call */
  int64_t n;
  float lr, grad_scale;      /* grad_scale = 1 / world for the mean over ranks */
} FgrnnPeerStep;
FGRNN_API size_t fgrnn_peer_recv_bytes(int64_t n, int32_t world);
FGRNN_API size_t fgrnn_peer_state_bytes(void);
FGRNN_API int fgrnn_peer_alloc(size_t bytes, int32_t device, void** ptr, unsigned char* handle /* [FGRNN_PEER_HANDLE_BYTES] out */);
FGRNN_API int fgrnn_peer_open(const unsigned char* handle, int32_t device, void** ptr);
FGRNN_API int fgrnn_peer_close(void* ptr, int32_t device);
FGRNN_API int fgrnn_peer_free(void* ptr, int32_t device);
FGRNN_API int fgrnn_sgd_allreduce_peer(const FgrnnPeerStep* step, void* stream);

/* Tuning / test override of the launchers' tile configuration.  Keys are the names of the environment variables that set
   the same values at process start (FGRNN_TC_NS, FGRNN_TC_NT, FGRNN_TC_BR_NS, FGRNN_TC_WIDE, FGRNN_FAST_NL,
   FGRNN_SMEM_CFG); the environment is read once, not on the launch path.  value NULL or "" clears the override. */
FGRNN_API int fgrnn_debug_set_tuning(const char* name, const char* value);

/* Diagnostic (tests only): fill every SM's tensor memory and shared memory with a NaN pattern, enqueued on
   `stream`, so that the next launch cannot be saved by operands a predecessor left on chip.  The reference
   has no counterpart; the first-launch parity tests call it between launches. */
FGRNN_API int fgrnn_debug_poison_onchip(int device, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FASTGRNN_B200_H_ */
