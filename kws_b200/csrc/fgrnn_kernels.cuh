// Kernel-family argument blocks and host launchers (internal to libfastgrnn_b200.so).
#pragma once
#include "fgrnn_common.cuh"

namespace fgrnn {

// ---- forward -----------------------------------------------------------------------------
struct FwdArgs {
  Dims d;
  // canonical weights, row-major [K][N]:  Wc[I][H] | W1c[I][rW], W2c[rW][H];  Uc[H][H] | U1c[H][rU], U2c[rU][H]
  const float *Wc, *Uc, *W1c, *W2c, *U1c, *U2c;
  const float *bias_gate, *bias_update, *zeta, *nu;
  const float *gate_scale, *update_scale;     // optional per-unit factors on the pre-activations (null = 1)
  const void* x; int64_t xsb, xst;
  const float* h0;
  float* out; int64_t osb, ost;
  float* h_last; float* save_z; float* save_c;
};

int launch_gen_fwd(const FwdArgs& a, cudaStream_t stream);

// low-rank persistent FFMA forward (fgrnn_lr.cu): H = 256, W1/W2/U1/U2 resident in shared memory
bool lr_path_supports(const Dims& d);
int launch_lr_fwd(const FwdArgs& a, cudaStream_t stream);
// low-rank recurrence on the tensor cores (fgrnn_tc_lr.cu): H = 256, uRank <= 32, wRank <= 16, I <= 32; two chained tcgen05 stages
bool tc_lr_supports(const Dims& d);
int launch_tc_lr_fwd(const FwdArgs& a, cudaStream_t stream);

// ---- backward, serial part ---------------------------------------------------------------
struct BwdRecArgs {
  Dims d;
  const float *zeta, *nu;
  // transposed canonical recurrent weights: UT[n][k] = Uc[k][n];  U2T[H][rU];  U1T[rU][H]
  const float *UT, *U2T, *U1T;
  const float* grad_h; int64_t gsb, gst; int gt0;      // grad_h[b][t - gt0] for t >= gt0, zero upstream gradient before
  const float* hs; int64_t hsb, hst;
  const float* h0;
  const float *z_s, *c_s;
  float* dpre_ws;       // [T][B][H]
  float* rec_partial;   // [nCTA][2H+2]: d_bias_gate | d_bias_update | d_zeta(raw) | d_nu(raw)
  float* d_h0;
};

int gen_bwd_rec_ctas(const Dims& d);
int launch_gen_bwd_rec(const BwdRecArgs& a, cudaStream_t stream);

// ---- backward, T-parallel contractions ---------------------------------------------------
struct TnArgs {
  int M, B, T, K, N;
  const void* a; int64_t asb, ast; int a_dtype;
  int a_shift;            // 1: row (t,b) reads a[t-1,b] and h0[b] at t == 0  (the h_{t-1} operand)
  const float* a_h0;
  const float* dpre;      // [M][N], m = t*B + b
  float* partial;         // [nchunk][K][N]
  int rows_per_chunk;
};
int launch_gemm_tn_partial(const TnArgs& a, int nchunk, cudaStream_t stream);

struct NtArgs {
  int M, B, N, I;
  const float* dpre;      // [M][N]
  const float* Wf;        // canonical [I][N]
  float* dx; int64_t dsb, dst;
};
int launch_gemm_nt(const NtArgs& a, cudaStream_t stream);

struct PrepJob { const float* src; float* dst; int rows, cols, transpose; };
struct PrepJobs { int n; PrepJob job[8]; };
int launch_prep(const PrepJobs& jobs, cudaStream_t stream);

struct SmallGemm {
  const float* A; int lda, transA;
  const float* B; int ldb, transB;
  float* C; int transC;
  int M, N, K;
};
int launch_small_gemm(const SmallGemm& g, cudaStream_t stream);

struct ReduceArgs {
  int I, H, nchunk, nrec;
  const float *partW, *partU;
  float *dWc, *dUc;                 // destination (canonical or the caller's d_W/d_U)
  int dW_transpose, dU_transpose;   // write destination transposed (HI layout, full rank)
  const float* rec_partial;
  float *d_bias_gate, *d_bias_update, *d_zeta, *d_nu;
  const float *zeta, *nu;
};
int launch_reduce(const ReduceArgs& a, cudaStream_t stream);

// ---- persistent shared-memory family (fgrnn_smem.cu) ---------------------------------------
struct SmemFwdArgs {
  Dims d;
  int layout;                       // FGRNN_LAYOUT_*: weights are read in the caller's layout
  int fast_nl;                      // bit0: MUFU-based sigmoid, bit1: MUFU/polynomial tanh (default both)
  const float *W, *U;
  const float *bias_gate, *bias_update, *zeta, *nu;
  const void* x; int64_t xsb, xst;
  const float* h0;
  float* out; int64_t osb, ost;
  float* h_last; float* save_z; float* save_c;
};
struct SmemBwdArgs {
  Dims d;
  int layout;
  const float* U;
  const float *zeta, *nu;
  const float* grad_h; int64_t gsb, gst; int gt0;      // grad_h[b][t - gt0] for t >= gt0, zero upstream gradient before
  const float* hs; int64_t hsb, hst;
  const float* h0;
  const float *z_s, *c_s;
  float* dpre_ws;
  float* rec_partial;
  float* d_h0;
};
bool smem_path_supports(const Dims& d);
int smem_rows_per_cta(const Dims& d, int backward);
int launch_smem_fwd(const SmemFwdArgs& a, cudaStream_t stream);
int smem_bwd_rec_ctas(const Dims& d);
int launch_smem_bwd_rec(const SmemBwdArgs& a, cudaStream_t stream);

// ---- tcgen05 / TMEM family (fgrnn_tc.cu) ---------------------------------------------------
bool tc_path_supports(const Dims& d);
bool tc_x_tma_ok(const void* x, int64_t xsb, int64_t xst, int x_dtype, int B, int T);
int launch_tc_fwd(const SmemFwdArgs& a, cudaStream_t stream);
// wide shapes (fgrnn_tc_wx.cu): hoisted x.W GEMM + WX-stream recurrence; H = 128 or 256 (CTA pair), I <= 256
bool tc_wide_supports(const Dims& d);
size_t tc_wide_workspace_floats(const Dims& d);
int launch_tc_wide_fwd(const SmemFwdArgs& a, const float* gate_scale, const float* update_scale, float* wx_ws, cudaStream_t stream);

// T-parallel contractions dW, dU on the tensor cores (fgrnn_tc_bwd.cu); partials in the canonical orientation,
// one per CTA, summed by launch_reduce
struct TcContractLaunch {
  Dims d;
  const void* x; int64_t xsb, xst;
  const float* hs; int64_t hsb, hst;
  const float* h0;
  const float* dpre;
  float *partW, *partU;
};
bool tc_bwd_rec_supports(const Dims& d);
int tc_bwd_rec_ctas(const Dims& d);
int launch_tc_bwd_rec(const SmemBwdArgs& a, cudaStream_t stream);
bool tc_contract_supports(const Dims& d);
int tc_contract_ctas(const Dims& d);
int launch_tc_contract(const TcContractLaunch& c, cudaStream_t stream);
// reverse recurrence and contraction in one launch (the contraction CTAs follow the recurrence on the SMs it leaves free)
int tc_bwd_fused_contract_ctas(const Dims& d);           // 0: do not fuse
int launch_tc_bwd_fused(const SmemBwdArgs& a, const TcContractLaunch& c, int ncontract, int* progress, cudaStream_t stream);

}  // namespace fgrnn
