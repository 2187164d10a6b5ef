// Persistent shared-memory kernel family (FGRNN_PATH_SMEM): one launch runs all T steps.
//
// Data layout (per CTA, BM = 8*TM batch rows):
//   U_s [H][H]   recurrent weights, canonical (k-major rows, n contiguous), resident for all T steps
//   W_s [I][H]   input weights, resident
//   h_s [BM][H+4] current hidden state tile (row-major, padded: conflict-free LDS.128 along k)
//   x_s [2][BM][I+4] double-buffered x_t tile (prefetched one step ahead through registers)
// Thread tiling: 8 warps = 2 (rows) x 4 (cols); warp tile (4*TM) x 32; lane = (ty 0..3, tx 0..7);
// a thread owns rows {ty + 4*i}, i < TM, of its warp's row block and 4 consecutive columns, so
//   - A fragments (x_t / h_{t-1}) are LDS.128 along k, 4 distinct padded rows per request (1 wavefront)
//   - B fragments (W / U rows) are LDS.128, 8 lanes x 16 B contiguous, broadcast over ty (1 wavefront)
//   - h_t is produced by the thread that owns the same (row, col) in the next step's epilogue,
//     written back to h_s as STS.128 and to HBM as STG.128 (8 lanes = one full 128 B line).
// Per step: pre = x_t.W + h_{t-1}.U (rnn.py:277-289) in fp32 FFMA, then the fused gate update
// (rnn.py:290-295).  Backward: the same structure on delta_{t-1} = z*G + dpre.U^T
// (cuda/fastgrnn_cuda_kernel.cu:110,537), with dpre materialised for the T-parallel contractions.
#include <cstdlib>

#include "fgrnn_kernels.cuh"

namespace fgrnn {

constexpr int SM_THREADS = 256;
constexpr int SM_WARPS_N = 4;       // H = 128 -> 4 warps across the 128 columns

template <int TM>
__device__ __forceinline__ void mma_block(float (&acc)[TM][4], const float* __restrict__ A_s, int lda,
                                          const float* __restrict__ B_s, int ldb, int K) {
  // acc[i][j] += sum_k A_s[(4*i)*lda + k] * B_s[k*ldb + j]  (A_s, B_s pre-offset to this thread's row / col)
#pragma unroll 2
  for (int kb = 0; kb < K; kb += 4) {
    float4 av[TM];
    float4 bv[4];
#pragma unroll
    for (int i = 0; i < TM; ++i) av[i] = *reinterpret_cast<const float4*>(A_s + (4 * i) * lda + kb);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) bv[kk] = *reinterpret_cast<const float4*>(B_s + (kb + kk) * ldb);
#pragma unroll
    for (int i = 0; i < TM; ++i) {
      const float a4[4] = {av[i].x, av[i].y, av[i].z, av[i].w};
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        acc[i][0] = fmaf(a4[kk], bv[kk].x, acc[i][0]);
        acc[i][1] = fmaf(a4[kk], bv[kk].y, acc[i][1]);
        acc[i][2] = fmaf(a4[kk], bv[kk].z, acc[i][2]);
        acc[i][3] = fmaf(a4[kk], bv[kk].w, acc[i][3]);
      }
    }
  }
}

// load a [rows][cols] fp32 matrix into shared memory, optionally transposing (dst is [cols][rows] then)
__device__ __forceinline__ void load_matrix(float* dst, const float* __restrict__ src, int rows, int cols,
                                            bool transpose, int tid) {
  const int total = rows * cols;
  if (!transpose) {
    for (int e = tid * 4; e < total; e += SM_THREADS * 4)
      *reinterpret_cast<float4*>(dst + e) = __ldg(reinterpret_cast<const float4*>(src + e));
  } else {
    for (int e = tid; e < total; e += SM_THREADS) {
      const int r = e / cols, c = e - r * cols;
      dst[c * rows + r] = __ldg(src + e);
    }
  }
}

template <int H, int TM, int MINB, bool DEFNL>
__global__ void __launch_bounds__(SM_THREADS, MINB) smem_fwd_kernel(const SmemFwdArgs a) {
  static_assert(H == 32 * SM_WARPS_N, "column tiling assumes H == 128");
  constexpr int BM = 8 * TM;
  constexpr int HS = H + 4;
  extern __shared__ __align__(16) float smem[];
  const Dims d = a.d;
  const int I = d.I, XS = I + 4;
  float* U_s = smem;                  // [H][H]
  float* W_s = U_s + H * H;           // [I][H]
  float* h_s = W_s + I * H;           // [BM][HS]
  float* x_s = h_s + BM * HS;         // [2][BM][XS]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int wm = warp / SM_WARPS_N, wn = warp % SM_WARPS_N;
  const int ty = lane >> 3, tx = lane & 7;
  const int c0 = wn * 32 + tx * 4;
  const int rbase = wm * (4 * TM) + ty;          // this thread's rows: rbase + 4*i
  const int row0 = blockIdx.x * BM;
  const float sz = sigmoid_f(__ldg(a.zeta)), sn = sigmoid_f(__ldg(a.nu));

  // ---- one-time: weights and initial state into shared memory -------------------------------
  const bool hi = a.layout == FGRNN_LAYOUT_HI;
  load_matrix(U_s, a.U, H, H, hi, tid);          // HI stores U^T (rnn.py:793): transpose on the way in
  if (hi) load_matrix(W_s, a.W, H, I, true, tid);    // W[H][I] -> W_s[I][H]
  else load_matrix(W_s, a.W, I, H, false, tid);
  for (int e = tid; e < BM * (H / 4); e += SM_THREADS) {
    const int r = e / (H / 4), q = e - r * (H / 4);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (a.h0 != nullptr && row0 + r < d.B) v = __ldg(reinterpret_cast<const float4*>(a.h0 + (size_t)(row0 + r) * H) + q);
    *reinterpret_cast<float4*>(h_s + r * HS + q * 4) = v;
  }
  const int IQ = I >> 2;                          // float4 chunks per x row
  const int nchunk = BM * IQ;
  constexpr int XQ = 4;                           // up to 4 chunks per thread per step (BM*I <= 4096)
  float4 xr[XQ];
  auto fetch_x = [&](int t) {
#pragma unroll
    for (int q = 0; q < XQ; ++q) {
      const int e = tid + q * SM_THREADS;
      xr[q] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (e < nchunk) {
        const int r = e / IQ, kq = e - r * IQ;
        if (row0 + r < d.B) {
          const int64_t off = (int64_t)(row0 + r) * a.xsb + (int64_t)t * a.xst + kq * 4;
          if (d.x_dtype == FGRNN_BF16) {
            const uint2 raw = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(a.x) + off));
            xr[q].x = __uint_as_float(raw.x << 16); xr[q].y = __uint_as_float(raw.x & 0xffff0000u);
            xr[q].z = __uint_as_float(raw.y << 16); xr[q].w = __uint_as_float(raw.y & 0xffff0000u);
          } else {
            xr[q] = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(a.x) + off));
          }
        }
      }
    }
  };
  auto stash_x = [&](float* xb) {
#pragma unroll
    for (int q = 0; q < XQ; ++q) {
      const int e = tid + q * SM_THREADS;
      if (e < nchunk) {
        const int r = e / IQ, kq = e - r * IQ;
        *reinterpret_cast<float4*>(xb + r * XS + kq * 4) = xr[q];
      }
    }
  };
  fetch_x(0);
  stash_x(x_s);
  float bg[4], bu[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) { bg[j] = __ldg(a.bias_gate + c0 + j); bu[j] = __ldg(a.bias_update + c0 + j); }
  __syncthreads();

  for (int t = 0; t < d.T; ++t) {
    const float* xc = x_s + (t & 1) * BM * XS;
    float* xn = x_s + ((t & 1) ^ 1) * BM * XS;
    if (t + 1 < d.T) fetch_x(t + 1);              // latency hidden behind the FFMA loop below

    float acc[TM][4];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    mma_block<TM>(acc, xc + rbase * XS, XS, W_s + c0, H, I);       // x_t . W      (rnn.py:278)
    mma_block<TM>(acc, h_s + rbase * HS, HS, U_s + c0, H, H);      // + h_{t-1}.U  (rnn.py:284,289)
    __syncthreads();                                               // every warp is done reading h_s

#pragma unroll
    for (int i = 0; i < TM; ++i) {
      const int r = rbase + 4 * i;
      float* hp = h_s + r * HS + c0;
      const float4 hold = *reinterpret_cast<const float4*>(hp);
      const float ho[4] = {hold.x, hold.y, hold.z, hold.w};
      float hn[4], zz[4], cc[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float pre = acc[i][j];
        zz[j] = DEFNL ? act<FGRNN_NL_SIGMOID>(pre + bg[j]) : act_rt(d.gate_nl, pre + bg[j]);   // rnn.py:290
        cc[j] = DEFNL ? act<FGRNN_NL_TANH>(pre + bu[j]) : act_rt(d.update_nl, pre + bu[j]);    // rnn.py:292
        hn[j] = zz[j] * ho[j] + (sz * (1.0f - zz[j]) + sn) * cc[j];                            // rnn.py:294-295
      }
      *reinterpret_cast<float4*>(hp) = make_float4(hn[0], hn[1], hn[2], hn[3]);
      const int row = row0 + r;
      if (row < d.B) {
        if (a.out) *reinterpret_cast<float4*>(a.out + (size_t)row * a.osb + (size_t)t * a.ost + c0) = make_float4(hn[0], hn[1], hn[2], hn[3]);
        if (a.save_z) {
          const size_t sidx = ((size_t)t * d.B + row) * H + c0;
          *reinterpret_cast<float4*>(a.save_z + sidx) = make_float4(zz[0], zz[1], zz[2], zz[3]);
          *reinterpret_cast<float4*>(a.save_c + sidx) = make_float4(cc[0], cc[1], cc[2], cc[3]);
        }
        if (a.h_last && t == d.T - 1) *reinterpret_cast<float4*>(a.h_last + (size_t)row * H + c0) = make_float4(hn[0], hn[1], hn[2], hn[3]);
      }
    }
    if (t + 1 < d.T) stash_x(xn);
    __syncthreads();                                               // h_s / x_s ready for step t+1
  }
}

// ---------------------------------------------------------------------------------------------
// backward, serial part
// ---------------------------------------------------------------------------------------------
template <int H, int TM, int MINB>
__global__ void __launch_bounds__(SM_THREADS, MINB) smem_bwd_rec_kernel(const SmemBwdArgs a) {
  constexpr int BM = 8 * TM;
  constexpr int HS = H + 4;
  extern __shared__ __align__(16) float smem[];
  const Dims d = a.d;
  float* UT_s = smem;                  // [H(n)][H(k)] : UT_s[n][k] = U[k][n]
  float* dp_s = UT_s + H * H;          // [BM][HS]  dpre_t
  float* red_s = dp_s + BM * HS;       // [2][H] + [2][8] cross-thread reduction scratch
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int wm = warp / SM_WARPS_N, wn = warp % SM_WARPS_N;
  const int ty = lane >> 3, tx = lane & 7;
  const int c0 = wn * 32 + tx * 4;
  const int rbase = wm * (4 * TM) + ty;
  const int row0 = blockIdx.x * BM;
  const float sz = sigmoid_f(__ldg(a.zeta)), sn = sigmoid_f(__ldg(a.nu));
  // HI layout stores U^T already (rnn.py:793); IH needs the transpose
  load_matrix(UT_s, a.U, H, H, a.layout == FGRNN_LAYOUT_IH, tid);
  for (int e = tid; e < 2 * H; e += SM_THREADS) red_s[e] = 0.f;

  float acc[TM][4];                    // delta, carried in registers by the owning thread
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  float dbg[4] = {0.f, 0.f, 0.f, 0.f}, dbu[4] = {0.f, 0.f, 0.f, 0.f};
  float dzeta = 0.f, dnu = 0.f;
  __syncthreads();

  for (int t = d.T - 1; t >= 0; --t) {
#pragma unroll
    for (int i = 0; i < TM; ++i) {
      const int r = rbase + 4 * i;
      const int row = row0 + r;
      float dp[4] = {0.f, 0.f, 0.f, 0.f};
      if (row < d.B) {
        const size_t sidx = ((size_t)t * d.B + row) * H + c0;
        const float4 g4 = __ldg(reinterpret_cast<const float4*>(a.grad_h + (size_t)row * a.gsb + (size_t)t * a.gst + c0));
        const float4 z4 = __ldg(reinterpret_cast<const float4*>(a.z_s + sidx));
        const float4 c4 = __ldg(reinterpret_cast<const float4*>(a.c_s + sidx));
        float4 h4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (t > 0) h4 = __ldg(reinterpret_cast<const float4*>(a.hs + (size_t)row * a.hsb + (size_t)(t - 1) * a.hst + c0));
        else if (a.h0) h4 = __ldg(reinterpret_cast<const float4*>(a.h0 + (size_t)row * H + c0));
        const float gg[4] = {g4.x, g4.y, g4.z, g4.w}, zz[4] = {z4.x, z4.y, z4.z, z4.w};
        const float cc[4] = {c4.x, c4.y, c4.z, c4.w}, hh[4] = {h4.x, h4.y, h4.z, h4.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float G = gg[j] + acc[i][j];                                             // cu:474
          const float dc = (sz * (1.0f - zz[j]) + sn) * dact_rt(d.update_nl, cc[j]) * G;  // cu:111
          const float dz = (hh[j] - sz * cc[j]) * dact_rt(d.gate_nl, zz[j]) * G;         // cu:112
          dp[j] = dc + dz;                                                               // cu:115
          dbu[j] += dc; dbg[j] += dz;                                                    // cu:113-114
          dzeta = fmaf((1.0f - zz[j]) * cc[j], G, dzeta);                                // cu:116
          dnu = fmaf(cc[j], G, dnu);                                                     // cu:117
          acc[i][j] = zz[j] * G;                                                         // cu:110
        }
        *reinterpret_cast<float4*>(a.dpre_ws + sidx) = make_float4(dp[0], dp[1], dp[2], dp[3]);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
      }
      *reinterpret_cast<float4*>(dp_s + r * HS + c0) = make_float4(dp[0], dp[1], dp[2], dp[3]);
    }
    __syncthreads();
    mma_block<TM>(acc, dp_s + rbase * HS, HS, UT_s + c0, H, H);      // + dpre_t . U^T  (cu:537)
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int row = row0 + rbase + 4 * i;
    if (a.d_h0 && row < d.B)
      *reinterpret_cast<float4*>(a.d_h0 + (size_t)row * H + c0) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
  }
  // per-CTA partials: columns are shared by the 4 ty lanes of a warp and by the 2 row-warps
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float g = dbg[j], u = dbu[j];
    g += __shfl_xor_sync(0xffffffffu, g, 8);  g += __shfl_xor_sync(0xffffffffu, g, 16);
    u += __shfl_xor_sync(0xffffffffu, u, 8);  u += __shfl_xor_sync(0xffffffffu, u, 16);
    dbg[j] = g; dbu[j] = u;
  }
  dzeta = warp_sum(dzeta);
  dnu = warp_sum(dnu);
  float* zn_s = red_s + 2 * H;     // [2][8]
  if (lane == 0) { zn_s[warp] = dzeta; zn_s[8 + warp] = dnu; }
  // two row-warps add into the same column slots one after the other (deterministic order)
  for (int w = 0; w < 2; ++w) {
    if (wm == w && ty == 0) {
#pragma unroll
      for (int j = 0; j < 4; ++j) { red_s[c0 + j] += dbg[j]; red_s[H + c0 + j] += dbu[j]; }
    }
    __syncthreads();
  }
  float* part = a.rec_partial + (size_t)blockIdx.x * (2 * H + 2);
  for (int e = tid; e < 2 * H; e += SM_THREADS) part[e] = red_s[e];
  if (tid == 0) {
    float s0 = 0.f, s1 = 0.f;
    for (int w = 0; w < 8; ++w) { s0 += zn_s[w]; s1 += zn_s[8 + w]; }
    part[2 * H] = s0;
    part[2 * H + 1] = s1;
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static size_t smem_fwd_bytes(int H, int I, int TM) {
  const int BM = 8 * TM;
  return sizeof(float) * (size_t)(H * H + I * H + BM * (H + 4) + 2 * BM * (I + 4));
}
static size_t smem_bwd_bytes(int H, int TM) {
  const int BM = 8 * TM;
  return sizeof(float) * (size_t)(H * H + BM * (H + 4) + 2 * H + 16);
}

bool smem_path_supports(const Dims& d) {
  // full-rank, H == 128, I a multiple of 4 with W resident next to U
  return d.rW == 0 && d.rU == 0 && d.H == 128 && d.I % 4 == 0 && d.I >= 4 && d.I <= 64;
}

int smem_rows_per_cta(const Dims& d, int num_sms) {
  // 32-row CTAs, two per SM: fills the machine at small per-GPU batches and lets one CTA's
  // epilogue / barrier stalls hide behind the other's FFMA loop. Needs W small enough for 2 CTAs/SM.
  (void)num_sms;
  if (const char* env = getenv("FGRNN_SMEM_ROWS")) {       // tuning / test override
    const int v = atoi(env);
    if (v == 64) return 64;
    if (v == 32 && smem_fwd_bytes(d.H, d.I, 4) * 2 + 2048 <= 227 * 1024) return 32;
  }
  if (smem_fwd_bytes(d.H, d.I, 4) * 2 + 2048 <= 227 * 1024) return 32;
  return 64;
}

template <int TM, int MINB>
static int launch_fwd_t(const SmemFwdArgs& a, cudaStream_t stream) {
  const Dims& d = a.d;
  const bool defnl = d.gate_nl == FGRNN_NL_SIGMOID && d.update_nl == FGRNN_NL_TANH;
  const size_t smem = smem_fwd_bytes(d.H, d.I, TM);
  const unsigned grid = (unsigned)((d.B + 8 * TM - 1) / (8 * TM));
  if (defnl) {
    auto k = smem_fwd_kernel<128, TM, MINB, true>;
    FGRNN_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<grid, SM_THREADS, smem, stream>>>(a);
  } else {
    auto k = smem_fwd_kernel<128, TM, MINB, false>;
    FGRNN_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<grid, SM_THREADS, smem, stream>>>(a);
  }
  FGRNN_LAUNCH_CHECK("smem_fwd_kernel");
  return FGRNN_OK;
}

int launch_smem_fwd(const SmemFwdArgs& a, int rows_per_cta, cudaStream_t stream) {
  if (rows_per_cta == 32) return launch_fwd_t<4, 2>(a, stream);
  return launch_fwd_t<8, 1>(a, stream);
}

int smem_bwd_rec_ctas(const Dims& d, int rows_per_cta) { return (d.B + rows_per_cta - 1) / rows_per_cta; }

int launch_smem_bwd_rec(const SmemBwdArgs& a, int rows_per_cta, cudaStream_t stream) {
  const Dims& d = a.d;
  const unsigned grid = (unsigned)smem_bwd_rec_ctas(d, rows_per_cta);
  if (rows_per_cta == 32) {
    auto k = smem_bwd_rec_kernel<128, 4, 2>;
    const size_t smem = smem_bwd_bytes(d.H, 4);
    FGRNN_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<grid, SM_THREADS, smem, stream>>>(a);
  } else {
    auto k = smem_bwd_rec_kernel<128, 8, 1>;
    const size_t smem = smem_bwd_bytes(d.H, 8);
    FGRNN_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<grid, SM_THREADS, smem, stream>>>(a);
  }
  FGRNN_LAUNCH_CHECK("smem_bwd_rec_kernel");
  return FGRNN_OK;
}

}  // namespace fgrnn
