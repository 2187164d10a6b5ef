// Generic kernel family: correct for every supported shape (any I, H, ranks, nonlinearity,
// layout, stride).  Weights are streamed through L1/L2 (they are <= a few hundred KB and stay
// cache resident); the batch is tiled GEN_BM rows per CTA and each thread owns output columns.
// The shape-specialised persistent kernels (fgrnn_smem.cu, fgrnn_tc.cu) take over for the
// shapes they cover; this family is the always-available sm_100a path, not a CPU fallback.
//
// Math follows rnn.py:273-297 (forward) and cuda/fastgrnn_cuda_kernel.cu:109-118,537-556
// (backward, with the correct tanh-gate derivative).
#include "fgrnn_kernels.cuh"

namespace fgrnn {

constexpr int GEN_BM = 8;     // batch rows per CTA
constexpr int GEN_NT = 128;   // threads per CTA

// acc[r] += sum_k A_s[r*lda + k] * Wg[k*ldw + n]   (A in shared memory, W in global/L1)
template <int BM>
__device__ __forceinline__ void tile_mac(float (&acc)[BM], const float* __restrict__ A_s, int lda,
                                         const float* __restrict__ Wg, int ldw, int K, int n) {
  int k = 0;
  const float* __restrict__ wp = Wg + n;               // walks down column n: one 64-bit add per 4 k instead of four IMAD.WIDE
  const size_t step4 = (size_t)4 * ldw;
  const float* ap = A_s;
  for (; k + 4 <= K; k += 4, wp += step4, ap += 4) {
    const float w0 = __ldg(wp);
    const float w1 = __ldg(wp + ldw);
    const float w2 = __ldg(wp + 2 * ldw);
    const float w3 = __ldg(wp + 3 * ldw);
#pragma unroll
    for (int r = 0; r < BM; ++r) {
      const float4 a = *reinterpret_cast<const float4*>(ap + r * lda);
      acc[r] = fmaf(a.x, w0, acc[r]);
      acc[r] = fmaf(a.y, w1, acc[r]);
      acc[r] = fmaf(a.z, w2, acc[r]);
      acc[r] = fmaf(a.w, w3, acc[r]);
    }
  }
  for (; k < K; ++k) {
    const float w0 = __ldg(Wg + (size_t)k * ldw + n);
#pragma unroll
    for (int r = 0; r < BM; ++r) acc[r] = fmaf(A_s[r * lda + k], w0, acc[r]);
  }
}

static inline int pad4(int v) { return (v + 3) & ~3; }
__device__ __forceinline__ int dpad4(int v) { return (v + 3) & ~3; }

// ---------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(GEN_NT) gen_fwd_kernel(const FwdArgs a) {
  extern __shared__ __align__(16) float smem[];
  constexpr int BM = GEN_BM;
  const Dims d = a.d;
  const int Hp = dpad4(d.H), Ip = dpad4(d.I), rWp = dpad4(d.rW), rUp = dpad4(d.rU);
  float* h_s = smem;                      // [2][BM][Hp]
  float* x_s = h_s + 2 * BM * Hp;         // [2][BM][Ip]
  float* tw_s = x_s + 2 * BM * Ip;        // [BM][rWp]
  float* tu_s = tw_s + BM * rWp;          // [BM][rUp]
  const int smem_floats = 2 * BM * Hp + 2 * BM * Ip + BM * rWp + BM * rUp;
  const int tid = threadIdx.x;
  const int row0 = blockIdx.x * BM;
  const int nrows = min(BM, d.B - row0);
  const float sz = sigmoid_f(__ldg(a.zeta)), sn = sigmoid_f(__ldg(a.nu));

  for (int i = tid; i < smem_floats; i += GEN_NT) smem[i] = 0.0f;
  __syncthreads();
  if (a.h0 != nullptr) {
    for (int i = tid; i < nrows * d.H; i += GEN_NT) {
      const int r = i / d.H, n = i - r * d.H;
      h_s[r * Hp + n] = a.h0[(size_t)(row0 + r) * d.H + n];
    }
  }
  for (int i = tid; i < nrows * d.I; i += GEN_NT) {
    const int r = i / d.I, k = i - r * d.I;
    x_s[r * Ip + k] = load_x(a.x, (int64_t)(row0 + r) * a.xsb + k, d.x_dtype);
  }
  __syncthreads();

  for (int t = 0; t < d.T; ++t) {
    const int cur = t & 1;
    const float* hc = h_s + cur * BM * Hp;
    float* hn = h_s + (cur ^ 1) * BM * Hp;
    const float* xc = x_s + cur * BM * Ip;
    float* xn = x_s + (cur ^ 1) * BM * Ip;
    if (t + 1 < d.T) {
      for (int i = tid; i < nrows * d.I; i += GEN_NT) {
        const int r = i / d.I, k = i - r * d.I;
        xn[r * Ip + k] = load_x(a.x, (int64_t)(row0 + r) * a.xsb + (int64_t)(t + 1) * a.xst + k, d.x_dtype);
      }
    }
    if (d.rW > 0 || d.rU > 0) {
      // low-rank first stage, rnn.py:280 (x.W1) and rnn.py:286 (h.U1)
      for (int j = tid; j < d.rW + d.rU; j += GEN_NT) {
        float acc[BM];
#pragma unroll
        for (int r = 0; r < BM; ++r) acc[r] = 0.0f;
        if (j < d.rW) {
          tile_mac<BM>(acc, xc, Ip, a.W1c, d.rW, d.I, j);
#pragma unroll
          for (int r = 0; r < BM; ++r) tw_s[r * rWp + j] = acc[r];
        } else {
          const int ju = j - d.rW;
          tile_mac<BM>(acc, hc, Hp, a.U1c, d.rU, d.H, ju);
#pragma unroll
          for (int r = 0; r < BM; ++r) tu_s[r * rUp + ju] = acc[r];
        }
      }
      __syncthreads();
    }
    for (int n = tid; n < d.H; n += GEN_NT) {
      float accw[BM], accu[BM];
#pragma unroll
      for (int r = 0; r < BM; ++r) { accw[r] = 0.0f; accu[r] = 0.0f; }
      if (d.rW == 0) tile_mac<BM>(accw, xc, Ip, a.Wc, d.H, d.I, n);          // rnn.py:278
      else           tile_mac<BM>(accw, tw_s, rWp, a.W2c, d.H, d.rW, n);     // rnn.py:280-281
      if (d.rU == 0) tile_mac<BM>(accu, hc, Hp, a.Uc, d.H, d.H, n);          // rnn.py:284
      else           tile_mac<BM>(accu, tu_s, rUp, a.U2c, d.H, d.rU, n);     // rnn.py:286-287
      const float bg = __ldg(a.bias_gate + n), bu = __ldg(a.bias_update + n);
      const float sg = a.gate_scale ? __ldg(a.gate_scale + n) : 1.0f, su = a.update_scale ? __ldg(a.update_scale + n) : 1.0f;
#pragma unroll
      for (int r = 0; r < BM; ++r) {
        if (r < nrows) {
          const float pre = accw[r] + accu[r];                               // rnn.py:289
          const float z = act_rt(d.gate_nl, fmaf(sg, pre, bg));              // rnn.py:290 (sg, su: folded BatchNorm, rnn.py:402-408)
          const float c = act_rt(d.update_nl, fmaf(su, pre, bu));            // rnn.py:292
          const float hold = hc[r * Hp + n];
          const float hnew = z * hold + (sz * (1.0f - z) + sn) * c;          // rnn.py:294-295
          hn[r * Hp + n] = hnew;
          const size_t row = (size_t)(row0 + r);
          if (a.out) a.out[row * a.osb + (size_t)t * a.ost + n] = hnew;
          if (a.save_z) a.save_z[((size_t)t * d.B + row) * d.H + n] = z;
          if (a.save_c) a.save_c[((size_t)t * d.B + row) * d.H + n] = c;
          if (a.h_last && t == d.T - 1) a.h_last[row * d.H + n] = hnew;
        }
      }
    }
    __syncthreads();
  }
}

size_t gen_fwd_smem_bytes(const Dims& d) {
  return sizeof(float) * (size_t)(2 * GEN_BM * pad4(d.H) + 2 * GEN_BM * pad4(d.I) +
                                  GEN_BM * pad4(d.rW) + GEN_BM * pad4(d.rU));
}

int launch_gen_fwd(const FwdArgs& a, cudaStream_t stream) {
  const size_t smem = gen_fwd_smem_bytes(a.d);
  if (smem > 48 * 1024) {
    FGRNN_CUDA_TRY(cudaFuncSetAttribute(gen_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  const unsigned grid = (unsigned)((a.d.B + GEN_BM - 1) / GEN_BM);
  gen_fwd_kernel<<<grid, GEN_NT, smem, stream>>>(a);
  FGRNN_LAUNCH_CHECK("gen_fwd_kernel");
  return FGRNN_OK;
}

// ---------------------------------------------------------------------------------------------
// backward: the serial part.  delta_{t-1} = z_t*G_t + dpre_t.U^T (cu:110,537); dpre_t is
// materialised to `dpre_ws` so that the T-parallel sums (dW, dU, dX) run as dense contractions
// afterwards.  Per-CTA partials of db_gate, db_update, d_zeta, d_nu go to `rec_partial`.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(GEN_NT) gen_bwd_rec_kernel(const BwdRecArgs a) {
  extern __shared__ __align__(16) float smem[];
  constexpr int BM = GEN_BM;
  const Dims d = a.d;
  const int Hp = dpad4(d.H), rUp = dpad4(d.rU);
  float* dl_s = smem;                     // [2][BM][Hp]  delta (cur / next)
  float* dp_s = dl_s + 2 * BM * Hp;       // [BM][Hp]     dpre_t
  float* tu_s = dp_s + BM * Hp;           // [BM][rUp]    dpre.U2^T
  float* red_s = tu_s + BM * rUp;         // [2][GEN_NT/32]
  const int smem_floats = 3 * BM * Hp + BM * rUp;
  const int tid = threadIdx.x;
  const int row0 = blockIdx.x * BM;
  const int nrows = min(BM, d.B - row0);
  const float sz = sigmoid_f(__ldg(a.zeta)), sn = sigmoid_f(__ldg(a.nu));
  for (int i = tid; i < smem_floats; i += GEN_NT) smem[i] = 0.0f;
  __syncthreads();

  float* part = a.rec_partial + (size_t)blockIdx.x * (2 * d.H + 2);
  for (int n = tid; n < 2 * d.H; n += GEN_NT) part[n] = 0.0f;   // each thread zeroes what it later owns
  float dzeta_acc = 0.0f, dnu_acc = 0.0f;

  for (int t = d.T - 1; t >= 0; --t) {
    const int cur = (d.T - 1 - t) & 1;
    const float* dc_s = dl_s + cur * BM * Hp;
    float* dn_s = dl_s + (cur ^ 1) * BM * Hp;
    for (int n = tid; n < d.H; n += GEN_NT) {
      float sum_dc = 0.0f, sum_dz = 0.0f;
#pragma unroll
      for (int r = 0; r < BM; ++r) {
        if (r < nrows) {
          const size_t row = (size_t)(row0 + r);
          const size_t sidx = ((size_t)t * d.B + row) * d.H + n;
          const float G = (t >= a.gt0 ? a.grad_h[row * a.gsb + (size_t)(t - a.gt0) * a.gst + n] : 0.0f) + dc_s[r * Hp + n];   // cu:474
          const float z = a.z_s[sidx], c = a.c_s[sidx];
          float hp;
          if (t > 0) hp = a.hs[row * a.hsb + (size_t)(t - 1) * a.hst + n];
          else hp = a.h0 ? a.h0[row * d.H + n] : 0.0f;
          const float dc = (sz * (1.0f - z) + sn) * dact_rt(d.update_nl, c) * G;     // cu:111
          const float dz = (hp - sz * c) * dact_rt(d.gate_nl, z) * G;                // cu:112
          const float dpre = dc + dz;                                                // cu:115
          sum_dc += dc; sum_dz += dz;
          dzeta_acc = fmaf((1.0f - z) * c, G, dzeta_acc);                            // cu:116
          dnu_acc = fmaf(c, G, dnu_acc);                                             // cu:117
          dp_s[r * Hp + n] = dpre;
          dn_s[r * Hp + n] = z * G;                                                  // cu:110
          a.dpre_ws[sidx] = dpre;
        }
      }
      part[n] += sum_dz;          // d bias_gate   (cu:114)
      part[d.H + n] += sum_dc;    // d bias_update (cu:113)
    }
    __syncthreads();
    if (d.rU > 0) {
      for (int j = tid; j < d.rU; j += GEN_NT) {
        float acc[BM];
#pragma unroll
        for (int r = 0; r < BM; ++r) acc[r] = 0.0f;
        tile_mac<BM>(acc, dp_s, Hp, a.U2T, d.rU, d.H, j);
#pragma unroll
        for (int r = 0; r < BM; ++r) tu_s[r * rUp + j] = acc[r];
      }
      __syncthreads();
    }
    for (int k = tid; k < d.H; k += GEN_NT) {
      float acc[BM];
#pragma unroll
      for (int r = 0; r < BM; ++r) acc[r] = 0.0f;
      if (d.rU == 0) tile_mac<BM>(acc, dp_s, Hp, a.UT, d.H, d.H, k);                 // cu:537
      else           tile_mac<BM>(acc, tu_s, rUp, a.U1T, d.H, d.rU, k);
#pragma unroll
      for (int r = 0; r < BM; ++r) dn_s[r * Hp + k] += acc[r];
    }
    __syncthreads();
  }
  const float* dfin = dl_s + (d.T & 1) * BM * Hp;
  if (a.d_h0) {
    for (int i = tid; i < nrows * d.H; i += GEN_NT) {
      const int r = i / d.H, n = i - r * d.H;
      a.d_h0[(size_t)(row0 + r) * d.H + n] = dfin[r * Hp + n];
    }
  }
  dzeta_acc = warp_sum(dzeta_acc);
  dnu_acc = warp_sum(dnu_acc);
  if ((tid & 31) == 0) { red_s[tid >> 5] = dzeta_acc; red_s[GEN_NT / 32 + (tid >> 5)] = dnu_acc; }
  __syncthreads();
  if (tid == 0) {
    float s0 = 0.0f, s1 = 0.0f;
    for (int w = 0; w < GEN_NT / 32; ++w) { s0 += red_s[w]; s1 += red_s[GEN_NT / 32 + w]; }
    part[2 * d.H] = s0;
    part[2 * d.H + 1] = s1;
  }
}

int gen_bwd_rec_ctas(const Dims& d) { return (d.B + GEN_BM - 1) / GEN_BM; }

int launch_gen_bwd_rec(const BwdRecArgs& a, cudaStream_t stream) {
  const Dims& d = a.d;
  const size_t smem = sizeof(float) * (size_t)(3 * GEN_BM * pad4(d.H) + GEN_BM * pad4(d.rU) + 2 * (GEN_NT / 32));
  if (smem > 48 * 1024) {
    FGRNN_CUDA_TRY(cudaFuncSetAttribute(gen_bwd_rec_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  gen_bwd_rec_kernel<<<(unsigned)gen_bwd_rec_ctas(d), GEN_NT, smem, stream>>>(a);
  FGRNN_LAUNCH_CHECK("gen_bwd_rec_kernel");
  return FGRNN_OK;
}

// ---------------------------------------------------------------------------------------------
// T-parallel contractions (FFMA versions; the tcgen05 versions live in fgrnn_tc.cu)
//   TN:  P[chunk][K][N] = sum_{m in chunk} A[m][:]^T (x) D[m][:]      (dW = X^T.dPre, dU = Hprev^T.dPre)
//   NT:  C[m][i]        = sum_n D[m][n] * Wf[i][n]                    (dX = dPre.W^T)
// ---------------------------------------------------------------------------------------------
constexpr int TN_TILE = 64;   // K and N tile
constexpr int TN_MB = 16;     // rows staged per iteration

__device__ __forceinline__ float tn_load_a(const TnArgs& a, int m, int k) {
  const int t = m / a.B, b = m - t * a.B;
  if (!a.a_shift) return load_x(a.a, (int64_t)b * a.asb + (int64_t)t * a.ast + k, a.a_dtype);
  if (t > 0) return reinterpret_cast<const float*>(a.a)[(int64_t)b * a.asb + (int64_t)(t - 1) * a.ast + k];
  return a.a_h0 ? a.a_h0[(size_t)b * a.K + k] : 0.0f;
}

__global__ void __launch_bounds__(256) gemm_tn_partial_kernel(const TnArgs a) {
  __shared__ __align__(16) float As[TN_MB][TN_TILE];
  __shared__ __align__(16) float Ds[TN_MB][TN_TILE];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int n0 = blockIdx.x * TN_TILE, k0 = blockIdx.y * TN_TILE;
  const int chunk = blockIdx.z;
  const int m_begin = chunk * a.rows_per_chunk;
  const int m_end = min(a.M, m_begin + a.rows_per_chunk);
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;
  for (int m0 = m_begin; m0 < m_end; m0 += TN_MB) {
#pragma unroll
    for (int q = 0; q < (TN_MB * TN_TILE) / 256; ++q) {
      const int e = tid + q * 256;
      const int mi = e / TN_TILE, cc = e - mi * TN_TILE;
      const int m = m0 + mi;
      float av = 0.0f, dv = 0.0f;
      if (m < m_end) {
        if (k0 + cc < a.K) av = tn_load_a(a, m, k0 + cc);
        if (n0 + cc < a.N) dv = a.dpre[(size_t)m * a.N + n0 + cc];
      }
      As[mi][cc] = av;
      Ds[mi][cc] = dv;
    }
    __syncthreads();
#pragma unroll
    for (int mi = 0; mi < TN_MB; ++mi) {
      const float4 av = *reinterpret_cast<const float4*>(&As[mi][ty * 4]);
      const float4 dv = *reinterpret_cast<const float4*>(&Ds[mi][tx * 4]);
      const float aa[4] = {av.x, av.y, av.z, av.w};
      const float dd[4] = {dv.x, dv.y, dv.z, dv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(aa[i], dd[j], acc[i][j]);
    }
    __syncthreads();
  }
  float* P = a.partial + (size_t)chunk * a.K * a.N;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int k = k0 + ty * 4 + i;
    if (k >= a.K) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n < a.N) P[(size_t)k * a.N + n] = acc[i][j];
    }
  }
}

int launch_gemm_tn_partial(const TnArgs& a, int nchunk, cudaStream_t stream) {
  dim3 grid((a.N + TN_TILE - 1) / TN_TILE, (a.K + TN_TILE - 1) / TN_TILE, nchunk);
  gemm_tn_partial_kernel<<<grid, 256, 0, stream>>>(a);
  FGRNN_LAUNCH_CHECK("gemm_tn_partial_kernel");
  return FGRNN_OK;
}

constexpr int NT_TM = 64, NT_NB = 16, NT_LD = 68;

// dx[m][i] = sum_n dpre[m][n] * Wf[i][n]  (cu:552 d_input).  Tile = 64 rows x TI input features (TI = 32 when I <= 32:
// the flagship I = 32 would waste half of a 64-wide tile), 16 n per pass; 16-byte global loads along n when N % 4 == 0.
template <int TI>
__global__ void __launch_bounds__(256) gemm_nt_kernel(const NtArgs a) {
  constexpr int CJ = TI / 16;                         // output columns per thread
  __shared__ __align__(16) float Ds[NT_NB][NT_LD];   // [n][m]
  __shared__ __align__(16) float Ws[NT_NB][NT_LD];   // [n][i]
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.x * NT_TM, i0 = blockIdx.y * TI;
  const bool vec = (a.N & 3) == 0 && ((reinterpret_cast<uintptr_t>(a.dpre) | reinterpret_cast<uintptr_t>(a.Wf)) & 15) == 0;
  float acc[4][CJ];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < CJ; ++j) acc[i][j] = 0.0f;
  for (int nb = 0; nb < a.N; nb += NT_NB) {
    if (vec) {                                        // one float4 of dpre (and of W for the first TI rows) per thread
      const int r = tid >> 2, n4 = (tid & 3) * 4;
      float4 dv = make_float4(0.f, 0.f, 0.f, 0.f), wv = dv;
      if (nb + n4 < a.N) {
        if (m0 + r < a.M) dv = *reinterpret_cast<const float4*>(a.dpre + (size_t)(m0 + r) * a.N + nb + n4);
        if (r < TI && i0 + r < a.I) wv = __ldg(reinterpret_cast<const float4*>(a.Wf + (size_t)(i0 + r) * a.N + nb + n4));
      }
      Ds[n4 + 0][r] = dv.x; Ds[n4 + 1][r] = dv.y; Ds[n4 + 2][r] = dv.z; Ds[n4 + 3][r] = dv.w;
      if (r < TI) { Ws[n4 + 0][r] = wv.x; Ws[n4 + 1][r] = wv.y; Ws[n4 + 2][r] = wv.z; Ws[n4 + 3][r] = wv.w; }
    } else {
#pragma unroll
      for (int q = 0; q < (NT_TM * NT_NB) / 256; ++q) {
        const int e = tid + q * 256;
        const int r = e / NT_NB, nn = e - r * NT_NB;
        float dv = 0.0f, wv = 0.0f;
        if (nb + nn < a.N) {
          if (m0 + r < a.M) dv = a.dpre[(size_t)(m0 + r) * a.N + nb + nn];
          if (r < TI && i0 + r < a.I) wv = a.Wf[(size_t)(i0 + r) * a.N + nb + nn];
        }
        Ds[nn][r] = dv;
        if (r < TI) Ws[nn][r] = wv;
      }
    }
    __syncthreads();
#pragma unroll
    for (int nn = 0; nn < NT_NB; ++nn) {
      const float4 dv = *reinterpret_cast<const float4*>(&Ds[nn][ty * 4]);
      const float dd[4] = {dv.x, dv.y, dv.z, dv.w};
      float ww[CJ];
#pragma unroll
      for (int j = 0; j < CJ; ++j) ww[j] = Ws[nn][tx * CJ + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < CJ; ++j) acc[i][j] = fmaf(dd[i], ww[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= a.M) continue;
    const int t = m / a.B, b = m - t * a.B;
#pragma unroll
    for (int j = 0; j < CJ; ++j) {
      const int ii = i0 + tx * CJ + j;
      if (ii < a.I) a.dx[(size_t)b * a.dsb + (size_t)t * a.dst + ii] = acc[i][j];
    }
  }
}

int launch_gemm_nt(const NtArgs& a, cudaStream_t stream) {
  if (a.I <= 32) {
    dim3 grid((a.M + NT_TM - 1) / NT_TM, 1);
    gemm_nt_kernel<32><<<grid, 256, 0, stream>>>(a);
  } else {
    dim3 grid((a.M + NT_TM - 1) / NT_TM, (a.I + 63) / 64);
    gemm_nt_kernel<64><<<grid, 256, 0, stream>>>(a);
  }
  FGRNN_LAUNCH_CHECK("gemm_nt_kernel");
  return FGRNN_OK;
}

// ---------------------------------------------------------------------------------------------
// small dense helpers on parameter-sized matrices
// ---------------------------------------------------------------------------------------------
__global__ void prep_kernel(const PrepJobs jobs) {
  const PrepJob j = jobs.job[blockIdx.y];
  const int total = j.rows * j.cols;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    const int r = e / j.cols, c = e - r * j.cols;
    if (j.transpose) j.dst[(size_t)c * j.rows + r] = j.src[e];
    else j.dst[e] = j.src[e];
  }
}

int launch_prep(const PrepJobs& jobs, cudaStream_t stream) {
  if (jobs.n == 0) return FGRNN_OK;
  dim3 grid(16, jobs.n);
  prep_kernel<<<grid, 256, 0, stream>>>(jobs);
  FGRNN_LAUNCH_CHECK("prep_kernel");
  return FGRNN_OK;
}

// C[M][N] (+)= opA(A)[M][K] . opB(B)[K][N]; one thread per output; optionally writes C transposed.
__global__ void small_gemm_kernel(const SmallGemm g) {
  const int total = g.M * g.N;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    const int m = e / g.N, n = e - m * g.N;
    float acc = 0.0f;
    for (int k = 0; k < g.K; ++k) {
      const float av = g.transA ? g.A[(size_t)k * g.lda + m] : g.A[(size_t)m * g.lda + k];
      const float bv = g.transB ? g.B[(size_t)n * g.ldb + k] : g.B[(size_t)k * g.ldb + n];
      acc = fmaf(av, bv, acc);
    }
    if (g.transC) g.C[(size_t)n * g.M + m] = acc;
    else g.C[(size_t)m * g.N + n] = acc;
  }
}

int launch_small_gemm(const SmallGemm& g, cudaStream_t stream) {
  const int total = g.M * g.N;
  if (total == 0) return FGRNN_OK;
  small_gemm_kernel<<<(total + 255) / 256, 256, 0, stream>>>(g);
  FGRNN_LAUNCH_CHECK("small_gemm_kernel");
  return FGRNN_OK;
}

// Deterministic reduction of the per-CTA / per-chunk partials (fixed summation order, no atomics):
//   job 0: dW canonical [I][H]  <- sum over nchunk TN partials      (written straight to d_W when it
//   job 1: dU canonical [H][H]  <- sum over nchunk TN partials       is full rank, honouring layout)
//   job 2: d_bias_gate, d_bias_update, d_zeta, d_nu <- sum over recurrence CTAs
__global__ void __launch_bounds__(256) reduce_kernel(const ReduceArgs a) {
  const int job = blockIdx.y;
  if (job < 2) {
    // 32 consecutive elements per block, the partials split over 8 thread groups; both sums run in a fixed
    // order, so the result is identical from run to run
    __shared__ float part_s[8][33];
    const float* P = job == 0 ? a.partW : a.partU;
    float* dst = job == 0 ? a.dWc : a.dUc;
    const int K = job == 0 ? a.I : a.H;
    const int transpose = job == 0 ? a.dW_transpose : a.dU_transpose;
    if (dst == nullptr) return;
    const int total = K * a.H;
    const int el = threadIdx.x & 31, grp = threadIdx.x >> 5;
    for (int e0 = blockIdx.x * 32; e0 < total; e0 += gridDim.x * 32) {
      const int e = e0 + el;
      float s = 0.0f;
      if (e < total)
        for (int c = grp; c < a.nchunk; c += 8) s += P[(size_t)c * total + e];
      part_s[grp][el] = s;
      __syncthreads();
      if (grp == 0 && e < total) {
        float r = part_s[0][el];
#pragma unroll
        for (int g = 1; g < 8; ++g) r += part_s[g][el];
        if (transpose) { const int k = e / a.H, n = e - k * a.H; dst[(size_t)n * K + k] = r; }
        else dst[e] = r;
      }
      __syncthreads();
    }
  } else {
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gstride = gridDim.x * blockDim.x;
    const int stride = 2 * a.H + 2;
    for (int e = gtid; e < stride; e += gstride) {
      float s = 0.0f;
      for (int c = 0; c < a.nrec; ++c) s += a.rec_partial[(size_t)c * stride + e];
      if (e < a.H) { if (a.d_bias_gate) a.d_bias_gate[e] = s; }
      else if (e < 2 * a.H) { if (a.d_bias_update) a.d_bias_update[e - a.H] = s; }
      else if (e == 2 * a.H) {
        const float sz = sigmoid_f(__ldg(a.zeta));
        if (a.d_zeta) a.d_zeta[0] = s * sz * (1.0f - sz);                 // cu:116,544
      } else {
        const float sn = sigmoid_f(__ldg(a.nu));
        if (a.d_nu) a.d_nu[0] = s * sn * (1.0f - sn);                     // cu:117,545
      }
    }
  }
}

int launch_reduce(const ReduceArgs& a, cudaStream_t stream) {
  dim3 grid(256, 3);
  reduce_kernel<<<grid, 256, 0, stream>>>(a);
  FGRNN_LAUNCH_CHECK("reduce_kernel");
  return FGRNN_OK;
}

}  // namespace fgrnn
