// Persistent shared-memory kernel family (FGRNN_PATH_SMEM): one launch runs all T steps.
//
// Data layout (per CTA, BM = 2*TM*WM batch rows, H = 128):
//   U_s [H][H]       recurrent weights, canonical (k-major rows, n contiguous), resident for all T steps
//   W_s [I][H]       input weights, resident
//   h_s [BM][H+16]   current hidden-state tile, row-major, padded
//   x_s [2][BM][I+4] double-buffered x_t tile (prefetched one step ahead through registers)
//
// Thread tiling (profiles/r01_lds_wavefront_rules.txt): on B200 an LDS.128 returns at most 256 B of
// register data per shared-memory wavefront and stays at its 2-wavefront minimum only when
// adjacent lane pairs share an address.  So within a warp  ty = lane & 1  selects one of two
// row groups and  tx = lane >> 1  one of 16 column chunks:
//   - A fragments (x_t / h_{t-1}): LDS.128 along k, two distinct padded rows per request  (2 wavefronts)
//   - B fragments (W / U rows):    LDS.128 along n, 16 chunks = 256 B, lane pairs share   (2 wavefronts)
//   - thread tile TM x TN (rows ty+2i, columns chunk tx of every 64-column block), warp tile
//     2*TM x 16*TN; shared wavefronts per FFMA = (TM+TN)/(2*TM*TN).
//   - h_t is produced by the thread that owns the same (row, col) in the next step, written to
//     h_s as STS.128 and to HBM as STG.128 (16 lanes = 256 contiguous bytes per row).
// Per step: pre = x_t.W + h_{t-1}.U (rnn.py:277-289) in fp32 FFMA, then the fused gate update
// (rnn.py:290-295).  Backward: the same structure on delta_{t-1} = z*G + dpre.U^T
// (cuda/fastgrnn_cuda_kernel.cu:110,537), with dpre materialised for the T-parallel contractions.
#include <cstdlib>
#include <cstring>

#include "fgrnn_kernels.cuh"

namespace fgrnn {

constexpr int SH = 128;            // hidden size this family is specialised for
constexpr int SHS = SH + 16;       // padded row stride of the state tile (floats)

template <int TM_, int TN_, int WM_, int WN_, int MINB_>
struct TileCfg {
  static constexpr int TM = TM_, TN = TN_, WM = WM_, WN = WN_, MINB = MINB_;
  static constexpr int NT = 32 * WM * WN;        // threads per CTA
  static constexpr int BM = 2 * TM * WM;         // batch rows per CTA
  static constexpr int NJ = TN / 4;              // 16-byte column chunks per thread
  static_assert(16 * TN * WN == SH, "warps must tile the 128 hidden columns");
};
using CfgA = TileCfg<8, 4, 4, 2, 1>;   // 64 rows, 256 threads, 1 CTA/SM : large batches
using CfgA7 = TileCfg<7, 4, 4, 2, 1>;  // 56 rows: 8192 rows -> 147 CTAs on 148 SMs
using CfgB = TileCfg<4, 8, 4, 1, 2>;   // 32 rows, 128 threads, 2 CTAs/SM
using CfgC = TileCfg<2, 4, 4, 2, 2>;   // 16 rows, 256 threads, 2 CTAs/SM : small per-GPU batches

template <int TM, int NJ>
__device__ __forceinline__ void mma_block(float (&acc)[TM][NJ * 4], const float* __restrict__ A_s, int lda,
                                          const float* __restrict__ B_s, int ldb, int K) {
  // acc[i][4j+q] += sum_k A_s[(2i)*lda + k] * B_s[k*ldb + 64j + q]   (A_s / B_s pre-offset per thread)
#pragma unroll 2
  for (int kb = 0; kb < K; kb += 4) {
    float4 av[TM];
    float4 bv[4][NJ];
#pragma unroll
    for (int i = 0; i < TM; ++i) av[i] = *reinterpret_cast<const float4*>(A_s + (2 * i) * lda + kb);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk)
#pragma unroll
      for (int j = 0; j < NJ; ++j) bv[kk][j] = *reinterpret_cast<const float4*>(B_s + (kb + kk) * ldb + 64 * j);
#pragma unroll
    for (int i = 0; i < TM; ++i) {
      const float a4[4] = {av[i].x, av[i].y, av[i].z, av[i].w};
#pragma unroll
      for (int kk = 0; kk < 4; ++kk)
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
          acc[i][4 * j + 0] = fmaf(a4[kk], bv[kk][j].x, acc[i][4 * j + 0]);
          acc[i][4 * j + 1] = fmaf(a4[kk], bv[kk][j].y, acc[i][4 * j + 1]);
          acc[i][4 * j + 2] = fmaf(a4[kk], bv[kk][j].z, acc[i][4 * j + 2]);
          acc[i][4 * j + 3] = fmaf(a4[kk], bv[kk][j].w, acc[i][4 * j + 3]);
        }
    }
  }
}

// load a [rows][cols] fp32 matrix into shared memory, optionally transposing (dst is [cols][rows] then)
template <int NT>
__device__ __forceinline__ void load_matrix(float* dst, const float* __restrict__ src, int rows, int cols,
                                            bool transpose, int tid) {
  const int total = rows * cols;
  if (!transpose) {
    for (int e = tid * 4; e < total; e += NT * 4)
      *reinterpret_cast<float4*>(dst + e) = __ldg(reinterpret_cast<const float4*>(src + e));
  } else {
    for (int e = tid; e < total; e += NT) {
      const int r = e / cols, c = e - r * cols;
      dst[c * rows + r] = __ldg(src + e);
    }
  }
}

template <typename Cfg, bool DEFNL>
__global__ void __launch_bounds__(Cfg::NT, Cfg::MINB) smem_fwd_kernel(const SmemFwdArgs a) {
  constexpr int TM = Cfg::TM, NJ = Cfg::NJ, TN = Cfg::TN, BM = Cfg::BM, NT = Cfg::NT, H = SH, HS = SHS;
  extern __shared__ __align__(16) float smem[];
  const Dims d = a.d;
  const int I = d.I, XS = I + 4;
  float* U_s = smem;                  // [H][H]
  float* W_s = U_s + H * H;           // [I][H]
  float* h_s = W_s + I * H;           // [BM][HS]
  float* x_s = h_s + BM * HS;         // [2][BM][XS]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int wm = warp / Cfg::WN, wn = warp % Cfg::WN;
  const int ty = lane & 1, tx = lane >> 1;
  const int cbase = wn * (16 * TN) + tx * 4;     // this thread's columns: cbase + 64*j + {0..3}
  const int rbase = wm * (2 * TM) + ty;          // this thread's rows:    rbase + 2*i
  const int row0 = blockIdx.x * BM;
  const float sz = sigmoid_f(__ldg(a.zeta)), sn = sigmoid_f(__ldg(a.nu));

  // ---- one-time: weights and initial state into shared memory -------------------------------
  const bool hi = a.layout == FGRNN_LAYOUT_HI;
  load_matrix<NT>(U_s, a.U, H, H, hi, tid);             // HI stores U^T (rnn.py:793): transpose on the way in
  if (hi) load_matrix<NT>(W_s, a.W, H, I, true, tid);   // W[H][I] -> W_s[I][H]
  else load_matrix<NT>(W_s, a.W, I, H, false, tid);
  for (int e = tid; e < BM * (H / 4); e += NT) {
    const int r = e / (H / 4), q = e - r * (H / 4);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (a.h0 != nullptr && row0 + r < d.B) v = __ldg(reinterpret_cast<const float4*>(a.h0 + (size_t)(row0 + r) * H) + q);
    *reinterpret_cast<float4*>(h_s + r * HS + q * 4) = v;
  }
  const int IQ = I >> 2;                          // 16-byte chunks per x row
  const int nchunk = BM * IQ;
  constexpr int XQ = (BM * 16 + NT - 1) / NT;     // chunks per thread per step at I = 64
  float4 xr[XQ];
  auto fetch_x = [&](int t) {
#pragma unroll
    for (int q = 0; q < XQ; ++q) {
      const int e = tid + q * NT;
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
      const int e = tid + q * NT;
      if (e < nchunk) {
        const int r = e / IQ, kq = e - r * IQ;
        *reinterpret_cast<float4*>(xb + r * XS + kq * 4) = xr[q];
      }
    }
  };
  fetch_x(0);
  stash_x(x_s);
  float bg[NJ * 4], bu[NJ * 4];
#pragma unroll
  for (int j = 0; j < NJ; ++j)
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      bg[4 * j + q] = __ldg(a.bias_gate + cbase + 64 * j + q);
      bu[4 * j + q] = __ldg(a.bias_update + cbase + 64 * j + q);
    }
  __syncthreads();

  for (int t = 0; t < d.T; ++t) {
    const float* xc = x_s + (t & 1) * BM * XS;
    float* xn = x_s + ((t & 1) ^ 1) * BM * XS;
    if (t + 1 < d.T) fetch_x(t + 1);              // latency hidden behind the FFMA loop below

    float acc[TM][NJ * 4];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
      for (int j = 0; j < NJ * 4; ++j) acc[i][j] = 0.f;
    // h.U first: its 128-term chain then starts from zero and keeps small partial sums, which
    // halves the fp32 accumulation error against the oracle (0.81 -> 0.43 of tolerance on the
    // config-1 case, profiles/r01_accumulation_order.txt); x.W's 32 terms are added last.
    mma_block<TM, NJ>(acc, h_s + rbase * HS, HS, U_s + cbase, H, H);      // h_{t-1}.U    (rnn.py:284)
    mma_block<TM, NJ>(acc, xc + rbase * XS, XS, W_s + cbase, H, I);       // + x_t.W      (rnn.py:278,289)
    __syncthreads();                                                      // every warp is done reading h_s

#pragma unroll
    for (int i = 0; i < TM; ++i) {
      const int r = rbase + 2 * i;
      const int row = row0 + r;
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        const int c = cbase + 64 * j;
        float* hp = h_s + r * HS + c;
        const float4 hold = *reinterpret_cast<const float4*>(hp);
        const float ho[4] = {hold.x, hold.y, hold.z, hold.w};
        float hn[4], zz[4], cc[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float pre = acc[i][4 * j + q];
          if (DEFNL) {
            zz[q] = (a.fast_nl & 1) ? sigmoid_fast(pre + bg[4 * j + q]) : sigmoid_f(pre + bg[4 * j + q]);   // rnn.py:290
            cc[q] = (a.fast_nl & 2) ? tanh_fast(pre + bu[4 * j + q]) : tanhf(pre + bu[4 * j + q]);          // rnn.py:292
          } else {
            zz[q] = act_rt(d.gate_nl, pre + bg[4 * j + q]);
            cc[q] = act_rt(d.update_nl, pre + bu[4 * j + q]);
          }
          hn[q] = zz[q] * ho[q] + (sz * (1.0f - zz[q]) + sn) * cc[q];                    // rnn.py:294-295
        }
        const float4 hv = make_float4(hn[0], hn[1], hn[2], hn[3]);
        *reinterpret_cast<float4*>(hp) = hv;
        if (row < d.B) {
          if (a.out) *reinterpret_cast<float4*>(a.out + (size_t)row * a.osb + (size_t)t * a.ost + c) = hv;
          if (a.save_z) {
            const size_t sidx = ((size_t)t * d.B + row) * H + c;
            *reinterpret_cast<float4*>(a.save_z + sidx) = make_float4(zz[0], zz[1], zz[2], zz[3]);
            *reinterpret_cast<float4*>(a.save_c + sidx) = make_float4(cc[0], cc[1], cc[2], cc[3]);
          }
          if (a.h_last && t == d.T - 1) *reinterpret_cast<float4*>(a.h_last + (size_t)row * H + c) = hv;
        }
      }
    }
    if (t + 1 < d.T) stash_x(xn);
    __syncthreads();                                                      // h_s / x_s ready for step t+1
  }
}

// ---------------------------------------------------------------------------------------------
// backward, serial part
// ---------------------------------------------------------------------------------------------
template <typename Cfg>
__global__ void __launch_bounds__(Cfg::NT, Cfg::MINB) smem_bwd_rec_kernel(const SmemBwdArgs a) {
  constexpr int TM = Cfg::TM, NJ = Cfg::NJ, TN = Cfg::TN, BM = Cfg::BM, NT = Cfg::NT, H = SH, HS = SHS;
  constexpr int NW = NT / 32;
  extern __shared__ __align__(16) float smem[];
  const Dims d = a.d;
  float* UT_s = smem;                  // [H(n)][H(k)] : UT_s[n][k] = U[k][n]
  float* dp_s = UT_s + H * H;          // [BM][HS]  dpre_t
  float* red_s = dp_s + BM * HS;       // [WM][2][H] per-row-warp column sums, then [2][NW] scalars
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int wm = warp / Cfg::WN, wn = warp % Cfg::WN;
  const int ty = lane & 1, tx = lane >> 1;
  const int cbase = wn * (16 * TN) + tx * 4;
  const int rbase = wm * (2 * TM) + ty;
  const int row0 = blockIdx.x * BM;
  const float sz = sigmoid_f(__ldg(a.zeta)), sn = sigmoid_f(__ldg(a.nu));
  // HI layout stores U^T already (rnn.py:793); IH needs the transpose
  load_matrix<NT>(UT_s, a.U, H, H, a.layout == FGRNN_LAYOUT_IH, tid);

  float acc[TM][NJ * 4];               // delta, carried in registers by the owning thread
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < NJ * 4; ++j) acc[i][j] = 0.f;
  float dbg[NJ * 4], dbu[NJ * 4];
#pragma unroll
  for (int j = 0; j < NJ * 4; ++j) { dbg[j] = 0.f; dbu[j] = 0.f; }
  float dzeta = 0.f, dnu = 0.f;
  __syncthreads();

  for (int t = d.T - 1; t >= 0; --t) {
#pragma unroll
    for (int i = 0; i < TM; ++i) {
      const int r = rbase + 2 * i;
      const int row = row0 + r;
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        const int c = cbase + 64 * j;
        float dp[4] = {0.f, 0.f, 0.f, 0.f};
        if (row < d.B) {
          const size_t sidx = ((size_t)t * d.B + row) * H + c;
          const float4 g4 = t >= a.gt0 ? __ldg(reinterpret_cast<const float4*>(a.grad_h + (size_t)row * a.gsb + (size_t)(t - a.gt0) * a.gst + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
          const float4 z4 = __ldg(reinterpret_cast<const float4*>(a.z_s + sidx));
          const float4 c4 = __ldg(reinterpret_cast<const float4*>(a.c_s + sidx));
          float4 h4 = make_float4(0.f, 0.f, 0.f, 0.f);
          if (t > 0) h4 = __ldg(reinterpret_cast<const float4*>(a.hs + (size_t)row * a.hsb + (size_t)(t - 1) * a.hst + c));
          else if (a.h0) h4 = __ldg(reinterpret_cast<const float4*>(a.h0 + (size_t)row * H + c));
          const float gg[4] = {g4.x, g4.y, g4.z, g4.w}, zz[4] = {z4.x, z4.y, z4.z, z4.w};
          const float cc[4] = {c4.x, c4.y, c4.z, c4.w}, hh[4] = {h4.x, h4.y, h4.z, h4.w};
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float G = gg[q] + acc[i][4 * j + q];                                      // cu:474
            const float dc = (sz * (1.0f - zz[q]) + sn) * dact_rt(d.update_nl, cc[q]) * G;  // cu:111
            const float dz = (hh[q] - sz * cc[q]) * dact_rt(d.gate_nl, zz[q]) * G;          // cu:112
            dp[q] = dc + dz;                                                                // cu:115
            dbu[4 * j + q] += dc; dbg[4 * j + q] += dz;                                     // cu:113-114
            dzeta = fmaf((1.0f - zz[q]) * cc[q], G, dzeta);                                 // cu:116
            dnu = fmaf(cc[q], G, dnu);                                                      // cu:117
            acc[i][4 * j + q] = zz[q] * G;                                                  // cu:110
          }
          *reinterpret_cast<float4*>(a.dpre_ws + sidx) = make_float4(dp[0], dp[1], dp[2], dp[3]);
        } else {
#pragma unroll
          for (int q = 0; q < 4; ++q) acc[i][4 * j + q] = 0.f;
        }
        *reinterpret_cast<float4*>(dp_s + r * HS + c) = make_float4(dp[0], dp[1], dp[2], dp[3]);
      }
    }
    __syncthreads();
    mma_block<TM, NJ>(acc, dp_s + rbase * HS, HS, UT_s + cbase, H, H);     // + dpre_t . U^T  (cu:537)
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int row = row0 + rbase + 2 * i;
    if (a.d_h0 && row < d.B) {
#pragma unroll
      for (int j = 0; j < NJ; ++j)
        *reinterpret_cast<float4*>(a.d_h0 + (size_t)row * H + cbase + 64 * j) =
            make_float4(acc[i][4 * j], acc[i][4 * j + 1], acc[i][4 * j + 2], acc[i][4 * j + 3]);
    }
  }
  // per-CTA partials, fixed summation order: the two ty lanes of a pair, then the WM row-warps
#pragma unroll
  for (int j = 0; j < NJ * 4; ++j) {
    dbg[j] += __shfl_xor_sync(0xffffffffu, dbg[j], 1);
    dbu[j] += __shfl_xor_sync(0xffffffffu, dbu[j], 1);
  }
  if (ty == 0) {
#pragma unroll
    for (int j = 0; j < NJ; ++j)
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        red_s[(wm * 2 + 0) * H + cbase + 64 * j + q] = dbg[4 * j + q];
        red_s[(wm * 2 + 1) * H + cbase + 64 * j + q] = dbu[4 * j + q];
      }
  }
  dzeta = warp_sum(dzeta);
  dnu = warp_sum(dnu);
  float* zn_s = red_s + Cfg::WM * 2 * H;     // [2][NW]
  if (lane == 0) { zn_s[warp] = dzeta; zn_s[NW + warp] = dnu; }
  __syncthreads();
  float* part = a.rec_partial + (size_t)blockIdx.x * (2 * H + 2);
  for (int e = tid; e < 2 * H; e += NT) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < Cfg::WM; ++w) s += red_s[w * 2 * H + e];
    part[e] = s;
  }
  if (tid == 0) {
    float s0 = 0.f, s1 = 0.f;
    for (int w = 0; w < NW; ++w) { s0 += zn_s[w]; s1 += zn_s[NW + w]; }
    part[2 * H] = s0;
    part[2 * H + 1] = s1;
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static size_t fwd_smem_bytes(int I, int BM) {
  return sizeof(float) * (size_t)(SH * SH + I * SH + BM * SHS + 2 * BM * (I + 4));
}
static size_t bwd_smem_bytes(int BM, int WM, int NW) {
  return sizeof(float) * (size_t)(SH * SH + BM * SHS + WM * 2 * SH + 2 * NW);
}

bool smem_path_supports(const Dims& d) {
  // full-rank, H == 128, I a multiple of 4 with W resident next to U
  return d.rW == 0 && d.rU == 0 && d.H == SH && d.I % 4 == 0 && d.I >= 4 && d.I <= 64;
}

static int cfg_rows(char c) { return c == 'A' ? CfgA::BM : c == '7' ? CfgA7::BM : c == 'B' ? CfgB::BM : CfgC::BM; }

// tile configuration: 'A' 64 rows, '7' 56 rows, 'B' 32 rows x 2 CTAs/SM, 'C' 16 rows x 2 CTAs/SM
static char pick_cfg(const Dims& d, bool backward) {
  if (tuning(TUNE_SMEM_CFG) != TUNE_UNSET) {                    // tuning / test override
    const char c = (char)tuning(TUNE_SMEM_CFG);
    const bool two_ok = fwd_smem_bytes(d.I, 32) * 2 + 2048 <= 227 * 1024 || backward;
    if (c == 'A' || c == '7') return c;
    if ((c == 'B' || c == 'C') && two_ok) return c;
  }
  const bool two_per_sm = backward || fwd_smem_bytes(d.I, 32) * 2 + 2048 <= 227 * 1024;
  if (two_per_sm && d.B <= 148 * 2 * 16) return 'C';           // small per-GPU batch: spread 16-row CTAs
  if (backward) return 'B';
  const int ctas56 = (d.B + 55) / 56, ctas64 = (d.B + 63) / 64;
  // one CTA per SM: prefer the tile height with fewer waves, then fewer idle rows
  const int waves56 = (ctas56 + 147) / 148, waves64 = (ctas64 + 147) / 148;
  return (waves56 * 56 <= waves64 * 64) ? '7' : 'A';
}

int smem_rows_per_cta(const Dims& d, int backward) { return cfg_rows(pick_cfg(d, backward != 0)); }

template <typename Cfg>
static int launch_fwd_t(const SmemFwdArgs& a, cudaStream_t stream) {
  const Dims& d = a.d;
  const bool defnl = d.gate_nl == FGRNN_NL_SIGMOID && d.update_nl == FGRNN_NL_TANH;
  const size_t smem = fwd_smem_bytes(d.I, Cfg::BM);
  const unsigned grid = (unsigned)((d.B + Cfg::BM - 1) / Cfg::BM);
  if (defnl) {
    auto k = smem_fwd_kernel<Cfg, true>;
    FGRNN_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<grid, Cfg::NT, smem, stream>>>(a);
  } else {
    auto k = smem_fwd_kernel<Cfg, false>;
    FGRNN_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<grid, Cfg::NT, smem, stream>>>(a);
  }
  FGRNN_LAUNCH_CHECK("smem_fwd_kernel");
  return FGRNN_OK;
}

int launch_smem_fwd(const SmemFwdArgs& a_, cudaStream_t stream) {
  SmemFwdArgs a = a_;
  a.fast_nl = 3;
  if (tuning(TUNE_FAST_NL) != TUNE_UNSET) a.fast_nl = tuning(TUNE_FAST_NL) & 3;   // bit0 sigmoid, bit1 tanh
  switch (pick_cfg(a.d, false)) {
    case 'A': return launch_fwd_t<CfgA>(a, stream);
    case '7': return launch_fwd_t<CfgA7>(a, stream);
    case 'B': return launch_fwd_t<CfgB>(a, stream);
    default: return launch_fwd_t<CfgC>(a, stream);
  }
}

template <typename Cfg>
static int launch_bwd_t(const SmemBwdArgs& a, cudaStream_t stream) {
  const size_t smem = bwd_smem_bytes(Cfg::BM, Cfg::WM, Cfg::NT / 32);
  const unsigned grid = (unsigned)((a.d.B + Cfg::BM - 1) / Cfg::BM);
  auto k = smem_bwd_rec_kernel<Cfg>;
  FGRNN_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k<<<grid, Cfg::NT, smem, stream>>>(a);
  FGRNN_LAUNCH_CHECK("smem_bwd_rec_kernel");
  return FGRNN_OK;
}

int smem_bwd_rec_ctas(const Dims& d) {
  const int rows = smem_rows_per_cta(d, 1);
  return (d.B + rows - 1) / rows;
}

int launch_smem_bwd_rec(const SmemBwdArgs& a, cudaStream_t stream) {
  switch (pick_cfg(a.d, true)) {
    case 'A': return launch_bwd_t<CfgA>(a, stream);
    case '7': return launch_bwd_t<CfgA7>(a, stream);
    case 'B': return launch_bwd_t<CfgB>(a, stream);
    default: return launch_bwd_t<CfgC>(a, stream);
  }
}

}  // namespace fgrnn
