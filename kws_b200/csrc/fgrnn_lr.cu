// Low-rank forward recurrence (rnn.py:280-287, factored order (x.W1).W2 + (h.U1).U2), hidden size 256:
// persistent FFMA kernel with W1, W2, U1, U2 resident in shared memory for all T steps.
//
//   CTA = 64 batch rows, 512 threads.  Per step:
//   stage 1  s[row][j] = h[row][:] . U1[:][j]  (j < rU)   and   sx[row][j] = x_t[row][:] . W1[:][j]  (j < rW)
//            h.U1: thread tile 4 rows x 4 ranks, K in blocks of 4 (8 LDS.128 per 64 FFMA); K = 256 is cut into KS slices
//            so that all 16 warps take part (one warp per scheduler cannot hide its own LDS latency: measured IPC
//            0.45), the KS partial sums meet in shared memory and are added in a short second pass.
//            x.W1 (K = I): 1 row x 2 ranks per thread, written directly.
//   stage 2  pre[row][n] = [s | sx][row][:] . [U2 ; W2][:][n]
//            thread tile 8 rows x 4 units (s broadcast within the warp, weights 16 B per lane): 12 LDS.128 per 128 FFMA,
//            followed by the gate update on the thread's 32 elements, h back to shared memory, 16-byte global stores.
// Algorithmic FLOPs per row-step 2*(256*32 + 32*16 + 48*256) = 41 984; the kernel is FFMA-issue bound.
// The tensor-core version of this path (two chained tcgen05 stages) is the round-2 item in DESIGN.md section 6.
#include "fgrnn_kernels.cuh"

namespace fgrnn {

constexpr int LR_H = 256, LR_BM = 64, LR_THREADS = 512;
constexpr int LR_HP = LR_H + 4;                 // padded h row: the 4 rows of a stage-1 tile fall into distinct banks
constexpr int LR_MAX_X4 = 2;                    // x tile: at most 64 x 64 floats = 1024 float4, two per thread

struct LrSmem { int U1, W1, V2, h, s, x, part, total, SP, KS; };      // offsets in floats
__host__ __device__ inline LrSmem lr_smem_layout(int I, int rW, int rU) {
  LrSmem L;
  L.SP = rU + rW + 4;
  L.U1 = 0;
  L.W1 = L.U1 + LR_H * rU;
  L.V2 = L.W1 + I * rW;                         // [rU + rW][H]: U2 stacked on W2
  L.h = L.V2 + (rU + rW) * LR_H;
  L.s = L.h + LR_BM * LR_HP;
  L.x = L.s + LR_BM * L.SP;                     // [2][BM][I]
  // K slices of the h.U1 stage: the largest power of two with KS * (16 * rU/4) threads <= 512
  L.KS = 1;
  while (L.KS * 2 * (4 * rU) <= LR_THREADS && LR_H % (8 * L.KS) == 0) L.KS *= 2;
  L.part = L.x + 2 * LR_BM * I;                 // [KS][BM][rU] partial sums
  L.total = L.part + L.KS * LR_BM * rU;
  return L;
}

__device__ __forceinline__ void lr_copy4(float* dst, const float* __restrict__ src, int n, int tid) {
  for (int i = tid; i < n / 4; i += LR_THREADS) reinterpret_cast<float4*>(dst)[i] = __ldg(reinterpret_cast<const float4*>(src) + i);
}

__device__ __forceinline__ void lr_fma4(float (&acc)[4], float s, const float4& w) {
  acc[0] = fmaf(s, w.x, acc[0]); acc[1] = fmaf(s, w.y, acc[1]); acc[2] = fmaf(s, w.z, acc[2]); acc[3] = fmaf(s, w.w, acc[3]);
}

// 4 rows x 4 outputs over K (multiple of 4): A rows in shared memory (stride lda), weights [K][ldw] in shared memory
template <int UNROLL>
__device__ __forceinline__ void lr_tile_4x4(float (&acc)[4][4], const float* A, int lda, const float* Wt, int ldw, int K) {
#pragma unroll UNROLL
  for (int k = 0; k < K; k += 4) {
    float4 av[4], wv[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) av[r] = *reinterpret_cast<const float4*>(A + r * lda + k);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) wv[kk] = *reinterpret_cast<const float4*>(Wt + (k + kk) * ldw);
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      lr_fma4(acc[r], av[r].x, wv[0]); lr_fma4(acc[r], av[r].y, wv[1]);
      lr_fma4(acc[r], av[r].z, wv[2]); lr_fma4(acc[r], av[r].w, wv[3]);
    }
  }
}

// RU_, RW_, I_ != 0: compile-time shape (the C4 shape 32/16/32 folds every shared-memory stride into immediates)
template <bool FAST_NL, int RU_, int RW_, int I_>
__global__ void __launch_bounds__(LR_THREADS, 1) lr_fwd_kernel(const FwdArgs a, const int bm) {
  extern __shared__ __align__(16) float sm[];
  const Dims d = a.d;
  const int I = I_ ? I_ : d.I, rW = RW_ ? RW_ : d.rW, rU = RU_ ? RU_ : d.rU, R = rU + rW;
  const LrSmem L = lr_smem_layout(I, rW, rU);
  float *U1s = sm + L.U1, *W1s = sm + L.W1, *V2s = sm + L.V2, *h_s = sm + L.h, *s_s = sm + L.s, *x_s = sm + L.x;
  const int SP = L.SP;
  const int tid = threadIdx.x;
  const int row0 = blockIdx.x * bm;              // bm <= 64 rows of this CTA (multiple of 8): the rest of the tile idles
  float* part_s = sm + L.part;

  // ---- prologue: weights, h0, x_0 -----------------------------------------------------------
  lr_copy4(U1s, a.U1c, LR_H * rU, tid);
  lr_copy4(W1s, a.W1c, I * rW, tid);
  lr_copy4(V2s, a.U2c, rU * LR_H, tid);
  lr_copy4(V2s + rU * LR_H, a.W2c, rW * LR_H, tid);
  for (int i = tid; i < LR_BM * (LR_H / 4); i += LR_THREADS) {
    const int r = i / (LR_H / 4), c4 = i - r * (LR_H / 4);
    const int row = row0 + r;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (a.h0 != nullptr && r < bm && row < d.B) v = __ldg(reinterpret_cast<const float4*>(a.h0 + (size_t)row * LR_H) + c4);
    *reinterpret_cast<float4*>(h_s + r * LR_HP + c4 * 4) = v;
  }
  // x tile tasks of this thread: (row, 4-feature chunk), step invariant
  const int xq = I / 4, xtasks = LR_BM * xq;
  int x_dst[LR_MAX_X4];
  int64_t x_src[LR_MAX_X4];
  bool x_live[LR_MAX_X4];
#pragma unroll
  for (int q = 0; q < LR_MAX_X4; ++q) {
    const int e = q * LR_THREADS + tid;
    const int r = e / xq, c4 = e - r * xq;
    x_live[q] = e < xtasks && r < bm && row0 + r < d.B;
    x_dst[q] = e < xtasks ? r * I + c4 * 4 : -1;
    x_src[q] = (int64_t)(row0 + r) * a.xsb + c4 * 4;
  }
  auto load_x4 = [&](int q, int t) -> float4 {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (x_live[q]) {
      const int64_t idx = x_src[q] + (int64_t)t * a.xst;
      if (d.x_dtype == FGRNN_BF16) {
        const uint2 p = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(a.x) + idx));
        v.x = __uint_as_float(p.x << 16); v.y = __uint_as_float(p.x & 0xffff0000u);
        v.z = __uint_as_float(p.y << 16); v.w = __uint_as_float(p.y & 0xffff0000u);
      } else {
        v = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(a.x) + idx));
      }
    }
    return v;
  };
#pragma unroll
  for (int q = 0; q < LR_MAX_X4; ++q)
    if (x_dst[q] >= 0) *reinterpret_cast<float4*>(x_s + x_dst[q]) = load_x4(q, 0);

  // stage-1 roles: h.U1 tile (4 rows x 4 ranks) of K slice `ks`; x.W1 items (1 row x 2 ranks); reduction items
  const int ujg = rU / 4, nU = 16 * ujg, KS = L.KS, kslice = LR_H / KS;
  const int ks = tid / nU, t1 = tid - ks * nU;
  const int s1_jg = t1 % ujg, s1_rg = t1 / ujg;
  const bool is_u = ks < KS && s1_rg * 4 < bm;
  const int witems = LR_BM * (rW / 2), wjp = rW / 2;
  const int ritems = LR_BM * ujg;
  // stage-2 / epilogue roles
  const int ng = tid & 63, rg = tid >> 6, n0 = ng * 4;
  const float4 bg = __ldg(reinterpret_cast<const float4*>(a.bias_gate) + ng);
  const float4 bu = __ldg(reinterpret_cast<const float4*>(a.bias_update) + ng);
  const float sz = sigmoid_f(__ldg(a.zeta)), sn = sigmoid_f(__ldg(a.nu));
  const int first_row = row0 + rg * 8;
  float* outp = a.out ? a.out + (size_t)first_row * a.osb + n0 : nullptr;
  float* zp = a.save_z ? a.save_z + (size_t)first_row * LR_H + n0 : nullptr;
  float* cp = a.save_c ? a.save_c + (size_t)first_row * LR_H + n0 : nullptr;
  const int rows_left = min(bm - rg * 8, d.B - first_row);
  const bool s2_active = rg * 8 < bm;
  __syncthreads();

  for (int t = 0; t < d.T; ++t) {
    const float* xc = x_s + (t & 1) * LR_BM * I;
    float* xn = x_s + ((t & 1) ^ 1) * LR_BM * I;
    float4 xnext[LR_MAX_X4];
    const bool more = t + 1 < d.T;
    if (more) {
#pragma unroll
      for (int q = 0; q < LR_MAX_X4; ++q) xnext[q] = load_x4(q, t + 1);        // lands during the two stages
    }
    // ---- stage 1 ----
    if (is_u) {
      float acc[4][4] = {};
      lr_tile_4x4<2>(acc, h_s + (s1_rg * 4) * LR_HP + ks * kslice, LR_HP, U1s + (ks * kslice) * rU + s1_jg * 4, rU, kslice);
      float* pp = part_s + (ks * LR_BM + s1_rg * 4) * rU + s1_jg * 4;
#pragma unroll
      for (int r = 0; r < 4; ++r) *reinterpret_cast<float4*>(pp + r * rU) = make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
    }
    for (int e = tid; e < witems; e += LR_THREADS) {
      const int row = e / wjp, jp = e - row * wjp;
      if (row >= bm) break;
      const float* xr = xc + row * I;
      const float* wp = W1s + jp * 2;
      float a0 = 0.f, a1 = 0.f;
      for (int k = 0; k < I; k += 4) {
        const float4 xv = *reinterpret_cast<const float4*>(xr + k);
        const float2 w0 = *reinterpret_cast<const float2*>(wp + (k + 0) * rW), w1 = *reinterpret_cast<const float2*>(wp + (k + 1) * rW);
        const float2 w2 = *reinterpret_cast<const float2*>(wp + (k + 2) * rW), w3 = *reinterpret_cast<const float2*>(wp + (k + 3) * rW);
        a0 = fmaf(xv.x, w0.x, a0); a1 = fmaf(xv.x, w0.y, a1); a0 = fmaf(xv.y, w1.x, a0); a1 = fmaf(xv.y, w1.y, a1);
        a0 = fmaf(xv.z, w2.x, a0); a1 = fmaf(xv.z, w2.y, a1); a0 = fmaf(xv.w, w3.x, a0); a1 = fmaf(xv.w, w3.y, a1);
      }
      *reinterpret_cast<float2*>(s_s + row * SP + rU + jp * 2) = make_float2(a0, a1);
    }
    __syncthreads();
    for (int e = tid; e < ritems; e += LR_THREADS) {             // s = sum of the K-slice partials, fixed order
      const int row = e / ujg, jg = e - row * ujg;
      if (row >= bm) break;
      float4 v = *reinterpret_cast<const float4*>(part_s + row * rU + jg * 4);
      for (int q = 1; q < KS; ++q) {
        const float4 w = *reinterpret_cast<const float4*>(part_s + (q * LR_BM + row) * rU + jg * 4);
        v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w;
      }
      *reinterpret_cast<float4*>(s_s + row * SP + jg * 4) = v;
    }
    __syncthreads();
    // ---- stage 2: 8 rows x 4 units ----
    float acc[8][4] = {};
    if (s2_active) {
      const float* sp = s_s + (rg * 8) * SP;
      const float* vp = V2s + n0;
#pragma unroll 1
      for (int j = 0; j < R; j += 4) {
        float4 vv[4];
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) vv[jj] = *reinterpret_cast<const float4*>(vp + (j + jj) * LR_H);
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          const float4 sv = *reinterpret_cast<const float4*>(sp + r * SP + j);
          lr_fma4(acc[r], sv.x, vv[0]); lr_fma4(acc[r], sv.y, vv[1]); lr_fma4(acc[r], sv.z, vv[2]); lr_fma4(acc[r], sv.w, vv[3]);
        }
      }
    }
    // ---- gate update (rnn.py:289-295) ----
    const bool last = t == d.T - 1;
    if (s2_active) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      float* hrow = h_s + (rg * 8 + r) * LR_HP + n0;
      const float4 ho = *reinterpret_cast<const float4*>(hrow);
      const float hold[4] = {ho.x, ho.y, ho.z, ho.w};
      const float bgv[4] = {bg.x, bg.y, bg.z, bg.w}, buv[4] = {bu.x, bu.y, bu.z, bu.w};
      float hn[4], zv[4], cv[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float pre = acc[r][c];
        zv[c] = FAST_NL ? sigmoid_fast(pre + bgv[c]) : act_rt(d.gate_nl, pre + bgv[c]);
        cv[c] = FAST_NL ? tanh_fast(pre + buv[c]) : act_rt(d.update_nl, pre + buv[c]);
        hn[c] = fmaf(zv[c], hold[c], (fmaf(sz, 1.0f - zv[c], sn)) * cv[c]);
      }
      *reinterpret_cast<float4*>(hrow) = make_float4(hn[0], hn[1], hn[2], hn[3]);
      if (r < rows_left) {
        if (outp) *reinterpret_cast<float4*>(outp + (size_t)r * a.osb) = make_float4(hn[0], hn[1], hn[2], hn[3]);
        if (zp) {
          *reinterpret_cast<float4*>(zp + (size_t)r * LR_H) = make_float4(zv[0], zv[1], zv[2], zv[3]);
          *reinterpret_cast<float4*>(cp + (size_t)r * LR_H) = make_float4(cv[0], cv[1], cv[2], cv[3]);
        }
        if (last && a.h_last) *reinterpret_cast<float4*>(a.h_last + (size_t)(first_row + r) * LR_H + n0) = make_float4(hn[0], hn[1], hn[2], hn[3]);
      }
    }
    }
    if (outp) outp += a.ost;
    if (zp) { zp += (size_t)d.B * LR_H; cp += (size_t)d.B * LR_H; }
    if (more) {
#pragma unroll
      for (int q = 0; q < LR_MAX_X4; ++q)
        if (x_dst[q] >= 0) *reinterpret_cast<float4*>(xn + x_dst[q]) = xnext[q];
    }
    __syncthreads();
  }
}

size_t lr_fwd_smem_bytes(const Dims& d) { return sizeof(float) * (size_t)lr_smem_layout(d.I, d.rW, d.rU).total; }

bool lr_path_supports(const Dims& d) {
  if (d.H != LR_H || d.rW <= 0 || d.rU <= 0) return false;
  if (d.rW % 4 || d.rU % 4 || d.I % 4 || d.I < 4 || d.I > 64 || d.rW > 32 || d.rU > 64) return false;
  return lr_fwd_smem_bytes(d) <= 227 * 1024;
}

// Rows per CTA: the batch is spread over whole rounds of 148 CTAs (C4: 32768 rows -> 4 rounds of 56-row CTAs
// instead of 3.46 waves of 64-row ones); small batches get proportionally short tiles and a shorter step.
static int lr_rows_per_cta(int B) {
  const int rounds = (B + 148 * LR_BM - 1) / (148 * LR_BM);
  int bm = (B + 148 * rounds - 1) / (148 * rounds);
  bm = (bm + 7) & ~7;
  return bm < 8 ? 8 : (bm > LR_BM ? LR_BM : bm);
}

int launch_lr_fwd(const FwdArgs& a, cudaStream_t stream) {
  if (a.d.B <= 0 || a.d.T <= 0) return FGRNN_OK;
  const size_t smem = lr_fwd_smem_bytes(a.d);
  const bool fast = a.d.gate_nl == FGRNN_NL_SIGMOID && a.d.update_nl == FGRNN_NL_TANH;
  const int bm = lr_rows_per_cta(a.d.B);
  const unsigned grid = (unsigned)((a.d.B + bm - 1) / bm);
  const bool c4 = a.d.rU == 32 && a.d.rW == 16 && a.d.I == 32;
  auto go = [&](auto kern) -> int {
    FGRNN_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, LR_THREADS, smem, stream>>>(a, bm);
    return FGRNN_OK;
  };
  int rc;
  if (fast) rc = c4 ? go(lr_fwd_kernel<true, 32, 16, 32>) : go(lr_fwd_kernel<true, 0, 0, 0>);
  else rc = c4 ? go(lr_fwd_kernel<false, 32, 16, 32>) : go(lr_fwd_kernel<false, 0, 0, 0>);
  if (rc) return rc;
  FGRNN_LAUNCH_CHECK("lr_fwd_kernel");
  return FGRNN_OK;
}

}  // namespace fgrnn
