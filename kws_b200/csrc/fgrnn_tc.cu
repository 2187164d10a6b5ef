// tcgen05 / TMEM kernel family (FGRNN_PATH_TCGEN05): the FastGRNN recurrence on the 5th-gen tensor cores.
//
// Formulation (weights-stationary, transposed):  pre_t^T = W^T.x_t^T + U^T.h_{t-1}^T   (rnn.py:277-289)
//   A operand = the weights, resident in TENSOR MEMORY for the whole kernel (lane = hidden unit n,
//               32-bit column c = fp16 pair k = 2c, 2c+1);  M = H = 128 is always a full-rate UMMA M
//   B operand = the streamed data in shared memory: x_t tile (K-major) and h_{t-1} tile (MN-major),
//               N = 32 batch rows per sub-tile
//   D         = fp32 accumulators in tensor memory, lane = hidden unit, column = batch row
//
// fp32 parity on fp16 tensor cores (measured and modelled in tools/tc_probe.cu, tools/fit_mma_model.py,
// tools/emulate_tc_schemes.py):
//   * every fp32 operand is split  v*2^s = hi + lo  (fp16, round-to-nearest, power-of-two pre-scale) and
//     each product is three MMAs  hi.hi + lo.hi + hi.lo  (lo.lo ~ 2^-22 is dropped);
//   * one tcgen05.mma aligns its 16 products and the accumulator to the largest exponent, TRUNCATES every
//     addend at 2^(emax-25) and truncates the sum to fp32 -- a toward-zero bias per MMA.  A single chain of
//     30 MMAs lands at 1.5-1.9x the (rtol 1e-5, atol 1e-6) budget.  So the small lo terms accumulate in
//     their own accumulator C (their truncation error is 2^-11 smaller) and the hi.hi terms are split over
//     two short chains M1, M2; the epilogue adds the three in fp32 round-to-nearest:  0.54-0.62 of the
//     budget against the oracle, the same as the FFMA kernel.
//
// One CTA = 64 batch rows = two sub-tiles of 32 that ping-pong: while the epilogue warps work on one
// sub-tile the tensor core runs the other one's 30 MMAs (18 cycles each at N = 32, A in TMEM).
//   warp 0      : MMA issuer (one elected lane), TMEM allocation
//   warps 1..3  : x path: TMA (cp.async.bulk.tensor, 3-D map over [B,T,I] by strides) -> raw ring ->
//                 fp16 hi/lo split -> K-major operand tiles
//   warps 4..11 : epilogue: tcgen05.ld C, M1, M2 -> gate update in registers (state h lives in registers,
//                 thread = hidden unit, 16 rows per sub-tile) -> STG h_t (128 B per warp per row) ->
//                 fp16 split -> 16-byte st.shared into the MN-major operand tile -> fence.proxy.async ->
//                 mbarrier
#include <cuda.h>
#include <cuda_fp16.h>

#include "fgrnn_kernels.cuh"

namespace fgrnn {

constexpr int TC_H = 128;                      // hidden size = UMMA M
constexpr int TC_NS = 32;                      // batch rows per sub-tile = UMMA N
constexpr int TC_NT = 2;                       // sub-tiles per CTA
constexpr int TC_ROWS = TC_NS * TC_NT;         // 64 batch rows per CTA
constexpr int TC_CONV_WARPS = 3;
constexpr int TC_EPI_WARPS = 8;
constexpr int TC_THREADS = 32 * (1 + TC_CONV_WARPS + TC_EPI_WARPS);   // 384
constexpr int TC_RAW_STAGES = 4;
constexpr int TC_MAX_KI = 64;                  // input features padded to a multiple of 16, <= 64
// tensor-memory column map
constexpr int TM_U_HI = 0, TM_U_LO = 64, TM_W_HI = 128, TM_W_LO = 160, TM_ACC = 192;
constexpr int TM_ACC_PER_TILE = 3 * TC_NS;     // C | M1 | M2
constexpr int TC_TMEM_COLS = 512;
constexpr int TC_SH_EXP = 4;                   // h is scaled by 2^4 before the fp16 split

// ---- raw PTX wrappers -------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  for (uint32_t spins = 0; !ok; ++spins) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (spins > (1u << 26)) __trap();      // protocol bug guard: fail loudly instead of hanging the GPU
  }
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[tmem] . B[smem], fp16 inputs, fp32 accumulate.  Executed by a whole warp; the instruction
// itself is predicated on one elected lane (straight-line SASS: ELECT / R2UR / UTCHMMA, 18 cycles per MMA
// at N = 32 -- an `if (lane == 0)` around it costs 45, tools/tc_probe2.cu).
__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p, q;\n\telect.sync _|q, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
               ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// arrives on the mbarrier when every MMA issued so far by the elected lane has completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t"
               "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr) : "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tma_load_3d(uint32_t dst_smem, const CUtensorMap* map, int c0, int c1, int c2, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
               ::"r"(dst_smem), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(bar) : "memory");
}

// ---- operand layouts (SWIZZLE_NONE canonical layouts, 128-byte core matrices) -------------------
// x tile, K-major [rows][KI]: core matrix = 8 rows x 16 B (8 k);  next 8 k: +128 B (LBO);  next 8 rows: +(KI/8)*128 B (SBO)
__device__ __forceinline__ uint64_t make_desc_kmajor(uint32_t smem_addr, int KI) {
  const uint64_t lbo = 128 >> 4, sbo = (uint64_t)((KI >> 3) * 128) >> 4;
  return (uint64_t)((smem_addr >> 4) & 0x3fff) | (lbo << 16) | (sbo << 32) | (1ull << 46);
}
// h tile, MN-major [k][rows]: core matrix = 8 k x 16 B (8 rows);  next 8 rows: +128 B (SBO);  next 8 k: +(NS/8)*128 B (LBO)
__device__ __forceinline__ uint64_t make_desc_mnmajor(uint32_t smem_addr) {
  const uint64_t sbo = 128 >> 4, lbo = (uint64_t)((TC_NS >> 3) * 128) >> 4;
  return (uint64_t)((smem_addr >> 4) & 0x3fff) | (lbo << 16) | (sbo << 32) | (1ull << 46);
}
constexpr uint32_t TC_H_KSTEP = (2 * (TC_NS >> 3) * 128) >> 4;     // descriptor advance per 16 k of the h tile
constexpr uint32_t TC_X_KSTEP = 256 >> 4;                          // ... of the x tile
// kind::f16: D fp32, A/B fp16, A K-major (TMEM), N = 32, M = 128; bit 16 = B is MN-major
constexpr uint32_t TC_IDESC_X = (1u << 4) | ((uint32_t)(TC_NS >> 3) << 17) | ((uint32_t)(TC_H >> 4) << 24);
constexpr uint32_t TC_IDESC_H = TC_IDESC_X | (1u << 16);

// v*scale = hi + lo with hi, lo fp16 (round to nearest); two values at once
__device__ __forceinline__ void split2(float a, float b, float scale, uint32_t& hi, uint32_t& lo) {
  a *= scale; b *= scale;
  a = fminf(fmaxf(a, -65504.f), 65504.f);
  b = fminf(fmaxf(b, -65504.f), 65504.f);
  const __half2 h = __floats2half2_rn(a, b);
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

struct TcArgs {
  SmemFwdArgs f;
  int KI;               // I rounded up to a multiple of 16
  int x_time_outer;     // tensor-map dimension order: 0 = {I, T, B}, 1 = {I, B, T}
};

struct TcSmemLayout {
  int h_op, x_op, raw, bars, misc, total;
  int x_tile_bytes, raw_stage_bytes;
};
__host__ __device__ inline TcSmemLayout tc_smem_layout(int I, int KI, int esz) {
  TcSmemLayout L;
  L.x_tile_bytes = TC_NS * KI * 2;
  L.raw_stage_bytes = TC_ROWS * I * esz;
  L.h_op = 0;                                                   // [NT][hi|lo][NS*128*2]
  L.x_op = L.h_op + TC_NT * 2 * TC_NS * TC_H * 2;               // [2 buffers][NT][hi|lo][x_tile_bytes]
  L.raw = L.x_op + 2 * TC_NT * 2 * L.x_tile_bytes;              // [RAW_STAGES][raw_stage_bytes], 128-byte aligned
  L.raw = (L.raw + 127) & ~127;
  L.bars = L.raw + TC_RAW_STAGES * L.raw_stage_bytes;
  L.bars = (L.bars + 15) & ~15;
  L.misc = L.bars + 24 * 8;
  L.total = L.misc + 256;
  return L;
}

// split without the range clamp (|v*scale| < 65504 is guaranteed by the caller)
__device__ __forceinline__ void split2_nc(float a, float b, float scale, uint32_t& hi, uint32_t& lo) {
  a *= scale; b *= scale;
  const __half2 h = __floats2half2_rn(a, b);
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

// gate update for one element (rnn.py:289-295): tot = 2^S * pre.
//   e_g = exp(-(pre + b_g)), e_u = exp(-2 (pre + b_u));  z = 1/(1+e_g);  c = (1-e_u)/(1+e_u);  one MUFU.RCP serves
//   both:  r = 1/((1+e_g)(1+e_u)).  tot is clamped from below (per-unit constant tmin) so that both exponents
//   stay <= 60 and the product stays finite; at the clamp z < 1e-18 and c = -1 to fp32 precision.
struct EpiConst { float kS, k2S, cg, cu, sz, szn, tmin; };
__device__ __forceinline__ float gate_update(float tot, float h, const EpiConst& k, float& z_out, float& c_out) {
  tot = fmaxf(tot, k.tmin);
  const float eg = ex2_approx(fmaf(tot, k.kS, k.cg));          // -(pre + b_g) * log2(e)
  const float eu = ex2_approx(fmaf(tot, k.k2S, k.cu));         // -2 (pre + b_u) * log2(e)
  const float a = 1.0f + eg, b = 1.0f + eu;
  const float r = rcp_approx(a * b);
  const float z = r * b;                                       // sigmoid(pre + b_g)           rnn.py:290
  const float c = (1.0f - eu) * (r * a);                       // tanh(pre + b_u)              rnn.py:292
  z_out = z; c_out = c;
  return fmaf(z, h, fmaf(-k.sz, z, k.szn) * c);                // z h + (sz (1 - z) + sn) c    rnn.py:294-295
}

// Epilogue main loop of one warp.  Thread = hidden unit n; it owns rows [rh*16, rh*16+16) of both sub-tiles.
struct EpiCtx {
  uint32_t bar_dfull, bar_hready;     // shared addresses of the [NT] barrier arrays
  uint32_t acc;                       // TMEM address: lane quadrant | TM_ACC + rh*16
  unsigned char* hop;                 // operand-tile address of (sub-tile 0, row group rh*2, k = n), hi part
  float* out;                         // &out[row0 + rh*16][t = 0][n]  (or null)
  float* zs; float* cs;               // &save[t = 0][row0 + rh*16][n]
  uint32_t out_row, out_step;         // element strides of `out`
  uint32_t zc_step;                   // B*H
  int rows_left;                      // B - (row0 + rh*16): rows >= rows_left (per sub-tile offset) are padding
  int T;
  float scale_h;
};

template <bool HAS_OUT, bool SAVE, bool MASKED>
__device__ __forceinline__ void epilogue_loop(const EpiCtx& cx, const EpiConst& kc, float (&hst)[TC_NT][16]) {
  float* outp = cx.out; float* zp = cx.zs; float* cp = cx.cs;
  for (int t = 0; t < cx.T; ++t) {
#pragma unroll
    for (int s = 0; s < TC_NT; ++s) {
      mbar_wait(cx.bar_dfull + s * 8, t & 1);
      tc_fence_after();
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        float vc[8], v1[8], v2[8];
        const uint32_t acc = cx.acc + s * TM_ACC_PER_TILE + g * 8;
        tmem_ld8(acc, vc);
        tmem_ld8(acc + TC_NS, v1);
        tmem_ld8(acc + 2 * TC_NS, v2);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int rj = s * TC_NS + g * 8 + j;                 // row offset from this thread's first row
          const float tot = (vc[j] + v2[j]) + v1[j];
          float z, c;
          const float hn = gate_update(tot, hst[s][g * 8 + j], kc, z, c);
          hst[s][g * 8 + j] = hn;
          if (!MASKED || rj < cx.rows_left) {
            if (HAS_OUT) outp[(uint32_t)rj * cx.out_row] = hn;
            if (SAVE) { zp[rj * TC_H] = z; cp[rj * TC_H] = c; }
          }
        }
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) split2_nc(hst[s][g * 8 + 2 * q], hst[s][g * 8 + 2 * q + 1], cx.scale_h, hi[q], lo[q]);
        unsigned char* p = cx.hop + s * (2 * TC_NS * TC_H * 2) + g * 128;
        *reinterpret_cast<uint4*>(p) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<uint4*>(p + TC_NS * TC_H * 2) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
      }
      fence_proxy_async_smem();                        // st.shared of the h tile -> visible to tcgen05.mma
      tc_fence_before();                               // tcgen05.ld of D done before the next MMAs overwrite it
      __syncwarp();
      if ((threadIdx.x & 31) == 0) mbar_arrive(cx.bar_hready + s * 8);
    }
    if (HAS_OUT) outp += cx.out_step;
    if (SAVE) { zp += cx.zc_step; cp += cx.zc_step; }
  }
}

__global__ void __launch_bounds__(TC_THREADS, 1) tc_fwd_kernel(const TcArgs ta, const __grid_constant__ CUtensorMap xmap) {
  extern __shared__ __align__(128) unsigned char sm[];
  const SmemFwdArgs& a = ta.f;
  const Dims d = a.d;
  const int I = d.I, KI = ta.KI;
  const int esz = d.x_dtype == FGRNN_BF16 ? 2 : 4;
  const TcSmemLayout L = tc_smem_layout(I, KI, esz);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + L.bars);
  uint32_t* tmem_base_s = reinterpret_cast<uint32_t*>(sm + L.misc);
  float* red_s = reinterpret_cast<float*>(sm + L.misc + 16);            // [2][12] max|U|, max|W| per warp

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row0 = blockIdx.x * TC_ROWS;
  const bool hi_layout = a.layout == FGRNN_LAYOUT_HI;
  // barrier map
  auto bar = [&](int i) { return smem_u32(&bars[i]); };
  const int B_HREADY = 0, B_DFULL = 2, B_XFULL = 4, B_XEMPTY = 6, B_RAWFULL = 8, B_RAWEMPTY = 12;

  // ---- prologue ---------------------------------------------------------------------------------
  if (warp == 0) tmem_alloc(smem_u32(tmem_base_s), TC_TMEM_COLS);
  if (tid == 32) {
    for (int s = 0; s < TC_NT; ++s) { mbar_init(bar(B_HREADY + s), TC_EPI_WARPS); mbar_init(bar(B_DFULL + s), 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(bar(B_XFULL + b), TC_CONV_WARPS); mbar_init(bar(B_XEMPTY + b), 1); }
    for (int st = 0; st < TC_RAW_STAGES; ++st) { mbar_init(bar(B_RAWFULL + st), 1); mbar_init(bar(B_RAWEMPTY + st), TC_CONV_WARPS); }
    fence_mbar_init();
  }
  // power-of-two operand scales from max|U|, max|W| (identical in every CTA)
  float mu = 0.f, mw = 0.f;
  for (int e = tid; e < TC_H * TC_H; e += TC_THREADS) mu = fmaxf(mu, fabsf(__ldg(a.U + e)));
  for (int e = tid; e < TC_H * I; e += TC_THREADS) mw = fmaxf(mw, fabsf(__ldg(a.W + e)));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mu = fmaxf(mu, __shfl_xor_sync(0xffffffffu, mu, o));
    mw = fmaxf(mw, __shfl_xor_sync(0xffffffffu, mw, o));
  }
  if (lane == 0) { red_s[warp] = mu; red_s[12 + warp] = mw; }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_base_s;
  for (int w = 0; w < TC_THREADS / 32; ++w) { mu = fmaxf(mu, red_s[w]); mw = fmaxf(mw, red_s[12 + w]); }
  // accumulators hold 2^S * pre:  (h*2^4).(U*2^(S-4))  and  (x*2^0).(W*2^S)
  int S = 40;
  if (mw > 0.f) S = min(S, (int)floorf(log2f(30000.f / mw)));
  if (mu > 0.f) S = min(S, (int)floorf(log2f(30000.f / mu)) + TC_SH_EXP);
  S = max(S, TC_SH_EXP - 14);
  const float scale_w = exp2f((float)S), scale_u = exp2f((float)(S - TC_SH_EXP)), scale_h = exp2f((float)TC_SH_EXP);
  const float unscale = exp2f((float)-S);

  if (warp == 0) {
    // =========================== MMA issuer =====================================================
    tc_fence_before();
    __syncthreads();                                   // weights in TMEM, h_{-1} / x_0 tiles under way
    tc_fence_after();
    const int nkx = KI >> 4;
    const bool x_has_lo = d.x_dtype != FGRNN_BF16;     // a bf16 value is one exact fp16 (plus an exact zero lo)
    uint64_t dH[TC_NT][2];
#pragma unroll
    for (int s = 0; s < TC_NT; ++s)
#pragma unroll
      for (int p = 0; p < 2; ++p) dH[s][p] = make_desc_mnmajor(smem_u32(sm + L.h_op + (s * 2 + p) * TC_NS * TC_H * 2));
    for (int t = 0; t < d.T; ++t) {
      const int xb = t & 1;
      mbar_wait(bar(B_XFULL + xb), (t >> 1) & 1);      // x_t operand tiles written
#pragma unroll
      for (int s = 0; s < TC_NT; ++s) {
        const uint64_t dXhi = make_desc_kmajor(smem_u32(sm + L.x_op + ((xb * TC_NT + s) * 2 + 0) * L.x_tile_bytes), KI);
        const uint64_t dXlo = make_desc_kmajor(smem_u32(sm + L.x_op + ((xb * TC_NT + s) * 2 + 1) * L.x_tile_bytes), KI);
        const uint32_t accC = tmem + TM_ACC + s * TM_ACC_PER_TILE, accM1 = accC + TC_NS, accM2 = accC + 2 * TC_NS;
        mbar_wait(bar(B_HREADY + s), t & 1);           // h_{t-1} operand tile written, D of step t-1 drained
        tc_fence_after();
        // C: the lo terms
        uint32_t acc = 0;
        for (int ks = 0; ks < nkx; ++ks) {
          umma_ts(accC, tmem + TM_W_LO + ks * 8, dXhi + ks * TC_X_KSTEP, TC_IDESC_X, acc); acc = 1;
          if (x_has_lo) umma_ts(accC, tmem + TM_W_HI + ks * 8, dXlo + ks * TC_X_KSTEP, TC_IDESC_X, 1);
        }
#pragma unroll
        for (int ks = 0; ks < TC_H / 16; ++ks) {
          umma_ts(accC, tmem + TM_U_LO + ks * 8, dH[s][0] + ks * TC_H_KSTEP, TC_IDESC_H, 1);
          umma_ts(accC, tmem + TM_U_HI + ks * 8, dH[s][1] + ks * TC_H_KSTEP, TC_IDESC_H, 1);
        }
        // M1: hi.hi of x.W and of the first three k-steps of h.U;  M2: the other five
        acc = 0;
        for (int ks = 0; ks < nkx; ++ks) { umma_ts(accM1, tmem + TM_W_HI + ks * 8, dXhi + ks * TC_X_KSTEP, TC_IDESC_X, acc); acc = 1; }
#pragma unroll
        for (int ks = 0; ks < 3; ++ks) umma_ts(accM1, tmem + TM_U_HI + ks * 8, dH[s][0] + ks * TC_H_KSTEP, TC_IDESC_H, 1);
#pragma unroll
        for (int ks = 3; ks < TC_H / 16; ++ks) umma_ts(accM2, tmem + TM_U_HI + ks * 8, dH[s][0] + ks * TC_H_KSTEP, TC_IDESC_H, ks > 3);
        umma_commit(bar(B_DFULL + s));                 // implies tcgen05.fence::before_thread_sync
      }
      umma_commit(bar(B_XEMPTY + xb));                 // both sub-tiles have consumed the x_t tiles
    }
  } else if (warp <= TC_CONV_WARPS) {
    // =========================== x path: TMA -> split -> operand tiles ============================
    const int cw = warp - 1;
    const uint32_t raw_bytes = (uint32_t)L.raw_stage_bytes;
    auto issue_tma = [&](int t) {
      const int st = t % TC_RAW_STAGES;
      mbar_expect_tx(bar(B_RAWFULL + st), raw_bytes);
      if (ta.x_time_outer) tma_load_3d(smem_u32(sm + L.raw + st * L.raw_stage_bytes), &xmap, 0, row0, t, bar(B_RAWFULL + st));
      else tma_load_3d(smem_u32(sm + L.raw + st * L.raw_stage_bytes), &xmap, 0, t, row0, bar(B_RAWFULL + st));
    };
    if (cw == 0 && lane == 0)
      for (int t = 0; t < TC_RAW_STAGES && t < d.T; ++t) issue_tma(t);
    tc_fence_before();
    __syncthreads();
    const int nch = KI >> 3, ntask = TC_ROWS * nch;
    for (int t = 0; t < d.T; ++t) {
      const int st = t % TC_RAW_STAGES, xb = t & 1;
      mbar_wait(bar(B_RAWFULL + st), (t / TC_RAW_STAGES) & 1);
      mbar_wait(bar(B_XEMPTY + xb), ((t >> 1) & 1) ^ 1);       // MMAs of step t-2 have finished with this buffer
      const unsigned char* raw = sm + L.raw + st * L.raw_stage_bytes;
      for (int e = cw * 32 + lane; e < ntask; e += TC_CONV_WARPS * 32) {
        const int row = e / nch, ch = e - row * nch;
        float v[8];
        if (ch * 8 < I) {
          if (esz == 4) {
            const float4 p0 = *reinterpret_cast<const float4*>(raw + (size_t)row * I * 4 + ch * 32);
            const float4 p1 = *reinterpret_cast<const float4*>(raw + (size_t)row * I * 4 + ch * 32 + 16);
            v[0] = p0.x; v[1] = p0.y; v[2] = p0.z; v[3] = p0.w; v[4] = p1.x; v[5] = p1.y; v[6] = p1.z; v[7] = p1.w;
          } else {
            const uint4 p = *reinterpret_cast<const uint4*>(raw + (size_t)row * I * 2 + ch * 16);
            const uint32_t w[4] = {p.x, p.y, p.z, p.w};
#pragma unroll
            for (int q = 0; q < 4; ++q) { v[2 * q] = __uint_as_float(w[q] << 16); v[2 * q + 1] = __uint_as_float(w[q] & 0xffff0000u); }
          }
        } else {
#pragma unroll
          for (int q = 0; q < 8; ++q) v[q] = 0.f;
        }
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) split2(v[2 * q], v[2 * q + 1], 1.0f, hi[q], lo[q]);
        const int s = row / TC_NS, r = row - s * TC_NS;
        unsigned char* xt = sm + L.x_op + ((xb * TC_NT + s) * 2) * L.x_tile_bytes + (r >> 3) * (nch * 128) + ch * 128 + (r & 7) * 16;
        *reinterpret_cast<uint4*>(xt) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<uint4*>(xt + L.x_tile_bytes) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
      }
      fence_proxy_async_smem();                        // st.shared of the operand tiles -> visible to tcgen05.mma
      __syncwarp();
      if (lane == 0) { mbar_arrive(bar(B_XFULL + xb)); mbar_arrive(bar(B_RAWEMPTY + st)); }
      if (cw == 0 && lane == 0 && t + TC_RAW_STAGES < d.T) {
        mbar_wait(bar(B_RAWEMPTY + st), (t / TC_RAW_STAGES) & 1);      // all converter warps are done with this stage
        fence_proxy_async_smem();
        issue_tma(t + TC_RAW_STAGES);
      }
      __syncwarp();
    }
  } else {
    // =========================== epilogue warps ===================================================
    const int ew = warp - 1 - TC_CONV_WARPS;           // 0..7
    const int quad = warp & 3;                         // TMEM lane quadrant this warp may access (= ew & 3)
    const int rh = ew >> 2;                            // which 16 of the sub-tile's 32 rows
    const int n = quad * 32 + lane;                    // hidden unit = TMEM lane
    const uint32_t lane_base = (uint32_t)(quad * 32) << 16;

    // weights -> tensor memory (A operands): this thread's row of U^T / W^T, fp16 hi/lo pairs along k
    {
      const int k_begin = rh * (TC_H / 2);
      for (int kc = 0; kc < TC_H / 2; kc += 16) {
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int k = k_begin + kc + 2 * j;
          const float v0 = hi_layout ? __ldg(a.U + (size_t)n * TC_H + k) : __ldg(a.U + (size_t)k * TC_H + n);
          const float v1 = hi_layout ? __ldg(a.U + (size_t)n * TC_H + k + 1) : __ldg(a.U + (size_t)(k + 1) * TC_H + n);
          split2(v0, v1, scale_u, hi[j], lo[j]);
        }
        tmem_st8(tmem + lane_base + TM_U_HI + ((k_begin + kc) >> 1), hi);
        tmem_st8(tmem + lane_base + TM_U_LO + ((k_begin + kc) >> 1), lo);
      }
      for (int kc = rh * 16; kc < KI; kc += 32) {
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int k = kc + 2 * j;
          float v0 = 0.f, v1 = 0.f;
          if (k < I) v0 = hi_layout ? __ldg(a.W + (size_t)n * I + k) : __ldg(a.W + (size_t)k * TC_H + n);
          if (k + 1 < I) v1 = hi_layout ? __ldg(a.W + (size_t)n * I + k + 1) : __ldg(a.W + (size_t)(k + 1) * TC_H + n);
          split2(v0, v1, scale_w, hi[j], lo[j]);
        }
        tmem_st8(tmem + lane_base + TM_W_HI + (kc >> 1), hi);
        tmem_st8(tmem + lane_base + TM_W_LO + (kc >> 1), lo);
      }
      tmem_st_wait();
    }

    EpiConst kc;
    {
      const float sz = sigmoid_f(__ldg(a.zeta)), sn = sigmoid_f(__ldg(a.nu));
      constexpr float LOG2E = 1.4426950408889634f;
      kc.kS = -LOG2E * unscale; kc.k2S = -2.0f * LOG2E * unscale;
      kc.cg = -LOG2E * __ldg(a.bias_gate + n); kc.cu = -2.0f * LOG2E * __ldg(a.bias_update + n);
      kc.sz = sz; kc.szn = sz + sn;
      // tot >= tmin  <=>  both exponents <= 60
      kc.tmin = fmaxf((60.0f - kc.cg) / kc.kS, (60.0f - kc.cu) / kc.k2S);
    }

    // state: this thread owns h[row][n] for 16 rows of each sub-tile
    float hst[TC_NT][16];
    unsigned char* hop = sm + L.h_op + (n >> 3) * ((TC_NS >> 3) * 128) + (rh * 2) * 128 + (n & 7) * 16;
#pragma unroll
    for (int s = 0; s < TC_NT; ++s) {
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int row = row0 + s * TC_NS + rh * 16 + j;
        hst[s][j] = (a.h0 && row < d.B) ? __ldg(a.h0 + (size_t)row * TC_H + n) : 0.f;
      }
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) split2(hst[s][g * 8 + 2 * q], hst[s][g * 8 + 2 * q + 1], scale_h, hi[q], lo[q]);
        unsigned char* p = hop + s * (2 * TC_NS * TC_H * 2) + g * 128;
        *reinterpret_cast<uint4*>(p) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<uint4*>(p + TC_NS * TC_H * 2) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
      }
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();                                   // matches the other roles' prologue barrier
    __syncwarp();
    if (lane == 0) { mbar_arrive(bar(B_HREADY + 0)); mbar_arrive(bar(B_HREADY + 1)); }   // phase 0: h_{-1} ready

    EpiCtx cx;
    cx.bar_dfull = bar(B_DFULL); cx.bar_hready = bar(B_HREADY);
    cx.acc = tmem + lane_base + TM_ACC + rh * 16;
    cx.hop = hop;
    const int first_row = row0 + rh * 16;
    cx.out = a.out ? a.out + (size_t)first_row * a.osb + n : nullptr;
    cx.zs = a.save_z ? a.save_z + (size_t)first_row * TC_H + n : nullptr;
    cx.cs = a.save_c ? a.save_c + (size_t)first_row * TC_H + n : nullptr;
    cx.out_row = (uint32_t)a.osb; cx.out_step = (uint32_t)a.ost; cx.zc_step = (uint32_t)d.B * TC_H;
    cx.rows_left = d.B - first_row; cx.T = d.T; cx.scale_h = scale_h;
    const bool masked = row0 + TC_ROWS > d.B;
    const int variant = (a.out ? 4 : 0) | (a.save_z ? 2 : 0) | (masked ? 1 : 0);
    switch (variant) {
      case 0: epilogue_loop<false, false, false>(cx, kc, hst); break;
      case 1: epilogue_loop<false, false, true>(cx, kc, hst); break;
      case 2: epilogue_loop<false, true, false>(cx, kc, hst); break;
      case 3: epilogue_loop<false, true, true>(cx, kc, hst); break;
      case 4: epilogue_loop<true, false, false>(cx, kc, hst); break;
      case 5: epilogue_loop<true, false, true>(cx, kc, hst); break;
      case 6: epilogue_loop<true, true, false>(cx, kc, hst); break;
      default: epilogue_loop<true, true, true>(cx, kc, hst); break;
    }
    if (a.h_last) {
#pragma unroll
      for (int s = 0; s < TC_NT; ++s)
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int row = first_row + s * TC_NS + j;
          if (row < d.B) a.h_last[(size_t)row * TC_H + n] = hst[s][j];
        }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, TC_TMEM_COLS);
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = []() -> EncodeTiledFn {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) return nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

bool tc_path_supports(const Dims& d) {
  // full-rank, H = 128, I a multiple of 8 up to 64, sigmoid gate / tanh update
  return d.rW == 0 && d.rU == 0 && d.H == TC_H && d.I >= 8 && d.I <= TC_MAX_KI && (d.I % 8) == 0 &&
         d.gate_nl == FGRNN_NL_SIGMOID && d.update_nl == FGRNN_NL_TANH;
}

// TMA needs a 16-byte aligned base and 16-byte multiples for the batch / time strides
bool tc_x_tma_ok(const void* x, int64_t xsb, int64_t xst, int x_dtype, int B, int T) {
  const int esz = x_dtype == FGRNN_BF16 ? 2 : 4;
  if (reinterpret_cast<uintptr_t>(x) & 15) return false;
  if ((xsb * esz) % 16 || (xst * esz) % 16) return false;
  (void)B; (void)T;
  return xsb > 0 && xst > 0;
}

int launch_tc_fwd(const SmemFwdArgs& a, cudaStream_t stream) {
  const Dims& d = a.d;
  if (d.B <= 0 || d.T <= 0) return FGRNN_OK;
  EncodeTiledFn encode = encode_tiled_fn();
  if (!encode) { set_error_detail("cuTensorMapEncodeTiled is not available from the driver"); return FGRNN_ERR_CUDA; }
  TcArgs ta{};
  ta.f = a;
  ta.KI = (d.I + 15) & ~15;
  const int esz = d.x_dtype == FGRNN_BF16 ? 2 : 4;
  ta.x_time_outer = a.xst > a.xsb ? 1 : 0;
  CUtensorMap map;
  cuuint64_t gdim[3], gstr[2];
  cuuint32_t box[3], estr[3] = {1, 1, 1};
  gdim[0] = (cuuint64_t)d.I;
  box[0] = (cuuint32_t)d.I;
  if (ta.x_time_outer) {
    gdim[1] = (cuuint64_t)d.B; gdim[2] = (cuuint64_t)d.T;
    gstr[0] = (cuuint64_t)a.xsb * esz; gstr[1] = (cuuint64_t)a.xst * esz;
    box[1] = TC_ROWS; box[2] = 1;
  } else {
    gdim[1] = (cuuint64_t)d.T; gdim[2] = (cuuint64_t)d.B;
    gstr[0] = (cuuint64_t)a.xst * esz; gstr[1] = (cuuint64_t)a.xsb * esz;
    box[1] = 1; box[2] = TC_ROWS;
  }
  const CUresult cr = encode(&map, d.x_dtype == FGRNN_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3,
                             const_cast<void*>(a.x), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                             CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) { set_error_detail("cuTensorMapEncodeTiled failed with CUresult %d", (int)cr); return FGRNN_ERR_CUDA; }
  const TcSmemLayout L = tc_smem_layout(d.I, ta.KI, esz);
  FGRNN_CUDA_TRY(cudaFuncSetAttribute(tc_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total));
  const unsigned grid = (unsigned)((d.B + TC_ROWS - 1) / TC_ROWS);
  tc_fwd_kernel<<<grid, TC_THREADS, L.total, stream>>>(ta, map);
  FGRNN_LAUNCH_CHECK("tc_fwd_kernel");
  return FGRNN_OK;
}

}  // namespace fgrnn
