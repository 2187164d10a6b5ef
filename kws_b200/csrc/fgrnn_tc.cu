// tcgen05 / TMEM kernel family (FGRNN_PATH_TCGEN05): the FastGRNN recurrence on the 5th-gen tensor cores.
//
// Formulation (weights-stationary, transposed):  pre_t^T = W^T.x_t^T + U^T.h_{t-1}^T   (rnn.py:277-289)
//   A operand = the weights, resident in TENSOR MEMORY for the whole kernel (lane = hidden unit n,
//               32-bit column c = fp16 pair k = 2c, 2c+1);  M = H = 128 is always a full-rate UMMA M
//   B operand = the streamed data in shared memory: x_t tile (K-major) and h_{t-1} tile (MN-major),
//               N = 32 batch rows per sub-tile (a tcgen05.mma with A in TMEM costs ~18 cycles for any N <= 32)
//   D         = fp32 accumulators in tensor memory, lane = hidden unit, column = batch row
//
// fp32 parity on fp16 tensor cores (measured and modelled in tools/tc_probe.cu, tools/fit_mma_model.py,
// tools/emulate_tc_schemes.py):
//   * every fp32 operand is split  v*2^s = hi + lo  (fp16, round-to-nearest, power-of-two pre-scale) and
//     each product is three MMAs  hi.hi + lo.hi + hi.lo  (lo.lo ~ 2^-22 is dropped);
//   * one tcgen05.mma aligns its 16 products and the accumulator to the largest exponent, TRUNCATES every
//     addend at 2^(emax-25) and truncates the sum to fp32 -- a toward-zero bias per MMA.  A single chain of
//     30 MMAs lands at 1.5-1.9x the (rtol 1e-5, atol 1e-6) budget.  So the small lo terms accumulate in
//     their own accumulators CA, CB (their truncation error is 2^-11 smaller) and the hi.hi terms are split over
//     two short chains M1, M2; the epilogue adds the four in fp32 round-to-nearest:  0.54-0.62 of the
//     budget against the oracle in emulation, 0.57-0.79 measured, the same class as the FFMA kernel.
//
// One CTA = 64 batch rows = two 32-row sub-tiles, each an independent step pipeline (MMA burst -> epilogue -> MMA
// burst ...) running half a period apart, so the tensor core works on one while the epilogue warps work on the other.
//   warps 0..15  : epilogue, 8 per sub-tile (4 TMEM lane quadrants x 2 row halves); thread = hidden unit, 16 rows:
//                  tcgen05.ld CA, CB, M1, M2 -> sum -> gate update on the packed fp32x2 pipe, one MUFU.EX2 and one
//                  MUFU.RCP per element, state h in registers for all T steps -> fp16 split -> 16-byte st.shared into
//                  the MN-major operand tile -> fence.proxy.async -> mbarrier -> (then) h_t into the warp's staging tile and out
//                  through one TMA tile store per warp (TC_TMA_STORE; scalar STG, 128 B per warp per row, when compiled out)
//   warps 16..19 : x path: TMA (cp.async.bulk.tensor, 3-D map over [B,T,I] by the caller's strides), 16 rows and a
//                  private 4-stage raw ring per warp -> fp16 hi/lo split -> K-major operand tiles (4 buffers)
//   warps 20..22 : MMA issuers: the 30 MMAs of a sub-tile step, 10 per warp, on one elected lane in a straight-line
//                  block; TMEM allocation (all 512 columns, so every tcgen05 address is a warp-uniform constant)
// Measured (tools/tc_trace.cu, profiles/r01_tc_fwd_ncu_summary.txt): the kernel is bound by the serial per-step chain
// of a sub-tile, not by HBM (34 %) or the tensor pipe (35 %).
#include <cstdlib>

#include "fgrnn_kernels.cuh"
#include "fgrnn_tc_common.cuh"

namespace fgrnn {

#ifdef TC_EXP_SCALAR_MATH      // experiment: scalar FFMA/FADD/FMUL instead of the packed fp32x2 instructions
__device__ __forceinline__ float2 s_ffma2(float2 a, float2 b, float2 c) { return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)); }
__device__ __forceinline__ float2 s_fadd2(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 s_fmul2(float2 a, float2 b) { return make_float2(a.x * b.x, a.y * b.y); }
#define __ffma2_rn s_ffma2
#define __fadd2_rn s_fadd2
#define __fmul2_rn s_fmul2
#endif

constexpr int TC_H = 128;                      // hidden size = UMMA M
constexpr int TC_CONV_WARPS = 4;                // each owns a quarter of the CTA's rows: own TMA box, own raw ring
constexpr int TC_XBUF = 4;                      // x operand tile buffers (the converters run up to 3 steps ahead)
constexpr int TC_EPI_WARPS = 16;                // 4 TMEM lane quadrants x (sub-tiles x row halves)
constexpr int TC_RAW_STAGES = 4;
constexpr int TC_MAX_KI = 64;                  // input features padded to a multiple of 16, <= 64
// tensor-memory column map
constexpr int TM_U_HI = 0, TM_U_LO = 64, TM_W_HI = 128, TM_W_LO = 160, TM_ACC = 192;
constexpr int TC_TMEM_COLS = 512;
#ifndef TC_CRIT_WAIT
#define TC_CRIT_WAIT mbar_wait      // waits on the step-critical path (mbar_wait_poll = spin without suspend hint)
#endif
#ifndef TC_NT4_ROLES
#define TC_NT4_ROLES 2
#endif
#ifndef TC_NT4_SPS
#define TC_NT4_SPS 2
#endif
#ifndef TC_STAGGER2_NS
#define TC_STAGGER2_NS (TC_STAGGER_NS / 2)
#endif
#ifndef TC_STAGGER_NS
#define TC_STAGGER_NS 500
#endif
#ifndef TC_TMA_STORE
#define TC_TMA_STORE 2          // 1 / 2: h_t (and z_t, c_t) leave through per-warp staging tiles and TMA tile stores (2: h_t is staged
                                // after the hand-off to the tensor core); 0: scalar st.global
#endif

// Developer trace (tools/tc_trace.cu defines FGRNN_TC_TRACE): clock64 stamps of CTA 0 for steps [16, 20)
#ifdef FGRNN_TC_TRACE
__device__ long long g_tc_trace[4 * 2 * 16];
#define TC_TRACE(t, s, slot)                                                                         \
  do {                                                                                               \
    if (blockIdx.x == 0 && (threadIdx.x & 31) == 0 && (t) >= 16 && (t) < 20)                         \
      g_tc_trace[(((t) - 16) * 2 + (s)) * 16 + (slot)] = clock64();                                  \
  } while (0)
__device__ unsigned long long g_tc_cta_time[1024 * 4];
__device__ __forceinline__ unsigned long long tc_globaltimer() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#define TC_CTA_TIME(slot) do { if (threadIdx.x == 128 && blockIdx.x < 1024) g_tc_cta_time[blockIdx.x * 4 + (slot)] = tc_globaltimer(); } while (0)
#else
#define TC_TRACE(t, s, slot) do { } while (0)
#define TC_CTA_TIME(slot) do { } while (0)
#endif

struct TcArgs {
  SmemFwdArgs f;
  int KI;               // I rounded up to a multiple of 16
  int x_time_outer;     // tensor-map dimension order: 0 = {I, T, B}, 1 = {I, B, T}
  int o_time_outer;     // the same for the output map (TMA stores)
};

struct TcSmemLayout {
  int h_op, x_op, raw, stage, bars, misc, total;
  int x_tile_bytes, raw_stage_bytes;
};
// Everything below depends on the sub-tile width NS = UMMA N (32 rows for large batches; 16 rows when the batch
// would not fill the SMs otherwise: the per-step chain of a 16-row sub-tile is ~30 % shorter).
// TC_NT = sub-tiles per CTA: 2 (one set of three MMA warps alternating between them) or 4 (two sets of two MMA
// warps, each set alternating between its own pair of sub-tiles; 16-row sub-tiles only: 4 x 4 x 16 accumulator columns).
// ALT (two 32-row sub-tiles only): all sixteen epilogue warps serve BOTH sub-tiles in turn (thread = hidden unit x 8 rows of
// each sub-tile) instead of eight warps per sub-tile.  The per-step chain of a sub-tile is MMA burst + epilogue; with ALT a
// sub-tile's epilogue has four warps per scheduler behind it instead of two and takes half the time, and the two sub-tiles
// alternate strictly (the epilogue warps work on one while the tensor core works on the other).  Same arithmetic per
// element, same accumulators: results are bit-identical to the non-ALT kernel.
// ACC2 (two sub-tiles only): TWO accumulators per sub-tile instead of four.  X takes every lo product FIRST and then the
// hi.hi products of x.W and of h.U k-steps 0..2; Y takes the hi.hi products of h.U k-steps 3..7.  The tensor core truncates
// each addend of an accumulate at 2^-25 of the largest one, so lo products that enter an accumulator while it is still
// small lose nothing: the emulator (tools/emulate_tc_schemes.py lofirst2) puts this order at 0.55-0.63 of the tolerance, the
// same as the four-accumulator scheme (0.54-0.62) -- and the epilogue reads half as much tensor memory per element
// (tcgen05.ld of four accumulators was the largest single cost of the epilogue).
// VR_ (0 = all): VALID rows per epilogue thread.  The CTA's row groups (one per converter warp = one per (sub-tile, row part)
// of the epilogue) each hold VR rows in the first VR of their RPT accumulator columns; the other columns are padding (zero x,
// zero h: batch rows are independent columns of the MMA, so the padding touches nothing).  An MMA costs the same for any
// N <= 32, so 8192 rows can run as 147 CTAs of 4 x 14 rows (all SMs busy) instead of 128 CTAs of 4 x 16 -- with the very same
// arithmetic per row.  Measured: no faster (launch_tc_fwd), opt-in.
template <int TC_NS, int TC_NT, bool ALT = false, bool ACC2 = false, int VR_ = 0>
struct TcFwd {
static_assert(!ACC2 || TC_NT == 2, "ACC2: two sub-tiles");
static_assert(TC_NT == 2 || (TC_NT == 4 && TC_NS == 16), "sub-tile configuration");
static_assert(!ALT || (TC_NT == 2 && TC_NS == 32), "ALT: two 32-row sub-tiles");
static constexpr int TC_MMA_ROLES = ACC2 ? 2 : (TC_NT == 2 ? 3 : TC_NT4_ROLES);       // warps sharing the 30 MMAs of a sub-tile step
static constexpr int TC_SPS = TC_NT == 2 ? 2 : TC_NT4_SPS;     // sub-tiles served by one set of MMA warps
static constexpr int TC_MMA_SETS = TC_NT / TC_SPS;
static constexpr int TC_MMA_WARPS = TC_MMA_ROLES * TC_MMA_SETS;
static constexpr int TC_THREADS = 32 * (TC_MMA_WARPS + TC_CONV_WARPS + TC_EPI_WARPS);   // 736 or 768
static constexpr int TC_ROWS = TC_NS * TC_NT;          // batch rows per CTA (64 or 32)
static constexpr int TC_RH = ALT ? TC_EPI_WARPS / 4 : TC_EPI_WARPS / TC_NT / 4;   // row parts per sub-tile (epilogue warps per lane quadrant)
static constexpr int TC_CONV_ROWS = TC_ROWS / TC_CONV_WARPS;   // rows per converter warp / TMA box
static constexpr int TM_ACC_PER_TILE = 4 * TC_NS;      // CA | CB | M1 | M2
static constexpr int RPT = TC_NS / TC_RH;              // rows per epilogue thread
static constexpr int PAIRS = RPT / 2;                  // row pairs on the packed fp32x2 pipe
static constexpr int NG = RPT / 8;                     // 8-row groups = 16-byte operand chunks per thread
static constexpr int VR = VR_ ? VR_ : RPT;             // valid rows per epilogue thread
static constexpr int VPAIRS = VR / 2;
static constexpr int CONV_VROWS = TC_CONV_ROWS / RPT * VR;     // valid rows per converter warp / TMA box
static constexpr int TC_VROWS = TC_ROWS / RPT * VR;            // batch rows per CTA
static_assert(VR % 2 == 0 && VR <= RPT && (VR == RPT || (!ALT && TC_CONV_ROWS == RPT)), "valid rows per thread");
// TMA stores: every epilogue warp owns staging tiles [h | z | c][RPT rows][32 units] fp32 (lane = unit: 128 contiguous bytes per
// row, conflict-free, immediate offsets) and one lane ships each as a {32, RPT} box; rows past the batch end are clipped by the
// TMA unit.  A scalar store costs four issue slots (IMAD.WIDE + 2 MOV + STG) in a kernel that is bound by them.
static constexpr bool TMA_ST = TC_TMA_STORE != 0 && !ALT;
static constexpr int STAGE_TILE = RPT * 32 * 4;

// ---- operand layouts (SWIZZLE_NONE canonical layouts, 128-byte core matrices) -------------------
// x tile, K-major [rows][KI]: core matrix = 8 rows x 16 B (8 k);  next 8 k: +128 B (LBO);  next 8 rows: +(KI/8)*128 B (SBO)
static __device__ __forceinline__ uint64_t make_desc_kmajor(uint32_t smem_addr, int KI) {
  const uint64_t lbo = 128 >> 4, sbo = (uint64_t)((KI >> 3) * 128) >> 4;
  return (uint64_t)((smem_addr >> 4) & 0x3fff) | (lbo << 16) | (sbo << 32) | (1ull << 46);
}
// h tile, MN-major [k][rows]: core matrix = 8 k x 16 B (8 rows);  next 8 rows: +128 B (SBO);  next 8 k: +(NS/8)*128 B (LBO)
static __device__ __forceinline__ uint64_t make_desc_mnmajor(uint32_t smem_addr) {
  const uint64_t sbo = 128 >> 4, lbo = (uint64_t)((TC_NS >> 3) * 128) >> 4;
  return (uint64_t)((smem_addr >> 4) & 0x3fff) | (lbo << 16) | (sbo << 32) | (1ull << 46);
}
static constexpr uint32_t TC_H_KSTEP = (2 * (TC_NS >> 3) * 128) >> 4;     // descriptor advance per 16 k of the h tile
static constexpr uint32_t TC_X_KSTEP = 256 >> 4;                          // ... of the x tile
// kind::f16: D fp32, A/B fp16, A K-major (TMEM), N = 32, M = 128; bit 16 = B is MN-major
static constexpr uint32_t TC_IDESC_X = (1u << 4) | ((uint32_t)(TC_NS >> 3) << 17) | ((uint32_t)(TC_H >> 4) << 24);
static constexpr uint32_t TC_IDESC_H = TC_IDESC_X | (1u << 16);


static __host__ __device__ inline TcSmemLayout tc_smem_layout(int I, int KI, int esz, bool save) {
  TcSmemLayout L;
  L.x_tile_bytes = TC_NS * KI * 2;
  L.raw_stage_bytes = (CONV_VROWS * I * esz + 127) & ~127;      // per converter warp; a TMA destination is 128-byte aligned
  L.h_op = 0;                                                   // [NT][hi|lo][NS*128*2]
  L.x_op = L.h_op + TC_NT * 2 * TC_NS * TC_H * 2;               // [XBUF][NT][hi|lo][x_tile_bytes]
  L.raw = L.x_op + TC_XBUF * TC_NT * 2 * L.x_tile_bytes;        // [CONV_WARPS][RAW_STAGES][raw_stage_bytes], 128-byte aligned
  L.raw = (L.raw + 127) & ~127;
  L.stage = (L.raw + TC_CONV_WARPS * TC_RAW_STAGES * L.raw_stage_bytes + 127) & ~127;      // [epilogue warp][h | z | c][STAGE_TILE]
  L.bars = L.stage + (TMA_ST ? TC_EPI_WARPS * (save ? 3 : 1) * STAGE_TILE : 0);
  L.misc = L.bars + 32 * 8;
  L.total = L.misc + 256;
  return L;
}

// Gate update of two rows at once on the packed fp32x2 pipe (FADD2 / FMUL2 / FFMA2), rnn.py:289-295.
//   tot = 2^S * pre.   e_g = exp(-(pre + b_g)),  e_u = exp(-2 (pre + b_u));   one MUFU.RCP serves both gates:
//   r = 1/((1+e_g)(1+e_u));  z = r (1+e_u) = sigmoid(pre + b_g);  c = 2 r (1+e_g) - 1 = tanh(pre + b_u);
//   h' = z (h - sz c) + (sz + sn) c   ( = z h + (sz (1 - z) + sn) c ).
// ONE_EX2: e_u = e_g^2 * exp(2 (b_g - b_u)) (one MUFU.EX2 per element; chosen per CTA when no unit's biases are more
// than 8 apart); tot is clamped from below (per-unit constant) so that e_g <= 2^30: z < 1e-9 and c = -1 there.
struct EpiConst { float2 kS, k2S, cg, cu, cu2, msz, szn; float tmin; };
#ifndef FGRNN_TC_C_FORM
#define FGRNN_TC_C_FORM 0       // 1: c = (1 - e_u) * (r a) instead of 2 r a - 1
#endif
#ifndef FGRNN_TC_ONE_EX2
#define FGRNN_TC_ONE_EX2 1      // 1: e_u = e_g^2 * exp(2 (b_g - b_u)) saves one MUFU.EX2 per element
#endif
template <bool ONE_EX2>
static __device__ __forceinline__ float2 gate_update2(float2 tot, float2 h, const EpiConst& k, float2& z, float2& c) {
  float2 ag, eg, eu;
  if (ONE_EX2) {
    // one clamp on tot keeps e_g <= 2^30; e_u = e_g^2 * ratio follows the clamped value consistently (tanh is -1 there)
    tot.x = fmax_nan(tot.x, k.tmin); tot.y = fmax_nan(tot.y, k.tmin);
    ag = __ffma2_rn(tot, k.kS, k.cg);                          // -(pre + b_g) * log2(e)
  } else {
    ag = __ffma2_rn(tot, k.kS, k.cg);
    ag.x = fmin_nan(ag.x, 60.0f); ag.y = fmin_nan(ag.y, 60.0f);      // the two exponents are clamped separately
  }
#ifdef TC_EXP_NO_MUFU
  eg = __fmul2_rn(ag, ag);
#else
  eg.x = ex2_approx(ag.x); eg.y = ex2_approx(ag.y);
#endif
  if (ONE_EX2) {
    eu = __fmul2_rn(__fmul2_rn(eg, eg), k.cu);                 // cu = exp(2 (b_g - b_u))
  } else {
    float2 au = __ffma2_rn(tot, k.k2S, k.cu2);                 // -2 (pre + b_u) * log2(e)
    au.x = fmin_nan(au.x, 60.0f); au.y = fmin_nan(au.y, 60.0f);
    eu.x = ex2_approx(au.x); eu.y = ex2_approx(au.y);
  }
  const float2 one = make_float2(1.0f, 1.0f);
  const float2 a = __fadd2_rn(eg, one), b = __fadd2_rn(eu, one);
  const float2 ab = __fmul2_rn(a, b);
  float2 r;
#ifdef TC_EXP_NO_MUFU
  r = __fmul2_rn(ab, ab);
#else
  r.x = rcp_approx(ab.x); r.y = rcp_approx(ab.y);
#endif
  z = __fmul2_rn(r, b);                                        // rnn.py:290
#if FGRNN_TC_C_FORM
  c = __fmul2_rn(__fadd2_rn(make_float2(-eu.x, -eu.y), one), __fmul2_rn(r, a));            // (1 - e_u) / (1 + e_u)
#else
  c = __ffma2_rn(__fmul2_rn(r, a), make_float2(2.0f, 2.0f), make_float2(-1.0f, -1.0f));   // rnn.py:292
#endif
  return __ffma2_rn(z, __ffma2_rn(k.msz, c, h), __fmul2_rn(k.szn, c));                      // rnn.py:294-295
}

// h (two rows) -> fp16 hi pair and fp16 lo pair (residual, exact subtraction); |h| < 65504
static __device__ __forceinline__ void split_pair(float2 h, uint32_t& hi, uint32_t& lo) {
#ifdef TC_EXP_NO_SPLIT
  hi = __float_as_uint(h.x); lo = __float_as_uint(h.y); return;
#endif
  const __half2 hh = __float22half2_rn(h);
  const float2 hf = __half22float2(hh);
  const __half2 hl = __float22half2_rn(__fadd2_rn(h, make_float2(-hf.x, -hf.y)));
  hi = *reinterpret_cast<const uint32_t*>(&hh);
  lo = *reinterpret_cast<const uint32_t*>(&hl);
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
  int rows_left;                      // B - first row of this thread: rows >= rows_left are padding
  int T, s;
  bool trace;
  // TMA stores
  const CUtensorMap *omap, *zmap, *cmap;
  float* stage;                       // this warp's staging tiles, this lane's column
  int o_unit0, o_row0, o_time_outer;  // box origin: first unit of the warp, first row
};

template <bool HAS_OUT, bool SAVE, bool MASKED, bool ONE_EX2>
static __device__ __forceinline__ void epilogue_loop(const EpiCtx& cx, const EpiConst& kc, float2 (&hst)[PAIRS]) {
  char* outp = reinterpret_cast<char*>(cx.out);
  float* zp = cx.zs; float* cp = cx.cs;
  const uint32_t row_bytes = cx.out_row * 4u;
  for (int t = 0; t < cx.T; ++t) {
    if (cx.trace) TC_TRACE(t, cx.s, 0);
#ifndef TC_EXP_NO_DWAIT
    TC_CRIT_WAIT(cx.bar_dfull, t & 1);
#endif
    tc_fence_after();
    if (cx.trace) TC_TRACE(t, cx.s, 1);
    if (TMA_ST && (HAS_OUT || SAVE)) {                 // the stores of step t-1 have read the staging tiles (issued a whole MMA burst ago)
      if ((threadIdx.x & 31) == 0) tma_store_wait_read<0>();
      __syncwarp();
    }
    uint32_t hi[PAIRS], lo[PAIRS];
#pragma unroll
    for (int q = VPAIRS; q < PAIRS; ++q) { hi[q] = 0u; lo[q] = 0u; }      // padding columns
#pragma unroll
    for (int g = 0; g < NG; ++g) {
      float va[8], vb[8], v1[8], v2[8];
      tmem_ld8(cx.acc + g * 8, va);
      tmem_ld8(cx.acc + TC_NS + g * 8, vb);
      if (!ACC2) {
        tmem_ld8(cx.acc + 2 * TC_NS + g * 8, v1);
        tmem_ld8(cx.acc + 3 * TC_NS + g * 8, v2);
      }
      tmem_ld_wait();
      if (cx.trace && g == 0) TC_TRACE(t, cx.s, 2);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (g * 4 + q >= VPAIRS) continue;
        const float2 corr = __fadd2_rn(make_float2(va[2 * q], va[2 * q + 1]), make_float2(vb[2 * q], vb[2 * q + 1]));     // ACC2: X + Y
        const float2 tot = ACC2 ? corr : __fadd2_rn(__fadd2_rn(corr, make_float2(v2[2 * q], v2[2 * q + 1])), make_float2(v1[2 * q], v1[2 * q + 1]));
        float2 z, c;
#ifdef TC_EXP_NO_MATH
        hst[g * 4 + q] = tot; z = tot; c = tot;
#else
        hst[g * 4 + q] = gate_update2<ONE_EX2>(tot, hst[g * 4 + q], kc, z, c);
#endif
        split_pair(hst[g * 4 + q], hi[g * 4 + q], lo[g * 4 + q]);
        if (SAVE && TMA_ST) {                            // training forward: z_s, c_s (cu:340-341) through the staging tiles
          const int rj = g * 8 + 2 * q;
          cx.stage[STAGE_TILE / 4 + rj * 32] = z.x; cx.stage[STAGE_TILE / 4 + (rj + 1) * 32] = z.y;
          cx.stage[2 * (STAGE_TILE / 4) + rj * 32] = c.x; cx.stage[2 * (STAGE_TILE / 4) + (rj + 1) * 32] = c.y;
        } else if (SAVE) {
          const int rj = g * 8 + 2 * q;
          if (!MASKED || rj < cx.rows_left) { zp[rj * TC_H] = z.x; cp[rj * TC_H] = c.x; }
          if (!MASKED || rj + 1 < cx.rows_left) { zp[(rj + 1) * TC_H] = z.y; cp[(rj + 1) * TC_H] = c.y; }
        }
      }
    }
#pragma unroll
    for (int g = 0; g < NG; ++g) {                     // one 16-byte chunk (8 rows of this unit) per group, hi and lo tile
      *reinterpret_cast<uint4*>(cx.hop + g * 128) = make_uint4(hi[4 * g], hi[4 * g + 1], hi[4 * g + 2], hi[4 * g + 3]);
      *reinterpret_cast<uint4*>(cx.hop + TC_NS * TC_H * 2 + g * 128) = make_uint4(lo[4 * g], lo[4 * g + 1], lo[4 * g + 2], lo[4 * g + 3]);
    }
    if (HAS_OUT && TMA_ST && TC_TMA_STORE == 1) {      // staging before the hand-off: one fence serves both
#pragma unroll
      for (int q = 0; q < VPAIRS; ++q) { cx.stage[(2 * q) * 32] = hst[q].x; cx.stage[(2 * q + 1) * 32] = hst[q].y; }
    }
    if (cx.trace) TC_TRACE(t, cx.s, 3);
    // hand h_t to the tensor core first: fence.proxy.async is MEMBAR.ALL.CTA + FENCE.VIEW.ASYNC and would
    // wait for the global stores too, so those are issued after the arrive and drain behind the next wait
    fence_proxy_async_smem();                          // st.shared of the h tile (and the staging tiles) -> visible to the async proxy
    tc_fence_before();                                 // tcgen05.ld of D done before the next MMAs overwrite it
    __syncwarp();
    if ((threadIdx.x & 31) == 0) mbar_arrive(cx.bar_hready);
    if (cx.trace) TC_TRACE(t, cx.s, 4);
    if (TMA_ST && (HAS_OUT || SAVE)) {
      if (HAS_OUT && TC_TMA_STORE == 2) {              // staging after the hand-off: off the step-critical chain, a fence of its own
#pragma unroll
        for (int q = 0; q < VPAIRS; ++q) { cx.stage[(2 * q) * 32] = hst[q].x; cx.stage[(2 * q + 1) * 32] = hst[q].y; }
        fence_proxy_async_smem();
        __syncwarp();
      }
      if ((threadIdx.x & 31) == 0) {
        const uint32_t src = smem_u32(cx.stage);       // lane 0: the tile base
        if (HAS_OUT) {
          if (cx.o_time_outer) tma_store_3d(cx.omap, cx.o_unit0, cx.o_row0, t, src);
          else tma_store_3d(cx.omap, cx.o_unit0, t, cx.o_row0, src);
        }
        if (SAVE) {
          tma_store_3d(cx.zmap, cx.o_unit0, cx.o_row0, t, src + STAGE_TILE);
          tma_store_3d(cx.cmap, cx.o_unit0, cx.o_row0, t, src + 2 * STAGE_TILE);
        }
        tma_store_commit();
      }
    } else if (HAS_OUT) {
#pragma unroll
      for (int q = 0; q < VPAIRS; ++q) {
        if (!MASKED || 2 * q < cx.rows_left) *reinterpret_cast<float*>(outp + (size_t)(2 * q) * row_bytes) = hst[q].x;
        if (!MASKED || 2 * q + 1 < cx.rows_left) *reinterpret_cast<float*>(outp + (size_t)(2 * q + 1) * row_bytes) = hst[q].y;
      }
      outp += (size_t)cx.out_step * 4u;
    }
    if (cx.trace) TC_TRACE(t, cx.s, 5);
    if (SAVE && !TMA_ST) { zp += cx.zc_step; cp += cx.zc_step; }
  }
  if (TMA_ST && (HAS_OUT || SAVE)) {
    if ((threadIdx.x & 31) == 0) tma_store_wait_all();
    __syncwarp();
  }
}

// ALT: the same step for rows [rh*8, rh*8 + 8) of sub-tile 0, then of sub-tile 1.  cx describes sub-tile 0; sub-tile 1 is at
// fixed offsets (barriers + 8 bytes, accumulators + TM_ACC_PER_TILE, operand tiles + 2 * NS * H * 2, rows + NS).
template <bool HAS_OUT, bool SAVE, bool MASKED, bool ONE_EX2>
static __device__ __forceinline__ void epilogue_loop_alt(const EpiCtx& cx, const EpiConst& kc, float2 (&hst)[TC_NT][PAIRS]) {
  static_assert(NG == 1, "ALT: one 8-row group per sub-tile and thread");
  char* outp = reinterpret_cast<char*>(cx.out);
  float* zp = cx.zs; float* cp = cx.cs;
  const uint32_t row_bytes = cx.out_row * 4u;
  for (int t = 0; t < cx.T; ++t) {
#pragma unroll
    for (int s = 0; s < TC_NT; ++s) {
      TC_CRIT_WAIT(cx.bar_dfull + 8 * s, t & 1);
      tc_fence_after();
      float va[8], vb[8], v1[8], v2[8];
      const uint32_t acc = cx.acc + s * TM_ACC_PER_TILE;
      tmem_ld8(acc, va);
      tmem_ld8(acc + TC_NS, vb);
      if (!ACC2) {
        tmem_ld8(acc + 2 * TC_NS, v1);
        tmem_ld8(acc + 3 * TC_NS, v2);
      }
      tmem_ld_wait();
      uint32_t hi[PAIRS], lo[PAIRS];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float2 corr = __fadd2_rn(make_float2(va[2 * q], va[2 * q + 1]), make_float2(vb[2 * q], vb[2 * q + 1]));     // ACC2: X + Y
        const float2 tot = ACC2 ? corr : __fadd2_rn(__fadd2_rn(corr, make_float2(v2[2 * q], v2[2 * q + 1])), make_float2(v1[2 * q], v1[2 * q + 1]));
        float2 z, c;
        hst[s][q] = gate_update2<ONE_EX2>(tot, hst[s][q], kc, z, c);
        split_pair(hst[s][q], hi[q], lo[q]);
        if (SAVE) {
          const int rj = s * TC_NS + 2 * q;
          if (!MASKED || rj < cx.rows_left) { zp[rj * TC_H] = z.x; cp[rj * TC_H] = c.x; }
          if (!MASKED || rj + 1 < cx.rows_left) { zp[(rj + 1) * TC_H] = z.y; cp[(rj + 1) * TC_H] = c.y; }
        }
      }
      unsigned char* hop = cx.hop + s * (2 * TC_NS * TC_H * 2);
      *reinterpret_cast<uint4*>(hop) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
      *reinterpret_cast<uint4*>(hop + TC_NS * TC_H * 2) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if ((threadIdx.x & 31) == 0) mbar_arrive(cx.bar_hready + 8 * s);
      if (HAS_OUT) {
        char* o = outp + (size_t)(s * TC_NS) * row_bytes;
#pragma unroll
        for (int q = 0; q < PAIRS; ++q) {
          if (!MASKED || s * TC_NS + 2 * q < cx.rows_left) *reinterpret_cast<float*>(o + (size_t)(2 * q) * row_bytes) = hst[s][q].x;
          if (!MASKED || s * TC_NS + 2 * q + 1 < cx.rows_left) *reinterpret_cast<float*>(o + (size_t)(2 * q + 1) * row_bytes) = hst[s][q].y;
        }
      }
    }
    if (HAS_OUT) outp += (size_t)cx.out_step * 4u;
    if (SAVE) { zp += cx.zc_step; cp += cx.zc_step; }
  }
}

// The 30 MMAs of one sub-tile step (K = 128 of h.U, K = 16*NKX of x.W, three fp16 products each), fully
// unrolled, split over three issuing warps so that the serial issue latency is a third:
//   role 0 -> CA : lo terms of x.W (W_lo.x_hi, W_hi.x_lo) and of h.U k-steps 0..2 (U_lo.h_hi, U_hi.h_lo)
//   role 1 -> CB : lo terms of h.U k-steps 3..7
//   role 2 -> M1 : hi.hi of x.W and of h.U k-steps 0..2;   M2 : hi.hi of h.U k-steps 3..7
// With two issuing warps per sub-tile (four sub-tiles per CTA): role 3 = CA + M1, role 4 = CB + M2, 15 MMAs each.
// Every TMEM column and descriptor offset is a compile-time constant on top of uniform bases.
template <int ROLE, int NKX, bool X_HAS_LO>
static __device__ __forceinline__ void issue_subtile_mmas(uint32_t tmem, uint32_t acc, uint64_t dXhi, uint64_t dXlo, uint64_t dHhi, uint64_t dHlo) {
  if (ROLE == 6) {          // ACC2, accumulator X: every lo product first, then hi.hi of x.W and of h.U k-steps 0..2
#pragma unroll
    for (int ks = 0; ks < NKX; ++ks) {
      umma_ts1(acc, tmem + TM_W_LO + ks * 8, dXhi + ks * TC_X_KSTEP, TC_IDESC_X, ks > 0);
      if (X_HAS_LO) umma_ts1(acc, tmem + TM_W_HI + ks * 8, dXlo + ks * TC_X_KSTEP, TC_IDESC_X, 1);
    }
#pragma unroll
    for (int ks = 0; ks < TC_H / 16; ++ks) {
      umma_ts1(acc, tmem + TM_U_LO + ks * 8, dHhi + ks * TC_H_KSTEP, TC_IDESC_H, 1);
      umma_ts1(acc, tmem + TM_U_HI + ks * 8, dHlo + ks * TC_H_KSTEP, TC_IDESC_H, 1);
    }
#pragma unroll
    for (int ks = 0; ks < NKX; ++ks) umma_ts1(acc, tmem + TM_W_HI + ks * 8, dXhi + ks * TC_X_KSTEP, TC_IDESC_X, 1);
#pragma unroll
    for (int ks = 0; ks < 3; ++ks) umma_ts1(acc, tmem + TM_U_HI + ks * 8, dHhi + ks * TC_H_KSTEP, TC_IDESC_H, 1);
    return;
  }
  if (ROLE == 7) {          // ACC2, accumulator Y: hi.hi of h.U k-steps 3..7
#pragma unroll
    for (int ks = 3; ks < TC_H / 16; ++ks) umma_ts1(acc + TC_NS, tmem + TM_U_HI + ks * 8, dHhi + ks * TC_H_KSTEP, TC_IDESC_H, ks > 3);
    return;
  }
  if (ROLE == 0 || ROLE == 3 || ROLE == 5) {
#pragma unroll
    for (int ks = 0; ks < NKX; ++ks) {
      umma_ts1(acc, tmem + TM_W_LO + ks * 8, dXhi + ks * TC_X_KSTEP, TC_IDESC_X, ks > 0);
      if (X_HAS_LO) umma_ts1(acc, tmem + TM_W_HI + ks * 8, dXlo + ks * TC_X_KSTEP, TC_IDESC_X, 1);
    }
#pragma unroll
    for (int ks = 0; ks < 3; ++ks) {
      umma_ts1(acc, tmem + TM_U_LO + ks * 8, dHhi + ks * TC_H_KSTEP, TC_IDESC_H, 1);
      umma_ts1(acc, tmem + TM_U_HI + ks * 8, dHlo + ks * TC_H_KSTEP, TC_IDESC_H, 1);
    }
  }
  if (ROLE == 1 || ROLE == 4 || ROLE == 5) {
#pragma unroll
    for (int ks = 3; ks < TC_H / 16; ++ks) {
      umma_ts1(acc + TC_NS, tmem + TM_U_LO + ks * 8, dHhi + ks * TC_H_KSTEP, TC_IDESC_H, ks > 3);
      umma_ts1(acc + TC_NS, tmem + TM_U_HI + ks * 8, dHlo + ks * TC_H_KSTEP, TC_IDESC_H, 1);
    }
  }
  if (ROLE == 2 || ROLE == 3 || ROLE == 5) {
#pragma unroll
    for (int ks = 0; ks < NKX; ++ks) umma_ts1(acc + 2 * TC_NS, tmem + TM_W_HI + ks * 8, dXhi + ks * TC_X_KSTEP, TC_IDESC_X, ks > 0);
#pragma unroll
    for (int ks = 0; ks < 3; ++ks) umma_ts1(acc + 2 * TC_NS, tmem + TM_U_HI + ks * 8, dHhi + ks * TC_H_KSTEP, TC_IDESC_H, 1);
  }
  if (ROLE == 2 || ROLE == 4 || ROLE == 5) {
#pragma unroll
    for (int ks = 3; ks < TC_H / 16; ++ks) umma_ts1(acc + 3 * TC_NS, tmem + TM_U_HI + ks * 8, dHhi + ks * TC_H_KSTEP, TC_IDESC_H, ks > 3);
  }
}

template <int ROLE>
static __device__ __forceinline__ void issue_subtile_dispatch(int variant, uint32_t tmem, uint32_t acc, uint64_t dXhi, uint64_t dXlo, uint64_t dHhi, uint64_t dHlo) {
  switch (variant) {
    case 0: issue_subtile_mmas<ROLE, 1, false>(tmem, acc, dXhi, dXlo, dHhi, dHlo); break;
    case 1: issue_subtile_mmas<ROLE, 1, true>(tmem, acc, dXhi, dXlo, dHhi, dHlo); break;
    case 2: issue_subtile_mmas<ROLE, 2, false>(tmem, acc, dXhi, dXlo, dHhi, dHlo); break;
    case 3: issue_subtile_mmas<ROLE, 2, true>(tmem, acc, dXhi, dXlo, dHhi, dHlo); break;
    case 4: issue_subtile_mmas<ROLE, 3, false>(tmem, acc, dXhi, dXlo, dHhi, dHlo); break;
    case 5: issue_subtile_mmas<ROLE, 3, true>(tmem, acc, dXhi, dXlo, dHhi, dHlo); break;
    case 6: issue_subtile_mmas<ROLE, 4, false>(tmem, acc, dXhi, dXlo, dHhi, dHlo); break;
    default: issue_subtile_mmas<ROLE, 4, true>(tmem, acc, dXhi, dXlo, dHhi, dHlo); break;
  }
}

static __device__ __forceinline__ void run(const TcArgs& ta, const CUtensorMap& xmap, const CUtensorMap& omap, const CUtensorMap& zmap,
                                           const CUtensorMap& cmap) {
  extern __shared__ __align__(128) unsigned char sm[];
  const SmemFwdArgs& a = ta.f;
  const Dims d = a.d;
  const int I = d.I, KI = ta.KI;
  const int esz = d.x_dtype == FGRNN_BF16 ? 2 : 4;
  const TcSmemLayout L = tc_smem_layout(I, KI, esz, a.save_z != nullptr);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + L.bars);
  uint32_t* tmem_base_s = reinterpret_cast<uint32_t*>(sm + L.misc);
  float* red_s = reinterpret_cast<float*>(sm + L.misc + 16);            // [3][16] max|U|, max|W|, max|b_g - b_u| per epilogue warp

  // warp index through a shuffle: the compiler then knows it is warp uniform and keeps everything derived from it
  // (MMA set / role, descriptors, barrier addresses) in uniform registers instead of R2UR-ing it per instruction
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
  const int row0 = blockIdx.x * TC_VROWS;
  const bool hi_layout = a.layout == FGRNN_LAYOUT_HI;
  // barrier map
  auto bar = [&](int i) { return smem_u32(&bars[i]); };
  const int B_HREADY = 0, B_DFULL = 4, B_XFULL = 8, B_XEMPTY = 12, B_RAWFULL = 16;   // RAWFULL: [conv warp][stage]

  // ---- prologue ---------------------------------------------------------------------------------
  TC_CTA_TIME(0);
  // warp roles: the scheduler favours high warp ids, so the latency-critical single-warp roles sit on top
  constexpr int W_CONV0 = TC_EPI_WARPS, W_MMA = TC_EPI_WARPS + TC_CONV_WARPS;
  if (warp == W_MMA) tmem_alloc(smem_u32(tmem_base_s), TC_TMEM_COLS);
  if (tid == 0) {
    for (int s = 0; s < TC_NT; ++s) { mbar_init(bar(B_HREADY + s), ALT ? TC_EPI_WARPS : TC_EPI_WARPS / TC_NT); mbar_init(bar(B_DFULL + s), TC_MMA_ROLES); }
    for (int b = 0; b < TC_XBUF; ++b) { mbar_init(bar(B_XFULL + b), TC_CONV_WARPS); mbar_init(bar(B_XEMPTY + b), TC_MMA_WARPS); }
    for (int st = 0; st < TC_CONV_WARPS * TC_RAW_STAGES; ++st) mbar_init(bar(B_RAWFULL + st), 1);
    fence_mbar_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // The CTA owns all 512 columns, so the allocation starts at TMEM address 0; using the constant keeps every
  // tcgen05 address warp-uniform (no R2UR in the MMA issue path).
  if (*tmem_base_s != 0u) __trap();
  constexpr uint32_t tmem = 0u;
#ifdef FGRNN_TC_FUZZ
  if (tid == 0) {
    tc_progress()[30] = bar(0);
    if (blockIdx.x == 0 && g_tc_stuck == 0) {
      unsigned long long v0, v1, v2, v3;
      asm volatile("ld.shared.b64 %0, [%1];" : "=l"(v0) : "r"(bar(B_HREADY)) : "memory");
      asm volatile("ld.shared.b64 %0, [%1];" : "=l"(v1) : "r"(bar(B_DFULL)) : "memory");
      asm volatile("ld.shared.b64 %0, [%1];" : "=l"(v2) : "r"(bar(B_XFULL)) : "memory");
      asm volatile("ld.shared.b64 %0, [%1];" : "=l"(v3) : "r"(bar(B_RAWFULL)) : "memory");
      printf("tc_fwd<%d,%d> barriers at %u; fresh words: count %d -> %016llx, count %d -> %016llx, count %d -> %016llx, count 1 -> %016llx\n",
             TC_NS, TC_NT, bar(0), TC_EPI_WARPS / TC_NT, v0, TC_MMA_ROLES, v1, TC_CONV_WARPS, v2, v3);
    }
  }
#endif

  if (warp >= W_MMA) {
    // =========================== MMA issuers ======================================================
    // a set of TC_MMA_ROLES warps shares the 30 MMAs of a sub-tile step and alternates between two sub-tiles
    const int set = (warp - W_MMA) / TC_MMA_ROLES, role = (warp - W_MMA) % TC_MMA_ROLES;
    const bool leader = elect_one();                   // the same lane issues every MMA and commit of this warp
    tc_fence_before();
    __syncthreads();                                   // weights in TMEM, h_{-1} / x_0 tiles under way
    tc_fence_after();
    const int nkx = KI >> 4;
    const bool x_has_lo = d.x_dtype != FGRNN_BF16;     // a bf16 value is one exact fp16 (plus an exact zero lo)
    const uint64_t dH0 = make_desc_mnmajor(smem_u32(sm + L.h_op));
    const uint32_t hlo_step = (uint32_t)(TC_NS * TC_H * 2) >> 4, htile_step = 2 * hlo_step;
    const uint64_t dX0 = make_desc_kmajor(smem_u32(sm + L.x_op), KI);
    const uint32_t xlo_step = (uint32_t)L.x_tile_bytes >> 4, xtile_step = 2 * xlo_step, xbuf_step = TC_NT * xtile_step;
    const int variant = (nkx - 1) * 2 + (x_has_lo ? 1 : 0);
    if (set) tc_spin_ns(set * TC_STAGGER2_NS);           // second set: a quarter period behind the first
    for (int t = 0; t < d.T; ++t) {
      const int xb = t % TC_XBUF;
      mbar_wait(bar(B_XFULL + xb), (t / TC_XBUF) & 1); // x_t operand tiles written
#pragma unroll
      for (int ss = 0; ss < TC_SPS; ++ss) {
        const int s = set * TC_SPS + ss;
        const uint64_t dXhi = dX0 + (uint64_t)(xb * xbuf_step + s * xtile_step), dXlo = dXhi + xlo_step;
        const uint64_t dHhi = dH0 + (uint64_t)(s * htile_step), dHlo = dHhi + hlo_step;
        const uint32_t acc = tmem + TM_ACC + s * TM_ACC_PER_TILE;
        if (role == 0 && set == 0) TC_TRACE(t, s, 8);
        TC_CRIT_WAIT(bar(B_HREADY + s), t & 1);        // h_{t-1} operand tile written, D of step t-1 drained
        tc_fence_after();
        if (role == 0 && set == 0) TC_TRACE(t, s, 9);
        tc_mark(20);
        if (leader) {
          tc_mark(21);
          if (ACC2) {
            if (role == 0) issue_subtile_dispatch<6>(variant, tmem, acc, dXhi, dXlo, dHhi, dHlo);
            else issue_subtile_dispatch<7>(variant, tmem, acc, dXhi, dXlo, dHhi, dHlo);
          } else if (TC_MMA_ROLES == 3) {
            if (role == 0) issue_subtile_dispatch<0>(variant, tmem, acc, dXhi, dXlo, dHhi, dHlo);
            else if (role == 1) issue_subtile_dispatch<1>(variant, tmem, acc, dXhi, dXlo, dHhi, dHlo);
            else issue_subtile_dispatch<2>(variant, tmem, acc, dXhi, dXlo, dHhi, dHlo);
          } else if (TC_MMA_ROLES == 1) {
            issue_subtile_dispatch<5>(variant, tmem, acc, dXhi, dXlo, dHhi, dHlo);
          } else {
            if (role == 0) issue_subtile_dispatch<3>(variant, tmem, acc, dXhi, dXlo, dHhi, dHlo);
            else issue_subtile_dispatch<4>(variant, tmem, acc, dXhi, dXlo, dHhi, dHlo);
          }
          tc_mark(22);
          umma_commit1(bar(B_DFULL + s));              // implies tcgen05.fence::before_thread_sync
          tc_mark(23);
        }
        __syncwarp();
        tc_mark(24);
        if (role == 0 && set == 0) TC_TRACE(t, s, 10);
        // start the two sub-tile pipelines half a period apart so that one is in its MMA phase while the
        // other is in its epilogue (they keep the offset: nothing couples them but shared pipes)
        if (t == 0 && ss == 0 && TC_SPS > 1) tc_spin_ns(TC_STAGGER_NS);
      }
      tc_mark(25);
      if (leader) umma_commit1(bar(B_XEMPTY + xb));    // this warp's MMAs have consumed the x_t tiles
      __syncwarp();
      tc_mark(26);
    }
  } else if (warp >= W_CONV0) {
    // =========================== x path: TMA -> split -> operand tiles ============================
    // Each converter warp owns 16 rows of the CTA's 64: its own TMA box, raw ring and barriers, so the four
    // never wait for one another.  The (row, 8-feature chunk) -> address mapping is step-invariant.
    const int cw = warp - W_CONV0;
    const uint32_t raw_bytes = (uint32_t)(CONV_VROWS * I * esz);           // bytes of one TMA box
    unsigned char* raw_base = sm + L.raw + cw * TC_RAW_STAGES * L.raw_stage_bytes;
    const int my_row0 = row0 + cw * CONV_VROWS;
    auto issue_tma = [&](int t) {
      const int st = t % TC_RAW_STAGES;
      const uint32_t fb = bar(B_RAWFULL + cw * TC_RAW_STAGES + st);
      mbar_expect_tx(fb, raw_bytes);
      if (ta.x_time_outer) tma_load_3d(smem_u32(raw_base + st * L.raw_stage_bytes), &xmap, 0, my_row0, t, fb);
      else tma_load_3d(smem_u32(raw_base + st * L.raw_stage_bytes), &xmap, 0, t, my_row0, fb);
    };
    if (lane == 0)
      for (int t = 0; t < TC_RAW_STAGES && t < d.T; ++t) issue_tma(t);
    tc_fence_before();
    __syncthreads();
    const int nch = KI >> 3, ntask = TC_CONV_ROWS * nch;          // <= 128 tasks: at most 4 per lane
    constexpr int MAXIT = TC_CONV_ROWS * (TC_MAX_KI / 8) / 32;
    uint32_t src_off[MAXIT], dst_off[MAXIT];
    bool live[MAXIT], pad[MAXIT];
#pragma unroll
    for (int it = 0; it < MAXIT; ++it) {
      const int e = it * 32 + lane;
      const int row = e / nch, ch = e - row * nch;
      live[it] = e < ntask;
      pad[it] = ch * 8 >= I || row >= CONV_VROWS;                // K padding chunk, padding row: zeros
      src_off[it] = (uint32_t)(row * I * esz + ch * 8 * esz);
      const int R = cw * TC_CONV_ROWS + row, sidx = R / TC_NS, r = R - sidx * TC_NS;
      dst_off[it] = (uint32_t)((sidx * 2) * L.x_tile_bytes + (r >> 3) * (nch * 128) + ch * 128 + (r & 7) * 16);
    }
    const uint32_t xbuf_bytes = (uint32_t)(TC_NT * 2 * L.x_tile_bytes);
    for (int t = 0; t < d.T; ++t) {
      const int st = t % TC_RAW_STAGES, xb = t % TC_XBUF;
      if (cw == 0) TC_TRACE(t, 0, 12);
      mbar_wait(bar(B_RAWFULL + cw * TC_RAW_STAGES + st), (t / TC_RAW_STAGES) & 1);
      if (cw == 0) TC_TRACE(t, 0, 13);
      mbar_wait(bar(B_XEMPTY + xb), ((t / TC_XBUF) & 1) ^ 1);    // MMAs of step t-XBUF have finished with this buffer
      if (cw == 0) TC_TRACE(t, 0, 14);
      const unsigned char* raw = raw_base + st * L.raw_stage_bytes;
      unsigned char* xdst = sm + L.x_op + xb * xbuf_bytes;
      float v[MAXIT][8];
#pragma unroll
      for (int it = 0; it < MAXIT; ++it) {
        if (live[it] && !pad[it]) {
          if (esz == 4) {
            const float4 p0 = *reinterpret_cast<const float4*>(raw + src_off[it]);
            const float4 p1 = *reinterpret_cast<const float4*>(raw + src_off[it] + 16);
            v[it][0] = p0.x; v[it][1] = p0.y; v[it][2] = p0.z; v[it][3] = p0.w;
            v[it][4] = p1.x; v[it][5] = p1.y; v[it][6] = p1.z; v[it][7] = p1.w;
          } else {
            const uint4 p = *reinterpret_cast<const uint4*>(raw + src_off[it]);
            const uint32_t w[4] = {p.x, p.y, p.z, p.w};
#pragma unroll
            for (int q = 0; q < 4; ++q) { v[it][2 * q] = __uint_as_float(w[q] << 16); v[it][2 * q + 1] = __uint_as_float(w[q] & 0xffff0000u); }
          }
        } else {
#pragma unroll
          for (int q = 0; q < 8; ++q) v[it][q] = 0.f;
        }
      }
#pragma unroll
      for (int it = 0; it < MAXIT; ++it) {
        if (live[it]) {
          uint32_t hi[4], lo[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) split2(v[it][2 * q], v[it][2 * q + 1], 1.0f, hi[q], lo[q]);
          *reinterpret_cast<uint4*>(xdst + dst_off[it]) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
          *reinterpret_cast<uint4*>(xdst + dst_off[it] + L.x_tile_bytes) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
      }
      fence_proxy_async_smem();                        // st.shared of the operand tiles -> visible to tcgen05.mma;
      __syncwarp();                                    // also orders this warp's raw reads before the next TMA write
      if (lane == 0) {
        mbar_arrive(bar(B_XFULL + xb));
        if (t + TC_RAW_STAGES < d.T) issue_tma(t + TC_RAW_STAGES);
      }
      if (cw == 0) TC_TRACE(t, 0, 15);
      __syncwarp();
    }
  } else {
    // =========================== epilogue warps ===================================================
    const int ew = warp;                               // 0..15
    const int quad = warp & 3;                         // TMEM lane quadrant this warp may access (= ew & 3)
    const int es = ALT ? 0 : (ew >> 2) % TC_NT;        // the sub-tile this warp serves (ALT: both, described from sub-tile 0)
    const int rh = ALT ? (ew >> 2) : (ew >> 2) / TC_NT; // which part of the sub-tile's rows
    const int n = quad * 32 + lane;                    // hidden unit = TMEM lane
    const uint32_t lane_base = (uint32_t)(quad * 32) << 16;

    // weights -> tensor memory (A operands): this thread's row of U^T / W^T, fp16 hi/lo pairs along k.
    // The four warps of a quadrant take 32 k each of U and 16 k each of W; everything is loaded in one batch,
    // the power-of-two scale comes from max|U|, max|W| over the CTA's copy (identical in every CTA).
    float scale_w, unscale;
    bool wide_bias;
    {
      const int part = ew >> 2;                        // 0..3
      float uv[32], wv[16];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int k = part * 32 + j;
        uv[j] = hi_layout ? __ldg(a.U + (size_t)n * TC_H + k) : __ldg(a.U + (size_t)k * TC_H + n);
      }
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int k = part * 16 + j;
        wv[j] = k < I ? (hi_layout ? __ldg(a.W + (size_t)n * I + k) : __ldg(a.W + (size_t)k * TC_H + n)) : 0.f;
      }
      float mu = 0.f, mw = 0.f;
#pragma unroll
      for (int j = 0; j < 32; ++j) mu = fmaxf(mu, fabsf(uv[j]));
#pragma unroll
      for (int j = 0; j < 16; ++j) mw = fmaxf(mw, fabsf(wv[j]));
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        mu = fmaxf(mu, __shfl_xor_sync(0xffffffffu, mu, o));
        mw = fmaxf(mw, __shfl_xor_sync(0xffffffffu, mw, o));
      }
      // a third reduced quantity: the largest gate/update bias distance of any unit (selects the activation form)
      float bd = part == 0 ? fabsf(__ldg(a.bias_gate + n) - __ldg(a.bias_update + n)) : 0.f;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) bd = fmaxf(bd, __shfl_xor_sync(0xffffffffu, bd, o));
      if (lane == 0) { red_s[ew] = mu; red_s[16 + ew] = mw; red_s[32 + ew] = bd; }
      asm volatile("bar.sync 1, %0;" ::"n"(TC_EPI_WARPS * 32) : "memory");      // epilogue warps only
#pragma unroll
      for (int w = 0; w < TC_EPI_WARPS; ++w) { mu = fmaxf(mu, red_s[w]); mw = fmaxf(mw, red_s[16 + w]); bd = fmaxf(bd, red_s[32 + w]); }
      wide_bias = !(bd <= 8.0f);                        // NaN biases also take the plain form
      // accumulators hold 2^S * pre:  h.(U*2^S) and x.(W*2^S); h and x themselves are split unscaled (an fp16
      // subnormal lo part still resolves 2^-24 absolute, i.e. fp32-level for |h| <= 1)
      int S = 40;
      if (mw > 0.f) S = min(S, (int)floorf(log2f(30000.f / mw)));
      if (mu > 0.f) S = min(S, (int)floorf(log2f(30000.f / mu)));
      S = max(S, -14);
      scale_w = exp2f((float)S);
      unscale = exp2f((float)-S);
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) split2(uv[c * 16 + 2 * j], uv[c * 16 + 2 * j + 1], scale_w, hi[j], lo[j]);
        tmem_st8(tmem + lane_base + TM_U_HI + part * 16 + c * 8, hi);
        tmem_st8(tmem + lane_base + TM_U_LO + part * 16 + c * 8, lo);
      }
      if (part * 16 < KI) {
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) split2(wv[2 * j], wv[2 * j + 1], scale_w, hi[j], lo[j]);
        tmem_st8(tmem + lane_base + TM_W_HI + part * 8, hi);
        tmem_st8(tmem + lane_base + TM_W_LO + part * 8, lo);
      }
      tmem_st_wait();
    }

    EpiConst kc;
    bool one_ex2;
    {
      const float sz = sigmoid_f(__ldg(a.zeta)), sn = sigmoid_f(__ldg(a.nu));
      constexpr float LOG2E = 1.4426950408889634f;
      const float bgv = __ldg(a.bias_gate + n), buv = __ldg(a.bias_update + n);
      const float kS = -LOG2E * unscale, cg = -LOG2E * bgv;
      kc.kS = make_float2(kS, kS); kc.cg = make_float2(cg, cg);
      kc.msz = make_float2(-sz, -sz); kc.szn = make_float2(sz + sn, sz + sn);
      const float k2S = -2.0f * LOG2E * unscale, cu2 = -2.0f * LOG2E * buv;
      kc.k2S = make_float2(k2S, k2S); kc.cu2 = make_float2(cu2, cu2);
      // e_u = e_g^2 * exp(2 (b_g - b_u)) saves one MUFU.EX2 per element; if any unit's two biases are more than 8
      // apart (exp(16) ~ 9e6 still leaves head-room) the whole CTA takes the two-EX2 form instead
      one_ex2 = FGRNN_TC_ONE_EX2 && !wide_bias;
      const float ratio = one_ex2 ? expf(2.0f * (bgv - buv)) : 1.0f;
      kc.cu = make_float2(ratio, ratio);
      // one-EX2 form: tot >= tmin  <=>  e_g <= 2^30, so (1+e_g)(1+e_g^2 ratio) stays finite (ratio <= e^16);
      // the two-EX2 form clamps each exponent at 60 instead
      kc.tmin = (30.0f - cg) / kS;
    }

    if constexpr (ALT) {
      // state: h[row][n] for rows [rh*8, rh*8 + 8) of BOTH sub-tiles
      float2 hst[TC_NT][PAIRS];
      unsigned char* hop = sm + L.h_op + (n >> 3) * ((TC_NS >> 3) * 128) + (rh * NG) * 128 + (n & 7) * 16;
      const int first_row = row0 + rh * RPT;           // of sub-tile 0; sub-tile 1: + TC_NS
#pragma unroll
      for (int s = 0; s < TC_NT; ++s) {
        uint32_t hi[PAIRS], lo[PAIRS];
#pragma unroll
        for (int q = 0; q < PAIRS; ++q) {
          const int row = first_row + s * TC_NS + 2 * q;
          const float v0 = (a.h0 && row < d.B) ? __ldg(a.h0 + (size_t)row * TC_H + n) : 0.f;
          const float v1 = (a.h0 && row + 1 < d.B) ? __ldg(a.h0 + (size_t)(row + 1) * TC_H + n) : 0.f;
          hst[s][q] = make_float2(v0, v1);
          split2(v0, v1, 1.0f, hi[q], lo[q]);
        }
        *reinterpret_cast<uint4*>(hop + s * (2 * TC_NS * TC_H * 2)) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<uint4*>(hop + s * (2 * TC_NS * TC_H * 2) + TC_NS * TC_H * 2) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
      }
      fence_proxy_async_smem();
      tc_fence_before();
      __syncthreads();                                 // matches the other roles' prologue barrier
      __syncwarp();
      if (lane == 0) { mbar_arrive(bar(B_HREADY)); mbar_arrive(bar(B_HREADY + 1)); }    // phase 0: h_{-1} ready
      EpiCtx cx;
      cx.bar_dfull = bar(B_DFULL); cx.bar_hready = bar(B_HREADY);
      cx.acc = tmem + lane_base + TM_ACC + rh * RPT;
      cx.hop = hop;
      cx.out = a.out ? a.out + (size_t)first_row * a.osb + n : nullptr;
      cx.zs = a.save_z ? a.save_z + (size_t)first_row * TC_H + n : nullptr;
      cx.cs = a.save_c ? a.save_c + (size_t)first_row * TC_H + n : nullptr;
      cx.out_row = (uint32_t)a.osb; cx.out_step = (uint32_t)a.ost; cx.zc_step = (uint32_t)d.B * TC_H;
      cx.rows_left = d.B - first_row; cx.T = d.T; cx.trace = false; cx.s = 0;
      const bool masked = row0 + TC_ROWS > d.B;
      const int variant = (one_ex2 ? 8 : 0) | (a.out ? 4 : 0) | (a.save_z ? 2 : 0) | (masked ? 1 : 0);
      switch (variant) {
        case 0: epilogue_loop_alt<false, false, false, false>(cx, kc, hst); break;
        case 1: epilogue_loop_alt<false, false, true, false>(cx, kc, hst); break;
        case 2: epilogue_loop_alt<false, true, false, false>(cx, kc, hst); break;
        case 3: epilogue_loop_alt<false, true, true, false>(cx, kc, hst); break;
        case 4: epilogue_loop_alt<true, false, false, false>(cx, kc, hst); break;
        case 5: epilogue_loop_alt<true, false, true, false>(cx, kc, hst); break;
        case 6: epilogue_loop_alt<true, true, false, false>(cx, kc, hst); break;
        case 7: epilogue_loop_alt<true, true, true, false>(cx, kc, hst); break;
        case 8: epilogue_loop_alt<false, false, false, true>(cx, kc, hst); break;
        case 9: epilogue_loop_alt<false, false, true, true>(cx, kc, hst); break;
        case 10: epilogue_loop_alt<false, true, false, true>(cx, kc, hst); break;
        case 11: epilogue_loop_alt<false, true, true, true>(cx, kc, hst); break;
        case 12: epilogue_loop_alt<true, false, false, true>(cx, kc, hst); break;
        case 13: epilogue_loop_alt<true, false, true, true>(cx, kc, hst); break;
        case 14: epilogue_loop_alt<true, true, false, true>(cx, kc, hst); break;
        default: epilogue_loop_alt<true, true, true, true>(cx, kc, hst); break;
      }
      if (a.h_last) {
#pragma unroll
        for (int s = 0; s < TC_NT; ++s) {
#pragma unroll
          for (int j = 0; j < RPT; ++j) {
            const int row = first_row + s * TC_NS + j;
            if (row < d.B) a.h_last[(size_t)row * TC_H + n] = (j & 1) ? hst[s][j >> 1].y : hst[s][j >> 1].x;
          }
        }
      }
    } else {
    // state: this thread owns h[row][n] for 16 rows of its sub-tile, kept as row pairs for the fp32x2 pipe
    float2 hst[PAIRS];
    unsigned char* hop = sm + L.h_op + es * (2 * TC_NS * TC_H * 2) + (n >> 3) * ((TC_NS >> 3) * 128) + (rh * NG) * 128 + (n & 7) * 16;
    const int first_row = row0 + (es * TC_RH + rh) * VR;
    {
      uint32_t hi[PAIRS], lo[PAIRS];
#pragma unroll
      for (int q = 0; q < PAIRS; ++q) {
        const int row = first_row + 2 * q;
        const float v0 = (a.h0 && q < VPAIRS && row < d.B) ? __ldg(a.h0 + (size_t)row * TC_H + n) : 0.f;
        const float v1 = (a.h0 && q < VPAIRS && row + 1 < d.B) ? __ldg(a.h0 + (size_t)(row + 1) * TC_H + n) : 0.f;
        hst[q] = make_float2(v0, v1);
        split2(v0, v1, 1.0f, hi[q], lo[q]);
      }
#pragma unroll
      for (int g = 0; g < NG; ++g) {
        *reinterpret_cast<uint4*>(hop + g * 128) = make_uint4(hi[4 * g], hi[4 * g + 1], hi[4 * g + 2], hi[4 * g + 3]);
        *reinterpret_cast<uint4*>(hop + TC_NS * TC_H * 2 + g * 128) = make_uint4(lo[4 * g], lo[4 * g + 1], lo[4 * g + 2], lo[4 * g + 3]);
      }
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();                                   // matches the other roles' prologue barrier
    __syncwarp();
    if (lane == 0) mbar_arrive(bar(B_HREADY + es));    // phase 0: h_{-1} ready

    TC_CTA_TIME(1);
    EpiCtx cx;
    cx.bar_dfull = bar(B_DFULL + es); cx.bar_hready = bar(B_HREADY + es);
    cx.acc = tmem + lane_base + TM_ACC + es * TM_ACC_PER_TILE + rh * RPT;
    cx.hop = hop;
    cx.out = a.out ? a.out + (size_t)first_row * a.osb + n : nullptr;
    cx.zs = a.save_z ? a.save_z + (size_t)first_row * TC_H + n : nullptr;
    cx.cs = a.save_c ? a.save_c + (size_t)first_row * TC_H + n : nullptr;
    cx.out_row = (uint32_t)a.osb; cx.out_step = (uint32_t)a.ost; cx.zc_step = (uint32_t)d.B * TC_H;
    cx.rows_left = d.B - first_row; cx.T = d.T; cx.trace = (ew & 11) == 0; cx.s = es;
    cx.omap = &omap; cx.zmap = &zmap; cx.cmap = &cmap;
    cx.stage = reinterpret_cast<float*>(sm + L.stage) + ew * (a.save_z ? 3 : 1) * (STAGE_TILE / 4) + lane;
    cx.o_unit0 = quad * 32; cx.o_row0 = first_row; cx.o_time_outer = ta.o_time_outer;
    const bool masked = !TMA_ST && row0 + TC_VROWS > d.B;       // the TMA unit clips the boxes at the batch end
    const int variant = (one_ex2 ? 8 : 0) | (a.out ? 4 : 0) | (a.save_z ? 2 : 0) | (masked ? 1 : 0);
    switch (variant) {
      case 0: epilogue_loop<false, false, false, false>(cx, kc, hst); break;
      case 1: if constexpr (!TMA_ST) epilogue_loop<false, false, true, false>(cx, kc, hst); break;
      case 2: epilogue_loop<false, true, false, false>(cx, kc, hst); break;
      case 3: if constexpr (!TMA_ST) epilogue_loop<false, true, true, false>(cx, kc, hst); break;
      case 4: epilogue_loop<true, false, false, false>(cx, kc, hst); break;
      case 5: if constexpr (!TMA_ST) epilogue_loop<true, false, true, false>(cx, kc, hst); break;
      case 6: epilogue_loop<true, true, false, false>(cx, kc, hst); break;
      case 7: if constexpr (!TMA_ST) epilogue_loop<true, true, true, false>(cx, kc, hst); break;
      case 8: epilogue_loop<false, false, false, true>(cx, kc, hst); break;
      case 9: if constexpr (!TMA_ST) epilogue_loop<false, false, true, true>(cx, kc, hst); break;
      case 10: epilogue_loop<false, true, false, true>(cx, kc, hst); break;
      case 11: if constexpr (!TMA_ST) epilogue_loop<false, true, true, true>(cx, kc, hst); break;
      case 12: epilogue_loop<true, false, false, true>(cx, kc, hst); break;
      case 13: if constexpr (!TMA_ST) epilogue_loop<true, false, true, true>(cx, kc, hst); break;
      case 14: epilogue_loop<true, true, false, true>(cx, kc, hst); break;
      default: if constexpr (!TMA_ST) epilogue_loop<true, true, true, true>(cx, kc, hst); break;
    }
    TC_CTA_TIME(2);
    if (a.h_last) {
#pragma unroll
      for (int j = 0; j < VR; ++j) {
        const int row = first_row + j;
        if (row < d.B) a.h_last[(size_t)row * TC_H + n] = (j & 1) ? hst[j >> 1].y : hst[j >> 1].x;
      }
    }
    }   // !ALT
  }
  tc_fence_before();
  __syncthreads();
  if (warp == W_MMA) tmem_dealloc(tmem, TC_TMEM_COLS);
}

};   // struct TcFwd

template <int NS, int NT, bool ALT = false, bool ACC2 = false, int VR = 0>
__global__ void __launch_bounds__((TcFwd<NS, NT, ALT, ACC2, VR>::TC_THREADS), 1)
tc_fwd_kernel(const TcArgs ta, const __grid_constant__ CUtensorMap xmap, const __grid_constant__ CUtensorMap omap,
              const __grid_constant__ CUtensorMap zmap, const __grid_constant__ CUtensorMap cmap) {
  TcFwd<NS, NT, ALT, ACC2, VR>::run(ta, xmap, omap, zmap, cmap);
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
bool tc_path_supports(const Dims& d) {
  // full-rank, H = 128, I a multiple of 8 up to 32, sigmoid gate / tanh update.  The kernel itself handles I <= 64, but
  // with four x k-steps the hi.hi chain M1 grows to 7 MMAs of larger addends and the result lands at 1.05x the state
  // tolerance against the oracle (0.80 against an fp64 evaluation; measured in round 2 at I = 64, T = 99): those shapes go
  // to the hoisted-projection kernels (fgrnn_tc_wx.cu), whose chains stay at 4 MMAs.  FGRNN_TC_WIDE=0 keeps them here.
  const int max_i = tuning(TUNE_TC_WIDE) == 0 ? TC_MAX_KI : 32;
  return d.rW == 0 && d.rU == 0 && d.H == TC_H && d.I >= 8 && d.I <= max_i && (d.I % 8) == 0 &&
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

template <int NS, int NT, bool ALT = false, bool ACC2 = false, int VR = 0>
static int launch_tc_fwd_ns(const SmemFwdArgs& a, cudaStream_t stream) {
  using K = TcFwd<NS, NT, ALT, ACC2, VR>;
  const Dims& d = a.d;
  TcArgs ta{};
  ta.f = a;
  ta.KI = (d.I + 15) & ~15;
  const int esz = d.x_dtype == FGRNN_BF16 ? 2 : 4;
  CUtensorMap map;
  const int rc = make_row_tile_map(&map, a.x, d.x_dtype == FGRNN_BF16, d.I, d.B, d.T, a.xsb, a.xst, K::CONV_VROWS, &ta.x_time_outer);
  if (rc) return rc;
  CUtensorMap omap = map, zmap = map, cmap = map;      // valid descriptors even where a tensor is absent
  if (K::TMA_ST) {
    int to = 1;
    if (a.out) { const int r2 = make_row_tile_map(&omap, a.out, false, TC_H, d.B, d.T, a.osb, a.ost, K::VR, &ta.o_time_outer, 32); if (r2) return r2; }
    if (a.save_z) {                                    // [T][B][H] contiguous
      int r2 = make_row_tile_map(&zmap, a.save_z, false, TC_H, d.B, d.T, TC_H, (int64_t)d.B * TC_H, K::VR, &to, 32);
      if (!r2) r2 = make_row_tile_map(&cmap, a.save_c, false, TC_H, d.B, d.T, TC_H, (int64_t)d.B * TC_H, K::VR, &to, 32);
      if (r2) return r2;
    }
  }
  const TcSmemLayout L = K::tc_smem_layout(d.I, ta.KI, esz, a.save_z != nullptr);
  FGRNN_CUDA_TRY(cudaFuncSetAttribute(tc_fwd_kernel<NS, NT, ALT, ACC2, VR>, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total));
  const unsigned grid = (unsigned)((d.B + K::TC_VROWS - 1) / K::TC_VROWS);
  tc_fwd_kernel<NS, NT, ALT, ACC2, VR><<<grid, K::TC_THREADS, L.total, stream>>>(ta, map, omap, zmap, cmap);
  FGRNN_LAUNCH_CHECK("tc_fwd_kernel");
  return FGRNN_OK;
}

// Tile configuration.  The per-step chain of a CTA does not depend on how many SMs are busy, and a tcgen05.mma costs
// the same ~19 cycles for N = 16 and N = 32, so:
//   batch fits one wave of 32-row CTAs  -> 2 x 16-row sub-tiles (shortest chain: ~2050 cycles per step);
//   larger batches                      -> 2 x 32-row sub-tiles (~3000 cycles per step for 64 rows).
// 4 x 16-row sub-tiles with two MMA warp sets (~3150 cycles) was the faster multi-wave configuration while the MMA
// warps still paid ~5 R2UR per MMA; with uniform descriptors 2 x 32 leads by 1-4 %.
// FGRNN_TC_NS=16|32 and FGRNN_TC_NT=2|4 override (tests, benchmarks); NT=4 implies NS=16.
int launch_tc_fwd(const SmemFwdArgs& a, cudaStream_t stream) {
  if (a.d.B <= 0 || a.d.T <= 0) return FGRNN_OK;
  int ns = a.d.B <= 148 * 32 ? 16 : 32, nt = 2;
  if (tuning(TUNE_TC_NT) == 4) nt = 4; else if (tuning(TUNE_TC_NT) == 2) nt = 2;
  if (tuning(TUNE_TC_NS) == 32) ns = 32; else if (tuning(TUNE_TC_NS) == 16) ns = 16;
  if (nt == 4) ns = 16;
  if (nt == 4) return launch_tc_fwd_ns<16, 4>(a, stream);
  // Measured on C2 / C5 (round 2, same box, A/B): two accumulators per sub-tile with the lo products first (FGRNN_TC_ACC2=1)
  // 0.1527 -> 0.1512 ms / 1.390 -> 1.364 ms at 32-row sub-tiles, 2 % SLOWER at 16-row sub-tiles (training forward, 2048 rows).
  // Opt-in: a 1-2 % gain is not worth giving up that every tile configuration computes the very same bits (a 4099-row
  // shard runs 16-row sub-tiles, the 8192-row batch 32-row ones; tests/test_gpu_fullsize.py compares them bit for bit).
  // Sixteen epilogue warps alternating between the sub-tiles (FGRNN_TC_ALT=1) is bit-identical but 3-5 % slower: opt-in too.
  const bool acc2 = tuning(TUNE_TC_ACC2) == 1, alt = tuning(TUNE_TC_ALT) == 1;
  // FGRNN_TC_VR=14: 56-row CTAs (TcFwd: VR_; 8192 rows then fill 147 SMs instead of 128).  Bit-identical, and measured on C2 / C5
  // (same box, A/B): 0.1504 -> 0.1496 ms / 1.332 -> 1.353 ms -- the per-step chain of a CTA does not get shorter with fewer rows
  // per epilogue thread, so more CTAs of fewer rows buy nothing here (they do in the low-rank kernel, fgrnn_tc_lr.cu).  Opt-in.
  if (ns == 32 && !acc2 && !alt && tuning(TUNE_TC_VR) == 14) return launch_tc_fwd_ns<32, 2, false, false, 14>(a, stream);
  if (ns == 16) return acc2 ? launch_tc_fwd_ns<16, 2, false, true>(a, stream) : launch_tc_fwd_ns<16, 2>(a, stream);
  if (acc2) return alt ? launch_tc_fwd_ns<32, 2, true, true>(a, stream) : launch_tc_fwd_ns<32, 2, false, true>(a, stream);
  return alt ? launch_tc_fwd_ns<32, 2, true>(a, stream) : launch_tc_fwd_ns<32, 2>(a, stream);
}

}  // namespace fgrnn
