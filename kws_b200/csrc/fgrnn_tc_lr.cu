// tcgen05 / TMEM kernel for the LOW-RANK recurrence (FGRNN_PATH_LOWRANK when the shape fits; BASELINE config 4:
// H = 256, wRank 16, uRank 32, I = 32):   pre_t = (x_t.W1).W2 + (h_{t-1}.U1).U2   in the reference's factored order
// (rnn.py:280-287), as two chained tensor-core stages per time step:
//
//   stage 1   D1^T[rank row][row] = A1 . [h_{t-1} ; x_t]^T            M = 128, K = 256 + KI, N = 32 batch rows.  The fp16 hi and
//             lo parts of the weights are STACKED on the M rows (lane quadrant q < 2: lanes 0..15 = U1_hi^T of ranks 16q..16q+15,
//             lanes 16..31 = U1_lo^T of the same ranks; quadrant 2: W1_hi^T | W1_lo^T; A is block diagonal), so one MMA with
//             B = h_hi yields hi.hi in the main rows AND U1_lo.h_hi in the correction rows: two MMAs per k-step, not three
//   hop       three warps (lane quadrants 0..2) read the rank rows (tcgen05.ld), add each main row and its correction row
//             (one shuffle), un-scale, split into fp16 hi/lo and write the MN-major [48][32] B operand of stage 2
//   stage 2   D2^T[unit][row] = [U2^T | W2^T] . [s ; sx]              two M = 128 tiles (256 units), K = 48, N = 32
//   epilogue  gate update (rnn.py:289-295) exactly as fgrnn_tc.cu: thread = hidden unit, 32 rows, h in registers
//
// Warp roles (24 warps): 0..15 epilogue -- ALL of them serve both sub-tiles in turn (thread = hidden unit x 16 rows of
// each sub-tile), so that a sub-tile's epilogue has four warps per scheduler behind it: the per-step chain of a sub-tile
// (stage 1 -> hop -> stage 2 -> epilogue) is what bounds the kernel, and with eight warps per sub-tile the epilogue alone
// was half of it (tools/trace_lowrank.py); 16..18 hop (lane quadrants 0..2, no global stores in flight when they fence; six hop
// warps -- one 16-row half of a sub-tile each, TL_HOP_HALVES=2 -- measured the same); 19..20 x path (TMA -> fp16 split, one
// sub-tile each); 21..24 MMA issue ([sub-tile][role]).  h_t leaves through per-warp staging tiles and TMA tile stores
// (TL_TMA_STORE, 1.146 -> 1.114 ms at C4: a scalar store costs four issue slots); the gate's reciprocal on the FMA pipe
// instead of MUFU.RCP (TL_RCP_NEWTON) was measured 6 % slower and is off.
//
// Both weight sets stay in TENSOR MEMORY for the whole kernel: stage 1 takes 144 columns (hi and lo share them, on different
// lanes), stage 2 96 (fp16 hi | lo).  D1 and D2 of a sub-tile are never live at the same time (D1 dies when
// the hop has read it, D2 when the epilogue has read it), so they ALIAS: two sub-tiles of 32 rows per CTA run half a
// period apart, each with its own pair of MMA-issuing warps.
//
// fp32 parity (tools/emulate_tc_lowrank2.py, bit-exact model of the tensor core's accumulate): every accumulator takes
// its small lo products FIRST and the hi.hi products after them -- the tensor core truncates each addend of an
// accumulate at 2^-25 of the largest, so a chain that interleaves them (1.3-1.9x the tolerance) loses what the lo-first
// order keeps (0.49-0.72x); the hi.hi chain of the 16 h k-steps is split over the two accumulators of D1.
#include <cuda.h>

#include "fgrnn_kernels.cuh"
#include "fgrnn_tc_common.cuh"

namespace fgrnn {

constexpr int TL_H = 256, TL_NS = 32, TL_NT = 2, TL_ROWS = TL_NS * TL_NT;
constexpr int TL_RU = 32, TL_RW = 16, TL_K2 = TL_RU + TL_RW;            // padded ranks; K of stage 2
#ifndef TL_HOP_HALVES
#define TL_HOP_HALVES 1           // 1: three hop warps (one per lane quadrant) convert both 16-row halves of a sub-tile; 2: six, one half each
#endif
#ifndef TL_RCP_NEWTON
#define TL_RCP_NEWTON 0           // 1: the reciprocal of the gate update on the FMA pipe (Newton) instead of MUFU.RCP
#endif
constexpr int TL_EPI_WARPS = 16, TL_HOP_WARPS = 3 * TL_HOP_HALVES, TL_CONV_WARPS = 2, TL_MMA_WARPS = 4;   // MMA warps: [sub-tile][role]
#ifndef TL_TMA_STORE
#define TL_TMA_STORE 1            // 1: h_t leaves through a per-warp staging tile and TMA tile stores; 0: scalar st.global per row
#endif
constexpr int TL_THREADS = 32 * (TL_EPI_WARPS + TL_HOP_WARPS + TL_CONV_WARPS + TL_MMA_WARPS);
constexpr int TL_RPT = TL_NS / 2;                     // rows of each sub-tile per epilogue thread
constexpr int TL_XBUF = 4, TL_RAW_STAGES = 4, TL_CONV_ROWS = TL_ROWS / TL_CONV_WARPS;
constexpr int TL_MAX_KI = 32;
// tensor-memory column map
constexpr uint32_t TLM_A1 = 0, TLM_A1X = 128;         // stage-1 weights (hi | lo stacked on the lanes): h part 128 columns, x part 16
constexpr uint32_t TLM_A2 = 144;                      // + m * 48 + {0: hi, 24: lo}
constexpr uint32_t TLM_ACC = 384;                     // + s * 64:  D1 = X | Y (32 columns each), aliased by D2 = tile 0 | tile 1
constexpr int TL_H_TILE = TL_H * TL_NS * 2;           // one fp16 [256][32] operand tile: 16 KB
constexpr int TL_S_TILE = TL_K2 * TL_NS * 2;          // one fp16 [48][32] operand tile: 3 KB
#ifndef TL_STAGGER_NS
#define TL_STAGGER_NS 900
#endif

// Developer trace (make trace -> -DFGRNN_TL_TRACE): clock64 stamps of CTA 0 for steps [16, 24), per sub-tile:
//   0 MMA role 0: HREADY seen   1 stage 1 issued   2 SREADY seen   3 stage 2 issued
//   4 hop warp 0: D1FULL seen   5 SREADY arrive    6 epilogue warp 0: DFULL seen   7 HREADY arrive   8 stores issued
//   9 epilogue warp 15: DFULL seen    10 its HREADY arrive
#ifdef FGRNN_TL_TRACE
__device__ long long g_tl_trace[8 * 2 * 16];
#define TL_TRACE(t, s, slot)                                                                         \
  do {                                                                                               \
    if (blockIdx.x == 0 && (threadIdx.x & 31) == 0 && (t) >= 16 && (t) < 24)                         \
      g_tl_trace[(((t) - 16) * 2 + (s)) * 16 + (slot)] = clock64();                                  \
  } while (0)
#else
#define TL_TRACE(t, s, slot) do { } while (0)
#endif

struct TlArgs {
  FwdArgs f;
  int KI;               // I rounded up to a multiple of 16 (16 or 32)
  int x_time_outer;
  int o_time_outer;     // coordinate order of the output map (TL_TMA_STORE)
};
struct TlSmem { int h_op, s_op, x_op, raw, stage, bars, misc, total; int x_tile_bytes, raw_stage_bytes; };
constexpr int TL_STAGE_TILE = 16 * 32 * 4;            // one epilogue warp's h_t box: 16 rows x 32 units fp32

__host__ __device__ inline TlSmem tl_smem_layout(int I, int KI, int esz) {
  TlSmem L;
  L.x_tile_bytes = TL_NS * KI * 2;
  L.raw_stage_bytes = (TL_CONV_ROWS * I * esz + 127) & ~127;    // sized for whole sub-tiles; a TMA destination is 128-byte aligned
  L.h_op = 0;                                                   // [NT][hi|lo][TL_H_TILE]
  L.s_op = L.h_op + TL_NT * 2 * TL_H_TILE;                      // [NT][hi|lo][TL_S_TILE]
  L.x_op = L.s_op + TL_NT * 2 * TL_S_TILE;                      // [XBUF][NT][hi|lo][x_tile_bytes]
  L.raw = (L.x_op + TL_XBUF * TL_NT * 2 * L.x_tile_bytes + 127) & ~127;
  L.stage = (L.raw + TL_CONV_WARPS * TL_RAW_STAGES * L.raw_stage_bytes + 127) & ~127;     // [epilogue warp][sub-tile][TL_STAGE_TILE]
  L.bars = L.stage + (TL_TMA_STORE ? TL_EPI_WARPS * TL_NT * TL_STAGE_TILE : 0);
  L.misc = L.bars + 48 * 8;
  L.total = L.misc + 512;
  return L;
}

// x tile, K-major [rows][KI]; h and s tiles, MN-major [k][rows] (fgrnn_tc.cu)
static __device__ __forceinline__ uint64_t tl_desc_kmajor(uint32_t smem_addr, int KI) {
  const uint64_t lbo = 128 >> 4, sbo = (uint64_t)((KI >> 3) * 128) >> 4;
  return (uint64_t)((smem_addr >> 4) & 0x3fff) | (lbo << 16) | (sbo << 32) | (1ull << 46);
}
constexpr uint32_t TL_MN_KSTEP = (2 * (TL_NS >> 3) * 128) >> 4;          // descriptor advance per 16 k of an MN-major tile
constexpr uint32_t TL_X_KSTEP = 256 >> 4;
constexpr uint32_t TL_IDESC_X = (1u << 4) | ((uint32_t)(TL_NS >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
constexpr uint32_t TL_IDESC_MN = TL_IDESC_X | (1u << 16);

struct TlEpiConst { float2 kS, k2S, cg, cu, cu2, msz, szn; float tmin; };

// rnn.py:289-295 for two rows of one unit on the packed fp32x2 pipe; same arithmetic as TcFwd::gate_update2
template <bool ONE_EX2>
static __device__ __forceinline__ float2 tl_gate_update2(float2 tot, float2 h, const TlEpiConst& k) {
  float2 ag, eg, eu;
  if (ONE_EX2) { tot.x = fmax_nan(tot.x, k.tmin); tot.y = fmax_nan(tot.y, k.tmin); }
  ag = __ffma2_rn(tot, k.kS, k.cg);                              // -(pre + b_g) log2(e)
  if (!ONE_EX2) { ag.x = fmin_nan(ag.x, 60.0f); ag.y = fmin_nan(ag.y, 60.0f); }
  eg.x = ex2_approx(ag.x); eg.y = ex2_approx(ag.y);
  if (ONE_EX2) {
    eu = __fmul2_rn(__fmul2_rn(eg, eg), k.cu);                   // e_u = e_g^2 exp(2 (b_g - b_u))
  } else {
    float2 au = __ffma2_rn(tot, k.k2S, k.cu2);
    au.x = fmin_nan(au.x, 60.0f); au.y = fmin_nan(au.y, 60.0f);
    eu.x = ex2_approx(au.x); eu.y = ex2_approx(au.y);
  }
  const float2 one = make_float2(1.0f, 1.0f);
  const float2 a = __fadd2_rn(eg, one), b = __fadd2_rn(eu, one);
  const float2 ab = __fmul2_rn(a, b);
  float2 r;
#if TL_RCP_NEWTON
  // 1 / ab on the FMA pipe (ab >= 1, finite): magic-constant estimate (<= 12 % off) + three Newton steps on the packed
  // fp32x2 pipe, <= 1 ulp -- the MUFU unit (16 lanes per SM and clock) then only serves the one EX2 per element
  r.x = __int_as_float(0x7EF311C7 - __float_as_int(ab.x)); r.y = __int_as_float(0x7EF311C7 - __float_as_int(ab.y));
  const float2 nab = make_float2(-ab.x, -ab.y);
#pragma unroll
  for (int it = 0; it < 3; ++it) r = __ffma2_rn(r, __ffma2_rn(nab, r, one), r);
#else
  r.x = rcp_approx(ab.x); r.y = rcp_approx(ab.y);
#endif
  const float2 z = __fmul2_rn(r, b);                                                          // rnn.py:290
  const float2 c = __ffma2_rn(__fmul2_rn(r, a), make_float2(2.0f, 2.0f), make_float2(-1.0f, -1.0f));   // rnn.py:292
  return __ffma2_rn(z, __ffma2_rn(k.msz, c, h), __fmul2_rn(k.szn, c));                        // rnn.py:294-295
}

struct TlEpiCtx {
  uint32_t bar_dfull, bar_hready;     // shared addresses of the [NT] barrier arrays (sub-tile s: + 8 s)
  uint32_t d2;                        // TMEM address of this thread's unit in D2 of sub-tile 0, its 16 columns (sub-tile s: + 64 s)
  unsigned char* hop;                 // h operand tile address of (sub-tile 0, k = unit, this thread's first row group), hi part
  float* out; uint32_t out_row, out_step;      // &out[first row of sub-tile 0][t = 0][unit]; element strides
  int rows_left, T, tr;               // rows_left: B - first row (sub-tile 0); tr: trace role
  const CUtensorMap* omap; float* stage;       // TL_TMA_STORE: output map; this warp's staging tiles [sub-tile][16 rows][32 units] (lane = unit)
  int o_unit0, o_row0, o_time_outer;           // box origin: first unit, first row of sub-tile 0
};

// Epilogue main loop of one warp: thread = hidden unit; columns [rh*16, rh*16 + VR) of sub-tile 0, then of sub-tile 1.
// VR = valid rows per thread and sub-tile (the other 16 - VR columns are padding: zero x, zero h, never stored): a CTA holds
// 4 VR batch rows, and the epilogue's share of the per-step chain shrinks with VR while the MMAs cost the same (fgrnn_tc.cu).
template <int VR, bool HAS_OUT, bool MASKED, bool ONE_EX2>
static __device__ __forceinline__ void tl_epilogue_loop(const TlEpiCtx& cx, const TlEpiConst& kc, float2 (&hst)[TL_NT][TL_RPT / 2]) {
  char* outp = reinterpret_cast<char*>(cx.out);
  const uint32_t row_bytes = cx.out_row * 4u;
  for (int t = 0; t < cx.T; ++t) {
#pragma unroll
    for (int s = 0; s < TL_NT; ++s) {
      mbar_wait(cx.bar_dfull + 8 * s, t & 1);
      tc_fence_after();
      if (cx.tr == 1) TL_TRACE(t, s, 6);
      if (cx.tr == 2) TL_TRACE(t, s, 9);
#pragma unroll
      for (int g = 0; g < TL_RPT / 8; ++g) {
        float v[8];
        tmem_ld8(cx.d2 + s * 64 + g * 8, v);
        tmem_ld_wait();
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int p = g * 4 + q;
          if (p >= VR / 2) { hi[q] = 0u; lo[q] = 0u; continue; }
          hst[s][p] = tl_gate_update2<ONE_EX2>(make_float2(v[2 * q], v[2 * q + 1]), hst[s][p], kc);
          const __half2 hh = __float22half2_rn(hst[s][p]);
          const float2 hf = __half22float2(hh);
          const __half2 hl = __float22half2_rn(__fadd2_rn(hst[s][p], make_float2(-hf.x, -hf.y)));
          hi[q] = *reinterpret_cast<const uint32_t*>(&hh);
          lo[q] = *reinterpret_cast<const uint32_t*>(&hl);
        }
        *reinterpret_cast<uint4*>(cx.hop + s * (2 * TL_H_TILE) + g * 128) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<uint4*>(cx.hop + s * (2 * TL_H_TILE) + TL_H_TILE + g * 128) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
      }
      if (HAS_OUT && TL_TMA_STORE) {
        // h_t -> this warp's staging tile (lane = unit: 128 contiguous bytes per row, immediate offsets); the TMA unit ships
        // the 16 x 32 box after the fence below and clips rows past the batch end
        float* st = cx.stage + s * (TL_STAGE_TILE / 4) + (threadIdx.x & 31);
#pragma unroll
        for (int q = 0; q < VR / 2; ++q) { st[(2 * q) * 32] = hst[s][q].x; st[(2 * q + 1) * 32] = hst[s][q].y; }
      }
      fence_proxy_async_smem();                        // st.shared of the h tile (and the staging tile) -> visible to the async proxy
      tc_fence_before();                               // tcgen05.ld of D2 done before stage 1 of the next step overwrites it
      __syncwarp();
      if ((threadIdx.x & 31) == 0) mbar_arrive(cx.bar_hready + 8 * s);
      if (cx.tr == 1) TL_TRACE(t, s, 7);
      if (cx.tr == 2) TL_TRACE(t, s, 10);
      if (HAS_OUT && TL_TMA_STORE) {
        if ((threadIdx.x & 31) == 0) {
          const uint32_t src = smem_u32(cx.stage + s * (TL_STAGE_TILE / 4));
          const int row = cx.o_row0 + s * 2 * VR;
          if (cx.o_time_outer)
            asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];"
                         ::"l"(reinterpret_cast<uint64_t>(cx.omap)), "r"(cx.o_unit0), "r"(row), "r"(t), "r"(src) : "memory");
          else
            asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];"
                         ::"l"(reinterpret_cast<uint64_t>(cx.omap)), "r"(cx.o_unit0), "r"(t), "r"(row), "r"(src) : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          // the other sub-tile's store may stay in flight; the one issued a step ago from the tile that is written next is read out
          asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        }
        __syncwarp();
      } else if (HAS_OUT) {
        char* o = outp + (uint32_t)(s * 2 * VR) * row_bytes;
#pragma unroll
        for (int q = 0; q < VR / 2; ++q) {
          if (!MASKED || s * 2 * VR + 2 * q < cx.rows_left) *reinterpret_cast<float*>(o + (uint32_t)(2 * q) * row_bytes) = hst[s][q].x;
          if (!MASKED || s * 2 * VR + 2 * q + 1 < cx.rows_left) *reinterpret_cast<float*>(o + (uint32_t)(2 * q + 1) * row_bytes) = hst[s][q].y;
        }
      }
      if (cx.tr == 1) TL_TRACE(t, s, 8);
    }
    if (HAS_OUT && !TL_TMA_STORE) outp += (size_t)cx.out_step * 4u;
  }
  if (HAS_OUT && TL_TMA_STORE) {
    if ((threadIdx.x & 31) == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");      // shared memory stays valid until the stores are done
    __syncwarp();
  }
}

// Stage 1 of one sub-tile step, straight-line on the elected lane.  Every accumulator takes its lo products first (B = the
// lo part of h / x: main rows get W_hi.v_lo) and the hi.hi products after them (B = the hi part: main rows get hi.hi, the
// correction rows W_lo.v_hi), see the header:
//   role 0 -> X: h k-steps 0..7 and x                  role 1 -> Y: h k-steps 8..15
template <int ROLE, int NKX, bool X_HAS_LO>
static __device__ __forceinline__ void tl_issue_stage1(uint32_t acc, uint64_t dHhi, uint64_t dHlo, uint64_t dXhi, uint64_t dXlo) {
  constexpr uint32_t tmem = 0u;
  constexpr int k0 = ROLE * 8;
#pragma unroll
  for (int ks = k0; ks < k0 + 8; ++ks) umma_ts1(acc, tmem + TLM_A1 + ks * 8, dHlo + ks * TL_MN_KSTEP, TL_IDESC_MN, ks > k0);
  if (ROLE == 0 && X_HAS_LO) {
#pragma unroll
    for (int ks = 0; ks < NKX; ++ks) umma_ts1(acc, tmem + TLM_A1X + ks * 8, dXlo + ks * TL_X_KSTEP, TL_IDESC_X, 1);
  }
#pragma unroll
  for (int ks = k0; ks < k0 + 8; ++ks) umma_ts1(acc, tmem + TLM_A1 + ks * 8, dHhi + ks * TL_MN_KSTEP, TL_IDESC_MN, 1);
  if (ROLE == 0) {
#pragma unroll
    for (int ks = 0; ks < NKX; ++ks) umma_ts1(acc, tmem + TLM_A1X + ks * 8, dXhi + ks * TL_X_KSTEP, TL_IDESC_X, 1);
  }
}

template <int VR>
__global__ void __launch_bounds__(TL_THREADS, 1) tc_lr_fwd_kernel(const TlArgs ta, const __grid_constant__ CUtensorMap xmap,
                                                                 const __grid_constant__ CUtensorMap omap) {
  extern __shared__ __align__(128) unsigned char sm[];
  const FwdArgs& a = ta.f;
  const Dims d = a.d;
  const int I = d.I, KI = ta.KI, rU = d.rU, rW = d.rW;
  const int esz = d.x_dtype == FGRNN_BF16 ? 2 : 4;
  const TlSmem L = tl_smem_layout(I, KI, esz);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + L.bars);
  uint32_t* tmem_base_s = reinterpret_cast<uint32_t*>(sm + L.misc);
  float* red_s = reinterpret_cast<float*>(sm + L.misc + 16);            // [3][16] reduction scratch, [48] = 2^-S1 for the hop warps
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
  const int row0 = blockIdx.x * (4 * VR);
  auto bar = [&](int i) { return smem_u32(&bars[i]); };
  // per sub-tile: HREADY (h_{t-1} tile written, D2 drained) | D1FULL | SREADY (s tile written, D1 drained) | DFULL
  const int B_HREADY = 0, B_D1FULL = 2, B_SREADY = 4, B_DFULL = 6, B_XFULL = 8, B_XEMPTY = 12, B_RAWFULL = 16;
  // TL_HOP_HALVES == 1: warps 16..18 hop, 19..20 x path; == 2: warps 16..23, (warp & 3) < 3 hop, == 3 x path
  constexpr int W_AUX0 = TL_EPI_WARPS, W_MMA = W_AUX0 + TL_HOP_WARPS + TL_CONV_WARPS;
  const bool x_warp = warp >= W_AUX0 && warp < W_MMA && (TL_HOP_HALVES == 2 ? (warp & 3) == 3 : warp >= W_AUX0 + 3);

  if (warp == W_MMA) tmem_alloc(smem_u32(tmem_base_s), 512);
  if (tid == 0) {
    for (int s = 0; s < TL_NT; ++s) {
      mbar_init(bar(B_HREADY + s), TL_EPI_WARPS);
      mbar_init(bar(B_D1FULL + s), 2);
      mbar_init(bar(B_SREADY + s), TL_HOP_WARPS);
      mbar_init(bar(B_DFULL + s), 2);
    }
    for (int b = 0; b < TL_XBUF; ++b) { mbar_init(bar(B_XFULL + b), TL_CONV_WARPS); mbar_init(bar(B_XEMPTY + b), TL_MMA_WARPS); }
    for (int st = 0; st < TL_CONV_WARPS * TL_RAW_STAGES; ++st) mbar_init(bar(B_RAWFULL + st), 1);
    fence_mbar_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (*tmem_base_s != 0u) __trap();
  constexpr uint32_t tmem = 0u;

  if (warp >= W_MMA) {
    // =========================== MMA issuers: two warps per sub-tile ==================================
    const int s = (warp - W_MMA) >> 1, role = (warp - W_MMA) & 1;
    const bool leader = elect_one();
    tc_fence_before();
    __syncthreads();                                   // weights in TMEM, h_{-1} / x_0 tiles under way
    tc_fence_after();
    const int nkx = KI >> 4;
    const bool x_has_lo = d.x_dtype != FGRNN_BF16;
    const uint64_t dHhi = make_desc_mn(smem_u32(sm + L.h_op + s * 2 * TL_H_TILE), TL_NS), dHlo = dHhi + (TL_H_TILE >> 4);
    const uint64_t dShi = make_desc_mn(smem_u32(sm + L.s_op + s * 2 * TL_S_TILE), TL_NS), dSlo = dShi + (TL_S_TILE >> 4);
    const uint64_t dX0 = tl_desc_kmajor(smem_u32(sm + L.x_op), KI);
    const uint32_t xlo_step = (uint32_t)L.x_tile_bytes >> 4, xtile_step = 2 * xlo_step, xbuf_step = TL_NT * xtile_step;
    const uint32_t acc1 = tmem + TLM_ACC + s * 64 + role * TL_NS;           // D1: X (role 0) | Y (role 1);  D2: tile `role`
    const uint32_t a2hi = tmem + TLM_A2 + role * 48, a2lo = a2hi + 24;
    const int variant = role ? 4 : (nkx - 1) * 2 + (x_has_lo ? 1 : 0);
    if (s) tc_spin_ns(TL_STAGGER_NS);                  // the second sub-tile starts half a period behind the first
    for (int t = 0; t < d.T; ++t) {
      const int xb = t % TL_XBUF;
      mbar_wait(bar(B_XFULL + xb), (t / TL_XBUF) & 1); // x_t operand tiles written
      mbar_wait(bar(B_HREADY + s), t & 1);             // h_{t-1} operand tile written, D2 of step t-1 drained
      tc_fence_after();
      if (role == 0) TL_TRACE(t, s, 0);
      if (leader) {
        const uint64_t dXhi = dX0 + (uint64_t)(xb * xbuf_step + s * xtile_step), dXlo = dXhi + xlo_step;
        switch (variant) {
          case 0: tl_issue_stage1<0, 1, false>(acc1, dHhi, dHlo, dXhi, dXlo); break;
          case 1: tl_issue_stage1<0, 1, true>(acc1, dHhi, dHlo, dXhi, dXlo); break;
          case 2: tl_issue_stage1<0, 2, false>(acc1, dHhi, dHlo, dXhi, dXlo); break;
          case 3: tl_issue_stage1<0, 2, true>(acc1, dHhi, dHlo, dXhi, dXlo); break;
          default: tl_issue_stage1<1, 0, false>(acc1, dHhi, dHlo, dXhi, dXlo); break;
        }
        umma_commit1(bar(B_D1FULL + s));
        umma_commit1(bar(B_XEMPTY + xb));              // this warp's MMAs have consumed the x_t tiles
      }
      __syncwarp();
      if (role == 0) TL_TRACE(t, s, 1);
      mbar_wait(bar(B_SREADY + s), t & 1);             // s tile written, D1 drained
      tc_fence_after();
      if (role == 0) TL_TRACE(t, s, 2);
      if (leader) {
        // ---- stage 2, unit tile `role`: K = 48; lo products first ----
#pragma unroll
        for (int ks = 0; ks < 3; ++ks) {
          umma_ts1(acc1, a2lo + ks * 8, dShi + ks * TL_MN_KSTEP, TL_IDESC_MN, ks > 0);
          umma_ts1(acc1, a2hi + ks * 8, dSlo + ks * TL_MN_KSTEP, TL_IDESC_MN, 1);
        }
#pragma unroll
        for (int ks = 0; ks < 3; ++ks) umma_ts1(acc1, a2hi + ks * 8, dShi + ks * TL_MN_KSTEP, TL_IDESC_MN, 1);
        umma_commit1(bar(B_DFULL + s));
      }
      __syncwarp();
      if (role == 0) TL_TRACE(t, s, 3);
    }
  } else if (x_warp) {
    // =========================== x path: TMA -> fp16 hi/lo split -> K-major operand tiles =========================
    // converter warp cw owns sub-tile cw: 32 rows, one TMA box per step, a private raw ring
    const int cw = TL_HOP_HALVES == 2 ? (warp - W_AUX0) >> 2 : warp - (W_AUX0 + 3);
    const uint32_t raw_bytes = (uint32_t)(2 * VR * I * esz);              // one TMA box: the sub-tile's 2 VR rows
    unsigned char* raw_base = sm + L.raw + cw * TL_RAW_STAGES * L.raw_stage_bytes;
    const int my_row0 = row0 + cw * 2 * VR;
    auto issue_tma = [&](int t) {
      const int st = t % TL_RAW_STAGES;
      const uint32_t fb = bar(B_RAWFULL + cw * TL_RAW_STAGES + st);
      mbar_expect_tx(fb, raw_bytes);
      if (ta.x_time_outer) tma_load_3d(smem_u32(raw_base + st * L.raw_stage_bytes), &xmap, 0, my_row0, t, fb);
      else tma_load_3d(smem_u32(raw_base + st * L.raw_stage_bytes), &xmap, 0, t, my_row0, fb);
    };
    if (lane == 0)
      for (int t = 0; t < TL_RAW_STAGES && t < d.T; ++t) issue_tma(t);
    tc_fence_before();
    __syncthreads();
    const int nch = KI >> 3, ntask = TL_CONV_ROWS * nch;          // <= 128 tasks: at most 4 per lane
    constexpr int MAXIT = TL_CONV_ROWS * (TL_MAX_KI / 8) / 32;
    uint32_t src_off[MAXIT], dst_off[MAXIT];
    bool live[MAXIT], pad[MAXIT];
#pragma unroll
    for (int it = 0; it < MAXIT; ++it) {
      const int e = it * 32 + lane;
      const int row = e / nch, ch = e - row * nch;      // row = column of the sub-tile: batch row (row >> 4) * VR + (row & 15) of it
      live[it] = e < ntask;
      pad[it] = ch * 8 >= I || (row & 15) >= VR;
      src_off[it] = (uint32_t)(((row >> 4) * VR + (row & 15)) * I * esz + ch * 8 * esz);
      dst_off[it] = (uint32_t)((cw * 2) * L.x_tile_bytes + (row >> 3) * (nch * 128) + ch * 128 + (row & 7) * 16);
    }
    const uint32_t xbuf_bytes = (uint32_t)(TL_NT * 2 * L.x_tile_bytes);
    for (int t = 0; t < d.T; ++t) {
      const int st = t % TL_RAW_STAGES, xb = t % TL_XBUF;
      mbar_wait(bar(B_RAWFULL + cw * TL_RAW_STAGES + st), (t / TL_RAW_STAGES) & 1);
      mbar_wait(bar(B_XEMPTY + xb), ((t / TL_XBUF) & 1) ^ 1);    // MMAs of step t - XBUF have finished with this buffer
      const unsigned char* raw = raw_base + st * L.raw_stage_bytes;
      unsigned char* xdst = sm + L.x_op + xb * xbuf_bytes;
#pragma unroll
      for (int it = 0; it < MAXIT; ++it) {
        if (live[it]) {
          float v[8];
          if (!pad[it]) {
            if (esz == 4) {
              const float4 p0 = *reinterpret_cast<const float4*>(raw + src_off[it]);
              const float4 p1 = *reinterpret_cast<const float4*>(raw + src_off[it] + 16);
              v[0] = p0.x; v[1] = p0.y; v[2] = p0.z; v[3] = p0.w; v[4] = p1.x; v[5] = p1.y; v[6] = p1.z; v[7] = p1.w;
            } else {
              const uint4 p = *reinterpret_cast<const uint4*>(raw + src_off[it]);
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
          *reinterpret_cast<uint4*>(xdst + dst_off[it]) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
          *reinterpret_cast<uint4*>(xdst + dst_off[it] + L.x_tile_bytes) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(bar(B_XFULL + xb));
        if (t + TL_RAW_STAGES < d.T) issue_tma(t + TL_RAW_STAGES);
      }
      __syncwarp();
    }
  } else if (warp >= W_AUX0) {
    // =========================== hop: D1 (rank rows) -> fp16 hi/lo B operand of stage 2 ============================
    // a hop warp reads lane quadrant q = warp & 3 (lanes 0..15 = the main rows of 16 ranks, lanes 16..31 = their correction
    // rows) for one 16-row half of the sub-tile
    const int quad = warp & 3;                         // the TMEM lane quadrant this warp may access
    const int half0 = TL_HOP_HALVES == 2 ? (warp - W_AUX0) >> 2 : 0;      // first 16-row half (columns of D1) of this warp
    const int k = quad * 16 + (lane & 15);             // k index of the stage-2 operand: U1 ranks 0..31, then W1 ranks
    const int half_lane = lane >> 4;                   // which 8 of a half's 16 batch rows this lane converts and stores
    const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
    tc_fence_before();
    __syncthreads();                                   // 2^-S1 is in shared memory
    const float unscale1 = red_s[48];
    unsigned char* sop0 = sm + L.s_op + (k >> 3) * ((TL_NS >> 3) * 128) + (k & 7) * 16;
    for (int t = 0; t < d.T; ++t) {
#pragma unroll
      for (int s = 0; s < TL_NT; ++s) {
        mbar_wait(bar(B_D1FULL + s), t & 1);
        tc_fence_after();
        if (quad == 0 && half0 == 0) TL_TRACE(t, s, 4);
        const uint32_t d1 = tmem + lane_base + TLM_ACC + s * 64;
#pragma unroll
        for (int half = half0; half < half0 + 2 / TL_HOP_HALVES; ++half) {
          float vx[16], vy[16];
          tmem_ld16(d1 + half * 16, vx);
          tmem_ld16(d1 + TL_NS + half * 16, vy);
          tmem_ld_wait();
          float mine[8];
#pragma unroll
          for (int c = 0; c < 16; ++c) {
            const float v = vx[c] + vy[c];
            const float tot = v + __shfl_xor_sync(0xffffffffu, v, 16);      // main row + correction row: the same sum on both lanes
            if (c < 8) mine[c] = tot; else if (half_lane) mine[c - 8] = tot;
          }
          uint32_t hi[4], lo[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) split2(mine[2 * q], mine[2 * q + 1], unscale1, hi[q], lo[q]);
          unsigned char* dst = sop0 + s * (2 * TL_S_TILE) + (half * 2 + half_lane) * 128;
          *reinterpret_cast<uint4*>(dst) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
          *reinterpret_cast<uint4*>(dst + TL_S_TILE) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
        fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(B_SREADY + s));
        if (quad == 0 && half0 == 0) TL_TRACE(t, s, 5);
      }
    }
  } else {
    // =========================== epilogue warps ===================================================
    const int ew = warp;                               // 0..15
    const int quad = ew & 3;                           // TMEM lane quadrant
    const int m = (ew >> 2) & 1;                       // unit tile
    const int rh = ew >> 3;                            // row half of each sub-tile
    const int part = ew >> 2;                          // 0..3: share of the weight upload
    const int ln = quad * 32 + lane;                   // TMEM lane
    const int gu = m * 128 + ln;                       // hidden unit of this thread
    const uint32_t lane_base = (uint32_t)(quad * 32) << 16;

    // ---- power-of-two scales from max |U1|, |W1| (stage 1) and max |U2|, |W2| (stage 2); largest bias distance
    float m1 = 0.f, m2 = 0.f;
    {
      const int et = ew * 32 + lane, nthr = TL_EPI_WARPS * 32;
      for (int i = et; i < TL_H * rU; i += nthr) { m1 = fmaxf(m1, fabsf(__ldg(a.U1c + i))); m2 = fmaxf(m2, fabsf(__ldg(a.U2c + i))); }
      for (int i = et; i < I * rW; i += nthr) m1 = fmaxf(m1, fabsf(__ldg(a.W1c + i)));
      for (int i = et; i < rW * TL_H; i += nthr) m2 = fmaxf(m2, fabsf(__ldg(a.W2c + i)));
    }
    float bd = rh == 0 ? fabsf(__ldg(a.bias_gate + gu) - __ldg(a.bias_update + gu)) : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, o));
      m2 = fmaxf(m2, __shfl_xor_sync(0xffffffffu, m2, o));
      bd = fmaxf(bd, __shfl_xor_sync(0xffffffffu, bd, o));
    }
    if (lane == 0) { red_s[ew] = m1; red_s[16 + ew] = m2; red_s[32 + ew] = bd; }
    asm volatile("bar.sync 1, %0;" ::"n"(TL_EPI_WARPS * 32) : "memory");      // epilogue warps only
#pragma unroll
    for (int w = 0; w < TL_EPI_WARPS; ++w) { m1 = fmaxf(m1, red_s[w]); m2 = fmaxf(m2, red_s[16 + w]); bd = fmaxf(bd, red_s[32 + w]); }
    const bool wide_bias = !(bd <= 8.0f);
    int S1 = 40, S2 = 40;
    if (m1 > 0.f) S1 = min(S1, (int)floorf(log2f(30000.f / m1)));
    if (m2 > 0.f) S2 = min(S2, (int)floorf(log2f(30000.f / m2)));
    S1 = max(S1, -14); S2 = max(S2, -14);
    const float scale1 = exp2f((float)S1);
    const float scale2 = exp2f((float)S2), unscale2 = exp2f((float)-S2);
    if (tid == 0) red_s[48] = exp2f((float)-S1);

    // ---- stage-1 weights -> tensor memory.  A1[lane][k]: quadrant q < 2: lanes 0..15 = the hi part of U1^T for ranks
    //      16q..16q+15, lanes 16..31 = the lo part of the same ranks (k = hidden unit); quadrant 2: W1^T likewise (k = input
    //      feature, its own columns); everything else zero (block diagonal).
    {
      const int rank = (quad & 1) * 16 + (lane & 15);
      const bool take_lo = (lane >> 4) != 0;
      const bool u_row = quad < 2 && rank < rU, w_row = quad == 2 && (lane & 15) < rW;
      for (int kb = part; kb < 16; kb += 4) {            // h part: 16 k-blocks of 16 units
        uint32_t sel[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int kk = kb * 16 + 2 * j;
          const float v0 = u_row ? __ldg(a.U1c + (size_t)kk * rU + rank) : 0.f;
          const float v1 = u_row ? __ldg(a.U1c + (size_t)(kk + 1) * rU + rank) : 0.f;
          uint32_t hi, lo;
          split2(v0, v1, scale1, hi, lo);
          sel[j] = take_lo ? lo : hi;
        }
        tmem_st8(tmem + lane_base + TLM_A1 + kb * 8, sel);
      }
      if (part < 2) {                                    // x part: 2 k-blocks of 16 features
        const int kb = part;
        uint32_t sel[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int kk = kb * 16 + 2 * j;
          const float v0 = (w_row && kk < I) ? __ldg(a.W1c + (size_t)kk * rW + (lane & 15)) : 0.f;
          const float v1 = (w_row && kk + 1 < I) ? __ldg(a.W1c + (size_t)(kk + 1) * rW + (lane & 15)) : 0.f;
          uint32_t hi, lo;
          split2(v0, v1, scale1, hi, lo);
          sel[j] = take_lo ? lo : hi;
        }
        tmem_st8(tmem + lane_base + TLM_A1X + kb * 8, sel);
      }
    }
    // ---- stage-2 weights: A2[tile][lane = unit][j]: j < 32 -> U2[j][unit], 32 <= j < 48 -> W2[j - 32][unit]
    {
      const int tile = part & 1, unit = tile * 128 + ln;
      const int kb0 = (part >> 1) ? 2 : 0, kb1 = (part >> 1) ? 3 : 2;
      for (int kb = kb0; kb < kb1; ++kb) {
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float v[2];
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int k = kb * 16 + 2 * j + e;
            v[e] = k < TL_RU ? (k < rU ? __ldg(a.U2c + (size_t)k * TL_H + unit) : 0.f)
                             : (k - TL_RU < rW ? __ldg(a.W2c + (size_t)(k - TL_RU) * TL_H + unit) : 0.f);
          }
          split2(v[0], v[1], scale2, hi[j], lo[j]);
        }
        tmem_st8(tmem + lane_base + TLM_A2 + tile * 48 + kb * 8, hi);
        tmem_st8(tmem + lane_base + TLM_A2 + tile * 48 + 24 + kb * 8, lo);
      }
    }
    tmem_st_wait();

    // ---- per-unit constants of the gate update
    TlEpiConst kc;
    bool one_ex2;
    {
      const float sz = sigmoid_f(__ldg(a.zeta)), sn = sigmoid_f(__ldg(a.nu));
      constexpr float LOG2E = 1.4426950408889634f;
      const float bgv = __ldg(a.bias_gate + gu), buv = __ldg(a.bias_update + gu);
      const float kS = -LOG2E * unscale2, cg = -LOG2E * bgv;
      kc.kS = make_float2(kS, kS); kc.cg = make_float2(cg, cg);
      kc.msz = make_float2(-sz, -sz); kc.szn = make_float2(sz + sn, sz + sn);
      const float k2S = -2.0f * LOG2E * unscale2, cu2 = -2.0f * LOG2E * buv;
      kc.k2S = make_float2(k2S, k2S); kc.cu2 = make_float2(cu2, cu2);
      one_ex2 = !wide_bias;
      const float ratio = one_ex2 ? expf(2.0f * (bgv - buv)) : 1.0f;
      kc.cu = make_float2(ratio, ratio);
      kc.tmin = (30.0f - cg) / kS;
    }

    // ---- state: h[row][gu] for rows [rh*16, rh*16 + 16) of both sub-tiles; h_{-1} operand tiles
    float2 hst[TL_NT][TL_RPT / 2];
    unsigned char* hop = sm + L.h_op + (gu >> 3) * ((TL_NS >> 3) * 128) + (rh * 2) * 128 + (gu & 7) * 16;
    const int first_row = row0 + rh * VR;              // of sub-tile 0; sub-tile 1: + 2 VR
#pragma unroll
    for (int s = 0; s < TL_NT; ++s) {
#pragma unroll
      for (int g = 0; g < TL_RPT / 8; ++g) {
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int row = first_row + s * 2 * VR + g * 8 + 2 * q;
          const bool valid = g * 8 + 2 * q < VR;
          const float v0 = (a.h0 && valid && row < d.B) ? __ldg(a.h0 + (size_t)row * TL_H + gu) : 0.f;
          const float v1 = (a.h0 && valid && row + 1 < d.B) ? __ldg(a.h0 + (size_t)(row + 1) * TL_H + gu) : 0.f;
          hst[s][g * 4 + q] = make_float2(v0, v1);
          split2(v0, v1, 1.0f, hi[q], lo[q]);
        }
        *reinterpret_cast<uint4*>(hop + s * (2 * TL_H_TILE) + g * 128) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<uint4*>(hop + s * (2 * TL_H_TILE) + TL_H_TILE + g * 128) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
      }
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();                                   // matches the other roles' prologue barrier
    __syncwarp();
    if (lane == 0) { mbar_arrive(bar(B_HREADY)); mbar_arrive(bar(B_HREADY + 1)); }    // phase 0: h_{-1} ready

    TlEpiCtx cx;
    cx.bar_dfull = bar(B_DFULL); cx.bar_hready = bar(B_HREADY);
    cx.d2 = tmem + lane_base + TLM_ACC + m * TL_NS + rh * TL_RPT;
    cx.hop = hop;
    cx.out = a.out ? a.out + (size_t)first_row * a.osb + gu : nullptr;
    cx.out_row = (uint32_t)a.osb; cx.out_step = (uint32_t)a.ost;
    cx.rows_left = d.B - first_row; cx.T = d.T;
    cx.tr = ew == 0 ? 1 : (ew == 15 ? 2 : 0);
    cx.omap = &omap; cx.stage = reinterpret_cast<float*>(sm + L.stage) + ew * TL_NT * (TL_STAGE_TILE / 4);
    cx.o_unit0 = m * 128 + quad * 32; cx.o_row0 = first_row; cx.o_time_outer = ta.o_time_outer;
    const bool masked = !TL_TMA_STORE && row0 + 4 * VR > d.B;      // the TMA unit clips its boxes at the batch end
    const int variant = (one_ex2 ? 4 : 0) | (a.out ? 2 : 0) | (masked ? 1 : 0);
    switch (variant) {
      case 0: tl_epilogue_loop<VR, false, false, false>(cx, kc, hst); break;
      case 1: tl_epilogue_loop<VR, false, true, false>(cx, kc, hst); break;
      case 2: tl_epilogue_loop<VR, true, false, false>(cx, kc, hst); break;
      case 3: tl_epilogue_loop<VR, true, true, false>(cx, kc, hst); break;
      case 4: tl_epilogue_loop<VR, false, false, true>(cx, kc, hst); break;
      case 5: tl_epilogue_loop<VR, false, true, true>(cx, kc, hst); break;
      case 6: tl_epilogue_loop<VR, true, false, true>(cx, kc, hst); break;
      default: tl_epilogue_loop<VR, true, true, true>(cx, kc, hst); break;
    }
    if (a.h_last) {
#pragma unroll
      for (int s = 0; s < TL_NT; ++s) {
#pragma unroll
        for (int j = 0; j < VR; ++j) {
          const int row = first_row + s * 2 * VR + j;
          if (row < d.B) a.h_last[(size_t)row * TL_H + gu] = (j & 1) ? hst[s][j >> 1].y : hst[s][j >> 1].x;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == W_MMA) tmem_dealloc(tmem, 512);
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
bool tc_lr_supports(const Dims& d) {
  return d.H == TL_H && d.rU > 0 && d.rU <= TL_RU && d.rW > 0 && d.rW <= TL_RW && d.I >= 8 && d.I <= TL_MAX_KI && (d.I % 8) == 0 &&
         d.gate_nl == FGRNN_NL_SIGMOID && d.update_nl == FGRNN_NL_TANH;
}

template <int VR>
static int launch_tc_lr_fwd_vr(const FwdArgs& a, cudaStream_t stream) {
  const Dims& d = a.d;
  TlArgs ta{};
  ta.f = a;
  ta.KI = (d.I + 15) & ~15;
  const int esz = d.x_dtype == FGRNN_BF16 ? 2 : 4;
  CUtensorMap map;
  const int rc = make_row_tile_map(&map, a.x, d.x_dtype == FGRNN_BF16, d.I, d.B, d.T, a.xsb, a.xst, 2 * VR, &ta.x_time_outer);
  if (rc) return rc;
  CUtensorMap omap = map;                              // a valid descriptor even when there is no output tensor
  if (a.out && TL_TMA_STORE) {
    const int rc2 = make_row_tile_map(&omap, a.out, false, TL_H, d.B, d.T, a.osb, a.ost, VR, &ta.o_time_outer, 32);
    if (rc2) return rc2;
  }
  const TlSmem L = tl_smem_layout(d.I, ta.KI, esz);
  FGRNN_CUDA_TRY(cudaFuncSetAttribute(tc_lr_fwd_kernel<VR>, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total));
  const unsigned grid = (unsigned)((d.B + 4 * VR - 1) / (4 * VR));
  tc_lr_fwd_kernel<VR><<<grid, TL_THREADS, L.total, stream>>>(ta, map, omap);
  FGRNN_LAUNCH_CHECK("tc_lr_fwd_kernel");
  return FGRNN_OK;
}

// Valid rows per epilogue thread.  All CTAs take the same time, so a launch lasts rounds x chain with rounds = ceil(CTAs / 148);
// the model below (chain ~ 3630 + 132 VR cycles per step: tools/trace_lowrank.py, 5750 at VR = 16 of which the epilogue 2120) only
// ranks the candidates.  32768 rows: 512 CTAs of 64 rows are 3.46 rounds and take 4; 586 CTAs of 56 rows are 3.96 and take 4
// shorter ones -- measured 1.126 -> 1.103 ms (same box, A/B).  FGRNN_TC_VR=16|14|12 overrides.
int launch_tc_lr_fwd(const FwdArgs& a, cudaStream_t stream) {
  if (a.d.B <= 0 || a.d.T <= 0) return FGRNN_OK;
  int vr = tuning(TUNE_TC_VR);
  if (vr != 16 && vr != 14 && vr != 12) {
    double best = 0.0;
    for (int cand = 16; cand >= 12; cand -= 2) {
      const int ctas = (a.d.B + 4 * cand - 1) / (4 * cand);
      const double cost = (double)((ctas + 147) / 148) * (3630.0 + 132.0 * cand);
      if (cand == 16 || cost < best * 0.98) { best = cost; vr = cand; }
    }
  }
  if (vr == 14) return launch_tc_lr_fwd_vr<14>(a, stream);
  if (vr == 12) return launch_tc_lr_fwd_vr<12>(a, stream);
  return launch_tc_lr_fwd_vr<16>(a, stream);
}

}  // namespace fgrnn

#ifdef FGRNN_TL_TRACE
extern "C" __attribute__((visibility("default"))) int fgrnn_debug_tl_trace(long long* host_out) {
  return cudaMemcpyFromSymbol(host_out, fgrnn::g_tl_trace, sizeof(long long) * 8 * 2 * 16) == cudaSuccess ? 0 : 6;
}
#endif
