// tcgen05 kernels of the backward pass (FGRNN_PATH_TCGEN05).
//
// tc_contract_kernel -- the T-parallel outer-product sums of BPTT (cuda/fastgrnn_cuda_kernel.cu:537-540):
//     dU[k][n] = sum_{t,b} h_{t-1}[b][k] * dpre_t[b][n]          dW[i][n] = sum_{t,b} x_t[b][i] * dpre_t[b][n]
// as split-K GEMMs over all M = T*B rows.  A chunk is 64 batch rows of one time step; CTAs take chunks
// round-robin (persistent, one CTA per SM) and keep their partial sums in tensor memory until the end:
//     dU  : A = h_{t-1} tile (MN-major: unit k contiguous), B = dpre tile (MN-major), M = N = 128
//     dW^T: A = dpre tile, B = x tile (MN-major: feature contiguous), M = 128, N = KI
// fp32 operands are split into bf16 hi + lo on their way into the shared-memory operand tiles (gradients sit far
// below the fp16 normal range; 16 mantissa bits are ample for the 1e-4 gradient tolerance); each product is
// hi.hi (accumulator *M) + lo.hi + hi.lo (accumulator *C).  16 converter warps
// stream the chunk from HBM (LDG.128, eight rows x 128 B per warp instruction), one warp issues the 24 MMAs of
// a chunk on an elected lane; two operand buffers let conversion of chunk c+1 overlap the MMAs of chunk c.
// Every CTA writes its partial [I][H] / [H][H] in the canonical orientation; fgrnn_generic.cu's reduce_kernel
// sums the partials in a fixed order (no atomics) into the caller's gradient tensors.
#include "fgrnn_kernels.cuh"
#include "fgrnn_tc_common.cuh"

namespace fgrnn {

constexpr int CT_H = 128;
constexpr int CT_ROWS = 64;                     // rows (= GEMM K) per chunk: four k-steps of 16
constexpr int CT_CONV_WARPS = 16;
constexpr int CT_THREADS = 32 * (CT_CONV_WARPS + 1);
constexpr int CT_TM_UM = 0, CT_TM_UC = 128, CT_TM_WM = 256, CT_TM_WC = 320;   // TMEM columns (WM/WC: KI <= 64 wide)
constexpr int CT_HTILE = CT_ROWS * CT_H * 2;    // bytes of one fp16 [64 x 128] operand tile

struct TcContractArgs {
  int B, T, I, KI;
  int x_dtype;
  const void* x; int64_t xsb, xst;
  const float* hs; int64_t hsb, hst;            // hidden states of the forward pass, [B,T,H] by strides
  const float* h0;                              // [B,H] or null (zeros)
  const float* dpre;                            // [T][B][H] contiguous
  float* partW;                                 // [gridDim.x][I][H]  (null: skip dW)
  float* partU;                                 // [gridDim.x][H][H]  (null: skip dU)
  int nbblk, nchunk;                            // ceil(B/64), T*nbblk
  // co-running with the reverse recurrence (tc_bwd_fused_kernel): chunks are taken in REVERSE time order and chunk
  // (t, rows) waits until the recurrence CTAs that own its rows have published T - t finished steps
  const int* progress;                          // [rec CTAs][16 epilogue warps] finished steps, null = dpre is complete
  int rec_rows, nrec;                           // rows per recurrence CTA, number of recurrence CTAs
};

struct CtSmem { int h, d, x, bars, total; int xtile; };
__host__ __device__ inline CtSmem ct_smem_layout(int KI) {
  CtSmem L;
  L.xtile = CT_ROWS * KI * 2;
  L.h = 0;                                      // [2 buffers][hi|lo][CT_HTILE]
  L.d = L.h + 2 * 2 * CT_HTILE;
  L.x = L.d + 2 * 2 * CT_HTILE;                 // [2 buffers][hi|lo][xtile]
  L.bars = L.x + 2 * 2 * L.xtile;
  L.total = L.bars + 64;
  return L;
}

__device__ __forceinline__ void ct_split8(const float (&v)[8], uint4& hi, uint4& lo) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) split2_bf16(v[2 * q], v[2 * q + 1], h[q], l[q]);
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}

static __device__ __forceinline__ void ct_run(const TcContractArgs& a, const int cta, const int ncta) {
  extern __shared__ __align__(128) unsigned char sm[];
  const CtSmem L = ct_smem_layout(a.KI);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + L.bars);      // [0,1] full, [2,3] empty, [4] done
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;   // shuffle: provably warp uniform
  auto bar = [&](int i) { return smem_u32(&bars[i]); };
  const bool want_u = a.partU != nullptr, want_w = a.partW != nullptr;

  if (warp == CT_CONV_WARPS) tmem_alloc(smem_u32(&tmem_base_s), 512);
  if (tid == 0) {
    mbar_init(bar(0), CT_CONV_WARPS); mbar_init(bar(1), CT_CONV_WARPS);
    mbar_init(bar(2), 1); mbar_init(bar(3), 1); mbar_init(bar(4), 1);
    fence_mbar_init();
  }
  tc_fence_before();
  asm volatile("bar.sync 2, %0;" ::"n"(CT_THREADS) : "memory");     // the 17 warps of the contraction (the fused kernel has 3 more, exited)
  tc_fence_after();
  if (tmem_base_s != 0u) __trap();               // the CTA owns all 512 columns
  constexpr uint32_t tmem = 0u;
  const int my_chunks = a.nchunk > cta ? (a.nchunk - 1 - cta) / ncta + 1 : 0;
  const bool follow = a.progress != nullptr;

  if (warp == CT_CONV_WARPS) {
    // =========================== MMA issuer =====================================================
    const bool leader = elect_one();
    const uint32_t idesc_u = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(CT_H >> 3) << 17) | ((uint32_t)(CT_H >> 4) << 24);
    const uint32_t idesc_w = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(a.KI >> 3) << 17) | ((uint32_t)(CT_H >> 4) << 24);
    const uint32_t hk = (2 * (CT_H >> 3) * 128) >> 4, xk = (uint32_t)(2 * (a.KI >> 3) * 128) >> 4;   // per 16 rows
    for (int j = 0; j < my_chunks; ++j) {
      const int b = j & 1;
      mbar_wait(bar(b), (j >> 1) & 1);
      tc_fence_after();
      if (leader) {
        const uint64_t dHhi = make_desc_mn(smem_u32(sm + L.h + (b * 2) * CT_HTILE), CT_H), dHlo = dHhi + (CT_HTILE >> 4);
        const uint64_t dDhi = make_desc_mn(smem_u32(sm + L.d + (b * 2) * CT_HTILE), CT_H), dDlo = dDhi + (CT_HTILE >> 4);
        const uint64_t dXhi = make_desc_mn(smem_u32(sm + L.x + (b * 2) * L.xtile), a.KI), dXlo = dXhi + ((uint32_t)L.xtile >> 4);
        const uint32_t acc0 = j > 0;
        if (want_u) {
#pragma unroll
          for (int ks = 0; ks < CT_ROWS / 16; ++ks) {
            umma_ss1(tmem + CT_TM_UM, dHhi + ks * hk, dDhi + ks * hk, idesc_u, acc0 | (ks > 0));
            umma_ss1(tmem + CT_TM_UC, dHlo + ks * hk, dDhi + ks * hk, idesc_u, acc0 | (ks > 0));
            umma_ss1(tmem + CT_TM_UC, dHhi + ks * hk, dDlo + ks * hk, idesc_u, 1);
          }
        }
        if (want_w) {
#pragma unroll
          for (int ks = 0; ks < CT_ROWS / 16; ++ks) {
            umma_ss1(tmem + CT_TM_WM, dDhi + ks * hk, dXhi + ks * xk, idesc_w, acc0 | (ks > 0));
            umma_ss1(tmem + CT_TM_WC, dDlo + ks * hk, dXhi + ks * xk, idesc_w, acc0 | (ks > 0));
            if (a.x_dtype != FGRNN_BF16) umma_ss1(tmem + CT_TM_WC, dDhi + ks * hk, dXlo + ks * xk, idesc_w, 1);   // bf16 x has no lo part
          }
        }
        umma_commit1(bar(2 + b));                      // operand buffer b may be refilled
        if (j == my_chunks - 1) umma_commit1(bar(4));  // all partial sums are final
      }
      __syncwarp();
    }
  } else {
    // =========================== converters: HBM -> fp16 hi/lo operand tiles ======================
    // warp task = (matrix, 8-row block rb, 32-column block cb); lane = (column group cgl = lane/8, row rl = lane%8):
    // a quarter warp writes one 128-byte core matrix, the warp reads 8 rows x 128 B.
    const int cgl = lane >> 3, rl = lane & 7;
    const int nxcb = (a.KI + 31) / 32;                  // column blocks of the x tile
    const int ntask = 64 + 8 * nxcb;                    // 32 (h) + 32 (dpre) + x
    constexpr int MAXT = (64 + 8 * 2 + CT_CONV_WARPS - 1) / CT_CONV_WARPS;   // <= 5 tasks per warp
    const int esz = a.x_dtype == FGRNN_BF16 ? 2 : 4;
    for (int j = 0; j < my_chunks; ++j) {
      const int fwd_chunk = cta + j * ncta;
      const int chunk = follow ? a.nchunk - 1 - fwd_chunk : fwd_chunk;          // following the recurrence: last time step first
      const int t = chunk / a.nbblk, b0 = (chunk - t * a.nbblk) * CT_ROWS;
      const int b = j & 1;
      if (follow) {
        // dpre_t of these 64 rows is written by the 16 epilogue warps of 64 / rec_rows recurrence CTAs: one flag per lane
        const int first = b0 / a.rec_rows, ncta_rows = CT_ROWS / a.rec_rows;
        const int fl = lane < ncta_rows * 16 ? first * 16 + lane : -1;
        const int need = a.T - t;                        // published counts are multiples of BR_PUBLISH, or T
        bool ok = fl < 0 || fl >= a.nrec * 16;
        for (unsigned spins = 0; !__all_sync(0xffffffffu, ok); ++spins) {
          if (!ok) {
            int v;
            asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(a.progress + fl) : "memory");
            ok = v >= need;
          }
          if (spins > (1u << 24)) __trap();              // the recurrence CTAs never wait for anybody: a bug, not a deadlock
          if (!ok) __nanosleep(200);
        }
        __syncwarp();
      }
      float v[MAXT][8];
#pragma unroll
      for (int k = 0; k < MAXT; ++k) {
        const int task = warp + k * CT_CONV_WARPS;
#pragma unroll
        for (int q = 0; q < 8; ++q) v[k][q] = 0.f;
        if (task < ntask) {
          const int mat = task < 32 ? 0 : (task < 64 ? 1 : 2);
          const int tt = task - mat * 32;
          const int rb = mat < 2 ? (tt >> 2) : (tt / nxcb), cb = mat < 2 ? (tt & 3) : (tt - rb * nxcb);
          const int row = b0 + rb * 8 + rl, col = cb * 32 + cgl * 8;
          if (row < a.B) {
            if (mat == 0) {               // h_{t-1}: the hidden-state tensor shifted by one step, h0 at t == 0
              const float* src = t > 0 ? a.hs + (int64_t)row * a.hsb + (int64_t)(t - 1) * a.hst + col
                                       : (a.h0 ? a.h0 + (size_t)row * CT_H + col : nullptr);
              if (src && want_u) {
                const float4 p0 = __ldg(reinterpret_cast<const float4*>(src)), p1 = __ldg(reinterpret_cast<const float4*>(src) + 1);
                v[k][0] = p0.x; v[k][1] = p0.y; v[k][2] = p0.z; v[k][3] = p0.w; v[k][4] = p1.x; v[k][5] = p1.y; v[k][6] = p1.z; v[k][7] = p1.w;
              }
            } else if (mat == 1) {
              // dpre may have been written moments ago by a recurrence CTA of the same launch: L2-coherent loads, never the
              // read-only path
              const float* src = a.dpre + ((size_t)t * a.B + row) * CT_H + col;
              const float4 p0 = __ldcg(reinterpret_cast<const float4*>(src)), p1 = __ldcg(reinterpret_cast<const float4*>(src) + 1);
              v[k][0] = p0.x; v[k][1] = p0.y; v[k][2] = p0.z; v[k][3] = p0.w; v[k][4] = p1.x; v[k][5] = p1.y; v[k][6] = p1.z; v[k][7] = p1.w;
            } else if (col < a.I && want_w) {
              const int64_t off = (int64_t)row * a.xsb + (int64_t)t * a.xst + col;
              if (esz == 4) {
                const float* src = reinterpret_cast<const float*>(a.x) + off;
                const float4 p0 = __ldg(reinterpret_cast<const float4*>(src)), p1 = __ldg(reinterpret_cast<const float4*>(src) + 1);
                v[k][0] = p0.x; v[k][1] = p0.y; v[k][2] = p0.z; v[k][3] = p0.w; v[k][4] = p1.x; v[k][5] = p1.y; v[k][6] = p1.z; v[k][7] = p1.w;
              } else {
                const uint4 p = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(a.x) + off));
                const uint32_t w[4] = {p.x, p.y, p.z, p.w};
#pragma unroll
                for (int q = 0; q < 4; ++q) { v[k][2 * q] = __uint_as_float(w[q] << 16); v[k][2 * q + 1] = __uint_as_float(w[q] & 0xffff0000u); }
              }
            }
          }
        }
      }
      if (j >= 2) mbar_wait(bar(2 + b), ((j >> 1) - 1) & 1);      // MMAs of chunk j-2 have finished with this buffer
#pragma unroll
      for (int k = 0; k < MAXT; ++k) {
        const int task = warp + k * CT_CONV_WARPS;
        if (task < ntask) {
          const int mat = task < 32 ? 0 : (task < 64 ? 1 : 2);
          const int tt = task - mat * 32;
          const int rb = mat < 2 ? (tt >> 2) : (tt / nxcb), cb = mat < 2 ? (tt & 3) : (tt - rb * nxcb);
          const int cg = cb * 4 + cgl;                  // 8-column group inside the tile
          uint4 hi, lo;
          ct_split8(v[k], hi, lo);
          if (mat < 2) {
            unsigned char* dst = sm + (mat == 0 ? L.h : L.d) + (b * 2) * CT_HTILE + rb * ((CT_H >> 3) * 128) + cg * 128 + rl * 16;
            *reinterpret_cast<uint4*>(dst) = hi;
            *reinterpret_cast<uint4*>(dst + CT_HTILE) = lo;
          } else if (cg * 8 < a.KI) {
            unsigned char* dst = sm + L.x + (b * 2) * L.xtile + rb * ((a.KI >> 3) * 128) + cg * 128 + rl * 16;
            *reinterpret_cast<uint4*>(dst) = hi;
            *reinterpret_cast<uint4*>(dst + L.xtile) = lo;
          }
        }
      }
      fence_proxy_async_smem();                         // st.shared of the operand tiles -> visible to tcgen05.mma
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(b));
    }
    // =========================== partial sums -> global (canonical [K][N]) =======================
    if (my_chunks > 0) {
      mbar_wait(bar(4), 0);
      tc_fence_after();
    }
    const int quad = warp & 3, part = warp >> 2;         // TMEM lane quadrant, column quarter
    const int m = quad * 32 + lane;
    const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
    if (want_u) {
      float* dst = a.partU + ((size_t)cta * CT_H + m) * CT_H + part * 32;      // row m = k, columns n
#pragma unroll
      for (int c0 = 0; c0 < 32; c0 += 16) {
        float vm[16], vc[16];
        if (my_chunks > 0) {
          tmem_ld16(tmem + lane_base + CT_TM_UM + part * 32 + c0, vm);
          tmem_ld16(tmem + lane_base + CT_TM_UC + part * 32 + c0, vc);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int q = 0; q < 16; ++q) { vm[q] = 0.f; vc[q] = 0.f; }
        }
#pragma unroll
        for (int q = 0; q < 16; q += 4)
          *reinterpret_cast<float4*>(dst + c0 + q) = make_float4(vm[q] + vc[q], vm[q + 1] + vc[q + 1], vm[q + 2] + vc[q + 2], vm[q + 3] + vc[q + 3]);
      }
    }
    if (want_w) {
      // D[m = n][col = i]: this warp takes columns [part*KI/4, (part+1)*KI/4) in steps of 4 (KI is a multiple of 16)
      const int cper = a.KI >> 2;
      for (int c0 = part * cper; c0 < (part + 1) * cper; c0 += 4) {
        float vm[4], vc[4];
        if (my_chunks > 0) {
          uint32_t r0[4], r1[4];
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r0[0]), "=r"(r0[1]), "=r"(r0[2]), "=r"(r0[3]) : "r"(tmem + lane_base + CT_TM_WM + c0) : "memory");
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r1[0]), "=r"(r1[1]), "=r"(r1[2]), "=r"(r1[3]) : "r"(tmem + lane_base + CT_TM_WC + c0) : "memory");
          tmem_ld_wait();
#pragma unroll
          for (int q = 0; q < 4; ++q) { vm[q] = __uint_as_float(r0[q]); vc[q] = __uint_as_float(r1[q]); }
        } else {
#pragma unroll
          for (int q = 0; q < 4; ++q) { vm[q] = 0.f; vc[q] = 0.f; }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (c0 + q < a.I) a.partW[((size_t)cta * a.I + c0 + q) * CT_H + m] = vm[q] + vc[q];
      }
    }
  }
  tc_fence_before();
  asm volatile("bar.sync 2, %0;" ::"n"(CT_THREADS) : "memory");     // the 17 warps of the contraction (the fused kernel has 3 more, exited)
  if (warp == CT_CONV_WARPS) tmem_dealloc(tmem, 512);
}

__global__ void __launch_bounds__(CT_THREADS, 1) tc_contract_kernel(const TcContractArgs a) {
  ct_run(a, (int)blockIdx.x, (int)gridDim.x);
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
bool tc_contract_supports(const Dims& d) {
  return d.H == CT_H && d.I >= 8 && d.I <= 64 && (d.I % 8) == 0;
}

int tc_contract_ctas(const Dims& d) {
  const int64_t nchunk = (int64_t)d.T * ((d.B + CT_ROWS - 1) / CT_ROWS);
  return (int)(nchunk < 148 ? (nchunk > 0 ? nchunk : 1) : 148);
}

int launch_tc_contract(const TcContractLaunch& c, cudaStream_t stream) {
  TcContractArgs a{};
  a.B = c.d.B; a.T = c.d.T; a.I = c.d.I; a.KI = (c.d.I + 15) & ~15;
  a.x_dtype = c.d.x_dtype;
  a.x = c.x; a.xsb = c.xsb; a.xst = c.xst;
  a.hs = c.hs; a.hsb = c.hsb; a.hst = c.hst; a.h0 = c.h0; a.dpre = c.dpre;
  a.partW = c.partW; a.partU = c.partU;
  a.nbblk = (c.d.B + CT_ROWS - 1) / CT_ROWS;
  a.nchunk = c.d.T * a.nbblk;
  const CtSmem L = ct_smem_layout(a.KI);
  FGRNN_CUDA_TRY(cudaFuncSetAttribute(tc_contract_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total));
  tc_contract_kernel<<<tc_contract_ctas(c.d), CT_THREADS, L.total, stream>>>(a);
  FGRNN_LAUNCH_CHECK("tc_contract_kernel");
  return FGRNN_OK;
}

// =============================================================================================
// tc_bwd_rec_kernel -- the serial part of BPTT (cuda/fastgrnn_cuda_kernel.cu:109-118, 537) as a persistent
// reverse kernel on the tensor cores, the mirror image of fgrnn_tc.cu:
//     G = g_t + delta;   dc = (sz (1 - z) + sn)(1 - c^2) G;   dz = (h_{t-1} - sz c) z (1 - z) G;   dpre = dc + dz
//     delta <- z G + dpre . U^T            (after t = 0: delta = d h0)
//   A operand = U (bf16 hi/lo) resident in tensor memory, lane = k (the unit of delta), column pair = n
//   B operand = dpre_t tile (bf16 hi/lo, MN-major) written by the epilogue warps, N = 32 rows per sub-tile
//   D         = CA | CB | M1 | M2 fp32 accumulators per sub-tile, 8 MMAs from each of three issuing warps
// One CTA = 64 batch rows = two sub-tiles owned by 8 epilogue warps each (thread = unit, 16 rows), one producer
// warp streams grad_h / z / c / h_{t-1} tiles into a 3-stage shared-memory ring with TMA
// (one 16 KB TMA box per array, 3-D maps over the caller's strides).  The bias / zeta / nu sums
// live in per-thread registers for the whole kernel and leave as one partial row per CTA.
// =============================================================================================
constexpr int BR_H = 128, BR_NT = 2;
#ifndef FGRNN_BR_PUBLISH
#define FGRNN_BR_PUBLISH 8
#endif
constexpr int BR_PUBLISH = FGRNN_BR_PUBLISH;                   // fused launch: steps between two progress reports to the contraction CTAs
constexpr int BR_EPI_WARPS = 16, BR_MMA_WARPS = 3;
constexpr int BR_W_PROD = BR_EPI_WARPS, BR_W_MMA = BR_EPI_WARPS + 1;
constexpr int BR_THREADS = 32 * (BR_EPI_WARPS + 1 + BR_MMA_WARPS);      // 640
constexpr int BR_TM_U_HI = 0, BR_TM_U_MID = 64, BR_TM_U_LO = 128, BR_TM_ACC = 192;
struct BrSmem { int stage, op, bars, red, total; };
struct BrMaps { CUtensorMap g, z, c, hs, h0; };

// BR_NS = rows per sub-tile (UMMA N): 32, or 16 when the batch fits one wave of 32-row CTAs (shorter per-step chain,
// twice as many SMs busy on small training batches).
template <int BR_NS>
struct BrK {
static constexpr int BR_ROWS = BR_NS * BR_NT;
static constexpr int RPT = BR_NS / 2, PAIRS = RPT / 2, NG = RPT / 8;    // rows / row pairs / 8-row groups per thread
static constexpr int BR_ARR = BR_NS * BR_H * 4;                         // ring unit: one [NS][128] fp32 tile
// The ring is counted in units; slot q (one sub-tile step: g | z | c | h_{t-1}) takes units (4q .. 4q+3) mod BR_RU.
// 16-row sub-tiles: 24 units = 6 whole slots; 32-row sub-tiles: 11 units (2.75 slots) is what fits beside the
// three operand tiles per sub-tile.
static constexpr int BR_RU = BR_NS == 32 ? 11 : 24;
static constexpr int BR_NB = (BR_RU + 3) / 4;                           // slot barriers (slots resident at once)
static constexpr int BR_OPT = BR_NS * BR_H * 2;                         // one bf16 operand tile
static constexpr int BR_TM_ACC_PER_TILE = 4 * BR_NS;
static constexpr uint32_t BR_IDESC = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(BR_NS >> 3) << 17) | ((uint32_t)(BR_H >> 4) << 24);
static constexpr uint32_t BR_KSTEP = (2 * (BR_NS >> 3) * 128) >> 4;

static __host__ __device__ inline BrSmem br_smem_layout() {
  BrSmem L;
  L.stage = 0;
  L.op = BR_RU * BR_ARR;                        // [NT][hi|mid|lo][BR_OPT]
  L.bars = L.op + BR_NT * 3 * BR_OPT;
  L.red = L.bars + 16 * 8;                      // 32 floats of reduction scratch
  L.total = L.red + 32 * 4;
  return L;
}

static __device__ __forceinline__ void run(const SmemBwdArgs& a, const BrMaps& maps, const int g_time_outer, const int hs_time_outer, int* progress = nullptr) {
  extern __shared__ __align__(128) unsigned char sm[];
  const BrSmem L = br_smem_layout();
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + L.bars);
  __shared__ uint32_t tmem_base_s;
  const Dims d = a.d;
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;   // shuffle: provably warp uniform
  const int row0 = blockIdx.x * BR_ROWS;
  auto bar = [&](int i) { return smem_u32(&bars[i]); };
  const int B_HREADY = 0, B_DFULL = 2, B_SFULL = 4, B_SEMPTY = 4 + BR_NB;      // operand ready | accumulators ready | ring

  if (warp == BR_W_MMA) tmem_alloc(smem_u32(&tmem_base_s), 512);
  if (tid == 0) {
    for (int s = 0; s < BR_NT; ++s) { mbar_init(bar(B_HREADY + s), BR_EPI_WARPS / BR_NT); mbar_init(bar(B_DFULL + s), BR_MMA_WARPS); }
    for (int st = 0; st < BR_NB; ++st) { mbar_init(bar(B_SFULL + st), 1); mbar_init(bar(B_SEMPTY + st), BR_EPI_WARPS / BR_NT); }
    fence_mbar_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (tmem_base_s != 0u) __trap();
  constexpr uint32_t tmem = 0u;
  const bool hp_zero_at_t0 = a.h0 == nullptr;

  if (warp >= BR_W_MMA) {
    // =========================== MMA issuers ======================================================
    const int role = warp - BR_W_MMA;
    const bool leader = elect_one();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint64_t dD0 = make_desc_mn(smem_u32(sm + L.op), BR_NS);
    for (int it = 0; it < d.T; ++it) {
#pragma unroll
      for (int s = 0; s < BR_NT; ++s) {
        const uint64_t dhi = dD0 + (uint64_t)(s * ((3 * BR_OPT) >> 4)), dmid = dhi + (BR_OPT >> 4), dlo = dmid + (BR_OPT >> 4);
        const uint32_t acc = tmem + BR_TM_ACC + s * BR_TM_ACC_PER_TILE;
        mbar_wait(bar(B_HREADY + s), it & 1);          // dpre_t operand tiles written, accumulators drained
        tc_fence_after();
        if (leader) {
          // six bf16 products of the three-term splits, 16 MMAs per issuing warp, grouped by magnitude so that the
          // tensor core's accumulate truncation (2^-25 of the largest addend) stays relative to each group:
          //   role 0 -> CA : U_mid.d_hi + U_hi.d_mid                     (2^-8)
          //   role 1 -> CB : U_lo.d_hi + U_hi.d_lo                       (2^-16)
          //   role 2 -> M1 / M2 : U_hi.d_hi + U_mid.d_mid, k-steps 0..3 / 4..7
          if (role == 0) {
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {
              umma_ts1(acc, tmem + BR_TM_U_MID + ks * 8, dhi + ks * BR_KSTEP, BR_IDESC, ks > 0);
              umma_ts1(acc, tmem + BR_TM_U_HI + ks * 8, dmid + ks * BR_KSTEP, BR_IDESC, 1);
            }
          } else if (role == 1) {
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {
              umma_ts1(acc + BR_NS, tmem + BR_TM_U_LO + ks * 8, dhi + ks * BR_KSTEP, BR_IDESC, ks > 0);
              umma_ts1(acc + BR_NS, tmem + BR_TM_U_HI + ks * 8, dlo + ks * BR_KSTEP, BR_IDESC, 1);
            }
          } else {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              umma_ts1(acc + 2 * BR_NS, tmem + BR_TM_U_HI + ks * 8, dhi + ks * BR_KSTEP, BR_IDESC, ks > 0);
              umma_ts1(acc + 2 * BR_NS, tmem + BR_TM_U_MID + ks * 8, dmid + ks * BR_KSTEP, BR_IDESC, 1);
            }
#pragma unroll
            for (int ks = 4; ks < 8; ++ks) {
              umma_ts1(acc + 3 * BR_NS, tmem + BR_TM_U_HI + ks * 8, dhi + ks * BR_KSTEP, BR_IDESC, ks > 4);
              umma_ts1(acc + 3 * BR_NS, tmem + BR_TM_U_MID + ks * 8, dmid + ks * BR_KSTEP, BR_IDESC, 1);
            }
          }
          umma_commit1(bar(B_DFULL + s));
        }
        __syncwarp();
      }
    }
  } else if (warp == BR_W_PROD) {
    // =========================== producer: one TMA box per array per sub-tile step ==================
    // grad_h and the hidden states by the caller's strides, z_s / c_s [T][B][H]; rows past the batch end are
    // zero-filled by the TMA unit, so the epilogue needs no masking of its sums.
    tc_fence_before();
    __syncthreads();
    if (lane == 0) {
      for (int it = 0; it < d.T; ++it) {
        const int t = d.T - 1 - it;
        for (int s = 0; s < BR_NT; ++s) {
          const int q = it * BR_NT + s;
          const int first = row0 + s * BR_NS;
          const bool with_hp = !(t == 0 && hp_zero_at_t0);
          // the units of slot q were last used by slots (4q - RU)/4 and (4q + 3 - RU)/4: both must be consumed
          if (4 * q >= BR_RU) { const int qa = (4 * q - BR_RU) / 4; mbar_wait(bar(B_SEMPTY + qa % BR_NB), (qa / BR_NB) & 1); }
          if (4 * q + 3 >= BR_RU && (BR_RU & 3)) { const int qb = (4 * q + 3 - BR_RU) / 4; mbar_wait(bar(B_SEMPTY + qb % BR_NB), (qb / BR_NB) & 1); }
          const uint32_t fb = bar(B_SFULL + q % BR_NB), ring = smem_u32(sm + L.stage);
          const int u0 = (4 * q) % BR_RU;
          auto unit = [&](int k) { return ring + (uint32_t)(((u0 + k) % BR_RU) * BR_ARR); };
          const bool with_g = t >= a.gt0;               // no upstream gradient before gt0: the tile is not fetched
          mbar_expect_tx(fb, (uint32_t)BR_ARR * ((with_hp ? 3u : 2u) + (with_g ? 1u : 0u)));
          if (with_g) { if (g_time_outer) tma_load_3d(unit(0), &maps.g, 0, first, t - a.gt0, fb); else tma_load_3d(unit(0), &maps.g, 0, t - a.gt0, first, fb); }
          tma_load_3d(unit(1), &maps.z, 0, first, t, fb);
          tma_load_3d(unit(2), &maps.c, 0, first, t, fb);
          if (t > 0) {
            if (hs_time_outer) tma_load_3d(unit(3), &maps.hs, 0, first, t - 1, fb); else tma_load_3d(unit(3), &maps.hs, 0, t - 1, first, fb);
          } else if (with_hp) {
            tma_load_3d(unit(3), &maps.h0, 0, first, 0, fb);
          }
        }
      }
    }
    __syncwarp();
  } else {
    // =========================== epilogue warps ===================================================
    const int ew = warp, quad = warp & 3, es = (ew >> 2) & 1, rh = ew >> 3;
    const int n = quad * 32 + lane;
    const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
    // U -> tensor memory: A[m = k][kk = n] = U_canonical[k][n]; this thread's row m = n (lane), the four warps of
    // a quadrant take 32 columns each
    {
      const int part = ew >> 2;
      const bool hi_layout = a.layout == FGRNN_LAYOUT_HI;
      float uv[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int kk = part * 32 + j;
        uv[j] = hi_layout ? __ldg(a.U + (size_t)kk * BR_H + n) : __ldg(a.U + (size_t)n * BR_H + kk);
      }
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t hi[8], mid[8], lo[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) split3_bf16(uv[c * 16 + 2 * j], uv[c * 16 + 2 * j + 1], hi[j], mid[j], lo[j]);
        tmem_st8(tmem + lane_base + BR_TM_U_HI + part * 16 + c * 8, hi);
        tmem_st8(tmem + lane_base + BR_TM_U_MID + part * 16 + c * 8, mid);
        tmem_st8(tmem + lane_base + BR_TM_U_LO + part * 16 + c * 8, lo);
      }
      tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();

    const float sz = sigmoid_f(__ldg(a.zeta)), sn = sigmoid_f(__ldg(a.nu));
    const float2 sz2 = make_float2(sz, sz), sn2 = make_float2(sn, sn), msz2 = make_float2(-sz, -sz), one2 = make_float2(1.f, 1.f);
    const int first_row = row0 + es * BR_NS + rh * RPT;
    const int rows_left = d.B - first_row;
    unsigned char* hop = sm + L.op + es * (3 * BR_OPT) + (n >> 3) * ((BR_NS >> 3) * 128) + (rh * NG) * 128 + (n & 7) * 16;
    const uint32_t acc = tmem + lane_base + BR_TM_ACC + es * BR_TM_ACC_PER_TILE + rh * RPT;
    float* dpre_p = a.dpre_ws + (size_t)first_row * BR_H + n;             // + t*B*H per step
    float2 carry[PAIRS];                                                       // z G of the previous (later) step
#pragma unroll
    for (int q = 0; q < PAIRS; ++q) carry[q] = make_float2(0.f, 0.f);
    float2 db_u = make_float2(0.f, 0.f), db_g = db_u, dze = db_u, dnu = db_u;

    for (int it = 0; it <= d.T; ++it) {
      const int t = d.T - 1 - it;
      float2 G[PAIRS];
#pragma unroll
      for (int q = 0; q < PAIRS; ++q) G[q] = carry[q];
      if (it > 0) {                                     // delta += dpre_{t+1} . U^T
        mbar_wait(bar(B_DFULL + es), (it - 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int g = 0; g < NG; ++g) {
          float va[8], vb[8], v1[8], v2[8];
          tmem_ld8(acc + g * 8, va);
          tmem_ld8(acc + BR_NS + g * 8, vb);
          tmem_ld8(acc + 2 * BR_NS + g * 8, v1);
          tmem_ld8(acc + 3 * BR_NS + g * 8, v2);
          tmem_ld_wait();
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float2 corr = __fadd2_rn(make_float2(va[2 * q], va[2 * q + 1]), make_float2(vb[2 * q], vb[2 * q + 1]));
            const float2 mm = __fadd2_rn(__fadd2_rn(corr, make_float2(v2[2 * q], v2[2 * q + 1])), make_float2(v1[2 * q], v1[2 * q + 1]));
            G[g * 4 + q] = __fadd2_rn(G[g * 4 + q], mm);
          }
        }
      }
      if (it == d.T) {                                  // delta after t = 0 is d h0
        if (a.d_h0) {
#pragma unroll
          for (int j = 0; j < RPT; ++j)
            if (j < rows_left) a.d_h0[(size_t)(first_row + j) * BR_H + n] = (j & 1) ? G[j >> 1].y : G[j >> 1].x;
        }
        break;
      }
      const int qi = it * BR_NT + es, sb = qi % BR_NB;
      mbar_wait_poll(bar(B_SFULL + sb), (qi / BR_NB) & 1);
      const int u0 = (4 * qi) % BR_RU;
      const float* ring = reinterpret_cast<const float*>(sm + L.stage) + (rh * RPT) * BR_H + n;
      const float* tg = ring + u0 * (BR_ARR / 4);
      const float* tz = ring + ((u0 + 1) % BR_RU) * (BR_ARR / 4);
      const float* tcc = ring + ((u0 + 2) % BR_RU) * (BR_ARR / 4);
      const float* th = ring + ((u0 + 3) % BR_RU) * (BR_ARR / 4);
      const bool hp_zero = t == 0 && hp_zero_at_t0, g_zero = t < a.gt0;
      uint32_t hi[PAIRS], mid[PAIRS], lo[PAIRS];
      float2 dp[PAIRS];
#pragma unroll
      for (int q = 0; q < PAIRS; ++q) {
        const int e = (2 * q) * BR_H;
        const float2 g = g_zero ? make_float2(0.f, 0.f) : make_float2(tg[e], tg[e + BR_H]);
        const float2 z = make_float2(tz[e], tz[e + BR_H]);
        const float2 c = make_float2(tcc[e], tcc[e + BR_H]);
        const float2 hp = hp_zero ? make_float2(0.f, 0.f) : make_float2(th[e], th[e + BR_H]);
        const float2 Gq = __fadd2_rn(G[q], g);
        const float2 w = __fadd2_rn(one2, make_float2(-z.x, -z.y));                       // 1 - z
        const float2 cG = __fmul2_rn(c, Gq);
        const float2 u = __ffma2_rn(make_float2(-c.x, -c.y), c, one2);                     // 1 - c^2
        const float2 dc = __fmul2_rn(__fmul2_rn(__ffma2_rn(sz2, w, sn2), u), Gq);          // cu:112
        const float2 zw = __fmul2_rn(z, w);                                                // sigmoid'(.) on the output
        const float2 dz = __fmul2_rn(__fmul2_rn(__ffma2_rn(msz2, c, hp), zw), Gq);         // cu:113
        dp[q] = __fadd2_rn(dc, dz);
        carry[q] = __fmul2_rn(z, Gq);                                                      // cu:110 d_old_h = z g
        db_u = __fadd2_rn(db_u, dc); db_g = __fadd2_rn(db_g, dz);
        dze = __ffma2_rn(w, cG, dze);                                                      // cu:116 (sigmoid' applied at the end)
        dnu = __fadd2_rn(dnu, cG);                                                         // cu:117
        split3_bf16(dp[q].x, dp[q].y, hi[q], mid[q], lo[q]);
      }
#pragma unroll
      for (int g = 0; g < NG; ++g) {
        *reinterpret_cast<uint4*>(hop + g * 128) = make_uint4(hi[4 * g], hi[4 * g + 1], hi[4 * g + 2], hi[4 * g + 3]);
        *reinterpret_cast<uint4*>(hop + BR_OPT + g * 128) = make_uint4(mid[4 * g], mid[4 * g + 1], mid[4 * g + 2], mid[4 * g + 3]);
        *reinterpret_cast<uint4*>(hop + 2 * BR_OPT + g * 128) = make_uint4(lo[4 * g], lo[4 * g + 1], lo[4 * g + 2], lo[4 * g + 3]);
      }
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) { mbar_arrive(bar(B_HREADY + es)); mbar_arrive(bar(B_SEMPTY + sb)); }
      float* dst = dpre_p + (size_t)t * d.B * BR_H;
#pragma unroll
      for (int q = 0; q < PAIRS; ++q) {
        if (2 * q < rows_left) dst[(2 * q) * BR_H] = dp[q].x;
        if (2 * q + 1 < rows_left) dst[(2 * q + 1) * BR_H] = dp[q].y;
      }
      if (progress && (((it + 1) & (BR_PUBLISH - 1)) == 0 || it == d.T - 1)) {
        // the contraction CTAs follow this kernel: publish "this warp's part of dpre is in memory up to this step".  A
        // gpu-scope fence under memory load costs about a microsecond (measured: one per step made the recurrence 70 %
        // slower), so the progress is published every BR_PUBLISH steps only.
        __threadfence();
        __syncwarp();
        if (lane == 0) asm volatile("st.relaxed.gpu.global.s32 [%0], %1;" ::"l"(progress + blockIdx.x * BR_EPI_WARPS + ew), "r"(it + 1) : "memory");
      }
    }

    // per-CTA partial row: d_bias_gate | d_bias_update | d_zeta (raw) | d_nu (raw); fixed summation order
    asm volatile("bar.sync 1, %0;" ::"n"(BR_EPI_WARPS * 32) : "memory");   // the ring is free: reuse it as scratch
    float* scr = reinterpret_cast<float*>(sm + L.stage);                    // [4 contributors][2][128] | [16 warps][2]
    const int contrib = es * 2 + rh;
    scr[(contrib * 2 + 0) * BR_H + n] = db_g.x + db_g.y;
    scr[(contrib * 2 + 1) * BR_H + n] = db_u.x + db_u.y;
    float vz = dze.x + dze.y, vn = dnu.x + dnu.y;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { vz += __shfl_xor_sync(0xffffffffu, vz, o); vn += __shfl_xor_sync(0xffffffffu, vn, o); }
    if (lane == 0) { scr[8 * BR_H + ew * 2] = vz; scr[8 * BR_H + ew * 2 + 1] = vn; }
    asm volatile("bar.sync 1, %0;" ::"n"(BR_EPI_WARPS * 32) : "memory");
    float* outp = a.rec_partial + (size_t)blockIdx.x * (2 * BR_H + 2);
    if (ew < 8) {                                        // 256 threads: one bias entry each
      const int which = ew >> 2, nn = (ew & 3) * 32 + lane;
      float s = 0.f;
#pragma unroll
      for (int cidx = 0; cidx < 4; ++cidx) s += scr[(cidx * 2 + which) * BR_H + nn];
      outp[which * BR_H + nn] = s;
    } else if (ew == 8 && lane < 2) {
      float s = 0.f;
      for (int w = 0; w < BR_EPI_WARPS; ++w) s += scr[8 * BR_H + w * 2 + lane];
      outp[2 * BR_H + lane] = s;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == BR_W_MMA) tmem_dealloc(tmem, 512);
}
};   // struct BrK

template <int NS>
__global__ void __launch_bounds__(BR_THREADS, 1) tc_bwd_rec_kernel(const SmemBwdArgs a, const __grid_constant__ BrMaps maps, const int g_time_outer, const int hs_time_outer) {
  BrK<NS>::run(a, maps, g_time_outer, hs_time_outer);
}

// The reverse recurrence keeps one CTA per 32 or 64 batch rows busy (64 of 148 SMs at the data-parallel training batch of
// 2048 rows) and the contraction needs its dpre_t only step by step: ONE launch runs both, recurrence CTAs first in the
// grid (they wait for nobody), contraction CTAs on the SMs left over, following the published progress.
template <int NS>
__global__ void __launch_bounds__(BR_THREADS, 1) tc_bwd_fused_kernel(const SmemBwdArgs a, const __grid_constant__ BrMaps maps, const int g_time_outer,
                                                                     const int hs_time_outer, const TcContractArgs c, int* progress) {
  if ((int)blockIdx.x < c.nrec) {
    BrK<NS>::run(a, maps, g_time_outer, hs_time_outer, progress);
  } else {
    if ((int)(threadIdx.x >> 5) > CT_CONV_WARPS) return;                       // the contraction uses 17 of the 20 warps
    ct_run(c, (int)blockIdx.x - c.nrec, (int)gridDim.x - c.nrec);
  }
}

bool tc_bwd_rec_supports(const Dims& d) {
  return d.rW == 0 && d.rU == 0 && d.H == BR_H && d.gate_nl == FGRNN_NL_SIGMOID && d.update_nl == FGRNN_NL_TANH;
}
// 16-row sub-tiles while the batch fits one wave of 32-row CTAs; FGRNN_TC_BR_NS=16|32 overrides (tests, benchmarks)
static int br_ns_for(int B) {
  int ns = B <= 148 * 32 ? 16 : 32;
  if (tuning(TUNE_TC_BR_NS) == 16) ns = 16; else if (tuning(TUNE_TC_BR_NS) == 32) ns = 32;
  return ns;
}
int tc_bwd_rec_ctas(const Dims& d) { const int rows = br_ns_for(d.B) * BR_NT; return (d.B + rows - 1) / rows; }

template <int BR_NS>
static int launch_tc_bwd_rec_ns(const SmemBwdArgs& a, cudaStream_t stream) {
  using K = BrK<BR_NS>;
  const Dims& d = a.d;
  BrMaps maps;
  int g_to = 0, hs_to = 0, dummy = 0, rc;
  if ((rc = make_row_tile_map(&maps.g, a.grad_h, false, BR_H, d.B, d.T - a.gt0, a.gsb, a.gst, BR_NS, &g_to))) return rc;
  if ((rc = make_row_tile_map(&maps.z, a.z_s, false, BR_H, d.B, d.T, BR_H, (int64_t)d.B * BR_H, BR_NS, &dummy))) return rc;
  if ((rc = make_row_tile_map(&maps.c, a.c_s, false, BR_H, d.B, d.T, BR_H, (int64_t)d.B * BR_H, BR_NS, &dummy))) return rc;
  // T == 1 never reads the hidden states (h_{t-1} is h0): any valid pointer keeps the descriptor well formed
  const float* hs = a.hs ? a.hs : a.z_s;
  if ((rc = make_row_tile_map(&maps.hs, hs, false, BR_H, d.B, d.T, a.hs ? a.hsb : BR_H, a.hs ? a.hst : (int64_t)d.B * BR_H, BR_NS, &hs_to))) return rc;
  const float* h0 = a.h0 ? a.h0 : a.z_s;
  if ((rc = make_row_tile_map(&maps.h0, h0, false, BR_H, d.B, 1, BR_H, (int64_t)d.B * BR_H, BR_NS, &dummy))) return rc;
  const BrSmem L = K::br_smem_layout();
  FGRNN_CUDA_TRY(cudaFuncSetAttribute(tc_bwd_rec_kernel<BR_NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total));
  tc_bwd_rec_kernel<BR_NS><<<(d.B + K::BR_ROWS - 1) / K::BR_ROWS, BR_THREADS, L.total, stream>>>(a, maps, g_to, hs_to);
  FGRNN_LAUNCH_CHECK("tc_bwd_rec_kernel");
  return FGRNN_OK;
}

int launch_tc_bwd_rec(const SmemBwdArgs& a, cudaStream_t stream) {
  if (a.d.B <= 0 || a.d.T <= 0) return FGRNN_OK;
  return br_ns_for(a.d.B) == 16 ? launch_tc_bwd_rec_ns<16>(a, stream) : launch_tc_bwd_rec_ns<32>(a, stream);
}

// Contraction CTAs of the fused launch: whatever the recurrence leaves free, 0 = do not fuse (the recurrence fills the GPU,
// or too few SMs would be left to follow it).
int tc_bwd_fused_contract_ctas(const Dims& d) {
  const int left = 148 - tc_bwd_rec_ctas(d);
  return left >= 48 ? left : 0;
}

template <int BR_NS>
static int launch_tc_bwd_fused_ns(const SmemBwdArgs& a, const TcContractLaunch& c, int ncontract, int* progress, cudaStream_t stream) {
  using K = BrK<BR_NS>;
  const Dims& d = a.d;
  BrMaps maps;
  int g_to = 0, hs_to = 0, dummy = 0, rc;
  if ((rc = make_row_tile_map(&maps.g, a.grad_h, false, BR_H, d.B, d.T - a.gt0, a.gsb, a.gst, BR_NS, &g_to))) return rc;
  if ((rc = make_row_tile_map(&maps.z, a.z_s, false, BR_H, d.B, d.T, BR_H, (int64_t)d.B * BR_H, BR_NS, &dummy))) return rc;
  if ((rc = make_row_tile_map(&maps.c, a.c_s, false, BR_H, d.B, d.T, BR_H, (int64_t)d.B * BR_H, BR_NS, &dummy))) return rc;
  const float* hs = a.hs ? a.hs : a.z_s;
  if ((rc = make_row_tile_map(&maps.hs, hs, false, BR_H, d.B, d.T, a.hs ? a.hsb : BR_H, a.hs ? a.hst : (int64_t)d.B * BR_H, BR_NS, &hs_to))) return rc;
  const float* h0 = a.h0 ? a.h0 : a.z_s;
  if ((rc = make_row_tile_map(&maps.h0, h0, false, BR_H, d.B, 1, BR_H, (int64_t)d.B * BR_H, BR_NS, &dummy))) return rc;
  TcContractArgs ca{};
  ca.B = d.B; ca.T = d.T; ca.I = d.I; ca.KI = (d.I + 15) & ~15;
  ca.x_dtype = d.x_dtype;
  ca.x = c.x; ca.xsb = c.xsb; ca.xst = c.xst;
  ca.hs = c.hs; ca.hsb = c.hsb; ca.hst = c.hst; ca.h0 = c.h0; ca.dpre = c.dpre;
  ca.partW = c.partW; ca.partU = c.partU;
  ca.nbblk = (d.B + CT_ROWS - 1) / CT_ROWS;
  ca.nchunk = d.T * ca.nbblk;
  ca.progress = progress; ca.rec_rows = K::BR_ROWS; ca.nrec = (d.B + K::BR_ROWS - 1) / K::BR_ROWS;
  const BrSmem L = K::br_smem_layout();
  const CtSmem CL = ct_smem_layout(ca.KI);
  const int smem = L.total > CL.total ? L.total : CL.total;
  FGRNN_CUDA_TRY(cudaMemsetAsync(progress, 0, sizeof(int) * (size_t)ca.nrec * BR_EPI_WARPS, stream));
  FGRNN_CUDA_TRY(cudaFuncSetAttribute(tc_bwd_fused_kernel<BR_NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  tc_bwd_fused_kernel<BR_NS><<<ca.nrec + ncontract, BR_THREADS, smem, stream>>>(a, maps, g_to, hs_to, ca, progress);
  FGRNN_LAUNCH_CHECK("tc_bwd_fused_kernel");
  return FGRNN_OK;
}

int launch_tc_bwd_fused(const SmemBwdArgs& a, const TcContractLaunch& c, int ncontract, int* progress, cudaStream_t stream) {
  if (a.d.B <= 0 || a.d.T <= 0) return FGRNN_OK;
  return br_ns_for(a.d.B) == 16 ? launch_tc_bwd_fused_ns<16>(a, c, ncontract, progress, stream)
                                : launch_tc_bwd_fused_ns<32>(a, c, ncontract, progress, stream);
}

}  // namespace fgrnn
