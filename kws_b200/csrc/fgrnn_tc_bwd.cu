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

__global__ void __launch_bounds__(CT_THREADS, 1) tc_contract_kernel(const TcContractArgs a) {
  extern __shared__ __align__(128) unsigned char sm[];
  const CtSmem L = ct_smem_layout(a.KI);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + L.bars);      // [0,1] full, [2,3] empty, [4] done
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  auto bar = [&](int i) { return smem_u32(&bars[i]); };
  const bool want_u = a.partU != nullptr, want_w = a.partW != nullptr;

  if (warp == CT_CONV_WARPS) tmem_alloc(smem_u32(&tmem_base_s), 512);
  if (tid == 0) {
    mbar_init(bar(0), CT_CONV_WARPS); mbar_init(bar(1), CT_CONV_WARPS);
    mbar_init(bar(2), 1); mbar_init(bar(3), 1); mbar_init(bar(4), 1);
    fence_mbar_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (tmem_base_s != 0u) __trap();               // the CTA owns all 512 columns
  constexpr uint32_t tmem = 0u;
  const int my_chunks = a.nchunk > (int)blockIdx.x ? (a.nchunk - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

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
      const int chunk = (int)blockIdx.x + j * (int)gridDim.x;
      const int t = chunk / a.nbblk, b0 = (chunk - t * a.nbblk) * CT_ROWS;
      const int b = j & 1;
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
              const float* src = a.dpre + ((size_t)t * a.B + row) * CT_H + col;
              const float4 p0 = __ldg(reinterpret_cast<const float4*>(src)), p1 = __ldg(reinterpret_cast<const float4*>(src) + 1);
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
      float* dst = a.partU + ((size_t)blockIdx.x * CT_H + m) * CT_H + part * 32;      // row m = k, columns n
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
          if (c0 + q < a.I) a.partW[((size_t)blockIdx.x * a.I + c0 + q) * CT_H + m] = vm[q] + vc[q];
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == CT_CONV_WARPS) tmem_dealloc(tmem, 512);
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

}  // namespace fgrnn
