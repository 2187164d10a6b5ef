// tcgen05 family, wide shapes (FGRNN_PATH_TCGEN05 when the fused kernel of fgrnn_tc.cu does not apply):
// the reference's default model (trainingConfig.py:9-30: 2 layers, hidden 256 / 128, 64 delta-MFCC features) and its
// BatchNorm variant.  Two kernels, as north-star items (1) and (2) prescribe:
//
//  tc_xw_kernel    -- the input projection hoisted out of the recurrence: WX[t][b][:] = x_t[b] . W for ALL T steps as
//                     one dense batched contraction on the tensor cores (rnn.py:278), fed by TMA.  W^T is the
//                     stationary A operand in tensor memory (fp16 hi/lo pairs, M = 128 hidden units per CTA column),
//                     x tiles stream through shared memory as the K-major B operand (TMA -> fp16 hi/lo split by
//                     converter warps), three products per k-step, hi.hi chains of <= 8 MMAs (accumulate-truncation,
//                     see fgrnn_tc.cu).  I up to 256, H = 128 or 256 (gridDim.y = H / 128).  HBM bound.
//  tc_wx_fwd_kernel -- the persistent recurrence  pre_t = WX_t + h_{t-1}.U  with the fused gate update (rnn.py:280-295):
//                     U^T stationary in tensor memory, h_{t-1} as the MN-major B operand written by the epilogue warps,
//                     WX_t tiles streamed by TMA into a shared-memory ring and added in the epilogue.
//                     H = 128: one CTA per 64 batch rows.  H = 256: a CTA PAIR (thread-block cluster of 2) per 64 rows --
//                     each CTA owns 128 hidden units (its U^T slice fills 256 of its 512 TMEM columns) and, every step,
//                     writes its half of the fp16 h_t operand tile into its own AND its partner's shared memory
//                     (st.shared::cluster), then arrives on both CTAs' barriers; operand tiles are double buffered by
//                     step parity so that the partner's MMAs of step t-1 are provably finished before they are overwritten.
//                     Gate: sigmoid or tanh; update: tanh; optional per-unit affine maps on the two pre-activations
//                     (eval-mode BatchNorm of rnn.py:316-452 folded: gate_scale / update_scale).
#include <cuda.h>

#include "fgrnn_kernels.cuh"
#include "fgrnn_tc_common.cuh"

namespace fgrnn {

constexpr int WX_HC = 128;                       // hidden units per CTA = UMMA M
constexpr int WX_NT = 2;                         // sub-tiles per CTA
constexpr int WX_EPI_WARPS = 16;

// ---- cluster / distributed-shared-memory PTX ----------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local_smem, uint32_t rank) {
  uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem), "r"(rank)); return r;
}
// Asynchronous 16-byte store into the partner CTA's shared memory; its completion is counted in bytes on the partner's
// mbarrier (complete_tx), so the hand-off needs no cluster-scope release: a `mbarrier.arrive.release.cluster` (and a full
// fence.proxy.async) compiles to MEMBAR.ALL.GPU, which waits for every global store of h_t / z_t / c_t still in flight --
// measured 1.6 us per step on the critical chain of the first version of the pair kernel.
__device__ __forceinline__ void st_async_v4(uint32_t cluster_addr, uint4 v, uint32_t cluster_mbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1,%2,%3,%4}, [%5];"
               ::"r"(cluster_addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"(cluster_mbar) : "memory");
}

// =====================================================================================================================
// tc_xw_kernel: WX = x . W
// =====================================================================================================================
// Pipeline unit = (tile, slab): a tile is 64 batch rows of one time step, a slab is 64 input features of it (one TMA box
// of 16 rows x 64 features per converter warp, four k-steps of MMAs).  Every unit has its own accumulator pair
// C (lo products, 8 MMAs) | M (hi.hi, 4 MMAs), double buffered per sub-tile; the epilogue drains it and adds the slabs of
// a tile in fp32 registers, round to nearest -- chains stay at <= 4 hi.hi MMAs whatever I is (the tensor core truncates
// every addend of an accumulate toward zero, fgrnn_tc.cu), which keeps x.W within 0.2 of the tolerance of an fp64 product.
struct XwArgs {
  Dims d;
  int layout;                      // weight layout of W
  const float* W;
  float* wx;                       // [T][B][H] contiguous
  int KI;                          // I rounded up to a multiple of 16 (<= 256)
  int KSW;                         // k width of a slab's operand tile: min(KI, 64)
  int BOXI;                        // features per TMA box = row pitch of the raw tile: I (one box per 16-row group and tile)
  int nslab;                       // ceil(KI / 64)
  int x_time_outer;
  int ntiles, nrb;                 // tiles = T * nrb, nrb = ceil(B / 64)
};

constexpr int XW_NS = 32, XW_ROWS = XW_NS * WX_NT;       // 64 rows of one time step per tile
constexpr int XW_CONV_WARPS = 4, XW_CONV_ROWS = XW_ROWS / XW_CONV_WARPS;
constexpr int XW_MMA_WARPS = 2;                           // role 0: lo products -> C, role 1: hi.hi -> M
constexpr int XW_THREADS = 32 * (WX_EPI_WARPS + XW_CONV_WARPS + XW_MMA_WARPS);
constexpr int XW_XBUF = 4;
__host__ __device__ inline int xw_raw_stages(int nslab) { return nslab > 1 ? 2 : 4; }      // raw tiles in flight per converter warp
constexpr int XW_TM_ACC = 256;                            // W_hi | W_lo take up to 2 x 128 columns (KI <= 256)

struct XwSmem { int x_op, raw, bars, misc, total; int x_tile_bytes, raw_stage_bytes; };
__host__ __device__ inline XwSmem xw_smem_layout(int BOXI, int KSW, int esz, int nslab) {
  XwSmem L;
  L.x_tile_bytes = XW_NS * KSW * 2;
  L.raw_stage_bytes = XW_CONV_ROWS * BOXI * esz;
  L.x_op = 0;                                                     // [XBUF][NT][hi|lo][x_tile_bytes]
  L.raw = (XW_XBUF * WX_NT * 2 * L.x_tile_bytes + 127) & ~127;    // [CONV_WARPS][RAW_STAGES][raw_stage_bytes]
  L.bars = (L.raw + XW_CONV_WARPS * xw_raw_stages(nslab) * L.raw_stage_bytes + 15) & ~15;
  L.misc = L.bars + 32 * 8;
  L.total = L.misc + 128;
  return L;
}

static __device__ __forceinline__ uint64_t xw_desc_kmajor(uint32_t smem_addr, int KSW) {
  const uint64_t lbo = 128 >> 4, sbo = (uint64_t)((KSW >> 3) * 128) >> 4;
  return (uint64_t)((smem_addr >> 4) & 0x3fff) | (lbo << 16) | (sbo << 32) | (1ull << 46);
}

__global__ void __launch_bounds__(XW_THREADS, 1) tc_xw_kernel(const XwArgs a, const __grid_constant__ CUtensorMap xmap) {
  extern __shared__ __align__(128) unsigned char sm[];
  const Dims d = a.d;
  const int I = d.I, KI = a.KI, H = d.H;
  const int esz = d.x_dtype == FGRNN_BF16 ? 2 : 4;
  const XwSmem L = xw_smem_layout(a.BOXI, a.KSW, esz, a.nslab);
  const int RS = xw_raw_stages(a.nslab);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + L.bars);
  uint32_t* tmem_base_s = reinterpret_cast<uint32_t*>(sm + L.misc);
  float* red_s = reinterpret_cast<float*>(sm + L.misc + 16);
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
  auto bar = [&](int i) { return smem_u32(&bars[i]); };
  const int B_DEMPTY = 0, B_DFULL = 4, B_XFULL = 8, B_XEMPTY = 12, B_RAWFULL = 16;    // D*: [sub-tile][buffer]; RAWFULL: [conv warp][stage]
  constexpr int W_CONV0 = WX_EPI_WARPS, W_MMA = WX_EPI_WARPS + XW_CONV_WARPS;
  const int unit0 = blockIdx.y * WX_HC;
  const int nkx = KI >> 4, nslab = a.nslab;
  const int my_tiles = a.ntiles > (int)blockIdx.x ? (a.ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  const int my_units = my_tiles * nslab;                       // pipeline units of this CTA

  if (warp == W_MMA) tmem_alloc(smem_u32(tmem_base_s), 512);
  if (tid == 0) {
    for (int s = 0; s < 2 * WX_NT; ++s) { mbar_init(bar(B_DEMPTY + s), WX_EPI_WARPS / WX_NT); mbar_init(bar(B_DFULL + s), XW_MMA_WARPS); }
    for (int b = 0; b < XW_XBUF; ++b) { mbar_init(bar(B_XFULL + b), XW_CONV_WARPS); mbar_init(bar(B_XEMPTY + b), XW_MMA_WARPS); }
    for (int st = 0; st < XW_CONV_WARPS * 4; ++st) mbar_init(bar(B_RAWFULL + st), 1);
    fence_mbar_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (*tmem_base_s != 0u) __trap();
  constexpr uint32_t tmem = 0u;
  const uint32_t TM_W_HI = 0, TM_W_LO = (uint32_t)nkx * 8;

  if (warp >= W_MMA) {
    // =========================== MMA issuers ======================================================
    // role 0 -> C: the lo products (W_lo.x_hi, W_hi.x_lo) of the slab's k-steps; role 1 -> M: hi.hi
    const int role = warp - W_MMA;
    const bool leader = elect_one();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const bool x_has_lo = d.x_dtype != FGRNN_BF16;
    const uint32_t idesc = (1u << 4) | ((uint32_t)(XW_NS >> 3) << 17) | ((uint32_t)(WX_HC >> 4) << 24);
    const uint64_t dX0 = xw_desc_kmajor(smem_u32(sm + L.x_op), a.KSW);
    const uint32_t xlo_step = (uint32_t)L.x_tile_bytes >> 4, xtile_step = 2 * xlo_step, xbuf_step = WX_NT * xtile_step;
    for (int u = 0; u < my_units; ++u) {
      const int sl = u % nslab, xb = u % XW_XBUF, db = u & 1;
      mbar_wait(bar(B_XFULL + xb), (u / XW_XBUF) & 1);
      const int ks0 = sl * 4, ks1 = min(nkx, ks0 + 4);
#pragma unroll
      for (int s = 0; s < WX_NT; ++s) {
        const uint64_t dXhi = dX0 + (uint64_t)(xb * xbuf_step + s * xtile_step), dXlo = dXhi + xlo_step;
        const uint32_t acc = tmem + XW_TM_ACC + (s * 2 + db) * 2 * XW_NS + role * XW_NS;
        mbar_wait(bar(B_DEMPTY + s * 2 + db), ((u >> 1) & 1) ^ 1);      // the epilogue has drained unit u-2 of this sub-tile
        tc_fence_after();
        if (leader) {
          for (int ks = ks0; ks < ks1; ++ks) {
            const uint32_t kd = (uint32_t)(ks - ks0) * 16;              // descriptor advance inside the slab tile
            if (role == 0) {
              umma_ts1(acc, tmem + TM_W_LO + ks * 8, dXhi + kd, idesc, ks > ks0);
              if (x_has_lo) umma_ts1(acc, tmem + TM_W_HI + ks * 8, dXlo + kd, idesc, 1);
            } else {
              umma_ts1(acc, tmem + TM_W_HI + ks * 8, dXhi + kd, idesc, ks > ks0);
            }
          }
          umma_commit1(bar(B_DFULL + s * 2 + db));
        }
        __syncwarp();
      }
      if (leader) umma_commit1(bar(B_XEMPTY + xb));
      __syncwarp();
    }
  } else if (warp >= W_CONV0) {
    // =========================== x path: TMA -> fp16 hi/lo split -> K-major operand tiles =============
    const int cw = warp - W_CONV0;
    const uint32_t raw_bytes = (uint32_t)L.raw_stage_bytes;
    unsigned char* raw_base = sm + L.raw + cw * RS * L.raw_stage_bytes;
    // one TMA box per tile and converter warp: 16 rows x ALL I features (1 KB contiguous per row at I = 256; the first
    // version fetched 64-feature slabs, 256-byte pieces of rows 101 KB apart, and read HBM at 2.6 TB/s)
    auto issue_tma = [&](int j) {
      const int tile = (int)blockIdx.x + j * (int)gridDim.x;
      const int t = tile / a.nrb, r0 = (tile - t * a.nrb) * XW_ROWS + cw * XW_CONV_ROWS;
      const int st = j % RS;
      const uint32_t fb = bar(B_RAWFULL + cw * 4 + st);
      mbar_expect_tx(fb, raw_bytes);
      if (a.x_time_outer) tma_load_3d(smem_u32(raw_base + st * L.raw_stage_bytes), &xmap, 0, r0, t, fb);
      else tma_load_3d(smem_u32(raw_base + st * L.raw_stage_bytes), &xmap, 0, t, r0, fb);
    };
    if (lane == 0)
      for (int j = 0; j < RS && j < my_tiles; ++j) issue_tma(j);
    tc_fence_before();
    __syncthreads();
    // 16 rows x 8 chunks of 8 features = 128 tasks per slab, 4 per lane
    const int nch_slab = a.KSW >> 3, boxi = a.BOXI;
    const uint32_t xbuf_bytes = (uint32_t)(WX_NT * 2 * L.x_tile_bytes);
    for (int u = 0; u < my_units; ++u) {
      const int xb = u % XW_XBUF;
      const int j = u / nslab, sl = u - j * nslab, st = j % RS;
      const int nch_here = min(nch_slab, (KI >> 3) - sl * 8);          // chunks of this slab that exist in K
      if (sl == 0) mbar_wait(bar(B_RAWFULL + cw * 4 + st), (j / RS) & 1);
      mbar_wait(bar(B_XEMPTY + xb), ((u / XW_XBUF) & 1) ^ 1);
      const unsigned char* raw = raw_base + st * L.raw_stage_bytes;
      unsigned char* xdst = sm + L.x_op + xb * xbuf_bytes;
      // 16 rows x 8 chunks of 8 features = 128 tasks, 4 per lane, on a DIAGONAL: the 8 lanes of a quarter warp take 8
      // different rows AND 8 different chunks, so that both the 32-byte reads of the raw tile (row stride 256 B; the two
      // 16-byte halves are read in opposite order by lanes 0-3 / 4-7) and the 16-byte writes into the K-major operand tile
      // (8 consecutive rows of one chunk = 128 B) are free of bank conflicts (the row-major mapping cost 8 wavefronts per
      // STS.128: 56 M excess wavefronts on the default model's second layer, 40 % of the kernel).
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        const int jj = lane & 7, row = (it & 1) * 8 + jj, ch = (jj + (lane >> 3) + 4 * (it >> 1)) & 7;
        if (ch < nch_slab) {
          float v[8];
          if (ch < nch_here && (sl * 8 + ch) * 8 < boxi) {           // inside the row (K padding chunks are zeros)
            const unsigned char* src = raw + (size_t)row * boxi * esz + (sl * 8 + ch) * 8 * esz;
            if (esz == 4) {
              const int sw = (jj >> 2) & 1;
              const float4 pa = *reinterpret_cast<const float4*>(src + sw * 16), pb = *reinterpret_cast<const float4*>(src + (sw ^ 1) * 16);
              const float4 p0 = sw ? pb : pa, p1 = sw ? pa : pb;
              v[0] = p0.x; v[1] = p0.y; v[2] = p0.z; v[3] = p0.w; v[4] = p1.x; v[5] = p1.y; v[6] = p1.z; v[7] = p1.w;
            } else {
              const uint4 p = *reinterpret_cast<const uint4*>(src);
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
          const int R = cw * XW_CONV_ROWS + row, sidx = R / XW_NS, r = R - sidx * XW_NS;
          unsigned char* dst = xdst + (sidx * 2) * L.x_tile_bytes + (r >> 3) * (nch_slab * 128) + ch * 128 + (r & 7) * 16;
          *reinterpret_cast<uint4*>(dst) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
          *reinterpret_cast<uint4*>(dst + L.x_tile_bytes) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(bar(B_XFULL + xb));
        if (sl == nslab - 1 && j + RS < my_tiles) issue_tma(j + RS);       // this warp is done with the raw tile of tile j
      }
      __syncwarp();
    }
  } else {
    // =========================== epilogue warps: W -> TMEM, then accumulators -> WX ====================
    const int ew = warp, quad = warp & 3, es = (ew >> 2) & 1, rh = ew >> 3, part = ew >> 2;
    const int n = quad * 32 + lane;
    const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
    const bool hi_layout = a.layout == FGRNN_LAYOUT_HI;
    // max |W| over this CTA's 128 columns -> power-of-two scale
    float mw = 0.f;
    for (int k = part; k < I; k += 4) mw = fmaxf(mw, fabsf(hi_layout ? __ldg(a.W + (size_t)(unit0 + n) * I + k) : __ldg(a.W + (size_t)k * H + unit0 + n)));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mw = fmaxf(mw, __shfl_xor_sync(0xffffffffu, mw, o));
    if (lane == 0) red_s[ew] = mw;
    asm volatile("bar.sync 1, %0;" ::"n"(WX_EPI_WARPS * 32) : "memory");
#pragma unroll
    for (int w = 0; w < WX_EPI_WARPS; ++w) mw = fmaxf(mw, red_s[w]);
    int S = 40;
    if (mw > 0.f) S = min(S, (int)floorf(log2f(30000.f / mw)));
    S = max(S, -14);
    const float scale_w = exp2f((float)S), unscale = exp2f((float)-S);
    // the four warps of a lane quadrant take the 16-k blocks part, part + 4, ...
    for (int kb = part; kb < nkx; kb += 4) {
      float wv[16];
#pragma unroll
      for (int jj = 0; jj < 16; ++jj) {
        const int k = kb * 16 + jj;
        wv[jj] = k < I ? (hi_layout ? __ldg(a.W + (size_t)(unit0 + n) * I + k) : __ldg(a.W + (size_t)k * H + unit0 + n)) : 0.f;
      }
      uint32_t hi[8], lo[8];
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) split2(wv[2 * jj], wv[2 * jj + 1], scale_w, hi[jj], lo[jj]);
      tmem_st8(tmem + lane_base + TM_W_HI + kb * 8, hi);
      tmem_st8(tmem + lane_base + TM_W_LO + kb * 8, lo);
    }
    tmem_st_wait();
    tc_fence_before();
    __syncthreads();

    const uint32_t acc0 = tmem + lane_base + XW_TM_ACC + es * 4 * XW_NS + rh * 16;
    int u = 0;
    int t = (int)blockIdx.x / a.nrb, rb = (int)blockIdx.x - t * a.nrb;       // tile = t * nrb + rb, advanced without divisions
    const int sub_row = es * XW_NS + rh * 16;
    const size_t Hs = (size_t)H;
    for (int j = 0; j < my_tiles; ++j) {
      const int row_first = rb * XW_ROWS + sub_row;
      float r[16];
#pragma unroll
      for (int q = 0; q < 16; ++q) r[q] = 0.f;
      for (int sl = 0; sl < nslab; ++sl, ++u) {
        const int db = u & 1;
        const uint32_t acc = acc0 + db * 2 * XW_NS;
        mbar_wait(bar(B_DFULL + es * 2 + db), (u >> 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          float vc[8], vm[8];
          tmem_ld8(acc + g * 8, vc);
          tmem_ld8(acc + XW_NS + g * 8, vm);
          tmem_ld_wait();
#pragma unroll
          for (int q = 0; q < 8; ++q) r[g * 8 + q] += vc[q] + vm[q];
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(B_DEMPTY + es * 2 + db));
      }
      float* dst = a.wx + ((size_t)t * d.B + row_first) * Hs + unit0 + n;
      if (row_first + 16 <= d.B) {
#pragma unroll
        for (int q = 0; q < 16; ++q) dst[q * Hs] = r[q] * unscale;
      } else {
#pragma unroll
        for (int q = 0; q < 16; ++q)
          if (row_first + q < d.B) dst[q * Hs] = r[q] * unscale;
      }
      rb += (int)gridDim.x;
      while (rb >= a.nrb) { rb -= a.nrb; ++t; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == W_MMA) tmem_dealloc(tmem, 512);
}

// =====================================================================================================================
// tc_wx_fwd_kernel: pre_t = WX_t + h_{t-1}.U, gate update, all T steps
// =====================================================================================================================
struct WxFwdArgs {
  Dims d;
  int layout;
  const float* U;
  const float *bias_gate, *bias_update, *zeta, *nu;
  const float *gate_scale, *update_scale;       // optional per-unit factors on the two pre-activations (null = 1)
  const float* h0;
  float* out; int64_t osb, ost;
  float* h_last; float* save_z; float* save_c;
};

template <int V> struct IntTag { static constexpr int value = V; };
template <bool V> struct BoolTag { static constexpr bool value = V; };
constexpr int WXF_MMA_WARPS = 3;
constexpr int WXF_THREADS = 32 * (WX_EPI_WARPS + 1 + WXF_MMA_WARPS);       // 640
struct WxfSmem { int h_op, ring, bars, misc, total; int h_tile_bytes, ring_stage_bytes, stages, nbuf; };

template <int NS, bool PAIR>
struct TcWxFwd {
  static constexpr int KH = PAIR ? 16 : 8;                 // k-steps of h.U
  static constexpr int HK = KH * 16;                       // total hidden units = K
  static constexpr int ROWS = NS * WX_NT;
  static constexpr int RPT = NS / 2, PAIRS = RPT / 2, NG = RPT / 8;
  static constexpr uint32_t TM_U_HI = 0, TM_U_LO = KH * 8, TM_ACC = 2 * KH * 8, TM_ACC_PER_TILE = 4 * NS;
  static constexpr int TMEM_COLS = 512;
  static constexpr uint32_t IDESC_H = (1u << 4) | (1u << 16) | ((uint32_t)(NS >> 3) << 17) | ((uint32_t)(WX_HC >> 4) << 24);
  static constexpr uint32_t H_KSTEP = (2 * (NS >> 3) * 128) >> 4;
  static_assert(TM_ACC + WX_NT * TM_ACC_PER_TILE <= 512, "tensor memory budget");

  static __host__ __device__ inline WxfSmem smem_layout() {
    WxfSmem L;
    L.h_tile_bytes = HK * NS * 2;
    L.nbuf = PAIR ? 2 : 1;
    L.ring_stage_bytes = NS * WX_HC * 4;
    L.stages = PAIR ? 2 : 4;
    L.h_op = 0;                                                  // [nbuf][NT][hi|lo][h_tile_bytes]
    L.ring = L.nbuf * WX_NT * 2 * L.h_tile_bytes;                // [NT][stages][ring_stage_bytes]
    L.bars = L.ring + WX_NT * L.stages * L.ring_stage_bytes;
    L.misc = L.bars + 32 * 8;
    L.total = L.misc + 256;
    return L;
  }
  static __device__ __forceinline__ uint64_t desc_mn(uint32_t smem_addr) {
    const uint64_t sbo = 128 >> 4, lbo = (uint64_t)((NS >> 3) * 128) >> 4;
    return (uint64_t)((smem_addr >> 4) & 0x3fff) | (lbo << 16) | (sbo << 32) | (1ull << 46);
  }

  struct EpiK { float2 un, kg, cg, ku, cu, msz, szn; float sg, bg, su, bu, tmin; };

  static __device__ __forceinline__ void run(const WxFwdArgs& a, const CUtensorMap& wxmap) {
    extern __shared__ __align__(128) unsigned char sm[];
    const Dims d = a.d;
    const WxfSmem L = smem_layout();
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + L.bars);
    uint32_t* tmem_base_s = reinterpret_cast<uint32_t*>(sm + L.misc);
    float* red_s = reinterpret_cast<float*>(sm + L.misc + 16);
    const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
    const uint32_t rank = PAIR ? cluster_ctarank() : 0u, peer = rank ^ 1u;
    const int row0 = (PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x) * ROWS;
    const int unit0 = (int)rank * WX_HC;
    auto bar = [&](int i) { return smem_u32(&bars[i]); };
    // HREADY: [sub-tile][step parity] -- the pair kernel uses one barrier per operand buffer (a step's 16 KB from the
    // partner are counted on the barrier of the buffer they land in, so traffic of adjacent steps can never mix);
    // the single-CTA kernel uses [sub-tile][0] only.  WX*: [sub-tile][stage]
    const int B_HREADY = 0, B_DFULL = 4, B_WXFULL = 6, B_WXEMPTY = 14;
    auto hr_bar = [&](int s_, int t_) { return bar(B_HREADY + s_ * 2 + (PAIR ? (t_ & 1) : 0)); };
    auto hr_parity = [&](int t_) { return (uint32_t)(PAIR ? (t_ >> 1) & 1 : t_ & 1); };
    constexpr int W_PROD = WX_EPI_WARPS, W_MMA = WX_EPI_WARPS + 1;
    constexpr int EPI_PER_TILE = WX_EPI_WARPS / WX_NT;

    if (warp == W_MMA) tmem_alloc(smem_u32(tmem_base_s), TMEM_COLS);
    if (tid == 0) {
      for (int s = 0; s < WX_NT; ++s) {
        mbar_init(bar(B_HREADY + s * 2), EPI_PER_TILE);
        mbar_init(bar(B_HREADY + s * 2 + 1), EPI_PER_TILE);
        mbar_init(bar(B_DFULL + s), WXF_MMA_WARPS);
        for (int st = 0; st < L.stages; ++st) { mbar_init(bar(B_WXFULL + s * 4 + st), 1); mbar_init(bar(B_WXEMPTY + s * 4 + st), EPI_PER_TILE); }
      }
      fence_mbar_init();
    }
    tc_fence_before();
    if (PAIR) cluster_sync_all(); else __syncthreads();        // barriers of BOTH CTAs are initialised before anyone arrives remotely
    tc_fence_after();
    if (*tmem_base_s != 0u) __trap();
    constexpr uint32_t tmem = 0u;

    if (warp >= W_MMA) {
      // =========================== MMA issuers ====================================================
      const int role = warp - W_MMA;
      const bool leader = elect_one();
      tc_fence_before();
      __syncthreads();
      tc_fence_after();
      const uint64_t dH0 = desc_mn(smem_u32(sm + L.h_op));
      const uint32_t hlo_step = (uint32_t)L.h_tile_bytes >> 4, htile_step = 2 * hlo_step, hbuf_step = WX_NT * htile_step;
      for (int t = 0; t < d.T; ++t) {
#pragma unroll
        for (int s = 0; s < WX_NT; ++s) {
          const uint64_t dHhi = dH0 + (uint64_t)((PAIR ? (t & 1) : 0) * hbuf_step + s * htile_step), dHlo = dHhi + hlo_step;
          const uint32_t acc = tmem + TM_ACC + s * TM_ACC_PER_TILE;
          mbar_wait(hr_bar(s, t), hr_parity(t));        // own half written by this CTA's epilogue, partner's half landed (tx bytes)
          if (PAIR) fence_proxy_async_smem();           // the partner's st.async data (generic proxy) -> visible to tcgen05.mma
          tc_fence_after();
          if (leader) {
            if (role == 0) {
#pragma unroll
              for (int ks = 0; ks < KH / 2; ++ks) {
                umma_ts1(acc, tmem + TM_U_LO + ks * 8, dHhi + ks * H_KSTEP, IDESC_H, ks > 0);
                umma_ts1(acc, tmem + TM_U_HI + ks * 8, dHlo + ks * H_KSTEP, IDESC_H, 1);
              }
            } else if (role == 1) {
#pragma unroll
              for (int ks = KH / 2; ks < KH; ++ks) {
                umma_ts1(acc + NS, tmem + TM_U_LO + ks * 8, dHhi + ks * H_KSTEP, IDESC_H, ks > KH / 2);
                umma_ts1(acc + NS, tmem + TM_U_HI + ks * 8, dHlo + ks * H_KSTEP, IDESC_H, 1);
              }
            } else {
#pragma unroll
              for (int ks = 0; ks < KH / 2; ++ks) umma_ts1(acc + 2 * NS, tmem + TM_U_HI + ks * 8, dHhi + ks * H_KSTEP, IDESC_H, ks > 0);
#pragma unroll
              for (int ks = KH / 2; ks < KH; ++ks) umma_ts1(acc + 3 * NS, tmem + TM_U_HI + ks * 8, dHhi + ks * H_KSTEP, IDESC_H, ks > KH / 2);
            }
            umma_commit1(bar(B_DFULL + s));
          }
          __syncwarp();
        }
      }
    } else if (warp == W_PROD) {
      // =========================== producer: one TMA box of WX_t per sub-tile step ==================
      tc_fence_before();
      __syncthreads();
      if (lane == 0) {
        for (int t = 0; t < d.T; ++t) {
          const int st = t % L.stages;
          for (int s = 0; s < WX_NT; ++s) {
            mbar_wait(bar(B_WXEMPTY + s * 4 + st), ((t / L.stages) & 1) ^ 1);
            const uint32_t fb = bar(B_WXFULL + s * 4 + st);
            mbar_expect_tx(fb, (uint32_t)L.ring_stage_bytes);
            tma_load_3d(smem_u32(sm + L.ring + (s * L.stages + st) * L.ring_stage_bytes), &wxmap, unit0, row0 + s * NS, t, fb);
          }
        }
      }
      __syncwarp();
    } else {
      // =========================== epilogue warps ===================================================
      const int ew = warp, quad = warp & 3, es = (ew >> 2) & 1, rh = ew >> 3, part = ew >> 2;
      const int n = quad * 32 + lane;                     // local unit = TMEM lane
      const int gu = unit0 + n;                           // global hidden unit
      const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
      const bool hi_layout = a.layout == FGRNN_LAYOUT_HI;
      // ---- U^T slice -> tensor memory.  A[m = n][k] = Uc[k][gu]; the four warps of a quadrant take 16-k blocks
      //      part, part + 4, ...; scale from max |U| over the CTA's slice
      float mu = 0.f;
      for (int k = part; k < HK; k += 4) mu = fmaxf(mu, fabsf(hi_layout ? __ldg(a.U + (size_t)gu * HK + k) : __ldg(a.U + (size_t)k * HK + gu)));
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mu = fmaxf(mu, __shfl_xor_sync(0xffffffffu, mu, o));
      // largest gate/update bias distance of any unit of this CTA: selects the activation form (fgrnn_tc.cu)
      float bd = part == 0 ? fabsf(__ldg(a.bias_gate + gu) - __ldg(a.bias_update + gu)) : 0.f;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) bd = fmaxf(bd, __shfl_xor_sync(0xffffffffu, bd, o));
      if (lane == 0) { red_s[ew] = mu; red_s[16 + ew] = bd; }
      asm volatile("bar.sync 1, %0;" ::"n"(WX_EPI_WARPS * 32) : "memory");
#pragma unroll
      for (int w = 0; w < WX_EPI_WARPS; ++w) { mu = fmaxf(mu, red_s[w]); bd = fmaxf(bd, red_s[16 + w]); }
      int S = 40;
      if (mu > 0.f) S = min(S, (int)floorf(log2f(30000.f / mu)));
      S = max(S, -14);
      const float scale_u = exp2f((float)S), unscale = exp2f((float)-S);
      // e_u = e_g^2 * exp(2 (b_g - b_u)) saves one MUFU.EX2 per element: sigmoid gate, no per-unit scales, biases <= 8 apart
      const bool one_ex2 = d.gate_nl == FGRNN_NL_SIGMOID && !a.gate_scale && !a.update_scale && bd <= 8.0f;
      for (int kb = part; kb < KH; kb += 4) {
        float uv[16];
#pragma unroll
        for (int jj = 0; jj < 16; ++jj) {
          const int k = kb * 16 + jj;
          uv[jj] = hi_layout ? __ldg(a.U + (size_t)gu * HK + k) : __ldg(a.U + (size_t)k * HK + gu);
        }
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) split2(uv[2 * jj], uv[2 * jj + 1], scale_u, hi[jj], lo[jj]);
        tmem_st8(tmem + lane_base + TM_U_HI + kb * 8, hi);
        tmem_st8(tmem + lane_base + TM_U_LO + kb * 8, lo);
      }
      tmem_st_wait();

      // ---- per-unit constants of the gate update (rnn.py:289-295; BatchNorm variant rnn.py:402-408 folded)
      EpiK kc;
      {
        constexpr float LOG2E = 1.4426950408889634f;
        const float sz = sigmoid_f(__ldg(a.zeta)), sn = sigmoid_f(__ldg(a.nu));
        const float sg = a.gate_scale ? __ldg(a.gate_scale + gu) : 1.0f, su = a.update_scale ? __ldg(a.update_scale + gu) : 1.0f;
        const float gf = d.gate_nl == FGRNN_NL_TANH ? -2.0f * LOG2E : -LOG2E;          // tanh(a) = 2 / (1 + e^{-2a}) - 1
        const float kg = gf * sg, cg = gf * __ldg(a.bias_gate + gu);
        const float ku = -2.0f * LOG2E * su, cu = -2.0f * LOG2E * __ldg(a.bias_update + gu);
        kc.un = make_float2(unscale, unscale);
        kc.kg = make_float2(kg, kg); kc.cg = make_float2(cg, cg);
        kc.ku = make_float2(ku, ku); kc.cu = make_float2(cu, cu);
        kc.msz = make_float2(-sz, -sz); kc.szn = make_float2(sz + sn, sz + sn);
        kc.sg = sg; kc.bg = __ldg(a.bias_gate + gu); kc.su = su; kc.bu = __ldg(a.bias_update + gu);
        if (one_ex2) {                                   // cu becomes the ratio exp(2 (b_g - b_u)); pre >= tmin keeps e_g <= 2^30
          const float ratio = expf(2.0f * (kc.bg - kc.bu));
          kc.cu = make_float2(ratio, ratio);
          kc.tmin = (30.0f - cg) / kg;
        }
      }
      const bool tanh_gate = d.gate_nl == FGRNN_NL_TANH;

      // ---- state and operand-tile addresses
      float2 hst[PAIRS];
      const int first_row = row0 + es * NS + rh * RPT;
      // chunk of (unit gu, row group) inside an MN-major [HK][NS] tile; + buffer / sub-tile / hi|lo offsets
      const uint32_t chunk_off = (uint32_t)((gu >> 3) * ((NS >> 3) * 128) + (rh * NG) * 128 + (gu & 7) * 16);
      const uint32_t tile_off = (uint32_t)(es * 2 * L.h_tile_bytes);
      const uint32_t hbuf_bytes = (uint32_t)(WX_NT * 2 * L.h_tile_bytes);
      unsigned char* hop0 = sm + L.h_op + tile_off + chunk_off;
      const uint32_t hop0_remote = PAIR ? mapa_u32(smem_u32(hop0), peer) : 0u;
      const uint32_t hr_remote0 = PAIR ? mapa_u32(bar(B_HREADY + es * 2), peer) : 0u;      // partner's HREADY[es][0]; [1] is +8
      constexpr uint32_t PEER_BYTES = 32u * NG * 2u * 16u;       // what one epilogue warp receives from its counterpart per step
      // h (for the MMAs of step tn) -> operand tiles of buffer tn & 1, here and in the partner CTA
      auto put_operand = [&](int tn, const uint32_t (&hi)[PAIRS], const uint32_t (&lo)[PAIRS]) {
        const uint32_t bo = PAIR ? (uint32_t)(tn & 1) * hbuf_bytes : 0u;
        const uint32_t rbar = hr_remote0 + (uint32_t)(tn & 1) * 8u;
#pragma unroll
        for (int g = 0; g < NG; ++g) {
          const uint4 vh = make_uint4(hi[4 * g], hi[4 * g + 1], hi[4 * g + 2], hi[4 * g + 3]);
          const uint4 vl = make_uint4(lo[4 * g], lo[4 * g + 1], lo[4 * g + 2], lo[4 * g + 3]);
          const uint32_t o = bo + g * 128;
          *reinterpret_cast<uint4*>(hop0 + o) = vh;
          *reinterpret_cast<uint4*>(hop0 + o + L.h_tile_bytes) = vl;
          if (PAIR) { st_async_v4(hop0_remote + o, vh, rbar); st_async_v4(hop0_remote + o + L.h_tile_bytes, vl, rbar); }
        }
      };
      auto hand_off = [&](int tn) {
        fence_proxy_async_smem();                          // this CTA's st.shared -> visible to its tcgen05.mma
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (PAIR) mbar_expect_tx(hr_bar(es, tn), PEER_BYTES);   // arrive + expect the counterpart warp's bytes
          else mbar_arrive(hr_bar(es, tn));
        }
      };
      {
        uint32_t hi[PAIRS], lo[PAIRS];
#pragma unroll
        for (int q = 0; q < PAIRS; ++q) {
          const int row = first_row + 2 * q;
          const float v0 = (a.h0 && row < d.B) ? __ldg(a.h0 + (size_t)row * HK + gu) : 0.f;
          const float v1 = (a.h0 && row + 1 < d.B) ? __ldg(a.h0 + (size_t)(row + 1) * HK + gu) : 0.f;
          hst[q] = make_float2(v0, v1);
          split2(v0, v1, 1.0f, hi[q], lo[q]);
        }
        put_operand(0, hi, lo);
      }
      fence_proxy_async_smem();
      tc_fence_before();
      __syncthreads();                                   // matches the other roles' prologue barrier; weights are in TMEM
      hand_off(0);                                       // phase 0: h_{-1} ready

      const uint32_t acc = tmem + lane_base + TM_ACC + es * TM_ACC_PER_TILE + rh * RPT;
      char* outp = a.out ? reinterpret_cast<char*>(a.out + (size_t)first_row * a.osb + gu) : nullptr;
      const uint32_t row_bytes = (uint32_t)a.osb * 4u, step_bytes = (uint32_t)a.ost * 4u;
      float* zp = a.save_z ? a.save_z + (size_t)first_row * HK + gu : nullptr;
      float* cp = a.save_c ? a.save_c + (size_t)first_row * HK + gu : nullptr;
      const uint32_t zc_step = (uint32_t)d.B * HK;
      const int rows_left = d.B - first_row;
      const float2 one = make_float2(1.0f, 1.0f), two = make_float2(2.0f, 2.0f), mone = make_float2(-1.0f, -1.0f);
      // The step loop, specialised at compile time on the activation form (0: sigmoid gate, one MUFU.EX2 per element;
      // 1: sigmoid gate, two; 2: tanh gate) and on whether z_t / c_t are stored (training forward, cu:340-341).
      auto step_loop = [&](auto mode_tag, auto save_tag) {
        constexpr int MODE = decltype(mode_tag)::value;
        constexpr bool SAVE = decltype(save_tag)::value;
        for (int t = 0; t < d.T; ++t) {
          const int st = t % L.stages;
          mbar_wait(bar(B_DFULL + es), t & 1);
          tc_fence_after();
          mbar_wait(bar(B_WXFULL + es * 4 + st), (t / L.stages) & 1);
          const float* wx = reinterpret_cast<const float*>(sm + L.ring + (es * L.stages + st) * L.ring_stage_bytes) + (rh * RPT) * WX_HC + n;
          uint32_t hi[PAIRS], lo[PAIRS];
#pragma unroll
          for (int g = 0; g < NG; ++g) {
            float va[8], vb[8], v1[8], v2[8];
            tmem_ld8(acc + g * 8, va);
            tmem_ld8(acc + NS + g * 8, vb);
            tmem_ld8(acc + 2 * NS + g * 8, v1);
            tmem_ld8(acc + 3 * NS + g * 8, v2);
            tmem_ld_wait();
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const int p = g * 4 + q;
              const float2 w2 = make_float2(wx[(2 * p) * WX_HC], wx[(2 * p + 1) * WX_HC]);
              const float2 corr = __fadd2_rn(make_float2(va[2 * q], va[2 * q + 1]), make_float2(vb[2 * q], vb[2 * q + 1]));
              const float2 tot = __fadd2_rn(__fadd2_rn(corr, make_float2(v2[2 * q], v2[2 * q + 1])), make_float2(v1[2 * q], v1[2 * q + 1]));
              const float2 pre = __ffma2_rn(tot, kc.un, w2);                       // rnn.py:289 (wComp + uComp)
              float2 z, c;
              if (MODE < 2) {
                // sigmoid gate: one MUFU.RCP serves both gates (fgrnn_tc.cu): r = 1 / ((1 + e_g)(1 + e_u))
                float2 eg, eu;
                if (MODE == 0) {
                  const float2 pc = make_float2(fmax_nan(pre.x, kc.tmin), fmax_nan(pre.y, kc.tmin));
                  const float2 ag = __ffma2_rn(pc, kc.kg, kc.cg);
                  eg = make_float2(ex2_approx(ag.x), ex2_approx(ag.y));
                  eu = __fmul2_rn(__fmul2_rn(eg, eg), kc.cu);
                } else {
                  float2 ag = __ffma2_rn(pre, kc.kg, kc.cg), au = __ffma2_rn(pre, kc.ku, kc.cu);
                  ag.x = fmin_nan(ag.x, 60.0f); ag.y = fmin_nan(ag.y, 60.0f);
                  au.x = fmin_nan(au.x, 60.0f); au.y = fmin_nan(au.y, 60.0f);
                  eg = make_float2(ex2_approx(ag.x), ex2_approx(ag.y)); eu = make_float2(ex2_approx(au.x), ex2_approx(au.y));
                }
                const float2 ga = __fadd2_rn(eg, one), ub = __fadd2_rn(eu, one);
                const float2 ab = __fmul2_rn(ga, ub);
                const float2 r = make_float2(rcp_approx(ab.x), rcp_approx(ab.y));
                z = __fmul2_rn(r, ub);                                              // rnn.py:290
                c = __ffma2_rn(__fmul2_rn(r, ga), two, mone);                       // rnn.py:292
              } else {
                // tanh gate: z multiplies h and crosses zero, where 2 / (1 + e) - 1 cancels; tanh_fast switches to an odd
                // polynomial there (<= 2 ulp everywhere)
                z = make_float2(tanh_fast(fmaf(kc.sg, pre.x, kc.bg)), tanh_fast(fmaf(kc.sg, pre.y, kc.bg)));
                c = make_float2(tanh_fast(fmaf(kc.su, pre.x, kc.bu)), tanh_fast(fmaf(kc.su, pre.y, kc.bu)));
              }
              hst[p] = __ffma2_rn(z, __ffma2_rn(kc.msz, c, hst[p]), __fmul2_rn(kc.szn, c));   // rnn.py:294-295
              if (SAVE) {
                const int rj = 2 * p;
                if (rj < rows_left) { zp[rj * HK] = z.x; cp[rj * HK] = c.x; }
                if (rj + 1 < rows_left) { zp[(rj + 1) * HK] = z.y; cp[(rj + 1) * HK] = c.y; }
              }
              const __half2 hh = __float22half2_rn(hst[p]);
              const float2 hf = __half22float2(hh);
              const __half2 hl = __float22half2_rn(__fadd2_rn(hst[p], make_float2(-hf.x, -hf.y)));
              hi[p] = *reinterpret_cast<const uint32_t*>(&hh);
              lo[p] = *reinterpret_cast<const uint32_t*>(&hl);
            }
          }
          if (t + 1 < d.T) {                                 // nobody reads h_{T-1} as an operand (and no store may be in flight
            put_operand(t + 1, hi, lo);                      // towards the partner when the pair leaves the kernel)
            hand_off(t + 1);
          }
          if (lane == 0) mbar_arrive(bar(B_WXEMPTY + es * 4 + st));
          if (outp) {
#pragma unroll
            for (int q = 0; q < PAIRS; ++q) {
              if (2 * q < rows_left) *reinterpret_cast<float*>(outp + (size_t)(2 * q) * row_bytes) = hst[q].x;
              if (2 * q + 1 < rows_left) *reinterpret_cast<float*>(outp + (size_t)(2 * q + 1) * row_bytes) = hst[q].y;
            }
            outp += step_bytes;
          }
          if (SAVE) { zp += zc_step; cp += zc_step; }
        }
      };
      using I0 = IntTag<0>; using I1 = IntTag<1>; using I2 = IntTag<2>; using BF = BoolTag<false>; using BT = BoolTag<true>;
      const int mode = tanh_gate ? 2 : (one_ex2 ? 0 : 1);
      switch (mode * 2 + (zp ? 1 : 0)) {
        case 0: step_loop(I0{}, BF{}); break;
        case 1: step_loop(I0{}, BT{}); break;
        case 2: step_loop(I1{}, BF{}); break;
        case 3: step_loop(I1{}, BT{}); break;
        case 4: step_loop(I2{}, BF{}); break;
        default: step_loop(I2{}, BT{}); break;
      }
      if (a.h_last) {
#pragma unroll
        for (int j = 0; j < RPT; ++j) {
          const int row = first_row + j;
          if (row < d.B) a.h_last[(size_t)row * HK + gu] = (j & 1) ? hst[j >> 1].y : hst[j >> 1].x;
        }
      }
    }
    tc_fence_before();
    if (PAIR) cluster_sync_all(); else __syncthreads();          // no CTA leaves while its partner may still write into it
    if (warp == W_MMA) tmem_dealloc(tmem, TMEM_COLS);
  }
};

template <int NS, bool PAIR>
__global__ void __launch_bounds__(WXF_THREADS, 1) tc_wx_fwd_kernel(const WxFwdArgs a, const __grid_constant__ CUtensorMap wxmap) {
  TcWxFwd<NS, PAIR>::run(a, wxmap);
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
bool tc_wide_supports(const Dims& d) {
  // full rank, H = 128 or 256, I a multiple of 8 up to 256, sigmoid or tanh gate, tanh update
  return d.rW == 0 && d.rU == 0 && (d.H == 128 || d.H == 256) && d.I >= 8 && d.I <= 256 && (d.I % 8) == 0 &&
         (d.gate_nl == FGRNN_NL_SIGMOID || d.gate_nl == FGRNN_NL_TANH) && d.update_nl == FGRNN_NL_TANH;
}
size_t tc_wide_workspace_floats(const Dims& d) { return (size_t)d.T * d.B * d.H; }

static int make_wx_map(CUtensorMap* map, const float* wx, int B, int T, int H, int box_rows) {
  EncodeTiledFn encode = encode_tiled_fn();
  if (!encode) { set_error_detail("cuTensorMapEncodeTiled is not available from the driver"); return FGRNN_ERR_CUDA; }
  cuuint64_t gdim[3] = {(cuuint64_t)H, (cuuint64_t)B, (cuuint64_t)T};
  cuuint64_t gstr[2] = {(cuuint64_t)H * 4, (cuuint64_t)B * H * 4};
  cuuint32_t box[3] = {(cuuint32_t)WX_HC, (cuuint32_t)box_rows, 1}, estr[3] = {1, 1, 1};
  const CUresult cr = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(wx), gdim, gstr, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) { set_error_detail("cuTensorMapEncodeTiled (WX) failed with CUresult %d", (int)cr); return FGRNN_ERR_CUDA; }
  return FGRNN_OK;
}

template <int NS, bool PAIR>
static int launch_wx_fwd_t(const WxFwdArgs& a, const CUtensorMap& map, cudaStream_t stream) {
  using K = TcWxFwd<NS, PAIR>;
  const WxfSmem L = K::smem_layout();
  FGRNN_CUDA_TRY(cudaFuncSetAttribute(tc_wx_fwd_kernel<NS, PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total));
  const unsigned tiles = (unsigned)((a.d.B + K::ROWS - 1) / K::ROWS);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(PAIR ? 2 * tiles : tiles);
  cfg.blockDim = dim3(WXF_THREADS);
  cfg.dynamicSmemBytes = (size_t)L.total;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = PAIR ? 2 : 1; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  FGRNN_CUDA_TRY(cudaLaunchKernelEx(&cfg, tc_wx_fwd_kernel<NS, PAIR>, a, map));
  FGRNN_LAUNCH_CHECK("tc_wx_fwd_kernel");
  return FGRNN_OK;
}

int launch_tc_wide_fwd(const SmemFwdArgs& s, const float* gate_scale, const float* update_scale, float* wx_ws, cudaStream_t stream) {
  const Dims& d = s.d;
  if (d.B <= 0 || d.T <= 0) return FGRNN_OK;
  // 1. WX = x . W for all T steps
  XwArgs xa{};
  xa.d = d; xa.layout = s.layout; xa.W = s.W; xa.wx = wx_ws;
  xa.KI = (d.I + 15) & ~15;
  xa.KSW = xa.KI < 64 ? xa.KI : 64;
  xa.BOXI = d.I;
  xa.nslab = (xa.KI + 63) / 64;
  xa.nrb = (d.B + XW_ROWS - 1) / XW_ROWS;
  xa.ntiles = d.T * xa.nrb;
  CUtensorMap xmap;
  int rc = make_row_tile_map(&xmap, s.x, d.x_dtype == FGRNN_BF16, d.I, d.B, d.T, s.xsb, s.xst, XW_CONV_ROWS, &xa.x_time_outer, xa.BOXI);
  if (rc) return rc;
  const int esz = d.x_dtype == FGRNN_BF16 ? 2 : 4;
  const XwSmem XL = xw_smem_layout(xa.BOXI, xa.KSW, esz, xa.nslab);
  FGRNN_CUDA_TRY(cudaFuncSetAttribute(tc_xw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, XL.total));
  const int hy = d.H / WX_HC;
  int gx = xa.ntiles < 148 * 4 / hy ? xa.ntiles : 148 / hy;           // persistent: one CTA per SM over both unit halves
  if (gx < 1) gx = 1;
  tc_xw_kernel<<<dim3((unsigned)gx, (unsigned)hy), XW_THREADS, XL.total, stream>>>(xa, xmap);
  FGRNN_LAUNCH_CHECK("tc_xw_kernel");
  // 2. the recurrence
  WxFwdArgs wa{};
  wa.d = d; wa.layout = s.layout; wa.U = s.U;
  wa.bias_gate = s.bias_gate; wa.bias_update = s.bias_update; wa.zeta = s.zeta; wa.nu = s.nu;
  wa.gate_scale = gate_scale; wa.update_scale = update_scale;
  wa.h0 = s.h0; wa.out = s.out; wa.osb = s.osb; wa.ost = s.ost;
  wa.h_last = s.h_last; wa.save_z = s.save_z; wa.save_c = s.save_c;
  CUtensorMap wxmap;
  if ((rc = make_wx_map(&wxmap, wx_ws, d.B, d.T, d.H, 32))) return rc;
  return d.H == 256 ? launch_wx_fwd_t<32, true>(wa, wxmap, stream) : launch_wx_fwd_t<32, false>(wa, wxmap, stream);
}

}  // namespace fgrnn
