// PTX wrappers shared by the tcgen05 / TMEM kernel family (mbarrier, tcgen05.mma / ld / st / commit, TMA, fp16 split).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "fgrnn_common.cuh"

namespace fgrnn {

// Busy delay on the SM clock (~2 cycles per ns).  The tcgen05 kernels never execute a plain `nanosleep` in a warp
// that issues tcgen05.mma / tcgen05.commit: measured in round 2 with the timing fuzzer (profiles/r02_first_launch_hunt.txt),
// NANOSLEEP between an MMA burst and its commit plus NANOSLEEP before the next barrier check left all three issuing
// warps blocked at their UTCHMMA / UTCBAR for good (a hard hang, never a wrong result), while the same delays spent
// spinning on the clock are harmless.  mbarrier.try_wait's own suspend (NANOSLEEP.SYNCS) is not affected.
__device__ __forceinline__ void tc_spin_ns(unsigned ns) {
  if (ns == 0u) return;
  const long long until = clock64() + 2ll * (long long)ns;
  while (clock64() < until) { }
}

// ---- timing fuzzer (developer builds: make fuzz -> -DFGRNN_TC_FUZZ) -----------------------------------
// Every synchronisation wrapper below calls tc_fuzz(site) first: a pseudo-random __nanosleep (1 in 16
// calls up to 16 us, 3 in 16 up to 1 us) that shuffles the relative timing of the warp roles.  A protocol that is
// correct gives bit-identical results under it (tools/first_launch_probe.cu compares against the quiet build).
#ifdef FGRNN_TC_FUZZ
static __device__ int g_tc_stuck = 0;      // set by the first wait that gives up; every later wait then returns at once
// progress board: every warp leaves (calls so far, last site) where a stuck warp's report can read it
__device__ __forceinline__ unsigned* tc_progress() { __shared__ unsigned board[2][32]; return &board[0][0]; }
__device__ __forceinline__ void tc_mark(unsigned site) {
  unsigned* b = tc_progress();
  const unsigned w = threadIdx.x >> 5;
  if (w < 32) { b[w] = b[w] + 1; b[32 + w] = site; }
}
#ifndef FGRNN_TC_FUZZ_SITES
#define FGRNN_TC_FUZZ_SITES 0xffffffffu    // bit i: site i sleeps (1 arrive, 2 expect_tx, 3/4 wait entry/exit, 5/6 poll wait, 7 commit, 8 TMA)
#endif
__device__ __forceinline__ void tc_fuzz(unsigned site) {
  tc_mark(site);
  if (!((FGRNN_TC_FUZZ_SITES >> site) & 1u)) return;
#ifdef FGRNN_TC_FUZZ_WARP_LO
  if ((threadIdx.x >> 5) < FGRNN_TC_FUZZ_WARP_LO || (threadIdx.x >> 5) > FGRNN_TC_FUZZ_WARP_HI) return;
#endif
  unsigned c, sm;
  asm volatile("mov.u32 %0, %%clock;" : "=r"(c));
  asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
  unsigned r = c ^ (sm * 0x9E3779B1u) ^ (site * 0x85EBCA6Bu) ^ ((threadIdx.x >> 5) * 0xC2B2AE35u);
  r ^= r >> 15; r *= 0x2C1B3C6Du; r ^= r >> 12; r *= 0x297A2D39u; r ^= r >> 15;
  // no shuffle: the lanes of a converged warp read the same %clock in the same issue slot, so r is warp uniform
  // wherever the caller is (and where it is not, the lanes merely sleep for different times and reconverge)
  const unsigned sel = r & 15u;
  const unsigned ns = sel == 0u ? ((r >> 4) & 0x3fffu) : (sel < 4u ? ((r >> 4) & 0x3ffu) : 0u);
#ifdef FGRNN_TC_FUZZ_NANOSLEEP          // see tc_spin_ns(): plain NANOSLEEP in the MMA-issuing warps can wedge tcgen05 issue
  if (ns) __nanosleep(ns);
#else
  tc_spin_ns(ns);
#endif
}
#else
#define tc_fuzz(site) do { } while (0)
#define tc_mark(site) do { } while (0)
#endif
#ifndef FGRNN_MBAR_SPIN_LIMIT
#define FGRNN_MBAR_SPIN_LIMIT (1u << 20)     // x 20 us per try_wait: ~20 s before the protocol-bug guard traps
#endif

// ---- raw PTX wrappers -------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  tc_fuzz(1);
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  tc_fuzz(2);
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Blocking wait: try_wait suspends the warp in hardware until the phase completes or the time hint (ns) runs
// out, so waiting warps do not burn issue slots polling (an unhinted try_wait loop took 25 % of all issued
// instructions away from the epilogue warps).
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  tc_fuzz(3);
  uint32_t ok = 0;
#ifdef FGRNN_TC_FUZZ
  unsigned long long t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
#endif
  for (uint32_t spins = 0; !ok; ++spins) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity), "r"(20000u) : "memory");
#ifdef FGRNN_TC_FUZZ
    if (*(volatile int*)&g_tc_stuck) break;
    unsigned long long t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    if (!ok && t1 - t0 > 2000000000ull) {     // fuzz build: 2 s of wall clock = deadlock; report who is stuck where and release everybody
      if ((threadIdx.x & 31) == 0 && blockIdx.x < 4) {
        printf("STUCK block %d warp %d barrier@%u parity %u after %u spins\n", (int)blockIdx.x, (int)(threadIdx.x >> 5), bar, parity, spins);
        if (atomicAdd(&tc_progress()[31], 1u) == 0u) {     // first reporter of the CTA dumps the board and the barrier words
          for (int w = 0; w < 24; ++w) printf("  board block %d warp %d: %u calls, last site %u\n", (int)blockIdx.x, w, tc_progress()[w], tc_progress()[32 + w]);
          const uint32_t base = tc_progress()[30];
          for (int i = 0; i < 32 && base; ++i) {
            unsigned long long v;
            asm volatile("ld.shared.b64 %0, [%1];" : "=l"(v) : "r"(base + 8 * i) : "memory");
            printf("  mbarrier[%d] block %d = %016llx\n", i, (int)blockIdx.x, v);
          }
        }
      }
      __nanosleep(1000000);                  // let the other stuck warps report before everything is released
      *(volatile int*)&g_tc_stuck = 1;
      break;
    }
#else
    if (spins > FGRNN_MBAR_SPIN_LIMIT) __trap();      // protocol bug guard: fail loudly instead of hanging the GPU
#endif
  }
  tc_fuzz(4);
}
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
// Polling wait (no suspend hint) for barriers completed by bulk-copy transaction counts on the critical path
__device__ __forceinline__ void mbar_wait_poll(uint32_t bar, uint32_t parity) {
  tc_fuzz(5);
  uint32_t ok = 0;
  for (uint32_t spins = 0; !ok; ++spins) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (spins > (1u << 26)) __trap();
  }
  tc_fuzz(6);
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
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\tselp.u32 %0, 1, 0, q;\n\t}" : "=r"(pred));
  return pred != 0;
}
// same, for code that already runs on one elected lane only
__device__ __forceinline__ void umma_ts1(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
               ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit1(uint32_t bar) {
  tc_fuzz(7);
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
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
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                 "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr) : "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tma_load_3d(uint32_t dst_smem, const CUtensorMap* map, int c0, int c1, int c2, uint32_t bar) {
  tc_fuzz(8);
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
               ::"r"(dst_smem), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(bar) : "memory");
}

// TMA tile store shared -> global (bulk async-group completion).  The source tile must have been made visible to the async
// proxy (fence.proxy.async after the st.shared) by every writing thread before the issuing thread gets here.
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, int c0, int c1, int c2, uint32_t src_smem) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(src_smem) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all but the newest N groups of this thread have finished READING their shared-memory source
template <int N>
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// v*scale = hi + lo with hi, lo fp16 (round to nearest); two values at once.  No clamp: NaN stays NaN and a value
// beyond the fp16 range (|v*scale| > 65504; weights are pre-scaled below 30000, so this means an input or state of
// that size) becomes Inf - Inf = NaN in the lo part -- the result is NaN, as loud as the reference's own NaN/Inf
// propagation, instead of a silently saturated finite number.
__device__ __forceinline__ void split2(float a, float b, float scale, uint32_t& hi, uint32_t& lo) {
  a *= scale; b *= scale;
  const __half2 h = __floats2half2_rn(a, b);
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

// v = hi + lo with hi, lo bf16 (round to nearest): 16 mantissa bits over the full fp32 exponent range -- the
// split used for gradients (tolerance 1e-4), whose magnitudes are far below the fp16 normal range
__device__ __forceinline__ void split2_bf16(float a, float b, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  const float2 hf = __bfloat1622float2(h);
  const __nv_bfloat162 l = __floats2bfloat162_rn(a - hf.x, b - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

// v = hi + mid + lo with three bf16 terms: 24 mantissa bits (fp32-exact up to the last rounding) over the full
// fp32 exponent range.  The reverse recurrence needs it: scalar gradients (zeta, nu) sum ~B*T*H signed terms and a
// 2^-17 relative error per element of delta survives the cancellation (measured: 0.9x the tolerance on an
// ill-conditioned case with the two-term split, 0.02x with three terms).
__device__ __forceinline__ void split3_bf16(float a, float b, uint32_t& hi, uint32_t& mid, uint32_t& lo) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  const float2 hf = __bfloat1622float2(h);
  const float ra = a - hf.x, rb = b - hf.y;                 // exact
  const __nv_bfloat162 m = __floats2bfloat162_rn(ra, rb);
  const float2 mf = __bfloat1622float2(m);
  const __nv_bfloat162 l = __floats2bfloat162_rn(ra - mf.x, rb - mf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  mid = *reinterpret_cast<const uint32_t*>(&m);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

// MN-major SWIZZLE_NONE operand tile [k][mn] (mn contiguous): core matrix = 8 k x 16 B (8 mn);
// next 8 mn: +128 B (SBO);  next 8 k: +(MN/8)*128 B (LBO).  Valid for A (idesc bit 15) and B (bit 16).
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t smem_addr, int MN) {
  const uint64_t sbo = 128 >> 4, lbo = (uint64_t)((MN >> 3) * 128) >> 4;
  return (uint64_t)((smem_addr >> 4) & 0x3fff) | (lbo << 16) | (sbo << 32) | (1ull << 46);
}
// D[tmem] (+)= A[smem] . B[smem]; runs on one (elected) lane
__device__ __forceinline__ void umma_ss1(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}


// ---- host side: TMA tensor maps -------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static inline EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = []() -> EncodeTiledFn {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) return nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

// Map over a [rows][steps][inner] tensor given by element strides (inner contiguous): box = `box_rows` rows of one
// step.  The two outer dimensions are ordered by stride; *time_outer tells the kernel the coordinate order:
// 0 -> {0, t, row}, 1 -> {0, row, t}.  Rows past the end are zero-filled by the TMA unit.
static inline int make_row_tile_map(CUtensorMap* map, const void* base, bool bf16, int inner, int rows, int steps,
                                    int64_t row_stride, int64_t step_stride, int box_rows, int* time_outer, int box_inner = 0) {
  EncodeTiledFn encode = encode_tiled_fn();
  if (!encode) { set_error_detail("cuTensorMapEncodeTiled is not available from the driver"); return FGRNN_ERR_CUDA; }
  const int esz = bf16 ? 2 : 4;
  *time_outer = step_stride >= row_stride ? 1 : 0;
  cuuint64_t gdim[3], gstr[2];
  cuuint32_t box[3], estr[3] = {1, 1, 1};
  gdim[0] = (cuuint64_t)inner; box[0] = (cuuint32_t)(box_inner > 0 ? box_inner : inner);   // box_inner: a feature slab; past `inner` the TMA unit zero-fills
  if (*time_outer) {
    gdim[1] = (cuuint64_t)rows; gdim[2] = (cuuint64_t)steps;
    gstr[0] = (cuuint64_t)row_stride * esz; gstr[1] = (cuuint64_t)step_stride * esz;
    box[1] = (cuuint32_t)box_rows; box[2] = 1;
  } else {
    gdim[1] = (cuuint64_t)steps; gdim[2] = (cuuint64_t)rows;
    gstr[0] = (cuuint64_t)step_stride * esz; gstr[1] = (cuuint64_t)row_stride * esz;
    box[1] = 1; box[2] = (cuuint32_t)box_rows;
  }
  const CUresult cr = encode(map, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(base),
                             gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                             CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) { set_error_detail("cuTensorMapEncodeTiled failed with CUresult %d", (int)cr); return FGRNN_ERR_CUDA; }
  return FGRNN_OK;
}

}  // namespace fgrnn
