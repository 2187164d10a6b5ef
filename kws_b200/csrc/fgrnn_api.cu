// C-ABI entry points of libfastgrnn_b200.so (declared in include/fastgrnn_b200.h):
// validation, workspace carving, kernel-family selection, launch sequencing.
#include <algorithm>
#include <atomic>
#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "fgrnn_kernels.cuh"

namespace fgrnn {

static thread_local char g_detail[512] = "";
static std::atomic<uint64_t> g_launches{0};

void set_error_detail(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_detail, sizeof(g_detail), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

static const char* const kTuneNames[TUNE_COUNT] = {"FGRNN_TC_NS", "FGRNN_TC_NT", "FGRNN_TC_BR_NS", "FGRNN_TC_WIDE", "FGRNN_FAST_NL", "FGRNN_SMEM_CFG", "FGRNN_TC_LR", "FGRNN_TC_ALT", "FGRNN_TC_ACC2", "FGRNN_TC_BWD_FUSED", "FGRNN_TC_VR"};
static std::atomic<int> g_tune[TUNE_COUNT];
static std::once_flag g_tune_once;
static int tune_parse(int key, const char* e) {
  if (key == TUNE_SMEM_CFG) return e[0] == 'A' && e[1] == '7' ? '7' : e[0];       // 'A' | 'A7' | 'B' | 'C'
  return atoi(e);
}
static void tune_init() {
  std::call_once(g_tune_once, [] {
    for (int k = 0; k < TUNE_COUNT; ++k) {
      const char* e = getenv(kTuneNames[k]);
      g_tune[k].store(e && e[0] ? tune_parse(k, e) : TUNE_UNSET);
    }
  });
}
int tuning(TuneKey key) { tune_init(); return g_tune[key].load(std::memory_order_relaxed); }

namespace {

constexpr int kMaxDim = 1024;   // I, H, rW, rU upper bound (generic-family shared-memory budget)

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_detail, sizeof(g_detail), fmt, ap);
  va_end(ap);
  return code;
}

bool nl_ok(int nl) { return nl >= FGRNN_NL_SIGMOID && nl <= FGRNN_NL_QUANT_SIGM4; }

int validate_problem(const FgrnnProblem& p) {
  if (p.abi_version != FGRNN_ABI_VERSION)
    return fail(FGRNN_ERR_VERSION, "desc.abi_version=%d, library=%d", p.abi_version, FGRNN_ABI_VERSION);
  if (p.B < 0 || p.T < 0) return fail(FGRNN_ERR_SHAPE, "B=%d T=%d must be >= 0", p.B, p.T);
  if (p.I < 1 || p.I > kMaxDim) return fail(FGRNN_ERR_SHAPE, "input size I=%d outside [1,%d]", p.I, kMaxDim);
  if (p.H < 1 || p.H > kMaxDim) return fail(FGRNN_ERR_SHAPE, "hidden size H=%d outside [1,%d]", p.H, kMaxDim);
  if (p.rW < 0 || p.rW > kMaxDim) return fail(FGRNN_ERR_SHAPE, "wRank=%d outside [0,%d]", p.rW, kMaxDim);
  if (p.rU < 0 || p.rU > kMaxDim) return fail(FGRNN_ERR_SHAPE, "uRank=%d outside [0,%d]", p.rU, kMaxDim);
  if ((int64_t)p.B * p.T > (int64_t)0x7fffffff / 4) return fail(FGRNN_ERR_SHAPE, "B*T=%lld too large", (long long)p.B * p.T);
  if (!nl_ok(p.gate_nl)) return fail(FGRNN_ERR_ENUM, "gate_nl=%d unknown", p.gate_nl);
  if (!nl_ok(p.update_nl)) return fail(FGRNN_ERR_ENUM, "update_nl=%d unknown", p.update_nl);
  if (p.weight_layout != FGRNN_LAYOUT_IH && p.weight_layout != FGRNN_LAYOUT_HI)
    return fail(FGRNN_ERR_ENUM, "weight_layout=%d unknown", p.weight_layout);
  if (p.x_dtype != FGRNN_F32 && p.x_dtype != FGRNN_BF16) return fail(FGRNN_ERR_ENUM, "x_dtype=%d unknown", p.x_dtype);
  if (p.force_path < -1 || p.force_path > FGRNN_PATH_LOWRANK) return fail(FGRNN_ERR_ENUM, "force_path=%d unknown", p.force_path);
  if (p.rW == 0 && !p.W) return fail(FGRNN_ERR_NULL, "W must be a CUDA tensor (NULL with wRank == 0)");
  if (p.rW > 0 && (!p.W1 || !p.W2)) return fail(FGRNN_ERR_NULL, "W1/W2 must be CUDA tensors (NULL with wRank > 0)");
  if (p.rU == 0 && !p.U) return fail(FGRNN_ERR_NULL, "U must be a CUDA tensor (NULL with uRank == 0)");
  if (p.rU > 0 && (!p.U1 || !p.U2)) return fail(FGRNN_ERR_NULL, "U1/U2 must be CUDA tensors (NULL with uRank > 0)");
  if (!p.bias_gate || !p.bias_update || !p.zeta || !p.nu)
    return fail(FGRNN_ERR_NULL, "bias_gate/bias_update/zeta/nu must be CUDA tensors (NULL)");
  if ((int64_t)p.B * p.T > 0 && !p.x) return fail(FGRNN_ERR_NULL, "input must be a CUDA tensor (NULL)");
  const uintptr_t xa = reinterpret_cast<uintptr_t>(p.x);
  if (xa % (p.x_dtype == FGRNN_BF16 ? 2 : 4)) return fail(FGRNN_ERR_ALIGN, "input pointer misaligned");
  return FGRNN_OK;
}

Dims dims_of(const FgrnnProblem& p) {
  Dims d;
  d.B = p.B; d.T = p.T; d.I = p.I; d.H = p.H; d.rW = p.rW; d.rU = p.rU;
  d.gate_nl = p.gate_nl; d.update_nl = p.update_nl; d.x_dtype = p.x_dtype;
  return d;
}

struct DeviceGuard {
  int prev = -1;
  bool ok = false;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) return;
    // cudaSetDevice even when the ordinal is already current: it binds the primary context to THIS thread, which the
    // driver-API tensor-map encoder (cuTensorMapEncodeTiled) needs -- autograd worker threads that never touched
    // the runtime otherwise get CUDA_ERROR_INVALID_CONTEXT.  (Legal during stream capture, unlike cudaFree(0).)
    if (cudaSetDevice(dev) != cudaSuccess) return;
    ok = true;
  }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

// ----------------------------------------------------------------------------------------
// forward planning
// ----------------------------------------------------------------------------------------
struct FwdPlan {
  int path;
  bool tc_wide;               // tcgen05 family, hoisted-projection kernels (fgrnn_tc_wx.cu)
  float* wx;                  // [T][B][H] workspace of the hoisted input projection
  // canonical weight pointers (either the caller's or workspace copies)
  float *Wc, *Uc, *W1c, *W2c, *U1c, *U2c;
  size_t ws_bytes;
};

bool aligned16(const void* ptr) { return (reinterpret_cast<uintptr_t>(ptr) & 15) == 0; }
bool mult4(int64_t v) { return (v & 3) == 0; }

// the persistent shared-memory family needs 16-byte vector access on every streamed tensor
bool has_scales(const FgrnnProblem& p) { return p.gate_scale || p.update_scale; }

bool smem_fwd_ok(const FgrnnForward& f) {
  const FgrnnProblem& p = f.p;
  if (has_scales(p) || !smem_path_supports(dims_of(p))) return false;
  const bool xok = p.x_dtype == FGRNN_BF16 ? (reinterpret_cast<uintptr_t>(p.x) & 7) == 0 : aligned16(p.x);
  return xok && mult4(p.x_stride_b) && mult4(p.x_stride_t) && aligned16(p.W) && aligned16(p.U) &&
         aligned16(p.h0) && aligned16(f.out) && mult4(f.out_stride_b) && mult4(f.out_stride_t) &&
         aligned16(f.h_last) && aligned16(f.save_z) && aligned16(f.save_c) && (!f.save_z == !f.save_c);
}

// the tcgen05 kernels keep output offsets in 32-bit registers: byte strides and B*H*4 must stay below 2^32
bool fits_u32_bytes(int64_t elems) { return elems >= 0 && elems < ((int64_t)1 << 30); }

// tcgen05 family: same streaming requirements; x is fetched by TMA (16-byte aligned base and strides)
bool tc_fwd_ok(const FgrnnForward& f) {
  const FgrnnProblem& p = f.p;
  // h_t, z_t, c_t leave through TMA tile stores: positive output strides (alignment: smem_fwd_ok); the three staging tiles of a
  // training forward do not fit beside the operand rings of I > 32
  if (f.out && (f.out_stride_b <= 0 || f.out_stride_t <= 0)) return false;
  if (f.save_z && p.I > 32) return false;
  return smem_fwd_ok(f) && tc_path_supports(dims_of(p)) &&
         tc_x_tma_ok(p.x, p.x_stride_b, p.x_stride_t, p.x_dtype, p.B, p.T) &&
         fits_u32_bytes(f.out_stride_b) && fits_u32_bytes(f.out_stride_t) && fits_u32_bytes((int64_t)p.B * p.H);
}

// tcgen05 family, wide shapes: x by TMA; everything else is accessed element-wise (4-byte alignment)
bool tc_wide_ok(const FgrnnForward& f) {
  const FgrnnProblem& p = f.p;
  return tc_wide_supports(dims_of(p)) && tc_x_tma_ok(p.x, p.x_stride_b, p.x_stride_t, p.x_dtype, p.B, p.T) &&
         fits_u32_bytes(f.out_stride_b) && fits_u32_bytes(f.out_stride_t) && fits_u32_bytes((int64_t)p.B * p.H) &&
         (!f.save_z == !f.save_c);
}

// low-rank FFMA family: 16-byte vector access on x, h0 and every output; weights are copied from their canonical form
bool lr_fwd_ok(const FgrnnForward& f) {
  const FgrnnProblem& p = f.p;
  if (has_scales(p) || !lr_path_supports(dims_of(p))) return false;
  const bool xok = p.x_dtype == FGRNN_BF16 ? (reinterpret_cast<uintptr_t>(p.x) & 7) == 0 : aligned16(p.x);
  return xok && mult4(p.x_stride_b) && mult4(p.x_stride_t) && aligned16(p.W1) && aligned16(p.W2) && aligned16(p.U1) &&
         aligned16(p.U2) && aligned16(p.bias_gate) && aligned16(p.bias_update) && aligned16(p.h0) && aligned16(f.out) &&
         mult4(f.out_stride_b) && mult4(f.out_stride_t) && aligned16(f.h_last) && aligned16(f.save_z) &&
         aligned16(f.save_c) && (!f.save_z == !f.save_c);
}

// low-rank family on the tensor cores (fgrnn_tc_lr.cu): x by TMA, inference forward only; FGRNN_TC_LR=0 keeps the FFMA kernel
bool tc_lr_ok(const FgrnnForward& f) {
  const FgrnnProblem& p = f.p;
  if (tuning(TUNE_TC_LR) == 0) return false;
  return lr_fwd_ok(f) && tc_lr_supports(dims_of(p)) && !f.save_z && !f.save_c &&
         tc_x_tma_ok(p.x, p.x_stride_b, p.x_stride_t, p.x_dtype, p.B, p.T) &&
         fits_u32_bytes(f.out_stride_b) && fits_u32_bytes(f.out_stride_t) && fits_u32_bytes((int64_t)p.B * p.H);
}

int select_fwd_path(const FgrnnForward& f) {
  if (f.p.force_path >= 0) return f.p.force_path;
  if (tc_fwd_ok(f) || tc_wide_ok(f)) return FGRNN_PATH_TCGEN05;
  if (lr_fwd_ok(f)) return FGRNN_PATH_LOWRANK;       // 5x the FFMA family at C2, faster per step even for one CTA
  return smem_fwd_ok(f) ? FGRNN_PATH_SMEM : FGRNN_PATH_GENERIC;
}

FwdPlan plan_forward(const FgrnnForward& f, void* ws) {
  const FgrnnProblem& p = f.p;
  FwdPlan pl{};
  pl.path = select_fwd_path(f);
  Carver cv(ws);
  pl.tc_wide = pl.path == FGRNN_PATH_TCGEN05 && !tc_fwd_ok(f);
  if (pl.path == FGRNN_PATH_TCGEN05 && !pl.tc_wide && tc_wide_ok(f) && tuning(TUNE_TC_WIDE) != TUNE_UNSET)
    pl.tc_wide = tuning(TUNE_TC_WIDE) != 0;                                       // tests: hoisted kernels on a shape the fused kernel covers
  if (pl.tc_wide) pl.wx = cv.take<float>(tc_wide_workspace_floats(dims_of(p)));
  if (p.weight_layout == FGRNN_LAYOUT_HI && (pl.path == FGRNN_PATH_GENERIC || pl.path == FGRNN_PATH_LOWRANK)) {
    if (p.rW == 0) pl.Wc = cv.take<float>((size_t)p.I * p.H);
    else { pl.W1c = cv.take<float>((size_t)p.I * p.rW); pl.W2c = cv.take<float>((size_t)p.rW * p.H); }
    if (p.rU == 0) pl.Uc = cv.take<float>((size_t)p.H * p.H);
    else { pl.U1c = cv.take<float>((size_t)p.H * p.rU); pl.U2c = cv.take<float>((size_t)p.rU * p.H); }
  }
  pl.ws_bytes = cv.total();
  return pl;
}

int validate_forward(const FgrnnForward& f) {
  int rc = validate_problem(f.p);
  if (rc) return rc;
  if (!f.out && !f.h_last && (int64_t)f.p.B * f.p.T > 0) return fail(FGRNN_ERR_NULL, "out and h_last are both NULL");
  const int path = select_fwd_path(f);
  if (path == FGRNN_PATH_SMEM && !smem_fwd_ok(f))
    return fail(FGRNN_ERR_SHAPE, "forced shared-memory path needs full-rank H=128, I%%4==0, I<=64 and 16-byte aligned tensors");
  if (path == FGRNN_PATH_LOWRANK && !lr_fwd_ok(f))
    return fail(FGRNN_ERR_SHAPE, "forced low-rank path needs H=256, wRank%%4==0 (<=32), uRank%%4==0 (<=64), I%%4==0 (<=64) and 16-byte aligned tensors");
  if (path == FGRNN_PATH_TCGEN05 && !tc_fwd_ok(f) && !tc_wide_ok(f))
    return fail(FGRNN_ERR_SHAPE, "forced tcgen05 path needs full rank, H=128 or 256, I%%8==0, I<=256, sigmoid or tanh gate, tanh update, 16-byte aligned input rows");
  if (has_scales(f.p) && path != FGRNN_PATH_GENERIC && !(path == FGRNN_PATH_TCGEN05 && tc_wide_ok(f)))
    return fail(FGRNN_ERR_SHAPE, "gate_scale / update_scale are supported by the generic and the wide tcgen05 kernels only");
  return FGRNN_OK;
}

// ----------------------------------------------------------------------------------------
// backward planning
// ----------------------------------------------------------------------------------------
struct BwdPlan {
  int path;
  int nchunk, rows_per_chunk, nrec, rows_per_cta;
  bool want_w, want_u;
  bool tc_contract;           // dW / dU sums on the tensor cores (fgrnn_tc_bwd.cu), one partial per CTA
  int fused_ctas;             // > 0: reverse recurrence + contraction in ONE launch, this many contraction CTAs
  int* progress;              // [nrec][16] progress flags of the fused launch
  float *UT, *U2T, *U1T;      // workspace copies (nullptr => use caller's pointer directly)
  float *Wf;                  // canonical [I][H] for d_x (nullptr => caller's W usable directly)
  float *dpre, *rec_partial, *partW, *partU, *dWc, *dUc;
  size_t ws_bytes;
};

bool smem_bwd_ok(const FgrnnBackward& g) {
  const FgrnnProblem& p = g.p;
  if (!smem_path_supports(dims_of(p))) return false;
  return aligned16(p.U) && aligned16(p.h0) && aligned16(g.grad_h) && mult4(g.grad_stride_b) && mult4(g.grad_stride_t) &&
         aligned16(g.hs) && mult4(g.hs_stride_b) && mult4(g.hs_stride_t) && aligned16(g.z_s) && aligned16(g.c_s) &&
         aligned16(g.d_h0);
}

// the tcgen05 contraction streams 32-byte row segments with 16-byte vector loads
bool tc_contract_ok(const FgrnnBackward& g) {
  const FgrnnProblem& p = g.p;
  if (p.force_path == FGRNN_PATH_GENERIC || p.force_path == FGRNN_PATH_SMEM) return false;   // keep the FFMA kernels testable
  if (!tc_contract_supports(dims_of(p))) return false;
  const int64_t xm = p.x_dtype == FGRNN_BF16 ? 8 : 4;
  return aligned16(p.x) && p.x_stride_b % xm == 0 && p.x_stride_t % xm == 0 && aligned16(g.hs) && mult4(g.hs_stride_b) &&
         mult4(g.hs_stride_t) && aligned16(p.h0);
}

// tcgen05 reverse recurrence: 512-byte bulk row copies need 16-byte aligned rows
bool tc_bwd_ok(const FgrnnBackward& g) {
  const FgrnnProblem& p = g.p;
  if (!tc_bwd_rec_supports(dims_of(p))) return false;
  // grad_h and the hidden states are fetched through TMA maps built from the caller's strides: the same
  // requirements as the forward's x map (positive strides, 16-byte multiples) -- an expanded gradient with a
  // zero stride (e.g. from out.sum((0,1)).backward()) falls through to the FFMA families instead of failing
  // in cuTensorMapEncodeTiled after the path has been chosen
  if (!tc_x_tma_ok(g.grad_h, g.grad_stride_b, g.grad_stride_t, FGRNN_F32, p.B, p.T - g.grad_t0)) return false;
  if (g.hs && !tc_x_tma_ok(g.hs, g.hs_stride_b, g.hs_stride_t, FGRNN_F32, p.B, p.T)) return false;
  if (!fits_u32_bytes((int64_t)p.B * p.H)) return false;
  return aligned16(p.U) && aligned16(p.h0) && aligned16(g.grad_h) && mult4(g.grad_stride_b) && mult4(g.grad_stride_t) &&
         aligned16(g.hs) && mult4(g.hs_stride_b) && mult4(g.hs_stride_t) && aligned16(g.z_s) && aligned16(g.c_s);
}

int select_bwd_path(const FgrnnBackward& g) {
  if (g.p.force_path == FGRNN_PATH_LOWRANK) return FGRNN_PATH_GENERIC;      // forward-only family
  if (g.p.force_path >= 0) return g.p.force_path;
  if (tc_bwd_ok(g)) return FGRNN_PATH_TCGEN05;
  return smem_bwd_ok(g) ? FGRNN_PATH_SMEM : FGRNN_PATH_GENERIC;
}

BwdPlan plan_backward(const FgrnnBackward& g, void* ws) {
  const FgrnnProblem& p = g.p;
  BwdPlan pl{};
  pl.path = select_bwd_path(g);
  const int64_t M = (int64_t)p.B * p.T;
  pl.nchunk = (int)std::min<int64_t>(128, std::max<int64_t>(1, M / 512));
  pl.rows_per_chunk = (int)((M + pl.nchunk - 1) / std::max(1, pl.nchunk));
  pl.rows_per_chunk = (pl.rows_per_chunk + 15) / 16 * 16;
  if (pl.rows_per_chunk < 16) pl.rows_per_chunk = 16;
  pl.rows_per_cta = pl.path == FGRNN_PATH_SMEM ? smem_rows_per_cta(dims_of(p), 1) : 0;
  pl.nrec = pl.path == FGRNN_PATH_TCGEN05 ? tc_bwd_rec_ctas(dims_of(p))
          : pl.path == FGRNN_PATH_SMEM ? smem_bwd_rec_ctas(dims_of(p)) : gen_bwd_rec_ctas(dims_of(p));
  pl.want_w = g.d_W || g.d_W1 || g.d_W2;
  pl.want_u = g.d_U || g.d_U1 || g.d_U2;
  pl.tc_contract = tc_contract_ok(g);
  if (pl.tc_contract) pl.nchunk = tc_contract_ctas(dims_of(p));
  if (pl.tc_contract && pl.path == FGRNN_PATH_TCGEN05 && (pl.want_w || pl.want_u) && tuning(TUNE_TC_BWD_FUSED) != 0) {
    pl.fused_ctas = tc_bwd_fused_contract_ctas(dims_of(p));
    if (pl.fused_ctas > 0) pl.nchunk = std::min(pl.fused_ctas, pl.nchunk);
    pl.fused_ctas = pl.fused_ctas > 0 ? pl.nchunk : 0;
  }
  Carver cv(ws);
  const bool ih = p.weight_layout == FGRNN_LAYOUT_IH;
  if (ih && pl.path == FGRNN_PATH_GENERIC) {
    if (p.rU == 0) pl.UT = cv.take<float>((size_t)p.H * p.H);
    else { pl.U2T = cv.take<float>((size_t)p.H * p.rU); pl.U1T = cv.take<float>((size_t)p.rU * p.H); }
  }
  if (g.d_x && (p.rW > 0 || !ih)) pl.Wf = cv.take<float>((size_t)p.I * p.H);
  pl.dpre = cv.take<float>((size_t)M * p.H);
  pl.rec_partial = cv.take<float>((size_t)pl.nrec * (2 * p.H + 2));
  if (pl.fused_ctas > 0) pl.progress = cv.take<int>((size_t)pl.nrec * 16);
  if (pl.want_w) pl.partW = cv.take<float>((size_t)pl.nchunk * p.I * p.H);
  if (pl.want_u) pl.partU = cv.take<float>((size_t)pl.nchunk * p.H * p.H);
  if (pl.want_w && p.rW > 0) pl.dWc = cv.take<float>((size_t)p.I * p.H);
  if (pl.want_u && p.rU > 0) pl.dUc = cv.take<float>((size_t)p.H * p.H);
  pl.ws_bytes = cv.total();
  return pl;
}

int validate_backward(const FgrnnBackward& g) {
  int rc = validate_problem(g.p);
  if (rc) return rc;
  if (has_scales(g.p)) return fail(FGRNN_ERR_SHAPE, "gate_scale / update_scale (folded eval-mode BatchNorm) are forward only");
  if ((int64_t)g.p.B * g.p.T > 0) {
    if (!g.grad_h) return fail(FGRNN_ERR_NULL, "grad_h must be a CUDA tensor (NULL)");
    if (g.grad_t0 < 0 || g.grad_t0 >= g.p.T) return fail(FGRNN_ERR_SHAPE, "grad_t0 = %d must be in [0, T = %d)", g.grad_t0, g.p.T);
    if (!g.hs && g.p.T > 1) return fail(FGRNN_ERR_NULL, "hidden_states must be a CUDA tensor (NULL)");
    if (!g.z_s || !g.c_s) return fail(FGRNN_ERR_NULL, "z / h_prime must be CUDA tensors (NULL)");
  }
  const int path = select_bwd_path(g);
  if (path == FGRNN_PATH_SMEM && !smem_bwd_ok(g))
    return fail(FGRNN_ERR_SHAPE, "forced shared-memory path needs full-rank H=128, I%%4==0, I<=64 and 16-byte aligned tensors");
  if (path == FGRNN_PATH_TCGEN05 && !tc_bwd_ok(g))
    return fail(FGRNN_ERR_SHAPE, "forced tcgen05 path needs full-rank H=128, sigmoid gate / tanh update and 16-byte aligned rows");
  return FGRNN_OK;
}

int check_device(int device) {
  int count = 0;
  FGRNN_CUDA_TRY(cudaGetDeviceCount(&count));
  if (device < 0 || device >= count) return fail(FGRNN_ERR_DEVICE, "device %d out of range (count %d)", device, count);
  static std::atomic<int> cc_major[64];
  int major = device < 64 ? cc_major[device].load() : 0;
  if (major == 0) {
    FGRNN_CUDA_TRY(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
    if (device < 64) cc_major[device].store(major);
  }
  if (major != 10) return fail(FGRNN_ERR_DEVICE, "device %d has compute capability %d.x; this library is sm_100a only", device, major);
  return FGRNN_OK;
}

}  // namespace
}  // namespace fgrnn

using namespace fgrnn;

extern "C" {

int fgrnn_abi_version(void) { return FGRNN_ABI_VERSION; }

int fgrnn_debug_set_tuning(const char* name, const char* value) {
  if (!name) return FGRNN_ERR_NULL;
  tune_init();
  for (int k = 0; k < TUNE_COUNT; ++k)
    if (!strcmp(name, kTuneNames[k])) {
      g_tune[k].store(value && value[0] ? tune_parse(k, value) : TUNE_UNSET);
      return FGRNN_OK;
    }
  return fail(FGRNN_ERR_ENUM, "unknown tuning key %s", name);
}
uint64_t fgrnn_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
const char* fgrnn_last_error_detail(void) { return g_detail; }

const char* fgrnn_strerror(int code) {
  switch (code) {
    case FGRNN_OK: return "ok";
    case FGRNN_ERR_NULL: return "required pointer is NULL";
    case FGRNN_ERR_SHAPE: return "unsupported shape";
    case FGRNN_ERR_ENUM: return "unknown enum value";
    case FGRNN_ERR_ALIGN: return "alignment requirement violated";
    case FGRNN_ERR_WORKSPACE: return "workspace missing or too small";
    case FGRNN_ERR_CUDA: return "CUDA error";
    case FGRNN_ERR_DEVICE: return "unsupported device";
    case FGRNN_ERR_VERSION: return "ABI version mismatch";
    default: return "unknown error";
  }
}

size_t fgrnn_forward_workspace_bytes(const FgrnnForward* f) {
  if (!f || validate_forward(*f)) return 0;
  return plan_forward(*f, nullptr).ws_bytes;
}

size_t fgrnn_backward_workspace_bytes(const FgrnnBackward* g) {
  if (!g || validate_backward(*g)) return 0;
  return plan_backward(*g, nullptr).ws_bytes;
}

int fgrnn_forward_plan(const FgrnnForward* f) {
  if (!f || validate_forward(*f)) return -1;
  return select_fwd_path(*f);
}

int fgrnn_backward_plan(const FgrnnBackward* g) {
  if (!g || validate_backward(*g)) return -1;
  return select_bwd_path(*g);
}

int64_t fgrnn_grad_bucket_layout(const FgrnnProblem* p, int64_t offsets[8]) {
  if (!p || !offsets) return -1;
  int64_t off = 0;
  const int64_t I = p->I, H = p->H, rW = p->rW, rU = p->rU;
  for (int i = 0; i < 8; ++i) offsets[i] = -1;
  // slots: 0 W|W1, 1 W2, 2 U|U1, 3 U2, 4 bias_gate, 5 bias_update, 6 zeta, 7 nu
  if (rW == 0) { offsets[0] = off; off += I * H; }
  else { offsets[0] = off; off += I * rW; offsets[1] = off; off += rW * H; }
  if (rU == 0) { offsets[2] = off; off += H * H; }
  else { offsets[2] = off; off += H * rU; offsets[3] = off; off += rU * H; }
  offsets[4] = off; off += H;
  offsets[5] = off; off += H;
  offsets[6] = off; off += 1;
  offsets[7] = off; off += 1;
  return off;
}

int fgrnn_forward(const FgrnnForward* f, void* stream_) {
  if (!f) return fail(FGRNN_ERR_NULL, "descriptor is NULL");
  g_detail[0] = 0;
  int rc = validate_forward(*f);
  if (rc) return rc;
  const FgrnnProblem& p = f->p;
  if ((int64_t)p.B * p.T == 0) return FGRNN_OK;
  rc = check_device(p.device);
  if (rc) return rc;
  DeviceGuard guard(p.device);
  if (!guard.ok) return fail(FGRNN_ERR_CUDA, "cudaSetDevice(%d) failed", p.device);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);

  FwdPlan pl = plan_forward(*f, f->workspace);
  if (pl.ws_bytes > 0) {
    if (!f->workspace || f->workspace_bytes < pl.ws_bytes)
      return fail(FGRNN_ERR_WORKSPACE, "forward needs %zu workspace bytes, got %zu", pl.ws_bytes, f->workspace_bytes);
    if (reinterpret_cast<uintptr_t>(f->workspace) % 256) return fail(FGRNN_ERR_ALIGN, "workspace must be 256-byte aligned");
  }

  if (pl.path == FGRNN_PATH_SMEM || pl.path == FGRNN_PATH_TCGEN05) {
    SmemFwdArgs s{};
    s.d = dims_of(p); s.layout = p.weight_layout; s.W = p.W; s.U = p.U;
    s.bias_gate = p.bias_gate; s.bias_update = p.bias_update; s.zeta = p.zeta; s.nu = p.nu;
    s.x = p.x; s.xsb = p.x_stride_b; s.xst = p.x_stride_t; s.h0 = p.h0;
    s.out = f->out; s.osb = f->out_stride_b; s.ost = f->out_stride_t;
    s.h_last = f->h_last; s.save_z = f->save_z; s.save_c = f->save_c;
    if (pl.tc_wide) return launch_tc_wide_fwd(s, p.gate_scale, p.update_scale, pl.wx, stream);
    return pl.path == FGRNN_PATH_TCGEN05 ? launch_tc_fwd(s, stream) : launch_smem_fwd(s, stream);
  }

  FwdArgs a{};
  a.d = dims_of(p);
  a.bias_gate = p.bias_gate; a.bias_update = p.bias_update; a.zeta = p.zeta; a.nu = p.nu;
  a.gate_scale = p.gate_scale; a.update_scale = p.update_scale;
  a.x = p.x; a.xsb = p.x_stride_b; a.xst = p.x_stride_t;
  a.h0 = p.h0;
  a.out = f->out; a.osb = f->out_stride_b; a.ost = f->out_stride_t;
  a.h_last = f->h_last; a.save_z = f->save_z; a.save_c = f->save_c;
  if (p.weight_layout == FGRNN_LAYOUT_IH) {
    a.Wc = p.W; a.Uc = p.U; a.W1c = p.W1; a.W2c = p.W2; a.U1c = p.U1; a.U2c = p.U2;
  } else {
    // FastGRNNCUDA layout (rnn.py:782-805): every matrix is stored transposed
    PrepJobs jobs{};
    auto add = [&](const float* src, float* dst, int rows, int cols) {
      jobs.job[jobs.n++] = PrepJob{src, dst, rows, cols, 1};
    };
    if (p.rW == 0) add(p.W, pl.Wc, p.H, p.I);
    else { add(p.W1, pl.W1c, p.rW, p.I); add(p.W2, pl.W2c, p.H, p.rW); }
    if (p.rU == 0) add(p.U, pl.Uc, p.H, p.H);
    else { add(p.U1, pl.U1c, p.rU, p.H); add(p.U2, pl.U2c, p.H, p.rU); }
    rc = launch_prep(jobs, stream);
    if (rc) return rc;
    a.Wc = pl.Wc; a.Uc = pl.Uc; a.W1c = pl.W1c; a.W2c = pl.W2c; a.U1c = pl.U1c; a.U2c = pl.U2c;
  }
  if (pl.path == FGRNN_PATH_LOWRANK) return tc_lr_ok(*f) ? launch_tc_lr_fwd(a, stream) : launch_lr_fwd(a, stream);
  return launch_gen_fwd(a, stream);
}

int fgrnn_backward(const FgrnnBackward* g, void* stream_) {
  if (!g) return fail(FGRNN_ERR_NULL, "descriptor is NULL");
  g_detail[0] = 0;
  int rc = validate_backward(*g);
  if (rc) return rc;
  const FgrnnProblem& p = g->p;
  rc = check_device(p.device);
  if (rc) return rc;
  DeviceGuard guard(p.device);
  if (!guard.ok) return fail(FGRNN_ERR_CUDA, "cudaSetDevice(%d) failed", p.device);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const bool ih = p.weight_layout == FGRNN_LAYOUT_IH;
  const int64_t M = (int64_t)p.B * p.T;

  if (M == 0) {
    // empty batch / sequence: all parameter gradients are zero, d_h0 = 0
    auto zero = [&](float* ptr, size_t n) -> int {
      if (ptr && n) FGRNN_CUDA_TRY(cudaMemsetAsync(ptr, 0, n * sizeof(float), stream));
      return FGRNN_OK;
    };
    if ((rc = zero(g->d_W, (size_t)p.I * p.H))) return rc;
    if ((rc = zero(g->d_U, (size_t)p.H * p.H))) return rc;
    if ((rc = zero(g->d_W1, (size_t)p.I * p.rW))) return rc;
    if ((rc = zero(g->d_W2, (size_t)p.rW * p.H))) return rc;
    if ((rc = zero(g->d_U1, (size_t)p.H * p.rU))) return rc;
    if ((rc = zero(g->d_U2, (size_t)p.rU * p.H))) return rc;
    if ((rc = zero(g->d_bias_gate, p.H))) return rc;
    if ((rc = zero(g->d_bias_update, p.H))) return rc;
    if ((rc = zero(g->d_zeta, 1))) return rc;
    if ((rc = zero(g->d_nu, 1))) return rc;
    if ((rc = zero(g->d_h0, (size_t)p.B * p.H))) return rc;
    return FGRNN_OK;
  }

  BwdPlan pl = plan_backward(*g, g->workspace);
  if (!g->workspace || g->workspace_bytes < pl.ws_bytes)
    return fail(FGRNN_ERR_WORKSPACE, "backward needs %zu workspace bytes, got %zu", pl.ws_bytes, g->workspace_bytes);
  if (reinterpret_cast<uintptr_t>(g->workspace) % 256) return fail(FGRNN_ERR_ALIGN, "workspace must be 256-byte aligned");

  // 1. parameter-sized layout fixes
  PrepJobs jobs{};
  auto add = [&](const float* src, float* dst, int rows, int cols, int tr) {
    jobs.job[jobs.n++] = PrepJob{src, dst, rows, cols, tr};
  };
  const float *UT = nullptr, *U2T = nullptr, *U1T = nullptr, *Wf = nullptr;
  if (pl.path == FGRNN_PATH_SMEM || pl.path == FGRNN_PATH_TCGEN05) {
    // the persistent kernels read U in the caller's layout
  } else if (ih) {
    if (p.rU == 0) { add(p.U, pl.UT, p.H, p.H, 1); UT = pl.UT; }
    else { add(p.U2, pl.U2T, p.rU, p.H, 1); add(p.U1, pl.U1T, p.H, p.rU, 1); U2T = pl.U2T; U1T = pl.U1T; }
  } else {
    UT = p.U; U2T = p.U2; U1T = p.U1;    // the HI layout *is* the transposed one
  }
  if (g->d_x) {
    if (p.rW == 0) {
      if (ih) Wf = p.W;
      else { add(p.W, pl.Wf, p.H, p.I, 1); Wf = pl.Wf; }
    } else {
      Wf = pl.Wf;
    }
  }
  if ((rc = launch_prep(jobs, stream))) return rc;
  if (g->d_x && p.rW > 0) {
    // materialise W = W1.W2 in the canonical [I][H] orientation (the reference does the same, cu:434-436)
    SmallGemm sg{};
    sg.M = p.I; sg.N = p.H; sg.K = p.rW; sg.C = pl.Wf; sg.transC = 0;
    if (ih) { sg.A = p.W1; sg.lda = p.rW; sg.transA = 0; sg.B = p.W2; sg.ldb = p.H; sg.transB = 0; }
    else    { sg.A = p.W1; sg.lda = p.I;  sg.transA = 1; sg.B = p.W2; sg.ldb = p.rW; sg.transB = 1; }
    if ((rc = launch_small_gemm(sg, stream))) return rc;
  }

  // 2. serial reverse recurrence
  if (pl.path == FGRNN_PATH_SMEM || pl.path == FGRNN_PATH_TCGEN05) {
    SmemBwdArgs s{};
    s.d = dims_of(p); s.layout = p.weight_layout; s.U = p.U; s.zeta = p.zeta; s.nu = p.nu;
    s.grad_h = g->grad_h; s.gsb = g->grad_stride_b; s.gst = g->grad_stride_t; s.gt0 = g->grad_t0;
    s.hs = g->hs; s.hsb = g->hs_stride_b; s.hst = g->hs_stride_t;
    s.h0 = p.h0; s.z_s = g->z_s; s.c_s = g->c_s;
    s.dpre_ws = pl.dpre; s.rec_partial = pl.rec_partial; s.d_h0 = g->d_h0;
    if (pl.fused_ctas > 0) {
      TcContractLaunch c{};
      c.d = dims_of(p);
      c.x = p.x; c.xsb = p.x_stride_b; c.xst = p.x_stride_t;
      c.hs = g->hs; c.hsb = g->hs_stride_b; c.hst = g->hs_stride_t; c.h0 = p.h0;
      c.dpre = pl.dpre; c.partW = pl.want_w ? pl.partW : nullptr; c.partU = pl.want_u ? pl.partU : nullptr;
      if ((rc = launch_tc_bwd_fused(s, c, pl.fused_ctas, pl.progress, stream))) return rc;
    } else
    if ((rc = pl.path == FGRNN_PATH_TCGEN05 ? launch_tc_bwd_rec(s, stream) : launch_smem_bwd_rec(s, stream))) return rc;
  } else {
  BwdRecArgs r{};
  r.d = dims_of(p);
  r.zeta = p.zeta; r.nu = p.nu;
  r.UT = UT; r.U2T = U2T; r.U1T = U1T;
  r.grad_h = g->grad_h; r.gsb = g->grad_stride_b; r.gst = g->grad_stride_t; r.gt0 = g->grad_t0;
  r.hs = g->hs; r.hsb = g->hs_stride_b; r.hst = g->hs_stride_t;
  r.h0 = p.h0; r.z_s = g->z_s; r.c_s = g->c_s;
  r.dpre_ws = pl.dpre; r.rec_partial = pl.rec_partial; r.d_h0 = g->d_h0;
  if ((rc = launch_gen_bwd_rec(r, stream))) return rc;
  }

  // 3. T-parallel outer-product sums as per-chunk partials (no atomics)
  if (pl.fused_ctas > 0) {
    // done by the fused launch above
  } else if (pl.tc_contract && (pl.want_w || pl.want_u)) {
    TcContractLaunch c{};
    c.d = dims_of(p);
    c.x = p.x; c.xsb = p.x_stride_b; c.xst = p.x_stride_t;
    c.hs = g->hs; c.hsb = g->hs_stride_b; c.hst = g->hs_stride_t; c.h0 = p.h0;
    c.dpre = pl.dpre; c.partW = pl.want_w ? pl.partW : nullptr; c.partU = pl.want_u ? pl.partU : nullptr;
    if ((rc = launch_tc_contract(c, stream))) return rc;
  } else {
  if (pl.want_w) {
    TnArgs t{};
    t.M = (int)M; t.B = p.B; t.T = p.T; t.K = p.I; t.N = p.H;
    t.a = p.x; t.asb = p.x_stride_b; t.ast = p.x_stride_t; t.a_dtype = p.x_dtype; t.a_shift = 0; t.a_h0 = nullptr;
    t.dpre = pl.dpre; t.partial = pl.partW; t.rows_per_chunk = pl.rows_per_chunk;
    if ((rc = launch_gemm_tn_partial(t, pl.nchunk, stream))) return rc;
  }
  if (pl.want_u) {
    TnArgs t{};
    t.M = (int)M; t.B = p.B; t.T = p.T; t.K = p.H; t.N = p.H;
    t.a = g->hs; t.asb = g->hs_stride_b; t.ast = g->hs_stride_t; t.a_dtype = FGRNN_F32; t.a_shift = 1; t.a_h0 = p.h0;
    t.dpre = pl.dpre; t.partial = pl.partU; t.rows_per_chunk = pl.rows_per_chunk;
    if ((rc = launch_gemm_tn_partial(t, pl.nchunk, stream))) return rc;
  }
  }

  // 4. deterministic tree/linear reduce of all partials
  ReduceArgs ra{};
  ra.I = p.I; ra.H = p.H; ra.nchunk = pl.nchunk; ra.nrec = pl.nrec;
  ra.partW = pl.partW; ra.partU = pl.partU;
  if (pl.want_w) {
    if (p.rW == 0) { ra.dWc = g->d_W; ra.dW_transpose = ih ? 0 : 1; }
    else { ra.dWc = pl.dWc; ra.dW_transpose = 0; }
  }
  if (pl.want_u) {
    if (p.rU == 0) { ra.dUc = g->d_U; ra.dU_transpose = ih ? 0 : 1; }
    else { ra.dUc = pl.dUc; ra.dU_transpose = 0; }
  }
  ra.rec_partial = pl.rec_partial;
  ra.d_bias_gate = g->d_bias_gate; ra.d_bias_update = g->d_bias_update;
  ra.d_zeta = g->d_zeta; ra.d_nu = g->d_nu; ra.zeta = p.zeta; ra.nu = p.nu;
  if ((rc = launch_reduce(ra, stream))) return rc;

  // 5. low-rank chain rule on the reduced matrices (cu:546-555), honouring the layout
  if (pl.want_w && p.rW > 0) {
    if (g->d_W1) {   // dW1c[i][j] = sum_n dWc[i][n] * W2c[j][n]
      SmallGemm sg{};
      sg.M = p.I; sg.N = p.rW; sg.K = p.H; sg.A = pl.dWc; sg.lda = p.H; sg.transA = 0;
      sg.B = p.W2; if (ih) { sg.ldb = p.H; sg.transB = 1; } else { sg.ldb = p.rW; sg.transB = 0; }
      sg.C = g->d_W1; sg.transC = ih ? 0 : 1;
      if ((rc = launch_small_gemm(sg, stream))) return rc;
    }
    if (g->d_W2) {   // dW2c[j][n] = sum_i W1c[i][j] * dWc[i][n]
      SmallGemm sg{};
      sg.M = p.rW; sg.N = p.H; sg.K = p.I; sg.A = p.W1;
      if (ih) { sg.lda = p.rW; sg.transA = 1; } else { sg.lda = p.I; sg.transA = 0; }
      sg.B = pl.dWc; sg.ldb = p.H; sg.transB = 0;
      sg.C = g->d_W2; sg.transC = ih ? 0 : 1;
      if ((rc = launch_small_gemm(sg, stream))) return rc;
    }
  }
  if (pl.want_u && p.rU > 0) {
    if (g->d_U1) {   // dU1c[k][j] = sum_n dUc[k][n] * U2c[j][n]
      SmallGemm sg{};
      sg.M = p.H; sg.N = p.rU; sg.K = p.H; sg.A = pl.dUc; sg.lda = p.H; sg.transA = 0;
      sg.B = p.U2; if (ih) { sg.ldb = p.H; sg.transB = 1; } else { sg.ldb = p.rU; sg.transB = 0; }
      sg.C = g->d_U1; sg.transC = ih ? 0 : 1;
      if ((rc = launch_small_gemm(sg, stream))) return rc;
    }
    if (g->d_U2) {   // dU2c[j][n] = sum_k U1c[k][j] * dUc[k][n]
      SmallGemm sg{};
      sg.M = p.rU; sg.N = p.H; sg.K = p.H; sg.A = p.U1;
      if (ih) { sg.lda = p.rU; sg.transA = 1; } else { sg.lda = p.H; sg.transA = 0; }
      sg.B = pl.dUc; sg.ldb = p.H; sg.transB = 0;
      sg.C = g->d_U2; sg.transC = ih ? 0 : 1;
      if ((rc = launch_small_gemm(sg, stream))) return rc;
    }
  }

  // 6. d_x = dPre . W^T (cu:538), optional
  if (g->d_x) {
    NtArgs n{};
    n.M = (int)M; n.B = p.B; n.N = p.H; n.I = p.I;
    n.dpre = pl.dpre; n.Wf = Wf; n.dx = g->d_x; n.dsb = g->dx_stride_b; n.dst = g->dx_stride_t;
    if ((rc = launch_gemm_nt(n, stream))) return rc;
  }
  return FGRNN_OK;
}

}  // extern "C"
