// Data-parallel training: the gradient all-reduce FUSED with the SGD step, over NVLink / NVSwitch peer memory.
//
// The data-parallel step of the keyword spotter (north star: one process per GPU, the batch sharded, the ~83 KB flat gradient
// bucket all-reduced, then torch.optim.SGD of trainClassifier.py:239-240) ends with  ncclAllReduce(bucket) -> sgd_flat_kernel.
// For a bucket this small the collective is pure latency.  Here the two are ONE kernel with ONE one-way NVLink latency in it:
//
//   * every rank owns a region (cudaMalloc + CUDA IPC handle, opened by all peers): its gradient bucket and a RECEIVE area
//     [2 step parities][source rank][lines];
//   * push: every thread packs two gradients and the step counter into a 16-byte line {g0, step, g1, step} and stores it
//     (one st.volatile.v4 per peer) straight into every peer's receive area -- posted writes through NVLink, all peers in
//     parallel through the NVSwitch, no fence and no flag round trip: each 8-byte half of a line carries its own flag, so a
//     line is valid exactly when both halves show the current step (the low-latency protocol of NCCL's LL kernels);
//   * reduce + update: the same thread polls ITS lines in its own receive area (local memory), adds the W contributions in
//     rank order 0..W-1 -- the same order on every rank, so all replicas compute the very same bits and cannot drift -- and
//     applies  p -= lr * (scale * sum);
//   * no closing handshake: the receive area is double buffered by step parity.  A rank overwrites parity p of a peer two steps
//     later, after it has passed a step in between, which needed that peer's lines of that step, which the peer pushed only
//     after its kernel of the step before had finished reading parity p.
// One-shot all-to-all: every rank sends (W-1) x 2 x 83 KB -- microseconds of NVLink time; ring or tree schedules only add hops
// at this size.  The step counters live in device memory (the kernel increments them itself), so a CUDA graph that captured the
// launch replays correctly.  A peer that never arrives (a crashed rank) does not hang the GPU: the poll gives up after
// PEER_TIMEOUT_NS, raises the error word of the state buffer and lets the kernel finish.
// History (tools/time_collective.py, two GPUs, 89.7 KB, per update inside a CUDA graph): ncclAllReduce + sgd_flat_kernel 14.7 us;
// a first version that PULLED the peers' buckets between two flag rounds (release / acquire at system scope) 22.5 us, 19.8 us
// with relaxed polls and one fence -- every round trip and every system-scope fence shows at this size.
#include "fgrnn_kernels.cuh"

namespace fgrnn {

constexpr int PEER_MAX = FGRNN_PEER_MAX_RANKS, PEER_CTAS = 32, PEER_THREADS = 256;
constexpr unsigned long long PEER_TIMEOUT_NS = 4000000000ull;

struct PeerArgs {
  float* params; float* reduced;            // local: parameters (updated), optional copy of the summed gradients
  const float* bucket;                      // local: this rank's gradients
  uint4* recv[PEER_MAX];                    // every rank's receive area as THIS process sees it: [2][world][nl] lines
  int* state;                               // local: [PEER_CTAS] step counters, [PEER_CTAS] error word
  long long n, nl;                          // floats, lines (two floats each)
  float lr, scale;
  int world, rank;
};

__device__ __forceinline__ void st_line(uint4* p, float a, float b, unsigned flag) {
  asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(__float_as_uint(a)), "r"(flag), "r"(__float_as_uint(b)), "r"(flag) : "memory");
}
__device__ __forceinline__ uint4 ld_line(const uint4* p) {
  uint4 v;
  asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long peer_now_ns() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }

__global__ void __launch_bounds__(PEER_THREADS) sgd_allreduce_peer_kernel(const PeerArgs a) {
  __shared__ int epoch_s;
  if (threadIdx.x == 0) epoch_s = a.state[blockIdx.x] + 1;
  __syncthreads();
  const unsigned epoch = (unsigned)epoch_s;
  const long long slot = (long long)(epoch & 1u) * a.world * a.nl;          // this step's half of a receive area
  const long long i0 = (long long)blockIdx.x * PEER_THREADS + threadIdx.x, stride = (long long)gridDim.x * PEER_THREADS;
  // ---- push this rank's gradients into every peer's receive area
  for (long long i = i0; i < a.nl; i += stride) {
    const float g0 = a.bucket[2 * i], g1 = 2 * i + 1 < a.n ? a.bucket[2 * i + 1] : 0.f;
    for (int r = 0; r < a.world; ++r)
      if (r != a.rank) st_line(a.recv[r] + slot + (long long)a.rank * a.nl + i, g0, g1, epoch);
  }
  // ---- collect the peers' lines (local polls), add in rank order, update
  const uint4* mine = a.recv[a.rank] + slot;
  bool gave_up = false;
  for (long long i = i0; i < a.nl; i += stride) {
    float s0 = 0.f, s1 = 0.f;
#pragma unroll 1
    for (int r = 0; r < a.world; ++r) {
      float v0, v1;
      if (r == a.rank) {
        v0 = a.bucket[2 * i]; v1 = 2 * i + 1 < a.n ? a.bucket[2 * i + 1] : 0.f;
      } else {
        const uint4* p = mine + (long long)r * a.nl + i;
        uint4 l = ld_line(p);
        unsigned long long t0 = 0;
        for (unsigned polls = 1; (l.y != epoch || l.w != epoch) && !gave_up; ++polls) {
          if ((polls & 1023u) == 0) {
            const unsigned long long now = peer_now_ns();
            if (t0 == 0) t0 = now;
            else if (now - t0 > PEER_TIMEOUT_NS) { atomicExch(a.state + PEER_CTAS, 1 + r); gave_up = true; }
          }
          l = ld_line(p);
        }
        v0 = __uint_as_float(l.x); v1 = __uint_as_float(l.z);
      }
      s0 = r == 0 ? v0 : s0 + v0; s1 = r == 0 ? v1 : s1 + v1;             // rank order: identical bits on every rank
    }
    a.params[2 * i] = fmaf(-a.lr, s0 * a.scale, a.params[2 * i]);
    if (a.reduced) a.reduced[2 * i] = s0;
    if (2 * i + 1 < a.n) {
      a.params[2 * i + 1] = fmaf(-a.lr, s1 * a.scale, a.params[2 * i + 1]);
      if (a.reduced) a.reduced[2 * i + 1] = s1;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) a.state[blockIdx.x] = (int)epoch;
}

}  // namespace fgrnn

using namespace fgrnn;

namespace {
struct DeviceScope {
  int prev = -1; bool ok = false;
  explicit DeviceScope(int device) { ok = cudaGetDevice(&prev) == cudaSuccess && cudaSetDevice(device) == cudaSuccess; }
  ~DeviceScope() { if (prev >= 0) cudaSetDevice(prev); }
};
}  // namespace

extern "C" size_t fgrnn_peer_recv_bytes(int64_t n, int32_t world) {
  if (n <= 0 || world < 1) return 0;
  return (size_t)2 * (size_t)world * (size_t)((n + 1) / 2) * sizeof(uint4);
}
extern "C" size_t fgrnn_peer_state_bytes(void) { return (size_t)(PEER_CTAS + 1) * sizeof(int); }

extern "C" int fgrnn_peer_alloc(size_t bytes, int32_t device, void** ptr, unsigned char* handle) {
  if (!ptr || !handle) { set_error_detail("peer_alloc: ptr / handle is NULL"); return FGRNN_ERR_NULL; }
  if (bytes == 0) { set_error_detail("peer_alloc: zero bytes"); return FGRNN_ERR_SHAPE; }
  DeviceScope ds(device);
  if (!ds.ok) { set_error_detail("peer_alloc: cannot select device %d", device); return FGRNN_ERR_DEVICE; }
  void* p = nullptr;
  FGRNN_CUDA_TRY(cudaMalloc(&p, bytes));
  cudaIpcMemHandle_t h;
  static_assert(sizeof(h) == FGRNN_PEER_HANDLE_BYTES, "CUDA IPC handle size");
  cudaError_t e = cudaMemset(p, 0, bytes);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    set_error_detail("peer_alloc: %s", cudaGetErrorString(e));
    cudaFree(p);
    return FGRNN_ERR_CUDA;
  }
  memcpy(handle, &h, sizeof(h));
  *ptr = p;
  return FGRNN_OK;
}

extern "C" int fgrnn_peer_open(const unsigned char* handle, int32_t device, void** ptr) {
  if (!ptr || !handle) { set_error_detail("peer_open: ptr / handle is NULL"); return FGRNN_ERR_NULL; }
  DeviceScope ds(device);
  if (!ds.ok) { set_error_detail("peer_open: cannot select device %d", device); return FGRNN_ERR_DEVICE; }
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  FGRNN_CUDA_TRY(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return FGRNN_OK;
}

extern "C" int fgrnn_peer_close(void* ptr, int32_t device) {
  if (!ptr) return FGRNN_OK;
  DeviceScope ds(device);
  if (!ds.ok) { set_error_detail("peer_close: cannot select device %d", device); return FGRNN_ERR_DEVICE; }
  FGRNN_CUDA_TRY(cudaIpcCloseMemHandle(ptr));
  return FGRNN_OK;
}

extern "C" int fgrnn_peer_free(void* ptr, int32_t device) {
  if (!ptr) return FGRNN_OK;
  DeviceScope ds(device);
  if (!ds.ok) { set_error_detail("peer_free: cannot select device %d", device); return FGRNN_ERR_DEVICE; }
  FGRNN_CUDA_TRY(cudaFree(ptr));
  return FGRNN_OK;
}

extern "C" int fgrnn_sgd_allreduce_peer(const FgrnnPeerStep* s, void* stream) {
  if (!s) { set_error_detail("sgd_allreduce_peer: descriptor is NULL"); return FGRNN_ERR_NULL; }
  if (s->abi_version != FGRNN_ABI_VERSION) { set_error_detail("sgd_allreduce_peer: abi_version %d, library %d", s->abi_version, FGRNN_ABI_VERSION); return FGRNN_ERR_VERSION; }
  if (s->world < 1 || s->world > PEER_MAX || s->rank < 0 || s->rank >= s->world) {
    set_error_detail("sgd_allreduce_peer: world %d (1..%d), rank %d", s->world, PEER_MAX, s->rank);
    return FGRNN_ERR_SHAPE;
  }
  if (!s->params || !s->state) { set_error_detail("sgd_allreduce_peer: params / state is NULL"); return FGRNN_ERR_NULL; }
  if (s->n <= 0) return FGRNN_OK;
  PeerArgs a{};
  a.params = s->params; a.reduced = s->reduced; a.bucket = s->bucket; a.state = s->state; a.n = s->n; a.nl = (s->n + 1) / 2;
  a.lr = s->lr; a.scale = s->grad_scale; a.world = s->world; a.rank = s->rank;
  if (!s->bucket) { set_error_detail("sgd_allreduce_peer: bucket is NULL"); return FGRNN_ERR_NULL; }
  for (int r = 0; r < s->world; ++r) {
    if (!s->recv[r]) { set_error_detail("sgd_allreduce_peer: receive area of rank %d is NULL", r); return FGRNN_ERR_NULL; }
    if (reinterpret_cast<uintptr_t>(s->recv[r]) & 15) { set_error_detail("sgd_allreduce_peer: receive area of rank %d must be 16-byte aligned", r); return FGRNN_ERR_ALIGN; }
    a.recv[r] = static_cast<uint4*>(s->recv[r]);
  }
  if ((reinterpret_cast<uintptr_t>(s->params) & 3) || (reinterpret_cast<uintptr_t>(s->bucket) & 3) || (reinterpret_cast<uintptr_t>(s->reduced) & 3)) {
    set_error_detail("sgd_allreduce_peer: params / bucket / reduced must be 4-byte aligned");
    return FGRNN_ERR_ALIGN;
  }
  DeviceScope ds(s->device);
  if (!ds.ok) { set_error_detail("sgd_allreduce_peer: cannot select device %d", s->device); return FGRNN_ERR_DEVICE; }
  // the grid depends on n only: every rank launches the same number of CTAs (CTA b pairs with CTA b of the peers)
  long long grid = (a.nl + PEER_THREADS - 1) / PEER_THREADS;
  if (grid > PEER_CTAS) grid = PEER_CTAS;
  if (grid < 1) grid = 1;
  sgd_allreduce_peer_kernel<<<(unsigned)grid, PEER_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(a);
  FGRNN_LAUNCH_CHECK("sgd_allreduce_peer_kernel");
  return FGRNN_OK;
}
