// Symmetric peer memory over NVLink / NVSwitch for the data-parallel step (SURVEY 8(b) last bullet, 8(e)): one "arena" per
// rank, allocated here with cudaMalloc, exported with CUDA IPC and mapped into every other rank of the box, so kernels and
// copy engines address a peer's buffers directly:
//   * the contrastive-loss GEMMs read the OTHER ranks' feature blocks in place (TMA loads on peer addresses) -- the all-gather
//     of reference loss.py:55-76 happens inside the logits kernel, tile by tile;
//   * small exchanges (row-LSE vectors, loss / d(scale) partial sums) are peer loads in a fixed rank order (deterministic);
//   * gradient buckets are pushed to every peer by the copy engines while backward runs on the SMs (no SM-resident
//     collective kernel competing with the persistent GEMMs) and summed in rank order inside the fused AdamW.
// Synchronisation: int32 flags in each arena.  flags[idx][src] of rank d is written by rank src (system-scope release, after
// its payload) and polled by rank d (acquire); values are monotonically increasing tickets, so "ready" is `>= ticket`.
#include <string.h>

#include "vl_host.h"
#include "vl_sm100.cuh"

namespace vl {

constexpr int kMaxPeers = 8;
constexpr int kMaxFlags = 64;

struct Comm {
  int rank = -1, world = 0;
  int device = -1;
  void* arena = nullptr;  // this rank's arena
  size_t bytes = 0;
  void* peer[kMaxPeers] = {nullptr};  // every rank's arena in this process's address space (peer[rank] == arena)
  bool connected = false;
};
static Comm g_comm;

__device__ __forceinline__ void st_release_sys(int* p, int v) { asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ int ld_acquire_sys(const int* p) {
  int v;
  asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

struct PeerPtrs {
  void* p[kMaxPeers];
};

// thread p publishes `value` into flags[idx][my_rank] of peer p's arena (flags live at offset 0 of every arena)
__global__ void comm_signal_kernel(PeerPtrs peers, int world, int my_rank, int idx, int value) {
  const int p = threadIdx.x;
  if (p >= world) return;
  __threadfence_system();  // everything this stream wrote before (payloads, local or through the copy engines) is visible first
  st_release_sys(reinterpret_cast<int*>(peers.p[p]) + idx * kMaxPeers + my_rank, value);
}

// thread p waits until flags[idx][p] of THIS arena reaches `value`; a peer that never arrives traps the launch after ~20 s
// instead of hanging the device
__global__ void comm_wait_kernel(const int* flags, int world, int idx, int value) {
  const int p = threadIdx.x;
  if (p >= world) return;
  const int* f = flags + idx * kMaxPeers + p;
  const long long t0 = clock64();
  unsigned spins = 0;
  while (ld_acquire_sys(f) < value) {
    __nanosleep(200);
    if ((++spins & 1023u) == 0 && clock64() - t0 > 40000000000ll) __trap();
  }
}

// out[i] = sum over ranks p (in rank order) of srcs[p][i]      (loss / d(scale) partial sums; deterministic on every rank)
__global__ void __launch_bounds__(256) peer_reduce_kernel(PeerPtrs srcs, int world, long long n, float* __restrict__ out) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float s = 0.f;
    for (int p = 0; p < world; ++p) s += reinterpret_cast<const float*>(srcs.p[p])[i];
    out[i] = s;
  }
}

// out[p * n + i] = srcs[p][i]      (row-LSE vectors of every rank, rank-major: the column LSEs of the opposite direction)
__global__ void __launch_bounds__(256) peer_gather_kernel(PeerPtrs srcs, int world, long long n, float* __restrict__ out) {
  const long long total = n * world;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int p = static_cast<int>(i / n);
    out[i] = reinterpret_cast<const float*>(srcs.p[p])[i - p * n];
  }
}

static int comm_ready(const char* who) {
  if (!g_comm.connected) {
    set_error("%s: vl_comm_init / vl_comm_connect have not completed", who);
    return VL_EINVAL;
  }
  return 0;
}

static PeerPtrs peer_ptrs_at(int64_t offset) {
  PeerPtrs pp;
  for (int i = 0; i < kMaxPeers; ++i) pp.p[i] = i < g_comm.world ? static_cast<char*>(g_comm.peer[i]) + offset : nullptr;
  return pp;
}

}  // namespace vl

using namespace vl;

extern "C" {

int vl_comm_init(int32_t rank, int32_t world, int64_t arena_bytes, void* handle_out, void** arena_out) {
  VL_CHECK_ARG(world >= 1 && world <= kMaxPeers && rank >= 0 && rank < world, "vl_comm_init: rank %d / world %d (at most %d ranks of one box)", rank, world, kMaxPeers);
  VL_CHECK_ARG(arena_bytes >= (int64_t)(kMaxFlags * kMaxPeers * sizeof(int)) && handle_out && arena_out, "vl_comm_init: bad arguments");
  VL_CHECK_ARG(g_comm.arena == nullptr, "vl_comm_init: already initialised (vl_comm_destroy first)");
  VL_CUDA(cudaGetDevice(&g_comm.device));
  VL_CUDA(cudaMalloc(&g_comm.arena, static_cast<size_t>(arena_bytes)));
  VL_CUDA(cudaMemset(g_comm.arena, 0, static_cast<size_t>(arena_bytes)));
  cudaIpcMemHandle_t h;
  VL_CUDA(cudaIpcGetMemHandle(&h, g_comm.arena));
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  memcpy(handle_out, &h, sizeof(h));
  g_comm.rank = rank;
  g_comm.world = world;
  g_comm.bytes = static_cast<size_t>(arena_bytes);
  g_comm.peer[rank] = g_comm.arena;
  g_comm.connected = world == 1;
  *arena_out = g_comm.arena;
  VL_CUDA(cudaDeviceSynchronize());
  return 0;
}

int vl_comm_connect(const void* all_handles) {
  VL_CHECK_ARG(g_comm.arena != nullptr && all_handles, "vl_comm_connect: call vl_comm_init first");
  for (int p = 0; p < g_comm.world; ++p) {
    if (p == g_comm.rank) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, static_cast<const char*>(all_handles) + 64 * p, sizeof(h));
    VL_CUDA(cudaIpcOpenMemHandle(&g_comm.peer[p], h, cudaIpcMemLazyEnablePeerAccess));
  }
  g_comm.connected = true;
  return 0;
}

int vl_comm_peer_ptr(int32_t peer, void** ptr_out) {
  if (int rc = comm_ready("vl_comm_peer_ptr")) return rc;
  VL_CHECK_ARG(peer >= 0 && peer < g_comm.world && ptr_out, "vl_comm_peer_ptr: bad peer %d", peer);
  *ptr_out = g_comm.peer[peer];
  return 0;
}

int vl_comm_destroy(void) {
  if (g_comm.arena == nullptr) return 0;
  cudaDeviceSynchronize();
  for (int p = 0; p < g_comm.world; ++p)
    if (p != g_comm.rank && g_comm.peer[p]) cudaIpcCloseMemHandle(g_comm.peer[p]);
  cudaFree(g_comm.arena);
  g_comm = Comm();
  return 0;
}

int vl_comm_signal(int32_t flag_idx, int32_t value, void* stream) {
  if (int rc = comm_ready("vl_comm_signal")) return rc;
  VL_CHECK_ARG(flag_idx >= 0 && flag_idx < kMaxFlags, "vl_comm_signal: flag index %d", flag_idx);
  comm_signal_kernel<<<1, 32, 0, reinterpret_cast<cudaStream_t>(stream)>>>(peer_ptrs_at(0), g_comm.world, g_comm.rank, flag_idx, value);
  return launch_check("comm_signal");
}

int vl_comm_wait(int32_t flag_idx, int32_t value, void* stream) {
  if (int rc = comm_ready("vl_comm_wait")) return rc;
  VL_CHECK_ARG(flag_idx >= 0 && flag_idx < kMaxFlags, "vl_comm_wait: flag index %d", flag_idx);
  comm_wait_kernel<<<1, 32, 0, reinterpret_cast<cudaStream_t>(stream)>>>(static_cast<const int*>(g_comm.arena), g_comm.world, flag_idx, value);
  return launch_check("comm_wait");
}

int vl_comm_peer_reduce_f32(int64_t offset, int64_t n, float* out, void* stream) {
  if (int rc = comm_ready("vl_comm_peer_reduce_f32")) return rc;
  VL_CHECK_ARG(out && n > 0 && offset >= 0 && offset % 4 == 0 && static_cast<size_t>(offset + n * 4) <= g_comm.bytes, "vl_comm_peer_reduce_f32: range outside the arena");
  const int g = static_cast<int>((n + 255) / 256 < 64 ? (n + 255) / 256 : 64);
  peer_reduce_kernel<<<g, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(peer_ptrs_at(offset), g_comm.world, n, out);
  return launch_check("peer_reduce");
}

int vl_comm_peer_gather_f32(int64_t offset, int64_t n, float* out, void* stream) {
  if (int rc = comm_ready("vl_comm_peer_gather_f32")) return rc;
  VL_CHECK_ARG(out && n > 0 && offset >= 0 && offset % 4 == 0 && static_cast<size_t>(offset + n * 4) <= g_comm.bytes, "vl_comm_peer_gather_f32: range outside the arena");
  const long long tot = n * g_comm.world;
  const int g = static_cast<int>((tot + 255) / 256 < 128 ? (tot + 255) / 256 : 128);
  peer_gather_kernel<<<g, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(peer_ptrs_at(offset), g_comm.world, n, out);
  return launch_check("peer_gather");
}

/* Feature exchange (reference gather_features, loss.py:20-78): publish this rank's packed bf16 block, already written at
 * `offset` of its own arena, under ticket `ticket` on flag `flag_idx`.  Nothing is copied: consumers (vl_gemm_bf16 with
 * b_peers) load the blocks in place over NVLink after seeing the ticket. */
int vl_allgather_features(int64_t offset, int64_t bytes, int32_t flag_idx, int32_t ticket, void* stream) {
  if (int rc = comm_ready("vl_allgather_features")) return rc;
  VL_CHECK_ARG(offset >= 0 && bytes > 0 && static_cast<size_t>(offset + bytes) <= g_comm.bytes, "vl_allgather_features: range outside the arena");
  return vl_comm_signal(flag_idx, ticket, stream);
}

/* Gradient exchange (the DDP all-reduce of reference pc_tri_main.py:378-380) on the copy engines: the `bytes` at `offset` of this
 * rank's arena (its own slot of the gradient region, slot stride = `slot_stride` bytes, slot index = rank) are copied into the
 * same slot of every peer's arena with cudaMemcpyAsync on `stream`, then ticket `ticket` is published on `flag_idx`.  The
 * consumer (vl_adamw_multi with n_src = world) waits for the flag and adds the world slots in rank order. */
int vl_allreduce_grads(int64_t offset, int64_t bytes, int32_t flag_idx, int32_t ticket, void* stream) {
  if (int rc = comm_ready("vl_allreduce_grads")) return rc;
  VL_CHECK_ARG(offset >= 0 && bytes > 0 && static_cast<size_t>(offset + bytes) <= g_comm.bytes, "vl_allreduce_grads: range outside the arena");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const char* src = static_cast<const char*>(g_comm.arena) + offset;
  for (int k = 1; k < g_comm.world; ++k) {
    const int p = (g_comm.rank + k) % g_comm.world;  // staggered so that the ranks do not all hit the same peer first
    VL_CUDA(cudaMemcpyAsync(static_cast<char*>(g_comm.peer[p]) + offset, src, static_cast<size_t>(bytes), cudaMemcpyDeviceToDevice, s));
  }
  return vl_comm_signal(flag_idx, ticket, s);
}
}
