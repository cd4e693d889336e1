// peer.cu — peer-visible device memory (CUDA IPC over NVLink / NVSwitch) and a device-side cross-GPU barrier.
//
// Multi-GPU training shards the embedding tables by row (owner = id mod N, local row = id div N).  Instead of
// staging rows through NCCL all-to-all buffers, the step kernels address the owners' shards DIRECTLY: the gather
// kernel's 128-bit loads and the update kernel's red.global.add.v4 go to peer-mapped pointers, so the exchange is
// part of the kernels that need the data (no extra launches, no host synchronisation, transfers overlap with the
// rest of the grid).  The reference has no multi-device path at all; this is the B200-native extension the
// north_star asks for (row-sharded tables, rows and gradients exchanged over NVLink).
#include <cstring>
#include "common.cuh"

namespace nncf {

struct PeerFlags {
  unsigned int* flags[16];   // flags[i] = the flag array (16 words) living on rank i
};

// Every rank writes `epoch` into slot [rank] of every peer's flag array, then waits until all slots of its own
// array reached `epoch`.  Bounded spin: a lost peer traps instead of hanging the GPU.
__global__ void peer_barrier_kernel(PeerFlags p, int n, int rank, unsigned int epoch) {
  const int i = threadIdx.x;
  if (i < n) {
    __threadfence_system();                              // my earlier peer writes / atomics are ordered before the flag
    volatile unsigned int* remote = p.flags[i];
    remote[rank] = epoch;
    __threadfence_system();
    volatile unsigned int* mine = p.flags[rank];
    unsigned int spins = 0;
    while (static_cast<int>(mine[i] - epoch) < 0) {
      if (++spins > (1u << 27)) __trap();
    }
  }
}

// Flag words in peer-visible memory for stream-ordered hand-offs between GPUs (the pipelined stratum rotation of
// nncf_b200/parallel.py): a producer's stream stores a monotonically increasing value AFTER the copies it covers, a
// consumer's stream waits until the word has reached the value it needs.  Bounded spin (~30 s): a lost peer traps.
__global__ void peer_signal_kernel(unsigned int* flag, unsigned int value) {
  __threadfence_system();                                // the stream's earlier peer writes are ordered before the flag
  *reinterpret_cast<volatile unsigned int*>(flag) = value;
  __threadfence_system();
}
__global__ void peer_wait_kernel(const unsigned int* flag, unsigned int value) {
  const volatile unsigned int* f = flag;
  unsigned int spins = 0;
  while (static_cast<int>(*f - value) < 0) {
    __nanosleep(256);
    if (++spins > 120000000u) __trap();
  }
  __threadfence_system();
}

}  // namespace nncf

using namespace nncf;

extern "C" int nncf_peer_alloc(size_t bytes, void** ptr_out, unsigned char* handle_out) {
  NNCF_CHECK_ARG(ptr_out && handle_out && bytes > 0, "nncf_peer_alloc: bad argument");
  void* p = nullptr;
  NNCF_CUDA(cudaMalloc(&p, bytes));
  NNCF_CUDA(cudaMemset(p, 0, bytes));
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    set_error(std::string("cudaIpcGetMemHandle: ") + cudaGetErrorString(e));
    return NNCF_ECUDA;
  }
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  memcpy(handle_out, &h, 64);
  *ptr_out = p;
  return NNCF_OK;
}

extern "C" int nncf_peer_open(const unsigned char* handle, void** ptr_out) {
  NNCF_CHECK_ARG(handle && ptr_out, "nncf_peer_open: null argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, 64);
  NNCF_CUDA(cudaIpcOpenMemHandle(ptr_out, h, cudaIpcMemLazyEnablePeerAccess));
  return NNCF_OK;
}

extern "C" int nncf_peer_close(void* ptr) {
  if (ptr) NNCF_CUDA(cudaIpcCloseMemHandle(ptr));
  return NNCF_OK;
}

extern "C" int nncf_peer_free(void* ptr) {
  if (ptr) NNCF_CUDA(cudaFree(ptr));
  return NNCF_OK;
}

extern "C" int nncf_peer_barrier(void* const* flag_ptrs, int n_ranks, int rank, unsigned int epoch, void* stream) {
  NNCF_CHECK_ARG(flag_ptrs && n_ranks >= 1 && n_ranks <= 16 && rank >= 0 && rank < n_ranks, "nncf_peer_barrier: bad argument");
  PeerFlags p{};
  for (int i = 0; i < n_ranks; ++i) p.flags[i] = static_cast<unsigned int*>(flag_ptrs[i]);
  peer_barrier_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(p, n_ranks, rank, epoch);
  NNCF_LAUNCH_OK();
  return NNCF_OK;
}


// ---- stream-ordered transfers between peer-visible allocations (copy engines over NVLink; no SM is used) ----------------
extern "C" int nncf_peer_copy(void* dst, const void* src, size_t bytes, void* stream) {
  NNCF_CHECK_ARG(dst && src, "nncf_peer_copy: null argument");
  if (bytes == 0) return NNCF_OK;
  NNCF_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, (cudaStream_t)stream));
  return NNCF_OK;
}
extern "C" int nncf_peer_signal(void* flag_ptr, unsigned int value, void* stream) {
  NNCF_CHECK_ARG(flag_ptr, "nncf_peer_signal: null flag");
  peer_signal_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(static_cast<unsigned int*>(flag_ptr), value);
  NNCF_LAUNCH_OK();
  return NNCF_OK;
}
extern "C" int nncf_peer_wait(const void* flag_ptr, unsigned int value, void* stream) {
  NNCF_CHECK_ARG(flag_ptr, "nncf_peer_wait: null flag");
  peer_wait_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(static_cast<const unsigned int*>(flag_ptr), value);
  NNCF_LAUNCH_OK();
  return NNCF_OK;
}
