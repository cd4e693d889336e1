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
