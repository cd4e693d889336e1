// common.cuh — error plumbing, launch accounting and the loss/gradient "epilogue" math shared by the
// CUDA-core (fp32) and tcgen05 (bf16) score kernels.
#pragma once
#include <cstdint>
#include <cstdio>
#include <string>
#include <atomic>
#include <cuda_runtime.h>
#include "../../include/nncf_b200.h"

namespace nncf {

void set_error(const std::string& msg);
extern std::atomic<int64_t> g_launches;
inline void count_launch(int n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

#define NNCF_CUDA(expr)                                                                              \
  do {                                                                                               \
    cudaError_t e__ = (expr);                                                                        \
    if (e__ != cudaSuccess) {                                                                        \
      ::nncf::set_error(std::string(#expr) + ": " + cudaGetErrorString(e__) + " (" + __FILE__ + ":" + \
                        std::to_string(__LINE__) + ")");                                             \
      return NNCF_ECUDA;                                                                             \
    }                                                                                                \
  } while (0)

#define NNCF_CHECK_ARG(cond, msg)           \
  do {                                      \
    if (!(cond)) {                          \
      ::nncf::set_error(std::string(msg));  \
      return NNCF_EINVAL;                   \
    }                                       \
  } while (0)

#define NNCF_LAUNCH_OK()                      \
  do {                                        \
    ::nncf::count_launch();                   \
    NNCF_CUDA(cudaGetLastError());            \
  } while (0)

// row-sharded tables (multi-GPU): row `id` lives on rank id % n at local row id / n; p[] are peer-mapped pointers
struct ShardPtrs {
  float* p[16];
  int n;                     // <= 1: not sharded
};

constexpr int kEpiWarps = 4;   // epilogue warps of the tcgen05 kernels (one per TMEM lane quadrant)

// Launch with programmatic dependent launch allowed: the kernel's CTAs may be scheduled while the preceding kernel of
// the stream is still running; the kernel itself calls pdl_wait() (griddepcontrol.wait) before it touches anything the
// predecessor wrote.  Only kernels that contain that wait may be launched this way.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

inline int ceil_div(int64_t a, int64_t b) { return static_cast<int>((a + b - 1) / b); }

// ------------------------------------------------------------------------------------------------
// Loss / dLoss/dScore for one score element.   ref: utils/objectives.py:78-117 (neg_shared),
// :163-220 (group_neg_shared); gradients per SURVEY.md Appendix A.
// ------------------------------------------------------------------------------------------------
struct EpiParams {
  int scheme;      // NNCF_SCHEME_NEG_SHARED / GROUP_NEG_SHARED
  int loss;        // NNCF_LOSS_*
  int B;           // valid rows
  int ncols;       // valid columns (B or n_unique)
  float w_neg;     // lambda / (ncols - 1)
  float gamma;
  float inv_b;     // 1 / B
  float inv_cnt;   // 1 / (B * ncols)   (pairwise losses average over every entry)
};

__device__ __forceinline__ EpiParams make_epi(int scheme, int loss, int B, int ncols, float lambda, float gamma) {
  EpiParams p;
  p.scheme = scheme; p.loss = loss; p.B = B; p.ncols = ncols;
  p.w_neg = lambda / static_cast<float>(ncols - 1);
  p.gamma = gamma;
  p.inv_b = 1.0f / static_cast<float>(B);
  p.inv_cnt = 1.0f / (static_cast<float>(B) * static_cast<float>(ncols));
  return p;
}

template <bool kFast>
__device__ __forceinline__ float softplus_f(float x) {   // log(1 + e^x)
  if (kFast) return fmaxf(x, 0.0f) + __logf(1.0f + __expf(-fabsf(x)));
  return fmaxf(x, 0.0f) + log1pf(expf(-fabsf(x)));
}
template <bool kFast>
__device__ __forceinline__ float sigmoid_f(float x) {
  if (kFast) return __fdividef(1.0f, 1.0f + __expf(-x));
  return 1.0f / (1.0f + expf(-x));
}

// s: score; is_pos: element is the row's positive; spos: the positive score it is compared with
// (neg_shared pairwise: S[j,j] of its column; group pairwise: P[i,pos_i] of its row).
// Returns g = dL/ds (without the diagonal/positive correction), a = dL/dD (pairwise only), and the
// element's loss contribution (already scaled so that a plain sum over elements gives the loss).
template <bool kFast>
__device__ __forceinline__ void epi_elem(const EpiParams& p, float s, bool is_pos, float spos, float& g, float& a,
                                         float& loss) {
  a = 0.0f;
  if (p.loss == NNCF_LOSS_SKIP_GRAM) {
    const float w = is_pos ? 1.0f : p.w_neg;
    const float sg = sigmoid_f<kFast>(s);
    loss = w * softplus_f<kFast>(is_pos ? -s : s) * p.inv_b;
    g = w * (sg - (is_pos ? 1.0f : 0.0f)) * p.inv_b;
  } else if (p.loss == NNCF_LOSS_MSE) {
    const float w = is_pos ? 1.0f : p.w_neg;
    const float t = s - (is_pos ? 1.0f : 0.0f);
    loss = w * t * t * p.inv_b;
    g = 2.0f * w * t * p.inv_b;
  } else if (p.loss == NNCF_LOSS_LOG_LOSS) {
    const float D = spos - s;
    loss = softplus_f<kFast>(-p.gamma * D) * p.inv_cnt;
    a = -p.gamma * sigmoid_f<kFast>(-p.gamma * D) * p.inv_cnt;
    g = -a;
  } else {   // max-margin
    const float D = spos - s;
    const float M = (p.scheme == NNCF_SCHEME_NEG_SHARED && is_pos) ? 0.0f : p.gamma;
    const float t = M - D;
    loss = fmaxf(t, 0.0f) * p.inv_cnt;
    a = (t > 0.0f) ? -p.inv_cnt : 0.0f;
    g = -a;
  }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

}  // namespace nncf
