// meanpool.cu — mean-of-word-vectors item encoder as a segmented gather-reduce (+ its backward scatter-add).
//
// ref: modules/content/mean_pool.py:27-33 (AverageEmbeddings.call: the mask tests `content != -1`, so with the
// 0-padded content matrix every one of the L positions counts and the pad token's row participates),
// models/model_framework.py:51-56 (content rows gathered per unique item id).
// One warp per item; lanes run along the embedding dimension so every word-row read is coalesced.
#include "common.cuh"

namespace nncf {

__global__ void __launch_bounds__(256)
meanpool_fwd_kernel(const float* __restrict__ W, int dw, const int32_t* __restrict__ content, int L,
                    const int32_t* __restrict__ item_ids, int n, float* __restrict__ out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int it = blockIdx.x * 8 + warp;
  if (it >= n) return;
  const int64_t row = item_ids ? item_ids[it] : it;
  const int32_t* c = content + row * L;
  float acc[8];
#pragma unroll
  for (int m = 0; m < 8; ++m) acc[m] = 0.0f;
  int cnt = 0;
  for (int l0 = 0; l0 < L; l0 += 32) {
    const int32_t mine = (l0 + lane < L) ? __ldg(c + l0 + lane) : -1;
    const int lim = min(32, L - l0);
    for (int t = 0; t < lim; ++t) {
      const int32_t w = __shfl_sync(0xffffffffu, mine, t);
      if (w < 0) continue;                       // `content != -1` mask
      ++cnt;
      const float* wr = W + (int64_t)w * dw;
#pragma unroll
      for (int m = 0; m < 8; ++m) {
        const int col = lane + 32 * m;
        if (col < dw) acc[m] += __ldg(wr + col);
      }
    }
  }
  const float inv = 1.0f / static_cast<float>(cnt);
#pragma unroll
  for (int m = 0; m < 8; ++m) {
    const int col = lane + 32 * m;
    if (col < dw) out[(int64_t)it * dw + col] = acc[m] * inv;
  }
}

__global__ void __launch_bounds__(256)
meanpool_bwd_kernel(float* __restrict__ dW, int dw, const int32_t* __restrict__ content, int L,
                    const int32_t* __restrict__ item_ids, int n, const float* __restrict__ dout) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int it = blockIdx.x * 8 + warp;
  if (it >= n) return;
  const int64_t row = item_ids ? item_ids[it] : it;
  const int32_t* c = content + row * L;
  // first pass: count valid positions and occurrences of the pad token 0 (one combined atomic for it)
  int cnt = 0, cnt0 = 0;
  for (int l = lane; l < L; l += 32) {
    const int32_t w = __ldg(c + l);
    cnt += (w >= 0);
    cnt0 += (w == 0);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    cnt0 += __shfl_xor_sync(0xffffffffu, cnt0, o);
  }
  float g[8];
  const float inv = 1.0f / static_cast<float>(cnt);
#pragma unroll
  for (int m = 0; m < 8; ++m) {
    const int col = lane + 32 * m;
    g[m] = (col < dw) ? dout[(int64_t)it * dw + col] * inv : 0.0f;
  }
  if (cnt0 > 0) {
#pragma unroll
    for (int m = 0; m < 8; ++m) {
      const int col = lane + 32 * m;
      if (col < dw) atomicAdd(dW + col, g[m] * static_cast<float>(cnt0));
    }
  }
  for (int l0 = 0; l0 < L; l0 += 32) {
    const int32_t mine = (l0 + lane < L) ? __ldg(c + l0 + lane) : -1;
    const int lim = min(32, L - l0);
    for (int t = 0; t < lim; ++t) {
      const int32_t w = __shfl_sync(0xffffffffu, mine, t);
      if (w <= 0) continue;
      float* wr = dW + (int64_t)w * dw;
#pragma unroll
      for (int m = 0; m < 8; ++m) {
        const int col = lane + 32 * m;
        if (col < dw) atomicAdd(wr + col, g[m]);
      }
    }
  }
}

}  // namespace nncf

using namespace nncf;

extern "C" int nncf_meanpool_fwd(const float* word_table_dev, int word_dim, const int32_t* content_dev, int content_len,
                                 const int32_t* item_ids_dev, int n_items, float* out_dev, void* stream) {
  NNCF_CHECK_ARG(n_items >= 0, "nncf_meanpool_fwd: n_items < 0");
  if (n_items == 0) return NNCF_OK;
  NNCF_CHECK_ARG(word_table_dev && content_dev && out_dev, "nncf_meanpool_fwd: null argument");
  NNCF_CHECK_ARG(word_dim >= 1 && word_dim <= 256, "nncf_meanpool_fwd: word_dim must be in [1, 256]");
  NNCF_CHECK_ARG(content_len >= 1, "nncf_meanpool_fwd: content_len must be >= 1");
  meanpool_fwd_kernel<<<ceil_div(n_items, 8), 256, 0, (cudaStream_t)stream>>>(word_table_dev, word_dim, content_dev,
                                                                              content_len, item_ids_dev, n_items, out_dev);
  NNCF_LAUNCH_OK();
  return NNCF_OK;
}

extern "C" int nncf_meanpool_bwd(float* grad_word_table_dev, int word_dim, const int32_t* content_dev, int content_len,
                                 const int32_t* item_ids_dev, int n_items, const float* grad_out_dev, void* stream) {
  NNCF_CHECK_ARG(n_items >= 0, "nncf_meanpool_bwd: n_items < 0");
  if (n_items == 0) return NNCF_OK;
  NNCF_CHECK_ARG(grad_word_table_dev && content_dev && grad_out_dev, "nncf_meanpool_bwd: null argument");
  NNCF_CHECK_ARG(word_dim >= 1 && word_dim <= 256, "nncf_meanpool_bwd: word_dim must be in [1, 256]");
  NNCF_CHECK_ARG(content_len >= 1, "nncf_meanpool_bwd: content_len must be >= 1");
  meanpool_bwd_kernel<<<ceil_div(n_items, 8), 256, 0, (cudaStream_t)stream>>>(grad_word_table_dev, word_dim, content_dev,
                                                                              content_len, item_ids_dev, n_items, grad_out_dev);
  NNCF_LAUNCH_OK();
  return NNCF_OK;
}
