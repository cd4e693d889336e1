// meanpool.cu — mean-of-word-vectors item encoder as a segmented gather-reduce (+ its backward scatter-add).
//
// ref: modules/content/mean_pool.py:27-33 (AverageEmbeddings.call: the mask tests `content != -1`, so with the
// 0-padded content matrix every one of the L positions counts and the pad token's row participates),
// models/model_framework.py:51-56 (content rows gathered per unique item id).
// One CTA per item: its 8 warps split the L word positions in chunks of 32 (lanes run along the embedding dimension, so
// every word-row read is coalesced) and keep EIGHT row loads in flight per lane; the partial sums meet in shared memory.
// (The first version - one warp per item walking its L rows one dependent load after the other - took 87 us for the C1
// batch of 512 unique items x 300 words, 0.04 of the HBM roofline: profiles/r02_kernels_summary.md.)
#include "common.cuh"

namespace nncf {

template <int NM>   // 32-column chunks of the word dimension: dw <= 32 NM
__global__ void __launch_bounds__(256)
meanpool_fwd_kernel(const float* __restrict__ W, int dw, const int32_t* __restrict__ content, int L,
                    const int32_t* __restrict__ item_ids, int n, float* __restrict__ out, const int32_t* __restrict__ n_valid) {
  __shared__ float part[8][32 * NM];
  __shared__ int cnts[8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int it = blockIdx.x;
  if (n_valid && it >= *n_valid) {                      // slot beyond the device-side count: a zero row, no id is read
    for (int col = threadIdx.x; col < dw; col += 256) out[(int64_t)it * dw + col] = 0.0f;
    return;
  }
  const int64_t row = item_ids ? item_ids[it] : it;
  const int32_t* c = content + row * L;
  float acc[NM];
#pragma unroll
  for (int m = 0; m < NM; ++m) acc[m] = 0.0f;
  int cnt = 0;
  for (int l0 = warp * 32; l0 < L; l0 += 256) {
    const int32_t mine = (l0 + lane < L) ? __ldg(c + l0 + lane) : -1;
#pragma unroll 1
    for (int t0 = 0; t0 < 32; t0 += 8) {
      float v[8][NM];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int32_t w = __shfl_sync(0xffffffffu, mine, t0 + j);      // -1 beyond L and for masked positions (`content != -1`)
        cnt += (w >= 0);
#pragma unroll
        for (int m = 0; m < NM; ++m) {
          const int col = lane + 32 * m;
          v[j][m] = (w >= 0 && col < dw) ? __ldg(W + (int64_t)w * dw + col) : 0.0f;
        }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int m = 0; m < NM; ++m) acc[m] += v[j][m];
    }
  }
#pragma unroll
  for (int m = 0; m < NM; ++m) part[warp][lane + 32 * m] = acc[m];
  if (lane == 0) cnts[warp] = cnt;
  __syncthreads();
  int total = 0;
#pragma unroll
  for (int w = 0; w < 8; ++w) total += cnts[w];
  const float inv = 1.0f / static_cast<float>(total);
  for (int col = threadIdx.x; col < dw; col += 256) {
    float sum = 0.0f;
#pragma unroll
    for (int w = 0; w < 8; ++w) sum += part[w][col];
    out[(int64_t)it * dw + col] = sum * inv;
  }
}

template <int NM>
__global__ void __launch_bounds__(256)
meanpool_bwd_kernel(float* __restrict__ dW, int dw, const int32_t* __restrict__ content, int L,
                    const int32_t* __restrict__ item_ids, int n, const float* __restrict__ dout, const int32_t* __restrict__ n_valid) {
  __shared__ int cnts[8], cnts0[8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int it = blockIdx.x;
  if (n_valid && it >= *n_valid) return;
  const int64_t row = item_ids ? item_ids[it] : it;
  const int32_t* c = content + row * L;
  // first pass: count valid positions and occurrences of the pad token 0 (one combined atomic row for it)
  int cnt = 0, cnt0 = 0;
  for (int l = threadIdx.x; l < L; l += 256) {
    const int32_t w = __ldg(c + l);
    cnt += (w >= 0);
    cnt0 += (w == 0);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    cnt0 += __shfl_xor_sync(0xffffffffu, cnt0, o);
  }
  if (lane == 0) { cnts[warp] = cnt; cnts0[warp] = cnt0; }
  __syncthreads();
  cnt = 0; cnt0 = 0;
#pragma unroll
  for (int w = 0; w < 8; ++w) { cnt += cnts[w]; cnt0 += cnts0[w]; }
  const float inv = 1.0f / static_cast<float>(cnt);
  if ((dw & 1) == 0) {
    // even word_dim: rows are 8-byte aligned, every lane adds PAIRS of columns with one red.global.add.v2.f32 (half the
    // reduction instructions of the scalar form; the kernel is bound by the L2's reduction rate)
    float2 g2[NM];
#pragma unroll
    for (int m = 0; m < NM; ++m) {
      const int col = 2 * (lane + 32 * m);
      g2[m] = (col < dw) ? make_float2(dout[(int64_t)it * dw + col] * inv, dout[(int64_t)it * dw + col + 1] * inv) : make_float2(0.f, 0.f);
    }
    auto add_row = [&](float* wr, float scale) {
#pragma unroll
      for (int m = 0; m < NM; ++m) {
        const int col = 2 * (lane + 32 * m);
        if (col < dw)
          asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(wr + col), "f"(g2[m].x * scale), "f"(g2[m].y * scale) : "memory");
      }
    };
    if (cnt0 > 0 && warp == 0) add_row(dW, static_cast<float>(cnt0));
    for (int l0 = warp * 32; l0 < L; l0 += 256) {
      const int32_t mine = (l0 + lane < L) ? __ldg(c + l0 + lane) : -1;
      for (int t = 0; t < 32; ++t) {
        const int32_t w = __shfl_sync(0xffffffffu, mine, t);
        if (w <= 0) continue;
        add_row(dW + (int64_t)w * dw, 1.0f);
      }
    }
    return;
  }
  float g[NM];
#pragma unroll
  for (int m = 0; m < NM; ++m) {
    const int col = lane + 32 * m;
    g[m] = (col < dw) ? dout[(int64_t)it * dw + col] * inv : 0.0f;
  }
  if (cnt0 > 0 && warp == 0) {
#pragma unroll
    for (int m = 0; m < NM; ++m) {
      const int col = lane + 32 * m;
      if (col < dw) atomicAdd(dW + col, g[m] * static_cast<float>(cnt0));
    }
  }
  for (int l0 = warp * 32; l0 < L; l0 += 256) {
    const int32_t mine = (l0 + lane < L) ? __ldg(c + l0 + lane) : -1;
    for (int t = 0; t < 32; ++t) {
      const int32_t w = __shfl_sync(0xffffffffu, mine, t);
      if (w <= 0) continue;
      float* wr = dW + (int64_t)w * dw;
#pragma unroll
      for (int m = 0; m < NM; ++m) {
        const int col = lane + 32 * m;
        if (col < dw) atomicAdd(wr + col, g[m]);
      }
    }
  }
}

}  // namespace nncf

using namespace nncf;

static int meanpool_fwd_impl(const float* word_table_dev, int word_dim, const int32_t* content_dev, int content_len,
                             const int32_t* item_ids_dev, int n_items, float* out_dev, const int32_t* n_valid_dev, void* stream) {
  NNCF_CHECK_ARG(n_items >= 0, "nncf_meanpool_fwd: n_items < 0");
  if (n_items == 0) return NNCF_OK;
  NNCF_CHECK_ARG(word_table_dev && content_dev && out_dev, "nncf_meanpool_fwd: null argument");
  NNCF_CHECK_ARG(word_dim >= 1 && word_dim <= 256, "nncf_meanpool_fwd: word_dim must be in [1, 256]");
  NNCF_CHECK_ARG(content_len >= 1, "nncf_meanpool_fwd: content_len must be >= 1");
  const int nm = (word_dim + 31) / 32;
#define NNCF_MP_FWD(NM) meanpool_fwd_kernel<NM><<<n_items, 256, 0, (cudaStream_t)stream>>>(word_table_dev, word_dim, content_dev, content_len, item_ids_dev, n_items, out_dev, n_valid_dev)
  if (nm <= 1) NNCF_MP_FWD(1); else if (nm <= 2) NNCF_MP_FWD(2); else if (nm <= 4) NNCF_MP_FWD(4); else NNCF_MP_FWD(8);
#undef NNCF_MP_FWD
  NNCF_LAUNCH_OK();
  return NNCF_OK;
}

static int meanpool_bwd_impl(float* grad_word_table_dev, int word_dim, const int32_t* content_dev, int content_len,
                             const int32_t* item_ids_dev, int n_items, const float* grad_out_dev, const int32_t* n_valid_dev, void* stream) {
  NNCF_CHECK_ARG(n_items >= 0, "nncf_meanpool_bwd: n_items < 0");
  if (n_items == 0) return NNCF_OK;
  NNCF_CHECK_ARG(grad_word_table_dev && content_dev && grad_out_dev, "nncf_meanpool_bwd: null argument");
  NNCF_CHECK_ARG(word_dim >= 1 && word_dim <= 256, "nncf_meanpool_bwd: word_dim must be in [1, 256]");
  NNCF_CHECK_ARG(content_len >= 1, "nncf_meanpool_bwd: content_len must be >= 1");
  const int nm = (word_dim + 31) / 32;
#define NNCF_MP_BWD(NM) meanpool_bwd_kernel<NM><<<n_items, 256, 0, (cudaStream_t)stream>>>(grad_word_table_dev, word_dim, content_dev, content_len, item_ids_dev, n_items, grad_out_dev, n_valid_dev)
  if (nm <= 1) NNCF_MP_BWD(1); else if (nm <= 2) NNCF_MP_BWD(2); else if (nm <= 4) NNCF_MP_BWD(4); else NNCF_MP_BWD(8);
#undef NNCF_MP_BWD
  NNCF_LAUNCH_OK();
  return NNCF_OK;
}

extern "C" int nncf_meanpool_fwd(const float* word_table_dev, int word_dim, const int32_t* content_dev, int content_len,
                                 const int32_t* item_ids_dev, int n_items, float* out_dev, void* stream) {
  return meanpool_fwd_impl(word_table_dev, word_dim, content_dev, content_len, item_ids_dev, n_items, out_dev, nullptr, stream);
}
extern "C" int nncf_meanpool_bwd(float* grad_word_table_dev, int word_dim, const int32_t* content_dev, int content_len,
                                 const int32_t* item_ids_dev, int n_items, const float* grad_out_dev, void* stream) {
  return meanpool_bwd_impl(grad_word_table_dev, word_dim, content_dev, content_len, item_ids_dev, n_items, grad_out_dev, nullptr, stream);
}
extern "C" int nncf_meanpool_fwd_n(const float* word_table_dev, int word_dim, const int32_t* content_dev, int content_len,
                                   const int32_t* item_ids_dev, int n_slots, const int32_t* n_valid_dev, float* out_dev, void* stream) {
  return meanpool_fwd_impl(word_table_dev, word_dim, content_dev, content_len, item_ids_dev, n_slots, out_dev, n_valid_dev, stream);
}
extern "C" int nncf_meanpool_bwd_n(float* grad_word_table_dev, int word_dim, const int32_t* content_dev, int content_len,
                                   const int32_t* item_ids_dev, int n_slots, const int32_t* n_valid_dev, const float* grad_out_dev,
                                   void* stream) {
  return meanpool_bwd_impl(grad_word_table_dev, word_dim, content_dev, content_len, item_ids_dev, n_slots, grad_out_dev, n_valid_dev, stream);
}

// =================================================================================================
// Dense-transform tail of the content towers on the block of UNIQUE items: BatchNorm (batch statistics over the first
// n = *n_valid rows, Keras defaults eps 1e-3 / momentum 0.99, ref: modules/content/mean_pool.py:90-95) + activation,
// forward and backward, with the row count in DEVICE memory so that a whole training step can be replayed as a CUDA
// graph (nncf_b200/model_framework.py: MeanPoolGraphStep).  One CTA per 32 columns; 8 warps stride over the rows.
// =================================================================================================
namespace nncf {

// Thread layout of the two tower kernels: a CTA of 256 threads takes kBnCols columns with 256 / kBnCols row lanes.  The first
// version took 32 columns with 8 row lanes: TWO CTAs for the d = 50 tower, each thread walking 64 rows three times (13.8 us
// forward, 40.5 us backward for 25,600 elements).  8 columns x 32 row lanes: 7 CTAs, 16 rows per thread.
constexpr int kBnCols = 8;
constexpr int kBnLanes = 256 / kBnCols;
__device__ __forceinline__ float cta_col_sum(float v, float (*red)[kBnCols + 1], int ry, int cx) {
  red[ry][cx] = v;
  __syncthreads();
  float t = 0.0f;
#pragma unroll
  for (int w = 0; w < kBnLanes; ++w) t += red[w][cx];                 // (fixed order: the result does not depend on scheduling)
  __syncthreads();
  return t;
}

// act: 0 = linear, 1 = relu, 2 = tanh.  use_bn = 0: activation only.
__global__ void __launch_bounds__(256)
tower_bn_act_fwd_kernel(const float* __restrict__ h, int rows, int d, const int32_t* __restrict__ n_valid, int use_bn, int act,
                        const float* __restrict__ gamma, const float* __restrict__ beta, float eps, float momentum,
                        float* running_mean, float* running_var, float* __restrict__ y, float* __restrict__ xhat,
                        float* __restrict__ rstd_out) {
  __shared__ float red[kBnLanes][kBnCols + 1];
  const int cx = threadIdx.x % kBnCols, ry = threadIdx.x / kBnCols;
  const int c = blockIdx.x * kBnCols + cx;
  const bool ok = c < d;
  const int n = min(*n_valid, rows);
  float mean = 0.0f, rstd = 1.0f, g = 1.0f, b = 0.0f;
  if (use_bn) {
    float s = 0.0f;
    if (ok) for (int r = ry; r < n; r += kBnLanes) s += h[(int64_t)r * d + c];
    mean = cta_col_sum(s, red, ry, cx) / static_cast<float>(n);
    float q = 0.0f;
    if (ok) for (int r = ry; r < n; r += kBnLanes) { const float t = h[(int64_t)r * d + c] - mean; q = fmaf(t, t, q); }
    const float var = cta_col_sum(q, red, ry, cx) / static_cast<float>(n);          // biased: what normalises the batch
    rstd = rsqrtf(var + eps);
    if (ok) { g = gamma[c]; b = beta[c]; }
    if (ok && ry == 0) {
      rstd_out[c] = rstd;
      running_mean[c] = (1.0f - momentum) * running_mean[c] + momentum * mean;
      running_var[c] = (1.0f - momentum) * running_var[c] + momentum * var * (static_cast<float>(n) / fmaxf(static_cast<float>(n) - 1.0f, 1.0f));
    }
  }
  if (!ok) return;
  for (int r = ry; r < rows; r += kBnLanes) {
    const int64_t o = (int64_t)r * d + c;
    float xh = 0.0f, v = 0.0f;
    if (r < n) {
      xh = use_bn ? (h[o] - mean) * rstd : h[o];
      v = use_bn ? fmaf(xh, g, b) : xh;
      v = act == 1 ? fmaxf(v, 0.0f) : (act == 2 ? tanhf(v) : v);
    }
    y[o] = v;                                           // rows beyond n: zeros
    xhat[o] = xh;
  }
}

// dy -> dh (through activation and BatchNorm), dgamma, dbeta.  y / xhat / rstd are the forward's outputs.
__global__ void __launch_bounds__(256)
tower_bn_act_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y, const float* __restrict__ xhat,
                        const float* __restrict__ rstd_in, int rows, int d, const int32_t* __restrict__ n_valid, int use_bn, int act,
                        const float* __restrict__ gamma, float* __restrict__ dh, float* __restrict__ dgamma, float* __restrict__ dbeta) {
  __shared__ float red[kBnLanes][kBnCols + 1];
  const int cx = threadIdx.x % kBnCols, ry = threadIdx.x / kBnCols;
  const int c = blockIdx.x * kBnCols + cx;
  const bool ok = c < d;
  const int n = min(*n_valid, rows);
  auto act_grad = [&](int64_t o) {
    const float g = dy[o], v = y[o];
    return act == 1 ? (v > 0.0f ? g : 0.0f) : (act == 2 ? g * (1.0f - v * v) : g);
  };
  float s1 = 0.0f, s2 = 0.0f;
  if (ok && use_bn)
    for (int r = ry; r < n; r += kBnLanes) { const int64_t o = (int64_t)r * d + c; const float g2 = act_grad(o); s1 += g2; s2 = fmaf(g2, xhat[o], s2); }
  const float sum_g = cta_col_sum(s1, red, ry, cx), sum_gx = cta_col_sum(s2, red, ry, cx);
  if (!ok) return;
  const float gm = use_bn ? gamma[c] : 1.0f, rs = use_bn ? rstd_in[c] : 1.0f, inv_n = 1.0f / static_cast<float>(n);
  if (use_bn && ry == 0) { dgamma[c] = sum_gx; dbeta[c] = sum_g; }
  for (int r = ry; r < rows; r += kBnLanes) {
    const int64_t o = (int64_t)r * d + c;
    float v = 0.0f;
    if (r < n) {
      const float g2 = act_grad(o);
      v = use_bn ? rs * gm * (g2 - sum_g * inv_n - xhat[o] * sum_gx * inv_n) : g2;
    }
    dh[o] = v;
  }
}

// ---- dense Adam for the tower's parameters, Keras-1 form (ref: utils/optimizer.py:108-147 is the reference's copy of it):
//   lr_t = lr sqrt(1 - beta2^t) / (1 - beta1^t);  m = beta1 m + (1 - beta1) g;  v = beta2 v + (1 - beta2) g^2;
//   p -= lr_t m / (sqrt(v) + eps)
// One launch for up to kDenseAdamMax tensors (blockIdx.y = tensor); the step count lives in device memory and is advanced
// by a one-thread kernel in front, so a call captured into a CUDA graph stays correct when replayed.  (torch's fused
// multi-tensor Adam spent 38 us per step on the 400k-element word table: 7 CTAs.)
constexpr int kDenseAdamMax = 8;
struct DenseAdamSegs {
  float* p[kDenseAdamMax]; const float* g[kDenseAdamMax]; float* m[kDenseAdamMax]; float* v[kDenseAdamMax];
  long long n[kDenseAdamMax];
};
__global__ void dense_adam_tick_kernel(long long* step, float* lr_out, double lr, double beta1, double beta2) {
  const long long t = ++(*step);
  *lr_out = static_cast<float>(lr * sqrt(1.0 - pow(beta2, static_cast<double>(t))) / (1.0 - pow(beta1, static_cast<double>(t))));
}
__global__ void __launch_bounds__(256)
dense_adam_kernel(DenseAdamSegs s, const float* __restrict__ lr_dev, float beta1, float beta2, float eps) {
  const int k = blockIdx.y;
  const long long n = s.n[k];
  const float lr_t = *lr_dev, c1 = 1.0f - beta1, c2 = 1.0f - beta2;
  float* p = s.p[k]; const float* g = s.g[k]; float* m = s.m[k]; float* v = s.v[k];
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
    const float gi = g[i];
    const float mi = beta1 * m[i] + c1 * gi, vi = beta2 * v[i] + c2 * gi * gi;
    m[i] = mi; v[i] = vi;
    p[i] -= lr_t * mi / (sqrtf(vi) + eps);
  }
}

}  // namespace nncf

extern "C" int nncf_dense_adam_step(int n_tensors, float* const* params_dev, const float* const* grads_dev, float* const* m_dev,
                                    float* const* v_dev, const int64_t* sizes, float lr, float beta1, float beta2, float eps,
                                    long long* step_dev, float* lr_t_dev, void* stream) {
  NNCF_CHECK_ARG(n_tensors >= 0 && params_dev && grads_dev && m_dev && v_dev && sizes && step_dev && lr_t_dev, "nncf_dense_adam_step: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  dense_adam_tick_kernel<<<1, 1, 0, st>>>(step_dev, lr_t_dev, (double)lr, (double)beta1, (double)beta2);
  NNCF_LAUNCH_OK();
  for (int k0 = 0; k0 < n_tensors; k0 += kDenseAdamMax) {
    DenseAdamSegs s{};
    const int cnt = n_tensors - k0 < kDenseAdamMax ? n_tensors - k0 : kDenseAdamMax;
    long long big = 0;
    for (int k = 0; k < cnt; ++k) {
      NNCF_CHECK_ARG(params_dev[k0 + k] && grads_dev[k0 + k] && m_dev[k0 + k] && v_dev[k0 + k] && sizes[k0 + k] >= 0, "nncf_dense_adam_step: bad tensor");
      s.p[k] = params_dev[k0 + k]; s.g[k] = grads_dev[k0 + k]; s.m[k] = m_dev[k0 + k]; s.v[k] = v_dev[k0 + k]; s.n[k] = sizes[k0 + k];
      big = sizes[k0 + k] > big ? sizes[k0 + k] : big;
    }
    if (big == 0) continue;
    const long long blocks = (big + 1023) / 1024;                      // ~4 elements per thread of the largest tensor
    dense_adam_kernel<<<dim3((unsigned)(blocks < 4096 ? blocks : 4096), cnt), 256, 0, st>>>(s, lr_t_dev, beta1, beta2, eps);
    NNCF_LAUNCH_OK();
  }
  return NNCF_OK;
}

extern "C" int nncf_tower_bn_act_fwd(const float* h_dev, int rows, int dim, const int32_t* n_valid_dev, int use_bn, int activation,
                                     const float* gamma_dev, const float* beta_dev, float eps, float momentum,
                                     float* running_mean_dev, float* running_var_dev, float* y_dev, float* xhat_dev,
                                     float* rstd_dev, void* stream) {
  NNCF_CHECK_ARG(h_dev && n_valid_dev && y_dev && xhat_dev, "nncf_tower_bn_act_fwd: null argument");
  NNCF_CHECK_ARG(rows >= 1 && dim >= 1 && activation >= 0 && activation <= 2, "nncf_tower_bn_act_fwd: bad sizes");
  if (use_bn) NNCF_CHECK_ARG(gamma_dev && beta_dev && running_mean_dev && running_var_dev && rstd_dev, "nncf_tower_bn_act_fwd: BatchNorm needs its parameters");
  tower_bn_act_fwd_kernel<<<ceil_div(dim, kBnCols), 256, 0, (cudaStream_t)stream>>>(h_dev, rows, dim, n_valid_dev, use_bn, activation, gamma_dev,
                                                                             beta_dev, eps, momentum, running_mean_dev, running_var_dev,
                                                                             y_dev, xhat_dev, rstd_dev);
  NNCF_LAUNCH_OK();
  return NNCF_OK;
}
extern "C" int nncf_tower_bn_act_bwd(const float* dy_dev, const float* y_dev, const float* xhat_dev, const float* rstd_dev, int rows,
                                     int dim, const int32_t* n_valid_dev, int use_bn, int activation, const float* gamma_dev,
                                     float* dh_dev, float* dgamma_dev, float* dbeta_dev, void* stream) {
  NNCF_CHECK_ARG(dy_dev && y_dev && xhat_dev && n_valid_dev && dh_dev, "nncf_tower_bn_act_bwd: null argument");
  NNCF_CHECK_ARG(rows >= 1 && dim >= 1 && activation >= 0 && activation <= 2, "nncf_tower_bn_act_bwd: bad sizes");
  if (use_bn) NNCF_CHECK_ARG(gamma_dev && rstd_dev && dgamma_dev && dbeta_dev, "nncf_tower_bn_act_bwd: BatchNorm needs its parameters");
  tower_bn_act_bwd_kernel<<<ceil_div(dim, kBnCols), 256, 0, (cudaStream_t)stream>>>(dy_dev, y_dev, xhat_dev, rstd_dev, rows, dim, n_valid_dev, use_bn,
                                                                             activation, gamma_dev, dh_dev, dgamma_dev, dbeta_dev);
  NNCF_LAUNCH_OK();
  return NNCF_OK;
}
