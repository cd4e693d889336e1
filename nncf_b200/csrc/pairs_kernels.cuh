// pairs_kernels.cuh — vectorised kernels of the PAIRS scheme ('original' / 'group_sample': row-wise dot on (1+k)B listed
// pairs, ref: modules/interaction/interaction_dot.py:92-99, utils/objectives.py:35-75) for 16-byte aligned rows without
// interaction-bias columns.  Included by train_step.cu after PairsArgs and the scalar kernels (which remain the general path).
//
// Why: ncu of the scalar kernels (profiles/r02_kernels_summary.md: 320 / 360 us for 208k pairs = 0.10 / 0.16 of the HBM
// roofline) showed one atomicAdd(double) per PAIR on the replica's loss word - 5,600 serialised L2 atomics per address -
// and a single pair of 512-byte rows in flight per warp behind two dependent loads (id -> row).  Here a warp takes FOUR
// pairs (ids first, then eight LDG.128 per lane in flight), and a CTA adds its loss share with one atomic.
#pragma once

namespace nncf {

constexpr int kPairsPerWarp = 4;

__device__ __forceinline__ void cta_loss_add(double* dst, float warp_value, float* sm8) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) sm8[warp] = warp_value;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.0f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += sm8[w];
    if (t != 0.0f) atomicAdd(dst, static_cast<double>(t));
  }
}

template <int NV>   // float4 chunks per lane: 1 for d <= 128, 2 for d <= 256
__global__ void __launch_bounds__(256)
pairs_score_vec_kernel(PairsArgs a) {
  __shared__ float sm8[8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = (1 + a.k) * a.B, r = blockIdx.y;
  const int row0 = (blockIdx.x * 8 + warp) * kPairsPerWarp;
  int64_t myu = -1, myc = -1;
  if (lane < kPairsPerWarp && row0 + lane < n) { myu = a.uid[r * a.ids_stride + row0 + lane]; myc = a.cid[r * a.ids_stride + row0 + lane]; }
  float4 x[kPairsPerWarp][NV], y[kPairsPerWarp][NV];
#pragma unroll
  for (int k = 0; k < kPairsPerWarp; ++k) {
    const int64_t iu = __shfl_sync(0xffffffffu, myu, k), ic = __shfl_sync(0xffffffffu, myc, k);
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const int c = 4 * (lane + 32 * v);
      const bool ok = iu >= 0 && c < a.d;
      x[k][v] = ok ? __ldg(reinterpret_cast<const float4*>(a.EU + iu * a.d + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
      y[k][v] = ok ? __ldg(reinterpret_cast<const float4*>(a.EV + ic * a.d + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  float reg = 0.0f;
#pragma unroll
  for (int k = 0; k < kPairsPerWarp; ++k) {
    float su = 0.0f, sv = 0.0f, dot = 0.0f;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      su += x[k][v].x * x[k][v].x + x[k][v].y * x[k][v].y + x[k][v].z * x[k][v].z + x[k][v].w * x[k][v].w;
      sv += y[k][v].x * y[k][v].x + y[k][v].y * y[k][v].y + y[k][v].z * y[k][v].z + y[k][v].w * y[k][v].w;
      dot += x[k][v].x * y[k][v].x + x[k][v].y * y[k][v].y + x[k][v].z * y[k][v].z + x[k][v].w * y[k][v].w;
    }
    su = warp_sum(su); sv = warp_sum(sv); dot = warp_sum(dot);
    const float iu = a.norm_u ? rsqrtf(fmaxf(su, 1e-12f)) : 1.0f;
    const float iv = a.norm_v ? rsqrtf(fmaxf(sv, 1e-12f)) : 1.0f;
    if (lane == k && row0 + k < n) {
      const int64_t o = (int64_t)r * n + row0 + k;
      a.s[o] = dot * iu * iv;
      a.invu[o] = iu;
      a.invv[o] = iv;
    }
    if (row0 + k < n) reg += su;
  }
  cta_loss_add(&a.loss[r], a.u_reg != 0.0f ? a.u_reg * reg / n : 0.0f, sm8);       // ref: utils/utilities.py:129-135
}

template <int NV>
__global__ void __launch_bounds__(256)
pairs_grad_vec_kernel(PairsArgs a) {
  __shared__ float sm8[8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int B = a.B, k = a.k, n = (1 + k) * B, r = blockIdx.y;
  const int row0 = (blockIdx.x * 8 + warp) * kPairsPerWarp;
  const float* S = a.s + (int64_t)r * n;
  const bool pointwise = a.loss_kind <= NNCF_LOSS_MSE;
  const float invB = 1.0f / B, w = a.lambda / k;
  // rows first (ids -> rows are two dependent loads): the loss / dL/ds arithmetic below runs while they are in flight
  int64_t myu = -1, myc = -1;
  if (lane < kPairsPerWarp && row0 + lane < n) { myu = a.uid[r * a.ids_stride + row0 + lane]; myc = a.cid[r * a.ids_stride + row0 + lane]; }
  float4 x[kPairsPerWarp][NV], y[kPairsPerWarp][NV];
#pragma unroll
  for (int q = 0; q < kPairsPerWarp; ++q) {
    const int64_t iu = __shfl_sync(0xffffffffu, myu, q), ic = __shfl_sync(0xffffffffu, myc, q);
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const int c = 4 * (lane + 32 * v);
      const bool ok = iu >= 0 && c < a.d;
      x[q][v] = ok ? __ldg(reinterpret_cast<const float4*>(a.EU + iu * a.d + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
      y[q][v] = ok ? __ldg(reinterpret_cast<const float4*>(a.EV + ic * a.d + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  float lsum = 0.0f;
  float gq[kPairsPerWarp];
#pragma unroll
  for (int q = 0; q < kPairsPerWarp; ++q) {
    const int row = row0 + q;
    float g = 0.0f, l = 0.0f;
    if (row < n) {                                                     // (warp-uniform)
      const float s = S[row];
      const bool is_pos = (pointwise && a.resp) ? (a.resp[(int64_t)r * n + row] == 1) : (row < B);
      if (a.loss_kind == NNCF_LOSS_SKIP_GRAM) {
        if (is_pos) { l = softplus_f<false>(-s) * invB; g = (sigmoid_f<false>(s) - 1.0f) * invB; }
        else { l = w * softplus_f<false>(s) * invB; g = w * sigmoid_f<false>(s) * invB; }
      } else if (a.loss_kind == NNCF_LOSS_MSE) {
        if (is_pos) { l = (1.0f - s) * (1.0f - s) * invB; g = -2.0f * (1.0f - s) * invB; }
        else { l = w * s * s * invB; g = 2.0f * w * s * invB; }
      } else {
        const float inv_cnt = 1.0f / (static_cast<float>(k) * B);
        if (is_pos) {                                                  // g+ = sum over this positive's k negatives of a_n
          float acc = 0.0f;
          for (int t = lane; t < k; t += 32) {
            const float dd = s - S[B + row * k + t];
            if (a.loss_kind == NNCF_LOSS_LOG_LOSS) acc += -a.gamma * sigmoid_f<false>(-a.gamma * dd) * inv_cnt;
            else acc += (a.gamma - dd > 0.0f) ? -inv_cnt : 0.0f;
          }
          g = warp_sum(acc);
        } else {
          const float dd = S[(row - B) / k] - s;
          if (a.loss_kind == NNCF_LOSS_LOG_LOSS) {
            l = softplus_f<false>(-a.gamma * dd) * inv_cnt;
            g = a.gamma * sigmoid_f<false>(-a.gamma * dd) * inv_cnt;
          } else {
            l = fmaxf(a.gamma - dd, 0.0f) * inv_cnt;
            g = (a.gamma - dd > 0.0f) ? inv_cnt : 0.0f;
          }
        }
      }
    }
    gq[q] = g;
    lsum += l;                                                         // (every lane holds the same l: counted once below)
  }
  const float reg = 2.0f * a.u_reg / n;
#pragma unroll
  for (int q = 0; q < kPairsPerWarp; ++q) {
    const int row = row0 + q;
    if (row >= n) break;
    const int64_t o = (int64_t)r * n + row;
    const float iu = a.invu[o], iv = a.invv[o], g = gq[q];
    const float gs = g * S[row];                                       // xhat . dxhat = g * s on both sides
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const int c = 4 * (lane + 32 * v);
      if (c >= a.d) continue;
      const float xs[4] = {x[q][v].x, x[q][v].y, x[q][v].z, x[q][v].w}, ys[4] = {y[q][v].x, y[q][v].y, y[q][v].z, y[q][v].w};
      float du[4], dv[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float xh = xs[e] * iu, yh = ys[e] * iv;
        du[e] = g * yh; dv[e] = g * xh;
        if (a.norm_u) du[e] = (du[e] - xh * gs) * iu;
        if (a.norm_v) dv[e] = (dv[e] - yh * gs) * iv;
        du[e] = fmaf(reg, xs[e], du[e]);
      }
      const float4 du4 = make_float4(du[0], du[1], du[2], du[3]), dv4 = make_float4(dv[0], dv[1], dv[2], dv[3]);
      if (a.dUrows) {
        *reinterpret_cast<float4*>(a.dUrows + o * a.d + c) = du4;
        *reinterpret_cast<float4*>(a.dVrows + o * a.d + c) = dv4;
      }
      if (a.grad_out_u && r == 0) *reinterpret_cast<float4*>(a.grad_out_u + (int64_t)row * a.d + c) = du4;
      if (a.grad_out_v && r == 0) *reinterpret_cast<float4*>(a.grad_out_v + (int64_t)row * a.d + c) = dv4;
    }
  }
  cta_loss_add(&a.loss[r], lsum, sm8);
}

// sparse SGD of the stored row gradients (second pass: every gradient of the step was taken on one snapshot of the tables)
template <int NV>
__global__ void __launch_bounds__(256)
rows_sgd_vec_kernel(const float* __restrict__ rows, const int32_t* __restrict__ ids, int64_t ids_stride, int n, int d, float lr,
                    float* table) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row0 = (blockIdx.x * 8 + warp) * kPairsPerWarp, r = blockIdx.y;
  int64_t myid = -1;
  if (lane < kPairsPerWarp && row0 + lane < n) myid = ids[r * ids_stride + row0 + lane];
  float4 g[kPairsPerWarp][NV];
#pragma unroll
  for (int q = 0; q < kPairsPerWarp; ++q)
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const int c = 4 * (lane + 32 * v);
      g[q][v] = (row0 + q < n && c < d) ? *reinterpret_cast<const float4*>(rows + ((int64_t)r * n + row0 + q) * d + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
  for (int q = 0; q < kPairsPerWarp; ++q) {
    const int64_t id = __shfl_sync(0xffffffffu, myid, q);
    if (id < 0) continue;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const int c = 4 * (lane + 32 * v);
      if (c < d) red_add_v4(table + id * d + c, -lr * g[q][v].x, -lr * g[q][v].y, -lr * g[q][v].z, -lr * g[q][v].w);
    }
  }
}

}  // namespace nncf
