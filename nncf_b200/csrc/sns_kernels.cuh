// sns_kernels.cuh — the 'sampled_neg_shared' view: B positive (user, item) rows followed by k sampled items that are
// shared as negatives by every positive of the batch.
//   ref: models/model_framework.py:138-143 (pred = [mul(U_front, C_front), matmul(U_front, C_back^T)] : [B, 1+k]),
//        utils/objectives.py:120-161 (get_sampled_neg_shared_loss), models/train_sampled_neg_shared.py:42-50 (batch
//        layout: the k back rows carry user id 0 from np.zeros((k, 3))).
// k is small (the demos use 10..20), so there is no GEMM shape in it: fp32, one warp per positive row, the k negative
// rows stay in L1/L2; the per-negative gradient sums are accumulated per CTA in shared memory first.
// Included by train_step.cu (uses its helpers).
#pragma once

namespace nncf {

struct SnsArgs {
  const float* EU; const float* EV;
  const int32_t* uid; const int32_t* cid;   // [R][B + k]
  int B, k, d;
  int norm_u, norm_v, loss_kind;
  float lambda, gamma, u_reg;
  double* loss;                // [R]
  float* dUrows; float* dVrows;   // [R][B + k][d] row gradients w.r.t. the RAW gathered rows
  float* dVn_hat;              // [R][k][d] accumulated dL/d(vhat_neg_j), zeroed by the caller
  float* grad_out_u; float* grad_out_v;   // optional [B + k][d] copies of replica 0
};

// one warp per positive row i
__global__ void __launch_bounds__(256)
sns_main_kernel(SnsArgs a) {
  extern __shared__ float s_acc[];                 // [k][d] CTA-local accumulators of dL/d(vhat_neg)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int B = a.B, k = a.k, d = a.d, n = B + k, r = blockIdx.y;
  for (int i = threadIdx.x; i < k * d; i += blockDim.x) s_acc[i] = 0.0f;
  __syncthreads();
  const int row = blockIdx.x * 8 + warp;
  if (row < B) {
    const int32_t* uid = a.uid + (int64_t)r * n;
    const int32_t* cid = a.cid + (int64_t)r * n;
    const float* u = a.EU + (int64_t)uid[row] * d;
    const float* v = a.EV + (int64_t)cid[row] * d;
    float su = 0.0f, sv = 0.0f, dot = 0.0f;
    for (int c = lane; c < d; c += 32) {
      const float x = __ldg(u + c), y = __ldg(v + c);
      su = fmaf(x, x, su); sv = fmaf(y, y, sv); dot = fmaf(x, y, dot);
    }
    su = warp_sum(su); sv = warp_sum(sv); dot = warp_sum(dot);
    const float iu = a.norm_u ? rsqrtf(fmaxf(su, 1e-12f)) : 1.0f;
    const float iv = a.norm_v ? rsqrtf(fmaxf(sv, 1e-12f)) : 1.0f;
    const float p0 = dot * iu * iv;
    const float invB = 1.0f / B, w = a.lambda / k, inv_cnt = 1.0f / (static_cast<float>(B) * k);
    const bool pairwise = a.loss_kind >= NNCF_LOSS_LOG_LOSS;
    float l = 0.0f, g0 = 0.0f;
    if (a.loss_kind == NNCF_LOSS_SKIP_GRAM) { l = softplus_f<false>(-p0) * invB; g0 = (sigmoid_f<false>(p0) - 1.0f) * invB; }
    else if (a.loss_kind == NNCF_LOSS_MSE) { l = (p0 - 1.0f) * (p0 - 1.0f) * invB; g0 = 2.0f * (p0 - 1.0f) * invB; }
    // dL/d(uhat_i) accumulates in registers: lane holds columns lane, lane + 32, ... (d <= 256 -> at most 8)
    float du[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) du[q] = 0.0f;
    for (int j = 0; j < k; ++j) {
      const float* vn = a.EV + (int64_t)cid[B + j] * d;
      float sn = 0.0f, dn = 0.0f;
      for (int c = lane; c < d; c += 32) {
        const float x = __ldg(u + c), y = __ldg(vn + c);
        sn = fmaf(y, y, sn); dn = fmaf(x, y, dn);
      }
      sn = warp_sum(sn); dn = warp_sum(dn);
      const float in = a.norm_v ? rsqrtf(fmaxf(sn, 1e-12f)) : 1.0f;
      const float pn = dn * iu * in;
      float gn;
      if (a.loss_kind == NNCF_LOSS_SKIP_GRAM) { l += w * softplus_f<false>(pn) * invB; gn = w * sigmoid_f<false>(pn) * invB; }
      else if (a.loss_kind == NNCF_LOSS_MSE) { l += w * pn * pn * invB; gn = 2.0f * w * pn * invB; }
      else {
        const float D = p0 - pn;
        float aa;
        if (a.loss_kind == NNCF_LOSS_LOG_LOSS) { l += softplus_f<false>(-a.gamma * D) * inv_cnt; aa = -a.gamma * sigmoid_f<false>(-a.gamma * D) * inv_cnt; }
        else { l += fmaxf(a.gamma - D, 0.0f) * inv_cnt; aa = (a.gamma - D > 0.0f) ? -inv_cnt : 0.0f; }
        g0 += aa; gn = -aa;
      }
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int c = lane + 32 * q;
        if (c < d) {
          du[q] = fmaf(gn, __ldg(vn + c) * in, du[q]);
          atomicAdd(&s_acc[j * d + c], gn * __ldg(u + c) * iu);
        }
      }
    }
    (void)pairwise;
    // positive pair, then normalise-backward and the activity regulariser on the raw user row
    const int64_t o = ((int64_t)r * n + row) * d;
    float uhat_dot = 0.0f;                        // uhat . dL/duhat
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int c = lane + 32 * q;
      if (c < d) {
        du[q] = fmaf(g0, __ldg(v + c) * iv, du[q]);
        uhat_dot = fmaf(__ldg(u + c) * iu, du[q], uhat_dot);
      }
    }
    uhat_dot = warp_sum(uhat_dot);
    const float reg = 2.0f * a.u_reg / n;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int c = lane + 32 * q;
      if (c < d) {
        const float x = __ldg(u + c), y = __ldg(v + c);
        float gu = du[q];
        if (a.norm_u) gu = (gu - x * iu * uhat_dot) * iu;
        gu = fmaf(reg, x, gu);
        float gv = g0 * x * iu;                   // dL/d(vhat_i)
        if (a.norm_v) gv = (gv - y * iv * (g0 * p0)) * iv;
        a.dUrows[o + c] = gu;
        a.dVrows[o + c] = gv;
        if (a.grad_out_u && r == 0) a.grad_out_u[(int64_t)row * d + c] = gu;
        if (a.grad_out_v && r == 0) a.grad_out_v[(int64_t)row * d + c] = gv;
      }
    }
    if (lane == 0) {
      if (a.u_reg != 0.0f) l += a.u_reg * su / n;
      atomicAdd(&a.loss[r], static_cast<double>(l));
    }
  }
  __syncthreads();
  float* acc = a.dVn_hat + (int64_t)r * k * d;
  for (int i = threadIdx.x; i < k * d; i += blockDim.x) {
    const float x = s_acc[i];
    if (x != 0.0f) atomicAdd(acc + i, x);
  }
}

// the k back rows: item gradient = normalise-backward of the accumulated dL/d(vhat_neg_j); user gradient = the
// regulariser on the dummy user row (Keras regularises the Embedding output of ALL B + k inputs)
__global__ void __launch_bounds__(256)
sns_back_kernel(SnsArgs a) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int B = a.B, k = a.k, d = a.d, n = B + k, r = blockIdx.y;
  const int j = blockIdx.x * 8 + warp;
  if (j >= k) return;
  const int32_t uidj = a.uid[(int64_t)r * n + B + j], cidj = a.cid[(int64_t)r * n + B + j];
  const float* u = a.EU + (int64_t)uidj * d;
  const float* v = a.EV + (int64_t)cidj * d;
  const float* g = a.dVn_hat + ((int64_t)r * k + j) * d;
  float sv = 0.0f, su = 0.0f, vg = 0.0f;
  for (int c = lane; c < d; c += 32) {
    const float y = __ldg(v + c), x = __ldg(u + c);
    sv = fmaf(y, y, sv); su = fmaf(x, x, su); vg = fmaf(y, g[c], vg);
  }
  sv = warp_sum(sv); su = warp_sum(su); vg = warp_sum(vg);
  const float in = a.norm_v ? rsqrtf(fmaxf(sv, 1e-12f)) : 1.0f;
  const float reg = 2.0f * a.u_reg / n;
  const int64_t o = ((int64_t)r * n + B + j) * d;
  for (int c = lane; c < d; c += 32) {
    float gv = g[c];
    if (a.norm_v) gv = (gv - __ldg(v + c) * in * (vg * in)) * in;
    const float gu = reg * __ldg(u + c);
    a.dVrows[o + c] = gv;
    a.dUrows[o + c] = gu;
    if (a.grad_out_u && r == 0) a.grad_out_u[(int64_t)(B + j) * d + c] = gu;
    if (a.grad_out_v && r == 0) a.grad_out_v[(int64_t)(B + j) * d + c] = gv;
  }
  if (lane == 0 && a.u_reg != 0.0f) atomicAdd(&a.loss[r], static_cast<double>(a.u_reg * su / n));
}

}  // namespace nncf
