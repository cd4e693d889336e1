// train_step.cu — fused training step of the sampling-and-scoring loop.
//
// One step on R independent batches (replicas):
//   [unique_kernel]      group_neg_shared only: tf.unique (first-occurrence order) of the item ids
//   gather_rows_kernel   ids -> rows: fp32 staging (l2-normalised if asked), bf16 "tile image" for the tensor
//                        cores, 1/||x||, zeroed gradient accumulators                       (HBM-bound part)
//   [pos_score_kernel]   pairwise losses: positive score per column / per row
//   score_grad_*         S = U V^T tile by tile, loss + dL/dS in the epilogue, dU = G V, dV = G^T U; the
//                        score matrix lives in TMEM / registers only                       (tensor-bound part)
//   finalize_kernel      positive/diagonal corrections, l2-normalise backward, activity regulariser,
//                        sparse SGD scatter-add (atomics) or row gradients for lazy Adam    (HBM-bound part)
//   [adam kernels]       duplicate-summing + lazy Adam on the touched rows
// The PAIRS scheme ('original' / 'group_sample': row-wise dot on (1+k)B listed pairs) has its own two kernels.
//
// ref: models/model_framework.py:40-65,85-143; modules/interaction/interaction_dot.py:92-107;
//      utils/objectives.py:35-220; utils/utilities.py:122-135; utils/optimizer.py:108-147.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <vector>
#include "common.cuh"
#include "sm100.cuh"
#include "score_tc.cuh"

namespace nncf {

// =================================================================================================
// tf.unique (first occurrence order), one block per replica.   ref: models/model_framework.py:45-48
// =================================================================================================
constexpr int kUniqueThreads = 1024;
constexpr int kUniqueHashMax = 2048;      // up to this many ids the first occurrences come from a shared-memory hash table

__global__ void __launch_bounds__(kUniqueThreads)
unique_kernel(const int32_t* __restrict__ ids_all, int64_t ids_stride, int n, int32_t* __restrict__ uniq_all,
              int32_t* __restrict__ inv_all, int32_t* __restrict__ nuniq_all, int out_stride, int hash_slots) {
  extern __shared__ int32_t sm[];
  int32_t* s_ids = sm;            // [n]
  int32_t* s_first = sm + n;      // [n]
  int32_t* s_rank = sm + 2 * n;   // [n]
  __shared__ int s_warp_tot[32];
  __shared__ int s_carry;
  const int r = blockIdx.x;
  const int32_t* ids = ids_all + r * ids_stride;
  int32_t* uniq = uniq_all + (int64_t)r * out_stride;
  int32_t* inv = inv_all + (int64_t)r * out_stride;
  const int tid = threadIdx.x;
  pdl_launch_dependents();
  for (int i = tid; i < n; i += blockDim.x) s_ids[i] = ids[i];     // (link ids are inputs of the step: nothing in the stream writes them)
  if (tid == 0) s_carry = 0;
  pdl_wait();        // the previous step's kernels may still read uniq / inverse / n_unique (no-op under a plain launch)
  if (hash_slots > 0) {
    // first occurrence of every id through an open-addressing table of (id << 32 | position) entries: insert with
    // compare-and-swap into an empty slot or atomicMin into the id's slot (the smallest position wins: the result is
    // exactly the quadratic scan's, in ~2 probes per id instead of up to n comparisons - 23 us -> ~3 us at n = 512)
    unsigned long long* tab = reinterpret_cast<unsigned long long*>(sm + 3 * n + ((3 * n) & 1));       // 8-byte aligned
    constexpr unsigned long long kEmpty = ~0ull;
    const unsigned int mask = static_cast<unsigned int>(hash_slots - 1);
    for (int i = tid; i < hash_slots; i += blockDim.x) tab[i] = kEmpty;
    __syncthreads();
    for (int i = tid; i < n; i += blockDim.x) {
      const unsigned int key = static_cast<unsigned int>(s_ids[i]);
      const unsigned long long packed = (static_cast<unsigned long long>(key) << 32) | static_cast<unsigned int>(i);
      unsigned int slot = (key * 2654435761u) >> 7 & mask;
      while (true) {
        unsigned long long cur = *reinterpret_cast<volatile unsigned long long*>(&tab[slot]);
        if (cur == kEmpty) {
          cur = atomicCAS(&tab[slot], kEmpty, packed);
          if (cur == kEmpty) break;                                   // the slot is mine
        }
        if (static_cast<unsigned int>(cur >> 32) == key) { atomicMin(&tab[slot], packed); break; }
        slot = (slot + 1) & mask;
      }
    }
    __syncthreads();
    for (int i = tid; i < n; i += blockDim.x) {
      const unsigned int key = static_cast<unsigned int>(s_ids[i]);
      unsigned int slot = (key * 2654435761u) >> 7 & mask;
      while (static_cast<unsigned int>(tab[slot] >> 32) != key) slot = (slot + 1) & mask;
      s_first[i] = static_cast<int32_t>(tab[slot] & 0xFFFFFFFFull);
    }
  } else {
    __syncthreads();
    for (int i = tid; i < n; i += blockDim.x) {
      const int32_t me = s_ids[i];
      int f = i;
      for (int j = 0; j < i; ++j)
        if (s_ids[j] == me) { f = j; break; }
      s_first[i] = f;
    }
  }
  __syncthreads();
  // exclusive scan of is_first flags, chunk by chunk
  const int lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  for (int base = 0; base < n; base += blockDim.x) {
    const int i = base + tid;
    const int flag = (i < n && s_first[i] == i) ? 1 : 0;
    int x = flag;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) s_warp_tot[warp] = x;
    __syncthreads();
    if (warp == 0) {
      int t = (lane < nwarps) ? s_warp_tot[lane] : 0;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int y = __shfl_up_sync(0xffffffffu, t, o);
        if (lane >= o) t += y;
      }
      s_warp_tot[lane] = t;   // inclusive over warps
    }
    __syncthreads();
    const int warp_off = (warp == 0) ? 0 : s_warp_tot[warp - 1];
    const int excl = s_carry + warp_off + x - flag;
    if (i < n) s_rank[i] = excl;
    __syncthreads();
    if (tid == 0) s_carry += s_warp_tot[nwarps - 1];
    __syncthreads();
  }
  for (int i = tid; i < n; i += blockDim.x) {
    const int f = s_first[i];
    inv[i] = s_rank[f];
    if (f == i) uniq[s_rank[i]] = s_ids[i];
  }
  if (tid == 0) nuniq_all[r] = s_carry;
}

// =================================================================================================
// gather rows -> staging.  One warp per (padded) row.
// =================================================================================================
struct GatherArgs {
  ShardPtrs shards;
  const float* table;        // [n_rows_table, d] or NULL
  const float* dense_rows;   // [count, d] (used when table == NULL)
  const int32_t* ids;        // [R][ids_stride] (ignored for dense rows)
  int64_t ids_stride;
  const int32_t* count_dev;  // optional per-replica valid-row count (group_neg_shared n_unique), else NULL
  int count;                 // valid rows when count_dev == NULL
  int rows_pad;              // rows per replica in the staging buffers (multiple of 128)
  int d, dp;                 // true / padded dim
  int normalize;
  int write_img;
  int write_xf;              // write the fp32 staging rows (skipped when nothing downstream reads them)
  int zero_grad;             // zero dX (needed by the atomically-accumulating fp32 path only)
  float* Xf;                 // [R][rows_pad][dp]
  float* inv;                // [R][rows_pad]
  uint8_t* img;              // [R][rows_pad/128][dp/64][16 KiB]
  float* dX;                 // [R][rows_pad][dp]  zeroed here
  float* corr;               // [R][rows_pad]      zeroed here
  int d_emb;                 // columns [0, d_emb) are the embedding (l2-normalised if asked); [d_emb, d) are interaction-bias
                             // plumbing (a bias and a constant 1) that passes through unscaled.  0 = all d columns
  unsigned long long* tl;    // developer timeline (NNCF_TIMELINE): [0] first CTA start, [1] first CTA past the wait, [2] last end
};

__global__ void __launch_bounds__(256)
gather_rows_kernel(GatherArgs a0, GatherArgs a1) {
  pdl_launch_dependents();
  pdl_wait();
  const GatherArgs& a = blockIdx.z ? a1 : a0;     // z = 0 user side, z = 1 item side: one launch for both
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + warp;
  const int r = blockIdx.y;
  if (row >= a.rows_pad) return;
  const int count = a.count_dev ? a.count_dev[r] : a.count;
  const int nchunk = a.dp / 64;
  float x[8];   // columns 2*lane + 64*m (+1) for m < nchunk (dp <= 256)
#pragma unroll
  for (int m = 0; m < 8; ++m) x[m] = 0.0f;
  if (row < count) {
    const float* src;
    if (a.table) {
      const int64_t id = a.ids[r * a.ids_stride + row];
      src = a.table + id * a.d;
    } else {
      src = a.dense_rows + (int64_t)row * a.d;
    }
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      if (m < nchunk) {
        const int c = m * 64 + 2 * lane;
        if (c < a.d) x[2 * m] = __ldg(src + c);
        if (c + 1 < a.d) x[2 * m + 1] = __ldg(src + c + 1);
      }
    }
  }
  float inv = 1.0f;
  if (a.normalize) {
    const int de = a.d_emb > 0 ? a.d_emb : a.d;   // interaction-bias columns are not part of the embedding
    float ss = 0.0f;
#pragma unroll
    for (int m = 0; m < 8; ++m) { const int c = (m >> 1) * 64 + 2 * lane + (m & 1); ss += (c < de) ? x[m] * x[m] : 0.0f; }
    ss = warp_sum(ss);
    inv = rsqrtf(fmaxf(ss, 1e-12f));   // tf.nn.l2_normalize: x * rsqrt(max(sum(x^2), 1e-12))
    if (row >= count) inv = 1.0f;
#pragma unroll
    for (int m = 0; m < 8; ++m) { const int c = (m >> 1) * 64 + 2 * lane + (m & 1); if (c < de) x[m] *= inv; }
  }
  const int64_t rowoff = ((int64_t)r * a.rows_pad + row);
  float* xf = a.Xf + rowoff * a.dp;
  float* dx = a.dX + rowoff * a.dp;
#pragma unroll
  for (int m = 0; m < 4; ++m) {
    if (m < nchunk) {
      const int c = m * 64 + 2 * lane;
      if (a.write_xf) *reinterpret_cast<float2*>(xf + c) = make_float2(x[2 * m], x[2 * m + 1]);
      if (a.zero_grad) *reinterpret_cast<float2*>(dx + c) = make_float2(0.0f, 0.0f);
    }
  }
  if (lane == 0) {
    a.inv[rowoff] = inv;
    a.corr[rowoff] = 0.0f;
  }
  if (a.write_img) {
    uint8_t* blk = a.img + ((int64_t)r * (a.rows_pad / 128) + (row >> 7)) * nchunk * kSubBytes;
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      if (m < nchunk) {
        uint8_t* p = blk + m * kSubBytes + sw128_offset(row & 127, 2 * lane);
        *reinterpret_cast<uint32_t*>(p) = pack_bf16x2(x[2 * m], x[2 * m + 1]);
      }
    }
  }
}

// positive score per batch row:  spos[b] = <U_b, V_pos(b)>  (pos(b) = b for neg_shared, inverse[b] for group).
// In bf16 mode the operands are rounded to bf16 first so the value matches what the tensor cores see.
// Optionally (u_reg != 0) the same pass adds the activity regulariser's loss term u_reg * sum_d mean_b U_raw[b,d]^2 with
// U_raw = Uf / inv (ref: utils/utilities.py:129-135): one double atomic per CTA.  Launched with programmatic dependent
// launch behind the gather (it was two plain launches, pos_score + reg_loss: two exposed launch gaps and a second pass
// over the user rows).
__global__ void __launch_bounds__(256)
pos_score_kernel(const float* __restrict__ Uf, const float* __restrict__ Vf, const int32_t* __restrict__ inverse,
                 int rows_pad, int dp, int B, int round_bf16, float* __restrict__ spos,
                 const float* __restrict__ inv, float u_reg, int d_emb, double* loss) {
  __shared__ float part[8];
  pdl_launch_dependents();
  pdl_wait();                                   // the gather's staging rows must have landed
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x * 8 + warp;
  const int r = blockIdx.y;
  float reg = 0.0f;
  if (b < B) {
    const int64_t base = (int64_t)r * rows_pad;
    const int p = inverse ? inverse[base + b] : b;
    const float* u = Uf + (base + b) * dp;
    const float* v = Vf + (base + p) * dp;
    const int ncol = d_emb > 0 ? d_emb : dp;    // the regulariser sees the embedding columns only
    float acc = 0.0f, ss = 0.0f;
    for (int c = lane; c < dp; c += 32) {
      float a = u[c], w = v[c];
      if (c < ncol) ss = fmaf(a, a, ss);
      if (round_bf16) { a = __bfloat162float(__float2bfloat16(a)); w = __bfloat162float(__float2bfloat16(w)); }
      acc += a * w;
    }
    acc = warp_sum(acc);
    if (lane == 0) spos[base + b] = acc;
    if (u_reg != 0.0f) {
      ss = warp_sum(ss);
      const float iv = inv[base + b];
      reg = ss / (iv * iv);
    }
  }
  if (u_reg != 0.0f) {                           // (uniform: every thread of the CTA reaches the barrier)
    if (lane == 0) part[warp] = reg;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.0f;
#pragma unroll
      for (int w = 0; w < 8; ++w) t += part[w];
      if (t != 0.0f) atomicAdd(&loss[r], static_cast<double>(u_reg) * static_cast<double>(t) / static_cast<double>(B));
    }
  }
}

// =================================================================================================
// score + gradient tiles, CUDA-core fp32 path (precision = fp32).  CTA = 64 x 64 tile of S.
// =================================================================================================
struct ScoreArgs {
  const float* Uf; const float* Vf;      // [R][rows_pad][dp]
  const uint8_t* Uimg; const uint8_t* Vimg;
  float* dU; float* dV;                  // [R][rows_pad][dp]
  float* corrU; float* corrV;            // [R][rows_pad]
  const float* spos;                     // [R][rows_pad]
  const int32_t* inverse;                // [R][rows_pad] (group) or NULL
  const int32_t* ncols_dev;              // [R] (group) or NULL
  double* loss;                          // [R]
  int rows_pad, dp, B, scheme, loss_kind;
  float lambda, gamma;
};

constexpr int kSimtTile = 64;

__global__ void __launch_bounds__(256)
score_grad_simt_kernel(ScoreArgs a) {
  extern __shared__ float smf[];
  const int dp = a.dp, ld = dp + 1;
  float* Us = smf;                       // [64][ld]
  float* Vs = Us + kSimtTile * ld;       // [64][ld]
  float* Gs = Vs + kSimtTile * ld;       // [64][65]
  __shared__ float s_colA[kSimtTile];
  __shared__ float s_rowA[kSimtTile];
  __shared__ float s_loss[8];
  const int r = blockIdx.z;
  const int ncols = a.ncols_dev ? a.ncols_dev[r] : a.B;
  const int i0 = blockIdx.y * kSimtTile, j0 = blockIdx.x * kSimtTile;
  if (j0 >= ncols) return;
  const int tid = threadIdx.x;
  const int64_t base = (int64_t)r * a.rows_pad;
  for (int idx = tid; idx < kSimtTile * dp; idx += 256) {
    const int rr = idx / dp, c = idx - rr * dp;
    Us[rr * ld + c] = a.Uf[(base + i0 + rr) * dp + c];
    Vs[rr * ld + c] = a.Vf[(base + j0 + rr) * dp + c];
  }
  if (tid < kSimtTile) { s_colA[tid] = 0.0f; s_rowA[tid] = 0.0f; }
  __syncthreads();
  const int ty = tid >> 4, tx = tid & 15;
  float acc[4][4];
#pragma unroll
  for (int x = 0; x < 4; ++x)
#pragma unroll
    for (int y = 0; y < 4; ++y) acc[x][y] = 0.0f;
  for (int k = 0; k < dp; ++k) {
    float u[4], v[4];
#pragma unroll
    for (int x = 0; x < 4; ++x) { u[x] = Us[(ty * 4 + x) * ld + k]; v[x] = Vs[(tx * 4 + x) * ld + k]; }
#pragma unroll
    for (int x = 0; x < 4; ++x)
#pragma unroll
      for (int y = 0; y < 4; ++y) acc[x][y] = fmaf(u[x], v[y], acc[x][y]);
  }
  const EpiParams ep = make_epi(a.scheme, a.loss_kind, a.B, ncols, a.lambda, a.gamma);
  const bool pairwise = a.loss_kind >= NNCF_LOSS_LOG_LOSS;
  float lsum = 0.0f;
#pragma unroll
  for (int x = 0; x < 4; ++x) {
    const int il = ty * 4 + x, i = i0 + il;
    const int posc = (a.scheme == NNCF_SCHEME_GROUP_NEG_SHARED && i < a.B) ? a.inverse[base + i] : i;
#pragma unroll
    for (int y = 0; y < 4; ++y) {
      const int jl = tx * 4 + y, j = j0 + jl;
      float g = 0.0f, av = 0.0f, l = 0.0f;
      if (i < a.B && j < ncols) {
        float sp = 0.0f;
        if (pairwise) sp = (a.scheme == NNCF_SCHEME_NEG_SHARED) ? a.spos[base + j] : a.spos[base + i];
        epi_elem<false>(ep, acc[x][y], j == posc, sp, g, av, l);
        lsum += l;
        if (pairwise) {
          if (a.scheme == NNCF_SCHEME_NEG_SHARED) atomicAdd(&s_colA[jl], av);
          else atomicAdd(&s_rowA[il], av);
        }
      }
      Gs[il * 65 + jl] = g;
    }
  }
  lsum = warp_sum(lsum);
  if ((tid & 31) == 0) s_loss[tid >> 5] = lsum;
  __syncthreads();
  if (tid == 0) {
    float t = 0.0f;
    for (int w = 0; w < 8; ++w) t += s_loss[w];
    atomicAdd(&a.loss[r], static_cast<double>(t));
  }
  if (pairwise && tid < kSimtTile) {
    if (a.scheme == NNCF_SCHEME_NEG_SHARED) { if (j0 + tid < ncols) atomicAdd(&a.corrV[base + j0 + tid], s_colA[tid]); }
    else { if (i0 + tid < a.B) atomicAdd(&a.corrU[base + i0 + tid], s_rowA[tid]); }
  }
  // dU[i][k] += sum_j G[i][j] V[j][k];  dV[j][k] += sum_i G[i][j] U[i][k]
  for (int idx = tid; idx < kSimtTile * dp; idx += 256) {
    const int rr = idx / dp, c = idx - rr * dp;
    float su = 0.0f, sv = 0.0f;
    for (int t = 0; t < kSimtTile; ++t) {
      su = fmaf(Gs[rr * 65 + t], Vs[t * ld + c], su);
      sv = fmaf(Gs[t * 65 + rr], Us[t * ld + c], sv);
    }
    if (i0 + rr < a.B) atomicAdd(&a.dU[(base + i0 + rr) * dp + c], su);
    if (j0 + rr < ncols) atomicAdd(&a.dV[(base + j0 + rr) * dp + c], sv);
  }
}

// (the tcgen05 score + gradient kernel lives in score_tc.cuh)

// =================================================================================================
// finalize: corrections, normalise-backward, regulariser, optimizer.  One warp per batch row.
// =================================================================================================
struct FinalizeArgs {
  ShardPtrs shards;
  int side;                  // 0 = user rows, 1 = item rows
  int scheme, pairwise;
  const float* Xf;           // this side's staged rows        [R][rows_pad][dp]
  const float* Of;           // the other side's staged rows
  const float* inv;          // this side's 1/||x||
  float* dX;                 // this side's gradient accumulators (finalised in place)
  float* dO;                 // other side's accumulators (group pairwise: user side adds into item rows)
  const float* corr_self;    // group: corrU (side 0).  neg_shared: corrV (both sides use the column sums)
  const int32_t* inverse;    // group: [R][rows_pad]
  const int32_t* count_dev;  // per-replica valid rows (group item side) or NULL
  int count;
  int rows_pad, d, dp;
  int normalize;
  int write_back;            // store the finalised row gradient back to dX (lazy Adam reads it)
  int need_x;                // rows of Xf are read (normalise-backward, corrections or regulariser)
  float reg_scale;           // 2 * u_reg / rows  (user side) else 0
  int d_emb;                 // embedding columns (normalise-backward and regulariser act on [0, d_emb)); 0 = all
  int frozen0, frozen1;      // columns whose gradient is dropped (interaction-bias plumbing: the constant 1, an unused bias), -1 = none
  // optimizer
  int optimizer;
  float lr;
  float* table;
  const int32_t* ids; int64_t ids_stride;      // table row of each batch row
  float* grad_out;           // optional [count][d] copy of the final row gradients (replica 0 only)
  // loss hand-off, folded into the step's LAST finalize launch (saves a memset node and a one-block kernel per step):
  // block (0,0,0) publishes the R accumulated losses as fp32 and re-zeroes the accumulators for the next step
  double* loss_acc;          // [R] or NULL
  float* loss_out;           // [R] or NULL
  int n_replicas;
};

__device__ __forceinline__ void finalize_publish_loss(const FinalizeArgs& a) {
  if (a.loss_acc && (blockIdx.x | blockIdx.y | blockIdx.z) == 0) {
    for (int r = threadIdx.x; r < a.n_replicas; r += blockDim.x) {
      if (a.loss_out) a.loss_out[r] = static_cast<float>(a.loss_acc[r]);
      a.loss_acc[r] = 0.0;
    }
  }
}

__global__ void __launch_bounds__(256)
finalize_kernel(FinalizeArgs a0, FinalizeArgs a1) {
  pdl_launch_dependents();
  pdl_wait();
  const FinalizeArgs& a = blockIdx.z ? a1 : a0;
  finalize_publish_loss(a0);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + warp;
  const int r = blockIdx.y;
  const int count = a.count_dev ? a.count_dev[r] : a.count;
  if (row >= count) return;
  const int64_t base = (int64_t)r * a.rows_pad;
  const float* x = a.Xf + (base + row) * a.dp;
  float* dx = a.dX + (base + row) * a.dp;
  float g[8], xv[8];
#pragma unroll
  for (int m = 0; m < 8; ++m) {
    const int c = lane + 32 * m;
    g[m] = (c < a.dp) ? dx[c] : 0.0f;
    xv[m] = (a.need_x && c < a.dp) ? x[c] : 0.0f;
  }
  if (a.pairwise) {
    if (a.scheme == NNCF_SCHEME_NEG_SHARED) {
      // G[j,j] += colsum_j  =>  dU_j += colsum_j V_j ; dV_j += colsum_j U_j
      const float cs = a.corr_self[base + row];
      const float* o = a.Of + (base + row) * a.dp;
#pragma unroll
      for (int m = 0; m < 8; ++m) {
        const int c = lane + 32 * m;
        if (c < a.dp) g[m] = fmaf(cs, o[c], g[m]);
      }
    } else if (a.side == 0) {
      // G[i,pos_i] += rowsum_i  =>  dU_i += rowsum_i V_pos ; dV_pos += rowsum_i U_i (atomic, item side runs later)
      const float rs = a.corr_self[base + row];
      const int p = a.inverse[base + row];
      const float* o = a.Of + (base + p) * a.dp;
      float* dop = a.dO + (base + p) * a.dp;
#pragma unroll
      for (int m = 0; m < 8; ++m) {
        const int c = lane + 32 * m;
        if (c < a.dp) {
          g[m] = fmaf(rs, o[c], g[m]);
          atomicAdd(dop + c, rs * xv[m]);
        }
      }
    }
  }
  float invn = 1.0f;
  const int de = a.d_emb > 0 ? a.d_emb : a.dp;
  if (a.normalize) {
    invn = a.inv[base + row];
    float dot = 0.0f;
#pragma unroll
    for (int m = 0; m < 8; ++m) dot += (lane + 32 * m < de) ? g[m] * xv[m] : 0.0f;
    dot = warp_sum(dot);
#pragma unroll
    for (int m = 0; m < 8; ++m) if (lane + 32 * m < de) g[m] = (g[m] - xv[m] * dot) * invn;
  }
  if (a.reg_scale != 0.0f) {
    // regulariser acts on the UN-normalised row: x_raw = xhat / inv
    const float s = a.reg_scale / invn;
#pragma unroll
    for (int m = 0; m < 8; ++m) if (lane + 32 * m < de) g[m] = fmaf(s, xv[m], g[m]);
  }
#pragma unroll
  for (int m = 0; m < 8; ++m) {
    const int c = lane + 32 * m;
    if (c == a.frozen0 || c == a.frozen1) g[m] = 0.0f;
    if (c < a.dp) dx[c] = g[m];
  }
  if (a.grad_out && r == 0) {
#pragma unroll
    for (int m = 0; m < 8; ++m) {
      const int c = lane + 32 * m;
      if (c < a.d) a.grad_out[(int64_t)row * a.d + c] = g[m];
    }
  }
  if (a.optimizer == NNCF_OPT_SGD && a.table) {
    const int64_t id = a.ids[r * a.ids_stride + row];
    float* t = a.table + id * a.d;
#pragma unroll
    for (int m = 0; m < 8; ++m) {
      const int c = lane + 32 * m;
      if (c < a.d) atomicAdd(t + c, -a.lr * g[m]);
    }
  }
}

}  // namespace nncf
#include "row_kernels.cuh"
#include "sns_kernels.cuh"
namespace nncf {

// =================================================================================================
// lazy Adam on touched rows with duplicate summing.   ref: utils/optimizer.py:108-147
//   owner[id] (direct-address scratch, INT_MAX when idle) = smallest global batch position holding id.
// =================================================================================================
__global__ void adam_owner_kernel(const int32_t* __restrict__ ids, int64_t ids_stride, int count,
                                  const int32_t* __restrict__ count_dev, int rows_pad, int32_t* owner) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
  const int cnt = count_dev ? count_dev[r] : count;
  if (row >= cnt) return;
  atomicMin(&owner[ids[r * ids_stride + row]], r * rows_pad + row);
}
// non-owner rows add their gradient into the owner's row
__global__ void __launch_bounds__(256)
adam_combine_kernel(const int32_t* __restrict__ ids, int64_t ids_stride, int count, const int32_t* __restrict__ count_dev,
                    int rows_pad, int dp, const int32_t* __restrict__ owner, float* dX) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + warp, r = blockIdx.y;
  const int cnt = count_dev ? count_dev[r] : count;
  if (row >= cnt) return;
  const int me = r * rows_pad + row;
  const int own = owner[ids[r * ids_stride + row]];
  if (own == me) return;
  const float* src = dX + (int64_t)me * dp;
  float* dst = dX + (int64_t)own * dp;
  for (int c = lane; c < dp; c += 32) atomicAdd(dst + c, src[c]);
}
__global__ void __launch_bounds__(256)
adam_apply_kernel(const int32_t* __restrict__ ids, int64_t ids_stride, int count, const int32_t* __restrict__ count_dev,
                  int rows_pad, int d, int dp, int32_t* owner, const float* __restrict__ dX, float* table, float* m,
                  float* v, float lr_t, float beta1, float beta2, float eps, const float* lr_dev) {
  if (lr_dev) lr_t = *lr_dev;                     // device step clock (CUDA-graph replay): see adam_tick_kernel
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + warp, r = blockIdx.y;
  const int cnt = count_dev ? count_dev[r] : count;
  if (row >= cnt) return;
  const int me = r * rows_pad + row;
  const int64_t id = ids[r * ids_stride + row];
  if (owner[id] != me) return;
  const float* g = dX + (int64_t)me * dp;
  for (int c = lane; c < d; c += 32) {
    const int64_t o = id * d + c;
    const float gg = g[c];
    const float mm = beta1 * m[o] + (1.0f - beta1) * gg;
    const float vv = beta2 * v[o] + (1.0f - beta2) * gg * gg;
    m[o] = mm; v[o] = vv;
    table[o] -= lr_t * mm / (sqrtf(vv) + eps);
  }
}
__global__ void adam_reset_kernel(const int32_t* __restrict__ ids, int64_t ids_stride, int count,
                                  const int32_t* __restrict__ count_dev, int32_t* owner) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
  const int cnt = count_dev ? count_dev[r] : count;
  if (row >= cnt) return;
  owner[ids[r * ids_stride + row]] = 0x7fffffff;
}
// ---- vectorised, both tables per launch, programmatic dependent launch (the matmul schemes with 16-byte aligned rows).
// The four kernels above are 8 dependent launches per step for the two tables (+ gather, score, finalize = 11): at
// R = 37, B = 512 the step took 90 us, most of it launch gaps.  Here: 3 launches for both tables (owner, combine,
// apply), the reset folded into apply (the owner row re-arms owner[id] itself: any value != my position still means
// "not the owner" to a duplicate that looks later), 128-bit accesses, several rows in flight per warp.
struct AdamSide {
  const int32_t* ids; int64_t ids_stride; int count; const int32_t* count_dev;
  int32_t* owner; float* dX; float* table; float* m; float* v;
};
__global__ void adam_owner2_kernel(AdamSide s0, AdamSide s1, int rows_pad) {
  pdl_launch_dependents();
  pdl_wait();
  const AdamSide& s = blockIdx.z ? s1 : s0;
  if (!s.ids) return;
  const int row = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
  const int cnt = s.count_dev ? s.count_dev[r] : s.count;
  if (row >= cnt) return;
  atomicMin(&s.owner[s.ids[r * s.ids_stride + row]], r * rows_pad + row);
}
template <int NV>   // float4 chunks per lane: 1 for dp <= 128, 2 for dp = 256
__global__ void __launch_bounds__(256)
adam_combine_vec_kernel(AdamSide s0, AdamSide s1, int rows_pad, int dp) {
  pdl_launch_dependents();
  pdl_wait();
  const AdamSide& s = blockIdx.z ? s1 : s0;
  if (!s.ids) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + warp, r = blockIdx.y;
  const int cnt = s.count_dev ? s.count_dev[r] : s.count;
  if (row >= cnt) return;
  const int me = r * rows_pad + row;
  const int own = s.owner[s.ids[r * s.ids_stride + row]];
  if (own == me) return;
  const float* src = s.dX + (int64_t)me * dp;
  float* dst = s.dX + (int64_t)own * dp;
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const int c = 4 * (lane + 32 * v);
    if (c < dp) {
      const float4 g4 = *reinterpret_cast<const float4*>(src + c);
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + c), "f"(g4.x), "f"(g4.y), "f"(g4.z), "f"(g4.w) : "memory");
    }
  }
}
template <int NV>
__global__ void __launch_bounds__(256)
adam_apply_vec_kernel(AdamSide s0, AdamSide s1, int rows_pad, int d, int dp, float lr_t, float beta1, float beta2, float eps,
                      const float* lr_dev) {
  pdl_launch_dependents();
  pdl_wait();
  if (lr_dev) lr_t = *lr_dev;
  const AdamSide& s = blockIdx.z ? s1 : s0;
  if (!s.ids) return;
  constexpr int kRows = 4;                              // rows in flight per warp (id -> owner -> rows: dependent loads)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row0 = (blockIdx.x * 8 + warp) * kRows, r = blockIdx.y;
  const int cnt = s.count_dev ? s.count_dev[r] : s.count;
  if (row0 >= cnt) return;
  int64_t myid = -1;
  if (lane < kRows && row0 + lane < cnt) {
    const int64_t id = s.ids[r * s.ids_stride + row0 + lane];
    if (s.owner[id] == r * rows_pad + row0 + lane) myid = id;       // only the owner row applies (it holds the summed gradient)
  }
  float4 g[kRows][NV], mm[kRows][NV], vv[kRows][NV], pp[kRows][NV];
  int64_t ids[kRows];
#pragma unroll
  for (int k = 0; k < kRows; ++k) {
    ids[k] = __shfl_sync(0xffffffffu, myid, k);
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const int c = 4 * (lane + 32 * v);
      if (ids[k] >= 0 && c < d) {
        g[k][v] = *reinterpret_cast<const float4*>(s.dX + (int64_t)(r * rows_pad + row0 + k) * dp + c);
        const int64_t o = ids[k] * d + c;
        mm[k][v] = *reinterpret_cast<const float4*>(s.m + o);
        vv[k][v] = *reinterpret_cast<const float4*>(s.v + o);
        pp[k][v] = *reinterpret_cast<const float4*>(s.table + o);
      }
    }
  }
  const float c1 = 1.0f - beta1, c2 = 1.0f - beta2;
#pragma unroll
  for (int k = 0; k < kRows; ++k) {
    if (ids[k] < 0) continue;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const int c = 4 * (lane + 32 * v);
      if (c < d) {
        float4 G = g[k][v], M = mm[k][v], V = vv[k][v], P = pp[k][v];
        M.x = beta1 * M.x + c1 * G.x; M.y = beta1 * M.y + c1 * G.y; M.z = beta1 * M.z + c1 * G.z; M.w = beta1 * M.w + c1 * G.w;
        V.x = beta2 * V.x + c2 * G.x * G.x; V.y = beta2 * V.y + c2 * G.y * G.y; V.z = beta2 * V.z + c2 * G.z * G.z; V.w = beta2 * V.w + c2 * G.w * G.w;
        P.x -= lr_t * M.x / (sqrtf(V.x) + eps); P.y -= lr_t * M.y / (sqrtf(V.y) + eps);
        P.z -= lr_t * M.z / (sqrtf(V.z) + eps); P.w -= lr_t * M.w / (sqrtf(V.w) + eps);
        const int64_t o = ids[k] * d + c;
        *reinterpret_cast<float4*>(s.m + o) = M;
        *reinterpret_cast<float4*>(s.v + o) = V;
        *reinterpret_cast<float4*>(s.table + o) = P;
      }
    }
    if (lane == 0) s.owner[ids[k]] = 0x7fffffff;          // re-arm (was adam_reset_kernel)
  }
}
// ---- folded form (the plain matmul step: nothing to post-process): the score kernel's drain has ALREADY summed every
// position's gradient row into a per-table accumulation table keyed by id (bulk reductions at the L2, the same mechanism
// as the fused sparse SGD update; duplicates - also across replicas - sum there), so the owner / combine launches and the
// [R][rows][d] gradient staging disappear: gather, score, this kernel.  The first position of an id to swap the step's
// tag into claim[id] applies the _apply_sparse rule (ref: utils/optimizer.py:108-134) to the row and re-zeroes its
// accumulator; the other positions of the same id do nothing.
struct AdamAccSide {
  const int32_t* ids; int64_t ids_stride; int count; const int32_t* count_dev;
  int32_t* claim; float* acc; float* table; float* m; float* v;
};
template <int NV>
__global__ void __launch_bounds__(256)
adam_apply_accum_kernel(AdamAccSide s0, AdamAccSide s1, int tag, int d, float lr_t, float beta1, float beta2, float eps,
                        const float* lr_dev) {
  pdl_launch_dependents();
  const AdamAccSide& s = blockIdx.z ? s1 : s0;
  if (!s.ids) return;
  constexpr int kRows = 4;                              // rows in flight per warp (id -> claim -> rows: dependent accesses)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row0 = (blockIdx.x * 8 + warp) * kRows, r = blockIdx.y;
  // link ids are inputs of the step: fetched before waiting for the score kernel (the group scheme's compact ids are not)
  int64_t myid = -1;
  const bool early = s.count_dev == nullptr;
  if (early && lane < kRows && row0 + lane < s.count) myid = s.ids[r * s.ids_stride + row0 + lane];
  pdl_wait();
  if (lr_dev) lr_t = *lr_dev;
  const int cnt = s.count_dev ? s.count_dev[r] : s.count;
  if (row0 >= cnt) return;
  if (!early && lane < kRows && row0 + lane < cnt) myid = s.ids[r * s.ids_stride + row0 + lane];
  if (myid >= 0 && atomicExch(&s.claim[myid], tag) == tag) myid = -1;      // somebody else applies this id
  float4 g[kRows][NV], mm[kRows][NV], vv[kRows][NV], pp[kRows][NV];
  int64_t ids[kRows];
#pragma unroll
  for (int k = 0; k < kRows; ++k) {
    ids[k] = __shfl_sync(0xffffffffu, myid, k);
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const int c = 4 * (lane + 32 * v);
      if (ids[k] >= 0 && c < d) {
        const int64_t o = ids[k] * d + c;
        g[k][v] = *reinterpret_cast<const float4*>(s.acc + o);
        mm[k][v] = *reinterpret_cast<const float4*>(s.m + o);
        vv[k][v] = *reinterpret_cast<const float4*>(s.v + o);
        pp[k][v] = *reinterpret_cast<const float4*>(s.table + o);
      }
    }
  }
  const float c1 = 1.0f - beta1, c2 = 1.0f - beta2;
#pragma unroll
  for (int k = 0; k < kRows; ++k) {
    if (ids[k] < 0) continue;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const int c = 4 * (lane + 32 * v);
      if (c < d) {
        float4 G = g[k][v], M = mm[k][v], V = vv[k][v], P = pp[k][v];
        M.x = beta1 * M.x + c1 * G.x; M.y = beta1 * M.y + c1 * G.y; M.z = beta1 * M.z + c1 * G.z; M.w = beta1 * M.w + c1 * G.w;
        V.x = beta2 * V.x + c2 * G.x * G.x; V.y = beta2 * V.y + c2 * G.y * G.y; V.z = beta2 * V.z + c2 * G.z * G.z; V.w = beta2 * V.w + c2 * G.w * G.w;
        P.x -= lr_t * M.x / (sqrtf(V.x) + eps); P.y -= lr_t * M.y / (sqrtf(V.y) + eps);
        P.z -= lr_t * M.z / (sqrtf(V.z) + eps); P.w -= lr_t * M.w / (sqrtf(V.w) + eps);
        const int64_t o = ids[k] * d + c;
        *reinterpret_cast<float4*>(s.m + o) = M;
        *reinterpret_cast<float4*>(s.v + o) = V;
        *reinterpret_cast<float4*>(s.table + o) = P;
        *reinterpret_cast<float4*>(s.acc + o) = make_float4(0.f, 0.f, 0.f, 0.f);     // re-armed for the next step
      }
    }
  }
}
// Device step clock: when a training step is captured into a CUDA graph and replayed, a step count kept on the host (and the
// bias-corrected rate lr_t derived from it, passed by value) would be frozen into the graph.  With the clock enabled
// (nncf_trainer_set_device_clock) this one-thread kernel advances the count in device memory and leaves
// lr_t = lr sqrt(1 - beta2^t) / (1 - beta1^t) (ref: utils/optimizer.py:109-111) where the apply kernels read it.
__global__ void adam_tick_kernel(long long* step, float* lr_out, double lr, double beta1, double beta2) {
  const long long t = ++(*step);
  *lr_out = static_cast<float>(lr * sqrt(1.0 - pow(beta2, static_cast<double>(t))) / (1.0 - pow(beta1, static_cast<double>(t))));
}
__global__ void fill_i32_kernel(int32_t* p, int64_t n, int32_t v) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

// =================================================================================================
// PAIRS scheme: row-wise dot on (1+k)B listed pairs.   ref: interaction_dot.py:92-99, objectives.py:35-75
// =================================================================================================
struct PairsArgs {
  const float* EU; const float* EV;
  const int32_t* uid; const int32_t* cid; int64_t ids_stride;
  int B, k, d;
  int norm_u, norm_v, loss_kind;
  float lambda, gamma, u_reg;
  float* s;        // [R][n]   scores
  float* invu;     // [R][n]
  float* invv;     // [R][n]
  double* loss;    // [R]
  // second pass
  float* dUrows; float* dVrows;   // [R][n][d] final row gradients (optional unless Adam / grad_out)
  int optimizer; float lr;
  float* tableU; float* tableV;
  float* grad_out_u; float* grad_out_v;
  int d_emb;             // embedding columns (0 = all d); [d_emb, d) = interaction-bias plumbing, see GatherArgs
  int fu0, fu1, fv0, fv1;   // frozen columns of the user / item rows (-1 = none)
  const int32_t* resp;   // optional [R][n] response labels: pointwise losses take positive = (label == 1) from here
                         // (ref: utils/objectives.py:59-70 weights by y_true); NULL = positives are rows [0, B)
};

// One double atomic on loss[r] per CTA instead of one per row (same-address atomics serialise: ~1.5 ns each): the warps of a
// CTA add into a shared accumulator and the last one to arrive forwards the sum.  `active` = warps of this CTA that have a row.
struct CtaLoss { float sum; int arrived; };
__device__ __forceinline__ void cta_loss_init(CtaLoss* c) {
  if (threadIdx.x == 0) { c->sum = 0.0f; c->arrived = 0; }
  __syncthreads();
}
__device__ __forceinline__ void cta_loss_add(CtaLoss* c, float l, int active, double* dst) {   // lane 0 of every active warp, once
  if (l != 0.0f) atomicAdd(&c->sum, l);
  __threadfence_block();
  if (atomicAdd(&c->arrived, 1) == active - 1) {
    const float t = atomicAdd(&c->sum, 0.0f);
    if (t != 0.0f) atomicAdd(dst, static_cast<double>(t));
  }
}

__global__ void __launch_bounds__(256)
pairs_score_kernel(PairsArgs a) {
  __shared__ CtaLoss cl;
  cta_loss_init(&cl);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = (1 + a.k) * a.B;
  const int row = blockIdx.x * 8 + warp, r = blockIdx.y;
  if (row >= n) return;
  const float* u = a.EU + (int64_t)a.uid[r * a.ids_stride + row] * a.d;
  const float* v = a.EV + (int64_t)a.cid[r * a.ids_stride + row] * a.d;
  const int de = a.d_emb > 0 ? a.d_emb : a.d;
  float su = 0.0f, sv = 0.0f, dot = 0.0f, dotb = 0.0f;
  for (int c = lane; c < a.d; c += 32) {
    const float x = __ldg(u + c), y = __ldg(v + c);
    if (c < de) { su = fmaf(x, x, su); sv = fmaf(y, y, sv); dot = fmaf(x, y, dot); }
    else dotb = fmaf(x, y, dotb);                 // ubias * 1 + 1 * cbias (ref: interaction_dot.py:96-99)
  }
  su = warp_sum(su); sv = warp_sum(sv); dot = warp_sum(dot); dotb = warp_sum(dotb);
  const float iu = a.norm_u ? rsqrtf(fmaxf(su, 1e-12f)) : 1.0f;
  const float iv = a.norm_v ? rsqrtf(fmaxf(sv, 1e-12f)) : 1.0f;
  if (lane == 0) {
    const int64_t o = (int64_t)r * n + row;
    a.s[o] = dot * iu * iv + dotb;
    a.invu[o] = iu;
    a.invv[o] = iv;
    if (a.u_reg != 0.0f) cta_loss_add(&cl, a.u_reg * su / n, min(8, n - static_cast<int>(blockIdx.x) * 8), &a.loss[r]);
  }
}

__global__ void __launch_bounds__(256)
pairs_grad_kernel(PairsArgs a) {
  __shared__ CtaLoss cl;
  cta_loss_init(&cl);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int B = a.B, k = a.k, n = (1 + k) * B;
  const int row = blockIdx.x * 8 + warp, r = blockIdx.y;
  if (row >= n) return;
  const float* S = a.s + (int64_t)r * n;
  const float s = S[row];
  const bool pointwise = a.loss_kind <= NNCF_LOSS_MSE;
  const bool is_pos = (pointwise && a.resp) ? (a.resp[(int64_t)r * n + row] == 1) : (row < B);
  const float invB = 1.0f / B, w = a.lambda / k;
  float g = 0.0f, l = 0.0f;
  if (a.loss_kind == NNCF_LOSS_SKIP_GRAM) {
    if (is_pos) { l = softplus_f<false>(-s) * invB; g = (sigmoid_f<false>(s) - 1.0f) * invB; }
    else { l = w * softplus_f<false>(s) * invB; g = w * sigmoid_f<false>(s) * invB; }
  } else if (a.loss_kind == NNCF_LOSS_MSE) {
    if (is_pos) { l = (1.0f - s) * (1.0f - s) * invB; g = -2.0f * (1.0f - s) * invB; }
    else { l = w * s * s * invB; g = 2.0f * w * s * invB; }
  } else {
    const float inv_cnt = 1.0f / (static_cast<float>(k) * B);
    if (is_pos) {
      // g+ = sum over this positive's k negatives of a_n (each lane takes some negatives)
      float acc = 0.0f;
      for (int t = lane; t < k; t += 32) {
        const float dd = s - S[B + row * k + t];
        if (a.loss_kind == NNCF_LOSS_LOG_LOSS) acc += -a.gamma * sigmoid_f<false>(-a.gamma * dd) * inv_cnt;
        else acc += (a.gamma - dd > 0.0f) ? -inv_cnt : 0.0f;
      }
      g = warp_sum(acc);
    } else {
      const int p = (row - B) / k;
      const float dd = S[p] - s;
      if (a.loss_kind == NNCF_LOSS_LOG_LOSS) {
        l = softplus_f<false>(-a.gamma * dd) * inv_cnt;
        g = a.gamma * sigmoid_f<false>(-a.gamma * dd) * inv_cnt;
      } else {
        l = fmaxf(a.gamma - dd, 0.0f) * inv_cnt;
        g = (a.gamma - dd > 0.0f) ? inv_cnt : 0.0f;
      }
    }
  }
  if (lane == 0) cta_loss_add(&cl, l, min(8, n - static_cast<int>(blockIdx.x) * 8), &a.loss[r]);
  const int64_t uidx = a.uid[r * a.ids_stride + row], cidx = a.cid[r * a.ids_stride + row];
  const float* u = a.EU + uidx * a.d;
  const float* v = a.EV + cidx * a.d;
  const int64_t o = (int64_t)r * n + row;
  const float iu = a.invu[o], iv = a.invv[o];
  // dU_hat = g * V_hat, dV_hat = g * U_hat; then normalise-backward: dx = (dxhat - xhat (xhat . dxhat)) * inv
  // with xhat . dxhat = g * s_emb for both sides (s_emb = the embedding part of the score: bias columns pass through).
  const int de = a.d_emb > 0 ? a.d_emb : a.d;
  float se = s;
  if (de < a.d) {
    float dote = 0.0f;
    for (int c = lane; c < de; c += 32) dote = fmaf(__ldg(u + c), __ldg(v + c), dote);
    se = warp_sum(dote) * iu * iv;
  }
  const float reg = 2.0f * a.u_reg / n;
  for (int c = lane; c < a.d; c += 32) {
    const float x = __ldg(u + c), y = __ldg(v + c);
    const bool emb = c < de;
    const float xh = emb ? x * iu : x, yh = emb ? y * iv : y;
    float du = g * yh, dv = g * xh;
    if (a.norm_u && emb) du = (du - xh * (g * se)) * iu;
    if (a.norm_v && emb) dv = (dv - yh * (g * se)) * iv;
    if (emb) du = fmaf(reg, x, du);
    if (c == a.fu0 || c == a.fu1) du = 0.0f;
    if (c == a.fv0 || c == a.fv1) dv = 0.0f;
    if (a.dUrows) { a.dUrows[o * a.d + c] = du; a.dVrows[o * a.d + c] = dv; }
    if (a.grad_out_u && r == 0) a.grad_out_u[(int64_t)row * a.d + c] = du;
    if (a.grad_out_v && r == 0) a.grad_out_v[(int64_t)row * a.d + c] = dv;
  }
}
// SGD scatter for the PAIRS scheme runs as a second pass over the stored row gradients so that every
// gradient of the step is computed from the same snapshot of the tables.
__global__ void __launch_bounds__(256)
rows_sgd_kernel(const float* __restrict__ rows, const int32_t* __restrict__ ids, int64_t ids_stride, int n, int d,
                float lr, float* table) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + warp, r = blockIdx.y;
  if (row >= n) return;
  const float* g = rows + ((int64_t)r * n + row) * d;
  float* t = table + (int64_t)ids[r * ids_stride + row] * d;
  for (int c = lane; c < d; c += 32) atomicAdd(t + c, -lr * g[c]);
}

}  // namespace nncf
#include "pairs_kernels.cuh"
namespace nncf {

__global__ void loss_out_kernel(const double* __restrict__ loss, int R, float* out) {
  const int r = threadIdx.x;
  if (r < R) out[r] = static_cast<float>(loss[r]);
}
__global__ void __launch_bounds__(256)
reg_loss_kernel(const float* __restrict__ Uf, const float* __restrict__ inv, int rows_pad, int dp,
                int rows, float u_reg, double* loss, int d_emb) {
  // u_reg * sum_d mean_b U_raw[b,d]^2, U_raw = Uf / inv        ref: utils/utilities.py:129-135
  // Grid-stride over the rows, ONE double atomic per CTA: the first version added every row's term with its own atomic on
  // loss[r] - 16,384 same-address double atomics at C5 = 25 of the step's 530 us.
  __shared__ float part[8];
  pdl_launch_dependents();
  pdl_wait();                                                               // the gather's staging rows must have landed
  const int r = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ncol = d_emb > 0 ? d_emb : dp;                                  // embedding columns only
  float acc = 0.0f;
  for (int row = blockIdx.x * 8 + warp; row < rows; row += gridDim.x * 8) {
    const int64_t base = (int64_t)r * rows_pad + row;
    const float* x = Uf + base * dp;
    float ss = 0.0f;
    for (int c = lane; c < ncol; c += 32) ss = fmaf(x[c], x[c], ss);
    ss = warp_sum(ss);
    const float iv = inv[base];
    acc += ss / (iv * iv);
  }
  if (lane == 0) part[warp] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.0f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += part[w];
    atomicAdd(&loss[r], static_cast<double>(u_reg) * static_cast<double>(t) / static_cast<double>(rows));
  }
}

}  // namespace nncf

// =================================================================================================
// host side
// =================================================================================================
using namespace nncf;

__global__ void timeline_init_kernel(unsigned long long* tl, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int k = i & 15;
  if (i < n) tl[i] = (k == 0 || k == 1 || k == 3 || k == 4) ? ~0ull : 0ull;   // min slots / max slots
}

struct nncf_trainer {
  nncf_step_config cfg;
  int rows;        // user-side rows per batch: B or (1+k)B
  int rows_pad;    // padded to 128
  int dp, nsub;
  int64_t adam_t = 0;
  // workspace
  float *Uf = nullptr, *Vf = nullptr, *invU = nullptr, *invV = nullptr, *dU = nullptr, *dV = nullptr;
  float *corrU = nullptr, *corrV = nullptr, *spos = nullptr;
  uint8_t *Uimg = nullptr, *Vimg = nullptr;
  double* loss = nullptr;
  unsigned int* loss_count = nullptr;   // fused mode: per-replica arrival counters of the in-kernel loss hand-off
  int* gather_flags = nullptr;          // self-gather mode (score_tc.cuh): [R][2][rows_pad / 128] step sequence numbers
  unsigned long long* gather_count = nullptr;
  int gather_seq = 0;
  uint8_t* gx_buf = nullptr;            // G' exchange (score_tc.cuh): [R][2][nblk][nblk][32 KiB] bf16 gradient tiles, allocated at the first eligible step
  int* gx_flags = nullptr;              // [R][2][nblk][nblk] step sequence numbers
  int gx_seq = 0;
  int resident_ctas = 0;                // CTAs of the score kernel that fit on the device at once
  int32_t *uniq = nullptr, *inverse = nullptr, *nuniq = nullptr;
  int32_t *ownerU = nullptr, *ownerV = nullptr;
  int64_t ownerU_n = 0, ownerV_n = 0;
  // folded lazy Adam: per-table gradient accumulators [rows][d] (zero between steps) and claim words [rows]
  long long* step_dev = nullptr;        // device step clock (see adam_tick_kernel); lr_dev != nullptr = enabled
  float* lr_dev = nullptr;
  float *accU = nullptr, *accV = nullptr;
  int32_t *claimU = nullptr, *claimV = nullptr;
  int64_t accU_rows = 0, accV_rows = 0;
  float *ps = nullptr;   // PAIRS scores
  float *dVn = nullptr;  // sampled_neg_shared: [R][k][d] accumulated dL/d(vhat_neg)
  bool tc_attr_set = false;
  const int32_t *hint_next_uid = nullptr, *hint_next_cid = nullptr;   // ids of the step after this call's last one (L2 prefetch hint)
  bool loss_published = false;   // the step's finalize launch already wrote loss_out and re-zeroed the accumulators
  // row-sharded multi-GPU mode (nncf_trainer_set_shards)
  int n_shards = 1, rank = 0;
  float* ushards[16] = {nullptr};
  float* ishards[16] = {nullptr};
  void* flags[16] = {nullptr};
  unsigned int epoch = 0;
  // host-fed mode (nncf_train_steps_host): a ring of chunk buffers (ids and losses of up to host_chunk steps each), copy
  // streams, hand-off events
  static constexpr int kHostBufs = 3;
  static constexpr int kHostChunkMax = 16;
  int host_chunk = 0;                                                  // steps per chunk buffer (fixed at the first call)
  int32_t* h_ids[kHostBufs] = {nullptr, nullptr, nullptr};             // [2][host_chunk * R * rows] device staging (uids then cids)
  float* h_loss[kHostBufs] = {nullptr, nullptr, nullptr};              // [host_chunk][R] device
  cudaStream_t s_h2d = nullptr, s_d2h = nullptr;
  cudaEvent_t ev_ready[kHostBufs] = {}, ev_read[kHostBufs] = {};
  cudaEvent_t ev_step[kHostBufs] = {};                                 // chunk finished (its losses may leave, its id buffer is free)
  // developer timeline (env NNCF_TIMELINE=<file>): per step 16 stamp slots written by the kernels themselves,
  // dumped as text when the trainer is destroyed (tools/timeline.py reads it)
  unsigned long long* timeline = nullptr;
  int64_t tl_step = 0;
  static constexpr int kTlSteps = 4096;
  // optional per-phase device timing (CUDA events on the launching stream)
  bool profile = false;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  double phase_ms[3] = {0.0, 0.0, 0.0};
  int64_t phase_steps = 0;
};
#define NNCF_PROFILE_MARK(t, i, st) do { if ((t)->profile) NNCF_CUDA(cudaEventRecord((t)->ev[i], st)); } while (0)

template <typename T>
static int dev_alloc(T** p, size_t n) {
  NNCF_CUDA(cudaMalloc(reinterpret_cast<void**>(p), n * sizeof(T)));
  NNCF_CUDA(cudaMemset(*p, 0, n * sizeof(T)));
  return 0;
}

extern "C" int nncf_trainer_create(const nncf_step_config* cfg, nncf_trainer_t** out) {
  NNCF_CHECK_ARG(cfg && out, "nncf_trainer_create: null argument");
  NNCF_CHECK_ARG(cfg->scheme >= 0 && cfg->scheme <= 3, "unknown scheme");
  NNCF_CHECK_ARG(cfg->loss >= 0 && cfg->loss <= 3, "[ERROR!] loss not specified.");
  NNCF_CHECK_ARG(cfg->precision == NNCF_PREC_FP32 || cfg->precision == NNCF_PREC_BF16, "unknown precision");
  NNCF_CHECK_ARG(cfg->batch_size_p >= 2, "batch_size_p must be >= 2");
  NNCF_CHECK_ARG(cfg->dim >= 1 && cfg->dim <= 256, "dim must be in [1, 256]");
  NNCF_CHECK_ARG(cfg->replicas >= 1 && cfg->replicas <= 1024, "replicas must be in [1, 1024]");
  NNCF_CHECK_ARG(cfg->optimizer >= 0 && cfg->optimizer <= 2, "unknown optimizer");
  NNCF_CHECK_ARG(cfg->interaction_bias >= 0 && cfg->interaction_bias <= 3, "ERROR! Unknown interation bias");
  if (cfg->interaction_bias) {
    NNCF_CHECK_ARG(cfg->dim >= 3, "interaction bias: dim counts the two bias columns");
    NNCF_CHECK_ARG(cfg->scheme != NNCF_SCHEME_SAMPLED_NEG_SHARED, "interaction bias is not available for sampled_neg_shared");
  }
  if (cfg->scheme == NNCF_SCHEME_PAIRS || cfg->scheme == NNCF_SCHEME_SAMPLED_NEG_SHARED)
    NNCF_CHECK_ARG(cfg->num_negatives >= 1, "num_negatives must be >= 1");
  if (cfg->scheme == NNCF_SCHEME_SAMPLED_NEG_SHARED)
    NNCF_CHECK_ARG((int64_t)cfg->num_negatives * cfg->dim * 4 <= 160 * 1024, "sampled_neg_shared: num_negatives * dim too large for the shared-memory accumulators");
  if (cfg->scheme == NNCF_SCHEME_GROUP_NEG_SHARED)
    NNCF_CHECK_ARG(cfg->batch_size_p <= 12288, "group_neg_shared supports batch_size_p <= 12288");
  auto* t = new nncf_trainer();
  t->cfg = *cfg;
  const int R = cfg->replicas;
  t->rows = cfg->scheme == NNCF_SCHEME_PAIRS ? (1 + cfg->num_negatives) * cfg->batch_size_p
          : cfg->scheme == NNCF_SCHEME_SAMPLED_NEG_SHARED ? cfg->batch_size_p + cfg->num_negatives : cfg->batch_size_p;
  t->rows_pad = (t->rows + 127) / 128 * 128;
  t->dp = cfg->dim <= 64 ? 64 : (cfg->dim <= 128 ? 128 : 256);   // tensor-core kernels exist for dp = 64 / 128 / 256
  t->nsub = t->dp / 64;
  const size_t nrow = (size_t)R * t->rows_pad, nel = nrow * t->dp;
  int rc = 0;
  rc |= dev_alloc(&t->loss, (size_t)R);
  rc |= dev_alloc(&t->loss_count, (size_t)R);
  rc |= dev_alloc(&t->gather_flags, (size_t)R * 2 * (t->rows_pad / 128));
  rc |= dev_alloc(&t->gather_count, (size_t)1);
  {
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    t->resident_ctas = 2 * sms;         // ScoreTcCfg::kMinBlocks = 2 for dp <= 128 (the only shapes that self-gather)
  }
  if (cfg->scheme == NNCF_SCHEME_SAMPLED_NEG_SHARED) {
    rc |= dev_alloc(&t->dU, (size_t)R * t->rows * cfg->dim);
    rc |= dev_alloc(&t->dV, (size_t)R * t->rows * cfg->dim);
    rc |= dev_alloc(&t->dVn, (size_t)R * cfg->num_negatives * cfg->dim);
  } else if (cfg->scheme == NNCF_SCHEME_PAIRS) {
    rc |= dev_alloc(&t->ps, (size_t)R * t->rows);
    rc |= dev_alloc(&t->invU, (size_t)R * t->rows);
    rc |= dev_alloc(&t->invV, (size_t)R * t->rows);
    rc |= dev_alloc(&t->dU, (size_t)R * t->rows * cfg->dim);
    rc |= dev_alloc(&t->dV, (size_t)R * t->rows * cfg->dim);
  } else {
    rc |= dev_alloc(&t->Uf, nel); rc |= dev_alloc(&t->Vf, nel);
    rc |= dev_alloc(&t->dU, nel); rc |= dev_alloc(&t->dV, nel);
    rc |= dev_alloc(&t->invU, nrow); rc |= dev_alloc(&t->invV, nrow);
    rc |= dev_alloc(&t->corrU, nrow); rc |= dev_alloc(&t->corrV, nrow); rc |= dev_alloc(&t->spos, nrow);
    if (cfg->precision == NNCF_PREC_BF16) {
      rc |= dev_alloc(&t->Uimg, nel * 2); rc |= dev_alloc(&t->Vimg, nel * 2);
    }
    if (cfg->scheme == NNCF_SCHEME_GROUP_NEG_SHARED) {
      rc |= dev_alloc(&t->uniq, nrow); rc |= dev_alloc(&t->inverse, nrow); rc |= dev_alloc(&t->nuniq, (size_t)R);
    }
  }
  if (const char* tl = getenv("NNCF_TIMELINE")) {
    if (*tl) {
      rc |= dev_alloc(&t->timeline, (size_t)nncf_trainer::kTlSteps * 16);
      if (!rc) timeline_init_kernel<<<ceil_div(nncf_trainer::kTlSteps * 16, 256), 256>>>(t->timeline, nncf_trainer::kTlSteps * 16);
    }
  }
  if (rc) { nncf_trainer_destroy(t); return NNCF_ECUDA; }
  *out = t;
  return NNCF_OK;
}

extern "C" int nncf_trainer_destroy(nncf_trainer_t* t) {
  if (!t) return NNCF_OK;
  if (t->timeline) {
    const int64_t n = t->tl_step < nncf_trainer::kTlSteps ? t->tl_step : nncf_trainer::kTlSteps;
    std::vector<unsigned long long> h((size_t)nncf_trainer::kTlSteps * 16);
    cudaDeviceSynchronize();
    if (cudaMemcpy(h.data(), t->timeline, h.size() * 8, cudaMemcpyDeviceToHost) == cudaSuccess) {
      if (FILE* f = fopen(getenv("NNCF_TIMELINE") ? getenv("NNCF_TIMELINE") : "/dev/null", "a")) {
        fprintf(f, "# trainer R=%d B=%d d=%d steps=%lld (slot = step %% %d)\n", t->cfg.replicas, t->cfg.batch_size_p, t->cfg.dim, (long long)t->tl_step, nncf_trainer::kTlSteps);
        for (int64_t i = 0; i < n; ++i) {
          for (int k = 0; k < 16; ++k) fprintf(f, "%llu%c", h[i * 16 + k], k == 15 ? '\n' : ' ');
        }
        fclose(f);
      }
    }
    cudaFree(t->timeline);
  }
  void* ptrs[] = {t->Uf, t->Vf, t->invU, t->invV, t->dU, t->dV, t->corrU, t->corrV, t->spos, t->Uimg, t->Vimg,
                  t->loss, t->loss_count, t->uniq, t->inverse, t->nuniq, t->ownerU, t->ownerV, t->ps, t->dVn,
                  t->gather_flags, t->gather_count, t->gx_buf, t->gx_flags, t->accU, t->accV, t->claimU, t->claimV, t->step_dev, t->lr_dev};
  for (void* p : ptrs) if (p) cudaFree(p);
  for (int i = 0; i < 4; ++i) if (t->ev[i]) cudaEventDestroy(t->ev[i]);
  for (int i = 0; i < nncf_trainer::kHostBufs; ++i) {
    if (t->h_ids[i]) cudaFree(t->h_ids[i]);
    if (t->h_loss[i]) cudaFree(t->h_loss[i]);
    if (t->ev_ready[i]) cudaEventDestroy(t->ev_ready[i]);
    if (t->ev_read[i]) cudaEventDestroy(t->ev_read[i]);
  }
  for (cudaEvent_t e : t->ev_step) if (e) cudaEventDestroy(e);
  if (t->s_h2d) cudaStreamDestroy(t->s_h2d);
  if (t->s_d2h) cudaStreamDestroy(t->s_d2h);
  delete t;
  return NNCF_OK;
}

extern "C" int nncf_peer_barrier(void* const* flag_ptrs, int n_ranks, int rank, unsigned int epoch, void* stream);

extern "C" int nncf_trainer_set_shards(nncf_trainer_t* t, int n_shards, int rank, void* const* user_shards,
                                       void* const* item_shards, void* const* barrier_flags) {
  NNCF_CHECK_ARG(t, "nncf_trainer_set_shards: null trainer");
  NNCF_CHECK_ARG(n_shards >= 1 && n_shards <= 16 && rank >= 0 && rank < n_shards, "nncf_trainer_set_shards: bad rank / n_shards");
  if (n_shards > 1) {
    NNCF_CHECK_ARG(user_shards && item_shards && barrier_flags, "nncf_trainer_set_shards: null pointer arrays");
    NNCF_CHECK_ARG(t->cfg.scheme != NNCF_SCHEME_PAIRS, "sharded tables: matmul schemes only");
    NNCF_CHECK_ARG(t->cfg.optimizer != NNCF_OPT_LAZY_ADAM, "sharded tables: sparse SGD only (optimizer state is not sharded yet)");
    NNCF_CHECK_ARG(t->cfg.dim % 4 == 0, "sharded tables need dim % 4 == 0");
    for (int i = 0; i < n_shards; ++i) {
      t->ushards[i] = static_cast<float*>(user_shards[i]);
      t->ishards[i] = static_cast<float*>(item_shards[i]);
      t->flags[i] = barrier_flags[i];
    }
  }
  t->n_shards = n_shards;
  t->rank = rank;
  t->epoch = 0;
  return NNCF_OK;
}

extern "C" int nncf_trainer_set_device_clock(nncf_trainer_t* t, int enable) {
  NNCF_CHECK_ARG(t, "nncf_trainer_set_device_clock: null trainer");
  if (enable && !t->lr_dev) {
    NNCF_CUDA(cudaMalloc(reinterpret_cast<void**>(&t->step_dev), sizeof(long long)));
    NNCF_CUDA(cudaMalloc(reinterpret_cast<void**>(&t->lr_dev), sizeof(float)));
    const long long t0 = t->adam_t;                  // continue from the host count
    NNCF_CUDA(cudaMemcpy(t->step_dev, &t0, sizeof(long long), cudaMemcpyHostToDevice));
    NNCF_CUDA(cudaMemset(t->lr_dev, 0, sizeof(float)));
  } else if (!enable && t->lr_dev) {
    long long td = 0;
    NNCF_CUDA(cudaMemcpy(&td, t->step_dev, sizeof(long long), cudaMemcpyDeviceToHost));
    t->adam_t = td;
    cudaFree(t->step_dev); cudaFree(t->lr_dev);
    t->step_dev = nullptr; t->lr_dev = nullptr;
  }
  return NNCF_OK;
}

extern "C" int nncf_trainer_set_profile(nncf_trainer_t* t, int enable) {
  NNCF_CHECK_ARG(t, "nncf_trainer_set_profile: null trainer");
  if (enable && !t->ev[0])
    for (int i = 0; i < 4; ++i) NNCF_CUDA(cudaEventCreate(&t->ev[i]));
  t->profile = enable != 0;
  t->phase_ms[0] = t->phase_ms[1] = t->phase_ms[2] = 0.0;
  t->phase_steps = 0;
  return NNCF_OK;
}
extern "C" int nncf_trainer_get_profile(nncf_trainer_t* t, double* phase_ms_out, int64_t* steps_out) {
  NNCF_CHECK_ARG(t && phase_ms_out && steps_out, "nncf_trainer_get_profile: null argument");
  for (int p = 0; p < 3; ++p) phase_ms_out[p] = t->phase_ms[p];
  *steps_out = t->phase_steps;
  return NNCF_OK;
}

// dynamic shared memory of unique_kernel for n ids: three int arrays, plus (n <= kUniqueHashMax) the hash table of the next
// power of two >= 2 n 64-bit entries
static size_t unique_smem_bytes(int n, int* hash_slots) {
  size_t sm = (size_t)3 * n * sizeof(int32_t);
  *hash_slots = 0;
  if (n <= nncf::kUniqueHashMax) {
    int slots = 64;
    while (slots < 2 * n) slots *= 2;
    *hash_slots = slots;
    sm = ((size_t)3 * n + ((3 * n) & 1)) * sizeof(int32_t) + (size_t)slots * sizeof(unsigned long long);
  }
  return sm;
}

extern "C" int nncf_unique_first_occurrence(const int32_t* ids_dev, int n, int32_t* unique_ids_dev, int32_t* inverse_dev,
                                            int32_t* n_unique_dev, void* stream) {
  NNCF_CHECK_ARG(ids_dev && unique_ids_dev && inverse_dev && n_unique_dev, "nncf_unique_first_occurrence: null argument");
  NNCF_CHECK_ARG(n >= 1 && n <= 12288, "nncf_unique_first_occurrence: n must be in [1, 12288]");
  int slots = 0;
  const size_t sm = unique_smem_bytes(n, &slots);
  if (sm > 48 * 1024) NNCF_CUDA(cudaFuncSetAttribute(unique_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
  unique_kernel<<<1, kUniqueThreads, sm, (cudaStream_t)stream>>>(ids_dev, 0, n, unique_ids_dev, inverse_dev, n_unique_dev, 0, slots);
  NNCF_LAUNCH_OK();
  return NNCF_OK;
}

static int ensure_owner(int32_t** owner, int64_t* have, int64_t need, cudaStream_t st) {
  if (*have >= need) return 0;
  if (*owner) cudaFree(*owner);
  NNCF_CUDA(cudaMalloc(reinterpret_cast<void**>(owner), need * sizeof(int32_t)));
  fill_i32_kernel<<<ceil_div(need, 256), 256, 0, st>>>(*owner, need, 0x7fffffff);
  NNCF_LAUNCH_OK();
  *have = need;
  return 0;
}

static int ensure_accum(float** acc, int32_t** claim, int64_t* have, int64_t rows, int d, cudaStream_t st) {
  if (*have >= rows) return 0;
  if (*acc) cudaFree(*acc);
  if (*claim) cudaFree(*claim);
  *acc = nullptr; *claim = nullptr; *have = 0;
  NNCF_CUDA(cudaMalloc(reinterpret_cast<void**>(acc), (size_t)rows * d * sizeof(float)));
  NNCF_CUDA(cudaMalloc(reinterpret_cast<void**>(claim), (size_t)rows * sizeof(int32_t)));
  NNCF_CUDA(cudaMemsetAsync(*acc, 0, (size_t)rows * d * sizeof(float), st));
  NNCF_CUDA(cudaMemsetAsync(*claim, 0, (size_t)rows * sizeof(int32_t), st));
  *have = rows;
  return 0;
}

static int run_adam2(nncf_trainer* t, const AdamSide& su, const AdamSide& sv, int rows_stride, int d, int dp, float lr_t,
                     cudaStream_t st) {
  const int R = t->cfg.replicas;
  const int count = su.count > sv.count ? su.count : sv.count;
  NNCF_CUDA(launch_pdl(adam_owner2_kernel, dim3(ceil_div(count, 256), R, 2), dim3(256), 0, st, su, sv, rows_stride));
  if (dp <= 128) {
    NNCF_CUDA(launch_pdl(adam_combine_vec_kernel<1>, dim3(ceil_div(count, 8), R, 2), dim3(256), 0, st, su, sv, rows_stride, dp));
    NNCF_CUDA(launch_pdl(adam_apply_vec_kernel<1>, dim3(ceil_div(count, 32), R, 2), dim3(256), 0, st, su, sv, rows_stride, d, dp, lr_t,
                         t->cfg.beta1, t->cfg.beta2, t->cfg.epsilon, t->lr_dev));
  } else {
    NNCF_CUDA(launch_pdl(adam_combine_vec_kernel<2>, dim3(ceil_div(count, 8), R, 2), dim3(256), 0, st, su, sv, rows_stride, dp));
    NNCF_CUDA(launch_pdl(adam_apply_vec_kernel<2>, dim3(ceil_div(count, 32), R, 2), dim3(256), 0, st, su, sv, rows_stride, d, dp, lr_t,
                         t->cfg.beta1, t->cfg.beta2, t->cfg.epsilon, t->lr_dev));
  }
  return 0;
}

static int run_adam(nncf_trainer* t, const int32_t* ids, int64_t ids_stride, int count, const int32_t* count_dev,
                    int rows_stride, int d, int dp, int32_t* owner, float* dX, float* table, float* m, float* v,
                    float lr_t, cudaStream_t st) {
  const int R = t->cfg.replicas;
  dim3 g1(ceil_div(count, 256), R), g8(ceil_div(count, 8), R);
  adam_owner_kernel<<<g1, 256, 0, st>>>(ids, ids_stride, count, count_dev, rows_stride, owner);
  NNCF_LAUNCH_OK();
  adam_combine_kernel<<<g8, 256, 0, st>>>(ids, ids_stride, count, count_dev, rows_stride, dp, owner, dX);
  NNCF_LAUNCH_OK();
  adam_apply_kernel<<<g8, 256, 0, st>>>(ids, ids_stride, count, count_dev, rows_stride, d, dp, owner, dX, table, m, v,
                                        lr_t, t->cfg.beta1, t->cfg.beta2, t->cfg.epsilon, t->lr_dev);
  NNCF_LAUNCH_OK();
  adam_reset_kernel<<<g1, 256, 0, st>>>(ids, ids_stride, count, count_dev, owner);
  NNCF_LAUNCH_OK();
  return 0;
}

static int step_matmul(nncf_trainer* t, const nncf_tables* tb, const int32_t* uid, const int32_t* cid,
                       const nncf_step_io* io, bool last, float* loss_out_step, const int32_t* next_uid,
                       const int32_t* next_cid, cudaStream_t st) {
  const nncf_step_config& c = t->cfg;
  const int R = c.replicas, B = c.batch_size_p, d = c.dim, dp = t->dp, rp = t->rows_pad;
  const bool group = c.scheme == NNCF_SCHEME_GROUP_NEG_SHARED;
  const bool pairwise = c.loss >= NNCF_LOSS_LOG_LOSS;
  const bool bf16 = c.precision == NNCF_PREC_BF16;
  const bool dense_items = tb->item_table == nullptr;
  // interaction bias (ref: modules/interaction/interaction_dot.py:96-107): the tables carry two extra columns, users
  // (ubias, 1), items (1, cbias), so that <u, v> = <u_emb, v_emb> + ubias + cbias comes out of the same contraction;
  // they are excluded from l2-normalisation and the regulariser, and the gradient of a constant / unused column is dropped
  const int bias = c.interaction_bias;
  const int d_emb = bias ? d - 2 : 0;
  NNCF_PROFILE_MARK(t, 0, st);
  const bool want_row_grads = last && io && (io->grad_user_rows_dev || io->grad_item_rows_dev);
  // plain sparse SGD with nothing to post-process: the score kernel's drain applies the update itself, one bulk async
  // reduction (TMA engine) per row, and publishes the loss, so the step is two launches (NNCF_FUSE_SGD=0 disables it).
  // (A first version issued red.global.add.v4 from the epilogue warps: +9.7 us in the kernel for the 11 us it saved.)
  static const bool fuse_env = [] { const char* e = getenv("NNCF_FUSE_SGD"); return !e || atoi(e) != 0; }();
  // (the activity regulariser rides along: loss term and gradient are added by the user-side CTAs of the score kernel)
  const bool fuse_sgd = fuse_env && bf16 && c.optimizer == NNCF_OPT_SGD && !c.norm_u && !c.norm_v && !pairwise &&
                        (d % 4 == 0) && !dense_items && !want_row_grads && !bias;
  // lazy Adam with nothing to post-process: the score kernel publishes the loss itself (as in the fused SGD mode) and
  // writes the finished gradient blocks, the Adam kernels read them: no finalize launch
  const bool adam_plain = bf16 && c.optimizer == NNCF_OPT_LAZY_ADAM && !c.norm_u && !c.norm_v && !pairwise &&
                          (d % 4 == 0) && !dense_items && !want_row_grads && !bias && t->n_shards <= 1;
  const bool drain_reg = (fuse_sgd || adam_plain) && c.u_reg != 0.0f;    // regulariser gradient added by the score kernel's drain
  // folded lazy Adam (NNCF_ADAM_FOLD=0 keeps the owner / combine / apply form): the drain sums the gradient rows into
  // per-table accumulators keyed by id, one apply launch follows (see adam_apply_accum_kernel)
  const bool adam_fold = adam_plain && !t->lr_dev && [] { const char* e = getenv("NNCF_ADAM_FOLD"); return !e || atoi(e) != 0; }();
  if (adam_fold) {
    NNCF_CHECK_ARG(tb->user_m && tb->user_v && tb->item_m && tb->item_v, "lazy Adam needs m / v tables");
    if (ensure_accum(&t->accU, &t->claimU, &t->accU_rows, tb->n_users, d, st)) return NNCF_ECUDA;
    if (ensure_accum(&t->accV, &t->claimV, &t->accV_rows, tb->n_items, d, st)) return NNCF_ECUDA;
  }
  // the same accumulators behind the finalize pass (l2-normalised rows, pairwise losses, ...: whatever keeps the step off the
  // plain path): the finalize kernel's sparse-SGD form adds its finished gradient rows into them (red.global.add.v4 with
  // lr = -1) and ONE apply launch follows instead of owner / combine / apply (group_neg_shared + log-loss, R = 37: the
  // optimizer phase was 49 of the step's 110 us).  Not under a device step clock: the claim tag is the host's step count.
  const bool adam_fold_fin = !adam_plain && c.optimizer == NNCF_OPT_LAZY_ADAM && (d % 4 == 0) && !bias && t->n_shards <= 1 && !t->lr_dev &&
                             [] { const char* e = getenv("NNCF_ADAM_FOLD"); return !e || atoi(e) != 0; }();
  if (adam_fold_fin) {
    NNCF_CHECK_ARG(tb->user_m && tb->user_v && (dense_items || (tb->item_m && tb->item_v)), "lazy Adam needs m / v tables");
    if (ensure_accum(&t->accU, &t->claimU, &t->accU_rows, tb->n_users, d, st)) return NNCF_ECUDA;
    if (!dense_items) if (ensure_accum(&t->accV, &t->claimV, &t->accV_rows, tb->n_items, d, st)) return NNCF_ECUDA;
  }
  const bool drain_adds = fuse_sgd || adam_fold;     // the drain reduces into a table: no gradient staging blocks
  // t->loss is zero on entry: zeroed at creation and re-zeroed by whoever publishes the step's loss (the last finalize
  // launch, or in fused mode the last side-0 CTA of each replica inside the score kernel)
  t->loss_published = true;   // by the last finalize launch, or by the score kernel itself in fused mode
  const int32_t* item_ids = cid;
  int64_t item_stride = B;
  if (group) {
    if (dense_items) {
      // the framework tower already called nncf_unique_first_occurrence; it hands us inverse + n_unique
      NNCF_CHECK_ARG(io && io->inverse_dev && io->n_unique_dev, "dense group_neg_shared needs inverse_dev and n_unique_dev");
      NNCF_CUDA(cudaMemcpyAsync(t->inverse, io->inverse_dev, sizeof(int32_t) * B, cudaMemcpyDeviceToDevice, st));
      NNCF_CUDA(cudaMemcpyAsync(t->nuniq, io->n_unique_dev, sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
    } else {
      int slots = 0;
      const size_t sm = unique_smem_bytes(B, &slots);
      if (sm > 48 * 1024)
        NNCF_CUDA(cudaFuncSetAttribute(unique_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
      NNCF_CUDA(launch_pdl(unique_kernel, dim3(R), dim3(kUniqueThreads), sm, st, (const int32_t*)cid, (int64_t)B, B, t->uniq, t->inverse, t->nuniq, rp, slots));
      NNCF_LAUNCH_OK();
      item_ids = t->uniq;
      item_stride = rp;
    }
  }
  // gather
  GatherArgs gu{};
  gu.table = tb->user_table; gu.ids = uid; gu.ids_stride = B; gu.count = B; gu.rows_pad = rp; gu.d = d; gu.dp = dp;
  gu.normalize = c.norm_u; gu.write_img = bf16; gu.zero_grad = bf16 ? 0 : 1;
  gu.write_xf = (!bf16 || c.norm_u || c.norm_v || pairwise || (c.u_reg != 0.0f && !drain_reg)) ? 1 : 0; gu.Xf = t->Uf; gu.inv = t->invU; gu.img = t->Uimg; gu.dX = t->dU;
  gu.corr = t->corrU;
  unsigned long long* tl = nullptr;
  if (t->timeline) {
    tl = t->timeline + (t->tl_step % nncf_trainer::kTlSteps) * 16;
    if (t->tl_step >= nncf_trainer::kTlSteps) timeline_init_kernel<<<1, 16, 0, st>>>(tl, 16);
    ++t->tl_step;
  }
  gu.tl = tl;
  gu.d_emb = d_emb;
  GatherArgs gv = gu;
  gv.table = tb->item_table; gv.dense_rows = dense_items ? io->item_rows_dev : nullptr;
  gv.ids = item_ids; gv.ids_stride = item_stride; gv.count_dev = group ? t->nuniq : nullptr;
  gv.normalize = c.norm_v; gv.Xf = t->Vf; gv.inv = t->invV; gv.img = t->Vimg; gv.dX = t->dV; gv.corr = t->corrV;
  const bool sharded = t->n_shards > 1;
  if (sharded) {
    NNCF_CHECK_ARG(!dense_items, "sharded tables: embedding-table models only");
    gu.shards.n = gv.shards.n = t->n_shards;
    for (int i = 0; i < t->n_shards; ++i) { gu.shards.p[i] = t->ushards[i]; gv.shards.p[i] = t->ishards[i]; }
  }
  const bool vec = (d % 4 == 0) && !bias;   // 16-byte aligned rows: 128-bit loads / vector reductions (the scalar kernels carry the bias-column logic)
  // split sweep (score_tc.cuh): when a step is too small to fill the device (R = 1: 8 CTAs), `split` CTAs share an owner block
  // and add their partial gradient blocks.  NNCF_SPLIT overrides (developer switch).
  const int active_ctas = ceil_div(B, 128) * 2 * R;
  int split = 1;
  if (bf16 && vec && !dense_items && t->n_shards <= 1) {
    const int tn = 64, nt_min = ceil_div(B, tn);     // (group: fewer unique items than B leave some split CTAs without tiles; they exit at once)
    // Cost model in units of one tile: a CTA takes (tiles / split + c0) with c0 ~ 4 tiles of prologue + drain, the grid runs
    // in ceil(CTAs / resident) waves.  It reproduces the measurements: R = 1, B = 512 (8 CTAs, 8 tiles): 12.7 / 10.2 /
    // 9.0 / 10.5 us per step for split 1 / 2 / 4 / 8; and it removes wave quantisation where a step is a fractional number
    // of waves: C5 (256 CTAs of 256 tiles on 148 one-CTA SMs = 1.73 waves) and B = 4,096 x 5 replicas (320 CTAs on 296 slots).
    const int resident = dp <= 128 ? t->resident_ctas : t->resident_ctas / 2;     // dp = 256 kernels: one CTA per SM
    double best = 1e30;
    for (int sp = 1; sp <= 4 && sp <= nt_min; sp *= 2) {                         // (8-way measured slower than 4-way: 8x the partial updates)
      const double waves = ceil_div((int64_t)active_ctas * sp, resident);
      const double cost = waves * (static_cast<double>(nt_min) / sp + 4.0);
      if (cost < best * 0.97) { best = cost; split = sp; }                         // (needs a 3 % predicted gain to split further)
    }
    const int split_env = [] { const char* e = getenv("NNCF_SPLIT"); return e ? atoi(e) : 0; }();   // (read per step: the tests switch it)
    if (split_env >= 1 && split_env <= nt_min) split = split_env;
  }
  if (split > 1 && !drain_adds) gu.zero_grad = 1;    // partial blocks are summed with red.global.add
  // self-gather (opt-in, NNCF_SELF_GATHER=1): the score kernel's CTAs gather their own rows (no gather launch, no hand-off).
  // Needs the fused drain (nothing else reads the staging buffers), one local table, dp <= 128 and the whole grid resident
  // at once: CTAs wait for each other's image blocks, and no drain may start before the last gather has ended.
  // Measured on B200 and NOT the default: 27.9 vs 26.9 us per step at R = 37, 18.8 vs 15.1 us at R = 1 (B = 512, d = 128,
  // uniform ids).  The separate gather kernel has 1,184 CTAs x 8 warps x 4 rows in flight and its launch overlaps the score
  // kernel's prologue through programmatic dependent launch; 8 warps x 8 rows per CTA behind the prologue do not beat it.
  const bool self_env = [] { const char* e = getenv("NNCF_SELF_GATHER"); return e && atoi(e) != 0; }();   // (read per step: the tests switch it)
  if (split > 1 && !drain_adds) gv.zero_grad = 1;
  const bool self_gather = self_env && fuse_sgd && vec && !group && !sharded && dp <= 128 && !t->timeline && split == 1 &&
                           active_ctas <= t->resident_ctas;
  // gather / score / finalize are launched with programmatic dependent launch: each calls griddepcontrol.wait before it
  // reads what its predecessor wrote, so only launch latency and prologues overlap
  if (self_gather) { /* the score kernel gathers */ }
  else if (vec) {
    // rows per warp: 8 when the grid at 4 would not fit in one wave next to the resident CTAs of the following score kernel
    // (768 of an SM's 2,048 threads), else 4
    const bool fat = (size_t)2 * R * rp * 8 > (size_t)(t->resident_ctas / 2) * 1280;
    cudaError_t ge;
    if (dp <= 128) ge = fat ? launch_pdl(gather_rows_vec_kernel<1, 8>, dim3(rp / 64, R, 2), dim3(256), 0, st, gu, gv)
                            : launch_pdl(gather_rows_vec_kernel<1, 4>, dim3(rp / 32, R, 2), dim3(256), 0, st, gu, gv);
    else ge = fat ? launch_pdl(gather_rows_vec_kernel<2, 8>, dim3(rp / 64, R, 2), dim3(256), 0, st, gu, gv)
                  : launch_pdl(gather_rows_vec_kernel<2, 4>, dim3(rp / 32, R, 2), dim3(256), 0, st, gu, gv);
    NNCF_CUDA(ge);
  }
  else NNCF_CUDA(launch_pdl(gather_rows_kernel, dim3(rp / 8, R, 2), dim3(256), 0, st, gu, gv));
  if (!self_gather) count_launch();
  const bool reg_loss_here = c.u_reg != 0.0f && !drain_reg;   // (the fused step adds the regulariser's loss and gradient in the score kernel)
  if (pairwise) {
    NNCF_CUDA(launch_pdl(pos_score_kernel, dim3(ceil_div(B, 8), R), dim3(256), 0, st, (const float*)t->Uf, (const float*)t->Vf,
                         (const int32_t*)(group ? t->inverse : nullptr), rp, dp, B, bf16 ? 1 : 0, t->spos, (const float*)t->invU,
                         reg_loss_here ? c.u_reg : 0.0f, d_emb, t->loss));
  }
  if (reg_loss_here && !pairwise) {
    NNCF_CUDA(launch_pdl(reg_loss_kernel, dim3(std::min(ceil_div(B, 8), 4 * std::max(t->resident_ctas / 2, 1)), R), dim3(256), 0, st,
                         (const float*)t->Uf, (const float*)t->invU, rp, dp, B, c.u_reg, t->loss, d_emb));
  }
  NNCF_PROFILE_MARK(t, 1, st);
  // score + grad
  ScoreArgs sa{};
  sa.Uf = t->Uf; sa.Vf = t->Vf; sa.Uimg = t->Uimg; sa.Vimg = t->Vimg; sa.dU = t->dU; sa.dV = t->dV;
  sa.corrU = t->corrU; sa.corrV = t->corrV; sa.spos = t->spos; sa.inverse = group ? t->inverse : nullptr;
  sa.ncols_dev = group ? t->nuniq : nullptr; sa.loss = t->loss; sa.rows_pad = rp; sa.dp = dp; sa.B = B;
  sa.scheme = c.scheme; sa.loss_kind = c.loss; sa.lambda = c.neg_loss_weight; sa.gamma = c.loss_gamma;
  if (bf16) {
    ScoreTcArgs ta{};
    ta.Uimg = t->Uimg; ta.Vimg = t->Vimg; ta.dU = t->dU; ta.dV = t->dV; ta.corrU = t->corrU; ta.corrV = t->corrV;
    ta.spos = t->spos; ta.inverse = sa.inverse; ta.ncols_dev = sa.ncols_dev; ta.loss = t->loss; ta.rows_pad = rp;
    ta.B = B; ta.scheme = c.scheme; ta.loss_kind = c.loss; ta.lambda = c.neg_loss_weight; ta.gamma = c.loss_gamma;
    ta.loss_count = (fuse_sgd || adam_plain) ? t->loss_count : nullptr; ta.loss_out = (fuse_sgd || adam_plain) ? loss_out_step : nullptr;
    {   // drain form of the fused update: bulk reductions; NNCF_DRAIN_VEC=1 selects vector reductions (developer switch)
      const char* e = getenv("NNCF_DRAIN_VEC");
      ta.drain_vec = e ? atoi(e) : 0;     // measured at R = 1, split 4: 8.9 us per step with bulk reductions, 10.0 with vector reductions
    }
    {   // duplicate folding in the fused drain (one MATCH per warp), per side, where a step's rows crowd the table: more than
        // one position per 20 table rows (a stratified block at N >= 2; not the 1M-row tables of one GPU, where it costs 0.3 us
        // per step and folds next to nothing).  NNCF_DEDUP=0 / 1 switches both sides off / on.
      const char* e = getenv("NNCF_DEDUP");
      const int64_t pos = (int64_t)R * B;
      const int crowd_u = pos * 20 > tb->n_users ? 1 : 0, crowd_v = (!dense_items && pos * 20 > tb->n_items) ? 2 : 0;
      ta.dedup = !drain_adds ? 0 : (e ? (atoi(e) ? 3 : 0) : (crowd_u | crowd_v));
    }
    ta.split = split; ta.reg_scale = drain_reg ? 2.0f * c.u_reg / static_cast<float>(B) : 0.0f;
    ta.fuse_sgd = drain_adds ? 1 : 0; ta.d = d; ta.neg_lr = -c.learn_rate; ta.table_u = tb->user_table; ta.table_v = tb->item_table;
    ta.pf_table_u = tb->user_table; ta.pf_table_v = tb->item_table;
    if (adam_fold) { ta.table_u = t->accU; ta.table_v = t->accV; ta.neg_lr = 1.0f; }
    {   // folded lazy Adam: L2 prefetch of this step's m / v rows for the apply launch (NNCF_ADAM_PF=0: off, 3: accumulator rows
        // as well).  Measured at C3, R = 37: 45.9 us per step without, 43.7 with m / v, 44.4 with the accumulator rows too.
      static const int pf_env = [] { const char* e = getenv("NNCF_ADAM_PF"); return e ? atoi(e) : 1; }();
      if (adam_fold && !group && pf_env && (d % 4 == 0)) {
        ta.pf_now[0][0] = tb->user_m; ta.pf_now[0][1] = tb->user_v; ta.pf_now[1][0] = tb->item_m; ta.pf_now[1][1] = tb->item_v;
        if (pf_env & 2) { ta.pf_now[0][2] = t->accU; ta.pf_now[1][2] = t->accV; }
        ta.pf_now_count = R * B;
      }
    }
    ta.ids_u = uid; ta.ids_stride_u = B; ta.ids_v = item_ids; ta.ids_stride_v = item_stride;
    ta.shards_u = gu.shards; ta.shards_v = gv.shards;
    ta.tl = tl;
    // G' exchange (two-sided: every block's scores, sigmoids and loss computed once, by one side; the other side contracts
    // the exchanged bf16 tiles), opt-in with NNCF_GX=1: neg_shared with a pointwise loss, an even number of full 128-row
    // blocks, no split sweep.  Not under a device step clock (= the caller captures steps into a CUDA graph: the sequence
    // number is a kernel argument).  Measured slower than the one-sided kernel (C3: 25.2 vs 21.1 us per step,
    // tools/gpu_experiments_r02.sh gx), kept with its parity tests.
    {
      const bool gx_env = [] { const char* e = getenv("NNCF_GX"); return e && atoi(e) != 0; }();    // (read per step: the tests switch it)
      const bool gx_ok = gx_env && !group && !pairwise && dp <= 128 && (B % 256 == 0) && split == 1 && !self_gather &&
                         !dense_items && vec && !t->lr_dev && !NNCF_SCORE_TEAMS;
      if (gx_ok) {
        const size_t nb = (size_t)rp / 128;
        if (!t->gx_buf) {
          NNCF_CUDA(cudaMalloc(reinterpret_cast<void**>(&t->gx_buf), (size_t)R * 2 * nb * nb * 2 * kSubBytes));
          if (int rc3 = dev_alloc(&t->gx_flags, (size_t)R * 2 * nb * nb)) return rc3;
        }
        ta.gx = 1; ta.gx_seq = ++t->gx_seq; ta.gx_buf = t->gx_buf; ta.gx_flags = t->gx_flags;
      }
    }
    if (self_gather) {
      ta.self_gather = 1; ta.gather_seq = ++t->gather_seq; ta.gather_flags = t->gather_flags; ta.gather_count = t->gather_count;
      ta.gather_target = static_cast<unsigned long long>(t->gather_seq) * static_cast<unsigned long long>(active_ctas);
    }
    static const bool prefetch_env = [] { const char* e = getenv("NNCF_PREFETCH"); return !e || atoi(e) != 0; }();
    if (prefetch_env && next_uid && next_cid && !sharded && !dense_items && (d % 4 == 0)) {
      ta.next_ids_u = next_uid; ta.next_ids_v = next_cid; ta.next_count = R * B;
      ta.n_rows_u = tb->n_users; ta.n_rows_v = tb->n_items;
    }
    if (sharded && fuse_sgd) {
      // every rank has finished READING its peers' rows (gather) before any drain starts updating them
      if (int rc2 = nncf_peer_barrier(t->flags, t->n_shards, t->rank, ++t->epoch, st)) return rc2;
    }
    int rc = 0;
    switch (t->nsub) {
      case 1: rc = launch_score_tc_nsub1(ta, rp / 128, R, st); break;
      case 2: rc = launch_score_tc_nsub2(ta, rp / 128, R, st); break;
      default: rc = launch_score_tc_nsub4(ta, rp / 128, R, st); break;
    }
    if (rc) return rc;
  } else {
    const size_t sm = ((size_t)2 * kSimtTile * (dp + 1) + kSimtTile * 65) * sizeof(float);
    if (!t->tc_attr_set) {
      NNCF_CUDA(cudaFuncSetAttribute(score_grad_simt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
      t->tc_attr_set = true;
    }
    const int nt = ceil_div(B, kSimtTile);
    score_grad_simt_kernel<<<dim3(nt, nt, R), 256, sm, st>>>(sa);
    NNCF_LAUNCH_OK();
  }
  NNCF_PROFILE_MARK(t, 2, st);
  // finalize (user side first: for group pairwise it adds into the item accumulators)
  const bool sgd = c.optimizer == NNCF_OPT_SGD;
  FinalizeArgs fu{};
  fu.side = 0; fu.scheme = c.scheme; fu.pairwise = pairwise; fu.Xf = t->Uf; fu.Of = t->Vf; fu.inv = t->invU;
  fu.dX = t->dU; fu.dO = t->dV; fu.corr_self = group ? t->corrU : t->corrV; fu.inverse = group ? t->inverse : nullptr;
  fu.count = B; fu.rows_pad = rp; fu.d = d; fu.dp = dp; fu.normalize = c.norm_u; fu.need_x = (c.norm_u || pairwise || c.u_reg != 0.0f) ? 1 : 0;
  fu.write_back = (c.optimizer == NNCF_OPT_LAZY_ADAM) ? 1 : 0;
  fu.d_emb = d_emb; fu.frozen0 = bias ? d - 1 : -1; fu.frozen1 = (bias && !(bias & 1)) ? d - 2 : -1;   // users: (ubias, 1)
  fu.reg_scale = 2.0f * c.u_reg / B; fu.optimizer = c.optimizer; fu.lr = c.learn_rate; fu.table = tb->user_table;
  fu.ids = uid; fu.ids_stride = B; fu.grad_out = (last && io) ? io->grad_user_rows_dev : nullptr;
  FinalizeArgs fv = fu;
  fv.frozen0 = bias ? d - 2 : -1; fv.frozen1 = (bias && !(bias & 2)) ? d - 1 : -1;                      // items: (1, cbias)
  fv.side = 1; fv.Xf = t->Vf; fv.Of = t->Uf; fv.inv = t->invV; fv.dX = t->dV; fv.dO = nullptr; fv.corr_self = t->corrV;
  fv.count_dev = group ? t->nuniq : nullptr; fv.normalize = c.norm_v; fv.need_x = (c.norm_v || pairwise) ? 1 : 0; fv.reg_scale = 0.0f;
  fv.table = dense_items ? nullptr : tb->item_table; fv.ids = item_ids; fv.ids_stride = item_stride;
  fv.grad_out = (last && io) ? io->grad_item_rows_dev : nullptr;
  if (adam_fold_fin) {       // finished rows are ADDED into the per-table accumulators (the sparse-SGD form of the kernel with lr = -1)
    fu.optimizer = NNCF_OPT_SGD; fu.lr = -1.0f; fu.table = t->accU; fu.write_back = 0;
    fv.optimizer = NNCF_OPT_SGD; fv.lr = -1.0f; fv.table = dense_items ? nullptr : t->accV; fv.write_back = 0;
  }
  if (sharded && !fuse_sgd) {
    fu.shards = gu.shards; fv.shards = gv.shards;
    // every rank has finished READING its peers' rows (gather) before anybody starts updating them
    if (int rc = nncf_peer_barrier(t->flags, t->n_shards, t->rank, ++t->epoch, st)) return rc;
  }
  auto launch_finalize = [&](const FinalizeArgs& x, const FinalizeArgs& y, int nz) {
    cudaError_t e;
    if (vec && dp <= 128) e = launch_pdl(finalize_vec_kernel<1>, dim3(ceil_div(B, 32), R, nz), dim3(256), 0, st, x, y);
    else if (vec) e = launch_pdl(finalize_vec_kernel<2>, dim3(ceil_div(B, 32), R, nz), dim3(256), 0, st, x, y);
    else e = launch_pdl(finalize_kernel, dim3(ceil_div(B, 8), R, nz), dim3(256), 0, st, x, y);
    if (e != cudaSuccess) ::nncf::set_error(std::string("finalize launch: ") + cudaGetErrorString(e));
  };
  if (fuse_sgd) {
    // nothing left to do: the update was applied by the score kernel's drain
  } else if (adam_plain) {
    // nothing to post-process: the Adam kernels below read the gradient blocks as the score kernel left them
  } else if (group && pairwise) {
    // the user side adds the positive-column corrections into the item accumulators: strictly before the item side
    launch_finalize(fu, fu, 1);
    NNCF_LAUNCH_OK();
    fv.loss_acc = t->loss; fv.loss_out = loss_out_step; fv.n_replicas = R;
    launch_finalize(fv, fv, 1);
    NNCF_LAUNCH_OK();
  } else {
    fu.loss_acc = t->loss; fu.loss_out = loss_out_step; fu.n_replicas = R;
    launch_finalize(fu, fv, 2);
    NNCF_LAUNCH_OK();
  }
  if (sharded) {
    // all updates of this step have landed before the next step's gathers read the tables
    if (int rc = nncf_peer_barrier(t->flags, t->n_shards, t->rank, ++t->epoch, st)) return rc;
  }
  (void)sgd;
  if (c.optimizer == NNCF_OPT_LAZY_ADAM) {
    NNCF_CHECK_ARG(tb->user_m && tb->user_v, "lazy Adam needs user_m / user_v");
    t->adam_t += 1;
    if (t->lr_dev) {                                   // device step clock: the count and lr_t live in device memory
      adam_tick_kernel<<<1, 1, 0, st>>>(t->step_dev, t->lr_dev, (double)c.learn_rate, (double)c.beta1, (double)c.beta2);
      NNCF_LAUNCH_OK();
    }
    const double b1t = pow((double)c.beta1, (double)t->adam_t), b2t = pow((double)c.beta2, (double)t->adam_t);
    const float lr_t = (float)(c.learn_rate * sqrt(1.0 - b2t) / (1.0 - b1t));   // optimizer.py:109-111
    if (ensure_owner(&t->ownerU, &t->ownerU_n, tb->n_users, st)) return NNCF_ECUDA;
    if (!dense_items) {
      NNCF_CHECK_ARG(tb->item_m && tb->item_v, "lazy Adam needs item_m / item_v");
      if (ensure_owner(&t->ownerV, &t->ownerV_n, tb->n_items, st)) return NNCF_ECUDA;
    }
    if (adam_fold || adam_fold_fin) {
      AdamAccSide su{uid, B, B, nullptr, t->claimU, t->accU, tb->user_table, tb->user_m, tb->user_v};
      AdamAccSide sv{item_ids, item_stride, B, group ? t->nuniq : nullptr, t->claimV, t->accV, tb->item_table, tb->item_m, tb->item_v};
      if (dense_items) sv = AdamAccSide{};                               // (no item table: the tower owns the item side)
      const int tag = static_cast<int>(t->adam_t & 0x7fffffff);          // (claim words start at 0, adam_t at 1)
      if (dp <= 128)
        NNCF_CUDA(launch_pdl(adam_apply_accum_kernel<1>, dim3(ceil_div(B, 32), R, 2), dim3(256), 0, st, su, sv, tag, d, lr_t,
                             c.beta1, c.beta2, c.epsilon, t->lr_dev));
      else
        NNCF_CUDA(launch_pdl(adam_apply_accum_kernel<2>, dim3(ceil_div(B, 32), R, 2), dim3(256), 0, st, su, sv, tag, d, lr_t,
                             c.beta1, c.beta2, c.epsilon, t->lr_dev));
      count_launch();
    } else if (vec) {
      // both tables per launch, 3 launches (owner, combine, apply + re-arm), 128-bit accesses, programmatic dependent launch
      AdamSide su{uid, B, B, nullptr, t->ownerU, t->dU, tb->user_table, tb->user_m, tb->user_v};
      AdamSide sv{};
      if (!dense_items) sv = AdamSide{item_ids, item_stride, B, group ? t->nuniq : nullptr, t->ownerV, t->dV, tb->item_table, tb->item_m, tb->item_v};
      if (run_adam2(t, su, sv, rp, d, dp, lr_t, st)) return NNCF_ECUDA;
    } else {
      if (run_adam(t, uid, B, B, nullptr, rp, d, dp, t->ownerU, t->dU, tb->user_table, tb->user_m, tb->user_v, lr_t, st))
        return NNCF_ECUDA;
      if (!dense_items) {
        if (run_adam(t, item_ids, item_stride, B, group ? t->nuniq : nullptr, rp, d, dp, t->ownerV, t->dV, tb->item_table,
                     tb->item_m, tb->item_v, lr_t, st))
          return NNCF_ECUDA;
      }
    }
  }
  NNCF_PROFILE_MARK(t, 3, st);
  if (last && io && group && !dense_items) {
    if (io->unique_ids_dev) NNCF_CUDA(cudaMemcpyAsync(io->unique_ids_dev, t->uniq, sizeof(int32_t) * B, cudaMemcpyDeviceToDevice, st));
    if (io->inverse_dev) NNCF_CUDA(cudaMemcpyAsync(io->inverse_dev, t->inverse, sizeof(int32_t) * B, cudaMemcpyDeviceToDevice, st));
    if (io->n_unique_dev) NNCF_CUDA(cudaMemcpyAsync(io->n_unique_dev, t->nuniq, sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
  }
  return NNCF_OK;
}

static int apply_row_grads(nncf_trainer* t, const nncf_tables* tb, const int32_t* uid, const int32_t* cid, int n, cudaStream_t st);

static int step_pairs(nncf_trainer* t, const nncf_tables* tb, const int32_t* uid, const int32_t* cid,
                      const nncf_step_io* io, bool last, const int32_t* resp, cudaStream_t st) {
  const nncf_step_config& c = t->cfg;
  NNCF_CHECK_ARG(tb->item_table, "PAIRS scheme needs an item embedding table");
  const int R = c.replicas, n = t->rows, d = c.dim;
  NNCF_CUDA(cudaMemsetAsync(t->loss, 0, sizeof(double) * R, st));
  PairsArgs pa{};
  pa.EU = tb->user_table; pa.EV = tb->item_table; pa.uid = uid; pa.cid = cid; pa.ids_stride = n;
  pa.B = c.batch_size_p; pa.k = c.num_negatives; pa.d = d; pa.norm_u = c.norm_u; pa.norm_v = c.norm_v;
  pa.loss_kind = c.loss; pa.lambda = c.neg_loss_weight; pa.gamma = c.loss_gamma; pa.u_reg = c.u_reg;
  pa.s = t->ps; pa.invu = t->invU; pa.invv = t->invV; pa.loss = t->loss;
  pa.dUrows = t->dU; pa.dVrows = t->dV; pa.optimizer = c.optimizer; pa.lr = c.learn_rate;
  pa.tableU = tb->user_table; pa.tableV = tb->item_table;
  pa.grad_out_u = (last && io) ? io->grad_user_rows_dev : nullptr;
  pa.grad_out_v = (last && io) ? io->grad_item_rows_dev : nullptr;
  pa.resp = resp;
  { const int bias = c.interaction_bias;
    pa.d_emb = bias ? d - 2 : 0;
    pa.fu0 = bias ? d - 1 : -1; pa.fu1 = (bias && !(bias & 1)) ? d - 2 : -1;
    pa.fv0 = bias ? d - 2 : -1; pa.fv1 = (bias && !(bias & 2)) ? d - 1 : -1; }
  dim3 g8(ceil_div(n, 8), R), g32(ceil_div(n, 8 * kPairsPerWarp), R);
  const bool vec = (d % 4 == 0) && !c.interaction_bias && d <= 256;     // 16-byte aligned rows, no bias columns: pairs_kernels.cuh
  NNCF_PROFILE_MARK(t, 0, st);
  if (vec && d <= 128) pairs_score_vec_kernel<1><<<g32, 256, 0, st>>>(pa);
  else if (vec) pairs_score_vec_kernel<2><<<g32, 256, 0, st>>>(pa);
  else pairs_score_kernel<<<g8, 256, 0, st>>>(pa);
  NNCF_LAUNCH_OK();
  NNCF_PROFILE_MARK(t, 1, st);
  if (vec && d <= 128) pairs_grad_vec_kernel<1><<<g32, 256, 0, st>>>(pa);
  else if (vec) pairs_grad_vec_kernel<2><<<g32, 256, 0, st>>>(pa);
  else pairs_grad_kernel<<<g8, 256, 0, st>>>(pa);
  NNCF_LAUNCH_OK();
  NNCF_PROFILE_MARK(t, 2, st);
  if (int rc = apply_row_grads(t, tb, uid, cid, n, st)) return rc;
  NNCF_PROFILE_MARK(t, 3, st);
  return NNCF_OK;
}

// per-position row gradients t->dU / t->dV [R][n][d] -> sparse SGD (atomic scatter-add) or lazy Adam
static int apply_row_grads(nncf_trainer* t, const nncf_tables* tb, const int32_t* uid, const int32_t* cid, int n, cudaStream_t st) {
  const nncf_step_config& c = t->cfg;
  const int R = c.replicas, d = c.dim;
  dim3 g8(ceil_div(n, 8), R);
  if (c.optimizer == NNCF_OPT_SGD) {
    const dim3 g32(ceil_div(n, 8 * kPairsPerWarp), R);
    if (d % 4 == 0 && d <= 128) {
      rows_sgd_vec_kernel<1><<<g32, 256, 0, st>>>(t->dU, uid, n, n, d, c.learn_rate, tb->user_table);
      NNCF_LAUNCH_OK();
      rows_sgd_vec_kernel<1><<<g32, 256, 0, st>>>(t->dV, cid, n, n, d, c.learn_rate, tb->item_table);
      NNCF_LAUNCH_OK();
    } else if (d % 4 == 0 && d <= 256) {
      rows_sgd_vec_kernel<2><<<g32, 256, 0, st>>>(t->dU, uid, n, n, d, c.learn_rate, tb->user_table);
      NNCF_LAUNCH_OK();
      rows_sgd_vec_kernel<2><<<g32, 256, 0, st>>>(t->dV, cid, n, n, d, c.learn_rate, tb->item_table);
      NNCF_LAUNCH_OK();
    } else {
      rows_sgd_kernel<<<g8, 256, 0, st>>>(t->dU, uid, n, n, d, c.learn_rate, tb->user_table);
      NNCF_LAUNCH_OK();
      rows_sgd_kernel<<<g8, 256, 0, st>>>(t->dV, cid, n, n, d, c.learn_rate, tb->item_table);
      NNCF_LAUNCH_OK();
    }
  } else if (c.optimizer == NNCF_OPT_LAZY_ADAM) {
    NNCF_CHECK_ARG(tb->user_m && tb->user_v && tb->item_m && tb->item_v, "lazy Adam needs m / v tables");
    t->adam_t += 1;
    if (t->lr_dev) {                                   // device step clock: the count and lr_t live in device memory
      adam_tick_kernel<<<1, 1, 0, st>>>(t->step_dev, t->lr_dev, (double)c.learn_rate, (double)c.beta1, (double)c.beta2);
      NNCF_LAUNCH_OK();
    }
    const double b1t = pow((double)c.beta1, (double)t->adam_t), b2t = pow((double)c.beta2, (double)t->adam_t);
    const float lr_t = (float)(c.learn_rate * sqrt(1.0 - b2t) / (1.0 - b1t));
    if (ensure_owner(&t->ownerU, &t->ownerU_n, tb->n_users, st)) return NNCF_ECUDA;
    if (ensure_owner(&t->ownerV, &t->ownerV_n, tb->n_items, st)) return NNCF_ECUDA;
    if (run_adam(t, uid, n, n, nullptr, n, d, d, t->ownerU, t->dU, tb->user_table, tb->user_m, tb->user_v, lr_t, st))
      return NNCF_ECUDA;
    if (run_adam(t, cid, n, n, nullptr, n, d, d, t->ownerV, t->dV, tb->item_table, tb->item_m, tb->item_v, lr_t, st))
      return NNCF_ECUDA;
  }
  return NNCF_OK;
}

// sampled_neg_shared: B positives + k shared sampled negatives per batch (ref: models/train_sampled_neg_shared.py)
static int step_sns(nncf_trainer* t, const nncf_tables* tb, const int32_t* uid, const int32_t* cid,
                    const nncf_step_io* io, bool last, cudaStream_t st) {
  const nncf_step_config& c = t->cfg;
  NNCF_CHECK_ARG(tb->item_table, "sampled_neg_shared needs an item embedding table");
  const int R = c.replicas, n = t->rows, d = c.dim, k = c.num_negatives, B = c.batch_size_p;
  NNCF_CUDA(cudaMemsetAsync(t->loss, 0, sizeof(double) * R, st));
  NNCF_CUDA(cudaMemsetAsync(t->dVn, 0, sizeof(float) * (size_t)R * k * d, st));
  SnsArgs a{};
  a.EU = tb->user_table; a.EV = tb->item_table; a.uid = uid; a.cid = cid; a.B = B; a.k = k; a.d = d;
  a.norm_u = c.norm_u; a.norm_v = c.norm_v; a.loss_kind = c.loss; a.lambda = c.neg_loss_weight; a.gamma = c.loss_gamma;
  a.u_reg = c.u_reg; a.loss = t->loss; a.dUrows = t->dU; a.dVrows = t->dV; a.dVn_hat = t->dVn;
  a.grad_out_u = (last && io) ? io->grad_user_rows_dev : nullptr;
  a.grad_out_v = (last && io) ? io->grad_item_rows_dev : nullptr;
  const size_t sm = (size_t)k * d * sizeof(float);
  if (sm > 48 * 1024 && !t->tc_attr_set) {
    NNCF_CUDA(cudaFuncSetAttribute(sns_main_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    t->tc_attr_set = true;
  }
  NNCF_PROFILE_MARK(t, 0, st);
  NNCF_PROFILE_MARK(t, 1, st);
  sns_main_kernel<<<dim3(ceil_div(B, 8), R), 256, sm, st>>>(a);
  NNCF_LAUNCH_OK();
  sns_back_kernel<<<dim3(ceil_div(k, 8), R), 256, 0, st>>>(a);
  NNCF_LAUNCH_OK();
  NNCF_PROFILE_MARK(t, 2, st);
  if (int rc = apply_row_grads(t, tb, uid, cid, n, st)) return rc;
  NNCF_PROFILE_MARK(t, 3, st);
  return NNCF_OK;
}

extern "C" int nncf_train_steps(nncf_trainer_t* t, const nncf_tables* tables, const int32_t* user_ids_dev,
                                const int32_t* item_ids_dev, int64_t n_steps, const nncf_step_io* io, void* stream) {
  NNCF_CHECK_ARG(t && tables && user_ids_dev && item_ids_dev, "nncf_train_steps: null argument");
  NNCF_CHECK_ARG(tables->user_table, "nncf_train_steps: user_table is required");
  NNCF_CHECK_ARG(n_steps >= 0, "nncf_train_steps: n_steps < 0");
  const bool dense_items = tables->item_table == nullptr;
  if (dense_items) {
    NNCF_CHECK_ARG(io && io->item_rows_dev, "dense item side needs io->item_rows_dev");
    NNCF_CHECK_ARG(t->cfg.replicas == 1 && n_steps <= 1, "dense item side supports one batch per call");
    NNCF_CHECK_ARG(t->cfg.scheme < NNCF_SCHEME_PAIRS, "dense item side is available for the matmul schemes only");
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int R = t->cfg.replicas;
  const int64_t per_step = (int64_t)R * t->rows;
  for (int64_t s = 0; s < n_steps; ++s) {
    const bool last = (s + 1 == n_steps);
    int rc;
    float* loss_out_step = (io && io->loss_out_dev) ? io->loss_out_dev + s * R : nullptr;
    t->loss_published = false;
    if (t->cfg.scheme == NNCF_SCHEME_PAIRS)
      rc = step_pairs(t, tables, user_ids_dev + s * per_step, item_ids_dev + s * per_step, io, last,
                      (io && io->response_dev) ? io->response_dev + s * per_step : nullptr, st);
    else if (t->cfg.scheme == NNCF_SCHEME_SAMPLED_NEG_SHARED)
      rc = step_sns(t, tables, user_ids_dev + s * per_step, item_ids_dev + s * per_step, io, last, st);
    else
      rc = step_matmul(t, tables, user_ids_dev + s * per_step, item_ids_dev + s * per_step, io, last, loss_out_step,
                       last ? (t->hint_next_uid ? t->hint_next_uid : (io ? io->next_user_ids_dev : nullptr)) : user_ids_dev + (s + 1) * per_step,
                       last ? (t->hint_next_cid ? t->hint_next_cid : (io ? io->next_item_ids_dev : nullptr)) : item_ids_dev + (s + 1) * per_step, st);
    if (rc) return rc;
    if (t->profile) {
      NNCF_CUDA(cudaEventSynchronize(t->ev[3]));
      for (int p = 0; p < 3; ++p) {
        float ms = 0.0f;
        NNCF_CUDA(cudaEventElapsedTime(&ms, t->ev[p], t->ev[p + 1]));
        t->phase_ms[p] += ms;
      }
      t->phase_steps += 1;
    }
    if (loss_out_step && !t->loss_published) {
      loss_out_kernel<<<1, R < 32 ? 32 : ((R + 31) / 32 * 32), 0, st>>>(t->loss, R, io->loss_out_dev + s * R);
      NNCF_LAUNCH_OK();
    }
  }
  return NNCF_OK;
}

// Host-fed training loop: the link ids of every batch live in HOST memory, as the reference's `train` array does (it
// slices a NumPy array per batch and feeds it through feed_dict, ref: models/train_neg_shared.py:46-50), and every
// batch's loss goes back to the host (what train_on_batch returns).  The ids travel on a copy stream into a ring of
// device chunk buffers while earlier steps compute; the losses of a chunk's steps return on a second copy stream as soon as
// the chunk has finished.  Nothing is skipped: every step's ids are copied H2D, every step's losses D2H; the host is only
// synchronised once, at the end.
// The loop is bound by the HOST's CUDA calls (3-5 us each on the boxes this was measured on; tools/host_fed_trace.py shows
// the device idling at every point where the host has more than a launch pair to issue between two steps).  The first
// version made 12 calls per step - and its two stream waits + event between consecutive steps also broke the
// programmatic-dependent-launch overlap of step s + 1's gather with step s's drain: 26 us per C3 step against 21.7
// device-fed.  Now ids AND losses travel in chunks of 1, 4, 16, 16, ... steps (the short first chunk lets the call start
// after one step's worth of ids): two H2D copies, one D2H copy, three events and two waits per CHUNK, and per step nothing
// but the step's own two launches.  The ids of chunk k + 2 are enqueued right AFTER chunk k's launches, never in front of
// them.  Measured and dropped: the kernel storing its losses straight into pinned host memory (no copy at all) - the PCIe
// write sits at the end of the step's last CTA and the next step waits for it: 26-28 us per step.
extern "C" int nncf_train_steps_host(nncf_trainer_t* t, const nncf_tables* tables, const int32_t* user_ids_host,
                                     const int32_t* item_ids_host, int64_t n_steps, float* loss_out_host, void* stream) {
  NNCF_CHECK_ARG(t && tables && user_ids_host && item_ids_host, "nncf_train_steps_host: null argument");
  NNCF_CHECK_ARG(tables->user_table && tables->item_table, "nncf_train_steps_host: embedding-table models only");
  NNCF_CHECK_ARG(n_steps >= 0, "nncf_train_steps_host: n_steps < 0");
  constexpr int NB = nncf_trainer::kHostBufs;
  cudaStream_t st = (cudaStream_t)stream;
  const int R = t->cfg.replicas;
  const int64_t per_step = (int64_t)R * t->rows;
  if (!t->s_h2d) {
    // longest chunk: a power of four, up to 16 steps or ~2 MB of ids per side (NNCF_HOST_CHUNK overrides)
    int64_t want = std::max<int64_t>(1, (1 << 19) / per_step);
    if (const char* e = getenv("NNCF_HOST_CHUNK")) want = std::max(1, atoi(e));
    int ch = 1;
    while (ch * 4 <= want && ch * 4 <= nncf_trainer::kHostChunkMax) ch *= 4;
    t->host_chunk = ch;
    NNCF_CUDA(cudaStreamCreateWithFlags(&t->s_h2d, cudaStreamNonBlocking));
    NNCF_CUDA(cudaStreamCreateWithFlags(&t->s_d2h, cudaStreamNonBlocking));
    for (int i = 0; i < NB; ++i) {
      NNCF_CUDA(cudaMalloc(reinterpret_cast<void**>(&t->h_ids[i]), 2 * ch * per_step * sizeof(int32_t)));
      NNCF_CUDA(cudaMalloc(reinterpret_cast<void**>(&t->h_loss[i]), (size_t)ch * R * sizeof(float)));
      NNCF_CUDA(cudaEventCreateWithFlags(&t->ev_ready[i], cudaEventDisableTiming));
      NNCF_CUDA(cudaEventCreateWithFlags(&t->ev_read[i], cudaEventDisableTiming));
      NNCF_CUDA(cudaEventCreateWithFlags(&t->ev_step[i], cudaEventDisableTiming));
    }
  }
  const int CH = t->host_chunk;
  int ramp = 0;                                                        // chunks 0 .. ramp - 1 have 1, 4, 16, ... steps, the rest CH
  while (((int64_t)1 << (2 * ramp)) < CH) ++ramp;
  auto pow4 = [](int64_t k) -> int64_t { return (int64_t)1 << (2 * k); };
  const int64_t ramp_steps = (pow4(ramp) - 1) / 3;                     // 1 + 4 + ... + 4^(ramp - 1)
  const int64_t cid_off = (int64_t)CH * per_step;                      // the cids of a chunk buffer start here
  auto first = [&](int64_t k) -> int64_t { return k <= ramp ? (pow4(k) - 1) / 3 : ramp_steps + (k - ramp) * CH; };
  auto len = [&](int64_t k) -> int64_t { return std::min<int64_t>(k < ramp ? pow4(k) : CH, n_steps - first(k)); };
  const int64_t n_chunks = n_steps <= ramp_steps ? [&] { int64_t k = 0; while (first(k) < n_steps) ++k; return k; }()
                                                 : ramp + (n_steps - ramp_steps + CH - 1) / CH;
  auto enqueue_ids = [&](int64_t k) -> int {
    const int b = static_cast<int>(k % NB);
    const int64_t s0 = first(k), n = len(k);
    if (k >= NB) NNCF_CUDA(cudaStreamWaitEvent(t->s_h2d, t->ev_step[b], 0));   // chunk k - NB has finished with buffer b
    NNCF_CUDA(cudaMemcpyAsync(t->h_ids[b], user_ids_host + s0 * per_step, n * per_step * sizeof(int32_t), cudaMemcpyHostToDevice, t->s_h2d));
    NNCF_CUDA(cudaMemcpyAsync(t->h_ids[b] + cid_off, item_ids_host + s0 * per_step, n * per_step * sizeof(int32_t), cudaMemcpyHostToDevice, t->s_h2d));
    NNCF_CUDA(cudaEventRecord(t->ev_ready[b], t->s_h2d));
    return NNCF_OK;
  };
  // developer trace (NNCF_HOST_TRACE=1): host-side wall clock of the call's phases on stderr
  static const bool trace = [] { const char* e = getenv("NNCF_HOST_TRACE"); return e && atoi(e) != 0; }();
  auto now_us = [] { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e6 + ts.tv_nsec * 1e-3; };
  const double t_in = trace ? now_us() : 0.0;
  nncf_step_io io{};
  if (n_chunks > 0) if (int rc = enqueue_ids(0)) return rc;
  if (n_chunks > 1) if (int rc = enqueue_ids(1)) return rc;          // (behind chunk 0's one step it would land after that step has ended)
  for (int64_t k = 0; k < n_chunks; ++k) {
    const int b = static_cast<int>(k % NB), bn = static_cast<int>((k + 1) % NB);
    NNCF_CUDA(cudaStreamWaitEvent(st, t->ev_ready[b], 0));
    if (k >= NB && loss_out_host) NNCF_CUDA(cudaStreamWaitEvent(st, t->ev_read[b], 0));   // the loss slots of buffer b have been read back
    const int64_t n = len(k);
    for (int64_t j = 0; j < n; ++j) {
      const int32_t* uid = t->h_ids[b] + j * per_step;
      io.loss_out_dev = t->h_loss[b] + j * R;
      // L2 prefetch hint: the ids of the next step (a hint: if the next chunk has not landed yet the kernel prefetches
      // stale rows, never faults)
      const bool more = (j + 1 < n) || (k + 1 < n_chunks);
      t->hint_next_uid = !more ? nullptr : (j + 1 < n ? uid + per_step : t->h_ids[bn]);
      t->hint_next_cid = !more ? nullptr : (j + 1 < n ? uid + cid_off + per_step : t->h_ids[bn] + cid_off);
      const int rc = nncf_train_steps(t, tables, uid, uid + cid_off, 1, &io, st);
      t->hint_next_uid = t->hint_next_cid = nullptr;
      if (rc) return rc;
    }
    NNCF_CUDA(cudaEventRecord(t->ev_step[b], st));                     // chunk k has finished: its losses may leave, its id buffer is free
    if (loss_out_host) {                                               // the chunk's n x R losses -> host
      NNCF_CUDA(cudaStreamWaitEvent(t->s_d2h, t->ev_step[b], 0));
      NNCF_CUDA(cudaMemcpyAsync(loss_out_host + first(k) * R, t->h_loss[b], n * R * sizeof(float), cudaMemcpyDeviceToHost, t->s_d2h));
      NNCF_CUDA(cudaEventRecord(t->ev_read[b], t->s_d2h));
    }
    // ids of the chunks ahead: two in flight, enqueued BEHIND this chunk's launches (the device never waits for the host
    // to finish copy calls before it gets a step to run)
    if (k + 2 < n_chunks) if (int rc = enqueue_ids(k + 2)) return rc;
  }
  const double t_enq = trace ? now_us() : 0.0;
  NNCF_CUDA(cudaStreamSynchronize(st));
  const double t_sync = trace ? now_us() : 0.0;
  NNCF_CUDA(cudaStreamSynchronize(t->s_d2h));
  if (trace) fprintf(stderr, "[nncf_train_steps_host] %lld steps, %lld chunks (max %d): enqueue %.1f us, compute-stream sync +%.1f us, loss-stream sync +%.1f us\n",
                     (long long)n_steps, (long long)n_chunks, CH, t_enq - t_in, t_sync - t_enq, now_us() - t_sync);
  return NNCF_OK;
}

// =================================================================================================
// stand-alone row gather / sparse row update (used by framework towers and by the row-sharded multi-GPU path:
// owners gather rows for their peers and apply the gradients they receive back)
// =================================================================================================
namespace nncf {
__global__ void __launch_bounds__(256)
gather_plain_kernel(const float* __restrict__ table, int d, const int32_t* __restrict__ ids, int64_t n, float* __restrict__ out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * 8 + warp;
  if (row >= n) return;
  const float* src = table + (int64_t)ids[row] * d;
  float* dst = out + row * d;
  for (int c = lane; c < d; c += 32) dst[c] = __ldg(src + c);
}
}  // namespace nncf

struct nncf_updater {
  int optimizer;
  float lr, beta1, beta2, eps;
  int64_t t = 0;
  int32_t* owner = nullptr;
  int64_t owner_n = 0;
};

extern "C" int nncf_gather_rows(const float* table_dev, int dim, const int32_t* ids_dev, int64_t n, float* out_dev,
                                void* stream) {
  NNCF_CHECK_ARG(n >= 0 && dim >= 1, "nncf_gather_rows: bad sizes");
  if (n == 0) return NNCF_OK;
  NNCF_CHECK_ARG(table_dev && ids_dev && out_dev, "nncf_gather_rows: null argument");
  gather_plain_kernel<<<ceil_div(n, 8), 256, 0, (cudaStream_t)stream>>>(table_dev, dim, ids_dev, n, out_dev);
  NNCF_LAUNCH_OK();
  return NNCF_OK;
}

extern "C" int nncf_updater_create(int optimizer, float learn_rate, float beta1, float beta2, float epsilon,
                                   nncf_updater_t** out) {
  NNCF_CHECK_ARG(out, "nncf_updater_create: null argument");
  NNCF_CHECK_ARG(optimizer == NNCF_OPT_SGD || optimizer == NNCF_OPT_LAZY_ADAM, "nncf_updater_create: optimizer must be SGD or LAZY_ADAM");
  auto* u = new nncf_updater();
  u->optimizer = optimizer; u->lr = learn_rate; u->beta1 = beta1; u->beta2 = beta2; u->eps = epsilon;
  *out = u;
  return NNCF_OK;
}
extern "C" int nncf_updater_destroy(nncf_updater_t* u) {
  if (!u) return NNCF_OK;
  if (u->owner) cudaFree(u->owner);
  delete u;
  return NNCF_OK;
}
extern "C" int nncf_updater_begin_step(nncf_updater_t* u) {
  NNCF_CHECK_ARG(u, "nncf_updater_begin_step: null updater");
  u->t += 1;
  return NNCF_OK;
}
extern "C" int nncf_updater_apply(nncf_updater_t* u, float* table_dev, float* m_dev, float* v_dev, int64_t n_table_rows,
                                  int dim, const int32_t* ids_dev, int64_t n, float* grads_dev, void* stream) {
  NNCF_CHECK_ARG(u && table_dev, "nncf_updater_apply: null argument");
  NNCF_CHECK_ARG(n >= 0 && n < (int64_t)0x7fffffff && dim >= 1, "nncf_updater_apply: bad sizes");
  if (n == 0) return NNCF_OK;
  NNCF_CHECK_ARG(ids_dev && grads_dev, "nncf_updater_apply: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  const int cnt = static_cast<int>(n);
  if (u->optimizer == NNCF_OPT_SGD) {
    rows_sgd_kernel<<<dim3(ceil_div(n, 8), 1), 256, 0, st>>>(grads_dev, ids_dev, 0, cnt, dim, u->lr, table_dev);
    NNCF_LAUNCH_OK();
    return NNCF_OK;
  }
  NNCF_CHECK_ARG(m_dev && v_dev, "nncf_updater_apply: lazy Adam needs m / v");
  NNCF_CHECK_ARG(u->t >= 1, "nncf_updater_apply: call nncf_updater_begin_step first");
  if (ensure_owner(&u->owner, &u->owner_n, n_table_rows, st)) return NNCF_ECUDA;
  const double b1t = pow((double)u->beta1, (double)u->t), b2t = pow((double)u->beta2, (double)u->t);
  const float lr_t = (float)(u->lr * sqrt(1.0 - b2t) / (1.0 - b1t));
  dim3 g1(ceil_div(n, 256), 1), g8(ceil_div(n, 8), 1);
  adam_owner_kernel<<<g1, 256, 0, st>>>(ids_dev, 0, cnt, nullptr, 0, u->owner);
  NNCF_LAUNCH_OK();
  adam_combine_kernel<<<g8, 256, 0, st>>>(ids_dev, 0, cnt, nullptr, 0, dim, u->owner, grads_dev);
  NNCF_LAUNCH_OK();
  adam_apply_kernel<<<g8, 256, 0, st>>>(ids_dev, 0, cnt, nullptr, 0, dim, dim, u->owner, grads_dev, table_dev, m_dev, v_dev,
                                        lr_t, u->beta1, u->beta2, u->eps, nullptr);
  NNCF_LAUNCH_OK();
  adam_reset_kernel<<<g1, 256, 0, st>>>(ids_dev, 0, cnt, nullptr, u->owner);
  NNCF_LAUNCH_OK();
  return NNCF_OK;
}
