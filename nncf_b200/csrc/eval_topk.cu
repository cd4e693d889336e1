// eval_topk.cu — whole@k evaluation: users x candidate-items score GEMM with a per-row top-k epilogue (the
// score matrix never reaches memory), top-k merge across item splits, ranking metrics from CSR truth, and the
// given@k pair scorer / per-user AP+AUC.
//
// ref: utils/objectives.py:296-321 (test_eval_mat), :333-370 (evaluate_mat), :231-294 (given);
//      utils/metrics_ranking.py:6-35 (eval_multiple), :38-61 (eval_multiple_original).
// Tie rule (declared realisation of the reference's random tie-break): score descending, then lowest column.
#include <cmath>
#include "common.cuh"
#include "sm100.cuh"

namespace nncf {

// ------------------------------------------------------------------------------------------------
// 64-bit ordering keys: larger key = better candidate.  key 0 = empty slot.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t make_key(float s, uint32_t col) {
  s += 0.0f;                                   // -0 -> +0 so that equal scores tie on the column only
  uint32_t b = __float_as_uint(s);
  b ^= (b >> 31) ? 0xFFFFFFFFu : 0x80000000u;
  return (static_cast<uint64_t>(b) << 32) | static_cast<uint64_t>(0xFFFFFFFFu - col);
}
__device__ __forceinline__ float key_score(uint64_t k) {
  uint32_t b = static_cast<uint32_t>(k >> 32);
  b = (b & 0x80000000u) ? (b ^ 0x80000000u) : ~b;
  return __uint_as_float(b);
}
__device__ __forceinline__ uint32_t key_col(uint64_t k) { return 0xFFFFFFFFu - static_cast<uint32_t>(k); }

// warp-cooperative bitonic sort, descending, n a power of two (<= 256), data in shared memory
__device__ __forceinline__ void warp_bitonic_desc(uint64_t* a, int n, int lane) {
  for (int k = 2; k <= n; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = lane; t < (n >> 1); t += 32) {
        const int lo = ((t & ~(j - 1)) << 1) | (t & (j - 1));   // index with bit j clear
        const int hi = lo | j;
        const bool desc = ((lo & k) == 0);
        const uint64_t x = a[lo], y = a[hi];
        const bool sw = desc ? (x < y) : (x > y);
        if (sw) { a[lo] = y; a[hi] = x; }
      }
      __syncwarp();
    }
  }
}

constexpr int kPend = 64;          // pending candidates per row (smem)
constexpr int kMaxKP = 128;        // largest padded k

struct TopkShared {
  uint64_t* pend;     // [128][kPend]
  uint64_t* scratch;  // [4 warps][256]
};

// Merge row `il`'s pending candidates into its sorted list in global memory; returns the new threshold key.
__device__ __forceinline__ uint64_t flush_row(uint64_t* list, int KP, int k, int sortn, const uint64_t* pend_row, int cnt,
                                           uint64_t* scratch, int lane) {
  for (int t = lane; t < sortn; t += 32) {
    uint64_t v = 0;
    if (t < KP) v = list[t];
    else if (t - KP < cnt) v = pend_row[t - KP];
    scratch[t] = v;
  }
  __syncwarp();
  warp_bitonic_desc(scratch, sortn, lane);
  for (int t = lane; t < KP; t += 32) list[t] = (t < k) ? scratch[t] : 0ull;
  const uint64_t kth = scratch[k - 1];
  __syncwarp();
  return kth;
}

// One epilogue step: thread (= row) looks at 32 consecutive scores of its row.
struct RowState {
  float thr;          // score of the current k-th best (-inf until k candidates are known)
  uint64_t thr_key;   // its full ordering key (0 = none): ties are decided on keys, so tiles may arrive in any order
  int cnt;
};

__device__ __forceinline__ void consider32(const float* v, uint32_t col0, uint32_t col_end, RowState& st,
                                           uint64_t* pend_row, bool row_ok) {
  // maxima of the four contiguous groups of 8 (independent chains; pairs fold into 3-input FMNMX3).  Almost every
  // chunk has a candidate in SOME lane of the warp, so the slow path must be cheap: it only walks the groups whose
  // maximum reaches the threshold instead of all 32 elements.
  float gm[4];
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    const float* w = v + g * 8;
    gm[g] = fmaxf(fmaxf(fmaxf(w[0], w[1]), fmaxf(w[2], w[3])), fmaxf(fmaxf(w[4], w[5]), fmaxf(w[6], w[7])));
  }
  const float m = fmaxf(fmaxf(gm[0], gm[1]), fmaxf(gm[2], gm[3]));
  if (row_ok && m >= st.thr) {
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      if (gm[g] >= st.thr) {
#pragma unroll
        for (int t = g * 8; t < g * 8 + 8; ++t) {
          if (v[t] >= st.thr && col0 + t < col_end) {
            const uint64_t key = make_key(v[t], col0 + t);
            if (key > st.thr_key) { pend_row[st.cnt] = key; ++st.cnt; }
          }
        }
      }
    }
  }
}

__device__ __forceinline__ void flush_if_needed(RowState& st, uint64_t* list_base /*row 0 of this warp's quadrant*/,
                                                int64_t list_row_stride, int KP, int k, int sortn, TopkShared sh, int q,
                                                int lane, bool force) {
  unsigned need = __ballot_sync(0xffffffffu, force ? (st.cnt > 0) : (st.cnt > kPend - 32));
  while (need) {
    const int rl = __ffs(need) - 1;
    need &= need - 1;
    const int cnt = __shfl_sync(0xffffffffu, st.cnt, rl);
    const uint64_t kth = flush_row(list_base + rl * list_row_stride, KP, k, sortn, sh.pend + (q * 32 + rl) * kPend, cnt,
                                   sh.scratch + q * 256, lane);
    if (lane == rl) { st.thr_key = kth; st.thr = kth ? key_score(kth) : -INFINITY; st.cnt = 0; }
  }
}

// ------------------------------------------------------------------------------------------------
// fp32 rows -> bf16 tile image (zero padded)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
rows_to_img_kernel(const float* __restrict__ rows, int64_t n, int d, int nsub, uint8_t* __restrict__ img, int64_t n_pad) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * 8 + warp;
  if (row >= n_pad) return;
  uint8_t* blk = img + (row >> 7) * nsub * kSubBytes;
  for (int m = 0; m < nsub; ++m) {
    const int c = m * 64 + 2 * lane;
    float x0 = 0.0f, x1 = 0.0f;
    if (row < n) {
      if (c < d) x0 = __ldg(rows + row * d + c);
      if (c + 1 < d) x1 = __ldg(rows + row * d + c + 1);
    }
    *reinterpret_cast<uint32_t*>(blk + m * kSubBytes + sw128_offset(row & 127, 2 * lane)) = pack_bf16x2(x0, x1);
  }
}

// ------------------------------------------------------------------------------------------------
// tcgen05 eval kernel.  CTA (ub, sp): 128 users x item tiles [tile_begin, tile_end) of split sp.
//   warp 0 producer, warp 1 MMA issuer, warps 2..5 epilogue.  4 S accumulators of 128 columns in TMEM.
// ------------------------------------------------------------------------------------------------
struct EvalArgs {
  const uint8_t* Uimg; const uint8_t* Vimg;   // tile images
  const float* Uf; const float* Vf;           // fp32 rows (SIMT path)
  int64_t n_users, n_items;
  int d, dp;
  int k, KP, sortn;
  int nsplit, tiles_per_split;
  uint64_t* lists;                            // [n_user_blocks*128][nsplit][KP]
};

template <int NSUB>
struct EvalCfg {
  static constexpr int kStages = NSUB <= 2 ? 3 : (NSUB == 3 ? 2 : 1);
  static constexpr int kSBufs = 4;
  static constexpr size_t kSmemBytes = (size_t)NSUB * kSubBytes * (1 + kStages) + 128 * kPend * 8 + 4 * 256 * 8 +
                                       1024 + 256;
};

template <int NSUB>
__global__ void __launch_bounds__(192, 1)
eval_topk_tc_kernel(EvalArgs a) {
  using C = EvalCfg<NSUB>;
  constexpr int DP = 64 * NSUB;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sU = smem;
  uint8_t* sV = sU + NSUB * kSubBytes;
  TopkShared sh;
  sh.pend = reinterpret_cast<uint64_t*>(sV + C::kStages * NSUB * kSubBytes);
  sh.scratch = sh.pend + 128 * kPend;
  uint64_t* bars = sh.scratch + 4 * 256;
  uint64_t* u_full = bars;
  uint64_t* v_full = bars + 1;    // [3]
  uint64_t* v_empty = bars + 4;   // [3]
  uint64_t* s_full = bars + 7;    // [4]
  uint64_t* s_empty = bars + 11;  // [4]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 15);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ub = blockIdx.x, sp = blockIdx.y;
  const int n_tiles = static_cast<int>((a.n_items + 127) >> 7);
  const int t0 = sp * a.tiles_per_split;
  const int t1 = min(n_tiles, t0 + a.tiles_per_split);
  const int nj = max(0, t1 - t0);
  // CTAs sweep the item tiles in rotated order so that at any moment they read DIFFERENT tiles (no L2 hot spot on
  // one tile); the top-k logic is order independent (ties are decided on full keys)
  const int rot = nj > 0 ? static_cast<int>((static_cast<unsigned>(ub) * 37u + static_cast<unsigned>(sp) * 11u) % static_cast<unsigned>(nj)) : 0;

  if (tid == 0) {
    mbar_init(u_full, 1);
    for (int s = 0; s < 3; ++s) { mbar_init(&v_full[s], 1); mbar_init(&v_empty[s], 1); }
    for (int s = 0; s < 4; ++s) { mbar_init(&s_full[s], 1); mbar_init(&s_empty[s], kEpiWarps); }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    if (lane == 0 && nj > 0) {
      const uint8_t* gU = a.Uimg + (size_t)ub * NSUB * kSubBytes;
      mbar_expect_tx(u_full, NSUB * kSubBytes);
      for (int s = 0; s < NSUB; ++s) bulk_g2s(sU + s * kSubBytes, gU + (size_t)s * kSubBytes, kSubBytes, u_full);
      for (int j = 0; j < nj; ++j) {
        const int st = j % C::kStages;
        mbar_wait(&v_empty[st], ((j / C::kStages) & 1) ^ 1);
        mbar_expect_tx(&v_full[st], NSUB * kSubBytes);
        const int jj = (j + rot) % nj;
        const uint8_t* gV = a.Vimg + (size_t)(t0 + jj) * NSUB * kSubBytes;
        for (int s = 0; s < NSUB; ++s)
          bulk_g2s(sV + (st * NSUB + s) * kSubBytes, gV + (size_t)s * kSubBytes, kSubBytes, &v_full[st]);
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && nj > 0) {
      const uint32_t idesc = make_idesc_bf16(128, 128, 0, 0);
      mbar_wait(u_full, 0);
      for (int j = 0; j < nj; ++j) {
        const int st = j % C::kStages, sb = j % C::kSBufs;
        mbar_wait(&v_full[st], (j / C::kStages) & 1);
        mbar_wait(&s_empty[sb], ((j / C::kSBufs) & 1) ^ 1);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < DP / 16; ++k) {
          const uint64_t ad = make_smem_desc(smem_u32(sU + (k >> 2) * kSubBytes) + (k & 3) * 32, 16, 1024);
          const uint64_t bd = make_smem_desc(smem_u32(sV + (st * NSUB + (k >> 2)) * kSubBytes) + (k & 3) * 32, 16, 1024);
          umma_bf16(tmem + sb * 128, ad, bd, idesc, k > 0);
        }
        umma_commit(&s_full[sb]);
        umma_commit(&v_empty[st]);
      }
    }
  } else {
    const int q = warp & 3;
    const int il = q * 32 + lane;
    const int64_t urow = (int64_t)ub * 128 + il;
    const bool row_ok = urow < a.n_users;
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    RowState rs; rs.thr = -INFINITY; rs.thr_key = 0; rs.cnt = 0;
    uint64_t* pend_row = sh.pend + il * kPend;
    const int64_t lstride = (int64_t)a.nsplit * a.KP;
    uint64_t* list_base = a.lists + ((int64_t)ub * 128 + q * 32) * lstride + (int64_t)sp * a.KP;
    const uint32_t col_end = static_cast<uint32_t>(a.n_items);
    for (int j = 0; j < nj; ++j) {
      const int sb = j % C::kSBufs;
      mbar_wait(&s_full[sb], (j / C::kSBufs) & 1);
      tc_fence_after();
      // pull the whole 128-column row of the tile into registers with four back-to-back TMEM loads, hand the
      // accumulator back to the MMA warp immediately, then filter from registers
      float v[4][32];
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_ld32(tmem + lane_addr + sb * 128 + c * 32, v[c]);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_empty[sb]);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        consider32(v[c], static_cast<uint32_t>((t0 + (j + rot) % nj) * 128 + c * 32), col_end, rs, pend_row, row_ok);
        flush_if_needed(rs, list_base, lstride, a.KP, a.k, a.sortn, sh, q, lane, false);
      }
    }
    flush_if_needed(rs, list_base, lstride, a.KP, a.k, a.sortn, sh, q, lane, true);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// ------------------------------------------------------------------------------------------------
// CUDA-core fp32 eval kernel (precision = fp32): 128 threads, thread = user row, 32 item columns at a time.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
eval_topk_simt_kernel(EvalArgs a) {
  extern __shared__ __align__(16) uint8_t smem_raw2[];
  const int dp = a.dp;
  float* Vs = reinterpret_cast<float*>(smem_raw2);           // [32][dp]
  TopkShared sh;
  sh.pend = reinterpret_cast<uint64_t*>(Vs + 32 * dp);
  sh.scratch = sh.pend + 128 * kPend;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ub = blockIdx.x, sp = blockIdx.y;
  const int n_tiles = static_cast<int>((a.n_items + 127) >> 7);
  const int t0 = sp * a.tiles_per_split;
  const int t1 = min(n_tiles, t0 + a.tiles_per_split);
  const int il = tid, q = warp;
  const int64_t urow = (int64_t)ub * 128 + il;
  const bool row_ok = urow < a.n_users;
  RowState rs; rs.thr = -INFINITY; rs.thr_key = 0; rs.cnt = 0;
  uint64_t* pend_row = sh.pend + il * kPend;
  const int64_t lstride = (int64_t)a.nsplit * a.KP;
  uint64_t* list_base = a.lists + ((int64_t)ub * 128 + q * 32) * lstride + (int64_t)sp * a.KP;
  const uint32_t col_end = static_cast<uint32_t>(a.n_items);
  for (int64_t c0 = (int64_t)t0 * 128; c0 < (int64_t)t1 * 128 && c0 < a.n_items; c0 += 32) {
    __syncthreads();
    for (int idx = tid; idx < 32 * dp; idx += 128) {
      const int rr = idx / dp, c = idx - rr * dp;
      const int64_t it = c0 + rr;
      Vs[rr * dp + c] = (it < a.n_items && c < a.d) ? a.Vf[it * a.d + c] : 0.0f;
    }
    __syncthreads();
    float acc[32];
#pragma unroll
    for (int t = 0; t < 32; ++t) acc[t] = 0.0f;
    // the user's row is read straight from global memory (each thread walks its own row; L1 serves the lines)
    const float* u = a.Uf + (row_ok ? urow : 0) * a.d;
    for (int kk = 0; kk < dp; kk += 4) {
      const float u0 = (kk < a.d) ? __ldg(u + kk) : 0.0f, u1 = (kk + 1 < a.d) ? __ldg(u + kk + 1) : 0.0f;
      const float u2 = (kk + 2 < a.d) ? __ldg(u + kk + 2) : 0.0f, u3 = (kk + 3 < a.d) ? __ldg(u + kk + 3) : 0.0f;
#pragma unroll
      for (int t = 0; t < 32; ++t) {
        const float4 vv = *reinterpret_cast<const float4*>(Vs + t * dp + kk);
        acc[t] = fmaf(u0, vv.x, acc[t]);
        acc[t] = fmaf(u1, vv.y, acc[t]);
        acc[t] = fmaf(u2, vv.z, acc[t]);
        acc[t] = fmaf(u3, vv.w, acc[t]);
      }
    }
    consider32(acc, static_cast<uint32_t>(c0), col_end, rs, pend_row, row_ok);
    flush_if_needed(rs, list_base, lstride, a.KP, a.k, a.sortn, sh, q, lane, false);
  }
  flush_if_needed(rs, list_base, lstride, a.KP, a.k, a.sortn, sh, q, lane, true);
}

// merge the per-split sorted lists of each user and decode: one warp per user
__global__ void __launch_bounds__(128)
topk_merge_kernel(const uint64_t* __restrict__ lists, int64_t n_users, int nsplit, int KP, int k, int sortn2,
                  int32_t* __restrict__ ids, float* __restrict__ scores) {
  __shared__ uint64_t scratch_all[4][256];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t u = (int64_t)blockIdx.x * 4 + warp;
  if (u >= n_users) return;
  uint64_t* sc = scratch_all[warp];
  const uint64_t* base = lists + u * nsplit * KP;
  for (int t = lane; t < KP; t += 32) sc[t] = base[t];
  for (int t = KP + lane; t < sortn2; t += 32) sc[t] = 0;
  __syncwarp();
  for (int s = 1; s < nsplit; ++s) {
    for (int t = lane; t < KP; t += 32) sc[KP + t] = base[(int64_t)s * KP + t];
    __syncwarp();
    warp_bitonic_desc(sc, sortn2, lane);
    for (int t = KP + lane; t < sortn2; t += 32) sc[t] = 0;
    __syncwarp();
  }
  for (int t = lane; t < k; t += 32) {
    const uint64_t key = sc[t];
    ids[u * k + t] = key ? static_cast<int32_t>(key_col(key)) : -1;
    scores[u * k + t] = key ? key_score(key) : -INFINITY;
  }
}

// ------------------------------------------------------------------------------------------------
// metrics.   ref: utils/metrics_ranking.py:14-33; users with no relevant candidate are excluded
// (utils/objectives.py:316).  One thread per user; deterministic two-level fp64 reduction.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
metrics_kernel(const int32_t* __restrict__ ids, int64_t n_users, int k, const int64_t* __restrict__ indptr,
               const int32_t* __restrict__ cols, float* __restrict__ per_user, double* __restrict__ partial) {
  const int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  double ap = 0.0, rc = 0.0, pr = 0.0, kept = 0.0;
  if (u < n_users) {
    const int64_t b = indptr[u], e = indptr[u + 1];
    const int64_t nhits = e - b;
    if (nhits > 0) {
      double run = 0.0, sumap = 0.0;
      for (int i = 0; i < k; ++i) {
        const int32_t c = ids[u * k + i];
        if (c < 0) continue;
        int64_t lo = b, hi = e;
        while (lo < hi) {
          const int64_t mid = (lo + hi) >> 1;
          if (cols[mid] < c) lo = mid + 1; else hi = mid;
        }
        if (lo < e && cols[lo] == c) { run += 1.0; sumap += run / (i + 1.0); }
      }
      const double denom = nhits < k ? static_cast<double>(nhits) : static_cast<double>(k);
      ap = sumap / denom; rc = run / static_cast<double>(nhits); pr = run / k; kept = 1.0;
    }
    if (per_user) {
      per_user[u * 3 + 0] = static_cast<float>(ap);
      per_user[u * 3 + 1] = static_cast<float>(rc);
      per_user[u * 3 + 2] = static_cast<float>(pr);
    }
  }
  __shared__ double sred[4][256];
  sred[0][threadIdx.x] = ap; sred[1][threadIdx.x] = rc; sred[2][threadIdx.x] = pr; sred[3][threadIdx.x] = kept;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s)
      for (int m = 0; m < 4; ++m) sred[m][threadIdx.x] += sred[m][threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x < 4) partial[(int64_t)blockIdx.x * 4 + threadIdx.x] = sred[threadIdx.x][0];
}
__global__ void metrics_final_kernel(const double* __restrict__ partial, int nblocks, double* sums) {
  const int m = threadIdx.x;
  if (m < 4) {
    double s = 0.0;
    for (int b = 0; b < nblocks; ++b) s += partial[(int64_t)b * 4 + m];
    sums[m] = s;
  }
}

// ------------------------------------------------------------------------------------------------
// given@k: pair scores and per-group AP / AUC
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
score_pairs_kernel(const float* __restrict__ EU, const float* __restrict__ EV, int d, const int32_t* __restrict__ uid,
                   const int32_t* __restrict__ cid, int64_t n, float* __restrict__ out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t p = (int64_t)blockIdx.x * 8 + warp;
  if (p >= n) return;
  const float* u = EU + (int64_t)uid[p] * d;
  const float* v = EV + (int64_t)cid[p] * d;
  float acc = 0.0f;
  for (int c = lane; c < d; c += 32) acc = fmaf(__ldg(u + c), __ldg(v + c), acc);
  acc = warp_sum(acc);
  if (lane == 0) out[p] = acc;
}

// One warp per group.  rank-by-counting (O(n^2 / 32)); groups in the 'given' test lists are small.
__global__ void __launch_bounds__(128)
eval_given_kernel(const float* __restrict__ scores, const int32_t* __restrict__ truth, const int64_t* __restrict__ indptr,
                  int64_t n_groups, float* __restrict__ out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t gidx = (int64_t)blockIdx.x * 4 + warp;
  if (gidx >= n_groups) return;
  const int64_t b = indptr[gidx], e = indptr[gidx + 1];
  const int64_t n = e - b;
  double ap_acc = 0.0, rank_sum = 0.0;
  int64_t npos = 0;
  for (int64_t i = lane; i < n; i += 32) {
    if (truth[b + i] == 0) continue;
    ++npos;
    const float si = scores[b + i];
    // descending position (ties: lower index first) and number of positives at or before it
    int64_t better = 0, better_pos = 0, less = 0, equal = 0;
    for (int64_t j = 0; j < n; ++j) {
      const float sj = scores[b + j];
      const bool bt = (sj > si) || (sj == si && j < i);
      better += bt;
      better_pos += (bt && truth[b + j] != 0);
      less += (sj < si);
      equal += (sj == si);
    }
    ap_acc += static_cast<double>(better_pos + 1) / static_cast<double>(better + 1);
    rank_sum += static_cast<double>(less) + 0.5 * static_cast<double>(equal + 1);   // average rank (1-based)
  }
  for (int o = 16; o > 0; o >>= 1) {
    ap_acc += __shfl_xor_sync(0xffffffffu, ap_acc, o);
    rank_sum += __shfl_xor_sync(0xffffffffu, rank_sum, o);
    npos += __shfl_xor_sync(0xffffffffu, npos, o);
  }
  if (lane == 0) {
    const int64_t nneg = n - npos;
    out[gidx * 2 + 0] = npos > 0 ? static_cast<float>(ap_acc / static_cast<double>(npos)) : 0.0f;
    out[gidx * 2 + 1] = (npos > 0 && nneg > 0)
                            ? static_cast<float>((rank_sum - 0.5 * npos * (npos + 1)) / (static_cast<double>(npos) * nneg))
                            : NAN;
  }
}

}  // namespace nncf

using namespace nncf;

static int kpad_of(int k) { return (k + 31) / 32 * 32; }
static int pow2_ge(int x) { int p = 1; while (p < x) p <<= 1; return p; }

struct EvalPlan {
  int dp, nsub, KP, sortn, nsplit, tiles_per_split, n_tiles;
  int64_t n_ub, users_pad, items_pad;
  size_t off_uimg, off_vimg, off_lists, off_end;
};

static int make_plan(int64_t n_users, int64_t n_items, int dim, int topk, int precision, EvalPlan* p) {
  NNCF_CHECK_ARG(n_users >= 1 && n_items >= 1, "eval: empty user or item set");
  NNCF_CHECK_ARG(dim >= 1 && dim <= 256, "eval: dim must be in [1, 256]");
  NNCF_CHECK_ARG(topk >= 1 && topk <= kMaxKP, "[ERROR] eval_topk must be in [1, 128] when eval_scheme=whole");
  NNCF_CHECK_ARG(n_items < (int64_t)0x7fffffff, "eval: too many candidate items");
  p->dp = (dim + 63) / 64 * 64;
  p->nsub = p->dp / 64;
  p->KP = kpad_of(topk);
  p->sortn = pow2_ge(p->KP + kPend);
  p->n_ub = (n_users + 127) / 128;
  p->users_pad = p->n_ub * 128;
  p->n_tiles = static_cast<int>((n_items + 127) / 128);
  p->items_pad = (int64_t)p->n_tiles * 128;
  // enough CTAs to fill 148 SMs twice when the user dimension alone cannot
  int nsplit = 1;
  if (p->n_ub < 296) nsplit = static_cast<int>((296 + p->n_ub - 1) / p->n_ub);
  if (nsplit > p->n_tiles) nsplit = p->n_tiles;
  if (nsplit > 64) nsplit = 64;
  p->tiles_per_split = (p->n_tiles + nsplit - 1) / nsplit;
  p->nsplit = (p->n_tiles + p->tiles_per_split - 1) / p->tiles_per_split;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 1023) / 1024 * 1024; return o; };
  p->off_uimg = take(precision == NNCF_PREC_BF16 ? (size_t)p->users_pad * p->dp * 2 : 0);
  p->off_vimg = take(precision == NNCF_PREC_BF16 ? (size_t)p->items_pad * p->dp * 2 : 0);
  p->off_lists = take((size_t)p->users_pad * p->nsplit * p->KP * 8);
  p->off_end = off + 1024;
  return 0;
}

extern "C" size_t nncf_eval_topk_workspace_bytes(int64_t n_users, int64_t n_items, int dim, int topk, int precision) {
  EvalPlan p;
  if (make_plan(n_users, n_items, dim, topk, precision, &p)) return 0;
  return p.off_end;
}

template <int NSUB>
static int launch_eval_tc(const EvalArgs& ea, const EvalPlan& p, cudaStream_t st) {
  using C = EvalCfg<NSUB>;
  NNCF_CUDA(cudaFuncSetAttribute(eval_topk_tc_kernel<NSUB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)C::kSmemBytes));
  eval_topk_tc_kernel<NSUB><<<dim3((unsigned)p.n_ub, p.nsplit), 192, C::kSmemBytes, st>>>(ea);
  NNCF_LAUNCH_OK();
  return 0;
}

extern "C" int nncf_eval_topk(const float* user_rows_dev, int64_t n_users, const float* item_rows_dev, int64_t n_items,
                              int dim, int topk, int precision, int32_t* topk_ids_dev, float* topk_scores_dev,
                              void* workspace_dev, size_t workspace_bytes, void* stream) {
  NNCF_CHECK_ARG(user_rows_dev && item_rows_dev && topk_ids_dev && topk_scores_dev && workspace_dev,
                 "nncf_eval_topk: null argument");
  NNCF_CHECK_ARG(precision == NNCF_PREC_FP32 || precision == NNCF_PREC_BF16, "unknown precision");
  EvalPlan p;
  if (int rc = make_plan(n_users, n_items, dim, topk, precision, &p)) return rc;
  NNCF_CHECK_ARG(workspace_bytes >= p.off_end, "nncf_eval_topk: workspace too small");
  NNCF_CHECK_ARG(n_users <= (int64_t)65535 * 128 * 512, "eval: too many users for one call");
  cudaStream_t st = (cudaStream_t)stream;
  uint8_t* ws = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(workspace_dev) + 1023) & ~uintptr_t(1023));
  EvalArgs ea{};
  ea.Uf = user_rows_dev; ea.Vf = item_rows_dev; ea.n_users = n_users; ea.n_items = n_items; ea.d = dim; ea.dp = p.dp;
  ea.k = topk; ea.KP = p.KP; ea.sortn = p.sortn; ea.nsplit = p.nsplit; ea.tiles_per_split = p.tiles_per_split;
  ea.lists = reinterpret_cast<uint64_t*>(ws + p.off_lists);
  NNCF_CUDA(cudaMemsetAsync(ea.lists, 0, (size_t)p.users_pad * p.nsplit * p.KP * 8, st));
  if (precision == NNCF_PREC_BF16) {
    ea.Uimg = ws + p.off_uimg; ea.Vimg = ws + p.off_vimg;
    rows_to_img_kernel<<<ceil_div(p.users_pad, 8), 256, 0, st>>>(user_rows_dev, n_users, dim, p.nsub,
                                                                 const_cast<uint8_t*>(ea.Uimg), p.users_pad);
    NNCF_LAUNCH_OK();
    rows_to_img_kernel<<<ceil_div(p.items_pad, 8), 256, 0, st>>>(item_rows_dev, n_items, dim, p.nsub,
                                                                 const_cast<uint8_t*>(ea.Vimg), p.items_pad);
    NNCF_LAUNCH_OK();
    int rc;
    switch (p.nsub) {
      case 1: rc = launch_eval_tc<1>(ea, p, st); break;
      case 2: rc = launch_eval_tc<2>(ea, p, st); break;
      case 3: rc = launch_eval_tc<3>(ea, p, st); break;
      default: rc = launch_eval_tc<4>(ea, p, st); break;
    }
    if (rc) return rc;
  } else {
    const size_t sm = ((size_t)32 * p.dp) * 4 + 128 * kPend * 8 + 4 * 256 * 8;
    NNCF_CUDA(cudaFuncSetAttribute(eval_topk_simt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    eval_topk_simt_kernel<<<dim3((unsigned)p.n_ub, p.nsplit), 128, sm, st>>>(ea);
    NNCF_LAUNCH_OK();
  }
  topk_merge_kernel<<<ceil_div(n_users, 4), 128, 0, st>>>(ea.lists, n_users, p.nsplit, p.KP, topk, pow2_ge(2 * p.KP),
                                                          topk_ids_dev, topk_scores_dev);
  NNCF_LAUNCH_OK();
  return NNCF_OK;
}

extern "C" int nncf_eval_metrics(const int32_t* topk_ids_dev, int64_t n_users, int topk, const int64_t* truth_indptr_dev,
                                 const int32_t* truth_cols_dev, float* per_user_dev, double* sums_dev, void* stream) {
  NNCF_CHECK_ARG(topk_ids_dev && truth_indptr_dev && truth_cols_dev && sums_dev, "nncf_eval_metrics: null argument");
  NNCF_CHECK_ARG(n_users >= 1 && topk >= 1, "nncf_eval_metrics: bad sizes");
  cudaStream_t st = (cudaStream_t)stream;
  const int nb = ceil_div(n_users, 256);
  double* partial = nullptr;
  NNCF_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&partial), (size_t)nb * 4 * sizeof(double), st));
  metrics_kernel<<<nb, 256, 0, st>>>(topk_ids_dev, n_users, topk, truth_indptr_dev, truth_cols_dev, per_user_dev, partial);
  NNCF_LAUNCH_OK();
  metrics_final_kernel<<<1, 32, 0, st>>>(partial, nb, sums_dev);
  NNCF_LAUNCH_OK();
  NNCF_CUDA(cudaFreeAsync(partial, st));
  return NNCF_OK;
}

extern "C" int nncf_score_pairs(const float* user_table_dev, const float* item_table_dev, int dim,
                                const int32_t* user_ids_dev, const int32_t* item_ids_dev, int64_t n_pairs,
                                float* scores_dev, void* stream) {
  NNCF_CHECK_ARG(user_table_dev && item_table_dev && user_ids_dev && item_ids_dev && scores_dev, "nncf_score_pairs: null argument");
  if (n_pairs == 0) return NNCF_OK;
  score_pairs_kernel<<<ceil_div(n_pairs, 8), 256, 0, (cudaStream_t)stream>>>(user_table_dev, item_table_dev, dim,
                                                                             user_ids_dev, item_ids_dev, n_pairs, scores_dev);
  NNCF_LAUNCH_OK();
  return NNCF_OK;
}

extern "C" int nncf_eval_given(const float* scores_dev, const int32_t* truth_dev, const int64_t* seg_indptr_dev,
                               int64_t n_groups, float* per_group_dev, void* stream) {
  NNCF_CHECK_ARG(scores_dev && truth_dev && seg_indptr_dev && per_group_dev, "nncf_eval_given: null argument");
  if (n_groups == 0) return NNCF_OK;
  eval_given_kernel<<<ceil_div(n_groups, 4), 128, 0, (cudaStream_t)stream>>>(scores_dev, truth_dev, seg_indptr_dev,
                                                                             n_groups, per_group_dev);
  NNCF_LAUNCH_OK();
  return NNCF_OK;
}
