// eval_topk.cu — whole@k evaluation: users x candidate-items score GEMM with a per-row top-k epilogue (the
// score matrix never reaches memory), top-k merge across item splits, ranking metrics from CSR truth, and the
// given@k pair scorer / per-user AP+AUC.
//
// ref: utils/objectives.py:296-321 (test_eval_mat), :333-370 (evaluate_mat), :231-294 (given);
//      utils/metrics_ranking.py:6-35 (eval_multiple), :38-61 (eval_multiple_original).
// Tie rule (declared realisation of the reference's random tie-break): score descending, then lowest column.
#include <cmath>
#include <cstdlib>
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
__device__ __forceinline__ uint64_t shfl_u64(uint64_t v, int src) {
  const uint32_t hi = __shfl_sync(0xffffffffu, static_cast<uint32_t>(v >> 32), src);
  const uint32_t lo = __shfl_sync(0xffffffffu, static_cast<uint32_t>(v), src);
  return (static_cast<uint64_t>(hi) << 32) | lo;
}
__device__ __forceinline__ uint64_t shfl_xor_u64(uint64_t v, int m) {
  const uint32_t hi = __shfl_xor_sync(0xffffffffu, static_cast<uint32_t>(v >> 32), m);
  const uint32_t lo = __shfl_xor_sync(0xffffffffu, static_cast<uint32_t>(v), m);
  return (static_cast<uint64_t>(hi) << 32) | lo;
}

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

constexpr int kMaxKP = 128;        // largest padded k

// ------------------------------------------------------------------------------------------------
// Streaming exact top-k, one row per thread of the epilogue warp (row = TMEM lane).
//   Every row keeps an UNSORTED set of its k best keys in shared memory plus, in registers, the smallest of them
//   (thr_key, at position minpos) — the exact running threshold, so the number of candidates is the minimum a
//   streaming scan can have (~k ln(I/k) per row).
//   Fast path per 128-column tile: the row maximum (tree of FMNMX3) against the threshold, ONE ballot per warp.
//   Hit path (some lane's maximum reaches its threshold — most tiles, because a warp covers 32 rows): warp-
//   COOPERATIVE, no per-lane divergent loops: the hitting lane parks a 32-score chunk in shared memory, lane t tests
//   column t, a ballot yields the candidates, and each candidate replaces the row's current minimum followed by a
//   warp arg-min over the k keys.  (Measured on B200: per-lane divergent candidate handling cost ~500 cycles per
//   candidate and a sort-based pending-buffer merge ~6k cycles per flush — together 3x the tensor-core pipeline.)
// ------------------------------------------------------------------------------------------------
struct RowState {
  float thr;          // score of thr_key (-inf while the set is not full)
  uint64_t thr_key;   // smallest key of the row's set (0 = an empty slot exists)
  int minpos;         // its position in the set
};

// chunk = 32 scores of lane L (already staged in `stage`), columns col0 .. col0+31
__device__ __forceinline__ void coop_chunk(int L, const float* stage, uint32_t col0, uint32_t col_end, RowState& st,
                                           uint64_t* list_L, int k, int lane) {
  const float x = stage[lane];
  const uint32_t col = col0 + lane;
  const uint64_t key = make_key(x, col);
  uint64_t tk = shfl_u64(st.thr_key, L);
  unsigned cm = __ballot_sync(0xffffffffu, (col < col_end) && (key > tk));
  while (cm) {
    const int t = __ffs(cm) - 1;
    cm &= cm - 1;
    const uint64_t kc = shfl_u64(key, t);
    if (kc > tk) {                                      // warp-uniform (tk may have risen since the ballot)
      const int mp = __shfl_sync(0xffffffffu, st.minpos, L);
      if (lane == 0) list_L[mp] = kc;
      __syncwarp();
      uint64_t best = ~0ull;
      int bpos = 0;
      for (int i = lane; i < k; i += 32) {
        const uint64_t kk = list_L[i];
        if (kk < best) { best = kk; bpos = i; }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const uint64_t ob = shfl_xor_u64(best, o);
        const int op = __shfl_xor_sync(0xffffffffu, bpos, o);
        if (ob < best || (ob == best && op < bpos)) { best = ob; bpos = op; }
      }
      tk = best;
      if (lane == L) { st.thr_key = best; st.minpos = bpos; st.thr = best ? key_score(best) : -INFINITY; }
    }
  }
}

// one 128-column tile held in registers as v[4][32]; lists = this warp's 32 rows x KP keys in shared memory
__device__ __forceinline__ void filter_tile(const float (&v)[4][32], uint32_t col0, uint32_t col_end, RowState& st, bool row_ok,
                                            uint64_t* lists_warp, int KP, int k, float* stage, int lane, unsigned own_mask,
                                            bool skip_hits = false) {
  float gm[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    float m8[4];
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const float* w = v[c] + g * 8;
      m8[g] = fmaxf(fmaxf(fmaxf(w[0], w[1]), fmaxf(w[2], w[3])), fmaxf(fmaxf(w[4], w[5]), fmaxf(w[6], w[7])));
    }
    gm[c] = fmaxf(fmaxf(m8[0], m8[1]), fmaxf(m8[2], m8[3]));
  }
  const float m = fmaxf(fmaxf(gm[0], gm[1]), fmaxf(gm[2], gm[3]));
  // several warps read the same 32 TMEM lanes; each handles only the rows it owns (own_mask), so the serial
  // candidate work of a quadrant is split between them while every row still has exactly one owner
  unsigned hits = __ballot_sync(0xffffffffu, row_ok && m >= st.thr) & own_mask;
  if (skip_hits) { if (hits == 0x12345678u) st.minpos = 1; st.thr = fmaxf(st.thr, m - 1.0f); return; }   // developer switch: fast path only
  while (hits) {
    const int L = __ffs(hits) - 1;
    hits &= hits - 1;
    const float thr_l = __shfl_sync(0xffffffffu, st.thr, L);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const float gmc = __shfl_sync(0xffffffffu, gm[c], L);
      if (gmc >= thr_l && col0 + c * 32 < col_end) {    // warp-uniform
        if (lane == L) {
#pragma unroll
          for (int t = 0; t < 32; t += 4)
            *reinterpret_cast<float4*>(stage + t) = make_float4(v[c][t], v[c][t + 1], v[c][t + 2], v[c][t + 3]);
        }
        __syncwarp();
        coop_chunk(L, stage, col0 + c * 32, col_end, st, lists_warp + L * KP, k, lane);
        __syncwarp();
      }
    }
  }
}

// same for ONE 32-column chunk (CUDA-core kernel)
__device__ __forceinline__ void filter_chunk(const float (&v)[32], uint32_t col0, uint32_t col_end, RowState& st, bool row_ok,
                                             uint64_t* lists_warp, int KP, int k, float* stage, int lane) {
  float m = v[0];
#pragma unroll
  for (int t = 1; t < 32; ++t) m = fmaxf(m, v[t]);
  unsigned hits = __ballot_sync(0xffffffffu, row_ok && m >= st.thr);
  while (hits) {
    const int L = __ffs(hits) - 1;
    hits &= hits - 1;
    if (lane == L) {
#pragma unroll
      for (int t = 0; t < 32; t += 4) *reinterpret_cast<float4*>(stage + t) = make_float4(v[t], v[t + 1], v[t + 2], v[t + 3]);
    }
    __syncwarp();
    coop_chunk(L, stage, col0, col_end, st, lists_warp + L * KP, k, lane);
    __syncwarp();
  }
}

// copy this warp's 32 unsorted sets to the global per-(user, slot) lists (sorted later by topk_merge_kernel)
__device__ __forceinline__ void store_lists(const uint64_t* lists_warp, int KP, uint64_t* glists_row0, int64_t gstride, int lane,
                                            int nrows = 32) {
  for (int r = 0; r < nrows; ++r)
    for (int i = lane; i < KP; i += 32) glists_row0[r * gstride + i] = lists_warp[r * KP + i];
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
// tcgen05 eval kernel.  CTA (ub, sp): 128 users x the item tiles of split sp.
//   warp 0 bulk-copy producer, warp 1 MMA issuer, warps 2..5 epilogue (TMEM lane quadrant = warp & 3);
//   4 S accumulators of 128 columns in TMEM so the MMA runs up to four tiles ahead of the filter.
//   CL > 1: thread-block cluster of CL CTAs (CL consecutive user blocks) sweeping the SAME item tiles: every CTA loads
//   1/CL of each V tile and multicasts it into all CL shared memories, so a tile crosses L2 -> SM once per CL*128 users.
// ------------------------------------------------------------------------------------------------
struct EvalArgs {
  const uint8_t* Uimg; const uint8_t* Vimg;   // tile images
  const float* Uf; const float* Vf;           // fp32 rows (CUDA-core path)
  int64_t n_users, n_items;
  int d, dp;
  int k, KP;
  int nsplit, tiles_per_split;
  int nstages;                                // V stages in flight (shared memory left after the top-k sets)
  uint64_t* lists;                            // [n_user_blocks*128][nsplit][KP]  unsorted k-best sets per split
  int prefetch_ahead;                         // second-generation kernel: L2 prefetch distance in tiles (0 = off)
  int cap;                                    // second-generation kernel: keys per row buffer (multiple of 32, >= KP + 32)
  int dbg_mode;                               // developer switch (env NNCF_EVAL_DBG): 1 = skip the filter, 2 = also the TMEM loads, 3 = row maxima + ballot only
  unsigned long long* stats;                  // developer counters (env NNCF_EVAL_STATS): [0] tiles x warps, [1] tiles with a hit, [2] hit rows,
                                              // [3] keys appended, [4] compactions, [5] cycles in the hit path, [6] cycles in the tile loop
  int dbg_pipe;                               // developer switch (env NNCF_EVAL_PIPE), second generation: bit 0 = no item copies after the first fill, bit 1 = no MMAs, bit 2 = exact (radix-select) compaction on the hot path
};

constexpr int kEvalRowGroups = 2;                 // epilogue warps per TMEM lane quadrant; each owns 32 / kEvalRowGroups rows
constexpr int kEvalEpiWarps = 4 * kEvalRowGroups;
constexpr int kEvalThreads = 64 + 32 * kEvalEpiWarps;
constexpr int kEvalCluster = 2;     // measured (37,888 users x 1M items, k = 50): CL=1 26.9 ms, CL=2 26.8 ms (pipeline alone 10.6 ms =
                                    // 914 TFLOP/s), CL=4 40.0 ms — a bigger cluster couples more CTAs to the slowest filter

__host__ __device__ inline size_t eval_smem_bytes(int nsub, int nstages, int KP) {
  return (size_t)nsub * kSubBytes * (1 + nstages) + (size_t)128 * KP * 8 + 8 * 32 * 4 + 1024 + 256;
}

template <int NSUB, int CL>
__global__ void __launch_bounds__(kEvalThreads, 1)
eval_topk_tc_kernel(EvalArgs a) {
  constexpr int DP = 64 * NSUB;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // (pointer arithmetic keeps the shared address space: LDS / STS, not generic LD / ST)
  const int nst = a.nstages;
  uint8_t* sU = smem;
  uint8_t* sV = sU + NSUB * kSubBytes;
  uint64_t* lists_sm = reinterpret_cast<uint64_t*>(sV + nst * NSUB * kSubBytes);   // [128][KP]
  float* stage_all = reinterpret_cast<float*>(lists_sm + 128 * a.KP);              // [8 warps][32]
  uint64_t* bars = reinterpret_cast<uint64_t*>(stage_all + 8 * 32);
  uint64_t* u_full = bars;
  uint64_t* v_full = bars + 1;    // [4]
  uint64_t* v_empty = bars + 5;   // [4]
  uint64_t* s_full = bars + 9;    // [4]
  uint64_t* s_empty = bars + 13;  // [4]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 17);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ub = blockIdx.x, sp = blockIdx.y;
  const int n_tiles = static_cast<int>((a.n_items + 127) >> 7);
  const int t0 = sp * a.tiles_per_split;
  const int t1 = min(n_tiles, t0 + a.tiles_per_split);
  const int nj = max(0, t1 - t0);
  // clusters sweep the item tiles in rotated order so that at any moment different clusters read DIFFERENT tiles; the
  // top-k logic is order independent (ties are decided on full keys)
  const int rot = nj > 0 ? static_cast<int>((static_cast<unsigned>(ub / CL) * 37u + static_cast<unsigned>(sp) * 11u) % static_cast<unsigned>(nj)) : 0;
  const uint32_t crank = CL > 1 ? cluster_ctarank() : 0u;
  constexpr uint16_t kMask = static_cast<uint16_t>((1u << CL) - 1u);

  for (int i = tid; i < 128 * a.KP; i += kEvalThreads) lists_sm[i] = 0ull;
  if (tid == 0) {
    mbar_init(u_full, 1);
    for (int s = 0; s < 4; ++s) { mbar_init(&v_full[s], 1); mbar_init(&v_empty[s], CL); }
    for (int s = 0; s < 4; ++s) { mbar_init(&s_full[s], 1); mbar_init(&s_empty[s], kEvalEpiWarps); }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();     // peers' barriers are initialised before anybody multicasts into this CTA
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    if (lane == 0 && nj > 0) {
      const uint8_t* gU = a.Uimg + (size_t)ub * NSUB * kSubBytes;
      mbar_expect_tx(u_full, NSUB * kSubBytes);
      for (int s = 0; s < NSUB; ++s) bulk_g2s(sU + s * kSubBytes, gU + (size_t)s * kSubBytes, kSubBytes, u_full);
      for (int j = 0; j < nj; ++j) {
        const int st = j % nst;
        mbar_wait(&v_empty[st], ((j / nst) & 1) ^ 1);
        mbar_expect_tx(&v_full[st], NSUB * kSubBytes);
        const int jj = (j + rot) % nj;
        const uint8_t* gV = a.Vimg + (size_t)(t0 + jj) * NSUB * kSubBytes;
        if (CL > 1) {
          constexpr uint32_t kSlice = NSUB * kSubBytes / CL;      // my share of the tile, delivered to every CTA
          bulk_g2s_multicast(sV + st * NSUB * kSubBytes + crank * kSlice, gV + (size_t)crank * kSlice, kSlice, &v_full[st], kMask);
        } else {
          for (int s = 0; s < NSUB; ++s)
            bulk_g2s(sV + (st * NSUB + s) * kSubBytes, gV + (size_t)s * kSubBytes, kSubBytes, &v_full[st]);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && nj > 0) {
      const uint32_t idesc = make_idesc_bf16(128, 128, 0, 0);
      mbar_wait(u_full, 0);
      for (int j = 0; j < nj; ++j) {
        const int st = j % nst, sb = j & 3;
        mbar_wait(&v_full[st], (j / nst) & 1);
        mbar_wait(&s_empty[sb], ((j >> 2) & 1) ^ 1);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < DP / 16; ++k) {
          const uint64_t ad = make_smem_desc(smem_u32(sU + (k >> 2) * kSubBytes) + (k & 3) * 32, 16, 1024);
          const uint64_t bd = make_smem_desc(smem_u32(sV + (st * NSUB + (k >> 2)) * kSubBytes) + (k & 3) * 32, 16, 1024);
          umma_bf16(tmem + sb * 128, ad, bd, idesc, k > 0);
        }
        umma_commit(&s_full[sb]);
        if (CL > 1) umma_commit_multicast(&v_empty[st], kMask);   // the stage is rewritten by all CL producers
        else umma_commit(&v_empty[st]);
      }
    }
  } else {
    const int q = warp & 3;
    const int il = q * 32 + lane;
    const int64_t urow = (int64_t)ub * 128 + il;
    const bool row_ok = urow < a.n_users;
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    RowState rs; rs.thr = -INFINITY; rs.thr_key = 0; rs.minpos = 0;
    uint64_t* lists_warp = lists_sm + (size_t)q * 32 * a.KP;
    float* stage = stage_all + (warp - 2) * 32;
    const int rg = (warp - 2) >> 2;
    constexpr int kOwn = 32 / kEvalRowGroups;
    const unsigned own_mask = (kOwn == 32 ? 0xffffffffu : ((1u << kOwn) - 1u)) << (rg * kOwn);
    const uint32_t col_end = static_cast<uint32_t>(a.n_items);
    for (int j = 0; j < nj; ++j) {
      const int sb = j & 3;
      mbar_wait(&s_full[sb], (j >> 2) & 1);
      tc_fence_after();
      // pull the whole 128-column row of the tile into registers with four back-to-back TMEM loads, hand the
      // accumulator back to the MMA warp immediately, then filter from registers
      float v[4][32];
      if (a.dbg_mode < 2) {
#pragma unroll
        for (int c = 0; c < 4; ++c) tmem_ld32(tmem + lane_addr + sb * 128 + c * 32, v[c]);
        tmem_ld_wait();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_empty[sb]);
      if (a.dbg_mode == 1 || a.dbg_mode == 2) { if (a.dbg_mode < 2 && v[0][0] == 123456.0f) rs.minpos = 1; continue; }
      filter_tile(v, static_cast<uint32_t>((t0 + (j + rot) % nj) * 128), col_end, rs, row_ok, lists_warp, a.KP, a.k, stage, lane, own_mask,
                  a.dbg_mode == 3);
    }
    __syncwarp();
    const int64_t gstride = (int64_t)a.nsplit * a.KP;
    store_lists(lists_warp + rg * kOwn * a.KP, a.KP,
                a.lists + ((int64_t)ub * 128 + q * 32 + rg * kOwn) * gstride + (int64_t)sp * a.KP, gstride, lane, kOwn);
  }
  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();     // nobody exits while a peer may still multicast into it or signal its barriers
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// ------------------------------------------------------------------------------------------------
// tcgen05 eval kernel, second generation (the default).  Measured on B200 with the kernel above (37,888 users x 1M
// items, d = 128, k = 50): the MMA pipeline ALONE needed 1,326 cycles per 128 x 128 x 128 tile against a 512-cycle
// tensor floor, because an SS-mode M = N = 128 MMA fetches 8 KiB of operands per 64 cycles = the whole 128 B/clk
// shared-memory port, on top of the bulk-copy fills; and the exact streaming insertion cost ~1,900 warp-cycles per
// candidate (a serial arg-min chain), 3,373 cycles per tile in total.  Changes:
//   * the user block is the A operand of every MMA of the CTA: it is written ONCE into tensor memory (DP / 2 columns,
//     two bf16 per column) and read from there (tcgen05.mma with a TMEM A operand), so shared memory serves only the
//     item tiles; 3 accumulators of 128 columns (3 x 128 + DP / 2 <= 512 columns);
//   * every accumulator is read from TMEM once: 4 epilogue warps, one per lane quadrant, 32 rows each (kEval2RowGroups);
//   * top-k by APPEND + lazy compaction: a row keeps up to CAP > k keys in shared memory and a threshold that is exact
//     as of its last compaction.  A 32-score chunk that reaches the threshold appends all its candidates with ONE
//     ballot (no per-candidate loop); when fewer than 32 slots are left the warp selects the row's k largest keys
//     (radix select on the ordered score bits in registers: 32 x REDUX, then ties by column) and raises the
//     threshold.  A stale threshold only admits extra candidates, never loses one, so the result is still exact:
//     score descending, then lowest column.  ~k (CAP - k)^-1 ln(I / k) compactions per row instead of an arg-min per
//     candidate.
// ------------------------------------------------------------------------------------------------
constexpr int kEval2RowGroups = 1;                // epilogue warps per TMEM lane quadrant.  Measured (37,888 x 1M, k = 50): 1 -> 19.5 ms,
                                                  // 2 (both read the accumulator, 16 rows each, row maxima duplicated) -> 21.9 ms
constexpr int kEval2EpiWarps = 4 * kEval2RowGroups;
constexpr int kEval2Threads = 64 + 32 * kEval2EpiWarps;
constexpr int kEval2Acc = 3;
constexpr int kEval2MaxNK = 6;                    // CAP <= 192 keys per row

__host__ __device__ inline size_t eval2_smem_bytes(int nsub, int nstages, int cap) {
  return (size_t)nsub * kSubBytes * nstages + (size_t)128 * (cap + 1) * 8 + kEval2EpiWarps * 32 * 4 + 1024 + 256;
}

// the row's n > k keys (buf[0..n)) -> its k largest at buf[0..k); returns the score of the k-th as the new threshold.
// Warp-cooperative; n, k warp-uniform; keys are unique (they embed the column).
__device__ __forceinline__ float compact_row(uint64_t* buf, int n, int k, int lane) {
  uint32_t hi[kEval2MaxNK], lo[kEval2MaxNK];
#pragma unroll
  for (int i = 0; i < kEval2MaxNK; ++i) {
    const int idx = lane + 32 * i;
    const uint64_t key = idx < n ? buf[idx] : 0ull;
    hi[i] = static_cast<uint32_t>(key >> 32); lo[i] = static_cast<uint32_t>(key);
  }
  __syncwarp();
  uint32_t T = 0;                                   // largest T with #(hi >= T) >= k  ==  high word of the k-th key
#pragma unroll 1
  for (int b = 31; b >= 0; --b) {
    const uint32_t c = T | (1u << b);
    int cnt = 0;
#pragma unroll
    for (int i = 0; i < kEval2MaxNK; ++i) cnt += (hi[i] >= c) ? 1 : 0;
    if (__reduce_add_sync(0xffffffffu, cnt) >= k) T = c;
  }
  int gt = 0, eq = 0;
#pragma unroll
  for (int i = 0; i < kEval2MaxNK; ++i) { gt += (hi[i] > T) ? 1 : 0; eq += (hi[i] == T) ? 1 : 0; }
  gt = __reduce_add_sync(0xffffffffu, gt);
  eq = __reduce_add_sync(0xffffffffu, eq);
  const int need = k - gt;                          // 1 <= need <= eq keys of score T survive
  uint32_t TL = 0;
  if (eq > need) {                                  // equal scores straddle the boundary: lowest columns (largest low words) win
#pragma unroll 1
    for (int b = 31; b >= 0; --b) {
      const uint32_t c = TL | (1u << b);
      int cnt = 0;
#pragma unroll
      for (int i = 0; i < kEval2MaxNK; ++i) cnt += (hi[i] == T && lo[i] >= c) ? 1 : 0;
      if (__reduce_add_sync(0xffffffffu, cnt) >= need) TL = c;
    }
  }
  int base = 0;
  const unsigned lt = (1u << lane) - 1u;
#pragma unroll
  for (int i = 0; i < kEval2MaxNK; ++i) {
    const bool keep = hi[i] > T || (hi[i] == T && lo[i] >= TL && hi[i] != 0u);
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    if (keep) buf[base + __popc(m & lt)] = (static_cast<uint64_t>(hi[i]) << 32) | lo[i];
    base += __popc(m);
  }
  __syncwarp();
  return key_score(static_cast<uint64_t>(T) << 32);
}

// Coarse compaction (the hot-path variant): one histogram round instead of a 32-step radix select.  The row's keys are
// bucketed by score into 32 equal bins between the row's smallest and largest score (monotone in the score), the bins are
// suffix-summed, and every key in or above the highest bin B whose suffix count reaches k is kept: kept >= k keys, all
// greater than every dropped key, so the smallest kept score is a valid lower bound of the row's k-th best and no key that
// can still belong to the top k is lost (exactness is preserved; the final per-row compaction is the exact one).
// ~300 cycles against ~2,500: a compaction holds up its whole CTA pair (an accumulator is recycled only when all 8 filter
// warps have read it), and with ~0.5 compactions per tile per pair that stall, not the per-candidate work, was what
// held whole@k at ~2,200 cycles per tile.  Falls back to the exact select when the bins cannot separate (ties).
// Returns the new threshold; *n_out = keys kept.
__device__ __forceinline__ float compact_row_coarse(uint64_t* buf, int n, int k, int max_keep, int lane, int* hist, int* n_out) {
  uint32_t hi[kEval2MaxNK], lo[kEval2MaxNK];
  uint32_t hmax = 0u, hmin = 0xffffffffu;
#pragma unroll
  for (int i = 0; i < kEval2MaxNK; ++i) {
    const int idx = lane + 32 * i;
    const uint64_t key = idx < n ? buf[idx] : 0ull;
    hi[i] = static_cast<uint32_t>(key >> 32); lo[i] = static_cast<uint32_t>(key);
    if (hi[i]) { hmax = max(hmax, hi[i]); hmin = min(hmin, hi[i]); }
  }
  hist[lane] = 0;
  hmax = __reduce_max_sync(0xffffffffu, hmax);
  hmin = __reduce_min_sync(0xffffffffu, hmin);
  const float smax = key_score(static_cast<uint64_t>(hmax) << 32), smin = key_score(static_cast<uint64_t>(hmin) << 32);
  const float scale = 32.0f / (smax - smin);
  __syncwarp();
  int bin[kEval2MaxNK];
#pragma unroll
  for (int i = 0; i < kEval2MaxNK; ++i) {
    bin[i] = -1;
    if (hi[i]) {
      const float sc = key_score(static_cast<uint64_t>(hi[i]) << 32);
      bin[i] = min(31, max(0, static_cast<int>((sc - smin) * scale)));
      atomicAdd(&hist[bin[i]], 1);
    }
  }
  __syncwarp();
  int suf = hist[lane];
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int t = __shfl_down_sync(0xffffffffu, suf, d);
    if (lane + d < 32) suf += t;
  }
  const unsigned okm = __ballot_sync(0xffffffffu, suf >= k);      // suf is non-increasing in the bin index; bin 0 holds all n >= k keys
  const int B = 31 - __clz(okm | 1u);
  const int kept = __shfl_sync(0xffffffffu, suf, B);
  if (!(smax > smin) || kept > max_keep || kept < k) {             // bins cannot separate (equal scores): exact select
    __syncwarp();
    *n_out = k;
    return compact_row(buf, n, k, lane);
  }
  uint32_t kmin = 0xffffffffu;
#pragma unroll
  for (int i = 0; i < kEval2MaxNK; ++i) if (bin[i] >= B) kmin = min(kmin, hi[i]);
  kmin = __reduce_min_sync(0xffffffffu, kmin);
  int base = 0;
  const unsigned lt = (1u << lane) - 1u;
#pragma unroll
  for (int i = 0; i < kEval2MaxNK; ++i) {
    const bool keep = bin[i] >= B;
    const unsigned mk = __ballot_sync(0xffffffffu, keep);
    if (keep) buf[base + __popc(mk & lt)] = (static_cast<uint64_t>(hi[i]) << 32) | lo[i];
    base += __popc(mk);
  }
  __syncwarp();
  *n_out = kept;
  return key_score(static_cast<uint64_t>(kmin) << 32);
}

// out-of-line copies for the selector warps of the third generation (few live registers around the call; the
// second generation inlines: a call there spills the 128 score registers)
__device__ __noinline__ float compact_row_call(uint64_t* buf, int n, int k, int lane) { return compact_row(buf, n, k, lane); }
__device__ __noinline__ float compact_row_coarse_call(uint64_t* buf, int n, int k, int max_keep, int lane, int* hist, int* n_out) {
  return compact_row_coarse(buf, n, k, max_keep, lane, hist, n_out);
}

#ifdef NNCF_EVAL_STATS_BUILD
struct FStat { unsigned tiles_hit, hit_rows, appended, compactions; };
#define FSTAT_ADD(field, x) do { fstat.field += (x); } while (0)
#else
struct FStat {};
#define FSTAT_ADD(field, x) do { } while (0)
#endif

struct Row2 {
  float thr;     // a lower bound of the row's k-th best score (exact as of the last compaction; -inf before k keys exist)
  int cnt;       // keys in the row's buffer
};

// (A compact variant of this filter - one run-time loop over the row's passing chunks, compactions out of line, 38 KB of
// SASS instead of 160 KB - was SLOWER (17.8 vs 15.7 ms at k = 50): the cost is the dependent chain per hit row, not
// instruction fetch.  The third generation takes the candidate path off the reading warps instead.)
__device__ __forceinline__ void filter_tile2(const float (&v)[4][32], uint32_t col0, uint32_t col_end, Row2& st, bool row_ok,
                                             uint64_t* bufs_warp, int CAP, int k, float* stage, int lane, unsigned own_mask,
                                             bool skip_hits, bool exact_compaction, FStat& fstat) {
  float gm[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    float m8[4];
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const float* w = v[c] + g * 8;
      m8[g] = fmaxf(fmaxf(fmaxf(w[0], w[1]), fmaxf(w[2], w[3])), fmaxf(fmaxf(w[4], w[5]), fmaxf(w[6], w[7])));
    }
    gm[c] = fmaxf(fmaxf(m8[0], m8[1]), fmaxf(m8[2], m8[3]));
  }
  const float m = fmaxf(fmaxf(gm[0], gm[1]), fmaxf(gm[2], gm[3]));
  unsigned hits = __ballot_sync(0xffffffffu, row_ok && m >= st.thr) & own_mask;
  if (skip_hits) { if (hits == 0x12345678u) st.cnt = 1; st.thr = fmaxf(st.thr, m - 1.0f); return; }   // developer switch
  const unsigned lt = (1u << lane) - 1u;
  FSTAT_ADD(tiles_hit, hits ? 1 : 0); FSTAT_ADD(hit_rows, __popc(hits));
  while (hits) {
    const int L = __ffs(hits) - 1;
    hits &= hits - 1;
    float thr_l = __shfl_sync(0xffffffffu, st.thr, L);
    int cnt_l = __shfl_sync(0xffffffffu, st.cnt, L);
    uint64_t* buf = bufs_warp + (size_t)L * (CAP + 1);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const float gmc = __shfl_sync(0xffffffffu, gm[c], L);
      if (gmc >= thr_l && col0 + c * 32 < col_end) {    // warp-uniform
        if (lane == L) {
#pragma unroll
          for (int t = 0; t < 32; t += 4)
            *reinterpret_cast<float4*>(stage + t) = make_float4(v[c][t], v[c][t + 1], v[c][t + 2], v[c][t + 3]);
        }
        __syncwarp();
        const float x = stage[lane];
        const uint32_t col = col0 + c * 32 + lane;
        const bool cand = (col < col_end) && (x >= thr_l);
        const unsigned cm = __ballot_sync(0xffffffffu, cand);
        if (cand) buf[cnt_l + __popc(cm & lt)] = make_key(x, col);
        cnt_l += __popc(cm);
        FSTAT_ADD(appended, __popc(cm));
        __syncwarp();
        if (cnt_l > CAP - 32) {                         // fewer than 32 free slots: keep (about) the k best, raise the threshold
          int kept = k;
          FSTAT_ADD(compactions, 1);
          if (exact_compaction) thr_l = compact_row(buf, cnt_l, k, lane);
          else thr_l = compact_row_coarse(buf, cnt_l, k, CAP - 48, lane, reinterpret_cast<int*>(stage), &kept);
          cnt_l = kept;
        }
      }
    }
    if (lane == L) { st.thr = thr_l; st.cnt = cnt_l; }
  }
}

template <int NSUB, int CL>
__global__ void __launch_bounds__(kEval2Threads, 1)
eval_topk_tc2_kernel(EvalArgs a) {
  constexpr int DP = 64 * NSUB;
  constexpr int kColU = kEval2Acc * 128;            // TMEM: 3 accumulators, then the user block (DP / 2 columns)
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // (pointer arithmetic keeps the shared address space: LDS / STS, not generic LD / ST)
  const int nst = a.nstages, CAP = a.cap;
  uint8_t* sV = smem;
  uint64_t* bufs_sm = reinterpret_cast<uint64_t*>(sV + nst * NSUB * kSubBytes);    // [128][CAP + 1] (padded rows)
  float* stage_all = reinterpret_cast<float*>(bufs_sm + 128 * (CAP + 1));          // [filter warps][128]: a hit row's passing chunks / the compaction histogram
  uint64_t* bars = reinterpret_cast<uint64_t*>(stage_all + kEval2EpiWarps * 32);
  uint64_t* u_full = bars;
  uint64_t* v_full = bars + 1;    // [4]
  uint64_t* v_empty = bars + 5;   // [4]
  uint64_t* s_full = bars + 9;    // [3]
  uint64_t* s_empty = bars + 13;  // [3]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 17);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ub = blockIdx.x, sp = blockIdx.y;
  const int n_tiles = static_cast<int>((a.n_items + 127) >> 7);
  const int t0 = sp * a.tiles_per_split;
  const int t1 = min(n_tiles, t0 + a.tiles_per_split);
  const int nj = max(0, t1 - t0);
  const int rot = nj > 0 ? static_cast<int>((static_cast<unsigned>(ub / CL) * 37u + static_cast<unsigned>(sp) * 11u) % static_cast<unsigned>(nj)) : 0;
  const uint32_t crank = CL > 1 ? cluster_ctarank() : 0u;
  constexpr uint16_t kMask = static_cast<uint16_t>((1u << CL) - 1u);

  if (tid == 0) {
    mbar_init(u_full, 4);
    for (int s = 0; s < 4; ++s) { mbar_init(&v_full[s], 1); mbar_init(&v_empty[s], CL); }
    for (int s = 0; s < kEval2Acc; ++s) { mbar_init(&s_full[s], 1); mbar_init(&s_empty[s], kEval2EpiWarps); }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();     // peers' barriers are initialised before anybody multicasts into this CTA
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    if (lane == 0 && nj > 0) {
      for (int j = 0; j < nj; ++j) {
        const int st = j % nst;
        mbar_wait(&v_empty[st], ((j / nst) & 1) ^ 1);
        if ((a.dbg_pipe & 1) && j >= nst) { mbar_arrive(&v_full[st]); continue; }   // developer switch: stale stage contents
        mbar_expect_tx(&v_full[st], NSUB * kSubBytes);
        const int jj = (j + rot) % nj;
        const uint8_t* gV = a.Vimg + (size_t)(t0 + jj) * NSUB * kSubBytes;
        constexpr uint32_t kSlice = NSUB * kSubBytes / CL;        // my share of the tile, delivered to every CTA
        if (a.prefetch_ahead > 0 && j + a.prefetch_ahead < nj) {
          // the copy is latency-bound (3 stages cover ~1.5k cycles, a DRAM miss takes longer): pull my share of a tile
          // further ahead into L2 now
          const uint8_t* gP = a.Vimg + (size_t)(t0 + (j + a.prefetch_ahead + rot) % nj) * NSUB * kSubBytes + (size_t)crank * kSlice;
          asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gP), "r"(kSlice) : "memory");
        }
        if (CL > 1) {
          bulk_g2s_multicast(sV + st * NSUB * kSubBytes + crank * kSlice, gV + (size_t)crank * kSlice, kSlice, &v_full[st], kMask);
        } else {
          for (int s = 0; s < NSUB; ++s)
            bulk_g2s(sV + (st * NSUB + s) * kSubBytes, gV + (size_t)s * kSubBytes, kSubBytes, &v_full[st]);
        }
      }
    }
  } else if (warp == 1) {
    if (nj > 0) {
      // The whole warp runs this loop (warp-uniform control flow and operands -> uniform registers); one elected lane issues.
      // An `if (lane == 0)` body made ptxas rebuild the descriptor and move the tensor-memory addresses through an
      // ELECT / R2UR.BROADCAST / BRA.U.ANY sequence before every MMA: 163 cycles per instruction against 64 of tensor work.
      const uint32_t idesc = make_idesc_bf16(128, 128, 0, 0);
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
      const uint64_t desc0 = make_smem_desc(smem_u32(sV), 16, 1024);
      mbar_wait(u_full, 0);                                       // the user block sits in TMEM
      tc_fence_after();
      for (int j = 0; j < nj; ++j) {
        const int st = j % nst, sb = j % kEval2Acc;
        mbar_wait(&v_full[st], (j / nst) & 1);
        mbar_wait(&s_empty[sb], ((j / kEval2Acc) & 1) ^ 1);
        tc_fence_after();
        if (elect_one()) {
          if (!(a.dbg_pipe & 2)) {
            const uint64_t dst = desc0 + static_cast<uint64_t>((st * NSUB * kSubBytes) >> 4);
            const uint32_t dcol = tmem_u + sb * 128;
#pragma unroll
            for (int k = 0; k < DP / 16; ++k)                     // K = 16 per MMA = 8 TMEM columns of the user block
              umma_bf16_ts(dcol, tmem_u + kColU + 8 * k, dst + static_cast<uint64_t>(((k >> 2) * kSubBytes + (k & 3) * 32) >> 4),
                           idesc, k > 0);
          }
          umma_commit(&s_full[sb]);
          if (CL > 1) umma_commit_multicast(&v_empty[st], kMask);   // the stage is rewritten by all CL producers
          else umma_commit(&v_empty[st]);
        }
        __syncwarp();
      }
    }
  } else {
    const int q = warp & 3;
    const int il = q * 32 + lane;
    const int64_t urow = (int64_t)ub * 128 + il;
    const bool row_ok = urow < a.n_users;
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    const int rg = (warp - 2) >> 2;                                // row group inside the quadrant
    constexpr int kOwn = 32 / kEval2RowGroups;
    const unsigned own_mask = (kOwn == 32 ? 0xffffffffu : ((1u << kOwn) - 1u)) << (rg * kOwn);
    // my user's row -> bf16 pairs -> TMEM columns kColU .. (row = lane): the A operand of every MMA of this CTA
    if (rg == 0) {
      const float* u = a.Uf + (row_ok ? urow : 0) * a.d;
#pragma unroll 1
      for (int c0 = 0; c0 < DP; c0 += 32) {
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int c = c0 + 2 * i;
          const float x0 = (row_ok && c < a.d) ? __ldg(u + c) : 0.0f;
          const float x1 = (row_ok && c + 1 < a.d) ? __ldg(u + c + 1) : 0.0f;
          pk[i] = pack_bf16x2(x0, x1);
        }
        tmem_st16(tmem + lane_addr + kColU + c0 / 2, pk);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(u_full);
    }
    Row2 rs; rs.thr = -INFINITY; rs.cnt = 0;
    FStat fstat{};
    long long cyc_hit = 0; (void)cyc_hit;
    uint64_t* bufs_warp = bufs_sm + (size_t)q * 32 * (CAP + 1);
    float* stage = stage_all + (warp - 2) * 32;
    const uint32_t col_end = static_cast<uint32_t>(a.n_items);
    for (int j = 0; j < nj; ++j) {
      const int sb = j % kEval2Acc;
      mbar_wait(&s_full[sb], (j / kEval2Acc) & 1);
      tc_fence_after();
      float v[4][32];
      if (a.dbg_mode != 2) {
#pragma unroll
        for (int c = 0; c < 4; ++c) tmem_ld32(tmem + lane_addr + sb * 128 + c * 32, v[c]);
        tmem_ld_wait();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_empty[sb]);
      if (a.dbg_mode == 1 || a.dbg_mode == 2) { if (a.dbg_mode < 2 && v[0][0] == 123456.0f) rs.cnt = 1; continue; }
      const uint32_t col0 = static_cast<uint32_t>((t0 + (j + rot) % nj) * 128);
      filter_tile2(v, col0, col_end, rs, row_ok, bufs_warp, CAP, a.k, stage, lane, own_mask, a.dbg_mode == 3, (a.dbg_pipe & 4) != 0, fstat);
    }
    __syncwarp();
    // final compaction of every row to its k best, then the row's KP slots go to the global per-(user, split) lists
    const int64_t gstride = (int64_t)a.nsplit * a.KP;
    uint64_t* grow0 = a.lists + ((int64_t)ub * 128 + q * 32) * gstride + (int64_t)sp * a.KP;
    for (int r = rg * kOwn; r < (rg + 1) * kOwn; ++r) {
      int n = __shfl_sync(0xffffffffu, rs.cnt, r);
      uint64_t* buf = bufs_warp + (size_t)r * (CAP + 1);
      if (n > a.k) { compact_row(buf, n, a.k, lane); n = a.k; }
      for (int i = lane; i < a.KP; i += 32) grow0[r * gstride + i] = i < n ? buf[i] : 0ull;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();     // nobody exits while a peer may still multicast into it or signal its barriers
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// ------------------------------------------------------------------------------------------------
// Third generation: a CTA PAIR (cluster of 2 = the two SMs of a TPC) on one tcgen05.mma.cta_group::2.
//   M = 256 users (128 per CTA, each CTA's block in its OWN tensor memory as the A operand), N = 128 items of which
//   each CTA stages only 64 rows (the first / second 8 KiB of every 16 KiB sub-tile of the item image): an item tile
//   enters each SM once per 256 users, so the L2 -> SM stream per SM is half of the second generation's, and a stage
//   is half as big, so the same shared memory holds twice as many tiles in flight.
//   Leader (cluster rank 0): warp 1 lane 0 issues the MMAs and commits with a multicast arrive to both CTAs' barriers.
//   Peer: warp 1 lane 0 relays "my half of stage st has landed" to the leader's p_full[st] (remote mbarrier arrive);
//   the epilogue warps of both CTAs release accumulators on the leader's s_empty (remote arrive).
//   Epilogue (filter) identical to the second generation.
// ------------------------------------------------------------------------------------------------
constexpr int kEval3MaxStages = 8;
// Warp roles of the third generation: 0 producer, 1 MMA issuer / relay, 2..9 readers (two per TMEM lane quadrant: columns
// 0..63 and 64..127 of every accumulator, so nothing is read twice), 10..17 selectors (two per quadrant: rows 0..15 and
// 16..31 of the quadrant; one ring per (reader, row half)).
// A reader loads its 32 x 64 scores, releases the accumulator, takes the row / chunk maxima, and - for rows that reach their
// threshold - only STAGES the passing 32-column chunks into its shared-memory ring; it never touches the k-best sets.  The
// selector of a row half owns its row buffers and thresholds: it drains its two rings (two entries at a time, two
// independent dependency chains: compare, ballot, append, compaction) and publishes each raised threshold for the
// readers' next tiles (a stale threshold only forwards extra chunks: the result stays exact).  Why: ncu pc-sampling of the one-warp filter (profiles/r01c) showed the four filter
// warps busy 83 % of the time, 3/4 of it in the candidate path, every sample a fixed-latency dependency stall - one warp
// per scheduler, nothing to hide a ~700-cycle chain per hit row behind - while an accumulator is recycled only when all
// 8 filter warps of the pair have read it.  Variants measured first (37,888 users x 1M items, k = 50; 15.7 ms before):
// two readers per quadrant with 16 rows each 18.6 ms (duplicated TMEM reads), per-lane compare-and-branch 30 ms,
// per-lane branch-free scan down the maxima hierarchy 19.9 ms; one reader + one selector per quadrant 14.3 ms (the
// reader alone: wait 310, TMEM load 210, 128 maxima 430, staging 700 cycles per tile); two readers + one selector
// 12.4 ms (selector 100 % busy, ~150 dependent instructions per pair of entries); two readers + two selectors 12.1 ms
// (this version; readers 80 % busy: TMEM load 215, 64 maxima 260, threshold 125, staging + publishing 345 cycles per tile);
// four readers (32 columns each) + two selectors on multi-producer rings (slots reserved with a shared-memory atomic,
// made visible by a lap tag) 12.8 ms: the fixed cost per reader and tile (barrier wait, fences, arrive; the pipeline
// alone went from 959 to 1,010 cycles per tile with 26 warps) outweighs the smaller chunk.
// Also slower: a reader that keeps one TMEM load in flight behind the processing of the other chunk (13.2 ms): it holds
// the accumulator one chunk longer, and the accumulator ring (3 deep) is what the MMA waits for.
constexpr int kEval3Ring = 16;                    // staged chunks per ring (power of two, >= 16 = one chunk of every row of a row half)
constexpr int kEval3Threads = 64 + 32 * 16;
constexpr int kEvalDefaultGen = 3;                // NNCF_EVAL_GEN overrides (2 = second generation, 3 = CTA pair; the plan falls back to 2 when
                                                  // the pair's rings and row buffers leave fewer than 2 item stages, e.g. k > 64, or for a single user block)

__host__ __device__ inline size_t eval3_smem_bytes(int nsub, int nstages, int cap) {
  return (size_t)nsub * (kSubBytes / 2) * nstages + (size_t)128 * (cap + 1) * 8 + 16 * kEval3Ring * (128 + 8) /*rings + meta*/ +
         8 * 128 /*histograms*/ + 128 * 4 /*thresholds*/ + 256 /*ring counters*/ + 1024 + 512;
}

template <int NSUB>
__global__ void __launch_bounds__(kEval3Threads, 1)
eval_topk_tc3_kernel(EvalArgs a) {
  constexpr int DP = 64 * NSUB;
  constexpr int kHalf = kSubBytes / 2;              // 64 item rows x 64 bf16 of one sub-tile
  constexpr int kStage = NSUB * kHalf;
  constexpr int kColU = kEval2Acc * 128;            // TMEM: 3 accumulators, then the user block (DP / 2 columns)
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // (pointer arithmetic keeps the shared address space: LDS / STS, not generic LD / ST)
  const int nst = a.nstages, CAP = a.cap;
  uint8_t* sV = smem;
  uint64_t* bufs_sm = reinterpret_cast<uint64_t*>(sV + (size_t)nst * kStage);      // [128][CAP + 1] (padded rows)
  float* ring = reinterpret_cast<float*>(bufs_sm + 128 * (CAP + 1));               // [16 = (quadrant, column half, row half)][kEval3Ring][32] staged chunks
  uint2* ring_meta = reinterpret_cast<uint2*>(ring + 16 * kEval3Ring * 32);        // [16][kEval3Ring] (row inside the quadrant, first column)
  int* hist_all = reinterpret_cast<int*>(ring_meta + 16 * kEval3Ring);             // [8 selectors][32] compaction histograms
  float* thr_sm = reinterpret_cast<float*>(hist_all + 8 * 32);                     // [128] thresholds published by the selectors
  volatile int* ring_ctl = reinterpret_cast<volatile int*>(thr_sm + 128);          // [16] head, [16] tail, [16] finished
  uint64_t* bars = reinterpret_cast<uint64_t*>(const_cast<int*>(ring_ctl) + 64);
  uint64_t* u_full = bars;                          // leader's: both user blocks are in tensor memory (8 warp arrivals)
  uint64_t* v_full = bars + 1;                      // [8] my half of the stage has landed
  uint64_t* p_full = bars + 9;                      // [8] leader's: the peer's half has landed (relayed)
  uint64_t* v_empty = bars + 17;                    // [8] the MMAs reading the stage have completed (multicast commit)
  uint64_t* s_full = bars + 25;                     // [3] accumulator ready (multicast commit)
  uint64_t* s_empty = bars + 28;                    // [3] leader's: accumulator drained by the 8 epilogue warps of the pair
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 31);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ub = blockIdx.x, sp = blockIdx.y;
  const int n_tiles = static_cast<int>((a.n_items + 127) >> 7);
  const int t0 = sp * a.tiles_per_split;
  const int t1 = min(n_tiles, t0 + a.tiles_per_split);
  const int nj = max(0, t1 - t0);
  const int rot = nj > 0 ? static_cast<int>((static_cast<unsigned>(ub >> 1) * 37u + static_cast<unsigned>(sp) * 11u) % static_cast<unsigned>(nj)) : 0;
  const uint32_t crank = cluster_ctarank();
  const bool leader = crank == 0;

  if (tid == 0) {
    mbar_init(u_full, 2 * 4);
    for (int s = 0; s < kEval3MaxStages; ++s) { mbar_init(&v_full[s], 1); mbar_init(&p_full[s], 1); mbar_init(&v_empty[s], 1); }
    for (int s = 0; s < kEval2Acc; ++s) { mbar_init(&s_full[s], 1); mbar_init(&s_empty[s], 2 * 8); }
    mbar_fence_init();
  }
  if (tid < 128) thr_sm[tid] = -INFINITY;
  if (tid < 64) ring_ctl[tid] = 0;
  if (warp == 1) tmem_alloc_pair(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                 // both CTAs' barriers and tensor memory exist before anything crosses the pair
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    if (lane == 0 && nj > 0) {
      for (int j = 0; j < nj; ++j) {
        const int st = j % nst;
        mbar_wait(&v_empty[st], ((j / nst) & 1) ^ 1);
        if ((a.dbg_pipe & 1) && j >= nst) { mbar_arrive(&v_full[st]); continue; }   // developer switch: stale stage contents
        mbar_expect_tx(&v_full[st], kStage);
        const uint8_t* gV = a.Vimg + (size_t)(t0 + (j + rot) % nj) * NSUB * kSubBytes + (size_t)crank * kHalf;
#pragma unroll
        for (int s = 0; s < NSUB; ++s)
          bulk_g2s(sV + (size_t)st * kStage + s * kHalf, gV + (size_t)s * kSubBytes, kHalf, &v_full[st]);
      }
    }
  } else if (warp == 1) {
    if (nj > 0) {
      if (leader) {
        // warp-uniform loop, one elected lane issues (see the second generation)
        const uint32_t idesc = make_idesc_bf16(256, 128, 0, 0);
        const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
        const uint64_t desc0 = make_smem_desc(smem_u32(sV), 16, 1024);
        mbar_wait(u_full, 0);                                       // both user blocks sit in tensor memory
        tc_fence_after();
        for (int j = 0; j < nj; ++j) {
          const int st = j % nst, sb = j % kEval2Acc;
          mbar_wait(&v_full[st], (j / nst) & 1);
          mbar_wait(&p_full[st], (j / nst) & 1);
          mbar_wait(&s_empty[sb], ((j / kEval2Acc) & 1) ^ 1);
          tc_fence_after();
          if (elect_one()) {
            if (!(a.dbg_pipe & 2)) {
              const uint64_t dst = desc0 + static_cast<uint64_t>((st * kStage) >> 4);
              const uint32_t dcol = tmem_u + sb * 128;
#pragma unroll
              for (int k = 0; k < DP / 16; ++k)                     // K = 16 per MMA = 8 TMEM columns of the user blocks
                umma_bf16_ts_pair(dcol, tmem_u + kColU + 8 * k, dst + static_cast<uint64_t>(((k >> 2) * kHalf + (k & 3) * 32) >> 4),
                                  idesc, k > 0);
            }
            umma_commit_pair(&s_full[sb], 3);
            umma_commit_pair(&v_empty[st], 3);
          }
          __syncwarp();
        }
      } else if (lane == 0) {
        const uint32_t p_full0 = mapa_shared(smem_u32(p_full), 0);
        for (int j = 0; j < nj; ++j) {
          const int st = j % nst;
          mbar_wait(&v_full[st], (j / nst) & 1);
          mbar_arrive_cluster(p_full0 + 8u * st);
        }
      }
    }
  } else if (warp < 10) {
    // ------------------------------------------------------------------------------ readers (two per TMEM lane quadrant)
    const int q = warp & 3;
    const int h = (warp - 2) >> 2;                        // my column half of every accumulator
    const int rid = q * 2 + h;                            // my ring
    const int il = q * 32 + lane;
    const int64_t urow = (int64_t)ub * 128 + il;
    const bool row_ok = urow < a.n_users;
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    const uint32_t s_empty0 = mapa_shared(smem_u32(s_empty), 0);
    // my user's row -> bf16 pairs -> TMEM columns kColU .. (row = lane): my CTA's half of the A operand of every MMA
    if (h == 0) {
      const uint32_t u_full0 = mapa_shared(smem_u32(u_full), 0);
      const float* u = a.Uf + (row_ok ? urow : 0) * a.d;
#pragma unroll 1
      for (int c0 = 0; c0 < DP; c0 += 32) {
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int c = c0 + 2 * i;
          const float x0 = (row_ok && c < a.d) ? __ldg(u + c) : 0.0f;
          const float x1 = (row_ok && c + 1 < a.d) ? __ldg(u + c + 1) : 0.0f;
          pk[i] = pack_bf16x2(x0, x1);
        }
        tmem_st16(tmem + lane_addr + kColU + c0 / 2, pk);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(u_full0);
    }
    float* my_ring = ring + rid * 2 * kEval3Ring * 32;    // my two rings: row halves 0 and 1
    uint2* my_meta = ring_meta + rid * 2 * kEval3Ring;
    volatile float* my_thr = thr_sm + il;
    volatile int* head_sm = ring_ctl + rid * 2;
    volatile int* tail_sm = ring_ctl + 16 + rid * 2;
    int head0 = 0, head1 = 0, pub0 = 0, pub1 = 0, tail_seen0 = 0, tail_seen1 = 0;
    const unsigned lt = (1u << lane) - 1u;
    const int rh = lane >> 4;                             // my row's half of the quadrant
    const uint32_t col_end = static_cast<uint32_t>(a.n_items);
    auto publish = [&]() {                                // make the staged chunks visible, then the new heads
      if (head0 != pub0 || head1 != pub1) {
        __syncwarp();                                     // (orders the lanes' stores before lane 0's release)
        if (lane == 0) { if (head0 != pub0) st_release_cta_smem(head_sm, head0); if (head1 != pub1) st_release_cta_smem(head_sm + 1, head1); }
        pub0 = head0; pub1 = head1;
      }
    };
    for (int j = 0; j < nj; ++j) {
      const int sb = j % kEval2Acc;
      const float thr = *my_thr;                          // possibly stale (lower): forwards extra chunks, never loses one;
                                                          // read before the wait so that its latency hides behind it
      mbar_wait(&s_full[sb], (j / kEval2Acc) & 1);
      tc_fence_after();
      float v[2][32];
      if (a.dbg_mode != 2) {
#pragma unroll
        for (int c = 0; c < 2; ++c) tmem_ld32(tmem + lane_addr + sb * 128 + h * 64 + c * 32, v[c]);
        tmem_ld_wait();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(s_empty0 + 8u * sb);
      if (a.dbg_mode == 1 || a.dbg_mode == 2) { if (a.dbg_mode < 2 && v[0][0] == 123456.0f) head0 = 1; continue; }
      const uint32_t col0 = static_cast<uint32_t>((t0 + (j + rot) % nj) * 128 + h * 64);
      float gm[2];
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        float m8[4];
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const float* w = v[c] + g * 8;
          m8[g] = fmaxf(fmaxf(fmaxf(w[0], w[1]), fmaxf(w[2], w[3])), fmaxf(fmaxf(w[4], w[5]), fmaxf(w[6], w[7])));
        }
        gm[c] = fmaxf(fmaxf(m8[0], m8[1]), fmaxf(m8[2], m8[3]));
      }
      const float m = fmaxf(gm[0], gm[1]);
      const bool hit = row_ok && m >= thr;
      if (!__any_sync(0xffffffffu, hit) || a.dbg_mode == 3) continue;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const bool pc = hit && gm[c] >= thr && col0 + c * 32 < col_end;
        const unsigned bc = __ballot_sync(0xffffffffu, pc);
        if (!bc) continue;
        const int n0 = __popc(bc & 0x0000ffffu), n1 = __popc(bc & 0xffff0000u);
        if (head0 + n0 - tail_seen0 > kEval3Ring || head1 + n1 - tail_seen1 > kEval3Ring) {   // a ring is full: wait for its selector
          publish();                                      // (it can only drain what it can see)
          uint32_t spins = 0;
          for (;;) {
            tail_seen0 = ld_acquire_cta_smem(tail_sm); tail_seen1 = ld_acquire_cta_smem(tail_sm + 1);
            if (head0 + n0 - tail_seen0 <= kEval3Ring && head1 + n1 - tail_seen1 <= kEval3Ring) break;
            __nanosleep(40);                              // (a hot spin would take issue slots from the selectors on this scheduler)
            if (++spins > 20000000u) __trap();
          }
        }
        if (pc) {
          const int pos = rh ? head1 + __popc(bc & 0xffff0000u & lt) : head0 + __popc(bc & 0x0000ffffu & lt);
          const int slot = rh * kEval3Ring + (pos & (kEval3Ring - 1));
          float* dst = my_ring + slot * 32;
#pragma unroll
          for (int t = 0; t < 32; t += 4)
            *reinterpret_cast<float4*>(dst + t) = make_float4(v[c][t], v[c][t + 1], v[c][t + 2], v[c][t + 3]);
          my_meta[slot] = make_uint2(static_cast<uint32_t>(lane), col0 + c * 32);
        }
        head0 += n0; head1 += n1;
      }
      publish();
    }
    __syncwarp();
    if (lane == 0) { st_release_cta_smem(ring_ctl + 32 + rid * 2, 1); st_release_cta_smem(ring_ctl + 32 + rid * 2 + 1, 1); }   // finished: the selectors drain what is left and write the lists
  } else {
    // ------------------------------------------------------------------------------ selectors (own the k-best sets)
    const int q = warp & 3;
    const int rh = (warp - 10) >> 2;                      // my row half: rows 16 rh .. 16 rh + 15 of the quadrant
    const int ra = ((q * 2 + 0) * 2 + rh), rb = ((q * 2 + 1) * 2 + rh);   // my rings: column halves 0 and 1
    const float* ring_a = ring + ra * kEval3Ring * 32;
    const float* ring_b = ring + rb * kEval3Ring * 32;
    const uint2* meta_a = ring_meta + ra * kEval3Ring;
    const uint2* meta_b = ring_meta + rb * kEval3Ring;
    volatile int* ctl = ring_ctl;                         // [r] head, [16 + r] tail, [32 + r] finished
    int* hist = hist_all + (warp - 10) * 32;
    uint64_t* bufs_warp = bufs_sm + q * 32 * (CAP + 1);
    const unsigned lt = (1u << lane) - 1u;
    const uint32_t col_end = static_cast<uint32_t>(a.n_items);
    const bool exact_compaction = (a.dbg_pipe & 4) != 0;
    const int k = a.k;
    float thr = -INFINITY;                                // lane = row inside the quadrant
    int cnt = 0, tail_a = 0, tail_b = 0;
    uint32_t idle = 0;
    // fewer than 32 free slots in row L: keep (about) the k best, raise and publish the threshold
    auto make_room = [&](int L, uint64_t* buf, int& cnt_l) {
      if (cnt_l > CAP - 32) {
        __syncwarp();
        int kept = k;
        const float t = exact_compaction ? compact_row_call(buf, cnt_l, k, lane) : compact_row_coarse_call(buf, cnt_l, k, CAP - 48, lane, hist, &kept);
        cnt_l = kept;
        if (lane == L) { thr = t; thr_sm[q * 32 + L] = t; }
      }
    };
    auto do_one = [&](const uint2 me, const float x) {
      const int L = static_cast<int>(me.x);
      const float thr_l = __shfl_sync(0xffffffffu, thr, L);
      const uint32_t col = me.y + lane;
      const bool cand = (col < col_end) && (x >= thr_l);
      const unsigned cm = __ballot_sync(0xffffffffu, cand);
      if (!cm) return;
      int cnt_l = __shfl_sync(0xffffffffu, cnt, L);
      uint64_t* buf = bufs_warp + L * (CAP + 1);
      if (cand) buf[cnt_l + __popc(cm & lt)] = make_key(x, col);
      cnt_l += __popc(cm);
      make_room(L, buf, cnt_l);
      if (lane == L) cnt = cnt_l;
    };
    for (;;) {
      const int fin = ld_acquire_cta_smem(ctl + 32 + ra) & ld_acquire_cta_smem(ctl + 32 + rb);                  // (read before the heads: what was published before `finished` is seen)
      const int head_a = ld_acquire_cta_smem(ctl + ra), head_b = ld_acquire_cta_smem(ctl + rb);
      if (head_a == tail_a && head_b == tail_b) {
        if (fin) break;
        if (++idle > 40000000u) __trap();
        __nanosleep(20);
        continue;
      }
      idle = 0;
      // two entries per step while both rings have some: two independent dependency chains
#pragma unroll 1
      while (tail_a != head_a && tail_b != head_b) {
        const int sa = tail_a & (kEval3Ring - 1), sb2 = tail_b & (kEval3Ring - 1);
        ++tail_a; ++tail_b;
        const uint2 ma = meta_a[sa], mb = meta_b[sb2];
        const float xa = ring_a[sa * 32 + lane], xb = ring_b[sb2 * 32 + lane];
        const int La = static_cast<int>(ma.x), Lb = static_cast<int>(mb.x);
        if (La == Lb) { do_one(ma, xa); do_one(mb, xb); continue; }     // same row: one buffer, in sequence
        const float thr_a = __shfl_sync(0xffffffffu, thr, La), thr_b = __shfl_sync(0xffffffffu, thr, Lb);
        const uint32_t col_a = ma.y + lane, col_b = mb.y + lane;
        const bool cand_a = (col_a < col_end) && (xa >= thr_a), cand_b = (col_b < col_end) && (xb >= thr_b);
        const unsigned cm_a = __ballot_sync(0xffffffffu, cand_a), cm_b = __ballot_sync(0xffffffffu, cand_b);
        if (!(cm_a | cm_b)) continue;
        int cnt_a = __shfl_sync(0xffffffffu, cnt, La), cnt_b = __shfl_sync(0xffffffffu, cnt, Lb);
        uint64_t* buf_a = bufs_warp + La * (CAP + 1);
        uint64_t* buf_b = bufs_warp + Lb * (CAP + 1);
        if (cand_a) buf_a[cnt_a + __popc(cm_a & lt)] = make_key(xa, col_a);
        if (cand_b) buf_b[cnt_b + __popc(cm_b & lt)] = make_key(xb, col_b);
        cnt_a += __popc(cm_a); cnt_b += __popc(cm_b);
        make_room(La, buf_a, cnt_a);
        make_room(Lb, buf_b, cnt_b);
        if (lane == La) cnt = cnt_a;
        if (lane == Lb) cnt = cnt_b;
      }
#pragma unroll 1
      for (; tail_a != head_a; ++tail_a) {
        const int sl = tail_a & (kEval3Ring - 1);
        do_one(meta_a[sl], ring_a[sl * 32 + lane]);
      }
#pragma unroll 1
      for (; tail_b != head_b; ++tail_b) {
        const int sl = tail_b & (kEval3Ring - 1);
        do_one(meta_b[sl], ring_b[sl * 32 + lane]);
      }
      __syncwarp();
      if (lane == 0) { st_release_cta_smem(ctl + 16 + ra, tail_a); st_release_cta_smem(ctl + 16 + rb, tail_b); }   // the slots up to here may be overwritten
    }
    __syncwarp();
    // final (exact) compaction of every row to its k best, then the row's KP slots go to the global per-(user, split) lists
    const int64_t gstride = (int64_t)a.nsplit * a.KP;
    uint64_t* grow0 = a.lists + ((int64_t)ub * 128 + q * 32) * gstride + (int64_t)sp * a.KP;
    for (int r = 16 * rh; r < 16 * rh + 16; ++r) {
      int n = __shfl_sync(0xffffffffu, cnt, r);
      uint64_t* buf = bufs_warp + r * (CAP + 1);
      if (n > a.k) { compact_row_call(buf, n, a.k, lane); n = a.k; }
      __syncwarp();
      for (int i = lane; i < a.KP; i += 32) grow0[r * gstride + i] = i < n ? buf[i] : 0ull;
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                 // the pair's MMAs read both shared memories and write both tensor memories: leave together
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair(tmem, 512);
  }
}

// ------------------------------------------------------------------------------------------------
// CUDA-core fp32 eval kernel (precision = fp32): 128 threads, thread = user row, 32 item columns at a time.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
eval_topk_simt_kernel(EvalArgs a) {
  extern __shared__ __align__(16) uint8_t smem_raw2[];
  const int dp = a.dp;
  float* Vs = reinterpret_cast<float*>(smem_raw2);                      // [32][dp]
  uint64_t* lists_sm = reinterpret_cast<uint64_t*>(Vs + 32 * dp);       // [128][KP]
  float* stage_all = reinterpret_cast<float*>(lists_sm + 128 * a.KP);   // [4 warps][32]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ub = blockIdx.x, sp = blockIdx.y;
  const int n_tiles = static_cast<int>((a.n_items + 127) >> 7);
  const int t0 = sp * a.tiles_per_split;
  const int t1 = min(n_tiles, t0 + a.tiles_per_split);
  for (int i = tid; i < 128 * a.KP; i += 128) lists_sm[i] = 0ull;
  const int il = tid, q = warp;
  const int64_t urow = (int64_t)ub * 128 + il;
  const bool row_ok = urow < a.n_users;
  RowState rs; rs.thr = -INFINITY; rs.thr_key = 0; rs.minpos = 0;
  uint64_t* lists_warp = lists_sm + (size_t)q * 32 * a.KP;
  float* stage = stage_all + q * 32;
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
    filter_chunk(acc, static_cast<uint32_t>(c0), col_end, rs, row_ok, lists_warp, a.KP, a.k, stage, lane);
  }
  __syncwarp();
  const int64_t gstride = (int64_t)a.nsplit * a.KP;
  store_lists(lists_warp, a.KP, a.lists + ((int64_t)ub * 128 + q * 32) * gstride + (int64_t)sp * a.KP, gstride, lane);
}

// sort + merge the per-split k-best sets of each user and decode: one warp per user
__global__ void __launch_bounds__(128)
topk_merge_kernel(const uint64_t* __restrict__ lists, int64_t n_users, int nsplit, int KP, int k, int sortn2,
                  int32_t* __restrict__ ids, float* __restrict__ scores) {
  __shared__ uint64_t scratch_all[4][256];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t u = (int64_t)blockIdx.x * 4 + warp;
  if (u >= n_users) return;
  uint64_t* sc = scratch_all[warp];
  const uint64_t* base = lists + u * nsplit * KP;
  for (int t = lane; t < sortn2; t += 32) sc[t] = (t < KP) ? base[t] : 0ull;
  __syncwarp();
  warp_bitonic_desc(sc, sortn2, lane);
  for (int s = 1; s < nsplit; ++s) {
    for (int t = KP + lane; t < sortn2; t += 32) sc[t] = (t - KP < KP) ? base[(int64_t)s * KP + (t - KP)] : 0ull;
    __syncwarp();
    warp_bitonic_desc(sc, sortn2, lane);
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
                  int64_t n_groups, int topk, float* __restrict__ out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t gidx = (int64_t)blockIdx.x * 4 + warp;
  if (gidx >= n_groups) return;
  const int64_t b = indptr[gidx], e = indptr[gidx + 1];
  const int64_t n = e - b;
  const int64_t k = topk >= 0 ? topk : n;          // eval_multiple_original: k = len(rec) when topk < 0 (metrics_ranking.py:45)
  double ap_acc = 0.0, rank_sum = 0.0;
  int64_t npos = 0, hits_k = 0;
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
    if (better < k) {                                // inside the top k of this user's list
      ap_acc += static_cast<double>(better_pos + 1) / static_cast<double>(better + 1);
      ++hits_k;
    }
    rank_sum += static_cast<double>(less) + 0.5 * static_cast<double>(equal + 1);   // average rank (1-based)
  }
  for (int o = 16; o > 0; o >>= 1) {
    ap_acc += __shfl_xor_sync(0xffffffffu, ap_acc, o);
    rank_sum += __shfl_xor_sync(0xffffffffu, rank_sum, o);
    npos += __shfl_xor_sync(0xffffffffu, npos, o);
    hits_k += __shfl_xor_sync(0xffffffffu, hits_k, o);
  }
  if (lane == 0) {
    const int64_t nneg = n - npos;
    const double denom = static_cast<double>(npos < k ? npos : k);          // min(nhits, k)
    out[gidx * 4 + 0] = (npos > 0 && k > 0) ? static_cast<float>(ap_acc / denom) : 0.0f;
    out[gidx * 4 + 1] = (npos > 0 && nneg > 0)
                            ? static_cast<float>((rank_sum - 0.5 * npos * (npos + 1)) / (static_cast<double>(npos) * nneg))
                            : NAN;
    out[gidx * 4 + 2] = npos > 0 ? static_cast<float>(static_cast<double>(hits_k) / static_cast<double>(npos)) : 0.0f;
    out[gidx * 4 + 3] = (npos > 0 && k > 0) ? static_cast<float>(static_cast<double>(hits_k) / static_cast<double>(k)) : 0.0f;
  }
}

}  // namespace nncf

using namespace nncf;

static int kpad_of(int k) { return (k + 31) / 32 * 32; }
static int pow2_ge(int x) { int p = 1; while (p < x) p <<= 1; return p; }

struct EvalPlan {
  int dp, nsub, KP, nsplit, nstages, tiles_per_split, n_tiles, cl;
  int v2, cap, nstages2;      // second-generation tensor-core kernel: usable, keys per row buffer, V stages
  int v3;                     // third generation (CTA pair, cta_group::2) selected: cap / nstages2 describe it, cl = 2
  int64_t n_ub_grid;
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
  p->n_ub = (n_users + 127) / 128;
  p->cl = (precision == NNCF_PREC_BF16 && p->n_ub >= kEvalCluster) ? kEvalCluster : 1;
  { const char* e = getenv("NNCF_EVAL_CL"); if (e && precision == NNCF_PREC_BF16) { const int v = atoi(e); if (v == 1 || v == 2 || v == 4 || v == 8) p->cl = v; } }   // developer override (8: second-generation kernel only)
  {
    const int padto = (precision == NNCF_PREC_BF16 && p->n_ub >= 2 && p->cl < 2) ? 2 : p->cl;   // a CTA pair needs an even block count
    p->n_ub_grid = (p->n_ub + p->cl - 1) / p->cl * p->cl;
    p->users_pad = (p->n_ub + padto - 1) / padto * padto * 128;
  }
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
  p->nstages = 0;
  p->v2 = 0; p->cap = 0; p->nstages2 = 0;
  if (precision == NNCF_PREC_BF16) {
    for (int st = 4; st >= 1; --st)
      if (eval_smem_bytes(p->nsub, st, p->KP) <= 232448) { p->nstages = st; break; }
    // second generation: the largest row buffer (fewest compactions) that still leaves 3 (else 2) item-tile stages
    const bool v1_env = [] { const char* e = getenv("NNCF_EVAL_V1"); return e && atoi(e) != 0; }();
    int st_hi = 3, cap_hi = 2 * p->KP > 192 ? 192 : (2 * p->KP < 96 ? 96 : 2 * p->KP);
    { const char* e = getenv("NNCF_EVAL_NST"); if (e) { const int v = atoi(e); if (v >= 2 && v <= 4) st_hi = v; } }     // developer overrides
    { const char* e = getenv("NNCF_EVAL_CAP"); if (e) { const int v = atoi(e); if (v % 32 == 0 && v >= p->KP + 32 && v <= 192) cap_hi = v; } }
    for (int st = st_hi; st >= 2 && !p->v2 && !v1_env; --st)
      for (int cap = cap_hi; cap >= p->KP + 32; cap -= 32)
        if (eval2_smem_bytes(p->nsub, st, cap) <= 232448) { p->v2 = 1; p->cap = cap; p->nstages2 = st; break; }
    if (p->v2) p->nstages = p->nstages2;
    // third generation (CTA pair): half-size stages; the largest row buffer that leaves >= 3 (else 2) stages in flight
    const int gen_env = [] { const char* e = getenv("NNCF_EVAL_GEN"); return e ? atoi(e) : kEvalDefaultGen; }();
    p->v3 = 0;
    if (p->v2 && gen_env >= 3 && p->n_ub >= 2) {
      int best_st = 0, best_cap = 0;
      for (int want = 3; want >= 2 && !best_st; --want)   // (measured: 4 stages 876, 2 stages 1,266 cycles per tile for the pipeline alone)
        for (int cap = cap_hi; cap >= p->KP + 32 && !best_st; cap -= 32) {
          int st = 0;
          for (int t = kEval3MaxStages; t >= want; --t) if (eval3_smem_bytes(p->nsub, t, cap) <= 232448) { st = t; break; }
          if (st) { best_st = st; best_cap = cap; }
        }
      // large k (k = 100 of the C4 configuration): the row buffers need not be a multiple of 32 keys - a row compacts when
      // fewer than 32 slots are free and keeps at most cap - 48 >= k keys - so the smallest buffer that leaves two item
      // stages keeps the pair kernel in play (k = 100, d = 128: 152 keys per row, 2 stages)
      for (int want = 3; want >= 2 && !best_st; --want)
        for (int cap = p->KP + 24; cap >= (topk + 48 + 7) / 8 * 8 && !best_st; cap -= 8) {
          int st = 0;
          for (int t = kEval3MaxStages; t >= want; --t) if (eval3_smem_bytes(p->nsub, t, cap) <= 232448) { st = t; break; }
          if (st) { best_st = st; best_cap = cap; }
        }
      { const char* e = getenv("NNCF_EVAL_NST3"); if (e && best_st) { const int v = atoi(e); if (v >= 2 && v <= best_st) best_st = v; } }   // developer override
      if (best_st) {
        p->v3 = 1; p->cap = best_cap; p->nstages2 = best_st; p->nstages = best_st; p->cl = 2;
        p->n_ub_grid = (p->n_ub + 1) / 2 * 2;
        // (the workspace offsets above were sized with the cluster size in force then; the pair needs users_pad even in blocks)
      }
    }
    if (p->nstages == 0) {
      set_error("eval (bf16): top-k sets of k > 64 do not fit next to dim > 128 operands; use k <= 64 or precision fp32");
      return NNCF_EUNSUPPORTED;
    }
  }
  return 0;
}

extern "C" size_t nncf_eval_topk_workspace_bytes(int64_t n_users, int64_t n_items, int dim, int topk, int precision) {
  EvalPlan p;
  if (make_plan(n_users, n_items, dim, topk, precision, &p)) return 0;
  return p.off_end;
}

template <int NSUB, int CL>
static int launch_eval_tc_cl(const EvalArgs& ea, const EvalPlan& p, cudaStream_t st) {
  const size_t smem = eval_smem_bytes(NSUB, p.nstages, p.KP);
  NNCF_CUDA(cudaFuncSetAttribute(eval_topk_tc_kernel<NSUB, CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)p.n_ub_grid, p.nsplit, 1);
  cfg.blockDim = dim3(kEvalThreads, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  NNCF_CUDA(cudaLaunchKernelEx(&cfg, eval_topk_tc_kernel<NSUB, CL>, ea));
  NNCF_LAUNCH_OK();
  return 0;
}
template <int NSUB, int CL>
static int launch_eval_tc2_cl(const EvalArgs& ea, const EvalPlan& p, cudaStream_t st) {
  const size_t smem = eval2_smem_bytes(NSUB, p.nstages2, p.cap);
  NNCF_CUDA(cudaFuncSetAttribute(eval_topk_tc2_kernel<NSUB, CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)p.n_ub_grid, p.nsplit, 1);
  cfg.blockDim = dim3(kEval2Threads, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  NNCF_CUDA(cudaLaunchKernelEx(&cfg, eval_topk_tc2_kernel<NSUB, CL>, ea));
  NNCF_LAUNCH_OK();
  return 0;
}
template <int NSUB>
static int launch_eval_tc3(const EvalArgs& ea, const EvalPlan& p, cudaStream_t st) {
  const size_t smem = eval3_smem_bytes(NSUB, p.nstages2, p.cap);
  NNCF_CUDA(cudaFuncSetAttribute(eval_topk_tc3_kernel<NSUB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)p.n_ub_grid, p.nsplit, 1);
  cfg.blockDim = dim3(kEval3Threads, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  NNCF_CUDA(cudaLaunchKernelEx(&cfg, eval_topk_tc3_kernel<NSUB>, ea));
  NNCF_LAUNCH_OK();
  return 0;
}
template <int NSUB>
static int launch_eval_tc(const EvalArgs& ea, const EvalPlan& p, cudaStream_t st) {
  if (p.v3) return launch_eval_tc3<NSUB>(ea, p, st);
  if (p.v2) {
    if (p.cl == 8) return launch_eval_tc2_cl<NSUB, 8>(ea, p, st);
    if (p.cl == 4) return launch_eval_tc2_cl<NSUB, 4>(ea, p, st);
    if (p.cl == 2) return launch_eval_tc2_cl<NSUB, 2>(ea, p, st);
    return launch_eval_tc2_cl<NSUB, 1>(ea, p, st);
  }
  if (p.cl == 4) return launch_eval_tc_cl<NSUB, 4>(ea, p, st);
  if (p.cl == 2) return launch_eval_tc_cl<NSUB, 2>(ea, p, st);
  return launch_eval_tc_cl<NSUB, 1>(ea, p, st);
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
  ea.k = topk; ea.KP = p.KP; ea.nsplit = p.nsplit; ea.nstages = p.nstages; ea.tiles_per_split = p.tiles_per_split;
  ea.cap = p.cap;
  { const char* e = getenv("NNCF_EVAL_PF"); ea.prefetch_ahead = e ? atoi(e) : 0; }   // measured: no effect (the stream is not DRAM-latency-bound)
  { const char* e = getenv("NNCF_EVAL_DBG"); ea.dbg_mode = e ? atoi(e) : 0; }
  { const char* e = getenv("NNCF_EVAL_PIPE"); ea.dbg_pipe = e ? atoi(e) : 0; }
#ifdef NNCF_EVAL_STATS_BUILD
  static unsigned long long* stats_dev = nullptr;
  if (getenv("NNCF_EVAL_STATS")) {
    if (!stats_dev) cudaMalloc(&stats_dev, 8 * sizeof(unsigned long long));
    cudaMemsetAsync(stats_dev, 0, 8 * sizeof(unsigned long long), st);
    ea.stats = stats_dev;
  }
#endif
  ea.lists = reinterpret_cast<uint64_t*>(ws + p.off_lists);
  if (precision == NNCF_PREC_BF16) {
    ea.Uimg = ws + p.off_uimg; ea.Vimg = ws + p.off_vimg;
    if (!p.v2) {   // (the second-generation kernel converts its user block itself, straight into tensor memory)
      rows_to_img_kernel<<<ceil_div(p.users_pad, 8), 256, 0, st>>>(user_rows_dev, n_users, dim, p.nsub,
                                                                   const_cast<uint8_t*>(ea.Uimg), p.users_pad);
      NNCF_LAUNCH_OK();
    }
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
    const size_t sm = ((size_t)32 * p.dp) * 4 + (size_t)128 * p.KP * 8 + 4 * 32 * 4;
    NNCF_CUDA(cudaFuncSetAttribute(eval_topk_simt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    eval_topk_simt_kernel<<<dim3((unsigned)p.n_ub, p.nsplit), 128, sm, st>>>(ea);
    NNCF_LAUNCH_OK();
  }
  topk_merge_kernel<<<ceil_div(n_users, 4), 128, 0, st>>>(ea.lists, n_users, p.nsplit, p.KP, topk, pow2_ge(2 * p.KP),
                                                          topk_ids_dev, topk_scores_dev);
  NNCF_LAUNCH_OK();
#ifdef NNCF_EVAL_STATS_BUILD
  if (ea.stats) {
    unsigned long long h[8];
    cudaStreamSynchronize(st);
    cudaMemcpy(h, ea.stats, sizeof(h), cudaMemcpyDeviceToHost);
    fprintf(stderr, "[eval stats] warp-tiles %llu, with a hit %.3f, hit rows / warp-tile %.3f, appended / warp-tile %.3f, compactions / warp-tile %.4f, "
            "filter cycles / warp-tile %.0f (lane-0 counts are per warp for appended only in the per-lane filter)\n",
            h[0], (double)h[1] / h[0], (double)h[2] / h[0], (double)h[3] / h[0], (double)h[4] / h[0], (double)h[5] / h[0]);
  }
#endif
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
                               int64_t n_groups, int topk, float* per_group_dev, void* stream) {
  NNCF_CHECK_ARG(scores_dev && truth_dev && seg_indptr_dev && per_group_dev, "nncf_eval_given: null argument");
  if (n_groups == 0) return NNCF_OK;
  NNCF_CHECK_ARG(topk == -1 || topk >= 1, "nncf_eval_given: topk must be -1 (whole list) or >= 1");
  eval_given_kernel<<<ceil_div(n_groups, 4), 128, 0, (cudaStream_t)stream>>>(scores_dev, truth_dev, seg_indptr_dev,
                                                                             n_groups, topk, per_group_dev);
  NNCF_LAUNCH_OK();
  return NNCF_OK;
}
