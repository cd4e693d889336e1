// score_tc.cuh — score + gradient on the 5th-gen tensor cores (precision = bf16), one-sided formulation.
//
// Every CTA owns ONE 128-row block of one side X (X = U rows i when side = 0, X = V rows j when side = 1) and sweeps
// the 128-row blocks Y_t of the other side:
//     MMA1  S'_t  = X_ob Y_t^T       A, B K-major (smem)      -> TMEM, double buffered
//     epi   G'_t  = dL/dS'_t         TMEM -> registers -> bf16 -> back into TMEM, IN PLACE over the S' buffer
//                                    (+ loss / positive-score corrections, see below)
//     MMA2  dX_ob += G'_t Y_t        A from TMEM, B MN-major  -> TMEM accumulator (DP columns), drained ONCE
// G' never touches shared memory: the tile loop is bound by the 128 B/clk shared-memory port (MMA operand fetches, Y tile
// fills), and routing G' through TMEM takes its 16 KiB write, 16 KiB operand read, the proxy fence and the swizzled
// stores out of every tile.  tcgen05.mma instructions execute in issue order, so MMA1(t+2) overwriting the buffer that
// MMA2(t) reads needs no barrier.
// so the gradient of the owned rows is complete inside the CTA: no atomics, no second drain, and the score matrix
// exists only in TMEM.  The two sides recompute S (side 1 sees its transpose), which costs a second pass of the
// cheap K = d contraction but removes the B^2 d / 128 float atomics of a one-pass scheme (measured: 21k cycles per
// tile on B200).  side 0 also produces the loss and, for group pairwise losses, the per-row sums of dL/dD; side 1
// produces the per-column sums for neg_shared pairwise losses (both are per-thread running sums here).
//
// The kernel is specialised at compile time on <NSUB (= dp/64), LOSS, GROUP>: the epilogue is straight-line,
// branch-free code (a first version that switched on the loss per element ran 24k cycles per tile: the unrolled
// 4-way switch blew the instruction cache).
//
// Tile width TN (swept rows per tile) = 64.  For dp <= 128 a CTA needs ~100 KiB of shared memory and 256 TMEM columns, so
// TWO CTAs are resident per SM: one CTA's prologue (barrier init, TMEM alloc, X load) and drain overlap the other's tile
// loop, and 16 epilogue warps per SM hide the MUFU / TMEM-load latencies (measured at B = 512: the 1-CTA/SM version spent
// ~8k of its ~21k cycles per CTA outside the tile loop).  dp = 256: one CTA per SM, four 32 KiB stages (128-row tiles with
// two 64 KiB stages left the next-but-one copy exposed).
//
// Warp roles: warp 0 = bulk-copy (TMA engine) producer (+ the loss hand-off), warp 1 = MMA issuer, warps 2..9 = epilogue
// (warp w reads TMEM lanes 32*(w&3).. and the column half (w-2)>>2 (TN/2 columns) of every S tile), warps 10..11 = L2
// prefetch of the next step's rows (and of this step's Adam state) + the G' exchange's flag publisher.
//
// Two variants live behind compile-time switches, both measured slower and documented where they are declared (DESIGN.md
// 8.1): NNCF_SCORE_TEAMS (four 32-column S' buffers, the two warp groups on alternate tiles) and the G' exchange
// (template parameter GX: every block computed once by one side, the other side contracts the exchanged bf16 tiles).
#pragma once
#include "common.cuh"
#include "sm100.cuh"

namespace nncf {

// developer ablations (tools/score_bench.cu builds one binary per bit; the library is always built with 0):
//   1 no epilogue math   2 no Y tile copies   4 no MMA2   8 no MMA1   16 no loss   32 no table update in the drain   64 no L2 prefetch
#ifndef NNCF_ABLATE
#define NNCF_ABLATE 0
#endif

constexpr int kScoreEpiWarps = 8;
constexpr int kScoreThreads = 64 + 32 * kScoreEpiWarps + 64;   // 12 warps: registers are allocated in groups of 4 warps

struct ScoreTcArgs {
  const uint8_t* Uimg; const uint8_t* Vimg;   // [R][rows_pad/128][NSUB][16 KiB]
  float* dU; float* dV;                       // [R][rows_pad][dp]   plain stores (each row has one owner)
  float* corrU; float* corrV;                 // [R][rows_pad]
  const float* spos;                          // [R][rows_pad] positive score of batch row b
  const int32_t* inverse;                     // group: [R][rows_pad] compact column of batch row i, else NULL
  const int32_t* ncols_dev;                   // group: [R] number of unique items, else NULL
  double* loss;                               // [R]
  int rows_pad, B, scheme, loss_kind;
  float lambda, gamma;
  long long* dbg;                             // optional [grid][64] clock64 stamps (developer tool)
  unsigned long long* tl;                     // optional step timeline: [3] first CTA start, [4] first CTA past the wait, [5] last end
  // fused sparse SGD: the drain adds -lr * dX straight into the embedding rows instead of writing dX for a separate
  // update kernel: one bulk async reduction (cp.reduce.async.bulk .add.f32, TMA engine, performed at the L2; duplicates
  // sum there) per owned row, issued from the staged fp32 block in shared memory.  Legal because every row of this step
  // was gathered by the PREVIOUS kernel, so nothing in this kernel reads the tables.
  int fuse_sgd;
  // activity regulariser folded into the drain (ref: utils/utilities.py:129-135): the user-side CTAs add
  // reg_scale * X (reg_scale = 2 u_reg / B) to their gradient block from the bf16 X block in shared memory, so the
  // reference's default u_reg = 1e-6 keeps the two-launch step.  0 = off.
  float reg_scale;
  // split sweep: `split` CTAs share one owner block, each sweeps 1/split of the tiles and ADDS its partial gradient block
  // (bulk reductions into the table rows in fused mode, red.global.add.v4 into the zeroed dX blocks otherwise).  A batch of
  // 512 is only 8 CTAs: at R = 1 (the reference's sequential loop) the sweep is the critical path of the step.
  int split;
  int drain_vec;   // fused drain with vector reductions instead of bulk reductions (small grids), see the drain
  int dedup;       // fused drain, bit `side`: rows of a warp's 32 that share an id are summed in shared memory first (crowded ids)
  // fused mode has no finalize launch: the last side-0 CTA of a replica publishes the replica's loss and re-arms the
  // accumulators (loss_count[R] arrival counters, zero between steps)
  unsigned int* loss_count;
  float* loss_out;                            // [R] or NULL
  // L2 prefetch of the NEXT step's rows (ids known in advance; the two spare warps of every CTA issue one bulk L2
  // prefetch per row while the tensor pipe works, so the next gather finds its rows in L2).  NULL = no next step.
  const int32_t* next_ids_u; const int32_t* next_ids_v;
  int next_count;                             // ids per side (R * B)
  int64_t n_rows_u, n_rows_v;                 // table sizes (ids are clamped: a hint must never fault)
  int d;                                      // true embedding dim (table row stride), d % 4 == 0 when fused
  float neg_lr;
  float* table_u; float* table_v;             // what the fused drain adds into: the embedding tables (sparse SGD) or the
                                              // per-table gradient accumulators of the folded lazy Adam (neg_lr = 1)
  const float* pf_table_u; const float* pf_table_v;   // tables whose rows the spare warps prefetch for the next step
  // lazy Adam (folded): the rows of THIS step's ids in the optimizer-state tables (m, v) and in the gradient accumulators are
  // pulled into L2 by the same warps while the tile loops run, so the apply launch behind this kernel - a claim -> rows
  // dependency chain, latency-bound - finds them there.  NULL = off.
  const float* pf_now[2][3];                          // [side][m, v, acc]
  int pf_now_count;                                   // ids per side (R * B)
  const int32_t* ids_u; const int32_t* ids_v; // table row of each owner row: ids_x[r * stride_x + row]
  int64_t ids_stride_u, ids_stride_v;
  ShardPtrs shards_u, shards_v;
  // self-gather (fused SGD, neg_shared, dp <= 128, one table, the whole grid co-resident): there is no gather kernel.
  // Every CTA gathers ITS OWN 128 rows (ids -> fp32 rows -> bf16 tile image) straight into its shared-memory X block
  // and into the global image, where the CTAs of the other side read them as their Y tiles once the block's flag
  // carries this step's sequence number.  The tables are read and updated inside one kernel now, so no drain may
  // start before every CTA has finished reading: gather_count (monotonic) must reach gather_target first.
  // G' exchange ("two-sided": every score, sigmoid and loss term is computed ONCE; neg_shared with a pointwise loss,
  // dp <= 128, B % 256 == 0, no split).  The 128 x 128 block (user block a, item block b) is computed by the user-side CTA
  // of a when a + b is even and by the item-side CTA of b when it is odd, so EVERY CTA runs the score / loss / gradient
  // epilogue on half of its sweep (its "own" blocks: MMA1 -> epilogue -> MMA2 as before).  It also stores each own G' tile
  // (bf16) into gx_buf as the 128B-swizzled tile image [128 of my rows][128 B = 64 swept columns] - an MN-major A operand
  // for the CTA on the other side of the block - and raises gx_flags[(r, side, my block, swept block)] = gx_seq once both
  // 64-column halves are written.  The other half of its gradient comes from those images: after its own tiles the
  // producer warp waits for the flag of each "foreign" block, copies the image and the swept block's rows with the
  // bulk-copy engine, and the MMA warp accumulates dX += G'^T Y - a plain contraction, no epilogue.  Own blocks are swept
  // and foreign blocks consumed in the same rotating order, so a CTA's i-th foreign block is its producer's i-th own
  // block.  All flags of a step carry the step's sequence number.
  int gx;
  int gx_seq;
  uint8_t* gx_buf;                            // [R][2 sides][nblk mine][nblk swept][2 halves][16 KiB]
  int* gx_flags;                              // [R][2 sides][nblk mine][nblk swept]
  int self_gather;
  int gather_seq;                             // this step's sequence number (> 0, increasing)
  int* gather_flags;                          // [R][2][rows_pad / 128]
  unsigned long long* gather_count;           // CTAs that have finished gathering, summed over all steps
  unsigned long long gather_target;           // = gather_seq * (active CTAs per step)
};

// declared here, defined in score_tc_nsub{1,2,4}.cu (one translation unit per NSUB so they compile in parallel)
int launch_score_tc_nsub1(const ScoreTcArgs& a, int nblk, int R, cudaStream_t st);
int launch_score_tc_nsub2(const ScoreTcArgs& a, int nblk, int R, cudaStream_t st);
int launch_score_tc_nsub4(const ScoreTcArgs& a, int nblk, int R, cudaStream_t st);

template <int NSUB>
struct ScoreTcCfg {
  static constexpr int DP = 64 * NSUB;
#ifndef NNCF_SCORE_TN4
#define NNCF_SCORE_TN4 64
#endif
  // swept rows per tile = columns of one S' tile.  dp = 256 used 128-row tiles with 2 stages of 64 KiB: a stage is held
  // from MMA1 through the epilogue to MMA2, so the next-but-one tile's copy (64 KiB, ~2k cycles) could only start after
  // MMA2 and was fully exposed (tensor pipe ~50 % busy at C5).  64-row tiles: 32 KiB stages, four of them in flight.
#ifndef NNCF_SCORE_TEAMS
#define NNCF_SCORE_TEAMS 0
#endif
  // "teams" (compile-time option, measured and NOT used): 32-row tiles, FOUR S' buffers of 32 columns, the two groups of four
  // epilogue warps take alternate tiles instead of the two column halves of every tile.  The idea: the tile loop is a
  // dependent chain per S' buffer (S' ready -> TMEM load -> G' -> TMEM store -> barrier -> MMA2 -> MMA1 of the tile that
  // reuses the buffer -> commit: ~600 cycles with everything but the synchronisation switched off, tools/score_bench.cu
  // ablation 63), tensor memory fixes the columns in flight (128 next to the accumulator) but not their granularity, so four
  // shorter chains should fill the bubbles of two long ones.  Measured (tools/gpu_experiments_r02.sh teams, parity green): C3 22.9 vs 21.0 us
  // per step, B = 4,096: 130.6 vs 104.8 us.  The loop is not waiting for round trips, it is MUFU-bound: one MUFU.TANH per
  // score = 8.2k cycles per SM sub-partition and step against a measured 12.2k for the loop, and N = 32 MMAs only double the
  // issue work of the MMA warp.  What halves the MUFU work is computing every sigmoid once: the G' exchange below.
  static constexpr bool kTeams = (NSUB <= 2) && NNCF_SCORE_TEAMS;
  static constexpr int TN = kTeams ? 32 : (NSUB <= 2 ? 64 : NNCF_SCORE_TN4);
  static constexpr int CW = kTeams ? 32 : TN / 2;     // S' columns per epilogue warp
  static constexpr int kBufs = kTeams ? 4 : 2;        // S' buffers in tensor memory
  static constexpr int kYBytes = TN * 128;            // one [TN rows x 64 bf16] piece of a Y tile (8 or 16 KiB)
#ifndef NNCF_SCORE_STAGES
#define NNCF_SCORE_STAGES 4
#endif
  // Two resident CTAs of a tcgen05 kernel have (228 KiB - 2 x (1 KiB reserved + 1 KiB tcgen05 block)) / 2 = 112 KiB of
  // dynamic shared memory each (measured with tools/occ_probe.cu), barriers included.
  static constexpr int kStages = kTeams ? 2 * NNCF_SCORE_STAGES : ((NSUB <= 2 || TN == 64) ? NNCF_SCORE_STAGES : 2);   // Y tiles in flight (the bulk-copy latency is ~2k cycles)
  static_assert(kStages <= 8, "barrier block: eight Y stages");
  static constexpr int kColDX = kBufs * TN;           // the S' buffers take TMEM columns [0, kBufs TN)
  static constexpr int kTmemCols = (kColDX + DP) <= 256 ? 256 : 512;
  static constexpr int kMinBlocks = NSUB <= 2 ? 2 : 1;   // resident CTAs per SM
  // no alignment slack: the dynamic shared window starts 1024-byte aligned (checked at kernel entry); two CTAs of
  // dp = 128 need 2 x (112 KiB + 256 B + 1 KiB reserved) <= 228 KiB
  // layout: [barriers, 1 KiB][X block][Y stages]; at the drain the fp32 staging block [128][DP + 4] overlays the Y stages
  // only, so the bf16 X block stays readable (the activity regulariser adds reg_scale * x there)
  static constexpr size_t kBarBytes = 1024;
  static constexpr size_t kYAllBytes = (size_t)kStages * NSUB * kYBytes;
  static constexpr size_t kStageBytes = (size_t)128 * (DP + 4) * 4;
  static constexpr size_t kSmemBytes = kBarBytes + (size_t)NSUB * kSubBytes + (kYAllBytes > kStageBytes ? kYAllBytes : kStageBytes);
};

__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// packed bf16x2 arithmetic (two elements per instruction, and per MUFU op)
__device__ __forceinline__ uint32_t bf2_mul(uint32_t a, uint32_t b) {
  uint32_t d; asm("mul.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d;
}
__device__ __forceinline__ uint32_t bf2_fma(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t d; asm("fma.rn.bf16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d;
}
__device__ __forceinline__ uint32_t bf2_tanh(uint32_t a) {
  uint32_t d; asm("tanh.approx.bf16x2 %0, %1;" : "=r"(d) : "r"(a)); return d;
}
constexpr uint32_t kBf2Half = 0x3F003F00u;   // (0.5, 0.5)

struct EpiConst {
  float w_neg, gamma, inv_b, inv_cnt, ns_margin;   // ns_margin = gamma for neg_shared (zero margin at the positive), else 0
  float inv_wneg;   // 1 / w_neg
  float g_scale;    // skip-gram stores G' = (w / w_neg) (sigmoid - pos) in bf16 and applies w_neg / B to the fp32
                    // accumulator at the drain (keeps sigmoid's full bf16 precision, no constant rounded to bf16)
};

// skip-gram fast path: 32 scores of a full tile without positives.  G' = sigmoid(s) through packed bf16x2 math
// (1 MUFU.TANH per PAIR); the loss uses sum softplus(s) = sum max(s,0) - ln(prod sigmoid(|s|)): ONE MUFU.LG2 per chunk.
template <bool NEED_LOSS>
__device__ __forceinline__ void epi_chunk_sg_fast(const float (&v)[32], uint32_t (&pk)[16], float& lraw) {
  float summax = 0.0f, prod0 = 1.0f, prod1 = 1.0f;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const uint32_t xb = pack_bf16x2(v[2 * i], v[2 * i + 1]);
    const uint32_t t = bf2_tanh(bf2_mul(xb, kBf2Half));                 // tanh(s / 2)
    pk[i] = bf2_fma(t, kBf2Half, kBf2Half);                             // sigmoid(s)
    if (NEED_LOSS) {
      const uint32_t sa = bf2_fma(t & 0x7FFF7FFFu, kBf2Half, kBf2Half); // sigmoid(|s|) in [0.5, 1]
      prod0 *= __uint_as_float(sa << 16);
      prod1 *= __uint_as_float(sa & 0xFFFF0000u);
      summax += fmaxf(v[2 * i], 0.0f) + fmaxf(v[2 * i + 1], 0.0f);
    }
  }
  if (NEED_LOSS) lraw += summax - __logf(prod0 * prod1);                // product of 32 values >= 2^-32: safe in fp32
}

// One element, straight-line.  w = loss weight of the element, posf = 1 for the row's positive else 0,
// sp = positive score it is compared with (pairwise losses).
template <int LOSS, bool NEED_LOSS>
__device__ __forceinline__ void epi_val(const EpiConst& c, float s, float w, float posf, float sp, float& g, float& a,
                                        float& l) {
  a = 0.0f; l = 0.0f;
  // The epilogue is MUFU-bound on B200 (16 special-function lanes per SM), so every transcendental counts:
  //   sigmoid(|x|) = 0.5 * tanh(|x| / 2) + 0.5                      one MUFU.TANH instead of EX2 + RCP
  //   softplus(x)  = max(x, 0) - ln(sigmoid(|x|))                    one MUFU.LG2, argument in [0.5, 1)
  if (LOSS == NNCF_LOSS_SKIP_GRAM) {
    const float sa = fmaf(0.5f, tanh_approx(0.5f * fabsf(s)), 0.5f);            // sigmoid(|s|)
    const float sg = (s >= 0.0f) ? sa : 1.0f - sa;                              // sigmoid(s)
    if (NEED_LOSS) l = w * (fmaxf(s, 0.0f) - __logf(sa) - posf * s) * c.inv_b;  // softplus(s) [- s for the positive]
    g = w * c.inv_wneg * (sg - posf);                                           // G' (scaled by w_neg / B at the drain)
  } else if (LOSS == NNCF_LOSS_MSE) {
    const float t = s - posf;
    if (NEED_LOSS) l = w * t * t * c.inv_b;
    g = 2.0f * w * t * c.inv_b;
  } else if (LOSS == NNCF_LOSS_LOG_LOSS) {
    const float x = -c.gamma * (sp - s);                                        // -gamma * D
    const float sa = fmaf(0.5f, tanh_approx(0.5f * fabsf(x)), 0.5f);
    const float sg = (x >= 0.0f) ? sa : 1.0f - sa;                              // sigmoid(-gamma D)
    if (NEED_LOSS) l = (fmaxf(x, 0.0f) - __logf(sa)) * c.inv_cnt;
    a = -c.gamma * sg * c.inv_cnt;
    g = -a;
  } else {
    const float t = (c.gamma - posf * c.ns_margin) - (sp - s);                  // M - D
    if (NEED_LOSS) l = fmaxf(t, 0.0f) * c.inv_cnt;
    a = (t > 0.0f) ? -c.inv_cnt : 0.0f;
    g = (t > 0.0f) ? 1.0f : 0.0f;                                               // G' is an indicator; 1 / count is applied at the drain (g_scale)
  }
}

// max-margin fast path: 32 scores of a full tile without positives.  The gradient of relu(gamma - (sp - s)) is an
// indicator: G' = 1 or 0 as a bf16 bit pattern (one add, one compare, one select per element), scaled by 1 / count at the
// drain; the hinge sum and the number of active elements are accumulated raw and scaled once per chunk.
template <bool SIDE1, bool GROUP>
__device__ __forceinline__ void epi_chunk_mm_fast(const EpiConst& c, const float (&v)[32], int x0, bool row_ok, float my_sp,
                                                  const float* __restrict__ spos_row, float& lsum, float& asum, uint32_t (&pk)[16]) {
  constexpr bool kSpMine = (SIDE1 != GROUP);        // positive score constant along my row
  float hinge = 0.0f;
  int cnt = 0;
  const float gs_mine = c.gamma - my_sp;
#pragma unroll
  for (int u4 = 0; u4 < 32; u4 += 4) {
    float4 sv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (!kSpMine) sv = __ldg(reinterpret_cast<const float4*>(spos_row + x0 + u4));
    const float gs[4] = {kSpMine ? gs_mine : c.gamma - sv.x, kSpMine ? gs_mine : c.gamma - sv.y,
                         kSpMine ? gs_mine : c.gamma - sv.z, kSpMine ? gs_mine : c.gamma - sv.w};
#pragma unroll
    for (int k = 0; k < 4; k += 2) {
      const float t0 = gs[k] + v[u4 + k], t1 = gs[k + 1] + v[u4 + k + 1];       // M - D with M = gamma (no positive here)
      const bool p0 = t0 > 0.0f, p1 = t1 > 0.0f;
      pk[(u4 + k) >> 1] = (p0 ? 0x00003F80u : 0u) | (p1 ? 0x3F800000u : 0u);    // bf16 (1.0 | 0.0) pair
      if (!SIDE1) hinge += fmaxf(t0, 0.0f) + fmaxf(t1, 0.0f);
      cnt += (p0 ? 1 : 0) + (p1 ? 1 : 0);
    }
  }
  if (row_ok) {
    if (!SIDE1) lsum += hinge * c.inv_cnt;
    asum -= static_cast<float>(cnt) * c.inv_cnt;
  } else {
#pragma unroll
    for (int i = 0; i < 16; ++i) pk[i] = 0u;
  }
}

// 32 consecutive swept columns of my row.  GENERAL = false: the tile is full and holds no positive (fast path).
template <int LOSS, bool SIDE1, bool GROUP, bool GENERAL>
__device__ __forceinline__ void epi_chunk(const EpiConst& c, float (&v)[32], int x0, int n_other, bool row_ok, int o,
                                          int my_posc, float my_sp, const int32_t* __restrict__ inv_row,
                                          const float* __restrict__ spos_row, float& lsum, float& asum,
                                          uint32_t (&pk)[16]) {
  constexpr bool kPairwise = LOSS >= NNCF_LOSS_LOG_LOSS;
  constexpr bool kSpMine = kPairwise && (SIDE1 != GROUP);       // positive score constant along my row
  constexpr bool kPosByInverse = SIDE1 && GROUP;                // positive test needs inverse[x] of the swept row
#pragma unroll
  for (int u4 = 0; u4 < 32; u4 += 4) {
    int4 iv = make_int4(0, 0, 0, 0);
    float4 sv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (GENERAL && kPosByInverse) iv = __ldg(reinterpret_cast<const int4*>(inv_row + x0 + u4));
    if (kPairwise && !kSpMine) sv = __ldg(reinterpret_cast<const float4*>(spos_row + x0 + u4));
    const int ivs[4] = {iv.x, iv.y, iv.z, iv.w};
    const float svs[4] = {sv.x, sv.y, sv.z, sv.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int u = u4 + k;
      const int x = x0 + u;
      float w = c.w_neg, posf = 0.0f;
      if (GENERAL) {
        const bool pos = kPosByInverse ? (ivs[k] == o) : (x == my_posc);
        w = pos ? 1.0f : c.w_neg;
        posf = pos ? 1.0f : 0.0f;
      }
      const float sp = kSpMine ? my_sp : svs[k];
      float g, aa, ll;
      epi_val<LOSS, !SIDE1>(c, v[u], w, posf, sp, g, aa, ll);
      const bool valid = GENERAL ? (row_ok && x < n_other) : row_ok;
      v[u] = valid ? g : 0.0f;
      if (!SIDE1) lsum += valid ? ll : 0.0f;
      if (kPairwise) asum += valid ? aa : 0.0f;
    }
  }
#pragma unroll
  for (int i = 0; i < 16; ++i) pk[i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
}

template <int NSUB, int LOSS, bool GROUP, bool GX = false>
__global__ void __launch_bounds__(kScoreThreads, ScoreTcCfg<NSUB>::kMinBlocks)
score_grad_tc_kernel(ScoreTcArgs a) {
  static_assert(!GX || (!GROUP && LOSS < NNCF_LOSS_LOG_LOSS && NSUB <= 2 && !ScoreTcCfg<NSUB>::kTeams), "G' exchange: neg_shared, pointwise loss, dp <= 128");
  using C = ScoreTcCfg<NSUB>;
  constexpr int DP = C::DP;
  constexpr int TN = C::TN;
  constexpr int CW = C::CW;
  constexpr int NV = CW / 32;                     // 32-column TMEM loads per epilogue warp per tile
  constexpr bool kPairwise = LOSS >= NNCF_LOSS_LOG_LOSS;
  extern __shared__ __align__(1024) uint8_t smem[];
  if (smem_u32(smem) & 1023u) __trap();           // SWIZZLE_128B operands need 1024-byte aligned tiles
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  uint8_t* sX = smem + C::kBarBytes;
  uint8_t* sY = sX + NSUB * kSubBytes;
  uint64_t* x_full = bars + 0;
  uint64_t* y_full = bars + 1;      // [kStages <= 8]
  uint64_t* y_empty = bars + 9;     // [kStages <= 8]
  uint64_t* s_full = bars + 17;     // [kBufs <= 4]  S'(t) is in TMEM buffer t % kBufs
  uint64_t* g_full = bars + 21;     // [kBufs <= 4]  G'(t) has replaced it
  uint64_t* dx_full = bars + 25;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 26);
  float* loss_slots = reinterpret_cast<float*>(bars + 28);   // [8] per-epilogue-warp loss sums, handed to warp 0 (see below)
  uint32_t* gx_cnt = reinterpret_cast<uint32_t*>(bars + 32);   // G' exchange: epilogue warps that have stored their part of a tile (monotonic)
  constexpr int kBufs = C::kBufs;
  constexpr int kGArrive = C::kTeams ? kScoreEpiWarps / 2 : kScoreEpiWarps;   // epilogue warps that write one G' tile

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nblk_grid = a.rows_pad >> 7;
  const int ob = blockIdx.x % nblk_grid, sp = blockIdx.x / nblk_grid, side = blockIdx.y, r = blockIdx.z;
  pdl_launch_dependents();                         // the next kernel may start its prologue (it waits for us before reading)
  if (a.tl && tid == 0) atomicMin(&a.tl[3], global_timer_ns());
  if (GROUP) pdl_wait();                           // n_unique comes from a preceding kernel
  const int ncols = GROUP ? a.ncols_dev[r] : a.B;
  const int n_owner = side == 0 ? a.B : ncols;     // valid rows on the owner side
  const int n_other = side == 0 ? ncols : a.B;     // valid rows on the swept side
  if (ob * 128 >= n_owner) return;                 // whole CTA exits together (no barrier touched yet)
  const int nt_all = (n_other + TN - 1) / TN;
  const int t_begin = (nt_all * sp) / a.split;        // this CTA sweeps tiles [t_begin, t_begin + nt)
  const int nt = (nt_all * (sp + 1)) / a.split - t_begin;
  if (nt <= 0) return;                             // (the host never asks for more splits than tiles)
  const int nblk = a.rows_pad >> 7;
  const int64_t base = (int64_t)r * a.rows_pad;
  // G' exchange: my i-th own block is swept block (ob + side + 2 i) % nblk, my i-th foreign block (ob + side - 1 - 2 i) % nblk
  // (nblk is even); nt_loop = tiles this CTA runs through MMA1 / epilogue / MMA2
  constexpr int kTilesPerBlk = 128 / C::TN;
  const int nt_loop = GX ? nt / 2 : nt;
  auto gx_tile = [&](int it) -> int {                    // sweep position of my it-th own tile
    return ((ob + side + 2 * (it / kTilesPerBlk)) % nblk) * kTilesPerBlk + it % kTilesPerBlk;
  };
  constexpr bool bal_loss = (LOSS == NNCF_LOSS_SKIP_GRAM && !GROUP) || GX;  // both sides hold a share of the loss (see the epilogue)
  constexpr int kGxStages = 2;                           // foreign blocks arrive in stages of kGxRows rows: [G' 2 x kGxPiece][rows NSUB x kGxPiece]
  constexpr int kGxRows = NSUB >= 2 ? 64 : 32;           // (dp = 64: the Y stages it overlays are 32 KiB)
  constexpr uint32_t kGxPiece = kGxRows * 128;           // kGxRows rows of one [128 x 64] sub-tile: contiguous in the tile image
  constexpr uint32_t kGxStageBytes = (2 + NSUB) * kGxPiece;
  static_assert(!GX || kGxStages * kGxStageBytes <= C::kYAllBytes, "G' exchange stages overlay the Y stages (the X block stays: the regulariser reads it)");
  uint64_t* gx_full = bars + 34;    // [2]
  uint64_t* gx_empty = bars + 36;   // [2]
  const uint8_t* gX = (side == 0 ? a.Uimg : a.Vimg) + ((int64_t)r * nblk + ob) * NSUB * kSubBytes;
  const uint8_t* gY = (side == 0 ? a.Vimg : a.Uimg) + (int64_t)r * nblk * NSUB * kSubBytes;

  if (tid == 0) {
    mbar_init(x_full, 1);
    for (int s = 0; s < C::kStages; ++s) { mbar_init(&y_full[s], 1); mbar_init(&y_empty[s], 1); }
    for (int s = 0; s < kBufs; ++s) { mbar_init(&s_full[s], 1); mbar_init(&g_full[s], kGArrive); }
    mbar_init(dx_full, 1);
    for (int s = 0; s < kGxStages; ++s) { mbar_init(&gx_full[s], 1); mbar_init(&gx_empty[s], 1); }
    *gx_cnt = 0u;
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, C::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (!GROUP) pdl_wait();                          // barriers, TMEM and descriptors are set up; now wait for the gather
  long long tl_c0 = 0; unsigned long long tl_g0 = 0;
  if (a.tl && tid == 0) {
    tl_g0 = global_timer_ns(); tl_c0 = clock64();
    atomicMin(&a.tl[4], tl_g0); atomicMax(&a.tl[8], tl_g0);
  }
  long long* dbg = a.dbg ? a.dbg + ((blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 64 : nullptr;
#define NNCF_STAMP(slot) do { if (dbg) dbg[slot] = clock64(); } while (0)
  if (tid == 0) NNCF_STAMP(0);
  if (tid == 0 && dbg) { dbg[40] = static_cast<long long>(global_timer_ns()); dbg[42] = static_cast<long long>(sm_id()); }

  if (warp == 0) {
    // ------------------------------------------------------------------------------ producer
    if (lane == 0) {
      if (!a.self_gather) {
        mbar_expect_tx(x_full, NSUB * kSubBytes);
        for (int s = 0; s < NSUB; ++s) bulk_g2s(sX + s * kSubBytes, gX + (size_t)s * kSubBytes, kSubBytes, x_full);
      }
      int ready_blk = -1;                                // self-gather: blocks of the swept side known to be in the image
      for (int it = 0; it < nt_loop; ++it) {
        const int t = GX ? gx_tile(it) : t_begin + it;
        const int st = it % C::kStages;
        if (a.self_gather && ((t * TN) >> 7) > ready_blk) {
          ready_blk = (t * TN) >> 7;
          const int* flag = a.gather_flags + (r * 2 + (1 - side)) * nblk + ready_blk;
          uint32_t spins = 0;
          while (ld_acquire_gpu(flag) != a.gather_seq) { __nanosleep(20); if (++spins > 20000000u) __trap(); }
          fence_proxy_async_global();                    // the rows were written by generic stores; the bulk copy reads them
        }
        mbar_wait(&y_empty[st], ((it / C::kStages) & 1) ^ 1);
        if (NNCF_ABLATE & 2) { mbar_arrive(&y_full[st]); continue; }
        mbar_expect_tx(&y_full[st], NSUB * C::kYBytes);
        // rows [t TN, (t+1) TN) of the swept side: a contiguous piece of each [128 x 64] sub-tile of block (t TN) / 128
        const uint8_t* src = gY + (size_t)((t * TN) >> 7) * NSUB * kSubBytes + (size_t)((t * TN) & 127) * 128;
        for (int s = 0; s < NSUB; ++s)
          bulk_g2s(sY + (st * NSUB + s) * C::kYBytes, src + (size_t)s * kSubBytes, C::kYBytes, &y_full[st]);
      }
      if (GX) {
        // foreign blocks: the Y stages are free once the MMA2 of my last own tiles has read them (the waits a tile
        // nt_loop + j would do); then, per foreign block, the flag of the CTA that computed it, and two stages of 64 rows:
        // its G' image (A operand, K = its rows) and its rows (B operand)
        for (int j = 0; j < C::kStages; ++j) mbar_wait(&y_empty[(nt_loop + j) % C::kStages], (((nt_loop + j) / C::kStages) & 1) ^ 1);
        int i = 0;
        for (int f = 0; f < nblk / 2; ++f) {
          const int fb = ((ob + side - 1 - 2 * f) % nblk + nblk) % nblk;
          const int64_t pair = (((int64_t)r * 2 + (1 - side)) * nblk + fb) * nblk + ob;
          const int* flag = a.gx_flags + pair;
          uint32_t spins = 0;
          while (ld_acquire_gpu(flag) != a.gx_seq) { __nanosleep(32); if (++spins > 40000000u) __trap(); }
          fence_proxy_async_global();                    // G' was written with generic stores; the bulk copy reads it
          const uint8_t* gsrc = a.gx_buf + pair * 2 * kSubBytes;
          const uint8_t* ysrc = gY + (size_t)fb * NSUB * kSubBytes;
          for (int kh = 0; kh < 128 / kGxRows; ++kh, ++i) {
            const int st = i % kGxStages;
            mbar_wait(&gx_empty[st], ((i / kGxStages) & 1) ^ 1);
            mbar_expect_tx(&gx_full[st], kGxStageBytes);
            uint8_t* dst = sY + st * kGxStageBytes;
            for (int m = 0; m < 2; ++m) bulk_g2s(dst + m * kGxPiece, gsrc + (size_t)m * kSubBytes + kh * kGxPiece, kGxPiece, &gx_full[st]);
            for (int s2 = 0; s2 < NSUB; ++s2) bulk_g2s(dst + (2 + s2) * kGxPiece, ysrc + (size_t)s2 * kSubBytes + kh * kGxPiece, kGxPiece, &gx_full[st]);
          }
        }
      }
    }
    if (side == 0 || bal_loss) {
      // loss hand-off for the whole CTA (the epilogue warps have moved on to the drain)
      __syncwarp();                                      // lane 0 comes out of the producer loop: the barrier below is .aligned
      asm volatile("bar.sync 2, %0;" ::"n"(32 * kScoreEpiWarps + 32) : "memory");
      if (lane == 0) {
        float cta_loss = 0.0f;
#pragma unroll
        for (int w = 0; w < kScoreEpiWarps; ++w) cta_loss += loss_slots[w];
        atomicAdd(&a.loss[r], static_cast<double>(cta_loss));
        if (a.loss_count) {
          // last arriving CTA of the replica publishes the loss and re-arms the accumulators.  One acq_rel atomic on the
          // counter orders my loss contribution before it and everybody's contributions before the read below
          // (instead of __threadfence() = MEMBAR.SC.GPU on either side of a relaxed atomic)
          // (split CTAs whose tile range is empty left before the barriers: min(split, tiles) of them arrive per owner block)
          const unsigned int expect = static_cast<unsigned int>((n_owner + 127) >> 7) * static_cast<unsigned int>(a.split < nt_all ? a.split : nt_all) *
                                      (bal_loss ? 2u : 1u);
          unsigned int arrived;
          asm volatile("atom.add.acq_rel.gpu.global.u32 %0, [%1], 1;" : "=r"(arrived) : "l"(a.loss_count + r) : "memory");
          if (arrived + 1u == expect) {
            const double total = atomicAdd(&a.loss[r], 0.0);
            if (a.loss_out) a.loss_out[r] = static_cast<float>(total);
            a.loss[r] = 0.0;
            a.loss_count[r] = 0u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------------------ MMA issuer
    // The whole warp runs the loop (warp-uniform control flow, descriptors and tensor-memory addresses in uniform
    // registers, one UIADD3.64 per MMA); one elected lane issues.  Under `if (lane == 0)` ptxas rebuilt every descriptor
    // and moved the addresses through an ELECT / R2UR.BROADCAST / BRA.U.ANY sequence: ~160 cycles per MMA, against 32-64
    // cycles of tensor work.
    {
      const uint32_t idesc_s = make_idesc_bf16(128, TN, 0, 0);
      const uint32_t idesc_dx = make_idesc_bf16(128, DP, 0, 1);
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
      const uint64_t xdesc0 = make_smem_desc(smem_u32(sX), 16, 1024);            // K-major X block
      const uint64_t ydesc0 = make_smem_desc(smem_u32(sY), 16, 1024);            // K-major view of a Y stage (MMA1)
      const uint64_t ydesc0_mn = make_smem_desc(smem_u32(sY), C::kYBytes, 1024); // MN-major view of the same bytes (MMA2)
      auto issue_mma1 = [&](int t) {
        const int st = t % C::kStages, sb = t % kBufs;
        mbar_wait(&y_full[st], (t / C::kStages) & 1);
        // buffer sb was last read by MMA2(t - kBufs), issued earlier by this warp: the tensor pipe runs in issue order
        if (elect_one()) {
          const uint64_t yd = ydesc0 + static_cast<uint64_t>((st * NSUB * C::kYBytes) >> 4);
          const uint32_t dcol = tmem_u + sb * TN;
#pragma unroll
          for (int k = 0; k < ((NNCF_ABLATE & 8) ? 0 : DP / 16); ++k)
            umma_bf16(dcol, xdesc0 + static_cast<uint64_t>(((k >> 2) * kSubBytes + (k & 3) * 32) >> 4),
                      yd + static_cast<uint64_t>(((k >> 2) * C::kYBytes + (k & 3) * 32) >> 4), idesc_s, k > 0);
          umma_commit(&s_full[sb]);
        }
        __syncwarp();
      };
      mbar_wait(x_full, 0);
      if (lane == 0) NNCF_STAMP(1);
      constexpr int kAhead = kBufs - 1;         // MMA1 runs this many tiles ahead of MMA2
      for (int p = 0; p < kAhead && p < nt_loop; ++p) issue_mma1(p);
      if (lane == 0) NNCF_STAMP(2);
      for (int t = 0; t < nt_loop; ++t) {
        const int st = t % C::kStages, sb = t % kBufs;
        if (t + kAhead < nt_loop) issue_mma1(t + kAhead);
        mbar_wait(&g_full[sb], (t / kBufs) & 1);
        if (lane == 0 && t < 8) NNCF_STAMP(8 + t);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t yd = ydesc0_mn + static_cast<uint64_t>((st * NSUB * C::kYBytes) >> 4);
          const uint32_t acol0 = tmem_u + sb * TN;
#pragma unroll
          for (int k = 0; k < ((NNCF_ABLATE & 4) ? 0 : TN / 16); ++k) {   // K = the TN swept rows of this tile, 16 per MMA = 8 TMEM columns of G'
            // K rows 16k.. were written by epilogue column-half (16k) / CW at column offset ((16k) % CW) / 2 of its own range
            umma_bf16_ts(tmem_u + C::kColDX, acol0 + ((16 * k) / CW) * CW + ((16 * k) % CW) / 2,
                         yd + static_cast<uint64_t>((k * 2048) >> 4), idesc_dx, (t > 0) || (k > 0));
          }
          umma_commit(&y_empty[st]);
        }
        __syncwarp();
      }
      if (GX) {
        // foreign blocks: dX += G'^T Y, both operands MN-major (rows = the other side's rows = K), 64 rows per stage
        const uint32_t idesc3 = make_idesc_bf16(128, DP, 1, 1);
        const uint64_t gdesc0 = make_smem_desc(smem_u32(sY), kGxPiece, 1024);                  // A: [kGxRows rows][2 x 64 of my rows]
        const uint64_t ydesc3 = make_smem_desc(smem_u32(sY) + 2 * kGxPiece, kGxPiece, 1024);   // B: [kGxRows rows][NSUB x 64 columns of d]
        for (int i = 0; i < (nblk / 2) * (128 / kGxRows); ++i) {
          const int st = i % kGxStages;
          mbar_wait(&gx_full[st], (i / kGxStages) & 1);
          tc_fence_after();
          if (elect_one()) {
            const uint64_t off = static_cast<uint64_t>((st * kGxStageBytes) >> 4);
#pragma unroll
            for (int k = 0; k < kGxRows / 16; ++k)
              umma_bf16(tmem_u + C::kColDX, gdesc0 + off + static_cast<uint64_t>((k * 2048) >> 4),
                        ydesc3 + off + static_cast<uint64_t>((k * 2048) >> 4), idesc3, 1u);
            umma_commit(&gx_empty[st]);
          }
          __syncwarp();
        }
      }
      if (elect_one()) umma_commit(dx_full);    // (elect.sync picks the same lane for the same mask: all MMAs are its own)
      __syncwarp();
      if (lane == 0) NNCF_STAMP(3);
    }
  } else if (warp < 2 + kScoreEpiWarps) {
    // ------------------------------------------------------------------------------ epilogue warps
    const int ew = warp - 2;
    const int q = warp & 3;                     // TMEM lane quadrant this warp may access
    const int h = ew >> 2;                      // column half of each S' tile (CW columns)
    const int ol = q * 32 + lane;               // row inside the owned block
    const int o = ob * 128 + ol;                // owner-side index (i on side 0, j on side 1)
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    const int nc = ncols > 1 ? ncols : 2;
    EpiConst ec;
    ec.w_neg = a.lambda / static_cast<float>(nc - 1);
    ec.gamma = a.gamma;
    ec.inv_b = 1.0f / static_cast<float>(a.B);
    ec.inv_cnt = 1.0f / (static_cast<float>(a.B) * static_cast<float>(nc));
    ec.ns_margin = GROUP ? 0.0f : a.gamma;
    ec.inv_wneg = 1.0f / ec.w_neg;
    ec.g_scale = (LOSS == NNCF_LOSS_SKIP_GRAM) ? ec.w_neg * ec.inv_b : ((LOSS == NNCF_LOSS_MAX_MARGIN) ? ec.inv_cnt : 1.0f);
    if (a.fuse_sgd) ec.g_scale *= a.neg_lr;     // the drain adds -lr * dX straight into the table rows
    const bool row_ok = o < n_owner;
    // side 0: the positive column of my row.  side 1 (neg_shared): my own index (the diagonal).
    const int my_posc = (side == 0 && GROUP) ? (row_ok ? a.inverse[base + o] : -1) : o;
    const bool sp_mine = kPairwise && ((side == 0) == GROUP);
    const float my_sp = (sp_mine && row_ok) ? a.spos[base + o] : 0.0f;
    const int32_t* inv_row = GROUP ? a.inverse + base : nullptr;
    const float* spos_row = a.spos + base;
    float lsum = 0.0f, asum = 0.0f, lraw = 0.0f;

    if (a.self_gather) {
      // gather my CTA's 128 rows: this warp takes rows 16 ew .. 16 ew + 15 (ids first: the id -> row dependency is two
      // L2 round trips; the rows were pulled into L2 by the previous step's prefetch warps)
      const int32_t* gids = (side == 0 ? a.ids_u + r * a.ids_stride_u : a.ids_v + r * a.ids_stride_v) + ob * 128 + ew * 16;
      const float* gtab = side == 0 ? a.table_u : a.table_v;
      const int rows_here = n_owner - (ob * 128 + ew * 16);                     // valid rows of my 16
      int64_t myid = -1;
      if (lane < 16 && lane < rows_here) myid = gids[lane];
      const int c = 4 * lane;                                                   // my 4 columns of every row
      uint8_t* gimg = const_cast<uint8_t*>(gX);
#pragma unroll 1
      for (int k0 = 0; k0 < 16; k0 += 8) {                                      // 8 row loads in flight (the register budget is 80)
        float4 x[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int64_t id = __shfl_sync(0xffffffffu, myid, k0 + k);
          x[k] = (id >= 0 && c < a.d) ? __ldg(reinterpret_cast<const float4*>(gtab + id * a.d + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        if (c < DP) {
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            uint2 pk2;
            pk2.x = pack_bf16x2(x[k].x, x[k].y);
            pk2.y = pack_bf16x2(x[k].z, x[k].w);
            const uint32_t off = (c >> 6) * kSubBytes + sw128_offset(ew * 16 + k0 + k, c & 63);
            *reinterpret_cast<uint2*>(sX + off) = pk2;                           // my X operand
            *reinterpret_cast<uint2*>(gimg + off) = pk2;                         // the other side's Y tiles
          }
        }
      }
      fence_proxy_async();                                                      // sX is read by the tensor core (async proxy)
      asm volatile("bar.sync 1, %0;" ::"n"(32 * kScoreEpiWarps) : "memory");    // epilogue warps only; orders their image
                                                                                // stores before the (cumulative) release below
      if (tid == 64) {
        st_release_gpu(a.gather_flags + (r * 2 + side) * nblk + ob, a.gather_seq);
        atomicAdd(a.gather_count, 1ull);
        mbar_arrive(x_full);
      }
    }

    // teams: this warp's group of four takes tiles h, h + 2, ...; otherwise every warp takes its column half of every tile
    constexpr int kItStep = C::kTeams ? 2 : 1;
    const int hcol = C::kTeams ? 0 : h * CW;            // my first column inside a tile
    for (int it = C::kTeams ? h : 0; it < nt_loop; it += kItStep) {
      const int t = GX ? gx_tile(it) : t_begin + it;    // tile position in the sweep; `it` indexes the pipeline state
      const int sb = it % kBufs;
      mbar_wait(&s_full[sb], (it / kBufs) & 1);
      if (warp == 2 && lane == 0 && it < 8) NNCF_STAMP(16 + it);
      tc_fence_after();
      float v[NV][32];
#pragma unroll
      for (int j = 0; j < NV; ++j) tmem_ld32(tmem + lane_addr + sb * TN + hcol + 32 * j, v[j]);
      tmem_ld_wait();
      const int x0 = t * TN + hcol;                     // swept index of my first column
      // fast path: a full tile that cannot contain a positive (neg_shared: only the diagonal tiles have them)
      constexpr bool kFastSg = (LOSS == NNCF_LOSS_SKIP_GRAM);
      constexpr bool kBalanceLoss = kFastSg && !GROUP;
      const bool on_diag = ((t * TN) >> 7) == ob;       // this tile crosses the diagonal of my 128-row block
      const bool full = (t * TN + TN <= n_other);
      // neg_shared skip-gram, full diagonal tile: it holds at most ONE positive per row (column == row).  Run the packed
      // fast path on all scores and patch that single element afterwards instead of the per-element general path.
      const bool diag_fast = kFastSg && !GROUP && on_diag && full;
      const bool general = (GROUP || on_diag || !full) && !diag_fast;
      uint32_t pk[NV][16];
      bool do_patch = false;
      uint16_t patch_bits = 0;
      const int pidx = o - x0;                          // my positive's column inside this warp's CW columns (if any)
      if (diag_fast && pidx >= 0 && pidx < CW) {
        float sp = 0.0f;
#pragma unroll
        for (int j = 0; j < NV; ++j)
#pragma unroll
          for (int u = 0; u < 32; ++u) sp = (pidx == 32 * j + u) ? v[j][u] : sp;
        const float sa = fmaf(0.5f, tanh_approx(0.5f * fabsf(sp)), 0.5f);
        const float sg = (sp >= 0.0f) ? sa : 1.0f - sa;
        const __nv_bfloat16 pb = __float2bfloat16(ec.inv_wneg * (sg - 1.0f));      // G' of the positive
        patch_bits = *reinterpret_cast<const uint16_t*>(&pb);
        do_patch = row_ok;
        if (side == 0 && row_ok) {
          const float sps = fmaxf(sp, 0.0f) - __logf(sa);                           // softplus(s)
          lsum += ec.inv_b * (sps - sp) - ec.w_neg * ec.inv_b * sps;                // positive's own term minus what the fast path adds
        }
      }
      // who adds the loss of a fast-path tile: the loss code triples the fast path (279 vs 98 instructions per 32 scores), so
      // for neg_shared skip-gram the two sides share it by the parity of the TN x TN block the scores sit in (every block is
      // seen once by each side); blocks that touch a ragged edge stay with side 0, whose general path handles them
      bool loss_here = (side == 0);
      if (NNCF_ABLATE & 16) loss_here = false;
      else if (GX) loss_here = true;                    // every block is computed once, by one side: that side adds its loss
      else if (kBalanceLoss) {
        const bool take1 = (((o / TN) + t) & 1) && ((o / TN) * TN + TN <= n_owner);   // parity-1 block, my row block is full
        loss_here = (side == 0) ? !take1 : take1;
      }
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        if (NNCF_ABLATE & 1) {
#pragma unroll
          for (int i = 0; i < 16; ++i) pk[j][i] = pack_bf16x2(v[j][2 * i], v[j][2 * i + 1]);
        } else if (general) {
          if (side == 0) epi_chunk<LOSS, false, GROUP, true>(ec, v[j], x0 + 32 * j, n_other, row_ok, o, my_posc, my_sp, inv_row, spos_row, lsum, asum, pk[j]);
          else epi_chunk<LOSS, true, GROUP, true>(ec, v[j], x0 + 32 * j, n_other, row_ok, o, my_posc, my_sp, inv_row, spos_row, lsum, asum, pk[j]);
        } else if (kFastSg) {
          if (loss_here) epi_chunk_sg_fast<true>(v[j], pk[j], lraw);
          else epi_chunk_sg_fast<false>(v[j], pk[j], lraw);
        } else if (LOSS == NNCF_LOSS_MAX_MARGIN) {
          if (side == 0) epi_chunk_mm_fast<false, GROUP>(ec, v[j], x0 + 32 * j, row_ok, my_sp, spos_row, lsum, asum, pk[j]);
          else epi_chunk_mm_fast<true, GROUP>(ec, v[j], x0 + 32 * j, row_ok, my_sp, spos_row, lsum, asum, pk[j]);
        } else {
          // (pointwise losses: the two variants differ only in who adds the loss; under the G' exchange whoever computes a block does)
          if (side == 0 || GX) epi_chunk<LOSS, false, GROUP, false>(ec, v[j], x0 + 32 * j, n_other, row_ok, o, my_posc, my_sp, inv_row, spos_row, lsum, asum, pk[j]);
          else epi_chunk<LOSS, true, GROUP, false>(ec, v[j], x0 + 32 * j, n_other, row_ok, o, my_posc, my_sp, inv_row, spos_row, lsum, asum, pk[j]);
        }
      }
      if (do_patch) {
        // branch-free: every lane of a patch warp has a DIFFERENT pidx (== its lane within the chunk), so an if-chain
        // here diverges 32 ways (measured: +4k cycles on each diagonal tile); selects cost 3 instructions per word
        const bool odd = (pidx & 1) != 0;
        const uint32_t keep = odd ? 0x0000FFFFu : 0xFFFF0000u;
        const uint32_t ins = odd ? (static_cast<uint32_t>(patch_bits) << 16) : static_cast<uint32_t>(patch_bits);
        const int pw = pidx >> 1;
#pragma unroll
        for (int j = 0; j < NV; ++j)
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const uint32_t patched = (pk[j][i] & keep) | ins;
            pk[j][i] = (pw == 16 * j + i) ? patched : pk[j][i];
          }
      }
      if (GX) {
        // G' exchange: my 32 (= CW) item columns of row ol go into the image the item-side CTA of this item block copies
        // as an MN-major A operand: row = user row, 16-byte chunk c of the 128-byte row at position c ^ (row & 7)
        const int tb = t * TN;
        uint8_t* gdst = a.gx_buf + (((((int64_t)r * 2 + side) * nblk + ob) * nblk + (tb >> 7)) * 2 + ((tb >> 6) & 1)) * kSubBytes + ol * 128;
#pragma unroll
        for (int j = 0; j < NV; ++j)
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int c = (((tb + hcol + 32 * j) & 63) >> 3) + i;
            *reinterpret_cast<uint4*>(gdst + ((c ^ (ol & 7)) << 4)) = make_uint4(pk[j][4 * i], pk[j][4 * i + 1], pk[j][4 * i + 2], pk[j][4 * i + 3]);
          }
      }
      // G' goes back into TMEM over the first half of my own S' columns (two bf16 per column): the A operand of MMA2
#pragma unroll
      for (int j = 0; j < NV; ++j) tmem_st16(tmem + lane_addr + sb * TN + hcol + 16 * j, pk[j]);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&g_full[sb]);
        if (GX) red_release_cta_smem_add(gx_cnt, 1u);   // (after the __syncwarp: covers the whole warp's G' stores)
      }
      if (warp == 2 && lane == 0 && it < 8) NNCF_STAMP(24 + it);
    }
    // (the loss hand-off comes BEFORE the drain: its values are final once the tile loop has ended, and its two gpu-scope
    //  fences would otherwise sit behind ~20 MB of bulk reductions: 19 % of this kernel's stall samples were there)
    // per-row sums of dL/dD: the two column halves of a row live in two warps -> atomics on two addends only
    if (kPairwise && row_ok) {
      if (side == 0 && GROUP) atomicAdd(a.corrU + base + o, asum);
      if (side == 1 && !GROUP) atomicAdd(a.corrV + base + o, asum);
    }
    if (side == 0 || bal_loss) {   // (balanced loss: both sides hold a share)
      // The replica's loss is summed with a gpu-scope atomic, a fence and an arrival counter (the last arriver publishes):
      // ~4 us when the epilogue warps did it themselves in front of (or behind) the drain - the in-kernel timeline showed
      // the drain + update is only 2.6 us.  The warps now leave their sums in shared memory, arrive on a named barrier
      // without waiting, and the idle producer warp does the hand-off while they drain.
      if (row_ok) lsum += ec.w_neg * ec.inv_b * lraw;      // fast-path elements: all negatives, weight w_neg / B
      if (a.reg_scale != 0.0f && side == 0 && t_begin == 0 && row_ok) {
        // activity regulariser, loss term: u_reg * sum_d mean_b x[b,d]^2 = (reg_scale / 2) * sum over my half row of x^2
        // (ref: utils/utilities.py:129-135), x from the bf16 X block; its gradient is added at the drain
        float ss = 0.0f;
#pragma unroll
        for (int q8 = 0; q8 < DP / 16; ++q8) {
          const int c = h * (DP / 2) + 8 * q8;
          const uint4 w = *reinterpret_cast<const uint4*>(sX + (c >> 6) * kSubBytes + sw128_offset(ol, c & 63));
          const uint32_t ws[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float lo = __uint_as_float(ws[e] << 16), hi = __uint_as_float(ws[e] & 0xFFFF0000u);
            ss = fmaf(lo, lo, fmaf(hi, hi, ss));
          }
        }
        lsum = fmaf(0.5f * a.reg_scale, ss, lsum);
      }
      lsum = warp_sum(lsum);
      if (lane == 0) loss_slots[ew] = lsum;
      __syncwarp();
      asm volatile("bar.arrive 2, %0;" ::"n"(32 * kScoreEpiWarps + 32) : "memory");
    }
    // drain the accumulated gradient of the owned rows: this warp takes columns [h*DP/2, (h+1)*DP/2)
    mbar_wait(dx_full, 0);
    if (warp == 2 && lane == 0) NNCF_STAMP(4);
    if (a.tl && warp == 2 && lane == 0) atomicMax(&a.tl[6], global_timer_ns());   // last CTA leaves its tile loop
    tc_fence_after();
    {
      // TMEM rows live in lanes, so a direct store would scatter 32 rows per instruction (measured: 4.4k cycles).
      // Transpose through shared memory (the Y/G buffers are idle now) and write whole rows, coalesced.
      constexpr int LD = DP + 4;                                  // padded row: STS.128 at the 4-wavefront minimum
      float* stage = reinterpret_cast<float*>(sY);               // the Y stages are idle once dx_full has fired; X stays intact
      // activity regulariser (user side): + reg_scale * x, x from the bf16 X block (4 LDS.128 per 32 columns of my row)
      const bool with_reg = (a.reg_scale != 0.0f) && side == 0 && t_begin == 0;     // (ONE of the split CTAs adds it: the one that sweeps tile 0)
      const float reg_s = with_reg ? a.reg_scale * (a.fuse_sgd ? a.neg_lr : 1.0f) : 0.0f;
#pragma unroll 1
      for (int c0 = h * (DP / 2); c0 < (h + 1) * (DP / 2); c0 += 32) {
        float v[32];
        tmem_ld32(tmem + lane_addr + C::kColDX + c0, v);
        tmem_ld_wait();
#pragma unroll
        for (int u = 0; u < 32; ++u) v[u] *= ec.g_scale;
        if (with_reg) {
#pragma unroll
          for (int q8 = 0; q8 < 4; ++q8) {
            const int c = c0 + 8 * q8;
            const uint4 w = *reinterpret_cast<const uint4*>(sX + (c >> 6) * kSubBytes + sw128_offset(ol, c & 63));
            const uint32_t ws[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              v[8 * q8 + 2 * e] = fmaf(reg_s, __uint_as_float(ws[e] << 16), v[8 * q8 + 2 * e]);
              v[8 * q8 + 2 * e + 1] = fmaf(reg_s, __uint_as_float(ws[e] & 0xFFFF0000u), v[8 * q8 + 2 * e + 1]);
            }
          }
        }
#pragma unroll
        for (int u = 0; u < 32; u += 4)
          *reinterpret_cast<float4*>(stage + ol * LD + c0 + u) = make_float4(v[u], v[u + 1], v[u + 2], v[u + 3]);
      }
      if (a.fuse_sgd) fence_proxy_async();                                     // staged rows are read by the TMA engine
      if (a.self_gather && tid == 64) {
        // nobody may still be READING the tables when the first update lands (all CTAs of the grid are co-resident and
        // their gathers ended ~20 us ago: this never spins in practice, it makes the one-snapshot semantics unconditional)
        uint32_t spins = 0;
        while (ld_acquire_gpu_u64(a.gather_count) < a.gather_target) { __nanosleep(64); if (++spins > 20000000u) __trap(); }
      }
      asm volatile("bar.sync 1, %0;" ::"n"(32 * kScoreEpiWarps) : "memory");   // epilogue warps only
      if (a.fuse_sgd) {
        // (the staged rows already carry the factor -lr, see g_scale)
        const ShardPtrs& sh = side == 0 ? a.shards_u : a.shards_v;
        float* table = side == 0 ? a.table_u : a.table_v;
        const int32_t* ids = (side == 0 ? a.ids_u + r * a.ids_stride_u : a.ids_v + r * a.ids_stride_v) + ob * 128;
        if (!a.drain_vec) {
          // one bulk reduction (TMA engine, performed at the L2) per owned row
          const int row = tid - 64;                                  // epilogue threads 0..255: the first 128 take a row each
          int64_t id = (row < 128 && ob * 128 + row < n_owner) ? ids[row] : -1;
          if (((a.dedup >> side) & 1) && row < 128) {
            // Crowded ids (a stratified block at N = 8 spreads 18,944 links over 62,500 item rows: its hottest row takes
            // ~10 % of them) make thousands of 512-byte reductions queue on the same L2 lines: the score kernel went from
            // 21.9 us (N = 1) to 29.6 us (N = 8).  Rows of a warp's 32 that share an id are summed in shared memory first
            // and leave as ONE reduction: one MATCH per warp, no block barrier, nothing to do when there are no duplicates.
            // (Folding over the whole 128-row block with two block barriers cost 3.5 us per step and gained 2.6 at N = 8.)
            const unsigned same = __match_any_sync(0xffffffffu, static_cast<int32_t>(id));
            const int leader = __ffs(same) - 1;
            unsigned dup = __ballot_sync(0xffffffffu, id >= 0 && leader != lane);
            while (dup) {                                                          // warp-cooperative: lanes along the row
              const int L = __ffs(dup) - 1;
              dup &= dup - 1;
              const int src = (row & ~31) + L, dst = (row & ~31) + __shfl_sync(0xffffffffu, leader, L);
              for (int c = 4 * lane; c < a.d; c += 128) {
                const float4 g4 = *reinterpret_cast<const float4*>(stage + src * LD + c);
                float4* q = reinterpret_cast<float4*>(stage + dst * LD + c);
                float4 t4 = *q;
                t4.x += g4.x; t4.y += g4.y; t4.z += g4.z; t4.w += g4.w;
                *q = t4;
              }
              __syncwarp();
            }
            if (id >= 0 && leader != lane) id = -1;                                // folded into the warp's first occurrence
            fence_proxy_async();                                                   // the sums are read by the TMA engine
            __syncwarp();
          }
          if (!(NNCF_ABLATE & 32) && id >= 0) {
            float* trow = sh.n > 1 ? sh.p[id % sh.n] + (id / sh.n) * a.d : table + id * a.d;
            bulk_reduce_add_f32_s2g(trow, stage + row * LD, static_cast<uint32_t>(a.d) * 4u);
            bulk_commit_group();
            bulk_wait_group_read0();                                 // shared memory must outlive the engine's reads
          }
        } else {
          // alternative (NNCF_DRAIN_VEC=1): vector reductions from the staged rows, a warp instruction covers one whole row
          // (coalesced 128 B lines), fire and forget.  Measured: no faster than the bulk form at R = 37 (the L2 reduction
          // throughput bounds both) and slower at R = 1 (10.0 vs 8.9 us per step)
          const int nrow = min(128, n_owner - ob * 128);
          int64_t myid = 0;
          if (lane < 16 && ew + 8 * lane < nrow) myid = ids[ew + 8 * lane];
#pragma unroll 4
          for (int k = 0; k < 16; ++k) {
            const int row = ew + 8 * k;
            const int64_t id = __shfl_sync(0xffffffffu, myid, k);
            if (row >= nrow) break;
            float* trow = sh.n > 1 ? sh.p[id % sh.n] + (id / sh.n) * a.d : table + id * a.d;
            for (int c = 4 * lane; c < a.d; c += 128) {
              const float4 g4 = *reinterpret_cast<const float4*>(stage + row * LD + c);
              red_add_v4(trow + c, g4.x, g4.y, g4.z, g4.w);
            }
          }
        }
      } else {
        float4* dst = reinterpret_cast<float4*>((side == 0 ? a.dU : a.dV) + (base + (int64_t)ob * 128) * DP);
        for (int row = ew; row < 128; row += kScoreEpiWarps) {
          if (ob * 128 + row >= n_owner) break;
          const float4* src = reinterpret_cast<const float4*>(stage + row * LD);
          if (a.split > 1) {            // partial blocks of the split CTAs sum in the (zeroed) gradient block
            for (int c = lane; c < DP / 4; c += 32) { const float4 g4 = src[c]; red_add_v4(reinterpret_cast<float*>(dst + row * (DP / 4) + c), g4.x, g4.y, g4.z, g4.w); }
          } else {
            for (int c = lane; c < DP / 4; c += 32) dst[row * (DP / 4) + c] = src[c];
          }
        }
      }
    }
    if (warp == 2 && lane == 0) NNCF_STAMP(5);
    tc_fence_before();
  }
  else {
   if (a.next_ids_u && !(NNCF_ABLATE & 64)) {
    // ------------------------------------------------------------------------------ spare warps: L2 prefetch
    const int cta = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
    const int ncta = gridDim.x * gridDim.y * gridDim.z;
    const bool vside = warp == 2 + kScoreEpiWarps + 1;
    const int32_t* ids = vside ? a.next_ids_v : a.next_ids_u;
    const float* table = vside ? a.pf_table_v : a.pf_table_u;
    const int64_t nrows = vside ? a.n_rows_v : a.n_rows_u;
    const uint32_t row_bytes = static_cast<uint32_t>(a.d) * 4u;
    for (int i = cta * 32 + lane; i < a.next_count; i += ncta * 32) {
      int64_t id = ids[i];
      id = id < 0 ? 0 : (id >= nrows ? nrows - 1 : id);
      asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(table + id * a.d), "r"(row_bytes) : "memory");
    }
   }
   if (a.pf_now_count > 0 && !(NNCF_ABLATE & 64)) {
    const int cta = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
    const int ncta = gridDim.x * gridDim.y * gridDim.z;
    const int vside = warp == 2 + kScoreEpiWarps + 1 ? 1 : 0;
    const int32_t* ids = vside ? a.ids_v : a.ids_u;
    const int64_t stride = vside ? a.ids_stride_v : a.ids_stride_u;
    const uint32_t row_bytes = static_cast<uint32_t>(a.d) * 4u;
    for (int i = cta * 32 + lane; i < a.pf_now_count; i += ncta * 32) {
      const int64_t id = ids[(i / a.B) * stride + i % a.B];
#pragma unroll
      for (int k = 0; k < 3; ++k)
        if (a.pf_now[vside][k])
          asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a.pf_now[vside][k] + id * a.d), "r"(row_bytes) : "memory");
    }
   }
   if (GX && warp == 2 + kScoreEpiWarps && lane == 0) {
    // G' exchange: publish every item block of my G' rows once all epilogue warps have stored both of its tiles.  The
    // warps' stores are ordered before their release-add on the counter, my acquire-load of the counter before the
    // gpu-scope release store of the flag (cumulative), so whoever acquires the flag sees the image.
    for (int j = 0; j < nblk / 2; ++j) {
      const uint32_t need = static_cast<uint32_t>(kScoreEpiWarps * kTilesPerBlk * (j + 1));
      uint32_t spins = 0;
      while (ld_acquire_cta_smem_u32(gx_cnt) < need) { __nanosleep(32); if (++spins > 40000000u) __trap(); }
      st_release_gpu(a.gx_flags + (((int64_t)r * 2 + side) * nblk + ob) * nblk + (ob + side + 2 * j) % nblk, a.gx_seq);
    }
   }
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, C::kTmemCols);
  }
  if (tid == 0) NNCF_STAMP(6);
  if (tid == 0 && dbg) dbg[41] = static_cast<long long>(global_timer_ns());
  if (a.tl && tid == 0) {
    const unsigned long long g1 = global_timer_ns();
    atomicMax(&a.tl[5], g1);
    if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) { a.tl[9] = static_cast<unsigned long long>(clock64() - tl_c0); a.tl[10] = g1 - tl_g0; }
  }
#undef NNCF_STAMP
}

// host-side dispatch over <LOSS, GROUP> for one NSUB (instantiated by score_tc_nsub*.cu)
template <int NSUB, int LOSS, bool GROUP, bool GX = false>
int launch_score_tc_one(const ScoreTcArgs& a, int nblk, int R, cudaStream_t st) {
  using C = ScoreTcCfg<NSUB>;
  static bool attr_set = false;
  if (!attr_set) {
    NNCF_CUDA(cudaFuncSetAttribute(score_grad_tc_kernel<NSUB, LOSS, GROUP, GX>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)C::kSmemBytes));
    // two resident CTAs per SM need the full shared-memory carve-out
    NNCF_CUDA(cudaFuncSetAttribute(score_grad_tc_kernel<NSUB, LOSS, GROUP, GX>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                   (int)cudaSharedmemCarveoutMaxShared));
    attr_set = true;
  }
  NNCF_CUDA(launch_pdl(score_grad_tc_kernel<NSUB, LOSS, GROUP, GX>, dim3(nblk * a.split, 2, R), dim3(kScoreThreads), C::kSmemBytes, st, a));
  count_launch();
  return 0;
}

// G' exchange instantiations exist for dp <= 128 and the pointwise losses of neg_shared (the host asks for nothing else)
template <int NSUB, int LOSS>
int launch_score_tc_gx(const ScoreTcArgs& a, int nblk, int R, cudaStream_t st) {
  if constexpr (NSUB <= 2 && !ScoreTcCfg<NSUB>::kTeams) return launch_score_tc_one<NSUB, LOSS, false, true>(a, nblk, R, st);
  else { (void)a; (void)nblk; (void)R; (void)st; ::nncf::set_error("G' exchange: dp <= 128 only"); return NNCF_EUNSUPPORTED; }
}

template <int NSUB>
int launch_score_tc_all(const ScoreTcArgs& a, int nblk, int R, cudaStream_t st) {
  const bool g = a.scheme == NNCF_SCHEME_GROUP_NEG_SHARED;
  if (a.gx) return a.loss_kind == NNCF_LOSS_SKIP_GRAM ? launch_score_tc_gx<NSUB, 0>(a, nblk, R, st) : launch_score_tc_gx<NSUB, 1>(a, nblk, R, st);
  switch (a.loss_kind) {
    case NNCF_LOSS_SKIP_GRAM: return g ? launch_score_tc_one<NSUB, 0, true>(a, nblk, R, st) : launch_score_tc_one<NSUB, 0, false>(a, nblk, R, st);
    case NNCF_LOSS_MSE: return g ? launch_score_tc_one<NSUB, 1, true>(a, nblk, R, st) : launch_score_tc_one<NSUB, 1, false>(a, nblk, R, st);
    case NNCF_LOSS_LOG_LOSS: return g ? launch_score_tc_one<NSUB, 2, true>(a, nblk, R, st) : launch_score_tc_one<NSUB, 2, false>(a, nblk, R, st);
    default: return g ? launch_score_tc_one<NSUB, 3, true>(a, nblk, R, st) : launch_score_tc_one<NSUB, 3, false>(a, nblk, R, st);
  }
}

}  // namespace nncf
