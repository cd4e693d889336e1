// batch_builder.cu — device batch-index builders.
//
//   nncf_permute_rows         np.random.shuffle(train)          ref: models/train_neg_shared.py:37
//   nncf_group_shuffle        group_shuffle_train(...)          ref: configs/data_utils.py:218-241
//   nncf_assemble_pairs_batch positives + k sampled negatives   ref: models/train_original.py:49-56,
//                                                                    models/train_group_sample.py:75-85
// The random permutations are drawn by the host from the shared stream (same order as the reference); the
// device applies them and does the STABLE sort by key (LSD radix sort, 8 bits per pass, stable block-local
// ranking) so that the result is bit-identical to the reference function run with a stable argsort.
#include <cstring>
#include "common.cuh"

namespace nncf {

__global__ void __launch_bounds__(256)
permute_rows_kernel(const int32_t* __restrict__ train, int64_t n, const int64_t* __restrict__ perm, int32_t* __restrict__ out) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const int64_t s = perm[r];
  out[r * 3 + 0] = train[s * 3 + 0];
  out[r * 3 + 1] = train[s * 3 + 1];
  out[r * 3 + 2] = train[s * 3 + 2];
}

__global__ void __launch_bounds__(256)
make_keys_kernel(const int32_t* __restrict__ train, int64_t n, int col, const int64_t* __restrict__ iidx,
                 const int64_t* __restrict__ row_perm, uint32_t* __restrict__ keys, uint32_t* __restrict__ vals) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const int64_t src = row_perm[r];
  keys[r] = static_cast<uint32_t>(iidx[train[src * 3 + col]]);
  vals[r] = static_cast<uint32_t>(src);
}

constexpr int kRadixThreads = 256;
constexpr int kRadixItems = 16;                                  // rounds per block
constexpr int kRadixChunk = kRadixThreads * kRadixItems;         // 4096 elements per block

__global__ void __launch_bounds__(kRadixThreads)
radix_hist_kernel(const uint32_t* __restrict__ keys, int64_t n, int shift, int nblocks, uint32_t* __restrict__ hist) {
  __shared__ uint32_t h[256];
  h[threadIdx.x] = 0;
  __syncthreads();
  const int64_t base = (int64_t)blockIdx.x * kRadixChunk;
  for (int r = 0; r < kRadixItems; ++r) {
    const int64_t i = base + (int64_t)r * kRadixThreads + threadIdx.x;
    if (i < n) atomicAdd(&h[(keys[i] >> shift) & 255u], 1u);
  }
  __syncthreads();
  hist[(int64_t)threadIdx.x * nblocks + blockIdx.x] = h[threadIdx.x];   // digit-major
}

// exclusive scan of a uint32 array, three kernels (block sums -> scan of sums -> add back)
constexpr int kScanBlock = 1024;
__global__ void __launch_bounds__(kScanBlock)
scan_block_kernel(uint32_t* __restrict__ data, int64_t n, uint32_t* __restrict__ block_sums) {
  __shared__ uint32_t warp_tot[32];
  const int64_t i = (int64_t)blockIdx.x * kScanBlock + threadIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t v = (i < n) ? data[i] : 0u;
  uint32_t x = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) warp_tot[warp] = x;
  __syncthreads();
  if (warp == 0) {
    uint32_t t = warp_tot[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t y = __shfl_up_sync(0xffffffffu, t, o);
      if (lane >= o) t += y;
    }
    warp_tot[lane] = t;
  }
  __syncthreads();
  const uint32_t off = (warp == 0) ? 0u : warp_tot[warp - 1];
  if (i < n) data[i] = off + x - v;
  if (threadIdx.x == kScanBlock - 1) block_sums[blockIdx.x] = off + x;
}
__global__ void __launch_bounds__(kScanBlock)
scan_sums_kernel(uint32_t* __restrict__ sums, int n) {   // single block, serial over chunks
  __shared__ uint32_t warp_tot[32];
  __shared__ uint32_t carry;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < n; base += kScanBlock) {
    const int i = base + threadIdx.x;
    const uint32_t v = (i < n) ? sums[i] : 0u;
    uint32_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) warp_tot[warp] = x;
    __syncthreads();
    if (warp == 0) {
      uint32_t t = warp_tot[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, t, o);
        if (lane >= o) t += y;
      }
      warp_tot[lane] = t;
    }
    __syncthreads();
    const uint32_t off = carry + ((warp == 0) ? 0u : warp_tot[warp - 1]);
    if (i < n) sums[i] = off + x - v;
    __syncthreads();
    if (threadIdx.x == kScanBlock - 1) carry = off + x;
    __syncthreads();
  }
}
__global__ void __launch_bounds__(kScanBlock)
scan_add_kernel(uint32_t* __restrict__ data, int64_t n, const uint32_t* __restrict__ block_sums) {
  const int64_t i = (int64_t)blockIdx.x * kScanBlock + threadIdx.x;
  if (i < n) data[i] += block_sums[blockIdx.x];
}

// stable scatter: elements are ranked in (round, thread) order inside the block's chunk
__global__ void __launch_bounds__(kRadixThreads)
radix_scatter_kernel(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in, int64_t n, int shift,
                     int nblocks, const uint32_t* __restrict__ offsets, uint32_t* __restrict__ keys_out,
                     uint32_t* __restrict__ vals_out) {
  __shared__ uint32_t running[256];
  __shared__ uint32_t warp_cnt[kRadixThreads / 32][256];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  running[tid] = offsets[(int64_t)tid * nblocks + blockIdx.x];
  for (int w = 0; w < kRadixThreads / 32; ++w) warp_cnt[w][tid] = 0;
  __syncthreads();
  const int64_t base = (int64_t)blockIdx.x * kRadixChunk;
  for (int r = 0; r < kRadixItems; ++r) {
    const int64_t i = base + (int64_t)r * kRadixThreads + tid;
    const bool ok = i < n;
    uint32_t key = 0, val = 0, dg = 0;
    if (ok) { key = keys_in[i]; val = vals_in[i]; dg = (key >> shift) & 255u; }
    // invalid lanes get a digit value outside 0..255 so that they only match each other
    const uint32_t peers = __match_any_sync(0xffffffffu, ok ? dg : 0xffffu);
    const uint32_t rank_in_warp = __popc(peers & ((1u << lane) - 1u));
    if (ok && rank_in_warp == 0) warp_cnt[warp][dg] = __popc(peers);
    __syncthreads();
    if (ok) {
      uint32_t pos = running[dg] + rank_in_warp;
      for (int w = 0; w < warp; ++w) pos += warp_cnt[w][dg];
      keys_out[pos] = key;
      vals_out[pos] = val;
    }
    __syncthreads();
    {
      uint32_t tot = 0;
      for (int w = 0; w < kRadixThreads / 32; ++w) { tot += warp_cnt[w][tid]; warp_cnt[w][tid] = 0; }
      running[tid] += tot;
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256)
group_emit_kernel(const int32_t* __restrict__ train, int64_t n, const uint32_t* __restrict__ sorted_src,
                  const int64_t* __restrict__ block_perm, int chop, int32_t* __restrict__ out) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  int64_t q = p;
  if (chop > 0) {
    const int64_t bulk = (n / chop) * chop;
    if (p < bulk) q = block_perm[p / chop] * chop + (p % chop);
  }
  const int64_t s = sorted_src[q];
  out[p * 3 + 0] = train[s * 3 + 0];
  out[p * 3 + 1] = train[s * 3 + 1];
  out[p * 3 + 2] = train[s * 3 + 2];
}

__global__ void __launch_bounds__(256)
assemble_pairs_kernel(const int32_t* __restrict__ pos, int B, int k, const int32_t* __restrict__ negs, int neg_col,
                      int neg_sign, int32_t* __restrict__ out) {
  const int n = (1 + k) * B;
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  int32_t row[3];
  if (r < B) {
    row[0] = pos[r * 3]; row[1] = pos[r * 3 + 1]; row[2] = pos[r * 3 + 2];
  } else {
    const int p = (r - B) / k;
    row[0] = pos[p * 3]; row[1] = pos[p * 3 + 1];
    row[neg_col] = negs[r - B];
    row[2] = neg_sign;
  }
  out[r * 3] = row[0]; out[r * 3 + 1] = row[1]; out[r * 3 + 2] = row[2];
}

// presample: the whole epoch's links with their k sampled negatives in one array.
//   layout 0 (shuffle_st 'original' / 'reverse'): row p(1+k) is positive p, the k rows after it are its negatives
//            (train_p.repeat(1+k); column neg_col <- samples; train[::1+k] = train_p)   ref: models/train_presample.py:46-60
//   layout 1 (all other shuffle_st): the N positives, then positive p's negatives at N + p k + j
//            (np.vstack((train_p, train_n)))                                            ref: models/train_presample.py:61-65
__global__ void __launch_bounds__(256)
presample_assemble_kernel(const int32_t* __restrict__ pos, int64_t n, int k, const int32_t* __restrict__ negs, int neg_col,
                          int neg_sign, int layout, int32_t* __restrict__ out) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n * (1 + k)) return;
  int64_t p, j;                                     // positive index, negative slot (-1 = the positive itself)
  if (layout == 0) { p = r / (1 + k); j = r % (1 + k) - 1; }
  else if (r < n) { p = r; j = -1; }
  else { p = (r - n) / k; j = (r - n) % k; }
  int32_t row[3] = {pos[p * 3], pos[p * 3 + 1], pos[p * 3 + 2]};
  if (j >= 0) { row[neg_col] = negs[p * k + j]; row[2] = neg_sign; }
  out[r * 3] = row[0]; out[r * 3 + 1] = row[1]; out[r * 3 + 2] = row[2];
}

// sampled_neg_shared: batch b = its B positive rows followed by k rows (user 0, item negs[b k + j])
//   ref: models/train_sampled_neg_shared.py:28,46-49 (train_batch_n = np.zeros((k, 3)); column 1 <- sample_batch(k))
__global__ void __launch_bounds__(256)
sns_assemble_kernel(const int32_t* __restrict__ train, int64_t n_batches, int B, int k, const int32_t* __restrict__ negs,
                    int32_t* __restrict__ uid, int32_t* __restrict__ cid) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int rows = B + k;
  if (r >= n_batches * rows) return;
  const int64_t b = r / rows;
  const int i = static_cast<int>(r - b * rows);
  if (i < B) { uid[r] = train[(b * B + i) * 3]; cid[r] = train[(b * B + i) * 3 + 1]; }
  else { uid[r] = 0; cid[r] = negs[b * k + (i - B)]; }
}

}  // namespace nncf

using namespace nncf;

extern "C" int nncf_permute_rows(const int32_t* train_dev, int64_t n_rows, const int64_t* row_perm_dev, int32_t* out_dev,
                                 void* stream) {
  NNCF_CHECK_ARG(n_rows >= 0, "nncf_permute_rows: n_rows < 0");
  if (n_rows == 0) return NNCF_OK;
  NNCF_CHECK_ARG(train_dev && row_perm_dev && out_dev, "nncf_permute_rows: null argument");
  permute_rows_kernel<<<ceil_div(n_rows, 256), 256, 0, (cudaStream_t)stream>>>(train_dev, n_rows, row_perm_dev, out_dev);
  NNCF_LAUNCH_OK();
  return NNCF_OK;
}

struct ShufflePlan {
  int nblocks;
  int64_t hist_n;
  int scan_blocks;
  size_t off_keys[2], off_vals[2], off_hist, off_sums, off_end;
};
static void make_shuffle_plan(int64_t n, ShufflePlan* p) {
  p->nblocks = ceil_div(n > 0 ? n : 1, kRadixChunk);
  p->hist_n = (int64_t)256 * p->nblocks;
  p->scan_blocks = ceil_div(p->hist_n, kScanBlock);
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) / 256 * 256; return o; };
  for (int i = 0; i < 2; ++i) { p->off_keys[i] = take((size_t)n * 4); p->off_vals[i] = take((size_t)n * 4); }
  p->off_hist = take((size_t)p->hist_n * 4);
  p->off_sums = take((size_t)p->scan_blocks * 4);
  p->off_end = off + 256;
}

extern "C" size_t nncf_group_shuffle_workspace_bytes(int64_t n_rows, int64_t n_keys) {
  (void)n_keys;
  ShufflePlan p;
  make_shuffle_plan(n_rows, &p);
  return p.off_end;
}

extern "C" int nncf_group_shuffle(const int32_t* train_dev, int64_t n_rows, int col, const int64_t* iidx_dev, int64_t n_keys,
                                  const int64_t* row_perm_dev, const int64_t* block_perm_dev, int chop, int32_t* out_dev,
                                  void* workspace_dev, size_t workspace_bytes, void* stream) {
  NNCF_CHECK_ARG(n_rows >= 0 && n_rows < (int64_t)0xffffffffll, "nncf_group_shuffle: n_rows out of range");
  if (n_rows == 0) return NNCF_OK;
  NNCF_CHECK_ARG(train_dev && iidx_dev && row_perm_dev && out_dev && workspace_dev, "nncf_group_shuffle: null argument");
  NNCF_CHECK_ARG(col == 0 || col == 1, "nncf_group_shuffle: col must be 0 (user) or 1 (item)");
  NNCF_CHECK_ARG(n_keys >= 1 && n_keys < (int64_t)0xffffffffll, "nncf_group_shuffle: n_keys out of range");
  NNCF_CHECK_ARG(chop >= 0, "nncf_group_shuffle: chop < 0");
  NNCF_CHECK_ARG(chop == 0 || n_rows / chop == 0 || block_perm_dev, "nncf_group_shuffle: block_perm required when chop > 0");
  ShufflePlan p;
  make_shuffle_plan(n_rows, &p);
  NNCF_CHECK_ARG(workspace_bytes >= p.off_end, "nncf_group_shuffle: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  uint8_t* ws = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(workspace_dev) + 255) & ~uintptr_t(255));
  uint32_t* keys[2] = {reinterpret_cast<uint32_t*>(ws + p.off_keys[0]), reinterpret_cast<uint32_t*>(ws + p.off_keys[1])};
  uint32_t* vals[2] = {reinterpret_cast<uint32_t*>(ws + p.off_vals[0]), reinterpret_cast<uint32_t*>(ws + p.off_vals[1])};
  uint32_t* hist = reinterpret_cast<uint32_t*>(ws + p.off_hist);
  uint32_t* sums = reinterpret_cast<uint32_t*>(ws + p.off_sums);
  make_keys_kernel<<<ceil_div(n_rows, 256), 256, 0, st>>>(train_dev, n_rows, col, iidx_dev, row_perm_dev, keys[0], vals[0]);
  NNCF_LAUNCH_OK();
  int bits = 1;
  while (bits < 32 && (1ll << bits) < n_keys) ++bits;
  int cur = 0;
  for (int shift = 0; shift < bits; shift += 8) {
    radix_hist_kernel<<<p.nblocks, kRadixThreads, 0, st>>>(keys[cur], n_rows, shift, p.nblocks, hist);
    NNCF_LAUNCH_OK();
    scan_block_kernel<<<p.scan_blocks, kScanBlock, 0, st>>>(hist, p.hist_n, sums);
    NNCF_LAUNCH_OK();
    scan_sums_kernel<<<1, kScanBlock, 0, st>>>(sums, p.scan_blocks);
    NNCF_LAUNCH_OK();
    scan_add_kernel<<<p.scan_blocks, kScanBlock, 0, st>>>(hist, p.hist_n, sums);
    NNCF_LAUNCH_OK();
    radix_scatter_kernel<<<p.nblocks, kRadixThreads, 0, st>>>(keys[cur], vals[cur], n_rows, shift, p.nblocks, hist,
                                                              keys[cur ^ 1], vals[cur ^ 1]);
    NNCF_LAUNCH_OK();
    cur ^= 1;
  }
  group_emit_kernel<<<ceil_div(n_rows, 256), 256, 0, st>>>(train_dev, n_rows, vals[cur], block_perm_dev, chop, out_dev);
  NNCF_LAUNCH_OK();
  return NNCF_OK;
}

extern "C" int nncf_assemble_pairs_batch(const int32_t* pos_dev, int B, int k, const int32_t* negs_dev, int neg_col,
                                         int neg_sign, int32_t* out_dev, void* stream) {
  NNCF_CHECK_ARG(pos_dev && negs_dev && out_dev, "nncf_assemble_pairs_batch: null argument");
  NNCF_CHECK_ARG(B >= 1 && k >= 1, "nncf_assemble_pairs_batch: bad sizes");
  NNCF_CHECK_ARG(neg_col == 0 || neg_col == 1, "nncf_assemble_pairs_batch: neg_col must be 0 or 1");
  assemble_pairs_kernel<<<ceil_div((int64_t)(1 + k) * B, 256), 256, 0, (cudaStream_t)stream>>>(pos_dev, B, k, negs_dev,
                                                                                               neg_col, neg_sign, out_dev);
  NNCF_LAUNCH_OK();
  return NNCF_OK;
}

extern "C" int nncf_presample_assemble(const int32_t* pos_dev, int64_t n_links, int k, const int32_t* negs_dev, int neg_col,
                                       int neg_sign, int layout, int32_t* out_dev, void* stream) {
  NNCF_CHECK_ARG(pos_dev && negs_dev && out_dev, "nncf_presample_assemble: null argument");
  NNCF_CHECK_ARG(n_links >= 1 && k >= 1, "nncf_presample_assemble: bad sizes");
  NNCF_CHECK_ARG(neg_col == 0 || neg_col == 1, "nncf_presample_assemble: neg_col must be 0 or 1");
  NNCF_CHECK_ARG(layout == 0 || layout == 1, "nncf_presample_assemble: layout must be 0 (interleaved) or 1 (stacked)");
  presample_assemble_kernel<<<ceil_div(n_links * (1 + k), 256), 256, 0, (cudaStream_t)stream>>>(pos_dev, n_links, k, negs_dev,
                                                                                              neg_col, neg_sign, layout, out_dev);
  NNCF_LAUNCH_OK();
  return NNCF_OK;
}

extern "C" int nncf_assemble_sns_batches(const int32_t* train_dev, int64_t n_batches, int B, int k, const int32_t* negs_dev,
                                         int32_t* user_ids_dev, int32_t* item_ids_dev, void* stream) {
  NNCF_CHECK_ARG(train_dev && negs_dev && user_ids_dev && item_ids_dev, "nncf_assemble_sns_batches: null argument");
  NNCF_CHECK_ARG(n_batches >= 1 && B >= 1 && k >= 1, "nncf_assemble_sns_batches: bad sizes");
  sns_assemble_kernel<<<ceil_div(n_batches * (B + k), 256), 256, 0, (cudaStream_t)stream>>>(train_dev, n_batches, B, k, negs_dev,
                                                                                          user_ids_dev, item_ids_dev);
  NNCF_LAUNCH_OK();
  return NNCF_OK;
}
