// group_sampler.cu — device GroupSampler: "first sample a group (item), then sample its positive members (users),
// followed by sampling negative members".   ref: configs/data_utils.py:244-408 (class GroupSampler), used by
// models/train_group_neg_shared.py:33-37,56 and models/train_group_sample.py:31-36,86-87 when
// group_shuffling_trick is False.
//
// What the reference does per batch in a Python loop over groups, done here in one launch for a whole epoch of batches:
//   * groups ~ degree^1 over the group column            (:281-282: get_sampler(..., neg_sampling_power=1.))
//   * `chop` positive members per group, uniform WITH replacement from the group's member list (:320-323)
//   * sample_with_negs: random_rounding(k * chop * p_n/p_d[group]) negative members per group from the member
//     sampler (:357-362), top-up rounds of extra groups until B(1+k) rows exist (:366-381), positives first, then
//     negatives, truncated to B(1+k) (:383-399)
// Data in HBM: CSR of the train links by group (indptr int64[G+1], members int32[N] in train order), two alias tables
// (8 B / id), p_n/p_d as float[G].  Randomness: Philox4x32-10, one key per purpose, counters derived from
// (call counter, batch, slot): every output row is a pure function of the seed, so results do not depend on the grid.
// The reference draws from np.random / a time-seeded LCG (not reproducible): parity is distributional (tests).
#include <cmath>
#include <vector>
#include "common.cuh"
#include "alias.cuh"

namespace nncf {

struct GsArgs {
  const uint2* group_table; uint32_t n_groups;       // alias table over group ids (degree^1)
  const uint2* member_table; uint32_t n_members;     // alias table of the negative-member sampler
  const int64_t* indptr; const int32_t* members;     // CSR by group
  const float* pnd;                                  // p_n / p_d per group id
  uint64_t key; uint64_t call;                       // Philox key, per-call counter block
  int chop, B, k, neg_sign, group_by_user;
  int32_t* out;                                      // [n_batches][rows][3]
  int32_t* n_pos;                                    // [n_batches] or NULL
  int32_t* cand;                                     // workspace [n_batches][cmax][2] (group id, first negative row)
  int cmax;
  int* err;
};

// Philox counter of (batch, slot): 2^24 slots per batch (range-checked by the entry points), batches up to 2^40.  The call
// number is not packed into the counter (its field used to overlap the batch field after 2^20 batches or calls): it is
// mixed into the KEY on the host (gs_args), so no two calls share a stream however large an epoch is.
__device__ __forceinline__ uint64_t gs_ctr(uint64_t /*call*/, uint64_t batch, uint64_t slot) {
  return (batch << 24) ^ slot;
}

__device__ __forceinline__ void gs_write(const GsArgs& a, int32_t* row, int32_t member, int32_t group, int32_t label) {
  row[0] = a.group_by_user ? group : member;         // (:330-331, :401-402: columns swapped when grouping by user)
  row[1] = a.group_by_user ? member : group;
  row[2] = label;
}

// uniform member of a group's list (np.random.choice(gms, chop): with replacement)
__device__ __forceinline__ int32_t gs_member(const GsArgs& a, int32_t group, uint64_t ctr) {
  uint32_t r[4];
  Philox::gen(ctr, a.key ^ 0x6d656d62ull, r);
  const int64_t lo = a.indptr[group], cnt = a.indptr[group + 1] - lo;
  const uint64_t x = (static_cast<uint64_t>(r[0]) << 32) | r[1];
  return a.members[lo + static_cast<int64_t>(__umul64hi(x, static_cast<uint64_t>(cnt)))];
}

// sample(batch_size_p): one thread per output row
__global__ void __launch_bounds__(256)
group_sample_kernel(GsArgs a, int64_t n_rows_total) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rows_total) return;
  const int64_t b = i / a.B;
  const int row = static_cast<int>(i - b * a.B);
  const int g = row / a.chop;
  uint32_t r[4];
  Philox::gen(gs_ctr(a.call, b, g), a.key, r);
  const int32_t group = alias_draw(a.group_table, a.n_groups, r);
  const int32_t member = gs_member(a, group, gs_ctr(a.call, b, row));
  gs_write(a, a.out + i * 3, member, group, 1);
}

// sample_with_negs(batch_size_p, k): one CTA per batch
__global__ void __launch_bounds__(256)
group_sample_negs_kernel(GsArgs a) {
  __shared__ int s_scan[256];
  __shared__ int s_total, s_used, s_neg_total;
  const int b = blockIdx.x, tid = threadIdx.x;
  const int whole = a.B * (1 + a.k);
  int32_t* cand = a.cand + (int64_t)b * a.cmax * 2;
  // candidate c: its group and its number of negatives are pure functions of (key, call, batch, c)
  auto candidate = [&](int c, int32_t& group, int& nneg) {
    uint32_t r[4];
    Philox::gen(gs_ctr(a.call, b, c), a.key, r);
    group = alias_draw(a.group_table, a.n_groups, r);
    const float x = static_cast<float>(a.k) * a.chop * a.pnd[group];
    const int xi = static_cast<int>(x);
    const float u = static_cast<float>(r[3] >> 8) * (1.0f / 16777216.0f);
    nneg = xi + ((u < x - static_cast<float>(xi)) ? 1 : 0);                 // random_rounding (:357)
  };
  // running exclusive prefix of negatives over candidates [from, to); returns the new running total
  auto extend = [&](int from, int to, int run) {
    for (int c0 = from; c0 < to; c0 += 256) {
      const int c = c0 + tid;
      int32_t group = 0; int nneg = 0;
      if (c < to) candidate(c, group, nneg);
      s_scan[tid] = nneg;
      __syncthreads();
      for (int off = 1; off < 256; off <<= 1) {
        const int v = tid >= off ? s_scan[tid - off] : 0;
        __syncthreads();
        s_scan[tid] += v;
        __syncthreads();
      }
      if (c < to) { cand[2 * c] = group; cand[2 * c + 1] = run + s_scan[tid] - nneg; }
      run += s_scan[255];
      __syncthreads();
    }
    return run;
  };
  int used = (a.B + a.chop - 1) / a.chop;               // strict_return_shape: ceil(B / chop) groups (:347-348)
  int negs = extend(0, used, 0);
  for (int round = 0; round < 10; ++round) {            // top-up rounds (:366-381; the reference asserts after 10)
    const int diff = whole - (used * a.chop + negs);
    if (diff <= 0) break;
    int add = diff <= 100 ? diff : diff / (a.chop * (1 + a.k)) + 100;
    if (used + add > a.cmax) add = a.cmax - used;
    if (add <= 0) break;
    negs = extend(used, used + add, negs);
    used += add;
  }
  if (tid == 0) {
    s_used = used; s_neg_total = negs; s_total = used * a.chop + negs;
    if (s_total < whole) atomicExch(a.err, 1);          // reference: assert False after 10 rounds
  }
  __syncthreads();
  used = s_used; negs = s_neg_total;
  const int P = min(used * a.chop, whole);
  int32_t* out = a.out + (int64_t)b * whole * 3;
  for (int i = tid; i < P; i += 256) {
    const int32_t group = cand[2 * (i / a.chop)];
    gs_write(a, out + (int64_t)i * 3, gs_member(a, group, gs_ctr(a.call, b, i)), group, 1);
  }
  const int nn = min(negs, whole - P);
  for (int j = tid; j < nn; j += 256) {
    // candidate owning negative j: last c with first_neg[c] <= j (candidates without negatives share a start)
    int lo = 0, hi = used - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (cand[2 * mid + 1] <= j) lo = mid; else hi = mid - 1;
    }
    uint32_t r[4];
    Philox::gen(gs_ctr(a.call, b, j), a.key ^ 0x6e656773ull, r);
    gs_write(a, out + (int64_t)(P + j) * 3, alias_draw(a.member_table, a.n_members, r), cand[2 * lo], a.neg_sign);
  }
  for (int i = P + nn + tid; i < whole; i += 256) gs_write(a, out + (int64_t)i * 3, 0, 0, 0);   // only after an error
  if (tid == 0 && a.n_pos) a.n_pos[b] = P;
}

}  // namespace nncf

using namespace nncf;

struct nncf_group_sampler {
  uint2 *group_table = nullptr, *member_table = nullptr;
  int64_t* indptr = nullptr;
  int32_t* members = nullptr;
  float* pnd = nullptr;
  int32_t* cand = nullptr; size_t cand_elems = 0;
  int* err = nullptr;
  uint32_t n_groups = 0, n_members = 0;
  int chop = 1, neg_sign = 0, group_by_user = 0;
  uint64_t key = 0, call = 0;
};

extern "C" int nncf_group_sampler_destroy(nncf_group_sampler_t* g) {
  if (!g) return NNCF_OK;
  void* ptrs[] = {g->group_table, g->member_table, g->indptr, g->members, g->pnd, g->cand, g->err};
  for (void* p : ptrs) if (p) cudaFree(p);
  delete g;
  return NNCF_OK;
}

template <typename T>
static bool gs_upload(T** dst, const std::vector<T>& src) {
  if (cudaMalloc(reinterpret_cast<void**>(dst), sizeof(T) * (src.empty() ? 1 : src.size())) != cudaSuccess) return false;
  return src.empty() || cudaMemcpy(*dst, src.data(), sizeof(T) * src.size(), cudaMemcpyHostToDevice) == cudaSuccess;
}

extern "C" int nncf_group_sampler_create(const int32_t* train_host, int64_t n_links, int group_by, int chop, int neg_dist,
                                         int neg_sign, double neg_sampling_power, uint64_t rand_seed,
                                         nncf_group_sampler_t** out) {
  NNCF_CHECK_ARG(train_host && out, "nncf_group_sampler_create: null argument");
  NNCF_CHECK_ARG(n_links >= 1, "nncf_group_sampler_create: empty train array");
  NNCF_CHECK_ARG(group_by == 0 || group_by == 1, "[ERROR] Illegal group_by");
  NNCF_CHECK_ARG(neg_dist >= 0 && neg_dist <= 2, "[ERROR] Illegal neg_dist");
  NNCF_CHECK_ARG(chop >= 1, "nncf_group_sampler_create: chop must be >= 1");
  const int gcol = group_by == 0 ? 1 : 0, mcol = 1 - gcol;        // (:267-272)
  int32_t gmax = 0, mmax = 0;
  for (int64_t i = 0; i < n_links; ++i) {
    const int32_t g = train_host[3 * i + gcol], m = train_host[3 * i + mcol];
    NNCF_CHECK_ARG(g >= 0 && m >= 0, "nncf_group_sampler_create: negative id");
    if (g > gmax) gmax = g;
    if (m > mmax) mmax = m;
  }
  const int G = gmax + 1, M = mmax + 1;
  // CSR by group, members in train order (:301-304)
  std::vector<int64_t> indptr(G + 1, 0);
  std::vector<double> gdeg(G, 0.0), mdeg(M, 0.0);
  for (int64_t i = 0; i < n_links; ++i) { indptr[train_host[3 * i + gcol] + 1] += 1; gdeg[train_host[3 * i + gcol]] += 1.0; mdeg[train_host[3 * i + mcol]] += 1.0; }
  for (int g = 0; g < G; ++g) indptr[g + 1] += indptr[g];
  std::vector<int32_t> members(n_links);
  { std::vector<int64_t> cur(indptr.begin(), indptr.end() - 1);
    for (int64_t i = 0; i < n_links; ++i) members[cur[train_host[3 * i + gcol]]++] = train_host[3 * i + mcol]; }
  // group sampler: unigram, power 1 (:281-282); member sampler: neg_dist, neg_sampling_power (:283-285; 'uniform' sets
  // every positive degree to 1, configs/data_utils.py:203-204)
  std::vector<double> mw(M);
  for (int m = 0; m < M; ++m) mw[m] = mdeg[m] == 0.0 ? 0.0 : (neg_dist == 0 ? std::pow(mdeg[m], neg_sampling_power) : 1.0);
  // p_n / p_d (:287-299): 1 for unigram and uniform_no_correction, (1 / #groups) / (degree share) for uniform
  int64_t gset = 0;
  for (int g = 0; g < G; ++g) gset += gdeg[g] > 0.0;
  std::vector<float> pnd(G, 0.0f);
  for (int g = 0; g < G; ++g)
    if (gdeg[g] > 0.0) pnd[g] = neg_dist == 1 ? static_cast<float>(1.0 / static_cast<double>(gset) / (gdeg[g] / static_cast<double>(n_links))) : 1.0f;
  auto* gs = new nncf_group_sampler();
  gs->n_groups = G; gs->n_members = M; gs->chop = chop; gs->neg_sign = neg_sign; gs->group_by_user = group_by;
  gs->key = rand_seed;
  std::vector<uint2> gtab, mtab; std::vector<float> prob; std::vector<int32_t> alias;
  bool ok = build_alias_table(gdeg, gtab, prob, alias) && build_alias_table(mw, mtab, prob, alias);
  ok = ok && gs_upload(&gs->group_table, gtab) && gs_upload(&gs->member_table, mtab) && gs_upload(&gs->indptr, indptr) &&
       gs_upload(&gs->members, members) && gs_upload(&gs->pnd, pnd);
  ok = ok && cudaMalloc(reinterpret_cast<void**>(&gs->err), sizeof(int)) == cudaSuccess &&
       cudaMemset(gs->err, 0, sizeof(int)) == cudaSuccess;
  if (!ok) {
    set_error(std::string("nncf_group_sampler_create: ") + cudaGetErrorString(cudaGetLastError()));
    nncf_group_sampler_destroy(gs);
    return NNCF_ECUDA;
  }
  *out = gs;
  return NNCF_OK;
}

static GsArgs gs_args(nncf_group_sampler* g, int B, int k, int32_t* out, int32_t* n_pos) {
  GsArgs a{};
  a.group_table = g->group_table; a.n_groups = g->n_groups; a.member_table = g->member_table; a.n_members = g->n_members;
  a.indptr = g->indptr; a.members = g->members; a.pnd = g->pnd;
  a.call = g->call++;
  {  // per-call key: splitmix64 of (seed key, call number)
    uint64_t z = g->key + 0x9E3779B97F4A7C15ull * (a.call + 1);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; z ^= z >> 31;
    a.key = z;
  }
  a.chop = g->chop; a.B = B; a.k = k; a.neg_sign = g->neg_sign; a.group_by_user = g->group_by_user;
  a.out = out; a.n_pos = n_pos; a.err = g->err;
  return a;
}

extern "C" int nncf_group_sampler_sample(nncf_group_sampler_t* g, int batch_size_p, int n_batches, int32_t* out_dev,
                                         void* stream) {
  NNCF_CHECK_ARG(g && out_dev, "nncf_group_sampler_sample: null argument");
  NNCF_CHECK_ARG(batch_size_p >= 1 && batch_size_p < (1 << 24) && n_batches >= 0, "nncf_group_sampler_sample: bad sizes");
  if (n_batches == 0) return NNCF_OK;
  GsArgs a = gs_args(g, batch_size_p, 0, out_dev, nullptr);
  const int64_t n = (int64_t)n_batches * batch_size_p;
  group_sample_kernel<<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(a, n);
  NNCF_LAUNCH_OK();
  return NNCF_OK;
}

extern "C" int nncf_group_sampler_sample_with_negs(nncf_group_sampler_t* g, int batch_size_p, int k, int n_batches,
                                                   int32_t* out_dev, int32_t* n_pos_dev, void* stream) {
  NNCF_CHECK_ARG(g && out_dev, "nncf_group_sampler_sample_with_negs: null argument");
  NNCF_CHECK_ARG(batch_size_p >= 1 && k >= 1 && n_batches >= 0, "nncf_group_sampler_sample_with_negs: bad sizes");
  NNCF_CHECK_ARG((int64_t)batch_size_p * (1 + k) < (1 << 24), "nncf_group_sampler_sample_with_negs: batch too large");
  if (n_batches == 0) return NNCF_OK;
  const int ng0 = (batch_size_p + g->chop - 1) / g->chop;
  const int cmax = ng0 + 10 * (batch_size_p / g->chop + 101);          // bound of the 10 top-up rounds
  const size_t need = (size_t)n_batches * cmax * 2;
  if (need > g->cand_elems) {
    if (g->cand) cudaFree(g->cand);
    g->cand = nullptr; g->cand_elems = 0;
    NNCF_CUDA(cudaMalloc(reinterpret_cast<void**>(&g->cand), need * sizeof(int32_t)));
    g->cand_elems = need;
  }
  GsArgs a = gs_args(g, batch_size_p, k, out_dev, n_pos_dev);
  a.cand = g->cand; a.cmax = cmax;
  group_sample_negs_kernel<<<n_batches, 256, 0, (cudaStream_t)stream>>>(a);
  NNCF_LAUNCH_OK();
  return NNCF_OK;
}

/* 1 if a sample_with_negs batch could not be filled within the reference's 10 top-up rounds (the reference asserts);
 * synchronises the device; clears the flag */
extern "C" int nncf_group_sampler_check(nncf_group_sampler_t* g, int* failed_out) {
  NNCF_CHECK_ARG(g && failed_out, "nncf_group_sampler_check: null argument");
  int h = 0;
  NNCF_CUDA(cudaMemcpy(&h, g->err, sizeof(int), cudaMemcpyDeviceToHost));
  if (h) NNCF_CUDA(cudaMemset(g->err, 0, sizeof(int)));
  *failed_out = h;
  return NNCF_OK;
}
