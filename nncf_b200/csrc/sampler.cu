// sampler.cu — device negative sampler: Walker/Vose alias table in HBM + counter-based Philox4x32-10.
//
// Replaces the reference's 1e8-entry int lookup table + 64-bit LCG (ref: sampler/nodesampler.cpp:14-49,
// :23-26, :69-76) with the same target distribution p_i ∝ dist[i]^power for dist[i] > 0; ids with zero
// degree are never returned (nodesampler.cpp:34,39).  One draw = one 8-byte table read + 4 bytes written.
#include <cmath>
#include <cstring>
#include <vector>
#include "common.cuh"

namespace nncf {

struct Philox {
  static constexpr uint32_t kM0 = 0xD2511F53u, kM1 = 0xCD9E8D57u, kW0 = 0x9E3779B9u, kW1 = 0xBB67AE85u;
  __host__ __device__ static inline void round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#ifdef __CUDA_ARCH__
    const uint32_t hi0 = __umulhi(kM0, c[0]), hi1 = __umulhi(kM1, c[2]);
#else
    const uint32_t hi0 = static_cast<uint32_t>((static_cast<uint64_t>(kM0) * c[0]) >> 32);
    const uint32_t hi1 = static_cast<uint32_t>((static_cast<uint64_t>(kM1) * c[2]) >> 32);
#endif
    const uint32_t lo0 = kM0 * c[0], lo1 = kM1 * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
  }
  // Philox4x32-10: counter (128 bit) x key (64 bit) -> 4 x 32 random bits
  __host__ __device__ static inline void gen(uint64_t ctr, uint64_t key, uint32_t (&out)[4]) {
    uint32_t c[4] = {static_cast<uint32_t>(ctr), static_cast<uint32_t>(ctr >> 32), 0u, 0u};
    uint32_t k0 = static_cast<uint32_t>(key), k1 = static_cast<uint32_t>(key >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      round(c, k0, k1);
      k0 += kW0; k1 += kW1;
    }
    out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
  }
};

// table entry: x = acceptance probability as float bits, y = alias id
__global__ void __launch_bounds__(256)
sample_kernel(const uint2* __restrict__ table, uint32_t n, uint64_t key, uint64_t ctr0, int64_t count,
              int32_t* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  uint32_t r[4];
  Philox::gen(ctr0 + static_cast<uint64_t>(i), key, r);
  const uint64_t x = (static_cast<uint64_t>(r[0]) << 32) | r[1];
  const uint32_t bin = static_cast<uint32_t>(__umul64hi(x, static_cast<uint64_t>(n)));
  const float u = static_cast<float>(r[2] >> 8) * (1.0f / 16777216.0f);
  const uint2 e = __ldg(table + bin);
  out[i] = (u < __uint_as_float(e.x)) ? static_cast<int32_t>(bin) : static_cast<int32_t>(e.y);
}

}  // namespace nncf

using namespace nncf;

struct nncf_sampler {
  uint2* table = nullptr;
  int n = 0;
  uint64_t key = 0;
  uint64_t counter = 0;
  std::vector<float> prob_host;
  std::vector<int32_t> alias_host;
};

extern "C" int nncf_sampler_create(const double* dist_host, int dist_size, double neg_sampling_power, uint64_t rand_seed,
                                   nncf_sampler_t** out) {
  NNCF_CHECK_ARG(dist_host && out, "nncf_sampler_create: null argument");
  NNCF_CHECK_ARG(dist_size >= 1, "nncf_sampler_create: dist_size must be >= 1");
  const int n = dist_size;
  std::vector<double> w(n);
  double sum = 0.0;
  int best = -1;
  for (int i = 0; i < n; ++i) {
    const double deg = dist_host[i];
    NNCF_CHECK_ARG(deg >= 0.0 && std::isfinite(deg), "nncf_sampler_create: dist must be finite and non-negative");
    w[i] = (deg == 0.0) ? 0.0 : std::pow(deg, neg_sampling_power);   // zero degree node will not be sampled
    sum += w[i];
    if (w[i] > 0.0 && (best < 0 || w[i] > w[best])) best = i;
  }
  NNCF_CHECK_ARG(best >= 0 && sum > 0.0, "nncf_sampler_create: distribution has no positive entry");
  // Vose's alias method
  std::vector<double> p(n);
  std::vector<int32_t> alias(n), small, large;
  small.reserve(n); large.reserve(n);
  for (int i = 0; i < n; ++i) {
    p[i] = w[i] / sum * n;
    alias[i] = best;
    (p[i] < 1.0 ? small : large).push_back(i);
  }
  while (!small.empty() && !large.empty()) {
    const int s = small.back(); small.pop_back();
    const int l = large.back();
    alias[s] = l;
    p[l] = (p[l] + p[s]) - 1.0;
    if (p[l] < 1.0) { large.pop_back(); small.push_back(l); }
  }
  for (int i : large) p[i] = 1.0;
  for (int i : small) {                       // numerical leftovers: ~1 for real ids, exactly 0 for zero-degree ids
    if (w[i] > 0.0) p[i] = 1.0;
    else { p[i] = 0.0; alias[i] = best; }
  }
  auto* s = new nncf_sampler();
  s->n = n;
  s->key = rand_seed;
  s->prob_host.resize(n);
  s->alias_host = alias;
  std::vector<uint2> tab(n);
  for (int i = 0; i < n; ++i) {
    float pf = static_cast<float>(p[i]);
    if (w[i] == 0.0) pf = 0.0f;
    if (pf > 1.0f) pf = 1.0f;
    s->prob_host[i] = pf;
    uint32_t bits;
    memcpy(&bits, &pf, 4);
    tab[i] = make_uint2(bits, static_cast<uint32_t>(alias[i]));
  }
  cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&s->table), sizeof(uint2) * n);
  if (e == cudaSuccess) e = cudaMemcpy(s->table, tab.data(), sizeof(uint2) * n, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    set_error(std::string("nncf_sampler_create: ") + cudaGetErrorString(e));
    if (s->table) cudaFree(s->table);
    delete s;
    return NNCF_ECUDA;
  }
  *out = s;
  return NNCF_OK;
}

extern "C" int nncf_sampler_destroy(nncf_sampler_t* s) {
  if (!s) return NNCF_OK;
  if (s->table) cudaFree(s->table);
  delete s;
  return NNCF_OK;
}

extern "C" int nncf_sampler_seek(nncf_sampler_t* s, uint64_t counter) {
  NNCF_CHECK_ARG(s, "nncf_sampler_seek: null sampler");
  s->counter = counter;
  return NNCF_OK;
}

extern "C" int nncf_sampler_sample_batch_dev(nncf_sampler_t* s, int64_t n, int32_t* out_dev, void* stream) {
  NNCF_CHECK_ARG(s && (out_dev || n == 0), "nncf_sampler_sample_batch_dev: null argument");
  NNCF_CHECK_ARG(n >= 0, "nncf_sampler_sample_batch_dev: n < 0");
  if (n == 0) return NNCF_OK;
  sample_kernel<<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(s->table, static_cast<uint32_t>(s->n), s->key,
                                                                    s->counter, n, out_dev);
  NNCF_LAUNCH_OK();
  s->counter += static_cast<uint64_t>(n);
  return NNCF_OK;
}

extern "C" int nncf_sampler_sample_batch_host(nncf_sampler_t* s, int64_t n, int32_t* out_host) {
  NNCF_CHECK_ARG(s && (out_host || n == 0), "nncf_sampler_sample_batch_host: null argument");
  NNCF_CHECK_ARG(n >= 0, "nncf_sampler_sample_batch_host: n < 0");
  if (n == 0) return NNCF_OK;
  int32_t* tmp = nullptr;
  NNCF_CUDA(cudaMalloc(reinterpret_cast<void**>(&tmp), sizeof(int32_t) * n));
  int rc = nncf_sampler_sample_batch_dev(s, n, tmp, nullptr);
  if (rc == NNCF_OK) {
    cudaError_t e = cudaMemcpy(out_host, tmp, sizeof(int32_t) * n, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { set_error(cudaGetErrorString(e)); rc = NNCF_ECUDA; }
  }
  cudaFree(tmp);
  return rc;
}

extern "C" int nncf_sampler_export_table(nncf_sampler_t* s, float* prob_host, int32_t* alias_host) {
  NNCF_CHECK_ARG(s && prob_host && alias_host, "nncf_sampler_export_table: null argument");
  memcpy(prob_host, s->prob_host.data(), sizeof(float) * s->n);
  memcpy(alias_host, s->alias_host.data(), sizeof(int32_t) * s->n);
  return NNCF_OK;
}
