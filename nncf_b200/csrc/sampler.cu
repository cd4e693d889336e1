// sampler.cu — device negative sampler: Walker/Vose alias table in HBM + counter-based Philox4x32-10.
//
// Replaces the reference's 1e8-entry int lookup table + 64-bit LCG (ref: sampler/nodesampler.cpp:14-49,
// :23-26, :69-76) with the same target distribution p_i ∝ dist[i]^power for dist[i] > 0; ids with zero
// degree are never returned (nodesampler.cpp:34,39).  One draw = one 8-byte table read + 4 bytes written.
#include <cmath>
#include <cstring>
#include <vector>
#include "common.cuh"
#include "alias.cuh"

namespace nncf {

// table entry: x = acceptance probability as float bits, y = alias id
__global__ void __launch_bounds__(256)
sample_kernel(const uint2* __restrict__ table, uint32_t n, uint64_t key, uint64_t ctr0, int64_t count,
              int32_t* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  uint32_t r[4];
  Philox::gen(ctr0 + static_cast<uint64_t>(i), key, r);
  out[i] = alias_draw(table, n, r);
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
  for (int i = 0; i < n; ++i) {
    const double deg = dist_host[i];
    NNCF_CHECK_ARG(deg >= 0.0 && std::isfinite(deg), "nncf_sampler_create: dist must be finite and non-negative");
    w[i] = (deg == 0.0) ? 0.0 : std::pow(deg, neg_sampling_power);   // zero degree node will not be sampled
  }
  auto* s = new nncf_sampler();
  s->n = n;
  s->key = rand_seed;
  std::vector<uint2> tab;
  if (!build_alias_table(w, tab, s->prob_host, s->alias_host)) {
    delete s;
    set_error("nncf_sampler_create: distribution has no positive entry");
    return NNCF_EINVAL;
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
