// alias.cuh — pieces shared by the negative sampler (sampler.cu) and the group sampler (group_sampler.cu):
// counter-based Philox4x32-10, Walker/Vose alias-table construction (host) and the one-read alias draw (device).
#pragma once
#include <cstdint>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>

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


// table entry: x = acceptance probability as float bits, y = alias id.  r = one Philox block: r[0..1] pick the bin
// (64-bit multiply-high: unbiased to 2^-64 n), r[2] is the 24-bit acceptance test.
__device__ __forceinline__ int32_t alias_draw(const uint2* __restrict__ table, uint32_t n, const uint32_t (&r)[4]) {
  const uint64_t x = (static_cast<uint64_t>(r[0]) << 32) | r[1];
  const uint32_t bin = static_cast<uint32_t>(__umul64hi(x, static_cast<uint64_t>(n)));
  const float u = static_cast<float>(r[2] >> 8) * (1.0f / 16777216.0f);
  const uint2 e = __ldg(table + bin);
  return (u < __uint_as_float(e.x)) ? static_cast<int32_t>(bin) : static_cast<int32_t>(e.y);
}

// Vose's alias method over non-negative weights w (zero weight = never drawn: probability 0 and an alias with
// positive weight).  Returns false when no weight is positive.
inline bool build_alias_table(const std::vector<double>& w, std::vector<uint2>& tab, std::vector<float>& prob,
                              std::vector<int32_t>& alias) {
  const int n = static_cast<int>(w.size());
  double sum = 0.0;
  int best = -1;
  for (int i = 0; i < n; ++i) {
    sum += w[i];
    if (w[i] > 0.0 && (best < 0 || w[i] > w[best])) best = i;
  }
  if (best < 0 || !(sum > 0.0)) return false;
  std::vector<double> p(n);
  std::vector<int32_t> small, large;
  alias.assign(n, best);
  small.reserve(n); large.reserve(n);
  for (int i = 0; i < n; ++i) {
    p[i] = w[i] / sum * n;
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
  for (int i : small) {                       // numerical leftovers: ~1 for real ids, exactly 0 for zero-weight ids
    if (w[i] > 0.0) p[i] = 1.0;
    else { p[i] = 0.0; alias[i] = best; }
  }
  prob.resize(n);
  tab.resize(n);
  for (int i = 0; i < n; ++i) {
    float pf = static_cast<float>(p[i]);
    if (w[i] == 0.0) pf = 0.0f;
    if (pf > 1.0f) pf = 1.0f;
    prob[i] = pf;
    uint32_t bits;
    memcpy(&bits, &pf, 4);
    tab[i] = make_uint2(bits, static_cast<uint32_t>(alias[i]));
  }
  return true;
}

}  // namespace nncf
