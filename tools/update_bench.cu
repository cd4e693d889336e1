// update_bench.cu — developer tool: how fast can one super-step's sparse SGD update (37,888 rows x 512 B into random
// rows of two 1M x 128 fp32 tables, power-law ids with duplicates, rows L2-warm from the gather) be applied?
//   A  red.global.add.v4.f32, warp per row (the finalize kernel's way)
//   B  plain ld.v4 + add + st.v4 (NOT duplicate-safe: speed of the non-atomic path only)
//   C  cp.reduce.async.bulk add.f32 from shared memory, one 512 B row per thread (the fused drain's way)
//   D  plain st.v4 (write-only lower bound)
// usage: update_bench [rows] [exponent]
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <random>
#include <algorithm>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e_), __LINE__); return 2; } } while (0)
constexpr int D = 128;
__global__ void warm(const float* __restrict__ t, const int* __restrict__ ids, int n, float* sink) {
  int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, l = threadIdx.x & 31;
  if (w >= n) return;
  float4 v = __ldg(reinterpret_cast<const float4*>(t + (size_t)ids[w] * D) + l);
  if (v.x == 123456.f) sink[0] = v.y;
}
__global__ void upd_red(float* t, const int* __restrict__ ids, const float* __restrict__ g, int n) {
  int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, l = threadIdx.x & 31;
  if (w >= n) return;
  float4 v = reinterpret_cast<const float4*>(g + (size_t)w * D)[l];
  float* p = t + (size_t)ids[w] * D + 4 * l;
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__global__ void upd_plain(float* t, const int* __restrict__ ids, const float* __restrict__ g, int n) {
  int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, l = threadIdx.x & 31;
  if (w >= n) return;
  float4 v = reinterpret_cast<const float4*>(g + (size_t)w * D)[l];
  float4* p = reinterpret_cast<float4*>(t + (size_t)ids[w] * D) + l;
  float4 o = *p; o.x += v.x; o.y += v.y; o.z += v.z; o.w += v.w; *p = o;
}
__global__ void upd_store(float* t, const int* __restrict__ ids, const float* __restrict__ g, int n) {
  int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, l = threadIdx.x & 31;
  if (w >= n) return;
  float4 v = reinterpret_cast<const float4*>(g + (size_t)w * D)[l];
  reinterpret_cast<float4*>(t + (size_t)ids[w] * D)[l] = v;
}
__global__ void __launch_bounds__(128) upd_bulk(float* t, const int* __restrict__ ids, const float* __restrict__ g, int n) {
  extern __shared__ __align__(128) float sm[];
  const int row0 = blockIdx.x * 128;
  for (int i = threadIdx.x; i < 128 * D / 4; i += 128) {
    int r = i / (D / 4);
    if (row0 + r < n) reinterpret_cast<float4*>(sm)[i] = reinterpret_cast<const float4*>(g + (size_t)row0 * D)[i];
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  const int r = threadIdx.x;
  if (row0 + r < n) {
    float* dst = t + (size_t)ids[row0 + r] * D;
    unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(sm + r * D));
    asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(dst), "r"(s), "r"(D * 4) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  }
}
int main(int argc, char** argv) {
  int n = argc > 1 ? atoi(argv[1]) : 37888;
  double expo = argc > 2 ? atof(argv[2]) : 1.0;
  const int NR = 1000000;
  std::vector<double> cdf(NR); double acc = 0; for (int i = 0; i < NR; ++i) { acc += pow(i + 10.0, -expo); cdf[i] = acc; }
  std::mt19937_64 rng(1); std::vector<int> perm(NR); for (int i = 0; i < NR; ++i) perm[i] = i; std::shuffle(perm.begin(), perm.end(), rng);
  std::vector<int> ids(n); std::uniform_real_distribution<double> U(0, acc);
  for (int i = 0; i < n; ++i) { int r = int(std::lower_bound(cdf.begin(), cdf.end(), U(rng)) - cdf.begin()); ids[i] = perm[std::min(r, NR - 1)]; }
  { std::vector<int> s = ids; std::sort(s.begin(), s.end()); int uniq = int(std::unique(s.begin(), s.end()) - s.begin());
    std::vector<int> c = ids; std::sort(c.begin(), c.end()); int once = 0, maxrun = 0; for (int i = 0; i < n;) { int j = i; while (j < n && c[j] == c[i]) ++j; if (j - i == 1) ++once; maxrun = std::max(maxrun, j - i); i = j; }
    printf("rows %d, distinct ids %d, rows whose id occurs once %d (%.1f%%), hottest id x%d\n", n, uniq, once, 100.0 * once / n, maxrun); }
  float *t, *g, *sink; int* dids;
  CK(cudaMalloc(&t, (size_t)NR * D * 4)); CK(cudaMemset(t, 0, (size_t)NR * D * 4));
  CK(cudaMalloc(&g, (size_t)n * D * 4)); CK(cudaMemset(g, 0, (size_t)n * D * 4));
  CK(cudaMalloc(&sink, 4)); CK(cudaMalloc(&dids, n * 4)); CK(cudaMemcpy(dids, ids.data(), n * 4, cudaMemcpyHostToDevice));
  float* flush; size_t fl = 512u << 20; CK(cudaMalloc(&flush, fl));
  CK(cudaFuncSetAttribute(upd_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * D * 4));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int blocks = (n * 32 + 255) / 256;
  const char* names[4] = {"A red.v4", "B plain RMW", "C bulk reduce", "D plain store"};
  for (int warmL2 = 1; warmL2 >= 0; --warmL2)
    for (int k = 0; k < 4; ++k) {
      float tot = 0; const int reps = 10;
      for (int it = 0; it < reps; ++it) {
        CK(cudaMemsetAsync(flush, it, fl));                       // evict
        if (warmL2) warm<<<blocks, 256>>>(t, dids, n, sink);      // the step's gather leaves its rows in L2
        cudaEventRecord(e0);
        if (k == 0) upd_red<<<blocks, 256>>>(t, dids, g, n);
        else if (k == 1) upd_plain<<<blocks, 256>>>(t, dids, g, n);
        else if (k == 2) upd_bulk<<<(n + 127) / 128, 128, 128 * D * 4>>>(t, dids, g, n);
        else upd_store<<<blocks, 256>>>(t, dids, g, n);
        cudaEventRecord(e1);
        CK(cudaDeviceSynchronize());
        float ms; cudaEventElapsedTime(&ms, e0, e1); tot += ms;
      }
      printf("%-14s rows %s: %.2f us  (%.0f GB/s of update payload)\n", names[k], warmL2 ? "L2-warm" : "cold   ", tot / reps * 1e3, (double)n * D * 4 / (tot / reps * 1e-3) / 1e9);
    }
  return 0;
}
