// score_bench.cu — developer tool: runs score_grad_tc_kernel<2> on synthetic tile images, times it with CUDA events
// and prints the in-kernel clock64 timeline of CTA 0 (stamps: 0 start, 1 X landed, 2 first MMA1 issued, 8+t G(t) ready
// at the MMA warp, 16+t S(t) ready at the epilogue, 24+t G(t) written, 3 all MMAs issued, 4 dX ready, 5 drained, 6 end).
// usage: score_bench [B] [R] [fuse]     fuse = 1: the drain applies the sparse SGD update to two 1M x 128 tables (power-law ids)
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <random>
#include <cmath>
#include "../nncf_b200/csrc/score_tc.cuh"
namespace nncf { int launch_score_tc_nsub1(const ScoreTcArgs&, int, int, cudaStream_t) { return 0; } int launch_score_tc_nsub2(const ScoreTcArgs&, int, int, cudaStream_t) { return 0; } int launch_score_tc_nsub4(const ScoreTcArgs&, int, int, cudaStream_t) { return 0; }
void set_error(const std::string&) {} std::atomic<int64_t> g_launches{0}; }
using namespace nncf;
// rewrites the tile images with generic stores from every SM, like the gather kernel does right before the score kernel
__global__ void touch_images(uint4* a, uint4* b, size_t n16, unsigned salt) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) {
    a[i] = make_uint4(salt & 0x3c003c00u, 0x3c003c00u & (unsigned)i, 0, 0); b[i] = make_uint4(0x3c003c00u & salt, 0, 0x38003800u & (unsigned)i, 0);
  }
}
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e_), __LINE__); return 2; } } while (0)
int main(int argc, char** argv) {
  int B = argc > 1 ? atoi(argv[1]) : 512, R = argc > 2 ? atoi(argv[2]) : 1, fuse = argc > 3 ? atoi(argv[3]) : 0, touch = argc > 4 ? atoi(argv[4]) : 0; const double expo_scale = argc > 5 ? atof(argv[5]) : 1.0;
  constexpr int NSUB = 2, DP = 128;
  int rp = (B + 127) / 128 * 128, nblk = rp / 128;
  size_t img = (size_t)R * rp * DP * 2, nel = (size_t)R * rp * DP;
  uint8_t *U, *V; float *dU, *dV, *cU, *cV, *sp; double* loss; long long* dbg;
  CK(cudaMalloc(&U, img)); CK(cudaMalloc(&V, img)); CK(cudaMemset(U, 0, img)); CK(cudaMemset(V, 0, img));
  CK(cudaMalloc(&dU, nel * 4)); CK(cudaMalloc(&dV, nel * 4));
  CK(cudaMalloc(&cU, R * rp * 4)); CK(cudaMalloc(&cV, R * rp * 4)); CK(cudaMalloc(&sp, R * rp * 4));
  CK(cudaMalloc(&loss, R * 8)); CK(cudaMemset(loss, 0, R * 8));
  size_t ndbg = (size_t)nblk * 8 * 2 * R * 64;
  CK(cudaMalloc(&dbg, ndbg * 8)); CK(cudaMemset(dbg, 0, ndbg * 8));
  ScoreTcArgs a{};
  a.Uimg = U; a.Vimg = V; a.dU = dU; a.dV = dV; a.corrU = cU; a.corrV = cV; a.spos = sp; a.loss = loss;
  a.rows_pad = rp; a.B = B; a.scheme = NNCF_SCHEME_NEG_SHARED; a.loss_kind = NNCF_LOSS_SKIP_GRAM; a.lambda = 128.f; a.gamma = 10.f;
  a.dbg = dbg; a.split = argc > 6 ? atoi(argv[6]) : 1;
  if (fuse) {
    const int NR = 1000000;
    float *tu, *tv; int *iu, *iv; unsigned int* lc;
    CK(cudaMalloc(&tu, (size_t)NR * DP * 4)); CK(cudaMalloc(&tv, (size_t)NR * DP * 4));
    CK(cudaMemset(tu, 0, (size_t)NR * DP * 4)); CK(cudaMemset(tv, 0, (size_t)NR * DP * 4));
    CK(cudaMalloc(&lc, R * 4)); CK(cudaMemset(lc, 0, R * 4));
    std::mt19937_64 rng(1);
    auto draw = [&](double expo, std::vector<int>& out) {
      std::vector<double> cdf(NR); double acc = 0; for (int i = 0; i < NR; ++i) { acc += pow(i + 10.0, -expo); cdf[i] = acc; }
      std::vector<int> perm(NR); for (int i = 0; i < NR; ++i) perm[i] = i; std::shuffle(perm.begin(), perm.end(), rng);
      std::uniform_real_distribution<double> U(0, acc);
      for (auto& x : out) { int r = int(std::lower_bound(cdf.begin(), cdf.end(), U(rng)) - cdf.begin()); x = perm[std::min(r, NR - 1)]; }
    };
    std::vector<int> hu((size_t)R * B), hv((size_t)R * B); draw(0.8 * expo_scale, hu); draw(1.0 * expo_scale, hv);
    CK(cudaMalloc(&iu, hu.size() * 4)); CK(cudaMalloc(&iv, hv.size() * 4));
    CK(cudaMemcpy(iu, hu.data(), hu.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(iv, hv.data(), hv.size() * 4, cudaMemcpyHostToDevice));
    a.fuse_sgd = 1; a.d = DP; a.neg_lr = -0.01f; a.table_u = tu; a.table_v = tv; a.ids_u = iu; a.ids_v = iv; a.ids_stride_u = a.ids_stride_v = B;
    a.loss_count = lc; a.shards_u.n = a.shards_v.n = 1;
  }
  using C = ScoreTcCfg<NSUB>;
  CK(cudaFuncSetAttribute(score_grad_tc_kernel<NSUB, 0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::kSmemBytes));
  CK(cudaFuncSetAttribute(score_grad_tc_kernel<NSUB, 0, false>, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
  { int nb = 0; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, score_grad_tc_kernel<NSUB, 0, false>, kScoreThreads, C::kSmemBytes)); cudaFuncAttributes fa; CK(cudaFuncGetAttributes(&fa, score_grad_tc_kernel<NSUB, 0, false>));
    printf("resident CTAs per SM: %d (dyn smem %zu B, static %zu B, regs %d, maxdyn %d, carveout %d)\n", nb, (size_t)C::kSmemBytes, fa.sharedSizeBytes, fa.numRegs, fa.maxDynamicSharedSizeBytes, fa.preferredShmemCarveout);
    for (size_t sz = 100 * 1024; sz <= C::kSmemBytes; sz += 1024) { int q = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&q, score_grad_tc_kernel<NSUB, 0, false>, kScoreThreads, sz); if (q < 2) { printf("  occupancy drops to %d at dyn smem %zu\n", q, sz); break; } } }
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int it = 0; it < 3; ++it) score_grad_tc_kernel<NSUB, 0, false><<<dim3(nblk * a.split, 2, R), kScoreThreads, C::kSmemBytes>>>(a);
  CK(cudaDeviceSynchronize());
  cudaEventRecord(e0);
  for (int it = 0; it < 20; ++it) {
    if (touch) touch_images<<<1184, 256>>>((uint4*)U, (uint4*)V, img / 16, it);
    score_grad_tc_kernel<NSUB, 0, false><<<dim3(nblk * a.split, 2, R), kScoreThreads, C::kSmemBytes>>>(a);
  }
  cudaEventRecord(e1);
  CK(cudaDeviceSynchronize());
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  printf("B=%d R=%d grid=%d CTAs: %.2f us per launch\n", B, R, nblk * 2 * R, ms / 20 * 1e3);
  std::vector<long long> h(ndbg);
  CK(cudaMemcpy(h.data(), dbg, ndbg * 8, cudaMemcpyDeviceToHost));
  long long t0 = h[0];
  for (int i = 0; i < 40; ++i) if (h[i]) printf("  stamp %2d : +%lld cyc\n", i, h[i] - t0);
  // all CTAs of the LAST launch: global-timer start/end relative to the earliest start, SM id, in-CTA cycles
  int ncta = nblk * 2 * R;
  long long g0 = h[40];
  for (int c = 0; c < ncta; ++c) g0 = std::min(g0, h[(size_t)c * 64 + 40]);
  long long last_end = 0, sum_cyc = 0, max_cyc = 0; std::vector<int> per_sm(256, 0);
  for (int c = 0; c < ncta; ++c) { const long long* d = &h[(size_t)c * 64]; last_end = std::max(last_end, d[41] - g0); sum_cyc += d[6] - d[0]; max_cyc = std::max(max_cyc, d[6] - d[0]); per_sm[d[42] & 255]++; }
  int sm_used = 0, sm_max = 0; for (int s = 0; s < 256; ++s) { if (per_sm[s]) ++sm_used; sm_max = std::max(sm_max, per_sm[s]); }
  printf("  CTAs %d on %d SMs (max %d per SM); last CTA ends +%lld ns after the first starts; CTA cycles mean %lld max %lld\n", ncta, sm_used, sm_max, last_end, sum_cyc / ncta, max_cyc);
  if (getenv("NNCF_DUMP_CTAS")) for (int c = 0; c < ncta; ++c) { const long long* d = &h[(size_t)c * 64]; printf("    cta %3d (ob %d side %d r %d) sm %3lld start +%6lld ns end +%6lld ns  cycles %lld  loop %lld drain %lld\n", c, c % nblk, (c / nblk) % 2, c / (2 * nblk), d[42], d[40] - g0, d[41] - g0, d[6] - d[0], d[3] - d[2], d[5] - d[4]); }
  return 0;
}
