// score_bench.cu — developer tool: runs score_grad_tc_kernel<2> on synthetic tile images, times it with CUDA events
// and prints the in-kernel clock64 timeline of CTA 0 (stamps: 0 start, 1 X landed, 2 first MMA1 issued, 8+t G(t) ready
// at the MMA warp, 16+t S(t) ready at the epilogue, 24+t G(t) written, 3 all MMAs issued, 4 dX ready, 5 drained, 6 end).
// usage: score_bench [B] [R]
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../nncf_b200/csrc/score_tc.cuh"
namespace nncf { int launch_score_tc_nsub1(const ScoreTcArgs&, int, int, cudaStream_t) { return 0; } int launch_score_tc_nsub2(const ScoreTcArgs&, int, int, cudaStream_t) { return 0; } int launch_score_tc_nsub4(const ScoreTcArgs&, int, int, cudaStream_t) { return 0; }
void set_error(const std::string&) {} std::atomic<int64_t> g_launches{0}; }
using namespace nncf;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e_), __LINE__); return 2; } } while (0)
int main(int argc, char** argv) {
  int B = argc > 1 ? atoi(argv[1]) : 512, R = argc > 2 ? atoi(argv[2]) : 1;
  constexpr int NSUB = 2, DP = 128;
  int rp = (B + 127) / 128 * 128, nblk = rp / 128;
  size_t img = (size_t)R * rp * DP * 2, nel = (size_t)R * rp * DP;
  uint8_t *U, *V; float *dU, *dV, *cU, *cV, *sp; double* loss; long long* dbg;
  CK(cudaMalloc(&U, img)); CK(cudaMalloc(&V, img)); CK(cudaMemset(U, 0, img)); CK(cudaMemset(V, 0, img));
  CK(cudaMalloc(&dU, nel * 4)); CK(cudaMalloc(&dV, nel * 4));
  CK(cudaMalloc(&cU, R * rp * 4)); CK(cudaMalloc(&cV, R * rp * 4)); CK(cudaMalloc(&sp, R * rp * 4));
  CK(cudaMalloc(&loss, R * 8)); CK(cudaMemset(loss, 0, R * 8));
  size_t ndbg = (size_t)nblk * 2 * R * 64;
  CK(cudaMalloc(&dbg, ndbg * 8)); CK(cudaMemset(dbg, 0, ndbg * 8));
  ScoreTcArgs a{};
  a.Uimg = U; a.Vimg = V; a.dU = dU; a.dV = dV; a.corrU = cU; a.corrV = cV; a.spos = sp; a.loss = loss;
  a.rows_pad = rp; a.B = B; a.scheme = NNCF_SCHEME_NEG_SHARED; a.loss_kind = NNCF_LOSS_SKIP_GRAM; a.lambda = 128.f; a.gamma = 10.f;
  a.dbg = dbg;
  using C = ScoreTcCfg<NSUB>;
  CK(cudaFuncSetAttribute(score_grad_tc_kernel<NSUB, 0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::kSmemBytes));
  CK(cudaFuncSetAttribute(score_grad_tc_kernel<NSUB, 0, false>, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
  { int nb = 0; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, score_grad_tc_kernel<NSUB, 0, false>, kScoreThreads, C::kSmemBytes)); cudaFuncAttributes fa; CK(cudaFuncGetAttributes(&fa, score_grad_tc_kernel<NSUB, 0, false>));
    printf("resident CTAs per SM: %d (dyn smem %zu B, static %zu B, regs %d, maxdyn %d, carveout %d)\n", nb, (size_t)C::kSmemBytes, fa.sharedSizeBytes, fa.numRegs, fa.maxDynamicSharedSizeBytes, fa.preferredShmemCarveout);
    for (size_t sz = 100 * 1024; sz <= C::kSmemBytes; sz += 1024) { int q = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&q, score_grad_tc_kernel<NSUB, 0, false>, kScoreThreads, sz); if (q < 2) { printf("  occupancy drops to %d at dyn smem %zu\n", q, sz); break; } } }
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int it = 0; it < 3; ++it) score_grad_tc_kernel<NSUB, 0, false><<<dim3(nblk, 2, R), kScoreThreads, C::kSmemBytes>>>(a);
  CK(cudaDeviceSynchronize());
  cudaEventRecord(e0);
  for (int it = 0; it < 20; ++it) score_grad_tc_kernel<NSUB, 0, false><<<dim3(nblk, 2, R), kScoreThreads, C::kSmemBytes>>>(a);
  cudaEventRecord(e1);
  CK(cudaDeviceSynchronize());
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  printf("B=%d R=%d grid=%d CTAs: %.2f us per launch\n", B, R, nblk * 2 * R, ms / 20 * 1e3);
  std::vector<long long> h(64);
  CK(cudaMemcpy(h.data(), dbg, 64 * 8, cudaMemcpyDeviceToHost));
  long long t0 = h[0];
  for (int i = 0; i < 64; ++i) if (h[i]) printf("  stamp %2d : +%lld cyc\n", i, h[i] - t0);
  return 0;
}
