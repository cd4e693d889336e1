// occ_probe.cu — developer tool: which resource decides whether TWO 384-thread CTAs that use tcgen05 / TMEM are resident on
// one SM?  Prints the occupancy API's answer for four kernel variants and measures real co-residency (CTAs record %smid and
// globaltimer around a 20 us spin; two CTAs of one SM overlapping in time = resident together).
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>
extern __shared__ __align__(1024) unsigned char sm[];
__device__ __forceinline__ uint64_t gtime() { uint64_t t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ uint32_t smid() { uint32_t t; asm volatile("mov.u32 %0, %%smid;" : "=r"(t)); return t; }
template <int REGS, int TMEM_COLS>
__global__ void __launch_bounds__(384, 2) k(uint64_t* o, float* sink) {
  __shared__ uint32_t slot;
  float acc[REGS];
#pragma unroll
  for (int i = 0; i < REGS; ++i) acc[i] = threadIdx.x * 0.5f + i;
  uint32_t taddr = 0;
  if (TMEM_COLS > 0) {
    if (threadIdx.x < 32) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)), "r"(TMEM_COLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    __syncthreads();
    taddr = slot;
  }
  const uint64_t t0 = gtime();
  while (gtime() - t0 < 20000) {
#pragma unroll
    for (int i = 0; i < REGS; ++i) acc[i] = acc[i] * 1.0001f + 0.5f;
  }
  const uint64_t t1 = gtime();
  float s = 0; 
#pragma unroll
  for (int i = 0; i < REGS; ++i) s += acc[i];
  if (s == 12345.678f) sink[0] = s + sm[threadIdx.x];
  if (threadIdx.x == 0) { o[blockIdx.x * 3] = smid(); o[blockIdx.x * 3 + 1] = t0; o[blockIdx.x * 3 + 2] = t1; }
  __syncthreads();
  if (TMEM_COLS > 0 && threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(TMEM_COLS) : "memory");
}
template <int REGS, int TMEM_COLS>
void run(const char* name, int smem) {
  auto kern = k<REGS, TMEM_COLS>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
  cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, kern);
  int nb = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, 384, smem);
  const int grid = 296;
  uint64_t* d; float* sink; cudaMalloc(&d, grid * 24); cudaMalloc(&sink, 4);
  kern<<<grid, 384, smem>>>(d, sink);
  cudaError_t e = cudaDeviceSynchronize();
  std::vector<uint64_t> h(grid * 3); cudaMemcpy(h.data(), d, grid * 24, cudaMemcpyDeviceToHost);
  int overlap = 0; uint64_t tmin = ~0ull, tmax = 0;
  for (int i = 0; i < grid; ++i) { if (h[i*3+1] < tmin) tmin = h[i*3+1]; if (h[i*3+2] > tmax) tmax = h[i*3+2];
    for (int j = i + 1; j < grid; ++j) if (h[i*3] == h[j*3] && h[i*3+1] < h[j*3+2] && h[j*3+1] < h[i*3+2]) ++overlap; }
  printf("%-28s regs %3d dyn smem %6d: occupancy API %d, co-resident CTA pairs %d, wall %.1f us (%s)\n", name, fa.numRegs, smem, nb, overlap,
         (tmax - tmin) * 1e-3, cudaGetErrorString(e));
  cudaFree(d); cudaFree(sink);
}
int main() {
  run<8, 0>("few regs, no TMEM", 114944);
  run<56, 0>("~80 regs, no TMEM", 114944);
  run<8, 256>("few regs, TMEM 256", 114944);
  run<56, 256>("~80 regs, TMEM 256", 114944);
  run<56, 256>("~80 regs, TMEM 256, 64K", 65536);
  return 0;
}
