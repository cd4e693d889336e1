// mufu_bench.cu — developer tool: issue cost (cycles per warp instruction per SM sub-partition) of the special-function
// and packed-math instructions the score epilogue can be built from.  One CTA of 512 threads (4 warps per sub-partition),
// 8 independent dependency chains per thread, clock64 around 2048 iterations.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define OPS(X) X(0, "tanh.approx.f32") X(1, "ex2.approx.ftz.f32") X(2, "rcp.approx.ftz.f32") X(3, "lg2.approx.ftz.f32") \
  X(4, "tanh.approx.bf16x2") X(5, "ex2.approx.ftz.bf16x2") X(6, "fma.rn.bf16x2") X(7, "fma.rn.f32") X(8, "tanh.approx.f16x2") \
  X(9, "ex2.approx.f16x2") X(10, "rsqrt.approx.ftz.f32") X(11, "fma.rn.f16x2")
template <int OP>
__device__ __forceinline__ uint32_t op(uint32_t x) {
  uint32_t y;
  if (OP == 0) asm volatile("tanh.approx.f32 %0, %1;" : "=r"(y) : "r"(x));
  if (OP == 1) asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=r"(y) : "r"(x));
  if (OP == 2) asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=r"(y) : "r"(x));
  if (OP == 3) asm volatile("lg2.approx.ftz.f32 %0, %1;" : "=r"(y) : "r"(x));
  if (OP == 4) asm volatile("tanh.approx.bf16x2 %0, %1;" : "=r"(y) : "r"(x));
  if (OP == 5) asm volatile("ex2.approx.ftz.bf16x2 %0, %1;" : "=r"(y) : "r"(x));
  if (OP == 6) asm volatile("fma.rn.bf16x2 %0, %1, %1, %1;" : "=r"(y) : "r"(x));
  if (OP == 7) asm volatile("fma.rn.f32 %0, %1, %1, %1;" : "=r"(y) : "r"(x));
  if (OP == 8) asm volatile("tanh.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x));
  if (OP == 9) asm volatile("ex2.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x));
  if (OP == 10) asm volatile("rsqrt.approx.ftz.f32 %0, %1;" : "=r"(y) : "r"(x));
  if (OP == 11) asm volatile("fma.rn.f16x2 %0, %1, %1, %1;" : "=r"(y) : "r"(x));
  return y;
}
template <int OP>
__global__ void __launch_bounds__(512) k(uint32_t* out, long long* cyc) {
  uint32_t v[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = 0x3e003e00u + threadIdx.x * 8 + i;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < 2048; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = op<OP>(v[i]);
  }
  const long long t1 = clock64();
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s ^= v[i];
  out[threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
  uint32_t* o; long long* c; cudaMalloc(&o, 4096); cudaMalloc(&c, 8);
#define X(i, name) { k<i><<<1, 512>>>(o, c); k<i><<<1, 512>>>(o, c); cudaDeviceSynchronize(); long long h; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost); \
    printf("%-24s %6.2f cycles per warp instruction per sub-partition\n", name, (double)h / (2048.0 * 8 * 4)); }
  OPS(X)
  return 0;
}
