// umma_probe.cu — developer tool (not part of the product library): checks, on a real B200, that the
// shared-memory descriptors in csrc/sm100.cuh describe the "tile image" layout correctly for the three
// operand forms the fused training kernel uses:
//   (1) S  = U  * V^T   A K-major,  B K-major
//   (2) dU = G  * V     A K-major,  B MN-major
//   (3) dV = G^T* U     A MN-major, B MN-major
// Inputs are small integers / 8 so every product and sum is exact in fp32: expected max error is 0.
// usage: umma_probe [d] [lbo_mn] [sbo_mn] [lbo_k] [sbo_k]
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include <cuda_runtime.h>
#include "../nncf_b200/csrc/sm100.cuh"

using namespace nncf;

#define CK(x)                                                                          \
  do {                                                                                 \
    cudaError_t e_ = (x);                                                              \
    if (e_ != cudaSuccess) {                                                           \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);  \
      return 2;                                                                        \
    }                                                                                  \
  } while (0)

__device__ __forceinline__ float gmap(float s) {
  int q = static_cast<int>(s * 64.0f);
  int m = ((q % 7) + 7) % 7 - 3;
  return static_cast<float>(m);
}

__global__ void __launch_bounds__(128, 1)
probe_kernel(const uint8_t* __restrict__ Uimg, const uint8_t* __restrict__ Vimg, int d, float* __restrict__ S,
             float* __restrict__ dU, float* __restrict__ dV, uint32_t lbo_mn, uint32_t sbo_mn, uint32_t lbo_k,
             uint32_t sbo_k) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int nsub = d / 64;
  uint8_t* sU = smem;
  uint8_t* sV = sU + nsub * kSubBytes;
  uint8_t* sG = sV + nsub * kSubBytes;   // 2 sub-tiles
  __shared__ uint64_t bars[3];
  __shared__ uint32_t tmem_slot;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_init(&bars[2], 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc(&tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;

  if (tid == 0) {
    mbar_expect_tx(&bars[0], 2u * nsub * kSubBytes);
    for (int s = 0; s < nsub; ++s) {
      bulk_g2s(sU + s * kSubBytes, Uimg + (size_t)s * kSubBytes, kSubBytes, &bars[0]);
      bulk_g2s(sV + s * kSubBytes, Vimg + (size_t)s * kSubBytes, kSubBytes, &bars[0]);
    }
  }
  mbar_wait(&bars[0], 0);

  // (1) S = U V^T
  if (tid == 0) {
    tc_fence_after();
    const uint32_t idesc = make_idesc_bf16(128, 128, 0, 0);
    for (int k = 0; k < d / 16; ++k) {
      int sub = k / 4, kk = k % 4;
      uint64_t ad = make_smem_desc(smem_u32(sU + sub * kSubBytes) + kk * 32, lbo_k, sbo_k);
      uint64_t bd = make_smem_desc(smem_u32(sV + sub * kSubBytes) + kk * 32, lbo_k, sbo_k);
      umma_bf16(tmem + 0, ad, bd, idesc, k > 0);
    }
    umma_commit(&bars[1]);
  }
  mbar_wait(&bars[1], 0);
  tc_fence_after();

  // epilogue 1: S -> global, G -> smem tile image
  {
    const int row = warp * 32 + lane;
    for (int c0 = 0; c0 < 128; c0 += 32) {
      float v[32];
      tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
      tmem_ld_wait();
      for (int c = 0; c < 32; ++c) S[row * 128 + c0 + c] = v[c];
      for (int c = 0; c < 32; c += 2) {
        int col = c0 + c;
        uint32_t packed = pack_bf16x2(gmap(v[c]), gmap(v[c + 1]));
        uint8_t* base = sG + (col / 64) * kSubBytes;
        *reinterpret_cast<uint32_t*>(base + sw128_offset(row, col % 64)) = packed;
      }
    }
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();

  if (tid == 0) {
    tc_fence_after();
    // (2) dU = G V : A = G K-major (M=i, K=j), B = V MN-major (N=d, K=j)
    {
      const uint32_t idesc = make_idesc_bf16(128, d, 0, 1);
      for (int k = 0; k < 8; ++k) {   // K = 128 (j), 16 per MMA
        int sub = k / 4, kk = k % 4;
        uint64_t ad = make_smem_desc(smem_u32(sG + sub * kSubBytes) + kk * 32, lbo_k, sbo_k);
        // B: rows = k (j index), 16 rows per MMA => 16*128 bytes
        uint64_t bd = make_smem_desc(smem_u32(sV) + k * 16 * 128, lbo_mn, sbo_mn);
        umma_bf16(tmem + 128, ad, bd, idesc, k > 0);
      }
    }
    umma_commit(&bars[2]);
  }
  mbar_wait(&bars[2], 0);
  tc_fence_after();
  {
    const int row = warp * 32 + lane;
    for (int c0 = 0; c0 < d; c0 += 32) {
      float v[32];
      tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + 128 + c0, v);
      tmem_ld_wait();
      for (int c = 0; c < 32; ++c) dU[row * d + c0 + c] = v[c];
    }
  }
  tc_fence_before();
  __syncthreads();
  if (tid == 0) {
    tc_fence_after();
    // (3) dV = G^T U : A = G MN-major (M=j, K=i), B = U MN-major (N=d, K=i)
    {
      const uint32_t idesc = make_idesc_bf16(128, d, 1, 1);
      for (int k = 0; k < 8; ++k) {
        uint64_t ad = make_smem_desc(smem_u32(sG) + k * 16 * 128, lbo_mn, sbo_mn);
        uint64_t bd = make_smem_desc(smem_u32(sU) + k * 16 * 128, lbo_mn, sbo_mn);
        umma_bf16(tmem + 128, ad, bd, idesc, k > 0);
      }
    }
    umma_commit(&bars[1]);
  }
  mbar_wait(&bars[1], 1);
  tc_fence_after();
  {
    const int row = warp * 32 + lane;
    for (int c0 = 0; c0 < d; c0 += 32) {
      float v[32];
      tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + 128 + c0, v);
      tmem_ld_wait();
      for (int c = 0; c < 32; ++c) dV[row * d + c0 + c] = v[c];
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

static float hgmap(float s) {
  int q = static_cast<int>(s * 64.0f);
  int m = ((q % 7) + 7) % 7 - 3;
  return static_cast<float>(m);
}

int main(int argc, char** argv) {
  int d = argc > 1 ? atoi(argv[1]) : 128;
  uint32_t lbo_mn = argc > 2 ? atoi(argv[2]) : kSubBytes;
  uint32_t sbo_mn = argc > 3 ? atoi(argv[3]) : 1024;
  uint32_t lbo_k = argc > 4 ? atoi(argv[4]) : 16;
  uint32_t sbo_k = argc > 5 ? atoi(argv[5]) : 1024;
  if (d % 64 != 0 || d < 64 || d > 256) { printf("bad d\n"); return 1; }
  const int nsub = d / 64;
  std::vector<float> U(128 * d), V(128 * d);
  uint32_t st = 12345u;
  auto rnd = [&]() { st = st * 1664525u + 1013904223u; return (int)((st >> 16) % 9) - 4; };
  for (auto& x : U) x = rnd() / 8.0f;
  for (auto& x : V) x = rnd() / 8.0f;
  std::vector<uint8_t> Uimg(nsub * kSubBytes), Vimg(nsub * kSubBytes);
  for (int r = 0; r < 128; ++r)
    for (int c = 0; c < d; ++c) {
      __nv_bfloat16 hu = __float2bfloat16(U[r * d + c]), hv = __float2bfloat16(V[r * d + c]);
      size_t off = (size_t)(c / 64) * kSubBytes + sw128_offset(r, c % 64);
      memcpy(&Uimg[off], &hu, 2);
      memcpy(&Vimg[off], &hv, 2);
    }
  // CPU reference
  std::vector<float> S(128 * 128), G(128 * 128), dU(128 * d, 0.f), dV(128 * d, 0.f);
  for (int i = 0; i < 128; ++i)
    for (int j = 0; j < 128; ++j) {
      float a = 0;
      for (int k = 0; k < d; ++k) a += U[i * d + k] * V[j * d + k];
      S[i * 128 + j] = a;
      G[i * 128 + j] = hgmap(a);
    }
  for (int i = 0; i < 128; ++i)
    for (int j = 0; j < 128; ++j)
      for (int k = 0; k < d; ++k) {
        dU[i * d + k] += G[i * 128 + j] * V[j * d + k];
        dV[j * d + k] += G[i * 128 + j] * U[i * d + k];
      }
  uint8_t *dUimg, *dVimg;
  float *gS, *gdU, *gdV;
  CK(cudaMalloc(&dUimg, Uimg.size()));
  CK(cudaMalloc(&dVimg, Vimg.size()));
  CK(cudaMalloc(&gS, S.size() * 4));
  CK(cudaMalloc(&gdU, dU.size() * 4));
  CK(cudaMalloc(&gdV, dV.size() * 4));
  CK(cudaMemcpy(dUimg, Uimg.data(), Uimg.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dVimg, Vimg.data(), Vimg.size(), cudaMemcpyHostToDevice));
  CK(cudaMemset(gS, 0xff, S.size() * 4));
  CK(cudaMemset(gdU, 0xff, dU.size() * 4));
  CK(cudaMemset(gdV, 0xff, dV.size() * 4));
  size_t smem = (size_t)(2 * nsub + 2) * kSubBytes + 1024;
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  probe_kernel<<<1, 128, smem>>>(dUimg, dVimg, d, gS, gdU, gdV, lbo_mn, sbo_mn, lbo_k, sbo_k);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  std::vector<float> hS(S.size()), hdU(dU.size()), hdV(dV.size());
  CK(cudaMemcpy(hS.data(), gS, S.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hdU.data(), gdU, dU.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hdV.data(), gdV, dV.size() * 4, cudaMemcpyDeviceToHost));
  auto maxerr = [](const std::vector<float>& a, const std::vector<float>& b) {
    double m = 0;
    for (size_t i = 0; i < a.size(); ++i) {
      double e = std::fabs((double)a[i] - (double)b[i]);
      if (!(e == e)) e = 1e30;
      if (e > m) m = e;
    }
    return m;
  };
  double eS = maxerr(S, hS), eU = maxerr(dU, hdU), eV = maxerr(dV, hdV);
  printf("probe d=%d lbo_mn=%u sbo_mn=%u lbo_k=%u sbo_k=%u : errS=%g errdU=%g errdV=%g %s\n", d, lbo_mn, sbo_mn,
         lbo_k, sbo_k, eS, eU, eV, (eS == 0 && eU == 0 && eV == 0) ? "ALL_OK" : "MISMATCH");
  return 0;
}
