// Experiment: do tcgen05 shared-memory descriptors accept start addresses that are 128-byte-row aligned but not
// 1024-byte (swizzle atom) aligned?  Decides whether a halo'd activation patch can be loaded ONCE and addressed by
// nine shifted descriptors. Build: nvcc -gencode arch=compute_100a,code=sm_100a -o exp_desc exp_desc.cu ../../pytorch-camvid_b200/csrc/api.cu
#include <vector>
#include <cstdio>
#include <cstdlib>
#include "../../pytorch-camvid_b200/csrc/common.cuh"
#include "../../pytorch-camvid_b200/csrc/sm100.cuh"
#include "../../pytorch-camvid_b200/csrc/tma_host.h"
using namespace cvb;

__device__ __forceinline__ uint64_t desc_ex(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t base_off) {
  uint64_t d = smem_desc_sw128(saddr, lbo, sbo);
  d |= static_cast<uint64_t>(base_off & 7) << 49;
  return d;
}

// mode 0: K-major A rows [r0, r0+128) (SBO 1024);  mode 1: K-major A, 8-row groups 16 rows apart (SBO 2048)
// mode 2: MN-major A and B, K rows [r0, r0+16)
__global__ void __launch_bounds__(128, 1) k(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                           int mode, int r0, int use_base, float* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;              // 512 rows x 128 B = 64 KB
  uint8_t* sB = smem + 65536;      // 128 rows x 128 B = 16 KB
  __shared__ uint64_t bar, mbar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init(&mbar, 1); fence_mbar_init(); }
  if (warp == 0) { tmem_alloc(&slot, 128); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = slot;
  if (threadIdx.x == 0) {
    mbar_expect_tx(&bar, 65536 + 16384);
    tma_load_2d(sA, &tmA, &bar, 0, 0);
    tma_load_2d(sA + 32768, &tmA, &bar, 0, 256);
    tma_load_2d(sB, &tmB, &bar, 0, 0);
    mbar_wait(&bar, 0);
    tc_fence_after();
    const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
    if (mode == 0 || mode == 1 || mode == 3) {
      constexpr uint32_t idesc = idesc_bf16_f32(128, 64, false, false);
      const uint32_t sbo = mode == 0 ? 1024 : (mode == 1 ? 2048 : 1280);
      for (int kk = 0; kk < 4; ++kk) {
        uint32_t sa = a0 + r0 * 128 + kk * 32;
        uint32_t bo = use_base ? ((sa >> 7) & 7) : 0;
        umma_bf16(tm, desc_ex(sa, 16, sbo, bo), smem_desc_sw128(b0 + kk * 32, 16, 1024), idesc, kk ? 1u : 0u);
      }
    } else {
      // MN-major: A = 2 units (M 0..63 at sA, M 64..127 at sA+32768: the second 256-row load), K rows r0..r0+15
      constexpr uint32_t idesc = idesc_bf16_f32(128, 64, true, true);
      uint32_t sa = a0 + r0 * 128, sb = b0 + r0 * 128;
      uint32_t boa = use_base ? ((sa >> 7) & 7) : 0, bob = use_base ? ((sb >> 7) & 7) : 0;
      umma_bf16(tm, desc_ex(sa, 32768, 1024, boa), desc_ex(sb, 8192, 1024, bob), idesc, 0u);
    }
    umma_commit(&mbar);
    mbar_wait(&mbar, 0);
  }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  for (int c0 = 0; c0 < 64; c0 += 32) {
    uint32_t r[32];
    tmem_ld_32x32(tm + (static_cast<uint32_t>(warp * 32) << 16) + c0, r);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) out[(warp * 32 + lane) * 64 + c0 + j] = __uint_as_float(r[j]);
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tm, 128); }
}

static float bf(float v) { return __bfloat162float(__float2bfloat16(v)); }

int main() {
  const int RA = 512, RB = 128;
  std::vector<__nv_bfloat16> hA(RA * 64), hB(RB * 64);
  std::vector<float> fA(RA * 64), fB(RB * 64);
  srand(1);
  for (int i = 0; i < RA * 64; ++i) { fA[i] = bf((rand() % 17 - 8) / 8.f); hA[i] = __float2bfloat16(fA[i]); }
  for (int i = 0; i < RB * 64; ++i) { fB[i] = bf((rand() % 13 - 6) / 4.f); hB[i] = __float2bfloat16(fB[i]); }
  __nv_bfloat16 *dA, *dB; float* dOut;
  cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dOut, 128 * 64 * 4);
  cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
  CUtensorMap tmA, tmB;
  if (make_mat_tmap(&tmA, dA, RA, 64, 256) || make_mat_tmap(&tmB, dB, RB, 64, 128)) { printf("tmap failed: %s\n", cvb_last_error()); return 1; }
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 + 16384 + 2048);
  std::vector<float> out(128 * 64);
  for (int mode = 0; mode < 4; ++mode)
    for (int use_base = 0; use_base < 2; ++use_base) {
      printf("mode %d use_base %d:", mode, use_base);
      for (int r0 : {0, 1, 2, 3, 7, 8, 9, 10, 15, 16, 17, 33}) {
        if (mode == 2 && r0 + 16 > 128) continue;
        cudaMemset(dOut, 0, 128 * 64 * 4);
        k<<<1, 128, 65536 + 16384 + 2048>>>(tmA, tmB, mode, r0, use_base, dOut);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf(" r0=%d CUDA error %s\n", r0, cudaGetErrorString(e)); return 2; }
        cudaMemcpy(out.data(), dOut, out.size() * 4, cudaMemcpyDeviceToHost);
        double maxerr = 0;
        for (int m = 0; m < 128; ++m)
          for (int n = 0; n < 64; ++n) {
            double ref = 0;
            if (mode == 0) for (int kk = 0; kk < 64; ++kk) ref += fA[(r0 + m) * 64 + kk] * fB[n * 64 + kk];
            else if (mode == 1) for (int kk = 0; kk < 64; ++kk) ref += fA[(r0 + (m / 8) * 16 + (m % 8)) * 64 + kk] * fB[n * 64 + kk];
            else if (mode == 3) for (int kk = 0; kk < 64; ++kk) ref += fA[(r0 + (m / 8) * 10 + (m % 8)) * 64 + kk] * fB[n * 64 + kk];
            else for (int kk = 0; kk < 16; ++kk) ref += fA[((m / 64) * 256 + r0 + kk) * 64 + (m % 64)] * fB[(r0 + kk) * 64 + n];
            double d = fabs(ref - out[m * 64 + n]);
            if (d > maxerr) maxerr = d;
          }
        printf(" r0=%d:%s", r0, maxerr < 1e-3 ? "OK" : "BAD");
      }
      printf("\n");
    }
  return 0;
}
