// Experiment 2: what slows tcgen05.mma below its stand-alone rate? mode bits:
//   1 = tcgen05.commit to an mbarrier after every 8 MMAs      2 = A descriptor: SBO 1280, start at an odd 128 B row
//   4 = alternate between two accumulators every 4 MMAs        8 = a second thread streams 16 KB TMA boxes into smem
//  16 = B walks over 8 different tiles (as a weight ring does)
#include <cstdio>
#include <vector>
#include "../../pytorch-camvid_b200/csrc/common.cuh"
#include "../../pytorch-camvid_b200/csrc/sm100.cuh"
#include "../../pytorch-camvid_b200/csrc/tma_host.h"
using namespace cvb;

template <int N>
__global__ void __launch_bounds__(128, 1) rate(const __grid_constant__ CUtensorMap tm2d, long long* out, int iters, int mode) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  // [0,64K) A tiles, [64K,128K) B tiles, [128K,192K) TMA landing ring
  __shared__ uint64_t mbar, cbar, tbar[4];
  __shared__ uint32_t slot;
  __shared__ volatile int stop;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 128 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    mbar_init(&mbar, 1); mbar_init(&cbar, 1);
    for (int i = 0; i < 4; ++i) mbar_init(&tbar[i], 1);
    fence_mbar_init();
    stop = 0;
  }
  if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = slot;
  if (threadIdx.x == 0) {
    constexpr uint32_t idesc = idesc_bf16_f32(128, N, false, false);
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 65536);
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      uint32_t aa = a0 + (i & 1) * 16384;
      uint32_t sbo = 1024;
      if (mode & 2) { aa += ((i % 9) * 3 + 1) * 128; sbo = 1280; }
      const uint32_t bb = b0 + ((mode & 16) ? (i & 7) * (N * 128 > 8192 ? 8192 : N * 128) : 0);
      const uint32_t d = tm + ((mode & 4) ? (i & 1) * N : 0);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        umma_bf16(d, smem_desc_sw128(aa + k * 32, 16, sbo), smem_desc_sw128(bb + k * 32, 16, 1024), idesc, 1u);
      if ((mode & 1) && (i & 1)) umma_commit(&cbar);
    }
    umma_commit(&mbar);
    mbar_wait(&mbar, 0);
    long long t1 = clock64();
    stop = 1;
    if (blockIdx.x == 0) out[0] = t1 - t0;
  } else if (threadIdx.x == 32 && (mode & 8)) {
    uint32_t ph[4] = {0, 0, 0, 0};
    int n = 0;
    for (int i = 0; i < 4; ++i) {
      mbar_expect_tx(&tbar[i], 16384);
      tma_load_2d(smem + 131072 + i * 16384, &tm2d, &tbar[i], 0, ((blockIdx.x * 37 + n++) % 64) * 128);
    }
    int s = 0;
    while (!stop) {
      mbar_wait(&tbar[s], ph[s]);
      ph[s] ^= 1;
      mbar_expect_tx(&tbar[s], 16384);
      tma_load_2d(smem + 131072 + s * 16384, &tm2d, &tbar[s], 0, ((blockIdx.x * 37 + n++) % 64) * 128);
      s = (s + 1) & 3;
    }
    for (int i = 0; i < 4; ++i) { mbar_wait(&tbar[s], ph[s]); ph[s] ^= 1; s = (s + 1) & 3; }
    if (blockIdx.x == 0) out[1] = n;
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tm, 512); }
}

template <int N>
void run(const CUtensorMap& tmap, long long* d, int grid, int mode) {
  cudaFuncSetAttribute(rate<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int iters = 2000;
  for (int rep = 0; rep < 2; ++rep) rate<N><<<grid, 128, 200 * 1024>>>(tmap, d, iters, mode);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[2] = {0, 0};
  cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  printf("N=%3d mode=%2d: %s  %.1f clk per MMA (ideal %d)  tma boxes %lld (%.1f B/clk)\n", N, mode, cudaGetErrorString(e),
         double(h[0]) / (iters * 4), N / 2, (mode & 8) ? h[1] : 0, (mode & 8) ? double(h[1]) * 16384 / h[0] : 0.0);
}

int main() {
  long long* d;
  cudaMalloc(&d, 16);
  void* g;
  cudaMalloc(&g, 64 * 128 * 128);
  cudaMemset(g, 0, 64 * 128 * 128);
  CUtensorMap tmap;
  if (make_mat_tmap(&tmap, g, 64 * 128, 64, 128)) { printf("tmap: %s\n", cvb_last_error()); return 1; }
  for (int mode : {0, 1, 2, 4, 8, 16, 9, 31}) {
    run<64>(tmap, d, 148, mode);
    run<128>(tmap, d, 148, mode);
    run<256>(tmap, d, 148, mode);
  }
  return 0;
}
