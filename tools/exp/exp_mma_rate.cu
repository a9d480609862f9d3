// Experiment: tcgen05.mma issue/execute rate for M=128, N in {64,128,256}, K=16 bf16, operands resident in shared
// memory (SS) or A in tensor memory (TS). One CTA per SM, no TMA in the loop.
#include <cstdio>
#include <vector>
#include "../../pytorch-camvid_b200/csrc/common.cuh"
#include "../../pytorch-camvid_b200/csrc/sm100.cuh"
using namespace cvb;

__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc),
      "r"(idesc), "r"(acc) : "memory");
}

template <int N, int TS>
__global__ void __launch_bounds__(128, 1) rate(long long* out, int iters, int a_rows_stride, int mn) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t mbar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (64 + 32) * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(&mbar, 1); fence_mbar_init(); }
  if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = mn ? idesc_bf16_f32(128, N, true, true) : idesc_bf16_f32(128, N, false, false);
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 65536);
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      // 4 k-steps over a 64-wide K tile; A walks over 4 different 16 KB tiles, B over its 64-wide tile
      const uint32_t aa = a0 + (i & 3) * 16384 + (i & 1) * a_rows_stride * 128;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (TS) umma_ts(tm, tm + 256 + k * 8, smem_desc_sw128(b0 + k * 32, 16, 1024), idesc, 1u);
        else if (mn) umma_bf16(tm, smem_desc_sw128(aa + k * 2048, 8192, 1024), smem_desc_sw128(b0 + k * 2048, 8192, 1024), idesc, 1u);
        else umma_bf16(tm, smem_desc_sw128(aa + k * 32, 16, 1024), smem_desc_sw128(b0 + k * 32, 16, 1024), idesc, 1u);
      }
    }
    umma_commit(&mbar);
    mbar_wait(&mbar, 0);
    long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tm, 512); }
}

template <int N, int TS>
void run(const char* tag, long long* d, int grid, int mn = 0) {
  cudaFuncSetAttribute(rate<N, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int iters = 2000;
  for (int rep = 0; rep < 2; ++rep) rate<N, TS><<<grid, 128, 100 * 1024>>>(d, iters, 0, mn);
  cudaError_t e = cudaDeviceSynchronize();
  long long h = 0;
  cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  printf("%s mn=%d N=%d grid=%d: %s  %.1f clk per MMA (ideal %d)\n", tag, mn, N, grid, cudaGetErrorString(e),
         double(h) / (iters * 4), N / 2);
}

int main() {
  long long* d;
  cudaMalloc(&d, 8);
  for (int grid : {148}) {
    run<64, 0>("SS", d, grid, 1);
    run<128, 0>("SS", d, grid, 1);
    run<256, 0>("SS", d, grid, 1);
    run<64, 0>("SS", d, grid);
    run<128, 0>("SS", d, grid);
    run<256, 0>("SS", d, grid);
    run<64, 1>("TS", d, grid);
    run<128, 1>("TS", d, grid);
    run<256, 1>("TS", d, grid);
  }
  return 0;
}
