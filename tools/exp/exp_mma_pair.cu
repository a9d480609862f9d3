// Experiment: issue rate of tcgen05.mma.cta_group::2 (a CTA pair runs ONE M = 256 MMA; each SM reads its own 128 rows of A
// and only HALF of B's N rows from its shared memory) against the single-CTA M = 128 MMA of the same per-SM work.
// Question behind it: the weight-gradient and cout <= 128 kernels are bound by operand rows read from
// shared memory (~0.44 clk per 32-byte row): does the pair form lift that bound?   K-major operands, 128-byte swizzle.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o exp_mma_pair exp_mma_pair.cu -lcuda && ./exp_mma_pair
#include <cstdio>
#include "../../pytorch-camvid_b200/csrc/common.cuh"
#include "../../pytorch-camvid_b200/csrc/sm100.cuh"
using namespace cvb;

// Defined by api.cu in the library; the experiment links without it.
namespace cvb { void set_error(const char*, ...) {} int sm_count() { return 148; } }

// walk != 0: consecutive iterations (4 MMAs = one 64-wide K chunk) read DIFFERENT A and B tiles, as a pipeline of TMA
// stages does; walk == 0: the same B tile every time (what an operand cache, if there is one, would love).
template <int N, bool PAIR, bool MN>
__global__ void __launch_bounds__(128, 1) rate(long long* out, int iters, int walk, int fill, const uint8_t* src) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t mbar, tbar[4];
  __shared__ uint32_t slot;
  __shared__ volatile int stop;
  const int warp = threadIdx.x >> 5;
  const uint32_t rank = PAIR ? cluster_ctarank() : 0;
  // [0, 32K) two A tiles, [32K, 96K) two B tiles, [96K, 160K) landing ring of the bulk-copy streamer
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;  // bf16 1.0
  if (threadIdx.x == 0) {
    mbar_init(&mbar, 1);
    for (int i = 0; i < 4; ++i) mbar_init(&tbar[i], 1);
    fence_mbar_init();
    stop = 0;
  }
  if (warp == 0) {
    if (PAIR) { tmem_alloc_pair(&slot, 512); tmem_relinquish_pair(); } else { tmem_alloc(&slot, 512); tmem_relinquish(); }
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();
  tc_fence_after();
  const uint32_t tm = slot;
  if (warp == 1) {
    if (!PAIR || rank == 0) {
      // MN: both operands MN-major as in the weight-gradient kernels (a K = 16 slice = 16 rows of 128 B per 64-wide atom,
      // atoms 8 KB apart, K groups of 8 rows 1 KB apart); else K-major as in the forward kernels
      constexpr uint32_t idesc = idesc_bf16_f32(PAIR ? 256 : 128, N, MN, MN);
      constexpr uint32_t hi = desc_hi_sw128(1024);
      const uint32_t a0 = desc_lo(smem_u32(smem), MN ? 8192 : 16), b0 = desc_lo(smem_u32(smem + 32768), MN ? 8192 : 16);
      long long t0 = clock64();
      for (int i = 0; i < iters; ++i) {
        constexpr uint32_t b_tile = (PAIR ? N / 2 : N) * 128;  // bytes of B one SM holds per K chunk
        const uint32_t aa = a0 + (i & 1) * (16384 >> 4);
        const uint32_t bb = b0 + (walk ? (i & 1) * (b_tile >> 4) : 0);
        const uint32_t d = tm + (i & 1) * N;
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            if (PAIR) umma_bf16_lohi_pair(d, aa + (MN ? 128 : 2) * k, hi, bb + (MN ? 128 : 2) * k, hi, idesc, 1u);
            else umma_bf16_lohi(d, aa + (MN ? 128 : 2) * k, hi, bb + (MN ? 128 : 2) * k, hi, idesc, 1u);
          }
        }
        __syncwarp();
      }
      if (elect_one()) {
        if (PAIR) umma_commit_pair(&mbar); else umma_commit(&mbar);
      }
      __syncwarp();
      mbar_wait(&mbar, 0);
      long long t1 = clock64();
      stop = 1;
      if (blockIdx.x == 0 && (threadIdx.x & 31) == 0) out[0] = t1 - t0;
    } else {
      mbar_wait(&mbar, 0);  // the leader's commit arrives on both CTAs' barriers
    }
  }
  if (warp == 2 && (threadIdx.x & 31) == 0 && fill > 0 && (!PAIR || rank == 0)) {
    // streams 16 KB bulk copies global -> shared memory, `fill` of them in flight, while the MMAs run: what the TMA producer
    // of a real kernel does to the shared-memory write port
    uint32_t ph[4] = {0, 0, 0, 0};
    long long n = 0;
    auto issue = [&](int s_) {
      mbar_expect_tx(&tbar[s_], 16384);
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                       smem_u32(smem + 98304 + s_ * 16384)),
                   "l"(src + ((blockIdx.x * 37 + n) % 512) * 16384), "r"(16384), "r"(smem_u32(&tbar[s_]))
                   : "memory");
      ++n;
    };
    for (int i = 0; i < fill; ++i) issue(i);
    int s_ = 0;
    while (!stop) {
      mbar_wait(&tbar[s_], ph[s_]);
      ph[s_] ^= 1;
      issue(s_);
      s_ = (s_ + 1) % fill;
    }
    for (int i = 0; i < fill; ++i) {
      mbar_wait(&tbar[s_], ph[s_]);
      ph[s_] ^= 1;
      s_ = (s_ + 1) % fill;
    }
    if (blockIdx.x == 0) out[1] = n;
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();
  if (warp == 0) {
    tc_fence_after();
    if (PAIR) tmem_dealloc_pair(tm, 512); else tmem_dealloc(tm, 512);
  }
}

template <int N, bool PAIR, bool MN>
void run(long long* d, int grid, int walk, int fill, const uint8_t* src) {
  auto kern = rate<N, PAIR, MN>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int iters = 2000;
  cudaError_t e = cudaSuccess;
  for (int rep = 0; rep < 2 && e == cudaSuccess; ++rep) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(128);
    cfg.dynamicSmemBytes = 200 * 1024;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = PAIR ? 2 : 1;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    e = cudaLaunchKernelEx(&cfg, kern, d, iters, walk, fill, src);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
  }
  long long hh[2] = {0, 0};
  cudaMemcpy(hh, d, 16, cudaMemcpyDeviceToHost);
  const long long h = hh[0];
  const double clk = double(h) / (iters * 4);
  // per-SM MACs of one instruction: 128 x N x 16 in both forms (the pair's M = 256 is split over two SMs)
  printf("%s %s N=%3d %s fill %d (%.0f B/clk written by bulk copies): %s  %.1f clk per MMA; per SM: tensor pipe needs %d, operand rows A 128 + B %d -> model %.0f clk\n",
         MN ? "MN-major" : "K-major ", PAIR ? "pair M=256" : "single M=128", N, walk ? "walking tiles" : "same B tile  ", fill, fill ? double(hh[1]) * 16384 / double(h) : 0.0, cudaGetErrorString(e), clk, N / 2, PAIR ? N / 2 : N,
         0.44 * (128 + (PAIR ? N / 2 : N)));
}

int main() {
  long long* d;
  cudaMalloc(&d, 16);
  uint8_t* src;
  cudaMalloc(&src, 512 * 16384);
  cudaMemset(src, 0x3c, 512 * 16384);
  for (int fill : {0, 4}) {
    run<64, false, false>(d, 148, 1, fill, src);
    run<128, false, false>(d, 148, 1, fill, src);
    run<256, false, false>(d, 148, 1, fill, src);
    run<128, true, false>(d, 148, 1, fill, src);
    run<256, true, false>(d, 148, 1, fill, src);
    run<64, false, true>(d, 148, 1, fill, src);
    run<128, false, true>(d, 148, 1, fill, src);
    run<192, false, true>(d, 148, 1, fill, src);
    run<256, false, true>(d, 148, 1, fill, src);
    run<128, true, true>(d, 148, 1, fill, src);
    run<256, true, true>(d, 148, 1, fill, src);
  }
  return 0;
}
