// 3x3 / stride 1 / pad 1 convolution (and its data gradient) as an implicit GEMM on tcgen05.
//   reference: nn.Conv2d(cin, cout, 3, padding=1) in models/unet.py:11 and models/segnet.py:8, and the input-gradient
//   half of aten::convolution_backward reached through loss.backward() (train.py:131).
//
// GEMM view:  D[M = pixels][N = cout] = sum over (tap, cin) of  A[pixel shifted by tap][cin] * B[cout][tap][cin]
//   * A tile = a TN x TH x TW patch of pixels (<= 128 rows) x 64 channels, fetched by ONE 4-D TMA box per (tap, cin
//     chunk). The box is shifted by the tap offset; TMA zero-fills whatever falls outside the image, which is the
//     convolution's zero padding -- no im2col buffer, no predication in the producer.
//   * B tile = BN x 64 slice of the packed weights [cout][tap][cin] (K-major), one 2-D TMA box.
//   * both land in 128B-swizzled shared memory and feed tcgen05.mma (M=128, N=BN, K=16 x4 per stage), fp32
//     accumulators live in TMEM, double-buffered so the epilogue of tile i overlaps the MMAs of tile i+1.
//   * persistent CTAs (one per SM), warp-specialised: warp0 = TMA producer, warp1 = MMA issuer, warp2 = TMEM
//     allocator, warps 4-7 = epilogue (TMEM -> registers -> bf16 NHWC stores, plus per-channel sum / sum-of-squares
//     for the BatchNorm batch statistics, or the folded eval-mode BN+ReLU).
#include <stdlib.h>

#include "common.cuh"
#include "sm100.cuh"
#include "tma_host.h"

namespace cvb {

struct FpropParams {
  int N, H, W;
  int cin_chunks, taps, cin_pad, cout_pad;
  int cout_store;   // channels of y that exist in memory (<= cout_pad; the transposed cout = 64 kernel stores only these)
  int ksteps_last;  // K = 16 steps of the last 64-channel chunk that hold real channels (x.c need not fill the chunk)
  int TW, TH, TN;
  int tiles_w, tiles_h, tiles_n, n_tiles, total_tiles;
  uint32_t a_bytes;
  __nv_bfloat16* y;
  long long ysn, ysh, ysw;
  float* stat_partials;
  const float* scale;
  const float* shift;
  int relu;
  // data-gradient mode: BatchNorm+ReLU backward statistics of the consuming block (cvb_conv_epilogue::bwd_*)
  const __nv_bfloat16* by;
  long long bysn, bysh, bysw;
  const float* bscale;
  const float* bshift;
  float* bpartials;
  int debug;  // development knobs (CVB_DEBUG env): bit 0 = skip the output stores, bit 1 = MMA-thread wait trace
  long long* trace;  // with debug bit 1: the stat_partials buffer reinterpreted (statistics are then not produced)
};

constexpr int kFpropThreads = 256;
constexpr int kABytes = 128 * 128;  // 128 pixel rows x 64 bf16
constexpr int kStatFloats = 4 * 2 * 1024;  // one private [sum | sumsq][cout_pad <= 1024] row per epilogue warp

template <int BN>
struct FpropCfg {
  static constexpr int kStages = BN == 256 ? 4 : (BN == 128 ? 6 : 8);
  static constexpr int kBBytes = BN * 128;
  static constexpr int kTmemCols = 2 * BN < 32 ? 32 : 2 * BN;  // double-buffered accumulator (power of two)
  static constexpr int kSmem = 1024 /*align slack*/ + kStages * (kABytes + kBBytes) + kStatFloats * 4 + 256;
};

// Column sums of a 32(lane) x 32(value) tile: after the exchange, lane L holds sum over lanes of v[L].
__device__ __forceinline__ float warp_column_sum(float (&v)[32], int lane) {
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    bool up = lane & 16;
    float send = up ? v[i] : v[i + 16];
    float keep = up ? v[i + 16] : v[i];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    bool up = lane & 8;
    float send = up ? v[i] : v[i + 8];
    float keep = up ? v[i + 8] : v[i];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    bool up = lane & 4;
    float send = up ? v[i] : v[i + 4];
    float keep = up ? v[i + 4] : v[i];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    bool up = lane & 2;
    float send = up ? v[i] : v[i + 2];
    float keep = up ? v[i + 2] : v[i];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  }
  {
    bool up = lane & 1;
    float send = up ? v[0] : v[1];
    float keep = up ? v[1] : v[0];
    v[0] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
  }
  return v[0];
}

// 32 accumulator columns of one output pixel: (eval: folded BatchNorm + ReLU) -> bf16 -> four 16-byte stores.
__device__ __forceinline__ void epilogue_store(const FpropParams& p, const uint32_t (&r)[32], __nv_bfloat16* dst, int ch,
                                               bool valid) {
  if (!valid || (p.debug & 1)) return;
  float v[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
  if (p.scale) {
    const float* sc = p.scale + ch;
    const float* sh = p.shift + ch;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      v[j] = fmaf(v[j], __ldg(sc + j), __ldg(sh + j));
      if (p.relu) v[j] = fmaxf(v[j], 0.f);
    }
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    uint4 o;
    o.x = pack_bf16x2(v[q * 8 + 0], v[q * 8 + 1]);
    o.y = pack_bf16x2(v[q * 8 + 2], v[q * 8 + 3]);
    o.z = pack_bf16x2(v[q * 8 + 4], v[q * 8 + 5]);
    o.w = pack_bf16x2(v[q * 8 + 6], v[q * 8 + 7]);
    *reinterpret_cast<uint4*>(dst + q * 8) = o;
  }
}

// Epilogue of one 32-row slab of an accumulator tile: TMEM -> registers -> (eval: folded BN + ReLU) -> bf16 NHWC store,
// plus per-channel sum / sum of squares of the fp32 accumulators for the BatchNorm batch statistics (train).
template <int BN>
__device__ __forceinline__ void epilogue_rows(const FpropParams& p, uint32_t taddr, __nv_bfloat16* dst, int ch0,
                                              bool valid, int lane, float* s_stats) {
#pragma unroll 1
  for (int c0 = 0; c0 < BN; c0 += 32) {
    uint32_t r[32];
    tmem_ld_32x32(taddr + c0, r);
    tmem_ld_wait();
    float v[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
    if (p.scale) {
      const float* sc = p.scale + ch0 + c0;
      const float* sh = p.shift + ch0 + c0;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        v[j] = fmaf(v[j], __ldg(sc + j), __ldg(sh + j));
        if (p.relu) v[j] = fmaxf(v[j], 0.f);
      }
    }
    if (valid && !(p.debug & 1)) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint4 o;
        o.x = pack_bf16x2(v[q * 8 + 0], v[q * 8 + 1]);
        o.y = pack_bf16x2(v[q * 8 + 2], v[q * 8 + 3]);
        o.z = pack_bf16x2(v[q * 8 + 4], v[q * 8 + 5]);
        o.w = pack_bf16x2(v[q * 8 + 6], v[q * 8 + 7]);
        *reinterpret_cast<uint4*>(dst + c0 + q * 8) = o;
      }
    }
    if (p.stat_partials) {
      float sq[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        v[j] = valid ? v[j] : 0.f;
        sq[j] = v[j] * v[j];
      }
      float s1 = warp_column_sum(v, lane);
      float s2 = warp_column_sum(sq, lane);
      // s_stats is this warp's private row and a lane always owns the same columns: plain read-modify-write
      s_stats[ch0 + c0 + lane] += s1;
      s_stats[p.cout_pad + ch0 + c0 + lane] += s2;
    }
  }
}

template <int BN>
__global__ void __launch_bounds__(kFpropThreads, 1)
conv_fprop_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const FpropParams p) {
  using Cfg = FpropCfg<BN>;
  constexpr int S = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = sA + S * kABytes;
  float* s_stats = reinterpret_cast<float*>(sB + S * Cfg::kBBytes);
  uint64_t* full = reinterpret_cast<uint64_t*>(s_stats + kStatFloats);
  uint64_t* empty = full + S;
  uint64_t* tfull = empty + S;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < S; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 4);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, Cfg::kTmemCols);
    tmem_relinquish();
  }
  for (int i = threadIdx.x; i < kStatFloats; i += kFpropThreads) s_stats[i] = 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int kb_total = p.cin_chunks * p.taps;

  if (warp == 0) {
    if (lane == 0) {
      // ------------------------------- TMA producer -------------------------------
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const int nt = tile % p.n_tiles;
        int mt = tile / p.n_tiles;
        const int w0 = (mt % p.tiles_w) * p.TW;
        mt /= p.tiles_w;
        const int h0 = (mt % p.tiles_h) * p.TH;
        const int n0 = (mt / p.tiles_h) * p.TN;
        for (int chunk = 0; chunk < p.cin_chunks; ++chunk) {
          for (int tap = 0; tap < p.taps; ++tap) {
            const int dr = p.taps == 9 ? tap / 3 - 1 : 0;
            const int ds = p.taps == 9 ? tap % 3 - 1 : 0;
            mbar_wait(&empty[stage], phase ^ 1);
            mbar_expect_tx(&full[stage], p.a_bytes + Cfg::kBBytes);
            tma_load_4d(sA + stage * kABytes, &tmA, &full[stage], chunk * 64, w0 + ds, h0 + dr, n0);
            tma_load_2d(sB + stage * Cfg::kBBytes, &tmB, &full[stage], tap * p.cin_pad + chunk * 64, nt * BN);
            if (++stage == S) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------- MMA issuer (whole warp, one elected lane issues) -------------------------------
    constexpr uint32_t idesc = idesc_bf16_f32(128, BN, false, false);
    constexpr uint32_t hi = desc_hi_sw128(1024);
    const uint32_t a_lo0 = desc_lo(smem_u32(sA), 16), b_lo0 = desc_lo(smem_u32(sB), 16);
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(&tempty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d = tmem_base + acc * BN;
      for (int kb = 0; kb < kb_total; ++kb) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        const uint32_t a_lo = a_lo0 + stage * (kABytes >> 4), b_lo = b_lo0 + stage * (Cfg::kBBytes >> 4);
        const int ks = kb >= kb_total - p.taps ? p.ksteps_last : 4;  // the producer walks chunks outermost: last chunk = last taps
        if (elect_one()) {
          umma_bf16_lohi(d, a_lo, hi, b_lo, hi, idesc, kb != 0 ? 1u : 0u);
          if (ks > 1) umma_bf16_lohi(d, a_lo + 2, hi, b_lo + 2, hi, idesc, 1u);
          if (ks > 2) umma_bf16_lohi(d, a_lo + 4, hi, b_lo + 4, hi, idesc, 1u);
          if (ks > 3) umma_bf16_lohi(d, a_lo + 6, hi, b_lo + 6, hi, idesc, 1u);
          umma_commit(&empty[stage]);
        }
        __syncwarp();
        if (++stage == S) {
          stage = 0;
          phase ^= 1;
        }
      }
      if (elect_one()) umma_commit(&tfull[acc]);
      __syncwarp();
    }
  } else if (warp >= 4) {
    // --------------------------------- epilogue -----------------------------------
    const int ew = warp - 4;
    const int row = ew * 32 + lane;
    const int thw = p.TH * p.TW;
    const int tn_ = row / thw;
    const int rem = row - tn_ * thw;
    const int ty_ = rem / p.TW;
    const int tx_ = rem - ty_ * p.TW;
    const bool row_in_box = row < p.TN * thw;
    // A single 64-wide channel tile (the im2col'd first layer): per-thread statistic accumulators, reduced across lanes
    // once per kernel (see the halo kernel); otherwise the per-tile shuffle reduction into this warp's shared-memory row.
    constexpr int kRegCh = BN == 64 ? 64 : 1;
    const bool reg_stats = BN == 64 && p.n_tiles == 1 && p.stat_partials != nullptr;
    float S[kRegCh], Q[kRegCh];
#pragma unroll
    for (int j = 0; j < kRegCh; ++j) S[j] = Q[j] = 0.f;
    float* my_stats = s_stats + ew * 2 * p.cout_pad;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int nt = tile % p.n_tiles;
      int mt = tile / p.n_tiles;
      const int w = (mt % p.tiles_w) * p.TW + tx_;
      mt /= p.tiles_w;
      const int h = (mt % p.tiles_h) * p.TH + ty_;
      const int n = (mt / p.tiles_h) * p.TN + tn_;
      const bool valid = row_in_box && n < p.N && h < p.H && w < p.W;
      __nv_bfloat16* dst = p.y + n * p.ysn + h * p.ysh + w * p.ysw + nt * BN;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + acc * BN;
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      if (BN == 64 && reg_stats) {
#pragma unroll
        for (int c = 0; c < kRegCh / 32; ++c) {
          uint32_t r[32];
          tmem_ld_32x32(taddr + c * 32, r);
          tmem_ld_wait();
          epilogue_store(p, r, dst + c * 32, c * 32, valid);
          if (valid) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float v = __uint_as_float(r[j]);
              S[(c * 32 + j) % kRegCh] += v;
              Q[(c * 32 + j) % kRegCh] = fmaf(v, v, Q[(c * 32 + j) % kRegCh]);
            }
          }
        }
      } else {
        epilogue_rows<BN>(p, taddr, dst, nt * BN, valid, lane, my_stats);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
    }
    if (BN == 64 && reg_stats) {
#pragma unroll
      for (int c = 0; c < kRegCh / 32; ++c) {
        float sv[32], qv[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          sv[j] = S[(c * 32 + j) % kRegCh];
          qv[j] = Q[(c * 32 + j) % kRegCh];
        }
        const float s1 = warp_column_sum(sv, lane), s2 = warp_column_sum(qv, lane);
        my_stats[c * 32 + lane] += s1;
        my_stats[p.cout_pad + c * 32 + lane] += s2;
      }
    }
  }

  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (p.stat_partials) {
    float* dstp = p.stat_partials + static_cast<long long>(blockIdx.x) * 2 * p.cout_pad;
    const int n2 = 2 * p.cout_pad;
    for (int i = threadIdx.x; i < n2; i += kFpropThreads)
      dstp[i] = (s_stats[i] + s_stats[n2 + i]) + (s_stats[2 * n2 + i] + s_stats[3 * n2 + i]);
  }
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Halo variant for the wide-and-shallow layers (cout_pad <= 128: the full- and half-resolution stages).  There the
// kernel above is bound by L2 -> shared-memory traffic, not by the tensor pipe: every tap re-fetches the (shifted)
// activation tile and every 128-pixel tile re-fetches the weights.  Here
//   * a CTA tile is 8 (w) x 32 (h) pixels = two 128-row accumulators; its activation patch (10 x 34 pixels x 64
//     channels, zero-filled by TMA outside the image) is fetched ONCE per 64-channel chunk;
//   * the nine taps are nine shifted views of that patch: the tcgen05 shared-memory descriptor starts at patch row
//     ((dr+1)*10 + ds+1) and strides 10 rows (1280 B) between 8-pixel groups -- the 128-byte swizzle is a function of
//     the absolute shared-memory address, so a start that is not 1024-byte aligned addresses the same data TMA wrote
//     (verified on B200: tools/exp/exp_desc.cu);
//   * each weight tile (tap, chunk) is fetched once per CTA tile and used by both accumulators.
// L2 -> smem bytes per 128x64x576 MMA block drop from 216 KB to 58 KB (BN = 64).
// ---------------------------------------------------------------------------------------------------------------
constexpr int kHaloW = 8, kHaloH = 32;
constexpr int kHaloThreads = 384;  // 4 control warps + 8 epilogue warps (4 per accumulator half)
constexpr int kPatchRows = (kHaloH + 2) * (kHaloW + 2);  // 340 pixel rows of 128 B
constexpr int kPatchBytes = kPatchRows * 128;             // 43520
constexpr int kPatchStride = 45056;                       // rounded up to 1 KB

template <int BN>
struct HaloCfg {
  static constexpr int kStagesA = 3;
  static constexpr int kStagesB = BN == 64 ? 8 : 5;
  static constexpr int kBBytes = BN * 128;
  static constexpr int kTmemCols = 4 * BN;  // 2 halves x double buffering
  static constexpr int kSmem = 1024 + kStagesA * kPatchStride + kStagesB * kBBytes + 4 * 2 * BN * 4 + 512;
};

template <int BN>
__global__ void __launch_bounds__(kHaloThreads, 1)
conv_fprop_halo_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                       const FpropParams p) {
  using Cfg = HaloCfg<BN>;
  constexpr int SA = Cfg::kStagesA, SB = Cfg::kStagesB;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = sA + SA * kPatchStride;
  float* s_stats = reinterpret_cast<float*>(sB + SB * Cfg::kBBytes);
  uint64_t* fullA = reinterpret_cast<uint64_t*>(s_stats + 4 * 2 * BN);
  uint64_t* emptyA = fullA + SA;
  uint64_t* fullB = emptyA + SA;
  uint64_t* emptyB = fullB + SB;
  uint64_t* tfull = emptyB + SB;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < SA; ++i) {
      mbar_init(&fullA[i], 1);
      mbar_init(&emptyA[i], 1);
    }
    for (int i = 0; i < SB; ++i) {
      mbar_init(&fullB[i], 1);
      mbar_init(&emptyB[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 8);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, Cfg::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // 384 threads x 168 registers at launch; the four control warps give registers back, the eight epilogue warps take
  // them (64 or 128 private statistic accumulators each)
  // (setmaxnreg sits inside the role branches: placed before them, ptxas applies the smaller bound to every role)
  if (warp < 4) {
  setmaxnreg_dec<40>();
  if (warp == 0) {
    if (lane == 0) {
      // ------------------------------- activation-patch producer -------------------------------
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        int mt = tile;
        const int w0 = (mt % p.tiles_w) * kHaloW;
        mt /= p.tiles_w;
        const int h0 = (mt % p.tiles_h) * kHaloH;
        const int n0 = mt / p.tiles_h;
        for (int chunk = 0; chunk < p.cin_chunks; ++chunk) {
          mbar_wait(&emptyA[stage], phase ^ 1);
          mbar_expect_tx(&fullA[stage], kPatchBytes);
          tma_load_4d(sA + stage * kPatchStride, &tmA, &fullA[stage], chunk * 64, w0 - 1, h0 - 1, n0);
          if (++stage == SA) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 3) {
    if (lane == 0) {
      // ------------------------------- weight-tile producer -------------------------------
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        for (int chunk = 0; chunk < p.cin_chunks; ++chunk) {
          for (int tap = 0; tap < 9; ++tap) {
            mbar_wait(&emptyB[stage], phase ^ 1);
            mbar_expect_tx(&fullB[stage], Cfg::kBBytes);
            tma_load_2d(sB + stage * Cfg::kBBytes, &tmB, &fullB[stage], tap * p.cin_pad + chunk * 64, 0);
            if (++stage == SB) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------- MMA issuer (whole warp, one elected lane issues) -------------------------------
    constexpr uint32_t idesc = idesc_bf16_f32(128, BN, false, false);
    constexpr uint32_t a_hi = desc_hi_sw128((kHaloW + 2) * 128), b_hi = desc_hi_sw128(1024);
    const uint32_t a_lo0 = desc_lo(smem_u32(sA), 16), b_lo0 = desc_lo(smem_u32(sB), 16);
    int sa = 0, sb = 0;
    uint32_t pa = 0, pb = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int h0 = ((tile / p.tiles_w) % p.tiles_h) * kHaloH;
      const bool two = h0 + 16 < p.H;  // lower half entirely below the image: skip its MMAs
      mbar_wait(&tempty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d0 = tmem_base + acc * 2 * BN;
      for (int chunk = 0; chunk < p.cin_chunks; ++chunk) {
        mbar_wait(&fullA[sa], pa);
        const uint32_t a_st = a_lo0 + sa * (kPatchStride >> 4);
        for (int tap = 0; tap < 9; ++tap) {
          mbar_wait(&fullB[sb], pb);
          tc_fence_after();
          const uint32_t b_lo = b_lo0 + sb * (Cfg::kBBytes >> 4);
          const int dr = tap / 3, ds = tap - dr * 3;  // already offset by +1 (patch origin is (h0-1, w0-1))
          const uint32_t a_lo = a_st + (dr * (kHaloW + 2) + ds) * 8;  // 128-byte rows in 16-byte units
          const uint32_t first = (chunk | tap) != 0 ? 1u : 0u;
          if (elect_one()) {
            umma_bf16_lohi(d0, a_lo, a_hi, b_lo, b_hi, idesc, first);
            umma_bf16_lohi(d0, a_lo + 2, a_hi, b_lo + 2, b_hi, idesc, 1u);
            umma_bf16_lohi(d0, a_lo + 4, a_hi, b_lo + 4, b_hi, idesc, 1u);
            umma_bf16_lohi(d0, a_lo + 6, a_hi, b_lo + 6, b_hi, idesc, 1u);
            if (two) {
              constexpr uint32_t half = 16 * (kHaloW + 2) * 8;
              umma_bf16_lohi(d0 + BN, a_lo + half, a_hi, b_lo, b_hi, idesc, first);
              umma_bf16_lohi(d0 + BN, a_lo + half + 2, a_hi, b_lo + 2, b_hi, idesc, 1u);
              umma_bf16_lohi(d0 + BN, a_lo + half + 4, a_hi, b_lo + 4, b_hi, idesc, 1u);
              umma_bf16_lohi(d0 + BN, a_lo + half + 6, a_hi, b_lo + 6, b_hi, idesc, 1u);
            }
            umma_commit(&emptyB[sb]);
          }
          __syncwarp();
          if (++sb == SB) {
            sb = 0;
            pb ^= 1;
          }
        }
        if (elect_one()) umma_commit(&emptyA[sa]);
        __syncwarp();
        if (++sa == SA) {
          sa = 0;
          pa ^= 1;
        }
      }
      if (elect_one()) umma_commit(&tfull[acc]);
      __syncwarp();
    }
  }
  } else {
    setmaxnreg_inc<232>();
    // --------------------------------- epilogue -----------------------------------
    // Warps 4-7 own output channels [0, BN/2), warps 8-11 [BN/2, BN), of BOTH accumulator halves; a warp reads the TMEM
    // lane quadrant warp % 4, i.e. tile rows ew*4 .. ew*4+3 of each half. BatchNorm statistics: every thread keeps a
    // private running sum / sum of squares per channel it sees (its pixel changes from tile to tile, its channels never
    // do), so the cross-lane reduction happens ONCE per kernel instead of once per tile -- per tile the statistics cost
    // one FADD + one FFMA per element (the shuffle-based per-tile reduction was ~10x that and issue-bound).
    constexpr int CH = BN / 2;  // channels per warp
    const int ew = warp & 3;
    const int cbase = ((warp - 4) >> 2) * CH;
    const int row = ew * 32 + lane;
    const int ty_ = row >> 3, tx_ = row & 7;
    float S[CH], Q[CH];
#pragma unroll
    for (int j = 0; j < CH; ++j) S[j] = Q[j] = 0.f;
    const bool want_stats = p.stat_partials != nullptr;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      int mt = tile;
      const int w = (mt % p.tiles_w) * kHaloW + tx_;
      mt /= p.tiles_w;
      const int h0 = (mt % p.tiles_h) * kHaloH;
      const int n = mt / p.tiles_h;
      const int halves = (h0 + 16 < p.H) ? 2 : 1;  // the MMA warp skips a lower half that lies below the image
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      for (int half = 0; half < halves; ++half) {
        const int h = h0 + half * 16 + ty_;
        const bool valid = h < p.H && w < p.W;
        __nv_bfloat16* dst = p.y + n * p.ysn + h * p.ysh + w * p.ysw + cbase;
        const uint32_t t_a = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + (acc * 2 + half) * BN + cbase;
#pragma unroll
        for (int c = 0; c < CH / 32; ++c) {
          uint32_t r[32];
          tmem_ld_32x32(t_a + c * 32, r);
          tmem_ld_wait();
          epilogue_store(p, r, dst + c * 32, cbase + c * 32, valid);
          if (want_stats && valid) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float v = __uint_as_float(r[j]);
              S[c * 32 + j] += v;
              Q[c * 32 + j] = fmaf(v, v, Q[c * 32 + j]);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
    }
    // one cross-lane reduction per kernel; s_stats[warp quadrant][sum | sumsq][BN]
#pragma unroll
    for (int c = 0; c < CH / 32; ++c) {
      float s[32], q[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        s[j] = S[c * 32 + j];
        q[j] = Q[c * 32 + j];
      }
      const float s1 = warp_column_sum(s, lane), s2 = warp_column_sum(q, lane);
      s_stats[ew * 2 * BN + cbase + c * 32 + lane] = s1;
      s_stats[ew * 2 * BN + BN + cbase + c * 32 + lane] = s2;
    }
  }

  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (p.stat_partials) {
    float* dstp = p.stat_partials + static_cast<long long>(blockIdx.x) * 2 * p.cout_pad;
    for (int i = threadIdx.x; i < 2 * BN; i += kHaloThreads)
      dstp[i] = (s_stats[i] + s_stats[2 * BN + i]) + (s_stats[4 * BN + i] + s_stats[6 * BN + i]);
  }
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// CTA-pair variant of the halo kernel (cta_group::2). In the single-CTA kernel an MMA of N = 64 reads 4 KB of A and
// 2 KB of B from shared memory (48-60 clk at 128 B/clk) for 32 clk of tensor work: the full-resolution 64-channel
// layers are shared-memory bound. Here two CTAs of one TPC each own an 8 x 32 pixel tile and its activation patch, each
// loads HALF of every weight tile, and the leader issues M = 256 MMAs that span both tiles: per SM an MMA now reads
// 4 KB + 1 KB. Everything else (patch reuse across the nine taps, per-thread BatchNorm statistics) is the kernel above.
//   barriers: full* live in the leader (two producer arrivals, transaction bytes of both CTAs' TMA loads); empty* and
//   tfull are signalled in both CTAs by multicast commits; tempty lives in the leader and collects the epilogue warps
//   of both CTAs.
// ---------------------------------------------------------------------------------------------------------------
template <int BN>
struct Halo2Cfg {
  static constexpr int kStagesA = 3;
  static constexpr int kStagesB = BN == 64 ? 20 : 10;
  static constexpr int kBBytes = (BN / 2) * 128;  // this CTA's half of a weight tile
  static constexpr int kTmemCols = 4 * BN;
  static constexpr int kSmem = 1024 + kStagesA * kPatchStride + kStagesB * kBBytes + 4 * 2 * BN * 4 + 512;
};

template <int BN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kHaloThreads, 1)
conv_fprop_halo2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                        const FpropParams p) {
  using Cfg = Halo2Cfg<BN>;
  constexpr int SA = Cfg::kStagesA, SB = Cfg::kStagesB;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = sA + SA * kPatchStride;
  float* s_stats = reinterpret_cast<float*>(sB + SB * Cfg::kBBytes);
  uint64_t* fullA = reinterpret_cast<uint64_t*>(s_stats + 4 * 2 * BN);
  uint64_t* emptyA = fullA + SA;
  uint64_t* fullB = emptyA + SA;
  uint64_t* emptyB = fullB + SB;
  uint64_t* tfull = emptyB + SB;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  const int pair_tiles = (p.total_tiles + 1) >> 1;  // tile pair t = tiles (2t, 2t+1); an odd last tile is duplicated

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < SA; ++i) {
      mbar_init(&fullA[i], 2);
      mbar_init(&emptyA[i], 1);
    }
    for (int i = 0; i < SB; ++i) {
      mbar_init(&fullB[i], 2);
      mbar_init(&emptyB[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 16);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc_pair(tmem_slot, Cfg::kTmemCols);
    tmem_relinquish_pair();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer's barriers and TMEM exist before anything is signalled across the pair
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (warp < 4) {
  setmaxnreg_dec<40>();
  if (warp == 0) {
    if (lane == 0) {
      // ------------------------------- activation-patch producer (own tile) -------------------------------
      int stage = 0;
      uint32_t phase = 0;
      for (int tp = pair; tp < pair_tiles; tp += n_pairs) {
        int mt = min(2 * tp + static_cast<int>(rank), p.total_tiles - 1);
        const int w0 = (mt % p.tiles_w) * kHaloW;
        mt /= p.tiles_w;
        const int h0 = (mt % p.tiles_h) * kHaloH;
        const int n0 = mt / p.tiles_h;
        for (int chunk = 0; chunk < p.cin_chunks; ++chunk) {
          mbar_wait(&emptyA[stage], phase ^ 1);
          if (rank == 0) mbar_expect_tx(&fullA[stage], 2 * kPatchBytes);
          tma_load_4d_pair(sA + stage * kPatchStride, &tmA, &fullA[stage], chunk * 64, w0 - 1, h0 - 1, n0);
          if (rank != 0) mbar_arrive_leader(&fullA[stage]);
          if (++stage == SA) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 3) {
    if (lane == 0) {
      // ------------------------------- weight-tile producer (own half of the cout rows) -------------------------------
      int stage = 0;
      uint32_t phase = 0;
      for (int tp = pair; tp < pair_tiles; tp += n_pairs) {
        for (int chunk = 0; chunk < p.cin_chunks; ++chunk) {
          for (int tap = 0; tap < 9; ++tap) {
            mbar_wait(&emptyB[stage], phase ^ 1);
            if (rank == 0) mbar_expect_tx(&fullB[stage], 2 * Cfg::kBBytes);
            tma_load_2d_pair(sB + stage * Cfg::kBBytes, &tmB, &fullB[stage], tap * p.cin_pad + chunk * 64,
                             static_cast<int>(rank) * (BN / 2));
            if (rank != 0) mbar_arrive_leader(&fullB[stage]);
            if (++stage == SB) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1 && rank == 0) {
    // ------------------------------- MMA issuer (leader CTA only) -------------------------------
    constexpr uint32_t idesc = idesc_bf16_f32(256, BN, false, false);
    constexpr uint32_t a_hi = desc_hi_sw128((kHaloW + 2) * 128), b_hi = desc_hi_sw128(1024);
    const uint32_t a_lo0 = desc_lo(smem_u32(sA), 16), b_lo0 = desc_lo(smem_u32(sB), 16);
    int sa = 0, sb = 0;
    uint32_t pa = 0, pb = 0;
    int it = 0;
    for (int tp = pair; tp < pair_tiles; tp += n_pairs, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int t0 = 2 * tp, t1 = min(2 * tp + 1, p.total_tiles - 1);
      const int h00 = ((t0 / p.tiles_w) % p.tiles_h) * kHaloH, h01 = ((t1 / p.tiles_w) % p.tiles_h) * kHaloH;
      const bool two = (h00 + 16 < p.H) || (h01 + 16 < p.H);  // lower halves of both tiles below the image: skip
      mbar_wait(&tempty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d0 = tmem_base + acc * 2 * BN;
      for (int chunk = 0; chunk < p.cin_chunks; ++chunk) {
        mbar_wait(&fullA[sa], pa);
        const uint32_t a_st = a_lo0 + sa * (kPatchStride >> 4);
        for (int tap = 0; tap < 9; ++tap) {
          mbar_wait(&fullB[sb], pb);
          tc_fence_after();
          const uint32_t b_lo = b_lo0 + sb * (Cfg::kBBytes >> 4);
          const int dr = tap / 3, ds = tap - dr * 3;
          const uint32_t a_lo = a_st + (dr * (kHaloW + 2) + ds) * 8;
          const uint32_t first = (chunk | tap) != 0 ? 1u : 0u;
          if (elect_one()) {
            umma_bf16_lohi_pair(d0, a_lo, a_hi, b_lo, b_hi, idesc, first);
            umma_bf16_lohi_pair(d0, a_lo + 2, a_hi, b_lo + 2, b_hi, idesc, 1u);
            umma_bf16_lohi_pair(d0, a_lo + 4, a_hi, b_lo + 4, b_hi, idesc, 1u);
            umma_bf16_lohi_pair(d0, a_lo + 6, a_hi, b_lo + 6, b_hi, idesc, 1u);
            if (two) {
              constexpr uint32_t half = 16 * (kHaloW + 2) * 8;
              umma_bf16_lohi_pair(d0 + BN, a_lo + half, a_hi, b_lo, b_hi, idesc, first);
              umma_bf16_lohi_pair(d0 + BN, a_lo + half + 2, a_hi, b_lo + 2, b_hi, idesc, 1u);
              umma_bf16_lohi_pair(d0 + BN, a_lo + half + 4, a_hi, b_lo + 4, b_hi, idesc, 1u);
              umma_bf16_lohi_pair(d0 + BN, a_lo + half + 6, a_hi, b_lo + 6, b_hi, idesc, 1u);
            }
            umma_commit_pair(&emptyB[sb]);
          }
          __syncwarp();
          if (++sb == SB) {
            sb = 0;
            pb ^= 1;
          }
        }
        if (elect_one()) umma_commit_pair(&emptyA[sa]);
        __syncwarp();
        if (++sa == SA) {
          sa = 0;
          pa ^= 1;
        }
      }
      if (elect_one()) umma_commit_pair(&tfull[acc]);
      __syncwarp();
    }
  }
  } else {
    setmaxnreg_inc<232>();
    // --------------------------------- epilogue (own tile, own TMEM) -----------------------------------
    constexpr int CH = BN / 2;
    const int ew = warp & 3;
    const int cbase = ((warp - 4) >> 2) * CH;
    const int row = ew * 32 + lane;
    const int ty_ = row >> 3, tx_ = row & 7;
    float S[CH], Q[CH];
#pragma unroll
    for (int j = 0; j < CH; ++j) S[j] = Q[j] = 0.f;
    const bool want_stats = p.stat_partials != nullptr;
    int it = 0;
    for (int tp = pair; tp < pair_tiles; tp += n_pairs, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      int mt = 2 * tp + static_cast<int>(rank);
      const bool real = mt < p.total_tiles;  // the duplicate of an odd last tile is computed but not stored
      mt = min(mt, p.total_tiles - 1);
      const int w = (mt % p.tiles_w) * kHaloW + tx_;
      mt /= p.tiles_w;
      const int h0 = (mt % p.tiles_h) * kHaloH;
      const int n = mt / p.tiles_h;
      const int halves = (h0 + 16 < p.H) ? 2 : 1;
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      for (int half = 0; half < halves; ++half) {
        const int h = h0 + half * 16 + ty_;
        const bool valid = real && h < p.H && w < p.W;
        __nv_bfloat16* dst = p.y + n * p.ysn + h * p.ysh + w * p.ysw + cbase;
        const uint32_t t_a = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + (acc * 2 + half) * BN + cbase;
#pragma unroll
        for (int c = 0; c < CH / 32; ++c) {
          uint32_t r[32];
          tmem_ld_32x32(t_a + c * 32, r);
          tmem_ld_wait();
          epilogue_store(p, r, dst + c * 32, cbase + c * 32, valid);
          if (want_stats && valid) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float v = __uint_as_float(r[j]);
              S[c * 32 + j] += v;
              Q[c * 32 + j] = fmaf(v, v, Q[c * 32 + j]);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(&tempty[acc]);
    }
#pragma unroll
    for (int c = 0; c < CH / 32; ++c) {
      float s[32], q[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        s[j] = S[c * 32 + j];
        q[j] = Q[c * 32 + j];
      }
      const float s1 = warp_column_sum(s, lane), s2 = warp_column_sum(q, lane);
      s_stats[ew * 2 * BN + cbase + c * 32 + lane] = s1;
      s_stats[ew * 2 * BN + BN + cbase + c * 32 + lane] = s2;
    }
  }

  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (p.stat_partials) {
    float* dstp = p.stat_partials + static_cast<long long>(blockIdx.x) * 2 * p.cout_pad;
    for (int i = threadIdx.x; i < 2 * BN; i += kHaloThreads)
      dstp[i] = (s_stats[i] + s_stats[2 * BN + i]) + (s_stats[4 * BN + i] + s_stats[6 * BN + i]);
  }
  cluster_sync_all();  // neither CTA leaves (or frees TMEM) while the other may still touch its barriers / TMEM
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, Cfg::kTmemCols);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Transposed variant for cout_pad == 64 (the full-resolution stages and every data gradient that lands on 64
// channels). With pixels on M, an MMA reads its 128 activation rows from shared memory in ~60 clk whatever N is
// (measured: 60 / 64 / 128 clk at N = 64 / 128 / 256), so N = 64 keeps the tensor pipe half idle. Here the roles are
// swapped: D^T[(shift, cout)][pixel] = A[(shift, cout)][cin] * B[pixel][cin]^T with N = 256 pixels per MMA.
//   * M = 128 = two 64-row blocks of weights. Block s holds the tap (dr - s, dc) of the activation view (dr, dc) the MMA
//     reads, so block s accumulates the OUTPUT ROW 2i + s from the pixel rows 2i + dr: one accumulator covers two
//     output rows per pixel-row pair, 12 views (dr = -1..2, dc = -1..1) carry the 18 (tap, shift) products (75 % of
//     the MMA rows do useful work; taps that do not exist for a block are zero rows, fetched as an out-of-bounds TMA box).
//   * B = a view of a ROW-PARITY patch: even image rows (dr = 0, 2) or odd ones (dr = -1, 1), 33 rows x 10 pixels, fetched
//     once per 64-channel chunk through a 5-D tensor map (rows split into (pair, parity)); the view of (dr, dc) starts
//     at patch row k0 (0 for dr = 0 / -1, 1 for dr = 2 / +1), patch column dc + 1, i.e. 128-byte row k0 * 10 + dc + 1, with
//     SBO = 10 pixels, like the taps of the halo kernel.
//   * tile = 8 (w) x 64 (h) output pixels = 256 accumulator columns, double-buffered in TMEM (512 columns).
//   * epilogue: a thread owns ONE (shift, channel) lane: BatchNorm statistics are two scalars per thread with no
//     cross-lane work at all; stores are 2-byte, 32 lanes = 64 contiguous bytes of one pixel.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kTrW = 8, kTrI = 32;
constexpr int kTrPitch = kTrW + 2;
constexpr int kTrPatchRows = (kTrI + 1) * kTrPitch;  // 330 pixel rows of 128 B
constexpr int kTrPatchBytes = kTrPatchRows * 128;    // 42240
constexpr int kTrPatchStride = 43008;                // rounded up to 1 KB
constexpr int kTrStagesP = 3, kTrStagesW = 5;
constexpr int kTrWBytes = 128 * 128;  // [2 shifts x 64 cout][64 cin]
constexpr int kTrThreads = 384;
constexpr int kTrSmem = 1024 + kTrStagesP * kTrPatchStride + kTrStagesW * kTrWBytes + 8 * 2048 + 8 * 64 * 4 + 256;

__device__ __forceinline__ int tr_cols(int H, int h0) {  // accumulator columns of a tile: 8 per row pair in the image
  int icnt = min(kTrI, (H - h0 + 1) >> 1);
  icnt = (icnt + 1) & ~1;  // N must be a multiple of 16
  return icnt * kTrW;
}

__global__ void __launch_bounds__(kTrThreads, 1)
conv_fprop_tr64_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
                       const FpropParams p) {
  constexpr int SP = kTrStagesP, SW = kTrStagesW;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sP = smem;
  uint8_t* sW = sP + SP * kTrPatchStride;
  uint8_t* sT = sW + SW * kTrWBytes;                                // [8 epilogue warps][32 pixels][32 channels] bf16
  float* s_stats = reinterpret_cast<float*>(sT + 8 * 2048);         // [8 epilogue warps][sum | sumsq][32 lanes]
  uint64_t* fullP = reinterpret_cast<uint64_t*>(s_stats + 8 * 64);
  uint64_t* emptyP = fullP + SP;
  uint64_t* fullW = emptyP + SP;
  uint64_t* emptyW = fullW + SW;
  uint64_t* tfull = emptyW + SW;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmW);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < SP; ++i) {
      mbar_init(&fullP[i], 1);
      mbar_init(&emptyP[i], 1);
    }
    for (int i = 0; i < SW; ++i) {
      mbar_init(&fullW[i], 1);
      mbar_init(&emptyW[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 8);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ------------------------------- row-parity patch producer -------------------------------
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        int mt = tile;
        const int w0 = (mt % p.tiles_w) * kTrW;
        mt /= p.tiles_w;
        const int h0 = (mt % p.tiles_h) * (2 * kTrI);
        const int n0 = mt / p.tiles_h;
        for (int chunk = 0; chunk < p.cin_chunks; ++chunk) {
          for (int par = 0; par < 2; ++par) {  // even rows h0 + 2k, then odd rows h0 - 1 + 2k  (k = 0..32)
            mbar_wait(&emptyP[stage], phase ^ 1);
            mbar_expect_tx(&fullP[stage], kTrPatchBytes);
            tma_load_5d(sP + stage * kTrPatchStride, &tmX, &fullP[stage], chunk * 64, w0 - 1, par, h0 / 2 - par, n0);
            if (++stage == SP) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 3) {
    if (lane == 0) {
      // ------------------------------- weight producer: [tap of shift 0 | tap of shift 1] per view -------------------------------
      int stage = 0;
      uint32_t phase = 0;
      const int oob = 9 * p.cin_pad;  // a box that starts past the last column is all zeros
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        for (int chunk = 0; chunk < p.cin_chunks; ++chunk) {
          for (int v = 0; v < 12; ++v) {
            const int par = v / 6, r = (v % 6) / 3, dc = v % 3;
            const int dr = par == 0 ? 2 * r : 2 * r - 1;             // view row offset: 0, 2 | -1, 1
            const int tap_a = dr <= 1 ? (dr + 1) * 3 + dc : -1;      // shift 0 uses tap (dr, dc)
            const int tap_b = dr >= 0 ? dr * 3 + dc : -1;            // shift 1 uses tap (dr - 1, dc)
            const int col_a = tap_a >= 0 ? tap_a * p.cin_pad + chunk * 64 : oob;
            const int col_b = tap_b >= 0 ? tap_b * p.cin_pad + chunk * 64 : oob;
            mbar_wait(&emptyW[stage], phase ^ 1);
            mbar_expect_tx(&fullW[stage], kTrWBytes);
            tma_load_2d(sW + stage * kTrWBytes, &tmW, &fullW[stage], col_a, 0);
            tma_load_2d(sW + stage * kTrWBytes + kTrWBytes / 2, &tmW, &fullW[stage], col_b, 0);
            if (++stage == SW) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------- MMA issuer -------------------------------
    constexpr uint32_t idesc0 = idesc_bf16_f32(128, 0, false, false);
    constexpr uint32_t w_hi = desc_hi_sw128(1024), x_hi = desc_hi_sw128(kTrPitch * 128);
    const uint32_t w_lo0 = desc_lo(smem_u32(sW), 16), x_lo0 = desc_lo(smem_u32(sP), 16);
    int sp = 0, sw = 0;
    uint32_t pp = 0, pw = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int h0 = ((tile / p.tiles_w) % p.tiles_h) * (2 * kTrI);
      const uint32_t idesc = idesc0 | (static_cast<uint32_t>(tr_cols(p.H, h0) >> 3) << 17);
      mbar_wait(&tempty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d = tmem_base + acc * 256;
      for (int chunk = 0; chunk < p.cin_chunks; ++chunk) {
        for (int par = 0; par < 2; ++par) {
          mbar_wait(&fullP[sp], pp);
          const uint32_t x_st = x_lo0 + sp * (kTrPatchStride >> 4);
          for (int v = 0; v < 6; ++v) {
            const int r = v / 3, dc = v - r * 3;
            mbar_wait(&fullW[sw], pw);
            tc_fence_after();
            const uint32_t w_lo = w_lo0 + sw * (kTrWBytes >> 4);
            const uint32_t x_lo = x_st + (r * kTrPitch + dc) * 8;  // view start: patch row r, column dc (16-byte units)
            const uint32_t first = (chunk | par | v) != 0 ? 1u : 0u;
            const int ks = chunk == p.cin_chunks - 1 ? p.ksteps_last : 4;
            if (elect_one()) {
              umma_bf16_lohi(d, w_lo, w_hi, x_lo, x_hi, idesc, first);
              if (ks > 1) umma_bf16_lohi(d, w_lo + 2, w_hi, x_lo + 2, x_hi, idesc, 1u);
              if (ks > 2) umma_bf16_lohi(d, w_lo + 4, w_hi, x_lo + 4, x_hi, idesc, 1u);
              if (ks > 3) umma_bf16_lohi(d, w_lo + 6, w_hi, x_lo + 6, x_hi, idesc, 1u);
              umma_commit(&emptyW[sw]);
            }
            __syncwarp();
            if (++sw == SW) {
              sw = 0;
              pw ^= 1;
            }
          }
          if (elect_one()) umma_commit(&emptyP[sp]);
          __syncwarp();
          if (++sp == SP) {
            sp = 0;
            pp ^= 1;
          }
        }
      }
      if (elect_one()) umma_commit(&tfull[acc]);
      __syncwarp();
    }
  } else if (warp >= 4) {
    // --------------------------------- epilogue -----------------------------------
    // TMEM lane = (shift, channel): warp quadrant wq = warp % 4 reads lanes 32 wq .. 32 wq + 31; warps 4-7 take accumulator
    // columns [0, 128), warps 8-11 [128, 256). Column q = pixel (row pair q / 8, column q % 8) of the tile. A block of 32
    // channels x 32 pixels is turned around through 2 KB of shared memory (2-byte stores, one 64-byte row per pixel) so
    // that global memory sees 16-byte stores; a warp is a serial instruction stream, and at 2-byte global stores with
    // per-element addressing the epilogue (~20 instructions per element) ran 2x longer than the tile's MMAs (ncu).
    const int wq = warp & 3, chalf = (warp - 4) >> 2;
    const int shift = wq >> 1, co = (wq & 1) * 32 + lane;
    const float sc = p.scale ? __ldg(p.scale + co) : 1.f, sh = p.scale ? __ldg(p.shift + co) : 0.f;
    const bool want_stats = p.stat_partials != nullptr;
    const bool affine = p.scale != nullptr, relu = p.relu != 0;
    uint8_t* tbuf = sT + (warp - 4) * 2048;
    const uint32_t tb_st = smem_u32(tbuf) + lane * 2;                       // + pixel * 64
    const uint32_t tb_ld = smem_u32(tbuf) + (lane >> 2) * 64 + (lane & 3) * 16;  // + 8-pixel group * 512
    const int jj = lane >> 2, part = lane & 3;  // this lane's pixel column / 8-channel group in the 16-byte phase
    float S = 0.f, Q = 0.f;
    // BatchNorm+ReLU backward statistics of the consuming block, gathered in the 16-byte phase (a thread's 8 channels
    // never change): g = da * [y * scale + shift > 0], sums of g and g * y
    const bool bwd = p.bpartials != nullptr;
    float bsc[8], bsh[8], BS[8], BQ[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      bsc[j] = bwd ? __ldg(p.bscale + (wq & 1) * 32 + part * 8 + j) : 0.f;
      bsh[j] = bwd ? __ldg(p.bshift + (wq & 1) * 32 + part * 8 + j) : 0.f;
      BS[j] = BQ[j] = 0.f;
    }
    int it = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      int mt = tile;
      const int w0 = (mt % p.tiles_w) * kTrW;
      mt /= p.tiles_w;
      const int h0 = (mt % p.tiles_h) * (2 * kTrI);
      const int n = mt / p.tiles_h;
      const int ncols = tr_cols(p.H, h0);
      const int wvalid = p.W - w0;                       // columns j < wvalid are inside the image
      const int ivalid = (p.H - h0 - shift + 1) >> 1;    // row pairs i < ivalid have row 2i + shift inside the image
      __nv_bfloat16* gbase = p.y + n * p.ysn + (h0 + shift) * p.ysh + (w0 + jj) * p.ysw + (wq & 1) * 32 + part * 8;
      const __nv_bfloat16* ybase = p.by + n * p.bysn + (h0 + shift) * p.bysh + (w0 + jj) * p.bysw + (wq & 1) * 32 + part * 8;
      // y of the consuming block for the backward statistics: its addresses do not depend on the accumulator, so block
      // b + 1 is fetched while block b is processed (block 0 before the wait for the MMAs)
      uint4 ynext[4];
      auto fetch_y = [&](int blk, uint4 (&dst)[4]) {
        const int ib = (chalf * 128 + blk * 32) >> 3;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          dst[k] = (bwd && blk < 4 && jj < wvalid && ib + k < ivalid && chalf * 128 + blk * 32 < ncols)
                       ? ldg16(ybase + 2 * (ib + k) * p.bysh)
                       : make_uint4(0u, 0u, 0u, 0u);
      };
      fetch_y(0, ynext);
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
#pragma unroll 1
      for (int b = 0; b < 4; ++b) {
        const int c0 = chalf * 128 + b * 32;
        if (c0 >= ncols) break;
        const int i0 = c0 >> 3;
        uint4 ycur[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) ycur[k] = ynext[k];
        fetch_y(b + 1, ynext);
        uint32_t r[32];
        tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(wq * 32) << 16) + acc * 256 + c0, r);
        tmem_ld_wait();
        if (want_stats) {
          if (wvalid >= kTrW && i0 + 4 <= ivalid) {  // block entirely inside the image
#pragma unroll
            for (int t = 0; t < 32; ++t) {
              const float v = __uint_as_float(r[t]);
              S += v;
              Q = fmaf(v, v, Q);
            }
          } else {
#pragma unroll
            for (int t = 0; t < 32; ++t) {
              const float v = (i0 + (t >> 3) < ivalid && (t & 7) < wvalid) ? __uint_as_float(r[t]) : 0.f;
              S += v;
              Q = fmaf(v, v, Q);
            }
          }
        }
#pragma unroll
        for (int t = 0; t < 32; ++t) {
          float v = __uint_as_float(r[t]);
          if (affine) {
            v = fmaf(v, sc, sh);
            if (relu) v = fmaxf(v, 0.f);
          }
          const __nv_bfloat16 hv = __float2bfloat16_rn(v);
          asm volatile("st.shared.b16 [%0], %1;" ::"r"(tb_st + t * 64), "h"(*reinterpret_cast<const uint16_t*>(&hv)) : "memory");
        }
        __syncwarp();
        if (!(p.debug & 1) && jj < wvalid && (wq & 1) * 32 + part * 8 < p.cout_store) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            uint4 o;
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(o.x), "=r"(o.y), "=r"(o.z), "=r"(o.w) : "r"(tb_ld + k * 512) : "memory");
            if (i0 + k < ivalid) {
              *reinterpret_cast<uint4*>(gbase + 2 * (i0 + k) * p.ysh) = o;
              if (bwd) {
                float g[8], yv[8];
                unpack8(o, g);
                unpack8(ycur[k], yv);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  const float gj = fmaf(yv[j], bsc[j], bsh[j]) > 0.f ? g[j] : 0.f;
                  BS[j] += gj;
                  BQ[j] = fmaf(gj, yv[j], BQ[j]);
                }
              }
            }
          }
        }
        __syncwarp();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
    }
    if (!bwd) {
      s_stats[(warp - 4) * 64 + lane] = S;
      s_stats[(warp - 4) * 64 + 32 + lane] = Q;
    } else {
      // lanes with the same 8-channel group differ in their pixel column (lane >> 2): fold them, lanes 0-3 write
#pragma unroll
      for (int j = 0; j < 8; ++j) {
#pragma unroll
        for (int m = 4; m < 32; m <<= 1) {
          BS[j] += __shfl_xor_sync(0xffffffffu, BS[j], m);
          BQ[j] += __shfl_xor_sync(0xffffffffu, BQ[j], m);
        }
      }
      if (lane < 4) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          s_stats[(warp - 4) * 64 + lane * 8 + j] = BS[j];
          s_stats[(warp - 4) * 64 + 32 + lane * 8 + j] = BQ[j];
        }
      }
    }
  }

  __syncwarp();
  tc_fence_before();
  __syncthreads();
  float* partials_out = p.stat_partials ? p.stat_partials : p.bpartials;
  if (partials_out && threadIdx.x < 128) {
    // channel c is held by the warps with (wq & 1) == c / 32: both shifts, both column halves
    const int which = threadIdx.x >> 6, c = threadIdx.x & 63, half = c >> 5, l = c & 31;
    float a = 0.f;
#pragma unroll
    for (int ch = 0; ch < 2; ++ch)
#pragma unroll
      for (int sft = 0; sft < 2; ++sft) a += s_stats[(ch * 4 + sft * 2 + half) * 64 + which * 32 + l];
    partials_out[static_cast<long long>(blockIdx.x) * 2 * p.cout_pad + which * p.cout_pad + c] = a;
  }
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Transposed variant for cout_pad == 128 (the half-resolution stages and the data gradients that land on 128 channels).
// Same idea as the cout = 64 kernel without the row-shift trick: the 128 output channels ARE the M dimension,
// D^T[cout][pixel] = W_tap[cout][cin] * X_tap[pixel][cin]^T, nine taps = nine shifted views of one 10 x 34 pixel patch
// (the halo kernel's patch), N = 256 pixels (8 w x 32 h) per MMA. Per unit of work an MMA now reads 128 + 256 operand
// rows instead of 2 x (128 + 128): 1.3x fewer (the halo kernel ran at the operand-row bound, ~113 clk per N = 128 MMA).
// Epilogue as in the cout = 64 kernel: a thread owns one channel lane (statistics = two scalars), blocks of 32 channels x
// 32 pixels are turned through shared memory into 16-byte stores.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kT8StagesP = 2, kT8StagesW = 7;
constexpr int kT8WBytes = 128 * 128;  // [128 cout][64 cin]
constexpr int kT8Smem = 1024 + kT8StagesP * kPatchStride + kT8StagesW * kT8WBytes + 8 * 2048 + 8 * 64 * 4 + 256;

__device__ __forceinline__ int t8_cols(int H, int h0) {
  int rows = min(kHaloH, H - h0);
  rows = (rows + 1) & ~1;  // N must be a multiple of 16
  return rows * kHaloW;
}

__global__ void __launch_bounds__(kTrThreads, 1)
conv_fprop_tr128_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
                        const FpropParams p) {
  constexpr int SP = kT8StagesP, SW = kT8StagesW;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sP = smem;
  uint8_t* sW = sP + SP * kPatchStride;
  uint8_t* sT = sW + SW * kT8WBytes;                                // [8 epilogue warps][32 pixels][32 channels] bf16
  float* s_stats = reinterpret_cast<float*>(sT + 8 * 2048);         // [8 epilogue warps][sum | sumsq][32 lanes]
  uint64_t* fullP = reinterpret_cast<uint64_t*>(s_stats + 8 * 64);
  uint64_t* emptyP = fullP + SP;
  uint64_t* fullW = emptyP + SP;
  uint64_t* emptyW = fullW + SW;
  uint64_t* tfull = emptyW + SW;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmW);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < SP; ++i) {
      mbar_init(&fullP[i], 1);
      mbar_init(&emptyP[i], 1);
    }
    for (int i = 0; i < SW; ++i) {
      mbar_init(&fullW[i], 1);
      mbar_init(&emptyW[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 8);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ------------------------------- activation-patch producer -------------------------------
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        int mt = tile;
        const int w0 = (mt % p.tiles_w) * kHaloW;
        mt /= p.tiles_w;
        const int h0 = (mt % p.tiles_h) * kHaloH;
        const int n0 = mt / p.tiles_h;
        for (int chunk = 0; chunk < p.cin_chunks; ++chunk) {
          mbar_wait(&emptyP[stage], phase ^ 1);
          mbar_expect_tx(&fullP[stage], kPatchBytes);
          tma_load_4d(sP + stage * kPatchStride, &tmX, &fullP[stage], chunk * 64, w0 - 1, h0 - 1, n0);
          if (++stage == SP) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 3) {
    if (lane == 0) {
      // ------------------------------- weight producer: one [128 cout][64 cin] tile per (chunk, tap) -------------------------------
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        for (int chunk = 0; chunk < p.cin_chunks; ++chunk) {
          for (int tap = 0; tap < 9; ++tap) {
            mbar_wait(&emptyW[stage], phase ^ 1);
            mbar_expect_tx(&fullW[stage], kT8WBytes);
            tma_load_2d(sW + stage * kT8WBytes, &tmW, &fullW[stage], tap * p.cin_pad + chunk * 64, 0);
            if (++stage == SW) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------- MMA issuer -------------------------------
    constexpr uint32_t idesc0 = idesc_bf16_f32(128, 0, false, false);
    constexpr uint32_t w_hi = desc_hi_sw128(1024), x_hi = desc_hi_sw128((kHaloW + 2) * 128);
    const uint32_t w_lo0 = desc_lo(smem_u32(sW), 16), x_lo0 = desc_lo(smem_u32(sP), 16);
    int sp = 0, sw = 0;
    uint32_t pp = 0, pw = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int h0 = ((tile / p.tiles_w) % p.tiles_h) * kHaloH;
      const uint32_t idesc = idesc0 | (static_cast<uint32_t>(t8_cols(p.H, h0) >> 3) << 17);
      mbar_wait(&tempty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d = tmem_base + acc * 256;
      for (int chunk = 0; chunk < p.cin_chunks; ++chunk) {
        mbar_wait(&fullP[sp], pp);
        const uint32_t x_st = x_lo0 + sp * (kPatchStride >> 4);
        for (int tap = 0; tap < 9; ++tap) {
          mbar_wait(&fullW[sw], pw);
          tc_fence_after();
          const uint32_t w_lo = w_lo0 + sw * (kT8WBytes >> 4);
          const int dr = tap / 3, ds = tap - dr * 3;
          const uint32_t x_lo = x_st + (dr * (kHaloW + 2) + ds) * 8;
          const uint32_t first = (chunk | tap) != 0 ? 1u : 0u;
          if (elect_one()) {
            umma_bf16_lohi(d, w_lo, w_hi, x_lo, x_hi, idesc, first);
            umma_bf16_lohi(d, w_lo + 2, w_hi, x_lo + 2, x_hi, idesc, 1u);
            umma_bf16_lohi(d, w_lo + 4, w_hi, x_lo + 4, x_hi, idesc, 1u);
            umma_bf16_lohi(d, w_lo + 6, w_hi, x_lo + 6, x_hi, idesc, 1u);
            umma_commit(&emptyW[sw]);
          }
          __syncwarp();
          if (++sw == SW) {
            sw = 0;
            pw ^= 1;
          }
        }
        if (elect_one()) umma_commit(&emptyP[sp]);
        __syncwarp();
        if (++sp == SP) {
          sp = 0;
          pp ^= 1;
        }
      }
      if (elect_one()) umma_commit(&tfull[acc]);
      __syncwarp();
    }
  } else if (warp >= 4) {
    // --------------------------------- epilogue -----------------------------------
    // TMEM lane = output channel: warp quadrant wq = warp % 4 reads lanes 32 wq .. 32 wq + 31; warps 4-7 take accumulator
    // columns [0, 128), warps 8-11 [128, 256). Column q = pixel (row q / 8, column q % 8) of the tile.
    const int wq = warp & 3, chalf = (warp - 4) >> 2;
    const int co = wq * 32 + lane;
    const float sc = p.scale ? __ldg(p.scale + co) : 1.f, sh = p.scale ? __ldg(p.shift + co) : 0.f;
    const bool want_stats = p.stat_partials != nullptr;
    const bool affine = p.scale != nullptr, relu = p.relu != 0;
    uint8_t* tbuf = sT + (warp - 4) * 2048;
    const uint32_t tb_st = smem_u32(tbuf) + lane * 2;
    const uint32_t tb_ld = smem_u32(tbuf) + (lane >> 2) * 64 + (lane & 3) * 16;
    const int jj = lane >> 2, part = lane & 3;
    float S = 0.f, Q = 0.f;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      int mt = tile;
      const int w0 = (mt % p.tiles_w) * kHaloW;
      mt /= p.tiles_w;
      const int h0 = (mt % p.tiles_h) * kHaloH;
      const int n = mt / p.tiles_h;
      const int ncols = t8_cols(p.H, h0);
      const int wvalid = p.W - w0;   // columns j < wvalid are inside the image
      const int rvalid = p.H - h0;   // tile rows r < rvalid are inside the image
      __nv_bfloat16* gbase = p.y + n * p.ysn + h0 * p.ysh + (w0 + jj) * p.ysw + wq * 32 + part * 8;
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
#pragma unroll 1
      for (int b = 0; b < 4; ++b) {
        const int c0 = chalf * 128 + b * 32;
        if (c0 >= ncols) break;
        const int r0 = c0 >> 3;
        uint32_t r[32];
        tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(wq * 32) << 16) + acc * 256 + c0, r);
        tmem_ld_wait();
        if (want_stats) {
          if (wvalid >= kHaloW && r0 + 4 <= rvalid) {
#pragma unroll
            for (int t = 0; t < 32; ++t) {
              const float v = __uint_as_float(r[t]);
              S += v;
              Q = fmaf(v, v, Q);
            }
          } else {
#pragma unroll
            for (int t = 0; t < 32; ++t) {
              const float v = (r0 + (t >> 3) < rvalid && (t & 7) < wvalid) ? __uint_as_float(r[t]) : 0.f;
              S += v;
              Q = fmaf(v, v, Q);
            }
          }
        }
#pragma unroll
        for (int t = 0; t < 32; ++t) {
          float v = __uint_as_float(r[t]);
          if (affine) {
            v = fmaf(v, sc, sh);
            if (relu) v = fmaxf(v, 0.f);
          }
          const __nv_bfloat16 hv = __float2bfloat16_rn(v);
          asm volatile("st.shared.b16 [%0], %1;" ::"r"(tb_st + t * 64), "h"(*reinterpret_cast<const uint16_t*>(&hv)) : "memory");
        }
        __syncwarp();
        if (!(p.debug & 1) && jj < wvalid) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            uint4 o;
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(o.x), "=r"(o.y), "=r"(o.z), "=r"(o.w) : "r"(tb_ld + k * 512) : "memory");
            if (r0 + k < rvalid) *reinterpret_cast<uint4*>(gbase + (r0 + k) * p.ysh) = o;
          }
        }
        __syncwarp();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
    }
    s_stats[(warp - 4) * 64 + lane] = S;
    s_stats[(warp - 4) * 64 + 32 + lane] = Q;
  }

  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (p.stat_partials && threadIdx.x < 256) {
    // channel c is held by quadrant c / 32 of both column halves
    const int which = threadIdx.x >> 7, c = threadIdx.x & 127, q = c >> 5, l = c & 31;
    const float a = s_stats[q * 64 + which * 32 + l] + s_stats[(4 + q) * 64 + which * 32 + l];
    p.stat_partials[static_cast<long long>(blockIdx.x) * 2 * p.cout_pad + which * p.cout_pad + c] = a;
  }
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// Pixel tile (TW x TH x TN <= 128) that wastes the fewest MMA rows for this image size.
static void pick_tile(int N, int H, int W, int* TW, int* TH, int* TN) {
  double best = -1.0;
  int bw = 1, bh = 1, bn = 1;
  for (int tn = 1; tn <= 128 && tn <= N; tn *= 2) {
    for (int th = 1; th * tn <= 128 && th <= H; ++th) {
      int tw = 128 / (tn * th);
      if (tw > W) tw = W;
      if (tw > 256) tw = 256;
      if (tw < 1) continue;
      // also try the widest tw that divides W exactly
      int cands[2] = {tw, tw};
      for (int d = tw; d >= 1; --d)
        if (W % d == 0) {
          cands[1] = d;
          break;
        }
      for (int k = 0; k < 2; ++k) {
        int w_ = cands[k];
        long long tiles = 1LL * ((N + tn - 1) / tn) * ((H + th - 1) / th) * ((W + w_ - 1) / w_);
        double eff = static_cast<double>(1LL * N * H * W) / static_cast<double>(tiles * 128);
        // prefer efficiency, then no batch folding, then wider rows (longer contiguous global runs)
        double score = eff - 1e-4 * (tn > 1) + 1e-6 * w_;
        if (score > best) {
          best = score;
          bw = w_;
          bh = th;
          bn = tn;
        }
      }
    }
  }
  *TW = bw;
  *TH = bh;
  *TN = bn;
}

template <int BN>
static int launch_fprop(const CUtensorMap& tmA, const CUtensorMap& tmB, const FpropParams& p, cudaStream_t st) {
  using Cfg = FpropCfg<BN>;
  static bool configured = false;  // benign race: the attribute call is idempotent
  if (!configured) {
    CVB_CUDA(cudaFuncSetAttribute(conv_fprop_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmem));
    configured = true;
  }
  conv_fprop_kernel<BN><<<sm_count(), kFpropThreads, Cfg::kSmem, st>>>(tmA, tmB, p);
  CVB_LAUNCH_CHECK();
  return CVB_OK;
}

template <int BN>
static int launch_fprop_halo(const CUtensorMap& tmA, const CUtensorMap& tmB, const FpropParams& p, cudaStream_t st) {
  using Cfg = HaloCfg<BN>;
  static bool configured = false;
  if (!configured) {
    CVB_CUDA(cudaFuncSetAttribute(conv_fprop_halo_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmem));
    configured = true;
  }
  conv_fprop_halo_kernel<BN><<<sm_count(), kHaloThreads, Cfg::kSmem, st>>>(tmA, tmB, p);
  CVB_LAUNCH_CHECK();
  return CVB_OK;
}

template <int BN>
static int launch_fprop_halo2(const CUtensorMap& tmA, const CUtensorMap& tmB, const FpropParams& p, cudaStream_t st) {
  using Cfg = Halo2Cfg<BN>;
  static bool configured = false;
  if (!configured) {
    CVB_CUDA(cudaFuncSetAttribute(conv_fprop_halo2_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmem));
    configured = true;
  }
  conv_fprop_halo2_kernel<BN><<<sm_count() & ~1, kHaloThreads, Cfg::kSmem, st>>>(tmA, tmB, p);
  CVB_LAUNCH_CHECK();
  return CVB_OK;
}

static int launch_fprop_tr64(const CUtensorMap& tmX, const CUtensorMap& tmW, const FpropParams& p, cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    CVB_CUDA(cudaFuncSetAttribute(conv_fprop_tr64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kTrSmem));
    configured = true;
  }
  const int grid = sm_count();  // every CTA writes its row of the BatchNorm partials (zeros when it has no tile)
  conv_fprop_tr64_kernel<<<grid, kTrThreads, kTrSmem, st>>>(tmX, tmW, p);
  CVB_LAUNCH_CHECK();
  return CVB_OK;
}

static int launch_fprop_tr128(const CUtensorMap& tmX, const CUtensorMap& tmW, const FpropParams& p, cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    CVB_CUDA(cudaFuncSetAttribute(conv_fprop_tr128_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kT8Smem));
    configured = true;
  }
  conv_fprop_tr128_kernel<<<sm_count(), kTrThreads, kT8Smem, st>>>(tmX, tmW, p);
  CVB_LAUNCH_CHECK();
  return CVB_OK;
}

// Kernel selection shared by the entry point and the cvb_conv3x3_fprop_fuses_bwd_stats query.
static bool uses_tr64(const cvb_view& x, const cvb_view& y, int taps) {
  static int tr_mode = -1;  // CVB_TR64=0 keeps the pixel-major halo kernel for cout = 64 (A/B measurements)
  if (tr_mode < 0) {
    const char* e = getenv("CVB_TR64");
    tr_mode = e ? atoi(e) : 1;
  }
  return taps == 9 && y.c <= 64 && (x.h % 2) == 0 && x.w >= kTrW && tr_mode != 0;
}

}  // namespace cvb

using namespace cvb;

extern "C" int cvb_conv3x3_fprop(cvb_view x, const void* wpack, int taps, cvb_view y, const cvb_conv_epilogue* ep,
                                 void* stream) {
  int rc = check_view(x, "conv_fprop.x");
  if (rc) return rc;
  rc = check_view(y, "conv_fprop.y");
  if (rc) return rc;
  CVB_REQUIRE(wpack != nullptr, CVB_ERR_INVALID_ARG, "conv_fprop: null weights");
  CVB_REQUIRE((reinterpret_cast<uintptr_t>(wpack) & 15) == 0, CVB_ERR_INVALID_ARG, "conv_fprop: weights not 16-byte aligned");
  CVB_REQUIRE(taps == 9 || taps == 1, CVB_ERR_INVALID_ARG, "conv_fprop: taps must be 9 or 1 (got %d)", taps);
  CVB_REQUIRE(x.n == y.n && x.h == y.h && x.w == y.w, CVB_ERR_INVALID_ARG,
              "conv_fprop: x %dx%dx%d and y %dx%dx%d spatial shapes differ", x.n, x.h, x.w, y.n, y.h, y.w);
  CVB_REQUIRE((x.c % 16) == 0 && (y.c % 8) == 0, CVB_ERR_UNSUPPORTED,
              "conv_fprop: cin must be a multiple of 16 and cout a multiple of 8 (cin %d, cout %d)", x.c, y.c);
  // y may hold fewer channels than the 64-padded GEMM computes (12 classes -> 16 channels in memory): only the
  // transposed cout = 64 kernel stores a channel subset
  const int cout_pad = (y.c + 63) / 64 * 64;
  const bool narrow_out = y.c != cout_pad;
  CVB_REQUIRE(!narrow_out || uses_tr64(x, y, taps), CVB_ERR_UNSUPPORTED,
              "conv_fprop: cout %d is not a multiple of 64 and the cout = 64 kernel does not apply to these views", y.c);
  CVB_REQUIRE(y.c <= 1024, CVB_ERR_UNSUPPORTED, "conv_fprop: cout %d > 1024", y.c);

  FpropParams p;
  memset(&p, 0, sizeof(p));
  p.N = x.n; p.H = x.h; p.W = x.w;
  // The weights are packed for cin padded to 64; x itself may stop short of the last chunk (TMA zero-fills the rest of
  // the box, the MMA issuer skips the all-zero K steps). Only the generic and the cout = 64 kernels implement that.
  const int cin_pad = (x.c + 63) / 64 * 64;
  const bool narrow = x.c != cin_pad;
  p.cin_pad = cin_pad; p.cout_pad = cout_pad; p.cout_store = y.c;
  p.cin_chunks = cin_pad / 64;
  p.ksteps_last = narrow ? (x.c % 64 + 15) / 16 : 4;
  p.taps = taps;
  pick_tile(x.n, x.h, x.w, &p.TW, &p.TH, &p.TN);
  p.tiles_w = (x.w + p.TW - 1) / p.TW;
  p.tiles_h = (x.h + p.TH - 1) / p.TH;
  p.tiles_n = (x.n + p.TN - 1) / p.TN;
  const int BN = (cout_pad % 256 == 0) ? 256 : ((cout_pad % 128 == 0) ? 128 : 64);
  p.n_tiles = cout_pad / BN;
  long long total = 1LL * p.tiles_w * p.tiles_h * p.tiles_n * p.n_tiles;
  CVB_REQUIRE(total < (1LL << 31), CVB_ERR_UNSUPPORTED, "conv_fprop: too many tiles");
  p.total_tiles = static_cast<int>(total);
  p.a_bytes = static_cast<uint32_t>(p.TW * p.TH * p.TN) * 128u;
  {
    static int dbg = -1;
    if (dbg < 0) {
      const char* e = getenv("CVB_DEBUG");
      dbg = e ? atoi(e) : 0;
    }
    p.debug = dbg;
    if ((dbg & 2) && ep && ep->stat_partials) p.trace = reinterpret_cast<long long*>(ep->stat_partials);
  }
  p.y = static_cast<__nv_bfloat16*>(y.ptr);
  p.ysn = y.sn; p.ysh = y.sh; p.ysw = y.sw;
  if (ep) {
    p.stat_partials = (p.debug & 2) ? nullptr : ep->stat_partials;
    p.scale = ep->scale;
    p.shift = ep->shift;
    p.relu = ep->relu;
    CVB_REQUIRE((ep->scale == nullptr) == (ep->shift == nullptr), CVB_ERR_INVALID_ARG,
                "conv_fprop: scale and shift must both be given or both be NULL");
  }
  CUtensorMap tmA, tmB;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool want_bwd = ep && ep->bwd_partials != nullptr;
  if (want_bwd) {
    CVB_REQUIRE(ep->bwd_scale && ep->bwd_shift && ep->bwd_y.ptr, CVB_ERR_INVALID_ARG, "conv_fprop: incomplete bwd_* epilogue");
    CVB_REQUIRE(!ep->stat_partials && !ep->scale, CVB_ERR_INVALID_ARG,
                "conv_fprop: the bwd_* epilogue excludes stat_partials and scale/shift");
    rc = check_view(ep->bwd_y, "conv_fprop.bwd_y");
    if (rc) return rc;
    CVB_REQUIRE(same_shape(ep->bwd_y, y), CVB_ERR_INVALID_ARG, "conv_fprop: bwd_y and y shapes differ");
    CVB_REQUIRE(uses_tr64(x, y, taps) && !narrow_out, CVB_ERR_UNSUPPORTED,
                "conv_fprop: no kernel with the bwd_* epilogue for these views (ask cvb_conv3x3_fprop_fuses_bwd_stats)");
    p.by = static_cast<const __nv_bfloat16*>(ep->bwd_y.ptr);
    p.bysn = ep->bwd_y.sn; p.bysh = ep->bwd_y.sh; p.bysw = ep->bwd_y.sw;
    p.bscale = ep->bwd_scale; p.bshift = ep->bwd_shift; p.bpartials = ep->bwd_partials;
  }
  if (uses_tr64(x, y, taps)) {
    // cout = 64: transposed kernel (weights on M, 256 pixels on N)
    p.TW = kTrW; p.TH = 2 * kTrI; p.TN = 1;
    p.tiles_w = (x.w + kTrW - 1) / kTrW;
    p.tiles_h = (x.h + 2 * kTrI - 1) / (2 * kTrI);
    p.tiles_n = x.n;
    p.n_tiles = 1;
    long long tr_total = 1LL * p.tiles_w * p.tiles_h * p.tiles_n;
    CVB_REQUIRE(tr_total < (1LL << 31), CVB_ERR_UNSUPPORTED, "conv_fprop: too many tiles");
    p.total_tiles = static_cast<int>(tr_total);
    rc = make_act_tmap_rowpairs(&tmA, x, kTrW + 2, kTrI + 1);
    if (rc) return rc;
    rc = make_mat_tmap(&tmB, wpack, cout_pad, 9LL * cin_pad, 64);
    if (rc) return rc;
    return launch_fprop_tr64(tmA, tmB, p, st);
  }
  const char* t8_env = getenv("CVB_TR128");  // CVB_TR128=0 keeps the pixel-major halo kernel for cout = 128 (A/B)
  if (taps == 9 && y.c == 128 && !narrow && x.w >= kHaloW && x.h >= 2 && !(t8_env && atoi(t8_env) == 0)) {
    // cout = 128: transposed kernel (channels on M, 256 pixels on N)
    p.TW = kHaloW; p.TH = kHaloH; p.TN = 1;
    p.tiles_w = (x.w + kHaloW - 1) / kHaloW;
    p.tiles_h = (x.h + kHaloH - 1) / kHaloH;
    p.tiles_n = x.n;
    p.n_tiles = 1;
    long long t8_total = 1LL * p.tiles_w * p.tiles_h * p.tiles_n;
    CVB_REQUIRE(t8_total < (1LL << 31), CVB_ERR_UNSUPPORTED, "conv_fprop: too many tiles");
    p.total_tiles = static_cast<int>(t8_total);
    rc = make_act_tmap(&tmA, x, kHaloW + 2, kHaloH + 2, 1);
    if (rc) return rc;
    rc = make_mat_tmap(&tmB, wpack, y.c, 9LL * cin_pad, 128);
    if (rc) return rc;
    return launch_fprop_tr128(tmA, tmB, p, st);
  }
  if (taps == 9 && y.c <= 128 && !narrow && x.w >= kHaloW && x.h >= 16) {
    // wide-and-shallow layer: halo kernel (one patch fetch per chunk, nine shifted descriptor views)
    p.TW = kHaloW; p.TH = kHaloH; p.TN = 1;
    p.tiles_w = (x.w + kHaloW - 1) / kHaloW;
    p.tiles_h = (x.h + kHaloH - 1) / kHaloH;
    p.tiles_n = x.n;
    p.n_tiles = 1;
    p.total_tiles = p.tiles_w * p.tiles_h * p.tiles_n;
    rc = make_act_tmap(&tmA, x, kHaloW + 2, kHaloH + 2, 1);
    if (rc) return rc;
    // CVB_HALO_PAIR=1 selects the CTA-pair kernel: correct, but measured ~2x SLOWER than the single-CTA kernel on
    // B200, so it is off by default; read per call so that a test can exercise it.
    const char* pe = getenv("CVB_HALO_PAIR");
    const int pair_mode = pe ? atoi(pe) : 0;
    const bool use_pair = pair_mode != 0 && p.total_tiles >= 2 && sm_count() >= 2 && (sm_count() & 1) == 0;
    rc = make_mat_tmap(&tmB, wpack, y.c, 9LL * cin_pad, use_pair ? y.c / 2 : y.c);
    if (rc) return rc;
    if (use_pair) return y.c == 64 ? launch_fprop_halo2<64>(tmA, tmB, p, st) : launch_fprop_halo2<128>(tmA, tmB, p, st);
    return y.c == 64 ? launch_fprop_halo<64>(tmA, tmB, p, st) : launch_fprop_halo<128>(tmA, tmB, p, st);
  }
  rc = make_act_tmap(&tmA, x, p.TW, p.TH, p.TN);
  if (rc) return rc;
  rc = make_mat_tmap(&tmB, wpack, y.c, 1LL * taps * cin_pad, BN);
  if (rc) return rc;
  switch (BN) {
    case 256: return launch_fprop<256>(tmA, tmB, p, st);
    case 128: return launch_fprop<128>(tmA, tmB, p, st);
    default: return launch_fprop<64>(tmA, tmB, p, st);
  }
}

extern "C" int cvb_conv3x3_fprop_fuses_bwd_stats(cvb_view x, cvb_view y, int taps) {
  if (x.n != y.n || x.h != y.h || x.w != y.w || (x.c % 16) != 0) return 0;
  return uses_tr64(x, y, taps) ? 1 : 0;
}
