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
#include "common.cuh"
#include "sm100.cuh"
#include "tma_host.h"

namespace cvb {

struct FpropParams {
  int N, H, W;
  int cin_chunks, taps, cin_pad, cout_pad;
  int TW, TH, TN;
  int tiles_w, tiles_h, tiles_n, n_tiles, total_tiles;
  uint32_t a_bytes;
  __nv_bfloat16* y;
  long long ysn, ysh, ysw;
  float* stat_partials;
  const float* scale;
  const float* shift;
  int relu;
};

constexpr int kFpropThreads = 256;
constexpr int kABytes = 128 * 128;  // 128 pixel rows x 64 bf16
constexpr int kStatFloats = 2 * 1024;

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
    if (lane == 0) {
      // ------------------------------- MMA issuer ---------------------------------
      constexpr uint32_t idesc = idesc_bf16_f32(128, BN, false, false);
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
          const uint32_t a_addr = smem_u32(sA + stage * kABytes);
          const uint32_t b_addr = smem_u32(sB + stage * Cfg::kBBytes);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            umma_bf16(d, smem_desc_sw128(a_addr + k * 32, 16, 1024), smem_desc_sw128(b_addr + k * 32, 16, 1024),
                      idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty[stage]);
          if (++stage == S) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&tfull[acc]);
      }
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
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t r[32];
        tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + acc * BN + c0, r);
        tmem_ld_wait();
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
        if (p.scale) {
          const float* sc = p.scale + nt * BN + c0;
          const float* sh = p.shift + nt * BN + c0;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            v[j] = fmaf(v[j], __ldg(sc + j), __ldg(sh + j));
            if (p.relu) v[j] = fmaxf(v[j], 0.f);
          }
        }
        if (valid) {
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
          atomicAdd(&s_stats[nt * BN + c0 + lane], s1);
          atomicAdd(&s_stats[p.cout_pad + nt * BN + c0 + lane], s2);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
    }
  }

  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (p.stat_partials) {
    float* dstp = p.stat_partials + static_cast<long long>(blockIdx.x) * 2 * p.cout_pad;
    for (int i = threadIdx.x; i < 2 * p.cout_pad; i += kFpropThreads) dstp[i] = s_stats[i];
  }
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
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
  CVB_REQUIRE((x.c % 64) == 0 && (y.c % 64) == 0, CVB_ERR_UNSUPPORTED,
              "conv_fprop: channels must be padded to multiples of 64 (cin %d, cout %d)", x.c, y.c);
  CVB_REQUIRE(y.c <= 1024, CVB_ERR_UNSUPPORTED, "conv_fprop: cout %d > 1024", y.c);

  FpropParams p;
  memset(&p, 0, sizeof(p));
  p.N = x.n; p.H = x.h; p.W = x.w;
  p.cin_pad = x.c; p.cout_pad = y.c;
  p.cin_chunks = x.c / 64;
  p.taps = taps;
  pick_tile(x.n, x.h, x.w, &p.TW, &p.TH, &p.TN);
  p.tiles_w = (x.w + p.TW - 1) / p.TW;
  p.tiles_h = (x.h + p.TH - 1) / p.TH;
  p.tiles_n = (x.n + p.TN - 1) / p.TN;
  const int BN = (y.c % 256 == 0) ? 256 : ((y.c % 128 == 0) ? 128 : 64);
  p.n_tiles = y.c / BN;
  long long total = 1LL * p.tiles_w * p.tiles_h * p.tiles_n * p.n_tiles;
  CVB_REQUIRE(total < (1LL << 31), CVB_ERR_UNSUPPORTED, "conv_fprop: too many tiles");
  p.total_tiles = static_cast<int>(total);
  p.a_bytes = static_cast<uint32_t>(p.TW * p.TH * p.TN) * 128u;
  p.y = static_cast<__nv_bfloat16*>(y.ptr);
  p.ysn = y.sn; p.ysh = y.sh; p.ysw = y.sw;
  if (ep) {
    p.stat_partials = ep->stat_partials;
    p.scale = ep->scale;
    p.shift = ep->shift;
    p.relu = ep->relu;
    CVB_REQUIRE((ep->scale == nullptr) == (ep->shift == nullptr), CVB_ERR_INVALID_ARG,
                "conv_fprop: scale and shift must both be given or both be NULL");
  }
  CUtensorMap tmA, tmB;
  rc = make_act_tmap(&tmA, x, p.TW, p.TH, p.TN);
  if (rc) return rc;
  rc = make_mat_tmap(&tmB, wpack, y.c, 1LL * taps * x.c, BN);
  if (rc) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (BN) {
    case 256: return launch_fprop<256>(tmA, tmB, p, st);
    case 128: return launch_fprop<128>(tmA, tmB, p, st);
    default: return launch_fprop<64>(tmA, tmB, p, st);
  }
}
