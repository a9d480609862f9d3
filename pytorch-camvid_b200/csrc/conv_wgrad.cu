// Weight gradient of the 3x3 / stride 1 / pad 1 convolution on tcgen05.
//   reference: the weight-gradient half of aten::convolution_backward reached through loss.backward() (train.py:131)
//   for nn.Conv2d(cin, cout, 3, padding=1) in models/unet.py:11 and models/segnet.py:8.
//
//   dW[co][ci][tap] = sum over pixels p of  dy[p][co] * x[p + tap offset][ci]
//
// GEMM view (K = pixels): D_tap[M = ci][N = co] = X_tap^T * dY.  Both operands are "MN-major": the TMA box
// [64 pixels][64 channels] lands as 64 rows of 128 B (128B swizzle) and tcgen05 reads it transposed, so no
// transposed copy of the activations is ever materialised.
//   * "unit"   = one such 8 KB box. x units are fetched with the box shifted by the tap offset (TMA zero-fills outside
//                the image = the convolution padding); dy units are unshifted and shared by every tap of the item.
//   * M = 128 = two x units: two 64-channel chunks of one tap (cin >= 128) or two taps of a 64-channel input.
//   * a CTA owns one work item = (co tile, ci tile, tap group, K split) and keeps one fp32 accumulator per unit pair
//     in TMEM for its whole pixel range; split-K partials go to a workspace that a second kernel reduces and
//     transposes to the OIHW fp32 layout of nn.Conv2d.weight.grad (deterministic, no atomics).
#include "common.cuh"
#include "sm100.cuh"
#include "tma_host.h"

namespace cvb {

constexpr int kWgradThreads = 256;
constexpr int kUnitBytes = 64 * 128;
constexpr int kWgradSmemBudget = 200 * 1024;

struct WgradParams {
  int PW, PH, PN;                 // pixel patch of one K block (PW*PH*PN == 64)
  int tiles_w, tiles_h, tiles_n;  // patches per image row / column / batch
  int kt_total;                   // K blocks in the whole tensor
  int taps;                       // 9 or 1
  int cin_pad, cout_pad;
  int BN, CM, T;                  // co tile, 64-channel ci chunks per tap in M, taps per item
  int n_co_tiles, n_ci_tiles, n_tap_groups, splits;
  int units_pad;                  // x units per stage, rounded up to even
  int stages;
  float* ws;                      // [splits][taps][cin_pad][cout_pad]
};

__global__ void __launch_bounds__(kWgradThreads, 1)
conv_wgrad_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmDY,
                  const WgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int b_units = p.BN / 64;
  const int stage_bytes = (p.units_pad + b_units) * kUnitBytes;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + p.stages * stage_bytes);
  uint64_t* empty = full + p.stages;
  uint64_t* tfull = empty + p.stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // work item decode
  int item = blockIdx.x;
  const int co_tile = item % p.n_co_tiles;
  item /= p.n_co_tiles;
  const int ci_tile = item % p.n_ci_tiles;
  item /= p.n_ci_tiles;
  const int tg = item % p.n_tap_groups;
  const int split = item / p.n_tap_groups;
  const int t0 = tg * p.T;
  const int tcount = min(p.T, p.taps - t0);
  const int units = tcount * p.CM;
  const int accs = (units + 1) >> 1;
  const int kb_begin = static_cast<int>(1LL * p.kt_total * split / p.splits);
  const int kb_end = static_cast<int>(1LL * p.kt_total * (split + 1) / p.splits);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmDY);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < p.stages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(tfull, 1);
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
      // ------------------------------- TMA producer -------------------------------
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t tx_bytes = static_cast<uint32_t>(units + b_units) * kUnitBytes;
      for (int kb = kb_begin; kb < kb_end; ++kb) {
        int t = kb;
        const int pw0 = (t % p.tiles_w) * p.PW;
        t /= p.tiles_w;
        const int ph0 = (t % p.tiles_h) * p.PH;
        const int pn0 = (t / p.tiles_h) * p.PN;
        uint8_t* sA = smem + stage * stage_bytes;
        uint8_t* sB = sA + p.units_pad * kUnitBytes;
        mbar_wait(&empty[stage], phase ^ 1);
        mbar_expect_tx(&full[stage], tx_bytes);
        for (int u = 0; u < units; ++u) {
          const int tap = t0 + u / p.CM;
          const int chunk = ci_tile * p.CM + u % p.CM;
          const int dr = p.taps == 9 ? tap / 3 - 1 : 0;
          const int ds = p.taps == 9 ? tap % 3 - 1 : 0;
          tma_load_4d(sA + u * kUnitBytes, &tmX, &full[stage], chunk * 64, pw0 + ds, ph0 + dr, pn0);
        }
        for (int j = 0; j < b_units; ++j)
          tma_load_4d(sB + j * kUnitBytes, &tmDY, &full[stage], co_tile * p.BN + j * 64, pw0, ph0, pn0);
        if (++stage == p.stages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ------------------------------- MMA issuer ---------------------------------
      const uint32_t idesc = idesc_bf16_f32(128, p.BN, true, true);
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = kb_begin; kb < kb_end; ++kb) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        const uint32_t a_base = smem_u32(smem + stage * stage_bytes);
        const uint32_t b_base = a_base + p.units_pad * kUnitBytes;
        for (int j = 0; j < accs; ++j) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            // 16 pixel rows per MMA = 2 groups of 8 rows (SBO 1024 B); 64-channel atoms are one unit apart (LBO).
            const uint64_t adesc = smem_desc_sw128(a_base + (2 * j) * kUnitBytes + k * 2048, kUnitBytes, 1024);
            const uint64_t bdesc = smem_desc_sw128(b_base + k * 2048, kUnitBytes, 1024);
            umma_bf16(tmem_base + j * p.BN, adesc, bdesc, idesc, (kb > kb_begin || k > 0) ? 1u : 0u);
          }
        }
        umma_commit(&empty[stage]);
        if (++stage == p.stages) {
          stage = 0;
          phase ^= 1;
        }
      }
      umma_commit(tfull);
    }
  } else if (warp >= 4) {
    // --------------------------------- epilogue -----------------------------------
    const int ew = warp - 4;
    const int row = ew * 32 + lane;
    if (kb_end > kb_begin) {
      mbar_wait(tfull, 0);
      tc_fence_after();
    }
    for (int j = 0; j < accs; ++j) {
      const int u = 2 * j + (row >> 6);
      const bool valid = u < units;
      const int tap = t0 + u / p.CM;
      const int ci = (ci_tile * p.CM + u % p.CM) * 64 + (row & 63);
      float* dst = p.ws + ((static_cast<long long>(split) * p.taps + tap) * p.cin_pad + ci) * p.cout_pad +
                   co_tile * p.BN;
      for (int c0 = 0; c0 < p.BN; c0 += 32) {
        uint32_t r[32];
        if (kb_end > kb_begin) {
          tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + j * p.BN + c0, r);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int q = 0; q < 32; ++q) r[q] = 0;
        }
        if (valid) {
#pragma unroll
          for (int q = 0; q < 8; ++q)
            *reinterpret_cast<uint4*>(dst + c0 + q * 4) = make_uint4(r[q * 4], r[q * 4 + 1], r[q * 4 + 2], r[q * 4 + 3]);
        }
      }
    }
  }
  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// out[co][ci][tap] = sum_s ws[s][tap][ci][co]; a block transposes a 32(co) x 32(ci) x taps brick through smem so both
// the workspace reads and the OIHW writes are coalesced.
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float* __restrict__ ws, int splits, int taps,
                                                           int cin_pad, int cout_pad, int cout, int cin_eff,
                                                           float* __restrict__ out) {
  extern __shared__ float tile[];  // [taps][32][33]
  const int co0 = blockIdx.x * 32, ci0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  const long long split_stride = 1LL * taps * cin_pad * cout_pad;
  for (int tap = 0; tap < taps; ++tap) {
    for (int i = ty; i < 32; i += 8) {
      const int ci = ci0 + i, co = co0 + tx;
      float acc = 0.f;
      if (ci < cin_pad && co < cout_pad) {
        const float* src = ws + (1LL * tap * cin_pad + ci) * cout_pad + co;
        for (int s = 0; s < splits; ++s) acc += src[s * split_stride];
      }
      tile[(tap * 32 + i) * 33 + tx] = acc;
    }
  }
  __syncthreads();
  // write: for each co row, (ci, tap) is contiguous in OIHW
  const int row_elems = 32 * taps;
  for (int j = ty; j < 32; j += 8) {
    const int co = co0 + j;
    if (co >= cout) continue;
    for (int e = tx; e < row_elems; e += 32) {
      const int i = e / taps, tap = e - i * taps;
      const int ci = ci0 + i;
      if (ci < cin_eff) out[(1LL * co * cin_eff + ci) * taps + tap] = tile[(tap * 32 + i) * 33 + j];
    }
  }
}

struct WgradPlan {
  WgradParams p;
  int grid;
  int smem;
  long long ws_bytes;
};

static int plan_wgrad(const cvb_view& x, const cvb_view& dy, int taps, WgradPlan* plan) {
  CVB_REQUIRE(taps == 9 || taps == 1, CVB_ERR_INVALID_ARG, "conv_wgrad: taps must be 9 or 1 (got %d)", taps);
  CVB_REQUIRE(x.n == dy.n && x.h == dy.h && x.w == dy.w, CVB_ERR_INVALID_ARG,
              "conv_wgrad: x %dx%dx%d and dy %dx%dx%d spatial shapes differ", x.n, x.h, x.w, dy.n, dy.h, dy.w);
  CVB_REQUIRE((x.c % 64) == 0 && (dy.c % 64) == 0, CVB_ERR_UNSUPPORTED,
              "conv_wgrad: channels must be padded to multiples of 64 (cin %d, cout %d)", x.c, dy.c);
  WgradParams& p = plan->p;
  memset(&p, 0, sizeof(p));
  // K block = 64-pixel patch; pick the power-of-two shape that wastes the fewest rows on this image size
  double best = -1.0;
  for (int pn = 1; pn <= 64; pn *= 2)
    for (int ph = 1; ph * pn <= 64; ph *= 2) {
      int pw = 64 / (pn * ph);
      if (pw > 256) continue;
      long long tiles = 1LL * ((x.n + pn - 1) / pn) * ((x.h + ph - 1) / ph) * ((x.w + pw - 1) / pw);
      double eff = static_cast<double>(1LL * x.n * x.h * x.w) / static_cast<double>(tiles * 64);
      double score = eff - 1e-4 * (pn > 1) + 1e-6 * pw;
      if (score > best) {
        best = score;
        p.PW = pw; p.PH = ph; p.PN = pn;
      }
    }
  p.tiles_w = (x.w + p.PW - 1) / p.PW;
  p.tiles_h = (x.h + p.PH - 1) / p.PH;
  p.tiles_n = (x.n + p.PN - 1) / p.PN;
  long long kt = 1LL * p.tiles_w * p.tiles_h * p.tiles_n;
  CVB_REQUIRE(kt < (1LL << 31), CVB_ERR_UNSUPPORTED, "conv_wgrad: too many K blocks");
  p.kt_total = static_cast<int>(kt);
  p.taps = taps;
  p.cin_pad = x.c;
  p.cout_pad = dy.c;
  p.BN = (dy.c % 256 == 0) ? 256 : ((dy.c % 128 == 0) ? 128 : 64);
  p.CM = (x.c % 128 == 0) ? 2 : 1;
  const int max_accs = 512 / p.BN;
  const int max_units = p.BN == 256 ? 4 : 6;  // keeps a stage <= 64 KB so three stages fit
  int T = (max_units / p.CM);
  if (T > 2 * max_accs / p.CM) T = 2 * max_accs / p.CM;
  if (T > taps) T = taps;
  if (T < 1) T = 1;
  p.T = T;
  p.n_co_tiles = dy.c / p.BN;
  p.n_ci_tiles = x.c / (64 * p.CM);
  p.n_tap_groups = (taps + T - 1) / T;
  p.units_pad = ((T * p.CM + 1) / 2) * 2;
  const int stage_bytes = (p.units_pad + p.BN / 64) * kUnitBytes;
  p.stages = kWgradSmemBudget / stage_bytes;
  if (p.stages > 8) p.stages = 8;
  CVB_REQUIRE(p.stages >= 2, CVB_ERR_UNSUPPORTED, "conv_wgrad: stage of %d bytes does not pipeline", stage_bytes);
  const int items = p.n_co_tiles * p.n_ci_tiles * p.n_tap_groups;
  int splits = (2 * sm_count() + items - 1) / items;
  int max_splits = p.kt_total / 8;
  if (max_splits < 1) max_splits = 1;
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  p.splits = splits;
  plan->grid = items * splits;
  plan->smem = 1024 + p.stages * stage_bytes + 256;
  plan->ws_bytes = 1LL * splits * taps * x.c * dy.c * 4;
  return CVB_OK;
}

}  // namespace cvb

using namespace cvb;

extern "C" int64_t cvb_conv3x3_wgrad_workspace_bytes(cvb_view x, cvb_view dy, int taps) {
  WgradPlan plan;
  int rc = plan_wgrad(x, dy, taps, &plan);
  if (rc) return rc;
  return plan.ws_bytes;
}

extern "C" int cvb_conv3x3_wgrad(cvb_view x, cvb_view dy, int taps, float* dw, int cout, int cin, void* workspace,
                                 int64_t workspace_bytes, void* stream) {
  int rc = check_view(x, "conv_wgrad.x");
  if (rc) return rc;
  rc = check_view(dy, "conv_wgrad.dy");
  if (rc) return rc;
  CVB_REQUIRE(dw && workspace, CVB_ERR_INVALID_ARG, "conv_wgrad: null pointer");
  WgradPlan plan;
  rc = plan_wgrad(x, dy, taps, &plan);
  if (rc) return rc;
  const int cin_eff = taps == 9 ? cin : cin * 9;
  CVB_REQUIRE(cout > 0 && cout <= dy.c && cin > 0 && cin_eff <= x.c, CVB_ERR_INVALID_ARG,
              "conv_wgrad: cout %d / cin %d do not fit the padded views (%d / %d)", cout, cin, dy.c, x.c);
  CVB_REQUIRE(workspace_bytes >= plan.ws_bytes, CVB_ERR_INVALID_ARG, "conv_wgrad: workspace %lld < required %lld",
              (long long)workspace_bytes, plan.ws_bytes);
  CVB_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 15) == 0, CVB_ERR_INVALID_ARG, "conv_wgrad: workspace not 16-byte aligned");
  plan.p.ws = static_cast<float*>(workspace);
  CUtensorMap tmX, tmDY;
  rc = make_act_tmap(&tmX, x, plan.p.PW, plan.p.PH, plan.p.PN);
  if (rc) return rc;
  rc = make_act_tmap(&tmDY, dy, plan.p.PW, plan.p.PH, plan.p.PN);
  if (rc) return rc;
  static bool configured = false;
  if (!configured) {
    CVB_CUDA(cudaFuncSetAttribute(conv_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  1024 + kWgradSmemBudget + 256));
    configured = true;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  conv_wgrad_kernel<<<plan.grid, kWgradThreads, plan.smem, st>>>(tmX, tmDY, plan.p);
  CVB_LAUNCH_CHECK();
  dim3 rgrid((dy.c + 31) / 32, (x.c + 31) / 32);
  wgrad_reduce_kernel<<<rgrid, 256, taps * 32 * 33 * sizeof(float), st>>>(plan.p.ws, plan.p.splits, taps, x.c, dy.c,
                                                                         cout, cin_eff, dw);
  CVB_LAUNCH_CHECK();
  return CVB_OK;
}
