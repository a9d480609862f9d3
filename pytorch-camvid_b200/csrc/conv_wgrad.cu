// Weight gradient of the 3x3 / stride 1 / pad 1 convolution on tcgen05.
//   reference: the weight-gradient half of aten::convolution_backward reached through loss.backward() (train.py:131)
//   for nn.Conv2d(cin, cout, 3, padding=1) in models/unet.py:11 and models/segnet.py:8.
//
//   dW[co][ci][tap] = sum over pixels p of  dy[p][co] * x[p + tap offset][ci]
//
// GEMM view (K = pixels): D_tap[M = ci][N = co] = X_tap^T * dY.  Both operands are "MN-major": a pixel is one 128-byte
// shared-memory row of 64 channels (128B swizzle, exactly what TMA writes) and tcgen05 reads it transposed, so no
// transposed copy of the activations is ever materialised.
//   * pixel tile = 8 (w) x 16 (h); one K=16 MMA slice = two tile rows of 8 pixels (SBO = row pitch).
//   * the activation patch of a tile (10 x 18 pixels, halo included, zero-filled by TMA outside the image = the
//     convolution padding) is fetched ONCE per 64-channel chunk; every tap is a shifted descriptor view of it (start
//     row (dr*10 + ds), any 128-byte row is a legal start: the swizzle is a function of the absolute smem address,
//     tools/exp/exp_desc.cu). The dy tile is fetched once and shared by all taps.
//   * M = 128 = two 64-row atoms: two input-channel chunks of one tap (cin >= 128; LBO = chunk stride in smem) or two
//     taps of a 64-channel input (LBO = distance between the two tap views).
//   * work item = (co tile, ci tile, tap group): one fp32 accumulator per atom pair in TMEM, up to 512 columns.
//   * stream-K partition: the (item, pixel tile) space is flattened and cut into one EQUAL contiguous range per CTA
//     (grid = number of SMs, a single balanced wave whatever the item count). A CTA that crosses an item boundary
//     drains its accumulators to the workspace (one "part" of that item) and carries on with the next item. Parts are
//     summed and transposed to the OIHW fp32 layout of nn.Conv2d.weight.grad by two small kernels (deterministic, no
//     atomics).
#include "common.cuh"
#include "sm100.cuh"
#include "tma_host.h"

namespace cvb {

constexpr int kWgradThreads = 256;
constexpr int kWTileW = 8, kWTileH = 16;
constexpr int kWPitch = kWTileW + 2;                       // patch row pitch in pixels
constexpr int kWPatchRows = (kWTileH + 2) * kWPitch;      // 180
constexpr int kWPatchBytes = kWPatchRows * 128;           // 23040
constexpr int kWPatchStride = 23552;                      // rounded up to 1 KB
constexpr int kWDyBytes = kWTileW * kWTileH * 128;        // 16384 per 64-channel unit
constexpr int kWgradSmemBudget = 225 * 1024;
constexpr int kMaxAccs = 8;
constexpr int kMaxWaves = 4;

struct WgradParams {
  int N, H, W;
  int tiles_w, tiles_h, total_tiles;
  int th;    // pixel-tile height (even, <= kWTileH): 16 unless a shorter tile wastes much less of a small image
  int taps;  // 9 or 1
  int cin_pad, cout_pad;
  int CM, T;  // 64-channel ci chunks per M=128 accumulator (1 or 2), taps per work item
  int n_co_tiles, n_ci_tiles, n_tap_groups, items;
  int grid;               // CTAs = ranges of the flattened (item, tile) space
  long long total_units;  // items * total_tiles
  // Waves (conv_wgrad_kernel): the items are processed in up to kMaxWaves groups, each a stream-K partition of its own
  // whose grid is a multiple of its item count (lockstep, see plan_wgrad); CTA c takes part in wave w if c < wave_grid[w].
  int n_waves;
  int wave_item0[4], wave_items[4], wave_grid[4];
  int slots;              // workspace parts per item
  int stages, stage_bytes;
  // cluster == 2 (conv_wgrad_kernel only): CTAs 2v, 2v + 1 form a cluster and play ONE range of the partition (virtual CTA
  // v) on two neighbouring ci tiles (2t, 2t + 1) of the same (co tile, tap group): they read the same dy tiles, so each
  // fetches half of them and multicasts. Items, n_ci_tiles, grids and parts are then counted in virtual CTAs / tile pairs.
  int cluster;
  int pair;  // cluster == 2 only: 1 = the CTA-pair instantiation (cta_group::2), 0 = two single-CTA MMAs sharing dy by multicast
  int patch_stride, dy_bytes;  // conv_wgrad_kernel: bytes between the ci chunks of a stage's x patch / its 64-channel dy units
  float* ws;  // [slots][taps][cin_pad][cout_pad]; part k of an item lives in slot k (k < parts of that item)
};

// Stream-K bookkeeping shared by host and device: CTA c owns units [c*U/G, (c+1)*U/G).
__host__ __device__ inline long long sk_begin(long long U, int G, int c) { return U * c / G; }
__host__ __device__ inline int sk_owner(long long U, int G, long long u) {  // the CTA whose range contains unit u
  int c = static_cast<int>(u * G / U);
  while (c + 1 < G && sk_begin(U, G, c + 1) <= u) ++c;
  while (c > 0 && sk_begin(U, G, c) > u) --c;
  return c;
}

// Issues every K slice of one pixel tile for ACCS accumulators (compile-time: no per-MMA predication; a k-step is one
// 32-bit add on each descriptor low word).
template <int BN, int ACCS, bool PAIR>
__device__ __forceinline__ void issue_tile(uint32_t tmem_base, uint32_t x_lo, uint32_t d_lo, const uint32_t (&acc_lo)[kMaxAccs],
                                           int slices, uint32_t first) {
  constexpr uint32_t idesc = idesc_bf16_f32(PAIR ? 256 : 128, BN, true, true);
  constexpr uint32_t a_hi = desc_hi_sw128(kWPitch * 128);  // K groups: consecutive tile rows of the patch
  constexpr uint32_t b_hi = desc_hi_sw128(kWTileW * 128);  // dy tile rows are dense
#pragma unroll 1
  for (int s = 0; s < slices; ++s) {
    const uint32_t xs = x_lo + s * (2 * kWPitch * 8), ds = d_lo + s * (2 * kWTileW * 8);
    const uint32_t accum = (first | static_cast<uint32_t>(s)) != 0 ? 1u : 0u;
#pragma unroll
    for (int j = 0; j < ACCS; ++j) {
      if (PAIR) umma_bf16_lohi_pair(tmem_base + j * BN, xs + acc_lo[j], a_hi, ds, b_hi, idesc, accum);
      else umma_bf16_lohi(tmem_base + j * BN, xs + acc_lo[j], a_hi, ds, b_hi, idesc, accum);
    }
  }
}

// PAIR: the two CTAs of a cluster run ONE M = 256 MMA (cta_group::2) on the two ci tiles they own: each supplies its own x
// patch (its 128 A rows) and only HALF of the dy tile (BN / 2 of the N columns), the leader (cluster rank 0) issues, and
// each CTA's TMEM receives its own 128 accumulator rows -- the epilogue and the workspace layout do not change. What
// changes is the number of bytes DELIVERED to an SM per MAC (x + dy / 2 instead of x + dy: -28 % at BN = 256, -20 % at 128),
// which is what bounds this kernel (42-45 B/clk per SM when all SMs pull; multicast into both CTAs delivers the same bytes
// and gained nothing: profiles/r02z_wgrad_cluster.txt). Protocol as in conv_fprop_halo2_kernel: loads of both CTAs are
// credited to the LEADER's full barrier, the leader's commits are multicast to both CTAs' empty / tfull barriers, both
// epilogues report to the leader's tempty.
template <int BN, bool PAIR>
__global__ void __launch_bounds__(kWgradThreads, 1)
conv_wgrad_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmDY,
                  const WgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int b_units = PAIR ? BN / 128 : BN / 64;  // 64-channel dy boxes this CTA holds per stage
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + p.stages * p.stage_bytes);
  uint64_t* empty = full + p.stages;
  uint64_t* tfull = empty + p.stages;
  uint64_t* tempty = tfull + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int cl = p.cluster;
  const int crank = cl == 2 ? static_cast<int>(cluster_ctarank()) : 0;
  const int vcta = cl == 2 ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);

  // per wave: this (virtual) CTA's unit range [u_begin, u_end) of the wave's flattened (local item, tile) space
#define CVB_WAVE_RANGE(w)                                                                                         \
  const long long wave_units = 1LL * p.wave_items[w] * p.total_tiles;                                             \
  const bool in_wave = vcta < p.wave_grid[w];                                                                     \
  const long long u_begin = in_wave ? sk_begin(wave_units, p.wave_grid[w], vcta) : 0;                             \
  const long long u_end = in_wave ? sk_begin(wave_units, p.wave_grid[w], vcta + 1) : 0;                           \
  const int item_first = static_cast<int>(u_begin / p.total_tiles);                                               \
  const int item_last = u_end > u_begin ? static_cast<int>((u_end - 1) / p.total_tiles) : item_first - 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmDY);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < p.stages; ++i) {
      mbar_init(&full[i], PAIR ? 2 : 1);  // PAIR: the leader's expect_tx + the peer's "my loads are issued"
      // a stage is free when every CTA of the cluster that it was multicast to has consumed it (PAIR: one multicast commit)
      mbar_init(&empty[i], PAIR ? 1 : cl);
    }
    mbar_init(tfull, 1);
    mbar_init(tempty, PAIR ? 8 : 4);  // PAIR: the epilogue warps of both CTAs report to the leader
    fence_mbar_init();
  }
  if (warp == 2) {
    if (PAIR) {
      tmem_alloc_pair(tmem_slot, 512);
      tmem_relinquish_pair();
    } else {
      tmem_alloc(tmem_slot, 512);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (cl == 2) cluster_sync_all();  // the peer's barriers exist before anything is signalled across the cluster
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ------------------------------- TMA producer -------------------------------
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t tx_bytes = static_cast<uint32_t>(p.CM) * (p.th + 2) * kWPitch * 128 + b_units * p.th * kWTileW * 128;
      for (int w = 0; w < p.n_waves; ++w) {
      CVB_WAVE_RANGE(w)
      for (int litem = item_first; litem <= item_last; ++litem) {
        const int item = p.wave_item0[w] + litem;
        const int co_tile = item % p.n_co_tiles;
        const int ci_tile = ((item / p.n_co_tiles) % p.n_ci_tiles) * cl + crank;
        const long long base = 1LL * litem * p.total_tiles;
        const int tile_begin = static_cast<int>(max(u_begin, base) - base);
        const int tile_end = static_cast<int>(min(u_end, base + p.total_tiles) - base);
        for (int tile = tile_begin; tile < tile_end; ++tile) {
          int t = tile;
          const int w0 = (t % p.tiles_w) * kWTileW;
          t /= p.tiles_w;
          const int h0 = (t % p.tiles_h) * p.th;
          const int n0 = t / p.tiles_h;
          uint8_t* sX = smem + stage * p.stage_bytes;
          uint8_t* sD = sX + p.CM * p.patch_stride;
          mbar_wait(&empty[stage], phase ^ 1);
          if (PAIR) {
            if (crank == 0) mbar_expect_tx(&full[stage], 2 * tx_bytes);  // both CTAs' boxes are credited to the leader
            for (int c = 0; c < p.CM; ++c)
              tma_load_4d_pair(sX + c * p.patch_stride, &tmX, &full[stage], (ci_tile * p.CM + c) * 64, w0 - 1, h0 - 1, n0);
#pragma unroll
            for (int j = 0; j < b_units; ++j)  // this CTA's half of the N columns
              tma_load_4d_pair(sD + j * p.dy_bytes, &tmDY, &full[stage], co_tile * BN + crank * (BN / 2) + j * 64, w0, h0, n0);
            if (crank != 0) mbar_arrive_leader(&full[stage]);
          } else {
            mbar_expect_tx(&full[stage], tx_bytes);
            for (int c = 0; c < p.CM; ++c)
              tma_load_4d(sX + c * p.patch_stride, &tmX, &full[stage], (ci_tile * p.CM + c) * 64, w0 - 1, h0 - 1, n0);
#pragma unroll
            for (int j = 0; j < b_units; ++j) {
              if (cl == 1)
                tma_load_4d(sD + j * p.dy_bytes, &tmDY, &full[stage], co_tile * BN + j * 64, w0, h0, n0);
              else if ((j & 1) == crank)  // this CTA's half of the dy tile, delivered to both CTAs
                tma_load_4d_mcast(sD + j * p.dy_bytes, &tmDY, &full[stage], co_tile * BN + j * 64, w0, h0, n0, 3);
            }
          }
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
      }
    }
  } else if (warp == 1 && (!PAIR || crank == 0)) {
    // ------------------------------- MMA issuer (whole warp, one elected lane issues; PAIR: the leader CTA only) -------
    const uint32_t x_lo0 = desc_lo(smem_u32(smem), 0);
    const uint32_t d_lo0 = desc_lo(smem_u32(smem) + p.CM * p.patch_stride, p.dy_bytes);
    const uint32_t stage_lo = static_cast<uint32_t>(p.stage_bytes) >> 4;
    int stage = 0;
    uint32_t phase = 0;
    int seg = 0;
    for (int w = 0; w < p.n_waves; ++w) {
    CVB_WAVE_RANGE(w)
    for (int litem = item_first; litem <= item_last; ++litem, ++seg) {
      const int item = p.wave_item0[w] + litem;
      const int tg = item / (p.n_co_tiles * p.n_ci_tiles);
      const int t0 = tg * p.T;
      const int tcount = min(p.T, p.taps - t0);
      const int accs = (tcount * p.CM + 1) >> 1;
      // per accumulator: descriptor low word of its first atom relative to the stage base (row offset of the tap
      // view, in 16-byte units) with the LBO field = distance to the second atom
      uint32_t acc_lo[kMaxAccs];
#pragma unroll
      for (int j = 0; j < kMaxAccs; ++j) {
        const int tap_a = p.CM == 2 ? t0 + j : t0 + 2 * j;
        const int tap_b = p.CM == 2 ? tap_a : tap_a + 1;
        const int ra = p.taps == 9 ? (tap_a / 3) * kWPitch + tap_a % 3 : kWPitch + 1;
        const int rb = p.taps == 9 ? (tap_b / 3) * kWPitch + tap_b % 3 : kWPitch + 2;
        // second atom: the other ci chunk of the same tap (CM == 2) or the next tap's view of the same chunk
        const uint32_t lbo = p.CM == 2 ? static_cast<uint32_t>(p.patch_stride) : static_cast<uint32_t>(rb - ra) * 128;
        acc_lo[j] = static_cast<uint32_t>(ra) * 8 + ((lbo >> 4) << 16);
      }
      const long long base = 1LL * litem * p.total_tiles;
      const int tile_begin = static_cast<int>(max(u_begin, base) - base);
      const int tile_end = static_cast<int>(min(u_end, base + p.total_tiles) - base);
      // the epilogue must have drained the previous item's accumulators
      mbar_wait(tempty, (seg & 1) ^ 1);
      tc_fence_after();
      for (int tile = tile_begin; tile < tile_end; ++tile) {
        const int h0 = ((tile / p.tiles_w) % p.tiles_h) * p.th;
        const int rows = min(p.th, p.H - h0);
        const int slices = (rows + 1) >> 1;  // K slices of two tile rows that touch the image
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        const uint32_t x_lo = x_lo0 + stage * stage_lo, d_lo = d_lo0 + stage * stage_lo;
        const uint32_t first = tile != tile_begin ? 1u : 0u;
        if (elect_one()) {
          switch (accs) {
            case 1: issue_tile<BN, 1, PAIR>(tmem_base, x_lo, d_lo, acc_lo, slices, first); break;
            case 2: issue_tile<BN, 2, PAIR>(tmem_base, x_lo, d_lo, acc_lo, slices, first); break;
            case 3: if (3 * BN <= 512) issue_tile<BN, 3, PAIR>(tmem_base, x_lo, d_lo, acc_lo, slices, first); break;
            case 4: if (4 * BN <= 512) issue_tile<BN, 4, PAIR>(tmem_base, x_lo, d_lo, acc_lo, slices, first); break;
            case 5: if (5 * BN <= 512) issue_tile<BN, 5, PAIR>(tmem_base, x_lo, d_lo, acc_lo, slices, first); break;
            case 6: if (6 * BN <= 512) issue_tile<BN, 6, PAIR>(tmem_base, x_lo, d_lo, acc_lo, slices, first); break;
            case 7: if (7 * BN <= 512) issue_tile<BN, 7, PAIR>(tmem_base, x_lo, d_lo, acc_lo, slices, first); break;
            default: if (8 * BN <= 512) issue_tile<BN, 8, PAIR>(tmem_base, x_lo, d_lo, acc_lo, slices, first); break;
          }
          if (PAIR) umma_commit_pair(&empty[stage]);
          else if (cl == 1) umma_commit(&empty[stage]);
          else umma_commit_mcast(&empty[stage], 3);
        }
        __syncwarp();
        if (++stage == p.stages) {
          stage = 0;
          phase ^= 1;
        }
      }
      if (elect_one()) {
        if (PAIR) umma_commit_pair(tfull);
        else umma_commit(tfull);
      }
      __syncwarp();
    }
    }
  } else if (warp >= 4) {
    // --------------------------------- epilogue (once per item segment) -----------------------------------
    const int ew = warp - 4;
    const int row = ew * 32 + lane;
    const int atom = row >> 6, r = row & 63;
    int seg = 0;
    for (int w = 0; w < p.n_waves; ++w) {
    CVB_WAVE_RANGE(w)
    for (int litem = item_first; litem <= item_last; ++litem, ++seg) {
      const int item = p.wave_item0[w] + litem;
      const int co_tile = item % p.n_co_tiles;
      const int ci_tile = ((item / p.n_co_tiles) % p.n_ci_tiles) * cl + crank;
      const int tg = item / (p.n_co_tiles * p.n_ci_tiles);
      const int t0 = tg * p.T;
      const int tcount = min(p.T, p.taps - t0);
      const int units = tcount * p.CM;
      const int accs = (units + 1) >> 1;
      // this CTA's part number within the item = distance from the (virtual) CTA that owns the item's first unit
      const int part = vcta - sk_owner(wave_units, p.wave_grid[w], 1LL * litem * p.total_tiles);
      mbar_wait(tfull, seg & 1);
      tc_fence_after();
      for (int j = 0; j < accs; ++j) {
        int tap, ci;
        bool valid;
        if (p.CM == 2) {
          tap = t0 + j;
          ci = (ci_tile * 2 + atom) * 64 + r;
          valid = true;
        } else {
          const int u = 2 * j + atom;
          valid = u < units;
          tap = t0 + u;
          ci = ci_tile * 64 + r;
        }
        float* dst = p.ws + ((static_cast<long long>(part) * p.taps + tap) * p.cin_pad + ci) * p.cout_pad + co_tile * BN;
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
          uint32_t v[32];
          tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + j * BN + c0, v);
          tmem_ld_wait();
          if (valid) {
#pragma unroll
            for (int q = 0; q < 8; ++q)
              *reinterpret_cast<uint4*>(dst + c0 + q * 4) = make_uint4(v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (PAIR) mbar_arrive_leader(tempty);
        else mbar_arrive(tempty);
      }
    }
    }
  }
#undef CVB_WAVE_RANGE
  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (cl == 2) cluster_sync_all();  // neither CTA leaves while the peer's multicasts / commits may still target it
  if (warp == 2) {
    tc_fence_after();
    if (PAIR) tmem_dealloc_pair(tmem_base, 512);
    else tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Row-shift variant for cout_pad == 64. An MMA pays for every operand row it reads from shared memory, so N = 64 leaves
// the tensor pipe half idle (ncu: 50 % active on these layers). Here the ROW offsets of the taps move to the dy operand:
//   dW[co][ci][(dr, dc)] = sum_q  x[q + (0, dc)][ci] * dy[q - (dr, 0)][co]
// A = the x patch (16 rows x 10 pixels, column halo only) viewed at column offset dc, two views per M = 128;
// B = the dy patch (18 rows x 8 pixels, row halo, zero-filled by TMA outside the image) as N = 192 = three 64-wide atoms,
// atom j = the patch shifted down by j rows = tap row dr = 1 - j. Two MMAs per K slice (views {dc0, dc1} and {dc1, dc2};
// the duplicated dc1 rows are dropped by the epilogue) replace five N = 64 MMAs. One work item = one 64-channel ci chunk;
// stream-K partition, workspace layout and the reduction kernels are those of the kernel above.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kRsXRows = kWTileH * kWPitch;          // 160
constexpr int kRsXBytes = kRsXRows * 128;            // 20480
constexpr int kRsDRows = (kWTileH + 2) * kWTileW;    // 144
constexpr int kRsDBytes = kRsDRows * 128;            // 18432
constexpr int kRsStageBytes = kRsXBytes + kRsDBytes; // 38912
constexpr int kRsStages = 5;

__global__ void __launch_bounds__(kWgradThreads, 1)
conv_wgrad_rs64_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmDY,
                       const WgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int S = kRsStages;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + S * kRsStageBytes);
  uint64_t* empty = full + S;
  uint64_t* tfull = empty + S;
  uint64_t* tempty = tfull + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const long long u_begin = sk_begin(p.total_units, p.grid, blockIdx.x);
  const long long u_end = sk_begin(p.total_units, p.grid, blockIdx.x + 1);
  const int item_first = static_cast<int>(u_begin / p.total_tiles);
  const int item_last = u_end > u_begin ? static_cast<int>((u_end - 1) / p.total_tiles) : item_first - 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmDY);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < S; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(tfull, 1);
    mbar_init(tempty, 4);
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
      for (int item = item_first; item <= item_last; ++item) {
        const long long base = 1LL * item * p.total_tiles;
        const int tile_begin = static_cast<int>(max(u_begin, base) - base);
        const int tile_end = static_cast<int>(min(u_end, base + p.total_tiles) - base);
        for (int tile = tile_begin; tile < tile_end; ++tile) {
          int t = tile;
          const int w0 = (t % p.tiles_w) * kWTileW;
          t /= p.tiles_w;
          const int h0 = (t % p.tiles_h) * kWTileH;
          const int n0 = t / p.tiles_h;
          uint8_t* sX = smem + stage * kRsStageBytes;
          mbar_wait(&empty[stage], phase ^ 1);
          mbar_expect_tx(&full[stage], kRsStageBytes);
          tma_load_4d(sX, &tmX, &full[stage], item * 64, w0 - 1, h0, n0);
          tma_load_4d(sX + kRsXBytes, &tmDY, &full[stage], 0, w0, h0 - 1, n0);
          if (++stage == S) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------- MMA issuer -------------------------------
    constexpr uint32_t idesc = idesc_bf16_f32(128, 192, true, true);
    constexpr uint32_t a_hi = desc_hi_sw128(kWPitch * 128);  // K groups: consecutive rows of the x patch
    constexpr uint32_t b_hi = desc_hi_sw128(kWTileW * 128);  // dy patch rows are dense
    // LBO: A atoms = neighbouring column views (one pixel = 128 B apart); B atoms = the patch one row (8 pixels) further down
    const uint32_t x_lo0 = desc_lo(smem_u32(smem), 128);
    const uint32_t d_lo0 = desc_lo(smem_u32(smem) + kRsXBytes, kWTileW * 128);
    constexpr uint32_t stage_lo = kRsStageBytes >> 4;
    int stage = 0;
    uint32_t phase = 0;
    int seg = 0;
    for (int item = item_first; item <= item_last; ++item, ++seg) {
      const long long base = 1LL * item * p.total_tiles;
      const int tile_begin = static_cast<int>(max(u_begin, base) - base);
      const int tile_end = static_cast<int>(min(u_end, base + p.total_tiles) - base);
      mbar_wait(tempty, (seg & 1) ^ 1);
      tc_fence_after();
      for (int tile = tile_begin; tile < tile_end; ++tile) {
        const int h0 = ((tile / p.tiles_w) % p.tiles_h) * kWTileH;
        const int rows = min(kWTileH, p.H - h0);
        const int slices = (rows + 1) >> 1;
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        const uint32_t x_lo = x_lo0 + stage * stage_lo, d_lo = d_lo0 + stage * stage_lo;
        const uint32_t first = tile != tile_begin ? 1u : 0u;
        if (elect_one()) {
#pragma unroll 1
          for (int s = 0; s < slices; ++s) {
            const uint32_t xs = x_lo + s * (2 * kWPitch * 8), ds = d_lo + s * (2 * kWTileW * 8);
            const uint32_t accum = (first | static_cast<uint32_t>(s)) != 0 ? 1u : 0u;
            umma_bf16_lohi(tmem_base, xs, a_hi, ds, b_hi, idesc, accum);            // column views dc = 0, 1
            umma_bf16_lohi(tmem_base + 192, xs + 8, a_hi, ds, b_hi, idesc, accum);  // column views dc = 1, 2
          }
          umma_commit(&empty[stage]);
        }
        __syncwarp();
        if (++stage == S) {
          stage = 0;
          phase ^= 1;
        }
      }
      if (elect_one()) umma_commit(tfull);
      __syncwarp();
    }
  } else if (warp >= 4) {
    // --------------------------------- epilogue (once per item segment) -----------------------------------
    const int ew = warp - 4;
    const int row = ew * 32 + lane;
    const int atom = row >> 6, r = row & 63;
    int seg = 0;
    for (int item = item_first; item <= item_last; ++item, ++seg) {
      const int part = static_cast<int>(blockIdx.x) - sk_owner(p.total_units, p.grid, 1LL * item * p.total_tiles);
      const int ci = item * 64 + r;
      mbar_wait(tfull, seg & 1);
      tc_fence_after();
#pragma unroll 1
      for (int a = 0; a < 2; ++a) {
        const int dc = a + atom;             // accumulator 0: views (0, 1); accumulator 1: views (1, 2)
        const bool keep = !(a == 1 && atom == 0);  // view 1 was already taken from accumulator 0
#pragma unroll 1
        for (int j = 0; j < 3; ++j) {
          const int tap = (2 - j) * 3 + dc;  // atom j of N = tap row dr = 1 - j
          float* dst = p.ws + ((static_cast<long long>(part) * p.taps + tap) * p.cin_pad + ci) * p.cout_pad;
#pragma unroll 1
          for (int c0 = 0; c0 < 64; c0 += 32) {
            uint32_t v[32];
            tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + a * 192 + j * 64 + c0, v);
            tmem_ld_wait();
            if (keep) {
#pragma unroll
              for (int q = 0; q < 8; ++q)
                *reinterpret_cast<uint4*>(dst + c0 + q * 4) = make_uint4(v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty);
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

// Number of workspace parts (= CTAs whose unit range touched it) of a work item: the same integer arithmetic the main
// kernel uses for its ranges.
__device__ __forceinline__ int item_parts(const WgradParams& p, int item) {
  int w = 0;
  while (w + 1 < p.n_waves && item >= p.wave_item0[w + 1]) ++w;
  const long long wu = 1LL * p.wave_items[w] * p.total_tiles, u0 = 1LL * (item - p.wave_item0[w]) * p.total_tiles;
  return sk_owner(wu, p.wave_grid[w], u0 + p.total_tiles - 1) - sk_owner(wu, p.wave_grid[w], u0) + 1;
}

// out[co][ci][tap] = sum over the parts s of the item that owns the element of ws[s][tap][ci][co].  ONE kernel for the
// split-K reduction and the transposition to nn.Conv2d.weight.grad's OIHW layout (round 1 ran two, 46 launches per UNet
// step). A block owns a brick of 32 co x BCI ci x all taps: it sums the parts with coalesced 128-byte reads (co fastest;
// the part count is a property of the (tap group, ci tile, co tile) item and a brick lies inside one ci tile and one co
// tile: at most n_tap_groups items per brick, their part counts are worked out once per block), transposes the brick
// through shared memory and writes runs of BCI*taps contiguous floats per co row. The brick is kept in OIHW order
// ([ci][tap] rows of 33 floats) so that both the stores of the first phase and the loads of the second are free of bank
// conflicts and the second phase needs no index arithmetic (the [tap][ci] order cost 8-way conflicts and a division per
// element). BCI shrinks for small layers so that the grid still fills the GPU. Deterministic: fixed summation order, no
// atomics.
// Items that stream-K cut into many parts (the cout = 64 layers: ONE item per 64 input channels, a part per SM) would make
// a thread walk 74-148 parts one dependent load after another: there the block's warps split the PARTS between them
// (`ps` part groups, chosen per block from its items' part counts, at most ps_max = what the host sized the brick
// buffer for) and the second phase adds the groups up.
template <int TAPS>
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float* __restrict__ ws, const WgradParams p, int BN,
                                                           int cout, int cin_eff, int x_c, int bci, int ps_max,
                                                           float* __restrict__ out) {
  extern __shared__ float tile[];  // [ps][bci][TAPS][33]
  __shared__ int s_parts[16];
  const int cin_pad = p.cin_pad, cout_pad = p.cout_pad;
  const int co0 = blockIdx.x * 32, ci0 = blockIdx.y * bci;
  const int tx = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x < p.n_tap_groups && threadIdx.x < 16) {
    const int cl = p.cluster > 1 ? p.cluster : 1;  // cluster mode: an item = a PAIR of ci tiles
    const int item = (static_cast<int>(threadIdx.x) * p.n_ci_tiles + ci0 / (64 * p.CM * cl)) * p.n_co_tiles + co0 / BN;
    s_parts[threadIdx.x] = item_parts(p, item);
  }
  __syncthreads();
  int max_parts = 1;
  for (int g = 0; g < p.n_tap_groups && g < 16; ++g) max_parts = max(max_parts, s_parts[g]);
  int ps = max_parts >= 32 ? 8 : (max_parts >= 8 ? 4 : (max_parts >= 3 ? 2 : 1));
  if (ps > ps_max) ps = ps_max;
  const int pg = warp % ps, wr = warp / ps, row_step = (8 / ps) * 4;  // part group / row group of this warp
  const long long ss = (1LL * TAPS * cin_pad * cout_pad) >> 2;  // part stride in float4
  const int elems = TAPS * bci;  // (tap, i) pairs, each a 32-wide co row
  const int gstride = elems * 33;
  // 16-byte loads: a thread owns 4 consecutive co of one (tap, ci) row; a warp covers 4 rows x 128 bytes per load; three
  // rows per thread are in flight, then up to four further parts of a row
  const int q = tx & 7, rsub = tx >> 3;
  const int co = co0 + q * 4;
  auto add = [](float4& a, const float4& v) { a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w; };
  for (int e0 = wr * 4 + rsub; e0 < elems; e0 += 3 * row_step) {
    float4 acc[3];
    const float4* src[3];
    int parts[3], slot[3];
#pragma unroll
    for (int u = 0; u < 3; ++u) {
      const int e = e0 + u * row_step;
      acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      parts[u] = 0;
      slot[u] = -1;
      src[u] = nullptr;
      if (e < elems) {
        const int tap = e / bci, i = e - tap * bci;
        const int ci = ci0 + i;
        slot[u] = pg * gstride + (i * TAPS + tap) * 33 + q * 4;
        if (ci < cin_pad && ci < x_c && co < cout_pad) {
          parts[u] = s_parts[tap / p.T];
          src[u] = reinterpret_cast<const float4*>(ws + (1LL * tap * cin_pad + ci) * cout_pad + co);
          if (pg < parts[u]) acc[u] = __ldcs(src[u] + pg * ss);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < 3; ++u) {
      if (parts[u] > pg + ps) {  // split items: this group's remaining parts, four in flight
        float4 a1 = make_float4(0.f, 0.f, 0.f, 0.f), a2 = a1, a3 = a1;
        int sidx = pg + ps;
        for (; sidx + 3 * ps < parts[u]; sidx += 4 * ps) {
          const float4 v0 = __ldcs(src[u] + sidx * ss), v1 = __ldcs(src[u] + (sidx + ps) * ss),
                       v2 = __ldcs(src[u] + (sidx + 2 * ps) * ss), v3 = __ldcs(src[u] + (sidx + 3 * ps) * ss);
          add(acc[u], v0); add(a1, v1); add(a2, v2); add(a3, v3);
        }
        for (; sidx < parts[u]; sidx += ps) add(acc[u], __ldcs(src[u] + sidx * ss));
        add(a1, a3);
        add(acc[u], a2);
        add(acc[u], a1);
      }
      if (slot[u] >= 0) {
        float* t = tile + slot[u];
        t[0] = acc[u].x; t[1] = acc[u].y; t[2] = acc[u].z; t[3] = acc[u].w;
      }
    }
  }
  __syncthreads();
  // OIHW: for a fixed co the brick's (ci, tap) elements are contiguous, in the order the brick is stored
  const int ci_n = min(bci, cin_eff - ci0);
  const int row_elems = ci_n * TAPS;
  for (int j = warp; j < 32; j += 8) {
    const int c = co0 + j;
    if (c >= cout) continue;
    float* dst = out + (1LL * c * cin_eff + ci0) * TAPS;
    for (int e = tx; e < row_elems; e += 32) {
      float v = tile[e * 33 + j];
      for (int g = 1; g < ps; ++g) v += tile[g * gstride + e * 33 + j];
      dst[e] = v;
    }
  }
}

struct WgradPlan {
  WgradParams p;
  int BN;
  bool rowshift;  // cout_pad == 64: conv_wgrad_rs64_kernel
  int grid;
  int smem;
  long long ws_bytes;
};

static int wgrad_max_clusters(int BN, bool pair);

static int plan_wgrad(const cvb_view& x, const cvb_view& dy, int taps, WgradPlan* plan) {
  CVB_REQUIRE(taps == 9 || taps == 1, CVB_ERR_INVALID_ARG, "conv_wgrad: taps must be 9 or 1 (got %d)", taps);
  CVB_REQUIRE(x.n == dy.n && x.h == dy.h && x.w == dy.w, CVB_ERR_INVALID_ARG,
              "conv_wgrad: x %dx%dx%d and dy %dx%dx%d spatial shapes differ", x.n, x.h, x.w, dy.n, dy.h, dy.w);
  CVB_REQUIRE((x.c % 16) == 0 && (dy.c % 16) == 0, CVB_ERR_UNSUPPORTED,
              "conv_wgrad: cin and cout must be multiples of 16 (cin %d, cout %d)", x.c, dy.c);
  // dy may hold fewer channels than the 64-padded GEMM (12 classes -> 16 channels in memory): zero fill as for x; only the
  // cout = 64 row-shift kernel is used that way
  const int cout_pad = (dy.c + 63) / 64 * 64;
  // x may stop short of its last 64-channel chunk (the im2col'd first layer keeps 32 channels in memory): TMA zero-fills
  // the rest of the box, the corresponding gradient rows come out as zeros and are never copied out
  const int cin_pad = (x.c + 63) / 64 * 64;
  WgradParams& p = plan->p;
  memset(&p, 0, sizeof(p));
  p.N = x.n; p.H = x.h; p.W = x.w;
  p.tiles_w = (x.w + kWTileW - 1) / kWTileW;
  p.th = kWTileH;
  if (x.h < 2 * kWTileH && !(taps == 9 && cout_pad == 64)) {
    // small images (22 or 11 rows at the bottom of the networks): a 16-row tile would be up to a third padding
    int best = (x.h + kWTileH - 1) / kWTileH * kWTileH;
    for (int th = kWTileH - 2; th >= 8; th -= 2) {
      const int cover = (x.h + th - 1) / th * th;
      if (cover * 100 < best * 85) {
        best = cover;
        p.th = th;
      }
    }
  }
  {
    // CVB_WGRAD_TH = 8 / 12: force a shorter pixel tile on the layers that would otherwise pipeline only two stages
    static int th_env = -1;
    if (th_env < 0) {
      const char* e = getenv("CVB_WGRAD_TH");
      th_env = e ? atoi(e) : 0;
    }
    if (th_env >= 8 && th_env < p.th && (th_env % 2) == 0 && !(taps == 9 && cout_pad == 64) && cout_pad >= 128 &&
        cin_pad >= 128)
      p.th = th_env;
  }
  p.tiles_h = (x.h + p.th - 1) / p.th;
  long long tiles = 1LL * p.tiles_w * p.tiles_h * x.n;
  CVB_REQUIRE(tiles < (1LL << 31), CVB_ERR_UNSUPPORTED, "conv_wgrad: too many pixel tiles");
  p.total_tiles = static_cast<int>(tiles);
  p.taps = taps;
  p.cin_pad = cin_pad;
  p.cout_pad = cout_pad;
  // CVB_WGRAD_BN=128: N = 128 tiles also where cout is a multiple of 256 (3 taps x 128 columns per item instead of 2 x 256:
  // fewer operand bytes per MAC, twice the co tiles) -- A/B knob
  static int bn_cap = -1;
  if (bn_cap < 0) {
    const char* e = getenv("CVB_WGRAD_BN");
    bn_cap = e ? atoi(e) : 256;
  }
  const int BN = (cout_pad % 256 == 0 && bn_cap >= 256) ? 256 : ((cout_pad % 128 == 0) ? 128 : 64);
  plan->BN = BN;
  // CVB_WGRAD_RS=0 keeps the N = 64 kernel for cout = 64 (A/B measurements); read per call
  const char* rs_env = getenv("CVB_WGRAD_RS");
  plan->rowshift = taps == 9 && cout_pad == 64 && !(rs_env && atoi(rs_env) == 0);
  CVB_REQUIRE(plan->rowshift || dy.c == cout_pad, CVB_ERR_UNSUPPORTED,
              "conv_wgrad: cout %d is not a multiple of 64 and the cout = 64 kernel does not apply", dy.c);
  if (plan->rowshift) {
    p.CM = 1;
    p.T = 9;
    p.n_tap_groups = 1;
    p.n_co_tiles = 1;
    p.n_ci_tiles = cin_pad / 64;
    p.stage_bytes = kRsStageBytes;
    p.stages = kRsStages;
    p.items = p.n_ci_tiles;
    p.total_units = 1LL * p.items * p.total_tiles;
    p.grid = static_cast<int>(p.total_units < sm_count() ? p.total_units : sm_count());
    int slots_rs = 1;
    for (int item = 0; item < p.items; ++item) {
      const long long u0 = 1LL * item * p.total_tiles;
      const int parts = sk_owner(p.total_units, p.grid, u0 + p.total_tiles - 1) - sk_owner(p.total_units, p.grid, u0) + 1;
      if (parts > slots_rs) slots_rs = parts;
    }
    p.slots = slots_rs;
    p.n_waves = 1;
    p.wave_item0[0] = 0; p.wave_items[0] = p.items; p.wave_grid[0] = p.grid;
    plan->grid = p.grid;
    plan->smem = 1024 + kRsStages * kRsStageBytes + 256;
    plan->ws_bytes = 1LL * slots_rs * taps * cin_pad * cout_pad * 4;
    return CVB_OK;
  }
  p.cluster = 1;
  p.CM = (cin_pad % 128 == 0) ? 2 : 1;
  // accumulators: accs * BN <= 512 TMEM columns, accs <= kMaxAccs; one accumulator = two 64-row atoms
  int max_accs = 512 / BN;
  if (max_accs > kMaxAccs) max_accs = kMaxAccs;
  const int max_taps = p.CM == 2 ? max_accs : 2 * max_accs;  // taps per item
  p.n_tap_groups = (taps + max_taps - 1) / max_taps;
  p.T = (taps + p.n_tap_groups - 1) / p.n_tap_groups;  // balanced groups
  p.n_tap_groups = (taps + p.T - 1) / p.T;
  p.n_co_tiles = cout_pad / BN;
  p.n_ci_tiles = cin_pad / (64 * p.CM);
  // Cluster of two along ci (CVB_WGRAD_CLUSTER = 1 or 2): neighbouring ci tiles of one (co tile, tap group) read the same dy
  // tiles. 1 = each CTA fetches half of them and multicasts (fewer L2 reads, same bytes delivered per SM: measured neutral);
  // 2 = the two CTAs run one M = 256 MMA as a pair and each keeps only HALF of the dy tile (fewer bytes delivered per SM,
  // the bound of this kernel: 42-45 B/clk per SM is what arrives when all SMs pull, the BN = 256 pipeline asks for 54).
  // Needs dy tiles of >= 2 boxes and an even number of ci tiles.
  int G = sm_count();
  p.pair = 0;
  {
    static int want = -1;
    if (want < 0) {
      const char* e = getenv("CVB_WGRAD_CLUSTER");
      want = e ? atoi(e) : 0;
    }
    if (want && BN >= 128 && (p.n_ci_tiles % 2) == 0 && (G % 2) == 0) {
      const int pairs = wgrad_max_clusters(BN, want == 2);
      if (pairs > 0) {
        p.cluster = 2;
        p.pair = want == 2 ? 1 : 0;
        p.n_ci_tiles /= 2;
        G = pairs < G / 2 ? pairs : G / 2;
      }
    }
  }
  // shared-memory slots of a stage follow the tile height (1 KB granularity: swizzle atoms): shorter tiles = smaller stages
  // = a deeper pipeline within the same 225 KB; a CTA of a pair holds half of the dy tile
  p.patch_stride = ((p.th + 2) * kWPitch * 128 + 1023) / 1024 * 1024;
  p.dy_bytes = p.th * kWTileW * 128;
  p.stage_bytes = p.CM * p.patch_stride + (BN / 64) / (p.pair ? 2 : 1) * p.dy_bytes;
  p.stages = (kWgradSmemBudget - 2048) / p.stage_bytes;
  if (p.stages > 6) p.stages = 6;
  CVB_REQUIRE(p.stages >= 2, CVB_ERR_UNSUPPORTED, "conv_wgrad: stage of %d bytes does not pipeline", p.stage_bytes);
  p.items = p.n_co_tiles * p.n_ci_tiles * p.n_tap_groups;
  p.total_units = 1LL * p.items * p.total_tiles;
  // Work partition. Stream-K: the (item, pixel tile) space is flattened and cut into equal contiguous ranges, one per
  // CTA. Lockstep: the grid of a partition is a MULTIPLE of its item count, so every item is cut at the same tile
  // boundaries, the CTAs of different items sweep the SAME pixel tiles at the same time and the x / dy tiles they share
  // are fetched from HBM once and hit in L2 by the others (with a plain 148-way cut the items drift apart and the big
  // layers re-read their operands 2.6-7.8x from HBM: ncu, 1.2 GB per launch against 354 MB on 256->128 at 180x240).
  // Waves: item counts that do not divide the SM count well (40, 80, 160 ...) are processed as successive lockstep
  // partitions, e.g. 80 items = 74 items x 2 CTAs, then 6 items x 24 CTAs; per CTA the waves add up to an equal share.
  // CVB_WGRAD_LOCKSTEP=0 falls back to one plain stream-K partition (A/B).
  const char* ls_env = getenv("CVB_WGRAD_LOCKSTEP");
  const bool lockstep = !(ls_env && atoi(ls_env) == 0);
  p.n_waves = 0;
  int remaining = p.items, off = 0;
  while (remaining > 0) {
    int take = remaining, grid_w;
    const long long units = 1LL * remaining * p.total_tiles;
    if (!lockstep || p.n_waves == kMaxWaves - 1) {
      grid_w = static_cast<int>(units < G ? units : G);  // plain stream-K over everything that is left
    } else if (remaining >= G) {
      take = G;  // one whole item per CTA: no split-K at all
      grid_w = G;
    } else {
      int S = G / remaining;
      if (remaining * S * 100 < G * 93) {  // would idle more than 7 % of the SMs: split the items instead
        S += 1;
        take = G / S;
      }
      if (S > p.total_tiles) S = p.total_tiles;
      grid_w = take * S;
    }
    p.wave_item0[p.n_waves] = off;
    p.wave_items[p.n_waves] = take;
    p.wave_grid[p.n_waves] = grid_w;
    ++p.n_waves;
    off += take;
    remaining -= take;
  }
  p.grid = 0;
  int slots = 1;
  for (int w = 0; w < p.n_waves; ++w) {
    if (p.wave_grid[w] > p.grid) p.grid = p.wave_grid[w];
    const long long wu = 1LL * p.wave_items[w] * p.total_tiles;
    for (int li = 0; li < p.wave_items[w]; ++li) {
      const long long u0 = 1LL * li * p.total_tiles;
      const int parts = sk_owner(wu, p.wave_grid[w], u0 + p.total_tiles - 1) - sk_owner(wu, p.wave_grid[w], u0) + 1;
      if (parts > slots) slots = parts;
    }
  }
  p.slots = slots;
  plan->grid = p.grid * p.cluster;
  plan->smem = 1024 + p.stages * p.stage_bytes + 256;
  plan->ws_bytes = 1LL * slots * taps * cin_pad * cout_pad * 4;
  return CVB_OK;
}

template <int BN, bool PAIR>
static int configure_wgrad() {
  static bool configured = false;
  if (!configured) {
    CVB_CUDA(cudaFuncSetAttribute(conv_wgrad_kernel<BN, PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  kWgradSmemBudget));
    configured = true;
  }
  return CVB_OK;
}

static void cluster_config(cudaLaunchConfig_t* cfg, cudaLaunchAttribute* at, int grid, int smem, cudaStream_t st) {
  memset(cfg, 0, sizeof(*cfg));
  cfg->gridDim = dim3(grid);
  cfg->blockDim = dim3(kWgradThreads);
  cfg->dynamicSmemBytes = smem;
  cfg->stream = st;
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg->attrs = at;
  cfg->numAttrs = 1;
}

// How many 2-CTA clusters of the weight-gradient kernel the device runs at once (1 CTA per SM: the two CTAs of a cluster
// need two free SMs of one GPC); 0 if clusters cannot be used. Asked once per instantiation.
template <int BN, bool PAIR>
static int max_clusters_bn() {
  static int cached = -1;
  if (cached < 0) {
    cached = 0;
    if (configure_wgrad<BN, PAIR>() == CVB_OK) {
      cudaLaunchConfig_t cfg;
      cudaLaunchAttribute at[1];
      cluster_config(&cfg, at, sm_count(), kWgradSmemBudget, nullptr);
      int n = 0;
      if (cudaOccupancyMaxActiveClusters(&n, conv_wgrad_kernel<BN, PAIR>, &cfg) == cudaSuccess) cached = n;
      else (void)cudaGetLastError();
    }
  }
  return cached;
}

static int wgrad_max_clusters(int BN, bool pair) {
  if (BN == 256) return pair ? max_clusters_bn<256, true>() : max_clusters_bn<256, false>();
  if (BN == 128) return pair ? max_clusters_bn<128, true>() : max_clusters_bn<128, false>();
  return 0;
}

template <int BN, bool PAIR>
static int launch_wgrad_as(const CUtensorMap& tmX, const CUtensorMap& tmDY, const WgradPlan& plan, cudaStream_t st) {
  int rc = configure_wgrad<BN, PAIR>();
  if (rc) return rc;
  if (plan.p.cluster == 2) {
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute at[1];
    cluster_config(&cfg, at, plan.grid, plan.smem, st);
    CVB_CUDA(cudaLaunchKernelEx(&cfg, conv_wgrad_kernel<BN, PAIR>, tmX, tmDY, plan.p));
  } else {
    conv_wgrad_kernel<BN, PAIR><<<plan.grid, kWgradThreads, plan.smem, st>>>(tmX, tmDY, plan.p);
  }
  CVB_LAUNCH_CHECK();
  return CVB_OK;
}

template <int BN>
static int launch_wgrad(const CUtensorMap& tmX, const CUtensorMap& tmDY, const WgradPlan& plan, cudaStream_t st) {
  if (BN >= 128 && plan.p.pair) return launch_wgrad_as<(BN >= 128 ? BN : 128), true>(tmX, tmDY, plan, st);
  return launch_wgrad_as<BN, false>(tmX, tmDY, plan, st);
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
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (plan.rowshift) {
    rc = make_act_tmap(&tmX, x, kWTileW + 2, kWTileH, 1);
    if (rc) return rc;
    rc = make_act_tmap(&tmDY, dy, kWTileW, kWTileH + 2, 1);
    if (rc) return rc;
    static bool configured = false;
    if (!configured) {
      CVB_CUDA(cudaFuncSetAttribute(conv_wgrad_rs64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    kWgradSmemBudget));
      configured = true;
    }
    conv_wgrad_rs64_kernel<<<plan.grid, kWgradThreads, plan.smem, st>>>(tmX, tmDY, plan.p);
    CVB_LAUNCH_CHECK();
  } else {
  rc = make_act_tmap(&tmX, x, kWTileW + 2, plan.p.th + 2, 1);
  if (rc) return rc;
  rc = make_act_tmap(&tmDY, dy, kWTileW, plan.p.th, 1);
  if (rc) return rc;
  switch (plan.BN) {
    case 256: rc = launch_wgrad<256>(tmX, tmDY, plan, st); break;
    case 128: rc = launch_wgrad<128>(tmX, tmDY, plan, st); break;
    default: rc = launch_wgrad<64>(tmX, tmDY, plan, st); break;
  }
  if (rc) return rc;
  }
  // reduction bricks: part groups for items cut into many parts (ps_max * bci <= 32 bounds the brick buffer), then shrink
  // the ci extent until the grid has a few hundred blocks
  const int cin_pad = plan.p.cin_pad, cout_pad = plan.p.cout_pad;
  const int ps_max = plan.p.slots >= 32 ? 8 : (plan.p.slots >= 8 ? 4 : (plan.p.slots >= 3 ? 2 : 1));
  int bci = 32 / ps_max;
  while (bci > 1 && 1LL * ((cout_pad + 31) / 32) * ((cin_pad + bci - 1) / bci) < 2 * sm_count()) bci >>= 1;
  dim3 rgrid((cout_pad + 31) / 32, (x.c + bci - 1) / bci);
  const size_t rsmem = static_cast<size_t>(ps_max) * taps * bci * 33 * sizeof(float);
  if (taps == 9)
    wgrad_reduce_kernel<9><<<rgrid, 256, rsmem, st>>>(plan.p.ws, plan.p, plan.BN, cout, cin_eff, x.c, bci, ps_max, dw);
  else
    wgrad_reduce_kernel<1><<<rgrid, 256, rsmem, st>>>(plan.p.ws, plan.p, plan.BN, cout, cin_eff, x.c, bci, ps_max, dw);
  CVB_LAUNCH_CHECK();
  return CVB_OK;
}
