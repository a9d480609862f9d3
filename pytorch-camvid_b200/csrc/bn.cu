// BatchNorm2d(+ReLU) training-mode kernels on NHWC bf16 views (reference: models/unet.py:12-13, models/segnet.py:9-10).
// All HBM-bound: 128-bit accesses, one thread = 8 channels of one pixel, fp32 math.
#include "common.cuh"

namespace cvb {

constexpr int kThreads = 256;

// ---------------------------------------------------------------------------------------------------------------
// Per-channel reductions: every block strides over pixels, each thread owns one 8-channel group, partial sums are
// combined in shared memory and written as one row of partials[row][2][C].
//   MODE 0: (sum y, sum y^2)                         -- batch statistics
//   MODE 1: g = da*[y*scale+shift > 0]; (sum g, sum g*y) -- BatchNorm+ReLU backward reduction
// ---------------------------------------------------------------------------------------------------------------
// reverse != 0: pixels are visited from the last to the first (an experiment in L2 reuse -- the tensors these passes read
// were just written front to back and only their END can still be cached; measured neutral). The
// partial sums are the same numbers either way up to fp32 summation order.
template <int MODE>
__global__ void __launch_bounds__(kThreads) bn_reduce_kernel(View y, View da, const float* __restrict__ scale,
                                                              const float* __restrict__ shift,
                                                              float* __restrict__ partials, int reverse) {
  __shared__ float red[kThreads * 2];
  const int CV = y.c >> 3;
  const int ppb = kThreads / CV;  // pixels handled per block iteration (CV divides kThreads, checked on host)
  const int cv = threadIdx.x % CV;
  const int pl = threadIdx.x / CV;
  const unsigned npix = static_cast<unsigned>(y.n) * y.h * y.w;

  float s1[8], s2[8], sc[8], sh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s1[j] = s2[j] = 0.f;
  if (MODE == 1) {
    ld8f(scale + cv * 8, sc);
    ld8f(shift + cv * 8, sh);
  }
  const unsigned step = gridDim.x * ppb;
  for (unsigned it = blockIdx.x * ppb + pl; it < npix; it += 2 * step) {
    const bool has2 = it + step < npix;
    const unsigned pix = reverse ? npix - 1u - it : it;
    const unsigned pix2 = has2 ? (reverse ? npix - 1u - (it + step) : it + step) : pix;
    // all loads of both pixels are issued before the arithmetic
    const uint4 uy = ldg16(y.p + poff(y, pix) + cv * 8);
    uint4 uy2 = make_uint4(0, 0, 0, 0), ud = uy2, ud2 = uy2;
    if (has2) uy2 = ldg16(y.p + poff(y, pix2) + cv * 8);
    if (MODE == 1) {
      ud = ldg16(da.p + poff(da, pix) + cv * 8);
      if (has2) ud2 = ldg16(da.p + poff(da, pix2) + cv * 8);
    }
#pragma unroll
    for (int rep = 0; rep < 2; ++rep) {
      if (rep == 1 && !has2) break;
      float fy[8], fd[8];
      unpack8(rep == 0 ? uy : uy2, fy);
      if (MODE == 0) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          s1[j] += fy[j];
          s2[j] = fmaf(fy[j], fy[j], s2[j]);
        }
      } else {
        unpack8(rep == 0 ? ud : ud2, fd);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float g = fmaf(fy[j], sc[j], sh[j]) > 0.f ? fd[j] : 0.f;
          s1[j] += g;
          s2[j] = fmaf(g, fy[j], s2[j]);
        }
      }
    }
  }
  // Cross-thread combine in eight rounds through a 2 KB buffer (round j = channel j of every 8-channel group): the
  // kernel must stay co-resident with the weight-gradient kernel of the previous block, which leaves < 5 KB of an SM's
  // shared memory free (engine.py runs it on a side stream under these HBM-bound passes).
  const int C = y.c;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    red[threadIdx.x * 2] = s1[j];
    red[threadIdx.x * 2 + 1] = s2[j];
    __syncthreads();
    if (threadIdx.x < 2 * CV) {  // thread (which, group) sums its column over the ppb pixel lanes
      const int which = threadIdx.x / CV, ocv = threadIdx.x - which * CV;
      float acc = 0.f;
      for (int l = 0; l < ppb; ++l) acc += red[(l * CV + ocv) * 2 + which];
      partials[(1LL * blockIdx.x * 2 + which) * C + ocv * 8 + j] = acc;
    }
    __syncthreads();
  }
}

static int reduce_launch_cfg(const cvb_view& v, int rows, int* grid) {
  int CV = v.c / 8;
  CVB_REQUIRE(CV <= kThreads && (kThreads % CV) == 0, CVB_ERR_UNSUPPORTED,
              "bn reduce: channels %d must divide %d", v.c, kThreads * 8);
  CVB_REQUIRE(rows > 0, CVB_ERR_INVALID_ARG, "bn reduce: rows must be positive");
  CVB_REQUIRE(fits_u32(v), CVB_ERR_UNSUPPORTED, "bn reduce: view too large for 32-bit indexing");
  *grid = rows;
  return CVB_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// finalize kernels: tiny and latency-bound, and they sit on the critical path twice per block (forward statistics,
// backward coefficients). A block owns 8 channels; a warp load covers 4 partial rows x 8 channels (four 32-byte
// sectors), 16 warps stride over the rows with 4 loads of each quantity in flight, so the dependent-load chain is
// rows / 256 deep instead of rows / 64. Combine: shuffles over the 4 row lanes, then 2 KB of shared memory (the kernels
// must fit next to a resident weight-gradient CTA, see bn_reduce_kernel). Lanes 0-7 of warp 0 finish the 8 channels.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kFinThreads = 512;
constexpr int kFinWarps = kFinThreads / 32;
constexpr int kFinCh = 8;  // channels per block

__device__ __forceinline__ void reduce_partial_rows(const float* __restrict__ partials, int rows, int pstride, int ch,
                                                    bool ch_ok, double* s1_out, double* s2_out) {
  __shared__ double sh[2][kFinWarps][kFinCh];
  const int lane = threadIdx.x & 31, wq = threadIdx.x >> 5;
  const int rsub = lane >> 3;  // which of the 4 rows of a warp load
  double s1 = 0.0, s2 = 0.0;
  if (ch_ok) {
    constexpr int RS = kFinWarps * 4;  // rows covered by one load of every warp
    int r = wq * 4 + rsub;
    for (; r + 3 * RS < rows; r += 4 * RS) {
      float a0 = partials[(2LL * r) * pstride + ch], b0 = partials[(2LL * r + 1) * pstride + ch];
      float a1 = partials[(2LL * (r + RS)) * pstride + ch], b1 = partials[(2LL * (r + RS) + 1) * pstride + ch];
      float a2 = partials[(2LL * (r + 2 * RS)) * pstride + ch], b2 = partials[(2LL * (r + 2 * RS) + 1) * pstride + ch];
      float a3 = partials[(2LL * (r + 3 * RS)) * pstride + ch], b3 = partials[(2LL * (r + 3 * RS) + 1) * pstride + ch];
      s1 += (static_cast<double>(a0) + a1) + (static_cast<double>(a2) + a3);
      s2 += (static_cast<double>(b0) + b1) + (static_cast<double>(b2) + b3);
    }
    for (; r < rows; r += RS) {
      s1 += static_cast<double>(partials[(2LL * r) * pstride + ch]);
      s2 += static_cast<double>(partials[(2LL * r + 1) * pstride + ch]);
    }
  }
  s1 += __shfl_xor_sync(0xffffffffu, s1, 8);
  s2 += __shfl_xor_sync(0xffffffffu, s2, 8);
  s1 += __shfl_xor_sync(0xffffffffu, s1, 16);
  s2 += __shfl_xor_sync(0xffffffffu, s2, 16);
  if (lane < kFinCh) {
    sh[0][wq][lane] = s1;
    sh[1][wq][lane] = s2;
  }
  __syncthreads();
  if (threadIdx.x < kFinCh) {
    s1 = s2 = 0.0;
#pragma unroll
    for (int q = 0; q < kFinWarps; ++q) {
      s1 += sh[0][q][lane];
      s2 += sh[1][q][lane];
    }
  }
  *s1_out = s1;
  *s2_out = s2;
}

__global__ void __launch_bounds__(kFinThreads)
bn_finalize_kernel(const float* __restrict__ partials, int rows, int c, int c_pad, int pstride, double inv_count,
                   double unbias, const float* __restrict__ gamma, const float* __restrict__ beta,
                   const float* __restrict__ conv_bias, float* running_mean, float* running_var, float momentum,
                   float eps, float* mean, float* invstd, float* scale, float* shift) {
  const int ch = blockIdx.x * kFinCh + (threadIdx.x & (kFinCh - 1));
  double s1, s2;
  reduce_partial_rows(partials, rows, pstride, ch, ch < c, &s1, &s2);
  if (threadIdx.x >= kFinCh || ch >= c_pad) return;
  if (ch >= c) {
    scale[ch] = 0.f;
    shift[ch] = 0.f;
    if (mean) mean[ch] = 0.f;
    if (invstd) invstd[ch] = 0.f;
    return;
  }
  double m = s1 * inv_count;
  double var = s2 * inv_count - m * m;
  if (var < 0.0) var = 0.0;
  float is = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
  float sc = gamma[ch] * is;
  mean[ch] = static_cast<float>(m);
  invstd[ch] = is;
  scale[ch] = sc;
  shift[ch] = beta[ch] - static_cast<float>(m) * sc;
  if (running_mean) {
    float mb = static_cast<float>(m) + (conv_bias ? conv_bias[ch] : 0.f);
    running_mean[ch] = (1.f - momentum) * running_mean[ch] + momentum * mb;
    running_var[ch] = (1.f - momentum) * running_var[ch] + momentum * static_cast<float>(var * unbias);
  }
}

__global__ void __launch_bounds__(kFinThreads)
bn_bwd_finalize_kernel(const float* __restrict__ partials, int rows, int c, int c_pad, int pstride, double inv_count,
                       const float* __restrict__ gamma, const float* __restrict__ mean,
                       const float* __restrict__ invstd, float* dgamma, float* dbeta, float* coef) {
  const int ch = blockIdx.x * kFinCh + (threadIdx.x & (kFinCh - 1));
  double sg, sgy;
  reduce_partial_rows(partials, rows, pstride, ch, ch < c, &sg, &sgy);
  if (threadIdx.x >= kFinCh || ch >= c_pad) return;
  if (ch >= c) {
    coef[ch] = 0.f;
    coef[c_pad + ch] = 0.f;
    coef[2 * c_pad + ch] = 0.f;
    return;
  }
  double m = mean[ch], is = invstd[ch], g = gamma[ch];
  double dg = is * (sgy - m * sg);  // sum g * xhat
  double db = sg;
  dgamma[ch] = static_cast<float>(dg);
  dbeta[ch] = static_cast<float>(db);
  double sc = g * is;
  // dy = sc*g - sc*db/N - sc*(y-m)*is*dg/N
  coef[ch] = static_cast<float>(sc);
  coef[c_pad + ch] = static_cast<float>(-sc * is * dg * inv_count);
  coef[2 * c_pad + ch] = static_cast<float>(-sc * db * inv_count + sc * m * is * dg * inv_count);
}

// ---------------------------------------------------------------------------------------------------------------
// apply kernels
// ---------------------------------------------------------------------------------------------------------------
// Elementwise kernels: a thread's 8-channel group is the same in every iteration whenever the grid stride is a
// multiple of C/8 (always for the power-of-two widths of these networks), so the per-channel vectors are loaded once
// into registers (HOIST); otherwise they are re-read from L1 per element.
template <bool HOIST>
__global__ void __launch_bounds__(kThreads) bn_relu_apply_kernel(View y, View a, const float* __restrict__ scale,
                                                                  const float* __restrict__ shift, int reverse) {
  const unsigned total = static_cast<unsigned>(y.n) * y.h * y.w * (y.c >> 3);
  const unsigned stride = gridDim.x * kThreads;
  float sc[8], sh[8];
  if (HOIST) {
    unsigned pix0, cv0;
    const unsigned i0 = blockIdx.x * kThreads + threadIdx.x;
    split_cv(y, reverse ? total - 1u - (i0 < total ? i0 : 0u) : i0, pix0, cv0);
    ld8f(scale + cv0 * 8, sc);
    ld8f(shift + cv0 * 8, sh);
  }
  // four independent elements per iteration: all loads are issued before the first store
  for (unsigned i = blockIdx.x * kThreads + threadIdx.x; i < total; i += 4 * stride) {
    uint4 u[4];
    unsigned pix[4], cv[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const unsigned iq = i + q * stride;
      pix[q] = 0xffffffffu;
      if (iq < total) {
        split_cv(y, reverse ? total - 1u - iq : iq, pix[q], cv[q]);
        u[q] = ldg16(y.p + poff(y, pix[q]) + cv[q] * 8);
      }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (pix[q] == 0xffffffffu) continue;
      float f[8];
      unpack8(u[q], f);
      if (!HOIST) {
        ld8f(scale + cv[q] * 8, sc);
        ld8f(shift + cv[q] * 8, sh);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = fmaxf(fmaf(f[j], sc[j], sh[j]), 0.f);
      stg16(a.p + poff(a, pix[q]) + cv[q] * 8, pack8(f));
    }
  }
}

template <bool HOIST>
__global__ void __launch_bounds__(kThreads) bn_relu_bwd_apply_kernel(View da, View y, View dy,
                                                                      const float* __restrict__ scale,
                                                                      const float* __restrict__ shift,
                                                                      const float* __restrict__ coef, int c_pad,
                                                                      int reverse) {
  const unsigned total = static_cast<unsigned>(y.n) * y.h * y.w * (y.c >> 3);
  const unsigned stride = gridDim.x * kThreads;
  float sc[8], sh[8], c0[8], c1[8], c2[8];
  if (HOIST) {
    unsigned pix0, cv0;
    const unsigned i0 = blockIdx.x * kThreads + threadIdx.x;
    split_cv(y, reverse ? total - 1u - (i0 < total ? i0 : 0u) : i0, pix0, cv0);
    ld8f(scale + cv0 * 8, sc);
    ld8f(shift + cv0 * 8, sh);
    ld8f(coef + cv0 * 8, c0);
    ld8f(coef + c_pad + cv0 * 8, c1);
    ld8f(coef + 2 * c_pad + cv0 * 8, c2);
  }
  for (unsigned i = blockIdx.x * kThreads + threadIdx.x; i < total; i += 2 * stride) {
    uint4 uy[2], ud[2];
    unsigned pix[2], cv[2];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const unsigned iq = i + q * stride;
      pix[q] = 0xffffffffu;
      if (iq < total) {
        split_cv(y, reverse ? total - 1u - iq : iq, pix[q], cv[q]);
        uy[q] = ldg16(y.p + poff(y, pix[q]) + cv[q] * 8);
        ud[q] = ldg16(da.p + poff(da, pix[q]) + cv[q] * 8);
      }
    }
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      if (pix[q] == 0xffffffffu) continue;
      if (!HOIST) {
        ld8f(scale + cv[q] * 8, sc);
        ld8f(shift + cv[q] * 8, sh);
        ld8f(coef + cv[q] * 8, c0);
        ld8f(coef + c_pad + cv[q] * 8, c1);
        ld8f(coef + 2 * c_pad + cv[q] * 8, c2);
      }
      float fy[8], fd[8], o[8];
      unpack8(uy[q], fy);
      unpack8(ud[q], fd);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float g = fmaf(fy[j], sc[j], sh[j]) > 0.f ? fd[j] : 0.f;
        o[j] = fmaf(g, c0[j], fmaf(fy[j], c1[j], c2[j]));
      }
      stg16(dy.p + poff(dy, pix[q]) + cv[q] * 8, pack8(o));
    }
  }
}


// ---------------------------------------------------------------------------------------------------------------
// The network's last block, fused with the module boundary (cross-layer fusion): its
// activation IS the fp32 NCHW logits tensor the module returns (models/unet.py:156, models/segnet.py:119), and the
// gradient that enters its backward pass IS the fp32 NCHW dlogits of the loss.
//   forward : logits[n,c,h,w] = float(bf16(relu(y*scale+shift)))  -- cvb_bn_relu_apply + cvb_nhwc_bf16_to_nchw_f32 in one
//             pass; the bf16 activation is never materialised (nothing reads it: the backward mask comes from y)
//   backward: da = bf16(dlogits) as NHWC + (sum g, sum g*y)       -- cvb_nchw_f32_to_nhwc_bf16 + cvb_bn_relu_bwd_reduce
// Thread = one pixel (consecutive threads = consecutive pixels: every plane access of a warp is one 128-byte line, the
// NHWC side is 32 B x CV per thread, contiguous across the warp). CV = 8-channel vectors of the narrow class tensor.
// ---------------------------------------------------------------------------------------------------------------
template <int CV>
__global__ void __launch_bounds__(kThreads) bn_relu_apply_nchw_kernel(View y, const float* __restrict__ scale,
                                                                       const float* __restrict__ shift,
                                                                       float* __restrict__ dst, int c_dst) {
  const unsigned hw = static_cast<unsigned>(y.h) * y.w;
  const unsigned total = static_cast<unsigned>(y.n) * hw;
  float sc[CV][8], sh[CV][8];
#pragma unroll
  for (int v = 0; v < CV; ++v) {
    ld8f(scale + v * 8, sc[v]);
    ld8f(shift + v * 8, sh[v]);
  }
  for (unsigned i = blockIdx.x * kThreads + threadIdx.x; i < total; i += gridDim.x * kThreads) {
    const unsigned n = i / hw, p = i - n * hw;
    const __nv_bfloat16* sp = y.p + poff(y, i);
    uint4 u[CV];
#pragma unroll
    for (int v = 0; v < CV; ++v) u[v] = ldg16(sp + v * 8);
    float* dp = dst + static_cast<long long>(n) * c_dst * hw + p;
#pragma unroll
    for (int v = 0; v < CV; ++v) {
      float f[8];
      unpack8(u[v], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = fmaxf(fmaf(f[j], sc[v][j], sh[v][j]), 0.f);
      unpack8(pack8(f), f);  // the value the unfused pair would have stored as bf16
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (v * 8 + j < c_dst) dp[static_cast<long long>(v * 8 + j) * hw] = f[j];
    }
  }
}

template <int CV>
__global__ void __launch_bounds__(kThreads) nchw_to_nhwc_bn_reduce_kernel(const float* __restrict__ src, int c_src,
                                                                           View da, View y,
                                                                           const float* __restrict__ scale,
                                                                           const float* __restrict__ shift,
                                                                           float* __restrict__ partials) {
  __shared__ float red[kThreads / 32][2 * CV * 8];
  const unsigned hw = static_cast<unsigned>(y.h) * y.w;
  const unsigned total = static_cast<unsigned>(y.n) * hw;
  float sc[CV][8], sh[CV][8], s1[CV][8], s2[CV][8];
#pragma unroll
  for (int v = 0; v < CV; ++v) {
    ld8f(scale + v * 8, sc[v]);
    ld8f(shift + v * 8, sh[v]);
#pragma unroll
    for (int j = 0; j < 8; ++j) s1[v][j] = s2[v][j] = 0.f;
  }
  for (unsigned i = blockIdx.x * kThreads + threadIdx.x; i < total; i += gridDim.x * kThreads) {
    const unsigned n = i / hw, p = i - n * hw;
    const float* sp = src + static_cast<long long>(n) * c_src * hw + p;
    uint4 uy[CV];
    float g[CV][8];
#pragma unroll
    for (int v = 0; v < CV; ++v) uy[v] = ldg16(y.p + poff(y, i) + v * 8);
#pragma unroll
    for (int v = 0; v < CV; ++v)
#pragma unroll
      for (int j = 0; j < 8; ++j) g[v][j] = (v * 8 + j < c_src) ? __ldcs(sp + static_cast<long long>(v * 8 + j) * hw) : 0.f;
    __nv_bfloat16* dp = da.p + poff(da, i);
#pragma unroll
    for (int v = 0; v < CV; ++v) {
      const uint4 pk = pack8(g[v]);
      stg16(dp + v * 8, pk);
      float fy[8], fd[8];
      unpack8(pk, fd);  // sums of the bf16 values the apply pass will read
      unpack8(uy[v], fy);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float ge = fmaf(fy[j], sc[v][j], sh[v][j]) > 0.f ? fd[j] : 0.f;
        s1[v][j] += ge;
        s2[v][j] = fmaf(ge, fy[j], s2[v][j]);
      }
    }
  }
  // every thread holds all channels: fold the 32 lanes by shuffles, then the 8 warps through 1 KB of shared memory
  const int lane = threadIdx.x & 31, wq = threadIdx.x >> 5;
#pragma unroll
  for (int v = 0; v < CV; ++v)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float a = s1[v][j], b = s2[v][j];
#pragma unroll
      for (int m = 16; m > 0; m >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, m);
        b += __shfl_xor_sync(0xffffffffu, b, m);
      }
      if (lane == 0) {
        red[wq][v * 8 + j] = a;
        red[wq][CV * 8 + v * 8 + j] = b;
      }
    }
  __syncthreads();
  if (threadIdx.x < 2 * CV * 8) {
    float acc = 0.f;
#pragma unroll
    for (int q = 0; q < kThreads / 32; ++q) acc += red[q][threadIdx.x];
    const int which = threadIdx.x / (CV * 8), ch = threadIdx.x - which * (CV * 8);
    partials[(1LL * blockIdx.x * 2 + which) * (CV * 8) + ch] = acc;
  }
}

}  // namespace cvb

using namespace cvb;

extern "C" int cvb_bn_stats(cvb_view y, float* partials, int rows, void* stream) {
  int rc = check_view(y, "bn_stats.y");
  if (rc) return rc;
  CVB_REQUIRE(partials, CVB_ERR_INVALID_ARG, "bn_stats: null partials");
  int grid;
  rc = reduce_launch_cfg(y, rows, &grid);
  if (rc) return rc;
  View vy = to_dev(y);
  bn_reduce_kernel<0><<<grid, kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      vy, vy, nullptr, nullptr, partials, 0);
  CVB_LAUNCH_CHECK();
  return CVB_OK;
}

extern "C" int cvb_bn_relu_bwd_reduce(cvb_view da, cvb_view y, const float* scale, const float* shift,
                                      float* partials, int rows, int reverse, void* stream) {
  int rc = check_view(y, "bn_bwd_reduce.y");
  if (rc) return rc;
  rc = check_view(da, "bn_bwd_reduce.da");
  if (rc) return rc;
  CVB_REQUIRE(same_shape(da, y), CVB_ERR_INVALID_ARG, "bn_bwd_reduce: da and y shapes differ");
  CVB_REQUIRE(scale && shift && partials, CVB_ERR_INVALID_ARG, "bn_bwd_reduce: null pointer");
  int grid;
  rc = reduce_launch_cfg(y, rows, &grid);
  if (rc) return rc;
  bn_reduce_kernel<1><<<grid, kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      to_dev(y), to_dev(da), scale, shift, partials, reverse);
  CVB_LAUNCH_CHECK();
  return CVB_OK;
}

extern "C" int cvb_bn_finalize(const float* partials, int rows, int c, int c_pad, int64_t count, const float* gamma,
                               const float* beta, const float* conv_bias, float* running_mean, float* running_var,
                               float momentum, float eps, float* mean, float* invstd, float* scale, float* shift,
                               void* stream) {
  CVB_REQUIRE(partials && gamma && beta && mean && invstd && scale && shift, CVB_ERR_INVALID_ARG,
              "bn_finalize: null pointer");
  CVB_REQUIRE(rows > 0 && c > 0 && c_pad >= c && count > 0, CVB_ERR_INVALID_ARG, "bn_finalize: bad sizes");
  CVB_REQUIRE((running_mean == nullptr) == (running_var == nullptr), CVB_ERR_INVALID_ARG,
              "bn_finalize: running_mean and running_var must both be given or both be NULL");
  double inv = 1.0 / static_cast<double>(count);
  double unbias = count > 1 ? static_cast<double>(count) / static_cast<double>(count - 1) : 1.0;
  bn_finalize_kernel<<<(c_pad + kFinCh - 1) / kFinCh, kFinThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      partials, rows, c, c_pad, c_pad, inv, unbias, gamma, beta, conv_bias, running_mean, running_var, momentum, eps,
      mean, invstd, scale, shift);
  CVB_LAUNCH_CHECK();
  return CVB_OK;
}

extern "C" int cvb_bn_bwd_finalize(const float* partials, int rows, int c, int c_pad, int64_t count,
                                   const float* gamma, const float* mean, const float* invstd, float* dgamma,
                                   float* dbeta, float* coef, void* stream) {
  CVB_REQUIRE(partials && gamma && mean && invstd && dgamma && dbeta && coef, CVB_ERR_INVALID_ARG,
              "bn_bwd_finalize: null pointer");
  CVB_REQUIRE(rows > 0 && c > 0 && c_pad >= c && count > 0, CVB_ERR_INVALID_ARG, "bn_bwd_finalize: bad sizes");
  bn_bwd_finalize_kernel<<<(c_pad + kFinCh - 1) / kFinCh, kFinThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      partials, rows, c, c_pad, c_pad, 1.0 / static_cast<double>(count), gamma, mean, invstd, dgamma, dbeta, coef);
  CVB_LAUNCH_CHECK();
  return CVB_OK;
}

extern "C" int cvb_bn_relu_apply(cvb_view y, const float* scale, const float* shift, cvb_view a, int reverse,
                                 void* stream) {
  int rc = check_view(y, "bn_relu_apply.y");
  if (rc) return rc;
  rc = check_view(a, "bn_relu_apply.a");
  if (rc) return rc;
  CVB_REQUIRE(same_shape(y, a), CVB_ERR_INVALID_ARG, "bn_relu_apply: shapes differ");
  CVB_REQUIRE(fits_u32(y), CVB_ERR_UNSUPPORTED, "bn_relu_apply: view too large for 32-bit indexing");
  CVB_REQUIRE(scale && shift, CVB_ERR_INVALID_ARG, "bn_relu_apply: null scale/shift");
  long long total = 1LL * y.n * y.h * y.w * (y.c / 8);
  const int grid = ew_grid((total + 3) / 4, kThreads);
  if (kThreads % (y.c / 8) == 0)
    bn_relu_apply_kernel<true><<<grid, kThreads, 0, static_cast<cudaStream_t>(stream)>>>(to_dev(y), to_dev(a), scale, shift,
                                                                                          reverse);
  else
    bn_relu_apply_kernel<false><<<grid, kThreads, 0, static_cast<cudaStream_t>(stream)>>>(to_dev(y), to_dev(a), scale, shift,
                                                                                           reverse);
  CVB_LAUNCH_CHECK();
  return CVB_OK;
}

extern "C" int cvb_bn_relu_bwd_apply(cvb_view da, cvb_view y, const float* scale, const float* shift,
                                     const float* coef, cvb_view dy, int reverse, void* stream) {
  int rc = check_view(y, "bn_bwd_apply.y");
  if (rc) return rc;
  rc = check_view(da, "bn_bwd_apply.da");
  if (rc) return rc;
  rc = check_view(dy, "bn_bwd_apply.dy");
  if (rc) return rc;
  CVB_REQUIRE(same_shape(y, da) && same_shape(y, dy), CVB_ERR_INVALID_ARG, "bn_bwd_apply: shapes differ");
  CVB_REQUIRE(fits_u32(y), CVB_ERR_UNSUPPORTED, "bn_bwd_apply: view too large for 32-bit indexing");
  CVB_REQUIRE(scale && shift && coef, CVB_ERR_INVALID_ARG, "bn_bwd_apply: null pointer");
  long long total = 1LL * y.n * y.h * y.w * (y.c / 8);
  const int grid = ew_grid((total + 1) / 2, kThreads);
  if (kThreads % (y.c / 8) == 0)
    bn_relu_bwd_apply_kernel<true><<<grid, kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        to_dev(da), to_dev(y), to_dev(dy), scale, shift, coef, y.c, reverse);
  else
    bn_relu_bwd_apply_kernel<false><<<grid, kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        to_dev(da), to_dev(y), to_dev(dy), scale, shift, coef, y.c, reverse);
  CVB_LAUNCH_CHECK();
  return CVB_OK;
}

extern "C" int cvb_bn_relu_apply_nchw_f32(cvb_view y, const float* scale, const float* shift, float* dst, int c_dst,
                                          void* stream) {
  int rc = check_view(y, "bn_relu_apply_nchw.y");
  if (rc) return rc;
  CVB_REQUIRE(scale && shift && dst, CVB_ERR_INVALID_ARG, "bn_relu_apply_nchw: null pointer");
  CVB_REQUIRE(c_dst > 0 && c_dst <= y.c, CVB_ERR_INVALID_ARG, "bn_relu_apply_nchw: c_dst=%d does not fit the %d channels of y",
              c_dst, y.c);
  CVB_REQUIRE(y.c == 8 || y.c == 16, CVB_ERR_UNSUPPORTED,
              "bn_relu_apply_nchw: the fused boundary kernel serves class tensors of 8 or 16 channels in memory (got %d); "
              "use cvb_bn_relu_apply + cvb_nhwc_bf16_to_nchw_f32", y.c);
  CVB_REQUIRE(1LL * y.n * y.h * y.w < (1LL << 31), CVB_ERR_UNSUPPORTED, "bn_relu_apply_nchw: view too large for 32-bit indexing");
  const long long total = 1LL * y.n * y.h * y.w;
  const int grid = ew_grid(total, kThreads);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (y.c == 16)
    bn_relu_apply_nchw_kernel<2><<<grid, kThreads, 0, st>>>(to_dev(y), scale, shift, dst, c_dst);
  else
    bn_relu_apply_nchw_kernel<1><<<grid, kThreads, 0, st>>>(to_dev(y), scale, shift, dst, c_dst);
  CVB_LAUNCH_CHECK();
  return CVB_OK;
}

extern "C" int cvb_nchw_f32_to_nhwc_bf16_bn_reduce(const float* src, int c_src, cvb_view da, cvb_view y,
                                                   const float* scale, const float* shift, float* partials, int rows,
                                                   void* stream) {
  int rc = check_view(y, "nchw_to_nhwc_bn_reduce.y");
  if (rc) return rc;
  rc = check_view(da, "nchw_to_nhwc_bn_reduce.da");
  if (rc) return rc;
  CVB_REQUIRE(same_shape(da, y), CVB_ERR_INVALID_ARG, "nchw_to_nhwc_bn_reduce: da and y shapes differ");
  CVB_REQUIRE(src && scale && shift && partials && rows > 0, CVB_ERR_INVALID_ARG, "nchw_to_nhwc_bn_reduce: null pointer / rows");
  CVB_REQUIRE(c_src > 0 && c_src <= y.c, CVB_ERR_INVALID_ARG, "nchw_to_nhwc_bn_reduce: c_src=%d does not fit the %d channels",
              c_src, y.c);
  CVB_REQUIRE(y.c == 8 || y.c == 16, CVB_ERR_UNSUPPORTED,
              "nchw_to_nhwc_bn_reduce: the fused boundary kernel serves class tensors of 8 or 16 channels in memory (got %d); "
              "use cvb_nchw_f32_to_nhwc_bf16 + cvb_bn_relu_bwd_reduce", y.c);
  CVB_REQUIRE(1LL * y.n * y.h * y.w < (1LL << 31), CVB_ERR_UNSUPPORTED, "nchw_to_nhwc_bn_reduce: view too large for 32-bit indexing");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (y.c == 16)
    nchw_to_nhwc_bn_reduce_kernel<2><<<rows, kThreads, 0, st>>>(src, c_src, to_dev(da), to_dev(y), scale, shift, partials);
  else
    nchw_to_nhwc_bn_reduce_kernel<1><<<rows, kThreads, 0, st>>>(src, c_src, to_dev(da), to_dev(y), scale, shift, partials);
  CVB_LAUNCH_CHECK();
  return CVB_OK;
}
