// nn.Upsample(scale_factor=2, mode='bilinear', align_corners=True) forward / backward on NHWC bf16 views
// (reference: models/unet.py:25,29). Source index math follows ATen's upsample_bilinear2d (fp32):
//   scale = (in-1)/(out-1);  src = scale*dst;  i0 = (int)src;  i1 = i0 + (i0 < in-1);  l1 = src - i0;  l0 = 1 - l1.
#include "common.cuh"

namespace cvb {

constexpr int kThreads = 256;

__device__ __forceinline__ void src_index(float scale, int dst, int in, int& i0, int& i1, float& l0, float& l1) {
  float src = scale * static_cast<float>(dst);
  i0 = static_cast<int>(src);
  if (i0 > in - 1) i0 = in - 1;
  i1 = i0 + ((i0 < in - 1) ? 1 : 0);
  l1 = src - static_cast<float>(i0);
  l0 = 1.f - l1;
}

// x-interpolated row: lx0 * in[y][x0] + lx1 * in[y][x1] for 8 channels. BN: the source is a raw conv output y and the
// value interpolated is a = bf16(relu(y*scale + shift)), exactly what cvb_bn_relu_apply would have stored.
template <bool BN>
__device__ __forceinline__ void lerp_row(const __nv_bfloat16* p0, const __nv_bfloat16* p1, float lx0, float lx1,
                                         const float (&sc)[8], const float (&sh)[8], float (&r)[8]) {
  const uint4 ua = ldg16(p0), ub = ldg16(p1);
  float a[8], b[8];
  unpack8(ua, a);
  unpack8(ub, b);
  if (BN) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      a[j] = fmaxf(fmaf(a[j], sc[j], sh[j]), 0.f);
      b[j] = fmaxf(fmaf(b[j], sc[j], sh[j]), 0.f);
    }
    unpack8(pack8(a), a);
    unpack8(pack8(b), b);
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) r[j] = lx0 * a[j] + lx1 * b[j];
}

// One thread = one 8-channel vector of one output COLUMN, marching down a strip of `rows` output rows. The
// x-interpolated values of the two source rows in use stay in registers: a source row is fetched once per strip (two
// 16-byte loads) instead of once per output pixel that touches it -- about one load per store instead of four.
// Consecutive threads = consecutive channel vectors, then consecutive columns: every access of a warp is one
// contiguous run. The kernel is bound by HBM WRITES (8 of its 10 bytes per input element are stores; write-only
// bandwidth is 3.9 TB/s against 5.9 TB/s reading), which is also why the BatchNorm+ReLU of the block
// that produced the source can ride along for free (BN = true, cross-layer fusion: that block's activation is never
// written -- nothing else reads it).
template <bool BN>
__global__ void __launch_bounds__(kThreads) bilinear2x_fwd_kernel(View x, View out, float sy, float sx, int rows,
                                                                   int strips, const float* __restrict__ scale,
                                                                   const float* __restrict__ shift) {
  const unsigned CV = static_cast<unsigned>(x.c) >> 3;
  const unsigned total = 1u * out.n * strips * out.w * CV;
  const unsigned xsh = static_cast<unsigned>(x.sh), osh = static_cast<unsigned>(out.sh);
  float sc[8], sh[8];
  unsigned cv_loaded = 0xffffffffu;
  for (unsigned i = blockIdx.x * kThreads + threadIdx.x; i < total; i += gridDim.x * kThreads) {
    unsigned t, cv;
    split_cv(out, i, t, cv);
    if (BN && cv != cv_loaded) {  // a thread keeps its channel group whenever the grid stride is a multiple of C / 8
      ld8f(scale + cv * 8, sc);
      ld8f(shift + cv * 8, sh);
      cv_loaded = cv;
    }
    const int ox = static_cast<int>(t % out.w);
    t /= out.w;
    const int s = static_cast<int>(t % strips);
    const int n = static_cast<int>(t / strips);
    int x0, x1;
    float lx0, lx1;
    src_index(sx, ox, x.w, x0, x1, lx0, lx1);
    const __nv_bfloat16* c0 = x.p + n * x.sn + x0 * x.sw + cv * 8;
    const __nv_bfloat16* c1 = x.p + n * x.sn + x1 * x.sw + cv * 8;
    float ra[8], rb[8];
    int r = -2;  // source row held in ra; rb holds the row below it (or the same row at the bottom edge)
    const int oy0 = s * rows, oy_end = min(out.h, oy0 + rows);
    // running output pointer and 32-bit source-row offsets (one image fits 32-bit element offsets: host check)
    __nv_bfloat16* po = out.p + n * out.sn + static_cast<long long>(oy0) * out.sh + ox * out.sw + cv * 8;
    for (int oy = oy0; oy < oy_end; ++oy, po += osh) {
      int y0, y1;
      float ly0, ly1;
      src_index(sy, oy, x.h, y0, y1, ly0, ly1);
      if (y0 != r) {
        if (y0 == r + 1) {
#pragma unroll
          for (int j = 0; j < 8; ++j) ra[j] = rb[j];
        } else {
          const unsigned o0 = static_cast<unsigned>(y0) * xsh;
          lerp_row<BN>(c0 + o0, c1 + o0, lx0, lx1, sc, sh, ra);
        }
        const unsigned o1 = static_cast<unsigned>(y1) * xsh;
        lerp_row<BN>(c0 + o1, c1 + o1, lx0, lx1, sc, sh, rb);
        r = y0;
      }
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = ly0 * ra[j] + ly1 * rb[j];
      stg16(po, pack8(o));
    }
  }
}

// Gather form of the adjoint: input pixel (iy,ix) sums every output pixel whose 4-tap stencil touches it.
__device__ __forceinline__ float tap_weight(float scale, int dst, int in, int target) {
  int i0, i1;
  float l0, l1;
  src_index(scale, dst, in, i0, i1, l0, l1);
  return (i0 == target ? l0 : 0.f) + (i1 == target ? l1 : 0.f);
}

// Output positions whose stencil touches input position `target`, with their weights (at most 6 for a x2 upsample:
// sources lie in the open interval (target-1, target+1), 2/scale wide).
template <int NT>
__device__ __forceinline__ int gather_taps(float scale, float inv, int in, int out, int target, int (&idx)[NT],
                                           float (&wgt)[NT]) {
  const int lo = max(0, static_cast<int>(floorf((target - 1) * inv)) - 1);
  const int hi = min(out - 1, static_cast<int>(ceilf((target + 1) * inv)) + 1);
  int cnt = 0;
#pragma unroll
  for (int k = 0; k < NT; ++k) {
    idx[k] = 0;
    wgt[k] = 0.f;
  }
  for (int o = lo; o <= hi; ++o) {
    const float w = tap_weight(scale, o, in, target);
    if (w != 0.f) {
#pragma unroll
      for (int k = 0; k < NT; ++k)  // (static indexing keeps idx / wgt in registers)
        if (k == cnt) {
          idx[k] = o;
          wgt[k] = w;
        }
      ++cnt;
    }
  }
  const int first = idx[0];
#pragma unroll
  for (int k = 0; k < NT; ++k)
    if (k >= cnt) idx[k] = first;  // unused slots re-read the first tap with weight 0
  return cnt < NT ? cnt : NT;
}

// Backward, separable: one thread = one 8-channel vector of one input COLUMN and a strip of `rows` input rows. It walks
// the output rows whose stencils touch the strip; per output row it gathers along x (the <= 6 output columns that touch
// its input column, found once per thread) and adds the result into two rolling row accumulators (source rows y0 and
// y0 + 1 of that output row). 4-5 loads per output row instead of 16-25 per input pixel.
// NT = tap slots along x: 4 is exact for every 2x align_corners upsampling of an input wider than 1 (each input
// column is touched by 4 output columns, 3 at the edges -- counted on the host, which falls back to 6 slots otherwise).
// The loop body is kept lean on purpose: ncu (profiles/r02n_elementwise_full.md) showed the previous version ISSUE-bound
// at 227 instructions per (thread, output row) -- 64-bit address products per tap, L2 prefetches, the row bounds tested
// inside the loop -- against ~100 of payload (4 loads, 32 conversions, 48 FMAs). Now: tap offsets are 32-bit element
// offsets computed once per thread, the row pointer advances by one add, the first / last output row of the strip are
// found before the loop (no continue / break inside), and since consecutive output rows map to source rows at most one
// apart (scale < 1/2) the accumulator roll is a single `if`.
template <int NT>
__global__ void __launch_bounds__(kThreads) bilinear2x_bwd_kernel(View dout, View dx, float sy, float sx, float isy,
                                                                   float isx, int rows, int strips) {
  const unsigned CV = static_cast<unsigned>(dx.c) >> 3;
  const unsigned total = 1u * dx.n * strips * dx.w * CV;
  const unsigned osh = static_cast<unsigned>(dout.sh);  // element strides of one image fit 32 bits (checked on the host)
  for (unsigned i = blockIdx.x * kThreads + threadIdx.x; i < total; i += gridDim.x * kThreads) {
    unsigned t, cv;
    split_cv(dx, i, t, cv);
    const int ix = static_cast<int>(t % dx.w);
    t /= dx.w;
    const int s = static_cast<int>(t % strips);
    const int n = static_cast<int>(t / strips);
    int ox[NT];
    float wx[NT];
    gather_taps<NT>(sx, isx, dx.w, dout.w, ix, ox, wx);
    unsigned toff[NT];
#pragma unroll
    for (int b = 0; b < NT; ++b) toff[b] = static_cast<unsigned>(ox[b]) * static_cast<unsigned>(dout.sw) + cv * 8;
    __nv_bfloat16* dst = dx.p + n * dx.sn + ix * dx.sw + cv * 8;
    const int ra = s * rows, rb = min(dx.h, ra + rows);  // owned input rows [ra, rb)
    // output rows [lo, hi] = those whose stencil (y0, y1) meets [ra, rb): conservative bounds, then exact
    int lo = max(0, static_cast<int>(floorf((ra - 1) * isy)) - 1);
    int hi = min(dout.h - 1, static_cast<int>(ceilf(rb * isy)) + 1);
    {
      int y0, y1;
      float l0, l1;
      for (;; ++lo) {
        src_index(sy, lo, dx.h, y0, y1, l0, l1);
        if (y1 >= ra || lo >= hi) break;
      }
      for (;; --hi) {
        src_index(sy, hi, dx.h, y0, y1, l0, l1);
        if (y0 < rb || hi <= lo) break;
      }
    }
    const __nv_bfloat16* rowp = dout.p + n * dout.sn + static_cast<long long>(lo) * dout.sh;
    float A[8], B[8];  // accumulators of input rows r and r + 1
#pragma unroll
    for (int j = 0; j < 8; ++j) A[j] = B[j] = 0.f;
    int r = ra - 1;
    for (int oy = lo; oy <= hi; ++oy, rowp += osh) {
      uint4 u[NT];
#pragma unroll
      for (int b = 0; b < NT; ++b) u[b] = ldg16(rowp + toff[b]);  // unused slots re-read tap 0 with weight 0
      int y0, y1;
      float ly0, ly1;
      src_index(sy, oy, dx.h, y0, y1, ly0, ly1);
      if (y0 != r) {  // row r is complete (y0 == r + 1; at the very first row possibly further: nothing owned lies between)
        if (r >= ra) stg16(dst + r * dx.sh, pack8(A));
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          A[j] = y0 == r + 1 ? B[j] : 0.f;
          B[j] = 0.f;
        }
        r = y0;
      }
      float g[8];
      {
        float v[8];
        unpack8(u[0], v);
#pragma unroll
        for (int j = 0; j < 8; ++j) g[j] = wx[0] * v[j];
      }
#pragma unroll
      for (int b = 1; b < NT; ++b) {
        float v[8];
        unpack8(u[b], v);
#pragma unroll
        for (int j = 0; j < 8; ++j) g[j] = fmaf(wx[b], v[j], g[j]);
      }
      const float la = y1 == y0 ? ly0 + ly1 : ly0, lb = y1 == y0 ? 0.f : ly1;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        A[j] = fmaf(la, g[j], A[j]);
        B[j] = fmaf(lb, g[j], B[j]);
      }
    }
    if (r >= ra && r < rb) stg16(dst + r * dx.sh, pack8(A));
    if (r + 1 >= ra && r + 1 < rb) stg16(dst + (r + 1) * dx.sh, pack8(B));
    for (int z = max(r + 2, ra); z < rb; ++z) stg16(dst + z * dx.sh, make_uint4(0u, 0u, 0u, 0u));  // untouched rows
  }
}

}  // namespace cvb

using namespace cvb;

static int up_shapes_ok(const cvb_view& x, const cvb_view& out, const char* who) {
  CVB_REQUIRE(out.n == x.n && out.c == x.c && out.h == 2 * x.h && out.w == 2 * x.w, CVB_ERR_INVALID_ARG,
              "%s: output %dx%dx%dx%d is not the 2x upsample of %dx%dx%dx%d", who, out.n, out.h, out.w, out.c, x.n,
              x.h, x.w, x.c);
  return CVB_OK;
}

static inline float ac_scale(int in, int out) { return out > 1 ? static_cast<float>(in - 1) / static_cast<float>(out - 1) : 0.f; }

// rows per thread: long strips amortise the strip's first row fetches, but the grid must still fill the machine
static int strip_rows(int n, int h, int w, int cv, int want) {
  int rows = want;
  while (rows > 2 && 1LL * n * ((h + rows - 1) / rows) * w * cv < 2LL * sm_count() * 2048) rows >>= 1;
  return rows;
}

static int launch_bilinear_fwd(cvb_view x, cvb_view out, const float* scale, const float* shift, void* stream) {
  int rc = check_view(x, "bilinear.x");
  if (rc) return rc;
  rc = check_view(out, "bilinear.out");
  if (rc) return rc;
  rc = up_shapes_ok(x, out, "bilinear2x_fwd");
  if (rc) return rc;
  CVB_REQUIRE(fits_u32(out), CVB_ERR_UNSUPPORTED, "bilinear2x_fwd: view too large for 32-bit indexing");
  CVB_REQUIRE(1LL * x.h * x.sh < (1LL << 31) && out.sh < (1LL << 31), CVB_ERR_UNSUPPORTED,
              "bilinear2x_fwd: one source image exceeds 32-bit element offsets");
  const int cv = out.c / 8;
  const int rows = strip_rows(out.n, out.h, out.w, cv, 16);
  const int strips = (out.h + rows - 1) / rows;
  long long total = 1LL * out.n * strips * out.w * cv;
  const int grid = ew_grid(total, kThreads);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const float fsy = ac_scale(x.h, out.h), fsx = ac_scale(x.w, out.w);
  if (scale)
    bilinear2x_fwd_kernel<true><<<grid, kThreads, 0, st>>>(to_dev(x), to_dev(out), fsy, fsx, rows, strips, scale, shift);
  else
    bilinear2x_fwd_kernel<false><<<grid, kThreads, 0, st>>>(to_dev(x), to_dev(out), fsy, fsx, rows, strips, nullptr, nullptr);
  CVB_LAUNCH_CHECK();
  return CVB_OK;
}

extern "C" int cvb_bilinear2x_fwd(cvb_view x, cvb_view out, void* stream) {
  return launch_bilinear_fwd(x, out, nullptr, nullptr, stream);
}

extern "C" int cvb_bn_relu_bilinear2x_fwd(cvb_view y, const float* scale, const float* shift, cvb_view out, void* stream) {
  CVB_REQUIRE(scale && shift, CVB_ERR_INVALID_ARG, "bn_relu_bilinear2x_fwd: null scale/shift");
  return launch_bilinear_fwd(y, out, scale, shift, stream);
}

// largest number of output positions whose stencil touches one input position (same fp32 index math as the kernels)
static int max_adjoint_taps(int in, int out) {
  const float scale = ac_scale(in, out);
  int best = 0, run_target = -1, run = 0;
  // taps of input position t = #{o: i0(o) == t with l0 != 0} + #{o: i1(o) == t != i0(o) with l1 != 0}; both sets are
  // contiguous in o, so two running counters indexed by position suffice
  static thread_local int cnt[4096];
  if (in > 4096) return 6;
  for (int t = 0; t < in; ++t) cnt[t] = 0;
  for (int o = 0; o < out; ++o) {
    const float src = scale * static_cast<float>(o);
    int i0 = static_cast<int>(src);
    if (i0 > in - 1) i0 = in - 1;
    const int i1 = i0 + ((i0 < in - 1) ? 1 : 0);
    const float l1 = src - static_cast<float>(i0), l0 = 1.f - l1;
    if (i1 == i0) {
      if (l0 + l1 != 0.f) ++cnt[i0];
    } else {
      if (l0 != 0.f) ++cnt[i0];
      if (l1 != 0.f) ++cnt[i1];
    }
  }
  (void)run_target; (void)run;
  for (int t = 0; t < in; ++t) best = cnt[t] > best ? cnt[t] : best;
  return best;
}

extern "C" int cvb_bilinear2x_bwd(cvb_view dout, cvb_view dx, void* stream) {
  int rc = check_view(dout, "bilinear_bwd.dout");
  if (rc) return rc;
  rc = check_view(dx, "bilinear_bwd.dx");
  if (rc) return rc;
  rc = up_shapes_ok(dx, dout, "bilinear2x_bwd");
  if (rc) return rc;
  float sy = ac_scale(dx.h, dout.h), sx = ac_scale(dx.w, dout.w);
  float isy = sy > 0.f ? 1.f / sy : static_cast<float>(dout.h), isx = sx > 0.f ? 1.f / sx : static_cast<float>(dout.w);
  CVB_REQUIRE(fits_u32(dout), CVB_ERR_UNSUPPORTED, "bilinear2x_bwd: view too large for 32-bit indexing");
  const int cv = dx.c / 8;
  const int rows = strip_rows(dx.n, dx.h, dx.w, cv, 16);
  const int strips = (dx.h + rows - 1) / rows;
  long long total = 1LL * dx.n * strips * dx.w * cv;
  const int grid = ew_grid(total, kThreads);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CVB_REQUIRE(1LL * dout.h * dout.sh < (1LL << 31), CVB_ERR_UNSUPPORTED,
              "bilinear2x_bwd: one image of dout exceeds 32-bit element offsets");
  if (max_adjoint_taps(dx.w, dout.w) <= 4)
    bilinear2x_bwd_kernel<4><<<grid, kThreads, 0, st>>>(to_dev(dout), to_dev(dx), sy, sx, isy, isx, rows, strips);
  else
    bilinear2x_bwd_kernel<6><<<grid, kThreads, 0, st>>>(to_dev(dout), to_dev(dx), sy, sx, isy, isx, rows, strips);
  CVB_LAUNCH_CHECK();
  return CVB_OK;
}
