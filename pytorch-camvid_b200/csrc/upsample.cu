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

__global__ void __launch_bounds__(kThreads) bilinear2x_fwd_kernel(View x, View out, float sy, float sx) {
  const int CV = x.c >> 3;
  const unsigned total = 1u * out.n * out.h * out.w * CV;
  for (unsigned i = blockIdx.x * kThreads + threadIdx.x; i < total; i += gridDim.x * kThreads) {
    unsigned t, cv;
    split_cv(out, i, t, cv);
    int ox = static_cast<int>(t % out.w);
    t /= out.w;
    int oy = static_cast<int>(t % out.h);
    int n = static_cast<int>(t / out.h);
    int y0, y1, x0, x1;
    float ly0, ly1, lx0, lx1;
    src_index(sy, oy, x.h, y0, y1, ly0, ly1);
    src_index(sx, ox, x.w, x0, x1, lx0, lx1);
    float a[8], b[8], c[8], d[8], o[8];
    unpack8(ldg16(x.p + voff(x, n, y0, x0) + cv * 8), a);
    unpack8(ldg16(x.p + voff(x, n, y0, x1) + cv * 8), b);
    unpack8(ldg16(x.p + voff(x, n, y1, x0) + cv * 8), c);
    unpack8(ldg16(x.p + voff(x, n, y1, x1) + cv * 8), d);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = ly0 * (lx0 * a[j] + lx1 * b[j]) + ly1 * (lx0 * c[j] + lx1 * d[j]);
    stg16(out.p + voff(out, n, oy, ox) + cv * 8, pack8(o));
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
__device__ __forceinline__ int gather_taps(float scale, float inv, int in, int out, int target, int (&idx)[6],
                                           float (&wgt)[6]) {
  const int lo = max(0, static_cast<int>(floorf((target - 1) * inv)) - 1);
  const int hi = min(out - 1, static_cast<int>(ceilf((target + 1) * inv)) + 1);
  int cnt = 0;
  for (int o = lo; o <= hi; ++o) {
    const float w = tap_weight(scale, o, in, target);
    if (w != 0.f && cnt < 6) {
      idx[cnt] = o;
      wgt[cnt] = w;
      ++cnt;
    }
  }
  return cnt;
}

__global__ void __launch_bounds__(kThreads) bilinear2x_bwd_kernel(View dout, View dx, float sy, float sx, float isy,
                                                                   float isx) {
  const int CV = dx.c >> 3;
  const unsigned total = 1u * dx.n * dx.h * dx.w * CV;
  for (unsigned i = blockIdx.x * kThreads + threadIdx.x; i < total; i += gridDim.x * kThreads) {
    unsigned pix, cv;
    split_cv(dx, i, pix, cv);
    const int ix = static_cast<int>(pix % dx.w);
    const unsigned t = pix / dx.w;
    const int iy = static_cast<int>(t % dx.h);
    const int n = static_cast<int>(t / dx.h);
    int oy[6], ox[6];
    float wy[6], wx[6];
    const int ny = gather_taps(sy, isy, dx.h, dout.h, iy, oy, wy);
    const int nx = gather_taps(sx, isx, dx.w, dout.w, ix, ox, wx);
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
    for (int a = 0; a < 6; ++a) {
      if (a >= ny) break;
      const __nv_bfloat16* rowp = dout.p + n * dout.sn + oy[a] * dout.sh + cv * 8;
#pragma unroll
      for (int b = 0; b < 6; ++b) {
        if (b >= nx) break;
        float g[8];
        unpack8(ldg16(rowp + ox[b] * dout.sw), g);
        const float wgt = wy[a] * wx[b];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = fmaf(wgt, g[j], acc[j]);
      }
    }
    stg16(dx.p + voff(dx, n, iy, ix) + cv * 8, pack8(acc));
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

extern "C" int cvb_bilinear2x_fwd(cvb_view x, cvb_view out, void* stream) {
  int rc = check_view(x, "bilinear.x");
  if (rc) return rc;
  rc = check_view(out, "bilinear.out");
  if (rc) return rc;
  rc = up_shapes_ok(x, out, "bilinear2x_fwd");
  if (rc) return rc;
  long long total = 1LL * out.n * out.h * out.w * (out.c / 8);
  bilinear2x_fwd_kernel<<<ew_grid(total, kThreads), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      to_dev(x), to_dev(out), ac_scale(x.h, out.h), ac_scale(x.w, out.w));
  CVB_LAUNCH_CHECK();
  return CVB_OK;
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
  long long total = 1LL * dx.n * dx.h * dx.w * (dx.c / 8);
  bilinear2x_bwd_kernel<<<ew_grid(total, kThreads), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      to_dev(dout), to_dev(dx), sy, sx, isy, isx);
  CVB_LAUNCH_CHECK();
  return CVB_OK;
}
