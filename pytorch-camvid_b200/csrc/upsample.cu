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

// One thread = VPT 8-channel vectors of ONE output pixel (vector indices t, t + TPP, ...: consecutive threads still
// touch consecutive 16-byte vectors), so the source-index arithmetic is paid once per VPT vectors.
template <int VPT>
__global__ void __launch_bounds__(kThreads) bilinear2x_fwd_kernel(View x, View out, float sy, float sx) {
  const unsigned CV = static_cast<unsigned>(x.c) >> 3;
  const unsigned TPP = CV / VPT;  // threads per pixel
  const unsigned total = 1u * out.n * out.h * out.w * TPP;
  for (unsigned i = blockIdx.x * kThreads + threadIdx.x; i < total; i += gridDim.x * kThreads) {
    unsigned t = i / TPP;
    const unsigned tv = i - t * TPP;
    const int ox = static_cast<int>(t % out.w);
    t /= out.w;
    const int oy = static_cast<int>(t % out.h);
    const int n = static_cast<int>(t / out.h);
    int y0, y1, x0, x1;
    float ly0, ly1, lx0, lx1;
    src_index(sy, oy, x.h, y0, y1, ly0, ly1);
    src_index(sx, ox, x.w, x0, x1, lx0, lx1);
    const __nv_bfloat16* p00 = x.p + voff(x, n, y0, x0);
    const __nv_bfloat16* p01 = x.p + voff(x, n, y0, x1);
    const __nv_bfloat16* p10 = x.p + voff(x, n, y1, x0);
    const __nv_bfloat16* p11 = x.p + voff(x, n, y1, x1);
    __nv_bfloat16* po = out.p + voff(out, n, oy, ox);
    uint4 ua[VPT], ub[VPT], uc[VPT], ud[VPT];
#pragma unroll
    for (int v = 0; v < VPT; ++v) {
      const unsigned off = (tv + v * TPP) * 8;
      ua[v] = ldg16(p00 + off);
      ub[v] = ldg16(p01 + off);
      uc[v] = ldg16(p10 + off);
      ud[v] = ldg16(p11 + off);
    }
#pragma unroll
    for (int v = 0; v < VPT; ++v) {
      float a[8], b[8], c[8], d[8], o[8];
      unpack8(ua[v], a);
      unpack8(ub[v], b);
      unpack8(uc[v], c);
      unpack8(ud[v], d);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = ly0 * (lx0 * a[j] + lx1 * b[j]) + ly1 * (lx0 * c[j] + lx1 * d[j]);
      stg16(po + (tv + v * TPP) * 8, pack8(o));
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

template <int VPT>
__global__ void __launch_bounds__(kThreads) bilinear2x_bwd_kernel(View dout, View dx, float sy, float sx, float isy,
                                                                   float isx) {
  const unsigned CV = static_cast<unsigned>(dx.c) >> 3;
  const unsigned TPP = CV / VPT;
  const unsigned total = 1u * dx.n * dx.h * dx.w * TPP;
  for (unsigned i = blockIdx.x * kThreads + threadIdx.x; i < total; i += gridDim.x * kThreads) {
    unsigned t = i / TPP;
    const unsigned tv = i - t * TPP;
    const int ix = static_cast<int>(t % dx.w);
    t /= dx.w;
    const int iy = static_cast<int>(t % dx.h);
    const int n = static_cast<int>(t / dx.h);
    int oy[6], ox[6];
    float wy[6], wx[6];
    const int ny = gather_taps(sy, isy, dx.h, dout.h, iy, oy, wy);
    const int nx = gather_taps(sx, isx, dx.w, dout.w, ix, ox, wx);
    float acc[VPT][8];
#pragma unroll
    for (int v = 0; v < VPT; ++v)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[v][j] = 0.f;
#pragma unroll
    for (int a = 0; a < 6; ++a) {
      if (a >= ny) break;
      const __nv_bfloat16* rowp = dout.p + n * dout.sn + oy[a] * dout.sh + tv * 8;
#pragma unroll
      for (int b = 0; b < 6; ++b) {
        if (b >= nx) break;
        const float wgt = wy[a] * wx[b];
        const __nv_bfloat16* pp = rowp + ox[b] * dout.sw;
        uint4 u[VPT];
#pragma unroll
        for (int v = 0; v < VPT; ++v) u[v] = ldg16(pp + v * TPP * 8);
#pragma unroll
        for (int v = 0; v < VPT; ++v) {
          float g[8];
          unpack8(u[v], g);
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[v][j] = fmaf(wgt, g[j], acc[v][j]);
        }
      }
    }
    __nv_bfloat16* po = dx.p + voff(dx, n, iy, ix);
#pragma unroll
    for (int v = 0; v < VPT; ++v) stg16(po + (tv + v * TPP) * 8, pack8(acc[v]));
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
  CVB_REQUIRE(fits_u32(out), CVB_ERR_UNSUPPORTED, "bilinear2x_fwd: view too large for 32-bit indexing");
  const int cv = out.c / 8;
  const int vpt = (cv % 4 == 0 && cv >= 32) ? 4 : ((cv % 2 == 0 && cv >= 16) ? 2 : 1);
  long long total = 1LL * out.n * out.h * out.w * (cv / vpt);
  const int grid = ew_grid(total, kThreads);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const float fsy = ac_scale(x.h, out.h), fsx = ac_scale(x.w, out.w);
  if (vpt == 4) bilinear2x_fwd_kernel<4><<<grid, kThreads, 0, st>>>(to_dev(x), to_dev(out), fsy, fsx);
  else if (vpt == 2) bilinear2x_fwd_kernel<2><<<grid, kThreads, 0, st>>>(to_dev(x), to_dev(out), fsy, fsx);
  else bilinear2x_fwd_kernel<1><<<grid, kThreads, 0, st>>>(to_dev(x), to_dev(out), fsy, fsx);
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
  CVB_REQUIRE(fits_u32(dout), CVB_ERR_UNSUPPORTED, "bilinear2x_bwd: view too large for 32-bit indexing");
  const int cv = dx.c / 8;
  const int vpt = (cv % 4 == 0 && cv >= 32) ? 4 : ((cv % 2 == 0 && cv >= 16) ? 2 : 1);
  long long total = 1LL * dx.n * dx.h * dx.w * (cv / vpt);
  const int grid = ew_grid(total, kThreads);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (vpt == 4) bilinear2x_bwd_kernel<4><<<grid, kThreads, 0, st>>>(to_dev(dout), to_dev(dx), sy, sx, isy, isx);
  else if (vpt == 2) bilinear2x_bwd_kernel<2><<<grid, kThreads, 0, st>>>(to_dev(dout), to_dev(dx), sy, sx, isy, isx);
  else bilinear2x_bwd_kernel<1><<<grid, kThreads, 0, st>>>(to_dev(dout), to_dev(dx), sy, sx, isy, isx);
  CVB_LAUNCH_CHECK();
  return CVB_OK;
}
