// MaxPool2d(2,2) with indices / MaxUnpool2d(2) on NHWC bf16 views (reference: models/unet.py:92,
// models/segnet.py:79-80,86-116). One thread = 8 channels of one 2x2 window; 128-bit loads/stores.
// Tie rule is torch's: scan (0,0),(0,1),(1,0),(1,1); take val if (val > max) || isnan(val)  -> first max wins,
// last NaN wins. The window position (0..3) is kept as a uint8 code per output element.
#include "common.cuh"

namespace cvb {

constexpr int kThreads = 256;

__device__ __forceinline__ void window_max(const float (&v0)[8], const float (&v1)[8], const float (&v2)[8],
                                           const float (&v3)[8], float (&m)[8], uint32_t (&code)[8]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float best = v0[j];
    uint32_t k = 0;
    if (v1[j] > best || v1[j] != v1[j]) { best = v1[j]; k = 1; }
    if (v2[j] > best || v2[j] != v2[j]) { best = v2[j]; k = 2; }
    if (v3[j] > best || v3[j] != v3[j]) { best = v3[j]; k = 3; }
    m[j] = best;
    code[j] = k;
  }
}

__device__ __forceinline__ uint2 pack_code(const uint32_t (&c)[8]) {
  uint2 r;
  r.x = c[0] | (c[1] << 8) | (c[2] << 16) | (c[3] << 24);
  r.y = c[4] | (c[5] << 8) | (c[6] << 16) | (c[7] << 24);
  return r;
}
__device__ __forceinline__ void unpack_code(const uint2& r, uint32_t (&c)[8]) {
  c[0] = r.x & 255; c[1] = (r.x >> 8) & 255; c[2] = (r.x >> 16) & 255; c[3] = r.x >> 24;
  c[4] = r.y & 255; c[5] = (r.y >> 8) & 255; c[6] = (r.y >> 16) & 255; c[7] = r.y >> 24;
}

// Iteration domain: ceil(h/2) x ceil(w/2) windows so the odd last row/col of `a` is still produced when FUSE_BN.
template <bool FUSE_BN>
__global__ void __launch_bounds__(kThreads) maxpool_fwd_kernel(View x, View a, View out,
                                                                const float* __restrict__ scale,
                                                                const float* __restrict__ shift,
                                                                uint8_t* __restrict__ code) {
  const int CV = x.c >> 3;
  const int HO = (x.h + 1) >> 1, WO = (x.w + 1) >> 1;
  const unsigned total = 1u * x.n * HO * WO * CV;
  // a thread keeps its channel group across iterations when the grid stride is a multiple of CV: load scale / shift once
  const bool hoist = FUSE_BN && (kThreads % CV) == 0;
  float sc[8], sh[8];
  if (hoist) {
    const int cv0 = static_cast<int>((blockIdx.x * kThreads + threadIdx.x) % CV);
    ld8f(scale + cv0 * 8, sc);
    ld8f(shift + cv0 * 8, sh);
  }
  for (unsigned i = blockIdx.x * kThreads + threadIdx.x; i < total; i += gridDim.x * kThreads) {
    unsigned t, cvu;
    split_cv(x, i, t, cvu);
    const int cv = static_cast<int>(cvu);
    int wo = static_cast<int>(t % WO);
    t /= WO;
    int ho = static_cast<int>(t % HO);
    int n = static_cast<int>(t / HO);
    const int h0 = 2 * ho, w0 = 2 * wo;
    const bool hv = (h0 + 1) < x.h, wv = (w0 + 1) < x.w;
    float v[4][8];
    if (FUSE_BN && !hoist) {
      ld8f(scale + cv * 8, sc);
      ld8f(shift + cv * 8, sh);
    }
    // all window loads first, then the arithmetic and the stores
    uint4 raw[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int dh = k >> 1, dw = k & 1;
      const bool valid = (dh == 0 || hv) && (dw == 0 || wv);
      if (valid) raw[k] = ldg16(x.p + voff(x, n, h0 + dh, w0 + dw) + cv * 8);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int dh = k >> 1, dw = k & 1;
      const bool valid = (dh == 0 || hv) && (dw == 0 || wv);
      if (valid) {
        unpack8(raw[k], v[k]);
        if (FUSE_BN) {
#pragma unroll
          for (int j = 0; j < 8; ++j) v[k][j] = fmaxf(fmaf(v[k][j], sc[j], sh[j]), 0.f);
          uint4 pk = pack8(v[k]);
          stg16(a.p + voff(a, n, h0 + dh, w0 + dw) + cv * 8, pk);
          unpack8(pk, v[k]);  // pool over the values as stored (bf16), so codes match the saved activation
        }
      }
    }
    if (hv && wv) {
      float m[8];
      uint32_t cd[8];
      window_max(v[0], v[1], v[2], v[3], m, cd);
      stg16(out.p + voff(out, n, ho, wo) + cv * 8, pack8(m));
      if (code) {
        long long co = ((1LL * n * out.h + ho) * out.w + wo) * x.c + cv * 8;
        *reinterpret_cast<uint2*>(code + co) = pack_code(cd);
      }
    }
  }
}

// dx[window] = dout at code position (else 0). Covers the whole dx view including the odd last row/col.
// Also serves as MaxUnpool forward (dx == unpooled output, dout == pooled input).
template <bool ACCUM, bool RECOMPUTE>
__global__ void __launch_bounds__(kThreads) pool_scatter_kernel(View dout, View xin, View dx,
                                                                 const uint8_t* __restrict__ code) {
  const int CV = dx.c >> 3;
  const int HO = (dx.h + 1) >> 1, WO = (dx.w + 1) >> 1;
  const unsigned total = 1u * dx.n * HO * WO * CV;
  for (unsigned i = blockIdx.x * kThreads + threadIdx.x; i < total; i += gridDim.x * kThreads) {
    int cv = static_cast<int>(i % CV);
    unsigned t = i / CV;
    int wo = static_cast<int>(t % WO);
    t /= WO;
    int ho = static_cast<int>(t % HO);
    int n = static_cast<int>(t / HO);
    const int h0 = 2 * ho, w0 = 2 * wo;
    const bool hv = (h0 + 1) < dx.h, wv = (w0 + 1) < dx.w;
    const bool full = hv && wv && ho < dout.h && wo < dout.w;
    float g[8];
    uint32_t cd[8];
    if (full) {
      unpack8(ldg16(dout.p + voff(dout, n, ho, wo) + cv * 8), g);
      if (RECOMPUTE) {
        float v[4][8], m[8];
#pragma unroll
        for (int k = 0; k < 4; ++k) unpack8(ldg16(xin.p + voff(xin, n, h0 + (k >> 1), w0 + (k & 1)) + cv * 8), v[k]);
        window_max(v[0], v[1], v[2], v[3], m, cd);
      } else {
        long long co = ((1LL * n * dout.h + ho) * dout.w + wo) * dx.c + cv * 8;
        unpack_code(__ldg(reinterpret_cast<const uint2*>(code + co)), cd);
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int dh = k >> 1, dw = k & 1;
      const bool valid = (dh == 0 || hv) && (dw == 0 || wv);
      if (!valid) continue;
      __nv_bfloat16* p = dx.p + voff(dx, n, h0 + dh, w0 + dw) + cv * 8;
      float o[8];
      if (ACCUM) {
        if (!full) continue;  // nothing to add
        unpack8(*reinterpret_cast<const uint4*>(p), o);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] += (cd[j] == static_cast<uint32_t>(k)) ? g[j] : 0.f;
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = (full && cd[j] == static_cast<uint32_t>(k)) ? g[j] : 0.f;
      }
      stg16(p, pack8(o));
    }
  }
}

// dx[n,ho,wo,c] = dout[n, 2ho+dh, 2wo+dw, c] with (dh,dw) from the code  (MaxUnpool backward)
__global__ void __launch_bounds__(kThreads) pool_gather_kernel(View dout, View dx, const uint8_t* __restrict__ code) {
  const int CV = dx.c >> 3;
  const unsigned total = 1u * dx.n * dx.h * dx.w * CV;
  for (unsigned i = blockIdx.x * kThreads + threadIdx.x; i < total; i += gridDim.x * kThreads) {
    int cv = static_cast<int>(i % CV);
    unsigned t = i / CV;
    int wo = static_cast<int>(t % dx.w);
    t /= dx.w;
    int ho = static_cast<int>(t % dx.h);
    int n = static_cast<int>(t / dx.h);
    uint32_t cd[8];
    long long co = ((1LL * n * dx.h + ho) * dx.w + wo) * dx.c + cv * 8;
    unpack_code(__ldg(reinterpret_cast<const uint2*>(code + co)), cd);
    float v[4][8], o[8];
#pragma unroll
    for (int k = 0; k < 4; ++k)
      unpack8(ldg16(dout.p + voff(dout, n, 2 * ho + (k >> 1), 2 * wo + (k & 1)) + cv * 8), v[k]);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float r = v[0][j];
      r = cd[j] == 1 ? v[1][j] : r;
      r = cd[j] == 2 ? v[2][j] : r;
      r = cd[j] == 3 ? v[3][j] : r;
      o[j] = r;
    }
    stg16(dx.p + voff(dx, n, ho, wo) + cv * 8, pack8(o));
  }
}

// MaxPool backward FUSED with the BatchNorm+ReLU backward reduction of the block whose activation was pooled (SURVEY
// section 8(f) rank 1, producer side): the kernel that last writes `da` -- the gradient w.r.t. that block's activation -- also
// emits the per-channel sums (g, g*y) with g = da * [y*scale+shift > 0] that cvb_bn_relu_bwd_reduce would otherwise
// re-read da and y from HBM for: per element 6.75 B (y, da in/out, dout/4, code/4) instead of 6.5 + 4.
// One thread = 8 channels of one COLUMN of a row pair (the two pixels (2hp, w), (2hp+1, w) share their window's pooled
// gradient and code: 6 loads for 2 pixels; a whole window per thread needed 128+ registers and ran slower than the two
// kernels it replaced). A thread keeps one channel group; blocks stride over (row pair, column chunk) items, one
// partial row per block. Sums are taken from the bf16-rounded da that is stored, so the result is bit-identical to the
// unfused pair. The first version indexed pixels linearly (two divisions and three 64-bit stride products per pixel)
// and was ISSUE-bound at 267 instructions per pixel vector (ncu, profiles/r02n_elementwise_full.md); here the row
// bases are computed once per item, the item -> (image, row pair, chunk) split uses host-made reciprocals, and the
// window codes are compared as bytes.
struct PoolBwdGeom {
  unsigned chunks;       // column chunks (of ppb columns) per row pair
  unsigned hp;           // row pairs per image = ceil(h / 2)
  unsigned items;        // n * hp * chunks
  unsigned magic_chunks; // ceil(2^32 / chunks)      (exact quotient for every item index: checked on the host)
  unsigned magic_hp;     // ceil(2^32 / hp)
};

__device__ __forceinline__ unsigned div_magic(unsigned x, unsigned d, unsigned magic) {
  return d == 1 ? x : __umulhi(x, magic);
}

template <bool ACCUM>
__global__ void __launch_bounds__(kThreads, 3) pool_bwd_bn_reduce_kernel(View dout, View y, View dx,
                                                                       const uint8_t* __restrict__ code,
                                                                       const float* __restrict__ scale,
                                                                       const float* __restrict__ shift,
                                                                       float* __restrict__ partials, PoolBwdGeom gm) {
  __shared__ float red[kThreads * 2];
  const int CV = dx.c >> 3;
  const int ppb = kThreads / CV;
  const int cv = threadIdx.x % CV;
  const int pl = threadIdx.x / CV;
  float sc[8], sh[8], s1[8], s2[8];
  ld8f(scale + cv * 8, sc);
  ld8f(shift + cv * 8, sh);
#pragma unroll
  for (int j = 0; j < 8; ++j) s1[j] = s2[j] = 0.f;
  // 32-bit element strides (one image of each view and the code tensor fit 32-bit offsets: checked on the host); only the
  // image base is a 64-bit product
  const unsigned ysw = static_cast<unsigned>(y.sw), ysh = static_cast<unsigned>(y.sh), ysn = static_cast<unsigned>(y.sn);
  const unsigned xsw = static_cast<unsigned>(dx.sw), xsh = static_cast<unsigned>(dx.sh), xsn = static_cast<unsigned>(dx.sn);
  const unsigned osw = static_cast<unsigned>(dout.sw), osh = static_cast<unsigned>(dout.sh), osn = static_cast<unsigned>(dout.sn);
  const unsigned H = static_cast<unsigned>(dx.h), W = static_cast<unsigned>(dx.w), HO = static_cast<unsigned>(dout.h),
                 WO = static_cast<unsigned>(dout.w), C = static_cast<unsigned>(dx.c);
  const unsigned cvo = static_cast<unsigned>(cv) * 8u;
  for (unsigned it = blockIdx.x; it < gm.items; it += gridDim.x) {
    const unsigned rp = div_magic(it, gm.chunks, gm.magic_chunks);  // (image, row pair)
    const unsigned chunk = it - rp * gm.chunks;
    const unsigned n = div_magic(rp, gm.hp, gm.magic_hp);
    const unsigned hp = rp - n * gm.hp;
    const unsigned w = chunk * static_cast<unsigned>(ppb) + static_cast<unsigned>(pl);
    if (w >= W) continue;
    const unsigned h0 = 2u * hp, wo = w >> 1;
    const bool two = h0 + 1u < H;                  // odd last row: a single pixel, no window
    const bool full = two && hp < HO && wo < WO;   // odd last column: no window, the gradient passes through
    const __nv_bfloat16* yp = y.p + static_cast<size_t>(n) * ysn + (h0 * ysh + w * ysw + cvo);
    __nv_bfloat16* dp = dx.p + static_cast<size_t>(n) * xsn + (h0 * xsh + w * xsw + cvo);
    uint4 uy[2], ud[2], ug = make_uint4(0u, 0u, 0u, 0u);
    uint2 uc = make_uint2(0u, 0u);
    uy[0] = ldg16(yp);
    uy[1] = two ? ldg16(yp + ysh) : make_uint4(0u, 0u, 0u, 0u);
    ud[0] = ud[1] = make_uint4(0u, 0u, 0u, 0u);
    if (ACCUM) {
      ud[0] = *reinterpret_cast<const uint4*>(dp);
      if (two) ud[1] = *reinterpret_cast<const uint4*>(dp + xsh);
    }
    if (full) {
      ug = ldg16(dout.p + static_cast<size_t>(n) * osn + (hp * osh + wo * osw + cvo));
      uc = __ldg(reinterpret_cast<const uint2*>(code + (((n * HO + hp) * WO + wo) * C + cvo)));
    }
    float g[8];
    unpack8(ug, g);
    // the window position of this column's two pixels: (w & 1) in the top row, 2 + (w & 1) in the bottom row; a code
    // byte can only equal one of them, and none when the window does not exist (ug = 0 then anyway)
    const unsigned k0 = w & 1u, k1 = k0 + 2u;
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      if (q == 1 && !two) break;
      float o[8], fy[8];
      unpack8(ud[q], o);  // zeros unless ACCUM
      unpack8(uy[q], fy);
      const unsigned kq = q == 0 ? k0 : k1;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const unsigned cj = ((j < 4 ? uc.x : uc.y) >> (8 * (j & 3))) & 0xffu;
        o[j] += (full && cj == kq) ? g[j] : 0.f;
      }
      const uint4 pk = pack8(o);
      if (!ACCUM || full) stg16(dp + q * xsh, pk);
      unpack8(pk, o);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float ge = fmaf(fy[j], sc[j], sh[j]) > 0.f ? o[j] : 0.f;
        s1[j] += ge;
        s2[j] = fmaf(ge, fy[j], s2[j]);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {  // cross-thread combine through 2 KB (see bn_reduce_kernel: must fit beside a wgrad CTA)
    red[threadIdx.x * 2] = s1[j];
    red[threadIdx.x * 2 + 1] = s2[j];
    __syncthreads();
    if (threadIdx.x < 2 * CV) {
      const int which = threadIdx.x / CV, ocv = threadIdx.x - which * CV;
      float acc = 0.f;
      for (int l = 0; l < ppb; ++l) acc += red[(l * CV + ocv) * 2 + which];
      partials[(1LL * blockIdx.x * 2 + which) * dx.c + ocv * 8 + j] = acc;
    }
    __syncthreads();
  }
}

__global__ void pool_code_to_index_kernel(const uint8_t* __restrict__ code, int n, int ho, int wo, int c, int w_in,
                                          int64_t* __restrict__ idx) {
  const unsigned total = 1u * n * c * ho * wo;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    int x = static_cast<int>(i % wo);
    unsigned t = i / wo;
    int y = static_cast<int>(t % ho);
    t /= ho;
    int ch = static_cast<int>(t % c);
    int b = static_cast<int>(t / c);
    uint32_t k = code[((1LL * b * ho + y) * wo + x) * c + ch];
    idx[i] = 1LL * (2 * y + (k >> 1)) * w_in + (2 * x + (k & 1));
  }
}

}  // namespace cvb

using namespace cvb;

static int pool_shapes_ok(const cvb_view& x, const cvb_view& out, const char* who) {
  CVB_REQUIRE(out.n == x.n && out.c == x.c && out.h == x.h / 2 && out.w == x.w / 2, CVB_ERR_INVALID_ARG,
              "%s: pooled view %dx%dx%dx%d does not match input %dx%dx%dx%d (floor mode)", who, out.n, out.h, out.w,
              out.c, x.n, x.h, x.w, x.c);
  CVB_REQUIRE(x.h >= 2 && x.w >= 2, CVB_ERR_INVALID_ARG, "%s: input smaller than the 2x2 window", who);
  return CVB_OK;
}

extern "C" int cvb_maxpool2x2_fwd(cvb_view x, cvb_view out, uint8_t* code, void* stream) {
  int rc = check_view(x, "maxpool.x");
  if (rc) return rc;
  rc = check_view(out, "maxpool.out");
  if (rc) return rc;
  rc = pool_shapes_ok(x, out, "maxpool_fwd");
  if (rc) return rc;
  long long total = 1LL * x.n * ((x.h + 1) / 2) * ((x.w + 1) / 2) * (x.c / 8);
  maxpool_fwd_kernel<false><<<ew_grid(total, kThreads), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      to_dev(x), to_dev(x), to_dev(out), nullptr, nullptr, code);
  CVB_LAUNCH_CHECK();
  return CVB_OK;
}

extern "C" int cvb_bn_relu_maxpool2x2_fwd(cvb_view y, const float* scale, const float* shift, cvb_view a,
                                          cvb_view out, uint8_t* code, void* stream) {
  int rc = check_view(y, "bn_relu_maxpool.y");
  if (rc) return rc;
  rc = check_view(a, "bn_relu_maxpool.a");
  if (rc) return rc;
  rc = check_view(out, "bn_relu_maxpool.out");
  if (rc) return rc;
  CVB_REQUIRE(same_shape(y, a), CVB_ERR_INVALID_ARG, "bn_relu_maxpool: y and a shapes differ");
  CVB_REQUIRE(scale && shift, CVB_ERR_INVALID_ARG, "bn_relu_maxpool: null scale/shift");
  rc = pool_shapes_ok(y, out, "bn_relu_maxpool_fwd");
  if (rc) return rc;
  long long total = 1LL * y.n * ((y.h + 1) / 2) * ((y.w + 1) / 2) * (y.c / 8);
  maxpool_fwd_kernel<true><<<ew_grid(total, kThreads), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      to_dev(y), to_dev(a), to_dev(out), scale, shift, code);
  CVB_LAUNCH_CHECK();
  return CVB_OK;
}

extern "C" int cvb_maxpool2x2_bwd(cvb_view dout, const uint8_t* code, cvb_view x_or_null, cvb_view dx,
                                  int accumulate, void* stream) {
  int rc = check_view(dout, "maxpool_bwd.dout");
  if (rc) return rc;
  rc = check_view(dx, "maxpool_bwd.dx");
  if (rc) return rc;
  rc = pool_shapes_ok(dx, dout, "maxpool_bwd");
  if (rc) return rc;
  CVB_REQUIRE(code != nullptr || x_or_null.ptr != nullptr, CVB_ERR_INVALID_ARG,
              "maxpool_bwd: need either the index codes or the forward input");
  long long total = 1LL * dx.n * ((dx.h + 1) / 2) * ((dx.w + 1) / 2) * (dx.c / 8);
  int grid = ew_grid(total, kThreads);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (code) {
    if (accumulate)
      pool_scatter_kernel<true, false><<<grid, kThreads, 0, st>>>(to_dev(dout), to_dev(dx), to_dev(dx), code);
    else
      pool_scatter_kernel<false, false><<<grid, kThreads, 0, st>>>(to_dev(dout), to_dev(dx), to_dev(dx), code);
  } else {
    rc = check_view(x_or_null, "maxpool_bwd.x");
    if (rc) return rc;
    CVB_REQUIRE(same_shape(x_or_null, dx), CVB_ERR_INVALID_ARG, "maxpool_bwd: x and dx shapes differ");
    if (accumulate)
      pool_scatter_kernel<true, true><<<grid, kThreads, 0, st>>>(to_dev(dout), to_dev(x_or_null), to_dev(dx), nullptr);
    else
      pool_scatter_kernel<false, true><<<grid, kThreads, 0, st>>>(to_dev(dout), to_dev(x_or_null), to_dev(dx), nullptr);
  }
  CVB_LAUNCH_CHECK();
  return CVB_OK;
}

extern "C" int cvb_maxpool2x2_bwd_bn_reduce(cvb_view dout, const uint8_t* code, cvb_view y, const float* scale,
                                            const float* shift, cvb_view dx, int accumulate, float* partials, int rows,
                                            void* stream) {
  int rc = check_view(dout, "maxpool_bwd_bn_reduce.dout");
  if (rc) return rc;
  rc = check_view(dx, "maxpool_bwd_bn_reduce.dx");
  if (rc) return rc;
  rc = check_view(y, "maxpool_bwd_bn_reduce.y");
  if (rc) return rc;
  rc = pool_shapes_ok(dx, dout, "maxpool_bwd_bn_reduce");
  if (rc) return rc;
  CVB_REQUIRE(same_shape(y, dx), CVB_ERR_INVALID_ARG, "maxpool_bwd_bn_reduce: y and dx shapes differ");
  CVB_REQUIRE(scale && shift && partials && rows > 0, CVB_ERR_INVALID_ARG, "maxpool_bwd_bn_reduce: null pointer / rows");
  const int CV = dx.c / 8;
  CVB_REQUIRE(CV <= kThreads && (kThreads % CV) == 0, CVB_ERR_UNSUPPORTED,
              "maxpool_bwd_bn_reduce: channels %d must divide %d", dx.c, kThreads * 8);
  CVB_REQUIRE(1LL * dx.n * dx.h * dx.w < (1LL << 31), CVB_ERR_UNSUPPORTED,
              "maxpool_bwd_bn_reduce: view too large for 32-bit indexing");
  CVB_REQUIRE(code, CVB_ERR_INVALID_ARG, "maxpool_bwd_bn_reduce: null index codes (cvb_bn_relu_maxpool2x2_fwd writes them)");
  CVB_REQUIRE(1LL * dx.h * dx.sh < (1LL << 31) && 1LL * dout.h * dout.sh < (1LL << 31) && 1LL * y.h * y.sh < (1LL << 31) &&
                  dx.sn < (1LL << 31) && dout.sn < (1LL << 31) && y.sn < (1LL << 31) &&
                  1LL * dout.n * dout.h * dout.w * dx.c < (1LL << 32),
              CVB_ERR_UNSUPPORTED, "maxpool_bwd_bn_reduce: one image (or the code tensor) exceeds 32-bit element offsets");
  PoolBwdGeom gm;
  const int ppb = kThreads / CV;
  gm.chunks = static_cast<unsigned>((dx.w + ppb - 1) / ppb);
  gm.hp = static_cast<unsigned>((dx.h + 1) / 2);
  const long long items = 1LL * dx.n * gm.hp * gm.chunks;
  gm.items = static_cast<unsigned>(items);
  gm.magic_chunks = static_cast<unsigned>(((1ULL << 32) + gm.chunks - 1) / gm.chunks);
  gm.magic_hp = static_cast<unsigned>(((1ULL << 32) + gm.hp - 1) / gm.hp);
  // x / d == umulhi(x, ceil(2^32 / d)) for every x < 2^32 / d (round-up reciprocal: error x * (d - 1) / 2^32 < 1 / d... checked)
  CVB_REQUIRE(items * gm.chunks < (1LL << 32) && items * gm.hp < (1LL << 32), CVB_ERR_UNSUPPORTED,
              "maxpool_bwd_bn_reduce: view too large for the reciprocal index split");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (accumulate)
    pool_bwd_bn_reduce_kernel<true><<<rows, kThreads, 0, st>>>(to_dev(dout), to_dev(y), to_dev(dx), code, scale, shift, partials, gm);
  else
    pool_bwd_bn_reduce_kernel<false><<<rows, kThreads, 0, st>>>(to_dev(dout), to_dev(y), to_dev(dx), code, scale, shift, partials, gm);
  CVB_LAUNCH_CHECK();
  return CVB_OK;
}

extern "C" int cvb_maxunpool2x2_fwd(cvb_view x, const uint8_t* code, cvb_view out, void* stream) {
  int rc = check_view(x, "maxunpool.x");
  if (rc) return rc;
  rc = check_view(out, "maxunpool.out");
  if (rc) return rc;
  CVB_REQUIRE(code, CVB_ERR_INVALID_ARG, "maxunpool_fwd: null index codes");
  rc = pool_shapes_ok(out, x, "maxunpool_fwd");
  if (rc) return rc;
  long long total = 1LL * out.n * ((out.h + 1) / 2) * ((out.w + 1) / 2) * (out.c / 8);
  pool_scatter_kernel<false, false><<<ew_grid(total, kThreads), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      to_dev(x), to_dev(out), to_dev(out), code);
  CVB_LAUNCH_CHECK();
  return CVB_OK;
}

extern "C" int cvb_maxunpool2x2_bwd(cvb_view dout, const uint8_t* code, cvb_view dx, void* stream) {
  int rc = check_view(dout, "maxunpool_bwd.dout");
  if (rc) return rc;
  rc = check_view(dx, "maxunpool_bwd.dx");
  if (rc) return rc;
  CVB_REQUIRE(code, CVB_ERR_INVALID_ARG, "maxunpool_bwd: null index codes");
  rc = pool_shapes_ok(dout, dx, "maxunpool_bwd");
  if (rc) return rc;
  long long total = 1LL * dx.n * dx.h * dx.w * (dx.c / 8);
  pool_gather_kernel<<<ew_grid(total, kThreads), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      to_dev(dout), to_dev(dx), code);
  CVB_LAUNCH_CHECK();
  return CVB_OK;
}

extern "C" int cvb_pool_code_to_index(const uint8_t* code, int n, int ho, int wo, int c, int w_in, int64_t* idx_nchw,
                                      void* stream) {
  CVB_REQUIRE(code && idx_nchw, CVB_ERR_INVALID_ARG, "pool_code_to_index: null pointer");
  CVB_REQUIRE(n > 0 && ho > 0 && wo > 0 && c > 0 && w_in >= 2 * wo, CVB_ERR_INVALID_ARG,
              "pool_code_to_index: bad sizes");
  long long total = 1LL * n * c * ho * wo;
  pool_code_to_index_kernel<<<ew_grid(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(code, n, ho, wo, c,
                                                                                                 w_in, idx_nchw);
  CVB_LAUNCH_CHECK();
  return CVB_OK;
}
