// Layout conversion at the nn.Module boundary (NCHW fp32 <-> NHWC bf16), first-layer im2col, weight packing for the
// implicit-GEMM convolutions and view zero-fill. Memory-bound; pixel-fastest thread mapping keeps the NCHW side
// coalesced, each thread moves one 16-byte NHWC vector.
#include "common.cuh"

namespace cvb {

constexpr int kThreads = 256;

// thread = one pixel: its channels are gathered from the NCHW planes (consecutive threads read consecutive pixels of a
// plane: coalesced) and the whole NHWC row of the pixel is written as consecutive 16-byte vectors, so a warp writes
// 32 complete rows (no partially written 128-byte lines left for other blocks to finish).
template <int CVT>  // CVT > 0: compile-time number of 8-channel vectors (loop fully unrolled, loads batched)
__global__ void __launch_bounds__(kThreads) nchw_f32_to_nhwc_bf16_kernel(const float* __restrict__ src, int c_src,
                                                                          View dst) {
  const int CV = CVT > 0 ? CVT : (dst.c >> 3);
  const unsigned hw = static_cast<unsigned>(dst.h) * dst.w;
  const unsigned total = static_cast<unsigned>(dst.n) * hw;
  for (unsigned i = blockIdx.x * kThreads + threadIdx.x; i < total; i += gridDim.x * kThreads) {
    const unsigned n = i / hw, p = i - n * hw;
    const float* sp = src + static_cast<long long>(n) * c_src * hw + p;
    __nv_bfloat16* dp = dst.p + poff(dst, i);
#pragma unroll
    for (int cv = 0; cv < CV; ++cv) {
      if (cv * 8 >= c_src) {  // padding channels
        stg16(dp + cv * 8, make_uint4(0, 0, 0, 0));
        continue;
      }
      float f[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = cv * 8 + j;
        f[j] = c < c_src ? __ldg(sp + static_cast<long long>(c) * hw) : 0.f;
      }
      stg16(dp + cv * 8, pack8(f));
    }
  }
}

__global__ void __launch_bounds__(kThreads) nhwc_bf16_to_nchw_f32_kernel(View src, float* __restrict__ dst,
                                                                          int c_dst) {
  const int CV = (c_dst + 7) >> 3;
  const unsigned hw = static_cast<unsigned>(src.h) * src.w;
  const unsigned total = static_cast<unsigned>(src.n) * hw * CV;
  for (unsigned i = blockIdx.x * kThreads + threadIdx.x; i < total; i += gridDim.x * kThreads) {
    const unsigned p = i % hw;
    const unsigned t = i / hw;
    const int cv = static_cast<int>(t % CV);
    const int n = static_cast<int>(t / CV);
    float f[8];
    unpack8(ldg16(src.p + poff(src, n * hw + p) + cv * 8), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      int c = cv * 8 + j;
      if (c < c_dst) dst[(1LL * n * c_dst + c) * hw + p] = f[j];
    }
  }
}

// dst channel k = ci*9 + r*3 + s  <-  x[n, ci, h+r-1, w+s-1]; thread = one pixel (see above)
__global__ void __launch_bounds__(kThreads) im2col3x3_kernel(const float* __restrict__ src, int c_src, View dst) {
  const int CV = dst.c >> 3;
  const unsigned hw = static_cast<unsigned>(dst.h) * dst.w;
  const unsigned total = static_cast<unsigned>(dst.n) * hw;
  const int kmax = c_src * 9;
  for (unsigned i = blockIdx.x * kThreads + threadIdx.x; i < total; i += gridDim.x * kThreads) {
    const unsigned n = i / hw, p = i - n * hw;
    const int h = static_cast<int>(p / dst.w), w = static_cast<int>(p - (p / dst.w) * dst.w);
    const float* sp = src + static_cast<long long>(n) * c_src * hw;
    __nv_bfloat16* dp = dst.p + poff(dst, i);
    for (int cv = 0; cv < CV; ++cv) {
      float f[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int k = cv * 8 + j;
        float v = 0.f;
        if (k < kmax) {
          const int ci = k / 9, tap = k - ci * 9;
          const int hh = h + tap / 3 - 1, ww = w + tap % 3 - 1;
          if (hh >= 0 && hh < dst.h && ww >= 0 && ww < dst.w)
            v = __ldg(sp + static_cast<long long>(ci) * hw + static_cast<unsigned>(hh) * dst.w + ww);
        }
        f[j] = v;
      }
      stg16(dp + cv * 8, pack8(f));
    }
  }
}

// The case the networks use (3 input channels -> 27 of 64 destination channels): everything is a compile-time index,
// 27 predicated coalesced loads and eight 16-byte stores per pixel.
__global__ void __launch_bounds__(kThreads) im2col3x3_c3_kernel(const float* __restrict__ src, View dst) {
  const unsigned hw = static_cast<unsigned>(dst.h) * dst.w;
  const unsigned total = static_cast<unsigned>(dst.n) * hw;
  const int H = dst.h, W = dst.w;
  for (unsigned i = blockIdx.x * kThreads + threadIdx.x; i < total; i += gridDim.x * kThreads) {
    const unsigned n = i / hw, p = i - n * hw;
    const int h = static_cast<int>(p / W), w = static_cast<int>(p - (p / W) * W);
    const float* sp = src + static_cast<long long>(n) * 3 * hw + p;
    float f[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) f[k] = 0.f;
#pragma unroll
    for (int ci = 0; ci < 3; ++ci)
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int q = 0; q < 3; ++q) {
          const int hh = h + r - 1, ww = w + q - 1;
          if (hh >= 0 && hh < H && ww >= 0 && ww < W)
            f[ci * 9 + r * 3 + q] = __ldg(sp + static_cast<long long>(ci) * hw + (r - 1) * W + (q - 1));
        }
    __nv_bfloat16* dp = dst.p + poff(dst, i);
#pragma unroll
    for (int cv = 0; cv < 8; ++cv) {
      if (cv * 8 >= dst.c) break;  // 32-channel destination: the padding beyond channel 31 is never materialised
      uint4 u = make_uint4(0, 0, 0, 0);
      if (cv < 4) {
        u.x = pk2(f[cv * 8 + 0], f[cv * 8 + 1]);
        u.y = pk2(f[cv * 8 + 2], f[cv * 8 + 3]);
        u.z = pk2(f[cv * 8 + 4], f[cv * 8 + 5]);
        u.w = pk2(f[cv * 8 + 6], f[cv * 8 + 7]);
      }
      stg16(dp + cv * 8, u);
    }
  }
}

// dst[co][tap][ci] (bf16) <- w[co][ci][tap] (fp32 OIHW); taps==1: dst[co][k] <- w[co][k], k < cin*9
__global__ void pack_w_fprop_kernel(const float* __restrict__ w, int cout, int cin, int taps, int cout_pad,
                                    int cin_pad, __nv_bfloat16* __restrict__ dst) {
  const unsigned total = 1u * cout_pad * taps * cin_pad;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    int ci = static_cast<int>(i % cin_pad);
    unsigned t = i / cin_pad;
    int tap = static_cast<int>(t % taps);
    int co = static_cast<int>(t / taps);
    float v = 0.f;
    if (co < cout) {
      if (taps == 9) {
        if (ci < cin) v = w[(1LL * co * cin + ci) * 9 + tap];
      } else {
        if (ci < cin * 9) v = w[1LL * co * cin * 9 + ci];
      }
    }
    dst[i] = __float2bfloat16_rn(v);
  }
}

// dst[ci][tap'][co] (bf16) <- w[co][ci][8 - tap'] : the 180-degree rotated, in/out-transposed filter
__global__ void pack_w_dgrad_kernel(const float* __restrict__ w, int cout, int cin, int cout_pad, int cin_pad,
                                    __nv_bfloat16* __restrict__ dst) {
  const unsigned total = 1u * cin_pad * 9 * cout_pad;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    int co = static_cast<int>(i % cout_pad);
    unsigned t = i / cout_pad;
    int tap = static_cast<int>(t % 9);
    int ci = static_cast<int>(t / 9);
    float v = 0.f;
    if (co < cout && ci < cin) v = w[(1LL * co * cin + ci) * 9 + (8 - tap)];
    dst[i] = __float2bfloat16_rn(v);
  }
}

// One brick = 32 output x 32 input channels x 9 taps of one layer: read once (288 contiguous floats per output channel),
// transposed through shared memory into both packed operands with 64-byte contiguous bf16 runs.
__global__ void __launch_bounds__(256) pack_w_batch_kernel(const cvb_pack_entry* __restrict__ table) {
  __shared__ float tile[32][289];
  const cvb_pack_entry e = table[blockIdx.z];
  const int co0 = blockIdx.x * 32, ci0 = blockIdx.y * 32;
  if (co0 >= e.cout_pad || ci0 >= e.cin_pad) return;
  const int lane = threadIdx.x & 31, wq = threadIdx.x >> 5;
  const int cin = static_cast<int>(e.cin), cout = static_cast<int>(e.cout);
  for (int r = wq; r < 32; r += 8) {
    const int co = co0 + r;
    for (int k = lane; k < 288; k += 32) {
      const int ci = ci0 + k / 9;
      tile[r][k] = (co < cout && ci < cin) ? __ldg(e.w + (static_cast<long long>(co) * cin + ci0) * 9 + k) : 0.f;
    }
  }
  __syncthreads();
  __nv_bfloat16* df = static_cast<__nv_bfloat16*>(e.dst_fprop);
  __nv_bfloat16* dd = static_cast<__nv_bfloat16*>(e.dst_dgrad);
  const int cin_pad = static_cast<int>(e.cin_pad), cout_pad = static_cast<int>(e.cout_pad);
  for (int q = wq; q < 288; q += 8) {  // (row, tap) pairs; lane = the contiguous channel
    const int r = q / 9, tap = q - r * 9;
    // fprop operand [co][tap][ci]: row = output channel, lane = input channel
    df[(static_cast<long long>(co0 + r) * 9 + tap) * cin_pad + ci0 + lane] = __float2bfloat16_rn(tile[r][lane * 9 + tap]);
    // dgrad operand [ci][tap'][co] = w[co][ci][8 - tap']: row = input channel, lane = output channel
    if (dd) dd[(static_cast<long long>(ci0 + r) * 9 + tap) * cout_pad + co0 + lane] = __float2bfloat16_rn(tile[lane][r * 9 + (8 - tap)]);
  }
}

__global__ void __launch_bounds__(kThreads) zero_view_kernel(View v) {
  const int CV = v.c >> 3;
  const unsigned total = 1u * v.n * v.h * v.w * CV;
  for (unsigned i = blockIdx.x * kThreads + threadIdx.x; i < total; i += gridDim.x * kThreads) {
    unsigned pix, cv;
    split_cv(v, i, pix, cv);
    stg16(v.p + poff(v, pix) + cv * 8, make_uint4(0, 0, 0, 0));
  }
}

}  // namespace cvb

using namespace cvb;

extern "C" int cvb_nchw_f32_to_nhwc_bf16(const float* src, int c_src, cvb_view dst, void* stream) {
  int rc = check_view(dst, "nchw_to_nhwc.dst");
  if (rc) return rc;
  CVB_REQUIRE(src && c_src > 0 && c_src <= dst.c, CVB_ERR_INVALID_ARG, "nchw_to_nhwc: bad source (c_src=%d, dst.c=%d)",
              c_src, dst.c);
  long long total = 1LL * dst.n * dst.h * dst.w;
  if (dst.c == 64)
    nchw_f32_to_nhwc_bf16_kernel<8><<<ew_grid(total, kThreads), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        src, c_src, to_dev(dst));
  else
    nchw_f32_to_nhwc_bf16_kernel<0><<<ew_grid(total, kThreads), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        src, c_src, to_dev(dst));
  CVB_LAUNCH_CHECK();
  return CVB_OK;
}

extern "C" int cvb_nhwc_bf16_to_nchw_f32(cvb_view src, float* dst, int c_dst, void* stream) {
  int rc = check_view(src, "nhwc_to_nchw.src");
  if (rc) return rc;
  CVB_REQUIRE(dst && c_dst > 0 && c_dst <= src.c, CVB_ERR_INVALID_ARG, "nhwc_to_nchw: bad destination (c_dst=%d, src.c=%d)",
              c_dst, src.c);
  long long total = 1LL * src.n * src.h * src.w * ((c_dst + 7) / 8);
  nhwc_bf16_to_nchw_f32_kernel<<<ew_grid(total, kThreads), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      to_dev(src), dst, c_dst);
  CVB_LAUNCH_CHECK();
  return CVB_OK;
}

extern "C" int cvb_im2col3x3_nchw_f32(const float* src, int c_src, cvb_view dst, void* stream) {
  int rc = check_view(dst, "im2col.dst");
  if (rc) return rc;
  CVB_REQUIRE(src && c_src > 0 && c_src * 9 <= dst.c, CVB_ERR_INVALID_ARG,
              "im2col3x3: 9*c_src=%d does not fit the %d destination channels", c_src * 9, dst.c);
  long long total = 1LL * dst.n * dst.h * dst.w;
  if (c_src == 3 && (dst.c == 64 || dst.c == 32))
    im2col3x3_c3_kernel<<<ew_grid(total, kThreads), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(src, to_dev(dst));
  else
    im2col3x3_kernel<<<ew_grid(total, kThreads), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(src, c_src,
                                                                                                    to_dev(dst));
  CVB_LAUNCH_CHECK();
  return CVB_OK;
}

extern "C" int cvb_pack_weights_fprop(const float* w, int cout, int cin, int taps, int cout_pad, int cin_pad,
                                      void* dst, void* stream) {
  CVB_REQUIRE(w && dst, CVB_ERR_INVALID_ARG, "pack_weights_fprop: null pointer");
  CVB_REQUIRE(taps == 9 || taps == 1, CVB_ERR_INVALID_ARG, "pack_weights_fprop: taps must be 9 or 1");
  CVB_REQUIRE(cout > 0 && cin > 0 && cout_pad >= cout && cin_pad >= (taps == 9 ? cin : cin * 9), CVB_ERR_INVALID_ARG,
              "pack_weights_fprop: padded sizes too small");
  long long total = 1LL * cout_pad * taps * cin_pad;
  pack_w_fprop_kernel<<<ew_grid(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      w, cout, cin, taps, cout_pad, cin_pad, static_cast<__nv_bfloat16*>(dst));
  CVB_LAUNCH_CHECK();
  return CVB_OK;
}

extern "C" int cvb_pack_weights_dgrad(const float* w, int cout, int cin, int cout_pad, int cin_pad, void* dst,
                                      void* stream) {
  CVB_REQUIRE(w && dst, CVB_ERR_INVALID_ARG, "pack_weights_dgrad: null pointer");
  CVB_REQUIRE(cout > 0 && cin > 0 && cout_pad >= cout && cin_pad >= cin, CVB_ERR_INVALID_ARG,
              "pack_weights_dgrad: padded sizes too small");
  long long total = 1LL * cin_pad * 9 * cout_pad;
  pack_w_dgrad_kernel<<<ew_grid(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      w, cout, cin, cout_pad, cin_pad, static_cast<__nv_bfloat16*>(dst));
  CVB_LAUNCH_CHECK();
  return CVB_OK;
}

extern "C" int cvb_pack_weights_batch(const cvb_pack_entry* table, int count, int max_cout_pad, int max_cin_pad,
                                      void* stream) {
  CVB_REQUIRE(table && count > 0 && count <= 65535, CVB_ERR_INVALID_ARG, "pack_weights_batch: bad table");
  CVB_REQUIRE(max_cout_pad > 0 && max_cin_pad > 0 && (max_cout_pad % 32) == 0 && (max_cin_pad % 32) == 0,
              CVB_ERR_INVALID_ARG, "pack_weights_batch: padded extents must be positive multiples of 32");
  dim3 grid(max_cout_pad / 32, max_cin_pad / 32, count);
  pack_w_batch_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(table);
  CVB_LAUNCH_CHECK();
  return CVB_OK;
}

extern "C" int cvb_zero_view(cvb_view v, void* stream) {
  int rc = check_view(v, "zero_view");
  if (rc) return rc;
  long long total = 1LL * v.n * v.h * v.w * (v.c / 8);
  zero_view_kernel<<<ew_grid(total, kThreads), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(to_dev(v));
  CVB_LAUNCH_CHECK();
  return CVB_OK;
}
