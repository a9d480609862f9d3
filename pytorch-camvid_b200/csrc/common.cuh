// Host/device helpers shared by every translation unit of libcamvid_b200.so.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "../../include/camvid_b200.h"

namespace cvb {

// thread-local error string behind cvb_last_error()
void set_error(const char* fmt, ...);
int sm_count();

#define CVB_REQUIRE(cond, code, ...)  \
  do {                                \
    if (!(cond)) {                    \
      ::cvb::set_error(__VA_ARGS__);  \
      return (code);                  \
    }                                 \
  } while (0)

#define CVB_CUDA(expr)                                                                       \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess) {                                                                 \
      ::cvb::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return CVB_ERR_CUDA;                                                                   \
    }                                                                                        \
  } while (0)

#define CVB_LAUNCH_CHECK()                                                                    \
  do {                                                                                        \
    cudaError_t _e = cudaGetLastError();                                                      \
    if (_e != cudaSuccess) {                                                                  \
      ::cvb::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
      return CVB_ERR_CUDA;                                                                    \
    }                                                                                         \
  } while (0)

inline int check_view(const cvb_view& v, const char* name) {
  CVB_REQUIRE(v.ptr != nullptr, CVB_ERR_INVALID_ARG, "%s: null pointer", name);
  CVB_REQUIRE(v.n > 0 && v.h > 0 && v.w > 0 && v.c > 0, CVB_ERR_INVALID_ARG, "%s: empty view %dx%dx%dx%d", name, v.n,
              v.h, v.w, v.c);
  CVB_REQUIRE((v.c % 8) == 0, CVB_ERR_INVALID_ARG, "%s: channels %d not a multiple of 8", name, v.c);
  CVB_REQUIRE((reinterpret_cast<uintptr_t>(v.ptr) & 15) == 0, CVB_ERR_INVALID_ARG, "%s: pointer not 16-byte aligned",
              name);
  CVB_REQUIRE((v.sn % 8) == 0 && (v.sh % 8) == 0 && (v.sw % 8) == 0 && v.sw >= v.c, CVB_ERR_INVALID_ARG,
              "%s: strides (%lld,%lld,%lld) must be multiples of 8 with sw >= c", name, (long long)v.sn,
              (long long)v.sh, (long long)v.sw);
  return CVB_OK;
}

inline bool same_shape(const cvb_view& a, const cvb_view& b) {
  return a.n == b.n && a.h == b.h && a.w == b.w && a.c == b.c;
}

// Grid for grid-stride elementwise kernels: enough CTAs to fill the machine a few times over, never more than needed.
inline int ew_grid(int64_t items, int threads, int per_sm = 8) {
  static int env_per_sm = -1;  // development knob: CVB_EW_PER_SM overrides the resident-blocks-per-SM target
  if (env_per_sm < 0) {
    const char* e = getenv("CVB_EW_PER_SM");
    env_per_sm = e ? atoi(e) : 0;
  }
  if (env_per_sm > 0) per_sm = env_per_sm;
  int64_t need = (items + threads - 1) / threads;
  int64_t cap = static_cast<int64_t>(sm_count()) * per_sm;
  return static_cast<int>(need < 1 ? 1 : (need < cap ? need : cap));
}

// ---- device side -------------------------------------------------------------------------------
struct View {  // device copy of cvb_view with typed pointer
  __nv_bfloat16* p;
  int n, h, w, c;
  long long sn, sh, sw;
  int dense;     // pixels are equally spaced: offset(pixel index) = index * sw (true for whole buffers and channel slices)
  int cv_shift;  // log2(c / 8) when c / 8 is a power of two, else -1
};
inline View to_dev(const cvb_view& v) {
  View d;
  d.p = static_cast<__nv_bfloat16*>(v.ptr);
  d.n = v.n; d.h = v.h; d.w = v.w; d.c = v.c;
  d.sn = v.sn; d.sh = v.sh; d.sw = v.sw;
  d.dense = (v.sh == v.w * v.sw && v.sn == v.h * v.sh) ? 1 : 0;
  const int cv = v.c / 8;
  d.cv_shift = -1;
  for (int s = 0; s < 12; ++s)
    if ((1 << s) == cv) d.cv_shift = s;
  return d;
}
// Elementwise kernels index (pixel, 8-channel vector) pairs with 32-bit arithmetic; hosts check this bound.
inline bool fits_u32(const cvb_view& v) { return 1LL * v.n * v.h * v.w * (v.c / 8) < (1LL << 31); }

__device__ __forceinline__ long long voff(const View& v, int n, int h, int w) {
  return n * v.sn + h * v.sh + w * v.sw;
}
// element index -> (pixel, channel vector) without 64-bit division
__device__ __forceinline__ void split_cv(const View& v, unsigned i, unsigned& pix, unsigned& cv) {
  if (v.cv_shift >= 0) {
    pix = i >> v.cv_shift;
    cv = i & ((1u << v.cv_shift) - 1u);
  } else {
    const unsigned CV = static_cast<unsigned>(v.c) >> 3;
    pix = i / CV;
    cv = i - pix * CV;
  }
}
// offset of a linear pixel index (n, h, w flattened over the view's own extents)
__device__ __forceinline__ long long poff(const View& v, unsigned pix) {
  if (v.dense) return static_cast<long long>(pix) * v.sw;
  const unsigned w = pix % static_cast<unsigned>(v.w);
  const unsigned t = pix / static_cast<unsigned>(v.w);
  const unsigned h = t % static_cast<unsigned>(v.h);
  const unsigned n = t / static_cast<unsigned>(v.h);
  return n * v.sn + h * v.sh + w * v.sw;
}

// 8 bf16 <-> 8 floats
__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  f[0] = __uint_as_float(u.x << 16); f[1] = __uint_as_float(u.x & 0xFFFF0000u);
  f[2] = __uint_as_float(u.y << 16); f[3] = __uint_as_float(u.y & 0xFFFF0000u);
  f[4] = __uint_as_float(u.z << 16); f[5] = __uint_as_float(u.z & 0xFFFF0000u);
  f[6] = __uint_as_float(u.w << 16); f[7] = __uint_as_float(u.w & 0xFFFF0000u);
}
__device__ __forceinline__ uint32_t pk2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 u;
  u.x = pk2(f[0], f[1]); u.y = pk2(f[2], f[3]); u.z = pk2(f[4], f[5]); u.w = pk2(f[6], f[7]);
  return u;
}
__device__ __forceinline__ uint4 ldg16(const __nv_bfloat16* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
__device__ __forceinline__ void stg16(__nv_bfloat16* p, const uint4& v) { *reinterpret_cast<uint4*>(p) = v; }
// streaming variants for single-use data (evict-first: keep L2 for the tensors the next kernel re-reads)
__device__ __forceinline__ uint4 ldg16_cs(const __nv_bfloat16* p) { return __ldcs(reinterpret_cast<const uint4*>(p)); }
// L2 prefetch of the 128-byte line holding p: costs no register and no scoreboard entry, so a thread that walks a strip
// row by row (dependent iterations, few loads in flight) can keep several ROWS of DRAM requests outstanding
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void ld8f(const float* p, float (&f)[8]) {
  float4 a = __ldg(reinterpret_cast<const float4*>(p));
  float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}

}  // namespace cvb
