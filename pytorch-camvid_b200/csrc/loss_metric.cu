// nn.CrossEntropyLoss forward+backward fused (reference: train.py:105,130-131; eval.py:42,58) and
// argmax + confusion matrix (reference: train.py:191-194, utils.py:162-228, legacy/metrics.py:22-30).
// One thread = one pixel; fp32 math; HBM-bound.
#include "common.cuh"

namespace cvb {

constexpr int kThreads = 256;
constexpr int kMaxClasses = 32;  // classes held in registers

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// block-reduce (loss, count) and add to the double accumulators with one atomic pair per block
__device__ __forceinline__ void block_accumulate(float loss, float cnt, double* out) {
  __shared__ float s_loss[kThreads / 32], s_cnt[kThreads / 32];
  loss = warp_sum(loss);
  cnt = warp_sum(cnt);
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
    s_loss[warp] = loss;
    s_cnt[warp] = cnt;
  }
  __syncthreads();
  if (warp == 0) {
    float l = lane < kThreads / 32 ? s_loss[lane] : 0.f;
    float c = lane < kThreads / 32 ? s_cnt[lane] : 0.f;
    l = warp_sum(l);
    c = warp_sum(c);
    if (lane == 0) {
      atomicAdd(out, static_cast<double>(l));
      atomicAdd(out + 1, static_cast<double>(c));
    }
  }
}

// softmax / loss / gradient of one pixel held in registers
__device__ __forceinline__ float ce_pixel(float (&v)[kMaxClasses], int c, long long tgt, bool counted, float gs) {
  float mx = -INFINITY;
#pragma unroll
  for (int k = 0; k < kMaxClasses; ++k)
    if (k < c) mx = fmaxf(mx, v[k]);
  float se = 0.f, vt = 0.f;
#pragma unroll
  for (int k = 0; k < kMaxClasses; ++k)
    if (k < c) {
      float e = __expf(v[k] - mx);
      if (k == tgt) vt = v[k];
      v[k] = e;
      se += e;
    }
  float inv = counted ? gs / se : 0.f;
#pragma unroll
  for (int k = 0; k < kMaxClasses; ++k)
    if (k < c) v[k] = v[k] * inv - ((counted && k == tgt) ? gs : 0.f);
  return counted ? (__logf(se) + mx - vt) : 0.f;
}

__global__ void __launch_bounds__(kThreads) ce_nchw_f32_kernel(const float* __restrict__ logits,
                                                                const int64_t* __restrict__ target, int n, int c,
                                                                long long hw, long long ignore_index, double* out,
                                                                float* __restrict__ dlogits, float grad_scale,
                                                                const float* __restrict__ grad_scale_dev) {
  const long long total = 1LL * n * hw;
  const float gs = grad_scale * (grad_scale_dev ? __ldg(grad_scale_dev) : 1.f);
  float loss = 0.f, cnt = 0.f;
  for (long long i = 1LL * blockIdx.x * kThreads + threadIdx.x; i < total; i += 1LL * gridDim.x * kThreads) {
    long long b = i / hw, p = i % hw;
    const float* src = logits + b * c * hw + p;
    float v[kMaxClasses];
#pragma unroll
    for (int k = 0; k < kMaxClasses; ++k)
      if (k < c) v[k] = __ldg(src + k * hw);
    long long tgt = target[i];
    bool counted = (tgt != ignore_index) && tgt >= 0 && tgt < c;
    loss += ce_pixel(v, c, tgt, counted, gs);
    cnt += counted ? 1.f : 0.f;
    if (dlogits) {
      float* dst = dlogits + b * c * hw + p;
#pragma unroll
      for (int k = 0; k < kMaxClasses; ++k)
        if (k < c) dst[k * hw] = v[k];
    }
  }
  block_accumulate(loss, cnt, out);
}

// Fast path of the two NCHW fp32 kernels for a compile-time class count and hw % 4 == 0: one thread = four consecutive
// pixels, one 16-byte load per class plane (a warp reads 512 contiguous bytes of each plane), 32-bit indexing.
template <int C>
__global__ void __launch_bounds__(kThreads) ce_nchw_f32_vec4_kernel(const float* __restrict__ logits,
                                                                     const int64_t* __restrict__ target, unsigned n,
                                                                     unsigned hw4, long long ignore_index, double* out,
                                                                     float* __restrict__ dlogits, float grad_scale,
                                                                     const float* __restrict__ grad_scale_dev) {
  const unsigned total = n * hw4;
  const float gs = grad_scale * (grad_scale_dev ? __ldg(grad_scale_dev) : 1.f);
  float loss = 0.f, cnt = 0.f;
  for (unsigned i = blockIdx.x * kThreads + threadIdx.x; i < total; i += gridDim.x * kThreads) {
    const unsigned b = i / hw4, q = i - b * hw4;
    const float4* src = reinterpret_cast<const float4*>(logits) + static_cast<size_t>(b) * C * hw4 + q;
    float4 v[C];
#pragma unroll
    for (int k = 0; k < C; ++k) v[k] = __ldg(src + static_cast<size_t>(k) * hw4);
    const longlong2 t01 = __ldg(reinterpret_cast<const longlong2*>(target) + 2 * static_cast<size_t>(i));
    const longlong2 t23 = __ldg(reinterpret_cast<const longlong2*>(target) + 2 * static_cast<size_t>(i) + 1);
    const long long tg[4] = {t01.x, t01.y, t23.x, t23.y};
#pragma unroll
    for (int px = 0; px < 4; ++px) {
      float x[C];
#pragma unroll
      for (int k = 0; k < C; ++k) x[k] = px == 0 ? v[k].x : (px == 1 ? v[k].y : (px == 2 ? v[k].z : v[k].w));
      const long long tgt = tg[px];
      const bool counted = (tgt != ignore_index) && tgt >= 0 && tgt < C;
      float mx = x[0];
#pragma unroll
      for (int k = 1; k < C; ++k) mx = fmaxf(mx, x[k]);
      float se = 0.f, vt = 0.f;
#pragma unroll
      for (int k = 0; k < C; ++k) {
        const float e = __expf(x[k] - mx);
        if (k == tgt) vt = x[k];
        x[k] = e;
        se += e;
      }
      const float inv = counted ? gs / se : 0.f;
#pragma unroll
      for (int k = 0; k < C; ++k) {
        const float g = x[k] * inv - ((counted && k == tgt) ? gs : 0.f);
        if (px == 0) v[k].x = g; else if (px == 1) v[k].y = g; else if (px == 2) v[k].z = g; else v[k].w = g;
      }
      loss += counted ? (__logf(se) + mx - vt) : 0.f;
      cnt += counted ? 1.f : 0.f;
    }
    if (dlogits) {
      float4* dst = reinterpret_cast<float4*>(dlogits) + static_cast<size_t>(b) * C * hw4 + q;
#pragma unroll
      for (int k = 0; k < C; ++k) dst[static_cast<size_t>(k) * hw4] = v[k];
    }
  }
  block_accumulate(loss, cnt, out);
}

// per-warp private histograms (8 x c*c counters) cut shared-memory atomic contention on skewed label distributions
__device__ __forceinline__ void cm_flush_warps(const unsigned int* hist, int cc, int64_t* cm) {
  __syncthreads();
  for (int i = threadIdx.x; i < cc; i += blockDim.x) {
    unsigned int v = 0;
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) v += hist[w * cc + i];
    if (v) atomicAdd(reinterpret_cast<unsigned long long*>(cm) + i, static_cast<unsigned long long>(v));
  }
}

template <int C>
__global__ void __launch_bounds__(kThreads) argmax_confmat_nchw_vec4_kernel(const float* __restrict__ logits,
                                                                             const int64_t* __restrict__ gt, unsigned n,
                                                                             unsigned hw4, int64_t* __restrict__ pred,
                                                                             int64_t* cm) {
  __shared__ unsigned int hist[(kThreads / 32) * C * C];
  for (int i = threadIdx.x; i < (kThreads / 32) * C * C; i += kThreads) hist[i] = 0;
  __syncthreads();
  unsigned int* mine = hist + (threadIdx.x >> 5) * C * C;
  const unsigned total = n * hw4;
  for (unsigned i = blockIdx.x * kThreads + threadIdx.x; i < total; i += gridDim.x * kThreads) {
    const unsigned b = i / hw4, q = i - b * hw4;
    const float4* src = reinterpret_cast<const float4*>(logits) + static_cast<size_t>(b) * C * hw4 + q;
    float4 v[C];
#pragma unroll
    for (int k = 0; k < C; ++k) v[k] = __ldg(src + static_cast<size_t>(k) * hw4);
    const longlong2 g01 = __ldg(reinterpret_cast<const longlong2*>(gt) + 2 * static_cast<size_t>(i));
    const longlong2 g23 = __ldg(reinterpret_cast<const longlong2*>(gt) + 2 * static_cast<size_t>(i) + 1);
    const long long gl[4] = {g01.x, g01.y, g23.x, g23.y};
    long long arg[4];
#pragma unroll
    for (int px = 0; px < 4; ++px) {
      float best = px == 0 ? v[0].x : (px == 1 ? v[0].y : (px == 2 ? v[0].z : v[0].w));
      int a = 0;
#pragma unroll
      for (int k = 1; k < C; ++k) {
        const float x = px == 0 ? v[k].x : (px == 1 ? v[k].y : (px == 2 ? v[k].z : v[k].w));
        if (x > best || (x != x && best == best)) {  // first max; NaN is treated as maximal (torch.argmax)
          best = x;
          a = k;
        }
      }
      arg[px] = a;
      if (gl[px] >= 0 && gl[px] < C) atomicAdd(&mine[gl[px] * C + a], 1u);
    }
    if (pred) {
      longlong2* dp = reinterpret_cast<longlong2*>(pred) + 2 * static_cast<size_t>(i);
      dp[0] = make_longlong2(arg[0], arg[1]);
      dp[1] = make_longlong2(arg[2], arg[3]);
    }
  }
  cm_flush_warps(hist, C * C, cm);
}

__global__ void __launch_bounds__(kThreads) ce_nhwc_bf16_kernel(View logits, int c, const int64_t* __restrict__ target,
                                                                 long long ignore_index, double* out, View dl,
                                                                 bool write_grad, float grad_scale,
                                                                 const float* __restrict__ grad_scale_dev) {
  const long long total = 1LL * logits.n * logits.h * logits.w;
  const float gs = grad_scale * (grad_scale_dev ? __ldg(grad_scale_dev) : 1.f);
  const int cvs = (c + 7) >> 3;
  float loss = 0.f, cnt = 0.f;
  for (long long i = 1LL * blockIdx.x * kThreads + threadIdx.x; i < total; i += 1LL * gridDim.x * kThreads) {
    int w = static_cast<int>(i % logits.w);
    long long t = i / logits.w;
    int h = static_cast<int>(t % logits.h);
    int b = static_cast<int>(t / logits.h);
    const __nv_bfloat16* src = logits.p + voff(logits, b, h, w);
    float v[kMaxClasses];
#pragma unroll
    for (int g = 0; g < kMaxClasses / 8; ++g)
      if (g < cvs) {
        float f[8];
        unpack8(ldg16(src + g * 8), f);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[g * 8 + j] = f[j];
      }
    long long tgt = target[i];
    bool counted = (tgt != ignore_index) && tgt >= 0 && tgt < c;
    loss += ce_pixel(v, c, tgt, counted, gs);
    cnt += counted ? 1.f : 0.f;
    if (write_grad) {
      __nv_bfloat16* dst = dl.p + voff(dl, b, h, w);
      const int dcv = dl.c >> 3;
#pragma unroll
      for (int g = 0; g < kMaxClasses / 8; ++g)
        if (g < dcv) {
          float f[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) f[j] = (g * 8 + j < c) ? v[g * 8 + j] : 0.f;
          stg16(dst + g * 8, pack8(f));
        }
      for (int g = kMaxClasses / 8; g < dcv; ++g) stg16(dst + g * 8, make_uint4(0, 0, 0, 0));
    }
  }
  block_accumulate(loss, cnt, out);
}

// ---- confusion matrix --------------------------------------------------------------------------
__device__ __forceinline__ void cm_flush(const unsigned int* hist, int cc, int64_t* cm) {
  __syncthreads();
  for (int i = threadIdx.x; i < cc; i += blockDim.x) {
    unsigned int v = hist[i];
    if (v) atomicAdd(reinterpret_cast<unsigned long long*>(cm) + i, static_cast<unsigned long long>(v));
  }
}

__global__ void __launch_bounds__(kThreads) confmat_kernel(const int64_t* __restrict__ pred,
                                                            const int64_t* __restrict__ gt, long long count, int c,
                                                            long long ignore_label, int clamp_oob, int64_t* cm) {
  extern __shared__ unsigned int hist[];  // [warps][c*c]
  const int cc = c * c;
  for (int i = threadIdx.x; i < (kThreads / 32) * cc; i += kThreads) hist[i] = 0;
  __syncthreads();
  unsigned int* mine = hist + (threadIdx.x >> 5) * cc;
  for (long long i = 1LL * blockIdx.x * kThreads + threadIdx.x; i < count; i += 1LL * gridDim.x * kThreads) {
    long long p = __ldg(pred + i), g = __ldg(gt + i);
    if (g == ignore_label) continue;
    if (clamp_oob) {
      if (p < 0 || p >= c) p = c - 1;
      if (g < 0 || g >= c) g = c - 1;
    }
    if (p >= 0 && p < c && g >= 0 && g < c) atomicAdd(&mine[g * c + p], 1u);
  }
  cm_flush_warps(hist, cc, cm);
}

__global__ void __launch_bounds__(kThreads) argmax_confmat_nchw_kernel(const float* __restrict__ logits,
                                                                        const int64_t* __restrict__ gt, int n, int c,
                                                                        long long hw, int64_t* __restrict__ pred,
                                                                        int64_t* cm) {
  extern __shared__ unsigned int hist[];
  for (int i = threadIdx.x; i < c * c; i += kThreads) hist[i] = 0;
  __syncthreads();
  const long long total = 1LL * n * hw;
  for (long long i = 1LL * blockIdx.x * kThreads + threadIdx.x; i < total; i += 1LL * gridDim.x * kThreads) {
    long long b = i / hw, p = i % hw;
    const float* src = logits + b * c * hw + p;
    float best = __ldg(src);
    int arg = 0;
    for (int k = 1; k < c; ++k) {
      float v = __ldg(src + k * hw);
      if (v > best || (v != v && best == best)) {  // first max; NaN is treated as maximal (torch.argmax)
        best = v;
        arg = k;
      }
    }
    if (pred) pred[i] = arg;
    long long g = gt[i];
    if (g >= 0 && g < c) atomicAdd(&hist[g * c + arg], 1u);
  }
  cm_flush(hist, c * c, cm);
}

__global__ void __launch_bounds__(kThreads) argmax_confmat_nhwc_kernel(View logits, int c,
                                                                        const int64_t* __restrict__ gt,
                                                                        int64_t* __restrict__ pred, int64_t* cm) {
  extern __shared__ unsigned int hist[];
  for (int i = threadIdx.x; i < c * c; i += kThreads) hist[i] = 0;
  __syncthreads();
  const long long total = 1LL * logits.n * logits.h * logits.w;
  const int cvs = (c + 7) >> 3;
  for (long long i = 1LL * blockIdx.x * kThreads + threadIdx.x; i < total; i += 1LL * gridDim.x * kThreads) {
    int w = static_cast<int>(i % logits.w);
    long long t = i / logits.w;
    int h = static_cast<int>(t % logits.h);
    int b = static_cast<int>(t / logits.h);
    const __nv_bfloat16* src = logits.p + voff(logits, b, h, w);
    float best = 0.f;
    int arg = 0;
    for (int g = 0; g < cvs; ++g) {
      float f[8];
      unpack8(ldg16(src + g * 8), f);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        int k = g * 8 + j;
        if (k == 0) {
          best = f[0];
        } else if (k < c && (f[j] > best || (f[j] != f[j] && best == best))) {
          best = f[j];
          arg = k;
        }
      }
    }
    if (pred) pred[i] = arg;
    long long gl = gt[i];
    if (gl >= 0 && gl < c) atomicAdd(&hist[gl * c + arg], 1u);
  }
  cm_flush(hist, c * c, cm);
}

}  // namespace cvb

using namespace cvb;

extern "C" int cvb_softmax_ce_nchw_f32(const float* logits, const int64_t* target, int n, int c, int h, int w,
                                       int64_t ignore_index, double* loss_sum_count, float* dlogits, float grad_scale,
                                       const float* grad_scale_dev, void* stream) {
  CVB_REQUIRE(logits && target && loss_sum_count, CVB_ERR_INVALID_ARG, "softmax_ce: null pointer");
  CVB_REQUIRE(n > 0 && h > 0 && w > 0, CVB_ERR_INVALID_ARG, "softmax_ce: empty input");
  CVB_REQUIRE(c > 0 && c <= kMaxClasses, CVB_ERR_UNSUPPORTED, "softmax_ce: %d classes (max %d)", c, kMaxClasses);
  long long hw = 1LL * h * w;
  const bool aligned = (reinterpret_cast<uintptr_t>(logits) & 15) == 0 && (reinterpret_cast<uintptr_t>(target) & 15) == 0 &&
                       (dlogits == nullptr || (reinterpret_cast<uintptr_t>(dlogits) & 15) == 0);
  if (c == 12 && hw % 4 == 0 && aligned && 1LL * n * hw < (1LL << 32)) {  // the CamVid class count, vectorised
    const unsigned hw4 = static_cast<unsigned>(hw / 4);
    ce_nchw_f32_vec4_kernel<12><<<ew_grid(1LL * n * hw4, kThreads), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        logits, target, static_cast<unsigned>(n), hw4, ignore_index, loss_sum_count, dlogits, grad_scale, grad_scale_dev);
    CVB_LAUNCH_CHECK();
    return CVB_OK;
  }
  ce_nchw_f32_kernel<<<ew_grid(1LL * n * hw, kThreads), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      logits, target, n, c, hw, ignore_index, loss_sum_count, dlogits, grad_scale, grad_scale_dev);
  CVB_LAUNCH_CHECK();
  return CVB_OK;
}

extern "C" int cvb_softmax_ce_nhwc_bf16(cvb_view logits, int c, const int64_t* target, int64_t ignore_index,
                                        double* loss_sum_count, cvb_view dlogits, float grad_scale,
                                        const float* grad_scale_dev, void* stream) {
  int rc = check_view(logits, "softmax_ce.logits");
  if (rc) return rc;
  CVB_REQUIRE(target && loss_sum_count, CVB_ERR_INVALID_ARG, "softmax_ce: null pointer");
  CVB_REQUIRE(c > 0 && c <= kMaxClasses && c <= logits.c, CVB_ERR_UNSUPPORTED, "softmax_ce: %d classes (max %d, view has %d)",
              c, kMaxClasses, logits.c);
  bool wg = dlogits.ptr != nullptr;
  View dl = to_dev(logits);
  if (wg) {
    rc = check_view(dlogits, "softmax_ce.dlogits");
    if (rc) return rc;
    CVB_REQUIRE(dlogits.n == logits.n && dlogits.h == logits.h && dlogits.w == logits.w && dlogits.c >= c,
                CVB_ERR_INVALID_ARG, "softmax_ce: dlogits view does not match logits");
    dl = to_dev(dlogits);
  }
  long long total = 1LL * logits.n * logits.h * logits.w;
  ce_nhwc_bf16_kernel<<<ew_grid(total, kThreads), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      to_dev(logits), c, target, ignore_index, loss_sum_count, dl, wg, grad_scale, grad_scale_dev);
  CVB_LAUNCH_CHECK();
  return CVB_OK;
}

static int cm_check(int c) {
  CVB_REQUIRE(c > 0 && c <= 64, CVB_ERR_UNSUPPORTED, "confusion matrix: %d classes (max 64)", c);
  return CVB_OK;
}

extern "C" int cvb_confusion_matrix(const int64_t* pred, const int64_t* gt, int64_t count, int c,
                                    int64_t ignore_label, int clamp_oob, int64_t* cm, void* stream) {
  CVB_REQUIRE(cm, CVB_ERR_INVALID_ARG, "confusion_matrix: null cm");
  int rc = cm_check(c);
  if (rc) return rc;
  if (count == 0) return CVB_OK;  // empty input: nothing to add (sklearn returns zeros)
  CVB_REQUIRE(pred && gt && count > 0, CVB_ERR_INVALID_ARG, "confusion_matrix: null pointer or negative count");
  CVB_REQUIRE(c <= 38, CVB_ERR_UNSUPPORTED, "confusion_matrix: %d classes (per-warp counters fit 38)", c);
  confmat_kernel<<<ew_grid(count, kThreads, 8), kThreads, (kThreads / 32) * c * c * sizeof(unsigned int),
                   static_cast<cudaStream_t>(stream)>>>(pred, gt, count, c, ignore_label, clamp_oob, cm);
  CVB_LAUNCH_CHECK();
  return CVB_OK;
}

extern "C" int cvb_argmax_confusion_nchw_f32(const float* logits, const int64_t* gt, int n, int c, int h, int w,
                                             int64_t* pred_or_null, int64_t* cm, void* stream) {
  CVB_REQUIRE(logits && gt && cm, CVB_ERR_INVALID_ARG, "argmax_confusion: null pointer");
  CVB_REQUIRE(n > 0 && h > 0 && w > 0, CVB_ERR_INVALID_ARG, "argmax_confusion: empty input");
  int rc = cm_check(c);
  if (rc) return rc;
  long long hw = 1LL * h * w;
  const bool aligned = (reinterpret_cast<uintptr_t>(logits) & 15) == 0 && (reinterpret_cast<uintptr_t>(gt) & 15) == 0 &&
                       (pred_or_null == nullptr || (reinterpret_cast<uintptr_t>(pred_or_null) & 15) == 0);
  if (c == 12 && hw % 4 == 0 && aligned && 1LL * n * hw < (1LL << 32)) {
    const unsigned hw4 = static_cast<unsigned>(hw / 4);
    argmax_confmat_nchw_vec4_kernel<12><<<ew_grid(1LL * n * hw4, kThreads, 8), kThreads, 0,
                                          static_cast<cudaStream_t>(stream)>>>(logits, gt, static_cast<unsigned>(n), hw4,
                                                                                pred_or_null, cm);
    CVB_LAUNCH_CHECK();
    return CVB_OK;
  }
  argmax_confmat_nchw_kernel<<<ew_grid(1LL * n * hw, kThreads, 4), kThreads, c * c * sizeof(unsigned int),
                               static_cast<cudaStream_t>(stream)>>>(logits, gt, n, c, hw, pred_or_null, cm);
  CVB_LAUNCH_CHECK();
  return CVB_OK;
}

extern "C" int cvb_argmax_confusion_nhwc_bf16(cvb_view logits, int c, const int64_t* gt, int64_t* pred_or_null,
                                              int64_t* cm, void* stream) {
  int rc = check_view(logits, "argmax_confusion.logits");
  if (rc) return rc;
  CVB_REQUIRE(gt && cm, CVB_ERR_INVALID_ARG, "argmax_confusion: null pointer");
  rc = cm_check(c);
  if (rc) return rc;
  CVB_REQUIRE(c <= logits.c, CVB_ERR_INVALID_ARG, "argmax_confusion: %d classes but view has %d channels", c, logits.c);
  long long total = 1LL * logits.n * logits.h * logits.w;
  argmax_confmat_nhwc_kernel<<<ew_grid(total, kThreads, 4), kThreads, c * c * sizeof(unsigned int),
                               static_cast<cudaStream_t>(stream)>>>(to_dev(logits), c, gt, pred_or_null, cm);
  CVB_LAUNCH_CHECK();
  return CVB_OK;
}
