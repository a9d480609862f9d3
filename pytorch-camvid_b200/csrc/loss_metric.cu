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

// Label tensors come as torch int64 (the reference API: masks.cuda(), train.py:127) or as the uint8 masks the device
// input stage keeps (cvb_label_type). LT = bytes per label.
template <int LT>
__device__ __forceinline__ long long load_label(const void* p, size_t i) {
  if (LT == 1) return static_cast<long long>(__ldg(static_cast<const uint8_t*>(p) + i));
  return __ldg(static_cast<const long long*>(p) + i);
}
// four consecutive labels starting at index 4*i4 (16-byte / 4-byte aligned base checked on the host)
template <int LT>
__device__ __forceinline__ void load_label4(const void* p, size_t i4, long long (&t)[4]) {
  if (LT == 1) {
    const uchar4 u = __ldg(static_cast<const uchar4*>(p) + i4);
    t[0] = u.x; t[1] = u.y; t[2] = u.z; t[3] = u.w;
  } else {
    const longlong2 a = __ldg(static_cast<const longlong2*>(p) + 2 * i4);
    const longlong2 b = __ldg(static_cast<const longlong2*>(p) + 2 * i4 + 1);
    t[0] = a.x; t[1] = a.y; t[2] = b.x; t[3] = b.y;
  }
}

// scratch layout of the loss kernels (double[4], zeroed by the caller): [0] sum of -log p[target] over counted pixels,
// [1] number of counted pixels (filled by the count pre-pass when reduction = mean, else by the main kernel),
// [2] number of labels that are neither ignore_index nor in [0, c), [3] block ticket.
__device__ __forceinline__ void block_accumulate(float loss, float cnt, float bad, bool add_count, int mean,
                                                 double* scratch, float* loss_out) {
  __shared__ float s_loss[kThreads / 32], s_cnt[kThreads / 32], s_bad[kThreads / 32];
  __shared__ bool s_last;
  loss = warp_sum(loss);
  cnt = warp_sum(cnt);
  bad = warp_sum(bad);
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
    s_loss[warp] = loss;
    s_cnt[warp] = cnt;
    s_bad[warp] = bad;
  }
  __syncthreads();
  if (warp == 0) {
    float l = lane < kThreads / 32 ? s_loss[lane] : 0.f;
    float c = lane < kThreads / 32 ? s_cnt[lane] : 0.f;
    float b = lane < kThreads / 32 ? s_bad[lane] : 0.f;
    l = warp_sum(l);
    c = warp_sum(c);
    b = warp_sum(b);
    if (lane == 0) {
      atomicAdd(scratch, static_cast<double>(l));
      if (add_count) atomicAdd(scratch + 1, static_cast<double>(c));
      if (b != 0.f) atomicAdd(scratch + 2, static_cast<double>(b));
      __threadfence();
      const double ticket = atomicAdd(scratch + 3, 1.0);
      s_last = ticket == static_cast<double>(gridDim.x - 1);
      if (s_last && loss_out) {  // every block's sums are visible: finish the reduction here, no host arithmetic
        __threadfence();
        volatile double* v = scratch;
        const double sum = v[0], count = v[1], invalid = v[2];
        double r = mean ? sum / count : sum;  // all pixels ignored: 0/0 = NaN, like torch
        // torch raises (device-side assert) on a label outside [0, c) that is not ignore_index; there is no assert to
        // raise across a C ABI without killing the context, so the loss is poisoned instead: it cannot go unnoticed
        if (invalid > 0.0) r = __longlong_as_double(0x7ff8000000000000LL);
        *loss_out = static_cast<float>(r);
      }
    }
  }
}

// reduction = mean: number of counted labels, BEFORE the main pass (its gradients are scaled by 1 / count)
template <int LT>
__global__ void __launch_bounds__(kThreads) ce_count_kernel(const void* __restrict__ target, long long total, int c,
                                                             long long ignore_index, double* scratch) {
  float cnt = 0.f;
  for (long long i = 1LL * blockIdx.x * kThreads + threadIdx.x; i < total; i += 1LL * gridDim.x * kThreads) {
    const long long t = load_label<LT>(target, static_cast<size_t>(i));
    cnt += (t != ignore_index && t >= 0 && t < c) ? 1.f : 0.f;
  }
  __shared__ float s_cnt[kThreads / 32];
  cnt = warp_sum(cnt);
  if ((threadIdx.x & 31) == 0) s_cnt[threadIdx.x >> 5] = cnt;
  __syncthreads();
  if (threadIdx.x < 32) {
    float c2 = threadIdx.x < kThreads / 32 ? s_cnt[threadIdx.x] : 0.f;
    c2 = warp_sum(c2);
    if (threadIdx.x == 0 && c2 != 0.f) atomicAdd(scratch + 1, static_cast<double>(c2));
  }
}

// gradient scale of one launch: grad_scale, divided by the counted-pixel total of the pre-pass for reduction = mean
__device__ __forceinline__ float ce_grad_scale(float grad_scale, int mean, const double* scratch) {
  if (!mean) return grad_scale;
  return static_cast<float>(static_cast<double>(grad_scale) / *reinterpret_cast<const volatile double*>(scratch + 1));
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

template <int LT>
__global__ void __launch_bounds__(kThreads) ce_nchw_f32_kernel(const float* __restrict__ logits,
                                                                const void* __restrict__ target, int n, int c,
                                                                long long hw, long long ignore_index, int mean,
                                                                double* scratch, float* loss_out,
                                                                float* __restrict__ dlogits, float grad_scale) {
  const long long total = 1LL * n * hw;
  const float gs = ce_grad_scale(grad_scale, mean, scratch);
  float loss = 0.f, cnt = 0.f, bad = 0.f;
  for (long long i = 1LL * blockIdx.x * kThreads + threadIdx.x; i < total; i += 1LL * gridDim.x * kThreads) {
    long long b = i / hw, p = i % hw;
    const float* src = logits + b * c * hw + p;
    float v[kMaxClasses];
#pragma unroll
    for (int k = 0; k < kMaxClasses; ++k)
      if (k < c) v[k] = __ldg(src + k * hw);
    long long tgt = load_label<LT>(target, static_cast<size_t>(i));
    const bool in_range = tgt >= 0 && tgt < c;
    bool counted = (tgt != ignore_index) && in_range;
    bad += (tgt != ignore_index && !in_range) ? 1.f : 0.f;
    loss += ce_pixel(v, c, tgt, counted, gs);
    cnt += counted ? 1.f : 0.f;
    if (dlogits) {
      float* dst = dlogits + b * c * hw + p;
#pragma unroll
      for (int k = 0; k < kMaxClasses; ++k)
        if (k < c) dst[k * hw] = v[k];
    }
  }
  block_accumulate(loss, cnt, bad, !mean, mean, scratch, loss_out);
}

// Fast path of the two NCHW fp32 kernels for a compile-time class count and hw % 4 == 0: one thread = four consecutive
// pixels, one 16-byte load per class plane (a warp reads 512 contiguous bytes of each plane), 32-bit indexing.
template <int C, int LT>
__global__ void __launch_bounds__(kThreads) ce_nchw_f32_vec4_kernel(const float* __restrict__ logits,
                                                                     const void* __restrict__ target, unsigned n,
                                                                     unsigned hw4, long long ignore_index, int mean,
                                                                     double* scratch, float* loss_out,
                                                                     float* __restrict__ dlogits, float grad_scale) {
  const unsigned total = n * hw4;
  const float gs = ce_grad_scale(grad_scale, mean, scratch);
  float loss = 0.f, cnt = 0.f, bad = 0.f;
  for (unsigned i = blockIdx.x * kThreads + threadIdx.x; i < total; i += gridDim.x * kThreads) {
    const unsigned b = i / hw4, q = i - b * hw4;
    const float4* src = reinterpret_cast<const float4*>(logits) + static_cast<size_t>(b) * C * hw4 + q;
    float4 v[C];
#pragma unroll
    for (int k = 0; k < C; ++k) v[k] = __ldg(src + static_cast<size_t>(k) * hw4);
    long long tg[4];
    load_label4<LT>(target, static_cast<size_t>(i), tg);
#pragma unroll
    for (int px = 0; px < 4; ++px) {
      float x[C];
#pragma unroll
      for (int k = 0; k < C; ++k) x[k] = px == 0 ? v[k].x : (px == 1 ? v[k].y : (px == 2 ? v[k].z : v[k].w));
      const long long tgt = tg[px];
      const bool in_range = tgt >= 0 && tgt < C;
      const bool counted = (tgt != ignore_index) && in_range;
      bad += (tgt != ignore_index && !in_range) ? 1.f : 0.f;
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
  block_accumulate(loss, cnt, bad, !mean, mean, scratch, loss_out);
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

template <int C, int LT>
__global__ void __launch_bounds__(kThreads) argmax_confmat_nchw_vec4_kernel(const float* __restrict__ logits,
                                                                             const void* __restrict__ gt, unsigned n,
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
    long long gl[4];
    load_label4<LT>(gt, static_cast<size_t>(i), gl);
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

template <int LT>
__global__ void __launch_bounds__(kThreads) ce_nhwc_bf16_kernel(View logits, int c, const void* __restrict__ target,
                                                                 long long ignore_index, int mean, double* scratch,
                                                                 float* loss_out, View dl, bool write_grad,
                                                                 float grad_scale) {
  const long long total = 1LL * logits.n * logits.h * logits.w;
  const float gs = ce_grad_scale(grad_scale, mean, scratch);
  const int cvs = (c + 7) >> 3;
  float loss = 0.f, cnt = 0.f, bad = 0.f;
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
    long long tgt = load_label<LT>(target, static_cast<size_t>(i));
    const bool in_range = tgt >= 0 && tgt < c;
    bool counted = (tgt != ignore_index) && in_range;
    bad += (tgt != ignore_index && !in_range) ? 1.f : 0.f;
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
  block_accumulate(loss, cnt, bad, !mean, mean, scratch, loss_out);
}

// ---- confusion matrix --------------------------------------------------------------------------
__device__ __forceinline__ void cm_flush(const unsigned int* hist, int cc, int64_t* cm) {
  __syncthreads();
  for (int i = threadIdx.x; i < cc; i += blockDim.x) {
    unsigned int v = hist[i];
    if (v) atomicAdd(reinterpret_cast<unsigned long long*>(cm) + i, static_cast<unsigned long long>(v));
  }
}

template <int LT>
__global__ void __launch_bounds__(kThreads) confmat_kernel(const void* __restrict__ pred,
                                                            const void* __restrict__ gt, long long count, int c,
                                                            long long ignore_label, int clamp_oob, int64_t* cm) {
  extern __shared__ unsigned int hist[];  // [warps][c*c]
  const int cc = c * c;
  for (int i = threadIdx.x; i < (kThreads / 32) * cc; i += kThreads) hist[i] = 0;
  __syncthreads();
  unsigned int* mine = hist + (threadIdx.x >> 5) * cc;
  for (long long i = 1LL * blockIdx.x * kThreads + threadIdx.x; i < count; i += 1LL * gridDim.x * kThreads) {
    long long p = load_label<LT>(pred, static_cast<size_t>(i)), g = load_label<LT>(gt, static_cast<size_t>(i));
    if (g == ignore_label) continue;
    if (clamp_oob) {
      if (p < 0 || p >= c) p = c - 1;
      if (g < 0 || g >= c) g = c - 1;
    }
    if (p >= 0 && p < c && g >= 0 && g < c) atomicAdd(&mine[g * c + p], 1u);
  }
  cm_flush_warps(hist, cc, cm);
}

template <int LT>
__global__ void __launch_bounds__(kThreads) argmax_confmat_nchw_kernel(const float* __restrict__ logits,
                                                                        const void* __restrict__ gt, int n, int c,
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
    long long g = load_label<LT>(gt, static_cast<size_t>(i));
    if (g >= 0 && g < c) atomicAdd(&hist[g * c + arg], 1u);
  }
  cm_flush(hist, c * c, cm);
}

template <int LT>
__global__ void __launch_bounds__(kThreads) argmax_confmat_nhwc_kernel(View logits, int c,
                                                                        const void* __restrict__ gt,
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
    long long gl = load_label<LT>(gt, static_cast<size_t>(i));
    if (gl >= 0 && gl < c) atomicAdd(&hist[gl * c + arg], 1u);
  }
  cm_flush(hist, c * c, cm);
}

}  // namespace cvb

using namespace cvb;

static int label_type_ok(int label_type, const char* who) {
  CVB_REQUIRE(label_type == CVB_LABEL_I64 || label_type == CVB_LABEL_U8, CVB_ERR_INVALID_ARG,
              "%s: label_type must be CVB_LABEL_I64 (8) or CVB_LABEL_U8 (1), got %d", who, label_type);
  return CVB_OK;
}

template <int LT>
static void launch_ce_count(const void* target, long long total, int c, long long ignore_index, double* scratch,
                            cudaStream_t st) {
  ce_count_kernel<LT><<<ew_grid(total, kThreads * 8), kThreads, 0, st>>>(target, total, c, ignore_index, scratch);
}

extern "C" int cvb_softmax_ce_nchw_f32(const float* logits, const void* target, int label_type, int n, int c, int h,
                                       int w, int64_t ignore_index, int mean, double* scratch, float* loss_out,
                                       float* dlogits, float grad_scale, void* stream) {
  CVB_REQUIRE(logits && target && scratch && loss_out, CVB_ERR_INVALID_ARG, "softmax_ce: null pointer");
  CVB_REQUIRE(n > 0 && h > 0 && w > 0, CVB_ERR_INVALID_ARG, "softmax_ce: empty input");
  CVB_REQUIRE(c > 0 && c <= kMaxClasses, CVB_ERR_UNSUPPORTED, "softmax_ce: %d classes (max %d)", c, kMaxClasses);
  int rc = label_type_ok(label_type, "softmax_ce");
  if (rc) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool u8 = label_type == CVB_LABEL_U8;
  long long hw = 1LL * h * w;
  if (mean) {
    if (u8) launch_ce_count<1>(target, n * hw, c, ignore_index, scratch, st);
    else launch_ce_count<8>(target, n * hw, c, ignore_index, scratch, st);
    CVB_LAUNCH_CHECK();
  }
  const bool aligned = (reinterpret_cast<uintptr_t>(logits) & 15) == 0 && (reinterpret_cast<uintptr_t>(target) & (u8 ? 3 : 15)) == 0 &&
                       (dlogits == nullptr || (reinterpret_cast<uintptr_t>(dlogits) & 15) == 0);
  if (c == 12 && hw % 4 == 0 && aligned && 1LL * n * hw < (1LL << 32)) {  // the CamVid class count, vectorised
    const unsigned hw4 = static_cast<unsigned>(hw / 4);
    const int grid = ew_grid(1LL * n * hw4, kThreads);
    if (u8)
      ce_nchw_f32_vec4_kernel<12, 1><<<grid, kThreads, 0, st>>>(logits, target, static_cast<unsigned>(n), hw4, ignore_index,
                                                                mean, scratch, loss_out, dlogits, grad_scale);
    else
      ce_nchw_f32_vec4_kernel<12, 8><<<grid, kThreads, 0, st>>>(logits, target, static_cast<unsigned>(n), hw4, ignore_index,
                                                                mean, scratch, loss_out, dlogits, grad_scale);
    CVB_LAUNCH_CHECK();
    return CVB_OK;
  }
  const int grid = ew_grid(1LL * n * hw, kThreads);
  if (u8)
    ce_nchw_f32_kernel<1><<<grid, kThreads, 0, st>>>(logits, target, n, c, hw, ignore_index, mean, scratch, loss_out, dlogits,
                                                     grad_scale);
  else
    ce_nchw_f32_kernel<8><<<grid, kThreads, 0, st>>>(logits, target, n, c, hw, ignore_index, mean, scratch, loss_out, dlogits,
                                                     grad_scale);
  CVB_LAUNCH_CHECK();
  return CVB_OK;
}

extern "C" int cvb_softmax_ce_nhwc_bf16(cvb_view logits, int c, const void* target, int label_type,
                                        int64_t ignore_index, int mean, double* scratch, float* loss_out,
                                        cvb_view dlogits, float grad_scale, void* stream) {
  int rc = check_view(logits, "softmax_ce.logits");
  if (rc) return rc;
  CVB_REQUIRE(target && scratch && loss_out, CVB_ERR_INVALID_ARG, "softmax_ce: null pointer");
  CVB_REQUIRE(c > 0 && c <= kMaxClasses && c <= logits.c, CVB_ERR_UNSUPPORTED, "softmax_ce: %d classes (max %d, view has %d)",
              c, kMaxClasses, logits.c);
  rc = label_type_ok(label_type, "softmax_ce");
  if (rc) return rc;
  bool wg = dlogits.ptr != nullptr;
  View dl = to_dev(logits);
  if (wg) {
    rc = check_view(dlogits, "softmax_ce.dlogits");
    if (rc) return rc;
    CVB_REQUIRE(dlogits.n == logits.n && dlogits.h == logits.h && dlogits.w == logits.w && dlogits.c >= c,
                CVB_ERR_INVALID_ARG, "softmax_ce: dlogits view does not match logits");
    dl = to_dev(dlogits);
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool u8 = label_type == CVB_LABEL_U8;
  long long total = 1LL * logits.n * logits.h * logits.w;
  if (mean) {
    if (u8) launch_ce_count<1>(target, total, c, ignore_index, scratch, st);
    else launch_ce_count<8>(target, total, c, ignore_index, scratch, st);
    CVB_LAUNCH_CHECK();
  }
  if (u8)
    ce_nhwc_bf16_kernel<1><<<ew_grid(total, kThreads), kThreads, 0, st>>>(to_dev(logits), c, target, ignore_index, mean, scratch,
                                                                         loss_out, dl, wg, grad_scale);
  else
    ce_nhwc_bf16_kernel<8><<<ew_grid(total, kThreads), kThreads, 0, st>>>(to_dev(logits), c, target, ignore_index, mean, scratch,
                                                                         loss_out, dl, wg, grad_scale);
  CVB_LAUNCH_CHECK();
  return CVB_OK;
}

static int cm_check(int c) {
  CVB_REQUIRE(c > 0 && c <= 64, CVB_ERR_UNSUPPORTED, "confusion matrix: %d classes (max 64)", c);
  return CVB_OK;
}

extern "C" int cvb_confusion_matrix(const void* pred, const void* gt, int label_type, int64_t count, int c,
                                    int64_t ignore_label, int clamp_oob, int64_t* cm, void* stream) {
  CVB_REQUIRE(cm, CVB_ERR_INVALID_ARG, "confusion_matrix: null cm");
  int rc = cm_check(c);
  if (rc) return rc;
  rc = label_type_ok(label_type, "confusion_matrix");
  if (rc) return rc;
  if (count == 0) return CVB_OK;  // empty input: nothing to add (sklearn returns zeros)
  CVB_REQUIRE(pred && gt && count > 0, CVB_ERR_INVALID_ARG, "confusion_matrix: null pointer or negative count");
  CVB_REQUIRE(c <= 38, CVB_ERR_UNSUPPORTED, "confusion_matrix: %d classes (per-warp counters fit 38)", c);
  const size_t smem = (kThreads / 32) * c * c * sizeof(unsigned int);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (label_type == CVB_LABEL_U8)
    confmat_kernel<1><<<ew_grid(count, kThreads, 8), kThreads, smem, st>>>(pred, gt, count, c, ignore_label, clamp_oob, cm);
  else
    confmat_kernel<8><<<ew_grid(count, kThreads, 8), kThreads, smem, st>>>(pred, gt, count, c, ignore_label, clamp_oob, cm);
  CVB_LAUNCH_CHECK();
  return CVB_OK;
}

extern "C" int cvb_argmax_confusion_nchw_f32(const float* logits, const void* gt, int label_type, int n, int c, int h,
                                             int w, int64_t* pred_or_null, int64_t* cm, void* stream) {
  CVB_REQUIRE(logits && gt && cm, CVB_ERR_INVALID_ARG, "argmax_confusion: null pointer");
  CVB_REQUIRE(n > 0 && h > 0 && w > 0, CVB_ERR_INVALID_ARG, "argmax_confusion: empty input");
  int rc = cm_check(c);
  if (rc) return rc;
  rc = label_type_ok(label_type, "argmax_confusion");
  if (rc) return rc;
  const bool u8 = label_type == CVB_LABEL_U8;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  long long hw = 1LL * h * w;
  const bool aligned = (reinterpret_cast<uintptr_t>(logits) & 15) == 0 && (reinterpret_cast<uintptr_t>(gt) & (u8 ? 3 : 15)) == 0 &&
                       (pred_or_null == nullptr || (reinterpret_cast<uintptr_t>(pred_or_null) & 15) == 0);
  if (c == 12 && hw % 4 == 0 && aligned && 1LL * n * hw < (1LL << 32)) {
    const unsigned hw4 = static_cast<unsigned>(hw / 4);
    const int grid = ew_grid(1LL * n * hw4, kThreads, 8);
    if (u8)
      argmax_confmat_nchw_vec4_kernel<12, 1><<<grid, kThreads, 0, st>>>(logits, gt, static_cast<unsigned>(n), hw4, pred_or_null, cm);
    else
      argmax_confmat_nchw_vec4_kernel<12, 8><<<grid, kThreads, 0, st>>>(logits, gt, static_cast<unsigned>(n), hw4, pred_or_null, cm);
    CVB_LAUNCH_CHECK();
    return CVB_OK;
  }
  const int grid = ew_grid(1LL * n * hw, kThreads, 4);
  if (u8)
    argmax_confmat_nchw_kernel<1><<<grid, kThreads, c * c * sizeof(unsigned int), st>>>(logits, gt, n, c, hw, pred_or_null, cm);
  else
    argmax_confmat_nchw_kernel<8><<<grid, kThreads, c * c * sizeof(unsigned int), st>>>(logits, gt, n, c, hw, pred_or_null, cm);
  CVB_LAUNCH_CHECK();
  return CVB_OK;
}

extern "C" int cvb_argmax_confusion_nhwc_bf16(cvb_view logits, int c, const void* gt, int label_type,
                                              int64_t* pred_or_null, int64_t* cm, void* stream) {
  int rc = check_view(logits, "argmax_confusion.logits");
  if (rc) return rc;
  CVB_REQUIRE(gt && cm, CVB_ERR_INVALID_ARG, "argmax_confusion: null pointer");
  rc = cm_check(c);
  if (rc) return rc;
  rc = label_type_ok(label_type, "argmax_confusion");
  if (rc) return rc;
  CVB_REQUIRE(c <= logits.c, CVB_ERR_INVALID_ARG, "argmax_confusion: %d classes but view has %d channels", c, logits.c);
  long long total = 1LL * logits.n * logits.h * logits.w;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int grid = ew_grid(total, kThreads, 4);
  if (label_type == CVB_LABEL_U8)
    argmax_confmat_nhwc_kernel<1><<<grid, kThreads, c * c * sizeof(unsigned int), st>>>(to_dev(logits), c, gt, pred_or_null, cm);
  else
    argmax_confmat_nhwc_kernel<8><<<grid, kThreads, c * c * sizeof(unsigned int), st>>>(to_dev(logits), c, gt, pred_or_null, cm);
  CVB_LAUNCH_CHECK();
  return CVB_OK;
}
