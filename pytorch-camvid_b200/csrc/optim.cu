// Fused multi-tensor AdamW: every parameter of the network in ONE launch.
//   reference: optimizer = optim.AdamW(net.parameters(), lr, weight_decay) / optimizer.step() in train.py:100,133
//   (torch.optim.AdamW, amsgrad=False, maximize=False). SURVEY section 8(f) rank 4: the stock foreach implementation
//   makes ~8 passes over 34.5 M parameters (460 us per step in the ncu launch list); one pass reads p, g, m, v and
//   writes p, m, v = 28 B per parameter.
//
// Math per element (identical to torch's single-tensor formulation, fp32):
//   p *= 1 - lr * wd;  m += (1 - b1) * (g - m);  v = b2 * v + (1 - b2) * g * g;
//   p -= (lr / (1 - b1^t)) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
#include "common.cuh"

namespace cvb {

constexpr int kOptThreads = 256;
constexpr int kOptChunk = 16384;  // elements per block

// one block = one chunk of one tensor
__device__ __forceinline__ void adamw_chunk(const cvb_adamw_entry* __restrict__ table, const int2* __restrict__ chunks,
                                            float decay, float b1c, float b2, float b2c, float step_size,
                                            float inv_bc2_sqrt, float eps) {
  const int2 ck = chunks[blockIdx.x];  // (tensor index, chunk index within the tensor)
  const cvb_adamw_entry e = table[ck.x];
  const long long begin = 1LL * ck.y * kOptChunk;
  const long long end = begin + kOptChunk < e.numel ? begin + kOptChunk : e.numel;
  float* __restrict__ p = e.param;
  const float* __restrict__ g = e.grad;
  float* __restrict__ m = e.exp_avg;
  float* __restrict__ v = e.exp_avg_sq;
  const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                     reinterpret_cast<uintptr_t>(v)) & 15) == 0;
  auto upd = [&](float& pp, float gg, float& mm, float& vv) {
    pp *= decay;
    mm += b1c * (gg - mm);
    vv = b2 * vv + b2c * gg * gg;
    pp -= step_size * mm / (sqrtf(vv) * inv_bc2_sqrt + eps);
  };
  if (vec) {
    const long long n4 = (end - begin) >> 2;
    for (long long i = threadIdx.x; i < n4; i += kOptThreads) {
      const long long o = begin + 4 * i;
      float4 P = *reinterpret_cast<float4*>(p + o);
      const float4 G = __ldcs(reinterpret_cast<const float4*>(g + o));
      float4 M = *reinterpret_cast<float4*>(m + o);
      float4 V = *reinterpret_cast<float4*>(v + o);
      upd(P.x, G.x, M.x, V.x);
      upd(P.y, G.y, M.y, V.y);
      upd(P.z, G.z, M.z, V.z);
      upd(P.w, G.w, M.w, V.w);
      *reinterpret_cast<float4*>(p + o) = P;
      *reinterpret_cast<float4*>(m + o) = M;
      *reinterpret_cast<float4*>(v + o) = V;
    }
    for (long long o = begin + 4 * n4 + threadIdx.x; o < end; o += kOptThreads) upd(p[o], g[o], m[o], v[o]);
  } else {
    for (long long o = begin + threadIdx.x; o < end; o += kOptThreads) upd(p[o], g[o], m[o], v[o]);
  }
}

__global__ void __launch_bounds__(kOptThreads) adamw_kernel(const cvb_adamw_entry* __restrict__ table,
                                                            const int2* __restrict__ chunks, float decay, float b1c,
                                                            float b2, float b2c, float step_size, float inv_bc2_sqrt,
                                                            float eps) {
  adamw_chunk(table, chunks, decay, b1c, b2, b2c, step_size, inv_bc2_sqrt, eps);
}

// The seven scalar factors read from DEVICE memory: what a CUDA graph needs, whose kernel arguments are frozen at
// capture while lr (OneCycleLR) and the bias corrections change every step.
__global__ void __launch_bounds__(kOptThreads) adamw_dev_kernel(const cvb_adamw_entry* __restrict__ table,
                                                                const int2* __restrict__ chunks,
                                                                const float* __restrict__ f) {
  adamw_chunk(table, chunks, __ldg(f), __ldg(f + 1), __ldg(f + 2), __ldg(f + 3), __ldg(f + 4), __ldg(f + 5), __ldg(f + 6));
}

// scalar factors of one step, in double like torch computes them on the host for python-float hyper-parameters
static int adamw_factors(float lr, float beta1, float beta2, float eps, float weight_decay, int64_t step, float* f) {
  CVB_REQUIRE(step >= 1 && beta1 >= 0.f && beta1 < 1.f && beta2 >= 0.f && beta2 < 1.f && eps >= 0.f,
              CVB_ERR_INVALID_ARG, "adamw: bad hyper-parameters (step %lld, betas %g %g, eps %g)", (long long)step,
              beta1, beta2, eps);
  const double bc1 = 1.0 - pow(static_cast<double>(beta1), static_cast<double>(step));
  const double bc2 = 1.0 - pow(static_cast<double>(beta2), static_cast<double>(step));
  f[0] = static_cast<float>(1.0 - static_cast<double>(lr) * static_cast<double>(weight_decay));  // decay
  f[1] = 1.f - beta1;
  f[2] = beta2;
  f[3] = 1.f - beta2;
  f[4] = static_cast<float>(static_cast<double>(lr) / bc1);  // step size
  f[5] = static_cast<float>(1.0 / sqrt(bc2));
  f[6] = eps;
  f[7] = 0.f;
  return CVB_OK;
}

}  // namespace cvb

using namespace cvb;

extern "C" int cvb_adamw_chunk_elems(void) { return kOptChunk; }

extern "C" int cvb_adamw_step(const cvb_adamw_entry* table, const int32_t* chunks, int n_chunks, float lr, float beta1,
                              float beta2, float eps, float weight_decay, int64_t step, void* stream) {
  CVB_REQUIRE(table && chunks && n_chunks > 0, CVB_ERR_INVALID_ARG, "adamw_step: empty table");
  float f[8];
  int rc = adamw_factors(lr, beta1, beta2, eps, weight_decay, step, f);
  if (rc) return rc;
  adamw_kernel<<<n_chunks, kOptThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      table, reinterpret_cast<const int2*>(chunks), f[0], f[1], f[2], f[3], f[4], f[5], f[6]);
  CVB_LAUNCH_CHECK();
  return CVB_OK;
}

extern "C" int cvb_adamw_factors(float lr, float beta1, float beta2, float eps, float weight_decay, int64_t step,
                                 float* factors_host) {
  CVB_REQUIRE(factors_host, CVB_ERR_INVALID_ARG, "adamw_factors: null pointer");
  return adamw_factors(lr, beta1, beta2, eps, weight_decay, step, factors_host);
}

extern "C" int cvb_adamw_step_dev(const cvb_adamw_entry* table, const int32_t* chunks, int n_chunks,
                                  const float* factors_dev, void* stream) {
  CVB_REQUIRE(table && chunks && n_chunks > 0 && factors_dev, CVB_ERR_INVALID_ARG, "adamw_step_dev: null pointer / empty table");
  adamw_dev_kernel<<<n_chunks, kOptThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      table, reinterpret_cast<const int2*>(chunks), factors_dev);
  CVB_LAUNCH_CHECK();
  return CVB_OK;
}
