// Device input stage: transforms.ToTensor + transforms.Normalize of the reference
// (transforms.py:485-538) and the `.long()` of the mask, applied on the GPU to the uint8 HWC image / uint8 mask that
// cv2 and the dataset produce (dataset/camvid.py:161-173), so that 11 MB instead of 55 MB per 16-image batch cross
// PCIe (train.py:126-127 copies fp32 NCHW images and int64 masks).
//
// Arithmetic = torch's, rounding for rounding (checked bit-exactly against the reference transforms, tests/golden):
//   ToTensor:   img.float() / 255.0                -> IEEE fp32 division by 255
//   Normalize:  img.sub_(mean[c]).div_(std[c])     -> fp32 subtraction, then IEEE fp32 division (two roundings: the
//                                                     intrinsics below keep nvcc from contracting them into an FMA)
#include "common.cuh"

namespace cvb {

constexpr int kThreads = 256;

struct NormParams {
  float mean[4];
  float std[4];
};

__device__ __forceinline__ float to_tensor_normalize(unsigned v, float mean, float std) {
  return __fdiv_rn(__fsub_rn(__fdiv_rn(static_cast<float>(v), 255.0f), mean), std);
}

// One thread = four consecutive pixels of one image: 12 contiguous bytes in (three 32-bit loads; 4 * C bytes is a
// multiple of 4 for every C), one 16-byte store per channel plane out, one 4-byte load + two 16-byte stores for the
// mask. Requires h*w % 4 == 0 (host checks; the scalar kernel below covers the rest).
template <int C>
__global__ void __launch_bounds__(kThreads) input_stage_vec4_kernel(const uint8_t* __restrict__ img, unsigned n,
                                                                     unsigned hw4, NormParams np,
                                                                     float* __restrict__ out,
                                                                     const uint8_t* __restrict__ mask,
                                                                     int64_t* __restrict__ mask_out) {
  const unsigned total = n * hw4;
  for (unsigned i = blockIdx.x * kThreads + threadIdx.x; i < total; i += gridDim.x * kThreads) {
    const unsigned b = i / hw4, q = i - b * hw4;
    if (img) {
      const uint32_t* src = reinterpret_cast<const uint32_t*>(img) + static_cast<size_t>(i) * C;  // 4 px * C bytes
      uint32_t wd[C];
#pragma unroll
      for (int k = 0; k < C; ++k) wd[k] = __ldg(src + k);
      float v[4][C];
#pragma unroll
      for (int e = 0; e < 4 * C; ++e) {  // byte e of the 4-pixel run = pixel e / C, channel e % C
        const unsigned byte = (wd[e >> 2] >> (8 * (e & 3))) & 0xffu;
        v[e / C][e % C] = to_tensor_normalize(byte, np.mean[e % C], np.std[e % C]);
      }
#pragma unroll
      for (int c = 0; c < C; ++c) {
        float4* dst = reinterpret_cast<float4*>(out) + (static_cast<size_t>(b) * C + c) * hw4 + q;
        *dst = make_float4(v[0][c], v[1][c], v[2][c], v[3][c]);
      }
    }
    if (mask_out) {
      const uchar4 m = __ldg(reinterpret_cast<const uchar4*>(mask) + i);
      longlong2* dm = reinterpret_cast<longlong2*>(mask_out) + 2 * static_cast<size_t>(i);
      dm[0] = make_longlong2(m.x, m.y);
      dm[1] = make_longlong2(m.z, m.w);
    }
  }
}

__global__ void __launch_bounds__(kThreads) input_stage_kernel(const uint8_t* __restrict__ img, long long n,
                                                                long long hw, int c, NormParams np,
                                                                float* __restrict__ out,
                                                                const uint8_t* __restrict__ mask,
                                                                int64_t* __restrict__ mask_out) {
  const long long total = n * hw;
  for (long long i = 1LL * blockIdx.x * kThreads + threadIdx.x; i < total; i += 1LL * gridDim.x * kThreads) {
    const long long b = i / hw, p = i - b * hw;
    if (img)
      for (int k = 0; k < c; ++k)
        out[(b * c + k) * hw + p] = to_tensor_normalize(__ldg(img + i * c + k), np.mean[k], np.std[k]);
    if (mask_out) mask_out[i] = __ldg(mask + i);
  }
}

}  // namespace cvb

using namespace cvb;

extern "C" int cvb_input_stage_u8(const uint8_t* img_u8, int n, int h, int w, int c, const float* mean_host,
                                  const float* std_host, float* out_nchw, const uint8_t* mask_u8, int64_t* mask_i64,
                                  void* stream) {
  CVB_REQUIRE(n > 0 && h > 0 && w > 0, CVB_ERR_INVALID_ARG, "input_stage: empty batch");
  CVB_REQUIRE((img_u8 != nullptr) == (out_nchw != nullptr), CVB_ERR_INVALID_ARG,
              "input_stage: image source and destination must both be given or both be NULL");
  CVB_REQUIRE((mask_i64 == nullptr) || (mask_u8 != nullptr), CVB_ERR_INVALID_ARG, "input_stage: mask_i64 without mask_u8");
  CVB_REQUIRE(img_u8 || mask_i64, CVB_ERR_INVALID_ARG, "input_stage: nothing to do");
  NormParams np;
  for (int k = 0; k < 4; ++k) {
    np.mean[k] = 0.f;
    np.std[k] = 1.f;
  }
  if (img_u8) {
    CVB_REQUIRE(c >= 1 && c <= 4, CVB_ERR_UNSUPPORTED, "input_stage: %d channels (1..4 supported)", c);
    CVB_REQUIRE(mean_host && std_host, CVB_ERR_INVALID_ARG, "input_stage: null mean / std");
    for (int k = 0; k < c; ++k) {
      np.mean[k] = mean_host[k];
      np.std[k] = std_host[k];
    }
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long hw = 1LL * h * w;
  auto al = [](const void* p, uintptr_t a) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) & (a - 1)) == 0; };
  if ((!img_u8 || c == 3) && hw % 4 == 0 && n * hw < (1LL << 32) && al(img_u8, 4) && al(out_nchw, 16) && al(mask_u8, 4) &&
      al(mask_i64, 16)) {
    const unsigned hw4 = static_cast<unsigned>(hw / 4);
    input_stage_vec4_kernel<3><<<ew_grid(1LL * n * hw4, kThreads), kThreads, 0, st>>>(
        img_u8, static_cast<unsigned>(n), hw4, np, out_nchw, mask_u8, mask_i64);
  } else {
    input_stage_kernel<<<ew_grid(n * hw, kThreads), kThreads, 0, st>>>(img_u8, n, hw, c, np, out_nchw, mask_u8, mask_i64);
  }
  CVB_LAUNCH_CHECK();
  return CVB_OK;
}
