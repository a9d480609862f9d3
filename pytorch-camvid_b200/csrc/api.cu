// Library-level plumbing of libcamvid_b200.so: error string, device queries, TMA tensor-map construction.
#include <stdarg.h>
#include <string.h>

#include <mutex>

#include "common.cuh"
#include "tma_host.h"

namespace cvb {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int sm_count() {
  static int cached = 0;
  if (cached > 0) return cached;
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) return 148;
  cached = n;
  return n;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

int make_act_tmap(CUtensorMap* out, const cvb_view& v, int box_w, int box_h, int box_n) {
  EncodeTiledFn fn = encode_fn();
  CVB_REQUIRE(fn != nullptr, CVB_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the CUDA driver");
  // channels need not fill the 64-wide box: what lies beyond v.c is out of bounds = zero-filled, at no HBM cost (the
  // 27-of-64 im2col'd first layer and the 12-of-64 output layer keep narrow tensors in memory this way)
  CVB_REQUIRE((v.c % 8) == 0, CVB_ERR_UNSUPPORTED, "GEMM operand view needs channels %% 8 == 0 (got %d)", v.c);
  CVB_REQUIRE(box_w >= 1 && box_w <= 256 && box_h >= 1 && box_h <= 256 && box_n >= 1 && box_n <= 256,
              CVB_ERR_INVALID_ARG, "bad TMA box %dx%dx%d", box_w, box_h, box_n);
  cuuint64_t dims[4] = {static_cast<cuuint64_t>(v.c), static_cast<cuuint64_t>(v.w), static_cast<cuuint64_t>(v.h),
                        static_cast<cuuint64_t>(v.n)};
  cuuint64_t strides[3] = {static_cast<cuuint64_t>(v.sw) * 2, static_cast<cuuint64_t>(v.sh) * 2,
                           static_cast<cuuint64_t>(v.sn) * 2};
  cuuint32_t box[4] = {64, static_cast<cuuint32_t>(box_w), static_cast<cuuint32_t>(box_h),
                       static_cast<cuuint32_t>(box_n)};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, v.ptr, dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CVB_REQUIRE(r == CUDA_SUCCESS, CVB_ERR_CUDA,
              "cuTensorMapEncodeTiled(activation %dx%dx%dx%d strides %lld,%lld,%lld box %d,%d,%d) failed: %d", v.n, v.h,
              v.w, v.c, (long long)v.sn, (long long)v.sh, (long long)v.sw, box_w, box_h, box_n, static_cast<int>(r));
  return CVB_OK;
}

int make_act_tmap_rowpairs(CUtensorMap* out, const cvb_view& v, int box_w, int box_pairs) {
  EncodeTiledFn fn = encode_fn();
  CVB_REQUIRE(fn != nullptr, CVB_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the CUDA driver");
  CVB_REQUIRE((v.c % 8) == 0 && (v.h % 2) == 0, CVB_ERR_UNSUPPORTED,
              "row-pair view needs channels %% 8 == 0 and an even height (got c %d, h %d)", v.c, v.h);
  cuuint64_t dims[5] = {static_cast<cuuint64_t>(v.c), static_cast<cuuint64_t>(v.w), 2,
                        static_cast<cuuint64_t>(v.h / 2), static_cast<cuuint64_t>(v.n)};
  cuuint64_t strides[4] = {static_cast<cuuint64_t>(v.sw) * 2, static_cast<cuuint64_t>(v.sh) * 2,
                           static_cast<cuuint64_t>(v.sh) * 4, static_cast<cuuint64_t>(v.sn) * 2};
  cuuint32_t box[5] = {64, static_cast<cuuint32_t>(box_w), 1, static_cast<cuuint32_t>(box_pairs), 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, v.ptr, dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CVB_REQUIRE(r == CUDA_SUCCESS, CVB_ERR_CUDA, "cuTensorMapEncodeTiled(row pairs %dx%dx%dx%d box %d,%d) failed: %d",
              v.n, v.h, v.w, v.c, box_w, box_pairs, static_cast<int>(r));
  return CVB_OK;
}

int make_mat_tmap(CUtensorMap* out, const void* ptr, long long rows, long long cols, int box_rows) {
  EncodeTiledFn fn = encode_fn();
  CVB_REQUIRE(fn != nullptr, CVB_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the CUDA driver");
  CVB_REQUIRE((cols % 64) == 0 && box_rows >= 1 && box_rows <= 256, CVB_ERR_INVALID_ARG,
              "bad matrix tensor map: %lld x %lld, box rows %d", rows, cols, box_rows);
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(cols) * 2};
  cuuint32_t box[2] = {64, static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CVB_REQUIRE(r == CUDA_SUCCESS, CVB_ERR_CUDA, "cuTensorMapEncodeTiled(matrix %lld x %lld) failed: %d", rows, cols,
              static_cast<int>(r));
  return CVB_OK;
}

}  // namespace cvb

extern "C" const char* cvb_last_error(void) { return cvb::g_err; }
extern "C" int cvb_abi_version(void) { return CVB_ABI_VERSION; }
extern "C" int cvb_sm_count(void) {
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess ||
      cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) {
    cvb::set_error("no CUDA device");
    return CVB_ERR_NO_DEVICE;
  }
  return n;
}
extern "C" int cvb_conv_stat_rows(void) { return cvb::sm_count(); }
extern "C" void cvb_shutdown(void) {}
