// Host-side plumbing: thread-local error string and TMA descriptor encoding through the driver entry points
// (resolved at run time via cudart, so the shared library loads on machines without libcuda).
#include <stdarg.h>
#include <stdio.h>

#include "common.cuh"

namespace vb {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* get_error() { return g_err; }

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const int*, const int*, cuuint32_t, cuuint32_t, const cuuint32_t*,
                                   CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                   CUtensorMapFloatOOBfill);

static EncodeTiledFn g_tiled = nullptr;
static EncodeIm2colFn g_im2col = nullptr;

static int resolve() {
  if (g_tiled && g_im2col) return VB_OK;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  VB_CHECK_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  if (qres != cudaDriverEntryPointSuccess || !fn) {
    set_error("cuTensorMapEncodeTiled not available from the driver");
    return VB_ERR_CUDA;
  }
  g_tiled = reinterpret_cast<EncodeTiledFn>(fn);
  fn = nullptr;
  VB_CHECK_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &fn, cudaEnableDefault, &qres));
  if (qres != cudaDriverEntryPointSuccess || !fn) {
    set_error("cuTensorMapEncodeIm2col not available from the driver");
    return VB_ERR_CUDA;
  }
  g_im2col = reinterpret_cast<EncodeIm2colFn>(fn);
  return VB_OK;
}

static size_t dtype_bytes(CUtensorMapDataType t) {
  switch (t) {
    case CU_TENSOR_MAP_DATA_TYPE_BFLOAT16:
    case CU_TENSOR_MAP_DATA_TYPE_FLOAT16:
      return 2;
    case CU_TENSOR_MAP_DATA_TYPE_FLOAT32:
    case CU_TENSOR_MAP_DATA_TYPE_TFLOAT32:
      return 4;
    default:
      return 1;
  }
}

int encode_tma_2d(CUtensorMap* map, CUtensorMapDataType dtype, const void* base, uint64_t inner, uint64_t outer,
                  uint64_t outer_stride_bytes, uint32_t box_inner, uint32_t box_outer, CUtensorMapSwizzle swz) {
  int rc = resolve();
  if (rc) return rc;
  VB_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA: base pointer %p not 16-byte aligned", base);
  VB_REQUIRE(outer_stride_bytes % 16 == 0, "TMA: row stride %llu not a multiple of 16 bytes",
             (unsigned long long)outer_stride_bytes);
  VB_REQUIRE(box_inner * dtype_bytes(dtype) <= 128 || swz == CU_TENSOR_MAP_SWIZZLE_NONE,
             "TMA: inner box exceeds the swizzle span");
  VB_REQUIRE(box_outer <= 256 && box_inner <= 256, "TMA: box dims must be <= 256");
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {outer_stride_bytes};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_tiled(map, dtype, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d): inner=%llu outer=%llu stride=%llu box=%ux%u", (int)r,
              (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)outer_stride_bytes, box_inner,
              box_outer);
    return VB_ERR_CUDA;
  }
  return VB_OK;
}

int encode_tma_4d_nhwc(CUtensorMap* map, CUtensorMapDataType dtype, const void* base, int N, int H, int W, int C,
                       uint32_t box_c, uint32_t box_w, uint32_t box_h, CUtensorMapSwizzle swz) {
  int rc = resolve();
  if (rc) return rc;
  const size_t eb = dtype_bytes(dtype);
  VB_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA 4d: base pointer %p not 16-byte aligned", base);
  VB_REQUIRE(((size_t)C * eb) % 16 == 0, "TMA 4d: C*elem must be a multiple of 16 bytes (C=%d)", C);
  VB_REQUIRE(box_c * eb <= 128 && box_w <= 256 && box_h <= 256, "TMA 4d: box too large");
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)C * eb, (cuuint64_t)W * C * eb, (cuuint64_t)H * W * C * eb};
  cuuint32_t box[4] = {box_c, box_w, box_h, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = g_tiled(map, dtype, 4, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(4d) failed (%d): N=%d H=%d W=%d C=%d box=%ux%ux%u", (int)r, N, H, W, C, box_c,
              box_w, box_h);
    return VB_ERR_CUDA;
  }
  return VB_OK;
}

// NHWC activation tensor viewed by TMA im2col mode.  The "bounding box" corners follow the cuDNN/CUTLASS
// convention: lower = -pad_lo, upper = pad_hi - (filter - 1); base-pixel coordinates passed to the load are
// (q*stride - pad_lo_w, p*stride - pad_lo_h, n) and the filter tap is passed as the {s, r} offsets.
int encode_tma_im2col_nhwc(CUtensorMap* map, CUtensorMapDataType dtype, const void* base, int N, int H, int W, int C,
                           int pad_lo_h, int pad_lo_w, int pad_hi_h, int pad_hi_w, int R, int S, int stride,
                           uint32_t channels_per_pixel, uint32_t pixels_per_column, CUtensorMapSwizzle swz,
                           long long pixel_stride, long long row_stride, long long img_stride) {
  int rc = resolve();
  if (rc) return rc;
  const size_t eb = dtype_bytes(dtype);
  VB_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA im2col: base pointer %p not 16-byte aligned", base);
  VB_REQUIRE(((size_t)C * eb) % 16 == 0, "TMA im2col: C*elem must be a multiple of 16 bytes (C=%d)", C);
  VB_REQUIRE(pixels_per_column <= 1024 && channels_per_pixel <= 256, "TMA im2col: box too large");
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  // element strides between pixels / rows / images; a pixel stride below C gives overlapping pixel windows
  const cuuint64_t ps = pixel_stride > 0 ? (cuuint64_t)pixel_stride : (cuuint64_t)C;
  const cuuint64_t rs = row_stride > 0 ? (cuuint64_t)row_stride : (cuuint64_t)W * ps;
  const cuuint64_t is = img_stride > 0 ? (cuuint64_t)img_stride : (cuuint64_t)H * rs;
  VB_REQUIRE((ps * eb) % 16 == 0 && (rs * eb) % 16 == 0 && (is * eb) % 16 == 0,
             "TMA im2col: strides must be multiples of 16 bytes");
  cuuint64_t strides[3] = {ps * eb, rs * eb, is * eb};
  int lower[2] = {-pad_lo_w, -pad_lo_h};
  int upper[2] = {pad_hi_w - (S - 1), pad_hi_h - (R - 1)};
  cuuint32_t estr[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
  CUresult r = g_im2col(map, dtype, 4, const_cast<void*>(base), dims, strides, lower, upper, channels_per_pixel,
                        pixels_per_column, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeIm2col failed (%d): N=%d H=%d W=%d C=%d R=%d S=%d stride=%d pad=(%d,%d,%d,%d)", (int)r,
              N, H, W, C, R, S, stride, pad_lo_h, pad_lo_w, pad_hi_h, pad_hi_w);
    return VB_ERR_CUDA;
  }
  // Driver workaround mirrored from CUTLASS (copy_traits_sm90_im2col.hpp): for small tensors (< 128 KiB)
  // drivers <= 13.1 set a descriptor bit that makes im2col loads fault; clear it.
  int drv = 0;
  if (cudaDriverGetVersion(&drv) == cudaSuccess && drv <= 13010) {
    const size_t bytes = (size_t)N * is * eb;
    if (bytes < 131072) reinterpret_cast<uint64_t*>(map)[1] &= ~(1ull << 21);
  }
  return VB_OK;
}

}  // namespace vb
