// Host-side TMA descriptor (CUtensorMap) creation without linking libcuda: the driver entry point is
// resolved at run time through the runtime API, so the extension imports on a machine without a GPU.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <stdexcept>
#include <string>

namespace tma {

using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    if (e != cudaSuccess || p == nullptr || q != cudaDriverEntryPointSuccess)
      throw std::runtime_error("cuTensorMapEncodeTiled not available from the driver");
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// cuTensorMapEncodeTiled is a DRIVER call and needs a current context on the calling thread.  Threads that never made a
// runtime call for their device (an autograd / pipeline worker whose device is the default device 0: the c10 device guard
// skips cudaSetDevice when the index does not change) have none and get CUDA_ERROR_INVALID_CONTEXT (201): bind the primary
// context with a runtime call and encode again.
template <typename... A>
inline CUresult encode_with_context(A... a) {
  CUresult r = get_encode_fn()(a...);
  if (r == CUDA_ERROR_INVALID_CONTEXT) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaSetDevice(dev);
    cudaFree(nullptr);
    r = get_encode_fn()(a...);
  }
  return r;
}

// 2-D row-major tensor [rows, cols] with `row_pitch_bytes`; box = {box_cols, box_rows}; 128B swizzle when
// box_cols * elem_size == 128, otherwise no swizzle.
inline CUtensorMap make_2d(const void* base, CUtensorMapDataType dt, int elem_bytes, uint64_t rows, uint64_t cols,
                           uint64_t row_pitch_bytes, uint32_t box_rows, uint32_t box_cols, bool swizzle128 = true) {
  CUtensorMap m;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {row_pitch_bytes};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = encode_with_context(&m, dt, (cuuint32_t)2, const_cast<void*>(base), (const cuuint64_t*)dims, (const cuuint64_t*)strides,
                                   (const cuuint32_t*)box, (const cuuint32_t*)estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE,
                               swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    throw std::runtime_error("cuTensorMapEncodeTiled failed with code " + std::to_string(static_cast<int>(r)) +
                             " (rows=" + std::to_string(rows) + " cols=" + std::to_string(cols) +
                             " pitch=" + std::to_string(row_pitch_bytes) + ")");
  return m;
}

// 3-D [d2, d1, d0] tensor with byte strides for d1/d2 (d0 contiguous); box {b0, b1, b2}
inline CUtensorMap make_3d(const void* base, CUtensorMapDataType dt, uint64_t d0, uint64_t d1, uint64_t d2,
                           uint64_t stride1_bytes, uint64_t stride2_bytes, uint32_t b0, uint32_t b1, uint32_t b2,
                           bool swizzle128 = true) {
  CUtensorMap m;
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {stride1_bytes, stride2_bytes};
  cuuint32_t box[3] = {b0, b1, b2};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = encode_with_context(&m, dt, (cuuint32_t)3, const_cast<void*>(base), (const cuuint64_t*)dims, (const cuuint64_t*)strides,
                                   (const cuuint32_t*)box, (const cuuint32_t*)estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE,
                               swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    throw std::runtime_error("cuTensorMapEncodeTiled(3d) failed with code " + std::to_string(static_cast<int>(r)));
  return m;
}

}  // namespace tma
