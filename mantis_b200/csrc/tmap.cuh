// Host-side TMA tensor-map construction shared by the tcgen05 kernels.  cuTensorMapEncodeTiled is resolved through
// cudaGetDriverEntryPoint so that the library has no link-time dependency on libcuda (it must load on a CPU-only box).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <errno.h>

extern "C" void mb200_set_last_error(const char* msg);

namespace mbtmap {

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static inline PFN_encodeTiled get_encode() {
  static const PFN_encodeTiled fn = [] {
    void* p = nullptr; cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      return (PFN_encodeTiled)p;
    return (PFN_encodeTiled) nullptr;
  }();
  return fn;
}

// 2-D bf16 map over a row-major [rows, cols] matrix with leading dimension ld (elements); 128B swizzle, zero OOB fill.
static inline int make_2d(CUtensorMap* tm, const void* base, long long rows, long long cols, long long ld, int box_cols,
                          int box_rows) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) { mb200_set_last_error("cuTensorMapEncodeTiled unavailable"); return -ENOSYS; }
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { mb200_set_last_error("cuTensorMapEncodeTiled (2-D) failed"); return -EINVAL; }
  return 0;
}

// 2-D map used as a TMA-STORE / TMA-REDUCE destination: bf16 (box_cols x box_rows elements) or fp32
static inline int make_2d_store(CUtensorMap* tm, void* base, long long rows, long long cols, long long ld, int box_cols,
                                int box_rows, bool f32 = false) {
  if (!f32) return make_2d(tm, base, rows, cols, ld, box_cols, box_rows);
  PFN_encodeTiled enc = get_encode();
  if (!enc) { mb200_set_last_error("cuTensorMapEncodeTiled unavailable"); return -ENOSYS; }
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { mb200_set_last_error("cuTensorMapEncodeTiled (fp32 store) failed"); return -EINVAL; }
  return 0;
}

// 3-D bf16 map over x[slabs][rows][cols] (cols contiguous, dense): box (box_cols, box_rows, 1), 128B swizzle
static inline int make_3d(CUtensorMap* tm, const void* base, long long slabs, long long rows, long long cols, int box_cols,
                          int box_rows) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) { mb200_set_last_error("cuTensorMapEncodeTiled unavailable"); return -ENOSYS; }
  cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)slabs};
  cuuint64_t strides[2] = {(cuuint64_t)cols * 2, (cuuint64_t)cols * rows * 2};
  cuuint32_t box[3] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { mb200_set_last_error("cuTensorMapEncodeTiled (3-D) failed"); return -EINVAL; }
  return 0;
}

// 4-D bf16 map over x[B, S, H, hd] (element strides sb, ss, sh; hd contiguous): dims (hd, H, S, B), box (64, 1, rows, 1)
static inline int make_bshd(CUtensorMap* tm, const void* base, int B, int S, int H, int hd, long long sb, long long ss,
                            long long sh, int box_rows) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) { mb200_set_last_error("cuTensorMapEncodeTiled unavailable"); return -ENOSYS; }
  cuuint64_t dims[4] = {(cuuint64_t)hd, (cuuint64_t)H, (cuuint64_t)S, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)sh * 2, (cuuint64_t)ss * 2, (cuuint64_t)sb * 2};
  cuuint32_t box[4] = {64, 1, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { mb200_set_last_error("cuTensorMapEncodeTiled (4-D) failed"); return -EINVAL; }
  return 0;
}

}  // namespace mbtmap
