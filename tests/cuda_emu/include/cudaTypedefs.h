// tests/cuda_emu -- stand-in for <cudaTypedefs.h>: the tensor-map descriptor of the TMA engine and its encoder
// (TEST INFRASTRUCTURE ONLY).  The emulated descriptor keeps what a 2-D tiled FP64 copy needs.
#pragma once
#include <stdint.h>

typedef uint64_t cuuint64_t;
typedef uint32_t cuuint32_t;
typedef int CUresult;
enum { CUDA_SUCCESS = 0, CUDA_ERROR_INVALID_VALUE = 1 };
enum CUtensorMapDataType { CU_TENSOR_MAP_DATA_TYPE_FLOAT64 = 10 };
enum CUtensorMapInterleave { CU_TENSOR_MAP_INTERLEAVE_NONE = 0 };
enum CUtensorMapSwizzle { CU_TENSOR_MAP_SWIZZLE_NONE = 0, CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_SWIZZLE_128B };
enum CUtensorMapL2promotion { CU_TENSOR_MAP_L2_PROMOTION_NONE = 0, CU_TENSOR_MAP_L2_PROMOTION_L2_256B = 3 };
enum CUtensorMapFloatOOBfill { CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE = 0 };

struct alignas(64) CUtensorMap {
    const unsigned char* base;
    uint64_t dim[2];       // elements: {inner, outer}
    uint64_t row_stride;   // bytes between outer indices
    uint32_t box[2];       // elements: {inner, outer}
    uint32_t swizzle;
    uint32_t elem_bytes;
    uint64_t pad[9];
};
static_assert(sizeof(CUtensorMap) == 128, "descriptor size of the real type");

typedef CUresult (*PFN_cuTensorMapEncodeTiled_v12000)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                                      const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                                      CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                                      CUtensorMapFloatOOBfill);
