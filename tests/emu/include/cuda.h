// cuda.h of the CPU kernel-logic harness (tests/emu): the tensor-map types of the driver API, with
// a transparent descriptor instead of the opaque hardware one.  TEST INFRASTRUCTURE ONLY.
#pragma once

#include <cstdint>

typedef uint32_t cuuint32_t;
typedef uint64_t cuuint64_t;
typedef int CUresult;
enum { CUDA_SUCCESS = 0, CUDA_ERROR_INVALID_VALUE = 1 };
enum CUtensorMapDataType { CU_TENSOR_MAP_DATA_TYPE_FLOAT64 = 10 };
enum CUtensorMapInterleave { CU_TENSOR_MAP_INTERLEAVE_NONE = 0 };
enum CUtensorMapSwizzle {
    CU_TENSOR_MAP_SWIZZLE_NONE = 0,
    CU_TENSOR_MAP_SWIZZLE_32B,
    CU_TENSOR_MAP_SWIZZLE_64B,
    CU_TENSOR_MAP_SWIZZLE_128B
};
enum CUtensorMapL2promotion { CU_TENSOR_MAP_L2_PROMOTION_NONE = 0, CU_TENSOR_MAP_L2_PROMOTION_L2_128B = 2 };
enum CUtensorMapFloatOOBfill { CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE = 0 };

struct alignas(64) CUtensorMap_st {
    // what cuTensorMapEncodeTiled was told (fp64 elements)
    double *base;
    int rank;
    int swizzle;
    uint64_t dims[3];
    uint64_t stride_bytes[3];   // stride_bytes[0] = 8
    uint32_t box[3];
    uint32_t pad_[5];
};
typedef CUtensorMap_st CUtensorMap;
static_assert(sizeof(CUtensorMap) == 128, "keep the size of the real descriptor");
