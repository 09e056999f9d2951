// Host-side CUtensorMap construction through the driver entry point (no link-time dependency on libcuda).
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>

#include "common.cuh"

namespace rd {

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_encodeTiled tensormap_encoder() {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_encodeTiled>(p);
    }
    return fn;
}

// 4-D map over a [B, S, H, 64] 16-bit tensor (element strides sb, ss, sh; innermost contiguous), dims ordered
// {D, S, H, B}; box = {64, box_rows, 1, 1}, SWIZZLE_128B (one 128-byte row per token), OOB rows read as zero.
inline int make_qkv_map(CUtensorMap* m, const void* base, rd_dtype dt, int B, int H, int S, int64_t sb, int64_t ss,
                        int64_t sh, int box_rows) {
    PFN_encodeTiled enc = tensormap_encoder();
    if (!enc) { set_error("cuTensorMapEncodeTiled not available from the driver"); return RD_ERR_CUDA; }
    const cuuint64_t dims[4] = {64, (cuuint64_t)S, (cuuint64_t)H, (cuuint64_t)B};
    if (B == 1) sb = (int64_t)S * ss;   // unused stride; keep it legal
    if (H == 1) sh = 64;
    const cuuint64_t strides[3] = {(cuuint64_t)ss * 2, (cuuint64_t)sh * 2, (cuuint64_t)sb * 2};
    const cuuint32_t box[4] = {64, (cuuint32_t)box_rows, 1, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(m, dt == RD_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4,
                     const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r); return RD_ERR_CUDA; }
    return RD_OK;
}

}  // namespace rd
