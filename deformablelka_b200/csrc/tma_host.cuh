// tma_host.cuh -- host-side tensor-map construction for channels-last activations [B][D][H][W][C] (fp32).
// The driver's encoder is resolved through cudaGetDriverEntryPoint, so the library does not link libcuda.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

namespace dlka {

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_tiled_fn()
{
    static EncodeTiledFn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
            qres != cudaDriverEntryPointSuccess)
            p = nullptr;
        return (EncodeTiledFn)p;
    }();
    return fn;
}

// box = (bc channels, bw, bh, bd voxels, 1 sample); traversal stride `dil` along w and h (box extents are given in
// loaded voxels); elements outside the volume are zero-filled by the hardware.  Returns false on failure.
inline bool make_tmap_cl5(CUtensorMap *tm, const float *x, int B, int C, int D, int H, int W, int bc, int bw, int bh, int bd, int dil)
{
    EncodeTiledFn encode = encode_tiled_fn();
    if (!encode) return false;
    const cuuint64_t gdim[5] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)B};
    const cuuint64_t gstr[4] = {(cuuint64_t)C * 4, (cuuint64_t)W * C * 4, (cuuint64_t)H * W * C * 4, (cuuint64_t)D * H * W * C * 4};
    const cuuint32_t box[5] = {(cuuint32_t)bc, (cuuint32_t)(bw * dil), (cuuint32_t)(bh * dil), (cuuint32_t)bd, 1};
    const cuuint32_t estr[5] = {1, (cuuint32_t)dil, (cuuint32_t)dil, 1, 1};
    return encode(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<float *>(x), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace dlka
