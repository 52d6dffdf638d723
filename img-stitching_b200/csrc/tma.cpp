// Host side of tma.h: cuTensorMapEncodeTiled reached through cudaGetDriverEntryPoint (no libcuda link dependency).
#include "tma.h"

#include <mutex>

namespace pano {

namespace {
using EncodeFn = CUresult (*)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                              const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                              CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeFn encodeFn()
{
    static EncodeFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeFn>(p);
        (void)cudaGetLastError();
    });
    return fn;
}
}  // namespace

bool tma_encode_words3d(CUtensorMap *map, const void *base, int width, int height, int images, size_t row_bytes, size_t img_bytes,
                        int box_w, int box_h)
{
    EncodeFn fn = encodeFn();
    if (!fn || !map || !base) return false;
    if ((reinterpret_cast<uintptr_t>(base) & 15) || (row_bytes & 15) || (img_bytes & 15)) return false;
    if (box_w < 1 || box_w > 256 || box_h < 1 || box_h > 256 || (box_w * 4) % 16) return false;
    const cuuint64_t dims[3] = {(cuuint64_t)width, (cuuint64_t)height, (cuuint64_t)(images > 0 ? images : 1)};
    const cuuint64_t strides[2] = {(cuuint64_t)row_bytes, (cuuint64_t)img_bytes};
    const cuuint32_t box[3] = {(cuuint32_t)box_w, (cuuint32_t)box_h, 1u};
    const cuuint32_t estr[3] = {1u, 1u, 1u};
    const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, const_cast<void *>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

}  // namespace pano
