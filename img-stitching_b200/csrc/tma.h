// TMA (cp.async.bulk.tensor) helpers for the tile-staging kernels: host-side descriptor encoding through the driver
// entry point (the library links cudart statically and never links libcuda), device-side mbarrier + bulk-tensor PTX.
// Used by cubic5_kernel (frontend.cu) and warp_tile_kernel<.., kSrc4> (kernels.cu): both gather from word-per-pixel
// images whose tile footprint is a 2-D box, so the staged layout IS the source layout and the copy engine's
// out-of-bounds zero fill is exactly cv::remap's BORDER_CONSTANT(0) (include/nvcam.hpp:909).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>

namespace pano {

// 3-D tensor of 32-bit words: [images][height][width], rows `row_bytes` apart, images `img_bytes` apart; box = box_w x
// box_h words of one image, no swizzle, zero fill outside.  false: the driver entry point is unavailable or the
// geometry violates a TMA constraint (16-byte aligned base and strides, box_w <= 256) -- callers fall back to the
// LDG/STS staging loop.
bool tma_encode_words3d(CUtensorMap *map, const void *base, int width, int height, int images, size_t row_bytes,
                        size_t img_bytes, int box_w, int box_h);

#ifdef __CUDACC__
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");      // visible to the async proxy
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity)
{
    uint32_t done;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!done);
}

// one box of a 3-D tensor -> shared memory; completion is counted in bytes on `bar`
__device__ __forceinline__ void tma_load_3d(void *smem_dst, const CUtensorMap *map, int x, int y, int z, uint64_t *bar)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar))
                 : "memory");
}
#endif

}  // namespace pano
