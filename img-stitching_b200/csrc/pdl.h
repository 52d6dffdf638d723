// Programmatic dependent launch for the per-frame kernel chain (sm_90+).
//
// Every kernel of the chain (front end -> warp -> pyrDown levels -> coarsest -> collapse levels) depends on the
// one before it, and at the top of the pyramid -- or with ONE frame-set in flight (pano_process, strip split) -- a
// kernel runs for a few microseconds: the launch + block-scheduling gap between two kernels is then as long as the
// kernels themselves.  Launched with cudaLaunchAttributeProgrammaticStreamSerialization a kernel may become resident while
// its predecessor is still draining; its first statement, pdl_enter(), lets ITS successor do the same
// (`griddepcontrol.launch_dependents`) and then blocks (`griddepcontrol.wait`) until the predecessor has completed and
// its writes are visible.  Nothing before pdl_enter() touches global memory and no block returns before it, so the
// data flow is exactly that of serialised launches (the wait is transitive: the predecessor waited for its own);
// what overlaps is launch latency, block dispatch and the tail of the previous grid.  Both instructions are no-ops
// in a kernel launched without the attribute.  Only kernels that call pdl_enter() may go through launch_chain().
// PANO_NO_PDL=1 launches the chain without the attribute (A/B measurements).
#pragma once
#include <cuda_runtime.h>

#include <cstdlib>
#include <utility>

namespace pano {

#ifdef __CUDACC__
__device__ __forceinline__ void pdl_enter()
{
    asm volatile("griddepcontrol.launch_dependents;");
    asm volatile("griddepcontrol.wait;" ::: "memory");
}
#endif

inline bool pdl_enabled()
{
    static const bool on = getenv("PANO_NO_PDL") == nullptr;
    return on;
}

// Largest grid (blocks) that is launched with the attribute: see the measurements in DESIGN.md section 4.
inline size_t pdl_max_blocks()
{
    static const size_t n = getenv("PANO_PDL_MAX_BLOCKS") ? (size_t)atoll(getenv("PANO_PDL_MAX_BLOCKS")) : (size_t)148 * 8;
    return n;
}

template <typename... P, typename... A>
inline cudaError_t launch_chain(void (*kern)(P...), dim3 grid, dim3 block, cudaStream_t stream, A &&...args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = 0;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = (pdl_enabled() && (size_t)grid.x * grid.y * grid.z <= pdl_max_blocks()) ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, std::forward<A>(args)...);
}

}  // namespace pano
