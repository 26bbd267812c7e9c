// launch.cuh -- kernel launch with the programmatic-dependent-launch attribute.
#pragma once

#include <cuda_runtime.h>

namespace spmvb200 {

// Launch with cudaLaunchAttributeProgrammaticStreamSerialization when `pdl` is set: the kernel may
// start while its predecessor in the stream drains; it must execute griddepcontrol.wait before it
// touches anything the predecessor wrote.
template <typename... KArgs, typename... Args>
static cudaError_t launch_kernel(void (*kernel)(KArgs...), unsigned grid, unsigned block, size_t smem,
                                 cudaStream_t stream, bool pdl, Args... args)
{
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

}  // namespace spmvb200
