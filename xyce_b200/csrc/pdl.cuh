// xyce_b200 -- programmatic dependent launch for chains of small kernels on one stream.
// A kernel launched through launch_pdl may be scheduled while its predecessor in the stream is still draining
// (the launch latency and block scheduling overlap the predecessor's tail); it must execute pdl_wait() before it
// touches global memory: that instruction returns once the predecessor grid has completed and its writes are
// visible.  No kernel here signals early (griddepcontrol.launch_dependents), so ordering is exactly stream order.
// pdl_wait() in a kernel launched the ordinary way is a no-op.
#pragma once
#include <cuda_runtime.h>

namespace xb {

__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <class... KArgs, class... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

}  // namespace xb
