// Host-side launch helper: kernel launches that allow programmatic dependent launch (see common.cuh, pdl_wait).
#pragma once

#include <cuda_runtime.h>

#include <utility>

namespace ddl {

extern int g_pdl;        // set_pdl / DDL_PDL: 0 = plain stream order, 1 (default) = opted-in kernels are launched with
                         // programmaticStreamSerializationAllowed, 2 / 3 = plus explicit triggers (common.cuh)

// `cluster` > 1 launches thread-block clusters of that many CTAs along x.  Only kernels that execute pdl_wait() before
// their first global-memory access may be launched through this helper.
template <typename... P, typename... A>
cudaError_t launch_pdl(void (*kernel)(P...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, int cluster,
                       A&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int n = 0;
  if (cluster > 1) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = static_cast<unsigned>(cluster);
    attr[n].val.clusterDim.y = 1;
    attr[n].val.clusterDim.z = 1;
    ++n;
  }
  if (g_pdl) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  cfg.attrs = attr;
  cfg.numAttrs = static_cast<unsigned>(n);
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<A>(args)...);
}

}  // namespace ddl
