// Host-side CUtensorMap construction. libcuda is not linked at build time (the
// build box has no driver); cuTensorMapEncodeTiled is resolved at run time
// through cudaGetDriverEntryPoint.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace vpd {

// Encodes a bf16 tiled tensor map with 128-byte swizzle (or none).
// dims/strides are innermost-first; strides_bytes[0] is implied (2 bytes) and
// ignored. Returns 0 on success, else sets the library error string.
int encode_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                     const uint64_t* strides_bytes, const uint32_t* box, bool swizzle128);

void set_error(const char* fmt, ...);
void count_launches(int n);
long long launch_count();
const char* get_error();

#define VPD_CHECK_CUDA(expr)                                                        \
  do {                                                                              \
    cudaError_t _e = (expr);                                                        \
    if (_e != cudaSuccess) {                                                        \
      ::vpd::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),      \
                       __FILE__, __LINE__);                                         \
      return -1;                                                                    \
    }                                                                               \
  } while (0)

// after a <<<>>> launch: bump the launch counter and check for launch errors
#define VPD_LAUNCHED(n)                       \
  do {                                        \
    ::vpd::count_launches(n);                 \
    VPD_CHECK_CUDA(cudaGetLastError());       \
  } while (0)

bool pdl_enabled();

// Kernel launch with the programmatic-stream-serialization attribute (PDL) and an
// optional thread-block cluster of `cluster` CTAs along x.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel_cluster(int cluster, void (*kernel)(KArgs...), dim3 grid,
                                         dim3 block, size_t smem, cudaStream_t stream,
                                         Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int n = 0;
  if (pdl_enabled()) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  if (cluster > 1) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = cluster;
    attr[n].val.clusterDim.y = 1;
    attr[n].val.clusterDim.z = 1;
    ++n;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                                 cudaStream_t stream, Args&&... args) {
  return launch_kernel_cluster(1, kernel, grid, block, smem, stream, static_cast<Args&&>(args)...);
}

#define VPD_REQUIRE(cond, ...)          \
  do {                                  \
    if (!(cond)) {                      \
      ::vpd::set_error(__VA_ARGS__);    \
      return -1;                        \
    }                                   \
  } while (0)

}  // namespace vpd
