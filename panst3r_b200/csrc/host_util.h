// Host-side helpers shared by the .cu translation units: error plumbing for the C ABI and
// TMA tensor-map encoding through the driver entry point (no link-time dependency on libcuda).
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

namespace pst3r {

// Error codes returned across the C ABI (0 = success).
enum {
  PST3R_OK = 0,
  PST3R_ERR_INVALID = -1,   // bad argument / unsupported shape
  PST3R_ERR_CUDA = -2,      // CUDA runtime error (launch failure etc.)
  PST3R_ERR_DRIVER = -3,    // driver entry point / tensor map encode failure
  PST3R_ERR_NODEVICE = -4,  // no sm_100 device
};

void set_last_error(const char* fmt, ...);
const char* get_last_error();

#define PST3R_CHECK_ARG(cond, ...)          \
  do {                                      \
    if (!(cond)) {                          \
      pst3r::set_last_error(__VA_ARGS__);   \
      return pst3r::PST3R_ERR_INVALID;      \
    }                                       \
  } while (0)

#define PST3R_CHECK_CUDA(expr)                                                                  \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess) {                                                                    \
      pst3r::set_last_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                            __LINE__);                                                          \
      (void)cudaGetLastError(); /* reported through our return code: do not leave it for the next caller */ \
      return pst3r::PST3R_ERR_CUDA;                                                             \
    }                                                                                           \
  } while (0)

// Encode a bf16 (or any 2-byte) tiled tensor map with SWIZZLE_128B.
//   rank      : 2..5
//   dims[i]   : extent in elements, dims[0] is the contiguous dimension
//   strides[i]: stride in BYTES of dimension i (i >= 1); strides[0] is implied (elem size)
//   box[i]    : box extent in elements; box[0] * elem_size must be <= 128 for SWIZZLE_128B
// Returns 0 on success.
int encode_tmap(CUtensorMap* out, const void* base, int elem_bytes, int rank, const uint64_t* dims,
                const uint64_t* strides_bytes, const uint32_t* box, bool swizzle128 = true);

int num_sms();
// SMs the launchers may fill (<= num_sms()).  Persistent kernels size their grids with it and the tile / split
// heuristics count waves against it, so that two streams can share the GPU side by side (the latency-bound sequential
// memory build next to the throughput-bound DINOv2 encoder).  Set through pst3r_set_sm_budget(); 0 = all SMs.
int sm_budget();
void set_sm_budget(int n);

// Programmatic dependent launch (PDL): the kernel may start while its stream predecessor drains; every kernel
// launched this way executes griddepcontrol.wait (pdl_wait() in common.cuh) before touching global memory.
// PST3R_PDL=0 in the environment disables the attribute (plain stream order).
bool pdl_enabled();
int set_pdl(int on);  // returns the previous setting
bool split_k_enabled();
int set_split_k(int on);  // returns the previous setting

template <typename... KArgs, typename... Args>
cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

}  // namespace pst3r
