// Host-side helpers shared by the C-ABI entry points: last-error string, launch counter,
// TMA tensor-map encoding through the driver entry point (no link-time libcuda dependency, so the
// library also loads on a box without a driver for the symbol-export check).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace emote {

int set_error(const char* msg);
int set_error_cuda(const char* what, cudaError_t e);
void count_launch(int n = 1);

// rank-2..5 bf16 (elem_bytes 2) or fp32 (elem_bytes 4) tensor map, 128B swizzle by default, zero OOB fill. dims/box innermost-first;
// strides (bytes) for dims 1..rank-1.
int make_tensor_map(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                    const uint32_t* box, bool swizzle128 = true, int elem_bytes = 2);

#define EMOTE_CHECK_LAUNCH(name)                                   \
  do {                                                             \
    cudaError_t e__ = cudaGetLastError();                          \
    if (e__ != cudaSuccess) return emote::set_error_cuda(name, e__); \
    emote::count_launch();                                         \
  } while (0)

}  // namespace emote
