// Host-side helpers shared by the C-ABI entry points: last-error string, launch counter,
// TMA tensor-map encoding through the driver entry point (no link-time libcuda dependency, so the
// library also loads on a box without a driver for the symbol-export check).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <utility>

namespace emote {

int set_error(const char* msg);
int set_error_cuda(const char* what, cudaError_t e);
void count_launch(int n = 1);

// rank-2..5 bf16 (elem_bytes 2) or fp32 (elem_bytes 4) tensor map, zero OOB fill. dims/box innermost-first; strides (bytes)
// for dims 1..rank-1.  swizzle: 0 = none, 32 / 64 / 128 = that TMA swizzle span in bytes (1 is accepted as 128).
int make_tensor_map(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                    const uint32_t* box, int swizzle = 128, int elem_bytes = 2);

// All kernels are launched through here: cudaLaunchKernelEx, with programmatic stream serialization (PDL) when
// enabled (EMOTE_PDL=1 or emote_set_pdl(1); off by default — it measured slower on the UNet step graph); the kernels
// gate their global-memory accesses with griddepcontrol.wait, a no-op for ordinary launches.
bool pdl_enabled();
// emote_set_tuning knobs: launch-geometry / variant choices that can be switched inside one process for A/B timing.
// tuning(id, dflt) returns the library default until a value >= 0 was set.
enum { TUNE_GN_REDUCE = 0, TUNE_GN_APPLY_BLOCKS, TUNE_LN_WARPS, TUNE_TEMPORAL_WARPS, TUNE_COUNT };
int tuning(int id, int dflt);
int set_tuning(const char* key, int value);
template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                 Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

// cudaFuncSetAttribute applies to the device that is current when it is called: a launcher remembers, per device
// ordinal, whether it has raised its kernel's dynamic shared-memory limit there.  Two threads racing through pending()
// both configure — the call is idempotent.
struct PerDeviceOnce {
  std::atomic<uint64_t> mask{0};
  bool pending(int* dev) {
    if (cudaGetDevice(dev) != cudaSuccess) *dev = 0;
    return !((mask.load(std::memory_order_acquire) >> (*dev & 63)) & 1ull);
  }
  void done(int dev) { mask.fetch_or(1ull << (dev & 63), std::memory_order_release); }
};

#define EMOTE_CHECK_LAUNCH(name)                                   \
  do {                                                             \
    cudaError_t e__ = cudaGetLastError();                          \
    if (e__ != cudaSuccess) return emote::set_error_cuda(name, e__); \
    emote::count_launch();                                         \
  } while (0)

}  // namespace emote
