#include "host_utils.h"

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "emote_b200.h"

#ifdef EMOTE_OPERAND_BF16
#define EMOTE_TMAP_OP16 CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
#else
#define EMOTE_TMAP_OP16 CU_TENSOR_MAP_DATA_TYPE_FLOAT16
#endif

namespace emote {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_pdl(int enabled);
int set_error(const char* msg) {
  std::snprintf(g_err, sizeof(g_err), "%s", msg);
  return EMOTE_ERR_INVALID;
}
int set_error_cuda(const char* what, cudaError_t e) {
  std::snprintf(g_err, sizeof(g_err), "%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
  return EMOTE_ERR_CUDA;
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

static std::atomic<int> g_pdl{-1};
bool pdl_enabled() {
  int v = g_pdl.load(std::memory_order_relaxed);
  if (v < 0) {
    const char* e = std::getenv("EMOTE_PDL");
    v = (e && e[0] == '1') ? 1 : 0;  // opt-in: measured 2-3 % SLOWER on the UNet step graph (B200, driver 580)
    g_pdl.store(v, std::memory_order_relaxed);
  }
  return v != 0;
}
void set_pdl(int enabled) { g_pdl.store(enabled ? 1 : 0, std::memory_order_relaxed); }

// measurement knobs (emote_set_tuning): value < 0 = library default
static const char* const g_tuning_keys[TUNE_COUNT] = {"gn_reduce", "gn_apply_blocks", "ln_warps", "temporal_warps"};
static std::atomic<int> g_tuning[TUNE_COUNT] = {{-1}, {-1}, {-1}, {-1}};
int tuning(int id, int dflt) {
  int v = g_tuning[id].load(std::memory_order_relaxed);
  if (id == TUNE_GN_REDUCE && v < 0) {
    const char* e = std::getenv("EMOTE_GN_REDUCE");
    v = (e && e[0] == 's') ? 0 : 1;
    g_tuning[id].store(v, std::memory_order_relaxed);
  }
  return v < 0 ? dflt : v;
}
int set_tuning(const char* key, int value) {
  for (int i = 0; i < TUNE_COUNT; ++i)
    if (key && std::strcmp(key, g_tuning_keys[i]) == 0) {
      g_tuning[i].store(value, std::memory_order_relaxed);
      return 0;
    }
  return set_error("emote_set_tuning: unknown key");
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;
static std::once_flag g_encode_once;

static void resolve_encode() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  if (e == cudaSuccess && q == cudaDriverEntryPointSuccess) g_encode = reinterpret_cast<EncodeTiledFn>(fn);
}

int make_tensor_map(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                    const uint32_t* box, int swizzle, int elem_bytes) {
  std::call_once(g_encode_once, resolve_encode);
  if (!g_encode) return set_error("cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) return set_error("TMA base pointer must be 16-byte aligned");
  cuuint64_t gdims[5];
  cuuint64_t gstr[4];
  cuuint32_t gbox[5];
  cuuint32_t estr[5];
  for (int i = 0; i < rank; ++i) {
    gdims[i] = dims[i];
    gbox[i] = box[i];
    estr[i] = 1;
    if (box[i] == 0 || box[i] > 256) return set_error("TMA box extent out of range");
  }
  for (int i = 0; i + 1 < rank; ++i) {
    gstr[i] = strides_bytes[i];
    if (strides_bytes[i] % 16 != 0) return set_error("TMA global strides must be multiples of 16 bytes");
  }
  CUresult r = g_encode(out, elem_bytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : EMOTE_TMAP_OP16, (cuuint32_t)rank, const_cast<void*>(base), gdims, gstr,
                        gbox, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        swizzle == 0 ? CU_TENSOR_MAP_SWIZZLE_NONE : swizzle == 32 ? CU_TENSOR_MAP_SWIZZLE_32B
                        : swizzle == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[128];
    std::snprintf(buf, sizeof(buf), "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return set_error(buf);
  }
  return 0;
}

}  // namespace emote

extern "C" const char* emote_last_error(void) { return emote::g_err; }
extern "C" long long emote_launch_count(void) { return emote::g_launches.load(std::memory_order_relaxed); }
extern "C" int emote_abi_version(void) { return EMOTE_ABI_VERSION; }
extern "C" int emote_operand_dtype(void) {
#ifdef EMOTE_OPERAND_BF16
  return EMOTE_OP_BF16;
#else
  return EMOTE_OP_F16;
#endif
}
extern "C" void emote_set_pdl(int enabled) { emote::set_pdl(enabled); }
extern "C" int emote_set_tuning(const char* key, int32_t value) { return emote::set_tuning(key, value); }
