// Weight-stationary variant of the tcgen05 GEMM for the small-K (K <= 320), many-row bf16-output GEMMs at the
// 320-channel level: fused QKV (N = 960), cross-attention q (N = 320) and the GEGLU up-projection (N = 2560) over
// M = 131072 tokens (orig_attention.py:606-650, 825-827).
//
// With 128 x 160 tiles and K = 320 the streaming kernel (gemm_tcgen05.cu) moves 80 KB of A plus 100 KB of W per tile
// through L2 -> shared memory for only 1600 tensor cycles of work: those launches sit on the L2 -> SM delivery limit
// (~12 TB/s), not on the tensor pipe (profiles/r01_ncu_summary.md).  Here every CTA is pinned to ONE column tile: its
// whole 160 x K weight panel is TMA-loaded once and stays resident in shared memory (<= 100 KB); only the A row tiles
// stream through a 5-stage ring, which more than halves the operand traffic per tile.  CTAs that share a row tile
// (same slot, different column tile) run in lock step, so the A tile is fetched from HBM once and served from L2.
// TMEM double buffering and the staged TMA-store epilogue (OUT_MODE 1 of gemm_epilogue.cuh) are unchanged.
#include "common.cuh"
#include "gemm_epilogue.cuh"
#include "emote_b200.h"
#include "host_utils.h"

namespace emote {

template <int BN, int KB_MAX>
struct BresSmem {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int B_RES = KB_MAX * B_BYTES;
  static constexpr int STAGES = 5;
  static constexpr int OUT_BYTES = BM * BN * 2;
  static constexpr int TOTAL = B_RES + STAGES * A_BYTES + OUT_BYTES + 256 + 1024;
  static constexpr int TMEM_COLS = (2 * BN <= 256) ? 256 : 512;
};

template <int BN, int KB_MAX, int EPI_WARPS>
__global__ void __launch_bounds__(64 + 32 * EPI_WARPS, 1)
gemm_bres_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                         const __grid_constant__ CUtensorMap tmC, const GemmDev p, const int slots) {
  using S = BresSmem<BN, KB_MAX>;
  pdl_launch_early();
  extern __shared__ uint8_t smem_raw_bres[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw_bres) + 1023) & ~uintptr_t(1023));
  uint8_t* sB = smem;                         // resident weight panel: KB_MAX blocks of [BN rows][64 k] (128B swizzle)
  uint8_t* sA = smem + S::B_RES;              // A ring
  uint8_t* stage_out = sA + S::STAGES * S::A_BYTES;
  uint8_t* bar_base = stage_out + S::OUT_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(bar_base);
  uint64_t* empty_bar = full_bar + S::STAGES;
  uint64_t* tmem_full = empty_bar + S::STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint64_t* b_full = tmem_empty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(b_full + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmC);
    for (int s = 0; s < S::STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], EPI_WARPS);
    }
    mbar_init(b_full, 1);
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, S::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  // this CTA's column tile and its share of the row tiles: tm = slot, slot + slots, ...
  const int tn = blockIdx.x % p.tiles_n;
  const int slot = blockIdx.x / p.tiles_n;
  const int n0 = tn * BN;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      mbar_expect_tx(b_full, static_cast<uint32_t>(p.num_kb) * S::B_BYTES);
      for (int kb = 0; kb < p.num_kb; ++kb) tma_load_2d(sB + kb * S::B_BYTES, &tmB, b_full, kb * BK, n0);
    }
    __syncwarp();
    int stage = 0;
    uint32_t phase = 0;
    for (int tm = slot; tm < p.tiles_m; tm += slots) {
      const int m0 = tm * BM;
      for (int kb = 0; kb < p.num_kb; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (elect_one()) {
          mbar_expect_tx(&full_bar[stage], S::A_BYTES);
          tma_load_2d(sA + stage * S::A_BYTES, &tmA, &full_bar[stage], kb * BK, m0);
        }
        __syncwarp();
        if (++stage == S::STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
    pdl_launch_late();
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = umma_idesc_op16(BM, BN);
    int stage = 0;
    uint32_t phase = 0;
    int as = 0;
    uint32_t aphase = 0;
    mbar_wait(b_full, 0);
    tc_fence_after();
    for (int tm = slot; tm < p.tiles_m; tm += slots) {
      mbar_wait(&tmem_empty[as], aphase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(as * BN);
      for (int kb = 0; kb < p.num_kb; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t da = umma_desc_sw128(smem_u32(sA + stage * S::A_BYTES));
          const uint64_t db = umma_desc_sw128(smem_u32(sB + kb * S::B_BYTES));
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            umma_f16(d_tmem, da + static_cast<uint64_t>(2 * k), db + static_cast<uint64_t>(2 * k), idesc,
                     (kb | k) != 0 ? 1u : 0u);
          umma_commit(&empty_bar[stage]);
        }
        __syncwarp();
        if (++stage == S::STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
      if (elect_one()) umma_commit(&tmem_full[as]);
      __syncwarp();
      if (++as == 2) {
        as = 0;
        aphase ^= 1;
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps: TMEM -> bf16 tile in smem ->
    // one TMA bulk store per tile (same protocol as OUT_MODE 1 of gemm_tcgen05.cu)
    const int ew = warp - 2;
    const int quarter = warp & 3;
    const int part = ew >> 2;
    int as = 0;
    uint32_t aphase = 0;
    // GEGLU tiles are half as wide (BN/2 outputs): the staging area holds two of them, so tile i is written while the
    // bulk store of tile i-1 is still reading its buffer (only the store of tile i-2 has to be done)
    const bool two_buf = p.geglu != 0;
    int it = 0;
    for (int tm = slot; tm < p.tiles_m; tm += slots, ++it) {
      const uint32_t tbase = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(as * BN);
      uint8_t* stage_buf = stage_out + ((two_buf && (it & 1)) ? BM * (BN / 2) * 2 : 0);
      if (it > (two_buf ? 1 : 0)) {  // the bulk store that last used this buffer must have read it before it is rewritten
        if (threadIdx.x == 64) {
          if (two_buf) bulk_wait_read1(); else bulk_wait_read0();
        }
        named_bar_sync(2, EPI_WARPS * 32);
      }
      gemm_epilogue_tile<BN, EPI_WARPS, false, 1>(p, tbase, tm * BM, n0, tn, quarter, part, lane, stage_buf, [&]() {
        mbar_wait(&tmem_full[as], aphase);
        tc_fence_after();
      });
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[as]);
      fence_proxy_async_smem();
      named_bar_sync(1, EPI_WARPS * 32);
      if (threadIdx.x == 64) {
        store_bf16_boxes<BN>(&tmC, stage_buf, p, tn, tm * BM);
        bulk_commit();
      }
      if (++as == 2) {
        as = 0;
        aphase ^= 1;
      }
    }
    if (threadIdx.x == 64) bulk_wait0();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, S::TMEM_COLS);
  }
}

// Whether the weight-stationary kernel applies: bf16 staged-store epilogue, BN = 160, the whole K panel resident
// (<= 5 k-blocks), and enough row tiles per CTA to amortise the one-time panel load.
bool gemm_bres_applicable(int M, int N, int K, int bn, int num_sms) {
  if (bn != 160 || K > 5 * BK) return false;
  const int tiles_m = (M + BM - 1) / BM, tiles_n = (N + bn - 1) / bn;
  if (tiles_n > num_sms) return false;
  int slots = num_sms / tiles_n;
  if (slots > tiles_m) slots = tiles_m;
  return tiles_m >= 4 * slots;
}

int launch_gemm_bres(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, GemmDev& p, int num_sms,
                     cudaStream_t stream) {
  constexpr int BN = 160, KB_MAX = 5, EPI = 16;
  using S = BresSmem<BN, KB_MAX>;
  static PerDeviceOnce configured;
  int dev__ = 0;
  if (configured.pending(&dev__)) {
    cudaError_t e = cudaFuncSetAttribute(gemm_bres_tcgen05_kernel<BN, KB_MAX, EPI>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL);
    if (e != cudaSuccess) return set_error_cuda("cudaFuncSetAttribute(gemm_bres)", e);
    configured.done(dev__);
  }
  p.tiles_m = (p.M + BM - 1) / BM;
  p.tiles_n = (p.N + BN - 1) / BN;
  int slots = num_sms / p.tiles_n;
  if (slots > p.tiles_m) slots = p.tiles_m;
  launch_kernel(gemm_bres_tcgen05_kernel<BN, KB_MAX, EPI>, dim3(slots * p.tiles_n), dim3(64 + 32 * EPI), S::TOTAL, stream,
                tmA, tmB, tmC, p, slots);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error_cuda("gemm_bres launch", e);
  count_launch();
  return 0;
}

}  // namespace emote
