// Persistent warp-specialised tcgen05 GEMM / implicit-GEMM 3x3 convolution for sm_100a.
//
//   out[M,N] = epilogue( A[M,K] (bf16, K-major)  x  W[N,K]^T (bf16, K-major) )       fp32 accumulate in TMEM
//
// * operands staged by TMA (128B swizzle) into a multi-stage smem ring, consumed by tcgen05.mma
//   (UMMA 128 x BN x 16, cta_group::1) issued from a single thread; accumulators double-buffered in TMEM
//   so the epilogue of tile i overlaps the main loop of tile i+1.
// * conv mode: A is an NHWC bf16 image tensor; every 3x3 tap is one 4-D TMA box whose out-of-bounds
//   rows/columns are zero-filled by the TMA unit (= the conv zero padding), so no im2col is materialised.
//   Replaces cuDNN conv2d behind InflatedConv3d (reference magicanimate/models/resnet.py:30-38).
// * fused epilogues: bias, per-sample time-embedding bias (resnet.py:186-189), fp32 residual add and
//   1/output_scale_factor (resnet.py:202-205), GEGLU (orig_attention.py:817-827), bf16 or fp32 store.
#include "common.cuh"
#include "gemm_epilogue.cuh"
#include "emote_b200.h"
#include "host_utils.h"

namespace emote {

// warp0 = TMA producer, warp1 = MMA issuer + TMEM allocator, then EPI_WARPS epilogue warps.  Two variants:
//   HAS_ADD (residual and/or per-sample bias): 8 epilogue warps that prefetch their whole fp32 residual span into
//            registers BEFORE waiting for the accumulator, so ~80 KB/SM of residual reads overlap the main loop;
//   plain / GEGLU: 16 epilogue warps (4 per TMEM lane quarter) to hide TMEM and issue latency.
constexpr int gemm_threads(int epi_warps) { return 64 + 32 * epi_warps; }
// mode 2 (fp32 result staged for TMA stores, residual staged in by TMA) adds warp 2 = store warp: it bulk-stores each column
// half of a finished tile and re-arms that half of the staging buffer (next tile's residual boxes) while the epilogue warps
// work on the other half, so nobody waits for a whole-tile store -> reload round trip.
constexpr int gemm_threads_mode(int epi_warps, int out_mode) { return gemm_threads(epi_warps) + (out_mode == 2 ? 32 : 0); }

template <int BN, int OUT_MODE = 0>
struct GemmSmem {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = ((BN <= 128) ? 6 : (BN <= 160 ? 5 : 4)) - (OUT_MODE == 2 ? 1 : 0);
  // output tile staged for the TMA store: bf16 (mode 1) or fp32 residual-in / result-out boxes (mode 2)
  static constexpr int OUT_BYTES = OUT_MODE == 1 ? BM * BN * 2 : (OUT_MODE == 2 ? BM * BN * 4 : 0);
  static constexpr int BAR_BYTES = 256;
  static constexpr int TOTAL = STAGES * STAGE_BYTES + OUT_BYTES + BAR_BYTES + 1024 /*align slack*/;
  static constexpr int TMEM_COLS = (2 * BN <= 256) ? 256 : 512;
};

template <int BN, int EPI_WARPS, bool HAS_ADD, int OUT_MODE>
__global__ void __launch_bounds__(gemm_threads_mode(EPI_WARPS, OUT_MODE), 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                         const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmR,
                         const GemmDev p) {
  using S = GemmSmem<BN, OUT_MODE>;
  constexpr bool TMA_OUT = OUT_MODE != 0;
  constexpr bool SPLIT = OUT_MODE == 2;                 // store warp + column halves (see gemm_threads_mode)
  constexpr int FIRST_EPI = SPLIT ? 3 : 2;              // first epilogue warp
  constexpr int NBOX = BN / 32;                         // fp32 staging boxes of [128 rows][32 columns] per tile
  constexpr int NB_A = (NBOX + 1) / 2;                  // boxes in column half 0 (3 of 5 at BN = 160)
  pdl_launch_early();
  extern __shared__ uint8_t smem_raw[];
  // 128B-swizzled tiles need 1024 B alignment.
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stage_out = smem + S::STAGES * S::STAGE_BYTES;
  uint8_t* bar_base = smem + S::STAGES * S::STAGE_BYTES + S::OUT_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(bar_base);
  uint64_t* empty_bar = full_bar + S::STAGES;
  uint64_t* tmem_full = empty_bar + S::STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  uint64_t* res_full = reinterpret_cast<uint64_t*>(bar_base + 192);  // [2] mode 2: column half i of the staging buffer is armed
  uint64_t* half_done = reinterpret_cast<uint64_t*>(bar_base + 208); // [2] mode 2: the epilogue warps finished column half i

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < S::STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], EPI_WARPS);  // one arrive per epilogue warp
    }
    mbar_init(&res_full[0], 1);
    mbar_init(&res_full[1], 1);
    mbar_init(&half_done[0], EPI_WARPS);
    mbar_init(&half_done[1], EPI_WARPS);
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
  pdl_wait();  // barriers, TMEM and descriptors are set up; operands of earlier kernels may be read from here on

  const int num_tiles = p.tiles_m * p.tiles_n;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (warp-uniform loop, one
    // elected lane issues)
    {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int tm = tile / p.tiles_n;
        const int tn = tile - tm * p.tiles_n;
        const int m0 = tm * BM;
        const int n0 = tn * BN;
        int img0 = 0, y0 = 0, x0 = 0;
        if (p.taps > 1) {
          const TileOrigin o = tile_origin(p, m0);
          img0 = o.img0; y0 = o.y0; x0 = o.x0;
        }
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (elect_one()) {
            uint8_t* sa = smem + stage * S::STAGE_BYTES;
            uint8_t* sb = sa + S::A_BYTES;
            mbar_expect_tx(&full_bar[stage], S::STAGE_BYTES);
            if (p.taps > 1) {
              const int tap = kb / p.kb_per_tap;
              const int kc = kb - tap * p.kb_per_tap;
              const int dy = tap / 3 - 1;
              const int dx = tap - (tap / 3) * 3 - 1;
              tma_load_4d(sa, &tmA, &full_bar[stage], kc * BK, x0 + dx, y0 + dy, img0);
            } else {
              tma_load_2d(sa, &tmA, &full_bar[stage], kb * BK, m0);
            }
            tma_load_2d(sb, &tmB, &full_bar[stage], kb * BK, n0);
          }
          __syncwarp();
          if (++stage == S::STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
    pdl_launch_late();
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (warp-uniform loop, one
    // elected lane issues)
    {
      constexpr uint32_t idesc = umma_idesc_op16(BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&tmem_empty[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(as * BN);
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t sa = smem_u32(smem + stage * S::STAGE_BYTES);
            const uint32_t sb = sa + S::A_BYTES;
            const uint64_t da = umma_desc_sw128(sa);
            const uint64_t db = umma_desc_sw128(sb);
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              // advance 16 bf16 = 32 B inside the 128 B swizzle row: +2 in the (addr >> 4) field
              umma_f16(d_tmem, da + static_cast<uint64_t>(2 * k), db + static_cast<uint64_t>(2 * k), idesc,
                       (kb | k) != 0 ? 1u : 0u);
            }
            umma_commit(&empty_bar[stage]);  // frees the smem slot when these MMAs retire
          }
          __syncwarp();
          if (++stage == S::STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (elect_one()) umma_commit(&tmem_full[as]);  // accumulator ready for the epilogue
        __syncwarp();
        if (++as == 2) {
          as = 0;
          aphase ^= 1;
        }
      }
    }
  } else if (SPLIT && warp == 2) {
    // ------------------------------------------------------------------ mode 2: store warp (one lane)
    // Column half h of the staging buffer = boxes [h ? NB_A : 0, h ? NBOX : NB_A).  res_full[h] completes when half h is
    // armed for the next tile: its residual boxes have landed (arrive.expect_tx + TMA loads), or — no residual / no
    // valid box — the previous store has been read (plain arrive).  half_done[h] completes when all epilogue warps have
    // finished half h (their staging writes fenced for the async proxy).
    if (lane == 0) {
      const bool has_res = p.residual != nullptr;
      auto arm = [&](int tile, int h) {
        const int tm = tile / p.tiles_n;
        const int tn = tile - tm * p.tiles_n;
        const int b0 = h ? NB_A : 0, b1 = h ? NBOX : NB_A;
        int nb = 0;
        for (int b = b0; b < b1; ++b) nb += (tn * BN + b * 32 < p.N) ? 1 : 0;
        if (!has_res || nb == 0) {
          mbar_arrive(&res_full[h]);
          return;
        }
        mbar_expect_tx(&res_full[h], static_cast<uint32_t>(nb) * (BM * 128));
        for (int b = b0; b < b0 + nb; ++b)
          tile_box_load(p, stage_out + b * (BM * 128), &tmR, &res_full[h], tn * BN + b * 32, tm * BM);
      };
      auto prefetch = [&](int tile) {   // residual boxes of a later tile -> L2
        if (!has_res || tile >= num_tiles) return;
        const int ntm = tile / p.tiles_n, ntn = tile - ntm * p.tiles_n;
        for (int b = 0; b < NBOX; ++b)
          if (ntn * BN + b * 32 < p.N) tile_box_prefetch(p, &tmR, ntn * BN + b * 32, ntm * BM);
      };
      if (static_cast<int>(blockIdx.x) < num_tiles) {
        arm(blockIdx.x, 0);
        arm(blockIdx.x, 1);
      }
      uint32_t dphase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int tm = tile / p.tiles_n;
        const int tn = tile - tm * p.tiles_n;
        const int nxt = tile + static_cast<int>(gridDim.x);
        prefetch(nxt);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          mbar_wait(&half_done[h], dphase);
          const int b0 = h ? NB_A : 0, b1 = h ? NBOX : NB_A;
          for (int b = b0; b < b1; ++b)
            if (tn * BN + b * 32 < p.N) tile_box_store(p, &tmC, stage_out + b * (BM * 128), tn * BN + b * 32, tm * BM);
          bulk_commit();
          if (nxt < num_tiles) {
            bulk_wait_read0();   // the boxes of this half (and anything older) have been read: they may be overwritten
            arm(nxt, h);
          }
        }
        dphase ^= 1;
      }
      bulk_wait0();
    }
  } else if (SPLIT) {
    // ------------------------------------------------------------------ mode 2: epilogue warps, two column halves per tile
    const int quarter = warp & 3;
    const int part = (warp - FIRST_EPI) >> 2;
    constexpr int COLS_A = NB_A * 32;
    int as = 0;
    uint32_t aphase = 0, rphase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int tm = tile / p.tiles_n;
      const int tn = tile - tm * p.tiles_n;
      const uint32_t tbase = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(as * BN);
      mbar_wait(&res_full[0], rphase);
      gemm_epilogue_tile<BN, EPI_WARPS, HAS_ADD, 2, 0, COLS_A>(p, tbase, tm * BM, tn * BN, tn, quarter, part, lane, stage_out,
                                                               [&]() {
                                                                 mbar_wait(&tmem_full[as], aphase);
                                                                 tc_fence_after();
                                                               });
      fence_proxy_async_smem();   // staging writes -> visible to the TMA (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive(&half_done[0]);
      mbar_wait(&res_full[1], rphase);
      gemm_epilogue_tile<BN, EPI_WARPS, HAS_ADD, 2, COLS_A, BN - COLS_A>(p, tbase, tm * BM, tn * BN, tn, quarter, part, lane,
                                                                         stage_out, []() {});
      tc_fence_before();          // release the accumulator stage back to the MMA warp
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&tmem_empty[as]);
        mbar_arrive(&half_done[1]);
      }
      rphase ^= 1;
      if (++as == 2) {
        as = 0;
        aphase ^= 1;
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps (TMEM -> regs -> global)
    // EPI_WARPS warps: warp w may touch TMEM lanes 32*(w%4)..+31; the warps of a lane quarter take the 8-column
    // chunks of the tile round-robin.  tcgen05.ld.16x256b hands thread (g = lane/4, t = lane%4) the mma-style
    // fragment rows {g, g+8} x columns {2t, 2t+1}: a quad covers one full 32-byte sector of an fp32 row, so global
    // reads of the residual and writes of the output are sector-complete.  The loop is software pipelined: the TMEM
    // loads and the residual / bias loads of chunk i+1 are in flight while chunk i is finished and stored.
    const int ew = warp - 2;
    const int quarter = warp & 3;
    const int part = ew >> 2;
    int as = 0;
    uint32_t aphase = 0;
    // mode 1, GEGLU: tiles are half as wide (BN/2 outputs), the staging area holds two of them — tile i is written while
    // the bulk store of tile i-1 is still reading its buffer (only the store of tile i-2 has to be done)
    const bool two_buf = TMA_OUT && p.geglu != 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int tm = tile / p.tiles_n;
      const int tn = tile - tm * p.tiles_n;
      const uint32_t tbase = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(as * BN);
      uint8_t* stage_buf = stage_out + ((two_buf && (it & 1)) ? BM * (BN / 2) * 2 : 0);
      if constexpr (TMA_OUT) {
        if (it > (two_buf ? 1 : 0)) {  // the bulk store that last used this buffer must have read it before it is rewritten
          if (threadIdx.x == 64) {
            if (two_buf) bulk_wait_read1(); else bulk_wait_read0();
          }
          named_bar_sync(2, EPI_WARPS * 32);
        }
      }
      gemm_epilogue_tile<BN, EPI_WARPS, HAS_ADD, OUT_MODE>(p, tbase, tm * BM, tn * BN, tn, quarter, part, lane, stage_buf,
                                                           [&]() {
                                                             mbar_wait(&tmem_full[as], aphase);
                                                             tc_fence_after();
                                                           });
      // release the accumulator stage back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[as]);
      if constexpr (TMA_OUT) {   // mode 1: the staged 16-bit tile leaves through one bulk tensor store
        fence_proxy_async_smem();                 // staging writes -> visible to the TMA (async proxy)
        named_bar_sync(1, EPI_WARPS * 32);
        if (threadIdx.x == 64) {
          store_bf16_boxes<BN>(&tmC, stage_buf, p, tn, tm * BM);
          bulk_commit();
        }
      }
      if (++as == 2) {
        as = 0;
        aphase ^= 1;
      }
    }
    if constexpr (TMA_OUT) {
      if (threadIdx.x == 64) bulk_wait0();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, S::TMEM_COLS);
  }
}

// --------------------------------------------------------------------------- host side
int launch_gemm_pair(int bn, const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const CUtensorMap& tmR,
                     int out_mode, GemmDev& p, cudaStream_t stream);  // gemm2_tcgen05.cu
bool gemm_bres_applicable(int M, int N, int K, int bn, int num_sms);                                   // gemm_bres_tcgen05.cu
int launch_gemm_bres(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, GemmDev& p, int num_sms,
                     cudaStream_t stream);
bool gemm_skinny_applicable(const EmoteGemmArgs* a);                                                    // gemm_skinny.cu
int launch_gemm_skinny(const void* A, const void* Wt, void* out, const EmoteGemmArgs* a, cudaStream_t stream);
static int g_num_sms = 0;
static int num_sms() {
  if (g_num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  return g_num_sms;
}

template <int BN, int EPI_WARPS, bool HAS_ADD, int OUT_MODE>
static int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const CUtensorMap& tmR,
                       GemmDev& p, cudaStream_t stream) {
  using S = GemmSmem<BN, OUT_MODE>;
  static PerDeviceOnce configured;
  int dev__ = 0;
  if (configured.pending(&dev__)) {
    cudaError_t e = cudaFuncSetAttribute(gemm_bf16_tcgen05_kernel<BN, EPI_WARPS, HAS_ADD, OUT_MODE>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL);
    if (e != cudaSuccess) return set_error_cuda("cudaFuncSetAttribute(gemm)", e);
    configured.done(dev__);
  }
  p.tiles_m = p.patch ? p.subtiles : (p.M + BM - 1) / BM;
  p.tiles_n = (p.N + BN - 1) / BN;
  const int tiles = p.tiles_m * p.tiles_n;
  const int grid = tiles < num_sms() ? tiles : num_sms();
  launch_kernel(gemm_bf16_tcgen05_kernel<BN, EPI_WARPS, HAS_ADD, OUT_MODE>, dim3(grid), dim3(gemm_threads_mode(EPI_WARPS, OUT_MODE)), S::TOTAL, stream, tmA, tmB, tmC, tmR, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error_cuda("gemm launch", e);
  count_launch();
  return 0;
}

template <int BN>
static int dispatch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const CUtensorMap& tmR,
                         int out_mode, GemmDev& p, cudaStream_t stream) {
  if (out_mode >= 2) return launch_gemm<BN, 16, false, 2>(tmA, tmB, tmC, tmR, p, stream);
  if (!p.geglu && (p.residual != nullptr || p.row_bias != nullptr))
    return launch_gemm<BN, 8, true, 0>(tmA, tmB, tmC, tmR, p, stream);
  if (out_mode == 1) return launch_gemm<BN, 16, false, 1>(tmA, tmB, tmC, tmR, p, stream);
  return launch_gemm<BN, 16, false, 0>(tmA, tmB, tmC, tmR, p, stream);
}

}  // namespace emote

using namespace emote;

extern "C" int emote_gemm_bf16(const void* A, const void* Wt, void* out, const EmoteGemmArgs* a, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (!A || !Wt || !out || !a) return set_error("emote_gemm_bf16: null pointer");
  if (a->M <= 0 || a->N <= 0 || a->K <= 0) return set_error("emote_gemm_bf16: non-positive dims");
  if (a->K % 8 != 0) return set_error("emote_gemm_bf16: K must be a multiple of 8 (16-byte TMA rows)");
  const bool conv = a->conv_taps == 9;
  if (a->conv_taps != 1 && a->conv_taps != 9) return set_error("emote_gemm_bf16: conv_taps must be 1 or 9");
  const bool geglu = a->epilogue == EMOTE_EPI_GEGLU;
  const bool act_gelu = a->epilogue == EMOTE_EPI_GELU;
  if (a->epilogue != EMOTE_EPI_LINEAR && !geglu && !act_gelu) return set_error("emote_gemm_bf16: unknown epilogue");
  if (act_gelu && (a->row_bias || a->colstats)) return set_error("emote_gemm_bf16: the GELU epilogue takes no row_bias / colstats");
  if (geglu && (a->N % 2 != 0 || a->out_dtype != EMOTE_DT_OP16 || a->residual || a->row_bias))
    return set_error("emote_gemm_bf16: GEGLU epilogue needs even N, bf16 output, no residual/row_bias");
  const bool mode2_ok = a->tma_store != 2 && a->out_dtype == EMOTE_DT_F32 && !geglu &&
                        (!a->row_bias || (a->rows_per_group > 0 && a->rows_per_group % 128 == 0)) &&
                        (reinterpret_cast<uintptr_t>(out) & 15) == 0 && (reinterpret_cast<uintptr_t>(a->residual) & 15) == 0;
  const int bn = a->block_n ? a->block_n : ((a->N % 160 == 0) ? 160 : 128);
  if (bn != 128 && bn != 160) return set_error("emote_gemm_bf16: block_n must be 128 or 160");
  if (geglu && a->N % bn != 0) return set_error("emote_gemm_bf16: GEGLU needs N % block_n == 0");
  if ((a->out_dtype == EMOTE_DT_OP16 && a->ldc % 8 != 0) || (a->out_dtype == EMOTE_DT_F32 && a->ldc % 4 != 0))
    return set_error("emote_gemm_bf16: ldc must keep rows 16-byte aligned");
  if (a->residual && a->ldr % 4 != 0) return set_error("emote_gemm_bf16: ldr must be a multiple of 4");
  // M <= 8 (time-embedding products): weight-streaming GEMV instead of a 128-row tensor-core tile
  if (gemm_skinny_applicable(a)) return launch_gemm_skinny(A, Wt, out, a, stream);

  GemmDev p{};
  p.M = a->M; p.N = a->N; p.K = a->K;
  p.taps = a->conv_taps;
  p.bias = a->bias; p.row_bias = a->row_bias;
  p.rows_per_group = a->rows_per_group > 0 ? a->rows_per_group : 1;
  p.residual = a->residual; p.ldr = a->ldr;
  p.out_scale = a->out_scale;
  p.geglu = geglu ? 1 : 0;
  p.act_gelu = act_gelu ? 1 : 0;
  p.out_bf16 = a->out_dtype == EMOTE_DT_OP16;
  p.ldc = a->ldc;
  p.out = out;
  if (a->colstats) {
    if (p.out_bf16 || geglu) return set_error("emote_gemm_bf16: colstats needs a plain fp32 output");
    if (a->stats_rows <= 0 || a->stats_rows % 32 != 0 || a->M % a->stats_rows != 0 || a->N % 2 != 0 ||
        (reinterpret_cast<uintptr_t>(a->colstats) & 15) != 0)
      return set_error("emote_gemm_bf16: stats_rows must be a multiple of 32 that divides M (even N, 16-byte aligned slots)");
    p.colstats = a->colstats;
    p.stats_rows = a->stats_rows;
  }

  CUtensorMap tmA, tmB;
  bool patch_rowbias_per_row = false;   // patch mode with bias groups smaller than an image group: per-row bias (mode 0)
  if (conv) {
    const int C = a->C, H = a->H, W = a->W, NI = a->n_img;
    if (C <= 0 || C % 64 != 0) return set_error("emote_gemm_bf16(conv): C must be a multiple of 64");
    if (a->K != 9 * C) return set_error("emote_gemm_bf16(conv): K must equal 9*C");
    if ((long long)NI * H * W != a->M) return set_error("emote_gemm_bf16(conv): M must equal n_img*H*W");
    // 128 output pixels per sub-tile = one TMA box {64 channels, bw, bh, bnimg}: the widest power-of-two run of a row,
    // then rows, then images.  When the runs are whole rows (bw == W) or single-row (bh == 1) the 128 pixels are
    // consecutive output rows; otherwise (96x96, 48x48, 24x24, 12x12 ... maps) the sub-tile is a 2-D patch ("patch mode").
    auto gcd = [](int a, int b) { while (b) { const int t = a % b; a = b; b = t; } return a; };
    const int bw = gcd(W, 128);
    const int bh = gcd(H, 128 / bw);
    const int bnimg = 128 / (bw * bh);
    p.H = H; p.W = W; p.bw = bw; p.bh = bh; p.bnimg = bnimg; p.n_img = NI;
    const bool contiguous = bw == 128 || (bw == W && (bnimg == 1 || bh == H));
    p.patch = contiguous ? 0 : 1;
    if (p.patch) {
      p.tiles_x = W / bw;
      p.tiles_y = H / bh;
      p.subtiles = ((NI + bnimg - 1) / bnimg) * p.tiles_x * p.tiles_y;
      const long long group_rows = (long long)H * W * bnimg;   // a sub-tile never leaves its group of bnimg images
      if (p.colstats && p.stats_rows % group_rows != 0)
        return set_error("emote_gemm_bf16(conv): stats_rows must be a multiple of H*W*images-per-tile for this map size");
      if (a->row_bias && p.rows_per_group % group_rows != 0) patch_rowbias_per_row = true;
    }
    p.kb_per_tap = C / 64;
    p.num_kb = 9 * p.kb_per_tap;
    uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)NI};
    uint64_t strides[3] = {(uint64_t)C * 2, (uint64_t)W * C * 2, (uint64_t)H * W * C * 2};
    uint32_t box[4] = {64, (uint32_t)bw, (uint32_t)bh, (uint32_t)bnimg};
    if (int rc = make_tensor_map(&tmA, A, 4, dims, strides, box)) return rc;
  } else {
    // lda < K is allowed: rows then overlap in memory — the zero-copy operand of a strided 1-D convolution over a
    // [T, C] token matrix (row t = tokens stride*t .. stride*t + k - 1; lda = stride*C, K = k*C)
    if (a->lda % 8 != 0 || a->lda <= 0) return set_error("emote_gemm_bf16: lda must be a positive multiple of 8");
    p.kb_per_tap = (a->K + BK - 1) / BK;
    p.num_kb = p.kb_per_tap;
    uint64_t dims[2] = {(uint64_t)a->K, (uint64_t)a->M};
    uint64_t strides[1] = {(uint64_t)a->lda * 2};
    uint32_t box[2] = {64, 128};
    if (int rc = make_tensor_map(&tmA, A, 2, dims, strides, box)) return rc;
  }
  // CTA pairs (cta_group::2, UMMA M = 256, B tile split over the two SMs) for the tensor-bound shapes: the 3x3
  // convolutions and the large-K GEMMs, when there are enough 256-row tiles to occupy the 74 pairs.
  const long long pair_tiles = ((long long)(a->M + 255) / 256) * ((a->N + bn - 1) / bn);
  bool use_pair = (conv || a->K >= 1024) && a->M >= 256 && pair_tiles >= 64;
  if (p.patch && p.subtiles < 2) use_pair = false;
  if (a->pair_mode == 1) use_pair = true;
  if (a->pair_mode == 2) use_pair = false;
  {
    uint64_t dims[2] = {(uint64_t)a->K, (uint64_t)a->N};
    uint64_t strides[1] = {(uint64_t)a->K * 2};
    uint32_t box[2] = {64, (uint32_t)(use_pair ? bn / 2 : bn)};
    if (int rc = make_tensor_map(&tmB, Wt, 2, dims, strides, box)) return rc;
  }
  // Staged TMA epilogues (tma_store != 2):
  //   mode 1: bf16 outputs without residual adds (QKV / q projections, GEGLU) leave through one bulk tensor store;
  //   mode 2: fp32 outputs of the HBM/epilogue-bound shapes (K <= 4096; the deep-K convolutions keep their full operand
  //           ring): the fp32 residual tile is TMA-loaded into swizzled staging boxes, added in place and bulk-stored
  //           (per-sample row_bias is folded into the column bias when a tile never straddles a group).
  CUtensorMap tmC = tmB, tmR = tmB;
  int out_mode = 0;
  if (a->tma_store != 2) {
    if (p.out_bf16 && !(p.residual || p.row_bias)) {
      out_mode = p.patch ? 0 : 1;   // patch mode: 16-bit outputs leave through the register path
    } else if (mode2_ok && !patch_rowbias_per_row && a->K <= 4096) {
      out_mode = 2;
    }
  }
  if (out_mode == 1) {
    const int n_out = geglu ? a->N / 2 : a->N;
    uint64_t dims[2] = {(uint64_t)n_out, (uint64_t)a->M};
    uint64_t strides[1] = {(uint64_t)a->ldc * 2};
    uint32_t box[2] = {(uint32_t)(geglu ? bn / 2 : bn), 128};
    if (int rc = make_tensor_map(&tmC, out, 2, dims, strides, box, /*swizzle=*/0)) return rc;
  } else if (out_mode >= 2 && p.patch) {
    // the fp32 output / residual as [n_img, H, W, N] tensors: the staging boxes of a sub-tile are {32, bw, bh, bnimg}
    uint64_t dims[4] = {(uint64_t)a->N, (uint64_t)p.W, (uint64_t)p.H, (uint64_t)p.n_img};
    uint64_t strides[3] = {(uint64_t)a->ldc * 4, (uint64_t)p.W * a->ldc * 4, (uint64_t)p.H * p.W * a->ldc * 4};
    uint32_t box[4] = {32, (uint32_t)p.bw, (uint32_t)p.bh, (uint32_t)p.bnimg};
    if (int rc = make_tensor_map(&tmC, out, 4, dims, strides, box, true, 4)) return rc;
    tmR = tmC;
    if (p.residual) {
      uint64_t rs[3] = {(uint64_t)a->ldr * 4, (uint64_t)p.W * a->ldr * 4, (uint64_t)p.H * p.W * a->ldr * 4};
      if (int rc = make_tensor_map(&tmR, p.residual, 4, dims, rs, box, true, 4)) return rc;
    }
  } else if (out_mode >= 2) {
    uint64_t dims[2] = {(uint64_t)a->N, (uint64_t)a->M};
    uint64_t strides[1] = {(uint64_t)a->ldc * 4};
    uint32_t box[2] = {32, 128};
    if (int rc = make_tensor_map(&tmC, out, 2, dims, strides, box, true, 4)) return rc;
    tmR = tmC;
    if (p.residual) {
      strides[0] = (uint64_t)a->ldr * 4;
      if (int rc = make_tensor_map(&tmR, p.residual, 2, dims, strides, box, true, 4)) return rc;
    }
  }
  if (use_pair) return launch_gemm_pair(bn, tmA, tmB, tmC, tmR, out_mode, p, stream);
  // Weight-stationary kernel for the small-K, many-row bf16-output GEMMs (QKV / q / GEGLU at the 320-channel level),
  // which are L2 -> SM delivery bound in the streaming kernel (pair_mode == 3 disables it).
  if (!conv && out_mode == 1 && a->pair_mode != 3 && gemm_bres_applicable(a->M, a->N, a->K, bn, num_sms()))
    return launch_gemm_bres(tmA, tmB, tmC, p, num_sms(), stream);
  if (bn == 160) return dispatch_gemm<160>(tmA, tmB, tmC, tmR, out_mode, p, stream);
  return dispatch_gemm<128>(tmA, tmB, tmC, tmR, out_mode, p, stream);
}
