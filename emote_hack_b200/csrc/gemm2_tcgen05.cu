// CTA-pair (cta_group::2) variant of the tcgen05 GEMM / implicit-GEMM conv for the tensor-bound shapes.
//
// Two CTAs of a cluster (one TPC) cooperate on a 256 x BN output tile: each CTA TMA-loads its own 128 rows of A and
// HALF of the B tile (BN/2 weight rows) and owns a 128 x BN accumulator in its TMEM; the leader CTA issues
// tcgen05.mma.cta_group::2 (UMMA M = 256) which reads A from each CTA's shared memory and the two B halves from both.
// Per CTA and k-block this moves (16 KB + BN/2 * 128 B) through shared memory instead of (16 KB + BN * 128 B): the
// single-CTA kernel is shared-memory-bandwidth bound on the large-K convolutions (profiles/r01_ncu_summary.md:
// tensor pipe 55 %, 72 KB of smem traffic per 320 MMA cycles).
// Everything else (persistent tile loop, TMEM double buffering, fused epilogues) follows gemm_tcgen05.cu.
#include "common.cuh"
#include "gemm_epilogue.cuh"
#include "emote_b200.h"
#include "host_utils.h"

namespace emote {

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `p` (a local shared-memory pointer) inside CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_shared(const void* p, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(p)), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish2() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma2_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      :
      : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive (once the MMAs issued so far have retired) on the barrier at the same offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma2_commit_multicast(uint64_t* bar) {
  const uint16_t mask = 0x3;
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(mask)
      : "memory");
}
// TMA loads of a CTA pair: data lands in the executing CTA's smem, completion bytes are credited to the LEADER's barrier
__device__ __forceinline__ void tma2_load_2d(void* smem_dst, const void* desc, uint32_t leader_bar, int32_t c0, int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], "
      "[%2];"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(desc)), "r"(leader_bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma2_load_4d(void* smem_dst, const void* desc, uint32_t leader_bar, int32_t c0, int32_t c1,
                                             int32_t c2, int32_t c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, "
      "%6}], [%2];"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(desc)), "r"(leader_bar), "r"(c0), "r"(c1), "r"(c2),
        "r"(c3)
      : "memory");
}

template <int BN, int OUT_MODE = 0>
struct Gemm2Smem {
  static constexpr int A_BYTES = BM * BK * 2;            // 16 KB: this CTA's 128 rows
  static constexpr int B_BYTES = (BN / 2) * BK * 2;      // this CTA's half of the B tile
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = ((BN <= 128) ? 8 : 7) - (OUT_MODE == 1 ? 1 : (OUT_MODE == 2 ? ((BN <= 128) ? 2 : 2) : 0));
  static constexpr int OUT_BYTES = OUT_MODE == 1 ? BM * BN * 2 : (OUT_MODE == 2 ? BM * BN * 4 : 0);
  static constexpr int TOTAL = STAGES * STAGE_BYTES + OUT_BYTES + 256 + 1024;
  static constexpr int TMEM_COLS = (2 * BN <= 256) ? 256 : 512;
};

template <int BN, int EPI_WARPS, bool HAS_ADD, int OUT_MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(64 + 32 * EPI_WARPS + (OUT_MODE == 2 ? 32 : 0), 1)
gemm2_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                          const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmR,
                          const GemmDev p) {
  using S = Gemm2Smem<BN, OUT_MODE>;
  constexpr bool TMA_OUT = OUT_MODE != 0;
  constexpr bool SPLIT = OUT_MODE == 2;     // store warp (warp 2) + column halves: see gemm_tcgen05.cu
  constexpr int FIRST_EPI = SPLIT ? 3 : 2;
  constexpr int NBOX = BN / 32;
  constexpr int NB_A = (NBOX + 1) / 2;
  pdl_launch_early();
  extern __shared__ uint8_t smem_raw2[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw2) + 1023) & ~uintptr_t(1023));
  uint8_t* stage_out = smem + S::STAGES * S::STAGE_BYTES;
  uint8_t* bar_base = smem + S::STAGES * S::STAGE_BYTES + S::OUT_BYTES;
  uint64_t* res_full = reinterpret_cast<uint64_t*>(bar_base + 192);  // [2] per CTA, mode 2: column half i of the staging buffer is armed
  uint64_t* half_done = reinterpret_cast<uint64_t*>(bar_base + 208); // [2] per CTA, mode 2: the epilogue warps finished column half i
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(bar_base);   // used in the leader: bytes of BOTH CTAs
  uint64_t* empty_bar = full_bar + S::STAGES;                    // per CTA, multicast commit
  uint64_t* tmem_full = empty_bar + S::STAGES;                   // per CTA, multicast commit
  uint64_t* tmem_empty = tmem_full + 2;                          // used in the leader: epilogue warps of both CTAs
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < S::STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], 2 * EPI_WARPS);
    }
    mbar_init(&res_full[0], 1);
    mbar_init(&res_full[1], 1);
    mbar_init(&half_done[0], EPI_WARPS);
    mbar_init(&half_done[1], EPI_WARPS);
    mbar_fence_init();
  }
  cluster_sync_all();  // barriers of both CTAs are initialised before anyone signals across the pair
  if (warp == 1) {
    tmem_alloc2(tmem_slot, S::TMEM_COLS);
    tmem_relinquish2();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();  // set-up done; operands of earlier kernels may be read from here on

  const int num_tiles = p.tiles_m * p.tiles_n;   // tiles_m counts 256-row pair tiles
  const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (one per CTA; warp-uniform
    // loop, one elected lane issues)
    {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
        const int tm = tile / p.tiles_n;
        const int tn = tile - tm * p.tiles_n;
        const int m0 = tm * (2 * BM) + static_cast<int>(rank) * BM;
        const int nb0 = tn * BN + static_cast<int>(rank) * (BN / 2);
        int img0 = 0, y0 = 0, x0 = 0;
        if (p.taps > 1) {
          const TileOrigin o = tile_origin(p, m0);   // a sub-tile past the last one maps beyond n_img: zero-filled loads
          img0 = o.img0; y0 = o.y0; x0 = o.x0;
        }
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (elect_one()) {
            uint8_t* sa = smem + stage * S::STAGE_BYTES;
            uint8_t* sb = sa + S::A_BYTES;
            const uint32_t lbar = mapa_shared(&full_bar[stage], 0);
            if (leader) mbar_expect_tx(&full_bar[stage], 2 * S::STAGE_BYTES);
            if (p.taps > 1) {
              const int tap = kb / p.kb_per_tap;
              const int kc = kb - tap * p.kb_per_tap;
              const int dy = tap / 3 - 1;
              const int dx = tap - (tap / 3) * 3 - 1;
              tma2_load_4d(sa, &tmA, lbar, kc * BK, x0 + dx, y0 + dy, img0);
            } else {
              tma2_load_2d(sa, &tmA, lbar, kb * BK, m0);
            }
            tma2_load_2d(sb, &tmB, lbar, kb * BK, nb0);
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
    // ------------------------------------------------------------------ MMA issuer: one elected lane of the LEADER
    // CTA's warp 1 (warp-uniform loop)
    if (leader) {
      constexpr uint32_t idesc = umma_idesc_op16(2 * BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
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
            for (int k = 0; k < BK / 16; ++k)
              umma2_f16(d_tmem, da + static_cast<uint64_t>(2 * k), db + static_cast<uint64_t>(2 * k), idesc,
                        (kb | k) != 0 ? 1u : 0u);
            umma2_commit_multicast(&empty_bar[stage]);  // frees the slot in both CTAs
          }
          __syncwarp();
          if (++stage == S::STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (elect_one()) umma2_commit_multicast(&tmem_full[as]);  // both CTAs' epilogues may drain their half
        __syncwarp();
        if (++as == 2) {
          as = 0;
          aphase ^= 1;
        }
      }
    }
  } else if (SPLIT && warp == 2) {
    // ------------------------------------------------------------------ mode 2: store warp of this CTA (one lane)
    if (lane == 0) {
      const bool has_res = p.residual != nullptr;
      auto row_base_of = [&](int tile) { return (tile / p.tiles_n) * (2 * BM) + static_cast<int>(rank) * BM; };
      auto arm = [&](int tile, int h) {
        const int tn = tile % p.tiles_n;
        const int b0 = h ? NB_A : 0, b1 = h ? NBOX : NB_A;
        int nb = 0;
        for (int b = b0; b < b1; ++b) nb += (tn * BN + b * 32 < p.N) ? 1 : 0;
        if (!has_res || nb == 0) {
          mbar_arrive(&res_full[h]);
          return;
        }
        mbar_expect_tx(&res_full[h], static_cast<uint32_t>(nb) * (BM * 128));
        for (int b = b0; b < b0 + nb; ++b)
          tile_box_load(p, stage_out + b * (BM * 128), &tmR, &res_full[h], tn * BN + b * 32, row_base_of(tile));
      };
      auto prefetch = [&](int tile) {
        if (!has_res || tile >= num_tiles) return;
        const int ntn = tile % p.tiles_n;
        for (int b = 0; b < NBOX; ++b)
          if (ntn * BN + b * 32 < p.N) tile_box_prefetch(p, &tmR, ntn * BN + b * 32, row_base_of(tile));
      };
      if (pair < num_tiles) {
        arm(pair, 0);
        arm(pair, 1);
      }
      uint32_t dphase = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
        const int tn = tile % p.tiles_n;
        const int row_base = row_base_of(tile);
        const bool rows_ok = p.patch ? (row_base / BM < p.subtiles) : (row_base < p.M);   // the second CTA of the last pair may own no valid rows
        const int nxt = tile + num_pairs;
        prefetch(nxt);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          mbar_wait(&half_done[h], dphase);
          const int b0 = h ? NB_A : 0, b1 = h ? NBOX : NB_A;
          for (int b = b0; b < b1; ++b)
            if (rows_ok && tn * BN + b * 32 < p.N)
              tile_box_store(p, &tmC, stage_out + b * (BM * 128), tn * BN + b * 32, row_base);
          bulk_commit();
          if (nxt < num_tiles) {
            bulk_wait_read0();
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
    for (int tile = pair; tile < num_tiles; tile += num_pairs) {
      const int tm = tile / p.tiles_n;
      const int tn = tile - tm * p.tiles_n;
      const int row_base = tm * (2 * BM) + static_cast<int>(rank) * BM;
      const uint32_t tbase = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(as * BN);
      mbar_wait(&res_full[0], rphase);
      gemm_epilogue_tile<BN, EPI_WARPS, HAS_ADD, 2, 0, COLS_A>(p, tbase, row_base, tn * BN, tn, quarter, part, lane, stage_out,
                                                               [&]() {
                                                                 mbar_wait(&tmem_full[as], aphase);
                                                                 tc_fence_after();
                                                               });
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&half_done[0]);
      mbar_wait(&res_full[1], rphase);
      gemm_epilogue_tile<BN, EPI_WARPS, HAS_ADD, 2, COLS_A, BN - COLS_A>(p, tbase, row_base, tn * BN, tn, quarter, part, lane,
                                                                         stage_out, []() {});
      tc_fence_before();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive_cluster(mapa_shared(&tmem_empty[as], 0));
        mbar_arrive(&half_done[1]);
      }
      rphase ^= 1;
      if (++as == 2) {
        as = 0;
        aphase ^= 1;
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps (each CTA drains its 128 rows)
    const int ew = warp - 2;
    const int quarter = warp & 3;
    const int part = ew >> 2;
    int as = 0;
    uint32_t aphase = 0;
    const bool two_buf = TMA_OUT && p.geglu != 0;   // GEGLU: two half-width staging buffers (see gemm_tcgen05.cu)
    int it = 0;
    for (int tile = pair; tile < num_tiles; tile += num_pairs, ++it) {
      const int tm = tile / p.tiles_n;
      const int tn = tile - tm * p.tiles_n;
      const int row_base = tm * (2 * BM) + static_cast<int>(rank) * BM;
      const uint32_t tbase = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(as * BN);
      uint8_t* stage_buf = stage_out + ((two_buf && (it & 1)) ? BM * (BN / 2) * 2 : 0);
      if constexpr (TMA_OUT) {
        if (it > (two_buf ? 1 : 0)) {   // the bulk store that last used this buffer must have read it before it is rewritten
          if (threadIdx.x == 64) {
            if (two_buf) bulk_wait_read1(); else bulk_wait_read0();
          }
          named_bar_sync(2, EPI_WARPS * 32);
        }
      }
      gemm_epilogue_tile<BN, EPI_WARPS, HAS_ADD, OUT_MODE>(p, tbase, row_base, tn * BN, tn, quarter, part, lane, stage_buf,
                                                           [&]() {
                                                             mbar_wait(&tmem_full[as], aphase);
                                                             tc_fence_after();
                                                           });
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_shared(&tmem_empty[as], 0));
      if constexpr (TMA_OUT) {   // mode 1: the staged 16-bit tile leaves through one bulk tensor store
        fence_proxy_async_smem();
        named_bar_sync(1, EPI_WARPS * 32);
        if (threadIdx.x == 64) {
          // the second CTA of the last pair may own no valid rows
          if (p.patch ? (row_base / BM < p.subtiles) : (row_base < p.M)) store_bf16_boxes<BN>(&tmC, stage_buf, p, tn, row_base);
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
  cluster_sync_all();  // nobody leaves (or frees TMEM) while the peer may still signal / read
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, S::TMEM_COLS);
  }
}

static int g2_num_sms = 0;

template <int BN, int EPI_WARPS, bool HAS_ADD, int OUT_MODE>
static int launch_gemm2(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const CUtensorMap& tmR,
                        GemmDev& p, cudaStream_t stream) {
  using S = Gemm2Smem<BN, OUT_MODE>;
  static PerDeviceOnce configured;
  int dev__ = 0;
  if (configured.pending(&dev__)) {
    cudaError_t e = cudaFuncSetAttribute(gemm2_bf16_tcgen05_kernel<BN, EPI_WARPS, HAS_ADD, OUT_MODE>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL);
    if (e != cudaSuccess) return set_error_cuda("cudaFuncSetAttribute(gemm2)", e);
    configured.done(dev__);
  }
  p.tiles_m = p.patch ? (p.subtiles + 1) / 2 : (p.M + 2 * BM - 1) / (2 * BM);
  p.tiles_n = (p.N + BN - 1) / BN;
  const int tiles = p.tiles_m * p.tiles_n;
  if (g2_num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g2_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g2_num_sms <= 0) g2_num_sms = 148;
  }
  const int max_pairs = g2_num_sms / 2;
  const int pairs = tiles < max_pairs ? tiles : max_pairs;
  launch_kernel(gemm2_bf16_tcgen05_kernel<BN, EPI_WARPS, HAS_ADD, OUT_MODE>, dim3(2 * pairs), dim3(64 + 32 * EPI_WARPS + (OUT_MODE == 2 ? 32 : 0)), S::TOTAL, stream, tmA, tmB, tmC, tmR, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error_cuda("gemm2 launch", e);
  count_launch();
  return 0;
}

// entry used by emote_gemm_bf16 (gemm_tcgen05.cu) for the shapes routed to CTA pairs
int launch_gemm_pair(int bn, const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const CUtensorMap& tmR,
                     int out_mode, GemmDev& p, cudaStream_t stream) {
  const bool has_add = !p.geglu && (p.residual != nullptr || p.row_bias != nullptr);
  if (bn == 160) {
    if (out_mode == 2) return launch_gemm2<160, 16, false, 2>(tmA, tmB, tmC, tmR, p, stream);
    if (has_add) return launch_gemm2<160, 8, true, 0>(tmA, tmB, tmC, tmR, p, stream);
    if (out_mode == 1) return launch_gemm2<160, 16, false, 1>(tmA, tmB, tmC, tmR, p, stream);
    return launch_gemm2<160, 16, false, 0>(tmA, tmB, tmC, tmR, p, stream);
  }
  if (out_mode == 2) return launch_gemm2<128, 16, false, 2>(tmA, tmB, tmC, tmR, p, stream);
  if (has_add) return launch_gemm2<128, 8, true, 0>(tmA, tmB, tmC, tmR, p, stream);
  if (out_mode == 1) return launch_gemm2<128, 16, false, 1>(tmA, tmB, tmC, tmR, p, stream);
  return launch_gemm2<128, 16, false, 0>(tmA, tmB, tmC, tmR, p, stream);
}

}  // namespace emote
