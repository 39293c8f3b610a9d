// tcgen05 flash attention for the spatial self-attention / reference-attention layers (head_dim 40 / 80 / 160) and the
// wav2vec2 encoder (head_dim 64): >80% of the attention time of a UNet call is spent here (N = 4096 / 1024 keys per image).
//
//   per CTA: 128 query rows of one (image, head).  Per 64-key tile:
//     S  = Q K^T          tcgen05.mma 128 x 64 x 16 (x KSTEPS), Q/K K-major 128B-swizzled smem, S in TMEM (fp32)
//     P  = exp2(S*c - m)  4 softmax warps, one thread per query row (TMEM lane == row: no shuffles), P -> smem (16-bit)
//     O += P V            tcgen05.mma 128 x NPV x 16 (x4) ACCUMULATING in TMEM, A = P (K-major), B = V consumed MN-major
//                         straight from its natural [key][d] layout (no transpose); row sums from a P x ones MMA
//   one thread of warp 0 streams K/V tiles with TMA into a ring (128B swizzle; zero-fills the tail), warp 1 issues the
//   MMAs, warps 2-5 do the softmax.  S and P are double buffered so the softmax of tile j overlaps S(j+1) and P V(j-1).
//
// Same semantics as flash_attn_kernel in attention.cu (two key/value segments, per-batch visibility of segment 1,
// ragged tails); reference arithmetic: orig_attention.py:655-684, mutual_self_attention.py:239-255.
#include "attention_tc.cuh"

#include <cstdlib>

namespace emote {

// ---------------------------------------------------------------------------------------------------------------------
// O and the row sums accumulate in TMEM (PV / P x ones MMAs with the accumulate flag) and are rescaled LAZILY.
//
// The round-1 kernel folded every tile's P V into per-thread registers (O = O*corr + PV: a TMEM read of the whole PV tile
// plus D FMAs per row and tile, 48-80 live registers, 153 registers per thread).  Timing showed one CTA needs ~1600 cycles
// per 64-key tile whatever else runs on the SM: the softmax warp's dependent chain (S load -> max -> 64 exponentials ->
// P store -> P V fold) bounds it, not MUFU (48 % busy) or the tensor pipe (27 %).  Here the softmax thread only produces
// P (120 registers; head_dim 40, N = 4096, 32 x 8 heads: 1.62 -> 1.51 ms; ncu: issue slots 60 %, XU 34 %, tensor 29 %,
// ~450 issued instructions per warp and tile — the kernel is issue-bound on the exponentials' scalar work).  It keeps the
// exponent reference m_used of its row and moves it only when the running maximum has grown by more than 2^8 (log2 units);
// then — rarely after the first tiles — the thread rescales its own O row and row sum in TMEM (tcgen05.ld -> multiply ->
// tcgen05.st) before publishing P.  P is therefore bounded by 2^8 instead of 1 (fp32 accumulation; exact after the final
// division by the equally scaled row sum).  S and P stay double-buffered so S(j+1) is computed while softmax(j) runs.
template <int D>
struct TcCfg {
  static constexpr int ATOMS = (D + 63) / 64;
  static constexpr int KSTEPS = (D + 15) / 16;
  static constexpr int NPV = KSTEPS * 16;
  static constexpr int CH = D / 8;
  static constexpr int STAGES = (D <= 96) ? 3 : 2;       // K/V ring (head_dim 160: 2 stages of 48 KB fit next to Q and P)
  static constexpr int Q_BYTES = ATOMS * TC_BQ * 128;
  static constexpr int KV_TILE = ATOMS * TC_BKV * 128;
  static constexpr int STAGE = 2 * KV_TILE;
  static constexpr int P_BYTES = TC_BQ * 128;
  static constexpr int ONES_BYTES = 16 * 128;
  static constexpr int SMEM = Q_BYTES + STAGES * STAGE + 2 * P_BYTES + ONES_BYTES + 256 + 1024;
  static constexpr int O_COL0 = 2 * TC_BKV;             // S double buffer occupies columns [0, 128)
  static constexpr int SUM_COL0 = O_COL0 + NPV;         // 16 columns: every one holds the row sum (column 0 is used)
  static constexpr int USED_COLS = SUM_COL0 + 16;
  static constexpr int TMEM_COLS = USED_COLS <= 256 ? 256 : 512;
  static constexpr int CTAS = (2 * SMEM <= 226 * 1024 && TMEM_COLS <= 256) ? 2 : 1;
};

template <int D, int EMU>
__global__ void __launch_bounds__(TC_THREADS, TcCfg<D>::CTAS)
flash_attn_tc_kernel(const __grid_constant__ CUtensorMap tmK0, const __grid_constant__ CUtensorMap tmV0,
                      const __grid_constant__ CUtensorMap tmK1, const __grid_constant__ CUtensorMap tmV1,
                      const AttnTcDev p) {
  using C = TcCfg<D>;
  pdl_launch_early();
  extern __shared__ uint8_t tc_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(tc_smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sKV = sQ + C::Q_BYTES;
  uint8_t* sP = sKV + C::STAGES * C::STAGE;
  uint8_t* sOnes = sP + 2 * C::P_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sOnes + C::ONES_BYTES);
  uint64_t* kv_full = bars;                   // [STAGES] TMA (expect_tx) -> MMA
  uint64_t* kv_empty = kv_full + C::STAGES;   // [STAGES] MMA commit -> loader
  uint64_t* s_full = kv_empty + C::STAGES;    // [2] MMA commit -> softmax: S of tile j in buffer j&1
  uint64_t* p_full = s_full + 2;              // [2] softmax (4 warp arrivals) -> MMA: P of tile j stored, S buffer j&1 free
  uint64_t* o_done = p_full + 2;              // [2] MMA commit -> softmax: P V of tile j accumulated, P buffer j&1 free
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_done + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.z, h = blockIdx.y;
  const int q0 = blockIdx.x * TC_BQ;

  const int n1 = (p.n1 > 0 && b >= p.kv1_first) ? p.n1 : 0;
  const int tiles0 = (p.n0 + TC_BKV - 1) / TC_BKV;
  const int tiles1 = (n1 + TC_BKV - 1) / TC_BKV;
  const int ntiles = tiles0 + tiles1;

  {
    uint4* z = reinterpret_cast<uint4*>(smem);
    const int n16 = (C::Q_BYTES + C::STAGES * C::STAGE + 2 * C::P_BYTES) / 16;
    for (int i = threadIdx.x; i < n16; i += TC_THREADS) z[i] = make_uint4(0u, 0u, 0u, 0u);
    uint4* o = reinterpret_cast<uint4*>(sOnes);
    for (int i = threadIdx.x; i < C::ONES_BYTES / 16; i += TC_THREADS)
      o[i] = make_uint4(OP16_ONE_PAIR, OP16_ONE_PAIR, OP16_ONE_PAIR, OP16_ONE_PAIR);
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&s_full[s], 1);
      mbar_init(&p_full[s], 4);
      mbar_init(&o_done[s], 1);
    }
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, C::TMEM_COLS);
    tmem_relinquish();
  }
  __syncthreads();
  pdl_wait();
  {
    const op16* qg = p.q + (long long)b * p.q_bs + (long long)q0 * p.q_rs + h * p.d;
    const int nvq = p.nq - q0;
    for (int i = threadIdx.x; i < TC_BQ * C::CH; i += TC_THREADS) {
      const int r = i / C::CH, c = i - r * C::CH;
      const uint32_t dst = smem_u32(sQ) + (c >> 3) * (TC_BQ * 128) + r * 128 + (((c & 7) ^ (r & 7)) << 4);
      cp_async16_zfill(dst, qg + (long long)r * p.q_rs + c * 8, r < nvq);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ K/V loader: one thread, TMA
    // Each K / V tile is 64 keys x (ATOMS x 64) columns starting at column h*D of the [batch][key][heads*D] view: the
    // box over-reads up to 64-D%64 columns of the NEXT head (or zero-fill past the row end).  Harmless: the matching Q
    // columns are zero (QK^T) and the extra P*V columns are never stored.  Rows past the segment end are zero-filled.
    for (int j = 0; j < ntiles; ++j) {
      const int stage = j % C::STAGES;
      if (j >= C::STAGES) mbar_wait(&kv_empty[stage], ((j / C::STAGES) - 1) & 1);
      if (elect_one()) {
        uint8_t* kdst = sKV + stage * C::STAGE;
        uint8_t* vdst = kdst + C::KV_TILE;
        mbar_expect_tx(&kv_full[stage], C::STAGE);
        const bool seg0 = j < tiles0;
        const CUtensorMap* mk = seg0 ? &tmK0 : &tmK1;
        const CUtensorMap* mv = seg0 ? &tmV0 : &tmV1;
        const int row = (seg0 ? j : j - tiles0) * TC_BKV;
        const int bidx = seg0 ? b / p.kv0_div : b / p.kv1_div;
#pragma unroll
        for (int a = 0; a < C::ATOMS; ++a) {
          tma_load_3d(kdst + a * (TC_BKV * 128), mk, &kv_full[stage], h * D + a * 64, row, bidx);
          tma_load_3d(vdst + a * (TC_BKV * 128), mv, &kv_full[stage], h * D + a * 64, row, bidx);
        }
      }
      __syncwarp();
    }
    pdl_launch_late();
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc_s = umma_idesc_op16(TC_BQ, TC_BKV);
    constexpr uint32_t idesc_pv = umma_idesc_op16_bmn(TC_BQ, C::NPV);
    constexpr uint32_t idesc_sum = umma_idesc_op16(TC_BQ, 16);
    auto issue_s = [&](int j) {
      const int stage = j % C::STAGES;
      mbar_wait(&kv_full[stage], (j / C::STAGES) & 1);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t ka = smem_u32(sKV + stage * C::STAGE);
        const uint32_t qa = smem_u32(sQ);
        const uint32_t d_tmem = tmem_base + (j & 1) * TC_BKV;
#pragma unroll
        for (int ks = 0; ks < C::KSTEPS; ++ks) {
          const uint64_t da = umma_desc_sw128(qa + (ks >> 2) * (TC_BQ * 128)) + static_cast<uint64_t>(2 * (ks & 3));
          const uint64_t db = umma_desc_sw128(ka + (ks >> 2) * (TC_BKV * 128)) + static_cast<uint64_t>(2 * (ks & 3));
          umma_f16(d_tmem, da, db, idesc_s, ks != 0 ? 1u : 0u);
        }
        umma_commit(&s_full[j & 1]);
      }
      __syncwarp();
    };
    issue_s(0);
    for (int j = 0; j < ntiles; ++j) {
      if (j + 1 < ntiles) issue_s(j + 1);   // S buffer (j+1)&1 was released by p_full(j-1), waited last iteration
      mbar_wait(&p_full[j & 1], (j >> 1) & 1);   // P(j) is in shared memory and O was rescaled if it had to be
      tc_fence_after();
      if (elect_one()) {
        const int stage = j % C::STAGES;
        const uint32_t va = smem_u32(sKV + stage * C::STAGE) + C::KV_TILE;
        const uint32_t pa = smem_u32(sP + (j & 1) * C::P_BYTES);
        const uint32_t acc = j != 0 ? 1u : 0u;
#pragma unroll
        for (int ks = 0; ks < TC_BKV / 16; ++ks) {
          const uint64_t da = umma_desc_sw128(pa) + static_cast<uint64_t>(2 * ks);
          const uint64_t db = umma_desc_sw128_mn(va + ks * 16 * 128, TC_BKV * 128);
          umma_f16(tmem_base + C::O_COL0, da, db, idesc_pv, (ks != 0) ? 1u : acc);
        }
        const uint32_t oa = smem_u32(sOnes);
#pragma unroll
        for (int ks = 0; ks < TC_BKV / 16; ++ks)
          umma_f16(tmem_base + C::SUM_COL0, umma_desc_sw128(pa) + static_cast<uint64_t>(2 * ks),
                   umma_desc_sw128(oa) + static_cast<uint64_t>(2 * ks), idesc_sum, (ks != 0) ? 1u : acc);
        umma_commit(&o_done[j & 1]);
        umma_commit(&kv_empty[stage]);
      }
      __syncwarp();
    }
  } else {
    // ------------------------------------------------------------------ softmax warps: thread == query row
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
    const uint32_t t_o = tmem_base + lane_addr + C::O_COL0;
    const uint32_t t_sum = tmem_base + lane_addr + C::SUM_COL0;
    float m_used = -INFINITY;
    const float sc = p.scale_log2;

    for (int j = 0; j < ntiles; ++j) {
      const int nvalid = (j < tiles0) ? (p.n0 - j * TC_BKV) : (n1 - (j - tiles0) * TC_BKV);
      mbar_wait(&s_full[j & 1], (j >> 1) & 1);
      tc_fence_after();
      const uint32_t t_s = tmem_base + lane_addr + (j & 1) * TC_BKV;
      uint32_t s0[32], s1[32];
      tmem_ld32(t_s, s0);
      tmem_ld32(t_s + 32, s1);
      tmem_ld_wait();
      if (nvalid < TC_BKV) {
#pragma unroll
        for (int k = 0; k < 32; ++k) {
          if (k >= nvalid) s0[k] = 0xff800000u;
          if (k + 32 >= nvalid) s1[k] = 0xff800000u;
        }
      }
      float mxa[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) mxa[k] = fmaxf(__uint_as_float(s0[k]), __uint_as_float(s1[k]));
#pragma unroll
      for (int k = 8; k < 32; ++k) mxa[k & 7] = fmax3(mxa[k & 7], __uint_as_float(s0[k]), __uint_as_float(s1[k]));
      const float mx = fmaxf(fmax3(mxa[0], mxa[1], mxa[2]), fmax3(fmax3(mxa[3], mxa[4], mxa[5]), mxa[6], mxa[7]));
      // lazy rescale: warp-uniform decision (tcgen05.ld / st are warp-collective); rows that did not need it move too
      const bool need = (mx - m_used) * sc > TC2_TAU;   // true on the first tile (m_used = -inf)
      if (__any_sync(0xffffffffu, need)) {
        const float m_new = fmaxf(m_used, mx);
        if (j > 0) {
          // every P V issued so far must have landed in O before it is rescaled; P V (j) waits for this warp's p_full
          mbar_wait(&o_done[(j - 1) & 1], ((j - 1) >> 1) & 1);
          tc_fence_after();
          const float f = ex2f((m_used - m_new) * sc);
#pragma unroll
          for (int c = 0; c < C::NPV / 16; ++c) {
            uint32_t r[16];
            tmem_ld16(t_o + c * 16, r);
            tmem_ld_wait();
#pragma unroll
            for (int k = 0; k < 16; ++k) r[k] = __float_as_uint(__uint_as_float(r[k]) * f);
            tmem_st16(t_o + c * 16, r);
          }
          const uint32_t l = tmem_ld1(t_sum);
          tmem_ld_wait();
          tmem_st1(t_sum, __float_as_uint(__uint_as_float(l) * f));
          tmem_st_wait();
        }
        m_used = m_new;
      }
      // P buffer j&1 was read by P V (j-2)
      if (j >= 2) mbar_wait(&o_done[j & 1], ((j - 2) >> 1) & 1);
      const float nmsc = -(m_used * sc);
      // row r of the P tile is 128 B; its 16-byte chunk c lives at chunk c ^ (r & 7) (128B swizzle): one XOR per store
      const uint32_t prow = (smem_u32(sP) + (j & 1) * C::P_BYTES + row * 128) ^ ((row & 7) << 4);
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        float pv[8];
#pragma unroll
        for (int k = 0; k < 8; k += 2) {
          const int idx = c * 8 + k;
          const float sa = __uint_as_float(idx < 32 ? s0[idx] : s1[idx - 32]);
          const float sb = __uint_as_float(idx < 32 ? s0[idx + 1] : s1[idx - 31]);
          float x0, x1;
          ffma2_bcast(x0, x1, sa, sb, sc, nmsc);
          if (k >= 8 - 2 * EMU) {
            exp2_poly2(pv[k], pv[k + 1], x0, x1);
          } else {
            pv[k] = ex2f(x0);
            pv[k + 1] = ex2f(x1);
          }
        }
        sts128(prow ^ (c << 4), pack_op16x2(pv[0], pv[1]), pack_op16x2(pv[2], pv[3]), pack_op16x2(pv[4], pv[5]),
               pack_op16x2(pv[6], pv[7]));
      }
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[j & 1]);
    }
    // ---- epilogue: O / l
    mbar_wait(&o_done[(ntiles - 1) & 1], ((ntiles - 1) >> 1) & 1);
    tc_fence_after();
    const uint32_t lsum = tmem_ld1(t_sum);
    tmem_ld_wait();
    const float l = __uint_as_float(lsum);
    const float inv = l > 0.f ? 1.f / l : 0.f;
    const int qrow = q0 + row;
    op16* og = p.out + (long long)b * p.o_bs + (long long)qrow * p.o_rs + h * p.d;
#pragma unroll
    for (int c = 0; c < C::NPV / 16; ++c) {   // 16 columns at a time keeps the register footprint flat for wide heads
      uint32_t o[16];
      tmem_ld16(t_o + c * 16, o);
      tmem_ld_wait();
      if (qrow < p.nq) {
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          if (c * 16 + hf * 8 < D) {
            uint4 w;
            w.x = pack_op16x2(__uint_as_float(o[hf * 8 + 0]) * inv, __uint_as_float(o[hf * 8 + 1]) * inv);
            w.y = pack_op16x2(__uint_as_float(o[hf * 8 + 2]) * inv, __uint_as_float(o[hf * 8 + 3]) * inv);
            w.z = pack_op16x2(__uint_as_float(o[hf * 8 + 4]) * inv, __uint_as_float(o[hf * 8 + 5]) * inv);
            w.w = pack_op16x2(__uint_as_float(o[hf * 8 + 6]) * inv, __uint_as_float(o[hf * 8 + 7]) * inv);
            *reinterpret_cast<uint4*>(og + c * 16 + hf * 8) = w;
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

static int make_all_kv_maps(const AttnTcDev& p, int batch, CUtensorMap& mk0, CUtensorMap& mv0, CUtensorMap& mk1,
                            CUtensorMap& mv1) {
  const int cols = p.heads * p.d;
  const int nb0 = (batch + p.kv0_div - 1) / p.kv0_div;
  if (int rc = make_kv_map(&mk0, p.k0, cols, p.n0, p.kv0_rs, p.kv0_bs, nb0)) return rc;
  if (int rc = make_kv_map(&mv0, p.v0, cols, p.n0, p.kv0_rs, p.kv0_bs, nb0)) return rc;
  if (p.n1 > 0) {
    const int nb1 = (batch + p.kv1_div - 1) / p.kv1_div;
    if (int rc = make_kv_map(&mk1, p.k1, cols, p.n1, p.kv1_rs, p.kv1_bs, nb1)) return rc;
    if (int rc = make_kv_map(&mv1, p.v1, cols, p.n1, p.kv1_rs, p.kv1_bs, nb1)) return rc;
  } else {
    mk1 = mk0;
    mv1 = mv0;
  }
  return 0;
}

template <int D, int EMU>
static int launch_tc(const AttnTcDev& p, int batch, cudaStream_t stream) {
  using C = TcCfg<D>;
  static PerDeviceOnce configured;
  int dev__ = 0;
  if (configured.pending(&dev__)) {
    cudaError_t e = cudaFuncSetAttribute(flash_attn_tc_kernel<D, EMU>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
    if (e != cudaSuccess) return set_error_cuda("cudaFuncSetAttribute(flash_attn_tc2)", e);
    configured.done(dev__);
  }
  CUtensorMap mk0, mv0, mk1, mv1;
  if (int rc = make_all_kv_maps(p, batch, mk0, mv0, mk1, mv1)) return rc;
  dim3 grid((p.nq + TC_BQ - 1) / TC_BQ, p.heads, batch);
  launch_kernel(flash_attn_tc_kernel<D, EMU>, dim3(grid), dim3(TC_THREADS), C::SMEM, stream, mk0, mv0, mk1, mv1, p);
  EMOTE_CHECK_LAUNCH("emote_attention_tc_bf16");
  return 0;
}

template <int D>
static int dispatch_tc(int emu, const AttnTcDev& p, int batch, cudaStream_t stream) {
  if (emu == 0) return launch_tc<D, 0>(p, batch, stream);
  if (emu == 2) return launch_tc<D, 2>(p, batch, stream);
  return launch_tc<D, 1>(p, batch, stream);
}

}  // namespace emote

using namespace emote;

extern "C" int emote_attention_tc_supported(int32_t head_dim) {
  return (head_dim == 40 || head_dim == 64 || head_dim == 80 || head_dim == 160) ? 1 : 0;
}

extern "C" int emote_attention_tc_bf16(const EmoteAttnArgs* a, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (!a || !a->q || !a->k0 || !a->v0 || !a->out) return set_error("emote_attention_tc_bf16: null pointer");
  if (a->batch <= 0 || a->heads <= 0 || a->nq <= 0 || a->n0 <= 0 || a->n1 < 0)
    return set_error("emote_attention_tc_bf16: bad sizes");
  if (!emote_attention_tc_supported(a->head_dim))
    return set_error("emote_attention_tc_bf16: head_dim must be 40, 64, 80 or 160");
  if (a->n1 > 0 && (!a->k1 || !a->v1)) return set_error("emote_attention_tc_bf16: segment 1 pointers missing");
  if (a->batch > 65535 || a->heads > 65535) return set_error("emote_attention_tc_bf16: grid too large");
  const int64_t strides[] = {a->q_batch_stride, a->q_row_stride, a->kv0_batch_stride, a->kv0_row_stride,
                             a->kv1_batch_stride, a->kv1_row_stride, a->o_batch_stride, a->o_row_stride};
  for (int64_t s : strides)
    if (s % 8 != 0) return set_error("emote_attention_tc_bf16: strides must keep rows 16-byte aligned");
  AttnTcDev p{};
  p.q = (const op16*)a->q; p.k0 = (const op16*)a->k0; p.v0 = (const op16*)a->v0;
  p.k1 = (const op16*)a->k1; p.v1 = (const op16*)a->v1; p.out = (op16*)a->out;
  p.heads = a->heads; p.d = a->head_dim; p.nq = a->nq; p.n0 = a->n0; p.n1 = a->n1;
  p.q_bs = a->q_batch_stride; p.q_rs = a->q_row_stride;
  p.kv0_bs = a->kv0_batch_stride; p.kv0_rs = a->kv0_row_stride;
  p.kv1_bs = a->kv1_batch_stride; p.kv1_rs = a->kv1_row_stride;
  p.o_bs = a->o_batch_stride; p.o_rs = a->o_row_stride;
  p.kv0_div = a->kv0_batch_div > 0 ? a->kv0_batch_div : 1;
  p.kv1_div = a->kv1_batch_div > 0 ? a->kv1_batch_div : 1;
  p.kv1_first = a->kv1_first_batch;
  p.scale_log2 = a->scale * 1.4426950408889634f;
  // share of the exponentials moved from MUFU.EX2 to the FMA pipe: 0, 1/4 or 1/2 (EMOTE_ATTN_EMU = 0 / 1 / 2, dev knob).
  // Measured on B200 (32 images x 8 heads): head_dim 40, N = 4096: 1.64 / 1.55 / 1.50 ms; head_dim 80: 1/4 is best.
  static const int emu_env = [] {
    const char* e = std::getenv("EMOTE_ATTN_EMU");
    return (e && e[0] >= '0' && e[0] <= '2') ? e[0] - '0' : -1;
  }();
  const int emu = emu_env >= 0 ? emu_env : (a->head_dim == 40 ? 2 : 1);
  switch (a->head_dim) {
    case 40: return dispatch_tc<40>(emu, p, a->batch, stream);
    case 64: return dispatch_tc<64>(emu, p, a->batch, stream);
    case 80: return dispatch_tc<80>(emu, p, a->batch, stream);
    default: return dispatch_tc<160>(emu, p, a->batch, stream);
  }
}
