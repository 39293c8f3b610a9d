// tcgen05 flash attention for ONE wide head (head_dim 512): the mid-block attention of the SD VAE decoder / encoder
// (legacy diffusers `AttentionBlock`, vendored at reference orig_attention.py:253-385: softmax(q k^T / sqrt(C)) v with a
// single 512-dim head over the 64 x 64 = 4096 positions of every frame).  Replaces the per-frame loop of
// QK^T GEMM -> [4096, 4096] fp32 scores -> row softmax -> PV GEMM (the reference does the same with baddbmm + softmax + bmm).
//
//   grid (N / 128 query tiles, 2 halves of the value / output columns, frames); 192 threads: warp 0 TMA loader,
//   warp 1 MMA issuer, warps 2-5 softmax (thread == query row), exactly the roles of flash_attn_tc_kernel.
//   Q tile [128 x 512] stays resident in shared memory (8 swizzle atoms, 128 KB).  Per 64-key tile:
//     S  = sum over 8 atoms Q_a K_a^T    K streams through a 4-stage ring of [64 keys x 64 dims] chunks (8 KB)
//     P  = exp2(S c - m_used)            lazy rescale of O / row sums in TMEM as in the v2 kernel
//     O += P V_half                      V half-tile [64 keys x 256] MN-major, accumulator 256 TMEM columns
//   TMEM: S double buffer 128 + O 256 + row sums 16 = 400 of 512 columns; shared memory 128 + 32 + 32 + 16 + 2 KB.
//   The two column halves recompute S (1.5x the FLOPs of an ideal kernel; O for all 512 columns plus S does not fit TMEM).
#include "attention_tc.cuh"

namespace emote {

constexpr int W_D = 512;                 // head dim
constexpr int W_ATOMS = W_D / 64;        // 8
constexpr int W_DV = 256;                // value / output columns per CTA
constexpr int W_VATOMS = W_DV / 64;      // 4
constexpr int W_KSTAGES = 4;
constexpr int W_Q_BYTES = W_ATOMS * TC_BQ * 128;       // 128 KB
constexpr int W_KCHUNK = TC_BKV * 128;                 // 8 KB
constexpr int W_V_BYTES = W_VATOMS * TC_BKV * 128;     // 32 KB
constexpr int W_P_BYTES = TC_BQ * 128;                 // 16 KB
constexpr int W_ONES_BYTES = 16 * 128;
constexpr int W_SMEM = W_Q_BYTES + W_KSTAGES * W_KCHUNK + W_V_BYTES + W_P_BYTES + W_ONES_BYTES + 256 + 1024;
constexpr int W_O_COL0 = 2 * TC_BKV;
constexpr int W_SUM_COL0 = W_O_COL0 + W_DV;
constexpr int W_TMEM_COLS = 512;

struct AttnWideDev {
  const op16* q;
  op16* out;
  int nq, nk;
  long long q_bs, q_rs, o_bs, o_rs;
  float scale_log2;
};

template <int EMU>
__global__ void __launch_bounds__(TC_THREADS, 1)
flash_attn_wide_kernel(const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                       const AttnWideDev p) {
  pdl_launch_early();
  extern __shared__ uint8_t wide_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(wide_smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + W_Q_BYTES;
  uint8_t* sV = sK + W_KSTAGES * W_KCHUNK;
  uint8_t* sP = sV + W_V_BYTES;
  uint8_t* sOnes = sP + W_P_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sOnes + W_ONES_BYTES);
  uint64_t* k_full = bars;                    // [KSTAGES]
  uint64_t* k_empty = k_full + W_KSTAGES;     // [KSTAGES]
  uint64_t* v_full = k_empty + W_KSTAGES;     // V half-tile of tile j landed (phase j)
  uint64_t* v_empty = v_full + 1;             // P V (j) retired: V / P buffers free (phase j)
  uint64_t* s_full = v_empty + 1;             // [2]
  uint64_t* p_full = s_full + 2;              // [2] 4 warp arrivals
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(p_full + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int img = blockIdx.z, half = blockIdx.y;
  const int q0 = blockIdx.x * TC_BQ;
  const int ntiles = (p.nk + TC_BKV - 1) / TC_BKV;

  {
    uint4* z = reinterpret_cast<uint4*>(smem);
    const int n16 = (W_Q_BYTES + W_KSTAGES * W_KCHUNK + W_V_BYTES + W_P_BYTES) / 16;
    for (int i = threadIdx.x; i < n16; i += TC_THREADS) z[i] = make_uint4(0u, 0u, 0u, 0u);
    uint4* o = reinterpret_cast<uint4*>(sOnes);
    for (int i = threadIdx.x; i < W_ONES_BYTES / 16; i += TC_THREADS)
      o[i] = make_uint4(OP16_ONE_PAIR, OP16_ONE_PAIR, OP16_ONE_PAIR, OP16_ONE_PAIR);
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < W_KSTAGES; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], 1);
    }
    mbar_init(v_full, 1);
    mbar_init(v_empty, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&s_full[s], 1);
      mbar_init(&p_full[s], 4);
    }
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, W_TMEM_COLS);
    tmem_relinquish();
  }
  __syncthreads();
  pdl_wait();
  {
    // Q tile: 128 rows x 64 chunks of 16 B; atom = chunk / 8, physical slot = (chunk % 8) ^ (row % 8)
    const op16* qg = p.q + (long long)img * p.q_bs + (long long)q0 * p.q_rs;
    const int nvq = p.nq - q0;
    constexpr int CH = W_D / 8;
    for (int i = threadIdx.x; i < TC_BQ * CH; i += TC_THREADS) {
      const int r = i / CH, c = i - r * CH;
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
    // ------------------------------------------------------------------ loader: K chunks through the ring, V half-tiles
    int kc = 0;   // running K-chunk counter
    for (int j = 0; j < ntiles; ++j) {
      for (int a = 0; a < W_ATOMS; ++a, ++kc) {
        const int stage = kc % W_KSTAGES;
        if (kc >= W_KSTAGES) mbar_wait(&k_empty[stage], ((kc / W_KSTAGES) - 1) & 1);
        if (elect_one()) {
          mbar_expect_tx(&k_full[stage], W_KCHUNK);
          tma_load_3d(sK + stage * W_KCHUNK, &tmK, &k_full[stage], a * 64, j * TC_BKV, img);
        }
        __syncwarp();
      }
      if (j > 0) mbar_wait(v_empty, (j - 1) & 1);   // P V (j-1) has read the V buffer
      if (elect_one()) {
        mbar_expect_tx(v_full, W_V_BYTES);
#pragma unroll
        for (int a = 0; a < W_VATOMS; ++a)
          tma_load_3d(sV + a * (TC_BKV * 128), &tmV, v_full, half * W_DV + a * 64, j * TC_BKV, img);
      }
      __syncwarp();
    }
    pdl_launch_late();
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc_s = umma_idesc_op16(TC_BQ, TC_BKV);
    constexpr uint32_t idesc_pv = umma_idesc_op16_bmn(TC_BQ, W_DV);
    constexpr uint32_t idesc_sum = umma_idesc_op16(TC_BQ, 16);
    int kc = 0;
    auto issue_s = [&](int j) {
      const uint32_t d_tmem = tmem_base + (j & 1) * TC_BKV;
      for (int a = 0; a < W_ATOMS; ++a, ++kc) {
        const int stage = kc % W_KSTAGES;
        mbar_wait(&k_full[stage], (kc / W_KSTAGES) & 1);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t qa = smem_u32(sQ) + a * (TC_BQ * 128);
          const uint32_t ka = smem_u32(sK + stage * W_KCHUNK);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            umma_f16(d_tmem, umma_desc_sw128(qa) + static_cast<uint64_t>(2 * ks),
                     umma_desc_sw128(ka) + static_cast<uint64_t>(2 * ks), idesc_s, (a | ks) != 0 ? 1u : 0u);
          umma_commit(&k_empty[stage]);
          if (a == W_ATOMS - 1) umma_commit(&s_full[j & 1]);
        }
        __syncwarp();
      }
    };
    issue_s(0);
    for (int j = 0; j < ntiles; ++j) {
      if (j + 1 < ntiles) issue_s(j + 1);   // S buffer (j+1)&1 was released by p_full(j-1), waited last iteration
      mbar_wait(&p_full[j & 1], (j >> 1) & 1);
      mbar_wait(v_full, j & 1);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t va = smem_u32(sV);
        const uint32_t pa = smem_u32(sP);
        const uint32_t acc = j != 0 ? 1u : 0u;
#pragma unroll
        for (int ks = 0; ks < TC_BKV / 16; ++ks)
          umma_f16(tmem_base + W_O_COL0, umma_desc_sw128(pa) + static_cast<uint64_t>(2 * ks),
                   umma_desc_sw128_mn(va + ks * 16 * 128, TC_BKV * 128), idesc_pv, (ks != 0) ? 1u : acc);
        const uint32_t oa = smem_u32(sOnes);
#pragma unroll
        for (int ks = 0; ks < TC_BKV / 16; ++ks)
          umma_f16(tmem_base + W_SUM_COL0, umma_desc_sw128(pa) + static_cast<uint64_t>(2 * ks),
                   umma_desc_sw128(oa) + static_cast<uint64_t>(2 * ks), idesc_sum, (ks != 0) ? 1u : acc);
        umma_commit(v_empty);
      }
      __syncwarp();
    }
  } else {
    // ------------------------------------------------------------------ softmax warps: thread == query row
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
    const uint32_t t_o = tmem_base + lane_addr + W_O_COL0;
    const uint32_t t_sum = tmem_base + lane_addr + W_SUM_COL0;
    float m_used = -INFINITY;
    const float sc = p.scale_log2;
    uint8_t* prow = sP + row * 128;

    for (int j = 0; j < ntiles; ++j) {
      const int nvalid = p.nk - j * TC_BKV;
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
      // the single P buffer was read by P V (j-1); the same wait orders this thread's O accesses after that MMA
      if (j > 0) {
        mbar_wait(v_empty, (j - 1) & 1);
        tc_fence_after();
      }
      const bool need = (mx - m_used) * sc > TC2_TAU;
      if (__any_sync(0xffffffffu, need)) {
        const float m_new = fmaxf(m_used, mx);
        if (j > 0) {
          const float f = ex2f((m_used - m_new) * sc);
#pragma unroll
          for (int c = 0; c < W_DV / 16; ++c) {
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
      const float nmsc = -(m_used * sc);
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
        uint4 w;
        w.x = pack_op16x2(pv[0], pv[1]); w.y = pack_op16x2(pv[2], pv[3]);
        w.z = pack_op16x2(pv[4], pv[5]); w.w = pack_op16x2(pv[6], pv[7]);
        *reinterpret_cast<uint4*>(prow + ((c ^ (row & 7)) << 4)) = w;
      }
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[j & 1]);
    }
    mbar_wait(v_empty, (ntiles - 1) & 1);
    tc_fence_after();
    const uint32_t lsum = tmem_ld1(t_sum);
    tmem_ld_wait();
    const float l = __uint_as_float(lsum);
    const float inv = l > 0.f ? 1.f / l : 0.f;
    const int qrow = q0 + row;
    op16* og = p.out + (long long)img * p.o_bs + (long long)qrow * p.o_rs + half * W_DV;
#pragma unroll 4
    for (int c = 0; c < W_DV / 16; ++c) {
      uint32_t o[16];
      tmem_ld16(t_o + c * 16, o);
      tmem_ld_wait();
      if (qrow < p.nq) {
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
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

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, W_TMEM_COLS);
  }
}

}  // namespace emote

using namespace emote;

extern "C" int emote_attention_wide_bf16(const void* q, const void* k, const void* v, void* out, int32_t batch, int32_t nq,
                                         int32_t nk, int32_t head_dim, int64_t q_batch_stride, int64_t q_row_stride,
                                         int64_t kv_batch_stride, int64_t kv_row_stride, int64_t o_batch_stride,
                                         int64_t o_row_stride, float scale, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (!q || !k || !v || !out) return set_error("emote_attention_wide_bf16: null pointer");
  if (head_dim != W_D) return set_error("emote_attention_wide_bf16: head_dim must be 512");
  if (batch <= 0 || batch > 65535 || nq <= 0 || nk <= 0) return set_error("emote_attention_wide_bf16: bad sizes");
  const int64_t strides[] = {q_batch_stride, q_row_stride, kv_batch_stride, kv_row_stride, o_batch_stride, o_row_stride};
  for (int64_t s : strides)
    if (s % 8 != 0) return set_error("emote_attention_wide_bf16: strides must keep rows 16-byte aligned");
  static PerDeviceOnce configured;
  int dev__ = 0;
  if (configured.pending(&dev__)) {
    cudaError_t e = cudaFuncSetAttribute(flash_attn_wide_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, W_SMEM);
    if (e != cudaSuccess) return set_error_cuda("cudaFuncSetAttribute(flash_attn_wide)", e);
    configured.done(dev__);
  }
  CUtensorMap mk, mv;
  if (int rc = make_kv_map(&mk, k, W_D, nk, kv_row_stride, kv_batch_stride, batch)) return rc;
  if (int rc = make_kv_map(&mv, v, W_D, nk, kv_row_stride, kv_batch_stride, batch)) return rc;
  AttnWideDev p{};
  p.q = (const op16*)q; p.out = (op16*)out; p.nq = nq; p.nk = nk;
  p.q_bs = q_batch_stride; p.q_rs = q_row_stride; p.o_bs = o_batch_stride; p.o_rs = o_row_stride;
  p.scale_log2 = scale * 1.4426950408889634f;
  dim3 grid((nq + TC_BQ - 1) / TC_BQ, W_D / W_DV, batch);
  launch_kernel(flash_attn_wide_kernel<1>, dim3(grid), dim3(TC_THREADS), W_SMEM, stream, mk, mv, p);
  EMOTE_CHECK_LAUNCH("emote_attention_wide_bf16");
  return 0;
}
