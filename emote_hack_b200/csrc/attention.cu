// Attention kernels (bf16 operands, fp32 softmax + accumulation, warp-level mma.sync tensor-core tiles).
//
//  flash_attn_kernel      spatial self-attention, reference-attention (two key/value segments:
//                         [self tokens | ReferenceNet bank]) and text/audio cross-attention.
//                         Replaces baddbmm -> softmax -> bmm of CrossAttention._attention
//                         (orig_attention.py:655-684) without materialising the score matrix, and the
//                         torch.cat + recompute-uncond trick of mutual_self_attention.py:239-255.
//  temporal_attn_kernel   VersatileAttention over the frame axis (motion_module.py:275-334): one warp per
//                         (sample, pixel, head); the (b f) d c <-> (b d) f c rearranges are address arithmetic.
#include "common.cuh"
#include "emote_b200.h"
#include "host_utils.h"

namespace emote {

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, bool valid) {
  const uint32_t d = smem_u32(smem_dst);
  const int sz = valid ? 16 : 0;  // src-size 0 -> zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                               uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32." EMOTE_OP16_PTX "." EMOTE_OP16_PTX ".f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

struct AttnDev {
  const op16 *q, *k0, *v0, *k1, *v1;
  op16* out;
  int heads, d;
  int nq, n0, n1;
  long long q_bs, q_rs, kv0_bs, kv0_rs, kv1_bs, kv1_rs, o_bs, o_rs;
  int kv0_div, kv1_div, kv1_first;
  float scale_log2;
};

constexpr int FA_BKV = 64;
constexpr int FA_THREADS = 128;

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Load `rows` x d bf16 (row stride rs elements) into smem [rows][DP+8]; rows >= nvalid and cols >= d are zero.
template <int DP>
__device__ __forceinline__ void fa_load_tile(op16* s, const op16* g, long long rs, int rows,
                                             int nvalid, int d) {
  constexpr int PITCH = DP + 8;
  constexpr int CH = DP / 8;  // 16-byte chunks per padded row
  const int dch = d >> 3;
  for (int i = threadIdx.x; i < rows * CH; i += FA_THREADS) {
    const int r = i / CH;
    const int c = i - r * CH;
    const bool ok = (r < nvalid) && (c < dch);
    const op16* src = ok ? g + (long long)r * rs + c * 8 : g;
    cp_async16(s + r * PITCH + c * 8, src, ok);
  }
}

// MT = 16-row m-tiles per warp: 4 warps x MT x 16 query rows per CTA.  With MT = 2 every K / V fragment fetched by
// ldmatrix feeds two independent MMAs (half the shared-memory traffic per FLOP, twice the MMA-level parallelism).
template <int DP, int MT>
__global__ void __launch_bounds__(FA_THREADS, (DP <= 96) ? 2 : 1) flash_attn_kernel(const AttnDev p) {
  pdl_prologue();
  constexpr int PITCH = DP + 8;
  constexpr int KS = DP / 16;  // k-steps over the head dim
  constexpr int NT = DP / 8;   // output n-tiles
  constexpr int BQ = 64 * MT;
  extern __shared__ __align__(16) uint8_t fa_smem[];
  op16* sQ = reinterpret_cast<op16*>(fa_smem);
  op16* sK = sQ + BQ * PITCH;           // [2][64][PITCH]
  op16* sV = sK + 2 * 64 * PITCH;       // [2][64][PITCH]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int b = blockIdx.z, h = blockIdx.y;
  const int q0 = blockIdx.x * BQ;

  const op16* qg = p.q + (long long)b * p.q_bs + (long long)q0 * p.q_rs + h * p.d;
  const op16* k0g = p.k0 + (long long)(b / p.kv0_div) * p.kv0_bs + h * p.d;
  const op16* v0g = p.v0 + (long long)(b / p.kv0_div) * p.kv0_bs + h * p.d;
  const int n1 = (p.n1 > 0 && b >= p.kv1_first) ? p.n1 : 0;
  const op16* k1g = nullptr;
  const op16* v1g = nullptr;
  if (n1 > 0) {
    k1g = p.k1 + (long long)(b / p.kv1_div) * p.kv1_bs + h * p.d;
    v1g = p.v1 + (long long)(b / p.kv1_div) * p.kv1_bs + h * p.d;
  }
  const int tiles0 = (p.n0 + FA_BKV - 1) / FA_BKV;
  const int tiles1 = (n1 + FA_BKV - 1) / FA_BKV;
  const int ntiles = tiles0 + tiles1;

  auto issue_kv = [&](int tile, int buf) {
    const op16 *kg, *vg;
    long long rs;
    int nvalid;
    if (tile < tiles0) {
      kg = k0g + (long long)tile * FA_BKV * p.kv0_rs;
      vg = v0g + (long long)tile * FA_BKV * p.kv0_rs;
      rs = p.kv0_rs;
      nvalid = p.n0 - tile * FA_BKV;
    } else {
      const int t1 = tile - tiles0;
      kg = k1g + (long long)t1 * FA_BKV * p.kv1_rs;
      vg = v1g + (long long)t1 * FA_BKV * p.kv1_rs;
      rs = p.kv1_rs;
      nvalid = n1 - t1 * FA_BKV;
    }
    if (nvalid > FA_BKV) nvalid = FA_BKV;
    fa_load_tile<DP>(sK + buf * 64 * PITCH, kg, rs, 64, nvalid, p.d);
    fa_load_tile<DP>(sV + buf * 64 * PITCH, vg, rs, 64, nvalid, p.d);
  };

  {
    int nvq = p.nq - q0;
    if (nvq > BQ) nvq = BQ;
    fa_load_tile<DP>(sQ, qg, p.q_rs, BQ, nvq, p.d);
  }
  issue_kv(0, 0);
  cp_async_commit();

  float o_acc[MT][NT][4];
  float m_run[MT][2], l_run[MT][2];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
    for (int i = 0; i < NT; ++i) o_acc[mt][i][0] = o_acc[mt][i][1] = o_acc[mt][i][2] = o_acc[mt][i][3] = 0.f;
    m_run[mt][0] = m_run[mt][1] = -INFINITY;
    l_run[mt][0] = l_run[mt][1] = 0.f;
  }
  uint32_t qf[MT][KS][4];
  const int qrow_w = warp * 16 * MT;  // first query row (within the CTA tile) owned by this warp

  for (int tile = 0; tile < ntiles; ++tile) {
    const int buf = tile & 1;
    if (tile + 1 < ntiles) {
      issue_kv(tile + 1, buf ^ 1);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    if (tile == 0) {
      // Q fragments stay in registers for the whole kernel
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
          const int row = qrow_w + mt * 16 + (lane & 15);
          const int col = ks * 16 + (lane >> 4) * 8;
          ldsm_x4(smem_u32(sQ + row * PITCH + col), qf[mt][ks][0], qf[mt][ks][1], qf[mt][ks][2], qf[mt][ks][3]);
        }
    }
    const op16* sKb = sK + buf * 64 * PITCH;
    const op16* sVb = sV + buf * 64 * PITCH;

    // ---- S = Q K^T  (MT x 16 x 64 per warp)
    float s_acc[MT][8][4];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int i = 0; i < 8; ++i) s_acc[mt][i][0] = s_acc[mt][i][1] = s_acc[mt][i][2] = s_acc[mt][i][3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
      for (int np = 0; np < 4; ++np) {  // pairs of n-tiles (16 kv rows)
        uint32_t b0, b1, b2, b3;
        const int row = np * 16 + (lane & 7) + (lane >> 4) * 8;
        const int col = ks * 16 + ((lane >> 3) & 1) * 8;
        ldsm_x4(smem_u32(sKb + row * PITCH + col), b0, b1, b2, b3);
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          mma_bf16_16816(s_acc[mt][2 * np], qf[mt][ks][0], qf[mt][ks][1], qf[mt][ks][2], qf[mt][ks][3], b0, b1);
          mma_bf16_16816(s_acc[mt][2 * np + 1], qf[mt][ks][0], qf[mt][ks][1], qf[mt][ks][2], qf[mt][ks][3], b2, b3);
        }
      }
    }
    // ---- mask the ragged tail of the segment
    int nvalid;
    if (tile < tiles0) nvalid = p.n0 - tile * FA_BKV; else nvalid = n1 - (tile - tiles0) * FA_BKV;
    if (nvalid < FA_BKV) {
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
          const int c = nt * 8 + 2 * t;
          if (c >= nvalid) s_acc[mt][nt][0] = s_acc[mt][nt][2] = -INFINITY;
          if (c + 1 >= nvalid) s_acc[mt][nt][1] = s_acc[mt][nt][3] = -INFINITY;
        }
    }
    // ---- online softmax (rows g and g+8 of each 16-row m-tile); P packed as A fragments
    uint32_t pf[MT][4][4];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        mx[0] = fmaxf(mx[0], fmaxf(s_acc[mt][nt][0], s_acc[mt][nt][1]));
        mx[1] = fmaxf(mx[1], fmaxf(s_acc[mt][nt][2], s_acc[mt][nt][3]));
      }
      float corr[2], msc[2];
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
        mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
        const float m_new = fmaxf(m_run[mt][r], mx[r]);
        corr[r] = (m_run[mt][r] == -INFINITY) ? 0.f : ex2_approx((m_run[mt][r] - m_new) * p.scale_log2);
        msc[r] = (m_new == -INFINITY) ? 0.f : m_new * p.scale_log2;
        m_run[mt][r] = m_new;
      }
      float rs[2] = {0.f, 0.f};
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const float p0 = ex2_approx(fmaf(s_acc[mt][nt][0], p.scale_log2, -msc[0]));
        const float p1 = ex2_approx(fmaf(s_acc[mt][nt][1], p.scale_log2, -msc[0]));
        const float p2 = ex2_approx(fmaf(s_acc[mt][nt][2], p.scale_log2, -msc[1]));
        const float p3 = ex2_approx(fmaf(s_acc[mt][nt][3], p.scale_log2, -msc[1]));
        rs[0] += p0 + p1;
        rs[1] += p2 + p3;
        const int kk = nt >> 1;
        if ((nt & 1) == 0) {
          pf[mt][kk][0] = pack_op16x2(p0, p1);
          pf[mt][kk][1] = pack_op16x2(p2, p3);
        } else {
          pf[mt][kk][2] = pack_op16x2(p0, p1);
          pf[mt][kk][3] = pack_op16x2(p2, p3);
        }
      }
#pragma unroll
      for (int r = 0; r < 2; ++r) l_run[mt][r] = l_run[mt][r] * corr[r] + rs[r];
#pragma unroll
      for (int i = 0; i < NT; ++i) {
        o_acc[mt][i][0] *= corr[0]; o_acc[mt][i][1] *= corr[0];
        o_acc[mt][i][2] *= corr[1]; o_acc[mt][i][3] *= corr[1];
      }
    }
    // ---- O += P V
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
      for (int np = 0; np < NT / 2; ++np) {
        uint32_t b0, b1, b2, b3;
        const int row = kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
        const int col = np * 16 + (lane >> 4) * 8;
        ldsm_x4_t(smem_u32(sVb + row * PITCH + col), b0, b1, b2, b3);
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          mma_bf16_16816(o_acc[mt][2 * np], pf[mt][kk][0], pf[mt][kk][1], pf[mt][kk][2], pf[mt][kk][3], b0, b1);
          mma_bf16_16816(o_acc[mt][2 * np + 1], pf[mt][kk][0], pf[mt][kk][1], pf[mt][kk][2], pf[mt][kk][3], b2, b3);
        }
      }
    }
    __syncthreads();  // everyone done with buf before it is refilled
  }

  // ---- finalise: O / l, stage through this warp's Q rows, 16-byte stores
  const int dch = p.d >> 3;
  op16* og = p.out + (long long)b * p.o_bs + h * p.d;
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      l_run[mt][r] += __shfl_xor_sync(0xffffffffu, l_run[mt][r], 1);
      l_run[mt][r] += __shfl_xor_sync(0xffffffffu, l_run[mt][r], 2);
    }
    const float inv0 = l_run[mt][0] > 0.f ? 1.f / l_run[mt][0] : 0.f;
    const float inv1 = l_run[mt][1] > 0.f ? 1.f / l_run[mt][1] : 0.f;
    op16* sO = sQ + (qrow_w + mt * 16) * PITCH;
#pragma unroll
    for (int i = 0; i < NT; ++i) {
      *reinterpret_cast<uint32_t*>(sO + g * PITCH + i * 8 + 2 * t) =
          pack_op16x2(o_acc[mt][i][0] * inv0, o_acc[mt][i][1] * inv0);
      *reinterpret_cast<uint32_t*>(sO + (g + 8) * PITCH + i * 8 + 2 * t) =
          pack_op16x2(o_acc[mt][i][2] * inv1, o_acc[mt][i][3] * inv1);
    }
  }
  __syncwarp();
  for (int i = lane; i < 16 * MT * dch; i += 32) {
    const int r = i / dch, c = i - r * dch;
    const int qrow = q0 + qrow_w + r;
    if (qrow < p.nq) {
      const uint4 v = *reinterpret_cast<const uint4*>(sQ + (qrow_w + r) * PITCH + c * 8);
      *reinterpret_cast<uint4*>(og + (long long)qrow * p.o_rs + c * 8) = v;
    }
  }
}

// --------------------------------------------------------------------------- short-key-set attention
// Cross-attention to the text (77 tokens) / audio (5 tokens) context: the whole K and V of one (context, head) fit in
// shared memory, so a CTA loads them ONCE and then streams 64-row query tiles (double buffered) through a single-pass
// softmax.  The generic flash kernel paid its per-CTA setup (K/V tile loads, pipeline fill) for only 77 keys of work.
// NK16 = number of 16-key groups covering the key set (keys beyond n0 are zero-filled and masked).
template <int DP, int NK16>
__global__ void __launch_bounds__(FA_THREADS) short_kv_attn_kernel(const AttnDev p, int q_tiles_per_cta) {
  pdl_prologue();
  constexpr int PITCH = DP + 8;
  constexpr int KS = DP / 16;
  constexpr int NT = DP / 8;
  constexpr int NKEY = NK16 * 16;
  constexpr int SNT = NK16 * 2;  // score n-tiles of 8 keys
  extern __shared__ __align__(16) uint8_t sk_smem[];
  op16* sK = reinterpret_cast<op16*>(sk_smem);
  op16* sV = sK + NKEY * PITCH;
  op16* sQ = sV + NKEY * PITCH;  // [2][64][PITCH]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int b = blockIdx.z, h = blockIdx.y;
  const op16* kg = p.k0 + (long long)(b / p.kv0_div) * p.kv0_bs + h * p.d;
  const op16* vg = p.v0 + (long long)(b / p.kv0_div) * p.kv0_bs + h * p.d;
  const op16* qb = p.q + (long long)b * p.q_bs + h * p.d;
  op16* ob = p.out + (long long)b * p.o_bs + h * p.d;
  const int n_qtiles = (p.nq + 63) / 64;
  const int qt0 = blockIdx.x * q_tiles_per_cta;
  int qt1 = qt0 + q_tiles_per_cta;
  if (qt1 > n_qtiles) qt1 = n_qtiles;
  if (qt0 >= qt1) return;

  fa_load_tile<DP>(sK, kg, p.kv0_rs, NKEY, p.n0, p.d);
  fa_load_tile<DP>(sV, vg, p.kv0_rs, NKEY, p.n0, p.d);
  {
    int nv = p.nq - qt0 * 64;
    fa_load_tile<DP>(sQ, qb + (long long)qt0 * 64 * p.q_rs, p.q_rs, 64, nv > 64 ? 64 : nv, p.d);
  }
  cp_async_commit();
  const int dch = p.d >> 3;

  for (int qt = qt0; qt < qt1; ++qt) {
    const int buf = (qt - qt0) & 1;
    if (qt + 1 < qt1) {
      int nv = p.nq - (qt + 1) * 64;
      fa_load_tile<DP>(sQ + (buf ^ 1) * 64 * PITCH, qb + (long long)(qt + 1) * 64 * p.q_rs, p.q_rs, 64, nv > 64 ? 64 : nv, p.d);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    op16* sQb = sQ + buf * 64 * PITCH;
    float s_acc[SNT][4];
#pragma unroll
    for (int i = 0; i < SNT; ++i) s_acc[i][0] = s_acc[i][1] = s_acc[i][2] = s_acc[i][3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      uint32_t a0, a1, a2, a3;
      ldsm_x4(smem_u32(sQb + (warp * 16 + (lane & 15)) * PITCH + ks * 16 + (lane >> 4) * 8), a0, a1, a2, a3);
#pragma unroll
      for (int np = 0; np < NK16; ++np) {
        uint32_t b0, b1, b2, b3;
        const int row = np * 16 + (lane & 7) + (lane >> 4) * 8;
        const int col = ks * 16 + ((lane >> 3) & 1) * 8;
        ldsm_x4(smem_u32(sK + row * PITCH + col), b0, b1, b2, b3);
        mma_bf16_16816(s_acc[2 * np], a0, a1, a2, a3, b0, b1);
        mma_bf16_16816(s_acc[2 * np + 1], a0, a1, a2, a3, b2, b3);
      }
    }
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int nt = 0; nt < SNT; ++nt) {
      const int c = nt * 8 + 2 * t;
      if (c >= p.n0) s_acc[nt][0] = s_acc[nt][2] = -INFINITY;
      if (c + 1 >= p.n0) s_acc[nt][1] = s_acc[nt][3] = -INFINITY;
      mx[0] = fmaxf(mx[0], fmaxf(s_acc[nt][0], s_acc[nt][1]));
      mx[1] = fmaxf(mx[1], fmaxf(s_acc[nt][2], s_acc[nt][3]));
    }
    float sum[2] = {0.f, 0.f};
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
      mx[r] *= p.scale_log2;
    }
    uint32_t pf[NK16][4];
#pragma unroll
    for (int nt = 0; nt < SNT; ++nt) {
      const float p0 = ex2_approx(fmaf(s_acc[nt][0], p.scale_log2, -mx[0]));
      const float p1 = ex2_approx(fmaf(s_acc[nt][1], p.scale_log2, -mx[0]));
      const float p2 = ex2_approx(fmaf(s_acc[nt][2], p.scale_log2, -mx[1]));
      const float p3 = ex2_approx(fmaf(s_acc[nt][3], p.scale_log2, -mx[1]));
      sum[0] += p0 + p1;
      sum[1] += p2 + p3;
      const int kk = nt >> 1;
      if ((nt & 1) == 0) {
        pf[kk][0] = pack_op16x2(p0, p1);
        pf[kk][1] = pack_op16x2(p2, p3);
      } else {
        pf[kk][2] = pack_op16x2(p0, p1);
        pf[kk][3] = pack_op16x2(p2, p3);
      }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      sum[r] += __shfl_xor_sync(0xffffffffu, sum[r], 1);
      sum[r] += __shfl_xor_sync(0xffffffffu, sum[r], 2);
    }
    const float inv0 = 1.f / sum[0], inv1 = 1.f / sum[1];
    float o_acc[NT][4];
#pragma unroll
    for (int i = 0; i < NT; ++i) o_acc[i][0] = o_acc[i][1] = o_acc[i][2] = o_acc[i][3] = 0.f;
#pragma unroll
    for (int kk = 0; kk < NK16; ++kk) {
#pragma unroll
      for (int np = 0; np < NT / 2; ++np) {
        uint32_t b0, b1, b2, b3;
        const int row = kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
        const int col = np * 16 + (lane >> 4) * 8;
        ldsm_x4_t(smem_u32(sV + row * PITCH + col), b0, b1, b2, b3);
        mma_bf16_16816(o_acc[2 * np], pf[kk][0], pf[kk][1], pf[kk][2], pf[kk][3], b0, b1);
        mma_bf16_16816(o_acc[2 * np + 1], pf[kk][0], pf[kk][1], pf[kk][2], pf[kk][3], b2, b3);
      }
    }
    // stage the warp's 16 output rows through its (consumed) Q rows, then 16-byte stores
    __syncwarp();
    op16* sO = sQb + warp * 16 * PITCH;
#pragma unroll
    for (int i = 0; i < NT; ++i) {
      *reinterpret_cast<uint32_t*>(sO + g * PITCH + i * 8 + 2 * t) = pack_op16x2(o_acc[i][0] * inv0, o_acc[i][1] * inv0);
      *reinterpret_cast<uint32_t*>(sO + (g + 8) * PITCH + i * 8 + 2 * t) =
          pack_op16x2(o_acc[i][2] * inv1, o_acc[i][3] * inv1);
    }
    __syncwarp();
    for (int i = lane; i < 16 * dch; i += 32) {
      const int r = i / dch, c = i - r * dch;
      const int qrow = qt * 64 + warp * 16 + r;
      if (qrow < p.nq) {
        const uint4 v = *reinterpret_cast<const uint4*>(sO + r * PITCH + c * 8);
        *reinterpret_cast<uint4*>(ob + (long long)qrow * p.o_rs + c * 8) = v;
      }
    }
    __syncthreads();  // Q buffer `buf` is refilled two iterations later; everyone must be done with it
  }
}

// --------------------------------------------------------------------------- temporal attention
// TA_WARPS heads of one (sample, pixel) per block: 4 by default, 8 (a whole 8-head row per block) through the
// "temporal_warps" knob
template <int DP, int FP, int TA_WARPS>
__global__ void __launch_bounds__(TA_WARPS * 32) temporal_attn_kernel(const op16* __restrict__ qkv,
                                                                        op16* __restrict__ out, int F, int HW,
                                                                        int heads, int d, float scale_log2) {
  pdl_prologue();
  constexpr int PITCH = DP + 8;
  constexpr int KS = DP / 16;
  constexpr int NT = DP / 8;
  constexpr int MT = FP / 16;
  constexpr int SNT = FP / 8;  // score n-tiles
  extern __shared__ __align__(16) uint8_t ta_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int hgroups = (heads + TA_WARPS - 1) / TA_WARPS;
  const long long pix = blockIdx.x / hgroups;  // b*HW + p
  const int h = (int)(blockIdx.x % hgroups) * TA_WARPS + warp;
  if (h >= heads) return;
  const long long b = pix / HW, pidx = pix % HW;
  const int C = heads * d;
  op16* sQ = reinterpret_cast<op16*>(ta_smem) + warp * 3 * FP * PITCH;
  op16* sK = sQ + FP * PITCH;
  op16* sV = sK + FP * PITCH;

  // rows of this (b, pixel): token (b*F + f)*HW + pidx
  constexpr int CH = DP / 8;
  const int dch = d >> 3;
  for (int i = lane; i < 3 * FP * CH; i += 32) {
    const int which = i / (FP * CH);
    const int rem = i - which * FP * CH;
    const int f = rem / CH, c = rem - f * CH;
    const bool ok = (f < F) && (c < dch);
    const op16* src = qkv;
    if (ok) src = qkv + ((b * F + f) * HW + pidx) * (3LL * C) + (long long)which * C + h * d + c * 8;
    cp_async16(sQ + which * FP * PITCH + f * PITCH + c * 8, src, ok);
  }
  cp_async_commit();
  cp_async_wait<0>();
  __syncwarp();

#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
    float s_acc[SNT][4];
#pragma unroll
    for (int i = 0; i < SNT; ++i) s_acc[i][0] = s_acc[i][1] = s_acc[i][2] = s_acc[i][3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      uint32_t a0, a1, a2, a3;
      ldsm_x4(smem_u32(sQ + (mt * 16 + (lane & 15)) * PITCH + ks * 16 + (lane >> 4) * 8), a0, a1, a2, a3);
#pragma unroll
      for (int np = 0; np < SNT / 2; ++np) {
        uint32_t b0, b1, b2, b3;
        const int row = np * 16 + (lane & 7) + (lane >> 4) * 8;
        const int col = ks * 16 + ((lane >> 3) & 1) * 8;
        ldsm_x4(smem_u32(sK + row * PITCH + col), b0, b1, b2, b3);
        mma_bf16_16816(s_acc[2 * np], a0, a1, a2, a3, b0, b1);
        mma_bf16_16816(s_acc[2 * np + 1], a0, a1, a2, a3, b2, b3);
      }
    }
    if (F < FP) {
#pragma unroll
      for (int nt = 0; nt < SNT; ++nt) {
        const int c = nt * 8 + 2 * t;
        if (c >= F) s_acc[nt][0] = s_acc[nt][2] = -INFINITY;
        if (c + 1 >= F) s_acc[nt][1] = s_acc[nt][3] = -INFINITY;
      }
    }
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int nt = 0; nt < SNT; ++nt) {
      mx[0] = fmaxf(mx[0], fmaxf(s_acc[nt][0], s_acc[nt][1]));
      mx[1] = fmaxf(mx[1], fmaxf(s_acc[nt][2], s_acc[nt][3]));
    }
    float sum[2] = {0.f, 0.f};
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
      mx[r] *= scale_log2;
    }
    uint32_t pf[SNT / 2][4];
#pragma unroll
    for (int nt = 0; nt < SNT; ++nt) {
      const float p0 = exp2f(s_acc[nt][0] * scale_log2 - mx[0]);
      const float p1 = exp2f(s_acc[nt][1] * scale_log2 - mx[0]);
      const float p2 = exp2f(s_acc[nt][2] * scale_log2 - mx[1]);
      const float p3 = exp2f(s_acc[nt][3] * scale_log2 - mx[1]);
      sum[0] += p0 + p1;
      sum[1] += p2 + p3;
      const int kk = nt >> 1;
      if ((nt & 1) == 0) {
        pf[kk][0] = pack_op16x2(p0, p1);
        pf[kk][1] = pack_op16x2(p2, p3);
      } else {
        pf[kk][2] = pack_op16x2(p0, p1);
        pf[kk][3] = pack_op16x2(p2, p3);
      }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      sum[r] += __shfl_xor_sync(0xffffffffu, sum[r], 1);
      sum[r] += __shfl_xor_sync(0xffffffffu, sum[r], 2);
    }
    const float inv0 = 1.f / sum[0], inv1 = 1.f / sum[1];
    float o_acc[NT][4];
#pragma unroll
    for (int i = 0; i < NT; ++i) o_acc[i][0] = o_acc[i][1] = o_acc[i][2] = o_acc[i][3] = 0.f;
#pragma unroll
    for (int kk = 0; kk < SNT / 2; ++kk) {
#pragma unroll
      for (int np = 0; np < NT / 2; ++np) {
        uint32_t b0, b1, b2, b3;
        const int row = kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
        const int col = np * 16 + (lane >> 4) * 8;
        ldsm_x4_t(smem_u32(sV + row * PITCH + col), b0, b1, b2, b3);
        mma_bf16_16816(o_acc[2 * np], pf[kk][0], pf[kk][1], pf[kk][2], pf[kk][3], b0, b1);
        mma_bf16_16816(o_acc[2 * np + 1], pf[kk][0], pf[kk][1], pf[kk][2], pf[kk][3], b2, b3);
      }
    }
    // all lanes have consumed Q rows of this m-tile -> reuse them as the output staging area
    __syncwarp();
    op16* sO = sQ + mt * 16 * PITCH;
#pragma unroll
    for (int i = 0; i < NT; ++i) {
      *reinterpret_cast<uint32_t*>(sO + g * PITCH + i * 8 + 2 * t) = pack_op16x2(o_acc[i][0] * inv0, o_acc[i][1] * inv0);
      *reinterpret_cast<uint32_t*>(sO + (g + 8) * PITCH + i * 8 + 2 * t) =
          pack_op16x2(o_acc[i][2] * inv1, o_acc[i][3] * inv1);
    }
  }
  __syncwarp();
  for (int i = lane; i < FP * dch; i += 32) {
    const int f = i / dch, c = i - f * dch;
    if (f < F) {
      const uint4 v = *reinterpret_cast<const uint4*>(sQ + f * PITCH + c * 8);
      *reinterpret_cast<uint4*>(out + ((b * F + f) * HW + pidx) * (long long)C + h * d + c * 8) = v;
    }
  }
}

template <int DP>
static int launch_flash(const AttnDev& p, int batch, cudaStream_t stream) {
  // 32 query rows per warp while the accumulators still fit the register file, 16 for the widest heads
  constexpr int MT = (DP <= 80) ? 2 : 1;
  constexpr int BQ = 64 * MT;
  constexpr int SMEM = (BQ + 4 * 64) * (DP + 8) * 2;
  static PerDeviceOnce configured;
  int dev__ = 0;
  if (configured.pending(&dev__)) {
    cudaError_t e = cudaFuncSetAttribute(flash_attn_kernel<DP, MT>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    if (e != cudaSuccess) return set_error_cuda("cudaFuncSetAttribute(flash_attn)", e);
    configured.done(dev__);
  }
  dim3 grid((p.nq + BQ - 1) / BQ, p.heads, batch);
  launch_kernel(flash_attn_kernel<DP, MT>, dim3(grid), dim3(FA_THREADS), SMEM, stream, p);
  EMOTE_CHECK_LAUNCH("emote_attention_bf16");
  return 0;
}

template <int DP, int NK16>
static int launch_short_kv(const AttnDev& p, int batch, cudaStream_t stream) {
  constexpr int SMEM = (2 * NK16 * 16 + 2 * 64) * (DP + 8) * 2;
  static PerDeviceOnce configured;
  int dev__ = 0;
  if (configured.pending(&dev__)) {
    cudaError_t e =
        cudaFuncSetAttribute(short_kv_attn_kernel<DP, NK16>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    if (e != cudaSuccess) return set_error_cuda("cudaFuncSetAttribute(short_kv_attn)", e);
    configured.done(dev__);
  }
  const int n_qtiles = (p.nq + 63) / 64;
  // ~4 CTAs per SM worth of work items, each streaming several query tiles over one resident K/V
  int per_cta = (int)(((long long)n_qtiles * p.heads * batch + 148 * 4 - 1) / (148 * 4));
  if (per_cta < 1) per_cta = 1;
  if (per_cta > n_qtiles) per_cta = n_qtiles;
  dim3 grid((n_qtiles + per_cta - 1) / per_cta, p.heads, batch);
  launch_kernel(short_kv_attn_kernel<DP, NK16>, dim3(grid), dim3(FA_THREADS), SMEM, stream, p, per_cta);
  EMOTE_CHECK_LAUNCH("emote_attention_bf16");
  return 0;
}

template <int DP>
static int dispatch_short_kv(const AttnDev& p, int batch, cudaStream_t stream) {
  const int nk16 = (p.n0 + 15) / 16;
  if (nk16 <= 1) return launch_short_kv<DP, 1>(p, batch, stream);
  if (nk16 <= 2) return launch_short_kv<DP, 2>(p, batch, stream);
  if (nk16 <= 5) return launch_short_kv<DP, 5>(p, batch, stream);
  return launch_short_kv<DP, 8>(p, batch, stream);
}

template <int DP, int FP, int TA_WARPS>
static int launch_temporal_w(const op16* qkv, op16* out, int B, int F, int HW, int heads, int d,
                             float scale_log2, cudaStream_t stream) {
  constexpr int SMEM = TA_WARPS * 3 * FP * (DP + 8) * 2;
  static PerDeviceOnce configured;
  int dev__ = 0;
  if (configured.pending(&dev__)) {
    cudaError_t e = cudaFuncSetAttribute(temporal_attn_kernel<DP, FP, TA_WARPS>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    if (e != cudaSuccess) return set_error_cuda("cudaFuncSetAttribute(temporal_attn)", e);
    configured.done(dev__);
  }
  const int hgroups = (heads + TA_WARPS - 1) / TA_WARPS;
  const long long blocks = (long long)B * HW * hgroups;
  launch_kernel(temporal_attn_kernel<DP, FP, TA_WARPS>, dim3((unsigned)blocks), dim3(TA_WARPS * 32), SMEM, stream, qkv,
                out, F, HW, heads, d, scale_log2);
  EMOTE_CHECK_LAUNCH("emote_temporal_attention_bf16");
  return 0;
}

template <int DP, int FP>
static int launch_temporal(const op16* qkv, op16* out, int B, int F, int HW, int heads, int d,
                           float scale_log2, cudaStream_t stream) {
  if constexpr (8 * 3 * FP * (DP + 8) * 2 <= 200 * 1024) {   // eight heads' Q / K / V must fit one block's shared memory
    if (tuning(TUNE_TEMPORAL_WARPS, 4) == 8 && heads > 4)
      return launch_temporal_w<DP, FP, 8>(qkv, out, B, F, HW, heads, d, scale_log2, stream);
  }
  return launch_temporal_w<DP, FP, 4>(qkv, out, B, F, HW, heads, d, scale_log2, stream);
}

}  // namespace emote

using namespace emote;

#define EMOTE_DP_DISPATCH(dp, CALL)                         \
  switch (dp) {                                             \
    case 16: { constexpr int DPV = 16; CALL; } break;       \
    case 32: { constexpr int DPV = 32; CALL; } break;       \
    case 48: { constexpr int DPV = 48; CALL; } break;       \
    case 64: { constexpr int DPV = 64; CALL; } break;       \
    case 80: { constexpr int DPV = 80; CALL; } break;       \
    case 96: { constexpr int DPV = 96; CALL; } break;       \
    case 128: { constexpr int DPV = 128; CALL; } break;     \
    case 160: { constexpr int DPV = 160; CALL; } break;     \
    default: return set_error("attention: unsupported head_dim (max 160, multiple of 8)"); \
  }

static int padded_head_dim(int d) {
  const int c[] = {16, 32, 48, 64, 80, 96, 128, 160};
  for (int v : c)
    if (d <= v) return v;
  return -1;
}

extern "C" int emote_attention_bf16(const EmoteAttnArgs* a, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (!a || !a->q || !a->k0 || !a->v0 || !a->out) return set_error("emote_attention_bf16: null pointer");
  if (a->batch <= 0 || a->heads <= 0 || a->nq <= 0 || a->n0 <= 0 || a->n1 < 0)
    return set_error("emote_attention_bf16: bad sizes");
  if (a->head_dim % 8 != 0 || a->head_dim <= 0) return set_error("emote_attention_bf16: head_dim must be a multiple of 8");
  if (a->n1 > 0 && (!a->k1 || !a->v1)) return set_error("emote_attention_bf16: segment 1 pointers missing");
  if (a->batch > 65535 || a->heads > 65535) return set_error("emote_attention_bf16: grid too large");
  const int64_t strides[] = {a->q_batch_stride, a->q_row_stride, a->kv0_batch_stride, a->kv0_row_stride,
                             a->kv1_batch_stride, a->kv1_row_stride, a->o_batch_stride, a->o_row_stride};
  for (int64_t s : strides)
    if (s % 8 != 0) return set_error("emote_attention_bf16: strides must keep rows 16-byte aligned");
  AttnDev p{};
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
  const int dp = padded_head_dim(a->head_dim);
  if (a->n1 == 0 && a->n0 <= 128 && a->nq >= 256) {  // text / audio context: K/V resident in smem, queries streamed
    EMOTE_DP_DISPATCH(dp, return dispatch_short_kv<DPV>(p, a->batch, stream));
  }
  EMOTE_DP_DISPATCH(dp, return launch_flash<DPV>(p, a->batch, stream));
  return 0;
}

extern "C" int emote_temporal_attention_bf16(const void* qkv, void* out, int32_t B, int32_t F, int32_t HW,
                                             int32_t heads, int32_t head_dim, float scale, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (!qkv || !out || B <= 0 || F <= 0 || HW <= 0 || heads <= 0) return set_error("emote_temporal_attention_bf16: bad arguments");
  if (F > 32) return set_error("emote_temporal_attention_bf16: at most 32 frames per window");
  if (head_dim % 8 != 0) return set_error("emote_temporal_attention_bf16: head_dim must be a multiple of 8");
  const int dp = padded_head_dim(head_dim);
  const float sl2 = scale * 1.4426950408889634f;
  const op16* qp = (const op16*)qkv;
  op16* op = (op16*)out;
  if (F <= 16) {
    EMOTE_DP_DISPATCH(dp, return (launch_temporal<DPV, 16>(qp, op, B, F, HW, heads, head_dim, sl2, stream)));
  } else {
    EMOTE_DP_DISPATCH(dp, return (launch_temporal<DPV, 32>(qp, op, B, F, HW, heads, head_dim, sl2, stream)));
  }
  return 0;
}
