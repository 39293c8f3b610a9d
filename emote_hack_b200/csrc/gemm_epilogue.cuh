// Shared by the 1-CTA and the CTA-pair GEMM kernels: launch parameters and the per-tile epilogue
// (TMEM -> registers -> fused bias / residual / GEGLU -> global).
#pragma once
#include "common.cuh"

namespace emote {

constexpr int BM = 128;       // rows of the output tile owned by one CTA (one TMEM lane per row)
constexpr int BK = 64;        // 64 bf16 = 128 B = one swizzle row

struct GemmDev {
  int M, N, K;
  int num_kb;        // K blocks of 64 (over all taps)
  int kb_per_tap;    // C/64 in conv mode
  int taps;          // 1 or 9
  int H, W;          // conv image dims
  int bw, bh;        // TMA box extents in W and H (bw*bh*bnimg = 128)
  int bnimg, n_img;  // images per box / images in the tensor
  int patch;         // conv only: the 128 rows of a sub-tile are a 2-D patch bw x bh (x bnimg images) of the image instead
                     // of 128 consecutive pixels (W neither divides nor is a multiple of 128: 96x96, 48x48, 24x24 ... maps)
  int tiles_x, tiles_y;  // patch mode: patches per image row / column
  int subtiles;          // patch mode: number of 128-row sub-tiles (image groups x tiles_x x tiles_y)
  int tiles_m, tiles_n;
  const float* bias;
  const float* row_bias;
  int rows_per_group;
  const float* residual;
  int ldr;
  float out_scale;
  int geglu;
  int act_gelu;      // EMOTE_EPI_GELU: exact (erf) GELU of acc + bias, applied before the residual add
  int out_bf16;
  int ldc;
  void* out;
  float* colstats;   // optional fused GroupNorm statistics of the fp32 output, else null: one (sum, sum of squares) slot per
                     // 32-row quarter of a sub-tile and column, [sub-tiles * 4][N][2] fp32, written with plain stores
  int stats_rows;    // rows per statistics batch (validated on the host: a 32-row quarter never straddles two batches)
};


// ---- where the 128 rows of a sub-tile live.  `row_base` is the sub-tile index x 128: in the ordinary layout that IS the first
// global row; in patch mode (implicit-GEMM conv over maps whose rows do not pack into 128-pixel runs) sub-tile
// st = row_base / 128 is patch (tx, ty) of image group st / (tiles_x * tiles_y), and local row lr is pixel
// (x0 + lr % bw, y0 + (lr / bw) % bh) of image img0 + lr / (bw * bh).
struct TileOrigin { int img0, y0, x0; };
__device__ __forceinline__ TileOrigin tile_origin(const GemmDev& p, int row_base) {
  TileOrigin o;
  if (p.patch) {
    const int st = row_base / BM;
    const int tpg = p.tiles_x * p.tiles_y;
    const int grp = st / tpg, rem = st - grp * tpg;
    const int ty = rem / p.tiles_x;
    o.img0 = grp * p.bnimg;
    o.y0 = ty * p.bh;
    o.x0 = (rem - ty * p.tiles_x) * p.bw;
  } else {
    const int hw = p.H * p.W;
    o.img0 = row_base / hw;
    const int rem = row_base - o.img0 * hw;
    o.y0 = rem / p.W;
    o.x0 = rem - o.y0 * p.W;
  }
  return o;
}
// global output row of local row lr of the sub-tile, or -1 when that row does not exist
__device__ __forceinline__ int tile_global_row(const GemmDev& p, int row_base, int lr) {
  if (!p.patch) {
    const int r = row_base + lr;
    return r < p.M ? r : -1;
  }
  const TileOrigin o = tile_origin(p, row_base);
  const int pix = p.bw * p.bh;
  const int li = lr / pix, q = lr - li * pix;
  const int img = o.img0 + li;
  const int yy = q / p.bw;
  return img < p.n_img ? (img * p.H + o.y0 + yy) * p.W + o.x0 + (q - yy * p.bw) : -1;
}
// fp32 staging boxes [128 rows][32 floats] <-> global memory: 2-D tensor maps over [M, N] in the ordinary layout, 4-D maps
// over [n_img, H, W, N] with a {32, bw, bh, bnimg} box in patch mode (same order of the 128 rows in shared memory)
__device__ __forceinline__ void tma_store_4d(const void* desc, const void* smem_src, int32_t c0, int32_t c1, int32_t c2,
                                             int32_t c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(desc)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_prefetch_4d(const void* desc, int32_t c0, int32_t c1, int32_t c2, int32_t c3) {
  asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];" ::"l"(
                   reinterpret_cast<uint64_t>(desc)),
               "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tile_box_store(const GemmDev& p, const void* desc, const void* smem_src, int col,
                                               int row_base) {
  if (p.patch) {
    const TileOrigin o = tile_origin(p, row_base);
    tma_store_4d(desc, smem_src, col, o.x0, o.y0, o.img0);
  } else {
    tma_store_2d(desc, smem_src, col, row_base);
  }
}
__device__ __forceinline__ void tile_box_load(const GemmDev& p, void* smem_dst, const void* desc, uint64_t* bar, int col,
                                              int row_base) {
  if (p.patch) {
    const TileOrigin o = tile_origin(p, row_base);
    tma_load_4d(smem_dst, desc, bar, col, o.x0, o.y0, o.img0);
  } else {
    tma_load_2d(smem_dst, desc, bar, col, row_base);
  }
}
__device__ __forceinline__ void tile_box_prefetch(const GemmDev& p, const void* desc, int col, int row_base) {
  if (p.patch) {
    const TileOrigin o = tile_origin(p, row_base);
    tma_prefetch_4d(desc, col, o.x0, o.y0, o.img0);
  } else {
    tma_prefetch_2d(desc, col, row_base);
  }
}

// OUT_MODE 1: bulk-store the staged bf16 tile (issued by one thread; TMA clips rows >= M and columns >= N).
template <int BN>
__device__ __forceinline__ void store_bf16_boxes(const CUtensorMap* tmC, const uint8_t* stage, const GemmDev& p, int tn,
                                                 int row_base) {
  tma_store_2d(tmC, stage, p.geglu ? tn * (BN / 2) : tn * BN, row_base);
}

// One output tile of one CTA.  `tbase` = TMEM address of the accumulator stage for this warp's lane quarter,
// `row_base` = first global row of the tile, (`n0`, `tn`) = first column / column-tile index.  The caller has NOT yet
// waited for the accumulator: `wait_full()` is invoked after the residual prefetch has been issued.
// OUT_MODE 0: results go straight to global memory (8-byte / 4-byte stores).
// OUT_MODE 1: bf16 results are staged in shared memory (`stage`, dense [128][BN or BN/2] bf16) for one bulk tensor store
//             per tile issued by the caller — full-line L2 writes instead of 16-byte partial-sector stores.  (Narrow
//             swizzled boxes — 32 columns / 64B swizzle, 16 / 32B — remove the 4-way bank conflicts of the dense layout
//             but measured slower: GEGLU K=320 241 -> 315 us; the bulk store prefers one wide box.)
// OUT_MODE 2: fp32 results; `stage` holds BN/32 boxes of [128 rows][32 floats] in the TMA 128-byte swizzle.  When
//             p.residual is set the caller has TMA-loaded the fp32 residual tile into the same boxes: the epilogue
//             adds in place and the caller bulk-stores the boxes (the residual stream never touches the LSU path).
// COL0 / NCOLS (multiples of 32; linear epilogue only): restrict the call to tile columns [COL0, COL0 + NCOLS) — the split
// mode-2 pipeline finishes and stores a tile in two column halves.
template <int BN, int EPI_WARPS, bool HAS_ADD, int OUT_MODE, int COL0 = 0, int NCOLS = BN, class WaitFull>
__device__ __forceinline__ void gemm_epilogue_tile(const GemmDev& p, uint32_t tbase, int row_base, int n0, int tn,
                                                   int quarter, int part, int lane, void* stage_, WaitFull wait_full) {
  constexpr bool TMA_OUT = OUT_MODE == 1;
  constexpr int NP = EPI_WARPS / 4;   // warps per lane quarter
  const int g = lane >> 2, t = lane & 3;
  // this thread's rows: local rows quarter*32 + g + 8*i, i = 0..3 (global row -1 = not there)
  int rows[4];
  bool rok[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    rows[i] = tile_global_row(p, row_base, quarter * 32 + g + 8 * i);
    rok[i] = rows[i] >= 0;
  }
  const int first_row = tile_global_row(p, row_base, 0);   // selects the tile's bias group / statistics batch

  if (!p.geglu) {
    constexpr int NCT = NCOLS / 8;                  // 8-column chunks in the (sub-)tile
    constexpr int NCH = NCT / NP;                   // contiguous chunks per warp
    static_assert(NCT % NP == 0, "tile columns must split evenly over the warps of a lane quarter");
    const int c_first = COL0 / 8 + part * NCH;
    size_t ooff[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) ooff[i] = (size_t)(rok[i] ? rows[i] : 0) * p.ldc;
    // ---- residual / per-sample bias of this warp's whole column span, issued before the accumulator wait
    float add[HAS_ADD ? NCH : 1][8];
    if constexpr (HAS_ADD) {
      const bool has_res = p.residual != nullptr, has_rb = p.row_bias != nullptr;
#pragma unroll
      for (int ci = 0; ci < NCH; ++ci) {
        const int col = n0 + (c_first + ci) * 8 + 2 * t;
        const bool c0ok = col < p.N, c1ok = col + 1 < p.N;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int row = rows[i];
          float v0 = 0.f, v1 = 0.f;
          if (rok[i]) {
            if (has_res) {
              const float* rp = p.residual + (size_t)row * p.ldr + col;
              if (c1ok) { const float2 r2 = *reinterpret_cast<const float2*>(rp); v0 = r2.x; v1 = r2.y; }
              else if (c0ok) v0 = rp[0];
            }
            if (has_rb) {
              const float* bp = p.row_bias + (size_t)(row / p.rows_per_group) * p.N + col;
              if (c1ok) { const float2 r2 = *reinterpret_cast<const float2*>(bp); v0 += r2.x; v1 += r2.y; }
              else if (c0ok) v0 += bp[0];
            }
          }
          add[ci][2 * i] = v0;
          add[ci][2 * i + 1] = v1;
        }
      }
    }
    // ---- column bias (+ the tile's per-sample bias in mode 2) of this warp's span: loaded here, before the accumulator
    // wait, so the global-load latency is off the per-chunk critical path (it used to be paid once per 8-column chunk)
    float bia[NCH][2];
#pragma unroll
    for (int ci = 0; ci < NCH; ++ci) {
      const int col = n0 + (c_first + ci) * 8 + 2 * t;
      const bool c0ok = col < p.N, c1ok = col + 1 < p.N;
      float b0 = 0.f, b1 = 0.f;
      if (p.bias) {
        if (c0ok) b0 = __ldg(p.bias + col);
        if (c1ok) b1 = __ldg(p.bias + col + 1);
      }
      if constexpr (OUT_MODE == 2) {
        if (p.row_bias && first_row >= 0) {  // host guarantees the 128 rows of a tile share one group
          const float* rb = p.row_bias + (size_t)(first_row / p.rows_per_group) * p.N + col;
          if (c0ok) b0 += __ldg(rb);
          if (c1ok) b1 += __ldg(rb + 1);
        }
      }
      bia[ci][0] = b0;
      bia[ci][1] = b1;
    }
    uint32_t acc[2][8];
    auto issue = [&](int ci, uint32_t (&a)[8]) {
      const uint32_t col_t = static_cast<uint32_t>((c_first + ci) * 8);
      uint32_t lo[4], hi[4];
      tmem_ld_16x256b_x1(tbase + col_t, lo);
      tmem_ld_16x256b_x1(tbase + (16u << 16) + col_t, hi);
#pragma unroll
      for (int k = 0; k < 4; ++k) { a[k] = lo[k]; a[4 + k] = hi[k]; }
    };
    auto finish = [&](int ci, const uint32_t (&a)[8]) {
      const int col = n0 + (c_first + ci) * 8 + 2 * t;
      const bool c0ok = col < p.N, c1ok = col + 1 < p.N;
      const float b0 = bia[ci][0], b1 = bia[ci][1];
      const bool stats = p.colstats != nullptr;   // fused GroupNorm statistics of the values being written
      float cs0 = 0.f, cq0 = 0.f, cs1 = 0.f, cq1 = 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        // rows g, g+8 come from the low 16-lane load, rows g+16, g+24 from the high one
        const int ri = (i >> 1) * 4 + (i & 1) * 2;
        float v0 = __uint_as_float(a[ri]) + b0, v1 = __uint_as_float(a[ri + 1]) + b1;
        if (p.act_gelu) { v0 = gelu_erf_fast(v0); v1 = gelu_erf_fast(v1); }
        if constexpr (HAS_ADD) { v0 += add[ci][2 * i]; v1 += add[ci][2 * i + 1]; }
        float* sp = nullptr;
        if constexpr (OUT_MODE == 2) {
          const int lr = quarter * 32 + g + 8 * i, lc = (c_first + ci) * 8 + 2 * t;
          sp = reinterpret_cast<float*>(stage_) + (lc >> 5) * (BM * 32) + lr * 32 +
               ((((lc & 31) >> 2) ^ (lr & 7)) << 2) + (lc & 3);
          if (p.residual) {
            const float2 r2 = *reinterpret_cast<const float2*>(sp);
            v0 += r2.x; v1 += r2.y;
          }
        }
        v0 *= p.out_scale; v1 *= p.out_scale;
        if (stats && rok[i]) {
          cs0 += v0; cq0 = fmaf(v0, v0, cq0);
          cs1 += v1; cq1 = fmaf(v1, v1, cq1);
        }
        if constexpr (OUT_MODE == 2) {
          *reinterpret_cast<float2*>(sp) = make_float2(v0, v1);
        } else if constexpr (TMA_OUT) {
          const int lr = quarter * 32 + g + 8 * i, lc = (c_first + ci) * 8 + 2 * t;
          *reinterpret_cast<uint32_t*>(reinterpret_cast<op16*>(stage_) + lr * BN + lc) = pack_op16x2(v0, v1);
        } else if (rok[i] && c0ok) {
          if (p.out_bf16) {
            op16* o = reinterpret_cast<op16*>(p.out) + ooff[i] + col;
            if (c1ok) *reinterpret_cast<uint32_t*>(o) = pack_op16x2(v0, v1);
            else o[0] = float2op16(v0);
          } else {
            float* o = reinterpret_cast<float*>(p.out) + ooff[i] + col;
            if (c1ok) *reinterpret_cast<float2*>(o) = make_float2(v0, v1);
            else o[0] = v0;
          }
        }
      }
      if (stats) {
        // this thread summed its 4 rows; fold the 8 row groups of the warp (lanes differing in g) in a fixed order, then
        // one lane per column pair stores the 32-row partial into the quarter's slot: no atomics, no zero-fill, and the
        // statistics (hence the whole network) are bit-reproducible from run to run
#pragma unroll
        for (int off = 4; off < 32; off <<= 1) {
          cs0 += __shfl_xor_sync(0xffffffffu, cs0, off);
          cq0 += __shfl_xor_sync(0xffffffffu, cq0, off);
          cs1 += __shfl_xor_sync(0xffffffffu, cs1, off);
          cq1 += __shfl_xor_sync(0xffffffffu, cq1, off);
        }
        if (g == 0 && first_row >= 0) {   // (the second CTA of the last pair may own a sub-tile that does not exist)
          float* cs = p.colstats + ((size_t)(row_base / 32 + quarter) * p.N + col) * 2;
          if (c1ok) *reinterpret_cast<float4*>(cs) = make_float4(cs0, cq0, cs1, cq1);
          else if (c0ok) *reinterpret_cast<float2*>(cs) = make_float2(cs0, cq0);
        }
      }
    };
    wait_full();
    issue(0, acc[0]);
#pragma unroll
    for (int ci = 0; ci < NCH; ++ci) {
      tmem_ld_wait();
      if (ci + 1 < NCH) issue(ci + 1, acc[(ci + 1) & 1]);
      finish(ci, acc[ci & 1]);
    }
  } else {
    // GEGLU: tile columns [0, BN/2) hold the value half, [BN/2, BN) the gate half of the same BN/2 output
    // features (weights are packed that way); out = (value + b_v) * gelu_erf(gate + b_g).
    constexpr int HALF = BN / 2;
    constexpr int NCT = HALF / 8;
    constexpr int NCH = (NCT + NP - 1) / NP;   // round-robin over the warps of the quarter (may be uneven)
    const int n_out = p.N / 2;
    size_t ooff[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) ooff[i] = (size_t)(rok[i] ? rows[i] : 0) * p.ldc;
    uint32_t av[2][8], ag[2][8];
    float bb[NCH][4];   // value / gate biases of this warp's chunks, loaded before the accumulator wait
#pragma unroll
    for (int ci = 0; ci < NCH; ++ci) {
      const int cc = part + ci * NP;
      bb[ci][0] = bb[ci][1] = bb[ci][2] = bb[ci][3] = 0.f;
      if (cc < NCT && p.bias) {
        const int tc = cc * 8 + 2 * t;
        bb[ci][0] = __ldg(p.bias + n0 + tc); bb[ci][1] = __ldg(p.bias + n0 + tc + 1);
        bb[ci][2] = __ldg(p.bias + n0 + HALF + tc); bb[ci][3] = __ldg(p.bias + n0 + HALF + tc + 1);
      }
    }
    auto issue = [&](int ci, uint32_t (&v)[8], uint32_t (&gt)[8]) {
      const int cc = part + ci * NP;
      if (cc < NCT) {
        const uint32_t col_t = static_cast<uint32_t>(cc * 8);
        uint32_t lo[4], hi[4], glo[4], ghi[4];
        tmem_ld_16x256b_x1(tbase + col_t, lo);
        tmem_ld_16x256b_x1(tbase + (16u << 16) + col_t, hi);
        tmem_ld_16x256b_x1(tbase + HALF + col_t, glo);
        tmem_ld_16x256b_x1(tbase + (16u << 16) + HALF + col_t, ghi);
#pragma unroll
        for (int k = 0; k < 4; ++k) { v[k] = lo[k]; v[4 + k] = hi[k]; gt[k] = glo[k]; gt[4 + k] = ghi[k]; }
      }
    };
    auto finish = [&](int ci, const uint32_t (&v)[8], const uint32_t (&gt)[8], const float (&b)[4]) {
      const int cc = part + ci * NP;
      if (cc < NCT) {
        const int ocol = tn * HALF + cc * 8 + 2 * t;
        const bool cok = ocol < n_out;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int ri = (i >> 1) * 4 + (i & 1) * 2;
          const float a0 = __uint_as_float(v[ri]) + b[0], a1 = __uint_as_float(v[ri + 1]) + b[1];
          const float q0 = __uint_as_float(gt[ri]) + b[2], q1 = __uint_as_float(gt[ri + 1]) + b[3];
          const uint32_t packed = pack_op16x2(a0 * gelu_sig(q0), a1 * gelu_sig(q1));
          if constexpr (TMA_OUT) {
            const int lr = quarter * 32 + g + 8 * i, lc = cc * 8 + 2 * t;
            *reinterpret_cast<uint32_t*>(reinterpret_cast<op16*>(stage_) + lr * HALF + lc) = packed;
          } else if (rok[i] && cok)
            *reinterpret_cast<uint32_t*>(reinterpret_cast<op16*>(p.out) + ooff[i] + ocol) = packed;
        }
      }
    };
    wait_full();
    issue(0, av[0], ag[0]);
#pragma unroll
    for (int ci = 0; ci < NCH; ++ci) {
      tmem_ld_wait();
      if (ci + 1 < NCH) issue(ci + 1, av[(ci + 1) & 1], ag[(ci + 1) & 1]);
      finish(ci, av[ci & 1], ag[ci & 1], bb[ci]);
    }
  }
}

}  // namespace emote
