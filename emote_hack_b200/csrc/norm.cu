// HBM-bound normalisation kernels over tokens-major fp32 activations (coalesced, vectorised):
//   * GroupNorm statistics (sum / sum of squares per (group batch, group)) with channel-concat support
//   * GroupNorm apply (+SiLU) -> bf16 GEMM/conv operand (+ optional raw bf16 copy for the shortcut conv)
//   * LayerNorm (+ additive temporal positional table) -> bf16 GEMM operand
// Reference ops: nn.GroupNorm resnet.py:180,191 / attention.py:124 / motion_module.py:147 /
// unet_controlnet.py:476; nn.LayerNorm attention.py:204-232, motion_module.py:208-213.
#include "common.cuh"
#include "emote_b200.h"
#include "host_utils.h"

namespace emote {

constexpr int GN_MAX_PASS = 4;
constexpr int GN_MAX_GROUPS = 64;

// blockDim = (PX, TY); thread (px, ty) owns channel pairs px + pass*PX and rows ty, ty+TY, ...
// Blocks walk the tensor BACKWARDS (last rows first): the producer (a GEMM epilogue) wrote the rows in ascending order,
// so the tail is what is still resident in L2; gn_apply then walks forwards over the rows this kernel touched last.
template <int PASSES>
__global__ void __launch_bounds__(512) gn_stats_kernel(const float* __restrict__ x, int C_src, int c_offset, int cpg,
                                                       int groups, long long rows_per_batch, int rows_per_block,
                                                       double* __restrict__ sums) {
  pdl_prologue();
  // fp64 accumulation end to end: E[x^2] - mean^2 cancels badly in fp32 when |mean| >> std, and the atomics'
  // ordering would otherwise leak ~1e-6 run-to-run noise into every normalised value.
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;

  const int batch = gridDim.y - 1 - blockIdx.y;
  const long long r0 = (long long)(gridDim.x - 1 - blockIdx.x) * rows_per_block;
  long long r1 = r0 + rows_per_block;
  if (r1 > rows_per_batch) r1 = rows_per_batch;
  const float* xb = x + ((long long)batch * rows_per_batch) * C_src;

  double s[PASSES], q[PASSES];
#pragma unroll
  for (int i = 0; i < PASSES; ++i) s[i] = q[i] = 0.0;

  // 8 rows in flight per thread and pass (independent 8-byte loads); the 16 values are pre-summed in fp32 in a fixed
  // order (deterministic), everything after that is fp64.
  const long long stride = blockDim.y;
  for (long long r = r0 + threadIdx.y; r < r1; r += 8 * stride) {
#pragma unroll
    for (int ps = 0; ps < PASSES; ++ps) {
      const int cp = threadIdx.x + ps * blockDim.x;
      float2 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const long long rr = r + u * stride;
        v[u] = (rr < r1) ? __ldg(reinterpret_cast<const float2*>(xb + rr * C_src) + cp) : make_float2(0.f, 0.f);
      }
      float sf = 0.f, qf = 0.f;
#pragma unroll
      for (int u = 0; u < 8; u += 2) {
        sf += (v[u].x + v[u].y) + (v[u + 1].x + v[u + 1].y);
        qf += (v[u].x * v[u].x + v[u].y * v[u].y) + (v[u + 1].x * v[u + 1].x + v[u + 1].y * v[u + 1].y);
      }
      s[ps] += (double)sf;
      q[ps] += (double)qf;
    }
  }
  // block reduction without shared-memory fp64 atomics (those are CAS loops and dominated the kernel): every thread
  // parks its partials, then one thread per (group, statistic) sums the <= cpg/2 * blockDim.y slots of its group
  __shared__ double part[2][PASSES][512];
#pragma unroll
  for (int ps = 0; ps < PASSES; ++ps) {
    part[0][ps][tid] = s[ps];
    part[1][ps][tid] = q[ps];
  }
  __syncthreads();
  if (tid < groups * 2) {
    const int g = tid >> 1, which = tid & 1;
    int c_lo = g * cpg, c_hi = c_lo + cpg;            // channel range of the group in the concatenated tensor
    if (c_lo < c_offset) c_lo = c_offset;
    if (c_hi > c_offset + C_src) c_hi = c_offset + C_src;
    double acc = 0.0;
    for (int c = c_lo; c < c_hi; c += 2) {
      const int cp = (c - c_offset) >> 1;
      const int ps = cp / (int)blockDim.x, tx = cp - ps * (int)blockDim.x;
      for (int ty = 0; ty < (int)blockDim.y; ++ty) acc += part[which][ps][ty * blockDim.x + tx];
    }
    if (c_lo < c_hi) atomicAdd(&sums[(long long)batch * groups * 2 + tid], acc);
  }
}

// Group statistics from the per-quarter column slots a GEMM epilogue stored (EmoteGemmArgs.colstats): the per-slot walk
// (kept as the A/B partner of the flat fold below and for > 2^30 pairs per group).
__global__ void __launch_bounds__(256) gn_colstats_reduce_kernel(const float* __restrict__ slots, int C_src, int c_offset,
                                                                 int cpg, int groups, int slots_per_batch, int n_batches,
                                                                 double* __restrict__ sums, int overwrite) {
  pdl_prologue();
  // one block per (batch, group): thread t walks slots t, t + 256, ... of the batch over the group's columns of this
  // source, in a fixed order; fp64 from the first addition on; fixed-order tree over the block -> deterministic
  const int batch = blockIdx.x / groups, g = blockIdx.x - batch * groups;
  int c_lo = g * cpg, c_hi = c_lo + cpg;
  if (c_lo < c_offset) c_lo = c_offset;
  if (c_hi > c_offset + C_src) c_hi = c_offset + C_src;
  const int nc = c_hi - c_lo;
  double* o = sums + ((long long)batch * groups + g) * 2;
  if (nc <= 0) {
    if (overwrite && threadIdx.x < 2) o[threadIdx.x] = 0.0;
    return;
  }
  double s = 0.0, q = 0.0;
  const float* base = slots + ((long long)batch * slots_per_batch * C_src + (c_lo - c_offset)) * 2;
  for (int sl = threadIdx.x; sl < slots_per_batch; sl += blockDim.x) {
    const float2* row = reinterpret_cast<const float2*>(base + (long long)sl * C_src * 2);
    for (int c = 0; c < nc; ++c) {
      const float2 v = __ldg(row + c);
      s += (double)v.x;
      q += (double)v.y;
    }
  }
  __shared__ double rs[256], rq[256];
  rs[threadIdx.x] = s;
  rq[threadIdx.x] = q;
  __syncthreads();
  for (int off = 128; off > 0; off >>= 1) {
    if (threadIdx.x < off) {
      rs[threadIdx.x] += rs[threadIdx.x + off];
      rq[threadIdx.x] += rq[threadIdx.x + off];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    if (overwrite) { o[0] = rs[0]; o[1] = rq[0]; }
    else { o[0] += rs[0]; o[1] += rq[0]; }
  }
}

// Same result slots, flat walk: the (slot, column) pairs of one (batch, group) are numbered e = slot * nc + column and
// thread t takes e = t, t + NT, ...; eight independent loads are in flight per thread before the first fp64 addition, so a
// 2048-slot batch costs 3 rounds of memory latency with NT = 1024 (the per-slot walk above: one round per slot and column
// batch, ~24).  Fixed mapping + fixed-order warp / block folds -> deterministic.
template <int NT>
__global__ void __launch_bounds__(NT) gn_colstats_reduce_flat_kernel(const float* __restrict__ slots, int C_src,
                                                                     int c_offset, int cpg, int groups,
                                                                     int slots_per_batch, double* __restrict__ sums,
                                                                     int overwrite) {
  pdl_prologue();
  const int batch = blockIdx.x / groups, g = blockIdx.x - batch * groups;
  int c_lo = g * cpg, c_hi = c_lo + cpg;
  if (c_lo < c_offset) c_lo = c_offset;
  if (c_hi > c_offset + C_src) c_hi = c_offset + C_src;
  const int nc = c_hi - c_lo;
  double* o = sums + ((long long)batch * groups + g) * 2;
  if (nc <= 0) {
    if (overwrite && threadIdx.x < 2) o[threadIdx.x] = 0.0;
    return;
  }
  const float2* base =
      reinterpret_cast<const float2*>(slots) + ((long long)batch * slots_per_batch * C_src + (c_lo - c_offset));
  const int total = slots_per_batch * nc;   // < 2^30 (checked by the launcher): 32-bit index arithmetic
  double s = 0.0, q = 0.0;
  for (int e0 = threadIdx.x; e0 < total; e0 += 8 * NT) {
    // branch-free: out-of-range elements re-read the last valid one and are masked after all eight loads were issued
    // (a guarded load makes the compiler convert each value right behind its load, serialising the latencies)
    float2 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int e = min(e0 + u * NT, total - 1);
      const int sl = e / nc;
      v[u] = __ldg(base + (long long)sl * C_src + (e - sl * nc));
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const bool ok = e0 + u * NT < total;
      s += ok ? (double)v[u].x : 0.0;
      q += ok ? (double)v[u].y : 0.0;
    }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    s += __shfl_down_sync(0xffffffffu, s, off);
    q += __shfl_down_sync(0xffffffffu, q, off);
  }
  __shared__ double ws[NT / 32], wq[NT / 32];
  if ((threadIdx.x & 31) == 0) {
    ws[threadIdx.x >> 5] = s;
    wq[threadIdx.x >> 5] = q;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double S = 0.0, Q = 0.0;
#pragma unroll
    for (int w = 0; w < NT / 32; ++w) {
      S += ws[w];
      Q += wq[w];
    }
    if (overwrite) { o[0] = S; o[1] = Q; }
    else { o[0] += S; o[1] += Q; }
  }
}

// y = x * scale[c] + shift[c] with scale = rstd*gamma, shift = beta - mean*rstd*gamma staged in shared memory per
// block; blockDim = (octets per pass, rows in parallel); GN_APPLY_U rows per thread are in flight per batch.
constexpr int GN_APPLY_U = 4;  // rows per thread in flight
__global__ void __launch_bounds__(256) gn_apply_kernel(const float* __restrict__ x, int C_src, int c_offset, int C_total,
                                                       int cpg, int groups, long long rows_per_batch, int rows_per_block,
                                                       const double* __restrict__ sums, const float* __restrict__ gamma,
                                                       const float* __restrict__ beta, float eps, int act_silu,
                                                       op16* __restrict__ out,
                                                       op16* __restrict__ raw_out) {
  pdl_prologue();
  extern __shared__ __align__(16) float gn_smem[];
  float* s_scale = gn_smem;
  float* s_shift = gn_smem + C_src;
  const int batch = blockIdx.y;
  {
    // mean / rstd once per group (fp64 divide + square root), then one fp32 multiply-add pair per channel
    __shared__ float g_mean[GN_MAX_GROUPS], g_rstd[GN_MAX_GROUPS];
    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    const int nthr = blockDim.x * blockDim.y;
    const double cnt = (double)rows_per_batch * (double)cpg;
    for (int g = tid; g < groups; g += nthr) {
      const double sm = sums[((long long)batch * groups + g) * 2];
      const double sq = sums[((long long)batch * groups + g) * 2 + 1];
      const double mean = sm / cnt;
      double var = sq / cnt - mean * mean;
      if (var < 0.0) var = 0.0;
      g_mean[g] = (float)mean;
      g_rstd[g] = (float)(1.0 / sqrt(var + (double)eps));
    }
    __syncthreads();
    for (int c = tid; c < C_src; c += nthr) {
      const int ct = c_offset + c;
      const int g = ct / cpg;
      const float sc = g_rstd[g] * gamma[ct];
      s_scale[c] = sc;
      s_shift[c] = beta[ct] - g_mean[g] * sc;
    }
    __syncthreads();
  }
  const int opr = C_src >> 3;
  const long long r0 = (long long)blockIdx.x * rows_per_block;
  long long r1 = r0 + rows_per_block;
  if (r1 > rows_per_batch) r1 = rows_per_batch;
  const long long row_base = (long long)batch * rows_per_batch;
  const long long stride = blockDim.y;
  for (long long r = r0 + threadIdx.y; r < r1; r += GN_APPLY_U * stride) {
    for (int o = threadIdx.x; o < opr; o += blockDim.x) {
      const int c0 = o << 3;
      float4 a[GN_APPLY_U], b[GN_APPLY_U];
#pragma unroll
      for (int u = 0; u < GN_APPLY_U; ++u) {   // all loads of the batch are in flight before the first store
        const long long rr = r + u * stride;
        if (rr < r1) {
          const float* xr = x + (row_base + rr) * C_src + c0;
          a[u] = __ldg(reinterpret_cast<const float4*>(xr));
          b[u] = __ldg(reinterpret_cast<const float4*>(xr + 4));
        }
      }
      const float4 sa = *reinterpret_cast<const float4*>(s_scale + c0);
      const float4 sb = *reinterpret_cast<const float4*>(s_scale + c0 + 4);
      const float4 ha = *reinterpret_cast<const float4*>(s_shift + c0);
      const float4 hb = *reinterpret_cast<const float4*>(s_shift + c0 + 4);
      const float sc[8] = {sa.x, sa.y, sa.z, sa.w, sb.x, sb.y, sb.z, sb.w};
      const float sh[8] = {ha.x, ha.y, ha.z, ha.w, hb.x, hb.y, hb.z, hb.w};
#pragma unroll
      for (int u = 0; u < GN_APPLY_U; ++u) {
        const long long rr = r + u * stride;
        if (rr < r1) {
          const float v[8] = {a[u].x, a[u].y, a[u].z, a[u].w, b[u].x, b[u].y, b[u].z, b[u].w};
          float y[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float tt = fmaf(v[j], sc[j], sh[j]);
            y[j] = act_silu ? silu_f(tt) : tt;
          }
          uint4 w;
          w.x = pack_op16x2(y[0], y[1]); w.y = pack_op16x2(y[2], y[3]);
          w.z = pack_op16x2(y[4], y[5]); w.w = pack_op16x2(y[6], y[7]);
          *reinterpret_cast<uint4*>(out + (row_base + rr) * C_total + c_offset + c0) = w;
          if (raw_out) {
            uint4 q;
            q.x = pack_op16x2(v[0], v[1]); q.y = pack_op16x2(v[2], v[3]);
            q.z = pack_op16x2(v[4], v[5]); q.w = pack_op16x2(v[6], v[7]);
            *reinterpret_cast<uint4*>(raw_out + (row_base + rr) * C_total + c_offset + c0) = q;
          }
        }
      }
    }
  }
}

constexpr int LN_MAX_V = 32;  // float2 per lane -> C <= 2048

// One warp per row, NV float2 per lane (C = 64*NV).  NV is a template parameter so the row lives in exactly
// 2*NV registers (occupancy) and all loads of a row are issued back to back (bytes in flight).
template <int NV>
__global__ void __launch_bounds__(256) layernorm_kernel(const float* __restrict__ x, long long M, int C_rt,
                                                        const float* __restrict__ gamma, const float* __restrict__ beta,
                                                        float eps, const float* __restrict__ pe, int pe_rows_per_frame,
                                                        int pe_frames, op16* __restrict__ out,
                                                        float* __restrict__ out_f32) {
  pdl_prologue();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // rows are visited last-to-first: the producing GEMM wrote them in ascending order, so the tail is still in L2, and
  // the rows written here last (the head) are the first ones the consuming GEMM loads.
  const long long row = M - 1 - ((long long)blockIdx.x * (blockDim.x >> 5) + warp);
  if (row < 0) return;
  const int nv = (NV > 0) ? NV : (C_rt >> 6);
  constexpr int CAP = (NV > 0) ? NV : LN_MAX_V;
  const int C = nv << 6;
  const float2* xr = reinterpret_cast<const float2*>(x + row * C);
  float2 v[CAP];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < CAP; ++i) {
    if (i < nv) {
      v[i] = __ldg(xr + lane + i * 32);
      s += v[i].x + v[i].y;
    }
  }
  const float mean = warp_sum(s) / (float)C;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < CAP; ++i) {
    if (i < nv) {
      const float a = v[i].x - mean, b = v[i].y - mean;
      q += a * a + b * b;
    }
  }
  const float rstd = rsqrtf(warp_sum(q) / (float)C + eps);
  const float* per = nullptr;
  if (pe) per = pe + (long long)((row / pe_rows_per_frame) % pe_frames) * C;
  uint32_t* o = reinterpret_cast<uint32_t*>(out + row * C);
#pragma unroll
  for (int i = 0; i < CAP; ++i) {
    if (i < nv) {
      const int c = 2 * (lane + i * 32);
      const float2 g = __ldg(reinterpret_cast<const float2*>(gamma + c));
      const float2 b = __ldg(reinterpret_cast<const float2*>(beta + c));
      float y0 = (v[i].x - mean) * rstd * g.x + b.x;
      float y1 = (v[i].y - mean) * rstd * g.y + b.y;
      if (per) {
        const float2 p2 = __ldg(reinterpret_cast<const float2*>(per + c));
        y0 += p2.x;
        y1 += p2.y;
      }
      if (out) o[lane + i * 32] = pack_op16x2(y0, y1);
      if (out_f32) *reinterpret_cast<float2*>(out_f32 + row * C + c) = make_float2(y0, y1);
    }
  }
}

__global__ void softmax_rows_kernel(const float* __restrict__ s, int N, float scale, op16* __restrict__ out) {
  pdl_prologue();
  // one block per row
  const long long row = blockIdx.x;
  const float* sr = s + row * N;
  __shared__ float red[32];
  float m = -INFINITY;
  for (int i = threadIdx.x; i < N; i += blockDim.x) m = fmaxf(m, sr[i] * scale);
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : -INFINITY;
    t = warp_max(t);
    if (threadIdx.x == 0) red[0] = t;
  }
  __syncthreads();
  m = red[0];
  __syncthreads();
  float sum = 0.f;
  for (int i = threadIdx.x; i < N; i += blockDim.x) sum += __expf(sr[i] * scale - m);
  sum = warp_sum(sum);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sum;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    t = warp_sum(t);
    if (threadIdx.x == 0) red[0] = t;
  }
  __syncthreads();
  const float inv = 1.0f / red[0];
  op16* o = out + row * N;
  for (int i = threadIdx.x; i < N; i += blockDim.x) o[i] = float2op16(__expf(sr[i] * scale - m) * inv);
}

}  // namespace emote

using namespace emote;

static int gn_check(int C_src, int c_offset, int C_total, int groups, const char* who) {
  if (groups <= 0 || groups > GN_MAX_GROUPS || C_total % groups != 0) return set_error(who);
  const int cpg = C_total / groups;
  if (cpg % 2 != 0 || c_offset % 8 != 0 || C_src % 8 != 0 || c_offset + C_src > C_total || C_total % 8 != 0)
    return set_error(who);
  return 0;
}

extern "C" int emote_gn_stats(const float* x, int32_t C_src, int32_t c_offset, int32_t C_total, int32_t groups,
                              int64_t rows_per_batch, int32_t n_batches, double* sums, int32_t zero_first,
                              void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (!x || !sums || rows_per_batch <= 0 || n_batches <= 0) return set_error("emote_gn_stats: bad arguments");
  if (gn_check(C_src, c_offset, C_total, groups, "emote_gn_stats: unsupported channel/group configuration")) return EMOTE_ERR_INVALID;
  if (zero_first) {
    cudaError_t e = cudaMemsetAsync(sums, 0, sizeof(double) * 2 * groups * (size_t)n_batches, stream);
    if (e != cudaSuccess) return set_error_cuda("emote_gn_stats memset", e);
  }
  const int P = C_src / 2;
  int passes = 1;
  while (passes <= GN_MAX_PASS && !(P % passes == 0 && P / passes <= 512)) ++passes;
  if (passes > GN_MAX_PASS) return set_error("emote_gn_stats: C_src too wide");
  const int PX = P / passes;
  int TY = 512 / PX;
  if (TY < 1) TY = 1;
  // one batch of loads covers 8*TY rows; small tensors get one batch per block (more blocks than SMs), large ones up
  // to 4 batches per block to amortise the block reduction and its atomics
  const long long total_rows = rows_per_batch * (long long)n_batches;
  long long k = total_rows / (8LL * TY * 1184);
  if (k < 1) k = 1;
  if (k > 4) k = 4;
  const int rows_per_block = (int)(8 * TY * k);
  const long long chunks = (rows_per_batch + rows_per_block - 1) / rows_per_block;
  dim3 grid((unsigned)chunks, (unsigned)n_batches), block(PX, TY);
  const int cpg = C_total / groups;
  switch (passes) {
    case 1: launch_kernel(gn_stats_kernel<1>, dim3(grid), dim3(block), 0, stream, x, C_src, c_offset, cpg, groups, rows_per_batch, rows_per_block, sums); break;
    case 2: launch_kernel(gn_stats_kernel<2>, dim3(grid), dim3(block), 0, stream, x, C_src, c_offset, cpg, groups, rows_per_batch, rows_per_block, sums); break;
    case 3: launch_kernel(gn_stats_kernel<3>, dim3(grid), dim3(block), 0, stream, x, C_src, c_offset, cpg, groups, rows_per_batch, rows_per_block, sums); break;
    default: launch_kernel(gn_stats_kernel<4>, dim3(grid), dim3(block), 0, stream, x, C_src, c_offset, cpg, groups, rows_per_batch, rows_per_block, sums); break;
  }
  EMOTE_CHECK_LAUNCH("emote_gn_stats");
  return 0;
}

extern "C" int emote_gn_colstats_reduce(const float* colstats, int32_t C_src, int32_t c_offset, int32_t C_total,
                                        int32_t groups, int32_t slots_per_batch, int32_t n_batches, double* sums,
                                        int32_t zero_first, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (!colstats || !sums || slots_per_batch <= 0 || n_batches <= 0)
    return set_error("emote_gn_colstats_reduce: bad arguments");
  if (gn_check(C_src, c_offset, C_total, groups, "emote_gn_colstats_reduce: unsupported channel/group configuration"))
    return EMOTE_ERR_INVALID;
  // emote_set_tuning("gn_reduce", 0) / EMOTE_GN_REDUCE=slots selects the per-slot walk for A/B timing
  const bool per_slot_walk = tuning(TUNE_GN_REDUCE, 1) == 0;
  const dim3 grid((unsigned)(n_batches * groups));
  const int cpg = C_total / groups;
  const long long per_group = (long long)slots_per_batch * cpg;   // (slot, column) pairs one block folds
  if (per_slot_walk || per_group >= (1LL << 30))
    launch_kernel(gn_colstats_reduce_kernel, grid, dim3(256), 0, stream, colstats, C_src, c_offset, cpg, groups,
                  slots_per_batch, n_batches, sums, zero_first);
  else if (per_group > 8 * 512)
    launch_kernel(gn_colstats_reduce_flat_kernel<1024>, grid, dim3(1024), 0, stream, colstats, C_src, c_offset, cpg,
                  groups, slots_per_batch, sums, zero_first);
  else
    launch_kernel(gn_colstats_reduce_flat_kernel<256>, grid, dim3(256), 0, stream, colstats, C_src, c_offset, cpg, groups,
                  slots_per_batch, sums, zero_first);
  EMOTE_CHECK_LAUNCH("emote_gn_colstats_reduce");
  return 0;
}

extern "C" int emote_gn_apply(const float* x, int32_t C_src, int32_t c_offset, int32_t C_total, int32_t groups,
                              int64_t rows_per_batch, int32_t n_batches, const double* sums, const float* gamma,
                              const float* beta, float eps, int32_t act_silu, void* out_bf16, void* raw_out_bf16,
                              void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (!x || !sums || !gamma || !beta || !out_bf16 || rows_per_batch <= 0 || n_batches <= 0)
    return set_error("emote_gn_apply: bad arguments");
  if (gn_check(C_src, c_offset, C_total, groups, "emote_gn_apply: unsupported channel/group configuration")) return EMOTE_ERR_INVALID;
  const int opr = C_src / 8;
  int bx = opr < 256 ? opr : 256;
  while (opr % bx != 0) --bx;  // octets per pass divide the row evenly
  int by = 256 / bx;
  if (by < 1) by = 1;
  const long long total_rows = rows_per_batch * (long long)n_batches;
  // rows per block: aim at ~16 blocks per SM (2368) on large tensors, one load batch per block on small ones
  int target_blocks = tuning(TUNE_GN_APPLY_BLOCKS, 2368);
  if (target_blocks < 1) target_blocks = 2368;
  long long k = total_rows / ((long long)GN_APPLY_U * by * target_blocks);
  if (k < 1) k = 1;
  const int kmax = tuning(TUNE_GN_APPLY_BLOCKS, 0) > 0 ? 8 : 4;   // the knob also lifts the cap for the sweep
  if (k > kmax) k = kmax;
  const int rows_per_block = (int)(GN_APPLY_U * by * k);
  const long long chunks = (rows_per_batch + rows_per_block - 1) / rows_per_block;
  dim3 grid((unsigned)chunks, (unsigned)n_batches), block(bx, by);
  const size_t smem = 2 * (size_t)C_src * sizeof(float);
  launch_kernel(gn_apply_kernel, dim3(grid), dim3(block), smem, stream, x, C_src, c_offset, C_total, C_total / groups, groups, rows_per_batch,
                                                 rows_per_block, sums, gamma, beta, eps, act_silu,
                                                 reinterpret_cast<op16*>(out_bf16),
                                                 reinterpret_cast<op16*>(raw_out_bf16));
  EMOTE_CHECK_LAUNCH("emote_gn_apply");
  return 0;
}

static int layernorm_impl(const float* x, int64_t M, int32_t C, const float* gamma, const float* beta, float eps,
                          const float* pe, int32_t pe_rows_per_frame, int32_t pe_frames, void* out_bf16, float* out_f32,
                          void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (!x || !gamma || !beta || (!out_bf16 && !out_f32) || M <= 0) return set_error("emote_layernorm: bad arguments");
  if (C % 64 != 0 || C > 64 * LN_MAX_V) return set_error("emote_layernorm: C must be a multiple of 64 and <= 2048");
  if (pe && (pe_rows_per_frame <= 0 || pe_frames <= 0)) return set_error("emote_layernorm: bad positional table dims");
  int warps = tuning(TUNE_LN_WARPS, 4);
  if (warps < 1 || warps > 8) warps = 4;
  const long long blocks = (M + warps - 1) / warps;
  op16* o = reinterpret_cast<op16*>(out_bf16);
#define EMOTE_LN(NVV) launch_kernel(layernorm_kernel<NVV>, dim3((unsigned)blocks), dim3(warps * 32), 0, stream, \
      x, M, C, gamma, beta, eps, pe, pe_rows_per_frame, pe_frames, o, out_f32)
  switch (C / 64) {
    case 1: EMOTE_LN(1); break;
    case 2: EMOTE_LN(2); break;
    case 4: EMOTE_LN(4); break;
    case 5: EMOTE_LN(5); break;
    case 8: EMOTE_LN(8); break;
    case 10: EMOTE_LN(10); break;
    case 20: EMOTE_LN(20); break;
    default: EMOTE_LN(0); break;
  }
#undef EMOTE_LN
  EMOTE_CHECK_LAUNCH("emote_layernorm");
  return 0;
}

extern "C" int emote_layernorm(const float* x, int64_t M, int32_t C, const float* gamma, const float* beta, float eps,
                               const float* pe, int32_t pe_rows_per_frame, int32_t pe_frames, void* out_bf16,
                               void* stream_) {
  if (!out_bf16) return set_error("emote_layernorm: bad arguments");
  return layernorm_impl(x, M, C, gamma, beta, eps, pe, pe_rows_per_frame, pe_frames, out_bf16, nullptr, stream_);
}

extern "C" int emote_layernorm_dual(const float* x, int64_t M, int32_t C, const float* gamma, const float* beta, float eps,
                                    void* out_bf16, float* out_f32, void* stream_) {
  return layernorm_impl(x, M, C, gamma, beta, eps, nullptr, 0, 0, out_bf16, out_f32, stream_);
}

extern "C" int emote_softmax_rows_bf16(const float* scores, int64_t R, int32_t N, float scale, void* out_bf16,
                                       void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (!scores || !out_bf16 || R <= 0 || N <= 0) return set_error("emote_softmax_rows_bf16: bad arguments");
  launch_kernel(softmax_rows_kernel, dim3((unsigned)R), dim3(256), 0, stream, scores, N, scale, reinterpret_cast<op16*>(out_bf16));
  EMOTE_CHECK_LAUNCH("emote_softmax_rows_bf16");
  return 0;
}
