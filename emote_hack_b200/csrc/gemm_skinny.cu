// Skinny GEMM (M <= 8 rows): out[m][n] = bias[n] + sum_k a[m][k] * w[n][k], bf16 operands, fp32 accumulate.
//
// The time-embedding path of every UNet call is 24 such products on [2, 1280] activations (TimestepEmbedding
// linear_1 / linear_2, embeddings.py:206-218, and the 22 per-resnet time_emb_proj, resnet.py:186).  Through the
// 128-row tcgen05 tile kernel each of them is one CTA column walking 20 k-blocks in series (17 us, all latency); here
// they are weight-streaming GEMVs: one warp per output column, the activations in shared memory, 16-byte coalesced
// weight reads (3.3 MB for 1280 x 1280, HBM/L2-bound at a few microseconds).
#include "common.cuh"
#include "emote_b200.h"
#include "host_utils.h"

namespace emote {

constexpr int SK_MAX_M = 8;
constexpr int SK_WARPS = 8;

__global__ void __launch_bounds__(SK_WARPS * 32)
gemm_skinny_kernel(const op16* __restrict__ a, int lda, const op16* __restrict__ w,
                   const float* __restrict__ bias, int M, int N, int K, float out_scale, void* __restrict__ out, int ldc,
                   int out_bf16) {
  pdl_prologue();
  extern __shared__ uint4 sk_smem[];   // [M][K/8] 16-byte chunks of a
  const int kc = K >> 3;
  for (int i = threadIdx.x; i < M * kc; i += blockDim.x) {
    const int m = i / kc, c = i - m * kc;
    sk_smem[i] = *reinterpret_cast<const uint4*>(a + (size_t)m * lda + c * 8);
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int n = blockIdx.x * SK_WARPS + warp; n < N; n += gridDim.x * SK_WARPS) {
    float acc[SK_MAX_M];
#pragma unroll
    for (int m = 0; m < SK_MAX_M; ++m) acc[m] = 0.f;
    const uint4* wr = reinterpret_cast<const uint4*>(w + (size_t)n * K);
    for (int c = lane; c < kc; c += 32) {
      const uint4 wv = __ldg(wr + c);
      const op16x2* w2 = reinterpret_cast<const op16x2*>(&wv);
      float wf[8];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = op16x2_to_float2(w2[j]);
        wf[2 * j] = f.x;
        wf[2 * j + 1] = f.y;
      }
#pragma unroll
      for (int m = 0; m < SK_MAX_M; ++m) {
        if (m < M) {
          const uint4 av = sk_smem[m * kc + c];
          const op16x2* a2 = reinterpret_cast<const op16x2*>(&av);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 f = op16x2_to_float2(a2[j]);
            acc[m] = fmaf(f.x, wf[2 * j], acc[m]);
            acc[m] = fmaf(f.y, wf[2 * j + 1], acc[m]);
          }
        }
      }
    }
    const float b = bias ? __ldg(bias + n) : 0.f;
#pragma unroll
    for (int m = 0; m < SK_MAX_M; ++m) {
      if (m < M) {
        const float v = (warp_sum(acc[m]) + b) * out_scale;
        if (lane == 0) {
          if (out_bf16) reinterpret_cast<op16*>(out)[(size_t)m * ldc + n] = float2op16(v);
          else reinterpret_cast<float*>(out)[(size_t)m * ldc + n] = v;
        }
      }
    }
  }
}

bool gemm_skinny_applicable(const EmoteGemmArgs* a) {
  return a->conv_taps == 1 && a->M <= SK_MAX_M && a->K % 8 == 0 && a->lda % 8 == 0 && a->epilogue == EMOTE_EPI_LINEAR &&
         !a->residual && !a->row_bias && !a->colstats && (size_t)a->M * a->K * 2 <= 96 * 1024 && a->pair_mode == 0 &&
         a->block_n == 0;
}

int launch_gemm_skinny(const void* A, const void* Wt, void* out, const EmoteGemmArgs* a, cudaStream_t stream) {
  if ((reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(Wt) & 15))
    return set_error("emote_gemm_bf16(skinny): operands must be 16-byte aligned");
  const size_t smem = (size_t)a->M * a->K * 2;
  static PerDeviceOnce configured;
  int dev__ = 0;
  if (smem > 48 * 1024 && configured.pending(&dev__)) {
    cudaError_t e = cudaFuncSetAttribute(gemm_skinny_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    if (e != cudaSuccess) return set_error_cuda("cudaFuncSetAttribute(gemm_skinny)", e);
    configured.done(dev__);
  }
  int blocks = (a->N + SK_WARPS - 1) / SK_WARPS;
  if (blocks > 148 * 4) blocks = 148 * 4;
  launch_kernel(gemm_skinny_kernel, dim3(blocks), dim3(SK_WARPS * 32), smem, stream,
                reinterpret_cast<const op16*>(A), a->lda, reinterpret_cast<const op16*>(Wt), a->bias, a->M,
                a->N, a->K, a->out_scale, out, a->ldc, a->out_dtype == EMOTE_DT_OP16 ? 1 : 0);
  EMOTE_CHECK_LAUNCH("emote_gemm_bf16(skinny)");
  return 0;
}

}  // namespace emote
