// Audio front-end kernels: the pieces of wav2vec2-base (`Wav2VecFeatureExtractor`, reference Net.py:607-667
// -> transformers Wav2Vec2Model) that are not plain GEMM / LayerNorm / attention launches, and the SpeedEncoder
// (Net.py:198-258).  All HBM-light: a 10 s clip is 160 000 samples -> 499 frames of 768 features.
//   * waveform statistics + normalisation + layer-0 im2col (Wav2Vec2FeatureExtractor.zero_mean_unit_var_norm, Conv1d k=10 s=5)
//   * per-channel GroupNorm (num_groups == num_channels) over time + GELU  (Wav2Vec2GroupNormConvLayer)
//   * [T, C] tokens -> zero-padded group-major [G, T + pad, C/G] operand of the grouped positional convolution
//   * SpeedEncoder: tanh bucket encoding + Linear -> ReLU -> Linear
#include "common.cuh"
#include "emote_b200.h"
#include "host_utils.h"

namespace emote {

// stats[0] = mean, stats[1] = 1 / sqrt(var + eps)  (population variance), one block
__global__ void __launch_bounds__(1024) wave_stats_kernel(const float* __restrict__ x, long long n, float eps,
                                                          float* __restrict__ stats) {
  pdl_prologue();
  __shared__ double rs[32], rq[32];
  double s = 0.0, q = 0.0;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) {
    const double v = (double)x[i];
    s += v;
    q += v * v;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    q += __shfl_xor_sync(0xffffffffu, q, o);
  }
  if ((threadIdx.x & 31) == 0) {
    rs[threadIdx.x >> 5] = s;
    rq[threadIdx.x >> 5] = q;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double ts = 0.0, tq = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
      ts += rs[w];
      tq += rq[w];
    }
    const double mean = ts / (double)n;
    double var = tq / (double)n - mean * mean;
    if (var < 0.0) var = 0.0;
    stats[0] = (float)mean;
    stats[1] = (float)(1.0 / sqrt(var + (double)eps));
  }
}

// out[t, j] = op16((x[stride*t + j] - mean) * rstd) for j < k, 0 for k <= j < kpad
__global__ void wave_im2col_kernel(const float* __restrict__ x, const float* __restrict__ stats, long long T_out, int k,
                                   int stride, int kpad, op16* __restrict__ out) {
  pdl_prologue();
  const float mean = stats ? stats[0] : 0.f, rstd = stats ? stats[1] : 1.f;
  const long long total = T_out * kpad;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(i % kpad);
    const long long t = i / kpad;
    out[i] = float2op16(j < k ? (x[t * stride + j] - mean) * rstd : 0.f);
  }
}

// per-channel (sum, sum of squares) over the rows of x [T, C]: blockDim.x threads = C/2 channel pairs
__global__ void chan_stats_kernel(const float* __restrict__ x, long long T, int C, int rows_per_block,
                                  double* __restrict__ sums) {
  pdl_prologue();
  const long long r0 = (long long)blockIdx.x * rows_per_block;
  long long r1 = r0 + rows_per_block;
  if (r1 > T) r1 = T;
  for (int cp = threadIdx.x; cp < C / 2; cp += blockDim.x) {
    double s0 = 0.0, q0 = 0.0, s1 = 0.0, q1 = 0.0;
    for (long long r = r0; r < r1; ++r) {
      const float2 v = __ldg(reinterpret_cast<const float2*>(x + r * C) + cp);
      s0 += v.x; q0 += (double)v.x * v.x;
      s1 += v.y; q1 += (double)v.y * v.y;
    }
    atomicAdd(&sums[(2 * cp) * 2], s0);
    atomicAdd(&sums[(2 * cp) * 2 + 1], q0);
    atomicAdd(&sums[(2 * cp + 1) * 2], s1);
    atomicAdd(&sums[(2 * cp + 1) * 2 + 1], q1);
  }
}

// y = gelu_erf((x - mean_c) * rstd_c * gamma_c + beta_c) -> op16
__global__ void chan_norm_gelu_kernel(const float* __restrict__ x, long long T, int C, const double* __restrict__ sums,
                                      const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                      op16* __restrict__ out) {
  pdl_prologue();
  const long long total = T * (C / 2);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int cp = (int)(i % (C / 2));
    const float2 v = __ldg(reinterpret_cast<const float2*>(x) + i);
    float y[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int c = 2 * cp + k;
      const double mean = sums[2 * c] / (double)T;
      double var = sums[2 * c + 1] / (double)T - mean * mean;
      if (var < 0.0) var = 0.0;
      const float rstd = (float)(1.0 / sqrt(var + (double)eps));
      const float xv = k == 0 ? v.x : v.y;
      y[k] = gelu_erf_fast((xv - (float)mean) * rstd * gamma[c] + beta[c]);
    }
    reinterpret_cast<uint32_t*>(out)[i] = pack_op16x2(y[0], y[1]);
  }
}

// x fp32 [T, C] -> op16 [G, T + pad_front + pad_back, C/G], zero padded in time
__global__ void tokens_to_groups_kernel(const float* __restrict__ x, long long T, int C, int G, int pad_front,
                                        int pad_back, op16* __restrict__ out) {
  pdl_prologue();
  const int cg = C / G;
  const long long Tp = T + pad_front + pad_back;
  const long long total = (long long)G * Tp * cg;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % cg);
    const long long tp = (i / cg) % Tp;
    const int g = (int)(i / (cg * Tp));
    const long long t = tp - pad_front;
    out[i] = float2op16((t >= 0 && t < T) ? x[t * C + g * cg + c] : 0.f);
  }
}

// one block per sample: v_i = tanh((s - center_i) / radius_i * 3); h = relu(W1 v + b1); out = W2 h + b2     (fp32)
__global__ void speed_encoder_kernel(const float* __restrict__ speeds, const float* __restrict__ centers,
                                     const float* __restrict__ radii, int nb, const float* __restrict__ w1,
                                     const float* __restrict__ b1, const float* __restrict__ w2,
                                     const float* __restrict__ b2, int E, float* __restrict__ out) {
  pdl_prologue();
  extern __shared__ float se_smem[];
  float* v = se_smem;        // [nb]
  float* h = se_smem + nb;   // [E]
  const float s = speeds[blockIdx.x];
  for (int i = threadIdx.x; i < nb; i += blockDim.x) v[i] = tanhf((s - centers[i]) / radii[i] * 3.0f);
  __syncthreads();
  for (int e = threadIdx.x; e < E; e += blockDim.x) {
    float acc = b1[e];
    for (int i = 0; i < nb; ++i) acc = fmaf(w1[e * nb + i], v[i], acc);
    h[e] = fmaxf(acc, 0.f);
  }
  __syncthreads();
  for (int e = threadIdx.x; e < E; e += blockDim.x) {
    float acc = b2[e];
    for (int i = 0; i < E; ++i) acc = fmaf(w2[e * E + i], h[i], acc);
    out[(long long)blockIdx.x * E + e] = acc;
  }
}

static inline unsigned audio_grid(long long n, int threads) {
  long long b = (n + threads - 1) / threads;
  if (b > 148LL * 16) b = 148LL * 16;
  if (b < 1) b = 1;
  return (unsigned)b;
}

}  // namespace emote

using namespace emote;
#define STREAM(s) reinterpret_cast<cudaStream_t>(s)

extern "C" int emote_wave_stats(const float* wave, int64_t n, float eps, float* stats, void* stream) {
  if (!wave || !stats || n <= 0) return set_error("emote_wave_stats: bad arguments");
  launch_kernel(wave_stats_kernel, dim3(1), dim3(1024), 0, STREAM(stream), wave, (long long)n, eps, stats);
  EMOTE_CHECK_LAUNCH("emote_wave_stats");
  return 0;
}

extern "C" int emote_wave_im2col(const float* wave, int64_t n, const float* stats, int32_t kernel, int32_t stride,
                                 int32_t kpad, void* out_op16, void* stream) {
  if (!wave || !out_op16 || n < kernel || kernel <= 0 || stride <= 0 || kpad < kernel || kpad % 8 != 0)
    return set_error("emote_wave_im2col: bad arguments (kpad must be a multiple of 8 >= kernel, n >= kernel)");
  const long long T_out = (n - kernel) / stride + 1;
  launch_kernel(wave_im2col_kernel, dim3(audio_grid(T_out * kpad, 256)), dim3(256), 0, STREAM(stream), wave, stats, T_out,
                kernel, stride, kpad, reinterpret_cast<op16*>(out_op16));
  EMOTE_CHECK_LAUNCH("emote_wave_im2col");
  return 0;
}

extern "C" int emote_channel_norm_gelu(const float* x, int64_t T, int32_t C, const float* gamma, const float* beta,
                                       float eps, double* sums_scratch, void* out_op16, void* stream) {
  if (!x || !gamma || !beta || !sums_scratch || !out_op16 || T <= 0 || C <= 0 || C % 2 != 0)
    return set_error("emote_channel_norm_gelu: bad arguments (C must be even)");
  cudaError_t e = cudaMemsetAsync(sums_scratch, 0, sizeof(double) * 2 * (size_t)C, STREAM(stream));
  if (e != cudaSuccess) return set_error_cuda("emote_channel_norm_gelu memset", e);
  const int rows_per_block = 64;
  const int threads = C / 2 < 256 ? C / 2 : 256;
  launch_kernel(chan_stats_kernel, dim3((unsigned)((T + rows_per_block - 1) / rows_per_block)), dim3(threads), 0,
                STREAM(stream), x, (long long)T, C, rows_per_block, sums_scratch);
  EMOTE_CHECK_LAUNCH("emote_channel_norm_gelu(stats)");
  launch_kernel(chan_norm_gelu_kernel, dim3(audio_grid((long long)T * (C / 2), 256)), dim3(256), 0, STREAM(stream), x,
                (long long)T, C, (const double*)sums_scratch, gamma, beta, eps, reinterpret_cast<op16*>(out_op16));
  EMOTE_CHECK_LAUNCH("emote_channel_norm_gelu");
  return 0;
}

extern "C" int emote_tokens_to_groups(const float* x, int64_t T, int32_t C, int32_t groups, int32_t pad_front,
                                      int32_t pad_back, void* out_op16, void* stream) {
  if (!x || !out_op16 || T <= 0 || C <= 0 || groups <= 0 || C % groups != 0 || pad_front < 0 || pad_back < 0)
    return set_error("emote_tokens_to_groups: bad arguments");
  const long long total = (long long)groups * (T + pad_front + pad_back) * (C / groups);
  launch_kernel(tokens_to_groups_kernel, dim3(audio_grid(total, 256)), dim3(256), 0, STREAM(stream), x, (long long)T, C,
                groups, pad_front, pad_back, reinterpret_cast<op16*>(out_op16));
  EMOTE_CHECK_LAUNCH("emote_tokens_to_groups");
  return 0;
}

extern "C" int emote_speed_encoder(const float* speeds, int32_t batch, const float* centers, const float* radii,
                                   int32_t n_buckets, const float* w1, const float* b1, const float* w2, const float* b2,
                                   int32_t embed_dim, float* out, void* stream) {
  if (!speeds || !centers || !radii || !w1 || !b1 || !w2 || !b2 || !out || batch <= 0 || n_buckets <= 0 || embed_dim <= 0)
    return set_error("emote_speed_encoder: bad arguments");
  if ((size_t)(n_buckets + embed_dim) * sizeof(float) > 48 * 1024) return set_error("emote_speed_encoder: dims too large");
  launch_kernel(speed_encoder_kernel, dim3((unsigned)batch), dim3(128), (size_t)(n_buckets + embed_dim) * sizeof(float),
                STREAM(stream), speeds, centers, radii, n_buckets, w1, b1, w2, b2, embed_dim, out);
  EMOTE_CHECK_LAUNCH("emote_speed_encoder");
  return 0;
}
