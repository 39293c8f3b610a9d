/* emote_b200.h — C ABI of libemote_b200.so (B200 / sm_100a kernels for the Emote-hack denoising hot path).
 *
 * The reference (johndpope/Emote-hack) has no FFI: its hot path bottoms out in ATen/cuDNN/cuBLAS library
 * calls made from Python nn.Modules.  This header is the boundary a maintainer would bind
 * (ctypes stub in INTEGRATION.md); each entry point names the reference call it replaces (file:line under
 * the reference checkout).
 *
 * Conventions: raw device pointers + explicit sizes, caller-owned buffers, no allocation and no host sync
 * inside, work is enqueued on `stream` (a cudaStream_t passed as void*).  Every function returns 0 on success,
 * non-zero on error; emote_last_error() gives the message (thread-local).  "tokens-major" means the
 * activation tensor [b, c, f, h, w] stored channels-last, i.e. row index = ((b*F + f)*H + y)*W + x, C contiguous.
 */
#ifndef EMOTE_B200_H
#define EMOTE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EMOTE_ABI_VERSION 2
#define EMOTE_ERR_INVALID 1
#define EMOTE_ERR_CUDA 2

/* The library is built for ONE 16-bit tensor-core operand type ("op16"): fp16 in the default build
 * (libemote_b200.so; the reference pipeline runs fp16, magicanimate/pipelines/animation.py:96-100) or bf16
 * (libemote_b200_bf16.so, built with -DEMOTE_OPERAND_BF16).  emote_operand_dtype() tells which.  Wherever this header
 * says "bf16" / "_bf16" for an operand buffer it means that op16 type (the names predate the fp16 build); statistics,
 * residual stream, accumulators and softmax are fp32 in both builds. */
#define EMOTE_OP_BF16 1
#define EMOTE_OP_F16 2
int emote_operand_dtype(void); /* EMOTE_OP_BF16 or EMOTE_OP_F16 */

/* EmoteGemmArgs.out_dtype */
#define EMOTE_DT_F32 0
#define EMOTE_DT_OP16 1

#define EMOTE_EPI_LINEAR 0
#define EMOTE_EPI_GEGLU 1
#define EMOTE_EPI_GELU 2   /* out = out_scale * (gelu_erf(acc + bias) + residual): wav2vec2 conv / feed-forward layers */

const char* emote_last_error(void);
long long emote_launch_count(void); /* kernels launched by this library so far (bench.py "gpu_launches") */
int emote_abi_version(void);
/* Programmatic dependent launch of the library's kernels (default off; EMOTE_PDL=1 in the environment enables it):
 * each kernel's set-up may overlap the tail of the previous one; results are unaffected. */
void emote_set_pdl(int enabled);
/* Measurement knobs (A/B timing of kernel variants inside one process; results are the same up to fp64 summation order).
 * key "gn_reduce": 1 = flat fold of the GroupNorm statistic slots (default), 0 = the per-slot walk.  Returns 0, or
 * EMOTE_ERR_INVALID for an unknown key. */
int emote_set_tuning(const char* key, int32_t value);

/* ------------------------------------------------------------------------------------------------ GEMM / conv
 * out[M,N] = epilogue(A[M,K] x Wt[N,K]^T), bf16 operands, fp32 accumulation on tcgen05 tensor cores.
 * conv_taps = 9: A is an NHWC bf16 tensor [n_img,H,W,C]; computes the 3x3 / stride 1 / pad 1 convolution as an
 * implicit GEMM with Wt packed as [N, (ky*3+kx)*C + c].
 * Replaces: nn.Linear / nn.Conv2d(1x1) / InflatedConv3d(3x3) — resnet.py:30-38,180-205; attention.py:126-150;
 * orig_attention.py:606-650,776-781,825-827; motion_module.py:149-159.
 * Epilogue: v = acc + bias[n] + row_bias[m / rows_per_group, n] + residual[m, n]; out = out_scale * v.
 * EMOTE_EPI_GEGLU: Wt/bias rows are packed per block_n tile as [value rows | gate rows]; out[M, N/2] bf16 =
 * (acc_v + b_v) * gelu_erf(acc_g + b_g)   (orig_attention.py:817-827). */
typedef struct {
  int32_t M, N, K;
  int32_t lda;            /* A row pitch in elements (plain GEMM); may be < K: overlapping rows = the zero-copy operand of a
                             strided 1-D convolution over a [T, C] token matrix (lda = stride*C, K = kernel*C) */
  int32_t conv_taps;      /* 1 or 9 */
  int32_t n_img, H, W, C; /* conv mode */
  const float* bias;      /* [N] or NULL */
  const float* row_bias;  /* [ceil(M/rows_per_group), N] or NULL */
  int32_t rows_per_group;
  const float* residual;  /* [M, ldr] fp32 or NULL; may alias out */
  int32_t ldr;
  float out_scale;
  int32_t epilogue;       /* EMOTE_EPI_* */
  int32_t out_dtype;      /* EMOTE_DT_F32 or EMOTE_DT_OP16 */
  int32_t ldc;            /* out row pitch in elements */
  int32_t block_n;        /* 0 = auto (160 if N % 160 == 0 else 128) */
  int32_t pair_mode;      /* 0 = auto, 1 = force CTA pairs (cta_group::2, 256-row tiles), 2 = force single CTA,
                             3 = auto without the weight-stationary kernel (small-K bf16-output shapes) */
  int32_t tma_store;      /* 0 = auto (staged TMA-store epilogues where they apply), 2 = never (register path; dev / A-B runs) */
  float* colstats;        /* optional fused GroupNorm statistics of the fp32 output: one (sum, sum of squares) fp32 slot per
                             32-row quarter of a 128-row sub-tile and column, [ceil(M/128)*4][N][2] (conv over 2-D patch
                             tiles: sub-tile order), every slot written with a plain store: no zero-fill, no atomics,
                             bit-reproducible.  Slots of statistics batch b: [b*stats_rows/32, (b+1)*stats_rows/32). */
  int32_t stats_rows;     /* rows per statistics batch; multiple of 32 that divides M */
} EmoteGemmArgs;
int emote_gemm_bf16(const void* A, const void* Wt, void* out, const EmoteGemmArgs* args, void* stream);

/* ------------------------------------------------------------------------------------------------ GroupNorm
 * Two-kernel GroupNorm over tokens-major fp32 activations (nn.GroupNorm at resnet.py:180,191 [5-D: statistics
 * span all frames of a sample], attention.py:124 and motion_module.py:147 [per frame], unet_controlnet.py:476).
 * A "group batch" is the set of rows sharing statistics (F*H*W rows for the 5-D norm, H*W for per-frame).
 * The input may be one of several channel-concatenated sources (torch.cat at unet_3d_blocks.py:629,731):
 * source channels [0,C_src) map to channels [c_offset, c_offset+C_src) of the C_total-wide normalised tensor.
 * sums: [n_batches, groups, 2] doubles (sum, sum of squares); zeroed by the call when zero_first != 0. */
int emote_gn_stats(const float* x, int32_t C_src, int32_t c_offset, int32_t C_total, int32_t groups,
                   int64_t rows_per_batch, int32_t n_batches, double* sums, int32_t zero_first, void* stream);
/* GroupNorm statistics from the slots emote_gemm_bf16 stored (EmoteGemmArgs.colstats) while it wrote the tensor being
 * normalised: replaces the emote_gn_stats pass over that source (same `sums` layout and concat semantics: c_offset /
 * C_total; zero_first = overwrite instead of accumulate).  `slots_per_batch` = rows of one GroupNorm batch / 32 (e.g. the
 * f frames of a sample for the 5-D GroupNorm of resnet.py:180).  Fixed summation order, fp64: deterministic. */
int emote_gn_colstats_reduce(const float* colstats, int32_t C_src, int32_t c_offset, int32_t C_total, int32_t groups,
                             int32_t slots_per_batch, int32_t n_batches, double* sums, int32_t zero_first,
                             void* stream);

/* y = (x-mean)*rstd*gamma+beta [-> SiLU if act_silu]; written as bf16 into out[row, c_offset + c] (pitch C_total).
 * raw_out (optional, same layout) receives the un-normalised input rounded to bf16 (shortcut-conv operand). */
int emote_gn_apply(const float* x, int32_t C_src, int32_t c_offset, int32_t C_total, int32_t groups,
                   int64_t rows_per_batch, int32_t n_batches, const double* sums, const float* gamma,
                   const float* beta, float eps, int32_t act_silu, void* out_bf16, void* raw_out_bf16, void* stream);

/* ------------------------------------------------------------------------------------------------ LayerNorm
 * out_bf16[M,C] = LN(x[M,C]) * gamma + beta (+ pe[(row / pe_rows_per_frame) % pe_frames, :]).
 * nn.LayerNorm at attention.py:204,225,232 and motion_module.py:208,213; the additive table is the temporal
 * sinusoidal PositionalEncoding applied after the norm (motion_module.py:246-248,283). */
int emote_layernorm(const float* x, int64_t M, int32_t C, const float* gamma, const float* beta, float eps,
                    const float* pe, int32_t pe_rows_per_frame, int32_t pe_frames, void* out_bf16, void* stream);
/* same, additionally writing the normalised rows in fp32 (post-LN transformers — wav2vec2's encoder layers,
 * transformers Wav2Vec2EncoderLayer — continue their fp32 residual stream from the LayerNorm output); out_bf16 may be
 * NULL when only the fp32 copy is wanted */
int emote_layernorm_dual(const float* x, int64_t M, int32_t C, const float* gamma, const float* beta, float eps,
                         void* out_bf16, float* out_f32, void* stream);

/* ------------------------------------------------------------------------------------------------ attention
 * Flash-style softmax(Q K^T * scale) V on bf16 with fp32 softmax/accumulation (orig_attention.py:655-684).
 * Keys/values come from up to two segments that are logically concatenated along the key axis
 * (mutual_self_attention.py:239-241: [self tokens | reference bank]); segment 1 is only visible to batches
 * b >= kv1_first_batch (the unconditional CFG half must not see the bank, mutual_self_attention.py:244-255).
 * Row r, head h of batch b lives at base + b_idx*batch_stride + r*row_stride + h*head_dim (elements). */
typedef struct {
  const void* q;
  const void* k0;
  const void* v0;
  const void* k1; /* NULL when n1 == 0 */
  const void* v1;
  void* out;
  int32_t batch, heads, head_dim;
  int32_t nq, n0, n1;
  int64_t q_batch_stride, q_row_stride;
  int64_t kv0_batch_stride, kv0_row_stride;
  int64_t kv1_batch_stride, kv1_row_stride;
  int64_t o_batch_stride, o_row_stride;
  int32_t kv0_batch_div; /* kv batch index = b / div (context shared by the frames of a sample) */
  int32_t kv1_batch_div;
  int32_t kv1_first_batch;
  float scale;
} EmoteAttnArgs;
int emote_attention_bf16(const EmoteAttnArgs* args, void* stream);
/* Same contract on the tcgen05 tensor cores (S and P*V accumulate in TMEM, V consumed MN-major): head_dim 40 / 80,
 * i.e. the 64x64 and 32x32 spatial self-attention / reference-attention layers that dominate the attention time. */
int emote_attention_tc_bf16(const EmoteAttnArgs* args, void* stream);
int emote_attention_tc_supported(int32_t head_dim); /* 1 if emote_attention_tc_bf16 handles this head_dim */

/* One wide head (head_dim 512): the mid-block attention of the SD VAE (legacy AttentionBlock, orig_attention.py:253-385,
 * `vae.decode` / `vae.encode` at EMOAnimationPipeline.py:301,412) as a tcgen05 flash kernel batched over frames — no
 * [N, N] score matrix, no per-frame loop.  q / k / v: op16 rows of 512 (e.g. views into a fused QKV GEMM output),
 * row r of image b at base + b*batch_stride + r*row_stride (elements); out likewise. */
int emote_attention_wide_bf16(const void* q, const void* k, const void* v, void* out, int32_t batch, int32_t nq, int32_t nk,
                              int32_t head_dim, int64_t q_batch_stride, int64_t q_row_stride, int64_t kv_batch_stride,
                              int64_t kv_row_stride, int64_t o_batch_stride, int64_t o_row_stride, float scale,
                              void* stream);

/* Temporal self-attention over the frame axis (VersatileAttention, motion_module.py:275-334): for every
 * (sample b, pixel p, head h) attends over the F frames.  qkv: tokens-major [B, F, HW, 3*heads*head_dim] bf16
 * (q | k | v), out: [B, F, HW, heads*head_dim] bf16.  The (b f) d c <-> (b d) f c rearranges
 * (motion_module.py:280,332) are expressed as strides, never materialised.  F <= 32. */
int emote_temporal_attention_bf16(const void* qkv, void* out, int32_t B, int32_t F, int32_t HW, int32_t heads,
                                  int32_t head_dim, float scale, void* stream);

/* row softmax of fp32 scores [R, N] (scaled) -> bf16 probabilities; used by the VAE mid-block attention
 * (single 512-wide head: orig_attention.py:360-376). */
int emote_softmax_rows_bf16(const float* scores, int64_t R, int32_t N, float scale, void* out_bf16, void* stream);

/* ------------------------------------------------------------------------------------------------ layout / gather
 * conv_in operand: latents [B, Cl<=7, F, H, W] fp32 -> im2col rows [B*F*H*W, 64] bf16 (cols (ky*3+kx)*Cl + c, rest 0)
 * after x*pre_scale and an optional pointwise Cl x Cl linear (VAE post_quant_conv).  unet_controlnet.py:411,
 * EMOAnimationPipeline.py:293-301. */
int emote_latent_im2col(const float* latent, int32_t B, int32_t Cl, int32_t F, int32_t H, int32_t W, float pre_scale,
                        const float* pw_weight, const float* pw_bias, void* out_bf16, void* stream);
/* generic 3x3 pad-1 im2col from tokens-major fp32 [n_img,H,W,C] -> bf16 [n_img*Ho*Wo, 9*C], stride 1 or 2
 * (Downsample3D, resnet.py:99-110; also the fallback for image sizes the TMA conv tile cannot cover). */
int emote_im2col3x3(const float* x, int32_t n_img, int32_t H, int32_t W, int32_t C, int32_t stride, void* out_bf16,
                    void* stream);
/* stride-2 3x3 gather with the VAE encoder's asymmetric padding: rows/cols padded by (0, 1) instead of (1, 1)
 * (diffusers Downsample2D(padding=0): F.pad(x, (0,1,0,1)) then conv stride 2 — the `vae.encode` of
 * EMOAnimationPipeline.py:412); out [n_img*(H/2)*(W/2), 9*C] bf16, H and W even. */
int emote_im2col3x3_s2_pad01(const float* x, int32_t n_img, int32_t H, int32_t W, int32_t C, void* out_bf16,
                             void* stream);
/* same gather from a bf16 NHWC source (fallback path of the fused-norm convs) */
int emote_im2col3x3_bf16(const void* x, int32_t n_img, int32_t H, int32_t W, int32_t C, int32_t stride,
                         void* out_bf16, void* stream);
/* nearest-neighbour x2 upsample in H and W (Upsample3D, resnet.py:74), fp32 -> bf16 conv operand */
int emote_upsample2x(const float* x, int32_t n_img, int32_t H, int32_t W, int32_t C, void* out_bf16, void* stream);
/* nearest-neighbour resize to a forced Ho x Wo (F.interpolate(size=output_size), resnet.py:76: the `upsample_size` the UNet
 * forwards when the latent size is not a multiple of 2**num_upsamplers, unet_controlnet.py:355-364,453-460) */
int emote_upsample_nearest(const float* x, int32_t n_img, int32_t H, int32_t W, int32_t C, int32_t Ho, int32_t Wo,
                           void* out_bf16, void* stream);
/* out_bf16[row, c_offset + c] = bf16(x[row, c]) (pitch C_total) */
int emote_cast_bf16(const float* x, int64_t rows, int32_t C_src, int32_t c_offset, int32_t C_total, void* out_bf16,
                    void* stream);
/* out_bf16[i] = bf16(silu(x[i]))  (time embedding activation, resnet.py:186) */
int emote_silu_bf16(const float* x, int64_t n, void* out_bf16, void* stream);
/* tokens-major [B,F,H,W,C] fp32 <-> [B,C,F,H,W] fp32 */
int emote_tokens_to_ncfhw(const float* tok, int32_t B, int32_t C, int32_t F, int32_t HW, float* out, void* stream);
int emote_ncfhw_to_tokens(const float* x, int32_t B, int32_t C, int32_t F, int32_t HW, float* out, void* stream);
/* out = a + b (ControlNet residual adds, unet_controlnet.py:430-447) */
int emote_add_f32(const float* a, const float* b, float* out, int64_t n, void* stream);
/* sinusoidal timestep embedding -> bf16 [B, dim] (embeddings.py:28-68) */
int emote_timestep_embedding(const float* timesteps, int32_t B, int32_t dim, int32_t flip_sin_to_cos,
                             float freq_shift, void* out_bf16, void* stream);

/* ------------------------------------------------------------------------------------------------ sampler
 * Fused window-average + classifier-free guidance + DDIM update (EMOAnimationPipeline.py:790-817):
 *   eps = eps_u/cnt + g * (eps_c/cnt - eps_u/cnt);  x0 = (x - sqrt(1-a_t) eps)/sqrt(a_t);
 *   x_prev = sqrt(a_prev) x0 + sqrt(1 - a_prev - sigma^2) eps + sigma z.
 * sigma = eta * sqrt((1-a_prev)/(1-a_t) (1 - a_t/a_prev)) is computed by the caller (0 for the reference's eta = 0, then
 * `noise` may be NULL); noise: [n] standard-normal z.
 * noise_pred: [2, n] (uncond, cond) accumulated predictions; counter: [n_frames] visit counts, frame index of
 * element i = (i / inner) % n_frames; latents updated in place.  zero_noise_pred != 0: the accumulator is cleared as it
 * is consumed (the per-timestep torch.zeros of EMOAnimationPipeline.py:702-706). */
int emote_cfg_ddim_step(float* latents, float* noise_pred, const float* counter, int64_t n, int32_t n_frames,
                        int64_t inner, float guidance_scale, float alpha_t, float alpha_prev, const float* noise,
                        float sigma, int32_t zero_noise_pred, void* stream);
/* The same update for a ready epsilon [n] (no guidance, no window average): `scheduler.step(noise_pred, t, latents)` as a
 * stand-alone call (EMOAnimationPipeline.py:817) and, with the two alphas exchanged, the inversion `next_step` (:379-400). */
int emote_ddim_step(float* latents, const float* eps, int64_t n, float alpha_t, float alpha_prev, const float* noise,
                    float sigma, void* stream);
/* Context-window bookkeeping of the denoise loop (EMOAnimationPipeline.py:759-763 `latents[:, :, c]` + repeat, :790-794
 * `noise_pred[:, :, c] += pred`) over fp32 tensors viewed as [outer, frames, inner] (inner contiguous, multiple of 4):
 *   gather:       dst[o, j, :] = src[(o % src_mod) + src_off, frame_idx[j], :]         o < n_outer, j < wlen
 *   scatter-add:  dst[o + dst_off, frame_idx[j], :] += src[o, j, :]                   (frames of a window are distinct)
 * frame_idx: device int32 [wlen].  Also serves the per-frame audio-token context gather (Net.py:646-667 tokens). */
int emote_gather_frames(const float* src, float* dst, const int32_t* frame_idx, int32_t n_outer, int32_t wlen,
                        int32_t F_src, int64_t inner, int32_t src_mod, int32_t src_off, void* stream);
int emote_scatter_add_frames(const float* src, float* dst, const int32_t* frame_idx, int32_t n_outer, int32_t wlen,
                             int32_t F_dst, int64_t inner, int32_t dst_off, void* stream);
/* p[0..n) = value (timestep scalar of a captured UNet step, accumulator clears) */
int emote_fill_f32(float* p, float value, int64_t n, void* stream);
/* decoded frames: tokens-major [n,H*W,ld>=3] fp32 -> clamp(x/2+0.5,0,1) as fp32 [n,3,H,W] and/or uint8 [n,3,H,W]
 * (EMOAnimationPipeline.py:304); the uint8 copy uses the truncating cast of the reference's writer (utils/util.py:28) */
int emote_vae_postprocess(const float* tok, int32_t n_img, int32_t HW, int32_t ld, float* out_f32, uint8_t* out_u8,
                          void* stream);
/* frames for the video writer — the loop of save_videos_grid (magicanimate/utils/util.py:21-30): videos [b,c,t,h,w]
 * fp32 (c = 1 or 3) -> uint8 [t, Hg, Wg, 3] with torchvision.utils.make_grid's layout per frame (nrow samples per row,
 * `padding` zero pixels around each; b == 1: no padding, Hg = h, Wg = w; else Hg = (h+padding)*ceil(b/min(nrow,b))+padding,
 * Wg = (w+padding)*min(nrow,b)+padding), optional (x+1)/2, then numpy's truncating float->uint8 cast. */
int emote_video_grid_u8(const float* videos, int32_t b, int32_t c, int32_t t, int32_t h, int32_t w, int32_t nrow,
                        int32_t padding, int32_t rescale, uint8_t* out, void* stream);

/* ------------------------------------------------------------------------------------------------ audio front-end
 * wav2vec2-base forward behind `Wav2VecFeatureExtractor` (Net.py:607-667 -> transformers Wav2Vec2Model).
 * The convolutions, linears, LayerNorms and the 12-head attention run on emote_gemm_bf16 (EMOTE_EPI_GELU, overlapping-row
 * operands for the strided 1-D convs), emote_layernorm(_dual) and emote_attention_bf16; these entries cover the rest. */
/* stats[0] = mean, stats[1] = 1/sqrt(var + eps) of the waveform (Wav2Vec2FeatureExtractor.zero_mean_unit_var_norm) */
int emote_wave_stats(const float* wave, int64_t n, float eps, float* stats, void* stream);
/* layer-0 operand: out[t, j] = op16((wave[stride*t + j] - mean) * rstd), j < kernel, zero up to kpad (multiple of 8);
 * t < (n - kernel)/stride + 1; stats may be NULL (no normalisation) */
int emote_wave_im2col(const float* wave, int64_t n, const float* stats, int32_t kernel, int32_t stride, int32_t kpad,
                      void* out_op16, void* stream);
/* GroupNorm with one channel per group over the time axis + exact GELU (Wav2Vec2GroupNormConvLayer): x fp32 [T, C] ->
 * op16 [T, C]; sums_scratch: 2*C doubles (cleared by the call) */
int emote_channel_norm_gelu(const float* x, int64_t T, int32_t C, const float* gamma, const float* beta, float eps,
                            double* sums_scratch, void* out_op16, void* stream);
/* x fp32 [T, C] -> op16 [groups, pad_front + T + pad_back, C/groups] zero padded in time: row t of group g then holds the
 * `kernel` consecutive frames of the grouped positional convolution contiguously (Wav2Vec2PositionalConvEmbedding) */
int emote_tokens_to_groups(const float* x, int64_t T, int32_t C, int32_t groups, int32_t pad_front, int32_t pad_back,
                           void* out_op16, void* stream);
/* SpeedEncoder (Net.py:198-258): v_i = tanh((s - center_i)/radius_i * 3); out = W2 relu(W1 v + b1) + b2, fp32.
 * speeds [batch]; w1 [embed_dim, n_buckets]; w2 [embed_dim, embed_dim]; out [batch, embed_dim] */
int emote_speed_encoder(const float* speeds, int32_t batch, const float* centers, const float* radii, int32_t n_buckets,
                        const float* w1, const float* b1, const float* w2, const float* b2, int32_t embed_dim, float* out,
                        void* stream);

#ifdef __cplusplus
}
#endif
#endif /* EMOTE_B200_H */
