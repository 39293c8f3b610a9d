// HBM-bound gather / layout / sampler kernels (coalesced 16-byte accesses along the channel axis).
#include "common.cuh"
#include "emote_b200.h"
#include "host_utils.h"

namespace emote {

static inline unsigned grid_for(long long n, int threads, long long cap = 148LL * 32) {
  long long b = (n + threads - 1) / threads;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (unsigned)b;
}

// latents [B, Cl, F, H, W] fp32 -> rows [B*F*H*W, 64] bf16; col (ky*3+kx)*Cl + c
__global__ void latent_im2col_kernel(const float* __restrict__ lat, int B, int Cl, int F, int H, int W, float pre_scale,
                                     const float* __restrict__ pw_w, const float* __restrict__ pw_b,
                                     op16* __restrict__ out) {
  pdl_prologue();
  const long long total = (long long)B * F * H * W;
  const long long plane = (long long)H * W;
  for (long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x; m < total;
       m += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(m % W);
    const int y = (int)((m / W) % H);
    const long long bf = m / plane;
    const int f = (int)(bf % F);
    const long long b = bf / F;
    __align__(16) op16 row[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) row[i] = float2op16(0.f);
    for (int ky = 0; ky < 3; ++ky) {
      for (int kx = 0; kx < 3; ++kx) {
        const int yy = y + ky - 1, xx = x + kx - 1;
        if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
        float v[8];
        for (int c = 0; c < Cl; ++c)
          v[c] = lat[(((b * Cl + c) * F + f) * H + yy) * W + xx] * pre_scale;
        if (pw_w) {
          float u[8];
          for (int o = 0; o < Cl; ++o) {
            float acc = pw_b ? pw_b[o] : 0.f;
            for (int c = 0; c < Cl; ++c) acc += pw_w[o * Cl + c] * v[c];
            u[o] = acc;
          }
          for (int c = 0; c < Cl; ++c) v[c] = u[c];
        }
        for (int c = 0; c < Cl; ++c) row[(ky * 3 + kx) * Cl + c] = float2op16(v[c]);
      }
    }
    uint4* o = reinterpret_cast<uint4*>(out + m * 64);
    const uint4* r = reinterpret_cast<const uint4*>(row);
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = r[i];
  }
}

template <typename T>
__device__ __forceinline__ uint4 load8_as_bf16(const T* p);
template <>
__device__ __forceinline__ uint4 load8_as_bf16<float>(const float* p) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p));
  const float4 b = __ldg(reinterpret_cast<const float4*>(p + 4));
  uint4 w;
  w.x = pack_op16x2(a.x, a.y); w.y = pack_op16x2(a.z, a.w);
  w.z = pack_op16x2(b.x, b.y); w.w = pack_op16x2(b.z, b.w);
  return w;
}
template <>
__device__ __forceinline__ uint4 load8_as_bf16<op16>(const op16* p) {
  return __ldg(reinterpret_cast<const uint4*>(p));
}

// x [n,H,W,C] -> out [n*Ho*Wo, 9*C] bf16 (3x3, pad 1, given stride)
template <typename T>
__global__ void im2col3x3_kernel(const T* __restrict__ x, int n_img, int H, int W, int C, int stride, int Ho, int Wo,
                                 op16* __restrict__ out, int pad_lo) {
  pdl_prologue();
  const int oct = C >> 3;
  const long long total = (long long)n_img * Ho * Wo * 9 * oct;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % oct);
    long long r = i / oct;
    const int tap = (int)(r % 9);
    r /= 9;
    const int xo = (int)(r % Wo);
    const int yo = (int)((r / Wo) % Ho);
    const long long img = r / ((long long)Wo * Ho);
    const int yy = yo * stride + tap / 3 - pad_lo;
    const int xx = xo * stride + tap % 3 - pad_lo;
    uint4 w = make_uint4(0u, 0u, 0u, 0u);
    if (yy >= 0 && yy < H && xx >= 0 && xx < W) w = load8_as_bf16<T>(x + ((img * H + yy) * W + xx) * C + c8 * 8);
    *reinterpret_cast<uint4*>(out + r * 9LL * C + (long long)tap * C + c8 * 8) = w;
  }
}

// nearest-neighbour resize to Ho x Wo: source index floor(dst * in / out) (F.interpolate(mode="nearest")); Ho = 2H, Wo = 2W
// is the x2 upsample of Upsample3D
__global__ void upsample_nearest_kernel(const float* __restrict__ x, int n_img, int H, int W, int C, int Ho, int Wo,
                                        op16* __restrict__ out) {
  pdl_prologue();
  const int oct = C >> 3;
  const float sy = (float)H / (float)Ho, sx = (float)W / (float)Wo;
  const long long total = (long long)n_img * Ho * Wo * oct;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % oct);
    long long r = i / oct;
    const int xo = (int)(r % Wo);
    const int yo = (int)((r / Wo) % Ho);
    const long long img = r / ((long long)Wo * Ho);
    const int ys = min((int)floorf((float)yo * sy), H - 1), xs = min((int)floorf((float)xo * sx), W - 1);
    const uint4 w = load8_as_bf16<float>(x + ((img * H + ys) * W + xs) * C + c8 * 8);
    *reinterpret_cast<uint4*>(out + r * C + c8 * 8) = w;
  }
}

__global__ void cast_bf16_kernel(const float* __restrict__ x, long long rows, int C_src, int c_offset, int C_total,
                                 op16* __restrict__ out) {
  pdl_prologue();
  const int oct = C_src >> 3;
  const long long total = rows * oct;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / oct;
    const int c8 = (int)(i - r * oct);
    *reinterpret_cast<uint4*>(out + r * C_total + c_offset + c8 * 8) = load8_as_bf16<float>(x + r * C_src + c8 * 8);
  }
}

__global__ void silu_bf16_kernel(const float* __restrict__ x, long long n, op16* __restrict__ out) {
  pdl_prologue();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = float2op16(silu_f(x[i]));
}

// tok [B,F,HW,C] <-> x [B,C,F,HW]
__global__ void tokens_to_ncfhw_kernel(const float* __restrict__ tok, int B, int C, int F, int HW,
                                       float* __restrict__ out) {
  pdl_prologue();
  const long long total = (long long)B * C * F * HW;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int p = (int)(i % HW);
    const int f = (int)((i / HW) % F);
    const int c = (int)((i / ((long long)HW * F)) % C);
    const long long b = i / ((long long)HW * F * C);
    out[i] = tok[((b * F + f) * HW + p) * C + c];
  }
}
__global__ void ncfhw_to_tokens_kernel(const float* __restrict__ x, int B, int C, int F, int HW,
                                       float* __restrict__ out) {
  pdl_prologue();
  const long long total = (long long)B * C * F * HW;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const int p = (int)((i / C) % HW);
    const int f = (int)((i / ((long long)C * HW)) % F);
    const long long b = i / ((long long)C * HW * F);
    out[i] = x[((b * C + c) * F + f) * HW + p];
  }
}

__global__ void add_f32_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out,
                               long long n) {
  pdl_prologue();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = a[i] + b[i];
}

// embeddings.py:28-68 (scale = 1, max_period = 10000): [sin | cos], optionally flipped to [cos | sin]
__global__ void timestep_embedding_kernel(const float* __restrict__ ts, int B, int dim, int flip, float freq_shift,
                                          op16* __restrict__ out) {
  pdl_prologue();
  const int half = dim / 2;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * dim) return;
  const int b = i / dim, j = i % dim;
  float v = 0.f;
  if (j < 2 * half) {
    const int k = j % half;
    const bool second = j >= half;
    const float e = expf(-logf(10000.0f) * (float)k / ((float)half - freq_shift));
    const float arg = ts[b] * e;
    const bool is_sin = flip ? second : !second;
    v = is_sin ? sinf(arg) : cosf(arg);
  }
  out[i] = float2op16(v);
}

// eps = eps_u/cnt [+ g (eps_c/cnt - eps_u/cnt)];  x0 = (x - sqrt(1-a_t) eps)/sqrt(a_t);
// x_prev = sqrt(a_prev) x0 + sqrt(1 - a_prev - sigma^2) eps [+ sigma z]      (DDIM, Song et al. 2021 eq. 12 / 16)
__global__ void cfg_ddim_kernel(float* __restrict__ lat, float* __restrict__ np_u, float* __restrict__ np_c,
                                const float* __restrict__ counter, const float* __restrict__ noise, long long n,
                                int n_frames, long long inner, float gs, float sa_t, float s1a_t, float sa_p, float dir_p,
                                float sigma, int zero_after) {
  pdl_prologue();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float cnt = counter ? counter[(i / inner) % n_frames] : 1.f;
    float eps = np_u[i] / cnt;
    if (np_c) {
      const float ec = np_c[i] / cnt;
      eps = eps + gs * (ec - eps);
    }
    const float x = lat[i];
    const float x0 = (x - s1a_t * eps) / sa_t;
    float xp = sa_p * x0 + dir_p * eps;
    if (noise) xp = fmaf(sigma, noise[i], xp);
    lat[i] = xp;
    if (zero_after) {  // the accumulator is ready for the next timestep's windows (EMOAnimationPipeline.py:702-706)
      np_u[i] = 0.f;
      if (np_c) np_c[i] = 0.f;
    }
  }
}

// Window bookkeeping of the denoise loop (EMOAnimationPipeline.py:759-763, 790-794) as two gather / scatter kernels over
// tensors viewed as [outer, frames, inner] (inner contiguous, a multiple of 4 floats):
//   gather:       dst[o, j, :]              = src[(o % src_mod) + src_off, idx[j], :]
//   scatter-add:  dst[o + dst_off, idx[j], :] += src[o, j, :]          (frames of one window are distinct: no atomics)
__global__ void gather_frames_kernel(const float4* __restrict__ src, float4* __restrict__ dst, const int* __restrict__ idx,
                                     int n_outer, int wlen, int F_src, long long inner4, int src_mod, int src_off) {
  pdl_prologue();
  const long long total = (long long)n_outer * wlen * inner4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long k = i % inner4;
    const int j = (int)((i / inner4) % wlen);
    const int o = (int)(i / (inner4 * wlen));
    const long long so = (o % src_mod) + src_off;
    dst[i] = __ldg(src + (so * F_src + idx[j]) * inner4 + k);
  }
}
__global__ void scatter_add_frames_kernel(const float4* __restrict__ src, float4* __restrict__ dst,
                                          const int* __restrict__ idx, int n_outer, int wlen, int F_dst, long long inner4,
                                          int dst_off) {
  pdl_prologue();
  const long long total = (long long)n_outer * wlen * inner4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long k = i % inner4;
    const int j = (int)((i / inner4) % wlen);
    const int o = (int)(i / (inner4 * wlen));
    float4* d = dst + ((long long)(o + dst_off) * F_dst + idx[j]) * inner4 + k;
    const float4 a = *d, b = __ldg(src + i);
    *d = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
  }
}
__global__ void fill_f32_kernel(float* __restrict__ p, float v, long long n) {
  pdl_prologue();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    p[i] = v;
}

// tok [n, HW, ld>=3] -> [n, 3, HW]
__global__ void vae_post_kernel(const float* __restrict__ tok, int n_img, int HW, int ld, float* __restrict__ of,
                                uint8_t* __restrict__ ou) {
  pdl_prologue();
  const long long total = (long long)n_img * 3 * HW;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int p = (int)(i % HW);
    const int c = (int)((i / HW) % 3);
    const long long img = i / (3LL * HW);
    float v = tok[(img * HW + p) * ld + c] * 0.5f + 0.5f;
    v = fminf(fmaxf(v, 0.f), 1.f);
    if (of) of[i] = v;
    if (ou) ou[i] = (uint8_t)(v * 255.0f);   // truncation, like `(x * 255).numpy().astype(np.uint8)` (magicanimate/utils/util.py:28)
  }
}

// videos [b, c, t, h, w] fp32 -> uint8 [t, Hg, Wg, 3]: torchvision.utils.make_grid (nrow images per row, `pad` black pixels
// around each; a single sample is returned without padding; one channel is replicated to three), optional
// (x + 1) / 2, then the truncating uint8 cast — the frame loop of save_videos_grid (magicanimate/utils/util.py:21-30).
__global__ void video_grid_kernel(const float* __restrict__ v, int b, int c, int t, int h, int w, int xmaps, int pad,
                                  int Hg, int Wg, int rescale, uint8_t* __restrict__ out) {
  pdl_prologue();
  const long long total = (long long)t * Hg * Wg;
  const int cell_h = h + pad, cell_w = w + pad;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int gx = (int)(i % Wg);
    const int gy = (int)((i / Wg) % Hg);
    const int f = (int)(i / ((long long)Wg * Hg));
    int k = -1, py = 0, px = 0;
    if (b == 1) {
      k = 0; py = gy; px = gx;
    } else {
      const int cy = gy / cell_h, cx = gx / cell_w;
      py = gy - cy * cell_h - pad;
      px = gx - cx * cell_w - pad;
      if (py >= 0 && px >= 0 && cx < xmaps && cy * xmaps + cx < b) k = cy * xmaps + cx;
    }
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
      float x = 0.f;
      if (k >= 0) x = v[((((long long)k * c + (c == 1 ? 0 : ch)) * t + f) * h + py) * w + px];
      if (rescale) x = (x + 1.0f) / 2.0f;
      out[i * 3 + ch] = (uint8_t)(int)(x * 255.0f);   // numpy's float -> uint8 cast: truncate, low 8 bits
    }
  }
}

}  // namespace emote

using namespace emote;
#define STREAM(s) reinterpret_cast<cudaStream_t>(s)

extern "C" int emote_latent_im2col(const float* latent, int32_t B, int32_t Cl, int32_t F, int32_t H, int32_t W,
                                   float pre_scale, const float* pw_weight, const float* pw_bias, void* out_bf16,
                                   void* stream) {
  if (!latent || !out_bf16 || B <= 0 || F <= 0 || H <= 0 || W <= 0) return set_error("emote_latent_im2col: bad arguments");
  if (Cl <= 0 || Cl > 7) return set_error("emote_latent_im2col: latent channels must be in [1,7] (9*Cl <= 64)");
  const long long total = (long long)B * F * H * W;
  launch_kernel(latent_im2col_kernel, dim3(grid_for(total, 128)), dim3(128), 0, STREAM(stream), latent, B, Cl, F, H, W, pre_scale, pw_weight, pw_bias, reinterpret_cast<op16*>(out_bf16));
  EMOTE_CHECK_LAUNCH("emote_latent_im2col");
  return 0;
}

template <typename T>
static int im2col_impl(const T* x, int n_img, int H, int W, int C, int stride, void* out, void* stream) {
  if (!x || !out || n_img <= 0 || H <= 0 || W <= 0 || C <= 0 || C % 8 != 0) return set_error("emote_im2col3x3: bad arguments");
  if (stride != 1 && stride != 2) return set_error("emote_im2col3x3: stride must be 1 or 2");
  const int Ho = (H + 2 - 3) / stride + 1, Wo = (W + 2 - 3) / stride + 1;
  const long long total = (long long)n_img * Ho * Wo * 9 * (C / 8);
  launch_kernel(im2col3x3_kernel<T>, dim3(grid_for(total, 256)), dim3(256), 0, STREAM(stream), x, n_img, H, W, C, stride, Ho, Wo,
                                                                       reinterpret_cast<op16*>(out), 1);
  EMOTE_CHECK_LAUNCH("emote_im2col3x3");
  return 0;
}
extern "C" int emote_im2col3x3(const float* x, int32_t n_img, int32_t H, int32_t W, int32_t C, int32_t stride,
                               void* out_bf16, void* stream) {
  return im2col_impl<float>(x, n_img, H, W, C, stride, out_bf16, stream);
}
extern "C" int emote_im2col3x3_s2_pad01(const float* x, int32_t n_img, int32_t H, int32_t W, int32_t C, void* out_bf16,
                                        void* stream) {
  if (!x || !out_bf16 || n_img <= 0 || H <= 0 || W <= 0 || C <= 0 || C % 8 != 0 || H % 2 != 0 || W % 2 != 0)
    return set_error("emote_im2col3x3_s2_pad01: bad arguments (C % 8 == 0, even H and W)");
  const int Ho = H / 2, Wo = W / 2;
  const long long total = (long long)n_img * Ho * Wo * 9 * (C / 8);
  launch_kernel(im2col3x3_kernel<float>, dim3(grid_for(total, 256)), dim3(256), 0, STREAM(stream), x, n_img, H, W, C, 2, Ho, Wo,
                reinterpret_cast<op16*>(out_bf16), 0);
  EMOTE_CHECK_LAUNCH("emote_im2col3x3_s2_pad01");
  return 0;
}
extern "C" int emote_im2col3x3_bf16(const void* x, int32_t n_img, int32_t H, int32_t W, int32_t C, int32_t stride,
                                    void* out_bf16, void* stream) {
  return im2col_impl<op16>(reinterpret_cast<const op16*>(x), n_img, H, W, C, stride, out_bf16, stream);
}

extern "C" int emote_upsample2x(const float* x, int32_t n_img, int32_t H, int32_t W, int32_t C, void* out_bf16,
                                void* stream) {
  if (!x || !out_bf16 || n_img <= 0 || H <= 0 || W <= 0 || C <= 0 || C % 8 != 0) return set_error("emote_upsample2x: bad arguments");
  const long long total = (long long)n_img * 4 * H * W * (C / 8);
  launch_kernel(upsample_nearest_kernel, dim3(grid_for(total, 256)), dim3(256), 0, STREAM(stream), x, n_img, H, W, C,
                2 * H, 2 * W, reinterpret_cast<op16*>(out_bf16));
  EMOTE_CHECK_LAUNCH("emote_upsample2x");
  return 0;
}

extern "C" int emote_upsample_nearest(const float* x, int32_t n_img, int32_t H, int32_t W, int32_t C, int32_t Ho,
                                      int32_t Wo, void* out_bf16, void* stream) {
  if (!x || !out_bf16 || n_img <= 0 || H <= 0 || W <= 0 || Ho <= 0 || Wo <= 0 || C <= 0 || C % 8 != 0)
    return set_error("emote_upsample_nearest: bad arguments (C must be a multiple of 8)");
  const long long total = (long long)n_img * Ho * Wo * (C / 8);
  launch_kernel(upsample_nearest_kernel, dim3(grid_for(total, 256)), dim3(256), 0, STREAM(stream), x, n_img, H, W, C, Ho, Wo,
                reinterpret_cast<op16*>(out_bf16));
  EMOTE_CHECK_LAUNCH("emote_upsample_nearest");
  return 0;
}

extern "C" int emote_cast_bf16(const float* x, int64_t rows, int32_t C_src, int32_t c_offset, int32_t C_total,
                               void* out_bf16, void* stream) {
  if (!x || !out_bf16 || rows <= 0 || C_src <= 0 || C_src % 8 != 0 || c_offset % 8 != 0 || C_total % 8 != 0 ||
      c_offset + C_src > C_total)
    return set_error("emote_cast_bf16: bad arguments");
  launch_kernel(cast_bf16_kernel, dim3(grid_for(rows * (C_src / 8), 256)), dim3(256), 0, STREAM(stream), x, rows, C_src, c_offset, C_total, reinterpret_cast<op16*>(out_bf16));
  EMOTE_CHECK_LAUNCH("emote_cast_bf16");
  return 0;
}

extern "C" int emote_silu_bf16(const float* x, int64_t n, void* out_bf16, void* stream) {
  if (!x || !out_bf16 || n <= 0) return set_error("emote_silu_bf16: bad arguments");
  launch_kernel(silu_bf16_kernel, dim3(grid_for(n, 256)), dim3(256), 0, STREAM(stream), x, n, reinterpret_cast<op16*>(out_bf16));
  EMOTE_CHECK_LAUNCH("emote_silu_bf16");
  return 0;
}

extern "C" int emote_tokens_to_ncfhw(const float* tok, int32_t B, int32_t C, int32_t F, int32_t HW, float* out,
                                     void* stream) {
  if (!tok || !out || B <= 0 || C <= 0 || F <= 0 || HW <= 0) return set_error("emote_tokens_to_ncfhw: bad arguments");
  launch_kernel(tokens_to_ncfhw_kernel, dim3(grid_for((long long)B * C * F * HW, 256)), dim3(256), 0, STREAM(stream), tok, B, C, F, HW, out);
  EMOTE_CHECK_LAUNCH("emote_tokens_to_ncfhw");
  return 0;
}
extern "C" int emote_ncfhw_to_tokens(const float* x, int32_t B, int32_t C, int32_t F, int32_t HW, float* out,
                                     void* stream) {
  if (!x || !out || B <= 0 || C <= 0 || F <= 0 || HW <= 0) return set_error("emote_ncfhw_to_tokens: bad arguments");
  launch_kernel(ncfhw_to_tokens_kernel, dim3(grid_for((long long)B * C * F * HW, 256)), dim3(256), 0, STREAM(stream), x, B, C, F, HW, out);
  EMOTE_CHECK_LAUNCH("emote_ncfhw_to_tokens");
  return 0;
}

extern "C" int emote_add_f32(const float* a, const float* b, float* out, int64_t n, void* stream) {
  if (!a || !b || !out || n <= 0) return set_error("emote_add_f32: bad arguments");
  launch_kernel(add_f32_kernel, dim3(grid_for(n, 256)), dim3(256), 0, STREAM(stream), a, b, out, n);
  EMOTE_CHECK_LAUNCH("emote_add_f32");
  return 0;
}

extern "C" int emote_timestep_embedding(const float* timesteps, int32_t B, int32_t dim, int32_t flip_sin_to_cos,
                                        float freq_shift, void* out_bf16, void* stream) {
  if (!timesteps || !out_bf16 || B <= 0 || dim <= 1) return set_error("emote_timestep_embedding: bad arguments");
  const int n = B * dim;
  launch_kernel(timestep_embedding_kernel, dim3((n + 127) / 128), dim3(128), 0, STREAM(stream), timesteps, B, dim, flip_sin_to_cos, freq_shift,
                                                                         reinterpret_cast<op16*>(out_bf16));
  EMOTE_CHECK_LAUNCH("emote_timestep_embedding");
  return 0;
}

extern "C" int emote_gather_frames(const float* src, float* dst, const int32_t* frame_idx, int32_t n_outer, int32_t wlen,
                                   int32_t F_src, int64_t inner, int32_t src_mod, int32_t src_off, void* stream) {
  if (!src || !dst || !frame_idx || n_outer <= 0 || wlen <= 0 || F_src <= 0 || inner <= 0 || src_mod <= 0 || src_off < 0)
    return set_error("emote_gather_frames: bad arguments");
  if (inner % 4 != 0 || (reinterpret_cast<uintptr_t>(src) & 15) || (reinterpret_cast<uintptr_t>(dst) & 15))
    return set_error("emote_gather_frames: inner must be a multiple of 4 floats and the buffers 16-byte aligned");
  const long long total = (long long)n_outer * wlen * (inner / 4);
  launch_kernel(gather_frames_kernel, dim3(grid_for(total, 256)), dim3(256), 0, STREAM(stream),
                reinterpret_cast<const float4*>(src), reinterpret_cast<float4*>(dst), frame_idx, n_outer, wlen, F_src,
                (long long)(inner / 4), src_mod, src_off);
  EMOTE_CHECK_LAUNCH("emote_gather_frames");
  return 0;
}

extern "C" int emote_scatter_add_frames(const float* src, float* dst, const int32_t* frame_idx, int32_t n_outer,
                                        int32_t wlen, int32_t F_dst, int64_t inner, int32_t dst_off, void* stream) {
  if (!src || !dst || !frame_idx || n_outer <= 0 || wlen <= 0 || F_dst <= 0 || inner <= 0 || dst_off < 0)
    return set_error("emote_scatter_add_frames: bad arguments");
  if (inner % 4 != 0 || (reinterpret_cast<uintptr_t>(src) & 15) || (reinterpret_cast<uintptr_t>(dst) & 15))
    return set_error("emote_scatter_add_frames: inner must be a multiple of 4 floats and the buffers 16-byte aligned");
  const long long total = (long long)n_outer * wlen * (inner / 4);
  launch_kernel(scatter_add_frames_kernel, dim3(grid_for(total, 256)), dim3(256), 0, STREAM(stream),
                reinterpret_cast<const float4*>(src), reinterpret_cast<float4*>(dst), frame_idx, n_outer, wlen, F_dst,
                (long long)(inner / 4), dst_off);
  EMOTE_CHECK_LAUNCH("emote_scatter_add_frames");
  return 0;
}

extern "C" int emote_fill_f32(float* p, float value, int64_t n, void* stream) {
  if (!p || n <= 0) return set_error("emote_fill_f32: bad arguments");
  launch_kernel(fill_f32_kernel, dim3(grid_for(n, 256)), dim3(256), 0, STREAM(stream), p, value, (long long)n);
  EMOTE_CHECK_LAUNCH("emote_fill_f32");
  return 0;
}

static int ddim_launch(float* latents, float* np_u, float* np_c, const float* counter, const float* noise, int64_t n,
                       int32_t n_frames, int64_t inner, float gs, float alpha_t, float alpha_prev, float sigma,
                       int32_t zero_after, void* stream, const char* who) {
  if (!latents || !np_u || n <= 0) return set_error("emote_(cfg_)ddim_step: bad arguments");
  if (counter && (n_frames <= 0 || inner <= 0)) return set_error("emote_(cfg_)ddim_step: bad counter geometry");
  if (!(alpha_t > 0.f && alpha_t <= 1.f && alpha_prev > 0.f && alpha_prev <= 1.f))
    return set_error("emote_(cfg_)ddim_step: alphas must lie in (0,1]");
  if (sigma < 0.f || sigma * sigma > 1.f - alpha_prev + 1e-7f || (sigma > 0.f && !noise))
    return set_error("emote_(cfg_)ddim_step: sigma must satisfy 0 <= sigma^2 <= 1 - alpha_prev and come with a noise tensor");
  const float dir = sqrtf(fmaxf(1.f - alpha_prev - sigma * sigma, 0.f));
  launch_kernel(cfg_ddim_kernel, dim3(grid_for(n, 256)), dim3(256), 0, STREAM(stream), latents, np_u, np_c, counter,
                sigma > 0.f ? noise : (const float*)nullptr, (long long)n, n_frames > 0 ? n_frames : 1,
                (long long)(inner > 0 ? inner : 1), gs, sqrtf(alpha_t), sqrtf(1.f - alpha_t), sqrtf(alpha_prev), dir, sigma,
                zero_after);
  EMOTE_CHECK_LAUNCH(who);
  return 0;
}

extern "C" int emote_cfg_ddim_step(float* latents, float* noise_pred, const float* counter, int64_t n,
                                   int32_t n_frames, int64_t inner, float guidance_scale, float alpha_t,
                                   float alpha_prev, const float* noise, float sigma, int32_t zero_noise_pred,
                                   void* stream) {
  return ddim_launch(latents, noise_pred, noise_pred ? noise_pred + n : nullptr, counter, noise, n, n_frames, inner,
                     guidance_scale, alpha_t, alpha_prev, sigma, zero_noise_pred, stream, "emote_cfg_ddim_step");
}

extern "C" int emote_ddim_step(float* latents, const float* eps, int64_t n, float alpha_t, float alpha_prev,
                               const float* noise, float sigma, void* stream) {
  return ddim_launch(latents, const_cast<float*>(eps), nullptr, nullptr, noise, n, 1, 1, 1.f, alpha_t, alpha_prev, sigma, 0,
                     stream, "emote_ddim_step");
}

extern "C" int emote_vae_postprocess(const float* tok, int32_t n_img, int32_t HW, int32_t ld, float* out_f32,
                                     uint8_t* out_u8, void* stream) {
  if (!tok || n_img <= 0 || HW <= 0 || ld < 3 || (!out_f32 && !out_u8)) return set_error("emote_vae_postprocess: bad arguments");
  launch_kernel(vae_post_kernel, dim3(grid_for((long long)n_img * 3 * HW, 256)), dim3(256), 0, STREAM(stream), tok, n_img, HW, ld, out_f32, out_u8);
  EMOTE_CHECK_LAUNCH("emote_vae_postprocess");
  return 0;
}

extern "C" int emote_video_grid_u8(const float* videos, int32_t b, int32_t c, int32_t t, int32_t h, int32_t w, int32_t nrow,
                                   int32_t padding, int32_t rescale, uint8_t* out, void* stream) {
  if (!videos || !out || b <= 0 || (c != 1 && c != 3) || t <= 0 || h <= 0 || w <= 0 || nrow <= 0 || padding < 0)
    return set_error("emote_video_grid_u8: bad arguments (channels must be 1 or 3)");
  const int xmaps = nrow < b ? nrow : b;
  const int ymaps = (b + xmaps - 1) / xmaps;
  const int Hg = b == 1 ? h : (h + padding) * ymaps + padding;
  const int Wg = b == 1 ? w : (w + padding) * xmaps + padding;
  launch_kernel(video_grid_kernel, dim3(grid_for((long long)t * Hg * Wg, 256)), dim3(256), 0, STREAM(stream), videos, b, c, t,
                h, w, xmaps, padding, Hg, Wg, rescale, out);
  EMOTE_CHECK_LAUNCH("emote_video_grid_u8");
  return 0;
}
