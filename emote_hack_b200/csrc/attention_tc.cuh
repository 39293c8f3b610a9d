// Shared pieces of the tcgen05 attention kernels (attention_tc.cu: head_dim 40 / 64 / 80 / 160 multi-head;
// attention_wide.cu: the 512-wide single head of the VAE mid block): launch parameters, packed-FMA / polynomial exp2
// helpers, TMEM load / store wrappers, MN-major UMMA descriptors and the K/V tensor maps.
#pragma once
#include "common.cuh"
#include "emote_b200.h"
#include "host_utils.h"

namespace emote {

struct AttnTcDev {
  const op16 *q, *k0, *v0, *k1, *v1;
  op16* out;
  int heads, d;
  int nq, n0, n1;
  long long q_bs, q_rs, kv0_bs, kv0_rs, kv1_bs, kv1_rs, o_bs, o_rs;
  int kv0_div, kv1_div, kv1_first;
  float scale_log2;
};

constexpr int TC_BQ = 128;
constexpr int TC_BKV = 64;
constexpr int TC_STAGES = 3;   // K/V ring

constexpr int TC_THREADS = 192;  // warp0 loader, warp1 MMA, warps 2-5 softmax

// 16-byte store through the shared window (32-bit address: no generic-address arithmetic in the P-tile loop)
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
__device__ __forceinline__ void cp_async16_zfill(uint32_t dst, const void* src, bool valid) {
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// packed fp32x2 FMA (FFMA2): {d.x, d.y} = {a.x, a.y} * {b, b} + {c, c}
__device__ __forceinline__ void ffma2_bcast(float& d0, float& d1, float a0, float a1, float b, float c) {
  unsigned long long aa, bb, cc, dd;
  asm("mov.b64 %0, {%1, %2};" : "=l"(aa) : "f"(a0), "f"(a1));
  asm("mov.b64 %0, {%1, %1};" : "=l"(bb) : "f"(b));
  asm("mov.b64 %0, {%1, %1};" : "=l"(cc) : "f"(c));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(dd) : "l"(aa), "l"(bb), "l"(cc));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d0), "=f"(d1) : "l"(dd));
}
// {d.x, d.y} = {a.x, a.y} * {b, b} + {c.x, c.y}
__device__ __forceinline__ void ffma2_acc(float& d0, float& d1, float a0, float a1, float b, float c0, float c1) {
  unsigned long long aa, bb, cc, dd;
  asm("mov.b64 %0, {%1, %2};" : "=l"(aa) : "f"(a0), "f"(a1));
  asm("mov.b64 %0, {%1, %1};" : "=l"(bb) : "f"(b));
  asm("mov.b64 %0, {%1, %2};" : "=l"(cc) : "f"(c0), "f"(c1));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(dd) : "l"(aa), "l"(bb), "l"(cc));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d0), "=f"(d1) : "l"(dd));
}
// {d.x, d.y} = {a.x, a.y} * {b.x, b.y} + {c, c}
__device__ __forceinline__ void ffma2_vvb(float& d0, float& d1, float a0, float a1, float b0, float b1, float c) {
  unsigned long long aa, bb, cc, dd;
  asm("mov.b64 %0, {%1, %2};" : "=l"(aa) : "f"(a0), "f"(a1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(bb) : "f"(b0), "f"(b1));
  asm("mov.b64 %0, {%1, %1};" : "=l"(cc) : "f"(c));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(dd) : "l"(aa), "l"(bb), "l"(cc));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d0), "=f"(d1) : "l"(dd));
}
// 2^x for a pair on the FMA / ALU pipes instead of MUFU.EX2 (the softmax loop is MUFU-bound at head_dim 40: 160 FLOP
// per exponential).  Cody-Waite: n = round(x) through the 1.5 * 2^23 magic add, f = x - n in [-0.5, 0.5], 2^f by a
// degree-3 minimax polynomial (max rel. error 7.5e-5, 26x below the bf16 rounding of P), exponent spliced in with one
// shift-add.  x is clamped at -126 (masked keys arrive as -inf and leave as 2^-126 ~ 0).
__device__ __forceinline__ void exp2_poly2(float& r0, float& r1, float x0, float x1) {
  constexpr float MAGIC = 12582912.f;  // 1.5 * 2^23: low mantissa bits of (x + MAGIC) hold round(x)
  x0 = fmaxf(x0, -126.f);
  x1 = fmaxf(x1, -126.f);
  float xf0, xf1, n0, n1, f0, f1, p0, p1;
  ffma2_bcast(xf0, xf1, x0, x1, 1.0f, MAGIC);
  ffma2_bcast(n0, n1, xf0, xf1, 1.0f, -MAGIC);
  ffma2_acc(f0, f1, n0, n1, -1.0f, x0, x1);
  ffma2_bcast(p0, p1, f0, f1, 0.055171505f, 0.24261077f);
  ffma2_vvb(p0, p1, p0, p1, f0, f1, 0.69326097f);
  ffma2_vvb(p0, p1, p0, p1, f0, f1, 0.99992812f);
  r0 = __int_as_float(__float_as_int(p0) + (__float_as_int(xf0) << 23));
  r1 = __int_as_float(__float_as_int(p1) + (__float_as_int(xf1) << 23));
}
__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
__device__ __forceinline__ uint32_t tmem_ld1(uint32_t taddr) {
  uint32_t r;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(taddr) : "memory");
  return r;
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}

// UMMA smem descriptor, MN-major operand, 128B swizzle: rows = K index (128 B = 64 MN elements each), 8-row groups
// 1024 B apart (SBO), 64-element MN atoms `lbo_bytes` apart (LBO).
__device__ __forceinline__ uint64_t umma_desc_sw128_mn(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
__host__ __device__ constexpr uint32_t umma_idesc_op16_bmn(int M, int N) {  // A K-major, B MN-major
  return umma_idesc_op16(M, N) | (1u << 16);
}

constexpr float TC2_TAU = 8.0f;   // lazy-rescale threshold in log2 units: P <= 2^8

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      :
      : "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st1(uint32_t taddr, uint32_t r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr), "r"(r) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// [batch][key][cols] bf16 view starting at `base` (the k or v pointer, i.e. already offset to its first column)
static inline int make_kv_map(CUtensorMap* m, const void* base, int cols, int nkeys, long long row_stride, long long batch_stride,
                       int nbatch) {
  uint64_t dims[3] = {(uint64_t)cols, (uint64_t)nkeys, (uint64_t)nbatch};
  uint64_t strides[2] = {(uint64_t)row_stride * 2, (uint64_t)(nbatch > 1 ? batch_stride : row_stride * nkeys) * 2};
  uint32_t box[3] = {64, (uint32_t)TC_BKV, 1};
  return make_tensor_map(m, base, 3, dims, strides, box);
}


}  // namespace emote
