/*
 * oracle/matmul_oracle.c — CPU restatement of pyg-lib's segment_matmul / grouped_matmul.
 *
 * TEST INFRASTRUCTURE ONLY (see sampler_oracle.c header): never linked into the product.
 *
 * Parity status: PINNED against the reference's own tests (test/ops/test_matmul.py:14-93:
 * out[ptr[i]:ptr[i+1]] == inputs[ptr[i]:ptr[i+1]] @ other[i], atol 1e-6 fp32 / 1e-2 bf16) and
 * against oracle/_ref (the reference's CPU kernel compiled here) in tests/test_oracle_golden.py.
 *
 * Restates pyg_lib/csrc/ops/cpu/matmul_kernel.cpp:410-439 (segment_matmul_kernel: split rows by
 * ptr, one at::matmul_out per segment, :195-201) with the arithmetic contract of ATen's CPU
 * matmul: fp32 accumulation, one rounding to the storage dtype.
 */
#include <stdint.h>
#include <string.h>
#include <omp.h>

static inline float bf16_to_f32(uint16_t h) {
  uint32_t u = ((uint32_t)h) << 16; float f; memcpy(&f, &u, 4); return f;
}
static inline uint16_t f32_to_bf16(float f) { /* round-to-nearest-even, NaN preserved */
  uint32_t u; memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);
  u += 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}
static inline float f16_to_f32(uint16_t h) {
  uint32_t s = (uint32_t)(h & 0x8000u) << 16, e = (h >> 10) & 0x1fu, m = h & 0x3ffu, u;
  if (e == 0) {
    if (m == 0) u = s;
    else { int sh = 0; while (!(m & 0x400u)) { m <<= 1; sh++; } m &= 0x3ffu; u = s | ((uint32_t)(113 - sh) << 23) | (m << 13); }
  } else if (e == 31) u = s | 0x7f800000u | (m << 13);
  else u = s | ((e + 112) << 23) | (m << 13);
  float f; memcpy(&f, &u, 4); return f;
}
static inline uint16_t f32_to_f16(float f) {
  uint32_t x; memcpy(&x, &f, 4);
  uint32_t sign = (x >> 16) & 0x8000u; x &= 0x7fffffffu;
  if (x >= 0x7f800000u) return (uint16_t)(sign | 0x7c00u | (x > 0x7f800000u ? 0x200u : 0));
  if (x >= 0x477ff000u) return (uint16_t)(sign | 0x7c00u);            /* overflow -> inf */
  if (x < 0x33000001u) return (uint16_t)sign;                          /* underflow -> 0 */
  if (x < 0x38800000u) {                                               /* subnormal */
    uint32_t shift = 113 - (x >> 23); uint32_t m = (x & 0x7fffffu) | 0x800000u;
    uint32_t r = m >> (shift + 13); uint32_t rem = m & ((1u << (shift + 13)) - 1), half = 1u << (shift + 12);
    if (rem > half || (rem == half && (r & 1))) r++;
    return (uint16_t)(sign | r);
  }
  uint32_t r = ((x - 0x38000000u) >> 13); uint32_t rem = x & 0x1fffu;
  if (rem > 0x1000u || (rem == 0x1000u && (r & 1))) r++;
  return (uint16_t)(sign | r);
}

/* dtype codes shared with include/pyg_b200.h: 0=f32, 1=bf16, 2=f16 */
static inline float ld(const void* p, int64_t i, int dt) {
  return dt == 0 ? ((const float*)p)[i] : dt == 1 ? bf16_to_f32(((const uint16_t*)p)[i]) : f16_to_f32(((const uint16_t*)p)[i]);
}
static inline void st(void* p, int64_t i, int dt, float v) {
  if (dt == 0) ((float*)p)[i] = v; else if (dt == 1) ((uint16_t*)p)[i] = f32_to_bf16(v); else ((uint16_t*)p)[i] = f32_to_f16(v);
}

/* out[r, :] = x[r, :] @ w[seg(r)]   x:[N,K] w:[B,K,M] out:[N,M] row-major, ptr:[B+1] */
void oracle_segment_matmul(const void* x, const int64_t* ptr, const void* w, void* out, int64_t K, int64_t M,
                           int64_t B, int dtype) {
  for (int64_t b = 0; b < B; ++b) {
    const int64_t r0 = ptr[b], r1 = ptr[b + 1];
#pragma omp parallel for schedule(static)
    for (int64_t r = r0; r < r1; ++r) {
      float acc[1024];
      for (int64_t m0 = 0; m0 < M; m0 += 1024) {
        int64_t mw = M - m0 < 1024 ? M - m0 : 1024;
        for (int64_t m = 0; m < mw; ++m) acc[m] = 0.f;
        for (int64_t k = 0; k < K; ++k) {
          float a = ld(x, r * K + k, dtype);
          const int64_t wb = (b * K + k) * M + m0;
          if (dtype == 0) { const float* wp = (const float*)w + wb; for (int64_t m = 0; m < mw; ++m) acc[m] += a * wp[m]; }
          else for (int64_t m = 0; m < mw; ++m) acc[m] += a * ld(w, wb + m, dtype);
        }
        for (int64_t m = 0; m < mw; ++m) st(out, r * M + m0 + m, dtype, acc[m]);
      }
    }
  }
}

/* grouped_matmul (cpu/matmul_kernel.cpp:281-312): independent [n_i,k_i]@[k_i,m_i], contiguous. */
void oracle_matmul(const void* a, const void* b, void* c, int64_t n, int64_t k, int64_t m, int dtype) {
  int64_t ptr[2] = {0, n};
  oracle_segment_matmul(a, ptr, b, c, k, m, 1, dtype);
}

int oracle_num_threads(void) { return omp_get_max_threads(); }
