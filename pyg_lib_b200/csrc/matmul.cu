// pyg_lib_b200/csrc/matmul.cu — segment_matmul / grouped_matmul entry points of the C ABI and the
// general (any shape, any stride, fp32-exact) grouped GEMM they fall back to.
//
// Replaces pyg_lib/csrc/ops/cuda/matmul_kernel.cu (CUTLASS 2.x sm80 GemmGrouped, fp32 only):
//   * problems are described ON THE DEVICE (built from `ptr` by a tiny kernel), so a device-resident
//     `ptr` needs no D2H sync (the reference does size_from_ptr(ptr).cpu(), matmul_kernel.cu:307);
//   * one launch walks a flat tile list over all problems (ragged segments load-balance by tiles);
//   * bf16 / fp16 / fp32 storage with fp32 accumulation; optional fused bias.
// The bf16/fp16 fast path for TMA-aligned segment_matmul lives in matmul_tcgen05.cu; this file is the
// path for fp32 ('highest' precision == the reference's SIMT kernels, matmul_kernel.cu:227-282),
// odd shapes, transposed views (grouped_matmul backward) and the weight gradient.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <vector>

#include "common.cuh"

namespace pygb200 {

int segment_matmul_tcgen05(const void* x, const i64* ptr_dev, const void* w, const void* bias, void* out, i64 N, i64 K,
                           i64 M, i64 B, int dtype, cudaStream_t st);  // matmul_tcgen05.cu
bool tcgen05_supported(i64 N, i64 K, i64 M, i64 B, int dtype, const void* x, const void* w, const void* out);

namespace {

struct Problem {       // C[n,m] = A[n,k] @ B[k,m] (+ bias[m]);  element strides
  const void* a; const void* b; void* c; const void* bias;
  i64 n, k, m;
  i64 sa0, sa1, sb0, sb1, ldc;
  i64 tile0;           // index of this problem's first tile in the flat tile list
};

constexpr int BM = 64, BN = 64, BK = 16, MM_NT = 256;

template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <> __device__ __forceinline__ float to_f<__half>(__half v) { return __half2float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ __half from_f<__half>(float v) { return __float2half_rn(v); }

// problems[b] for segment b of a segment_matmul; also the flat tile prefix (single block).
//   mode 0: forward   C_b = X[ptr_b:ptr_b+1] @ W[b]            (n=len, k=K, m=M)
//   mode 1: wgrad     dW[b] = X_b^T @ dY_b                      (n=K, k=len, m=M)
__global__ void k_build_segments(Problem* probs, i64* total_tiles, const i64* __restrict__ ptr, const char* x,
                                 const char* w, const char* bias, char* out, i64 K, i64 M, i64 B, int esize, int mode) {
  __shared__ i64 s_carry;
  __shared__ i64 s_w[32];
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (i64 base = 0; base < B; base += blockDim.x) {
    const i64 b = base + threadIdx.x;
    i64 tiles = 0;
    Problem p;
    if (b < B) {
      const i64 r0 = ptr[b], len = ptr[b + 1] - r0;
      if (mode == 0) {
        p.a = x + r0 * K * esize; p.b = w + b * K * M * esize; p.c = out + r0 * M * esize;
        p.bias = bias ? bias + b * M * esize : nullptr;
        p.n = len; p.k = K; p.m = M; p.sa0 = K; p.sa1 = 1; p.sb0 = M; p.sb1 = 1; p.ldc = M;
      } else {
        p.a = x + r0 * K * esize; p.b = w + r0 * M * esize; p.c = out + b * K * M * esize; p.bias = nullptr;
        p.n = K; p.k = len; p.m = M; p.sa0 = 1; p.sa1 = K; p.sb0 = M; p.sb1 = 1; p.ldc = M;
      }
      tiles = (p.n > 0 && p.m > 0) ? ceil_div(p.n, BM) * ceil_div(p.m, BN) : 0;
    }
    i64 inc = tiles;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const i64 o = __shfl_up_sync(0xffffffffu, inc, d);
      if (lane >= d) inc += o;
    }
    if (lane == 31) s_w[wid] = inc;
    __syncthreads();
    i64 pre = 0, tot = 0;
    for (int q = 0; q < (int)(blockDim.x >> 5); ++q) { if (q < wid) pre += s_w[q]; tot += s_w[q]; }
    const i64 c0 = s_carry;
    if (b < B) { p.tile0 = c0 + pre + inc - tiles; probs[b] = p; }
    __syncthreads();
    if (threadIdx.x == 0) s_carry = c0 + tot;
    __syncthreads();
  }
  if (threadIdx.x == 0) *total_tiles = s_carry;
}

// Generic grouped GEMM: each block takes tiles from the flat list (grid-stride), finds the owning
// problem by binary search on tile0, and computes a 64x64 tile with a 4x4 register micro-tile.
template <typename T>
__global__ void __launch_bounds__(MM_NT) k_grouped_gemm(const Problem* __restrict__ probs, i64 P,
                                                         const i64* __restrict__ total_tiles_ptr) {
  __shared__ float sA[BK][BM + 4];
  __shared__ float sB[BK][BN + 4];
  const i64 total = *total_tiles_ptr;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;  // 16 x 16 threads, each 4x4 outputs
  for (i64 tile = blockIdx.x; tile < total; tile += gridDim.x) {
    i64 lo = 0, hi = P - 1;
    while (lo < hi) {  // last problem with tile0 <= tile (skipping empty problems, which share tile0)
      const i64 mid = (lo + hi + 1) >> 1;
      if (probs[mid].tile0 <= tile) lo = mid; else hi = mid - 1;
    }
    const Problem p = probs[lo];
    const i64 lt = tile - p.tile0, tn = ceil_div(p.m, BN);
    const i64 row0 = (lt / tn) * BM, col0 = (lt % tn) * BN;
    const T* __restrict__ A = (const T*)p.a;
    const T* __restrict__ Bm = (const T*)p.b;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (i64 k0 = 0; k0 < p.k; k0 += BK) {
      // A tile [BM x BK]: pick the thread mapping that walks the contiguous dimension fastest
      for (int e = threadIdx.x; e < BM * BK; e += MM_NT) {
        int r, kk;
        if (p.sa1 == 1) { kk = e % BK; r = e / BK; } else { r = e % BM; kk = e / BM; }
        const i64 gr = row0 + r, gk = k0 + kk;
        sA[kk][r] = (gr < p.n && gk < p.k) ? to_f<T>(A[gr * p.sa0 + gk * p.sa1]) : 0.f;
      }
      for (int e = threadIdx.x; e < BK * BN; e += MM_NT) {
        int c, kk;
        if (p.sb1 == 1) { c = e % BN; kk = e / BN; } else { kk = e % BK; c = e / BK; }
        const i64 gk = k0 + kk, gc = col0 + c;
        sB[kk][c] = (gk < p.k && gc < p.m) ? to_f<T>(Bm[gk * p.sb0 + gc * p.sb1]) : 0.f;
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < BK; ++kk) {
        float a[4], b[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = sA[kk][ty * 4 + i];
#pragma unroll
        for (int j = 0; j < 4; ++j) b[j] = sB[kk][tx * 4 + j];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
      __syncthreads();
    }
    T* __restrict__ C = (T*)p.c;
    const T* __restrict__ bias = (const T*)p.bias;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const i64 gr = row0 + ty * 4 + i;
      if (gr >= p.n) continue;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const i64 gc = col0 + tx * 4 + j;
        if (gc < p.m) {
          float v = acc[i][j];
          if (bias) v += to_f<T>(bias[gc]);
          C[gr * p.ldc + gc] = from_f<T>(v);
        }
      }
    }
  }
}

int launch_grouped(const Problem* probs_dev, i64 P, const i64* total_dev, i64 tiles_bound, int dtype, cudaStream_t st) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  i64 g = tiles_bound < 1 ? 1 : tiles_bound;
  if (g > (i64)sms * 8) g = (i64)sms * 8;
  void* tk = prof_begin(st);
  switch (dtype) {
    case PYGB200_F32: k_grouped_gemm<float><<<(int)g, MM_NT, 0, st>>>(probs_dev, P, total_dev); break;
    case PYGB200_BF16: k_grouped_gemm<__nv_bfloat16><<<(int)g, MM_NT, 0, st>>>(probs_dev, P, total_dev); break;
    case PYGB200_F16: k_grouped_gemm<__half><<<(int)g, MM_NT, 0, st>>>(probs_dev, P, total_dev); break;
    default: set_error("matmul: unknown dtype"); return PYGB200_ERR_ARG;
  }
  prof_end(tk, "grouped_gemm", st, tiles_bound);
  PYGB_LAUNCH_CHECK();
  return PYGB200_OK;
}

inline int esize_of(int dtype) { return dtype == PYGB200_F32 ? 4 : 2; }

int segment_generic(const void* x, const i64* ptr_dev, const void* w, const void* bias, void* out, i64 N, i64 K, i64 M,
                    i64 B, int dtype, int mode, cudaStream_t st) {
  if (B == 0) return PYGB200_OK;
  char* scratch = nullptr;
  const size_t bytes = (size_t)B * sizeof(Problem) + 16;
  PYGB_CUDA(cudaMallocAsync((void**)&scratch, bytes, st));
  Problem* probs = (Problem*)(scratch + 16);
  i64* total = (i64*)scratch;
  k_build_segments<<<1, 1024, 0, st>>>(probs, total, ptr_dev, (const char*)x, (const char*)w, (const char*)bias,
                                       (char*)out, K, M, B, esize_of(dtype), mode);
  PYGB_LAUNCH_CHECK();
  const i64 bound = mode == 0 ? (ceil_div(N, BM) + B) * ceil_div(M, BN) : B * ceil_div(K, BM) * ceil_div(M, BN);
  int rc = launch_grouped(probs, B, total, bound, dtype, st);
  cudaFreeAsync(scratch, st);
  return rc;
}

}  // namespace
}  // namespace pygb200

using namespace pygb200;

extern "C" int pygb200_segment_matmul(const void* x, const int64_t* ptr_dev, const void* w, const void* bias, void* out,
                                      int64_t N, int64_t K, int64_t M, int64_t B, int dtype, unsigned flags,
                                      void* stream) {
  PYGB_CHECK(N >= 0 && K >= 0 && M >= 0 && B >= 0, PYGB200_ERR_ARG, "segment_matmul: negative size");
  PYGB_CHECK(dtype == PYGB200_F32 || dtype == PYGB200_BF16 || dtype == PYGB200_F16, PYGB200_ERR_ARG,
             "segment_matmul: dtype must be f32, bf16 or f16");
  if (N == 0 || M == 0 || B == 0) return PYGB200_OK;
  PYGB_CHECK(x && ptr_dev && w && out, PYGB200_ERR_ARG, "segment_matmul: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  if (K == 0) {
    PYGB_CHECK(bias == nullptr, PYGB200_ERR_UNSUPPORTED, "segment_matmul: K == 0 with bias");
    PYGB_CUDA(cudaMemsetAsync(out, 0, (size_t)N * M * (dtype == PYGB200_F32 ? 4 : 2), st));
    return PYGB200_OK;
  }
  if (!(flags & PYGB200_MM_FORCE_SIMT) && tcgen05_supported(N, K, M, B, dtype, x, w, out))
    return segment_matmul_tcgen05(x, (const i64*)ptr_dev, w, bias, out, N, K, M, B, dtype, st);
  return segment_generic(x, (const i64*)ptr_dev, w, bias, out, N, K, M, B, dtype, 0, st);
}

extern "C" int pygb200_segment_matmul_wgrad(const void* x, const int64_t* ptr_dev, const void* dy, void* dw, int64_t N,
                                            int64_t K, int64_t M, int64_t B, int dtype, unsigned flags, void* stream) {
  (void)flags;
  PYGB_CHECK(N >= 0 && K >= 0 && M >= 0 && B >= 0, PYGB200_ERR_ARG, "segment_matmul_wgrad: negative size");
  PYGB_CHECK(dtype == PYGB200_F32 || dtype == PYGB200_BF16 || dtype == PYGB200_F16, PYGB200_ERR_ARG,
             "segment_matmul_wgrad: dtype must be f32, bf16 or f16");
  if (K == 0 || M == 0 || B == 0) return PYGB200_OK;
  PYGB_CHECK(ptr_dev && dw && (N == 0 || (x && dy)), PYGB200_ERR_ARG, "segment_matmul_wgrad: null pointer");
  // empty segments produce zero gradients: the generic kernel runs the k-loop zero times and stores 0
  return segment_generic(x, (const i64*)ptr_dev, dy, nullptr, dw, N, K, M, B, dtype, 1, (cudaStream_t)stream);
}

extern "C" int pygb200_grouped_matmul(const pygb200_gemm_problem* ps, int64_t count, int dtype, unsigned flags,
                                      void* stream) {
  (void)flags;
  PYGB_CHECK(count >= 0 && (ps || count == 0), PYGB200_ERR_ARG, "grouped_matmul: bad arguments");
  PYGB_CHECK(dtype == PYGB200_F32 || dtype == PYGB200_BF16 || dtype == PYGB200_F16, PYGB200_ERR_ARG,
             "grouped_matmul: dtype must be f32, bf16 or f16");
  if (count == 0) return PYGB200_OK;
  cudaStream_t st = (cudaStream_t)stream;
  std::vector<Problem> h((size_t)count);
  i64 tiles = 0;
  for (i64 i = 0; i < count; ++i) {
    const auto& q = ps[i];
    PYGB_CHECK(q.n >= 0 && q.k >= 0 && q.m >= 0, PYGB200_ERR_ARG, "grouped_matmul: negative size");
    Problem p;
    p.a = q.a; p.b = q.b; p.c = q.c; p.bias = nullptr;
    p.n = q.n; p.k = q.k; p.m = q.m;
    if (q.a_colmajor) { p.sa0 = 1; p.sa1 = q.lda; } else { p.sa0 = q.lda; p.sa1 = 1; }
    if (q.b_colmajor) { p.sb0 = 1; p.sb1 = q.ldb; } else { p.sb0 = q.ldb; p.sb1 = 1; }
    p.ldc = q.ldc;
    p.tile0 = tiles;
    tiles += (q.n > 0 && q.m > 0) ? ceil_div(q.n, BM) * ceil_div(q.m, BN) : 0;
    h[(size_t)i] = p;
  }
  if (tiles == 0) return PYGB200_OK;
  char* scratch = nullptr;
  const size_t bytes = (size_t)count * sizeof(Problem) + 16;
  PYGB_CUDA(cudaMallocAsync((void**)&scratch, bytes, st));
  // pageable source: the copy is staged by the driver before the call returns, so `h` may die here
  PYGB_CUDA(cudaMemcpyAsync(scratch, &tiles, 8, cudaMemcpyHostToDevice, st));
  PYGB_CUDA(cudaMemcpyAsync(scratch + 16, h.data(), (size_t)count * sizeof(Problem), cudaMemcpyHostToDevice, st));
  int rc = launch_grouped((const Problem*)(scratch + 16), count, (const i64*)scratch, tiles, dtype, st);
  cudaFreeAsync(scratch, st);
  return rc;
}
