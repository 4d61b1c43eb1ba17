// pyg_lib_b200/csrc/matmul_grouped_tc.cu — general grouped GEMM on Blackwell tensor cores (bf16 / fp16).
//
// Replaces the reference's TensorOp grouped GEMM (pyg_lib/csrc/ops/cuda/matmul_kernel.cu:121-287: CUTLASS 2.x
// GemmGrouped, any K / M) for everything the specialised segment kernel of matmul_tcgen05.cu does not take:
//   * pyg::grouped_matmul (matmul_kernel.cu:289-302) — independent problems C_i = A_i @ B_i, including the transposed
//     VIEWS its backward passes in (dX = dY @ W^T, dW = X^T @ dY): an operand stored transposed is simply the other
//     UMMA major-ness (A MN-major / B K-major), selected per problem in the instruction descriptor — no copies;
//   * pyg::segment_matmul with K or M outside {64,128,192,256} (any K, M that are multiples of 8 elements — the
//     TMA row-pitch rule), any number of segments; the problem list is built ON THE DEVICE from `ptr`.
// One persistent CTA per SM, 256 threads, warp-specialised like the segment kernel:
//   warp 0 lane 0  TMA producer: per K step of 64 one stage = A tile [128 x 64] + B tile [64 x BN] (BN <= 256), 4-stage
//                  mbarrier ring; out-of-range K / M / rows are zero-filled by TMA, so tails need no special cases;
//   warp 1 lane 0  tcgen05.mma issuer (kind::f16, M = 128, N = BN rounded up to 16), accumulating over the K steps
//                  into one of two TMEM accumulators;
//   warp 2         TMEM allocator;   warps 4-7  epilogue: tcgen05.ld -> (+bias) -> bf16/fp16 -> predicated 16-byte
//                  row stores (a thread owns one output row; 64 contiguous bytes per TMEM load).
// Tiles of all problems form one flat list dealt round-robin to the CTAs.  Per-problem tensor maps live in global
// memory (uploaded with the problem list; the producer fences them in with fence.proxy.tensormap::generic).
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "common.cuh"

namespace pygb200 {
int* mm_error_flag_dev();   // matmul.cu
namespace {

#include "tcgen05_ptx.cuh"

constexpr int GT_TM = 128;       // rows per tile == UMMA M
constexpr int GT_BN = 256;       // max columns per tile == max UMMA N (two accumulators fill the 512 TMEM columns)
constexpr int GT_BK = 64;        // K elements per stage == one 128-byte swizzle row
constexpr int GT_STAGES = 4;
constexpr int GT_NT = 256;
constexpr u32 GT_A_BYTES = GT_TM * GT_BK * 2;            // 16 KB
constexpr u32 GT_B_GROUP = 64 * GT_BK * 2;               // 8 KB per 64 columns of B
constexpr u32 GT_STAGE_BYTES = GT_A_BYTES + (GT_BN / 64) * GT_B_GROUP;   // 48 KB

struct GProb {
  const CUtensorMap* map_a;    // K-major A: dims [k, rows]; MN-major A (transposed view): dims [n, k]
  const CUtensorMap* map_b;    // MN-major B: dims [m, k, z]; K-major B (transposed view): dims [k, m, z]
  unsigned char* c;            // row-major output of this problem
  const unsigned char* bias;   // [m] or null
  i64 ldc;                     // elements
  int n, k, m;
  int a_row0;                  // K-major A: row offset of this problem inside map_a (segment mode)
  int b_z;                     // third coordinate into map_b (segment index)
  int a_mn, b_k;               // operand stored transposed
  int c_vec;                   // rows of C are 16-byte aligned
  int tile0, tiles_n;          // first flat tile of the problem; number of column tiles
  int ksplit, ksteps_chunk;    // split-K: number of K chunks (1 = none) and K steps (of 64) per chunk
  float* part;                 // split-K: [ksplit][n, m] fp32 partial products, added in chunk order by k_gt_finish
};

__device__ __forceinline__ void tma_load_3d(u32 dst, const CUtensorMap* map, int c0, int c1, int c2, u32 bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst),
      "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
      : "memory");
}
// a tensor map that was written to global memory through the generic proxy (the host's copy) must be fenced
// into the tensormap proxy before the TMA unit reads it
__device__ __forceinline__ void fence_tensormap(const CUtensorMap* map) {
  asm volatile("fence.proxy.tensormap::generic.acquire.sys [%0], 128;" ::"l"(map) : "memory");
}

struct TileRef { int p, i, j, kc; };
__device__ __forceinline__ TileRef find_tile(const GProb* __restrict__ probs, int P, int t) {
  int lo = 0, hi = P - 1;
  while (lo < hi) {   // largest p with tile0 <= t (empty problems share their successor's tile0 and are skipped)
    const int mid = (lo + hi + 1) >> 1;
    if (probs[mid].tile0 <= t) lo = mid; else hi = mid - 1;
  }
  const int local = t - probs[lo].tile0, tn = probs[lo].tiles_n, tm = (probs[lo].n + GT_TM - 1) / GT_TM;
  const int mn = local % (tm * tn);   // K chunks outermost: the chunks of one output tile run on different CTAs
  return {lo, mn / tn, mn % tn, local / (tm * tn)};
}

template <bool BF16>
__global__ void __launch_bounds__(GT_NT, 1) k_grouped_tc(const GProb* __restrict__ probs, int P, const int* __restrict__ total_ptr) {
  extern __shared__ unsigned char smem_raw[];
  const u32 raw = smem_u32(smem_raw);
  const u32 base = (raw + 1023u) & ~1023u;
  unsigned char* sm = smem_raw + (base - raw);
  const u32 off_bar = GT_STAGES * GT_STAGE_BYTES;
  const u32 bar0 = base + off_bar;
  auto FULL = [&](int s) { return bar0 + 8u * (u32)s; };
  auto EMPTY = [&](int s) { return bar0 + 8u * (u32)(GT_STAGES + s); };
  auto T_FULL = [&](int s) { return bar0 + 8u * (u32)(2 * GT_STAGES + s); };
  auto T_EMPTY = [&](int s) { return bar0 + 8u * (u32)(2 * GT_STAGES + 2 + s); };
  constexpr int NBARS = 2 * GT_STAGES + 4;
  u32* tmem_slot = reinterpret_cast<u32*>(sm + off_bar + NBARS * 8);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 1 && lane == 0) {
    for (int s = 0; s < GT_STAGES; ++s) { mbar_init(FULL(s), 1); mbar_init(EMPTY(s), 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(T_FULL(s), 1); mbar_init(T_EMPTY(s), 128); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    fence_proxy_async();
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const u32 tmem_base = *tmem_slot;
  const int total = *total_ptr;

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      int stage = 0, last_p = -1;
      u32 phase = 0;
      for (int t = blockIdx.x; t < total; t += gridDim.x) {
        const TileRef tr = find_tile(probs, P, t);
        const GProb& pb = probs[tr.p];
        if (tr.p != last_p) { fence_tensormap(pb.map_a); fence_tensormap(pb.map_b); last_p = tr.p; }
        const int cols = min(GT_BN, pb.m - tr.j * GT_BN), groups = (cols + 63) >> 6;
        const int ks0 = tr.kc * pb.ksteps_chunk, ks1 = min((pb.k + GT_BK - 1) / GT_BK, ks0 + pb.ksteps_chunk);
        for (int ks = ks0; ks < ks1; ++ks) {
          mbar_wait(EMPTY(stage), phase ^ 1);
          const u32 a_dst = base + stage * GT_STAGE_BYTES, b_dst = a_dst + GT_A_BYTES;
          mbar_expect_tx(FULL(stage), GT_A_BYTES + groups * GT_B_GROUP);
          if (!pb.a_mn) {   // box [64 k x 128 rows] -> [128][128 B]
            tma_load_2d(a_dst, pb.map_a, ks * GT_BK, pb.a_row0 + tr.i * GT_TM, FULL(stage));
          } else {          // two boxes [64 n x 64 k]: MN-major atoms, 8 KB per 64 rows of the tile
            tma_load_2d(a_dst, pb.map_a, tr.i * GT_TM, ks * GT_BK, FULL(stage));
            tma_load_2d(a_dst + 8192, pb.map_a, tr.i * GT_TM + 64, ks * GT_BK, FULL(stage));
          }
          for (int gq = 0; gq < groups; ++gq) {
            if (!pb.b_k) tma_load_3d(b_dst + gq * GT_B_GROUP, pb.map_b, tr.j * GT_BN + gq * 64, ks * GT_BK, pb.b_z, FULL(stage));   // [64 k][128 B]
            else tma_load_3d(b_dst + gq * GT_B_GROUP, pb.map_b, ks * GT_BK, tr.j * GT_BN + gq * 64, pb.b_z, FULL(stage));          // [64 cols][128 B]
          }
          if (++stage == GT_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    int stage = 0, acc = 0;
    u32 phase = 0, t_phase = 0;
    const u32 fmt = BF16 ? 1u : 0u;
    for (int t = blockIdx.x; t < total; t += gridDim.x) {
      const TileRef tr = find_tile(probs, P, t);
      const GProb& pb = probs[tr.p];
      const int cols = min(GT_BN, pb.m - tr.j * GT_BN);
      const u32 un = (u32)((cols + 15) & ~15);   // UMMA N: multiple of 16 (columns past m are zero-filled B)
      // D = f32 (bit 4), A/B format (bits 7-9 / 10-12), A major (bit 15), B major (bit 16; 1 = MN-major), N >> 3, M >> 4
      const u32 idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((u32)(pb.a_mn ? 1 : 0) << 15) | ((u32)(pb.b_k ? 0 : 1) << 16) |
                        ((un >> 3) << 17) | ((u32)(GT_TM >> 4) << 24);
      const int ks0 = tr.kc * pb.ksteps_chunk, ksteps = min((pb.k + GT_BK - 1) / GT_BK, ks0 + pb.ksteps_chunk);
      mbar_wait(T_EMPTY(acc), ((t_phase >> acc) & 1u) ^ 1u);
      const u32 d_tmem = tmem_base + (u32)(acc * GT_BN);
      for (int ks = ks0; ks < ksteps; ++ks) {
        mbar_wait(FULL(stage), phase);
        tc_fence_after();
        if (lane == 0) {
          const u32 a_base = base + stage * GT_STAGE_BYTES, b_base = a_base + GT_A_BYTES;
          const int kk_n = min(4, (pb.k - ks * GT_BK + 15) >> 4);   // K = 16 per instruction; skip all-zero steps of the K tail
          for (int kk = 0; kk < kk_n; ++kk) {
            const u64 adesc = pb.a_mn ? make_desc(a_base + kk * 2048, 8192, 1024) : make_desc(a_base + kk * 32, 16, 1024);
            const u64 bdesc = pb.b_k ? make_desc(b_base + kk * 32, 16, 1024) : make_desc(b_base + kk * 2048, GT_B_GROUP, 1024);
            tc_mma_f16(d_tmem, adesc, bdesc, idesc, (ks > ks0 || kk > 0) ? 1u : 0u);
          }
          tc_commit(EMPTY(stage));
          if (ks == ksteps - 1) tc_commit(T_FULL(acc));
        }
        __syncwarp();
        if (++stage == GT_STAGES) { stage = 0; phase ^= 1; }
      }
      t_phase ^= 1u << acc;
      acc ^= 1;
    }
  } else if (warp >= 4) {
    // ================================ epilogue (128 threads) ================================
    const int q = warp & 3, r = q * 32 + lane;
    int acc = 0;
    u32 t_phase = 0;
    for (int t = blockIdx.x; t < total; t += gridDim.x) {
      const TileRef tr = find_tile(probs, P, t);
      const GProb& pb = probs[tr.p];
      const int col0 = tr.j * GT_BN, cols = min(GT_BN, pb.m - col0);
      const i64 grow = (i64)tr.i * GT_TM + r;
      const bool row_ok = grow < pb.n;
      unsigned char* crow = pb.c + (grow * pb.ldc + col0) * 2;
      mbar_wait(T_FULL(acc), (t_phase >> acc) & 1u);
      t_phase ^= 1u << acc;
      tc_fence_after();
      for (int c = 0; c < cols; c += 32) {
        u32 v[32];
        tc_ld_32x32(tmem_base + (u32)(acc * GT_BN + c) + ((u32)(q * 32) << 16), v);
        tc_wait_ld();
        if (!row_ok) continue;
        const int nc = min(32, cols - c);
        if (pb.ksplit > 1) {   // fp32 partial of this K chunk; bias / rounding happen in k_gt_finish
          float* dst = pb.part + ((i64)tr.kc * pb.n + grow) * pb.m + col0 + c;
          if (nc == 32 && (pb.m & 3) == 0) {
#pragma unroll
            for (int j = 0; j < 8; ++j) reinterpret_cast<uint4*>(dst)[j] = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) if (j < nc) dst[j] = __uint_as_float(v[j]);
          }
          continue;
        }
        if (pb.bias) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (j < nc) v[j] = __float_as_uint(__uint_as_float(v[j]) + ld_bias<BF16>(pb.bias, col0 + c + j));
        }
        if (nc == 32 && pb.c_vec) {
          uint4* dst = reinterpret_cast<uint4*>(crow + c * 2);
#pragma unroll
          for (int j = 0; j < 4; ++j)
            dst[j] = make_uint4(pack2<BF16>(__uint_as_float(v[8 * j]), __uint_as_float(v[8 * j + 1])),
                                pack2<BF16>(__uint_as_float(v[8 * j + 2]), __uint_as_float(v[8 * j + 3])),
                                pack2<BF16>(__uint_as_float(v[8 * j + 4]), __uint_as_float(v[8 * j + 5])),
                                pack2<BF16>(__uint_as_float(v[8 * j + 6]), __uint_as_float(v[8 * j + 7])));
        } else {
          unsigned short* dst = reinterpret_cast<unsigned short*>(crow + c * 2);
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (j < nc) dst[j] = (unsigned short)(pack2<BF16>(__uint_as_float(v[j]), 0.f) & 0xffffu);
        }
      }
      tc_fence_before();
      mbar_arrive(T_EMPTY(acc));
      acc ^= 1;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// segment mode: problems[b] for segment b of a segment_matmul (all share map_a / map_b), flat tile prefix; one block
__global__ void k_gt_build_segments(GProb* probs, int* total, const i64* __restrict__ ptr, const CUtensorMap* map_a,
                                    const CUtensorMap* map_b, unsigned char* out, const unsigned char* bias, int K, int M, int B,
                                    int c_vec, i64 N, int* err) {
  __shared__ int s_carry;
  __shared__ int s_w[32];
  if (threadIdx.x == 0) s_carry = 0;
  {   // same validation of `ptr` as the other segment kernels (build_tile_prefix, matmul_tcgen05.cu)
    int bad = 0;
    for (int b = threadIdx.x; b < B; b += blockDim.x) bad |= (ptr[b + 1] < ptr[b]) | (b == 0 && ptr[0] != 0) | (b == B - 1 && ptr[B] != N);
    if (__syncthreads_or(bad)) {
      if (threadIdx.x == 0) { *total = 0; if (err) *err = 1; }
      return;
    }
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int tiles_n = (M + GT_BN - 1) / GT_BN;
  for (int base = 0; base < B; base += blockDim.x) {
    const int b = base + threadIdx.x;
    int tiles = 0;
    GProb p;
    if (b < B) {
      const i64 r0 = ptr[b], len = ptr[b + 1] - r0;
      p.map_a = map_a; p.map_b = map_b; p.c = out + r0 * M * 2; p.bias = bias ? bias + (i64)b * M * 2 : nullptr;
      p.ldc = M; p.n = (int)len; p.k = K; p.m = M; p.a_row0 = (int)r0; p.b_z = b; p.a_mn = 0; p.b_k = 0; p.c_vec = c_vec;
      p.tiles_n = tiles_n; p.ksplit = 1; p.ksteps_chunk = (K + GT_BK - 1) / GT_BK; p.part = nullptr;
      tiles = (int)((len + GT_TM - 1) / GT_TM) * tiles_n;
    }
    int inc = tiles;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int o = __shfl_up_sync(0xffffffffu, inc, d);
      if (lane >= d) inc += o;
    }
    if (lane == 31) s_w[wid] = inc;
    __syncthreads();
    int pre = 0, tot = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { if (w < wid) pre += s_w[w]; tot += s_w[w]; }
    const int c0 = s_carry;
    if (b < B) { p.tile0 = c0 + pre + inc - tiles; probs[b] = p; }
    __syncthreads();
    if (threadIdx.x == 0) s_carry = c0 + tot;
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = s_carry;
}

// split-K epilogue: C = sum over the K chunks' partials, in chunk order (bit-reproducible), rounded to the storage type
template <bool BF16>
__global__ void k_gt_finish(const GProb* __restrict__ probs, int P) {
  for (int b = blockIdx.y; b < P; b += gridDim.y) {
    const GProb pb = probs[b];
    if (pb.ksplit <= 1) continue;
    const i64 nm = (i64)pb.n * pb.m;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < nm; i += (i64)gridDim.x * blockDim.x) {
      float v = 0.f;
      for (int c = 0; c < pb.ksplit; ++c) v += pb.part[(i64)c * nm + i];
      const i64 r = i / pb.m, col = i - r * pb.m;
      if (pb.bias) v += ld_bias<BF16>(pb.bias, col);
      reinterpret_cast<unsigned short*>(pb.c)[r * pb.ldc + col] = (unsigned short)(pack2<BF16>(v, 0.f) & 0xffffu);
    }
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn gt_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// 16-bit tensor, innermost dimension contiguous: dims / strides (bytes, for dims 1..) / box, 128-byte swizzle, zero fill
int gt_map(CUtensorMap* m, const void* ptr, bool bf16, int rank, const cuuint64_t* dims, const cuuint64_t* strides, const cuuint32_t* box) {
  EncodeTiledFn enc = gt_encode();
  PYGB_CHECK(enc != nullptr, PYGB200_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint32_t estr[3] = {1, 1, 1};
  const CUresult r = enc(m, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, (cuuint32_t)rank,
                         const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with code " + std::to_string((int)r));
    return PYGB200_ERR_CUDA;
  }
  return PYGB200_OK;
}

constexpr size_t GT_SMEM = 1024 + (size_t)GT_STAGES * GT_STAGE_BYTES + (2 * GT_STAGES + 4) * 8 + 16;

int gt_launch(const GProb* probs, int P, const int* total, i64 tiles_bound, bool bf16, cudaStream_t st) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = (int)std::max<i64>(1, std::min<i64>(tiles_bound, sms));
  void* tk = prof_begin(st);
  if (bf16) {
    PYGB_CUDA(cudaFuncSetAttribute(k_grouped_tc<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GT_SMEM));
    k_grouped_tc<true><<<grid, GT_NT, GT_SMEM, st>>>(probs, P, total);
  } else {
    PYGB_CUDA(cudaFuncSetAttribute(k_grouped_tc<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GT_SMEM));
    k_grouped_tc<false><<<grid, GT_NT, GT_SMEM, st>>>(probs, P, total);
  }
  prof_end(tk, "grouped_gemm", st, tiles_bound);
  PYGB_LAUNCH_CHECK();
  return PYGB200_OK;
}

inline bool al16(const void* p) { return ((uintptr_t)p & 15) == 0; }

}  // namespace

// ---- pyg::grouped_matmul: can every problem go through the tensor cores?  (TMA: 16-byte aligned bases and row pitches)
bool grouped_tc_supported(const pygb200_gemm_problem* ps, i64 count, int dtype) {
  if (dtype != PYGB200_BF16 && dtype != PYGB200_F16) return false;
  static const bool off = getenv("PYGB200_NO_GROUPED_TC") != nullptr;
  if (off || count < 1 || count > 65536) return false;
  i64 tiles = 0;
  for (i64 i = 0; i < count; ++i) {
    const auto& q = ps[i];
    if (q.n == 0 || q.m == 0) continue;
    if (q.n >= ((i64)1 << 31) || q.k >= ((i64)1 << 31) || q.m >= ((i64)1 << 31)) return false;
    if (q.k == 0) continue;   // empty contraction: the output is zeroed by a memset, no tile is scheduled
    if (!al16(q.a) || !al16(q.b) || q.lda % 8 != 0 || q.ldb % 8 != 0) return false;
    // the extent along the contiguous dimension must fit the pitch
    if (q.lda < (q.a_colmajor ? q.n : q.k) || q.ldb < (q.b_colmajor ? q.k : q.m) || q.ldc < q.m) return false;
    tiles += ((q.n + GT_TM - 1) / GT_TM) * ((q.m + GT_BN - 1) / GT_BN);
  }
  return tiles > 0 && tiles < ((i64)1 << 31);
}

int grouped_matmul_tc(const pygb200_gemm_problem* ps, i64 count, int dtype, cudaStream_t st) {
  const bool bf16 = dtype == PYGB200_BF16;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  // Split-K plan: a weight-gradient-shaped list (few output tiles, very long K: dW = X^T dY of a HeteroDictLinear) would
  // keep a handful of CTAs busy for thousands of K steps.  When the output tiles do not fill the machine twice, every
  // problem's K steps are cut into chunks (>= 8 steps = 512 of K each) that become work items of their own; chunk c
  // stores an fp32 partial and k_gt_finish adds the partials in order.  256 MB of partials at most.
  i64 base_tiles = 0;
  for (i64 i = 0; i < count; ++i)
    if (ps[i].n > 0 && ps[i].m > 0 && ps[i].k > 0) base_tiles += ((ps[i].n + GT_TM - 1) / GT_TM) * ((ps[i].m + GT_BN - 1) / GT_BN);
  const i64 want = base_tiles > 0 && base_tiles < 2 * (i64)sms ? (2 * (i64)sms + base_tiles - 1) / base_tiles : 1;
  std::vector<int> ksplit((size_t)count, 1);
  size_t part_bytes = 0;
  for (i64 i = 0; i < count && want > 1; ++i) {
    const auto& q = ps[i];
    if (q.n <= 0 || q.m <= 0 || q.k <= 0) continue;
    const i64 ksteps = (q.k + GT_BK - 1) / GT_BK;
    const i64 sp = std::max<i64>(1, std::min<i64>(want, ksteps / 8));
    const size_t bytes = (size_t)sp * q.n * q.m * 4;
    if (sp > 1 && part_bytes + bytes <= ((size_t)256 << 20)) { ksplit[(size_t)i] = (int)sp; part_bytes += (bytes + 255) & ~(size_t)255; }
  }
  // one upload: [tensor maps: 2 per problem | problem list | tile total]; partials behind it
  const size_t map_bytes = (size_t)count * 2 * sizeof(CUtensorMap);
  const size_t prob_off = (map_bytes + 255) & ~(size_t)255, total_off = prob_off + (((size_t)count * sizeof(GProb) + 15) & ~(size_t)15);
  const size_t part_off = (total_off + 16 + 255) & ~(size_t)255;
  std::vector<unsigned char> h(total_off + 16, 0);
  AsyncScratch sc;   // (freed on every return path)
  if (int e = sc.alloc(part_off + part_bytes, st)) return e;
  unsigned char* dbuf = (unsigned char*)sc.p;
  CUtensorMap* hmaps = reinterpret_cast<CUtensorMap*>(h.data());
  GProb* hp = reinterpret_cast<GProb*>(h.data() + prob_off);
  const CUtensorMap* dmaps = reinterpret_cast<const CUtensorMap*>(dbuf);
  i64 tiles = 0;
  size_t part_used = 0;
  bool any_split = false;
  int rc = PYGB200_OK;
  for (i64 i = 0; i < count && rc == PYGB200_OK; ++i) {
    const auto& q = ps[i];
    GProb p;
    memset(&p, 0, sizeof(p));
    p.tile0 = (int)tiles; p.tiles_n = 1; p.ksplit = 1; p.ksteps_chunk = 1;
    p.n = (int)q.n; p.k = (int)q.k; p.m = (int)q.m;
    if (q.n > 0 && q.m > 0 && q.k == 0) {   // empty contraction: zeros
      if (cudaMemset2DAsync(q.c, (size_t)q.ldc * 2, 0, (size_t)q.m * 2, (size_t)q.n, st) != cudaSuccess) rc = PYGB200_ERR_CUDA;
    } else if (q.n > 0 && q.m > 0) {
      CUtensorMap ma, mb;
      cuuint64_t d[3], s[2]; cuuint32_t box[3];
      if (!q.a_colmajor) { d[0] = (cuuint64_t)q.k; d[1] = (cuuint64_t)q.n; s[0] = (cuuint64_t)q.lda * 2; box[0] = 64; box[1] = GT_TM; }
      else { d[0] = (cuuint64_t)q.n; d[1] = (cuuint64_t)q.k; s[0] = (cuuint64_t)q.lda * 2; box[0] = 64; box[1] = 64; }
      rc = gt_map(&ma, q.a, bf16, 2, d, s, box);
      if (rc != PYGB200_OK) break;
      if (!q.b_colmajor) { d[0] = (cuuint64_t)q.m; d[1] = (cuuint64_t)q.k; }
      else { d[0] = (cuuint64_t)q.k; d[1] = (cuuint64_t)q.m; }
      d[2] = 1; s[0] = (cuuint64_t)q.ldb * 2; s[1] = s[0] * d[1]; box[0] = 64; box[1] = 64; box[2] = 1;
      rc = gt_map(&mb, q.b, bf16, 3, d, s, box);
      if (rc != PYGB200_OK) break;
      hmaps[2 * i] = ma; hmaps[2 * i + 1] = mb;
      p.map_a = dmaps + 2 * i; p.map_b = dmaps + 2 * i + 1;
      p.c = reinterpret_cast<unsigned char*>(q.c); p.bias = nullptr; p.ldc = q.ldc;
      p.a_mn = q.a_colmajor ? 1 : 0; p.b_k = q.b_colmajor ? 1 : 0;
      p.c_vec = (al16(q.c) && q.ldc % 8 == 0) ? 1 : 0;
      p.tiles_n = (int)((q.m + GT_BN - 1) / GT_BN);
      const int ksteps = (int)((q.k + GT_BK - 1) / GT_BK);
      p.ksplit = ksplit[(size_t)i];
      p.ksteps_chunk = (ksteps + p.ksplit - 1) / p.ksplit;
      p.ksplit = (ksteps + p.ksteps_chunk - 1) / p.ksteps_chunk;   // (no empty chunk)
      if (p.ksplit > 1) {
        p.part = reinterpret_cast<float*>(dbuf + part_off + part_used);
        part_used += ((size_t)p.ksplit * q.n * q.m * 4 + 255) & ~(size_t)255;
        any_split = true;
      }
      tiles += ((q.n + GT_TM - 1) / GT_TM) * p.tiles_n * p.ksplit;
    }
    hp[i] = p;
  }
  if (rc == PYGB200_OK && tiles > 0) {
    *reinterpret_cast<int*>(h.data() + total_off) = (int)tiles;
    // pageable source: the copy is staged by the driver before the call returns, so `h` may die at the end of the scope
    if (cudaMemcpyAsync(dbuf, h.data(), h.size(), cudaMemcpyHostToDevice, st) != cudaSuccess) rc = PYGB200_ERR_CUDA;
    if (rc == PYGB200_OK) rc = gt_launch(reinterpret_cast<const GProb*>(dbuf + prob_off), (int)count, reinterpret_cast<const int*>(dbuf + total_off), tiles, bf16, st);
    if (rc == PYGB200_OK && any_split) {
      const dim3 grid(64, (unsigned)std::min<i64>(count, 4096));
      if (bf16) k_gt_finish<true><<<grid, 256, 0, st>>>(reinterpret_cast<const GProb*>(dbuf + prob_off), (int)count);
      else k_gt_finish<false><<<grid, 256, 0, st>>>(reinterpret_cast<const GProb*>(dbuf + prob_off), (int)count);
      count_launch();
      if (cudaGetLastError() != cudaSuccess) rc = PYGB200_ERR_CUDA;
    }
  }
  return rc;
}

// ---- pyg::segment_matmul for the shapes the specialised kernel does not take
bool segment_tc_general_supported(i64 N, i64 K, i64 M, i64 B, int dtype, const void* x, const void* w, const void* out) {
  if (dtype != PYGB200_BF16 && dtype != PYGB200_F16) return false;
  static const bool off = getenv("PYGB200_NO_GROUPED_TC") != nullptr;
  if (off) return false;
  if (K < 1 || M < 1 || K % 8 != 0 || M % 8 != 0) return false;
  if (B < 1 || N < 1 || N >= ((i64)1 << 31) || B >= ((i64)1 << 24) || K >= ((i64)1 << 31) || M >= ((i64)1 << 31)) return false;
  if ((N / GT_TM + B) * ((M + GT_BN - 1) / GT_BN) >= ((i64)1 << 31)) return false;
  return al16(x) && al16(w) && al16(out);
}

int segment_matmul_tc_general(const void* x, const i64* ptr_dev, const void* w, const void* bias, void* out, i64 N, i64 K, i64 M,
                              i64 B, int dtype, cudaStream_t st) {
  const bool bf16 = dtype == PYGB200_BF16;
  CUtensorMap hm[2];
  {
    cuuint64_t d[3] = {(cuuint64_t)K, (cuuint64_t)N, 1}, s[2] = {(cuuint64_t)K * 2, 0};
    cuuint32_t box[3] = {64, GT_TM, 1};
    if (int e = gt_map(&hm[0], x, bf16, 2, d, s, box)) return e;
    cuuint64_t dw[3] = {(cuuint64_t)M, (cuuint64_t)K, (cuuint64_t)B}, sw[2] = {(cuuint64_t)M * 2, (cuuint64_t)K * M * 2};
    cuuint32_t bw[3] = {64, 64, 1};
    if (int e = gt_map(&hm[1], w, bf16, 3, dw, sw, bw)) return e;
  }
  const size_t prob_off = 256, total_off = prob_off + (((size_t)B * sizeof(GProb) + 15) & ~(size_t)15);
  AsyncScratch sc;
  if (int e = sc.alloc(total_off + 16, st)) return e;
  unsigned char* dbuf = (unsigned char*)sc.p;
  int rc = PYGB200_OK;
  if (cudaMemcpyAsync(dbuf, hm, sizeof(hm), cudaMemcpyHostToDevice, st) != cudaSuccess) rc = PYGB200_ERR_CUDA;
  if (rc == PYGB200_OK) {
    const CUtensorMap* dm = reinterpret_cast<const CUtensorMap*>(dbuf);
    k_gt_build_segments<<<1, 1024, 0, st>>>(reinterpret_cast<GProb*>(dbuf + prob_off), reinterpret_cast<int*>(dbuf + total_off), ptr_dev, dm,
                                            dm + 1, reinterpret_cast<unsigned char*>(out), reinterpret_cast<const unsigned char*>(bias),
                                            (int)K, (int)M, (int)B, 1 /* rows of out are M * 2 bytes apart, M % 8 == 0 */, N, mm_error_flag_dev());
    count_launch();
    if (cudaGetLastError() != cudaSuccess) rc = PYGB200_ERR_CUDA;
  }
  if (rc == PYGB200_OK)
    rc = gt_launch(reinterpret_cast<const GProb*>(dbuf + prob_off), (int)B, reinterpret_cast<const int*>(dbuf + total_off),
                   (N / GT_TM + B) * ((M + GT_BN - 1) / GT_BN), bf16, st);
  return rc;
}

}  // namespace pygb200
