// pyg_lib_b200/csrc/matmul_tcgen05.cu — segment_matmul for bf16/fp16 on Blackwell tensor cores.
//
// Replaces the reference's CUTLASS 2.x sm80 grouped GEMM (pyg_lib/csrc/ops/cuda/matmul_kernel.cu:167-190,
// mma.sync TF32, fp32 storage only) with a hand-written persistent sm_100a kernel:
//
//   * one CTA per SM, warp-specialised: warp 0 = TMA producer, warp 1 = tcgen05.mma issuer,
//     warp 2 = TMEM allocator, warps 4-7 = epilogue (TMEM -> registers -> bf16 -> smem -> TMA store);
//   * the flat tile list (ceil(len_b / 128) row tiles per segment) is derived on the device from `ptr`
//     (no host sync); every CTA takes a contiguous chunk so consecutive tiles share W[b];
//   * A row tiles [128 x K] stream through a 3-stage TMA ring (SWIZZLE_128B, K-major);
//     W[b] ([K x M], M contiguous == "MN-major" B operand) is TMA-loaded once per segment into a
//     double-buffered slot and consumed in place — no transpose pass;
//   * accumulators live in TMEM (two [128 x M] fp32 buffers) so the epilogue of tile t overlaps the
//     MMAs of tile t+1; bias is fused in the epilogue;
//   * ragged tails: a tile never crosses a segment end; TMA loads may over-read rows of the next
//     segment (harmless), full tiles leave through TMA stores, partial tiles through predicated
//     16-byte row stores.
//
// The shape is HBM-bound (AI = 63.75 FLOP/B at K=M=128), so the design goal is bytes in flight per SM,
// not MMA issue rate: 3 x 32 KB of A loads are outstanding while one tile computes.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"

namespace pygb200 {
int* mm_error_flag_dev();   // matmul.cu: device address of the mapped "invalid ptr" flag
namespace {

constexpr int TM = 128;            // rows per tile == UMMA M == TMEM lanes
constexpr int MAX_SEG = 1024;      // segments handled by the in-kernel tile prefix (else generic path)
constexpr int NTHREADS = 256;

#include "tcgen05_ptx.cuh"

struct SegParams {
  const i64* ptr;
  const void* bias;   // [B, M] or null
  void* out;          // [N, M]
  i64 N;
  int K, M, B;
  int G;              // tiles per scheduling chunk (chunks are dealt round-robin to CTAs); 0 = one chunk per CTA
  int* err;           // mapped host flag: raised when `ptr` is not a valid segment pointer
};

// Tile order of one CTA: chunks of G consecutive tiles, chunk c -> CTA c % gridDim.  All CTAs therefore
// sweep the same moving window of the row space (DRAM page locality, like a grid-stride copy) while
// consecutive tiles of a CTA still share W[b].
struct TileIter {
  int G, grid, total, chunk, t, end;
  __device__ TileIter(int G_, int grid_, int total_, int bid) : G(G_), grid(grid_), total(total_), chunk(bid) {
    t = chunk * G; end = min(total, t + G);
  }
  __device__ bool valid() const { return t < total; }
  __device__ int peek_next() const {   // tile after the current one (>= total if none)
    if (t + 1 < end) return t + 1;
    const long long n = (long long)(chunk + grid) * G;
    return n < total ? (int)n : total;
  }
  __device__ void next() {
    if (++t >= end) { chunk += grid; const long long n = (long long)chunk * G; t = n < total ? (int)n : total; end = min(total, t + G); }
  }
};

// Prologue shared by the three segment kernels: tile_pre[b] = number of 128-row tiles of the segments before b
// (tile_pre[B] = total), from `ptr` on the device — the reference does size_from_ptr(ptr).cpu() here
// (matmul_kernel.cu:307).  All NTHREADS threads; a later __syncthreads() of the caller publishes the result.  `ptr` is
// validated on the way (ptr[0] == 0, non-decreasing, ptr[B] == N — what the reference's split_with_sizes enforces on the
// host): a bad ptr leaves the kernel without work and raises the mapped flag `err`, which the next matmul call reports.
__device__ __forceinline__ void build_tile_prefix(const i64* __restrict__ ptr, int B, i64 N, int* tile_pre, int* err) {
  __shared__ int s_part[NTHREADS];
  const int per = (B + NTHREADS - 1) / NTHREADS;
  int loc = 0, bad = 0;
  for (int j = 0; j < per; ++j) {
    const int b = threadIdx.x * per + j;
    if (b < B) {
      const i64 lo = ptr[b], hi = ptr[b + 1];
      bad |= (hi < lo) | (b == 0 && lo != 0) | (b == B - 1 && hi != N);
      loc += (int)((hi - lo + TM - 1) / TM);
    }
  }
  s_part[threadIdx.x] = loc;
  if (__syncthreads_or(bad)) {   // (also the barrier that publishes s_part)
    for (int b = threadIdx.x; b <= B; b += NTHREADS) tile_pre[b] = 0;
    if (threadIdx.x == 0 && blockIdx.x == 0 && err) *err = 1;
    return;
  }
  int pre = 0;
  for (int t = 0; t < (int)threadIdx.x; ++t) pre += s_part[t];
  for (int j = 0; j < per; ++j) {
    const int b = threadIdx.x * per + j;
    if (b < B) {
      tile_pre[b] = pre;
      pre += (int)((ptr[b + 1] - ptr[b] + TM - 1) / TM);
    }
  }
  if (threadIdx.x == 0) {
    int tot = 0;
    for (int t = 0; t < NTHREADS; ++t) tot += s_part[t];
    tile_pre[B] = tot;
  }
}

// dynamic smem layout (1024-aligned): A ring | W double buffer | out staging | barriers | tile prefix
template <bool BF16, int A_STAGES, int W_BUFS, int O_BUFS>
__global__ void __launch_bounds__(NTHREADS, 1)
k_segment_matmul_tc(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w,
                    const __grid_constant__ CUtensorMap map_o, const SegParams P) {
  extern __shared__ unsigned char smem_raw[];
  const u32 raw = smem_u32(smem_raw);
  const u32 base = (raw + 1023u) & ~1023u;
  unsigned char* sm = smem_raw + (base - raw);
  const int K = P.K, M = P.M, KH = K / 64, NH = M / 64;
  const u32 a_bytes = TM * K * 2, w_bytes = K * M * 2, o_bytes = TM * M * 2;
  const u32 off_a = 0, off_w = off_a + A_STAGES * a_bytes, off_o = off_w + W_BUFS * w_bytes, off_bar = off_o + O_BUFS * o_bytes;
  u64* bars = reinterpret_cast<u64*>(sm + off_bar);
  // barrier indices
  const u32 bar0 = base + off_bar;
  auto A_FULL = [&](int s) { return bar0 + 8u * (u32)s; };
  auto A_EMPTY = [&](int s) { return bar0 + 8u * (u32)(A_STAGES + s); };
  auto W_FULL = [&](int s) { return bar0 + 8u * (u32)(2 * A_STAGES + s); };
  auto W_EMPTY = [&](int s) { return bar0 + 8u * (u32)(2 * A_STAGES + 2 + s); };
  auto T_FULL = [&](int s) { return bar0 + 8u * (u32)(2 * A_STAGES + 4 + s); };
  auto T_EMPTY = [&](int s) { return bar0 + 8u * (u32)(2 * A_STAGES + 6 + s); };
  constexpr int NBARS = 2 * A_STAGES + 8;
  u32* tmem_slot = reinterpret_cast<u32*>(bars + NBARS);
  int* tile_pre = reinterpret_cast<int*>(tmem_slot + 2);  // [B + 1] exclusive prefix of tiles per segment

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const u32 tmem_cols = (2 * M <= 32) ? 32 : (2 * M <= 64) ? 64 : (2 * M <= 128) ? 128 : (2 * M <= 256) ? 256 : 512;

  // ---- prologue: tile prefix over segments (all threads), barriers, TMEM
  build_tile_prefix(P.ptr, P.B, P.N, tile_pre, P.err);
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < A_STAGES; ++s) { mbar_init(A_FULL(s), 1); mbar_init(A_EMPTY(s), 1); }
    for (int s = 0; s < 2; ++s) {
      mbar_init(W_FULL(s), 1); mbar_init(W_EMPTY(s), 1);
      mbar_init(T_FULL(s), 1); mbar_init(T_EMPTY(s), 128);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    fence_proxy_async();
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_o) : "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const u32 tmem_base = *tmem_slot;

  const int total_tiles = tile_pre[P.B];
  const int per_cta = (total_tiles + (int)gridDim.x - 1) / (int)gridDim.x;
  const int G = P.G > 0 ? P.G : max(per_cta, 1);
  const bool has_work = (long long)blockIdx.x * G < total_tiles;

  // segment lookup: largest b with tile_pre[b] <= t and a non-empty segment (tile_pre[b+1] > t)
  auto seg_of = [&](int t) {
    int lo = 0, hi = P.B - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (tile_pre[mid] <= t) lo = mid; else hi = mid - 1;
    }
    return lo;
  };

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (lane == 0 && has_work) {
      TileIter it(G, (int)gridDim.x, total_tiles, (int)blockIdx.x);
      int seg = seg_of(it.t), cur_seg = -1, wbuf = W_BUFS - 1, a_stage = 0;
      u32 a_phase = 0, w_phase = 0;   // w_phase bit i = parity of W slot i
      for (; it.valid(); it.next()) {
        const int t = it.t;
        while (tile_pre[seg + 1] <= t) ++seg;
        if (seg != cur_seg) {
          cur_seg = seg;
          wbuf = (wbuf + 1 == W_BUFS) ? 0 : wbuf + 1;
          mbar_wait(W_EMPTY(wbuf), ((w_phase >> wbuf) & 1u) ^ 1u);
          w_phase ^= 1u << wbuf;
          mbar_expect_tx(W_FULL(wbuf), w_bytes);
          for (int h = 0; h < NH; ++h)  // box [64 cols of M x K rows] -> [K][128 B], MN-major SW128 atoms
            tma_load_2d(base + off_w + wbuf * w_bytes + h * (K * 128), &map_w, h * 64, seg * K, W_FULL(wbuf));
        }
        const i64 row0 = P.ptr[seg] + (i64)(t - tile_pre[seg]) * TM;
        mbar_wait(A_EMPTY(a_stage), a_phase ^ 1);
        mbar_expect_tx(A_FULL(a_stage), a_bytes);
        for (int h = 0; h < KH; ++h)  // box [64 cols of K x 128 rows] -> [128][128 B], K-major SW128 atoms
          tma_load_2d(base + off_a + a_stage * a_bytes + h * (TM * 128), &map_a, h * 64, (int)row0, A_FULL(a_stage));
        if (++a_stage == A_STAGES) { a_stage = 0; a_phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    if (has_work) {
      // instruction descriptor: D=f32 (bit 4), A/B format (bits 7-9, 10-12: 1 = bf16, 0 = f16),
      // A K-major (bit 15 = 0), B MN-major (bit 16 = 1), N>>3 at bit 17, M>>4 at bit 24
      const u32 fmt = BF16 ? 1u : 0u;
      const u32 idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | (1u << 16) | ((u32)(M >> 3) << 17) | ((u32)(TM >> 4) << 24);
      TileIter it(G, (int)gridDim.x, total_tiles, (int)blockIdx.x);
      int seg = seg_of(it.t), cur_seg = -1, wbuf = W_BUFS - 1, a_stage = 0, acc = 0;
      u32 a_phase = 0, w_phase = 0, t_phase = 0;
      for (; it.valid(); it.next()) {
        const int t = it.t;
        while (tile_pre[seg + 1] <= t) ++seg;
        if (seg != cur_seg) {
          cur_seg = seg;
          wbuf = (wbuf + 1 == W_BUFS) ? 0 : wbuf + 1;
          mbar_wait(W_FULL(wbuf), (w_phase >> wbuf) & 1u);
          w_phase ^= 1u << wbuf;
        }
        mbar_wait(T_EMPTY(acc), ((t_phase >> acc) & 1u) ^ 1u);
        mbar_wait(A_FULL(a_stage), a_phase);
        tc_fence_after();
        if (lane == 0) {
          const u32 a_base = base + off_a + a_stage * a_bytes, w_base = base + off_w + wbuf * w_bytes;
          const u32 d_tmem = tmem_base + (u32)(acc * M);
          for (int kk = 0; kk < K / 16; ++kk) {
            // A: half (kk/4) of 64 K-elements, 32 bytes per 16-element step inside the 128 B swizzle row
            const u64 adesc = make_desc(a_base + (kk >> 2) * (TM * 128) + (kk & 3) * 32, 16, 1024);
            // B: 16 K-rows = 2 atoms of 8 rows x 128 B; LBO = stride between 64-column groups of N
            const u64 bdesc = make_desc(w_base + kk * 2048, (u32)(K * 128), 1024);
            tc_mma_f16(d_tmem, adesc, bdesc, idesc, kk > 0 ? 1u : 0u);
          }
          tc_commit(A_EMPTY(a_stage));   // smem A stage reusable once these MMAs retire
          tc_commit(T_FULL(acc));        // accumulator ready for the epilogue
          const int tn = it.peek_next();
          const bool last_of_seg = (tn >= total_tiles) || (tile_pre[seg + 1] <= tn);
          if (last_of_seg) tc_commit(W_EMPTY(wbuf));
        }
        __syncwarp();
        if (++a_stage == A_STAGES) { a_stage = 0; a_phase ^= 1; }
        t_phase ^= 1u << acc;
        acc ^= 1;
      }
    }
  } else if (warp >= 4) {
    // ================================ epilogue (128 threads) ================================
    const int q = warp & 3;                 // TMEM lane quarter owned by this warp
    const int r = q * 32 + lane;            // row inside the tile == TMEM lane
    const int et = threadIdx.x - 128;
    if (has_work) {
      TileIter it(G, (int)gridDim.x, total_tiles, (int)blockIdx.x);
      int seg = seg_of(it.t), acc = 0, obuf = 0;
      u32 t_phase = 0;
      for (; it.valid(); it.next()) {
        const int t = it.t;
        while (tile_pre[seg + 1] <= t) ++seg;
        const i64 row0 = P.ptr[seg] + (i64)(t - tile_pre[seg]) * TM;
        const i64 rem = P.ptr[seg + 1] - row0;
        const int valid = rem < TM ? (int)rem : TM;
        unsigned char* stage = sm + off_o + obuf * o_bytes;
        mbar_wait(T_FULL(acc), (t_phase >> acc) & 1u);
        t_phase ^= 1u << acc;
        tc_fence_after();
        if (valid == TM) {
          // the TMA store that last read this staging buffer must have finished reading it
          if (et == 0) {
            if (O_BUFS == 1) tma_store_wait_read(); else asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
          }
          asm volatile("bar.sync 1, 128;" ::: "memory");
        }
        for (int h = 0; h < NH; ++h) {   // 64 output columns per step: two TMEM loads in flight
          u32 v[64];
          const u32 taddr = tmem_base + (u32)(acc * M + h * 64) + ((u32)(q * 32) << 16);
          tc_ld_32x32(taddr, v);
          tc_ld_32x32(taddr + 32, v + 32);
          tc_wait_ld();
          u32 pk[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float f0 = __uint_as_float(v[2 * j]), f1 = __uint_as_float(v[2 * j + 1]);
            if (P.bias) {
              f0 += ld_bias<BF16>(P.bias, (i64)seg * M + h * 64 + 2 * j);
              f1 += ld_bias<BF16>(P.bias, (i64)seg * M + h * 64 + 2 * j + 1);
            }
            pk[j] = pack2<BF16>(f0, f1);
          }
          if (valid == TM) {
            // staging layout == TMA SWIZZLE_128B box [64 cols x 128 rows]: row pitch 128 B, 16 B chunk ^ (row & 7)
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int ch = j ^ (r & 7);
              *reinterpret_cast<uint4*>(stage + h * (TM * 128) + r * 128 + ch * 16) =
                  make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
            }
          } else if (r < valid) {
            uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<unsigned char*>(P.out) + ((row0 + r) * M + h * 64) * 2);
#pragma unroll
            for (int j = 0; j < 8; ++j) dst[j] = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
          }
        }
        // accumulator drained -> MMA warp may overwrite it
        tc_fence_before();
        mbar_arrive(T_EMPTY(acc));
        if (valid == TM) {
          fence_proxy_async();
          asm volatile("bar.sync 1, 128;" ::: "memory");
          if (et == 0) {
            for (int h = 0; h < NH; ++h) tma_store_2d(&map_o, base + off_o + obuf * o_bytes + h * (TM * 128), h * 64, (int)row0);
            tma_store_commit();
          }
          obuf = (obuf + 1 == O_BUFS) ? 0 : obuf + 1;
        }
        acc ^= 1;
      }
      if (et == 0) tma_store_wait_all();
    }
  }

  // ---- teardown
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
}

// ------------------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// 2-D row-major [rows, cols] tensor of 16-bit elements, box [box_rows x 64 cols], 128 B swizzle
int make_map(CUtensorMap* m, const void* ptr, i64 rows, i64 cols, int box_rows, bool bf16) {
  EncodeTiledFn enc = get_encode();
  PYGB_CHECK(enc != nullptr, PYGB200_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(ptr), dims,
                   strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with code " + std::to_string((int)r));
    return PYGB200_ERR_CUDA;
  }
  return PYGB200_OK;
}

size_t smem_needed(i64 K, i64 M, i64 B, int as, int ws, int os) {
  return 1024 + (size_t)as * TM * K * 2 + (size_t)ws * K * M * 2 + (size_t)os * TM * M * 2 + (2 * as + 8) * 8 + 16 +
         (size_t)(B + 2) * 4;
}
constexpr size_t SMEM_LIMIT = 227 * 1024;

// pipeline depths: deepest A ring that fits, preferring two W slots and two output staging buffers
struct Variant { int as, ws, os; };
// order = preference (measured on B200, profiles/matmul_variants_r1.md): two staging buffers matter most
const Variant kVariants[] = {{4, 1, 2}, {3, 1, 2}, {3, 2, 2}, {2, 2, 1}, {3, 1, 1}, {3, 2, 1}, {2, 1, 1}, {4, 2, 1}, {4, 1, 1}, {5, 1, 1}};
constexpr int N_VARIANTS = sizeof(kVariants) / sizeof(kVariants[0]);
int pick_variant(i64 K, i64 M, i64 B) {
  static int forced = -2;
  if (forced == -2) {
    const char* e = getenv("PYGB200_MM_VARIANT");
    forced = e ? atoi(e) : -1;
  }
  if (forced >= 0 && forced < N_VARIANTS && smem_needed(K, M, B, kVariants[forced].as, kVariants[forced].ws, kVariants[forced].os) <= SMEM_LIMIT)
    return forced;
  for (int i = 0; i < N_VARIANTS; ++i)
    if (smem_needed(K, M, B, kVariants[i].as, kVariants[i].ws, kVariants[i].os) <= SMEM_LIMIT) return i;
  return -1;
}

template <bool BF16, int AS, int WS, int OS>
int launch_variant(const CUtensorMap& ma, const CUtensorMap& mw, const CUtensorMap& mo, const SegParams& P, int grid,
                   size_t smem, cudaStream_t st) {
  PYGB_CUDA(cudaFuncSetAttribute(k_segment_matmul_tc<BF16, AS, WS, OS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_segment_matmul_tc<BF16, AS, WS, OS><<<grid, NTHREADS, smem, st>>>(ma, mw, mo, P);
  return PYGB200_OK;
}
template <bool BF16>
int launch_any(int v, const CUtensorMap& ma, const CUtensorMap& mw, const CUtensorMap& mo, const SegParams& P, int grid,
               size_t smem, cudaStream_t st) {
  switch (v) {
    case 0: return launch_variant<BF16, 4, 1, 2>(ma, mw, mo, P, grid, smem, st);
    case 1: return launch_variant<BF16, 3, 1, 2>(ma, mw, mo, P, grid, smem, st);
    case 2: return launch_variant<BF16, 3, 2, 2>(ma, mw, mo, P, grid, smem, st);
    case 3: return launch_variant<BF16, 2, 2, 1>(ma, mw, mo, P, grid, smem, st);
    case 4: return launch_variant<BF16, 3, 1, 1>(ma, mw, mo, P, grid, smem, st);
    case 5: return launch_variant<BF16, 3, 2, 1>(ma, mw, mo, P, grid, smem, st);
    case 6: return launch_variant<BF16, 2, 1, 1>(ma, mw, mo, P, grid, smem, st);
    case 7: return launch_variant<BF16, 4, 2, 1>(ma, mw, mo, P, grid, smem, st);
    case 8: return launch_variant<BF16, 4, 1, 1>(ma, mw, mo, P, grid, smem, st);
    default: return launch_variant<BF16, 5, 1, 1>(ma, mw, mo, P, grid, smem, st);
  }
}


// ===================================================================================== weight gradient
// dW[b] = X_b^T @ dY_b  (X [N,K], dY [N,M], dW [B,K,M]).  Replaces the B torch::matmul calls + at::stack of
// SegmentMatmul::backward (pyg_lib/csrc/ops/autograd/matmul_kernel.cpp:92-107).
//   * UMMA shape: D[K=128 x M] += A[128 x 16] * B[16 x M], contraction over the ROWS of the segment.
//     Both operands are consumed in place as MN-major tiles: a TMA box [64 cols x 128 rows] of X (or dY) is
//     exactly the canonical SW128 MN-major layout with the row index as the K dimension.
//   * Row tiles are dealt to the CTAs as contiguous ranges; a CTA accumulates in TMEM while consecutive tiles
//     stay in one segment and flushes at a segment change with fp32 atomics into a zeroed [B,K,M] scratch
//     (two accumulators, so the flush overlaps the next run's MMAs).  A final pass rounds to bf16/fp16.
//   * Tails: rows of the tile past the segment end belong to the next segment -> the MMA warp zeroes them in
//     shared memory (both operands) before issuing, then fences towards the async proxy.
constexpr int WG_STAGES = 3;

struct WgradParams {
  const i64* ptr;
  float* acc;    // [B, K, M] fp32: segments whose row tiles all belong to one CTA are stored here directly
  float* part;   // [grid, 2, K, M] fp32: a CTA's partial sums of the (at most two) segments it shares with its neighbours
  int* tile_pre_out;   // [B + 1] tile prefix for k_wgrad_finish (written by block 0)
  i64 N;
  int K, M, B;
  int* err;
};

template <bool BF16>
__global__ void __launch_bounds__(NTHREADS, 1)
k_segment_wgrad_tc(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_dy, const WgradParams P) {
  extern __shared__ unsigned char smem_raw[];
  const u32 raw = smem_u32(smem_raw);
  const u32 base = (raw + 1023u) & ~1023u;
  unsigned char* sm = smem_raw + (base - raw);
  const int K = P.K, M = P.M, KH = K / 64, NH = M / 64;
  const u32 x_bytes = TM * K * 2, y_bytes = TM * M * 2, st_bytes = x_bytes + y_bytes;
  const u32 off_bar = WG_STAGES * st_bytes;
  const u32 bar0 = base + off_bar;
  auto FULL = [&](int s) { return bar0 + 8u * (u32)s; };
  auto EMPTY = [&](int s) { return bar0 + 8u * (u32)(WG_STAGES + s); };
  auto T_FULL = [&](int s) { return bar0 + 8u * (u32)(2 * WG_STAGES + s); };
  auto T_EMPTY = [&](int s) { return bar0 + 8u * (u32)(2 * WG_STAGES + 2 + s); };
  constexpr int NBARS = 2 * WG_STAGES + 4;
  u64* bars = reinterpret_cast<u64*>(sm + off_bar);
  u32* tmem_slot = reinterpret_cast<u32*>(bars + NBARS);
  int* tile_pre = reinterpret_cast<int*>(tmem_slot + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const u32 tmem_cols = (2 * M <= 32) ? 32 : (2 * M <= 64) ? 64 : (2 * M <= 128) ? 128 : (2 * M <= 256) ? 256 : 512;
  build_tile_prefix(P.ptr, P.B, P.N, tile_pre, P.err);
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < WG_STAGES; ++s) { mbar_init(FULL(s), 1); mbar_init(EMPTY(s), 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(T_FULL(s), 1); mbar_init(T_EMPTY(s), 128); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    fence_proxy_async();
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_dy) : "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const u32 tmem_base = *tmem_slot;

  const int total_tiles = tile_pre[P.B];
  const int per_cta = (total_tiles + (int)gridDim.x - 1) / (int)gridDim.x;
  const int t_begin = (int)blockIdx.x * per_cta;
  const int t_end = min(total_tiles, t_begin + per_cta);
  if (blockIdx.x == 0) for (int b = threadIdx.x; b <= P.B; b += NTHREADS) P.tile_pre_out[b] = tile_pre[b];
  auto seg_of = [&](int t) {
    int lo = 0, hi = P.B - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (tile_pre[mid] <= t) lo = mid; else hi = mid - 1;
    }
    return lo;
  };

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (lane == 0 && t_begin < t_end) {
      int seg = seg_of(t_begin), stage = 0;
      u32 phase = 0;
      for (int t = t_begin; t < t_end; ++t) {
        while (tile_pre[seg + 1] <= t) ++seg;
        const i64 row0 = P.ptr[seg] + (i64)(t - tile_pre[seg]) * TM;
        mbar_wait(EMPTY(stage), phase ^ 1);
        mbar_expect_tx(FULL(stage), st_bytes);
        const u32 xs = base + stage * st_bytes, ys = xs + x_bytes;
        for (int h = 0; h < KH; ++h) tma_load_2d(xs + h * (TM * 128), &map_x, h * 64, (int)row0, FULL(stage));
        for (int h = 0; h < NH; ++h) tma_load_2d(ys + h * (TM * 128), &map_dy, h * 64, (int)row0, FULL(stage));
        if (++stage == WG_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    if (t_begin < t_end) {
      // D = f32, A and B both MN-major (bits 15 and 16), N = M(out cols) >> 3, M = K(=128) >> 4
      const u32 fmt = BF16 ? 1u : 0u;
      const u32 idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | (1u << 15) | (1u << 16) | ((u32)(M >> 3) << 17) | ((u32)(K >> 4) << 24);
      int seg = seg_of(t_begin), stage = 0, acc = 0;
      u32 phase = 0, t_phase = 0;
      bool run_open = false;   // an accumulation run (consecutive tiles of one segment) is in progress
      for (int t = t_begin; t < t_end; ++t) {
        while (tile_pre[seg + 1] <= t) ++seg;
        const i64 row0 = P.ptr[seg] + (i64)(t - tile_pre[seg]) * TM;
        const i64 rem = P.ptr[seg + 1] - row0;
        const int valid = rem < TM ? (int)rem : TM;
        if (!run_open) mbar_wait(T_EMPTY(acc), ((t_phase >> acc) & 1u) ^ 1u);
        mbar_wait(FULL(stage), phase);
        const u32 xs = base + stage * st_bytes, ys = xs + x_bytes;
        if (valid < TM) {
          // rows [valid, 128) belong to the next segment: zero them in both operand tiles (each row is one
          // 128-byte line per 64-column half, whatever the swizzle)
          unsigned char* xp = sm + stage * st_bytes;
          const int halves = KH + NH, chunks = (TM - valid) * 8;
          for (int h = 0; h < halves; ++h)
            for (int c = lane; c < chunks; c += 32)
              *reinterpret_cast<uint4*>(xp + h * (TM * 128) + (valid + (c >> 3)) * 128 + (c & 7) * 16) = make_uint4(0, 0, 0, 0);
          fence_proxy_async();
          __syncwarp();
        }
        tc_fence_after();
        if (lane == 0) {
          const u32 d_tmem = tmem_base + (u32)(acc * M);
          for (int kk = 0; kk < TM / 16; ++kk) {
            const u64 adesc = make_desc(xs + kk * 2048, (u32)(TM * 128), 1024);
            const u64 bdesc = make_desc(ys + kk * 2048, (u32)(TM * 128), 1024);
            tc_mma_f16(d_tmem, adesc, bdesc, idesc, (run_open || kk > 0) ? 1u : 0u);
          }
          tc_commit(EMPTY(stage));
        }
        run_open = true;
        const bool last_of_run = (t + 1 >= t_end) || (tile_pre[seg + 1] <= t + 1);
        if (last_of_run) {
          if (lane == 0) tc_commit(T_FULL(acc));
          t_phase ^= 1u << acc;
          acc ^= 1;
          run_open = false;
        }
        __syncwarp();
        if (++stage == WG_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ================================ epilogue: flush finished runs ================================
    const int q = warp & 3;
    const int r = q * 32 + lane;   // accumulator lane == column of X == row of dW[b]
    if (t_begin < t_end) {
      int seg = seg_of(t_begin), acc = 0;
      u32 t_phase = 0;
      for (int t = t_begin; t < t_end; ++t) {
        while (tile_pre[seg + 1] <= t) ++seg;
        const bool last_of_run = (t + 1 >= t_end) || (tile_pre[seg + 1] <= t + 1);
        if (!last_of_run) continue;
        mbar_wait(T_FULL(acc), (t_phase >> acc) & 1u);
        t_phase ^= 1u << acc;
        tc_fence_after();
        // Deterministic reduction (the reference's per-segment torch::matmul is deterministic,
        // ops/autograd/matmul_kernel.cpp:92-107): no atomics.  A segment whose tiles all lie inside this CTA's range
        // is stored to acc[seg]; otherwise this run is one of the CTA's at most two partial sums — slot 0 if the run
        // begins at the CTA's first tile, slot 1 if it is a later run that continues past its last tile — and
        // k_wgrad_finish adds the partials of a segment in CTA order.
        const bool exclusive = tile_pre[seg] >= t_begin && tile_pre[seg + 1] <= t_end;
        const int run_begin = max(tile_pre[seg], t_begin);
        float* dst = exclusive ? P.acc + ((i64)seg * K + r) * M
                               : P.part + (((i64)blockIdx.x * 2 + (run_begin == t_begin ? 0 : 1)) * K + r) * M;
        for (int c0 = 0; c0 < M; c0 += 32) {
          u32 v[32];
          tc_ld_32x32(tmem_base + (u32)(acc * M + c0) + ((u32)(q * 32) << 16), v);
          tc_wait_ld();
#pragma unroll
          for (int j = 0; j < 8; ++j)
            reinterpret_cast<uint4*>(dst + c0)[j] = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        }
        tc_fence_before();
        mbar_arrive(T_EMPTY(acc));
        acc ^= 1;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
}

// dW[b] in the storage type: 0 for an empty segment, acc[b] for a segment one CTA had to itself, else the sum of the
// CTAs' partial sums in CTA order (fixed order => bit-reproducible).  blockIdx.y = segment.
template <bool BF16>
__global__ void k_wgrad_finish(const float* __restrict__ acc, const float* __restrict__ part, const int* __restrict__ tile_pre,
                               void* __restrict__ dw, int KM, int grid_main) {
  const int b = blockIdx.y;
  const int t0 = tile_pre[b], t1 = tile_pre[b + 1], total = tile_pre[gridDim.y];
  const int per_cta = (total + grid_main - 1) / grid_main;
  const int c0 = t1 > t0 ? t0 / per_cta : 0, c1 = t1 > t0 ? (t1 - 1) / per_cta : -1;
  u32* out = reinterpret_cast<u32*>(dw) + (i64)b * (KM / 2);
  for (int i = (blockIdx.x * blockDim.x + threadIdx.x) * 2; i < KM; i += gridDim.x * blockDim.x * 2) {
    float2 v = make_float2(0.f, 0.f);
    if (c1 == c0) {
      v = *reinterpret_cast<const float2*>(acc + (i64)b * KM + i);
    } else {
      for (int c = c0; c <= c1; ++c) {
        const int slot = t0 <= c * per_cta ? 0 : 1;
        const float2 p = *reinterpret_cast<const float2*>(part + ((i64)c * 2 + slot) * KM + i);
        v.x += p.x; v.y += p.y;
      }
    }
    out[i >> 1] = pack2<BF16>(v.x, v.y);
  }
}

size_t wgrad_smem(i64 K, i64 M, i64 B) {
  return 1024 + (size_t)WG_STAGES * (TM * K * 2 + TM * M * 2) + (2 * WG_STAGES + 4) * 8 + 16 + (size_t)(B + 2) * 4;
}

}  // namespace

bool tcgen05_supported(i64 N, i64 K, i64 M, i64 B, int dtype, const void* x, const void* w, const void* out) {
  if (dtype != PYGB200_BF16 && dtype != PYGB200_F16) return false;
  if (K < 64 || K > 256 || K % 64 != 0) return false;
  if (M < 64 || M > 256 || M % 64 != 0) return false;
  if (B < 1 || B > MAX_SEG || N < 1 || N >= ((i64)1 << 31) || B * K >= ((i64)1 << 31)) return false;
  if (((uintptr_t)x | (uintptr_t)w | (uintptr_t)out) & 15) return false;
  if (pick_variant(K, M, B) < 0) return false;
  return true;
}

int segment_matmul_tcgen05(const void* x, const i64* ptr_dev, const void* w, const void* bias, void* out, i64 N, i64 K,
                           i64 M, i64 B, int dtype, cudaStream_t st) {
  const bool bf16 = dtype == PYGB200_BF16;
  CUtensorMap ma, mw, mo;
  if (int e = make_map(&ma, x, N, K, TM, bf16)) return e;
  if (int e = make_map(&mw, w, B * K, M, (int)K, bf16)) return e;
  if (int e = make_map(&mo, out, N, M, TM, bf16)) return e;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const i64 max_tiles = N / TM + B;
  SegParams P;
  P.ptr = ptr_dev; P.bias = bias; P.out = out; P.N = N; P.K = (int)K; P.M = (int)M; P.B = (int)B; P.err = mm_error_flag_dev();
  {
    // Scheduling chunk: ~4 consecutive tiles per chunk, chunks dealt round-robin, sized so that every CTA
    // gets (nearly) the same whole number of chunks.  Measured: 96 us vs 103-110 us for one contiguous
    // chunk per CTA at N=2^20, K=M=128 (all SMs sweep one DRAM window together).
    const i64 per_cta = (max_tiles + sms - 1) / sms;
    const i64 rounds = std::max<i64>(1, (per_cta + 2) / 4);
    P.G = (int)std::max<i64>(1, (per_cta + rounds - 1) / rounds);
    if (const char* e = getenv("PYGB200_MM_G")) P.G = atoi(e);
  }
  const int v = pick_variant(K, M, B);
  PYGB_CHECK(v >= 0, PYGB200_ERR_UNSUPPORTED, "segment_matmul_tcgen05: shape does not fit shared memory");
  size_t smem = smem_needed(K, M, B, kVariants[v].as, kVariants[v].ws, kVariants[v].os);
  int grid = (int)(max_tiles < sms ? (max_tiles < 1 ? 1 : max_tiles) : sms);
  void* tk = prof_begin(st);
  if (int e = bf16 ? launch_any<true>(v, ma, mw, mo, P, grid, smem, st) : launch_any<false>(v, ma, mw, mo, P, grid, smem, st)) return e;
  prof_end(tk, "segment_matmul", st, N);
  PYGB_LAUNCH_CHECK();
  return PYGB200_OK;
}

// ===================================================================================== TF32 forward
// fp32 storage, TF32 tensor-core math (tcgen05.mma kind::tf32) — taken when the caller allows TF32
// (torch.get_float32_matmul_precision() != 'highest', like the reference's TensorOp kernel,
// pyg_lib/csrc/ops/cuda/matmul_kernel.cu:159-190).  Same structure as the 16-bit kernel with 4-byte elements:
// a 128-byte swizzle row holds 32 elements, one MMA covers K = 8.  A streams in [128 rows x 64 cols] stages
// (two TMA boxes each); the fp32 result leaves in 32-column groups through a double-buffered 16 KB staging
// tile.  Unlike the 16-bit kernel W is NOT consumed in place: MN-major 32-bit operands need the special
// 128B_BASE32B swizzle (the first version with plain SWIZZLE_128B produced zeros), so W is transposed once per
// call into a [B, M, K] scratch (4 MB at C3, ~3 us) and enters as an ordinary K-major operand (K*M*4 <= 64 KB
// resident per segment).
namespace {

constexpr int TF_A_STAGES = 4;
constexpr u32 TF_A_BYTES = TM * 64 * 4;    // one stage
constexpr u32 TF_O_BYTES = TM * 32 * 4;    // one staging tile

__device__ __forceinline__ void tc_mma_tf32(u32 tmem_d, u64 adesc, u64 bdesc, u32 idesc, u32 accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

struct Tf32Params {
  const i64* ptr;
  const float* bias;
  float* out;
  i64 N;
  int K, M, B, G;
  int* err;
};

__global__ void __launch_bounds__(NTHREADS, 1)
k_segment_matmul_tf32(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w,
                      const __grid_constant__ CUtensorMap map_o, const Tf32Params P) {
  extern __shared__ unsigned char smem_raw[];
  const u32 raw = smem_u32(smem_raw);
  const u32 base = (raw + 1023u) & ~1023u;
  unsigned char* sm = smem_raw + (base - raw);
  const int K = P.K, M = P.M, KC = K / 64, NG = M / 32;
  const u32 w_bytes = (u32)K * M * 4;
  const u32 off_w = TF_A_STAGES * TF_A_BYTES, off_o = off_w + w_bytes, off_bar = off_o + 2 * TF_O_BYTES;
  const u32 bar0 = base + off_bar;
  auto A_FULL = [&](int s) { return bar0 + 8u * (u32)s; };
  auto A_EMPTY = [&](int s) { return bar0 + 8u * (u32)(TF_A_STAGES + s); };
  const u32 W_FULL = bar0 + 8u * (2 * TF_A_STAGES), W_EMPTY = W_FULL + 8;
  auto T_FULL = [&](int s) { return bar0 + 8u * (u32)(2 * TF_A_STAGES + 2 + s); };
  auto T_EMPTY = [&](int s) { return bar0 + 8u * (u32)(2 * TF_A_STAGES + 4 + s); };
  constexpr int NBARS = 2 * TF_A_STAGES + 6;
  u64* bars = reinterpret_cast<u64*>(sm + off_bar);
  u32* tmem_slot = reinterpret_cast<u32*>(bars + NBARS);
  int* tile_pre = reinterpret_cast<int*>(tmem_slot + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const u32 tmem_cols = (2 * M <= 32) ? 32 : (2 * M <= 64) ? 64 : (2 * M <= 128) ? 128 : (2 * M <= 256) ? 256 : 512;
  build_tile_prefix(P.ptr, P.B, P.N, tile_pre, P.err);
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < TF_A_STAGES; ++s) { mbar_init(A_FULL(s), 1); mbar_init(A_EMPTY(s), 1); }
    mbar_init(W_FULL, 1); mbar_init(W_EMPTY, 1);
    for (int s = 0; s < 2; ++s) { mbar_init(T_FULL(s), 1); mbar_init(T_EMPTY(s), 128); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    fence_proxy_async();
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_o) : "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const u32 tmem_base = *tmem_slot;
  const int total_tiles = tile_pre[P.B];
  const int per_cta = (total_tiles + (int)gridDim.x - 1) / (int)gridDim.x;
  const int G = P.G > 0 ? P.G : max(per_cta, 1);
  const bool has_work = (long long)blockIdx.x * G < total_tiles;
  auto seg_of = [&](int t) {
    int lo = 0, hi = P.B - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (tile_pre[mid] <= t) lo = mid; else hi = mid - 1;
    }
    return lo;
  };

  if (warp == 0) {
    if (lane == 0 && has_work) {
      TileIter it(G, (int)gridDim.x, total_tiles, (int)blockIdx.x);
      int seg = seg_of(it.t), cur_seg = -1, stage = 0;
      u32 a_phase = 0, w_phase = 0;
      for (; it.valid(); it.next()) {
        const int t = it.t;
        while (tile_pre[seg + 1] <= t) ++seg;
        if (seg != cur_seg) {
          cur_seg = seg;
          mbar_wait(W_EMPTY, w_phase ^ 1);
          w_phase ^= 1;
          mbar_expect_tx(W_FULL, w_bytes);
          for (int bx = 0; bx < K / 32; ++bx)   // W^T[b]: box [32 cols of K x M rows] -> [M][128 B], K-major SW128 atoms
            tma_load_2d(base + off_w + bx * (M * 128), &map_w, bx * 32, seg * M, W_FULL);
        }
        const i64 row0 = P.ptr[seg] + (i64)(t - tile_pre[seg]) * TM;
        for (int c = 0; c < KC; ++c) {
          mbar_wait(A_EMPTY(stage), a_phase ^ 1);
          mbar_expect_tx(A_FULL(stage), TF_A_BYTES);
          for (int h = 0; h < 2; ++h)     // box [32 cols of K x 128 rows] -> [128][128 B]: K-major SW128 atoms
            tma_load_2d(base + stage * TF_A_BYTES + h * (TM * 128), &map_a, c * 64 + h * 32, (int)row0, A_FULL(stage));
          if (++stage == TF_A_STAGES) { stage = 0; a_phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (has_work) {
      // D = f32 (bit 4), A/B = TF32 (format 2 at bits 7 and 10), both K-major, N >> 3, M >> 4
      const u32 idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((u32)(M >> 3) << 17) | ((u32)(TM >> 4) << 24);
      TileIter it(G, (int)gridDim.x, total_tiles, (int)blockIdx.x);
      int seg = seg_of(it.t), cur_seg = -1, stage = 0, acc = 0;
      u32 a_phase = 0, w_phase = 0, t_phase = 0;
      for (; it.valid(); it.next()) {
        const int t = it.t;
        while (tile_pre[seg + 1] <= t) ++seg;
        if (seg != cur_seg) {
          cur_seg = seg;
          mbar_wait(W_FULL, w_phase);
          w_phase ^= 1;
        }
        mbar_wait(T_EMPTY(acc), ((t_phase >> acc) & 1u) ^ 1u);
        const u32 d_tmem = tmem_base + (u32)(acc * M);
        for (int c = 0; c < KC; ++c) {
          mbar_wait(A_FULL(stage), a_phase);
          tc_fence_after();
          if (lane == 0) {
            const u32 a_base = base + stage * TF_A_BYTES, w_base = base + off_w;
            for (int kk = 0; kk < 8; ++kk) {   // K = 8 per instruction: 32 bytes inside the 128 B swizzle row
              const u64 adesc = make_desc(a_base + (kk >> 2) * (TM * 128) + (kk & 3) * 32, 16, 1024);
              const int kg = c * 8 + kk;   // K step of 8 inside the whole K: box kg/4, 32 bytes per step inside the row
              const u64 bdesc = make_desc(w_base + (kg >> 2) * (M * 128) + (kg & 3) * 32, 16, 1024);
              tc_mma_tf32(d_tmem, adesc, bdesc, idesc, (c > 0 || kk > 0) ? 1u : 0u);
            }
            tc_commit(A_EMPTY(stage));
          }
          __syncwarp();
          if (++stage == TF_A_STAGES) { stage = 0; a_phase ^= 1; }
        }
        if (lane == 0) {
          tc_commit(T_FULL(acc));
          const int tn = it.peek_next();
          if ((tn >= total_tiles) || (tile_pre[seg + 1] <= tn)) tc_commit(W_EMPTY);
        }
        __syncwarp();
        t_phase ^= 1u << acc;
        acc ^= 1;
      }
    }
  } else if (warp >= 4) {
    const int q = warp & 3, r = q * 32 + lane, et = threadIdx.x - 128;
    if (has_work) {
      TileIter it(G, (int)gridDim.x, total_tiles, (int)blockIdx.x);
      int seg = seg_of(it.t), acc = 0, obuf = 0;
      u32 t_phase = 0;
      for (; it.valid(); it.next()) {
        const int t = it.t;
        while (tile_pre[seg + 1] <= t) ++seg;
        const i64 row0 = P.ptr[seg] + (i64)(t - tile_pre[seg]) * TM;
        const i64 rem = P.ptr[seg + 1] - row0;
        const int valid = rem < TM ? (int)rem : TM;
        mbar_wait(T_FULL(acc), (t_phase >> acc) & 1u);
        t_phase ^= 1u << acc;
        tc_fence_after();
        for (int g = 0; g < NG; ++g) {
          u32 v[32];
          tc_ld_32x32(tmem_base + (u32)(acc * M + g * 32) + ((u32)(q * 32) << 16), v);
          tc_wait_ld();
          if (P.bias) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + P.bias[(i64)seg * M + g * 32 + j]);
          }
          if (valid == TM) {
            unsigned char* stage_o = sm + off_o + obuf * TF_O_BYTES;
            if (et == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");   // the store that last read this buffer
            asm volatile("bar.sync 1, 128;" ::: "memory");
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int ch = j ^ (r & 7);
              *reinterpret_cast<uint4*>(stage_o + r * 128 + ch * 16) = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            }
            fence_proxy_async();
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (et == 0) {
              tma_store_2d(&map_o, base + off_o + obuf * TF_O_BYTES, g * 32, (int)row0);
              tma_store_commit();
            }
            obuf ^= 1;
          } else if (r < valid) {
            uint4* dst = reinterpret_cast<uint4*>(P.out + (row0 + r) * M + g * 32);
#pragma unroll
            for (int j = 0; j < 8; ++j) dst[j] = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          }
        }
        tc_fence_before();
        mbar_arrive(T_EMPTY(acc));
        acc ^= 1;
      }
      if (et == 0) tma_store_wait_all();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
}

// W[b] [K, M] -> Wt[b] [M, K]
__global__ void k_transpose_w_f32(const float* __restrict__ w, float* __restrict__ wt, int K, int M) {
  __shared__ float tile[32][33];
  const float* src = w + (size_t)blockIdx.z * K * M;
  float* dst = wt + (size_t)blockIdx.z * K * M;
  const int k0 = blockIdx.y * 32, m0 = blockIdx.x * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int k = k0 + i, m = m0 + threadIdx.x;
    tile[i][threadIdx.x] = (k < K && m < M) ? src[(size_t)k * M + m] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int m = m0 + i, k = k0 + threadIdx.x;
    if (m < M && k < K) dst[(size_t)m * K + k] = tile[threadIdx.x][i];
  }
}

// 2-D row-major fp32 tensor, box [box_rows x 32 cols] (128 bytes), 128 B swizzle
int make_map_f32(CUtensorMap* m, const void* ptr, i64 rows, i64 cols, int box_rows) {
  EncodeTiledFn enc = get_encode();
  PYGB_CHECK(enc != nullptr, PYGB200_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * 4};
  cuuint32_t box[2] = {32, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(f32) failed with code " + std::to_string((int)r));
    return PYGB200_ERR_CUDA;
  }
  return PYGB200_OK;
}

size_t tf32_smem(i64 K, i64 M, i64 B) {
  return 1024 + (size_t)TF_A_STAGES * TF_A_BYTES + (size_t)K * M * 4 + 2 * TF_O_BYTES + (2 * TF_A_STAGES + 6) * 8 + 16 +
         (size_t)(B + 2) * 4;
}

}  // namespace

bool tf32_supported(i64 N, i64 K, i64 M, i64 B, const void* x, const void* w, const void* out) {
  if (K < 64 || K > 256 || K % 64 != 0) return false;
  if (M < 32 || M > 256 || M % 32 != 0) return false;
  if (B < 1 || B > MAX_SEG || N < 1 || N >= ((i64)1 << 31) || B * K >= ((i64)1 << 31)) return false;
  if (((uintptr_t)x | (uintptr_t)w | (uintptr_t)out) & 15) return false;
  return tf32_smem(K, M, B) <= SMEM_LIMIT;
}

int segment_matmul_tf32(const void* x, const i64* ptr_dev, const void* w, const void* bias, void* out, i64 N, i64 K, i64 M,
                        i64 B, cudaStream_t st) {
  AsyncScratch sc;   // (freed on every return path)
  if (int e = sc.alloc((size_t)B * K * M * 4, st)) return e;
  float* wt = (float*)sc.p;
  k_transpose_w_f32<<<dim3((unsigned)((M + 31) / 32), (unsigned)((K + 31) / 32), (unsigned)B), dim3(32, 8), 0, st>>>(
      (const float*)w, wt, (int)K, (int)M);
  PYGB_LAUNCH_CHECK();
  CUtensorMap ma, mw, mo;
  if (int e = make_map_f32(&ma, x, N, K, TM)) return e;
  if (int e = make_map_f32(&mw, wt, B * M, K, (int)M)) return e;
  if (int e = make_map_f32(&mo, out, N, M, TM)) return e;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const i64 max_tiles = N / TM + B;
  Tf32Params P;
  P.ptr = ptr_dev; P.bias = (const float*)bias; P.out = (float*)out; P.N = N; P.K = (int)K; P.M = (int)M; P.B = (int)B; P.err = mm_error_flag_dev();
  {
    const i64 per_cta = (max_tiles + sms - 1) / sms;
    const i64 rounds = std::max<i64>(1, (per_cta + 2) / 4);
    P.G = (int)std::max<i64>(1, (per_cta + rounds - 1) / rounds);
  }
  const size_t smem = tf32_smem(K, M, B);
  const int grid = (int)(max_tiles < sms ? (max_tiles < 1 ? 1 : max_tiles) : sms);
  PYGB_CUDA(cudaFuncSetAttribute(k_segment_matmul_tf32, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  void* tk = prof_begin(st);
  k_segment_matmul_tf32<<<grid, NTHREADS, smem, st>>>(ma, mw, mo, P);
  prof_end(tk, "segment_matmul", st, N);
  PYGB_LAUNCH_CHECK();
  return PYGB200_OK;
}

bool wgrad_tcgen05_supported(i64 N, i64 K, i64 M, i64 B, int dtype, const void* x, const void* dy, const void* dw) {
  if (dtype != PYGB200_BF16 && dtype != PYGB200_F16) return false;
  if (K != 128) return false;                                  // UMMA M dimension (rows of dW[b])
  if (M < 64 || M > 256 || M % 64 != 0) return false;
  if (B < 1 || B > MAX_SEG || N < 1 || N >= ((i64)1 << 31)) return false;
  if (((uintptr_t)x | (uintptr_t)dy | (uintptr_t)dw) & 15) return false;
  if (wgrad_smem(K, M, B) > SMEM_LIMIT) return false;
  return true;
}

int segment_wgrad_tcgen05(const void* x, const i64* ptr_dev, const void* dy, void* dw, i64 N, i64 K, i64 M, i64 B, int dtype,
                          cudaStream_t st) {
  const bool bf16 = dtype == PYGB200_BF16;
  CUtensorMap mx, my;
  if (int e = make_map(&mx, x, N, K, TM, bf16)) return e;
  if (int e = make_map(&my, dy, N, M, TM, bf16)) return e;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const i64 max_tiles = N / TM + B;
  const int grid = (int)(max_tiles < sms ? (max_tiles < 1 ? 1 : max_tiles) : sms);
  // scratch: acc [B,K,M] | part [grid,2,K,M] | tile prefix [B+1]   (nothing needs zeroing: every word that is read was stored)
  const size_t acc_bytes = (size_t)B * K * M * sizeof(float), part_bytes = (size_t)grid * 2 * K * M * sizeof(float);
  AsyncScratch sc;   // (freed on every return path)
  if (int e = sc.alloc(acc_bytes + part_bytes + ((size_t)B + 1) * sizeof(int), st)) return e;
  float* acc = (float*)sc.p;
  WgradParams P;
  P.ptr = ptr_dev; P.acc = acc; P.part = acc + (size_t)B * K * M; P.tile_pre_out = reinterpret_cast<int*>(P.part + (size_t)grid * 2 * K * M);
  P.N = N; P.K = (int)K; P.M = (int)M; P.B = (int)B; P.err = mm_error_flag_dev();
  const size_t smem = wgrad_smem(K, M, B);
  void* tk = prof_begin(st);
  if (bf16) {
    PYGB_CUDA(cudaFuncSetAttribute(k_segment_wgrad_tc<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_segment_wgrad_tc<true><<<grid, NTHREADS, smem, st>>>(mx, my, P);
  } else {
    PYGB_CUDA(cudaFuncSetAttribute(k_segment_wgrad_tc<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_segment_wgrad_tc<false><<<grid, NTHREADS, smem, st>>>(mx, my, P);
  }
  prof_end(tk, "segment_matmul", st, N);
  PYGB_LAUNCH_CHECK();
  const int KM = (int)(K * M);
  const dim3 fgrid((unsigned)std::min<i64>((KM / 2 + 255) / 256, 64), (unsigned)B);
  if (bf16) k_wgrad_finish<true><<<fgrid, 256, 0, st>>>(acc, P.part, P.tile_pre_out, dw, KM, grid);
  else k_wgrad_finish<false><<<fgrid, 256, 0, st>>>(acc, P.part, P.tile_pre_out, dw, KM, grid);
  PYGB_LAUNCH_CHECK();
  return PYGB200_OK;
}

}  // namespace pygb200
