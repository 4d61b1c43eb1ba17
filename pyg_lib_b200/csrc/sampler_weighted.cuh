// Biased (edge_weight) neighbour sampling — the reference's _biased_sample (neighbor_kernel.cpp:245-285), bit for bit.
// Included by sampler.cu inside its anonymous namespace (uses PassArgs, NodeRec, last_block, mt_temper ...).
//
// What the reference does per frontier node with `count` = the hop's fan-out and `population` = its degree:
//   * count < 0, or no replacement and count >= population: the whole neighbourhood, in row order, no random numbers;
//   * replacement:  at::multinomial(weight, count, true) — a float32 running sum of the row's weights divided by the
//     total, the last bucket forced to 1, and per sample one double u = (random64() & (2^53-1)) * 2^-53 located by a
//     lower-bound binary search (ATen/native/cpu/MultinomialKernel.cpp).  Two engine outputs per sample, high word first;
//   * no replacement:  key = rand.log() / weight with rand = at::empty_like(weight).uniform_(), then key.topk(count):
//     one engine output per NEIGHBOUR, u = (word & 0xFFFFFF) * 2^-24, torch's CPU float log (MKL vmsLn, reproduced as
//     "correctly rounded log + table of deviations", mkl_logf_table.inc), an IEEE float32 division, and the `count`
//     largest keys in descending order.  at::topk on CPU is std::partial_sort / std::nth_element + std::sort of
//     (value, index) pairs compared BY VALUE ONLY (NaN first): whenever two of the count + 1 largest keys are equal —
//     masked neighbours (weight 0 -> key -inf) are the common case — which of them is taken, and in which order, is
//     decided by libstdc++'s algorithms.  The warp finds the count + 1 largest keys in parallel; if it sees such a tie,
//     one lane replays libstdc++ (introselect, introsort, heap select — restated below) on the row.
// The random words of a pass are contiguous in the engine's output stream: a frontier node's words start where the
// previous node's end, so one scan over the degrees (k_w_count) places every node; the host reads the totals (one
// synchronisation per biased pass), extends the raw mt19937 stream and sizes the key scratch, then k_w_sample draws.
// The RandintEngine the reference constructs is never used on this path, but its constructor has consumed 256 outputs.
#pragma once

struct WArgs {
  const float* weight;         // the relation's edge weights
  float* skey; u32* sidx;      // scratch: keys / cumulative sums (and, for the libstdc++ replay, indices) of the rows that draw
  const u32* lbits;            // 2^24-bit map of the k where MKL's logf(k * 2^-24) is not the correctly rounded value
  const u32* ltab; int ltab_n; // ... and the sorted list (k << 1 | up) that says in which direction
  i64 wbase;                   // raw-stream index of the pass's first word
};
enum { W_ERR_NEG = 3, W_ERR_INF = 4, W_ERR_SUM = 5, W_ERR_CATEGORIES = 6 };

__global__ void k_w_log_bitmap(const u32* __restrict__ tab, int n, u32* __restrict__ bits) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { const u32 k = tab[i] >> 1; atomicOr(&bits[k >> 5], 1u << (k & 31)); }
}

// torch.log(float(k24) * 2^-24) on CPU
__device__ __forceinline__ float w_mkl_logf(u32 k24, const WArgs& w) {
  if (k24 == 0) return __int_as_float(0xff800000);               // log(0) = -inf
  const float u = (float)k24 * 5.9604644775390625e-08f;          // exact
  float L = __double2float_rn(log((double)u));                   // correctly rounded (margin: tools/make_logf_table.py)
  if ((__ldg(&w.lbits[k24 >> 5]) >> (k24 & 31)) & 1u) {
    int lo = 0, hi = w.ltab_n - 1;
    while (lo <= hi) {
      const int mid = (lo + hi) >> 1;
      const u32 v = __ldg(&w.ltab[mid]);
      if ((v >> 1) == k24) { L = __int_as_float(__float_as_int(L) + ((v & 1u) ? 1 : -1)); break; }
      if ((v >> 1) < k24) lo = mid + 1; else hi = mid - 1;
    }
  }
  return L;
}

// ---------------------------------------------------------------------------------- count
struct W3 { u64 e, w, s; };   // emitted edges, engine words, scratch entries
__device__ __forceinline__ W3 w3_add(const W3& a, const W3& b) { return W3{a.e + b.e, a.w + b.w, a.s + b.s}; }
__device__ __forceinline__ W3 w3_shfl_up(const W3& v, int d) {
  return W3{__shfl_up_sync(0xffffffffu, v.e, d), __shfl_up_sync(0xffffffffu, v.w, d), __shfl_up_sync(0xffffffffu, v.s, d)};
}
// ordered exclusive scan over the NT threads of a block
__device__ __forceinline__ W3 w3_block_scan(W3 v, W3* total) {
  __shared__ W3 s_w[NT / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  W3 inc = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const W3 o = w3_shfl_up(inc, d);
    if (lane >= d) inc = w3_add(inc, o);
  }
  __syncthreads();   // (s_w of a previous call has been read)
  if (lane == 31) s_w[warp] = inc;
  __syncthreads();
  W3 pre{0, 0, 0}, tot{0, 0, 0};
#pragma unroll
  for (int i = 0; i < NT / 32; ++i) {
    if (i < warp) pre = w3_add(pre, s_w[i]);
    tot = w3_add(tot, s_w[i]);
  }
  *total = tot;
  return W3{pre.e + inc.e - v.e, pre.w + inc.w - v.w, pre.s + inc.s - v.s};
}

// how a frontier node is sampled on the biased path (neighbor_kernel.cpp:46-49,257-278)
__device__ __forceinline__ W3 w_classify(i64 deg, i64 k, int replace) {
  if (deg == 0 || k == 0) return W3{0, 0, 0};
  if (k < 0 || (!replace && k >= deg)) return W3{(u64)deg, 0, 0};
  if (replace) return W3{(u64)k, 2 * (u64)k, (u64)deg};
  return W3{(u64)k, (u64)deg, (u64)deg};
}

// NodeRec on this path: loc_off = edges emitted by earlier nodes of the tile, pf[0..1] = engine words consumed by them,
// pf[2..3] = their scratch entries; per tile: tile_off = first edge, tile_pos = first word (relative to the pass),
// ((i64*)tile_func)[2 * tile + 1] = first scratch entry.
template <typename idx_t>
__global__ void __launch_bounds__(NT) k_w_count(const PassArgs a) {
  pdl_enter();
  deferred_lookup(a);   // (wide-table schedule: local ids of the previous pass; must precede the ticket, like in k_count)
  const i64 begin = a.st[a.o_src_begin], end = a.st[a.o_src_end];
  const i64 F = end - begin;
  const i64 ntiles = ceil_div(F, NT);
  i64* tile_scr = reinterpret_cast<i64*>(a.tile_func);
  const idx_t* __restrict__ rowptr = (const idx_t*)a.rowptr;
  for (i64 tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const i64 i = tile * NT + threadIdx.x;
    i64 rs = 0, deg = 0;
    W3 v{0, 0, 0};
    if (i < F) {
      const i64 node = a.src_nodes[begin + i];
      rs = (i64)rowptr[node];
      deg = (i64)rowptr[node + 1] - rs;
      v = w_classify(deg, a.fanout, a.replace);
      if (a.replace && v.w && deg > (1 << 24)) a.st[ST_ERROR] = W_ERR_CATEGORIES;   // at::multinomial: "number of categories cannot exceed 2^24"
    }
    W3 tot;
    const W3 ex = w3_block_scan(v, &tot);
    if (i < F) {
      NodeRec r;
      r.rs = rs; r.deg = (u32)deg; r.loc_off = (u32)ex.e;
      r.pf[0] = (u32)ex.w; r.pf[1] = (u32)(ex.w >> 32); r.pf[2] = (u32)ex.s; r.pf[3] = (u32)(ex.s >> 32);
      a.rec[i] = r;
    }
    if (threadIdx.x == 0) { a.tile_out[tile] = (i64)tot.e; a.tile_pos[tile] = (i64)tot.w; tile_scr[2 * tile] = (i64)tot.s; }
  }
  if (last_block(&a.st[ST_TICKET_A])) {
    W3 carry{0, 0, 0};
    for (i64 base = 0; base < ntiles; base += NT) {
      const i64 t = base + threadIdx.x;
      W3 v{0, 0, 0};
      if (t < ntiles) v = W3{(u64)__ldcg(&a.tile_out[t]), (u64)__ldcg(&a.tile_pos[t]), (u64)__ldcg(&tile_scr[2 * t])};
      W3 tot;
      const W3 ex = w3_block_scan(v, &tot);
      if (t < ntiles) {
        a.tile_off[t] = (i64)(carry.e + ex.e); a.tile_pos[t] = (i64)(carry.w + ex.w); tile_scr[2 * t + 1] = (i64)(carry.s + ex.s);
      }
      carry = w3_add(carry, tot);
    }
    if (threadIdx.x == 0) {
      a.st[ST_PASS_F] = F;
      a.st[ST_PASS_E] = (i64)carry.e;
      a.st[ST_W_WORDS] = (i64)carry.w;
      a.st[ST_W_SCR] = (i64)carry.s;
      a.st[ST_PASS_BASE] = a.st[a.o_rel_edges];
      a.st[a.o_rel_edges] += (i64)carry.e;
      a.st[a.o_eph] = (i64)carry.e;
    }
  }
}

#include "topk_replay.h"

// ---------------------------------------------------------------------------------- draws: one warp per frontier node
// total order of the parallel selection: NaN first, then value descending, then index ascending, packed so that the
// LARGEST 64-bit code is the next pick
__device__ __forceinline__ u64 w_code(float x, u32 idx) {
  const u32 b = __float_as_uint(x);
  const u32 o = (x != x) ? 0xfffffffeu : ((b & 0x80000000u) ? ~b : (b | 0x80000000u));   // (+inf is 0xff800000; ~0 stays free: "before the first pick")
  return ((u64)o << 32) | (u64)(0xffffffffu - idx);
}

// comp-equivalent keys (neither goes before the other), from the value part of two codes: equal, or +0 / -0
__device__ __forceinline__ bool w_same_key(u32 o1, u32 o2) {
  return o1 == o2 || (o1 == 0x80000000u && o2 == 0x7fffffffu) || (o1 == 0x7fffffffu && o2 == 0x80000000u);
}
constexpr int W_LIST_MAX = 32;   // fan-out + 1 candidates per lane live in shared memory; beyond that the keys are re-scanned per pick

template <typename idx_t>
__global__ void __launch_bounds__(NT) k_w_sample(const PassArgs a, const WArgs w) {
  extern __shared__ u64 w_lists[];   // [warp][M][32 lanes]: every lane's M best codes so far
  pdl_enter();
  const i64 F = a.st[ST_PASS_F], begin = a.st[a.o_src_begin], pbase = a.st[ST_PASS_BASE];
  const int lane = threadIdx.x & 31;
  const i64 nwarps = (i64)gridDim.x * (NT / 32);
  const idx_t* __restrict__ col = (const idx_t*)a.col;
  const i64* tile_scr = reinterpret_cast<const i64*>(a.tile_func);
  const i64 k = a.fanout;
  const int M = (!a.replace && k >= 0 && k + 1 <= W_LIST_MAX) ? (int)k + 1 : 0;
  u64* L = w_lists + (size_t)(threadIdx.x >> 5) * M * 32 + lane;   // entry m at L[m * 32]
  for (i64 i = (i64)blockIdx.x * (NT / 32) + (threadIdx.x >> 5); i < F; i += nwarps) {
    const NodeRec r = a.rec[i];
    const i64 deg = r.deg;
    if (deg == 0 || k == 0) continue;
    const i64 tile = i / NT;
    const i64 off = pbase + a.tile_off[tile] + r.loc_off;   // first output slot of the node
    const i64 src_pos = begin + i;                          // local id of the source node (neighbor_kernel.cpp:438)
    auto emit = [&](i64 j, i64 e) { a.row[off + j] = src_pos; a.eid[off + j] = e; a.colv[off + j] = (i64)col[e]; };
    if (k < 0 || (!a.replace && k >= deg)) {   // the whole neighbourhood
      for (i64 j = lane; j < deg; j += 32) emit(j, r.rs + j);
      continue;
    }
    const i64 wpos = w.wbase + a.tile_pos[tile] + (i64)((u64)r.pf[0] | ((u64)r.pf[1] << 32));
    const i64 spos = tile_scr[2 * tile + 1] + (i64)((u64)r.pf[2] | ((u64)r.pf[3] << 32));
    float* key = w.skey + spos;
    const float* __restrict__ wt = w.weight + r.rs;
    const int n = (int)deg;
    if (a.replace) {
      // ---- at::multinomial with replacement: a float32 running sum IN ROW ORDER (that is its definition), normalised.
      // 32 weights per coalesced load; every lane replays the 32 additions from shuffles (same chain, same roundings)
      // and keeps the prefix that ends at its element.
      int err = 0; float sum = 0.f;
      for (int base = 0; base < n; base += 32) {
        const int j = base + lane;
        const float v = j < n ? wt[j] : 0.f;
        const int bad = j < n ? (!(v >= 0.f) ? W_ERR_NEG : (isinf(v) ? W_ERR_INF : 0)) : 0;   // "entry < 0" (NaN fails it too), "inf or NaN"
        const unsigned bm = __ballot_sync(0xffffffffu, bad != 0);
        if (bm && !err) err = __shfl_sync(0xffffffffu, bad, __ffs(bm) - 1);   // the first offending entry decides, like the reference's loop
        const int m = n - base < 32 ? n - base : 32;
        float mine = 0.f;
        for (int q = 0; q < m; ++q) {
          sum = __fadd_rn(sum, __shfl_sync(0xffffffffu, v, q));
          if (lane == q) mine = sum;
        }
        if (j < n) key[j] = mine;
      }
      if (!err && !(sum > 0.f)) err = W_ERR_SUM;
      if (err) {   // reported by the host when the run ends; the node's slots still get valid edges so that the rest of the pass stays in bounds
        if (lane == 0) a.st[ST_ERROR] = err;
        for (i64 t = lane; t < k; t += 32) emit(t, r.rs);
        continue;
      }
      __syncwarp();
      for (int j = lane; j < n; j += 32) key[j] = __fdiv_rn(key[j], sum);
      __syncwarp();
      if (lane == 0) key[n - 1] = 1.0f;
      __syncwarp();
      for (i64 t = lane; t < k; t += 32) {
        const u64 hi = mt_temper(a.raw[wpos + 2 * t]), lo = mt_temper(a.raw[wpos + 2 * t + 1]);
        const double u = (double)(((hi << 32) | lo) & ((1ull << 53) - 1)) * 1.1102230246251565e-16;   // * 2^-53
        int l = 0, h = n;
        while (h - l > 0) {
          const int mid = l + (h - l) / 2;
          if ((double)key[mid] < u) l = mid + 1; else h = mid;
        }
        emit(t, r.rs + l);
      }
      continue;
    }
    // ---- weighted reservoir keys (kept in `key` for the libstdc++ replay), the k largest in descending order.
    // One pass: four independent elements per lane in flight (engine word, weight, log-table bit), every lane keeps its own
    // k + 1 best codes in shared memory — the k + 1 best of the row are among them.
    int cnt = 0, mnpos = 0;
    u64 mn = ~0ull;
    for (int jb = lane; jb < n; jb += 128) {
      u32 wd[4]; float ww[4], kv[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int j = jb + 32 * q;
        wd[q] = j < n ? a.raw[wpos + j] : 0u;
        ww[q] = j < n ? wt[j] : 1.f;
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) kv[q] = __fdiv_rn(w_mkl_logf(mt_temper(wd[q]) & 0xffffffu, w), ww[q]);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int j = jb + 32 * q;
        if (j < n) {
          key[j] = kv[q];
          if (M) {
            const u64 c = w_code(kv[q], (u32)j);
            if (cnt < M) {
              L[cnt * 32] = c;
              if (c < mn) { mn = c; mnpos = cnt; }
              ++cnt;
            } else if (c > mn) {
              L[mnpos * 32] = c;
              mn = ~0ull;
              for (int m = 0; m < M; ++m) { const u64 v = L[m * 32]; if (v < mn) { mn = v; mnpos = m; } }
            }
          }
        }
      }
    }
    __syncwarp();
    u64 prev = ~0ull;
    u32 prev_o = 0;
    bool tie = false;
    const int rounds = (int)k + 1;   // (n > k here)
    for (int t = 0; t < rounds; ++t) {
      u64 best = 0;   // (every real code is > 0: the index part alone is >= 2^32 - n)
      if (M) {
        for (int m = 0; m < cnt; ++m) { const u64 c = L[m * 32]; if (c < prev && c > best) best = c; }
      } else {
        for (int j = lane; j < n; j += 32) { const u64 c = w_code(key[j], (u32)j); if (c < prev && c > best) best = c; }
      }
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) { const u64 o = __shfl_xor_sync(0xffffffffu, best, d); best = o > best ? o : best; }
      const u32 idx = 0xffffffffu - (u32)best, o = (u32)(best >> 32);
      if (t > 0 && w_same_key(o, prev_o)) tie = true;
      if (t < k && lane == (t & 31)) emit(t, r.rs + idx);
      prev = best; prev_o = o;
    }
    if (tie) {   // two of the k + 1 largest keys are equal: libstdc++ decides (one lane; rare outside masked rows)
      u32* sidx = w.sidx + spos;
      for (int j = lane; j < n; j += 32) sidx[j] = (u32)j;
      __syncwarp();
      if (lane == 0) w_topk_replay(WPairs{key, sidx}, n, (int)k);
      __syncwarp();
      for (i64 t = lane; t < k; t += 32) emit(t, r.rs + sidx[t]);
    }
    __syncwarp();   // (the lists are reused by this warp's next node)
  }
}
