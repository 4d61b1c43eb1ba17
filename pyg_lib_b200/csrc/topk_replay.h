// at::topk's CPU tie behaviour, replayed: see sampler_weighted.cuh.  Plain C++ (also compiled for the host by
// tests/test_weighted_oracle.py, which checks it against torch.topk on tie-heavy inputs); device code includes it
// inside pygb200's anonymous namespace.
#pragma once
#include <stdint.h>
#ifdef __CUDACC__
#define PYGB_HD __host__ __device__
#else
#define PYGB_HD
#endif

// ---------------------------------------------------------------------------------- libstdc++ replay (one lane)
// std::partial_sort / std::nth_element / std::sort (bits/stl_algo.h, bits/stl_heap.h of GCC's libstdc++) on the pairs
// (K[j], I[j]) with the comparator at::topk uses for `largest` (ATen/native/cpu/TopKImpl.h): x before y iff x is NaN and
// y is not, or x > y.  Restated from the algorithms' published structure; checked against torch.topk on tie-heavy inputs
// through the identical numpy restatement in tests/weighted_model.py.
struct WPairs {
  float* K; uint32_t* I;
  PYGB_HD inline void move(int d, int s) const { K[d] = K[s]; I[d] = I[s]; }
  PYGB_HD inline void swap(int x, int y) const {
    const float k = K[x]; const uint32_t i = I[x]; K[x] = K[y]; I[x] = I[y]; K[y] = k; I[y] = i;
  }
  PYGB_HD inline void put(int d, float k, uint32_t i) const { K[d] = k; I[d] = i; }
};
PYGB_HD inline bool w_before(float x, float y) { return ((x != x) && !(y != y)) || (x > y); }

PYGB_HD inline void w_push_heap(const WPairs& a, int first, int hole, int top, float vk, uint32_t vi) {
  int parent = (hole - 1) / 2;
  while (hole > top && w_before(a.K[first + parent], vk)) { a.move(first + hole, first + parent); hole = parent; parent = (hole - 1) / 2; }
  a.put(first + hole, vk, vi);
}
PYGB_HD inline void w_adjust_heap(const WPairs& a, int first, int hole, int len, float vk, uint32_t vi) {
  const int top = hole;
  int child = hole;
  while (child < (len - 1) / 2) {
    child = 2 * (child + 1);
    if (w_before(a.K[first + child], a.K[first + child - 1])) --child;
    a.move(first + hole, first + child); hole = child;
  }
  if ((len & 1) == 0 && child == (len - 2) / 2) {
    child = 2 * (child + 1);
    a.move(first + hole, first + child - 1); hole = child - 1;
  }
  w_push_heap(a, first, hole, top, vk, vi);
}
PYGB_HD inline void w_make_heap(const WPairs& a, int first, int last) {
  const int len = last - first;
  if (len < 2) return;
  for (int parent = (len - 2) / 2;; --parent) {
    w_adjust_heap(a, first, parent, len, a.K[first + parent], a.I[first + parent]);
    if (parent == 0) return;
  }
}
PYGB_HD inline void w_pop_heap(const WPairs& a, int first, int last, int result) {
  const float vk = a.K[result]; const uint32_t vi = a.I[result];
  a.move(result, first);
  w_adjust_heap(a, first, 0, last - first, vk, vi);
}
PYGB_HD inline void w_heap_select(const WPairs& a, int first, int middle, int last) {
  w_make_heap(a, first, middle);
  for (int i = middle; i < last; ++i)
    if (w_before(a.K[i], a.K[first])) w_pop_heap(a, first, middle, i);
}
PYGB_HD inline void w_sort_heap(const WPairs& a, int first, int last) {
  while (last - first > 1) { --last; w_pop_heap(a, first, last, last); }
}
PYGB_HD inline void w_move_median_to_first(const WPairs& a, int r, int x, int y, int z) {
  if (w_before(a.K[x], a.K[y])) {
    if (w_before(a.K[y], a.K[z])) a.swap(r, y);
    else if (w_before(a.K[x], a.K[z])) a.swap(r, z);
    else a.swap(r, x);
  } else if (w_before(a.K[x], a.K[z])) a.swap(r, x);
  else if (w_before(a.K[y], a.K[z])) a.swap(r, z);
  else a.swap(r, y);
}
PYGB_HD inline int w_partition_pivot(const WPairs& a, int first, int last) {
  const int mid = first + (last - first) / 2;
  w_move_median_to_first(a, first, first + 1, mid, last - 1);
  const int pivot = first;
  int lo = first + 1, hi = last;
  while (true) {
    while (w_before(a.K[lo], a.K[pivot])) ++lo;
    --hi;
    while (w_before(a.K[pivot], a.K[hi])) --hi;
    if (!(lo < hi)) return lo;
    a.swap(lo, hi);
    ++lo;
  }
}
PYGB_HD inline void w_unguarded_linear_insert(const WPairs& a, int last) {
  const float vk = a.K[last]; const uint32_t vi = a.I[last];
  int next = last - 1;
  while (w_before(vk, a.K[next])) { a.move(last, next); last = next; --next; }
  a.put(last, vk, vi);
}
PYGB_HD inline void w_insertion_sort(const WPairs& a, int first, int last) {
  if (first == last) return;
  for (int i = first + 1; i < last; ++i) {
    if (w_before(a.K[i], a.K[first])) {
      const float vk = a.K[i]; const uint32_t vi = a.I[i];
      for (int j = i; j > first; --j) a.move(j, j - 1);
      a.put(first, vk, vi);
    } else w_unguarded_linear_insert(a, i);
  }
}
PYGB_HD inline int w_lg(int n) {
#ifdef __CUDA_ARCH__
  return 31 - __clz(n);
#else
  return 31 - __builtin_clz((unsigned)n);
#endif
}
PYGB_HD inline void w_nth_element(const WPairs& a, int first, int nth, int last) {
  if (first == last || nth == last) return;
  int depth = w_lg(last - first) * 2;
  while (last - first > 3) {
    if (depth == 0) { w_heap_select(a, first, nth + 1, last); a.swap(first, nth); return; }
    --depth;
    const int cut = w_partition_pivot(a, first, last);
    if (cut <= nth) first = cut; else last = cut;
  }
  w_insertion_sort(a, first, last);
}
PYGB_HD inline void w_sort(const WPairs& a, int first, int last) {
  if (first == last) return;
  // __introsort_loop: the recursion on [cut, last) only touches that range, so an explicit stack replays it in any order
  int sf[64], sl[64], sd[64], sp = 0;
  sf[0] = first; sl[0] = last; sd[0] = w_lg(last - first) * 2; sp = 1;
  while (sp > 0) {
    --sp;
    int f = sf[sp], l = sl[sp], d = sd[sp];
    while (l - f > 16) {
      if (d == 0) { w_heap_select(a, f, l, l); w_sort_heap(a, f, l); break; }
      --d;
      const int cut = w_partition_pivot(a, f, l);
      if (sp < 64) { sf[sp] = cut; sl[sp] = l; sd[sp] = d; ++sp; }
      l = cut;
    }
  }
  if (last - first > 16) {   // __final_insertion_sort
    w_insertion_sort(a, first, first + 16);
    for (int i = first + 16; i < last; ++i) w_unguarded_linear_insert(a, i);
  } else w_insertion_sort(a, first, last);
}
// at::topk(k) of the n pairs: afterwards I[0..k) are the returned indices
PYGB_HD inline void w_topk_replay(const WPairs& a, int n, int k) {
  if ((int64_t)k * 64 <= (int64_t)n) { w_heap_select(a, 0, k, n); w_sort_heap(a, 0, k); }
  else { w_nth_element(a, 0, k - 1, n); w_sort(a, 0, k - 1); }
}

