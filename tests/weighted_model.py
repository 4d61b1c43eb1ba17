"""A word-level model of what the CUDA kernels of the biased (edge_weight) sampler compute, in numpy — independent of the ATen
calls oracle/weighted.py makes, so that the two pin each other on the CPU (tests/test_weighted_oracle.py) before any GPU runs:

  * raw mt19937 words of the torch CPU generator; uniform_ on float32 = (word & 0xFFFFFF) * 2^-24, one word per element
    (at::uniform_real_distribution<float>); a double in [0,1) = ((hi << 32 | lo) & (2^53-1)) * 2^-53, hi word first
  * torch.log on CPU float32 (MKL vmsLn) = correctly rounded log + pyg_lib_b200/csrc/mkl_logf_table.inc
  * at::topk on CPU = libstdc++ std::partial_sort (k * 64 <= n) or std::nth_element + std::sort of (value, index) pairs
    compared by value only, NaN first (ATen/native/cpu/TopKImpl.h) — restated below; without ties among the k + 1 largest
    keys the result is simply "descending by key", which is what the GPU's fast path computes
  * at::multinomial(replacement=True, n > 1) = float32 running sum, divided by the total, last bucket forced to 1, a lower-bound
    binary search per double sample (ATen/native/cpu/MultinomialKernel.cpp)
"""
import math
import os
import sys
from typing import Dict, List, Tuple

import numpy as np
import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, os.path.join(ROOT, 'tools'))


# ------------------------------------------------------------------------------------------ random words
class Words:
    """The next outputs of torch's default CPU generator, read without advancing it."""

    def __init__(self):
        raw = torch.get_rng_state().numpy()
        st = np.frombuffer(raw[24:24 + 624 * 8].tobytes(), dtype=np.uint64).astype(np.uint32)
        left = int(np.frombuffer(raw[8:12].tobytes(), dtype=np.int32)[0])
        self.rs = np.random.RandomState()
        self.rs.set_state(('MT19937', st, 625 - left))
        self.buf = np.zeros(0, dtype=np.uint32)
        self.pos = 0

    def take(self, n: int) -> np.ndarray:
        while self.buf.size - self.pos < n:
            more = self.rs.randint(0, 1 << 32, size=max(4096, n), dtype=np.uint64).astype(np.uint32)
            self.buf = np.concatenate([self.buf[self.pos:], more]); self.pos = 0
        out = self.buf[self.pos:self.pos + n]
        self.pos += n
        return out


# ------------------------------------------------------------------------------------------ log
_TABLE = None


def mkl_logf(k24: np.ndarray) -> np.ndarray:
    """torch.log(float32(k * 2^-24)) bit for bit."""
    global _TABLE
    if _TABLE is None:
        from make_logf_table import read_inc
        t = read_inc()
        _TABLE = dict(zip((t >> 1).tolist(), (t & 1).tolist()))
    u = k24.astype(np.float32) * np.float32(2.0 ** -24)
    with np.errstate(divide='ignore'):
        cr = np.log(u.astype(np.float64)).astype(np.float32)
    bits = cr.view(np.int32).copy()
    for i, k in enumerate(k24.tolist()):
        d = _TABLE.get(k)
        if d is not None:
            bits[i] += 1 if d else -1
    return bits.view(np.float32)


# ------------------------------------------------------------------------------------------ libstdc++ top-k
def comp(x, y):
    xn, yn = math.isnan(x[0]), math.isnan(y[0])
    return (xn and not yn) or (x[0] > y[0])


def _push_heap(a, first, hole, top, val):
    parent = (hole - 1) // 2
    while hole > top and comp(a[first + parent], val):
        a[first + hole] = a[first + parent]; hole = parent; parent = (hole - 1) // 2
    a[first + hole] = val


def _adjust_heap(a, first, hole, ln, val):
    top = hole; child = hole
    while child < (ln - 1) // 2:
        child = 2 * (child + 1)
        if comp(a[first + child], a[first + child - 1]): child -= 1
        a[first + hole] = a[first + child]; hole = child
    if (ln & 1) == 0 and child == (ln - 2) // 2:
        child = 2 * (child + 1)
        a[first + hole] = a[first + child - 1]; hole = child - 1
    _push_heap(a, first, hole, top, val)


def _make_heap(a, first, last):
    ln = last - first
    if ln < 2: return
    parent = (ln - 2) // 2
    while True:
        _adjust_heap(a, first, parent, ln, a[first + parent])
        if parent == 0: return
        parent -= 1


def _pop_heap(a, first, last, result):
    val = a[result]; a[result] = a[first]
    _adjust_heap(a, first, 0, last - first, val)


def _heap_select(a, first, middle, last):
    _make_heap(a, first, middle)
    for i in range(middle, last):
        if comp(a[i], a[first]): _pop_heap(a, first, middle, i)


def _sort_heap(a, first, last):
    while last - first > 1:
        last -= 1; _pop_heap(a, first, last, last)


def _partial_sort(a, first, middle, last):
    _heap_select(a, first, middle, last); _sort_heap(a, first, middle)


def _move_median_to_first(a, r, x, y, z):
    if comp(a[x], a[y]):
        if comp(a[y], a[z]): a[r], a[y] = a[y], a[r]
        elif comp(a[x], a[z]): a[r], a[z] = a[z], a[r]
        else: a[r], a[x] = a[x], a[r]
    elif comp(a[x], a[z]): a[r], a[x] = a[x], a[r]
    elif comp(a[y], a[z]): a[r], a[z] = a[z], a[r]
    else: a[r], a[y] = a[y], a[r]


def _unguarded_partition(a, first, last, pivot):
    while True:
        while comp(a[first], a[pivot]): first += 1
        last -= 1
        while comp(a[pivot], a[last]): last -= 1
        if not first < last: return first
        a[first], a[last] = a[last], a[first]; first += 1


def _partition_pivot(a, first, last):
    mid = first + (last - first) // 2
    _move_median_to_first(a, first, first + 1, mid, last - 1)
    return _unguarded_partition(a, first + 1, last, first)


def _unguarded_linear_insert(a, last):
    val = a[last]; nxt = last - 1
    while comp(val, a[nxt]):
        a[last] = a[nxt]; last = nxt; nxt -= 1
    a[last] = val


def _insertion_sort(a, first, last):
    if first == last: return
    for i in range(first + 1, last):
        if comp(a[i], a[first]):
            val = a[i]; a[first + 1:i + 1] = a[first:i]; a[first] = val
        else: _unguarded_linear_insert(a, i)


def _lg(n): return n.bit_length() - 1


def _nth_element(a, first, nth, last):
    if first == last or nth == last: return
    depth = _lg(last - first) * 2
    while last - first > 3:
        if depth == 0:
            _heap_select(a, first, nth + 1, last); a[first], a[nth] = a[nth], a[first]; return
        depth -= 1
        cut = _partition_pivot(a, first, last)
        if cut <= nth: first = cut
        else: last = cut
    _insertion_sort(a, first, last)


def _introsort_loop(a, first, last, depth):
    while last - first > 16:
        if depth == 0:
            _partial_sort(a, first, last, last); return
        depth -= 1
        cut = _partition_pivot(a, first, last)
        _introsort_loop(a, cut, last, depth)
        last = cut


def _std_sort(a, first, last):
    if first == last: return
    _introsort_loop(a, first, last, _lg(last - first) * 2)
    if last - first > 16:
        _insertion_sort(a, first, first + 16)
        for i in range(first + 16, last): _unguarded_linear_insert(a, i)
    else: _insertion_sort(a, first, last)


def topk_libstdcxx(vals, k: int) -> List[int]:
    """Indices at::topk(k) returns for a 1-D float tensor on CPU, ties included."""
    n = len(vals); a = [(float(v), j) for j, v in enumerate(vals)]
    if k == 0: return []
    if k * 64 <= n: _partial_sort(a, 0, k, n)
    else:
        _nth_element(a, 0, k - 1, n)
        _std_sort(a, 0, k - 1)
    return [a[j][1] for j in range(k)]


def topk_gpu(vals, k: int) -> Tuple[List[int], bool]:
    """The GPU's way: k rounds of arg-max in the total order (NaN first, value descending, index ascending); returns the
    indices and whether two neighbours among the k + 1 first are comp-equivalent (then the exact emulation above decides)."""
    n = len(vals)
    order = sorted(range(n), key=lambda j: (0 if math.isnan(vals[j]) else 1, -vals[j] if not math.isnan(vals[j]) else 0.0, j))
    top = order[:min(k + 1, n)]
    tie = any(not comp((float(vals[top[t]]), 0), (float(vals[top[t + 1]]), 0)) for t in range(len(top) - 1))
    return order[:k], tie


# ------------------------------------------------------------------------------------------ per-node draws
def biased_indices(words: Words, weight: np.ndarray, count: int, replace: bool, stats=None):
    """None = full neighbourhood; else in-row indices in emission order.  Consumes words like the reference."""
    n = weight.size
    if count < 0 or (not replace and count >= n):
        return None
    if replace:
        assert count > 1, 'at::multinomial(n_sample=1) takes the exponential_ path (MKL VSL stream): not modelled'
        cum = np.empty(n, np.float32); s = np.float32(0)
        for j in range(n):
            s = np.float32(s + weight[j]); cum[j] = s
        cum = (cum / s).astype(np.float32)
        cum[n - 1] = 1
        w = words.take(2 * count)
        out = []
        for j in range(count):
            v = (int(w[2 * j]) << 32) | int(w[2 * j + 1])
            u = (v & ((1 << 53) - 1)) * 2.0 ** -53
            lo, hi = 0, n
            while hi - lo > 0:
                mid = lo + (hi - lo) // 2
                if float(cum[mid]) < u: lo = mid + 1
                else: hi = mid
            out.append(lo)
        return out
    k24 = (words.take(n) & np.uint32(0xFFFFFF)).astype(np.int64)
    with np.errstate(divide='ignore', invalid='ignore'):
        key = (mkl_logf(k24) / weight.astype(np.float32)).astype(np.float32)
    idx, tie = topk_gpu(key.tolist(), count)
    if tie:
        if stats is not None: stats['ties'] = stats.get('ties', 0) + 1
        idx = topk_libstdcxx(key.tolist(), count)
    return idx


def neighbor_sample(rowptr, col, seed, num_neighbors, edge_weight, replace=False, stats=None):
    """Homogeneous, non-disjoint.  Returns (row, col, node, eid, nph, eph, words consumed incl. the engine's 256)."""
    rp, cl, sd = rowptr.tolist(), col.tolist(), seed.tolist()
    w = edge_weight.numpy()
    words = Words(); words.take(256)
    ids: Dict[int, int] = {}
    nodes = []
    for v in sd:
        nodes.append(v); ids.setdefault(v, len(ids))
    rows, cols, eids, nph, eph = [], [], [], [len(sd)], []
    begin, end = 0, len(sd)
    used = 256
    for count in num_neighbors:
        n_e = 0
        for i in range(begin, end):
            v = nodes[i]; rs, re = rp[v], rp[v + 1]
            if re == rs or count == 0: continue
            before = words.pos
            idx = biased_indices(words, w[rs:re], count, replace, stats)
            edges = range(rs, re) if idx is None else [rs + j for j in idx]
            for e in edges:
                d = cl[e]
                if d not in ids:
                    ids[d] = len(ids); nodes.append(d)
                rows.append(i); cols.append(ids[d]); eids.append(e); n_e += 1
        begin, end = end, len(nodes)
        nph.append(end - begin); eph.append(n_e)
    T = lambda x: torch.tensor(x, dtype=seed.dtype)   # noqa: E731
    # words consumed: recomputed from the degrees (what the kernels' count pass does)
    return T(rows), T(cols), T(nodes), T(eids), nph, eph
