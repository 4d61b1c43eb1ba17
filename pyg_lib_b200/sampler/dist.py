"""Frontier-sharded neighbor sampling of ONE batch across the GPUs of a box (BASELINE
configs[4]).  The reference has no multi-GPU sampler; its own decomposition of distributed sampling is
`dist_neighbor_sample` (sample without relabel, neighbor_kernel.cpp:296-303,957-978) + merge + relabel.
Here: the CSR is replicated, every rank calls `dist_neighbor_sample` with the same arguments and the
same CPU generator state (`torch.manual_seed(s)` on every rank); per hop each rank draws the edges of
its slice of the frontier and every rank returns the SAME tensors as the single-GPU op (and as the
reference).  Two transports (include/pyg_b200.h, `pygb200_sampler_run_sharded`):

  * peer memory (default): the sampling kernel stores its edges straight into every rank's exchange
    region over NVLink, dedup is partitioned by key hash, refs are reduced slice-wise over the peers —
    no host sync, no collective call per hop (csrc/sampler_v2.cuh).  torch.distributed is only used to
    swap the CUDA IPC handles of the regions when they are (re)allocated.
  * `transport='collective'` (and always for disjoint runs): the drawn edge ids are all-gathered with
    torch.distributed broadcasts (NCCL over NVLink) and dedup / relabel run replicated.

Plumbing is torch.distributed; the sampling itself is `pygb200_sampler_run_sharded` of the C ABI.
"""
import ctypes as C
import os.path as osp
from typing import List, Optional, Tuple

import torch
import torch.distributed as dist
from torch import Tensor

from .._rng import MT19937, read_default_cpu_engine, write_default_cpu_engine

_ALLGATHER_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_int64), C.c_int32, C.c_void_p)
_EXCHANGE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64)


class _Shard(C.Structure):
    _fields_ = [('rank', C.c_int32), ('world', C.c_int32), ('allgather', _ALLGATHER_FN), ('user', C.c_void_p),
                ('exchange', _EXCHANGE_FN)]


def allgather_blobs(mine: bytes, device: torch.device, group=None) -> bytes:
    """Host all-gather of equally sized blobs in rank order (used to swap CUDA IPC handles)."""
    world = dist.get_world_size(group)
    backend = dist.get_backend(group)
    t = torch.frombuffer(bytearray(mine), dtype=torch.uint8)
    if backend == 'nccl':
        t = t.to(device)
    out = torch.empty(world * t.numel(), dtype=torch.uint8, device=t.device)
    dist.all_gather_into_tensor(out, t, group=group) if backend == 'nccl' else \
        dist.all_gather(list(out.view(world, -1).unbind(0)), t, group=group)
    return bytes(out.cpu().numpy().tobytes())


class _Relation(C.Structure):
    _fields_ = [('rowptr', C.c_void_p), ('col', C.c_void_p), ('num_src_nodes', C.c_int64), ('num_edges', C.c_int64),
                ('src_type', C.c_int32), ('dst_type', C.c_int32)]


class _DevBuf:
    """Zero-copy int64 view of a raw device pointer for torch (CUDA array interface)."""
    def __init__(self, ptr: int, n: int):
        self.__cuda_array_interface__ = {'shape': (n,), 'typestr': '<i8', 'data': (ptr, False), 'version': 3}


def segment_bounds(total: int, world: int) -> List[int]:
    """Even split of `total` frontier nodes by index: rank q gets [total*q//world, total*(q+1)//world)."""
    return [total * q // world for q in range(world + 1)]


def allgather_segments(buf: Tensor, seg_begin: List[int], group=None) -> None:
    """In-place all-gather of variable-length contiguous segments: after the call
    buf[seg_begin[q]:seg_begin[q+1]] holds rank q's data on every rank.  One broadcast per non-empty
    segment, issued back to back (NCCL executes them on its stream in order)."""
    world = dist.get_world_size(group)
    works = []
    for q in range(world):
        a, b = int(seg_begin[q]), int(seg_begin[q + 1])
        if b > a:
            src = dist.get_global_rank(group, q) if group is not None else q
            works.append(dist.broadcast(buf[a:b], src=src, group=group, async_op=True))
    for w in works:
        w.wait()


_lib = None
_handles = {}


def _abi():
    global _lib
    if _lib is None:
        _lib = C.CDLL(osp.join(osp.dirname(osp.dirname(osp.abspath(__file__))), 'libpyg_b200.so'))
        _lib.pygb200_last_error.restype = C.c_char_p
    return _lib


def _handle(device: torch.device, stream: int):
    key = (device.index, stream)
    if key not in _handles:
        h = C.c_void_p()
        with torch.cuda.device(device):
            if _abi().pygb200_sampler_create(C.byref(h)) != 0:
                raise RuntimeError('pyg_lib_b200: ' + _abi().pygb200_last_error().decode())
        _handles[key] = h
    return _handles[key]


def dist_neighbor_sample(rowptr: Tensor, col: Tensor, seed: Tensor, num_neighbors: List[int], csc: bool = False,
                         replace: bool = False, disjoint: bool = False, return_edge_id: bool = True,
                         group=None, transport: str = 'peer') -> Tuple[Tensor, Tensor, Tensor, Optional[Tensor], List[int], List[int]]:
    """Same contract as `neighbor_sample` (uniform sampling, fan-outs >= 0); collective over `group`.
    `transport`: 'peer' (peer-memory stores + key-partitioned dedup; non-disjoint runs, <= 16 ranks) or 'collective'
    (edge ids all-gathered with torch.distributed, replicated dedup)."""
    assert transport in ('peer', 'collective')
    assert rowptr.is_cuda and col.is_cuda and seed.is_cuda, 'dist_neighbor_sample expects CUDA tensors'
    assert rowptr.dtype == col.dtype == seed.dtype and seed.dtype in (torch.int64, torch.int32)
    assert rowptr.is_contiguous() and col.is_contiguous() and seed.is_contiguous()
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    lib = _abi()
    dev = seed.device
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream(dev).cuda_stream
        h = _handle(dev, stream)
        L = len(num_neighbors)
        rel = _Relation(rowptr.data_ptr(), col.data_ptr(), rowptr.numel() - 1, col.numel(), 0, 0)
        seeds = (C.c_void_p * 1)(seed.data_ptr())
        n_seeds = (C.c_int64 * 1)(seed.numel())
        nn = (C.c_int64 * max(L, 1))(*num_neighbors)
        nph, eph = (C.c_int64 * (L + 1))(), (C.c_int64 * max(L, 1))()
        n_nodes, n_edges = (C.c_int64 * 1)(), (C.c_int64 * 1)()
        idx32 = seed.dtype == torch.int32
        flags = (1 if replace else 0) | (2 if disjoint else 0) | (4 if idx32 else 0)
        err = []

        def _cb(user, buf, seg_begin, w, strm):
            try:
                seg = [seg_begin[i] for i in range(w + 1)]
                if seg[w] > 0:
                    t = torch.as_tensor(_DevBuf(buf, seg[w]), device=dev)
                    allgather_segments(t, seg, group)
                return 0
            except Exception as ex:  # noqa: never let an exception cross the C boundary
                err.append(ex)
                return 1

        def _xcb(user, mine, out, nbytes):
            try:
                blob = C.string_at(mine, nbytes)
                C.memmove(out, allgather_blobs(blob, dev, group), nbytes * world)
                return 0
            except Exception as ex:  # noqa
                err.append(ex)
                return 1

        peer = transport == 'peer' and not disjoint and world <= 16 and len(num_neighbors) > 0 and \
            min(num_neighbors) >= 0 and rowptr.numel() - 1 < 0xffffffff
        shard = _Shard(rank, world, _ALLGATHER_FN(_cb), None, _EXCHANGE_FN(_xcb) if peer else _EXCHANGE_FN())
        mt = read_default_cpu_engine()
        rc = lib.pygb200_sampler_run_sharded(h, 1, 1, L, C.byref(rel), seeds, n_seeds, nn, flags, C.byref(mt), nph, eph,
                                             n_nodes, n_edges, C.c_void_p(stream), C.byref(shard))
        if rc != 0:
            raise RuntimeError('pyg_lib_b200: ' + lib.pygb200_last_error().decode() + (f' ({err[0]!r})' if err else ''))
        write_default_cpu_engine(mt)
        row = torch.empty(n_edges[0], dtype=seed.dtype, device=dev)
        colv = torch.empty_like(row)
        eid = torch.empty_like(row) if return_edge_id else None
        node = torch.empty((n_nodes[0], 2) if disjoint else (n_nodes[0],), dtype=seed.dtype, device=dev)
        lib.pygb200_sampler_export_edges(h, 0, C.c_void_p(row.data_ptr()), C.c_void_p(colv.data_ptr()),
                                         C.c_void_p(eid.data_ptr()) if return_edge_id else None, int(idx32), C.c_void_p(stream))
        lib.pygb200_sampler_export_nodes(h, 0, C.c_void_p(node.data_ptr()), int(idx32), C.c_void_p(stream))
    if csc:
        row, colv = colv, row
    return row, colv, node, eid, list(nph), list(eph)[:L]


__all__ = ['dist_neighbor_sample', 'allgather_segments', 'segment_bounds']
