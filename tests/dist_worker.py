"""Worker for the multi-process sampler tests (launched by torch.distributed.run).
usage: dist_worker.py <backend> <mode>     mode: 'segments' (CPU host logic) | 'sample' (GPU parity)"""
import os
import os.path as osp
import sys

import torch
import torch.distributed as dist

ROOT = osp.dirname(osp.dirname(osp.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, osp.join(ROOT, 'tests'))


def main():
    backend, mode = sys.argv[1], sys.argv[2]
    rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
    os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
    if mode == 'segments':
        dist.init_process_group(backend)
        from pyg_lib_b200.sampler.dist import allgather_segments, segment_bounds
        for total in (0, 1, 7, 1000):
            seg = segment_bounds(total, world)
            assert seg[0] == 0 and seg[-1] == total and all(b >= a for a, b in zip(seg, seg[1:]))
            buf = torch.full((total,), -1, dtype=torch.int64)
            buf[seg[rank]:seg[rank + 1]] = torch.arange(seg[rank], seg[rank + 1]) * 10 + rank
            allgather_segments(buf, seg)
            exp = torch.cat([torch.arange(seg[q], seg[q + 1]) * 10 + q for q in range(world)]) if total else buf
            assert torch.equal(buf, exp), (rank, total)
        # host all-gather of equally sized blobs (swaps the CUDA IPC handles of the peer-memory transport)
        from pyg_lib_b200.sampler.dist import allgather_blobs
        blob = bytes([(rank * 7 + i) % 256 for i in range(64)])
        got = allgather_blobs(blob, torch.device('cpu'))
        assert got == b''.join(bytes([(q * 7 + i) % 256 for i in range(64)]) for q in range(world)), rank
        # uneven, with empty segments
        seg = [0, 0, 5][:world + 1] if world == 2 else segment_bounds(9, world)
        buf = torch.zeros(seg[-1], dtype=torch.int64)
        buf[seg[rank]:seg[rank + 1]] = rank + 1
        allgather_segments(buf, seg)
        assert all(int(buf[i]) == q + 1 for q in range(world) for i in range(seg[q], seg[q + 1]))
    else:
        ngpu = torch.cuda.device_count()
        dev = torch.device('cuda', rank % ngpu)
        torch.cuda.set_device(dev)
        dist.init_process_group(backend, device_id=dev if backend == 'nccl' else None)
        import pyg_lib_b200 as P
        from graphs import random_csr
        from oracle import oracle as O
        rowptr, col = random_csr(20000, 30, 0, big=[(5, 70000), (77, 65540)])
        seed = torch.randperm(20000, generator=torch.Generator().manual_seed(5))[:512]
        seed[3], seed[9] = 5, 77
        d = [t.to(dev) for t in (rowptr, col, seed)]
        # both transports: peer memory (IPC-mapped exchange regions, key-partitioned dedup) and collective
        # (edge ids all-gathered by torch.distributed, replicated dedup); disjoint runs always take the latter
        for kw in (dict(), dict(replace=True), dict(csc=True, return_edge_id=False), dict(disjoint=True),
                   dict(transport='collective'), dict(transport='collective', replace=True)):
            for nn in ([15, 10], [4, 3, 2], [40]):
                torch.manual_seed(11)
                exp = [O.neighbor_sample(rowptr, col, seed, nn, **{k: v for k, v in kw.items() if k != 'transport'}) for _ in range(2)]
                s_exp = torch.get_rng_state()
                torch.manual_seed(11)
                for i in range(2):
                    out = P.sampler.dist_neighbor_sample(d[0], d[1], d[2], nn, **kw)
                    for a, b in zip(out[:4], exp[i][:4]):
                        assert (a is None and b is None) or torch.equal(a.cpu(), b), (rank, kw, nn)
                    assert out[4] == exp[i][4] and out[5] == exp[i][5]
                assert torch.equal(torch.get_rng_state()[:24 + 624 * 8], s_exp[:24 + 624 * 8])
        # the single-GPU op interleaves with the sharded one on the same generator
        torch.manual_seed(3)
        e1 = O.neighbor_sample(rowptr, col, seed, [5, 5]); e2 = O.neighbor_sample(rowptr, col, seed, [5, 5])
        torch.manual_seed(3)
        o1 = P.sampler.neighbor_sample(d[0], d[1], d[2], [5, 5]); o2 = P.sampler.dist_neighbor_sample(d[0], d[1], d[2], [5, 5])
        assert torch.equal(o1[0].cpu(), e1[0]) and torch.equal(o2[0].cpu(), e2[0]) and torch.equal(o2[2].cpu(), e2[2])
        torch.cuda.synchronize()
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print('DIST_OK')


if __name__ == '__main__':
    main()
