"""Biased (edge_weight) sampling: time per call on the GPU next to the reference's CPU kernel.

    python tools/weighted_bench.py > gpurun_out/weighted_bench.json

ogbn-products-shaped CSR, uniform float32 weights; 1024 seeds [15,10] (the loader-sized batch of BASELINE configs[1]) and 65,536
seeds; with and without replacement.  The reference (oracle/_ref, its own biased_sample: at::multinomial / uniform_ + log + topk
per frontier node) runs in a subprocess on one core and on all cores of the box."""
import json
import os
import subprocess
import sys

import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))


def main():
    import pyg_lib_b200 as P
    from graphs import lognormal_csr
    dev = 'cuda:0'
    n, e = 2_449_029, 123_718_280
    rowptr, col = lognormal_csr(n, e, seed=1, device=dev)
    w = torch.rand(e, device=dev, generator=torch.Generator(device=dev).manual_seed(3))
    perm = torch.randperm(n, generator=torch.Generator().manual_seed(2)).to(dev)
    res = {'graph': 'ogbn-products-shaped CSR (2,449,029 nodes / 123,718,280 edges, int64), uniform float32 edge weights', 'gpu': {}, 'reference': {}}
    for batch in (1024, 65536):
        for replace in (False, True):
            iters = 20 if batch == 1024 else 5
            torch.manual_seed(1)
            edges = 0
            for i in range(3):
                P.sampler.neighbor_sample(rowptr, col, perm[i * batch:(i + 1) * batch], [15, 10], edge_weight=w, replace=replace)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(iters):
                out = P.sampler.neighbor_sample(rowptr, col, perm[(3 + i) * batch:(4 + i) * batch], [15, 10], edge_weight=w, replace=replace)
                edges += out[0].numel()
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / iters
            torch.manual_seed(1)
            un = 0
            e0.record()
            for i in range(iters):
                un += P.sampler.neighbor_sample(rowptr, col, perm[(3 + i) * batch:(4 + i) * batch], [15, 10], replace=replace)[0].numel()
            e1.record(); torch.cuda.synchronize()
            import ctypes as C
            abi = C.CDLL(os.path.join(os.path.dirname(P.__file__), 'libpyg_b200.so'))
            abi.pygb200_profile_read.argtypes = [C.c_char_p, C.c_void_p, C.c_void_p, C.c_void_p]
            abi.pygb200_profile_enable(1)
            for i in range(iters):
                P.sampler.neighbor_sample(rowptr, col, perm[(3 + i) * batch:(4 + i) * batch], [15, 10], edge_weight=w, replace=replace)
            torch.cuda.synchronize()
            abi.pygb200_profile_enable(0)
            prof = {}
            for name in ('count', 'mtgen', 'sample', 'insert', 'mark', 'assign', 'seed', 'final', 'cleanup'):
                msv, ln, wk = C.c_double(), C.c_int64(), C.c_int64()
                abi.pygb200_profile_read(name.encode(), C.byref(msv), C.byref(ln), C.byref(wk))
                if ln.value:
                    prof[name] = round(msv.value / iters * 1e3, 1)
            res['gpu'][f'{batch} seeds [15,10] replace={replace}'] = dict(kernel_us_per_call=prof, ms_per_call=ms, edges_per_call=edges / iters, edges_per_s=edges / iters / ms * 1e3,
                                                                      uniform_ms_per_call=e0.elapsed_time(e1) / iters)
    del rowptr, col, w
    torch.cuda.empty_cache()
    cores = len(os.sched_getaffinity(0))
    for workers in (1, cores):
        r = subprocess.run([sys.executable, os.path.join(ROOT, 'oracle', 'ref_bench.py'), 'sampler', '--workers', str(workers), '--calls', '3',
                            '--graph', 'products', '--weighted'], capture_output=True, text=True)
        line = [ln for ln in r.stdout.splitlines() if ln.startswith('REFBENCH ')]
        res['reference'][f'{workers} worker(s), 1024 seeds [15,10]'] = json.loads(line[0][9:]) if line else {'error': r.stderr[-800:]}
    print(json.dumps(res))


if __name__ == '__main__':
    main()
