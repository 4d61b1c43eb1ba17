"""Golden fixtures of BIASED (edge_weight) sampling, produced by the REFERENCE itself (oracle/_ref/libpyg_ref.so built from
/root/reference by oracle/build_ref.sh).  Run in this container only:

    python tests/golden/make_golden_weighted.py

Writes tests/golden/weighted_outputs.npz.  Inputs are regenerated from tests/graphs.py (WEIGHTED_CASES,
HETERO_WEIGHTED_CASES); outputs and the CPU generator state after the call are stored."""
import os
import os.path as osp
import sys

import numpy as np
import torch

HERE = osp.dirname(osp.abspath(__file__))
ROOT = osp.dirname(osp.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, osp.join(ROOT, 'tests'))

torch.ops.load_library(osp.join(ROOT, 'oracle', '_ref', 'libpyg_ref.so'))
torch.set_num_threads(1)

from graphs import WEIGHTED_CASES, HETERO_WEIGHTED_CASES, build_weighted, build_hetero_weighted  # noqa


def np_(t):
    return t.numpy() if isinstance(t, torch.Tensor) else np.asarray(t, dtype=np.int64)


def main():
    out = {}
    for name, case in WEIGHTED_CASES.items():
        rowptr, col, seed, w = build_weighted(case)
        torch.manual_seed(case['rng_seed'])
        r = torch.ops.pyg.neighbor_sample(rowptr, col, seed, case['num_neighbors'], None, None, None, w, case.get('csc', False),
                                          case.get('replace', False), True, case.get('disjoint', False), 'uniform', True)
        for k, v in zip(('row', 'col', 'node', 'eid', 'nph', 'eph'), r):
            out[f'homo/{name}/{k}'] = np_(v)
        out[f'homo/{name}/rng_after'] = torch.get_rng_state().numpy()[:24 + 624 * 8].copy()
        print('homo', name, r[0].numel(), 'edges', r[2].shape[0], 'nodes')
    for name, case in HETERO_WEIGHTED_CASES.items():
        nt, et, rp, cl, sd, nn, wd = build_hetero_weighted(case)
        torch.manual_seed(case['rng_seed'])
        r = torch.ops.pyg.hetero_neighbor_sample(nt, et, rp, cl, sd, nn, None, None, None, wd, case.get('csc', False),
                                                 case.get('replace', False), True, case.get('disjoint', False), 'uniform', True)
        for i, key in enumerate(('row', 'col', 'node', 'eid', 'nph', 'eph')):
            for k, v in r[i].items():
                out[f'hetero/{name}/{key}/{k}'] = np_(v)
        out[f'hetero/{name}/rng_after'] = torch.get_rng_state().numpy()[:24 + 624 * 8].copy()
        print('hetero', name, sum(v.numel() for v in r[0].values()), 'edges')
    path = osp.join(HERE, 'weighted_outputs.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, os.path.getsize(path), 'bytes')


if __name__ == '__main__':
    main()
