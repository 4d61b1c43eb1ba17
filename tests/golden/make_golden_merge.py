"""Golden fixtures for pyg::merge_sampler_outputs from the REFERENCE itself (oracle/_ref/libpyg_ref.so).
Run in this container only:  python tests/golden/make_golden_merge.py  ->  tests/golden/merge_outputs.npz"""
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

from graphs import MERGE_CASES, build_merge  # noqa


def ref_dist(rowptr, col, seed, k, replace=False, disjoint=False):
    return torch.ops.pyg.dist_neighbor_sample(rowptr, col, seed, k, None, None, None, None, False, replace, True, disjoint, 'uniform')


def main():
    out = {}
    for name, case in MERGE_CASES.items():
        a = build_merge(case, ref_dist)
        r = torch.ops.pyg.merge_sampler_outputs(*a)
        out[f'{name}/node'] = r[0].numpy(); out[f'{name}/edge'] = r[1].numpy(); out[f'{name}/counts'] = np.asarray(r[3], dtype=np.int64)
        if r[2] is not None:
            out[f'{name}/batch'] = r[2].numpy()
        print(name, len(a[3]), 'seeds ->', r[0].numel(), 'neighbours')
    np.savez_compressed(osp.join(HERE, 'merge_outputs.npz'), **out)


if __name__ == '__main__':
    main()
