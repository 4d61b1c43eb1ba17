"""Golden fixtures for pyg::relabel_neighborhood from the REFERENCE itself (oracle/_ref/libpyg_ref.so).
Run in this container only:  python tests/golden/make_golden_relabel.py  ->  tests/golden/relabel_outputs.npz"""
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

from graphs import RELABEL_CASES, build_relabel  # noqa


def ref_dist(rowptr, col, seed, k, replace=False, disjoint=False):
    return torch.ops.pyg.dist_neighbor_sample(rowptr, col, seed, k, None, None, None, None, False, replace, True, disjoint, 'uniform')


def main():
    out = {}
    for name, case in RELABEL_CASES.items():
        seed, sampled, counts, num_nodes, batch, csc, disjoint = build_relabel(case, ref_dist)
        r = torch.ops.pyg.relabel_neighborhood(seed, sampled, counts, num_nodes, batch, csc, disjoint)
        out[f'{name}/row'] = r[0].numpy(); out[f'{name}/col'] = r[1].numpy()
        print(name, sampled.numel(), 'sampled ->', int(r[1].max()) + 1 if r[1].numel() else 0, 'ids')
    np.savez_compressed(osp.join(HERE, 'relabel_outputs.npz'), **out)


if __name__ == '__main__':
    main()
