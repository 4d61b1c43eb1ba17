"""Golden fixtures for pyg::hetero_relabel_neighborhood from the REFERENCE itself (oracle/_ref/libpyg_ref.so).
Run in this container only:  python tests/golden/make_golden_hetero_relabel.py  ->  tests/golden/hetero_relabel_outputs.npz"""
import os.path as osp
import sys

import numpy as np
import torch

HERE = osp.dirname(osp.abspath(__file__))
ROOT = osp.dirname(osp.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, osp.join(ROOT, 'tests'))
torch.ops.load_library(osp.join(ROOT, 'oracle', '_ref', 'libpyg_ref.so'))
torch.set_num_threads(1)   # (the reference groups edge types by dst type per thread; one thread = list order)

from graphs import HETERO_RELABEL_CASES, build_hetero_relabel  # noqa


def main():
    out = {}
    for name, case in HETERO_RELABEL_CASES.items():
        nt, et, seed_d, sampled_d, counts_d, nn_d, batch_d, csc, disjoint = build_hetero_relabel(case)
        r = torch.ops.pyg.hetero_relabel_neighborhood(nt, et, seed_d, sampled_d, counts_d, nn_d, batch_d, csc, disjoint)
        for k, v in r[0].items(): out[f'{name}/row/{k}'] = v.numpy()
        for k, v in r[1].items(): out[f'{name}/col/{k}'] = v.numpy()
        print(name, {k: v.numel() for k, v in r[0].items()})
    np.savez_compressed(osp.join(HERE, 'hetero_relabel_outputs.npz'), **out)


if __name__ == '__main__':
    main()
