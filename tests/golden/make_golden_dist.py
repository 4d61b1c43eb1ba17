"""Golden fixtures for pyg::dist_neighbor_sample from the REFERENCE itself (oracle/_ref/libpyg_ref.so).
Run in this container only:  python tests/golden/make_golden_dist.py  ->  tests/golden/dist_outputs.npz"""
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

from graphs import DIST_CASES, build_dist  # noqa


def main():
    out = {}
    for name, case in DIST_CASES.items():
        rowptr, col, seed = build_dist(case)
        torch.manual_seed(case['rng_seed'])
        r = torch.ops.pyg.dist_neighbor_sample(rowptr, col, seed, case['k'], None, None, None, None, False,
                                               case.get('replace', False), True, case.get('disjoint', False), 'uniform')
        out[f'{name}/node'] = r[0].numpy(); out[f'{name}/eid'] = r[1].numpy(); out[f'{name}/cumsum'] = np.asarray(r[2], dtype=np.int64)
        out[f'{name}/rng_after'] = torch.get_rng_state().numpy()[:24 + 624 * 8].copy()
        print(name, seed.numel(), 'seeds ->', r[1].numel(), 'edges')
    np.savez_compressed(osp.join(HERE, 'dist_outputs.npz'), **out)


if __name__ == '__main__':
    main()
