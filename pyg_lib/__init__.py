"""Drop-in alias: `import pyg_lib` resolves to the B200 package so PyG finds
`pyg_lib.ops.segment_matmul`, `pyg_lib.sampler.neighbor_sample`, ... unchanged."""
import sys

import pyg_lib_b200 as _impl
from pyg_lib_b200 import __version__, cuda_version, ops, sampler  # noqa: F401

sys.modules[__name__ + '.ops'] = ops
sys.modules[__name__ + '.sampler'] = sampler
