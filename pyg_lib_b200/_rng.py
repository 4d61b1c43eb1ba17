"""Read / write torch's default CPU mt19937 engine as the C ABI's `pygb200_mt19937` struct from Python
(used by the ctypes-level multi-GPU orchestration; the single-GPU ops do this in C++, sampler_op.cpp).

Layout of torch.get_rng_state() (CPUGeneratorImplStateLegacy, ATen/CPUGeneratorImpl.h):
u64 seed | i32 left | i32 seeded | u64 next | u64 state[624] | normal-distribution cache."""
import ctypes as C

import numpy as np
import torch

_OFF_LEFT, _OFF_NEXT, _OFF_STATE = 8, 16, 24


class MT19937(C.Structure):
    _fields_ = [('state', C.c_uint32 * 624), ('left', C.c_int32), ('next', C.c_int32)]


def read_default_cpu_engine() -> MT19937:
    raw = torch.get_rng_state().numpy()
    mt = MT19937()
    st = np.frombuffer(raw[_OFF_STATE:_OFF_STATE + 624 * 8].tobytes(), dtype=np.uint64).astype(np.uint32)
    C.memmove(mt.state, st.ctypes.data, 624 * 4)
    mt.left = int(np.frombuffer(raw[_OFF_LEFT:_OFF_LEFT + 4].tobytes(), dtype=np.int32)[0])
    mt.next = int(np.frombuffer(raw[_OFF_NEXT:_OFF_NEXT + 8].tobytes(), dtype=np.uint64)[0])
    return mt


def write_default_cpu_engine(mt: MT19937) -> None:
    raw = torch.get_rng_state().clone()
    arr = raw.numpy()
    st = np.ctypeslib.as_array(mt.state).astype(np.uint64)
    arr[_OFF_STATE:_OFF_STATE + 624 * 8] = np.frombuffer(st.tobytes(), dtype=np.uint8)
    arr[_OFF_LEFT:_OFF_LEFT + 4] = np.frombuffer(np.int32(mt.left).tobytes(), dtype=np.uint8)
    arr[_OFF_NEXT:_OFF_NEXT + 8] = np.frombuffer(np.uint64(mt.next).tobytes(), dtype=np.uint8)
    torch.set_rng_state(raw)
