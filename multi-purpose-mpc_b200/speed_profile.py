"""Host wrapper of the device speed-profile QP (csrc/speed_profile.cu; reference rp.py:289-354)."""
import ctypes as C

import numpy as np

from . import _lib


def solve_speed_profile(li, v_max, v_min, a_min, a_max, cfg=None, return_info=False):
    L = _lib.load()
    li = np.ascontiguousarray(li, dtype=np.float64)
    v_max = np.ascontiguousarray(v_max, dtype=np.float64)
    n = len(v_max)
    assert len(li) >= n - 1
    v = np.empty(n)
    it, st = C.c_int32(0), C.c_int32(0)
    _lib._check(L.mpc_speed_profile(_lib._dp(li), _lib._dp(v_max), n, float(v_min), float(a_min), float(a_max),
                                    None if cfg is None else C.byref(cfg), _lib._dp(v), C.byref(it), C.byref(st)))
    if return_info:
        return v, it.value, st.value
    return v
