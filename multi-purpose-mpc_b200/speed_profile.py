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


def solve_speed_profiles(li_list, v_max_list, v_min, a_min, a_max, cfg=None, return_info=False):
    """T tracks at once (one CTA per track, mpc_speed_profile_batch): li_list[t] has n_t - 1 segment lengths,
    v_max_list[t] has n_t speed bounds.  Returns the list of v arrays (and iteration counts / statuses)."""
    L = _lib.load()
    T = len(v_max_list)
    off = np.zeros(T + 1, np.int32)
    off[1:] = np.cumsum([len(v) for v in v_max_list])
    li = np.zeros(int(off[-1]))
    vm = np.zeros(int(off[-1]))
    for t in range(T):
        n = len(v_max_list[t])
        assert len(li_list[t]) >= n - 1
        li[off[t]:off[t] + n - 1] = np.asarray(li_list[t], dtype=np.float64)[:n - 1]
        vm[off[t]:off[t + 1]] = v_max_list[t]
    v = np.empty(int(off[-1]))
    it = np.zeros(T, np.int32)
    st = np.zeros(T, np.int32)
    _lib._check(L.mpc_speed_profile_batch(_lib._dp(li), _lib._dp(vm), _lib._ip(off), T, float(v_min), float(a_min),
                                          float(a_max), None if cfg is None else C.byref(cfg), _lib._dp(v), _lib._ip(it),
                                          _lib._ip(st)))
    out = [v[off[t]:off[t + 1]].copy() for t in range(T)]
    if return_info:
        return out, it, st
    return out
