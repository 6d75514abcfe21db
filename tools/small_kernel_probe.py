#!/usr/bin/env python
"""Per-level cost of the single-CTA small-level kernel: times one V- and one W(gamma = 2)-visit for every top size
and generation, and derives own(n) from T(n, gamma) = own(n) + gamma * T(n_coarser, gamma)."""
import ctypes
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pmg_b200 as pmg  # noqa: E402

L = pmg.lib()
L.pmg_bench_small.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_double)]
for version in (2, 3):
    for gamma in (1, 2):
        t = {}
        for n in (5, 9, 17, 33, 65) + ((129, 257) if version == 3 else ()):  # 129 / 257: the cluster kernel
            v = ctypes.c_double()
            pmg.check(L.pmg_bench_small(n, gamma, 200, version, ctypes.byref(v)))
            t[n] = v.value
        own = {n: t[n] - (gamma * t[(n - 1) // 2 + 1] if n > 5 else 0.0) for n in t}
        print(json.dumps({"version": version, "gamma": gamma, "us_per_launch": {k: round(v, 2) for k, v in t.items()},
                          "own_us_per_visit_incl_launch_and_io_at_top": {k: round(v, 2) for k, v in own.items()}}), flush=True)

# per-level cycle profile inside one launch (thread 0 of CTA 0): own cycles of all visits of each level
L.pmg_coarse_profile.argtypes = [ctypes.POINTER(ctypes.c_longlong)]
for n, gamma in ((65, 1), (65, 2), (129, 1), (129, 2), (257, 1), (257, 2)):
    v = ctypes.c_double()
    pmg.check(L.pmg_bench_small(n, gamma, 3, 3, ctypes.byref(v)))
    prof = (ctypes.c_longlong * 16)()
    pmg.check(L.pmg_coarse_profile(prof))
    visits = lambda k: gamma ** (int(n - 1).bit_length() - 1 - k)
    print(json.dumps({"top": n, "gamma": gamma, "kernel_cycles": prof[0],
                      "own_cycles_per_level": {str(2 ** k + 1): prof[k] for k in range(1, 9) if prof[k]},
                      "cycles_per_visit": {str(2 ** k + 1): round(prof[k] / visits(k)) for k in range(1, 9) if prof[k]}}), flush=True)
