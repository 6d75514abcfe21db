#!/usr/bin/env python
"""Driver for ncu: W(gamma = 2) cycles at N = 129 with the cluster kernel off, so that every cycle launches
k_coarse_local<65> with gamma = 2 (the subtree a W-cycle at N = 16385 visits 256 times).  PMG_CLUSTER=0 python tools/profile_coarse_w.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pmg_b200 as pmg  # noqa: E402

s = pmg.Solver(129, omega=2.0 / 3.0, use_graph=0, gamma=2)
s.set_rhs_sine()
s.zero_guess()
for _ in range(3):
    s.cycle(pmg.W)
print("W cycle at N=129: %.4f ms" % s.last_ms)
s.close()
