#!/usr/bin/env python
"""Small single-GPU runs for compute-sanitizer (memcheck / racecheck): every kernel family on grids that finish under
the tool -- fused passes, the cross-cycle pass, the 16-CTA cluster kernel (DSMEM), the single-CTA coarse kernel, the
operator engine, the other smoothers, PCG, and the multi-GPU kernel flavours on slabs inside one device
(tests/test_gpu_slab_kernels.py).  Results are compared with the oracle as usual: the tool must report 0 errors AND the
numbers must still be right.   compute-sanitizer --tool memcheck python tools/sanitizer_cases.py [quick]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cpu_checkers as cc  # noqa: E402
import pmg_b200 as pmg  # noqa: E402

quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
orc = cc.load("orc")
ok = True
cases = [(257, pmg.V, 1), (257, pmg.W, 2), (129, pmg.V, 1)] if not quick else [(257, pmg.V, 1), (129, pmg.W, 2)]
for n, kind, gamma in cases:
    f = cc.random_rhs(n, seed=5)
    want = np.zeros((n, n))
    for _ in range(2):
        orc.cycle(want, f, kind=cc.W if kind == pmg.W else cc.V, omega=2.0 / 3.0, eps=0.0, alpha=gamma)
    with pmg.Solver(n, omega=2.0 / 3.0, gamma=gamma) as s:
        s.set_rhs(f)
        s.zero_guess()
        for _ in range(2):
            s.cycle(kind)
        same = np.array_equal(s.get_solution(), want)
        s.zero_guess()
        k, hist = s.solve(kind, rel_tol=1e-8, max_cycles=40)   # cross-cycle path for V
        conv = hist[-1] < 1e-8 * hist[0]
        print("n=%d kind=%d: 2 cycles bit-identical=%s, solve %d cycles converged=%s cluster_top=%d" % (n, kind, same, k, conv, s.cluster_top))
        ok &= same and bool(conv)
if not quick:
    n = 65
    f = cc.random_rhs(n, seed=6)
    for sm in (pmg.SMOOTHER_RBGS, pmg.SMOOTHER_GS_LEX, pmg.SMOOTHER_CHEBYSHEV):
        with pmg.Solver(n, smoother=sm, prolong_mode=pmg.PROLONG_FULL) as s:
            s.set_rhs(f)
            s.zero_guess()
            s.cycle(pmg.V)
    with pmg.Solver(n, omega=2.0 / 3.0, prolong_mode=pmg.PROLONG_FULL) as s:
        s.set_rhs(f)
        s.zero_guess()
        k, hist = s.pcg(precond=1, rel_tol=1e-8, max_iter=30)
        print("pcg n=%d: %d steps" % (n, k))
        ok &= hist[-1] < 1e-8 * hist[0]
    import pytest
    rc = pytest.main(["-q", "-x", os.path.join(ROOT, "tests", "test_gpu_slab_kernels.py"), "-m", "gpu", "-k", "129 or 65 or 257"])
    ok &= (rc == 0 or rc == 5)
print("sanitizer_cases:", "OK" if ok else "FAILED")
sys.exit(0 if ok else 1)
