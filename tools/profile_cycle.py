#!/usr/bin/env python
"""Small driver for ncu: `cycle` runs 1 warm-up + 2 timed V(2,2) cycles at N (graphs off so that every
kernel is its own launch); `passes` runs the two level-0 fused passes alone (1 warm-up + 2 launches each).
`wcycle` is `cycle` with W(gamma = 2) recursion; `solve` runs pmg_solve for 4 cycles with graphs off (the cross-cycle path:
Pass A once, then per cycle the coarse part, k_cross and the convergence kernel).
`fcycle` runs 1 warm-up + 2 timed F-cycle passes (the runner's wrapper: restrict phi to the coarsest grid, nested iteration up).
Usage: python tools/profile_cycle.py {cycle|wcycle|fcycle|solve|passes} [N]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pmg_b200 as pmg  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "cycle"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 16385
s = pmg.Solver(n, omega=2.0 / 3.0, use_graph=0, gamma=2 if mode == "wcycle" else 1)
s.set_rhs_sine()
s.zero_guess()
if mode == "solve":
    s.close()
    s = pmg.Solver(n, omega=2.0 / 3.0)
    s.set_rhs_sine()
    for _ in range(2):
        s.zero_guess()
        k, hist = s.solve(pmg.V, rel_tol=0.0, max_cycles=4)
    print("N=%d solve of %d cycles: %.4f ms -> %.4f ms per cycle; norms %s" % (n, k, s.last_ms, s.last_ms / k, list(hist)))
elif mode == "fcycle":
    s.cycle(pmg.F)
    t = []
    for _ in range(2):
        s.zero_guess()
        s.cycle(pmg.F)
        t.append(s.last_ms)
    print("N=%d F-cycle pass ms: %s" % (n, t))
elif mode in ("cycle", "wcycle"):
    kind = pmg.W if mode == "wcycle" else pmg.V
    s.cycle(kind)
    t = []
    for _ in range(2):
        s.cycle(kind)
        t.append(s.last_ms)
    print("N=%d cycle ms (no graph): %s" % (n, t))
else:
    print("N=%d passA ms %.4f  passB+norm ms %.4f" % (n, s.bench_pass(0, 0, 2), s.bench_pass(1, 0, 2)))
s.close()
